/*
 * CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's epipolar mask and Pluecker/ray embedding:
 *   - CamContextI2V.get_epipolar_mask   R/model/camcontexti2v.py:202-271 (pix2coord: R/model/modules/epipolar.py:32-34)
 *   - CameraControlLVDM.ray_condition   R/model/base.py:112-174
 *
 * The mask is boolean and must be reproduced bit-exactly.  The reference evaluates it with fp32
 * torch ops on CPU: two 3-term contractions (F @ grid^T and lines^T @ grid^T), a 2-norm, a divide, abs
 * and a compare against float32(d*sqrt(2)/2).  Executed in the build container the reference's
 * contractions round like a k-ordered fused chain  acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc)
 * (re-verified by oracle/refgen/make_golden.py against the reference itself),
 * the norm is sqrt(l0*l0 + l1*l1) with every operation rounded separately, and the divide is IEEE.
 * fmaf() below is correctly rounded on every platform, so this file gives the same bits on any host.
 *
 * Compile with -O2 -ffp-contract=off (no implicit contraction: every FMA here is written explicitly).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline float dot3_chain(float a0, float b0, float a1, float b1, float a2, float b2) {
    float acc = a0 * b0;
    acc = fmaf(a1, b1, acc);
    acc = fmaf(a2, b2, acc);
    return acc;
}

/* Fm: [B, T1, T2, 3, 3] fp32.  out: [B, T1*HW, T2*HW] uint8 (1 = attend), rows = (t1, pixel i), cols = (t2, pixel j).
 * T1 = T2 for the UNet's temporal blocks; T1 = 16 target frames, T2 = 1 + n context frames for the adaptor's conditional mask
 * (compute_conditional_epipolar_mask, R/model/camcontexti2v.py:493-521). */
void epi_mask_oracle_rect(const float *Fm, int B, int T1, int T2, int H, int W, int d, uint8_t *out) {
    const int HW = H * W;
    const size_t L2 = (size_t)T2 * HW;
    const float thr = (float)((double)d * sqrt(2.0) / 2.0);
    const float off = (float)d / 2.0f - 0.5f;
    for (int b = 0; b < B; ++b)
        for (int t1 = 0; t1 < T1; ++t1)
            for (int t2 = 0; t2 < T2; ++t2) {
                const float *f = Fm + (((size_t)b * T1 + t1) * T2 + t2) * 9;
                for (int i = 0; i < HW; ++i) {
                    const float xi = (float)(i % W) * (float)d + off;
                    const float yi = (float)(i / W) * (float)d + off;
                    float l0 = dot3_chain(f[0], xi, f[1], yi, f[2], 1.0f);
                    float l1 = dot3_chain(f[3], xi, f[4], yi, f[5], 1.0f);
                    float l2 = dot3_chain(f[6], xi, f[7], yi, f[8], 1.0f);
                    const float s0 = l0 * l0;
                    const float s1 = l1 * l1;
                    const float nrm = sqrtf(s0 + s1);
                    l0 = l0 / nrm;
                    l1 = l1 / nrm;
                    l2 = l2 / nrm;
                    uint8_t *row = out + (((size_t)b * T1 + t1) * HW + i) * L2 + (size_t)t2 * HW;
                    for (int j = 0; j < HW; ++j) {
                        const float xj = (float)(j % W) * (float)d + off;
                        const float yj = (float)(j / W) * (float)d + off;
                        const float dist = fabsf(dot3_chain(l0, xj, l1, yj, l2, 1.0f));
                        row[j] = dist < thr;   /* NaN (degenerate line) compares false, as in torch */
                    }
                }
            }
}

void epi_mask_oracle(const float *Fm, int B, int T, int H, int W, int d, uint8_t *out) { epi_mask_oracle_rect(Fm, B, T, T, H, W, d, out); }

/*
 * K: [B, T, 3, 3], c2w: [B, T, 4, 4] -> out [B, 6, T, H, W].
 * plucker != 0: channels = [o x d, d]; plucker == 0 ("ray", CamI2V): channels = [o, d].
 * Floating point, tolerance-checked (not bit-exact): d = normalize(((i+.5-cx)/fx, (j+.5-cy)/fy, 1)) @ R^T.
 */
void plucker_oracle(const float *K, const float *c2w, int B, int T, int H, int W, int plucker, float *out) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) {
            const float *k = K + ((size_t)b * T + t) * 9;
            const float *m = c2w + ((size_t)b * T + t) * 16;
            const float fx = k[0], fy = k[4], cx = k[2], cy = k[5];
            const float ox = m[3], oy = m[7], oz = m[11];
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float dx = ((float)x + 0.5f - cx) / fx;
                    float dy = ((float)y + 0.5f - cy) / fy;
                    float dz = 1.0f;
                    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
                    dx /= n; dy /= n; dz /= n;
                    const float rx = dx * m[0] + dy * m[1] + dz * m[2];
                    const float ry = dx * m[4] + dy * m[5] + dz * m[6];
                    const float rz = dx * m[8] + dy * m[9] + dz * m[10];
                    float c[6];
                    if (plucker) {
                        c[0] = oy * rz - oz * ry;
                        c[1] = oz * rx - ox * rz;
                        c[2] = ox * ry - oy * rx;
                    } else {
                        c[0] = ox; c[1] = oy; c[2] = oz;
                    }
                    c[3] = rx; c[4] = ry; c[5] = rz;
                    for (int ch = 0; ch < 6; ++ch)
                        out[(((size_t)b * 6 + ch) * T + t) * HW + (size_t)y * W + x] = c[ch];
                }
        }
}
