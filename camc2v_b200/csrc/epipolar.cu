// Camera-geometry kernels.
//   epipolar mask   CamContextI2V.get_epipolar_mask, R/model/camcontexti2v.py:202-271 (bit-exact, boolean)
//   Pluecker / ray  CameraControlLVDM.ray_condition, R/model/base.py:112-174
// The hot path never materialises the mask (attn_fa.cu reads the packed per-sample mask or evaluates the predicate in-tile); this kernel exists for the
// reference-facing API (`sample_locs_dict`) and for the bit-exactness tests.  It writes one byte per pair:
// pure HBM-write bound (268 MB per sample at 32x32x16), 16 mask bytes per thread, 512 B per warp store.
#include "common.cuh"
#include "kernels.h"

namespace c2v {

__global__ void __launch_bounds__(256) epipolar_mask_kernel(const float* __restrict__ Fm, uint8_t* __restrict__ out, int T1, int T, int H, int W,
                                                            int d, float thr, float off) {
    // grid: x = query row (t1, i) of one batch element, y = batch.  Each warp sweeps key chunks of 16.
    // F is [B, T1, T, 3, 3]: T1 query frames x T key frames (T1 = T in the UNet; 16 x (1 + n) for the adaptor's conditional mask)
    const int HW = H * W;
    const int64_t L = (int64_t)T * HW;            // keys per row
    const int b = blockIdx.y;
    const int row = blockIdx.x;
    const int t1 = row / HW, i = row % HW;
    const float xi = __fadd_rn(__fmul_rn((float)(i % W), (float)d), off);
    const float yi = __fadd_rn(__fmul_rn((float)(i / W), (float)d), off);
    uint8_t* orow = out + ((size_t)b * T1 * HW + row) * L;
    const int chunks = (int)(L / 16);
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
        const int key0 = c * 16;
        int cur_t2 = -1;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int t2 = (key0 + e) / HW;            // constant over the chunk whenever HW is a multiple of 16
            if (t2 != cur_t2) {
                cur_t2 = t2;
                const float* f = Fm + (((size_t)b * T1 + t1) * T + t2) * 9;
                const float a0 = __fmaf_rn(f[2], 1.0f, __fmaf_rn(f[1], yi, __fmul_rn(f[0], xi)));
                const float a1 = __fmaf_rn(f[5], 1.0f, __fmaf_rn(f[4], yi, __fmul_rn(f[3], xi)));
                const float a2 = __fmaf_rn(f[8], 1.0f, __fmaf_rn(f[7], yi, __fmul_rn(f[6], xi)));
                const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
                l0 = __fdiv_rn(a0, nrm);
                l1 = __fdiv_rn(a1, nrm);
                l2 = __fdiv_rn(a2, nrm);
            }
            const int pj = key0 + e - t2 * HW;
            const float xj = __fadd_rn(__fmul_rn((float)(pj % W), (float)d), off);
            const float yj = __fadd_rn(__fmul_rn((float)(pj / W), (float)d), off);
            const float dist = fabsf(__fmaf_rn(l2, 1.0f, __fmaf_rn(l1, yj, __fmul_rn(l0, xj))));
            if (dist < thr) w[e >> 2] |= 1u << ((e & 3) * 8);
        }
        *reinterpret_cast<uint4*>(orow + key0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

int epipolar_mask_launch(const float* F, uint8_t* out, int B, int T1, int T, int H, int W, int d, cudaStream_t st) {
    const int HW = H * W;
    if (((int64_t)T * HW) % 16 != 0 || B <= 0 || B > 65535 || T1 <= 0) return ERR_UNSUPPORTED;
    const float thr = (float)((double)d * sqrt(2.0) / 2.0);
    const float off = (float)d / 2.0f - 0.5f;
    epipolar_mask_kernel<<<dim3(T1 * HW, B), 256, 0, st>>>(F, out, T1, T, H, W, d, thr, off);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

__global__ void __launch_bounds__(256) plucker_kernel(const float* __restrict__ K, const float* __restrict__ c2w, float* __restrict__ out, int T,
                                                      int H, int W, int plucker) {
    const int bt = blockIdx.y;                  // b*T + t
    const int b = bt / T, t = bt % T;
    const float* k = K + (size_t)bt * 9;
    const float* m = c2w + (size_t)bt * 16;
    const float fx = k[0], fy = k[4], cx = k[2], cy = k[5];
    const float ox = m[3], oy = m[7], oz = m[11];
    const int HW = H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const int x = p % W, y = p / W;
        float dx = __fdiv_rn((float)x + 0.5f - cx, fx);
        float dy = __fdiv_rn((float)y + 0.5f - cy, fy);
        float dz = 1.0f;
        const float n = __fsqrt_rn(dx * dx + dy * dy + dz * dz);
        dx = __fdiv_rn(dx, n); dy = __fdiv_rn(dy, n); dz = __fdiv_rn(dz, n);
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
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) out[(((size_t)b * 6 + ch) * T + t) * HW + p] = c[ch];
    }
}

int plucker_launch(const float* K, const float* c2w, float* out, int B, int T, int H, int W, int plucker, cudaStream_t st) {
    if (B * T > 65535) return ERR_UNSUPPORTED;
    int gx = (H * W + 255) / 256;
    if (gx > 64) gx = 64;
    plucker_kernel<<<dim3(gx, B * T), 256, 0, st>>>(K, c2w, out, T, H, W, plucker);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
