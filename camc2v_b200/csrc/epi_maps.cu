// Per-sample derived forms of the epipolar mask for the attention kernel (attn_fa.cu): the packed 1-bit mask and the
// (128-query x 64-key)-tile occupancy map with its longest-first CTA order.  The reference materialises a bool mask [B, L, L]
// (268 MB per sample at 32x32x16, R/model/camcontexti2v.py:202-271) and reads it in every layer; here the predicate is evaluated
// once per sample with the reference's exact fp32 operation order (camcontexti2v.py:229-239; FMA-chain contraction, separately
// rounded norm, IEEE sqrt / div) into 1 bit per pair.
#include "attn.h"
#include "common.cuh"

namespace c2v {

constexpr int AT_BM = 128;   // query rows per CTA
constexpr int AT_BN = 64;    // keys per tile
struct EpiLine {
    float l0, l1, l2;
};

// Normalised epipolar line of query pixel (xi, yi) in frame t2 (camcontexti2v.py:229-236).
__device__ __forceinline__ EpiLine epi_line(const float* __restrict__ f, float xi, float yi) {
    float a0 = __fmaf_rn(f[2], 1.0f, __fmaf_rn(f[1], yi, __fmul_rn(f[0], xi)));
    float a1 = __fmaf_rn(f[5], 1.0f, __fmaf_rn(f[4], yi, __fmul_rn(f[3], xi)));
    float a2 = __fmaf_rn(f[8], 1.0f, __fmaf_rn(f[7], yi, __fmul_rn(f[6], xi)));
    const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
    EpiLine l;
    l.l0 = __fdiv_rn(a0, nrm);
    l.l1 = __fdiv_rn(a1, nrm);
    l.l2 = __fdiv_rn(a2, nrm);
    return l;
}

// ------------------------------------------------------------------------------------------------
// Packed epipolar mask: out[b][q_tile][k_chunk][r] bit i = mask[b][128 q_tile + r][32 k_chunk + i], evaluated with exactly the
// arithmetic of the attention kernel's in-kernel predicate (attn_fa.cu MODE 2; and hence of the reference, camcontexti2v.py:229-239).  One CTA = 128 queries x
// one key frame: the normalised line is computed once per (query, frame), then 32 predicates per word.
// ------------------------------------------------------------------------------------------------
template <int LOGW, int D>
__global__ void __launch_bounds__(128) epi_bitmask_kernel(const float* __restrict__ Fm, unsigned int* __restrict__ out, int T, float thr) {
    constexpr int W = 1 << LOGW, HW = W * W, RPC = 32 / W, CPF = HW / 32;      // chunks per frame
    constexpr float DF = (float)D, OFFC = (float)D * 0.5f - 0.5f;
    const int t2 = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
    const int r = threadIdx.x;
    const int qi = qt * AT_BM + r;
    const int t1 = qi >> (2 * LOGW), pix = qi & (HW - 1);
    const float xi = (float)(pix & (W - 1)) * DF + OFFC, yi = (float)(pix >> LOGW) * DF + OFFC;
    const EpiLine line = epi_line(Fm + (((size_t)b * T + t1) * T + t2) * 9, xi, yi);
    float l0x[W];
#pragma unroll
    for (int x = 0; x < W; ++x) l0x[x] = __fmul_rn(line.l0, (float)x * DF + OFFC);
    const size_t n_chunks = (size_t)T * CPF;
    unsigned int* o = out + (((size_t)b * gridDim.y + qt) * n_chunks + (size_t)t2 * CPF) * AT_BM + r;
    for (int c = 0; c < CPF; ++c) {
        unsigned int word = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float yr = (float)(c * RPC + (i >> LOGW)) * DF + OFFC;
            const float wv = __fadd_rn(__fmaf_rn(line.l1, yr, l0x[i & (W - 1)]), line.l2);
            word |= (fabsf(wv) < thr ? 1u : 0u) << i;
        }
        o[(size_t)c * AT_BM] = word;
    }
}

// Any grid (e.g. the 4x4 level at d = 64, where a 32-key chunk spans two frames): one thread per (query, 32-key chunk), the
// predicate in the general form of the reference (pixel centres x*d + d/2 - 0.5, camcontexti2v.py:213-239).  Once per sample.
__global__ void __launch_bounds__(128) epi_bitmask_generic_kernel(const float* __restrict__ Fm, unsigned int* __restrict__ out, int T, int H, int W,
                                                                  int d, float thr, float off) {
    const int HW = H * W, L = T * HW;
    const int c = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
    const int r = threadIdx.x;
    const int qi = qt * AT_BM + r;                   // L % 128 == 0 (checked by the launcher)
    const int t1 = qi / HW, pix = qi % HW;
    const float xi = __fadd_rn(__fmul_rn((float)(pix % W), (float)d), off), yi = __fadd_rn(__fmul_rn((float)(pix / W), (float)d), off);
    const float* Frow = Fm + ((size_t)b * T + t1) * T * 9;
    int cur_t2 = -1;
    EpiLine line = {0.f, 0.f, 0.f};
    unsigned int word = 0;
    for (int i = 0; i < 32; ++i) {
        const int key = c * 32 + i;
        if (key >= L) break;
        const int t2 = key / HW;
        if (t2 != cur_t2) {
            cur_t2 = t2;
            line = epi_line(Frow + t2 * 9, xi, yi);
        }
        const int pj = key - t2 * HW;
        const float xj = __fadd_rn(__fmul_rn((float)(pj % W), (float)d), off);
        const float yj = __fadd_rn(__fmul_rn((float)(pj / W), (float)d), off);
        const float dist = fabsf(__fadd_rn(__fmaf_rn(line.l1, yj, __fmul_rn(line.l0, xj)), line.l2));
        word |= (dist < thr ? 1u : 0u) << i;
    }
    out[(((size_t)b * gridDim.y + qt) * gridDim.x + c) * AT_BM + r] = word;
}

int epi_bitmask_launch(const float* F, unsigned int* out, int B, int T, int H, int W, int d, cudaStream_t st) {
    const int L = T * H * W;
    if (L % AT_BM != 0 || L % 32 != 0 || B > 65535) return ERR_UNSUPPORTED;
    const float thr = (float)((double)d * sqrt(2.0) / 2.0);
    dim3 grid(T, L / AT_BM, B);
#define C2V_BM(LW, DD) epi_bitmask_kernel<LW, DD><<<grid, 128, 0, st>>>(F, out, T, thr)
    if (H == W && W == 32 && d == 8) C2V_BM(5, 8);
    else if (H == W && W == 16 && d == 16) C2V_BM(4, 16);
    else if (H == W && W == 8 && d == 32) C2V_BM(3, 32);
    else if (H == W && W == 16 && d == 8) C2V_BM(4, 8);
    else if (H == W && W == 8 && d == 16) C2V_BM(3, 16);
    else {
        if (L / AT_BM > 65535) return ERR_UNSUPPORTED;
        epi_bitmask_generic_kernel<<<dim3(L / 32, L / AT_BM, B), 128, 0, st>>>(F, out, T, H, W, d, thr, (float)d / 2.0f - 0.5f);
    }
#undef C2V_BM
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// Epipolar tile map: bit (q_tile, k_tile) = "some query of the 128-query tile may see some key of the 64-key tile".
// Same conservative per-image-row interval test (and the same rounding margin) as the in-tile row skip of attn_fa_kernel,
// so a cleared bit implies every chunk of that tile would have been skipped anyway: results are bit-identical with and
// without the map.  F is constant over the 25 steps x 2 passes x 16 layers of a sample, so the map is built once per sample.
// ------------------------------------------------------------------------------------------------
template <int LOGW, int D>
__global__ void __launch_bounds__(128) epi_tile_map_kernel(const float* __restrict__ Fm, unsigned int* __restrict__ map, int T, int n_ktiles,
                                                           int words, float thr) {
    constexpr int W = 1 << LOGW, HW = W * W;
    constexpr float DF = (float)D, OFFC = (float)D * 0.5f - 0.5f;
    const int b = blockIdx.y, qt = blockIdx.x;
    const int L = T * HW;
    const int qi = min(qt * AT_BM + (int)threadIdx.x, L - 1);
    const int t1 = qi >> (2 * LOGW), pix = qi & (HW - 1);
    const float xi = (float)(pix & (W - 1)) * DF + OFFC, yi = (float)(pix >> LOGW) * DF + OFFC;
    const float* Frow = Fm + ((size_t)b * T + t1) * T * 9;
    unsigned int* out = map + ((size_t)b * gridDim.x + qt) * (words + 1);     // last word of a row: LPT order (epi_tile_order_kernel)
    int cur_t2 = -1;
    EpiLine line = {0.f, 0.f, 0.f};
    float thr_m = 0.f;
    unsigned int word = 0;
    for (int j = 0; j < n_ktiles; ++j) {
        bool maybe = false;
        for (int rr = 0; rr < AT_BN / W; ++rr) {
            const int key0 = j * AT_BN + rr * W;
            if (key0 >= L) break;
            const int t2 = key0 >> (2 * LOGW);
            if (t2 != cur_t2) {
                cur_t2 = t2;
                line = epi_line(Frow + t2 * 9, xi, yi);
                const float cmax = (float)(W - 1) * DF + OFFC;
                thr_m = thr + 1e-6f + 4e-7f * (fabsf(line.l0) * cmax + fabsf(line.l1) * cmax + fabsf(line.l2));
            }
            const float yr = (float)((key0 & (HW - 1)) >> LOGW) * DF + OFFC;
            const float w0 = __fadd_rn(__fmaf_rn(line.l1, yr, __fmul_rn(line.l0, OFFC)), line.l2);
            const float w1 = __fadd_rn(__fmaf_rn(line.l1, yr, __fmul_rn(line.l0, (float)(W - 1) * DF + OFFC)), line.l2);
            maybe |= !((w0 > thr_m && w1 > thr_m) || (w0 < -thr_m && w1 < -thr_m));
        }
        const int any = __syncthreads_or(maybe ? 1 : 0);
        if (threadIdx.x == 0) {
            if (any) word |= 1u << (j & 31);
            if ((j & 31) == 31 || j == n_ktiles - 1) {
                out[j >> 5] = word;
                word = 0;
            }
        }
    }
}

// order[rank] = query tile with the rank-th largest number of visited key tiles (ties by index); stored in the extra word of
// row `rank` of the map.  One CTA per batch element, rank sort (n <= 1024 query tiles).
__global__ void __launch_bounds__(1024) epi_tile_order_kernel(unsigned int* __restrict__ map, int nq, int words) {
    __shared__ int cnt[1024];
    unsigned int* m = map + (size_t)blockIdx.x * nq * (words + 1);
    const int i = threadIdx.x;
    if (i < nq) {
        int c = 0;
        for (int w = 0; w < words; ++w) c += __popc(m[(size_t)i * (words + 1) + w]);
        cnt[i] = c;
    }
    __syncthreads();
    if (i < nq) {
        int rank = 0;
        for (int j = 0; j < nq; ++j) rank += (cnt[j] > cnt[i]) || (cnt[j] == cnt[i] && j < i);
        m[(size_t)rank * (words + 1) + words] = (unsigned int)i;
    }
}

int epi_tile_map_launch(const float* F, unsigned int* map, int B, int T, int H, int W, int d, cudaStream_t st) {
    if (H != W) return ERR_UNSUPPORTED;
    const int L = T * H * W;
    const int nq = (L + AT_BM - 1) / AT_BM, nk = (L + AT_BN - 1) / AT_BN, words = (nk + 31) / 32;
    const float thr = (float)((double)d * sqrt(2.0) / 2.0);
    dim3 grid(nq, B);
#define C2V_MAP(LW, DD) epi_tile_map_kernel<LW, DD><<<grid, 128, 0, st>>>(F, map, T, nk, words, thr)
    if (W == 32 && d == 8) C2V_MAP(5, 8);
    else if (W == 16 && d == 16) C2V_MAP(4, 16);
    else if (W == 8 && d == 32) C2V_MAP(3, 32);
    else if (W == 16 && d == 8) C2V_MAP(4, 8);
    else if (W == 8 && d == 16) C2V_MAP(3, 16);
    else return ERR_UNSUPPORTED;
#undef C2V_MAP
    if (nq > 1024) return ERR_UNSUPPORTED;
    epi_tile_order_kernel<<<B, 1024, 0, st>>>(map, nq, words);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
