// Layout changes, small glue ops and the fused CFG + DDIM update.  All HBM-bound: vectorised, coalesced,
// one read and one write per element.
#include "common.cuh"
#include "kernels.h"

namespace c2v {

static inline int grid_for(int64_t work, int threads, int cap = 148 * 16) {
    int64_t g = (work + threads - 1) / threads;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
// [B, C, S] fp32  <->  channels-last [B, S, Cpad]     (b c t h w <-> (b t) (h w) c of modified_forwards.py:48,130)
// ------------------------------------------------------------------------------------------------
template <bool OUT_BF16>
__global__ void to_cl_kernel(const float* __restrict__ in, void* __restrict__ out, int C, int S, int Cpad) {
    pdl_entry();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, s = s0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && s < S) ? in[((size_t)b * C + c) * S + s] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int s = s0 + i, c = c0 + threadIdx.x;
        if (s < S && c < Cpad) {
            const float v = tile[threadIdx.x][i];
            const size_t o = ((size_t)b * S + s) * Cpad + c;
            if (OUT_BF16)
                reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16(v);
            else
                reinterpret_cast<float*>(out)[o] = v;
        }
    }
}

__global__ void from_cl_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int S) {
    pdl_entry();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int s = s0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && s < S) ? in[((size_t)b * S + s) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, s = s0 + threadIdx.x;
        if (c < C && s < S) out[((size_t)b * C + c) * S + s] = tile[threadIdx.x][i];
    }
}

int to_channels_last_launch(const float* in, void* out, int B, int C, int S, int Cpad, int out_bf16, cudaStream_t st) {
    if (Cpad < C || B <= 0 || B > 65535) return ERR_BAD_ARG;
    dim3 grid((S + 31) / 32, (Cpad + 31) / 32, B), block(32, 8);
    if (out_bf16)
        C2V_CHECK_CUDA(launch(to_cl_kernel<true>, grid, block, 0, st, in, out, C, S, Cpad));
    else
        C2V_CHECK_CUDA(launch(to_cl_kernel<false>, grid, block, 0, st, in, out, C, S, Cpad));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

int from_channels_last_launch(const float* in, float* out, int B, int C, int S, cudaStream_t st) {
    if (B <= 0 || B > 65535) return ERR_BAD_ARG;
    dim3 grid((S + 31) / 32, (C + 31) / 32, B), block(32, 8);
    C2V_CHECK_CUDA(launch(from_cl_kernel, grid, block, 0, st, in, out, C, S));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// channel concat of the skip connection (modified_forwards.py:108), optional bf16 copy for the 1x1 skip conv
// ------------------------------------------------------------------------------------------------
// 16-bit copies of UNNORMALISED fp32 data (the residual stream feeding the 1x1 skip convolution): scaled by a power of two chosen by
// the caller (the consumer's weights carry the inverse, so the product is unchanged) and, in the IEEE-half build, saturated at the
// largest finite half instead of overflowing to inf.
__device__ __forceinline__ float sat16(float x) {
#if C2V_OPERAND_IS_FP16
    return fminf(fmaxf(x, -65504.0f), 65504.0f);
#else
    return x;
#endif
}

__global__ void concat_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ of, __nv_bfloat16* __restrict__ ob,
                              int64_t rows, int Ca, int Cb, float s16) {
    pdl_entry();
    const int nv = (Ca + Cb) >> 2, nva = Ca >> 2;
    const int64_t total = rows * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nv;
        const int cv = (int)(i - r * nv);
        const float4 v = cv < nva ? *reinterpret_cast<const float4*>(a + r * Ca + cv * 4)
                                  : *reinterpret_cast<const float4*>(b + r * Cb + (cv - nva) * 4);
        if (of) *reinterpret_cast<float4*>(of + i * 4) = v;
        if (ob)
            *reinterpret_cast<uint2*>(ob + i * 4) =
                make_uint2(pack_bf16(sat16(v.x * s16), sat16(v.y * s16)), pack_bf16(sat16(v.z * s16), sat16(v.w * s16)));
    }
}

int concat_channels_launch(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, float scale16,
                           cudaStream_t st) {
    if (Ca % 4 || Cb % 4) return ERR_UNSUPPORTED;
    C2V_CHECK_CUDA(launch(concat_kernel, dim3(grid_for(rows * ((Ca + Cb) >> 2), 256)), dim3(256), 0, st, a, b, out_f32,
                          reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, Ca, Cb, scale16));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n4, int64_t n, float s) {
    pdl_entry();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(in + i * 4);
        *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(pack_bf16(sat16(v.x * s), sat16(v.y * s)), pack_bf16(sat16(v.z * s), sat16(v.w * s)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n4 * 4; i < n; ++i) out[i] = __float2bfloat16(sat16(in[i] * s));
}

int cast_bf16_launch(const float* in, void* out, int64_t n, float scale, cudaStream_t st) {
    C2V_CHECK_CUDA(launch(cast_bf16_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, st, in, reinterpret_cast<__nv_bfloat16*>(out), n / 4, n, scale));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// nearest 2x upsample feeding the Upsample conv (openaimodel3d.py:101-105)
__global__ void upsample2x_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H, int W, int C) {
    pdl_entry();
    const int nv = C >> 2;
    const int64_t total = (int64_t)N * 4 * H * W * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % nv);
        int64_t r = i / nv;
        const int xo = (int)(r % (2 * W)); r /= 2 * W;
        const int yo = (int)(r % (2 * H));
        const int n = (int)(r / (2 * H));
        const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + (yo >> 1)) * W + (xo >> 1)) * C + cv * 4);
        *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

int upsample2x_launch(const float* in, void* out, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 4) return ERR_UNSUPPORTED;
    C2V_CHECK_CUDA(launch(upsample2x_kernel, dim3(grid_for((int64_t)N * 4 * H * W * (C >> 2), 256)), dim3(256), 0, st, in, reinterpret_cast<__nv_bfloat16*>(out), N, H, W, C));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// im2col of the stride-2 Downsample conv (openaimodel3d.py:66-70): row (n, yo, xo), column tap*C + c
// pad_lo = 1: symmetric padding 1 (UNet Downsample, openaimodel3d.py:68-70); pad_lo = 0: zero padding on the right / bottom only
// (VAE encoder Downsample, ae_modules.py:102-106)
__global__ void im2col_s2_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H, int W, int C, int pad_lo) {
    pdl_entry();
    const int nv = C >> 2, Ho = H >> 1, Wo = W >> 1;
    const int64_t total = (int64_t)N * Ho * Wo * 9 * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % nv);
        int64_t r = i / nv;
        const int tap = (int)(r % 9); r /= 9;
        const int xo = (int)(r % Wo); r /= Wo;
        const int yo = (int)(r % Ho);
        const int n = (int)(r / Ho);
        const int y = 2 * yo + tap / 3 - pad_lo, x = 2 * xo + tap % 3 - pad_lo;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + y) * W + x) * C + cv * 4);
        *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

int im2col_s2_launch(const float* in, void* out, int N, int H, int W, int C, int pad_lo, cudaStream_t st) {
    if (C % 4 || H % 2 || W % 2) return ERR_UNSUPPORTED;
    C2V_CHECK_CUDA(launch(im2col_s2_kernel, dim3(grid_for((int64_t)N * (H / 2) * (W / 2) * 9 * (C >> 2), 256)), dim3(256), 0, st, in, reinterpret_cast<__nv_bfloat16*>(out), N, H, W, C, pad_lo));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int C, int64_t dst_bstride,
                                 int ldd) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * C; i += gridDim.x * blockDim.x) {
        const int r = i / C, c = i - r * C;
        dst[(size_t)b * dst_bstride + (size_t)r * ldd + c] = src[i];
    }
}

int copy_rows_launch(const void* src, void* dst, int rows, int C, int B, int64_t dst_bstride, int ldd, cudaStream_t st) {
    copy_rows_kernel<<<dim3(grid_for((int64_t)rows * C, 256, 64), B), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                                                  reinterpret_cast<__nv_bfloat16*>(dst), rows, C, dst_bstride, ldd);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// small-M linear: one warp per output feature, weights streamed once (bf16, 16-byte loads)
// ------------------------------------------------------------------------------------------------
constexpr int SK_MAXM = 8;
constexpr int SK_SMEM_FLOATS = 12288;   // 48 KB of staged activations per CTA
// Activations (after the optional SiLU) are staged once per CTA in shared memory; every warp then streams one weight row
// with all of its 16-byte loads in flight and reduces over K with shuffles.
__global__ void __launch_bounds__(256) skinny_linear_kernel(const float* __restrict__ in, const __nv_bfloat16* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out, int M, int N, int K,
                                                            int silu_in, int mb) {
    pdl_entry();
    extern __shared__ float xs[];           // [mb][K]
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (int m0 = 0; m0 < M; m0 += mb) {
        const int mc = min(mb, M - m0);
        __syncthreads();
        for (int i = threadIdx.x * 4; i < mc * K; i += blockDim.x * 4) {
            float4 x = *reinterpret_cast<const float4*>(in + (size_t)m0 * K + i);
            if (silu_in) {
                x.x = x.x / (1.0f + expf(-x.x));
                x.y = x.y / (1.0f + expf(-x.y));
                x.z = x.z / (1.0f + expf(-x.z));
                x.w = x.w / (1.0f + expf(-x.w));
            }
            *reinterpret_cast<float4*>(xs + i) = x;
        }
        __syncthreads();
        if (n >= N) continue;
        float acc[SK_MAXM];
#pragma unroll
        for (int m = 0; m < SK_MAXM; ++m) acc[m] = 0.f;
        const __nv_bfloat16* wr = w + (size_t)n * K;
#pragma unroll 4
        for (int k = lane * 8; k < K; k += 256) {
            const uint4 wv = *reinterpret_cast<const uint4*>(wr + k);
            const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
            float wf[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                wf[2 * e] = __low2float(wh[e]);
                wf[2 * e + 1] = __high2float(wh[e]);
            }
#pragma unroll
            for (int m = 0; m < SK_MAXM; ++m) {
                if (m < mc) {
                    const float4 x0 = *reinterpret_cast<const float4*>(xs + m * K + k);
                    const float4 x1 = *reinterpret_cast<const float4*>(xs + m * K + k + 4);
                    acc[m] = fmaf(x0.x, wf[0], acc[m]);
                    acc[m] = fmaf(x0.y, wf[1], acc[m]);
                    acc[m] = fmaf(x0.z, wf[2], acc[m]);
                    acc[m] = fmaf(x0.w, wf[3], acc[m]);
                    acc[m] = fmaf(x1.x, wf[4], acc[m]);
                    acc[m] = fmaf(x1.y, wf[5], acc[m]);
                    acc[m] = fmaf(x1.z, wf[6], acc[m]);
                    acc[m] = fmaf(x1.w, wf[7], acc[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < SK_MAXM; ++m) {
            if (m < mc) {                                   // warp-uniform
                const float s = warp_sum(acc[m]);
                if (lane == 0) out[(size_t)(m0 + m) * N + n] = s + (bias ? bias[n] : 0.f);
            }
        }
    }
}

int skinny_linear_launch(const float* in, const void* w, const float* bias, float* out, int M, int N, int K, int silu_in, cudaStream_t st) {
    if (K % 8 || K > SK_SMEM_FLOATS) return ERR_UNSUPPORTED;
    int mb = SK_SMEM_FLOATS / K;
    if (mb > SK_MAXM) mb = SK_MAXM;
    if (mb > M) mb = M;
    C2V_CHECK_CUDA(launch(skinny_linear_kernel, dim3((N + 7) / 8), dim3(256), (size_t)mb * K * sizeof(float), st, in,
                          reinterpret_cast<const __nv_bfloat16*>(w), bias, out, M, N, K, silu_in, mb));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// sinusoidal embedding (utils_diffusion.py:8-28): [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i / half)
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int n, int dim) {
    pdl_entry();
    const int half = dim >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * half; i += gridDim.x * blockDim.x) {
        const int r = i / half, c = i - r * half;
        const float f = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)c), (float)half));
        const float a = __fmul_rn((float)t[r], f);
        out[(size_t)r * dim + c] = cosf(a);
        out[(size_t)r * dim + half + c] = sinf(a);
    }
}

int timestep_embedding_launch(const int64_t* t, float* out, int n, int dim, cudaStream_t st) {
    if (dim % 2) return ERR_UNSUPPORTED;
    C2V_CHECK_CUDA(launch(timestep_embedding_kernel, dim3(grid_for((int64_t)n * (dim / 2), 128, 64)), dim3(128), 0, st, t, out, n, dim));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// Fused CFG combine + guidance rescale + DDIM update (ddim.py:262-346, utils_diffusion.py:147-158).
// One CTA per sample: the reference issues ~15 elementwise / reduction launches here.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

// model_output = e_u + scale (e_c - e_u) [+ cam_w (e_c - e_nc)]: the optional third term is the reference's camera guidance
// (ddim.py:268-280; e_nc = the conditional pass WITHOUT the camera condition, cam_w = (camera_cfg - 1) * scheduler weight)
__device__ __forceinline__ float cfg_combine(float c, float u, float scale, const float* __restrict__ enc, int64_t i, float cam_w) {
    float e = __fadd_rn(u, __fmul_rn(scale, __fsub_rn(c, u)));
    if (enc) e = __fadd_rn(e, __fmul_rn(cam_w, __fsub_rn(c, enc[i])));
    return e;
}

__global__ void __launch_bounds__(1024) cfg_ddim_kernel(const float* __restrict__ x, const float* __restrict__ ec, const float* __restrict__ eu,
                                                        const float* __restrict__ enc, const float* __restrict__ noise, float* __restrict__ x_prev,
                                                        float* __restrict__ pred_x0, int64_t n, float scale, float cam_w, float phi, float a_t,
                                                        float a_prev, float sigma_t, float sqrt_one_minus_at) {
    pdl_entry();
    __shared__ double sh[32];
    const size_t base = (size_t)blockIdx.x * n;
    x += base; ec += base; eu += base; noise += base; x_prev += base; pred_x0 += base;
    if (enc) enc += base;
    float ratio = 1.f;
    if (phi > 0.f) {
        double sc = 0.0, se = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const float c = ec[i], u = eu[i];
            sc += c;
            se += cfg_combine(c, u, scale, enc, i, cam_w);
        }
        const double mc = block_sum(sc, sh) / (double)n;
        const double me = block_sum(se, sh) / (double)n;
        double vc = 0.0, ve = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const float c = ec[i], u = eu[i];
            const float e = cfg_combine(c, u, scale, enc, i, cam_w);
            vc += ((double)c - mc) * ((double)c - mc);
            ve += ((double)e - me) * ((double)e - me);
        }
        vc = block_sum(vc, sh);
        ve = block_sum(ve, sh);
        ratio = (float)sqrt(vc / (double)(n - 1)) / (float)sqrt(ve / (double)(n - 1));
    }
    const float sqrt_at = sqrtf(a_t), sqrt_aprev = sqrtf(a_prev);
    const float dir = sqrtf(fmaxf(1.0f - a_prev - sigma_t * sigma_t, 0.f));
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float c = ec[i], u = eu[i];
        float e = cfg_combine(c, u, scale, enc, i, cam_w);
        if (phi > 0.f) e = __fadd_rn(__fmul_rn(phi, __fmul_rn(e, ratio)), __fmul_rn(1.0f - phi, e));
        const float p0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
        pred_x0[i] = p0;
        x_prev[i] = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_aprev, p0), __fmul_rn(dir, e)), __fmul_rn(sigma_t, noise[i]));
    }
}

// Cluster form (round 2): the one-CTA-per-sample kernel above sits alone at the end of every step (64 us for 65 536 elements: the
// two UNet passes join before it, nothing overlaps it).  Here CS CTAs per sample form a thread-block cluster; every thread keeps its
// <= 4 float4 of e_c / e_u (/ e_nc) in registers across the three phases (means, variances, update), and the per-sample sums are
// combined through distributed shared memory in fixed rank order (deterministic).  Same arithmetic per element as above.
constexpr int CFG_CS = 8, CFG_THREADS = 512, CFG_NV = 6;

__device__ __forceinline__ double cluster_sum(double v, double* sh, double* slot, int cs) {
    const double part = block_sum(v, sh);
    if (threadIdx.x == 0) *slot = part;
    cluster_sync_all();
    double t = 0.0;
    for (int rk = 0; rk < cs; ++rk) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(slot)), "r"(rk));
        double x;
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(x) : "r"(ra) : "memory");
        t += x;
    }
    return t;
}

__global__ void __launch_bounds__(CFG_THREADS) cfg_ddim_cluster_kernel(const float* __restrict__ x, const float* __restrict__ ec,
                                                                       const float* __restrict__ eu, const float* __restrict__ enc,
                                                                       const float* __restrict__ noise, float* __restrict__ x_prev,
                                                                       float* __restrict__ pred_x0, int64_t n, float scale, float cam_w, float phi,
                                                                       float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at) {
    pdl_entry();
    __shared__ double sh[32];
    __shared__ double slots[4];
    const size_t base = (size_t)blockIdx.y * n;
    const int64_t chunk = n / CFG_CS;                    // floats per CTA (a multiple of 4)
    const int64_t c0 = (int64_t)blockIdx.x * chunk;
    const int nv = (int)(chunk >> 2);
    float4 c4[CFG_NV], u4[CFG_NV], e4[CFG_NV];
#pragma unroll
    for (int k = 0; k < CFG_NV; ++k) {
        const int v = threadIdx.x + k * CFG_THREADS;
        if (v < nv) {
            const size_t o = base + c0 + (size_t)v * 4;
            c4[k] = *reinterpret_cast<const float4*>(ec + o);
            u4[k] = *reinterpret_cast<const float4*>(eu + o);
            const float* en = enc ? enc + o : nullptr;
            e4[k].x = cfg_combine(c4[k].x, u4[k].x, scale, en, 0, cam_w);
            e4[k].y = cfg_combine(c4[k].y, u4[k].y, scale, en, 1, cam_w);
            e4[k].z = cfg_combine(c4[k].z, u4[k].z, scale, en, 2, cam_w);
            e4[k].w = cfg_combine(c4[k].w, u4[k].w, scale, en, 3, cam_w);
        }
    }
    float ratio = 1.f;
    if (phi > 0.f) {
        double sc = 0.0, se = 0.0;
#pragma unroll
        for (int k = 0; k < CFG_NV; ++k)
            if ((int)(threadIdx.x + k * CFG_THREADS) < nv) {
                sc += ((double)c4[k].x + (double)c4[k].y) + ((double)c4[k].z + (double)c4[k].w);
                se += ((double)e4[k].x + (double)e4[k].y) + ((double)e4[k].z + (double)e4[k].w);
            }
        const double mc = cluster_sum(sc, sh, &slots[0], CFG_CS) / (double)n;
        const double me = cluster_sum(se, sh, &slots[1], CFG_CS) / (double)n;
        double vc = 0.0, ve = 0.0;
#pragma unroll
        for (int k = 0; k < CFG_NV; ++k)
            if ((int)(threadIdx.x + k * CFG_THREADS) < nv) {
                const double dcx = (double)c4[k].x - mc, dcy = (double)c4[k].y - mc, dcz = (double)c4[k].z - mc, dcw = (double)c4[k].w - mc;
                const double dex = (double)e4[k].x - me, dey = (double)e4[k].y - me, dez = (double)e4[k].z - me, dew = (double)e4[k].w - me;
                vc += (dcx * dcx + dcy * dcy) + (dcz * dcz + dcw * dcw);
                ve += (dex * dex + dey * dey) + (dez * dez + dew * dew);
            }
        vc = cluster_sum(vc, sh, &slots[2], CFG_CS);
        ve = cluster_sum(ve, sh, &slots[3], CFG_CS);
        ratio = (float)sqrt(vc / (double)(n - 1)) / (float)sqrt(ve / (double)(n - 1));
    }
    const float sqrt_at = sqrtf(a_t), sqrt_aprev = sqrtf(a_prev);
    const float dir = sqrtf(fmaxf(1.0f - a_prev - sigma_t * sigma_t, 0.f));
#pragma unroll
    for (int k = 0; k < CFG_NV; ++k) {
        const int v = threadIdx.x + k * CFG_THREADS;
        if (v < nv) {
            const size_t o = base + c0 + (size_t)v * 4;
            const float4 xv = *reinterpret_cast<const float4*>(x + o);
            const float4 nz = *reinterpret_cast<const float4*>(noise + o);
            const float ev[4] = {e4[k].x, e4[k].y, e4[k].z, e4[k].w}, xs[4] = {xv.x, xv.y, xv.z, xv.w}, ns[4] = {nz.x, nz.y, nz.z, nz.w};
            float p0[4], xp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float e = ev[j];
                if (phi > 0.f) e = __fadd_rn(__fmul_rn(phi, __fmul_rn(e, ratio)), __fmul_rn(1.0f - phi, e));
                p0[j] = __fdiv_rn(__fsub_rn(xs[j], __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
                xp[j] = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_aprev, p0[j]), __fmul_rn(dir, e)), __fmul_rn(sigma_t, ns[j]));
            }
            *reinterpret_cast<float4*>(pred_x0 + o) = make_float4(p0[0], p0[1], p0[2], p0[3]);
            *reinterpret_cast<float4*>(x_prev + o) = make_float4(xp[0], xp[1], xp[2], xp[3]);
        }
    }
    if (phi > 0.f) cluster_sync_all();      // no CTA exits while a peer may still read its slots
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int cfg_ddim_update_launch(const float* x, const float* ec, const float* eu, const float* enc, const float* noise, float* x_prev,
                           float* pred_x0, int B, int64_t n, float scale, float cam_w, float phi, float a_t, float a_prev, float sigma_t,
                           float sqrt_one_minus_at, cudaStream_t st) {
    if (B <= 0 || n <= 1) return ERR_BAD_ARG;
    if (n % (CFG_CS * 4) == 0 && n / CFG_CS <= (int64_t)CFG_THREADS * 4 * CFG_NV && B <= 65535 && aligned16(x) && aligned16(ec) && aligned16(eu) &&
        aligned16(noise) && aligned16(x_prev) && aligned16(pred_x0) && (!enc || aligned16(enc))) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(CFG_CS, B);
        cfg.blockDim = dim3(CFG_THREADS);
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        int na = 0;
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CFG_CS;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
        if (pdl_enabled()) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        cfg.attrs = at;
        cfg.numAttrs = na;
        C2V_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cfg_ddim_cluster_kernel, x, ec, eu, enc, noise, x_prev, pred_x0, n, scale, cam_w, phi, a_t, a_prev,
                                          sigma_t, sqrt_one_minus_at));
        return OK;
    }
    C2V_CHECK_CUDA(launch(cfg_ddim_kernel, dim3(B), dim3(1024), 0, st, x, ec, eu, enc, noise, x_prev, pred_x0, n, scale, cam_w, phi, a_t, a_prev, sigma_t, sqrt_one_minus_at));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
