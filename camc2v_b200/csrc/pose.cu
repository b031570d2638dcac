// Kernels of the camera pose encoder (SURVEY f-2; reference camera_pose_encoder.py:295-376) that the UNet path does not
// already provide: PixelUnshuffle into channel-last rows, 2x2 average pooling, ReLU, and temporal self-attention with a
// head dim other than 64 (320/8 = 40, 640/8 = 80, 1280/8 = 160).  Once per sample, HBM-bound; the GEMMs / convs / layer norms
// around them are the UNet's own kernels.
#include "common.cuh"
#include "kernels.h"

namespace c2v {

static inline int grid_for(int64_t work, int threads, int cap = 148 * 16) {
    const int64_t g = (work + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// nn.PixelUnshuffle(r) on '(b f) c h w' (camera_pose_encoder.py:359-361) -> operand-dtype rows (b, f, y, x) x [c*r*r + dy*r + dx]
__global__ void pixel_unshuffle_cl_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int C, int T, int H, int W, int r) {
    const int Ho = H / r, Wo = W / r, Co = C * r * r;
    const int64_t total = (int64_t)B * T * Ho * Wo * Co;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Co);
        int64_t row = i / Co;
        const int x = (int)(row % Wo); row /= Wo;
        const int y = (int)(row % Ho); row /= Ho;
        const int t = (int)(row % T);
        const int b = (int)(row / T);
        const int dx = co % r, dy = (co / r) % r, c = co / (r * r);
        out[i] = __float2bfloat16(in[((((size_t)b * C + c) * T + t) * H + (y * r + dy)) * W + (x * r + dx)]);
    }
}

int pixel_unshuffle_cl_launch(const float* in, void* out, int B, int C, int T, int H, int W, int r, cudaStream_t st) {
    if (r <= 0 || H % r || W % r) return ERR_BAD_ARG;
    pixel_unshuffle_cl_kernel<<<grid_for((int64_t)B * C * T * H * W, 256), 256, 0, st>>>(in, reinterpret_cast<__nv_bfloat16*>(out), B, C, T, H, W, r);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// nn.AvgPool2d(2, 2) (Downsample with use_conv=False, camera_pose_encoder.py:229-231) on channel-last fp32 rows (n, y, x);
// writes the fp32 result and (optionally) its operand-dtype copy for the next GEMM.
__global__ void avgpool2_cl_kernel(const float* __restrict__ in, float* __restrict__ out, __nv_bfloat16* __restrict__ out_b, int N, int H, int W,
                                   int C) {
    const int nv = C >> 2, Ho = H >> 1, Wo = W >> 1;
    const int64_t total = (int64_t)N * Ho * Wo * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % nv);
        int64_t row = i / nv;
        const int xo = (int)(row % Wo); row /= Wo;
        const int yo = (int)(row % Ho);
        const int n = (int)(row / Ho);
        const float* p = in + (((size_t)n * H + 2 * yo) * W + 2 * xo) * C + cv * 4;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + C);
        const float4 c = *reinterpret_cast<const float4*>(p + (size_t)W * C), d = *reinterpret_cast<const float4*>(p + (size_t)W * C + C);
        float4 o;
        o.x = (a.x + b.x + c.x + d.x) * 0.25f;
        o.y = (a.y + b.y + c.y + d.y) * 0.25f;
        o.z = (a.z + b.z + c.z + d.z) * 0.25f;
        o.w = (a.w + b.w + c.w + d.w) * 0.25f;
        *reinterpret_cast<float4*>(out + i * 4) = o;
        if (out_b) *reinterpret_cast<uint2*>(out_b + i * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
}

int avgpool2_cl_launch(const float* in, float* out, void* out_b, int N, int H, int W, int C, cudaStream_t st) {
    if (C % 4 || H % 2 || W % 2) return ERR_UNSUPPORTED;
    avgpool2_cl_kernel<<<grid_for((int64_t)N * (H / 2) * (W / 2) * (C / 4), 256), 256, 0, st>>>(in, out, reinterpret_cast<__nv_bfloat16*>(out_b), N, H, W,
                                                                                               C);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// in-place ReLU on an operand-dtype tensor (ResnetBlock.act, camera_pose_encoder.py:262)
__global__ void relu_kernel(__nv_bfloat16* __restrict__ x, int64_t n8) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = *reinterpret_cast<uint4*>(x + i * 8);
        __nv_bfloat162* e = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(e[j]);
            e[j] = __floats2bfloat162_rn(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f));
        }
        *reinterpret_cast<uint4*>(x + i * 8) = v;
    }
}

int relu_launch(void* x, int64_t n, cudaStream_t st) {
    if (n % 8) return ERR_UNSUPPORTED;
    relu_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x), n / 8);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// Temporal self-attention over T <= 16 frames per (b, pixel, head) with any head dim D (multiple of 8, <= 160): the
// AttnProcessor2_0 math of diffusers' Attention (softmax(q k^T / sqrt(D)) v, fp32 accumulation) on the packed q | k | v rows
// [B, T, HW, 3*heads*D].  One warp per item; q / k / v staged in shared memory, scores 8 per lane, output D/32 dims per lane.
constexpr int TH_WARPS = 2, TH_MAX_T = 16, TH_MAX_D = 160;

__global__ void __launch_bounds__(TH_WARPS * 32) attn_temporal_hd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B,
                                                                         int T, int HW, int heads, int D, float scale) {
    __shared__ __align__(16) __nv_bfloat16 sm[TH_WARPS][3][TH_MAX_T][TH_MAX_D];
    __shared__ float sp[TH_WARPS][TH_MAX_T][TH_MAX_T + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * TH_WARPS + warp;
    if (item >= (int64_t)B * HW * heads) return;
    const int head = (int)(item % heads);
    const int64_t bp = item / heads;
    const int pix = (int)(bp % HW);
    const int b = (int)(bp / HW);
    const int C = heads * D, ld = 3 * C, dv = D >> 3;
    for (int i = lane; i < 3 * T * dv; i += 32) {
        const int c8 = i % dv, t = (i / dv) % T, m = i / (dv * T);
        const int64_t row = ((int64_t)b * T + t) * HW + pix;
        *reinterpret_cast<uint4*>(&sm[warp][m][t][c8 * 8]) = *reinterpret_cast<const uint4*>(qkv + row * ld + m * C + head * D + c8 * 8);
    }
    __syncwarp();
    // scores: lane -> query t = lane / 2, keys s0 = (lane & 1) * 8 .. +8
    const int t = lane >> 1, s0 = (lane & 1) * 8;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    if (t < T) {
        for (int d = 0; d < D; d += 2) {
            const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sm[warp][0][t][d]));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (s0 + j < T) {
                    const float2 k = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sm[warp][1][s0 + j][d]));
                    s[j] = fmaf(q.x, k.x, fmaf(q.y, k.y, s[j]));
                }
            }
        }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j] = (s0 + j < T) ? s[j] * scale : -INFINITY;
        mx = fmaxf(mx, s[j]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j] = (s0 + j < T) ? __expf(s[j] - mx) : 0.f;
        sum += s[j];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float inv = 1.f / sum;
    if (t < T) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sp[warp][t][s0 + j] = s[j] * inv;
    }
    __syncwarp();
    // output: lane -> pairs of dims; loop over (t, d2)
    const int d2n = D >> 1;
    for (int i = lane; i < T * d2n; i += 32) {
        const int d2 = i % d2n, tq = i / d2n;
        float ox = 0.f, oy = 0.f;
        for (int k = 0; k < T; ++k) {
            const float p = sp[warp][tq][k];
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sm[warp][2][k][d2 * 2]));
            ox = fmaf(p, v.x, ox);
            oy = fmaf(p, v.y, oy);
        }
        const int64_t orow = ((int64_t)b * T + tq) * HW + pix;
        *reinterpret_cast<uint32_t*>(out + orow * C + head * D + d2 * 2) = pack_bf16(ox, oy);
    }
}

int attention_temporal_hd_launch(const void* qkv, void* out, int B, int T, int HW, int heads, int D, cudaStream_t st) {
    if (T < 1 || T > TH_MAX_T || D < 8 || D > TH_MAX_D || D % 8) return ERR_UNSUPPORTED;
    const int64_t total = (int64_t)B * HW * heads;
    attn_temporal_hd_kernel<<<(unsigned)((total + TH_WARPS - 1) / TH_WARPS), TH_WARPS * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), B, T, HW, heads, D, 1.f / sqrtf((float)D));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
