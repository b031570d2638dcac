// Temporal self-attention: sequence length T (8 or 16 frames) per pixel, head dim 64.
//   replaces the einsum softmax path of CrossAttention.forward for the TemporalTransformer
//   (R/lvdm/modules/attention.py:105-129 with q = k = v of length T; attn1 and attn2 of
//   R/model/modules/modified_forwards.py:529-534).
// 0.007 TFLOP per UNet pass: latency / HBM bound, so a 128-row tcgen05 tile would be >87 % padding.
// One warp per (batch, pixel, head): q/k/v rows (128 B each) are staged in shared memory with coalesced
// 16-byte loads straight from the packed QKV GEMM output in its native [B, T, HW, 3*H*64] layout (the
// "(b hw) t c" rearrange of the reference is just a stride here), scores and softmax live in registers.
#include "common.cuh"
#include "kernels.h"

namespace c2v {

constexpr int TA_WARPS = 4;
constexpr int TA_LD = 66;   // padded row stride in bf16 elements (33 words: conflict-free row-per-lane reads)

template <int T>
__global__ void __launch_bounds__(TA_WARPS * 32) attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                      int B, int HW, int heads) {
    constexpr int G = 32 / T;          // lanes sharing one query row
    constexpr int KPL = T / G;         // keys per lane
    constexpr int DPL = 64 / G;        // output dims per lane
    __shared__ __nv_bfloat16 sm[TA_WARPS][3][T][TA_LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * TA_WARPS + warp;   // (b, pix, head)
    const int64_t total = (int64_t)B * HW * heads;
    if (item >= total) return;
    const int head = (int)(item % heads);
    const int64_t bp = item / heads;
    const int pix = (int)(bp % HW);
    const int b = (int)(bp / HW);
    const int C = heads * 64;
    const int ld = 3 * C;

    // ---- stage q, k, v: 8 lanes x 16 B per 128-byte row, 4 rows per instruction ----
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int t0 = 0; t0 < T; t0 += 4) {
            const int t = t0 + (lane >> 3);
            const size_t row = ((size_t)b * T + t) * HW + pix;
            const uint4 v = *reinterpret_cast<const uint4*>(qkv + row * ld + m * C + head * 64 + (lane & 7) * 8);
            uint32_t* dst = reinterpret_cast<uint32_t*>(&sm[warp][m][t][(lane & 7) * 8]);
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    }
    __syncwarp();

    const int i = lane % T, h = lane / T;
    // ---- scores for keys [h*KPL, (h+1)*KPL) ----
    float s[KPL];
#pragma unroll
    for (int jj = 0; jj < KPL; ++jj) s[jj] = 0.f;
    const __nv_bfloat162* qrow = reinterpret_cast<const __nv_bfloat162*>(&sm[warp][0][i][0]);
#pragma unroll 8
    for (int d2 = 0; d2 < 32; ++d2) {
        const float2 qv = __bfloat1622float2(qrow[d2]);
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) {
            const float2 kv = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&sm[warp][1][h * KPL + jj][0])[d2]);
            s[jj] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, s[jj]));
        }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < KPL; ++jj) {
        s[jj] *= 0.125f;
        mx = fmaxf(mx, s[jj]);
    }
#pragma unroll
    for (int o = T; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < KPL; ++jj) {
        s[jj] = __expf(s[jj] - mx);
        sum += s[jj];
    }
#pragma unroll
    for (int o = T; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;

    // ---- out[i, h*DPL : (h+1)*DPL] = sum_j p[i,j] v[j, :] ----
    float acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[d] = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, s[jj], g * T + i) * inv;   // probability of key g*KPL+jj for row i
            const __nv_bfloat162* vrow = reinterpret_cast<const __nv_bfloat162*>(&sm[warp][2][g * KPL + jj][h * DPL]);
#pragma unroll
            for (int d2 = 0; d2 < DPL / 2; ++d2) {
                const float2 vv = __bfloat1622float2(vrow[d2]);
                acc[2 * d2] = fmaf(pj, vv.x, acc[2 * d2]);
                acc[2 * d2 + 1] = fmaf(pj, vv.y, acc[2 * d2 + 1]);
            }
        }
    }
    const size_t orow = ((size_t)b * T + i) * HW + pix;
    __nv_bfloat16* o = out + orow * C + head * 64 + h * DPL;
#pragma unroll
    for (int d = 0; d < DPL; d += 8)
        *reinterpret_cast<uint4*>(o + d) = make_uint4(pack_bf16(acc[d], acc[d + 1]), pack_bf16(acc[d + 2], acc[d + 3]),
                                                      pack_bf16(acc[d + 4], acc[d + 5]), pack_bf16(acc[d + 6], acc[d + 7]));
}

int attention_temporal_launch(const void* qkv, void* out, int B, int T, int HW, int heads, cudaStream_t st) {
    const int64_t total = (int64_t)B * HW * heads;
    const int grid = (int)((total + TA_WARPS - 1) / TA_WARPS);
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    switch (T) {
        case 8: attn_temporal_kernel<8><<<grid, TA_WARPS * 32, 0, st>>>(q, o, B, HW, heads); break;
        case 16: attn_temporal_kernel<16><<<grid, TA_WARPS * 32, 0, st>>>(q, o, B, HW, heads); break;
        default: return ERR_UNSUPPORTED;
    }
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
