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
                                                                      int B, int HW, int heads, int ldo) {
    pdl_entry();
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
    __nv_bfloat16* o = out + orow * ldo + head * 64 + h * DPL;
#pragma unroll
    for (int d = 0; d < DPL; d += 8)
        *reinterpret_cast<uint4*>(o + d) = make_uint4(pack_bf16(acc[d], acc[d + 1]), pack_bf16(acc[d + 2], acc[d + 3]),
                                                      pack_bf16(acc[d + 4], acc[d + 5]), pack_bf16(acc[d + 6], acc[d + 7]));
}

// ------------------------------------------------------------------------------------------------
// T = 16: the 16 x 16 x 64 problem of one (pixel, head) is exactly two warp-level tensor-core shapes, so one warp does
//   S = Q K^T   8 x mma.sync m16n8k16 (4 k-steps x 2 key tiles), fragments by ldmatrix from the staged rows
//   softmax     on the accumulator fragments (a query row lives in one quad: two xor-shuffles)
//   O = P V     8 x mma.sync (P re-used straight from the S accumulators as the A fragment, V by ldmatrix.trans)
// ~150 instructions per (pixel, head) instead of ~2000 for the scalar version below; loads and stores stay 16-byte coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int TM_LD = 72;    // smem row stride in elements (144 B: 16-byte aligned rows, conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if C2V_OPERAND_IS_FP16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#endif
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(TA_WARPS * 32) attn_temporal16_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                            int B, int HW, int heads, int ldo) {
    pdl_entry();
    constexpr int T = 16;
    __shared__ __align__(16) __nv_bfloat16 sm[TA_WARPS][3][T][TM_LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * TA_WARPS + warp;   // (b, pix, head)
    if (item >= (int64_t)B * HW * heads) return;
    const int head = (int)(item % heads);
    const int64_t bp = item / heads;
    const int pix = (int)(bp % HW), b = (int)(bp / HW);
    const int C = heads * 64, ld = 3 * C;
    // ---- stage q, k, v: 8 lanes x 16 B per 128-byte row, 4 rows per instruction (all 12 loads in flight) ----
    uint4 st[12];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int t0 = 0; t0 < T; t0 += 4) {
            const size_t row = ((size_t)b * T + t0 + (lane >> 3)) * HW + pix;
            st[m * 4 + t0 / 4] = *reinterpret_cast<const uint4*>(qkv + row * ld + m * C + head * 64 + (lane & 7) * 8);
        }
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int t0 = 0; t0 < T; t0 += 4)
            *reinterpret_cast<uint4*>(&sm[warp][m][t0 + (lane >> 3)][(lane & 7) * 8]) = st[m * 4 + t0 / 4];
    __syncwarp();
    // ---- S = Q K^T (16 queries x 16 keys), fp32 accumulators: s[nt] = keys 8nt..8nt+7 ----
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4], kb[4];
        ldsm_x4(a, &sm[warp][0][(lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + (lane >> 4) * 8]);
        ldsm_x4(kb, &sm[warp][1][(lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
        mma_16816(s[0], a, kb[0], kb[1]);
        mma_16816(s[1], a, kb[2], kb[3]);
    }
    // ---- softmax over the 16 keys of rows g = lane / 4 (elements 0, 1) and g + 8 (elements 2, 3); quad = one row ----
    float m0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1])), m1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    }
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = __expf((s[nt][0] - m0) * 0.125f); s[nt][1] = __expf((s[nt][1] - m0) * 0.125f);
        s[nt][2] = __expf((s[nt][2] - m1) * 0.125f); s[nt][3] = __expf((s[nt][3] - m1) * 0.125f);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    // ---- O = P V: the S accumulator fragments are the A fragment of P (16 x 16) ----
    uint32_t pa[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]), pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __syncwarp();                       // every lane has read Q: its rows become the output staging buffer
#pragma unroll
    for (int np = 0; np < 4; ++np) {    // two 8-wide d tiles per ldmatrix.x4.trans
        uint32_t vb[4];
        ldsm_x4_t(vb, &sm[warp][2][(lane & 7) + ((lane >> 3) & 1) * 8][(2 * np + (lane >> 4)) * 8]);
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        mma_16816(o0, pa, vb[0], vb[1]);
        mma_16816(o1, pa, vb[2], vb[3]);
        const int g = lane >> 2, c2 = (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(&sm[warp][0][g][16 * np + c2]) = pack_bf16(o0[0] * i0, o0[1] * i0);
        *reinterpret_cast<uint32_t*>(&sm[warp][0][g + 8][16 * np + c2]) = pack_bf16(o0[2] * i1, o0[3] * i1);
        *reinterpret_cast<uint32_t*>(&sm[warp][0][g][16 * np + 8 + c2]) = pack_bf16(o1[0] * i0, o1[1] * i0);
        *reinterpret_cast<uint32_t*>(&sm[warp][0][g + 8][16 * np + 8 + c2]) = pack_bf16(o1[2] * i1, o1[3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int t0 = 0; t0 < T; t0 += 4) {
        const int t = t0 + (lane >> 3);
        const size_t orow = ((size_t)b * T + t) * HW + pix;
        *reinterpret_cast<uint4*>(out + orow * ldo + head * 64 + (lane & 7) * 8) = *reinterpret_cast<const uint4*>(&sm[warp][0][t][(lane & 7) * 8]);
    }
}

int attention_temporal_launch(const void* qkv, void* out, int B, int T, int HW, int heads, int ldo, cudaStream_t st) {
    if (ldo == 0) ldo = heads * 64;
    if (ldo < heads * 64 || ldo % 8 != 0) return ERR_UNSUPPORTED;      // 16-byte row stores
    const int64_t total = (int64_t)B * HW * heads;
    const int grid = (int)((total + TA_WARPS - 1) / TA_WARPS);
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    switch (T) {
        case 8: C2V_CHECK_CUDA(launch(attn_temporal_kernel<8>, dim3(grid), dim3(TA_WARPS * 32), 0, st, q, o, B, HW, heads, ldo)); break;
        case 16: C2V_CHECK_CUDA(launch(attn_temporal16_mma_kernel, dim3(grid), dim3(TA_WARPS * 32), 0, st, q, o, B, HW, heads, ldo)); break;
        default: return ERR_UNSUPPORTED;
    }
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
