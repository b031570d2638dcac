// Dense attention with a SHORT key sequence (the API routes lk <= 128 here; the kernel handles up to 256), head dim 64: text
// cross-attention (77 keys), per-frame image cross-attention (16 keys), spatial self-attention of the 8x8 / 4x4 levels (64 / 16).
//   replaces the same reference calls as attn_fa.cu (R/lvdm/modules/attention.py:105-144, 177, 189) for these shapes.
// A 128 x 64 tcgen05 tile pipeline is mostly fixed latency here (one or two key tiles per CTA: ~20 us per launch whatever the
// size); with all of K and V resident in shared memory the problem is a warp-level one:
//   one warp = 16 query rows; S = Q K^T by mma.sync m16n8k16 (fragments by ldmatrix), online softmax over 64-key blocks on the
//   accumulator fragments (a row lives in one quad), O += P V with P re-used from the S accumulators and V by ldmatrix.trans.
// 8 warps (128 queries) per CTA share the K / V rows of their (kv batch, head).  HBM-bound: q + out once, K / V from L2.
#include "attn.h"
#include "common.cuh"

namespace c2v {

constexpr int AS_WARPS = 8;
constexpr int AS_LD = 72;          // smem row stride in elements (144 B: 16-byte aligned rows, conflict-free ldmatrix)
constexpr int AS_MAX_LK = 256;

__device__ __forceinline__ void as_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void as_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void as_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if C2V_OPERAND_IS_FP16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#endif
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnSmallArgs {
    const __nv_bfloat16 *q, *k, *v;
    __nv_bfloat16* out;
    int lq, lk, kv_div;
    int ldq, ldk, ldv, ldo;
    long long q_bstride, k_bstride, v_bstride, o_bstride;
    float scale_log2, out_scale;
    int accumulate;
};

__global__ void __launch_bounds__(AS_WARPS * 32, 2) attn_small_kernel(const AttnSmallArgs p) {
    pdl_entry();
    extern __shared__ __align__(16) uint8_t as_smem[];
    const int lk_pad = (p.lk + 15) & ~15;
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(as_smem);
    __nv_bfloat16* sV = sK + (size_t)lk_pad * AS_LD;
    __nv_bfloat16* sQ = sV + (size_t)lk_pad * AS_LD;               // [AS_WARPS][16][AS_LD]: Q rows, later the output staging
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, b = blockIdx.z, bkv = b / p.kv_div;
    const int q0 = blockIdx.x * (AS_WARPS * 16) + warp * 16;
    // ---- K, V rows of this (kv batch, head) -> smem: 8 lanes x 16 B per 128-byte row; rows >= lk are zero ----
    const __nv_bfloat16* kg = p.k + (size_t)bkv * p.k_bstride + head * 64;
    const __nv_bfloat16* vg = p.v + (size_t)bkv * p.v_bstride + head * 64;
    for (int r = threadIdx.x >> 3; r < lk_pad; r += AS_WARPS * 4) {
        uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
        if (r < p.lk) {
            kk = *reinterpret_cast<const uint4*>(kg + (size_t)r * p.ldk + (lane & 7) * 8);
            vv = *reinterpret_cast<const uint4*>(vg + (size_t)r * p.ldv + (lane & 7) * 8);
        }
        *reinterpret_cast<uint4*>(sK + (size_t)r * AS_LD + (lane & 7) * 8) = kk;
        *reinterpret_cast<uint4*>(sV + (size_t)r * AS_LD + (lane & 7) * 8) = vv;
    }
    // ---- this warp's 16 query rows ----
    __nv_bfloat16* myQ = sQ + (size_t)warp * 16 * AS_LD;
    const __nv_bfloat16* qg = p.q + (size_t)b * p.q_bstride + head * 64;
#pragma unroll
    for (int t0 = 0; t0 < 16; t0 += 4) {
        const int r = q0 + t0 + (lane >> 3);
        uint4 qq = make_uint4(0u, 0u, 0u, 0u);
        if (r < p.lq) qq = *reinterpret_cast<const uint4*>(qg + (size_t)r * p.ldq + (lane & 7) * 8);
        *reinterpret_cast<uint4*>(myQ + (t0 + (lane >> 3)) * AS_LD + (lane & 7) * 8) = qq;
    }
    __syncthreads();
    if (q0 >= p.lq) return;                                          // warp-uniform; no block-level sync below
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) as_ldsm_x4(qa[ks], myQ + ((lane & 7) + ((lane >> 3) & 1) * 8) * AS_LD + ks * 16 + (lane >> 4) * 8);

    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // rows g = lane / 4 and g + 8 (running max in the log2 domain)
    const int c2 = (lane & 3) * 2;
    for (int kb0 = 0; kb0 < lk_pad; kb0 += 64) {
        // ---- S block: 16 queries x up to 64 keys ----
        float s[8][4];
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
            s[2 * ntp][0] = s[2 * ntp][1] = s[2 * ntp][2] = s[2 * ntp][3] = 0.f;
            s[2 * ntp + 1][0] = s[2 * ntp + 1][1] = s[2 * ntp + 1][2] = s[2 * ntp + 1][3] = 0.f;
            if (kb0 + ntp * 16 < lk_pad) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t kf[4];
                    as_ldsm_x4(kf, sK + (size_t)(kb0 + ntp * 16 + (lane & 7) + (lane >> 4) * 8) * AS_LD + ks * 16 + ((lane >> 3) & 1) * 8);
                    as_mma(s[2 * ntp], qa[ks], kf[0], kf[1]);
                    as_mma(s[2 * ntp + 1], qa[ks], kf[2], kf[3]);
                }
            }
        }
        // ---- mask keys >= lk, block row max ----
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int key = kb0 + nt * 8 + c2;
            if (key >= p.lk) s[nt][0] = s[nt][2] = -INFINITY;
            if (key + 1 >= p.lk) s[nt][1] = s[nt][3] = -INFINITY;
            bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
            bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int x = 1; x < 4; x <<= 1) {
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, x));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, x));
        }
        const float n0 = fmaxf(m0, bm0 * p.scale_log2), n1 = fmaxf(m1, bm1 * p.scale_log2);   // every block has >= 1 valid key
        const float a0 = fast_exp2(m0 - n0), a1 = fast_exp2(m1 - n1);                          // exp2(-inf) = 0 on the first block
        m0 = n0;
        m1 = n1;
        l0 *= a0;
        l1 *= a1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= a0; o[nt][1] *= a0;
            o[nt][2] *= a1; o[nt][3] *= a1;
        }
        // ---- probabilities (kept in the accumulator registers), O += P V ----
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = fast_exp2(fmaf(s[nt][0], p.scale_log2, -m0)); s[nt][1] = fast_exp2(fmaf(s[nt][1], p.scale_log2, -m0));
            s[nt][2] = fast_exp2(fmaf(s[nt][2], p.scale_log2, -m1)); s[nt][3] = fast_exp2(fmaf(s[nt][3], p.scale_log2, -m1));
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                              // 16 keys per k-step
            if (kb0 + kk * 16 < lk_pad) {
                const uint32_t pa[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                                        pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t vf[4];
                    as_ldsm_x4_t(vf, sV + (size_t)(kb0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * AS_LD + (2 * np + (lane >> 4)) * 8);
                    as_mma(o[2 * np], pa, vf[0], vf[1]);
                    as_mma(o[2 * np + 1], pa, vf[2], vf[3]);
                }
            }
        }
    }
    // ---- epilogue: O / l * out_scale (+ previous out) -> 16-bit -> smem (this warp's Q rows) -> 16-byte row stores ----
#pragma unroll
    for (int x = 1; x < 4; x <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, x);
        l1 += __shfl_xor_sync(0xffffffffu, l1, x);
    }
    const float i0 = l0 > 0.f ? p.out_scale / l0 : 0.f, i1 = l1 > 0.f ? p.out_scale / l1 : 0.f;
    const int g = lane >> 2;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(myQ + g * AS_LD + nt * 8 + c2) = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
        *reinterpret_cast<uint32_t*>(myQ + (g + 8) * AS_LD + nt * 8 + c2) = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
    }
    __syncwarp();
    __nv_bfloat16* og = p.out + (size_t)b * p.o_bstride + head * 64;
#pragma unroll
    for (int t0 = 0; t0 < 16; t0 += 4) {
        const int r = q0 + t0 + (lane >> 3);
        if (r < p.lq) {
            uint4 val = *reinterpret_cast<const uint4*>(myQ + (t0 + (lane >> 3)) * AS_LD + (lane & 7) * 8);
            uint4* dst = reinterpret_cast<uint4*>(og + (size_t)r * p.ldo + (lane & 7) * 8);
            if (p.accumulate) {
                const uint4 prev = *dst;
                const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&val);
                const __nv_bfloat162* c = reinterpret_cast<const __nv_bfloat162*>(&prev);
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__low2float(a[e]) + __low2float(c[e]), __high2float(a[e]) + __high2float(c[e]));
                val = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *dst = val;
        }
    }
}

int attn_small_launch(const void* q, const void* k, const void* v, void* out, int bq, int lq, int lk, int heads, int kv_div, int ldq, int ldk,
                      int ldv, int ldo, long long q_bstride, long long k_bstride, long long v_bstride, long long o_bstride, float scale_log2,
                      float out_scale, int accumulate, cudaStream_t st) {
    if (lk <= 0 || lk > AS_MAX_LK) return ERR_UNSUPPORTED;
    AttnSmallArgs a;
    a.q = reinterpret_cast<const __nv_bfloat16*>(q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(v);
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.lq = lq; a.lk = lk; a.kv_div = kv_div;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
    a.q_bstride = q_bstride; a.k_bstride = k_bstride; a.v_bstride = v_bstride; a.o_bstride = o_bstride;
    a.scale_log2 = scale_log2; a.out_scale = out_scale; a.accumulate = accumulate;
    const int lk_pad = (lk + 15) & ~15;
    const size_t smem = ((size_t)2 * lk_pad + AS_WARPS * 16) * AS_LD * sizeof(__nv_bfloat16);
    static bool attr_set = false;
    if (!attr_set) {
        C2V_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(((size_t)2 * AS_MAX_LK + AS_WARPS * 16) * AS_LD * sizeof(__nv_bfloat16))));
        attr_set = true;
    }
    C2V_CHECK_CUDA(launch(attn_small_kernel, dim3((lq + AS_WARPS * 16 - 1) / (AS_WARPS * 16), heads, bq), dim3(AS_WARPS * 32), smem, st, a));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
