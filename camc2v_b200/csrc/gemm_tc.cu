// tcgen05 GEMM for every dense contraction of the UNet step that is not attention:
//   * nn.Linear            (attention.py:58-62, 277, 301, 348, 378, 434, 451-455; epipolar.py:55-63; ...)
//   * Conv2d 3x3 / 1x1     (openaimodel3d.py:151-155, 175-187, 68-70, 96, 386, 561-565)  as implicit GEMM
//   * Conv3d (3,1,1)       (openaimodel3d.py:255-266)                                     as implicit GEMM
//
//   out[m, n] = sum_{tap, k} A[row(m) shifted by tap, k] * Wt[n, tap*Cin + k]  (+ bias[n] + rowbias[m / rows_per_group, n]
//               + residual[m, n]);   optional GEGLU epilogue  out = x * gelu_erf(gate).
//
// Data layout: activations are channels-last bf16 ([rows, C]); A tiles are fetched by TMA straight from
// the activation tensor — for convolutions a 4-D tensor map (C, W, H, N) / (C, HW, T, B) is sampled at
// tap-shifted coordinates and TMA's out-of-bounds zero fill provides the padding, so no im2col buffer
// ever exists in HBM.  Weights are [N, taps*Cin] bf16 (K-major).  Both operands land in 128B-swizzled
// shared memory and feed tcgen05.mma (M=128, N=BN, K=16) with the fp32 accumulator in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global).  3-stage smem ring, 2 CTAs per SM so that one
// CTA's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "gemm_tc.h"

namespace c2v {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MIN_STAGES = 3;    // the TMA epilogue stages BN/32 x 16 KB chunks in the idle pipeline buffers
constexpr int MAX_STAGES = 8;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int VEC_BYTES = 2 * BN * 4;      // bias / row-bias of the tile's columns, staged by the epilogue warps while the main loop runs
    static constexpr int total(int stages) { return stages * STAGE_BYTES + 256 + VEC_BYTES + 1024; }  // + barriers/tmem ptr + vectors + 1024B alignment slack
};

// EPI_WARPS = 4 everywhere except the GEGLU projections, whose erf epilogue (2x the main loop's time at K = 320) is split over
// two warps per TMEM lane quarter (EPI_WARPS = 8, 320 threads): warps 2..5 take the even 16-column chunks, warps 6..9 the odd ones.
template <int BN, int EPI_WARPS = 4>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 2) gemm_tc_kernel(const __grid_constant__ GemmKernelArgs p) {
    using S = GemmSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.stages;                          // smem ring depth chosen per launch (gemm_tc_launch)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* acc_bar = empty_bar + MAX_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint64_t* res_bar = acc_bar + 2;                      // [BN / 32] residual-chunk arrival barriers (TMA epilogue)
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + 256);      // [BN] bias, [BN] row bias (TMA epilogue)
    float* s_rb = s_bias + BN;

    const int warp = threadIdx.x >> 5;
    const int m_tile = blockIdx.x;
    const int n0 = blockIdx.y * BN;
    const int total_iters = p.taps * p.k_chunks;
    const int it_begin = (int)((long long)total_iters * blockIdx.z / p.splits);
    const int it_end = (int)((long long)total_iters * (blockIdx.z + 1) / p.splits);
    const int iters = it_end - it_begin;
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_bar, 1);
        for (int c = 0; c < BN / 32; ++c) mbar_init(&res_bar[c], 1);
        if (p.tma_epi) {
            tma_prefetch_desc(&p.tmO);
            if (p.residual && p.splits == 1) tma_prefetch_desc(&p.tmR);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_entry();      // everything above is CTA-local set-up and overlaps the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            // first output row of this tile -> base coordinates of the A box
            const int m0 = m_tile * p.tile_rows;
            int c1 = 0, c2 = 0, c3 = 0;
            if (p.a_mode == A_PLAIN) {
                c1 = m0;
            } else if (p.a_mode == A_CONV2D) {          // rows ordered (n, y, x)
                const int hw = p.dim1 * p.dim2;
                c3 = m0 / hw;
                c2 = (m0 % hw) / p.dim1;
                c1 = (m0 % hw) % p.dim1;           // 0 unless the image is wider than a tile (then a multiple of 128)
            } else {                                    // A_CONVT: rows ordered (b, t, p)
                const int thw = p.dim1 * p.dim2;
                c3 = m0 / thw;
                c2 = (m0 % thw) / p.dim1;
                c1 = m0 % p.dim1;
            }
            const uint32_t tx = (uint32_t)p.tile_rows * BK * 2 + S::B_BYTES;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                const int git = it_begin + it;
                const int tap = git / p.k_chunks, kc = git - tap * p.k_chunks;
                int d1 = 0, d2 = 0;
                if (p.taps > 1) {
                    if (p.a_mode == A_CONV2D) {
                        d1 = tap % 3 - 1;
                        d2 = tap / 3 - 1;
                    } else if (p.a_mode == A_CONVT) {
                        d2 = tap - 1;
                    }
                }
                mbar_wait<100>(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                uint8_t* b_dst = a_dst + S::A_BYTES;
                mbar_expect_tx(&full_bar[s], tx);
                if (p.a_mode == A_PLAIN)
                    tma_load_2d(a_dst, &p.tmA, &full_bar[s], kc * BK, c1);
                else
                    tma_load_4d(a_dst, &p.tmA, &full_bar[s], kc * BK, c1 + d1, c2 + d2, c3);
                tma_load_2d(b_dst, &p.tmB, &full_bar[s], git * BK, n0);
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait<20>(&full_bar[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t b_addr = a_addr + S::A_BYTES;
                const uint64_t adesc = umma_desc_sw128(a_addr);
                const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // advance 16 bf16 = 32 B inside the 128B swizzle atom: +2 in the (addr >> 4) field
                    umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
                }
                umma_commit(&empty_bar[s]);
                if (it == iters - 1) umma_commit(acc_bar);
            }
            __syncwarp();
            if (++s == STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int lg = warp & 3;                       // TMEM lane group this warp may access
        const int r = lg * 32 + lane_id();             // row inside the tile
        const int m = m_tile * p.tile_rows + r;
        const bool row_ok = (r < p.tile_rows) && (m < p.M);
        const bool part = p.splits > 1;                 // split-K: store the raw fp32 partial tile, epilogue terms are applied by the reduce
        const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
        const float* rb = (!part && p.rowbias && row_ok) ? p.rowbias + (size_t)(m / p.rows_per_group) * p.N : nullptr;
        const float* bias = part ? nullptr : p.bias;
        if (EPI_WARPS > 4 && (warp >= 6) && p.epi != EPI_GEGLU) {
            // the second epilogue group only exists for the GEGLU epilogue
        } else if (p.tma_epi) {
            // ---- TMA epilogue: residual tile fetched by TMA into the (now idle) pipeline stages with every 32-column chunk
            //      in flight at once; results staged in shared memory (hardware swizzle, conflict-free) and written back by
            //      TMA stores, so every global access of the epilogue is a full-line bulk transfer.
            constexpr int NCH = BN / 32;
            const int m0 = m_tile * p.tile_rows;
            const bool has_res = p.residual != nullptr && !part;
            const bool out_f32 = part || !p.out_bf16;
            const bool leader = threadIdx.x == 64;      // warp 2, lane 0: owns the bulk async-group of the stores
            // The tile's bias (and its row bias, when all rows of the tile share one) are fetched into shared memory NOW, under the
            // main loop: read from global inside the chunk loop, each 32-column chunk paid one exposed L2 / DRAM round trip for a
            // line no one has touched in this pass (ncu: 21 % of the epilogue warps' stall samples of the K = 320 linears).
            const int last_row = min(m0 + p.tile_rows, p.M) - 1;
            const bool rb_tile = !part && p.rowbias && (m0 / p.rows_per_group == last_row / p.rows_per_group);
            if (bias || rb_tile) {
                const float* rbt = rb_tile ? p.rowbias + (size_t)(m0 / p.rows_per_group) * p.N : nullptr;
                for (int i = threadIdx.x - 64; i < BN; i += 128) {
                    const int n = n0 + i;
                    s_bias[i] = (bias && n < p.N) ? bias[n] : 0.f;
                    s_rb[i] = (rbt && n < p.N) ? rbt[n] : 0.f;
                }
                named_bar_sync(1, 128);
            }
            mbar_wait<200>(acc_bar, 0);                 // every MMA has completed: accumulator valid, smem stages free
            tc_fence_after();
            {
            if (leader && has_res) {
                for (int c = 0; c < NCH; ++c) {
                    if (n0 + c * 32 >= p.N) break;
                    mbar_expect_tx(&res_bar[c], (uint32_t)p.tile_rows * 128);
                    tma_load_2d(smem + c * 16384, &p.tmR, &res_bar[c], n0 + c * 32, m0);
                }
            }
            const int flush = (!out_f32 && has_res) ? 1 : 3;
            int c_flushed = 0;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                if (n0 + c * 32 >= p.N) break;          // CTA-uniform
                uint32_t v[32];
                tmem_ld32(trow + c * 32, v);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                uint8_t* buf = smem + c * 16384;
                uint8_t* rowp = buf + r * 128;
                if (has_res) {
                    mbar_wait(&res_bar[c], 0);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 x = *reinterpret_cast<const float4*>(rowp + ((k ^ (r & 7)) << 4));
                        f[4 * k] += x.x; f[4 * k + 1] += x.y; f[4 * k + 2] += x.z; f[4 * k + 3] += x.w;
                    }
                }
                const int nb = n0 + c * 32;
                if (bias) {                                           // columns past N hold 0 and are clipped by the TMA store
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c * 32 + j);
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                if (rb_tile) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_rb + c * 32 + j);
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                } else if (rb) {                                      // the tile spans several row-bias groups (8x8 / 4x4 levels)
                    if (nb + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(rb + nb + j);
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                        }
                    } else {                                          // ragged last N tile: never read past the row of the row bias
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.N) f[j] += rb[nb + j];
                    }
                }
                if (p.epi == EPI_GELU && !part) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf_f(f[j]);
                }
                if (out_f32) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<float4*>(rowp + ((k ^ (r & 7)) << 4)) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                } else {
                    if (has_res) named_bar_sync(1, 128);        // bf16 rows are packed tighter than the fp32 residual rows they replace
                    uint8_t* orow = buf + r * 64;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<uint4*>(orow + ((k ^ ((r >> 1) & 3)) << 4)) =
                            make_uint4(pack_bf16(f[8 * k], f[8 * k + 1]), pack_bf16(f[8 * k + 2], f[8 * k + 3]),
                                       pack_bf16(f[8 * k + 4], f[8 * k + 5]), pack_bf16(f[8 * k + 6], f[8 * k + 7]));
                }
                // Hand finished chunks to the TMA store engine.  Each hand-over costs a proxy fence + a 128-thread barrier, so
                // chunks are flushed in batches of `flush` (per chunk only when bf16 rows alias other threads' residual rows).
                const bool last = (c + 1 == NCH) || (n0 + (c + 1) * 32 >= p.N);
                if (c + 1 - c_flushed >= flush || last) {
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                    if (leader) {
                        for (int cf = c_flushed; cf <= c; ++cf) tma_store_3d(&p.tmO, smem + cf * 16384, n0 + cf * 32, m0, blockIdx.z);
                        tma_store_commit();
                    }
                    c_flushed = c + 1;
                }
            }
            if (leader) tma_store_wait_read_all();
            tc_fence_before();
            }
        } else {
        mbar_wait<200>(acc_bar, 0);
        tc_fence_after();
        const float* resid = part ? nullptr : p.residual;
        const int out_bf16 = part ? 0 : p.out_bf16;
        const int ldo = part ? p.N : p.ldo;
        void* const outp = part ? (void*)(reinterpret_cast<float*>(p.out) + (size_t)blockIdx.z * p.M * p.N) : p.out;
        if (p.epi == EPI_GEGLU) {
            constexpr int HALF = BN / 2;
            const int no = blockIdx.y * HALF;          // output column base
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.ldo + no;
            constexpr int NGRP = EPI_WARPS / 4;
            const int grp = (warp - 2) >> 2;
#pragma unroll 1
            for (int c = grp * 16; c < HALF; c += 16 * NGRP) {
                uint32_t xv[16], gv[16];
                tmem_ld16(trow + c, xv);
                tmem_ld16(trow + HALF + c, gv);
                float bx[16], bg[16];                  // bias of the value / gate columns (same for every row: L1 broadcast)
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 a4 = *reinterpret_cast<const float4*>(p.bias + n0 + c + j);
                        const float4 g4 = *reinterpret_cast<const float4*>(p.bias + n0 + HALF + c + j);
                        bx[j] = a4.x; bx[j + 1] = a4.y; bx[j + 2] = a4.z; bx[j + 3] = a4.w;
                        bg[j] = g4.x; bg[j + 1] = g4.y; bg[j + 2] = g4.z; bg[j + 3] = g4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) bx[j] = bg[j] = 0.f;
                }
                tmem_ld_wait();
                if (row_ok) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2)
                    {
                        const float2 y = geglu_fast2(fadd2(make_float2(__uint_as_float(xv[j]), __uint_as_float(xv[j + 1])), make_float2(bx[j], bx[j + 1])),
                                                     fadd2(make_float2(__uint_as_float(gv[j]), __uint_as_float(gv[j + 1])), make_float2(bg[j], bg[j + 1])));
                        pk[j / 2] = pack_bf16(y.x, y.y);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(o + c);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + c, v);
                tmem_ld_wait();
                if (row_ok && n0 + c < p.N) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    const int nvalid = min(32, p.N - (n0 + c));
                    if (nvalid == 32) {
                        if (bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(bias + n0 + c + j);
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (rb) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(rb + n0 + c + j);
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (resid) {
                            const float* rs = resid + (size_t)m * p.ldr + n0 + c;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(rs + j);
                                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                            }
                        }
                        if (p.epi == EPI_GELU && !part) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_erf_f(f[j]);
                        }
                        if (out_bf16) {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(outp) + (size_t)m * ldo + n0 + c;
#pragma unroll
                            for (int j = 0; j < 32; j += 8)
                                *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                                              pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
                        } else {
                            float* o = reinterpret_cast<float*>(outp) + (size_t)m * ldo + n0 + c;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        }
                    } else {
                        // ragged N tail (e.g. the 4-channel output conv): scalar path
#pragma unroll
                        for (int j = 0; j < 32; ++j) {                // static indices: f[] must stay in registers
                            if (j >= nvalid) break;
                            float x = f[j];
                            const int n = n0 + c + j;
                            if (bias) x += bias[n];
                            if (rb) x += rb[n];
                            if (resid) x += resid[(size_t)m * p.ldr + n];
                            if (p.epi == EPI_GELU && !part) x = gelu_erf_f(x);
                            if (out_bf16)
                                reinterpret_cast<__nv_bfloat16*>(outp)[(size_t)m * ldo + n] = __float2bfloat16(x);
                            else
                                reinterpret_cast<float*>(outp)[(size_t)m * ldo + n] = x;
                        }
                    }
                }
            }
        }
        tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BN, int EPI_WARPS = 4>
static int launch(const GemmKernelArgs& a, int m_tiles, int n_tiles, cudaStream_t st) {
    constexpr int THREADS = 64 + 32 * EPI_WARPS;
    using S = GemmSmem<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        C2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    // Ring depth: a grid that puts at most one CTA on an SM has the whole 227 KB to itself and is latency-bound on the
    // K loop (few tiles, deep K: the 16x16 / 8x8 / 4x4 levels), so it gets as many stages as fit; otherwise stay within
    // half an SM so that two CTAs are co-resident and one's epilogue overlaps the other's main loop.
    GemmKernelArgs b = a;
    const int ctas = m_tiles * n_tiles * a.splits;
    const int budget = (ctas <= 148 ? 227 : 113) * 1024 - 1024;      // 1 KB per CTA is reserved by the system
    int stages = (budget - 256 - S::VEC_BYTES - 1024) / S::STAGE_BYTES;
    const int iters = (a.taps * a.k_chunks + a.splits - 1) / a.splits;
    if (stages > iters) stages = iters;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const int min_stages = a.epi == EPI_GEGLU ? 2 : MIN_STAGES;       // the GEGLU epilogue stores straight from registers
    if (stages < min_stages) stages = min_stages;
    b.stages = stages;
    C2V_CHECK_CUDA(launch(gemm_tc_kernel<BN, EPI_WARPS>, dim3(m_tiles, n_tiles, a.splits), dim3(THREADS), S::total(stages), st, b));
    return OK;
}

// out = sum_z ws[z] + bias + rowbias + residual  (deterministic split-K reduction, fused epilogue)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, const float* __restrict__ bias,
                                                            const float* __restrict__ rowbias, int rows_per_group,
                                                            const float* __restrict__ residual, int ldr, void* __restrict__ out, int ldo,
                                                            int out_bf16) {
    pdl_entry();
    const int nv = N >> 2;
    const size_t plane = (size_t)M * N;
    const long long total = (long long)M * nv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / nv);
        const int n = (int)(i - (long long)m * nv) * 4;
        // every operand of the element is requested before the first add: one memory round trip per element (splits <= 8)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 y[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) y[z] = z < splits ? *reinterpret_cast<const float4*>(ws + z * plane + (size_t)m * N + n) : z4;
        const float4 b0 = bias ? *reinterpret_cast<const float4*>(bias + n) : z4;
        const float4 b1 = rowbias ? *reinterpret_cast<const float4*>(rowbias + (size_t)(m / rows_per_group) * N + n) : z4;
        const float4 b2 = residual ? *reinterpret_cast<const float4*>(residual + (size_t)m * ldr + n) : z4;
        float4 x = y[0];
#pragma unroll
        for (int z = 1; z < 8; ++z) {
            if (z < splits) {
                x.x += y[z].x; x.y += y[z].y; x.z += y[z].z; x.w += y[z].w;
            }
        }
        if (bias) {
            x.x += b0.x; x.y += b0.y; x.z += b0.z; x.w += b0.w;
        }
        if (rowbias) {
            x.x += b1.x; x.y += b1.y; x.z += b1.z; x.w += b1.w;
        }
        if (residual) {
            x.x += b2.x; x.y += b2.y; x.z += b2.z; x.w += b2.w;
        }
        if (out_bf16)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)m * ldo + n) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
        else
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)m * ldo + n) = x;
    }
}

int splitk_reduce_launch(const float* ws, int splits, int M, int N, const float* bias, const float* rowbias, int rows_per_group,
                         const float* residual, int ldr, void* out, int ldo, int out_bf16, cudaStream_t st) {
    if (splits > 8) return ERR_UNSUPPORTED;
    long long work = (long long)M * (N >> 2);
    int grid = (int)((work + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    C2V_CHECK_CUDA(launch(splitk_reduce_kernel, dim3(grid), dim3(256), 0, st, ws, splits, M, N, bias, rowbias, rows_per_group, residual, ldr, out, ldo,
                          out_bf16));
    return OK;
}

int gemm_tc_launch(const GemmKernelArgs& a, int bn, int m_tiles, int n_tiles, cudaStream_t st) {
    if (a.epi == EPI_GEGLU) {      // two epilogue warp groups for the erf epilogue (-10 % on the 32x32-level FF projection)
        switch (bn) {
            case 64: return launch<64, 8>(a, m_tiles, n_tiles, st);
            case 128: return launch<128, 8>(a, m_tiles, n_tiles, st);
            case 160: return launch<160, 8>(a, m_tiles, n_tiles, st);
            case 256: return launch<256, 8>(a, m_tiles, n_tiles, st);
            default: return ERR_UNSUPPORTED;
        }
    }
    switch (bn) {
        case 64: return launch<64>(a, m_tiles, n_tiles, st);
        case 128: return launch<128>(a, m_tiles, n_tiles, st);
        case 160: return launch<160>(a, m_tiles, n_tiles, st);
        case 256: return launch<256>(a, m_tiles, n_tiles, st);
        default: return ERR_UNSUPPORTED;
    }
}

}  // namespace c2v
