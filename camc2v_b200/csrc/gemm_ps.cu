// Persistent tcgen05 GEMM for the multi-wave projections of the UNet step: the GEGLU projections (attention.py:431-438) and the
// fused q|k|v / q projections (attention.py:58-62; epipolar.py:55-63) of the 32x32 and 16x16 levels - plain [M, K] x [N, K]^T
// products with a 16-bit output, no residual, and several times more output tiles than the machine has SMs.
//
// gemm_tc.cu runs ONE tile per CTA (two CTAs per SM): a tile's epilogue warps sleep through its main loop, its tensor pipe idles
// through its epilogue, and every tile pays barrier init + TMEM allocation (ncu, round 2: 26 % of the GEGLU kernel's stall samples
// are epilogue warps waiting for the accumulator, the tensor pipe is 24 % active).  Here one CTA per SM walks the tile list
//     tile = blockIdx.x + i * gridDim.x      (n fastest: the CTAs of a wave share A row blocks and all of B through L2)
// with the shared-memory ring running straight through tile boundaries and the fp32 accumulator DOUBLE-BUFFERED in tensor memory
// (2 x BN columns): the MMA warp fills buffer (i+1)&1 while the epilogue warps drain buffer i&1, so the main loop of tile i+1,
// the operand traffic of tile i+2 and the epilogue of tile i overlap inside one CTA.
//
// B-RESIDENT mode (short K: the 32x32 level, K = 320, and the 16x16-level q|k|v, K = 640).  These products are bound by L2 -> shared
// memory operand traffic, not by the tensor pipe: a 128 x 256 x 320 tile moves 240 KB of operands for 1.3 us of MMA.  When the whole
// K extent of a weight tile fits in shared memory next to a small A ring, the CTA loads its [BN, K] weight tile ONCE and then walks
// M: CTAs are grouped per N tile (P = #SM / n_tiles CTAs per group, CTA r of a group takes M tiles r, r + P, ...), so per output tile
// only the 128 x K activation block is fetched (80 KB instead of 240 KB at K = 320, N tile 256).  For the linear epilogue the N tile
// is chosen here (up to 256 wide, independent of gemm_tc's tile) to minimise the number of passes over A.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = epilogue (4 for the linear epilogue, 8 for the
// erf GEGLU epilogue: two warps per TMEM lane quarter, lower / upper half of the tile's output columns).  A lane owns an accumulator
// ROW, so both epilogues stage their 16-bit results per warp in shared memory and write them back with several lanes per row.
#include "common.cuh"
#include "gemm_tc.h"

namespace c2v {

namespace {
constexpr int PS_BM = 128;
constexpr int PS_BK = 64;
constexpr int PS_MAX_STAGES = 8;
constexpr int PS_STAGE_OUT = 4 * 4096;        // staged output: 4 KB per warp (linear epilogue, 4 warps) / 2 KB per warp (GEGLU, 8 warps)
}  // namespace

// P == 0: streaming mode (A and B tiles through the ring); P > 0: B-resident mode with P CTAs per N tile.
template <int BN, int EPI_WARPS, bool GEGLU, bool BRES>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1) gemm_ps_kernel(const __grid_constant__ GemmKernelArgs p, int m_tiles, int n_tiles, int P) {
    constexpr int A_BYTES = PS_BM * PS_BK * 2;
    constexpr int B_BYTES = BN * PS_BK * 2;
    constexpr int STAGE_BYTES = BRES ? A_BYTES : A_BYTES + B_BYTES;
    constexpr uint32_t ACC_STRIDE = BN <= 128 ? 128 : 256;        // TMEM columns between the two accumulators
    constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.stages;
    const int k_iters = p.k_chunks;
    uint8_t* ring = smem + (BRES ? k_iters * B_BYTES : 0);        // B-resident: [k_iters][BN x 64] weight tile first, then the A ring
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + PS_MAX_STAGES;
    uint64_t* acc_full = empty_bar + PS_MAX_STAGES;               // [2] MMA -> epilogue: accumulator b is complete
    uint64_t* acc_empty = acc_full + 2;                           // [2] epilogue -> MMA: accumulator b has been read out
    uint64_t* b_bar = acc_empty + 2;                              // B-resident: the weight tile has landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_bar + 1);
    uint8_t* stage_out = reinterpret_cast<uint8_t*>(full_bar) + 256;       // epilogue staging (PS_STAGE_OUT bytes)

    const int warp = threadIdx.x >> 5;
    // tile walk of this CTA: (first, step, end) over a linear index that is the tile number (streaming) or the M tile (B-resident)
    const int t_first = BRES ? (int)(blockIdx.x % P) : (int)blockIdx.x;
    const int t_step = BRES ? P : (int)gridDim.x;
    const int t_end = BRES ? m_tiles : m_tiles * n_tiles;
    const int n_fixed = BRES ? (int)(blockIdx.x / P) : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], EPI_WARPS);
        }
        mbar_init(b_bar, 1);
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
    pdl_entry();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            if (BRES) {
                mbar_expect_tx(b_bar, (uint32_t)(k_iters * B_BYTES));
                for (int kc = 0; kc < k_iters; ++kc) tma_load_2d(smem + kc * B_BYTES, &p.tmB, b_bar, kc * PS_BK, n_fixed * BN);
            }
            for (int tile = t_first; tile < t_end; tile += t_step) {
                const int m_tile = BRES ? tile : tile / n_tiles, n_tile = BRES ? n_fixed : tile - m_tile * n_tiles;
                for (int kc = 0; kc < k_iters; ++kc) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* a_dst = ring + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(a_dst, &p.tmA, &full_bar[s], kc * PS_BK, m_tile * PS_BM);
                    if (!BRES) tma_load_2d(a_dst + A_BYTES, &p.tmB, &full_bar[s], kc * PS_BK, n_tile * BN);
                    if (++s == STAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(PS_BM, BN, 0, 0);
        int s = 0;
        uint32_t ph = 0;
        int i = 0;
        if (BRES) mbar_wait(b_bar, 0);
        for (int tile = t_first; tile < t_end; tile += t_step, ++i) {
            const int b = i & 1;
            mbar_wait(&acc_empty[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);      // first use of each buffer passes at once
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)b * ACC_STRIDE;
            for (int kc = 0; kc < k_iters; ++kc) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(ring + s * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a_addr);
                    const uint64_t bdesc = umma_desc_sw128(BRES ? smem_u32(smem + kc * B_BYTES) : a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < PS_BK / 16; ++k) umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
                    umma_commit(&empty_bar[s]);
                    if (kc == k_iters - 1) umma_commit(&acc_full[b]);
                }
                __syncwarp();
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int lg = warp & 3;                       // TMEM lane quarter this warp may access (lane = accumulator row)
        const int grp = (warp - 2) >> 2;               // 0 / 1: which of the two warps of this lane quarter (EPI_WARPS == 8)
        int i = 0;
        for (int tile = t_first; tile < t_end; tile += t_step, ++i) {
            const int m_tile = BRES ? tile : tile / n_tiles, n_tile = BRES ? n_fixed : tile - m_tile * n_tiles;
            const int b = i & 1;
            const int n0 = n_tile * BN;
            mbar_wait(&acc_full[b], (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)b * ACC_STRIDE;
            // 16-column chunks: tcgen05.ld -> wait -> compute -> stage in shared memory -> coalesced write-back.  (A register double
            // buffer that keeps the tcgen05.ld of chunk k+1 in flight under the math of chunk k, 16 instead of 8 epilogue warps and
            // four chunks per wait were all measured and gave nothing: profiles/r02i_ab_ps_variants.txt, r02o_ab_*.)
            if constexpr (GEGLU) {
                constexpr int HALF = BN / 2;
                const float* bias_x = p.bias ? p.bias + n0 : nullptr;
                auto fetch = [&](int c, uint32_t (&xv)[16], uint32_t (&gv)[16], float4 (&bx)[4], float4 (&bg)[4]) {
                    tmem_ld16(trow + c, xv);
                    tmem_ld16(trow + HALF + c, gv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        bx[j] = bias_x ? *reinterpret_cast<const float4*>(bias_x + c + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        bg[j] = bias_x ? *reinterpret_cast<const float4*>(bias_x + HALF + c + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                // Staged write-back (see the linear epilogue below): the two warps of a lane quarter take the lower / upper half of the
                // tile's output chunks; a warp stages two chunks (32 columns = 64 B per row, XOR-swizzled 16-byte units) in its own 2 KB
                // and writes them back with 4 lanes per row - 8 rows x 64 B per store instruction instead of 32 rows x 32 B.
                constexpr int NCH = HALF / 16, C0 = (NCH + 1) / 2;
                const int ch_lo = grp == 0 ? 0 : C0, ch_hi = grp == 0 ? C0 : NCH;
                uint8_t* wst = stage_out + (warp - 2) * 2048;
                const int ln = lane_id();
                __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(m_tile * PS_BM + lg * 32) * p.ldo + n_tile * HALF;
                const int rows_ok = min(32, p.M - (m_tile * PS_BM + lg * 32));
                auto stage = [&](int slot, const uint32_t (&xv)[16], const uint32_t (&gv)[16], const float4 (&bx)[4], const float4 (&bg)[4]) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 x01 = fadd2(make_float2(__uint_as_float(xv[4 * j]), __uint_as_float(xv[4 * j + 1])), make_float2(bx[j].x, bx[j].y));
                        const float2 g01 = fadd2(make_float2(__uint_as_float(gv[4 * j]), __uint_as_float(gv[4 * j + 1])), make_float2(bg[j].x, bg[j].y));
                        const float2 x23 = fadd2(make_float2(__uint_as_float(xv[4 * j + 2]), __uint_as_float(xv[4 * j + 3])), make_float2(bx[j].z, bx[j].w));
                        const float2 g23 = fadd2(make_float2(__uint_as_float(gv[4 * j + 2]), __uint_as_float(gv[4 * j + 3])), make_float2(bg[j].z, bg[j].w));
                        const float2 y01 = geglu_fast2(x01, g01), y23 = geglu_fast2(x23, g23);
                        pk[2 * j] = pack_bf16(y01.x, y01.y);
                        pk[2 * j + 1] = pack_bf16(y23.x, y23.y);
                    }
                    const int sw = (ln >> 1) & 3;
                    *reinterpret_cast<uint4*>(wst + ln * 64 + (((2 * slot) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(wst + ln * 64 + (((2 * slot + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                };
                uint32_t xa[16], ga[16];
                float4 bxa[4], bga[4];
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ch += 2) {
                    const bool two = ch + 1 < ch_hi;
                    fetch(ch * 16, xa, ga, bxa, bga);
                    tmem_ld_wait();
                    stage(0, xa, ga, bxa, bga);
                    if (two) {
                        fetch(ch * 16 + 16, xa, ga, bxa, bga);
                        tmem_ld_wait();
                        stage(1, xa, ga, bxa, bga);
                    }
                    __syncwarp();
                    const int u = ln & 3;
                    if (u < (two ? 4 : 2)) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int rr = it * 8 + (ln >> 2);
                            if (rr < rows_ok) {
                                const uint4 val = *reinterpret_cast<const uint4*>(wst + rr * 64 + ((u ^ ((rr >> 1) & 3)) << 4));
                                *reinterpret_cast<uint4*>(obase + (size_t)rr * p.ldo + ch * 16 + u * 8) = val;
                            }
                        }
                    }
                    __syncwarp();
                }
            } else {
                // Linear epilogue (EPI_WARPS == 4: one warp per TMEM lane quarter).  A lane owns a ROW of the accumulator, so storing
                // from registers sends 32 different 128-byte lines per warp instruction, 16-32 bytes each (ncu on the first version:
                // 26 % of the kernel's stall samples are the epilogue waiting to re-use registers that queued stores still hold).
                // Instead the warp stages 64 columns (128 bytes per row, XOR-swizzled 16-byte units: conflict-free both ways) in its
                // own 4 KB of shared memory and writes them back with 8 lanes per row: 4 full lines per store instruction.
                static_assert(GEGLU || EPI_WARPS == 4, "the staged linear epilogue assumes one warp per lane quarter");
                const float* bias_n = p.bias ? p.bias + n0 : nullptr;
                const bool gelu = p.epi == EPI_GELU;
                uint8_t* wst = stage_out + (warp - 2) * 4096;
                const int ln = lane_id();
                const int rsw = ln & 7;                                     // this lane's row (= lane) modulo 8: swizzle key of its writes
                __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(m_tile * PS_BM + lg * 32) * p.ldo + n0;
                const int rows_ok = min(32, p.M - (m_tile * PS_BM + lg * 32));      // valid rows of this warp's slice (<= 0: none)
#pragma unroll 1
                for (int c = 0; c < BN; c += 64) {
                    uint32_t v[4][16];
                    float4 bv[4][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (c + 16 * q < BN) {
                            tmem_ld16(trow + c + 16 * q, v[q]);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                bv[q][j] = bias_n ? *reinterpret_cast<const float4*>(bias_n + c + 16 * q + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (c + 16 * q < BN) {
                            float f[16];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[4 * j] = bv[q][j].x + __uint_as_float(v[q][4 * j]);
                                f[4 * j + 1] = bv[q][j].y + __uint_as_float(v[q][4 * j + 1]);
                                f[4 * j + 2] = bv[q][j].z + __uint_as_float(v[q][4 * j + 2]);
                                f[4 * j + 3] = bv[q][j].w + __uint_as_float(v[q][4 * j + 3]);
                            }
                            if (gelu) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) f[j] = gelu_erf_f(f[j]);
                            }
                            *reinterpret_cast<uint4*>(wst + ln * 128 + (((2 * q) ^ rsw) << 4)) =
                                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                            *reinterpret_cast<uint4*>(wst + ln * 128 + (((2 * q + 1) ^ rsw) << 4)) =
                                make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
                        }
                    }
                    __syncwarp();
                    const int u = ln & 7;                                   // 16-byte unit of the row this lane writes back
                    if (c + u * 8 < BN) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rr = it * 4 + (ln >> 3);
                            if (rr < rows_ok) {
                                const uint4 val = *reinterpret_cast<const uint4*>(wst + rr * 128 + ((u ^ (rr & 7)) << 4));
                                *reinterpret_cast<uint4*>(obase + (size_t)rr * p.ldo + c + u * 8) = val;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // this warp has read its part of accumulator b: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&acc_empty[b]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

constexpr int PS_SMEM_BUDGET = 226 * 1024;     // of the 227 KB an sm_100 CTA may use
constexpr int PS_SMEM_BASE = 256 + 1024;       // barriers / TMEM pointer + 1024-byte alignment slack
constexpr int ps_fixed(bool) { return PS_SMEM_BASE + PS_STAGE_OUT; }      // + the epilogue's staging (both epilogues)      // + the linear epilogue's staging

template <int BN, int EPI_WARPS, bool GEGLU, bool BRES>
static int ps_launch(const GemmKernelArgs& a, const PsPlan& pl, int m_tiles, cudaStream_t st) {
    constexpr int A_BYTES = PS_BM * PS_BK * 2, B_BYTES = BN * PS_BK * 2;
    constexpr int STAGE_BYTES = BRES ? A_BYTES : A_BYTES + B_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        C2V_CHECK_CUDA(cudaFuncSetAttribute(gemm_ps_kernel<BN, EPI_WARPS, GEGLU, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GemmKernelArgs b = a;
    const int resident = BRES ? a.k_chunks * B_BYTES : 0;
    int stages = (PS_SMEM_BUDGET - ps_fixed(GEGLU) - resident) / STAGE_BYTES;
    if (stages > PS_MAX_STAGES) stages = PS_MAX_STAGES;
    if (stages < 2) return ERR_UNSUPPORTED;
    b.stages = stages;
    const int n_tiles = a.N / BN;
    const int tiles = m_tiles * n_tiles;
    const int grid = BRES ? pl.P * n_tiles : (tiles < sm_count() ? tiles : sm_count());
    C2V_CHECK_CUDA(launch(gemm_ps_kernel<BN, EPI_WARPS, GEGLU, BRES>, dim3(grid), dim3(64 + 32 * EPI_WARPS),
                          (size_t)resident + (size_t)stages * STAGE_BYTES + ps_fixed(GEGLU), st, b, m_tiles, n_tiles, pl.P));
    return OK;
}

// Which persistent form (if any) a GEMM takes.  Base conditions: plain A, one tap, no split-K, 16-bit output, no residual / row
// bias, full 128-row A boxes, at least `min_tiles` output tiles of the default width (below that the one-tile-per-CTA kernel
// already fits the machine in one wave).  C2V_GEMM_PS (A/B switch): 0 = off, 1 = streaming only, 2 = B-resident where it fits (default).
PsPlan gemm_ps_plan(const GemmKernelArgs& a, int bn_default) {
    static const int mode = [] { const char* e = getenv("C2V_GEMM_PS"); return e ? atoi(e) : 2; }();
    PsPlan none{0, 0, 0};
    if (mode <= 0) return none;
    if (a.a_mode != A_PLAIN || a.taps != 1 || a.splits != 1 || !a.out_bf16 || a.residual || a.rowbias) return none;
    if (a.tile_rows != PS_BM || (a.ldo % 8) != 0) return none;
    const int m_tiles = (a.M + PS_BM - 1) / PS_BM;
    if (a.N % bn_default != 0 || m_tiles * (a.N / bn_default) < 300) return none;
    const int sms = sm_count();
    PsPlan best = none;
    if (mode >= 2) {
        // B-resident candidates: GEGLU weights are interleaved per N tile of the default width, so that width is fixed
        const int cands_lin[] = {256, 240, 192, 160, 128};
        const int cands_geglu[] = {bn_default};
        const int* cands = a.epi == EPI_GEGLU ? cands_geglu : cands_lin;
        const int nc = a.epi == EPI_GEGLU ? 1 : 5;
        double best_cost = 1e30;
        for (int ci = 0; ci < nc; ++ci) {
            const int bn = cands[ci];
            if (bn != 128 && bn != 160 && bn != 192 && bn != 240 && bn != 256) continue;
            if (a.N % bn != 0) continue;
            const int resident = a.k_chunks * bn * PS_BK * 2;
            const int stages = (PS_SMEM_BUDGET - ps_fixed(a.epi == EPI_GEGLU) - resident) / (PS_BM * PS_BK * 2);
            if (stages < 3) continue;
            const int n_tiles = a.N / bn;
            if (n_tiles > sms) continue;
            int P = sms / n_tiles;
            if (P > m_tiles) P = m_tiles;
            const int per_cta = (m_tiles + P - 1) / P;
            // per tile: the A block through L2 (~45 KB/us per SM) against the MMA (128 x bn x 64 per k chunk at 8192 flop/clk, 1.9 GHz)
            const double t_a = a.k_chunks * 16.0 / 45.0, t_mma = a.k_chunks * (2.0 * PS_BM * bn * PS_BK) / (8192.0 * 1900.0);
            const double cost = per_cta * (t_a > t_mma ? t_a : t_mma) + resident / 1024.0 / 45.0;
            if (cost < best_cost) {
                best_cost = cost;
                best = PsPlan{2, bn, P};
            }
        }
        if (best.mode) return best;
    }
    if (bn_default != 128 && bn_default != 160 && bn_default != 256) return none;
    return PsPlan{1, bn_default, 0};
}

template <int BN>
static int ps_dispatch(const GemmKernelArgs& a, const PsPlan& pl, int m_tiles, cudaStream_t st) {
    if (a.epi == EPI_GEGLU)
        return pl.mode == 2 ? ps_launch<BN, 8, true, true>(a, pl, m_tiles, st) : ps_launch<BN, 8, true, false>(a, pl, m_tiles, st);
    return pl.mode == 2 ? ps_launch<BN, 4, false, true>(a, pl, m_tiles, st) : ps_launch<BN, 4, false, false>(a, pl, m_tiles, st);
}

int gemm_ps_launch(const GemmKernelArgs& a, const PsPlan& pl, cudaStream_t st) {
    const int m_tiles = (a.M + PS_BM - 1) / PS_BM;
    switch (pl.bn) {
        case 128: return ps_dispatch<128>(a, pl, m_tiles, st);
        case 160: return ps_dispatch<160>(a, pl, m_tiles, st);
        case 192: return ps_dispatch<192>(a, pl, m_tiles, st);
        case 240: return ps_dispatch<240>(a, pl, m_tiles, st);
        case 256: return ps_dispatch<256>(a, pl, m_tiles, st);
        default: return ERR_UNSUPPORTED;
    }
}

}  // namespace c2v
