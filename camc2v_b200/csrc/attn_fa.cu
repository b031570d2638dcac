// Flash attention on tcgen05 tensor cores with P kept in tensor memory: head dim 64, 16-bit operands, fp32 softmax.
//   replaces xformers.ops.memory_efficient_attention (R/lvdm/modules/attention.py:177,189), the einsum softmax path
//   (attention.py:105-129) and F.scaled_dot_product_attention with the boolean epipolar mask (R/model/modules/epipolar.py:99).
//
// Round-2 rebuild of the round-1 pipeline (profiles/r01h_ncu_full.txt: ~2050 cycles per 64-key tile against ~256 cycles of MMA; four
// softmax warps made two TMEM passes per tile - max + masked-score write-back, then exp - P went through shared memory with a
// proxy fence, and softmax(j+1) waited for PV(j)).  One CTA = 128 query rows of one (batch, head), 64-key tiles (two image rows of
// a 32x32 latent frame keep the epipolar tile map fine-grained), two CTAs per SM.  Per tile:
//     S = Q K^T          tcgen05.mma 128x64x16 (x4), operands K-major in 128B-swizzled smem (TMA); S double-buffered in TMEM
//     P_h = softmax      the two 32-key HALVES of a tile are two INDEPENDENT online-softmax streams, each with its own reference
//                        maximum, row sum and O accumulator (split-K style): a stream never needs the other half's maximum, so each
//                        half gets its own warp - 8 softmax warps per CTA, 4 per SM sub-partition with two CTAs per SM, which is what
//                        keeps the MUFU (exp2) pipe busy: with head dim 64 the kernel is exp2 / issue bound, not MMA bound
//                        (8192 exp2 per tile = 512 cycles of the SM's 16-lane MUFU against 256 cycles of MMA).
//                        One pass: scores are read from TMEM once, the mask is applied in registers (a packed 32-bit word per row and
//                        32-key chunk; masked scores become -inf, nothing is written back), then max, exp2, sum.
//                        LAZY rescale (FA-4 style): the reference maximum of a stream is raised only when a tile exceeds it by more
//                        than 2^8 (probabilities are then <= 256: exact in bf16 / fp16 and in the fp32 sums), so the O accumulator is
//                        touched by the softmax warps only in those rare tiles and nothing waits for PV(j-1) on the common path.
//                        P stays in TENSOR MEMORY: the 16-bit probabilities are stored (tcgen05.st) over the first 16 of the 32 S
//                        columns they came from - no shared-memory P tile, no fence.proxy.async.
//     O_h += P_h V_h     tcgen05.mma with the A operand FROM TMEM (128x64x16, x2 per half), V as MN-major operand straight from its
//                        [key, d] layout.  QK(j+2) overwrites S buffer j&1 (= P(j)) only after PV(j): the tensor pipe executes one
//                        thread's MMAs in issue order.
// Epilogue: out = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_h = 2^(m_h - max(m_0, m_1)); the (m, l) pairs of a row are exchanged
// once through shared memory, each thread of the pair writes half of the head dim.
// TMEM: 2 x 64 S/P columns + 2 x 64 O columns = 256.  Packed-mask words are fetched one tile ahead.
//
// Mask forms (all produce the same 32-bit word per (row, 32-key chunk), so outputs are bit-identical across forms):
//   MODE 0  dense / ragged tails / register-token segment, and the per-sample packed epipolar mask (c2v_epipolar_bitmask)
//   MODE 1  materialised byte mask in the reference's format, and the exact epipolar predicate on arbitrary grids (rolled loop)
//   MODE 2  the exact epipolar predicate on square power-of-two grids, evaluated in-kernel from the 3x3 fundamental matrices with
//           the reference's fp32 operation order (R/model/camcontexti2v.py:229-239)
#include "attn.h"
#include "common.cuh"

namespace c2v {

constexpr int FA_BM = 128;   // query rows per CTA
constexpr int FA_BN = 64;    // keys per tile
constexpr int FA_D = 64;
#ifndef C2V_FA_KV_STAGES
#define C2V_FA_KV_STAGES 4
#endif
#ifndef C2V_FA_MIN_CTAS
#define C2V_FA_MIN_CTAS 2
#endif
// debug builds only (-DC2V_FA_TIMING=1): warp 2 of a few CTAs prints its per-phase cycle totals (clock64) at the end
#ifndef C2V_FA_TIMING
#define C2V_FA_TIMING 0
#endif
#if C2V_FA_TIMING
#define FA_T(i) do { const long long _t = clock64(); tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define FA_T(i) do { } while (0)
#endif
constexpr int FA_KV_STAGES = C2V_FA_KV_STAGES;
// threads per query row: 2 = each of the two 32-key halves of a tile has its own warp (8 softmax warps, 4 per SM sub-partition with
// two CTAs per SM: the softmax is MUFU / latency bound, so the extra warps are what keeps the exp2 pipe busy); 1 = one thread per row
// processes both halves one after the other (4 softmax warps).  Measured on one box (dense 16x1024x1024x5 heads / 32x32 epipolar
// layer): 2 -> 55.3 / 268 us, 1 -> 63.6 / 330 us.  p_full arrivals: one elected lane per warp (per-thread arrivals measured the same).
#ifndef C2V_FA_SPLIT
#define C2V_FA_SPLIT 2
#endif
#define FA_SPLIT C2V_FA_SPLIT
constexpr int FA_NH = 2 / FA_SPLIT;                  // key halves per thread
constexpr int FA_THREADS = 64 + 128 * FA_SPLIT;
constexpr float FA_TAU = 8.0f;                       // lazy-rescale threshold (log2 units): P <= 2^8

constexpr int FA_Q_BYTES = FA_BM * FA_D * 2;         // 16 KB
constexpr int FA_K_BYTES = FA_BN * FA_D * 2;         // 8 KB
constexpr int FA_V_BYTES = FA_BN * FA_D * 2;         // 8 KB
constexpr int FA_OFF_Q = 0;
constexpr int FA_OFF_K = FA_OFF_Q + FA_Q_BYTES;
constexpr int FA_OFF_V = FA_OFF_K + FA_KV_STAGES * FA_K_BYTES;
constexpr int FA_OFF_BAR = FA_OFF_V + FA_KV_STAGES * FA_V_BYTES;
constexpr int FA_OFF_LIST = FA_OFF_BAR + 256;
constexpr int FA_MAX_TILES = 1024;
constexpr int FA_OFF_XCH = FA_OFF_LIST + FA_MAX_TILES * 2;       // (m, l) of the two streams of a row, exchanged once in the epilogue
constexpr int FA_SMEM = FA_OFF_XCH + 2 * FA_BM * 8;

constexpr uint32_t FA_TM_S = 0;                      // S buffers: 2 x 64 fp32 columns; P_h(j) = first 16 columns of chunk h of S(j)
constexpr uint32_t FA_TM_O = 2 * FA_BN;              // O accumulators of the two key-half streams: 2 x 64 columns
constexpr uint32_t FA_TMEM_COLS = 256;

constexpr uint32_t FA_NEG_INF = 0xff800000u;

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P, 16-bit, K-major: lane = row, one 32-bit column = two consecutive k) is read
// straight from tensor memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

struct FaLine {
    float l0, l1, l2;
};
// Normalised epipolar line of query pixel (xi, yi) in frame t2 (camcontexti2v.py:229-236); same arithmetic as epi_maps.cu / epipolar.cu
__device__ __forceinline__ FaLine fa_epi_line(const float* __restrict__ f, float xi, float yi) {
    float a0 = __fmaf_rn(f[2], 1.0f, __fmaf_rn(f[1], yi, __fmul_rn(f[0], xi)));
    float a1 = __fmaf_rn(f[5], 1.0f, __fmaf_rn(f[4], yi, __fmul_rn(f[3], xi)));
    float a2 = __fmaf_rn(f[8], 1.0f, __fmaf_rn(f[7], yi, __fmul_rn(f[6], xi)));
    const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
    FaLine l;
    l.l0 = __fdiv_rn(a0, nrm);
    l.l1 = __fdiv_rn(a1, nrm);
    l.l2 = __fdiv_rn(a2, nrm);
    return l;
}

// max of 32 scores held as raw bits: four independent chains of 3-input maxima (FMNMX3)
__device__ __forceinline__ float fa_max32(const uint32_t (&v)[32]) {
    float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        a0 = fmaxf(a0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        a1 = fmaxf(a1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        a2 = fmaxf(a2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
        a3 = fmaxf(a3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
    }
    return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}

template <int LOGW, int D, int MODE>
__global__ void __launch_bounds__(FA_THREADS, C2V_FA_MIN_CTAS) attn_fa_kernel(const __grid_constant__ AttnKernelArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;     // [FA_KV_STAGES <= 4]
    uint64_t* kv_empty = bars + 5;    // [FA_KV_STAGES <= 4]
    uint64_t* s_full = bars + 9;      // [2]     S(j) is in TMEM buffer j & 1
    uint64_t* p_full = bars + 11;     // [2][2]  P_h(j) has been stored over chunk h of S(j): index (j & 1) * 2 + h
    uint64_t* pv_done = bars + 15;    // [2]     one phase per PV_h(j)
    uint64_t* o_final = bars + 17;    // the last PV
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);
    int* n_act_s = reinterpret_cast<int*>(bars + 19);

    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int bkv = b / p.kv_div;
    const int n_main = (p.lk + FA_BN - 1) / FA_BN;
    const int n_tiles = n_main + (p.lk2 > 0 ? 1 : 0);       // last tile = register-token segment

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("camc2v_b200: attention smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&p.tmQ);
        tma_prefetch_desc(&p.tmK);
        tma_prefetch_desc(&p.tmV);
        if (p.lk2 > 0) {
            tma_prefetch_desc(&p.tmK2);
            tma_prefetch_desc(&p.tmV2);
        }
        mbar_init(q_full, 1);
        for (int s = 0; s < FA_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&pv_done[i], 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);      // one elected lane of each of the 4 warps that own a chunk
        mbar_init(o_final, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, FA_TMEM_COLS);
        tmem_relinquish();
    }
    pdl_entry();      // CTA-local set-up above overlaps the previous kernel's tail; global memory is touched only below
    // CTA -> (query tile, head).  With an epipolar tile map the CTAs are issued heaviest query tile first (longest-processing-time
    // order, all heads of a tile back to back): the work per query tile varies 2-3x with the number of key tiles visited and the
    // grid is only ~2 waves deep, so in-order dispatch would otherwise leave a long tail.
    int q_tile = blockIdx.x, head = blockIdx.y;
    if (p.tile_map) {
        const int lin = blockIdx.x + gridDim.x * blockIdx.y;
        head = lin % gridDim.y;
        q_tile = (int)p.tile_map[((size_t)b * gridDim.x + lin / gridDim.y) * p.tile_map_words + (p.tile_map_words - 1)];
    }
    const int q0 = q_tile * FA_BM;
    // key tiles this query tile visits (epipolar tile map: tiles that cannot contain an unmasked pair are never loaded)
    uint16_t* tile_list = reinterpret_cast<uint16_t*>(smem + FA_OFF_LIST);
    if (warp == 2) {
        const uint32_t* map = p.tile_map ? p.tile_map + ((size_t)b * gridDim.x + q_tile) * p.tile_map_words : nullptr;
        int cnt = 0;
        for (int j0 = 0; j0 < n_tiles; j0 += 32) {
            const int j = j0 + lane_id();
            bool act = j < n_tiles;
            if (act && map && j < n_main) act = (map[j >> 5] >> (j & 31)) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, act);
            if (act) tile_list[cnt + __popc(bal & ((1u << lane_id()) - 1u))] = (uint16_t)j;
            cnt += __popc(bal);
        }
        if (lane_id() == 0) *n_act_s = cnt;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int n_act = *n_act_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (n_act > 0 && elect_one()) {
            mbar_expect_tx(q_full, FA_Q_BYTES);
            tma_load_3d(smem + FA_OFF_Q, &p.tmQ, q_full, head * FA_D, q0, b);
            for (int it = 0; it < n_act; ++it) {
                const int j = tile_list[it];
                const int s = it % FA_KV_STAGES;
                const uint32_t ph = (it / FA_KV_STAGES) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&kv_full[s], FA_K_BYTES + FA_V_BYTES);
                if (j < n_main) {
                    tma_load_3d(smem + FA_OFF_K + s * FA_K_BYTES, &p.tmK, &kv_full[s], head * FA_D, j * FA_BN, bkv);
                    tma_load_3d(smem + FA_OFF_V + s * FA_V_BYTES, &p.tmV, &kv_full[s], head * FA_D, j * FA_BN, bkv);
                } else {
                    tma_load_3d(smem + FA_OFF_K + s * FA_K_BYTES, &p.tmK2, &kv_full[s], head * FA_D, 0, 0);
                    tma_load_3d(smem + FA_OFF_V + s * FA_V_BYTES, &p.tmV2, &kv_full[s], head * FA_D, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (n_act > 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(FA_BM, FA_BN, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(FA_BM, FA_D, 0, 1);   // A = P from TMEM (K-major), B = V is MN-major
            const uint32_t q_addr = smem_u32(smem + FA_OFF_Q);
            auto issue_qk = [&](int j) {                  // S(j) -> TMEM buffer j & 1
                const int s = j % FA_KV_STAGES;
                mbar_wait(&kv_full[s], (j / FA_KV_STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t qd = umma_desc_sw128(q_addr);
                    const uint64_t kd = umma_desc_sw128(smem_u32(smem + FA_OFF_K + s * FA_K_BYTES));
#pragma unroll
                    for (int k = 0; k < FA_D / 16; ++k)
                        umma_bf16_ss(tmem_base + FA_TM_S + (uint32_t)(j & 1) * FA_BN, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
                    umma_commit(&s_full[j & 1]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            if (n_act > 1) issue_qk(1);
            for (int j = 0; j < n_act; ++j) {
                const int s = j % FA_KV_STAGES;
#pragma unroll
                for (int h = 0; h < 2; ++h) {             // the two 32-key halves are independent streams with their own accumulator
                    mbar_wait(&p_full[(j & 1) * 2 + h], (j >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t v_addr = smem_u32(smem + FA_OFF_V + s * FA_V_BYTES) + h * 32 * 128;
                        const uint32_t p_tmem = tmem_base + FA_TM_S + (uint32_t)(j & 1) * FA_BN + h * 32;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint64_t vd = umma_desc_sw128(v_addr + ks * 16 * 128);
                            umma_bf16_ts(tmem_base + FA_TM_O + h * FA_D, p_tmem + ks * 8, vd, idesc_pv, (j | ks) != 0);
                        }
                        umma_commit(&pv_done[h]);
                        if (h == 1) {
                            umma_commit(&kv_empty[s]);
                            if (j == n_act - 1) umma_commit(o_final);
                        }
                    }
                    __syncwarp();
                }
                // the tensor pipe executes this thread's MMAs in issue order: QK(j+2) may overwrite buffer j & 1 (= P(j)) now
                if (j + 2 < n_act) issue_qk(j + 2);
            }
        }
    } else {
        // ===================== softmax / rescale / epilogue (warps 2 .. 2 + 4 * FA_SPLIT) =====================
        // TMEM lane quarter lg (a warp may only touch lanes 32 * (warp % 4) ..), row r of the query tile; with FA_SPLIT = 2 the
        // warp pair (w, w + 4) shares rows and each warp owns one 32-key half of every tile.
        const int lg = warp & 3;
        const int half0 = (warp - 2) >> 2;                         // first (only, if FA_SPLIT == 2) half this thread processes
        const int r = lg * 32 + lane_id();
        const int qi = q0 + r;                                     // query index inside the batch
        const uint32_t t_lane = (uint32_t)(lg * 32) << 16;
        const uint32_t t_s0 = tmem_base + FA_TM_S + t_lane;
        const uint32_t t_o0 = tmem_base + FA_TM_O + t_lane;
        const bool epi = p.epi_F != nullptr;
        const bool use_words = MODE == 0 && epi && p.bitmask != nullptr;
        const unsigned char* mrow = (MODE == 1 && p.mask) ? p.mask + (size_t)b * p.mask_bstride + (size_t)min(qi, p.lq - 1) * p.lk : nullptr;
        // epipolar query geometry (MODE 1 / 2 only)
        const int HW = p.epi_H * p.epi_W;
        float xi = 0.f, yi = 0.f;
        const float* Frow = nullptr;
        if (MODE != 0 && epi) {
            const int qc = min(qi, p.lq - 1);
            const int t1 = qc / HW, pix = qc % HW;
            xi = __fadd_rn(__fmul_rn((float)(pix % p.epi_W), (float)p.epi_d), p.epi_off);
            yi = __fadd_rn(__fmul_rn((float)(pix / p.epi_W), (float)p.epi_d), p.epi_off);
            Frow = p.epi_F + ((size_t)b * p.epi_T + t1) * p.epi_T * 9;
        }
        int cur_t2 = -1;
        FaLine line = {0.f, 0.f, 0.f};
        float thr_m = 0.f;
        float l0x[MODE == 2 ? (1 << LOGW) : 1];

        // Per stream (key half h of every tile): reference maximum of the exponentials (log2 domain, raised lazily) and row sum.
        float m_ref[FA_NH], l_sum[FA_NH];
#pragma unroll
        for (int i = 0; i < FA_NH; ++i) {
            m_ref[i] = -INFINITY;
            l_sum[i] = 0.f;
        }
        const uint32_t* wbase = use_words ? p.bitmask + ((size_t)b * gridDim.x + q_tile) * (size_t)(p.lk >> 5) * FA_BM + r : nullptr;

        // Mask word of this row for 32-key chunk c of key tile jt: bit i = key (chunk base + i) is attended.
        auto chunk_word = [&](int jt, int c) -> uint32_t {
            const bool main_seg = jt < n_main;
            const int klim = main_seg ? p.lk : p.lk2;
            const int key0 = (main_seg ? jt * FA_BN : 0) + c * 32;
            if (use_words && main_seg) return __ldg(wbase + ((size_t)jt * 2 + c) * FA_BM);
            if (MODE == 2 && epi && main_seg) {
                constexpr int W = 1 << (MODE == 2 ? LOGW : 5);
                constexpr int RPC = 32 / W;
                constexpr float DF = (float)D, OFFC = (float)D * 0.5f - 0.5f;
                const int t2 = key0 >> (2 * LOGW);
                if (t2 != cur_t2) {                     // warp-uniform: once per key frame
                    cur_t2 = t2;
                    line = fa_epi_line(Frow + t2 * 9, xi, yi);
                    const float cmax = (float)(W - 1) * DF + OFFC;
                    thr_m = p.epi_thr + 1e-6f + 4e-7f * (fabsf(line.l0) * cmax + fabsf(line.l1) * cmax + fabsf(line.l2));
#pragma unroll
                    for (int x = 0; x < W; ++x) l0x[x] = __fmul_rn(line.l0, (float)x * DF + OFFC);
                }
                const int py0 = (key0 & (W * W - 1)) >> LOGW;
                float yr[RPC];
                bool maybe = false;
#pragma unroll
                for (int rr = 0; rr < RPC; ++rr) {
                    yr[rr] = (float)(py0 + rr) * DF + OFFC;
                    // the distance is linear in x: both row ends beyond threshold + margin on the same side => no key of the row passes
                    const float w0 = __fadd_rn(__fmaf_rn(line.l1, yr[rr], __fmul_rn(line.l0, OFFC)), line.l2);
                    const float w1 = __fadd_rn(__fmaf_rn(line.l1, yr[rr], __fmul_rn(line.l0, (float)(W - 1) * DF + OFFC)), line.l2);
                    maybe |= !((w0 > thr_m && w1 > thr_m) || (w0 < -thr_m && w1 < -thr_m));
                }
                uint32_t word = 0;
                if (maybe) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float wv = __fadd_rn(__fmaf_rn(line.l1, yr[i >> LOGW], l0x[i & (W - 1)]), line.l2);
                        word |= (fabsf(wv) < p.epi_thr ? 1u : 0u) << i;
                    }
                }
                return word;
            }
            if (MODE == 1 && main_seg && mrow) {
                // materialised byte mask (reference format, bool / uint8 [B, lq, lk]): 32 bytes -> one word
                uint32_t word = 0;
                if (key0 + 32 <= p.lk && (p.lk & 15) == 0) {
                    const uint4 m0 = *reinterpret_cast<const uint4*>(mrow + key0);
                    const uint4 m1 = *reinterpret_cast<const uint4*>(mrow + key0 + 16);
                    const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8)      // non-zero byte -> bit: 0xFF per byte, top bits gathered by a multiply
                        word |= (((((__vcmpne4(mw[w8], 0u) >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << (4 * w8);
                } else {
#pragma unroll 1
                    for (int i = 0; i < 32; ++i) {
                        const int key = key0 + i;
                        word |= ((key < klim && mrow[key] != 0) ? 1u : 0u) << i;
                    }
                }
                return word;
            }
            if (MODE == 1 && main_seg && epi) {
                // exact predicate on an arbitrary grid (non power-of-two, or a chunk spanning several frames): rolled loop
                uint32_t word = 0;
#pragma unroll 1
                for (int i = 0; i < 32; ++i) {
                    const int key = key0 + i;
                    bool ok = key < klim;
                    if (ok) {
                        const int t2 = key / HW;
                        if (t2 != cur_t2) {
                            cur_t2 = t2;
                            line = fa_epi_line(Frow + t2 * 9, xi, yi);
                        }
                        const int pj = key - t2 * HW;
                        const float xj = __fadd_rn(__fmul_rn((float)(pj % p.epi_W), (float)p.epi_d), p.epi_off);
                        const float yj = __fadd_rn(__fmul_rn((float)(pj / p.epi_W), (float)p.epi_d), p.epi_off);
                        const float dist = fabsf(__fadd_rn(__fmaf_rn(line.l1, yj, __fmul_rn(line.l0, xj)), line.l2));
                        ok = dist < p.epi_thr;
                    }
                    word |= (ok ? 1u : 0u) << i;
                }
                return word;
            }
            // dense keys (and the never-masked register-token segment): everything below klim
            const int nv = klim - key0;
            return nv >= 32 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
        };
        // true when chunk_word is the all-ones constant for every row (dense tile fully inside the sequence): no load, no votes
        auto chunk_dense = [&](int jt, int c) -> bool {
            const bool main_seg = jt < n_main;
            return !(main_seg && (epi || mrow)) && (main_seg ? jt * FA_BN : 0) + c * 32 + 32 <= (main_seg ? p.lk : p.lk2);
        };

        // One 32-key chunk of one stream: S columns -> mask -> max -> (rarely) raise the reference maximum and rescale this stream's
        // O accumulator -> exp2 -> 16-bit pairs stored over the first 16 of the 32 S columns just read.
        auto chunk = [&](int j, int h, int hi, uint32_t word, bool dense) {
            const uint32_t t_sc = t_s0 + (uint32_t)(j & 1) * FA_BN + h * 32;
            uint32_t w[16];
            const bool act = dense || __any_sync(0xffffffffu, word != 0u);
            if (!act) {                                               // no row of this warp attends a key of the chunk
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = 0u;
                tmem_st16(t_sc, w);
                return;
            }
            uint32_t v[32];
            tmem_ld32(t_sc, v);
            tmem_ld_wait();
            if (!dense && !__all_sync(0xffffffffu, word == 0xffffffffu)) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (word >> i) & 1u ? v[i] : FA_NEG_INF;
            }
            const float mt = fa_max32(v) * p.scale_log2;               // scale > 0; -inf stays -inf
            if (__any_sync(0xffffffffu, mt > m_ref[hi] + FA_TAU)) {   // first attended key of a row: m_ref = -inf -> true
                const float m_new = fmaxf(m_ref[hi], mt);
                const float alpha = (m_ref[hi] == -INFINITY) ? 0.f : fast_exp2(m_ref[hi] - m_new);
                l_sum[hi] *= alpha;
                m_ref[hi] = m_new;
                if (j > 0) {                                          // O_h holds PV_h(0..j-1): wait for PV_h(j-1), rescale in place
                    mbar_wait(&pv_done[h], (j - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o[16];
                        tmem_ld16(t_o0 + h * FA_D + c * 16, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st16(t_o0 + h * FA_D + c * 16, o);
                    }
                }
            }
            const float m_use = (m_ref[hi] == -INFINITY) ? 0.f : m_ref[hi];
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_use, -m_use);
            float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
                // exp2 on the MUFU pipe.  (A Cody-Waite / degree-4 polynomial 2^x on the FMA pipe for 1 or 2 of every 8 elements, FA-4
                // style, measured SLOWER here - 55 -> 61 / 68 us dense, 268 -> 307 / 344 us epipolar: with head dim 64 the softmax warps
                // are issue-bound as much as MUFU-bound - and was removed.)
                const float2 e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                if (i & 1) lb = fadd2(lb, e);
                else la = fadd2(la, e);
                w[i] = pack_bf16(e.x, e.y);
            }
            l_sum[hi] += (la.x + la.y) + (lb.x + lb.y);
            tmem_st16(t_sc, w);
        };

#if C2V_FA_TIMING
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
        const long long tstart = tlast;
#endif
        // mask words one tile ahead of their use (global loads of the packed mask)
        uint32_t wnext[FA_NH];
        bool dnext[FA_NH];
        auto fetch = [&](int it) {
            const int jt = tile_list[it];
#pragma unroll
            for (int hi = 0; hi < FA_NH; ++hi) {
                dnext[hi] = chunk_dense(jt, half0 + hi);
                wnext[hi] = dnext[hi] ? 0xffffffffu : chunk_word(jt, half0 + hi);
            }
        };
        if (n_act > 0) fetch(0);
        for (int j = 0; j < n_act; ++j) {
            uint32_t wcur[FA_NH];
            bool dcur[FA_NH];
#pragma unroll
            for (int hi = 0; hi < FA_NH; ++hi) {
                wcur[hi] = wnext[hi];
                dcur[hi] = dnext[hi];
            }
            if (j + 1 < n_act) fetch(j + 1);
            FA_T(0);
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            FA_T(1);
#pragma unroll
            for (int hi = 0; hi < FA_NH; ++hi) {
                const int h = half0 + hi;
                chunk(j, h, hi, wcur[hi], dcur[hi]);
                FA_T(2 + hi);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&p_full[(j & 1) * 2 + h]);
            }
            FA_T(4);
        }
#if C2V_FA_TIMING
        if (lane_id() == 0 && warp == 2 && blockIdx.y == 0 && blockIdx.z == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2))
            printf("fa_timing cta %d tiles %d cycles/tile: fetch %lld waitS %lld chunkA %lld chunkB %lld arrive %lld | total %lld\n", blockIdx.x,
                   n_act, tacc[0] / n_act, tacc[1] / n_act, tacc[2] / n_act, tacc[3] / n_act, tacc[4] / n_act, (clock64() - tstart) / n_act);
#endif
        // ---- epilogue: merge the two streams (different reference maxima), normalise, 16-bit -> global ----
        // out = (a0 O_0 + a1 O_1) / (a0 l_0 + a1 l_1), a_h = 2^(m_h - max(m_0, m_1)); this thread writes head-dim columns
        // [ocol0, ocol0 + FA_OCOLS) of its row.
        float mm[2], ll[2];
#if FA_SPLIT == 2
        float2* xch = reinterpret_cast<float2*>(smem + FA_OFF_XCH);
        xch[half0 * FA_BM + r] = make_float2(m_ref[0], l_sum[0]);
        named_bar_sync(1 + lg, 64);                                   // the warp pair that shares these rows
        const float2 oth = xch[(half0 ^ 1) * FA_BM + r];
        mm[half0] = m_ref[0];
        ll[half0] = l_sum[0];
        mm[half0 ^ 1] = oth.x;
        ll[half0 ^ 1] = oth.y;
#else
        mm[0] = m_ref[0]; ll[0] = l_sum[0]; mm[1] = m_ref[1]; ll[1] = l_sum[1];
#endif
        const float mmax = fmaxf(mm[0], mm[1]);
        const float a0 = (mm[0] == -INFINITY) ? 0.f : fast_exp2(mm[0] - mmax);
        const float a1 = (mm[1] == -INFINITY) ? 0.f : fast_exp2(mm[1] - mmax);
        const float l_run = a0 * ll[0] + a1 * ll[1];
        if (n_act > 0) {
            mbar_wait(o_final, 0);
            tc_fence_after();
        }
        const float inv = (l_run > 1e-30f && n_act > 0) ? p.out_scale / l_run : 0.f;
        const float s0 = a0 * inv, s1 = a1 * inv;
        const bool row_ok = qi < p.lq;
        constexpr int FA_OCOLS = FA_D / FA_SPLIT;
        const int ocol0 = (FA_SPLIT == 2 ? half0 : 0) * FA_OCOLS;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)b * p.o_bstride + (size_t)qi * p.ldo + head * FA_D + ocol0;
#pragma unroll
        for (int c = 0; c < FA_OCOLS / 16; ++c) {
            uint32_t oa[16], ob[16];
            if (n_act > 0) {
                tmem_ld16(t_o0 + ocol0 + c * 16, oa);
                tmem_ld16(t_o0 + FA_D + ocol0 + c * 16, ob);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) oa[i] = ob[i] = 0u;
            }
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        f[e] = inv != 0.f ? __uint_as_float(oa[i + e]) * s0 + __uint_as_float(ob[i + e]) * s1 : 0.f;
                    uint4* dst = reinterpret_cast<uint4*>(orow + c * 16 + i);
                    if (p.accumulate) {
                        const uint4 prev = *dst;
                        const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[2 * e] += __low2float(ph[e]);
                            f[2 * e + 1] += __high2float(ph[e]);
                        }
                    }
                    *dst = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FA_TMEM_COLS);
    }
}

template <int LOGW, int D, int MODE>
static int launch_fa(const AttnKernelArgs& a, int q_tiles, int heads, int batch, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        C2V_CHECK_CUDA(cudaFuncSetAttribute(attn_fa_kernel<LOGW, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        attr_set = true;
    }
    C2V_CHECK_CUDA(launch(attn_fa_kernel<LOGW, D, MODE>, dim3(q_tiles, heads, batch), dim3(FA_THREADS), FA_SMEM, st, a));
    return OK;
}

int attn_fa_launch(const AttnKernelArgs& a, int q_tiles, int heads, int batch, cudaStream_t st) {
    if ((a.lk + FA_BN - 1) / FA_BN + (a.lk2 > 0 ? 1 : 0) > FA_MAX_TILES || a.lk2 > FA_BN) return ERR_UNSUPPORTED;
    if (!a.epi_F && !a.mask) return launch_fa<0, 1, 0>(a, q_tiles, heads, batch, st);                 // dense
    if (a.epi_F && a.bitmask && a.lk % FA_BN == 0) return launch_fa<0, 1, 0>(a, q_tiles, heads, batch, st);   // per-sample packed mask
    AttnKernelArgs g = a;
    g.bitmask = nullptr;
    if (a.epi_F && a.epi_H == a.epi_W && a.lk % FA_BN == 0) {
        const int w = a.epi_W, d = a.epi_d;
        if (w == 32 && d == 8) return launch_fa<5, 8, 2>(g, q_tiles, heads, batch, st);
        if (w == 16 && d == 16) return launch_fa<4, 16, 2>(g, q_tiles, heads, batch, st);
        if (w == 8 && d == 32) return launch_fa<3, 32, 2>(g, q_tiles, heads, batch, st);
        if (w == 16 && d == 8) return launch_fa<4, 8, 2>(g, q_tiles, heads, batch, st);
        if (w == 8 && d == 16) return launch_fa<3, 16, 2>(g, q_tiles, heads, batch, st);
    }
    g.tile_map = nullptr;                     // the tile map is defined for the power-of-two grids only
    return launch_fa<0, 1, 1>(g, q_tiles, heads, batch, st);
}

}  // namespace c2v
