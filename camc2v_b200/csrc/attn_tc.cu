// Flash-style attention on tcgen05 tensor cores, head dim 64, bf16 operands, fp32 softmax.
//   replaces xformers.ops.memory_efficient_attention (R/lvdm/modules/attention.py:177,189), the einsum
//   softmax path (attention.py:105-129) and F.scaled_dot_product_attention with the boolean epipolar mask
//   (R/model/modules/epipolar.py:99).
//
// One CTA = 128 query rows of one (batch, head).  Per 64-key tile:
//     S = Q K^T          tcgen05.mma 128x64x16 (x4), both operands K-major in 128B-swizzled smem (TMA); S is double-buffered
//                        in TMEM so that S(j+1) is computed while the softmax warps are still reading S(j)
//     P = softmax tile   4 warps, one query row per thread, S read from TMEM in two passes (max + mask write-back, then exp)
//     O += P V           tcgen05.mma 128x64x16 (x4), P from smem (written swizzled by the softmax warps),
//                        V as MN-major operand straight from its natural [key, d] layout
// The running output O stays in TMEM; it is rescaled in place (tcgen05.ld/st) only when a row maximum grows.
// 64-key tiles (two image rows of a 32x32 latent frame) keep the epipolar tile map fine-grained: only tiles that can
// contain an attended pair are ever loaded, multiplied or soft-maxed.
//
// Epipolar mode: the reference materialises a bool mask [B, L, L] (268 MB per sample at 32x32x16) and
// reads it in every layer.  Here the mask is evaluated inside the softmax pass from the 3x3 fundamental
// matrices (9 floats per frame pair), with the reference's exact fp32 operation order
// (R/model/camcontexti2v.py:229-239; FMA-chain contraction, separately rounded norm, IEEE sqrt/div),
// so masked-out keys are bit-identical to the reference's mask and no mask ever touches HBM.
#include "attn_tc.h"
#include "common.cuh"

namespace c2v {

constexpr int AT_BM = 128;   // query rows per CTA
constexpr int AT_BN = 64;    // keys per tile
constexpr int AT_D = 64;
#ifndef C2V_AT_S_BUFS
#define C2V_AT_S_BUFS 2
#endif
#ifndef C2V_AT_KV_STAGES
#define C2V_AT_KV_STAGES 4
#endif
// 1: load the packed epipolar mask words of tile j+1 while tile j is processed.  The end-of-round capture
// (profiles/r01h_ncu_full.txt, SASS page) puts 10 % of the kernel's stall samples on the first use of the mask word: with S
// double-buffered the wait for S(j) is short, so a load issued right before it is not hidden.  Written at the end of round 1
// with no GPU time left to measure it: off until it has been A/B-ed (registers: see DESIGN.md section 6).
#ifndef C2V_AT_MASK_PREFETCH
#define C2V_AT_MASK_PREFETCH 0
#endif
#ifndef C2V_AT_MIN_CTAS
#define C2V_AT_MIN_CTAS 2
#endif
constexpr int AT_S_BUFS = C2V_AT_S_BUFS;        // S accumulators in TMEM
constexpr int AT_KV_STAGES = C2V_AT_KV_STAGES;
constexpr int AT_MIN_CTAS = C2V_AT_MIN_CTAS;    // resident CTAs per SM the register budget is sized for
constexpr int AT_THREADS = 192;

constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;        // 16 KB
constexpr int AT_K_BYTES = AT_BN * AT_D * 2;        // 8 KB
constexpr int AT_V_BYTES = AT_BN * AT_D * 2;        // 8 KB
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;       // 16 KB (128 rows x one 128-byte swizzle atom)
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_K = AT_OFF_Q + AT_Q_BYTES;
constexpr int AT_OFF_V = AT_OFF_K + AT_KV_STAGES * AT_K_BYTES;
constexpr int AT_OFF_P = AT_OFF_V + AT_KV_STAGES * AT_V_BYTES;
constexpr int AT_OFF_BAR = AT_OFF_P + AT_P_BYTES;
constexpr int AT_OFF_LIST = AT_OFF_BAR + 256;                   // active key-tile list (uint16), <= AT_MAX_TILES entries
constexpr int AT_MAX_TILES = 1024;                              // 65 472 keys + the register-token segment
constexpr int AT_SMEM = AT_OFF_LIST + AT_MAX_TILES * 2;

constexpr uint32_t AT_TM_S = 0;                       // S accumulator(s): AT_S_BUFS x 64 columns
constexpr uint32_t AT_TM_O = AT_S_BUFS * AT_BN;       // O accumulator: 64 columns
constexpr uint32_t AT_TMEM_COLS = AT_TM_O + AT_D <= 128 ? 128 : 256;

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

// LOGW >= 3 selects the fast epipolar predicate for square 2^LOGW x 2^LOGW key grids with pixel pitch D
// (LOGW = 0: dense attention, materialised masks and arbitrary grids).
constexpr uint32_t NEG_INF_BITS = 0xff800000u;

// max of 32 scores held as raw bits: four independent chains of 3-input maxima (FMNMX3)
__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
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

template <int LOGW, int D>
__global__ void __launch_bounds__(AT_THREADS, AT_MIN_CTAS) attn_tc_kernel(const __grid_constant__ AttnKernelArgs p) {
    constexpr bool FAST = LOGW >= 3;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;    // [AT_KV_STAGES]
    uint64_t* kv_empty = bars + 5;   // [AT_KV_STAGES]  (AT_KV_STAGES <= 4)
    uint64_t* s_full = bars + 9;     // [2]  S(j) is in TMEM buffer j & 1
    uint64_t* s_free = bars + 11;    // [2]
    uint64_t* p_full = bars + 13;
    uint64_t* pv_done = bars + 14;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5;
    // CTA -> (query tile, head).  With an epipolar tile map the CTAs are issued heaviest query tile first (longest-processing-
    // time order, all heads of a tile back to back): the per-tile work varies by 2-3x with the number of key tiles visited and
    // the grid is only ~2 waves deep, so the hardware's in-order CTA dispatch would otherwise leave a long tail.
    const int b = blockIdx.z;
    int q_tile = blockIdx.x, head = blockIdx.y;
    if (p.tile_map) {
        const int lin = blockIdx.x + gridDim.x * blockIdx.y;
        head = lin % gridDim.y;
        q_tile = (int)p.tile_map[((size_t)b * gridDim.x + lin / gridDim.y) * p.tile_map_words + (p.tile_map_words - 1)];
    }
    const int q0 = q_tile * AT_BM;
    const int bkv = b / p.kv_div;
    const int n_main = (p.lk + AT_BN - 1) / AT_BN;
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
        for (int s = 0; s < AT_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
        }
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, AT_TMEM_COLS);
        tmem_relinquish();
    }
    // Key tiles this query tile has to visit.  With an epipolar tile map (one bit per (query tile, key tile), built once per
    // sample by epi_tile_map_kernel with the same conservative test as the in-tile row skip) tiles that cannot contain an
    // unmasked pair are never loaded, multiplied or soft-maxed.
    uint16_t* tile_list = reinterpret_cast<uint16_t*>(smem + AT_OFF_LIST);
    int* n_act_s = reinterpret_cast<int*>(bars + 16);
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
        if (elect_one()) {
            mbar_expect_tx(q_full, AT_Q_BYTES);
            tma_load_3d(smem + AT_OFF_Q, &p.tmQ, q_full, head * AT_D, q0, b);
            for (int it = 0; it < n_act; ++it) {
                const int j = tile_list[it];
                const int s = it % AT_KV_STAGES;
                const uint32_t ph = (it / AT_KV_STAGES) & 1;
                mbar_wait<200>(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&kv_full[s], AT_K_BYTES + AT_V_BYTES);
                if (j < n_main) {
                    tma_load_3d(smem + AT_OFF_K + s * AT_K_BYTES, &p.tmK, &kv_full[s], head * AT_D, j * AT_BN, bkv);
                    tma_load_3d(smem + AT_OFF_V + s * AT_V_BYTES, &p.tmV, &kv_full[s], head * AT_D, j * AT_BN, bkv);
                } else {
                    tma_load_3d(smem + AT_OFF_K + s * AT_K_BYTES, &p.tmK2, &kv_full[s], head * AT_D, 0, 0);
                    tma_load_3d(smem + AT_OFF_V + s * AT_V_BYTES, &p.tmV2, &kv_full[s], head * AT_D, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);   // B = V is MN-major
        const uint32_t q_addr = smem_u32(smem + AT_OFF_Q);
        const uint32_t p_addr = smem_u32(smem + AT_OFF_P);
        auto issue_qk = [&](int j) {                  // j = iteration index over the active-tile list; S(j) -> TMEM buffer j & 1
            const int s = j % AT_KV_STAGES;
            mbar_wait<40>(&kv_full[s], (j / AT_KV_STAGES) & 1);
            if (j >= AT_S_BUFS) mbar_wait<40>(&s_free[j % AT_S_BUFS], ((j - AT_S_BUFS) / AT_S_BUFS) & 1);   // softmax(j - bufs) done
            tc_fence_after();
            if (elect_one()) {
                const uint64_t qd = umma_desc_sw128(q_addr);
                const uint64_t kd = umma_desc_sw128(smem_u32(smem + AT_OFF_K + s * AT_K_BYTES));
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k)
                    umma_bf16_ss(tmem_base + AT_TM_S + (uint32_t)(j % AT_S_BUFS) * AT_BN, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
                umma_commit(&s_full[j % AT_S_BUFS]);
            }
            __syncwarp();
        };
        mbar_wait<40>(q_full, 0);
        issue_qk(0);
        for (int j = 0; j < n_act; ++j) {
            const int s = j % AT_KV_STAGES;
            if (j + 1 < n_act) issue_qk(j + 1);       // S(j+1) is computed while the softmax warps work on S(j)
            mbar_wait<40>(p_full, j & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t v_addr = smem_u32(smem + AT_OFF_V + s * AT_V_BYTES);
#pragma unroll
                for (int ks = 0; ks < AT_BN / 16; ++ks) {
                    const uint64_t pd = umma_desc_sw128(p_addr) + 2 * ks;
                    const uint64_t vd = umma_desc_sw128(v_addr + ks * 16 * 128);
                    umma_bf16_ss(tmem_base + AT_TM_O, pd, vd, idesc_pv, (j | ks) != 0);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax / correction / epilogue (warps 2..5) =====================
        const int lg = warp & 3;
        const int r = lg * 32 + lane_id();
        const int qi = q0 + r;                                     // query index inside the batch
        const uint32_t t_s0 = tmem_base + AT_TM_S + ((uint32_t)(lg * 32) << 16);
        const uint32_t t_o = tmem_base + AT_TM_O + ((uint32_t)(lg * 32) << 16);
        const bool epi = p.epi_F != nullptr;
        const unsigned char* mrow = p.mask ? p.mask + (size_t)b * p.mask_bstride + (size_t)min(qi, p.lq - 1) * p.lk : nullptr;
        // epipolar query geometry
        const int HW = p.epi_H * p.epi_W;
        float xi = 0.f, yi = 0.f;
        const float* Frow = nullptr;
        if (epi) {
            const int qc = min(qi, p.lq - 1);
            const int t1 = qc / HW, pix = qc % HW;
            xi = __fadd_rn(__fmul_rn((float)(pix % p.epi_W), (float)p.epi_d), p.epi_off);
            yi = __fadd_rn(__fmul_rn((float)(pix / p.epi_W), (float)p.epi_d), p.epi_off);
            Frow = p.epi_F + ((size_t)b * p.epi_T + t1) * p.epi_T * 9;
        }
        int cur_t2 = -1;
        EpiLine line = {0.f, 0.f, 0.f};
        float thr_m = 0.f;         // threshold + rounding margin for the conservative row test (fast path)
        float l0x[FAST ? (1 << LOGW) : 1];   // line.l0 * x_j for the W pixel columns of a key row, refreshed once per key frame

        float m_run = -INFINITY;   // running max, already multiplied by scale*log2(e)
        float l_run = 0.f;
        uint8_t* p_row = smem + AT_OFF_P + (r >> 3) * 1024 + (r & 7) * 128;

#if C2V_AT_MASK_PREFETCH
        // Packed mask words one tile ahead (see the macro's comment): bw_n holds the words of tile j at the top of iteration j.
        uint32_t bw_n[AT_BN / 32] = {};
        auto load_mask_words = [&](int jt_, uint32_t (&w)[AT_BN / 32]) {
#pragma unroll
            for (int c = 0; c < AT_BN / 32; ++c) w[c] = 0u;
            if (FAST && p.bitmask && jt_ < n_main) {
                const uint32_t* mrow_w = p.bitmask + (((size_t)b * gridDim.x + q_tile) * (size_t)(p.lk >> 5) + (size_t)jt_ * (AT_BN / 32)) * AT_BM + r;
#pragma unroll
                for (int c = 0; c < AT_BN / 32; ++c) w[c] = __ldg(mrow_w + (size_t)c * AT_BM);
            }
        };
        if (n_act > 0) load_mask_words(tile_list[0], bw_n);
#endif

        for (int j = 0; j < n_act; ++j) {
            const int jt = tile_list[j];              // key tile index (j counts visited tiles: barrier parities)
            uint32_t bw[AT_BN / 32] = {};             // packed mask words of this row for the tile's 32-key chunks
#if C2V_AT_MASK_PREFETCH
#pragma unroll
            for (int c = 0; c < AT_BN / 32; ++c) bw[c] = bw_n[c];
            if (j + 1 < n_act) load_mask_words(tile_list[j + 1], bw_n);
#else
            if (FAST && p.bitmask && jt < n_main) {   // issued before the wait for S: the load latency hides behind QK^T
                const uint32_t* mrow_w = p.bitmask + (((size_t)b * gridDim.x + q_tile) * (size_t)(p.lk >> 5) + (size_t)jt * (AT_BN / 32)) * AT_BM + r;
#pragma unroll
                for (int c = 0; c < AT_BN / 32; ++c) bw[c] = __ldg(mrow_w + (size_t)c * AT_BM);
            }
#endif
            mbar_wait(&s_full[j % AT_S_BUFS], (j / AT_S_BUFS) & 1);
            tc_fence_after();
            const uint32_t t_s = t_s0 + (uint32_t)(j % AT_S_BUFS) * AT_BN;
            uint32_t anyc = 0;       // chunks in which at least one lane of this warp has a valid key (warp-uniform)
            bool wrote = false;      // some chunk of S was rewritten with masked scores (warp-uniform)
            float mx = -INFINITY;
            const bool main_seg = jt < n_main;
            const int klim = main_seg ? p.lk : p.lk2;
            const int tile_key0 = main_seg ? jt * AT_BN : 0;
            // ---- pass 1: row max; masked / out-of-range scores are replaced by -inf IN TMEM (tcgen05.st), so that pass 2 is
            //      the same predicate-free exp2 loop for every flavour of attention (exp2(-inf) = 0) ----
            if (FAST && epi && main_seg && p.bitmask) {
                // Packed mask (c2v_epipolar_bitmask, built once per sample with the same arithmetic): one bit test per element
#pragma unroll
                for (int c = 0; c < AT_BN / 32; ++c) {
                    if (__any_sync(0xffffffffu, bw[c] != 0u)) {
                        uint32_t v[32];
                        tmem_ld32(t_s + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (bw[c] >> i) & 1u ? v[i] : NEG_INF_BITS;
                        anyc |= 1u << c;
                        wrote = true;
                        tmem_st32(t_s + c * 32, v);
                        mx = fmaxf(mx, max32(v));
                    }
                }
            } else if (FAST && epi && main_seg) {
                // Square power-of-two key grid: a 32-key chunk is RPC whole image rows of one frame, pixel x of column i
                // is a compile-time constant and the reference's mask predicate costs FMUL+FFMA+FADD+FSETP per element.
                constexpr int W = 1 << (FAST ? LOGW : 5);
                constexpr int RPC = 32 / W;
                constexpr float DF = (float)D, OFFC = (float)D * 0.5f - 0.5f;
#pragma unroll
                for (int c = 0; c < AT_BN / 32; ++c) {
                    const int key0 = tile_key0 + c * 32;
                    const int t2 = key0 >> (2 * LOGW);
                    if (t2 != cur_t2) {                     // warp-uniform: once per key frame
                        cur_t2 = t2;
                        line = epi_line(Frow + t2 * 9, xi, yi);
                        const float cmax = (float)(W - 1) * DF + OFFC;
                        thr_m = p.epi_thr + 1e-6f + 4e-7f * (fabsf(line.l0) * cmax + fabsf(line.l1) * cmax + fabsf(line.l2));
#pragma unroll
                        for (int x = 0; x < W; ++x) l0x[x] = __fmul_rn(line.l0, (float)x * DF + OFFC);   // the FMUL of the predicate
                    }
                    const int py0 = (key0 & (W * W - 1)) >> LOGW;
                    float yr[RPC];
                    bool maybe = false;
#pragma unroll
                    for (int rr = 0; rr < RPC; ++rr) {
                        yr[rr] = (float)(py0 + rr) * DF + OFFC;          // exact: small integers
                        // the distance is linear in x: if both row ends are beyond threshold+margin on the same side, no key
                        // of this image row can satisfy the exact predicate
                        const float w0 = __fadd_rn(__fmaf_rn(line.l1, yr[rr], __fmul_rn(line.l0, OFFC)), line.l2);
                        const float w1 = __fadd_rn(__fmaf_rn(line.l1, yr[rr], __fmul_rn(line.l0, (float)(W - 1) * DF + OFFC)), line.l2);
                        maybe |= !((w0 > thr_m && w1 > thr_m) || (w0 < -thr_m && w1 < -thr_m));
                    }
                    if (__any_sync(0xffffffffu, maybe)) {
                        uint32_t v[32];
                        tmem_ld32(t_s + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float wv = __fadd_rn(__fmaf_rn(line.l1, yr[i >> LOGW], l0x[i & (W - 1)]), line.l2);
                            v[i] = fabsf(wv) < p.epi_thr ? v[i] : NEG_INF_BITS;
                        }
                        const float cm = max32(v);
                        if (__any_sync(0xffffffffu, cm > -INFINITY)) {
                            anyc |= 1u << c;
                            wrote = true;
                            tmem_st32(t_s + c * 32, v);
                            mx = fmaxf(mx, cm);
                        }
                    }
                }
            } else {
                const bool plain = !(epi && main_seg) && !(mrow && main_seg);
#pragma unroll
                for (int c = 0; c < AT_BN / 32; ++c) {
                    const int key0 = tile_key0 + c * 32;
                    if (plain && key0 >= klim) continue;          // chunk past the last key (ragged tail, register tokens)
                    anyc |= 1u << c;
                    uint32_t v[32];
                    tmem_ld32(t_s + c * 32, v);
                    tmem_ld_wait();
                    if (plain && key0 + 32 <= klim) {             // dense attention, full chunk: no predicate, S stays as it is
                        mx = fmaxf(mx, max32(v));
                        continue;
                    }
                    if (plain) {                                  // ragged last chunk: keys [0, nv) valid, everything in registers
                        const int nv = klim - key0;               // 1..31, warp-uniform
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = i < nv ? v[i] : NEG_INF_BITS;
                    } else {
                        uint32_t mw[8];
                        if (mrow && main_seg) {
                            if (key0 + 32 <= p.lk && (p.lk & 15) == 0) {
                                const uint4 m0 = *reinterpret_cast<const uint4*>(mrow + key0);
                                const uint4 m1 = *reinterpret_cast<const uint4*>(mrow + key0 + 16);
                                mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
                                mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
                            } else {
#pragma unroll
                                for (int w8 = 0; w8 < 8; ++w8) {
                                    uint32_t wv = 0;
                                    for (int e = 0; e < 4; ++e) {
                                        const int key = key0 + w8 * 4 + e;
                                        if (key < p.lk && mrow[key]) wv |= 1u << (8 * e);
                                    }
                                    mw[w8] = wv;
                                }
                            }
                        }
                        if (epi && main_seg) {                    // generic grid (non power-of-two / 4x4): slow but exact
                            uint32_t okbits = 0;                  // rolled predicate loop; scores stay in registers
#pragma unroll 1
                            for (int i = 0; i < 32; ++i) {
                                const int key = key0 + i;
                                bool ok = key < klim;
                                if (mrow) ok = ok && ((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0;
                                if (ok) {
                                    const int t2 = key / HW;
                                    if (t2 != cur_t2) {
                                        cur_t2 = t2;
                                        line = epi_line(Frow + t2 * 9, xi, yi);
                                    }
                                    const int pj = key - t2 * HW;
                                    const float xj = __fadd_rn(__fmul_rn((float)(pj % p.epi_W), (float)p.epi_d), p.epi_off);
                                    const float yj = __fadd_rn(__fmul_rn((float)(pj / p.epi_W), (float)p.epi_d), p.epi_off);
                                    const float dist = fabsf(__fadd_rn(__fmaf_rn(line.l1, yj, __fmul_rn(line.l0, xj)), line.l2));
                                    ok = dist < p.epi_thr;
                                }
                                okbits |= (ok ? 1u : 0u) << i;
                            }
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = (okbits >> i) & 1u ? v[i] : NEG_INF_BITS;
                        } else {                                  // materialised mask bytes (+ ragged tail)
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const bool ok = (key0 + i < klim) && ((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0;
                                v[i] = ok ? v[i] : NEG_INF_BITS;
                            }
                        }
                    }
                    wrote = true;
                    tmem_st32(t_s + c * 32, v);
                    mx = fmaxf(mx, max32(v));
                }
            }
            if (wrote) tmem_st_wait();
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_use);
            l_run *= alpha;
            // ---- PV(j-1) complete: the P buffer is free and O is stable; rescale O in place if some row maximum grew ----
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld32(t_o + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(t_o + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            // ---- pass 2: probabilities, written straight into the 128B-swizzled K-major P tile of the PV MMA ----
            float2 l2 = make_float2(0.f, 0.f);
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_use, -m_use);
#pragma unroll
            for (int c = 0; c < AT_BN / 32; ++c) {
                if ((anyc >> c) & 1u) {
                    uint32_t v[32];
                    tmem_ld32(t_s + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // pass 2 is issue-bound: the scale-and-shift and the row sums go two elements per instruction
                            const float2 x = ffma2(make_float2(__uint_as_float(v[g * 8 + 2 * i]), __uint_as_float(v[g * 8 + 2 * i + 1])), sc2, nm2);
                            const float2 e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                            l2 = fadd2(l2, e);
                            w[i] = pack_bf16(e.x, e.y);
                        }
                        *reinterpret_cast<uint4*>(p_row + (((c * 4 + g) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(p_row + (((c * 4 + g) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            l_run += l2.x + l2.y;
            m_run = m_new;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&s_free[j % AT_S_BUFS]);        // S(j) fully consumed: QK(j + bufs) may overwrite this buffer
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> global ----
        if (n_act > 0) mbar_wait(pv_done, (n_act - 1) & 1);
        tc_fence_after();
        const float inv = (l_run > 0.f && n_act > 0) ? p.out_scale / l_run : 0.f;
        const bool row_ok = qi < p.lq;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)b * p.o_bstride + (size_t)qi * p.ldo + head * AT_D;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + c * 32, o);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = inv != 0.f ? __uint_as_float(o[i + e]) * inv : 0.f;
                    uint4* dst = reinterpret_cast<uint4*>(orow + c * 32 + i);
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
        tmem_dealloc(tmem_base, AT_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// Packed epipolar mask: out[b][q_tile][k_chunk][r] bit i = mask[b][128 q_tile + r][32 k_chunk + i], evaluated with exactly the
// arithmetic of the in-kernel predicate above (and hence of the reference, camcontexti2v.py:229-239).  One CTA = 128 queries x
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
// Same conservative per-image-row interval test (and the same rounding margin) as the in-tile row skip of attn_tc_kernel,
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

template <int LOGW, int D>
static int launch_attn(const AttnKernelArgs& a, int q_tiles, int heads, int batch, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        C2V_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<LOGW, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_set = true;
    }
    attn_tc_kernel<LOGW, D><<<dim3(q_tiles, heads, batch), AT_THREADS, AT_SMEM, st>>>(a);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

int attn_tc_launch(const AttnKernelArgs& a, int q_tiles, int heads, int batch, cudaStream_t st) {
    if ((a.lk + AT_BN - 1) / AT_BN + (a.lk2 > 0 ? 1 : 0) > AT_MAX_TILES || a.lk2 > AT_BN) return ERR_UNSUPPORTED;
    if (a.epi_F && a.epi_H == a.epi_W && a.lk % AT_BN == 0) {
        const int w = a.epi_W, d = a.epi_d;
        if (w == 32 && d == 8) return launch_attn<5, 8>(a, q_tiles, heads, batch, st);
        if (w == 16 && d == 16) return launch_attn<4, 16>(a, q_tiles, heads, batch, st);
        if (w == 8 && d == 32) return launch_attn<3, 32>(a, q_tiles, heads, batch, st);
        if (w == 16 && d == 8) return launch_attn<4, 8>(a, q_tiles, heads, batch, st);
        if (w == 8 && d == 16) return launch_attn<3, 16>(a, q_tiles, heads, batch, st);
    }
    AttnKernelArgs g = a;
    g.tile_map = nullptr;                     // the tile map / packed mask are defined for the power-of-two grids only
    g.bitmask = nullptr;
    return launch_attn<0, 1>(g, q_tiles, heads, batch, st);
}

}  // namespace c2v
