// extern "C" surface of libcamc2v_b200.so (declared in include/camc2v_b200.h).
// Argument validation, TMA descriptor construction and kernel dispatch; no allocation, no sync.
#include <math.h>
#include <stdlib.h>

#include "../../include/camc2v_b200.h"
#include "attn.h"
#include "common.cuh"
#include "gemm_tc.h"
#include "kernels.h"
#include "tma_host.h"

using namespace c2v;

extern "C" {

int c2v_abi_version(void) { return 1; }

#ifdef C2V_OPERAND_FP16
int c2v_operand_dtype(void) { return 1; }
#else
int c2v_operand_dtype(void) { return 0; }
#endif

const char* c2v_status_string(int s) {
    switch (s) {
        case OK: return "ok";
        case ERR_BAD_ARG: return "bad argument";
        case ERR_CUDA: return "CUDA error";
        case ERR_TMA_ENCODE: return "TMA descriptor encode failed";
        case ERR_UNSUPPORTED: return "unsupported shape";
        default: return "unknown status";
    }
}

int c2v_gemm_tile_n(int N, int epi) {
    if (epi == C2V_EPI_GEGLU) {
        // both halves (value | gate) of an output-column block live in one N tile
        // these short-K GEMMs are bound by L2 -> smem operand traffic: the widest tile has the highest FLOP per loaded byte
        // (the 32x32-level FF and init_attn; at the deeper levels M is small and the 2-stage ring of a 256-wide tile starves)
        if (N % 256 == 0 && N <= 4096) return 256;
        if (N % 160 == 0) return 160;
        if (N % 128 == 0) return 128;
        if (N % 64 == 0) return 64;
        return 0;
    }
    if (N % 160 == 0) return 160;
    if (N % 128 == 0) return 128;
    if (N <= 64) return 64;
    return 128;   // ragged last tile: TMA zero fill + masked stores
}

int c2v_gemm_splitk(int M, int N, int Cin, int taps, int epi) {
    // Fill the machine (148 SMs x 2 resident CTAs) when the output has few tiles but the K loop is deep: the 16x16, 8x8 and
    // 4x4 levels of the UNet at batch 1.  Each split keeps >= 15 K-iterations so the partial-tile traffic stays amortised.
    if (epi != C2V_EPI_LINEAR) return 1;
    const int bn = c2v_gemm_tile_n(N, epi);
    const int ctas = ((M + 127) / 128) * ((N + bn - 1) / bn);
    const int iters = taps * (Cin / 64);
    if (ctas >= 120 || iters < 30) return 1;   // >= 0.8 wave of tiles already: the reduce pass costs more than the idle SMs
    int s = (296 + ctas / 2) / ctas;           // round to the nearest multiple of one full wave
    if (s > 8) s = 8;
    if (s > iters / 15) s = iters / 15;
    return s < 2 ? 1 : s;
}

int c2v_gemm_persistent_plan(int M, int N, int Cin, int epi, int out_bf16, int has_residual, int* plan3) {
    if (!plan3 || M <= 0 || N <= 0 || Cin <= 0 || Cin % 64 != 0) return ERR_BAD_ARG;
    const int bn = c2v_gemm_tile_n(N, epi);
    if (bn == 0) return ERR_UNSUPPORTED;
    GemmKernelArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M;
    a.N = N;
    a.k_chunks = Cin / 64;
    a.taps = 1;
    a.a_mode = A_PLAIN;
    a.epi = epi;
    a.out_bf16 = out_bf16;
    a.residual = has_residual ? reinterpret_cast<const float*>(plan3) : nullptr;      // only tested for null-ness by the planner
    a.ldo = epi == C2V_EPI_GEGLU ? N / 2 : N;
    a.splits = c2v_gemm_splitk(M, N, Cin, 1, epi);
    a.tile_rows = 128;
    const PsPlan pl = gemm_ps_plan(a, bn);
    plan3[0] = pl.mode;
    plan3[1] = pl.mode ? pl.bn : bn;
    plan3[2] = pl.P;
    return OK;
}

int c2v_gemm(const c2v_gemm_desc* d, void* stream) {
    if (!d || !d->a || !d->w || !d->out) return ERR_BAD_ARG;
    if (d->M <= 0 || d->N <= 0 || d->Cin <= 0 || d->Cin % 64 != 0 || d->N % 4 != 0) return ERR_UNSUPPORTED;
    if (d->ldo % 4 != 0 || (d->residual && d->ldr % 4 != 0)) return ERR_UNSUPPORTED;
    if (d->taps != 1 && d->taps != 3 && d->taps != 9) return ERR_BAD_ARG;
    if (d->rowbias && d->rows_per_group <= 0) return ERR_BAD_ARG;
    int bn = c2v_gemm_tile_n(d->N, d->epi);
    if (bn == 0) return ERR_UNSUPPORTED;
    if (d->epi != C2V_EPI_GEGLU && d->splitk <= 1 && d->N % 64 == 0 && bn > 64) {
        // few output tiles and no split-K (shallow K): narrower N tiles put more CTAs (more TMA streams) on the machine
        const int mt = (d->M + 127) / 128;
        if (mt * ((d->N + bn - 1) / bn) < 120) bn = 64;
    }
    if (d->epi == C2V_EPI_GEGLU && (!d->out_bf16 || d->rowbias || d->residual)) return ERR_BAD_ARG;
    if (d->epi == C2V_EPI_GELU && d->residual) return ERR_BAD_ARG;          // activation of the GEMM result only
    if (d->epi < 0 || d->epi > C2V_EPI_GELU) return ERR_BAD_ARG;

    GemmKernelArgs a;
    memset(&a, 0, sizeof(a));
    a.M = d->M;
    a.N = d->N;
    a.k_chunks = d->Cin / 64;
    a.taps = d->taps;
    a.a_mode = d->a_mode;
    a.epi = d->epi;
    a.bias = d->bias;
    a.rowbias = d->rowbias;
    a.rows_per_group = d->rows_per_group > 0 ? d->rows_per_group : 1;
    a.residual = d->residual;
    a.ldr = d->ldr;
    a.out = d->out;
    a.ldo = d->ldo;
    a.out_bf16 = d->out_bf16;
    a.splits = d->splitk > 1 ? d->splitk : 1;
    if (a.splits > 1) {
        if (d->epi != C2V_EPI_LINEAR || a.splits > a.taps * a.k_chunks) return ERR_BAD_ARG;
        if (a.splits > 8) return ERR_UNSUPPORTED;
        if (!d->ws) return ERR_BAD_ARG;          // split-K needs the caller's fp32 workspace [splitk, M, N]
        a.out = d->ws;
    }

    if (d->a_mode == C2V_A_PLAIN) {
        if (d->taps != 1 || d->lda < d->Cin) return ERR_BAD_ARG;
        const uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)d->M};
        const uint64_t strides[1] = {(uint64_t)d->lda * 2};
        const uint32_t box[2] = {64, 128};
        if (!make_tmap_bf16(&a.tmA, d->a, 2, dims, strides, box)) return ERR_TMA_ENCODE;
        a.tile_rows = 128;
    } else if (d->a_mode == C2V_A_CONV2D) {
        const int W = d->d1, H = d->d2, NB = d->nb;
        if (W <= 0 || H <= 0 || NB <= 0 || d->M != NB * H * W) return ERR_BAD_ARG;
        if (W > 128 ? (W % 128 != 0) : (128 % W != 0)) return ERR_UNSUPPORTED;
        int bh, bnimg, bw = W;
        if (W > 128) {             // wide images (VAE decoder, 256 x 256): one tile = 128 consecutive pixels of one image row
            bw = 128;
            bh = 1;
            bnimg = 1;
        } else if (W * H >= 128) {
            if ((W * H) % 128 != 0) return ERR_UNSUPPORTED;
            bh = 128 / W;
            bnimg = 1;
        } else {
            bh = H;
            if (128 % (W * H) != 0) return ERR_UNSUPPORTED;
            bnimg = 128 / (W * H);
            if (bnimg > NB) bnimg = NB;
            if (NB % bnimg != 0) return ERR_UNSUPPORTED;
        }
        const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
        const uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)W * d->Cin * 2, (uint64_t)H * W * d->Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnimg};
        if (!make_tmap_bf16(&a.tmA, d->a, 4, dims, strides, box)) return ERR_TMA_ENCODE;
        a.dim1 = W;
        a.dim2 = H;
        a.tile_rows = bw * bh * bnimg;
    } else if (d->a_mode == C2V_A_CONVT) {
        const int HW = d->d1, T = d->d2, B = d->nb;
        if (HW <= 0 || T <= 0 || B <= 0 || d->M != B * T * HW) return ERR_BAD_ARG;
        int bhw, bt;
        if (HW >= 128) {
            if (HW % 128 != 0) return ERR_UNSUPPORTED;
            bhw = 128;
            bt = 1;
        } else {
            if (128 % HW != 0) return ERR_UNSUPPORTED;
            bhw = HW;
            bt = 128 / HW;
            if (bt > T) bt = T;
            if (T % bt != 0) return ERR_UNSUPPORTED;
        }
        const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)HW, (uint64_t)T, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)d->Cin * 2, (uint64_t)HW * d->Cin * 2, (uint64_t)T * HW * d->Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)bhw, (uint32_t)bt, 1};
        if (!make_tmap_bf16(&a.tmA, d->a, 4, dims, strides, box)) return ERR_TMA_ENCODE;
        a.dim1 = HW;
        a.dim2 = T;
        a.tile_rows = bhw * bt;
    } else {
        return ERR_BAD_ARG;
    }
    const PsPlan ps = gemm_ps_plan(a, bn);        // persistent form for the multi-wave 16-bit-output projections (gemm_ps.cu)
    if (ps.mode) bn = ps.bn;
    {
        const uint64_t ktot = (uint64_t)d->taps * d->Cin;
        const uint64_t dims[2] = {ktot, (uint64_t)d->N};
        const uint64_t strides[1] = {ktot * 2};
        const uint32_t box[2] = {64, (uint32_t)bn};
        if (!make_tmap_bf16(&a.tmB, d->w, 2, dims, strides, box)) return ERR_TMA_ENCODE;
    }
    if (ps.mode) return gemm_ps_launch(a, ps, (cudaStream_t)stream);
    if (d->epi != C2V_EPI_GEGLU) {
        // TMA epilogue descriptors (residual load, output / split-K partial store).  When the output rows are not 16-byte
        // aligned (e.g. a 4-column bf16 output) the kernel falls back to its direct-store epilogue.
        const bool part = a.splits > 1;
        const bool o_f32 = part || !d->out_bf16;
        const void* obase = part ? (const void*)d->ws : (const void*)d->out;
        const uint64_t ld = part ? (uint64_t)d->N : (uint64_t)d->ldo;
        const uint64_t esz = o_f32 ? 4 : 2;
        bool ok = (ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0;
        if (d->residual && !part) ok = ok && ((uint64_t)d->ldr * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0;
        if (ok) {
            const uint64_t odims[3] = {(uint64_t)d->N, (uint64_t)d->M, (uint64_t)a.splits};
            const uint64_t ostr[2] = {ld * esz, (uint64_t)d->M * ld * esz};
            const uint32_t obox[3] = {32, (uint32_t)a.tile_rows, 1};
            if (!make_tmap(&a.tmO, obase, 3, odims, ostr, obox, o_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (C2V_OPERAND_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16),
                           o_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
                return ERR_TMA_ENCODE;
            if (d->residual && !part) {
                const uint64_t rdims[2] = {(uint64_t)d->N, (uint64_t)d->M};
                const uint64_t rstr[1] = {(uint64_t)d->ldr * 4};
                const uint32_t rbox[2] = {32, (uint32_t)a.tile_rows};
                if (!make_tmap(&a.tmR, d->residual, 2, rdims, rstr, rbox, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B))
                    return ERR_TMA_ENCODE;
            }
            a.tma_epi = 1;
        }
    }
    const int m_tiles = (d->M + a.tile_rows - 1) / a.tile_rows;
    const int n_tiles = (d->N + bn - 1) / bn;
    if (n_tiles > 65535) return ERR_UNSUPPORTED;
    const int rc = gemm_tc_launch(a, bn, m_tiles, n_tiles, (cudaStream_t)stream);
    if (rc != OK || a.splits == 1) return rc;
    return splitk_reduce_launch(d->ws, a.splits, d->M, d->N, d->bias, d->rowbias, a.rows_per_group, d->residual, d->ldr, d->out, d->ldo,
                                d->out_bf16, (cudaStream_t)stream);
}

int c2v_skinny_linear(const float* in, const void* w, const float* bias, float* out, int M, int N, int K, int silu_in, void* stream) {
    if (!in || !w || !out || M <= 0 || N <= 0) return ERR_BAD_ARG;
    return skinny_linear_launch(in, w, bias, out, M, N, K, silu_in, (cudaStream_t)stream);
}

int c2v_timestep_embedding(const int64_t* t, float* out, int n, int dim, void* stream) {
    if (!t || !out || n <= 0) return ERR_BAD_ARG;
    return timestep_embedding_launch(t, out, n, dim, (cudaStream_t)stream);
}

int c2v_groupnorm_silu(const float* x, const float* gamma, const float* beta, void* out, float* ws, int ns, int rows, int C, float eps, int silu,
                       void* stream) {
    if (!x || !gamma || !beta || !out || !ws) return ERR_BAD_ARG;
    return groupnorm_silu_launch(x, gamma, beta, out, ws, ns, rows, C, eps, silu, (cudaStream_t)stream);
}

int64_t c2v_groupnorm_ws_floats(int ns, int rows, int C) { return groupnorm_ws_floats(ns, rows, C); }

int c2v_layernorm(const float* x, const float* gamma, const float* beta, void* out, const float* add, void* out2, float* out_f32, int rows,
                  int C, float eps, int ld_out2, void* stream) {
    if (!x || !gamma || !beta || !out || ld_out2 < 0) return ERR_BAD_ARG;
    return layernorm_launch(x, gamma, beta, out, add, out2, out_f32, rows, C, eps, ld_out2, (cudaStream_t)stream);
}

int c2v_epipolar_tile_map_words(int T, int H, int W);

int c2v_softmax_rows(const float* x, void* out, int rows, int n, float scale, void* stream) {
    if (!x || !out) return ERR_BAD_ARG;
    return softmax_rows_launch(x, out, rows, n, scale, (cudaStream_t)stream);
}

int c2v_attention(const c2v_attn_desc* d, void* stream) {
    if (!d || !d->q || !d->k || !d->v || !d->out) return ERR_BAD_ARG;
    if (d->bq <= 0 || d->lq <= 0 || d->lk <= 0 || d->heads <= 0 || d->kv_div <= 0) return ERR_BAD_ARG;
    if (d->heads > 65535 || d->bq > 65535) return ERR_UNSUPPORTED;
    if ((d->ldq | d->ldk | d->ldv | d->ldo) % 8 != 0) return ERR_UNSUPPORTED;
    if (!d->epi_F && !d->mask && !d->k2 && !d->v2 && d->lk2 <= 0 && d->lk <= 128) {
        // short key sequences (text / per-frame image cross-attention, self-attention of the 8x8 / 4x4 levels): all of K and V fit
        // in shared memory (measured: faster up to 128 keys, slower at 256 where every 128-query CTA would re-stage 74 KB), warp-level mma.sync kernel (attn_small.cu) instead of the 128 x 64 tcgen05 tile pipeline
        return attn_small_launch(d->q, d->k, d->v, d->out, d->bq, d->lq, d->lk, d->heads, d->kv_div, d->ldq, d->ldk, d->ldv, d->ldo,
                                 d->q_bstride, d->k_bstride, d->v_bstride, d->o_bstride, 0.125f * 1.4426950408889634f, d->out_scale,
                                 d->accumulate, (cudaStream_t)stream);
    }
    AttnKernelArgs a;
    memset(&a, 0, sizeof(a));
    const int hd = d->heads * 64;
    const int bk = (d->bq + d->kv_div - 1) / d->kv_div;
    const uint32_t qbox[3] = {64, 128, 1};    // 128 queries per CTA
    const uint32_t box[3] = {64, 64, 1};      // 64 keys per tile
    {
        const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)d->lq, (uint64_t)d->bq};
        const uint64_t strides[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->q_bstride * 2};
        if (!make_tmap_bf16(&a.tmQ, d->q, 3, dims, strides, qbox)) return ERR_TMA_ENCODE;
    }
    {
        const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)d->lk, (uint64_t)bk};
        const uint64_t sk[2] = {(uint64_t)d->ldk * 2, (uint64_t)d->k_bstride * 2};
        const uint64_t sv[2] = {(uint64_t)d->ldv * 2, (uint64_t)d->v_bstride * 2};
        if (!make_tmap_bf16(&a.tmK, d->k, 3, dims, sk, box)) return ERR_TMA_ENCODE;
        if (!make_tmap_bf16(&a.tmV, d->v, 3, dims, sv, box)) return ERR_TMA_ENCODE;
    }
    a.out = d->out;
    a.lq = d->lq;
    a.lk = d->lk;
    a.kv_div = d->kv_div;
    a.ldo = d->ldo;
    a.o_bstride = d->o_bstride;
    a.scale_log2 = 0.125f * 1.4426950408889634f;
    a.out_scale = d->out_scale;
    a.accumulate = d->accumulate;
    if (d->k2 || d->v2 || d->lk2 > 0) {
        if (!d->k2 || !d->v2 || d->lk2 <= 0 || d->lk2 > 64 || (d->ldk2 | d->ldv2) % 8 != 0) return ERR_BAD_ARG;
        const uint64_t dims[3] = {(uint64_t)hd, (uint64_t)d->lk2, 1};
        const uint64_t sk[2] = {(uint64_t)d->ldk2 * 2, (uint64_t)d->ldk2 * 2 * d->lk2};
        const uint64_t sv[2] = {(uint64_t)d->ldv2 * 2, (uint64_t)d->ldv2 * 2 * d->lk2};
        if (!make_tmap_bf16(&a.tmK2, d->k2, 3, dims, sk, box)) return ERR_TMA_ENCODE;
        if (!make_tmap_bf16(&a.tmV2, d->v2, 3, dims, sv, box)) return ERR_TMA_ENCODE;
        a.lk2 = d->lk2;
    }
    if (d->epi_F && d->mask) return ERR_BAD_ARG;
    a.epi_F = d->epi_F;
    if (d->epi_F) {
        if (d->epi_T <= 0 || d->epi_H <= 0 || d->epi_W <= 0 || d->epi_d <= 0) return ERR_BAD_ARG;
        if (d->lq != d->epi_T * d->epi_H * d->epi_W || d->lk != d->lq) return ERR_BAD_ARG;
        a.epi_T = d->epi_T;
        a.epi_H = d->epi_H;
        a.epi_W = d->epi_W;
        a.epi_d = d->epi_d;
        a.epi_thr = (float)((double)d->epi_d * sqrt(2.0) / 2.0);
        a.epi_off = (float)d->epi_d / 2.0f - 0.5f;
    }
    a.mask = d->mask;
    a.mask_bstride = d->mask_bstride;
    if (d->epi_tile_map && d->epi_F) {
        a.tile_map = d->epi_tile_map;
        a.tile_map_words = c2v_epipolar_tile_map_words(d->epi_T, d->epi_H, d->epi_W);
    }
    if (d->epi_bitmask && d->epi_F && d->lq % 128 == 0) a.bitmask = d->epi_bitmask;
    return attn_fa_launch(a, (d->lq + 127) / 128, d->heads, d->bq, (cudaStream_t)stream);
}

int c2v_attention_temporal(const void* qkv, void* out, int B, int T, int HW, int heads, int ldo, void* stream) {
    if (!qkv || !out || B <= 0 || HW <= 0 || heads <= 0 || ldo < 0) return ERR_BAD_ARG;
    return attention_temporal_launch(qkv, out, B, T, HW, heads, ldo, (cudaStream_t)stream);
}

int c2v_attention_temporal_hd(const void* qkv, void* out, int B, int T, int HW, int heads, int head_dim, void* stream) {
    if (!qkv || !out || B <= 0 || HW <= 0 || heads <= 0) return ERR_BAD_ARG;
    return attention_temporal_hd_launch(qkv, out, B, T, HW, heads, head_dim, (cudaStream_t)stream);
}

int c2v_pixel_unshuffle_cl(const float* in, void* out, int B, int C, int T, int H, int W, int r, void* stream) {
    if (!in || !out || B <= 0 || C <= 0 || T <= 0 || H <= 0 || W <= 0) return ERR_BAD_ARG;
    return pixel_unshuffle_cl_launch(in, out, B, C, T, H, W, r, (cudaStream_t)stream);
}

int c2v_avgpool2_cl(const float* in, float* out, void* out_16, int N, int H, int W, int C, void* stream) {
    if (!in || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0) return ERR_BAD_ARG;
    return avgpool2_cl_launch(in, out, out_16, N, H, W, C, (cudaStream_t)stream);
}

int c2v_relu(void* x, int64_t n, void* stream) {
    if (!x || n < 0) return ERR_BAD_ARG;
    if (n == 0) return OK;
    return relu_launch(x, n, (cudaStream_t)stream);
}

int c2v_epipolar_mask(const float* F, uint8_t* out, int B, int T, int H, int W, int d, void* stream) {
    if (!F || !out) return ERR_BAD_ARG;
    return epipolar_mask_launch(F, out, B, T, T, H, W, d, (cudaStream_t)stream);
}

int c2v_epipolar_mask_rect(const float* F, uint8_t* out, int B, int T1, int T2, int H, int W, int d, void* stream) {
    if (!F || !out) return ERR_BAD_ARG;
    return epipolar_mask_launch(F, out, B, T1, T2, H, W, d, (cudaStream_t)stream);
}

// bitmap words of one query-tile row (64-key tiles) + 1 word holding the longest-first issue order
int c2v_epipolar_tile_map_words(int T, int H, int W) { return ((T * H * W + 63) / 64 + 31) / 32 + 1; }

int64_t c2v_epipolar_bitmask_words(int T, int H, int W) { return (int64_t)T * H * W * (int64_t)(T * H * W / 32); }

int c2v_epipolar_bitmask(const float* F, uint32_t* out, int B, int T, int H, int W, int d, void* stream) {
    if (!F || !out || B <= 0) return ERR_BAD_ARG;
    return epi_bitmask_launch(F, out, B, T, H, W, d, (cudaStream_t)stream);
}

int c2v_epipolar_tile_map(const float* F, uint32_t* map, int B, int T, int H, int W, int d, void* stream) {
    if (!F || !map || B <= 0 || B > 65535) return ERR_BAD_ARG;
    return epi_tile_map_launch(F, map, B, T, H, W, d, (cudaStream_t)stream);
}

int c2v_plucker(const float* K, const float* c2w, float* out, int B, int T, int H, int W, int plucker, void* stream) {
    if (!K || !c2w || !out) return ERR_BAD_ARG;
    return plucker_launch(K, c2w, out, B, T, H, W, plucker, (cudaStream_t)stream);
}

int c2v_to_channels_last(const float* in, void* out, int B, int C, int S, int Cpad, int out_bf16, void* stream) {
    if (!in || !out) return ERR_BAD_ARG;
    return to_channels_last_launch(in, out, B, C, S, Cpad, out_bf16, (cudaStream_t)stream);
}

int c2v_from_channels_last(const float* in, float* out, int B, int C, int S, void* stream) {
    if (!in || !out) return ERR_BAD_ARG;
    return from_channels_last_launch(in, out, B, C, S, (cudaStream_t)stream);
}

int c2v_concat_channels(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, void* stream) {
    if (!a || !b || (!out_f32 && !out_bf16)) return ERR_BAD_ARG;
    return concat_channels_launch(a, b, out_f32, out_bf16, rows, Ca, Cb, 1.0f, (cudaStream_t)stream);
}

int c2v_concat_channels_scaled(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, float scale16,
                               void* stream) {
    if (!a || !b || (!out_f32 && !out_bf16) || !(scale16 > 0.0f)) return ERR_BAD_ARG;
    return concat_channels_launch(a, b, out_f32, out_bf16, rows, Ca, Cb, scale16, (cudaStream_t)stream);
}

int c2v_cast_bf16(const float* in, void* out, int64_t n, void* stream) {
    if (!in || !out || n <= 0) return ERR_BAD_ARG;
    return cast_bf16_launch(in, out, n, 1.0f, (cudaStream_t)stream);
}

int c2v_cast_bf16_scaled(const float* in, void* out, int64_t n, float scale, void* stream) {
    if (!in || !out || n <= 0 || !(scale > 0.0f)) return ERR_BAD_ARG;
    return cast_bf16_launch(in, out, n, scale, (cudaStream_t)stream);
}

int c2v_upsample2x(const float* in, void* out, int N, int H, int W, int C, void* stream) {
    if (!in || !out) return ERR_BAD_ARG;
    return upsample2x_launch(in, out, N, H, W, C, (cudaStream_t)stream);
}

int c2v_im2col_s2(const float* in, void* out, int N, int H, int W, int C, void* stream) {
    if (!in || !out) return ERR_BAD_ARG;
    return im2col_s2_launch(in, out, N, H, W, C, 1, (cudaStream_t)stream);
}

int c2v_im2col_s2_pad(const float* in, void* out, int N, int H, int W, int C, int pad_lo, void* stream) {
    if (!in || !out || (pad_lo != 0 && pad_lo != 1)) return ERR_BAD_ARG;
    return im2col_s2_launch(in, out, N, H, W, C, pad_lo, (cudaStream_t)stream);
}

int c2v_copy_rows(const void* src, void* dst, int rows, int C, int B, int64_t dst_bstride, int ldd, void* stream) {
    if (!src || !dst || rows <= 0 || B <= 0) return ERR_BAD_ARG;
    return copy_rows_launch(src, dst, rows, C, B, dst_bstride, ldd, (cudaStream_t)stream);
}

int c2v_cfg_ddim_update(const float* x, const float* e_cond, const float* e_uncond, const float* noise, float* x_prev, float* pred_x0, int B,
                        int64_t n, float scale, float guidance_rescale, float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at,
                        void* stream) {
    if (!x || !e_cond || !e_uncond || !noise || !x_prev || !pred_x0) return ERR_BAD_ARG;
    return cfg_ddim_update_launch(x, e_cond, e_uncond, nullptr, noise, x_prev, pred_x0, B, n, scale, 0.f, guidance_rescale, a_t, a_prev, sigma_t,
                                  sqrt_one_minus_at, (cudaStream_t)stream);
}

int c2v_cfg_ddim_update_cam(const float* x, const float* e_cond, const float* e_uncond, const float* e_cond_nocam, const float* noise,
                            float* x_prev, float* pred_x0, int B, int64_t n, float scale, float cam_weight, float guidance_rescale, float a_t,
                            float a_prev, float sigma_t, float sqrt_one_minus_at, void* stream) {
    if (!x || !e_cond || !e_uncond || !e_cond_nocam || !noise || !x_prev || !pred_x0) return ERR_BAD_ARG;
    return cfg_ddim_update_launch(x, e_cond, e_uncond, e_cond_nocam, noise, x_prev, pred_x0, B, n, scale, cam_weight, guidance_rescale, a_t, a_prev,
                                  sigma_t, sqrt_one_minus_at, (cudaStream_t)stream);
}

}  // extern "C"
