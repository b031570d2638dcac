/*
 * camc2v_b200 — C ABI of the B200-native (sm_100a) denoising hot path of CamContextI2V.
 *
 * The reference (LDenninger/CamC2V) is pure Python and has no FFI: the seam it offers is the
 * `nn.Module.forward` contract of its UNet blocks.  The Python shims in
 * `camc2v_b200/` own those signatures and call the entry points below through ctypes with raw device
 * pointers; a maintainer of the reference would bind them the same way.
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - nothing allocates: outputs and workspaces are caller-owned;
 *   - return value: 0 = ok, 1 = bad argument, 2 = CUDA error, 3 = TMA descriptor error, 4 = unsupported shape;
 *   - activations are channels-last: a "token matrix" is [rows, C] with C contiguous; the canonical
 *     video layout is [B, T, H*W, C], so spatial (per frame), temporal (per pixel) and epipolar
 *     (all T*H*W tokens) attention are all views of the same buffer.
 *
 * Each declaration cites the reference code (R = CamContextI2V/) whose arithmetic it replaces.
 */
#ifndef CAMC2V_B200_H
#define CAMC2V_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library / ABI version; also proves the shared object is the CUDA build (smoke tests check it). */
int c2v_abi_version(void);
/* 16-bit operand type this build of the library uses for every `bf16` buffer of this header: 0 = bfloat16 (default build),
 * 1 = IEEE half (the -DC2V_OPERAND_FP16 build, libcamc2v_b200_fp16.so): same kernels, same speed, ~8x smaller operand
 * rounding error; accumulation / statistics / softmax / residual stream are fp32 in both. */
int c2v_operand_dtype(void);
/* Human-readable string for a status code returned by any entry point (host pointer, static storage). */
const char* c2v_status_string(int status);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 tensor cores: nn.Linear, Conv2d 3x3/1x1, Conv3d (3,1,1).
 *   replaces: every nn.Linear of R/lvdm/modules/attention.py:58-62,277,301,348,378,434,451-455 and
 *   R/model/modules/epipolar.py:55-63; nn.Conv2d of R/lvdm/modules/networks/openaimodel3d.py:151-155,
 *   175-187, 68-70, 96, 386, 561-565; nn.Conv3d (3,1,1) of openaimodel3d.py:255-266.
 * ---------------------------------------------------------------------------------------------- */
enum { C2V_A_PLAIN = 0, C2V_A_CONV2D = 1, C2V_A_CONVT = 2 };
enum { C2V_EPI_LINEAR = 0, C2V_EPI_GEGLU = 1, C2V_EPI_GELU = 2 };   /* GELU: out = gelu_erf(A W^T + bias) (resampler.py:31-38) */

typedef struct c2v_gemm_desc {
    const void* a;         /* bf16 activations.  PLAIN: [M, lda];  CONV2D: [nb, d2(H), d1(W), Cin];  CONVT: [nb(B), d2(T), d1(HW), Cin] */
    const void* w;         /* bf16 weights [N, taps*Cin], K contiguous, tap-major (tap = ky*3+kx or kt) */
    const float* bias;     /* fp32 [N] or NULL */
    const float* rowbias;  /* fp32 [M / rows_per_group, N] or NULL (ResBlock timestep-embedding add, openaimodel3d.py:219-228) */
    const float* residual; /* fp32 [M, ldr] or NULL, added in the epilogue */
    void* out;             /* fp32 or bf16 [M, ldo] */
    int M, N, Cin, taps;   /* taps: 1 (linear / 1x1), 3 (temporal conv), 9 (3x3 conv, pad 1, stride 1) */
    int a_mode;            /* C2V_A_* */
    int nb, d1, d2;        /* see `a` */
    int lda;               /* PLAIN: row stride of A in elements (>= Cin) */
    int rows_per_group;    /* rows sharing one rowbias row */
    int ldr, ldo;
    int out_bf16;          /* 0: fp32 out, 1: bf16 out */
    int epi;               /* C2V_EPI_*.  GEGLU (attention.py:431-438): w/bias rows are pre-interleaved per N tile
                              (see camc2v_b200.ops.geglu_interleave); out is bf16 [M, N/2] */
    int splitk;            /* 0/1: none.  > 1: split the K loop over `splitk` CTAs per tile (deep-K, few-tile GEMMs at the 16x16 /
                              8x8 / 4x4 levels); the partials are reduced deterministically (fixed order) with the fused epilogue */
    float* ws;             /* required when splitk > 1: fp32 scratch of splitk*M*N floats; partial tiles go through it (L2-resident
                              at these sizes) and a second kernel reduces them in fixed order. */
} c2v_gemm_desc;

int c2v_gemm(const c2v_gemm_desc* d, void* stream);
/* N-tile width the GEMM will use for a given N (needed to interleave GEGLU weights). */
int c2v_gemm_tile_n(int N, int epi);
/* Split-K factor this library recommends for a GEMM (1 = none). */
int c2v_gemm_splitk(int M, int N, int Cin, int taps, int epi);
/* Which kernel form a plain linear (one tap, contiguous output) takes, for inspection / tests:
 * plan3 = {mode, N tile, CTAs per N tile}; mode 0 = one output tile per CTA (gemm_tc.cu), 1 = persistent (one CTA per SM walks the
 * tile list, accumulator double-buffered in tensor memory), 2 = persistent with the [N tile, K] weight tile resident in shared
 * memory (gemm_ps.cu).  The persistent forms are taken by 16-bit-output products without residual and with >= 300 output tiles
 * (the GEGLU and q|k|v projections of the 32x32 / 16x16 levels). */
int c2v_gemm_persistent_plan(int M, int N, int Cin, int epi, int out_bf16, int has_residual, int* plan3);

/* Small-M linear (time/fps embedding MLPs, ResBlock emb_layers; openaimodel3d.py:168-174, 370-380):
 * out[m,n] = sum_k act(in[m,k]) * w[n,k] + bias[n];  act = SiLU if silu_in.  in/out fp32, w bf16. */
int c2v_skinny_linear(const float* in, const void* w_bf16, const float* bias, float* out, int M, int N, int K, int silu_in, void* stream);

/* Sinusoidal timestep embedding (R/lvdm/models/utils_diffusion.py:8-28): t int64 [n] -> fp32 [n, dim] = [cos | sin]. */
int c2v_timestep_embedding(const int64_t* t, float* out, int n, int dim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation (HBM-bound, fused with the activation and the bf16 cast that feeds the next GEMM).
 * ---------------------------------------------------------------------------------------------- */
/* GroupNorm(32 groups) [+ SiLU] over channels-last fp32 x [ns, rows, C]; statistics per (sample, group)
 * over rows x C/32 in fp32 (GroupNormSpecific, R/lvdm/basics.py:78-80; nn.GroupNorm eps 1e-6 of
 * attention.py:273,343; eps 1e-5 + SiLU of openaimodel3d.py:151-153,175-177,255-265).  out: bf16 [ns*rows, C].
 * `ws` is fp32 scratch of at least c2v_groupnorm_ws_floats(ns, rows, C) floats. */
int c2v_groupnorm_silu(const float* x, const float* gamma, const float* beta, void* out_bf16, float* ws,
                       int ns, int rows, int C, float eps, int silu, void* stream);
int64_t c2v_groupnorm_ws_floats(int ns, int rows, int C);

/* Row softmax: out[r, :] = softmax(scale * x[r, :]), fp32 [rows, n] -> 16-bit operands [rows, n] (n % 4 == 0).  The softmax of the
 * single-head 512-wide attention block of the VAE decoder (ae_modules.py:53-80), whose scores / P.V products go through c2v_gemm. */
int c2v_softmax_rows(const float* x, void* out_bf16, int rows, int n, float scale, void* stream);

/* LayerNorm over the last dim (nn.LayerNorm, attention.py:232-234), fp32 in -> bf16 out.  If `add` is not
 * NULL a second output out2 = LN(x) + add is produced (normed_x + pluker features,
 * R/model/modules/modified_forwards.py:508-520); add is fp32 [rows, C].  out_f32 (optional) receives the
 * un-rounded normalised rows (CameraCtrl adds cc_projection(...) to them, cameractrl_modified_modules.py:235-239).
 * ld_out2: row stride of out2 in elements (0 = C; a multiple of 4) — the temporal block writes LN(x) + pluker straight into its
 * column block of the K-concatenated operand of the fused output projection (see c2v_attention_temporal). */
int c2v_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, const float* add, void* out2_bf16,
                  float* out_f32, int rows, int C, float eps, int ld_out2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention.
 * ---------------------------------------------------------------------------------------------- */
/* Flash-style softmax(Q K^T / sqrt(64)) V on tcgen05, head dim 64, bf16 in/out, fp32 softmax.
 *   replaces xformers.ops.memory_efficient_attention (attention.py:177,189), the einsum path
 *   (attention.py:105-129) and F.scaled_dot_product_attention with the bool epipolar mask (epipolar.py:99).
 * q: [bq, lq, heads*64] (row stride ldq elements), k/v: [bk, lk, heads*64] (strides ldk/ldv); query batch b uses
 * kv batch b / kv_div (context broadcast over frames).  out: [bq, lq, heads*64] (stride ldo).
 * accumulate != 0: out = out + out_scale * result (image cross-attention sum, attention.py:140-144).
 * k2/v2 (optional, lk2 <= 64 rows, [lk2, heads*64], shared by every batch): an extra, never-masked key
 *   segment — the epipolar register tokens (epipolar.py:86-96; softmax is invariant to key order, so they
 *   are appended instead of prepended).
 * Masking of the main segment, at most one of:
 *   epi_F != NULL : query i = (t1, pixel) and key j = (t2, pixel) of a [T, H, W] grid (lq == lk == T*H*W) attend
 *                   iff the reference's epipolar mask (camcontexti2v.py:202-271) is true; evaluated in-kernel,
 *                   bit-exact, from epi_F fp32 [bq, T, T, 3, 3] — no mask is ever read from HBM;
 *   mask  != NULL : a materialised mask in the reference's own format, uint8/bool [bq, lq, lk] (drop-in for
 *                   an unmodified `sample_locs_dict`). */
typedef struct c2v_attn_desc {
    const void* q; const void* k; const void* v; void* out;
    int bq, lq, lk, heads;
    int ldq, ldk, ldv, ldo;
    int64_t q_bstride, k_bstride, v_bstride, o_bstride; /* batch strides in elements */
    int kv_div;
    float out_scale; int accumulate;
    const void* k2; const void* v2; int lk2, ldk2, ldv2;
    const float* epi_F; int epi_T, epi_H, epi_W, epi_d;
    const uint8_t* mask; int64_t mask_bstride;
    const uint32_t* epi_tile_map;  /* optional (with epi_F): output of c2v_epipolar_tile_map for the same F and grid; key tiles that
                                      cannot hold an unmasked pair are skipped entirely (results are bit-identical without it) */
    const uint32_t* epi_bitmask;   /* optional (with epi_F): output of c2v_epipolar_bitmask for the same F and grid: the mask packed
                                      to 1 bit per (query, key) pair, so the kernel tests a bit instead of re-evaluating the
                                      predicate in every layer / pass / step (results are bit-identical without it) */
} c2v_attn_desc;
int c2v_attention(const c2v_attn_desc* d, void* stream);

/* Temporal self-attention over T <= 32 frames per pixel (attention.py:105-129 with q=k=v of length T):
 * qkv bf16 [B, T, HW, 3*heads*64] packed (q | k | v), out bf16 [B, T, HW, heads*64] with row stride ldo elements (0 = heads*64;
 * a multiple of 8).  A stride wider than the row lets the caller collect several attention outputs side by side: the
 * camera-conditioned temporal block (modified_forwards.py:505-536) sums pluker_projection(n + p), Epipolar(n + p).to_out and
 * attn1(n).to_out into the stream, which is ONE c2v_gemm over the K-concatenated operand [n + p | attn1 | epipolar]. */
int c2v_attention_temporal(const void* qkv, void* out, int B, int T, int HW, int heads, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera pose encoder (SURVEY f-2; CamContextI2V/model/modules/camera_pose_encoder.py:295-376): the ops the UNet path does
 * not already provide.  16-bit tensors are in the library's operand dtype (c2v_operand_dtype).
 * -----------------------------------------------------------------------------------------------*/
/* Temporal self-attention with head dim `head_dim` (multiple of 8, <= 160; 320/8, 640/8, 1280/8 in the shipped config) over
 * T <= 16 frames: diffusers Attention / AttnProcessor2_0 as used by TemporalSelfAttention (camera_pose_encoder.py:101-160).
 * qkv [B, T, HW, 3*heads*head_dim] packed (q | k | v), out [B, T, HW, heads*head_dim]. */
int c2v_attention_temporal_hd(const void* qkv, void* out, int B, int T, int HW, int heads, int head_dim, void* stream);
/* rearrange 'b c f h w -> (b f) c h w' + nn.PixelUnshuffle(r) (camera_pose_encoder.py:359-361): fp32 [B, C, T, H, W] ->
 * 16-bit channel-last rows [(b, f, y, x), C*r*r], channel = c*r*r + dy*r + dx. */
int c2v_pixel_unshuffle_cl(const float* in, void* out, int B, int C, int T, int H, int W, int r, void* stream);
/* nn.AvgPool2d(2, 2) (Downsample with use_conv=False, camera_pose_encoder.py:212-231) on channel-last fp32 rows [N, H, W, C];
 * out fp32 [N, H/2, W/2, C], out_16 (optional) the same in the operand dtype. */
int c2v_avgpool2_cl(const float* in, float* out, void* out_16, int N, int H, int W, int C, void* stream);
/* in-place ReLU (ResnetBlock.act, camera_pose_encoder.py:262) on n 16-bit values, n % 8 == 0. */
int c2v_relu(void* x, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera geometry.
 * ---------------------------------------------------------------------------------------------- */
/* Materialise the epipolar mask (camcontexti2v.py:202-271): F fp32 [B,T,T,3,3] -> uint8 [B, T*H*W, T*H*W]. Bit-exact. */
int c2v_epipolar_mask(const float* F, uint8_t* out, int B, int T, int H, int W, int d, void* stream);
/* Rectangular form: F fp32 [B,T1,T2,3,3] (T1 query frames, T2 key frames) -> uint8 [B, T1*H*W, T2*H*W]: the conditional mask
 * between the 16 target frames and the 1 + n context frames that MultiLatentEpipolarAdaptor consumes
 * (compute_conditional_epipolar_mask, camcontexti2v.py:493-521).  Bit-exact. */
int c2v_epipolar_mask_rect(const float* F, uint8_t* out, int B, int T1, int T2, int H, int W, int d, void* stream);
/* Conservative tile-occupancy bitmap of the epipolar mask for 128x64 (query, key) tiles of a square power-of-two grid
 * (W in {8,16,32}): map[b][q_tile][word] bit j = key tile j may contain an attended pair.  Words per row =
 * c2v_epipolar_tile_map_words(T,H,W); the LAST word of row r is not part of the bitmap: it holds the index of the query
 * tile with the r-th largest number of visited key tiles, and c2v_attention issues its CTAs in that (longest first) order.
 * Returns 4 (unsupported) for other grids: callers then simply pass no map. */
int c2v_epipolar_tile_map(const float* F, uint32_t* map, int B, int T, int H, int W, int d, void* stream);
int c2v_epipolar_tile_map_words(int T, int H, int W);
/* The mask of camcontexti2v.py:202-271 packed to bits in the layout c2v_attention reads with coalesced 128-byte loads:
 * out[b][q_tile][k_chunk][r] (uint32), bit i = mask[b][128*q_tile + r][32*k_chunk + i]; T*H*W * T*H*W / 8 bytes per batch
 * element (32 MB at 32x32x16, against 268 MB for the reference's bool mask).  Same supported grids as the tile map. */
int c2v_epipolar_bitmask(const float* F, uint32_t* out, int B, int T, int H, int W, int d, void* stream);
int64_t c2v_epipolar_bitmask_words(int T, int H, int W);      /* uint32 words per batch element */
/* Pluecker / ray embedding (R/model/base.py:112-174): K fp32 [B,T,3,3], c2w fp32 [B,T,4,4] -> fp32 [B,6,T,H,W]. */
int c2v_plucker(const float* K, const float* c2w, float* out, int B, int T, int H, int W, int plucker, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / elementwise glue on the residual stream.
 * ---------------------------------------------------------------------------------------------- */
/* [B, C, S] fp32 -> channels-last [B, S, Cpad] (channels >= C zero-filled); out fp32 or bf16. */
int c2v_to_channels_last(const float* in, void* out, int B, int C, int S, int Cpad, int out_bf16, void* stream);
/* channels-last fp32 [B, S, C] -> [B, C, S] fp32. */
int c2v_from_channels_last(const float* in, float* out, int B, int C, int S, void* stream);
/* out = cat([a, b], channel) for channels-last fp32 [rows, Ca] / [rows, Cb]; writes fp32 (out_f32) and/or bf16 (out_bf16). */
int c2v_concat_channels(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, void* stream);
/* fp32 -> bf16 cast of n elements. */
int c2v_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* The same two with the 16-bit copy multiplied by `scale` (a power of two) first and, in the IEEE-half build, saturated at +-65504
 * instead of overflowing to inf.  For 16-bit copies of UNNORMALISED data: the fp32 residual stream / skip concat that feeds the
 * ResBlock's 1x1 skip convolution (R/lvdm/modules/networks/openaimodel3d.py:197-236 computes it in fp32 / autocast): the host stores
 * the copy at 2^-8 and packs that convolution's weights at 2^8, so activations up to 1.6e7 stay finite and the product is unchanged. */
int c2v_concat_channels_scaled(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, float scale16,
                               void* stream);
int c2v_cast_bf16_scaled(const float* in, void* out_bf16, int64_t n, float scale, void* stream);
/* nearest 2x upsample (openaimodel3d.py:101-103) of channels-last fp32 [N,H,W,C] -> bf16 [N,2H,2W,C]. */
int c2v_upsample2x(const float* in, void* out_bf16, int N, int H, int W, int C, void* stream);
/* im2col for the stride-2 3x3 Downsample conv (openaimodel3d.py:68-70): fp32 [N,H,W,C] -> bf16 [N*(H/2)*(W/2), 9*C]. */
int c2v_im2col_s2(const float* in, void* out_bf16, int N, int H, int W, int C, void* stream);
/* Same with the padding made explicit: pad_lo = 1 is c2v_im2col_s2; pad_lo = 0 pads on the right / bottom only, which is the
 * `F.pad(x, (0,1,0,1))` + stride-2 conv of the VAE encoder's Downsample (ae_modules.py:102-106). */
int c2v_im2col_s2_pad(const float* in, void* out_bf16, int N, int H, int W, int C, int pad_lo, void* stream);
/* out[b, r + row_off, :] = src[r, :] for r < rows (bf16): writes the pre-projected register tokens in front of K / V. */
int c2v_copy_rows(const void* src_bf16, void* dst_bf16, int rows, int C, int B, int64_t dst_bstride, int ldd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampler: fused classifier-free-guidance combine + guidance rescale + DDIM update
 *   (DDIMSampler.p_sample_ddim, R/lvdm/models/samplers/ddim.py:262-346; rescale_noise_cfg,
 *    R/lvdm/models/utils_diffusion.py:147-158).  All tensors fp32 [B, n]; one CTA per sample.
 *   e = e_u + scale (e_c - e_u);  e = phi e std(e_c)/std(e) + (1-phi) e;  pred_x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);
 *   x_prev = sqrt(a_prev) pred_x0 + sqrt(max(1-a_prev-sigma^2,0)) e + sigma noise.
 * ---------------------------------------------------------------------------------------------- */
int c2v_cfg_ddim_update(const float* x, const float* e_cond, const float* e_uncond, const float* noise, float* x_prev, float* pred_x0,
                        int B, int64_t n, float scale, float guidance_rescale, float a_t, float a_prev, float sigma_t,
                        float sqrt_one_minus_at, void* stream);
/* The same with the reference's camera guidance (ddim.py:268-280, `camera_cfg != 1`): a third UNet pass e_cond_nocam (the
 * conditional pass without the camera condition) enters as  model_output += cam_weight * (e_cond - e_cond_nocam)  before the
 * guidance rescale; cam_weight = (camera_cfg - 1) * scheduler weight (1 for "constant", cos((1 - t/999) pi/2) for "cosine"). */
int c2v_cfg_ddim_update_cam(const float* x, const float* e_cond, const float* e_uncond, const float* e_cond_nocam, const float* noise,
                            float* x_prev, float* pred_x0, int B, int64_t n, float scale, float cam_weight, float guidance_rescale, float a_t,
                            float a_prev, float sigma_t, float sqrt_one_minus_at, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAMC2V_B200_H */
