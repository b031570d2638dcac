// Kernel-side argument block of the tcgen05 flash attention (attn_fa.cu) and launchers of the attention family.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace c2v {

struct AttnKernelArgs {
    CUtensorMap tmQ, tmK, tmV;   // rank-3 bf16 maps (heads*64, rows, batch), box (64, 128, 1)
    CUtensorMap tmK2, tmV2;      // optional extra key segment shared by all batches (epipolar register tokens)
    int lk2;                     // rows of the extra segment (0 = none, <= 128)
    void* out;                   // bf16
    int lq, lk, kv_div;
    int ldo;
    long long o_bstride;
    float scale_log2;            // 64^-0.5 * log2(e)
    float out_scale;
    int accumulate;
    // epipolar mask over the main key segment: evaluated from F (null F and null mask => dense attention)
    const float* epi_F;          // [B, T, T, 3, 3]
    int epi_T, epi_H, epi_W, epi_d;
    // alternatively a materialised mask in the reference's format: uint8/bool [B, lq, lk]
    const unsigned char* mask;
    long long mask_bstride;
    // optional epipolar tile map: bit j of row (b, q_tile) = key tile j may contain an unmasked pair
    const unsigned int* tile_map;
    int tile_map_words;
    // optional packed mask [B][q_tile][k_chunk][128] (bit i of a word = key 32*k_chunk + i)
    const unsigned int* bitmask;
    float epi_thr, epi_off;      // float32(d*sqrt(2)/2), d/2 - 0.5
};

// attn_fa.cu: single-pass softmax in two key-half streams, lazy rescale, P kept in tensor memory, PV as a TMEM-operand MMA
int attn_fa_launch(const AttnKernelArgs& a, int q_tiles, int heads, int batch, cudaStream_t st);
// attn_small.cu: dense attention with lk <= 256 on warp-level tensor cores (all of K / V resident in shared memory)
int attn_small_launch(const void* q, const void* k, const void* v, void* out, int bq, int lq, int lk, int heads, int kv_div, int ldq, int ldk,
                      int ldv, int ldo, long long q_bstride, long long k_bstride, long long v_bstride, long long o_bstride, float scale_log2,
                      float out_scale, int accumulate, cudaStream_t st);
int epi_bitmask_launch(const float* F, unsigned int* out, int B, int T, int H, int W, int d, cudaStream_t st);
int epi_tile_map_launch(const float* F, unsigned int* map, int B, int T, int H, int W, int d, cudaStream_t st);

}  // namespace c2v
