// Kernel-side argument block of the tcgen05 GEMM / implicit-GEMM convolution (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace c2v {

enum AMode : int { A_PLAIN = 0, A_CONV2D = 1, A_CONVT = 2 };
enum EpiMode : int { EPI_LINEAR = 0, EPI_GEGLU = 1, EPI_GELU = 2 };   // GELU: linear epilogue + exact erf GELU (no residual)

struct GemmKernelArgs {
    CUtensorMap tmA;        // activations (bf16): rank 2 [K, M] or rank 4 (C, W, H, N) / (C, HW, T, B)
    CUtensorMap tmB;        // weights (bf16) [N, taps*Cin], K-major
    CUtensorMap tmR;        // residual (fp32) [M, ldr], box (32, tile_rows), 128B swizzle   (EPI_LINEAR with residual)
    CUtensorMap tmO;        // output [splits, M, ldo]: fp32 box (32, tile_rows, 1) SW128 | bf16 box (32, tile_rows, 1) SW64
    int tma_epi;            // 1: EPI_LINEAR epilogue goes through TMA (tmR / tmO valid)
    int M, N;               // output rows / columns
    int k_chunks;           // Cin / 64
    int taps;               // 1, 3 (temporal conv) or 9 (3x3 conv)
    int a_mode;             // AMode
    int dim1, dim2;         // A_CONV2D: W, H    A_CONVT: HW, T
    int tile_rows;          // valid rows per M tile (<= 128)
    int epi;                // EpiMode
    const float* bias;      // [N] or null
    const float* rowbias;   // [M / rows_per_group, N] or null (timestep-embedding add of ResBlock)
    int rows_per_group;
    const float* residual;  // fp32 [M, ldr] or null
    int ldr;
    void* out;              // fp32 or bf16 [M, ldo]
    int ldo;
    int out_bf16;
    int stages;             // smem ring depth (set by gemm_tc_launch)
    int splits;             // split-K factor (grid.z); > 1: raw fp32 partials go to `out` = ws[z][M][N], reduced by splitk_reduce_kernel
};

int gemm_tc_launch(const GemmKernelArgs& a, int bn, int m_tiles, int n_tiles, cudaStream_t st);
// gemm_ps.cu: persistent variant (one CTA per SM, double-buffered TMEM accumulator) for multi-wave 16-bit-output projections
struct PsPlan {
    int mode;   // 0: not taken; 1: streaming (A and B tiles through the ring); 2: B-resident (weight tile loaded once per CTA)
    int bn;     // N tile width the weight tensor map must be built with
    int P;      // B-resident: CTAs per N tile
};
PsPlan gemm_ps_plan(const GemmKernelArgs& a, int bn_default);      // needs M, N, k_chunks, taps, a_mode, epi, splits, out_bf16, residual, rowbias, tile_rows, ldo
int gemm_ps_launch(const GemmKernelArgs& a, const PsPlan& pl, cudaStream_t st);
int splitk_reduce_launch(const float* ws, int splits, int M, int N, const float* bias, const float* rowbias, int rows_per_group,
                         const float* residual, int ldr, void* out, int ldo, int out_bf16, cudaStream_t st);

}  // namespace c2v
