// Internal launcher declarations shared between the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace c2v {

// norm.cu
int64_t groupnorm_ws_floats(int ns, int rows, int C);
int groupnorm_silu_launch(const float* x, const float* gamma, const float* beta, void* out, float* ws, int ns, int rows, int C, float eps,
                          int silu, cudaStream_t st);
int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, const float* add, void* out2, float* out_f32,
                     int rows, int C, float eps, int ld2, cudaStream_t st);

int softmax_rows_launch(const float* x, void* out, int rows, int n, float scale, cudaStream_t st);

// elementwise.cu
int to_channels_last_launch(const float* in, void* out, int B, int C, int S, int Cpad, int out_bf16, cudaStream_t st);
int from_channels_last_launch(const float* in, float* out, int B, int C, int S, cudaStream_t st);
int concat_channels_launch(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t rows, int Ca, int Cb, float scale16,
                           cudaStream_t st);
int cast_bf16_launch(const float* in, void* out, int64_t n, float scale, cudaStream_t st);
int upsample2x_launch(const float* in, void* out, int N, int H, int W, int C, cudaStream_t st);
int im2col_s2_launch(const float* in, void* out, int N, int H, int W, int C, int pad_lo, cudaStream_t st);
int copy_rows_launch(const void* src, void* dst, int rows, int C, int B, int64_t dst_bstride, int ldd, cudaStream_t st);
int skinny_linear_launch(const float* in, const void* w, const float* bias, float* out, int M, int N, int K, int silu_in, cudaStream_t st);
int timestep_embedding_launch(const int64_t* t, float* out, int n, int dim, cudaStream_t st);
int cfg_ddim_update_launch(const float* x, const float* ec, const float* eu, const float* enc, const float* noise, float* x_prev,
                           float* pred_x0, int B, int64_t n, float scale, float cam_w, float phi, float a_t, float a_prev, float sigma_t,
                           float sqrt_one_minus_at, cudaStream_t st);

// attn_t16.cu
int attention_temporal_launch(const void* qkv, void* out, int B, int T, int HW, int heads, int ldo, cudaStream_t st);

// pose.cu
int pixel_unshuffle_cl_launch(const float* in, void* out, int B, int C, int T, int H, int W, int r, cudaStream_t st);
int avgpool2_cl_launch(const float* in, float* out, void* out_b, int N, int H, int W, int C, cudaStream_t st);
int relu_launch(void* x, int64_t n, cudaStream_t st);
int attention_temporal_hd_launch(const void* qkv, void* out, int B, int T, int HW, int heads, int D, cudaStream_t st);

// epipolar.cu
int epipolar_mask_launch(const float* F, uint8_t* out, int B, int T1, int T2, int H, int W, int d, cudaStream_t st);
int plucker_launch(const float* K, const float* c2w, float* out, int B, int T, int H, int W, int plucker, cudaStream_t st);

}  // namespace c2v
