// HBM-bound normalisation kernels on the channels-last fp32 residual stream.  Each writes the bf16
// operand of the GEMM that follows, so norm + activation + cast is one read and one (half-size) write.
//   GroupNorm(32) [+SiLU]: R/lvdm/basics.py:78-89 (GroupNormSpecific, fp32 statistics),
//                          R/lvdm/modules/attention.py:273,343 (eps 1e-6),
//                          R/lvdm/modules/networks/openaimodel3d.py:151-153,175-177,255-265 (eps 1e-5 + SiLU)
//   LayerNorm:             R/lvdm/modules/attention.py:232-234
#include "common.cuh"
#include "kernels.h"

namespace c2v {

constexpr int GN_THREADS = 512;
constexpr int GN_MAX_SLOTS = 2;   // C/4 <= GN_THREADS * GN_MAX_SLOTS  ->  C <= 4096

// Thread mapping shared by both GroupNorm kernels: the C/4 float4 columns of a row are spread over the
// threads (slot s handles column-vector tid%nvec + s*GN_THREADS when nvec > GN_THREADS); when a row is
// narrower than the block several rows are processed side by side.  Every thread therefore owns a fixed
// set of channels and every global access is a fully coalesced float4.
struct GnMap {
    int nvec, slots, rows_par, rsub, cv0;
    bool active;
};
__device__ __forceinline__ GnMap gn_map(int C) {
    GnMap m;
    m.nvec = C >> 2;
    if (m.nvec >= GN_THREADS) {
        m.slots = (m.nvec + GN_THREADS - 1) / GN_THREADS;
        m.rows_par = 1;
        m.rsub = 0;
        m.cv0 = threadIdx.x;
        m.active = true;
    } else {
        m.slots = 1;
        m.rows_par = GN_THREADS / m.nvec;
        m.rsub = threadIdx.x / m.nvec;
        m.cv0 = threadIdx.x % m.nvec;
        m.active = m.rsub < m.rows_par;
    }
    return m;
}

// partial (sum, sumsq) per (sample, slab, group)
__global__ void __launch_bounds__(GN_THREADS, 2) gn_stats_kernel(const float* __restrict__ x, float* __restrict__ ws, int rows, int C,
                                                              int rows_per_slab) {
    pdl_entry();
    // Per-thread fp32 partials are combined across the CTA in a FIXED order (bit-reproducible run to run) without atomics: every
    // thread stages (sum, sumsq) of the <= 2 groups its float4 column touches in shared memory, then one warp per group adds the
    // staged values of that group's columns (lane-strided, doubles) and finishes with a shuffle tree.  (Round 1 used 64-bit
    // fixed-point shared-memory atomics here: ATOMS.CAST.SPIN loops with ~16 threads contending per address.)
    __shared__ float s_part[GN_MAX_SLOTS][4][GN_THREADS];
    const int n = blockIdx.y, slab = blockIdx.x, nslab = gridDim.x;
    const GnMap m = gn_map(C);
    const int cg = C / 32;
    const int r0 = slab * rows_per_slab;
    const int r1 = min(rows, r0 + rows_per_slab);
    const float* xs = x + (size_t)n * rows * C;
    for (int s = 0; s < GN_MAX_SLOTS; ++s) {
        float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
        const int cv = m.cv0 + s * GN_THREADS;
        if (m.active && s < m.slots && cv < m.nvec) {
            float a[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
            const float* col = xs + cv * 4;
            const size_t step = (size_t)m.rows_par * C;
            int r = r0 + m.rsub;
            // 4 independent 16-byte loads in flight per thread
            // the last batch is predicated (zero rows add nothing) rather than walked row by row: one round trip, not up to three
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (; r < r1; r += 4 * m.rows_par) {
                const float* p0 = col + (size_t)r * C;
                const float4 v0 = *reinterpret_cast<const float4*>(p0);
                const float4 v1 = r + m.rows_par < r1 ? *reinterpret_cast<const float4*>(p0 + step) : z4;
                const float4 v2 = r + 2 * m.rows_par < r1 ? *reinterpret_cast<const float4*>(p0 + 2 * step) : z4;
                const float4 v3 = r + 3 * m.rows_par < r1 ? *reinterpret_cast<const float4*>(p0 + 3 * step) : z4;
                a[0] += (v0.x + v1.x) + (v2.x + v3.x); q[0] += fmaf(v0.x, v0.x, v1.x * v1.x) + fmaf(v2.x, v2.x, v3.x * v3.x);
                a[1] += (v0.y + v1.y) + (v2.y + v3.y); q[1] += fmaf(v0.y, v0.y, v1.y * v1.y) + fmaf(v2.y, v2.y, v3.y * v3.y);
                a[2] += (v0.z + v1.z) + (v2.z + v3.z); q[2] += fmaf(v0.z, v0.z, v1.z * v1.z) + fmaf(v2.z, v2.z, v3.z * v3.z);
                a[3] += (v0.w + v1.w) + (v2.w + v3.w); q[3] += fmaf(v0.w, v0.w, v1.w * v1.w) + fmaf(v2.w, v2.w, v3.w * v3.w);
            }
            const int gA = (cv * 4) / cg;           // cg >= 2: a float4 column touches at most the groups gA and gA + 1
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if ((cv * 4 + e) / cg == gA) {
                    sA += a[e]; qA += q[e];
                } else {
                    sB += a[e]; qB += q[e];
                }
            }
        }
        s_part[s][0][threadIdx.x] = sA; s_part[s][1][threadIdx.x] = qA; s_part[s][2][threadIdx.x] = sB; s_part[s][3][threadIdx.x] = qB;
    }
    __syncthreads();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const bool wide = m.nvec >= GN_THREADS;
        for (int g = warp; g < 32; g += GN_THREADS / 32) {
            const int cvlo = (g * cg) >> 2, cvhi = ((g + 1) * cg - 1) >> 2;
            const int ncv = cvhi - cvlo + 1;
            const int items = m.rows_par * ncv;
            double sd = 0.0, qd = 0.0;
            for (int idx = lane; idx < items; idx += 32) {
                const int rs = idx / ncv, c = cvlo + (idx - rs * ncv);
                const int t = wide ? (c % GN_THREADS) : rs * m.nvec + c;
                const int sl = wide ? (c / GN_THREADS) : 0;
                const int tgA = (c * 4) / cg;
                if (tgA == g) {
                    sd += (double)s_part[sl][0][t]; qd += (double)s_part[sl][1][t];
                } else {                               // the column starts in group g - 1 and ends in g
                    sd += (double)s_part[sl][2][t]; qd += (double)s_part[sl][3][t];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sd += __shfl_xor_sync(0xffffffffu, sd, o);
                qd += __shfl_xor_sync(0xffffffffu, qd, o);
            }
            if (lane == 0) {
                float* w = ws + (((size_t)n * nslab + slab) * 32 + g) * 2;
                w[0] = (float)sd;
                w[1] = (float)qd;
            }
        }
    }
}

__global__ void __launch_bounds__(GN_THREADS, 2) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ ws,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ out, int rows, int C, int rows_per_slab,
                                                              float eps, int silu) {
    pdl_entry();
    __shared__ float s_mean[32], s_rstd[32];
    __shared__ double s_ps[16][32], s_pq[16][32];
    const int n = blockIdx.y, slab = blockIdx.x, nslab = gridDim.x;
    {   // reduce the per-slab partials: 16 threads per group in parallel, then a fixed-order sum (deterministic)
        const int g = threadIdx.x & 31, part = threadIdx.x >> 5;       // 512 threads = 16 parts x 32 groups
        double s = 0.0, q = 0.0;
        const float* w = ws + ((size_t)n * nslab * 32 + g) * 2;
        for (int i = part; i < nslab; i += 16) {
            s += (double)w[(size_t)i * 64];
            q += (double)w[(size_t)i * 64 + 1];
        }
        s_ps[part][g] = s;
        s_pq[part][g] = q;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            s += s_ps[i][threadIdx.x];
            q += s_pq[i][threadIdx.x];
        }
        const double cnt = (double)rows * (double)(C / 32);
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const GnMap m = gn_map(C);
    if (!m.active) return;
    const int cg = C / 32;
    const int r0 = slab * rows_per_slab;
    const int r1 = min(rows, r0 + rows_per_slab);
    const float* xs = x + (size_t)n * rows * C;
    __nv_bfloat16* os = out + (size_t)n * rows * C;
    for (int s = 0; s < m.slots; ++s) {
        const int cv = m.cv0 + s * GN_THREADS;
        if (cv >= m.nvec) break;
        float sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = cv * 4 + e;
            const int g = c / cg;
            sc[e] = s_rstd[g] * gamma[c];
            sh[e] = beta[c] - s_mean[g] * sc[e];
        }
        const size_t step = (size_t)m.rows_par * C;
        const float* col = xs + cv * 4;
        __nv_bfloat16* ocol = os + cv * 4;
        auto emit = [&](const float4 v, size_t off) {
            float y0 = fmaf(v.x, sc[0], sh[0]), y1 = fmaf(v.y, sc[1], sh[1]), y2 = fmaf(v.z, sc[2], sh[2]), y3 = fmaf(v.w, sc[3], sh[3]);
            if (silu) {
                y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3);
            }
            *reinterpret_cast<uint2*>(ocol + off) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
        };
        int r = r0 + m.rsub;
        for (; r < r1; r += 4 * m.rows_par) {      // last batch predicated: one round trip instead of a row-by-row tail
            const size_t o0 = (size_t)r * C;
            const bool k1 = r + m.rows_par < r1, k2 = r + 2 * m.rows_par < r1, k3 = r + 3 * m.rows_par < r1;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v0 = *reinterpret_cast<const float4*>(col + o0);
            const float4 v1 = k1 ? *reinterpret_cast<const float4*>(col + o0 + step) : z4;
            const float4 v2 = k2 ? *reinterpret_cast<const float4*>(col + o0 + 2 * step) : z4;
            const float4 v3 = k3 ? *reinterpret_cast<const float4*>(col + o0 + 3 * step) : z4;
            emit(v0, o0);
            if (k1) emit(v1, o0 + step);
            if (k2) emit(v2, o0 + 2 * step);
            if (k3) emit(v3, o0 + 3 * step);
        }
    }
}

static int gn_slabs(int ns, int rows, int C, int* rows_per_slab) {
    // ~2 CTAs per SM in total, but never fewer than 4 rows per thread (= one unrolled batch of independent 16-byte loads):
    // the small tensors of the 8x8 / 4x4 levels are latency-bound, so they want many CTAs with ONE round trip each rather
    // than a few CTAs walking 16 rows serially.
    int target = 296 / (ns > 0 ? ns : 1);
    if (target < 1) target = 1;
    const int nvec = C >> 2;
    const int rows_par = nvec >= GN_THREADS ? 1 : GN_THREADS / nvec;
    const int min_rows = 4 * rows_par;
    int slabs = (rows + min_rows - 1) / min_rows;
    if (slabs > target) slabs = target;
    if (slabs < 1) slabs = 1;
    *rows_per_slab = (rows + slabs - 1) / slabs;
    return (rows + *rows_per_slab - 1) / *rows_per_slab;
}

int64_t groupnorm_ws_floats(int ns, int rows, int C) {
    int rps;
    const int slabs = gn_slabs(ns, rows, C, &rps);
    return (int64_t)ns * slabs * 64;
}

int groupnorm_silu_launch(const float* x, const float* gamma, const float* beta, void* out, float* ws, int ns, int rows, int C, float eps,
                          int silu, cudaStream_t st) {
    if (C % 32 != 0 || C < 64 || (C >> 2) > GN_THREADS * GN_MAX_SLOTS || ns <= 0 || rows <= 0) return ERR_UNSUPPORTED;   // >= 2 channels per group
    if (ns > 65535) return ERR_UNSUPPORTED;
    int rps;
    const int slabs = gn_slabs(ns, rows, C, &rps);
    C2V_CHECK_CUDA(launch(gn_stats_kernel, dim3(slabs, ns), dim3(GN_THREADS), 0, st, x, ws, rows, C, rps));
    C2V_CHECK_CUDA(launch(gn_apply_kernel, dim3(slabs, ns), dim3(GN_THREADS), 0, st, x, (const float*)ws, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out),
                          rows, C, rps, eps, silu));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass mean / variance, fp32).
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 10;   // float4 per lane -> C <= 1280

// Each warp walks rows row0, row0 + stride, ...; the loads of the next row are issued before the statistics of the current
// one, so a warp keeps two rows of traffic in flight (a one-row-per-warp grid is latency-bound: 2.9 TB/s at 16384 x 320).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                        const float* __restrict__ add, __nv_bfloat16* __restrict__ out2,
                                                        float* __restrict__ out_f32, int rows, int C, float eps, int ld2) {
    pdl_entry();
    const int wpb = blockDim.x >> 5, stride = gridDim.x * wpb;
    int row = blockIdx.x * wpb + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 2;
    float4 v[NV], nx[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int cv = lane + i * 32;
        if (cv < nvec) v[i] = *reinterpret_cast<const float4*>(x + (size_t)row * C + cv * 4);
    }
    for (; row < rows; row += stride) {
        const int nrow = row + stride;
        if (nrow < rows) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int cv = lane + i * 32;
                if (cv < nvec) nx[i] = *reinterpret_cast<const float4*>(x + (size_t)nrow * C + cv * 4);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int cv = lane + i * 32;
            if (cv < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int cv = lane + i * 32;
            if (cv < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int cv = lane + i * 32;
            if (cv < nvec) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + cv * 4));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + cv * 4));
                const float y0 = (v[i].x - mean) * rstd * g.x + bb.x;
                const float y1 = (v[i].y - mean) * rstd * g.y + bb.y;
                const float y2 = (v[i].z - mean) * rstd * g.z + bb.z;
                const float y3 = (v[i].w - mean) * rstd * g.w + bb.w;
                *reinterpret_cast<uint2*>(out + (size_t)row * C + cv * 4) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
                if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * C + cv * 4) = make_float4(y0, y1, y2, y3);
                if (add) {
                    const float4 p = *reinterpret_cast<const float4*>(add + (size_t)row * C + cv * 4);
                    *reinterpret_cast<uint2*>(out2 + (size_t)row * ld2 + cv * 4) =
                        make_uint2(pack_bf16(y0 + p.x, y1 + p.y), pack_bf16(y2 + p.z, y3 + p.w));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = nx[i];
    }
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, const float* add, void* out2, float* out_f32,
                     int rows, int C, float eps, int ld2, cudaStream_t st) {
    if (C % 4 != 0 || (C >> 2) > LN_MAXV * 32 || rows <= 0) return ERR_UNSUPPORTED;
    if (ld2 == 0) ld2 = C;
    if (ld2 < C || ld2 % 4 != 0) return ERR_UNSUPPORTED;      // 8-byte stores into out2
    if (add && !out2) return ERR_BAD_ARG;
    const int nv = ((C >> 2) + 31) / 32;
    // few rows: 4-warp CTAs, one row per warp, spread over all SMs; many rows: 8-warp CTAs, 4 per SM, each warp walking its rows
    const int wpb = rows <= 148 * 8 * 2 ? 4 : 8;
    int grid = (rows + wpb - 1) / wpb;
    if (grid > 148 * 4) grid = 148 * 4;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(out2);
    if (nv <= 3)
        C2V_CHECK_CUDA(launch(layernorm_kernel<3>, dim3(grid), dim3(wpb * 32), 0, st, x, gamma, beta, o, add, o2, out_f32, rows, C, eps, ld2));
    else if (nv <= 5)
        C2V_CHECK_CUDA(launch(layernorm_kernel<5>, dim3(grid), dim3(wpb * 32), 0, st, x, gamma, beta, o, add, o2, out_f32, rows, C, eps, ld2));
    else
        C2V_CHECK_CUDA(launch(layernorm_kernel<LN_MAXV>, dim3(grid), dim3(wpb * 32), 0, st, x, gamma, beta, o, add, o2, out_f32, rows, C, eps, ld2));
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

// ------------------------------------------------------------------------------------------------
// Row softmax: out[r, :] = softmax(scale * x[r, :]) as 16-bit operands.  One warp per row, fp32 statistics.
// Used by the single-head, 512-wide attention block of the VAE decoder (R/lvdm/modules/networks/ae_modules.py:53-80), whose
// head dim does not fit the 64-wide flash kernels: scores and P.V go through the GEMM kernel, the softmax through this one.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int n,
                                                           float scale_log2) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)row * n;
    float mx = -INFINITY;
    for (int i = lane * 4; i < n; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m = mx * scale_log2;
    float sum = 0.f;
    for (int i = lane * 4; i < n; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        sum += (fast_exp2(fmaf(v.x, scale_log2, -m)) + fast_exp2(fmaf(v.y, scale_log2, -m))) +
               (fast_exp2(fmaf(v.z, scale_log2, -m)) + fast_exp2(fmaf(v.w, scale_log2, -m)));
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __nv_bfloat16* o = out + (size_t)row * n;
    for (int i = lane * 4; i < n; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        *reinterpret_cast<uint2*>(o + i) = make_uint2(pack_bf16(fast_exp2(fmaf(v.x, scale_log2, -m)) * inv, fast_exp2(fmaf(v.y, scale_log2, -m)) * inv),
                                                      pack_bf16(fast_exp2(fmaf(v.z, scale_log2, -m)) * inv, fast_exp2(fmaf(v.w, scale_log2, -m)) * inv));
    }
}

int softmax_rows_launch(const float* x, void* out, int rows, int n, float scale, cudaStream_t st) {
    if (rows <= 0 || n <= 0 || n % 4 != 0) return ERR_UNSUPPORTED;
    softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(out), rows, n, scale * 1.4426950408889634f);
    C2V_CHECK_CUDA(cudaGetLastError());
    return OK;
}

}  // namespace c2v
