// Shared device helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 / TMEM
// wrappers written as inline PTX, plus small math utilities.  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// 16-bit operand type of every tensor-core GEMM / attention operand and of every bf16 activation buffer.
// Default: bfloat16.  -DC2V_OPERAND_FP16 builds the same kernels with IEEE half operands (what the reference's own
// "16-mixed" autocast uses): 11 significand bits instead of 8, i.e. ~8x smaller operand rounding error at identical
// speed; accumulation, statistics, softmax and the residual stream stay fp32 either way.  The kernels are written against
// the bf16 names; this block re-points those names at the half type, so one flag flips the whole library consistently.
// ------------------------------------------------------------------------------------------------
#ifdef C2V_OPERAND_FP16
#define __nv_bfloat16 __half
#define __nv_bfloat162 __half2
#define __bfloat1622float2 __half22float2
#define __float2bfloat16 __float2half
#define __floats2bfloat162_rn __floats2half2_rn
#define C2V_OPERAND_IS_FP16 1
#else
#define C2V_OPERAND_IS_FP16 0
#endif

namespace c2v {

// ------------------------------------------------------------------------------------------------
// status codes returned by every extern "C" entry point (include/camc2v_b200.h)
// ------------------------------------------------------------------------------------------------
enum Status : int { OK = 0, ERR_BAD_ARG = 1, ERR_CUDA = 2, ERR_TMA_ENCODE = 3, ERR_UNSUPPORTED = 4 };

#define C2V_CHECK_CUDA(expr)                      \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return c2v::ERR_CUDA; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the step is launched with the programmatic-stream-serialization
// attribute (c2v::launch below) and starts with pdl_entry(): `launch_dependents` lets the NEXT kernel of the stream / graph branch
// be scheduled as soon as every CTA of this one is resident, so its prologue (barrier init, TMEM allocation, descriptor prefetch)
// and its launch latency overlap this kernel's tail; `wait` then blocks until the PREVIOUS kernel has completed and its writes are
// visible.  Rule: no global memory access (read or write) before pdl_entry().  C2V_PDL=0 in the environment turns the
// attribute off (the two instructions are then no-ops) for A/B measurements.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("C2V_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint expires) instead of
// returning at once, so a waiting warp issues no instructions - a software poll loop (try_wait + nanosleep + branch) of the
// TMA / MMA warps measured ~30 % of all issued instructions in the attention kernel and stole issue slots from the softmax warps.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 10000u) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU (each failed try_wait has already slept up to 10 us).
// The template parameter is kept for the call sites that used to select a nanosleep back-off; the hardware suspend makes it moot.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 1000000u) {
            printf("camc2v_b200: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA tiled loads (global -> shared), completion on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// TMA tiled store (shared -> global), bulk async-group completion (per issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrive on `bar` once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
// For 128B-swizzled tiles whose rows are exactly 128 B (64 bf16): 8-row groups are 1024 B apart (SBO),
// LBO is unused.  Valid for K-major operands (rows = M/N index) and for MN-major operands (rows = K index).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled layouts)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation
// (cute::UMMA::InstrDescriptor): c_format[4,6)=1 (F32), a_format[7,10)=1, b_format[10,13)=1 (BF16),
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    // c_format = f32 (bit 4); a_format / b_format (bits 7 / 10): 1 = bf16, 0 = f16
    return (1u << 4) | ((C2V_OPERAND_IS_FP16 ? 0u : 1u) << 7) | ((C2V_OPERAND_IS_FP16 ? 0u : 1u) << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 (or 16) consecutive fp32 columns; thread i gets lane (row) i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// misc math
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Packed fp32 pairs (sm_100 FFMA2 / FADD2): two IEEE fp32 operations per issue slot, bit-identical to the scalar forms.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// x * gelu_erf(g) with erf by Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7): 1 RCP + 1 EX2 + 12 FMA-class ops instead of
// libdevice erff.  Used where the result is rounded to bf16 anyway (GEGLU gate, attention.py:437-438).
//   z = |g|/sqrt2, t = 1/(1 + p z), erf|.| = 1 - poly(t) exp(-z^2);  x*gelu(g) = h + h*erf(g) with h = 0.5*g*x
__device__ __forceinline__ float geglu_fast(float x, float g) {
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(g), 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float e = fast_exp2((g * g) * (-0.5f * 1.4426950408889634f));
    const float erfv = copysignf(fmaf(-poly, e, 1.0f), g);
    const float h = (0.5f * g) * x;
    return fmaf(h, erfv, h);
}
__device__ __forceinline__ float gelu_erf_fast(float g) { return geglu_fast(1.0f, g); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// Two geglu_fast() at once on packed fp32 pairs (FFMA2 / FMUL2: one issue slot per pair for 11 of the 15 FMA-class operations; the
// two MUFU operations, |g| and copysign stay scalar).  Same operation sequence per element as geglu_fast, with the polynomial
// evaluated on negated coefficients (fma(-a, t, -b) == -fma(a, t, b) exactly), so the results are bit-identical to the scalar form.
__device__ __forceinline__ float2 geglu_fast2(float2 x, float2 g) {
    float2 t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(g.x), 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(g.y), 1.0f)));
    float2 q = ffma2(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
    q = ffma2(q, t, make_float2(-1.421413741f, -1.421413741f));
    q = ffma2(q, t, make_float2(0.284496736f, 0.284496736f));
    q = ffma2(q, t, make_float2(-0.254829592f, -0.254829592f));
    q = fmul2(q, t);                                                   // = -poly(t) * t
    const float2 ea = fmul2(fmul2(g, g), make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
    const float2 e = make_float2(fast_exp2(ea.x), fast_exp2(ea.y));
    const float2 r = ffma2(q, e, make_float2(1.0f, 1.0f));             // 1 - poly * e
    const float2 erfv = make_float2(copysignf(r.x, g.x), copysignf(r.y, g.y));
    const float2 h = fmul2(fmul2(make_float2(0.5f, 0.5f), g), x);
    return ffma2(h, erfv, h);
}
// ---- thread-block clusters: barrier + distributed shared memory ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cta_smem_addr, uint32_t cta_rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(cta_smem_addr), "r"(cta_rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace c2v
