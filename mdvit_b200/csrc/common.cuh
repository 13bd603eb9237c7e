// Shared device helpers for the mdvit_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define MDV_OK 0
#define MDV_ERR_ARG (-1)
#define MDV_ERR_UNSUPPORTED (-2)
#define MDV_ERR_DRIVER (-3)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
extern long long g_mdv_launches;
#define MDV_CHECK_LAUNCH()                          \
    do {                                            \
        ++g_mdv_launches;                           \
        cudaError_t e__ = cudaGetLastError();       \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

#define MDV_NUM_SMS 148

// ----------------------------------------------------------------------------- programmatic dependent launch (PDL)
// A training step is ~4500 mostly short kernels on one stream; ~28% of its time is batch-independent launch / ramp
// latency.  Every kernel is launched with the programmatic-stream-serialization attribute and begins with
// griddepcontrol.wait (blocks until the preceding grid has completed and its writes are visible) followed by
// griddepcontrol.launch_dependents: the NEXT kernel's launch, block scheduling and prologue then overlap this kernel's
// execution instead of following it.  The wait is the first instruction, so no dependent read or write can be early.
#define MDV_PDL_SYNC()                                                  \
    do {                                                                \
        asm volatile("griddepcontrol.wait;" ::: "memory");              \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    } while (0)

extern int g_mdv_pdl;   // 1: launch with the PDL attribute (default), 0: plain launches (MDV_NO_PDL=1 or mdv_set_pdl(0))

template <typename... KArgs, typename... Args>
static inline cudaError_t mdv_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_mdv_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int mdv_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------- math
// Exact-erf GELU (nn.GELU default, mpvit.py:61).  erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. fp32
// round-off level) on the MUFU units: one ex2 + one rcp + 7 FMAs, a third of libdevice erff's instruction count — the
// GEMM epilogues that apply it are issue-bound (8 warps/SM), see DESIGN.md.
// Returns Phi(x) = 0.5 (1 + erf(x / sqrt 2)); *e_out = exp(-x^2 / 2).
__device__ __forceinline__ float gauss_cdf(float x, float* e_out) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    const float e = __expf(-z * z);
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float h = 0.5f * poly * t * e;          // 0.5 * erfc(|z|)
    *e_out = e;
    return x >= 0.0f ? 1.0f - h : h;
}
__device__ __forceinline__ float gelu_erf(float x) {
    float e;
    return x * gauss_cdf(x, &e);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float e;
    const float phi = gauss_cdf(x, &e);
    return fmaf(x * 0.3989422804014327f, e, phi);
}
// Packed (two elements per instruction, Blackwell FFMA2/FMUL2/FADD2) versions for the issue-bound GEMM epilogues.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GELU(x) and (optionally) GELU'(x) for two elements with the fewest issue slots (the fused MLP epilogue is issue-bound):
//   a = |x|, t = 1/(1 + p a/sqrt2), e = exp(-x^2/2), hh = 0.5 erfc(a/sqrt2) = (poly(t) t e) with 0.5 folded into poly
//   GELU(x)  = relu(x) - a hh                     (x >= 0: x (1 - hh);  x < 0: x hh)
//   GELU'(x) = Phi(x) + x phi(x),  Phi = x >= 0 ? 1 - hh : hh,  phi = e / sqrt(2 pi)
// Same A&S 7.1.26 erf as gauss_cdf2 (|abs err| <= 1.5e-7).
template <bool GRAD>
__device__ __forceinline__ void gelu_pair(float2 x, float2* g, float2* dg) {
    const float ax = fabsf(x.x), ay = fabsf(x.y);
    const float2 t = make_float2(rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f)),
                                 rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, ay, 1.0f)));
    const float2 w = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f));   // -x^2/2 * log2(e)
    const float2 e = make_float2(ex2_approx(w.x), ex2_approx(w.y));
    float2 poly = __ffma2_rn(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    poly = __ffma2_rn(poly, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    const float2 hh = __fmul2_rn(__fmul2_rn(poly, t), e);
    // relu(x) - |x| hh   (the |.| and the negation are free source modifiers of the scalar FFMA)
    g->x = fmaf(-ax, hh.x, fmaxf(x.x, 0.0f));
    g->y = fmaf(-ay, hh.y, fmaxf(x.y, 0.0f));
    if (GRAD) {
        const float2 phi = make_float2(x.x >= 0.0f ? 1.0f - hh.x : hh.x, x.y >= 0.0f ? 1.0f - hh.y : hh.y);
        *dg = __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), e, phi);
    }
}
__device__ __forceinline__ float2 gauss_cdf2(float2 x, float2* e_out) {
    const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752f, 0.70710678118654752f));
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
    const float2 t = make_float2(__fdividef(1.0f, d.x), __fdividef(1.0f, d.y));
    const float2 w = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    const float2 e = make_float2(ex2_approx(w.x), ex2_approx(w.y));                       // exp(-z^2)
    float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
    poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
    poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
    // g = 0.5 - 0.5 erfc(|z|) >= 0;  Phi = 0.5 + copysign(g, x)
    const float2 h = __fmul2_rn(__fmul2_rn(poly, t), e);
    const float2 g = __ffma2_rn(h, make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));
    *e_out = e;
    return __fadd2_rn(make_float2(copysignf(g.x, x.x), copysignf(g.y, x.y)), make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    float2 g, dg;
    gelu_pair<false>(x, &g, &dg);
    return g;
}
__device__ __forceinline__ float2 gelu_erf_grad2(float2 x) {
    float2 e;
    const float2 phi = gauss_cdf2(x, &e);
    return __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), e, phi);
}

__device__ __forceinline__ float hardswish_f(float x) { return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f); }
__device__ __forceinline__ float hardswish_grad(float x) {
    return x <= -3.0f ? 0.0f : (x >= 3.0f ? 1.0f : (2.0f * x + 3.0f) * (1.0f / 6.0f));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

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

// Block-wide sum of `v`; result valid in every thread. `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.0f;
    r = warp_sum(r);
    return r;
}

// ----------------------------------------------------------------------------- counter-based RNG
// Dropout / DropPath masks are a pure function of (seed, step, stream, element index) so the
// backward pass regenerates them instead of storing them.  rng[0] = seed, rng[1] = step counter
// (device memory, bumped once per optimizer step so that CUDA-graph replays see fresh masks).
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t rng_key(const unsigned long long* rng, uint32_t stream) {
    unsigned long long s = rng ? rng[0] : 0ull, c = rng ? rng[1] : 0ull;
    uint32_t k = mix32((uint32_t)s ^ 0x9E3779B9u);
    k = mix32(k ^ (uint32_t)(s >> 32));
    k = mix32(k ^ (uint32_t)c * 0x27D4EB2Fu);
    k = mix32(k ^ stream * 0x165667B1u);
    return k;
}
// One 32-bit hash decides TWO consecutive elements (16 bits each): element idx is kept iff the (idx & 1)-th half of
// hash(idx >> 1) is >= thresh16 = round(p * 65536).  Kernels that walk consecutive elements hash once per pair.
__device__ __forceinline__ uint32_t drop_hash(uint32_t key, uint32_t pair) { return mix32(pair * 0x9E3779B1u ^ key); }
__device__ __forceinline__ float drop_lo(uint32_t h, uint32_t thresh16, float inv_keep) { return (h & 0xFFFFu) >= thresh16 ? inv_keep : 0.0f; }
__device__ __forceinline__ float drop_hi(uint32_t h, uint32_t thresh16, float inv_keep) { return (h >> 16) >= thresh16 ? inv_keep : 0.0f; }
// keep-scale for element `idx`: 0 if dropped, 1/(1-p) if kept.
__device__ __forceinline__ float drop_scale(uint32_t key, unsigned long long idx, uint32_t thresh16, float inv_keep) {
    const uint32_t h = drop_hash(key, (uint32_t)(idx >> 1));
    return (idx & 1ull) ? drop_hi(h, thresh16, inv_keep) : drop_lo(h, thresh16, inv_keep);
}
__host__ __device__ __forceinline__ uint32_t drop_thresh(float p) {
    const float t = p * 65536.0f + 0.5f;
    return t >= 65535.0f ? 65535u : (uint32_t)t;
}

// ----------------------------------------------------------------------------- vector io
__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(b);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&r);
}
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
