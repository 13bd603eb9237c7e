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

static inline int mdv_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------- math
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
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
// keep-scale for element `idx`: 0 if dropped, 1/(1-p) if kept. thresh = p * 2^32.
__device__ __forceinline__ float drop_scale(uint32_t key, unsigned long long idx, uint32_t thresh, float inv_keep) {
    uint32_t h = mix32((uint32_t)idx * 0x9E3779B1u ^ key);
    h = mix32(h ^ (uint32_t)(idx >> 32) ^ 0x632BE5ABu);
    return h >= thresh ? inv_keep : 0.0f;
}
__host__ __device__ __forceinline__ uint32_t drop_thresh(float p) {
    double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
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
