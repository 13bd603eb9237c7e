// softmax(Q K^T * scale) V with the domain-adapter head gate, on tcgen05 / TMEM / TMA (sm_100a).
//
// This is the attention of TransFuse_S_adapt's DeiT-S branch — Attention_Sup.forward between the qkv and proj Linears
// (Models/Hybrid_models/TransFuseFolder/vision_transformer.py:149-169): N = 256 tokens (16 x 16 patches of a 256 x 256
// image), 6 heads x 64, followed by the same softmax-over-heads DA gate as MDViT's factorized attention
// (vision_transformer.py:160-164).  BASELINE.json's north_star names it "the QK^T.softmax.V attention".
//
// One CTA per (image, head) handles NTILE = N / 128 query tiles of 128 rows that SHARE the K / V tiles (one load, and the softmax of
// one tile overlaps the MMAs of the other: the single-tile version was a serial load -> MMA -> softmax -> MMA -> store chain with
// one CTA per SM); the whole key/value sequence of a head (N <= 256) is resident, so there is no
// online-softmax rescaling: S = Q K^T is ONE 128 x N UMMA into TMEM, the softmax runs row-per-thread out of TMEM, P goes to
// shared memory as the bf16 K-major A operand of the second UMMA, O = P V accumulates in TMEM, and the epilogue applies
// 1/rowsum and the gate.
//   last warp            TMEM allocator; one thread issues the TMA loads (Q tiles, K, V: 128B-swizzled) and all MMAs
//   warps 4t .. 4t+3     softmax + epilogue of query tile t: thread = query row (TMEM lane); tcgen05.ld 32 columns at a time
// The O accumulator of a tile reuses the first 64 TMEM columns of its S tile (S is dead once P has been written).
// V is consumed as an MN-major B operand straight from its [token, channel] layout (no transpose anywhere).
#include "../../include/mdvit_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace {
using namespace tc;

constexpr int BM = 128;      // query rows per CTA
constexpr int D = 64;        // head dim
constexpr int MAXN = 256;    // keys per head

struct SdpaParams {
    int B, N, C, heads;
    float scale_log2e;       // scale * log2(e)
    const float* gate;       // [B, C] or NULL
    bf16* out;               // [B, N, C]
    float* lse;              // [B, heads, N] (row max * scale + ln(rowsum)) or NULL
};

template <int NTILE>
__global__ void __launch_bounds__(32 * (4 * NTILE + 1), 1)
    sdpa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const SdpaParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t ld_bar, s_bar[NTILE], p_bar[NTILE], o_bar[NTILE];
    __shared__ uint32_t tmem_base_slot;
    constexpr int CW = 4 * NTILE;                // the control warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.N, C = p.C;
    const int h = blockIdx.y, b = blockIdx.z;
    const int nkb = N / 64;                      // 64-key blocks
    uint8_t* sQ = smem;                          // NTILE x [128 x 64] bf16 K-major                 16 KB each
    uint8_t* sK = sQ + NTILE * BM * D * 2;       // [N x 64] bf16 K-major (B operand of S = Q K^T)  32 KB
    uint8_t* sV = sK + MAXN * D * 2;             // nkb x [64 keys x 64 ch] MN-major chunks         32 KB
    uint8_t* sP = sV + MAXN * D * 2;             // NTILE x nkb x [128 x 64] bf16 K-major k-blocks  64 KB each

    if (warp == CW && lane == 0) {
        mbar_init(&ld_bar, 1);
        for (int t = 0; t < NTILE; ++t) {
            mbar_init(&s_bar[t], 1);
            mbar_init(&p_bar[t], 4);
            mbar_init(&o_bar[t], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    }
    if (warp == CW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // tile t: S in TMEM columns [256 t, 256 t + N); its O accumulator later takes the first 64 of them
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == CW) {
        if (lane == 0) {
            const int row0 = b * N;                    // first token row of this image in the [B*N, 3C] matrix
            const int q0 = blockIdx.x * NTILE * BM;    // first query row of this CTA inside the image
            mbar_expect_tx(&ld_bar, (uint32_t)(NTILE * BM * D * 2 + 2 * N * D * 2));
            for (int t = 0; t < NTILE; ++t) tma_load_2d(sQ + t * (BM * D * 2), &tmQ, h * D, row0 + q0 + t * BM, &ld_bar);
            tma_load_2d(sK, &tmK, C + h * D, row0, &ld_bar);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sV + kb * 8192, &tmV, 2 * C + h * D, row0 + kb * 64, &ld_bar);
            mbar_wait(&ld_bar, 0);
            tc_fence_after();
            // S[128, N] = Q[128, 64] . K[N, 64]^T   (both K-major; +32 B per K=16 step)
            const uint64_t dk = make_desc(0, 16, 1024);
            const uint32_t idesc1 = make_idesc(BM, N, false);
            for (int t = 0; t < NTILE; ++t) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem_base_slot + t * MAXN, dk + ((smem_u32(sQ) + t * (BM * D * 2) + k * 32) >> 4), dk + ((smem_u32(sK) + k * 32) >> 4),
                                idesc1, k != 0 ? 1u : 0u);
                tc_commit(&s_bar[t]);
            }
            // O[128, 64] = P[128, N] . V[N, 64]: A K-major from sP, B MN-major (keys are its K dimension, rows of sV)
            const uint32_t idesc2 = make_idesc(BM, D, false) | (1u << 16);
            const uint64_t dv = make_desc(0, 8192, 1024);
            for (int t = 0; t < NTILE; ++t) {
                mbar_wait(&p_bar[t], 0);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_bf16(tmem_base_slot + t * MAXN, dk + ((smem_u32(sP) + (t * (MAXN / 64) + kb) * (BM * 128) + k * 32) >> 4),
                                    dv + ((smem_u32(sV) + kb * 8192 + k * 2048) >> 4), idesc2, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(&o_bar[t]);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax + epilogue: thread = query row of tile t
        const int t = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int mt = blockIdx.x * NTILE + t;
        const uint32_t tmem_s = tmem_base_slot + t * MAXN, tmem_o = tmem_s;
        const uint32_t lane_taddr = (uint32_t)((warp & 3) * 32) << 16;
        mbar_wait(&s_bar[t], 0);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_s + lane_taddr + (uint32_t)c0, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        const float mneg = -mx * p.scale_log2e;
        float sum = 0.f;
        const uint32_t p_addr = smem_u32(sP) + (uint32_t)t * (MAXN / 64) * (BM * 128);
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_s + lane_taddr + (uint32_t)c0, v);
            tc_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, mneg));
                const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, mneg));
                pk[j] = f2_to_bf2(e0, e1);
                const float2 rq = bf2_to_f2(pk[j]);          // the row sum is that of the ROUNDED probabilities the MMA will see
                sum += rq.x + rq.y;
            }
            // K-major 128B-swizzled k-block kb = c0 / 64; this chunk covers 16-byte parts (c0 % 64) / 8 .. +3 of row r
            const uint32_t base = p_addr + (uint32_t)(c0 >> 6) * (BM * 128) + (uint32_t)r * 128u;
            const int part0 = (c0 & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) sts128(base + (uint32_t)(((part0 + j) ^ (r & 7)) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_bar[t]);
        const float inv = 1.0f / sum;
        const int n = mt * BM + r;                       // token index inside the image
        if (p.lse) p.lse[((size_t)b * p.heads + h) * N + n] = mx * p.scale_log2e * 0.6931471805599453f + __logf(sum);
        mbar_wait(&o_bar[t], 0);
        tc_fence_after();
        bf16* orow = p.out + ((size_t)b * N + n) * C + h * D;
        const float* grow = p.gate ? p.gate + (size_t)b * C + h * D : nullptr;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_o + lane_taddr + (uint32_t)c0, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t pk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + 8 * j + 2 * u;
                    float a0 = __uint_as_float(v[8 * j + 2 * u]) * inv, a1 = __uint_as_float(v[8 * j + 2 * u + 1]) * inv;
                    if (grow) {
                        a0 *= __ldg(grow + c);
                        a1 *= __ldg(grow + c + 1);
                    }
                    pk[u] = f2_to_bf2(a0, a1);
                }
                *reinterpret_cast<uint4*>(orow + c0 + 8 * j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "r"(512u) : "memory");
    }
}

}  // namespace

extern "C" int mdv_sdpa_fwd(const void* qkv_bf16, const float* gate, void* out_bf16, float* lse, int B, int N, int C, int heads, float scale,
                            void* stream) {
    if (!qkv_bf16 || !out_bf16 || B <= 0) return MDV_ERR_ARG;
    if (heads <= 0 || C != heads * D || (N != 128 && N != 256)) return MDV_ERR_UNSUPPORTED;
    SdpaParams p = {};
    p.B = B; p.N = N; p.C = C; p.heads = heads;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.gate = gate;
    p.out = (bf16*)out_bf16;
    p.lse = lse;
    CUtensorMap tq, tk, tv;
    const long long rows = (long long)B * N;
    int rc = make_map(&tq, qkv_bf16, 2, 3 * C, rows, 3 * C, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tk, qkv_bf16, 2, 3 * C, rows, 3 * C, 64, N, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tv, qkv_bf16, 2, 3 * C, rows, 3 * C, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool configured = false;
    const size_t smem1 = BM * D * 2 + 2 * MAXN * D * 2 + (size_t)(MAXN / 64) * BM * 128 + 1024;
    const size_t smem2 = 2 * BM * D * 2 + 2 * MAXN * D * 2 + (size_t)2 * (MAXN / 64) * BM * 128 + 1024;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sdpa_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sdpa_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (N == 2 * BM) mdv_launch(sdpa_fwd_kernel<2>, dim3(1, heads, B), dim3(32 * 9), smem2, (cudaStream_t)stream, tq, tk, tv, p);
    else mdv_launch(sdpa_fwd_kernel<1>, dim3(N / BM, heads, B), dim3(32 * 5), smem1, (cudaStream_t)stream, tq, tk, tv, p);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
