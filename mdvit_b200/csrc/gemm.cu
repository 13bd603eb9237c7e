// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   mdv_gemm_nt : C[M,N] = epilogue( A[M,K] . W[N,K]^T )      (both operands K-major; Linear / 1x1 conv fwd + dgrad)
//   mdv_gemm_tn : C[P,Q] += A[R,P]^T . B[R,Q]                 (both operands MN-major; weight gradients, split over R,
//                                                              fp32 atomics into the gradient buffer)
//
// One CTA computes one 128 x BN output tile.  Warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> smem transpose -> coalesced global stores).
// Operand tiles are staged in 128B-swizzled shared memory by TMA through an mbarrier ring; the fp32 accumulator
// lives in TMEM.  Replaces the aten::addmm / cudnn 1x1-conv calls behind nn.Linear / nn.Conv2d(k=1) in
// reference Models/Transformer/mdvit.py:288,310, mpvit.py:72-76, Decoders.py:197,59,317-333.
#include <cuda.h>

#include "../../include/mdvit_b200.h"
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 6;
constexpr int EPI_ROW_F = 68;                       // padded fp32 row of the 32x64 staging tile
constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_F * 4;  // 8704
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;       // 34816

struct GemmParams {
    int M, N, K;          // NT: output M x N, reduce K.  TN: output P(=M) x Q(=N), reduce R(=K)
    int stages;
    int kb_per_split;     // TN: k-blocks per blockIdx.z
    MdvGemmEpi epi;
    int atomic;           // TN: atomicAdd into fp32 out
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=BF16 [7,10),[10,13), a_major 15, b_major 16,
// N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 1u << 7;
    d |= 1u << 10;
    if (mn_major) d |= (1u << 15) | (1u << 16);
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

__device__ __forceinline__ void epilogue_pair(const GemmParams& p, float v0, float v1, int row, int col, uint32_t dkey,
                                              uint32_t dthr, float dinv) {
    const MdvGemmEpi& e = p.epi;
    if (e.bias) {
        v0 += __ldg(e.bias + col);
        v1 += __ldg(e.bias + col + 1);
    }
    if (e.out_preact) *reinterpret_cast<uint32_t*>((bf16*)e.out_preact + (size_t)row * e.ld_preact + col) = f2_to_bf2(v0, v1);
    if (e.act == MDV_ACT_GELU) {
        v0 = gelu_erf(v0);
        v1 = gelu_erf(v1);
    }
    if (e.mul_gelu_grad) {
        float2 u = bf2_to_f2(*reinterpret_cast<const uint32_t*>((const bf16*)e.mul_gelu_grad + (size_t)row * e.ld_mul + col));
        v0 *= gelu_erf_grad(u.x);
        v1 *= gelu_erf_grad(u.y);
    }
    if (dthr) {
        unsigned long long idx = (unsigned long long)row * (unsigned)p.N + (unsigned)col;
        v0 *= drop_scale(dkey, idx, dthr, dinv);
        v1 *= drop_scale(dkey, idx + 1, dthr, dinv);
    }
    if (e.rowscale) {
        float s = __ldg(e.rowscale + row / e.rows_per_scale);
        v0 *= s;
        v1 *= s;
    }
    if (e.residual) {
        float2 r = *reinterpret_cast<const float2*>(e.residual + (size_t)row * e.ld_res + col);
        v0 += r.x;
        v1 += r.y;
    }
    if (p.atomic) {
        float* o = (float*)e.out + (size_t)row * e.ldc + col;
        atomicAdd(o, v0);
        atomicAdd(o + 1, v1);
    } else if (e.out_bf16) {
        *reinterpret_cast<uint32_t*>((bf16*)e.out + (size_t)row * e.ldc + col) = f2_to_bf2(v0, v1);
    } else {
        float2* o = reinterpret_cast<float2*>((float*)e.out + (size_t)row * e.ldc + col);
        if (e.accumulate) {
            float2 old = *o;
            v0 += old.x;
            v1 += old.y;
        }
        *o = make_float2(v0, v1);
    }
}

template <int BN, bool TN>
__global__ void __launch_bounds__(192) gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    constexpr uint32_t A_BYTES = BM * BK * 2;
    constexpr uint32_t B_BYTES = BN * BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    const int total_kb = (p.K + BK - 1) / BK;
    int kb0 = 0, num_kb = total_kb;
    if (TN) {
        kb0 = blockIdx.z * p.kb_per_split;
        num_kb = min(p.kb_per_split, total_kb - kb0);
    }
    const int stages = p.stages;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < stages; ++i) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], 1);
            }
            mbar_init(&tmem_full_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (num_kb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int s = kb % stages;
                    const uint32_t ph = (kb / stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    const int kc = (kb0 + kb) * BK;
                    if (!TN) {
                        tma_load_2d(sa, &tmA, kc, m_tile * BM, &full_bar[s]);
                        tma_load_2d(sb, &tmB, kc, n_tile * BN, &full_bar[s]);
                    } else {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, m_tile * BM + c * 64, kc, &full_bar[s]);
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, n_tile * BN + c * 64, kc, &full_bar[s]);
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = make_idesc(BM, BN, TN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int s = kb % stages;
                    const uint32_t ph = (kb / stages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        uint64_t ad, bd;
                        if (!TN) {  // K-major: 8-row groups 1024B apart; +32B per K=16 step inside the 128B swizzle span
                            ad = make_desc(sa + k * 32, 16, 1024);
                            bd = make_desc(sb + k * 32, 16, 1024);
                        } else {    // MN-major: 64-element MN chunks 8192B apart (LBO), 8 k-rows = 1024B (SBO); K=16 -> +2048B
                            ad = make_desc(sa + k * 2048, 8192, 1024);
                            bd = make_desc(sb + k * 2048, 8192, 1024);
                        }
                        tc_mma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&tmem_full_bar);
            }
        } else {
            mbar_wait(&tmem_full_bar, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int row_base = m_tile * BM + q * 32;
            float* stg = reinterpret_cast<float*>(smem + (size_t)(warp - 2) * EPI_WARP_BYTES);
            const int n0 = n_tile * BN;
            uint32_t dthr = 0, dkey = 0;
            float dinv = 1.0f;
            if (p.epi.dropout_p > 0.0f) {
                dthr = drop_thresh(p.epi.dropout_p);
                dinv = 1.0f / (1.0f - p.epi.dropout_p);
                dkey = rng_key((const unsigned long long*)p.epi.rng, p.epi.drop_stream);
            }
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 64) {
                if (n0 + c0 >= p.N) break;
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tc_ld32(taddr, v);
                tc_ld32(taddr + 32, v + 32);
                tc_wait_ld();
                float4* dst = reinterpret_cast<float4*>(stg + lane * EPI_ROW_F);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                         __uint_as_float(v[4 * j + 3]));
                __syncwarp();
                const int col = n0 + c0 + 2 * lane;
                if (col < p.N) {
                    const int rmax = min(32, p.M - row_base);
                    for (int r = 0; r < rmax; ++r) {
                        float2 a = *reinterpret_cast<const float2*>(stg + r * EPI_ROW_F + 2 * lane);
                        epilogue_pair(p, a.x, a.y, row_base + r, col, dkey, dthr, dinv);
                    }
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch `ld` elements.
int make_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return MDV_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return MDV_ERR_ARG;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MDV_OK : MDV_ERR_DRIVER;
}

int g_force_bn = 0, g_force_stages = 0, g_force_split = 0;

int pick_bn(int N, long long m_tiles) {
    if (g_force_bn) return g_force_bn;
    const int cands[3] = {256, 128, 64};
    int best = 64;
    double best_w = 1e9;
    for (int i = 0; i < 3; ++i) {
        int bn = cands[i];
        double w = (double)mdv_cdiv(N, bn) * bn / N;
        if (w <= 1.07) {
            best = bn;
            best_w = w;
            break;
        }
        if (w < best_w - 1e-9) {
            best_w = w;
            best = bn;
        }
    }
    // small grids: prefer more, narrower tiles so that all 148 SMs get work
    while (best > 64 && m_tiles * mdv_cdiv(N, best) < MDV_NUM_SMS) best >>= 1;
    return best;
}

template <int BN, bool TN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid, cudaStream_t st) {
    const size_t stage_bytes = (size_t)(BM * BK * 2 + BN * BK * 2);
    size_t smem = (size_t)p.stages * stage_bytes;
    if (smem < (size_t)EPI_BYTES) smem = EPI_BYTES;
    smem += 1024;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        if (e != cudaSuccess) return (int)e;
        configured = 200 * 1024;
    }
    gemm_kernel<BN, TN><<<grid, 192, smem, st>>>(ta, tb, p);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

}  // namespace

extern "C" int mdv_gemm_tune(int force_bn, int force_stages, int force_split) {
    g_force_bn = force_bn;
    g_force_stages = force_stages;
    g_force_split = force_split;
    return MDV_OK;
}

extern "C" int mdv_gemm_nt(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const MdvGemmEpi* epi,
                           void* stream) {
    if (!A || !W || !epi || !epi->out || M <= 0 || N <= 0 || K <= 0) return MDV_ERR_ARG;
    if ((N & 1) || (K & 7) || (epi->ldc & 1)) return MDV_ERR_ARG;
    const long long m_tiles = mdv_cdiv(M, BM);
    if (m_tiles > 65535) return MDV_ERR_UNSUPPORTED;
    const int bn = pick_bn(N, m_tiles);
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.epi = *epi;
    p.atomic = 0;
    p.kb_per_split = 0;
    const int kb = mdv_cdiv(K, BK);
    int stages = bn == 256 ? 4 : (bn == 128 ? 4 : 6);
    if (g_force_stages) stages = g_force_stages;
    p.stages = stages < kb ? stages : kb;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, K, M, lda, BK, BM);
    if (rc) return rc;
    rc = make_map(&tb, W, K, N, ldw, BK, bn);
    if (rc) return rc;
    dim3 grid(mdv_cdiv(N, bn), (unsigned)m_tiles, 1);
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 256: return launch<256, false>(ta, tb, p, grid, st);
        case 128: return launch<128, false>(ta, tb, p, grid, st);
        default: return launch<64, false>(ta, tb, p, grid, st);
    }
}

extern "C" int mdv_gemm_tn(const void* A, int lda, const void* B, int ldb, int R, int P, int Q, float* C, int ldc,
                           void* stream) {
    if (!A || !B || !C || R <= 0 || P <= 0 || Q <= 0) return MDV_ERR_ARG;
    if ((P & 7) || (Q & 7) || (ldc & 1)) return MDV_ERR_ARG;
    const long long p_tiles = mdv_cdiv(P, BM);
    int bn = g_force_bn ? g_force_bn : (Q % 256 == 0 || Q > 512 ? 256 : (Q % 128 == 0 ? 128 : 64));
    if (Q <= 64) bn = 64;
    const int q_tiles = mdv_cdiv(Q, bn);
    const int kb = mdv_cdiv(R, BK);
    // split the reduction so that the grid is a few waves of the 148 SMs; >= 4 k-blocks per CTA
    int want = (4 * MDV_NUM_SMS) / (int)(p_tiles * q_tiles);
    if (want < 1) want = 1;
    int splits = kb / 4 < want ? (kb / 4 > 0 ? kb / 4 : 1) : want;
    if (g_force_split) splits = g_force_split;
    if (splits > kb) splits = kb;
    GemmParams p;
    p.M = P; p.N = Q; p.K = R;
    p.kb_per_split = mdv_cdiv(kb, splits);
    splits = mdv_cdiv(kb, p.kb_per_split);
    MdvGemmEpi e = {};
    e.out = C;
    e.ldc = ldc;
    p.epi = e;
    p.atomic = 1;
    int stages = g_force_stages ? g_force_stages : 4;
    p.stages = stages < p.kb_per_split ? stages : p.kb_per_split;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, P, R, lda, 64, BK);
    if (rc) return rc;
    rc = make_map(&tb, B, Q, R, ldb, 64, BK);
    if (rc) return rc;
    dim3 grid(q_tiles, (unsigned)p_tiles, splits);
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 256: return launch<256, true>(ta, tb, p, grid, st);
        case 128: return launch<128, true>(ta, tb, p, grid, st);
        default: return launch<64, true>(ta, tb, p, grid, st);
    }
}
