// Fused two-GEMM MLP kernels for sm_100a: the hidden activation of Mlp (mpvit.py:71-78) never makes the
// HBM round trip between its two Linears.
//
//   mdv_mlp_fwd : out = residual + rowscale * dropout( GELU(A W1^T + b1) [dropout] W2^T + b2 )         (mpvit.py:72-77, mdvit.py:357-359)
//                 optionally also stores hact = dropout(GELU(.)) and u = GELU'(.) * mask/(1-p) for the backward pass
//   mdv_mlp_bwd : dA = ((dY W2) * u) W1     with du = (dY W2) * u optionally stored (weight gradients need it) and
//                 colsum(du) accumulated (fc1 bias gradient)
//
// Both are the same pipeline:  mid = f(A . B1^T) in 64-column chunks of the hidden dimension;  out += mid . B2^T.
// One persistent CTA per SM loops over 128-token tiles.
//   warp 0        TMA producer: the tile's A operand (K-major, 128B swizzle), a ring of weight stages {B1 chunk [64 x C], B2
//                 chunk [C x 64]} (L2-resident), and in the backward the u tile of the chunk
//   warp 1        tcgen05.mma issuer of GEMM1 (one thread), free-running up to two chunks ahead of the epilogue: the tensor
//                 core computes the next pre-activations while the epilogue warps are busy with the current ones
//   warp 3        tcgen05.mma issuer of GEMM2 (one thread).  Two issuing threads, because one thread's serial instruction
//                 stream (waits, descriptors, MMAs, commits for both GEMMs of every chunk) was the kernel's critical path.
//                 S (GEMM1 accumulators, 2 x 64 columns) and Y (GEMM2 accumulators, 2 x C columns) both live in TMEM
//   warp 2        TMEM allocator, then TMA-store issuer for the mid / u tiles (training only)
//   warps 4..19   epilogue, two groups of 8 warps working on alternate chunks (so the fixed latencies of one chunk hide under
//                 the other's): tcgen05.ld S -> bias, GELU (+ GELU', dropout) or multiply by u -> bf16 -> the 128B-swizzled
//                 K-major shared-memory tile that IS the A operand of GEMM2 (and the source box of the TMA store);
//                 per tile, all 16 warps: tcgen05.ld Y -> bias, dropout, DropPath scale, residual -> global
// C in {64, 128} (encoder / decoder stages 0 and 1: 85% of the model's MLP time); other widths use mdv_gemm_nt.
#include <stdlib.h>

#include "../../include/mdvit_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int HC = 64;
constexpr int NEPI = 16;
constexpr int THREADS = 32 * (4 + NEPI);
constexpr int MAX_WS = 4;
constexpr int TILE_BYTES = BM * HC * 2;      // one 128 x 64 bf16 K-major k-block: 16 KB

struct MlpParams {
    int M, C, hidden, n_chunks, m_tiles;
    int a_bufs, w_stages;
    int mode;                 // 0: forward (GELU), 1: backward (multiply by u)
    int store_mid, store_aux, load_aux;
    int gelu_grad;            // forward: also produce u = GELU'(x) * mask/(1-p)
    float drop_p;
    uint32_t drop_stream1, drop_stream2;
    const unsigned long long* rng;
    const float* bias1;
    const float* bias2;
    const float* residual;
    const float* rowscale;
    int rows_per_scale;
    float* colsum1;
    float* out;
};

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// MODE 0: forward (TRAIN: also u = GELU' * mask, dropout; !TRAIN: inference, nothing but GELU in the inner loop); MODE 1: backward
template <int MODE, bool TRAIN>
__global__ void __launch_bounds__(THREADS, 1)
    mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                     const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmMid,
                     const __grid_constant__ CUtensorMap tmAux, const MlpParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], w1_full[MAX_WS], w1_empty[MAX_WS], w2_full[MAX_WS], w2_empty[MAX_WS], s_full[2],
        s_empty[2], mid_full[2], mid_empty[2], aux_full[2], aux_empty[2], y_full[2], y_empty[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, KB = C / 64;
    const uint32_t a_bytes = (uint32_t)KB * TILE_BYTES;           // A tile: KB k-blocks of 128 x 64
    const uint32_t b1_bytes = (uint32_t)KB * (HC * 128);           // B1 chunk: KB k-blocks of 64 rows x 128 B
    const uint32_t b2_bytes = (uint32_t)C * 128;                   // B2 chunk: C rows x 128 B
    const uint32_t w_bytes = b1_bytes + b2_bytes;
    uint8_t* sA = smem;
    uint8_t* sW = sA + (size_t)p.a_bufs * a_bytes;
    uint8_t* sMid = sW + (size_t)p.w_stages * w_bytes;
    uint8_t* sAux = sMid + 2 * TILE_BYTES;
    // [hidden] floats: forward = fc1 bias (read from shared memory in the inner loop), backward = bias-gradient column sums
    float* cs_sm = reinterpret_cast<float*>(sAux + ((p.store_aux || p.load_aux) ? 2 * TILE_BYTES : 0));

    const int my_tiles = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * p.n_chunks;        // chunks this CTA processes, numbered g = tile * n_chunks + chunk

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB2)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], NEPI / 2);
            mbar_init(&mid_full[i], NEPI / 2);
            mbar_init(&mid_empty[i], 1 + ((p.store_mid || p.store_aux) ? 1 : 0));
            mbar_init(&aux_full[i], 1);
            mbar_init(&aux_empty[i], NEPI / 2);
            mbar_init(&y_full[i], 1);
            mbar_init(&y_empty[i], NEPI);
        }
        for (int i = 0; i < MAX_WS; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
            mbar_init(&w2_full[i], 1);
            mbar_init(&w2_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (MODE == 1 && p.colsum1) {
        for (int i = threadIdx.x; i < p.hidden; i += blockDim.x) cs_sm[i] = 0.f;
    }
    if (MODE == 0) {       // (parameters: not produced by the preceding kernel, safe to read before griddepcontrol.wait)
        for (int i = threadIdx.x; i < p.hidden; i += blockDim.x) cs_sm[i] = __ldg(p.bias1 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_y = tmem_base;                       // 2 x C columns
    const uint32_t tmem_s = tmem_base + 2u * (uint32_t)C;    // 2 x 64 columns
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int m0 = ((int)blockIdx.x + lt * (int)gridDim.x) * BM;
                const int ab = lt % p.a_bufs, ak = lt / p.a_bufs;
                mbar_wait_spin(&a_empty[ab], (ak & 1) ^ 1);
                mbar_expect_tx(&a_full[ab], a_bytes);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + (size_t)ab * a_bytes + kb * TILE_BYTES, &tmA, kb * 64, m0, &a_full[ab]);
                for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
                    const int ws = g % p.w_stages, wk = g / p.w_stages;
                    uint8_t* sb1 = sW + (size_t)ws * w_bytes;
                    mbar_wait_spin(&w1_empty[ws], (wk & 1) ^ 1);
                    mbar_expect_tx(&w1_full[ws], b1_bytes);
                    for (int kb = 0; kb < KB; ++kb) tma_load_2d(sb1 + kb * (HC * 128), &tmB1, kb * 64, ch * HC, &w1_full[ws]);
                    mbar_wait_spin(&w2_empty[ws], (wk & 1) ^ 1);
                    mbar_expect_tx(&w2_full[ws], b2_bytes);
                    tma_load_2d(sb1 + b1_bytes, &tmB2, ch * HC, 0, &w2_full[ws]);
                    if (p.load_aux) {
                        const int xb = g & 1;
                        mbar_wait_spin(&aux_empty[xb], ((g >> 1) & 1) ^ 1);
                        mbar_expect_tx(&aux_full[xb], TILE_BYTES);
                        tma_load_2d(sAux + xb * TILE_BYTES, &tmAux, ch * HC, m0, &aux_full[xb]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ GEMM1 issuer: S[g & 1] = A . B1(chunk)^T
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc(BM, HC, false);
            const uint64_t dbase = make_desc(0, 16, 1024);       // descriptor with a zero start address: add (addr >> 4)
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int ab = lt % p.a_bufs, ak = lt / p.a_bufs;
                mbar_wait_spin(&a_full[ab], ak & 1);
                const uint32_t a0 = smem_u32(sA + (size_t)ab * a_bytes);
                for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
                    const int ws = g % p.w_stages, wk = g / p.w_stages;
                    const int sb = g & 1;
                    mbar_wait_spin(&w1_full[ws], wk & 1);
                    mbar_wait_spin(&s_empty[sb], ((g >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t b0 = smem_u32(sW + (size_t)ws * w_bytes);
                    const uint32_t d = tmem_s + (uint32_t)sb * HC;
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t ad = dbase + ((a0 + kb * TILE_BYTES) >> 4);
                        const uint64_t bd = dbase + ((b0 + kb * (HC * 128)) >> 4);
                        tc_mma_bf16(d, ad, bd, idesc1, kb != 0 ? 1u : 0u);
                        tc_mma_bf16(d, ad + 2, bd + 2, idesc1, 1u);        // +32 B per K=16 step inside the 128 B swizzle span
                        tc_mma_bf16(d, ad + 4, bd + 4, idesc1, 1u);
                        tc_mma_bf16(d, ad + 6, bd + 6, idesc1, 1u);
                    }
                    tc_commit(&s_full[sb]);
                    tc_commit(&w1_empty[ws]);
                }
                tc_commit(&a_empty[ab]);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ GEMM2 issuer: Y[tile & 1] += mid(chunk) . B2(chunk)^T
        if (lane == 0) {
            const uint32_t idesc2 = make_idesc(BM, C, false);
            const uint64_t dbase = make_desc(0, 16, 1024);
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int yb = lt & 1;
                mbar_wait_spin(&y_empty[yb], ((lt >> 1) & 1) ^ 1);
                const uint32_t d = tmem_y + (uint32_t)yb * C;
                for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
                    const int ws = g % p.w_stages, wk = g / p.w_stages;
                    const int mb = g & 1;
                    mbar_wait_spin(&w2_full[ws], wk & 1);
                    mbar_wait_spin(&mid_full[mb], (g >> 1) & 1);
                    tc_fence_after();
                    const uint64_t ad = dbase + (smem_u32(sMid + mb * TILE_BYTES) >> 4);
                    const uint64_t bd = dbase + (smem_u32(sW + (size_t)ws * w_bytes + b1_bytes) >> 4);
                    tc_mma_bf16(d, ad, bd, idesc2, ch != 0 ? 1u : 0u);
                    tc_mma_bf16(d, ad + 2, bd + 2, idesc2, 1u);
                    tc_mma_bf16(d, ad + 4, bd + 4, idesc2, 1u);
                    tc_mma_bf16(d, ad + 6, bd + 6, idesc2, 1u);
                    tc_commit(&mid_empty[mb]);
                    tc_commit(&w2_empty[ws]);
                }
                tc_commit(&y_full[yb]);
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ TMA store issuer (mid / u tiles -> HBM)
        if (lane == 0 && (p.store_mid || p.store_aux)) {
            int g = 0;
            for (int lt = 0; lt < my_tiles; ++lt) {
                const int m0 = ((int)blockIdx.x + lt * (int)gridDim.x) * BM;
                for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
                    const int mb = g & 1;
                    mbar_wait_spin(&mid_full[mb], (g >> 1) & 1);
                    if (p.store_mid) tma_store_2d(&tmMid, sMid + mb * TILE_BYTES, ch * HC, m0);
                    if (p.store_aux) tma_store_2d(&tmAux, sAux + mb * TILE_BYTES, ch * HC, m0);
                    tma_commit();
                    tma_wait_read<0>();              // the tile has been read out of shared memory: the buffer may be rewritten
                    mbar_arrive(&mid_empty[mb]);
                }
            }
            tma_wait_all();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int e = warp - 4;
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int grp = (e >> 2) & 1;        // chunk parity this warp works on
        const int half = e >> 3;             // which 32 of the chunk's 64 columns
        const int sub = e >> 2;              // which quarter of Y's columns (final epilogue)
        const int r_tile = q * 32 + lane;    // row inside the tile
        uint32_t dthr = 0, dkey1 = 0, dkey2 = 0;
        float dinv = 1.0f;
        if (MODE == 0 && TRAIN && p.drop_p > 0.0f) {
            dthr = drop_thresh(p.drop_p);
            dinv = 1.0f / (1.0f - p.drop_p);
            dkey1 = rng_key(p.rng, p.drop_stream1);
            dkey2 = rng_key(p.rng, p.drop_stream2);
        }
        const uint32_t lane_taddr = (uint32_t)(q * 32) << 16;
        const uint32_t mid_addr = smem_u32(sMid), aux_addr = smem_u32(sAux), cs_addr = smem_u32(cs_sm);
        // swizzled offsets of this lane's four 16-byte chunks inside a 128 x 128B tile
        uint32_t off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) off[j] = (uint32_t)r_tile * 128u + (uint32_t)(((4 * half + j) ^ (r_tile & 7)) << 4);

        auto final_epilogue = [&](int lt) {
            const int yb = lt & 1;
            const int m0 = ((int)blockIdx.x + lt * (int)gridDim.x) * BM;
            const int row = m0 + r_tile;
            const bool row_ok = row < p.M;
            mbar_wait(&y_full[yb], (lt >> 1) & 1);
            tc_fence_after();
            const int ncol = C / 4;          // columns of Y this warp owns: [sub * ncol, +ncol)
            float rs = 1.0f;
            if (MODE == 0 && p.rowscale && row_ok) rs = __ldg(p.rowscale + row / p.rows_per_scale);
            for (int cc = 0; cc < ncol; cc += 16) {
                const int col0 = sub * ncol + cc;
                uint32_t v[16];
                tc_ld16(tmem_y + lane_taddr + (uint32_t)(yb * C + col0), v);
                tc_wait_ld();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (MODE == 0) {
                    if (p.bias2) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0 + j));
                            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                        }
                    }
                    if (dthr) {
                        const uint32_t pbase = (uint32_t)(((unsigned long long)row * (unsigned)C + (unsigned)col0) >> 1);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t h = drop_hash(dkey2, pbase + j);
                            f[2 * j] *= drop_lo(h, dthr, dinv);
                            f[2 * j + 1] *= drop_hi(h, dthr, dinv);
                        }
                    }
                    if (p.rowscale) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] *= rs;
                    }
                    if (p.residual && row_ok) {
                        const float* rp = p.residual + (size_t)row * C + col0;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                            f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                        }
                    }
                }
                if (row_ok) {
                    float* op = p.out + (size_t)row * C + col0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&y_empty[yb]);
        };

        int finalized = 0;                   // tiles [0, finalized) have had their output written by this warp
        for (int g = grp; g < total; g += 2) {
            const int lt = g / p.n_chunks, ch = g - lt * p.n_chunks;
            const int m0 = ((int)blockIdx.x + lt * (int)gridDim.x) * BM;
            const int row = m0 + r_tile;
            const bool row_ok = row < p.M;
            const int sb = g & 1;            // == grp: each group owns one S stage, one mid buffer, one u buffer
            const uint32_t par = (uint32_t)(g >> 1) & 1u;
            mbar_wait(&s_full[sb], par);
            tc_fence_after();
            uint32_t v[32];
            tc_ld32(tmem_s + lane_taddr + (uint32_t)(sb * HC + half * 32), v);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);      // S is in registers: the tensor core may overwrite this stage
            const int col0 = ch * HC + half * 32;           // first hidden column of this lane's 32 values
            uint32_t pk[16];
            if (MODE == 0) {
                uint32_t uk[TRAIN ? 16 : 1];
                uint32_t hw = 0;
                const uint32_t pbase = (uint32_t)(((unsigned long long)row * (unsigned)p.hidden + (unsigned)col0) >> 1);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const uint4 b4 = lds128(cs_addr + (uint32_t)(col0 + 4 * j4) * 4u);      // broadcast read
                    const float2 bb[2] = {make_float2(__uint_as_float(b4.x), __uint_as_float(b4.y)),
                                          make_float2(__uint_as_float(b4.z), __uint_as_float(b4.w))};
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int j = 2 * j4 + u;                 // pair index: columns 2j, 2j+1
                        const float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb[u]);
                        float2 gl, gr;
                        gelu_pair<TRAIN>(x, &gl, &gr);
                        if (TRAIN) {
                            if (dthr) {
                                // the hidden-activation mask lives only in this kernel (u carries it to the backward pass), so it
                                // need not follow the library-wide one-hash-per-pair convention: one full hash per 8 elements, then
                                // a multiply-add + xorshift step per further pair (3 instructions instead of 10; issue-bound loop)
                                if ((j & 3) == 0) hw = drop_hash(dkey1, pbase + j);
                                else {
                                    hw = hw * 0x9E3779B1u + 0x7F4A7C15u;
                                    hw ^= hw >> 15;
                                }
                                const uint32_t h = hw;
                                const float2 sc = make_float2(drop_lo(h, dthr, dinv), drop_hi(h, dthr, dinv));
                                gl = __fmul2_rn(gl, sc);
                                gr = __fmul2_rn(gr, sc);
                            }
                            uk[j] = f2_to_bf2(gr.x, gr.y);
                        }
                        pk[j] = f2_to_bf2(gl.x, gl.y);
                    }
                }
                mbar_wait(&mid_empty[sb], par ^ 1);
                const uint32_t mt = mid_addr + sb * TILE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) sts128(mt + off[j], make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
                if (TRAIN) {
                    const uint32_t at = aux_addr + sb * TILE_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts128(at + off[j], make_uint4(uk[4 * j], uk[4 * j + 1], uk[4 * j + 2], uk[4 * j + 3]));
                }
            } else {
                mbar_wait(&aux_full[sb], par);
                const uint32_t ua = aux_addr + sb * TILE_BYTES;
                uint4 uq[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) uq[j] = lds128(ua + off[j]);
                // (the u buffer is handed back to the TMA producer further down, AFTER the stores that consume these loads)
                float d[32];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t uu[4] = {uq[j].x, uq[j].y, uq[j].z, uq[j].w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = 8 * j + 2 * u;
                        const float2 r = __fmul2_rn(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), bf2_to_f2(uu[u]));
                        pk[4 * j + u] = f2_to_bf2(r.x, r.y);
                        d[c] = r.x;
                        d[c + 1] = r.y;
                    }
                }
                mbar_wait(&mid_empty[sb], par ^ 1);
                const uint32_t mt = mid_addr + sb * TILE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) sts128(mt + off[j], make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
                // Hand the u buffer back to the TMA producer only now: the stores above DEPEND on the u loads, so the loads have
                // returned.  An mbarrier arrive does not wait for the thread's outstanding shared-memory loads (it is executed by
                // a different unit than the LSU queue they sit in) and MEMBAR.CTA did not close the window either: arriving
                // right after issuing the loads let the next u tile land under loads still in flight — a few lanes of a
                // warp then saw rows of the wrong tile, in roughly one run out of six.
                __syncwarp();
                if (lane == 0) mbar_arrive(&aux_empty[sb]);
                if (p.colsum1) {
                    // column sums of this warp's 32 x 32 block by a transposing butterfly: after the step with offset `o` a
                    // lane keeps half of its columns, each summed over twice as many rows; lane l ends with column l
                    if (!row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) d[j] = 0.f;
                    }
#pragma unroll
                    for (int w = 16, o = 16; w >= 1; w >>= 1, o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < w; ++i) {
                            const float send = upper ? d[i] : d[i + w];
                            const float keep = upper ? d[i + w] : d[i];
                            d[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    atomicAdd(cs_sm + col0 + lane, d[0]);
                }
            }
            fence_async_smem();          // generic-proxy writes -> visible to the tensor core / TMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&mid_full[sb]);
            // outputs of the tiles before this one: their last GEMM2 has had this chunk's epilogue time to finish
            while (finalized < lt) final_epilogue(finalized++);
        }
        while (finalized < my_tiles) final_epilogue(finalized++);
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == 1 && p.colsum1) {
        for (int i = threadIdx.x; i < p.hidden; i += blockDim.x) {
            const float s = cs_sm[i];
            if (s != 0.f) atomicAdd(p.colsum1 + i, s);
        }
    }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int MODE, bool TRAIN>
int launch_mlp(const CUtensorMap& ta, const CUtensorMap& tb1, const CUtensorMap& tb2, const CUtensorMap& tmid, const CUtensorMap& taux,
               MlpParams& p, cudaStream_t st) {
    const int KB = p.C / 64;
    const size_t a_bytes = (size_t)KB * TILE_BYTES, w_bytes = (size_t)256 * p.C;
    const size_t fixed = 2 * TILE_BYTES + ((p.store_aux || p.load_aux) ? 2 * TILE_BYTES : 0) + (size_t)p.hidden * 4 + 1024;
    const size_t budget = 226 * 1024 - 1024;
    p.a_bufs = 2;
    p.w_stages = MAX_WS;
    while (fixed + p.a_bufs * a_bytes + p.w_stages * w_bytes > budget) {
        if (p.w_stages > 3) --p.w_stages;
        else if (p.a_bufs > 1) --p.a_bufs;
        else if (p.w_stages > 2) --p.w_stages;
        else return MDV_ERR_UNSUPPORTED;
    }
    const size_t smem = fixed + p.a_bufs * a_bytes + p.w_stages * w_bytes;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<MODE, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024));
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int grid = p.m_tiles < MDV_NUM_SMS ? p.m_tiles : MDV_NUM_SMS;
    mdv_launch((mlp_fused_kernel<MODE, TRAIN>), dim3(grid), dim3(THREADS), smem, st, ta, tb1, tb2, tmid, taux, p);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

int mlp_shape_ok(int M, int C, int hidden) {
    if (M <= 0 || (C != 64 && C != 128) || hidden < HC || hidden % HC || hidden > 2048) return 0;
    return 1;
}

}  // namespace

extern "C" int mdv_mlp_supported(int C, int hidden) { return mlp_shape_ok(1, C, hidden); }

extern "C" int mdv_mlp_fwd(const void* a, const void* w1, const float* b1, const void* w2, const float* b2, const float* residual,
                           float* out, void* hact_out, void* u_out, int M, int C, int hidden, float drop_p, const void* rng,
                           uint32_t drop_stream1, uint32_t drop_stream2, const float* rowscale, int rows_per_scale, void* stream) {
    if (!a || !w1 || !b1 || !w2 || !out) return MDV_ERR_ARG;
    if (!mlp_shape_ok(M, C, hidden)) return MDV_ERR_UNSUPPORTED;
    if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !rng) || (rowscale && rows_per_scale <= 0)) return MDV_ERR_ARG;
    MlpParams p = {};
    p.M = M; p.C = C; p.hidden = hidden; p.n_chunks = hidden / HC; p.m_tiles = mdv_cdiv(M, BM);
    p.mode = 0;
    p.store_mid = hact_out != nullptr;
    p.store_aux = u_out != nullptr;
    p.gelu_grad = u_out != nullptr;
    p.drop_p = drop_p; p.drop_stream1 = drop_stream1; p.drop_stream2 = drop_stream2;
    p.rng = (const unsigned long long*)rng;
    p.bias1 = b1; p.bias2 = b2; p.residual = residual; p.rowscale = rowscale; p.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
    p.out = out;
    CUtensorMap ta, tb1, tb2, tmid, taux;
    int rc = make_map(&ta, a, 2, C, M, C, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tb1, w1, 2, C, hidden, C, 64, HC, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tb2, w2, 2, hidden, C, hidden, 64, C, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    tmid = ta; taux = ta;
    if (hact_out) {
        rc = make_map(&tmid, hact_out, 2, hidden, M, hidden, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (u_out) {
        rc = make_map(&taux, u_out, 2, hidden, M, hidden, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (hact_out || u_out || drop_p > 0.f) {
        if (!hact_out || !u_out) return MDV_ERR_ARG;       // training: both saved tensors
        return launch_mlp<0, true>(ta, tb1, tb2, tmid, taux, p, (cudaStream_t)stream);
    }
    return launch_mlp<0, false>(ta, tb1, tb2, tmid, taux, p, (cudaStream_t)stream);
}

extern "C" int mdv_mlp_bwd(const void* dy, const void* w2t, const void* u, const void* w1t, void* du_out, float* dx_out, float* colsum1,
                           int M, int C, int hidden, void* stream) {
    if (!dy || !w2t || !u || !w1t || !dx_out) return MDV_ERR_ARG;
    if (!mlp_shape_ok(M, C, hidden)) return MDV_ERR_UNSUPPORTED;
    MlpParams p = {};
    p.M = M; p.C = C; p.hidden = hidden; p.n_chunks = hidden / HC; p.m_tiles = mdv_cdiv(M, BM);
    p.mode = 1;
    p.store_mid = du_out != nullptr;
    p.load_aux = 1;
    p.colsum1 = colsum1;
    p.rows_per_scale = 1;
    p.out = dx_out;
    CUtensorMap ta, tb1, tb2, tmid, taux;
    int rc = make_map(&ta, dy, 2, C, M, C, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tb1, w2t, 2, C, hidden, C, 64, HC, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&tb2, w1t, 2, hidden, C, hidden, 64, C, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map(&taux, u, 2, hidden, M, hidden, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    tmid = taux;
    if (du_out) {
        rc = make_map(&tmid, du_out, 2, hidden, M, hidden, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    return launch_mlp<1, false>(ta, tb1, tb2, tmid, taux, p, (cudaStream_t)stream);
}
