// Backward of softmax(Q K^T) V attention with the DA head gate (Attention_Sup.forward, TransFuse_S_adapt's DeiT branch:
// Models/Hybrid_models/TransFuseFolder/vision_transformer.py:149-169).  Forward: csrc/sdpa.cu.
//
//   Y = g * (P V),  P = softmax(s Q K^T)            (g: per-channel gate of this image, s: head_dim^-0.5)
//   dO = g * dY;   D_i = sum_c dO_ic O_ic = sum_c dY_ic Y_ic;   dgate_c = sum_i dY_ic Y_ic / g_c
//   dP = dO V^T;   dS = P * (dP - D);   dQ = s dS K;   dK = s dS^T Q;   dV = P^T dO
//
// One CTA (8 warps) per (head, image); N <= 256 tokens of 64 channels, so Q, K, V and dO of the head live in shared memory as
// bf16 rows (pitch padded to 144 B: conflict-free ldmatrix).  Everything is bf16 mma.sync m16n8k16 with fp32 accumulation, P is
// recomputed from the forward's row log-sum-exp.  Two passes, no atomics: pass A gives every warp 16-query tiles (dQ), pass B
// 16-key tiles (dK, dV); the accumulator layout of two adjacent n-tiles IS the A-fragment layout of the next product, so P and
// dS never leave registers.  An HBM-light op (reads qkv, Y, dY; writes dqkv): the 7 small GEMMs per head are 59 MFLOP.
#include "../../include/mdvit_b200.h"
#include "common.cuh"

namespace {

constexpr int D = 64;
constexpr int PITCH = 144;      // bytes per staged row (64 bf16 + 16 B pad)
constexpr int MAXN = 256;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

struct Frag16 {      // A operand, 16 rows x 64 columns: 4 k-steps
    uint32_t a[4][4];
};
// rows r0..r0+15 of a staged [row][64] matrix as A fragments (row-major: plain ldmatrix)
__device__ __forceinline__ void load_a(Frag16& f, uint32_t base, int r0, int lane) {
    // matrices: (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
    const uint32_t off = (uint32_t)((r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm4(f.a[ks], base + off + ks * 32);
}
// B fragments of T[r0..r0+15][ks*16..+15]^T, i.e. B(k = column, n = row) for two n-tiles (rows r0..+7, r0+8..+15): the stored
// row-major [row][column] matrix is exactly the "col" operand layout -> plain ldmatrix.  b[0..1]: n-tile 0, b[2..3]: n-tile 1
__device__ __forceinline__ void load_b_rows(uint32_t (&b)[4], uint32_t base, int r0, int ks, int lane) {
    // matrices: (rows 0-7, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 0-7), (rows 8-15, k 8-15)
    const uint32_t off = (uint32_t)((r0 + (lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 16 + ks * 32);
    ldsm4(b, base + off);
}
// B fragments of T[r0..r0+15][n0..n0+15] as B(k = row, n = column) for two n-tiles (columns n0..+7, n0+8..+15): needs .trans
__device__ __forceinline__ void load_b_cols(uint32_t (&b)[4], uint32_t base, int r0, int n0, int lane) {
    // matrices: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
    const uint32_t off = (uint32_t)((r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (n0 + (lane >> 4) * 8) * 2);
    ldsm4t(b, base + off);
}

__global__ void __launch_bounds__(256, 1) sdpa_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ gate,
                                                           const bf16* __restrict__ yout, const float* __restrict__ lse,
                                                           const bf16* __restrict__ dy, bf16* __restrict__ dqkv,
                                                           float* __restrict__ dgate, int N, int C, int heads, float scale) {
    MDV_PDL_SYNC();
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + MAXN * PITCH;
    uint8_t* sV = sK + MAXN * PITCH;
    uint8_t* sG = sV + MAXN * PITCH;                                  // dO = g * dY
    float* sL = reinterpret_cast<float*>(sG + MAXN * PITCH);          // row log-sum-exp
    float* sD = sL + MAXN;                                            // D_i
    float* sCol = sD + MAXN;                                          // [32][64] partial column sums of dY * Y
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x, b = blockIdx.y;
    const size_t tok0 = (size_t)b * N;

    {   // ---- stage: thread = 8 channels (16 B) of rows prow, prow + 32, ...
        const int prt = threadIdx.x & 7, prow = threadIdx.x >> 3;
        const int c0 = h * D + prt * 8;
        float g8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g8[j] = gate ? gate[(size_t)b * C + c0 + j] : 1.f;
        float col[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) col[j] = 0.f;
        for (int r = prow; r < N; r += 32) {
            const bf16* qrow = qkv + (tok0 + r) * 3 * C + c0;
            const uint4 q = *reinterpret_cast<const uint4*>(qrow);
            const uint4 k = *reinterpret_cast<const uint4*>(qrow + C);
            const uint4 v = *reinterpret_cast<const uint4*>(qrow + 2 * C);
            const uint4 dyv = *reinterpret_cast<const uint4*>(dy + (tok0 + r) * C + c0);
            const uint4 yv = *reinterpret_cast<const uint4*>(yout + (tok0 + r) * C + c0);
            const uint32_t so = (uint32_t)(r * PITCH + prt * 16);
            *reinterpret_cast<uint4*>(sQ + so) = q;
            *reinterpret_cast<uint4*>(sK + so) = k;
            *reinterpret_cast<uint4*>(sV + so) = v;
            const uint32_t d4[4] = {dyv.x, dyv.y, dyv.z, dyv.w}, y4[4] = {yv.x, yv.y, yv.z, yv.w};
            uint32_t o4[4];
            float rs = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float2 dd = bf2_to_f2(d4[w]), yy = bf2_to_f2(y4[w]);
                const float p0 = dd.x * yy.x, p1 = dd.y * yy.y;
                rs += p0 + p1;
                col[2 * w] += p0;
                col[2 * w + 1] += p1;
                o4[w] = f2_to_bf2(dd.x * g8[2 * w], dd.y * g8[2 * w + 1]);
            }
            *reinterpret_cast<uint4*>(sG + so) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            rs += __shfl_xor_sync(0xffffffffu, rs, 1);
            rs += __shfl_xor_sync(0xffffffffu, rs, 2);
            rs += __shfl_xor_sync(0xffffffffu, rs, 4);
            if (prt == 0) {
                sD[r] = rs;
                sL[r] = lse[((size_t)b * heads + h) * N + r];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sCol[prow * D + prt * 8 + j] = col[j];
    }
    __syncthreads();
    if (dgate && gate && threadIdx.x < D) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s += sCol[r * D + threadIdx.x];
        dgate[(size_t)b * C + h * D + threadIdx.x] = s / gate[(size_t)b * C + h * D + threadIdx.x];
    }
    const uint32_t bQ = (uint32_t)__cvta_generic_to_shared(sQ), bK = (uint32_t)__cvta_generic_to_shared(sK),
                   bV = (uint32_t)__cvta_generic_to_shared(sV), bG = (uint32_t)__cvta_generic_to_shared(sG);
    const int g = lane >> 2, t = lane & 3;
    const int ntile = N / 16;

    // ---- pass A: dQ.  A warp owns 16 query rows and walks the keys 16 at a time.
    for (int rt = warp; rt < ntile; rt += 8) {
        const int r0 = rt * 16;
        Frag16 fq, fg;
        load_a(fq, bQ, r0, lane);
        load_a(fg, bG, r0, lane);
        const float l0 = sL[r0 + g], l1 = sL[r0 + g + 8], d0 = sD[r0 + g], d1 = sD[r0 + g + 8];
        float dq[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) dq[n][j] = 0.f;
        for (int kb = 0; kb < ntile; ++kb) {
            const int k0 = kb * 16;
            float s[2][4], dp[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[n][j] = dp[n][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t bk[4], bv[4];
                load_b_rows(bk, bK, k0, ks, lane);
                load_b_rows(bv, bV, k0, ks, lane);
                mma16816(s[0], fq.a[ks], bk[0], bk[1]);
                mma16816(s[1], fq.a[ks], bk[2], bk[3]);
                mma16816(dp[0], fg.a[ks], bv[0], bv[1]);
                mma16816(dp[1], fg.a[ks], bv[2], bv[3]);
            }
            uint32_t a[4];      // dS (16 queries x 16 keys) as the A operand of dS . K
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const float p0 = __expf(scale * s[n][0] - l0), p1 = __expf(scale * s[n][1] - l0);
                const float p2 = __expf(scale * s[n][2] - l1), p3 = __expf(scale * s[n][3] - l1);
                a[2 * n] = f2_to_bf2(p0 * (dp[n][0] - d0), p1 * (dp[n][1] - d0));
                a[2 * n + 1] = f2_to_bf2(p2 * (dp[n][2] - d1), p3 * (dp[n][3] - d1));
            }
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t bb[4];
                load_b_cols(bb, bK, k0, n2 * 16, lane);
                mma16816(dq[2 * n2], a, bb[0], bb[1]);
                mma16816(dq[2 * n2 + 1], a, bb[2], bb[3]);
            }
        }
        bf16* o0 = dqkv + (tok0 + r0 + g) * 3 * C + h * D + 2 * t;
        bf16* o1 = o0 + (size_t)8 * 3 * C;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<uint32_t*>(o0 + n * 8) = f2_to_bf2(scale * dq[n][0], scale * dq[n][1]);
            *reinterpret_cast<uint32_t*>(o1 + n * 8) = f2_to_bf2(scale * dq[n][2], scale * dq[n][3]);
        }
    }

    // ---- pass B: dK, dV.  A warp owns 16 keys and walks the queries 16 at a time (everything transposed).
    for (int kt = warp; kt < ntile; kt += 8) {
        const int k0 = kt * 16;
        Frag16 fk, fv;
        load_a(fk, bK, k0, lane);
        load_a(fv, bV, k0, lane);
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) dk[n][j] = dv[n][j] = 0.f;
        for (int qb = 0; qb < ntile; ++qb) {
            const int q0 = qb * 16;
            float st[2][4], dpt[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) st[n][j] = dpt[n][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t bq[4], bg[4];
                load_b_rows(bq, bQ, q0, ks, lane);
                load_b_rows(bg, bG, q0, ks, lane);
                mma16816(st[0], fk.a[ks], bq[0], bq[1]);
                mma16816(st[1], fk.a[ks], bq[2], bq[3]);
                mma16816(dpt[0], fv.a[ks], bg[0], bg[1]);
                mma16816(dpt[1], fv.a[ks], bg[2], bg[3]);
            }
            uint32_t ap[4], as[4];      // P^T and dS^T (16 keys x 16 queries) as A operands
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int qc = q0 + n * 8 + 2 * t;      // the two query columns this thread holds in n-tile n
                const float l0 = sL[qc], l1 = sL[qc + 1], d0 = sD[qc], d1 = sD[qc + 1];
                const float p0 = __expf(scale * st[n][0] - l0), p1 = __expf(scale * st[n][1] - l1);
                const float p2 = __expf(scale * st[n][2] - l0), p3 = __expf(scale * st[n][3] - l1);
                ap[2 * n] = f2_to_bf2(p0, p1);
                ap[2 * n + 1] = f2_to_bf2(p2, p3);
                as[2 * n] = f2_to_bf2(p0 * (dpt[n][0] - d0), p1 * (dpt[n][1] - d1));
                as[2 * n + 1] = f2_to_bf2(p2 * (dpt[n][2] - d0), p3 * (dpt[n][3] - d1));
            }
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t bb[4];
                load_b_cols(bb, bG, q0, n2 * 16, lane);
                mma16816(dv[2 * n2], ap, bb[0], bb[1]);
                mma16816(dv[2 * n2 + 1], ap, bb[2], bb[3]);
                load_b_cols(bb, bQ, q0, n2 * 16, lane);
                mma16816(dk[2 * n2], as, bb[0], bb[1]);
                mma16816(dk[2 * n2 + 1], as, bb[2], bb[3]);
            }
        }
        bf16* o0 = dqkv + (tok0 + k0 + g) * 3 * C + C + h * D + 2 * t;
        bf16* o1 = o0 + (size_t)8 * 3 * C;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            *reinterpret_cast<uint32_t*>(o0 + n * 8) = f2_to_bf2(scale * dk[n][0], scale * dk[n][1]);
            *reinterpret_cast<uint32_t*>(o1 + n * 8) = f2_to_bf2(scale * dk[n][2], scale * dk[n][3]);
            *reinterpret_cast<uint32_t*>(o0 + C + n * 8) = f2_to_bf2(dv[n][0], dv[n][1]);
            *reinterpret_cast<uint32_t*>(o1 + C + n * 8) = f2_to_bf2(dv[n][2], dv[n][3]);
        }
    }
}

}  // namespace

extern "C" int mdv_sdpa_bwd(const void* qkv_bf16, const float* gate, const void* out_bf16, const float* lse, const void* dout_bf16,
                            void* dqkv_bf16, float* dgate, int B, int N, int C, int heads, float scale, void* stream) {
    if (!qkv_bf16 || !out_bf16 || !lse || !dout_bf16 || !dqkv_bf16 || B <= 0) return MDV_ERR_ARG;
    if (heads <= 0 || C != heads * D || (N != 128 && N != 256)) return MDV_ERR_UNSUPPORTED;
    const size_t smem = (size_t)4 * MAXN * PITCH + 2 * MAXN * sizeof(float) + 32 * D * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sdpa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    mdv_launch(sdpa_bwd_kernel, dim3(heads, B), dim3(256), smem, (cudaStream_t)stream, (const bf16*)qkv_bf16, gate, (const bf16*)out_bf16, lse,
               (const bf16*)dout_bf16, (bf16*)dqkv_bf16, dgate, N, C, heads, scale);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
