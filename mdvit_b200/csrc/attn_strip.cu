// Factorized attention, second generation: "strip" kernels.
//
// lane == one PAIR of adjacent channels (one bf16x2 word), so a warp spans 64 channels with 128-byte coalesced global
// accesses and conflict-free 128-byte shared-memory rows, and every fp32 multiply-add is a packed FFMA2 over the pair.
// A thread owns a horizontal strip of TX = 8 pixels: each staged input value is reused for up to WIN taps x TX outputs
// from registers, so the 3x3 / 5x5 / 7x7 depthwise relative-position convolution (mpvit.py:296-318) costs ~0.2 shared
// loads per multiply-add instead of 1.  The per-head Ch x Ch matrix products (Q.A, and in backward dF.A^T, S.dA, V.dA^T)
// are done with warp shuffles inside the lanes of a head.  Cross-token reductions (K^T V and Q^T dF) write per-chunk
// partials that a second tiny kernel sums in a fixed order: no atomics, bit-reproducible forward.
//
// Channel groups: a warp handles CPW = 64 channels (8 heads at Ch=8, 4 at Ch=16, 1 at Ch=64) or one 40-channel head
// (20 active lanes) at Ch=40.  The convolution window of a group is that of its last head (windows grow with the head).
#include <cstdlib>

#include "attn_internal.cuh"

namespace {

constexpr int TX = 8;        // pixels per strip
constexpr int TILE_W = 16;   // spatial tile (pixels) handled by one block
constexpr int TILE_H = 16;

template <int CH>
struct Cfg {
    static constexpr int LPH = CH / 2;                        // lanes per head
    static constexpr int HPW = CH <= 16 ? 32 / LPH : 1;       // heads per warp
    static constexpr int ACT = LPH * HPW;                     // active lanes (32, 32, 20, 32)
    static constexpr int CPW = 2 * ACT;                       // channels per warp / group
    static constexpr int KR = CH <= 16 ? CH : CH / 2;         // k-range per block of the outer-product kernel
    static constexpr int KSPLIT = CH / KR;
    static_assert(KR * KSPLIT == CH, "k-range split");
};

__host__ __device__ inline int win_of_head(int h) { return h < 2 ? 3 : (h < 5 ? 5 : 7); }

__device__ __forceinline__ float2 up2(uint32_t v) { return bf2_to_f2(v); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// per-lane view of the filter bank: (window, pointer to this channel's taps, pointer to its bias)
__device__ __forceinline__ void crpe_of_channel(const CrpeW& cw, int c, int CH, const float*& w, const float*& b, int& win) {
    const int h = c / CH;
    const int grp = h < 2 ? 0 : (h < 5 ? 1 : 2);
    const int cl = c - (grp == 0 ? 0 : (grp == 1 ? 2 * CH : 5 * CH));
    win = 3 + 2 * grp;
    // (selects, not cw.w[grp]: a dynamically indexed kernel parameter is copied to local memory)
    w = (grp == 0 ? cw.w[0] : (grp == 1 ? cw.w[1] : cw.w[2])) + (size_t)cl * win * win;
    b = (grp == 0 ? cw.b[0] : (grp == 1 ? cw.b[1] : cw.b[2])) + cl;
}

// ------------------------------------------------------------------------------------------------ cross-token sums
// MODE 0:  part[k,v] = sum_n exp(K[n,k]-kmax[k]) V[n,v];  zpart[k] = sum_n exp(K[n,k]-kmax[k])      (forward, App. E)
// MODE 1:  part[k,v] = sum_n Q[n,k] (g[v] dY[n,v])                                                  (backward dA / scale)
// grid = (token chunk, group * KSPLIT + ksplit, B).  A warp walks its tokens 4 at a time; lane (pair v) accumulates
// acc[k] for the KR values of k of this split, fetching P[n,k] from the lane that owns k with shuffles.
template <int CH, int MODE>
__global__ void __launch_bounds__(256) attn_outer_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                          const float* __restrict__ gate, const float* __restrict__ kmax,
                                                          float* __restrict__ part, float* __restrict__ zpart, int N, int C,
                                                          int rows_per_block, int nchunk) {
    MDV_PDL_SYNC();
    using G = Cfg<CH>;
    constexpr int KR = G::KR;
    __shared__ float2 red[KR][32];
    __shared__ float2 zred[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, chunk = blockIdx.x;
    const int grp = blockIdx.y / G::KSPLIT, ks = blockIdx.y % G::KSPLIT;
    const bool act = lane < G::ACT;
    const int c0 = grp * G::CPW + 2 * (act ? lane : 0);            // this lane's channel pair
    const int hb = (lane / G::LPH) * G::LPH;                        // first lane of this lane's head
    const int r0 = chunk * rows_per_block, r1 = min(N, r0 + rows_per_block);
    const bf16* base = qkv + (size_t)b * N * 3 * C;
    float2 km = make_float2(0.f, 0.f), gt = make_float2(1.f, 1.f);
    if (MODE == 0) km = *reinterpret_cast<const float2*>(kmax + (size_t)b * C + c0);
    if (MODE == 1 && gate) gt = *reinterpret_cast<const float2*>(gate + (size_t)b * C + c0);
    float2 acc[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) acc[k] = make_float2(0.f, 0.f);
    float2 zacc = make_float2(0.f, 0.f);
    constexpr int U = 4;
    for (int n0 = r0 + warp * U; n0 < r1; n0 += 8 * U) {
        uint32_t pw[U], vw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + u;
            const bool ok = act && n < r1;
            const size_t row = (size_t)n * 3 * C;
            if (MODE == 0) {
                pw[u] = ok ? *reinterpret_cast<const uint32_t*>(base + row + C + c0) : 0u;
                vw[u] = ok ? *reinterpret_cast<const uint32_t*>(base + row + 2 * C + c0) : 0u;
            } else {
                pw[u] = ok ? *reinterpret_cast<const uint32_t*>(base + row + c0) : 0u;
                vw[u] = ok ? *reinterpret_cast<const uint32_t*>(dy + ((size_t)b * N + n) * C + c0) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = act && (n0 + u) < r1;
            float2 p = up2(pw[u]);
            float2 v = up2(vw[u]);
            if (MODE == 0) {
                p = ok ? make_float2(__expf(p.x - km.x), __expf(p.y - km.y)) : make_float2(0.f, 0.f);
                zacc.x += p.x;
                zacc.y += p.y;
            } else {
                v = mul2(v, gt);
            }
#pragma unroll
            for (int kk = 0; kk < KR / 2; ++kk) {
                const float2 pp = shfl2(p, hb + ks * (KR / 2) + kk);
                acc[2 * kk] = fma2(splat(pp.x), v, acc[2 * kk]);
                acc[2 * kk + 1] = fma2(splat(pp.y), v, acc[2 * kk + 1]);
            }
        }
    }
    // deterministic block reduction: warps add in order
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                float2 t = w == 0 ? make_float2(0.f, 0.f) : red[k][lane];
                red[k][lane] = make_float2(t.x + acc[k].x, t.y + acc[k].y);
            }
            if (MODE == 0) {
                float2 t = w == 0 ? make_float2(0.f, 0.f) : zred[lane];
                zred[lane] = make_float2(t.x + zacc.x, t.y + zacc.y);
            }
        }
        __syncthreads();
    }
    // part[b][chunk][c = h*CH + k][v]  (v pair of this lane)
    for (int idx = threadIdx.x; idx < KR * 32; idx += 256) {
        const int k = idx >> 5, l = idx & 31;
        if (l >= G::ACT) continue;
        const int cc = grp * G::CPW + 2 * l;
        const int h = cc / CH, v = cc % CH;
        const int kg = ks * KR + k;
        float* dst = part + (((size_t)(b * nchunk + chunk) * C) + h * CH + kg) * CH + v;
        *reinterpret_cast<float2*>(dst) = red[k][l];
    }
    if (MODE == 0 && ks == 0 && threadIdx.x < G::ACT)
        *reinterpret_cast<float2*>(zpart + (size_t)(b * nchunk + chunk) * C + grp * G::CPW + 2 * threadIdx.x) = zred[threadIdx.x];
}

// A[b,c,v] = sum_chunks part / sum_chunks zpart ; zsum[b,c]
__global__ void attn_combine_fwd_kernel(const float* __restrict__ part, const float* __restrict__ zpart, float* __restrict__ A,
                                        float* __restrict__ zsum, int C, int Ch, int nchunk, long long total) {
    MDV_PDL_SYNC();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bc = i / Ch;
    const int b = (int)(bc / C), c = (int)(bc % C);
    float s = 0.f, z = 0.f;
    for (int k = 0; k < nchunk; ++k) {
        s += part[((size_t)(b * nchunk + k) * C + c) * Ch + (i % Ch)];
        z += zpart[(size_t)(b * nchunk + k) * C + c];
    }
    A[i] = s / z;
    if ((i % Ch) == 0) zsum[bc] = z;
}

// dA[b,c,v] = scale * sum_chunks part ;  rk[b,c] = sum_v A[b,c,v] dA[b,c,v]      (one warp per (b,c) row)
__global__ void attn_combine_bwd_kernel(const float* __restrict__ part, const float* __restrict__ A, float* __restrict__ dA,
                                        float* __restrict__ rk, float scale, int C, int Ch, int nchunk, int rows) {
    MDV_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const int bc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bc >= rows) return;
    const int b = bc / C, c = bc % C;
    float r = 0.f;
    for (int v = lane; v < Ch; v += 32) {
        float s = 0.f;
        for (int k = 0; k < nchunk; ++k) s += part[((size_t)(b * nchunk + k) * C + c) * Ch + v];
        s *= scale;
        dA[(size_t)bc * Ch + v] = s;
        r += s * A[(size_t)bc * Ch + v];
    }
    r = warp_sum(r);
    if (lane == 0) rk[bc] = r;
}

// ------------------------------------------------------------------------------------------------ tile helpers
struct TileGeom {
    int ty0, tx0, th, tw, R, PW, PH, nsx;   // PW: padded row pitch (strips rounded up + halo); nsx strips per row
};
__device__ __forceinline__ TileGeom tile_geom(int H, int Wd, int WIN, int TX = 8) {
    TileGeom g;
    const int TH = min(H, TILE_H), TW = min(Wd, TILE_W);
    const int tiles_x = (Wd + TW - 1) / TW;
    g.ty0 = (blockIdx.x / tiles_x) * TH;
    g.tx0 = (blockIdx.x % tiles_x) * TW;
    g.th = min(TH, H - g.ty0);
    g.tw = min(TW, Wd - g.tx0);
    g.R = WIN >> 1;
    g.nsx = (g.tw + TX - 1) / TX;
    g.PW = g.nsx * TX + 2 * g.R;
    g.PH = g.th + 2 * g.R;
    return g;
}
__host__ __device__ constexpr int tile_words(int WIN) {   // uint32 words of one halo tile (worst case geometry)
    return (TILE_H + WIN - 1) * (((TILE_W + TX - 1) / TX) * TX + WIN - 1) * 32;
}

// stage src[b, pos, c0 + 2*lane .. +1] (bf16x2 words) for all halo positions; zero outside the image / for idle lanes
template <int ACT>
__device__ __forceinline__ void load_halo(uint32_t* dst, const bf16* __restrict__ src_b, int ld, int cg0, const TileGeom& g, int H, int Wd) {
    const int total = g.PH * g.PW * 8;                               // 8 x 16-byte parts per 128-byte position
    constexpr int UB = 8;                                            // loads in flight per thread
    for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * UB) {
        uint4 v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * blockDim.x;
            const int pos = e >> 3, part = e & 7;
            const int py = pos / g.PW, px = pos - py * g.PW;
            const int y = g.ty0 - g.R + py, x = g.tx0 - g.R + px;
            v[u] = make_uint4(0, 0, 0, 0);
            if (e < total && part * 4 < ACT && y >= 0 && y < H && x >= 0 && x < Wd && px < g.tw + 2 * g.R)
                v[u] = *reinterpret_cast<const uint4*>(src_b + (size_t)(y * Wd + x) * ld + cg0 + part * 8);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < total) *reinterpret_cast<uint4*>(dst + (e >> 3) * 32 + (e & 7) * 4) = v[u];
        }
    }
}

// zero-padded WIN x WIN taps of every lane's channel pair: sW[tap][lane] (float2)
template <int CH, int WIN>
__device__ __forceinline__ void load_taps(float2* sW, const CrpeW& cw, int cg0, int nact) {
    for (int e = threadIdx.x; e < WIN * WIN * 32; e += blockDim.x) {
        const int tap = e >> 5, l = e & 31;
        float2 w = make_float2(0.f, 0.f);
        if (l < nact) {
            const int i = tap / WIN, j = tap % WIN;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float* wp; const float* bp; int wc;
                crpe_of_channel(cw, cg0 + 2 * l + u, CH, wp, bp, wc);
                const int o = (WIN - wc) >> 1, ii = i - o, jj = j - o;
                const float t = (ii >= 0 && ii < wc && jj >= 0 && jj < wc) ? __ldg(wp + ii * wc + jj) : 0.f;
                if (u == 0) w.x = t; else w.y = t;
            }
        }
        sW[e] = w;
    }
}

// ------------------------------------------------------------------------------------------------ forward
// Y[n,v] = g[v] * ( s * sum_k Q[n,k] A[k,v] + Q[n,v] * (dwconv(V)[n,v] + b[v]) )         (mdvit.py:293-304)
template <int CH, int WIN>
__device__ __forceinline__ void attn_fwd_strip_body(const bf16* __restrict__ qkv, const float* __restrict__ A,
                                                    const float* __restrict__ gate, const CrpeW& cw, bf16* __restrict__ out,
                                                    bf16* __restrict__ eout, float scale, int H, int Wd, int C, uint8_t* smem_s) {
    using G = Cfg<CH>;
    const int grp0 = 0;
    uint32_t* sV = reinterpret_cast<uint32_t*>(smem_s);
    float2* sW = reinterpret_cast<float2*>(sV + tile_words(WIN));
    float4* sA = reinterpret_cast<float4*>(sW + WIN * WIN * 32);      // [CH/2][32]: (A[2kk][v0], A[2kk+1][v0], A[2kk][v1], A[2kk+1][v1])
    const int N = H * Wd;
    const int b = blockIdx.z, cg0 = (blockIdx.y + grp0) * G::CPW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileGeom g = tile_geom(H, Wd, WIN);
    const bf16* qkv_b = qkv + (size_t)b * N * 3 * C;
    load_halo<G::ACT>(sV, qkv_b + 2 * C, 3 * C, cg0, g, H, Wd);
    load_taps<CH, WIN>(sW, cw, cg0, G::ACT);
    for (int e = threadIdx.x; e < (CH / 2) * 32; e += blockDim.x) {
        const int kk = e >> 5, l = e & 31;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < G::ACT) {
            const int c = cg0 + 2 * l, h = c / CH, v = c % CH;
            const float* src = A + ((size_t)b * C + h * CH + 2 * kk) * CH + v;
            const float2 r0 = *reinterpret_cast<const float2*>(src), r1 = *reinterpret_cast<const float2*>(src + CH);
            a = make_float4(r0.x, r1.x, r0.y, r1.y);
        }
        sA[e] = a;
    }
    const bool act = lane < G::ACT;
    const int c0 = cg0 + 2 * (act ? lane : 0);
    const int hb = (lane / G::LPH) * G::LPH;
    float2 bias = make_float2(0.f, 0.f), gt = make_float2(1.f, 1.f);
    if (act) {
        const float* wp; const float* bp; int wc;
        crpe_of_channel(cw, c0, CH, wp, bp, wc);
        bias.x = __ldg(bp);
        crpe_of_channel(cw, c0 + 1, CH, wp, bp, wc);
        bias.y = __ldg(bp);
        if (gate) gt = *reinterpret_cast<const float2*>(gate + (size_t)b * C + c0);
    }
    __syncthreads();
    const int nstrips = g.th * g.nsx;
    for (int s = warp; s < nstrips; s += 8) {
        const int py = s / g.nsx, px0 = (s % g.nsx) * TX;
        const size_t n0 = (size_t)(g.ty0 + py) * Wd + g.tx0 + px0;
        uint32_t qw[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) qw[t] = (act && px0 + t < g.tw) ? *reinterpret_cast<const uint32_t*>(qkv_b + (n0 + t) * 3 * C + c0) : 0u;
        float2 e[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) e[t] = bias;
#pragma unroll 1
        for (int i = 0; i < WIN; ++i) {
            const uint32_t* row = sV + ((py + i) * g.PW + px0) * 32 + lane;
            float2 in[TX + WIN - 1];
#pragma unroll
            for (int t = 0; t < TX + WIN - 1; ++t) in[t] = up2(row[t * 32]);
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const float2 w = sW[(i * WIN + j) * 32 + lane];
#pragma unroll
                for (int t = 0; t < TX; ++t) e[t] = fma2(w, in[t + j], e[t]);
            }
        }
        // Q.A per head: the (even k, odd k) halves of a packed accumulator are summed at the end, so the shuffled
        // bf16x2 word (Q[2kk], Q[2kk+1]) is used as a packed operand directly
        float2 f0[TX], f1[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) f0[t] = f1[t] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int kk = 0; kk < CH / 2; ++kk) {
            const float4 a = sA[kk * 32 + lane];
            const float2 a0 = make_float2(a.x, a.y), a1 = make_float2(a.z, a.w);
#pragma unroll
            for (int t = 0; t < TX; ++t) {
                const float2 qq = up2(__shfl_sync(0xffffffffu, qw[t], hb + kk));
                f0[t] = fma2(qq, a0, f0[t]);
                f1[t] = fma2(qq, a1, f1[t]);
            }
        }
        float2 fa[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) fa[t] = make_float2(f0[t].x + f0[t].y, f1[t].x + f1[t].y);
        if (act) {
#pragma unroll
            for (int t = 0; t < TX; ++t) {
                if (px0 + t >= g.tw) break;
                const float2 q = up2(qw[t]);
                const float2 y = mul2(gt, fma2(splat(scale), fa[t], mul2(q, e[t])));
                *reinterpret_cast<uint32_t*>(out + ((size_t)b * N + n0 + t) * C + c0) = f2_to_bf2(y.x, y.y);
                // E = dwconv(V) + b is kept for the backward pass (dQ needs it; recomputing it there costs a third of that kernel)
                if (eout) *reinterpret_cast<uint32_t*>(eout + ((size_t)b * N + n0 + t) * C + c0) = f2_to_bf2(e[t].x, e[t].y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// dQ = s dF.A^T + dF*E ; dK = S*(V.dA^T - r) ; dV = S.dA + convT(dE) ; dE = g dY Q ; dF = g dY            
// dgate[b,c] += sum_n dY Y / g ;  dWconv[c,tap] += sum_n dE[n,c] V[n+tap,c] ;  dbconv[c] += sum_n dE[n,c]
// EXT: the three per-head mat-vecs (and with them dQ and dK) are done by attn_mm_bwd_kernel on the tensor cores; this
// kernel then only does the convolutional part: dV = dv_part + conv^T(dE), the CRPE weight gradients and the gate sums.
//
// Shared memory: dE at every halo position as fp32 pairs (no unpacking in the 7x7 loops), V as bf16 pairs (weight-gradient
// pass only), the FLIPPED taps, the three per-head matrices.  The halo tile has a fixed pitch of TILE_W + WIN - 1 positions,
// so all position arithmetic is by compile-time constants; FULL = the block's tile is a whole 16 x 16 one (no bounds tests).
// Pass B (weight gradients) gives every warp ONE kernel row i and a band of tile rows: 7 live accumulators, no division,
// and the per-warp results are parked over the dE tile and summed in a fixed order (no shared-memory atomics).
constexpr int BWD_THREADS = 512;   // 16 warps share one tile (register bound: one block per SM)
constexpr int BWD_TX = 4;          // activation-gradient strips: 4 pixels keep the per-thread arrays within 128 registers
constexpr int BWD_TXB = 8;         // weight-gradient strips
#ifndef MDV_MVU
#define MDV_MVU 2
#endif
#ifndef MDV_MVJ
#define MDV_MVJ 2
#endif
constexpr int MVU = MDV_MVU;       // pixels per group of the in-kernel mat-vecs (Ch <= 16)
constexpr int MVJ = MDV_MVJ;       // unroll of their j loop

template <int WIN>
__host__ __device__ constexpr int bwd_tile_pos() { return (TILE_H + WIN - 1) * (TILE_W + WIN - 1); }

template <int CH, int WIN, int TX, bool EXT, bool FULL>
__device__ __forceinline__ void attn_bwd_strip_body(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                    const bf16* __restrict__ yout, const float* __restrict__ gate,
                                                    const float* __restrict__ A, const float* __restrict__ dA,
                                                    const float* __restrict__ rk, const float* __restrict__ kmax,
                                                    const float* __restrict__ zsum, const CrpeW& cw, const CrpeG& cg,
                                                    const bf16* __restrict__ ein, bf16* __restrict__ dqkv,
                                                    float* __restrict__ dgate, float* __restrict__ dbias_qkv, float scale, int H,
                                                    int Wd, int C, uint8_t* smem_s) {
    using G = Cfg<CH>;
    constexpr int NT = BWD_THREADS, NW = NT / 32;
    constexpr int R = WIN >> 1, PW = TILE_W + WIN - 1, NPOS = bwd_tile_pos<WIN>(), NTAP = WIN * WIN;
    const bool want_w = cg.w[0] != nullptr;
    float2* sE = reinterpret_cast<float2*>(smem_s);                  // [NPOS][32] dE, fp32 pairs
    uint32_t* sV = reinterpret_cast<uint32_t*>(sE + NPOS * 32);      // [NPOS][32] V, bf16 pairs (weight-gradient pass only)
    float2* sW = reinterpret_cast<float2*>(sV + (want_w ? NPOS * 32 : 0));   // [NTAP][32] flipped taps
    // per (j pair, lane): the two j-values of a matrix entry for each of the lane's two channels v0, v1
    float4* sAt = reinterpret_cast<float4*>(sW + NTAP * 32);         // (A[v0][2jj],  A[v0][2jj+1],  A[v1][2jj],  A[v1][2jj+1])
    float4* sdA = sAt + (CH / 2) * 32;                               // (dA[2jj][v0], dA[2jj+1][v0], dA[2jj][v1], dA[2jj+1][v1])
    float4* sdAt = sdA + (CH / 2) * 32;                              // (dA[v0][2jj], dA[v0][2jj+1], dA[v1][2jj], dA[v1][2jj+1])
    const int N = H * Wd;
    const int b = blockIdx.z, cg0 = blockIdx.y * G::CPW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ty0, tx0, th, tw;
    {
        const int TH = min(H, TILE_H), TW = min(Wd, TILE_W);
        const int tiles_x = (Wd + TW - 1) / TW;
        ty0 = (blockIdx.x / tiles_x) * TH;
        tx0 = (blockIdx.x % tiles_x) * TW;
        th = FULL ? TILE_H : min(TH, H - ty0);
        tw = FULL ? TILE_W : min(TW, Wd - tx0);
    }
    const bf16* qkv_b = qkv + (size_t)b * N * 3 * C;
    const bf16* dy_b = dy + (size_t)b * N * C;
    const int npos = (th + 2 * R) * PW;
    {   // staging: a thread always handles the same 8-channel part of a position
        const int part = threadIdx.x & 7;
        const bool pact = part * 4 < G::ACT;
        const int cc = cg0 + part * 8;
        float2 gg[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
            gg[w] = (gate && pact) ? *reinterpret_cast<const float2*>(gate + (size_t)b * C + cc + 2 * w) : make_float2(1.f, 1.f);
        constexpr int UB = 4, PSTEP = NT / 8;
        // dE = g * dY * Q at every halo position (2 x UB 16-byte requests in flight per thread)
        for (int p0 = threadIdx.x >> 3; p0 < npos; p0 += PSTEP * UB) {
            uint4 dv[UB], qv[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int pos = p0 + u * PSTEP;
                const int py = pos / PW, px = pos - py * PW;
                const int y = ty0 - R + py, x = tx0 - R + px;
                dv[u] = qv[u] = make_uint4(0, 0, 0, 0);
                if (pos < npos && pact && y >= 0 && y < H && x >= 0 && x < Wd && px < tw + 2 * R) {
                    const size_t n = (size_t)y * Wd + x;
                    dv[u] = *reinterpret_cast<const uint4*>(dy_b + n * C + cc);
                    qv[u] = *reinterpret_cast<const uint4*>(qkv_b + n * 3 * C + cc);
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int pos = p0 + u * PSTEP;
                if (pos >= npos) continue;
                const float2 o0 = mul2(mul2(gg[0], up2(dv[u].x)), up2(qv[u].x)), o1 = mul2(mul2(gg[1], up2(dv[u].y)), up2(qv[u].y));
                const float2 o2 = mul2(mul2(gg[2], up2(dv[u].z)), up2(qv[u].z)), o3 = mul2(mul2(gg[3], up2(dv[u].w)), up2(qv[u].w));
                float4* dst = reinterpret_cast<float4*>(sE + pos * 32 + part * 4);
                dst[0] = make_float4(o0.x, o0.y, o1.x, o1.y);
                dst[1] = make_float4(o2.x, o2.y, o3.x, o3.y);
            }
        }
        if (want_w) {
            constexpr int UV = 8;
            for (int p0 = threadIdx.x >> 3; p0 < npos; p0 += PSTEP * UV) {
                uint4 v[UV];
#pragma unroll
                for (int u = 0; u < UV; ++u) {
                    const int pos = p0 + u * PSTEP;
                    const int py = pos / PW, px = pos - py * PW;
                    const int y = ty0 - R + py, x = tx0 - R + px;
                    v[u] = make_uint4(0, 0, 0, 0);
                    if (pos < npos && pact && y >= 0 && y < H && x >= 0 && x < Wd && px < tw + 2 * R)
                        v[u] = *reinterpret_cast<const uint4*>(qkv_b + ((size_t)y * Wd + x) * 3 * C + 2 * C + cc);
                }
#pragma unroll
                for (int u = 0; u < UV; ++u) {
                    const int pos = p0 + u * PSTEP;
                    if (pos < npos) *reinterpret_cast<uint4*>(sV + pos * 32 + part * 4) = v[u];
                }
            }
        }
    }
    const bool act = lane < G::ACT;
    const int c0 = cg0 + 2 * (act ? lane : 0);
    {   // flipped, zero-padded WIN x WIN taps of every lane's channel pair: sW[tap][lane] = w[NTAP - 1 - tap]
        const float* wp0; const float* wp1; const float* bp; int wc0, wc1;
        crpe_of_channel(cw, c0, CH, wp0, bp, wc0);
        crpe_of_channel(cw, c0 + 1, CH, wp1, bp, wc1);
        for (int tap = warp; tap < NTAP; tap += NW) {
            const int ft = NTAP - 1 - tap, i = ft / WIN, j = ft - i * WIN;
            float2 w = make_float2(0.f, 0.f);
            if (act) {
                const int o0 = (WIN - wc0) >> 1, o1 = (WIN - wc1) >> 1;
                const int i0 = i - o0, j0 = j - o0, i1 = i - o1, j1 = j - o1;
                if (i0 >= 0 && i0 < wc0 && j0 >= 0 && j0 < wc0) w.x = __ldg(wp0 + i0 * wc0 + j0);
                if (i1 >= 0 && i1 < wc1 && j1 >= 0 && j1 < wc1) w.y = __ldg(wp1 + i1 * wc1 + j1);
            }
            sW[tap * 32 + lane] = w;
        }
    }
    for (int e = threadIdx.x; e < (EXT ? 0 : (CH / 2) * 32); e += NT) {
        const int jj = e >> 5, l = e & 31;
        float4 at = make_float4(0.f, 0.f, 0.f, 0.f), da = at, dat = at;
        if (l < G::ACT) {
            const int c = cg0 + 2 * l, h = c / CH, v = c % CH;
            const size_t hbase = ((size_t)b * C + h * CH) * CH;
            const float2 a0 = *reinterpret_cast<const float2*>(A + hbase + (size_t)v * CH + 2 * jj);
            const float2 a1 = *reinterpret_cast<const float2*>(A + hbase + (size_t)(v + 1) * CH + 2 * jj);
            at = make_float4(a0.x, a0.y, a1.x, a1.y);
            const float2 d0 = *reinterpret_cast<const float2*>(dA + hbase + (size_t)(2 * jj) * CH + v);
            const float2 d1 = *reinterpret_cast<const float2*>(dA + hbase + (size_t)(2 * jj + 1) * CH + v);
            da = make_float4(d0.x, d1.x, d0.y, d1.y);
            const float2 t0 = *reinterpret_cast<const float2*>(dA + hbase + (size_t)v * CH + 2 * jj);
            const float2 t1 = *reinterpret_cast<const float2*>(dA + hbase + (size_t)(v + 1) * CH + 2 * jj);
            dat = make_float4(t0.x, t0.y, t1.x, t1.y);
        }
        sAt[e] = at;
        sdA[e] = da;
        sdAt[e] = dat;
    }
    const int hb = (lane / G::LPH) * G::LPH;
    float2 gt = make_float2(1.f, 1.f), km = make_float2(0.f, 0.f), zi = make_float2(0.f, 0.f), rr = make_float2(0.f, 0.f);
    if (act) {
        const size_t bc = (size_t)b * C + c0;
        if (gate) gt = *reinterpret_cast<const float2*>(gate + bc);
        if (!EXT) {
            km = *reinterpret_cast<const float2*>(kmax + bc);
            const float2 z = *reinterpret_cast<const float2*>(zsum + bc);
            zi = make_float2(1.f / z.x, 1.f / z.y);
            rr = *reinterpret_cast<const float2*>(rk + bc);
        }
    }
    __syncthreads();
    const int nsx = FULL ? TILE_W / TX : (tw + TX - 1) / TX;
    const int nstrips = th * nsx;
    float2 gacc = make_float2(0.f, 0.f);
    float2 bq = make_float2(0.f, 0.f), bk = bq, bv = bq;      // column sums of dQ, dK, dV: the qkv Linear's bias gradient
    // ---- pass A: activation gradients
    for (int s = warp; s < nstrips; s += NW) {
        const int py = s / nsx, px0 = (s - py * nsx) * TX;
        const size_t n0 = (size_t)(ty0 + py) * Wd + tx0 + px0;
        // this strip's own pixels: issued before the convolution loop so their latency hides behind the math
        uint32_t kw[TX], dw[TX], yw[TX], ew[TX], vw[TX];
        float2 dvp[TX];
        const bf16* qrow = qkv_b + n0 * 3 * C + c0;
        const bf16* drow_in = dy_b + n0 * C + c0;
        const size_t nb = ((size_t)b * N + n0) * C + c0;
#pragma unroll
        for (int t = 0; t < TX; ++t) {
            const bool ok = FULL ? act : (act && px0 + t < tw);
            kw[t] = (ok && !EXT) ? *reinterpret_cast<const uint32_t*>(qrow + t * 3 * C + C) : 0u;
            vw[t] = (ok && !EXT) ? *reinterpret_cast<const uint32_t*>(qrow + t * 3 * C + 2 * C) : 0u;
            dw[t] = (ok && (!EXT || gate)) ? *reinterpret_cast<const uint32_t*>(drow_in + t * C) : 0u;
            ew[t] = (ok && !EXT) ? *reinterpret_cast<const uint32_t*>(ein + nb + t * C) : 0u;
            yw[t] = (ok && gate) ? *reinterpret_cast<const uint32_t*>(yout + nb + t * C) : 0u;
            // (the mat-vec part of dV was parked, as bf16, in the dV slot of dqkv by attn_mm_bwd_kernel)
            dvp[t] = (ok && EXT) ? up2(*reinterpret_cast<const uint32_t*>(dqkv + 3 * nb - 2 * c0 + 2 * C + t * 3 * C)) : make_float2(0.f, 0.f);
        }
        // transposed convolution of dE (flipped taps): dV_conv[n] = sum_ij w[i,j] dE[n - (i-R, j-R)]
        float2 tc[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) tc[t] = make_float2(0.f, 0.f);
        {
            const float2* rowe = sE + (py * PW + px0) * 32 + lane;
            const float2* wrow = sW + lane;
#pragma unroll 1
            for (int i = 0; i < WIN; ++i, rowe += PW * 32, wrow += WIN * 32) {
                float2 ine[TX + WIN - 1];
#pragma unroll
                for (int t = 0; t < TX + WIN - 1; ++t) ine[t] = rowe[t * 32];
#pragma unroll
                for (int j = 0; j < WIN; ++j) {
                    const float2 wf = wrow[j * 32];
#pragma unroll
                    for (int t = 0; t < TX; ++t) tc[t] = fma2(wf, ine[t + j], tc[t]);
                }
            }
        }
        bf16* drow = dqkv + 3 * nb - 2 * c0;       // = dqkv + ((b N + n0) 3C + c0)
        if constexpr (EXT) {
            // dQ, dK and the mat-vec part of dV come from attn_mm_bwd_kernel: finish dV, the gate sums and the bias sums
            if (act) {
#pragma unroll
                for (int t = 0; t < TX; ++t) {
                    if (!FULL && px0 + t >= tw) break;
                    gacc = fma2(up2(dw[t]), up2(yw[t]), gacc);
                    const float2 dv = make_float2(dvp[t].x + tc[t].x, dvp[t].y + tc[t].y);
                    bv.x += dv.x; bv.y += dv.y;
                    *reinterpret_cast<uint32_t*>(drow + t * 3 * C + 2 * C) = f2_to_bf2(dv.x, dv.y);
                }
            }
        } else {
        float2 dF[TX], S[TX];
#pragma unroll
        for (int t = 0; t < TX; ++t) {
            const bool ok = FULL ? act : (act && px0 + t < tw);
            const float2 kk = up2(kw[t]), d = up2(dw[t]);
            dF[t] = mul2(gt, d);
            S[t] = ok ? make_float2(__expf(kk.x - km.x) * zi.x, __expf(kk.y - km.y) * zi.y) : make_float2(0.f, 0.f);
            gacc = fma2(d, up2(yw[t]), gacc);
        }
        // three per-head mat-vecs, MVU pixels at a time (register pressure); packed accumulators hold the (even j, odd j)
        // partial sums, so the shuffled pairs are used as packed operands without splatting
#pragma unroll
        for (int half = 0; half < TX / MVU; ++half) {
            float2 q0[MVU], q1[MVU], v0[MVU], v1[MVU], k0[MVU], k1[MVU];
#pragma unroll
            for (int u = 0; u < MVU; ++u) q0[u] = q1[u] = v0[u] = v1[u] = k0[u] = k1[u] = make_float2(0.f, 0.f);
#pragma unroll (MVJ)
            for (int jj = 0; jj < CH / 2; ++jj) {
                const float4 at = sAt[jj * 32 + lane], da = sdA[jj * 32 + lane], dt = sdAt[jj * 32 + lane];
#pragma unroll
                for (int u = 0; u < MVU; ++u) {
                    const int t = half * MVU + u;
                    const float2 f = shfl2(dF[t], hb + jj);
                    const float2 ss = shfl2(S[t], hb + jj);
                    const float2 vv = up2(__shfl_sync(0xffffffffu, vw[t], hb + jj));
                    q0[u] = fma2(f, make_float2(at.x, at.y), q0[u]);
                    q1[u] = fma2(f, make_float2(at.z, at.w), q1[u]);
                    v0[u] = fma2(ss, make_float2(da.x, da.y), v0[u]);
                    v1[u] = fma2(ss, make_float2(da.z, da.w), v1[u]);
                    k0[u] = fma2(vv, make_float2(dt.x, dt.y), k0[u]);
                    k1[u] = fma2(vv, make_float2(dt.z, dt.w), k1[u]);
                }
            }
            if (act) {
#pragma unroll
                for (int u = 0; u < MVU; ++u) {
                    const int t = half * MVU + u;
                    if (!FULL && px0 + t >= tw) break;
                    const float2 sq = make_float2(q0[u].x + q0[u].y, q1[u].x + q1[u].y);
                    const float2 sv = make_float2(v0[u].x + v0[u].y, v1[u].x + v1[u].y);
                    const float2 sk = make_float2(k0[u].x + k0[u].y, k1[u].x + k1[u].y);
                    const float2 dq = fma2(splat(scale), sq, mul2(dF[t], up2(ew[t])));
                    const float2 dk = mul2(S[t], make_float2(sk.x - rr.x, sk.y - rr.y));
                    const float2 dv = make_float2(sv.x + tc[t].x, sv.y + tc[t].y);
                    bq.x += dq.x; bq.y += dq.y; bk.x += dk.x; bk.y += dk.y; bv.x += dv.x; bv.y += dv.y;
                    *reinterpret_cast<uint32_t*>(drow + t * 3 * C) = f2_to_bf2(dq.x, dq.y);
                    *reinterpret_cast<uint32_t*>(drow + t * 3 * C + C) = f2_to_bf2(dk.x, dk.y);
                    *reinterpret_cast<uint32_t*>(drow + t * 3 * C + 2 * C) = f2_to_bf2(dv.x, dv.y);
                }
            }
        }
        }
    }
    // ---- pass B: convolution weight / bias gradients.  warp -> (kernel row i, band of tile rows); one spare warp sums the bias
    constexpr int NPARTS = NW / WIN, NSLOT = NTAP + 1;
    float2 gw[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) gw[j] = make_float2(0.f, 0.f);
    float2 gb = make_float2(0.f, 0.f);
    const int item_i = warp / NPARTS, item_p = warp - item_i * NPARTS;
    if (want_w) {
        if (item_i < WIN) {
            const int rows = (th + NPARTS - 1) / NPARTS;
            const int r0 = item_p * rows, r1 = min(th, r0 + rows);
            const int nsxb = FULL ? TILE_W / BWD_TXB : (tw + BWD_TXB - 1) / BWD_TXB;
            for (int py = r0; py < r1; ++py) {
                for (int sx = 0; sx < nsxb; ++sx) {
                    const int px0 = sx * BWD_TXB;
                    const uint32_t* rowv = sV + ((py + item_i) * PW + px0) * 32 + lane;
                    const float2* ce = sE + ((py + R) * PW + px0 + R) * 32 + lane;
                    float2 inv[BWD_TXB + WIN - 1], de[BWD_TXB];
#pragma unroll
                    for (int t = 0; t < BWD_TXB + WIN - 1; ++t) inv[t] = up2(rowv[t * 32]);
#pragma unroll
                    for (int t = 0; t < BWD_TXB; ++t)      // (halo columns past the tile hold the neighbours' dE)
                        de[t] = (FULL || px0 + t < tw) ? ce[t * 32] : make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < WIN; ++j)
#pragma unroll
                        for (int t = 0; t < BWD_TXB; ++t) gw[j] = fma2(de[t], inv[t + j], gw[j]);
                }
            }
        } else if (warp == WIN * NPARTS) {
            for (int py = 0; py < th; ++py) {
                const float2* ce = sE + ((py + R) * PW + R) * 32 + lane;
                for (int px = 0; px < tw; ++px) {
                    const float2 d = ce[px * 32];
                    gb.x += d.x; gb.y += d.y;
                }
            }
        }
    }
    __syncthreads();     // every warp is done with the dE tile: the per-warp sums are parked over it
    float2* sR = reinterpret_cast<float2*>(smem_s);                   // [NW][4][32]: dQ, dK, dV column sums and the gate sum
    float* sP = reinterpret_cast<float*>(sR + NW * 4 * 32);           // [NPARTS][NSLOT][SPP] weight-gradient partials
    constexpr int SPP = 66;                                           // (row pitch: the final sum reads columns)
    sR[(warp * 4 + 0) * 32 + lane] = bq;
    sR[(warp * 4 + 1) * 32 + lane] = bk;
    sR[(warp * 4 + 2) * 32 + lane] = bv;
    sR[(warp * 4 + 3) * 32 + lane] = gacc;
    if (want_w) {
        if (item_i < WIN) {
#pragma unroll
            for (int j = 0; j < WIN; ++j)
                *reinterpret_cast<float2*>(sP + (item_p * NSLOT + item_i * WIN + j) * SPP + 2 * lane) = gw[j];
        } else if (warp == WIN * NPARTS) {
            *reinterpret_cast<float2*>(sP + NTAP * SPP + 2 * lane) = gb;
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int q = threadIdx.x >> 5;          // 0: dQ bias, 1: dK bias, 2: dV bias, 3: gate
        float2 sum = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float2 v = sR[(w * 4 + q) * 32 + lane];
            sum.x += v.x; sum.y += v.y;
        }
        if (act) {
            if (q == 3) {
                if (gate && dgate) {
                    atomicAdd(dgate + (size_t)b * C + c0, sum.x / gt.x);
                    atomicAdd(dgate + (size_t)b * C + c0 + 1, sum.y / gt.y);
                }
            } else if (dbias_qkv && (!EXT || q == 2)) {
                atomicAdd(dbias_qkv + q * C + c0, sum.x);
                atomicAdd(dbias_qkv + q * C + c0 + 1, sum.y);
            }
        }
    }
    if (!want_w) return;   // block-uniform
    // lanes walk the taps of one channel: consecutive addresses of the gradient array
    for (int o = threadIdx.x; o < NSLOT * G::CPW; o += NT) {
#ifndef MDV_EXP_NO_TAIL
        const int cl_ = o / NSLOT, t = o - cl_ * NSLOT;
        float sum = sP[t * SPP + cl_];
        if (t < NTAP)
#pragma unroll
            for (int p = 1; p < NPARTS; ++p) sum += sP[(p * NSLOT + t) * SPP + cl_];
        const int c = cg0 + cl_;
        const int h = c / CH;
        const int grp = h < 2 ? 0 : (h < 5 ? 1 : 2);
        const int cl = c - (grp == 0 ? 0 : (grp == 1 ? 2 * CH : 5 * CH));
        const int wc = 3 + 2 * grp;
        float* gwp = grp == 0 ? cg.w[0] : (grp == 1 ? cg.w[1] : cg.w[2]);
        float* gbp = grp == 0 ? cg.b[0] : (grp == 1 ? cg.b[1] : cg.b[2]);
        if (t < NTAP) {
            const int o2 = (WIN - wc) >> 1;
            const int ti = t / WIN;
            const int ii = ti - o2, jj = t - ti * WIN - o2;
            if (ii >= 0 && ii < wc && jj >= 0 && jj < wc) atomicAdd(gwp + (size_t)cl * wc * wc + ii * wc + jj, sum);
        } else {
            atomicAdd(gbp + cl, sum);
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------------ backward mat-vecs on tensor cores
// For Ch >= 40 (stages 2 and 3) the three per-head contractions of the backward,
//     T1 = dF A^T  (-> dQ = s T1 + dF*E),   T2 = V dA^T  (-> dK = S*(T2 - r)),   T3 = S dA  (-> mat-vec part of dV),
// cost 3*Ch^2 multiply-adds per token and head: as shuffled FFMA2 mat-vecs they were 60-65% of attn_bwd_strip_kernel<40/64>.
// Here a warp takes 16 tokens of one head and runs them as bf16 mma.sync m16n8k16 (fp32 accumulate): the token rows are the
// A fragments (loaded straight from global memory in fragment layout), the Ch x Ch matrices (A, dA, dA^T as bf16 in shared
// memory, row pitch padded against bank conflicts) the B fragments.  S = softmax_N(K) is computed once in A-fragment layout,
// which for m16n8k16 coincides with the accumulator layout of n-tiles (2t, 2t+1): the same registers serve as the A operand
// of T3 and as the elementwise factor of dK.  dQ and dK are final here; the mat-vec part of dV is parked (bf16) in the dV
// slot of dqkv for attn_bwd_strip_kernel<CH, EXT> to add the transposed convolution.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------ cross-token sums on tensor cores
// Ch >= 40 (stages 2 / 3): part[k,v] = sum_n P[n,k] V'[n,v] is a (Ch x tokens).(tokens x Ch) product per head — as shuffled FFMA2
// outer products it was issue bound at 14x its HBM time.  Here: mma.sync m16n8k16 bf16, fp32 accumulation.  A tile of 64 tokens
// of both operands is staged in shared memory as plain [token][channel] bf16 rows (pitch padded by 16 B: conflict-free
// ldmatrix); ldmatrix.trans turns the token-major rows into the K-major fragments.  One warp per 8-column n-tile, all m-tiles.
//   MODE 0: P = exp(K - kmax) is split into bf16 hi + lo parts (two MMAs per fragment): the sums keep fp32-level accuracy;
//           zpart is accumulated from the fp32 values during staging.
//   MODE 1: P = Q, V' = dY; the gate g[v] is a column factor and is applied to the finished sums.
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}

// A CTA covers one channel group: a whole head for Ch >= 40, 64 channels (8 or 4 heads) for Ch = 8 / 16 — there an n-tile only
// needs the m-tile that holds its own head (for Ch = 8 half of that 16-row tile belongs to the neighbouring head and is dropped).
template <int CH, int MODE>
__global__ void __launch_bounds__(32 * ((CH <= 16 ? 64 : CH) / 8)) attn_outer_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                                       const float* __restrict__ gate, const float* __restrict__ kmax,
                                                                       float* __restrict__ part, float* __restrict__ zpart, int N, int C,
                                                                       int rows_per_block, int nchunk) {
    MDV_PDL_SYNC();
    constexpr int CPB = CH <= 16 ? 64 : CH;    // channels per CTA
    constexpr int NP = CPB / 8;                // 16-byte parts per token row = warps = n-tiles
    constexpr int NT = 32 * NP;
    constexpr int MTALL = (CPB + 15) / 16;     // m-tiles of the staged tile (Ch = 40: rows 40..47 are zero padding)
    constexpr int MT = CH <= 16 ? 1 : MTALL;   // m-tiles a warp multiplies
    constexpr int PITCH = MTALL * 32 + 16;     // bytes per staged token row
    constexpr int TT = 64;                     // tokens per tile
    __shared__ __align__(16) uint8_t sPh[TT * PITCH];
    __shared__ __align__(16) uint8_t sPl[MODE == 0 ? TT * PITCH : 16];
    __shared__ __align__(16) uint8_t sV[TT * PITCH];
    __shared__ float sZ[MODE == 0 ? 32 * CPB : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, cg0 = blockIdx.y * CPB, chunk = blockIdx.x;
    const int r0 = chunk * rows_per_block, r1 = min(N, r0 + rows_per_block);
    const int prt = threadIdx.x % NP, prow = threadIdx.x / NP;       // staging role: 8 channels of token (tile row prow, prow + 32)
    const int cb = cg0 + prt * 8;                                     // first channel of this thread's part
    const int m_base = CH <= 16 ? ((warp * 8) / 16) * 16 : 0;         // first row (group channel) of this warp's m-tiles
    const bf16* base = qkv + (size_t)b * N * 3 * C;
    // zero the padding columns once (never written again)
    for (int i = threadIdx.x; i < TT * PITCH / 16; i += NT) {
        reinterpret_cast<uint4*>(sPh)[i] = make_uint4(0, 0, 0, 0);
        if (MODE == 0) reinterpret_cast<uint4*>(sPl)[i] = make_uint4(0, 0, 0, 0);
    }
    float km[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        km[j] = MODE == 0 ? kmax[(size_t)b * C + cb + j] : 0.f;
        z[j] = 0.f;
    }
    float acc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
    const uint32_t sph = (uint32_t)__cvta_generic_to_shared(sPh), spl = (uint32_t)__cvta_generic_to_shared(sPl),
                   sv = (uint32_t)__cvta_generic_to_shared(sV);
    // ldmatrix row addresses of this lane: A (x4.trans): token = k0 + (lane & 7) + ((lane >> 4) << 3), column = m0 + ((lane >> 3) & 1) * 8
    const uint32_t a_off = (uint32_t)(((lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 16 + m_base * 2);
    // B (x2.trans): token = k0 + (lane & 7) + ((lane >> 3) & 1) * 8, column = 8 * warp
    const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + warp * 16);
    __syncthreads();
    for (int t0 = r0; t0 < r1; t0 += TT) {
        // ---- stage the tile (two token rows per thread)
        uint4 pv[2], vv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int n = t0 + prow + 32 * u;
            pv[u] = vv[u] = make_uint4(0, 0, 0, 0);
            if (n < r1) {
                if (MODE == 0) {
                    pv[u] = *reinterpret_cast<const uint4*>(base + (size_t)n * 3 * C + C + cb);
                    vv[u] = *reinterpret_cast<const uint4*>(base + (size_t)n * 3 * C + 2 * C + cb);
                } else {
                    pv[u] = *reinterpret_cast<const uint4*>(base + (size_t)n * 3 * C + cb);
                    vv[u] = *reinterpret_cast<const uint4*>(dy + ((size_t)b * N + n) * C + cb);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int n = t0 + prow + 32 * u;
            const uint32_t so = (uint32_t)((prow + 32 * u) * PITCH + prt * 16);
            if (MODE == 0) {
                const uint32_t w4[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 kk = bf2_to_f2(w4[w]);
                    float2 p = make_float2(0.f, 0.f);
                    if (n < r1) p = make_float2(__expf(kk.x - km[2 * w]), __expf(kk.y - km[2 * w + 1]));
                    z[2 * w] += p.x;
                    z[2 * w + 1] += p.y;
                    hi[w] = f2_to_bf2(p.x, p.y);
                    const float2 hf = bf2_to_f2(hi[w]);
                    lo[w] = f2_to_bf2(p.x - hf.x, p.y - hf.y);
                }
                *reinterpret_cast<uint4*>(sPh + so) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(sPl + so) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {
                *reinterpret_cast<uint4*>(sPh + so) = pv[u];
            }
            *reinterpret_cast<uint4*>(sV + so) = vv[u];
        }
        __syncthreads();
        // ---- 4 k-steps of 16 tokens
#pragma unroll
        for (int ks = 0; ks < TT / 16; ++ks) {
            uint32_t b0, b1;
            ldsm_x2_t(b0, b1, sv + ks * 16 * PITCH + b_off);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                uint32_t a[4];
                ldsm_x4_t(a, sph + ks * 16 * PITCH + m * 32 + a_off);
                mma_bf16_16816(acc[m], a, b0, b1);
                if (MODE == 0) {
                    ldsm_x4_t(a, spl + ks * 16 * PITCH + m * 32 + a_off);
                    mma_bf16_16816(acc[m], a, b0, b1);
                }
            }
        }
        __syncthreads();
    }
    // ---- results: acc[m] = rows (group channel) m_base + 16 m + g (+8), columns (group channel) 8 warp + 2 t (+1);
    //      part[b][chunk][c = head * CH + k][v]: only the rows of the column's own head are kept
    const int g = lane >> 2, t = lane & 3;
    const int vg = warp * 8 + 2 * t;                  // column within the group
    const int hv = vg / CH;                           // its head within the group
    float2 gt = make_float2(1.f, 1.f);
    if (MODE == 1 && gate) gt = *reinterpret_cast<const float2*>(gate + (size_t)b * C + cg0 + vg);
    float* pbase = part + ((size_t)(b * nchunk + chunk) * C + cg0) * CH;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int kg = m_base + 16 * m + g + 8 * half;      // row within the group
            if (kg < CPB && kg / CH == hv)
                *reinterpret_cast<float2*>(pbase + (size_t)kg * CH + (vg - hv * CH)) = make_float2(acc[m][2 * half] * gt.x, acc[m][2 * half + 1] * gt.y);
        }
    }
    if (MODE == 0) {
        // zpart[c] = sum over this chunk's tokens of exp(K - kmax): every thread summed 8 channels over its rows
#pragma unroll
        for (int j = 0; j < 8; ++j) sZ[prow * CPB + prt * 8 + j] = z[j];
        __syncthreads();
        for (int c = threadIdx.x; c < CPB; c += NT) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) s += sZ[r * CPB + c];
            zpart[(size_t)(b * nchunk + chunk) * C + cg0 + c] = s;
        }
    }
}


__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

// Token tiles of 64 are staged through shared memory with coalesced 16-byte loads (g dY, S = softmax_N(K) in bf16 and fp32, V, E as
// [token][channel] rows), the A fragments come from ldmatrix, and the results (dQ, dK, the mat-vec part of dV) overwrite the
// staged rows of their own warp before one coalesced store — the first version loaded every fragment word straight from global
// memory (4-byte accesses, half-used sectors, ~70 load instructions per 16 tokens) and was latency bound at 4-5x its HBM time.
template <int CH>
__global__ void __launch_bounds__(128) attn_mm_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                          const bf16* __restrict__ ein, const float* __restrict__ gate,
                                                          const float* __restrict__ A, const float* __restrict__ dA,
                                                          const float* __restrict__ rk, const float* __restrict__ kmax,
                                                          const float* __restrict__ zsum, bf16* __restrict__ dqkv,
                                                          float* __restrict__ dbias_qkv, float scale, int N, int C, int tok_per_cta) {
    MDV_PDL_SYNC();
    constexpr int KS = (CH + 15) / 16;      // k-steps of 16
    constexpr int NT = CH / 8;              // n-tiles of 8 = 16-byte parts of a token row
    constexpr int LD = KS * 16 + 8;         // matrix row pitch (bf16): +8 keeps the B-fragment loads conflict-free
    constexpr int PITCH = KS * 32 + 16;     // bytes per staged token row (padded channels + 16 B: conflict-free ldmatrix)
    constexpr int P32 = CH + 2;             // floats per row of the fp32 S tile
    constexpr int TT = 64;                  // tokens per tile: 16 per warp
    extern __shared__ __align__(16) uint8_t smem_mm[];
    bf16* sA = reinterpret_cast<bf16*>(smem_mm);
    bf16* sdA = sA + CH * LD;
    bf16* sdAt = sdA + CH * LD;
    uint8_t* tF = reinterpret_cast<uint8_t*>(sdAt + CH * LD);      // g * dY      -> dQ
    uint8_t* tV = tF + TT * PITCH;                                  // V           -> dK
    uint8_t* tS = tV + TT * PITCH;                                  // S (bf16)    -> mat-vec part of dV
    uint8_t* tE = tS + TT * PITCH;                                  // E = dwconv(V) + b
    float* tS32 = reinterpret_cast<float*>(tE + TT * PITCH);        // S (fp32): the elementwise factor of dK
    float* sbias = tS32 + TT * P32;                                 // [2][CH]
    float* srk = sbias + 2 * CH;                                    // [CH]
    float* sG = srk + CH;                                           // gate, kmax, 1 / zsum of this head's channels
    float* sKm = sG + CH;
    float* sZi = sKm + CH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int h = blockIdx.y, b = blockIdx.z;
    const size_t hbase = ((size_t)b * C + h * CH) * CH;
    const int cb = b * C + h * CH;                        // index of this head's first channel in [B, C] tables
    for (int e = threadIdx.x; e < CH * LD; e += blockDim.x) {
        const int r = e / LD, c = e % LD;
        const bool in = c < CH;
        sA[e] = __float2bfloat16_rn(in ? A[hbase + (size_t)r * CH + c] : 0.f);
        sdA[e] = __float2bfloat16_rn(in ? dA[hbase + (size_t)r * CH + c] : 0.f);
        sdAt[e] = __float2bfloat16_rn(in ? dA[hbase + (size_t)c * CH + r] : 0.f);
    }
    for (int e = threadIdx.x; e < CH; e += blockDim.x) {
        sbias[e] = sbias[CH + e] = 0.f;
        sG[e] = gate ? gate[cb + e] : 1.f;
        sKm[e] = kmax[cb + e];
        sZi[e] = 1.f / zsum[cb + e];
    }
    for (int e = threadIdx.x; e < 4 * TT * PITCH / 16; e += blockDim.x) reinterpret_cast<uint4*>(tF)[e] = make_uint4(0, 0, 0, 0);   // padding columns
    __syncthreads();
    // r_k = sum_n S[n,k] dS[n,k] = sum_v A[k,v] dA[k,v] — with the SAME bf16-rounded dA the mat-vec below uses, so that
    // sum_n dK[n,k] = sum_n S (dS - r) still cancels to fp32 round-off (the softmax over tokens is shift-invariant: the K bias
    // has an identically-zero gradient, and AdamW would turn a 2^-9-sized residual into +-lr steps); `rk` is the fp32 one
    for (int k = threadIdx.x; k < CH; k += blockDim.x) {
        float r = 0.f;
        for (int v = 0; v < CH; ++v) r = fmaf(A[hbase + (size_t)k * CH + v], __bfloat162float(sdA[k * LD + v]), r);
        srk[k] = r;
    }
    float bq[NT][2], bk[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) bq[j][0] = bq[j][1] = bk[j][0] = bk[j][1] = 0.f;
    const uint32_t* wA = reinterpret_cast<const uint32_t*>(sA);
    const uint32_t* wdA = reinterpret_cast<const uint32_t*>(sdA);
    const uint32_t* wdAt = reinterpret_cast<const uint32_t*>(sdAt);
    auto bfrag = [&](const uint32_t* m, int j, int t, uint32_t& b0, uint32_t& b1) {
        const int o = ((8 * j + gid) * LD + 16 * t + 2 * tig) >> 1;      // 32-bit word index
        b0 = m[o];
        b1 = m[o + 4];
    };
    const uint32_t uF = (uint32_t)__cvta_generic_to_shared(tF), uV = (uint32_t)__cvta_generic_to_shared(tV),
                   uS = (uint32_t)__cvta_generic_to_shared(tS);
    // ldmatrix (x4) addresses of this lane inside its warp's 16 rows: (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), ...
    const uint32_t a_off = (uint32_t)((warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16);
    // a CTA owns `tok_per_cta` tokens of this (image, head): the matrix staging above is amortised over all of them
    const int t_begin = blockIdx.x * tok_per_cta, t_end = min(N, t_begin + tok_per_cta);
    for (int n0 = t_begin; n0 < t_end; n0 += TT) {
        __syncthreads();      // (first pass: srk and the tables are ready; later: the previous tile's stores are done)
        // ---- stage 64 tokens: thread = (token row, 8-channel part)
        for (int it = threadIdx.x; it < TT * NT; it += 128) {
            const int row = it / NT, prt = it - row * NT;
            const int n = n0 + row;
            uint4 d4 = make_uint4(0, 0, 0, 0), k4 = d4, v4 = d4, e4 = d4;
            if (n < t_end) {
                const size_t tok = (size_t)b * N + n;
                d4 = *reinterpret_cast<const uint4*>(dy + tok * C + h * CH + prt * 8);
                k4 = *reinterpret_cast<const uint4*>(qkv + tok * 3 * C + C + h * CH + prt * 8);
                v4 = *reinterpret_cast<const uint4*>(qkv + tok * 3 * C + 2 * C + h * CH + prt * 8);
                e4 = *reinterpret_cast<const uint4*>(ein + tok * C + h * CH + prt * 8);
            }
            const uint32_t dw[4] = {d4.x, d4.y, d4.z, d4.w}, kw[4] = {k4.x, k4.y, k4.z, k4.w};
            uint32_t fw[4], sw[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int c = prt * 8 + 2 * w;
                const float2 d = up2(dw[w]), kk = up2(kw[w]);
                fw[w] = f2_to_bf2(sG[c] * d.x, sG[c + 1] * d.y);
                float2 sv = make_float2(0.f, 0.f);
                if (n < t_end) sv = make_float2(__expf(kk.x - sKm[c]) * sZi[c], __expf(kk.y - sKm[c + 1]) * sZi[c + 1]);
                sw[w] = f2_to_bf2(sv.x, sv.y);
                *reinterpret_cast<float2*>(tS32 + row * P32 + c) = sv;
            }
            const uint32_t so = (uint32_t)(row * PITCH + prt * 16);
            *reinterpret_cast<uint4*>(tF + so) = make_uint4(fw[0], fw[1], fw[2], fw[3]);
            *reinterpret_cast<uint4*>(tS + so) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
            *reinterpret_cast<uint4*>(tV + so) = v4;
            *reinterpret_cast<uint4*>(tE + so) = e4;
        }
        __syncthreads();
        // ---- this warp's 16 tokens: A fragments of dF, V, S
        uint32_t fdf[KS][4], fv[KS][4], fs[KS][4];
#pragma unroll
        for (int t = 0; t < KS; ++t) {
            ldsm_x4(fdf[t], uF + a_off + t * 32);
            ldsm_x4(fv[t], uV + a_off + t * 32);
            ldsm_x4(fs[t], uS + a_off + t * 32);
        }
        __syncwarp();      // every lane has its fragments: the rows may now be overwritten with the results
        const int rl[2] = {warp * 16 + gid, warp * 16 + gid + 8};       // tile rows of this thread's accumulator halves
        const bool rok[2] = {n0 + rl[0] < t_end, n0 + rl[1] < t_end};
        // ---- T1 -> dQ = s T1 + dF * E          (written over the dF rows)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < KS; ++t) {
                uint32_t b0, b1;
                bfrag(wA, j, t, b0, b1);
                mma_bf16_16816(acc, fdf[t], b0, b1);
            }
            const int c = 8 * j + 2 * tig;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t* pf = reinterpret_cast<uint32_t*>(tF + rl[r] * PITCH + c * 2);
                const float2 e = up2(*reinterpret_cast<const uint32_t*>(tE + rl[r] * PITCH + c * 2));
                const float2 df = up2(*pf);
                const float qx = fmaf(scale, acc[2 * r], df.x * e.x), qy = fmaf(scale, acc[2 * r + 1], df.y * e.y);
                if (rok[r]) {
                    bq[j][0] += qx;
                    bq[j][1] += qy;
                }
                *pf = f2_to_bf2(qx, qy);
            }
        }
        // ---- T2 -> dK = S * (T2 - r)           (written over the V rows)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < KS; ++t) {
                uint32_t b0, b1;
                bfrag(wdA, j, t, b0, b1);
                mma_bf16_16816(acc, fv[t], b0, b1);
            }
            const int c = 8 * j + 2 * tig;
            const float2 rr = make_float2(srk[c], srk[c + 1]);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float2 sv = *reinterpret_cast<const float2*>(tS32 + rl[r] * P32 + c);
                const float kx = sv.x * (acc[2 * r] - rr.x), ky = sv.y * (acc[2 * r + 1] - rr.y);
                if (rok[r]) {
                    bk[j][0] += kx;
                    bk[j][1] += ky;
                }
                *reinterpret_cast<uint32_t*>(tV + rl[r] * PITCH + c * 2) = f2_to_bf2(kx, ky);
            }
        }
        // ---- T3 -> mat-vec part of dV            (written over the S rows; parked in the dV slot of dqkv)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < KS; ++t) {
                uint32_t b0, b1;
                bfrag(wdAt, j, t, b0, b1);
                mma_bf16_16816(acc, fs[t], b0, b1);
            }
            const int c = 8 * j + 2 * tig;
#pragma unroll
            for (int r = 0; r < 2; ++r) *reinterpret_cast<uint32_t*>(tS + rl[r] * PITCH + c * 2) = f2_to_bf2(acc[2 * r], acc[2 * r + 1]);
        }
        __syncthreads();
        // ---- coalesced stores of the three result tiles
        for (int it = threadIdx.x; it < TT * NT; it += 128) {
            const int row = it / NT, prt = it - row * NT;
            const int n = n0 + row;
            if (n >= t_end) continue;
            bf16* o = dqkv + ((size_t)b * N + n) * 3 * C + h * CH + prt * 8;
            const uint32_t so = (uint32_t)(row * PITCH + prt * 16);
            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(tF + so);
            *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(tV + so);
            *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(tS + so);
        }
    }
    // ---- bias-gradient column sums of dQ and dK: over the 8 row groups of the warp, then the block, then one atomic per column
    if (dbias_qkv) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float a = bq[j][u], k2 = bk[j][u];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    k2 += __shfl_xor_sync(0xffffffffu, k2, o);
                }
                if (gid == 0) {
                    atomicAdd(&sbias[8 * j + 2 * tig + u], a);
                    atomicAdd(&sbias[CH + 8 * j + 2 * tig + u], k2);
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * CH; e += blockDim.x) {
            const int which = e / CH, c = e % CH;
            atomicAdd(dbias_qkv + which * C + h * CH + c, sbias[which * CH + c]);
        }
    }
}

template <int CH>
constexpr int attn_mm_smem_bytes() {
    constexpr int KS = (CH + 15) / 16, LD = KS * 16 + 8, PITCH = KS * 32 + 16;
    return 3 * CH * LD * 2 + 4 * 64 * PITCH + 64 * (CH + 2) * 4 + (2 * CH + CH + 3 * CH) * 4;
}

// One launch covers every channel group; the window of a group (that of its last head) selects the instantiation.
template <int CH>
__device__ __forceinline__ int win_of_group(int grp) { return win_of_head(((grp + 1) * Cfg<CH>::CPW - 1) / CH); }

template <int CH>
__global__ void __launch_bounds__(256) attn_fwd_strip_kernel(const bf16* __restrict__ qkv, const float* __restrict__ A,
                                                              const float* __restrict__ gate, CrpeW cw, bf16* __restrict__ out,
                                                              bf16* __restrict__ eout, float scale, int H, int Wd, int C) {
    MDV_PDL_SYNC();
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    const int win = win_of_group<CH>(blockIdx.y);
    if (win == 3) attn_fwd_strip_body<CH, 3>(qkv, A, gate, cw, out, eout, scale, H, Wd, C, smem_dyn);
    else if (win == 5) attn_fwd_strip_body<CH, 5>(qkv, A, gate, cw, out, eout, scale, H, Wd, C, smem_dyn);
    else attn_fwd_strip_body<CH, 7>(qkv, A, gate, cw, out, eout, scale, H, Wd, C, smem_dyn);
}

template <int CH, bool EXT>
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_strip_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                              const bf16* __restrict__ yout, const float* __restrict__ gate,
                                                              const float* __restrict__ A, const float* __restrict__ dA,
                                                              const float* __restrict__ rk, const float* __restrict__ kmax,
                                                              const float* __restrict__ zsum, CrpeW cw, CrpeG cg,
                                                              const bf16* __restrict__ ein, bf16* __restrict__ dqkv,
                                                              float* __restrict__ dgate, float* __restrict__ dbias_qkv, float scale,
                                                              int H, int Wd, int C) {
    MDV_PDL_SYNC();
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    const int win = win_of_group<CH>(blockIdx.y);
    const bool full = (H % TILE_H) == 0 && (Wd % TILE_W) == 0;
#define MDV_BWD_BODY(WIN_, FULL_) \
    attn_bwd_strip_body<CH, WIN_, BWD_TX, EXT, FULL_>(qkv, dy, yout, gate, A, dA, rk, kmax, zsum, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, H, Wd, C, smem_dyn)
    if (full) {
        if (win == 3) MDV_BWD_BODY(3, true);
        else if (win == 5) MDV_BWD_BODY(5, true);
        else MDV_BWD_BODY(7, true);
    } else {
        if (win == 3) MDV_BWD_BODY(3, false);
        else if (win == 5) MDV_BWD_BODY(5, false);
        else MDV_BWD_BODY(7, false);
    }
#undef MDV_BWD_BODY
}

// ------------------------------------------------------------------------------------------------ host side
template <typename K>
int set_smem(K kernel, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? MDV_OK : (int)e;
}

inline dim3 tile_grid(int B, int H, int W, int groups) {
    const int TH = H < TILE_H ? H : TILE_H, TW = W < TILE_W ? W : TILE_W;
    return dim3(mdv_cdiv(H, TH) * mdv_cdiv(W, TW), groups, B);
}

template <int CH>
int launch_fwd(const bf16* qkv, const float* A, const float* gate, const CrpeW& cw, bf16* out, bf16* eout, float scale, int B, int H, int W,
               int C, cudaStream_t st) {
    const int smem = tile_words(7) * 4 + 49 * 32 * 8 + CH * 32 * 8;     // sized for the widest window
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(attn_fwd_strip_kernel<CH>, smem);
        if (rc) return rc;
        configured = true;
    }
    mdv_launch(attn_fwd_strip_kernel<CH>, dim3(tile_grid(B, H, W, C / Cfg<CH>::CPW)), dim3(256), smem, st, qkv, A, gate, cw, out, eout, scale, H, W, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

template <int CH>
int launch_bwd(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* A, const float* dA, const float* rk,
               const float* kmax, const float* zsum, const CrpeW& cw, const CrpeG& cg, const bf16* ein, bf16* dqkv, float* dgate,
               float* dbias_qkv, float scale, int B, int H, int W, int C, cudaStream_t st) {
    // dE tile (fp32 pairs) + V tile (bf16 pairs) + taps + the three matrices, sized for the widest window
    const int full = bwd_tile_pos<7>() * 32 * 12 + 49 * 32 * 8 + (CH >= 40 ? 0 : 3 * (CH / 2) * 32 * 16);
    const int smem = cg.w[0] ? full : full - bwd_tile_pos<7>() * 32 * 4;      // activation-gradient-only pass: no V tile
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(attn_bwd_strip_kernel<CH, (CH >= 40)>, full);
        if (rc) return rc;
        configured = true;
    }
    constexpr bool EXT = CH >= 40;      // stages 2 / 3: the mat-vecs run on the tensor cores first
    if (EXT) {
        // tokens per CTA: the whole image up to 512 (more CTAs only while the grid would not fill the machine)
        int tpc = 64;
        while (tpc < 512 && (long long)mdv_cdiv(H * W, tpc) * (C / CH) * B > 4 * MDV_NUM_SMS) tpc *= 2;
        static bool mm_configured = false;
        if (!mm_configured) {
            int rc = set_smem(attn_mm_bwd_kernel<CH>, attn_mm_smem_bytes<CH>());
            if (rc) return rc;
            mm_configured = true;
        }
        mdv_launch(attn_mm_bwd_kernel<CH>, dim3(mdv_cdiv(H * W, tpc), C / CH, B), dim3(128), attn_mm_smem_bytes<CH>(), st, qkv, dy, ein, gate, A, dA, rk,
                   kmax, zsum, dqkv, dbias_qkv, scale, H * W, C, tpc);
        MDV_CHECK_LAUNCH();
    }
    mdv_launch((attn_bwd_strip_kernel<CH, EXT>), dim3(tile_grid(B, H, W, C / Cfg<CH>::CPW)), dim3(BWD_THREADS), smem, st, qkv, dy, yout, gate, A, dA, rk, kmax, zsum, cw, cg, ein, dqkv,
                                                                                       dgate, dbias_qkv, scale, H, W, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

// upper bound of chunks_for() over N: sizes the partial-sum scratch
int max_chunks(int B, int C, int Ch) {
    const int cpw = Ch <= 16 ? 64 : (Ch == 40 ? 40 : 64), ksplit = Ch <= 16 ? 1 : 2;
    int want = mdv_cdiv(4 * MDV_NUM_SMS, B * (C / cpw) * ksplit);
    if (want > 64) want = 64;
    return want < 1 ? 1 : want;
}

int chunks_for(int B, int N, int blocks_y) {
    int want = mdv_cdiv(4 * MDV_NUM_SMS, B * blocks_y);
    const int maxc = N / 64 > 0 ? N / 64 : 1;
    if (want > maxc) want = maxc;
    if (want > 64) want = 64;
    return want < 1 ? 1 : want;
}

template <int CH, int MODE>
int launch_outer(const bf16* qkv, const bf16* dy, const float* gate, const float* kmax, float* part, float* zpart, int B, int N, int C,
                 int& nchunk, cudaStream_t st) {
    // tensor-core version: one CTA per (token chunk, channel group, image); the chunk count stays within the scratch sized by
    // max_chunks().  (attn_outer_kernel, the FFMA2 version, is kept for A/B runs: MDV_ATTN_OUTER_FFMA=1.)
    static const bool ffma = getenv("MDV_ATTN_OUTER_FFMA") && atoi(getenv("MDV_ATTN_OUTER_FFMA")) != 0;
    if (!ffma) {
        constexpr int CPB = CH <= 16 ? 64 : CH;
        const int by = C / CPB;
        nchunk = chunks_for(B, N, by);
        const int mc = max_chunks(B, C, CH);
        if (nchunk > mc) nchunk = mc;
        int rpb = mdv_cdiv(N, nchunk);
        rpb = ((rpb + 63) / 64) * 64;
        nchunk = mdv_cdiv(N, rpb);
        mdv_launch((attn_outer_mma_kernel<CH, MODE>), dim3(dim3(nchunk, by, B)), dim3(32 * (CPB / 8)), 0, st, qkv, dy, gate, kmax, part, zpart, N, C, rpb,
                   nchunk);
        MDV_CHECK_LAUNCH();
        return MDV_OK;
    }
    using G = Cfg<CH>;
    const int by = (C / G::CPW) * G::KSPLIT;
    nchunk = chunks_for(B, N, by);
    int rpb = mdv_cdiv(N, nchunk);
    rpb = ((rpb + 31) / 32) * 32;
    nchunk = mdv_cdiv(N, rpb);
    mdv_launch((attn_outer_kernel<CH, MODE>), dim3(dim3(nchunk, by, B)), dim3(256), 0, st, qkv, dy, gate, kmax, part, zpart, N, C, rpb, nchunk);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

template <int CH>
int fwd_impl(const bf16* qkv, const float* gate, const CrpeW& cw, float* kmax, float* zsum, float* A, float* ws, bf16* out, bf16* eout,
             float scale, int B, int H, int W, int C, cudaStream_t st) {
    const int N = H * W;
    int nchunk = 0;
    float* part = ws;
    // zpart sits behind the largest possible partial buffer of this shape
    const size_t mc = (size_t)max_chunks(B, C, CH);
    float* zpart = ws + (size_t)B * mc * C * CH;
    int rc = launch_outer<CH, 0>(qkv, nullptr, nullptr, kmax, part, zpart, B, N, C, nchunk, st);
    if (rc) return rc;
    const long long tot = (long long)B * C * CH;
    mdv_launch(attn_combine_fwd_kernel, dim3(mdv_cdiv(tot, 256)), dim3(256), 0, st, part, zpart, A, zsum, C, CH, nchunk, tot);
    MDV_CHECK_LAUNCH();
    return launch_fwd<CH>(qkv, A, gate, cw, out, eout, scale, B, H, W, C, st);
}

template <int CH>
int bwd_impl(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* kmax, const float* zsum, const float* A,
             float* ws, const CrpeW& cw, const CrpeG& cg, const bf16* ein, bf16* dqkv, float* dgate, float* dbias_qkv, float scale, int B,
             int H, int W, int C, cudaStream_t st) {
    const int N = H * W;
    int nchunk = 0;
    float* part = ws;
    const size_t mc = (size_t)max_chunks(B, C, CH);
    float* dA = ws + (size_t)B * mc * C * CH + (size_t)B * mc * C;
    float* rk = dA + (size_t)B * C * CH;
    int rc = launch_outer<CH, 1>(qkv, dy, gate, nullptr, part, nullptr, B, N, C, nchunk, st);
    if (rc) return rc;
    mdv_launch(attn_combine_bwd_kernel, dim3(mdv_cdiv(B * C, 8)), dim3(256), 0, st, part, A, dA, rk, scale, C, CH, nchunk, B * C);
    MDV_CHECK_LAUNCH();
    return launch_bwd<CH>(qkv, dy, yout, gate, A, dA, rk, kmax, zsum, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, B, H, W, C, st);
}

}  // namespace

// scratch floats needed by attn_strip_fwd / attn_strip_bwd: partial sums (<= 64 chunks) + zpart + dA + rk
long long attn_strip_ws_floats(int B, int C, int Ch) {
    const long long mc = max_chunks(B, C, Ch);
    return (long long)B * mc * C * Ch + (long long)B * mc * C + (long long)B * C * Ch + (long long)B * C;
}

int attn_strip_fwd(const bf16* qkv, const float* gate, const CrpeW& cw, float* kmax, float* zsum, float* A, float* ws, bf16* out,
                   bf16* eout, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st) {
    switch (Ch) {
        case 8: return fwd_impl<8>(qkv, gate, cw, kmax, zsum, A, ws, out, eout, scale, B, H, W, C, st);
        case 16: return fwd_impl<16>(qkv, gate, cw, kmax, zsum, A, ws, out, eout, scale, B, H, W, C, st);
        case 40: return fwd_impl<40>(qkv, gate, cw, kmax, zsum, A, ws, out, eout, scale, B, H, W, C, st);
        case 64: return fwd_impl<64>(qkv, gate, cw, kmax, zsum, A, ws, out, eout, scale, B, H, W, C, st);
        default: return MDV_ERR_UNSUPPORTED;
    }
}

int attn_strip_bwd(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* kmax, const float* zsum,
                   const float* A, float* ws, const CrpeW& cw, const CrpeG& cg, const bf16* ein, bf16* dqkv, float* dgate,
                   float* dbias_qkv, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st) {
    switch (Ch) {
        case 8: return bwd_impl<8>(qkv, dy, yout, gate, kmax, zsum, A, ws, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, B, H, W, C, st);
        case 16: return bwd_impl<16>(qkv, dy, yout, gate, kmax, zsum, A, ws, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, B, H, W, C, st);
        case 40: return bwd_impl<40>(qkv, dy, yout, gate, kmax, zsum, A, ws, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, B, H, W, C, st);
        case 64: return bwd_impl<64>(qkv, dy, yout, gate, kmax, zsum, A, ws, cw, cg, ein, dqkv, dgate, dbias_qkv, scale, B, H, W, C, st);
        default: return MDV_ERR_UNSUPPORTED;
    }
}
