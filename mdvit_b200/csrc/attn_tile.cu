// Shared-memory tiled kernels for the token-parallel half of the factorized attention (phase 2 and its backward).
//
// A block owns a TH x TW pixel tile (<= 16 x 16) of one image and 32 consecutive channels; the V tile (and, in backward,
// the dE = g*dY*Q tile) is staged with its convolution halo in shared memory as [position][32 ch] bf16, so every tap of the
// 3x3 / 5x5 / 7x7 depthwise relative-position convolution (mpvit.py:306-316) is a conflict-free 64-byte shared read.
// lane == channel, so per-channel constants (filter taps, gate, softmax stats, and for Ch <= 16 the head's Ch x Ch
// matrices) live in registers and the per-head mat-vecs are warp shuffles.  For Ch = 40 / 64 the mat-vecs are done by a
// small shared-memory "head GEMM" kernel first and this kernel adds the convolution terms.
#include "attn_internal.cuh"

namespace {

constexpr int TILE = 16;

__device__ __forceinline__ int win_of_channel(int c, int Ch) {
    const int h = c / Ch;
    return h < 2 ? 3 : (h < 5 ? 5 : 7);
}
__device__ __forceinline__ void crpe_ptrs(const CrpeW& cw, int c, int Ch, const float*& w, const float*& b, int& win) {
    const int h = c / Ch;
    const int grp = h < 2 ? 0 : (h < 5 ? 1 : 2);
    const int cl = c - (grp == 0 ? 0 : (grp == 1 ? 2 * Ch : 5 * Ch));
    win = 3 + 2 * grp;
    w = cw.w[grp] + (size_t)cl * win * win;
    b = cw.b[grp] + cl;
}

struct TileGeom {
    int ty0, tx0, th, tw, R, PW, PH;
};
__device__ __forceinline__ TileGeom tile_geom(int H, int Wd, int WIN) {
    TileGeom g;
    const int TH = min(H, TILE), TW = min(Wd, TILE);
    const int tiles_x = (Wd + TW - 1) / TW;
    g.ty0 = (blockIdx.x / tiles_x) * TH;
    g.tx0 = (blockIdx.x % tiles_x) * TW;
    g.th = min(TH, H - g.ty0);
    g.tw = min(TW, Wd - g.tx0);
    g.R = WIN >> 1;
    g.PW = g.tw + 2 * g.R;
    g.PH = g.th + 2 * g.R;
    return g;
}

// stage src[b, pos, ch0 .. ch0+32) (bf16, row pitch ld) for all halo positions; zero outside the image
__device__ __forceinline__ void load_halo(bf16* dst, const bf16* __restrict__ src_b, int ld, int ch0, const TileGeom& g, int H, int Wd) {
    const int npos = g.PH * g.PW;
    for (int e = threadIdx.x; e < npos * 4; e += blockDim.x) {
        const int pos = e >> 2, part = e & 3;
        const int y = g.ty0 - g.R + pos / g.PW, x = g.tx0 - g.R + pos % g.PW;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < H && x >= 0 && x < Wd) v = *reinterpret_cast<const uint4*>(src_b + (size_t)(y * Wd + x) * ld + ch0 + part * 8);
        *reinterpret_cast<uint4*>(dst + pos * 32 + part * 8) = v;
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int WIN, int CH>
__device__ __forceinline__ void fwd_tile_body(const bf16* __restrict__ qkv, const float* __restrict__ A, const float* __restrict__ gate,
                                              const CrpeW& cw, bf16* __restrict__ out, float scale, int H, int Wd, int C, bf16* sV) {
    const int N = H * Wd;
    const int b = blockIdx.z, ch0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const TileGeom g = tile_geom(H, Wd, WIN);
    const bf16* qkv_b = qkv + (size_t)b * N * 3 * C;
    load_halo(sV, qkv_b + 2 * C, 3 * C, ch0, g, H, Wd);
    const int c = ch0 + lane;
    // per-channel constants
    float w[WIN * WIN];
    {
        const float* wp; const float* bp; int wc;
        crpe_ptrs(cw, c, CH, wp, bp, wc);
        const int o = (WIN - wc) >> 1;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const int ii = i - o, jj = j - o;
                w[i * WIN + j] = (ii >= 0 && ii < wc && jj >= 0 && jj < wc) ? __ldg(wp + ii * wc + jj) : 0.f;
            }
    }
    const float* wp_; const float* bp_; int wc_;
    crpe_ptrs(cw, c, CH, wp_, bp_, wc_);
    const float bias = __ldg(bp_);
    const float gt = gate ? __ldg(gate + (size_t)b * C + c) : 1.f;
    constexpr bool FUSED = CH <= 16;
    constexpr int NA = FUSED ? CH : 1;
    float Acol[NA];
    const int hb = (lane / (FUSED ? CH : 32)) * (FUSED ? CH : 32);   // first lane of this lane's head
    if (FUSED) {
        const int h = c / CH, v = c % CH;
#pragma unroll
        for (int k = 0; k < NA; ++k) Acol[k] = __ldg(A + ((size_t)b * C + h * CH + k) * CH + v);
    }
    __syncthreads();
    const int npix = g.th * g.tw;
    for (int p = warp; p < npix; p += nwarp) {
        const int py = p / g.tw, px = p % g.tw;
        const size_t n = (size_t)(g.ty0 + py) * Wd + (g.tx0 + px);
        const float q = __bfloat162float(qkv_b[n * 3 * C + c]);
        float e = bias;
        const bf16* sp = sV + (py * g.PW + px) * 32 + lane;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) e += w[i * WIN + j] * __bfloat162float(sp[(i * g.PW + j) * 32]);
        float fa;
        if (FUSED) {
            fa = 0.f;
#pragma unroll
            for (int k = 0; k < NA; ++k) fa += __shfl_sync(0xffffffffu, q, hb + k) * Acol[k];
            fa *= scale;
        } else {
            fa = __bfloat162float(out[((size_t)b * N + n) * C + c]);   // scale * Q.A written by the head-GEMM kernel
        }
        out[((size_t)b * N + n) * C + c] = __float2bfloat16_rn(gt * (fa + q * e));
    }
}

template <int CH>
__global__ void __launch_bounds__(256) attn_fwd_tile_kernel(const bf16* __restrict__ qkv, const float* __restrict__ A,
                                                             const float* __restrict__ gate, CrpeW cw, bf16* __restrict__ out,
                                                             float scale, int H, int Wd, int C) {
    extern __shared__ __align__(16) uint8_t smem_t[];
    bf16* sV = reinterpret_cast<bf16*>(smem_t);
    const int win = win_of_channel(blockIdx.y * 32 + 31, CH);   // windows grow with the channel index
    if (win == 3) fwd_tile_body<3, CH>(qkv, A, gate, cw, out, scale, H, Wd, C, sV);
    else if (win == 5) fwd_tile_body<5, CH>(qkv, A, gate, cw, out, scale, H, Wd, C, sV);
    else fwd_tile_body<7, CH>(qkv, A, gate, cw, out, scale, H, Wd, C, sV);
}

// head GEMM (Ch = 40 / 64): out[n, (h,v)] = bf16( scale * sum_k Q[n,h,k] A[h,k,v] );  block = 32 tokens x one head
template <int CH>
__global__ void __launch_bounds__(256) attn_head_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ A,
                                                             bf16* __restrict__ out, float scale, int N, int C) {
    __shared__ float sA[CH * CH];
    __shared__ float sQ[32 * CH];
    const int b = blockIdx.z, h = blockIdx.y, n0 = blockIdx.x * 32;
    const int nt = min(32, N - n0);
    for (int e = threadIdx.x; e < CH * CH; e += 256) sA[e] = __ldg(A + ((size_t)b * C + h * CH) * CH + e);
    for (int e = threadIdx.x; e < 32 * CH; e += 256) {
        const int t = e / CH, k = e % CH;
        sQ[e] = t < nt ? __bfloat162float(qkv[((size_t)b * N + n0 + t) * 3 * C + h * CH + k]) : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 32 * CH; o += 256) {
        const int t = o / CH, v = o % CH;
        if (t >= nt) continue;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < CH; ++k) s += sQ[t * CH + k] * sA[k * CH + v];
        out[((size_t)b * N + n0 + t) * C + h * CH + v] = __float2bfloat16_rn(scale * s);
    }
}

// ------------------------------------------------------------------------------------------------ backward (dQ, dK, dV)
template <int WIN, int CH>
__device__ __forceinline__ void bwd_tile_body(const bf16* __restrict__ qkv, const bf16* __restrict__ dy, const float* __restrict__ gate,
                                              const float* __restrict__ A, const float* __restrict__ dA, const float* __restrict__ rk,
                                              const float* __restrict__ kmax, const float* __restrict__ zsum, const CrpeW& cw,
                                              bf16* __restrict__ dqkv, float scale, int H, int Wd, int C, bf16* sV, bf16* sE) {
    const int N = H * Wd;
    const int b = blockIdx.z, ch0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const TileGeom g = tile_geom(H, Wd, WIN);
    const bf16* qkv_b = qkv + (size_t)b * N * 3 * C;
    const bf16* dy_b = dy + (size_t)b * N * C;
    load_halo(sV, qkv_b + 2 * C, 3 * C, ch0, g, H, Wd);
    {   // dE = g * dY * Q at every halo position
        const int npos = g.PH * g.PW;
        for (int e = threadIdx.x; e < npos * 4; e += blockDim.x) {
            const int pos = e >> 2, part = e & 3;
            const int y = g.ty0 - g.R + pos / g.PW, x = g.tx0 - g.R + pos % g.PW;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (y >= 0 && y < H && x >= 0 && x < Wd) {
                const size_t n = (size_t)y * Wd + x;
                const uint4 dv = *reinterpret_cast<const uint4*>(dy_b + n * C + ch0 + part * 8);
                const uint4 qv = *reinterpret_cast<const uint4*>(qkv_b + n * 3 * C + ch0 + part * 8);
                const uint32_t d4[4] = {dv.x, dv.y, dv.z, dv.w}, q4[4] = {qv.x, qv.y, qv.z, qv.w};
                uint32_t o4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 d = bf2_to_f2(d4[u]), q = bf2_to_f2(q4[u]);
                    const int cc = ch0 + part * 8 + 2 * u;
                    const float g0 = gate ? __ldg(gate + (size_t)b * C + cc) : 1.f, g1 = gate ? __ldg(gate + (size_t)b * C + cc + 1) : 1.f;
                    o4[u] = f2_to_bf2(g0 * d.x * q.x, g1 * d.y * q.y);
                }
                r = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            }
            *reinterpret_cast<uint4*>(sE + pos * 32 + part * 8) = r;
        }
    }
    const int c = ch0 + lane;
    float w[WIN * WIN];
    const float* wp; const float* bp; int wc;
    crpe_ptrs(cw, c, CH, wp, bp, wc);
    {
        const int o = (WIN - wc) >> 1;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const int ii = i - o, jj = j - o;
                w[i * WIN + j] = (ii >= 0 && ii < wc && jj >= 0 && jj < wc) ? __ldg(wp + ii * wc + jj) : 0.f;
            }
    }
    const float bias = __ldg(bp);
    const size_t bc = (size_t)b * C + c;
    const float gt = gate ? __ldg(gate + bc) : 1.f;
    const float km = __ldg(kmax + bc), zinv = 1.f / __ldg(zsum + bc), rkc = __ldg(rk + bc);
    constexpr bool FUSED = CH <= 16;
    constexpr int NA = FUSED ? CH : 1;
    float Arow[NA], dAcol[NA], dArow[NA];
    const int hb = (lane / (FUSED ? CH : 32)) * (FUSED ? CH : 32);
    if (FUSED) {
        const int h = c / CH, v = c % CH;
        const size_t hbase = ((size_t)b * C + h * CH) * CH;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            Arow[j] = __ldg(A + hbase + (size_t)v * CH + j);     // A[k=v][j]
            dArow[j] = __ldg(dA + hbase + (size_t)v * CH + j);   // dA[k=v][j]
            dAcol[j] = __ldg(dA + hbase + (size_t)j * CH + v);   // dA[j][v]
        }
    }
    __syncthreads();
    const int npix = g.th * g.tw;
    for (int p = warp; p < npix; p += nwarp) {
        const int py = p / g.tw, px = p % g.tw;
        const size_t n = (size_t)(g.ty0 + py) * Wd + (g.tx0 + px);
        const float q = __bfloat162float(qkv_b[n * 3 * C + c]);
        const float kk = __bfloat162float(qkv_b[n * 3 * C + C + c]);
        const float dyc = __bfloat162float(dy_b[n * C + c]);
        const bf16* spv = sV + (py * g.PW + px) * 32 + lane;
        const bf16* spe = sE + (py * g.PW + px) * 32 + lane;
        const float vv = __bfloat162float(spv[(g.R * g.PW + g.R) * 32]);
        float e = bias, tconv = 0.f;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                e += w[i * WIN + j] * __bfloat162float(spv[(i * g.PW + j) * 32]);
                // transposed conv: dE at n - (i-R, j-R)  ==  halo position (WIN-1-i, WIN-1-j)
                tconv += w[i * WIN + j] * __bfloat162float(spe[((WIN - 1 - i) * g.PW + (WIN - 1 - j)) * 32]);
            }
        const float dF = gt * dyc;
        const float S = __expf(kk - km) * zinv;
        float dq, dk, dv;
        bf16* drow = dqkv + ((size_t)b * N + n) * 3 * C;
        if (FUSED) {
            float sq = 0.f, sv = 0.f, sk = 0.f;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                sq += __shfl_sync(0xffffffffu, dF, hb + j) * Arow[j];
                sv += __shfl_sync(0xffffffffu, S, hb + j) * dAcol[j];
                sk += __shfl_sync(0xffffffffu, vv, hb + j) * dArow[j];
            }
            dq = scale * sq + dF * e;
            dk = S * (sk - rkc);
            dv = sv + tconv;
            drow[C + c] = __float2bfloat16_rn(dk);
        } else {
            dq = __bfloat162float(drow[c]) + dF * e;          // head-GEMM kernel left scale*sq, final dK and sv in dqkv
            dv = __bfloat162float(drow[2 * C + c]) + tconv;
        }
        drow[c] = __float2bfloat16_rn(dq);
        drow[2 * C + c] = __float2bfloat16_rn(dv);
    }
}

template <int CH>
__global__ void __launch_bounds__(256) attn_bwd_tile_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                             const float* __restrict__ gate, const float* __restrict__ A,
                                                             const float* __restrict__ dA, const float* __restrict__ rk,
                                                             const float* __restrict__ kmax, const float* __restrict__ zsum, CrpeW cw,
                                                             bf16* __restrict__ dqkv, float scale, int H, int Wd, int C) {
    extern __shared__ __align__(16) uint8_t smem_t[];
    bf16* sV = reinterpret_cast<bf16*>(smem_t);
    bf16* sE = sV + (TILE + 6) * (TILE + 6) * 32;
    const int win = win_of_channel(blockIdx.y * 32 + 31, CH);
    if (win == 3) bwd_tile_body<3, CH>(qkv, dy, gate, A, dA, rk, kmax, zsum, cw, dqkv, scale, H, Wd, C, sV, sE);
    else if (win == 5) bwd_tile_body<5, CH>(qkv, dy, gate, A, dA, rk, kmax, zsum, cw, dqkv, scale, H, Wd, C, sV, sE);
    else bwd_tile_body<7, CH>(qkv, dy, gate, A, dA, rk, kmax, zsum, cw, dqkv, scale, H, Wd, C, sV, sE);
}

// head GEMM backward (Ch = 40 / 64): dqkv.q = scale * dF.A^T ; dqkv.k = S * (V.dA^T - r) (final) ; dqkv.v = S.dA
template <int CH>
__global__ void __launch_bounds__(256) attn_head_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                             const float* __restrict__ gate, const float* __restrict__ At,
                                                             const float* __restrict__ dA, const float* __restrict__ dAt,
                                                             const float* __restrict__ rk, const float* __restrict__ kmax,
                                                             const float* __restrict__ zsum, bf16* __restrict__ dqkv, float scale, int N,
                                                             int C) {
    extern __shared__ __align__(16) uint8_t smem_h[];
    float* sAt = reinterpret_cast<float*>(smem_h);       // At[j][k] = A[k][j]
    float* sdA = sAt + CH * CH;                           // dA[j][v]
    float* sdAt = sdA + CH * CH;                          // dAt[j][k] = dA[k][j]
    float* sF = sdAt + CH * CH;                           // dF[t][j]
    float* sS = sF + 32 * CH;                             // S[t][j]
    float* sVv = sS + 32 * CH;                            // V[t][j]
    const int b = blockIdx.z, h = blockIdx.y, n0 = blockIdx.x * 32;
    const int nt = min(32, N - n0);
    const size_t hbase = ((size_t)b * C + h * CH) * CH;
    const size_t bc0 = (size_t)b * C + h * CH;
    for (int e = threadIdx.x; e < CH * CH; e += 256) {
        sAt[e] = __ldg(At + hbase + e);
        sdA[e] = __ldg(dA + hbase + e);
        sdAt[e] = __ldg(dAt + hbase + e);
    }
    for (int e = threadIdx.x; e < 32 * CH; e += 256) {
        const int t = e / CH, j = e % CH;
        float f = 0.f, s = 0.f, v = 0.f;
        if (t < nt) {
            const size_t tok = (size_t)b * N + n0 + t;
            f = (gate ? __ldg(gate + bc0 + j) : 1.f) * __bfloat162float(dy[tok * C + h * CH + j]);
            s = __expf(__bfloat162float(qkv[tok * 3 * C + C + h * CH + j]) - __ldg(kmax + bc0 + j)) / __ldg(zsum + bc0 + j);
            v = __bfloat162float(qkv[tok * 3 * C + 2 * C + h * CH + j]);
        }
        sF[e] = f; sS[e] = s; sVv[e] = v;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 32 * CH; o += 256) {
        const int t = o / CH, c = o % CH;
        if (t >= nt) continue;
        float sq = 0.f, sv = 0.f, sk = 0.f;
#pragma unroll 8
        for (int j = 0; j < CH; ++j) {
            sq += sF[t * CH + j] * sAt[j * CH + c];
            sv += sS[t * CH + j] * sdA[j * CH + c];
            sk += sVv[t * CH + j] * sdAt[j * CH + c];
        }
        bf16* drow = dqkv + ((size_t)b * N + n0 + t) * 3 * C + h * CH + c;
        drow[0] = __float2bfloat16_rn(scale * sq);
        drow[C] = __float2bfloat16_rn(sS[t * CH + c] * (sk - __ldg(rk + bc0 + c)));
        drow[2 * C] = __float2bfloat16_rn(sv);
    }
}

// ------------------------------------------------------------------------------------------------ backward (per-channel sums)
// dWconv[c,tap] += sum_n dE[n,c] V[n+tap,c];  dbconv[c] += sum_n dE[n,c];  dgate[b,c] += sum_n dY[n,c] y[n,c] / g[b,c]
template <int WIN, int CH>
__device__ __forceinline__ void wgrad_tile_body(const bf16* __restrict__ qkv, const bf16* __restrict__ dy, const bf16* __restrict__ yout,
                                                const float* __restrict__ gate, const CrpeG& cg, float* __restrict__ dgate, int H,
                                                int Wd, int C, bf16* sV, float* red) {
    const int N = H * Wd;
    const int b = blockIdx.z, ch0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const TileGeom g = tile_geom(H, Wd, WIN);
    const bf16* qkv_b = qkv + (size_t)b * N * 3 * C;
    load_halo(sV, qkv_b + 2 * C, 3 * C, ch0, g, H, Wd);
    const int c = ch0 + lane;
    const float gt = gate ? __ldg(gate + (size_t)b * C + c) : 1.f;
    float acc[WIN * WIN];
#pragma unroll
    for (int t = 0; t < WIN * WIN; ++t) acc[t] = 0.f;
    float accb = 0.f, accg = 0.f;
    __syncthreads();
    const int npix = g.th * g.tw;
    for (int p = warp; p < npix; p += nwarp) {
        const int py = p / g.tw, px = p % g.tw;
        const size_t n = (size_t)(g.ty0 + py) * Wd + (g.tx0 + px);
        const float q = __bfloat162float(qkv_b[n * 3 * C + c]);
        const float d = __bfloat162float(dy[((size_t)b * N + n) * C + c]);
        const float de = gt * d * q;
        accb += de;
        if (gate) accg += d * __bfloat162float(yout[((size_t)b * N + n) * C + c]);
        if (!cg.w[0]) continue;   // dgate-only pass (weight gradients off)
        const bf16* spv = sV + (py * g.PW + px) * 32 + lane;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) acc[i * WIN + j] += de * __bfloat162float(spv[(i * g.PW + j) * 32]);
    }
    __syncthreads();   // sV no longer needed; `red` may alias it
    constexpr int NR = WIN * WIN + 2;
    float* mine = red + ((size_t)warp * 32 + lane) * NR;
#pragma unroll
    for (int t = 0; t < WIN * WIN; ++t) mine[t] = acc[t];
    mine[WIN * WIN] = accb;
    mine[WIN * WIN + 1] = gate ? accg / gt : 0.f;
    __syncthreads();
    for (int o = threadIdx.x; o < 32 * NR; o += blockDim.x) {
        const int cc = o / NR, t = o % NR;
        float s = 0.f;
        for (int k = 0; k < nwarp; ++k) s += red[((size_t)k * 32 + cc) * NR + t];
        const int cgl = ch0 + cc;
        const int h = cgl / CH;
        const int grp = h < 2 ? 0 : (h < 5 ? 1 : 2);
        const int cl = cgl - (grp == 0 ? 0 : (grp == 1 ? 2 * CH : 5 * CH));
        const int wc = 3 + 2 * grp;
        if (t <= WIN * WIN && !cg.w[0]) continue;
        if (t < WIN * WIN) {
            const int o2 = (WIN - wc) >> 1;
            const int ii = t / WIN - o2, jj = t % WIN - o2;
            if (ii >= 0 && ii < wc && jj >= 0 && jj < wc) atomicAdd(cg.w[grp] + (size_t)cl * wc * wc + ii * wc + jj, s);
        } else if (t == WIN * WIN) {
            atomicAdd(cg.b[grp] + cl, s);
        } else if (dgate) {
            atomicAdd(dgate + (size_t)b * C + cgl, s);
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(256) attn_wgrad_tile_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                               const bf16* __restrict__ yout, const float* __restrict__ gate, CrpeG cg,
                                                               float* __restrict__ dgate, int H, int Wd, int C) {
    extern __shared__ __align__(16) uint8_t smem_t[];
    bf16* sV = reinterpret_cast<bf16*>(smem_t);
    float* red = reinterpret_cast<float*>(smem_t);
    const int win = win_of_channel(blockIdx.y * 32 + 31, CH);
    if (win == 3) wgrad_tile_body<3, CH>(qkv, dy, yout, gate, cg, dgate, H, Wd, C, sV, red);
    else if (win == 5) wgrad_tile_body<5, CH>(qkv, dy, yout, gate, cg, dgate, H, Wd, C, sV, red);
    else wgrad_tile_body<7, CH>(qkv, dy, yout, gate, cg, dgate, H, Wd, C, sV, red);
}

constexpr int HALO_BYTES = (TILE + 6) * (TILE + 6) * 32 * 2;          // 30976
constexpr int RED_BYTES = 8 * 32 * 51 * 4;                            // 52224

template <typename K>
int set_smem(K kernel, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? MDV_OK : (int)e;
}

inline dim3 tile_grid(int B, int H, int W, int C) {
    const int TH = H < TILE ? H : TILE, TW = W < TILE ? W : TILE;
    return dim3(mdv_cdiv(H, TH) * mdv_cdiv(W, TW), C / 32, B);
}

}  // namespace

template <int CH>
static int fwd_impl(const bf16* qkv, const float* A, const float* gate, const CrpeW& cw, bf16* out, float scale, int B, int H, int W,
                    int C, cudaStream_t st) {
    const int N = H * W;
    if (CH > 16) {
        attn_head_fwd_kernel<CH><<<dim3(mdv_cdiv(N, 32), C / CH, B), 256, 0, st>>>(qkv, A, out, scale, N, C);
        MDV_CHECK_LAUNCH();
    }
    attn_fwd_tile_kernel<CH><<<tile_grid(B, H, W, C), 256, HALO_BYTES, st>>>(qkv, A, gate, cw, out, scale, H, W, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

int attn_tile_fwd(const bf16* qkv, const float* A, const float* gate, const CrpeW& cw, bf16* out, float scale, int B, int H, int W, int C,
                  int Ch, cudaStream_t st) {
    switch (Ch) {
        case 8: return fwd_impl<8>(qkv, A, gate, cw, out, scale, B, H, W, C, st);
        case 16: return fwd_impl<16>(qkv, A, gate, cw, out, scale, B, H, W, C, st);
        case 40: return fwd_impl<40>(qkv, A, gate, cw, out, scale, B, H, W, C, st);
        case 64: return fwd_impl<64>(qkv, A, gate, cw, out, scale, B, H, W, C, st);
        default: return MDV_ERR_UNSUPPORTED;
    }
}

template <int CH>
static int bwd_impl(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* A, const float* At, const float* dA,
                    const float* dAt, const float* rk, const float* kmax, const float* zsum, const CrpeW& cw, const CrpeG& cg,
                    bf16* dqkv, float* dgate, float scale, int B, int H, int W, int C, cudaStream_t st) {
    const int N = H * W;
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(attn_bwd_tile_kernel<CH>, 2 * HALO_BYTES);
        if (rc) return rc;
        rc = set_smem(attn_wgrad_tile_kernel<CH>, RED_BYTES);
        if (rc) return rc;
        if (CH > 16) {
            rc = set_smem(attn_head_bwd_kernel<CH>, (3 * CH * CH + 3 * 32 * CH) * 4);
            if (rc) return rc;
        }
        configured = true;
    }
    if (CH > 16) {
        attn_head_bwd_kernel<CH><<<dim3(mdv_cdiv(N, 32), C / CH, B), 256, (3 * CH * CH + 3 * 32 * CH) * 4, st>>>(
            qkv, dy, gate, At, dA, dAt, rk, kmax, zsum, dqkv, scale, N, C);
        MDV_CHECK_LAUNCH();
    }
    attn_bwd_tile_kernel<CH><<<tile_grid(B, H, W, C), 256, 2 * HALO_BYTES, st>>>(qkv, dy, gate, A, dA, rk, kmax, zsum, cw, dqkv, scale, H, W, C);
    MDV_CHECK_LAUNCH();
    if (cg.w[0] || dgate) {
        attn_wgrad_tile_kernel<CH><<<tile_grid(B, H, W, C), 256, RED_BYTES, st>>>(qkv, dy, yout, gate, cg, dgate, H, W, C);
        MDV_CHECK_LAUNCH();
    }
    return MDV_OK;
}

int attn_tile_bwd(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* A, const float* At, const float* dA,
                  const float* dAt, const float* rk, const float* kmax, const float* zsum, const CrpeW& cw, const CrpeG& cg, bf16* dqkv,
                  float* dgate, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st) {
    switch (Ch) {
        case 8: return bwd_impl<8>(qkv, dy, yout, gate, A, At, dA, dAt, rk, kmax, zsum, cw, cg, dqkv, dgate, scale, B, H, W, C, st);
        case 16: return bwd_impl<16>(qkv, dy, yout, gate, A, At, dA, dAt, rk, kmax, zsum, cw, cg, dqkv, dgate, scale, B, H, W, C, st);
        case 40: return bwd_impl<40>(qkv, dy, yout, gate, A, At, dA, dAt, rk, kmax, zsum, cw, cg, dqkv, dgate, scale, B, H, W, C, st);
        case 64: return bwd_impl<64>(qkv, dy, yout, gate, A, At, dA, dAt, rk, kmax, zsum, cw, cg, dqkv, dgate, scale, B, H, W, C, st);
        default: return MDV_ERR_UNSUPPORTED;
    }
}
