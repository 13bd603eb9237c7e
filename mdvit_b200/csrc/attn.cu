// Factorized attention with convolutional relative position encoding and the per-domain head gate (DA).
//   reference: FactorAtt_ConvRelPosEnc_Sup.forward  Models/Transformer/mdvit.py:288-304
//              ConvRelPosEnc.forward                Models/Transformer/mpvit.py:296-318
//              domain_layer + softmax over heads    Models/Transformer/mdvit.py:272-276,301-303
//
// Layout: qkv is the token-major output of the QKV GEMM, bf16 [B, N, 3C] with channel order (q|k|v) x (head, Ch)
// (mdvit.py:288-290), so no permute/contiguous copy is ever made.  Per (batch, head):
//     m_k = max_n K[n,k]   Z_k = sum_n exp(K[n,k]-m_k)   A[k,v] = sum_n exp(K[n,k]-m_k) V[n,v] / Z_k        (phase 1)
//     Y[n,v] = g[v] * ( s * sum_k Q[n,k] A[k,v] + Q[n,v] * (dwconv(V)[n,v] + b[v]) )                         (phase 2)
// Everything here is HBM/L2-bound (matmul FLOPs < 1% of the model); the cross-token quantities are
// (2+Ch)*C floats per image.  Backward follows SURVEY.md App. E.
#include "../../include/mdvit_b200.h"
#include "common.cuh"

namespace {

struct CrpeW {               // three depthwise filters: heads [0,2) 3x3, [2,5) 5x5, [5,8) 7x7  (mdvit.py:423)
    const float* w[3];
    const float* b[3];
};
struct CrpeG {
    float* w[3];
    float* b[3];
};

__device__ __forceinline__ void crpe_lookup(int c, int Ch, int& grp, int& cl, int& win) {
    const int h = c / Ch;
    grp = h < 2 ? 0 : (h < 5 ? 1 : 2);
    cl = c - (grp == 0 ? 0 : (grp == 1 ? 2 * Ch : 5 * Ch));
    win = 3 + 2 * grp;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// ---------------------------------------------------------------------------------- phase 1a: column max of K
// kmax must be pre-filled with 0xFF bytes (identity of the atomic above).  thread = 2 channels, ty = row lane.
__global__ void __launch_bounds__(256) attn_colmax_kernel(const bf16* __restrict__ qkv, float* __restrict__ kmax, int N, int C,
                                                           int rows_per_block) {
    const int half = C >> 1;
    const int tx = threadIdx.x % half, ty = threadIdx.x / half, nty = blockDim.x / half;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    float m0 = -INFINITY, m1 = -INFINITY;
    const bf16* base = qkv + (size_t)b * N * 3 * C + C + tx * 2;
    for (int r = r0 + ty; r < r1; r += nty) {
        float2 v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(base + (size_t)r * 3 * C));
        m0 = fmaxf(m0, v.x);
        m1 = fmaxf(m1, v.y);
    }
    if (m0 > -INFINITY) atomic_max_float(kmax + (size_t)b * C + tx * 2, m0);
    if (m1 > -INFINITY) atomic_max_float(kmax + (size_t)b * C + tx * 2 + 1, m1);
}

// ---------------------------------------------------------------------------------- phase 1b / bwd: per-head outer-product sums
// MODE 0:  acc[k,v] += exp(K[n,k]-kmax[k]) * V[n,v];   zsum[k] += exp(K[n,k]-kmax[k])
// MODE 1:  acc[k,v] += scale * Q[n,k] * (g[v] * dY[n,v])                                  (dA of App. E)
// grid = (token chunks, head groups, B); block handles HPB heads; outputs HPB*CH*CH spread over 256 threads.
template <int CH, int HPB, int MODE>
__global__ void __launch_bounds__(256) attn_outer_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                          const float* __restrict__ gate, const float* __restrict__ kmax,
                                                          float* __restrict__ acc_out, float* __restrict__ zsum, float scale, int N,
                                                          int C, int rows_per_block) {
    constexpr int T = 32;
    constexpr int W = HPB * CH;                     // channels handled by this block
    constexpr int NOUT = HPB * CH * CH;
    constexpr int NACC = (NOUT + 255) / 256;
    __shared__ float sp[T][W];
    __shared__ float sr[T][W];
    const int b = blockIdx.z, h0 = blockIdx.y * HPB, c0 = h0 * CH;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    float acc[NACC];
    int pofs[NACC], rofs[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        acc[i] = 0.f;
        const int o = min((int)threadIdx.x + 256 * i, NOUT - 1);   // surplus slots alias the last output (never written back)
        const int hh = o / (CH * CH), k = (o / CH) % CH, v = o % CH;
        pofs[i] = hh * CH + k;
        rofs[i] = hh * CH + v;
    }
    float zacc = 0.f;
    const bf16* base = qkv + (size_t)b * N * 3 * C;
    for (int t0 = r0; t0 < r1; t0 += T) {
        const int tn = min(T, r1 - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < T * W; e += 256) {
            const int t = e / W, cc = e % W;
            float pv = 0.f, rv = 0.f;
            if (t < tn) {
                const size_t row = (size_t)(t0 + t) * 3 * C;
                if (MODE == 0) {
                    pv = __expf(__bfloat162float(base[row + C + c0 + cc]) - kmax[(size_t)b * C + c0 + cc]);
                    rv = __bfloat162float(base[row + 2 * C + c0 + cc]);
                } else {
                    pv = __bfloat162float(base[row + c0 + cc]);
                    rv = (gate ? gate[(size_t)b * C + c0 + cc] : 1.f) * __bfloat162float(dy[((size_t)b * N + t0 + t) * C + c0 + cc]);
                }
            }
            sp[t][cc] = pv;
            sr[t][cc] = rv;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += sp[t][pofs[i]] * sr[t][rofs[i]];
        }
        if (MODE == 0 && threadIdx.x < W) {
            float z = 0.f;
#pragma unroll 8
            for (int t = 0; t < T; ++t) z += sp[t][threadIdx.x];
            zacc += z;
        }
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const int o = threadIdx.x + 256 * i;
        if (o < NOUT) {
            const int hh = o / (CH * CH), k = (o / CH) % CH, v = o % CH;
            atomicAdd(acc_out + ((size_t)b * C + c0 + hh * CH + k) * CH + v, acc[i] * scale);
        }
    }
    if (MODE == 0 && threadIdx.x < W) atomicAdd(zsum + (size_t)b * C + c0 + threadIdx.x, zacc);
}

// A[b,c,v] /= Z[b,c]  and  At[b,h,v,k] = A[b,h,k,v]   (At serves the backward's row-k access pattern)
__global__ void attn_normalize_kernel(float* __restrict__ A, float* __restrict__ At, const float* __restrict__ zsum, int C, int Ch,
                                      long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int v = (int)(i % Ch);
    const long long bc = i / Ch;       // b*C + c
    const int c = (int)(bc % C);
    const int k = c % Ch;
    const float a = A[i] / zsum[bc];
    A[i] = a;
    At[(bc - k + v) * Ch + k] = a;
}

// ---------------------------------------------------------------------------------- phase 2: output
__device__ __forceinline__ float crpe_conv(const bf16* __restrict__ vbase /* V[b, 0, c] */, int ld, int y, int x, int H, int Wd,
                                           const float* __restrict__ w, int win) {
    const int r = win >> 1;
    float e = 0.f;
    for (int i = 0; i < win; ++i) {
        const int yy = y + i - r;
        if (yy < 0 || yy >= H) continue;
        for (int j = 0; j < win; ++j) {
            const int xx = x + j - r;
            if (xx < 0 || xx >= Wd) continue;
            e += __ldg(w + i * win + j) * __bfloat162float(vbase[(size_t)(yy * Wd + xx) * ld]);
        }
    }
    return e;
}

template <int CH>
__global__ void __launch_bounds__(256) attn_out_kernel(const bf16* __restrict__ qkv, const float* __restrict__ A,
                                                        const float* __restrict__ gate, CrpeW cw, bf16* __restrict__ out, float scale,
                                                        int B, int H, int Wd, int C) {
    const int N = H * Wd;
    const long long total = (long long)B * N * C;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const long long tok = idx / C;
    const int n = (int)(tok % N), b = (int)(tok / N);
    const int h = c / CH, v = c % CH;
    const bf16* qrow = qkv + (size_t)tok * 3 * C + h * CH;
    const float* Ab = A + ((size_t)b * C + h * CH) * CH + v;
    float fa = 0.f;
#pragma unroll
    for (int k8 = 0; k8 < CH / 8; ++k8) {
        const uint4 qv = *reinterpret_cast<const uint4*>(qrow + k8 * 8);
        const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 f = bf2_to_f2(qq[u]);
            fa += f.x * __ldg(Ab + (k8 * 8 + 2 * u) * CH) + f.y * __ldg(Ab + (k8 * 8 + 2 * u + 1) * CH);
        }
    }
    int grp, cl, win;
    crpe_lookup(c, CH, grp, cl, win);
    const float e = crpe_conv(qkv + (size_t)b * N * 3 * C + 2 * C + c, 3 * C, n / Wd, n % Wd, H, Wd, cw.w[grp] + (size_t)cl * win * win, win) +
                    __ldg(cw.b[grp] + cl);
    const float q = __bfloat162float(qrow[v]);
    const float g = gate ? __ldg(gate + (size_t)b * C + c) : 1.f;
    out[idx] = __float2bfloat16_rn(g * (scale * fa + q * e));
}

// ---------------------------------------------------------------------------------- backward: per-channel reductions
// dE[n,c] = g*dY*Q;  dWconv[c,tap] += dE[n,c] * V[n+tap,c];  dbconv[c] += dE;  dg[b,c] += dY[n,c]*y[n,c]/g[b,c]
// block = 32 channels x 8 token lanes over a token chunk of one image; grid = (C/32, chunks, B).
__global__ void __launch_bounds__(256) attn_bwd_chan_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                             const bf16* __restrict__ yout, const float* __restrict__ gate, CrpeG cg,
                                                             float* __restrict__ dgate, int H, int Wd, int C, int Ch,
                                                             int rows_per_block) {
    extern __shared__ float sh[];  // [8][32][52]
    const int N = H * Wd;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx, b = blockIdx.z;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
    float acc[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) acc[t] = 0.f;
    float accb = 0.f, accg = 0.f;
    int grp = 0, cl = 0, win = 3;
    if (c < C) {
        crpe_lookup(c, Ch, grp, cl, win);
        const int r = win >> 1;
        const float g = gate ? gate[(size_t)b * C + c] : 1.f;
        const bf16* base = qkv + (size_t)b * N * 3 * C;
        for (int n = r0 + ty; n < r1; n += 8) {
            const float d = __bfloat162float(dy[((size_t)b * N + n) * C + c]);
            const float q = __bfloat162float(base[(size_t)n * 3 * C + c]);
            const float de = g * d * q;
            accb += de;
            if (gate) accg += d * __bfloat162float(yout[((size_t)b * N + n) * C + c]);
            const int y = n / Wd, x = n % Wd;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int yy = y + i - r;
                if (i >= win || yy < 0 || yy >= H) continue;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int xx = x + j - r;
                    if (j >= win || xx < 0 || xx >= Wd) continue;
                    acc[i * 7 + j] += de * __bfloat162float(base[(size_t)(yy * Wd + xx) * 3 * C + 2 * C + c]);
                }
            }
        }
        if (gate) accg /= g;
    }
    float* mine = sh + ((size_t)ty * 32 + tx) * 52;
#pragma unroll
    for (int t = 0; t < 49; ++t) mine[t] = acc[t];
    mine[49] = accb;
    mine[50] = accg;
    __syncthreads();
    for (int o = threadIdx.x; o < 32 * 51; o += 256) {
        const int cc = o / 51, t = o % 51;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += sh[((size_t)k * 32 + cc) * 52 + t];
        const int cgl = blockIdx.x * 32 + cc;
        if (cgl >= C) continue;
        int g2, cl2, win2;
        crpe_lookup(cgl, Ch, g2, cl2, win2);
        if (t < 49) {
            const int i = t / 7, j = t % 7;
            if (i < win2 && j < win2) atomicAdd(cg.w[g2] + (size_t)cl2 * win2 * win2 + i * win2 + j, s);
        } else if (t == 49) {
            atomicAdd(cg.b[g2] + cl2, s);
        } else if (dgate) {
            atomicAdd(dgate + (size_t)b * C + cgl, s);
        }
    }
}

// dAt[b,h,v,k] = dA[b,h,k,v];  r[b,c=(h,k)] = sum_v A[b,c,v] * dA[b,c,v]      one block per (b, head)
__global__ void attn_bwd_mid_kernel(const float* __restrict__ A, const float* __restrict__ dA, float* __restrict__ dAt,
                                    float* __restrict__ rk, int C, int Ch) {
    const int b = blockIdx.y, h = blockIdx.x;
    const size_t base = ((size_t)b * C + h * Ch) * Ch;
    for (int o = threadIdx.x; o < Ch * Ch; o += blockDim.x) {
        const int k = o / Ch, v = o % Ch;
        dAt[base + (size_t)v * Ch + k] = dA[base + o];
    }
    for (int k = threadIdx.x; k < Ch; k += blockDim.x) {
        float s = 0.f;
        for (int v = 0; v < Ch; ++v) s += A[base + (size_t)k * Ch + v] * dA[base + (size_t)k * Ch + v];
        rk[(size_t)b * C + h * Ch + k] = s;
    }
}

// ---------------------------------------------------------------------------------- backward: dQ, dK, dV per (token, channel)
template <int CH>
__global__ void __launch_bounds__(256) attn_bwd_qkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dy,
                                                            const float* __restrict__ gate, const float* __restrict__ At,
                                                            const float* __restrict__ dA, const float* __restrict__ dAt,
                                                            const float* __restrict__ rk, const float* __restrict__ kmax,
                                                            const float* __restrict__ zsum, CrpeW cw, bf16* __restrict__ dqkv,
                                                            float scale, int B, int H, int Wd, int C) {
    const int N = H * Wd;
    const long long total = (long long)B * N * C;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const long long tok = idx / C;
    const int n = (int)(tok % N), b = (int)(tok / N);
    const int h = c / CH, v = c % CH;          // this thread's channel plays k for dQ/dK and v for dV
    const size_t bc0 = (size_t)b * C + h * CH;
    const bf16* row = qkv + (size_t)tok * 3 * C;
    const bf16* dyrow = dy + (size_t)tok * C + h * CH;
    const float* gb = gate ? gate + bc0 : nullptr;
    // sums over the head dimension
    float sq = 0.f;   // sum_v' dF[v'] * A[k=v][v']      -> dQ
    float sv = 0.f;   // sum_k  S[n,k] * dA[k][v]        -> dV
    float sk = 0.f;   // sum_v' V[n,v'] * dA[k=v][v']    -> dK
    const float* Atb = At + bc0 * CH + v;     // At[v'][k=v] = A[k][v'] : stride CH over v', lanes contiguous in k
    const float* dAb = dA + bc0 * CH + v;     // dA[k][v]               : stride CH over k,  lanes contiguous in v
    const float* dAtb = dAt + bc0 * CH + v;   // dAt[v'][k=v] = dA[k][v']
#pragma unroll 4
    for (int j = 0; j < CH; ++j) {
        const float dF = (gb ? __ldg(gb + j) : 1.f) * __bfloat162float(dyrow[j]);
        sq += dF * __ldg(Atb + (size_t)j * CH);
        const float S = __expf(__bfloat162float(row[C + h * CH + j]) - __ldg(kmax + bc0 + j)) / __ldg(zsum + bc0 + j);
        sv += S * __ldg(dAb + (size_t)j * CH);
        sk += __bfloat162float(row[2 * C + h * CH + j]) * __ldg(dAtb + (size_t)j * CH);
    }
    int grp, cl, win;
    crpe_lookup(c, CH, grp, cl, win);
    const float* w = cw.w[grp] + (size_t)cl * win * win;
    const int y = n / Wd, x = n % Wd, r = win >> 1;
    const bf16* imgb = qkv + (size_t)b * N * 3 * C;
    const bf16* dyb = dy + (size_t)b * N * C;
    const float g = gb ? __ldg(gb + v) : 1.f;
    // E[n,c] (forward conv of V) and the transposed conv of dE
    float e = __ldg(cw.b[grp] + cl), tconv = 0.f;
    for (int i = 0; i < win; ++i) {
        for (int j = 0; j < win; ++j) {
            const float wt = __ldg(w + i * win + j);
            const int yy = y + i - r, xx = x + j - r;
            if (yy >= 0 && yy < H && xx >= 0 && xx < Wd) e += wt * __bfloat162float(imgb[(size_t)(yy * Wd + xx) * 3 * C + 2 * C + c]);
            const int y2 = y - i + r, x2 = x - j + r;   // token n' with n' + (i-r, j-r) == n
            if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < Wd) {
                const size_t n2 = (size_t)(y2 * Wd + x2);
                tconv += wt * g * __bfloat162float(dyb[n2 * C + c]) * __bfloat162float(imgb[n2 * 3 * C + c]);
            }
        }
    }
    const float dFc = g * __bfloat162float(dyrow[v]);
    const float Sc = __expf(__bfloat162float(row[C + c]) - __ldg(kmax + bc0 + v)) / __ldg(zsum + bc0 + v);
    const float dk = Sc * (sk - __ldg(rk + bc0 + v));
    const float dv = sv + tconv;
    bf16* drow = dqkv + (size_t)tok * 3 * C;
    drow[c] = __float2bfloat16_rn(scale * sq + dFc * e);
    drow[C + c] = __float2bfloat16_rn(dk);
    drow[2 * C + c] = __float2bfloat16_rn(dv);
}

// ---------------------------------------------------------------------------------- DA gate
// z = W2 relu(W1[:,dom] + b1) + b2  (one-hot input => a column lookup);  g[h,v] = softmax over heads of z[h*Ch+v].
// One block per sample; label is the general [B, nd] fp32 vector so soft labels also work.
__global__ void da_gate_fwd_kernel(const float* __restrict__ label, const float* __restrict__ w1, const float* __restrict__ b1,
                                   const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ hid_out,
                                   float* __restrict__ gate, int nd, int hid, int C, int heads) {
    extern __shared__ float s[];   // hid + C
    float* sh = s;
    float* sz = s + hid;
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < hid; j += blockDim.x) {
        float a = b1[j];
        for (int d = 0; d < nd; ++d) a += w1[j * nd + d] * label[b * nd + d];
        a = fmaxf(a, 0.f);
        sh[j] = a;
        hid_out[(size_t)b * hid + j] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = b2[c];
        const float* wr = w2 + (size_t)c * hid;
        for (int j = 0; j < hid; ++j) a += wr[j] * sh[j];
        sz[c] = a;
    }
    __syncthreads();
    const int Ch = C / heads;
    for (int v = threadIdx.x; v < Ch; v += blockDim.x) {
        float m = -INFINITY;
        for (int h = 0; h < heads; ++h) m = fmaxf(m, sz[h * Ch + v]);
        float z = 0.f;
        for (int h = 0; h < heads; ++h) z += __expf(sz[h * Ch + v] - m);
        for (int h = 0; h < heads; ++h) gate[(size_t)b * C + h * Ch + v] = __expf(sz[h * Ch + v] - m) / z;
    }
}

// dz = g * (dg - sum_h g*dg);  dW2 += dz hid^T; db2 += dz; dhid = W2^T dz * (hid>0); dW1 += dhid label^T; db1 += dhid
__global__ void da_gate_bwd_kernel(const float* __restrict__ label, const float* __restrict__ w2, const float* __restrict__ hid_in,
                                   const float* __restrict__ gate, const float* __restrict__ dgate, float* __restrict__ dw1,
                                   float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, int nd, int hid, int C,
                                   int heads) {
    extern __shared__ float s[];   // dz[C] + hid[hid]
    float* dz = s;
    float* sh = s + C;
    const int b = blockIdx.x;
    const int Ch = C / heads;
    for (int j = threadIdx.x; j < hid; j += blockDim.x) sh[j] = hid_in[(size_t)b * hid + j];
    for (int v = threadIdx.x; v < Ch; v += blockDim.x) {
        float dot = 0.f;
        for (int h = 0; h < heads; ++h) dot += gate[(size_t)b * C + h * Ch + v] * dgate[(size_t)b * C + h * Ch + v];
        for (int h = 0; h < heads; ++h) {
            const size_t i = (size_t)b * C + h * Ch + v;
            dz[h * Ch + v] = gate[i] * (dgate[i] - dot);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(db2 + c, dz[c]);
        for (int j = 0; j < hid; ++j) atomicAdd(dw2 + (size_t)c * hid + j, dz[c] * sh[j]);
    }
    for (int j = threadIdx.x; j < hid; j += blockDim.x) {
        if (sh[j] <= 0.f) continue;
        float a = 0.f;
        for (int c = 0; c < C; ++c) a += w2[(size_t)c * hid + j] * dz[c];
        atomicAdd(db1 + j, a);
        for (int d = 0; d < nd; ++d) atomicAdd(dw1 + j * nd + d, a * label[b * nd + d]);
    }
}

template <int MODE>
int launch_outer(int Ch, const bf16* qkv, const bf16* dy, const float* gate, const float* kmax, float* acc, float* zsum, float scale,
                 int B, int N, int C, cudaStream_t st) {
    const int heads = C / Ch;
    int rpb = mdv_cdiv((long long)N * B * (Ch >= 40 ? heads : 1), 4 * MDV_NUM_SMS);
    rpb = ((rpb + 31) / 32) * 32;
    if (rpb < 32) rpb = 32;
    if (rpb > N) rpb = ((N + 31) / 32) * 32;
    const int chunks = mdv_cdiv(N, rpb);
    switch (Ch) {
        case 8: attn_outer_kernel<8, 8, MODE><<<dim3(chunks, heads / 8, B), 256, 0, st>>>(qkv, dy, gate, kmax, acc, zsum, scale, N, C, rpb); break;
        case 16: attn_outer_kernel<16, 8, MODE><<<dim3(chunks, heads / 8, B), 256, 0, st>>>(qkv, dy, gate, kmax, acc, zsum, scale, N, C, rpb); break;
        case 40: attn_outer_kernel<40, 1, MODE><<<dim3(chunks, heads, B), 256, 0, st>>>(qkv, dy, gate, kmax, acc, zsum, scale, N, C, rpb); break;
        case 64: attn_outer_kernel<64, 1, MODE><<<dim3(chunks, heads, B), 256, 0, st>>>(qkv, dy, gate, kmax, acc, zsum, scale, N, C, rpb); break;
        default: return MDV_ERR_UNSUPPORTED;
    }
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

}  // namespace

// Workspace layout (fp32, caller-owned, kept for backward): kmax[B*C] zsum[B*C] A[B*C*Ch] At[B*C*Ch]
extern "C" long long mdv_attn_stats_floats(int B, int C, int heads) { return (long long)B * C * (2 + 2 * (C / heads)); }

extern "C" int mdv_attn_fwd(const void* qkv_bf16, const float* gate, const float* crpe_w3, const float* crpe_b3,
                            const float* crpe_w5, const float* crpe_b5, const float* crpe_w7, const float* crpe_b7, float* stats,
                            void* out_bf16, int B, int H, int W, int C, int heads, void* stream) {
    if (!qkv_bf16 || !stats || !out_bf16 || heads != 8 || (C % 64)) return MDV_ERR_ARG;
    const int Ch = C / heads, N = H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const bf16* qkv = (const bf16*)qkv_bf16;
    float* kmax = stats;
    float* zsum = kmax + (size_t)B * C;
    float* A = zsum + (size_t)B * C;
    float* At = A + (size_t)B * C * Ch;
    cudaError_t e = cudaMemsetAsync(kmax, 0xFF, sizeof(float) * B * C, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(zsum, 0, sizeof(float) * ((size_t)B * C + (size_t)B * C * Ch), st);
    if (e != cudaSuccess) return (int)e;
    {
        const int half = C / 2;
        const int nty = 256 / half > 0 ? 256 / half : 1;
        int rpb = mdv_cdiv((long long)N * B, 4 * MDV_NUM_SMS);
        if (rpb < 8 * nty) rpb = 8 * nty;
        attn_colmax_kernel<<<dim3(mdv_cdiv(N, rpb), B), half * nty, 0, st>>>(qkv, kmax, N, C, rpb);
        MDV_CHECK_LAUNCH();
    }
    int rc = launch_outer<0>(Ch, qkv, nullptr, nullptr, kmax, A, zsum, 1.0f, B, N, C, st);
    if (rc) return rc;
    const long long tot = (long long)B * C * Ch;
    attn_normalize_kernel<<<mdv_cdiv(tot, 256), 256, 0, st>>>(A, At, zsum, C, Ch, tot);
    MDV_CHECK_LAUNCH();
    CrpeW cw = {{crpe_w3, crpe_w5, crpe_w7}, {crpe_b3, crpe_b5, crpe_b7}};
    const float scale = 1.0f / sqrtf((float)Ch);
    const long long total = (long long)B * N * C;
    const int blocks = mdv_cdiv(total, 256);
    bf16* out = (bf16*)out_bf16;
    switch (Ch) {
        case 8: attn_out_kernel<8><<<blocks, 256, 0, st>>>(qkv, A, gate, cw, out, scale, B, H, W, C); break;
        case 16: attn_out_kernel<16><<<blocks, 256, 0, st>>>(qkv, A, gate, cw, out, scale, B, H, W, C); break;
        case 40: attn_out_kernel<40><<<blocks, 256, 0, st>>>(qkv, A, gate, cw, out, scale, B, H, W, C); break;
        case 64: attn_out_kernel<64><<<blocks, 256, 0, st>>>(qkv, A, gate, cw, out, scale, B, H, W, C); break;
        default: return MDV_ERR_UNSUPPORTED;
    }
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

// ws: fp32 scratch of B*C*(2*Ch+1) floats.  dqkv bf16 [B,N,3C] is overwritten; crpe grads, dgate accumulate (+=).
extern "C" int mdv_attn_bwd(const void* qkv_bf16, const void* dy_bf16, const void* y_bf16, const float* gate, const float* crpe_w3,
                            const float* crpe_b3, const float* crpe_w5, const float* crpe_b5, const float* crpe_w7,
                            const float* crpe_b7, const float* stats, void* dqkv_bf16, float* dgate, float* dcrpe_w3,
                            float* dcrpe_b3, float* dcrpe_w5, float* dcrpe_b5, float* dcrpe_w7, float* dcrpe_b7, float* ws, int B,
                            int H, int W, int C, int heads, void* stream) {
    if (!qkv_bf16 || !dy_bf16 || !stats || !dqkv_bf16 || !ws || heads != 8 || (C % 64)) return MDV_ERR_ARG;
    if (gate && (!y_bf16 || !dgate)) return MDV_ERR_ARG;
    const int Ch = C / heads, N = H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const bf16* qkv = (const bf16*)qkv_bf16;
    const bf16* dy = (const bf16*)dy_bf16;
    const float* kmax = stats;
    const float* zsum = kmax + (size_t)B * C;
    const float* A = zsum + (size_t)B * C;
    const float* At = A + (size_t)B * C * Ch;
    float* dA = ws;
    float* dAt = dA + (size_t)B * C * Ch;
    float* rk = dAt + (size_t)B * C * Ch;
    const float scale = 1.0f / sqrtf((float)Ch);
    cudaError_t e = cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)B * C * Ch, st);
    if (e != cudaSuccess) return (int)e;
    int rc = launch_outer<1>(Ch, qkv, dy, gate, nullptr, dA, nullptr, scale, B, N, C, st);
    if (rc) return rc;
    attn_bwd_mid_kernel<<<dim3(heads, B), 256, 0, st>>>(A, dA, dAt, rk, C, Ch);
    MDV_CHECK_LAUNCH();
    {
        static bool attr_set = false;
        const int smem = 8 * 32 * 52 * (int)sizeof(float);
        if (!attr_set) {
            e = cudaFuncSetAttribute(attn_bwd_chan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        int rpb = mdv_cdiv((long long)N * B * (C / 32), 6 * MDV_NUM_SMS);
        if (rpb < 64) rpb = 64;
        CrpeG cg = {{dcrpe_w3, dcrpe_w5, dcrpe_w7}, {dcrpe_b3, dcrpe_b5, dcrpe_b7}};
        attn_bwd_chan_kernel<<<dim3(C / 32, mdv_cdiv(N, rpb), B), 256, smem, st>>>(qkv, dy, (const bf16*)y_bf16, gate, cg, dgate, H, W, C, Ch, rpb);
        MDV_CHECK_LAUNCH();
    }
    CrpeW cw = {{crpe_w3, crpe_w5, crpe_w7}, {crpe_b3, crpe_b5, crpe_b7}};
    const long long total = (long long)B * N * C;
    const int blocks = mdv_cdiv(total, 256);
    bf16* dqkv = (bf16*)dqkv_bf16;
    switch (Ch) {
        case 8: attn_bwd_qkv_kernel<8><<<blocks, 256, 0, st>>>(qkv, dy, gate, At, dA, dAt, rk, kmax, zsum, cw, dqkv, scale, B, H, W, C); break;
        case 16: attn_bwd_qkv_kernel<16><<<blocks, 256, 0, st>>>(qkv, dy, gate, At, dA, dAt, rk, kmax, zsum, cw, dqkv, scale, B, H, W, C); break;
        case 40: attn_bwd_qkv_kernel<40><<<blocks, 256, 0, st>>>(qkv, dy, gate, At, dA, dAt, rk, kmax, zsum, cw, dqkv, scale, B, H, W, C); break;
        case 64: attn_bwd_qkv_kernel<64><<<blocks, 256, 0, st>>>(qkv, dy, gate, At, dA, dAt, rk, kmax, zsum, cw, dqkv, scale, B, H, W, C); break;
        default: return MDV_ERR_UNSUPPORTED;
    }
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_da_gate_fwd(const float* label, const float* w1, const float* b1, const float* w2, const float* b2, float* hid_out,
                               float* gate, int B, int nd, int hid, int C, int heads, void* stream) {
    if (!label || !w1 || !w2 || !hid_out || !gate || C % heads) return MDV_ERR_ARG;
    da_gate_fwd_kernel<<<B, 128, (hid + C) * sizeof(float), (cudaStream_t)stream>>>(label, w1, b1, w2, b2, hid_out, gate, nd, hid, C, heads);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_da_gate_bwd(const float* label, const float* w2, const float* hid_in, const float* gate, const float* dgate,
                               float* dw1, float* db1, float* dw2, float* db2, int B, int nd, int hid, int C, int heads, void* stream) {
    if (!label || !w2 || !hid_in || !gate || !dgate) return MDV_ERR_ARG;
    da_gate_bwd_kernel<<<B, 128, (hid + C) * sizeof(float), (cudaStream_t)stream>>>(label, w2, hid_in, gate, dgate, dw1, db1, dw2, db2, nd, hid, C, heads);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
