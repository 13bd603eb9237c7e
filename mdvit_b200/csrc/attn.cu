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
#include "attn_internal.cuh"

namespace {

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

// ---------------------------------------------------------------------------------- DA gate
// z = W2 relu(W1 label + b1) + b2  (mdvit.py:272-276; a one-hot label makes the first layer a column lookup);
// g[h,v] = softmax over heads of z[h*Ch+v] (mdvit.py:301-303).  label is the general [B, nd] fp32 vector (soft labels work).
// Kernel 1: one warp per (sample, output channel): the W2 row is read coalesced and warp-reduced.
__global__ void __launch_bounds__(256) da_gate_z_kernel(const float* __restrict__ label, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ hid_out,
                                                         float* __restrict__ z_out, int nd, int hid, int C) {
    extern __shared__ float sh[];   // hid
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < hid; j += blockDim.x) {
        float a = b1[j];
        for (int d = 0; d < nd; ++d) a += w1[j * nd + d] * label[b * nd + d];
        a = fmaxf(a, 0.f);
        sh[j] = a;
        if (blockIdx.x == 0) hid_out[(size_t)b * hid + j] = a;
    }
    __syncthreads();
    const int c = blockIdx.x * 8 + warp;
    if (c >= C) return;
    const float* wr = w2 + (size_t)c * hid;
    float a = 0.f;
    for (int j = lane; j < hid; j += 32) a += __ldg(wr + j) * sh[j];
    a = warp_sum(a);
    if (lane == 0) z_out[(size_t)b * C + c] = a + b2[c];
}
// Kernel 2: softmax over the heads, one thread per (sample, v); in place on the z buffer.
__global__ void da_gate_softmax_kernel(float* __restrict__ gate, int B, int C, int heads) {
    const int Ch = C / heads;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Ch) return;
    const int b = i / Ch, v = i % Ch;
    float* g = gate + (size_t)b * C + v;
    float m = -INFINITY;
    for (int h = 0; h < heads; ++h) m = fmaxf(m, g[h * Ch]);
    float z = 0.f;
    for (int h = 0; h < heads; ++h) z += __expf(g[h * Ch] - m);
    const float inv = 1.f / z;
    for (int h = 0; h < heads; ++h) g[h * Ch] = __expf(g[h * Ch] - m) * inv;
}

// DA gate backward, step 1a (thread per (sample, v)): dz = g * (dg - sum_h g*dg)
__global__ void da_gate_bwd_dz_kernel(const float* __restrict__ gate, const float* __restrict__ dgate, float* __restrict__ dz_out,
                                      int B, int C, int heads) {
    const int Ch = C / heads;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Ch) return;
    const int b = i / Ch, v = i % Ch;
    const size_t o = (size_t)b * C + v;
    float dot = 0.f;
    for (int h = 0; h < heads; ++h) dot += gate[o + h * Ch] * dgate[o + h * Ch];
    for (int h = 0; h < heads; ++h) dz_out[o + h * Ch] = gate[o + h * Ch] * (dgate[o + h * Ch] - dot);
}
// step 1b: dhid[b,j] = (hid > 0) * sum_c W2[c,j] dz[b,c].  block = 32 hidden units x 8 channel lanes, grid (hid/32, B).
__global__ void __launch_bounds__(256) da_gate_bwd_dhid_kernel(const float* __restrict__ w2, const float* __restrict__ hid_in,
                                                                const float* __restrict__ dz, float* __restrict__ dhid_out, int hid,
                                                                int C) {
    __shared__ float red[8][33];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float a = 0.f;
    if (j < hid)
        for (int c = warp; c < C; c += 8) a += __ldg(w2 + (size_t)c * hid + j) * __ldg(dz + (size_t)b * C + c);
    red[warp][lane] = a;
    __syncthreads();
    if (warp == 0 && j < hid) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][lane];
        dhid_out[(size_t)b * hid + j] = hid_in[(size_t)b * hid + j] > 0.f ? t : 0.f;
    }
}
// step 2: dW2[c,j] += sum_b dz[b,c] hid[b,j]; db2[c] += sum_b dz[b,c]; dW1[j,d] += sum_b dhid[b,j] label[b,d]; db1[j] += sum_b dhid[b,j]
__global__ void da_gate_bwd2_kernel(const float* __restrict__ label, const float* __restrict__ hid_in, const float* __restrict__ dz,
                                    const float* __restrict__ dhid, float* __restrict__ dw1, float* __restrict__ db1,
                                    float* __restrict__ dw2, float* __restrict__ db2, int B, int nd, int hid, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n2 = C * hid;
    if (i < n2) {
        const int c = i / hid, j = i % hid;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dz[(size_t)b * C + c] * hid_in[(size_t)b * hid + j];
        dw2[i] += a;
    } else if (i < n2 + C) {
        const int c = i - n2;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dz[(size_t)b * C + c];
        db2[c] += a;
    } else if (i < n2 + C + hid * nd) {
        const int k = i - n2 - C, j = k / nd, d = k % nd;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dhid[(size_t)b * hid + j] * label[b * nd + d];
        dw1[k] += a;
    } else if (i < n2 + C + hid * nd + hid) {
        const int j = i - n2 - C - hid * nd;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dhid[(size_t)b * hid + j];
        db1[j] += a;
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
    return attn_tile_fwd(qkv, A, gate, cw, (bf16*)out_bf16, scale, B, H, W, C, Ch, st);
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
    CrpeG cg = {{dcrpe_w3, dcrpe_w5, dcrpe_w7}, {dcrpe_b3, dcrpe_b5, dcrpe_b7}};
    CrpeW cw = {{crpe_w3, crpe_w5, crpe_w7}, {crpe_b3, crpe_b5, crpe_b7}};
    return attn_tile_bwd(qkv, dy, (const bf16*)y_bf16, gate, A, At, dA, dAt, rk, kmax, zsum, cw, cg, (bf16*)dqkv_bf16, dgate, scale, B, H, W,
                         C, Ch, st);
}

extern "C" int mdv_da_gate_fwd(const float* label, const float* w1, const float* b1, const float* w2, const float* b2, float* hid_out,
                               float* gate, int B, int nd, int hid, int C, int heads, void* stream) {
    if (!label || !w1 || !w2 || !hid_out || !gate || C % heads) return MDV_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    da_gate_z_kernel<<<dim3(mdv_cdiv(C, 8), B), 256, hid * sizeof(float), st>>>(label, w1, b1, w2, b2, hid_out, gate, nd, hid, C);
    MDV_CHECK_LAUNCH();
    da_gate_softmax_kernel<<<mdv_cdiv(B * (C / heads), 128), 128, 0, st>>>(gate, B, C, heads);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

// ws: B*(C+hid) floats of scratch.  Gradients accumulate (+=).
extern "C" int mdv_da_gate_bwd(const float* label, const float* w2, const float* hid_in, const float* gate, const float* dgate,
                               float* dw1, float* db1, float* dw2, float* db2, float* ws, int B, int nd, int hid, int C, int heads,
                               void* stream) {
    if (!label || !w2 || !hid_in || !gate || !dgate || !ws || !dw1 || !db1 || !dw2 || !db2) return MDV_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float* dz = ws;
    float* dhid = ws + (size_t)B * C;
    da_gate_bwd_dz_kernel<<<mdv_cdiv(B * (C / heads), 128), 128, 0, st>>>(gate, dgate, dz, B, C, heads);
    MDV_CHECK_LAUNCH();
    da_gate_bwd_dhid_kernel<<<dim3(mdv_cdiv(hid, 32), B), 256, 0, st>>>(w2, hid_in, dz, dhid, hid, C);
    MDV_CHECK_LAUNCH();
    const int total = C * hid + C + hid * nd + hid;
    da_gate_bwd2_kernel<<<mdv_cdiv(total, 256), 256, 0, st>>>(label, hid_in, dz, dhid, dw1, db1, dw2, db2, B, nd, hid, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
