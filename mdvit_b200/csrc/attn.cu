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

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// ---------------------------------------------------------------------------------- phase 1a: column max of K
// kmax must be pre-filled with 0xFF bytes (identity of the atomic above).  thread = 2 channels, ty = row lane.
__global__ void __launch_bounds__(256) attn_colmax_kernel(const bf16* __restrict__ qkv, float* __restrict__ kmax, int N, int C,
                                                           int rows_per_block) {
    MDV_PDL_SYNC();
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

// ---------------------------------------------------------------------------------- DA gate
// z = W2 relu(W1 label + b1) + b2  (mdvit.py:272-276; a one-hot label makes the first layer a column lookup);
// g[h,v] = softmax over heads of z[h*Ch+v] (mdvit.py:301-303).  label is the general [B, nd] fp32 vector (soft labels work).
// Kernel 1: one warp per (sample, output channel): the W2 row is read coalesced and warp-reduced.
__global__ void __launch_bounds__(256) da_gate_z_kernel(const float* __restrict__ label, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ hid_out,
                                                         float* __restrict__ z_out, int nd, int hid, int C) {
    MDV_PDL_SYNC();
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
    MDV_PDL_SYNC();
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
    MDV_PDL_SYNC();
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
    MDV_PDL_SYNC();
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
                                    float* __restrict__ dw2, float* __restrict__ db2, int B, int nd, int hid, int C, int bsplit) {
    MDV_PDL_SYNC();
    // blockIdx.y owns a slice of the samples (the per-thread loop over B was the critical path); partial sums are added atomically
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = (B + bsplit - 1) / bsplit;
    const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
    const int n2 = C * hid;
    if (i < n2) {
        const int c = i / hid, j = i % hid;
        float a = 0.f;
        for (int b = b0; b < b1; ++b) a += dz[(size_t)b * C + c] * hid_in[(size_t)b * hid + j];
        atomicAdd(dw2 + i, a);
    } else if (i < n2 + C) {
        const int c = i - n2;
        float a = 0.f;
        for (int b = b0; b < b1; ++b) a += dz[(size_t)b * C + c];
        atomicAdd(db2 + c, a);
    } else if (i < n2 + C + hid * nd) {
        const int k = i - n2 - C, j = k / nd, d = k % nd;
        float a = 0.f;
        for (int b = b0; b < b1; ++b) a += dhid[(size_t)b * hid + j] * label[b * nd + d];
        atomicAdd(dw1 + k, a);
    } else if (i < n2 + C + hid * nd + hid) {
        const int j = i - n2 - C - hid * nd;
        float a = 0.f;
        for (int b = b0; b < b1; ++b) a += dhid[(size_t)b * hid + j];
        atomicAdd(db1 + j, a);
    }
}

}  // namespace

// stats (fp32, caller-owned, kept for backward): kmax[B*C] zsum[B*C] A[B*C*Ch]
extern "C" long long mdv_attn_stats_floats(int B, int C, int heads) { return (long long)B * C * (2 + C / heads); }
// scratch for one forward or backward call (partial sums of the cross-token reductions, dA, r)
extern "C" long long mdv_attn_ws_floats(int B, int C, int heads) { return attn_strip_ws_floats(B, C, C / heads); }

extern "C" int mdv_attn_fwd(const void* qkv_bf16, const float* gate, const float* crpe_w3, const float* crpe_b3,
                            const float* crpe_w5, const float* crpe_b5, const float* crpe_w7, const float* crpe_b7, float* stats,
                            float* ws, void* out_bf16, void* e_out_bf16, int B, int H, int W, int C, int heads, void* stream) {
    if (!qkv_bf16 || !stats || !ws || !out_bf16 || heads != 8 || (C % 64)) return MDV_ERR_ARG;
    const int Ch = C / heads, N = H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const bf16* qkv = (const bf16*)qkv_bf16;
    float* kmax = stats;
    float* zsum = kmax + (size_t)B * C;
    float* A = zsum + (size_t)B * C;
    cudaError_t e = cudaMemsetAsync(kmax, 0xFF, sizeof(float) * B * C, st);
    if (e != cudaSuccess) return (int)e;
    {
        const int half = C / 2;
        const int nty = 256 / half > 0 ? 256 / half : 1;
        int rpb = mdv_cdiv((long long)N * B, 4 * MDV_NUM_SMS);
        if (rpb < 8 * nty) rpb = 8 * nty;
        mdv_launch(attn_colmax_kernel, dim3(dim3(mdv_cdiv(N, rpb), B)), dim3(half * nty), 0, st, qkv, kmax, N, C, rpb);
        MDV_CHECK_LAUNCH();
    }
    CrpeW cw = {{crpe_w3, crpe_w5, crpe_w7}, {crpe_b3, crpe_b5, crpe_b7}};
    const float scale = 1.0f / sqrtf((float)Ch);
    return attn_strip_fwd(qkv, gate, cw, kmax, zsum, A, ws, (bf16*)out_bf16, (bf16*)e_out_bf16, scale, B, H, W, C, Ch, st);
}

// ws: mdv_attn_ws_floats() floats of scratch.  dqkv bf16 [B,N,3C] is overwritten; crpe grads (may all be NULL: skipped)
// and dgate accumulate (+=).
extern "C" int mdv_attn_bwd(const void* qkv_bf16, const void* dy_bf16, const void* y_bf16, const void* e_bf16, const float* gate, const float* crpe_w3,
                            const float* crpe_b3, const float* crpe_w5, const float* crpe_b5, const float* crpe_w7,
                            const float* crpe_b7, const float* stats, void* dqkv_bf16, float* dgate, float* dcrpe_w3,
                            float* dcrpe_b3, float* dcrpe_w5, float* dcrpe_b5, float* dcrpe_w7, float* dcrpe_b7, float* dbias_qkv,
                            float* ws, int B, int H, int W, int C, int heads, void* stream) {
    if (!qkv_bf16 || !dy_bf16 || !e_bf16 || !stats || !dqkv_bf16 || !ws || heads != 8 || (C % 64)) return MDV_ERR_ARG;
    if (gate && (!y_bf16 || !dgate)) return MDV_ERR_ARG;
    const int Ch = C / heads;
    const float* kmax = stats;
    const float* zsum = kmax + (size_t)B * C;
    const float* A = zsum + (size_t)B * C;
    const float scale = 1.0f / sqrtf((float)Ch);
    CrpeG cg = {{dcrpe_w3, dcrpe_w5, dcrpe_w7}, {dcrpe_b3, dcrpe_b5, dcrpe_b7}};
    CrpeW cw = {{crpe_w3, crpe_w5, crpe_w7}, {crpe_b3, crpe_b5, crpe_b7}};
    if (gate) {      // the tiles of an image add their gate sums with atomics: start from zero (a memset node, not a fill kernel)
        cudaError_t e = cudaMemsetAsync(dgate, 0, sizeof(float) * (size_t)B * C, (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    return attn_strip_bwd((const bf16*)qkv_bf16, (const bf16*)dy_bf16, (const bf16*)y_bf16, gate, kmax, zsum, A, ws, cw, cg,
                          (const bf16*)e_bf16, (bf16*)dqkv_bf16, dgate, dbias_qkv, scale, B, H, W, C, Ch, (cudaStream_t)stream);
}

extern "C" int mdv_da_gate_fwd(const float* label, const float* w1, const float* b1, const float* w2, const float* b2, float* hid_out,
                               float* gate, int B, int nd, int hid, int C, int heads, void* stream) {
    if (!label || !w1 || !w2 || !hid_out || !gate || C % heads) return MDV_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    mdv_launch(da_gate_z_kernel, dim3(dim3(mdv_cdiv(C, 8), B)), dim3(256), hid * sizeof(float), st, label, w1, b1, w2, b2, hid_out, gate, nd, hid, C);
    MDV_CHECK_LAUNCH();
    mdv_launch(da_gate_softmax_kernel, dim3(mdv_cdiv(B * (C / heads), 128)), dim3(128), 0, st, gate, B, C, heads);
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
    mdv_launch(da_gate_bwd_dz_kernel, dim3(mdv_cdiv(B * (C / heads), 128)), dim3(128), 0, st, gate, dgate, dz, B, C, heads);
    MDV_CHECK_LAUNCH();
    mdv_launch(da_gate_bwd_dhid_kernel, dim3(dim3(mdv_cdiv(hid, 32), B)), dim3(256), 0, st, w2, hid_in, dz, dhid, hid, C);
    MDV_CHECK_LAUNCH();
    const int total = C * hid + C + hid * nd + hid;
    const int bsplit = B >= 64 ? 8 : (B >= 16 ? 4 : 1);
    mdv_launch(da_gate_bwd2_kernel, dim3(mdv_cdiv(total, 256), bsplit), dim3(256), 0, st, label, hid_in, dz, dhid, dw1, db1, dw2, db2, B, nd, hid, C, bsplit);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
