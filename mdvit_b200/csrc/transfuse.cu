// TransFuse_S_adapt's convolutional side (Models/Hybrid_models/TransFuseFolder/TransFuse.py:182-283): the stencil / resize /
// pooling kernels around the tcgen05 GEMMs that run its dense convolutions.  Activations are NHWC fp32 ([B, H*W, C]).
#include "common.cuh"
#include "../../include/mdvit_b200.h"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// element loops index with 32-bit unsigned arithmetic (64-bit divisions by runtime integers cost ~100 instructions each)
inline bool fits_u32(long long total) { return total > 0 && total < 4000000000LL; }
inline int grid_for(long long total, int threads = 256) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)MDV_NUM_SMS * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

// ---------------------------------------------------------------------------------- generic k x k im2col / col2im
// col[(b,yo,xo), c*k*k + i*k + j] = in[b, yo*s - pad + i, xo*s - pad + j, c]   (0 outside the image and in the columns
// >= C*k*k up to the pitch ldc).  The column order is that of a flattened nn.Conv2d weight [Cout, Cin, k, k], so the GEMM's
// W operand is the parameter itself.  `in` is NHWC, or the NCHW input image (in_nchw = 1: resnet.conv1, TransFuse.py:231).
// KK / CC / LD: kernel size, channels and pitch at compile time (0 = the runtime arguments): the index arithmetic of the generic
// form (five divisions by runtime integers per element) made the 7x7 stem im2col instruction bound at 4x its HBM time.
template <typename TO, int KK, int CC, int LD>
__global__ void __launch_bounds__(256) im2col_k_kernel(const float* __restrict__ in, TO* __restrict__ col, int B, int Hi, int Wi,
                                                        int Ho, int Wo, int C_rt, int k_rt, int stride, int pad, int ldc_rt, int in_nchw) {
    MDV_PDL_SYNC();
    const int C = CC ? CC : C_rt, k = KK ? KK : k_rt, ldc = LD ? LD : ldc_rt;
    const int kk = k * k, K = C * kk;
    const int npix = B * Ho * Wo;
    const int lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
    // one warp per output pixel (its coordinates are decoded once), lanes over the columns: coalesced stores
    for (int pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < npix; pix += nwarp) {
        const int xo = pix % Wo, yo = (pix / Wo) % Ho, b = pix / (Wo * Ho);
        const int y0 = yo * stride - pad, x0 = xo * stride - pad;
        TO* dst = col + (size_t)pix * ldc;
        for (int q = lane; q < ldc; q += 32) {
            float v = 0.f;
            if (q < K) {
                const int c = q / kk, t = q - c * kk;
                const int i = t / k;
                const int yi = y0 + i, xi = x0 + (t - i * k);
                if (yi >= 0 && yi < Hi && xi >= 0 && xi < Wi)
                    v = in_nchw ? __ldg(in + (((size_t)b * C + c) * Hi + yi) * Wi + xi) : __ldg(in + (((size_t)b * Hi + yi) * Wi + xi) * C + c);
            }
            stf(dst + q, v);
        }
    }
}

// dx[b,y,x,c] = sum_ij dcol[(b,(y+pad-i)/s,(x+pad-j)/s), c*k*k + i*k + j]   (gather form of the transpose; NHWC dx)
__global__ void __launch_bounds__(256) col2im_k_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int B, int Hi, int Wi,
                                                        int Ho, int Wo, int C, int k, int stride, int pad, int ldc) {
    MDV_PDL_SYNC();
    const long long total = (long long)B * Hi * Wi * C;
    const int kk = k * k;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const unsigned pix = idx / C;
        const int x = (int)(pix % Wi);
        const int y = (int)((pix / Wi) % Hi);
        const int b = (int)(pix / (unsigned)(Wi * Hi));
        float a = 0.f;
        for (int i = 0; i < k; ++i) {
            const int t = y + pad - i;
            if (t < 0 || (t % stride)) continue;
            const int yo = t / stride;
            if (yo >= Ho) continue;
            for (int j = 0; j < k; ++j) {
                const int u = x + pad - j;
                if (u < 0 || (u % stride)) continue;
                const int xo = u / stride;
                if (xo >= Wo) continue;
                a += __ldg(dcol + (((size_t)b * Ho + yo) * Wo + xo) * ldc + c * kk + i * k + j);
            }
        }
        dx[idx] = a;
    }
}

// ---------------------------------------------------------------------------------- MaxPool2d(3, stride 2, padding 1)
// torchvision resnet34.maxpool (TransFuse.py:234).  idx keeps the tap (i*3+j) of the first maximum in scan order, which is
// the element nn.MaxPool2d routes the gradient to.
__global__ void __launch_bounds__(256) maxpool3s2_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              unsigned char* __restrict__ tap, int B, int Hi, int Wi, int Ho, int Wo, int C) {
    MDV_PDL_SYNC();
    const int c4n = C >> 2;
    const long long total = (long long)B * Ho * Wo * c4n;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = (int)(idx % c4n) * 4;
        const unsigned pix = idx / c4n;
        const int xo = (int)(pix % Wo);
        const int yo = (int)((pix / Wo) % Ho);
        const int b = (int)(pix / (unsigned)(Wo * Ho));
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int am[4] = {0, 0, 0, 0};
        bool first = true;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int yi = yo * 2 - 1 + i;
            if (yi < 0 || yi >= Hi) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int xi = xo * 2 - 1 + j;
                if (xi < 0 || xi >= Wi) continue;
                const float4 v = ld4(in + (((size_t)b * Hi + yi) * Wi + xi) * C + c);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (first || vv[e] > m[e]) { m[e] = vv[e]; am[e] = i * 3 + j; }
                first = false;
            }
        }
        st4(out + (size_t)pix * C + c, make_float4(m[0], m[1], m[2], m[3]));
        *reinterpret_cast<uchar4*>(tap + (size_t)pix * C + c) = make_uchar4(am[0], am[1], am[2], am[3]);
    }
}

__global__ void __launch_bounds__(256) maxpool3s2_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ tap,
                                                              float* __restrict__ din, int B, int Hi, int Wi, int Ho, int Wo, int C) {
    MDV_PDL_SYNC();
    const int c4n = C >> 2;
    const long long total = (long long)B * Hi * Wi * c4n;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = (int)(idx % c4n) * 4;
        const unsigned pix = idx / c4n;
        const int x = (int)(pix % Wi);
        const int y = (int)((pix / Wi) % Hi);
        const int b = (int)(pix / (unsigned)(Wi * Hi));
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int t = y + 1 - i;
            if (t < 0 || (t & 1)) continue;
            const int yo = t >> 1;
            if (yo >= Ho) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int u = x + 1 - j;
                if (u < 0 || (u & 1)) continue;
                const int xo = u >> 1;
                if (xo >= Wo) continue;
                const size_t o = (((size_t)b * Ho + yo) * Wo + xo) * C + c;
                const uchar4 tp = *reinterpret_cast<const uchar4*>(tap + o);
                const float4 d = ld4(dout + o);
                const int me = i * 3 + j;
                if (tp.x == me) a[0] += d.x;
                if (tp.y == me) a[1] += d.y;
                if (tp.z == me) a[2] += d.z;
                if (tp.w == me) a[3] += d.w;
            }
        }
        st4(din + (size_t)pix * C + c, make_float4(a[0], a[1], a[2], a[3]));
    }
}

// ---------------------------------------------------------------------------------- residual add + activation
// out = act(a + b)  (BasicBlock `out += identity; relu`, DoubleConv, Attention_block: TransFuse.py:589,617)
__global__ void __launch_bounds__(256) add_act_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                       long long n4, int act) {
    MDV_PDL_SYNC();
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 x = ld4(a + (size_t)i * 4), y = ld4(b + (size_t)i * 4);
        float4 r = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        if (act == MDV_ACT_RELU) r = make_float4(fmaxf(r.x, 0.f), fmaxf(r.y, 0.f), fmaxf(r.z, 0.f), fmaxf(r.w, 0.f));
        st4(out + (size_t)i * 4, r);
    }
}
// dx = dy * act'(.) evaluated from the activation's OUTPUT y (ReLU: y > 0)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                                       long long n4) {
    MDV_PDL_SYNC();
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 d = ld4(dy + (size_t)i * 4), v = ld4(y + (size_t)i * 4);
        st4(dx + (size_t)i * 4, make_float4(v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f));
    }
}

// ---------------------------------------------------------------------------------- bilinear resize, align_corners=True
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) in Up (TransFuse.py:559) and the three output maps
// (F.interpolate(..., align_corners=True), TransFuse.py:262-264): src = dst * (n_in - 1) / (n_out - 1).
__device__ __forceinline__ void ac_src(int d, float scale, int n_in, int& i0, int& i1, float& lam) {
    const float s = scale * (float)d;
    i0 = (int)s;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = s - (float)i0;
}
// weight with which output index d reads input index i (0 if it does not)
__device__ __forceinline__ float ac_weight(int d, int i, float scale, int n_in) {
    int i0, i1;
    float lam;
    ac_src(d, scale, n_in, i0, i1, lam);
    float w = 0.f;
    if (i0 == i) w += 1.f - lam;
    if (i1 == i) w += lam;      // i1 == i0 at the last index: the two taps add up to 1, as in the forward
    return w;
}

template <int V>      // V = 4: C % 4 == 0 (vector lanes); V = 1: any C
__global__ void __launch_bounds__(256) resize_ac_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Hi, int Wi,
                                                             int Ho, int Wo, int C, float sy, float sx) {
    MDV_PDL_SYNC();
    const int cn = C / V;
    const long long total = (long long)B * Ho * Wo * cn;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = (int)(idx % cn) * V;
        const unsigned pix = idx / cn;
        const int xo = (int)(pix % Wo);
        const int yo = (int)((pix / Wo) % Ho);
        const int b = (int)(pix / (unsigned)(Wo * Ho));
        int y0, y1, x0, x1;
        float ly, lx;
        ac_src(yo, sy, Hi, y0, y1, ly);
        ac_src(xo, sx, Wi, x0, x1, lx);
        const float* p = in + (size_t)b * Hi * Wi * C + c;
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        if (V == 4) {
            const float4 a = ld4(p + ((size_t)y0 * Wi + x0) * C), bq = ld4(p + ((size_t)y0 * Wi + x1) * C);
            const float4 cq = ld4(p + ((size_t)y1 * Wi + x0) * C), d = ld4(p + ((size_t)y1 * Wi + x1) * C);
            st4(out + (size_t)pix * C + c, make_float4(w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x, w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y,
                                                        w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z, w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w));
        } else {
            out[(size_t)pix * C + c] = w00 * __ldg(p + ((size_t)y0 * Wi + x0) * C) + w01 * __ldg(p + ((size_t)y0 * Wi + x1) * C) +
                                       w10 * __ldg(p + ((size_t)y1 * Wi + x0) * C) + w11 * __ldg(p + ((size_t)y1 * Wi + x1) * C);
        }
    }
}

// exact transpose in gather form: din[y,x] = sum over the output pixels that read (y,x), found by scanning the few output
// rows / columns whose source interval can contain it (deterministic, no atomics)
template <int V>
__global__ void __launch_bounds__(256) resize_ac_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B, int Hi, int Wi,
                                                             int Ho, int Wo, int C, float sy, float sx, float ry, float rx) {
    MDV_PDL_SYNC();
    const int cn = C / V;
    const long long total = (long long)B * Hi * Wi * cn;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = (int)(idx % cn) * V;
        const unsigned pix = idx / cn;
        const int x = (int)(pix % Wi);
        const int y = (int)((pix / Wi) % Hi);
        const int b = (int)(pix / (unsigned)(Wi * Hi));
        // outputs d with source in (y-1, y+1): d in ((y-1)*r, (y+1)*r), r = (n_out-1)/(n_in-1); widened by one on each side
        int ylo = (int)floorf((float)(y - 1) * ry) - 1, yhi = (int)ceilf((float)(y + 1) * ry) + 1;
        int xlo = (int)floorf((float)(x - 1) * rx) - 1, xhi = (int)ceilf((float)(x + 1) * rx) + 1;
        ylo = max(ylo, 0); yhi = min(yhi, Ho - 1); xlo = max(xlo, 0); xhi = min(xhi, Wo - 1);
        float a[V];
#pragma unroll
        for (int e = 0; e < V; ++e) a[e] = 0.f;
        const float* p = dout + (size_t)b * Ho * Wo * C + c;
        for (int yo = ylo; yo <= yhi; ++yo) {
            const float wy = ac_weight(yo, y, sy, Hi);
            if (wy == 0.f) continue;
            for (int xo = xlo; xo <= xhi; ++xo) {
                const float wx = ac_weight(xo, x, sx, Wi);
                if (wx == 0.f) continue;
                const float w = wy * wx;
                if (V == 4) {
                    const float4 d = ld4(p + ((size_t)yo * Wo + xo) * C);
                    a[0] += w * d.x; a[1 % V] += w * d.y; a[2 % V] += w * d.z; a[3 % V] += w * d.w;
                } else {
                    a[0] += w * __ldg(p + ((size_t)yo * Wo + xo) * C);
                }
            }
        }
        if (V == 4) st4(din + (size_t)pix * C + c, make_float4(a[0], a[1 % V], a[2 % V], a[3 % V]));
        else din[(size_t)pix * C + c] = a[0];
    }
}

// ---------------------------------------------------------------------------------- gates of BiFusion_block / Attention_block
// out[m, :] = [ g[m, :] * p[m]  |  x[m, :] * v[b(m), :]  |  bp[m, :] ]      (TransFuse.py:63-73: sigmoid(spatial) * g_in, sigmoid(fc2) * x_in,
// torch.cat([g, x, bp], 1); Attention_block `x * psi`, TransFuse.py:620, is the first part alone).  p [M] and v [B, C2] are the
// gates AFTER their sigmoids.  One warp per row, float4 lanes over the concatenated channels.
__global__ void __launch_bounds__(256) gate_cat_fwd_kernel(const float* __restrict__ g, const float* __restrict__ pg, const float* __restrict__ x,
                                                            const float* __restrict__ v, const float* __restrict__ bp, float* __restrict__ out,
                                                            int M, int C1, int C2, int C3, int rows_per_sample) {
    MDV_PDL_SYNC();
    const int Ct = C1 + C2 + C3, n4 = Ct >> 2;
    const int lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += nwarp) {
        const float pm = __ldg(pg + m);
        const float* vb = v ? v + (size_t)(m / rows_per_sample) * C2 : nullptr;
        for (int q = lane; q < n4; q += 32) {
            const int c = q << 2;
            float4 r;
            if (c < C1) {
                r = ld4(g + (size_t)m * C1 + c);
                r.x *= pm; r.y *= pm; r.z *= pm; r.w *= pm;
            } else if (c < C1 + C2) {
                r = ld4(x + (size_t)m * C2 + (c - C1));
                const float4 w = ld4(vb + (c - C1));
                r.x *= w.x; r.y *= w.y; r.z *= w.z; r.w *= w.w;
            } else {
                r = ld4(bp + (size_t)m * C3 + (c - C1 - C2));
            }
            st4(out + (size_t)m * Ct + c, r);
        }
    }
}
// dg = dout1 * p;  dp[m] = sum_c dout1 g;  dx = dout2 * v;  dv[b, c] += sum_{m in b} dout2 x;  dbp = dout3.
// A warp owns a band of consecutive rows of one sample and keeps its dv partial sums in registers (C2 <= 512).
__global__ void __launch_bounds__(256) gate_cat_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ g, const float* __restrict__ pg,
                                                            const float* __restrict__ x, const float* __restrict__ v, float* __restrict__ dg,
                                                            float* __restrict__ dp, float* __restrict__ dx, float* __restrict__ dv,
                                                            float* __restrict__ dbp, int M, int C1, int C2, int C3, int rows_per_sample, int band) {
    MDV_PDL_SYNC();
    const int Ct = C1 + C2 + C3;
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int bands_per_sample = (rows_per_sample + band - 1) / band;
    const int b = wid / bands_per_sample;
    const int r0 = b * rows_per_sample + (wid - b * bands_per_sample) * band;
    if (r0 >= M) return;
    const int r1 = min(min(r0 + band, (b + 1) * rows_per_sample), M);
    float4 acc[4];      // dv partial sums of this lane's chunks (C2 / 4 / 32 <= 4)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* vb = v ? v + (size_t)b * C2 : nullptr;
    for (int m = r0; m < r1; ++m) {
        const float pm = __ldg(pg + m);
        const float* d = dout + (size_t)m * Ct;
        float dot = 0.f;
        for (int c = lane << 2; c < C1; c += 128) {
            const float4 dd = ld4(d + c), gg = ld4(g + (size_t)m * C1 + c);
            dot += dd.x * gg.x + dd.y * gg.y + dd.z * gg.z + dd.w * gg.w;
            st4(dg + (size_t)m * C1 + c, make_float4(dd.x * pm, dd.y * pm, dd.z * pm, dd.w * pm));
        }
        dot = warp_sum(dot);
        if (lane == 0) dp[m] = dot;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (lane << 2) + i * 128;
            if (c < C2) {
                const float4 dd = ld4(d + C1 + c), xx = ld4(x + (size_t)m * C2 + c), w = ld4(vb + c);
                acc[i].x += dd.x * xx.x; acc[i].y += dd.y * xx.y; acc[i].z += dd.z * xx.z; acc[i].w += dd.w * xx.w;
                st4(dx + (size_t)m * C2 + c, make_float4(dd.x * w.x, dd.y * w.y, dd.z * w.z, dd.w * w.w));
            }
        }
        for (int c = lane << 2; c < C3; c += 128) st4(dbp + (size_t)m * C3 + c, ld4(d + C1 + C2 + c));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = (lane << 2) + i * 128;
        if (c < C2) {
            atomicAdd(dv + (size_t)b * C2 + c, acc[i].x);
            atomicAdd(dv + (size_t)b * C2 + c + 1, acc[i].y);
            atomicAdd(dv + (size_t)b * C2 + c + 2, acc[i].z);
            atomicAdd(dv + (size_t)b * C2 + c + 3, acc[i].w);
        }
    }
}

// ChannelPool (TransFuse.py:20-22): out[m] = (max_c x[m, c], mean_c x[m, c]); arg[m] = first maximal channel (where torch.max sends
// the gradient).  One warp per row.
__global__ void __launch_bounds__(256) channel_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int* __restrict__ arg, int M, int C) {
    MDV_PDL_SYNC();
    const int lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += nwarp) {
        float mx = -INFINITY, sum = 0.f;
        int am = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float t = __ldg(x + (size_t)m * C + c);
            sum += t;
            if (t > mx) { mx = t; am = c; }
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
        }
        if (lane == 0) {
            out[(size_t)m * 2] = mx;
            out[(size_t)m * 2 + 1] = sum / (float)C;
            arg[m] = am;
        }
    }
}
// dx[m, c] = dout[m, 1] / C + (c == arg[m]) dout[m, 0]
__global__ void __launch_bounds__(256) channel_pool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, float* __restrict__ dx,
                                                                int M, int C) {
    MDV_PDL_SYNC();
    const unsigned total = (unsigned)M * (unsigned)C;
    const float inv = 1.f / (float)C;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned m = idx / (unsigned)C;
        const int c = (int)(idx - m * (unsigned)C);
        dx[idx] = __ldg(dout + (size_t)m * 2 + 1) * inv + (c == __ldg(arg + m) ? __ldg(dout + (size_t)m * 2) : 0.f);
    }
}

// nn.Dropout2d on an NHWC map: whole (sample, channel) planes, mask = f(rng, stream, b*C + c) as in the rowdot kernels; the
// backward is the same call on the gradient.
__global__ void __launch_bounds__(256) dropout2d_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int C, int rows_per_sample,
                                                         float p, const unsigned long long* __restrict__ rng, uint32_t stream) {
    MDV_PDL_SYNC();
    const uint32_t thr = drop_thresh(p), key = rng_key(rng, stream);
    const float inv = 1.f / (1.f - p);
    const int c4n = C >> 2;
    const unsigned total = (unsigned)M * (unsigned)c4n;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned m = idx / (unsigned)c4n;
        const int c = (int)(idx - m * (unsigned)c4n) << 2;
        const unsigned long long e = (unsigned long long)(m / (unsigned)rows_per_sample) * C + c;
        const uint32_t h0 = drop_hash(key, (uint32_t)(e >> 1)), h1 = drop_hash(key, (uint32_t)(e >> 1) + 1);      // c % 4 == 0: pairs (c, c+1), (c+2, c+3)
        float4 v = ld4(x + (size_t)m * C + c);
        v.x *= drop_lo(h0, thr, inv); v.y *= drop_hi(h0, thr, inv); v.z *= drop_lo(h1, thr, inv); v.w *= drop_hi(h1, thr, inv);
        st4(out + (size_t)m * C + c, v);
    }
}

// ---------------------------------------------------------------------------------- structure_loss (multi_train_TransFuse.py:29-38)
// weit = 1 + 5 |avg_pool2d(mask, 31, stride 1, padding 15) - mask|   (count_include_pad: always / 961).
// Separable running sums: one block per (sample, row strip); rows first into shared memory, then columns.
__global__ void __launch_bounds__(256) boxsum_rows_kernel(const float* __restrict__ mask, float* __restrict__ tmp, int B, int H, int W, int R) {
    MDV_PDL_SYNC();
    const long long total = (long long)B * H * W;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const float* row = mask + ((size_t)idx - x);
        float s = 0.f;
        const int lo = max(x - R, 0), hi = min(x + R, W - 1);
        for (int u = lo; u <= hi; ++u) s += __ldg(row + u);
        tmp[idx] = s;
    }
}
__global__ void __launch_bounds__(256) boxsum_cols_weit_kernel(const float* __restrict__ tmp, const float* __restrict__ mask,
                                                                float* __restrict__ weit, int B, int H, int W, int R) {
    MDV_PDL_SYNC();
    const long long total = (long long)B * H * W;
    const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const int y = (int)((idx / W) % H);
        const float* img = tmp + ((size_t)idx - (size_t)y * W - x);
        float s = 0.f;
        const int lo = max(y - R, 0), hi = min(y + R, H - 1);
        for (int v = lo; v <= hi; ++v) s += __ldg(img + (size_t)v * W + x);
        weit[idx] = 1.f + 5.f * fabsf(s * inv - __ldg(mask + idx));
    }
}

// per-sample sums {sum weit, sum weit*bce, sum p*m*weit, sum (p+m)*weit} as doubles: sums[b*4 + .] +=
__global__ void __launch_bounds__(256) structure_sums_kernel(const float* __restrict__ pred, const float* __restrict__ mask,
                                                              const float* __restrict__ weit, double* __restrict__ sums, int HW, int blocks_per_sample) {
    MDV_PDL_SYNC();
    __shared__ float red[32];
    const int b = blockIdx.x / blocks_per_sample, part = blockIdx.x % blocks_per_sample;
    const size_t base = (size_t)b * HW;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int i = part * blockDim.x + threadIdx.x; i < HW; i += blocks_per_sample * blockDim.x) {
        const float z = __ldg(pred + base + i), m = __ldg(mask + base + i), w = __ldg(weit + base + i);
        // binary_cross_entropy_with_logits: max(z,0) - z m + log(1 + exp(-|z|))
        const float bce = fmaxf(z, 0.f) - z * m + log1pf(__expf(-fabsf(z)));
        const float p = 1.f / (1.f + __expf(-z));
        s0 += w; s1 += w * bce; s2 += p * m * w; s3 += (p + m) * w;
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red); s2 = block_sum(s2, red); s3 = block_sum(s3, red);
    if (threadIdx.x == 0) {
        atomicAdd(sums + b * 4 + 0, (double)s0);
        atomicAdd(sums + b * 4 + 1, (double)s1);
        atomicAdd(sums + b * 4 + 2, (double)s2);
        atomicAdd(sums + b * 4 + 3, (double)s3);
    }
}
// loss = mean_b( S1/S0 + 1 - (S2+1)/(S3-S2+1) )
__global__ void structure_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, int B) {
    MDV_PDL_SYNC();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double acc = 0.0;
        for (int b = 0; b < B; ++b) {
            const double S0 = sums[b * 4], S1 = sums[b * 4 + 1], S2 = sums[b * 4 + 2], S3 = sums[b * 4 + 3];
            acc += S1 / S0 + 1.0 - (S2 + 1.0) / (S3 - S2 + 1.0);
        }
        loss[0] = (float)(acc / B);
    }
}
// dpred = coef/B * weit * [ (p - m)/S0  -  p(1-p) * ( m/(U+1) - (I+1)(1-m)/(U+1)^2 ) ],  I = S2, U = S3 - S2
//   d wiou / d p_i = -[ m w (U+1) - (I+1)(w - m w) ] / (U+1)^2   (U = sum (p+m)w - sum p m w  =>  dU/dp = w - m w)
__global__ void __launch_bounds__(256) structure_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ mask,
                                                             const float* __restrict__ weit, const double* __restrict__ sums,
                                                             const float* __restrict__ gout, float coef, float* __restrict__ dpred, int B, int HW,
                                                             int accumulate) {
    MDV_PDL_SYNC();
    const long long total = (long long)B * HW;
    const float g = (gout ? gout[0] : 1.f) * coef / (float)B;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int b = (int)(idx / HW);
        const float S0 = (float)sums[b * 4], I1 = (float)(sums[b * 4 + 2] + 1.0), U1 = (float)(sums[b * 4 + 3] - sums[b * 4 + 2] + 1.0);
        const float z = __ldg(pred + idx), m = __ldg(mask + idx), w = __ldg(weit + idx);
        const float p = 1.f / (1.f + __expf(-z));
        const float dbce = (p - m) / S0;
        const float diou = -(m * U1 - I1 * (1.f - m)) / (U1 * U1);
        const float v = g * w * (dbce + p * (1.f - p) * diou);
        dpred[idx] = accumulate ? dpred[idx] + v : v;
    }
}

}  // namespace

// ====================================================================================== C ABI
extern "C" int mdv_im2col_k(const float* in, int in_nchw, void* col, int col_bf16, int B, int Hi, int Wi, int Ho, int Wo, int C, int k,
                            int stride, int pad, int ldc, void* stream) {
    if (!in || !col || k < 1 || stride < 1 || pad < 0 || ldc < C * k * k) return MDV_ERR_ARG;
    const long long total = (long long)B * Ho * Wo * 32;      // one warp per output pixel
    if (total <= 0 || (long long)B * Ho * Wo >= 2147483647LL) return MDV_ERR_ARG;
#define MDV_I2C(TO_, K_, C_, L_) mdv_launch((im2col_k_kernel<TO_, K_, C_, L_>), dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, in, (TO_*)col, B, Hi, Wi, Ho, Wo, C, k, stride, pad, ldc, in_nchw)
    if (col_bf16)
        MDV_I2C(bf16, 0, 0, 0);
    else if (k == 7 && C == 3 && ldc == 152)
        MDV_I2C(float, 7, 3, 152);      // resnet.conv1
    else if (k == 7 && C == 2 && ldc == 104)
        MDV_I2C(float, 7, 2, 104);      // BiFusion_block.spatial
    else
        MDV_I2C(float, 0, 0, 0);
#undef MDV_I2C
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_col2im_k(const float* dcol, float* dx, int B, int Hi, int Wi, int Ho, int Wo, int C, int k, int stride, int pad,
                            int ldc, void* stream) {
    if (!dcol || !dx || k < 1 || stride < 1 || pad < 0 || ldc < C * k * k) return MDV_ERR_ARG;
    if (!fits_u32((long long)B * Hi * Wi * C)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(col2im_k_kernel, dim3(grid_for((long long)B * Hi * Wi * C)), dim3(256), 0, (cudaStream_t)stream, dcol, dx, B, Hi, Wi, Ho, Wo, C, k, stride, pad, ldc);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_maxpool3s2_fwd(const float* in, float* out, void* tap_u8, int B, int Hi, int Wi, int C, void* stream) {
    if (!in || !out || !tap_u8 || (C & 3)) return MDV_ERR_ARG;
    const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
    if (!fits_u32((long long)B * Hi * Wi * C)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(maxpool3s2_fwd_kernel, dim3(grid_for((long long)B * Ho * Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, in, out, (unsigned char*)tap_u8, B, Hi, Wi, Ho, Wo, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_maxpool3s2_bwd(const float* dout, const void* tap_u8, float* din, int B, int Hi, int Wi, int C, void* stream) {
    if (!dout || !din || !tap_u8 || (C & 3)) return MDV_ERR_ARG;
    const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
    if (!fits_u32((long long)B * Hi * Wi * C)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(maxpool3s2_bwd_kernel, dim3(grid_for((long long)B * Hi * Wi * (C / 4))), dim3(256), 0, (cudaStream_t)stream, dout, (const unsigned char*)tap_u8, din, B, Hi, Wi, Ho, Wo, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_add_act(const float* a, const float* b, float* out, long long n, int act, void* stream) {
    if (!a || !b || !out || n <= 0 || (n & 3) || (act != MDV_ACT_NONE && act != MDV_ACT_RELU)) return MDV_ERR_ARG;
    if (!fits_u32(n / 4)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(add_act_kernel, dim3(grid_for(n / 4)), dim3(256), 0, (cudaStream_t)stream, a, b, out, n / 4, act);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream) {
    if (!dy || !y || !dx || n <= 0 || (n & 3)) return MDV_ERR_ARG;
    if (!fits_u32(n / 4)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(act_bwd_kernel, dim3(grid_for(n / 4)), dim3(256), 0, (cudaStream_t)stream, dy, y, dx, n / 4);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_resize_ac_fwd(const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, int C, void* stream) {
    if (!in || !out || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1 || C < 1) return MDV_ERR_ARG;
    if (!fits_u32((long long)B * Ho * Wo * C)) return MDV_ERR_UNSUPPORTED;
    const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    if (!(C & 3))
        mdv_launch(resize_ac_fwd_kernel<4>, dim3(grid_for((long long)B * Ho * Wo * (C / 4))), dim3(256), 0, (cudaStream_t)stream, in, out, B, Hi, Wi, Ho, Wo, C, sy, sx);
    else
        mdv_launch(resize_ac_fwd_kernel<1>, dim3(grid_for((long long)B * Ho * Wo * C)), dim3(256), 0, (cudaStream_t)stream, in, out, B, Hi, Wi, Ho, Wo, C, sy, sx);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_resize_ac_bwd(const float* dout, float* din, int B, int Hi, int Wi, int Ho, int Wo, int C, void* stream) {
    if (!dout || !din || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1 || C < 1) return MDV_ERR_ARG;
    if (!fits_u32((long long)B * Ho * Wo * C)) return MDV_ERR_UNSUPPORTED;
    const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    const float ry = Hi > 1 ? (float)(Ho - 1) / (float)(Hi - 1) : (float)Ho, rx = Wi > 1 ? (float)(Wo - 1) / (float)(Wi - 1) : (float)Wo;
    if (!(C & 3))
        mdv_launch(resize_ac_bwd_kernel<4>, dim3(grid_for((long long)B * Hi * Wi * (C / 4))), dim3(256), 0, (cudaStream_t)stream, dout, din, B, Hi, Wi, Ho, Wo, C, sy, sx, ry, rx);
    else
        mdv_launch(resize_ac_bwd_kernel<1>, dim3(grid_for((long long)B * Hi * Wi * C)), dim3(256), 0, (cudaStream_t)stream, dout, din, B, Hi, Wi, Ho, Wo, C, sy, sx, ry, rx);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_structure_weit(const float* mask, float* weit, float* ws, int B, int H, int W, void* stream) {
    if (!mask || !weit || !ws || B < 1 || H < 1 || W < 1) return MDV_ERR_ARG;
    if (!fits_u32((long long)B * H * W)) return MDV_ERR_UNSUPPORTED;
    const long long total = (long long)B * H * W;
    mdv_launch(boxsum_rows_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, mask, ws, B, H, W, 15);
    MDV_CHECK_LAUNCH();
    mdv_launch(boxsum_cols_weit_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, (const float*)ws, mask, weit, B, H, W, 15);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_structure_loss_fwd(const float* pred, const float* mask, const float* weit, void* sums, float* loss, int B, int HW,
                                      void* stream) {
    if (!pred || !mask || !weit || !sums || !loss || B < 1 || HW < 1) return MDV_ERR_ARG;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 4 * B, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    int bps = mdv_cdiv(HW, 256 * 8);
    if (bps < 1) bps = 1;
    mdv_launch(structure_sums_kernel, dim3(B * bps), dim3(256), 0, (cudaStream_t)stream, pred, mask, weit, (double*)sums, HW, bps);
    MDV_CHECK_LAUNCH();
    mdv_launch(structure_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const double*)sums, loss, B);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_structure_loss_bwd(const float* pred, const float* mask, const float* weit, const void* sums, const float* gout, float coef,
                                      float* dpred, int B, int HW, int accumulate, void* stream) {
    if (!pred || !mask || !weit || !sums || !dpred || B < 1 || HW < 1) return MDV_ERR_ARG;
    if (!fits_u32((long long)B * HW)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(structure_bwd_kernel, dim3(grid_for((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, pred, mask, weit, (const double*)sums, gout, coef, dpred, B, HW, accumulate);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_gate_cat_fwd(const float* g, const float* pgate, const float* x, const float* vgate, const float* bp, float* out, int M,
                                int C1, int C2, int C3, int rows_per_sample, void* stream) {
    if (!g || !pgate || !out || M <= 0 || C1 <= 0 || (C1 & 3) || (C2 & 3) || (C3 & 3) || rows_per_sample <= 0) return MDV_ERR_ARG;
    if ((C2 && (!x || !vgate)) || (C3 && !bp) || C2 > 512) return MDV_ERR_ARG;
    mdv_launch(gate_cat_fwd_kernel, dim3(grid_for((long long)M * 32)), dim3(256), 0, (cudaStream_t)stream, g, pgate, x, vgate, bp, out, M, C1, C2, C3, rows_per_sample);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_gate_cat_bwd(const float* dout, const float* g, const float* pgate, const float* x, const float* vgate, float* dg, float* dp,
                                float* dx, float* dv, float* dbp, int M, int C1, int C2, int C3, int rows_per_sample, void* stream) {
    if (!dout || !g || !pgate || !dg || !dp || M <= 0 || C1 <= 0 || (C1 & 3) || (C2 & 3) || (C3 & 3) || rows_per_sample <= 0) return MDV_ERR_ARG;
    if ((C2 && (!x || !vgate || !dx || !dv)) || (C3 && !dbp) || C2 > 512 || (M % rows_per_sample)) return MDV_ERR_ARG;
    if (C2) {
        cudaError_t e = cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)(M / rows_per_sample) * C2, (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    // bands of rows per warp: enough warps to fill the machine, long enough to amortise the dv atomics
    int band = rows_per_sample;
    while (band > 16 && (long long)(M / rows_per_sample) * ((rows_per_sample + band / 2 - 1) / (band / 2)) <= (long long)MDV_NUM_SMS * 64) band /= 2;
    const long long warps = (long long)(M / rows_per_sample) * ((rows_per_sample + band - 1) / band);
    mdv_launch(gate_cat_bwd_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, dout, g, pgate, x, vgate, dg, dp, dx, dv,
               dbp, M, C1, C2, C3, rows_per_sample, band);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_channel_pool_fwd(const float* x, float* out, void* arg_i32, int M, int C, void* stream) {
    if (!x || !out || !arg_i32 || M <= 0 || C <= 0) return MDV_ERR_ARG;
    mdv_launch(channel_pool_fwd_kernel, dim3(grid_for((long long)M * 32)), dim3(256), 0, (cudaStream_t)stream, x, out, (int*)arg_i32, M, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_channel_pool_bwd(const float* dout, const void* arg_i32, float* dx, int M, int C, void* stream) {
    if (!dout || !dx || !arg_i32 || M <= 0 || C <= 0) return MDV_ERR_ARG;
    if (!fits_u32((long long)M * C)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(channel_pool_bwd_kernel, dim3(grid_for((long long)M * C)), dim3(256), 0, (cudaStream_t)stream, dout, (const int*)arg_i32, dx, M, C);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}

extern "C" int mdv_dropout2d(const float* x, float* out, int M, int C, int rows_per_sample, float p, const void* rng, uint32_t drop_stream,
                             void* stream) {
    if (!x || !out || !rng || M <= 0 || C <= 0 || (C & 3) || rows_per_sample <= 0 || !(p > 0.f) || !(p < 1.f)) return MDV_ERR_ARG;
    if (!fits_u32((long long)M * C)) return MDV_ERR_UNSUPPORTED;
    mdv_launch(dropout2d_kernel, dim3(grid_for((long long)M * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, out, M, C, rows_per_sample, p,
               (const unsigned long long*)rng, drop_stream);
    MDV_CHECK_LAUNCH();
    return MDV_OK;
}
