/* mdvit_b200 — C ABI of the B200 (sm_100a) kernels behind the MDViT hot path.
 *
 * The reference (siyi-wind/MDViT) is pure PyTorch and has no FFI layer of its own; its drop-in
 * boundary is the nn.Module contract (Models/Transformer/mdvit.py:484-504 ctor, :667 forward).
 * This library sits underneath that module: every entry point replaces the aten op(s) the reference
 * module launches at the cited line.  The Python host (mdvit_b200/_lib.py) binds it with ctypes.
 *
 * Conventions: all pointers are DEVICE pointers unless noted; activations are token-major
 * ([B, H*W, C] == NHWC); `stream` is a cudaStream_t passed as void*; every function returns
 * 0 on success, a positive cudaError_t on a CUDA failure, or a negative MDV_ERR_* code on bad
 * arguments.  Nothing is allocated or freed by the library; workspaces are caller-owned.
 * Functions are re-entrant (no global mutable state apart from mdv_gemm_tune, a debug knob).
 */
#ifndef MDVIT_B200_H
#define MDVIT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDV_ACT_NONE 0
#define MDV_ACT_GELU 1
#define MDV_ACT_RELU 2
#define MDV_ACT_HSWISH 3

int mdv_version(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA) */
/* Epilogue applied to acc = A.W^T, in this order:
 *   v = acc + bias[n];  out_preact[m,n] = bf16(v);  v = act(v);  v *= gelu'(mul_gelu_grad[m,n]);
 *   v *= dropout_mask(rng, drop_stream, m*N+n)/(1-p);  v *= rowscale[m / rows_per_scale];
 *   v += residual[m,n];  out[m,n] (=|+=) v                                                   */
typedef struct MdvGemmEpi {
    const float* bias;          /* [N] fp32 or NULL */
    const float* residual;      /* [M, ld_res] fp32 or NULL */
    const void* mul_gelu_grad;  /* [M, ld_mul] bf16 pre-activation u, or NULL */
    void* out_preact;           /* [M, ld_preact] bf16 or NULL */
    void* out;                  /* [M, ldc] bf16 or fp32 */
    const float* rowscale;      /* [ceil(M / rows_per_scale)] fp32 or NULL (DropPath per-sample scale) */
    const void* rng;            /* device uint64[2] {seed, step}; may be NULL when dropout_p == 0 */
    int ld_res, ld_mul, ld_preact, ldc;
    int rows_per_scale;
    int out_bf16;               /* 1: out is bf16, 0: fp32 */
    int act;                    /* MDV_ACT_NONE | MDV_ACT_GELU */
    int accumulate;             /* fp32 out only: out += v */
    float dropout_p;
    uint32_t drop_stream;
} MdvGemmEpi;

/* C[M,N] = epi(A[M,K] . W[N,K]^T); A, W bf16 row-major with pitches lda, ldw (elements, multiples of 8).
 * nn.Linear / 1x1 Conv2d forward (mdvit.py:288,310; mpvit.py:72-76; Decoders.py:59,197,317-333) and,
 * with W := W^T, their input gradients. */
int mdv_gemm_nt(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const MdvGemmEpi* epi, void* stream);

/* C[P,Q] += A[R,P]^T . B[R,Q]  (fp32 atomics; A, B bf16 row-major).  Weight gradients of the above. */
int mdv_gemm_tn(const void* A, int lda, const void* B, int ldb, int R, int P, int Q, float* C, int ldc, void* stream);

/* Debug/tuning knob (0 = automatic): force tile N, pipeline stages, TN split count. */
int mdv_gemm_tune(int force_bn, int force_stages, int force_split);

#ifdef __cplusplus
}
#endif
#endif
