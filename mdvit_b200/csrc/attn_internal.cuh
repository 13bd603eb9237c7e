// Internal interface between attn.cu (C ABI, column max, DA gate) and attn_strip.cu (cross-token sums + strip kernels).
#pragma once
#include "../../include/mdvit_b200.h"
#include "common.cuh"

struct CrpeW {               // three depthwise filters: heads [0,2) 3x3, [2,5) 5x5, [5,8) 7x7  (mdvit.py:423)
    const float* w[3];
    const float* b[3];
};
struct CrpeG {
    float* w[3];
    float* b[3];
};

long long attn_strip_ws_floats(int B, int C, int Ch);
int attn_strip_fwd(const bf16* qkv, const float* gate, const CrpeW& cw, float* kmax, float* zsum, float* A, float* ws, bf16* out,
                   bf16* eout, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st);
int attn_strip_bwd(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* kmax, const float* zsum,
                   const float* A, float* ws, const CrpeW& cw, const CrpeG& cg, const bf16* ein, bf16* dqkv, float* dgate,
                   float* dbias_qkv, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st);
