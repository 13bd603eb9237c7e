// Internal interface between attn.cu (cross-token statistics, C ABI) and attn_tile.cu (token-parallel tiled kernels).
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

int attn_tile_fwd(const bf16* qkv, const float* A, const float* gate, const CrpeW& cw, bf16* out, float scale, int B, int H, int W, int C,
                  int Ch, cudaStream_t st);
int attn_tile_bwd(const bf16* qkv, const bf16* dy, const bf16* yout, const float* gate, const float* A, const float* At, const float* dA,
                  const float* dAt, const float* rk, const float* kmax, const float* zsum, const CrpeW& cw, const CrpeG& cg, bf16* dqkv,
                  float* dgate, float scale, int B, int H, int W, int C, int Ch, cudaStream_t st);
