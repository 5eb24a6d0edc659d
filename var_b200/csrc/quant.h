// Multi-scale residual quantizer launcher interface (no device code).
#pragma once
#include <cuda_runtime.h>

#include "attn.h"  // VB_MAX_SCALES

namespace vb {

struct QuantArgs {
  int B, Cvae, H, W, V, S;
  int ph[VB_MAX_SCALES], pw[VB_MAX_SCALES], phi_of_scale[VB_MAX_SCALES];
  int n_phi;
  float resi;
  const float* codebook;  // [V,Cvae]
  const float* phi_w;     // [n_phi,Cvae,Cvae,3,3]
  const float* phi_b;     // [n_phi,Cvae]
  int si_begin, si_end;   // scales processed: [si_begin, si_end)
  const float* f;         // encode: [B,Cvae,H,W] features; NULL = decode (idx is an input)
  float* f_rest;          // encode workspace [B,Cvae,H,W]
  float* f_hat;           // [B,Cvae,H,W] running reconstruction (in/out)
  int zero_fhat;          // 1: start from zeros
  int idx_concat;         // 1: idx = all scales concatenated ([B,l_s] blocks); 0: idx = scale si_begin only
  void* idx;              // int64
  float* fhat_list;       // optional [S,B,Cvae,H,W]: f_hat after every processed scale
  float* next_tokens;     // optional token-major area(f_hat -> next scale), [B, next_stride, Cvae]
  int next_stride;
  float* next_nchw;       // optional [B,Cvae,ph_next,pw_next] (single-scale step)
};

int quant_launch(const QuantArgs& a, cudaStream_t st);

}  // namespace vb
