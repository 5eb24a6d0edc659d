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
  // split encode: 0 fused (fp32 search inside), 1 init + pool scale si_begin, 2 update si_begin from idx + pool si_begin+1
  int split;
  float* z_out;           // [B*l, Cvae] pooled tokens of the pooled scale
  void* zb_out;           // [B*l, 64] bf16, zero padded
  float* zz_out;          // [B*l]
};

// Tensor-core nearest-neighbour search (quant_search_sm100.cu): idx_out[n] = argmin_v |z_n - e_v|^2, first index on
// ties, bit-identical to the fp32 search (bf16 UMMA distance filter + exact fp32 re-rank of the candidates).
struct QuantSearchArgs {
  const void* zb;         // [N, 64] bf16 tokens (K padded)
  const float* z;         // [N, 32] fp32 tokens
  const float* zz;        // [N]
  const void* cb_bf16;    // [V, 64] bf16 codebook (K padded)
  const float* codebook;  // [V, 32] fp32
  const float* ee;        // [V] |e_v|^2 (quant_prepare_codebook)
  int N, V;
  void* idx_out;          // int64 [N]
};
int quant_search_launch(const QuantSearchArgs& a, cudaStream_t st);
// cb_bf16[v, 0:32] = bf16(codebook[v]), [32:64] = 0; ee[v] = |e_v|^2 in the oracle's summation order
int quant_prepare_codebook(const float* codebook, void* cb_bf16, float* ee, int V, cudaStream_t st);

int quant_launch(const QuantArgs& a, cudaStream_t st);

}  // namespace vb
