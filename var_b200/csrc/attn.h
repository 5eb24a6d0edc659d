// Block-causal attention launcher interface (no device code).
#pragma once
#include <cuda_runtime.h>

#define VB_MAX_SCALES 16

namespace vb {

struct AttnArgs {
  const void* q;  // bf16 [n_seq, H, Lq, 64]   (L2-normalised, scaled)
  const void* k;  // bf16 [n_seq, H, Lmax, 64] (L2-normalised) ; rows >= visible range must be finite
  const void* v;  // bf16 [n_seq, H, Lmax, 64]
  void* out;      // bf16 [n_seq, Lq, H*64]
  int n_seq, H, Lq, Lmax;
  int q_pos0;                    // absolute sequence position of query row 0
  int n_scales;                  // pyramid levels
  int level_end[VB_MAX_SCALES];  // cumulative token count after each level
  float max_score;               // upper bound on |q.k| if the caller knows one (the per-head scale), else 0
  int q_log2;                    // q is pre-multiplied by log2(e) (needs 0 < max_score <= 43; bound in natural units)
};

int attn_launch(const AttnArgs& a, cudaStream_t st);
// Warp-per-item kernel for the first KV-cached scales (attn_small.cu): at most 32 queries and 64 visible keys per
// (sequence, head). kv_vis = keys visible to the last query row of the call.
bool attn_small_applies(const AttnArgs& a, int kv_vis);
int attn_small_launch(const AttnArgs& a, int kv_vis, cudaStream_t st);

}  // namespace vb
