// CFG-mix + top-k/top-p + sampling launcher interface (no device code).
#pragma once
#include <cuda_runtime.h>

namespace vb {

struct SampleArgs {
  const float* logits;  // use_cfg: [2B, l, V] (cond rows first, then uncond); else [B, l, V]
  int B, l, V;
  int use_cfg;
  double t;        // guidance strength of this scale: cfg * si / (S-1)   (models/var.py:172)
  const float* q;  // [B*l, V] Exp(1) noise
  int top_k;       // 0 = off
  float top_p;     // 0 = off
  void* idx_out;   // int64 [B, l]
  float* mixed_out;  // optional [B, l, V]: the mixed logits (before filtering)
};

int sample_launch(const SampleArgs& a, cudaStream_t st);

}  // namespace vb
