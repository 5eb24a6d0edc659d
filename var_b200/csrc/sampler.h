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
  // more_smooth (models/var.py:178-180): optional Gumbel soft embedding of the filtered row
  const float* q_gumbel;  // NULL = off; [B*l, V] Exp(1) noise (gumbel = -log q)
  float tau, logit_mul;   // softmax((x * logit_mul + gumbel) / tau)
  const float* codebook;  // [V, Cvae]
  int Cvae;
  float* h_out;           // [B*l, Cvae]
};

int sample_launch(const SampleArgs& a, cudaStream_t st);

// out[s,t] = sum_v softmax(mix(lc[s,t,:], lu[t,:]))_v * dists[gt[t], v]; lu may be null (no mixing); top_k 0 = all
int cfg_token_expected_dist(const float* lc, const float* lu, const int* gt, const float* t_row, const float* dists,
                            int n_seq, int L, int V, int top_k, float* out, cudaStream_t st);

// VAR.smooth_sampling token selection (models/var.py:483-536); see sampler.cu
int neighbor_select(const float* logits, int B, int l, int V, double t, const int* gt, const int* neighbors,
                    const float* dists, int n_nb, int cand_count, int thr_mode, float thr, float ratio, void* tok_out,
                    float* lp_out, float* dlp_out, cudaStream_t st);

}  // namespace vb
