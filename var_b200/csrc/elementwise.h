// HBM-bound helper kernels: launcher interface (no device code).
#pragma once
#include <cuda_runtime.h>

#include "attn.h"  // VB_MAX_SCALES

namespace vb {

struct AttnLevelsPOD {
  int n;
  int end[VB_MAX_SCALES];
};

// out bf16[M,C] = LN(x[M,C]) * (1 + scale[m / rows_per_seq]) + shift[m / rows_per_seq]; scale/shift row stride ada_ld
int ln_modulate(const float* x, const float* scale, const float* shift, int ada_ld, int rows_per_seq, void* out, int M,
                int C, float eps, cudaStream_t st);
// deferred-LayerNorm table inputs: out[4][n, C] bf16 = (1 + s1, h1, 1 + s2, h2) for columns taken from ada[n, ada_ld]
int ln_table_inputs(const float* s1, const float* h1, const float* s2, const float* h2, int ada_ld, void* out, int n, int C,
                    cudaStream_t st);
// out bf16[n_seq,C] = SiLU(class_emb[labels])
int cond_silu(const float* class_emb, const int* labels, void* out, int n_seq, int C, cudaStream_t st);
int expand_shared_aln(const float* shared, int shared_ld, const float* gss, float* ada, int ada_ld, int depth, int C,
                      int n_seq, cudaStream_t st);
int embed_tokens(const float* x_in, int n_x, int l_in, const int* labels, const float* class_emb,
                 const float* pos_start, const float* lvl_pos, const float* w_word_t, const float* b_word, float* out,
                 int n_seq, int l, int first_rows, int pos0, int C, int Cv, cudaStream_t st);
int score_finalize(const void* part, int n_tiles, const float* gt_logit, int n_seq, int L, int n_scales,
                   const int* level_end, float* tok_logp, float* per_scale, float* total, int first_pos,
                   cudaStream_t st);

int cfg_token_logprob(const float* lc, const float* lu, const int* gt, const float* t_row, int n_seq, int L, int V,
                      float* tok_logp, cudaStream_t st);
int scale_sums(const float* tok_logp, int n_seq, int L, int n_scales, const int* level_end, int first_pos,
               float* per_scale, float* total, cudaStream_t st);

}  // namespace vb
