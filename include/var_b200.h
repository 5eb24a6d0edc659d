/* var_b200 — C ABI of the B200-native VAR next-scale-prediction hot path.
 *
 * The reference (culiver/VAR, pure Python/PyTorch) has no FFI layer: the hot path sits behind nn.Module
 * methods. This header is the boundary a maintainer binds instead of the ATen calls those
 * methods make; every entry cites the reference lines it replaces. Conventions:
 *   - plain pointers and sizes, all device pointers unless noted "host";
 *   - the last argument is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return 0 on success, a negative VAR_B200_ERR_* code otherwise; var_b200_last_error() gives the message;
 *   - no allocation inside: workspaces are caller-provided (sizes from the *_workspace_bytes helpers);
 *   - bf16 tensors are raw 16-bit storage (torch.bfloat16), fp32 otherwise; indices are int64 where the
 *     reference returns LongTensors, int32 where noted.
 */
#ifndef VAR_B200_H
#define VAR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(VAR_B200_BUILD)
#define VAR_B200_API __attribute__((visibility("default")))
#else
#define VAR_B200_API
#endif

#define VAR_B200_OK 0
#define VAR_B200_ERR_ARG (-1)
#define VAR_B200_ERR_CUDA (-2)
#define VAR_B200_ERR_DRIVER (-3)
#define VAR_B200_ERR_WORKSPACE (-4)

/* Message describing the most recent error on the calling thread (never NULL). */
VAR_B200_API const char* var_b200_last_error(void);
/* Number of kernels this library has launched so far in the process. */
VAR_B200_API long long var_b200_launch_count(void);
/* Per-kernel CUDA-event timing (measurement aid for bench.py): between _begin and _end every kernel this library
 * launches is bracketed by events on its stream; _end synchronises the device and returns the summed milliseconds
 * and launch counts per kernel class (index = VAR_B200_PK_*). */
enum { VAR_B200_PK_GEMM_BIAS_F32 = 0, VAR_B200_PK_GEMM_BIAS_BF16, VAR_B200_PK_GEMM_GELU, VAR_B200_PK_GEMM_GATE_RESID,
       VAR_B200_PK_GEMM_QKV, VAR_B200_PK_GEMM_SCORE, VAR_B200_PK_ATTN, VAR_B200_PK_LN, VAR_B200_PK_EMBED,
       VAR_B200_PK_COND, VAR_B200_PK_SAMPLE, VAR_B200_PK_QUANT, VAR_B200_PK_SCORE_FIN, VAR_B200_PK_OTHER,
       VAR_B200_PK_COUNT };
VAR_B200_API void var_b200_profile_begin(void);
VAR_B200_API int var_b200_profile_end(double* ms_by_kind /* host */, long long* n_by_kind /* host */, int n_kinds);

/* ------------------------------------------------------------------------------------------------
 * GEMM family:  D[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation on tcgen05/TMEM.
 * Replaces F.linear / nn.Linear calls of the transformer: models/basic_var.py:93 (QKV), :119 (proj),
 * :52 (fc1/fc2), models/var.py:124 (head), basic_var.py:156,173 (adaLN linears).
 * ---------------------------------------------------------------------------------------------- */
enum {
  VAR_B200_EPI_BIAS_F32 = 0,   /* out fp32 = acc + bias */
  VAR_B200_EPI_BIAS_BF16 = 1,  /* out bf16 = acc + bias */
  VAR_B200_EPI_GELU_BF16 = 2,  /* out bf16 = gelu_tanh(acc + bias)              basic_var.py:40,52 */
  VAR_B200_EPI_GATE_RESID = 3, /* out fp32 = resid + gate[m / rows_per_seq] * (acc + bias)   basic_var.py:157-158 */
  VAR_B200_EPI_QKV = 4,        /* +[q_bias,0,v_bias]; q,k L2-normalised per head; q *= scale; scatter into q buffer
                                  and the preallocated K/V cache                  basic_var.py:93-109 */
  VAR_B200_EPI_SCORE = 5       /* per-row partial log-sum-exp and ground-truth logit eval_prob.py:446-452 */
};

#define VAR_B200_GEMM_EPI_PARTS 2
typedef struct var_b200_gemm_args {
  const void* A; /* [M,K] bf16 row-major */
  const void* W; /* [N,K] bf16 row-major (nn.Linear weight) */
  int M, N, K;
  int epilogue;      /* VAR_B200_EPI_* */
  int force_bn;      /* 0 = auto tile width, else 128 / 192 / 256; | 0x10000 forces the 1-CTA kernel */
  const float* bias; /* [N] or NULL (required for QKV / SCORE) */
  void* out;         /* [M,N] fp32 or bf16 */
  /* GATE_RESID */
  const float* resid; /* [M,N] fp32, may alias out */
  const float* gate;  /* gate[(m / rows_per_seq) * gate_ld + n] */
  int rows_per_seq;
  int gate_ld;
  /* QKV (N == 3*C, head_dim 64) */
  void* q_out;          /* bf16 [n_seq, H, rows_per_seq, 64] */
  void* k_cache;        /* bf16 [n_seq, H, Lmax, 64], rows written at pos0 + t */
  void* v_cache;        /* bf16 [n_seq, H, Lmax, 64] */
  const float* q_scale; /* [H] exp(min(scale_mul, ln 100)) */
  int C, H, pos0, Lmax;
  int no_l2norm;        /* 1: q *= q_scale[head] without the L2 normalisation of q and k (attn_l2_norm=False) */
  /* SCORE */
  const int32_t* gt; /* row m uses gt[m % gt_mod] */
  int gt_mod;        /* 0 = M */
  float* part;       /* [M, VAR_B200_GEMM_EPI_PARTS*ceil(N / var_b200_gemm_tile_n(N)), 2]: (max, sumexp) per row, tile and
                        epilogue warp of the row's lane quarter */
  float* gt_logit;   /* [M] */
  /* Deferred LayerNorm (see var_b200_blocks). Producer, GATE_RESID only, enabled by ln_a_out != NULL: */
  void* ln_a_out;         /* bf16 [M,N] = out * (1 + ln_scale[m / rows_per_seq]) */
  const float* ln_scale;  /* ln_scale[(m / rows_per_seq) * gate_ld + n]: adaLN scale of the LayerNorm that follows */
  float* ln_part_out;     /* [M, var_b200_gemm_ln_parts(M,N), 2]: partial (sum, sum of squares) of the new out rows */
  /* Consumer, QKV / GELU_BF16 only, enabled by ln_part_in != NULL: acc + bias is replaced by
   * rstd_m * (acc - mean_m * ln_u[label]) + ln_v[label], (mean, rstd) over ln_C columns from the partials. */
  const float* ln_part_in;
  int ln_parts, ln_C;
  float ln_eps;
  const float* ln_u;        /* [n_classes, N] */
  const float* ln_v;        /* [n_classes, N] (carries the bias) */
  const int32_t* ln_labels; /* [M / rows_per_seq] */
} var_b200_gemm_args_t;

VAR_B200_API int var_b200_gemm_bf16(const var_b200_gemm_args_t* args, void* stream);
/* Tile width (128/192/256) the GEMM uses for a given N. */
VAR_B200_API int var_b200_gemm_tile_n(int N);
/* Partials per row that a deferred-LayerNorm producer launch (GATE_RESID with ln_a_out) of this shape writes. */
VAR_B200_API int var_b200_gemm_ln_parts(int M, int N);

/* Hardware probe used by the test-suite to pin UMMA shared-memory descriptor encodings:
 * D[128,N] = A[128,64] * B, B = [N,64] (K-major) or [64,N] (MN-major). */
VAR_B200_API int var_b200_umma_probe(const void* A, const void* B, float* D, int N, int b_mn_major, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Block-causal attention (models/basic_var.py:98-117, mask of models/var.py:107-112), tcgen05/TMEM.
 * q: bf16 [n_seq,H,Lq,64], k/v: bf16 [n_seq,H,Lmax,64] (the preallocated KV cache), out: bf16 [n_seq,Lq,H*64].
 * A query at absolute position q_pos0 + i of pyramid level s attends keys [0, level_end[s]).
 * max_score: an upper bound on |q.k| if the caller knows one (for VAR the largest per-head scale
 * exp(min(scale_mul, ln 100)), basic_var.py:101, a model constant), else 0. With 0 < max_score <= 43 the softmax runs
 * against that fixed reference (no row maximum, eight softmax warps per CTA); otherwise the general kernel tracks a
 * per-row reference maximum and rebases on overflow. Both give softmax(q k^T + mask) v.
 * q_log2 != 0: q is already multiplied by log2(e) (var_b200_model_t folds it into q_scale), i.e. the scores are base-2
 * exponents and the bounded-score kernel feeds them to exp2 unchanged; requires 0 < max_score <= 43 (max_score stays in
 * natural units), else VB_ERR_ARG.
 * ---------------------------------------------------------------------------------------------- */
#define VAR_B200_MAX_SCALES 16
VAR_B200_API int var_b200_attention(const void* q, const void* k, const void* v, void* out, int n_seq, int H, int Lq,
                                    int Lmax, int q_pos0, int n_scales, const int* level_end /* host */,
                                    float max_score, int q_log2, void* stream);

/* LN(x)*(1+scale[seq])+shift[seq] -> bf16 (models/basic_var.py:157-158,174). scale/shift: row stride ada_ld. */
VAR_B200_API int var_b200_ln_modulate(const float* x, const float* scale, const float* shift, int ada_ld,
                                      int rows_per_seq, void* out_bf16, int M, int C, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-scale residual quantizer (models/quant.py). All tensors device fp32 NCHW unless noted; idx is int64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct var_b200_quant {
  int Cvae, V, n_scales;                 /* Cvae must be 32 */
  int ph[VAR_B200_MAX_SCALES], pw[VAR_B200_MAX_SCALES]; /* patch grid per scale; last = latent H,W */
  int phi_of_scale[VAR_B200_MAX_SCALES]; /* Phi conv used by each scale (quant.py:218-243) */
  int n_phi;
  float resi;                            /* |quant_resi| (quant.py:204) */
  const float* codebook;                 /* [V,Cvae] embedding.weight */
  const float* phi_w;                    /* [n_phi,Cvae,Cvae,3,3] */
  const float* phi_b;                    /* [n_phi,Cvae] */
} var_b200_quant_t;

/* f_to_idxBl_or_fhat (quant.py:135-166). idx_out: int64, scale blocks [B, ph*pw] concatenated.
 * fhat_list: NULL or [S,B,Cvae,H,W] (the to_fhat=True list). work: var_b200_quant_encode_workspace(qz, B) bytes.
 * search_mode 0: codebook search on the tensor cores (bf16 UMMA distance filter + exact fp32 re-rank, one search
 * launch per scale); 1: single fused kernel with an fp32 CUDA-core search. Both return identical indices. */
VAR_B200_API size_t var_b200_quant_encode_workspace(const var_b200_quant_t* qz, int B);
VAR_B200_API int var_b200_quant_encode(const var_b200_quant_t* qz, const float* f, int B, int64_t* idx_out,
                                       float* fhat_list, void* work, size_t work_bytes, int search_mode, void* stream);
/* idxBl_to_var_input (quant.py:169-184) and embed_to_fhat(all_to_max_scale=True) (quant.py:107-121) in one pass.
 * var_input: NULL or [B, L - l_0, Cvae]; fhat_list: NULL or [S,B,Cvae,H,W]; fhat_last: [B,Cvae,H,W] (required). */
VAR_B200_API int var_b200_quant_decode(const var_b200_quant_t* qz, const int64_t* idx, int B, float* var_input,
                                       float* fhat_list, float* fhat_last, void* stream);
/* get_next_autoregressive_input (quant.py:187-196) with h = codebook[idx]: f_hat updated in place;
 * next_tokens: NULL or [B, l_{si+1}, Cvae] (token-major word_embed input); next_nchw: NULL or [B,Cvae,ph,pw]. */
VAR_B200_API int var_b200_quant_next_input(const var_b200_quant_t* qz, int si, float* f_hat, const int64_t* idx_si, int B,
                                           float* next_tokens, float* next_nchw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CFG mix + top-k/top-p + sampling from caller-supplied Exp(1) noise (models/var.py:172-175, helpers.py:6-19).
 * logits: [2B,l,V] (cond then uncond) when use_cfg, else [B,l,V]. q: [B*l,V]. idx_out: int64 [B,l].
 * ---------------------------------------------------------------------------------------------- */
VAR_B200_API int var_b200_cfg_topk_sample(const float* logits, int B, int l, int V, int use_cfg, double t, const float* q,
                                          int top_k, float top_p, int64_t* idx_out, float* mixed_out, void* stream);
/* The same sampler plus the more_smooth soft embedding (models/var.py:178-180, helpers.py:22-36):
 * h_out[B*l, Cvae] = softmax((x * logit_mul - log(q_gumbel)) / tau) @ codebook over the top-k/top-p filtered row x,
 * q_gumbel [B*l, V] ~ Exp(1) drawn after q (the reference's order of generator use). */
VAR_B200_API int var_b200_cfg_topk_sample_smooth(const float* logits, int B, int l, int V, int use_cfg, double t,
                                                 const float* q, int top_k, float top_p, int64_t* idx_out,
                                                 float* mixed_out, const float* q_gumbel, float tau, float logit_mul,
                                                 const float* codebook, int Cvae, float* h_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VAR transformer (models/var.py, models/basic_var.py). Weights are caller-owned device buffers packed once:
 * GEMM weights bf16 in nn.Linear layout [out,in], everything else fp32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct var_b200_block_weights {
  const void* w_qkv;    /* bf16 [3C,C]   attn.mat_qkv.weight */
  const float* b_qkv;   /* [3C] = cat(q_bias, 0, v_bias)            basic_var.py:93 */
  const float* q_scale; /* [H]  = exp(min(scale_mul_1H11, ln 100))   basic_var.py:101 */
  const void* w_proj;   /* bf16 [C,C] */
  const float* b_proj;  /* [C] */
  const void* w_fc1;    /* bf16 [4C,C] */
  const float* b_fc1;   /* [4C] */
  const void* w_fc2;    /* bf16 [C,4C] */
  const float* b_fc2;   /* [C] */
  /* Deferred-LayerNorm tables (var_b200_ln_tables), per class c in [0, num_classes]; all four NULL = not built, the
   * block then runs a separate LayerNorm-modulate pass in front of the QKV and fc1 GEMMs. */
  const float* u_qkv;   /* [num_classes+1, 3C] = W_qkv (1 + scale1_c) */
  const float* v_qkv;   /* [num_classes+1, 3C] = W_qkv shift1_c + b_qkv */
  const float* u_fc1;   /* [num_classes+1, 4C] = W_fc1 (1 + scale2_c) */
  const float* v_fc1;   /* [num_classes+1, 4C] = W_fc1 shift2_c + b_fc1 */
} var_b200_block_weights_t;

typedef struct var_b200_model {
  int depth, C, H, V, Cvae, n_scales, num_classes, shared_aln;
  float norm_eps;
  int patch_nums[VAR_B200_MAX_SCALES];
  const var_b200_block_weights_t* blocks; /* HOST array [depth] */
  /* adaLN linears stacked along the output dimension. shared_aln=0: rows = [blocks.0.ada_lin (6C) ... blocks.{d-1}
   * (6C), head_nm.ada_lin (2C)]; shared_aln=1: rows = [shared_ada_lin (6C), head_nm.ada_lin (2C)]. */
  const void* w_ada;     /* bf16 [ada_rows, C] */
  const float* b_ada;    /* [ada_rows] */
  int ada_rows;
  const float* ada_gss;  /* shared_aln only: [depth,6,C] (blocks.i.ada_gss), else NULL */
  const void* w_head;    /* bf16 [V,C] */
  const float* b_head;   /* [V] */
  const float* w_word;   /* [Cvae,C] word_embed.weight TRANSPOSED (coalesced per-channel loads) */
  const float* b_word;   /* [C] */
  const float* class_emb; /* [num_classes+1, C] */
  const float* pos_start; /* [first_l, C] */
  const float* lvl_pos;   /* [L, C] = lvl_embed[lvl_1L] + pos_1LC   (var.py:153,207) */
  float attn_max_score;   /* max over blocks and heads of exp(min(scale_mul, ln 100)) (bound on |q.k|, natural units,
                             see var_b200_attention); 0 = unknown */
  int attn_q_log2;        /* blocks[].q_scale already carries a factor log2(e) (only with 0 < attn_max_score <= 43) */
  int attn_no_l2norm;     /* 1: attn_l2_norm=False (basic_var.py:72): q and k are NOT normalised, q_scale holds the softmax
                             scale 0.25/sqrt(head_dim) for every head, attn_max_score must be 0 (unbounded scores) */
} var_b200_model_t;

/* Row stride (floats) of the per-sequence adaLN parameter table: (6*depth + 2) * C.
 * Column layout: block i -> [gamma1,gamma2,scale1,scale2,shift1,shift2] at 6*i*C (basic_var.py:156);
 * head_nm -> [scale, shift] at 6*depth*C (basic_var.py:173). */
VAR_B200_API int var_b200_ada_ld(const var_b200_model_t* m);
/* ada_out[n_seq, ada_ld] = Linear(SiLU(class_emb[labels])) for every block and the head (loop-invariant over the
 * 10 AR scales; the reference recomputes it per block per call, basic_var.py:156). work: var_b200_ada_workspace(). */
VAR_B200_API size_t var_b200_ada_workspace(const var_b200_model_t* m, int n_seq);
VAR_B200_API int var_b200_ada_params(const var_b200_model_t* m, const int32_t* labels, int n_seq, float* ada_out,
                                     void* work, size_t work_bytes, void* stream);

/* Token embedding (var.py:200-207 / :153-154,185-187): rows t < first_rows are start tokens
 * (class_emb[label] + pos_start[t]), the rest word_embed(x_in[seq % n_x, t - first_rows]); + lvl_pos[pos0 + t]. */
VAR_B200_API int var_b200_embed(const var_b200_model_t* m, const float* x_in, int n_x, int l_in, const int32_t* labels,
                                int n_seq, int l, int first_rows, int pos0, float* x_out, void* stream);

/* depth x AdaLNSelfAttn.forward (basic_var.py:152-159) on x[n_seq, l, C] in place.
 * Teacher-forced: l = L, pos0 = 0, kv = scratch [2][n_seq,H,L,64] shared by all layers (kv_layer_stride = 0).
 * KV-cached step: l = pn^2, pos0 = tokens already cached, kv = [depth][2][n_seq,H,Lmax,64]
 * (kv_layer_stride = 2*n_seq*H*Lmax*64 elements). x_dump: NULL or [depth, n_seq*l, C] copies of x after each block. */
VAR_B200_API size_t var_b200_blocks_workspace(const var_b200_model_t* m, int n_seq, int l);
/* labels: NULL, or the int32 class of every sequence (the labels `ada` was computed from). With labels and the
 * blocks' u_/v_ tables present the adaLN LayerNorms (basic_var.py:157-158) cost no pass of their own: the proj / fc2
 * epilogues emit x_new * (1 + scale_next) in bf16 plus per-row partial sums, and the QKV / fc1 epilogues finish
 * LN(x)(1+scale)+shift = rstd * (acc - mean * U[label]) + V[label]. Only block 0's first LayerNorm runs as a pass. */
VAR_B200_API int var_b200_blocks(const var_b200_model_t* m, float* x, const float* ada, const int32_t* labels, int n_seq,
                                 int l, int pos0, void* kv, size_t kv_layer_stride, int Lmax, float* x_dump, void* work,
                                 size_t work_bytes, void* stream);
/* Builds block `block`'s deferred-LayerNorm tables from the adaLN table of ALL classes: ada_all[n_cls, ada_ld] =
 * var_b200_ada_params(labels = 0..n_cls-1). Outputs fp32 [n_cls,3C], [n_cls,3C], [n_cls,4C], [n_cls,4C]; work:
 * var_b200_ln_tables_workspace(m, n_cls) bytes. Four small GEMMs on the tcgen05 kernel; run once per weight pack. */
VAR_B200_API size_t var_b200_ln_tables_workspace(const var_b200_model_t* m, int n_cls);
VAR_B200_API int var_b200_ln_tables(const var_b200_model_t* m, int block, const float* ada_all, int n_cls, float* u_qkv,
                                    float* v_qkv, float* u_fc1, float* v_fc1, void* work, size_t work_bytes, void* stream);

/* get_logits (var.py:118-124): logits[n_seq*l, V] fp32 = head(LN(x)*(1+scale)+shift). work: blocks workspace. */
VAR_B200_API int var_b200_head_logits(const var_b200_model_t* m, const float* x, const float* ada, int n_seq, int l,
                                      float* logits, void* work, size_t work_bytes, void* stream);
/* Fused head + log-softmax + gather + per-sequence sum (eval_prob.py:441-463) without materialising logits.
 * gt: int32 [gt_rows] with gt_rows dividing n_seq*L; row r uses gt[r % gt_rows] (the token pyramid of the image is
 * shared by every candidate class). scores[n_seq] = sum_{t >= first_pos} log p(gt_t); per_scale: NULL or [n_seq,S];
 * tok_logp: NULL or [n_seq*L]. work: var_b200_score_workspace(). */
VAR_B200_API size_t var_b200_score_workspace(const var_b200_model_t* m, int n_seq, int l);
VAR_B200_API int var_b200_head_score(const var_b200_model_t* m, const float* x, const float* ada, int n_seq, int l,
                                     const int32_t* gt, int gt_rows, int first_pos, float* scores, float* per_scale,
                                     float* tok_logp, void* work, size_t work_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CFG-mixed teacher-forced scoring (var_analysis.py:320-346,437-466).
 * tok_logp[s,t] = log_softmax((1+t_row[t])*logits_cond[s,t,:] - t_row[t]*logits_uncond[t,:])[gt[t]],
 * t_row[t] = cfg * level(t)/(S-1) (device fp32 [L]); two rounded products and a subtraction as the reference.
 * var_b200_scale_sums: per_scale[s,i] = sum of tok_logp over level i (t >= first_pos), total[s] = sum over levels. */
VAR_B200_API int var_b200_cfg_token_logprob(const float* logits_cond, const float* logits_uncond, const int32_t* gt,
                                            const float* t_row, int n_seq, int L, int V, float* tok_logp, void* stream);
/* Token selection of VAR.smooth_sampling (models/var.py:483-536): logits fp32 [2B, l, V] (cond rows, then uncond) are
 * CFG-mixed with guidance t; per row the token is the arg-max of log_softmax(mixed) among the codebook neighbours of
 * the ground-truth token gt[B*l] (neighbors: device int32 [V, n_nb] = argsort(dists, 1)[:, :n_nb]; dists: fp32 [V, V]).
 * Count mode (thr_mode 0): the first cand_count neighbours; threshold mode: those with distance <= d_0 + (thr - d_0) *
 * ratio. Outputs: idx_out int64 [B*l], logp_out (selected log-probability), dlogp_out (log_softmax(-d) of the selected
 * candidate over all n_nb candidates). */
VAR_B200_API int var_b200_neighbor_select(const float* logits, int B, int l, int V, double t, const int32_t* gt,
                                          const int32_t* neighbors, const float* dists, int n_nb, int cand_count,
                                          int thr_mode, float thr, float ratio, void* idx_out, float* logp_out,
                                          float* dlogp_out, void* stream);

/* Expected codebook distance score (var_analysis.py:468-490, --mode l2_dist): tok_dist[s,t] =
 * sum_v p_v * dists[gt[t], v] with p = softmax of the mixed logits (logits_uncond NULL: no mixing), optionally
 * restricted to the top_k most probable tokens and renormalised (:476-486). dists: device fp32 [V, V] =
 * torch.cdist(codebook, codebook) (:256). The caller negates and sums per scale (var_b200_scale_sums). */
VAR_B200_API int var_b200_cfg_token_expected_dist(const float* logits_cond, const float* logits_uncond, const int32_t* gt,
                                                  const float* t_row, const float* dists, int n_seq, int L, int V,
                                                  int top_k, float* tok_dist, void* stream);
VAR_B200_API int var_b200_scale_sums(const float* tok_logp, int n_seq, int L, int n_scales, const int* level_end /* host */,
                                     int first_pos, float* per_scale, float* total, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3 / stride 1 / zero-pad 1 convolution of the VQVAE CNN (models/basic_vae.py:45-46,52-59,130,159,195,225) as an
 * implicit GEMM on the tcgen05 GEMM kernels: the nine taps are nine shifted TMA boxes of the NHWC bf16 activations
 * (out-of-image rows/columns are zero-filled by TMA). x: bf16 [B,H,W,Cin]; w_packed: bf16 [Cout, 9, kpt*64] with
 * kpt = ceil(Cin/64), w_packed[co, ky*3+kx, ci] = weight[co, ci, ky, kx], channels >= Cin zero; bias: fp32 [Cout] or
 * NULL; resid: NULL or bf16 [B,H,W,Cout] added to the bf16-rounded result (ResnetBlock shortcut, basic_vae.py:60);
 * out: bf16 [B,H,W,Cout]. Requires Cout % 32 == 0, Cin % 8 == 0, W | 128 or 128 | W, (H*W) % 128 == 0. */
VAR_B200_API int var_b200_conv3x3_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid,
                                       void* out, int B, int H, int W, int Cin, int Cout, void* stream);

/* Downsample2x of the VQVAE encoder (models/basic_vae.py:31-37): zero pad (0,1,0,1) then 3x3 / stride 2 / no padding, as
 * the same implicit GEMM with a traversal stride of 2 in the TMA box. x: bf16 [B,H,W,Cin] -> out: bf16 [B,H/2,W/2,Cout].
 * Requires H, W even, W/2 | 128 or W/2 == 128, (H/2 * W/2) % 128 == 0, Cout % 32 == 0, Cin % 8 == 0. */
VAR_B200_API int var_b200_conv3x3_s2_nhwc(const void* x, const void* w_packed, const float* bias, void* out, int B, int H,
                                          int W, int Cin, int Cout, void* stream);
/* 1x1 convolution (basic_vae.py:47 nin_shortcut, :69-71 AttnBlock, vqvae.py:48-49 quant_conv / post_quant_conv) = GEMM
 * over the pixels. x: bf16 [n_pixels, Cin]; w_packed: bf16 [Cout, ceil(Cin/64)*64] (columns >= Cin zero); resid: NULL or
 * bf16 [n_pixels, Cout]; out: bf16 [n_pixels, Cout]. Requires Cin % 8 == 0, Cout % 32 == 0. */
VAR_B200_API int var_b200_conv1x1_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid, void* out,
                                       long long n_pixels, int Cin, int Cout, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) on NHWC bf16 tensors: glue of the VQVAE CNN decoder/encoder around the cuDNN convolutions
 * (models/basic_vae.py:18-19,57-58,159,225). x, y: bf16 [B, HW, C]; gamma, beta: fp32 [C]. Deterministic. */
VAR_B200_API size_t var_b200_gn_workspace(int B, int HW, int C, int groups);
/* pre_bias: NULL or fp32 [C], added to x before the statistics (folds the bias of the producing convolution). */
VAR_B200_API int var_b200_gn_silu_nhwc(const void* x, const float* pre_bias, const float* gamma, const float* beta, void* y,
                                       int B, int HW, int C, int groups, float eps, int apply_silu, void* work,
                                       size_t work_bytes, void* stream);
/* The same convolution as var_b200_conv3x3_nhwc, whose epilogue also leaves the GroupNorm statistics of its output
 * behind (basic_vae.py:57-58: every GroupNorm of the CNN follows a convolution): gn_sums[b, g] = (sum, sum of squares)
 * over the pixels and the Cout/groups channels of group g of image b, computed from the stored bf16 values in a fixed
 * order. var_b200_gn_apply_nhwc consumes them: no statistics pass over the tensor. Requires H*W % 256 == 0. */
VAR_B200_API size_t var_b200_conv3x3_gn_workspace(int B, int H, int W, int Cout);
VAR_B200_API int var_b200_conv3x3_gn_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid, void* out,
                                          int B, int H, int W, int Cin, int Cout, int groups, float* gn_sums /* [B,groups,2] */,
                                          void* work, size_t work_bytes, void* stream);
VAR_B200_API int var_b200_gn_apply_nhwc(const void* x, const float* gn_sums, const float* gamma, const float* beta, void* y,
                                        int B, int HW, int C, int groups, float eps, int apply_silu, void* stream);
/* out = a (+ bias_a[c]) + b (+ bias_b[c]) on bf16 NHWC tensors of n_pixels x C; b, bias_a, bias_b may be NULL; out may
 * alias a (residual add of ResnetBlock with the convolution biases folded in, basic_vae.py:60). */
VAR_B200_API int var_b200_add_bias_nhwc(const void* a, const float* bias_a, const void* b, const float* bias_b, void* out,
                                        long long n_pixels, int C, void* stream);
/* y[B,2H,2W,C] = nearest-2x(x[B,H,W,C]) (+ bias[c]) (basic_vae.py:22-28). */
VAR_B200_API int var_b200_upsample2x_nhwc(const void* x, const float* bias, void* y, int B, int H, int W, int C, void* stream);

/* AttnBlock.forward of the VQVAE CNN (models/basic_vae.py:63-92) on NHWC bf16 activations, on var_b200's own kernels:
 * out = x + proj_out(softmax(q k^T * C^-0.5) v) with [q,k,v] = qkv(GroupNorm(x)) (single head over the HW pixels of
 * each image). x, out: bf16 [B, HW, C] (out may alias x); w_qkv: bf16 [3C, C] (the 1x1 convolution's weight, rows q, k,
 * v); w_proj: bf16 [C, C]; biases and GroupNorm affine fp32. The q k^T and P v products of all images run as ONE GEMM
 * launch each (block-diagonal batching of the tcgen05 GEMM). Requires HW % 256 == 0, HW <= 4096, C % 64 == 0. */
VAR_B200_API size_t var_b200_vae_attn_workspace(int B, int HW, int C, int groups);
VAR_B200_API int var_b200_vae_attn_block(const void* x, const float* gn_gamma, const float* gn_beta, int groups, float eps,
                                         const void* w_qkv, const float* b_qkv, const void* w_proj, const float* b_proj,
                                         void* out, int B, int HW, int C, void* work, size_t work_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAR_B200_H */
