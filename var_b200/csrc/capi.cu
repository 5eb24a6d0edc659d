// Model-level C-ABI (include/var_b200.h): adaLN parameter table, embedding, transformer blocks, head / fused
// likelihood score, quantizer, sampler. Pure orchestration: every heavy step is one of the sm_100a kernels.
#include "../../include/var_b200.h"

#include "attn.h"
#include "common.cuh"
#include "elementwise.h"
#include "gemm.h"
#include "host.h"
#include "quant.h"
#include "sampler.h"

#include <algorithm>

using namespace vb;

namespace {

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0, cap;
  Carver(void* p, size_t c) : base(reinterpret_cast<uint8_t*>(p)), cap(c) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off = align_up(off + bytes);
    return r;
  }
};

inline int model_L(const var_b200_model_t* m) {
  int L = 0;
  for (int i = 0; i < m->n_scales; ++i) L += m->patch_nums[i] * m->patch_nums[i];
  return L;
}

int check_model(const var_b200_model_t* m) {
  VB_REQUIRE(m != nullptr, "null model");
  VB_REQUIRE(m->depth > 0 && m->C > 0 && m->H * 64 == m->C, "model: depth=%d C=%d H=%d (head_dim must be 64)", m->depth,
             m->C, m->H);
  VB_REQUIRE(m->n_scales > 0 && m->n_scales <= VAR_B200_MAX_SCALES, "model: n_scales=%d", m->n_scales);
  VB_REQUIRE(m->V % 64 == 0 && m->C % 64 == 0, "model: V=%d and C=%d must be multiples of 64", m->V, m->C);
  VB_REQUIRE(m->blocks && m->w_ada && m->b_ada && m->w_head && m->b_head && m->w_word && m->b_word && m->class_emb &&
                 m->pos_start && m->lvl_pos,
             "model: null weight pointer");
  const int want = m->shared_aln ? 8 * m->C : (6 * m->depth + 2) * m->C;
  VB_REQUIRE(m->ada_rows == want, "model: ada_rows=%d, expected %d", m->ada_rows, want);
  VB_REQUIRE(!m->shared_aln || m->ada_gss, "model: shared_aln needs ada_gss");
  return VB_OK;
}

void level_ends(const var_b200_model_t* m, int* out) {
  int c = 0;
  for (int i = 0; i < m->n_scales; ++i) {
    c += m->patch_nums[i] * m->patch_nums[i];
    out[i] = c;
  }
}

// workspace layout shared by blocks / head / score
struct BlockWs {
  void *a, *q, *kvscratch, *h, *part, *gtl, *lnp;
  size_t bytes;
};
BlockWs carve_blocks(const var_b200_model_t* m, int n_seq, int l, void* work, size_t cap, bool score) {
  const size_t M = (size_t)n_seq * l, C = m->C;
  Carver cv(work, cap);
  BlockWs w{};
  w.a = cv.take(M * C * 2);        // LN-modulated activations (bf16), reused as attention output
  w.q = cv.take(M * C * 2);        // q (bf16)
  w.h = cv.take(M * 4 * C * 2);    // FFN hidden (bf16)
  w.lnp = cv.take(M * (size_t)gemm_ln_parts((int)M, m->C) * 8);  // deferred LayerNorm: partial (sum, sumsq) per row
  if (score) {
    const int nt = (m->V + gemm_pick_bn(m->V) - 1) / gemm_pick_bn(m->V);
    w.part = cv.take(M * nt * GEMM_EPI_SUB * 8);  // one (max, sumexp) per row, tile and epilogue sub-warp
    w.gtl = cv.take(M * 4);
  }
  w.bytes = cv.off;
  return w;
}

}  // namespace

extern "C" int var_b200_ada_ld(const var_b200_model_t* m) { return m ? (6 * m->depth + 2) * m->C : 0; }

extern "C" size_t var_b200_ada_workspace(const var_b200_model_t* m, int n_seq) {
  if (!m || n_seq <= 0) return 0;
  return align_up((size_t)n_seq * m->C * 2) + (m->shared_aln ? align_up((size_t)n_seq * 8 * m->C * 4) : 0);
}

extern "C" int var_b200_ada_params(const var_b200_model_t* m, const int32_t* labels, int n_seq, float* ada_out, void* work,
                                   size_t work_bytes, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(labels && ada_out && work && n_seq > 0, "ada_params: bad arguments");
  VB_REQUIRE(work_bytes >= var_b200_ada_workspace(m, n_seq), "ada_params: workspace %zu < %zu", work_bytes,
             var_b200_ada_workspace(m, n_seq));
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(work, work_bytes);
  void* cond = cv.take((size_t)n_seq * m->C * 2);
  rc = cond_silu(m->class_emb, labels, cond, n_seq, m->C, st);
  if (rc) return rc;
  const int ld = var_b200_ada_ld(m);
  GemmParams p{};
  p.M = n_seq; p.K = m->C; p.bias = m->b_ada;
  if (!m->shared_aln) {
    p.N = m->ada_rows;
    p.out = ada_out;
    return gemm_launch(cond, m->w_ada, p, EPI_BIAS_F32, st);
  }
  float* tmp = reinterpret_cast<float*>(cv.take((size_t)n_seq * 8 * m->C * 4));
  p.N = 8 * m->C;
  p.out = tmp;
  rc = gemm_launch(cond, m->w_ada, p, EPI_BIAS_F32, st);
  if (rc) return rc;
  rc = expand_shared_aln(tmp, 8 * m->C, m->ada_gss, ada_out, ld, m->depth, m->C, n_seq, st);
  if (rc) return rc;
  VB_CUDA_CHECK(cudaMemcpy2DAsync(ada_out + (size_t)6 * m->depth * m->C, (size_t)ld * 4, tmp + 6 * m->C,
                                  (size_t)8 * m->C * 4, (size_t)2 * m->C * 4, n_seq, cudaMemcpyDeviceToDevice, st));
  return VB_OK;
}

extern "C" int var_b200_embed(const var_b200_model_t* m, const float* x_in, int n_x, int l_in, const int32_t* labels,
                              int n_seq, int l, int first_rows, int pos0, float* x_out, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(pos0 >= 0 && pos0 + l <= model_L(m), "embed: positions [%d,%d) outside the sequence", pos0, pos0 + l);
  VB_REQUIRE(first_rows <= m->patch_nums[0] * m->patch_nums[0] || first_rows == 0, "embed: first_rows=%d", first_rows);
  return embed_tokens(x_in, n_x, l_in, labels, m->class_emb, m->pos_start, m->lvl_pos, m->w_word, m->b_word, x_out, n_seq,
                      l, first_rows, pos0, m->C, m->Cvae, (cudaStream_t)stream);
}

extern "C" size_t var_b200_blocks_workspace(const var_b200_model_t* m, int n_seq, int l) {
  if (!m || n_seq <= 0 || l <= 0) return 0;
  return carve_blocks(m, n_seq, l, nullptr, 0, false).bytes;
}
extern "C" size_t var_b200_score_workspace(const var_b200_model_t* m, int n_seq, int l) {
  if (!m || n_seq <= 0 || l <= 0) return 0;
  return carve_blocks(m, n_seq, l, nullptr, 0, true).bytes;
}

extern "C" size_t var_b200_ln_tables_workspace(const var_b200_model_t* m, int n_cls) {
  if (!m || n_cls <= 0) return 0;
  return align_up((size_t)4 * n_cls * m->C * 2);
}

extern "C" int var_b200_ln_tables(const var_b200_model_t* m, int block, const float* ada_all, int n_cls, float* u_qkv,
                                  float* v_qkv, float* u_fc1, float* v_fc1, void* work, size_t work_bytes, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(block >= 0 && block < m->depth && ada_all && n_cls > 0 && u_qkv && v_qkv && u_fc1 && v_fc1 && work,
             "ln_tables: bad arguments (block=%d n_cls=%d)", block, n_cls);
  VB_REQUIRE(work_bytes >= var_b200_ln_tables_workspace(m, n_cls), "ln_tables: workspace %zu < %zu", work_bytes,
             var_b200_ln_tables_workspace(m, n_cls));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = m->C, ld = var_b200_ada_ld(m);
  const float* ab = ada_all + (size_t)6 * block * C;  // gamma1,gamma2,scale1,scale2,shift1,shift2
  __nv_bfloat16* in = reinterpret_cast<__nv_bfloat16*>(work);
  rc = ln_table_inputs(ab + 2 * C, ab + 4 * C, ab + 3 * C, ab + 5 * C, ld, in, n_cls, C, st);
  if (rc) return rc;
  const var_b200_block_weights_t& bw = m->blocks[block];
  const size_t plane = (size_t)n_cls * C;
  struct { const void* A; const void* W; int N; const float* bias; float* out; } g[4] = {
      {in, bw.w_qkv, 3 * C, nullptr, u_qkv},
      {in + plane, bw.w_qkv, 3 * C, bw.b_qkv, v_qkv},
      {in + 2 * plane, bw.w_fc1, 4 * C, nullptr, u_fc1},
      {in + 3 * plane, bw.w_fc1, 4 * C, bw.b_fc1, v_fc1}};
  for (auto& t : g) {
    GemmParams p{};
    p.M = n_cls; p.N = t.N; p.K = C; p.bias = t.bias; p.out = t.out;
    rc = gemm_launch(t.A, t.W, p, EPI_BIAS_F32, st);
    if (rc) return rc;
  }
  return VB_OK;
}

extern "C" int var_b200_blocks(const var_b200_model_t* m, float* x, const float* ada, const int32_t* labels, int n_seq, int l,
                               int pos0, void* kv, size_t kv_layer_stride, int Lmax, float* x_dump, void* work,
                               size_t work_bytes, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(x && ada && kv && work && n_seq > 0 && l > 0, "blocks: bad arguments");
  const BlockWs w = carve_blocks(m, n_seq, l, work, work_bytes, false);
  VB_REQUIRE(work_bytes >= w.bytes, "blocks: workspace %zu < %zu", work_bytes, w.bytes);
  VB_REQUIRE(pos0 >= 0 && pos0 + l <= Lmax, "blocks: cache overflow pos0=%d l=%d Lmax=%d", pos0, l, Lmax);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = m->C, M = n_seq * l, ld = var_b200_ada_ld(m);
  const size_t cache_elems = (size_t)n_seq * m->H * Lmax * 64;
  AttnArgs at{};
  at.n_seq = n_seq; at.H = m->H; at.Lq = l; at.Lmax = Lmax; at.q_pos0 = pos0; at.n_scales = m->n_scales;
  at.max_score = m->attn_max_score;
  at.q_log2 = m->attn_q_log2;
  VB_REQUIRE(!m->attn_no_l2norm || (m->attn_max_score == 0.f && !m->attn_q_log2),
             "blocks: attn_no_l2norm needs attn_max_score = 0 and attn_q_log2 = 0 (unbounded scores: general attention kernel)");
  level_ends(m, at.level_end);
  VB_REQUIRE(pos0 + l <= at.level_end[m->n_scales - 1], "blocks: positions beyond the pyramid");
  // Deferred LayerNorm (gemm_sm100.cuh, LNF): possible when the caller names the class of every sequence and every block
  // carries its per-class tables. Buffers: w.a = LN1 operand / attention output, w.q = q, then the LN2 operand (q is
  // dead once the attention has run), w.lnp = the row statistics' partial sums (one consumer at a time).
  bool fused = labels != nullptr;
  for (int i = 0; i < m->depth && fused; ++i) {
    const var_b200_block_weights_t& bw = m->blocks[i];
    fused = bw.u_qkv && bw.v_qkv && bw.u_fc1 && bw.v_fc1;
  }
  const int ln_parts = gemm_ln_parts(M, C);
  auto consume = [&](GemmParams& p, const float* u, const float* v) {
    p.ln_part_in = reinterpret_cast<const float2*>(w.lnp); p.ln_parts = ln_parts; p.ln_C = C; p.ln_eps = m->norm_eps;
    p.ln_u = u; p.ln_v = v; p.ln_labels = labels; p.rows_per_seq = l;
  };
  for (int i = 0; i < m->depth; ++i) {
    const var_b200_block_weights_t& bw = m->blocks[i];
    const float* ab = ada + (size_t)6 * i * C;  // gamma1,gamma2,scale1,scale2,shift1,shift2
    __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(kv) + (size_t)i * kv_layer_stride;
    __nv_bfloat16* vc = kc + cache_elems;
    // x += gamma1 * proj(attn(LN(x)(1+scale1)+shift1))
    const bool ln1_deferred = fused && i > 0;  // the previous block's fc2 epilogue left x*(1+scale1) in w.a
    if (!ln1_deferred) {
      rc = ln_modulate(x, ab + 2 * C, ab + 4 * C, ld, l, w.a, M, C, m->norm_eps, st);
      if (rc) return rc;
    }
    GemmParams p{};
    p.M = M; p.N = 3 * C; p.K = C; p.bias = bw.b_qkv;
    p.q_out = reinterpret_cast<__nv_bfloat16*>(w.q); p.k_cache = kc; p.v_cache = vc; p.q_scale = bw.q_scale;
    p.C = C; p.H = m->H; p.pos0 = pos0; p.Lmax = Lmax; p.rows_per_seq = l; p.no_l2norm = m->attn_no_l2norm;
    if (ln1_deferred) consume(p, bw.u_qkv, bw.v_qkv);
    rc = gemm_launch(w.a, bw.w_qkv, p, EPI_QKV, st);
    if (rc) return rc;
    at.q = w.q; at.k = kc; at.v = vc; at.out = w.a;
    rc = attn_launch(at, st);
    if (rc) return rc;
    p = GemmParams{};
    p.M = M; p.N = C; p.K = C; p.bias = bw.b_proj; p.out = x; p.resid = x; p.gate = ab; p.rows_per_seq = l; p.gate_ld = ld;
    if (fused) {
      p.ln_a_out = reinterpret_cast<__nv_bfloat16*>(w.q); p.ln_scale = ab + 3 * C;
      p.ln_part_out = reinterpret_cast<float2*>(w.lnp);
    }
    rc = gemm_launch(w.a, bw.w_proj, p, EPI_GATE_RESID, st);
    if (rc) return rc;
    // x += gamma2 * fc2(gelu(fc1(LN(x)(1+scale2)+shift2)))
    if (!fused) {
      rc = ln_modulate(x, ab + 3 * C, ab + 5 * C, ld, l, w.a, M, C, m->norm_eps, st);
      if (rc) return rc;
    }
    p = GemmParams{};
    p.M = M; p.N = 4 * C; p.K = C; p.bias = bw.b_fc1; p.out = w.h;
    if (fused) consume(p, bw.u_fc1, bw.v_fc1);
    rc = gemm_launch(fused ? w.q : w.a, bw.w_fc1, p, EPI_GELU_BF16, st);
    if (rc) return rc;
    p = GemmParams{};
    p.M = M; p.N = C; p.K = 4 * C; p.bias = bw.b_fc2; p.out = x; p.resid = x; p.gate = ab + C; p.rows_per_seq = l;
    p.gate_ld = ld;
    if (fused && i + 1 < m->depth) {  // operand and statistics of the NEXT block's first LayerNorm
      p.ln_a_out = reinterpret_cast<__nv_bfloat16*>(w.a); p.ln_scale = ab + 6 * C + 2 * C;
      p.ln_part_out = reinterpret_cast<float2*>(w.lnp);
    }
    rc = gemm_launch(w.h, bw.w_fc2, p, EPI_GATE_RESID, st);
    if (rc) return rc;
    if (x_dump)
      VB_CUDA_CHECK(cudaMemcpyAsync(x_dump + (size_t)i * M * C, x, (size_t)M * C * 4, cudaMemcpyDeviceToDevice, st));
  }
  return VB_OK;
}

extern "C" int var_b200_head_logits(const var_b200_model_t* m, const float* x, const float* ada, int n_seq, int l,
                                    float* logits, void* work, size_t work_bytes, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(x && ada && logits && work, "head_logits: null pointer");
  const BlockWs w = carve_blocks(m, n_seq, l, work, work_bytes, false);
  VB_REQUIRE(work_bytes >= w.bytes, "head_logits: workspace %zu < %zu", work_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = m->C, M = n_seq * l, ld = var_b200_ada_ld(m);
  const float* ah = ada + (size_t)6 * m->depth * C;  // scale, shift
  rc = ln_modulate(x, ah, ah + C, ld, l, w.a, M, C, m->norm_eps, st);
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = m->V; p.K = C; p.bias = m->b_head; p.out = logits;
  return gemm_launch(w.a, m->w_head, p, EPI_BIAS_F32, st);
}

extern "C" int var_b200_head_score(const var_b200_model_t* m, const float* x, const float* ada, int n_seq, int l,
                                   const int32_t* gt, int gt_rows, int first_pos, float* scores, float* per_scale,
                                   float* tok_logp, void* work, size_t work_bytes, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  VB_REQUIRE(x && ada && gt && scores && work, "head_score: null pointer");
  VB_REQUIRE(l == model_L(m), "head_score: needs the full sequence (l=%d, L=%d)", l, model_L(m));
  VB_REQUIRE(gt_rows > 0 && ((size_t)n_seq * l) % gt_rows == 0 && gt_rows % l == 0, "head_score: gt_rows=%d", gt_rows);
  const BlockWs w = carve_blocks(m, n_seq, l, work, work_bytes, true);
  VB_REQUIRE(work_bytes >= w.bytes, "head_score: workspace %zu < %zu", work_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = m->C, M = n_seq * l, ld = var_b200_ada_ld(m);
  const float* ah = ada + (size_t)6 * m->depth * C;
  rc = ln_modulate(x, ah, ah + C, ld, l, w.a, M, C, m->norm_eps, st);
  if (rc) return rc;
  // NaN-fill: a ground-truth token outside [0, V) leaves its row unwritten and the score becomes NaN, never stale data
  VB_CUDA_CHECK(cudaMemsetAsync(w.gtl, 0xFF, (size_t)M * 4, st));
  GemmParams p{};
  p.M = M; p.N = m->V; p.K = C; p.bias = m->b_head;
  p.gt = gt; p.gt_mod = gt_rows; p.part = reinterpret_cast<float2*>(w.part); p.gt_logit = reinterpret_cast<float*>(w.gtl);
  rc = gemm_launch(w.a, m->w_head, p, EPI_SCORE, st);
  if (rc) return rc;
  int ends[VAR_B200_MAX_SCALES];
  level_ends(m, ends);
  const int bn = gemm_pick_bn(m->V);
  return score_finalize(w.part, GEMM_EPI_SUB * ((m->V + bn - 1) / bn), reinterpret_cast<float*>(w.gtl), n_seq, l, m->n_scales, ends,
                        tok_logp, per_scale, scores, first_pos, st);
}

extern "C" int var_b200_attention(const void* q, const void* k, const void* v, void* out, int n_seq, int H, int Lq, int Lmax,
                                  int q_pos0, int n_scales, const int* level_end, float max_score, int q_log2,
                                  void* stream) {
  VB_REQUIRE(level_end && n_scales > 0 && n_scales <= VAR_B200_MAX_SCALES, "attention: bad level table");
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.out = out; a.n_seq = n_seq; a.H = H; a.Lq = Lq; a.Lmax = Lmax; a.q_pos0 = q_pos0;
  a.n_scales = n_scales;
  a.max_score = max_score;
  a.q_log2 = q_log2;
  for (int i = 0; i < n_scales; ++i) a.level_end[i] = level_end[i];
  return attn_launch(a, (cudaStream_t)stream);
}

extern "C" int var_b200_ln_modulate(const float* x, const float* scale, const float* shift, int ada_ld, int rows_per_seq,
                                    void* out_bf16, int M, int C, float eps, void* stream) {
  return ln_modulate(x, scale, shift, ada_ld, rows_per_seq, out_bf16, M, C, eps, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ quantizer
static int fill_quant(const var_b200_quant_t* qz, int B, QuantArgs& a) {
  VB_REQUIRE(qz != nullptr, "quant: null descriptor");
  VB_REQUIRE(qz->n_scales > 0 && qz->n_scales <= VAR_B200_MAX_SCALES, "quant: n_scales=%d", qz->n_scales);
  a = QuantArgs{};
  a.B = B; a.Cvae = qz->Cvae; a.V = qz->V; a.S = qz->n_scales;
  a.H = qz->ph[qz->n_scales - 1]; a.W = qz->pw[qz->n_scales - 1];
  for (int i = 0; i < qz->n_scales; ++i) { a.ph[i] = qz->ph[i]; a.pw[i] = qz->pw[i]; a.phi_of_scale[i] = qz->phi_of_scale[i]; }
  a.n_phi = qz->n_phi; a.resi = qz->resi;
  a.codebook = qz->codebook; a.phi_w = qz->phi_w; a.phi_b = qz->phi_b;
  return VB_OK;
}

static size_t quant_tokens_max(const var_b200_quant_t* qz, int B) {
  size_t m = 0;
  for (int i = 0; i < qz->n_scales; ++i) m = std::max(m, (size_t)B * qz->ph[i] * qz->pw[i]);
  return m;
}

extern "C" size_t var_b200_quant_encode_workspace(const var_b200_quant_t* qz, int B) {
  if (!qz || B <= 0 || qz->n_scales <= 0) return 0;
  const size_t img = (size_t)qz->Cvae * qz->ph[qz->n_scales - 1] * qz->pw[qz->n_scales - 1];
  const size_t nt = quant_tokens_max(qz, B);
  size_t b = 0;
  b += align_up(2 * B * img * 4);          // f_rest, f_hat
  b += align_up(nt * qz->Cvae * 4);        // pooled tokens fp32
  b += align_up(nt * 64 * 2);              // pooled tokens bf16 (K padded to 64)
  b += align_up(nt * 4);                   // |z|^2
  b += align_up((size_t)qz->V * 64 * 2);   // codebook bf16 (K padded to 64)
  b += align_up((size_t)qz->V * 4);        // |e_v|^2
  return b;
}

extern "C" int var_b200_quant_encode(const var_b200_quant_t* qz, const float* f, int B, int64_t* idx_out, float* fhat_list,
                                     void* work, size_t work_bytes, int search_mode, void* stream) {
  QuantArgs a;
  int rc = fill_quant(qz, B, a);
  if (rc) return rc;
  VB_REQUIRE(f && idx_out && work && B > 0, "quant_encode: bad arguments");
  VB_REQUIRE(work_bytes >= var_b200_quant_encode_workspace(qz, B), "quant_encode: workspace %zu < %zu", work_bytes,
             var_b200_quant_encode_workspace(qz, B));
  VB_REQUIRE(search_mode == 0 || search_mode == 1, "quant_encode: search_mode=%d", search_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t img = (size_t)a.Cvae * a.H * a.W;
  const size_t nt = quant_tokens_max(qz, B);
  Carver cv(work, work_bytes);
  float* rest_hat = reinterpret_cast<float*>(cv.take(2 * B * img * 4));
  float* z = reinterpret_cast<float*>(cv.take(nt * a.Cvae * 4));
  void* zb = cv.take(nt * 64 * 2);
  float* zz = reinterpret_cast<float*>(cv.take(nt * 4));
  void* cb16 = cv.take((size_t)a.V * 64 * 2);
  float* ee = reinterpret_cast<float*>(cv.take((size_t)a.V * 4));
  a.f_rest = rest_hat; a.f_hat = rest_hat + (size_t)B * img;
  a.idx_concat = 1; a.idx = idx_out; a.fhat_list = fhat_list;
  if (search_mode == 1) {  // fused single kernel, fp32 CUDA-core search
    a.si_begin = 0; a.si_end = a.S; a.f = f; a.zero_fhat = 1; a.split = 0;
    return quant_launch(a, st);
  }
  // tensor-core search: [pool s0] -> for every scale: [UMMA filter + exact re-rank] -> [update s, pool s+1]
  rc = quant_prepare_codebook(a.codebook, cb16, ee, a.V, st);
  if (rc) return rc;
  a.z_out = z; a.zb_out = zb; a.zz_out = zz;
  a.si_begin = 0; a.si_end = 1; a.f = f; a.zero_fhat = 1; a.split = 1;
  rc = quant_launch(a, st);
  if (rc) return rc;
  a.f = nullptr; a.zero_fhat = 0; a.split = 2;
  size_t off = 0;
  for (int si = 0; si < a.S; ++si) {
    const int n_tok = B * a.ph[si] * a.pw[si];
    QuantSearchArgs sa{};
    sa.zb = zb; sa.z = z; sa.zz = zz; sa.cb_bf16 = cb16; sa.codebook = a.codebook; sa.ee = ee; sa.N = n_tok; sa.V = a.V;
    sa.idx_out = idx_out + off;
    rc = quant_search_launch(sa, st);
    if (rc) return rc;
    a.si_begin = si; a.si_end = si + 1;
    rc = quant_launch(a, st);
    if (rc) return rc;
    off += n_tok;
  }
  return VB_OK;
}

extern "C" int var_b200_quant_decode(const var_b200_quant_t* qz, const int64_t* idx, int B, float* var_input,
                                     float* fhat_list, float* fhat_last, void* stream) {
  QuantArgs a;
  int rc = fill_quant(qz, B, a);
  if (rc) return rc;
  VB_REQUIRE(idx && fhat_last && B > 0, "quant_decode: bad arguments");
  int L = 0;
  for (int i = 0; i < a.S; ++i) L += a.ph[i] * a.pw[i];
  a.si_begin = 0; a.si_end = a.S;
  a.f_hat = fhat_last; a.zero_fhat = 1;
  a.idx_concat = 1; a.idx = const_cast<int64_t*>(idx); a.fhat_list = fhat_list;
  a.next_tokens = var_input; a.next_stride = L - a.ph[0] * a.pw[0];
  return quant_launch(a, (cudaStream_t)stream);
}

extern "C" int var_b200_quant_next_input(const var_b200_quant_t* qz, int si, float* f_hat, const int64_t* idx_si, int B,
                                         float* next_tokens, float* next_nchw, void* stream) {
  QuantArgs a;
  int rc = fill_quant(qz, B, a);
  if (rc) return rc;
  VB_REQUIRE(f_hat && idx_si && B > 0 && si >= 0 && si < a.S, "quant_next_input: bad arguments (si=%d)", si);
  a.si_begin = si; a.si_end = si + 1;
  a.f_hat = f_hat; a.zero_fhat = 0;
  a.idx_concat = 0; a.idx = const_cast<int64_t*>(idx_si);
  a.next_tokens = next_tokens; a.next_nchw = next_nchw;
  a.next_stride = si + 1 < a.S ? a.ph[si + 1] * a.pw[si + 1] : 0;
  return quant_launch(a, (cudaStream_t)stream);
}

extern "C" int var_b200_cfg_topk_sample_smooth(const float* logits, int B, int l, int V, int use_cfg, double t,
                                               const float* q, int top_k, float top_p, int64_t* idx_out, float* mixed_out,
                                               const float* q_gumbel, float tau, float logit_mul, const float* codebook,
                                               int Cvae, float* h_out, void* stream) {
  SampleArgs a{};
  a.logits = logits; a.B = B; a.l = l; a.V = V; a.use_cfg = use_cfg; a.t = t; a.q = q; a.top_k = top_k; a.top_p = top_p;
  a.idx_out = idx_out; a.mixed_out = mixed_out;
  a.q_gumbel = q_gumbel; a.tau = tau; a.logit_mul = logit_mul; a.codebook = codebook; a.Cvae = Cvae; a.h_out = h_out;
  VB_REQUIRE(q_gumbel != nullptr, "cfg_topk_sample_smooth: q_gumbel is required (use var_b200_cfg_topk_sample otherwise)");
  return sample_launch(a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ CFG scoring
extern "C" int var_b200_cfg_token_logprob(const float* logits_cond, const float* logits_uncond, const int32_t* gt,
                                          const float* t_row, int n_seq, int L, int V, float* tok_logp, void* stream) {
  return cfg_token_logprob(logits_cond, logits_uncond, gt, t_row, n_seq, L, V, tok_logp, (cudaStream_t)stream);
}

extern "C" int var_b200_neighbor_select(const float* logits, int B, int l, int V, double t, const int32_t* gt,
                                        const int32_t* neighbors, const float* dists, int n_nb, int cand_count, int thr_mode,
                                        float thr, float ratio, void* idx_out, float* logp_out, float* dlogp_out,
                                        void* stream) {
  return neighbor_select(logits, B, l, V, t, gt, neighbors, dists, n_nb, cand_count, thr_mode, thr, ratio, idx_out, logp_out,
                         dlogp_out, (cudaStream_t)stream);
}

extern "C" int var_b200_cfg_token_expected_dist(const float* logits_cond, const float* logits_uncond, const int32_t* gt,
                                                const float* t_row, const float* dists, int n_seq, int L, int V, int top_k,
                                                float* tok_dist, void* stream) {
  return cfg_token_expected_dist(logits_cond, logits_uncond, gt, t_row, dists, n_seq, L, V, top_k, tok_dist,
                                 (cudaStream_t)stream);
}

extern "C" int var_b200_scale_sums(const float* tok_logp, int n_seq, int L, int n_scales, const int* level_end,
                                   int first_pos, float* per_scale, float* total, void* stream) {
  VB_REQUIRE(level_end != nullptr, "scale_sums: null level table");
  return scale_sums(tok_logp, n_seq, L, n_scales, level_end, first_pos, per_scale, total, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ sampler
extern "C" int var_b200_cfg_topk_sample(const float* logits, int B, int l, int V, int use_cfg, double t, const float* q,
                                        int top_k, float top_p, int64_t* idx_out, float* mixed_out, void* stream) {
  SampleArgs a{};
  a.logits = logits; a.B = B; a.l = l; a.V = V; a.use_cfg = use_cfg; a.t = t; a.q = q; a.top_k = top_k; a.top_p = top_p;
  a.idx_out = idx_out; a.mixed_out = mixed_out;
  return sample_launch(a, (cudaStream_t)stream);
}
