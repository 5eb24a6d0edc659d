// HBM-bound helper kernels of the VAR block: adaLN LayerNorm-modulate (A-operand producer for the QKV / FC1 / head
// GEMMs), token embedding, class-condition activation, and the likelihood-score reduction.
#include "elementwise.h"

#include "common.cuh"
#include "host.h"

namespace vb {

// ------------------------------------------------------------------------------------------------
// LN(x) * (1 + scale[seq]) + shift[seq] -> bf16        (basic_var.py:157-158,174; LayerNorm without affine)
// one warp per row; the row lives in registers between the statistics and the normalisation pass
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 20;   // float4 per lane: supports C <= 2560
constexpr int LN_ROWS = 32;   // rows of one sequence per CTA (8 warps x 4 rows)

// A CTA owns LN_ROWS consecutive rows of ONE sequence, so (1+scale) and shift are staged in shared memory once and
// the only global traffic per row is the 4C-byte read and the 2C-byte write (the first version re-read 8C bytes of
// adaLN parameters per row through L2 and ran at 2.3 TB/s).
template <int LN_NV>  // float4 per lane held in registers: C <= 128 * LN_NV
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   int ada_ld, int rows_per_seq, __nv_bfloat16* __restrict__ out, int C, float eps) {
  extern __shared__ float4 ln_sm[];  // [C/4] (1+scale), [C/4] shift
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.y;
  const int nvec = C >> 2;
  float4* s1 = ln_sm;
  float4* sh = ln_sm + nvec;
  {
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)seq * ada_ld);
    const float4* sf = reinterpret_cast<const float4*>(shift + (size_t)seq * ada_ld);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 a = __ldg(sc + i);
      s1[i] = make_float4(1.f + a.x, 1.f + a.y, 1.f + a.z, 1.f + a.w);
      sh[i] = __ldg(sf + i);
    }
  }
  __syncthreads();
  const int t_end = min(rows_per_seq, (int)(blockIdx.x + 1) * LN_ROWS);
  for (int t = blockIdx.x * LN_ROWS + warp; t < t_end; t += 8) {
    const size_t row = (size_t)seq * rows_per_seq + t;
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[LN_NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        v[i] = xr[idx];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(sq) / (float)C + eps);
    uint2* o = reinterpret_cast<uint2*>(out + row * C);
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        const float4 s = s1[idx], h = sh[idx];
        const float a = (v[i].x - mean) * rstd * s.x + h.x;
        const float b = (v[i].y - mean) * rstd * s.y + h.y;
        const float c = (v[i].z - mean) * rstd * s.z + h.z;
        const float d = (v[i].w - mean) * rstd * s.w + h.w;
        o[idx] = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
      }
    }
  }
}

int ln_modulate(const float* x, const float* scale, const float* shift, int ada_ld, int rows_per_seq, void* out, int M,
                int C, float eps, cudaStream_t st) {
  VB_REQUIRE(x && scale && shift && out, "ln_modulate: null pointer");
  VB_REQUIRE(C % 4 == 0 && C <= LN_MAXV * 128, "ln_modulate: C=%d unsupported", C);
  VB_REQUIRE(M > 0 && rows_per_seq > 0 && M % rows_per_seq == 0, "ln_modulate: bad M=%d rows_per_seq=%d", M, rows_per_seq);
  VB_REQUIRE(ada_ld % 4 == 0 && M / rows_per_seq <= 65535, "ln_modulate: bad ada_ld=%d or too many sequences", ada_ld);
  dim3 grid((rows_per_seq + LN_ROWS - 1) / LN_ROWS, M / rows_per_seq);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  vb::ProfScope prof_scope(vb::PK_LN, st);
  if (C <= 1024)
    ln_modulate_kernel<8><<<grid, 256, (size_t)C * 8, st>>>(x, scale, shift, ada_ld, rows_per_seq, o, C, eps);
  else if (C <= 2048)
    ln_modulate_kernel<16><<<grid, 256, (size_t)C * 8, st>>>(x, scale, shift, ada_ld, rows_per_seq, o, C, eps);
  else
    ln_modulate_kernel<LN_MAXV><<<grid, 256, (size_t)C * 8, st>>>(x, scale, shift, ada_ld, rows_per_seq, o, C, eps);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// A operands of the four table GEMMs of var_b200_ln_tables: bf16(1 + scale1), bf16(shift1), bf16(1 + scale2), bf16(shift2)
// ------------------------------------------------------------------------------------------------
__global__ void ln_table_inputs_kernel(const float* __restrict__ s1, const float* __restrict__ h1, const float* __restrict__ s2,
                                       const float* __restrict__ h2, int ada_ld, __nv_bfloat16* __restrict__ out, int n, int C) {
  const int r = blockIdx.x;
  const size_t plane = (size_t)n * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t src = (size_t)r * ada_ld + c, dst = (size_t)r * C + c;
    out[dst] = __float2bfloat16(1.f + s1[src]);
    out[plane + dst] = __float2bfloat16(h1[src]);
    out[2 * plane + dst] = __float2bfloat16(1.f + s2[src]);
    out[3 * plane + dst] = __float2bfloat16(h2[src]);
  }
}

int ln_table_inputs(const float* s1, const float* h1, const float* s2, const float* h2, int ada_ld, void* out, int n, int C,
                    cudaStream_t st) {
  VB_REQUIRE(s1 && h1 && s2 && h2 && out && n > 0 && C > 0, "ln_table_inputs: bad arguments");
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  ln_table_inputs_kernel<<<n, 256, 0, st>>>(s1, h1, s2, h2, ada_ld, reinterpret_cast<__nv_bfloat16*>(out), n, C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// cond = class_emb[label] ; A = bf16(SiLU(cond))      (basic_var.py:146-147 input of every ada_lin)
// ------------------------------------------------------------------------------------------------
__global__ void cond_silu_kernel(const float* __restrict__ class_emb, const int* __restrict__ labels,
                                 __nv_bfloat16* __restrict__ out, int n_seq, int C) {
  const int s = blockIdx.x;
  const float* e = class_emb + (size_t)__ldg(labels + s) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = e[c];
    out[(size_t)s * C + c] = __float2bfloat16(v / (1.f + expf(-v)));
  }
}

int cond_silu(const float* class_emb, const int* labels, void* out, int n_seq, int C, cudaStream_t st) {
  VB_REQUIRE(class_emb && labels && out && n_seq > 0, "cond_silu: bad arguments");
  vb::ProfScope prof_scope(vb::PK_COND, st);
  cond_silu_kernel<<<n_seq, 256, 0, st>>>(class_emb, labels, reinterpret_cast<__nv_bfloat16*>(out), n_seq, C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// shared_aln (models/var.py:15-18,80; basic_var.py:153-154): ada[s, blk*6C + j] = gss[blk, j] + shared[s, j]
__global__ void expand_shared_aln_kernel(const float* __restrict__ shared, int shared_ld, const float* __restrict__ gss,
                                         float* __restrict__ ada, int ada_ld, int depth, int sixC) {
  const int s = blockIdx.y, blk = blockIdx.x;
  for (int j = threadIdx.x; j < sixC; j += blockDim.x)
    ada[(size_t)s * ada_ld + (size_t)blk * sixC + j] = gss[(size_t)blk * sixC + j] + shared[(size_t)s * shared_ld + j];
}

int expand_shared_aln(const float* shared, int shared_ld, const float* gss, float* ada, int ada_ld, int depth, int C,
                      int n_seq, cudaStream_t st) {
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  expand_shared_aln_kernel<<<dim3(depth, n_seq), 256, 0, st>>>(shared, shared_ld, gss, ada, ada_ld, depth, 6 * C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// token embedding (models/var.py:200-207 teacher forced; :153-154,185-187 autoregressive)
//   row t <  first_rows : class_emb[label] + pos_start[t] + lvl_pos[pos0 + t]
//   row t >= first_rows : word_embed(x_in[seq % n_x, t - first_rows, :]) + lvl_pos[pos0 + t]
// ------------------------------------------------------------------------------------------------
constexpr int EMB_ROWS = 32;  // rows per CTA

// CTA = 256 channels x EMB_ROWS token positions of ONE input sequence. The n_seq output sequences share n_x input
// sequences (sequence s reads input s % n_x: the cond / uncond halves of a CFG batch share the sampled tokens, the 1000
// class hypotheses of likelihood scoring all share the image's tokens), so word_embed(x) + position / level embedding
// is computed once per (input sequence, token, channel) and stored to every output sequence that uses it; only the
// class-dependent first rows are per sequence. Each thread keeps its channel's 32 word_embed weights in registers
// (w_word is stored transposed [Cvae, C] so the load is coalesced); x_in rows are staged in shared memory.
__global__ void __launch_bounds__(256)
embed_kernel(const float* __restrict__ x_in, int n_x, int l_in, const int* __restrict__ labels,
             const float* __restrict__ class_emb, const float* __restrict__ pos_start,
             const float* __restrict__ lvl_pos, const float* __restrict__ w_word_t, const float* __restrict__ b_word,
             float* __restrict__ out, int n_seq, int l, int first_rows, int pos0, int C) {
  __shared__ __align__(16) float xin[EMB_ROWS][32];
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int xs = blockIdx.z;                 // input sequence
  const int t0 = blockIdx.y * EMB_ROWS;      // first token position of this CTA
  const int nr = min(EMB_ROWS, l - t0);
  for (int i = threadIdx.x; i < nr * 32; i += 256) {
    const int t = t0 + (i >> 5), k = i & 31;
    xin[i >> 5][k] = (t >= first_rows) ? x_in[((size_t)xs * l_in + (t - first_rows)) * 32 + k] : 0.f;
  }
  __syncthreads();
  if (c >= C) return;
  float w[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) w[k] = __ldg(w_word_t + (size_t)k * C + c);
  const float bw = __ldg(b_word + c);
  for (int i = 0; i < nr; ++i) {
    const int t = t0 + i;
    const float lp = __ldg(lvl_pos + (size_t)(pos0 + t) * C + c);
    if (t < first_rows) {  // class token rows: per output sequence
      const float ps = __ldg(pos_start + (size_t)t * C + c);
      for (int s = xs; s < n_seq; s += n_x)
        out[((size_t)s * l + t) * C + c] = (__ldg(class_emb + (size_t)__ldg(labels + s) * C + c) + ps) + lp;
    } else {
      // the row's 32 inputs as eight 16-byte broadcast loads
      const float4* xr = reinterpret_cast<const float4*>(xin[i]);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 xv = xr[k];
        acc = fmaf(xv.x, w[4 * k], acc);
        acc = fmaf(xv.y, w[4 * k + 1], acc);
        acc = fmaf(xv.z, w[4 * k + 2], acc);
        acc = fmaf(xv.w, w[4 * k + 3], acc);
      }
      const float v = (acc + bw) + lp;
      for (int s = xs; s < n_seq; s += n_x) out[((size_t)s * l + t) * C + c] = v;
    }
  }
}

int embed_tokens(const float* x_in, int n_x, int l_in, const int* labels, const float* class_emb,
                 const float* pos_start, const float* lvl_pos, const float* w_word_t, const float* b_word, float* out,
                 int n_seq, int l, int first_rows, int pos0, int C, int Cv, cudaStream_t st) {
  VB_REQUIRE(out && labels && class_emb && pos_start && lvl_pos && w_word_t && b_word, "embed: null pointer");
  VB_REQUIRE(Cv == 32, "embed: Cvae=%d unsupported (kernel is specialised for 32)", Cv);
  VB_REQUIRE(n_seq > 0 && l > 0, "embed: bad shape n_seq=%d l=%d", n_seq, l);
  VB_REQUIRE(first_rows >= l || (x_in && n_x > 0), "embed: x_in required");
  // without token input (only class rows) every output sequence is its own "input sequence"
  const int nx = (x_in && n_x > 0 && n_x <= n_seq) ? n_x : n_seq;
  dim3 grid((C + 255) / 256, (l + EMB_ROWS - 1) / EMB_ROWS, nx);
  VB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "embed: too many rows / sequences (l=%d, n_x=%d)", l, nx);
  vb::ProfScope prof_scope(vb::PK_EMBED, st);
  embed_kernel<<<grid, 256, 0, st>>>(x_in, nx, l_in, labels, class_emb, pos_start, lvl_pos, w_word_t, b_word, out, n_seq, l,
                                     first_rows, pos0, C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// likelihood score (eval_prob.py:446-463): log p(gt_t) = gt_logit_t - logsumexp_t ; summed per pyramid level.
// part[row, tile] = (max, sum exp(x - max)) from the head GEMM's EPI_SCORE epilogue.
// One CTA per sequence, fixed reduction order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_finalize_kernel(const float2* __restrict__ part, int n_tiles, const float* __restrict__ gt_logit, int L,
                      AttnLevelsPOD lv, float* __restrict__ tok_logp, float* __restrict__ per_scale,
                      float* __restrict__ total, int first_pos) {
  __shared__ float red[VB_MAX_SCALES][8];
  const int s = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[VB_MAX_SCALES];
#pragma unroll
  for (int i = 0; i < VB_MAX_SCALES; ++i) acc[i] = 0.f;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const size_t row = (size_t)s * L + t;
    const float2* p = part + row * n_tiles;
    float m = -INFINITY;
    for (int i = 0; i < n_tiles; ++i) m = fmaxf(m, p[i].x);
    float sum = 0.f;
    for (int i = 0; i < n_tiles; ++i) sum += p[i].y * expf(p[i].x - m);
    const float lp = gt_logit[row] - (m + logf(sum));
    if (tok_logp) tok_logp[row] = lp;
    int level = 0;
    while (level < lv.n - 1 && t >= lv.end[level]) ++level;
#pragma unroll
    for (int i = 0; i < VB_MAX_SCALES; ++i)
      if (i == level && t >= first_pos) acc[i] += lp;
  }
#pragma unroll
  for (int i = 0; i < VB_MAX_SCALES; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < lv.n; ++i) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[i][w];
      if (per_scale) per_scale[(size_t)s * lv.n + i] = v;
      tot += v;
    }
    total[s] = tot;
  }
}

int score_finalize(const void* part, int n_tiles, const float* gt_logit, int n_seq, int L, int n_scales,
                   const int* level_end, float* tok_logp, float* per_scale, float* total, int first_pos,
                   cudaStream_t st) {
  VB_REQUIRE(part && gt_logit && total && n_seq > 0 && L > 0, "score_finalize: bad arguments");
  VB_REQUIRE(n_scales > 0 && n_scales <= VB_MAX_SCALES, "score_finalize: n_scales=%d", n_scales);
  AttnLevelsPOD lv;
  lv.n = n_scales;
  for (int i = 0; i < VB_MAX_SCALES; ++i) lv.end[i] = level_end[i < n_scales ? i : n_scales - 1];
  vb::ProfScope prof_scope(vb::PK_SCORE_FIN, st);
  score_finalize_kernel<<<n_seq, 256, 0, st>>>(reinterpret_cast<const float2*>(part), n_tiles, gt_logit, L, lv, tok_logp,
                                               per_scale, total, first_pos);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// CFG-mixed teacher-forced scoring (var_analysis.py:320-346): x = (1+t)*cond - t*uncond with t = cfg * si/(S-1) per
// position, then log_softmax(x)[gt]. One CTA per (class sequence, position) row of V logits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cfg_token_logprob_kernel(const float* __restrict__ lc, const float* __restrict__ lu, const int* __restrict__ gt,
                         const float* __restrict__ t_row, int L, int V, float* __restrict__ tok_logp) {
  __shared__ float red[8];
  const int t = blockIdx.x, s = blockIdx.y;
  const float* c = lc + ((size_t)s * L + t) * V;
  const float* u = lu + (size_t)t * V;
  const float tr = __ldg(t_row + t), opt = 1.f + tr;
  const int g = __ldg(gt + t);
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) m = fmaxf(m, __fsub_rn(__fmul_rn(opt, c[v]), __fmul_rn(tr, u[v])));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) sum += expf(__fsub_rn(__fmul_rn(opt, c[v]), __fmul_rn(tr, u[v])) - m);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float xg = __fsub_rn(__fmul_rn(opt, c[g]), __fmul_rn(tr, u[g]));
    tok_logp[(size_t)s * L + t] = xg - (m + logf(tot));
  }
}

int cfg_token_logprob(const float* lc, const float* lu, const int* gt, const float* t_row, int n_seq, int L, int V,
                      float* tok_logp, cudaStream_t st) {
  VB_REQUIRE(lc && lu && gt && t_row && tok_logp && n_seq > 0 && L > 0 && V > 0, "cfg_token_logprob: bad arguments");
  VB_REQUIRE(n_seq <= 65535, "cfg_token_logprob: too many sequences");
  vb::ProfScope prof_scope(vb::PK_SCORE_FIN, st);
  cfg_token_logprob_kernel<<<dim3(L, n_seq), 256, 0, st>>>(lc, lu, gt, t_row, L, V, tok_logp);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// per-scale / total sums of token log-likelihoods (var_analysis.py:437-466), fixed reduction order
__global__ void __launch_bounds__(256)
scale_sums_kernel(const float* __restrict__ tok_logp, int L, AttnLevelsPOD lv, int first_pos, float* __restrict__ per_scale,
                  float* __restrict__ total) {
  __shared__ float red[VB_MAX_SCALES][8];
  const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[VB_MAX_SCALES];
#pragma unroll
  for (int i = 0; i < VB_MAX_SCALES; ++i) acc[i] = 0.f;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const float lp = tok_logp[(size_t)s * L + t];
    int level = 0;
    while (level < lv.n - 1 && t >= lv.end[level]) ++level;
#pragma unroll
    for (int i = 0; i < VB_MAX_SCALES; ++i)
      if (i == level && t >= first_pos) acc[i] += lp;
  }
#pragma unroll
  for (int i = 0; i < VB_MAX_SCALES; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < lv.n; ++i) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[i][w];
      if (per_scale) per_scale[(size_t)s * lv.n + i] = v;
      tot += v;
    }
    total[s] = tot;
  }
}

int scale_sums(const float* tok_logp, int n_seq, int L, int n_scales, const int* level_end, int first_pos,
               float* per_scale, float* total, cudaStream_t st) {
  VB_REQUIRE(tok_logp && total && n_seq > 0 && L > 0, "scale_sums: bad arguments");
  VB_REQUIRE(n_scales > 0 && n_scales <= VB_MAX_SCALES, "scale_sums: n_scales=%d", n_scales);
  AttnLevelsPOD lv;
  lv.n = n_scales;
  for (int i = 0; i < VB_MAX_SCALES; ++i) lv.end[i] = level_end[i < n_scales ? i : n_scales - 1];
  vb::ProfScope prof_scope(vb::PK_SCORE_FIN, st);
  scale_sums_kernel<<<n_seq, 256, 0, st>>>(tok_logp, L, lv, first_pos, per_scale, total);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
