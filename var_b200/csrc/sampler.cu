// Classifier-free-guidance logit mixing fused with top-k (+top-p) filtering and sampling from caller-supplied
// Exp(1) noise. One CTA per (image, position) row of V logits held in shared memory.
//
// Reference semantics: models/var.py:172-175 and models/helpers.py:6-19:
//   x = (1+t)*cond - t*uncond                      (two rounded products and a subtraction, no FMA)
//   top-k: thr = k-th largest; x < thr -> -inf      (ties with thr are kept)
//   top-p: ascending sort, softmax, cumsum <= 1-p removed, the largest is always kept
//   p = softmax(x); idx = argmax_v(p_v / q_v)       == torch.multinomial(p, 1) with q ~ Exp(1)
#include "sampler.h"

#include "common.cuh"
#include "host.h"

namespace vb {

constexpr int ST = 256;

__device__ __forceinline__ uint32_t f2key(float f) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < ST / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < ST / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}

// In-place ascending bitonic sort of (key, payload) pairs in shared memory; n is a power of two.
__device__ void bitonic_sort(float* key, int* val, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += ST) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const float a = key[i], b = key[ixj];
          const int va = val[i], vb2 = val[ixj];
          // total order: by key, ties by original index (stable, like torch.sort)
          const bool gt = (a > b) || (a == b && va > vb2);
          if (gt == up) { key[i] = b; key[ixj] = a; val[i] = vb2; val[ixj] = va; }
        }
      }
      __syncthreads();
    }
  }
}

// Exact k-th largest of xs[0..V) by a 4-pass radix select on order-preserving keys (whole CTA). hist is
// [ST/32][256]: one histogram per warp, so a shared-memory atomic only ever collides with lanes of its own warp (the
// top byte of a logit takes a handful of values: a single histogram serialised nearly all 4096 updates of the first
// pass), and the bin that holds the k-th largest is found by warp 0 with a suffix scan over the summed bins (eight
// bins per lane + one shuffle scan) instead of a 255-step loop on one thread. Ends with a barrier.
__device__ float block_kth_largest(const float* xs, int V, int k, int* hist, uint32_t* sel_prefix, int* sel_k) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { *sel_prefix = 0; *sel_k = k; }
  int* my = hist + warp * 256;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
#pragma unroll
    for (int i = 0; i < 8; ++i) my[lane + 32 * i] = 0;
    __syncthreads();
    const uint32_t prefix = *sel_prefix;
    const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int v = tid; v < V; v += ST) {
      const uint32_t key = f2key(xs[v]);
      if ((key & mask) == prefix) atomicAdd(&my[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (warp == 0) {
      // lane L owns bins [8L, 8L+8): c[i] = count of bin 8L+i over all warps; suffix sums from the top bin down
      int c[8], tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int a = 0;
#pragma unroll
        for (int w = 0; w < ST / 32; ++w) a += hist[w * 256 + 8 * lane + i];
        c[i] = a;
        tot += a;
      }
      int above = tot;  // inclusive suffix over lanes: counts of all bins >= 8L
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t2 = __shfl_down_sync(0xffffffffu, above, o);
        if (lane + o < 32) above += t2;
      }
      above -= tot;  // bins strictly above this lane's eight
      const int need = *sel_k;
      // the wanted bin is the highest b with (count of bins > b) < need <= (count of bins >= b); exactly one lane has it
      int acc = above, found = -1, need_out = 0;
#pragma unroll
      for (int i = 7; i >= 0; --i) {
        if (found < 0 && acc < need && need <= acc + c[i]) { found = 8 * lane + i; need_out = need - acc; }
        acc += c[i];
      }
      // fewer than `need` matching keys in total cannot happen (need <= remaining count by construction), but bin 0 is the
      // fallback of the serial version: keep it
      const unsigned has = __ballot_sync(0xffffffffu, found >= 0);
      if (has == 0) { if (lane == 0) { *sel_prefix = prefix; } }
      else if (found >= 0) { *sel_k = need_out; *sel_prefix = prefix | ((uint32_t)found << shift); }
    }
    __syncthreads();
  }
  return key2f(*sel_prefix);
}

template <bool SMOOTH>  // SMOOTH: also emit the more_smooth soft embedding (separate instantiation: 32 extra registers)
__global__ void __launch_bounds__(ST)
sample_kernel(const float* __restrict__ logits, int B, int l, int V, int use_cfg, float one_plus_t, float t,
              const float* __restrict__ q, int top_k, float top_p, float p_lim, long long* __restrict__ idx_out,
              float* __restrict__ mixed_out, int Vpow2, const float* __restrict__ q_gumbel, float tau, float logit_mul,
              const float* __restrict__ codebook, int Cvae, float* __restrict__ h_out) {
  extern __shared__ float sm[];
  float* xs = sm;                                   // [V]
  float* sk = xs + V;                               // [Vpow2] sort keys (top-p only)
  int* sv = reinterpret_cast<int*>(sk + Vpow2);     // [Vpow2] sort payload (top-p only)
  __shared__ float red[ST / 32];
  __shared__ int hist[(ST / 32) * 256];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_k;
  __shared__ float bval[ST / 32];
  __shared__ int bidx[ST / 32];
  __shared__ int topp_n;
  const int r = blockIdx.x;  // row = b*l + pos
  const int tid = threadIdx.x;
  const float* lc = logits + (size_t)r * V;
  const float* lu = logits + ((size_t)B * l + r) * V;
  for (int v = tid; v < V; v += ST) {
    float x = lc[v];
    if (use_cfg) x = __fsub_rn(__fmul_rn(one_plus_t, x), __fmul_rn(t, lu[v]));
    xs[v] = x;
    if (mixed_out) mixed_out[(size_t)r * V + v] = x;
  }
  __syncthreads();

  if (top_k > 0 && top_k < V) {
    const float thr = block_kth_largest(xs, V, top_k, hist, &sel_prefix, &sel_k);
    for (int v = tid; v < V; v += ST)
      if (xs[v] < thr) xs[v] = -INFINITY;
    __syncthreads();
  }

  if (top_p > 0.f) {
    // helpers.py:11-15: ascending stable sort, softmax over the sorted row, inclusive cumsum (double accumulator as
    // ATen's CPU cumsum), entries with cumsum <= 1-p removed except the last (largest).
    // Entries that top-k already removed are -inf: they sort to the front, add exactly 0.0 to every sum and are removed
    // again - so only the SURVIVORS are compacted, padded to a power of two and sorted (900 of 4096 with the README's
    // top_k: a 1024-element bitonic network, 55 instead of 78 steps on a quarter of the data). Every value the decision
    // depends on is computed exactly as on the full sorted row: the softmax denominator adds the survivors in the
    // per-thread groups their full-row positions would fall into, the probabilities are evaluated in parallel with the
    // same float operations, and thread 0 only runs the order-preserving double accumulation over them.
    const int lane = tid & 31;
    if (tid == 0) topp_n = 0;
    __syncthreads();
    for (int v0 = 0; v0 < V; v0 += ST) {
      const int v = v0 + tid;
      const float x = v < V ? xs[v] : -INFINITY;
      const bool keep = x != -INFINITY;
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      int base = 0;
      if (lane == 0 && mask) base = atomicAdd(&topp_n, __popc(mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) {
        const int pos = base + __popc(mask & ((1u << lane) - 1u));
        sk[pos] = x;
        sv[pos] = v;
      }
    }
    __syncthreads();
    const int ns = topp_n;                    // >= 1: the row maximum always survives top-k
    int n2 = 1;
    while (n2 < ns) n2 <<= 1;
    for (int i2 = ns + tid; i2 < n2; i2 += ST) { sk[i2] = INFINITY; sv[i2] = 0x7fffffff; }
    __syncthreads();
    bitonic_sort(sk, sv, n2);
    float m = -INFINITY;
    for (int i2 = tid; i2 < ns; i2 += ST) m = fmaxf(m, sk[i2]);
    m = block_max(m, red);
    // survivor j sits at position (V - ns) + j of the full sorted row, which thread ((V - ns) + j) % ST summed
    float ps = 0.f;
    for (int j2 = (((tid - (V - ns)) % ST) + ST) % ST; j2 < ns; j2 += ST) ps += expf(sk[j2] - m);
    const float sum = block_sum(ps, red);
    for (int i2 = tid; i2 < ns; i2 += ST) sk[i2] = expf(sk[i2] - m) / sum;  // sorted probabilities, in place
    __syncthreads();
    if (tid == 0) {
      double c = 0.0;
      const float lim = p_lim;
      int stop = 0;
      // The running sum is monotone (probabilities >= 0, rounding is monotone): if it is still <= lim after eight more
      // terms it was so after each of them. Eight independent loads / conversions per test keep the dependent chain to
      // the eight double additions; the block in which the limit is crossed is walked term by term.
      while (stop + 8 <= ns - 1) {
        double c8 = c;
#pragma unroll
        for (int u = 0; u < 8; ++u) c8 += (double)sk[stop + u];
        if (!((float)c8 <= lim)) break;
        c = c8;
        stop += 8;
      }
      for (; stop < ns - 1; ++stop) {  // nothing behind the first kept entry is removed
        c += (double)sk[stop];
        if (!((float)c <= lim)) break;
      }
      topp_n = stop;  // entries [0, stop) of the sorted survivors are removed
    }
    __syncthreads();
    const int stop = topp_n;
    for (int i2 = tid; i2 < stop; i2 += ST) xs[sv[i2]] = -INFINITY;
    __syncthreads();
  }

  float m = -INFINITY;
  for (int v = tid; v < V; v += ST) m = fmaxf(m, xs[v]);
  m = block_max(m, red);
  float ps = 0.f;
  for (int v = tid; v < V; v += ST) ps += expf(xs[v] - m);
  const float sum = block_sum(ps, red);
  const float* qr = q + (size_t)r * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = tid; v < V; v += ST) {
    const float ratio = (expf(xs[v] - m) / sum) / qr[v];
    if (ratio > best) { best = ratio; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((tid & 31) == 0) { bval[tid >> 5] = best; bidx[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < ST / 32; ++w)
      if (bval[w] > best || (bval[w] == best && bidx[w] < bi)) { best = bval[w]; bi = bidx[w]; }
    idx_out[r] = bi == 0x7fffffff ? 0 : bi;
  }
  if constexpr (!SMOOTH) return;
  // ---- more_smooth (models/var.py:178-180, helpers.py:22-36): soft embedding of the filtered row,
  //      h = softmax((x * (1 + ratio) - log(q_gumbel)) / tau) @ codebook, q_gumbel ~ Exp(1) drawn after the sampler noise
  __syncthreads();
  const float* qg = q_gumbel + (size_t)r * V;
  float zm = -INFINITY;
  for (int v = tid; v < V; v += ST) zm = fmaxf(zm, (xs[v] * logit_mul - logf(qg[v])) / tau);
  zm = block_max(zm, red);
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  float es = 0.f;
  for (int v = tid; v < V; v += ST) {
    const float e = expf((xs[v] * logit_mul - logf(qg[v])) / tau - zm);
    es += e;
    if (e != 0.f) {
      const float* er = codebook + (size_t)v * Cvae;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < Cvae) acc[c] = fmaf(e, __ldg(er + c), acc[c]);
    }
  }
  es = block_sum(es, red);
  float* hacc = xs;  // the row is no longer needed: reuse its first (ST/32)*32 floats for the cross-warp reduction
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float w = warp_sum(acc[c]);
    if ((tid & 31) == 0) hacc[(tid >> 5) * 32 + c] = w;
  }
  __syncthreads();
  if (tid < Cvae) {
    float h = 0.f;
    for (int w = 0; w < ST / 32; ++w) h += hacc[w * 32 + tid];
    h_out[(size_t)r * Cvae + tid] = h / es;
  }
}

int sample_launch(const SampleArgs& a, cudaStream_t st) {
  VB_REQUIRE(a.logits && a.q && a.idx_out, "sample: null pointer");
  VB_REQUIRE(a.B > 0 && a.l > 0 && a.V > 0, "sample: bad shape B=%d l=%d V=%d", a.B, a.l, a.V);
  VB_REQUIRE(a.top_k >= 0 && a.top_p >= 0.f && a.top_p <= 1.f, "sample: bad top_k=%d top_p=%f", a.top_k, a.top_p);
  int vp2 = 1;
  while (vp2 < a.V) vp2 <<= 1;
  const bool use_p = a.top_p > 0.f;
  const size_t smem = (size_t)a.V * 4 + (use_p ? (size_t)vp2 * 8 : 0);
  VB_REQUIRE(smem <= 200 * 1024, "sample: V=%d too large for the shared-memory sampler", a.V);
  if (a.q_gumbel != nullptr) {
    VB_REQUIRE(a.codebook && a.h_out && a.Cvae > 0 && a.Cvae <= 32 && a.tau > 0.f && a.V >= 256,
               "sample: more_smooth needs codebook, h_out, 0 < Cvae <= 32, tau > 0 (Cvae=%d tau=%f)", a.Cvae, a.tau);
  }
  static vb::SmemAttrCache attr_cache;  // static smem (~1.2 KB) counts against the 48 KB default limit
  if (smem > 40 * 1024 && vb::ensure_dyn_smem(attr_cache, smem, sample_kernel<false>, sample_kernel<true>)) return vb::VB_ERR_CUDA;
  const float opt = (float)(1.0 + a.t), tf = (float)a.t;
  vb::ProfScope prof_scope(vb::PK_SAMPLE, st);
  auto kern = a.q_gumbel != nullptr ? sample_kernel<true> : sample_kernel<false>;
  kern<<<a.B * a.l, ST, smem, st>>>(a.logits, a.B, a.l, a.V, a.use_cfg, opt, tf, a.q, a.top_k, a.top_p, (float)(1.0 - (double)a.top_p),
                                    reinterpret_cast<long long*>(a.idx_out), a.mixed_out, use_p ? vp2 : 0, a.q_gumbel, a.tau, a.logit_mul,
                                    a.codebook, a.Cvae, a.h_out);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// Expected codebook distance of the (optionally CFG-mixed, optionally top-k-renormalised) next-token distribution to
// the ground-truth token: out[s,t] = sum_v p_v * dists[gt_t, v]   (var_analysis.py:468-490, mode l2_dist; dists =
// torch.cdist(E, E)). One CTA per (class sequence, position) row. Ties with the k-th probability are all kept.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST)
expected_dist_kernel(const float* __restrict__ lc, const float* __restrict__ lu, const int* __restrict__ gt,
                     const float* __restrict__ t_row, const float* __restrict__ dists, int L, int V, int top_k,
                     float* __restrict__ out) {
  extern __shared__ float sm[];
  float* xs = sm;  // [V]
  __shared__ float red[ST / 32];
  __shared__ int hist[(ST / 32) * 256];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_k;
  const int t = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const float* c = lc + ((size_t)s * L + t) * V;
  const float* u = lu ? lu + (size_t)t * V : nullptr;
  const float tr = u ? __ldg(t_row + t) : 0.f, opt = 1.f + tr;
  float m = -INFINITY;
  for (int v = tid; v < V; v += ST) {
    float x = c[v];
    if (u) x = __fsub_rn(__fmul_rn(opt, x), __fmul_rn(tr, u[v]));
    xs[v] = x;
    m = fmaxf(m, x);
  }
  m = block_max(m, red);  // includes the barrier that publishes xs
  float thr = -INFINITY;
  if (top_k > 0 && top_k < V) thr = block_kth_largest(xs, V, top_k, hist, &sel_prefix, &sel_k);
  const float* d = dists + (size_t)__ldg(gt + t) * V;
  float se = 0.f, sd = 0.f;
  for (int v = tid; v < V; v += ST) {
    const float x = xs[v];
    if (x >= thr) {
      const float e = expf(x - m);
      se += e;
      sd = fmaf(e, __ldg(d + v), sd);
    }
  }
  se = block_sum(se, red);
  sd = block_sum(sd, red);
  if (tid == 0) out[(size_t)s * L + t] = sd / se;
}

int cfg_token_expected_dist(const float* lc, const float* lu, const int* gt, const float* t_row, const float* dists,
                            int n_seq, int L, int V, int top_k, float* out, cudaStream_t st) {
  VB_REQUIRE(lc && gt && dists && out && n_seq > 0 && L > 0 && V > 0, "cfg_token_expected_dist: bad arguments");
  VB_REQUIRE(lu == nullptr || t_row != nullptr, "cfg_token_expected_dist: t_row is required with uncond logits");
  VB_REQUIRE(n_seq <= 65535 && top_k >= 0 && (size_t)V * 4 <= 200 * 1024, "cfg_token_expected_dist: n_seq/top_k/V out of range");
  const size_t smem = (size_t)V * sizeof(float);
  static vb::SmemAttrCache attr_cache;
  if (smem > 40 * 1024 && vb::ensure_dyn_smem(attr_cache, smem, expected_dist_kernel)) return vb::VB_ERR_CUDA;
  vb::ProfScope prof_scope(vb::PK_SCORE_FIN, st);
  expected_dist_kernel<<<dim3(L, n_seq), ST, smem, st>>>(lc, lu, gt, t_row, dists, L, V, top_k, out);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

// ------------------------------------------------------------------------------------------------
// Neighbour-restricted arg-max token selection of VAR.smooth_sampling (models/var.py:483-536): one CTA per
// (image, position) row. x = CFG mix of the row's logits, lp = log_softmax(x); candidates are the first n_nb codebook
// neighbours of the ground-truth token (ascending L2 distance, table `neighbors` [V, n_nb]); a candidate j takes
// part if j < cand_count (count mode) or dists[gt, cand_j] <= d_0 + (thr - d_0) * ratio (threshold mode, thr_mode=1).
// Outputs the selected token, its log-probability and log_softmax(-d)[j*] over all n_nb candidates.
// First maximal candidate wins (torch.max on CPU); no candidate left -> candidate 0 (:521-527).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST)
neighbor_select_kernel(const float* __restrict__ logits, int B, int l, int V, float one_plus_t, float t,
                       const int* __restrict__ gt, const int* __restrict__ neighbors, const float* __restrict__ dists,
                       int n_nb, int cand_count, int thr_mode, float thr, float ratio, long long* __restrict__ tok_out,
                       float* __restrict__ lp_out, float* __restrict__ dlp_out) {
  extern __shared__ float sm[];
  float* xs = sm;  // [V]
  __shared__ float red[ST / 32];
  __shared__ float bval[ST / 32];
  __shared__ int bidx[ST / 32];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* lc = logits + (size_t)r * V;
  const float* lu = logits + ((size_t)B * l + r) * V;
  float m = -INFINITY;
  for (int v = tid; v < V; v += ST) {
    const float x = __fsub_rn(__fmul_rn(one_plus_t, lc[v]), __fmul_rn(t, lu[v]));
    xs[v] = x;
    m = fmaxf(m, x);
  }
  m = block_max(m, red);
  float se = 0.f;
  for (int v = tid; v < V; v += ST) se += expf(xs[v] - m);
  const float lse = m + logf(block_sum(se, red));
  const int g = __ldg(gt + r);
  const int* nb = neighbors + (size_t)g * n_nb;
  const float* dg = dists + (size_t)g * V;
  const float d0 = __ldg(dg + __ldg(nb));
  const float eff = d0 + (thr - d0) * ratio;
  // distance log-softmax over all n_nb candidates + masked arg-max of the token log-probabilities
  float dm = -INFINITY;
  for (int j = tid; j < n_nb; j += ST) dm = fmaxf(dm, -__ldg(dg + __ldg(nb + j)));
  dm = block_max(dm, red);
  float ds = 0.f, best = -INFINITY;
  int best_j = 0x7fffffff;
  for (int j = tid; j < n_nb; j += ST) {
    const int c = __ldg(nb + j);
    const float d = __ldg(dg + c);
    ds += expf(-d - dm);
    const bool in = thr_mode ? (d <= eff) : (j < cand_count);
    const float lp = in ? xs[c] - lse : -INFINITY;
    if (in && (lp > best || best_j == 0x7fffffff)) { best = lp; best_j = j; }  // j ascending per thread: first max kept
  }
  const float dlse = dm + logf(block_sum(ds, red));
  // block arg-max, ties -> smallest j
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (oj != 0x7fffffff && (best_j == 0x7fffffff || ov > best || (ov == best && oj < best_j))) { best = ov; best_j = oj; }
  }
  if ((tid & 31) == 0) { bval[tid >> 5] = best; bidx[tid >> 5] = best_j; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < ST / 32; ++w) {
      const float ov = bval[w];
      const int oj = bidx[w];
      if (oj != 0x7fffffff && (best_j == 0x7fffffff || ov > best || (ov == best && oj < best_j))) { best = ov; best_j = oj; }
    }
    if (best_j == 0x7fffffff) { best_j = 0; best = -INFINITY; }
    const int c = __ldg(nb + best_j);
    tok_out[r] = c;
    lp_out[r] = best;
    dlp_out[r] = -__ldg(dg + c) - dlse;
  }
}

int neighbor_select(const float* logits, int B, int l, int V, double t, const int* gt, const int* neighbors,
                    const float* dists, int n_nb, int cand_count, int thr_mode, float thr, float ratio, void* tok_out,
                    float* lp_out, float* dlp_out, cudaStream_t st) {
  VB_REQUIRE(logits && gt && neighbors && dists && tok_out && lp_out && dlp_out, "neighbor_select: null pointer");
  VB_REQUIRE(B > 0 && l > 0 && V > 0 && n_nb > 0 && n_nb <= V && (size_t)V * 4 <= 200 * 1024, "neighbor_select: bad shape");
  VB_REQUIRE(thr_mode || (cand_count >= 1 && cand_count <= n_nb), "neighbor_select: cand_count=%d out of [1,%d]", cand_count, n_nb);
  const size_t smem = (size_t)V * sizeof(float);
  static vb::SmemAttrCache attr_cache;
  if (smem > 40 * 1024 && vb::ensure_dyn_smem(attr_cache, smem, neighbor_select_kernel)) return vb::VB_ERR_CUDA;
  vb::ProfScope prof_scope(vb::PK_SAMPLE, st);
  neighbor_select_kernel<<<B * l, ST, smem, st>>>(logits, B, l, V, (float)(1.0 + t), (float)t, gt, neighbors, dists, n_nb,
                                                 cand_count, thr_mode, thr, ratio, reinterpret_cast<long long*>(tok_out),
                                                 lp_out, dlp_out);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
