#include <stdlib.h>

#include <algorithm>
// Host launcher for the tcgen05 GEMM family + the raw C-ABI entry used by tests and by the block driver.
#include "gemm.h"

#include "gemm_sm100.cuh"
#include "host.h"

namespace vb {

template <int BN, int EPI, int CTAS, bool LNF = false>
static int launch_one(const void* A, const void* W, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTAS>;
  CUtensorMap tmA, tmB;
  if (p.conv_kpt > 0) {
    // activations [B, H, W, Cin] as a 4-D map (C, W, H, B); a box = bw x bh pixels x 64 channels = one 128-row A tile
    const int Cin = p.conv_cin, nB = p.M / (p.conv_H * p.conv_W);
    const uint32_t bw = p.conv_W < GEMM_BM ? p.conv_W : GEMM_BM, bh = GEMM_BM / bw;
    const uint32_t sd = p.conv_stride > 1 ? (uint32_t)p.conv_stride : 1u;
    const uint64_t Wi = (uint64_t)p.conv_W * sd, Hi = (uint64_t)p.conv_H * sd;  // input extent
    uint64_t dims[4] = {(uint64_t)Cin, Wi, Hi, (uint64_t)nB};
    uint64_t str[3] = {(uint64_t)Cin * 2, Wi * Cin * 2, Hi * Wi * Cin * 2};
    uint32_t box[4] = {GEMM_BK, bw * sd, bh * sd, 1};  // traversed extent; delivers bw x bh pixels
    uint32_t es[4] = {1, sd, sd, 1};
    int r = make_tmap_bf16_sw128(&tmA, A, 4, dims, str, box, sd > 1 ? es : nullptr);
    if (r) return r;
  } else {
    const uint64_t acols = p.a_cols ? p.a_cols : p.K;
    uint64_t dims[2] = {acols, (uint64_t)p.M};
    uint64_t str[1] = {(uint64_t)(p.lda ? p.lda : acols) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int r = make_tmap_bf16_sw128(&tmA, A, 2, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)(p.w_cols ? p.w_cols : p.K), (uint64_t)(p.w_rows ? p.w_rows : p.N)};
    uint64_t str[1] = {(uint64_t)(p.ldw ? p.ldw : (p.w_cols ? p.w_cols : p.K)) * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(BN / CTAS)};
    int r = make_tmap_bf16_sw128(&tmB, W, 2, dims, str, box);
    if (r) return r;
  }
  auto kern = gemm_bf16_kernel<BN, EPI, CTAS, LNF>;
  static SmemAttrCache attr_cache;  // per instantiation, per device
  if (ensure_dyn_smem(attr_cache, Cfg::SMEM_BYTES, kern)) return VB_ERR_CUDA;
  const int m_tiles = (p.M + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int max_cl = sm_count() / CTAS;
  const int ncl = total < max_cl ? total : max_cl;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl * CTAS);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  vb::ProfScope prof_scope(EPI, st);
  VB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

template <int BN, int CTAS>
static int launch_bn(const void* A, const void* W, const GemmParams& p, int epi, cudaStream_t st) {
  switch (epi) {
    case EPI_BIAS_F32: return launch_one<BN, EPI_BIAS_F32, CTAS>(A, W, p, st);
    case EPI_BIAS_BF16: return launch_one<BN, EPI_BIAS_BF16, CTAS>(A, W, p, st);
    case EPI_GELU_BF16:
      return p.ln_part_in ? launch_one<BN, EPI_GELU_BF16, CTAS, true>(A, W, p, st) : launch_one<BN, EPI_GELU_BF16, CTAS>(A, W, p, st);
    case EPI_GATE_RESID:
      return p.ln_a_out ? launch_one<BN, EPI_GATE_RESID, CTAS, true>(A, W, p, st) : launch_one<BN, EPI_GATE_RESID, CTAS>(A, W, p, st);
    case EPI_QKV:
      return p.ln_part_in ? launch_one<BN, EPI_QKV, CTAS, true>(A, W, p, st) : launch_one<BN, EPI_QKV, CTAS>(A, W, p, st);
    case EPI_SCORE: return launch_one<BN, EPI_SCORE, CTAS>(A, W, p, st);
  }
  set_error("gemm: unknown epilogue %d", epi);
  return VB_ERR_ARG;
}

int gemm_pick_bn(int N) {
  // 256-wide tiles need the least operand traffic per FLOP (61 vs 71 B/clk/SM of L2 ingress for 192): they win as long
  // as their padding wastes under 3 % (d30 QKV, N = 5760: +7 % measured); otherwise the smallest padded N wins.
  int best = 256, best_pad = ((N + 255) / 256) * 256;
  if ((long long)best_pad * 100 <= (long long)N * 103) return 256;
  const int cands[2] = {192, 128};
  for (int bn : cands) {
    int pad = ((N + bn - 1) / bn) * bn;
    if (pad < best_pad) { best = bn; best_pad = pad; }
  }
  return best;
}

// Tile width by a small cost model (VAR_B200_GEMM_WIDE=0 restores gemm_pick_bn's N-only rule for A/B runs).
// A launch takes  waves x k-blocks x (cycles one tile spends per 64-deep k-block)  with
//   waves  = ceil(tiles / CTA pairs (or CTAs for the 1-CTA kernel)),
//   cycles = max(tensor time 2*BN, operand ingress (128 + BN/CTAS) rows x 128 B at ~64 B/clk/SM)
//          = 512 / 448 / 384 for BN = 256 / 192 / 128 on a CTA pair (profiles/r01_ncu_prof_gemm*.txt: the ingress cap).
// With many waves this prefers the widest tile whose padding it can afford (d30 QKV / proj / fc2 at B=256: 256, as
// measured); with few waves it removes the wave quantisation that the N-only rule left behind at 32 images per GPU
// (BASELINE configs[4] on 8 GPUs: M = 64 * l): e.g. M = 2304, N = 1920 is 90 tiles of 256x192 = two waves on 74 pairs,
// but 72 tiles of 256x256 = one. The SCORE epilogue sizes its partials with gemm_pick_bn(N) and keeps that choice.
static int gemm_pick_bn_mn(int M, int N, int epi) {
  const int bn0 = gemm_pick_bn(N);
  if (epi == EPI_SCORE) return bn0;
  static const int wide = [] { const char* e = getenv("VAR_B200_GEMM_WIDE"); return e ? atoi(e) : 1; }();  // A/B switch
  if (!wide) return bn0;
  const bool pair = M > GEMM_BM;
  const long long units = pair ? vb::sm_count() / 2 : vb::sm_count();
  const long long m_tiles = pair ? (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM) : 1;
  int best = bn0;
  long long best_cost = -1;
  const int cands[3] = {256, 192, 128};
  for (int bn : cands) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + units - 1) / units;
    const long long ingress = 2LL * (GEMM_BM + (pair ? bn / 2 : bn));
    const long long cyc = std::max<long long>(2LL * bn, ingress);
    const long long cost = waves * cyc;
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }  // ties: the wider tile (listed first)
  }
  return best;
}

int gemm_ln_parts(int M, int N) {
  const int bn = gemm_pick_bn_mn(M, N, EPI_GATE_RESID);
  return (N + bn - 1) / bn * GEMM_EPI_SUB;
}

// VAR_B200_L2HINT="<a><w>" with n / f / l = evict normal / first / last for the activation and the weight stream.
// Default "nn": evict-last weights save 0.2-0.7 GB of DRAM reads per d30 fc1 / fc2 launch under ncu but nothing inside
// the steps (d30 933.1 / 933.7 vs 933.8 ms, d16 scoring 323.5 / 324.0 vs 322.8 / 322.5 ms), and dead weights of earlier
// launches would linger in L2 with that priority.
static void l2_hints(unsigned long long* a, unsigned long long* w) {
  static const char* e = getenv("VAR_B200_L2HINT");
  auto dec = [](char c, unsigned long long dflt) {
    return c == 'f' ? L2_EVICT_FIRST : c == 'l' ? L2_EVICT_LAST : c == 'n' ? L2_EVICT_NORMAL : dflt;
  };
  *a = dec(e && e[0] ? e[0] : 0, L2_EVICT_NORMAL);
  *w = dec(e && e[0] && e[1] ? e[1] : 0, L2_EVICT_NORMAL);
}

int gemm_launch(const void* A, const void* W, const GemmParams& p_in, int epi, cudaStream_t st, int force_bn) {
  GemmParams p = p_in;
  if (!p.l2_hint_a && !p.l2_hint_w) l2_hints(&p.l2_hint_a, &p.l2_hint_w);
  VB_REQUIRE(A && W, "gemm: null operand");
  VB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  VB_REQUIRE(p.K % GEMM_BK == 0, "gemm: K=%d must be a multiple of %d", p.K, GEMM_BK);
  VB_REQUIRE(p.N % (epi == EPI_QKV ? 64 : 32) == 0, "gemm: N=%d must be a multiple of %d", p.N, epi == EPI_QKV ? 64 : 32);
  VB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "gemm: operands must be 16-byte aligned");
  if (epi == EPI_QKV) {
    VB_REQUIRE(p.N == 3 * p.C && p.C % 64 == 0 && p.H * 64 == p.C, "gemm/qkv: N=%d C=%d H=%d inconsistent", p.N, p.C, p.H);
    VB_REQUIRE(p.q_out && p.k_cache && p.v_cache && p.q_scale && p.bias, "gemm/qkv: null pointer");
    VB_REQUIRE(p.rows_per_seq > 0 && p.M % p.rows_per_seq == 0, "gemm/qkv: M=%d not a multiple of rows_per_seq=%d", p.M,
               p.rows_per_seq);
    VB_REQUIRE(p.pos0 >= 0 && p.pos0 + p.rows_per_seq <= p.Lmax, "gemm/qkv: cache overflow pos0=%d l=%d Lmax=%d", p.pos0,
               p.rows_per_seq, p.Lmax);
  } else if (epi == EPI_SCORE) {
    VB_REQUIRE(p.gt && p.part && p.gt_logit && p.bias && p.gt_mod > 0, "gemm/score: null pointer or gt_mod=%d", p.gt_mod);
  } else {
    VB_REQUIRE(p.out, "gemm: null output");
    if (epi == EPI_GATE_RESID) {
      VB_REQUIRE(p.resid && p.gate && p.rows_per_seq > 0, "gemm/gate: null pointer or rows_per_seq=%d", p.rows_per_seq);
      VB_REQUIRE(!p.ln_a_out || (p.ln_scale && p.ln_part_out && !force_bn), "gemm/gate: deferred LayerNorm needs ln_scale, "
                 "ln_part_out and the automatic tile width");
    }
  }
  if (p.ln_part_in) {
    VB_REQUIRE(epi == EPI_QKV || epi == EPI_GELU_BF16, "gemm: deferred LayerNorm input is built for the QKV and GELU epilogues");
    VB_REQUIRE(p.ln_u && p.ln_v && p.ln_labels && p.ln_parts > 0 && p.ln_C > 0 && p.rows_per_seq > 0,
               "gemm: deferred LayerNorm needs ln_u, ln_v, ln_labels, ln_parts, ln_C, rows_per_seq");
  }
  if (p.bd_rows > 0) {
    VB_REQUIRE(p.conv_kpt == 0 && p.bd_rows % (2 * GEMM_BM) == 0 && p.M % p.bd_rows == 0,
               "gemm: block-diagonal batching needs bd_rows=%d to be a multiple of %d and to divide M=%d", p.bd_rows,
               2 * GEMM_BM, p.M);
    VB_REQUIRE(epi == EPI_BIAS_F32 || epi == EPI_BIAS_BF16, "gemm: block-diagonal batching is built for the plain epilogues");
  }
  VB_REQUIRE(p.lda % 8 == 0 && p.ldw % 8 == 0, "gemm: lda=%d / ldw=%d must be multiples of 8 elements", p.lda, p.ldw);
  int bn = force_bn ? (force_bn & 0xffff) : gemm_pick_bn_mn(p.M, p.N, epi);
  if (!force_bn && epi == EPI_BIAS_BF16 && p.N <= 32) bn = 32;
  // CTA-pair tiles (256 x BN) unless the problem is a single 128-row tile or the caller forces 1-CTA (bit 16)
  const bool pair = !(force_bn & 0x10000) && p.M > GEMM_BM;
  if (bn == 160) {  // 160-wide tiles exist for the convolution epilogue only (channel counts 160 / 320 / 640 of the VQVAE)
    VB_REQUIRE(epi == EPI_BIAS_BF16, "gemm: tile width 160 is only built for EPI_BIAS_BF16");
    return pair ? launch_one<160, EPI_BIAS_BF16, 2>(A, W, p, st) : launch_one<160, EPI_BIAS_BF16, 1>(A, W, p, st);
  }
  if (bn == 32) {  // 32-wide tiles: the VQVAE's 3- and 32-channel convolutions (conv_out, quant_conv, post_quant_conv)
    VB_REQUIRE(epi == EPI_BIAS_BF16, "gemm: tile width 32 is only built for EPI_BIAS_BF16");
    return pair ? launch_one<32, EPI_BIAS_BF16, 2>(A, W, p, st) : launch_one<32, EPI_BIAS_BF16, 1>(A, W, p, st);
  }
  if (pair) {
    switch (bn) {
      case 256: return launch_bn<256, 2>(A, W, p, epi, st);
      case 192: return launch_bn<192, 2>(A, W, p, epi, st);
      case 128: return launch_bn<128, 2>(A, W, p, epi, st);
    }
  } else {
    switch (bn) {
      case 256: return launch_bn<256, 1>(A, W, p, epi, st);
      case 192: return launch_bn<192, 1>(A, W, p, epi, st);
      case 128: return launch_bn<128, 1>(A, W, p, epi, st);
    }
  }
  set_error("gemm: unsupported tile width %d", bn);
  return VB_ERR_ARG;
}

int conv3x3_launch(const void* x, const void* w_packed, const float* bias, const void* resid, void* out, int B, int H,
                   int W, int Cin, int Cout, cudaStream_t st, int stride, float2* gn_part) {
  VB_REQUIRE(x && w_packed && out && B > 0 && H > 0 && W > 0, "conv3x3: bad arguments");
  VB_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d (1 or 2)", stride);
  if (stride == 2) {  // H, W: input extent; the GEMM rows are the (H/2) x (W/2) output pixels
    VB_REQUIRE(H % 2 == 0 && W % 2 == 0 && W / 2 <= 128, "conv3x3/s2: H=%d W=%d must be even, W <= 256", H, W);
    H /= 2;
    W /= 2;
  }
  VB_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cout % 32 == 0, "conv3x3: Cin=%d must be a multiple of 8, Cout=%d of 32", Cin, Cout);
  VB_REQUIRE((W <= GEMM_BM ? GEMM_BM % W == 0 : W % GEMM_BM == 0) && (H * W) % GEMM_BM == 0,
             "conv3x3: H=%d W=%d not tileable into 128-pixel row patches", H, W);
  VB_REQUIRE((long long)B * H * W < (1ll << 31), "conv3x3: too many pixels");
  GemmParams p{};
  p.M = B * H * W;
  p.N = Cout;
  p.conv_kpt = (Cin + GEMM_BK - 1) / GEMM_BK;
  p.K = 9 * p.conv_kpt * GEMM_BK;
  p.conv_H = H;
  p.conv_W = W;
  p.bias = bias;
  p.out = out;
  p.resid_bf16 = reinterpret_cast<const __nv_bfloat16*>(resid);
  p.conv_cin = Cin;
  p.conv_stride = stride;
  p.gn_part = gn_part;
  return gemm_launch(x, w_packed, p, EPI_BIAS_BF16, st, Cout % 160 == 0 ? 160 : (Cout == 32 ? 32 : 0));
}

}  // namespace vb
