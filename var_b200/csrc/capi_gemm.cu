// C-ABI wrapper over the GEMM family (declared in include/var_b200.h).
#include "../../include/var_b200.h"
#include "gemm.h"
#include "host.h"
#include "common.cuh"

extern "C" int var_b200_gemm_bf16(const var_b200_gemm_args_t* a, void* stream) {
  using namespace vb;
  VB_REQUIRE(a != nullptr, "gemm: null args");
  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias;
  p.out = a->out;
  p.resid = a->resid;
  p.gate = a->gate;
  p.rows_per_seq = a->rows_per_seq;
  p.gate_ld = a->gate_ld;
  p.q_out = reinterpret_cast<__nv_bfloat16*>(a->q_out);
  p.k_cache = reinterpret_cast<__nv_bfloat16*>(a->k_cache);
  p.v_cache = reinterpret_cast<__nv_bfloat16*>(a->v_cache);
  p.q_scale = a->q_scale;
  p.C = a->C; p.H = a->H; p.pos0 = a->pos0; p.Lmax = a->Lmax;
  p.no_l2norm = a->no_l2norm;
  p.gt = a->gt;
  p.gt_mod = a->gt_mod > 0 ? a->gt_mod : a->M;
  p.part = reinterpret_cast<float2*>(a->part);
  p.gt_logit = a->gt_logit;
  p.ln_a_out = reinterpret_cast<__nv_bfloat16*>(a->ln_a_out);
  p.ln_scale = a->ln_scale;
  p.ln_part_out = reinterpret_cast<float2*>(a->ln_part_out);
  p.ln_part_in = reinterpret_cast<const float2*>(a->ln_part_in);
  p.ln_parts = a->ln_parts; p.ln_C = a->ln_C; p.ln_eps = a->ln_eps;
  p.ln_u = a->ln_u; p.ln_v = a->ln_v; p.ln_labels = a->ln_labels;
  return gemm_launch(a->A, a->W, p, a->epilogue, (cudaStream_t)stream, a->force_bn);
}

extern "C" int var_b200_gemm_tile_n(int N) { return vb::gemm_pick_bn(N); }
extern "C" int var_b200_gemm_ln_parts(int M, int N) { return (M > 0 && N > 0) ? vb::gemm_ln_parts(M, N) : 0; }

extern "C" int var_b200_conv3x3_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid, void* out,
                                     int B, int H, int W, int Cin, int Cout, void* stream) {
  return vb::conv3x3_launch(x, w_packed, bias, resid, out, B, H, W, Cin, Cout, (cudaStream_t)stream);
}

extern "C" int var_b200_conv3x3_s2_nhwc(const void* x, const void* w_packed, const float* bias, void* out, int B, int H, int W,
                                        int Cin, int Cout, void* stream) {
  return vb::conv3x3_launch(x, w_packed, bias, nullptr, out, B, H, W, Cin, Cout, (cudaStream_t)stream, 2);
}

extern "C" int var_b200_conv1x1_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid, void* out,
                                     long long n_pixels, int Cin, int Cout, void* stream) {
  using namespace vb;
  VB_REQUIRE(x && w_packed && out && n_pixels > 0 && n_pixels < (1ll << 31), "conv1x1: bad arguments");
  VB_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cout % 32 == 0, "conv1x1: Cin=%d must be a multiple of 8, Cout=%d of 32", Cin, Cout);
  GemmParams p{};
  p.M = (int)n_pixels;
  p.N = Cout;
  p.K = (Cin + 63) / 64 * 64;  // the GEMM's K block; weights are packed [Cout, K] with zero columns beyond Cin
  p.a_cols = Cin;                                   // the activation tail is zero-filled by TMA
  p.bias = bias;
  p.out = out;
  p.resid_bf16 = reinterpret_cast<const __nv_bfloat16*>(resid);
  return gemm_launch(x, w_packed, p, EPI_BIAS_BF16, (cudaStream_t)stream);
}
