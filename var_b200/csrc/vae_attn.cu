// Single-head spatial self-attention of the VQVAE CNN (reference: models/basic_vae.py:63-92 AttnBlock.forward) on NHWC
// bf16 activations, built from var_b200's own kernels only:
//   g      = GroupNorm(x)                                   gn_stats / gn_apply (csrc/groupnorm.cu)
//   [q|k]  = g Wqk^T + b_qk            [B*HW, 2C]           tcgen05 GEMM
//   v^T    = Wv g^T                    [C, B*HW]            tcgen05 GEMM (operands swapped: V arrives transposed, which is
//                                                           the K-major B operand the P V product needs)
//   S_b    = q_b k_b^T                 [B*HW, HW] fp32      tcgen05 GEMM, block-diagonal batching over the images
//   P      = softmax(S * C^-0.5)       bf16                 softmax_rows_kernel (one warp per query row)
//   O_b    = P_b v_b + b_v             [B*HW, C]            tcgen05 GEMM, block-diagonal along K (rows of P sum to one,
//                                                           so the value bias passes through the average unchanged)
//   out    = x + O Wp^T + b_p                               tcgen05 GEMM, shortcut fused in the epilogue
#include <algorithm>

#include "../../include/var_b200.h"
#include "common.cuh"
#include "gemm.h"
#include "host.h"

namespace vb {

// P[r, :] = softmax(scale * S[r, :]) for rows of `n` fp32 scores (n % 32 == 0, n <= 4096); one warp per row, the row
// stays in registers between the maximum, the exponentials and the normalisation (one read of S, one write of P).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, long long rows, int n, float scale_log2e) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* s = S + r * n;
  float mx = -INFINITY;
  for (int i = lane * 4; i < n; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float m2 = mx * scale_log2e;
  float sum = 0.f;
  for (int i = lane * 4; i < n; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    sum += (exp2f(fmaf(v.x, scale_log2e, -m2)) + exp2f(fmaf(v.y, scale_log2e, -m2))) +
           (exp2f(fmaf(v.z, scale_log2e, -m2)) + exp2f(fmaf(v.w, scale_log2e, -m2)));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __nv_bfloat16* p = P + r * n;
  for (int i = lane * 4; i < n; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    uint2 o;
    o.x = pack_bf16x2(exp2f(fmaf(v.x, scale_log2e, -m2)) * inv, exp2f(fmaf(v.y, scale_log2e, -m2)) * inv);
    o.y = pack_bf16x2(exp2f(fmaf(v.z, scale_log2e, -m2)) * inv, exp2f(fmaf(v.w, scale_log2e, -m2)) * inv);
    *reinterpret_cast<uint2*>(p + i) = o;
  }
}

struct VaeAttnWs {
  uint8_t *g, *qk, *vt, *s, *p, *o, *gn;
  size_t bytes;
};

static size_t au(size_t x) { return (x + 255) & ~(size_t)255; }

static VaeAttnWs carve_vae_attn(void* work, int B, int HW, int C, int groups) {
  VaeAttnWs w{};
  uint8_t* base = reinterpret_cast<uint8_t*>(work);
  size_t off = 0;
  const size_t M = (size_t)B * HW;
  auto take = [&](size_t n) { uint8_t* r = base ? base + off : nullptr; off += au(n); return r; };
  w.g = take(M * C * 2);
  w.qk = take(M * 2 * C * 2);
  w.vt = take(M * C * 2);
  w.s = take(M * HW * 4);
  w.p = take(M * HW * 2);
  w.o = take(M * C * 2);
  w.gn = take(var_b200_gn_workspace(B, HW, C, groups));
  w.bytes = off;
  return w;
}

}  // namespace vb

extern "C" size_t var_b200_vae_attn_workspace(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return 0;
  return vb::carve_vae_attn(nullptr, B, HW, C, groups).bytes;
}

extern "C" int var_b200_vae_attn_block(const void* x, const float* gn_gamma, const float* gn_beta, int groups, float eps,
                                       const void* w_qkv, const float* b_qkv, const void* w_proj, const float* b_proj,
                                       void* out, int B, int HW, int C, void* work, size_t work_bytes, void* stream) {
  using namespace vb;
  VB_REQUIRE(x && gn_gamma && gn_beta && w_qkv && b_qkv && w_proj && b_proj && out && work, "vae_attn: null pointer");
  VB_REQUIRE(B > 0 && HW > 0 && HW % 256 == 0 && HW <= 4096 && C % 64 == 0,
             "vae_attn: needs HW=%d a multiple of 256 (<= 4096) and C=%d a multiple of 64", HW, C);
  VB_REQUIRE((long long)B * HW < (1ll << 31) / 8, "vae_attn: too many pixels");
  const VaeAttnWs w = carve_vae_attn(work, B, HW, C, groups);
  VB_REQUIRE(work_bytes >= w.bytes, "vae_attn: workspace %zu < %zu", work_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * HW;
  int rc = var_b200_gn_silu_nhwc(x, nullptr, gn_gamma, gn_beta, w.g, B, HW, C, groups, eps, 0, w.gn,
                                 var_b200_gn_workspace(B, HW, C, groups), stream);
  if (rc) return rc;
  const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(w_qkv);
  GemmParams p{};
  // [q | k] = g [Wq; Wk]^T + [bq; bk]
  p.M = M; p.N = 2 * C; p.K = C; p.bias = b_qkv; p.out = w.qk;
  rc = gemm_launch(w.g, wq, p, EPI_BIAS_BF16, st);
  if (rc) return rc;
  // v^T = Wv g^T (no bias: added after the average)
  p = GemmParams{};
  p.M = C; p.N = M; p.K = C; p.out = w.vt;
  rc = gemm_launch(wq + (size_t)2 * C * C, w.g, p, EPI_BIAS_BF16, st);
  if (rc) return rc;
  // S_b = q_b k_b^T: q and k are column slices of the [M, 2C] buffer; image b multiplies rows [b*HW, (b+1)*HW) of k
  p = GemmParams{};
  p.M = M; p.N = HW; p.K = C; p.out = w.s;
  p.lda = 2 * C; p.ldw = 2 * C; p.bd_rows = HW; p.bd_w_row = HW; p.w_rows = M;
  rc = gemm_launch(w.qk, reinterpret_cast<const __nv_bfloat16*>(w.qk) + C, p, EPI_BIAS_F32, st);
  if (rc) return rc;
  {
    vb::ProfScope prof_scope(vb::PK_OTHER, st);
    const float sl2 = 1.4426950408889634f / sqrtf((float)C);  // w_ratio = C^-0.5 (basic_vae.py:70), in base-2 exponents
    softmax_rows_kernel<<<(M + 7) / 8, 256, 0, st>>>(reinterpret_cast<const float*>(w.s),
                                                     reinterpret_cast<__nv_bfloat16*>(w.p), M, HW, sl2);
    VB_CUDA_CHECK(cudaGetLastError());
    vb::count_launch();
  }
  // O_b = P_b v_b + b_v: W = v^T [C, B*HW], image b uses its columns [b*HW, (b+1)*HW)
  p = GemmParams{};
  p.M = M; p.N = C; p.K = HW; p.bias = b_qkv + 2 * C; p.out = w.o;
  p.ldw = M; p.bd_rows = HW; p.bd_w_k = HW; p.w_cols = M;
  rc = gemm_launch(w.p, w.vt, p, EPI_BIAS_BF16, st);
  if (rc) return rc;
  // out = x + proj_out(O)
  p = GemmParams{};
  p.M = M; p.N = C; p.K = C; p.bias = b_proj; p.out = out;
  p.resid_bf16 = reinterpret_cast<const __nv_bfloat16*>(x);
  return gemm_launch(w.o, w_proj, p, EPI_BIAS_BF16, st);
}
