// GroupNorm (+ SiLU) on NHWC bf16 activations for the VQVAE CNN decoder/encoder glue
// (reference: models/basic_vae.py:18-19 Normalize = GroupNorm(32, C, eps=1e-6, affine) followed by F.silu, :57-58,:159,:225).
// PyTorch runs this as moments + several elementwise passes over fp32/NCHW tensors (plus layout conversions around the
// cuDNN NHWC convolutions); here it is two HBM-bound passes over the bf16 NHWC tensor: (1) per-CTA partial sums in a
// fixed order (deterministic, no atomics), (2) normalise * gamma + beta, SiLU, store.
#include <algorithm>

#include "common.cuh"
#include "gemm.h"
#include "host.h"
#include "../../include/var_b200.h"

namespace vb {

constexpr int GN_THREADS = 256;
constexpr int GN_PIX_PER_CTA = 512;  // pixels reduced by one CTA of the statistics pass

// x: [B, HW, C] bf16. part: [B, nchunk, G, 2] fp32 (sum, sum of squares) of the chunk's pixels.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pre_bias, float* __restrict__ part, int HW,
                int C, int G) {
  __shared__ float red[GN_THREADS][8][2];  // per-thread (sum, sum of squares) of its 8 channels
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int nv = C >> 3;                  // 16-byte vectors per pixel
  const int ppp = GN_THREADS / nv;        // pixels per pass
  const int vl = threadIdx.x % nv, pl = threadIdx.x / nv;
  const int cpg = C / G;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const int p0 = chunk * GN_PIX_PER_CTA, p1 = min(HW, p0 + GN_PIX_PER_CTA);
  if (pl < ppp) {
    float pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pb[i] = pre_bias ? __ldg(pre_bias + vl * 8 + i) : 0.f;
    constexpr int UN = 4;  // four loads in flight per thread
    for (int p = p0 + pl; p < p1; p += UN * ppp) {
      uint4 v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u)
        if (p + u * ppp < p1) v[u] = *reinterpret_cast<const uint4*>(x + ((size_t)b * HW + p + u * ppp) * C + vl * 8);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (p + u * ppp >= p1) break;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = __bfloat1622float2(h[i]);
          f.x += pb[2 * i]; f.y += pb[2 * i + 1];
          s[2 * i] += f.x; q[2 * i] += f.x * f.x;
          s[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x][i][0] = s[i]; red[threadIdx.x][i][1] = q[i]; }
  __syncthreads();
  if (threadIdx.x < G) {  // one thread per group, fixed summation order: deterministic
    const int g = threadIdx.x;
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < ppp; ++r)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const int t = r * nv + (c >> 3), i = c & 7;
        ss += red[t][i][0];
        qq += red[t][i][1];
      }
    float* o = part + (((size_t)b * nchunk + chunk) * G + g) * 2;
    o[0] = ss;
    o[1] = qq;
  }
}

__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pre_bias, const float* __restrict__ part,
                int nchunk, const float* __restrict__ gamma,
                const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int HW, int C, int G, float eps, int silu,
                int pix_per_cta) {
  __shared__ float mean_s[64], rstd_s[64];
  const int b = blockIdx.y;
  const int cpg = C / G;
  if (threadIdx.x < G) {
    float s = 0.f, q = 0.f;
    for (int c = 0; c < nchunk; ++c) {  // fixed order
      const float* pp = part + (((size_t)b * nchunk + c) * G + threadIdx.x) * 2;
      s += pp[0]; q += pp[1];
    }
    const float n = (float)HW * (float)cpg;
    const float m = s / n;
    const float var = fmaxf(q / n - m * m, 0.f);
    mean_s[threadIdx.x] = m;
    rstd_s[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int nv = C >> 3;
  const int ppp = GN_THREADS / nv;
  const int vl = threadIdx.x % nv, pl = threadIdx.x / nv;
  if (pl >= ppp) return;
  float sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = vl * 8 + i, g = c / cpg;
    const float a = rstd_s[g] * __ldg(gamma + c);
    sc[i] = a;
    sf[i] = __ldg(beta + c) + ((pre_bias ? __ldg(pre_bias + c) : 0.f) - mean_s[g]) * a;
  }
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  // four pixels per thread and iteration: all loads are issued before the first use (memory-level parallelism);
  // SiLU(a) = a * sigmoid(a) = h + h * tanh(h), h = a / 2: one MUFU per element instead of exp + reciprocal
  constexpr int UN = 4;
  for (int p = p0 + pl; p < p1; p += UN * ppp) {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (p + u * ppp < p1) v[u] = *reinterpret_cast<const uint4*>(x + ((size_t)b * HW + p + u * ppp) * C + vl * 8);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (p + u * ppp >= p1) break;
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        float a = fmaf(f.x, sc[2 * i], sf[2 * i]), c = fmaf(f.y, sc[2 * i + 1], sf[2 * i + 1]);
        if (silu) {
          const float ha = 0.5f * a, hc = 0.5f * c;
          a = fmaf(ha, tanh_approx(ha), ha);
          c = fmaf(hc, tanh_approx(hc), hc);
        }
        ow[i] = pack_bf16x2(a, c);
      }
      *reinterpret_cast<uint4*>(y + ((size_t)b * HW + p + u * ppp) * C + vl * 8) = o;
    }
  }
}

// Statistics that the producing convolution's epilogue left behind (gemm_sm100.cuh, gn_part): per 128-pixel block and
// channel (sum, sum of squares) -> per image and group, in a fixed order: thread (g, k) sums the blocks k, k+8, ... of
// its group's channels, then the eight partial sums of a group are added in order. One CTA per image.
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float2* __restrict__ part, int T, int C, int G, float* __restrict__ sums) {
  __shared__ float2 red[256];
  const int b = blockIdx.x;
  const int cpg = C / G;
  const int g = threadIdx.x >> 3, k = threadIdx.x & 7;
  float s = 0.f, q = 0.f;
  if (g < G) {
    for (int t = k; t < T; t += 8) {
      const float2* pp = part + ((size_t)b * T + t) * C + g * cpg;
      for (int c = 0; c < cpg; ++c) {
        const float2 v = __ldg(pp + c);
        s += v.x;
        q += v.y;
      }
    }
  }
  red[threadIdx.x] = make_float2(s, q);
  __syncthreads();
  if (g < G && k == 0) {
    float ss = 0.f, qq = 0.f;
    for (int i = 0; i < 8; ++i) { ss += red[threadIdx.x + i].x; qq += red[threadIdx.x + i].y; }
    sums[((size_t)b * G + g) * 2] = ss;
    sums[((size_t)b * G + g) * 2 + 1] = qq;
  }
}

// out = a (+ bias_a[c]) + b (+ bias_b[c]); any of b / bias_a / bias_b may be null. In place allowed (out == a).
__global__ void __launch_bounds__(256)
add_bias_nhwc_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ bias_a, const __nv_bfloat16* b,
                     const float* __restrict__ bias_b, __nv_bfloat16* out, size_t n_vec, int C) {
  const int nv = C >> 3;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 256) {
    const int c0 = (int)(i % nv) * 8;
    const uint4 va = reinterpret_cast<const uint4*>(a)[i];
    uint4 vb2 = make_uint4(0, 0, 0, 0);
    if (b) vb2 = reinterpret_cast<const uint4*>(b)[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&va);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&vb2);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(ha[k]);
      if (b) { const float2 g = __bfloat1622float2(hb[k]); f.x += g.x; f.y += g.y; }
      if (bias_a) { f.x += __ldg(bias_a + c0 + 2 * k); f.y += __ldg(bias_a + c0 + 2 * k + 1); }
      if (bias_b) { f.x += __ldg(bias_b + c0 + 2 * k); f.y += __ldg(bias_b + c0 + 2 * k + 1); }
      ow[k] = pack_bf16x2(f.x, f.y);
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// nearest-neighbour 2x up-sampling (models/basic_vae.py:22-28), optional per-channel bias added on the fly
__global__ void __launch_bounds__(256)
upsample2x_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                       int B, int H, int W, int C) {
  const int nv = C >> 3;
  const size_t n_in = (size_t)B * H * W * nv;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_in; i += (size_t)gridDim.x * 256) {
    const int v = (int)(i % nv);
    size_t p = i / nv;
    const int xw = (int)(p % W); p /= W;
    const int yh = (int)(p % H);
    const int b = (int)(p / H);
    uint4 val = reinterpret_cast<const uint4*>(x)[i];
    if (bias) {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(h[k]);
        f.x += __ldg(bias + v * 8 + 2 * k); f.y += __ldg(bias + v * 8 + 2 * k + 1);
        h[k] = __floats2bfloat162_rn(f.x, f.y);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(y);
    const size_t row = (size_t)2 * W * nv;
    const size_t base = (((size_t)b * 2 * H + 2 * yh) * 2 * W + 2 * xw) * nv + v;
    o[base] = val; o[base + nv] = val; o[base + row] = val; o[base + row + nv] = val;
  }
}

}  // namespace vb

extern "C" int var_b200_add_bias_nhwc(const void* a, const float* bias_a, const void* b, const float* bias_b, void* out,
                                      long long n_pixels, int C, void* stream) {
  using namespace vb;
  VB_REQUIRE(a && out && n_pixels > 0 && C > 0 && C % 8 == 0, "add_bias_nhwc: bad arguments");
  const size_t n_vec = (size_t)n_pixels * (C >> 3);
  const int grid = (int)std::min<size_t>((n_vec + 255) / 256, (size_t)sm_count() * 16);
  vb::ProfScope prof_scope(vb::PK_OTHER, (cudaStream_t)stream);
  add_bias_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(a), bias_a,
                                                                reinterpret_cast<const __nv_bfloat16*>(b), bias_b,
                                                                reinterpret_cast<__nv_bfloat16*>(out), n_vec, C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

extern "C" int var_b200_upsample2x_nhwc(const void* x, const float* bias, void* y, int B, int H, int W, int C, void* stream) {
  using namespace vb;
  VB_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample2x_nhwc: bad arguments");
  const size_t n = (size_t)B * H * W * (C >> 3);
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
  vb::ProfScope prof_scope(vb::PK_OTHER, (cudaStream_t)stream);
  upsample2x_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), bias,
                                                                  reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

extern "C" size_t var_b200_gn_workspace(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || groups <= 0) return 0;
  const int nchunk = (HW + vb::GN_PIX_PER_CTA - 1) / vb::GN_PIX_PER_CTA;
  return (size_t)B * nchunk * groups * 2 * sizeof(float);
}

extern "C" int var_b200_gn_silu_nhwc(const void* x, const float* pre_bias, const float* gamma, const float* beta, void* y, int B,
                                     int HW, int C, int groups, float eps, int apply_silu, void* work, size_t work_bytes,
                                     void* stream) {
  using namespace vb;
  VB_REQUIRE(x && gamma && beta && y && work, "gn: null pointer");
  VB_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 && C / 8 <= GN_THREADS,
             "gn: unsupported shape B=%d HW=%d C=%d groups=%d", B, HW, C, groups);
  VB_REQUIRE(B <= 65535, "gn: batch too large");
  VB_REQUIRE(work_bytes >= var_b200_gn_workspace(B, HW, C, groups), "gn: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunk = (HW + GN_PIX_PER_CTA - 1) / GN_PIX_PER_CTA;
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  gn_stats_kernel<<<dim3(nchunk, B), GN_THREADS, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), pre_bias,
                                                          reinterpret_cast<float*>(work), HW, C, groups);
  VB_CUDA_CHECK(cudaGetLastError());
  const int ppc = 256;
  gn_apply_kernel<<<dim3((HW + ppc - 1) / ppc, B), GN_THREADS, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), pre_bias, reinterpret_cast<const float*>(work), nchunk, gamma, beta,
      reinterpret_cast<__nv_bfloat16*>(y), HW, C, groups, eps, apply_silu, ppc);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch(2);
  return VB_OK;
}

extern "C" size_t var_b200_conv3x3_gn_workspace(int B, int H, int W, int Cout) {
  if (B <= 0 || H <= 0 || W <= 0 || Cout <= 0) return 0;
  return (size_t)B * ((size_t)H * W / 128) * Cout * sizeof(float2);
}

extern "C" int var_b200_conv3x3_gn_nhwc(const void* x, const void* w_packed, const float* bias, const void* resid, void* out,
                                        int B, int H, int W, int Cin, int Cout, int groups, float* gn_sums, void* work,
                                        size_t work_bytes, void* stream) {
  using namespace vb;
  VB_REQUIRE(gn_sums && work && groups > 0 && groups <= 32 && Cout % groups == 0, "conv3x3_gn: bad arguments (groups=%d)", groups);
  VB_REQUIRE(work_bytes >= var_b200_conv3x3_gn_workspace(B, H, W, Cout), "conv3x3_gn: workspace too small");
  VB_REQUIRE(((size_t)H * W) % 256 == 0, "conv3x3_gn: H*W=%d must be a multiple of 256 (a CTA pair's tile stays inside one image)", H * W);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = conv3x3_launch(x, w_packed, bias, resid, out, B, H, W, Cin, Cout, st, 1, reinterpret_cast<float2*>(work));
  if (rc) return rc;
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  gn_finalize_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const float2*>(work), H * W / 128, Cout, groups, gn_sums);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

extern "C" int var_b200_gn_apply_nhwc(const void* x, const float* gn_sums, const float* gamma, const float* beta, void* y, int B,
                                      int HW, int C, int groups, float eps, int apply_silu, void* stream) {
  using namespace vb;
  VB_REQUIRE(x && gn_sums && gamma && beta && y, "gn_apply: null pointer");
  VB_REQUIRE(B > 0 && B <= 65535 && HW > 0 && C > 0 && C % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 &&
                 C / 8 <= GN_THREADS, "gn_apply: unsupported shape B=%d HW=%d C=%d groups=%d", B, HW, C, groups);
  cudaStream_t st = (cudaStream_t)stream;
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  const int ppc = 256;
  gn_apply_kernel<<<dim3((HW + ppc - 1) / ppc, B), GN_THREADS, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), nullptr, gn_sums, 1, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), HW, C,
      groups, eps, apply_silu, ppc);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}
