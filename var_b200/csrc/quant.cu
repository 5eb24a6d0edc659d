// Multi-scale residual vector quantizer of the VQVAE on sm_100a: one persistent CTA per image walks the scale
// pyramid (area down-sample -> codebook nearest neighbour -> code lookup -> bicubic up-sample -> Phi 3x3 conv ->
// f_hat += h, f_rest -= h), keeping the up-sampled map, the Phi weights and the pooled tokens in shared memory.
//
// Reference semantics: models/quant.py:135-166 (f_to_idxBl_or_fhat), :169-184 (idxBl_to_var_input), :187-196
// (get_next_autoregressive_input), :107-121 (embed_to_fhat), :199-206 (Phi).
//
// Arithmetic contract: this file is compiled with -fmad=false and every expression below is written in the exact
// operation order of oracle/quant_oracle.c, so indices and f_hat are bit-identical to the CPU oracle. The search is
// plain fp32 on CUDA cores (tcgen05 has no fp32 MMA; the whole search is 0.18 GFLOP per image).
#include "quant.h"

#include "common.cuh"
#include "host.h"

namespace vb {

constexpr int QT = 512;       // threads per CTA
constexpr int CV = 32;        // Cvae (lanes == channels)
constexpr int Q_TT = 2;       // tokens per thread in the search
constexpr int Q_CHUNK = 256;  // tokens searched per pass
constexpr int Q_MAXHW = 32;
constexpr int Q_SCRATCH = CV * Q_CHUNK + Q_CHUNK + 2 * QT * Q_TT;
constexpr int Q_WPHI = CV * 9 * CV + CV;
constexpr int Q_UNION = Q_SCRATCH > Q_WPHI ? Q_SCRATCH : Q_WPHI;

struct QuantKParams {
  int B, H, W, V, S;
  int ph[VB_MAX_SCALES], pw[VB_MAX_SCALES], phi[VB_MAX_SCALES];
  float resi, one_minus_resi;
  const float* codebook;  // [V,CV]
  const float* phi_w;     // [n_phi,CV,CV,3,3]
  const float* phi_b;     // [n_phi,CV]
  int si_begin, si_end;
  const float* f;  // encode: input features, else null
  float* f_rest;   // encode workspace
  float* f_hat;    // running reconstruction (in/out)
  int zero_fhat;
  int idx_concat;         // 1: idx holds all scales concatenated ([B,l_s] blocks), 0: idx holds scale si_begin only
  long long* idx;         // encode: out, otherwise in
  float* fhat_list;       // optional [S,B,CV,H,W]
  float* next_tokens;     // optional token-major area(f_hat -> next scale): row stride next_stride tokens per image
  int next_stride;
  float* next_nchw;       // optional [B,CV,ph_next,pw_next] (single-scale step only)
  // split encode (search done by the tensor-core kernel between launches)
  int split;              // 0: fused, 1: init + pool scale si_begin only, 2: update scale si_begin from idx, pool si_begin+1
  float* z_out;           // pooled tokens [B*l, CV] fp32 (token-major)
  __nv_bfloat16* zb_out;  // same, bf16, K padded to 64 with zeros: [B*l, 64]
  float* zz_out;          // |z|^2 per token [B*l] (sequential fp32 sum, the oracle's order)
};

__device__ __forceinline__ float q_cubic1(float x) {
  const float A = -0.75f;
  return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
}
__device__ __forceinline__ float q_cubic2(float x) {
  const float A = -0.75f;
  return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// windows of adaptive average pooling: [floor(o*I/O), ceil((o+1)*I/O))
__device__ __forceinline__ float area_at(const float* plane, int H, int W, int oh, int ow, int oy, int ox) {
  const int y0 = (oy * H) / oh, y1 = ((oy + 1) * H + oh - 1) / oh;
  const int x0 = (ox * W) / ow, x1 = ((ox + 1) * W + ow - 1) / ow;
  float s = 0.f;
  for (int iy = y0; iy < y1; ++iy)
    for (int ix = x0; ix < x1; ++ix) s = s + plane[iy * W + ix];
  s = s / (float)(y1 - y0);
  s = s / (float)(x1 - x0);
  return s;
}

__global__ void __launch_bounds__(QT, 1) quant_kernel(const QuantKParams p) {
  extern __shared__ float sm[];
  const int H = p.H, W = p.W, HW = H * W;
  const int PW = W + 2;
  const int plane = ((H + 2) * PW) | 1;  // odd stride: conflict-free when lanes index channels
  float* hup = sm;                       // [CV][plane] zero-bordered up-sampled map
  float* un = hup + CV * plane;          // union: Phi weights (conv phase) / search scratch (search phase)
  float* wphi = un;                      // [ci][tap][co]
  float* bphi = wphi + CV * 9 * CV;      // [co]
  float* zbuf = un;                      // [CV][Q_CHUNK] pooled tokens (channel-major)
  float* zz = zbuf + CV * Q_CHUNK;       // [Q_CHUNK]
  float* red_d = zz + Q_CHUNK;           // [QT*Q_TT]
  int* red_i = reinterpret_cast<int*>(red_d + QT * Q_TT);
  float* ee = un + Q_UNION;              // [V] (encode only)
  int* idx_s = reinterpret_cast<int*>(ee + (p.f ? p.V : 0));  // [HW]
  float* wy = reinterpret_cast<float*>(idx_s + HW);            // [Q_MAXHW][4]
  float* wx = wy + Q_MAXHW * 4;
  int* iy0 = reinterpret_cast<int*>(wx + Q_MAXHW * 4);
  int* ix0 = iy0 + Q_MAXHW;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const size_t img = (size_t)CV * HW;
  float* f_hat = p.f_hat + (size_t)b * img;
  float* f_rest = p.f_rest ? p.f_rest + (size_t)b * img : nullptr;
  const bool encode = p.f != nullptr && p.split == 0;         // search inside this kernel (fp32 CUDA cores)
  const bool upd_rest = f_rest != nullptr && (encode || p.split == 2);

  for (int i = tid; i < CV * plane; i += QT) hup[i] = 0.f;
  if (p.zero_fhat)
    for (size_t i = tid; i < img; i += QT) f_hat[i] = 0.f;
  if (p.f != nullptr) {
    const float* f = p.f + (size_t)b * img;
    for (size_t i = tid; i < img; i += QT) f_rest[i] = f[i];
  }
  // pooled tokens of scale `ps` -> global (input of the tensor-core search kernel)
  auto pool_out = [&](int ps) {
    const int ph = p.ph[ps], pw = p.pw[ps], l = ph * pw;
    const bool last = (ps == p.S - 1);
    for (int t0 = 0; t0 < l; t0 += Q_CHUNK) {
      const int lc = min(Q_CHUNK, l - t0);
      __syncthreads();
      for (int i = tid; i < CV * lc; i += QT) {
        const int c = i / lc, t = i - c * lc, tok = t0 + t;
        const float* pl = f_rest + (size_t)c * HW;
        zbuf[c * Q_CHUNK + t] = last ? pl[tok] : area_at(pl, H, W, ph, pw, tok / pw, tok % pw);
      }
      __syncthreads();
      for (int t = tid; t < lc; t += QT) {
        float s = 0.f;
        for (int c = 0; c < CV; ++c) { const float z = zbuf[c * Q_CHUNK + t]; s = s + z * z; }
        p.zz_out[(size_t)b * l + t0 + t] = s;
      }
      for (int i = tid; i < lc * 64; i += QT) {
        const int t = i >> 6, c = i & 63;
        const size_t tok = (size_t)b * l + t0 + t;
        const float z = c < CV ? zbuf[c * Q_CHUNK + t] : 0.f;
        if (c < CV) p.z_out[tok * CV + c] = z;
        p.zb_out[tok * 64 + c] = __float2bfloat16_rn(z);
      }
    }
    __syncthreads();
  };
  if (p.split == 1) {
    __syncthreads();
    pool_out(p.si_begin);
    return;
  }
  if (encode) {
    for (int v = tid; v < p.V; v += QT) {
      const float* e = p.codebook + (size_t)v * CV;
      float s = 0.f;
      for (int c = 0; c < CV; ++c) s = s + e[c] * e[c];
      ee[v] = s;
    }
  }
  long long idx_off = 0;  // offset of scale si inside the concatenated index buffer
  int tok_off = 0;        // offset of the next scale inside next_tokens
  if (p.idx_concat)
    for (int si = 0; si < p.si_begin; ++si) idx_off += (long long)p.B * p.ph[si] * p.pw[si];
  __syncthreads();

  for (int si = p.si_begin; si < p.si_end; ++si) {
    const int ph = p.ph[si], pw = p.pw[si], l = ph * pw;
    const bool last = (si == p.S - 1);
    long long* idx_g = p.idx + idx_off + (long long)b * l;

    if (encode) {
      // ---------------- nearest codebook entry for every pooled token ----------------
      for (int t0 = 0; t0 < l; t0 += Q_CHUNK) {
        const int lc = min(Q_CHUNK, l - t0);
        for (int i = tid; i < CV * lc; i += QT) {
          const int c = i / lc, t = i - c * lc, tok = t0 + t;
          const float* pl = f_rest + (size_t)c * HW;
          zbuf[c * Q_CHUNK + t] = last ? pl[tok] : area_at(pl, H, W, ph, pw, tok / pw, tok % pw);
        }
        __syncthreads();
        for (int t = tid; t < lc; t += QT) {
          float s = 0.f;
          for (int c = 0; c < CV; ++c) { const float z = zbuf[c * Q_CHUNK + t]; s = s + z * z; }
          zz[t] = s;
        }
        __syncthreads();
        const int G = (lc + Q_TT - 1) / Q_TT;       // token groups
        int S = QT / G;                             // code slices
        if (S > p.V) S = p.V;
        const int per = (p.V + S - 1) / S;          // codes per slice
        const int slice = tid / G, grp = tid - slice * G;
        if (slice < S) {
          float z[Q_TT][CV];
          float zzr[Q_TT];
#pragma unroll
          for (int u = 0; u < Q_TT; ++u) {
            const int t = grp * Q_TT + u;
            const int tc = t < lc ? t : lc - 1;
            zzr[u] = zz[tc];
#pragma unroll
            for (int c = 0; c < CV; ++c) z[u][c] = zbuf[c * Q_CHUNK + tc];
          }
          float best[Q_TT];
          int bidx[Q_TT];
#pragma unroll
          for (int u = 0; u < Q_TT; ++u) { best[u] = INFINITY; bidx[u] = 0; }
          const int v_begin = slice * per, v_end = min(p.V, v_begin + per);
          for (int v = v_begin; v < v_end; ++v) {
            const float4* e4 = reinterpret_cast<const float4*>(p.codebook + (size_t)v * CV);
            float e[CV];
#pragma unroll
            for (int c4 = 0; c4 < CV / 4; ++c4) {
              const float4 t4 = __ldg(e4 + c4);
              e[4 * c4] = t4.x; e[4 * c4 + 1] = t4.y; e[4 * c4 + 2] = t4.z; e[4 * c4 + 3] = t4.w;
            }
            const float eev = ee[v];
#pragma unroll
            for (int u = 0; u < Q_TT; ++u) {
              float dot = 0.f;
#pragma unroll
              for (int c = 0; c < CV; ++c) dot = dot + z[u][c] * e[c];
              const float d = (zzr[u] + eev) - 2.f * dot;
              if (d < best[u]) { best[u] = d; bidx[u] = v; }
            }
          }
#pragma unroll
          for (int u = 0; u < Q_TT; ++u) {
            const int t = grp * Q_TT + u;
            if (t < lc) { red_d[slice * lc + t] = best[u]; red_i[slice * lc + t] = bidx[u]; }
          }
        }
        __syncthreads();
        for (int t = tid; t < lc; t += QT) {
          float bd = red_d[t];
          int bi = red_i[t];
          for (int s = 1; s < S; ++s) {  // slices are ordered by code index: strict < keeps the first minimum
            const float d = red_d[s * lc + t];
            if (d < bd) { bd = d; bi = red_i[s * lc + t]; }
          }
          idx_s[t0 + t] = bi;
          idx_g[t0 + t] = bi;
        }
        __syncthreads();
      }
    } else {
      for (int t = tid; t < l; t += QT) idx_s[t] = (int)idx_g[t];
      __syncthreads();
    }

    // ---------------- code lookup + bicubic up-sample into the padded map ----------------
    if (!last) {
      if (tid < H + W) {
        const bool isy = tid < H;
        const int o = isy ? tid : tid - H;
        const int in = isy ? ph : pw, out = isy ? H : W;
        const float scale = (float)in / (float)out;
        const float s = scale * ((float)o + 0.5f) - 0.5f;
        const float fl = floorf(s);
        const float t = s - fl;
        const float x2 = 1.f - t;
        float* wv = (isy ? wy : wx) + 4 * o;
        wv[0] = q_cubic2(t + 1.f);
        wv[1] = q_cubic1(t);
        wv[2] = q_cubic1(x2);
        wv[3] = q_cubic2(x2 + 1.f);
        (isy ? iy0 : ix0)[o] = (int)fl;
      }
      __syncthreads();
      for (int pix = warp; pix < HW; pix += QT / 32) {
        const int oy = pix / W, ox = pix - oy * W;
        int xi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xi[j] = min(max(ix0[ox] - 1 + j, 0), pw - 1);
        float rows[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int yy = min(max(iy0[oy] - 1 + i, 0), ph - 1);
          const int* ir = idx_s + yy * pw;
          float a = __ldg(p.codebook + (size_t)ir[xi[0]] * CV + lane) * wx[4 * ox + 0];
          a = a + __ldg(p.codebook + (size_t)ir[xi[1]] * CV + lane) * wx[4 * ox + 1];
          a = a + __ldg(p.codebook + (size_t)ir[xi[2]] * CV + lane) * wx[4 * ox + 2];
          a = a + __ldg(p.codebook + (size_t)ir[xi[3]] * CV + lane) * wx[4 * ox + 3];
          rows[i] = a;
        }
        float o = rows[0] * wy[4 * oy + 0];
        o = o + rows[1] * wy[4 * oy + 1];
        o = o + rows[2] * wy[4 * oy + 2];
        o = o + rows[3] * wy[4 * oy + 3];
        hup[lane * plane + (oy + 1) * PW + ox + 1] = o;
      }
    } else {
      for (int pix = warp; pix < HW; pix += QT / 32) {
        const int oy = pix / W, ox = pix - oy * W;
        hup[lane * plane + (oy + 1) * PW + ox + 1] = __ldg(p.codebook + (size_t)idx_s[pix] * CV + lane);
      }
    }
    {  // Phi weights of this scale (the region doubles as search scratch, so reload every scale)
      const int cur_phi = p.phi[si];
      const float* w = p.phi_w + (size_t)cur_phi * CV * CV * 9;
      for (int i = tid; i < CV * CV * 9; i += QT) {
        const int co = i / (CV * 9), r = i - co * CV * 9, ci = r / 9, tap = r - ci * 9;
        wphi[(ci * 9 + tap) * CV + co] = __ldg(w + i);
      }
      if (tid < CV) bphi[tid] = __ldg(p.phi_b + (size_t)cur_phi * CV + tid);
    }
    __syncthreads();

    // ---------------- Phi: h*(1-r) + conv3x3(h)*r ; f_hat += ; f_rest -= ----------------
    for (int item = tid; item < HW * 2; item += QT) {
      const int pix = item % HW, half = item / HW;
      const int y = pix / W, x = pix - y * W;
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = bphi[half * 16 + k];
      for (int ci = 0; ci < CV; ++ci) {
        const float* hp = hup + ci * plane + y * PW + x;  // top-left of the 3x3 window in padded coordinates
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float xv = hp[(tap / 3) * PW + (tap % 3)];
          const float4* w4 = reinterpret_cast<const float4*>(wphi + (ci * 9 + tap) * CV + half * 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 w = w4[k4];
            acc[4 * k4 + 0] = acc[4 * k4 + 0] + w.x * xv;
            acc[4 * k4 + 1] = acc[4 * k4 + 1] + w.y * xv;
            acc[4 * k4 + 2] = acc[4 * k4 + 2] + w.z * xv;
            acc[4 * k4 + 3] = acc[4 * k4 + 3] + w.w * xv;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int co = half * 16 + k;
        const float hv = hup[co * plane + (y + 1) * PW + x + 1];
        const float hphi = hv * p.one_minus_resi + acc[k] * p.resi;
        const size_t g = (size_t)co * HW + pix;
        const float nh = f_hat[g] + hphi;
        f_hat[g] = nh;
        if (upd_rest) f_rest[g] = f_rest[g] - hphi;
        if (p.fhat_list) p.fhat_list[((size_t)si * p.B + b) * img + g] = nh;
      }
    }
    __syncthreads();

    // ---------------- next-scale input: area(f_hat -> next patch grid) ----------------
    if (!last && (p.next_tokens || p.next_nchw)) {
      const int nh = p.ph[si + 1], nw = p.pw[si + 1], nl = nh * nw;
      for (int i = tid; i < nl * CV; i += QT) {
        const int t = i / CV, c = i - t * CV;
        const float v = area_at(f_hat + (size_t)c * HW, H, W, nh, nw, t / nw, t % nw);
        if (p.next_tokens) p.next_tokens[((size_t)b * p.next_stride + tok_off + t) * CV + c] = v;
        if (p.next_nchw) p.next_nchw[((size_t)b * CV + c) * nl + t] = v;
      }
      tok_off += nl;
    }
    if (p.idx_concat) idx_off += (long long)p.B * l;
    __syncthreads();
    if (p.split == 2 && !last) pool_out(si + 1);
  }
}

static size_t quant_smem_bytes(int H, int W, int V, bool encode) {
  const int plane = ((H + 2) * (W + 2)) | 1;
  const size_t words = (size_t)CV * plane + Q_UNION + (encode ? V : 0) + (size_t)H * W + 2 * Q_MAXHW * 4 + 2 * Q_MAXHW;
  return words * 4;
}

int quant_launch(const QuantArgs& a, cudaStream_t st) {
  VB_REQUIRE(a.Cvae == CV, "quant: Cvae=%d unsupported (kernel is specialised for 32)", a.Cvae);
  VB_REQUIRE(a.B > 0 && a.S > 0 && a.S <= VB_MAX_SCALES, "quant: bad B=%d S=%d", a.B, a.S);
  VB_REQUIRE(a.H > 0 && a.W > 0 && a.H <= Q_MAXHW && a.W <= Q_MAXHW, "quant: latent %dx%d unsupported (max %d)", a.H, a.W,
             Q_MAXHW);
  VB_REQUIRE(a.ph[a.S - 1] == a.H && a.pw[a.S - 1] == a.W, "patch_hws[-1]=(%d, %d) != (H=%d, W=%d)", a.ph[a.S - 1],
             a.pw[a.S - 1], a.H, a.W);  // quant.py:144
  VB_REQUIRE(a.codebook && a.phi_w && a.phi_b && a.f_hat && a.idx, "quant: null pointer");
  VB_REQUIRE(a.si_begin >= 0 && a.si_begin < a.si_end && a.si_end <= a.S, "quant: bad scale range [%d,%d)", a.si_begin,
             a.si_end);
  VB_REQUIRE(!a.f || a.f_rest, "quant: encode needs the f_rest workspace");
  VB_REQUIRE(a.split >= 0 && a.split <= 2, "quant: bad split mode %d", a.split);
  VB_REQUIRE(a.V > 0, "quant: empty codebook");
  QuantKParams p{};
  p.B = a.B; p.H = a.H; p.W = a.W; p.V = a.V; p.S = a.S;
  for (int i = 0; i < a.S; ++i) {
    VB_REQUIRE(a.ph[i] > 0 && a.pw[i] > 0 && a.ph[i] <= a.H && a.pw[i] <= a.W, "quant: bad patch size at scale %d", i);
    p.ph[i] = a.ph[i]; p.pw[i] = a.pw[i]; p.phi[i] = a.phi_of_scale[i];
    VB_REQUIRE(p.phi[i] >= 0 && p.phi[i] < a.n_phi, "quant: phi index out of range at scale %d", i);
  }
  p.resi = a.resi;
  p.one_minus_resi = (float)(1.0 - (double)a.resi);
  p.codebook = a.codebook; p.phi_w = a.phi_w; p.phi_b = a.phi_b;
  p.si_begin = a.si_begin; p.si_end = a.si_end;
  p.f = a.f; p.f_rest = a.f_rest; p.f_hat = a.f_hat; p.zero_fhat = a.zero_fhat;
  p.idx_concat = a.idx_concat;
  p.idx = reinterpret_cast<long long*>(a.idx);
  p.fhat_list = a.fhat_list; p.next_tokens = a.next_tokens; p.next_stride = a.next_stride; p.next_nchw = a.next_nchw;
  p.split = a.split; p.z_out = a.z_out; p.zb_out = reinterpret_cast<__nv_bfloat16*>(a.zb_out); p.zz_out = a.zz_out;
  VB_REQUIRE(a.split == 0 || (a.z_out && a.zb_out && a.zz_out && a.f_rest), "quant: split mode needs pooled-token buffers");
  const size_t smem = quant_smem_bytes(a.H, a.W, a.V, a.f != nullptr && a.split == 0);
  VB_REQUIRE(smem <= 227 * 1024, "quant: shared memory %zu exceeds 227 KB (V=%d too large?)", smem, a.V);
  static vb::SmemAttrCache attr_cache;
  if (vb::ensure_dyn_smem(attr_cache, smem, quant_kernel)) return vb::VB_ERR_CUDA;
  vb::ProfScope prof_scope(vb::PK_QUANT, st);
  quant_kernel<<<a.B, QT, smem, st>>>(p);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
