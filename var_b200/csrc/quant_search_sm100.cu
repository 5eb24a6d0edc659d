// Codebook nearest-neighbour search on the tensor cores (models/quant.py:155-157):
//   idx_n = argmin_v ( |z_n|^2 + |e_v|^2 - 2 z_n . e_v ),  first index on ties.
//
// tcgen05 has no fp32 MMA and bf16/TF32 distances flip ~0.1-0.6 % of the argmins, so the GEMM is used
// as a FILTER: pass 1 computes the bf16 distance matrix tile by tile in TMEM and keeps the row minimum; pass 2
// recomputes it and re-ranks, in exact fp32 with the oracle's operation order, every code whose bf16 distance lies
// within a rigorous error margin of that minimum (on average ~1-2 codes per token). The result is therefore
// bit-identical to the full fp32 search (and to oracle/quant_oracle.c) while >99.9 % of the 2*N*V*32 FLOPs run on
// the tensor pipe.
//
// CTA = 128 tokens (persistent over token tiles). warp 0: TMA producer, warp 1: UMMA issuer (128 x 256 x 16, K padded
// 32 -> 64 with zeros so operands are plain 128B-swizzled K-major tiles), warps 2-17: sixteen scan warps, four per TMEM
// lane quarter (thread = token row), which deal the eight 32-code chunks of a tile between them and meet once per pass
// to combine their row minima through shared memory. (With one warp per quarter the kernel issued one instruction per
// 4.4 cycles per scheduler - profiles/r02_ncu_quant_search.details.csv: 137 us per launch whatever the grid, tensor
// pipe 6 % active; four warps per scheduler hide that dependent-issue latency.)
#include "common.cuh"
#include "host.h"
#include "quant.h"

namespace vb {

constexpr int QS_BM = 128, QS_BN = 256, QS_STAGES = 3;
constexpr int QS_SUB = 4;                         // scan warps per TMEM lane quarter
constexpr int QS_THREADS = 64 + 128 * QS_SUB;     // TMA warp + UMMA warp + 16 scan warps
constexpr int QS_A_BYTES = QS_BM * 64 * 2;  // 16 KB
constexpr int QS_B_BYTES = QS_BN * 64 * 2;  // 32 KB
constexpr int QS_CV = 32;

__global__ void quant_prep_codebook_kernel(const float* __restrict__ cb, __nv_bfloat16* __restrict__ out,
                                           float* __restrict__ ee, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * 64) return;
  const int v = i >> 6, c = i & 63;
  out[i] = __float2bfloat16_rn(c < QS_CV ? cb[(size_t)v * QS_CV + c] : 0.f);
  if (c == 0) {  // |e_v|^2 in the oracle's order (sequential, separately rounded multiply and add)
    const float* e = cb + (size_t)v * QS_CV;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < QS_CV; ++k) { const float t = __ldg(e + k); s = __fadd_rn(s, __fmul_rn(t, t)); }
    ee[v] = s;
  }
}

__global__ void __launch_bounds__(QS_THREADS, 1)
quant_search_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ z, const float* __restrict__ zz, const float* __restrict__ codebook,
                    const float* __restrict__ ee_g, long long* __restrict__ idx_out, int N, int V) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 + 2 * QS_STAGES + 4];  // a_full | a_free | full[3] | empty[3] | tfull[2] | tempty[2]
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red[QS_THREADS / 32];
  __shared__ float emax_s;
  __shared__ float dmin_x[QS_SUB][QS_BM], best_x[QS_SUB][QS_BM];  // per-pass partial results of the scan warps of a row
  __shared__ int bi_x[QS_SUB][QS_BM];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + QS_A_BYTES;
  float* ee = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + QS_A_BYTES + QS_STAGES * QS_B_BYTES);  // [V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_af = smem_u32(&bars[0]), bar_ae = smem_u32(&bars[1]);
  auto full_bar = [&](int s) { return smem_u32(&bars[2 + s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[2 + QS_STAGES + s]); };
  auto tfull_bar = [&](int s) { return smem_u32(&bars[2 + 2 * QS_STAGES + s]); };
  auto tempty_bar = [&](int s) { return smem_u32(&bars[4 + 2 * QS_STAGES + s]); };
  const int n_mt = (N + QS_BM - 1) / QS_BM;
  const int n_nt = (V + QS_BN - 1) / QS_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(bar_af, 1);
    mbar_init(bar_ae, 1);
    for (int s = 0; s < QS_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * QS_SUB); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  // |e_v|^2 (precomputed once per codebook by quant_prep_codebook_kernel) into shared memory, and max_v |e_v|
  float emax2 = 0.f;
  for (int v = V + threadIdx.x; v < ((V + 255) & ~255); v += QS_THREADS) ee[v] = INFINITY;  // tail of the last code tile
  for (int v = threadIdx.x; v < V; v += QS_THREADS) {
    const float s = __ldg(ee_g + v);
    ee[v] = s;
    emax2 = fmaxf(emax2, s);
  }
  emax2 = warp_max(emax2);
  if (lane == 0) red[warp] = emax2;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < QS_THREADS / 32; ++i) m = fmaxf(m, red[i]);
    emax_s = sqrtf(m);
  }
  __syncthreads();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int mt = blockIdx.x; mt < n_mt; mt += gridDim.x, ++it) {
        mbar_wait(bar_ae, (it & 1) ^ 1);  // previous token tile fully consumed by the tensor pipe
        mbar_expect_tx(bar_af, QS_A_BYTES);
        tma_load_2d(&tmA, bar_af, sA, 0, mt * QS_BM);
        for (int pass = 0; pass < 2; ++pass)
          for (int nt = 0; nt < n_nt; ++nt) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), QS_B_BYTES);
            tma_load_2d(&tmB, full_bar(stage), sB + stage * QS_B_BYTES, 0, nt * QS_BN);
            if (++stage == QS_STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(QS_BM, QS_BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0, acc_it = 0;
      for (int mt = blockIdx.x; mt < n_mt; mt += gridDim.x, ++it) {
        mbar_wait(bar_af, it & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k_sw128(sA);
        for (int pass = 0; pass < 2; ++pass)
          for (int nt = 0; nt < n_nt; ++nt, ++acc_it) {
            const int as = acc_it & 1;
            mbar_wait(tempty_bar(as), ((acc_it >> 1) & 1) ^ 1);
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_k_sw128(sB + stage * QS_B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + as * 256, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
            umma_commit(empty_bar(stage));
            umma_commit(tfull_bar(as));
            if (++stage == QS_STAGES) { stage = 0; phase ^= 1; }
          }
        umma_commit(bar_ae);  // all MMAs reading sA have completed
      }
    }
  } else {
    const int quarter = warp & 3, sub = (warp - 2) >> 2;  // TMEM lane quarter; which of its QS_SUB scan warps
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int row_cta = quarter * 32 + lane;
    int acc_it = 0;
    for (int mt = blockIdx.x; mt < n_mt; mt += gridDim.x) {
      const int row = mt * QS_BM + row_cta;
      const bool row_ok = row < N;
      float zr[QS_CV];
      {
        const float4* z4 = reinterpret_cast<const float4*>(z + (size_t)(row_ok ? row : 0) * QS_CV);
#pragma unroll
        for (int c4 = 0; c4 < QS_CV / 4; ++c4) {
          const float4 t = __ldg(z4 + c4);
          zr[4 * c4] = t.x; zr[4 * c4 + 1] = t.y; zr[4 * c4 + 2] = t.z; zr[4 * c4 + 3] = t.w;
        }
      }
      const float zzr = __ldg(zz + (row_ok ? row : 0));
      // |d_bf16 - d_fp32| <= 2 |dot_bf16 - dot| <= 2 * (2*2^-9 + 2^-18) |z||e| (+ fp32 accumulation noise): two codes are
      // compared, so the filter margin is twice that (2^-6 |z| max|e|), taken with slack (0.02) plus an absolute floor
      // for the fp32 rounding of the exact distances themselves.
      const float margin = 0.02f * sqrtf(zzr) * emax_s + 2e-5f * (zzr + emax_s * emax_s) + 1e-30f;
      float dmin = INFINITY;   // pass 1: minimum bf16 distance (without the row-constant |z|^2)
      float best = INFINITY;   // pass 2: exact fp32 distance
      int bi = 0x7fffffff;
      for (int pass = 0; pass < 2; ++pass) {
        const float thr = dmin + margin;
        for (int nt = 0; nt < n_nt; ++nt, ++acc_it) {
          const int as = acc_it & 1;
          mbar_wait(tfull_bar(as), (acc_it >> 1) & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + as * 256 + lane_off;
#pragma unroll 1
          for (int c = sub; c < QS_BN / 32; c += QS_SUB) {
            float acc[32];
            __syncwarp();
            tmem_ld_32x32(taddr + c * 32, acc);
            tmem_ld_wait_dep(acc);
            const int v0 = nt * QS_BN + c * 32;
            if (v0 >= V) break;
            if (pass == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float d = fmaf(-2.f, acc[j], ee[v0 + j]);
                dmin = (v0 + j < V) ? fminf(dmin, d) : dmin;
              }
            } else {
              unsigned cand = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float d = fmaf(-2.f, acc[j], ee[v0 + j]);
                cand |= (d <= thr && v0 + j < V) ? (1u << j) : 0u;
              }
              while (cand) {  // ascending code order inside the chunk; across chunks the (distance, index) pair decides
                const int j = __ffs(cand) - 1;
                cand &= cand - 1;
                const int v = v0 + j;
                const float4* e4 = reinterpret_cast<const float4*>(codebook + (size_t)v * QS_CV);
                float dot = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < QS_CV / 4; ++c4) {
                  const float4 e = __ldg(e4 + c4);
                  dot = __fadd_rn(dot, __fmul_rn(zr[4 * c4], e.x));
                  dot = __fadd_rn(dot, __fmul_rn(zr[4 * c4 + 1], e.y));
                  dot = __fadd_rn(dot, __fmul_rn(zr[4 * c4 + 2], e.z));
                  dot = __fadd_rn(dot, __fmul_rn(zr[4 * c4 + 3], e.w));
                }
                const float d = __fsub_rn(__fadd_rn(zzr, ee[v]), __fmul_rn(2.f, dot));
                if (d < best || (d == best && v < bi)) { best = d; bi = v; }  // first minimum (torch.argmin)
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        // the QS_SUB scan warps of this row quarter combine their partial results (named barrier per quarter)
        if (pass == 0) {
          dmin_x[sub][row_cta] = dmin;
          named_bar_sync(1 + quarter, 32 * QS_SUB);
#pragma unroll
          for (int i = 0; i < QS_SUB; ++i) dmin = fminf(dmin, dmin_x[i][row_cta]);
        } else {
          best_x[sub][row_cta] = best;
          bi_x[sub][row_cta] = bi;
          named_bar_sync(1 + quarter, 32 * QS_SUB);
          if (sub == 0) {
#pragma unroll
            for (int i = 1; i < QS_SUB; ++i) {
              const float d = best_x[i][row_cta];
              const int v = bi_x[i][row_cta];
              if (d < best || (d == best && v < bi)) { best = d; bi = v; }
            }
            if (row_ok) idx_out[row] = bi == 0x7fffffff ? 0 : bi;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int quant_prepare_codebook(const float* codebook, void* cb_bf16, float* ee, int V, cudaStream_t st) {
  VB_REQUIRE(codebook && cb_bf16 && ee && V > 0, "quant_prepare_codebook: bad arguments");
  vb::ProfScope prof_scope(vb::PK_OTHER, st);
  quant_prep_codebook_kernel<<<(V * 64 + 255) / 256, 256, 0, st>>>(codebook, reinterpret_cast<__nv_bfloat16*>(cb_bf16), ee, V);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

int quant_search_launch(const QuantSearchArgs& a, cudaStream_t st) {
  VB_REQUIRE(a.zb && a.z && a.zz && a.cb_bf16 && a.codebook && a.ee && a.idx_out, "quant_search: null pointer");
  VB_REQUIRE(a.N > 0 && a.V > 0, "quant_search: bad N=%d V=%d", a.N, a.V);
  const size_t smem = QS_A_BYTES + QS_STAGES * QS_B_BYTES + (size_t)((a.V + 255) & ~255) * 4 + 1024;
  VB_REQUIRE(smem <= 227 * 1024, "quant_search: codebook of %d entries does not fit the |e|^2 table in shared memory", a.V);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {64, (uint64_t)a.N};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, QS_BM};
    int r = make_tmap_bf16_sw128(&tmA, a.zb, 2, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)a.V};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, QS_BN};
    int r = make_tmap_bf16_sw128(&tmB, a.cb_bf16, 2, dims, str, box);
    if (r) return r;
  }
  static SmemAttrCache attr_cache;
  if (ensure_dyn_smem(attr_cache, smem, quant_search_kernel)) return VB_ERR_CUDA;
  const int n_mt = (a.N + QS_BM - 1) / QS_BM;
  const int grid = n_mt < sm_count() ? n_mt : sm_count();
  vb::ProfScope prof_scope(vb::PK_QUANT, st);
  quant_search_kernel<<<grid, QS_THREADS, smem, st>>>(tmA, tmB, a.z, a.zz, a.codebook, a.ee,
                                                      reinterpret_cast<long long*>(a.idx_out), a.N, a.V);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
