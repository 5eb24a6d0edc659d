// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (fp32 accumulate in TMEM)
//   * A and W are K-major (row-major activations, nn.Linear weight layout) and arrive through TMA
//     (128B-swizzled 64-column boxes) into a multi-stage smem ring
//   * one elected thread issues tcgen05.mma: CTAS=1 -> UMMA 128 x BN x 16 (cta_group::1);
//     CTAS=2 -> a CTA pair (cluster of 2 SMs) computes a 256 x BN tile with cta_group::2: each CTA stages its own
//     128 rows of A and HALF of the W tile, so the per-SM L2->SMEM traffic per FLOP drops by a third (the 1-CTA
//     kernel is capped near 60 % tensor-active by that ingress, profiles/r01_gemm_1cta_ncu.txt)
//   * accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
//     the main loop of tile i+1
//   * eight epilogue warps (4 * GEMM_EPI_SUB) read TMEM with tcgen05.ld (one accumulator row per thread) and apply the
//     fused epilogue of the VAR block (reference semantics cited per mode in gemm.h)
#pragma once
#include "common.cuh"
#include "gemm.h"

namespace vb {

constexpr int GEMM_BM = 128;        // accumulator rows per CTA (TMEM lanes)
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 4 * GEMM_EPI_SUB;
// Warp roles by warpgroup: warpgroup 0 = {TMA producer, MMA issuer + TMEM owner, two idle warps}, warpgroups 1.. = the
// epilogue warps. The split is on warpgroup boundaries so that setmaxnreg can move registers from the service
// warpgroup (few live values) to the epilogue warps: with 12 warps every SM sub-partition hosts 3 of them, which caps
// a uniform allocation at 168 registers per thread; after the hand-over the epilogue warps own 224.
constexpr int GEMM_SVC_WARPS = 4;
constexpr int GEMM_THREADS = 32 * (GEMM_SVC_WARPS + GEMM_EPI_WARPS);
constexpr uint32_t GEMM_SVC_REGS = 72, GEMM_EPI_REGS = 216;  // 128 * svc + 256 * epi == 384 * 168 (the CTA's pool)
constexpr int EPI_STG_LD = 20;      // floats per staging row (16 + 4 pad: 16-byte aligned)

template <int BN, int CTAS>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = GEMM_EPI_WARPS * 32 * EPI_STG_LD * 4;  // epilogue staging, one 32x20 fp32 tile per warp
  static constexpr int BUDGET = 232448 - 2048 - STG_BYTES;  // 227 KB per CTA minus alignment slack / static smem
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024;  // +1024: manual alignment slack
};

// LNF = deferred LayerNorm ("adaLN without a LayerNorm pass"):
//   producer (EPI_GATE_RESID): besides x_new the epilogue writes a' = bf16(x_new * (1 + scale_next[seq])) and, per row,
//     tile and epilogue sub-warp, the partial (sum, sum of squares) of x_new;
//   consumer (EPI_QKV, EPI_GELU_BF16): the GEMM runs on a' and the epilogue finishes the normalisation,
//     LN(x)(1+s)+sh = rstd*(a'W^T - mean*U[label]) + V[label], U = W(1+s), V = W sh + bias precomputed per class.
template <int BN, int EPI, int CTAS, bool LNF = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN, CTAS>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TILE_M = GEMM_BM * CTAS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_smem;
  // GroupNorm statistics of the conv epilogue: [sub-warp][chunk parity][lane quarter][channel of the chunk] (sum, sumsq)
  __shared__ float2 gn_red[EPI == EPI_BIAS_BF16 ? GEMM_EPI_SUB * 2 * 4 * 32 : 1];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (CTAS == 2) ? (int)cluster_ctarank() : 0;  // CTA 0 of the pair issues the MMAs
  const int cid = blockIdx.x / CTAS, ncl = gridDim.x / CTAS;   // tile scheduler runs per cluster

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[STAGES + s]); };
  auto tfull_bar = [&](int s) { return smem_u32(&bars[2 * STAGES + s]); };
  auto tempty_bar = [&](int s) { return smem_u32(&bars[2 * STAGES + 2 + s]); };

  const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // the (leader) producer's arrive.expect_tx; TMA bytes of both CTAs complete it
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit (multicast to both CTAs when CTAS == 2)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), GEMM_EPI_WARPS * CTAS);  // one arrive per epilogue warp of every CTA of the pair
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) tmem_alloc_cg2(smem_u32(&tmem_base_smem), 512);
    else tmem_alloc(smem_u32(&tmem_base_smem), 512);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // PDL: the set-up above overlapped the previous kernel's tail; from here on global memory is read and written
  pdl_wait();
  pdl_launch_dependents();

  // The two register regimes never merge again: each branch carries its own copy of the teardown and returns.
  auto teardown = [&]() {
    tc_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
      tc_fence_after();
      if constexpr (CTAS == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
  };
  if (warp < GEMM_SVC_WARPS) {
  setmaxnreg_dec<GEMM_SVC_REGS>();
  if (warp == 0) {
    // ------------------------------ TMA producer (every CTA loads its own operand slices) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t hint_a = p.l2_hint_a ? p.l2_hint_a : L2_EVICT_NORMAL, hint_w = p.l2_hint_w ? p.l2_hint_w : L2_EVICT_NORMAL;
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int m0 = m_blk * TILE_M + rank * GEMM_BM;  // first accumulator row (pixel) of this CTA
        // conv mode: the 128 rows are a bw x bh patch of one image; tap (dy,dx) reads the patch shifted by (dy-1,dx-1),
        // rows/columns outside the image come back as zeros from TMA (= the convolution's zero padding)
        int cx0 = 0, cy0 = 0, cb = 0;
        if (p.conv_kpt > 0) {
          const int hw = p.conv_H * p.conv_W;
          cb = m0 / hw;
          const int rem = m0 - cb * hw;
          cy0 = rem / p.conv_W;
          cx0 = rem - cy0 * p.conv_W;
        }
        // input coordinate of tap (0,0): stride 1 has zero padding 1 all round (x0-1, y0-1); the stride-2 convolution
        // pads only right / bottom (2*x0, 2*y0)
        const int cxs = p.conv_stride > 1 ? cx0 * p.conv_stride : cx0 - 1;
        const int cys = p.conv_stride > 1 ? cy0 * p.conv_stride : cy0 - 1;
        // block-diagonal batching: the W rows / columns this A row block multiplies (0, 0 for a plain GEMM)
        const int bd = p.bd_rows > 0 ? (m_blk * TILE_M) / p.bd_rows : 0;
        const int wn0 = n_blk * BN + bd * p.bd_w_row, wk0 = bd * p.bd_w_k;
        int tap = 0, kc = 0;  // conv mode: filter tap and channel block of k-block kb
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if constexpr (CTAS == 2) {
            // completion bytes of BOTH CTAs are credited to the leader's full barrier (it gates the pair's MMAs)
            const uint32_t fb = mapa_shared(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            if (p.conv_kpt > 0)
              tma_load_4d_cg2(&tmA, fb, sa, kc * GEMM_BK, cxs + tap % 3, cys + tap / 3, cb);
            else
              tma_load_2d_cg2_hint(&tmA, fb, sa, kb * GEMM_BK, m0, hint_a);
            tma_load_2d_cg2_hint(&tmB, fb, sa + Cfg::A_BYTES, wk0 + kb * GEMM_BK, wn0 + rank * (BN / 2), hint_w);
          } else {
            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            if (p.conv_kpt > 0)
              tma_load_4d(&tmA, full_bar(stage), sa, kc * GEMM_BK, cxs + tap % 3, cys + tap / 3, cb);
            else
              tma_load_2d_hint(&tmA, full_bar(stage), sa, kb * GEMM_BK, m0, hint_a);
            tma_load_2d_hint(&tmB, full_bar(stage), sa + Cfg::A_BYTES, wk0 + kb * GEMM_BK, wn0, hint_w);
          }
          if (++kc == p.conv_kpt) { kc = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            if constexpr (CTAS == 2) umma_bf16_ss_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          // smem slot reusable (in both CTAs) once these MMAs have read it
          if constexpr (CTAS == 2) umma_commit_cg2_mc(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete: wake the epilogue warps of both CTAs
        if constexpr (CTAS == 2) umma_commit_cg2_mc(tfull_bar(as), 3); else umma_commit(tfull_bar(as));
      }
    }
  }
  teardown();
  return;
  } else {
    setmaxnreg_inc<GEMM_EPI_REGS>();
    // ------------------------------ epilogue warps ------------------------------
    // GEMM_EPI_SUB warps per TMEM lane quarter, dealing the column chunks of the tile round-robin between them. The
    // epilogue is latency bound (one accumulator row per thread): the GELU epilogue costs a K=1024 tile 14 % (d16 fc1:
    // 1206 TFLOP/s against 1408 with the plain bias epilogue, tools/gemm_k_sweep.py); a third warp per quarter wins
    // back only 2-3 % of that and loses it again in the QKV epilogue (see gemm.h).
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - GEMM_SVC_WARPS) >> 2;  // which of the GEMM_EPI_SUB warps of that quarter (chunk c = half, half+SUB, ..)
    int it = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row = m_blk * TILE_M + rank * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + as * 256 + ((uint32_t)(quarter * 32) << 16);
      const int n_base = n_blk * BN;
      // private 32 x 20-float staging tile of this warp (transposes accumulator rows into coalesced global accesses)
      float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * Cfg::STAGE_BYTES) +
                   (warp - GEMM_SVC_WARPS) * (32 * EPI_STG_LD);
      const int rsub = lane >> 2, cg = lane & 3;
      const int row_w0 = m_blk * TILE_M + rank * GEMM_BM + quarter * 32;  // first row of this warp
      // deferred LayerNorm, consumer side: this thread's row statistics from the producer's partial sums, and the
      // rows of the per-class tables U, V (requested before the accumulator barrier: their latency overlaps the wait)
      float ln_rstd = 1.f, ln_nmr = 0.f;  // rstd and -mean * rstd
      const float* ln_urow = nullptr;
      const float* ln_vrow = nullptr;
      if constexpr (LNF && (EPI == EPI_QKV || EPI == EPI_GELU_BF16)) {
        const int rr = row_ok ? row : p.M - 1;
        const float2* pp = p.ln_part_in + (size_t)rr * p.ln_parts;
        float sx = 0.f, sxx = 0.f;
        for (int j = 0; j < p.ln_parts; ++j) {
          const float2 t = __ldg(pp + j);
          sx += t.x;
          sxx += t.y;
        }
        const float inv_c = 1.f / (float)p.ln_C;
        const float mean = sx * inv_c;
        ln_rstd = rsqrtf(fmaxf(sxx * inv_c - mean * mean, 0.f) + p.ln_eps);
        ln_nmr = -mean * ln_rstd;
        const size_t lab = (size_t)__ldg(p.ln_labels + rr / p.rows_per_seq) * p.N;
        ln_urow = p.ln_u + lab;
        ln_vrow = p.ln_v + lab;
      }
      // 32 columns of the row's U / V vectors. The window is consumed at the very start of a chunk (two FMAs per element)
      // and refilled for the next chunk right away, so the L2 latency of the refill hides behind the rest of the chunk's
      // work (GELU / normalisation, packing, stores); the first window of a tile is requested before the accumulator
      // barrier. No second register set is needed.
      float4 lq_u[8], lq_v[8];
      auto ln_fetch = [&](int col) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lq_u[j] = __ldg(reinterpret_cast<const float4*>(ln_urow + col) + j);
          lq_v[j] = __ldg(reinterpret_cast<const float4*>(ln_vrow + col) + j);
        }
      };
      auto ln_apply = [&](float* v) {  // v[0..32) = rstd * v - mean * rstd * U + V
        const float2 r2 = make_float2(ln_rstd, ln_rstd), m2 = make_float2(ln_nmr, ln_nmr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 t0 = __ffma2_rn(m2, make_float2(lq_u[j].x, lq_u[j].y), make_float2(lq_v[j].x, lq_v[j].y));
          const float2 t1 = __ffma2_rn(m2, make_float2(lq_u[j].z, lq_u[j].w), make_float2(lq_v[j].z, lq_v[j].w));
          const float2 a0 = __ffma2_rn(make_float2(v[4 * j], v[4 * j + 1]), r2, t0);
          const float2 a1 = __ffma2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), r2, t1);
          v[4 * j] = a0.x; v[4 * j + 1] = a0.y; v[4 * j + 2] = a1.x; v[4 * j + 3] = a1.y;
        }
      };
      if constexpr (LNF && (EPI == EPI_QKV || EPI == EPI_GELU_BF16)) {
        const int first = n_base + half * (EPI == EPI_QKV ? 64 : 32);
        if (first < p.N) ln_fetch(first);
      }

      mbar_wait_nocall(tfull_bar(as), aphase);
      tc_fence_after();

      if constexpr (EPI == EPI_QKV) {
        // destination row offsets of the four row groups this lane stores (fixed for the whole tile)
        size_t q_off[4], kv_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rg = row_w0 + rsub + 8 * i;
          const int sq = rg / p.rows_per_seq, tt = rg - sq * p.rows_per_seq;
          q_off[i] = ((size_t)sq * p.H * p.rows_per_seq + tt) * 64;
          kv_off[i] = ((size_t)sq * p.H * p.Lmax + p.pos0 + tt) * 64;
        }
#pragma unroll 1
        for (int c = half; c < BN / 64; c += GEMM_EPI_SUB) {
          const int n0 = n_base + c * 64;
          if (n0 >= p.N) break;
          float v[64];
          __syncwarp();
          tmem_ld_32x32(taddr + c * 64, v);
          tmem_ld_32x32(taddr + c * 64 + 32, v + 32);
          const int which = n0 / p.C;  // 0 q, 1 k, 2 v
          const int head = (n0 - which * p.C) >> 6;
          const float qs = which == 0 ? __ldg(p.q_scale + head) : 1.f;
          float2 ss2 = make_float2(0.f, 0.f);
          if constexpr (LNF) {
            // qkv = rstd * acc - mean * rstd * U[label] + V[label]   (V holds W sh + [q_bias, 0, v_bias]); 32 columns at a time
            tmem_ld_wait_dep(v);
            ln_apply(v);
            ln_fetch(n0 + 32);
#pragma unroll
            for (int j = 0; j < 32; j += 2) ss2 = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(v[j], v[j + 1]), ss2);
            tmem_ld_wait_dep(v + 32);
            ln_apply(v + 32);
            if (c + GEMM_EPI_SUB < BN / 64 && n0 + GEMM_EPI_SUB * 64 < p.N) ln_fetch(n0 + GEMM_EPI_SUB * 64);  // next chunk
#pragma unroll
            for (int j = 32; j < 64; j += 2) ss2 = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(v[j], v[j + 1]), ss2);
          } else {
            float4 bq[16];  // the chunk's bias, requested while the accumulator is still on its way from TMEM
#pragma unroll
            for (int j = 0; j < 16; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
            tmem_ld_wait_dep(v);
            tmem_ld_wait_dep(v + 32);
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const float4 b = bq[j >> 2];
              const float2 a0 = __fadd2_rn(make_float2(v[j], v[j + 1]), make_float2(b.x, b.y));
              const float2 a1 = __fadd2_rn(make_float2(v[j + 2], v[j + 3]), make_float2(b.z, b.w));
              ss2 = __ffma2_rn(a0, a0, ss2);
              ss2 = __ffma2_rn(a1, a1, ss2);
              v[j] = a0.x; v[j + 1] = a0.y; v[j + 2] = a1.x; v[j + 3] = a1.y;
            }
          }
          float mul = 1.f;
          if (p.no_l2norm) mul = qs;  // attn_l2_norm=False: softmax scale folded into q, k as it is (qs = 1 for k, v)
          else if (which < 2) mul = qs / fmaxf(sqrtf(ss2.x + ss2.y), 1e-12f);  // F.normalize(dim=-1), eps 1e-12
          const float2 mul2 = make_float2(mul, mul);
          const size_t head_off = (size_t)head * (which == 0 ? p.rows_per_seq : p.Lmax) * 64;
          __nv_bfloat16* const dst_base = (which == 0 ? p.q_out : (which == 1 ? p.k_cache : p.v_cache)) + head_off;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {  // 64-byte halves of the 128-byte head row
            uint4* srow = reinterpret_cast<uint4*>(stg + lane * EPI_STG_LD);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 32 * h2 + 8 * j;
              const float2 p0 = __fmul2_rn(make_float2(v[e + 0], v[e + 1]), mul2);
              const float2 p1 = __fmul2_rn(make_float2(v[e + 2], v[e + 3]), mul2);
              const float2 p2 = __fmul2_rn(make_float2(v[e + 4], v[e + 5]), mul2);
              const float2 p3 = __fmul2_rn(make_float2(v[e + 6], v[e + 7]), mul2);
              uint4 o;
              o.x = pack_bf16x2(p0.x, p0.y);
              o.y = pack_bf16x2(p1.x, p1.y);
              o.z = pack_bf16x2(p2.x, p2.y);
              o.w = pack_bf16x2(p3.x, p3.y);
              srow[j] = o;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rsub + 8 * i;
              if (row_w0 + r < p.M)
                *reinterpret_cast<uint4*>(dst_base + (which == 0 ? q_off[i] : kv_off[i]) + 32 * h2 + cg * 8) =
                    *reinterpret_cast<const uint4*>(stg + r * EPI_STG_LD + cg * 4);
            }
            __syncwarp();
          }
        }
      } else if constexpr (EPI == EPI_SCORE) {
        float run_max = -INFINITY, run_sum = 0.f;
        const int gt = row_ok ? __ldg(p.gt + (row % p.gt_mod)) : -1;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += GEMM_EPI_SUB) {
          const int n0 = n_base + c * 32;
          if (n0 >= p.N) break;
          float v[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait_dep(v);
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] += __ldg(p.bias + n0 + j);
            cmax = fmaxf(cmax, v[j]);
          }
          if (gt >= n0 && gt < n0 + 32) {
            float g = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) g = (gt - n0 == j) ? v[j] : g;
            p.gt_logit[row] = g;
          }
          const float new_max = fmaxf(run_max, cmax);
          float sacc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sacc += __expf(v[j] - new_max);
          run_sum = run_sum * __expf(run_max - new_max) + sacc;
          run_max = new_max;
        }
        // one partial per (row, tile, sub-warp); an empty share contributes (-inf, 0)
        if (row_ok) p.part[((size_t)row * n_tiles + n_blk) * GEMM_EPI_SUB + half] = make_float2(run_max, run_sum);
      } else if constexpr (EPI == EPI_GATE_RESID) {
        // Residual update needs coalesced reads of resid/gate: transpose 32x16 accumulator pieces through a private
        // smem tile so that each warp instruction touches 8 rows x 64 contiguous bytes. (Issuing the residual reads a
        // chunk ahead / before the accumulator barrier was measured 3-10 % SLOWER in an A/B on one box: the extra
        // loads in flight compete with the TMA operand stream for the SM's L2 ingress.)
        int seqs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // rows beyond M (tile tail) are clamped: their loads stay in bounds, nothing is stored
          const int rg = row_w0 + rsub + 8 * i;
          seqs[i] = (rg < p.M ? rg : p.M - 1) / p.rows_per_seq;
        }
        float ln_s1[4] = {0.f, 0.f, 0.f, 0.f}, ln_s2[4] = {0.f, 0.f, 0.f, 0.f};  // LNF: partial row sums of x_new
        // The four row groups of a lane nearly always belong to one sequence (rows_per_seq >= 32 and aligned): gate and
        // the next LayerNorm's scale are then one load per 16-column half instead of four, and the loads of BOTH halves
        // of a chunk (bias, gate, scale, 8 residual vectors) are requested together, right behind the TMEM load, so that
        // one memory round trip covers the chunk instead of two (same-box A/B, tools/gemm_one.py: proj + LN outputs d30
        // 1065 -> 985 us, d16 275 -> 242 us; fc2 unchanged or better). Requesting the NEXT chunk's loads as well was
        // measured too and is slower (d30 proj 1074-1240 us): that many loads in flight compete with the TMA operand
        // stream for the SM's L2 ingress.
        const bool one_seq = seqs[0] == seqs[3];
#pragma unroll 1
        for (int c = half; c < BN / 32; c += GEMM_EPI_SUB) {
          const int n0 = n_base + c * 32;
          if (n0 >= p.N) break;
          float v[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c * 32, v);
          float4 b4[2], g1[2], sc1[2], rs4[2][4];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int col = n0 + 16 * h2 + 4 * cg;
            b4[h2] = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            g1[h2] = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)seqs[0] * p.gate_ld + col));
            if constexpr (LNF) sc1[h2] = __ldg(reinterpret_cast<const float4*>(p.ln_scale + (size_t)seqs[0] * p.gate_ld + col));
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // all loads of the chunk before its stores: resid may alias out
              const int rg = row_w0 + rsub + 8 * i;
              if (rg < p.M) rs4[h2][i] = *reinterpret_cast<const float4*>(p.resid + (size_t)rg * p.N + col);
            }
          }
          tmem_ld_wait_dep(v);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float4* srow = reinterpret_cast<float4*>(stg + lane * EPI_STG_LD);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              srow[j] = make_float4(v[16 * h2 + 4 * j], v[16 * h2 + 4 * j + 1], v[16 * h2 + 4 * j + 2], v[16 * h2 + 4 * j + 3]);
            __syncwarp();
            const int col = n0 + 16 * h2 + 4 * cg;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rsub + 8 * i;
              const int rg = row_w0 + r;
              if (rg < p.M) {
                float4 g = g1[h2], sc = sc1[h2];
                if (!one_seq) {  // a warp that straddles a sequence boundary (small scales): per-row-group vectors
                  g = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)seqs[i] * p.gate_ld + col));
                  if constexpr (LNF) sc = __ldg(reinterpret_cast<const float4*>(p.ln_scale + (size_t)seqs[i] * p.gate_ld + col));
                }
                const float4 a = *reinterpret_cast<const float4*>(stg + r * EPI_STG_LD + 4 * cg);
                const float4 rs = rs4[h2][i];
                const float4 b = b4[h2];
                const float4 o = make_float4(fmaf(a.x + b.x, g.x, rs.x), fmaf(a.y + b.y, g.y, rs.y),
                                             fmaf(a.z + b.z, g.z, rs.z), fmaf(a.w + b.w, g.w, rs.w));
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)rg * p.N + col) = o;
                if constexpr (LNF) {
                  ln_s1[i] += (o.x + o.y) + (o.z + o.w);
                  ln_s2[i] += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
                  *reinterpret_cast<uint2*>(p.ln_a_out + (size_t)rg * p.N + col) =
                      make_uint2(pack_bf16x2(fmaf(o.x, sc.x, o.x), fmaf(o.y, sc.y, o.y)),
                                 pack_bf16x2(fmaf(o.z, sc.z, o.z), fmaf(o.w, sc.w, o.w)));
                }
              }
            }
            __syncwarp();
          }
        }
        if constexpr (LNF) {
          // one partial (sum, sum of squares) per row, N tile and sub-warp; a sub-warp without columns writes zeros
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = ln_s1[i], b = ln_s2[i];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            b += __shfl_xor_sync(0xffffffffu, b, 2);
            const int rg = row_w0 + rsub + 8 * i;
            if (cg == 0 && rg < p.M)
              p.ln_part_out[(size_t)rg * (n_tiles * GEMM_EPI_SUB) + n_blk * GEMM_EPI_SUB + half] = make_float2(a, b);
          }
        }
      } else {
        // software-pipelined: the TMEM load of chunk i+1 is in flight while chunk i is converted and stored
        constexpr int NCH = (BN / 32 + GEMM_EPI_SUB - 1) / GEMM_EPI_SUB;  // chunks per warp (dealt round-robin)
        float vbuf[2][32];
        float4 bbuf[2][8];  // bias of the chunk, fetched one chunk ahead like the accumulator (its latency was the
                            // largest long-scoreboard stall of the epilogue warps)
        int n_valid = 0;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          if ((half + GEMM_EPI_SUB * i) * 32 < BN && n_base + (half + GEMM_EPI_SUB * i) * 32 < p.N) n_valid = i + 1;
        const bool has_bias = !(LNF && EPI == EPI_GELU_BF16) && p.bias != nullptr;
        if (n_valid > 0) {
          __syncwarp();
          tmem_ld_32x32(taddr + half * 32, vbuf[0]);
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bbuf[0][j] = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + half * 32) + j);
          }
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          if (i >= n_valid) break;
          float* v = vbuf[i & 1];
          const int n0 = n_base + (half + GEMM_EPI_SUB * i) * 32;
          tmem_ld_wait_dep(v);
          if (i + 1 < n_valid) {
            __syncwarp();
            tmem_ld_32x32(taddr + (half + GEMM_EPI_SUB * (i + 1)) * 32, vbuf[(i + 1) & 1]);
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                bbuf[(i + 1) & 1][j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + GEMM_EPI_SUB * 32) + j);
            }
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = bbuf[i & 1][j];
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if constexpr (LNF && EPI == EPI_GELU_BF16) {  // fc1 pre-activation = rstd * acc - mean * rstd * U + V
            ln_apply(v);
            if (i + 1 < n_valid) ln_fetch(n0 + GEMM_EPI_SUB * 32);
          }
          if (!row_ok) {
            // nothing to store for rows beyond M (the TMEM loads stay warp-convergent)
          } else if constexpr (EPI == EPI_BIAS_F32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.N + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 g = gelu_tanh2(make_float2(v[j], v[j + 1]));
              v[j] = g.x;
              v[j + 1] = g.y;
            }
          }
          if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16) {
            // transpose the 32 rows x 64 B of this chunk through the warp's staging tile: each store instruction then
            // writes 8 rows x 64 contiguous bytes (whole sectors) instead of 32 scattered 16-byte pieces
            uint4* srow = reinterpret_cast<uint4*>(stg + lane * EPI_STG_LD);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              srow[j] = w;
            }
            __syncwarp();
            float gs[8], gq[8];  // GroupNorm statistics (conv epilogue): this lane's 8 channels over its 4 rows
#pragma unroll
            for (int e = 0; e < 8; ++e) { gs[e] = 0.f; gq[e] = 0.f; }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rsub + 8 * i;
              const int rg = row_w0 + r;
              if (rg < p.M) {
                uint4 w = *reinterpret_cast<const uint4*>(stg + r * EPI_STG_LD + cg * 4);
                const size_t off = (size_t)rg * p.N + n0 + cg * 8;
                if constexpr (EPI == EPI_BIAS_BF16) {
                  if (p.resid_bf16 != nullptr) {  // bf16 + bf16 like the eager `x + h` of the reference's ResnetBlock
                    const uint4 rr = *reinterpret_cast<const uint4*>(p.resid_bf16 + off);
                    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&w);
                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                    for (int e = 0; e < 4; ++e) a2[e] = __hadd2(a2[e], b2[e]);
                  }
                  if (p.gn_part != nullptr) {  // statistics of exactly the values that are stored
                    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 f = __bfloat1622float2(a2[e]);
                      gs[2 * e] += f.x; gq[2 * e] = fmaf(f.x, f.x, gq[2 * e]);
                      gs[2 * e + 1] += f.y; gq[2 * e + 1] = fmaf(f.y, f.y, gq[2 * e + 1]);
                    }
                  }
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = w;
              }
            }
            __syncwarp();
            if constexpr (EPI == EPI_BIAS_BF16) {
              if (p.gn_part != nullptr) {
                // rows of the warp: lanes with equal cg (xor 4, 8, 16); then the four lane quarters of the tile meet in
                // shared memory (named barrier per sub-warp group) and quarter 0 writes the 128-row partial, fixed order
#pragma unroll
                for (int e = 0; e < 8; ++e) {
#pragma unroll
                  for (int o = 4; o <= 16; o <<= 1) {
                    gs[e] += __shfl_xor_sync(0xffffffffu, gs[e], o);
                    gq[e] += __shfl_xor_sync(0xffffffffu, gq[e], o);
                  }
                }
                float2* red = gn_red + ((half * 2 + (i & 1)) * 4) * 32;
                if (lane < 4) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) red[quarter * 32 + cg * 8 + e] = make_float2(gs[e], gq[e]);
                }
                named_bar_sync(1 + half, 128);
                if (quarter == 0 && n0 + lane < p.N) {
                  float2 t = red[lane];
#pragma unroll
                  for (int q = 1; q < 4; ++q) { t.x += red[q * 32 + lane].x; t.y += red[q * 32 + lane].y; }
                  p.gn_part[(size_t)(row_w0 / GEMM_BM) * p.N + n0 + lane] = t;
                }
              }
            }
          }
        }
      }
      // release this accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // relaxed: the accumulator values are in registers already; do not wait for this tile's global stores
        if constexpr (CTAS == 2) mbar_arrive_relaxed_cluster(mapa_shared(tempty_bar(as), 0));  // the leader CTA's MMA thread waits
        else mbar_arrive_relaxed(tempty_bar(as));
      }
    }
    teardown();
  }
}

}  // namespace vb
