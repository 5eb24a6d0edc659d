// Block-causal attention for the VAR token pyramid on sm_100a (tcgen05 / TMEM / TMA).
//
// Reference semantics: models/basic_var.py:98-117 with the mask of models/var.py:107-112
//   out = softmax(q_hat k_hat^T + bias) v,  bias[i,j] = 0 if level(i) >= level(j) else -inf, softmax scale 1
// (q_hat / k_hat are already L2-normalised and scaled by the QKV GEMM epilogue). The mask is never materialised:
// a query at pyramid level s sees keys [0, level_end[s]), so each CTA only visits the key tiles below the
// largest level end among its 128 query rows and masks per row inside the last tiles. The KV-cached decode step
// (attn_bias=None, keys = all cached + current scale) is the same kernel with every query on the newest level.
//
// One CTA = 128 query rows of one (sequence, head). Warps 0-3: softmax (one row per thread == one TMEM lane),
// warp 4: TMA producer + UMMA issuer (one elected thread).
//   S = Q K^T   : UMMA 128x64x16 x4, Q and K tiles K-major SW128 in smem, S in TMEM columns [0,64)
//   P = exp2(..) : registers -> bf16 -> smem (K-major SW128, A operand of the second MMA)
//   O_j = P V   : UMMA 128x64x16 x4, V tile in its natural [key, d] layout = MN-major B operand, TMEM [64,128)
//   running output is kept in registers and rescaled on-line (flash-attention recurrence).
#include "attn.h"
#include "common.cuh"
#include "host.h"

namespace vb {

constexpr int ATT_BM = 128;  // query rows per CTA
constexpr int ATT_BN = 64;   // keys per tile
constexpr int ATT_D = 64;    // head dim
constexpr int ATT_THREADS = 160;
constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;   // 16 KB
constexpr int ATT_KV_BYTES = ATT_BN * ATT_D * 2;  // 8 KB
constexpr int ATT_P_BYTES = ATT_BM * ATT_BN * 2;  // 16 KB
constexpr int ATT_SMEM = ATT_Q_BYTES + 4 * ATT_KV_BYTES + ATT_P_BYTES + 1024;

struct AttnLevels {
  int n;
  int end[VB_MAX_SCALES];  // cumulative token count after each scale
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128_attn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;  // LBO (distance between 64-element N chunks; single chunk here)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 key rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int Lq, int H, int q_pos0,
            const AttnLevels lv) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[8];  // 0:q 1,2:k[2] 3,4:v[2] 5:s 6:p 7:o
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + ATT_Q_BYTES;
  const uint32_t sV = sK + 2 * ATT_KV_BYTES;
  const uint32_t sP = sV + 2 * ATT_KV_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int bh = seq * H + head;
  const uint32_t bar_q = smem_u32(&bars[0]), bar_s = smem_u32(&bars[5]), bar_p = smem_u32(&bars[6]),
                 bar_o = smem_u32(&bars[7]);
  auto bar_k = [&](int s) { return smem_u32(&bars[1 + s]); };
  auto bar_v = [&](int s) { return smem_u32(&bars[3 + s]); };

  // visible keys for the rows of this tile
  const int row0 = q_tile * ATT_BM;
  auto kv_end_of = [&](int row) {  // row index inside this call's query block
    const int pos = q_pos0 + (row < Lq ? row : Lq - 1);
    int e = lv.end[lv.n - 1];
    for (int s = lv.n - 1; s >= 0; --s)
      if (pos < lv.end[s]) e = lv.end[s];
    return e;
  };
  const int last_row = (row0 + ATT_BM - 1 < Lq) ? row0 + ATT_BM - 1 : Lq - 1;
  const int kv_max = kv_end_of(last_row);
  const int n_kt = (kv_max + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_k(s), 1); mbar_init(bar_v(s), 1); }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_smem), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);  // B (V) is MN-major
      mbar_expect_tx(bar_q, ATT_Q_BYTES);
      tma_load_3d(&tmQ, bar_q, sQ, 0, row0, bh);
      for (int j = 0; j < 2 && j < n_kt; ++j) {
        mbar_expect_tx(bar_k(j), ATT_KV_BYTES);
        tma_load_3d(&tmK, bar_k(j), sK + j * ATT_KV_BYTES, 0, j * ATT_BN, bh);
        mbar_expect_tx(bar_v(j), ATT_KV_BYTES);
        tma_load_3d(&tmV, bar_v(j), sV + j * ATT_KV_BYTES, 0, j * ATT_BN, bh);
      }
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kt; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // ---- S = Q K_j^T ----
        mbar_wait(bar_k(st), ph);
        tc_fence_after();
        {
          const uint64_t qd = umma_desc_k_sw128(sQ);
          const uint64_t kd = umma_desc_k_sw128(sK + st * ATT_KV_BYTES);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tmem, qd + 2 * k, kd + 2 * k, idesc_qk, k != 0);
        }
        umma_commit(bar_s);
        // ---- O_j = P_j V_j ----
        mbar_wait(bar_p, j & 1);  // softmax wrote P_j (and finished reading S_j, O_{j-1})
        tc_fence_after();
        // QK_j has completed (bar_s fired before bar_p could): K stage st is free -> prefetch K_{j+2}
        if (j + 2 < n_kt) {
          mbar_expect_tx(bar_k(st), ATT_KV_BYTES);
          tma_load_3d(&tmK, bar_k(st), sK + st * ATT_KV_BYTES, 0, (j + 2) * ATT_BN, bh);
        }
        mbar_wait(bar_v(st), ph);
        tc_fence_after();
        {
          const uint64_t pd = umma_desc_k_sw128(sP);
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            const uint64_t vd = umma_desc_mn_sw128_attn(sV + st * ATT_KV_BYTES + k * 2048);
            umma_bf16_ss(tmem + 64, pd + 2 * k, vd, idesc_pv, k != 0);
          }
        }
        umma_commit(bar_o);
        if (j + 2 < n_kt) {
          // V stage st is reusable once PV_j has read it
          mbar_wait(bar_o, j & 1);
          mbar_expect_tx(bar_v(st), ATT_KV_BYTES);
          tma_load_3d(&tmV, bar_v(st), sV + st * ATT_KV_BYTES, 0, (j + 2) * ATT_BN, bh);
        }
      }
    }
  } else {
    // ------------------------------ softmax / output warps ------------------------------
    const int row = row0 + warp * 32 + lane;
    const int kv_end = kv_end_of(row);
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t p_row = sP + (uint32_t)(warp * 32 + lane) * 128;
    const int sw = lane & 7;  // == row & 7
    constexpr float LOG2E = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[ATT_D];
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) o_acc[d] = 0.f;

    for (int j = 0; j < n_kt; ++j) {
      const int k0 = j * ATT_BN;
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // pass A: row maximum over the visible keys of this tile
      float m_tile = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s[32];
        __syncwarp();
        tmem_ld_32x32(t_lane + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = (k0 + c * 32 + i < kv_end) ? s[i] : -INFINITY;
          m_tile = fmaxf(m_tile, v);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);  // finite from the first tile on (key 0 is always visible)
      const float alpha = exp2f((m_run - m_new) * LOG2E);
      const float mneg = -m_new * LOG2E;
      float l_tile = 0.f;
      // pass B: probabilities -> bf16 -> swizzled smem (A operand of P V)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s[32];
        __syncwarp();
        tmem_ld_32x32(t_lane + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = (k0 + c * 32 + i < kv_end) ? exp2f(fmaf(s[i], LOG2E, mneg)) : 0.f;
          l_tile += p;
          s[i] = p;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 16-byte chunk (8 keys) index c*4+g, XOR-swizzled with row%8
          const uint32_t addr = p_row + (uint32_t)(((c * 4 + g) ^ sw) << 4);
          const uint32_t w0 = pack_bf16x2(s[8 * g + 0], s[8 * g + 1]), w1 = pack_bf16x2(s[8 * g + 2], s[8 * g + 3]);
          const uint32_t w2 = pack_bf16x2(s[8 * g + 4], s[8 * g + 5]), w3 = pack_bf16x2(s[8 * g + 6], s[8 * g + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                       : "memory");
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      // fold O_j into the running output
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float o[32];
        __syncwarp();
        tmem_ld_32x32(t_lane + 64 + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, o[i]);
      }
    }
    if (row < Lq) {
      const float inv = 1.f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)seq * Lq + row) * (size_t)(H * ATT_D) + head * ATT_D);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        w.x = pack_bf16x2(o_acc[8 * g + 0] * inv, o_acc[8 * g + 1] * inv);
        w.y = pack_bf16x2(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
        w.z = pack_bf16x2(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
        w.w = pack_bf16x2(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
        dst[g] = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

int attn_launch(const AttnArgs& a, cudaStream_t st) {
  VB_REQUIRE(a.q && a.k && a.v && a.out, "attn: null pointer");
  VB_REQUIRE(a.n_seq > 0 && a.H > 0 && a.Lq > 0 && a.Lmax > 0, "attn: bad shape n_seq=%d H=%d Lq=%d Lmax=%d", a.n_seq,
             a.H, a.Lq, a.Lmax);
  VB_REQUIRE(a.n_scales > 0 && a.n_scales <= VB_MAX_SCALES, "attn: n_scales=%d out of range", a.n_scales);
  VB_REQUIRE(a.q_pos0 >= 0 && a.q_pos0 + a.Lq <= a.level_end[a.n_scales - 1], "attn: queries [%d,%d) exceed sequence %d",
             a.q_pos0, a.q_pos0 + a.Lq, a.level_end[a.n_scales - 1]);
  VB_REQUIRE(a.level_end[a.n_scales - 1] <= a.Lmax, "attn: sequence %d exceeds cache rows %d", a.level_end[a.n_scales - 1],
             a.Lmax);
  VB_REQUIRE(a.n_seq <= 65535 && a.H <= 65535, "attn: grid too large");
  AttnLevels lv;
  lv.n = a.n_scales;
  for (int i = 0; i < VB_MAX_SCALES; ++i) lv.end[i] = i < a.n_scales ? a.level_end[i] : a.level_end[a.n_scales - 1];
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t nbh = (uint64_t)a.n_seq * a.H;
  {
    uint64_t dims[3] = {64, (uint64_t)a.Lq, nbh};
    uint64_t str[2] = {128, (uint64_t)a.Lq * 128};
    uint32_t box[3] = {64, ATT_BM, 1};
    int r = make_tmap_bf16_sw128(&tmQ, a.q, 3, dims, str, box);
    if (r) return r;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)a.Lmax, nbh};
    uint64_t str[2] = {128, (uint64_t)a.Lmax * 128};
    uint32_t box[3] = {64, ATT_BN, 1};
    int r = make_tmap_bf16_sw128(&tmK, a.k, 3, dims, str, box);
    if (r) return r;
    r = make_tmap_bf16_sw128(&tmV, a.v, 3, dims, str, box);
    if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VB_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid((a.Lq + ATT_BM - 1) / ATT_BM, a.H, a.n_seq);
  attn_kernel<<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, reinterpret_cast<__nv_bfloat16*>(a.out), a.Lq, a.H,
                                                   a.q_pos0, lv);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
