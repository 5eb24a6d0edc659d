// Attention for the first KV-cached scales of the pyramid (models/basic_var.py:98-117 with l <= 32 queries and at most
// 64 visible keys per (sequence, head): scales 0-4 of the 256 px pyramid, l = 1..25, keys = 1..55).
//
// The tcgen05 kernel (attn_sm100.cu) works on 128-query x 64-key tiles: here a work item would fill 1-20 % of the query
// rows of one tile, and its cost is the latency of the TMA -> UMMA -> softmax -> UMMA -> epilogue chain per item (about
// 4000 cycles per item per CTA whatever l is: profiles/r02_attn_scale_times.txt, 108-112 us per launch for 15 360 items).
// A whole item fits one warp: Q (<= 32 x 64), K and V (<= 64 x 64) go to the warp's private shared memory with plain
// 16-byte loads, S = Q K^T and O = P V are warp-level mma.sync m16n8k16 tiles (bf16 in, fp32 accumulate), the softmax
// runs on the accumulator fragments (row reductions inside a quad). No block-level barrier, no pipeline to fill.
// Same function as the large kernel: scores masked by the level limit of every query row, P rounded to bf16, fp32 row
// sums of the unrounded exponentials, O / l stored as bf16.
#include "attn.h"
#include "common.cuh"
#include "host.h"

namespace vb {

constexpr int AS_WARPS = 8;    // work items in flight per CTA (one per warp)
constexpr int AS_LD = 72;      // shared-memory row pitch in bf16 elements (144 B: ldmatrix rows hit distinct banks)
constexpr int AS_QROWS = 32, AS_KROWS = 64;
// Shared memory per warp is sized by the call: 16-row groups of Q, K and V that hold live rows ((16 + 2*16) x 144 B =
// 6.9 KB at scale 0, (32 + 2*64) x 144 B = 23 KB at scale 4), so the small scales keep up to 32 warps per SM resident.

struct AttnSmallLevels {
  int end[VB_MAX_SCALES];
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(AS_WARPS * 32)
attn_small_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                  __nv_bfloat16* __restrict__ out, int Lq, int H, int Lmax, int q_pos0, const __grid_constant__ AttnSmallLevels lv,
                  int n_items, int kv_max, float exp2_scale) {
  extern __shared__ __align__(16) uint8_t as_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (Lq + 15) >> 4;          // 16-row query tiles with live rows
  const int n_nt = (kv_max + 7) >> 3;       // 8-key tiles with visible keys
  const int n_kt = (kv_max + 15) >> 4;      // 16-key steps of the P V product
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(as_smem) + (size_t)warp * ((n_mt + 2 * n_kt) * 16 * AS_LD);
  __nv_bfloat16* sk = sq + n_mt * 16 * AS_LD;
  __nv_bfloat16* sv = sk + n_kt * 16 * AS_LD;
  const uint32_t sq_a = smem_u32(sq), sk_a = smem_u32(sk), sv_a = smem_u32(sv);
  pdl_wait();               // q / K / V of the QKV GEMM are complete and visible
  pdl_launch_dependents();

  auto kv_end_of = [&](int row) {  // visible keys of query row `row` (lv.end padded with the sequence length)
    const int pos = q_pos0 + row;
    int e = lv.end[VB_MAX_SCALES - 1];
#pragma unroll
    for (int s = VB_MAX_SCALES - 2; s >= 0; --s) e = (pos < lv.end[s]) ? lv.end[s] : e;
    return e;
  };
  const int g = lane >> 2, t4 = lane & 3;   // fragment row inside the 8-row group, column pair

  for (int item = blockIdx.x * AS_WARPS + warp; item < n_items; item += gridDim.x * AS_WARPS) {
    const int head = item % H, seq = item / H;
    const uint4* gq = reinterpret_cast<const uint4*>(q + (size_t)item * Lq * 64);
    const uint4* gk = reinterpret_cast<const uint4*>(k + (size_t)item * Lmax * 64);
    const uint4* gv = reinterpret_cast<const uint4*>(v + (size_t)item * Lmax * 64);
    __syncwarp();  // the previous item's ldmatrix reads are done before the tiles are overwritten
    // rows as eight 16-byte chunks, copied asynchronously (cp.async: every chunk of the item is in flight at once, one
    // memory round trip per item); rows beyond the live ones are zero (V: 0 * stale data must not make NaNs)
    for (int i = lane; i < n_mt * 16 * 8; i += 32) {
      const int r = i >> 3, c = i & 7;
      const uint32_t dst = sq_a + (uint32_t)((r * AS_LD + c * 8) * 2);
      if (r < Lq) cp_async_16(dst, gq + i);
      else *reinterpret_cast<uint4*>(sq + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
    }
    for (int i = lane; i < n_kt * 16 * 8; i += 32) {
      const int r = i >> 3, c = i & 7;
      const uint32_t off = (uint32_t)((r * AS_LD + c * 8) * 2);
      if (r < kv_max) {
        cp_async_16(sk_a + off, gk + i);
        cp_async_16(sv_a + off, gv + i);
      } else {
        *reinterpret_cast<uint4*>(sk + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sv + r * AS_LD + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
    cp_async_wait_all();
    __syncwarp();

    for (int mt = 0; mt < n_mt; ++mt) {
      // ---- S = Q K^T for 16 query rows x (8 * n_nt) keys
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { s[nt][0] = 0.f; s[nt][1] = 0.f; s[nt][2] = 0.f; s[nt][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(sq_a + (uint32_t)(((mt * 16 + (lane & 15)) * AS_LD + kk * 16 + (lane >> 4) * 8) * 2), a0, a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (nt < n_nt) {
            uint32_t b0, b1;
            ldsm_x2(sk_a + (uint32_t)(((nt * 8 + (lane & 7)) * AS_LD + kk * 16 + ((lane >> 3) & 1) * 8) * 2), b0, b1);
            mma_bf16_16816(s[nt], a0, a1, a2, a3, b0, b1);
          }
        }
      }
      // ---- softmax over the visible keys of rows r0 = mt*16 + g and r1 = r0 + 8 (row data lives in one quad)
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const int lim0 = kv_end_of(r0 < Lq ? r0 : Lq - 1), lim1 = kv_end_of(r1 < Lq ? r1 : Lq - 1);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t4;
        s[nt][0] = (c < lim0) ? s[nt][0] : -INFINITY;
        s[nt][1] = (c + 1 < lim0) ? s[nt][1] : -INFINITY;
        s[nt][2] = (c < lim1) ? s[nt][2] : -INFINITY;
        s[nt][3] = (c + 1 < lim1) ? s[nt][3] : -INFINITY;
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float o0 = m0 * exp2_scale, o1 = m1 * exp2_scale;  // key 0 is visible to every row: the maxima are finite
      float l0 = 0.f, l1 = 0.f;
      uint32_t p[8][2];  // bf16x2: rows r0 / r1, columns c, c+1
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float e0 = exp2f(fmaf(s[nt][0], exp2_scale, -o0)), e1 = exp2f(fmaf(s[nt][1], exp2_scale, -o0));
        const float e2 = exp2f(fmaf(s[nt][2], exp2_scale, -o1)), e3 = exp2f(fmaf(s[nt][3], exp2_scale, -o1));
        l0 += e0 + e1;
        l1 += e2 + e3;
        p[nt][0] = pack_bf16x2(e0, e1);
        p[nt][1] = pack_bf16x2(e2, e3);
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // ---- O = P V: the accumulator fragments of S are the A fragments of the second product
      float o[8][4];
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) { o[nd][0] = 0.f; o[nd][1] = 0.f; o[nd][2] = 0.f; o[nd][3] = 0.f; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < n_kt) {
          const uint32_t a0 = p[2 * j][0], a1 = p[2 * j][1], a2 = p[2 * j + 1][0], a3 = p[2 * j + 1][1];
#pragma unroll
          for (int nd = 0; nd < 8; ++nd) {
            uint32_t b0, b1;
            ldsm_x2_trans(sv_a + (uint32_t)(((j * 16 + (lane & 15)) * AS_LD + nd * 8) * 2), b0, b1);
            mma_bf16_16816(o[nd], a0, a1, a2, a3, b0, b1);
          }
        }
      }
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      __nv_bfloat16* out0 = out + ((size_t)seq * Lq + r0) * (size_t)(H * 64) + head * 64 + 2 * t4;
      __nv_bfloat16* out1 = out + ((size_t)seq * Lq + r1) * (size_t)(H * 64) + head * 64 + 2 * t4;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        if (r0 < Lq) *reinterpret_cast<uint32_t*>(out0 + nd * 8) = pack_bf16x2(o[nd][0] * i0, o[nd][1] * i0);
        if (r1 < Lq) *reinterpret_cast<uint32_t*>(out1 + nd * 8) = pack_bf16x2(o[nd][2] * i1, o[nd][3] * i1);
      }
    }
  }
}

bool attn_small_applies(const AttnArgs& a, int kv_vis) { return a.Lq <= AS_QROWS && kv_vis <= AS_KROWS; }

int attn_small_launch(const AttnArgs& a, int kv_vis, cudaStream_t st) {
  AttnSmallLevels lv;
  for (int i = 0; i < VB_MAX_SCALES; ++i) lv.end[i] = i < a.n_scales ? a.level_end[i] : a.level_end[a.n_scales - 1];
  const long long n_items = (long long)a.n_seq * a.H;
  VB_REQUIRE(n_items < (1ll << 31), "attn: too many work items");
  const int rows = ((a.Lq + 15) / 16 + 2 * ((kv_vis + 15) / 16)) * 16;
  const size_t smem = (size_t)AS_WARPS * rows * AS_LD * 2;
  static SmemAttrCache attr_cache;
  if (ensure_dyn_smem(attr_cache, (size_t)AS_WARPS * (AS_QROWS + 2 * AS_KROWS) * AS_LD * 2, attn_small_kernel)) return VB_ERR_CUDA;
  int grid = (int)((n_items + AS_WARPS - 1) / AS_WARPS);
  int per_sm = (int)((size_t)220 * 1024 / smem);  // CTAs of this size one SM holds (shared memory; 2048 threads)
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int cap = sm_count() * per_sm;    // persistent: every resident warp walks its share of the items
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(AS_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  // scores are base-2 exponents already when q carries log2(e) (q_log2), natural-log units otherwise
  const float exp2_scale = a.q_log2 ? 1.f : 1.4426950408889634f;
  vb::ProfScope prof_scope(vb::PK_ATTN, st);
  VB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_small_kernel, reinterpret_cast<const __nv_bfloat16*>(a.q),
                                   reinterpret_cast<const __nv_bfloat16*>(a.k), reinterpret_cast<const __nv_bfloat16*>(a.v),
                                   reinterpret_cast<__nv_bfloat16*>(a.out), a.Lq, a.H, a.Lmax, a.q_pos0, lv, (int)n_items, kv_vis,
                                   exp2_scale));
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}

}  // namespace vb
