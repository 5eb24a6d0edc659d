// Single-tile UMMA probe: D[128,N] = A[128,64] * B, with B either K-major ([N,64]) or MN-major ([64,N], the
// natural layout of V in attention). Used by tests to pin the shared-memory descriptor encodings on hardware.
#include "common.cuh"
#include "host.h"
#include "../../include/var_b200.h"

namespace vb {

// MN-major operand, 128B swizzle: rows are K indices, each row holds 64 contiguous MN elements (128 bytes);
// 8-row groups are 1024 bytes apart (SBO); LBO = distance between successive 64-element MN chunks.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int N,
                  int b_mn_major) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t b_bytes = b_mn_major ? 64 * 128 * (N / 64) : N * 128;
    mbar_expect_tx(smem_u32(&bars[0]), 16384 + b_bytes);
    tma_load_2d(&tmA, smem_u32(&bars[0]), sA, 0, 0);
    if (b_mn_major) {
      for (int c = 0; c < N / 64; ++c) tma_load_2d(&tmB, smem_u32(&bars[0]), sB + c * 8192, c * 64, 0);
    } else {
      tma_load_2d(&tmB, smem_u32(&bars[0]), sB, 0, 0);
    }
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    uint32_t idesc = umma_idesc_bf16(128, N);
    if (b_mn_major) idesc |= (1u << 16);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = umma_desc_k_sw128(sA) + 2 * k;
      const uint64_t bd = b_mn_major ? umma_desc_mn_sw128(sB + k * 2048, 8192) : umma_desc_k_sw128(sB) + 2 * k;
      umma_bf16_ss(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(smem_u32(&bars[1]));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N / 32; ++c) {
    float v[32];
    __syncwarp();
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait_dep(v);
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace vb

// A: [128,64] bf16. B: [N,64] bf16 (b_mn_major=0) or [64,N] bf16 (b_mn_major=1). D: [128,N] fp32. N in {64,128,256}.
extern "C" int var_b200_umma_probe(const void* A, const void* B, float* D, int N, int b_mn_major, void* stream) {
  using namespace vb;
  VB_REQUIRE(A && B && D, "probe: null pointer");
  VB_REQUIRE(N == 64 || N == 128 || N == 256, "probe: N=%d unsupported", N);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {64, 128};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 128};
    int r = make_tmap_bf16_sw128(&tmA, A, 2, dims, str, box);
    if (r) return r;
  }
  if (b_mn_major) {
    uint64_t dims[2] = {(uint64_t)N, 64};
    uint64_t str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 64};
    int r = make_tmap_bf16_sw128(&tmB, B, 2, dims, str, box);
    if (r) return r;
  } else {
    uint64_t dims[2] = {64, (uint64_t)N};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, (uint32_t)N};
    int r = make_tmap_bf16_sw128(&tmB, B, 2, dims, str, box);
    if (r) return r;
  }
  const int smem = 16384 + 32768 + 1024;
  VB_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, D, N, b_mn_major);
  VB_CUDA_CHECK(cudaGetLastError());
  vb::count_launch();
  return VB_OK;
}
