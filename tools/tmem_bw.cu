// Measurement aid (not product code): per-SM throughput of tcgen05.ld (TMEM read), tcgen05.st and MUFU.EX2 as a
// function of the number of warps, to decide what bounds the attention softmax warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ivar_b200/csrc tools/tmem_bw.cu -o tools/tmem_bw
#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace vb;

// mode 0: 2 x LDTM.x32 + wait per iteration (8 KB per warp-iteration)
// mode 1: STTM.x32 + wait per iteration (4 KB)
// mode 2: 48 MUFU.EX2 per iteration (one attention tile's worth at poly 1/4)
// mode 3: mode 0 + mode 2 interleaved (independent)
__global__ void __launch_bounds__(512, 1) bw_kernel(int mode, int iters, unsigned long long* clk, float* sink) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_smem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = 0.f;
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = (float)(threadIdx.x + i) * 1e-3f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0 || mode == 3) {
      tmem_ld_32x32(taddr, s);
      tmem_ld_32x32(taddr + 32, s + 32);
      tmem_ld_wait_dep(s);
      tmem_ld_wait_dep(s + 32);
      acc += s[0] + s[63];
    }
    if (mode == 1) {
      tmem_st_32x32(taddr, s);
      tmem_st_wait();
    }
    if (mode == 2 || mode == 3) {
      float e[48];
#pragma unroll
      for (int i = 0; i < 48; ++i) e[i] = fast_exp2(s[i] * 1e-3f - (float)it);
#pragma unroll
      for (int i = 0; i < 48; ++i) acc += e[i];
    }
  }
  const unsigned long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base_smem, 256);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* clk;
  float* sink;
  cudaMalloc(&clk, sizeof(unsigned long long) * 4096 * 16);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  const char* names[] = {"LDTM 2x32 cols (8 KB/warp-iter)", "STTM 32 cols (4 KB/warp-iter)", "48 MUFU.EX2/warp-iter",
                         "LDTM 2x32 + 48 MUFU"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
      for (int warps : {1, 4, 8, 16}) {
        if (warps * ctas_per_sm > 32) continue;
        const int grid = sms * ctas_per_sm, threads = warps * 32;
        // 2 CTAs/SM need both resident: 256 TMEM columns each, small smem -> they are
        bw_kernel<<<grid, threads>>>(mode, 10, clk, sink);
        cudaDeviceSynchronize();
        bw_kernel<<<grid, threads>>>(mode, iters, clk, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<unsigned long long> h(grid * warps);
        cudaMemcpy(h.data(), clk, h.size() * 8, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (auto v : h) mean += (double)v;
        mean /= h.size();
        const double per_iter = mean / iters;                      // clk per warp-iteration (all warps concurrent)
        const double per_sm = per_iter / (warps * ctas_per_sm);    // SM clk per warp-iteration of throughput
        printf("%-34s ctas/SM=%d warps/CTA=%2d : %8.1f clk per iteration per warp, %7.1f clk per warp-iteration per SM\n",
               names[mode], ctas_per_sm, warps, per_iter, per_sm);
      }
    }
  }
  return 0;
}
