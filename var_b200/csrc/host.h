// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

namespace vb {

void set_error(const char* fmt, ...);
const char* get_error();

// Encodes a tiled bf16 tensor map with 128-byte swizzle. dims/box are innermost-first.
// strides_bytes[i] is the byte stride of dimension i+1 (rank-1 entries). elem_strides: NULL (all 1) or the traversal
// stride per dimension: the box then covers box[i] elements of the tensor and delivers ceil(box[i] / elem_strides[i])
// of them (strided convolutions). Returns 0 or VB_ERR_DRIVER.
int make_tmap_bf16_sw128(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr);

// SM count of the CURRENT device (cached per device).
int sm_count();
// Current device ordinal (cudaGetDevice), -1 on error.
int current_device();
// number of kernels this library has launched in the process (bench.py's gpu_launches claim)
void count_launch(int n = 1);

// Optional per-kernel CUDA-event timing (var_b200_profile_begin / _end): every launcher opens a ProfScope on its
// stream; when profiling is off this is two predictable branches.
enum ProfKind : int {
  PK_GEMM_BIAS_F32 = 0, PK_GEMM_BIAS_BF16, PK_GEMM_GELU, PK_GEMM_GATE_RESID, PK_GEMM_QKV, PK_GEMM_SCORE,
  PK_ATTN, PK_LN, PK_EMBED, PK_COND, PK_SAMPLE, PK_QUANT, PK_SCORE_FIN, PK_OTHER, PK_COUNT
};
struct ProfScope {
  int kind;
  cudaStream_t st;
  void* rec;
  ProfScope(int kind, cudaStream_t st);
  ~ProfScope();
};

#define VB_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      vb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return vb::VB_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define VB_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      vb::set_error(__VA_ARGS__);    \
      return vb::VB_ERR_ARG;         \
    }                                \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: this cache remembers, per device,
// the largest size already granted to the kernels of one launcher, and serialises the (rare) slow path, so the
// launchers are safe with several GPUs in one process and with several host threads.
// VAR_B200_PDL=0 launches the GEMM / attention chain without programmatic dependent launch (A/B runs).
bool pdl_enabled();

constexpr int VB_MAX_DEVICES = 64;
struct SmemAttrCache {
  std::atomic<int> granted[VB_MAX_DEVICES];
  std::mutex mu;
  SmemAttrCache() { for (auto& g : granted) g.store(-1, std::memory_order_relaxed); }
};
template <class... Kerns>
inline int ensure_dyn_smem(SmemAttrCache& cache, size_t bytes, Kerns... kerns) {
  const int dev = current_device();
  const bool cached = dev >= 0 && dev < VB_MAX_DEVICES;
  if (cached && (int)bytes <= cache.granted[dev].load(std::memory_order_acquire)) return 0;
  std::lock_guard<std::mutex> lk(cache.mu);
  if (cached && (int)bytes <= cache.granted[dev].load(std::memory_order_acquire)) return 0;
  const cudaError_t errs[] = {cudaFuncSetAttribute(kerns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)...};
  for (cudaError_t e : errs)
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu) failed: %s", bytes, cudaGetErrorString(e));
      return -2;  // VB_ERR_CUDA (common.cuh)
    }
  if (cached) cache.granted[dev].store((int)bytes, std::memory_order_release);
  return 0;
}

}  // namespace vb
