// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

void set_error(const char* fmt, ...);
const char* get_error();

// Encodes a tiled bf16 tensor map with 128-byte swizzle. dims/box are innermost-first.
// strides_bytes[i] is the byte stride of dimension i+1 (rank-1 entries). Returns 0 or VB_ERR_DRIVER.
int make_tmap_bf16_sw128(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int sm_count();
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

}  // namespace vb
