#include <stdlib.h>
// Host-side runtime helpers: last-error string, driver entry point for cuTensorMapEncodeTiled
// (resolved at run time so the library loads on machines without libcuda), SM count.
#include "host.h"
#include "../../include/var_b200.h"

#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace vb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Descriptor cache: the block driver re-encodes the same (pointer, shape, box) maps on every call (weights and
// workspace buffers keep their addresses), and cuTensorMapEncodeTiled costs microseconds of host time each.
struct TmapKey {
  uint64_t w[21];  // ptr | rank | dims[5] | box[5] | strides[4] | element strides[5]
};
static std::unordered_map<std::string, CUtensorMap> g_tmaps;
static std::mutex g_tmaps_mu;

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VAR_B200_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

int make_tmap_bf16_sw128(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  TmapKey key{};
  key.w[0] = reinterpret_cast<uint64_t>(gptr);
  key.w[1] = (uint64_t)rank;
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = box[i];
    if (i + 1 < rank) key.w[12 + i] = strides_bytes[i];
    key.w[16 + i] = elem_strides ? elem_strides[i] : 1;
  }
  const std::string ks(reinterpret_cast<const char*>(&key), sizeof(key));
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    auto it = g_tmaps.find(ks);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return VB_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return VB_ERR_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  // VAR_B200_TMAP_L2PROMO = 0 / 64 / 128 / 256 (default 256): L2 promotion size of every tensor map (A/B runs)
  static const CUtensorMapL2promotion promo = [] {
    const char* e = getenv("VAR_B200_TMAP_L2PROMO");
    const int v = e ? atoi(e) : 256;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }();
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult=%d rank=%d dims=(%llu,%llu,%llu) box=(%u,%u,%u) ptr=%p", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, gptr);
    return VB_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmaps_mu);
    if (g_tmaps.size() > 65536) g_tmaps.clear();
    g_tmaps.emplace(ks, *out);
  }
  return VB_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------- per-kernel event timing
struct ProfRec {
  int kind;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec*> g_prof;
static std::mutex g_prof_mu;

ProfScope::ProfScope(int k, cudaStream_t s) : kind(k), st(s), rec(nullptr) {
  if (!g_prof_on) return;
  ProfRec* r = new ProfRec{k, nullptr, nullptr};
  if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
  cudaEventRecord(r->a, s);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  ProfRec* r = static_cast<ProfRec*>(rec);
  cudaEventRecord(r->b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}

int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static std::atomic<int> cache[VB_MAX_DEVICES];  // zero-initialised: 0 = not queried yet
  const int dev = current_device();
  if (dev < 0) return 148;
  if (dev < VB_MAX_DEVICES) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev < VB_MAX_DEVICES) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

}  // namespace vb

extern "C" const char* var_b200_last_error(void) { return vb::get_error(); }
extern "C" long long var_b200_launch_count(void) { return vb::g_launches.load(); }

extern "C" void var_b200_profile_begin(void) {
  std::lock_guard<std::mutex> lk(vb::g_prof_mu);
  for (auto* r : vb::g_prof) { cudaEventDestroy(r->a); cudaEventDestroy(r->b); delete r; }
  vb::g_prof.clear();
  vb::g_prof_on = true;
}

extern "C" int var_b200_profile_end(double* ms_by_kind, long long* n_by_kind, int n_kinds) {
  vb::g_prof_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return vb::VB_ERR_CUDA;
  std::lock_guard<std::mutex> lk(vb::g_prof_mu);
  for (int i = 0; i < n_kinds; ++i) { ms_by_kind[i] = 0.0; n_by_kind[i] = 0; }
  for (auto* r : vb::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess && r->kind < n_kinds) {
      ms_by_kind[r->kind] += ms;
      n_by_kind[r->kind] += 1;
    }
    cudaEventDestroy(r->a); cudaEventDestroy(r->b); delete r;
  }
  vb::g_prof.clear();
  return vb::VB_OK;
}
