// Host-side helpers: last-error record, TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda), device properties.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/vp_b200.h"
#include "common.cuh"

namespace vp {

struct ErrorState {
  int cuda_error = 0;
  char msg[256] = {0};
};
inline ErrorState& err_state() {
  static thread_local ErrorState s;
  return s;
}
inline int fail(int code, const char* what, int cuda_error = 0) {
  ErrorState& e = err_state();
  e.cuda_error = cuda_error;
  snprintf(e.msg, sizeof(e.msg), "%s", what);
  return code;
}
#define VP_CHECK_CUDA(expr)                                                       \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) return vp::fail(VP_ERR_CUDA, #expr, (int)_e);      \
  } while (0)
#define VP_REQUIRE(cond, code, what)              \
  do {                                            \
    if (!(cond)) return vp::fail((code), (what)); \
  } while (0)

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled encode_fn() {
  static PFN_tensorMapEncodeTiled fn = []() -> PFN_tensorMapEncodeTiled {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_tensorMapEncodeTiled>(f);
  }();
  return fn;
}

// bf16 tensor, innermost dimension contiguous, 128-byte swizzle, zero fill out of bounds.
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
inline int encode_tmap_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_tensorMapEncodeTiled fn = encode_fn();
  if (!fn) return fail(VP_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(VP_ERR_BAD_ALIGN, "TMA base pointer not 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) return fail(VP_ERR_BAD_ALIGN, "TMA stride not a multiple of 16 bytes");
    }
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VP_ERR_DRIVER, "cuTensorMapEncodeTiled failed", (int)r);
  return VP_OK;
}

// A denoise step launches ~330 kernels with 2-5 tensor maps each, and every step of a run uses the same buffers (the
// workspace and the weights are allocated once): the 128-byte descriptors are encoded once per (pointer, shape, box) and
// looked up afterwards.  A CUtensorMap holds nothing but the address and the geometry, so a cached copy stays valid for
// as long as a buffer of that geometry lives at that address; the table is simply dropped when it grows past a bound.
struct TmapKey {
  uint64_t w[12];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.w) h = (h ^ x) * 1099511628211ull;
    return (size_t)h;
  }
};
inline int make_tmap_bf16(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey k{};
  k.w[0] = reinterpret_cast<uint64_t>(ptr);
  k.w[1] = (uint64_t)rank;
  for (int i = 0; i < rank && i < 4; ++i) {
    k.w[2 + i] = dims[i];
    k.w[6 + i] = i > 0 ? strides_bytes[i - 1] : 0;
    k.w[10] |= (uint64_t)box[i] << (16 * i);
  }
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(k);
    if (it != cache.end()) {
      *map = it->second;
      return VP_OK;
    }
  }
  const int rc = encode_tmap_bf16(map, ptr, rank, dims, strides_bytes, box);
  if (rc != VP_OK) return rc;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(k, *map);
  return VP_OK;
}

// Per-device caches: the SM count and cudaFuncAttributeMaxDynamicSharedMemorySize are properties of (function, device), and a
// process may drive several GPUs (one thread-local runtime per virtual rank: parallel.py).
constexpr int kMaxDevices = 64;

inline int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  v = cache[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

// Opt a kernel into `smem_bytes` of dynamic shared memory once per (kernel, device).
inline int configure_once(const void* func, int smem_bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, unsigned long long> done;
  int dev = 0;
  VP_CHECK_CUDA(cudaGetDevice(&dev));
  VP_REQUIRE(dev >= 0 && dev < kMaxDevices, VP_ERR_UNSUPPORTED, "device ordinal out of range");
  std::lock_guard<std::mutex> g(mu);
  unsigned long long& mask = done[func];
  if ((mask >> dev) & 1ull) return VP_OK;
  VP_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  mask |= 1ull << dev;
  return VP_OK;
}

}  // namespace vp
