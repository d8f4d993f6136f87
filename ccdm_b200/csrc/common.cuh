// Error plumbing and launch bookkeeping shared by every translation unit of libccdm_b200.so.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/ccdm_b200.h"

namespace ccdm {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return CCDM_ERR_CUDA;
}

// Checks the launch that was just issued (graph-capture safe: no sync).
inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return CCDM_OK;
}

#define CCDM_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::ccdm::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

}  // namespace ccdm
