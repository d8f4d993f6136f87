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

// Row kernels (RMSNorm forward / backward) give each row of nchunk 16-byte vectors to G lanes holding kv vectors each,
// so a warp covers 32/G rows per iteration.  Pick the kv <= max_kv that keeps the most lanes busy.
inline void row_lane_plan(int nchunk, int max_kv, int* kv_out, int* g_out) {
  // smallest kv (fewest registers, most resident warps) that keeps >= 80 % of the lanes busy; else the best one
  int best_kv = 0, best_g = 0;
  double best = -1.0;
  for (int kv = 1; kv <= max_kv; ++kv) {
    if (nchunk > 32 * kv) continue;
    const int g = (nchunk + kv - 1) / kv;
    const double score = (double)(32 / g) * nchunk / (32.0 * kv);
    if (score >= 0.8) {
      best_kv = kv;
      best_g = g;
      break;
    }
    if (score > best + 1e-9) {
      best = score;
      best_kv = kv;
      best_g = g;
    }
  }
  *kv_out = best_kv;
  *g_out = best_g;
}

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