// TEST INFRASTRUCTURE: a minimal CUDA-on-host shim so that the CUDA-core kernels of a .cu file can be compiled by g++ and
// executed on the CPU exactly as written (index math, shared-memory staging, barriers, shuffles, atomics), one thread
// block at a time with one OS thread per CUDA thread.  Used by tests/test_kernels_hostsim.py to check kernels in a
// container without a GPU; it is NOT a fallback and nothing under ccdm_b200/ uses it.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "ccdm_b200.h"

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int4 { int x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
static const int cudaSuccess = 0;

static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static std::unique_ptr<std::barrier<>> g_bar;
static std::vector<double> g_shfl;      // one 8-byte slot per thread (float and double shuffles)

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static              // blocks run one at a time, so one static copy per kernel == per-block storage

static inline void __syncthreads() { g_bar->arrive_and_wait(); }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
// warp-scoped: the 32 (or fewer, in a trailing partial warp) threads of a warp must all take part, as on the device
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_bar;
template <typename T>
static inline T shfl_from(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  const unsigned t = threadIdx.x, w = t >> 5;
  memcpy(&g_shfl[t], &v, sizeof(T));
  g_warp_bar[w]->arrive_and_wait();
  unsigned src = (t & ~31u) | (src_lane & 31u);
  if (src >= blockDim.x) src = t;
  T r;
  memcpy(&r, &g_shfl[src], sizeof(T));
  g_warp_bar[w]->arrive_and_wait();
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return shfl_from(v, (threadIdx.x ^ (unsigned)lane_mask) & 31u); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const unsigned lane = threadIdx.x & 31u;
  return shfl_from(v, lane + delta > 31u ? lane : lane + delta);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src_lane) { return shfl_from(v, (unsigned)src_lane); }
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(x) expf(x)
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::max;
using std::min;

struct __nv_bfloat16 { uint16_t v; };
static inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = (uint32_t)h.v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat16 __float2bfloat16(float f) {          // round to nearest even, as the device intrinsic
  uint32_t u;
  memcpy(&u, &f, 4);
  __nv_bfloat16 h;
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) { h.v = 0x7FFF; return h; }
  u += 0x7FFFu + ((u >> 16) & 1u);
  h.v = (uint16_t)(u >> 16);
  return h;
}
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static float g_dyn_smem[64 * 1024];                                  // 256 KB of "dynamic shared memory"
static const int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 4; return 0; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline int atomicMax(int* p, int v) {
  std::atomic_ref<int> a(*p);
  int old = a.load();
  while (old < v && !a.compare_exchange_weak(old, v)) {}
  return old;
}
static inline unsigned atomicMin(unsigned* p, unsigned v) {
  std::atomic_ref<unsigned> a(*p);
  unsigned old = a.load();
  while (old > v && !a.compare_exchange_weak(old, v)) {}
  return old;
}
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }

// ---- what the kernels use from common.cuh / ptx.cuh
namespace ccdm {
static char g_err[512];
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
static inline int cuda_fail(cudaError_t, const char* what) { set_error("%s", what); return CCDM_ERR_CUDA; }
static inline int after_launch(const char*) { return CCDM_OK; }
static inline int num_sms() { return 148; }
static inline float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
static inline float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
// host equivalents of the small helpers in ptx.cuh (same results up to the approximate-instruction error)
static inline uint32_t pack_bf16(float lo, float hi) { return (uint32_t)__float2bfloat16(lo).v | ((uint32_t)__float2bfloat16(hi).v << 16); }
static inline void load8(const float* p, float (&o)[8]) { for (int j = 0; j < 8; ++j) o[j] = p[j]; }
static inline float silu_f(float v) { return v / (1.0f + expf(-v)); }
static inline float ex2_fast(float x) { return exp2f(x); }
static inline float tanh_fast(float x) { return tanhf(x); }
static inline float sigmoid_fast(float x) { return fmaf(0.5f, tanhf(0.5f * x), 0.5f); }
static inline float seg_sum(float v, int gl, int G, int lane) {
  if ((G & (G - 1)) == 0) {
    for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  for (int off = 16; off > 0; off >>= 1) {
    const float o = __shfl_down_sync(0xffffffffu, v, off);
    if (gl + off < G) v += o;
  }
  return __shfl_sync(0xffffffffu, v, lane - gl);
}
static inline void griddep_wait() {}
static inline void griddep_launch_dependents() {}
}  // namespace ccdm
#define CCDM_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::ccdm::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

static inline void launch_blocks(dim3 grid, dim3 block, const std::function<void()>& body) {
  gridDim = grid;
  blockDim = block;
  g_shfl.assign(block.x, 0.0);
  const unsigned n_warps = (block.x + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block.x);
        g_warp_bar.clear();
        for (unsigned w = 0; w < n_warps; ++w)
          g_warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, block.x - 32 * w)));
        std::vector<std::thread> ts;
        ts.reserve(block.x);
        for (unsigned t = 0; t < block.x; ++t)
          ts.emplace_back([&, t] {
            threadIdx = dim3(t, 0, 0);
            blockIdx = dim3(bx, by, bz);
            body();
            g_bar->arrive_and_drop();          // a thread that returned early must not block later barriers
            g_warp_bar[t >> 5]->arrive_and_drop();
          });
        for (auto& th : ts) th.join();
      }
}
// kernels without barriers / shuffles: the CUDA threads of every block run one after the other in the calling thread
static inline void launch_blocks_seq(dim3 grid, dim3 block, const std::function<void()>& body) {
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned t = 0; t < block.x; ++t) {
          threadIdx = dim3(t, 0, 0);
          blockIdx = dim3(bx, by, bz);
          body();
        }
}
#define LAUNCH_SEQ(kernel, grid, block, ...) launch_blocks_seq(dim3 grid, dim3 block, [&] { kernel(__VA_ARGS__); })
#define LAUNCH(kernel, grid, block, ...) launch_blocks(dim3 grid, dim3 block, [&] { kernel(__VA_ARGS__); })

extern "C" const char* hostsim_last_error() { return ccdm::g_err; }
