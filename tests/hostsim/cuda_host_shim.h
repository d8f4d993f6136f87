// TEST INFRASTRUCTURE: a minimal CUDA-on-host shim so that the CUDA-core kernels of a .cu file can be compiled by g++ and
// executed on the CPU exactly as written (index math, shared-memory staging, barriers, shuffles, atomics), one thread
// block at a time.  Every CUDA thread of the block is a cooperative fiber (ucontext) that runs until its next barrier or
// shuffle, so execution is deterministic and needs no OS threads; kernels without barriers run their threads in a plain
// loop.  Used by tests/test_kernels_hostsim*.py and tests/hostpath.py to check kernels in a container without a GPU; it is
// NOT a fallback and nothing under ccdm_b200/ uses it.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <setjmp.h>
#include <ucontext.h>
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

static dim3 threadIdx, blockIdx, blockDim, gridDim;
static std::vector<double> g_shfl;      // one 8-byte slot per thread (float and double shuffles)

// ---- cooperative fibers: one per CUDA thread of the running block
enum { WAIT_NONE = 0, WAIT_BLOCK = 1, WAIT_WARP = 2 };
// makecontext / setcontext start a fiber on its own stack; every later switch is _setjmp / _longjmp (no signal-mask system
// call, ~50x cheaper than swapcontext).  Built with _FORTIFY_SOURCE off: the fortified longjmp rejects cross-stack jumps.
struct Fiber {
  ucontext_t ctx;
  jmp_buf jb;
  std::vector<char> stack;
  bool started = false, done = false;
  int wait = WAIT_NONE;
};
static std::vector<Fiber> g_fib;
static jmp_buf g_sched_jb;
static unsigned g_cur = 0;
static const std::function<void()>* g_body = nullptr;
static inline void fiber_yield() {               // back to the scheduler; returns when the scheduler resumes this fiber
  if (!_setjmp(g_fib[g_cur].jb)) _longjmp(g_sched_jb, 1);
}
static inline void fiber_wait(int scope) {       // park the current fiber at a barrier of the given scope
  g_fib[g_cur].wait = scope;
  fiber_yield();
}
static inline void fiber_resume(unsigned t) {    // scheduler side: run fiber t until it yields or finishes
  if (_setjmp(g_sched_jb)) return;
  Fiber& f = g_fib[t];
  if (f.started) _longjmp(f.jb, 1);
  f.started = true;
  setcontext(&f.ctx);
}

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static              // blocks run one at a time, so one static copy per kernel == per-block storage

static inline void __syncthreads() { fiber_wait(WAIT_BLOCK); }
static inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
// warp-scoped: the 32 (or fewer, in a trailing partial warp) threads of a warp must all take part, as on the device
template <typename T>
static inline T shfl_from(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  const unsigned t = threadIdx.x;
  memcpy(&g_shfl[t], &v, sizeof(T));
  fiber_wait(WAIT_WARP);
  unsigned src = (t & ~31u) | (src_lane & 31u);
  if (src >= blockDim.x) src = t;
  T r;
  memcpy(&r, &g_shfl[src], sizeof(T));
  fiber_wait(WAIT_WARP);
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return shfl_from(v, (threadIdx.x ^ (unsigned)lane_mask) & 31u); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const unsigned lane = threadIdx.x & 31u;
  return shfl_from(v, lane + delta > 31u ? lane : lane + delta);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src_lane) { return shfl_from(v, (unsigned)src_lane); }
static inline void __syncwarp(unsigned = 0xffffffffu) { fiber_wait(WAIT_WARP); }
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
static inline int atomicMax(int* p, int v) { const int o = *p; if (o < v) *p = v; return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { const unsigned o = *p; if (o > v) *p = v; return o; }
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
#define CCDM_KEXP_SHIFT 0.0f
#define CCDM_ONE_PAIR 0x3F803F80u
static inline float tanh_silu(float x) { return tanhf(x); }
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
static inline void cp_async16(void* dst, const void* src, bool valid) { if (valid) memcpy(dst, src, 16); else memset(dst, 0, 16); }
static inline void cp_async_commit() {}
template <int kPending> static inline void cp_async_wait() {}
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

static void fiber_entry() {
  (*g_body)();
  g_fib[g_cur].done = true;
  _longjmp(g_sched_jb, 1);
}
// Round-robin scheduler: every runnable fiber runs until its next barrier / shuffle (or to the end); a barrier opens when all
// live fibers of its scope (block, or one warp) are parked at it.  A block that can neither run nor open a barrier is a
// deadlock in the kernel (divergent barrier) and aborts.
static inline void launch_blocks(dim3 grid, dim3 block, const std::function<void()>& body) {
  constexpr size_t kStack = 256 * 1024;
  gridDim = grid;
  blockDim = block;
  g_body = &body;
  g_shfl.assign(block.x, 0.0);
  if (g_fib.size() < block.x) g_fib.resize(block.x);
  const unsigned n = block.x, n_warps = (n + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz);
        for (unsigned t = 0; t < n; ++t) {
          Fiber& f = g_fib[t];
          if (f.stack.size() != kStack) f.stack.resize(kStack);
          f.started = f.done = false;
          f.wait = WAIT_NONE;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.data();
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
        }
        unsigned live = n;
        while (live) {
          bool progressed = false;
          for (unsigned t = 0; t < n; ++t) {
            Fiber& f = g_fib[t];
            if (f.done || f.wait != WAIT_NONE) continue;
            g_cur = t;
            threadIdx = dim3(t, 0, 0);
            fiber_resume(t);
            progressed = true;
            if (f.done) --live;
          }
          // open the barriers whose every live participant has arrived
          bool all_block = live > 0;
          for (unsigned t = 0; t < n && all_block; ++t)
            if (!g_fib[t].done && g_fib[t].wait != WAIT_BLOCK) all_block = false;
          if (all_block) {
            for (unsigned t = 0; t < n; ++t) g_fib[t].wait = WAIT_NONE;
            progressed = true;
          }
          for (unsigned w = 0; w < n_warps; ++w) {
            const unsigned lo = 32 * w, hi = std::min(n, lo + 32);
            bool all = false, any = false;
            for (unsigned t = lo; t < hi; ++t)
              if (!g_fib[t].done) {
                if (!any) all = true;
                any = true;
                if (g_fib[t].wait != WAIT_WARP) all = false;
              }
            if (any && all) {
              for (unsigned t = lo; t < hi; ++t) g_fib[t].wait = WAIT_NONE;
              progressed = true;
            }
          }
          if (!progressed) {
            fprintf(stderr, "hostsim: deadlock in block (%u,%u,%u): divergent barrier\n", bx, by, bz);
            abort();
          }
        }
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
// HOSTSIM_PROFILE=1: per-kernel launch counts and wall time, printed when the library is unloaded
#include <chrono>
#include <map>
#include <string>
struct HostsimProfile {
  std::map<std::string, std::pair<long, double>> rows;
  ~HostsimProfile() {
    if (!getenv("HOSTSIM_PROFILE")) return;
    for (auto& r : rows) fprintf(stderr, "hostsim %-44s %6ld launches %9.3f s\n", r.first.c_str(), r.second.first, r.second.second);
  }
};
static HostsimProfile g_profile;
template <typename F>
static inline void profiled(const char* name, F&& f) {
  const auto t0 = std::chrono::steady_clock::now();
  f();
  auto& r = g_profile.rows[name];
  r.first += 1;
  r.second += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
#define LAUNCH_SEQ(kernel, grid, block, ...) \
  profiled(#kernel, [&] { launch_blocks_seq(dim3 grid, dim3 block, [&] { kernel(__VA_ARGS__); }); })
#define LAUNCH(kernel, grid, block, ...) profiled(#kernel, [&] { launch_blocks(dim3 grid, dim3 block, [&] { kernel(__VA_ARGS__); }); })

extern "C" const char* hostsim_last_error() { return ccdm::g_err; }
