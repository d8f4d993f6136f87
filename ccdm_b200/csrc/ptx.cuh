// sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM / commit).
// Hand-written for this project; encodings follow the PTX ISA for sm_100a.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

// Precision tiers.  The library is built twice from the same sources: libccdm_b200.so stores activations and packed weights
// as bfloat16 (8-bit mantissa); libccdm_b200_f16.so (-DCCDM_ACT_F16) stores them as IEEE binary16 -- the 10-bit mantissa of
// TF32 -- and feeds them to the same tcgen05.mma kind::f16 with fp32 accumulation.  Only the 2-byte element format changes:
// layouts, TMA boxes, shared-memory plans and schedules are identical.  Everything below the macro layer is written against
// the bf16 names.
#ifdef CCDM_ACT_F16
#include <cuda_fp16.h>
#define __nv_bfloat16 __half
#define __float2bfloat16 __float2half_rn
#define __bfloat162float __half2float
#define CCDM_ONE_PAIR 0x3C003C00u               // two binary16 1.0
#define CCDM_MMA_AB_FORMAT 0u                   // tcgen05 instruction descriptor: A / B format F16
#define CCDM_KEXP_SHIFT 10.0f                   // see kexp_bound_kernel (linattn.cu)
#else
#define CCDM_ONE_PAIR 0x3F803F80u               // two bfloat16 1.0
#define CCDM_MMA_AB_FORMAT 1u                   // A / B format BF16
#define CCDM_KEXP_SHIFT 0.0f
#endif

namespace ccdm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (sticky error, process exits) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("ccdm: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

// Address-based variants for the single-lane pipelines (no generic->shared conversion, out-of-line slow path): the
// TMA-producer and MMA-issuer warps are latency-bound, so their loops must be as short as possible.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("ccdm: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  if (!ok) mbar_wait_slow(bar_addr, parity);
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(const CUtensorMap* m, uint32_t bar_addr, uint32_t dst_addr, int c0, int c1,
                                              int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// One lane of a converged warp (warp-uniform control flow around it keeps operands in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until every committed bulk store has finished READING shared memory (the buffer may be overwritten).
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Executed by one full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ----------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a 2-CTA cluster (one TPC) execute ONE tcgen05.mma of M = 256: each CTA supplies its own 128 A rows and
// HALF of the B rows from the same shared-memory offsets, and receives its own 128 accumulator rows in its own TMEM.
// The leader (cluster rank 0) issues; TMA loads of both CTAs signal the LEADER's mbarrier (bit 24 of a shared::cluster
// address is the pair rank); commits are multicast to the barriers of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint32_t bar_addr, uint32_t dst_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];"
      ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t bar_addr, uint32_t dst_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have completed.
__device__ __forceinline__ void umma_commit_2sm_a(uint32_t bar_addr) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar_addr), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// mbarrier arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand in the canonical SWIZZLE_128B layout
// (rows of 128 B, 8-row / 1024 B swizzle atoms, atoms stacked along M/N):
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused for SW128 K-major; 1)
//   bits [32,46) stride byte offset >> 4 (1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 (sm_100)     bits [61,64) layout type = 2 (SWIZZLE_128B)
constexpr uint64_t kUmmaDescSw128 = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                    (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
// `addr16` = shared address >> 4 (fits the 14-bit field for any address below 256 KiB), so advancing a descriptor
// by k bytes is a plain integer add of k/16.
__device__ __forceinline__ uint64_t umma_desc_sw128_a16(uint32_t addr16) { return kUmmaDescSw128 | addr16; }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return umma_desc_sw128_a16((smem_addr & 0x3FFFF) >> 4);
}

// Instruction descriptor for kind::f16: A/B = bf16 (format 1), D = fp32 (format 1), both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (CCDM_MMA_AB_FORMAT << 7) | (CCDM_MMA_AB_FORMAT << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// MN-major SWIZZLE_128B operand: LBO = stride between 64-element MN blocks, SBO = stride between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (CCDM_MMA_AB_FORMAT << 7) | (CCDM_MMA_AB_FORMAT << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor drains; it must not touch the predecessor's outputs before griddep_wait().
// LDGSTS: 16 bytes global -> shared without a register round trip (zero-filled when !valid); groups are per thread
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------- misc
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// 8 consecutive floats; vector loads when the address allows (scale/shift offsets are multiples of 8 floats in
// practice), scalar otherwise
__device__ __forceinline__ void load8(const float* p, float (&o)[8]) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __ldg(p + j);
  }
}

// 2^x, one MUFU (exp2f without -use_fast_math expands to a range-handling sequence of ~6 instructions)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh of the SiLU epilogues (x*sigmoid(x) = h + h*tanh(h), h = x/2).  bf16 build: tanh.approx (one MUFU, relative error
// 2^-11 -- below bfloat16's rounding).  binary16 build: that error would be as large as the storage rounding itself, so
// tanh(h) = 1 - 2/(1 + e^{2h}) from ex2.approx and a fast division (~1e-7).
__device__ __forceinline__ float tanh_silu(float h) {
#ifdef CCDM_ACT_F16
  return 1.f - __fdividef(2.f, ex2_fast(h * 2.8853900817779268f) + 1.f);
#else
  return tanh_fast(h);
#endif
}
// sigmoid(x) = 0.5 + 0.5 tanh(x/2): one MUFU, no division
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Sum over the G lanes of a row segment, broadcast to all of them.  Power-of-two G: xor butterfly.  Branch-free: five
// fixed shuffle steps whose contribution is masked (a loop over the run-time G compiled to a divergent-capable loop with
// BRA.DIV per step and showed up in the row kernels' issue budget).
__device__ __forceinline__ float seg_sum(float v, int gl, int G, int lane) {
  if ((G & (G - 1)) == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, off);
      v += off < G ? o : 0.f;
    }
    return v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float o = __shfl_down_sync(0xffffffffu, v, off);
    v += gl + off < G ? o : 0.f;
  }
  return __shfl_sync(0xffffffffu, v, lane - gl);
}

#ifdef CCDM_ACT_F16
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFFu))); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16))); }
#else
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
#endif

}  // namespace ccdm
