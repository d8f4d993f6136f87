// Tap-GEMM: persistent implicit-GEMM convolution / linear engine for sm_100a (tcgen05 + TMEM + TMA).
//
//   D[128 pixels x n_tile] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem via TMA)  .  B[n_tile x 64]^T (bf16, smem via TMA)
//
// One persistent CTA per SM walks a static list of 128-row output tiles (tb x th x tw boxes of output positions) for
// one (sub-problem z, channel tile n0) combination = blockIdx.y.
//   warp 0 / lane 0  TMA producer.  A K "load group" is one 4-D box {64 ch, tw, th+R-1, tb} of a source view read at
//                    the tap-shifted position; out-of-range coordinates are zero-filled by TMA (= conv zero padding).
//                    The box feeds R vertically adjacent filter taps: tap r starts r*tw rows (a multiple of the
//                    1024-byte swizzle atom) further down the same box, so those rows are fetched once, not R times.
//                    Weights are loaded ONCE per CTA and stay resident in shared memory when they fit; otherwise the
//                    R weight blocks of a group travel in the same pipeline stage as its box.
//   warp 1 / lane 0  tcgen05.mma issuer (M=128, N<=256 per instruction, K=16; two N halves for n_tile > 256).
//                    Accumulators live in TMEM, double-buffered (2 x n_tile columns) so tile i+1 is multiplied while
//                    tile i is in the epilogue.  tcgen05.commit frees smem stages / publishes finished accumulators.
//   warps 2..9       epilogue.  tcgen05.ld 32x32b hands every thread one pixel row: channel RMSNorm, scale/shift,
//                    SiLU, q-softmax, residual add and the bf16 pack are per-thread loops without shuffles.  The two
//                    warps that share a TMEM lane quarter split the output columns; the norm's sum of squares is
//                    recomputed by both from TMEM (cheaper than exchanging it).
//
// Replaces the nn.Conv2d / nn.Linear call sites listed in include/ccdm_b200.h.
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr int kMaxN = 512;

struct TapGemmMaps {
  CUtensorMap a[CCDM_MAX_SRC];
  CUtensorMap b;
  CUtensorMap o[CCDM_MAX_Z];              // bf16 output views (one per sub-problem) for the TMA-store epilogue
  CUtensorMap r[CCDM_MAX_SRC];            // the sources again with a box of th rows: residual groups (one tap, no halo rows)
};

struct TapGemmDev {
  int gW, gH, gB, tw, th, tb, tiles_w, tiles_h, tiles_m, n_tiles;
  int ngroups, R, nkb, n_rows, N, n_tile, n_sub, nsub, stages, acc_stages, w_batch_rows, b_resident;
  int n_inner;                            // channel tiles walked by the SAME CTA per pixel tile (1: one per blockIdx.y)
  uint32_t a_bytes, stage_bytes, res_bytes, out_bytes;
  uint32_t out_pitch;                     // dense-row output staging: bytes per pixel row (0: swizzled 64-channel panels)
  int tiles_per_cta, out_bufs, store_tma;
  const int4* sched;
  const int* ksteps;                      // [nz * ngroups] K steps (16 channels each) that hold data per load group, or nullptr
  uint32_t flags, tmem_cols;
  const float *bias, *rowss, *gain, *ss;
  int ss_ld, ss_off;
  const __nv_bfloat16* resid;
  long long rsW, rsH, rsB;
  void* out;
  long long osW, osH, osB;
  long long ooff[CCDM_MAX_Z];
  float* out_rowss;
  float q_scale, gain_mul;
  int q_cols;
  int epi_alt;                            // the two epilogue warp groups take alternate tiles (n_tile <= 64)
  int pair;                               // CTA pairs: one tcgen05.mma.cta_group::2 (M = 256) covers a tile of each CTA
  int halo;                               // 3x3 taps out of one halo box per 64-channel block (R == 9)
  uint32_t stage_tx;                      // bytes TMA delivers into one main stage (the stage stride is rounded to 1 KiB)
  int n_res;                              // CCDM_EPI_RESACC: 1x1 load groups after the main ones, accumulated in a SECOND
  uint32_t a_bytes_res;                   //   TMEM accumulator (res_conv / identity shortcut); their box bytes
  const float* res_bias;
  int head_n;                             // CCDM_EPI_HEAD: fused 1x1 head to head_n <= 4 fp32 NCHW planes
  const float *head_w, *head_b;
  float* head_out;
  long long hsC, hsB;
};

// aux shared-memory block (after the resident weights and the stage ring)
struct __align__(16) TapGemmAux {
  uint64_t a_full[kMaxStages], a_empty[kMaxStages], b_full, tmem_full[2], tmem_empty[2];
  uint32_t tmem_slot, pad_[3];
  float bias[kMaxN], gain[kMaxN];         // bias[n], g[n]*gain_mul (0 for padded channels)
  float gs[2][kMaxN], sh[2][kMaxN];       // per-tile g*(1+scale[b]), shift[b] (tiles inside one sample)
  float part[2][kTileM];                  // sum-of-squares exchange between the two column halves (tile parity)
  float bias2[128];                       // CCDM_EPI_RESACC: res_conv bias
  float headw[4][128];                    // CCDM_EPI_HEAD: head weights (0 for k >= head_n, n >= N)
  float hpart[2][4][kTileM];              // ... and the exchange of its partial dot products (split mode)
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
// alternate-tile mode: each group of four epilogue warps synchronises on its own named barrier (ids 1 and 2)
__device__ __forceinline__ void epi_bar_g(bool alt, int grp) {
  if (alt) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
  else epi_bar();
}

__device__ __forceinline__ float fast_silu(float v) { return __fdividef(v, 1.f + __expf(-v)); }

// Descriptor from a low word that already holds (address >> 4) | LBO bit; the high word is constant.
__device__ __forceinline__ uint64_t umma_desc_lo(uint32_t lo) {
  return (kUmmaDescSw128 & 0xFFFFFFFF00000000ull) | lo;
}

// Halo boxes (3x3, one box {64 ch, tw + 2, th + 2} per 64-channel block, tw == 8): tap (r, q) starts (r * (tw + 2) + q) rows
// = 128-byte units into the box -- NOT a multiple of the 8-row swizzle atom -- and the 8-row groups of the operand (one image
// row each) are (tw + 2) rows = 1280 bytes apart.  SWIZZLE_128B is a pure function of the shared-memory address for TMA and
// tcgen05.mma alike, so plain descriptors with that start address and SBO read the shifted windows correctly (base offset
// field 0; tools/ubench/shift_desc.cu is the experiment, profiles/r2_ubench_shift_desc.txt its output).
__device__ __forceinline__ uint64_t umma_desc_lo_halo(uint32_t lo) {
  constexpr uint64_t hi = (static_cast<uint64_t>(1280 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
  return hi | lo;
}

// kR vertically adjacent taps x 4 K-steps of one load group, fully unrolled (tap r: r*tw rows further down the box).
template <int kR, bool kPair = false>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t tap16, uint32_t b16,
                                           uint32_t idesc, bool first_group, int nk = 4) {
  if constexpr (kR == 9) {                                 // halo box: 3 x 3 taps, 10-row image pitch (tw == 8)
    // One filter row (12 MMAs) per loop iteration: fully unrolled, the 72 descriptor words of a group overflow the uniform
    // register file and every R2UR.FILL of a spilled one stalls the issuing lane (26 % of its samples in ncu).
    uint32_t b_t = b_lo;
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
      uint32_t a_t = a_lo + static_cast<uint32_t>(r) * 80u;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (r | q | k) != 0 ? 1u : (first_group ? 0u : 1u);
          if (kPair)
            umma_bf16_ss_2sm(d_tmem, umma_desc_lo_halo(a_t + 2 * k), umma_desc_lo(b_t + 2 * k), idesc, acc);
          else
            umma_bf16_ss(d_tmem, umma_desc_lo_halo(a_t + 2 * k), umma_desc_lo(b_t + 2 * k), idesc, acc);
        }
        a_t += 8u;
        b_t += b16;
      }
    }
    return;
  }
  if (nk != 4) {
    // A load group whose 64-channel block is only partly real (72 = 64 + 8 channels: the dim-72 models) issues the K steps
    // that hold data and nothing else -- one warp-uniform branch per GROUP (a predicate per instruction slowed the issuing
    // lane by 10 %): 45 instead of 72 instructions per tile for a 72 -> 72 3x3 layer.
#pragma unroll
    for (int r = 0; r < kR; ++r) {
#pragma unroll 1
      for (int k = 0; k < nk; ++k) {
        if (kPair)
          umma_bf16_ss_2sm(d_tmem, umma_desc_lo(a_lo + r * tap16 + 2 * k), umma_desc_lo(b_lo + r * b16 + 2 * k), idesc,
                           (r | k) != 0 ? 1u : (first_group ? 0u : 1u));
        else
          umma_bf16_ss(d_tmem, umma_desc_lo(a_lo + r * tap16 + 2 * k), umma_desc_lo(b_lo + r * b16 + 2 * k), idesc,
                       (r | k) != 0 ? 1u : (first_group ? 0u : 1u));
      }
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < kR; ++r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (kPair)
        umma_bf16_ss_2sm(d_tmem, umma_desc_lo(a_lo + r * tap16 + 2 * k), umma_desc_lo(b_lo + r * b16 + 2 * k), idesc,
                         (r | k) != 0 ? 1u : (first_group ? 0u : 1u));
      else
        umma_bf16_ss(d_tmem, umma_desc_lo(a_lo + r * tap16 + 2 * k), umma_desc_lo(b_lo + r * b16 + 2 * k), idesc,
                     (r | k) != 0 ? 1u : (first_group ? 0u : 1u));
    }
  }
}
template <bool kPair>
__device__ __forceinline__ void commit_bar(uint32_t bar_addr) {
  if (kPair) umma_commit_2sm_a(bar_addr);
  else umma_commit_a(bar_addr);
}

// MMA-issuer loop of the common case (weights resident in shared memory, one N sub-tile, one channel tile per CTA),
// specialised on the number of vertically adjacent taps: no per-group branching, barrier addresses as plain integers.
template <int kR, bool kPair>
__device__ __forceinline__ void mma_loop_resident(int t_begin, int t_end, int n_groups, int n_stages, uint32_t a_lo0,
                                                  uint32_t stage16, uint32_t b_lo0, uint32_t tap16, uint32_t b16,
                                                  uint32_t idesc, uint32_t tmem_base, uint32_t n_tile, int acc_mask,
                                                  int acc_shift, uint32_t full_bar, uint32_t empty_bar,
                                                  uint32_t tfull_bar, uint32_t tempty_bar, int n_res, uint32_t res_off,
                                                  const int* s_nk) {
  int s = 0;
  uint32_t ph = 0;
  uint32_t a_lo = a_lo0;
  uint32_t item = 0;
  for (int tile = t_begin; tile < t_end; ++tile, ++item) {
    const int as = item & acc_mask;
    mbar_wait_a(tempty_bar + 8 * as, ((item >> acc_shift) & 1) ^ 1u);    // epilogue has drained this accumulator stage
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + as * n_tile;
    uint32_t b_lo = b_lo0;
    for (int g = 0; g < n_groups; ++g) {
      const int nk = s_nk ? s_nk[g] : 4;                   // (read before the wait: its latency hides behind the barrier)
      mbar_wait_a(full_bar + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        issue_taps<kR, kPair>(d_tmem, a_lo, b_lo, tap16, b16, idesc, g == 0, nk);
        commit_bar<kPair>(empty_bar + 8 * s);
      }
      __syncwarp();
      b_lo += kR * b16;
      a_lo += stage16;
      if (++s == n_stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
    }
    for (int g = 0; g < n_res; ++g) {                      // shortcut (1x1) groups -> the second accumulator
      mbar_wait_a(full_bar + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        issue_taps<1, kPair>(d_tmem + res_off, a_lo, b_lo, tap16, b16, idesc, g == 0);
        commit_bar<kPair>(empty_bar + 8 * s);
      }
      __syncwarp();
      b_lo += b16;
      a_lo += stage16;
      if (++s == n_stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
    }
    if (elect_one()) commit_bar<kPair>(tfull_bar + 8 * as);
    __syncwarp();
  }
}

// MMA-issuer loop (one N sub-tile), specialised on the number of vertically adjacent taps and on whether the weights
// are resident in shared memory or ride in the stage behind the A box: no per-group branching, barrier addresses as
// plain integers.  n_inner > 1: every CTA walks several channel tiles of the same boxes (qkv), which stay in the ring
// until the last channel tile has read them.
template <int kR, bool kRes, bool kMultiN, bool kPair = false>
__device__ __forceinline__ void mma_loop_fast(int t_begin, int t_end, int n_groups, int n_stages, int n_inner_rt,
                                              uint32_t a_lo0, uint32_t stage16, uint32_t abytes16, uint32_t b_lo0,
                                              uint32_t nkb_b16, uint32_t tap16, uint32_t b16, uint32_t idesc,
                                              uint32_t tmem_base, uint32_t n_tile, int acc_mask, int acc_shift,
                                              uint32_t full_bar, uint32_t empty_bar, uint32_t tfull_bar,
                                              uint32_t tempty_bar, int n_res, uint32_t res_off, const int* s_nk) {
  const int n_inner = kMultiN ? n_inner_rt : 1;             // compile-time 1 for the common case: no restore code
  int s = 0;
  uint32_t ph = 0;
  uint32_t a_lo = a_lo0;
  uint32_t item = 0;
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int s_tile = s;
    const uint32_t ph_tile = ph, a_tile = a_lo;
    uint32_t b_nt = b_lo0;
    for (int nt = 0; nt < n_inner; ++nt, ++item, b_nt += nkb_b16) {
      const int as = item & acc_mask;
      mbar_wait_a(tempty_bar + 8 * as, ((item >> acc_shift) & 1) ^ 1u);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * n_tile;
      if (kMultiN) {
        s = s_tile;
        ph = ph_tile;
        a_lo = a_tile;
      }
      uint32_t b_lo = b_nt;
      const bool last_nt = !kMultiN || nt == n_inner - 1;
      for (int g = 0; g < n_groups; ++g) {
        const int nk = s_nk ? s_nk[g] : 4;
        if (!kMultiN || nt == 0) {
          mbar_wait_a(full_bar + 8 * s, ph);
          tc_fence_after();
        }
        if (elect_one()) {
          issue_taps<kR, kPair>(d_tmem, a_lo, kRes ? b_lo : a_lo + abytes16, tap16, b16, idesc, g == 0, nk);
          if (last_nt) commit_bar<kPair>(empty_bar + 8 * s);   // box reusable once the last MMAs have read it
        }
        __syncwarp();
        b_lo += kR * b16;
        a_lo += stage16;
        if (++s == n_stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
      }
      if (!kMultiN) {
        for (int g = 0; g < n_res; ++g) {                  // shortcut (1x1) groups -> the second accumulator
          mbar_wait_a(full_bar + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            issue_taps<1, kPair>(d_tmem + res_off, a_lo, kRes ? b_lo : a_lo + abytes16, tap16, b16, idesc, g == 0);
            commit_bar<kPair>(empty_bar + 8 * s);
          }
          __syncwarp();
          b_lo += b16;
          a_lo += stage16;
          if (++s == n_stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
        }
      }
      if (elect_one()) commit_bar<kPair>(tfull_bar + 8 * as);
      __syncwarp();
    }
  }
}

constexpr uint32_t kRuntimeFlags = 0x80000000u;            // template value: epilogue flags are read from the params

// kFlags: the CCDM_EPI_* set compiled into the epilogue (dead branches vanish), or kRuntimeFlags.
// kStoreTma: bf16 output through the shared-memory staging tile + TMA bulk store.
// kPair: CTA-pair build (tcgen05 cta_group::2; must be launched as 2-CTA clusters -- a kernel containing cta_group::2
// instructions cannot be launched without them, hence a separate instantiation).
template <uint32_t kFlags, bool kStoreTma, bool kPair>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const __grid_constant__ TapGemmMaps maps,
                                                              const TapGemmDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];        // 1024-byte aligned: SWIZZLE_128B atoms
  const uint32_t kflags = (kFlags == kRuntimeFlags) ? p.flags : kFlags;
  uint8_t* res_b = smem;                                   // resident weights (may be empty)
  uint8_t* ring = smem + p.res_bytes;                      // pipeline stages
  uint8_t* stg = ring + static_cast<size_t>(p.stages) * p.stage_bytes;   // output staging (TMA-store epilogue)
  TapGemmAux* aux = reinterpret_cast<TapGemmAux*>(stg + static_cast<size_t>(p.out_bufs) * p.out_bytes);
  int4* s_sched = reinterpret_cast<int4*>(aux + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int z = blockIdx.y / p.n_tiles;
  const int n_base = (blockIdx.y % p.n_tiles) * p.n_tile;       // first output channel of this CTA
  // CTA pairs: this CTA keeps (and loads) only its half of the weight rows of every K block
  constexpr bool pair = kPair;
  uint32_t crank = 0;
  if constexpr (kPair) crank = cluster_ctarank();
  const int b_rows = pair ? p.n_tile >> 1 : p.n_tile;
  const int b_bytes = b_rows * 128;
  const int w_row0 = n_base + static_cast<int>(crank) * b_rows;  // first weight row this CTA loads (within z)

  // ---------------------------------------------------------------- one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.a[0]);
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_2sm(&aux->tmem_slot, p.tmem_cols);
    else tmem_alloc(&aux->tmem_slot, p.tmem_cols);
  }
  if (tid == 64) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&aux->a_full[s], 1);
      mbar_init(&aux->a_empty[s], 1);
    }
    mbar_init(&aux->b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aux->tmem_full[s], 1);
      // (pair: the epilogue warps of BOTH CTAs release the accumulator stage on the leader's barrier)
      mbar_init(&aux->tmem_empty[s], (p.epi_alt ? kEpiWarps / 2 : kEpiWarps) * (pair ? 2 : 1));
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.n_tile * p.n_inner; i += kThreads) {
    const bool ok = (n_base + i) < p.N;
    aux->bias[i] = ((kflags & CCDM_EPI_BIAS) && ok) ? p.bias[n_base + i] : 0.f;
    aux->gain[i] = ((kflags & CCDM_EPI_RMSNORM) && ok) ? p.gain[n_base + i] * p.gain_mul : 0.f;
  }
  if (kflags & CCDM_EPI_HEAD) {
    for (int i = tid; i < 4 * 128; i += kThreads) {
      const int k = i >> 7, n = i & 127;
      aux->headw[k][n] = (k < p.head_n && n < p.N) ? p.head_w[k * p.N + n] : 0.f;
    }
  }
  if (kflags & CCDM_EPI_RESACC) {
    for (int i = tid; i < 128; i += kThreads)
      aux->bias2[i] = (p.res_bias && (n_base + i) < p.N) ? p.res_bias[n_base + i] : 0.f;
  }
  // (residual groups follow the nz * ngroups main entries; they exist only with nz == 1)
  for (int i = tid; i < p.ngroups + p.n_res; i += kThreads) s_sched[i] = p.sched[z * p.ngroups + i];
  int* const s_nk_tab = reinterpret_cast<int*>(s_sched + p.ngroups + p.n_res);
  if (p.ksteps)
    for (int i = tid; i < p.ngroups; i += kThreads) s_nk_tab[i] = min(4, max(1, p.ksteps[z * p.ngroups + i]));
  const int* const s_nk = p.ksteps ? s_nk_tab : nullptr;
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();                 // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;
  const int tiles_per_sample = p.tiles_w * p.tiles_h;
  // Launched with programmatic stream serialization: everything above (and the resident-weight loads below, which
  // no kernel of the forward pass writes) overlaps the previous kernel's drain.  The next kernel may start its own
  // prologue as soon as every CTA of this grid has passed this point.
  griddep_launch_dependents();
  // contiguous tile range per CTA: neighbouring tiles share halos in L2 and (mostly) the sample's scale/shift
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  // (pair: both CTAs walk the same NUMBER of tiles -- tiles past the end are dummies: TMA zero-fills, nothing is stored)
  const int t_end = pair ? t_begin + p.tiles_per_cta : min(p.tiles_m, t_begin + p.tiles_per_cta);

  if (warp == 0) {
    // ============================================================== TMA producer (warp-uniform loops, one lane issues)
    if (p.w_batch_rows != 0) griddep_wait();               // per-sample weights come from the previous kernel
    if (p.b_resident && elect_one()) {
      if constexpr (kPair) {                                // both halves complete on the leader's barrier
        if (crank == 0) mbar_arrive_expect_tx(&aux->b_full, 2u * static_cast<uint32_t>(p.nkb) * b_bytes);
        for (int kb = 0; kb < p.nkb; ++kb)
          tma_load_2d_2sm(&maps.b, smem_u32(&aux->b_full), smem_u32(res_b + static_cast<size_t>(kb) * b_bytes),
                          kb * kBlockK, z * p.n_rows + w_row0);
      } else {
      mbar_arrive_expect_tx(&aux->b_full, static_cast<uint32_t>(p.nkb * p.n_inner) * b_bytes);
      for (int nt = 0; nt < p.n_inner; ++nt)
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int sub = 0; sub < p.nsub; ++sub)
            tma_load_2d(&maps.b, &aux->b_full,
                        res_b + static_cast<size_t>(nt * p.nkb + kb) * b_bytes + sub * p.n_sub * 128, kb * kBlockK,
                        z * p.n_rows + n_base + nt * p.n_tile + sub * p.n_sub);
      }
    }
    __syncwarp();
    griddep_wait();                                        // activations below are the previous kernels' outputs
    // No divisions in the loop: the single producer lane is latency-bound, every instruction counts.
    int twi = t_begin % p.tiles_w, thi = (t_begin / p.tiles_w) % p.tiles_h, tbi = t_begin / tiles_per_sample;
    const int n_stages = p.stages, n_groups = p.ngroups;
    const uint32_t stage_bytes = p.stage_bytes;
    const bool b_res = p.b_resident != 0;
    int s = 0;                                             // ring slot of the next load group
    uint32_t ph = 0;                                       // ... and how many times the ring has wrapped (mod 2)
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int w0 = twi * p.tw, h0 = thi * p.th, b0 = tbi * p.tb;
      if (++twi == p.tiles_w) { twi = 0; if (++thi == p.tiles_h) { thi = 0; ++tbi; } }
      const int wrow = b0 * p.w_batch_rows + z * p.n_rows + n_base;
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(&aux->a_empty[s], ph ^ 1u);
        if constexpr (kPair) {
          // Both CTAs fill their own slot s; every byte is accounted on the LEADER's barrier, which its MMA lane waits on.
          if (elect_one()) {
            if (crank == 0) mbar_arrive_expect_tx(&aux->a_full[s], 2u * p.stage_tx);
            const int4 e = s_sched[g];
            const uint32_t st = smem_u32(ring + static_cast<size_t>(s) * stage_bytes);
            const uint32_t fb = smem_u32(&aux->a_full[s]);
            tma_load_4d_2sm(&maps.a[e.x], fb, st, e.w, w0 + e.y, h0 + e.z, b0);
            if (!b_res) {
              for (int r = 0; r < p.R; ++r)
                tma_load_2d_2sm(&maps.b, fb, st + p.a_bytes + r * b_bytes, (g * p.R + r) * kBlockK, z * p.n_rows + w_row0);
            }
          }
        } else if (elect_one()) {
          mbar_arrive_expect_tx(&aux->a_full[s], p.stage_tx);
          const int4 e = s_sched[g];
          uint8_t* st = ring + static_cast<size_t>(s) * stage_bytes;
          tma_load_4d(&maps.a[e.x], &aux->a_full[s], st, e.w, w0 + e.y, h0 + e.z, b0);
          if (!b_res) {
            for (int r = 0; r < p.R; ++r)
              for (int sub = 0; sub < p.nsub; ++sub)
                tma_load_2d(&maps.b, &aux->a_full[s], st + p.a_bytes + r * b_bytes + sub * p.n_sub * 128,
                            (g * p.R + r) * kBlockK, wrow + sub * p.n_sub);
          }
        }
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
      for (int g = 0; g < p.n_res; ++g) {                  // shortcut sources: one unshifted box of th rows per 64 channels
        mbar_wait(&aux->a_empty[s], ph ^ 1u);
        if (elect_one()) {
          const int4 e = s_sched[n_groups + g];
          const uint32_t st = smem_u32(ring + static_cast<size_t>(s) * stage_bytes);
          const uint32_t fb = smem_u32(&aux->a_full[s]);
          const uint32_t bytes = p.a_bytes_res + (b_res ? 0u : static_cast<uint32_t>(b_bytes));
          const int kcol = (n_groups * p.R + g) * kBlockK;
          if constexpr (kPair) {
            if (crank == 0) mbar_arrive_expect_tx(&aux->a_full[s], 2u * bytes);
            tma_load_4d_2sm(&maps.r[e.x], fb, st, e.w, w0, h0, b0);
            if (!b_res) tma_load_2d_2sm(&maps.b, fb, st + p.a_bytes, kcol, z * p.n_rows + w_row0);
          } else {
            mbar_arrive_expect_tx(&aux->a_full[s], bytes);
            tma_load_4d_a(&maps.r[e.x], fb, st, e.w, w0, h0, b0);
            if (!b_res)
              tma_load_2d(&maps.b, &aux->a_full[s], ring + static_cast<size_t>(s) * stage_bytes + p.a_bytes, kcol, wrow);
          }
        }
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ============================================================== MMA issuer (warp-uniform loops, one lane issues;
    // in a CTA pair only the leader's: its M = 256 instructions read both CTAs' shared memory and write both TMEMs)
    const uint32_t idesc = umma_idesc_bf16(pair ? 2 * kTileM : kTileM, p.n_sub);
    const uint32_t ring16 = (smem_u32(ring) & 0x3FFFF) >> 4;       // descriptor address fields, in 16-byte units
    const uint32_t res16 = (smem_u32(res_b) & 0x3FFFF) >> 4;
    const uint32_t stage16 = p.stage_bytes >> 4, abytes16 = p.a_bytes >> 4, b16 = static_cast<uint32_t>(b_bytes) >> 4;
    const uint32_t tap16 = static_cast<uint32_t>(p.tw) * 8;         // r*tw rows * 128 B / 16
    const uint32_t sub16 = static_cast<uint32_t>(p.n_sub) * 8;
    if (p.b_resident) {
      mbar_wait(&aux->b_full, 0);
      tc_fence_after();
    }
    // The issuing lane is latency-bound (a serial chain of uniform-datapath instructions), so the loop carries running
    // ring / phase counters instead of dividing, keeps the loop invariants in registers and unrolls the taps with
    // compile-time offsets.  Descriptor low words already contain the constant LBO bit: advancing is one add.
    const uint32_t desc_lo = static_cast<uint32_t>(kUmmaDescSw128);          // 1 << 16
    const uint32_t a_lo0 = ring16 | desc_lo, b_res_lo = res16 | desc_lo;
    const int n_stages = p.stages, n_groups = p.ngroups, n_inner = p.n_inner, R = p.R;
    const bool b_res = p.b_resident != 0, one_sub = p.nsub == 1;
    const uint32_t n_tile = p.n_tile, nkb_b16 = static_cast<uint32_t>(p.nkb) * b16, rb16 = static_cast<uint32_t>(R) * b16;
    const int acc_mask = p.acc_stages - 1, acc_shift = p.acc_stages >> 1;
    const int n_res = p.n_res;
    const uint32_t res_off = static_cast<uint32_t>(p.acc_stages) * n_tile;   // second accumulators sit behind the main ones
    if (one_sub && (R <= 3 || p.halo) && (n_inner == 1 || (R == 1 && b_res))) {
      const uint32_t full_bar = smem_u32(&aux->a_full[0]), empty_bar = smem_u32(&aux->a_empty[0]);
      const uint32_t tfull_bar = smem_u32(&aux->tmem_full[0]), tempty_bar = smem_u32(&aux->tmem_empty[0]);
#define CCDM_MMA_LOOP(KR, RES, MULTI)                                                                                  \
  mma_loop_fast<KR, RES, MULTI>(t_begin, t_end, n_groups, n_stages, n_inner, a_lo0, stage16, abytes16, b_res_lo, nkb_b16, \
                                tap16, b16, idesc, tmem_base, n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, \
                                tempty_bar, n_res, res_off, s_nk)
      if constexpr (kPair) {                               // host guarantees: one channel tile per CTA, R <= 3
#define CCDM_PAIR_RES(KR)                                                                                                 \
  mma_loop_resident<KR, true>(t_begin, t_end, n_groups, n_stages, a_lo0, stage16, b_res_lo, tap16, b16, idesc, tmem_base, \
                              n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk)
#define CCDM_PAIR_STR(KR)                                                                                                 \
  mma_loop_fast<KR, false, false, true>(t_begin, t_end, n_groups, n_stages, n_inner, a_lo0, stage16, abytes16, b_res_lo,  \
                                        nkb_b16, tap16, b16, idesc, tmem_base, n_tile, acc_mask, acc_shift, full_bar,     \
                                        empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk)
        if (b_res) {
          if (R == 9) CCDM_PAIR_RES(9);
          else if (R == 3) CCDM_PAIR_RES(3);
          else if (R == 2) CCDM_PAIR_RES(2);
          else CCDM_PAIR_RES(1);
        } else {
          if (R == 9) CCDM_PAIR_STR(9);
          else if (R == 3) CCDM_PAIR_STR(3);
          else if (R == 2) CCDM_PAIR_STR(2);
          else CCDM_PAIR_STR(1);
        }
#undef CCDM_PAIR_RES
#undef CCDM_PAIR_STR
      } else if (n_inner > 1) {                            // several channel tiles per box (qkv): weights always resident
        CCDM_MMA_LOOP(1, true, true);
      } else if (b_res) {
        if (R == 9)
          mma_loop_resident<9, false>(t_begin, t_end, n_groups, n_stages, a_lo0, stage16, b_res_lo, tap16, b16, idesc, tmem_base,
                               n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk);
        else if (R == 3)
          mma_loop_resident<3, false>(t_begin, t_end, n_groups, n_stages, a_lo0, stage16, b_res_lo, tap16, b16, idesc, tmem_base,
                               n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk);
        else if (R == 2)
          mma_loop_resident<2, false>(t_begin, t_end, n_groups, n_stages, a_lo0, stage16, b_res_lo, tap16, b16, idesc, tmem_base,
                               n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk);
        else
          mma_loop_resident<1, false>(t_begin, t_end, n_groups, n_stages, a_lo0, stage16, b_res_lo, tap16, b16, idesc, tmem_base,
                               n_tile, acc_mask, acc_shift, full_bar, empty_bar, tfull_bar, tempty_bar, n_res, res_off, s_nk);
      } else {
        if (R == 9) CCDM_MMA_LOOP(9, false, false);
        else if (R == 3) CCDM_MMA_LOOP(3, false, false);
        else if (R == 2) CCDM_MMA_LOOP(2, false, false);
        else CCDM_MMA_LOOP(1, false, false);
      }
#undef CCDM_MMA_LOOP
    } else {
    int s = 0;                                             // ring slot of the tile's next load group
    uint32_t ph = 0;
    uint32_t item = 0;                                     // (pixel tile, channel tile) pairs = accumulator uses
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int s_tile = s;
      const uint32_t ph_tile = ph;
      for (int nt = 0; nt < n_inner; ++nt, ++item) {
        const int as = item & acc_mask;
        const uint32_t aph = (item >> acc_shift) & 1;
        mbar_wait(&aux->tmem_empty[as], aph ^ 1u);         // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * n_tile;
        s = s_tile;                                        // the boxes of a tile stay resident for all its channel tiles
        ph = ph_tile;
        uint32_t b_lo = b_res_lo + nt * nkb_b16;
        for (int g = 0; g < n_groups; ++g) {
          if (nt == 0) {
            mbar_wait(&aux->a_full[s], ph);
            tc_fence_after();
          }
          const uint32_t a_lo = a_lo0 + s * stage16;
          if (!b_res) b_lo = a_lo + abytes16;
          if (elect_one()) {
            if (one_sub) {
              const int nk = s_nk ? s_nk[g] : 4;
              if (R == 3) issue_taps<3>(d_tmem, a_lo, b_lo, tap16, b16, idesc, g == 0, nk);
              else if (R == 1) issue_taps<1>(d_tmem, a_lo, b_lo, tap16, b16, idesc, g == 0, nk);
              else if (R == 2) issue_taps<2>(d_tmem, a_lo, b_lo, tap16, b16, idesc, g == 0, nk);
              else issue_taps<4>(d_tmem, a_lo, b_lo, tap16, b16, idesc, g == 0, nk);
            } else {
              for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  for (int sub = 0; sub < 2; ++sub)
                    umma_bf16_ss(d_tmem + sub * p.n_sub, umma_desc_lo(a_lo + r * tap16 + 2 * k),
                                 umma_desc_lo(b_lo + r * b16 + sub * sub16 + 2 * k), idesc, (g | r | k) != 0 ? 1u : 0u);
              }
            }
            if (nt == n_inner - 1) umma_commit(&aux->a_empty[s]);   // box reusable once the last MMAs have read it
          }
          __syncwarp();
          b_lo += rb16;
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(&aux->tmem_full[as]);   // accumulators of this (tile, channel tile) complete
        __syncwarp();
      }
    }
    }  // generic path
  } else if (warp >= 2) {
    // ============================================================== epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int half = ew >> 2;                              // which half of the columns this warp writes
    const int et = ew * 32 + lane;                         // 0..255 among the epilogue threads
    const int m = q * 32 + lane;
    const int lw = m % p.tw, lh = (m / p.tw) % p.th, lb = m / (p.tw * p.th);
    const uint32_t flags = kflags;
    const int nchunk = p.n_tile / 32;
    const int per_half = (nchunk + 1) >> 1;
    // Two ways to split the epilogue between the two groups of four warps: by column halves of the same tile, or
    // (narrow tiles) by tile parity -- each group then owns whole rows of every other tile, the two tiles' epilogues
    // overlap, the row's sum of squares needs no exchange and every barrier is group-local.
    const bool alt = p.epi_alt != 0;
    const int gi = alt ? half : 0;                         // this group's slot in the per-group shared-memory vectors
    const int et_g = alt ? (et & 127) : et;                // thread index inside the synchronising group
    const int g_threads = alt ? 128 : kEpiThreads;
    const int t_step = alt ? 2 : 1;
    const int c_lo = alt ? 0 : half * per_half, c_hi = alt ? nchunk : min(nchunk, c_lo + per_half);
    const bool tile_ss = (flags & CCDM_EPI_SS) && p.tb == 1;   // scale/shift uniform over the tile
    const bool use_rs = (flags & CCDM_EPI_ROWSCALE) != 0;
    const uint32_t trow_lane = static_cast<uint32_t>(q * 32) << 16;
    const int acc_mask = p.acc_stages - 1, acc_shift = p.acc_stages >> 1;     // acc_stages is 1 or 2
    const int ob_mask = p.out_bufs > 1 ? 1 : 0;
    const int npanels = (p.n_tile + 63) / 64;
    uint32_t r[32];
    griddep_wait();                                        // rowss / scale-shift / residual reads and all stores below
    int ss_b = -1;                                         // sample whose scale/shift currently sits in aux->gs/sh
    if (kStoreTma && et_g == 0) {
      for (int i = 0; i < CCDM_MAX_Z; ++i) tma_prefetch_desc(&maps.o[i]);
    }
    // tile coordinates advance incrementally (the CTA's tiles are contiguous): no divisions in the loop
    const int t_first = t_begin + (alt ? half : 0);
    int twi = t_first % p.tiles_w, thi = (t_first / p.tiles_w) % p.tiles_h, tbi = t_first / tiles_per_sample;
    auto advance = [&]() { if (++twi == p.tiles_w) { twi = 0; if (++thi == p.tiles_h) { thi = 0; ++tbi; } } };

    auto row_of = [&](int w0, int h0, int b0, bool& valid, int& bs, long long& pix) {
      const int w = w0 + lw, h = h0 + lh, b = b0 + lb;
      valid = (w < p.gW) && (h < p.gH) && (b < p.gB);
      bs = b < p.gB ? b : p.gB - 1;
      pix = (static_cast<long long>(bs) * p.gH + (h < p.gH ? h : 0)) * p.gW + (w < p.gW ? w : 0);
    };
    float rss_next = 1.f;                                  // A-row sum of squares, fetched one tile ahead
    if (use_rs && t_first < t_end) {
      bool v0; int b_; long long px;
      row_of(twi * p.tw, thi * p.th, tbi * p.tb, v0, b_, px);
      if (v0) rss_next = __ldg(p.rowss + px);
    }

    int lt = alt ? half - 2 : -1;                          // accumulator-use counter: (pixel tile, channel tile) pairs
    for (int tile = t_first; tile < t_end; tile += t_step) {
      const int w0 = twi * p.tw, h0 = thi * p.th, b0 = tbi * p.tb;
      advance();
      if (alt) advance();                                  // (twi, thi, tbi) now name this group's next tile
      bool valid; int bs; long long pix;
      row_of(w0, h0, b0, valid, bs, pix);
      const int w = w0 + lw, h = h0 + lh;

      if (tile_ss && b0 != ss_b) {                         // new sample (uniform over the group's threads): refresh
        epi_bar_g(alt, half);                              // everyone is done with the previous sample's vectors
        const float* ssrow = p.ss + static_cast<long long>(min(b0, p.gB - 1)) * p.ss_ld + p.ss_off + n_base;
        for (int c = et_g; c < p.n_tile; c += g_threads) {
          const bool ok = (n_base + c) < p.N;
          aux->gs[gi][c] = ok ? ((flags & CCDM_EPI_RMSNORM) ? aux->gain[c] : 1.f) * (1.f + ssrow[c]) : 0.f;
          aux->sh[gi][c] = ok ? ssrow[p.N + c] : 0.f;
        }
        epi_bar_g(alt, half);
        ss_b = b0;
      }
      // global operands are requested BEFORE waiting for the accumulators so their latency overlaps the MMAs:
      // next tile's A-row sum of squares and, when it fits in registers, this tile's residual rows
      const float rss_raw = rss_next;
      if (use_rs && tile + t_step < t_end) {
        bool v1; int b_; long long px;
        row_of(twi * p.tw, thi * p.th, tbi * p.tb, v1, b_, px);
        rss_next = v1 ? __ldg(p.rowss + px) : 1.f;
      }
      const long long ro = static_cast<long long>(bs) * p.rsB + static_cast<long long>(h) * p.rsH +
                           static_cast<long long>(w) * p.rsW + n_base;
      const bool pre_res = (flags & CCDM_EPI_RESID) && valid && (c_hi - c_lo) <= 2;
      uint4 rpre[2][4];
      if (pre_res) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (c_lo + cc < c_hi) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.resid + ro + (c_lo + cc) * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              rpre[cc][g] = (n_base + (c_lo + cc) * 32 + g * 8 < p.N) ? __ldg(rp + g) : make_uint4(0, 0, 0, 0);
          }
        }
      }

      for (int nt = 0; nt < p.n_inner; ++nt) {
      lt += t_step;
      const int as = lt & acc_mask;
      const uint32_t aph = (lt >> acc_shift) & 1;
      const int n0 = n_base + nt * p.n_tile;               // first output channel of this item
      const float* const s_bias = aux->bias + nt * p.n_tile;
      mbar_wait(&aux->tmem_full[as], aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + trow_lane + as * p.n_tile;
      const float rs = use_rs ? (valid ? 1.f / fmaxf(sqrtf(rss_raw), 1e-12f) : 0.f) : 1.f;
      const float2 rs2 = make_float2(rs, rs);

      float inv = 1.f;
      if (flags & CCDM_EPI_RMSNORM) {
        float2 sq = make_float2(0.f, 0.f);
        for (int c = 0; c < nchunk; ++c) {
          tmem_ld32(trow + c * 32, r);
          tmem_ld_wait();
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 a = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            const float2 v = use_rs ? __ffma2_rn(a, rs2, b2[i]) : __fadd2_rn(a, b2[i]);
            sq = __ffma2_rn(v, v, sq);
          }
        }
        inv = 1.f / fmaxf(sqrtf(sq.x + sq.y), 1e-12f);
      }
      const float2 inv2 = make_float2(inv, inv);

      const float* ssrow = ((flags & CCDM_EPI_SS) && !tile_ss)
                               ? p.ss + static_cast<long long>(bs) * p.ss_ld + p.ss_off + n0 : nullptr;
      const long long oo = p.ooff[z] + static_cast<long long>(bs) * p.osB + static_cast<long long>(h) * p.osH +
                           static_cast<long long>(w) * p.osW + n0;
      const uint32_t dense_pitch = p.out_pitch;            // != 0: whole pixel rows, un-swizzled (see the launcher)
      uint8_t* const stg_row = stg + static_cast<size_t>(lt & ob_mask) * p.out_bytes + m * (dense_pitch ? dense_pitch : 128u);
      float hd[4] = {0.f, 0.f, 0.f, 0.f};                  // CCDM_EPI_HEAD: this thread's partial head outputs
      float out_ss = 0.f, out_ss_hi = 0.f;                 // alternate-tile mode keeps the two column halves apart so that
                                                           // the sum rounds exactly as in the split mode (batch-shard invariance)
      if (kStoreTma && alt) {
        // A group re-uses its ONE staging buffer for every tile it owns: the bulk store of its previous tile must have
        // finished reading the buffer before anyone overwrites it (in split mode the two buffers alternate instead).
        if (et_g == 0) tma_store_wait_read0();
        epi_bar_g(true, half);
      }

      // The chunk body is instantiated per column class (0 plain, 1 q-softmax, 2 k-exp) so that the branches do not
      // merge with 32 live values (the merge cost one register move per value and chunk).
      auto chunk_body = [&](const int c, auto cmode_tag) {
        constexpr int kCMode = decltype(cmode_tag)::value;
        tmem_ld32(trow + c * 32, r);
        tmem_ld_wait();
        auto release_acc = [&]() {                         // last TMEM read of this tile by this warp: release the stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kPair) mbar_arrive_cluster(&aux->tmem_empty[as], 0);
            else mbar_arrive(&aux->tmem_empty[as]);
          }
        };
        if (c == c_hi - 1 && !(flags & CCDM_EPI_RESACC)) release_acc();
        float2 v[16];
        {
          const float2* b2 = reinterpret_cast<const float2*>(s_bias + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 a = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            v[i] = use_rs ? __ffma2_rn(a, rs2, b2[i]) : __fadd2_rn(a, b2[i]);
          }
        }
        if (flags & CCDM_EPI_RMSNORM) {
          if (tile_ss) {
            const float2* g2 = reinterpret_cast<const float2*>(aux->gs[gi] + c * 32);
            const float2* s2 = reinterpret_cast<const float2*>(aux->sh[gi] + c * 32);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(__fmul2_rn(v[i], inv2), g2[i], s2[i]);
          } else {
            const float2* g2 = reinterpret_cast<const float2*>(aux->gain + c * 32);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(__fmul2_rn(v[i], inv2), g2[i]);
          }
        }
        if (!(flags & CCDM_EPI_RMSNORM) && tile_ss) {      // conditional batch norm: per-(sample, channel) affine map
          const float2* g2 = reinterpret_cast<const float2*>(aux->gs[gi] + c * 32);
          const float2* s2 = reinterpret_cast<const float2*>(aux->sh[gi] + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], g2[i], s2[i]);
        }
        if (ssrow) {                                       // tiles spanning several samples: per-row scale/shift
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (n0 + c * 32 + 2 * i < p.N) {               // N is a multiple of 8: pairs are all-in or all-out
              const float2 sc = __ldg(reinterpret_cast<const float2*>(ssrow + c * 32) + i);
              const float2 sf = __ldg(reinterpret_cast<const float2*>(ssrow + p.N + c * 32) + i);
              v[i] = __ffma2_rn(v[i], make_float2(1.f + sc.x, 1.f + sc.y), sf);
            }
          }
        }
        if (flags & CCDM_EPI_SILU) {                       // x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU per element
          const float2 half2 = make_float2(0.5f, 0.5f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 hx = __fmul2_rn(v[i], half2);
            const float2 t = make_float2(tanh_silu(hx.x), tanh_silu(hx.y));
            v[i] = __ffma2_rn(hx, t, hx);
          }
        }
        if (flags & CCDM_EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
        }
        if (flags & CCDM_EPI_TANH) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(tanh_fast(v[i].x), tanh_fast(v[i].y));
        }
        if constexpr (kCMode == 1) {
          float mx = fmaxf(v[0].x, v[0].y);
#pragma unroll
          for (int i = 1; i < 16; ++i) mx = fmaxf(mx, fmaxf(v[i].x, v[i].y));
          const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
          const float2 nmx = make_float2(-mx * 1.4426950408889634f, -mx * 1.4426950408889634f);
          float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 e = __ffma2_rn(v[i], l2e, nmx);
            v[i].x = ex2_fast(e.x);
            v[i].y = ex2_fast(e.y);
            sum2 = __fadd2_rn(sum2, v[i]);
          }
          const float k = __fdividef(p.q_scale, sum2.x + sum2.y);
          const float2 k2 = make_float2(k, k);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __fmul2_rn(v[i], k2);
        }
        if constexpr (kCMode == 2) {
          const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 e = __fmul2_rn(v[i], l2e);
            v[i].x = ex2_fast(e.x);
            v[i].y = ex2_fast(e.y);
          }
        }
        if (flags & CCDM_EPI_RESACC) {                     // shortcut from the second accumulator (+ res_conv bias)
          tmem_ld32(trow + static_cast<uint32_t>(p.acc_stages * p.n_tile) + c * 32, r);
          tmem_ld_wait();
          if (c == c_hi - 1) release_acc();
          const float2* b2 = reinterpret_cast<const float2*>(aux->bias2 + c * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = __fadd2_rn(v[i], __fadd2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]));
        }
        if ((flags & CCDM_EPI_RESID) && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.resid + ro + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (n0 + c * 32 + g * 8 < p.N) {
              const uint4 u = pre_res ? (((c - c_lo) & 1) ? rpre[1][g] : rpre[0][g]) : __ldg(rp + g);
              v[g * 4 + 0] = __fadd2_rn(v[g * 4 + 0], make_float2(bf16_lo(u.x), bf16_hi(u.x)));
              v[g * 4 + 1] = __fadd2_rn(v[g * 4 + 1], make_float2(bf16_lo(u.y), bf16_hi(u.y)));
              v[g * 4 + 2] = __fadd2_rn(v[g * 4 + 2], make_float2(bf16_lo(u.z), bf16_hi(u.z)));
              v[g * 4 + 3] = __fadd2_rn(v[g * 4 + 3], make_float2(bf16_lo(u.w), bf16_hi(u.w)));
            }
          }
        }
        if (flags & CCDM_EPI_HEAD) {                       // the tile is consumed here: 1x1 head, nothing else is stored
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < p.head_n) {
              const float2* w2 = reinterpret_cast<const float2*>(aux->headw[k] + c * 32);
              float2 s2 = make_float2(hd[k], 0.f);
#pragma unroll
              for (int i = 0; i < 16; ++i) s2 = __ffma2_rn(v[i], w2[i], s2);
              hd[k] = s2.x + s2.y;
            }
          }
        } else if (flags & CCDM_EPI_OUT_F32) {
          if (valid) {
            float* op = reinterpret_cast<float*>(p.out) + oo + c * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (n0 + c * 32 + g * 4 < p.N)
                *reinterpret_cast<float4*>(op + g * 4) = make_float4(v[2 * g].x, v[2 * g].y, v[2 * g + 1].x, v[2 * g + 1].y);
          }
        } else {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + oo + c * 32;
          uint8_t* const srow = stg_row + (c >> 1) * 16384;                // 64-channel staging panel of this chunk
          const int j0 = (c & 1) * 4;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16(v[g * 4 + 0].x, v[g * 4 + 0].y);
            u.y = pack_bf16(v[g * 4 + 1].x, v[g * 4 + 1].y);
            u.z = pack_bf16(v[g * 4 + 2].x, v[g * 4 + 2].y);
            u.w = pack_bf16(v[g * 4 + 3].x, v[g * 4 + 3].y);
            if (kStoreTma) {
              if (dense_pitch) {                                         // dense pixel row
                if (n0 + c * 32 + g * 8 < p.N) *reinterpret_cast<uint4*>(stg_row + c * 64 + g * 16) = u;
              } else {                                                   // swizzled like a TMA SWIZZLE_128B box
                *reinterpret_cast<uint4*>(srow + (((j0 + g) ^ (m & 7)) << 4)) = u;
              }
            }
            else if (valid && n0 + c * 32 + g * 8 < p.N)
              *reinterpret_cast<uint4*>(op + g * 8) = u;
            if ((flags & CCDM_EPI_SUMSQ_OUT) && n0 + c * 32 + g * 8 < p.N) {
              const float a0 = bf16_lo(u.x), a1 = bf16_hi(u.x), a2 = bf16_lo(u.y), a3 = bf16_hi(u.y);
              const float a4 = bf16_lo(u.z), a5 = bf16_hi(u.z), a6 = bf16_lo(u.w), a7 = bf16_hi(u.w);
              const float sq8 = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
              if (alt && c >= per_half) out_ss_hi += sq8;
              else out_ss += sq8;
            }
          }
        }
      };
      for (int c = c_lo; c < c_hi; ++c) {
        const int col = n0 + c * 32;
        if ((flags & CCDM_EPI_QSOFTMAX) && col < p.q_cols) chunk_body(c, std::integral_constant<int, 1>{});
        else if ((flags & CCDM_EPI_KEXP) && col >= p.q_cols && col < 2 * p.q_cols) chunk_body(c, std::integral_constant<int, 2>{});
        else chunk_body(c, std::integral_constant<int, 0>{});
      }
      if (c_lo >= c_hi) {                                  // this warp owns no columns (n_tile == 32): still release
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_cluster(&aux->tmem_empty[as], 0);
          else mbar_arrive(&aux->tmem_empty[as]);
        }
      }
      if (flags & CCDM_EPI_HEAD) {
        if (!alt) {                                        // split mode: the two column halves hold partial dot products
          if (half == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) aux->hpart[lt & 1][k][m] = hd[k];
          }
          epi_bar();
          if (half == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) hd[k] += aux->hpart[lt & 1][k][m];
          }
        }
        if (valid && (alt || half == 0)) {
          float* ho = p.head_out + static_cast<long long>(bs) * p.hsB + static_cast<long long>(h) * p.gW + w;
          for (int k = 0; k < p.head_n; ++k) ho[k * p.hsC] = hd[k] + __ldg(p.head_b + k);
        }
      }
      if ((flags & CCDM_EPI_SUMSQ_OUT) && !alt && half == 1) aux->part[lt & 1][m] = out_ss;   // combine the column halves
      if (kStoreTma) {
        // staging complete -> one thread hands the tile to the TMA unit.  Before the barrier it makes sure the
        // PREVIOUS tile's bulk stores have finished reading their buffer, which the next tile will overwrite.
        fence_proxy_async_smem();
        if (et_g == 0) tma_store_wait_read0();
        epi_bar_g(alt, half);
        if (et_g == 0) {
          const uint8_t* sbuf = stg + static_cast<size_t>(lt & ob_mask) * p.out_bytes;
          if (p.out_pitch) {
            tma_store_4d(&maps.o[z], sbuf, n0, w0, h0, b0);
          } else {
            for (int pn = 0; pn < npanels; ++pn)
              if (n0 + pn * 64 < p.N) tma_store_4d(&maps.o[z], sbuf + pn * 16384, n0 + pn * 64, w0, h0, b0);
          }
          tma_store_commit();
          if (p.out_bufs == 1) tma_store_wait_read0();
        }
        if (p.out_bufs == 1) epi_bar_g(alt, half);
      } else if ((flags & CCDM_EPI_SUMSQ_OUT) && !alt) {
        epi_bar();
      }
      if ((flags & CCDM_EPI_SUMSQ_OUT) && valid) {
        if (alt) p.out_rowss[pix] = out_ss + out_ss_hi;
        else if (half == 0) p.out_rowss[pix] = out_ss + aux->part[lt & 1][m];
      }
      }  // channel tiles of this pixel tile
    }
  }

  // ---------------------------------------------------------------- teardown
  if (kStoreTma && (tid == 64 || (p.epi_alt && tid == 64 + 128))) tma_store_wait_all();   // the threads that issued bulk stores
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();                 // the peer is done with this CTA's barriers, shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_2sm(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ============================================================================ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  CCDM_REQUIRE(enc != nullptr, CCDM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  // The driver entry point needs a current context on the calling thread.  Threads that have not made a runtime call
  // yet (PyTorch's autograd workers run the backward nodes) bind the primary context here, once.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CCDM_REQUIRE(r == CUDA_SUCCESS, CCDM_ERR_BAD_ARG,
               "cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu "
               "box=%u,%u,%u,%u",
               (int)r, base, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
               (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
               (unsigned long long)strides_b[0], (unsigned long long)(rank > 2 ? strides_b[1] : 0),
               (unsigned long long)(rank > 3 ? strides_b[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
               rank > 3 ? box[3] : 0);
  return CCDM_OK;
}

// Un-swizzled box (inner extent up to 256 elements): the dense-row output staging of layers whose channel count is not a
// multiple of 64.
static int encode_map_bf16_plain(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                                 const cuuint64_t* strides_b, const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  CCDM_REQUIRE(enc != nullptr, CCDM_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CCDM_REQUIRE(r == CUDA_SUCCESS, CCDM_ERR_BAD_ARG, "cuTensorMapEncodeTiled (plain) failed (%d): box=%u,%u,%u,%u", (int)r,
               box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return CCDM_OK;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (c < static_cast<uint32_t>(n)) c <<= 1;
  return c;
}

template <uint32_t kFlags, bool kStoreTma, bool kPair = false>
static int launch_one(dim3 grid, size_t smem_bytes, cudaStream_t stream, const TapGemmMaps& maps, const TapGemmDev& p) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(tapgemm_kernel<kFlags, kStoreTma, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024);
  });
  if (err != cudaSuccess) return cuda_fail(err, "tapgemm: cudaFuncSetAttribute");
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (kPair) {                                             // CTA pairs: 2-CTA clusters along x (one TPC each)
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, tapgemm_kernel<kFlags, kStoreTma, kPair>, maps, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return cuda_fail(e, "tapgemm_kernel launch");
  return CCDM_OK;
}

// The epilogue flag sets the UNet actually uses are compiled as separate instantiations; anything else takes the
// runtime-flag instantiation (same code, branches kept).
static int launch_variant(uint32_t flags, bool tma, dim3 grid, size_t smem_bytes, cudaStream_t stream,
                          const TapGemmMaps& maps, const TapGemmDev& p) {
  if (p.pair) {                                            // CTA-pair builds (always TMA-store epilogues)
#define CCDM_PAIR_VARIANT(F) \
  case F:                    \
    return launch_one<F, true, true>(grid, smem_bytes, stream, maps, p);
    switch (flags) {
      CCDM_PAIR_VARIANT(0x1Du)
      CCDM_PAIR_VARIANT(0x35u)
      CCDM_PAIR_VARIANT(0xB5u)
      CCDM_PAIR_VARIANT(0x01u)
      CCDM_PAIR_VARIANT(0x25u)
      CCDM_PAIR_VARIANT(0x00u)
      CCDM_PAIR_VARIANT(0x2015u)
      CCDM_PAIR_VARIANT(0x2095u)
      default:
        break;
    }
#undef CCDM_PAIR_VARIANT
    return launch_one<kRuntimeFlags, true, true>(grid, smem_bytes, stream, maps, p);
  }
#define CCDM_VARIANT(F)                                                                        \
  case F:                                                                                      \
    return tma ? launch_one<F, true>(grid, smem_bytes, stream, maps, p)                        \
               : launch_one<F, false>(grid, smem_bytes, stream, maps, p);
  switch (flags) {
    CCDM_VARIANT(0x1Du)   // bias | rmsnorm | scale-shift | silu            (Block 1 of a ResnetBlock)
    CCDM_VARIANT(0x35u)   // bias | rmsnorm | silu | resid                  (Block 2 + residual)
    CCDM_VARIANT(0xB5u)   //   ... + sum of squares for the next PreNorm
    CCDM_VARIANT(0x01u)   // bias                                           (res_conv, down / up sampling convs)
    CCDM_VARIANT(0x243u)  // bias | rowscale | q-softmax | k-exp            (linear-attention qkv)
    CCDM_VARIANT(0x02u)   // rowscale                                       (bottleneck-attention qkv)
    CCDM_VARIANT(0x25u)   // bias | rmsnorm | resid                         (linear-attention to_out)
    CCDM_VARIANT(0x21u)   // bias | resid                                   (bottleneck-attention to_out)
    CCDM_VARIANT(0x00u)   // plain                                          (data gradients, per-sample context products)
    CCDM_VARIANT(0x409u)  // bias | scale-shift | relu                      (generator conv1 + CondBN + ReLU)
    CCDM_VARIANT(0x2015u) // bias | rmsnorm | silu | shortcut accumulator   (Block 2 + res_conv / identity in TMEM)
    CCDM_VARIANT(0x2095u) //   ... + sum of squares for the next PreNorm
    default:
      break;
  }
#undef CCDM_VARIANT
  if (flags == 0x101u) return launch_one<0x101u, false>(grid, smem_bytes, stream, maps, p);   // tc_mlp row-GEMM, fp32 out
  if (flags == 0x1035u) return launch_one<0x1035u, false>(grid, smem_bytes, stream, maps, p); // final block 2 + residual + 1x1 head
  if (flags == 0x3015u) return launch_one<0x3015u, false>(grid, smem_bytes, stream, maps, p); //   ... with the shortcut accumulator
  if (flags == 0x901u) return launch_one<0x901u, false>(grid, smem_bytes, stream, maps, p);   // generator output: tanh, fp32
  return tma ? launch_one<kRuntimeFlags, true>(grid, smem_bytes, stream, maps, p)
             : launch_one<kRuntimeFlags, false>(grid, smem_bytes, stream, maps, p);
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_tapgemm(const ccdm_tapgemm_args* a, void* stream) {
  CCDM_REQUIRE(a != nullptr, CCDM_ERR_BAD_ARG, "tapgemm: null args");
  CCDM_REQUIRE(a->n_src >= 1 && a->n_src <= CCDM_MAX_SRC, CCDM_ERR_BAD_ARG, "tapgemm: n_src=%d", a->n_src);
  CCDM_REQUIRE(a->tw > 0 && a->th > 0 && a->tb > 0 && a->tw * a->th * a->tb == kTileM, CCDM_ERR_BAD_ARG,
               "tapgemm: tile box %dx%dx%d must hold 128 positions", a->tw, a->th, a->tb);
  CCDM_REQUIRE(a->nz >= 1 && a->nz <= CCDM_MAX_Z && a->ngroups >= 1 && a->R >= 1 && (a->R <= 4 || (a->halo && a->R == 9)),
               CCDM_ERR_BAD_ARG, "tapgemm: nz=%d ngroups=%d R=%d", a->nz, a->ngroups, a->R);
  CCDM_REQUIRE(!a->halo || (a->R == 9 && a->tb == 1 && a->tw == 8 && a->nz == 1 && a->w_batch_rows == 0 && a->n_tile <= 256),
               CCDM_ERR_BAD_ARG, "tapgemm: halo boxes need R == 9, tile 8 x th x 1, nz == 1, shared weights, n_tile <= 256");
  CCDM_REQUIRE(a->R == 1 || (a->tb == 1 && a->tw % 8 == 0), CCDM_ERR_BAD_ARG,
               "tapgemm: vertical tap reuse (R=%d) needs tb == 1 and tw %% 8 == 0 (tile %dx%dx%d)", a->R, a->tw, a->th,
               a->tb);
  CCDM_REQUIRE(a->n_tile >= 32 && a->n_tile <= kMaxN && a->n_tile % 32 == 0, CCDM_ERR_UNSUPPORTED_SHAPE,
               "tapgemm: n_tile=%d must be a multiple of 32 in [32,512]", a->n_tile);
  CCDM_REQUIRE(a->N >= 1 && a->N % 8 == 0 && a->n_rows % a->n_tile == 0 && a->n_rows >= a->N,
               CCDM_ERR_UNSUPPORTED_SHAPE, "tapgemm: N=%d n_rows=%d n_tile=%d", a->N, a->n_rows, a->n_tile);
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_RMSNORM) || a->n_rows == a->n_tile, CCDM_ERR_UNSUPPORTED_SHAPE,
               "tapgemm: the RMSNorm epilogue needs all %d channels in one tile (n_tile=%d)", a->N, a->n_tile);
  CCDM_REQUIRE(a->sched && a->wpacked && (a->out || (a->flags & CCDM_EPI_HEAD)), CCDM_ERR_BAD_ARG,
               "tapgemm: null sched/wpacked/out");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_HEAD) ||
                   (a->head_w && a->head_b && a->head_out && a->head_n >= 1 && a->head_n <= 4 && a->n_rows == a->n_tile &&
                    a->n_tile <= 128 && a->nz == 1 && !(a->flags & (CCDM_EPI_OUT_F32 | CCDM_EPI_SUMSQ_OUT))),
               CCDM_ERR_BAD_ARG, "tapgemm: CCDM_EPI_HEAD needs head_w/head_b/head_out, 1 <= head_n <= 4, one N tile <= 128, nz == 1");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_BIAS) || a->bias, CCDM_ERR_BAD_ARG, "tapgemm: bias flag without pointer");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_ROWSCALE) || a->rowss, CCDM_ERR_BAD_ARG, "tapgemm: rowscale without rowss");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_RMSNORM) || a->gain, CCDM_ERR_BAD_ARG, "tapgemm: rmsnorm without gain");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_SS) || a->scale_shift, CCDM_ERR_BAD_ARG,
               "tapgemm: scale/shift flag without pointer (unet.py:145-149, sngan.py:28-36)");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_RESID) || a->resid, CCDM_ERR_BAD_ARG, "tapgemm: resid flag without pointer");
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_SUMSQ_OUT) || (a->out_rowss && a->n_rows == a->n_tile), CCDM_ERR_BAD_ARG,
               "tapgemm: sumsq output needs out_rowss and a single N tile");
  CCDM_REQUIRE(!(a->flags & (CCDM_EPI_QSOFTMAX | CCDM_EPI_KEXP)) || (a->q_cols > 0 && a->q_cols % 32 == 0),
               CCDM_ERR_BAD_ARG, "tapgemm: q_cols must be a positive multiple of 32");
  CCDM_REQUIRE(a->w_batch_rows == 0 || (a->tb == 1 && a->nz == 1 && a->w_batch_rows >= a->n_rows), CCDM_ERR_BAD_ARG,
               "tapgemm: per-sample weights need tb == 1, nz == 1 and w_batch_rows >= n_rows");

  const int box_h = a->halo ? a->th + 2 : a->th + a->R - 1;
  const int box_w = a->halo ? a->tw + 2 : a->tw;
  CCDM_REQUIRE(box_w <= 256 && box_h <= 256 && a->tb <= 256, CCDM_ERR_BAD_ARG, "tapgemm: TMA box too large");
  TapGemmMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < CCDM_MAX_SRC; ++i) {
    const ccdm_view& v = a->src[i < a->n_src ? i : 0];
    CCDM_REQUIRE(v.ptr && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, CCDM_ERR_BAD_ARG,
                 "tapgemm: source %d pointer must be 16-byte aligned", i);
    CCDM_REQUIRE(v.C > 0 && v.W > 0 && v.H > 0 && v.B > 0 && v.sW % 8 == 0 && v.sH % 8 == 0 && v.sB % 8 == 0,
                 CCDM_ERR_BAD_ARG, "tapgemm: source %d extents/strides (strides must be multiples of 8 elements)", i);
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.B};
    cuuint64_t str[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sB * 2};
    cuuint32_t box[4] = {kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)a->tb};
    int rc = encode_map_bf16(&maps.a[i], v.ptr, 4, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  const int nsub = a->n_tile > 256 ? 2 : 1;
  const int n_sub = a->n_tile / nsub;
  const int n_res = (a->flags & CCDM_EPI_RESACC) ? a->n_res : 0;
  CCDM_REQUIRE(!(a->flags & CCDM_EPI_RESACC) ||
                   (a->n_res >= 1 && a->nz == 1 && a->n_tile <= 128 && a->w_batch_rows == 0 && (a->R <= 3 || a->halo) &&
                    !(a->flags & CCDM_EPI_RESID)),
               CCDM_ERR_BAD_ARG,
               "tapgemm: CCDM_EPI_RESACC needs n_res >= 1, nz == 1, n_tile <= 128, shared weights, R <= 3 (or halo boxes) and no CCDM_EPI_RESID");
  const int nkb = a->ngroups * a->R + n_res;               // K blocks of the packed weights: main taps, then the shortcut's
  if (n_res > 0) {                                         // shortcut boxes: th rows (one tap, no halo rows)
    for (int i = 0; i < CCDM_MAX_SRC; ++i) {
      const ccdm_view& v = a->src[i < a->n_src ? i : 0];
      cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.B};
      cuuint64_t str[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sB * 2};
      cuuint32_t box[4] = {kBlockK, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->tb};
      int rc = encode_map_bf16(&maps.r[i], v.ptr, 4, dims, str, box);
      if (rc != CCDM_OK) return rc;
    }
  }
  CCDM_REQUIRE(n_sub % 16 == 0 && n_sub <= 256, CCDM_ERR_UNSUPPORTED_SHAPE, "tapgemm: n_sub=%d", n_sub);

  TapGemmDev p;
  std::memset(&p, 0, sizeof(p));
  p.gW = a->gW; p.gH = a->gH; p.gB = a->gB;
  p.tw = a->tw; p.th = a->th; p.tb = a->tb;
  p.tiles_w = (a->gW + a->tw - 1) / a->tw;
  p.tiles_h = (a->gH + a->th - 1) / a->th;
  const int tiles_b = (a->gB + a->tb - 1) / a->tb;
  p.tiles_m = p.tiles_w * p.tiles_h * tiles_b;
  p.n_tiles = a->n_rows / a->n_tile;
  p.n_inner = 1;
  p.ngroups = a->ngroups; p.R = a->R; p.nkb = nkb;
  p.w_batch_rows = a->w_batch_rows;
  p.n_rows = a->n_rows; p.N = a->N; p.n_tile = a->n_tile; p.n_sub = n_sub; p.nsub = nsub;
  p.sched = reinterpret_cast<const int4*>(a->sched);
  p.ksteps = a->halo ? nullptr : a->ksteps;
  p.flags = a->flags;
  p.acc_stages = a->n_tile <= 256 ? 2 : 1;
  p.tmem_cols = pow2_cols(a->n_tile * p.acc_stages * (n_res > 0 ? 2 : 1));
  p.n_res = n_res;
  p.res_bias = a->res_bias;
  p.a_bytes_res = (uint32_t)(a->th * a->tw * a->tb) * 128u;
  p.bias = a->bias; p.rowss = a->rowss; p.gain = a->gain; p.ss = a->scale_shift;
  p.ss_ld = a->ss_ld; p.ss_off = a->ss_off;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(a->resid);
  p.rsW = a->rsW; p.rsH = a->rsH; p.rsB = a->rsB;
  p.out = a->out; p.osW = a->osW; p.osH = a->osH; p.osB = a->osB;
  for (int i = 0; i < CCDM_MAX_Z; ++i) p.ooff[i] = a->ooff[i];
  p.out_rowss = a->out_rowss; p.q_scale = a->q_scale; p.q_cols = a->q_cols; p.gain_mul = a->gain_mul;
  p.head_n = a->head_n; p.head_w = a->head_w; p.head_b = a->head_b; p.head_out = a->head_out; p.hsC = a->hsC; p.hsB = a->hsB;

  // ---- channel tiles: by default one (z, n-tile) combination per blockIdx.y.  When the layer has several channel
  // tiles with very different epilogue costs (linear-attention qkv: softmax | exp | copy) and all their weights fit in
  // shared memory, every CTA walks ALL channel tiles of its pixel tiles instead: balanced, and A is fetched once.
  {
    const size_t all_b = (size_t)a->n_rows * nkb * 128;
    const bool plain_epi = !(a->flags & (CCDM_EPI_RMSNORM | CCDM_EPI_SS | CCDM_EPI_RESID | CCDM_EPI_SUMSQ_OUT | CCDM_EPI_RESACC));
    if (p.n_tiles > 1 && a->nz == 1 && a->w_batch_rows == 0 && plain_epi && a->n_rows <= kMaxN &&
        all_b <= 64 * 1024 && a->ngroups <= 2 && a->n_tile <= 256) {
      p.n_inner = p.n_tiles;
      p.n_tiles = 1;
    }
  }
  // ---- launch geometry: persistent CTAs, one (z, n-tile) combination per blockIdx.y
  const int combos = p.n_tiles * a->nz;
  const int sms = num_sms();
  int gx = sms / combos;                                           // never more CTAs than SMs: no second wave
  if (gx > p.tiles_m) gx = p.tiles_m;
  if (gx < 1) gx = 1;
  int tiles_per_cta = (p.tiles_m + gx - 1) / gx;
  gx = (p.tiles_m + tiles_per_cta - 1) / tiles_per_cta;            // contiguous ranges: no CTA without work
  // ---- CTA pairs (tcgen05 cta_group::2): two CTAs on one TPC run ONE M = 256 instruction per K step, each on its own
  // pixel tile, each holding half of the weight rows.  Halves the single issuing lane's work per tile (the measured
  // limiter of the 64-channel layers), the weight bytes in shared memory (128 -> 64 3x3 weights become resident) and the
  // B-operand fetch per MMA.  The i-th tile of the even CTA is multiplied together with the i-th tile of the odd CTA.
  {
    // Measured (profiles/r2_pair_layers.txt): pairs win where the K loop is long (>= 16 K blocks: 128 -> 64 3x3 +33 %,
    // 4x4-s2, every >= 128-channel 3x3) and lose on short, memory-bound layers (1x1, 64 -> 64 + residual), whose
    // per-tile time is dominated by load / store latency that the pair's cross-SM handshakes lengthen.
    static const int pair_env = [] { const char* e = getenv("CCDM_TAPGEMM_PAIR"); return e ? atoi(e) : 16; }();
    static const int pair_min_tiles = [] { const char* e = getenv("CCDM_TAPGEMM_PAIR_MINTILES"); return e ? atoi(e) : 1; }();
    p.pair = (pair_env != 0 && nkb >= pair_env && a->w_batch_rows == 0 && nsub == 1 && p.n_inner == 1 && (a->R <= 3 || a->halo) && a->n_tile <= 256 &&
              (a->n_tile / 2) % 16 == 0 && gx >= 2 && tiles_per_cta >= pair_min_tiles &&
              !(a->flags & (CCDM_EPI_OUT_F32 | CCDM_EPI_HEAD))) ? 1 : 0;
    if (p.pair) {
      if (gx & 1) ++gx;                                            // whole pairs; a trailing CTA may get dummy tiles only
      if (gx > sms / combos) {                                     // odd SM budget: re-balance over an even CTA count
        gx = (sms / combos) & ~1;
        tiles_per_cta = (p.tiles_m + gx - 1) / gx;
        gx = (p.tiles_m + tiles_per_cta - 1) / tiles_per_cta;
        if (gx & 1) ++gx;
      }
    }
  }
  p.tiles_per_cta = tiles_per_cta;
  {
    const cuuint64_t ktot = (cuuint64_t)nkb * kBlockK;
    cuuint64_t dims[2] = {ktot, a->w_batch_rows > 0 ? (cuuint64_t)a->w_batch_rows * a->gB
                                                    : (cuuint64_t)a->n_rows * a->nz};
    cuuint64_t str[1] = {ktot * 2};
    cuuint32_t box[2] = {kBlockK, (cuuint32_t)(p.pair ? n_sub / 2 : n_sub)};
    CCDM_REQUIRE((reinterpret_cast<uintptr_t>(a->wpacked) & 15) == 0, CCDM_ERR_BAD_ARG, "tapgemm: wpacked alignment");
    int rc = encode_map_bf16(&maps.b, a->wpacked, 2, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }

  // ---- shared-memory plan
  const uint32_t b_bytes = (uint32_t)(p.pair ? a->n_tile / 2 : a->n_tile) * 128u;
  const uint32_t a_tx = (uint32_t)(box_h * box_w * a->tb) * 128u;        // bytes one A box delivers
  p.a_bytes = (a_tx + 1023u) & ~1023u;                                   // stage layout: 1 KiB granules (halo boxes: 23040 B)
  p.halo = a->halo ? 1 : 0;
  const size_t aux_bytes = sizeof(TapGemmAux) + (size_t)(a->ngroups + n_res) * (sizeof(int4) + sizeof(int));
  size_t budget = 226 * 1024 - aux_bytes - 1024;
  // bf16 outputs up to 256 channels per CTA leave through shared memory and TMA bulk stores (coalesced, clipped at the
  // tensor edges); wider tiles (4x4 bottleneck layers) and fp32 outputs keep per-thread stores
  p.store_tma = (!(a->flags & (CCDM_EPI_OUT_F32 | CCDM_EPI_HEAD)) && a->n_tile <= 256) ? 1 : 0;
  p.out_bytes = p.store_tma ? (uint32_t)((a->n_tile + 63) / 64) * 16384u : 0;
  p.out_bufs = p.store_tma ? ((tiles_per_cta > 1 && p.out_bytes <= 32768) ? 2 : 1) : 0;
  const size_t res_all = (size_t)nkb * b_bytes * p.n_inner;
  // Resident weights matter more than a second output staging buffer: streamed, the weight blocks of every load group ride
  // with every box and cross L2 -> SM once per TILE (72 -> 72 3x3 at 64x64, batch 128: 1017 MB of L2 reads per launch, half of
  // them weights, 9.7 TB/s = the L2 limit; profiles/r2_notes.md).  Give the buffer up when that is what makes them fit.
  static const int res1_env = [] { const char* e = getenv("CCDM_TAPGEMM_RESIDENT_1BUF"); return e ? atoi(e) : 1; }();
  if (res1_env && p.out_bufs == 2 && a->w_batch_rows == 0 && tiles_per_cta >= 2 &&
      res_all + 3 * (size_t)p.a_bytes > budget - 2 * (size_t)p.out_bytes &&
      res_all + 3 * (size_t)p.a_bytes <= budget - (size_t)p.out_bytes)
    p.out_bufs = 1;
  budget -= (size_t)p.out_bufs * p.out_bytes;
  p.b_resident = (a->w_batch_rows == 0 && (tiles_per_cta >= 2 || p.n_inner > 1) &&
                  res_all + 3 * (size_t)p.a_bytes <= budget) ? 1 : 0;
  if (p.n_inner > 1 && !p.b_resident) {           // cannot happen with the limits above; keep the simple mode if so
    p.n_tiles = p.n_inner;
    p.n_inner = 1;
    CCDM_REQUIRE(false, CCDM_ERR_UNSUPPORTED_SHAPE, "tapgemm: internal: inner channel loop without resident weights");
  }
  p.res_bytes = p.b_resident ? (uint32_t)res_all : 0;
  p.stage_bytes = p.a_bytes + (p.b_resident ? 0 : (uint32_t)a->R * b_bytes);
  p.stage_tx = a_tx + (p.b_resident ? 0 : (uint32_t)a->R * b_bytes);
  int stages = (int)((budget - p.res_bytes) / p.stage_bytes);
  int useful = (a->ngroups + n_res) * (tiles_per_cta > 1 ? 2 : 1); // about two tiles of lookahead ...
  static const int min_stages = [] { const char* e = getenv("CCDM_TAPGEMM_MINSTAGES"); return e ? atoi(e) : 4; }();
  if (tiles_per_cta > 1 && useful < min_stages) useful = min_stages;   // ... but never fewer than 4 boxes in flight
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > useful) stages = useful;
  if (p.n_inner > 1 && stages < 2 * a->ngroups) stages = 2 * a->ngroups <= kMaxStages ? 2 * a->ngroups : a->ngroups;
  // narrow tiles: the two epilogue warp groups take alternate tiles (needs both accumulator stages and both staging
  // buffers, one channel tile per CTA and more than one tile per CTA)
  p.epi_alt = (a->n_tile <= 128 && p.n_inner == 1 && p.acc_stages == 2 && tiles_per_cta >= 2 &&
               (!p.store_tma || p.out_bufs == 2)) ? 1 : 0;
  CCDM_REQUIRE(stages >= 1, CCDM_ERR_UNSUPPORTED_SHAPE,
               "tapgemm: one pipeline stage (%u bytes) does not fit shared memory", p.stage_bytes);
  p.stages = stages;
  const size_t smem_bytes = p.res_bytes + (size_t)stages * p.stage_bytes + (size_t)p.out_bufs * p.out_bytes + aux_bytes + 1024;
  // Outputs whose channel count is not a multiple of 64 (the dim-72 models: 72, 144, ...) leave as ONE un-swizzled box of
  // whole pixel rows.  As 64-channel panels their last panel is a box clipped to a few channels and every 128-byte row of
  // the others starts off a 32-byte sector (row pitch 144 B): partial-sector writes, which tools/ubench/tma_rate.cu measures
  // at 2.5x the cost of aligned rows (profiles/r2_ubench_tma_rate.jsonl) and which made the stores, not the MMAs, the
  // limiter of those layers.
  static const int dense_env = [] { const char* e = getenv("CCDM_TAPGEMM_DENSE_STORE"); return e ? atoi(e) : 1; }();
  p.out_pitch = (dense_env && p.store_tma && a->N > 64 && a->N % 64 != 0 && a->N % 8 == 0 && a->N <= 256 && p.n_tiles == 1 &&
                 p.n_inner == 1)
                    ? (uint32_t)a->N * 2u : 0u;
  if (p.store_tma) {
    CCDM_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->osW % 8 == 0 && a->osH % 8 == 0 && a->osB % 8 == 0,
                 CCDM_ERR_BAD_ARG, "tapgemm: output view must be 16-byte aligned with strides in multiples of 8");
    for (int zz = 0; zz < CCDM_MAX_Z; ++zz) {
      const int zi = zz < a->nz ? zz : 0;
      CCDM_REQUIRE(a->ooff[zi] % 8 == 0, CCDM_ERR_BAD_ARG, "tapgemm: ooff[%d] must be a multiple of 8 elements", zi);
      cuuint64_t dims[4] = {(cuuint64_t)a->N, (cuuint64_t)a->gW, (cuuint64_t)a->gH, (cuuint64_t)a->gB};
      cuuint64_t str[3] = {(cuuint64_t)a->osW * 2, (cuuint64_t)a->osH * 2, (cuuint64_t)a->osB * 2};
      cuuint32_t box[4] = {p.out_pitch ? (cuuint32_t)a->N : kBlockK, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->tb};
      const __nv_bfloat16* obase = reinterpret_cast<const __nv_bfloat16*>(a->out) + a->ooff[zi];
      int rc = p.out_pitch ? encode_map_bf16_plain(&maps.o[zz], obase, 4, dims, str, box)
                           : encode_map_bf16(&maps.o[zz], obase, 4, dims, str, box);
      if (rc != CCDM_OK) return rc;
    }
  }

  dim3 grid((unsigned)gx, (unsigned)combos, 1);
  return launch_variant(a->flags, p.store_tma != 0, grid, smem_bytes, static_cast<cudaStream_t>(stream), maps, p);
}

// ---------------------------------------------------------------------------- weight packing

namespace ccdm {
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin_total, int ntaps,
                                    const int4* __restrict__ psched, int nkb, int n_rows,
                                    const float* __restrict__ cin_gain, float gain_mul,
                                    __nv_bfloat16* __restrict__ out, long long total, int nkb_total, int kb0) {
  // one thread per packed element: logical idx = ((z*n_rows + n)*nkb + kb)*64 + j, stored in a matrix of nkb_total K
  // blocks per row at block kb0 + kb (two weights -- a conv and its block's shortcut -- share one packed matrix)
  for (long long lidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; lidx < total;
       lidx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(lidx & 63);
    const long long t = lidx >> 6;
    const int kb = (int)(t % nkb);
    const long long zn = t / nkb;
    const int n = (int)(zn % n_rows);
    const int z = (int)(zn / n_rows);
    const long long idx = ((zn * nkb_total) + kb0 + kb) * 64 + j;
    const int4 e = psched[z * nkb + kb];
    float acc = 0.f;
    if (n < cout && j < e.y) {
      const int ci = e.x + j;
      const float* wp = w + ((long long)n * cin_total + ci) * ntaps;
      unsigned mask = (unsigned)e.z;
      for (int tp = 0; tp < ntaps; ++tp)
        if (mask & (1u << tp)) acc += wp[tp];
      acc *= gain_mul * (cin_gain ? cin_gain[ci] : 1.f);
    }
    out[idx] = __float2bfloat16(acc);
  }
}
}  // namespace ccdm

extern "C" int ccdm_pack_weights_at(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                                    int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain, float gain_mul,
                                    void* wpacked, int32_t nkb_total, int32_t kb0, void* stream) {
  CCDM_REQUIRE(w && psched && wpacked, CCDM_ERR_BAD_ARG, "pack_weights: null pointer");
  CCDM_REQUIRE(cout > 0 && cin_total > 0 && ntaps > 0 && ntaps <= 32 && nz > 0 && nkb > 0 && n_rows >= cout,
               CCDM_ERR_BAD_ARG, "pack_weights: bad sizes cout=%d cin=%d taps=%d nz=%d nkb=%d n_rows=%d", cout,
               cin_total, ntaps, nz, nkb, n_rows);
  CCDM_REQUIRE(kb0 >= 0 && kb0 + nkb <= nkb_total, CCDM_ERR_BAD_ARG, "pack_weights: K blocks [%d, %d) of %d", kb0,
               kb0 + nkb, nkb_total);
  const long long total = (long long)nz * n_rows * nkb * 64;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin_total, ntaps, reinterpret_cast<const int4*>(psched), nkb, n_rows, cin_gain, gain_mul,
      reinterpret_cast<__nv_bfloat16*>(wpacked), total, nkb_total, kb0);
  return after_launch("pack_weights_kernel");
}

extern "C" int ccdm_pack_weights(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                                 int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain, float gain_mul,
                                 void* wpacked, void* stream) {
  return ccdm_pack_weights_at(w, cout, cin_total, ntaps, psched, nz, nkb, n_rows, cin_gain, gain_mul, wpacked, nkb, 0, stream);
}
