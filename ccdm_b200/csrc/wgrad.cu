// Weight gradient of the tap-GEMM convolutions on tensor cores (tcgen05), the backward of every nn.Conv2d of
// CCDM_unified/models/unet.py (autograd of :77,81,139,160,165,195,198,225,226,326,341):
//
//   dW[co][ci][tap] = sum over output positions p of dZ[p][co] * X[p + shift(tap)][ci]
//
// Per load group of the FORWARD schedule (source view, dw, dh0, 64-channel slice c0; R vertically adjacent taps) these are
// R GEMMs over the position axis that share their operands.  The ACTIVATIONS are the A operand and the taps are stacked
// along M (two taps per instruction), the output gradient is the B operand:
//
//   D_p[(tap t0 | tap t0+1) x 64 ci][n co] = [X_t0 | X_t0+1]^T[128 x 128 pos] . dZ[128 pos x n co]     n = Cout tile <= 128
//
// Activations are position-major (channels contiguous), so both operands are "MN-major" for UMMA: a TMA box of
// {64 channels, tw, th, tb} lands as one 128-byte row per position in the SWIZZLE_128B layout.  The X box carries
// R-1 extra rows, exactly like the forward kernel: tap r starts r*tw rows (r*tw*128 bytes, a multiple of the 1024-byte
// swizzle atom) into it -- which is both the start offset of the A descriptor and its leading-dimension byte offset, i.e.
// rows 64..127 of A are the NEXT tap's view of the same box.  R = 3 issues the pairs (0,1) and (1,2) and drops the
// duplicate; R = 1 wastes half of M.  (The first version had dZ as A with M = 128 output channels and one N = 64
// instruction per tap: 3 instructions of 6 KB of shared-memory operand reads per K step against 2 of 6.5 KB here, and the
// shared-memory port, not the tensor pipe, is what bounds these shapes -- profiles/r2_notes.md.)
// Out-of-image positions are zero-filled by TMA (= the convolution's zero padding), out-of-range channels likewise.
//
// Work split: blockIdx.y = (z, group, 128-channel output tile), blockIdx.x = slice of the position tiles (split-K).
// Each CTA accumulates its slice in TMEM and adds it into the fp32 gradient with reductions (coalesced over ci).  The gradient is
// produced in the PACKED layout of the forward weights ([z][row][kb*64 + j], kb = group*R + r); ccdm_unpack_wgrad
// maps it back to [Cout][Cin][kh*kw] (summing the folded taps of the nearest-2x upsampling convolution).
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box);

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 4;
constexpr uint32_t kWgABlock = 128 * 128;        // one {64 ch x 128 pos} box of dZ

struct WgAux {
  uint64_t full[kWgMaxStages], empty[kWgMaxStages], acc_full;
  uint32_t tmem_slot;
};

struct WgMaps {
  CUtensorMap src[CCDM_MAX_SRC];
  CUtensorMap dz[CCDM_MAX_Z];
};

struct WgDev {
  const int4* sched;
  float* out;
  int ngroups, R, nkb, m_tiles, n_rows;
  int tw, th, tb, tiles_w, tiles_h, tiles_m, tiles_per_cta;
  int stages, N, partial;
  long long slot_stride;
  uint32_t b_bytes, stage_bytes;
};

// First tap of pair p: R = 3 -> (0,1), (1,2); R = 2 -> (0,1); R = 1 -> (0, unused).
__host__ __device__ constexpr int wg_pair_t0(int R, int p) { return R >= 2 ? (2 * p < R - 2 ? 2 * p : R - 2) : 0; }

// Tap pairs x 8 K-steps (16 positions each) of one position tile; descriptor low words hold (address >> 4) | LBO.
template <int kR>
__device__ __forceinline__ void wgrad_taps(uint32_t tmem_base, uint32_t x_lo, uint32_t dz_lo, uint32_t tap16,
                                           uint32_t idesc, uint32_t acc_first) {
  constexpr uint64_t hi = (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                          (static_cast<uint64_t>(2) << 61);
#pragma unroll
  for (int pr = 0; pr < (kR + 1) / 2; ++pr) {
    const uint32_t a_lo = x_lo + wg_pair_t0(kR, pr) * tap16;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16_ss(tmem_base + pr * 128, hi | (a_lo + k * 128), hi | (dz_lo + k * 128), idesc, k != 0 ? 1u : acc_first);
  }
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgDev p) {
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  uint8_t* ring = wg_smem;
  WgAux* aux = reinterpret_cast<WgAux*>(ring + (size_t)p.stages * p.stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int unit = blockIdx.y;
  const int mt = unit % p.m_tiles;
  const int g = (unit / p.m_tiles) % p.ngroups;
  const int z = unit / (p.m_tiles * p.ngroups);
  const int4 e = __ldg(&p.sched[z * p.ngroups + g]);               // {source view, dw, dh0, c0}
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(t0 + p.tiles_per_cta, p.tiles_m);
  const int n_here = min(128, p.N - mt * 128);                    // output channels of this tile
  const int ncols = (n_here + 15) & ~15;                           // N of the instruction

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.src[e.x]);
    tma_prefetch_desc(&maps.dz[z]);
  }
  if (warp == 1) tmem_alloc(&aux->tmem_slot, 256);
  if (tid == 64) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&aux->full[s], 1);
      mbar_init(&aux->empty[s], 1);
    }
    mbar_init(&aux->acc_full, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (single lane: no divisions, running
    // ring / phase counters, integer shared-memory addresses)
    int tx = t0 % p.tiles_w, ty = (t0 / p.tiles_w) % p.tiles_h, tz = t0 / (p.tiles_w * p.tiles_h);
    const uint32_t full_bar = smem_u32(&aux->full[0]), empty_bar = smem_u32(&aux->empty[0]);
    const uint32_t ring_a = smem_u32(ring);
    const int n_stages = p.stages;
    const CUtensorMap* const dzm = &maps.dz[z];
    const CUtensorMap* const srm = &maps.src[e.x];
    int s = 0;
    uint32_t ph = 0, st_a = ring_a;
    for (int t = t0; t < t1; ++t) {
      mbar_wait_a(empty_bar + 8 * s, ph ^ 1u);
      if (elect_one()) {
        const int cw = tx * p.tw, chh = ty * p.th, cb = tz * p.tb;
        mbar_arrive_expect_tx_a(full_bar + 8 * s, (ncols > 64 ? 2 : 1) * kWgABlock + p.b_bytes);
        tma_load_4d_a(dzm, full_bar + 8 * s, st_a, mt * 128, cw, chh, cb);
        if (ncols > 64) tma_load_4d_a(dzm, full_bar + 8 * s, st_a + kWgABlock, mt * 128 + 64, cw, chh, cb);
        tma_load_4d_a(srm, full_bar + 8 * s, st_a + 2 * kWgABlock, e.w, cw + e.y, chh + e.z, cb);
      }
      __syncwarp();
      if (++tx == p.tiles_w) { tx = 0; if (++ty == p.tiles_h) { ty = 0; ++tz; } }
      st_a += p.stage_bytes;
      if (++s == n_stages) { s = 0; ph ^= 1u; st_a = ring_a; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform loop, one elected lane).
    // The lane is latency-bound: descriptors advance by integer adds on a precomputed low word (address >> 4 | LBO),
    // the taps are unrolled per R, barriers are addressed as integers.
    const uint32_t idesc = umma_idesc_bf16_mn(128, static_cast<uint32_t>(ncols));
    const uint32_t full_bar = smem_u32(&aux->full[0]), empty_bar = smem_u32(&aux->empty[0]);
    const uint32_t stage16 = p.stage_bytes >> 4, tap16 = static_cast<uint32_t>(p.tw) * 8;
    const uint32_t dz_flags = static_cast<uint32_t>((kWgABlock >> 4) & 0x3FFF) << 16;       // B: LBO = 16 KiB block stride
    const uint32_t x_flags = (tap16 & 0x3FFF) << 16;                                        // A: LBO = one tap (tw rows)
    const uint32_t a_lo0 = ((smem_u32(ring) & 0x3FFFF) >> 4) | dz_flags;
    const uint32_t x_lo0 = (((smem_u32(ring) + 2 * kWgABlock) & 0x3FFFF) >> 4) | x_flags;
    const int n_stages = p.stages;
    int s = 0;
    uint32_t ph = 0, a_lo = a_lo0, x_lo = x_lo0;
    uint32_t first = 0;                                    // 0 for the first position tile: overwrite the accumulators
    for (int t = t0; t < t1; ++t) {
      mbar_wait_a(full_bar + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        if (p.R == 3) wgrad_taps<3>(tmem_base, x_lo, a_lo, tap16, idesc, first);
        else if (p.R == 2) wgrad_taps<2>(tmem_base, x_lo, a_lo, tap16, idesc, first);
        else wgrad_taps<1>(tmem_base, x_lo, a_lo, tap16, idesc, first);
        umma_commit_a(empty_bar + 8 * s);
      }
      __syncwarp();
      first = 1;
      a_lo += stage16;
      x_lo += stage16;
      if (++s == n_stages) { s = 0; ph ^= 1u; a_lo = a_lo0; x_lo = x_lo0; }
    }
    if (t1 > t0) {
      if (elect_one()) umma_commit(&aux->acc_full);
      __syncwarp();
    }
  } else if (t1 > t0 || p.partial) {
    // ---------------------------------------------------------------- epilogue: TMEM lane = (tap of the pair, ci), column = co
    // partial mode: plain stores into this slice's slot (zeros when the slice has no positions); else red.global.add --
    // whose throughput at L2 (one fp32 per clock and slice), not the tensor pipe, bounded the split-K layers: ~150 CTAs x
    // 37 k elements onto the same 37 k addresses cost 30-40 us of a 67 us launch.
    const int q = warp & 3;
    const int half = q >> 1, ci = (q & 1) * 32 + lane;
    const bool have = t1 > t0;
    if (have) {
      mbar_wait(&aux->acc_full, 0);
      tc_fence_after();
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const size_t row_ld = (size_t)p.nkb * 64;
    float* const out = p.out + (p.partial ? (size_t)blockIdx.x * (size_t)p.slot_stride : 0);
    for (int pr = 0; pr < (p.R + 1) / 2; ++pr) {
      const int tap = wg_pair_t0(p.R, pr) + half;
      if (tap >= p.R || tap < 2 * pr) continue;                    // unused half (R = 1) or the duplicate of pair (1,2)
      float* obase = out + ((size_t)z * p.n_rows + (size_t)mt * 128) * row_ld + (size_t)(g * p.R + tap) * 64 + ci;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        if (have) {
          tmem_ld32(trow + pr * 128 + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (p.partial) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_here) obase[(size_t)(c0 + i) * row_ld] = __uint_as_float(v[i]);           // 128 B per warp
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_here) atomicAdd(obase + (size_t)(c0 + i) * row_ld, __uint_as_float(v[i]));   // RED
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 layers on grids of 8 x 16 tiles: ALL nine taps of a 64-channel block out of ONE halo box.
//
// The kernel above is bound by L2 -> SM traffic, not by the tensor pipe: every (channel block, dw) group is its own CTA
// unit and re-reads the whole dZ tensor and a shifted copy of X (72 -> 72 at 64x64, batch 128: 6 units x 150 MB = 900 MB
// through L2 per launch, 189 us).  Here a unit is (channel block, <= 96 output channels): per 8 x 16 position tile ONE dZ
// box and ONE X box {64 ch, 10, 18} are loaded and the nine taps are addressed inside it.  SWIZZLE_128B is a pure function
// of the shared-memory address, so an operand may start at any 128-byte row (profiles/r2_ubench_shift_desc.txt; the
// forward's halo plan relies on the same fact): a K step is 16 positions = two image rows of 8 pixels, i.e. two 8-row
// groups 1280 bytes apart in the box (SBO = 1280); tap (r, q) starts (r*10 + q) rows into it.  Taps are stacked two per
// instruction along M exactly as above (LBO = the byte distance between the two taps' views):
//   pair 0..2: taps (0,q) | (1,q)  LBO = 10 rows;   pair 3: (2,0) | (2,1)  LBO = 1 row;   pair 4: (2,2) | unused.
// Five accumulators of <= 96 columns (480 of the 512 TMEM columns); packed K block of tap (r, q) = g*9 + r*3 + q, the
// layout of the forward's halo plan (plan.py: plan_conv(halo=True)), so ccdm_unpack_wgrad takes that plan's schedule.
constexpr uint32_t kWhXBytes = 10 * 18 * 128;             // halo box
constexpr uint32_t kWhXSlot = 23 * 1024;                  // + 512 bytes the unused half of pair 4 may read past the box
constexpr int kWhAccStride = 96;

struct WhDev {
  const int4* sched;
  float* out;
  int ngroups, nkb, n_tiles, n_tile, n_rows, N;
  int tiles_w, tiles_h, tiles_m, tiles_per_cta;
  int stages, dz_blocks, partial;
  long long slot_stride;
  uint32_t stage_bytes;
};

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgMaps maps, const WhDev p) {
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  uint8_t* ring = wg_smem;
  WgAux* aux = reinterpret_cast<WgAux*>(ring + (size_t)p.stages * p.stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int unit = blockIdx.y;
  const int nt = unit % p.n_tiles;
  const int g = (unit / p.n_tiles) % p.ngroups;
  const int z = unit / (p.n_tiles * p.ngroups);
  const int4 e = __ldg(&p.sched[z * p.ngroups + g]);               // {source view, -1, -1, c0}
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(t0 + p.tiles_per_cta, p.tiles_m);
  const int n0 = nt * p.n_tile;
  const int n_here = min(p.n_tile, p.N - n0);                      // output channels of this unit
  const int ncols = (n_here + 15) & ~15;                           // N of the instruction
  const uint32_t x_off = static_cast<uint32_t>(p.dz_blocks) * kWgABlock;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.src[e.x]);
    tma_prefetch_desc(&maps.dz[z]);
  }
  if (warp == 1) tmem_alloc(&aux->tmem_slot, 512);
  if (tid == 64) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&aux->full[s], 1);
      mbar_init(&aux->empty[s], 1);
    }
    mbar_init(&aux->acc_full, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int tx = t0 % p.tiles_w, ty = (t0 / p.tiles_w) % p.tiles_h, tz = t0 / (p.tiles_w * p.tiles_h);
    const uint32_t full_bar = smem_u32(&aux->full[0]), empty_bar = smem_u32(&aux->empty[0]);
    const uint32_t ring_a = smem_u32(ring);
    const int n_stages = p.stages;
    const CUtensorMap* const dzm = &maps.dz[z];
    const CUtensorMap* const srm = &maps.src[e.x];
    const uint32_t tx_bytes = (ncols > 64 ? 2u : 1u) * kWgABlock + kWhXBytes;
    int s = 0;
    uint32_t ph = 0, st_a = ring_a;
    for (int t = t0; t < t1; ++t) {
      mbar_wait_a(empty_bar + 8 * s, ph ^ 1u);
      if (elect_one()) {
        const int cw = tx * 8, chh = ty * 16;
        mbar_arrive_expect_tx_a(full_bar + 8 * s, tx_bytes);
        tma_load_4d_a(dzm, full_bar + 8 * s, st_a, n0, cw, chh, tz);
        if (ncols > 64) tma_load_4d_a(dzm, full_bar + 8 * s, st_a + kWgABlock, n0 + 64, cw, chh, tz);
        tma_load_4d_a(srm, full_bar + 8 * s, st_a + x_off, e.w, cw + e.y, chh + e.z, tz);
      }
      __syncwarp();
      if (++tx == p.tiles_w) { tx = 0; if (++ty == p.tiles_h) { ty = 0; ++tz; } }
      st_a += p.stage_bytes;
      if (++s == n_stages) { s = 0; ph ^= 1u; st_a = ring_a; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: 5 tap pairs x 8 K steps per tile
    const uint32_t idesc = umma_idesc_bf16_mn(128, static_cast<uint32_t>(ncols));
    const uint32_t full_bar = smem_u32(&aux->full[0]), empty_bar = smem_u32(&aux->empty[0]);
    constexpr uint64_t hi_x = (static_cast<uint64_t>(1280 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                              (static_cast<uint64_t>(2) << 61);
    constexpr uint64_t hi_dz = (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                               (static_cast<uint64_t>(2) << 61);
    const uint32_t stage16 = p.stage_bytes >> 4;
    const uint32_t dz_lo0 = ((smem_u32(ring) & 0x3FFFF) >> 4) | (static_cast<uint32_t>((kWgABlock >> 4) & 0x3FFF) << 16);
    const uint32_t x_lo0 = ((smem_u32(ring) + x_off) & 0x3FFFF) >> 4;
    const int n_stages = p.stages;
    int s = 0;
    uint32_t ph = 0, dz_lo = dz_lo0, x_lo = x_lo0;
    uint32_t first = 0;                                    // 0 for the first position tile: overwrite the accumulators
    for (int t = t0; t < t1; ++t) {
      mbar_wait_a(full_bar + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll 1
        for (int pr = 0; pr < 5; ++pr) {
          // start row of the pair's first tap in the box (8 sixteen-byte units per row) | LBO between its two taps
          const uint32_t row0 = pr < 3 ? static_cast<uint32_t>(pr) : (pr == 3 ? 20u : 22u);
          const uint32_t a_lo = x_lo + row0 * 8 + ((pr < 3 ? 80u : 8u) << 16);
          const uint32_t d = tmem_base + pr * kWhAccStride;
#pragma unroll
          for (int k = 0; k < 8; ++k)                     // K step k = image rows 2k, 2k+1 of the tile: 20 box rows on
            umma_bf16_ss(d, hi_x | (a_lo + k * 160), hi_dz | (dz_lo + k * 128), idesc, k != 0 ? 1u : first);
        }
        umma_commit_a(empty_bar + 8 * s);
      }
      __syncwarp();
      first = 1;
      dz_lo += stage16;
      x_lo += stage16;
      if (++s == n_stages) { s = 0; ph ^= 1u; dz_lo = dz_lo0; x_lo = x_lo0; }
    }
    if (t1 > t0) {
      if (elect_one()) umma_commit(&aux->acc_full);
      __syncwarp();
    }
  } else if (t1 > t0 || p.partial) {
    // ---------------------------------------------------------------- epilogue: TMEM lane = (tap of the pair, ci), column = co
    // (partial mode / red.add: see conv_wgrad_kernel)
    const int q = warp & 3;
    const int half = q >> 1, ci = (q & 1) * 32 + lane;
    const bool have = t1 > t0;
    if (have) {
      mbar_wait(&aux->acc_full, 0);
      tc_fence_after();
    }
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const size_t row_ld = (size_t)p.nkb * 64;
    float* const out = p.out + (p.partial ? (size_t)blockIdx.x * (size_t)p.slot_stride : 0);
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = pr < 3 ? pr + 3 * half : (pr == 3 ? 6 + half : (half == 0 ? 8 : -1));
      if (tap < 0) continue;
      float* obase = out + ((size_t)z * p.n_rows + (size_t)n0) * row_ld + (size_t)(g * 9 + tap) * 64 + ci;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        if (have) {
          tmem_ld32(trow + pr * kWhAccStride + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (p.partial) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_here) obase[(size_t)(c0 + i) * row_ld] = __uint_as_float(v[i]);           // 128 B per warp
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_here) atomicAdd(obase + (size_t)(c0 + i) * row_ld, __uint_as_float(v[i]));   // RED
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[n][cin0+j][t] += gain * packed[z][n][kb*64 + j] for every tap t in the block's tapmask (scatter; dW zeroed or
// accumulated by the caller -- the folded taps of the nearest-2x convolution land on the same element from several
// blocks, hence atomics)
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, int nslots, long long slot_stride,
                                    float* __restrict__ dw, int cout, int cin_total, int ntaps,
                                    const int4* __restrict__ psched, int nkb, int n_rows,
                                    const float* __restrict__ cin_gain, float gain_mul, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 63);
    const long long t = idx >> 6;
    const int kb = (int)(t % nkb);
    const long long zn = t / nkb;
    const int n = (int)(zn % n_rows);
    const int z = (int)(zn / n_rows);
    const int4 e = __ldg(&psched[z * nkb + kb]);
    if (n >= cout || j >= e.y) continue;
    const int ci = e.x + j;
    float acc = 0.f;
    for (int sl = 0; sl < nslots; ++sl) acc += packed[(long long)sl * slot_stride + idx];   // fixed order: reproducible
    const float v = acc * gain_mul * (cin_gain ? cin_gain[ci] : 1.f);
    float* dst = dw + ((long long)n * cin_total + ci) * ntaps;
    unsigned mask = (unsigned)e.z;
    while (mask) {
      const int tp = __ffs(mask) - 1;
      mask &= mask - 1;
      atomicAdd(dst + tp, v);
    }
  }
}

// transposed packing for the data gradient: block kb of sub-problem z holds, for j < nvalid and n < n_count,
//   sum over taps t in tapmask of W[cin0 + j][n_off + n][t]      (W is the forward [Cout][Cin_total][taps] tensor)
__global__ void pack_weights_t_kernel(const float* __restrict__ w, int cin_total, int ntaps,
                                      const int4* __restrict__ psched, int nkb, int n_rows, int n_off, int n_count,
                                      __nv_bfloat16* __restrict__ out, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 63);
    const long long t = idx >> 6;
    const int kb = (int)(t % nkb);
    const long long zn = t / nkb;
    const int n = (int)(zn % n_rows);
    const int z = (int)(zn / n_rows);
    const int4 e = psched[z * nkb + kb];
    float acc = 0.f;
    if (n < n_count && j < e.y) {
      const float* wp = w + ((long long)(e.x + j) * cin_total + (n_off + n)) * ntaps;
      const unsigned mask = (unsigned)e.z;
      for (int tp = 0; tp < ntaps; ++tp)
        if (mask & (1u << tp)) acc += wp[tp];
    }
    out[idx] = __float2bfloat16(acc);
  }
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_conv_wgrad(const ccdm_wgrad_args* a, void* stream) {
  CCDM_REQUIRE(a != nullptr, CCDM_ERR_BAD_ARG, "conv_wgrad: null args");
  CCDM_REQUIRE(a->n_src >= 1 && a->n_src <= CCDM_MAX_SRC, CCDM_ERR_BAD_ARG, "conv_wgrad: n_src=%d", a->n_src);
  CCDM_REQUIRE(a->tw > 0 && a->th > 0 && a->tb > 0 && a->tw * a->th * a->tb == 128, CCDM_ERR_BAD_ARG,
               "conv_wgrad: tile box %dx%dx%d must hold 128 positions", a->tw, a->th, a->tb);
  CCDM_REQUIRE(a->nz >= 1 && a->nz <= CCDM_MAX_Z && a->ngroups >= 1 && a->R >= 1 && (a->R <= 3 || a->R == 9), CCDM_ERR_BAD_ARG,
               "conv_wgrad: nz=%d ngroups=%d R=%d", a->nz, a->ngroups, a->R);
  CCDM_REQUIRE(a->R == 1 || a->R == 9 || (a->tb == 1 && a->tw % 8 == 0), CCDM_ERR_BAD_ARG,
               "conv_wgrad: vertical tap reuse (R=%d) needs tb == 1 and tw %% 8 == 0", a->R);
  CCDM_REQUIRE(a->sched && a->dz && a->wgrad_packed, CCDM_ERR_BAD_ARG, "conv_wgrad: null sched/dz/wgrad_packed");
  CCDM_REQUIRE(a->N >= 1 && a->N % 8 == 0 && a->n_rows >= a->N, CCDM_ERR_UNSUPPORTED_SHAPE, "conv_wgrad: N=%d n_rows=%d",
               a->N, a->n_rows);
  CCDM_REQUIRE((reinterpret_cast<uintptr_t>(a->wgrad_packed) & 15) == 0, CCDM_ERR_BAD_ARG, "conv_wgrad: output alignment");
  CCDM_REQUIRE(a->slots >= 0 && a->slots <= 65535 &&
                   (a->slots == 0 || a->slot_stride >= (int64_t)a->nz * a->n_rows * a->ngroups * a->R * 64),
               CCDM_ERR_BAD_ARG, "conv_wgrad: slots=%d slot_stride=%lld", a->slots, (long long)a->slot_stride);
  const bool halo = a->R == 9;                                     // plan_conv(halo=True): one box, nine taps
  CCDM_REQUIRE(a->R <= 3 || (halo && a->tw == 8 && a->th == 16 && a->tb == 1 && a->nz == 1), CCDM_ERR_BAD_ARG,
               "conv_wgrad: R=%d needs the 8 x 16 x 1 halo tile and nz == 1", a->R);
  const int box_h = halo ? 18 : a->th + a->R - 1;
  const int box_w = halo ? 10 : a->tw;
  WgMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < CCDM_MAX_SRC; ++i) {
    const ccdm_view& v = a->src[i < a->n_src ? i : 0];
    CCDM_REQUIRE(v.ptr && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, CCDM_ERR_BAD_ARG,
                 "conv_wgrad: source %d pointer must be 16-byte aligned", i);
    CCDM_REQUIRE(v.C > 0 && v.W > 0 && v.H > 0 && v.B > 0 && v.sW % 8 == 0 && v.sH % 8 == 0 && v.sB % 8 == 0,
                 CCDM_ERR_BAD_ARG, "conv_wgrad: source %d extents/strides", i);
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.B};
    cuuint64_t str[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sB * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)a->tb};
    int rc = encode_map_bf16(&maps.src[i], v.ptr, 4, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  CCDM_REQUIRE((reinterpret_cast<uintptr_t>(a->dz) & 15) == 0 && a->dsW % 8 == 0 && a->dsH % 8 == 0 && a->dsB % 8 == 0,
               CCDM_ERR_BAD_ARG, "conv_wgrad: dz view must be 16-byte aligned with strides in multiples of 8");
  for (int zz = 0; zz < CCDM_MAX_Z; ++zz) {
    const int zi = zz < a->nz ? zz : 0;
    CCDM_REQUIRE(a->doff[zi] % 8 == 0, CCDM_ERR_BAD_ARG, "conv_wgrad: doff[%d] must be a multiple of 8 elements", zi);
    cuuint64_t dims[4] = {(cuuint64_t)a->N, (cuuint64_t)a->gW, (cuuint64_t)a->gH, (cuuint64_t)a->gB};
    cuuint64_t str[3] = {(cuuint64_t)a->dsW * 2, (cuuint64_t)a->dsH * 2, (cuuint64_t)a->dsB * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->tb};
    int rc = encode_map_bf16(&maps.dz[zz], reinterpret_cast<const __nv_bfloat16*>(a->dz) + a->doff[zi], 4, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  if (halo) {
    WhDev h;
    std::memset(&h, 0, sizeof(h));
    h.sched = reinterpret_cast<const int4*>(a->sched);
    h.out = a->wgrad_packed;
    h.ngroups = a->ngroups; h.nkb = a->ngroups * 9;
    h.n_tiles = (a->N + kWhAccStride - 1) / kWhAccStride;          // <= 96 output channels per unit, evenly split
    h.n_tile = (((a->N + h.n_tiles - 1) / h.n_tiles) + 15) & ~15;
    h.n_tiles = (a->N + h.n_tile - 1) / h.n_tile;
    h.n_rows = a->n_rows; h.N = a->N;
    h.tiles_w = (a->gW + 7) / 8;
    h.tiles_h = (a->gH + 15) / 16;
    h.tiles_m = h.tiles_w * h.tiles_h * a->gB;
    const int units = a->ngroups * h.n_tiles;
    int ksplit = a->slots > 0 ? a->slots : (a->ksplit > 0 ? a->ksplit : num_sms() / units);
    if (ksplit < 1) ksplit = 1;
    if (ksplit > h.tiles_m && a->slots == 0) ksplit = h.tiles_m;
    h.tiles_per_cta = (h.tiles_m + ksplit - 1) / ksplit;
    if (a->slots == 0) ksplit = (h.tiles_m + h.tiles_per_cta - 1) / h.tiles_per_cta;
    h.partial = a->slots > 0;
    h.slot_stride = a->slot_stride;
    h.dz_blocks = h.n_tile > 64 ? 2 : 1;
    h.stage_bytes = (uint32_t)h.dz_blocks * kWgABlock + kWhXSlot;
    const int max_stages = h.dz_blocks == 2 ? 3 : 4;
    h.stages = h.tiles_per_cta < max_stages ? (h.tiles_per_cta < 2 ? 2 : h.tiles_per_cta) : max_stages;
    const size_t smem = (size_t)h.stages * h.stage_bytes + sizeof(WgAux) + 1024;
    CCDM_REQUIRE(smem <= 227 * 1024, CCDM_ERR_UNSUPPORTED_SHAPE, "conv_wgrad(halo): %zu bytes of shared memory", smem);
    static size_t attr_smem_h = 0;
    if (smem > attr_smem_h) {
      cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad(halo): cudaFuncSetAttribute");
      attr_smem_h = smem;
    }
    conv_wgrad_halo_kernel<<<dim3((unsigned)ksplit, (unsigned)units), kWgThreads, smem, (cudaStream_t)stream>>>(maps, h);
    return after_launch("conv_wgrad_halo_kernel");
  }
  WgDev p;
  std::memset(&p, 0, sizeof(p));
  p.sched = reinterpret_cast<const int4*>(a->sched);
  p.out = a->wgrad_packed;
  p.ngroups = a->ngroups; p.R = a->R; p.nkb = a->ngroups * a->R;
  p.m_tiles = (a->N + 127) / 128;
  p.n_rows = a->n_rows;
  p.tw = a->tw; p.th = a->th; p.tb = a->tb;
  p.tiles_w = (a->gW + a->tw - 1) / a->tw;
  p.tiles_h = (a->gH + a->th - 1) / a->th;
  p.tiles_m = p.tiles_w * p.tiles_h * ((a->gB + a->tb - 1) / a->tb);
  const int units = a->nz * a->ngroups * p.m_tiles;
  int ksplit = a->slots > 0 ? a->slots : (a->ksplit > 0 ? a->ksplit : num_sms() / units);   // about one wave of CTAs
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.tiles_m && a->slots == 0) ksplit = p.tiles_m;
  p.tiles_per_cta = (p.tiles_m + ksplit - 1) / ksplit;
  if (a->slots == 0) ksplit = (p.tiles_m + p.tiles_per_cta - 1) / p.tiles_per_cta;    // no CTA without work (partial mode:
  p.partial = a->slots > 0;                                                            // such a slice stores zeros)
  p.slot_stride = a->slot_stride;
  p.N = a->N;
  p.b_bytes = (uint32_t)(box_h * a->tw * a->tb) * 128u;
  p.stage_bytes = 2 * kWgABlock + p.b_bytes;
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  p.stages = p.tiles_per_cta < kWgMaxStages ? (p.tiles_per_cta < 2 ? 2 : p.tiles_per_cta) : kWgMaxStages;
  // R = 1: rows 64..127 of the A operand (the unused "next tap") reach up to tw rows past the box of the last stage
  const size_t smem = (size_t)p.stages * p.stage_bytes + sizeof(WgAux) + 1024 + (a->R == 1 ? (size_t)a->tw * 128 : 0);
  CCDM_REQUIRE(smem <= 227 * 1024, CCDM_ERR_UNSUPPORTED_SHAPE, "conv_wgrad: %zu bytes of shared memory", smem);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_wgrad: cudaFuncSetAttribute");
    attr_smem = smem;
  }
  conv_wgrad_kernel<<<dim3((unsigned)ksplit, (unsigned)units), kWgThreads, smem, (cudaStream_t)stream>>>(maps, p);
  return after_launch("conv_wgrad_kernel");
}

extern "C" int ccdm_unpack_wgrad_slots(const float* packed, int32_t nslots, int64_t slot_stride, float* dw, int32_t cout,
                                       int32_t cin_total, int32_t ntaps, const int32_t* psched, int32_t nz, int32_t nkb,
                                       int32_t n_rows, const float* cin_gain, float gain_mul, int32_t accumulate,
                                       void* stream) {
  CCDM_REQUIRE(packed && dw && psched, CCDM_ERR_BAD_ARG, "unpack_wgrad: null pointer");
  CCDM_REQUIRE(cout > 0 && cin_total > 0 && ntaps > 0 && ntaps <= 32 && nz > 0 && nkb > 0 && n_rows >= cout,
               CCDM_ERR_BAD_ARG, "unpack_wgrad: bad sizes");
  const long long total = (long long)nz * n_rows * nkb * 64;
  CCDM_REQUIRE(nslots >= 1 && (nslots == 1 || slot_stride >= total), CCDM_ERR_BAD_ARG,
               "unpack_wgrad: nslots=%d slot_stride=%lld", nslots, (long long)slot_stride);
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)cout * cin_total * ntaps * sizeof(float), s);
    if (e != cudaSuccess) return cuda_fail(e, "unpack_wgrad: cudaMemsetAsync");
  }
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  unpack_wgrad_kernel<<<(unsigned)blocks, 256, 0, s>>>(packed, nslots, (long long)slot_stride, dw, cout, cin_total, ntaps,
                                                       reinterpret_cast<const int4*>(psched), nkb, n_rows, cin_gain,
                                                       gain_mul, total);
  return after_launch("unpack_wgrad_kernel");
}

extern "C" int ccdm_unpack_wgrad(const float* packed, float* dw, int32_t cout, int32_t cin_total, int32_t ntaps,
                                 const int32_t* psched, int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain,
                                 float gain_mul, int32_t accumulate, void* stream) {
  return ccdm_unpack_wgrad_slots(packed, 1, 0, dw, cout, cin_total, ntaps, psched, nz, nkb, n_rows, cin_gain, gain_mul,
                                 accumulate, stream);
}

extern "C" int ccdm_pack_weights_t(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                                   int32_t nz, int32_t nkb, int32_t n_rows, int32_t n_off, int32_t n_count, void* wpacked,
                                   void* stream) {
  CCDM_REQUIRE(w && psched && wpacked, CCDM_ERR_BAD_ARG, "pack_weights_t: null pointer");
  CCDM_REQUIRE(cout > 0 && cin_total > 0 && ntaps > 0 && ntaps <= 32 && nz > 0 && nkb > 0 && n_off >= 0 && n_count > 0 &&
                   n_off + n_count <= cin_total && n_rows >= n_count,
               CCDM_ERR_BAD_ARG, "pack_weights_t: bad sizes cout=%d cin=%d taps=%d n_off=%d n_count=%d n_rows=%d", cout,
               cin_total, ntaps, n_off, n_count, n_rows);
  const long long total = (long long)nz * n_rows * nkb * 64;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_t_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      w, cin_total, ntaps, reinterpret_cast<const int4*>(psched), nkb, n_rows, n_off, n_count,
      reinterpret_cast<__nv_bfloat16*>(wpacked), total);
  return after_launch("pack_weights_t_kernel");
}
