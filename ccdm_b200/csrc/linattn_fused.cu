// Fused linear attention for inference (unet.py:202-216 inside Residual(PreNorm(.)), :66-72,:92-99), sm_100a.
//
// The unfused path writes q | p | v (3 x 128 channels per token, 1.26 GB per launch at 64x64 / batch 400) to HBM and reads
// it back twice.  Here the token tensor x is read twice and nothing but the output is written:
//
//   kernel 1  linattn_kv_kernel     x -> [k | v] = x W'_kv (tcgen05, K-major operands) -> p = exp(k/|x| - bound), v/|x|
//                                   (epilogue, bf16 into shared memory in the token-major layout a TMA box would have)
//                                   -> D2 += P^T V, S += P^T 1 (tcgen05, MN-major operands) accumulated in TMEM over a
//                                   UNIT of G consecutive 128-token tiles of one sample -> partial context per unit.
//   kernel 2  linattn_fold_parts    context = (sum of the sample's unit partials, FIXED order) / S;
//                                   wfold[b][c][h*32+d] = sum_e W_out[c][h*32+e] context[b][h][d][e]   (bf16)
//   kernel 3  linattn_qout_kernel   x -> q = x W'_q -> softmax over the 32 channels of each head * scale (epilogue, bf16 into
//                                   shared memory as a K-major A operand) -> y = q wfold_b^T (tcgen05) -> + bias ->
//                                   channel RMSNorm * g * sqrt(C) -> + x (the residual is the A tile still in shared
//                                   memory) -> bf16 -> TMA store.
//
// Units never straddle CTAs and their partials are summed in a fixed order, so the result does not depend on how many
// samples are in the batch (batch-shard invariance, bit-exact) -- no atomics anywhere.
//
// Warp roles (both tensor kernels): warp 0 TMA producer, warp 1 issues the first GEMM (x W^T), warps 2..17 epilogue, warp 18
// issues the second GEMM (context / output projection): the single issuing lane is a serial chain of uniform-datapath
// instructions (~100 cycles per MMA), so the two GEMMs get one issuing warp each.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box);

constexpr int kLaEpiWarps = 16;                 // four warps per TMEM lane quarter: latency-bound epilogues need the occupancy
constexpr int kLaThreads = 64 + 32 * kLaEpiWarps + 32;         // kv kernel: warp 0 TMA, warp 1 GEMM1, 16 epilogue warps, last warp GEMM2
constexpr int kLaMma2Warp = kLaThreads / 32 - 1;
constexpr int kLa2Threads = 64 + 32 * kLaEpiWarps;             // q-out kernel: one issuing warp (a second one measured no gain)
constexpr int kLaTok = 128;                     // tokens per tile = UMMA M
constexpr int kLaBlk = kLaTok * 128;            // one {64 channels x 128 tokens} shared-memory block: 16 KiB
constexpr int kLaMaxStages = 4;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// MN-major SWIZZLE_128B descriptor from a low word that already holds (address >> 4) | (LBO >> 4) << 16: advancing by k
// tokens' worth of bytes is one integer add on the low word; the high word (SBO = 1024 B, version, layout) is constant.
__device__ __forceinline__ uint64_t umma_desc_mn_lo(uint32_t lo) {
  constexpr uint64_t hi = (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                          (static_cast<uint64_t>(2) << 61);
  return hi | lo;
}
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

// ============================================================================ kernel 1: x -> per-unit context partials

struct La1Aux {
  uint64_t x_full[kLaMaxStages], x_empty[kLaMaxStages], w_full, d1_full[2], d1_empty[2], pv_full[2], pv_empty[2], d2_full,
      d2_empty;
  uint32_t tmem_slot, pad_[3];
  float kbias[128];                             // log2(e) * (-bound_d) of the 128 k channels
};

struct La1Params {
  int n, nkb, tps, G, ups, units_total, units_per_cta, x_stages, pv_bufs;
  const float* rowss;                           // [B*n] sum of squares of the token rows of x
  const float* kbias;                           // [384]: the k rows' softmax shifts at [128, 256)
  float* part;                                  // [B*ups][128 (h*32+d)][32 (e)]
  float* psum;                                  // [B*ups][128]
};

__global__ void __launch_bounds__(kLaThreads, 1) linattn_kv_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                   const __grid_constant__ CUtensorMap wmap,
                                                                   const La1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wres = smem;                                            // [kb][256 rows: k then v][128 B]
  uint8_t* xring = wres + static_cast<size_t>(p.nkb) * 2 * kLaBlk; // [stage][kb][128 tok][128 B]
  uint8_t* stg = xring + static_cast<size_t>(p.x_stages) * p.nkb * kLaBlk;   // [buf][P lo | P hi | V lo | V hi]
  uint8_t* ones = stg + static_cast<size_t>(p.pv_bufs) * 4 * kLaBlk;
  La1Aux* aux = reinterpret_cast<La1Aux*>(ones + kLaBlk);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&wmap);
  }
  if (warp == 1) tmem_alloc(&aux->tmem_slot, 512);
  if (tid == 64) {
    for (int s = 0; s < p.x_stages; ++s) {
      mbar_init(&aux->x_full[s], 1);
      mbar_init(&aux->x_empty[s], 1);
    }
    mbar_init(&aux->w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aux->d1_full[s], 1);
      mbar_init(&aux->d1_empty[s], kLaEpiWarps);
      mbar_init(&aux->pv_full[s], kLaEpiWarps);
      mbar_init(&aux->pv_empty[s], 1);
    }
    mbar_init(&aux->d2_full, 1);
    mbar_init(&aux->d2_empty, kLaEpiWarps);
    fence_mbar_init();
  }
  for (int i = tid; i < kLaBlk / 4; i += kLaThreads) reinterpret_cast<uint32_t*>(ones)[i] = CCDM_ONE_PAIR;   // bf16 1.0
  for (int i = tid; i < 128; i += kLaThreads) aux->kbias[i] = p.kbias[128 + i] * kLog2e;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;
  const uint32_t d2_col = 256, ds_col = 384;

  const int u_begin = blockIdx.x * p.units_per_cta;
  const int u_end = min(p.units_total, u_begin + p.units_per_cta);
  const int n_tiles = max(0, u_end - u_begin) * p.G;               // every unit has exactly G tiles
  const int x_stage_bytes = p.nkb * kLaBlk;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&aux->w_full, static_cast<uint32_t>(p.nkb) * 2 * kLaBlk);
      for (int kb = 0; kb < p.nkb; ++kb) {
        tma_load_2d(&wmap, &aux->w_full, wres + (kb * 2 + 0) * kLaBlk, kb * 64, 128);   // k rows of the qkv weight
        tma_load_2d(&wmap, &aux->w_full, wres + (kb * 2 + 1) * kLaBlk, kb * 64, 256);   // v rows
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int u = u_begin; u < u_end; ++u) {
      const int b = u / p.ups, tile0 = (u % p.ups) * p.G;
      for (int tl = 0; tl < p.G; ++tl) {
        mbar_wait(&aux->x_empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&aux->x_full[s], static_cast<uint32_t>(x_stage_bytes));
          for (int kb = 0; kb < p.nkb; ++kb)
            tma_load_3d(&xmap, &aux->x_full[s], xring + static_cast<size_t>(s) * x_stage_bytes + kb * kLaBlk, kb * 64,
                        (tile0 + tl) * kLaTok, b);
        }
        __syncwarp();
        if (++s == p.x_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ GEMM1 issuer: [k | v] = x W'^T, two 128-column items
    const uint32_t idesc1 = umma_idesc_bf16(128, 128);             // 128 tokens x 128 channels, both operands K-major
    mbar_wait(&aux->w_full, 0);
    tc_fence_after();
    const uint32_t w16 = (smem_u32(wres) & 0x3FFFF) >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int li = 0; li < n_tiles; ++li) {
      mbar_wait(&aux->x_full[s], ph);
      tc_fence_after();
      const uint32_t x16 = (smem_u32(xring + static_cast<size_t>(s) * x_stage_bytes) & 0x3FFFF) >> 4;
      for (int j = 0; j < 2; ++j) {
        mbar_wait(&aux->d1_empty[j], (static_cast<uint32_t>(li) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          for (int kb = 0; kb < p.nkb; ++kb) {
            const uint32_t a16 = x16 + kb * (kLaBlk >> 4);
            const uint32_t b16 = w16 + (kb * 2 + j) * (kLaBlk >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_base + j * 128, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k), idesc1,
                           (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&aux->d1_full[j]);
          if (j == 1) umma_commit(&aux->x_empty[s]);
        }
        __syncwarp();
      }
      if (++s == p.x_stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == kLaMma2Warp) {
    // ================================================================ GEMM2 issuer: D2 += P^T V, S += P^T 1 per tile
    const uint32_t idesc_ctx = umma_idesc_bf16_mn(128, 128);       // both operands token-major (MN-major)
    const uint32_t idesc_sum = umma_idesc_bf16_mn(128, 16);
    const uint32_t lbo = static_cast<uint32_t>(kLaBlk >> 4) << 16;
    const uint32_t ones_lo = ((smem_u32(ones) & 0x3FFFF) >> 4) | lbo;
    const uint32_t stg_lo = ((smem_u32(stg) & 0x3FFFF) >> 4) | lbo;
    int buf = 0, tl = 0;
    uint32_t use = 0, ul = 0;
    for (int t = 0; t < n_tiles; ++t) {
      if (tl == 0) {                                               // new unit: the previous unit's D2 must have been read
        mbar_wait(&aux->d2_empty, (ul & 1u) ^ 1u);
        tc_fence_after();
      }
      mbar_wait(&aux->pv_full[buf], use & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_lo = stg_lo + buf * (4 * kLaBlk >> 4);    // P^T: blocks 0, 1; V: blocks 2, 3 of the staging buffer
        const uint32_t v_lo = p_lo + (2 * kLaBlk >> 4);
#pragma unroll
        for (int k = 0; k < kLaTok / 16; ++k) {
          const uint32_t acc = (tl | k) != 0 ? 1u : 0u;
          const uint64_t adesc = umma_desc_mn_lo(p_lo + k * (2048 >> 4));
          umma_bf16_ss(tmem_base + d2_col, adesc, umma_desc_mn_lo(v_lo + k * (2048 >> 4)), idesc_ctx, acc);
          umma_bf16_ss(tmem_base + ds_col, adesc, umma_desc_mn_lo(ones_lo + k * (2048 >> 4)), idesc_sum, acc);
        }
        umma_commit(&aux->pv_empty[buf]);
        if (tl == p.G - 1) umma_commit(&aux->d2_full);
      }
      __syncwarp();
      if (++buf == p.pv_bufs) { buf = 0; ++use; }
      if (++tl == p.G) { tl = 0; ++ul; }
    }
  } else if (warp < 2 + kLaEpiWarps) {
    // ================================================================ epilogue (16 warps: 128 rows x 4 column quarters)
    const int ew = warp - 2;
    const int q = warp & 3;                                        // TMEM lane quarter this warp may read
    const int cq = ew >> 2;                                        // 32-channel quarter of every 128-channel item
    const int m = q * 32 + lane;                                   // token row inside the tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int blk = cq >> 1, cc = cq & 1;                          // 64-channel staging block and its 32-channel half
    uint32_t r[32];
    int li = 0, pend_u = -1;
    uint32_t flushes = 0;
    auto flush = [&](int u) {                                      // D2 / S of a finished unit -> global partials
      mbar_wait(&aux->d2_full, flushes & 1u);
      tc_fence_after();
      if (cq == 0) {
        tmem_ld32(trow + d2_col + q * 32, r);                      // head q: rows q*32+d, columns q*32 + e
        tmem_ld_wait();
        float4* o = reinterpret_cast<float4*>(p.part + (static_cast<long long>(u) * 128 + m) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
      } else if (cq == 1) {
        tmem_ld32(trow + ds_col, r);                               // every column of S holds sum_n p[n][q*32+d]
        tmem_ld_wait();
        p.psum[static_cast<long long>(u) * 128 + m] = __uint_as_float(r[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->d2_empty);
      ++flushes;
    };
    int buf = 0;
    uint32_t use = 0;                                              // how often the current staging buffer has been used
    const float* rsp = p.rowss + static_cast<long long>(u_begin) * p.G * kLaTok + m;   // units are contiguous token runs
    float rss_next = n_tiles > 0 ? __ldg(rsp) : 1.f;
    for (int u = u_begin; u < u_end; ++u) {
      for (int tl = 0; tl < p.G; ++tl, ++li) {
        const float rs = 1.f / fmaxf(sqrtf(rss_next), 1e-12f);
        rsp += kLaTok;
        if (li + 1 < n_tiles) rss_next = __ldg(rsp);               // next tile's row norm: its latency hides behind this tile
        uint8_t* sb = stg + static_cast<size_t>(buf) * 4 * kLaBlk + m * 128;
        mbar_wait(&aux->pv_empty[buf], (use & 1u) ^ 1u);           // the context MMAs that read this buffer are done
        // ---- item 0: p = exp(k / |x| - bound) = 2^(acc * rs*log2e + kbias)
        mbar_wait(&aux->d1_full[0], static_cast<uint32_t>(li) & 1u);
        tc_fence_after();
        {
          const float2 rs2 = make_float2(rs * kLog2e, rs * kLog2e);
          tmem_ld32(trow + cq * 32, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aux->d1_empty[0]);
          const float2* kb2 = reinterpret_cast<const float2*>(aux->kbias + cq * 32);
          uint8_t* dst = sb + (0 + blk) * kLaBlk;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float2 e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = make_float2(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1]));
              const float2 x2 = __ffma2_rn(a, rs2, kb2[g * 4 + i]);
              e[i] = make_float2(ex2_fast(x2.x), ex2_fast(x2.y));
            }
            uint4 o;
            o.x = pack_bf16(e[0].x, e[0].y);
            o.y = pack_bf16(e[1].x, e[1].y);
            o.z = pack_bf16(e[2].x, e[2].y);
            o.w = pack_bf16(e[3].x, e[3].y);
            *reinterpret_cast<uint4*>(dst + (((cc * 4 + g) ^ (m & 7)) << 4)) = o;
          }
        }
        // ---- item 1: v / |x|
        mbar_wait(&aux->d1_full[1], static_cast<uint32_t>(li) & 1u);
        tc_fence_after();
        {
          const float2 rs2 = make_float2(rs, rs);
          tmem_ld32(trow + 128 + cq * 32, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aux->d1_empty[1]);
          uint8_t* dst = sb + (2 + blk) * kLaBlk;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float2 e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              e[i] = __fmul2_rn(make_float2(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])), rs2);
            uint4 o;
            o.x = pack_bf16(e[0].x, e[0].y);
            o.y = pack_bf16(e[1].x, e[1].y);
            o.z = pack_bf16(e[2].x, e[2].y);
            o.w = pack_bf16(e[3].x, e[3].y);
            *reinterpret_cast<uint4*>(dst + (((cc * 4 + g) ^ (m & 7)) << 4)) = o;
          }
        }
        fence_proxy_async_smem();                                  // generic-proxy writes -> visible to the MMA unit
        __syncwarp();
        if (lane == 0) mbar_arrive(&aux->pv_full[buf]);
        if (++buf == p.pv_bufs) { buf = 0; ++use; }
        // the previous unit's accumulators are read one tile late: its last context MMAs have finished by now
        if (pend_u >= 0) {
          flush(pend_u);
          pend_u = -1;
        }
        if (tl == p.G - 1) pend_u = u;
      }
    }
    if (pend_u >= 0) flush(pend_u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ============================================================================ kernel 2: partials -> folded weights

// One CTA per sample, thread = (head h, channel d).  Sums the sample's unit partials in a fixed order, normalises by the
// column sum S (softmax over tokens) and folds the context into to_out[0]'s weight (unet.py:198,212-216).  W_out is staged
// in shared memory once per CTA (the first version read it through L1/L2 per output row and was latency-bound: 50-90 us);
// four output rows are accumulated at a time so the FMA chains overlap.
__global__ void __launch_bounds__(128) linattn_fold_parts_kernel(const float* __restrict__ part,
                                                                 const float* __restrict__ psum, int ups,
                                                                 const float* __restrict__ w_out, int C, int n_rows,
                                                                 __nv_bfloat16* __restrict__ wfold, int B) {
  extern __shared__ float fold_smem[];                             // [C][128] fp32 copy of W_out
  const int t = threadIdx.x;                                       // t = h*32 + d
  const int h = t >> 5;
  {
    const float4* src = reinterpret_cast<const float4*>(w_out);
    float4* dst = reinterpret_cast<float4*>(fold_smem);
    for (int i = t; i < C * 32; i += 128) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float cr[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) cr[e] = 0.f;
    float s = 0.f;
    for (int u = 0; u < ups; ++u) {
      const long long ub = static_cast<long long>(b) * ups + u;
      const float4* pr = reinterpret_cast<const float4*>(part + (ub * 128 + t) * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = __ldg(pr + j);
        cr[4 * j] += v.x; cr[4 * j + 1] += v.y; cr[4 * j + 2] += v.z; cr[4 * j + 3] += v.w;
      }
      s += __ldg(psum + ub * 128 + t);
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int e = 0; e < 32; ++e) cr[e] *= inv;
    __nv_bfloat16* wf = wfold + static_cast<long long>(b) * n_rows * 128 + t;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float4* w0 = reinterpret_cast<const float4*>(fold_smem + (c + 0) * 128 + h * 32);   // warp-uniform: broadcast reads
      const float4* w1 = reinterpret_cast<const float4*>(fold_smem + (c + 1) * 128 + h * 32);
      const float4* w2 = reinterpret_cast<const float4*>(fold_smem + (c + 2) * 128 + h * 32);
      const float4* w3 = reinterpret_cast<const float4*>(fold_smem + (c + 3) * 128 + h * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x0 = w0[j], x1 = w1[j], x2 = w2[j], x3 = w3[j];
        a0 = fmaf(x0.x, cr[4 * j], a0); a0 = fmaf(x0.y, cr[4 * j + 1], a0); a0 = fmaf(x0.z, cr[4 * j + 2], a0); a0 = fmaf(x0.w, cr[4 * j + 3], a0);
        a1 = fmaf(x1.x, cr[4 * j], a1); a1 = fmaf(x1.y, cr[4 * j + 1], a1); a1 = fmaf(x1.z, cr[4 * j + 2], a1); a1 = fmaf(x1.w, cr[4 * j + 3], a1);
        a2 = fmaf(x2.x, cr[4 * j], a2); a2 = fmaf(x2.y, cr[4 * j + 1], a2); a2 = fmaf(x2.z, cr[4 * j + 2], a2); a2 = fmaf(x2.w, cr[4 * j + 3], a2);
        a3 = fmaf(x3.x, cr[4 * j], a3); a3 = fmaf(x3.y, cr[4 * j + 1], a3); a3 = fmaf(x3.z, cr[4 * j + 2], a3); a3 = fmaf(x3.w, cr[4 * j + 3], a3);
      }
      wf[static_cast<long long>(c + 0) * 128] = __float2bfloat16(a0);
      wf[static_cast<long long>(c + 1) * 128] = __float2bfloat16(a1);
      wf[static_cast<long long>(c + 2) * 128] = __float2bfloat16(a2);
      wf[static_cast<long long>(c + 3) * 128] = __float2bfloat16(a3);
    }
    for (; c < C; ++c) {
      const float4* wr = reinterpret_cast<const float4*>(fold_smem + c * 128 + h * 32);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 wv = wr[j];
        acc = fmaf(wv.x, cr[4 * j], acc);
        acc = fmaf(wv.y, cr[4 * j + 1], acc);
        acc = fmaf(wv.z, cr[4 * j + 2], acc);
        acc = fmaf(wv.w, cr[4 * j + 3], acc);
      }
      wf[static_cast<long long>(c) * 128] = __float2bfloat16(acc);
    }
  }
}

// ============================================================================ kernel 3: x -> q -> out (+ norm + residual)

struct La2Aux {
  uint64_t x_full[kLaMaxStages], x_empty[kLaMaxStages], w_full, wf_full[2], wf_empty[2], d1_full[2], d1_empty[2], q_full[2],
      d2_full[2], d2_empty[2];
  uint32_t tmem_slot, pad_[3];
  float bias[128], gain[128];                   // to_out bias, g * sqrt(C) (0 for padded channels)
};

struct La2Params {
  int n, nkb, tps, tiles_total, tiles_per_cta, x_stages, wf_bufs, C, n_rows;
  const float* rowss;
  const float* bias;                            // to_out[0].bias [C]
  const float* gain;                            // to_out[1].g [C]
  float gain_mul, q_scale;
};

__global__ void __launch_bounds__(kLa2Threads, 1) linattn_qout_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                     const __grid_constant__ CUtensorMap wmap,
                                                                     const __grid_constant__ CUtensorMap fmap,
                                                                     const __grid_constant__ CUtensorMap omap,
                                                                     const La2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int wf_bytes = 2 * p.n_rows * 128;                         // two K blocks of [n_rows][128 B]
  const int wf_slot = (wf_bytes + 1023) & ~1023;
  uint8_t* wres = smem;                                            // [kb][128 q rows][128 B]
  uint8_t* xring = wres + static_cast<size_t>(p.nkb) * kLaBlk;     // [stage][kb][128 tok][128 B]
  uint8_t* wfr = xring + static_cast<size_t>(p.x_stages) * p.nkb * kLaBlk;    // [buf][2][n_rows][128 B]
  uint8_t* qstg = wfr + static_cast<size_t>(p.wf_bufs) * wf_slot;  // [group][2][128 tok][128 B]; reused as output staging
  La2Aux* aux = reinterpret_cast<La2Aux*>(qstg + 4 * kLaBlk);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&wmap);
    tma_prefetch_desc(&fmap);
    tma_prefetch_desc(&omap);
  }
  if (warp == 1) tmem_alloc(&aux->tmem_slot, 512);
  if (tid == 64) {
    for (int s = 0; s < p.x_stages; ++s) {
      mbar_init(&aux->x_full[s], 1);
      mbar_init(&aux->x_empty[s], 1 + kLaEpiWarps / 2);            // GEMM1 commit + the group's warps that read the residual
    }
    mbar_init(&aux->w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aux->wf_full[s], 1);
      mbar_init(&aux->wf_empty[s], 1);
      mbar_init(&aux->d1_full[s], 1);
      mbar_init(&aux->d1_empty[s], kLaEpiWarps / 2);
      mbar_init(&aux->q_full[s], kLaEpiWarps / 2);
      mbar_init(&aux->d2_full[s], 1);
      mbar_init(&aux->d2_empty[s], kLaEpiWarps / 2);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < 128; i += kLa2Threads) {
    aux->bias[i] = i < p.C ? p.bias[i] : 0.f;
    aux->gain[i] = i < p.C ? p.gain[i] * p.gain_mul : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;

  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_cta);
  const int n_tiles = max(0, t_end - t_begin);
  const int x_stage_bytes = p.nkb * kLaBlk;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&aux->w_full, static_cast<uint32_t>(p.nkb) * kLaBlk);
      for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(&wmap, &aux->w_full, wres + kb * kLaBlk, kb * 64, 0);   // q rows
    }
    __syncwarp();
    int s = 0, b_prev = -1, sl = -1;
    uint32_t ph = 0;
    for (int li = 0; li < n_tiles; ++li) {
      const int tile = t_begin + li, b = tile / p.tps, tok0 = (tile % p.tps) * kLaTok;
      mbar_wait(&aux->x_empty[s], ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&aux->x_full[s], static_cast<uint32_t>(x_stage_bytes));
        for (int kb = 0; kb < p.nkb; ++kb)
          tma_load_3d(&xmap, &aux->x_full[s], xring + static_cast<size_t>(s) * x_stage_bytes + kb * kLaBlk, kb * 64, tok0, b);
      }
      __syncwarp();
      if (++s == p.x_stages) { s = 0; ph ^= 1u; }
      if (b != b_prev) {                                           // new sample: its folded weights (after the x tile!)
        b_prev = b;
        ++sl;
        const int slot = sl % p.wf_bufs;
        mbar_wait(&aux->wf_empty[slot], (static_cast<uint32_t>(sl / p.wf_bufs) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&aux->wf_full[slot], static_cast<uint32_t>(wf_bytes));
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(&fmap, &aux->wf_full[slot], wfr + static_cast<size_t>(slot) * wf_slot + kb * p.n_rows * 128, kb * 64,
                        b * p.n_rows);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================================================ tcgen05 issuers.  GEMM1: q = x W'_q^T; GEMM2: y = q wfold_b^T.
    // Warp 1 interleaves GEMM1(t) and GEMM2(t-1).  (Giving GEMM2 its own issuing warp, as in the kv kernel, measured 233 vs
    // 241 us at 64x64 / batch 400: no gain, the epilogue is the limiter here.)
    const bool do1 = true, do2 = true;
    const uint32_t idesc1 = umma_idesc_bf16(128, 128);
    const uint32_t idesc2 = umma_idesc_bf16(128, static_cast<uint32_t>(p.n_rows));
    const uint32_t w16 = (smem_u32(wres) & 0x3FFFF) >> 4;
    const uint32_t q16 = (smem_u32(qstg) & 0x3FFFF) >> 4;
    const uint32_t f16 = (smem_u32(wfr) & 0x3FFFF) >> 4;
    if (do1) {
      mbar_wait(&aux->w_full, 0);
      tc_fence_after();
    }
    int b_prev = -1, sl = -1, slot = 0;
    int b = t_begin / p.tps, ti = t_begin % p.tps;                 // sample / tile-in-sample of the next GEMM2 tile
    auto gemm2 = [&](int t) {
      const int g = t & 1;
      const uint32_t k2 = static_cast<uint32_t>(t >> 1);
      const bool last_of_sample = (t == n_tiles - 1) || (ti == p.tps - 1);
      if (b != b_prev) {
        b_prev = b;
        ++sl;
        slot = sl % p.wf_bufs;
        mbar_wait(&aux->wf_full[slot], static_cast<uint32_t>(sl / p.wf_bufs) & 1u);
      }
      mbar_wait(&aux->q_full[g], k2 & 1u);
      mbar_wait(&aux->d2_empty[g], (k2 & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a16 = q16 + (g * 2 + kb) * (kLaBlk >> 4);
          const uint32_t b16 = f16 + ((slot * wf_slot + kb * p.n_rows * 128) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + 256 + g * 128, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k),
                         idesc2, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&aux->d2_full[g]);
        if (last_of_sample) umma_commit(&aux->wf_empty[slot]);
      }
      __syncwarp();
      if (++ti == p.tps) { ti = 0; ++b; }
    };
    int s = 0;
    uint32_t ph = 0;
    for (int li = 0; li < n_tiles; ++li) {
      if (do1) {
        const int g = li & 1;
        mbar_wait(&aux->x_full[s], ph);
        mbar_wait(&aux->d1_empty[g], (static_cast<uint32_t>(li >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t x16 = (smem_u32(xring + static_cast<size_t>(s) * x_stage_bytes) & 0x3FFFF) >> 4;
        if (elect_one()) {
          for (int kb = 0; kb < p.nkb; ++kb) {
            const uint32_t a16 = x16 + kb * (kLaBlk >> 4);
            const uint32_t b16 = w16 + kb * (kLaBlk >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_base + g * 128, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k), idesc1,
                           (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&aux->d1_full[g]);
          umma_commit(&aux->x_empty[s]);
        }
        __syncwarp();
        if (++s == p.x_stages) { s = 0; ph ^= 1u; }
      }
      if (do2) {
        if (do1) {                                                 // interleaved: one tile behind GEMM1
          if (li > 0) gemm2(li - 1);
        } else {
          gemm2(li);
        }
      }
    }
    if (do1 && do2 && n_tiles > 0) gemm2(n_tiles - 1);
  } else if (warp < 2 + kLaEpiWarps) {
    // ================================================================ epilogue: group g (8 warps) = tiles of parity g;
    // inside a group the two warps of a TMEM lane quarter split the columns (2 heads each; half of the output channels)
    const int ew = warp - 2;
    const int q = warp & 3;
    const int g = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int et_g = (ew & 7) * 32 + lane;
    const int m = q * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* const sq = qstg + static_cast<size_t>(g) * 2 * kLaBlk + m * 128;     // this thread's row in the group's staging
    const int nch = p.n_rows / 32;
    const int c2_lo = half == 0 ? 0 : (nch + 1) / 2, c2_hi = half == 0 ? (nch + 1) / 2 : nch;   // output chunks of this warp
    uint32_t r[32];
    if (et_g == 0) tma_prefetch_desc(&omap);
    // running tile coordinates of this group's tiles (no divisions in the loop)
    int tile = t_begin + g;
    int b = tile / p.tps, ti = tile % p.tps;
    int s = g % p.x_stages;
    const int s_step = 2 % p.x_stages;
    float rss_next = g < n_tiles ? __ldg(p.rowss + static_cast<long long>(tile) * kLaTok + m) : 1.f;
    for (int li = g; li < n_tiles; li += 2) {
      const int tok0 = ti * kLaTok, b_cur = b, s_cur = s;
      const uint32_t k2 = static_cast<uint32_t>(li >> 1);
      const float rs = 1.f / fmaxf(sqrtf(rss_next), 1e-12f);
      tile += 2;
      ti += 2;
      while (ti >= p.tps) { ti -= p.tps; ++b; }
      s += s_step;
      if (s >= p.x_stages) s -= p.x_stages;
      if (li + 2 < n_tiles) rss_next = __ldg(p.rowss + static_cast<long long>(tile) * kLaTok + m);   // tokens are contiguous
      // ---- epilogue 1: q = softmax over each head's 32 channels, * scale (unet.py:207,210); this warp: heads 2*half, +1
      mbar_wait(&aux->d1_full[g], k2 & 1u);
      tc_fence_after();
      if (et_g == 0) tma_store_wait_read0();                       // the previous tile's bulk store has left this buffer
      group_bar(g);
      const float k1 = rs * kLog2e;
      const float2 k12 = make_float2(k1, k1);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        tmem_ld32(trow + g * 128 + c * 32, r);
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aux->d1_empty[g]);
        }
        float mx = __uint_as_float(r[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        // softmax(rs * a): exponent (a - max a) * rs * log2e   (rs > 0, so the max commutes with the scaling).  The
        // exponentials overwrite the accumulator registers: 18 warps leave 96 registers per thread, a second 32-value array
        // spilled.
        const float2 nm2 = make_float2(-mx * k1, -mx * k1);
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x2 = __ffma2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), k12, nm2);
          const float2 ei = make_float2(ex2_fast(x2.x), ex2_fast(x2.y));
          r[2 * i] = __float_as_uint(ei.x);
          r[2 * i + 1] = __float_as_uint(ei.y);
          sum2 = __fadd2_rn(sum2, ei);
        }
        const float kk = __fdividef(p.q_scale, sum2.x + sum2.y);
        const float2 kk2 = make_float2(kk, kk);
        uint8_t* dst = sq + half * kLaBlk;                         // heads 2*half, 2*half+1 = K block `half` of q
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          uint4 o;
#define CCDM_E2(j) make_float2(__uint_as_float(r[2 * (j)]), __uint_as_float(r[2 * (j) + 1]))
          const float2 a0 = __fmul2_rn(CCDM_E2(gg * 4 + 0), kk2), a1 = __fmul2_rn(CCDM_E2(gg * 4 + 1), kk2);
          const float2 a2 = __fmul2_rn(CCDM_E2(gg * 4 + 2), kk2), a3 = __fmul2_rn(CCDM_E2(gg * 4 + 3), kk2);
#undef CCDM_E2
          o.x = pack_bf16(a0.x, a0.y);
          o.y = pack_bf16(a1.x, a1.y);
          o.z = pack_bf16(a2.x, a2.y);
          o.w = pack_bf16(a3.x, a3.y);
          *reinterpret_cast<uint4*>(dst + (((cc * 4 + gg) ^ (m & 7)) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->q_full[g]);
      // ---- epilogue 2: y = acc + bias -> RMSNorm * g * sqrt(C) -> + x -> bf16 (unet.py:88-89,198-199,72)
      const uint8_t* xrow = xring + static_cast<size_t>(s_cur) * x_stage_bytes + m * 128;
      mbar_wait(&aux->d2_full[g], k2 & 1u);
      tc_fence_after();
      const uint32_t d2 = trow + 256 + g * 128;
      float2 sq2 = make_float2(0.f, 0.f);
      for (int c = 0; c < nch; ++c) {                              // both warps of a quarter need the whole row's norm
        tmem_ld32(d2 + c * 32, r);
        tmem_ld_wait();
        const float2* b2 = reinterpret_cast<const float2*>(aux->bias + c * 32);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 v = __fadd2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), b2[i]);
          sq2 = __ffma2_rn(v, v, sq2);
        }
      }
      const float inv = 1.f / fmaxf(sqrtf(sq2.x + sq2.y), 1e-12f);
      const float2 inv2 = make_float2(inv, inv);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = c2_lo + j;
        if (c < c2_hi) {
          uint4 xr[4];                                             // residual rows: requested before the TMEM load completes
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            xr[gg] = make_uint4(0, 0, 0, 0);
            if ((c >> 1) < p.nkb)                                  // channels beyond the loaded K blocks are padding
              xr[gg] = *reinterpret_cast<const uint4*>(xrow + (c >> 1) * kLaBlk + ((((c & 1) * 4 + gg) ^ (m & 7)) << 4));
          }
          tmem_ld32(d2 + c * 32, r);
          tmem_ld_wait();
          const float2* b2 = reinterpret_cast<const float2*>(aux->bias + c * 32);
          const float2* g2 = reinterpret_cast<const float2*>(aux->gain + c * 32);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const uint4 x4 = xr[gg];
            float2 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = make_float2(__uint_as_float(r[gg * 8 + 2 * i]), __uint_as_float(r[gg * 8 + 2 * i + 1]));
              v[i] = __fmul2_rn(__fmul2_rn(__fadd2_rn(a, b2[gg * 4 + i]), inv2), g2[gg * 4 + i]);
            }
            v[0] = __fadd2_rn(v[0], make_float2(bf16_lo(x4.x), bf16_hi(x4.x)));
            v[1] = __fadd2_rn(v[1], make_float2(bf16_lo(x4.y), bf16_hi(x4.y)));
            v[2] = __fadd2_rn(v[2], make_float2(bf16_lo(x4.z), bf16_hi(x4.z)));
            v[3] = __fadd2_rn(v[3], make_float2(bf16_lo(x4.w), bf16_hi(x4.w)));
            uint4 o;
            o.x = pack_bf16(v[0].x, v[0].y);
            o.y = pack_bf16(v[1].x, v[1].y);
            o.z = pack_bf16(v[2].x, v[2].y);
            o.w = pack_bf16(v[3].x, v[3].y);
            *reinterpret_cast<uint4*>(sq + (c >> 1) * kLaBlk + ((((c & 1) * 4 + gg) ^ (m & 7)) << 4)) = o;
          }
        }
      }
      tc_fence_before();                                           // all TMEM reads of this tile by this warp are done
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&aux->d2_empty[g]);
        mbar_arrive(&aux->x_empty[s_cur]);                         // ... and so are the residual reads of the x stage
      }
      fence_proxy_async_smem();
      group_bar(g);
      if (et_g == 0) {
        const uint8_t* sbuf = qstg + static_cast<size_t>(g) * 2 * kLaBlk;
        for (int pn = 0; pn * 64 < p.C; ++pn) tma_store_3d(&omap, sbuf + pn * kLaBlk, pn * 64, tok0, b_cur);
        tma_store_commit();
      }
    }
    if (et_g == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ccdm

using namespace ccdm;

static int la_units(int tps) {                                     // tiles per unit: largest divisor of tps that is <= 8
  for (int g = 8; g > 1; --g)
    if (tps % g == 0) return g;
  return 1;
}

extern "C" int ccdm_linattn_fused_units(int32_t n) {
  if (n <= 0 || n % kLaTok != 0) return 0;
  const int tps = n / kLaTok;
  return tps / la_units(tps);
}

static int la_xmap(CUtensorMap* m, const void* x, int C, int n, int B) {
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)n * C * 2};
  cuuint32_t box[3] = {64, kLaTok, 1};
  return encode_map_bf16(m, x, 3, dims, str, box);
}

static int la_wmap(CUtensorMap* m, const void* wqkv, int nkb) {
  cuuint64_t dims[2] = {(cuuint64_t)nkb * 64, 384};
  cuuint64_t str[1] = {(cuuint64_t)nkb * 64 * 2};
  cuuint32_t box[2] = {64, 128};
  return encode_map_bf16(m, wqkv, 2, dims, str, box);
}

#define CCDM_LA_COMMON_CHECKS(what)                                                                                  \
  CCDM_REQUIRE(x && wqkv && rowss && B > 0, CCDM_ERR_BAD_ARG, what ": null pointer / empty batch");                  \
  CCDM_REQUIRE(C >= 8 && C <= 128 && C % 8 == 0, CCDM_ERR_UNSUPPORTED_SHAPE, what ": C=%d (needs C <= 128, C %% 8 == 0)", C); \
  CCDM_REQUIRE(n >= kLaTok && n % kLaTok == 0, CCDM_ERR_UNSUPPORTED_SHAPE, what ": n=%d must be a multiple of 128", n);        \
  CCDM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wqkv) & 15) == 0,          \
               CCDM_ERR_BAD_ARG, what ": 16-byte alignment")

extern "C" int ccdm_linattn_kv_partials(const void* x, int32_t B, int32_t n, int32_t C, const float* rowss, const void* wqkv,
                                        const float* kbias, float* part, float* psum, void* stream) {
  CCDM_LA_COMMON_CHECKS("linattn_kv_partials");
  CCDM_REQUIRE(kbias && part && psum, CCDM_ERR_BAD_ARG, "linattn_kv_partials: null kbias / part / psum");
  La1Params p;
  std::memset(&p, 0, sizeof(p));
  p.n = n;
  p.nkb = (C + 63) / 64;
  p.tps = n / kLaTok;
  p.G = la_units(p.tps);
  p.ups = p.tps / p.G;
  p.units_total = B * p.ups;
  const int sms = num_sms();
  int grid = sms < p.units_total ? sms : p.units_total;
  p.units_per_cta = (p.units_total + grid - 1) / grid;
  grid = (p.units_total + p.units_per_cta - 1) / p.units_per_cta;
  p.rowss = rowss; p.kbias = kbias; p.part = part; p.psum = psum;
  // shared memory: weights nkb*32K | x ring | P/V staging (64K per buffer) | ones 16K | aux
  const size_t fixed = (size_t)p.nkb * 2 * kLaBlk + kLaBlk + sizeof(La1Aux) + 1024;
  const size_t budget = 227 * 1024;
  p.pv_bufs = 2;
  p.x_stages = 3;
  auto total = [&]() { return fixed + (size_t)p.x_stages * p.nkb * kLaBlk + (size_t)p.pv_bufs * 4 * kLaBlk; };
  if (total() > budget) p.x_stages = 2;
  if (total() > budget) p.pv_bufs = 1;
  CCDM_REQUIRE(total() <= budget, CCDM_ERR_UNSUPPORTED_SHAPE, "linattn_kv_partials: shared memory plan does not fit");
  CUtensorMap xmap, wmap;
  int rc = la_xmap(&xmap, x, C, n, B);
  if (rc != CCDM_OK) return rc;
  rc = la_wmap(&wmap, wqkv, p.nkb);
  if (rc != CCDM_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "linattn_kv_kernel: cudaFuncSetAttribute");
    attr_set = true;
  }
  linattn_kv_kernel<<<grid, kLaThreads, total(), (cudaStream_t)stream>>>(xmap, wmap, p);
  return after_launch("linattn_kv_kernel");
}

extern "C" int ccdm_linattn_fold_partials(const float* part, const float* psum, int32_t B, int32_t units_per_sample,
                                          const float* w_out, int32_t C, int32_t n_rows, void* wfold, void* stream) {
  CCDM_REQUIRE(part && psum && w_out && wfold && B > 0 && units_per_sample > 0 && C > 0 && C <= 128 && n_rows >= C,
               CCDM_ERR_BAD_ARG, "linattn_fold_partials: bad args (C <= 128)");
  const size_t smem = (size_t)C * 128 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_fold_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 4);
    if (e != cudaSuccess) return cuda_fail(e, "linattn_fold_parts_kernel: cudaFuncSetAttribute");
    attr_set = true;
  }
  int grid = 3 * num_sms();                                        // <= 64 KB of shared memory each: three CTAs per SM
  if (grid > B) grid = B;
  linattn_fold_parts_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(part, psum, units_per_sample, w_out, C, n_rows,
                                                                      (__nv_bfloat16*)wfold, B);
  return after_launch("linattn_fold_parts_kernel");
}

extern "C" int ccdm_linattn_q_out(const void* x, int32_t B, int32_t n, int32_t C, const float* rowss, const void* wqkv,
                                  const void* wfold, int32_t n_rows, const float* bias, const float* gain, float gain_mul,
                                  float q_scale, void* out, void* stream) {
  CCDM_LA_COMMON_CHECKS("linattn_q_out");
  CCDM_REQUIRE(wfold && bias && gain && out && (reinterpret_cast<uintptr_t>(out) & 15) == 0, CCDM_ERR_BAD_ARG,
               "linattn_q_out: null / misaligned pointer");
  CCDM_REQUIRE(n_rows >= C && n_rows % 32 == 0 && n_rows <= 128, CCDM_ERR_UNSUPPORTED_SHAPE, "linattn_q_out: n_rows=%d", n_rows);
  La2Params p;
  std::memset(&p, 0, sizeof(p));
  p.n = n;
  p.nkb = (C + 63) / 64;
  p.tps = n / kLaTok;
  p.tiles_total = B * p.tps;
  p.C = C;
  p.n_rows = n_rows;
  const int sms = num_sms();
  int grid = sms < p.tiles_total ? sms : p.tiles_total;
  p.tiles_per_cta = (p.tiles_total + grid - 1) / grid;
  grid = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.rowss = rowss; p.bias = bias; p.gain = gain; p.gain_mul = gain_mul; p.q_scale = q_scale;
  const size_t wf_slot = ((size_t)2 * n_rows * 128 + 1023) & ~(size_t)1023;
  const size_t fixed = (size_t)p.nkb * kLaBlk + 4 * (size_t)kLaBlk + sizeof(La2Aux) + 1024;
  const size_t budget = 227 * 1024;
  p.x_stages = 4;
  p.wf_bufs = 2;
  auto total = [&]() { return fixed + (size_t)p.x_stages * p.nkb * kLaBlk + (size_t)p.wf_bufs * wf_slot; };
  if (total() > budget) p.x_stages = 3;
  if (total() > budget) p.wf_bufs = 1;
  CCDM_REQUIRE(total() <= budget, CCDM_ERR_UNSUPPORTED_SHAPE, "linattn_q_out: shared memory plan does not fit");
  CUtensorMap xmap, wmap, fmap, omap;
  int rc = la_xmap(&xmap, x, C, n, B);
  if (rc != CCDM_OK) return rc;
  rc = la_wmap(&wmap, wqkv, p.nkb);
  if (rc != CCDM_OK) return rc;
  {
    cuuint64_t dims[2] = {128, (cuuint64_t)B * n_rows};
    cuuint64_t str[1] = {256};
    cuuint32_t box[2] = {64, (cuuint32_t)n_rows};
    rc = encode_map_bf16(&fmap, wfold, 2, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  rc = la_xmap(&omap, out, C, n, B);
  if (rc != CCDM_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_qout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "linattn_qout_kernel: cudaFuncSetAttribute");
    attr_set = true;
  }
  linattn_qout_kernel<<<grid, kLa2Threads, total(), (cudaStream_t)stream>>>(xmap, wmap, fmap, omap, p);
  return after_launch("linattn_qout_kernel");
}
