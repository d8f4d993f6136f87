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
// Warp roles (both tensor kernels): warp 0 TMA producer, warp 1 tcgen05 issuer, warps 2..9 epilogue.
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box);

constexpr int kLaThreads = 320;
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
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

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
      mbar_init(&aux->d1_empty[s], 8);
      mbar_init(&aux->pv_full[s], 8);
      mbar_init(&aux->pv_empty[s], 1);
    }
    mbar_init(&aux->d2_full, 1);
    mbar_init(&aux->d2_empty, 8);
    fence_mbar_init();
  }
  for (int i = tid; i < kLaBlk / 4; i += kLaThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0
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
    // ================================================================ tcgen05 issuer
    const uint32_t idesc1 = umma_idesc_bf16(128, 128);             // [k | v] item: 128 tokens x 128 channels, K-major
    const uint32_t idesc_ctx = umma_idesc_bf16_mn(128, 128);       // P^T V, both operands token-major
    const uint32_t idesc_sum = umma_idesc_bf16_mn(128, 16);        // P^T 1
    const uint32_t ones_addr = smem_u32(ones);
    mbar_wait(&aux->w_full, 0);
    tc_fence_after();
    auto gemm2 = [&](int t) {                                      // context MMAs of local tile t
      const int tl = t % p.G, ul = t / p.G;
      const int buf = t % p.pv_bufs;
      const uint32_t use = static_cast<uint32_t>(t / p.pv_bufs);
      if (tl == 0) {                                               // new unit: the previous unit's D2 must have been read
        mbar_wait(&aux->d2_empty, (static_cast<uint32_t>(ul) & 1u) ^ 1u);
        tc_fence_after();
      }
      mbar_wait(&aux->pv_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t base = smem_u32(stg + static_cast<size_t>(buf) * 4 * kLaBlk);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kLaTok / 16; ++k) {
          const uint64_t adesc = umma_desc_mn_sw128(base + k * 2048, kLaBlk);                // P^T (128 channels)
          const uint64_t bdesc = umma_desc_mn_sw128(base + 2 * kLaBlk + k * 2048, kLaBlk);   // V   (128 channels)
          const uint64_t odesc = umma_desc_mn_sw128(ones_addr + k * 2048, kLaBlk);
          const uint32_t acc = (tl | k) != 0 ? 1u : 0u;
          umma_bf16_ss(tmem_base + d2_col, adesc, bdesc, idesc_ctx, acc);
          umma_bf16_ss(tmem_base + ds_col, adesc, odesc, idesc_sum, acc);
        }
        umma_commit(&aux->pv_empty[buf]);
        if (tl == p.G - 1) umma_commit(&aux->d2_full);
      }
      __syncwarp();
    };
    int s = 0;
    uint32_t ph = 0;
    for (int li = 0; li < n_tiles; ++li) {
      mbar_wait(&aux->x_full[s], ph);
      tc_fence_after();
      const uint32_t xa = smem_u32(xring + static_cast<size_t>(s) * x_stage_bytes);
      for (int j = 0; j < 2; ++j) {
        mbar_wait(&aux->d1_empty[j], (static_cast<uint32_t>(li) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          for (int kb = 0; kb < p.nkb; ++kb) {
            const uint32_t a16 = ((xa + kb * kLaBlk) & 0x3FFFF) >> 4;
            const uint32_t b16 = ((smem_u32(wres) + (kb * 2 + j) * kLaBlk) & 0x3FFFF) >> 4;
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
      if (li > 0) gemm2(li - 1);
      if (++s == p.x_stages) { s = 0; ph ^= 1u; }
    }
    if (n_tiles > 0) gemm2(n_tiles - 1);
  } else {
    // ================================================================ epilogue (8 warps: 128 rows x 2 column halves)
    const int ew = warp - 2;
    const int q = warp & 3;                                        // TMEM lane quarter this warp may read
    const int half = ew >> 2;                                      // 64-channel half of every 128-channel item
    const int m = q * 32 + lane;                                   // token row inside the tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t r[32];
    int li = 0, pend_u = -1;
    uint32_t flushes = 0;
    auto flush = [&](int u) {                                      // D2 / S of a finished unit -> global partials
      mbar_wait(&aux->d2_full, flushes & 1u);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld32(trow + d2_col + q * 32, r);                        // head q: rows q*32+d, columns q*32 + e
      tmem_ld32(trow + ds_col, sv);                                // every column of S holds sum_n p[n][q*32+d]
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->d2_empty);
      float4* o = reinterpret_cast<float4*>(p.part + (static_cast<long long>(u) * 128 + m) * 32 + half * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_float4(__uint_as_float(r[half * 16 + 4 * j]), __uint_as_float(r[half * 16 + 4 * j + 1]),
                           __uint_as_float(r[half * 16 + 4 * j + 2]), __uint_as_float(r[half * 16 + 4 * j + 3]));
      if (half == 0) p.psum[static_cast<long long>(u) * 128 + m] = __uint_as_float(sv[0]);
      ++flushes;
    };
    for (int u = u_begin; u < u_end; ++u) {
      const int b = u / p.ups, tile0 = (u % p.ups) * p.G;
      for (int tl = 0; tl < p.G; ++tl, ++li) {
        const long long row = static_cast<long long>(b) * p.n + static_cast<long long>(tile0 + tl) * kLaTok + m;
        const float rs = 1.f / fmaxf(sqrtf(__ldg(p.rowss + row)), 1e-12f);
        const int buf = li % p.pv_bufs;
        const uint32_t use = static_cast<uint32_t>(li / p.pv_bufs);
        uint8_t* sb = stg + static_cast<size_t>(buf) * 4 * kLaBlk + m * 128;
        mbar_wait(&aux->pv_empty[buf], (use & 1u) ^ 1u);           // the context MMAs that read this buffer are done
        // ---- item 0: p = exp(k / |x| - bound) = 2^(acc * rs*log2e + kbias)
        mbar_wait(&aux->d1_full[0], static_cast<uint32_t>(li) & 1u);
        tc_fence_after();
        {
          const float2 rs2 = make_float2(rs * kLog2e, rs * kLog2e);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            tmem_ld32(trow + (half * 2 + cc) * 32, r);
            tmem_ld_wait();
            if (cc == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&aux->d1_empty[0]);
            }
            const float2* kb2 = reinterpret_cast<const float2*>(aux->kbias + (half * 2 + cc) * 32);
            uint8_t* dst = sb + (0 + half) * kLaBlk;
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
        }
        // ---- item 1: v / |x|
        mbar_wait(&aux->d1_full[1], static_cast<uint32_t>(li) & 1u);
        tc_fence_after();
        {
          const float2 rs2 = make_float2(rs, rs);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            tmem_ld32(trow + 128 + (half * 2 + cc) * 32, r);
            tmem_ld_wait();
            if (cc == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&aux->d1_empty[1]);
            }
            uint8_t* dst = sb + (2 + half) * kLaBlk;
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
        }
        fence_proxy_async_smem();                                  // generic-proxy writes -> visible to the MMA unit
        __syncwarp();
        if (lane == 0) mbar_arrive(&aux->pv_full[buf]);
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
// column sum S (softmax over tokens) and folds the context into to_out[0]'s weight (unet.py:198,212-216).
__global__ void __launch_bounds__(128) linattn_fold_parts_kernel(const float* __restrict__ part,
                                                                 const float* __restrict__ psum, int ups,
                                                                 const float* __restrict__ w_out, int C, int n_rows,
                                                                 __nv_bfloat16* __restrict__ wfold) {
  const int b = blockIdx.x, t = threadIdx.x;                       // t = h*32 + d
  const int h = t >> 5;
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
  for (int c = 0; c < C; ++c) {
    const float4* wr = reinterpret_cast<const float4*>(w_out + static_cast<long long>(c) * 128 + h * 32);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 wv = __ldg(wr + j);
      acc = fmaf(wv.x, cr[4 * j], acc);
      acc = fmaf(wv.y, cr[4 * j + 1], acc);
      acc = fmaf(wv.z, cr[4 * j + 2], acc);
      acc = fmaf(wv.w, cr[4 * j + 3], acc);
    }
    wf[static_cast<long long>(c) * 128] = __float2bfloat16(acc);
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

__global__ void __launch_bounds__(kLaThreads, 1) linattn_qout_kernel(const __grid_constant__ CUtensorMap xmap,
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
      mbar_init(&aux->x_empty[s], 5);                              // GEMM1 commit + the four warps that read the residual
    }
    mbar_init(&aux->w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aux->wf_full[s], 1);
      mbar_init(&aux->wf_empty[s], 1);
      mbar_init(&aux->d1_full[s], 1);
      mbar_init(&aux->d1_empty[s], 4);
      mbar_init(&aux->q_full[s], 4);
      mbar_init(&aux->d2_full[s], 1);
      mbar_init(&aux->d2_empty[s], 4);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < 128; i += kLaThreads) {
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
    // ================================================================ tcgen05 issuer
    const uint32_t idesc1 = umma_idesc_bf16(128, 128);
    const uint32_t idesc2 = umma_idesc_bf16(128, static_cast<uint32_t>(p.n_rows));
    mbar_wait(&aux->w_full, 0);
    tc_fence_after();
    int b_prev = -1, sl = -1;
    auto gemm2 = [&](int t) {                                      // output projection of local tile t
      const int g = t & 1;
      const uint32_t k2 = static_cast<uint32_t>(t >> 1);
      const int b = (t_begin + t) / p.tps;
      const bool last_of_sample = (t == n_tiles - 1) || ((t_begin + t + 1) / p.tps != b);
      if (b != b_prev) {
        b_prev = b;
        ++sl;
        mbar_wait(&aux->wf_full[sl % p.wf_bufs], static_cast<uint32_t>(sl / p.wf_bufs) & 1u);
      }
      const int slot = sl % p.wf_bufs;
      mbar_wait(&aux->q_full[g], k2 & 1u);
      mbar_wait(&aux->d2_empty[g], (k2 & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t qa = smem_u32(qstg + static_cast<size_t>(g) * 2 * kLaBlk);
      const uint32_t fa = smem_u32(wfr + static_cast<size_t>(slot) * wf_slot);
      if (elect_one()) {
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a16 = ((qa + kb * kLaBlk) & 0x3FFFF) >> 4;
          const uint32_t b16 = ((fa + kb * p.n_rows * 128) & 0x3FFFF) >> 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + 256 + g * 128, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k),
                         idesc2, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&aux->d2_full[g]);
        if (last_of_sample) umma_commit(&aux->wf_empty[slot]);
      }
      __syncwarp();
    };
    int s = 0;
    uint32_t ph = 0;
    for (int li = 0; li < n_tiles; ++li) {
      const int g = li & 1;
      mbar_wait(&aux->x_full[s], ph);
      mbar_wait(&aux->d1_empty[g], (static_cast<uint32_t>(li >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t xa = smem_u32(xring + static_cast<size_t>(s) * x_stage_bytes);
      if (elect_one()) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          const uint32_t a16 = ((xa + kb * kLaBlk) & 0x3FFFF) >> 4;
          const uint32_t b16 = ((smem_u32(wres) + kb * kLaBlk) & 0x3FFFF) >> 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + g * 128, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k), idesc1,
                         (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&aux->d1_full[g]);
        umma_commit(&aux->x_empty[s]);
      }
      __syncwarp();
      if (li > 0) gemm2(li - 1);
      if (++s == p.x_stages) { s = 0; ph ^= 1u; }
    }
    if (n_tiles > 0) gemm2(n_tiles - 1);
  } else {
    // ================================================================ epilogue: group g = tiles of parity g
    const int ew = warp - 2;
    const int q = warp & 3;
    const int g = ew >> 2;
    const int et_g = (ew & 3) * 32 + lane;
    const int m = q * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* const sq = qstg + static_cast<size_t>(g) * 2 * kLaBlk + m * 128;     // this thread's row in the group's staging
    const int nch = p.n_rows / 32;
    uint32_t r[32];
    if (et_g == 0) tma_prefetch_desc(&omap);
    for (int li = g; li < n_tiles; li += 2) {
      const int tile = t_begin + li, b = tile / p.tps, tok0 = (tile % p.tps) * kLaTok;
      const uint32_t k2 = static_cast<uint32_t>(li >> 1);
      const int s = li % p.x_stages;
      const float rs = 1.f / fmaxf(sqrtf(__ldg(p.rowss + static_cast<long long>(b) * p.n + tok0 + m)), 1e-12f);
      // ---- epilogue 1: q = softmax over each head's 32 channels, * scale (unet.py:207,210)
      mbar_wait(&aux->d1_full[g], k2 & 1u);
      tc_fence_after();
      if (et_g == 0) tma_store_wait_read0();                       // the previous tile's bulk store has left this buffer
      group_bar(g);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(trow + g * 128 + c * 32, r);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aux->d1_empty[g]);
        }
        float mx = __uint_as_float(r[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        // softmax(rs * a): exponent (a - max a) * rs * log2e   (rs > 0, so the max commutes with the scaling)
        const float k1 = rs * kLog2e;
        const float2 k12 = make_float2(k1, k1), nm2 = make_float2(-mx * k1, -mx * k1);
        float2 e[16];
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x2 = __ffma2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), k12, nm2);
          e[i] = make_float2(ex2_fast(x2.x), ex2_fast(x2.y));
          sum2 = __fadd2_rn(sum2, e[i]);
        }
        const float kk = __fdividef(p.q_scale, sum2.x + sum2.y);
        const float2 kk2 = make_float2(kk, kk);
        uint8_t* dst = sq + (c >> 1) * kLaBlk;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          uint4 o;
          const float2 a0 = __fmul2_rn(e[gg * 4 + 0], kk2), a1 = __fmul2_rn(e[gg * 4 + 1], kk2);
          const float2 a2 = __fmul2_rn(e[gg * 4 + 2], kk2), a3 = __fmul2_rn(e[gg * 4 + 3], kk2);
          o.x = pack_bf16(a0.x, a0.y);
          o.y = pack_bf16(a1.x, a1.y);
          o.z = pack_bf16(a2.x, a2.y);
          o.w = pack_bf16(a3.x, a3.y);
          *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + gg) ^ (m & 7)) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->q_full[g]);
      // ---- epilogue 2: y = acc + bias -> RMSNorm * g * sqrt(C) -> + x -> bf16 (unet.py:88-89,198-199,72)
      mbar_wait(&aux->d2_full[g], k2 & 1u);
      tc_fence_after();
      const uint32_t d2 = trow + 256 + g * 128;
      float2 sq2 = make_float2(0.f, 0.f);
      for (int c = 0; c < nch; ++c) {
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
      const uint8_t* xrow = xring + static_cast<size_t>(s) * x_stage_bytes + m * 128;
      for (int c = 0; c < nch; ++c) {
        tmem_ld32(d2 + c * 32, r);
        tmem_ld_wait();
        if (c == nch - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aux->d2_empty[g]);
        }
        const float2* b2 = reinterpret_cast<const float2*>(aux->bias + c * 32);
        const float2* g2 = reinterpret_cast<const float2*>(aux->gain + c * 32);
        const bool have_x = (c >> 1) < p.nkb;                      // residual channels beyond the loaded K blocks are padding
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const int pos = (((c & 1) * 4 + gg) ^ (m & 7)) << 4;
          uint4 xr = make_uint4(0, 0, 0, 0);
          if (have_x) xr = *reinterpret_cast<const uint4*>(xrow + (c >> 1) * kLaBlk + pos);
          float2 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = make_float2(__uint_as_float(r[gg * 8 + 2 * i]), __uint_as_float(r[gg * 8 + 2 * i + 1]));
            v[i] = __fmul2_rn(__fmul2_rn(__fadd2_rn(a, b2[gg * 4 + i]), inv2), g2[gg * 4 + i]);
          }
          v[0] = __fadd2_rn(v[0], make_float2(bf16_lo(xr.x), bf16_hi(xr.x)));
          v[1] = __fadd2_rn(v[1], make_float2(bf16_lo(xr.y), bf16_hi(xr.y)));
          v[2] = __fadd2_rn(v[2], make_float2(bf16_lo(xr.z), bf16_hi(xr.z)));
          v[3] = __fadd2_rn(v[3], make_float2(bf16_lo(xr.w), bf16_hi(xr.w)));
          uint4 o;
          o.x = pack_bf16(v[0].x, v[0].y);
          o.y = pack_bf16(v[1].x, v[1].y);
          o.z = pack_bf16(v[2].x, v[2].y);
          o.w = pack_bf16(v[3].x, v[3].y);
          *reinterpret_cast<uint4*>(sq + (c >> 1) * kLaBlk + pos) = o;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->x_empty[s]);                // residual rows read: the x stage may be refilled
      fence_proxy_async_smem();
      group_bar(g);
      if (et_g == 0) {
        const uint8_t* sbuf = qstg + static_cast<size_t>(g) * 2 * kLaBlk;
        for (int pn = 0; pn * 64 < p.C; ++pn) tma_store_3d(&omap, sbuf + pn * kLaBlk, pn * 64, tok0, b);
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
  CCDM_REQUIRE(part && psum && w_out && wfold && B > 0 && units_per_sample > 0 && C > 0 && n_rows >= C, CCDM_ERR_BAD_ARG,
               "linattn_fold_partials: bad args");
  linattn_fold_parts_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(part, psum, units_per_sample, w_out, C, n_rows,
                                                                (__nv_bfloat16*)wfold);
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
  linattn_qout_kernel<<<grid, kLaThreads, total(), (cudaStream_t)stream>>>(xmap, wmap, fmap, omap, p);
  return after_launch("linattn_qout_kernel");
}
