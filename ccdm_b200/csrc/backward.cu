// Backward of the Block tail (RMSNorm -> (1+scale)*x+shift -> SiLU), CCDM_unified/models/unet.py:88-89,145-151.
// HBM-bound row kernel: reads dY and Z once, writes dZ once, and leaves three per-(sample, channel) sums from which
// the gradients of gain, bias, scale and shift follow (ccdm_block_bwd_finish).
#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

constexpr int kBwdThreads = 256;
constexpr int kBwdMaxC = 1024;

// Per-thread asynchronous staging ring (LDGSTS): every thread copies the 16-byte vectors it will consume itself kStages - 1
// row groups ahead into its OWN shared-memory slots, so the bytes in flight per SM no longer depend on registers or on the
// number of resident warps (the first version kept one row group ahead in registers and ran at 14-30 % of the HBM peak with
// 8-16 warps per SM).  No block barrier is needed in the row loop: a slot is written and read by the same thread.
constexpr int bwd_stages(int chunks) { return chunks == 1 ? 6 : chunks == 2 ? 4 : 3; }

// G lanes share one row; each owns kChunks chunks of 8 channels (see row_lane_plan); kMinCtas resident CTAs per SM
template <int kChunks, int kMinCtas, bool kSilu>
__global__ void __launch_bounds__(kBwdThreads, kMinCtas) block_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ z,
                                                                          uint4* __restrict__ dz, int rows_per_sample,
                                                                          int rows_per_block, int C, const float* __restrict__ gain,
                                                                          float gain_mul, const float* __restrict__ ss, int ss_ld,
                                                                          int ss_off, float* __restrict__ sums, int B,
                                                                          uint32_t flags, int G) {
  constexpr int kStages = bwd_stages(kChunks);
  __shared__ float coef_s[2][kBwdMaxC];                            // a = gain * (1 + scale), shift
  extern __shared__ float bwd_ring_f[];                            // [stage][chunk][z | dy][thread] 16-byte slots
  uint4* const ring = reinterpret_cast<uint4*>(bwd_ring_f);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = C >> 3;                                       // 8 channels (16 bytes) per chunk
  const int kRowsPerWarp = 32 / G;
  const int sub = lane / G, gl = lane - sub * G;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, rows_per_sample);
  for (int c = tid; c < C; c += kBwdThreads) {
    float sc = 0.f, sf = 0.f;
    if (flags & CCDM_EPI_SS) {
      sc = ss[(size_t)b * ss_ld + ss_off + c];
      sf = ss[(size_t)b * ss_ld + ss_off + C + c];
    }
    coef_s[0][c] = gain[c] * gain_mul * (1.f + sc);
    coef_s[1][c] = sf;
  }
  __syncthreads();

  float s1[kChunks][8], s2[kChunks][8], s3[kChunks][8];
#pragma unroll
  for (int k = 0; k < kChunks; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[k][j] = s2[k][j] = s3[k][j] = 0.f;

  const int rstep = (kBwdThreads / 32) * kRowsPerWarp;
  const int rb0 = r0 + warp * kRowsPerWarp;
  auto slot = [&](int st, int k, int which) { return ring + ((st * kChunks + k) * 2 + which) * kBwdThreads + tid; };
  auto fetch = [&](int rb, int st) {                               // one commit group per row group, also when it is empty
    const int r = rb + sub;
    const bool ok = sub < kRowsPerWarp && r < r1;
    const size_t off = ((size_t)b * rows_per_sample + (ok ? r : r0)) * nchunk;
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int ch = gl + G * k;
      const bool v = ok && ch < nchunk;
      cp_async16(slot(st, k, 0), z + off + (v ? ch : 0), v);
      cp_async16(slot(st, k, 1), dy + off + (v ? ch : 0), v);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < kStages - 1; ++st) fetch(rb0 + st * rstep, st);
  int st_rd = 0, st_wr = kStages - 1;
  for (int rb = rb0; rb < r1; rb += rstep) {
    const int r = rb + sub;
    const bool live = sub < kRowsPerWarp && r < r1;
    const size_t rowoff = ((size_t)b * rows_per_sample + r) * nchunk;
    fetch(rb + (kStages - 1) * rstep, st_wr);                      // rows past r1 become zero-filled slots
    cp_async_wait<kStages - 1>();
    // packed fp32x2 arithmetic throughout (FFMA2 / FMUL2 / FADD2)
    float2 zf[kChunks][4], gy[kChunks][4];
    float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const uint4 zu = *slot(st_rd, k, 0), du = *slot(st_rd, k, 1);
      zf[k][0] = make_float2(bf16_lo(zu.x), bf16_hi(zu.x));
      zf[k][1] = make_float2(bf16_lo(zu.y), bf16_hi(zu.y));
      zf[k][2] = make_float2(bf16_lo(zu.z), bf16_hi(zu.z));
      zf[k][3] = make_float2(bf16_lo(zu.w), bf16_hi(zu.w));
      gy[k][0] = make_float2(bf16_lo(du.x), bf16_hi(du.x));
      gy[k][1] = make_float2(bf16_lo(du.y), bf16_hi(du.y));
      gy[k][2] = make_float2(bf16_lo(du.z), bf16_hi(du.z));
      gy[k][3] = make_float2(bf16_lo(du.w), bf16_hi(du.w));
#pragma unroll
      for (int j = 0; j < 4; ++j) sq2 = __ffma2_rn(zf[k][j], zf[k][j], sq2);
    }
    if (++st_rd == kStages) st_rd = 0;
    if (++st_wr == kStages) st_wr = 0;
    const float sq = seg_sum(sq2.x + sq2.y, gl, G, lane);
    const float inv = rsqrtf(fmaxf(sq, 1e-24f));                   // 1 / max(|z|, 1e-12)
    const float2 inv2 = make_float2(inv, inv);
    const float2 one2 = make_float2(1.f, 1.f), half2 = make_float2(0.5f, 0.5f);
    float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int c0 = min(gl + G * k, nchunk - 1) * 8;              // lanes past the row read a valid coefficient slot
      const float4 a_lo = *reinterpret_cast<const float4*>(&coef_s[0][c0]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&coef_s[0][c0 + 4]);
      const float4 h_lo = *reinterpret_cast<const float4*>(&coef_s[1][c0]);
      const float4 h_hi = *reinterpret_cast<const float4*>(&coef_s[1][c0 + 4]);
      const float2 av[4] = {make_float2(a_lo.x, a_lo.y), make_float2(a_lo.z, a_lo.w), make_float2(a_hi.x, a_hi.y),
                            make_float2(a_hi.z, a_hi.w)};
      const float2 hv[4] = {make_float2(h_lo.x, h_lo.y), make_float2(h_lo.z, h_lo.w), make_float2(h_hi.x, h_hi.y),
                            make_float2(h_hi.z, h_hi.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 aj = av[j];
        const float2 zh = __fmul2_rn(zf[k][j], inv2);
        float2 du = gy[k][j];
        if constexpr (kSilu) {
          // silu'(u) = sig * (1 + u * (1 - sig)),  sig = 0.5 + 0.5 tanh(u / 2)
          const float2 u = __ffma2_rn(zh, aj, hv[j]);
          const float2 hu = __fmul2_rn(u, half2);
          const float2 sig = __ffma2_rn(make_float2(tanh_fast(hu.x), tanh_fast(hu.y)), half2, half2);
          const float2 t = __ffma2_rn(u, make_float2(1.f - sig.x, 1.f - sig.y), one2);
          du = __fmul2_rn(du, __fmul2_rn(sig, t));
        }
        const float2 s1v = __ffma2_rn(du, zh, make_float2(s1[k][2 * j], s1[k][2 * j + 1]));
        s1[k][2 * j] = s1v.x; s1[k][2 * j + 1] = s1v.y;
        const float2 s2v = __fadd2_rn(make_float2(s2[k][2 * j], s2[k][2 * j + 1]), du);
        s2[k][2 * j] = s2v.x; s2[k][2 * j + 1] = s2v.y;
        const float2 dzh = __fmul2_rn(du, aj);
        dot2 = __ffma2_rn(zh, dzh, dot2);
        zf[k][j] = zh;
        gy[k][j] = dzh;
      }
    }
    const float dot = seg_sum(dot2.x + dot2.y, gl, G, lane);
    const float2 ndot2 = make_float2(-dot, -dot);
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const int ch = gl + G * k;
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = __fmul2_rn(__ffma2_rn(zf[k][j], ndot2, gy[k][j]), inv2);
        const float2 s3v = __fadd2_rn(make_float2(s3[k][2 * j], s3[k][2 * j + 1]), o[j]);
        s3[k][2 * j] = s3v.x; s3[k][2 * j + 1] = s3v.y;
      }
      if (live && ch < nchunk)
        dz[rowoff + ch] = make_uint4(pack_bf16(o[0].x, o[0].y), pack_bf16(o[1].x, o[1].y), pack_bf16(o[2].x, o[2].y),
                                     pack_bf16(o[3].x, o[3].y));
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                                 // every thread is done with its ring slots
  // Column sums without shared-memory atomics (24 x kChunks of them per thread, up to 32 threads per address, made the first
  // version's tail longer than its row loop at C = 288): rows of one warp are folded with shuffles, each warp leaves its
  // partial in its own slab of the (now free) ring, and every output gets ONE global atomic per CTA.
  float* const slab = bwd_ring_f + warp * 3 * C;                   // [warp][3][C]; 8 * 3 * C * 4 <= ring bytes (launcher)
#pragma unroll
  for (int k = 0; k < kChunks; ++k) {
    const int ch = gl + G * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v1 = s1[k][j], v2 = s2[k][j], v3 = s3[k][j];
      for (int r = 1; r < kRowsPerWarp; ++r) {                     // lanes of row `sub` = 0 collect the other rows
        v1 += __shfl_down_sync(0xffffffffu, s1[k][j], r * G);
        v2 += __shfl_down_sync(0xffffffffu, s2[k][j], r * G);
        v3 += __shfl_down_sync(0xffffffffu, s3[k][j], r * G);
      }
      if (sub == 0 && ch < nchunk) {
        slab[ch * 8 + j] = v1;
        slab[C + ch * 8 + j] = v2;
        slab[2 * C + ch * 8 + j] = v3;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 3 * C; i += kBwdThreads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdThreads / 32; ++w) v += bwd_ring_f[w * 3 * C + i];
    const int w3 = i / C, c = i - w3 * C;
    atomicAdd(&sums[((size_t)w3 * B + b) * C + c], v);
  }
}

__global__ void __launch_bounds__(256) block_bwd_finish_kernel(const float* __restrict__ sums, int B, int C,
                                                               const float* __restrict__ gain, float gain_mul,
                                                               const float* __restrict__ ss, int ss_ld, int ss_off,
                                                               float* __restrict__ d_ss, float* __restrict__ dgain,
                                                               float* __restrict__ dbias) {
  // thread (b-lane, c): consecutive threads take consecutive channels (coalesced), 8 sample lanes per CTA
  __shared__ float red[2][8][32];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float dg = 0.f, db = 0.f;
  if (c < C) {
    const float g = gain[c] * gain_mul;
    for (int b = bl; b < B; b += 8) {
      const float s0 = sums[((size_t)0 * B + b) * C + c];
      const float s1 = sums[((size_t)1 * B + b) * C + c];
      db += sums[((size_t)2 * B + b) * C + c];
      const float sc = ss ? ss[(size_t)b * ss_ld + ss_off + c] : 0.f;
      dg = fmaf(1.f + sc, s0, dg);
      if (d_ss) {
        d_ss[(size_t)b * ss_ld + ss_off + c] = g * s0;
        d_ss[(size_t)b * ss_ld + ss_off + C + c] = s1;
      }
    }
  }
  red[0][bl][cl] = dg;
  red[1][bl][cl] = db;
  __syncthreads();
  if (bl == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      dg += red[0][i][cl];
      db += red[1][i][cl];
    }
    if (dgain) dgain[c] += gain_mul * dg;
    if (dbias) dbias[c] += db;
  }
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_block_bwd(const void* dy, const void* z, void* dz, int64_t rows, int32_t C, int32_t rows_per_sample,
                              const float* gain, float gain_mul, const float* scale_shift, int32_t ss_ld, int32_t ss_off,
                              float* sums, uint32_t flags, void* stream) {
  CCDM_REQUIRE(dy && z && dz && gain && sums, CCDM_ERR_BAD_ARG, "block_bwd: null pointer");
  CCDM_REQUIRE(rows > 0 && rows_per_sample > 0 && rows % rows_per_sample == 0, CCDM_ERR_BAD_ARG,
               "block_bwd: rows=%lld rows_per_sample=%d", (long long)rows, rows_per_sample);
  CCDM_REQUIRE(C > 0 && C % 8 == 0 && C <= kBwdMaxC, CCDM_ERR_UNSUPPORTED_SHAPE,
               "block_bwd: C=%d must be a multiple of 8 up to %d", C, kBwdMaxC);
  CCDM_REQUIRE(!(flags & CCDM_EPI_SS) || scale_shift, CCDM_ERR_BAD_ARG, "block_bwd: scale/shift flag without pointer");
  CCDM_REQUIRE((flags & ~(CCDM_EPI_SS | CCDM_EPI_SILU)) == 0, CCDM_ERR_BAD_ARG, "block_bwd: flags 0x%x", flags);
  const int B = (int)(rows / rows_per_sample);
  cudaStream_t s = (cudaStream_t)stream;
  int kv, G;
  row_lane_plan(C / 8, 4, &kv, &G);
  CCDM_REQUIRE(B <= 65535, CCDM_ERR_UNSUPPORTED_SHAPE, "block_bwd: %d samples", B);
  // Slabs of rows of one sample per CTA.  The number of slabs per sample minimises (waves of resident CTAs) x (row-loop
  // iterations per CTA + ~3 iterations' worth of prologue / column-sum epilogue): the first heuristic (3 CTAs per SM) left
  // 2-3 iterations per CTA at 16x16 / 8x8 resolutions, all overhead.
  const int slots = num_sms() * (kv <= 2 ? 2 : 1);
  const int rstep = (kBwdThreads / 32) * (32 / G);
  int max_split = (rows_per_sample + 63) / 64;
  if (max_split > 64) max_split = 64;
  int per_sample = 1;
  long long best_cost = -1;
  for (int p = 1; p <= max_split; ++p) {
    const int rpb = (rows_per_sample + p - 1) / p;
    const long long waves = ((long long)B * ((rows_per_sample + rpb - 1) / rpb) + slots - 1) / slots;
    const long long cost = waves * ((rpb + rstep - 1) / rstep + 3);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      per_sample = p;
    }
  }
  const int rows_per_block = (rows_per_sample + per_sample - 1) / per_sample;
  per_sample = (rows_per_sample + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)per_sample, (unsigned)B);
  const bool silu = (flags & CCDM_EPI_SILU) != 0;
#define CCDM_BWD(K)                                                                                                   \
  do {                                                                                                                \
    constexpr int ring_bytes = bwd_stages(K) * K * 2 * kBwdThreads * 16;                                              \
    constexpr int min_ctas = K <= 2 ? 2 : 1;                                                                          \
    static bool attr_done = false;                                                                                    \
    if (!attr_done) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(block_bwd_kernel<K, min_ctas, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           ring_bytes);                                                               \
      if (e == cudaSuccess)                                                                                           \
        e = cudaFuncSetAttribute(block_bwd_kernel<K, min_ctas, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 ring_bytes);                                                                         \
      if (e != cudaSuccess) return cuda_fail(e, "block_bwd: cudaFuncSetAttribute");                                   \
      attr_done = true;                                                                                               \
    }                                                                                                                 \
    if (silu)                                                                                                         \
      block_bwd_kernel<K, min_ctas, true><<<grid, kBwdThreads, ring_bytes, s>>>(                                      \
          (const uint4*)dy, (const uint4*)z, (uint4*)dz, rows_per_sample, rows_per_block, C, gain, gain_mul, scale_shift, \
          ss_ld, ss_off, sums, B, flags, G);                                                                          \
    else                                                                                                              \
      block_bwd_kernel<K, min_ctas, false><<<grid, kBwdThreads, ring_bytes, s>>>(                                     \
          (const uint4*)dy, (const uint4*)z, (uint4*)dz, rows_per_sample, rows_per_block, C, gain, gain_mul, scale_shift, \
          ss_ld, ss_off, sums, B, flags, G);                                                                          \
  } while (0)
  if (kv == 1) CCDM_BWD(1);
  else if (kv == 2) CCDM_BWD(2);
  else if (kv == 3) CCDM_BWD(3);
  else CCDM_BWD(4);
#undef CCDM_BWD
  return after_launch("block_bwd_kernel");
}

extern "C" int ccdm_block_bwd_finish(const float* sums, int32_t B, int32_t C, const float* gain, float gain_mul,
                                     const float* scale_shift, int32_t ss_ld, int32_t ss_off, float* d_ss, float* dgain,
                                     float* dbias, void* stream) {
  CCDM_REQUIRE(sums && gain && B > 0 && C > 0, CCDM_ERR_BAD_ARG, "block_bwd_finish: bad args");
  CCDM_REQUIRE(!d_ss || scale_shift, CCDM_ERR_BAD_ARG, "block_bwd_finish: d_ss without scale_shift");
  block_bwd_finish_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(sums, B, C, gain, gain_mul, scale_shift, ss_ld,
                                                                            ss_off, d_ss, dgain, dbias);
  return after_launch("block_bwd_finish_kernel");
}
