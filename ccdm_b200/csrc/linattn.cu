// Linear-attention context on tensor cores (tcgen05), unet.py:208,212:
//
//   context[b,h,d,e] = sum_n softmax_n(k)[b,h,d,n] * v[b,h,e,n]
//
// The qkv tap-GEMM epilogue already wrote p = exp(k - bound_d) (bound_d >= max_n k[d,n], see ccdm_kexp_bound), so the
// softmax is sum_n p v / sum_n p and both sums are GEMMs over the token axis:
//
//   D[128 x 128] = P^T[128 x n] . V[n x 128]       (all four heads at once; the diagonal 32x32 blocks are the contexts)
//   S[128 x 16]  = P^T[128 x n] . 1[n x 16]        (column sums, same A operand, B = a constant tile of ones)
//
// P and V are token-major in memory (channels contiguous), i.e. "MN-major" operands for UMMA: TMA boxes of
// {64 channels, 64 tokens} land as 128-byte rows per token in the SWIZZLE_128B layout, one 8 KiB block per 64-channel
// half; the descriptors use LBO = block stride, SBO = 1024 B (8 tokens) and the instruction descriptor sets the
// a_major / b_major bits.  One persistent CTA streams whole samples: the kernel is HBM-bound (reads k and v once).
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                    const cuuint32_t* box);

constexpr int kCtxTok = 64;                       // tokens per pipeline stage
constexpr int kCtxStages = 4;
constexpr int kCtxBlock = kCtxTok * 128;          // one {64 ch x 64 tok} box: 8 KiB
constexpr int kCtxStageBytes = 4 * kCtxBlock;     // P lo, P hi, V lo, V hi
constexpr int kCtxThreads = 192;

struct CtxAux {
  uint64_t full[kCtxStages], empty[kCtxStages], acc_full, acc_empty;
  uint32_t tmem_slot;
};

// The same kernel computes the backward token-axis product dctx = Q^T . dOut (unet.py:214 transposed): the A operand
// then comes from channels [a_ch0, a_ch0+128) of `a_map`, B from channels [b_ch0, b_ch0+128) of `b_map`, and the
// result is left un-normalised (normalize == 0).  `sums_out` ([B][128]) receives the column sums of A.
__global__ void __launch_bounds__(kCtxThreads, 1) linattn_context_mma_kernel(const __grid_constant__ CUtensorMap a_map,
                                                                             const __grid_constant__ CUtensorMap b_map,
                                                                             int a_ch0, int b_ch0, int normalize,
                                                                             float* __restrict__ ctx,
                                                                             float* __restrict__ sums_out, int B, int n,
                                                                             const float* __restrict__ w_out,
                                                                             __nv_bfloat16* __restrict__ wfold, int C,
                                                                             int n_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* ones = smem + kCtxStages * kCtxStageBytes;              // [64 tok][128 B] of bf16 1.0
  CtxAux* aux = reinterpret_cast<CtxAux*>(ones + kCtxBlock);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a_map);
    tma_prefetch_desc(&b_map);
  }
  if (warp == 1) tmem_alloc(&aux->tmem_slot, 256);
  if (tid == 64) {
    for (int s = 0; s < kCtxStages; ++s) {
      mbar_init(&aux->full[s], 1);
      mbar_init(&aux->empty[s], 1);
    }
    mbar_init(&aux->acc_full, 1);
    mbar_init(&aux->acc_empty, 4);
    fence_mbar_init();
  }
  for (int i = tid; i < kCtxBlock / 4; i += kCtxThreads) reinterpret_cast<uint32_t*>(ones)[i] = CCDM_ONE_PAIR;
  fence_proxy_async_smem();                                        // generic-proxy writes -> visible to the MMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_slot;
  const int nkt = (n + kCtxTok - 1) / kCtxTok;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    uint32_t it = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      for (int kt = 0; kt < nkt; ++kt, ++it) {
        const int s = it % kCtxStages;
        mbar_wait(&aux->empty[s], ((it / kCtxStages) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&aux->full[s], kCtxStageBytes);
          uint8_t* st = ring + s * kCtxStageBytes;
          // channels: q [0,128) | p = exp(k - bound) [128,256) | v [256,384); tokens past n are zero-filled
          for (int j = 0; j < 4; ++j)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(st + j * kCtxBlock)), "l"(reinterpret_cast<uint64_t>(j < 2 ? &a_map : &b_map)),
                "r"(smem_u32(&aux->full[s])), "r"(j < 2 ? a_ch0 + j * 64 : b_ch0 + (j - 2) * 64), "r"(kt * kCtxTok), "r"(b)
                : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc_ctx = umma_idesc_bf16_mn(128, 128);
    const uint32_t idesc_sum = umma_idesc_bf16_mn(128, 16);
    const uint32_t ones_addr = smem_u32(ones);
    uint32_t it = 0;
    int ls = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++ls) {
      mbar_wait(&aux->acc_empty, (ls & 1) ^ 1u);
      tc_fence_after();
      for (int kt = 0; kt < nkt; ++kt, ++it) {
        const int s = it % kCtxStages;
        mbar_wait(&aux->full[s], (it / kCtxStages) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(ring + s * kCtxStageBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kCtxTok / 16; ++k) {
            const uint64_t adesc = umma_desc_mn_sw128(base + k * 2048, kCtxBlock);                   // P^T
            const uint64_t bdesc = umma_desc_mn_sw128(base + 2 * kCtxBlock + k * 2048, kCtxBlock);   // V
            const uint64_t odesc = umma_desc_mn_sw128(ones_addr + k * 2048, kCtxBlock);
            const uint32_t acc = (kt | k) != 0 ? 1u : 0u;
            umma_bf16_ss(tmem_base, adesc, bdesc, idesc_ctx, acc);
            umma_bf16_ss(tmem_base + 128, adesc, odesc, idesc_sum, acc);
          }
          umma_commit(&aux->empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&aux->acc_full);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue: head h == TMEM lane quarter h
    const int h = warp & 3;
    const int d = lane;
    uint32_t r[32], sum[32];
    int ls = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++ls) {
      mbar_wait(&aux->acc_full, ls & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(h * 32) << 16);
      tmem_ld32(trow + h * 32, r);                                 // D[h*32+d][h*32 .. h*32+32)
      tmem_ld32(trow + 128, sum);                                  // column 128: sum_n p[n, h*32+d]
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux->acc_empty);
      const float inv = normalize ? 1.f / __uint_as_float(sum[0]) : 1.f;
      if (sums_out) sums_out[static_cast<long long>(b) * 128 + h * 32 + d] = __uint_as_float(sum[0]);
      float cr[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) cr[j] = __uint_as_float(r[j]) * inv;
      if (ctx) {
        float4* o = reinterpret_cast<float4*>(ctx + ((static_cast<long long>(b) * 4 + h) * 32 + d) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(cr[4 * j], cr[4 * j + 1], cr[4 * j + 2], cr[4 * j + 3]);
      }
      if (wfold) {
        // fold the context into the output projection (unet.py:214-216 + to_out[0]):
        //   wfold[b][c][h*32+d] = sum_e w_out[c][h*32+e] * ctx[b][h][d][e];  all lanes of a warp read the same w_out row
        __nv_bfloat16* wf = wfold + static_cast<long long>(b) * n_rows * 128 + h * 32 + d;
        for (int c = 0; c < n_rows; ++c) {
          float acc = 0.f;
          if (c < C) {
            const float4* wr = reinterpret_cast<const float4*>(w_out + static_cast<long long>(c) * 128 + h * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 wv = __ldg(wr + j);
              acc = fmaf(wv.x, cr[4 * j], acc);
              acc = fmaf(wv.y, cr[4 * j + 1], acc);
              acc = fmaf(wv.z, cr[4 * j + 2], acc);
              acc = fmaf(wv.w, cr[4 * j + 3], acc);
            }
          }
          wf[static_cast<long long>(c) * 128] = __float2bfloat16(acc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Generic-head-count variant on CUDA cores (p already exponentiated): one CTA per (sample, head).
__global__ void __launch_bounds__(128) linattn_context_simt_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   float* __restrict__ ctx, int n, int heads) {
  __shared__ float ps[64][33];
  __shared__ __align__(16) float vs[64][32];
  const int t = threadIdx.x;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * 32;
  const __nv_bfloat16* pbase = qkv + (long long)b * n * ld + heads * 32 + h * 32;
  const __nv_bfloat16* vbase = pbase + heads * 32;
  const int d = t >> 2, eg = t & 3;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float ssum = 0.f;
  for (int r0 = 0; r0 < n; r0 += 64) {
    const int rows = min(64, n - r0);
    for (int i = 0; i < 4; ++i) {
      const int row = (t >> 3) + 16 * i, c4 = (t & 7) * 4;
      float p4[4] = {0, 0, 0, 0}, v4[4] = {0, 0, 0, 0};
      if (row < rows) {
        const uint2 pu = __ldg(reinterpret_cast<const uint2*>(pbase + (long long)(r0 + row) * ld + c4));
        const uint2 vu = __ldg(reinterpret_cast<const uint2*>(vbase + (long long)(r0 + row) * ld + c4));
        p4[0] = bf16_lo(pu.x); p4[1] = bf16_hi(pu.x); p4[2] = bf16_lo(pu.y); p4[3] = bf16_hi(pu.y);
        v4[0] = bf16_lo(vu.x); v4[1] = bf16_hi(vu.x); v4[2] = bf16_lo(vu.y); v4[3] = bf16_hi(vu.y);
      }
      for (int j = 0; j < 4; ++j) {
        ps[row][c4 + j] = p4[j];
        vs[row][c4 + j] = v4[j];
      }
    }
    __syncthreads();
    for (int row = 0; row < rows; ++row) {
      const float pk = ps[row][d];
      const float4 va = *reinterpret_cast<const float4*>(&vs[row][eg * 8]);
      const float4 vb = *reinterpret_cast<const float4*>(&vs[row][eg * 8 + 4]);
      acc[0] = fmaf(pk, va.x, acc[0]); acc[1] = fmaf(pk, va.y, acc[1]);
      acc[2] = fmaf(pk, va.z, acc[2]); acc[3] = fmaf(pk, va.w, acc[3]);
      acc[4] = fmaf(pk, vb.x, acc[4]); acc[5] = fmaf(pk, vb.y, acc[5]);
      acc[6] = fmaf(pk, vb.z, acc[6]); acc[7] = fmaf(pk, vb.w, acc[7]);
      ssum += pk;
    }
    __syncthreads();
  }
  const float inv = 1.f / ssum;
  float* o = ctx + ((long long)blockIdx.x * 32 + d) * 32 + eg * 8;
  for (int j = 0; j < 8; ++j) o[j] = acc[j] * inv;
}

// bias[n] = -slack * ||wpacked[n, :]||_2 for n in [lo, hi), 0 elsewhere: the softmax shift of the k columns.
// |k[n,d]| <= ||W'_d|| because the PreNorm'd input row has unit length, so exp(k - bound) never overflows.
__global__ void kexp_bound_kernel(const __nv_bfloat16* __restrict__ wp, int n_rows, int K, int lo, int hi,
                                  float* __restrict__ bias) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float s = 0.f;
  if (row >= lo && row < hi)
    for (int k = lane; k < K; k += 32) {
      const float w = __bfloat162float(wp[(long long)row * K + k]);
      s = fmaf(w, w, s);
    }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  // binary16 build: p = exp(k - bound) would underflow where bfloat16 (fp32's exponent range) does not; a constant shift of
  // +10 (p <= e^10 = 22026 < 65504) moves the numerators into binary16's normal range and cancels in sum(p v) / sum(p)
  if (lane == 0) bias[row] = (row >= lo && row < hi) ? -1.01f * sqrtf(s) - 1e-3f + CCDM_KEXP_SHIFT : 0.f;
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_linattn_fold(const float* w_out, const float* ctx, void* wfold, int32_t B, int32_t C,
                                 int32_t n_rows, int32_t heads, void* stream);

static int launch_token_gemm(const void* a, int a_ld, int a_ch0, const void* bsrc, int b_ld, int b_ch0, int normalize,
                             float* ctx, float* sums_out, int B, int n, const float* w_out, void* wfold, int C, int n_rows,
                             cudaStream_t s) {
  CUtensorMap amap, bmap;
  cuuint32_t box[3] = {64, kCtxTok, 1};
  {
    cuuint64_t dims[3] = {(cuuint64_t)a_ld, (cuuint64_t)n, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)a_ld * 2, (cuuint64_t)n * a_ld * 2};
    int rc = encode_map_bf16(&amap, a, 3, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)b_ld, (cuuint64_t)n, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)b_ld * 2, (cuuint64_t)n * b_ld * 2};
    int rc = encode_map_bf16(&bmap, bsrc, 3, dims, str, box);
    if (rc != CCDM_OK) return rc;
  }
  const size_t smem = kCtxStages * kCtxStageBytes + kCtxBlock + sizeof(CtxAux) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linattn_context_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "linattn_context: cudaFuncSetAttribute");
    attr_set = true;
  }
  int grid = num_sms();
  if (grid > B) grid = B;
  linattn_context_mma_kernel<<<grid, kCtxThreads, smem, s>>>(amap, bmap, a_ch0, b_ch0, normalize, ctx, sums_out, B, n, w_out,
                                                             (__nv_bfloat16*)wfold, C, n_rows);
  return after_launch("linattn_context_mma_kernel");
}

extern "C" int ccdm_linattn_dcontext(const void* qkv, const void* dout, float* dctx, int32_t B, int32_t n, void* stream) {
  CCDM_REQUIRE(qkv && dout && dctx && B > 0 && n > 0, CCDM_ERR_BAD_ARG, "linattn_dcontext: bad args");
  return launch_token_gemm(qkv, 384, 0, dout, 128, 0, 0, dctx, nullptr, B, n, nullptr, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int ccdm_linattn_context(const void* qkv, float* ctx, float* colsum, int32_t B, int32_t n, int32_t heads,
                                    const float* w_out, void* wfold, int32_t C, int32_t n_rows, void* stream) {
  CCDM_REQUIRE(qkv && (ctx || wfold) && B > 0 && n > 0 && heads > 0, CCDM_ERR_BAD_ARG, "linattn_context: bad args");
  CCDM_REQUIRE(!wfold || (w_out && C > 0 && n_rows >= C), CCDM_ERR_BAD_ARG, "linattn_context: fold needs w_out, C, n_rows");
  cudaStream_t s = (cudaStream_t)stream;
  if (heads != 4) {
    CCDM_REQUIRE(ctx && !colsum, CCDM_ERR_BAD_ARG, "linattn_context: heads != 4 needs the ctx buffer and has no colsum output");
    linattn_context_simt_kernel<<<B * heads, 128, 0, s>>>((const __nv_bfloat16*)qkv, ctx, n, heads);
    int rc = after_launch("linattn_context_simt_kernel");
    if (rc != CCDM_OK || !wfold) return rc;
    return ccdm_linattn_fold(w_out, ctx, wfold, B, C, n_rows, heads, stream);
  }
  return launch_token_gemm(qkv, 384, 128, qkv, 384, 256, 1, ctx, colsum, B, n, w_out, wfold, C, n_rows, s);
}

extern "C" int ccdm_kexp_bound(const void* wpacked, int32_t n_rows, int32_t K, int32_t row_lo, int32_t row_hi,
                               float* bias, void* stream) {
  CCDM_REQUIRE(wpacked && bias && n_rows > 0 && K > 0 && row_lo >= 0 && row_hi <= n_rows, CCDM_ERR_BAD_ARG,
               "kexp_bound: bad args");
  kexp_bound_kernel<<<(n_rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)wpacked, n_rows, K, row_lo,
                                                                        row_hi, bias);
  return after_launch("kexp_bound_kernel");
}
