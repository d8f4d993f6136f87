// Fused clip + Adam over flat gradient / moment buffers (replaces
// torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step of CCDM_unified/trainer.py:137,724,733-734), and a
// multi-tensor lerp for the EMA (ema_pytorch.py:150-178).  HBM-bound: 28 bytes per parameter per step.
#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

// sumsq[0] += sum x^2 (double accumulation per block, one atomic per block)
__global__ void __launch_bounds__(512) sumsq_kernel(const float4* __restrict__ x, long long n4, const float* tail, int ntail,
                                                    double* __restrict__ out) {
  __shared__ double red[16];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) acc += (double)tail[threadIdx.x] * tail[threadIdx.x];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

__global__ void adam_begin_kernel(float* step, double* sumsq) {
  step[0] += 1.f;
  if (sumsq) sumsq[0] = 0.0;
}

// One CTA per chunk of one parameter tensor; g / m / v live in flat buffers at the same offset.
__global__ void __launch_bounds__(256) fused_adam_kernel(float* const* __restrict__ chunk_param,
                                                         const long long* __restrict__ chunk_off,
                                                         const int* __restrict__ chunk_n, const float* __restrict__ g_flat,
                                                         float* __restrict__ m_flat, float* __restrict__ v_flat,
                                                         const float* __restrict__ step, const double* __restrict__ sumsq,
                                                         float max_norm, float lr, float b1, float b2, float eps,
                                                         float weight_decay) {
  const int ci = blockIdx.x;
  float* __restrict__ p = chunk_param[ci];
  const long long off = chunk_off[ci];
  const int n = chunk_n[ci];
  const float t = step[0];
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  float coef = 1.f;                                               // clip_grad_norm_: min(1, max_norm / (norm + 1e-6))
  if (sumsq) coef = fminf(1.f, max_norm / ((float)sqrt(sumsq[0]) + 1e-6f));
  const float* g = g_flat + off;
  float* m = m_flat + off;
  float* v = v_flat + off;
  for (int i = threadIdx.x; i < n; i += 256) {
    float gi = g[i] * coef;
    float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// dst[i] += w * (src[i] - dst[i]) per chunk (EMA), w read from the device
__global__ void __launch_bounds__(256) multi_lerp_kernel(float* const* __restrict__ dst, const float* const* __restrict__ src,
                                                         const int* __restrict__ chunk_n, const float* __restrict__ w) {
  const int ci = blockIdx.x;
  float* __restrict__ d = dst[ci];
  const float* __restrict__ s = src[ci];
  const int n = chunk_n[ci];
  const float wt = w[0];
  for (int i = threadIdx.x; i < n; i += 256) d[i] = fmaf(wt, s[i] - d[i], d[i]);
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_fused_adam(void* const* chunk_param, const int64_t* chunk_off, const int32_t* chunk_n,
                               int32_t n_chunks, const float* g_flat, float* m_flat, float* v_flat, int64_t n_flat,
                               float* step, double* sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                               float weight_decay, void* stream) {
  CCDM_REQUIRE(chunk_param && chunk_off && chunk_n && g_flat && m_flat && v_flat && step && n_chunks > 0 && n_flat > 0,
               CCDM_ERR_BAD_ARG, "fused_adam: bad args");
  CCDM_REQUIRE((reinterpret_cast<uintptr_t>(g_flat) & 15) == 0, CCDM_ERR_BAD_ARG, "fused_adam: gradient buffer alignment");
  cudaStream_t s = (cudaStream_t)stream;
  adam_begin_kernel<<<1, 1, 0, s>>>(step, sumsq);
  int rc = after_launch("adam_begin_kernel");
  if (rc != CCDM_OK) return rc;
  if (sumsq) {
    const long long n4 = n_flat / 4;
    long long blocks = (n4 + 512 * 8 - 1) / (512 * 8);
    if (blocks > num_sms() * 4) blocks = num_sms() * 4;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<(unsigned)blocks, 512, 0, s>>>((const float4*)g_flat, n4, g_flat + n4 * 4, (int)(n_flat - n4 * 4), sumsq);
    rc = after_launch("sumsq_kernel");
    if (rc != CCDM_OK) return rc;
  }
  fused_adam_kernel<<<n_chunks, 256, 0, s>>>((float* const*)chunk_param, (const long long*)chunk_off, chunk_n, g_flat, m_flat,
                                             v_flat, step, sumsq, max_norm, lr, beta1, beta2, eps, weight_decay);
  return after_launch("fused_adam_kernel");
}

extern "C" int ccdm_multi_lerp(void* const* dst, const void* const* src, const int32_t* chunk_n, int32_t n_chunks,
                               const float* weight, void* stream) {
  CCDM_REQUIRE(dst && src && chunk_n && weight && n_chunks > 0, CCDM_ERR_BAD_ARG, "multi_lerp: bad args");
  multi_lerp_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((float* const*)dst, (const float* const*)src, chunk_n, weight);
  return after_launch("multi_lerp_kernel");
}
