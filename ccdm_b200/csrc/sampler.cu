// Guidance + sampler-step kernel and the training-side elementwise kernels (q_sample, vicinal loss).
// All HBM-bound: one pass (or a few L2-resident passes) over [B][C*H*W] fp32 state.
#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Sum K doubles across the block; result valid in every thread.  blockDim.x <= 1024.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch /* [32*K] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < K; ++k) scratch[warp * K + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[w * K + k];
    v[k] = s;
  }
}

// Coefficients of "guided = A*cond + Bn*null" before the std rescale, from per-sample reductions
// (unet.py:51-62, 365-371):  update = cond - null;  par = <update,unit> unit, unit = cond/max(|cond|,1e-12);
// update' = (update - par) + par*kpf;  scaled = cond + update' * (s-1).
struct CfgCoef {
  float a_cond, a_null;
};
__device__ __forceinline__ CfgCoef cfg_coef(double uc, double cc, float cond_scale, int remove_parallel, float kpf) {
  // scaled = cond + (s-1) * (u - (1-kpf) * par),  par = (uc / max(|c|,eps)^2) * c
  double pc = 0.0;
  if (remove_parallel) {
    const double nrm = fmax(sqrt(cc), 1e-12);
    pc = (1.0 - (double)kpf) * uc / (nrm * nrm);
  }
  const double s1 = (double)cond_scale - 1.0;
  CfgCoef r;
  r.a_cond = (float)(1.0 + s1 * (1.0 - pc));
  r.a_null = (float)(-s1);
  return r;
}

// One CTA per sample.  kStage: the sample's two network outputs are copied ONCE into shared memory (LDGSTS, no register
// round trip, every thread copies the vectors it reads itself so no block barrier is needed) and the state x_t is
// prefetched into registers while the reduction runs -- global memory is touched once: cond, null, x_t (+ noise) in,
// x_{t-1} (+ predictions) out.  64x64x3 samples need 96 KB, so two CTAs are resident per SM and a batch of up to 296
// samples is one wave.  !kStage: any chw; the second pass re-reads global memory (L2 hits).
// The guidance needs per-sample reductions (unet.py:51-62, 365-371): <cond - null, cond>, |cond|^2 and, for the std rescale,
// the unbiased std of cond and of the guided output.  All of them follow from ONE pass of second moments
// (sum c, n, c^2, n^2, cn accumulated in double: every product of two floats is exact there), instead of three dependent
// block reductions.
constexpr int kStepThreads = 512;
constexpr int kStepXPre = 6;                              // float4 of x_t prefetched per thread (64x64x3 / 4 / 512)

template <bool kStage>
__global__ void __launch_bounds__(kStepThreads, kStage ? 2 : 1) sampler_step_kernel(const ccdm_step_args a) {
  __shared__ double scratch[32 * 6];
  extern __shared__ float step_stage_f[];                  // kStage: [cond | null][chw]
  const int b = blockIdx.x;
  const long long base = (long long)b * a.chw;
  const float* cond = a.out_cond + base;
  const float* nul = a.out_null ? a.out_null + base : nullptr;
  const int nv = a.chw >> 2;                               // (kStage: chw % 4 == 0, checked by the launcher)
  float* x = a.x ? a.x + base : nullptr;
  float4* const sc4 = reinterpret_cast<float4*>(step_stage_f);
  float4* const sn4 = sc4 + nv;
  float4 xpre[kStepXPre];
  if constexpr (kStage) {
    for (int i4 = threadIdx.x; i4 < nv; i4 += kStepThreads) {
      cp_async16(sc4 + i4, reinterpret_cast<const float4*>(cond) + i4, true);
      if (nul) cp_async16(sn4 + i4, reinterpret_cast<const float4*>(nul) + i4, true);
    }
    cp_async_commit();
#pragma unroll
    for (int j = 0; j < kStepXPre; ++j) {
      const int i4 = threadIdx.x + j * kStepThreads;
      xpre[j] = (x && i4 < nv) ? *(reinterpret_cast<const float4*>(x) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait<0>();
  }
  const long long step = a.t_rows ? a.t_rows[b] : (a.step_counter ? (long long)*a.step_counter : 0);
  const float* cf = a.x ? a.coef + step * CCDM_STEP_NCOEF : nullptr;

  float a_cond = 1.f, a_null = 0.f, resc = 1.f;
  if (nul) {
    double v[6] = {0, 0, 0, 0, 0, 0};  // <u,c>, sum c^2, sum c, sum n, sum n^2, sum cn
    auto acc = [&](float cf_, float nf_) {
      const double c = cf_, n = nf_;
      v[0] += (c - n) * c;
      v[1] += c * c;
      v[2] += c;
      v[3] += n;
      v[4] += n * n;
      v[5] += c * n;
    };
    if constexpr (kStage) {
      for (int i4 = threadIdx.x; i4 < nv; i4 += kStepThreads) {
        const float4 c = sc4[i4], n = sn4[i4];
        acc(c.x, n.x);
        acc(c.y, n.y);
        acc(c.z, n.z);
        acc(c.w, n.w);
      }
    } else {
      for (int i = threadIdx.x; i < a.chw; i += blockDim.x) acc(cond[i], nul[i]);
    }
    block_sum<6>(v, scratch);
    const CfgCoef k = cfg_coef(v[0], v[1], a.cond_scale, a.remove_parallel, a.keep_parallel_frac);
    a_cond = k.a_cond;
    a_null = k.a_null;
    if (a.rescaled_phi != 0.f) {
      // unbiased std of cond and of scaled = a_cond * cond + a_null * null (torch.std default) from the moments
      const double N = (double)a.chw, ac = (double)a_cond, an = (double)a_null;
      const double var_c = fmax(v[1] - v[2] * v[2] / N, 0.0) / (N - 1.0);
      const double sum_s = ac * v[2] + an * v[3];
      const double sq_s = ac * ac * v[1] + 2.0 * ac * an * v[5] + an * an * v[4];
      const double var_s = fmax(sq_s - sum_s * sum_s / N, 0.0) / (N - 1.0);
      resc = (float)(sqrt(var_c) / sqrt(var_s)) * a.rescaled_phi + (1.f - a.rescaled_phi);
    }
  }

  const float c_recip = x ? cf[0] : 0.f, c_recipm1 = x ? cf[1] : 1.f, c_sa = x ? cf[2] : 0.f, c_s1m = x ? cf[3] : 0.f;
  // one element of the update: returns x_{t-1} (or the guided output when there is no state); may write the predictions
  auto elem = [&](int i, float c, float n, float xt, float nz) -> float {
    const float g = (a_cond * c + a_null * n) * resc;  // guided model output
    if (!x) return g;                                  // guidance only (ccdm_cfg_combine)
    float x0, eps;
    if (a.objective == CCDM_OBJ_PRED_NOISE) {
      x0 = c_recip * xt - c_recipm1 * g;
      if (a.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      eps = a.cfg_plus_plus ? n : g;
    } else if (a.objective == CCDM_OBJ_PRED_X0) {
      x0 = a.clip_x0 ? fminf(fmaxf(g, -1.f), 1.f) : g;
      float xe = x0;
      if (a.cfg_plus_plus) xe = a.clip_x0 ? fminf(fmaxf(n, -1.f), 1.f) : n;
      eps = (c_recip * xt - xe) / c_recipm1;
    } else {
      x0 = c_sa * xt - c_s1m * g;
      if (a.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      float xe = x0;
      if (a.cfg_plus_plus) {
        xe = c_sa * xt - c_s1m * n;
        if (a.clip_x0) xe = fminf(fmaxf(xe, -1.f), 1.f);
      }
      eps = (c_recip * xt - xe) / c_recipm1;
    }
    if (a.pred_noise) a.pred_noise[base + i] = eps;
    if (a.pred_x0) a.pred_x0[base + i] = x0;
    if (a.sampler == 2) return xt;  // predictions only: the state is left as it is
    float xn;
    if (a.sampler == 0) {  // DDIM, diffusion.py:450-464
      if (cf[7] != 0.f) {
        xn = x0;
      } else {
        xn = x0 * cf[4] + cf[5] * eps;
        if (cf[6] != 0.f) xn += cf[6] * nz;
      }
    } else {               // DDPM, diffusion.py:344-373 (x0 always clamped there)
      const float x0c = fminf(fmaxf(x0, -1.f), 1.f);
      xn = cf[8] * x0c + cf[9] * xt;
      if (cf[10] != 0.f) xn += cf[10] * nz;
    }
    return xn;
  };
  const bool need_noise = x && a.sampler != 2 && ((a.sampler == 0 && cf[7] == 0.f && cf[6] != 0.f) ||
                                                  (a.sampler == 1 && cf[10] != 0.f));
  const float* nzp = need_noise ? a.noise + base : nullptr;
  if constexpr (kStage) {
    float* dst = x ? x : a.pred_x0 + base;               // guidance only: the guided output goes to pred_x0
    auto one = [&](int i4, float4 xt) {
      const float4 c = sc4[i4], n = nul ? sn4[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 nz = nzp ? __ldg(reinterpret_cast<const float4*>(nzp) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = elem(4 * i4 + 0, c.x, n.x, xt.x, nz.x);
      o.y = elem(4 * i4 + 1, c.y, n.y, xt.y, nz.y);
      o.z = elem(4 * i4 + 2, c.z, n.z, xt.z, nz.z);
      o.w = elem(4 * i4 + 3, c.w, n.w, xt.w, nz.w);
      if (!x || a.sampler != 2) *(reinterpret_cast<float4*>(dst) + i4) = o;
    };
#pragma unroll
    for (int j = 0; j < kStepXPre; ++j) {
      const int i4 = threadIdx.x + j * kStepThreads;
      if (i4 < nv) one(i4, xpre[j]);
    }
    for (int i4 = threadIdx.x + kStepXPre * kStepThreads; i4 < nv; i4 += kStepThreads)
      one(i4, x ? *(reinterpret_cast<const float4*>(x) + i4) : make_float4(0.f, 0.f, 0.f, 0.f));
  } else {
    for (int i = threadIdx.x; i < a.chw; i += blockDim.x) {
      const float r = elem(i, cond[i], nul ? nul[i] : 0.f, x ? x[i] : 0.f, nzp ? nzp[i] : 0.f);
      if (!x) a.pred_x0[base + i] = r;
      else if (a.sampler != 2) x[i] = r;
    }
  }
}

// Staged variant for samples whose two network outputs fit the shared memory of one CTA (chw * 8 bytes <= 200 KB, i.e.
// up to 80x80x4); needs 16-byte aligned rows.  Larger samples (128x128 and up) take the streaming variant.
static int launch_sampler_step(const ccdm_step_args& k, cudaStream_t s) {
  const bool vec = (k.chw % 4 == 0) && ((reinterpret_cast<uintptr_t>(k.out_cond) & 15) == 0) &&
                   (!k.out_null || (reinterpret_cast<uintptr_t>(k.out_null) & 15) == 0) &&
                   (!k.x || (reinterpret_cast<uintptr_t>(k.x) & 15) == 0) &&
                   (!k.noise || (reinterpret_cast<uintptr_t>(k.noise) & 15) == 0) &&
                   (k.x || (reinterpret_cast<uintptr_t>(k.pred_x0) & 15) == 0);
  const size_t stage_bytes = (size_t)k.chw * 8;
  if (vec && stage_bytes <= 200 * 1024) {
    static size_t attr_bytes = 0;
    if (stage_bytes > attr_bytes) {
      cudaError_t e = cudaFuncSetAttribute(sampler_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "sampler_step: cudaFuncSetAttribute");
      attr_bytes = 200 * 1024;
    }
    sampler_step_kernel<true><<<k.B, kStepThreads, stage_bytes, s>>>(k);
  } else {
    sampler_step_kernel<false><<<k.B, kStepThreads, 0, s>>>(k);
  }
  return CCDM_OK;
}

// The step counter is bumped by its own 1-thread launch after the step kernel, so every CTA of the step kernel
// has read the old value (stream order) before it changes.
__global__ void advance_counter_kernel(int* c) { *c += 1; }

__global__ void broadcast_step_kernel(const long long* __restrict__ table, const int* __restrict__ counter,
                                      long long* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = table[*counter];
}

// ---------------------------------------------------------------------------- q_sample
__global__ void q_sample_kernel(const ccdm_qsample_args a) {
  const long long total = (long long)a.B * a.chw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.chw);
    const long long t = a.t[b];
    const float x0 = a.normalize ? a.img01[i] * 2.f - 1.f : a.img01[i];
    float nz = a.noise[i];
    if (a.cov) {
      if (a.keep[b]) nz *= sqrtf(a.cov[i]);
      else if (a.noise2) nz = a.noise2[i];
    }
    a.x0[i] = x0;
    a.noise_out[i] = nz;
    a.x_t[i] = a.sqrt_acp[t] * x0 + a.sqrt_1m_acp[t] * nz;
  }
}

// ---------------------------------------------------------------------------- vicinal loss
// per_sample[b] = loss_weight[t_b] * sum_i (out - target)^2 / cov      (cov = 1 on null rows / without Hy)
__global__ void __launch_bounds__(512) loss_rows_kernel(const ccdm_loss_args a) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const long long base = (long long)b * a.chw;
  const long long t = a.t[b];
  const float sa = a.sqrt_acp[t], s1 = a.sqrt_1m_acp[t];
  const bool use_cov = a.cov != nullptr && a.keep[b];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < a.chw; i += blockDim.x) {
    const float x0 = a.x0[base + i], nz = a.noise[base + i];
    const float target = a.objective == CCDM_OBJ_PRED_NOISE ? nz : (a.objective == CCDM_OBJ_PRED_X0 ? x0 : sa * nz - s1 * x0);
    const float d = a.model_out[base + i] - target;
    float e = d * d;
    if (use_cov) e /= a.cov[base + i];
    v[0] += e;
  }
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) a.per_sample[b] = a.loss_weight[t] * (float)v[0];
}

__global__ void loss_final_kernel(const ccdm_loss_args a) {
  __shared__ double scratch[32];
  double v[1] = {0.0};
  for (int b = threadIdx.x; b < a.B; b += blockDim.x)
    v[0] += (double)a.per_sample[b] * (a.row_weight ? (double)a.row_weight[b] : 1.0);
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) a.loss[0] = (float)(v[0] / ((double)a.B * a.chw));
}

__global__ void loss_grad_kernel(const ccdm_loss_args a) {
  const long long total = (long long)a.B * a.chw;
  const float norm = 2.f / ((float)a.B * (float)a.chw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.chw);
    const long long t = a.t[b];
    const float x0 = a.x0[i], nz = a.noise[i];
    const float target = a.objective == CCDM_OBJ_PRED_NOISE
                             ? nz
                             : (a.objective == CCDM_OBJ_PRED_X0 ? x0 : a.sqrt_acp[t] * nz - a.sqrt_1m_acp[t] * x0);
    float g = (a.model_out[i] - target) * norm * a.loss_weight[t] * (a.row_weight ? a.row_weight[b] : 1.f);
    if (a.cov && a.keep[b]) g /= a.cov[i];
    a.grad_out[i] = g;
  }
}

// In-batch vicinal weights; one thread per row i, loops over j.
__global__ void vicinal_weights_kernel(const float* __restrict__ proj, int B, int P, int euclid, int hard,
                                       const float* __restrict__ thr, float nu, const uint8_t* __restrict__ keep,
                                       float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float acc = 0.f;
  if (euclid) {
    for (int j = 0; j < B; ++j) {
      float d2 = 0.f;
      for (int p = 0; p < P; ++p) {
        const float d = proj[i * P + p] - proj[j * P + p];
        d2 += d * d;
      }
      const float dist = sqrtf(d2);
      acc += hard ? (dist <= thr[0] ? 1.f : 0.f) : expf(-nu * dist * dist);
    }
  } else {
    for (int p = 0; p < P; ++p) {
      float s = 0.f;
      for (int j = 0; j < B; ++j) {
        const float d = proj[i * P + p] - proj[j * P + p];
        s += hard ? (fabsf(d) <= thr[p] ? 1.f : 0.f) : expf(-nu * d * d);
      }
      acc += s / (float)P;
    }
  }
  acc /= (float)B;
  if (keep && !keep[i]) acc = 1.f;
  w[i] = acc;
}

}  // namespace ccdm

using namespace ccdm;

static int check_step(const ccdm_step_args* a) {
  CCDM_REQUIRE(a && a->out_cond && a->coef && a->B > 0 && a->chw > 1, CCDM_ERR_BAD_ARG, "sampler_step: bad args");
  CCDM_REQUIRE(a->noise || a->sampler != 1, CCDM_ERR_BAD_ARG, "sampler_step: the DDPM update needs a noise draw");
  CCDM_REQUIRE(a->objective >= 0 && a->objective <= 2 && a->sampler >= 0 && a->sampler <= 2, CCDM_ERR_BAD_ARG,
               "sampler_step: objective=%d sampler=%d", a->objective, a->sampler);
  CCDM_REQUIRE(a->out_null || a->cond_scale == 1.f, CCDM_ERR_BAD_ARG,
               "sampler_step: cond_scale != 1 needs the unconditional output");
  return CCDM_OK;
}

extern "C" int ccdm_sampler_step(const ccdm_step_args* a, void* stream) {
  int rc = check_step(a);
  if (rc != CCDM_OK) return rc;
  CCDM_REQUIRE(a->x, CCDM_ERR_BAD_ARG, "sampler_step: null state");
  ccdm_step_args k = *a;
  if (k.cond_scale == 1.f) k.out_null = nullptr;
  rc = launch_sampler_step(k, (cudaStream_t)stream);
  if (rc != CCDM_OK) return rc;
  rc = after_launch("sampler_step_kernel");
  if (rc != CCDM_OK) return rc;
  if (a->advance && a->step_counter) {
    advance_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a->step_counter);
    rc = after_launch("advance_counter_kernel");
  }
  return rc;
}

extern "C" int ccdm_broadcast_step_i64(const int64_t* table, const int32_t* step_counter, int64_t* out, int32_t n,
                                       void* stream) {
  CCDM_REQUIRE(table && step_counter && out && n > 0, CCDM_ERR_BAD_ARG, "broadcast_step: bad args");
  broadcast_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)table, step_counter,
                                                                           (long long*)out, n);
  return after_launch("broadcast_step_kernel");
}

extern "C" int ccdm_cfg_combine(const float* cond, const float* null_out, float* guided, int32_t B, int32_t chw,
                                float cond_scale, float rescaled_phi, int32_t remove_parallel,
                                float keep_parallel_frac, void* stream) {
  CCDM_REQUIRE(cond && null_out && guided && B > 0 && chw > 1, CCDM_ERR_BAD_ARG, "cfg_combine: bad args");
  ccdm_step_args k = {};
  k.out_cond = cond;
  k.out_null = null_out;
  k.x = nullptr;
  k.pred_x0 = guided;
  k.B = B;
  k.chw = chw;
  k.cond_scale = cond_scale;
  k.rescaled_phi = rescaled_phi;
  k.keep_parallel_frac = keep_parallel_frac;
  k.remove_parallel = remove_parallel;
  k.coef = nullptr;  // x == nullptr: guidance only, no sampler coefficients are read
  const int rc = launch_sampler_step(k, (cudaStream_t)stream);
  if (rc != CCDM_OK) return rc;
  return after_launch("sampler_step_kernel(cfg)");
}

extern "C" int ccdm_q_sample(const ccdm_qsample_args* a, void* stream) {
  CCDM_REQUIRE(a && a->img01 && a->noise && a->t && a->sqrt_acp && a->sqrt_1m_acp && a->x0 && a->noise_out && a->x_t,
               CCDM_ERR_BAD_ARG, "q_sample: null pointer");
  CCDM_REQUIRE(!a->cov || a->keep, CCDM_ERR_BAD_ARG, "q_sample: cov needs the keep mask");
  const long long total = (long long)a->B * a->chw;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  q_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
  return after_launch("q_sample_kernel");
}

extern "C" int ccdm_vicinal_loss(const ccdm_loss_args* a, void* stream) {
  CCDM_REQUIRE(a && a->model_out && a->x0 && a->noise && a->t && a->sqrt_acp && a->sqrt_1m_acp && a->loss_weight &&
                   a->per_sample && a->loss,
               CCDM_ERR_BAD_ARG, "vicinal_loss: null pointer");
  CCDM_REQUIRE(!a->cov || a->keep, CCDM_ERR_BAD_ARG, "vicinal_loss: cov needs the keep mask");
  CCDM_REQUIRE(a->objective >= 0 && a->objective <= 2, CCDM_ERR_BAD_ARG, "vicinal_loss: objective");
  cudaStream_t s = (cudaStream_t)stream;
  loss_rows_kernel<<<a->B, 512, 0, s>>>(*a);
  int rc = after_launch("loss_rows_kernel");
  if (rc != CCDM_OK) return rc;
  loss_final_kernel<<<1, 256, 0, s>>>(*a);
  rc = after_launch("loss_final_kernel");
  if (rc != CCDM_OK) return rc;
  if (a->grad_out) {
    const long long total = (long long)a->B * a->chw;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    loss_grad_kernel<<<(unsigned)blocks, 256, 0, s>>>(*a);
    rc = after_launch("loss_grad_kernel");
  }
  return rc;
}

extern "C" int ccdm_vicinal_weights(const float* proj, int32_t B, int32_t P, int32_t euclid, int32_t hard,
                                    const float* thr, float nu, const uint8_t* keep, float* w, void* stream) {
  CCDM_REQUIRE(proj && w && B > 0 && P > 0 && (!hard || thr), CCDM_ERR_BAD_ARG, "vicinal_weights: bad args");
  vicinal_weights_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(proj, B, P, euclid, hard, thr, nu, keep, w);
  return after_launch("vicinal_weights_kernel");
}
