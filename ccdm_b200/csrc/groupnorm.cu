// Vanilla (ADM-style) UNet pieces that the unified UNet does not have:
//   GroupNorm statistics + the per-(sample, channel) affine map that GroupNorm [+ (1+scale), shift] reduces to
//   (CCDM_vanilla/.../models/unet.py:88-90 norm_layer; call sites :99 conv1, :111 + :146 conv2, :160 attention, :323 out),
//   and the softmax attention over all H*W tokens with the reference's per-head [q|k|v] channel split (:165-175).
// GroupNorm needs a reduction over every pixel of a sample, so it cannot live in one conv tile's epilogue: the statistics
// are one HBM-bound read of the tensor, the apply step (ccdm_affine_act, SiLU) one read + one write in front of the conv.
#include <cfloat>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

// sums[b][0][c_off + c] += sum over the sample's pixels of x,  sums[b][1][c_off + c] += sum of x^2.
// x: bf16 [B][rows_per_sample][C] contiguous, C % 8 == 0, C <= 2048; sums: fp32 [B][2][ld].  grid = (slabs, B).
__global__ void __launch_bounds__(256) channel_stats_kernel(const uint4* __restrict__ x, int rows_per_sample, int C,
                                                            int slab, float* __restrict__ sums, int ld, int c_off) {
  __shared__ float acc_s[2 * 2048];
  const int nchunk = C >> 3;
  const int lanes = 256 / nchunk;                                  // row lanes per block
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < 2 * C; i += 256) acc_s[i] = 0.f;
  __syncthreads();
  const int ch = tid % nchunk, rl = tid / nchunk;
  if (rl < lanes) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint4* base = x + (long long)b * rows_per_sample * nchunk;
    const int r0 = blockIdx.x * slab, r1 = min(r0 + slab, rows_per_sample);
    for (int r = r0 + rl; r < r1; r += lanes) {
      const uint4 u = __ldg(base + (long long)r * nchunk + ch);
      const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                          bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] = fmaf(v[j], v[j], q[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&acc_s[ch * 8 + j], s[j]);
      atomicAdd(&acc_s[C + ch * 8 + j], q[j]);
    }
  }
  __syncthreads();
  float* dst = sums + (long long)b * 2 * ld + c_off;
  for (int i = tid; i < C; i += 256) {
    atomicAdd(dst + i, acc_s[i]);
    atomicAdd(dst + ld + i, acc_s[C + i]);
  }
}

// One thread per (sample, channel) of the concatenated channel axis.  GroupNorm(x)*gamma + beta, optionally followed by
// *(1+scale) + shift, is the affine map x*a + t with
//   a = rstd_g * gamma_c * (1 + scale_bc),   t = (beta_c - mean_g * rstd_g * gamma_c) * (1 + scale_bc) + shift_bc.
// Written per SOURCE in the [a-1 | t] layout ccdm_affine_act / ccdm_tapgemm read: channels [0,C0) -> coef[b][0..2*C0),
// channels [C0,Ctot) -> coef[b][2*C0 .. 2*Ctot).  A group may straddle the two sources (it is defined on the concatenation).
__global__ void groupnorm_coef_kernel(const float* __restrict__ sums, int Ctot, int G, float inv_count, float eps,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ ss_in, int ss_ld, int ss_off, int C0,
                                      float* __restrict__ coef, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ctot) return;
  const int b = i / Ctot, c = i - b * Ctot;
  const int cg = Ctot / G, g0 = (c / cg) * cg;
  const float* s = sums + (long long)b * 2 * Ctot;
  float sum = 0.f, sq = 0.f;
  for (int k = 0; k < cg; ++k) {
    sum += s[g0 + k];
    sq += s[Ctot + g0 + k];
  }
  const float mean = sum * inv_count;
  const float var = fmaxf(sq * inv_count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float a = rstd * gamma[c];
  float t = beta[c] - mean * a;
  if (ss_in) {
    const float sc = 1.f + ss_in[(long long)b * ss_ld + ss_off + c];
    a *= sc;
    t = fmaf(t, sc, ss_in[(long long)b * ss_ld + ss_off + Ctot + c]);
  }
  float* row = coef + (long long)b * 2 * Ctot;
  if (c < C0) {
    row[c] = a - 1.f;
    row[C0 + c] = t;
  } else {
    const int c1 = c - C0, C1 = Ctot - C0;
    row[2 * C0 + c1] = a - 1.f;
    row[2 * C0 + C1 + c1] = t;
  }
}

// ---------------------------------------------------------------------------- backward of norm -> affine -> activation
// Forward (per source): u = x*(1+s_bc) + t_bc, y = act(u), with (s, t) the forward coefficients of groupnorm_coef_kernel.
// d act(u)/du recomputed from x: act 0 none, 1 ReLU, 2 SiLU.
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == 1) return u > 0.f ? 1.f : 0.f;
  if (act == 2) {
    const float sg = 1.f / (1.f + __expf(-u));
    return sg * (1.f + u * (1.f - sg));
  }
  return 1.f;
}

// bsums[b][0][c_off+c] += sum_p du,  bsums[b][1][c_off+c] += sum_p du*x,  du = dy * act'(x*(1+s)+t).
// Same decomposition as channel_stats_kernel; dy, x: bf16 [B][rows_per_sample][C].  grid = (slabs, B).
__global__ void __launch_bounds__(256) norm_bwd_stats_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                             int rows_per_sample, int C, int slab,
                                                             const float* __restrict__ coef, int coef_ld, int coef_off,
                                                             int act, float* __restrict__ bsums, int ld, int c_off) {
  __shared__ float acc_s[2 * 2048];
  const int nchunk = C >> 3;
  const int lanes = 256 / nchunk;
  const int tid = threadIdx.x, b = blockIdx.y;
  for (int i = tid; i < 2 * C; i += 256) acc_s[i] = 0.f;
  __syncthreads();
  const int ch = tid % nchunk, rl = tid / nchunk;
  if (rl < lanes) {
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sc[8], sh[8];
    const float* crow = coef + (long long)b * coef_ld + coef_off;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = 1.f + crow[ch * 8 + j];
      sh[j] = crow[C + ch * 8 + j];
    }
    const long long base = (long long)b * rows_per_sample * nchunk;
    const int r0 = blockIdx.x * slab, r1 = min(r0 + slab, rows_per_sample);
    for (int r = r0 + rl; r < r1; r += lanes) {
      const uint4 ux = __ldg(x + base + (long long)r * nchunk + ch);
      const uint4 ud = __ldg(dy + base + (long long)r * nchunk + ch);
      const float xv[8] = {bf16_lo(ux.x), bf16_hi(ux.x), bf16_lo(ux.y), bf16_hi(ux.y),
                           bf16_lo(ux.z), bf16_hi(ux.z), bf16_lo(ux.w), bf16_hi(ux.w)};
      const float dv[8] = {bf16_lo(ud.x), bf16_hi(ud.x), bf16_lo(ud.y), bf16_hi(ud.y),
                           bf16_lo(ud.z), bf16_hi(ud.z), bf16_lo(ud.w), bf16_hi(ud.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = dv[j] * act_grad(fmaf(xv[j], sc[j], sh[j]), act);
        s1[j] += du;
        s2[j] = fmaf(du, xv[j], s2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&acc_s[ch * 8 + j], s1[j]);
      atomicAdd(&acc_s[C + ch * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  float* dst = bsums + (long long)b * 2 * ld + c_off;
  for (int i = tid; i < C; i += 256) {
    atomicAdd(dst + i, acc_s[i]);
    atomicAdd(dst + ld + i, acc_s[C + i]);
  }
}

// One thread per (sample, channel) of the concatenated axis.  With x^ = (x-mu_g) r_g, m = 1+scale_bc, S1 = sum_p du,
// S2 = sum_p du*x, Sx = sum_p du*x^ = r (S2 - mu S1), and the group means M1 = sum_{c in g} gamma_c m S1 / cnt,
// M2 = sum_{c in g} gamma_c m Sx / cnt:
//   dx = A*du + Bc*x + Cc,   A = r gamma m (the forward coefficient),  Bc = -r^2 M2,  Cc = r^2 M2 mu - r M1
//   dgamma_c += m Sx,  dbeta_c += m S1,  dscale_bc = gamma Sx + beta S1,  dshift_bc = S1.
// bcoef: per source [A | Bc | Cc] segments (source 0 at 0, source 1 at 3*C0); d_ss: [B][2*Ctot] = [dscale | dshift].
__global__ void groupnorm_bwd_coef_kernel(const float* __restrict__ sums, const float* __restrict__ bsums, int Ctot, int G,
                                          float inv_count, float eps, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, const float* __restrict__ ss_in, int ss_ld,
                                          int ss_off, int C0, float* __restrict__ bcoef, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta, float* __restrict__ d_ss, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ctot) return;
  const int b = i / Ctot, c = i - b * Ctot;
  const int cg = Ctot / G, g0 = (c / cg) * cg;
  const float* s = sums + (long long)b * 2 * Ctot;
  const float* bs = bsums + (long long)b * 2 * Ctot;
  float sum = 0.f, sq = 0.f;
  for (int k = 0; k < cg; ++k) {
    sum += s[g0 + k];
    sq += s[Ctot + g0 + k];
  }
  const float mean = sum * inv_count;
  const float var = fmaxf(sq * inv_count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int k = 0; k < cg; ++k) {
    const int ck = g0 + k;
    const float mk = ss_in ? 1.f + ss_in[(long long)b * ss_ld + ss_off + ck] : 1.f;
    const float gm = gamma[ck] * mk;
    m1 = fmaf(gm, bs[ck], m1);
    m2 = fmaf(gm, rstd * (bs[Ctot + ck] - mean * bs[ck]), m2);
  }
  m1 *= inv_count;
  m2 *= inv_count;
  const float m = ss_in ? 1.f + ss_in[(long long)b * ss_ld + ss_off + c] : 1.f;
  const float S1 = bs[c], Sx = rstd * (bs[Ctot + c] - mean * S1);
  const float A = rstd * gamma[c] * m;
  const float Bc = -rstd * rstd * m2;
  const float Cc = rstd * rstd * m2 * mean - rstd * m1;
  float* row = bcoef + (long long)b * 3 * Ctot;
  if (c < C0) {
    row[c] = A;
    row[C0 + c] = Bc;
    row[2 * C0 + c] = Cc;
  } else {
    const int c1 = c - C0, C1 = Ctot - C0;
    row[3 * C0 + c1] = A;
    row[3 * C0 + C1 + c1] = Bc;
    row[3 * C0 + 2 * C1 + c1] = Cc;
  }
  if (dgamma) atomicAdd(dgamma + c, m * Sx);
  if (dbeta) atomicAdd(dbeta + c, m * S1);
  if (d_ss) {
    d_ss[(long long)b * 2 * Ctot + c] = fmaf(gamma[c], Sx, beta[c] * S1);
    d_ss[(long long)b * 2 * Ctot + Ctot + c] = S1;
  }
}

// dx = A*du + Bc*x + Cc per element, du = dy * act'(x*(1+s)+t); one thread per 8-channel vector.
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                             uint4* __restrict__ dx, long long nvec, int nchunk,
                                                             int rows_per_sample, int C, const float* __restrict__ coef,
                                                             int coef_ld, int coef_off, const float* __restrict__ bcoef,
                                                             int bcoef_ld, int bcoef_off, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nchunk;
    const int ch = (int)(i - row * nchunk);
    const long long b = row / rows_per_sample;
    const float* crow = coef + b * coef_ld + coef_off + ch * 8;
    const float* brow = bcoef + b * bcoef_ld + bcoef_off + ch * 8;
    const uint4 ux = __ldg(x + i), ud = __ldg(dy + i);
    const float xv[8] = {bf16_lo(ux.x), bf16_hi(ux.x), bf16_lo(ux.y), bf16_hi(ux.y),
                         bf16_lo(ux.z), bf16_hi(ux.z), bf16_lo(ux.w), bf16_hi(ux.w)};
    const float dv[8] = {bf16_lo(ud.x), bf16_hi(ud.x), bf16_lo(ud.y), bf16_hi(ud.y),
                         bf16_lo(ud.z), bf16_hi(ud.z), bf16_lo(ud.w), bf16_hi(ud.w)};
    alignas(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float du = dv[j] * act_grad(fmaf(xv[j], 1.f + crow[j], crow[C + j]), act);
      o[j] = __float2bfloat16(fmaf(brow[j], du, fmaf(brow[C + j], xv[j], brow[2 * C + j])));
    }
    dx[i] = *reinterpret_cast<const uint4*>(o);
  }
}

// timestep_embedding (V/models/unet.py:40-57): out[b] = [cos(t*f_j) | sin(t*f_j)], f_j = exp(-ln(max_period)*j/half).
// (The unified UNet's SinusoidalPosEmb is sin | cos with the exponent divided by half-1: ccdm_time_features.)
__global__ void time_features_adm_kernel(const long long* __restrict__ t, int B, int dim, float log_period,
                                         float* __restrict__ out) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float f = expf(-log_period * (float)j / (float)half);
  const float a = (float)t[b] * f;
  out[(long long)b * dim + j] = cosf(a);
  out[(long long)b * dim + half + j] = sinf(a);
}

// softmax(q k^T * scale) v over n tokens for one (sample, head); grid = (B*heads, query blocks of 64), 128 threads:
// two threads per query, each owning the even / odd half of the head's DH dimensions (dot products are completed with one
// shuffle); keys / values are staged through shared memory 32 tokens at a time with a running (online) softmax.
// Channel layout of qkv [B][n][3*heads*DH]: head h starts at h*hs; q, k, v at +qo, +ko, +vo.
template <int DH>
__global__ void __launch_bounds__(128) attention_tokens_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                               __nv_bfloat16* __restrict__ out, int n, int heads,
                                                               float scale, int hs, int qo, int ko, int vo) {
  constexpr int HD = DH / 2, KT = 32;
  __shared__ float sk[KT][DH], sv[KT][DH];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int hid = heads * DH, ld = 3 * hid;
  const __nv_bfloat16* base = qkv + (long long)b * n * ld + h * hs;
  const int tid = threadIdx.x, half = tid & 1;
  const int i = blockIdx.y * 64 + (tid >> 1);
  const bool active = i < n;
  float q[HD], acc[HD];
#pragma unroll
  for (int dd = 0; dd < HD; ++dd) {
    q[dd] = active ? __bfloat162float(base[(long long)i * ld + qo + 2 * dd + half]) * scale : 0.f;
    acc[dd] = 0.f;
  }
  float m_run = -FLT_MAX, l_run = 0.f;
  for (int j0 = 0; j0 < n; j0 += KT) {
    const int rows = min(KT, n - j0);
    __syncthreads();
    for (int e = tid; e < rows * DH; e += 128) {
      const int tok = e / DH, dd = e % DH;
      const __nv_bfloat16* p = base + (long long)(j0 + tok) * ld + dd;
      sk[tok][dd] = __bfloat162float(p[ko]);
      sv[tok][dd] = __bfloat162float(p[vo]);
    }
    __syncthreads();
    for (int j = 0; j < rows; ++j) {
      float s = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) s = fmaf(q[dd], sk[j][2 * dd + half], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      const float m_new = fmaxf(m_run, s);
      const float corr = __expf(m_run - m_new), p = __expf(s - m_new);
      l_run = l_run * corr + p;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) acc[dd] = fmaf(acc[dd], corr, p * sv[j][2 * dd + half]);
      m_run = m_new;
    }
  }
  if (!active) return;
  const float inv = 1.f / l_run;
  __nv_bfloat16* o = out + ((long long)b * n + i) * hid + h * DH;
#pragma unroll
  for (int dd = 0; dd < HD; ++dd) o[2 * dd + half] = __float2bfloat16(acc[dd] * inv);
}

// Backward of attention_tokens_kernel.  One CTA per (sample, head); shared memory holds q, k, v, dO of the head as fp32
// [n][DH] plus per-query (max, 1/sum, D = <dO, O>).  SPLIT threads share one token, thread `part` owning the head
// dimensions d = SPLIT*dd + part (dot products are completed with xor shuffles).  Phase A (token = query): softmax
// statistics and dq.  Phase B (token = key): dk, dv.  Same channel layout parameters as the forward.
template <int DH, int SPLIT>
__global__ void __launch_bounds__(128) attention_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   const __nv_bfloat16* __restrict__ dout,
                                                                   __nv_bfloat16* __restrict__ dqkv, int n, int heads,
                                                                   float scale, int hs, int qo, int ko, int vo) {
  constexpr int HD = DH / SPLIT, TOK = 128 / SPLIT;
  extern __shared__ float att_smem[];
  float* sq = att_smem;
  float* sk = sq + (size_t)n * DH;
  float* sv = sk + (size_t)n * DH;
  float* sd = sv + (size_t)n * DH;
  float* st = sd + (size_t)n * DH;                                 // [n][3]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int hid = heads * DH, ld = 3 * hid;
  const __nv_bfloat16* base = qkv + (size_t)b * n * ld + h * hs;
  const __nv_bfloat16* dob = dout + (size_t)b * n * hid + h * DH;
  __nv_bfloat16* gbase = dqkv + (size_t)b * n * ld + h * hs;
  const int tid = threadIdx.x, part = tid % SPLIT, slot = tid / SPLIT;
  for (int e = tid; e < n * DH; e += 128) {
    const int tok = e / DH, dd = e % DH;
    const __nv_bfloat16* p = base + (size_t)tok * ld + dd;
    sq[e] = __bfloat162float(p[qo]) * scale;                       // q arrives pre-scaled in shared memory
    sk[e] = __bfloat162float(p[ko]);
    sv[e] = __bfloat162float(p[vo]);
    sd[e] = __bfloat162float(dob[(size_t)tok * hid + dd]);
  }
  __syncthreads();
  // ---- phase A: per query i
  for (int i0 = 0; i0 < n; i0 += TOK) {
    const int i = i0 + slot;
    const bool active = i < n;
    const int ir = active ? i : 0;
    float q[HD], g[HD], dq[HD];
#pragma unroll
    for (int dd = 0; dd < HD; ++dd) {
      q[dd] = sq[ir * DH + SPLIT * dd + part];
      g[dd] = sd[ir * DH + SPLIT * dd + part];
      dq[dd] = 0.f;
    }
    float m = -FLT_MAX;
    for (int j = 0; j < n; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) sc = fmaf(q[dd], sk[j * DH + SPLIT * dd + part], sc);
#pragma unroll
      for (int off = 1; off < SPLIT; off <<= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);
      m = fmaxf(m, sc);
    }
    float l = 0.f, dsum = 0.f;                                      // dsum = sum_j e_ij * <dO_i, v_j>
    for (int j = 0; j < n; ++j) {
      float sc = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        sc = fmaf(q[dd], sk[j * DH + SPLIT * dd + part], sc);
        gv = fmaf(g[dd], sv[j * DH + SPLIT * dd + part], gv);
      }
#pragma unroll
      for (int off = 1; off < SPLIT; off <<= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, off);
        gv += __shfl_xor_sync(0xffffffffu, gv, off);
      }
      const float e = __expf(sc - m);
      l += e;
      dsum = fmaf(e, gv, dsum);
    }
    const float inv = 1.f / l;
    const float D = dsum * inv;                                     // <dO_i, O_i>
    for (int j = 0; j < n; ++j) {
      float sc = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        sc = fmaf(q[dd], sk[j * DH + SPLIT * dd + part], sc);
        gv = fmaf(g[dd], sv[j * DH + SPLIT * dd + part], gv);
      }
#pragma unroll
      for (int off = 1; off < SPLIT; off <<= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, off);
        gv += __shfl_xor_sync(0xffffffffu, gv, off);
      }
      const float ds = __expf(sc - m) * inv * (gv - D);
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) dq[dd] = fmaf(ds, sk[j * DH + SPLIT * dd + part], dq[dd]);
    }
    if (active) {
      if (part == 0) {
        st[i * 3] = m;
        st[i * 3 + 1] = inv;
        st[i * 3 + 2] = D;
      }
      __nv_bfloat16* o = gbase + (size_t)i * ld + qo;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) o[SPLIT * dd + part] = __float2bfloat16(dq[dd] * scale);
    }
  }
  __syncthreads();
  // ---- phase B: per key j
  for (int j0 = 0; j0 < n; j0 += TOK) {
    const int j = j0 + slot;
    const bool active = j < n;
    const int jr = active ? j : 0;
    float kk[HD], vv[HD], dk[HD], dv[HD];
#pragma unroll
    for (int dd = 0; dd < HD; ++dd) {
      kk[dd] = sk[jr * DH + SPLIT * dd + part];
      vv[dd] = sv[jr * DH + SPLIT * dd + part];
      dk[dd] = 0.f;
      dv[dd] = 0.f;
    }
    for (int i = 0; i < n; ++i) {
      float sc = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        sc = fmaf(sq[i * DH + SPLIT * dd + part], kk[dd], sc);
        gv = fmaf(sd[i * DH + SPLIT * dd + part], vv[dd], gv);
      }
#pragma unroll
      for (int off = 1; off < SPLIT; off <<= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, off);
        gv += __shfl_xor_sync(0xffffffffu, gv, off);
      }
      const float pr = __expf(sc - st[i * 3]) * st[i * 3 + 1];
      const float ds = pr * (gv - st[i * 3 + 2]);
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        dk[dd] = fmaf(ds, sq[i * DH + SPLIT * dd + part], dk[dd]);   // q in shared memory already carries `scale`
        dv[dd] = fmaf(pr, sd[i * DH + SPLIT * dd + part], dv[dd]);
      }
    }
    if (active) {
      __nv_bfloat16* o = gbase + (size_t)j * ld;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        o[ko + SPLIT * dd + part] = __float2bfloat16(dk[dd]);
        o[vo + SPLIT * dd + part] = __float2bfloat16(dv[dd]);
      }
    }
  }
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_channel_stats(const void* x, int32_t B, int32_t rows_per_sample, int32_t C, float* sums, int32_t ld,
                                  int32_t c_off, int32_t zero_first, void* stream) {
  CCDM_REQUIRE(x && sums && B > 0 && rows_per_sample > 0 && c_off >= 0 && ld >= c_off + C, CCDM_ERR_BAD_ARG,
               "channel_stats: bad args");
  CCDM_REQUIRE(C > 0 && C % 8 == 0 && C <= 2048, CCDM_ERR_UNSUPPORTED_SHAPE, "channel_stats: C=%d (multiple of 8, <= 2048)", C);
  cudaStream_t s = (cudaStream_t)stream;
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * 2 * ld * sizeof(float), s);
    if (e != cudaSuccess) return cuda_fail(e, "channel_stats: memset");
  }
  const int lanes = 256 / (C / 8);
  int slabs = (num_sms() * 4 + B - 1) / B;                         // a few CTAs per SM over the whole batch
  const int max_slabs = (rows_per_sample + lanes * 4 - 1) / (lanes * 4);   // >= 4 rows per lane
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const int slab = (rows_per_sample + slabs - 1) / slabs;
  slabs = (rows_per_sample + slab - 1) / slab;
  channel_stats_kernel<<<dim3(slabs, B), 256, 0, s>>>((const uint4*)x, rows_per_sample, C, slab, sums, ld, c_off);
  return after_launch("channel_stats_kernel");
}

extern "C" int ccdm_groupnorm_coef(const float* sums, int32_t B, int32_t Ctot, int32_t groups, int64_t rows_per_sample,
                                   float eps, const float* gamma, const float* beta, const float* scale_shift,
                                   int32_t ss_ld, int32_t ss_off, int32_t C0, float* coef, void* stream) {
  CCDM_REQUIRE(sums && gamma && beta && coef && B > 0 && Ctot > 0 && rows_per_sample > 0, CCDM_ERR_BAD_ARG,
               "groupnorm_coef: bad args");
  CCDM_REQUIRE(groups > 0 && Ctot % groups == 0 && C0 > 0 && C0 <= Ctot, CCDM_ERR_BAD_ARG,
               "groupnorm_coef: Ctot=%d groups=%d C0=%d", Ctot, groups, C0);
  const int n = B * Ctot;
  const float inv_count = 1.f / ((float)rows_per_sample * (float)(Ctot / groups));
  groupnorm_coef_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sums, Ctot, groups, inv_count, eps, gamma, beta,
                                                                           scale_shift, ss_ld, ss_off, C0, coef, B);
  return after_launch("groupnorm_coef_kernel");
}

extern "C" int ccdm_norm_bwd_stats(const void* dy, const void* x, int32_t B, int32_t rows_per_sample, int32_t C,
                                   const float* coef, int32_t coef_ld, int32_t coef_off, int32_t act, float* bsums,
                                   int32_t ld, int32_t c_off, int32_t zero_first, void* stream) {
  CCDM_REQUIRE(dy && x && coef && bsums && B > 0 && rows_per_sample > 0 && c_off >= 0 && ld >= c_off + C &&
                   coef_ld >= coef_off + 2 * C && act >= 0 && act <= 2,
               CCDM_ERR_BAD_ARG, "norm_bwd_stats: bad args");
  CCDM_REQUIRE(C > 0 && C % 8 == 0 && C <= 2048, CCDM_ERR_UNSUPPORTED_SHAPE, "norm_bwd_stats: C=%d (multiple of 8, <= 2048)", C);
  cudaStream_t s = (cudaStream_t)stream;
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(bsums, 0, (size_t)B * 2 * ld * sizeof(float), s);
    if (e != cudaSuccess) return cuda_fail(e, "norm_bwd_stats: memset");
  }
  const int lanes = 256 / (C / 8);
  int slabs = (num_sms() * 4 + B - 1) / B;
  const int max_slabs = (rows_per_sample + lanes * 4 - 1) / (lanes * 4);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const int slab = (rows_per_sample + slabs - 1) / slabs;
  slabs = (rows_per_sample + slab - 1) / slab;
  norm_bwd_stats_kernel<<<dim3(slabs, B), 256, 0, s>>>((const uint4*)dy, (const uint4*)x, rows_per_sample, C, slab, coef,
                                                       coef_ld, coef_off, act, bsums, ld, c_off);
  return after_launch("norm_bwd_stats_kernel");
}

extern "C" int ccdm_groupnorm_bwd_coef(const float* sums, const float* bsums, int32_t B, int32_t Ctot, int32_t groups,
                                       int64_t rows_per_sample, float eps, const float* gamma, const float* beta,
                                       const float* scale_shift, int32_t ss_ld, int32_t ss_off, int32_t C0, float* bcoef,
                                       float* dgamma, float* dbeta, float* d_ss, void* stream) {
  CCDM_REQUIRE(sums && bsums && gamma && beta && bcoef && B > 0 && Ctot > 0 && rows_per_sample > 0, CCDM_ERR_BAD_ARG,
               "groupnorm_bwd_coef: bad args");
  CCDM_REQUIRE(groups > 0 && Ctot % groups == 0 && C0 > 0 && C0 <= Ctot, CCDM_ERR_BAD_ARG,
               "groupnorm_bwd_coef: Ctot=%d groups=%d C0=%d", Ctot, groups, C0);
  const int n = B * Ctot;
  const float inv_count = 1.f / ((float)rows_per_sample * (float)(Ctot / groups));
  groupnorm_bwd_coef_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sums, bsums, Ctot, groups, inv_count, eps,
                                                                               gamma, beta, scale_shift, ss_ld, ss_off, C0,
                                                                               bcoef, dgamma, dbeta, d_ss, B);
  return after_launch("groupnorm_bwd_coef_kernel");
}

extern "C" int ccdm_norm_bwd_apply(const void* dy, const void* x, void* dx, int64_t rows, int32_t C, int32_t rows_per_sample,
                                   const float* coef, int32_t coef_ld, int32_t coef_off, const float* bcoef,
                                   int32_t bcoef_ld, int32_t bcoef_off, int32_t act, void* stream) {
  CCDM_REQUIRE(dy && x && dx && coef && bcoef && rows > 0 && rows_per_sample > 0 && rows % rows_per_sample == 0 &&
                   act >= 0 && act <= 2,
               CCDM_ERR_BAD_ARG, "norm_bwd_apply: bad args");
  CCDM_REQUIRE(C > 0 && C % 8 == 0, CCDM_ERR_UNSUPPORTED_SHAPE, "norm_bwd_apply: C=%d", C);
  const long long nvec = rows * (C / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  norm_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)x, (uint4*)dx,
                                                                            nvec, C / 8, rows_per_sample, C, coef, coef_ld,
                                                                            coef_off, bcoef, bcoef_ld, bcoef_off, act);
  return after_launch("norm_bwd_apply_kernel");
}

extern "C" int ccdm_time_features_adm(const int64_t* t, int32_t B, int32_t dim, float max_period, float* out,
                                      void* stream) {
  CCDM_REQUIRE(t && out && B > 0 && dim >= 2 && dim % 2 == 0 && max_period > 1.f, CCDM_ERR_BAD_ARG,
               "time_features_adm: bad args");
  const int n = B * (dim / 2);
  time_features_adm_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)t, B, dim,
                                                                              logf(max_period), out);
  return after_launch("time_features_adm_kernel");
}

extern "C" int ccdm_attention_tokens(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t dim_head,
                                     float scale, int32_t head_major, void* stream) {
  CCDM_REQUIRE(qkv && out && B > 0 && n >= 1 && heads >= 1, CCDM_ERR_BAD_ARG, "attention_tokens: bad args");
  const int hid = heads * dim_head;
  const int hs = head_major ? 3 * dim_head : dim_head;
  const int qo = 0, ko = head_major ? dim_head : hid, vo = head_major ? 2 * dim_head : 2 * hid;
  dim3 grid(B * heads, (n + 63) / 64);
  const __nv_bfloat16* in = (const __nv_bfloat16*)qkv;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim_head) {
    case 16: attention_tokens_kernel<16><<<grid, 128, 0, s>>>(in, o, n, heads, scale, hs, qo, ko, vo); break;
    case 32: attention_tokens_kernel<32><<<grid, 128, 0, s>>>(in, o, n, heads, scale, hs, qo, ko, vo); break;
    case 64: attention_tokens_kernel<64><<<grid, 128, 0, s>>>(in, o, n, heads, scale, hs, qo, ko, vo); break;
    case 128: attention_tokens_kernel<128><<<grid, 128, 0, s>>>(in, o, n, heads, scale, hs, qo, ko, vo); break;
    default:
      CCDM_REQUIRE(false, CCDM_ERR_UNSUPPORTED_SHAPE, "attention_tokens: dim_head=%d (supported: 16, 32, 64, 128)", dim_head);
  }
  return after_launch("attention_tokens_kernel");
}

extern "C" int ccdm_attention_tokens_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t n, int32_t heads,
                                         int32_t dim_head, float scale, int32_t head_major, void* stream) {
  CCDM_REQUIRE(qkv && dout && dqkv && B > 0 && n >= 1 && heads >= 1, CCDM_ERR_BAD_ARG, "attention_tokens_bwd: bad args");
  const size_t smem = ((size_t)4 * n * dim_head + (size_t)3 * n) * sizeof(float);
  CCDM_REQUIRE(smem <= 200 * 1024, CCDM_ERR_UNSUPPORTED_SHAPE,
               "attention_tokens_bwd: %d tokens x %d do not fit shared memory", n, dim_head);
  const int hid = heads * dim_head;
  const int hs = head_major ? 3 * dim_head : dim_head;
  const int qo = 0, ko = head_major ? dim_head : hid, vo = head_major ? 2 * dim_head : 2 * hid;
  cudaStream_t s = (cudaStream_t)stream;
  const __nv_bfloat16* in = (const __nv_bfloat16*)qkv;
  const __nv_bfloat16* dob = (const __nv_bfloat16*)dout;
  __nv_bfloat16* o = (__nv_bfloat16*)dqkv;
  cudaError_t e = cudaSuccess;
  switch (dim_head) {
    case 16:
      e = cudaFuncSetAttribute(attention_tokens_bwd_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "attention_tokens_bwd: cudaFuncSetAttribute");
      attention_tokens_bwd_kernel<16, 2><<<B * heads, 128, smem, s>>>(in, dob, o, n, heads, scale, hs, qo, ko, vo);
      break;
    case 32:
      e = cudaFuncSetAttribute(attention_tokens_bwd_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "attention_tokens_bwd: cudaFuncSetAttribute");
      attention_tokens_bwd_kernel<32, 2><<<B * heads, 128, smem, s>>>(in, dob, o, n, heads, scale, hs, qo, ko, vo);
      break;
    case 64:
      e = cudaFuncSetAttribute(attention_tokens_bwd_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "attention_tokens_bwd: cudaFuncSetAttribute");
      attention_tokens_bwd_kernel<64, 2><<<B * heads, 128, smem, s>>>(in, dob, o, n, heads, scale, hs, qo, ko, vo);
      break;
    case 128:
      e = cudaFuncSetAttribute(attention_tokens_bwd_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "attention_tokens_bwd: cudaFuncSetAttribute");
      attention_tokens_bwd_kernel<128, 4><<<B * heads, 128, smem, s>>>(in, dob, o, n, heads, scale, hs, qo, ko, vo);
      break;
    default:
      CCDM_REQUIRE(false, CCDM_ERR_UNSUPPORTED_SHAPE, "attention_tokens_bwd: dim_head=%d (supported: 16, 32, 64, 128)", dim_head);
  }
  return after_launch("attention_tokens_bwd_kernel");
}
