// CUDA-core kernels around the tap-GEMM: stem / head boundary convs, attention cores, embedding MLPs.
// All are bandwidth- or latency-bound; see DESIGN.md for the per-kernel roofline.
#include <cfloat>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

// ============================================================================ stem on tensor cores: im2row + weight packing
// The 7x7 stem becomes a 4-tap tap-GEMM over a bf16 NHWC "im2row" tensor whose 64 channels hold a 2-row x 7-column
// window of the input:  rowimg[b,j,w, dr*7*Cin + s*Cin + c] = x[b,c,j-1+dr,w+s-3]  for j = 0..H (H+1 rows; dr in {0,1};
// zero outside the image and for the unused channels).  Vertical tap pair g reads rowimg at row offset 2g-2; the
// extra leading row keeps the pair (x[-1], x[0]) that straddles the top edge.
// One CTA per output row (b, j): the two input rows it needs (j-1 and j, Cin planes, 3 zero pixels on either side) are
// staged in shared memory with coalesced loads, then every thread builds 16-byte vectors (8 of the 64 channels) so that a
// warp writes 512 contiguous bytes.  (The first version gave each thread a whole pixel and 42 scattered global loads: the
// load/store unit, not HBM, bounded it at 1.85 TB/s.)  kCin > 0 makes the channel decomposition compile-time arithmetic.
template <int kCin>
__global__ void __launch_bounds__(256) stem_im2row_kernel(const float* __restrict__ x, uint4* __restrict__ out, int Cin_rt,
                                                          int H, int W) {
  extern __shared__ float stem_rows[];                             // [2][Cin][W + 6]
  const int Cin = kCin > 0 ? kCin : Cin_rt;
  const int WP = W + 6;
  const long long row = blockIdx.x;                                // b * (H + 1) + j
  const int h = (int)(row % (H + 1)) - 1;                          // h = j - 1
  const long long b = row / (H + 1);
  for (int i = threadIdx.x; i < 2 * Cin * WP; i += blockDim.x) {
    const int dr = i / (Cin * WP), rem = i - dr * Cin * WP, c = rem / WP, xx = rem - c * WP - 3;
    const int yy = h + dr;
    stem_rows[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((b * Cin + c) * H + yy) * W + xx) : 0.f;
  }
  __syncthreads();
  const int nval = 14 * Cin;
  for (int idx = threadIdx.x; idx < W * 8; idx += blockDim.x) {
    const int g = idx & 7, w = idx >> 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      float val = 0.f;
      if (ch < nval) {
        const int dr = ch / (7 * Cin), rem = ch % (7 * Cin), sx = rem / Cin, c = rem % Cin;
        val = stem_rows[(dr * Cin + c) * WP + w + sx];
      }
      v[j] = val;
    }
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    out[row * W * 8 + idx] = u;
  }
}

// packed[n][g*64 + dr*7*Cin + s*Cin + c] = w[n][c][2g+dr][s]  (0 for row 7 and for the unused channels)
__global__ void stem_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin,
                                 int n_rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * 256) return;
  const int n = idx / 256, k = idx % 256, g = k / 64, ch = k % 64;
  float val = 0.f;
  if (n < Cout && ch < 14 * Cin) {
    const int dr = ch / (7 * Cin), rem = ch % (7 * Cin), sx = rem / Cin, c = rem % Cin;
    const int r = 2 * g + dr;
    if (r < 7) val = w[((n * Cin + c) * 7 + r) * 7 + sx];
  }
  out[idx] = __float2bfloat16(val);
}

// ============================================================================ head: 1x1 conv, NHWC bf16 -> NCHW fp32
// unet.py:348,455.  One pixel per thread; reads the pixel row once, writes Cout planes coalesced.
__global__ void __launch_bounds__(256) head_conv1_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, long long npix_total, int HW, int Cin,
                                                         int Cout) {
  extern __shared__ float s_head[];  // [Cout][Cin] + [Cout]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) s_head[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_head[Cout * Cin + i] = bias[i];
  __syncthreads();
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npix_total) return;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = o < Cout ? s_head[Cout * Cin + o] : 0.f;
  const uint4* row = reinterpret_cast<const uint4*>(x + p * Cin);
  for (int c8 = 0; c8 < Cin / 8; ++c8) {
    const uint4 u = __ldg(row + c8);
    const float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                        bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < Cout) {
        const float* wr = s_head + o * Cin + c8 * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o] = fmaf(v[j], wr[j], acc[o]);
      }
    }
  }
  const long long b = p / HW, hw = p % HW;
  for (int o = 0; o < Cout; ++o) out[(b * Cout + o) * HW + hw] = acc[o];
}

// ============================================================================ linear attention: fold context into W_out
// wfold[b][c][h*32+d] = sum_e w_out[c][h*32+e] * ctx[b][h][d][e].  One CTA per (sample, 32-row slab of c).
__global__ void linattn_fold_kernel(const float* __restrict__ w_out, const float* __restrict__ ctx,
                                    __nv_bfloat16* __restrict__ wfold, int C, int n_rows, int heads) {
  extern __shared__ float s_fold[];          // ws[32][hid]
  const int hid = heads * 32;
  const int t = threadIdx.x;                 // column (h, d), blockDim.x == hid
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * 32;
  float cr[32];                              // ctx[b][h][d][0..32)
  const float* crow = ctx + ((long long)b * hid + t) * 32;
#pragma unroll
  for (int e4 = 0; e4 < 8; ++e4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(crow) + e4);
    cr[e4 * 4] = v.x; cr[e4 * 4 + 1] = v.y; cr[e4 * 4 + 2] = v.z; cr[e4 * 4 + 3] = v.w;
  }
  for (int i = t; i < 32 * hid; i += hid) {
    const int c = c0 + i / hid;
    s_fold[i] = c < C ? w_out[(long long)c * hid + i % hid] : 0.f;
  }
  __syncthreads();
  const int hbase = (t >> 5) * 32;
  for (int cc = 0; cc < 32; ++cc) {
    const int c = c0 + cc;
    if (c >= n_rows) break;
    float acc = 0.f;
    const float* wr = s_fold + cc * hid + hbase;
#pragma unroll
    for (int e = 0; e < 32; ++e) acc = fmaf(wr[e], cr[e], acc);
    wfold[((long long)b * n_rows + c) * hid + t] = __float2bfloat16(acc);
  }
}

// ============================================================================ standalone channel RMSNorm (+ tail)
// unet.py:88-89,145-151 for layers whose channel count does not fit one tap-GEMM tile (Cout > 512: the 576-wide
// bottleneck of the dim-72 UTKFace-64 model) or whose tiny pixel count wants the GEMM split over output channels
// (4x4 levels): z = conv + bias (bf16) in, out = [silu]( z/|z| * g*sqrt(C) * (1+scale) + shift ) [+ resid].
// One warp per pixel row, 16-byte loads; HBM-bound (reads z [+ resid], writes out).
// Lane mapping: a row of C channels is nchunk = C/8 16-byte vectors; G = ceil(nchunk / kV) lanes share a row, each
// owning kV vectors, so a warp covers 32/G rows per iteration (C = 64 -> 4 rows, C = 72 -> 3 rows of 9 lanes).  All
// rows of a CTA belong to one sample (blockIdx.y), so gain * (1 + scale) and shift live in registers.
constexpr int rms_stages(int kv) { return kv == 1 ? 6 : kv == 2 ? 4 : 3; }

template <int kV>
__global__ void __launch_bounds__(256) rmsnorm_act_kernel(const uint4* __restrict__ z, uint4* __restrict__ out,
                                                          int rows_per_sample, int rows_per_item, int items_per_sample,
                                                          int n_items, int C, int G, const float* __restrict__ gain,
                                                          float gain_mul, const float* __restrict__ ss, int ss_ld,
                                                          int ss_off, const uint4* __restrict__ resid,
                                                          float* __restrict__ out_rowss, uint32_t flags) {
  constexpr int kStages = rms_stages(kV);
  extern __shared__ float rms_ring_f[];                            // [stage][vector][z | resid][thread] 16-byte slots
  uint4* const ring = reinterpret_cast<uint4*>(rms_ring_f);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = C >> 3;
  const int rpw = 32 / G;                                          // rows per warp iteration
  const int sub = lane / G, gl = lane - sub * G;
  const int rstep = 8 * rpw;
  // persistent CTAs: work item = (sample, slab of rows); the grid is one resident wave
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / items_per_sample;
    const int r0 = (item - b * items_per_sample) * rows_per_item;
    const int r1 = min(r0 + rows_per_item, rows_per_sample);
    float2 a[kV][4], sh[kV][4];
#pragma unroll
    for (int k = 0; k < kV; ++k) {
      const int ch = gl + G * k;
      float g[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (ch < nchunk) {
        load8(gain + ch * 8, g);
        if (flags & CCDM_EPI_SS) {
          load8(ss + (size_t)b * ss_ld + ss_off + ch * 8, sc);
          load8(ss + (size_t)b * ss_ld + ss_off + C + ch * 8, sf);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[k][j] = make_float2(g[2 * j] * gain_mul * (1.f + sc[2 * j]), g[2 * j + 1] * gain_mul * (1.f + sc[2 * j + 1]));
        sh[k][j] = make_float2(sf[2 * j], sf[2 * j + 1]);
      }
    }
    // per-thread LDGSTS staging ring (see block_bwd_kernel): kStages - 1 row groups in flight without holding registers
    auto slot = [&](int st, int k, int which) { return ring + ((st * kV + k) * 2 + which) * 256 + tid; };
    auto fetch = [&](int rb, int st) {
      const int r = rb + sub;
      const bool ok = sub < rpw && r < r1;
      const size_t off = ((size_t)b * rows_per_sample + (ok ? r : r0)) * nchunk;
#pragma unroll
      for (int k = 0; k < kV; ++k) {
        const int ch = gl + G * k;
        const bool v = ok && ch < nchunk;
        cp_async16(slot(st, k, 0), z + off + (v ? ch : 0), v);
        if (flags & CCDM_EPI_RESID) cp_async16(slot(st, k, 1), resid + off + (v ? ch : 0), v);
      }
      cp_async_commit();
    };
    const int rb0 = r0 + warp * rpw;
#pragma unroll
    for (int st = 0; st < kStages - 1; ++st) fetch(rb0 + st * rstep, st);
    int st_rd = 0, st_wr = kStages - 1;
    for (int rb = rb0; rb < r1; rb += rstep) {
      const int r = rb + sub;
      const bool live = sub < rpw && r < r1;
      const size_t rowoff = ((size_t)b * rows_per_sample + r) * nchunk;
      fetch(rb + (kStages - 1) * rstep, st_wr);
      cp_async_wait<kStages - 1>();
      float2 v[kV][4];
      uint4 ru[kV];
      float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < kV; ++k) {
        const uint4 u = *slot(st_rd, k, 0);
        ru[k] = (flags & CCDM_EPI_RESID) ? *slot(st_rd, k, 1) : make_uint4(0, 0, 0, 0);
        v[k][0] = make_float2(bf16_lo(u.x), bf16_hi(u.x));
        v[k][1] = make_float2(bf16_lo(u.y), bf16_hi(u.y));
        v[k][2] = make_float2(bf16_lo(u.z), bf16_hi(u.z));
        v[k][3] = make_float2(bf16_lo(u.w), bf16_hi(u.w));
#pragma unroll
        for (int j = 0; j < 4; ++j) sq2 = __ffma2_rn(v[k][j], v[k][j], sq2);
      }
      if (++st_rd == kStages) st_rd = 0;
      if (++st_wr == kStages) st_wr = 0;
      const float ssq = seg_sum(sq2.x + sq2.y, gl, G, lane);
      const float inv = rsqrtf(fmaxf(ssq, 1e-24f));                // 1 / max(|z|, 1e-12)
      const float2 inv2 = make_float2(inv, inv);
      float out_ss = 0.f;
#pragma unroll
      for (int k = 0; k < kV; ++k) {
        const int ch = gl + G * k;
        float2 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __ffma2_rn(__fmul2_rn(v[k][j], inv2), a[k][j], sh[k][j]);
          if (flags & CCDM_EPI_SILU) {                             // t * sigmoid(t) = h + h * tanh(h), h = t / 2
            const float2 h = __fmul2_rn(t, make_float2(0.5f, 0.5f));
            t = __ffma2_rn(h, make_float2(tanh_silu(h.x), tanh_silu(h.y)), h);
          }
          o[j] = t;
        }
        if (flags & CCDM_EPI_RESID) {
          o[0] = __fadd2_rn(o[0], make_float2(bf16_lo(ru[k].x), bf16_hi(ru[k].x)));
          o[1] = __fadd2_rn(o[1], make_float2(bf16_lo(ru[k].y), bf16_hi(ru[k].y)));
          o[2] = __fadd2_rn(o[2], make_float2(bf16_lo(ru[k].z), bf16_hi(ru[k].z)));
          o[3] = __fadd2_rn(o[3], make_float2(bf16_lo(ru[k].w), bf16_hi(ru[k].w)));
        }
        uint4 w;
        w.x = pack_bf16(o[0].x, o[0].y); w.y = pack_bf16(o[1].x, o[1].y);
        w.z = pack_bf16(o[2].x, o[2].y); w.w = pack_bf16(o[3].x, o[3].y);
        if (live && ch < nchunk) out[rowoff + ch] = w;
        if (flags & CCDM_EPI_SUMSQ_OUT) {
          const float a0 = bf16_lo(w.x), a1 = bf16_hi(w.x), a2 = bf16_lo(w.y), a3 = bf16_hi(w.y);
          const float a4 = bf16_lo(w.z), a5 = bf16_hi(w.z), a6 = bf16_lo(w.w), a7 = bf16_hi(w.w);
          if (ch < nchunk) out_ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
        }
      }
      if (flags & CCDM_EPI_SUMSQ_OUT) {
        out_ss = seg_sum(out_ss, gl, G, lane);
        if (live && gl == 0) out_rowss[(size_t)b * rows_per_sample + r] = out_ss;
      }
    }
    cp_async_wait<0>();                                            // the (empty) groups past r1, before the next item's prologue
  }
}

// ============================================================================ generator helpers (models/sngan.py)
__global__ void condbn_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ weight, const float* __restrict__ bias,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, int B, int C,
                                   float* __restrict__ ss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float r = rsqrtf(var[c] + eps);
  const float a = r * (weight ? weight[c] : 1.f) * (1.f + (gamma ? gamma[i] : 0.f));
  ss[(size_t)b * 2 * C + c] = a - 1.f;
  ss[(size_t)b * 2 * C + C + c] = (beta ? beta[i] : 0.f) + (bias ? bias[c] : 0.f) - mean[c] * a;
}

// out = act(x * (1 + scale[b]) + shift[b]); one thread per 8-channel vector
__global__ void __launch_bounds__(256) affine_act_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                         long long nvec, int nchunk, int rows_per_sample, int C,
                                                         const float* __restrict__ ss, int ss_ld, int ss_off, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nchunk;
    const int ch = (int)(i - row * nchunk);
    const long long b = row / rows_per_sample;
    float sc[8], sf[8];
    load8(ss + b * ss_ld + ss_off + ch * 8, sc);
    load8(ss + b * ss_ld + ss_off + C + ch * 8, sf);
    const uint4 u = __ldg(x + i);
    float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = fmaf(v[j], 1.f + sc[j], sf[j]);
      if (act == 1) v[j] = fmaxf(v[j], 0.f);
      else if (act == 2) v[j] = silu_f(v[j]);
    }
    out[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// ============================================================================ bottleneck softmax attention
// unet.py:228-240.  grid = (sample*head, query blocks of 64); one thread per query token, keys / values staged
// through shared memory 64 tokens at a time with a running (online) softmax.  The shipped configs have 9 or 16
// tokens here; larger token counts just loop.
template <int DH>
__global__ void __launch_bounds__(64) attention_small_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ out, int n, int heads,
                                                             float scale) {
  __shared__ float sk[64][DH + 1], sv[64][DH + 1];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int hid = heads * DH, ld = 3 * hid;
  const __nv_bfloat16* base = qkv + (long long)b * n * ld + h * DH;
  const int i = blockIdx.y * 64 + threadIdx.x;
  const bool active = i < n;
  float q[DH], acc[DH];
#pragma unroll
  for (int dd = 0; dd < DH; ++dd) {
    q[dd] = active ? __bfloat162float(base[(long long)i * ld + dd]) * scale : 0.f;
    acc[dd] = 0.f;
  }
  float m_run = -FLT_MAX, l_run = 0.f;
  for (int j0 = 0; j0 < n; j0 += 64) {
    const int rows = min(64, n - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < rows * DH; e += 64) {
      const int tok = e / DH, dd = e % DH;
      const __nv_bfloat16* p = base + (long long)(j0 + tok) * ld + dd;
      sk[tok][dd] = __bfloat162float(p[hid]);
      sv[tok][dd] = __bfloat162float(p[2 * hid]);
    }
    __syncthreads();
    for (int j = 0; j < rows; ++j) {
      float s = 0.f;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) s = fmaf(q[dd], sk[j][dd], s);
      const float m_new = fmaxf(m_run, s);
      const float corr = __expf(m_run - m_new), p = __expf(s - m_new);
      l_run = l_run * corr + p;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) acc[dd] = acc[dd] * corr + p * sv[j][dd];
      m_run = m_new;
    }
  }
  if (!active) return;
  const float inv = 1.f / l_run;
  __nv_bfloat16* o = out + ((long long)b * n + i) * hid + h * DH;
#pragma unroll
  for (int dd = 0; dd < DH; ++dd) o[dd] = __float2bfloat16(acc[dd] * inv);
}

// ============================================================================ embedding MLPs
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case CCDM_ACT_RELU: return fmaxf(v, 0.f);
    case CCDM_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // nn.GELU() default (erf)
    case CCDM_ACT_SILU: return v / (1.f + expf(-v));
    default: return v;
  }
}

// One warp per output feature: dot products over the batch, optional BatchNorm1d (batch or running stats), act.
__global__ void __launch_bounds__(128) linear_small_kernel(const float* __restrict__ x, int B, int in_dim,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int out_dim, const float* __restrict__ bn_w,
                                                           const float* __restrict__ bn_b, float* bn_mean,
                                                           float* bn_var, int bn_train, int act, float* __restrict__ y,
                                                           long long y_ld) {
  const int o = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= out_dim) return;
  const float* wr = w + (long long)o * in_dim;
  const float bo = bias ? bias[o] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* xr = x + (long long)b * in_dim;
    float acc = 0.f;
    for (int k = lane; k < in_dim; k += 32) acc = fmaf(xr[k], wr[k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    acc += bo;
    if (bn_w && bn_train) {
      s1 += acc;
      s2 += (double)acc * acc;
      if (lane == 0) y[(long long)b * y_ld + o] = acc;
    } else {
      if (bn_w) acc = (acc - bn_mean[o]) * rsqrtf(bn_var[o] + 1e-5f) * bn_w[o] + bn_b[o];
      if (lane == 0) y[(long long)b * y_ld + o] = act_apply(acc, act);
    }
  }
  if (bn_w && bn_train) {
    const double mean = s1 / B;
    double var = s2 / B - mean * mean;
    if (var < 0) var = 0;
    const float inv = rsqrtf((float)var + 1e-5f);
    __syncwarp();
    for (int b = lane; b < B; b += 32) {
      const float v = (y[(long long)b * y_ld + o] - (float)mean) * inv * bn_w[o] + bn_b[o];
      y[(long long)b * y_ld + o] = act_apply(v, act);
    }
    if (lane == 0) {
      bn_mean[o] = 0.9f * bn_mean[o] + 0.1f * (float)mean;
      bn_var[o] = 0.9f * bn_var[o] + 0.1f * (float)(var * B / (B > 1 ? B - 1 : 1));
    }
  }
}

// Eval-mode variant (running statistics or no BatchNorm): one lane per batch row, one warp per (output feature,
// 32-row slab), so the batch dimension is parallel instead of a serial loop.
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ x, int B, int in_dim,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          int out_dim, const float* __restrict__ bn_w,
                                                          const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                                          const float* __restrict__ bn_var, int act,
                                                          float* __restrict__ y, long long y_ld) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y * 32 + (threadIdx.x & 31);
  if (o >= out_dim || b >= B) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)b * in_dim);
  const float4* wr = reinterpret_cast<const float4*>(w + (long long)o * in_dim);
  float acc = 0.f;
  for (int k = 0; k < in_dim / 4; ++k) {
    const float4 a = __ldg(xr + k), c = __ldg(wr + k);
    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
  }
  acc += bias ? bias[o] : 0.f;
  if (bn_w) acc = (acc - bn_mean[o]) * rsqrtf(bn_var[o] + 1e-5f) * bn_w[o] + bn_b[o];
  y[(long long)b * y_ld + o] = act_apply(acc, act);
}

// nn.GroupNorm(groups, C) on a [B, C] matrix followed by an activation, in place (the learned label-embedding MLPs,
// models/resnet_y2h.py:143-173, models/resnet_y2cov.py:149-179): one warp per (row, group); biased variance, eps inside
// the square root (torch semantics), fp32.
__global__ void __launch_bounds__(256) groupnorm_rows_kernel(float* __restrict__ x, int B, int C, int groups,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int act) {
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= B * groups) return;
  const int b = wid / groups, g = wid % groups, cg = C / groups;
  float* xr = x + (long long)b * C + g * cg;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < cg; i += 32) {
    const float v = xr[i];
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  const float mean = s1 / cg;
  const float var = fmaxf(s2 / cg - mean * mean, 0.f);
  const float inv = rsqrtf(var + eps);
  for (int i = lane; i < cg; i += 32) {
    const int c = g * cg + i;
    xr[i] = act_apply((xr[i] - mean) * inv * gamma[c] + beta[c], act);
  }
}

// unet.py:107-115: sin | cos of t * exp(-log(1e4) * i / (half-1))
__global__ void time_features_kernel(const long long* __restrict__ t, int B, int dim, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx % half;
  const float k = logf(10000.f) / (float)(half - 1);
  const float ang = (float)t[b] * expf((float)i * -k);
  out[(long long)b * dim + i] = sinf(ang);
  out[(long long)b * dim + half + i] = cosf(ang);
}

__global__ void select_null_kernel(float* __restrict__ c, const uint8_t* __restrict__ keep, int all_null,
                                   const float* __restrict__ null_emb, int B, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim;
  if (all_null || !keep[b]) c[idx] = null_emb[idx % dim];
}

__global__ void silu_concat_kernel(const float* __restrict__ t_emb, int dt, const float* __restrict__ c_emb, int dc,
                                   int B, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ld = dt + dc;
  if (idx >= B * ld) return;
  const int b = idx / ld, j = idx % ld;
  const float v = j < dt ? t_emb[(long long)b * dt + j] : c_emb[(long long)b * dc + (j - dt)];
  out[idx] = __float2bfloat16(v / (1.f + expf(-v)));
}

}  // namespace ccdm

using namespace ccdm;

// ---------------------------------------------------------------------------- C ABI

extern "C" int ccdm_stem_im2row(const float* x, void* rowimg, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                void* stream) {
  CCDM_REQUIRE(x && rowimg && B > 0 && Cin >= 1 && Cin <= 4 && H > 0 && W > 0, CCDM_ERR_BAD_ARG,
               "stem_im2row: B=%d Cin=%d H=%d W=%d", B, Cin, H, W);
  const unsigned blocks = (unsigned)((long long)B * (H + 1));
  const size_t smem = (size_t)2 * Cin * (W + 6) * sizeof(float);
  CCDM_REQUIRE(smem <= 48 * 1024, CCDM_ERR_UNSUPPORTED_SHAPE, "stem_im2row: W=%d", W);
  if (Cin == 3) stem_im2row_kernel<3><<<blocks, 256, smem, (cudaStream_t)stream>>>(x, (uint4*)rowimg, Cin, H, W);
  else if (Cin == 1) stem_im2row_kernel<1><<<blocks, 256, smem, (cudaStream_t)stream>>>(x, (uint4*)rowimg, Cin, H, W);
  else stem_im2row_kernel<0><<<blocks, 256, smem, (cudaStream_t)stream>>>(x, (uint4*)rowimg, Cin, H, W);
  return after_launch("stem_im2row_kernel");
}

extern "C" int ccdm_stem_pack(const float* w, void* wpacked, int32_t Cout, int32_t Cin, int32_t n_rows, void* stream) {
  CCDM_REQUIRE(w && wpacked && Cout > 0 && Cin >= 1 && Cin <= 4 && n_rows >= Cout, CCDM_ERR_BAD_ARG, "stem_pack: bad args");
  const int total = n_rows * 256;
  stem_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wpacked, Cout, Cin, n_rows);
  return after_launch("stem_pack_kernel");
}

extern "C" int ccdm_head_conv1(const void* x, const float* w, const float* bias, float* out, int32_t B, int32_t H,
                               int32_t W, int32_t Cin, int32_t Cout, void* stream) {
  CCDM_REQUIRE(x && w && bias && out, CCDM_ERR_BAD_ARG, "head_conv1: null pointer");
  CCDM_REQUIRE(Cin % 8 == 0 && Cout >= 1 && Cout <= 8 && Cin <= 1024, CCDM_ERR_UNSUPPORTED_SHAPE,
               "head_conv1: Cin=%d Cout=%d", Cin, Cout);
  const long long npix = (long long)B * H * W;
  const size_t smem = (size_t)(Cout * Cin + Cout) * sizeof(float);
  head_conv1_kernel<<<(unsigned)((npix + 255) / 256), 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, w, bias, out, npix, H * W, Cin, Cout);
  return after_launch("head_conv1_kernel");
}

extern "C" int ccdm_linattn_fold(const float* w_out, const float* ctx, void* wfold, int32_t B, int32_t C,
                                 int32_t n_rows, int32_t heads, void* stream) {
  CCDM_REQUIRE(w_out && ctx && wfold && B > 0 && C > 0 && n_rows >= C && heads > 0 && heads <= 8, CCDM_ERR_BAD_ARG,
               "linattn_fold: bad args");
  const int hid = heads * 32;
  dim3 grid(B, (n_rows + 31) / 32);
  linattn_fold_kernel<<<grid, hid, (size_t)32 * hid * sizeof(float), (cudaStream_t)stream>>>(
      w_out, ctx, (__nv_bfloat16*)wfold, C, n_rows, heads);
  return after_launch("linattn_fold_kernel");
}

extern "C" int ccdm_rmsnorm_act(const void* z, void* out, int64_t rows, int32_t C, int32_t rows_per_sample,
                                const float* gain, float gain_mul, const float* scale_shift, int32_t ss_ld,
                                int32_t ss_off, const void* resid, float* out_rowss, uint32_t flags, void* stream) {
  CCDM_REQUIRE(z && out && gain && rows > 0 && rows_per_sample > 0, CCDM_ERR_BAD_ARG, "rmsnorm_act: bad args");
  CCDM_REQUIRE(C % 8 == 0 && C <= 768, CCDM_ERR_UNSUPPORTED_SHAPE, "rmsnorm_act: C=%d (multiple of 8, <= 768)", C);
  CCDM_REQUIRE(!(flags & CCDM_EPI_SS) || scale_shift, CCDM_ERR_BAD_ARG, "rmsnorm_act: scale/shift pointer");
  CCDM_REQUIRE(!(flags & CCDM_EPI_RESID) || resid, CCDM_ERR_BAD_ARG, "rmsnorm_act: resid pointer");
  CCDM_REQUIRE(!(flags & CCDM_EPI_SUMSQ_OUT) || out_rowss, CCDM_ERR_BAD_ARG, "rmsnorm_act: out_rowss pointer");
  CCDM_REQUIRE(rows % rows_per_sample == 0, CCDM_ERR_BAD_ARG, "rmsnorm_act: rows=%lld rows_per_sample=%d",
               (long long)rows, rows_per_sample);
  const int B = (int)(rows / rows_per_sample);
  int kv, G;
  row_lane_plan(C / 8, 3, &kv, &G);
  // persistent grid: one resident wave of CTAs; work items = (sample, slab of >= 64 rows), ~6 items per CTA
  cudaStream_t st = (cudaStream_t)stream;
  static int occ[4] = {0, 0, 0, 0};
  const int ring_bytes = rms_stages(kv) * kv * 2 * 256 * 16;
  if (!occ[kv]) {
    int o = 1;
    const int rb = ring_bytes;
    cudaError_t e;
    if (kv == 1) {
      e = cudaFuncSetAttribute(rmsnorm_act_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rb);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rmsnorm_act_kernel<1>, 256, rb);
    } else if (kv == 2) {
      e = cudaFuncSetAttribute(rmsnorm_act_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, rb);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rmsnorm_act_kernel<2>, 256, rb);
    } else {
      e = cudaFuncSetAttribute(rmsnorm_act_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, rb);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rmsnorm_act_kernel<3>, 256, rb);
    }
    if (e != cudaSuccess) return cuda_fail(e, "rmsnorm_act: cudaFuncSetAttribute");
    occ[kv] = o > 0 ? o : 1;
  }
  const int slots = num_sms() * occ[kv];
  int per_sample = (slots * 6 + B - 1) / B;
  const int max_split = (rows_per_sample + 63) / 64;
  if (per_sample > max_split) per_sample = max_split;
  if (per_sample < 1) per_sample = 1;
  const int rows_per_item = (rows_per_sample + per_sample - 1) / per_sample;
  per_sample = (rows_per_sample + rows_per_item - 1) / rows_per_item;
  const int n_items = per_sample * B;
  const int grid = n_items < slots ? n_items : slots;
#define CCDM_RMS(K)                                                                                                   \
  rmsnorm_act_kernel<K><<<grid, 256, ring_bytes, st>>>((const uint4*)z, (uint4*)out, rows_per_sample, rows_per_item, per_sample, \
                                              n_items, C, G, gain, gain_mul, scale_shift, ss_ld, ss_off,               \
                                              (const uint4*)resid, out_rowss, flags)
  if (kv == 1) CCDM_RMS(1);
  else if (kv == 2) CCDM_RMS(2);
  else CCDM_RMS(3);
#undef CCDM_RMS
  return after_launch("rmsnorm_act_kernel");
}

extern "C" int ccdm_condbn_coef(const float* gamma, const float* beta, const float* weight, const float* bias,
                                const float* mean, const float* var, float eps, int32_t B, int32_t C, float* ss,
                                void* stream) {
  CCDM_REQUIRE(mean && var && ss && B > 0 && C > 0, CCDM_ERR_BAD_ARG, "condbn_coef: bad args");
  const int n = B * C;
  condbn_coef_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gamma, beta, weight, bias, mean, var, eps, B, C, ss);
  return after_launch("condbn_coef_kernel");
}

extern "C" int ccdm_affine_act(const void* x, void* out, int64_t rows, int32_t C, int32_t rows_per_sample,
                               const float* scale_shift, int32_t ss_ld, int32_t ss_off, int32_t act, void* stream) {
  CCDM_REQUIRE(x && out && scale_shift && rows > 0 && rows_per_sample > 0 && rows % rows_per_sample == 0, CCDM_ERR_BAD_ARG,
               "affine_act: bad args");
  CCDM_REQUIRE(C > 0 && C % 8 == 0 && act >= 0 && act <= 2, CCDM_ERR_UNSUPPORTED_SHAPE, "affine_act: C=%d act=%d", C, act);
  const long long nvec = rows * (C / 8);
  long long blocks = (nvec + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  affine_act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, nvec, C / 8,
                                                                        rows_per_sample, C, scale_shift, ss_ld, ss_off, act);
  return after_launch("affine_act_kernel");
}

extern "C" int ccdm_attention_small(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t dim_head,
                                    float scale, void* stream) {
  CCDM_REQUIRE(qkv && out && B > 0 && n >= 1 && heads >= 1, CCDM_ERR_BAD_ARG, "attention_small: bad args");
  dim3 grid(B * heads, (n + 63) / 64);
  const __nv_bfloat16* in = (const __nv_bfloat16*)qkv;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  cudaStream_t s = (cudaStream_t)stream;
  switch (dim_head) {
    case 16: attention_small_kernel<16><<<grid, 64, 0, s>>>(in, o, n, heads, scale); break;
    case 32: attention_small_kernel<32><<<grid, 64, 0, s>>>(in, o, n, heads, scale); break;
    case 64: attention_small_kernel<64><<<grid, 64, 0, s>>>(in, o, n, heads, scale); break;
    default:
      CCDM_REQUIRE(false, CCDM_ERR_UNSUPPORTED_SHAPE, "attention_small: dim_head=%d (supported: 16, 32, 64)", dim_head);
  }
  return after_launch("attention_small_kernel");
}

extern "C" int ccdm_linear_small(const float* x, int32_t B, int32_t in_dim, const float* w, const float* bias,
                                 int32_t out_dim, const float* bn_w, const float* bn_b, float* bn_mean, float* bn_var,
                                 int32_t bn_train, int32_t act, float* y, int64_t y_ld, void* stream) {
  CCDM_REQUIRE(x && w && y && B > 0 && in_dim > 0 && out_dim > 0 && y_ld >= out_dim, CCDM_ERR_BAD_ARG,
               "linear_small: bad args");
  CCDM_REQUIRE(!bn_w || (bn_b && bn_mean && bn_var), CCDM_ERR_BAD_ARG, "linear_small: incomplete BatchNorm pointers");
  if (!(bn_w && bn_train) && in_dim % 4 == 0) {
    dim3 grid((out_dim + 7) / 8, (B + 31) / 32);
    linear_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, in_dim, w, bias, out_dim, bn_w, bn_b, bn_mean,
                                                               bn_var, act, y, y_ld);
    return after_launch("linear_rows_kernel");
  }
  linear_small_kernel<<<(out_dim + 3) / 4, 128, 0, (cudaStream_t)stream>>>(x, B, in_dim, w, bias, out_dim, bn_w, bn_b,
                                                                           bn_mean, bn_var, bn_train, act, y, y_ld);
  return after_launch("linear_small_kernel");
}

extern "C" int ccdm_groupnorm_rows(float* x, int32_t B, int32_t C, int32_t groups, const float* gamma, const float* beta,
                                   float eps, int32_t act, void* stream) {
  CCDM_REQUIRE(x && gamma && beta && B > 0 && C > 0 && groups > 0 && C % groups == 0, CCDM_ERR_BAD_ARG,
               "groupnorm_rows: bad args (B=%d C=%d groups=%d)", B, C, groups);
  const int warps = B * groups;
  groupnorm_rows_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, B, C, groups, gamma, beta, eps, act);
  return after_launch("groupnorm_rows_kernel");
}

extern "C" int ccdm_time_features(const int64_t* t, int32_t B, int32_t dim, float* out, void* stream) {
  CCDM_REQUIRE(t && out && B > 0 && dim >= 4 && dim % 2 == 0, CCDM_ERR_BAD_ARG, "time_features: bad args");
  const int n = B * (dim / 2);
  time_features_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)t, B, dim, out);
  return after_launch("time_features_kernel");
}

extern "C" int ccdm_select_null(float* c, const uint8_t* keep, int32_t all_null, const float* null_emb, int32_t B,
                                int32_t dim, void* stream) {
  CCDM_REQUIRE(c && null_emb && (keep || all_null), CCDM_ERR_BAD_ARG, "select_null: bad args");
  const int n = B * dim;
  select_null_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c, keep, all_null, null_emb, B, dim);
  return after_launch("select_null_kernel");
}

extern "C" int ccdm_silu_concat_bf16(const float* t_emb, int32_t dt, const float* c_emb, int32_t dc, int32_t B,
                                     void* out, void* stream) {
  CCDM_REQUIRE(t_emb && c_emb && out && B > 0, CCDM_ERR_BAD_ARG, "silu_concat: bad args");
  const int n = B * (dt + dc);
  silu_concat_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t_emb, dt, c_emb, dc, B, (__nv_bfloat16*)out);
  return after_launch("silu_concat_kernel");
}
