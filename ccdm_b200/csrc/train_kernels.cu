// Training-only helper kernels (HBM-bound or tiny): bias gradients, the softmaxes of LinearAttention and their
// backward, the bottleneck attention backward, head / stem gradients.  Reference: autograd of
// CCDM_unified/models/unet.py:202-216 (LinearAttention), :228-240 (Attention), :271,:348 (stem / head convs).
#include <cfloat>

#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

// ---------------------------------------------------------------------------- column sums (bias gradients)
// out[c] += sum_rows x[row][c];  x bf16 [rows][C], C % 8 == 0, C <= 2048
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const uint4* __restrict__ x, long long rows, int C,
                                                          float* __restrict__ out) {
  __shared__ float acc_s[2048];
  const int nchunk = C >> 3;
  const int lanes = 256 / nchunk;                                  // row lanes per block
  const int tid = threadIdx.x;
  for (int i = tid; i < C; i += 256) acc_s[i] = 0.f;
  __syncthreads();
  const int ch = tid % nchunk, rl = tid / nchunk;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (long long r = (long long)blockIdx.x * lanes + rl; r < rows; r += (long long)gridDim.x * lanes) {
      const uint4 u = __ldg(x + r * nchunk + ch);
      a[0] += bf16_lo(u.x); a[1] += bf16_hi(u.x); a[2] += bf16_lo(u.y); a[3] += bf16_hi(u.y);
      a[4] += bf16_lo(u.z); a[5] += bf16_hi(u.z); a[6] += bf16_lo(u.w); a[7] += bf16_hi(u.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&acc_s[ch * 8 + j], a[j]);
  }
  __syncthreads();
  for (int i = tid; i < C; i += 256) atomicAdd(&out[i], acc_s[i]);
}

// ---------------------------------------------------------------------------- LinearAttention forward pieces
// kmax[b][c] = max over tokens of qkv[b][tok][128 + c], c < 128 (the softmax-over-tokens shift, unet.py:208).
// grid = (token slabs, B): each CTA reduces its slab and merges with an ordered-int atomic max (kmax pre-filled with
// -FLT_MAX by kmax_init_kernel).
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void kmax_init_kernel(float* __restrict__ kmax, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) kmax[i] = -FLT_MAX;
}

__global__ void __launch_bounds__(256) linattn_kmax_kernel(const __nv_bfloat16* __restrict__ qkv, int n, int slab,
                                                           float* __restrict__ kmax) {
  __shared__ float red[16][128];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int ch = tid & 15, rl = tid >> 4;
  const uint4* base = reinterpret_cast<const uint4*>(qkv + (size_t)b * n * 384 + 128);
  const int t0 = blockIdx.x * slab, t1 = min(t0 + slab, n);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -FLT_MAX;
  for (int t = t0 + rl; t < t1; t += 16) {
    const uint4 u = __ldg(base + (size_t)t * 48 + ch);
    m[0] = fmaxf(m[0], bf16_lo(u.x)); m[1] = fmaxf(m[1], bf16_hi(u.x));
    m[2] = fmaxf(m[2], bf16_lo(u.y)); m[3] = fmaxf(m[3], bf16_hi(u.y));
    m[4] = fmaxf(m[4], bf16_lo(u.z)); m[5] = fmaxf(m[5], bf16_hi(u.z));
    m[6] = fmaxf(m[6], bf16_lo(u.w)); m[7] = fmaxf(m[7], bf16_hi(u.w));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][ch * 8 + j] = m[j];
  __syncthreads();
  if (tid < 128) {
    float v = red[0][tid];
#pragma unroll
    for (int r = 1; r < 16; ++r) v = fmaxf(v, red[r][tid]);
    atomic_max_float(&kmax[(size_t)b * 128 + tid], v);
  }
}

// in place on qkv [B][n][384]:  q <- softmax over each head's 32 channels * scale;  k <- exp(k - kmax[b][c])
// one warp per token, lane l owns channels 4l .. 4l+3 of q and of k (8 lanes per head)
__global__ void __launch_bounds__(256) linattn_prep_kernel(__nv_bfloat16* __restrict__ qkv, long long tokens, int n,
                                                           const float* __restrict__ kmax, float scale) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= tokens) return;
  const int b = (int)(row / n);
  uint2* qp = reinterpret_cast<uint2*>(qkv + row * 384) + lane;
  uint2* kp = qp + 32;
  const uint2 qu = *qp, ku = *kp;
  float q[4] = {bf16_lo(qu.x), bf16_hi(qu.x), bf16_lo(qu.y), bf16_hi(qu.y)};
  float k[4] = {bf16_lo(ku.x), bf16_hi(ku.x), bf16_lo(ku.y), bf16_hi(ku.y)};
  float m = fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3]));
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[j] = __expf(q[j] - m);
    s += q[j];
  }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float inv = scale / s;
  const float4 km = *reinterpret_cast<const float4*>(kmax + (size_t)b * 128 + lane * 4);
  k[0] = __expf(k[0] - km.x); k[1] = __expf(k[1] - km.y); k[2] = __expf(k[2] - km.z); k[3] = __expf(k[3] - km.w);
  *qp = make_uint2(pack_bf16(q[0] * inv, q[1] * inv), pack_bf16(q[2] * inv, q[3] * inv));
  *kp = make_uint2(pack_bf16(k[0], k[1]), pack_bf16(k[2], k[3]));
}

// Per-sample block-diagonal 128x128 bf16 weights for ccdm_tapgemm (w_batch_rows = 128), from [B][4][32][32] fp32
// matrices m[b][h][d][e]:   transpose == 0: w[b][h*32+e][h*32+d] = m[b][h][d][e] * rs(d)     (out = m^T . in)
//                           transpose != 0: w[b][h*32+d][h*32+e] = m[b][h][d][e] * rs(d)     (out = m   . in)
// rs(d) = 1 / row_div[b][h*32+d] when row_div is given.
// diag_only: write just the four 32x32 diagonal blocks (idx enumerates [b][row][32 columns of the row's head]); the
// off-diagonal zeros of a persistent buffer are written once by the caller.
__global__ void __launch_bounds__(256) linattn_pack_blockdiag_kernel(const float* __restrict__ m,
                                                                     const float* __restrict__ row_div,
                                                                     int transpose, __nv_bfloat16* __restrict__ w,
                                                                     long long total, int diag_only) {
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  if (diag_only) {                                                 // idx = (b*128 + row)*32 + j  ->  full index
    const int j = (int)(idx & 31);
    const long long br = idx >> 5;
    const int row_ = (int)(br & 127);
    idx = (br << 7) + (row_ & ~31) + j;
  }
  const int col = (int)(idx & 127), row = (int)((idx >> 7) & 127);
  const long long b = idx >> 14;
  float v = 0.f;
  if ((col >> 5) == (row >> 5)) {
    const int h = row >> 5;
    const int d = transpose ? (row & 31) : (col & 31);
    const int e = transpose ? (col & 31) : (row & 31);
    v = m[((b * 4 + h) * 32 + d) * 32 + e];
    if (row_div) v /= row_div[b * 128 + h * 32 + d];
  }
  w[idx] = __float2bfloat16(v);
}

// c[b][h*32+d] = sum_e dctx[b][h][d][e] * ctx[b][h][d][e] / S[b][h*32+d]      (the -dS term of the k softmax)
__global__ void __launch_bounds__(128) linattn_bwd_rowdot_kernel(const float* __restrict__ ctx,
                                                                 const float* __restrict__ dctx,
                                                                 const float* __restrict__ S, float* __restrict__ c,
                                                                 long long rows) {
  const long long r = (long long)blockIdx.x * 128 + threadIdx.x;      // r = b*128 + h*32 + d
  if (r >= rows) return;
  const float4* a = reinterpret_cast<const float4*>(ctx + r * 32);
  const float4* g = reinterpret_cast<const float4*>(dctx + r * 32);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 x = a[j], y = g[j];
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  c[r] = acc / S[r];
}

// in place on dpre [B][n][384] = [dq_sm | dp_term | dv]  ->  [dq_raw | dk_raw | dv]   with qkv = [q_sm | p | v]:
//   dq_raw = q_sm * (dq_sm - <q_sm, dq_sm>_head / scale)        dk_raw = p * (dp_term - c[b][ch])
__global__ void __launch_bounds__(256) linattn_bwd_finish_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ dpre, long long tokens,
                                                                 int n, const float* __restrict__ c, float inv_scale) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= tokens) return;
  const int b = (int)(row / n);
  const uint2* qp = reinterpret_cast<const uint2*>(qkv + row * 384) + lane;
  uint2* dq = reinterpret_cast<uint2*>(dpre + row * 384) + lane;
  const uint2 qu = __ldg(qp), pu = __ldg(qp + 32);
  const uint2 gq = *dq, gp = *(dq + 32);
  const float q[4] = {bf16_lo(qu.x), bf16_hi(qu.x), bf16_lo(qu.y), bf16_hi(qu.y)};
  const float p[4] = {bf16_lo(pu.x), bf16_hi(pu.x), bf16_lo(pu.y), bf16_hi(pu.y)};
  const float a[4] = {bf16_lo(gq.x), bf16_hi(gq.x), bf16_lo(gq.y), bf16_hi(gq.y)};
  const float t[4] = {bf16_lo(gp.x), bf16_hi(gp.x), bf16_lo(gp.y), bf16_hi(gp.y)};
  float dot = q[0] * a[0] + q[1] * a[1] + q[2] * a[2] + q[3] * a[3];
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
  dot *= inv_scale;
  const float4 cc = *reinterpret_cast<const float4*>(c + (size_t)b * 128 + lane * 4);
  *dq = make_uint2(pack_bf16(q[0] * (a[0] - dot), q[1] * (a[1] - dot)), pack_bf16(q[2] * (a[2] - dot), q[3] * (a[3] - dot)));
  *(dq + 32) = make_uint2(pack_bf16(p[0] * (t[0] - cc.x), p[1] * (t[1] - cc.y)),
                          pack_bf16(p[2] * (t[2] - cc.z), p[3] * (t[3] - cc.w)));
}

// ---------------------------------------------------------------------------- bottleneck attention backward
// One CTA per (sample, head).  Shared memory holds q, k, v, dO of the head as fp32 [n][DH] plus per-query
// (max, 1/sum, D = <dO, O>).  Phase A (thread = query): statistics and dq.  Phase B (thread = key): dk, dv.
template <int DH>
__global__ void __launch_bounds__(128) attention_small_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                  const __nv_bfloat16* __restrict__ dout,
                                                                  __nv_bfloat16* __restrict__ dqkv, int n, int heads,
                                                                  float scale) {
  extern __shared__ float att_smem[];
  float* sq = att_smem;
  float* sk = sq + (size_t)n * DH;
  float* sv = sk + (size_t)n * DH;
  float* sd = sv + (size_t)n * DH;
  float* st = sd + (size_t)n * DH;                                 // [n][3]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int hid = heads * DH, ld = 3 * hid;
  const __nv_bfloat16* base = qkv + (size_t)b * n * ld + h * DH;
  const __nv_bfloat16* dob = dout + (size_t)b * n * hid + h * DH;
  for (int e = threadIdx.x; e < n * DH; e += 128) {
    const int tok = e / DH, dd = e % DH;
    const __nv_bfloat16* p = base + (size_t)tok * ld + dd;
    sq[e] = __bfloat162float(p[0]);
    sk[e] = __bfloat162float(p[hid]);
    sv[e] = __bfloat162float(p[2 * hid]);
    sd[e] = __bfloat162float(dob[(size_t)tok * hid + dd]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 128) {
    float q[DH], g[DH];
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) {
      q[dd] = sq[i * DH + dd] * scale;
      g[dd] = sd[i * DH + dd];
    }
    float m = -FLT_MAX;
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) s = fmaf(q[dd], sk[j * DH + dd], s);
      m = fmaxf(m, s);
    }
    float l = 0.f, dsum = 0.f;                                      // dsum = sum_j e_ij * <dO_i, v_j>
    for (int j = 0; j < n; ++j) {
      float s = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) {
        s = fmaf(q[dd], sk[j * DH + dd], s);
        gv = fmaf(g[dd], sv[j * DH + dd], gv);
      }
      const float e = __expf(s - m);
      l += e;
      dsum = fmaf(e, gv, dsum);
    }
    const float inv = 1.f / l;
    const float D = dsum * inv;                                     // <dO_i, O_i>
    float dq[DH];
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) dq[dd] = 0.f;
    for (int j = 0; j < n; ++j) {
      float s = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) {
        s = fmaf(q[dd], sk[j * DH + dd], s);
        gv = fmaf(g[dd], sv[j * DH + dd], gv);
      }
      const float ds = __expf(s - m) * inv * (gv - D);
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) dq[dd] = fmaf(ds, sk[j * DH + dd], dq[dd]);
    }
    st[i * 3] = m;
    st[i * 3 + 1] = inv;
    st[i * 3 + 2] = D;
    __nv_bfloat16* o = dqkv + ((size_t)b * n + i) * ld + h * DH;
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) o[dd] = __float2bfloat16(dq[dd] * scale);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += 128) {
    float kk[DH], vv[DH], dk[DH], dv[DH];
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) {
      kk[dd] = sk[j * DH + dd];
      vv[dd] = sv[j * DH + dd];
      dk[dd] = dv[dd] = 0.f;
    }
    for (int i = 0; i < n; ++i) {
      float s = 0.f, gv = 0.f;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) {
        s = fmaf(sq[i * DH + dd], kk[dd], s);
        gv = fmaf(sd[i * DH + dd], vv[dd], gv);
      }
      const float pij = __expf(s * scale - st[i * 3]) * st[i * 3 + 1];
      const float ds = pij * (gv - st[i * 3 + 2]) * scale;
#pragma unroll
      for (int dd = 0; dd < DH; ++dd) {
        dv[dd] = fmaf(pij, sd[i * DH + dd], dv[dd]);
        dk[dd] = fmaf(ds, sq[i * DH + dd], dk[dd]);
      }
    }
    __nv_bfloat16* o = dqkv + ((size_t)b * n + j) * ld + h * DH;
#pragma unroll
    for (int dd = 0; dd < DH; ++dd) {
      o[hid + dd] = __float2bfloat16(dk[dd]);
      o[2 * hid + dd] = __float2bfloat16(dv[dd]);
    }
  }
}

// ---------------------------------------------------------------------------- head 1x1 conv backward
// dh[b,p,c] = sum_o dout[b,o,p] * w[o][c];  dw[o][c] += sum_{b,p} dout[b,o,p] * h[b,p,c];  db[o] += sum dout[b,o,p]
// dout fp32 NCHW (Cout <= 4), h / dh bf16 NHWC.
__global__ void __launch_bounds__(256) head_conv1_bwd_kernel(const float* __restrict__ dout, const uint4* __restrict__ h,
                                                             const float* __restrict__ w, uint4* __restrict__ dh,
                                                             float* __restrict__ dw, float* __restrict__ db, int B, int HW,
                                                             int Cin, int Cout) {
  __shared__ float dw_s[4 * 2048];
  __shared__ float db_s[4];
  const int nchunk = Cin >> 3;
  const int lanes = 256 / nchunk;
  const int tid = threadIdx.x;
  for (int i = tid; i < Cout * Cin; i += 256) dw_s[i] = 0.f;
  if (tid < 4) db_s[tid] = 0.f;
  __syncthreads();
  const int ch = tid % nchunk, pl = tid / nchunk;
  float wr[4][8], acc[4][8], bacc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wr[o][j] = o < Cout ? w[o * Cin + ch * 8 + j] : 0.f;
      acc[o][j] = 0.f;
    }
  if (pl < lanes) {
    const long long total = (long long)B * HW;
    for (long long p = (long long)blockIdx.x * lanes + pl; p < total; p += (long long)gridDim.x * lanes) {
      const long long b = p / HW, pix = p - b * HW;
      float g[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) g[o] = o < Cout ? __ldg(dout + (b * Cout + o) * HW + pix) : 0.f;
      const uint4 u = __ldg(h + p * nchunk + ch);
      const float hv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                           bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          d[j] = fmaf(g[o], wr[o][j], d[j]);
          acc[o][j] = fmaf(g[o], hv[j], acc[o][j]);
        }
      }
      if (ch == 0)
#pragma unroll
        for (int o = 0; o < 4; ++o) bacc[o] += g[o];
      dh[p * nchunk + ch] = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (o < Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&dw_s[o * Cin + ch * 8 + j], acc[o][j]);
        if (ch == 0) atomicAdd(&db_s[o], bacc[o]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Cout * Cin; i += 256) atomicAdd(&dw[i], dw_s[i]);
  if (tid < Cout) atomicAdd(&db[tid], db_s[tid]);
}

// dW[n][c][kr][ks] (=|+=) packed[n][(kr/2)*64 + (kr%2)*7*Cin + ks*Cin + c]     (inverse of ccdm_stem_pack)
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ dw, int Cout, int Cin,
                                         int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Cin * 49) return;
  const int ks = idx % 7, kr = (idx / 7) % 7, c = (idx / 49) % Cin, n = idx / (49 * Cin);
  const float v = packed[(size_t)n * 256 + (kr >> 1) * 64 + (kr & 1) * 7 * Cin + ks * Cin + c];
  dw[idx] = accumulate ? dw[idx] + v : v;
}


// Batch construction on the device (trainer.py:461-482 `process_images`): gather B images of a device-resident uint8
// dataset [N][C][H][W] by index, apply the per-sample augmentation (rot90 by k quarter turns as np.rot90 over (H, W), then
// horizontal flip, then vertical flip -- the order of trainer.py:468-474) and normalise to [0, 1] fp32 NCHW (utils.py:182-186).
// aug[b]: bits 0-1 = k (quarter turns, needs H == W when non-zero), bit 2 = horizontal flip, bit 3 = vertical flip.
__global__ void gather_augment_u8_kernel(const uint8_t* __restrict__ images, const long long* __restrict__ idx,
                                         const uint8_t* __restrict__ aug, float* __restrict__ out, int B, int C, int H, int W,
                                         long long n_images) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    const unsigned a = aug ? aug[b] : 0u;
    int yy = (a & 8u) ? H - 1 - y : y;                    // undo the vertical flip
    int xx = (a & 4u) ? W - 1 - x : x;                    // undo the horizontal flip
    const unsigned k = a & 3u;                            // np.rot90: out[i][j] = in[j][n-1-i] for k = 1
    int sy = yy, sx = xx;
    if (k == 1) { sy = xx; sx = W - 1 - yy; }
    else if (k == 2) { sy = H - 1 - yy; sx = W - 1 - xx; }
    else if (k == 3) { sy = H - 1 - xx; sx = yy; }
    long long src = idx[b];
    src = src < 0 ? 0 : (src >= n_images ? n_images - 1 : src);
    out[i] = (float)images[((src * C + c) * H + sy) * W + sx] * (1.f / 255.f);
  }
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_colsum_bf16(const void* x, int64_t rows, int32_t C, float* out, void* stream) {
  CCDM_REQUIRE(x && out && rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, CCDM_ERR_BAD_ARG, "colsum_bf16: bad args");
  const int lanes = 256 / (C / 8);
  long long blocks = (rows + (long long)lanes * 16 - 1) / ((long long)lanes * 16);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  colsum_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, rows, C, out);
  return after_launch("colsum_bf16_kernel");
}

extern "C" int ccdm_linattn_prep(void* qkv, int32_t B, int32_t n, float* kmax, float scale, void* stream) {
  CCDM_REQUIRE(qkv && kmax && B > 0 && n > 0, CCDM_ERR_BAD_ARG, "linattn_prep: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  kmax_init_kernel<<<(B * 128 + 255) / 256, 256, 0, s>>>(kmax, B * 128);
  int rc = after_launch("kmax_init_kernel");
  if (rc != CCDM_OK) return rc;
  int slabs = (num_sms() * 4 + B - 1) / B;                         // a few CTAs per SM, >= 64 tokens each
  if (slabs > (n + 63) / 64) slabs = (n + 63) / 64;
  if (slabs < 1) slabs = 1;
  const int slab = (n + slabs - 1) / slabs;
  slabs = (n + slab - 1) / slab;
  linattn_kmax_kernel<<<dim3(slabs, B), 256, 0, s>>>((const __nv_bfloat16*)qkv, n, slab, kmax);
  rc = after_launch("linattn_kmax_kernel");
  if (rc != CCDM_OK) return rc;
  const long long tokens = (long long)B * n;
  linattn_prep_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, s>>>((__nv_bfloat16*)qkv, tokens, n, kmax, scale);
  return after_launch("linattn_prep_kernel");
}

extern "C" int ccdm_linattn_pack_blockdiag(const float* m, const float* row_div, int32_t transpose, void* w, int32_t B,
                                           void* stream) {
  CCDM_REQUIRE(m && w && B > 0, CCDM_ERR_BAD_ARG, "linattn_pack_blockdiag: bad args");
  const int diag_only = (transpose & 2) ? 1 : 0;                   // bit 1 of `transpose`: diagonal blocks only
  const long long total = (long long)B * 128 * (diag_only ? 32 : 128);
  linattn_pack_blockdiag_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      m, row_div, transpose & 1, (__nv_bfloat16*)w, total, diag_only);
  return after_launch("linattn_pack_blockdiag_kernel");
}

extern "C" int ccdm_linattn_bwd_rowdot(const float* ctx, const float* dctx, const float* S, float* c, int32_t B,
                                       void* stream) {
  CCDM_REQUIRE(ctx && dctx && S && c && B > 0, CCDM_ERR_BAD_ARG, "linattn_bwd_rowdot: bad args");
  const long long rows = (long long)B * 128;
  linattn_bwd_rowdot_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ctx, dctx, S, c, rows);
  return after_launch("linattn_bwd_rowdot_kernel");
}

extern "C" int ccdm_linattn_bwd_finish(const void* qkv, void* dpre, int32_t B, int32_t n, const float* c, float scale,
                                       void* stream) {
  CCDM_REQUIRE(qkv && dpre && c && B > 0 && n > 0 && scale > 0.f, CCDM_ERR_BAD_ARG, "linattn_bwd_finish: bad args");
  const long long tokens = (long long)B * n;
  linattn_bwd_finish_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)qkv, (__nv_bfloat16*)dpre, tokens, n, c, 1.f / scale);
  return after_launch("linattn_bwd_finish_kernel");
}

extern "C" int ccdm_attention_small_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t n,
                                        int32_t heads, int32_t dim_head, float scale, void* stream) {
  CCDM_REQUIRE(qkv && dout && dqkv && B > 0 && n >= 1 && heads >= 1, CCDM_ERR_BAD_ARG, "attention_small_bwd: bad args");
  const size_t smem = ((size_t)4 * n * dim_head + (size_t)3 * n) * sizeof(float);
  CCDM_REQUIRE(smem <= 200 * 1024, CCDM_ERR_UNSUPPORTED_SHAPE, "attention_small_bwd: %d tokens x %d do not fit shared memory",
               n, dim_head);
  cudaStream_t s = (cudaStream_t)stream;
#define CCDM_ATT_BWD(D)                                                                                               \
  {                                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(attention_small_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         200 * 1024);                                                                 \
    if (e != cudaSuccess) return cuda_fail(e, "attention_small_bwd: cudaFuncSetAttribute");                           \
    attention_small_bwd_kernel<D><<<B * heads, 128, smem, s>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, \
                                                               (__nv_bfloat16*)dqkv, n, heads, scale);                \
  }
  switch (dim_head) {
    case 16: CCDM_ATT_BWD(16) break;
    case 32: CCDM_ATT_BWD(32) break;
    case 64: CCDM_ATT_BWD(64) break;
    default:
      CCDM_REQUIRE(false, CCDM_ERR_UNSUPPORTED_SHAPE, "attention_small_bwd: dim_head=%d (supported: 16, 32, 64)", dim_head);
  }
#undef CCDM_ATT_BWD
  return after_launch("attention_small_bwd_kernel");
}

extern "C" int ccdm_head_conv1_bwd(const float* dout, const void* h, const float* w, void* dh, float* dw, float* db,
                                   int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream) {
  CCDM_REQUIRE(dout && h && w && dh && dw && db && B > 0 && H > 0 && W > 0, CCDM_ERR_BAD_ARG, "head_conv1_bwd: bad args");
  CCDM_REQUIRE(Cin % 8 == 0 && Cin <= 2048 && Cout >= 1 && Cout <= 4, CCDM_ERR_UNSUPPORTED_SHAPE,
               "head_conv1_bwd: Cin=%d Cout=%d", Cin, Cout);
  const int lanes = 256 / (Cin / 8);
  const long long total = (long long)B * H * W;
  long long blocks = (total + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  head_conv1_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, (const uint4*)h, w, (uint4*)dh, dw, db, B,
                                                                            H * W, Cin, Cout);
  return after_launch("head_conv1_bwd_kernel");
}

extern "C" int ccdm_stem_unpack_wgrad(const float* packed, float* dw, int32_t Cout, int32_t Cin, int32_t accumulate,
                                      void* stream) {
  CCDM_REQUIRE(packed && dw && Cout > 0 && Cin > 0 && 14 * Cin <= 64, CCDM_ERR_BAD_ARG, "stem_unpack_wgrad: bad args");
  const int total = Cout * Cin * 49;
  stem_unpack_wgrad_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed, dw, Cout, Cin, accumulate);
  return after_launch("stem_unpack_wgrad_kernel");
}

extern "C" int ccdm_gather_augment_u8(const uint8_t* images, int64_t n_images, const int64_t* idx, const uint8_t* aug,
                                      float* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  CCDM_REQUIRE(images && idx && out && n_images > 0 && B > 0 && C > 0 && H > 0 && W > 0, CCDM_ERR_BAD_ARG,
               "gather_augment_u8: bad args");
  const long long total = (long long)B * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_augment_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(images, (const long long*)idx, aug, out, B, C,
                                                                               H, W, (long long)n_images);
  return after_launch("gather_augment_u8_kernel");
}
