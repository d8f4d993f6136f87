/* ccdm_b200 -- C ABI of the B200-native CCDM denoiser hot path (libccdm_b200.so).
 *
 * The reference (eric98040/CCDM, CCDM_unified/) is pure PyTorch and has no FFI of its own; every entry point
 * below names the reference call site(s) it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it is an args struct;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it, no host sync, no allocation
 *     that outlives the call -> every entry point is CUDA-graph capturable;
 *   - return 0 on success, a negative CCDM_ERR_* otherwise; ccdm_last_error() gives a thread-local message;
 *   - activations between kernels are bf16 NHWC ("pixel rows" of C channels); NCHW fp32 only at the stem / head.
 * There is no CPU fallback: without a CUDA device every compute entry point returns CCDM_ERR_CUDA.
 */
#ifndef CCDM_B200_H
#define CCDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCDM_OK 0
#define CCDM_ERR_BAD_ARG (-1)
#define CCDM_ERR_UNSUPPORTED_SHAPE (-2)
#define CCDM_ERR_CUDA (-3)

int ccdm_version(void);
const char* ccdm_last_error(void);
/* Number of kernels this library has launched on this process so far (bench.py's "gpu_launches"). */
int64_t ccdm_launch_count(void);
/* sizeof() of the args structs as this library was compiled: 0 tapgemm, 1 view, 2 step, 3 qsample, 4 loss, 5 wgrad.
 * Lets a foreign-language binding verify its struct mirrors. */
int ccdm_struct_size(int which);

/* ------------------------------------------------------------------------------------------------------------
 * Tap-GEMM: the implicit-GEMM convolution / linear engine (tcgen05 + TMEM + TMA).
 *
 *   out[b,h,w,n] = epilogue( sum_{g,r}  A_{src[g]}[b, h+dh0[g]+r, w+dw[g], c0[g] : c0[g]+64] . Wp[n, (g*R+r)*64 : +64] )
 *
 * Replaces nn.Conv2d (3x3, 1x1, 4x4/s2 Downsample, nearest-2x Upsample+3x3) and nn.Linear call sites of
 * CCDM_unified/models/unet.py:77,81,139,160,165,195,198,225,226,326,341 together with the elementwise tail that
 * follows them there: RMSNorm (:88-89), (scale+1)*x+shift (:147-149), SiLU (:151), residual adds (:72,:187), the
 * PreNorm in front of the attention 1x1 (:97-99, folded as a per-row scale) and the q softmax (:207,:210).
 * ------------------------------------------------------------------------------------------------------------ */
#define CCDM_MAX_SRC 4
#define CCDM_MAX_Z 4

#define CCDM_EPI_BIAS 0x1u       /* + bias[n] */
#define CCDM_EPI_ROWSCALE 0x2u   /* acc *= 1/max(sqrt(rowss[pixel]),1e-12) before the bias (PreNorm fold) */
#define CCDM_EPI_RMSNORM 0x4u    /* v *= gain[n]*gain_mul / max(||v||_2 over n, 1e-12); needs N <= n_tile <= 512 */
#define CCDM_EPI_SS 0x8u         /* v = v*(1+scale[b,n]) + shift[b,n]; without CCDM_EPI_RMSNORM it needs tb == 1 */
#define CCDM_EPI_SILU 0x10u
#define CCDM_EPI_RESID 0x20u     /* v += resid[b,h,w,n] (bf16) */
#define CCDM_EPI_QSOFTMAX 0x40u  /* columns < q_cols: softmax over aligned groups of 32, times q_scale */
#define CCDM_EPI_SUMSQ_OUT 0x80u /* out_rowss[pixel] = sum_n (bf16-rounded out)^2 */
#define CCDM_EPI_OUT_F32 0x100u  /* out is fp32 instead of bf16 */
#define CCDM_EPI_KEXP 0x200u     /* columns in [q_cols, 2*q_cols): v = exp(v) (the k softmax numerator; bias carries
                                    the -bound shift from ccdm_kexp_bound, so v <= ~1 and nothing overflows) */
#define CCDM_EPI_RELU 0x400u     /* v = max(v, 0)   (generator blocks, models/sngan.py:76-80) */
#define CCDM_EPI_TANH 0x800u     /* v = tanh(v)     (generator output, models/sngan.py:128) */
#define CCDM_EPI_RESACC 0x2000u  /* shortcut in the SAME launch (ResnetBlock: h + res_conv(x), unet.py:165,187): n_res extra load
                                    groups -- unshifted 64-channel boxes of the sources named by sched entries
                                    [nz*ngroups, +n_res) -- are multiplied with the last n_res K blocks of wpacked into a
                                    second TMEM accumulator, and v += acc2[n] + res_bias[n] after the tail (identity
                                    shortcut = identity K block).  nz == 1, n_tile <= 128, shared weights, no CCDM_EPI_RESID */
#define CCDM_EPI_HEAD 0x1000u    /* fused 1x1 head (unet.py:348,455 final_conv after final_res_block): instead of storing the
                                    tile, head_out[b][k][h][w] = sum_n v[n]*head_w[k][n] + head_b[k] for k < head_n (fp32
                                    NCHW); needs one N tile (n_rows == n_tile <= 128), nz == 1; `out` may be NULL */

typedef struct ccdm_view {
  const void* ptr;    /* bf16; first element of the view (already offset for channel slices / parity planes) */
  int32_t C, W, H, B; /* extents seen by TMA; reads outside them return 0 (this is the conv zero padding) */
  int64_t sW, sH, sB; /* strides in elements; the channel stride is 1 */
} ccdm_view;

typedef struct ccdm_tapgemm_args {
  int32_t n_src;
  ccdm_view src[CCDM_MAX_SRC];
  int32_t gW, gH, gB; /* extents of the output-position grid the 128-row tiles walk over */
  int32_t tw, th, tb; /* tile box, tw*th*tb == 128 */
  int32_t nz;         /* sub-problems (output parity planes of the nearest-2x conv) */
  int32_t ngroups, R; /* K loop: ngroups TMA boxes per sub-problem, each th+R-1 rows tall and feeding R vertically
                         adjacent filter taps (tap r starts r*tw rows into the box); R > 1 needs tb == 1, tw % 8 == 0 */
  const int32_t* sched; /* device [nz*ngroups][4] = {src, dw, dh0, c0} */
  const void* wpacked;  /* device bf16 [nz*n_rows][ngroups*R*64], K contiguous, block (g*R + r) = group g, tap r */
  int32_t n_rows;       /* packed rows per sub-problem; multiple of n_tile */
  int32_t w_batch_rows; /* 0: one weight set.  >0: per-sample weights, sample b starts at row b*w_batch_rows
                           (linear-attention output projection with the context folded in); needs tb == 1 */
  int32_t N;            /* valid output channels */
  int32_t n_tile;       /* output channels per CTA: multiple of 32, <= 512 */
  uint32_t flags;       /* CCDM_EPI_* */
  const float* bias;
  const float* rowss;
  const float* gain;    /* RMSNorm g[n]; the kernel multiplies it by gain_mul (= sqrt(C), unet.py:89) */
  float gain_mul;
  const float* scale_shift; /* scale at [b*ss_ld + ss_off + n], shift at [b*ss_ld + ss_off + N + n] */
  int32_t ss_ld, ss_off;
  const void* resid;
  int64_t rsW, rsH, rsB;
  void* out;
  int64_t osW, osH, osB;
  int64_t ooff[CCDM_MAX_Z]; /* element offset of sub-problem z inside out */
  float* out_rowss;
  float q_scale;
  int32_t q_cols;
  /* CCDM_EPI_HEAD: head_w fp32 [head_n][N], head_b fp32 [head_n], head_out fp32 with plane stride hsC and sample stride
     hsB (elements; pixel (h, w) at h*gW + w inside a plane), 1 <= head_n <= 4 */
  int32_t halo;          /* 1: 3x3 (R == 9): every load group is ONE box {64 ch, tw+2, th+2} at (dw, dh0) = (-1, -1) feeding all nine
                            taps (K block g*9 + r*3 + q = tap (r, q)); needs tile 8 x th x 1.  A third of the fills of R == 3 */
  int32_t n_res;         /* CCDM_EPI_RESACC: shortcut load groups; wpacked then has ngroups*R + n_res K blocks per row */
  const float* res_bias; /* fp32 [N] or NULL */
  int32_t head_n;
  const float* head_w;
  const float* head_b;
  float* head_out;
  int64_t hsC, hsB;
  /* device int32 [nz*ngroups] or NULL: K steps (16 input channels each, 1..4) of each load group's 64-channel block that
     hold data; the rest are neither issued nor read from wpacked (a 72-channel source is a full block plus a block with ONE
     live K step).  NULL = 4 everywhere.  Ignored with halo boxes. */
  const int32_t* ksteps;
} ccdm_tapgemm_args;

int ccdm_tapgemm(const ccdm_tapgemm_args* args, void* stream);

/* Pack fp32 conv / linear weights [Cout][Cin_total][kh*kw] into the bf16 K-blocked layout ccdm_tapgemm reads.
 * psched: device [nz*nkb][4] = {cin0, nvalid, tapmask, 0}: block kb of sub-problem z holds, for j < nvalid,
 *   sum over taps t in tapmask of W[n][cin0+j][t] * (cin_gain ? cin_gain[cin0+j] : 1) * gain_mul     and 0 for j >= nvalid
 * (tap sums implement the nearest-2x upsample fold; cin_gain folds the PreNorm g of unet.py:97-99). */
int ccdm_pack_weights(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                      int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain, float gain_mul, void* wpacked,
                      void* stream);
/* Multi-tensor form of ccdm_pack_weights (mode 0) / ccdm_pack_weights_t (mode 1) for the training step, where every
 * convolution weight is re-packed once per optimizer step (autograd of the nn.Conv2d sites reached from trainer.py:724): one
 * launch over a DEVICE-resident job table.  total = nz * n_rows * nkb * 64 packed elements. */
typedef struct ccdm_pack_job {
  const float* w;
  void* out;
  const int32_t* psched;
  int32_t mode, cout, cin_total, ntaps, nkb, n_rows, n_off, n_count;
  int64_t total;
} ccdm_pack_job;
int ccdm_pack_multi(const ccdm_pack_job* jobs, int32_t njobs, void* stream);
/* The same, into K blocks [kb0, kb0 + nkb) of a packed matrix with nkb_total K blocks per row (a conv and its block's shortcut
 * share one matrix, CCDM_EPI_RESACC). */
int ccdm_pack_weights_at(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                         int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain, float gain_mul, void* wpacked,
                         int32_t nkb_total, int32_t kb0, void* stream);

/* Standalone channel RMSNorm + tail for rows of C contiguous bf16 channels (unet.py:88-89,145-151), used when a
 * layer's channels do not fit one tap-GEMM tile (C > 512) or its GEMM is split over output channels:
 *   out = [silu]( z/max(|z|,1e-12) * gain*gain_mul * (1+scale[b]) + shift[b] ) [+ resid];  flags: CCDM_EPI_SS | SILU |
 *   RESID | SUMSQ_OUT.  z already contains the conv bias.  Sample of a row = row / rows_per_sample. */
int ccdm_rmsnorm_act(const void* z, void* out, int64_t rows, int32_t C, int32_t rows_per_sample, const float* gain,
                     float gain_mul, const float* scale_shift, int32_t ss_ld, int32_t ss_off, const void* resid,
                     float* out_rowss, uint32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stem and head (NCHW fp32 <-> NHWC bf16 boundary).
 *   stem: unet.py:271,418  nn.Conv2d(in_channels, dim, 7, padding=3)  -> im2row + ccdm_tapgemm (below)
 *   head: unet.py:348,455  nn.Conv2d(dim, out_dim, 1)
 * ------------------------------------------------------------------------------------------------------------ */
/* Tensor-core stem: the 7x7 conv as a 4-tap ccdm_tapgemm (schedule {src 0, dw 0, dh0 2g-2, c0 0}, g = 0..3) over
 *   rowimg[b,j,w, dr*7*Cin + s*Cin + c] = x[b,c,j-1+dr,w+s-3], j = 0..H   (bf16 [B][H+1][W][64], zero outside the image)
 * with weights packed[n][g*64 + dr*7*Cin + s*Cin + c] = w[n][c][2g+dr][s]. */
int ccdm_stem_im2row(const float* x_nchw, void* rowimg, int32_t B, int32_t Cin, int32_t H, int32_t W, void* stream);
int ccdm_stem_pack(const float* w, void* wpacked, int32_t Cout, int32_t Cin, int32_t n_rows, void* stream);
int ccdm_head_conv1(const void* x_nhwc, const float* w, const float* bias, float* out_nchw, int32_t B, int32_t H,
                    int32_t W, int32_t Cin, int32_t Cout, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention cores.
 *   linear attention, unet.py:202-216: qkv is bf16 [B][n][3*heads*32]; the tap-GEMM epilogue has already turned
 *     q into softmax_d(q)*scale and k into p = exp(k - bound_d); context[b,h,d,e] = sum_n p[d,n] v[e,n] / sum_n p[d,n]
 *     (fp32 [B][heads][32][32]) runs on tcgen05 for heads == 4 (both operands token-major = MN-major UMMA)
 *   softmax attention at the bottleneck, unet.py:228-240 (n <= 64 tokens, dim_head <= 64)
 * ------------------------------------------------------------------------------------------------------------ */
/* When wfold != NULL the fold below is fused into the same kernel (ctx may then be NULL).  colsum (optional, heads == 4
 * only) receives S[b][h*32+d] = sum_n p[d,n], which the backward needs. */
int ccdm_linattn_context(const void* qkv, float* ctx, float* colsum, int32_t B, int32_t n, int32_t heads,
                         const float* w_out, void* wfold, int32_t C, int32_t n_rows, void* stream);
/* bias[n] = -1.01*||wpacked[n,:]||_2 - 1e-3 for n in [row_lo,row_hi), 0 elsewhere.  The PreNorm'd input rows have unit
 * length, so |k[n,d]| <= ||W'_d||: using this bound as the softmax shift needs no max pass over the tokens. */
int ccdm_kexp_bound(const void* wpacked, int32_t n_rows, int32_t K, int32_t row_lo, int32_t row_hi, float* bias,
                    void* stream);
/* Fold the per-sample context into the output projection (unet.py:214-216 + to_out[0] at :198):
 *   wfold[b][c][h*32+d] = sum_e w_out[c][h*32+e] * ctx[b][h][d][e]     bf16 [B][n_rows][heads*32], rows >= C zero
 * so that to_out(context^T . q) becomes one tap-GEMM over q with per-sample weights (w_batch_rows = n_rows). */
int ccdm_linattn_fold(const float* w_out, const float* ctx, void* wfold, int32_t B, int32_t C, int32_t n_rows,
                      int32_t heads, void* stream);
int ccdm_attention_small(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t dim_head,
                         float scale, void* stream);

/* Fused linear attention for inference (heads == 4, dim_head == 32, C <= 128 channels, n a multiple of 128 tokens):
 * Residual(PreNorm(LinearAttention)) of unet.py:66-72,92-99,202-216 without q | k | v ever reaching HBM.
 *   x      bf16 [B][n][C] token rows (NHWC), rowss fp32 [B*n] their sums of squares (the producing conv's
 *          CCDM_EPI_SUMSQ_OUT), wqkv the packed to_qkv weight [384][ceil(C/64)*64] with the PreNorm gain folded in
 *          (ccdm_pack_weights, cin_gain), kbias [384] the k rows' softmax shifts (ccdm_kexp_bound).
 * ccdm_linattn_kv_partials: per UNIT (ccdm_linattn_fused_units(n) per sample, a fixed run of 128-token tiles) the
 *   un-normalised context  part[b][u][h*32+d][e] = sum_{n in unit} p[n,d] v[n,e],  psum[b][u][h*32+d] = sum p[n,d]
 *   with p = exp(k/|x| - bound), v = v/|x|   (bf16-rounded, fp32 accumulation in TMEM).
 * ccdm_linattn_fold_partials: sums the units in a fixed order, divides by psum (softmax over tokens, unet.py:208) and
 *   folds the context into to_out[0]:  wfold[b][c][h*32+d] = sum_e w_out[c][h*32+e] ctx[b][h][d][e]  (bf16, rows >= C
 *   untouched: keep them zero).
 * ccdm_linattn_q_out: out = RMSNorm_g( softmax_d(q)*q_scale . wfold_b^T + bias ) * gain_mul + x, bf16 [B][n][C].
 * Results are independent of the batch size (no atomics, fixed summation order). */
int ccdm_linattn_fused_units(int32_t n);
int ccdm_linattn_kv_partials(const void* x, int32_t B, int32_t n, int32_t C, const float* rowss, const void* wqkv,
                             const float* kbias, float* part, float* psum, void* stream);
int ccdm_linattn_fold_partials(const float* part, const float* psum, int32_t B, int32_t units_per_sample,
                               const float* w_out, int32_t C, int32_t n_rows, void* wfold, void* stream);
int ccdm_linattn_q_out(const void* x, int32_t B, int32_t n, int32_t C, const float* rowss, const void* wqkv,
                       const void* wfold, int32_t n_rows, const float* bias, const float* gain, float gain_mul,
                       float q_scale, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Embedding MLPs (unet.py:102-115 sinusoid, :289-312 time / label MLPs with BatchNorm1d, :397-414 null-label
 * select, :158-161 SiLU in front of every tc_mlp).  fp32, batch-sized.
 * ------------------------------------------------------------------------------------------------------------ */
#define CCDM_ACT_NONE 0
#define CCDM_ACT_RELU 1
#define CCDM_ACT_GELU 2
#define CCDM_ACT_SILU 3
/* y[b, :] = act( BN( x[b, :] . W^T + bias ) ); bn_* may be NULL.  In training mode (bn_train != 0) the batch
 * statistics are used and running_mean / running_var are updated in place (momentum 0.1, unbiased variance). */
int ccdm_linear_small(const float* x, int32_t B, int32_t in_dim, const float* w, const float* bias, int32_t out_dim,
                      const float* bn_w, const float* bn_b, float* bn_mean, float* bn_var, int32_t bn_train,
                      int32_t act, float* y, int64_t y_ld, void* stream);
/* x[b, :] = act( GroupNorm(groups, C)(x[b, :]) * gamma + beta ), in place, fp32 [B][C] (nn.GroupNorm on a 2-D input:
 * statistics over the C/groups channels of each group of each row; biased variance, eps inside the sqrt).  With
 * ccdm_linear_small this is the forward of the learned label-embedding MLPs model_y2h / model_y2cov
 * (models/resnet_y2h.py:143-173, models/resnet_y2cov.py:149-179; called from label_embedding.py:1028-1031,1173-1176). */
int ccdm_groupnorm_rows(float* x, int32_t B, int32_t C, int32_t groups, const float* gamma, const float* beta, float eps,
                        int32_t act, void* stream);
int ccdm_time_features(const int64_t* t, int32_t B, int32_t dim, float* out, void* stream);
/* c[b,:] = keep[b] ? c[b,:] : null_emb[:]   (keep may be NULL with all_null != 0) */
int ccdm_select_null(float* c, const uint8_t* keep, int32_t all_null, const float* null_emb, int32_t B, int32_t dim,
                     void* stream);
/* out_bf16[b, 0:dt] = silu(t_emb[b]),  out_bf16[b, dt:dt+dc] = silu(c_emb[b]) */
int ccdm_silu_concat_bf16(const float* t_emb, int32_t dt, const float* c_emb, int32_t dc, int32_t B, void* out,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Guidance + sampler step (one kernel).
 *   CFG combine: unet.py:51-62,365-380 (orthogonal update, std rescale)
 *   model_predictions: diffusion.py:295-336;  DDIM update :454-464;  DDPM p_sample :338-374, q_posterior :284-293
 * coef is a device table [n_steps][CCDM_STEP_NCOEF] indexed by *step_counter (device int32), which the kernel
 * increments when `advance` != 0, so one captured graph replays for every step.
 * ------------------------------------------------------------------------------------------------------------ */
#define CCDM_OBJ_PRED_NOISE 0
#define CCDM_OBJ_PRED_X0 1
#define CCDM_OBJ_PRED_V 2
#define CCDM_STEP_NCOEF 12
/* coef row: [0] sqrt_recip_acp  [1] sqrt_recipm1_acp  [2] sqrt_acp  [3] sqrt_1m_acp
 *           [4] ddim sqrt(alpha_next)  [5] ddim c  [6] ddim sigma  [7] is_last (x = x0)
 *           [8] ddpm posterior_mean_coef1  [9] coef2  [10] exp(0.5*posterior_log_variance) (0 at t==0)  [11] unused */
typedef struct ccdm_step_args {
  const float* out_cond; /* [B][C][H][W] fp32 network output, conditional half */
  const float* out_null; /* unconditional half; NULL when cond_scale == 1 */
  float* x;              /* in: x_t, out: x_{t-1}  (fp32 NCHW) */
  const float* noise;    /* N(0,1) draw for this step (may be NULL when every sigma is 0) */
  float* pred_noise;     /* optional outputs for parity traces (may be NULL) */
  float* pred_x0;
  int32_t B, chw;
  float cond_scale, rescaled_phi, keep_parallel_frac;
  int32_t remove_parallel, objective, clip_x0, cfg_plus_plus;
  int32_t sampler;       /* 0 = DDIM update, 1 = DDPM update, 2 = predictions only (x is read, not written) */
  const float* coef;
  int32_t* step_counter; /* coef row = *step_counter (all samples), unless t_rows is given */
  int32_t advance;
  const int64_t* t_rows; /* optional [B]: per-sample coef row (model_predictions with arbitrary timesteps) */
} ccdm_step_args;
int ccdm_sampler_step(const ccdm_step_args* args, void* stream);
/* out[i] = table[*step_counter] for i < n  -- refreshes the timestep input of a captured step graph
 * (torch.full((batch,), time), diffusion.py:440 / :363). */
int ccdm_broadcast_step_i64(const int64_t* table, const int32_t* step_counter, int64_t* out, int32_t n, void* stream);
/* Guidance arithmetic alone (forward_with_cond_scale without the sampler): guided = f(cond, null). */
int ccdm_cfg_combine(const float* cond, const float* null_out, float* guided, int32_t B, int32_t chw,
                     float cond_scale, float rescaled_phi, int32_t remove_parallel, float keep_parallel_frac,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training-side elementwise kernels.
 *   q_sample: diffusion.py:487-499 fused with img*2-1 (:755) and the Hy noise scaling (:550-557)
 *   loss:     diffusion.py:570-594,597-730 (target select, MSE, /Hy, loss_weight[t], vicinal batch weights)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ccdm_qsample_args {
  const float* img01;    /* [B][chw] images; in [0,1] when normalize != 0, already in [-1,1] otherwise */
  const float* noise;    /* [B][chw] N(0,1) */
  const float* noise2;   /* [B][chw] second draw used for null rows under use_Hy (may be NULL) */
  const float* cov;      /* [B][chw] exp(-y2cov) or NULL */
  const uint8_t* keep;   /* [B] keep mask (1 = conditional row) */
  const int64_t* t;      /* [B] */
  const float* sqrt_acp; /* [T] */
  const float* sqrt_1m_acp;
  float* x0;             /* out: img*2-1 (normalize != 0) or a copy of img */
  float* noise_out;      /* out: the noise actually mixed in (scaled by sqrt(cov) on conditional rows) */
  float* x_t;            /* out */
  int32_t B, chw;
  int32_t normalize;
} ccdm_qsample_args;
int ccdm_q_sample(const ccdm_qsample_args* args, void* stream);

typedef struct ccdm_loss_args {
  const float* model_out; /* [B][chw] */
  const float* x0;
  const float* noise;
  const float* cov;       /* NULL unless use_Hy */
  const uint8_t* keep;
  const int64_t* t;
  const float* sqrt_acp;
  const float* sqrt_1m_acp;
  const float* loss_weight; /* [T] */
  const float* row_weight;  /* [B] vicinal batch weights (already 1 on null rows), or NULL for the plain mean */
  float* per_sample;        /* out [B]: loss_weight[t] * sum_chw (out-target)^2 / cov */
  float* loss;              /* out [1] */
  float* grad_out;          /* out [B][chw]: d loss / d model_out (may be NULL) */
  int32_t B, chw, objective;
} ccdm_loss_args;
int ccdm_vicinal_loss(const ccdm_loss_args* args, void* stream);
/* In-batch vicinal weights, diffusion.py:669-727 ("traditional" branch) and :602-664 (sliced branch):
 * proj is [B][P] (P = 1 and proj = labels for scalar labels; P = label_dim with `euclid` != 0 for the multi-
 * dimensional l2 branch; P = num_projections of already-projected labels for the sliced branch, where the
 * per-projection thresholds are thr[p]).  w[i] = (1/B) * (1/P or 1) * sum_j [ |d_ij| <= thr ]  or exp(-nu d_ij^2). */
int ccdm_vicinal_weights(const float* proj, int32_t B, int32_t P, int32_t euclid, int32_t hard, const float* thr,
                         float nu, const uint8_t* keep, float* w, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward building blocks of the training step (autograd of the call sites above; trainer.py:560-640 drives
 * loss.backward() through them in the reference).  The data gradient of every convolution is ccdm_tapgemm itself,
 * run over dY with weights packed by ccdm_pack_weights_t (filter flipped / channel roles swapped; the stride-2 and
 * nearest-2x convolutions trade places: see ccdm_b200/plan.py "*_dgrad").
 * ------------------------------------------------------------------------------------------------------------ */
/* Transposed packing: block kb of sub-problem z holds, for j < nvalid and n < n_count,
 *   sum over taps t in tapmask of W[cin0+j][n_off+n][t],   W = the forward [cout][cin_total][ntaps] tensor
 * (n_off / n_count select the input-channel window of one source of a concatenated forward input). */
int ccdm_pack_weights_t(const float* w, int32_t cout, int32_t cin_total, int32_t ntaps, const int32_t* psched,
                        int32_t nz, int32_t nkb, int32_t n_rows, int32_t n_off, int32_t n_count, void* wpacked,
                        void* stream);

/* Weight gradient of one tap-GEMM layer (tcgen05, positions as the K axis), accumulated (+=, fp32) in the PACKED
 * layout of the forward weights:  wgrad_packed[z][n][(g*R+r)*64 + j] += sum_p dZ_z[p][n] * src[g][p + tap(g,r)][c0[g]+j].
 * src / gW..tb / nz / ngroups / R / sched are the forward layer's; dz is the gradient w.r.t. the conv output (bias
 * included, before any norm), bf16, addressed like the forward `out` (strides dsW/dsH/dsB, plane offsets doff[z]).
 * ksplit = number of position slices (CTAs per (z, group, 128-row tile)); 0 = fill the GPU once.
 * R == 9 selects the HALO plan of a 3x3 / stride-1 layer (tile 8 x 16 x 1, nz == 1): a group is one 64-channel block
 * {source view, -1, -1, c0} and K block g*9 + r*3 + q is filter tap (r, q); one dZ box and one X box {64, 10, 18} per
 * tile feed all nine taps. */
typedef struct ccdm_wgrad_args {
  int32_t n_src;
  ccdm_view src[CCDM_MAX_SRC];
  const void* dz;
  int64_t dsW, dsH, dsB;
  int64_t doff[CCDM_MAX_Z];
  int32_t gW, gH, gB;
  int32_t tw, th, tb;
  int32_t nz, ngroups, R;
  const int32_t* sched;
  int32_t N, n_rows;
  float* wgrad_packed;
  int32_t ksplit;
  /* slots > 0: PARTIAL mode.  Exactly `slots` position slices are launched (ksplit is ignored) and slice s STORES its
   * partial gradient (no atomics, no pre-zeroed buffer; a slice without positions stores zeros) at
   * wgrad_packed + s*slot_stride; ccdm_unpack_wgrad_slots sums the slices in a fixed order (bit-reproducible).
   * slots == 0: every slice adds into wgrad_packed with red.global.add (buffer zeroed or accumulated by the caller). */
  int32_t slots;
  int64_t slot_stride;
} ccdm_wgrad_args;
int ccdm_conv_wgrad(const ccdm_wgrad_args* args, void* stream);
/* Packed gradient -> dW[cout][cin_total][ntaps] (= or +=): every (ci, t) gathers the packed blocks that hold it
 * (the folded taps of the nearest-2x convolution sum), times cin_gain[ci]*gain_mul when the forward packed with them. */
int ccdm_unpack_wgrad(const float* packed, float* dw, int32_t cout, int32_t cin_total, int32_t ntaps,
                      const int32_t* psched, int32_t nz, int32_t nkb, int32_t n_rows, const float* cin_gain,
                      float gain_mul, int32_t accumulate, void* stream);
/* Same, over `nslots` partial gradients `slot_stride` elements apart (ccdm_wgrad_args.slots), summed in slot order. */
int ccdm_unpack_wgrad_slots(const float* packed, int32_t nslots, int64_t slot_stride, float* dw, int32_t cout,
                            int32_t cin_total, int32_t ntaps, const int32_t* psched, int32_t nz, int32_t nkb,
                            int32_t n_rows, const float* cin_gain, float gain_mul, int32_t accumulate, void* stream);

/* Backward of the Block tail (unet.py:88-89,145-151): with zh = z/max(|z|,1e-12), n = zh*gain*gain_mul,
 * u = n*(1+scale[b]) + shift[b], y = silu(u):
 *   du = dy * silu'(u)          dz = (gain*gain_mul*(1+scale)*du - zh * <zh, same>) / |z|        (bf16 out)
 *   sums[0][b][c] += sum_rows du*zh      sums[1][b][c] += sum_rows du      sums[2][b][c] += sum_rows dz
 * flags: CCDM_EPI_SS | CCDM_EPI_SILU as in the forward.  sums is fp32 [3][B][C], zeroed by the caller. */
int ccdm_block_bwd(const void* dy, const void* z, void* dz, int64_t rows, int32_t C, int32_t rows_per_sample,
                   const float* gain, float gain_mul, const float* scale_shift, int32_t ss_ld, int32_t ss_off,
                   float* sums, uint32_t flags, void* stream);
/* Finish: d_ss[b][ss_off+c] = gain*gain_mul*sums0, d_ss[b][ss_off+C+c] = sums1 (when d_ss != NULL);
 * dgain[c] += gain_mul * sum_b (1+scale[b,c]) * sums0;  dbias[c] += sum_b sums2. */
int ccdm_block_bwd_finish(const float* sums, int32_t B, int32_t C, const float* gain, float gain_mul,
                          const float* scale_shift, int32_t ss_ld, int32_t ss_off, float* d_ss, float* dgain,
                          float* dbias, void* stream);

/* ---- training-only pieces (autograd of unet.py:202-216, :228-240, :271, :348); heads = 4, dim_head = 32 for the
 * linear attention, as in the reference (never overridden, unet.py:190,325,340) ---- */
/* out[c] += sum over rows of x[row][c] (bf16 [rows][C]): bias gradients of plain convolutions. */
int ccdm_colsum_bf16(const void* x, int64_t rows, int32_t C, float* out, void* stream);
/* In place on the raw qkv [B][n][384] of the to_qkv conv: q <- softmax over each head's channels * scale (:207,:210),
 * k <- exp(k - kmax[b][c]) with kmax the per-sample maximum over tokens (written to kmax [B][128]). */
int ccdm_linattn_prep(void* qkv, int32_t B, int32_t n, float* kmax, float scale, void* stream);
/* Per-sample block-diagonal [B][128][128] bf16 weights for ccdm_tapgemm (w_batch_rows = 128) from m = fp32
 * [B][4][32][32]: (transpose & 1) == 0 -> w[b][h*32+e][h*32+d] = m[b][h][d][e] / row_div[b][h*32+d]; else rows/cols
 * swapped.  transpose & 2: write only the four diagonal 32x32 blocks (the caller zeroed the buffer once). */
int ccdm_linattn_pack_blockdiag(const float* m, const float* row_div, int32_t transpose, void* w, int32_t B, void* stream);
/* dctx[b][h][d][e] = sum_n q_sm[b][n][h*32+d] * dout[b][n][h*32+e]  (tcgen05, same kernel as the context). */
int ccdm_linattn_dcontext(const void* qkv, const void* dout, float* dctx, int32_t B, int32_t n, void* stream);
/* c[b][h*32+d] = sum_e dctx*ctx / S  (the column-sum term of the token softmax backward). */
int ccdm_linattn_bwd_rowdot(const float* ctx, const float* dctx, const float* S, float* c, int32_t B, void* stream);
/* In place: dpre [B][n][384] = [dq_sm | dp_term | dv] -> gradient of the raw qkv, given qkv = [q_sm | p | v]. */
int ccdm_linattn_bwd_finish(const void* qkv, void* dpre, int32_t B, int32_t n, const float* c, float scale, void* stream);
/* Backward of ccdm_attention_small: dqkv [B][n][3*heads*dim_head] from the raw qkv and dout [B][n][heads*dim_head]. */
int ccdm_attention_small_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t n, int32_t heads,
                             int32_t dim_head, float scale, void* stream);
/* Backward of ccdm_head_conv1: dh (bf16 NHWC), dw [Cout][Cin] += , db [Cout] += from dout (fp32 NCHW). */
int ccdm_head_conv1_bwd(const float* dout_nchw, const void* h, const float* w, void* dh, float* dw, float* db, int32_t B,
                        int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);
/* Inverse of ccdm_stem_pack for the gradient ccdm_conv_wgrad leaves over the im2row tensor: dW [Cout][Cin][7][7]. */
int ccdm_stem_unpack_wgrad(const float* packed, float* dw, int32_t Cout, int32_t Cin, int32_t accumulate, void* stream);

/* Device-side batch construction (trainer.py:461-482 process_images + utils.py:164-211): out[b] = augment(images[idx[b]]) / 255,
 * fp32 NCHW in [0, 1] from a device-resident uint8 dataset [n_images][C][H][W].  aug[b] (may be NULL = none): bits 0-1 =
 * quarter turns k as np.rot90 over (H, W) (needs H == W), bit 2 = horizontal flip, bit 3 = vertical flip, applied in that
 * order (rotate, hflip, vflip) like the reference's Cell200 pipeline; UTKFace uses the hflip bit only. */
int ccdm_gather_augment_u8(const uint8_t* images, int64_t n_images, const int64_t* idx, const uint8_t* aug, float* out,
                           int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer step: clip_grad_norm_ + Adam of trainer.py:137,724,733-734 in three launches
 * over flat fp32 gradient / moment buffers, and the EMA lerp of ema_pytorch.py:150-178 in one.
 * ------------------------------------------------------------------------------------------------------------ */
/* chunk i covers chunk_n[i] elements of one parameter tensor starting at chunk_param[i]; its gradient and Adam
 * moments sit at g_flat / m_flat / v_flat + chunk_off[i].  step (device float) is incremented first (torch's
 * bias-correction convention); sumsq != NULL enables clipping: g *= min(1, max_norm / (||g_flat||_2 + 1e-6)), with the
 * squared norm left in sumsq[0] (device double).  n_flat = total elements of the flat buffers. */
int ccdm_fused_adam(void* const* chunk_param, const int64_t* chunk_off, const int32_t* chunk_n, int32_t n_chunks,
                    const float* g_flat, float* m_flat, float* v_flat, int64_t n_flat, float* step, double* sumsq,
                    float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* dst[i] += weight[0] * (src[i] - dst[i]) over n_chunks (pointer, pointer, count) triples. */
int ccdm_multi_lerp(void* const* dst, const void* const* src, const int32_t* chunk_n, int32_t n_chunks,
                    const float* weight, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * One-step generator: models/sngan.py:19-139 in eval mode.  The convolutions are
 * ccdm_tapgemm ("up2x3x3" for Upsample+conv1, "up2x1x1" for the bypass, epilogues CCDM_EPI_SS | CCDM_EPI_RELU for
 * ConditionalBatchNorm2d + ReLU and CCDM_EPI_TANH for the output); the two helpers below supply the rest.
 * ------------------------------------------------------------------------------------------------------------ */
/* (Conditional) BatchNorm2d with running statistics as a per-(sample, channel) affine map in the scale/shift layout
 * of ccdm_tapgemm (sngan.py:28-36): with r = rsqrt(var+eps), a = r * (weight ? weight[c] : 1) * (1 + (gamma ? gamma[b,c] : 0)):
 *   ss[b][c] = a - 1,   ss[b][C+c] = (beta ? beta[b,c] : 0) + (bias ? bias[c] : 0) - mean[c]*a        (ss is [B][2C]) */
int ccdm_condbn_coef(const float* gamma, const float* beta, const float* weight, const float* bias, const float* mean,
                     const float* var, float eps, int32_t B, int32_t C, float* ss, void* stream);
/* out = act(x*(1+scale[b]) + shift[b]) over bf16 rows of C channels (pre-activation CondBN + ReLU in front of a conv);
 * act: 0 none, 1 ReLU, 2 SiLU (the vanilla UNet's GroupNorm -> SiLU pre-activation). */
int ccdm_affine_act(const void* x, void* out, int64_t rows, int32_t C, int32_t rows_per_sample, const float* scale_shift,
                    int32_t ss_ld, int32_t ss_off, int32_t act, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Vanilla (ADM-style) UNet, SURVEY.md section 8f rank 4: CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/models/unet.py.
 * Its convolutions are ccdm_tapgemm (3x3, 1x1 shortcut over the concatenated sources, "down3x3s2" for Downsample :193-198,
 * "up2x3x3" for Upsample :178-190); nn.GroupNorm (:88-90; call sites :99, :111 + :146, :160, :323) is
 * ccdm_channel_stats -> ccdm_groupnorm_coef -> ccdm_affine_act (act 2 = SiLU) in front of the conv.
 * ------------------------------------------------------------------------------------------------------------ */
/* sums[b][0][c_off+c] += sum_pixels x[b][p][c],  sums[b][1][c_off+c] += sum_pixels x^2;  x bf16 [B][rows_per_sample][C]
 * (contiguous, C % 8 == 0, C <= 2048), sums fp32 [B][2][ld].  zero_first != 0 clears the whole sums buffer first
 * (a memset node on the stream); a concatenated input calls this once per source with its channel offset. */
int ccdm_channel_stats(const void* x, int32_t B, int32_t rows_per_sample, int32_t C, float* sums, int32_t ld,
                       int32_t c_off, int32_t zero_first, void* stream);
/* GroupNorm(groups, Ctot)(x)*gamma + beta [then *(1+scale[b]) + shift[b], unet.py:146] as a per-(sample, channel)
 * affine map in the [a-1 | t] scale/shift layout of ccdm_affine_act / ccdm_tapgemm, from sums [B][2][Ctot]:
 *   mean_g, var_g over rows_per_sample * Ctot/groups elements (biased variance, eps inside the sqrt);
 *   a = rstd_g*gamma[c]*(1+scale),  t = (beta[c] - mean_g*rstd_g*gamma[c])*(1+scale) + shift
 * scale_shift (may be NULL): fp32 [B][ss_ld], scale at ss_off + c, shift at ss_off + Ctot + c.
 * coef: fp32 [B][2*Ctot], one [a-1 | t] segment per source: channels [0,C0) at 0, channels [C0,Ctot) at 2*C0. */
int ccdm_groupnorm_coef(const float* sums, int32_t B, int32_t Ctot, int32_t groups, int64_t rows_per_sample, float eps,
                        const float* gamma, const float* beta, const float* scale_shift, int32_t ss_ld, int32_t ss_off,
                        int32_t C0, float* coef, void* stream);
/* Backward of [GroupNorm ->] x*(1+s)+t -> act (autograd of unet.py:99,111,146 reached from loss.backward()), three steps:
 *  1. ccdm_norm_bwd_stats:   bsums[b][0][c_off+c] += sum_p du, bsums[b][1][c_off+c] += sum_p du*x with du = dy*act'(x*(1+s)+t);
 *                            dy, x bf16 [B][rows_per_sample][C]; coef = the forward [s | t] segment of this source
 *                            (fp32 [B][coef_ld] at coef_off); act 0 none / 1 ReLU / 2 SiLU; bsums fp32 [B][2][ld].
 *  2. ccdm_groupnorm_bwd_coef: from the forward sums and bsums (both [B][2][Ctot]): per source [A | Bc | Cc] segments of
 *                            bcoef fp32 [B][3*Ctot] (source 0 at 0, source 1 at 3*C0) such that dx = A*du + Bc*x + Cc;
 *                            dgamma[c] += , dbeta[c] += (may be NULL); d_ss [B][2*Ctot] = [dscale | dshift] (may be NULL).
 *  3. ccdm_norm_bwd_apply:   dx (bf16) = A*du + Bc*x + Cc per element, du recomputed as in step 1. */
int ccdm_norm_bwd_stats(const void* dy, const void* x, int32_t B, int32_t rows_per_sample, int32_t C, const float* coef,
                        int32_t coef_ld, int32_t coef_off, int32_t act, float* bsums, int32_t ld, int32_t c_off,
                        int32_t zero_first, void* stream);
int ccdm_groupnorm_bwd_coef(const float* sums, const float* bsums, int32_t B, int32_t Ctot, int32_t groups,
                            int64_t rows_per_sample, float eps, const float* gamma, const float* beta,
                            const float* scale_shift, int32_t ss_ld, int32_t ss_off, int32_t C0, float* bcoef,
                            float* dgamma, float* dbeta, float* d_ss, void* stream);
int ccdm_norm_bwd_apply(const void* dy, const void* x, void* dx, int64_t rows, int32_t C, int32_t rows_per_sample,
                        const float* coef, int32_t coef_ld, int32_t coef_off, const float* bcoef, int32_t bcoef_ld,
                        int32_t bcoef_off, int32_t act, void* stream);
/* timestep_embedding (unet.py:40-57): out fp32 [B][dim] = [cos(t*f_j) | sin(t*f_j)], f_j = max_period^(-j/(dim/2)); dim even. */
int ccdm_time_features_adm(const int64_t* t, int32_t B, int32_t dim, float max_period, float* out, void* stream);
/* AttentionBlock core (unet.py:165-175): out[b][t][h*dh+d] = softmax_s(q_t . k_s * scale) v_s over all n tokens;
 * qkv bf16 [B][n][3*heads*dh].  head_major != 0: the reference's split (head h owns channels [3*dh*h, 3*dh*(h+1)) as
 * q | k | v); head_major == 0: [q heads | k heads | v heads] (the unified UNet's split).  dim_head in {16,32,64,128}. */
int ccdm_attention_tokens(const void* qkv, void* out, int32_t B, int32_t n, int32_t heads, int32_t dim_head, float scale,
                          int32_t head_major, void* stream);

/* Backward of ccdm_attention_tokens: dqkv (same layout as qkv) from the raw qkv and dout [B][n][heads*dim_head].  One CTA
 * per (sample, head) with the head's q, k, v, dO in shared memory: n * dim_head <= ~12 k elements. */
int ccdm_attention_tokens_bwd(const void* qkv, const void* dout, void* dqkv, int32_t B, int32_t n, int32_t heads,
                              int32_t dim_head, float scale, int32_t head_major, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CCDM_B200_H */
