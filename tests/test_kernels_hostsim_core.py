"""The CUDA-core kernels of the PRIMARY path (sampler.cu, kernels.cu, backward.cu, train_kernels.cu, optim.cu), compiled
for the HOST from their own source (tests/hostsim) and checked against the oracle / torch / autograd on CPU.

These kernels are GPU-verified (tests/test_gpu_*.py); running their source in the `-m "not gpu"` tier as well means a change
to their index arithmetic, reductions or argument checks is caught in a container without a GPU.  tcgen05 / TMA kernels
(tapgemm.cu, linattn.cu, wgrad.cu) cannot be run this way.  Not a fallback: the product loads only libccdm_b200.so.
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

import oracle
from ccdm_b200 import _lib as L
from oracle import diffusion_ref as D
from oracle.unet_ref import cfg_combine
from tests.hostsim.build import build


def _load(cu, names):
    h = C.CDLL(build(cu))
    for n in names:
        fn = getattr(h, n)
        fn.restype, fn.argtypes = L.SIGNATURES[n]
    return h


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def samp():
    return _load("sampler.cu", ["ccdm_sampler_step", "ccdm_cfg_combine", "ccdm_q_sample", "ccdm_vicinal_loss",
                                "ccdm_vicinal_weights", "ccdm_broadcast_step_i64"])


@pytest.fixture(scope="module")
def kern():
    return _load("kernels.cu", ["ccdm_rmsnorm_act", "ccdm_affine_act", "ccdm_head_conv1", "ccdm_attention_small",
                                "ccdm_linear_small", "ccdm_time_features", "ccdm_select_null", "ccdm_silu_concat_bf16",
                                "ccdm_condbn_coef", "ccdm_groupnorm_rows"])


@pytest.fixture(scope="module")
def bwd():
    return _load("backward.cu", ["ccdm_block_bwd", "ccdm_block_bwd_finish"])


@pytest.fixture(scope="module")
def trk():
    return _load("train_kernels.cu", ["ccdm_colsum_bf16", "ccdm_attention_small_bwd", "ccdm_head_conv1_bwd"])


@pytest.fixture(scope="module")
def opt():
    return _load("optim.cu", ["ccdm_fused_adam", "ccdm_multi_lerp"])


# ------------------------------------------------------------------------------------------------- sampler.cu

@pytest.mark.parametrize("remove_parallel,phi", [(1, 0.7), (0, 0.7), (1, 0.0)])
def test_cfg_combine(samp, remove_parallel, phi):
    g = torch.Generator().manual_seed(1)
    cond, null = torch.randn(3, 3, 8, 8, generator=g), torch.randn(3, 3, 8, 8, generator=g)
    out = torch.empty_like(cond)
    assert samp.ccdm_cfg_combine(cond.data_ptr(), null.data_ptr(), out.data_ptr(), 3, 192, 1.5, phi, remove_parallel, 0.0, None) == 0
    if remove_parallel:
        want = cfg_combine(cond, null, 1.5, phi)
    else:                                                   # plain CFG (vanilla tree)
        want = null + (cond - null) * 1.5
        if phi:
            dims = (1, 2, 3)
            want = want * (cond.std(dim=dims, keepdim=True) / want.std(dim=dims, keepdim=True)) * phi + want * (1 - phi)
    assert rel(out, want) < 1e-5


@pytest.mark.parametrize("objective", ["pred_x0", "pred_noise", "pred_v"])
@pytest.mark.parametrize("kind", ["ddim", "ddpm"])
def test_sampler_step_matches_oracle(samp, objective, kind):
    """One fused guidance + prediction + update step == oracle.model_predictions + the reference's update formulas."""
    sch = oracle.make_schedule(1000, "cosine", objective)
    g = torch.Generator().manual_seed(2)
    B, shape = 2, (2, 3, 8, 8)
    chw = 192
    x, cond, null, noise = (torch.randn(shape, generator=g) for _ in range(4))
    t, tn, eta, scale, phi = 640, 420, 0.5, 1.5, 0.7
    net = lambda xx, tt, e, p: cond if p == 0.0 else null                                   # noqa: E731
    tt = torch.full((B,), t, dtype=torch.long)
    eps, x0 = D.model_predictions(sch, net, x, tt, None, scale, phi, clip_x_start=(kind == "ddim"))
    row = torch.zeros(1, L.STEP_NCOEF)
    row[0, 0], row[0, 1] = sch.sqrt_recip_alphas_cumprod[t], sch.sqrt_recipm1_alphas_cumprod[t]
    row[0, 2], row[0, 3] = sch.sqrt_alphas_cumprod[t], sch.sqrt_one_minus_alphas_cumprod[t]
    if kind == "ddim":
        a, an = sch.alphas_cumprod[t], sch.alphas_cumprod[tn]
        sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        row[0, 4], row[0, 5], row[0, 6] = an.sqrt(), c, sigma
        want = x0 * an.sqrt() + c * eps + sigma * noise
    else:
        x0c = x0.clamp(-1, 1)
        row[0, 8], row[0, 9] = sch.posterior_mean_coef1[t], sch.posterior_mean_coef2[t]
        row[0, 10] = (0.5 * sch.posterior_log_variance_clipped[t]).exp()
        want = row[0, 8] * x0c + row[0, 9] * x + row[0, 10] * noise
    out2 = torch.cat([cond, null]).contiguous()
    xs = x.clone()
    pe, px = torch.empty(B, chw), torch.empty(B, chw)
    counter = torch.zeros(1, dtype=torch.int32)
    a_ = L.StepArgs()
    a_.out_cond, a_.out_null, a_.x, a_.noise = out2.data_ptr(), out2[B:].data_ptr(), xs.data_ptr(), noise.data_ptr()
    a_.pred_noise, a_.pred_x0, a_.B, a_.chw = pe.data_ptr(), px.data_ptr(), B, chw
    a_.cond_scale, a_.rescaled_phi, a_.keep_parallel_frac, a_.remove_parallel = scale, phi, 0.0, 1
    a_.objective, a_.cfg_plus_plus = L.OBJ[objective], 0
    a_.clip_x0, a_.sampler = (1, 0) if kind == "ddim" else (0, 1)
    a_.coef, a_.step_counter, a_.advance, a_.t_rows = row.data_ptr(), counter.data_ptr(), 1, None
    assert samp.ccdm_sampler_step(C.byref(a_), None) == 0
    assert counter.item() == 1
    assert rel(px.view(shape), x0) < 1e-4 and rel(pe.view(shape), eps) < 1e-4
    assert rel(xs, want) < 1e-4


@pytest.mark.parametrize("objective,use_Hy,weighted", [("pred_x0", True, True), ("pred_noise", False, True),
                                                       ("pred_v", True, False)])
def test_q_sample_and_vicinal_loss(samp, objective, use_Hy, weighted):
    sch = oracle.make_schedule(1000, "cosine", objective)
    g = torch.Generator().manual_seed(3)
    B, c, h, w = 4, 3, 4, 4
    chw = c * h * w
    img01 = torch.rand(B, c, h, w, generator=g)
    noise, noise2 = torch.randn(B, c, h, w, generator=g), torch.randn(B, c, h, w, generator=g)
    cov = (0.5 + torch.rand(B, c, h, w, generator=g)) if use_Hy else None
    keep = torch.tensor([1, 0, 1, 1], dtype=torch.uint8)
    t = torch.tensor([3, 500, 999, 42])
    x0, nz, x_t = torch.empty(B, chw), torch.empty(B, chw), torch.empty(B, chw)
    q = L.QSampleArgs()
    q.img01, q.noise, q.noise2, q.cov = img01.data_ptr(), noise.data_ptr(), noise2.data_ptr(), L.ptr(cov)
    q.keep, q.t, q.normalize = keep.data_ptr(), t.data_ptr(), 1
    q.sqrt_acp, q.sqrt_1m_acp = sch.sqrt_alphas_cumprod.data_ptr(), sch.sqrt_one_minus_alphas_cumprod.data_ptr()
    q.x0, q.noise_out, q.x_t, q.B, q.chw = x0.data_ptr(), nz.data_ptr(), x_t.data_ptr(), B, chw
    assert samp.ccdm_q_sample(C.byref(q), None) == 0
    want_x0 = img01 * 2 - 1
    want_nz = noise.clone()
    if use_Hy:
        want_nz = torch.where(keep.bool()[:, None, None, None], noise * cov.sqrt(), noise2)
    assert rel(x0.view_as(img01), want_x0) < 1e-6 and rel(nz.view_as(img01), want_nz) < 1e-6
    assert rel(x_t.view_as(img01), D.q_sample(sch, want_x0, t, want_nz)) < 1e-6
    # ---- loss + d loss / d model_out vs autograd of the reference formula (diffusion.py:550-594,713-735)
    out = torch.randn(B, c, h, w, generator=g).requires_grad_(True)
    row_w = (0.2 + torch.rand(B, generator=g)) if weighted else None
    if objective == "pred_noise":
        target = want_nz
    elif objective == "pred_x0":
        target = want_x0
    else:
        target = D._at(sch.sqrt_alphas_cumprod, t, want_x0) * want_nz - D._at(sch.sqrt_one_minus_alphas_cumprod, t, want_x0) * want_x0
    e = (out - target) ** 2
    if use_Hy:
        div = cov.clone()
        div[~keep.bool()] = 1.0
        e = e / div
    per = e.flatten(1).sum(1) * sch.loss_weight[t]
    want = (per * (row_w if weighted else 1.0)).sum() / (B * chw)
    want.backward()
    per_k, loss_k, grad_k = torch.empty(B), torch.empty(1), torch.empty(B, chw)
    la = L.LossArgs()
    od = out.detach().contiguous()
    la.model_out, la.x0, la.noise, la.cov = od.data_ptr(), x0.data_ptr(), nz.data_ptr(), L.ptr(cov)
    la.keep, la.t, la.sqrt_acp, la.sqrt_1m_acp = keep.data_ptr(), t.data_ptr(), q.sqrt_acp, q.sqrt_1m_acp
    la.loss_weight, la.row_weight = sch.loss_weight.data_ptr(), L.ptr(row_w)
    la.per_sample, la.loss, la.grad_out = per_k.data_ptr(), loss_k.data_ptr(), grad_k.data_ptr()
    la.B, la.chw, la.objective = B, chw, L.OBJ[objective]
    assert samp.ccdm_vicinal_loss(C.byref(la), None) == 0
    assert abs(loss_k.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel(grad_k.view_as(out), out.grad) < 1e-5


@pytest.mark.parametrize("vic,multi", [("hv", False), ("sv", False), ("hv", True), ("shv", True), ("ssv", True)])
def test_vicinal_weights_match_oracle(samp, vic, multi):
    g = torch.Generator().manual_seed(4)
    B, kappa = 9, 0.25
    labels = torch.rand(B, 3, generator=g) if multi else torch.rand(B, generator=g)
    keep = torch.ones(B, dtype=torch.uint8)
    hard, sliced = vic in ("hv", "shv"), vic in ("shv", "ssv")
    nu = 0.0 if hard else 1.0 / kappa ** 2
    w = torch.empty(B)
    if sliced and multi:
        v = torch.randn(2, 3, generator=g)
        proj = (labels @ F.normalize(v, dim=1, eps=1e-8).t()).contiguous()
        thr = (kappa * torch.norm(v, dim=1) + 1e-8).float().contiguous()
        assert samp.ccdm_vicinal_weights(proj.data_ptr(), B, 2, 0, int(hard), thr.data_ptr(), nu, keep.data_ptr(), w.data_ptr(), None) == 0
        want = D.vicinal_batch_weights(labels, vicinity_type=vic, kappa=kappa, num_projections=2, cached_vectors=v)
    else:
        thr = torch.full((1,), kappa)
        proj = labels.float().reshape(B, -1).contiguous()
        assert samp.ccdm_vicinal_weights(proj.data_ptr(), B, proj.shape[1], int(multi), int(hard), thr.data_ptr(), nu,
                                         keep.data_ptr(), w.data_ptr(), None) == 0
        want = D.vicinal_batch_weights(labels, vicinity_type="hv" if hard else "sv", kappa=kappa)
    assert rel(w, want) < 1e-5


# ------------------------------------------------------------------------------------------------- kernels.cu

@pytest.mark.parametrize("C_,flags,rows", [(64, L.EPI_SS | L.EPI_SILU, 40), (72, L.EPI_SILU | L.EPI_RESID | L.EPI_SUMSQ_OUT, 33),
                                           (128, L.EPI_RESID, 20), (576, L.EPI_SS | L.EPI_SILU, 12), (32, 0, 50)])
def test_rmsnorm_act(kern, C_, flags, rows):
    g = torch.Generator().manual_seed(5)
    B = 2
    z = torch.randn(B * rows, C_, generator=g).to(torch.bfloat16)
    gain = 1 + 0.1 * torch.randn(C_, generator=g)
    ss = 0.3 * torch.randn(B, 16 + 2 * C_, generator=g)
    resid = torch.randn(B * rows, C_, generator=g).to(torch.bfloat16)
    out = torch.empty_like(z)
    rowss = torch.empty(B * rows)
    assert kern.ccdm_rmsnorm_act(z.data_ptr(), out.data_ptr(), B * rows, C_, rows, gain.data_ptr(), math.sqrt(C_),
                                 ss.data_ptr() if flags & L.EPI_SS else None, ss.shape[1], 16,
                                 resid.data_ptr() if flags & L.EPI_RESID else None,
                                 rowss.data_ptr() if flags & L.EPI_SUMSQ_OUT else None, flags, None) == 0
    v = F.normalize(z.float(), dim=-1) * gain * math.sqrt(C_)
    if flags & L.EPI_SS:
        b = torch.arange(B * rows) // rows
        v = v * (1 + ss[b, 16:16 + C_]) + ss[b, 16 + C_:16 + 2 * C_]
    if flags & L.EPI_SILU:
        v = F.silu(v)
    if flags & L.EPI_RESID:
        v = v + resid.float()
    assert rel(out, v) < 6e-3
    if flags & L.EPI_SUMSQ_OUT:
        assert rel(rowss, out.float().pow(2).sum(-1)) < 1e-4


@pytest.mark.parametrize("act", [0, 1, 2])
def test_affine_act(kern, act):
    g = torch.Generator().manual_seed(6)
    B, rps, C_ = 3, 10, 40
    x = torch.randn(B * rps, C_, generator=g).to(torch.bfloat16)
    ss = 0.3 * torch.randn(B, 8 + 2 * C_, generator=g)
    out = torch.empty_like(x)
    assert kern.ccdm_affine_act(x.data_ptr(), out.data_ptr(), B * rps, C_, rps, ss.data_ptr(), ss.shape[1], 8, act, None) == 0
    v = x.float().reshape(B, rps, C_) * (1 + ss[:, None, 8:8 + C_]) + ss[:, None, 8 + C_:8 + 2 * C_]
    v = {0: lambda u: u, 1: F.relu, 2: F.silu}[act](v)
    assert rel(out.reshape(B, rps, C_), v) < 6e-3


def test_head_conv1_forward_and_backward(kern, trk):
    g = torch.Generator().manual_seed(7)
    B, H, W, Cin, Cout = 2, 5, 6, 64, 3
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / 8
    b = torch.randn(Cout, generator=g)
    out = torch.empty(B, Cout, H, W)
    assert kern.ccdm_head_conv1(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, None) == 0
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br)
    assert rel(out, ref) < 1e-5
    dout = torch.randn(B, Cout, H, W, generator=g)
    ref.backward(dout)
    dh = torch.empty_like(x)
    dw, db = torch.zeros(Cout, Cin), torch.zeros(Cout)
    assert trk.ccdm_head_conv1_bwd(dout.data_ptr(), x.data_ptr(), w.data_ptr(), dh.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                   B, H, W, Cin, Cout, None) == 0
    assert rel(dh.float().permute(0, 3, 1, 2), xr.grad) < 6e-3
    assert rel(dw, wr.grad.reshape(Cout, Cin)) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("dh,n", [(32, 16), (16, 9), (64, 70)])
def test_attention_small_forward_and_backward(kern, trk, dh, n):
    g = torch.Generator().manual_seed(8)
    B, heads = 2, 2
    hid = heads * dh
    scale = dh ** -0.5
    qkv = torch.randn(B, n, 3 * hid, generator=g).to(torch.bfloat16)
    out = torch.empty(B, n, hid, dtype=torch.bfloat16)
    assert kern.ccdm_attention_small(qkv.data_ptr(), out.data_ptr(), B, n, heads, dh, scale, None) == 0
    f = qkv.float().requires_grad_(True)
    q, k, v = (f[..., i * hid:(i + 1) * hid].reshape(B, n, heads, dh) for i in range(3))
    att = torch.einsum("bihd,bjhd->bhij", q * scale, k).softmax(-1)
    o = torch.einsum("bhij,bjhd->bihd", att, v).reshape(B, n, hid)
    assert rel(out, o) < 6e-3
    dout = torch.randn(B, n, hid, generator=g).to(torch.bfloat16)
    o.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    assert trk.ccdm_attention_small_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, n, heads, dh, scale, None) == 0
    assert rel(dqkv, f.grad) < 6e-3


@pytest.mark.parametrize("bn,train,in_dim,B", [(False, 0, 64, 5), (True, 0, 128, 40), (True, 1, 32, 7), (False, 0, 30, 3)])
def test_linear_small(kern, bn, train, in_dim, B):
    g = torch.Generator().manual_seed(9)
    out_dim = 24
    x = torch.randn(B, in_dim, generator=g)
    w, b = torch.randn(out_dim, in_dim, generator=g) / 6, torch.randn(out_dim, generator=g)
    bw, bb = 1 + 0.1 * torch.randn(out_dim, generator=g), 0.1 * torch.randn(out_dim, generator=g)
    rm, rv = 0.2 * torch.randn(out_dim, generator=g), 0.5 + torch.rand(out_dim, generator=g)
    rm0, rv0 = rm.clone(), rv.clone()
    y = torch.empty(B, out_dim)
    args = (bw.data_ptr(), bb.data_ptr(), rm.data_ptr(), rv.data_ptr()) if bn else (None, None, None, None)
    assert kern.ccdm_linear_small(x.data_ptr(), B, in_dim, w.data_ptr(), b.data_ptr(), out_dim, *args, train, L.ACT_RELU,
                                  y.data_ptr(), out_dim, None) == 0
    ref = F.linear(x, w, b)
    if bn:
        ref = F.batch_norm(ref, rm0.clone(), rv0.clone(), bw, bb, training=bool(train), momentum=0.1, eps=1e-5)
    assert rel(y, F.relu(ref)) < 1e-5
    if bn and train:                                        # running statistics updated like nn.BatchNorm1d
        m2, v2 = rm0.clone(), rv0.clone()
        F.batch_norm(F.linear(x, w, b), m2, v2, bw, bb, training=True, momentum=0.1, eps=1e-5)
        assert rel(rm, m2) < 1e-5 and rel(rv, v2) < 1e-5


def test_time_features_select_null_silu_concat(kern):
    from oracle.unet_ref import time_features
    t = torch.tensor([0, 1, 17, 500, 999])
    out = torch.empty(5, 64)
    assert kern.ccdm_time_features(t.data_ptr(), 5, 64, out.data_ptr(), None) == 0
    assert (out - time_features(t.float(), 64)).abs().max().item() < 2e-4
    g = torch.Generator().manual_seed(10)
    c, null = torch.randn(5, 12, generator=g), -torch.rand(12, generator=g)
    keep = torch.tensor([1, 0, 1, 0, 0], dtype=torch.uint8)
    want = torch.where(keep.bool()[:, None], c, null[None])
    assert kern.ccdm_select_null(c.data_ptr(), keep.data_ptr(), 0, null.data_ptr(), 5, 12, None) == 0
    assert torch.equal(c, want)
    te, ce = torch.randn(5, 8, generator=g), torch.randn(5, 16, generator=g)
    o = torch.empty(5, 24, dtype=torch.bfloat16)
    assert kern.ccdm_silu_concat_bf16(te.data_ptr(), 8, ce.data_ptr(), 16, 5, o.data_ptr(), None) == 0
    assert rel(o, F.silu(torch.cat([te, ce], 1))) < 6e-3


# ------------------------------------------------------------------------------------------------- backward.cu / train_kernels.cu

@pytest.mark.parametrize("C_,use_ss,silu,rows", [(64, True, True, 40), (72, False, True, 33), (128, True, False, 12),
                                                 (576, True, True, 6), (64, True, True, 300), (144, True, True, 100),
                                                 (288, True, True, 70), (576, False, True, 30)])
def test_block_backward_matches_autograd(bwd, C_, use_ss, silu, rows):
    g = torch.Generator().manual_seed(11)
    B = 3
    z = torch.randn(B * rows, C_, generator=g).to(torch.bfloat16)
    dy = torch.randn(B * rows, C_, generator=g).to(torch.bfloat16)
    gain = 1 + 0.1 * torch.randn(C_, generator=g)
    ss = 0.3 * torch.randn(B, 2 * C_, generator=g)
    flags = (L.EPI_SILU if silu else 0) | (L.EPI_SS if use_ss else 0)
    dz = torch.empty_like(z)
    zbuf = torch.zeros(3 * B * C_ + 2 * C_)
    sums, dgain, dbias = zbuf[:3 * B * C_], zbuf[3 * B * C_:3 * B * C_ + C_], zbuf[3 * B * C_ + C_:]
    gm = math.sqrt(C_)
    assert bwd.ccdm_block_bwd(dy.data_ptr(), z.data_ptr(), dz.data_ptr(), B * rows, C_, rows, gain.data_ptr(), gm,
                              ss.data_ptr() if use_ss else None, 2 * C_ if use_ss else 0, 0, sums.data_ptr(), flags, None) == 0
    d_ss = torch.empty_like(ss)
    assert bwd.ccdm_block_bwd_finish(sums.data_ptr(), B, C_, gain.data_ptr(), gm, ss.data_ptr() if use_ss else None,
                                     2 * C_ if use_ss else 0, 0, d_ss.data_ptr() if use_ss else None, dgain.data_ptr(),
                                     dbias.data_ptr(), None) == 0
    zr = z.float().reshape(B, rows, C_).requires_grad_(True)
    gr, sr = gain.clone().requires_grad_(True), ss.clone().requires_grad_(True)
    n = F.normalize(zr, dim=-1) * gr * gm
    if use_ss:
        n = n * (1 + sr[:, None, :C_]) + sr[:, None, C_:]
    y = F.silu(n) if silu else n
    y.backward(dy.float().reshape(B, rows, C_))
    assert rel(dz.reshape(B, rows, C_), zr.grad) < 8e-3
    assert rel(dgain, gr.grad) < 1e-3
    assert rel(dbias, zr.grad.sum((0, 1))) < 1e-3            # bias gradient = column sums of dz (fp32, before the bf16 store)
    if use_ss:
        assert rel(d_ss, sr.grad) < 1e-3


def test_colsum(trk):
    g = torch.Generator().manual_seed(12)
    x = torch.randn(77, 72, generator=g).to(torch.bfloat16)
    out = torch.full((72,), 2.0)
    assert trk.ccdm_colsum_bf16(x.data_ptr(), 77, 72, out.data_ptr(), None) == 0
    assert rel(out - 2.0, x.float().sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------- optim.cu

def test_fused_adam_matches_torch_adam_with_clipping(opt):
    g = torch.Generator().manual_seed(13)
    shapes = [(7, 5), (33,), (4, 3, 3, 3), (1,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    topt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.99), eps=1e-8)
    offs, tot = [], 0
    for p in ps:
        offs.append(tot)
        tot += (p.numel() + 3) // 4 * 4
    flat_g, m, v = torch.zeros(tot), torch.zeros(tot), torch.zeros(tot)
    step, sumsq = torch.zeros(1), torch.zeros(1, dtype=torch.float64)
    ptrs = torch.tensor([p.data_ptr() for p in ps], dtype=torch.int64)
    offt = torch.tensor(offs, dtype=torch.int64)
    ns = torch.tensor([p.numel() for p in ps], dtype=torch.int32)
    for it in range(4):
        grads = [torch.randn(s, generator=g) * (3.0 if it % 2 == 0 else 0.05) for s in shapes]      # clipped / not clipped
        flat_g.zero_()
        for r, gr, o in zip(ref, grads, offs):
            r.grad = gr.clone()
            flat_g[o:o + gr.numel()] = gr.flatten()
        want_norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        topt.step()
        assert opt.ccdm_fused_adam(ptrs.data_ptr(), offt.data_ptr(), ns.data_ptr(), len(ps), flat_g.data_ptr(), m.data_ptr(),
                                   v.data_ptr(), tot, step.data_ptr(), sumsq.data_ptr(), 1.0, 1e-2, 0.9, 0.99, 1e-8, 0.0, None) == 0
        assert abs(sumsq.sqrt().item() - want_norm.item()) < 1e-5 * want_norm.item()
        for p, r in zip(ps, ref):
            assert rel(p, r.detach()) < 1e-5, it


def test_multi_lerp(opt):
    g = torch.Generator().manual_seed(14)
    dst = [torch.randn(50, generator=g), torch.randn(3, 7, generator=g)]
    src = [torch.randn(50, generator=g), torch.randn(3, 7, generator=g)]
    want = [d.clone().lerp_(s, 0.25) for d, s in zip(dst, src)]
    dp = torch.tensor([d.data_ptr() for d in dst], dtype=torch.int64)
    sp = torch.tensor([s.data_ptr() for s in src], dtype=torch.int64)
    ns = torch.tensor([d.numel() for d in dst], dtype=torch.int32)
    w = torch.tensor([0.25])
    assert opt.ccdm_multi_lerp(dp.data_ptr(), sp.data_ptr(), ns.data_ptr(), 2, w.data_ptr(), None) == 0
    for d, x in zip(dst, want):
        assert rel(d, x) < 1e-6


# ------------------------------------------------------------------------------------------------- weight packing (tapgemm.cu / wgrad.cu)

@pytest.fixture(scope="module")
def packlib():
    from tests.hostsim.build import build_extract
    a = C.CDLL(build_extract("tapgemm.cu", ["pack_weights_kernel"], ["ccdm_pack_weights_at", "ccdm_pack_weights"]))
    b = C.CDLL(build_extract("wgrad.cu", ["unpack_wgrad_kernel", "pack_weights_t_kernel"],
                             ["ccdm_unpack_wgrad_slots", "ccdm_unpack_wgrad", "ccdm_pack_weights_t"]))
    for h, names in ((a, ["ccdm_pack_weights", "ccdm_pack_weights_at"]),
                     (b, ["ccdm_unpack_wgrad", "ccdm_unpack_wgrad_slots", "ccdm_pack_weights_t"])):
        for n in names:
            fn = getattr(h, n)
            fn.restype, fn.argtypes = L.SIGNATURES[n]
    return a, b


def _i32(rows):
    return torch.tensor(rows, dtype=torch.int32).contiguous()


@pytest.mark.parametrize("kind,cins,cout,reuse", [("3x3", (40, 24), 24, False), ("down3x3s2", (64,), 32, True),
                                                  ("up2x3x3", (32,), 32, True), ("1x1", (72,), 32, False)])
def test_pack_and_unpack_kernels_match_the_emulator(packlib, kind, cins, cout, reuse):
    """The real pack / unpack kernels (source extracted from tapgemm.cu / wgrad.cu) against tests/emu.py's packing rule,
    including the zero-weight padding taps of the vanilla UNet's stride-2 conv (tapmask 0)."""
    from ccdm_b200.plan import KB, plan_conv
    from tests.emu import pack_weights_emu
    pk, wg = packlib
    taps = {"3x3": 9, "1x1": 1, "down3x3s2": 9, "up2x3x3": 9, "down4x4s2": 16}[kind]
    k = int(math.isqrt(taps))
    g = torch.Generator().manual_seed(15)
    w = torch.randn(cout, sum(cins), k, k, generator=g)
    plan = plan_conv(kind, cins, cout, reuse_rows=reuse)
    n_rows = (cout + 31) // 32 * 32
    packed = torch.full((plan.nz * n_rows, plan.nkb * KB), 7.0).to(torch.bfloat16)
    ps = _i32(plan.psched)
    assert pk.ccdm_pack_weights(w.data_ptr(), cout, sum(cins), taps, ps.data_ptr(), plan.nz, plan.nkb, n_rows, None, 1.0,
                                packed.data_ptr(), None) == 0
    want = pack_weights_emu(plan, w, n_rows).reshape(plan.nz * n_rows, -1)
    assert torch.equal(packed, want.to(torch.bfloat16))
    # unpack(pack(w)) scatters every packed block back to its taps: folded taps receive the same value several times
    dw = torch.empty_like(w)
    assert wg.ccdm_unpack_wgrad(want.contiguous().data_ptr(), dw.data_ptr(), cout, sum(cins), taps, ps.data_ptr(), plan.nz,
                                plan.nkb, n_rows, None, 1.0, 0, None) == 0
    mult = torch.zeros(taps)
    cnt = torch.zeros(taps)
    for z in range(plan.nz):
        for kb in range(plan.nkb):
            c0, nv, mask, _ = plan.psched[z * plan.nkb + kb]
            if c0 == 0 and nv > 0:
                ts = [t for t in range(taps) if mask >> t & 1]
                for t in ts:
                    cnt[t] += 1
    assert (cnt > 0).all()                                  # every filter tap is covered by some block
    if kind != "up2x3x3":                                   # no tap folding: exact round trip
        assert rel(dw, w) < 1e-6


@pytest.mark.parametrize("kind,cin,cout,reuse", [("3x3", 32, 32, True), ("down3x3s2", 32, 32, True), ("down3x3s2", 40, 24, False)])
def test_transposed_pack_kernel_matches_the_emulator(packlib, kind, cin, cout, reuse):
    from ccdm_b200.plan import KB, plan_conv
    from tests.emu import pack_weights_emu
    _, wg = packlib
    taps = 9
    g = torch.Generator().manual_seed(16)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    plan = plan_conv(kind + "_dgrad", (cout,), cin, reuse_rows=reuse)
    n_rows = (cin + 31) // 32 * 32
    packed = torch.full((plan.nz * n_rows, plan.nkb * KB), 7.0).to(torch.bfloat16)
    ps = _i32(plan.psched)
    assert wg.ccdm_pack_weights_t(w.data_ptr(), cout, cin, taps, ps.data_ptr(), plan.nz, plan.nkb, n_rows, 0, cin,
                                  packed.data_ptr(), None) == 0
    want = pack_weights_emu(plan, w, n_rows).reshape(plan.nz * n_rows, -1)
    assert torch.equal(packed, want.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------- round-2 additions

@pytest.mark.parametrize("B,C_,groups,act", [(5, 128, 8, 1), (33, 512, 8, 1), (2, 48, 4, 0), (7, 4096, 8, 1)])
def test_groupnorm_rows_matches_torch(kern, B, C_, groups, act):
    """ccdm_groupnorm_rows = nn.GroupNorm(groups, C) on a [B, C] matrix + ReLU, in place (the learned label MLPs,
    models/resnet_y2h.py:143-173)."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, C_, generator=g) * 3 + 1
    gamma, beta = 1 + 0.3 * torch.randn(C_, generator=g), 0.2 * torch.randn(C_, generator=g)
    want = F.group_norm(x, groups, gamma, beta, 1e-5)
    want = F.relu(want) if act == 1 else want
    y = x.clone()
    assert kern.ccdm_groupnorm_rows(y.data_ptr(), B, C_, groups, gamma.data_ptr(), beta.data_ptr(), 1e-5, act, None) == 0
    assert rel(y, want) < 2e-6
    assert kern.ccdm_groupnorm_rows(y.data_ptr(), B, 50, 8, gamma.data_ptr(), beta.data_ptr(), 1e-5, act, None) == -1   # C % groups


def test_pack_multi_and_pack_at_match_the_single_tensor_kernels(packlib):
    """ccdm_pack_multi (one launch over a job table: the training step's weight re-pack, both layouts) and
    ccdm_pack_weights_at (a conv and its block's shortcut in ONE packed matrix, CCDM_EPI_RESACC) against the emulator."""
    from ccdm_b200.plan import KB, plan_conv
    from tests.emu import pack_weights_emu
    from tests.hostsim.build import build
    pm = C.CDLL(build("packmulti.cu"))
    pm.ccdm_pack_multi.restype, pm.ccdm_pack_multi.argtypes = L.SIGNATURES["ccdm_pack_multi"]
    pk, _ = packlib
    g = torch.Generator().manual_seed(21)
    specs = [("3x3", (40, 24), 24, 0), ("3x3", (72,), 72, 0), ("1x1", (96,), 32, 0), ("3x3_dgrad", (24,), 40, 1)]
    jobs, outs, wants, keep = [], [], [], []
    for kind, cins, cout, mode in specs:
        base = kind[:-6] if mode else kind
        taps = 9 if base == "3x3" else 1
        k = 3 if taps == 9 else 1
        plan = plan_conv(kind, cins, cout, reuse_rows=(taps == 9))
        n_rows = (cout + 31) // 32 * 32
        # forward weights are [Cout, Cin, k, k]; a dgrad plan packs the FORWARD weight [Cout_fwd = cins[0], Cin_fwd >= cout, ...]
        w = torch.randn(cout, sum(cins), k, k, generator=g) if not mode else torch.randn(cins[0], cout + 8, k, k, generator=g)
        packed = torch.full((plan.nz * n_rows, plan.nkb * KB), 7.0).to(torch.bfloat16)
        ps = _i32(plan.psched)
        j = L.PackJob()
        j.w, j.out, j.psched, j.mode = w.data_ptr(), packed.data_ptr(), ps.data_ptr(), mode
        j.cout, j.cin_total, j.ntaps, j.nkb, j.n_rows = w.shape[0], w.shape[1], taps, plan.nkb, n_rows
        j.n_off, j.n_count, j.total = (4 if mode else 0), cout, plan.nz * n_rows * plan.nkb * KB
        jobs.append(j)
        outs.append(packed)
        wants.append(pack_weights_emu(plan, w, n_rows, n_off=j.n_off, n_count=cout if mode else None).reshape(plan.nz * n_rows, -1))
        keep += [w, ps]
    table = torch.frombuffer(bytearray(b"".join(bytes(j) for j in jobs)), dtype=torch.uint8)
    assert pm.ccdm_pack_multi(table.data_ptr(), len(jobs), None) == 0
    for got, want in zip(outs, wants):
        assert torch.equal(got, want.to(torch.bfloat16))
    # pack_weights_at: main 3x3 weight in K blocks [0, 9), the 1x1 shortcut weight in block 9 of the same matrix
    main, sc = plan_conv("3x3", (48,), 32, reuse_rows=True), plan_conv("1x1", (48,), 32)
    w3, w1 = torch.randn(32, 48, 3, 3, generator=g), torch.randn(32, 48, 1, 1, generator=g)
    both = torch.full((32, (main.nkb + sc.nkb) * KB), 5.0).to(torch.bfloat16)
    p3, p1 = _i32(main.psched), _i32(sc.psched)
    assert pk.ccdm_pack_weights_at(w3.data_ptr(), 32, 48, 9, p3.data_ptr(), 1, main.nkb, 32, None, 1.0, both.data_ptr(),
                                   main.nkb + sc.nkb, 0, None) == 0
    assert pk.ccdm_pack_weights_at(w1.data_ptr(), 32, 48, 1, p1.data_ptr(), 1, sc.nkb, 32, None, 1.0, both.data_ptr(),
                                   main.nkb + sc.nkb, main.nkb, None) == 0
    want = torch.cat([pack_weights_emu(main, w3, 32)[0], pack_weights_emu(sc, w1, 32)[0]], dim=1)
    assert torch.equal(both, want.to(torch.bfloat16))
    assert pk.ccdm_pack_weights_at(w1.data_ptr(), 32, 48, 1, p1.data_ptr(), 1, sc.nkb, 32, None, 1.0, both.data_ptr(),
                                   main.nkb, main.nkb, None) == -1                       # K blocks out of range


def test_unpack_wgrad_slots_sums_partial_gradients_in_order(packlib):
    """ccdm_unpack_wgrad_slots over three partial gradients (ccdm_wgrad_args.slots, padded slot stride) equals
    ccdm_unpack_wgrad of their sum -- for the halo plan's K-block order too (tap (r, q) = block g*9 + r*3 + q)."""
    from ccdm_b200.plan import KB, plan_conv
    _, wg = packlib
    g = torch.Generator().manual_seed(21)
    for plan in (plan_conv("3x3", (72,), 40, halo=True), plan_conv("3x3", (40, 24), 24, reuse_rows=True)):
        cout, cin = plan.cout, sum(plan.cins)
        n_rows = (cout + 31) // 32 * 32
        size = plan.nz * n_rows * plan.nkb * KB
        stride = size + 192
        parts = torch.randn(3, stride, generator=g)
        ps = _i32(plan.psched)
        got = torch.empty(cout, cin, 3, 3)
        assert wg.ccdm_unpack_wgrad_slots(parts.data_ptr(), 3, stride, got.data_ptr(), cout, cin, 9, ps.data_ptr(), plan.nz,
                                          plan.nkb, n_rows, None, 1.0, 0, None) == 0
        total = ((parts[0, :size] + parts[1, :size]) + parts[2, :size]).contiguous()
        want = torch.empty_like(got)
        assert wg.ccdm_unpack_wgrad(total.data_ptr(), want.data_ptr(), cout, cin, 9, ps.data_ptr(), plan.nz, plan.nkb, n_rows,
                                    None, 1.0, 0, None) == 0
        assert torch.equal(got, want)
        acc = torch.ones_like(got)                                 # accumulate = 1 adds to what is there
        assert wg.ccdm_unpack_wgrad_slots(parts.data_ptr(), 3, stride, acc.data_ptr(), cout, cin, 9, ps.data_ptr(), plan.nz,
                                          plan.nkb, n_rows, None, 1.0, 1, None) == 0
        assert torch.allclose(acc, want + 1, atol=1e-6)
        assert wg.ccdm_unpack_wgrad_slots(parts.data_ptr(), 3, size - 1, got.data_ptr(), cout, cin, 9, ps.data_ptr(), plan.nz,
                                          plan.nkb, n_rows, None, 1.0, 0, None) != 0      # slots must not overlap


def test_plan_ksteps_tables():
    """ConvPlan.ksteps (ccdm_tapgemm_args.ksteps): live K steps of every load group's 64-channel block."""
    from ccdm_b200.plan import plan_conv
    assert plan_conv("3x3", (64,), 64, reuse_rows=True).ksteps is None
    assert plan_conv("3x3", (72,), 72, reuse_rows=True).ksteps == [[4], [1]] * 3
    assert plan_conv("3x3", (288,), 288, reuse_rows=True).ksteps == [[4], [4], [4], [4], [2]] * 3
    assert plan_conv("1x1", (72, 144), 72).ksteps == [[4], [1], [4], [4], [1]]
    assert plan_conv("3x3_dgrad", (72,), 144, reuse_rows=True).ksteps == [[4], [1]] * 3
    assert plan_conv("3x3", (72,), 72, halo=True).ksteps is None          # halo boxes issue whole blocks
    up = plan_conv("up2x3x3", (72,), 72, reuse_rows=True)
    assert up.ksteps is not None and len(up.ksteps) == up.nz * up.ngroups and {v[0] for v in up.ksteps} == {1, 4}
