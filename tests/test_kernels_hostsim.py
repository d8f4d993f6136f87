"""The CUDA-core kernels of ccdm_b200/csrc/groupnorm.cu, compiled for the HOST from their own source
(tests/hostsim: every CUDA thread a cooperative fiber, real barriers / shuffles / atomics) and checked against torch.

This exercises the kernels' index arithmetic, shared-memory staging, grid sizing and argument checks without a GPU.
It says nothing about performance and is not a fallback: the product loads only libccdm_b200.so (sm_100a).
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from tests.hostsim.build import build

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


@pytest.fixture(scope="module")
def lib():
    h = C.CDLL(build("groupnorm.cu"))
    h.ccdm_channel_stats.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp]
    h.ccdm_groupnorm_coef.argtypes = [vp, i32, i32, i32, i64, f32, vp, vp, vp, i32, i32, i32, vp, vp]
    h.ccdm_attention_tokens.argtypes = [vp, vp, i32, i32, i32, i32, f32, i32, vp]
    h.ccdm_time_features_adm.argtypes = [vp, i32, i32, f32, vp, vp]
    h.hostsim_last_error.restype = C.c_char_p
    return h


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("cs,groups,hw,with_ss", [((64,), 8, (8, 8), False), ((64, 32), 4, (6, 5), True),
                                                  ((72,), 8, (4, 7), True), ((128, 64), 8, (16, 16), True),
                                                  ((8,), 2, (3, 3), False)])
def test_groupnorm_stats_and_coef(lib, cs, groups, hw, with_ss):
    g = torch.Generator().manual_seed(5)
    B, (h, w), ctot = 3, hw, sum(cs)
    xs = [(torch.randn(B, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16).contiguous() for c in cs]
    gamma = 1 + 0.2 * torch.randn(ctot, generator=g)
    beta = 0.1 * torch.randn(ctot, generator=g)
    ss = 0.3 * torch.randn(B, 40 + 2 * ctot, generator=g) if with_ss else None
    sums = torch.full((B, 2, ctot), 7.0)                        # stale contents: zero_first must clear them
    off = 0
    for i, x in enumerate(xs):
        assert lib.ccdm_channel_stats(x.data_ptr(), B, h * w, x.shape[3], sums.data_ptr(), ctot, off, int(i == 0), None) == 0, \
            lib.hostsim_last_error()
        off += x.shape[3]
    cat = torch.cat([x.float() for x in xs], -1)
    want_sums = torch.stack([cat.sum((1, 2)), cat.pow(2).sum((1, 2))], 1)
    assert rel(sums, want_sums) < 1e-5
    coef = torch.empty(B, 2 * ctot)
    assert lib.ccdm_groupnorm_coef(sums.data_ptr(), B, ctot, groups, h * w, 1e-5, gamma.data_ptr(), beta.data_ptr(),
                                   ss.data_ptr() if with_ss else None, ss.shape[1] if with_ss else 0, 40, cs[0],
                                   coef.data_ptr(), None) == 0, lib.hostsim_last_error()
    # apply the coefficients per source (what ccdm_affine_act does) and compare with torch's group_norm
    outs, off = [], 0
    for x in xs:
        c = x.shape[3]
        outs.append(x.float() * (1 + coef[:, None, None, off:off + c]) + coef[:, None, None, off + c:off + 2 * c])
        off += 2 * c
    y = F.group_norm(cat.permute(0, 3, 1, 2), groups, gamma, beta, eps=1e-5)
    if with_ss:
        y = y * (1 + ss[:, 40:40 + ctot, None, None]) + ss[:, 40 + ctot:40 + 2 * ctot, None, None]
    assert rel(torch.cat(outs, -1), y.permute(0, 2, 3, 1)) < 2e-5


def test_channel_stats_rejects_bad_shapes(lib):
    x = torch.zeros(1, 4, 12, dtype=torch.bfloat16)
    s = torch.zeros(1, 2, 12)
    assert lib.ccdm_channel_stats(x.data_ptr(), 1, 4, 12, s.data_ptr(), 12, 0, 1, None) == -2      # C % 8 != 0
    assert lib.ccdm_channel_stats(x.data_ptr(), 1, 4, 8, s.data_ptr(), 12, 8, 1, None) == -1       # c_off + C > ld


@pytest.mark.parametrize("dh,n,head_major", [(16, 16, 0), (16, 70, 1), (32, 64, 1), (64, 33, 0), (128, 64, 1), (128, 9, 0)])
def test_attention_tokens(lib, dh, n, head_major):
    g = torch.Generator().manual_seed(n + dh)
    B, heads = 2, 2
    hid = heads * dh
    qkv = torch.randn(B, n, 3 * hid, generator=g).to(torch.bfloat16).contiguous()
    out = torch.zeros(B, n, hid, dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(dh)
    assert lib.ccdm_attention_tokens(qkv.data_ptr(), out.data_ptr(), B, n, heads, dh, scale, head_major, None) == 0
    f = qkv.float()
    v5 = f.reshape(B, n, heads, 3, dh).permute(0, 1, 3, 2, 4) if head_major else f.reshape(B, n, 3, heads, dh)
    q, k, v = v5[:, :, 0] * scale, v5[:, :, 1], v5[:, :, 2]
    att = torch.einsum("bihd,bjhd->bhij", q, k).softmax(-1)
    want = torch.einsum("bhij,bjhd->bihd", att, v).reshape(B, n, hid)
    assert rel(out, want) < 5e-3          # bf16 output rounding


def test_attention_tokens_rejects_unsupported_head_width(lib):
    assert lib.ccdm_attention_tokens(1, 1, 1, 4, 1, 24, 1.0, 0, None) == -2


def test_time_features_adm(lib):
    from oracle.vanilla_unet_ref import timestep_embedding
    t = torch.tensor([0, 1, 17, 500, 999])
    for dim in (32, 64, 128):
        out = torch.empty(5, dim)
        assert lib.ccdm_time_features_adm(t.data_ptr(), 5, dim, 10000.0, out.data_ptr(), None) == 0
        assert (out - timestep_embedding(t, dim)).abs().max().item() < 1e-4


@pytest.mark.parametrize("cs,groups,hw,with_ss,act", [((64,), 8, (8, 8), False, 2), ((64, 32), 4, (6, 5), True, 2),
                                                      ((72,), 8, (4, 7), True, 2), ((16, 8), 3, (5, 5), True, 0),
                                                      ((32,), 4, (4, 4), False, 1)])
def test_groupnorm_backward_chain_matches_autograd(lib, cs, groups, hw, with_ss, act):
    """norm_bwd_stats -> groupnorm_bwd_coef -> norm_bwd_apply == autograd of act(group_norm(cat(x)) [*(1+scale)+shift])
    w.r.t. x, gamma, beta, scale, shift -- with groups that straddle the two concatenated sources."""
    lib.ccdm_norm_bwd_stats.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.ccdm_groupnorm_bwd_coef.argtypes = [vp, vp, i32, i32, i32, i64, f32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.ccdm_norm_bwd_apply.argtypes = [vp, vp, vp, i64, i32, i32, vp, i32, i32, vp, i32, i32, i32, vp]
    g = torch.Generator().manual_seed(17)
    B, (h, w), ctot = 3, hw, sum(cs)
    xs = [(torch.randn(B, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16).contiguous() for c in cs]
    dys = [torch.randn(B, h, w, c, generator=g).to(torch.bfloat16).contiguous() for c in cs]
    gamma = 1 + 0.2 * torch.randn(ctot, generator=g)
    beta = 0.1 * torch.randn(ctot, generator=g)
    SS_OFF = 24
    ss = 0.3 * torch.randn(B, SS_OFF + 2 * ctot, generator=g) if with_ss else None
    ssp = ss.data_ptr() if with_ss else None
    ss_ld = ss.shape[1] if with_ss else 0
    # ---- forward statistics and coefficients (kernels under test in the forward tests above)
    sums = torch.zeros(B, 2, ctot)
    off = 0
    for i, x in enumerate(xs):
        assert lib.ccdm_channel_stats(x.data_ptr(), B, h * w, x.shape[3], sums.data_ptr(), ctot, off, int(i == 0), None) == 0
        off += x.shape[3]
    coef = torch.empty(B, 2 * ctot)
    assert lib.ccdm_groupnorm_coef(sums.data_ptr(), B, ctot, groups, h * w, 1e-5, gamma.data_ptr(), beta.data_ptr(), ssp,
                                   ss_ld, SS_OFF, cs[0], coef.data_ptr(), None) == 0
    # ---- backward
    bsums = torch.full((B, 2, ctot), 3.0)
    off = coff = 0
    for i, (x, dy) in enumerate(zip(xs, dys)):
        c = x.shape[3]
        assert lib.ccdm_norm_bwd_stats(dy.data_ptr(), x.data_ptr(), B, h * w, c, coef.data_ptr(), 2 * ctot, coff, act,
                                       bsums.data_ptr(), ctot, off, int(i == 0), None) == 0, lib.hostsim_last_error()
        off, coff = off + c, coff + 2 * c
    bcoef = torch.empty(B, 3 * ctot)
    dgamma, dbeta = torch.full((ctot,), 0.5), torch.full((ctot,), -0.25)      # accumulate semantics
    d_ss = torch.empty(B, 2 * ctot)
    assert lib.ccdm_groupnorm_bwd_coef(sums.data_ptr(), bsums.data_ptr(), B, ctot, groups, h * w, 1e-5, gamma.data_ptr(),
                                       beta.data_ptr(), ssp, ss_ld, SS_OFF, cs[0], bcoef.data_ptr(), dgamma.data_ptr(),
                                       dbeta.data_ptr(), d_ss.data_ptr(), None) == 0, lib.hostsim_last_error()
    dxs, coff, boff = [], 0, 0
    for x, dy in zip(xs, dys):
        c = x.shape[3]
        dx = torch.empty_like(x)
        assert lib.ccdm_norm_bwd_apply(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), B * h * w, c, h * w, coef.data_ptr(),
                                       2 * ctot, coff, bcoef.data_ptr(), 3 * ctot, boff, act, None) == 0
        dxs.append(dx)
        coff, boff = coff + 2 * c, boff + 3 * c
    # ---- autograd reference on the same (bf16-rounded) inputs
    xr = torch.cat([x.float() for x in xs], -1).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr.permute(0, 3, 1, 2), groups, gr, br, eps=1e-5)
    if with_ss:
        sc = ss[:, SS_OFF:SS_OFF + ctot].clone().requires_grad_(True)
        sh = ss[:, SS_OFF + ctot:SS_OFF + 2 * ctot].clone().requires_grad_(True)
        y = y * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
    y = {0: lambda v: v, 1: F.relu, 2: F.silu}[act](y)
    y.backward(torch.cat([d.float() for d in dys], -1).permute(0, 3, 1, 2))
    assert rel(torch.cat([d.float() for d in dxs], -1), xr.grad) < 6e-3          # bf16 output rounding
    assert rel(dgamma - 0.5, gr.grad) < 1e-4 and rel(dbeta + 0.25, br.grad) < 1e-4
    if with_ss:
        assert rel(d_ss[:, :ctot], sc.grad) < 1e-4 and rel(d_ss[:, ctot:], sh.grad) < 1e-4


@pytest.mark.parametrize("dh,n,head_major", [(16, 16, 0), (16, 70, 1), (32, 40, 1), (64, 33, 0), (128, 64, 1), (128, 9, 0)])
def test_attention_tokens_backward_matches_autograd(lib, dh, n, head_major):
    lib.ccdm_attention_tokens_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]
    g = torch.Generator().manual_seed(3 * n + dh)
    B, heads = 2, 2
    hid = heads * dh
    qkv = torch.randn(B, n, 3 * hid, generator=g).to(torch.bfloat16).contiguous()
    dout = torch.randn(B, n, hid, generator=g).to(torch.bfloat16).contiguous()
    dqkv = torch.zeros_like(qkv)
    scale = 1.0 / math.sqrt(dh)
    assert lib.ccdm_attention_tokens_bwd(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, n, heads, dh, scale, head_major,
                                         None) == 0, lib.hostsim_last_error()
    f = qkv.float().requires_grad_(True)
    v5 = f.reshape(B, n, heads, 3, dh).permute(0, 1, 3, 2, 4) if head_major else f.reshape(B, n, 3, heads, dh)
    att = torch.einsum("bihd,bjhd->bhij", v5[:, :, 0] * scale, v5[:, :, 1]).softmax(-1)
    o = torch.einsum("bhij,bjhd->bihd", att, v5[:, :, 2]).reshape(B, n, hid)
    o.backward(dout.float())
    assert rel(dqkv, f.grad) < 6e-3          # bf16 output rounding
