"""CPU checks of the backward DERIVATIONS the training kernels implement (tests/emu_train.py restates the kernels'
arithmetic step by step) against torch.autograd of the oracle's formulas.  fp32 on both sides: errors ~1e-6."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.emu_train import block_bwd_emu, linattn_core_fwd_emu, linattn_core_bwd_emu, attention_small_bwd_emu


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("use_ss,silu", [(True, True), (False, True), (False, False)])
def test_block_tail_backward_formula(use_ss, silu):
    g0 = torch.Generator().manual_seed(1)
    B, P, C = 3, 10, 24
    z = torch.randn(B, P, C, generator=g0, requires_grad=True)
    gain = (1 + 0.1 * torch.randn(C, generator=g0)).requires_grad_(True)
    ss = (0.3 * torch.randn(B, 2 * C, generator=g0)).requires_grad_(True) if use_ss else None
    dy = torch.randn(B, P, C, generator=g0)
    n = F.normalize(z, dim=-1) * gain * math.sqrt(C)           # unet.py:88-89
    if use_ss:
        n = n * (1 + ss[:, None, :C]) + ss[:, None, C:]         # unet.py:147-149
    y = F.silu(n) if silu else n
    y.backward(dy)
    dz, d_ss, dgain, dbias = block_bwd_emu(dy, z.detach(), gain.detach(), math.sqrt(C), ss.detach() if use_ss else None, silu)
    assert rel(dz, z.grad) < 1e-5
    assert rel(dgain, gain.grad) < 1e-5
    assert rel(dbias, z.grad.sum((0, 1))) < 1e-5
    if use_ss:
        assert rel(d_ss, ss.grad) < 1e-5


@pytest.mark.parametrize("n", [16, 50])
def test_linear_attention_core_backward_formula(n):
    g0 = torch.Generator().manual_seed(2)
    B, scale = 2, 32 ** -0.5
    qkv = (torch.randn(B, n, 384, generator=g0) * 1.5).requires_grad_(True)
    dout = torch.randn(B, n, 128, generator=g0)
    # the oracle's formulation (oracle/unet_ref.py::_linear_attention, unet.py:204-214), token-major
    q, k, v = (qkv[..., i * 128:(i + 1) * 128].reshape(B, n, 4, 32) for i in range(3))
    qs = q.softmax(-1) * scale
    ks = k.softmax(1)
    ctx = torch.einsum("bnhd,bnhe->bhde", ks, v)
    ref = torch.einsum("bhde,bnhd->bnhe", ctx, qs).reshape(B, n, 128)
    ref.backward(dout)
    out, saved = linattn_core_fwd_emu(qkv.detach(), scale)
    assert rel(out, ref.detach()) < 1e-5
    dqkv = linattn_core_bwd_emu(dout, saved, scale)
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        assert rel(dqkv[..., sl], qkv.grad[..., sl]) < 1e-4, name


@pytest.mark.parametrize("heads,dh,n", [(4, 32, 16), (2, 16, 9)])
def test_bottleneck_attention_backward_formula(heads, dh, n):
    g0 = torch.Generator().manual_seed(3)
    B, hid, scale = 2, heads * dh, dh ** -0.5
    qkv = torch.randn(B, n, 3 * hid, generator=g0, requires_grad=True)
    dout = torch.randn(B, n, hid, generator=g0)
    q, k, v = (qkv[..., i * hid:(i + 1) * hid].reshape(B, n, heads, dh) for i in range(3))
    att = torch.einsum("bihd,bjhd->bhij", q * scale, k).softmax(-1)          # unet.py:231-236
    o = torch.einsum("bhij,bjhd->bihd", att, v).reshape(B, n, hid)
    o.backward(dout)
    got = attention_small_bwd_emu(qkv.detach(), dout, heads, dh, scale)
    assert rel(got, qkv.grad) < 1e-5
