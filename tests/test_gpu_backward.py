"""GPU parity of the backward building blocks (ccdm_b200/backward.py) against torch autograd in fp32 on the same
bf16-rounded inputs: data gradient (tap-GEMM over dY), weight gradient (tcgen05 split-K kernel) and the Block tail.
Tolerances: gradients that leave as bf16 are compared at 1e-2 relative Frobenius error (bf16 has 8 bits of mantissa:
~4e-3 rounding), fp32 weight gradients at 2e-3 (fp32 accumulation order only)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

FWD = {
    "1x1": lambda x, w: F.conv2d(x, w),
    "3x3": lambda x, w: F.conv2d(x, w, padding=1),
    "down4x4s2": lambda x, w: F.conv2d(x, w, stride=2, padding=1),
    "up2x3x3": lambda x, w: F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1),
}
KSIZE = {"1x1": 1, "3x3": 3, "down4x4s2": 4, "up2x3x3": 3}

# kind, cins, cout, B, H, W
CASES = [
    ("3x3", (64,), 64, 3, 32, 32),          # R = 3 vertical reuse, 16x8 boxes
    ("3x3", (64, 64), 64, 2, 16, 16),       # concatenated skip input (decoder blocks)
    ("3x3", (128,), 256, 4, 8, 8),          # tb = 2 boxes, no reuse, two 128-row output tiles
    ("3x3", (72,), 144, 2, 16, 16),         # dim-72 widths: ragged channel slices
    ("1x1", (128,), 64, 2, 16, 16),
    ("1x1", (64, 128), 128, 2, 8, 8),       # res_conv of a concat block
    ("down4x4s2", (64,), 128, 2, 32, 32),
    ("up2x3x3", (128,), 64, 2, 8, 8),
    ("up2x3x3", (64,), 64, 2, 16, 16),
]


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def make_case(kind, cins, cout, b, h, w, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [torch.randn(b, c, h, w, generator=g).bfloat16().float().cuda() for c in cins]
    k = KSIZE[kind]
    wt = (torch.randn(cout, sum(cins), k, k, generator=g) / math.sqrt(sum(cins) * k * k)).bfloat16().float().cuda()
    x = torch.cat(xs, 1).requires_grad_(True)
    wp = wt.clone().requires_grad_(True)
    y = FWD[kind](x, wp)
    dy = torch.randn(y.shape, generator=g).bfloat16().float().cuda()
    y.backward(dy)
    return xs, wt, y.detach(), dy, x.grad, wp.grad


@pytest.mark.parametrize("kind,cins,cout,b,h,w", CASES)
def test_conv_forward_eager(kind, cins, cout, b, h, w):
    from ccdm_b200 import backward as bw
    xs, wt, y, _, _, _ = make_case(kind, cins, cout, b, h, w)
    bias = torch.linspace(-1, 1, cout, device="cuda")
    got = bw.conv_forward(kind, [nhwc(x).bfloat16() for x in xs], wt, bias)
    assert rel(got, nhwc(y) + bias) < 1e-2


@pytest.mark.parametrize("kind,cins,cout,b,h,w", CASES)
def test_conv_dgrad(kind, cins, cout, b, h, w):
    from ccdm_b200 import backward as bw
    xs, wt, _, dy, dx, _ = make_case(kind, cins, cout, b, h, w, seed=1)
    got = bw.conv_dgrad(kind, nhwc(dy).bfloat16(), wt, cins)
    off = 0
    for g, c in zip(got, cins):
        assert g.shape == (b, h, w, c)
        assert rel(g, nhwc(dx[:, off:off + c])) < 1e-2, (kind, off)
        off += c


@pytest.mark.parametrize("kind,cins,cout,b,h,w", CASES)
def test_conv_wgrad(kind, cins, cout, b, h, w):
    from ccdm_b200 import backward as bw
    xs, wt, _, dy, _, dw = make_case(kind, cins, cout, b, h, w, seed=2)
    got = bw.conv_wgrad(kind, [nhwc(x).bfloat16() for x in xs], nhwc(dy).bfloat16())
    assert got.shape == dw.shape
    assert rel(got, dw) < 2e-3
    # split-K invariance: one slice per unit vs many slices must agree to fp32 accumulation noise
    one = bw.conv_wgrad(kind, [nhwc(x).bfloat16() for x in xs], nhwc(dy).bfloat16(), ksplit=1)
    assert rel(one, dw) < 2e-3


def test_conv_wgrad_full_size_linearity():
    """RC-49 level-0 shape at a training batch: wgrad(a*dz) == a*wgrad(dz) and agreement with autograd on a slice."""
    from ccdm_b200 import backward as bw
    torch.manual_seed(3)
    x = torch.randn(32, 64, 64, 64, device="cuda").bfloat16()
    dz = torch.randn(32, 64, 64, 64, device="cuda").bfloat16()
    a = bw.conv_wgrad("3x3", [x], dz)
    b2 = bw.conv_wgrad("3x3", [x], (dz.float() * 2).bfloat16())
    assert rel(b2, 2 * a) < 1e-5
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros(64, 64, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(xr, w, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    assert rel(a, w.grad) < 2e-3


@pytest.mark.parametrize("c,hw,b,use_ss,silu", [(64, 16, 3, True, True), (128, 8, 2, False, True), (64, 8, 2, False, False),
                                                (576, 4, 2, True, True), (288, 8, 2, True, True)])
def test_block_backward(c, hw, b, use_ss, silu):
    from ccdm_b200 import backward as bw
    g0 = torch.Generator().manual_seed(4)
    z = torch.randn(b, hw, hw, c, generator=g0).bfloat16().cuda()
    dy = torch.randn(b, hw, hw, c, generator=g0).bfloat16().cuda()
    gain = (1 + 0.1 * torch.randn(c, generator=g0)).cuda()
    ld, off = 2 * c + 16, 8
    ss = (0.3 * torch.randn(b, ld, generator=g0)).cuda() if use_ss else None

    zr = z.float().requires_grad_(True)
    gr = gain.clone().requires_grad_(True)
    ssr = ss.clone().requires_grad_(True) if use_ss else None
    n = F.normalize(zr, dim=-1) * gr * math.sqrt(c)
    if use_ss:
        n = n * (1 + ssr[:, None, None, off:off + c]) + ssr[:, None, None, off + c:off + 2 * c]
    y = F.silu(n) if silu else n
    y.backward(dy.float())

    dz, d_ss, dgain, dbias = bw.block_backward(dy, z, gain, ss, off, silu)
    assert rel(dz, zr.grad) < 1e-2
    assert rel(dgain, gr.grad) < 2e-3
    assert rel(dbias, zr.grad.sum((0, 1, 2))) < 5e-3
    if use_ss:
        assert rel(d_ss[:, off:off + 2 * c], ssr.grad[:, off:off + 2 * c]) < 2e-3
        assert d_ss[:, :off].abs().max() == 0


def test_resblock_chain_backward():
    """conv3x3 -> Block tail -> conv3x3 chained through the CUDA blocks vs autograd (gradient w.r.t. the input and both
    weights): the composition the training program will be made of."""
    from ccdm_b200 import backward as bw
    g0 = torch.Generator().manual_seed(5)
    b, hw, c = 2, 16, 64
    x = torch.randn(b, hw, hw, c, generator=g0).bfloat16().cuda()
    w1 = (torch.randn(c, c, 3, 3, generator=g0) / 24).cuda()
    w2 = (torch.randn(c, c, 3, 3, generator=g0) / 24).cuda()
    b1 = (0.1 * torch.randn(c, generator=g0)).cuda()
    gain = torch.ones(c).cuda()
    ss = (0.2 * torch.randn(b, 2 * c, generator=g0)).cuda()
    dy = torch.randn(b, hw, hw, c, generator=g0).bfloat16().cuda()

    # CUDA path
    z1 = bw.conv_forward("3x3", [x], w1, b1)
    h1 = F.silu(F.normalize(z1.float(), dim=-1) * gain * 8 * (1 + ss[:, None, None, :c]) + ss[:, None, None, c:]).bfloat16()
    dh1, = bw.conv_dgrad("3x3", dy, w2, (c,))
    dw2 = bw.conv_wgrad("3x3", [h1], dy)
    dz1, d_ss, dgain, dbias = bw.block_backward(dh1, z1, gain, ss, 0, True)
    dx, = bw.conv_dgrad("3x3", dz1, w1, (c,))
    dw1 = bw.conv_wgrad("3x3", [x], dz1)

    # autograd in fp32 (weights rounded to bf16 as the packed copies are)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    w1r = w1.bfloat16().float().requires_grad_(True)
    w2r = w2.bfloat16().float().requires_grad_(True)
    b1r = b1.clone().requires_grad_(True)
    z = F.conv2d(xr, w1r, b1r, padding=1)
    n = F.normalize(z, dim=1) * 8 * (1 + ss[:, :c, None, None]) + ss[:, c:, None, None]
    y = F.conv2d(F.silu(n), w2r, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-2
    assert rel(dw2, w2r.grad) < 1e-2
    assert rel(dw1, w1r.grad) < 2e-2
    assert rel(dbias, b1r.grad) < 2e-2
