"""Host planning logic (tile boxes, K schedules, packing rule) against torch convolutions -- CPU only."""
import pytest
import torch
import torch.nn.functional as F

from ccdm_b200.plan import plan_conv, tile_box, n_tiling, TILE_M
from tests.emu import tapgemm_emu, assemble_parity


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("g", [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (128, 128), (192, 192), (96, 96),
                               (48, 48), (24, 24), (12, 12), (6, 6), (3, 3), (2, 2), (1, 1)])
@pytest.mark.parametrize("tb1", [False, True])
def test_tile_box(g, tb1):
    tw, th, tb = tile_box(g[0], g[1], force_tb1=tb1)
    assert tw * th * tb == TILE_M and max(tw, th, tb) <= 256
    if tb1:
        assert tb == 1
    # exact cover for the power-of-two grids the shipped configs use
    if g[0] in (64, 32, 16, 8, 4) and not tb1:
        assert g[0] % tw == 0 and g[1] % th == 0


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("cins,cout", [((64,), 64), ((128, 64), 64), ((32,), 96), ((72,), 72), ((96, 32), 32)])
def test_conv3x3(cins, cout, reuse):
    torch.manual_seed(0)
    xs = [torch.randn(2, c, 6, 5) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3)
    ref = nhwc(F.conv2d(torch.cat(xs, 1), w, padding=1))
    plan = plan_conv("3x3", cins, cout, reuse_rows=reuse)
    assert plan.R == (3 if reuse else 1) and plan.nkb == 9 * sum((c + 63) // 64 for c in cins)
    out = tapgemm_emu(plan, [nhwc(x) for x in xs], w, 6, 5)[0]
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("cins,cout", [((64,), 384), ((128, 64), 64), ((40,), 8)])
def test_conv1x1(cins, cout):
    torch.manual_seed(1)
    xs = [torch.randn(2, c, 4, 4) for c in cins]
    w = torch.randn(cout, sum(cins), 1, 1)
    ref = nhwc(F.conv2d(torch.cat(xs, 1), w))
    g = torch.rand(sum(cins)) + 0.5
    out = tapgemm_emu(plan_conv("1x1", cins, cout), [nhwc(x) for x in xs], w, 4, 4)[0]
    assert rel(out, ref) < 1e-5
    # cin_gain fold == scaling the input channels
    ref_g = nhwc(F.conv2d(torch.cat(xs, 1) * g[None, :, None, None], w))
    out_g = tapgemm_emu(plan_conv("1x1", cins, cout), [nhwc(x) for x in xs], w, 4, 4, cin_gain=g)[0]
    assert rel(out_g, ref_g) < 1e-5


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("c,cout,hw", [(64, 128, (8, 8)), (32, 32, (6, 4)), (72, 144, (4, 4))])
def test_down4x4s2(c, cout, hw, reuse):
    torch.manual_seed(2)
    x = torch.randn(2, c, *hw)
    w = torch.randn(cout, c, 4, 4)
    ref = nhwc(F.conv2d(x, w, stride=2, padding=1))
    plan = plan_conv("down4x4s2", (c,), cout, reuse_rows=reuse)
    assert plan.R == (2 if reuse else 1)
    out = tapgemm_emu(plan, [nhwc(x)], w, hw[0] // 2, hw[1] // 2)[0]
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("cins,cout,hw", [((64,), 64, (8, 8)), ((32,), 48, (6, 4)), ((128,), 128, (16, 16)),
                                          ((72, 24), 40, (4, 4))])
def test_down3x3s2(cins, cout, hw, reuse):
    """The vanilla UNet's Downsample: 3x3 / stride 2 / padding 1 over four parity planes."""
    torch.manual_seed(12)
    xs = [torch.randn(2, c, *hw) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3)
    ref = nhwc(F.conv2d(torch.cat(xs, 1), w, stride=2, padding=1))
    plan = plan_conv("down3x3s2", cins, cout, reuse_rows=reuse)
    assert plan.R == (2 if reuse else 1) and plan.n_views == 4
    out = tapgemm_emu(plan, [nhwc(x) for x in xs], w, hw[0] // 2, hw[1] // 2)[0]
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("c,cout,hw", [(64, 64, (4, 4)), (128, 64, (3, 5)), (32, 32, (1, 1))])
def test_up2x3x3(c, cout, hw, reuse):
    torch.manual_seed(3)
    x = torch.randn(2, c, *hw)
    w = torch.randn(cout, c, 3, 3)
    ref = nhwc(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1))
    plan = plan_conv("up2x3x3", (c,), cout, reuse_rows=reuse)
    assert plan.R == (2 if reuse else 1) and plan.nz == 4
    out4 = tapgemm_emu(plan, [nhwc(x)], w, hw[0], hw[1])
    assert rel(assemble_parity(out4), ref) < 1e-5


def test_n_tiling():
    assert n_tiling(64, True) == (64, 64)
    assert n_tiling(512, True) == (512, 512)
    assert n_tiling(72, True) == (96, 96)
    assert n_tiling(384, False) == (384, 128)
    assert n_tiling(9000, False)[1] == 128 and n_tiling(9000, False)[0] % 128 == 0
    with pytest.raises(ValueError):
        n_tiling(576, True)


# ----------------------------------------------------------------------------- data-gradient plans vs autograd

def _dgrad_ref(fn, x, w):
    x = x.clone().requires_grad_(True)
    y = fn(x, w)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
    y.backward(dy)
    return dy, x.grad


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("kind,cin,cout,hw", [("3x3", 64, 32, (6, 5)), ("3x3", 96, 64, (4, 4)), ("1x1", 40, 72, (3, 3)),
                                             ("down4x4s2", 32, 64, (8, 6)), ("up2x3x3", 64, 32, (3, 4)),
                                             ("down3x3s2", 32, 64, (8, 6)), ("down3x3s2", 72, 40, (4, 4))])
def test_dgrad_plans(kind, cin, cout, hw, reuse):
    torch.manual_seed(5)
    x = torch.randn(2, cin, *hw)
    k = {"3x3": 3, "1x1": 1, "down4x4s2": 4, "up2x3x3": 3, "down3x3s2": 3}[kind]
    w = torch.randn(cout, cin, k, k)
    fwd = {"3x3": lambda a, b: F.conv2d(a, b, padding=1), "1x1": lambda a, b: F.conv2d(a, b),
           "down4x4s2": lambda a, b: F.conv2d(a, b, stride=2, padding=1),
           "down3x3s2": lambda a, b: F.conv2d(a, b, stride=2, padding=1),
           "up2x3x3": lambda a, b: F.conv2d(F.interpolate(a, scale_factor=2, mode="nearest"), b, padding=1)}[kind]
    dy, dx = _dgrad_ref(fwd, x, w)
    plan = plan_conv(kind + "_dgrad", (cout,), cin, reuse_rows=reuse)
    assert plan.transposed
    if kind in ("down4x4s2", "down3x3s2"):        # four output parity planes of dx, each the size of dy
        out4 = tapgemm_emu(plan, [nhwc(dy)], w, dy.shape[2], dy.shape[3])
        got = assemble_parity(out4)
    else:
        got = tapgemm_emu(plan, [nhwc(dy)], w, hw[0], hw[1])[0]
    assert rel(got, nhwc(dx)) < 1e-5
    # concatenated forward input: gradient of the second source only = output-channel window of the same plan
    if kind == "3x3":
        half = cin // 2
        got2 = tapgemm_emu(plan_conv("3x3_dgrad", (cout,), cin - half, reuse_rows=reuse), [nhwc(dy)], w, hw[0], hw[1],
                           n_off=half, n_count=cin - half)[0]
        assert rel(got2, nhwc(dx)[..., half:]) < 1e-5


@pytest.mark.parametrize("cins,cout", [((64,), 64), ((64, 64), 64), ((72,), 72), ((40, 24), 32)])
def test_halo_plan_equals_conv(cins, cout):
    """3x3 with ONE halo box per 64-channel block (plan.halo, R = 9: K block g*9 + r*3 + q = tap (r, q), box origin (-1, -1)):
    the tap-GEMM semantics reproduce F.conv2d(padding=1), exactly like the three-box plan."""
    import torch.nn.functional as F
    from ccdm_b200.plan import plan_conv, halo_ok, HALO_TILE
    from tests.emu import tapgemm_emu
    g = torch.Generator().manual_seed(31)
    B, H, W = 2, 16, 8
    assert halo_ok("3x3", W, H) and not halo_ok("3x3", 12, 16) and not halo_ok("1x1", W, H) and HALO_TILE == (8, 16, 1)
    srcs = [torch.randn(B, H, W, c, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3, generator=g)
    plan = plan_conv("3x3", cins, cout, halo=True)
    assert plan.halo and plan.R == 9 and plan.ngroups == sum(-(-c // 64) for c in cins) and plan.nkb == 9 * plan.ngroups
    assert all(e[1] == -1 and e[2] == -1 for e in plan.sched)
    got = tapgemm_emu(plan, srcs, w, H, W)[0]
    want = F.conv2d(torch.cat(srcs, -1).permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    assert (got - want).abs().max() < 1e-3 * want.abs().max()


@pytest.mark.parametrize("cin,cout", [(64, 64), (72, 72), (40, 96)])
def test_halo_dgrad_plan_equals_autograd(cin, cout):
    """Data gradient of a 3x3 conv on halo boxes (plan_conv('3x3_dgrad', halo=True): one box per 64-channel block of dy,
    flipped filter taps), incl. the output-channel window of a concatenated forward input."""
    torch.manual_seed(9)
    H, W = 16, 8
    x = torch.randn(2, cin, H, W)
    w = torch.randn(cout, cin, 3, 3)
    dy, dx = _dgrad_ref(lambda a, b: F.conv2d(a, b, padding=1), x, w)
    plan = plan_conv("3x3_dgrad", (cout,), cin, halo=True)
    assert plan.halo and plan.transposed and plan.R == 9 and plan.ngroups == -(-cout // 64)
    got = tapgemm_emu(plan, [nhwc(dy)], w, H, W)[0]
    assert rel(got, nhwc(dx)) < 1e-5
    half = cin // 2
    got2 = tapgemm_emu(plan_conv("3x3_dgrad", (cout,), cin - half, halo=True), [nhwc(dy)], w, H, W, n_off=half,
                       n_count=cin - half)[0]
    assert rel(got2, nhwc(dx)[..., half:]) < 1e-5
