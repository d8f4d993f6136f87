"""One-step generator: host-side layout on CPU, CUDA path vs the oracle on the GPU.
Tolerance: bf16 activations through 9-11 convolutions -> relative error <= 2e-2 on the tanh output, PSNR >= 40 dB."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle.sngan_ref import GenSpec, generator_forward, make_state_dict, state_dict_shapes
from tests.golden.sngan_cases import GEN_CASES, GEN_SPECS, gen_inputs
from tests.emu import tapgemm_emu, assemble_parity
from tests.test_plan import nhwc, rel
from ccdm_b200.plan import plan_conv

G = os.path.join(os.path.dirname(__file__), "golden")


def make_gen(spec):
    import ccdm_b200.sngan as S
    return S.sngan_generator(dim_z=spec.dim_z, dim_embed=spec.dim_embed, nc=spec.nc, img_size=spec.img_size,
                             gene_ch=spec.gene_ch)


@pytest.mark.parametrize("name", list(GEN_SPECS))
def test_state_dict_layout_matches_reference(name):
    """Same keys, order and shapes as the reference module (recorded in the golden file), so checkpoints load."""
    spec = GEN_SPECS[name]
    net = make_gen(spec)
    sd = net.state_dict()
    gold = torch.load(os.path.join(G, "sngan.pt"), weights_only=True)[name]
    assert list(sd.keys()) == gold["keys"]
    shapes = state_dict_shapes(spec)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in shapes.items()}
    net.load_state_dict(make_state_dict(spec, 1), strict=True)


def test_up2x1x1_plan_is_upsample_then_1x1():
    torch.manual_seed(0)
    x = torch.randn(2, 40, 5, 6)
    w = torch.randn(24, 40, 1, 1)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w)
    plan = plan_conv("up2x1x1", (40,), 24)
    got = assemble_parity(tapgemm_emu(plan, [nhwc(x)], w, 5, 6))
    assert rel(got, nhwc(ref)) < 1e-5


def test_cpu_call_fails_loudly():
    net = make_gen(GEN_SPECS["g64"]).eval()
    with pytest.raises(RuntimeError):
        net(torch.randn(2, 32), torch.rand(2, 16))
    with pytest.raises(RuntimeError):
        net.train()(torch.randn(2, 32), torch.rand(2, 16))


def test_oracle_train_mode_vs_reference_golden():
    """Training-mode forward (batch-statistics BatchNorm2d, sngan.py:19-36): the oracle restatement reproduces the reference
    module's output and its updated running statistics (tests/golden/sngan.pt, case g64_train)."""
    spec = GEN_SPECS["g64"]
    sd = make_state_dict(spec, GEN_CASES["g64"][1])
    z, y = gen_inputs(spec, 6, seed=51)
    gold = torch.load(os.path.join(G, "sngan.pt"), weights_only=True)["g64_train"]
    stats = {}
    with torch.no_grad():
        out = generator_forward(sd, spec, z, y, stats=stats)
    assert ((out - gold["out"]).norm() / gold["out"].norm()).item() < 1e-5
    for k, v in stats.items():
        assert ((v - gold["stats"][k]).norm() / gold["stats"][k].norm().clamp_min(1e-12)).item() < 1e-5, k


@pytest.mark.gpu
def test_generator_train_mode_vs_reference_golden():
    """The CUDA path in training mode: ccdm_channel_stats batch statistics -> ccdm_condbn_coef -> affine + ReLU, running
    statistics updated like nn.BatchNorm2d; against the reference module's own train-mode output."""
    spec = GEN_SPECS["g64"]
    sd = make_state_dict(spec, GEN_CASES["g64"][1])
    net = make_gen(spec)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    z, y = gen_inputs(spec, 6, seed=51)
    out = net(z.cuda(), y.cuda()).cpu()
    gold = torch.load(os.path.join(G, "sngan.pt"), weights_only=True)["g64_train"]
    e = ((out - gold["out"]).norm() / gold["out"].norm()).item()
    print(f"g64 train mode: rel err vs reference {e:.3e}")
    assert e < 2e-2
    got = net.state_dict()
    for k, v in gold["stats"].items():
        if "num_batches" in k:
            assert int(got[k]) == int(v), k
        else:
            assert ((got[k].cpu() - v).norm() / v.norm().clamp_min(1e-12)).item() < 2e-2, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GEN_CASES))
def test_generator_vs_oracle_and_golden(name):
    sname, seed, batch = GEN_CASES[name]
    spec = GEN_SPECS[sname]
    sd = make_state_dict(spec, seed)
    net = make_gen(spec)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    z, y = gen_inputs(spec, batch)
    out = net(z.cuda(), y.cuda()).cpu()
    gold = torch.load(os.path.join(G, "sngan.pt"), weights_only=True)[name]["out"]
    ref = generator_forward(sd, spec, z, y)
    e_gold = ((out - gold).norm() / gold.norm()).item()
    e_ref = ((out - ref).norm() / ref.norm()).item()
    mse = ((out - gold).double() ** 2).mean().item()
    psnr = 10 * math.log10(4.0 / max(mse, 1e-20))          # images span [-1, 1]
    print(f"{name}: rel err vs reference golden {e_gold:.3e}, vs oracle {e_ref:.3e}, PSNR {psnr:.1f} dB")
    assert out.shape == gold.shape
    assert e_gold < 2e-2 and e_ref < 2e-2
    assert psnr >= 40.0


@pytest.mark.gpu
def test_generator_full_config_vs_oracle():
    """SURVEY config 5: dim_z 256, dim_embed 128, 192 px, gene_ch 48 (channels 768..48), oracle on the same GPU."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    spec = GenSpec(dim_z=256, dim_embed=128, nc=3, img_size=192, gene_ch=48)
    sd = make_state_dict(spec, 4)
    net = make_gen(spec)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(9)
    z = torch.randn(3, 256, generator=g).cuda()
    y = torch.rand(3, 128, generator=g).cuda()
    out = net(z, y)
    ref = generator_forward({k: v.cuda() for k, v in sd.items()}, spec, z, y)
    e = ((out - ref).norm() / ref.norm()).item()
    print(f"192 px generator: rel err {e:.3e}")
    assert e < 2e-2
    # batch invariance (eval-mode BatchNorm): sample 1 alone == sample 1 in the batch
    solo = net(z[1:2], y[1:2])
    assert ((solo - out[1:2]).norm() / out[1:2].norm()).item() < 2e-3
