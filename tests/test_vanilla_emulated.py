"""Whole-program wiring check of the vanilla (GroupNorm) UNet on CPU: interpret the program's records (documented
kernel semantics: bf16 storage, fp32 accumulation) and compare with the oracle AND with the reference's own outputs."""
import os

import pytest
import torch

from ccdm_b200.engine import WeightStore
from ccdm_b200.vanilla_unet import VanillaProgram, VanillaUnet
from oracle.vanilla_unet_ref import make_state_dict, state_dict_shapes, vanilla_unet_forward
from tests.emu_engine import run_program
from tests.golden.vanilla_cases import V_BATCH, V_CASES, V_SIZES, V_SPECS, keep_mask, vanilla_inputs

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vanilla_unet.pt"))


def build(sname, seed):
    s = V_SPECS[sname]
    net = VanillaUnet(embed_input_dim=s.embed_input_dim, cond_drop_prob=0.5, in_channels=s.in_channels,
                      model_channels=s.model_channels, num_res_blocks=s.num_res_blocks,
                      attention_resolutions=s.attention_resolutions, channel_mult=s.channel_mult, num_heads=s.num_heads,
                      num_groups=s.num_groups)
    net.load_state_dict(make_state_dict(s, seed), strict=True)       # reference checkpoint layout
    return s, net


@pytest.mark.parametrize("name", list(V_CASES))
def test_program_matches_oracle_and_reference(name):
    sname, seed, mode, kind = V_CASES[name]
    spec, net = build(sname, seed)
    net.train(mode == "train")
    x, t, classes = vanilla_inputs(sname)
    B, size = V_BATCH[sname], V_SIZES[sname]
    ws = WeightStore(torch.device("cpu"))
    prog = VanillaProgram(net, ws, B, B, size, size, mode == "train")
    keep = keep_mask(kind, B)
    prog.load_inputs(x, t, classes, keep.to(torch.uint8))
    out = run_program(prog, ws)
    with torch.no_grad():
        ref = vanilla_unet_forward(make_state_dict(spec, seed), spec, x, t, classes, keep, training=(mode == "train"))
    err = ((out - ref).norm() / ref.norm()).item()
    gold = GOLD[name]["out"]
    err_gold = ((out - gold).norm() / gold.norm()).item()
    print(f"{name}: rel L2 err vs oracle {err:.3e}, vs reference output {err_gold:.3e} (bf16 storage)")
    assert err < 2e-2 and err_gold < 2e-2


def test_state_dict_layout_matches_reference():
    for sname, spec in V_SPECS.items():
        _, net = build(sname, 0)
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == state_dict_shapes(spec), sname
        assert list(got) == list(state_dict_shapes(spec)), sname    # registration order == the reference's (pinned in golden)


def test_no_cpu_fallback():
    _, net = build("v_tiny", 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.eval()(torch.zeros(2, 3, 16, 16), torch.zeros(2, dtype=torch.long), torch.zeros(2, 16), cond_drop_prob=0.0)


def test_pair_batch_shares_the_input():
    """2B program with x_batch = B (what the guided sampler runs): halves == conditional / unconditional forwards."""
    spec, net = build("v_tiny", 1)
    net.eval()
    x, t, classes = vanilla_inputs("v_tiny")
    B, size = V_BATCH["v_tiny"], V_SIZES["v_tiny"]
    ws = WeightStore(torch.device("cpu"))
    prog = VanillaProgram(net, ws, 2 * B, B, size, size, False)
    keep = torch.cat([torch.ones(B), torch.zeros(B)]).to(torch.uint8)
    prog.load_inputs(x, torch.cat([t, t]), torch.cat([classes, classes]), keep)
    out = run_program(prog, ws)
    sd = make_state_dict(spec, 1)
    with torch.no_grad():
        c = vanilla_unet_forward(sd, spec, x, t, classes, torch.ones(B, dtype=torch.bool))
        n = vanilla_unet_forward(sd, spec, x, t, classes, torch.zeros(B, dtype=torch.bool))
    assert ((out[:B] - c).norm() / c.norm()).item() < 2e-2
    assert ((out[B:] - n).norm() / n.norm()).item() < 2e-2
