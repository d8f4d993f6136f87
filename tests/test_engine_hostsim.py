"""Whole engine programs on CPU with every CUDA-core kernel REAL: the program's records are executed through the engine's own
argument marshalling (ccdm_b200.engine._make_call) against host builds of the kernels' source (tests/hostsim); only the
tcgen05 kernels (ccdm_tapgemm, ccdm_linattn_context) stay restated in torch (tests/emu_engine.py).  Checks the marshalling
of every non-GEMM call and the kernels themselves in the context of a full UNet forward, against the oracle."""
import ctypes as C

import pytest
import torch

from ccdm_b200 import _lib as L
from ccdm_b200.engine import KernelRec, PackRec, TapGemmRec, UnetProgram, WeightStore, _make_call
from tests.emu_engine import run_kernel, run_tapgemm
from tests.hostsim.build import build, build_extract

EMULATED = {"linattn_context", "linattn_kv_partials", "linattn_q_out"}          # tcgen05


class HostLib:
    def __init__(self):
        handles = [C.CDLL(build("kernels.cu")), C.CDLL(build("train_kernels.cu")), C.CDLL(build("groupnorm.cu")),
                   C.CDLL(build_extract("tapgemm.cu", ["pack_weights_kernel"], ["ccdm_pack_weights_at", "ccdm_pack_weights"])),
                   C.CDLL(build_extract("linattn.cu", ["kexp_bound_kernel"], ["ccdm_kexp_bound"])),
                   C.CDLL(build_extract("linattn_fused.cu", ["linattn_fold_parts_kernel"], ["ccdm_linattn_fold_partials"]))]
        for name, (res, args) in L.SIGNATURES.items():
            for h in handles:
                fn = getattr(h, name, None)
                if fn is not None:
                    fn.restype, fn.argtypes = res, args
                    setattr(self, name, fn)
                    break
            else:
                setattr(self, name, None)                    # tcgen05 entry points: never called here


def run_program_hostsim(prog, weights, lib):
    keep = []
    with torch.no_grad():
        if hasattr(prog, "glue_in"):
            prog.glue_in()
        for r in weights.program.recs:                       # weight packs, kexp bounds, stem pack
            fn, args, name = _make_call(lib, r, keep)
            assert fn(*args, None) == 0, name
        for off, b in prog._tc_bias_srcs:
            weights.tc_bias[off:off + b.numel()].copy_(b.detach())
        if hasattr(prog, "_head_bias_src"):
            weights.head_bias[: prog._head_bias_src.numel()].copy_(prog._head_bias_src.detach())
        for r in prog.recs:
            if isinstance(r, TapGemmRec):
                run_tapgemm(r)
            elif isinstance(r, KernelRec) and r.kind in EMULATED:
                run_kernel(r)
            else:
                fn, args, name = _make_call(lib, r, keep)
                assert fn(*args, None) == 0, name
        if hasattr(prog, "glue_out"):
            prog.glue_out()
    return prog.out


@pytest.fixture(scope="module")
def lib():
    return HostLib()


@pytest.mark.parametrize("spec_name,mode", [("tiny", "eval"), ("cell", "train")])
def test_unified_program_with_real_kernels(lib, spec_name, mode):
    from oracle.unet_ref import unet_forward
    from tests.golden.cases import BATCH, SIZES, unet_inputs
    from tests.test_engine_emulated import build as build_net
    spec, net, sd = build_net(spec_name, 5)
    net.train(mode == "train")
    x, t, emb = unet_inputs(spec_name)
    B, size = BATCH[spec_name], SIZES[spec_name]
    ws = WeightStore(torch.device("cpu"))
    prog = UnetProgram(net, ws, B, B, size, size, mode == "train")
    keep = torch.tensor([True, False, True, False, True][:B])
    prog.x_in.copy_(x); prog.t_in.copy_(t); prog.emb_in.copy_(emb); prog.keep.copy_(keep.to(torch.uint8))
    out = run_program_hostsim(prog, ws, lib)
    with torch.no_grad():
        ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.5, training=(mode == "train"), keep_mask=keep)
    err = ((out - ref).norm() / ref.norm()).item()
    print(f"{spec_name}/{mode}: rel L2 err {err:.3e} (every CUDA-core kernel real)")
    assert err < 2e-2


def test_vanilla_program_with_real_kernels(lib):
    from ccdm_b200.vanilla_unet import VanillaProgram
    from oracle.vanilla_unet_ref import make_state_dict, vanilla_unet_forward
    from tests.golden.vanilla_cases import V_BATCH, V_SIZES, keep_mask, vanilla_inputs
    from tests.test_vanilla_emulated import build as build_net
    spec, net = build_net("v_tiny", 1)
    net.eval()
    x, t, classes = vanilla_inputs("v_tiny")
    B, size = V_BATCH["v_tiny"], V_SIZES["v_tiny"]
    ws = WeightStore(torch.device("cpu"))
    prog = VanillaProgram(net, ws, B, B, size, size, False)
    keep = keep_mask("mixed", B)
    prog.load_inputs(x, t, classes, keep.to(torch.uint8))
    out = run_program_hostsim(prog, ws, lib)
    with torch.no_grad():
        ref = vanilla_unet_forward(make_state_dict(spec, 1), spec, x, t, classes, keep)
    err = ((out - ref).norm() / ref.norm()).item()
    print(f"v_tiny: rel L2 err {err:.3e} (every CUDA-core kernel real)")
    assert err < 2e-2
