"""Run single tap-GEMM layers of the headline workload in isolation (for ncu / timing).

    python tools/prof_layer.py [conv64|conv128|qkv|all] [--batch 400] [--iters 5]
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ccdm_b200 import _lib as L  # noqa: E402
from ccdm_b200.engine import Program, TapGemmRec, WeightStore, nhwc_view, tapgemm_flops  # noqa: E402
from ccdm_b200.plan import plan_conv, tile_box, n_tiling, can_reuse_rows  # noqa: E402


def build(kind, B, H, W, cins, cout, flags, dev, tile=None, n_tile=None, reuse=True, halo=False):
    xs = [torch.randn(B, H, W, c, device=dev).to(torch.bfloat16) for c in cins]
    k = {"1x1": 1, "3x3": 3}[kind]
    conv = torch.nn.Conv2d(sum(cins), cout, k).to(dev)
    ws, prog = WeightStore(dev), Program(dev)
    tile = tile or tile_box(W, H, square=(kind != "1x1"))
    plan = plan_conv(kind, cins, cout, reuse_rows=(reuse and kind != "1x1" and can_reuse_rows(tile)))
    if halo and kind == "3x3":
        from ccdm_b200.plan import HALO_TILE
        plan, tile = plan_conv(kind, cins, cout, halo=True), HALO_TILE
    n_rows, nt = n_tiling(cout, bool(flags & (L.EPI_RMSNORM | L.EPI_SUMSQ_OUT)))
    nt = n_tile or nt
    pack = ws.add("w", conv.weight, plan, n_rows)
    out = torch.zeros(B, H, W, cout, device=dev, dtype=torch.bfloat16)
    rec = TapGemmRec("t", plan, [nhwc_view(x) for x in xs], W, H, B, tile, pack, pack.packed, pack.sched, n_rows, cout,
                     nt, flags | L.EPI_BIAS, out, (cout, W * cout, H * W * cout), bias=conv.bias)
    if flags & L.EPI_RMSNORM:
        rec.gain, rec.gain_mul = torch.ones(cout, device=dev), math.sqrt(cout)
    if flags & L.EPI_SS:
        rec.ss, rec.ss_ld, rec.ss_off = torch.randn(B, 2 * cout, device=dev) * 0.1, 2 * cout, 0
    if flags & L.EPI_RESID:
        rec.resid, rec.resid_strides = torch.randn_like(out), (cout, W * cout, H * W * cout)
    if flags & L.EPI_ROWSCALE:
        rec.rowss = torch.rand(B * H * W, device=dev) + 0.5
    if flags & L.EPI_QSOFTMAX:
        rec.q_scale, rec.q_cols = 32 ** -0.5, 128
    prog.recs.append(rec)
    prog.finalize()
    s = torch.cuda.current_stream().cuda_stream
    ws.refresh(s)
    return prog, rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="all")
    ap.add_argument("--batch", type=int, default=400)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tile", default=None, help="tw,th,tb override")
    ap.add_argument("--no-reuse", action="store_true")
    ap.add_argument("--halo", action="store_true", help="3x3 layers read their nine taps out of one halo box per tile")
    a = ap.parse_args()
    tile = tuple(int(v) for v in a.tile.split(",")) if a.tile else None
    dev = torch.device("cuda")
    B = a.batch
    cases = {
        "conv64": ("3x3", B, 64, 64, (64,), 64, L.EPI_RMSNORM | L.EPI_SS | L.EPI_SILU),
        "conv64plain": ("3x3", B, 64, 64, (64,), 64, 0),
        "conv72plain": ("3x3", B, 64, 64, (72,), 72, 0),          # UTKFace widths: a full block + an 8-channel block
        "conv128": ("3x3", B, 64, 64, (64, 64), 64, L.EPI_RMSNORM | L.EPI_SS | L.EPI_SILU),
        "conv128x128": ("3x3", B, 32, 32, (128,), 128, L.EPI_RMSNORM | L.EPI_SILU | L.EPI_RESID),
        "qkv": ("1x1", B, 64, 64, (64,), 384, L.EPI_ROWSCALE | L.EPI_QSOFTMAX),
        "res1x1": ("1x1", B, 64, 64, (64, 64), 64, 0),
    }
    names = list(cases) if a.which == "all" else a.which.split(",")
    for nm in names:
        prog, rec = build(*cases[nm], dev, tile=tile, reuse=not a.no_reuse, halo=a.halo)
        s = torch.cuda.current_stream().cuda_stream
        prog.run(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            prog.run(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"{nm:12s} tile={rec.tile} R={rec.plan.R} n_tile={rec.n_tile}: {ms*1e3:8.1f} us  {tapgemm_flops(rec)/ms/1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
