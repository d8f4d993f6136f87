"""GPU bring-up diagnostics (developer tool, not part of the product path).

    python tools/gpu_diag.py            # runs every case in its own subprocess with a timeout
    python tools/gpu_diag.py prog:tiny  # one case in-process

`prog:<spec>` runs a UnetProgram on the GPU and compares EVERY intermediate buffer with the CPU interpreter of
the same program (tests/emu_engine.py), printing the first ops that diverge.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["gemm:basic", "gemm:conv", "prog:rc_small", "prog:tiny", "prog:cell", "misc"]


def rel(a, b):
    import torch
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item(), (a - b).abs().max().item()


def case_gemm(which):
    import torch
    import torch.nn.functional as F
    from ccdm_b200 import _lib as L
    from ccdm_b200.engine import Program, TapGemmRec, WeightStore, nhwc_view, parity_views
    from ccdm_b200.plan import plan_conv, tile_box, n_tiling, can_reuse_rows
    dev = torch.device("cuda")
    torch.manual_seed(0)
    shapes = []
    if which == "basic":
        # (kind, B, H, W, cins, cout)
        shapes = [("1x1", 1, 2, 64, (64,), 64), ("1x1", 2, 16, 16, (128,), 128), ("1x1", 2, 8, 8, (64,), 384),
                  ("1x1", 1, 16, 16, (256,), 512), ("1x1", 3, 4, 4, (64,), 32)]
    else:
        shapes = [("3x3", 2, 16, 16, (64,), 64), ("3x3", 2, 8, 8, (128, 64), 128), ("3x3", 1, 64, 64, (64,), 64),
                  ("3x3", 3, 4, 4, (32,), 96), ("down4x4s2", 2, 16, 16, (64,), 128), ("up2x3x3", 2, 8, 8, (64,), 32),
                  ("3x3", 2, 6, 6, (72,), 72), ("3x3", 5, 3, 3, (512,), 512)]
    for kind, B, H, W, cins, cout in shapes:
        xs = [torch.randn(B, H, W, c, device=dev).to(torch.bfloat16) for c in cins]
        k = {"1x1": 1, "3x3": 3, "down4x4s2": 4, "up2x3x3": 3}[kind]
        conv = torch.nn.Conv2d(sum(cins), cout, k).to(dev)
        ws = WeightStore(dev)
        prog = Program(dev)
        oh, ow = (H // 2, W // 2) if kind == "down4x4s2" else ((2 * H, 2 * W) if kind == "up2x3x3" else (H, W))
        gh, gw = (H, W) if kind == "up2x3x3" else (oh, ow)
        tile = tile_box(gw, gh, square=(kind != "1x1"))
        plan = plan_conv(kind, cins, cout, reuse_rows=(kind != "1x1" and can_reuse_rows(tile)))
        n_rows, n_tile = n_tiling(cout, False)
        pack = ws.add("w", conv.weight, plan, n_rows)
        out = torch.zeros(B, oh, ow, cout, device=dev, dtype=torch.bfloat16)
        views = []
        for x in xs:
            views += parity_views(x) if plan.n_views == 4 else [nhwc_view(x)]
        if plan.out_parity:
            ostr = (2 * cout, 2 * ow * cout, oh * ow * cout)
            ooff = tuple((pa * ow + pb) * cout for pa in range(2) for pb in range(2))
        else:
            ostr, ooff = (cout, ow * cout, oh * ow * cout), (0, 0, 0, 0)
        rec = TapGemmRec("t", plan, views, gw, gh, B, tile, pack, pack.packed, pack.sched, n_rows, cout,
                         n_tile, L.EPI_BIAS, out, ostr, ooff, bias=conv.bias)
        prog.recs.append(rec)
        prog.finalize()
        s = torch.cuda.current_stream().cuda_stream
        ws.refresh(s)
        prog.run(s)
        torch.cuda.synchronize()
        xin = torch.cat([x.float() for x in xs], -1).permute(0, 3, 1, 2)
        wq = conv.weight.detach().to(torch.bfloat16).float()
        if kind == "up2x3x3":
            ref = F.conv2d(F.interpolate(xin, scale_factor=2, mode="nearest"), wq, conv.bias, padding=1)
        elif kind == "down4x4s2":
            ref = F.conv2d(xin, wq, conv.bias, stride=2, padding=1)
        else:
            ref = F.conv2d(xin, wq, conv.bias, padding=k // 2)
        ref = ref.permute(0, 2, 3, 1)
        r, m = rel(out, ref)
        print(f"{kind:10s} B={B} {H}x{W} cins={cins} cout={cout} tile={rec.tile} R={plan.R} n_tile={n_tile}: rel={r:.3e} max={m:.3e}",
              "OK" if r < 1e-2 else "FAIL", flush=True)
        if r >= 1e-2:
            d = (out.float() - ref).abs()
            bad = (d > 0.05 * ref.abs().max()).nonzero()
            print("   first bad idx (b,h,w,c):", bad[:8].tolist(), " n_bad", len(bad), "of", d.numel())
            print("   out[0,0,0,:8]", out[0, 0, 0, :8].float().tolist())
            print("   ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())


def case_prog(spec_name):
    import torch
    from ccdm_b200.engine import UnetProgram, WeightStore, TapGemmRec
    from tests.emu_engine import run_program
    from tests.test_engine_emulated import build
    from tests.golden.cases import SIZES, BATCH, unet_inputs
    dev = torch.device("cuda")
    spec, net_cpu, sd = build(spec_name, 5)
    net_cpu.eval()
    x, t, emb = unet_inputs(spec_name)
    B, size = BATCH[spec_name], SIZES[spec_name]
    keep = torch.tensor([1, 0, 1, 0, 1][:B], dtype=torch.uint8)
    ws_c = WeightStore(torch.device("cpu"))
    pc = UnetProgram(net_cpu, ws_c, B, B, size, size, False)
    pc.x_in.copy_(x); pc.t_in.copy_(t); pc.emb_in.copy_(emb); pc.keep.copy_(keep)
    run_program(pc, ws_c)

    _, net_g, _ = build(spec_name, 5)
    net_g = net_g.to(dev).eval()
    ws_g = WeightStore(dev)
    pg = UnetProgram(net_g, ws_g, B, B, size, size, False)
    pg.x_in.copy_(x.to(dev)); pg.t_in.copy_(t.to(dev)); pg.emb_in.copy_(emb.to(dev)); pg.keep.copy_(keep.to(dev))
    s = torch.cuda.current_stream().cuda_stream
    ws_g.refresh(s)
    torch.cuda.synchronize()
    nbad = 0
    for name, rec in ws_g.packs.items():
        if not hasattr(rec, "packed"):
            continue
        r, m = rel(rec.packed, ws_c.packs[name].packed)
        if r > 1e-3:
            nbad += 1
            print(f"  PACK {name}: rel={r:.3e}")
    print(f"packed weights: {len(ws_g.packs)} checked, {nbad} bad", flush=True)
    pg.run(s)
    torch.cuda.synchronize()
    shown = 0
    for name, bg in pg.bufs.items():
        bc = pc.bufs[name]
        if bg.dtype in (torch.int64, torch.uint8):
            continue
        r, m = rel(bg, bc)
        flag = "" if r < 2e-2 else "  <<<<<< DIVERGES"
        if flag or name in ("out", "ss_all", "stem") or name.endswith(".ctx"):
            print(f"  {name:32s} rel={r:.3e} max={m:.3e}{flag}", flush=True)
            shown += bool(flag)
        if shown >= 6:
            break
    r, m = rel(pg.out, pc.out)
    print(f"prog:{spec_name} final rel={r:.3e}", "OK" if r < 2e-2 else "FAIL")


def case_misc():
    import torch
    from ccdm_b200 import _lib as L
    from oracle.unet_ref import cfg_combine
    dev = torch.device("cuda")
    torch.manual_seed(1)
    c, n = torch.randn(4, 3, 16, 16, device=dev), torch.randn(4, 3, 16, 16, device=dev)
    out = torch.empty_like(c)
    s = torch.cuda.current_stream().cuda_stream
    for scale, phi in [(1.5, 0.7), (2.0, 0.0)]:
        L.check(L.lib().ccdm_cfg_combine(c.data_ptr(), n.data_ptr(), out.data_ptr(), 4, 768, scale, phi, 1, 0.0, s))
        torch.cuda.synchronize()
        print(f"cfg_combine s={scale} phi={phi}: rel={rel(out, cfg_combine(c, n, scale, phi))[0]:.3e}")


def main():
    if len(sys.argv) > 1:
        case = sys.argv[1]
        kind, _, arg = case.partition(":")
        {"gemm": case_gemm, "prog": case_prog, "misc": lambda _: case_misc()}[kind](arg)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for case in CASES:
        t0 = time.time()
        print(f"===== {case}", flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), case], cwd=ROOT, timeout=240,
                               capture_output=True, text=True)
            print(p.stdout[-6000:])
            if p.returncode != 0:
                print(f"[exit {p.returncode}] stderr tail:\n{p.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", (e.stdout or b"")[-2000:])
        print(f"===== {case} done in {time.time()-t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
