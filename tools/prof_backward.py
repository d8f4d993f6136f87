"""Times the backward building blocks at RC-49 64x64 training shapes (per-GPU batch 128) and prints one JSON line per
kernel: achieved TFLOP/s (tensor-core kernels) or GB/s (row kernel) against MEASURED_PEAKS.json."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccdm_b200 import backward as bw  # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(os.environ.get("B", 128))
    out = []
    shapes = [("3x3", (64,), 64, 64), ("3x3", (128,), 128, 32), ("3x3", (128, 128), 128, 32), ("3x3", (256,), 256, 16),
              ("3x3", (512,), 512, 8), ("down4x4s2", (64,), 64, 64), ("up2x3x3", (128,), 64, 32), ("1x1", (64,), 384, 64),
              # UTKFace widths (dim 72): BASELINE configs[1], the training step bench.py reports
              ("3x3", (72,), 72, 64), ("3x3", (72, 72), 72, 64), ("3x3", (144,), 144, 32), ("3x3", (288,), 288, 16)]
    for kind, cins, cout, hw in shapes:
        xs = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for c in cins]
        k = {"1x1": 1, "3x3": 3, "down4x4s2": 4, "up2x3x3": 3}[kind]
        w = torch.randn(cout, sum(cins), k, k, device="cuda") / math.sqrt(sum(cins) * k * k)
        z = bw.conv_forward(kind, xs, w)
        dz = torch.randn_like(z)
        flops = 2.0 * z.numel() * sum(cins) * k * k
        # wgrad: time the tensor-core kernel alone (the zero-fill and unpack are separate launches)
        tl = []
        for _ in range(6):
            bw.conv_wgrad(kind, xs, dz, timing=tl)
        torch.cuda.synchronize()
        t_wg = sum(a.elapsed_time(b) for a, b in tl[1:]) / (len(tl) - 1)
        t_wg_all = timed(lambda: bw.conv_wgrad(kind, xs, dz))
        t_dg = timed(lambda: bw.conv_dgrad(kind, dz, w, cins))
        t_fw = timed(lambda: bw.conv_forward(kind, xs, w))
        rec = dict(layer=f"{kind} {cins}->{cout} @{hw}x{hw} B={B}", gflop=round(flops / 1e9, 1),
                   wgrad_kernel_ms=round(t_wg, 3), wgrad_tflops=round(flops / t_wg / 1e9, 1),
                   wgrad_with_unpack_ms=round(t_wg_all, 3),
                   dgrad_eager_ms=round(t_dg, 3), dgrad_tflops=round(flops / t_dg / 1e9, 1),
                   fwd_eager_ms=round(t_fw, 3))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    for c, hw in [(64, 64), (128, 32), (256, 16), (512, 8)]:
        z = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
        dy = torch.randn_like(z)
        gain = torch.ones(c, device="cuda")
        ss = 0.1 * torch.randn(B, 2 * c, device="cuda")
        t = timed(lambda: bw.block_backward(dy, z, gain, ss, 0, True))
        gb = 3 * z.numel() * 2 / 1e9
        rec = dict(layer=f"block_bwd C={c} @{hw}x{hw} B={B}", ms=round(t, 3), gbytes=round(gb, 3),
                   gbps=round(gb / t * 1e3, 1), hbm_peak=PEAKS["hbm_gbs"], frac=round(gb / t * 1e3 / PEAKS["hbm_gbs"], 3))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/prof_backward.json", "w"), indent=1)


if __name__ == "__main__":
    main()
