"""HBM roofline of the vanilla UNet's GroupNorm kernels at the RC-49 64x64 script shapes:
ccdm_channel_stats (one read), ccdm_groupnorm_coef (tiny), ccdm_affine_act + SiLU (one read + one write), each timed
alone with CUDA events (L2 flushed between launches by cycling through > 126 MB of distinct buffers), reported as
algorithmic GB/s against MEASURED_PEAKS.json hbm_gbs.  One JSON line per shape -> gpurun_out/prof_groupnorm.jsonl.

  python tools/prof_groupnorm.py [--batch 400]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccdm_b200 import _lib as L  # noqa: E402

SHAPES = [(64, 64, 64), (64, 64, 128), (32, 32, 128), (32, 32, 256), (16, 16, 256), (16, 16, 512), (8, 8, 512), (8, 8, 1024)]


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=400)
    a = ap.parse_args()
    dev = torch.device("cuda")
    lib = L.lib()
    st = torch.cuda.current_stream().cuda_stream
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    os.makedirs("gpurun_out", exist_ok=True)
    for (h, w, c) in SHAPES:
        B = a.batch
        nbytes = B * h * w * c * 2
        copies = max(2, int(300e6 // nbytes) + 1)              # > 126 MB of distinct inputs: every launch misses L2
        xs = [torch.randn(B, h, w, c, device=dev).to(torch.bfloat16) for _ in range(copies)]
        out = torch.empty_like(xs[0])
        sums = torch.zeros(B, 2, c, device=dev)
        coef = torch.zeros(B, 2 * c, device=dev)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)

        def stats(i=0):
            L.check(lib.ccdm_channel_stats(xs[i % copies].data_ptr(), B, h * w, c, sums.data_ptr(), c, 0, 1, st))

        def coef_(i=0):
            L.check(lib.ccdm_groupnorm_coef(sums.data_ptr(), B, c, 8, h * w, 1e-5, gamma.data_ptr(), beta.data_ptr(), None, 0,
                                            0, c, coef.data_ptr(), st))

        def apply(i=0):
            L.check(lib.ccdm_affine_act(xs[i % copies].data_ptr(), out.data_ptr(), B * h * w, c, h * w, coef.data_ptr(),
                                        2 * c, 0, 2, st))

        ms_s, ms_c, ms_a = timed(stats), timed(coef_), timed(apply)
        rec = dict(shape=[B, h, w, c], stats_ms=round(ms_s, 4), stats_gbs=round(nbytes / ms_s / 1e6, 1),
                   stats_frac=round(nbytes / ms_s / 1e6 / peak, 3), coef_ms=round(ms_c, 4), apply_ms=round(ms_a, 4),
                   apply_gbs=round(2 * nbytes / ms_a / 1e6, 1), apply_frac=round(2 * nbytes / ms_a / 1e6 / peak, 3),
                   peak_gbs=peak)
        print(json.dumps(rec), flush=True)
        with open("gpurun_out/prof_groupnorm.jsonl", "a") as fh:
            fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
