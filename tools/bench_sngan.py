"""Throughput of the one-step generator forward:
sngan_generator(dim_z=256, dim_embed=128, img_size=192, gene_ch=48), eval mode, random-init weights, synthetic z / y.

  python tools/bench_sngan.py [--batch 64] [--steps 10] [--warmup 3] [--graph] [--cpu-seconds 10]

One JSON line: images/s (CUDA events), algorithmic TFLOP/s of the convolutions as the reference computes them
(3x3 convs on the UPSAMPLED tensor: the 2.25x saving of the folded nearest-2x conv is not credited) against
MEASURED_PEAKS.json, and the oracle (fp32 PyTorch port of the reference) timed on the host cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccdm_b200.sngan as S  # noqa: E402
from ccdm_b200 import _lib as L  # noqa: E402


def conv_flops(spec_channels, init_size, gene_ch, nc, dim_z):
    """2*MAC per image of dense + every conv at the resolution the reference runs it."""
    f = 2.0 * dim_z * init_size * init_size * spec_channels[0][0]
    s = init_size
    for ci, co in spec_channels:
        s *= 2
        f += 2.0 * s * s * (co * ci * 9 + co * co * 9 + co * ci)       # conv1 (on the upsampled map), conv2, bypass 1x1
    f += 2.0 * s * s * nc * gene_ch * 9
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="replay the forward from a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    a = ap.parse_args()
    torch.manual_seed(111)
    net = S.sngan_generator(dim_z=256, dim_embed=128, nc=3, img_size=192, gene_ch=48).cuda().eval()
    for m in net.modules():                                    # non-trivial running statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
    B = a.batch
    z = torch.randn(B, 256, device="cuda")
    y = torch.rand(B, 128, device="cuda")
    lib = L.lib()
    run = lambda: net(z, y)
    out = run()
    if a.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                run()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = run()
        run = g.replay
    for _ in range(a.warmup):
        run()
    torch.cuda.synchronize()
    l0 = lib.ccdm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (lib.ccdm_launch_count() - l0) // a.steps
    chans = [(768, 384), (384, 192), (192, 96), (96, 48), (48, 48)]
    gf = conv_flops(chans, 6, 48, 3, 256) / 1e9
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    rec = dict(metric="generator images/s", mode="cuda-graph" if a.graph else "eager", config="sngan_generator 192px gene_ch 48",
               batch=B, ms_per_forward=round(ms, 3), images_per_s=round(B / ms * 1e3, 1), gflop_per_image=round(gf, 2),
               algorithmic_tflops=round(gf * B / ms, 1), frac_of_sustained_bf16_peak=round(gf * B / ms / peaks["bf16_tflops_sustained"], 3),
               kernel_launches_per_forward=int(launches), finite=bool(torch.isfinite(out).all()))
    if a.cpu_seconds > 0:
        from oracle.sngan_ref import GenSpec, generator_forward
        spec = GenSpec(dim_z=256, dim_embed=128, nc=3, img_size=192, gene_ch=48)
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        zc, yc = z[:4].cpu(), y[:4].cpu()
        with torch.no_grad():
            generator_forward(sd, spec, zc[:1], yc[:1])
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < a.cpu_seconds:
                generator_forward(sd, spec, zc, yc)
                n += 4
            dt = time.perf_counter() - t0
        rec["cpu_baseline"] = dict(value=round(n / dt, 2), unit="images/s", cores=os.cpu_count(), kind="port",
                                   sample=f"{n} images in {dt:.1f} s, batch 4, oracle/sngan_ref.py (fp32 PyTorch CPU)")
    print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_sngan.jsonl", "a") as fh:
        fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
