#!/usr/bin/env python
"""Benchmark of the CCDM denoiser hot path on B200: DDIM images/sec, RC-49 64x64 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch: a full DDIM-250 guided sampling (cond_scale 1.5, 2 UNet
evaluations per denoising step) of ``--batch`` images per GPU with the RC-49 UNet (dim 64, mults 1-2-2-4-8),
random-init weights and synthetic continuous labels.  Prints ONE JSON line (rank 0):
  value    images/s with labels / embeddings already resident in HBM, device-timed (CUDA events), max over ranks
  e2e      the same through Trainer-style host buffers: pinned labels H2D, uint8 images D2H inside the timed region
  roofline the tap-GEMM (conv) kernel: algorithmic conv FLOPs / CUDA-event kernel time vs the measured bf16 peak
  cpu_baseline  the oracle port of the reference on the host cores, on a bounded sample of the same workload
``--impl reference`` times that CPU port alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ddim_images_per_sec"
UNIT = "images/s"
WORKLOAD = "RC-49 {size}x{size} CCDM UNet (dim 64, mults 1-2-2-4-8), random-init, DDIM-{steps} guided sampling (cond_scale 1.5)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=200, help="images per GPU per step (reference --samp_batch_size)")
    ap.add_argument("--ddim-steps", type=int, default=250)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step block")
    return ap.parse_args()


def config(args, world):
    return {"workload": WORKLOAD.format(size=args.size, steps=args.ddim_steps), "image_size": args.size, "ddim_steps": args.ddim_steps, "cond_scale": 1.5,
            "rescaled_phi": 0.7, "objective": "pred_x0", "batch_per_gpu": args.batch,
            "global_batch": args.batch * world, "parallelism": f"sample-sharded x{world}, no data-path collective",
            "l2": "per-step activations (GBs) exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- CPU reference arm (oracle port)

class CpuReference:
    """The oracle port of the reference (fp32 PyTorch on the host cores) driving guided DDIM steps of the same
    workload.  One `denoise_step(batch)` = 2 UNet forwards + guidance + DDIM update for `batch` images, i.e.
    1/ddim_steps of a bench step for that many images."""

    def __init__(self, size, seed=111):
        import oracle
        from oracle.unet_ref import UnetSpec, unet_forward, make_state_dict
        self.oracle, self.size = oracle, size
        self.spec = UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 8))
        self.sd = make_state_dict(self.spec, seed)
        self.sch = oracle.make_schedule(1000, "cosine", "pred_x0")
        self.net = lambda x, t, e, p: unet_forward(self.sd, self.spec, x, t, e, cond_drop_prob=p)
        self.pairs = oracle.diffusion_ref.ddim_time_pairs(1000, 250)
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)

    def denoise_steps(self, batch, n_steps):
        o = self.oracle
        emb = o.y2h_sinusoidal(torch.linspace(0, 1, batch), 128)
        t0 = time.perf_counter()
        with torch.inference_mode():
            img = torch.randn(batch, 3, self.size, self.size)
            for tm, tn in self.pairs[:n_steps]:
                tt = torch.full((batch,), tm, dtype=torch.long)
                eps, x0 = o.model_predictions(self.sch, self.net, img, tt, emb, 1.5, 0.7, clip_x_start=True)
                an = self.sch.alphas_cumprod[max(tn, 0)]
                img = x0 * an.sqrt() + (1 - an).sqrt() * eps
        return time.perf_counter() - t0


CPU_BATCH = 16      # the reference's own CPU-runnable case (BASELINE configs[0]: batch 16); measured 2x the img/s of batch 2


def cpu_baseline(args, budget_s=15.0, batch=CPU_BATCH):
    """Bounded sample: as many guided DDIM steps of `batch` images as fit in ~budget_s (at least one)."""
    ref = CpuReference(args.size)
    t_one = ref.denoise_steps(batch, 1)                      # also the warm-up
    n = max(1, min(8, int(budget_s / max(t_one, 1e-3))))
    dt = ref.denoise_steps(batch, n)
    ips = batch / (dt / n * args.ddim_steps)
    return {"value": ips, "unit": UNIT, "cores": ref.cores, "kind": "port",
            "sample": f"batch {batch}, {n} of {args.ddim_steps} guided DDIM steps (2 UNet forwards each) in {dt:.1f} s "
                      f"on {ref.cores} host threads, scaled by {args.ddim_steps}/{n}; oracle/ port of the reference "
                      f"(fp32 PyTorch CPU)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    ref = CpuReference(args.size)
    batch, n = CPU_BATCH, 1                                  # one bench step = a bounded sample: 1 of 250 DDIM steps
    for _ in range(args.warmup):
        ref.denoise_steps(batch, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.denoise_steps(batch, n)
    dt = (time.perf_counter() - t0) / args.steps
    ips = batch / (dt / n * args.ddim_steps)
    sample = (f"each step = batch {batch}, {n} of {args.ddim_steps} guided DDIM steps on {ref.cores} host threads, "
              f"scaled by {args.ddim_steps}/{n}; oracle/ port of the reference (the reference is pure PyTorch; its "
              f"CPU path is this arithmetic)")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args, world),
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- B200 arm

def run_b200(args):
    import ccdm_b200
    from ccdm_b200 import dist as D, _lib as L
    from ccdm_b200.engine import TapGemmRec, tapgemm_flops
    rank, local_rank, world = D.init("nccl" if int(os.environ.get("WORLD_SIZE", 1)) > 1 else None)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, S, size = args.batch, args.ddim_steps, args.size

    torch.manual_seed(111)
    net = ccdm_b200.Unet(dim=64, embed_input_dim=128, cond_drop_prob=0.1, dim_mults=(1, 2, 2, 4, 8), in_channels=3,
                         attn_dim_head=32, attn_heads=4).to(dev).eval()
    gd = ccdm_b200.GaussianDiffusion(net, image_size=size, timesteps=1000, sampling_timesteps=S, objective="pred_x0",
                                     beta_schedule="cosine", ddim_sampling_eta=0).to(dev).eval()
    embed = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=dev)
    shape = (B, 3, size, size)
    lo, hi = D.shard_bounds(B * world, rank, world)
    labels_host = torch.linspace(0, 1, B * world)[lo:hi].contiguous().pin_memory()
    labels_dev = labels_host.to(dev)
    emb_dev = embed.fn_y2h(labels_dev)

    def step_resident():
        return gd.ddim_sample(labels_emb=emb_dev, labels=labels_dev, shape=shape, cond_scale=1.5)

    out_host = torch.empty(shape, dtype=torch.uint8).pin_memory()

    def step_e2e():
        y = labels_host.to(dev, non_blocking=True)                          # H2D of this step's inputs
        img = gd.ddim_sample(labels_emb=embed.fn_y2h(y), labels=y, shape=shape, cond_scale=1.5)
        out_host.copy_((torch.clip(img, 0, 1) * 255.0).to(torch.uint8), non_blocking=True)   # trainer.py:853-854
        torch.cuda.synchronize()
        return out_host

    for _ in range(max(args.warmup, 1)):
        step_resident()
    torch.cuda.synchronize()
    launches0 = L.lib().ccdm_launch_count()

    # ---- value: device-timed, inputs resident
    with ClockSampler(local_rank) as clk:
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = clk.summary()
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: host buffers in, host uint8 images out
    step_e2e()
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    D.barrier()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e = {"value": world * B * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(labels_host.numel() * labels_host.element_size()),
           "d2h_bytes_per_step": int(out_host.numel())}

    # ---- kernels launched per step: graph replays of (broadcast + UNet program + step + counter) per DDIM step
    prog = net.engine().program(2 * B, B, size, size, False)
    per_denoise = len(prog.calls) + 3
    gpu_launches = args.steps * S * per_denoise

    # ---- roofline of the dominant kernel (tap-GEMM): eager pass with a CUDA event between launches
    prog.run_timed()
    timed = prog.run_timed()
    g_ms = sum(ms_ for r, ms_ in timed if isinstance(r, TapGemmRec))
    g_fl = sum(tapgemm_flops(r) for r, _ in timed if isinstance(r, TapGemmRec))
    n_g = sum(1 for r, _ in timed if isinstance(r, TapGemmRec))
    all_ms = sum(ms_ for _, ms_ in timed)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    burst = float(peaks.get("bf16_tflops", 1590.0))
    achieved = g_fl / (g_ms / 1e3) / 1e12
    # sanity of the FLOP accounting: no single launch may exceed the burst matmul peak unless its plan folds taps
    # (nearest-2x + 3x3: 9 reference taps are executed as 4 summed ones -- algorithmic credit, not tensor-pipe rate)
    for r, ms_ in timed:
        if isinstance(r, TapGemmRec) and not r.plan.out_parity:
            assert tapgemm_flops(r) / ms_ / 1e9 <= burst, (r.name, tapgemm_flops(r) / ms_ / 1e9, "TFLOP/s > burst peak")
    traffic = None                      # dram read+write bytes per launch (avg) from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_tapgemm_traffic.json")))
        if B == 200 and size == 64:
            traffic = tr["traffic_bytes_per_launch_avg"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "tapgemm_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                "launches_per_forward": n_g, "avg_launch_us": 1e3 * g_ms / max(n_g, 1),
                "flops_per_launch_avg": g_fl / max(n_g, 1), "share_of_unet_time": g_ms / all_ms,
                "how": "algorithmic conv FLOPs (2*B*Ho*Wo*Cout*Cin*kh*kw) of all tap-GEMM launches of one 2B UNet "
                       "forward / their summed CUDA-event durations (eager pass, same stream)"}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_layers.json"), "w") as f:
            rows = []
            for r, ms_ in timed:
                if isinstance(r, TapGemmRec):
                    fl = tapgemm_flops(r)
                    rows.append({"name": r.name, "kind": r.plan.kind, "grid": [r.gB, r.gH, r.gW], "cin": sum(r.plan.cins),
                                 "N": r.N, "n_tile": r.n_tile, "nkb": r.plan.nkb, "R": r.plan.R, "tile": list(r.tile), "ms": ms_, "tflops": fl / ms_ / 1e9,
                                 "folded_taps": bool(r.plan.out_parity)})
                else:
                    rows.append({"name": getattr(r, "kind", "?"), "ms": ms_})
            json.dump({"batch": 2 * B, "total_ms": all_ms, "rows": rows}, f, indent=1)

    # ---- training step (BASELINE configs[1]) so that the driver's BENCH / SCALE records carry it at every N
    train = None
    if not args.no_train:
        del prog, timed
        torch.cuda.empty_cache()
        try:
            train = train_block(rank, local_rank, world, dev, peaks)
        except Exception as e:                                             # never lose the headline line to the extra block
            train = {"error": f"{type(e).__name__}: {e}"[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": gpu_launches, "roofline": roofline}
    if train is not None:
        line["train"] = train
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def train_block(rank, local_rank, world, dev, peaks, batch=128, steps=5, warmup=3):
    """BASELINE.json configs[1]: UTKFace 64x64 CCDM training step (dim 72, mults 1-2-4-4-8; pred_x0, Hy, hard vicinal weights),
    batch 128 per GPU, bf16 tensor cores with fp32 master weights.  One step = q_sample -> UNet forward -> vicinal loss ->
    backward -> gradient all-reduce over NCCL (N > 1) -> clip -> Adam, replayed as ONE CUDA graph (train_graph.py).  CUDA
    events, max over ranks.  Algorithmic FLOPs = 3 x forward (fwd + dgrad + wgrad), 16.25 GF per image forward."""
    import math
    import ccdm_b200
    from ccdm_b200 import dist as D
    from ccdm_b200.optim import FusedAdam
    from ccdm_b200.train_graph import GraphedTrainStep
    size, gflop = 64, 16.25

    def sinusoid(y, dim):
        half = dim // 2
        f = torch.exp(-math.log(10000) * torch.arange(half, device=y.device, dtype=torch.float32) / half)
        a = y.reshape(-1)[:, None].float() * f[None]
        return torch.cat([torch.cos(a), torch.sin(a)], -1)

    torch.manual_seed(111)
    net = ccdm_b200.Unet(dim=72, embed_input_dim=128, cond_drop_prob=0.1, dim_mults=(1, 2, 4, 4, 8), in_channels=3,
                         attn_dim_head=32, attn_heads=4)
    n_el = 3 * size * size
    gd = ccdm_b200.GaussianDiffusion(net, image_size=size, objective="pred_x0", use_Hy=True,
                                     fn_y2cov=lambda y: (sinusoid(y, n_el) + 1) / 2, cond_drop_prob=0.1, timesteps=1000,
                                     vicinity_type="hv").to(dev).train()
    D.broadcast_parameters(gd)
    opt = FusedAdam([p for p in gd.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), max_grad_norm=1.0,
                    early_params=D.early_gradient_params(gd))
    g = torch.Generator().manual_seed(rank)
    img = torch.rand(batch, 3, size, size, generator=g).to(dev)
    labels = torch.rand(batch, generator=g).to(dev)
    emb = sinusoid(labels, 128)
    gstep = GraphedTrainStep(gd, opt, img, labels, emb, loss_kwargs=dict(vicinity_type="hv", kappa=0.05), warmup=warmup)
    gstep(img, labels, emb)
    torch.cuda.synchronize()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = gstep(img, labels, emb)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    tflops = 3 * gflop * batch * world / ms
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    return {"workload": "UTKFace 64x64 CCDM training step (dim 72, mults 1-2-4-4-8, pred_x0 + Hy, hard vicinal weights), "
                        "synthetic batch, whole step as one CUDA graph", "batch_per_gpu": batch, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "images_per_s": batch * world / ms * 1e3,
            "algorithmic_tflops": tflops, "frac_of_sustained_bf16_peak": tflops / world / peak,
            "gradient_exchange": "none (1 GPU)" if world == 1 else "NCCL all-reduce of the flat fp32 gradient inside the graph: decoder segment on a side stream under the encoder's backward, the rest after it",
            "loss": float(loss.item()), "dtype": "bf16 (fp32 master weights, fp32 accumulation)"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
