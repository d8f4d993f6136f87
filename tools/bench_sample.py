"""Guided sampling throughput of the other BASELINE.json sampling configurations (the headline RC-49 64x64 line is
bench.py):  --model sa128  SteeringAngle 128x128 CCDM, DDIM-250, covariance-embedded labels (use_Hy), cond_scale 1.5
            --model uk128  UTKFace 128x128 CcDPM, DDPM-1000 (`.sample`), pred_noise, cond_scale 2.0
            --model vrc64  vanilla (GroupNorm) RC-49 64x64 UNet, DDIM-250, cond_scale 1.5 (V/scripts/run_train_ccdm.sh)
CUDA-event timed, W warm-up + K timed samplings, max over ranks under torchrun; one JSON line.

  python tools/bench_sample.py --model sa128 [--batch 64] [--steps 2] [--warmup 1]
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccdm_b200  # noqa: E402
from ccdm_b200 import dist as D  # noqa: E402

MODELS = {   # forward GFLOP per image from SURVEY.md section 8d
    "sa128": dict(dim=64, dim_mults=(1, 2, 2, 4, 4, 8), size=128, gflop=42.2, sampler="ddim", S=250, objective="pred_x0",
                  use_Hy=True, scale=1.5),
    "uk128": dict(dim=64, dim_mults=(1, 2, 4, 4, 8, 8), size=128, gflop=52.08, sampler="ddpm", S=1000,
                  objective="pred_noise", use_Hy=False, scale=2.0),
    # vanilla tree (CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/scripts/run_train_ccdm.sh); 21.8 GFLOP / image
    "vrc64": dict(vanilla=True, size=64, gflop=21.8, sampler="ddim", S=250, objective="pred_x0", use_Hy=False, scale=1.5),
}


def sinusoid(y, dim):
    half = dim // 2
    f = torch.exp(-math.log(10000) * torch.arange(half, device=y.device, dtype=torch.float32) / half)
    a = y.reshape(-1)[:, None].float() * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sa128", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sampling-steps", type=int, default=0, help="override the number of denoising steps")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: fixed total batch, split over the ranks")
    a = ap.parse_args()
    rank, local, world = D.env_world()
    torch.cuda.set_device(local)
    D.init("nccl")
    dev = torch.device("cuda", local)
    m = MODELS[a.model]
    S = a.sampling_steps or m["S"]
    torch.manual_seed(111)
    n_el = 3 * m["size"] ** 2
    if m.get("vanilla"):
        net = ccdm_b200.VanillaUnet(embed_input_dim=128, cond_drop_prob=0.1, model_channels=64, num_res_blocks=2,
                                    attention_resolutions=(16, 32), channel_mult=(1, 2, 4, 8), num_heads=4, num_groups=8)
        gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=m["size"], objective=m["objective"], timesteps=1000,
                                                sampling_timesteps=S, ddim_sampling_eta=0.0).to(dev).eval()
    else:
        net = ccdm_b200.Unet(dim=m["dim"], embed_input_dim=128, cond_drop_prob=0.1, dim_mults=m["dim_mults"], in_channels=3,
                             attn_dim_head=32, attn_heads=4)
        gd = ccdm_b200.GaussianDiffusion(net, image_size=m["size"], objective=m["objective"], use_Hy=m["use_Hy"],
                                         fn_y2cov=(lambda y: (sinusoid(y, n_el) + 1) / 2) if m["use_Hy"] else None,
                                         cond_drop_prob=0.1, timesteps=1000, sampling_timesteps=S).to(dev).eval()
    B = a.batch if not a.global_batch else a.global_batch // world
    labels = torch.linspace(0, 1, B * world, device=dev)[rank * B:(rank + 1) * B]
    emb = sinusoid(labels, 128)
    shape = (B, 3, m["size"], m["size"])

    def run():
        if m.get("vanilla"):
            return gd.ddim_sample(emb, shape, cond_scale=m["scale"])
        if m["sampler"] == "ddim":
            return gd.ddim_sample(labels_emb=emb, labels=labels, shape=shape, cond_scale=m["scale"])
        return gd.sample(labels_emb=emb, labels=labels, cond_scale=m["scale"])

    for _ in range(a.warmup):
        img = run()
    torch.cuda.synchronize()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        img = run()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / a.steps, device=dev)
    if rank == 0:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        tf = 2 * S * m["gflop"] * B * world / ms                       # two UNet forwards per denoising step
        rec = dict(metric=f"{m['sampler']} images/s", model=a.model, image_size=m["size"], sampling_steps=S,
                   per_gpu_batch=B, scaling="strong" if a.global_batch else "weak", n_gpus=world, ms_per_sampling=round(ms, 1), images_per_s=round(B * world / ms * 1e3, 2),
                   algorithmic_tflops=round(tf, 1), frac_of_sustained_bf16_peak=round(tf / world / peaks["bf16_tflops_sustained"], 3),
                   finite=bool(torch.isfinite(img).all()), in_unit_range=bool(img.min() >= 0 and img.max() <= 1))
        print(json.dumps(rec), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/bench_sample.jsonl", "a") as fh:
            fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
