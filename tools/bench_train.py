"""Times one optimizer step of the training path: q_sample -> UNet forward ->
vicinal loss -> backward through ccdm_b200/train.py -> Adam.  CUDA events, W warm-up steps, K timed steps, max over
ranks; under torchrun it also averages gradients over the ranks (ccdm_b200.dist.all_reduce_gradients).

  python tools/bench_train.py [--model uk64|rc64|vrc64] [--batch 128] [--steps 5] [--warmup 3] [--breakdown]

Prints one JSON line: ms/step, images/s, algorithmic TFLOP/s (3 x forward FLOPs per image, the usual fwd+dgrad+wgrad
accounting) against MEASURED_PEAKS.json, and with --breakdown the forward / backward / optimizer split.
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccdm_b200  # noqa: E402
import ccdm_b200.optim  # noqa: E402,F401
from ccdm_b200 import dist as D  # noqa: E402

MODELS = {
    # forward GFLOP per image from SURVEY.md section 8d (conv + linear + bmm, 2*MAC)
    "uk64": dict(dim=72, dim_mults=(1, 2, 4, 4, 8), size=64, gflop=16.25),
    "rc64": dict(dim=64, dim_mults=(1, 2, 2, 4, 8), size=64, gflop=11.00),
    # BASELINE configs[4]: UTKFace 192x192 (UK192/run_ccdm.sh:18-33: dim 64, mults 1-2-2-4-4-8-8, 16 per GPU x accum 4)
    "uk192": dict(dim=64, dim_mults=(1, 2, 2, 4, 4, 8, 8), size=192, gflop=94.99),
    # vanilla (GroupNorm) tree, RC-49 64x64 script configuration (V/scripts/run_train_ccdm.sh: batch 128, pred_x0); eager only
    "vrc64": dict(vanilla=True, size=64, gflop=21.8),
}


def sinusoid(y, dim):
    half = dim // 2
    f = torch.exp(-math.log(10000) * torch.arange(half, device=y.device, dtype=torch.float32) / half)
    a = y.reshape(-1)[:, None].float() * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="uk64", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam + clip_grad_norm_ instead of FusedAdam")
    ap.add_argument("--graph", action="store_true", help="capture the whole step into a CUDA graph (train_graph.py)")
    ap.add_argument("--accum", type=int, default=1, help="micro-batches per optimizer step (gradient_accumulate_every); --graph only")
    a = ap.parse_args()
    rank, local, world = D.env_world()
    torch.cuda.set_device(local)
    D.init("nccl")
    m = MODELS[a.model]
    torch.manual_seed(111)
    n_el = 3 * m["size"] ** 2
    if m.get("vanilla"):
        if a.graph:
            raise SystemExit("--graph is not wired for the vanilla model (its UNet draws the label-drop mask itself)")
        net = ccdm_b200.VanillaUnet(embed_input_dim=128, cond_drop_prob=0.1, model_channels=64, num_res_blocks=2,
                                    attention_resolutions=(16, 32), channel_mult=(1, 2, 4, 8), num_heads=4, num_groups=8)
        gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=m["size"], objective="pred_x0", timesteps=1000).cuda().train()
    else:
        net = ccdm_b200.Unet(dim=m["dim"], embed_input_dim=128, cond_drop_prob=0.1, dim_mults=m["dim_mults"], in_channels=3,
                             attn_dim_head=32, attn_heads=4)
        fn_y2cov = lambda y: (sinusoid(y, n_el) + 1) / 2                  # sinusoidal covariance embedding, cov_dim = C*H*W
        gd = ccdm_b200.GaussianDiffusion(net, image_size=m["size"], objective="pred_x0", use_Hy=True, fn_y2cov=fn_y2cov,
                                         cond_drop_prob=0.1, timesteps=1000, vicinity_type="hv").cuda().train()
    D.broadcast_parameters(gd)
    if a.torch_adam:
        opt = torch.optim.Adam([p for p in gd.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99))
    else:
        opt = ccdm_b200.optim.FusedAdam([p for p in gd.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99),
                                        max_grad_norm=1.0,
                                        early_params=None if os.environ.get("CCDM_NO_OVERLAP") else D.early_gradient_params(gd))
    if not a.torch_adam:
        D.wire_overlap(gd, opt)                                        # decoder-segment all-reduce under the encoder's backward
    g = torch.Generator().manual_seed(rank)
    B = a.batch
    img = torch.rand(B, 3, m["size"], m["size"], generator=g).cuda()
    labels = torch.rand(B, generator=g).cuda()
    emb = sinusoid(labels, 128)
    ones = torch.ones(B, device="cuda")
    params = [p for p in gd.parameters() if p.requires_grad]
    lib = ccdm_b200._lib.lib()

    def step(ev=None):
        if ev:
            ev[0].record()
        if m.get("vanilla"):
            loss = gd(img, classes=emb, vicinal_weights=ones.clone())      # the trainer's per-sample weights (V/trainer.py)
        else:
            loss = gd(img, labels_emb=emb, labels=labels, vicinal_weights=ones, vicinity_type="hv", kappa=0.05)
        if ev:
            ev[1].record()
        opt.zero_grad(set_to_none=True)
        if not a.torch_adam:
            opt.arm_early_bucket(True)
        loss.backward()
        if ev:
            ev[2].record()
        if a.torch_adam:
            D.all_reduce_gradients(params)
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        else:
            opt.all_reduce_gradients()
        opt.step()
        if ev:
            ev[3].record()
        return loss

    if a.graph:
        from ccdm_b200.train_graph import GraphedTrainStep
        if a.accum > 1:                                                # micro-batches built inside the graph (same tensors: a throughput run)
            batch_fn = lambda: (img, labels, emb, ones, dict(vicinity_type="hv", kappa=0.05))
            gstep = GraphedTrainStep(gd, opt, batch_fn=batch_fn, accumulate=a.accum)
            step = lambda ev=None: gstep.replay()
        else:
            gstep = GraphedTrainStep(gd, opt, img, labels, emb, loss_kwargs=dict(vicinity_type="hv", kappa=0.05))
            step = lambda ev=None: gstep(img, labels, emb)
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    D.barrier()
    l0 = lib.ccdm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.steps)]
    e0.record()
    for k in range(a.steps):
        loss = step(evs[k] if (a.breakdown and not a.graph) else None)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / a.steps, device=torch.device("cuda", local))
    launches = (lib.ccdm_launch_count() - l0) // a.steps
    if rank == 0:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        tflops = 3 * m["gflop"] * B * a.accum * world / ms               # GFLOP / ms = TFLOP/s
        rec = dict(metric="train step", mode="cuda-graph" if a.graph else "eager", optimizer="torch" if a.torch_adam else "fused", model=a.model, per_gpu_batch=B, accumulate=a.accum, n_gpus=world, ms_per_step=round(ms, 2),
                   images_per_s=round(B * a.accum * world / ms * 1e3, 1), algorithmic_tflops=round(tflops, 1),
                   frac_of_sustained_bf16_peak=round(tflops / world / peaks["bf16_tflops_sustained"], 3),
                   kernel_launches_per_step=int(launches), loss=round(loss.item(), 5),
                   peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
        if a.breakdown and not a.graph:
            f = sum(e[0].elapsed_time(e[1]) for e in evs) / a.steps
            b = sum(e[1].elapsed_time(e[2]) for e in evs) / a.steps
            o = sum(e[2].elapsed_time(e[3]) for e in evs) / a.steps
            rec.update(forward_ms=round(f, 2), backward_ms=round(b, 2), allreduce_clip_adam_ms=round(o, 2))
        print(json.dumps(rec), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/bench_train.jsonl", "a") as fh:
            fh.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
