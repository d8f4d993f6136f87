"""Precision tiers on the device: network-prediction relative error of the bf16 and fp16 (TF32-class mantissa) builds against
the fp32 oracle on the RC-49 64x64 network (forward, guided forward, teacher-forced DDIM steps, free-running DDIM-20 PSNR),
and the guided DDIM throughput of both tiers.   python tools/prof_precision.py  -> gpurun_out/prof_precision.json"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccdm_b200  # noqa: E402
import oracle  # noqa: E402
from oracle.unet_ref import UnetSpec, unet_forward, unet_forward_cfg, make_state_dict  # noqa: E402

RC64 = UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    sd = {k: v.to(dev) for k, v in make_state_dict(RC64, 7).items()}
    out = {}
    B, size = 4, 64
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 3, size, size, generator=g).to(dev)
    t = torch.tensor([17, 333, 650, 990], device=dev)
    labels = torch.linspace(0.1, 0.9, B, device=dev)
    emb = oracle.y2h_sinusoidal(labels, 128)
    with torch.no_grad():
        ref = unet_forward(sd, RC64, x, t, emb, cond_drop_prob=0.0)
        ref_g, _ = unet_forward_cfg(sd, RC64, x, t, emb, cond_scale=1.5, rescaled_phi=0.7)
    for prec in ("bf16", "fp16"):
        net = ccdm_b200.Unet(dim=64, dim_mults=(1, 2, 2, 4, 8), cond_drop_prob=0.1, precision=prec)
        net.load_state_dict(make_state_dict(RC64, 7))
        net = net.to(dev).eval()
        rec = {"forward_rel_err": rel(net(x, t, emb, cond_drop_prob=0.0), ref)}
        gd_, _ = net.forward_with_cond_scale(x, t, emb, cond_scale=1.5, rescaled_phi=0.7)
        rec["guided_rel_err"] = rel(gd_, ref_g)
        for objective in ("pred_x0", "pred_noise"):
            S = 20
            gd = ccdm_b200.GaussianDiffusion(net, image_size=size, timesteps=1000, sampling_timesteps=S, objective=objective).to(dev).eval()
            sch = oracle.make_schedule(1000, "cosine", objective).to(dev)
            net_o = lambda xx, tt, e, p: unet_forward(sd, RC64, xx, tt, e, cond_drop_prob=p)
            shape = (B, 3, size, size)
            torch.manual_seed(5)
            tr_o = []
            ref_img = oracle.ddim_sample(sch, net_o, emb, shape, sampling_timesteps=S, cond_scale=1.5, trace=tr_o)
            torch.manual_seed(5)
            tr = []
            img = gd.ddim_sample(labels_emb=emb, labels=labels, shape=shape, cond_scale=1.5, trace=tr)
            mse = ((img - ref_img) ** 2).mean().item()
            rec[f"{objective}_ddim20_psnr_db"] = 10 * math.log10(1.0 / max(mse, 1e-20))
            rec[f"{objective}_step0_eps_rel_err"] = rel(tr[0][0], tr_o[0][0])
            # teacher forcing: the oracle's state at a few steps fed to the device network; error of its noise prediction
            errs = []
            pairs = oracle.diffusion_ref.ddim_time_pairs(1000, S)
            xs = [torch.randn(shape, generator=torch.Generator().manual_seed(50 + i)).to(dev) for i in range(4)]
            for i, (tm, _) in enumerate(pairs[::5][:4]):
                tt = torch.full((B,), tm, device=dev, dtype=torch.long)
                with torch.no_grad():
                    eps_o, x0_o = oracle.model_predictions(sch, net_o, xs[i], tt, emb, 1.5, 0.7, clip_x_start=True)
                p = gd.model_predictions(xs[i], tt, emb, cond_scale=1.5, rescaled_phi=0.7, clip_x_start=True)
                errs.append(rel(p.pred_noise, eps_o))
            rec[f"{objective}_teacher_forced_eps_rel_err_max"] = max(errs)
        # throughput of the tier: guided DDIM-50, batch 200 (same kernels, same bytes: expected equal)
        gd = ccdm_b200.GaussianDiffusion(net, image_size=size, timesteps=1000, sampling_timesteps=50, objective="pred_x0").to(dev).eval()
        lab = torch.linspace(0, 1, 200, device=dev)
        e200 = oracle.y2h_sinusoidal(lab, 128)
        for _ in range(2):
            gd.ddim_sample(labels_emb=e200, labels=lab, shape=(200, 3, size, size), cond_scale=1.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gd.ddim_sample(labels_emb=e200, labels=lab, shape=(200, 3, size, size), cond_scale=1.5)
        e1.record()
        torch.cuda.synchronize()
        rec["ddim250_equiv_images_per_s"] = 200 * 3 / (e0.elapsed_time(e1) / 1e3) * 50 / 250
        out[prec] = rec
        print(prec, json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/prof_precision.json", "w"), indent=1)


if __name__ == "__main__":
    main()
