"""Generate golden vectors by running the reference's OWN modules (CPU, build container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.pt

The reference cannot travel to the GPU box, so the outputs are committed as
small fixtures.  Inputs are not stored: weights come from
``oracle.make_state_dict`` (a pure function of key/shape/seed) and tensors from
seeded generators, both re-created by the tests.  Recipe: SURVEY.md §8(c).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference/CCDM_unified"

import torch  # noqa: E402
from torch import nn  # noqa: E402

from tests.golden.cases import (  # noqa: E402
    SPECS, unet_inputs, UNET_CASES, CFG_CASES, SAMPLER_CASES, LOSS_CASES, loss_inputs,
)
import oracle  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    from models import Unet            # type: ignore
    from diffusion import GaussianDiffusion  # type: ignore
    from label_embedding import LabelEmbed   # type: ignore
    return Unet, GaussianDiffusion, LabelEmbed


LabelEmbed = None


def label_hooks(spec, size, label_dim=1):
    """The reference's own sinusoidal fn_y2h / fn_y2cov (label_embedding.py:861,1035)."""
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        le = LabelEmbed(dataset=None, path_y2h="/tmp/ccdm_golden_y2h", path_y2cov="/tmp/ccdm_golden_y2cov",
                        y2h_type="sinusoidal", y2cov_type="sinusoidal", h_dim=spec.embed_input_dim,
                        cov_dim=spec.in_channels * size * size, nc=spec.in_channels,
                        device=torch.device("cpu"), label_dim=max(label_dim, 1), dim_combination="mean")
    return le.fn_y2h, le.fn_y2cov


def build_ref_unet(Unet, spec, seed, cond_drop_prob=0.1):
    net = Unet(dim=spec.dim, embed_input_dim=spec.embed_input_dim, cond_drop_prob=cond_drop_prob,
               dim_mults=spec.dim_mults, in_channels=spec.in_channels,
               attn_dim_head=spec.attn_dim_head, attn_heads=spec.attn_heads)
    net.load_state_dict(oracle.make_state_dict(spec, seed), strict=True)   # pins key names + shapes
    return net


def main():
    torch.set_num_threads(8)
    global LabelEmbed
    Unet, GaussianDiffusion, LabelEmbed = import_reference()
    out = {}

    # ---- schedules (bit-exact tables)
    sched = {}
    for (T, kind, obj) in [(1000, "cosine", "pred_noise"), (1000, "linear", "pred_x0"),
                           (1000, "cosine", "pred_v"), (200, "cosine", "pred_x0")]:
        spec = SPECS["tiny"]
        gd = GaussianDiffusion(nn.DataParallel(build_ref_unet(Unet, spec, 1)), image_size=16, timesteps=T,
                               objective=obj, beta_schedule=kind)
        sched[f"{T}_{kind}_{obj}"] = {n: getattr(gd, n).clone() for n in oracle.Schedule.NAMES}
    out["schedules"] = sched

    # ---- UNet forward
    unet = {}
    for name, (spec_name, seed, mode, p, mask_seed) in UNET_CASES.items():
        spec = SPECS[spec_name]
        net = build_ref_unet(Unet, spec, seed)
        net.train(mode == "train")
        x, t, emb = unet_inputs(spec_name)
        if mask_seed is not None:
            torch.manual_seed(mask_seed)
        with torch.no_grad():
            y = net(x, t, emb, cond_drop_prob=p)
        rec = {"out": y.clone()}
        if mode == "train":
            rec["bn"] = {k: v.clone() for k, v in net.state_dict().items() if "running_" in k}
        unet[name] = rec
    out["unet"] = unet

    # ---- classifier-free guidance combine
    cfg = {}
    for name, (spec_name, seed, scale, phi) in CFG_CASES.items():
        spec = SPECS[spec_name]
        net = build_ref_unet(Unet, spec, seed).eval()
        x, t, emb = unet_inputs(spec_name)
        with torch.no_grad():
            g, n = net.forward_with_cond_scale(x, t, emb, cond_scale=scale, rescaled_phi=phi)
        cfg[name] = {"guided": g.clone(), "null": n.clone()}
    out["cfg"] = cfg

    # ---- samplers
    samp = {}
    for name, c in SAMPLER_CASES.items():
        spec = SPECS[c["spec"]]
        net = build_ref_unet(Unet, spec, c["seed"]).eval()
        fn_y2h, fn_y2cov = label_hooks(spec, c["size"])
        gd = GaussianDiffusion(nn.DataParallel(net), image_size=c["size"], use_Hy=c["use_Hy"],
                               fn_y2cov=fn_y2cov if c["use_Hy"] else None, timesteps=c["T"],
                               sampling_timesteps=c["S"], objective=c["objective"],
                               ddim_sampling_eta=c["eta"]).eval()
        labels = torch.linspace(0.05, 0.95, c["B"])
        torch.manual_seed(c["rng"])
        with torch.inference_mode():
            if c["kind"] == "ddim":
                img = gd.ddim_sample(labels_emb=fn_y2h(labels), labels=labels,
                                     shape=(c["B"], spec.in_channels, c["size"], c["size"]),
                                     cond_scale=c["scale"])
            else:
                img = gd.sample(labels_emb=fn_y2h(labels), labels=labels, cond_scale=c["scale"])
        samp[name] = {"img": img.clone()}
    out["sampler"] = samp

    # ---- vicinal loss (+ gradients)
    loss = {}
    for name, c in LOSS_CASES.items():
        spec = SPECS[c["spec"]]
        net = build_ref_unet(Unet, spec, c["seed"], cond_drop_prob=c["p_drop"]).train()
        fn_y2h, fn_y2cov = label_hooks(spec, c["size"], c["label_dim"])
        gd = GaussianDiffusion(nn.DataParallel(net), image_size=c["size"], use_Hy=c["use_Hy"],
                               fn_y2cov=fn_y2cov if c["use_Hy"] else None, cond_drop_prob=c["p_drop"],
                               timesteps=1000, objective=c["objective"], vicinity_type=c["vic"]).train()
        img, labels, emb_in = loss_inputs(c)
        torch.manual_seed(c["rng"])
        kw = dict(labels_emb=fn_y2h(emb_in), labels=labels,
                  vicinal_weights=None if c["vic"] is None else torch.ones(len(labels)))
        if c["vic"] is not None:
            kw.update(vicinity_type=c["vic"], kappa=c["kappa"], vector_type="gaussian",
                      num_projections=c.get("nproj", 1))
        val = gd(img, **kw)
        val.backward()
        g = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        loss[name] = {"loss": val.detach().clone(),
                      "grad_sqnorm": sum(float((v.double() ** 2).sum()) for v in g.values()),
                      "grad_final_conv_bias": g["final_conv.bias"],
                      "grad_null_cond_emb": g["null_cond_emb"],
                      "grad_init_conv_bias": g["init_conv.bias"]}
    out["loss"] = loss

    for k, v in out.items():
        path = os.path.join(HERE, f"{k}.pt")
        torch.save(v, path)
        print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


if __name__ == "__main__":
    main()
