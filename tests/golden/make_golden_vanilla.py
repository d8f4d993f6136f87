"""Golden outputs of the reference's OWN vanilla UNet (CCDM_vanilla/.../CCGM/CCDM/models/unet.py), CPU, build container only.

    python tests/golden/make_golden_vanilla.py      # needs /root/reference; writes tests/golden/vanilla_unet.pt

Weights come from oracle.vanilla_unet_ref.make_state_dict (loaded with strict=True, which also pins key names, order
and shapes); inputs from seeded generators.  The Bernoulli label-drop mask (unet.py:352) is injected by replacing the
module's ``prob_mask_like`` so that reference and oracle see the same mask.  Only outputs are stored.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import importlib.util  # noqa: E402

import torch  # noqa: E402

from oracle.vanilla_unet_ref import make_state_dict  # noqa: E402
from tests.golden.vanilla_cases import (V_CASES, V_CFG_CASES, V_LOSS_CASES, V_LOSS_GRAD_KEYS, V_SAMPLER_CASES, V_SIZES, V_SPECS, V_BATCH, keep_mask,  # noqa: E402
                                        loss_inputs, sampler_classes, vanilla_inputs)

VREF = "/root/reference/CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(ref, s, seed):
    net = ref.Unet(embed_input_dim=s.embed_input_dim, cond_drop_prob=0.5, in_channels=s.in_channels,
                   model_channels=s.model_channels, out_channels=None, num_res_blocks=s.num_res_blocks,
                   attention_resolutions=s.attention_resolutions, dropout=0, channel_mult=s.channel_mult,
                   conv_resample=True, num_heads=s.num_heads, use_scale_shift_norm=True, learned_variance=False,
                   num_groups=s.num_groups)
    net.load_state_dict(make_state_dict(s, seed), strict=True)
    return net


def main():
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    ref = load(os.path.join(VREF, "models", "unet.py"), "ref_vanilla_unet")
    out = {}
    for name, (sname, seed, mode, kind) in V_CASES.items():
        s = V_SPECS[sname]
        net = build(ref, s, seed)
        net.train(mode == "train")
        x, t, classes = vanilla_inputs(sname)
        mask = keep_mask(kind, V_BATCH[sname])
        ref.prob_mask_like = lambda shape, prob, device, _m=mask: _m.clone()
        with torch.no_grad():
            y = net(x, t, classes, cond_drop_prob=0.5)
        out[name] = {"out": y.clone(), "keys": list(net.state_dict().keys())}
        print(name, tuple(y.shape), float(y.abs().mean()))
    # plain classifier-free guidance (V/diffusion.py:34-56): the function is restated by the oracle; here the reference's
    # own two forwards are combined with the reference's formula evaluated by the reference module's function
    dsrc = open(os.path.join(VREF, "diffusion.py")).read()
    start, end = dsrc.index("def forward_with_cond_scale"), dsrc.index("def extract")
    ns = {"torch": torch, "partial": __import__("functools").partial}
    exec(compile(dsrc[start:end], "ref_forward_with_cond_scale", "exec"), ns)      # the reference's own function text
    del ref.prob_mask_like
    ref_pm = load(os.path.join(VREF, "models", "unet.py"), "ref_vanilla_unet2")       # unpatched prob_mask_like
    for name, (sname, seed, cs, phi) in V_CFG_CASES.items():
        s = V_SPECS[sname]
        net = build(ref_pm, s, seed).eval()
        x, t, classes = vanilla_inputs(sname)
        with torch.no_grad():
            y = ns["forward_with_cond_scale"](x, t, classes, model=net, cond_scale=cs, rescaled_phi=phi)
        out[name] = {"out": y.clone()}
        print(name, tuple(y.shape), float(y.abs().mean()))
    torch.save(out, os.path.join(HERE, "vanilla_unet.pt"))

    # ---- sampling loops through the reference's own GaussianDiffusion (needs utils.py -> accelerate / ema_pytorch stubs)
    for m in ("accelerate",):
        stub = types.ModuleType(m)
        stub.Accelerator = object
        sys.modules.setdefault(m, stub)
    sys.path.insert(0, VREF)
    ref_diff = load(os.path.join(VREF, "diffusion.py"), "ref_vanilla_diffusion")
    samp = {}
    for name, c in V_SAMPLER_CASES.items():
        s = V_SPECS[c["spec"]]
        net = build(ref_pm, s, c["seed"]).eval()
        size = V_SIZES[c["spec"]]
        gd = ref_diff.GaussianDiffusion(torch.nn.DataParallel(net), image_size=size, timesteps=c["T"],
                                        sampling_timesteps=c["S"], objective=c["objective"],
                                        ddim_sampling_eta=c["eta"]).eval()
        classes = sampler_classes(c)
        torch.manual_seed(c["rng"])
        with torch.no_grad():
            if c["kind"] == "ddim":
                img = gd.ddim_sample(classes, (c["B"], s.in_channels, size, size), cond_scale=c["scale"],
                                     rescaled_phi=c["phi"])
            else:
                img = gd.sample(classes, cond_scale=c["scale"], rescaled_phi=c["phi"], preset_sampling_timesteps=c["S"])
        samp[name] = {"img": img.clone()}
        print(name, tuple(img.shape), float(img.mean()))
    torch.save(samp, os.path.join(HERE, "vanilla_sampler.pt"))

    # ---- training loss + gradients through the reference's own p_losses (label-drop mask injected, noise / t given)
    loss = {}
    for name, c in V_LOSS_CASES.items():
        s = V_SPECS[c["spec"]]
        net = build(ref, s, c["seed"]).train()
        mask = keep_mask(c["kind"], V_BATCH[c["spec"]])
        ref.prob_mask_like = lambda shape, prob, device, _m=mask: _m.clone()
        gd = ref_diff.GaussianDiffusion(torch.nn.DataParallel(net), image_size=V_SIZES[c["spec"]], timesteps=1000,
                                        objective=c["objective"]).train()
        x0, t, classes, noise, weights = loss_inputs(c)
        val = gd.p_losses(x0, t, classes=classes, noise=noise, vicinal_weights=None if weights is None else weights.clone())
        val.backward()
        g = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        loss[name] = {"loss": val.detach().clone(),
                      "grad_sqnorm": sum(float((v.double() ** 2).sum()) for v in g.values()),
                      **{"grad_" + k: g[k] for k in V_LOSS_GRAD_KEYS}}
        print(name, float(val), loss[name]["grad_sqnorm"])
    torch.save(loss, os.path.join(HERE, "vanilla_loss.pt"))


if __name__ == "__main__":
    main()
