"""Pins the oracle (oracle/*.py) to outputs of the reference's own modules (tests/golden/*.pt)."""
import os

import pytest
import torch

import oracle
from oracle.unet_ref import unet_forward, unet_forward_cfg, make_state_dict
from oracle import diffusion_ref as D
from tests.golden.cases import (SPECS, SIZES, unet_inputs, UNET_CASES, CFG_CASES, SAMPLER_CASES, LOSS_CASES,
                                loss_inputs)

G = os.path.join(os.path.dirname(__file__), "golden")
load = lambda n: torch.load(os.path.join(G, n + ".pt"), weights_only=True)


def close(a, b, tol=2e-5):
    err = (a.double() - b.double()).abs().max().item()
    ref = b.double().abs().max().item() + 1e-12
    assert err / ref < tol, f"rel err {err/ref:.3e}"


def test_schedule_tables_bit_exact():
    gold = load("schedules")
    for key, tabs in gold.items():
        T, kind, obj = key.split("_", 2)
        sch = oracle.make_schedule(int(T), kind, obj)
        for n in oracle.Schedule.NAMES:
            assert torch.equal(getattr(sch, n), tabs[n]), (key, n)


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_forward(name):
    spec_name, seed, mode, p, mask_seed = UNET_CASES[name]
    spec = SPECS[spec_name]
    sd = make_state_dict(spec, seed)
    x, t, emb = unet_inputs(spec_name)
    if mask_seed is not None:
        torch.manual_seed(mask_seed)
    upd = {}
    with torch.no_grad():
        y = unet_forward(sd, spec, x, t, emb, cond_drop_prob=p, training=(mode == "train"), bn_updates=upd)
    gold = load("unet")[name]
    close(y, gold["out"])
    if mode == "train":
        for k, v in gold["bn"].items():
            close(upd[k], v)


def test_rc64_headline_widths():
    """The headline network itself (dim 64, mults 1-2-2-4-8 at 64x64): oracle vs the reference's own outputs
    (tests/golden/rc64.pt, tests/golden/make_golden_rc64.py)."""
    from tests.golden.make_golden_rc64 import RC64, SEED, rc64_inputs
    x, t, emb = rc64_inputs()
    sd = make_state_dict(RC64, SEED)
    gold = load("rc64")
    torch.set_num_threads(8)
    with torch.no_grad():
        close(unet_forward(sd, RC64, x, t, emb, cond_drop_prob=0.0), gold["cond"], 5e-5)
        g, n = unet_forward_cfg(sd, RC64, x, t, emb, cond_scale=1.5, rescaled_phi=0.7)
    close(n, gold["null"], 5e-5)
    close(g, gold["guided"], 5e-5)


@pytest.mark.parametrize("name", list(CFG_CASES))
def test_cfg(name):
    spec_name, seed, scale, phi = CFG_CASES[name]
    spec = SPECS[spec_name]
    x, t, emb = unet_inputs(spec_name)
    with torch.no_grad():
        g, n = unet_forward_cfg(make_state_dict(spec, seed), spec, x, t, emb, cond_scale=scale, rescaled_phi=phi)
    gold = load("cfg")[name]
    close(g, gold["guided"])
    close(n, gold["null"])


def _net(sd, spec, training=False):
    return lambda x, t, e, p: unet_forward(sd, spec, x, t, e, cond_drop_prob=p, training=training)


@pytest.mark.parametrize("name", list(SAMPLER_CASES))
def test_sampler(name):
    c = SAMPLER_CASES[name]
    spec = SPECS[c["spec"]]
    sd = make_state_dict(spec, c["seed"])
    sch = oracle.make_schedule(c["T"], "cosine", c["objective"])
    labels = torch.linspace(0.05, 0.95, c["B"])
    emb = oracle.y2h_sinusoidal(labels, spec.embed_input_dim)
    shape = (c["B"], spec.in_channels, c["size"], c["size"])
    cov = None
    if c["use_Hy"]:
        cov = torch.exp(-oracle.y2cov_sinusoidal(labels, shape[1] * shape[2] * shape[3]).view(shape))
    torch.manual_seed(c["rng"])
    if c["kind"] == "ddim":
        img = oracle.ddim_sample(sch, _net(sd, spec), emb, shape, sampling_timesteps=c["S"], cond_scale=c["scale"],
                                 eta=c["eta"], init_cov=cov)
    else:
        img = oracle.ddpm_sample(sch, _net(sd, spec), emb, shape, sampling_timesteps=c["S"], cond_scale=c["scale"],
                                 init_cov=cov)
    close(img, load("sampler")[name]["img"], tol=2e-4)


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_loss_and_grads(name):
    c = LOSS_CASES[name]
    spec = SPECS[c["spec"]]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v)
          for k, v in make_state_dict(spec, c["seed"]).items()}
    sch = oracle.make_schedule(1000, "cosine", c["objective"])
    img, labels, emb_in = loss_inputs(c)
    n_el = spec.in_channels * c["size"] ** 2
    torch.manual_seed(c["rng"])
    t = torch.randint(0, 1000, (c["B"],)).long()                      # GaussianDiffusion.forward, diffusion.py:753
    net = lambda x, tt, e: unet_forward(sd, spec, x, tt, e, cond_drop_prob=c["p_drop"], training=True)
    kw = {}
    if c["vic"] is not None:
        kw = dict(vicinity_type=c["vic"], kappa=c["kappa"], num_projections=c.get("nproj", 1))
    val = oracle.p_losses(sch, net, img * 2 - 1, t, labels=labels,
                          labels_emb=oracle.y2h_sinusoidal(emb_in, spec.embed_input_dim),
                          cond_drop_prob=c["p_drop"], use_Hy=c["use_Hy"],
                          fn_y2cov=lambda y: oracle.y2cov_sinusoidal(y, n_el),
                          vicinal_weights=None if c["vic"] is None else torch.ones(c["B"]), **kw)
    gold = load("loss")[name]
    close(val.detach(), gold["loss"], tol=1e-4)
    val.backward()
    for k in ("final_conv.bias", "null_cond_emb", "init_conv.bias"):
        close(sd[k].grad, gold["grad_" + k.replace(".", "_")], tol=2e-3)
    sq = sum(float((v.grad.double() ** 2).sum()) for v in sd.values() if getattr(v, "grad", None) is not None)
    assert abs(sq - gold["grad_sqnorm"]) / gold["grad_sqnorm"] < 2e-3


# ----------------------------------------------------------------------------- one-step generator (sngan.py)

@pytest.mark.parametrize("name", ["g64", "g128", "g192_mono"])
def test_generator_oracle_vs_reference_golden(name):
    from oracle.sngan_ref import generator_forward, make_state_dict, state_dict_shapes
    from tests.golden.sngan_cases import GEN_CASES, GEN_SPECS, gen_inputs
    sname, seed, batch = GEN_CASES[name]
    spec = GEN_SPECS[sname]
    gold = torch.load(os.path.join(G, "sngan.pt"), weights_only=True)[name]
    assert list(state_dict_shapes(spec).keys()) == gold["keys"]          # names and registration order of the reference
    z, y = gen_inputs(spec, batch)
    with torch.no_grad():
        out = generator_forward(make_state_dict(spec, seed), spec, z, y)
    assert out.shape == gold["out"].shape
    assert ((out - gold["out"]).norm() / gold["out"].norm()).item() < 1e-5
