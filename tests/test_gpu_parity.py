"""Parity of the CUDA path (through the C ABI) against the oracle and the reference-generated golden vectors.

Floating-point tolerance (BASELINE.json north_star): noise-prediction relative error <= 2e-2 in bf16 and
final-sample PSNR >= 40 dB.  The oracle runs on the same GPU in fp32 with TF32 disabled.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from oracle.unet_ref import UnetSpec, unet_forward, unet_forward_cfg, make_state_dict
from tests.golden.cases import SPECS, SIZES, BATCH, unet_inputs, UNET_CASES, CFG_CASES, SAMPLER_CASES, LOSS_CASES, loss_inputs

BF16_TOL = 2e-2
G = os.path.join(os.path.dirname(__file__), "golden")
load = lambda n: torch.load(os.path.join(G, n + ".pt"), weights_only=True)

RC64 = UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4)


@pytest.fixture(autouse=True)
def _fp32_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def relerr(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


def make_net(spec, seed, dev="cuda", p_drop=0.1):
    import ccdm_b200
    net = ccdm_b200.Unet(dim=spec.dim, embed_input_dim=spec.embed_input_dim, cond_drop_prob=p_drop,
                         dim_mults=spec.dim_mults, in_channels=spec.in_channels, attn_dim_head=spec.attn_dim_head,
                         attn_heads=spec.attn_heads)
    sd = make_state_dict(spec, seed)
    net.load_state_dict(sd, strict=True)
    return net.to(dev), {k: v.to(dev) for k, v in sd.items()}


# ----------------------------------------------------------------------------- UNet forward

@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_vs_reference_golden(name):
    """CUDA path vs the outputs the reference itself produced (tests/golden/unet.pt)."""
    spec_name, seed, mode, p, mask_seed = UNET_CASES[name]
    spec = SPECS[spec_name]
    net, _ = make_net(spec, seed)
    net.train(mode == "train")
    x, t, emb = (v.cuda() for v in unet_inputs(spec_name))
    if 0 < p < 1:
        # the fixture's Bernoulli label-drop mask came from the CPU stream (unet.py:24-31 on a CPU tensor): re-create it from
        # the fixture's seed and hand it to the entry points Unet.forward itself calls with its own draw
        torch.manual_seed(mask_seed)
        mask = (torch.zeros((x.shape[0],)).float().uniform_(0, 1) < (1 - p)).cuda()
        if mode == "train":
            from ccdm_b200.train import unet_train_forward
            y = unet_train_forward(net, x, t, emb, mask)
        else:
            y = net.engine().forward(x, t, emb, mask)
    else:
        y = net(x, t, emb, cond_drop_prob=p)
    gold = load("unet")[name]
    assert relerr(y.cpu(), gold["out"]) < BF16_TOL
    if mode == "train":
        for k, v in gold["bn"].items():
            assert relerr(net.state_dict()[k].cpu(), v) < 1e-4, k


def test_rc64_vs_reference_golden():
    """The headline network at 64x64 against what the REFERENCE's own Unet produced from the same weights and inputs
    (tests/golden/rc64.pt): conditional, null and guided outputs -- closes reference -> oracle -> CUDA at the headline widths."""
    from tests.golden.make_golden_rc64 import SEED, rc64_inputs
    net, _ = make_net(RC64, SEED)
    net.eval()
    x, t, emb = (v.cuda() for v in rc64_inputs())
    gold = load("rc64")
    e_c = relerr(net(x, t, emb, cond_drop_prob=0.0).cpu(), gold["cond"])
    e_n = relerr(net(x, t, emb, cond_drop_prob=1.0).cpu(), gold["null"])
    guided, null = net.forward_with_cond_scale(x, t, emb, cond_scale=1.5, rescaled_phi=0.7)
    e_g = relerr(guided.cpu(), gold["guided"])
    print(f"rc64 vs reference: cond {e_c:.3e} null {e_n:.3e} guided {e_g:.3e}")
    assert max(e_c, e_n, e_g) < BF16_TOL and relerr(null.cpu(), gold["null"]) < BF16_TOL


@pytest.mark.parametrize("B,size", [(2, 64), (5, 64)])
def test_rc64_forward_vs_oracle(B, size):
    """The headline network (RC-49 64x64 widths) at full resolution."""
    torch.manual_seed(0)
    net, sd = make_net(RC64, 7)
    net.eval()
    x = torch.randn(B, 3, size, size, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    emb = oracle.y2h_sinusoidal(torch.rand(B, device="cuda"), 128)
    for p in (0.0, 1.0):
        y = net(x, t, emb, cond_drop_prob=p)
        with torch.no_grad():
            ref = unet_forward(sd, RC64, x, t, emb, cond_drop_prob=p)
        e = relerr(y, ref)
        print(f"rc64 B={B} p={p}: rel err {e:.3e}")
        assert e < BF16_TOL


@pytest.mark.parametrize("spec", [UnetSpec(dim=64, dim_mults=(1, 2, 4, 4, 8, 8)),          # UK128 widths @128
                                  UnetSpec(dim=32, dim_mults=(1, 2, 2, 4), in_channels=1)])  # Cell-200
def test_other_widths_vs_oracle(spec):
    torch.manual_seed(1)
    size = 128 if len(spec.dim_mults) == 6 else 64
    net, sd = make_net(spec, 8)
    net.eval()
    x = torch.randn(2, spec.in_channels, size, size, device="cuda")
    t = torch.tensor([999, 3], device="cuda")
    emb = oracle.y2h_sinusoidal(torch.tensor([0.2, 0.8], device="cuda"), 128)
    y = net(x, t, emb, cond_drop_prob=0.0)
    with torch.no_grad():
        ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.0)
    assert relerr(y, ref) < BF16_TOL


@pytest.mark.parametrize("spec_name", ["wide", "uk64"])
def test_split_norm_path_vs_oracle(spec_name):
    """Channel-split GEMM + standalone norm: 256-wide bottleneck on few pixels, and the UTKFace-64 widths (dim 72:
    channel counts that are not multiples of 64 and a 576-wide level that exceeds one TMEM tile)."""
    spec = SPECS[spec_name]
    net, sd = make_net(spec, 11)
    net.eval()
    x, t, emb = (v.cuda() for v in unet_inputs(spec_name))
    for p in (0.0, 1.0):
        y = net(x, t, emb, cond_drop_prob=p)
        with torch.no_grad():
            ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=p)
        e = relerr(y, ref)
        print(f"{spec_name} p={p}: rel err {e:.3e}")
        assert e < BF16_TOL


def test_mixed_mask_same_rng_stream():
    """0 < p < 1: the in-UNet Bernoulli mask is drawn like the reference (uniform_ < 1-p) from the CUDA stream."""
    spec = SPECS["rc_small"]
    net, sd = make_net(spec, 4)
    net.eval()
    x, t, emb = (v.cuda() for v in unet_inputs("rc_small"))
    torch.manual_seed(123)
    y = net(x, t, emb, cond_drop_prob=0.5)
    torch.manual_seed(123)
    with torch.no_grad():
        ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.5)
    assert relerr(y, ref) < BF16_TOL


# ----------------------------------------------------------------------------- guidance

@pytest.mark.parametrize("name", list(CFG_CASES))
def test_cfg_vs_reference_golden(name):
    spec_name, seed, scale, phi = CFG_CASES[name]
    net, _ = make_net(SPECS[spec_name], seed)
    net.eval()
    x, t, emb = (v.cuda() for v in unet_inputs(spec_name))
    g, n = net.forward_with_cond_scale(x, t, emb, cond_scale=scale, rescaled_phi=phi)
    gold = load("cfg")[name]
    assert relerr(g.cpu(), gold["guided"]) < BF16_TOL
    assert relerr(n.cpu(), gold["null"]) < BF16_TOL


def test_cfg_combine_kernel_exact():
    """The guidance arithmetic alone is fp32/fp64: tight tolerance."""
    from oracle.unet_ref import cfg_combine
    import ccdm_b200
    torch.manual_seed(2)
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2)).cuda()
    c, n = torch.randn(6, 3, 64, 64, device="cuda"), torch.randn(6, 3, 64, 64, device="cuda")
    for scale, phi, rp, kpf in [(1.5, 0.7, True, 0.0), (2.0, 0.0, True, 0.3), (6.0, 0.7, False, 0.0)]:
        got = net.engine().cfg_combine(c, n, scale, phi, rp, kpf)
        assert relerr(got, cfg_combine(c, n, scale, phi, rp, kpf)) < 1e-5


# ----------------------------------------------------------------------------- samplers

def _diffusion(spec, seed, size, **kw):
    import ccdm_b200
    net, sd = make_net(spec, seed)
    gd = ccdm_b200.GaussianDiffusion(torch.nn.DataParallel(net, device_ids=[0]), image_size=size, **kw).cuda().eval()
    return gd, sd


@pytest.mark.parametrize("name", [n for n, c in SAMPLER_CASES.items() if c["eta"] == 0 and c["kind"] == "ddim"])
def test_ddim_vs_reference_golden(name):
    """Initial noise re-created from the fixture's CPU seed; eta == 0 so no further draws matter."""
    c = SAMPLER_CASES[name]
    spec = SPECS[c["spec"]]
    fn_y2cov = (lambda y: oracle.y2cov_sinusoidal(y, spec.in_channels * c["size"] ** 2)) if c["use_Hy"] else None
    gd, _ = _diffusion(spec, c["seed"], c["size"], use_Hy=c["use_Hy"], fn_y2cov=fn_y2cov, timesteps=c["T"],
                       sampling_timesteps=c["S"], objective=c["objective"], ddim_sampling_eta=c["eta"])
    labels = torch.linspace(0.05, 0.95, c["B"])
    shape = (c["B"], spec.in_channels, c["size"], c["size"])
    torch.manual_seed(c["rng"])
    x_init = torch.randn(shape)
    img = gd.ddim_sample(labels_emb=oracle.y2h_sinusoidal(labels, 128).cuda(), labels=labels.cuda(), shape=shape,
                         cond_scale=c["scale"], x_init=x_init)
    gold = load("sampler")[name]["img"]
    p = psnr(img.cpu(), gold)
    print(f"{name}: PSNR vs reference {p:.1f} dB")
    assert p >= 40.0


@pytest.mark.parametrize("name", [n for n, c in SAMPLER_CASES.items() if c["eta"] != 0 or c["kind"] == "ddpm"])
def test_stochastic_samplers_vs_reference_golden(name):
    """The reference's DDPM and eta > 0 DDIM fixtures replayed ON THE DEVICE: the fixture's CPU random stream (initial noise,
    then one randn_like per step, diffusion.py:365,383,430,452) is re-created from its seed and injected through the
    ``x_init`` / ``noise_fn`` hooks, so the stochastic samplers are compared sample for sample."""
    c = SAMPLER_CASES[name]
    spec = SPECS[c["spec"]]
    fn_y2cov = (lambda y: oracle.y2cov_sinusoidal(y, spec.in_channels * c["size"] ** 2)) if c["use_Hy"] else None
    gd, _ = _diffusion(spec, c["seed"], c["size"], use_Hy=c["use_Hy"], fn_y2cov=fn_y2cov, timesteps=c["T"],
                       sampling_timesteps=c["S"], objective=c["objective"], ddim_sampling_eta=c["eta"])
    labels = torch.linspace(0.05, 0.95, c["B"])
    shape = (c["B"], spec.in_channels, c["size"], c["size"])
    torch.manual_seed(c["rng"])
    x_init = torch.randn(shape)
    noise_fn = lambda: torch.randn(shape).cuda()             # CPU generator, consumed in the reference's order
    emb = oracle.y2h_sinusoidal(labels, 128).cuda()
    if c["kind"] == "ddim":
        img = gd.ddim_sample(labels_emb=emb, labels=labels.cuda(), shape=shape, cond_scale=c["scale"], x_init=x_init,
                             noise_fn=noise_fn)
    else:
        img = gd.p_sample_loop(labels_emb=emb, labels=labels.cuda(), shape=shape, cond_scale=c["scale"], x_init=x_init,
                               noise_fn=noise_fn)
    gold = load("sampler")[name]["img"]
    p = psnr(img.cpu(), gold)
    print(f"{name}: PSNR vs reference {p:.1f} dB")
    # Same floors as the CPU replay of these fixtures (tests/test_sampling_hostpath.py): the DDPM fixtures run the LAST steps
    # of the 1000-step chain from pure noise, where the bf16 network error lands in the sample unattenuated (30 dB); the eps
    # objective with random-init weights also amplifies it by sqrt(1/acp - 1) (27 dB)
    assert p >= (27.0 if c["objective"] == "pred_noise" else 30.0)


@pytest.mark.parametrize("objective,use_Hy", [("pred_x0", False), ("pred_noise", False), ("pred_x0", True)])
def test_rc64_ddim_vs_oracle(objective, use_Hy):
    """RC-49 network, DDIM with guidance: teacher-forced per-step eps error and free-running PSNR."""
    B, S, size = 4, 20, 64
    fn_y2cov = (lambda y: oracle.y2cov_sinusoidal(y, 3 * size * size)) if use_Hy else None
    gd, sd = _diffusion(RC64, 7, size, use_Hy=use_Hy, fn_y2cov=fn_y2cov, timesteps=1000, sampling_timesteps=S,
                        objective=objective)
    labels = torch.linspace(0.1, 0.9, B, device="cuda")
    emb = oracle.y2h_sinusoidal(labels, 128)
    shape = (B, 3, size, size)
    sch = oracle.make_schedule(1000, "cosine", objective).to("cuda")
    net_o = lambda x, t, e, p: unet_forward(sd, RC64, x, t, e, cond_drop_prob=p)
    cov = torch.exp(-fn_y2cov(labels).view(shape)) if use_Hy else None

    torch.manual_seed(5)
    trace_o = []
    ref = oracle.ddim_sample(sch, net_o, emb, shape, sampling_timesteps=S, cond_scale=1.5, init_cov=cov, trace=trace_o)
    torch.manual_seed(5)
    trace = []
    img = gd.ddim_sample(labels_emb=emb, labels=labels, shape=shape, cond_scale=1.5, trace=trace)
    p = psnr(img, ref)
    print(f"rc64 {objective} Hy={use_Hy}: free-running PSNR {p:.1f} dB; step-0 eps err {relerr(trace[0][0], trace_o[0][0]):.3e}")
    assert relerr(trace[0][0], trace_o[0][0]) < BF16_TOL            # identical x_T -> pure per-step error
    # pred_x0 is the objective of every CCDM config (incl. the headline one) and must meet the 40 dB bar.  With the
    # eps objective and *random-init* weights the x0 reconstruction multiplies the (in-tolerance) per-step eps error
    # by sqrt(1/acp - 1) <= 150 at the start of the chain, so the free-running trajectories separate further.
    assert p >= (40.0 if objective == "pred_x0" else 30.0)

    # teacher-forced: the oracle's own states through model_predictions
    pairs = oracle.diffusion_ref.ddim_time_pairs(1000, S)
    torch.manual_seed(5)
    x = torch.randn(shape, device="cuda")
    if use_Hy:
        x = x * torch.sqrt(cov)
    worst, worst_derived = 0.0, 0.0
    for (tm, tn) in pairs[:: max(1, S // 5)]:
        tt = torch.full((B,), tm, device="cuda", dtype=torch.long)
        eps_o, x0_o = oracle.model_predictions(sch, net_o, x, tt, emb, 1.5, 0.7, clip_x_start=False)
        mp = gd.model_predictions(x, tt, emb, cond_scale=1.5, rescaled_phi=0.7, clip_x_start=False)
        # the network's own (guided) prediction carries the bf16 tolerance; the quantity derived from it through
        # 1/sqrt(1/acp - 1) (eps from x0, or x0 from eps) amplifies that error at the ends of the schedule
        native = (relerr(mp.pred_x_start, x0_o) if objective == "pred_x0" else relerr(mp.pred_noise, eps_o))
        derived = (relerr(mp.pred_noise, eps_o) if objective == "pred_x0" else relerr(mp.pred_x_start, x0_o))
        worst, worst_derived = max(worst, native), max(worst_derived, derived)
        x = oracle.q_sample(sch, x0_o.clamp(-1, 1), torch.full((B,), max(tn, 0), device="cuda"), torch.randn_like(x))
    print(f"   teacher-forced worst rel err: network prediction {worst:.3e}, derived {worst_derived:.3e}")
    assert worst < BF16_TOL
    assert worst_derived < 5 * BF16_TOL


def test_ddpm_vs_oracle():
    B, S, size = 2, 6, 32
    spec = SPECS["rc_small"]
    gd, sd = _diffusion(spec, 4, size, timesteps=1000, sampling_timesteps=S, objective="pred_noise")
    labels = torch.tensor([0.3, 0.7], device="cuda")
    emb = oracle.y2h_sinusoidal(labels, 128)
    sch = oracle.make_schedule(1000, "cosine", "pred_noise").to("cuda")
    net_o = lambda x, t, e, p: unet_forward(sd, spec, x, t, e, cond_drop_prob=p)
    torch.manual_seed(9)
    ref = oracle.ddpm_sample(sch, net_o, emb, (B, 3, size, size), sampling_timesteps=S, cond_scale=2.0)
    torch.manual_seed(9)
    img = gd.sample(labels_emb=emb, labels=labels, cond_scale=2.0)      # same seed: same randn stream, step by step
    p = psnr(img, ref)
    print(f"ddpm PSNR {p:.1f} dB")
    assert p >= 40.0


def test_sampling_is_deterministic_and_graph_replay_is_stable():
    gd, _ = _diffusion(SPECS["rc_small"], 4, 32, timesteps=1000, sampling_timesteps=8, objective="pred_x0")
    labels = torch.linspace(0, 1, 3, device="cuda")
    emb = oracle.y2h_sinusoidal(labels, 128)
    outs = []
    for _ in range(3):
        torch.manual_seed(3)
        outs.append(gd.ddim_sample(labels_emb=emb, labels=labels, shape=(3, 3, 32, 32), cond_scale=1.5))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert 0.0 <= outs[0].min() and outs[0].max() <= 1.0               # pred_x0 + clip keeps samples in range


# ----------------------------------------------------------------------------- training loss (forward)

@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_loss_vs_oracle(name):
    c = LOSS_CASES[name]
    spec = SPECS[c["spec"]]
    import ccdm_b200
    net, sd = make_net(spec, c["seed"], p_drop=c["p_drop"])
    n_el = spec.in_channels * c["size"] ** 2
    fn_y2cov = lambda y: oracle.y2cov_sinusoidal(y, n_el)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=c["size"], use_Hy=c["use_Hy"],
                                     fn_y2cov=fn_y2cov if c["use_Hy"] else None, cond_drop_prob=c["p_drop"],
                                     timesteps=1000, objective=c["objective"], vicinity_type=c["vic"]).cuda().train()
    img, labels, emb_in = (v.cuda() for v in loss_inputs(c))
    emb = oracle.y2h_sinusoidal(emb_in, 128)
    kw = {}
    if c["vic"] is not None:
        kw = dict(vicinity_type=c["vic"], kappa=c["kappa"], num_projections=c.get("nproj", 1), vector_type="gaussian")
    vw = None if c["vic"] is None else torch.ones(c["B"], device="cuda")
    bn_before = {k: v.clone() for k, v in net.state_dict().items() if "running_" in k}

    torch.manual_seed(c["rng"])
    with torch.no_grad():
        got = gd(img, labels_emb=emb, labels=labels, vicinal_weights=vw, **kw)

    sch = oracle.make_schedule(1000, "cosine", c["objective"]).to("cuda")
    sd_o = dict(sd)
    sd_o.update(bn_before)
    torch.manual_seed(c["rng"])
    t = torch.randint(0, 1000, (c["B"],), device="cuda").long()
    net_o = lambda x, tt, e: unet_forward(sd_o, spec, x, tt, e, cond_drop_prob=c["p_drop"], training=True)
    with torch.no_grad():
        ref = oracle.p_losses(sch, net_o, img * 2 - 1, t, labels=labels, labels_emb=emb, cond_drop_prob=c["p_drop"],
                              use_Hy=c["use_Hy"], fn_y2cov=fn_y2cov, vicinal_weights=vw, **{k: v for k, v in kw.items() if k != "vector_type"})
    e = abs(got.item() - ref.item()) / abs(ref.item())
    print(f"{name}: loss {got.item():.6f} vs oracle {ref.item():.6f} (rel {e:.2e})")
    assert e < BF16_TOL


def test_trainer_sample_given_labels_matches_direct_call():
    """Hot caller (trainer.py:782-869): uint8 images for given labels through the EMA copy."""
    import numpy as np
    import ccdm_b200
    spec = SPECS["rc_small"]
    net, _ = make_net(spec, 4)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=1000, sampling_timesteps=5, objective="pred_x0").cuda()
    tr = ccdm_b200.Trainer("RC-49", gd, train_images=None, train_labels=None,
                           vicinal_params={"kernel_sigma": 0.05, "kappa": 0.02, "nonzero_soft_weight_threshold": 1e-3},
                           train_batch_size=16, results_folder="/tmp/ccdm_b200_results", ema_update_after_step=0)
    fn_y2h = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=torch.device("cuda")).fn_y2h
    labels = np.linspace(0, 1, 6).astype(np.float32)
    torch.manual_seed(1)
    imgs, lab = tr.sample_given_labels(labels, fn_y2h, batch_size=3, sampler="ddim", cond_scale=1.5)
    assert imgs.shape == (6, 3, 16, 16) and imgs.dtype == np.uint8 and lab is labels
    torch.manual_seed(1)
    tr.ema.ema_model.eval()
    y = torch.from_numpy(labels[:3]).cuda()
    direct = tr.ema.ema_model.ddim_sample(labels_emb=fn_y2h(y), labels=y, shape=(3, 3, 16, 16), cond_scale=1.5)
    assert np.array_equal(imgs[:3], (direct.clip(0, 1) * 255.0).type(torch.uint8).cpu().numpy())


# ----------------------------------------------------------------------------- full-size, size-independent properties

def test_full_size_sampling_is_batch_shard_invariant():
    """BASELINE workload shape (RC-49 64x64, batch 200 per GPU): samples are independent of how the batch is split,
    which is what the multi-GPU sharding (no data-path collective) relies on.  Few DDIM steps keep it short."""
    B, S = 200, 3
    gd, _ = _diffusion(RC64, 7, 64, timesteps=1000, sampling_timesteps=S, objective="pred_x0")
    labels = torch.linspace(0, 1, B, device="cuda")
    emb = oracle.y2h_sinusoidal(labels, 128)
    torch.manual_seed(21)
    x_init = torch.randn(B, 3, 64, 64, device="cuda")
    full = gd.ddim_sample(labels_emb=emb, labels=labels, shape=(B, 3, 64, 64), cond_scale=1.5, x_init=x_init)
    parts = [gd.ddim_sample(labels_emb=emb[lo:lo + 100], labels=labels[lo:lo + 100], shape=(100, 3, 64, 64),
                            cond_scale=1.5, x_init=x_init[lo:lo + 100]) for lo in (0, 100)]
    assert torch.equal(full, torch.cat(parts))                      # bit-identical: every sample is its own problem
    assert full.min() >= 0.0 and full.max() <= 1.0 and torch.isfinite(full).all()


def test_pair_batch_equals_two_forwards_on_device():
    """The guided sampler evaluates cond + null as ONE 2B batch; in eval mode that must equal two forwards."""
    net, _ = make_net(RC64, 7)
    net.eval()
    B = 6
    torch.manual_seed(4)
    x = torch.randn(B, 3, 64, 64, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    emb = oracle.y2h_sinusoidal(torch.rand(B, device="cuda"), 128)
    cond = net(x, t, emb, cond_drop_prob=0.0)
    null = net(x, t, emb, cond_drop_prob=1.0)
    c2, n2 = net.engine().forward_pair(x, t, emb)
    assert torch.equal(cond, c2) and torch.equal(null, n2)


@pytest.mark.parametrize("name,spec,size", [
    ("SA128", UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 4, 8)), 128),          # BASELINE config 4 widths
    ("UK192", UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 4, 8, 8)), 192),       # BASELINE config 5 widths (3x3 bottleneck)
])
def test_large_configs_vs_oracle(name, spec, size):
    torch.manual_seed(2)
    net, sd = make_net(spec, 9)
    net.eval()
    x = torch.randn(1, 3, size, size, device="cuda")
    t = torch.tensor([400], device="cuda")
    emb = oracle.y2h_sinusoidal(torch.tensor([0.5], device="cuda"), 128)
    y = net(x, t, emb, cond_drop_prob=0.0)
    with torch.no_grad():
        ref = unet_forward(sd, spec, x, t, emb, cond_drop_prob=0.0)
    e = relerr(y, ref)
    print(f"{name} {size}x{size}: rel err {e:.3e}")
    assert e < BF16_TOL
