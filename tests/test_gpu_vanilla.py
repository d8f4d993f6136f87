"""GPU parity of the vanilla (GroupNorm) UNet path against the oracle and the reference's
own outputs (tests/golden/vanilla_unet.pt), through the C ABI.

First device run (round 2, B200): 61 of 62 passed unchanged; the one failure was the pair-vs-single tolerance below
(bf16 storage + atomically accumulated statistics: 0.8 % between two batch geometries), so the opt-in gate is gone.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu]

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vanilla_unet.pt"))
DEV = torch.device("cuda" if torch.cuda.is_available() else "cpu")


@pytest.fixture(autouse=True)
def _host_stand_in(monkeypatch):
    """Without a GPU (CCDM_RUN_UNVERIFIED=1 pytest tests/test_gpu_vanilla.py -m gpu in the build container) the same test
    bodies run against the host builds of the kernels (tests/hostpath.py): a dry run of the TESTS themselves."""
    if not torch.cuda.is_available():
        from tests import hostpath
        hostpath.install_engine(monkeypatch)
    yield


def _stream():
    return torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else None


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("cs,groups,hw,with_ss", [((64,), 8, (16, 16), False), ((64, 32), 4, (8, 8), True),
                                                  ((128, 64), 8, (32, 32), True), ((512,), 8, (8, 8), False),
                                                  ((256, 512), 8, (4, 4), True), ((72,), 8, (6, 10), True)])
def test_groupnorm_chain_matches_torch(cs, groups, hw, with_ss):
    """channel_stats -> groupnorm_coef -> affine_act(SiLU) == silu(group_norm(cat(x)) [*(1+scale)+shift]), per source;
    96- and 768-channel concatenations have groups that straddle the two sources."""
    from ccdm_b200 import _lib as L
    lib = L.lib()
    dev = DEV
    g = torch.Generator(device="cpu").manual_seed(5)
    B, (h, w), ctot = 3, hw, sum(cs)
    xs = [(torch.randn(B, h, w, c, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16).contiguous() for c in cs]
    gamma = (1 + 0.2 * torch.randn(ctot, generator=g)).to(dev)
    beta = (0.1 * torch.randn(ctot, generator=g)).to(dev)
    ss = (0.3 * torch.randn(B, 40 + 2 * ctot, generator=g)).to(dev) if with_ss else None
    sums = torch.full((B, 2, ctot), 7.0, device=dev)            # stale contents: zero_first must clear them
    off = 0
    for i, x in enumerate(xs):
        L.check(lib.ccdm_channel_stats(x.data_ptr(), B, h * w, x.shape[3], sums.data_ptr(), ctot, off, int(i == 0), _stream()))
        off += x.shape[3]
    coef = torch.empty(B, 2 * ctot, device=dev)
    L.check(lib.ccdm_groupnorm_coef(sums.data_ptr(), B, ctot, groups, h * w, 1e-5, gamma.data_ptr(), beta.data_ptr(),
                                    L.ptr(ss), ss.shape[1] if with_ss else 0, 40, cs[0], coef.data_ptr(), _stream()))
    outs, off = [], 0
    for x in xs:
        o = torch.empty_like(x)
        L.check(lib.ccdm_affine_act(x.data_ptr(), o.data_ptr(), B * h * w, x.shape[3], h * w, coef.data_ptr(),
                                    coef.shape[1], off, 2, _stream()))
        outs.append(o)
        off += 2 * x.shape[3]
    _sync()
    cat = torch.cat([x.float() for x in xs], -1)
    want_sums = torch.stack([cat.sum((1, 2)), cat.pow(2).sum((1, 2))], 1)
    assert rel(sums, want_sums) < 1e-4
    y = F.group_norm(cat.permute(0, 3, 1, 2), groups, gamma, beta, eps=1e-5)
    if with_ss:
        y = y * (1 + ss[:, 40:40 + ctot, None, None]) + ss[:, 40 + ctot:40 + 2 * ctot, None, None]
    want = F.silu(y).permute(0, 2, 3, 1)
    got = torch.cat([o.float() for o in outs], -1)
    assert rel(got, want) < 6e-3, rel(got, want)


@pytest.mark.parametrize("dh", [16, 32, 64, 128])
@pytest.mark.parametrize("n", [16, 64, 100, 256])
@pytest.mark.parametrize("head_major", [0, 1])
def test_attention_tokens_matches_torch(dh, n, head_major):
    from ccdm_b200 import _lib as L
    dev = DEV
    g = torch.Generator().manual_seed(n + dh)
    B, heads = 3, 2
    hid = heads * dh
    qkv = torch.randn(B, n, 3 * hid, generator=g).to(dev).to(torch.bfloat16).contiguous()
    out = torch.empty(B, n, hid, dtype=torch.bfloat16, device=dev)
    scale = 1.0 / math.sqrt(dh)
    L.check(L.lib().ccdm_attention_tokens(qkv.data_ptr(), out.data_ptr(), B, n, heads, dh, scale, head_major, _stream()))
    _sync()
    f = qkv.float()
    v5 = f.reshape(B, n, heads, 3, dh).permute(0, 1, 3, 2, 4) if head_major else f.reshape(B, n, 3, heads, dh)
    q, k, v = v5[:, :, 0] * scale, v5[:, :, 1], v5[:, :, 2]
    att = torch.einsum("bihd,bjhd->bhij", q, k).softmax(-1)
    want = torch.einsum("bhij,bjhd->bihd", att, v).reshape(B, n, hid)
    assert rel(out, want) < 1e-2


def test_time_features_adm():
    from ccdm_b200 import _lib as L
    from oracle.vanilla_unet_ref import timestep_embedding
    dev = DEV
    t = torch.tensor([0, 1, 17, 500, 999], device=dev)
    for dim in (32, 64, 128):
        out = torch.empty(5, dim, device=dev)
        L.check(L.lib().ccdm_time_features_adm(t.data_ptr(), 5, dim, 10000.0, out.data_ptr(), _stream()))
        _sync()
        assert (out - timestep_embedding(t, dim)).abs().max().item() < 2e-4


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, (64, 64)), (128, 128, (32, 32)), (256, 256, (16, 16)), (32, 48, (8, 8))])
def test_down3x3s2_matches_conv2d(cin, cout, hw):
    from ccdm_b200.backward import conv_forward
    dev = DEV
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(4, hw[0], hw[1], cin, generator=g).to(dev).to(torch.bfloat16).contiguous()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(dev)
    b = (0.1 * torch.randn(cout, generator=g)).to(dev)
    got = conv_forward("down3x3s2", [x], w, b)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel(got, want) < 6e-3


def _build(sname, seed, dev):
    from ccdm_b200.vanilla_unet import VanillaUnet
    from oracle.vanilla_unet_ref import make_state_dict
    from tests.golden.vanilla_cases import V_SPECS
    s = V_SPECS[sname]
    net = VanillaUnet(embed_input_dim=s.embed_input_dim, cond_drop_prob=0.5, in_channels=s.in_channels,
                      model_channels=s.model_channels, num_res_blocks=s.num_res_blocks,
                      attention_resolutions=s.attention_resolutions, channel_mult=s.channel_mult, num_heads=s.num_heads,
                      num_groups=s.num_groups)
    sd = make_state_dict(s, seed)
    net.load_state_dict(sd, strict=True)
    return s, net.to(dev), sd


def test_forward_matches_reference_outputs():
    """Whole network through VanillaEngine vs the reference's own outputs (golden) -- bf16 tolerance 2e-2 (BASELINE.json)."""
    from tests.golden.vanilla_cases import V_BATCH, V_CASES, keep_mask, vanilla_inputs
    dev = DEV
    for name, (sname, seed, mode, kind) in V_CASES.items():
        spec, net, _ = _build(sname, seed, dev)
        net.train(mode == "train")
        x, t, classes = (v.to(dev) for v in vanilla_inputs(sname))
        keep = keep_mask(kind, V_BATCH[sname]).to(dev)
        with torch.no_grad():
            out = net.engine().forward(x, t, classes, keep)
        err = rel(out.cpu(), GOLD[name]["out"])
        print(f"{name}: rel L2 err vs the reference's output {err:.3e}")
        assert err < 2e-2, (name, err)


def test_guidance_matches_reference_outputs():
    from tests.golden.vanilla_cases import V_CFG_CASES, vanilla_inputs
    dev = DEV
    for name, (sname, seed, cs, phi) in V_CFG_CASES.items():
        _, net, _ = _build(sname, seed, dev)
        net.eval()
        x, t, classes = (v.to(dev) for v in vanilla_inputs(sname))
        with torch.no_grad():
            out = net.forward_with_cond_scale(x, t, classes, cond_scale=cs, rescaled_phi=phi)
        err = rel(out.cpu(), GOLD[name]["out"])
        assert err < 3e-2, (name, err)


def test_rc49_config_pair_batch_and_oracle():
    """RC-49 64x64 script configuration at batch 8: the 2B pair batch equals two forwards, and both match the oracle."""
    from oracle.vanilla_unet_ref import vanilla_unet_forward
    dev = DEV
    spec, net, sd = _build("v_rc", 9, dev)
    net.eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 64, 64, generator=g).to(dev)
    t = torch.randint(0, 1000, (8,), generator=g).to(dev)
    classes = torch.rand(8, 128, generator=g).to(dev)
    with torch.no_grad():
        cond, null = net.engine().forward_pair(x, t, classes)
        c1 = net.engine().forward(x, t, classes, None)
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        ref_c = vanilla_unet_forward(sd_d, spec, x, t, classes, torch.ones(8, dtype=torch.bool, device=dev))
        ref_n = vanilla_unet_forward(sd_d, spec, x, t, classes, torch.zeros(8, dtype=torch.bool, device=dev))
    # not bit-equal: statistics are accumulated with atomics (order varies) and the two batch sizes tile differently;
    # measured 0.8 % on B200 (each is within 2e-2 of the fp32 oracle, asserted below)
    assert rel(cond, c1) < 1.5e-2
    assert rel(cond, ref_c) < 2e-2 and rel(null, ref_n) < 2e-2


def test_sampling_loops_match_reference_outputs():
    """VanillaGaussianDiffusion (CUDA-graph step loop) vs images sampled by the reference's own GaussianDiffusion on CPU
    with injected initial noise; eta = 0 cases only are comparable across devices (the CPU and CUDA Philox streams differ),
    so the stochastic cases are checked against the oracle run on the GPU with the same seed instead."""
    import ccdm_b200
    import oracle
    from oracle.vanilla_diffusion_ref import v_ddim_sample, v_ddpm_sample
    from oracle.vanilla_unet_ref import vanilla_forward_with_cond_scale
    from tests.golden.vanilla_cases import V_SAMPLER_CASES, V_SIZES, V_SPECS, sampler_classes
    dev = DEV
    for name, c in V_SAMPLER_CASES.items():
        spec, net, sd = _build(c["spec"], c["seed"], dev)
        net.eval()
        size = V_SIZES[c["spec"]]
        gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=size, timesteps=c["T"], sampling_timesteps=c["S"],
                                                objective=c["objective"], ddim_sampling_eta=c["eta"]).to(dev).eval()
        classes = sampler_classes(c).to(dev)
        shape = (c["B"], spec.in_channels, size, size)
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        guided = lambda x, t, cl, cs, phi: vanilla_forward_with_cond_scale(sd_d, spec, x, t, cl, cs, phi)   # noqa: E731
        sch = oracle.make_schedule(c["T"], "cosine", c["objective"]).to(dev)
        torch.manual_seed(c["rng"])
        if c["kind"] == "ddim":
            img = gd.ddim_sample(classes, shape, cond_scale=c["scale"], rescaled_phi=c["phi"])
        else:
            img = gd.sample(classes, cond_scale=c["scale"], rescaled_phi=c["phi"], preset_sampling_timesteps=c["S"])
        torch.manual_seed(c["rng"])
        if c["kind"] == "ddim":
            ref = v_ddim_sample(sch, guided, classes, shape, sampling_timesteps=c["S"], cond_scale=c["scale"], eta=c["eta"])
        else:
            ref = v_ddpm_sample(sch, guided, classes, shape, steps=c["S"], cond_scale=c["scale"], rescaled_phi=c["phi"])
        mse = ((img - ref) ** 2).mean().item()
        psnr = 10 * math.log10(1.0 / max(mse, 1e-20))
        print(f"{name}: PSNR vs the oracle (same seed, same device) {psnr:.1f} dB")
        # eps objective + random-init weights: in-tolerance per-step error is amplified by sqrt(1/acp - 1)
        assert psnr >= (30.0 if c["objective"] == "pred_noise" else 40.0), (name, psnr)


# ------------------------------------------------------------------------------------------------- training step

@pytest.mark.parametrize("cs,groups,hw,with_ss,act", [((64,), 8, (16, 16), False, 2), ((64, 32), 4, (8, 8), True, 2),
                                                      ((128, 64), 8, (32, 32), True, 2), ((512,), 8, (8, 8), False, 0),
                                                      ((256, 512), 8, (4, 4), True, 2)])
def test_groupnorm_node_backward_matches_autograd(cs, groups, hw, with_ss, act):
    from ccdm_b200.vanilla_train import GroupNormActFn
    dev = DEV
    g = torch.Generator().manual_seed(21)
    B, (h, w), ctot = 3, hw, sum(cs)
    xs = [(torch.randn(B, h, w, c, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16).requires_grad_(True) for c in cs]
    dys = [torch.randn(B, h, w, c, generator=g).to(dev).to(torch.bfloat16) for c in cs]
    gamma = (1 + 0.2 * torch.randn(ctot, generator=g)).to(dev).requires_grad_(True)
    beta = (0.1 * torch.randn(ctot, generator=g)).to(dev).requires_grad_(True)
    ss = (0.3 * torch.randn(B, 2 * ctot, generator=g)).to(dev).requires_grad_(True) if with_ss else None
    outs = GroupNormActFn.apply(groups, 1e-5, act, gamma, beta, ss, *xs)
    torch.autograd.backward(outs, dys)
    xr = torch.cat([x.detach().float() for x in xs], -1).requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    y = F.group_norm(xr.permute(0, 3, 1, 2), groups, gr, br, eps=1e-5)
    if with_ss:
        sr = ss.detach().clone().requires_grad_(True)
        y = y * (1 + sr[:, :ctot, None, None]) + sr[:, ctot:, None, None]
    y = {0: lambda v: v, 2: F.silu}[act](y)
    y.backward(torch.cat([d.float() for d in dys], -1).permute(0, 3, 1, 2))
    assert rel(torch.cat([o.float() for o in outs], -1), y.permute(0, 2, 3, 1)) < 6e-3
    assert rel(torch.cat([x.grad.float() for x in xs], -1), xr.grad) < 1e-2
    assert rel(gamma.grad, gr.grad) < 2e-3 and rel(beta.grad, br.grad) < 2e-3
    if with_ss:
        assert rel(ss.grad, sr.grad) < 2e-3


@pytest.mark.parametrize("dh,n", [(16, 64), (32, 256), (64, 100), (128, 64), (128, 16)])
def test_attention_tokens_backward_matches_autograd(dh, n):
    from ccdm_b200.vanilla_train import AttnTokensFn
    dev = DEV
    g = torch.Generator().manual_seed(n + dh)
    B, heads = 3, 4
    hid = heads * dh
    qkv = torch.randn(B, 1, n, 3 * hid, generator=g).to(dev).to(torch.bfloat16).requires_grad_(True)
    dout = torch.randn(B, 1, n, hid, generator=g).to(dev).to(torch.bfloat16)
    out = AttnTokensFn.apply(qkv, heads)
    out.backward(dout)
    f = qkv.detach().float().reshape(B, n, 3 * hid).requires_grad_(True)
    v5 = f.reshape(B, n, heads, 3, dh).permute(0, 1, 3, 2, 4)
    att = torch.einsum("bihd,bjhd->bhij", v5[:, :, 0] / math.sqrt(dh), v5[:, :, 1]).softmax(-1)
    o = torch.einsum("bhij,bjhd->bihd", att, v5[:, :, 2]).reshape(B, n, hid)
    o.backward(dout.float().reshape(B, n, hid))
    assert rel(out.reshape(B, n, hid), o) < 1e-2
    assert rel(qkv.grad.reshape(B, n, 3 * hid), f.grad) < 1.5e-2


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, (32, 32)), (128, 128, (16, 16)), (32, 48, (8, 8))])
def test_down3x3s2_gradients_match_autograd(cin, cout, hw):
    from ccdm_b200.train import ConvFn
    dev = DEV
    g = torch.Generator().manual_seed(cin)
    x = torch.randn(4, hw[0], hw[1], cin, generator=g).to(dev).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(dev).requires_grad_(True)
    b = (0.1 * torch.randn(cout, generator=g)).to(dev).requires_grad_(True)
    dy = torch.randn(4, hw[0] // 2, hw[1] // 2, cout, generator=g).to(dev).to(torch.bfloat16)
    ConvFn.apply("down3x3s2", w, b, None, x).backward(dy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.detach().to(torch.bfloat16).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    F.conv2d(xr, wr, br, stride=2, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(x.grad.float().permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert rel(w.grad, wr.grad) < 1e-2 and rel(b.grad, br.grad) < 2e-3


def test_training_step_gradients_match_oracle_autograd():
    """Whole vanilla UNet: forward + backward through the CUDA nodes vs fp32 autograd of the oracle (on the GPU)."""
    from ccdm_b200.vanilla_train import vanilla_train_forward
    from oracle.vanilla_unet_ref import vanilla_unet_forward
    from tests.golden.vanilla_cases import V_BATCH, keep_mask, vanilla_inputs
    dev = DEV
    for sname, kind, seed in (("v_tiny", "mixed", 7), ("v_attn", "cond", 8), ("v_rc", "mixed", 9)):
        spec, net, sd = _build(sname, seed, dev)
        net.train()
        x, t, classes = (v.to(dev) for v in vanilla_inputs(sname))
        keep = keep_mask(kind, V_BATCH[sname]).to(dev)
        dout = torch.randn(x.shape, generator=torch.Generator().manual_seed(9)).to(dev)
        out = vanilla_train_forward(net, x, t, classes, keep)
        out.backward(dout)
        sd_g = {k: (v.to(dev).clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k and
                    k != "null_classes_emb" else v.to(dev).clone()) for k, v in sd.items()}
        ref = vanilla_unet_forward(sd_g, spec, x, t, classes, keep, training=True)
        ref.backward(dout)
        names = [n for n, p in net.named_parameters() if p.requires_grad]
        got = torch.cat([net.get_parameter(n).grad.flatten() for n in names]).double()
        want = torch.cat([sd_g[n].grad.flatten() for n in names]).double()
        cos = (got @ want / (got.norm() * want.norm())).item()
        print(f"{sname}: output rel err {rel(out, ref):.3e}, gradient cosine {cos:.5f}")
        assert rel(out, ref) < 2e-2 and cos > 0.999, (sname, cos)


def test_p_losses_match_reference_outputs():
    """VanillaGaussianDiffusion.p_losses + backward vs the loss / gradients of the reference's own p_losses (golden)."""
    import ccdm_b200
    import ccdm_b200.vanilla_unet as VU
    from tests.golden.vanilla_cases import V_BATCH, V_LOSS_CASES, V_LOSS_GRAD_KEYS, V_SIZES, keep_mask, loss_inputs
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vanilla_loss.pt"))
    dev = DEV
    saved = VU.prob_mask_like
    try:
        for name, c in V_LOSS_CASES.items():
            spec, net, _ = _build(c["spec"], c["seed"], dev)
            net.train()
            mask = keep_mask(c["kind"], V_BATCH[c["spec"]]).to(dev)
            VU.prob_mask_like = lambda shape, prob, device, _m=mask: _m.clone()
            gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=V_SIZES[c["spec"]], timesteps=1000,
                                                    objective=c["objective"]).to(dev).train()
            x0, t, classes, noise, weights = (v.to(dev) if v is not None else None for v in loss_inputs(c))
            val = gd.p_losses(x0, t, classes=classes, noise=noise, vicinal_weights=weights)
            val.backward()
            ref = gold[name]
            assert abs(val.item() - ref["loss"].item()) < 2e-2 * abs(ref["loss"].item()), name
            for k in V_LOSS_GRAD_KEYS:
                gg, ww = net.get_parameter(k).grad.cpu(), ref["grad_" + k]
                cos = (gg.flatten().double() @ ww.flatten().double() / (gg.norm().double() * ww.norm().double())).item()
                assert cos > 0.995, (name, k, cos)
    finally:
        VU.prob_mask_like = saved
