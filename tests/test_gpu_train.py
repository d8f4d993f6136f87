"""GPU parity of the training path (ccdm_b200/train.py): every autograd node, then the whole UNet and the whole
training loss, against the oracle (oracle/unet_ref.py, fp32, TF32 off) differentiated by torch autograd.

Tolerances: the CUDA path keeps bf16 activations and bf16 gradients between nodes, so per-tensor gradient errors of
1-3 % are the bf16 floor for a 60-layer network; the tests ask for <= 5e-2 relative Frobenius error on each
parameter gradient of the single-node tests, and on the whole network a cosine similarity >= 0.99 over all parameter
gradients together plus <= 1e-1 per tensor (gradients of tiny tensors are noisier)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import oracle
from oracle.unet_ref import unet_forward, make_state_dict
from oracle import unet_ref as R
from tests.golden.cases import SPECS, SIZES, BATCH, unet_inputs, LOSS_CASES, loss_inputs
from tests.test_gpu_parity import make_net, relerr


@pytest.fixture(autouse=True)
def _fp32_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def cos(a, b):
    return (a.double().flatten() @ b.double().flatten() / (a.double().norm() * b.double().norm() + 1e-30)).item()


# ----------------------------------------------------------------------------- single nodes

@pytest.mark.parametrize("hw,b", [(16, 2), (8, 3), (4, 5), (32, 2)])
def test_linear_attention_core(hw, b):
    from ccdm_b200.train import LinAttnCoreFn
    g = torch.Generator().manual_seed(hw)
    qkv = (torch.randn(b, hw, hw, 384, generator=g) * 1.5).bfloat16().cuda()
    dout = torch.randn(b, hw, hw, 128, generator=g).bfloat16().cuda()
    scale = 32 ** -0.5

    q1 = qkv.clone().requires_grad_(True)
    out = LinAttnCoreFn.apply(q1, scale)
    out.backward(dout)

    x = nchw(qkv.float()).requires_grad_(True)
    v5 = x.view(b, 3, 4, 32, hw * hw)
    q, k, v = v5[:, 0], v5[:, 1], v5[:, 2]
    q = q.softmax(dim=-2) * scale
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, 128, hw, hw)
    o.backward(nchw(dout.float()))
    assert relerr(out.float(), nhwc(o.detach())) < 2e-2
    got, ref = q1.grad.float(), nhwc(x.grad)
    for name, sl in (("dq", slice(0, 128)), ("dk", slice(128, 256)), ("dv", slice(256, 384))):
        assert relerr(got[..., sl], ref[..., sl]) < 3e-2, name


@pytest.mark.parametrize("hw,b,heads,dh", [(4, 3, 4, 32), (8, 2, 4, 32), (4, 2, 2, 16), (16, 1, 4, 32)])
def test_attention_core(hw, b, heads, dh):
    from ccdm_b200.train import AttnCoreFn
    g = torch.Generator().manual_seed(hw + dh)
    hid = heads * dh
    qkv = torch.randn(b, hw, hw, 3 * hid, generator=g).bfloat16().cuda()
    dout = torch.randn(b, hw, hw, hid, generator=g).bfloat16().cuda()
    scale = dh ** -0.5
    q1 = qkv.clone().requires_grad_(True)
    out = AttnCoreFn.apply(q1, heads, dh, scale)
    out.backward(dout)

    x = nchw(qkv.float()).requires_grad_(True)
    v5 = x.view(b, 3, heads, dh, hw * hw)
    q, k, v = v5[:, 0] * scale, v5[:, 1], v5[:, 2]
    att = torch.einsum("bhdi,bhdj->bhij", q, k).softmax(dim=-1)
    o = torch.einsum("bhij,bhdj->bhid", att, v).permute(0, 1, 3, 2).reshape(b, hid, hw, hw)
    o.backward(nchw(dout.float()))
    assert relerr(out.float(), nhwc(o.detach())) < 2e-2
    assert relerr(q1.grad.float(), nhwc(x.grad)) < 2e-2


def test_stem_and_head():
    from ccdm_b200.train import StemFn, HeadFn
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 16, 16, generator=g).cuda()
    w = (torch.randn(64, 3, 7, 7, generator=g) / 12).cuda().requires_grad_(True)
    bias = (0.1 * torch.randn(64, generator=g)).cuda().requires_grad_(True)
    dz = torch.randn(3, 16, 16, 64, generator=g).bfloat16().cuda()
    out = StemFn.apply(x, w, bias)
    out.backward(dz)
    wr = w.detach().clone().requires_grad_(True)
    br = bias.detach().clone().requires_grad_(True)
    ref = F.conv2d(x, wr, br, padding=3)
    ref.backward(nchw(dz.float()))
    assert relerr(out.float(), nhwc(ref.detach())) < 2e-2
    assert relerr(w.grad, wr.grad) < 2e-2          # x is rounded to bf16 inside the im2row tensor
    assert relerr(bias.grad, br.grad) < 1e-3

    h = torch.randn(3, 16, 16, 64, generator=g).bfloat16().cuda().requires_grad_(True)
    w2 = (torch.randn(3, 64, 1, 1, generator=g) / 8).cuda().requires_grad_(True)
    b2 = (0.1 * torch.randn(3, generator=g)).cuda().requires_grad_(True)
    dout = torch.randn(3, 3, 16, 16, generator=g).cuda()
    y = HeadFn.apply(h, w2, b2)
    y.backward(dout)
    hr = nchw(h.detach().float()).requires_grad_(True)
    w2r, b2r = w2.detach().clone().requires_grad_(True), b2.detach().clone().requires_grad_(True)
    yr = F.conv2d(hr, w2r, b2r)
    yr.backward(dout)
    assert relerr(y, yr.detach()) < 1e-3
    assert relerr(h.grad.float(), nhwc(hr.grad)) < 1e-2
    assert relerr(w2.grad, w2r.grad) < 1e-3
    assert relerr(b2.grad, b2r.grad) < 1e-3


# ----------------------------------------------------------------------------- whole UNet

def _oracle_grads(spec, sd, x, t, emb, keep_mask, p_drop, dout):
    sd_o = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
            for k, v in sd.items()}
    y = unet_forward(sd_o, spec, x, t, emb, cond_drop_prob=p_drop, training=True, keep_mask=keep_mask)
    y.backward(dout)
    return y.detach(), {k: v.grad for k, v in sd_o.items() if isinstance(v, torch.Tensor) and v.requires_grad}


@pytest.mark.parametrize("spec_name,seed", [("tiny", 1), ("rc_small", 4), ("cell", 3), ("wide", 5), ("uk64", 6)])
def test_unet_backward_vs_oracle(spec_name, seed):
    spec = SPECS[spec_name]
    net, sd = make_net(spec, seed, p_drop=0.3)
    net.train()
    x, t, emb = (v.cuda() for v in unet_inputs(spec_name))
    b = x.shape[0]
    keep = torch.tensor([True, False, True, True, False][:b], device="cuda")
    dout = torch.randn(x.shape[0], spec.out_channels, *x.shape[2:], generator=torch.Generator().manual_seed(9)).cuda()

    from ccdm_b200.train import unet_train_forward
    if spec_name in ("rc_small", "wide"):
        # gradient buffers already attached (FusedAdam's flat views, or a second micro-batch): the kernels then
        # accumulate into .grad in place instead of handing tensors to autograd
        for p in net.parameters():
            p.grad = torch.zeros_like(p)
    y = unet_train_forward(net, x, t, emb, keep)
    y.backward(dout)
    y_ref, g_ref = _oracle_grads(spec, sd, x, t, emb, keep, 0.3, dout)
    assert relerr(y.detach(), y_ref) < 2e-2

    got_all, ref_all, worst = [], [], (0.0, "")
    for name, p in net.named_parameters():
        assert p.grad is not None, f"no gradient for {name}"
        r = g_ref[name]
        if r is None or r.norm() == 0 or name in ("cond_mlp_1.0.bias", "cond_mlp_2.0.bias"):
            continue          # a bias in front of a batch-statistics BatchNorm1d has an exactly-zero gradient: both are noise
        e = relerr(p.grad, r)
        if e > worst[0]:
            worst = (e, name)
        got_all.append(p.grad.flatten())
        ref_all.append(r.flatten())
    c = cos(torch.cat(got_all), torch.cat(ref_all))
    print(f"{spec_name}: output err {relerr(y.detach(), y_ref):.2e}, grad cosine {c:.5f}, worst tensor {worst[1]} {worst[0]:.2e}")
    assert c > 0.99
    assert worst[0] < 1e-1, worst


def test_rc64_backward_full_size_finite_and_linear():
    """RC-49 64x64 widths at full resolution: gradients finite, and backward is linear in the output gradient."""
    from tests.test_gpu_parity import RC64
    from ccdm_b200.train import unet_train_forward
    net, _ = make_net(RC64, 7)
    net.train()
    torch.manual_seed(0)
    x = torch.randn(4, 3, 64, 64, device="cuda")
    t = torch.randint(0, 1000, (4,), device="cuda")
    emb = oracle.y2h_sinusoidal(torch.rand(4, device="cuda"), 128)
    keep = torch.tensor([True, True, False, True], device="cuda")
    dout = torch.randn(4, 3, 64, 64, device="cuda")
    for bn in (net.cond_mlp_1[1], net.cond_mlp_2[1]):
        bn.momentum = 0.0                     # keep the running statistics fixed between the two passes
    grads = []
    for s in (1.0, 2.0):
        net.zero_grad(set_to_none=True)
        unet_train_forward(net, x, t, emb, keep).backward(dout * s)
        grads.append(torch.cat([p.grad.flatten() for p in net.parameters()]))
    assert torch.isfinite(grads[0]).all()
    assert relerr(grads[1], 2 * grads[0]) < 2e-2


# ----------------------------------------------------------------------------- training loss + optimizer step

@pytest.mark.parametrize("name", ["hv_x0_Hy", "sv_eps", "shv_scalar_v", "hv_rc_small"])
def test_loss_backward_vs_oracle(name):
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
    kw = dict(vicinity_type=c["vic"], kappa=c["kappa"], num_projections=c.get("nproj", 1), vector_type="gaussian")
    vw = torch.ones(c["B"], device="cuda")
    bn_before = {k: v.clone() for k, v in net.state_dict().items() if "running_" in k}

    torch.manual_seed(c["rng"])
    loss = gd(img, labels_emb=emb, labels=labels, vicinal_weights=vw, **kw)
    loss.backward()

    sch = oracle.make_schedule(1000, "cosine", c["objective"]).to("cuda")
    sd_o = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
            for k, v in sd.items()}
    sd_o.update(bn_before)
    torch.manual_seed(c["rng"])
    t = torch.randint(0, 1000, (c["B"],), device="cuda").long()
    net_o = lambda x, tt, e: unet_forward(sd_o, spec, x, tt, e, cond_drop_prob=c["p_drop"], training=True)
    ref = oracle.p_losses(sch, net_o, img * 2 - 1, t, labels=labels, labels_emb=emb, cond_drop_prob=c["p_drop"],
                          use_Hy=c["use_Hy"], fn_y2cov=fn_y2cov, vicinal_weights=vw,
                          **{k: v for k, v in kw.items() if k != "vector_type"})
    ref.backward()
    assert abs(loss.item() - ref.item()) / abs(ref.item()) < 2e-2
    got = torch.cat([p.grad.flatten() for _, p in net.named_parameters()])
    want = torch.cat([sd_o[n].grad.flatten() if sd_o[n].grad is not None else torch.zeros_like(p).flatten()
                      for n, p in net.named_parameters()])
    c_ = cos(got, want)
    print(f"{name}: loss {loss.item():.5f} vs {ref.item():.5f}; grad cosine {c_:.5f}, rel {relerr(got, want):.2e}")
    assert c_ > 0.99
    assert relerr(got, want) < 1e-1


def test_adam_steps_reduce_loss():
    """The reference's optimizer loop (trainer.py:617-734 condensed): torch Adam on our parameters, fixed batch."""
    import ccdm_b200
    spec = SPECS["rc_small"]
    net, _ = make_net(spec, 4, p_drop=0.1)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=1000, objective="pred_x0", cond_drop_prob=0.1).cuda().train()
    opt = torch.optim.Adam(gd.parameters(), lr=1e-3, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(0)
    img = torch.rand(8, 3, 16, 16, generator=g).cuda()
    labels = torch.rand(8, generator=g).cuda()
    emb = oracle.y2h_sinusoidal(labels, 128)
    losses = []
    for step in range(30):
        torch.manual_seed(100)                 # same timesteps / noise every step: the loss must go down
        loss = gd(img, labels_emb=emb, labels=labels, vicinal_weights=torch.ones(8, device="cuda"), vicinity_type="hv",
                  kappa=0.1)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", [round(v, 4) for v in losses[::5]])
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < 0.7 * losses[0]


def test_graphed_train_step():
    """The whole optimizer step replayed from a CUDA graph: fresh random draws per replay, loss goes down on a fixed
    batch, parameters stay finite, and it agrees with the eager path on what one step does to the loss scale."""
    import ccdm_b200
    from ccdm_b200.train_graph import GraphedTrainStep
    spec = SPECS["rc_small"]
    net, _ = make_net(spec, 4, p_drop=0.1)
    n_el = 3 * 16 * 16
    gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=1000, objective="pred_x0", cond_drop_prob=0.1,
                                     use_Hy=True, fn_y2cov=lambda y: cov_emb(y)).cuda().train()
    freq = torch.exp(-math.log(10000) * torch.arange(n_el // 2, dtype=torch.float32, device="cuda") / (n_el // 2))

    def cov_emb(y):            # sinusoidal covariance embedding with no host-side tensors (legal inside a capture)
        a = y.reshape(-1)[:, None] * freq[None]
        return (torch.cat([torch.cos(a), torch.sin(a)], -1) + 1) / 2

    opt = torch.optim.Adam(gd.parameters(), lr=1e-3, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(0)
    img = torch.rand(16, 3, 16, 16, generator=g).cuda()
    labels = torch.rand(16, generator=g).cuda()
    emb = oracle.y2h_sinusoidal(labels, 128)
    step = GraphedTrainStep(gd, opt, img, labels, emb, loss_kwargs=dict(vicinity_type="hv", kappa=0.1))
    losses = [step(img, labels, emb).item() for _ in range(40)]
    print("graphed losses", [round(v, 4) for v in losses[::8]])
    assert all(math.isfinite(v) for v in losses)
    assert len(set(round(v, 6) for v in losses[:5])) > 1          # new timesteps / noise on every replay
    assert sum(losses[-10:]) < 0.8 * sum(losses[:10])
    assert all(torch.isfinite(p).all() for p in gd.parameters())


def test_fused_adam_matches_torch_adam():
    """FusedAdam (clip + Adam over flat buffers) vs clip_grad_norm_ + torch.optim.Adam on identical gradients, 5 steps."""
    from ccdm_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 3, 7, 7), (64,), (1, 64, 1, 1), (128, 64, 3, 3), (7,), (300, 5)]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdam(pa, lr=3e-3, betas=(0.9, 0.99), max_grad_norm=1.0)
    ob = torch.optim.Adam(pb, lr=3e-3, betas=(0.9, 0.99))
    for step in range(5):
        gs = [torch.randn_like(p) * (10.0 if step % 2 == 0 else 0.01) for p in pa]      # clipped and unclipped steps
        oa.zero_grad()
        for p, g in zip(pa, gs):
            p.grad.add_(g)                     # what autograd's AccumulateGrad does with an attached view
        for p, g in zip(pb, gs):
            p.grad = g.clone()
        norm_b = torch.nn.utils.clip_grad_norm_(pb, 1.0)
        oa.step()
        ob.step()
        assert abs(oa.grad_norm.item() - norm_b.item()) / norm_b.item() < 1e-5
        for a, b in zip(pa, pb):
            assert relerr(a.detach(), b.detach()) < 1e-6
    sd = oa.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5
    assert relerr(sd["state"][3]["exp_avg"], ob.state_dict()["state"][3]["exp_avg"]) < 1e-6


def test_ema_multi_lerp_matches_foreach():
    import ccdm_b200
    from ccdm_b200.ema import EMA
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).cuda()
    ema = EMA(net, beta=0.99, update_after_step=0, update_every=1)
    ref = [p.detach().clone() for p in net.parameters()]
    for step in range(6):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        s = int(ema.step.item())
        ema.update()
        if s == 0:
            ref = [p.detach().clone() for p in net.parameters()]          # first call copies
        else:
            if s == 1:
                ref = [p.detach().clone() for p in net.parameters()]      # initted copy, then lerp with itself
            w = 1.0 - ema.get_current_decay()
            ref = [r.lerp(p.detach(), w) for r, p in zip(ref, net.parameters())]
    for e, r in zip(ema.ema_model.parameters(), ref):
        assert relerr(e, r) < 1e-5


def test_trainer_train_runs_optimizer_steps(tmp_path):
    """Trainer.train (trainer.py:537-780): vicinal batch construction on the device, loss, backward through the CUDA
    graph of autograd nodes, fused clip + Adam, EMA -- three optimizer steps on a toy data set."""
    import numpy as np
    import ccdm_b200
    from ccdm_b200.optim import FusedAdam
    spec = SPECS["rc_small"]
    net, _ = make_net(spec, 4, p_drop=0.1)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=1000, sampling_timesteps=5, objective="pred_x0",
                                     cond_drop_prob=0.1, vicinity_type="hv").cuda()
    rng = np.random.RandomState(0)
    images = rng.randint(0, 256, size=(64, 3, 16, 16)).astype(np.uint8)
    labels = rng.rand(64).astype(np.float32)
    tr = ccdm_b200.Trainer("RC-49", gd, train_images=images, train_labels=labels,
                           vicinal_params={"kernel_sigma": 0.05, "kappa": 0.1, "nonzero_soft_weight_threshold": 1e-3},
                           train_batch_size=16, train_num_steps=3, results_folder=str(tmp_path), vicinity_type="hv",
                           ema_update_after_step=0, ema_update_every=1)
    assert isinstance(tr.opt, FusedAdam)
    fn_y2h = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=torch.device("cuda")).fn_y2h
    before = torch.cat([p.detach().flatten().clone() for p in gd.parameters()])
    gd.train()
    tr.train(fn_y2h)
    after = torch.cat([p.detach().flatten() for p in gd.parameters()])
    assert tr.step == 3
    assert torch.isfinite(after).all() and (after - before).abs().max() > 0
    assert float(tr.opt.step_count.item()) == 3.0
    ema_p = torch.cat([p.detach().flatten() for p in tr.ema.ema_model.parameters()])
    assert torch.isfinite(ema_p).all()


def test_packed_weight_cache_follows_fused_updates():
    """FusedAdam / MultiLerp write parameters through raw pointers; the eval engine's packed bf16 weight
    cache is keyed on (data_ptr, _version) and must be invalidated by them.  eval forward -> fused step / EMA lerp ->
    eval forward again must equal a freshly built engine on the updated weights."""
    from ccdm_b200.optim import FusedAdam
    from ccdm_b200.ema import EMA
    spec = SPECS["tiny"]
    net, _ = make_net(spec, 3)
    net.eval()
    x, t, emb = (v.cuda() for v in unet_inputs("tiny"))
    with torch.no_grad():
        y0 = net(x, t, emb, cond_drop_prob=0.0).clone()
    # ---- fused optimizer step on synthetic gradients
    opt = FusedAdam(net.parameters(), lr=5e-2, betas=(0.9, 0.99), max_grad_norm=None)
    opt.zero_grad()
    g = torch.Generator(device="cuda").manual_seed(0)
    for p in net.parameters():
        p.grad.add_(torch.randn(p.shape, device="cuda", generator=g))
    opt.step()
    with torch.no_grad():
        y1 = net(x, t, emb, cond_drop_prob=0.0).clone()
    net2, _ = make_net(spec, 3)
    net2.load_state_dict(net.state_dict())
    net2.eval()
    with torch.no_grad():
        y1_fresh = net2(x, t, emb, cond_drop_prob=0.0)
    assert relerr(y1, y0) > 1e-2, "the optimizer step must change the output"
    assert relerr(y1, y1_fresh) < 1e-6, relerr(y1, y1_fresh)
    # ---- EMA lerp into a model whose engine already exists
    ema = EMA(net, beta=0.5, update_after_step=0, update_every=1)
    ema.ema_model.eval()
    with torch.no_grad():
        e0 = ema.ema_model(x, t, emb, cond_drop_prob=0.0).clone()
    for _ in range(3):                                   # step 0 copies, later steps lerp through ccdm_multi_lerp
        opt.zero_grad()
        for p in net.parameters():
            p.grad.add_(torch.randn(p.shape, device="cuda", generator=g))
        opt.step()
        ema.update()
    with torch.no_grad():
        e1 = ema.ema_model(x, t, emb, cond_drop_prob=0.0).clone()
    net3, _ = make_net(spec, 3)
    net3.load_state_dict(ema.ema_model.state_dict())
    net3.eval()
    with torch.no_grad():
        e1_fresh = net3(x, t, emb, cond_drop_prob=0.0)
    assert relerr(e1, e0) > 1e-3
    assert relerr(e1, e1_fresh) < 1e-6, relerr(e1, e1_fresh)


def test_trainer_graph_path_runs_without_host_syncs(tmp_path):
    """Trainer.train on the CUDA-graph path: the dataset lives on the device as uint8, a step is one graph
    replay (batch construction + 2 accumulated micro-batches + clip + Adam inside), and between the log points the loop issues
    NO host synchronisation -- asserted with torch.cuda.set_sync_debug_mode("error") around the replays."""
    import numpy as np
    import ccdm_b200
    spec = SPECS["rc_small"]
    net, _ = make_net(spec, 5, p_drop=0.1)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=1000, sampling_timesteps=5, objective="pred_x0",
                                     cond_drop_prob=0.1, vicinity_type="hv").cuda()
    rng = np.random.RandomState(0)
    images = rng.randint(0, 256, size=(96, 3, 16, 16)).astype(np.uint8)
    labels = np.round(rng.rand(96), 2).astype(np.float32)
    tr = ccdm_b200.Trainer("UTKFace", gd, train_images=images, train_labels=labels,
                           vicinal_params={"kernel_sigma": 0.05, "kappa": 0.1, "nonzero_soft_weight_threshold": 1e-3},
                           train_batch_size=8, gradient_accumulate_every=2, train_num_steps=12, save_every=100,
                           results_folder=str(tmp_path), vicinity_type="hv", kappa=0.1, sigma_delta=0.05,
                           ema_update_after_step=2, ema_update_every=2)
    fn_y2h = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=torch.device("cuda")).fn_y2h
    before = torch.cat([p.detach().flatten().clone() for p in gd.parameters()])
    gd.train()
    tr.train_num_steps = 9                                  # builds the graph (3 warm-up steps are real steps) and replays
    tr.train(fn_y2h)
    assert tr._graph_step is not None and tr.step == 9
    assert tr._images_dev.dtype == torch.uint8 and tr._images_dev.is_cuda
    tr.train_num_steps = 12
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        tr.train(fn_y2h)                                    # steps 9, 10, 11: pure replays, EMA lerps, no log point (step % 500)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert tr.step == 12 and float(tr.opt.step_count.item()) == 12.0
    after = torch.cat([p.detach().flatten() for p in gd.parameters()])
    assert torch.isfinite(after).all() and (after - before).abs().max() > 0
    assert torch.isfinite(tr._graph_step.loss).all() and float(tr._graph_step.loss) > 0
    ema_p = torch.cat([p.detach().flatten() for p in tr.ema.ema_model.parameters()])
    assert torch.isfinite(ema_p).all() and (ema_p - before).abs().max() > 0
    # the eval engine of the EMA model sees the graph-updated weights (packed-weight cache invalidation)
    y = torch.from_numpy(labels[:2]).cuda()
    tr.ema.ema_model.eval()
    img = tr.ema.ema_model.ddim_sample(labels_emb=fn_y2h(y), labels=y, shape=(2, 3, 16, 16), cond_scale=1.5)
    assert torch.isfinite(img).all()


def test_gather_augment_kernel_matches_numpy():
    """ccdm_gather_augment_u8 (trainer.py:461-482, utils.py:164-211): every augmentation code against numpy."""
    import numpy as np
    from ccdm_b200 import _lib as L
    rng = np.random.RandomState(3)
    imgs = rng.randint(0, 256, size=(7, 3, 12, 12)).astype(np.uint8)
    d_imgs = torch.from_numpy(imgs).cuda()
    idx = torch.tensor([(3 * i) % 7 for i in range(16)], dtype=torch.int64, device="cuda")
    aug = torch.arange(16, dtype=torch.uint8, device="cuda")
    out = torch.empty(16, 3, 12, 12, device="cuda")
    L.check(L.lib().ccdm_gather_augment_u8(d_imgs.data_ptr(), 7, idx.data_ptr(), aug.data_ptr(), out.data_ptr(), 16, 3, 12, 12,
                                           torch.cuda.current_stream().cuda_stream), "gather")
    out = out.cpu().numpy()
    for b in range(16):
        k, hf, vf = b & 3, (b >> 2) & 1, (b >> 3) & 1
        ref = np.rot90(imgs[(3 * b) % 7].astype(np.float32), k=k, axes=(1, 2))
        if hf:
            ref = ref[:, :, ::-1]
        if vf:
            ref = ref[:, ::-1, :]
        assert np.allclose(out[b], ref / 255.0, atol=1e-7), b
