"""Host logic of ccdm_b200.GaussianDiffusion.p_losses / forward on CPU: RNG order, label-drop handling, covariance (Hy)
weighting, in-batch vicinal weights for every vicinity type, the [B,1]-label broadcast quirk, loss value and the gradient
handed to the network -- with the q_sample / vicinal-loss / vicinal-weights kernels run from their own CUDA source compiled
for the host (tests/hostsim) and a small differentiable stand-in for the UNet.  The same stand-in and the same torch seed go
through oracle.p_losses (pinned to the reference's own p_losses by tests/golden/loss.pt); results must agree.

The UNet itself is covered elsewhere (tests/test_engine_emulated.py on CPU, tests/test_gpu_*.py on the device).
"""
import ctypes as C

import pytest
import torch
from torch import nn

import ccdm_b200
import oracle
from ccdm_b200 import _lib as L
from ccdm_b200.diffusion import GaussianDiffusion
from tests.hostsim.build import build


class StubDenoiser(nn.Module):
    """Differentiable stand-in with the attributes GaussianDiffusion reads and the training call signature of Unet."""

    def __init__(self, c, emb_dim):
        super().__init__()
        self.in_channels = self.out_dim = c
        self.random_or_learned_sinusoidal_cond = False
        self.mix = nn.Conv2d(c, c, 3, padding=1)
        self.emb = nn.Linear(emb_dim, c)
        self.null = nn.Parameter(torch.full((c,), -0.3))

    def forward(self, x, timesteps, labels_emb, keep_mask=None, cond_drop_prob=None):
        e = self.emb(labels_emb)
        if keep_mask is not None:
            e = torch.where(keep_mask[:, None], e, self.null[None].expand_as(e))
        return self.mix(x) * (1 + 0.1 * torch.sin(timesteps.float() / 100.0))[:, None, None, None] + e[:, :, None, None]


@pytest.fixture()
def host_sampler(monkeypatch):
    h = C.CDLL(build("sampler.cu"))

    class Lib:
        pass
    lib = Lib()
    for name in ("ccdm_q_sample", "ccdm_vicinal_loss", "ccdm_vicinal_weights"):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = L.SIGNATURES[name]
        setattr(lib, name, fn)
    h.hostsim_last_error.restype = C.c_char_p
    lib.ccdm_last_error = h.hostsim_last_error
    monkeypatch.setattr(L, "lib", lambda precision="bf16": lib)
    monkeypatch.setattr(GaussianDiffusion, "_stream", staticmethod(lambda: None))


CASES = {
    "hv_x0": dict(objective="pred_x0", vic="hv", use_Hy=False, label_dim=1, kappa=0.12),
    "hv_x0_Hy": dict(objective="pred_x0", vic="hv", use_Hy=True, label_dim=1, kappa=0.12),
    "sv_eps": dict(objective="pred_noise", vic="sv", use_Hy=False, label_dim=1, kappa=0.3),
    "shv_scalar_v": dict(objective="pred_v", vic="shv", use_Hy=False, label_dim=1, kappa=0.12),
    "shv_multi": dict(objective="pred_x0", vic="shv", use_Hy=False, label_dim=3, kappa=0.25, nproj=2),
    "ssv_multi_Hy": dict(objective="pred_noise", vic="ssv", use_Hy=True, label_dim=3, kappa=0.4, nproj=3),
    "hv_multi": dict(objective="pred_x0", vic="hv", use_Hy=False, label_dim=3, kappa=0.35),
    "hv_col_labels": dict(objective="pred_x0", vic="hv", use_Hy=False, label_dim=-1, kappa=0.12),      # labels shaped [B,1]
    "novic_eps": dict(objective="pred_noise", vic=None, use_Hy=False, label_dim=1, kappa=0.1),
    # distance l1 with [B,1] labels: abs(diff).sum(dim=2) gives ordinary [B] weights -- no broadcast quirk (diffusion.py:684-692)
    "hv_col_labels_l1": dict(objective="pred_x0", vic="hv", use_Hy=False, label_dim=-1, kappa=0.12, distance="l1"),
    "sv_multi_l1": dict(objective="pred_x0", vic="sv", use_Hy=False, label_dim=3, kappa=0.5, distance="l1"),
}


@pytest.mark.parametrize("name", list(CASES))
def test_p_losses_host_logic_matches_oracle(host_sampler, name):
    c = CASES[name]
    B, ch, size, emb_dim, p_drop = 8, 3, 8, 16, 0.3
    g = torch.Generator().manual_seed(77)
    img = torch.rand(B, ch, size, size, generator=g)
    d = c["label_dim"]
    labels = torch.rand(B, generator=g) if d == 1 else torch.rand(B, 1 if d == -1 else d, generator=g)
    le = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", y2cov_type="sinusoidal", h_dim=emb_dim, cov_dim=ch * size * size,
                              device=torch.device("cpu"), label_dim=max(d, 1), dim_combination="mean")
    emb = le.fn_y2h(labels)
    torch.manual_seed(5)
    net = StubDenoiser(ch, emb_dim)
    gd = GaussianDiffusion(net, image_size=size, use_Hy=c["use_Hy"], fn_y2cov=le.fn_y2cov if c["use_Hy"] else None,
                           cond_drop_prob=p_drop, timesteps=1000, objective=c["objective"], vicinity_type=c["vic"]).train()
    kw = dict(labels_emb=emb, labels=labels, vicinal_weights=None if c["vic"] is None else torch.ones(B))
    okw = {}
    if c["vic"] is not None:
        vec = torch.randn(c.get("nproj", 1), max(d, 1), generator=g)
        extra = dict(vicinity_type=c["vic"], kappa=c["kappa"], vector_type="gaussian", num_projections=c.get("nproj", 1),
                     cached_vectors=vec)
        if "distance" in c:
            extra["distance"] = c["distance"]
        kw.update(extra)
        okw.update(extra)
    # ---- product code (forward draws t, then p_losses draws the mask and the noise)
    torch.manual_seed(123)
    val = gd(img, **kw)
    val.backward()
    got = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad()
    # ---- oracle with the same stand-in network and the same RNG stream
    sch = oracle.make_schedule(1000, "cosine", c["objective"])
    torch.manual_seed(123)
    t = torch.randint(0, 1000, (B,)).long()
    keep_box = {}

    def net_train(x, tt, e):
        return net(x, tt, e, keep_mask=keep_box["keep"])

    # the oracle draws the keep mask itself; reproduce it for the stand-in by peeking at the same uniform draw
    state = torch.get_rng_state()
    keep_box["keep"] = torch.zeros(B).float().uniform_(0, 1) < (1 - p_drop)
    torch.set_rng_state(state)
    ref = oracle.p_losses(sch, net_train, img * 2 - 1, t, labels=labels, labels_emb=emb, cond_drop_prob=p_drop,
                          use_Hy=c["use_Hy"], fn_y2cov=le.fn_y2cov if c["use_Hy"] else None,
                          vicinal_weights=kw["vicinal_weights"], **okw)
    ref.backward()
    assert abs(val.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), (val.item(), ref.item())
    for k, p in net.named_parameters():
        assert ((got[k] - p.grad).norm() / p.grad.norm().clamp_min(1e-12)).item() < 1e-4, k


# ------------------------------------------------------------------------------------------------- sampling loops

class _StubProgram:
    """The surface _loop / _SamplerState use of an engine program: x_in / t_in / emb_in / keep / out / run(stream)."""

    def __init__(self, net, B, x_batch, H, W):
        c = net.in_channels
        self.net, self.B, self.x_batch = net, B, x_batch
        self.x_in = torch.zeros(x_batch, c, H, W)
        self.t_in = torch.zeros(B, dtype=torch.int64)
        self.emb_in = torch.zeros(B, net.emb.in_features)
        self.keep = torch.zeros(B, dtype=torch.uint8)
        self.out = torch.zeros(B, c, H, W)

    def run(self, stream):
        x = self.x_in.repeat(self.B // self.x_batch, 1, 1, 1)
        with torch.no_grad():
            self.out.copy_(self.net(x, self.t_in, self.emb_in, keep_mask=self.keep.bool()))


class _StubEngine:
    def __init__(self, net):
        self.net, self.progs = net, {}
        self.weights = type("W", (), {"refresh": staticmethod(lambda stream: None)})()

    def program(self, B, x_batch, H, W, training):
        return self.progs.setdefault((B, x_batch, H, W), _StubProgram(self.net, B, x_batch, H, W))


@pytest.fixture()
def host_loop(monkeypatch, host_sampler):
    import ccdm_b200.diffusion as DM
    lib = L.lib()
    h = C.CDLL(build("sampler.cu"))
    for name in ("ccdm_sampler_step", "ccdm_broadcast_step_i64"):
        fn = getattr(h, name)
        fn.restype, fn.argtypes = L.SIGNATURES[name]
        setattr(lib, name, fn)
    # no CUDA graph on the CPU: every step launches the same three calls eagerly
    monkeypatch.setattr(DM._SamplerState, "step", lambda self: self._launch(None))


@pytest.mark.parametrize("kind,objective,eta,use_Hy,scale", [("ddim", "pred_x0", 0.0, False, 1.5), ("ddim", "pred_noise", 0.5, False, 1.5),
                                                             ("ddim", "pred_x0", 0.0, True, 1.5), ("ddim", "pred_v", 0.0, False, 2.0),
                                                             ("ddpm", "pred_noise", 0.0, False, 2.0), ("ddpm", "pred_x0", 0.0, True, 1.5),
                                                             ("ddim", "pred_x0", 0.3, False, 1.0)])
def test_sampling_loop_host_logic_matches_oracle(host_loop, kind, objective, eta, use_Hy, scale):
    """_loop + _SamplerState (coefficient tables, device step counter, noise draws, Hy initial scaling, guided 2B batch) with
    ccdm_broadcast_step_i64 / ccdm_sampler_step run from source, against oracle.ddim_sample / ddpm_sample with the same
    stand-in network and torch seed."""
    B, ch, size, emb_dim, S = 3, 3, 8, 16, 5
    le = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", y2cov_type="sinusoidal", h_dim=emb_dim, cov_dim=ch * size * size,
                              device=torch.device("cpu"))
    labels = torch.linspace(0.1, 0.9, B)
    emb = le.fn_y2h(labels)
    torch.manual_seed(6)
    net = StubDenoiser(ch, emb_dim).eval()
    net.engine = lambda _e=_StubEngine(net): _e
    gd = GaussianDiffusion(net, image_size=size, use_Hy=use_Hy, fn_y2cov=le.fn_y2cov if use_Hy else None, timesteps=1000,
                           sampling_timesteps=S, objective=objective, ddim_sampling_eta=eta).eval()
    shape = (B, ch, size, size)
    torch.manual_seed(321)
    if kind == "ddim":
        img = gd.ddim_sample(labels_emb=emb, labels=labels, shape=shape, cond_scale=scale)
    else:
        img = gd.sample(labels_emb=emb, labels=labels, cond_scale=scale)
    sch = oracle.make_schedule(1000, "cosine", objective)

    def onet(x, t, e, p):
        keep = torch.ones(x.shape[0], dtype=torch.bool) if p == 0.0 else torch.zeros(x.shape[0], dtype=torch.bool)
        with torch.no_grad():
            return net(x, t, e, keep_mask=keep)

    cov = torch.exp(-le.fn_y2cov(labels).view(shape)) if use_Hy else None
    torch.manual_seed(321)
    if kind == "ddim":
        ref = oracle.ddim_sample(sch, onet, emb, shape, sampling_timesteps=S, cond_scale=scale, eta=eta, init_cov=cov)
    else:
        ref = oracle.ddpm_sample(sch, onet, emb, shape, sampling_timesteps=S, cond_scale=scale, init_cov=cov)
    assert (img - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("kind,objective,eta,phi", [("ddim", "pred_x0", 0.0, 0.2), ("ddim", "pred_noise", 1.0, 0.7), ("ddpm", "pred_x0", 1.0, 0.4)])
def test_vanilla_sampling_wrapper_matches_oracle(host_loop, kind, objective, eta, phi):
    """VanillaGaussianDiffusion's call-site details on CPU: plain CFG (cfg_remove_parallel = False), ddim_sample ignoring its
    rescaled_phi (always 0.7), DDPM over preset_sampling_timesteps -- against oracle.vanilla_diffusion_ref."""
    from oracle.vanilla_diffusion_ref import v_ddim_sample, v_ddpm_sample
    B, ch, size, emb_dim, S, scale = 3, 3, 8, 16, 4, 1.5
    g = torch.Generator().manual_seed(8)
    classes = torch.rand(B, emb_dim, generator=g)
    torch.manual_seed(6)
    net = StubDenoiser(ch, emb_dim).eval()
    net.cfg_remove_parallel = False
    net.engine = lambda _e=_StubEngine(net): _e
    gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=size, timesteps=1000, sampling_timesteps=S, objective=objective,
                                            ddim_sampling_eta=eta).eval()
    shape = (B, ch, size, size)
    torch.manual_seed(11)
    if kind == "ddim":
        img = gd.ddim_sample(classes, shape, cond_scale=scale, rescaled_phi=phi)
    else:
        img = gd.sample(classes, cond_scale=scale, rescaled_phi=phi, preset_sampling_timesteps=S)

    def guided(x, t, cl, cs, ph):                          # V/diffusion.py:34-56 over the stand-in network
        with torch.no_grad():
            cond = net(x, t, cl, keep_mask=torch.ones(B, dtype=torch.bool))
            null = net(x, t, cl, keep_mask=torch.zeros(B, dtype=torch.bool))
        s = null + (cond - null) * cs
        if ph == 0:
            return s
        dims = (1, 2, 3)
        return s * (cond.std(dim=dims, keepdim=True) / s.std(dim=dims, keepdim=True)) * ph + s * (1 - ph)

    sch = oracle.make_schedule(1000, "cosine", objective)
    torch.manual_seed(11)
    if kind == "ddim":
        ref = v_ddim_sample(sch, guided, classes, shape, sampling_timesteps=S, cond_scale=scale, eta=eta)
    else:
        ref = v_ddpm_sample(sch, guided, classes, shape, steps=S, cond_scale=scale, rescaled_phi=phi)
    assert (img - ref).abs().max().item() < 2e-4
    assert gd.sampling_timesteps == S and gd.ddim_sampling_eta == eta        # the preset overrides were restored
