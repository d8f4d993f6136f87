"""Node wiring of the vanilla (GroupNorm) UNet's training step on CPU: ccdm_b200.vanilla_train runs unchanged, with
  * every CUDA-core kernel executed from its own CUDA source compiled for the host (tests/hostsim),
  * the tap-GEMM based helpers (conv forward / data gradient / weight gradient / bias sums) replaced by torch restatements
    with bf16 storage (tests/hostpath.py),
and the parameter gradients compared with autograd of the oracle.  Checks saved tensors, coefficient layouts, the
concatenated-source handling, the scale/shift gradient path into the tc_mlp Linears and the padded stem / head weights."""
import math
import os

import pytest
import torch

from ccdm_b200.vanilla_train import vanilla_train_forward
from ccdm_b200.vanilla_unet import VanillaUnet
from oracle.vanilla_unet_ref import make_state_dict, vanilla_unet_forward
from tests import hostpath
from tests.golden.vanilla_cases import V_BATCH, V_SPECS, keep_mask, vanilla_inputs


@pytest.fixture()
def host_path(monkeypatch):
    hostpath.install(monkeypatch)


@pytest.mark.parametrize("sname,kind", [("v_tiny", "mixed"), ("v_attn", "cond")])
def test_training_step_gradients_match_oracle_autograd(host_path, sname, kind):
    spec = V_SPECS[sname]
    net = VanillaUnet(embed_input_dim=spec.embed_input_dim, cond_drop_prob=0.5, in_channels=spec.in_channels,
                      model_channels=spec.model_channels, num_res_blocks=spec.num_res_blocks,
                      attention_resolutions=spec.attention_resolutions, channel_mult=spec.channel_mult,
                      num_heads=spec.num_heads, num_groups=spec.num_groups)
    sd = make_state_dict(spec, 7)
    net.load_state_dict(sd, strict=True)
    net.train()
    x, t, classes = vanilla_inputs(sname)
    if sname == "v_attn":               # attention at every level: 8x8 input (64 / 16 / 4 tokens) keeps the fiber
        x = x[..., :8, :8].contiguous()  # simulation of the attention kernels to seconds
    keep = keep_mask(kind, V_BATCH[sname])
    dout = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    out = vanilla_train_forward(net, x, t, classes, keep)
    out.backward(dout)
    # oracle: fp32 autograd over the functional restatement
    sd_g = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k and k != "null_classes_emb"
                else v.clone()) for k, v in sd.items()}
    ref = vanilla_unet_forward(sd_g, spec, x, t, classes, keep, training=True)
    ref.backward(dout)
    assert ((out - ref).norm() / ref.norm()).item() < 2e-2
    names = [n for n, p in net.named_parameters() if p.requires_grad]
    assert all(net.get_parameter(n).grad is not None for n in names)
    got = torch.cat([net.get_parameter(n).grad.flatten() for n in names]).double()
    want = torch.cat([sd_g[n].grad.flatten() for n in names]).double()
    cos = (got @ want / (got.norm() * want.norm())).item()
    err = ((got - want).norm() / want.norm()).item()
    print(f"{sname}: output rel err {((out - ref).norm() / ref.norm()).item():.3e}, gradient cosine {cos:.5f}, rel L2 {err:.3e}")
    assert cos > 0.999 and err < 5e-2
    # the first / last convolutions (padded weights) and one GroupNorm gain individually
    for n in ("down_blocks.0.0.weight", "out.2.weight", "out.2.bias", "out.0.weight", "middle_block.0.tc_mlp.1.weight",
              "classes_emb.0.weight", "time_mlp.0.weight"):
        g, w = net.get_parameter(n).grad, sd_g[n].grad
        assert ((g - w).norm() / w.norm().clamp_min(1e-12)).item() < 6e-2, n


@pytest.mark.parametrize("name", ["v_loss_x0_vic", "v_loss_eps_plain", "v_loss_v_vic_attn"])
def test_p_losses_and_backward_match_the_reference(host_path, monkeypatch, name):
    """Product code end to end -- VanillaGaussianDiffusion.p_losses (ccdm_q_sample, VanillaUnet autograd nodes,
    ccdm_vicinal_loss from their own source) + loss.backward() -- against the loss and gradients the reference's own
    GaussianDiffusion.p_losses produced (tests/golden/vanilla_loss.pt), label-drop mask injected as in the generator."""
    import os
    import ccdm_b200
    import ccdm_b200.vanilla_unet as VU
    from ccdm_b200.diffusion import GaussianDiffusion
    from tests.golden.vanilla_cases import V_LOSS_CASES, V_LOSS_GRAD_KEYS, V_SIZES, loss_inputs
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vanilla_loss.pt"))[name]
    c = V_LOSS_CASES[name]
    spec = V_SPECS[c["spec"]]
    net = VanillaUnet(embed_input_dim=spec.embed_input_dim, cond_drop_prob=0.5, in_channels=spec.in_channels,
                      model_channels=spec.model_channels, num_res_blocks=spec.num_res_blocks,
                      attention_resolutions=spec.attention_resolutions, channel_mult=spec.channel_mult,
                      num_heads=spec.num_heads, num_groups=spec.num_groups)
    net.load_state_dict(make_state_dict(spec, c["seed"]), strict=True)
    net.train()
    mask = keep_mask(c["kind"], V_BATCH[c["spec"]])
    monkeypatch.setattr(VU, "prob_mask_like", lambda shape, prob, device: mask.clone())
    monkeypatch.setattr(VU, "_require_cuda", lambda x: None)
    monkeypatch.setattr(GaussianDiffusion, "_stream", staticmethod(lambda: None))
    gd = ccdm_b200.VanillaGaussianDiffusion(torch.nn.DataParallel(net), image_size=V_SIZES[c["spec"]], timesteps=1000,
                                            objective=c["objective"]).train()
    x0, t, classes, noise, weights = loss_inputs(c)
    w_in = None if weights is None else weights.clone()
    val = gd.p_losses(x0, t, classes=classes, noise=noise, vicinal_weights=w_in)
    val.backward()
    assert abs(val.item() - gold["loss"].item()) < 2e-2 * abs(gold["loss"].item()), (val.item(), gold["loss"].item())
    if w_in is not None:                                   # rows whose label was dropped were set to 1 in place (V:399)
        assert torch.equal(w_in[~mask], torch.ones_like(w_in[~mask])) and torch.equal(w_in[mask], weights[mask])
    sq = sum(float((p.grad.double() ** 2).sum()) for p in net.parameters() if p.grad is not None)
    print(f"{name}: loss {val.item():.5f} (reference {gold['loss'].item():.5f}), |grad|^2 {sq:.4f} (reference {gold['grad_sqnorm']:.4f})")
    assert abs(sq - gold["grad_sqnorm"]) < 8e-2 * gold["grad_sqnorm"]
    for k in V_LOSS_GRAD_KEYS:
        g, w = net.get_parameter(k).grad, gold["grad_" + k]
        cos = (g.flatten().double() @ w.flatten().double() / (g.norm().double() * w.norm().double())).item()
        assert cos > 0.995, (k, cos)


# ------------------------------------------------------------------------------------------------- VanillaTrainer

def _trainer(tmp_path, labels, threshold_type, kappa, sigma=0.02, batch=16, steps=2):
    import numpy as np
    import ccdm_b200
    spec = V_SPECS["v_tiny"]
    net = VanillaUnet(embed_input_dim=spec.embed_input_dim, cond_drop_prob=0.1, in_channels=3,
                      model_channels=spec.model_channels, num_res_blocks=spec.num_res_blocks,
                      attention_resolutions=spec.attention_resolutions, channel_mult=spec.channel_mult,
                      num_heads=spec.num_heads, num_groups=spec.num_groups)
    gd = ccdm_b200.VanillaGaussianDiffusion(net, image_size=8, timesteps=100, objective="pred_x0")
    labels = np.asarray(labels, dtype="float32")
    images = (np.random.RandomState(0).rand(len(labels), 3, 8, 8) * 255).astype("float32")
    return ccdm_b200.VanillaTrainer(gd, images, labels, dict(kernel_sigma=sigma, kappa=kappa, threshold_type=threshold_type,
                                                             nonzero_soft_weight_threshold=1e-3),
                                    train_batch_size=batch, train_num_steps=steps, save_every=steps,
                                    results_folder=str(tmp_path)), net


def test_vanilla_trainer_batch_construction(tmp_path):
    """V/trainer.py:221-289 vectorised: picks lie in the hard / soft vicinity of their (noisy) targets, soft weights follow
    exp(-kappa d^2), empty vicinities are re-drawn, and kappa == 0 switches the vicinity off."""
    torch.manual_seed(0)
    labels = torch.linspace(0, 1, 51)
    tr, _ = _trainer(tmp_path, labels.numpy(), "hard", 0.03, batch=64)
    idx, targets, w = tr.draw_batch()
    assert ((labels[idx] - targets).abs() <= 0.03 + 1e-6).all() and torch.equal(w, torch.ones(64))
    soft, _ = _trainer(tmp_path, labels.numpy(), "soft", 2000.0, batch=64)
    idx, targets, w = soft.draw_batch()
    d2 = (labels[idx] - targets) ** 2
    assert (d2 <= -math.log(1e-3) / 2000.0 + 1e-9).all()
    assert torch.allclose(w, torch.exp(-2000.0 * d2)) and (w >= 1e-3 - 1e-6).all()
    # sparse labels + wide noise: many first draws land in empty vicinities and must be re-drawn (never an assertion)
    sparse, _ = _trainer(tmp_path, [0.0, 0.5, 1.0], "hard", 0.01, sigma=0.05, batch=64)
    idx, targets, _ = sparse.draw_batch()
    assert ((torch.tensor([0.0, 0.5, 1.0])[idx] - targets).abs() <= 0.01 + 1e-6).all()
    plain, _ = _trainer(tmp_path, labels.numpy(), "hard", 0.0, batch=32)
    idx, targets, w = plain.draw_batch()
    assert targets is None and w is None and idx.shape == (32,)


def test_vanilla_trainer_step_end_to_end(host_path, monkeypatch, tmp_path):
    """VanillaTrainer.train for one optimizer step on CPU (kernels from source, convs restated): the loss is finite, the
    parameters move, the EMA and the checkpoint round trip work."""
    import ccdm_b200.vanilla_unet as VU
    from ccdm_b200.diffusion import GaussianDiffusion
    monkeypatch.setattr(VU, "_require_cuda", lambda x: None)
    monkeypatch.setattr(GaussianDiffusion, "_stream", staticmethod(lambda: None))
    torch.manual_seed(3)
    tr, net = _trainer(tmp_path, torch.linspace(0, 1, 40).numpy(), "soft", 500.0, batch=16, steps=1)
    w0 = net.out[2].weight.detach().clone()
    fn_y2h = lambda y: torch.stack([torch.sin((k + 1) * y) for k in range(16)], 1)        # noqa: E731  (any callable works)
    tr.train(fn_y2h)
    assert tr.step == 1 and not torch.equal(net.out[2].weight, w0)
    assert all(torch.isfinite(p).all() for p in net.parameters())
    assert (tmp_path / "model-1.pt").exists()
    log = (tmp_path / "log_loss_niters1.txt").read_text()
    assert "Step: 0, Loss:" in log and "nan" not in log.lower()
    tr2, net2 = _trainer(tmp_path, torch.linspace(0, 1, 40).numpy(), "soft", 500.0, batch=16, steps=1)
    tr2.load(1)
    assert tr2.step == 1 and torch.equal(net2.out[2].weight, net.out[2].weight)
