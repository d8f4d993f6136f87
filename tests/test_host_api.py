"""Host-side API surface that needs no GPU: EMA arithmetic, LabelEmbed hooks vs the oracle, schedule buffers."""
import copy

import pytest
import torch

import ccdm_b200
import oracle


def test_schedule_buffers_bit_exact_vs_reference_golden():
    import os
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "schedules.pt"), weights_only=True)
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2))
    for key, tabs in gold.items():
        T, kind, obj = key.split("_", 2)
        gd = ccdm_b200.GaussianDiffusion(net, image_size=16, timesteps=int(T), beta_schedule=kind, objective=obj)
        for name, ref in tabs.items():
            assert torch.equal(getattr(gd, name), ref), (key, name)
        assert list(dict(gd.named_buffers()).keys())[:13] == list(oracle.Schedule.NAMES)


def test_label_hooks_match_oracle():
    le = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", y2cov_type="sinusoidal", h_dim=128, cov_dim=3 * 8 * 8,
                              device=torch.device("cpu"), label_dim=3, dim_combination="mean")
    y1 = torch.rand(5)
    y3 = torch.rand(5, 3)
    assert torch.allclose(le.fn_y2h(y1), oracle.y2h_sinusoidal(y1, 128))
    assert torch.allclose(le.fn_y2h(y1[:, None]), oracle.y2h_sinusoidal(y1[:, None], 128))
    assert torch.allclose(le.fn_y2h(y3), oracle.y2h_sinusoidal(y3, 128))
    assert torch.allclose(le.fn_y2cov(y3), oracle.y2cov_sinusoidal(y3, 192))
    assert le.fn_y2cov(y1).min() >= 0 and le.fn_y2h(y1).max() <= 1
    with pytest.raises(NotImplementedError):
        ccdm_b200.LabelEmbed(y2h_type="resnet")


def test_ema_warmup_and_lerp():
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 4)
    ema = ccdm_b200.EMA(model, beta=0.9, update_after_step=2, update_every=1)
    w0 = model.weight.detach().clone()
    for step in range(6):
        with torch.no_grad():
            model.weight.add_(1.0)
        ema.update()
        if step <= 2:                                   # copies while warming up (ema_pytorch.py:155-158)
            assert torch.equal(ema.ema_model.weight, model.weight)
    assert not torch.equal(ema.ema_model.weight, model.weight)
    assert (ema.ema_model.weight < model.weight).all() and (ema.ema_model.weight > w0).all()
    assert 0.0 < ema.get_current_decay() <= 0.9


def test_unet_deepcopy_and_dataparallel_wrapper():
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2))
    gd = ccdm_b200.GaussianDiffusion(torch.nn.DataParallel(net), image_size=8, timesteps=100)
    twin = copy.deepcopy(gd)                             # what EMA does (ema_pytorch.py:69-74)
    assert twin.unet is not gd.unet and twin.unet._engine is None
    assert set(twin.state_dict()) == set(gd.state_dict())
    assert any(k.startswith("model.module.") for k in gd.state_dict())      # reference checkpoint key layout


def _cpu_trainer(labels, vicinity_type, kappa, tmp_path, **kw):
    import numpy as np
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2))
    gd = ccdm_b200.GaussianDiffusion(net, image_size=8, timesteps=100)
    n = len(labels)
    images = (np.random.RandomState(0).rand(n, 3, 8, 8) * 255).astype("float32")
    return ccdm_b200.Trainer("RC-49", gd, train_images=images, train_labels=np.asarray(labels, dtype="float32"),
                             vicinal_params=dict(kernel_sigma=0.05, kappa=kappa, nonzero_soft_weight_threshold=1e-3),
                             train_batch_size=16, results_folder=str(tmp_path), vicinity_type=vicinity_type, **kw)


def test_trainer_batch_construction_hard_vicinity(tmp_path):
    """Trainer.sample_real_indices (vectorised restatement of trainer.py:317-459): every pick lies in the hard vicinity
    of its target label, all vicinity members are reachable, and an empty vicinity falls back to the nearest label."""
    torch.manual_seed(0)
    labels = torch.linspace(0, 1, 101)
    tr = _cpu_trainer(labels.numpy(), "hv", 0.021, tmp_path)
    targets = torch.tensor([0.5, 0.1, 0.9, 0.333]).repeat(200)
    idx = tr.sample_real_indices(targets)
    assert ((labels[idx] - targets).abs() <= 0.021 + 1e-6).all()
    picked_for_half = set(idx[0::4].tolist())
    assert picked_for_half == {48, 49, 50, 51, 52}                  # uniform over the whole vicinity of 0.5
    # empty vicinity -> nearest neighbour (trainer.py:405-417)
    sparse = _cpu_trainer([0.0, 0.4, 1.0], "hv", 0.01, tmp_path)
    assert sparse.sample_real_indices(torch.tensor([0.3, 0.75, 0.05])).tolist() == [1, 2, 0]


def test_trainer_batch_construction_sliced_vicinity(tmp_path, monkeypatch):
    """Sliced hard vicinity with vector labels: the pick's projection on (one of) the random directions is within
    kappa * |v| of the target's (trainer.py:344-400)."""
    import ccdm_b200.trainer as T
    torch.manual_seed(1)
    labels = torch.rand(300, 3)
    v = torch.tensor([[0.6, -0.2, 1.1], [0.1, 0.9, -0.4]])
    monkeypatch.setattr(T, "generate_random_vectors", lambda kind, dim, n, device: v.to(device))
    tr = _cpu_trainer(labels.numpy(), "shv", 0.05, tmp_path, label_dim=3, num_projections=2)
    targets = torch.rand(64, 3)
    idx = tr.sample_real_indices(targets)
    vn = torch.nn.functional.normalize(v, dim=1)
    d = ((targets @ vn.t()) - (labels[idx] @ vn.t())).abs()         # [B, P]
    assert (d <= 0.05 * v.norm(dim=1)[None] + 1e-6).any(1).all()
