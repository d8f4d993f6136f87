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
