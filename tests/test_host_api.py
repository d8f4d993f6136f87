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
    with pytest.raises(FileNotFoundError):                      # learned MLPs load the reference's checkpoints (tests/test_label_mlp.py)
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
                             train_batch_size=16, results_folder=str(tmp_path), vicinity_type=vicinity_type, kappa=kappa,
                             sigma_delta=0.05, **kw)


def test_trainer_batch_construction_hard_vicinity(tmp_path):
    """Trainer.sample_real_indices (vectorised restatement of trainer.py:420-459): every pick lies in the hard vicinity
    of its target label, all vicinity members are reachable, and an empty vicinity falls back to a uniformly random sample
    (trainer.py:446-450), whose vicinal weight is then 0 (trainer.py:663-675)."""
    torch.manual_seed(0)
    labels = torch.linspace(0, 1, 101)
    tr = _cpu_trainer(labels.numpy(), "hv", 0.021, tmp_path)
    targets = torch.tensor([0.5, 0.1, 0.9, 0.333]).repeat(200)
    idx = tr.sample_real_indices(targets)
    assert ((labels[idx] - targets).abs() <= 0.021 + 1e-6).all()
    picked_for_half = set(idx[0::4].tolist())
    assert picked_for_half == {48, 49, 50, 51, 52}                  # uniform over the whole vicinity of 0.5
    # empty vicinity -> uniformly random over the whole set (trainer.py:446-450); the hard vicinal weight of such a pick is 0
    sparse = _cpu_trainer([0.0, 0.4, 1.0], "hv", 0.01, tmp_path)
    tg = torch.tensor([0.3, 0.75, 0.05]).repeat(100)
    picks = sparse.sample_real_indices(tg)
    assert set(picks[0::3].tolist()) == {0, 1, 2}
    w = sparse.vicinal_weights(sparse._train_labels_dev()[picks], tg)
    assert (w == 0).all()
    # inside the vicinity the weight is 1; the soft type weighs by exp(-dist^2 / kappa^2) (trainer.py:676-690)
    assert (tr.vicinal_weights(tr._train_labels_dev()[idx], targets) == 1).all()
    soft = _cpu_trainer(labels.numpy(), "sv", 0.5, tmp_path)
    ws = soft.vicinal_weights(torch.tensor([[0.2], [0.9]]), torch.tensor([[0.3], [0.9]]))
    assert torch.allclose(ws, torch.exp(-torch.tensor([0.01, 0.0]) / 0.25), atol=1e-6)


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


def test_trainer_sliced_search_prefers_samples_matched_by_more_projections(tmp_path, monkeypatch):
    """trainer.py:380-400: among the matches, the (up to) ten samples hit by the most projections are the candidates; when
    nothing matches the nearest neighbour is used (trainer.py:401-416)."""
    import ccdm_b200.trainer as T
    torch.manual_seed(2)
    # 12 samples on the x axis near the target + one sample that ALSO matches along y: it heads the candidate list (count 2),
    # the other nine candidates are count-1 matches (lowest indices first: this build's tie-break)
    labels = torch.tensor([[0.50 + 0.001 * i, 0.9] for i in range(12)] + [[0.5, 0.5]] + [[0.05, 0.05]])
    v = torch.tensor([[1.0, 0.0], [0.0, 1.0]])
    monkeypatch.setattr(T, "generate_random_vectors", lambda kind, dim, n, device: v.to(device))
    tr = _cpu_trainer(labels.numpy(), "shv", 0.02, tmp_path, label_dim=2, num_projections=2)
    idx = tr.sample_real_indices(torch.tensor([[0.5, 0.5]]).repeat(600, 1))
    assert set(idx.tolist()) == {12} | set(range(9))
    # one projection only: ten candidates (lowest indices first), all reachable
    monkeypatch.setattr(T, "generate_random_vectors", lambda kind, dim, n, device: v[:1].to(device))
    tr1 = _cpu_trainer(labels.numpy(), "shv", 0.02, tmp_path, label_dim=2, num_projections=1)
    idx = tr1.sample_real_indices(torch.tensor([[0.5, 0.0]]).repeat(400, 1))
    assert set(idx.tolist()) == set(range(10))
    # nothing within kappa on any projection: nearest neighbour
    far = tr1.sample_real_indices(torch.tensor([[0.2, 0.2]]))
    assert far.tolist() == [13]


def test_trainer_hyperparameters_follow_the_reference_rules(tmp_path):
    """compute_hyperparameters (trainer.py:173-252): rule of thumb and percentile, hard and soft vicinities."""
    import numpy as np
    rng = np.random.RandomState(0)
    labels = np.round(rng.rand(200), 2).astype("float32")
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2))
    gd = ccdm_b200.GaussianDiffusion(net, image_size=8, timesteps=100)
    images = (rng.rand(200, 3, 8, 8) * 255).astype("float32")
    vp = dict(kernel_sigma=0.05, kappa=0.02, nonzero_soft_weight_threshold=1e-3)

    def make(**kw):
        return ccdm_b200.Trainer("RC-49", gd, images, labels, vp, train_batch_size=16, results_folder=str(tmp_path), **kw)
    tr = make(vicinity_type="hv")
    uniq = np.unique(labels)
    assert abs(tr.sigma_delta - 1.06 * labels.astype("float64").std() * 200 ** -0.2) < 1e-6
    assert abs(tr.kappa - np.diff(np.sort(uniq)).max()) < 1e-6
    soft = make(vicinity_type="sv")
    assert abs(soft.kappa - 1 / np.diff(np.sort(uniq)).max() ** 2) / soft.kappa < 1e-5
    pct = make(vicinity_type="hv", hyperparameter="percentile", percentile=5.0, distance="l1")
    d = np.abs(labels[:, None].astype("float64") - labels[None, :].astype("float64"))[np.triu_indices(200, 1)]
    assert abs(pct.kappa - np.percentile(d, 5.0)) < 1e-6 and abs(pct.sigma_delta - pct.kappa / 3) < 1e-9
    given = make(vicinity_type="hv", kappa=0.123, sigma_delta=0.04)           # both given: used as they are
    assert given.kappa == 0.123 and given.sigma_delta == 0.04


def test_trainer_process_images_matches_the_reference_augmentation(tmp_path, monkeypatch):
    """process_images (trainer.py:461-482 + utils.py:164-211): rot90 (np.rot90 semantics), then horizontal, then vertical
    flip, then /255 -- checked code by code against numpy."""
    import numpy as np
    rng = np.random.RandomState(1)
    images = rng.randint(2, 256, size=(6, 1, 8, 8)).astype("float32")
    labels = np.linspace(0, 1, 6).astype("float32")
    net = ccdm_b200.Unet(dim=32, dim_mults=(1, 2), in_channels=1)
    gd = ccdm_b200.GaussianDiffusion(net, image_size=8, timesteps=100)
    tr = ccdm_b200.Trainer("Cell200", gd, images, labels, dict(kernel_sigma=0.05, kappa=0.2, nonzero_soft_weight_threshold=1e-3),
                           train_batch_size=16, results_folder=str(tmp_path), vicinity_type="hv", kappa=0.2, sigma_delta=0.05)
    codes = torch.arange(16, dtype=torch.uint8)
    monkeypatch.setattr(tr, "augmentation_bits", lambda n: codes[:n])
    idx = torch.arange(16) % 6
    out = tr.process_images(idx).numpy()
    for b in range(16):
        k, hf, vf = b & 3, (b >> 2) & 1, (b >> 3) & 1
        ref = np.rot90(images[b % 6], k=k, axes=(1, 2))
        if hf:
            ref = ref[:, :, ::-1]
        if vf:
            ref = ref[:, ::-1, :]
        assert np.allclose(out[b], ref / 255.0), b
    # UTKFace flips horizontally only; other data sets are left alone
    tr.data_name = "RC-49"
    monkeypatch.undo()
    assert tr.augmentation_bits(4) is None


def test_trainer_device_batch_is_well_formed(tmp_path):
    """device_batch: images in [0, 1], labels of the picked samples, hard weights in {0, 1}, sliced kwargs passed on."""
    torch.manual_seed(3)
    labels = torch.linspace(0, 1, 101)
    tr = _cpu_trainer(labels.numpy(), "hv", 0.03, tmp_path)
    le = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=torch.device("cpu"))
    images, lab, emb, w, kw = tr.device_batch(le.fn_y2h)
    assert images.shape == (16, 3, 8, 8) and 0 <= images.min() and images.max() <= 1
    assert lab.shape == (16,) and emb.shape == (16, 128) and kw == {}
    assert set(w.tolist()) <= {0.0, 1.0}
    trs = _cpu_trainer(torch.rand(50, 2).numpy(), "ssv", 5.0, tmp_path, label_dim=2, num_projections=3)
    images, lab, emb2, w, kw = trs.device_batch(lambda y: torch.zeros(len(y), 128))
    assert lab.shape == (16, 2) and (w == 1).all() and kw["vicinity_type"] == "ssv" and kw["num_projections"] == 3
