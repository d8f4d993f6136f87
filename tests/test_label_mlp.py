"""Learned label-embedding hooks: ``LabelEmbed(y2h_type="resnet", y2cov_type="resnet")``.

Reference: label_embedding.py:861-1178 (fn_y2h / fn_y2cov "resnet" branches), models/resnet_y2h.py:143-173,
models/resnet_y2cov.py:149-179.  Golden vectors: tests/golden/label_mlp.pt, produced by the reference's own modules
(tests/golden/make_golden_label_mlp.py).  CPU tier: the oracle restatement and the PRODUCT forward run over host builds of the
library's kernels (tests/hostpath.py); GPU tier: the same on the device, plus a larger cov_dim against the oracle."""
import os

import pytest
import torch
from torch import nn

import ccdm_b200
from ccdm_b200.label_embedding import model_y2h, model_y2cov
from oracle.label_mlp_ref import label_mlp_forward
from tests.golden.make_golden_label_mlp import CASES, randomize_affine, case_labels

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "label_mlp.pt"), weights_only=True)


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm().clamp_min(1e-12)).item()


def build_case(name, device):
    """Same RNG consumption as the generating script: seed -> LabelEmbed (combiner nets) -> model_y2h -> affine -> model_y2cov."""
    seed, h_dim, cov_dim, label_dim, comb, batch = CASES[name]
    torch.manual_seed(seed)
    le = ccdm_b200.LabelEmbed(dataset=None, path_y2h=None, path_y2cov=None, y2h_type="sinusoidal", y2cov_type=None,
                              h_dim=h_dim, cov_dim=cov_dim, nc=3, device=torch.device("cpu"), label_dim=label_dim,
                              dim_combination=comb)
    mh = model_y2h(dim_embed=h_dim)
    gen = torch.Generator().manual_seed(seed)
    randomize_affine(mh, gen)
    le.y2h_type, le.model_mlp_y2h = "resnet", mh
    if cov_dim is not None:
        mc = model_y2cov(dim_embed=cov_dim)
        randomize_affine(mc, gen)
        le.y2cov_type, le.model_mlp_y2cov = "resnet", mc
    if device.type == "cuda":
        le.device = device
        for attr in ("model_mlp_y2h", "model_mlp_y2cov", "h_attention_net", "h_cross_net"):
            if getattr(le, attr, None) is not None:
                setattr(le, attr, getattr(le, attr).to(device))
    return le, case_labels(seed, label_dim, batch).to(device)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_vs_reference_golden(name):
    le, labels = build_case(name, torch.device("cpu"))
    if CASES[name][3] == 1:                                   # scalar labels: the MLP itself
        assert rel(label_mlp_forward(le.model_mlp_y2h.state_dict(), labels), GOLD[name]["h"]) < 1e-5
        assert rel(label_mlp_forward(le.model_mlp_y2cov.state_dict(), labels), GOLD[name]["cov"]) < 1e-5
    elif CASES[name][4] == "mean":
        h = torch.stack([label_mlp_forward(le.model_mlp_y2h.state_dict(), labels[:, d]) for d in range(labels.shape[1])]).mean(0)
        assert rel(h, GOLD[name]["h"]) < 1e-5


@pytest.mark.parametrize("name", list(CASES))
def test_product_forward_on_host_kernels_vs_reference_golden(name, monkeypatch):
    """LabelEmbed.fn_y2h / fn_y2cov UNCHANGED, with ccdm_linear_small / ccdm_groupnorm_rows built for the host from source."""
    from tests import hostpath
    hostpath.install(monkeypatch)
    le, labels = build_case(name, torch.device("cpu"))
    assert rel(le.fn_y2h(labels), GOLD[name]["h"]) < 2e-5
    if "cov" in GOLD[name]:
        assert rel(le.fn_y2cov(labels), GOLD[name]["cov"]) < 2e-5


def test_checkpoint_discovery_and_loading(tmp_path):
    """The reference's two checkpoint locations and its nn.DataParallel key prefix (label_embedding.py:406-424,533-536)."""
    torch.manual_seed(5)
    src = model_y2h(dim_embed=64)
    d = tmp_path / "y2h"
    (d / "y2h_ckpt_in_train").mkdir(parents=True)
    torch.save({"net_state_dict": {"module." + k: v for k, v in src.state_dict().items()}},
               str(d / "y2h_ckpt_in_train" / "mlp_y2h_checkpoint_epoch_500.pth"))
    d2 = tmp_path / "y2cov"
    d2.mkdir()
    src2 = model_y2cov(dim_embed=48)
    torch.save({"net_state_dict": src2.state_dict()}, str(d2 / "ckpt_mlp_y2cov_epoch_500.pth"))
    le = ccdm_b200.LabelEmbed(dataset=None, path_y2h=str(d), path_y2cov=str(d2), y2h_type="resnet", y2cov_type="resnet",
                              h_dim=64, cov_dim=48, device=torch.device("cpu"))
    for a, b in ((le.model_mlp_y2h, src), (le.model_mlp_y2cov, src2)):
        assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    with pytest.raises(FileNotFoundError):
        ccdm_b200.LabelEmbed(dataset=None, path_y2h=str(tmp_path / "nowhere"), y2h_type="resnet", h_dim=64,
                             device=torch.device("cpu"))
    with pytest.raises(RuntimeError):                         # CPU tensors never reach a fallback
        le.fn_y2h(torch.rand(4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_vs_reference_golden(name):
    le, labels = build_case(name, torch.device("cuda"))
    assert rel(le.fn_y2h(labels), GOLD[name]["h"]) < 2e-5
    if "cov" in GOLD[name]:
        assert rel(le.fn_y2cov(labels), GOLD[name]["cov"]) < 2e-5


@pytest.mark.gpu
def test_gpu_rc64_sizes_vs_oracle():
    """The shipped RC-49 64x64 configuration: h_dim 128, cov_dim 3*64*64 = 12288 (a [200,4096] x [4096,12288] product)."""
    dev = torch.device("cuda")
    torch.manual_seed(21)
    mh, mc = model_y2h(128).to(dev), model_y2cov(3 * 64 * 64).to(dev)
    gen = torch.Generator().manual_seed(21)
    randomize_affine(mh, gen)
    randomize_affine(mc, gen)
    y = torch.rand(200, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    assert rel(mh(y), label_mlp_forward(mh.state_dict(), y)) < 1e-5
    out = mc(y)
    assert out.shape == (200, 12288) and rel(out, label_mlp_forward(mc.state_dict(), y)) < 1e-5
