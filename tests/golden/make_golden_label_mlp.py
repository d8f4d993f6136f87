"""Golden vectors for the learned label-embedding hooks: the reference's OWN ``model_y2h`` / ``model_y2cov`` modules
(models/resnet_y2h.py:143-173, models/resnet_y2cov.py:149-179) evaluated through the reference's OWN ``LabelEmbed.fn_y2h`` /
``fn_y2cov`` "resnet" branches (label_embedding.py:1028-1031, :1173-1176, multi-dimensional combiners :960-1005).

    python tests/golden/make_golden_label_mlp.py      # build container only (needs /root/reference); writes label_mlp.pt

Weights are not stored: ``torch.manual_seed(seed)`` followed by the module constructors reproduces them (the repo's holders
build the same layers in the same order), with the GroupNorm affine parameters re-drawn so that they are not the identity.
"""
import contextlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference/CCDM_unified"

import torch  # noqa: E402
from torch import nn  # noqa: E402

CASES = {
    # name: (seed, h_dim, cov_dim, label_dim, dim_combination, batch)
    "scalar": (11, 128, 3 * 8 * 8, 1, "mean", 6),
    "scalar_h64": (12, 64, 3 * 4 * 4, 1, "mean", 33),
    "multi_mean": (13, 128, 3 * 8 * 8, 3, "mean", 5),
    "multi_attention": (14, 128, None, 3, "attention", 5),
    "multi_cross": (15, 128, None, 2, "cross", 7),
}


def randomize_affine(model, gen):
    """GroupNorm weight / bias away from (1, 0) -- a trained network's are: the parity test must see them."""
    for m in model.modules():
        if isinstance(m, nn.GroupNorm):
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))


def case_labels(seed, label_dim, batch):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(batch, generator=g) if label_dim == 1 else torch.rand(batch, label_dim, generator=g)


def main():
    sys.path.insert(0, REF)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    from label_embedding import LabelEmbed          # type: ignore
    from models.resnet_y2h import model_y2h         # type: ignore
    from models.resnet_y2cov import model_y2cov     # type: ignore
    out = {}
    for name, (seed, h_dim, cov_dim, label_dim, comb, batch) in CASES.items():
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            le = LabelEmbed(dataset=None, path_y2h="/tmp/ccdm_golden_y2h", path_y2cov="/tmp/ccdm_golden_y2cov",
                            y2h_type="sinusoidal", y2cov_type=None, h_dim=h_dim, cov_dim=cov_dim, nc=3,
                            device=torch.device("cpu"), label_dim=label_dim, dim_combination=comb)
        mh = model_y2h(dim_embed=h_dim)
        gen = torch.Generator().manual_seed(seed)
        randomize_affine(mh, gen)
        le.y2h_type, le.model_mlp_y2h = "resnet", nn.DataParallel(mh)
        labels = case_labels(seed, label_dim, batch)
        rec = {}
        with torch.no_grad():
            rec["h"] = le.fn_y2h(labels).clone()
            if cov_dim is not None:
                mc = model_y2cov(dim_embed=cov_dim)
                randomize_affine(mc, gen)
                le.y2cov_type, le.model_mlp_y2cov = "resnet", nn.DataParallel(mc)
                rec["cov"] = le.fn_y2cov(labels).clone()
        out[name] = rec
        print(name, {k: tuple(v.shape) for k, v in rec.items()})
    torch.save(out, os.path.join(HERE, "label_mlp.pt"))


if __name__ == "__main__":
    main()
