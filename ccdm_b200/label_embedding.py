"""Continuous-label embedding hooks with the reference's interface (CCDM_unified/label_embedding.py:212-227,
:861 ``fn_y2h``, :1035 ``fn_y2cov``).

The hooks are the boundary of the hot path: arbitrary user callables ``labels -> [B, h_dim]`` /
``labels -> [B, C*H*W]`` plug into ``Trainer`` / ``GaussianDiffusion`` unchanged.  This class provides the
training-free families (sinusoidal, Gaussian Fourier; scalar or multi-dimensional labels combined by mean or
softmax weights).  They are a few tiny elementwise ops evaluated once per batch, so they stay in PyTorch
(SURVEY.md section 2.1 #9).  The learned families ("resnet" nets, attention / cross combiners) need the
out-of-scope embedding pre-training and raise.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class GaussianFourierProjection(nn.Module):     # label_embedding.py:18-34
    def __init__(self, embed_dim, scale=30.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embed_dim // 2) * scale, requires_grad=False)
        self.embed_dim = embed_dim

    def forward(self, x):
        proj = x[:, None] * (self.W[None, :]).to(x.device) * 2 * np.pi
        return torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1).view(len(proj), self.embed_dim)


_FREQS = {}


def _sinusoid(labels, dim):
    half = dim // 2
    key = (half, labels.device)
    if key not in _FREQS:      # uploaded once per device: no pageable host->device copy inside a captured training step
        _FREQS[key] = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(labels.device)
    freqs = _FREQS[key]
    args = labels.view(len(labels))[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class LabelEmbed:
    def __init__(self, dataset=None, path_y2h=None, path_y2cov=None, y2h_type="sinusoidal", y2cov_type=None,
                 h_dim=128, cov_dim=None, batch_size=128, nc=3, device=None, label_dim=1, dim_combination="cross"):
        assert y2h_type in ["resnet", "sinusoidal", "gaussian"]
        if y2cov_type is not None:
            assert y2cov_type in ["resnet", "sinusoidal", "gaussian"]
        if "resnet" in (y2h_type, y2cov_type):
            raise NotImplementedError("the learned y2h / y2cov networks need the embedding pre-training stage "
                                      "(label_embedding.py:345-859), which is outside the denoiser hot path; pass any "
                                      "callable as fn_y2h / fn_y2cov instead")
        self.dataset = dataset
        self.device = device if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.y2h_type, self.y2cov_type = y2h_type, y2cov_type
        self.h_dim = h_dim
        self.cov_dim = cov_dim if cov_dim is not None else 64 ** 2 * nc
        self.nc, self.label_dim, self.dim_combination = nc, label_dim, dim_combination
        if label_dim > 1:
            if dim_combination == "weighted":
                self.dim_weights = (torch.ones(label_dim) / label_dim).to(self.device)
            elif dim_combination != "mean":
                raise NotImplementedError(f"dim_combination={dim_combination!r} uses a learned combiner network")
        self._gfp = {}

    def _embed(self, labels, kind, dim, post, cache_tag):
        multi = len(labels.shape) > 1 and labels.shape[1] > 1
        if not multi:
            if kind == "sinusoidal":
                return post(_sinusoid(labels, dim))
            # the reference builds a fresh random projection on every scalar-label call (label_embedding.py:1022-1026)
            return post(GaussianFourierProjection(embed_dim=dim).to(labels.device)(labels))
        cols = range(labels.shape[1])
        if labels.shape[1] > 20:
            cols = list(range(0, labels.shape[1], max(1, labels.shape[1] // 10)))[:10]
        embs = []
        for d in cols:
            col = labels[:, d].view(len(labels))
            if kind == "sinusoidal":
                embs.append(post(_sinusoid(col, dim)))
            else:
                key = (cache_tag, d)
                if key not in self._gfp:
                    self._gfp[key] = GaussianFourierProjection(embed_dim=dim).to(labels.device)
                embs.append(post(self._gfp[key](col.unsqueeze(-1))))
        stacked = torch.stack(embs)
        if self.dim_combination == "weighted":
            return torch.sum(stacked * F.softmax(self.dim_weights, dim=0).view(-1, 1, 1), dim=0)
        return torch.mean(stacked, dim=0)

    def fn_y2h(self, labels):
        """labels [B] / [B,1] / [B,D] -> [B, h_dim] in [0,1]  (label_embedding.py:861-1033)."""
        return self._embed(labels, self.y2h_type, self.h_dim, lambda e: (e + 1) / 2, "h")

    def fn_y2cov(self, labels):
        """labels -> positive [B, cov_dim]  (label_embedding.py:1035-1178)."""
        return self._embed(labels, self.y2cov_type, self.cov_dim, lambda e: e + 1, "cov")
