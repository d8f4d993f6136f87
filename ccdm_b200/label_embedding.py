"""Continuous-label embedding hooks with the reference's interface (CCDM_unified/label_embedding.py:212-227,
:861 ``fn_y2h``, :1035 ``fn_y2cov``).

The hooks are the boundary of the hot path: arbitrary user callables ``labels -> [B, h_dim]`` /
``labels -> [B, C*H*W]`` plug into ``Trainer`` / ``GaussianDiffusion`` unchanged.  This class provides the
training-free families (sinusoidal, Gaussian Fourier; a few tiny elementwise ops evaluated once per batch: PyTorch,
SURVEY.md section 2.1 #9) and the INFERENCE of the learned family every reference script uses
(``--y2h_embed_type resnet --y2cov_embed_type resnet``, scripts/RC64/linux/run_ccdm.sh:25): the label MLPs ``model_y2h`` /
``model_y2cov`` (models/resnet_y2h.py:143-173, models/resnet_y2cov.py:149-179: Linear -> GroupNorm(8) -> ReLU x4, Linear ->
ReLU; the last y2cov layer is a [B,4096] x [4096, C*H*W] product) run as C-ABI kernels (``ccdm_linear_small`` +
``ccdm_groupnorm_rows``, fp32 like the reference).  Their weights come from the reference's own checkpoints
(``ckpt_mlp_y2h_epoch_500.pth`` / ``y2h_ckpt_in_train/mlp_y2h_checkpoint_epoch_500.pth`` under ``path_y2h``, same for y2cov;
label_embedding.py:383-424,656-700) or from a state dict; TRAINING those networks (the ResNet-34 embedding pre-training,
label_embedding.py:345-859) is outside the denoiser hot path and is not built: a missing checkpoint raises.
Multi-dimensional labels combine the per-dimension embeddings by mean / softmax weights / the reference's (randomly
initialised, never trained) attention and cross networks for ``h``; the ``cross_attention`` combiner and the cov-sized
combiner networks (cov_dim x 2 cov_dim Linear layers) raise.
"""
import math
import os

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


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("ccdm_b200: the learned label MLPs run on the CUDA path only (no CPU fallback)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _LabelMLP(nn.Module):
    """Parameter holder with the reference's ``main`` Sequential layout (state-dict keys ``main.{0,1,3,4,...}.weight``);
    ``forward`` runs the C-ABI kernels: per layer ``ccdm_linear_small`` (fp32 row kernel) then ``ccdm_groupnorm_rows``
    (GroupNorm(8) + ReLU in place) -- no torch arithmetic, and no CPU path."""

    def __init__(self, widths, num_groups):
        super().__init__()
        layers = []
        for i in range(len(widths) - 1):
            layers.append(nn.Linear(widths[i], widths[i + 1]))
            if i < len(widths) - 2:
                layers.append(nn.GroupNorm(num_groups, widths[i + 1]))
            layers.append(nn.ReLU())
        self.main = nn.Sequential(*layers)
        self.num_groups = num_groups

    @torch.no_grad()
    def forward(self, y):
        from . import _lib as L
        y = y.reshape(-1, 1).float() + 1e-8             # resnet_y2h.py:170 / resnet_y2cov.py:176
        _require_cuda(y)
        stream = _stream()
        x = y.contiguous()
        mods = list(self.main)
        i = 0
        while i < len(mods):
            lin = mods[i]
            gn = mods[i + 1] if isinstance(mods[i + 1], nn.GroupNorm) else None
            out = torch.empty(x.shape[0], lin.out_features, device=x.device, dtype=torch.float32)
            L.check(L.lib().ccdm_linear_small(x.data_ptr(), x.shape[0], lin.in_features, lin.weight.data_ptr(),
                                              lin.bias.data_ptr(), lin.out_features, None, None, None, None, 0,
                                              L.ACT_NONE if gn is not None else L.ACT_RELU, out.data_ptr(),
                                              lin.out_features, stream), "label MLP linear")
            if gn is not None:
                L.check(L.lib().ccdm_groupnorm_rows(out.data_ptr(), out.shape[0], gn.num_channels, gn.num_groups,
                                                    gn.weight.data_ptr(), gn.bias.data_ptr(), float(gn.eps), L.ACT_RELU,
                                                    stream), "label MLP groupnorm")
            x = out
            i += 3 if gn is not None else 2
        return x


def model_y2h(dim_embed=128, num_groups=8):         # models/resnet_y2h.py:143-173
    return _LabelMLP([1] + [dim_embed] * 5, num_groups)


def model_y2cov(dim_embed=64 * 64 * 3, num_groups=8):   # models/resnet_y2cov.py:149-179
    return _LabelMLP([1, 512, 1024, 2048, 4096, dim_embed], num_groups)


class LabelEmbed:
    def __init__(self, dataset=None, path_y2h=None, path_y2cov=None, y2h_type="sinusoidal", y2cov_type=None,
                 h_dim=128, cov_dim=None, batch_size=128, nc=3, device=None, label_dim=1, dim_combination="cross"):
        assert y2h_type in ["resnet", "sinusoidal", "gaussian"]
        if y2cov_type is not None:
            assert y2cov_type in ["resnet", "sinusoidal", "gaussian"]
        self.dataset = dataset
        self.path_y2h, self.path_y2cov = path_y2h, path_y2cov
        self.device = device if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.y2h_type, self.y2cov_type = y2h_type, y2cov_type
        self.h_dim = h_dim
        self.cov_dim = cov_dim if cov_dim is not None else 64 ** 2 * nc
        self.nc, self.label_dim, self.dim_combination = nc, label_dim, dim_combination
        if label_dim > 1:                               # label_embedding.py:283-341 (same construction order: same RNG draws)
            if dim_combination == "attention":
                self.h_attention_net = nn.Sequential(nn.Linear(h_dim, h_dim // 2), nn.ReLU(),
                                                     nn.Linear(h_dim // 2, 1)).to(self.device)
                if y2cov_type is not None:
                    raise NotImplementedError("cov_attention_net (cov_dim x cov_dim/2 Linear) is not built")
            elif dim_combination == "weighted":
                self.dim_weights = (torch.ones(label_dim) / label_dim).to(self.device)
            elif dim_combination == "cross":
                self.h_cross_net = nn.Sequential(nn.Linear(h_dim * label_dim, h_dim * 2), nn.LayerNorm(h_dim * 2), nn.ReLU(),
                                                 nn.Linear(h_dim * 2, h_dim), nn.LayerNorm(h_dim)).to(self.device)
                if y2cov_type is not None:
                    raise NotImplementedError("cov_cross_net (cov_dim*D x 2 cov_dim Linear) is not built")
            elif dim_combination != "mean":
                raise NotImplementedError(f"dim_combination={dim_combination!r} is not built")
        self._gfp = {}
        self.model_mlp_y2h = self.model_mlp_y2cov = None
        if y2h_type == "resnet":
            self.model_mlp_y2h = model_y2h(dim_embed=self.h_dim).to(self.device).eval()
            self._load_mlp(self.model_mlp_y2h, path_y2h, "y2h")
        if y2cov_type == "resnet":
            self.model_mlp_y2cov = model_y2cov(dim_embed=self.cov_dim).to(self.device).eval()
            self._load_mlp(self.model_mlp_y2cov, path_y2cov, "y2cov")

    # ------------------------------------------------------------------ learned label MLPs
    @staticmethod
    def find_mlp_checkpoint(path, name, epochs_mlp=500):
        """The reference's two locations (label_embedding.py:406-424 / :680-697)."""
        if path is None:
            return None
        for cand in (os.path.join(path, f"{name}_ckpt_in_train", f"mlp_{name}_checkpoint_epoch_{epochs_mlp}.pth"),
                     os.path.join(path, f"ckpt_mlp_{name}_epoch_{epochs_mlp}.pth")):
            if os.path.isfile(cand):
                return cand
        return None

    def _load_mlp(self, model, path, name):
        ckpt = self.find_mlp_checkpoint(path, name)
        if ckpt is None:
            raise FileNotFoundError(
                f"no trained {name} label MLP under {path!r} (expected ckpt_mlp_{name}_epoch_500.pth or "
                f"{name}_ckpt_in_train/mlp_{name}_checkpoint_epoch_500.pth, as written by the reference's LabelEmbed); the "
                "embedding pre-training itself is outside the denoiser hot path -- run it with the reference, or call "
                f"load_mlp_state('{name}', state_dict)")
        state = torch.load(ckpt, map_location=self.device, weights_only=True)["net_state_dict"]
        self.load_mlp_state(name, state)

    def load_mlp_state(self, name, state):
        """``state``: the MLP's state dict, with or without the ``module.`` prefix of the reference's nn.DataParallel wrapper."""
        model = {"y2h": self.model_mlp_y2h, "y2cov": self.model_mlp_y2cov}[name]
        if model is None:
            model = (model_y2h(dim_embed=self.h_dim) if name == "y2h" else model_y2cov(dim_embed=self.cov_dim)).to(self.device).eval()
            setattr(self, f"model_mlp_{name}", model)
        model.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in state.items()}, strict=True)

    def _embed(self, labels, kind, dim, post, cache_tag):
        if kind == "resnet":
            return self._embed_mlp(labels, self.model_mlp_y2h if cache_tag == "h" else self.model_mlp_y2cov, cache_tag)
        return self._embed_free(labels, kind, dim, post, cache_tag)

    def _combine(self, stacked, cache_tag):
        """[D, B, dim] -> [B, dim] (label_embedding.py:960-1005 / :1110-1143)."""
        if self.dim_combination == "weighted":
            return torch.sum(stacked * F.softmax(self.dim_weights, dim=0).view(-1, 1, 1), dim=0)
        if self.dim_combination == "attention" and cache_tag == "h":
            x = stacked.permute(1, 0, 2)
            with torch.no_grad():
                w = F.softmax(self.h_attention_net(x).squeeze(-1), dim=1).unsqueeze(-1)
            return torch.sum(x * w, dim=1)
        if self.dim_combination == "cross" and cache_tag == "h":
            x = stacked.permute(1, 0, 2)
            with torch.no_grad():
                return self.h_cross_net(x.reshape(x.shape[0], -1))
        return torch.mean(stacked, dim=0)

    def _embed_mlp(self, labels, model, cache_tag):
        multi = len(labels.shape) > 1 and labels.shape[1] > 1
        if not multi:
            return model(labels)
        cols = range(labels.shape[1])
        if labels.shape[1] > 20:
            cols = list(range(0, labels.shape[1], max(1, labels.shape[1] // 10)))[:10]
        return self._combine(torch.stack([model(labels[:, d].reshape(-1, 1)) for d in cols]), cache_tag)

    def _embed_free(self, labels, kind, dim, post, cache_tag):
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
        return self._combine(torch.stack(embs), cache_tag)

    def fn_y2h(self, labels):
        """labels [B] / [B,1] / [B,D] -> [B, h_dim] in [0,1]  (label_embedding.py:861-1033)."""
        return self._embed(labels, self.y2h_type, self.h_dim, lambda e: (e + 1) / 2, "h")

    def fn_y2cov(self, labels):
        """labels -> positive [B, cov_dim]  (label_embedding.py:1035-1178)."""
        return self._embed(labels, self.y2cov_type, self.cov_dim, lambda e: e + 1, "cov")
