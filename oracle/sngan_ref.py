"""fp32 restatement of the one-step generator ``sngan_generator`` in eval mode
(CCDM_unified/models/sngan.py:19-139), functional over a state dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference's own module by
``tests/golden/make_golden_sngan.py`` -> ``tests/golden/sngan.pt`` (replayed by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass(frozen=True)
class GenSpec:
    dim_z: int = 128
    dim_embed: int = 128
    nc: int = 3
    img_size: int = 64
    gene_ch: int = 32
    ch_multi: Tuple[int, ...] = (16, 8, 4, 2, 1)

    @property
    def init_size(self) -> int:                       # sngan.py:103-106
        return 4 if self.img_size in (64, 128) else 6

    @property
    def block_channels(self):                         # sngan.py:113-118
        c = [self.gene_ch * m for m in self.ch_multi[:5]]
        pairs = [(c[i], c[i + 1]) for i in range(4)]
        if self.img_size in (128, 192):
            pairs.append((c[4], self.gene_ch))
        return pairs


def _bn_shapes(name, c, affine):
    s = {}
    if affine:
        s[f"{name}.weight"] = (c,)
        s[f"{name}.bias"] = (c,)
    s[f"{name}.running_mean"] = (c,)
    s[f"{name}.running_var"] = (c,)
    s[f"{name}.num_batches_tracked"] = ()
    return s


def state_dict_shapes(spec: GenSpec) -> Dict[str, tuple]:
    """Keys in the reference's registration order (dense, final, genblock0..), incl. the shared-module aliases."""
    s: Dict[str, tuple] = {}
    c0 = spec.gene_ch * spec.ch_multi[0]
    s["dense.weight"] = (spec.init_size ** 2 * c0, spec.dim_z)
    s["dense.bias"] = (spec.init_size ** 2 * c0,)
    s.update(_bn_shapes("final.0", spec.gene_ch, True))
    s["final.2.weight"] = (spec.nc, spec.gene_ch, 3, 3)
    s["final.2.bias"] = (spec.nc,)
    for i, (ci, co) in enumerate(spec.block_channels):
        n = f"genblock{i}"
        s[f"{n}.conv1.weight"] = (co, ci, 3, 3)
        s[f"{n}.conv1.bias"] = (co,)
        s[f"{n}.conv2.weight"] = (co, co, 3, 3)
        s[f"{n}.conv2.bias"] = (co,)
        for j, c in ((1, ci), (2, co)):
            s.update(_bn_shapes(f"{n}.condbn{j}.bn", c, False))
            s[f"{n}.condbn{j}.embed_gamma.weight"] = (c, spec.dim_embed)
            s[f"{n}.condbn{j}.embed_beta.weight"] = (c, spec.dim_embed)
        s.update(_bn_shapes(f"{n}.model.0", ci, True))
        s[f"{n}.model.3.weight"] = (co, ci, 3, 3)            # alias of conv1
        s[f"{n}.model.3.bias"] = (co,)
        s.update(_bn_shapes(f"{n}.model.4", co, True))
        s[f"{n}.model.6.weight"] = (co, co, 3, 3)            # alias of conv2
        s[f"{n}.model.6.bias"] = (co,)
        s[f"{n}.bypass_conv.weight"] = (co, ci, 1, 1)
        s[f"{n}.bypass_conv.bias"] = (co,)
        s[f"{n}.bypass.1.weight"] = (co, ci, 1, 1)           # alias of bypass_conv
        s[f"{n}.bypass.1.bias"] = (co,)
    return s


_ALIASES = {"model.3.": "conv1.", "model.6.": "conv2.", "bypass.1.": "bypass_conv."}


def make_state_dict(spec: GenSpec, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic synthetic weights (a pure function of key, shape and seed); aliased keys share their tensor."""
    out: Dict[str, Tensor] = {}
    for key, shape in state_dict_shapes(spec).items():
        src = key
        for a, t in _ALIASES.items():
            src = src.replace(a, t)
        if src != key:
            out[key] = out[src]
            continue
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.tensor(5, dtype=torch.long)
        elif key.endswith("running_var"):
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            out[key] = 0.3 * torch.randn(shape, generator=g)
        elif key.endswith("embed_gamma.weight") or key.endswith("embed_beta.weight"):
            out[key] = 0.5 * torch.randn(shape, generator=g) / math.sqrt(shape[1])
        elif len(shape) == 1 and key.endswith("weight"):
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            out[key] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for n in shape[1:]:
                fan_in *= n
            scale = 1.5 if key.startswith("final.2") else 0.7           # keeps the tanh output out of saturation
            out[key] = torch.randn(shape, generator=g) * scale / math.sqrt(fan_in)
    return out


def _bn_eval(sd, name, x, affine, stats=None):
    """nn.BatchNorm2d, sngan.py:23,54,58,122: eval mode (running statistics) or, with ``stats`` (a dict that receives the
    updated running statistics), training mode: batch statistics, momentum 0.1, unbiased variance in the running update."""
    w = sd[f"{name}.weight"] if affine else None
    b = sd[f"{name}.bias"] if affine else None
    if stats is None:
        return F.batch_norm(x, sd[f"{name}.running_mean"], sd[f"{name}.running_var"], w, b, False, 0.0, 1e-5)
    rm, rv = sd[f"{name}.running_mean"].clone(), sd[f"{name}.running_var"].clone()
    out = F.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5)
    stats[f"{name}.running_mean"], stats[f"{name}.running_var"] = rm, rv
    return out


def _condbn(sd, name, x, y, stats=None):
    """ConditionalBatchNorm2d.forward, sngan.py:28-36."""
    out = _bn_eval(sd, f"{name}.bn", x, False, stats)
    gamma = F.linear(y, sd[f"{name}.embed_gamma.weight"])[:, :, None, None]
    beta = F.linear(y, sd[f"{name}.embed_beta.weight"])[:, :, None, None]
    return out + out * gamma + beta


def generator_forward(sd: Dict[str, Tensor], spec: GenSpec, z: Tensor, y: Tensor, stats=None) -> Tensor:
    """sngan_generator.forward, sngan.py:130-139, conditional branch (sngan.py:73-83); eval mode, or training mode when a
    ``stats`` dict is passed (it receives the updated running statistics)."""
    c0 = spec.gene_ch * spec.ch_multi[0]
    out = F.linear(z.view(z.size(0), -1), sd["dense.weight"], sd["dense.bias"]).view(-1, c0, spec.init_size, spec.init_size)
    for i in range(len(spec.block_channels)):
        n = f"genblock{i}"
        h = F.relu(_condbn(sd, f"{n}.condbn1", out, y, stats))
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        h = F.conv2d(h, sd[f"{n}.conv1.weight"], sd[f"{n}.conv1.bias"], padding=1)
        h = F.relu(_condbn(sd, f"{n}.condbn2", h, y, stats))
        h = F.conv2d(h, sd[f"{n}.conv2.weight"], sd[f"{n}.conv2.bias"], padding=1)
        by = F.conv2d(F.interpolate(out, scale_factor=2, mode="nearest"), sd[f"{n}.bypass_conv.weight"],
                      sd[f"{n}.bypass_conv.bias"])
        out = h + by
    out = F.relu(_bn_eval(sd, "final.0", out, True, stats))
    return torch.tanh(F.conv2d(out, sd["final.2.weight"], sd["final.2.bias"], padding=1))
