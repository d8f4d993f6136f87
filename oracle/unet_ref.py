"""Functional fp32 restatement of the CCDM_unified conditional UNet (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The network is evaluated straight from a ``state_dict`` whose keys and shapes
are the reference's (``CCDM_unified/models/unet.py:244-348``), with plain
``torch.nn.functional`` calls in fp32.  Every function cites the reference
lines it restates.  Pinned by ``tests/golden`` (outputs of the reference's own
modules); it holds no code from the reference.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LIN_ATTN_HEADS = 4      # fixed in the reference irrespective of CLI: unet.py:190,325,340
LIN_ATTN_DIM_HEAD = 32


@dataclass(frozen=True)
class UnetSpec:
    """Constructor arguments that shape the network (unet.py:245-260)."""
    dim: int = 64
    dim_mults: Tuple[int, ...] = (1, 2, 4, 8)
    in_channels: int = 3
    embed_input_dim: int = 128
    attn_dim_head: int = 32
    attn_heads: int = 4
    out_dim: Optional[int] = None
    init_dim: Optional[int] = None

    @property
    def stem_dim(self) -> int:
        return self.init_dim if self.init_dim is not None else self.dim

    @property
    def level_dims(self):
        widths = [self.stem_dim] + [self.dim * m for m in self.dim_mults]
        return list(zip(widths[:-1], widths[1:]))          # unet.py:273-274

    @property
    def emb_dim(self) -> int:
        return self.dim * 4                                 # time_dim == cond_emb_dim, unet.py:278,307

    @property
    def out_channels(self) -> int:
        return self.out_dim if self.out_dim is not None else self.in_channels


# --------------------------------------------------------------------------- shapes / weights

def _resblock_shapes(name, cin, cout, emb2):
    s = {
        f"{name}.tc_mlp.1.weight": (2 * cout, emb2),
        f"{name}.tc_mlp.1.bias": (2 * cout,),
        f"{name}.block1.proj.weight": (cout, cin, 3, 3),
        f"{name}.block1.proj.bias": (cout,),
        f"{name}.block1.norm.g": (1, cout, 1, 1),
        f"{name}.block2.proj.weight": (cout, cout, 3, 3),
        f"{name}.block2.proj.bias": (cout,),
        f"{name}.block2.norm.g": (1, cout, 1, 1),
    }
    if cin != cout:                                         # unet.py:165
        s[f"{name}.res_conv.weight"] = (cout, cin, 1, 1)
        s[f"{name}.res_conv.bias"] = (cout,)
    return s


def _linattn_shapes(name, c):
    hid = LIN_ATTN_HEADS * LIN_ATTN_DIM_HEAD
    return {
        f"{name}.fn.fn.to_qkv.weight": (3 * hid, c, 1, 1),
        f"{name}.fn.fn.to_out.0.weight": (c, hid, 1, 1),
        f"{name}.fn.fn.to_out.0.bias": (c,),
        f"{name}.fn.fn.to_out.1.g": (1, c, 1, 1),
        f"{name}.fn.norm.g": (1, c, 1, 1),
    }


def state_dict_shapes(spec: UnetSpec) -> Dict[str, tuple]:
    """Key -> shape map of ``Unet(...).state_dict()`` (unet.py:263-348)."""
    d, e = spec.dim, spec.emb_dim
    s: Dict[str, tuple] = {
        "null_cond_emb": (d,),
        "init_conv.weight": (spec.stem_dim, spec.in_channels, 7, 7),
        "init_conv.bias": (spec.stem_dim,),
        "time_mlp.1.weight": (e, d), "time_mlp.1.bias": (e,),
        "time_mlp.3.weight": (e, e), "time_mlp.3.bias": (e,),
    }
    for nm, (o, i) in (("cond_mlp_1", (d, spec.embed_input_dim)), ("cond_mlp_2", (e, d))):
        s.update({f"{nm}.0.weight": (o, i), f"{nm}.0.bias": (o,), f"{nm}.1.weight": (o,),
                  f"{nm}.1.bias": (o,), f"{nm}.1.running_mean": (o,), f"{nm}.1.running_var": (o,),
                  f"{nm}.1.num_batches_tracked": ()})
    lv = spec.level_dims
    for k, (ci, co) in enumerate(lv):
        last = k == len(lv) - 1
        s.update(_resblock_shapes(f"downs.{k}.0", ci, ci, 2 * e))
        s.update(_resblock_shapes(f"downs.{k}.1", ci, ci, 2 * e))
        s.update(_linattn_shapes(f"downs.{k}.2", ci))
        s[f"downs.{k}.3.weight"] = (co, ci, 3, 3) if last else (co, ci, 4, 4)
        s[f"downs.{k}.3.bias"] = (co,)
    mid = lv[-1][1]
    hid = spec.attn_heads * spec.attn_dim_head
    s.update(_resblock_shapes("mid_block1", mid, mid, 2 * e))
    s.update({"mid_attn.fn.fn.to_qkv.weight": (3 * hid, mid, 1, 1),
              "mid_attn.fn.fn.to_out.weight": (mid, hid, 1, 1),
              "mid_attn.fn.fn.to_out.bias": (mid,),
              "mid_attn.fn.norm.g": (1, mid, 1, 1)})
    s.update(_resblock_shapes("mid_block2", mid, mid, 2 * e))
    for k, (ci, co) in enumerate(reversed(lv)):
        last = k == len(lv) - 1
        s.update(_resblock_shapes(f"ups.{k}.0", co + ci, co, 2 * e))
        s.update(_resblock_shapes(f"ups.{k}.1", co + ci, co, 2 * e))
        s.update(_linattn_shapes(f"ups.{k}.2", co))
        nm = f"ups.{k}.3" if last else f"ups.{k}.3.1"       # Sequential(Upsample, Conv) unless last
        s[f"{nm}.weight"] = (ci, co, 3, 3)
        s[f"{nm}.bias"] = (ci,)
    s.update(_resblock_shapes("final_res_block", 2 * spec.stem_dim, spec.stem_dim, 2 * e))
    s["final_conv.weight"] = (spec.out_channels, spec.stem_dim, 1, 1)
    s["final_conv.bias"] = (spec.out_channels,)
    return s


def make_state_dict(spec: UnetSpec, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic synthetic weights that depend only on (key, shape, seed).

    Not an init scheme of the reference: it lets the golden generator and the
    tests rebuild identical weights without shipping them.
    """
    out = {}
    for key, shape in state_dict_shapes(spec).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.tensor(3, dtype=torch.long)
        elif key.endswith("running_var"):
            out[key] = (0.5 + torch.rand(shape, generator=g)).to(dtype)
        elif key.endswith("running_mean"):
            out[key] = (0.2 * torch.randn(shape, generator=g)).to(dtype)
        elif key.endswith(".g") or key.endswith(".1.weight") and len(shape) == 1:
            out[key] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif key == "null_cond_emb":
            out[key] = (-torch.randn(shape, generator=g).abs()).to(dtype)
        elif key.endswith("bias"):
            out[key] = (0.05 * torch.randn(shape, generator=g)).to(dtype)
        else:
            fan_in = 1
            for n in shape[1:]:
                fan_in *= n
            out[key] = (torch.randn(shape, generator=g) / math.sqrt(fan_in)).to(dtype)
    return out


# --------------------------------------------------------------------------- building blocks

def channel_rms(x: Tensor, g: Tensor) -> Tensor:
    """unet.py:88-89 -- per-pixel L2 normalisation over channels, times g*sqrt(C)."""
    return F.normalize(x, dim=1) * g * (x.shape[1] ** 0.5)


def time_features(t: Tensor, dim: int) -> Tensor:
    """unet.py:107-115 -- sin|cos, exponent divided by (half-1)."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -k)
    ang = t.view(-1)[:, None] * freq[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def _bn1d(sd, name, x, training, momentum_out=None):
    """nn.BatchNorm1d (unet.py:300,310): batch stats in training, running stats in eval."""
    w, b = sd[f"{name}.weight"], sd[f"{name}.bias"]
    if training:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        if momentum_out is not None:
            n = x.shape[0]
            momentum_out[f"{name}.running_mean"] = 0.9 * sd[f"{name}.running_mean"] + 0.1 * mean
            momentum_out[f"{name}.running_var"] = 0.9 * sd[f"{name}.running_var"] + 0.1 * var * n / max(n - 1, 1)
    else:
        mean, var = sd[f"{name}.running_mean"], sd[f"{name}.running_var"]
    return (x - mean) / torch.sqrt(var + 1e-5) * w + b


def _conv_block(sd, name, x, scale_shift=None):
    """Block.forward, unet.py:143-152."""
    y = F.conv2d(x, sd[f"{name}.proj.weight"], sd[f"{name}.proj.bias"], padding=1)
    y = channel_rms(y, sd[f"{name}.norm.g"])
    if scale_shift is not None:
        scale, shift = scale_shift
        y = y * (scale + 1) + shift
    return F.silu(y)


def _resblock(sd, name, x, t_emb, c_emb):
    """ResnetBlock.forward, unet.py:167-187."""
    tc = torch.cat((t_emb, c_emb), dim=1)
    tc = F.linear(F.silu(tc), sd[f"{name}.tc_mlp.1.weight"], sd[f"{name}.tc_mlp.1.bias"])
    scale, shift = tc[:, :, None, None].chunk(2, dim=1)
    h = _conv_block(sd, f"{name}.block1", x, (scale, shift))
    h = _conv_block(sd, f"{name}.block2", h)
    if f"{name}.res_conv.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.res_conv.weight"], sd[f"{name}.res_conv.bias"])
    return h + x


def _linear_attention(sd, name, x):
    """Residual(PreNorm(LinearAttention)), unet.py:66-72,91-99,202-216."""
    b, c, hh, ww = x.shape
    heads, dh = LIN_ATTN_HEADS, LIN_ATTN_DIM_HEAD
    xn = channel_rms(x, sd[f"{name}.fn.norm.g"])
    qkv = F.conv2d(xn, sd[f"{name}.fn.fn.to_qkv.weight"]).view(b, 3, heads, dh, hh * ww)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]               # each [b, heads, dh, n]
    q = q.softmax(dim=-2) * dh ** -0.5                      # over head channels
    k = k.softmax(dim=-1)                                   # over tokens
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    o = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, heads * dh, hh, ww)
    o = F.conv2d(o, sd[f"{name}.fn.fn.to_out.0.weight"], sd[f"{name}.fn.fn.to_out.0.bias"])
    o = channel_rms(o, sd[f"{name}.fn.fn.to_out.1.g"])
    return o + x


def _mid_attention(sd, name, x, heads, dh):
    """Residual(PreNorm(Attention)), unet.py:228-240."""
    b, c, hh, ww = x.shape
    xn = channel_rms(x, sd[f"{name}.fn.norm.g"])
    qkv = F.conv2d(xn, sd[f"{name}.fn.fn.to_qkv.weight"]).view(b, 3, heads, dh, hh * ww)
    q, k, v = qkv[:, 0] * dh ** -0.5, qkv[:, 1], qkv[:, 2]
    att = torch.einsum("bhdi,bhdj->bhij", q, k).softmax(dim=-1)
    o = torch.einsum("bhij,bhdj->bhid", att, v)             # [b, heads, n, dh]
    o = o.permute(0, 1, 3, 2).reshape(b, heads * dh, hh, ww)
    o = F.conv2d(o, sd[f"{name}.fn.fn.to_out.weight"], sd[f"{name}.fn.fn.to_out.bias"])
    return o + x


def _strip(sd):
    """Accept reference checkpoints keyed ``model.module.*`` / ``module.*``."""
    for pre in ("model.module.", "module."):
        if any(k.startswith(pre) for k in sd):
            return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    return sd


# --------------------------------------------------------------------------- forward

def unet_forward(sd, spec: UnetSpec, x: Tensor, t: Tensor, labels_emb: Tensor, *,
                 cond_drop_prob: float, training: bool = False,
                 keep_mask: Optional[Tensor] = None, bn_updates: Optional[dict] = None,
                 return_bottleneck: bool = False) -> Tensor:
    """Unet.forward, unet.py:382-455.

    ``keep_mask`` here is the mask the reference draws *inside* the UNet
    (unet.py:404); when None and 0 < p < 1 it is drawn the same way
    (uniform_(0,1) < 1-p) so that the RNG stream matches.  The reference's own
    ``keep_mask`` argument is dead code and has no counterpart.
    """
    sd = _strip(sd)
    b = x.shape[0]
    c = F.relu(_bn1d(sd, "cond_mlp_1.1",
                     F.linear(labels_emb, sd["cond_mlp_1.0.weight"], sd["cond_mlp_1.0.bias"]),
                     training, bn_updates))
    if cond_drop_prob > 0:
        if keep_mask is None:
            p = 1 - cond_drop_prob
            if p == 1:
                keep_mask = torch.ones(b, dtype=torch.bool, device=x.device)
            elif p == 0:
                keep_mask = torch.zeros(b, dtype=torch.bool, device=x.device)
            else:
                keep_mask = torch.zeros(b, device=x.device).float().uniform_(0, 1) < p
        c = torch.where(keep_mask[:, None], c, sd["null_cond_emb"][None, :].expand(b, -1))
    c = F.relu(_bn1d(sd, "cond_mlp_2.1",
                     F.linear(c, sd["cond_mlp_2.0.weight"], sd["cond_mlp_2.0.bias"]),
                     training, bn_updates))

    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    stem = x
    te = time_features(t, spec.dim)
    te = F.linear(te, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    te = F.linear(F.gelu(te), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])

    skips = []
    nlev = len(spec.level_dims)
    for k in range(nlev):
        x = _resblock(sd, f"downs.{k}.0", x, te, c)
        skips.append(x)
        x = _resblock(sd, f"downs.{k}.1", x, te, c)
        x = _linear_attention(sd, f"downs.{k}.2", x)
        skips.append(x)
        w, bias = sd[f"downs.{k}.3.weight"], sd[f"downs.{k}.3.bias"]
        x = F.conv2d(x, w, bias, padding=1) if k == nlev - 1 else F.conv2d(x, w, bias, stride=2, padding=1)

    x = _resblock(sd, "mid_block1", x, te, c)
    if return_bottleneck:
        return x
    x = _mid_attention(sd, "mid_attn", x, spec.attn_heads, spec.attn_dim_head)
    x = _resblock(sd, "mid_block2", x, te, c)

    for k in range(nlev):
        x = _resblock(sd, f"ups.{k}.0", torch.cat((x, skips.pop()), dim=1), te, c)
        x = _resblock(sd, f"ups.{k}.1", torch.cat((x, skips.pop()), dim=1), te, c)
        x = _linear_attention(sd, f"ups.{k}.2", x)
        if k == nlev - 1:
            x = F.conv2d(x, sd[f"ups.{k}.3.weight"], sd[f"ups.{k}.3.bias"], padding=1)
        else:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[f"ups.{k}.3.1.weight"], sd[f"ups.{k}.3.1.bias"], padding=1)

    x = _resblock(sd, "final_res_block", torch.cat((x, stem), dim=1), te, c)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])


def cfg_combine(cond: Tensor, null: Tensor, cond_scale: float, rescaled_phi: float,
                remove_parallel_component: bool = True, keep_parallel_frac: float = 0.0) -> Tensor:
    """Guidance arithmetic of forward_with_cond_scale + project, unet.py:51-62,365-380."""
    upd = cond - null
    if remove_parallel_component:
        shp = upd.shape
        u, y = upd.flatten(1).double(), cond.flatten(1).double()
        unit = F.normalize(y, dim=-1)
        par = (u * unit).sum(-1, keepdim=True) * unit
        orth = u - par
        upd = orth.view(shp).to(cond.dtype) + par.view(shp).to(cond.dtype) * keep_parallel_frac
    scaled = cond + upd * (cond_scale - 1.0)
    if rescaled_phi == 0.0:
        return scaled
    dims = tuple(range(1, scaled.ndim))
    resc = scaled * (cond.std(dim=dims, keepdim=True) / scaled.std(dim=dims, keepdim=True))
    return resc * rescaled_phi + scaled * (1.0 - rescaled_phi)


def unet_forward_cfg(sd, spec, x, t, labels_emb, *, cond_scale=1.0, rescaled_phi=0.0,
                     remove_parallel_component=True, keep_parallel_frac=0.0, training=False):
    """forward_with_cond_scale, unet.py:350-380.  Returns (guided, null).

    The reference returns a bare tensor when cond_scale == 1 and its only caller
    then fails to unpack it; here that case returns (cond, None).
    """
    cond = unet_forward(sd, spec, x, t, labels_emb, cond_drop_prob=0.0, training=training)
    if cond_scale == 1:
        return cond, None
    null = unet_forward(sd, spec, x, t, labels_emb, cond_drop_prob=1.0, training=training)
    return cfg_combine(cond, null, cond_scale, rescaled_phi, remove_parallel_component,
                       keep_parallel_frac), null
