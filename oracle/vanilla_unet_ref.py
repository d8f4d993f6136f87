"""Functional fp32 restatement of the *vanilla* CCDM UNet (GroupNorm + SiLU, ADM-style) -- test oracle.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Reference: ``CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/models/unet.py`` (``V/`` below).
The network is evaluated straight from a ``state_dict`` with the reference's keys and shapes; every function cites
the reference lines it restates.  Pinned by ``tests/golden/vanilla_unet.pt`` (outputs of the reference's own module,
``tests/golden/make_golden_vanilla.py``); it holds no code from the reference.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass(frozen=True)
class VanillaSpec:
    """Constructor arguments of V/models/unet.py:205-221 (script values: V/scripts/run_train_ccdm.sh)."""
    model_channels: int = 64
    channel_mult: Tuple[int, ...] = (1, 2, 4, 8)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (16, 32)
    num_heads: int = 4
    num_groups: int = 8
    in_channels: int = 3
    embed_input_dim: int = 128
    out_channels: Optional[int] = None

    @property
    def emb_dim(self) -> int:                       # time_embed_dim == cond_embed_dim, V/models/unet.py:242,250
        return self.model_channels * 4

    @property
    def out_ch(self) -> int:
        return self.out_channels if self.out_channels is not None else self.in_channels


# --------------------------------------------------------------------------- structure

def layout(spec: VanillaSpec):
    """Replays the constructor loops (V/models/unet.py:265-321) -> (down, middle, up) lists of layer descriptors.

    Each entry is a list of ("res", cin, cout, has_cond) | ("attn", c) | ("down", c) | ("up", c) | ("stem", cin, cout)."""
    mc = spec.model_channels
    down = [[("stem", spec.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(spec.channel_mult):
        for _ in range(spec.num_res_blocks):
            layers = [("res", ch, mult * mc, True)]
            ch = mult * mc
            if ds in spec.attention_resolutions:
                layers.append(("attn", ch))
            down.append(layers)
            chans.append(ch)
        if level != len(spec.channel_mult) - 1:
            down.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch, False), ("attn", ch), ("res", ch, ch, False)]
    up = []
    for level, mult in list(enumerate(spec.channel_mult))[::-1]:
        for i in range(spec.num_res_blocks + 1):
            layers = [("res", ch + chans.pop(), mc * mult, True)]
            ch = mc * mult
            if ds in spec.attention_resolutions:
                layers.append(("attn", ch))
            if level and i == spec.num_res_blocks:
                layers.append(("up", ch))
                ds //= 2
            up.append(layers)
    return down, middle, up


def _layer_shapes(name: str, layer, emb: int) -> Dict[str, tuple]:
    kind = layer[0]
    if kind == "stem":
        return {f"{name}.weight": (layer[2], layer[1], 3, 3), f"{name}.bias": (layer[2],)}
    if kind == "res":                               # V/models/unet.py:93-121
        _, cin, cout, has_cond = layer
        s = {f"{name}.conv1.0.weight": (cin,), f"{name}.conv1.0.bias": (cin,),
             f"{name}.conv1.2.weight": (cout, cin, 3, 3), f"{name}.conv1.2.bias": (cout,),
             f"{name}.tc_mlp.1.weight": (2 * cout, emb * (2 if has_cond else 1)), f"{name}.tc_mlp.1.bias": (2 * cout,),
             f"{name}.conv2.0.weight": (cout,), f"{name}.conv2.0.bias": (cout,),
             f"{name}.conv2.3.weight": (cout, cout, 3, 3), f"{name}.conv2.3.bias": (cout,)}
        if cin != cout:
            s[f"{name}.shortcut.weight"] = (cout, cin, 1, 1)
            s[f"{name}.shortcut.bias"] = (cout,)
        return s
    if kind == "attn":                              # V/models/unet.py:155-163
        c = layer[1]
        return {f"{name}.norm.weight": (c,), f"{name}.norm.bias": (c,), f"{name}.qkv.weight": (3 * c, c, 1, 1),
                f"{name}.proj.weight": (c, c, 1, 1), f"{name}.proj.bias": (c,)}
    if kind == "down":                              # V/models/unet.py:193-198
        c = layer[1]
        return {f"{name}.op.weight": (c, c, 3, 3), f"{name}.op.bias": (c,)}
    if kind == "up":                                # V/models/unet.py:178-183
        c = layer[1]
        return {f"{name}.conv.weight": (c, c, 3, 3), f"{name}.conv.bias": (c,)}
    raise ValueError(kind)


def state_dict_shapes(spec: VanillaSpec) -> Dict[str, tuple]:
    """Key -> shape map of the reference ``Unet(...).state_dict()`` in registration order."""
    mc, e = spec.model_channels, spec.emb_dim
    s: Dict[str, tuple] = {
        "null_classes_emb": (e,),
        "time_mlp.0.weight": (e, mc), "time_mlp.0.bias": (e,),
        "time_mlp.2.weight": (e, e), "time_mlp.2.bias": (e,),
        "classes_emb.0.weight": (e, spec.embed_input_dim), "classes_emb.0.bias": (e,),
        "classes_emb.1.weight": (e,), "classes_emb.1.bias": (e,), "classes_emb.1.running_mean": (e,),
        "classes_emb.1.running_var": (e,), "classes_emb.1.num_batches_tracked": (),
    }
    down, middle, up = layout(spec)
    for i, layers in enumerate(down):
        for j, layer in enumerate(layers):
            s.update(_layer_shapes(f"down_blocks.{i}.{j}", layer, e))
    for j, layer in enumerate(middle):
        s.update(_layer_shapes(f"middle_block.{j}", layer, e))
    for i, layers in enumerate(up):
        for j, layer in enumerate(layers):
            s.update(_layer_shapes(f"up_blocks.{i}.{j}", layer, e))
    s.update({"out.0.weight": (mc,), "out.0.bias": (mc,),
              "out.2.weight": (spec.out_ch, mc, 3, 3), "out.2.bias": (spec.out_ch,)})
    return s


def make_state_dict(spec: VanillaSpec, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic synthetic weights that depend only on (key, shape, seed) -- not an init scheme of the reference."""
    out = {}
    for key, shape in state_dict_shapes(spec).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 104729 * seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.tensor(3, dtype=torch.long)
        elif key.endswith("running_var"):
            out[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            out[key] = 0.2 * torch.randn(shape, generator=g)
        elif key == "null_classes_emb":
            out[key] = -torch.randn(shape, generator=g).abs()
        elif key.endswith("weight") and len(shape) == 1:            # GroupNorm / BatchNorm gains
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            out[key] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for n in shape[1:]:
                fan_in *= n
            out[key] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return out


# --------------------------------------------------------------------------- building blocks

def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """V/models/unet.py:40-57 -- cos | sin, exponent divided by half (not half - 1)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.reshape(-1)[:, None].float() * freqs[None]
    emb = torch.cat([args.cos(), args.sin()], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, name: str, x: Tensor, groups: int) -> Tensor:
    return F.group_norm(x, groups, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-5)     # nn.GroupNorm default eps


def _resblock(sd, name: str, x: Tensor, t_emb: Tensor, c_emb: Optional[Tensor], groups: int) -> Tensor:
    """ResidualBlock.forward with use_scale_shift_norm=True, V/models/unet.py:124-151 (dropout p = 0)."""
    h = F.conv2d(F.silu(_gn(sd, f"{name}.conv1.0", x, groups)), sd[f"{name}.conv1.2.weight"], sd[f"{name}.conv1.2.bias"],
                 padding=1)
    tc = t_emb if c_emb is None else torch.cat((t_emb, c_emb), dim=1)
    tc = F.linear(F.silu(tc), sd[f"{name}.tc_mlp.1.weight"], sd[f"{name}.tc_mlp.1.bias"])[:, :, None, None]
    scale, shift = tc.chunk(2, dim=1)
    h = _gn(sd, f"{name}.conv2.0", h, groups) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[f"{name}.conv2.3.weight"], sd[f"{name}.conv2.3.bias"], padding=1)
    if f"{name}.shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.shortcut.weight"], sd[f"{name}.shortcut.bias"])
    return h + x


def _attention(sd, name: str, x: Tensor, heads: int, groups: int) -> Tensor:
    """AttentionBlock.forward, V/models/unet.py:165-175: the 3C qkv channels are split per head FIRST
    (reshape to [B*heads, 3C/heads, T]) and only then chunked into q | k | v."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(_gn(sd, f"{name}.norm", x, groups), sd[f"{name}.qkv.weight"])
    q, k, v = qkv.reshape(b * heads, -1, hh * ww).chunk(3, dim=1)
    scale = 1.0 / math.sqrt(math.sqrt(c // heads))
    att = torch.einsum("bct,bcs->bts", q * scale, k * scale).softmax(dim=-1)
    h = torch.einsum("bts,bcs->bct", att, v).reshape(b, -1, hh, ww)
    return F.conv2d(h, sd[f"{name}.proj.weight"], sd[f"{name}.proj.bias"]) + x


def _run_layers(sd, prefix: str, layers, h, t_emb, c_emb, spec: VanillaSpec):
    """TimestepEmbedSequential.forward, V/models/unet.py:78-84."""
    for j, layer in enumerate(layers):
        name = f"{prefix}.{j}"
        kind = layer[0]
        if kind == "stem":
            h = F.conv2d(h, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=1)
        elif kind == "res":
            h = _resblock(sd, name, h, t_emb, c_emb if layer[3] else None, spec.num_groups)
        elif kind == "attn":
            h = _attention(sd, name, h, spec.num_heads, spec.num_groups)
        elif kind == "down":
            h = F.conv2d(h, sd[f"{name}.op.weight"], sd[f"{name}.op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"], padding=1)
    return h


def vanilla_unet_forward(sd: Dict[str, Tensor], spec: VanillaSpec, x: Tensor, t: Tensor, classes: Tensor,
                         keep_mask: Optional[Tensor] = None, training: bool = False) -> Tensor:
    """Unet.forward, V/models/unet.py:329-378.  ``keep_mask`` (bool [B], True = keep the label embedding) replaces the
    Bernoulli draw of :352; None keeps every row (cond_drop_prob = 0)."""
    t_emb = timestep_embedding(t, spec.model_channels)
    t_emb = F.linear(F.silu(F.linear(t_emb, sd["time_mlp.0.weight"], sd["time_mlp.0.bias"])),
                     sd["time_mlp.2.weight"], sd["time_mlp.2.bias"])
    c = F.linear(classes, sd["classes_emb.0.weight"], sd["classes_emb.0.bias"])
    if training:
        mean, var = c.mean(0), c.var(0, unbiased=False)
    else:
        mean, var = sd["classes_emb.1.running_mean"], sd["classes_emb.1.running_var"]
    c = F.relu((c - mean) / torch.sqrt(var + 1e-5) * sd["classes_emb.1.weight"] + sd["classes_emb.1.bias"])
    if keep_mask is not None:
        c = torch.where(keep_mask[:, None], c, sd["null_classes_emb"][None].expand_as(c))
    down, middle, up = layout(spec)
    hs: List[Tensor] = []
    h = x
    for i, layers in enumerate(down):
        h = _run_layers(sd, f"down_blocks.{i}", layers, h, t_emb, c, spec)
        hs.append(h)
    h = _run_layers(sd, "middle_block", middle, h, t_emb, None, spec)
    for i, layers in enumerate(up):
        h = _run_layers(sd, f"up_blocks.{i}", layers, torch.cat([h, hs.pop()], dim=1), t_emb, c, spec)
    h = F.silu(_gn(sd, "out.0", h, spec.num_groups))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def vanilla_forward_with_cond_scale(sd, spec, x, t, classes, cond_scale: float = 3.0, rescaled_phi: float = 0.0):
    """V/diffusion.py:34-56 -- plain classifier-free guidance with the optional std rescale."""
    b = x.shape[0]
    logits = vanilla_unet_forward(sd, spec, x, t, classes, torch.ones(b, dtype=torch.bool, device=x.device))
    if cond_scale == 1:
        return logits
    null = vanilla_unet_forward(sd, spec, x, t, classes, torch.zeros(b, dtype=torch.bool, device=x.device))
    scaled = null + (logits - null) * cond_scale
    if rescaled_phi == 0.0:
        return scaled
    dims = tuple(range(1, scaled.ndim))
    rescaled = scaled * (logits.std(dim=dims, keepdim=True) / scaled.std(dim=dims, keepdim=True))
    return rescaled * rescaled_phi + scaled * (1.0 - rescaled_phi)
