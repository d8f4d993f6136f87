"""Training-mode vanilla (GroupNorm) UNet: forward and backward of ``Unet.forward``
(CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/models/unet.py:329-378, ``V/`` below) as ``torch.autograd.Function`` nodes over
the CUDA library, so that the reference's step (``loss = diffusion(...); loss.backward()``, V/trainer.py) runs unchanged.

=====================  ==================================================  ==========================================
node                   forward                                             backward
=====================  ==================================================  ==========================================
``GroupNormActFn``     ccdm_channel_stats -> ccdm_groupnorm_coef ->        ccdm_norm_bwd_stats -> ccdm_groupnorm_bwd_coef
                       ccdm_affine_act (per concatenated source)           -> ccdm_norm_bwd_apply (dx, dgamma, dbeta,
                                                                           dscale, dshift)
``train.ConvFn``       ccdm_tapgemm (bias [+ residual]); kinds 3x3, 1x1,   tap-GEMM data gradient, ccdm_conv_wgrad,
                       down3x3s2, up2x3x3                                  ccdm_colsum_bf16
``AttnTokensFn``       ccdm_attention_tokens (per-head q|k|v split)        ccdm_attention_tokens_bwd
=====================  ==================================================  ==========================================

The conditioning path (timestep features, ``time_mlp``, ``classes_emb`` with BatchNorm1d batch statistics, the tc_mlp
Linears: [B, <= 1024] matrices) runs as PyTorch library calls under autograd, as in :mod:`ccdm_b200.train`.  The 3-channel
stem and head use zero-padded weights (64 input columns / 8 output rows) built under autograd, so their weight gradients are
slices of the padded ones.  The network output of this path is the bf16 result of the last tap-GEMM cast to fp32.

STATUS: not yet run on a GPU (tests/test_gpu_vanilla.py, opt-in).  On CPU: the kernels run from their own source (host build,
tests/hostsim) match autograd; this module, unchanged, over those kernels and C-ABI level stand-ins for the tcgen05 entry
points reproduces the reference's own loss and gradients (tests/test_vanilla_train_emulated.py, tests/hostpath.py).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function

from . import _lib as L
from . import backward as K
from .train import ConvFn, _c, _grad_slot
from .vanilla_unet import ACT_AFFINE_NONE, ACT_AFFINE_SILU, AttentionBlock, Downsample, ResidualBlock, Upsample


# ------------------------------------------------------------------------------------------------- eager helpers

def groupnorm_act_forward(srcs: List[torch.Tensor], groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
                          ss: Optional[torch.Tensor], act: int):
    """act(GroupNorm(cat(srcs)) [*(1+scale)+shift]) per source.  ``ss`` fp32 [B, 2*Ctot] = scale | shift.
    Returns (outs, sums [B,2,Ctot], coef [B,2*Ctot]); the last two are what the backward needs besides the inputs."""
    lib, st = L.lib(), K._stream()
    b, h, w, _ = srcs[0].shape
    ctot = sum(s.shape[3] for s in srcs)
    dev = srcs[0].device
    sums = torch.empty(b, 2, ctot, dtype=torch.float32, device=dev)
    off = 0
    for i, s in enumerate(srcs):
        L.check(lib.ccdm_channel_stats(s.data_ptr(), b, h * w, s.shape[3], sums.data_ptr(), ctot, off, int(i == 0), st),
                "channel_stats")
        off += s.shape[3]
    coef = torch.empty(b, 2 * ctot, dtype=torch.float32, device=dev)
    L.check(lib.ccdm_groupnorm_coef(sums.data_ptr(), b, ctot, groups, h * w, eps, gamma.data_ptr(), beta.data_ptr(),
                                    L.ptr(ss), ss.shape[1] if ss is not None else 0, 0, srcs[0].shape[3], coef.data_ptr(), st),
            "groupnorm_coef")
    outs, off = [], 0
    for s in srcs:
        c = s.shape[3]
        o = torch.empty_like(s)
        L.check(lib.ccdm_affine_act(s.data_ptr(), o.data_ptr(), b * h * w, c, h * w, coef.data_ptr(), 2 * ctot, off, act, st),
                "affine_act")
        outs.append(o)
        off += 2 * c
    return outs, sums, coef


def groupnorm_act_backward(dys: List[torch.Tensor], srcs: List[torch.Tensor], sums, coef, groups: int, eps: float,
                           gamma, beta, ss, act: int, dgamma: torch.Tensor, dbeta: torch.Tensor, want_dx: bool = True):
    """Returns (dxs, d_ss or None); ``dgamma`` / ``dbeta`` (fp32 [Ctot]) receive ``+=``."""
    lib, st = L.lib(), K._stream()
    b, h, w, _ = srcs[0].shape
    ctot = sum(s.shape[3] for s in srcs)
    dev = srcs[0].device
    bsums = torch.empty(b, 2, ctot, dtype=torch.float32, device=dev)
    off = coff = 0
    for i, (s, dy) in enumerate(zip(srcs, dys)):
        c = s.shape[3]
        L.check(lib.ccdm_norm_bwd_stats(dy.data_ptr(), s.data_ptr(), b, h * w, c, coef.data_ptr(), 2 * ctot, coff, act,
                                        bsums.data_ptr(), ctot, off, int(i == 0), st), "norm_bwd_stats")
        off, coff = off + c, coff + 2 * c
    bcoef = torch.empty(b, 3 * ctot, dtype=torch.float32, device=dev)
    d_ss = torch.empty(b, 2 * ctot, dtype=torch.float32, device=dev) if ss is not None else None
    L.check(lib.ccdm_groupnorm_bwd_coef(sums.data_ptr(), bsums.data_ptr(), b, ctot, groups, h * w, eps, gamma.data_ptr(),
                                        beta.data_ptr(), L.ptr(ss), ss.shape[1] if ss is not None else 0, 0, srcs[0].shape[3],
                                        bcoef.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), L.ptr(d_ss), st),
            "groupnorm_bwd_coef")
    dxs: List[Optional[torch.Tensor]] = []
    coff = boff = 0
    for s, dy in zip(srcs, dys):
        c = s.shape[3]
        dx = None
        if want_dx:
            dx = torch.empty_like(s)
            L.check(lib.ccdm_norm_bwd_apply(dy.data_ptr(), s.data_ptr(), dx.data_ptr(), b * h * w, c, h * w, coef.data_ptr(),
                                            2 * ctot, coff, bcoef.data_ptr(), 3 * ctot, boff, act, st), "norm_bwd_apply")
        dxs.append(dx)
        coff, boff = coff + 2 * c, boff + 3 * c
    return dxs, d_ss


# ------------------------------------------------------------------------------------------------- nodes

class GroupNormActFn(Function):
    """act(GroupNorm(cat(srcs)) [*(1+scale)+shift]) (V:99,111,146,160,323): one output per concatenated source."""

    @staticmethod
    def forward(ctx, groups, eps, act, gamma, beta, ss, *srcs):
        srcs = [_c(s) for s in srcs]
        ss = _c(ss.float()) if ss is not None else None
        outs, sums, coef = groupnorm_act_forward(srcs, groups, eps, gamma, beta, ss, act)
        ctx.groups, ctx.eps, ctx.act, ctx.has_ss, ctx.n = groups, eps, act, ss is not None, len(srcs)
        ctx.params = (gamma, beta)
        ctx.save_for_backward(gamma, beta, sums, coef, *([ss] if ss is not None else []), *srcs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        saved = list(ctx.saved_tensors)
        gamma, beta, sums, coef = saved[:4]
        ss = saved[4] if ctx.has_ss else None
        srcs = saved[5:] if ctx.has_ss else saved[4:]
        dys = [_c(d) if d is not None else torch.zeros_like(s) for d, s in zip(douts, srcs)]
        gparam, bparam = ctx.params
        gslot, bslot = _grad_slot(gparam), _grad_slot(bparam)
        dgamma = gslot if gslot is not None else torch.zeros_like(gamma)
        dbeta = bslot if bslot is not None else torch.zeros_like(beta)
        dxs, d_ss = groupnorm_act_backward(dys, srcs, sums, coef, ctx.groups, ctx.eps, gamma, beta, ss, ctx.act, dgamma, dbeta,
                                           want_dx=any(ctx.needs_input_grad[6:]))
        return (None, None, None, None if gslot is not None else dgamma, None if bslot is not None else dbeta, d_ss, *dxs)


class AttnTokensFn(Function):
    """AttentionBlock core (V:166-173) over the raw qkv [B,h,w,3C] with the reference's per-head q|k|v split."""

    @staticmethod
    def forward(ctx, qkv, heads):
        qkv = _c(qkv)
        b, h, w, c3 = qkv.shape
        hid = c3 // 3
        dh = hid // heads
        out = torch.empty(b, h, w, hid, dtype=torch.bfloat16, device=qkv.device)
        ctx.heads, ctx.dh = heads, dh
        L.check(L.lib().ccdm_attention_tokens(qkv.data_ptr(), out.data_ptr(), b, h * w, heads, dh, 1.0 / math.sqrt(dh), 1,
                                              K._stream()), "attention_tokens")
        ctx.save_for_backward(qkv)
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        b, h, w, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        L.check(L.lib().ccdm_attention_tokens_bwd(qkv.data_ptr(), _c(dout).data_ptr(), dqkv.data_ptr(), b, h * w, ctx.heads,
                                                  ctx.dh, 1.0 / math.sqrt(ctx.dh), 1, K._stream()), "attention_tokens_bwd")
        return dqkv, None


# ------------------------------------------------------------------------------------------------- composition

def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """V:40-57 -- cos | sin, exponent divided by half."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.reshape(-1)[:, None].float() * freqs[None]
    return torch.cat([args.cos(), args.sin()], dim=-1)


def _gn(gn: nn.GroupNorm, srcs, ss, act):
    return GroupNormActFn.apply(gn.num_groups, gn.eps, act, gn.weight, gn.bias, ss, *srcs)


def _resblock(mod: ResidualBlock, srcs, t_act, tc_act):
    """ResidualBlock.forward (V:124-151).  ``t_act`` = SiLU(t_emb), ``tc_act`` = SiLU(cat(t_emb, c_emb))."""
    lin = mod.tc_mlp[1]
    ss = F.linear(tc_act if mod.cond_channels > 0 else t_act, lin.weight, lin.bias)       # [B, 2*Cout]: scale | shift
    a1 = _gn(mod.conv1[0], srcs, None, ACT_AFFINE_SILU)
    h1 = ConvFn.apply("3x3", mod.conv1[2].weight, mod.conv1[2].bias, None, *a1)
    (a2,) = _gn(mod.conv2[0], [h1], ss, ACT_AFFINE_SILU)
    if isinstance(mod.shortcut, nn.Conv2d):
        res = ConvFn.apply("1x1", mod.shortcut.weight, mod.shortcut.bias, None, *srcs)
    else:
        res = srcs[0]
    return ConvFn.apply("3x3", mod.conv2[3].weight, mod.conv2[3].bias, res, a2)


def _attention(mod: AttentionBlock, x):
    """AttentionBlock.forward (V:165-175)."""
    (xn,) = _gn(mod.norm, [x], None, ACT_AFFINE_NONE)
    qkv = ConvFn.apply("1x1", mod.qkv.weight, None, None, xn)
    o = AttnTokensFn.apply(qkv, mod.num_heads)
    return ConvFn.apply("1x1", mod.proj.weight, mod.proj.bias, x, o)


def _run_layers(seq, srcs, t_act, tc_act):
    x = None
    for mod in seq:
        if isinstance(mod, ResidualBlock):
            x = _resblock(mod, srcs if x is None else [x], t_act, tc_act)
        elif isinstance(mod, AttentionBlock):
            x = _attention(mod, x if x is not None else srcs[0])
        elif isinstance(mod, Downsample):
            x = ConvFn.apply("down3x3s2", mod.op.weight, mod.op.bias, None, x if x is not None else srcs[0])
        elif isinstance(mod, Upsample):
            x = ConvFn.apply("up2x3x3", mod.conv.weight, mod.conv.bias, None, x if x is not None else srcs[0])
        else:
            raise TypeError(type(mod))
    return x


def vanilla_train_forward(net, x: torch.Tensor, t: torch.Tensor, classes: torch.Tensor,
                          keep_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Unet.forward (V:329-378) in training mode with an autograd graph; returns the fp32 NCHW prediction.

    ``keep_mask`` is the Bernoulli mask ``VanillaUnet.forward`` drew (None when cond_drop_prob == 0)."""
    b = x.shape[0]
    # conditioning (V:341-359): tiny fp32 matrices, PyTorch library calls under autograd
    t_emb = net.time_mlp(timestep_embedding(t, net.model_channels))
    c = net.classes_emb(classes.float().reshape(b, -1))
    if keep_mask is not None:
        c = torch.where(keep_mask[:, None], c, net.null_classes_emb[None, :].expand(b, -1).to(c.dtype))
    t_act = F.silu(t_emb)
    tc_act = F.silu(torch.cat((t_emb, c), dim=1))

    # stem (V:266): NCHW fp32 -> 64-channel NHWC bf16 (layout glue), weight columns >= in_channels are zeros
    stem = net.down_blocks[0][0]
    cin = net.in_channels
    x_pad = torch.zeros(b, x.shape[2], x.shape[3], 64, dtype=torch.bfloat16, device=x.device)
    x_pad[..., :cin] = x.permute(0, 2, 3, 1)
    h = ConvFn.apply("3x3", F.pad(stem.weight, (0, 0, 0, 0, 0, 64 - cin)), stem.bias, None, x_pad)
    hs = [h]
    for i in range(1, len(net.down_blocks)):
        h = _run_layers(net.down_blocks[i], [h], t_act, tc_act)
        hs.append(h)
    h = _run_layers(net.middle_block, [h], t_act, tc_act)
    for seq in net.up_blocks:
        h = _run_layers(seq, [h, hs.pop()], t_act, tc_act)
    # head (V:322-326,377): GroupNorm -> SiLU -> 3x3 conv; output rows padded to 8 (the GEMM's channel granularity)
    (a,) = _gn(net.out[0], [h], None, ACT_AFFINE_SILU)
    conv = net.out[2]
    co = conv.weight.shape[0]
    y = ConvFn.apply("3x3", F.pad(conv.weight, (0, 0, 0, 0, 0, 0, 0, 8 - co)), F.pad(conv.bias, (0, 8 - co)), None, a)
    return y[..., :co].permute(0, 3, 1, 2).float()
