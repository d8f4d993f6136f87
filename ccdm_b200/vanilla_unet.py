"""The *vanilla* CCDM UNet (ADM-style: GroupNorm -> SiLU -> conv, scale-shift conditioning, QKV attention) on the CUDA path.

Drop-in for ``CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/models/unet.py:203-378`` (``V/`` below; SURVEY.md section 8f
rank 4): same constructor keywords, ``forward(x, timesteps, classes, cond_drop_prob, return_null_indx)`` signature and
``state_dict`` keys / shapes / registration order.  The sub-modules only own parameters; the arithmetic is a flat
program of C-ABI calls (same machinery as :mod:`ccdm_b200.engine`):

  ResidualBlock (V:124-151)   x --channel_stats--> groupnorm_coef --> affine_act(SiLU) --> tap-GEMM 3x3 (+bias)
                              --channel_stats--> groupnorm_coef(with (1+scale), shift of the tc_mlp row-GEMM)
                              --> affine_act(SiLU) --> tap-GEMM 3x3 (+bias, + shortcut residual)
                              shortcut 1x1 and every conv over ``torch.cat([h, skip])`` are dual-source K loops; the
                              GroupNorm of a concatenation is computed per source (groups may straddle the two)
  AttentionBlock (V:165-175)  GroupNorm --> qkv 1x1 --> ccdm_attention_tokens (per-head q|k|v split) --> proj 1x1 + x
  Downsample (V:193-198)      3x3 / stride 2 as taps over four strided parity views ("down3x3s2")
  Upsample (V:178-190)        nearest-2x folded into four output-parity 2x2 convs ("up2x3x3")
  embeddings (V:242-262)      ccdm_time_features_adm, ccdm_linear_small (SiLU / BatchNorm1d + ReLU), ccdm_select_null;
                              all tc_mlp Linears as ONE row-GEMM (middle blocks see only the time half: zero K blocks)

Boundary glue in PyTorch (layout only): NCHW fp32 -> the first 3 of 64 bf16 NHWC channels for the 3x3 stem, and the
8-channel fp32 NHWC output of the last conv -> NCHW.  There is no CPU / PyTorch fallback for the arithmetic.

The program exposes the same ``x_in / t_in / emb_in / keep / out / run(stream)`` surface as the unified UNet's, so
:class:`ccdm_b200.VanillaGaussianDiffusion` drives it from the same CUDA-graph sampling loop.

STATUS: not yet run on a GPU (tests/test_gpu_vanilla.py, opt-in until the first device run).  On CPU the unchanged product code
over host builds of its CUDA-core kernels and a C-ABI level stand-in for ccdm_tapgemm reproduces the reference's own outputs
(0.8-1.0 %, the bf16 storage level) and samples (40-49 dB): tests/test_vanilla_emulated.py, tests/test_sampling_hostpath.py.
Training (``loss.backward()``) runs through ccdm_b200/vanilla_train.py.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib as L
from .engine import TapGemmRec, UnetProgram, WeightStore, nhwc_view
from .plan import KB, can_reuse_rows, n_tiling, plan_conv, tile_box

ACT_AFFINE_NONE, ACT_AFFINE_SILU = 0, 2          # ccdm_affine_act codes


def _holder_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} only holds parameters; run the whole ccdm_b200.VanillaUnet instead")


def prob_mask_like(shape, prob, device):          # V:31-37
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def _current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("ccdm_b200.VanillaUnet runs on sm_100a only (there is no CPU fallback)")


class TimestepEmbedSequential(nn.Sequential):     # V:72-84
    forward = _holder_forward


class ResidualBlock(nn.Module):                   # V:93-121 (same registration order)
    def __init__(self, in_channels, out_channels, time_channels, cond_channels, dropout, use_scale_shift_norm=False,
                 num_groups=32):
        super().__init__()
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is never used by the reference (V/main.py:336)")
        self.in_channels, self.out_channels, self.cond_channels = in_channels, out_channels, cond_channels
        self.conv1 = nn.Sequential(nn.GroupNorm(num_groups, in_channels), nn.SiLU(),
                                   nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))
        self.tc_mlp = nn.Sequential(nn.SiLU(), nn.Linear(int(time_channels) + int(cond_channels), 2 * out_channels))
        self.conv2 = nn.Sequential(nn.GroupNorm(num_groups, out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                   nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1))
        self.shortcut = (nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels
                         else nn.Identity())

    forward = _holder_forward


class AttentionBlock(nn.Module):                  # V:155-163
    def __init__(self, channels, num_heads=1, num_groups=32):
        super().__init__()
        assert channels % num_heads == 0
        self.num_heads = num_heads
        self.norm = nn.GroupNorm(num_groups, channels)
        self.qkv = nn.Conv2d(channels, channels * 3, kernel_size=1, bias=False)
        self.proj = nn.Conv2d(channels, channels, kernel_size=1)

    forward = _holder_forward


class Upsample(nn.Module):                        # V:178-183
    def __init__(self, channels, use_conv):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("conv_resample=False is never used by the reference (V/main.py:334)")
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, padding=1)

    forward = _holder_forward


class Downsample(nn.Module):                      # V:193-198
    def __init__(self, channels, use_conv):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("conv_resample=False builds nn.AvgPool2d(stride=2), which raises in the reference too")
        self.op = nn.Conv2d(channels, channels, kernel_size=3, stride=2, padding=1)

    forward = _holder_forward


class VanillaUnet(nn.Module):
    """``Unet`` of V/models/unet.py:203-378."""

    def __init__(self, embed_input_dim=128, cond_drop_prob=0.5, in_channels=3, model_channels=128, out_channels=None,
                 num_res_blocks=2, attention_resolutions=(8, 16), dropout=0, channel_mult=(1, 2, 4, 8),
                 conv_resample=True, num_heads=4, use_scale_shift_norm=True, learned_variance=False, num_groups=32):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is never used by the reference (V/main.py:332)")
        if channel_mult[0] != 1:
            raise ValueError("the output head is built for model_channels inputs (V:322-326): channel_mult[0] must be 1")
        self.embed_input_dim, self.cond_drop_prob = embed_input_dim, cond_drop_prob
        self.in_channels, self.model_channels = in_channels, model_channels
        self.out_channels = out_channels if out_channels is not None else in_channels * (2 if learned_variance else 1)
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, tuple(attention_resolutions)
        self.dropout, self.channel_mult, self.conv_resample = dropout, tuple(channel_mult), conv_resample
        self.num_heads, self.num_groups = num_heads, num_groups
        # what ccdm_b200.GaussianDiffusion reads from a denoiser: output width, time-embedding kind, guidance flavour
        self.out_dim = self.out_channels
        self.random_or_learned_sinusoidal_cond = False
        self.cfg_remove_parallel = False           # plain CFG (V/diffusion.py:34-56): no orthogonal-update projection
        mc = model_channels
        e = mc * 4
        self.time_mlp = nn.Sequential(nn.Linear(mc, e), nn.SiLU(), nn.Linear(e, e))
        self.classes_emb = nn.Sequential(nn.Linear(embed_input_dim, e), nn.BatchNorm1d(e), nn.ReLU())
        self.null_classes_emb = nn.Parameter(-1 * torch.abs(torch.randn(e)), requires_grad=False)

        res = lambda i, o, c: ResidualBlock(i, o, e, c, dropout, use_scale_shift_norm=use_scale_shift_norm,
                                            num_groups=num_groups)
        att = lambda c: AttentionBlock(c, num_heads=num_heads, num_groups=num_groups)
        self.down_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, mc, kernel_size=3, padding=1))])
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * mc, e)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(att(ch))
                self.down_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.down_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch, 0), att(ch), res(ch, ch, 0))
        self.up_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + chans.pop(), mc * mult, e)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(att(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample))
                    ds //= 2
                self.up_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(num_groups, ch), nn.SiLU(),
                                 nn.Conv2d(mc, self.out_channels, kernel_size=3, padding=1))
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def __deepcopy__(self, memo):                 # EMA deep-copies the model: the engine (device buffers) is not copied
        import copy
        eng, self._engine = self._engine, None
        try:
            new = type(self).__new__(type(self))
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        return st

    def engine(self) -> "VanillaEngine":
        dev = self.null_classes_emb.device
        if dev.type != "cuda":
            raise RuntimeError("ccdm_b200.VanillaUnet runs on sm_100a only: move the model to a CUDA device "
                               "(there is no CPU fallback)")
        if self._engine is None or self._engine.device != dev:
            self._engine = VanillaEngine(self)
        return self._engine

    # ------------------------------------------------------------------ reference API
    def forward(self, x, timesteps, classes, cond_drop_prob=None, return_null_indx=False):
        """V:329-378.  The Bernoulli label-drop mask is drawn exactly as ``prob_mask_like`` does (V:31-37)."""
        p = self.cond_drop_prob if cond_drop_prob is None else cond_drop_prob
        b = x.shape[0]
        mask = None
        if p > 0:
            mask = prob_mask_like((b,), 1 - p, x.device)
            self.keep_mask = mask
        if self.training and torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()):
            # training step: autograd graph over the CUDA kernels (ccdm_b200/vanilla_train.py); BatchNorm1d uses batch statistics
            _require_cuda(x)
            from .vanilla_train import vanilla_train_forward
            out = vanilla_train_forward(self, x, timesteps, classes, mask)
        else:
            out = self.engine().forward(x, timesteps, classes, mask)
        if return_null_indx:
            return out, torch.where(self.keep_mask == False)[0]      # noqa: E712  (V:375)
        return out

    def forward_with_cond_scale(self, x, timesteps, classes, cond_scale=3.0, rescaled_phi=0.0):
        """V/diffusion.py:34-56 (a module-level function there): plain classifier-free guidance + optional std rescale."""
        eng = self.engine()
        if cond_scale == 1:
            return eng.forward(x, timesteps, classes, None)
        if self.training:
            cond = eng.forward(x, timesteps, classes, None)
            null = eng.forward(x, timesteps, classes, torch.zeros(x.shape[0], device=x.device, dtype=torch.bool))
        else:
            cond, null = eng.forward_pair(x, timesteps, classes)
        out = torch.empty_like(cond)
        L.check(L.lib().ccdm_cfg_combine(cond.data_ptr(), null.data_ptr(), out.data_ptr(), cond.shape[0], cond[0].numel(),
                                         float(cond_scale), float(rescaled_phi), 0, 0.0,
                                         _current_stream()), "cfg_combine")
        return out


# --------------------------------------------------------------------------------------------- engine

class VanillaEngine:
    def __init__(self, net: VanillaUnet):
        self.net = net
        self.device = net.null_classes_emb.device
        self.weights = WeightStore(self.device)
        self.programs: Dict[tuple, "VanillaProgram"] = {}
        self._ptr_stamp = None

    def program(self, B, x_batch, H, W, training) -> "VanillaProgram":
        """Same signature as UnetEngine.program: ``x_batch`` < B (= B/2) means the cond / null halves share one input."""
        ptrs = tuple(p.data_ptr() for p in self.net.parameters())
        if ptrs != self._ptr_stamp:               # parameters were re-allocated: rebuild everything
            self.programs.clear()
            self.weights = WeightStore(self.device)
            self._ptr_stamp = ptrs
        key = (B, x_batch, H, W, bool(training))
        prog = self.programs.get(key)
        if prog is None:
            with torch.inference_mode(False):
                prog = VanillaProgram(self.net, self.weights, B, x_batch, H, W, training)
            self.programs[key] = prog
        return prog

    def _run(self, prog, x, t, classes, keep_rows):
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, model on {self.device}")
        prog.load_inputs(x, t, classes, keep_rows)
        stream = _current_stream()
        self.weights.refresh(stream)
        prog.run(stream)
        if self.net.training:
            self.net.classes_emb[1].num_batches_tracked += 1
        return prog.out

    def forward(self, x, t, classes, keep_mask):
        B, _, H, W = x.shape
        prog = self.program(B, B, H, W, self.net.training)
        keep = torch.ones(B, dtype=torch.uint8, device=self.device) if keep_mask is None else keep_mask.to(torch.uint8)
        return self._run(prog, x, t, classes, keep).clone()

    def forward_pair(self, x, t, classes):
        """Conditional and unconditional evaluations as one 2B batch (eval mode: BatchNorm1d uses running statistics and
        GroupNorm is per sample, so this equals two forwards).  Returns views of the persistent output buffer."""
        B, _, H, W = x.shape
        prog = self.program(2 * B, B, H, W, False)
        if prog.pair_keep is None:
            with torch.inference_mode(False):
                prog.pair_keep = torch.cat([torch.ones(B, dtype=torch.uint8), torch.zeros(B, dtype=torch.uint8)]).to(self.device)
        out = self._run(prog, x, torch.cat([t.reshape(-1), t.reshape(-1)]), torch.cat([classes, classes]), prog.pair_keep)
        return out[:B], out[B:]


def partial_k_plan(k_total: int, k_valid: int, cout: int):
    """1x1 plan over ``k_total`` input channels whose weight only has the first ``k_valid`` columns: the remaining K
    blocks are packed as zeros (nvalid = 0).  Lets the middle blocks' tc_mlp (time embedding only, V:301-305) share the
    [SiLU(t_emb) | SiLU(c_emb)] row-GEMM of the conditional blocks."""
    plan = plan_conv("1x1", [k_total], cout)
    ps = [(c0, max(0, min(nv, k_valid - c0)), m, z) for (c0, nv, m, z) in plan.psched]
    return dataclasses.replace(plan, psched=ps)


class VanillaProgram(UnetProgram):
    """Flat C-ABI program of one VanillaUnet evaluation at a fixed (batch, resolution, mode)."""

    def __init__(self, net: VanillaUnet, weights: WeightStore, B, x_batch, H, W, training):
        # (UnetProgram.__init__ reads the unified UNet's attributes, so the base Program is initialised directly)
        from .engine import Program
        Program.__init__(self, net.null_classes_emb.device)
        self.net, self.weights = net, weights
        assert B % x_batch == 0
        self.B, self.x_batch, self.H, self.W, self.training = B, x_batch, H, W, training
        self.pair_keep = None
        down = 2 ** (len(net.channel_mult) - 1)
        if H % down or W % down:
            raise ValueError(f"image size {H}x{W} must be a multiple of {down}")
        self._build()
        self.finalize()

    # ------------------------------------------------------------------ boundary
    def load_inputs(self, x, t, classes, keep_rows):
        self.x_in.copy_(x)
        self.t_in.copy_(t.reshape(-1))
        self.emb_in.copy_(classes.reshape(self.B, -1))
        self.keep.copy_(keep_rows)

    def glue_in(self):
        """Layout glue, NCHW fp32 ``x_in`` -> the first channels of the NHWC bf16 stem input (every x_batch slab)."""
        xb, cin = self.x_batch, self.net.in_channels
        src = self.x_in.permute(0, 2, 3, 1)
        for rep in range(self.B // xb):
            self.x_pad[rep * xb:(rep + 1) * xb, ..., :cin].copy_(src)

    def glue_out(self):
        """Layout glue, 8-channel fp32 NHWC output of the last conv -> NCHW fp32 ``out``."""
        self.out.copy_(self.out8[..., : self.net.out_channels].permute(0, 3, 1, 2))

    # ------------------------------------------------------------------ GroupNorm
    def gn_coef(self, name, srcs: List[torch.Tensor], gn: nn.GroupNorm, ss_off: Optional[int] = None):
        """Per-(sample, channel) affine coefficients of GroupNorm over the concatenation of ``srcs``."""
        B, h, w = self.B, srcs[0].shape[1], srcs[0].shape[2]
        ctot = sum(s.shape[3] for s in srcs)
        assert ctot == gn.num_channels and len(srcs) <= 2
        sums = self.buf(name + ".sums", (B, 2, ctot), torch.float32)
        off = 0
        for i, s in enumerate(srcs):
            self.kernel("channel_stats", x=s, B=B, rows=h * w, C=s.shape[3], sums=sums, ld=ctot, c_off=off,
                        zero_first=int(i == 0))
            off += s.shape[3]
        coef = self.buf(name + ".coef", (B, 2 * ctot), torch.float32)
        ss = self.bufs["ss_all"] if ss_off is not None else None
        self.kernel("groupnorm_coef", sums=sums, B=B, Ctot=ctot, groups=gn.num_groups, rows=h * w, eps=gn.eps,
                    gamma=gn.weight, beta=gn.bias, ss=ss, ss_ld=ss.shape[1] if ss is not None else 0,
                    ss_off=ss_off or 0, C0=srcs[0].shape[3], coef=coef)
        return coef

    def gn_apply(self, name, srcs, coef, act):
        outs, off = [], 0
        for i, s in enumerate(srcs):
            _, h, w, c = s.shape
            o = self.act(f"{name}.a{i}", h, w, c)
            self.kernel("affine_act", x=s, out=o, rows=self.B * h * w, C=c, rows_per_sample=h * w, ss=coef,
                        ss_ld=coef.shape[1], ss_off=off, act=act)
            outs.append(o)
            off += 2 * c
        return outs

    # ------------------------------------------------------------------ network pieces
    def v_resblock(self, name, mod: ResidualBlock, srcs, h, w, ss_off):
        cout = mod.out_channels
        a1 = self.gn_apply(name + ".conv1", srcs, self.gn_coef(name + ".conv1.0", srcs, mod.conv1[0]), ACT_AFFINE_SILU)
        h1 = self.act(name + ".h", h, w, cout)
        self.conv(name + ".conv1.2", "3x3", a1, mod.conv1[2], h1)
        a2 = self.gn_apply(name + ".conv2", [h1], self.gn_coef(name + ".conv2.0", [h1], mod.conv2[0], ss_off=ss_off),
                           ACT_AFFINE_SILU)
        if isinstance(mod.shortcut, nn.Conv2d):
            res = self.act(name + ".res", h, w, cout)
            self.conv(name + ".shortcut", "1x1", srcs, mod.shortcut, res)
        else:
            assert len(srcs) == 1
            res = srcs[0]
        out = self.act(name + ".out", h, w, cout)
        self.conv(name + ".conv2.3", "3x3", a2, mod.conv2[3], out, L.EPI_RESID, resid=res)
        return out

    def v_attention(self, name, mod: AttentionBlock, x, h, w):
        C = x.shape[3]
        heads, dh = mod.num_heads, C // mod.num_heads
        xn = self.gn_apply(name + ".norm", [x], self.gn_coef(name + ".norm", [x], mod.norm), ACT_AFFINE_NONE)
        qkv = self.act(name + ".qkv_out", h, w, 3 * C)
        self.conv(name + ".qkv", "1x1", xn, mod.qkv, qkv)
        ao = self.act(name + ".attn", h, w, C)
        # q*s and k*s with s = dh^-1/4 (V:169-170)  ==  logits * dh^-1/2
        self.kernel("attention_tokens", qkv=qkv, out=ao, B=self.B, n=h * w, heads=heads, dim_head=dh,
                    scale=1.0 / math.sqrt(dh), head_major=1)
        out = self.act(name + ".out", h, w, C)
        self.conv(name + ".proj", "1x1", [ao], mod.proj, out, L.EPI_RESID, resid=x)
        return out

    def _run_layers(self, prefix, seq, srcs, h, w, ss_offs):
        """One TimestepEmbedSequential (V:78-84).  Returns (tensor, h, w)."""
        x = None
        for j, mod in enumerate(seq):
            name = f"{prefix}.{j}"
            if isinstance(mod, ResidualBlock):
                x = self.v_resblock(name, mod, srcs if x is None else [x], h, w, ss_offs[name])
            elif isinstance(mod, AttentionBlock):
                x = self.v_attention(name, mod, x if x is not None else srcs[0], h, w)
            elif isinstance(mod, Downsample):
                src = x if x is not None else srcs[0]
                h, w = h // 2, w // 2
                x = self.act(name + ".out", h, w, mod.op.weight.shape[0])
                self.conv(name + ".op", "down3x3s2", [src], mod.op, x)
            elif isinstance(mod, Upsample):
                src = x if x is not None else srcs[0]
                h, w = h * 2, w * 2
                x = self.act(name + ".out", h, w, mod.conv.weight.shape[0])
                self.conv(name + ".conv", "up2x3x3", [src], mod.conv, x)
            else:
                raise TypeError(type(mod))
        return x, h, w

    def _resblocks(self):
        net, out = self.net, []
        for i, seq in enumerate(net.down_blocks):
            out += [(f"down_blocks.{i}.{j}", m) for j, m in enumerate(seq) if isinstance(m, ResidualBlock)]
        out += [(f"middle_block.{j}", m) for j, m in enumerate(net.middle_block) if isinstance(m, ResidualBlock)]
        for i, seq in enumerate(net.up_blocks):
            out += [(f"up_blocks.{i}.{j}", m) for j, m in enumerate(seq) if isinstance(m, ResidualBlock)]
        return out

    # ------------------------------------------------------------------ whole network
    def _build(self):
        net, B, H, W, dev = self.net, self.B, self.H, self.W, self.device
        mc, e = net.model_channels, net.model_channels * 4
        self.x_in = self.buf("x_in", (self.x_batch, net.in_channels, H, W), torch.float32)
        self.out = self.buf("out", (B, net.out_channels, H, W), torch.float32)
        self.x_pad = self.buf("x_pad", (B, H, W, KB), torch.bfloat16)
        self.x_pad.zero_()                              # channels >= in_channels stay zero (and meet zero weights)
        self.t_in = self.buf("t_in", (B,), torch.int64)
        self.emb_in = self.buf("emb_in", (B, net.embed_input_dim), torch.float32)
        self.keep = self.buf("keep", (B,), torch.uint8)
        self.out8 = self.buf("out8", (B, H, W, 8), torch.float32)

        # ---- embeddings (V:341-359)
        tf = self.buf("t_feat", (B, mc), torch.float32)
        self.kernel("time_features_adm", t=self.t_in, B=B, dim=mc, max_period=10000.0, out=tf)
        t1 = self.buf("t_hidden", (B, e), torch.float32)
        self.kernel("linear_small", x=tf, B=B, in_dim=mc, w=net.time_mlp[0].weight, bias=net.time_mlp[0].bias, out_dim=e,
                    act=L.ACT_SILU, y=t1)
        temb = self.buf("t_emb", (B, e), torch.float32)
        self.kernel("linear_small", x=t1, B=B, in_dim=e, w=net.time_mlp[2].weight, bias=net.time_mlp[2].bias, out_dim=e,
                    act=L.ACT_NONE, y=temb)
        cemb = self.buf("c_emb", (B, e), torch.float32)
        self.kernel("linear_small", x=self.emb_in, B=B, in_dim=net.embed_input_dim, w=net.classes_emb[0].weight,
                    bias=net.classes_emb[0].bias, out_dim=e, bn=net.classes_emb[1], bn_train=self.training,
                    act=L.ACT_RELU, y=cemb)
        self.kernel("select_null", c=cemb, keep=self.keep, null_emb=net.null_classes_emb, B=B, dim=e)
        tc = self.buf("tc_silu", (1, 1, B, 2 * e), torch.bfloat16)          # "image" of B pixels in one row
        self.kernel("silu_concat_bf16", t_emb=temb, dt=e, c_emb=cemb, dc=e, B=B, out=tc)

        # ---- every ResidualBlock's tc_mlp Linear (V:102-105,131-141) as one row-GEMM into ss_all[B, sum 2*Cout]
        blocks = self._resblocks()
        ss_offs, tot = {}, 0
        for nm, mod in blocks:
            ss_offs[nm] = tot
            tot += 2 * mod.out_channels
        n_rows = (tot + 127) // 128 * 128
        ss_all = self.buf("ss_all", (B, n_rows), torch.float32)
        plan = plan_conv("1x1", [2 * e], n_rows)
        shared = self.weights.packs["tc_mlp.0"].packed if "tc_mlp.0" in self.weights.packs else torch.zeros(
            n_rows, plan.nkb * KB, dtype=torch.bfloat16, device=dev)
        if "tc_bias" not in self.weights.__dict__:
            self.weights.tc_bias = torch.zeros(n_rows, dtype=torch.float32, device=dev)
        for i, (nm, mod) in enumerate(blocks):
            lin = mod.tc_mlp[1]
            self.weights.add(f"tc_mlp.{i}", lin.weight, partial_k_plan(2 * e, lin.weight.shape[1], 2 * mod.out_channels),
                             2 * mod.out_channels, shared=shared, row_off=ss_offs[nm])
        self._tc_bias_srcs = [(ss_offs[nm], mod.tc_mlp[1].bias) for nm, mod in blocks]
        sched = torch.tensor(plan.sched, dtype=torch.int32, device=dev)
        self.recs.append(TapGemmRec("tc_mlp", plan, [nhwc_view(tc)], B, 1, 1, (128, 1, 1), None, shared, sched, n_rows,
                                    n_rows, 128, L.EPI_BIAS | L.EPI_OUT_F32, ss_all, (n_rows, 0, 0),
                                    bias=self.weights.tc_bias))

        # ---- stem (V:266): 3x3 conv over the 64-channel padded input; only the first in_channels weight columns exist
        stem_conv = net.down_blocks[0][0]
        x = self.act("down_blocks.0.0.out", H, W, mc)
        self._conv_padded_cin("down_blocks.0.0", stem_conv, self.x_pad, net.in_channels, x)

        # ---- down path (V:362-365)
        h, w = H, W
        hs = [x]
        for i in range(1, len(net.down_blocks)):
            x, h, w = self._run_layers(f"down_blocks.{i}", net.down_blocks[i], [x], h, w, ss_offs)
            hs.append(x)
        # ---- middle (V:367)
        x, h, w = self._run_layers("middle_block", net.middle_block, [x], h, w, ss_offs)
        # ---- up path (V:369-371): torch.cat([h, hs.pop()]) == two sources
        for i, seq in enumerate(net.up_blocks):
            x, h, w = self._run_layers(f"up_blocks.{i}", seq, [x, hs.pop()], h, w, ss_offs)
        assert not hs and (h, w) == (H, W)

        # ---- head (V:322-326,377): GroupNorm -> SiLU -> 3x3 conv to out_channels; rows padded to 8, fp32 NHWC
        a = self.gn_apply("out", [x], self.gn_coef("out.0", [x], net.out[0]), ACT_AFFINE_SILU)
        self._conv_head("out.2", net.out[2], a[0])

    def _conv_padded_cin(self, name, conv_mod, src64, cin_valid, out):
        """3x3 conv whose input tensor is padded to 64 channels while the weight has ``cin_valid`` columns."""
        cout = conv_mod.weight.shape[0]
        gh, gw = out.shape[1], out.shape[2]
        tile = tile_box(gw, gh, square=True)
        plan = plan_conv("3x3", [KB], cout, reuse_rows=can_reuse_rows(tile))
        plan = dataclasses.replace(plan, psched=[(c0, min(nv, cin_valid), m, z) for (c0, nv, m, z) in plan.psched])
        n_rows, n_tile = n_tiling(cout, False)
        pack = self.weights.add(f"{name}/R{plan.R}", conv_mod.weight, plan, n_rows)
        self.recs.append(TapGemmRec(name, plan, [nhwc_view(src64)], gw, gh, self.B, tile, pack, pack.packed, pack.sched,
                                    n_rows, cout, n_tile, L.EPI_BIAS, out, (cout, gw * cout, gh * gw * cout),
                                    bias=conv_mod.bias, algo_flops=2.0 * self.B * gh * gw * cout * cin_valid * 9))

    def _conv_head(self, name, conv_mod, src):
        """3x3 conv to <= 8 output channels, written as fp32 NHWC [B,H,W,8] (rows >= out_channels are zero weights)."""
        cout = conv_mod.weight.shape[0]
        assert cout <= 8
        gh, gw = src.shape[1], src.shape[2]
        tile = tile_box(gw, gh, square=True)
        plan = plan_conv("3x3", [src.shape[3]], cout, reuse_rows=can_reuse_rows(tile))
        n_rows, n_tile = n_tiling(8, False)
        pack = self.weights.add(f"{name}/R{plan.R}", conv_mod.weight, plan, n_rows)
        if "head_bias" not in self.weights.__dict__:
            self.weights.head_bias = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._head_bias_src = conv_mod.bias
        self.recs.append(TapGemmRec(name, plan, [nhwc_view(src)], gw, gh, self.B, tile, pack, pack.packed, pack.sched,
                                    n_rows, 8, n_tile, L.EPI_BIAS | L.EPI_OUT_F32, self.out8, (8, gw * 8, gh * gw * 8),
                                    bias=self.weights.head_bias,
                                    algo_flops=2.0 * self.B * gh * gw * cout * src.shape[3] * 9))

    def run(self, stream: int):
        """``stream`` must be torch's current stream: the two layout copies at the boundary are torch kernels (they are
        captured with the rest when the sampler records its CUDA graph)."""
        if stream != _current_stream():
            raise RuntimeError("VanillaProgram.run: launch on torch's current stream")
        b = self._head_bias_src
        if getattr(self.weights, "_head_bias_stamp", None) != (b.data_ptr(), b._version):
            self.weights.head_bias[: b.numel()].copy_(b.detach())
            self.weights._head_bias_stamp = (b.data_ptr(), b._version)
        self.glue_in()
        super().run(stream)
        self.glue_out()
