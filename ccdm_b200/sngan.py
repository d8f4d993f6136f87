"""One-step generator of the DMD2-M distillation: ``sngan_generator`` with the
constructor, ``forward(z, y)`` signature and ``state_dict`` layout of CCDM_unified/models/sngan.py:19-139, evaluated by
the CUDA library in eval mode (BatchNorm running statistics), which is how ``dmd.py:878-881`` samples from it.

Forward of one ``ResBlockGenerator`` (sngan.py:72-83), with every BatchNorm2d(+conditional gain) reduced to a
per-(sample, channel) affine map ``x*a + s`` by ``ccdm_condbn_coef``:

    h0 = relu(x*a1 + s1)                       ccdm_affine_act                      (pre-activation, HBM-bound)
    h1 = relu((up2x(h0) * W1 + b1)*a2 + s2)    ccdm_tapgemm "up2x3x3", epilogue BIAS | SS | RELU  (nearest-2x folded:
                                                four output-parity 2x2 convs, 2.25x fewer MACs than 3x3 on the 4x tensor)
    by = up2x(x) * Wb + bb                     ccdm_tapgemm "up2x1x1"  (1x1 commutes with the upsampling)
    out = h1 * W2 + b2 + by                    ccdm_tapgemm "3x3", epilogue BIAS | RESID

The dense layer is a row tap-GEMM whose weight rows are permuted once so that its output already is the NHWC
``[B, s, s, C0]`` tensor; the output conv (gene_ch -> nc, tanh) writes fp32 with the weight rows padded to 8.
Training mode (batch statistics) is not built: ``forward`` raises in ``.train()``.  No CPU / PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from . import _lib as L
from . import backward as K
from .plan import KB, n_tiling


class ConditionalBatchNorm2d(nn.Module):            # sngan.py:19-36 (parameter holder)
    def __init__(self, num_features, dim_embed):
        super().__init__()
        self.num_features = num_features
        self.bn = nn.BatchNorm2d(num_features, affine=False)
        self.embed_gamma = nn.Linear(dim_embed, num_features, bias=False)
        self.embed_beta = nn.Linear(dim_embed, num_features, bias=False)


class ResBlockGenerator(nn.Module):                 # sngan.py:39-83 (parameter holder, same registration order)
    def __init__(self, in_channels, out_channels, dim_embed, bias=True):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, padding=1, bias=bias)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, padding=1, bias=bias)
        nn.init.xavier_uniform_(self.conv1.weight.data, math.sqrt(2))
        nn.init.xavier_uniform_(self.conv2.weight.data, math.sqrt(2))
        self.condbn1 = ConditionalBatchNorm2d(in_channels, dim_embed)
        self.condbn2 = ConditionalBatchNorm2d(out_channels, dim_embed)
        self.relu = nn.ReLU()
        self.upsample = nn.Upsample(scale_factor=2)
        # the reference also registers the unconditional branch, which shares conv1 / conv2 (state-dict keys model.*)
        self.model = nn.Sequential(nn.BatchNorm2d(in_channels), nn.ReLU(), nn.Upsample(scale_factor=2), self.conv1,
                                   nn.BatchNorm2d(out_channels), nn.ReLU(), self.conv2)
        self.bypass_conv = nn.Conv2d(in_channels, out_channels, 1, 1, padding=0, bias=bias)
        nn.init.xavier_uniform_(self.bypass_conv.weight.data, 1.0)
        self.bypass = nn.Sequential(nn.Upsample(scale_factor=2), self.bypass_conv)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(z):
    if not z.is_cuda:
        raise RuntimeError("ccdm_b200.sngan_generator runs on sm_100a only (there is no CPU fallback)")


class sngan_generator(nn.Module):                   # noqa: N801  (the reference's class name)
    def __init__(self, dim_z=128, dim_embed=128, nc=3, img_size=64, gene_ch=32, ch_multi=None):
        super().__init__()
        self.dim_z, self.dim_embed, self.gene_ch, self.img_size, self.nc = dim_z, dim_embed, gene_ch, img_size, nc
        assert self.img_size in [64, 128, 192]
        if ch_multi is None:
            ch_multi = [16, 8, 4, 2, 1]
        assert len(ch_multi) >= 5
        self.ch_multi = ch_multi
        self.init_size = 4 if self.img_size in [64, 128] else 6
        self.dense = nn.Linear(self.dim_z, self.init_size * self.init_size * gene_ch * ch_multi[0], bias=True)
        final = nn.Conv2d(gene_ch, nc, 3, stride=1, padding=1, bias=True)
        nn.init.xavier_uniform_(self.dense.weight.data, 1.)
        nn.init.xavier_uniform_(final.weight.data, 1.)
        self.final = final                            # re-registered below inside the Sequential, as in the reference
        chans = [gene_ch * m for m in ch_multi[:5]]
        for i in range(4):
            setattr(self, f"genblock{i}", ResBlockGenerator(chans[i], chans[i + 1], dim_embed=dim_embed))
        if self.img_size in [128, 192]:
            self.genblock4 = ResBlockGenerator(chans[4], gene_ch, dim_embed=dim_embed)
        self.final = nn.Sequential(nn.BatchNorm2d(gene_ch), nn.ReLU(), final, nn.Tanh())
        self._cache = {}

    # ------------------------------------------------------------------ helpers
    def _blocks(self) -> List[ResBlockGenerator]:
        n = 5 if self.img_size in [128, 192] else 4
        return [getattr(self, f"genblock{i}") for i in range(n)]

    def _pack(self, kind, weight, cins, cout, gw, gh, plan=None, tile=None):
        """bf16 K-blocked copy of a conv weight, cached until the parameter changes (data_ptr, _version)."""
        if plan is None:
            plan, tile = K._plan_for(kind, cins, cout, gw, gh)
        key = (kind, weight.data_ptr(), weight._version, plan.R, cout)
        hit = self._cache.get(key)
        n_rows, n_tile = n_tiling(cout, False)
        dev = weight.device
        sched = K._dev_i32(plan.sched, dev)
        if hit is None:
            packed = torch.empty(plan.nz * n_rows, plan.nkb * KB, dtype=torch.bfloat16, device=dev)
            psched = K._dev_i32(plan.psched, dev)
            L.check(L.lib().ccdm_pack_weights(weight.data_ptr(), cout, weight.shape[1], weight[0, 0].numel(),
                                              psched.data_ptr(), plan.nz, plan.nkb, n_rows, None, 1.0, packed.data_ptr(),
                                              _stream()), "pack_weights")
            stale = [k for k in self._cache if isinstance(k, tuple) and k[:2] == key[:2] and k != key]
            for k in stale:
                del self._cache[k]
            self._cache[key] = hit = packed
        return plan, tile, hit, sched, n_rows, n_tile

    def _conv(self, kind, src, weight, bias, *, flags=0, ss=None, resid=None, out_f32=False, cout=None):
        b, h, w, cin = src.shape
        cout = cout or weight.shape[0]
        up = kind.startswith("up2x")
        oh, ow = (2 * h, 2 * w) if up else (h, w)
        plan, tile = K._plan_for(kind, (cin,), cout, w, h)
        plan, tile, packed, sched, n_rows, n_tile = self._pack(kind, weight, (cin,), cout, w, h, plan, tile)
        out = torch.empty(b, oh, ow, cout, dtype=torch.float32 if out_f32 else torch.bfloat16, device=src.device)
        ostr, ooff = K._out_geometry(out, plan.out_parity)
        K._launch_tapgemm(plan, tile, [K._view(src)], w, h, b, packed, sched, n_rows, cout, n_tile, out, ostr, ooff, bias,
                          resid, flags=flags | (L.EPI_OUT_F32 if out_f32 else 0), ss=ss)
        return out

    def _batch_stats(self, bn: nn.BatchNorm2d, x: torch.Tensor):
        """Training-mode BatchNorm2d statistics of a bf16 NHWC tensor (sngan.py:23,30: ``self.bn(x)`` with the module in
        train mode): per-(sample, channel) sums by ``ccdm_channel_stats``, folded over the batch; biased variance for the
        normalisation, running statistics updated with momentum and the unbiased variance exactly like nn.BatchNorm2d."""
        b, h, w, c = x.shape
        sums = torch.empty(b, 2, c, dtype=torch.float32, device=x.device)
        L.check(L.lib().ccdm_channel_stats(x.data_ptr(), b, h * w, c, sums.data_ptr(), c, 0, 1, _stream()), "channel_stats")
        n = b * h * w
        tot = sums.sum(0)
        mean = tot[0] / n
        var = (tot[1] / n - mean * mean).clamp_min_(0.0)
        if bn.track_running_stats and bn.running_mean is not None:
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var * (n / max(n - 1, 1)), alpha=m)
            bn.num_batches_tracked += 1
        return mean.contiguous(), var.contiguous()

    def _coef(self, bn: nn.BatchNorm2d, y: Optional[torch.Tensor], cbn: Optional[ConditionalBatchNorm2d], b: int,
              stats=None):
        """[B, 2C] scale | shift of (Conditional)BatchNorm2d: running statistics (eval) or the given batch ``stats``."""
        lib, st = L.lib(), _stream()
        c = bn.num_features
        dev = bn.running_mean.device
        mean, var = (bn.running_mean, bn.running_var) if stats is None else stats
        gamma = beta = None
        if cbn is not None:
            gamma = torch.empty(b, c, dtype=torch.float32, device=dev)
            beta = torch.empty(b, c, dtype=torch.float32, device=dev)
            for lin, dst in ((cbn.embed_gamma, gamma), (cbn.embed_beta, beta)):
                L.check(lib.ccdm_linear_small(y.data_ptr(), b, y.shape[1], lin.weight.data_ptr(), None, c, None, None, None,
                                              None, 0, L.ACT_NONE, dst.data_ptr(), c, st), "linear_small")
        ss = torch.empty(b, 2 * c, dtype=torch.float32, device=dev)
        L.check(lib.ccdm_condbn_coef(L.ptr(gamma), L.ptr(beta), L.ptr(bn.weight) if bn.affine else None,
                                     L.ptr(bn.bias) if bn.affine else None, mean.data_ptr(),
                                     var.data_ptr(), float(bn.eps), b, c, ss.data_ptr(), st), "condbn_coef")
        return ss

    @staticmethod
    def _affine_relu(x, ss):
        b, h, w, c = x.shape
        out = torch.empty_like(x)
        L.check(L.lib().ccdm_affine_act(x.data_ptr(), out.data_ptr(), b * h * w, c, h * w, ss.data_ptr(), ss.shape[1], 0, 1,
                                        _stream()), "affine_act")
        return out

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, z, y):
        """sngan.py:130-139; ``y`` is the embedded label ``fn_y2h(labels)`` [B, dim_embed].  Returns NCHW fp32 in (-1, 1).
        Eval mode: running BatchNorm statistics, the second CondBN + ReLU of every block fused into conv1's epilogue.
        Train mode (forward only -- the generator's backward belongs to the out-of-scope DMD2 trainer): batch statistics,
        running statistics updated in place as nn.BatchNorm2d does; conv1's output is then normalised by a separate pass."""
        train = self.training
        _require_cuda(z)
        if y is None:
            raise NotImplementedError("the unconditional branch (sngan.py:84-85) is never taken by dmd.py")
        b = z.shape[0]
        s, c0 = self.init_size, self.gene_ch * self.ch_multi[0]
        y = y.reshape(b, -1).float().contiguous()
        zb = z.reshape(b, 1, 1, -1).to(torch.bfloat16).contiguous()
        # dense (sngan.py:132-133): rows permuted (c, h, w) -> (h, w, c) so that the GEMM output is NHWC
        key = (self.dense.weight.data_ptr(), self.dense.weight._version, self.dense.bias._version)
        if self._cache.get("dense_key") != key:
            wd = self.dense.weight.detach().view(c0, s * s, self.dim_z).permute(1, 0, 2).reshape(s * s * c0, self.dim_z, 1, 1)
            self._cache["dense_w"] = wd.contiguous()
            self._cache["dense_b"] = self.dense.bias.detach().view(c0, s * s).t().reshape(-1).contiguous()
            self._cache["dense_key"] = key
        x = self._conv("1x1", zb.view(1, 1, b, -1), self._cache["dense_w"], self._cache["dense_b"])
        x = x.view(b, s, s, c0)
        for blk in self._blocks():
            ss1 = self._coef(blk.condbn1.bn, y, blk.condbn1, b, self._batch_stats(blk.condbn1.bn, x) if train else None)
            h0 = self._affine_relu(x, ss1)
            if train:                                             # batch statistics of conv1's output: cannot live in its epilogue
                h1raw = self._conv("up2x3x3", h0, blk.conv1.weight, blk.conv1.bias)
                ss2 = self._coef(blk.condbn2.bn, y, blk.condbn2, b, self._batch_stats(blk.condbn2.bn, h1raw))
                h1 = self._affine_relu(h1raw, ss2)
            else:
                ss2 = self._coef(blk.condbn2.bn, y, blk.condbn2, b)
                h1 = self._conv("up2x3x3", h0, blk.conv1.weight, blk.conv1.bias, flags=L.EPI_RELU, ss=ss2)
            by = self._conv("up2x1x1", x, blk.bypass_conv.weight, blk.bypass_conv.bias)
            x = self._conv("3x3", h1, blk.conv2.weight, blk.conv2.bias, resid=by)
        bn, conv = self.final[0], self.final[2]
        h = self._affine_relu(x, self._coef(bn, None, None, b, self._batch_stats(bn, x) if train else None))
        # output conv: nc (3) rows padded to 8 (the GEMM's channel granularity); fp32 NHWC, tanh in the epilogue
        key = (conv.weight.data_ptr(), conv.weight._version, conv.bias._version)
        if self._cache.get("final_key") != key:
            w8 = torch.zeros(8, conv.weight.shape[1], 3, 3, dtype=torch.float32, device=z.device)
            w8[: self.nc] = conv.weight.detach()
            b8 = torch.zeros(8, dtype=torch.float32, device=z.device)
            b8[: self.nc] = conv.bias.detach()
            self._cache.update(final_w=w8, final_b=b8, final_key=key)
        o = self._conv("3x3", h, self._cache["final_w"], self._cache["final_b"], flags=L.EPI_TANH, out_f32=True, cout=8)
        return o[..., : self.nc].permute(0, 3, 1, 2).contiguous()
