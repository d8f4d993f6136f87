"""Training-mode UNet: the forward and backward of ``Unet.forward`` (CCDM_unified/models/unet.py:382-455) as
``torch.autograd.Function`` nodes over the CUDA library, so that the reference's training step
(``loss = diffusion(...); loss.backward(); opt.step()``, trainer.py:617-734) runs unchanged.

Every node keeps NHWC bf16 activations and calls hand-written kernels only:

=====================  ====================================================  =======================================
node                   forward                                               backward
=====================  ====================================================  =======================================
``ConvFn``             ccdm_tapgemm (bias [+ residual])                      tap-GEMM data gradient, ccdm_conv_wgrad,
                                                                             ccdm_colsum_bf16
``ConvBlockFn``        ccdm_tapgemm -> z, ccdm_rmsnorm_act -> h              ccdm_block_bwd(+finish), dgrad, wgrad
``RmsNormFn``          ccdm_rmsnorm_act (PreNorm)                            ccdm_block_bwd
``LinAttnCoreFn``      ccdm_linattn_prep / _context, per-sample tap-GEMM     ccdm_linattn_dcontext / _bwd_rowdot /
                                                                             _pack_blockdiag, 3 per-sample tap-GEMMs,
                                                                             ccdm_linattn_bwd_finish
``AttnCoreFn``         ccdm_attention_small                                  ccdm_attention_small_bwd
``StemFn``             ccdm_stem_im2row + tap-GEMM                           ccdm_conv_wgrad + ccdm_stem_unpack_wgrad
``HeadFn``             ccdm_head_conv1                                       ccdm_head_conv1_bwd
=====================  ====================================================  =======================================

The conditioning path (sinusoidal time features, ``time_mlp``, ``cond_mlp_1/2`` with BatchNorm1d in batch-statistics
mode, the 23 ``tc_mlp`` Linears: [B,<=512] matrices, <0.1 % of the FLOPs) runs as plain PyTorch library calls (cuBLAS)
under autograd; it is glue, exactly the "plain library GEMM" case.  The saved-for-backward tensors are the bf16 conv
inputs and pre-norm conv outputs.  There is no CPU or PyTorch fallback for the nodes above.
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
from .plan import n_tiling, tile_box


def _require_cuda(x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError("ccdm_b200.Unet runs on sm_100a only (there is no CPU fallback)")


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _grad_slot(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """``p.grad`` when a gradient buffer is already attached to the parameter (FusedAdam's flat-buffer views, or any
    earlier accumulation): the kernels then add into it directly and the node returns None for that input, which saves
    the temporary, its zero fill and autograd's accumulate kernel.  None otherwise (autograd gets the tensor)."""
    if p is None or not isinstance(p, torch.nn.Parameter):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device or g.numel() != p.numel():
        return None
    return g


# ------------------------------------------------------------------------------------------------- nodes

class ConvFn(Function):
    """conv(kind)(cat(srcs)) + bias [+ resid]   (res_conv, down / up sampling, to_qkv, attention to_out)"""

    @staticmethod
    def forward(ctx, kind, weight, bias, resid, *srcs):
        out = K.conv_forward(kind, srcs, weight, bias, resid)
        ctx.kind, ctx.has_bias, ctx.has_resid = kind, bias is not None, resid is not None
        ctx.params = (weight, bias)
        ctx.save_for_backward(weight, *srcs)
        return out

    @staticmethod
    def backward(ctx, dz):
        dz = _c(dz)
        weight, *srcs = ctx.saved_tensors
        wparam, bparam = ctx.params
        cins = [s.shape[3] for s in srcs]
        dsrcs: List[Optional[torch.Tensor]] = [None] * len(srcs)
        if any(ctx.needs_input_grad[4:]):
            dsrcs = K.conv_dgrad(ctx.kind, dz, weight, cins)
        dw = db = None
        if ctx.needs_input_grad[1]:
            slot = _grad_slot(wparam)
            dw = K.conv_wgrad(ctx.kind, srcs, dz, accumulate_into=slot)
            dw = None if slot is not None else dw
        if ctx.has_bias and ctx.needs_input_grad[2]:
            slot = _grad_slot(bparam)
            db = K.colsum(dz, accumulate_into=slot)
            db = None if slot is not None else db
        return (None, dw, db, dz if ctx.has_resid else None, *dsrcs)


class ConvBlockFn(Function):
    """Block.forward (unet.py:143-152) [+ residual]: conv -> RMSNorm(g) -> (1+scale)*x+shift -> SiLU, keeping the
    pre-norm conv output z for the backward."""

    @staticmethod
    def forward(ctx, kind, silu, weight, bias, gain, ss, resid, *srcs):
        z = K.conv_forward(kind, srcs, weight, bias)
        ss = _c(ss) if ss is not None else None
        g = gain.reshape(-1)
        out = K.rmsnorm_act(z, g, ss, silu, resid)
        ctx.kind, ctx.silu, ctx.has_ss, ctx.has_resid = kind, silu, ss is not None, resid is not None
        ctx.gain_shape = gain.shape
        ctx.params = (weight, bias, gain)
        ctx.save_for_backward(weight, g, z, *( [ss] if ss is not None else [] ), *srcs)
        return out

    @staticmethod
    def backward(ctx, dh):
        dh = _c(dh)
        saved = list(ctx.saved_tensors)
        weight, g, z = saved[:3]
        ss = saved[3] if ctx.has_ss else None
        srcs = saved[4:] if ctx.has_ss else saved[3:]
        wparam, bparam, gparam = ctx.params
        gslot, bslot, wslot = _grad_slot(gparam), _grad_slot(bparam), _grad_slot(wparam)
        dz, d_ss, dgain, dbias = K.block_backward(dh, z, g, ss, 0, ctx.silu, dgain_into=gslot, dbias_into=bslot)
        cins = [s.shape[3] for s in srcs]
        dsrcs: List[Optional[torch.Tensor]] = [None] * len(srcs)
        if any(ctx.needs_input_grad[7:]):
            dsrcs = K.conv_dgrad(ctx.kind, dz, weight, cins)
        dw = K.conv_wgrad(ctx.kind, srcs, dz, accumulate_into=wslot)
        return (None, None, None if wslot is not None else dw, dbias,
                dgain.view(ctx.gain_shape) if dgain is not None else None, d_ss, dh if ctx.has_resid else None, *dsrcs)


class RmsNormFn(Function):
    """PreNorm's RMSNorm (unet.py:88-89,97-99): x / |x| * g * sqrt(C)."""

    @staticmethod
    def forward(ctx, x, gain):
        g = gain.reshape(-1)
        ctx.gain_shape = gain.shape
        ctx.params = (gain,)
        ctx.save_for_backward(x, g)
        return K.rmsnorm_act(x, g)

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        gslot = _grad_slot(ctx.params[0])
        dx, _, dgain, _ = K.block_backward(_c(dy), x, g, None, 0, False, dgain_into=gslot)
        return dx, dgain.view(ctx.gain_shape) if dgain is not None else None


class LinAttnCoreFn(Function):
    """qkv [B,h,w,384] -> out [B,h,w,128]: q softmax over channels * scale, k softmax over tokens, context = k v^T,
    out = context^T q  (unet.py:204-214; 4 heads x 32 as in the reference)."""

    @staticmethod
    def forward(ctx, qkv_raw, scale):
        b, h, w, c3 = qkv_raw.shape
        assert c3 == 384, "linear attention kernels are written for 4 heads x 32 channels (unet.py:190)"
        n = h * w
        lib, st, dev = L.lib(), K._stream(), qkv_raw.device
        qkv = qkv_raw.clone()
        kmax = torch.empty(b, 128, dtype=torch.float32, device=dev)
        L.check(lib.ccdm_linattn_prep(qkv.data_ptr(), b, n, kmax.data_ptr(), float(scale), st), "linattn_prep")
        cmat = torch.empty(b, 4, 32, 32, dtype=torch.float32, device=dev)
        colsum = torch.empty(b, 128, dtype=torch.float32, device=dev)
        L.check(lib.ccdm_linattn_context(qkv.data_ptr(), cmat.data_ptr(), colsum.data_ptr(), b, n, 4, None, None, 0, 0, st),
                "linattn_context")
        wb = torch.empty(b, 128, 128, dtype=torch.bfloat16, device=dev)
        L.check(lib.ccdm_linattn_pack_blockdiag(cmat.data_ptr(), None, 0, wb.data_ptr(), b, st), "linattn_pack_blockdiag")
        out = torch.empty(b, h, w, 128, dtype=torch.bfloat16, device=dev)
        K.per_sample_linear(qkv, 0, wb, out, 0)
        ctx.scale = float(scale)
        ctx.save_for_backward(qkv, cmat, colsum)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        qkv, cmat, colsum = ctx.saved_tensors
        b, h, w, _ = qkv.shape
        n = h * w
        lib, st, dev = L.lib(), K._stream(), qkv.device
        dctx = torch.empty_like(cmat)
        L.check(lib.ccdm_linattn_dcontext(qkv.data_ptr(), dout.data_ptr(), dctx.data_ptr(), b, n, st), "linattn_dcontext")
        cvec = torch.empty(b, 128, dtype=torch.float32, device=dev)
        L.check(lib.ccdm_linattn_bwd_rowdot(cmat.data_ptr(), dctx.data_ptr(), colsum.data_ptr(), cvec.data_ptr(), b, st),
                "linattn_bwd_rowdot")
        wq = torch.empty(b, 128, 128, dtype=torch.bfloat16, device=dev)    # dq_sm = ctx . dout
        wp = torch.empty_like(wq)                                          # dp    = G . v,  G = dctx / S
        wv = torch.empty_like(wq)                                          # dv    = G^T . p
        L.check(lib.ccdm_linattn_pack_blockdiag(cmat.data_ptr(), None, 1, wq.data_ptr(), b, st), "pack_blockdiag")
        L.check(lib.ccdm_linattn_pack_blockdiag(dctx.data_ptr(), colsum.data_ptr(), 1, wp.data_ptr(), b, st), "pack_blockdiag")
        L.check(lib.ccdm_linattn_pack_blockdiag(dctx.data_ptr(), colsum.data_ptr(), 0, wv.data_ptr(), b, st), "pack_blockdiag")
        dpre = torch.empty_like(qkv)
        K.per_sample_linear(dout, 0, wq, dpre, 0)
        K.per_sample_linear(qkv, 256, wp, dpre, 128)
        K.per_sample_linear(qkv, 128, wv, dpre, 256)
        L.check(lib.ccdm_linattn_bwd_finish(qkv.data_ptr(), dpre.data_ptr(), b, n, cvec.data_ptr(), ctx.scale, st),
                "linattn_bwd_finish")
        return dpre, None


class AttnCoreFn(Function):
    """Bottleneck softmax attention on the raw qkv (unet.py:231-239)."""

    @staticmethod
    def forward(ctx, qkv, heads, dim_head, scale):
        b, h, w, _ = qkv.shape
        out = torch.empty(b, h, w, heads * dim_head, dtype=torch.bfloat16, device=qkv.device)
        L.check(L.lib().ccdm_attention_small(qkv.data_ptr(), out.data_ptr(), b, h * w, heads, dim_head, float(scale),
                                             K._stream()), "attention_small")
        ctx.cfg = (heads, dim_head, float(scale))
        ctx.save_for_backward(qkv)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, = ctx.saved_tensors
        heads, dim_head, scale = ctx.cfg
        b, h, w, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        L.check(L.lib().ccdm_attention_small_bwd(qkv.data_ptr(), _c(dout).data_ptr(), dqkv.data_ptr(), b, h * w, heads,
                                                 dim_head, scale, K._stream()), "attention_small_bwd")
        return dqkv, None, None, None


def _stem_geometry(h, w, cout, dev):
    plan = K._plan_cached_stem(cout)
    n_rows, n_tile = n_tiling(cout, False)
    return plan, n_rows, n_tile, tile_box(w, h), K._dev_i32(plan.sched, dev)


class StemFn(Function):
    """init_conv (unet.py:271,418): 7x7 conv of the fp32 NCHW input, as a 4-tap tap-GEMM over the im2row tensor."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        b, cin, h, w = x.shape
        cout = weight.shape[0]
        lib, st, dev = L.lib(), K._stream(), x.device
        plan, n_rows, n_tile, tile, sched = _stem_geometry(h, w, cout, dev)
        rowimg = torch.empty(b, h + 1, w, 64, dtype=torch.bfloat16, device=dev)
        L.check(lib.ccdm_stem_im2row(_c(x).data_ptr(), rowimg.data_ptr(), b, cin, h, w, st), "stem_im2row")
        packed = torch.empty(n_rows, 256, dtype=torch.bfloat16, device=dev)
        L.check(lib.ccdm_stem_pack(weight.data_ptr(), packed.data_ptr(), cout, cin, n_rows, st), "stem_pack")
        out = torch.empty(b, h, w, cout, dtype=torch.bfloat16, device=dev)
        K._launch_tapgemm(plan, tile, [K._view(rowimg)], w, h, b, packed, sched, n_rows, cout, n_tile, out,
                          (cout, w * cout, h * w * cout), (0, 0, 0, 0), bias)
        ctx.save_for_backward(rowimg)
        ctx.shape = (cin, cout, h, w)
        return out

    @staticmethod
    def backward(ctx, dz):
        dz = _c(dz)
        rowimg, = ctx.saved_tensors
        cin, cout, h, w = ctx.shape
        dev = dz.device
        plan, n_rows, _, tile, sched = _stem_geometry(h, w, cout, dev)
        gpacked = K.wgrad_packed(plan, tile, [K._view(rowimg)], dz, w, h, sched, cout, n_rows)
        dw = torch.empty(cout, cin, 7, 7, dtype=torch.float32, device=dev)
        L.check(L.lib().ccdm_stem_unpack_wgrad(gpacked.data_ptr(), dw.data_ptr(), cout, cin, 0, K._stream()),
                "stem_unpack_wgrad")
        return None, dw, K.colsum(dz)


class HeadFn(Function):
    """final_conv (unet.py:348,455): 1x1 conv to the fp32 NCHW output."""

    @staticmethod
    def forward(ctx, hfeat, weight, bias):
        b, h, w, cin = hfeat.shape
        cout = weight.shape[0]
        out = torch.empty(b, cout, h, w, dtype=torch.float32, device=hfeat.device)
        L.check(L.lib().ccdm_head_conv1(hfeat.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), b, h, w, cin,
                                        cout, K._stream()), "head_conv1")
        ctx.save_for_backward(hfeat, weight)
        return out

    @staticmethod
    def backward(ctx, dout):
        hfeat, weight = ctx.saved_tensors
        b, h, w, cin = hfeat.shape
        cout = weight.shape[0]
        dev = hfeat.device
        dout = _c(dout.float())
        dh = torch.empty_like(hfeat)
        dw = torch.zeros(cout, cin, dtype=torch.float32, device=dev)
        db = torch.zeros(cout, dtype=torch.float32, device=dev)
        L.check(L.lib().ccdm_head_conv1_bwd(dout.data_ptr(), hfeat.data_ptr(), weight.data_ptr(), dh.data_ptr(),
                                            dw.data_ptr(), db.data_ptr(), b, h, w, cin, cout, K._stream()), "head_conv1_bwd")
        return dh, dw.view_as(weight), db


# ------------------------------------------------------------------------------------------------- composition

def time_features(t: torch.Tensor, dim: int) -> torch.Tensor:
    """SinusoidalPosEmb (unet.py:107-115): sin | cos with exponent step log(10000)/(half-1)."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -k)
    ang = t.reshape(-1)[:, None] * freq[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def _resblock(mod, srcs, ss):
    """ResnetBlock.forward (unet.py:167-187); ``ss`` = this block's tc_mlp output [B, 2*Cout] (scale | shift)."""
    b1, b2 = mod.block1, mod.block2
    h1 = ConvBlockFn.apply("3x3", True, b1.proj.weight, b1.proj.bias, b1.norm.g, ss, None, *srcs)
    if isinstance(mod.res_conv, nn.Conv2d):
        res = ConvFn.apply("1x1", mod.res_conv.weight, mod.res_conv.bias, None, *srcs)
    else:
        res = srcs[0]
    return ConvBlockFn.apply("3x3", True, b2.proj.weight, b2.proj.bias, b2.norm.g, None, res, h1)


def _linear_attention(mod, x):
    """Residual(PreNorm(LinearAttention)) (unet.py:66-72,91-99,202-216)."""
    pre, att = mod.fn, mod.fn.fn
    assert att.heads == 4 and att.dim_head == 32, "linear attention kernels are written for 4 heads x 32 (unet.py:190)"
    xn = RmsNormFn.apply(x, pre.norm.g)
    qkv = ConvFn.apply("1x1", att.to_qkv.weight, None, None, xn)
    o = LinAttnCoreFn.apply(qkv, att.scale)
    conv, norm = att.to_out[0], att.to_out[1]
    return ConvBlockFn.apply("1x1", False, conv.weight, conv.bias, norm.g, None, x, o)


def _mid_attention(mod, x):
    """Residual(PreNorm(Attention)) (unet.py:228-240)."""
    pre, att = mod.fn, mod.fn.fn
    xn = RmsNormFn.apply(x, pre.norm.g)
    qkv = ConvFn.apply("1x1", att.to_qkv.weight, None, None, xn)
    o = AttnCoreFn.apply(qkv, att.heads, att.dim_head, att.scale)
    return ConvFn.apply("1x1", att.to_out.weight, att.to_out.bias, x, o)


def unet_train_forward(net, x: torch.Tensor, t: torch.Tensor, labels_emb: torch.Tensor,
                       keep_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Unet.forward in training mode with an autograd graph; returns the fp32 NCHW prediction.

    ``keep_mask`` is the Bernoulli mask ``Unet.forward`` drew (None when cond_drop_prob == 0)."""
    _require_cuda(x)
    b = x.shape[0]
    # conditioning (unet.py:397-414,421-423): tiny fp32 matrices, PyTorch library calls under autograd
    c = net.cond_mlp_1(labels_emb.float())
    if keep_mask is not None:
        c = torch.where(keep_mask[:, None], c, net.null_cond_emb[None, :].expand(b, -1).to(c.dtype))
    c = net.cond_mlp_2(c)
    te = time_features(t, net.dim)
    te = net.time_mlp[3](net.time_mlp[2](net.time_mlp[1](te)))
    tc = F.silu(torch.cat((te, c), dim=1))
    # every ResnetBlock's tc_mlp Linear (unet.py:158-161,171-176) reads the same [SiLU(t) | SiLU(c)] rows: ONE library GEMM over
    # the concatenated weights (and one dgrad + one wgrad GEMM in the backward) instead of 23-31 small ones
    blocks = []
    for lvl in net.downs:
        blocks += [lvl[0], lvl[1]]
    blocks += [net.mid_block1, net.mid_block2]
    for lvl in net.ups:
        blocks += [lvl[0], lvl[1]]
    blocks.append(net.final_res_block)
    ss_all = F.linear(tc, torch.cat([m.tc_mlp[1].weight for m in blocks]), torch.cat([m.tc_mlp[1].bias for m in blocks]))
    ss_it = iter(torch.split(ss_all, [m.tc_mlp[1].weight.shape[0] for m in blocks], dim=1))
    K.ARENA.begin_step(x.device)
    K.PACKS.begin_step(x.device)                                   # every conv weight re-packed (both layouts) in one launch

    stem = StemFn.apply(x.float(), net.init_conv.weight, net.init_conv.bias)
    h = stem
    skips = []
    nlev = len(net.downs)
    for k, (b1, b2, attn, down) in enumerate(net.downs):
        h = _resblock(b1, [h], next(ss_it))
        skips.append(h)
        h = _resblock(b2, [h], next(ss_it))
        h = _linear_attention(attn, h)
        skips.append(h)
        h = ConvFn.apply("3x3" if k == nlev - 1 else "down4x4s2", down.weight, down.bias, None, h)
    h = _resblock(net.mid_block1, [h], next(ss_it))
    h = _mid_attention(net.mid_attn, h)
    h = _resblock(net.mid_block2, [h], next(ss_it))
    hook = getattr(net, "_early_grad_hook", None)
    if hook is not None and h.requires_grad:
        # the gradient of the decoder's input exists once every decoder node has run, i.e. once all decoder weight gradients
        # are in the optimizer's flat buffer: start their all-reduce now, under the encoder's backward (dist.wire_overlap)
        h.register_hook(lambda g, _hook=hook: (_hook(), None)[1])
    for k, (b1, b2, attn, up) in enumerate(net.ups):
        h = _resblock(b1, [h, skips.pop()], next(ss_it))
        h = _resblock(b2, [h, skips.pop()], next(ss_it))
        h = _linear_attention(attn, h)
        if k == nlev - 1:
            h = ConvFn.apply("3x3", up.weight, up.bias, None, h)
        else:
            h = ConvFn.apply("up2x3x3", up[1].weight, up[1].bias, None, h)
    h = _resblock(net.final_res_block, [h, stem], next(ss_it))
    return HeadFn.apply(h, net.final_conv.weight, net.final_conv.bias)
