"""Backward building blocks of the training step, over NHWC bf16 activations (autograd of the conv / Block call
sites of CCDM_unified/models/unet.py:77-89,139-165,195-226,326-341; the reference reaches them through
``loss.backward()`` in trainer.py).

Three kernels carry ~all of the backward FLOPs and bytes:

* data gradient   ``conv_dgrad``   = ``ccdm_tapgemm`` over dY with weights packed by ``ccdm_pack_weights_t``
  (flipped filter; the 4x4/stride-2 and nearest-2x+3x3 convolutions trade schedules, see plan.py ``*_dgrad``);
* weight gradient ``conv_wgrad``   = ``ccdm_conv_wgrad`` (tcgen05, positions as the K axis, split-K) +
  ``ccdm_unpack_wgrad``;
* Block tail      ``block_backward`` = ``ccdm_block_bwd`` (+ ``_finish``): RMSNorm / scale-shift / SiLU.

These are eager, allocation-per-call wrappers; ccdm_b200/train.py chains them (as torch.autograd.Function nodes) into
the backward of the whole UNet.  There is no fallback: every function runs the CUDA library or raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import weakref
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .plan import HALO_TILE, KB, ConvPlan, can_reuse_rows, halo_ok, n_tiling, plan_conv, tile_box

_KIND_TAPS = {"1x1": 1, "3x3": 9, "down4x4s2": 16, "down3x3s2": 9, "up2x3x3": 9}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _view(t: torch.Tensor, c_off: int = 0, c: Optional[int] = None) -> L.View:
    b, h, w, ct = t.shape
    return L.View(t.data_ptr() + 2 * c_off, ct - c_off if c is None else c, w, h, b, ct, w * ct, h * w * ct)


def _parity_views(t: torch.Tensor) -> List[L.View]:
    b, h, w, ct = t.shape
    return [L.View(t.data_ptr() + 2 * (pr * w + pq) * ct, ct, (w - pq + 1) // 2, (h - pr + 1) // 2, b, 2 * ct,
                   2 * w * ct, h * w * ct) for pr in range(2) for pq in range(2)]


_I32_CACHE: dict = {}


def _dev_i32(rows, device) -> torch.Tensor:
    """Schedule tables are tiny and immutable: one device copy per (table, device)."""
    key = (id(rows), str(device))
    hit = _I32_CACHE.get(key)
    if hit is None or hit[0] is not rows:
        with torch.inference_mode(False):
            hit = (rows, torch.tensor(rows, dtype=torch.int32, device=device).contiguous())
        _I32_CACHE[key] = hit
    return hit[1]


def _check(t: torch.Tensor, what: str):
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.dim() == 4):
        raise ValueError(f"{what}: expected a contiguous CUDA bf16 [B,H,W,C] tensor, got {tuple(t.shape)} {t.dtype} {t.device}")


def _plan_for(kind: str, cins: Sequence[int], cout: int, gw: int, gh: int) -> Tuple[ConvPlan, Tuple[int, int, int]]:
    return _plan_cached(kind, tuple(cins), cout, gw, gh)


@functools.lru_cache(maxsize=None)
def _plan_cached(kind, cins, cout, gw, gh):
    base = kind[:-6] if kind.endswith("_dgrad") else kind
    tile = tile_box(gw, gh, square=(base != "1x1"))
    reuse = base != "1x1" and can_reuse_rows(tile)
    return plan_conv(kind, cins, cout, reuse_rows=reuse), tile


WGRAD_HALO = os.environ.get("CCDM_WGRAD_HALO", "1") != "0"       # A/B switch
KSTEPS = os.environ.get("CCDM_KSTEPS", "1") != "0"               # A/B switch: skip the empty K steps of partial channel blocks


@functools.lru_cache(maxsize=None)
def _wgrad_plan_for(kind, cins, cout, gw, gh):
    """Weight-gradient plan: 3x3 layers on grids of 8 x 16 tiles take the halo plan (R == 9 in ccdm_wgrad_args: one dZ box
    and one X halo box per tile feed all nine taps, so dZ and X cross L2 -> SM once per 64-channel block instead of three
    times); everything else follows the forward's load groups."""
    # Measured (profiles/r2_notes.md): the halo kernel wins whenever the last 128-channel tile of the output would be mostly
    # padding for the group kernel (Cout = 64, 72, 144, 288 ...); for Cout = 128 / 256 its 96-column accumulators need one
    # more pass over X than the group kernel's 128-row tiles and it loses 8-15 %.
    if WGRAD_HALO and halo_ok(kind, gw, gh) and 0 < cout % 128 <= 96:
        return plan_conv(kind, cins, cout, halo=True), HALO_TILE
    return _plan_cached(kind, cins, cout, gw, gh)


TRAIN_HALO = os.environ.get("CCDM_TRAIN_HALO", "1") != "0"       # A/B switch


@functools.lru_cache(maxsize=None)
def _conv_plan_for(kind, cins, cout, gw, gh):
    """Plan of a training-mode forward convolution or data gradient: 3x3 layers whose (pair-split) weights stay resident in
    shared memory take the halo plan (one box per 64-channel block and tile instead of three), under the inference engine's
    criterion.  Measured (profiles/r2_notes.md): 64 -> 64 at 64x64, batch 128: forward 57 -> 51 us, data gradient 57 -> 54 us;
    the 72-channel layers, whose 108 KB of weights only fit next to three boxes without the second output staging buffer,
    LOSE with it (forward 116 -> 145 us) and stay on the three-box plan."""
    base = kind[:-6] if kind.endswith("_dgrad") else kind
    if TRAIN_HALO and halo_ok(base, gw, gh) and cout <= 128:
        nkb = 9 * sum(-(-c // KB) for c in cins)
        resident = nkb * ((cout + 31) // 32 * 32) * 128 // (2 if nkb >= 16 else 1)
        if resident <= 96 * 1024:
            return plan_conv(kind, cins, cout, halo=True), HALO_TILE
    return _plan_cached(kind, cins, cout, gw, gh)


def _launch_tapgemm(plan: ConvPlan, tile, views: List[L.View], gw, gh, gb, wpacked, sched, n_rows, n, n_tile, out,
                    ostr, ooff, bias=None, resid: Optional[torch.Tensor] = None, w_batch_rows: int = 0, out_c_off: int = 0,
                    flags: int = 0, ss: Optional[torch.Tensor] = None):
    a = L.TapGemmArgs()
    a.n_src = len(views)
    for i, v in enumerate(views):
        a.src[i] = v
    a.gW, a.gH, a.gB = gw, gh, gb
    a.tw, a.th, a.tb = tile
    a.nz, a.ngroups, a.R = plan.nz, plan.ngroups, plan.R
    a.halo = int(plan.halo)
    a.sched, a.wpacked = sched.data_ptr(), wpacked.data_ptr()
    a.ksteps = _dev_i32(plan.ksteps, sched.device).data_ptr() if (KSTEPS and plan.ksteps is not None) else None
    a.n_rows, a.w_batch_rows, a.N, a.n_tile = n_rows, w_batch_rows, n, n_tile
    a.flags = flags | (L.EPI_BIAS if bias is not None else 0) | (L.EPI_RESID if resid is not None else 0)
    a.bias = L.ptr(bias)
    if ss is not None:
        a.flags |= L.EPI_SS
        a.scale_shift, a.ss_ld, a.ss_off = ss.data_ptr(), ss.shape[1], 0
    if resid is not None:
        rc = resid.shape[3]
        a.resid = resid.data_ptr()
        a.rsW, a.rsH, a.rsB = rc, resid.shape[2] * rc, resid.shape[1] * resid.shape[2] * rc
    a.out = out.data_ptr() + out.element_size() * out_c_off
    a.osW, a.osH, a.osB = ostr
    for i in range(L.MAX_Z):
        a.ooff[i] = ooff[i]
    L.check(L.lib().ccdm_tapgemm(C.byref(a), _stream()), f"tapgemm[{plan.kind}]")


def _out_geometry(out: torch.Tensor, parity: bool):
    _, h, w, co = out.shape
    if parity:
        return (2 * co, 2 * w * co, h * w * co), tuple((pa * w + pb) * co for pa in range(2) for pb in range(2))
    return (co, w * co, h * w * co), (0, 0, 0, 0)


def conv_forward(kind: str, srcs: Sequence[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 resid: Optional[torch.Tensor] = None):
    """z = conv(concat(srcs)) + bias [+ resid] with the plain epilogue (what a training forward keeps for the backward)."""
    for s in srcs:
        _check(s, "conv_forward source")
    b, h, w, _ = srcs[0].shape
    cout = weight.shape[0]
    oh, ow = {"1x1": (h, w), "3x3": (h, w), "down4x4s2": (h // 2, w // 2), "down3x3s2": ((h + 1) // 2, (w + 1) // 2),
              "up2x3x3": (2 * h, 2 * w)}[kind]
    gh, gw = (h, w) if kind == "up2x3x3" else (oh, ow)
    plan, tile = _conv_plan_for(kind, tuple(s.shape[3] for s in srcs), cout, gw, gh)
    n_rows, n_tile = n_tiling(cout, False)
    dev = weight.device
    sched, psched = _dev_i32(plan.sched, dev), _dev_i32(plan.psched, dev)
    packed, need = PACKS.get(weight, 0, kind, plan, n_rows, 0, cout, psched)
    if need:
        L.check(L.lib().ccdm_pack_weights(weight.data_ptr(), cout, weight.shape[1], _KIND_TAPS[kind], psched.data_ptr(),
                                          plan.nz, plan.nkb, n_rows, None, 1.0, packed.data_ptr(), _stream()), "pack_weights")
    views: List[L.View] = []
    for s in srcs:
        views += _parity_views(s) if plan.n_views == 4 else [_view(s)]
    out = torch.empty(b, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    ostr, ooff = _out_geometry(out, plan.out_parity)
    _launch_tapgemm(plan, tile, views, gw, gh, b, packed, sched, n_rows, cout, n_tile, out, ostr, ooff, bias, resid)
    return out


def per_sample_linear(src: torch.Tensor, c_off: int, w_batch: torch.Tensor, out: torch.Tensor, out_c_off: int):
    """out[b,h,w,out_c_off:out_c_off+128] = W_b . src[b,h,w,c_off:c_off+128] with one 128x128 bf16 matrix per sample
    (the block-diagonal context products of the linear attention, forward and backward)."""
    b, h, w, _ = src.shape
    plan, _ = _plan_for("1x1", (128,), 128, w, h)
    tile = tile_box(w, h, force_tb1=True)
    sched = _dev_i32(plan.sched, src.device)
    co = out.shape[3]
    _launch_tapgemm(plan, tile, [_view(src, c_off, 128)], w, h, b, w_batch, sched, 128, 128, 128, out,
                    (co, w * co, h * w * co), (0, 0, 0, 0), w_batch_rows=128, out_c_off=out_c_off)
    return out


def rmsnorm_act(z: torch.Tensor, gain: torch.Tensor, scale_shift: Optional[torch.Tensor] = None, silu: bool = False,
                resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = [silu]( z/|z| * gain*sqrt(C) * (1+scale) + shift ) [+ resid]   (unet.py:88-89,145-151)"""
    _check(z, "rmsnorm_act z")
    b, h, w, c = z.shape
    out = torch.empty_like(z)
    flags = (L.EPI_SILU if silu else 0) | (L.EPI_SS if scale_shift is not None else 0) | (L.EPI_RESID if resid is not None else 0)
    ld = scale_shift.shape[1] if scale_shift is not None else 0
    L.check(L.lib().ccdm_rmsnorm_act(z.data_ptr(), out.data_ptr(), b * h * w, c, h * w, gain.data_ptr(), math.sqrt(c),
                                     L.ptr(scale_shift), ld, 0, L.ptr(resid), None, flags, _stream()), "rmsnorm_act")
    return out


def colsum(x: torch.Tensor, accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [C] column sums of a bf16 [..., C] tensor (bias gradient of a plain convolution); added to
    ``accumulate_into`` when given."""
    c = x.shape[-1]
    out = accumulate_into if accumulate_into is not None else torch.zeros(c, dtype=torch.float32, device=x.device)
    L.check(L.lib().ccdm_colsum_bf16(x.data_ptr(), x.numel() // c, c, out.data_ptr(), _stream()), "colsum_bf16")
    return out


def conv_dgrad(kind: str, dy: torch.Tensor, weight: torch.Tensor, cins: Sequence[int]) -> List[torch.Tensor]:
    """Gradients w.r.t. each concatenated source of ``conv(kind)``: list of [B,H,W,cin_i] bf16 tensors."""
    _check(dy, "conv_dgrad dy")
    b, oh, ow, cout = dy.shape
    if cout != weight.shape[0] or sum(cins) != weight.shape[1]:
        raise ValueError(f"conv_dgrad: dy has {cout} channels, sources {tuple(cins)}, weight {tuple(weight.shape)}")
    h, w = {"1x1": (oh, ow), "3x3": (oh, ow), "down4x4s2": (2 * oh, 2 * ow), "down3x3s2": (2 * oh, 2 * ow),
            "up2x3x3": (oh // 2, ow // 2)}[kind]
    # the GEMM grid: positions of dx for stride-1 kinds and the upsampling conv, of dy (= one parity plane of dx) for
    # the stride-2 conv
    gh, gw = (oh, ow) if kind in ("down4x4s2", "down3x3s2") else (h, w)
    dev = weight.device
    outs = []
    n_off = 0
    for cin in cins:
        plan, tile = _conv_plan_for(kind + "_dgrad", (cout,), cin, gw, gh)
        n_rows, n_tile = n_tiling(cin, False)
        sched, psched = _dev_i32(plan.sched, dev), _dev_i32(plan.psched, dev)
        packed, need = PACKS.get(weight, 1, kind, plan, n_rows, n_off, cin, psched)
        if need:
            L.check(L.lib().ccdm_pack_weights_t(weight.data_ptr(), cout, weight.shape[1], _KIND_TAPS[kind], psched.data_ptr(),
                                                plan.nz, plan.nkb, n_rows, n_off, cin, packed.data_ptr(), _stream()),
                    "pack_weights_t")
        views = _parity_views(dy) if plan.n_views == 4 else [_view(dy)]
        dx = torch.empty(b, h, w, cin, dtype=torch.bfloat16, device=dev)
        ostr, ooff = _out_geometry(dx, plan.out_parity)
        _launch_tapgemm(plan, tile, views, gw, gh, b, packed, sched, n_rows, cin, n_tile, dx, ostr, ooff)
        outs.append(dx)
        n_off += cin
    return outs


class ZeroArena:
    """Pre-zeroed fp32 scratch for the backward pass: the packed weight-gradient accumulators (one per convolution, red.add
    targets) and the per-(sample, channel) sums of the block-tail backward each needed their own ``torch.zeros`` -- ~160 fill
    launches per step.  ``begin_step`` zeroes ONE arena sized by the previous step's demand; ``zeros`` hands out slices.
    Only for buffers that die inside the node that asked for them (nothing taken from here may be returned to autograd)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.used = 0          # elements handed out since the last begin_step
        self.clean = 0         # elements zeroed by the last begin_step (slices beyond it fall back to torch.zeros)
        self.demand = 0        # elements asked for since the last begin_step (sizes the next one)

    def begin_step(self, device):
        need = self.demand
        if need and (self.buf is None or self.buf.numel() < need or self.buf.device != device):
            self.buf = torch.empty(need + need // 8, dtype=torch.float32, device=device)
        self.clean = 0
        if self.buf is not None and self.buf.device == device and need:
            self.clean = min(need, self.buf.numel())
            self.buf[:self.clean].zero_()
        self.used, self.demand = 0, 0

    def zeros(self, n: int, device) -> torch.Tensor:
        n_al = (n + 63) // 64 * 64                        # 256-byte aligned slices (red.global.add.v4 targets)
        self.demand += n_al
        if self.buf is not None and self.buf.device == device and self.used + n_al <= self.clean:
            out = self.buf[self.used:self.used + n]
            self.used += n_al
            return out
        return torch.zeros(n, dtype=torch.float32, device=device)


ARENA = ZeroArena()


class Scratch:
    """Un-initialised fp32 scratch that dies inside the node that asked for it (the partial weight gradients of one
    convolution: written by ccdm_conv_wgrad, summed by ccdm_unpack_wgrad_slots right after, on the same stream).  ONE buffer
    per device, grown to the largest request -- stream order keeps consecutive users apart, and its address is stable once
    grown, which CUDA-graph capture needs (the warm-up steps before capture see every request size)."""

    def __init__(self):
        self.buf = {}

    def take(self, n: int, device) -> torch.Tensor:
        b = self.buf.get(device)
        if b is None or b.numel() < n:
            b = torch.empty(n + n // 8, dtype=torch.float32, device=device)
            self.buf[device] = b
        return b[:n]


SCRATCH = Scratch()


class PackCache:
    """Persistent packed bf16 copies of the convolution weights of the TRAINING path (forward layout and the transposed layout of
    the data gradient) plus the job table of ``ccdm_pack_multi``: ``begin_step`` re-packs every known weight in ONE launch
    (weights change once per optimizer step), so ``conv_forward`` / ``conv_dgrad`` launch their own pack only for a weight they
    see for the first time or whose ``_version`` moved since the batched launch."""

    def __init__(self):
        self.entries = {}                 # key -> [packed, weakref(weight), version stamp, job index]
        self.jobs = []                    # L.PackJob structs (host)
        self.table = None                 # device copy of the job table
        self.table_len = 0

    def get(self, weight, mode, kind, plan, n_rows, n_off, n_count, psched):
        """(packed tensor, True when the caller has to pack it now)."""
        key = (weight.data_ptr(), tuple(weight.shape), mode, plan.kind, plan.cins, plan.cout, plan.R, n_rows, n_off)
        ent = self.entries.get(key)
        if ent is not None and ent[1]() is weight:
            stale = ent[2] != weight._version
            ent[2] = weight._version
            return ent[0], stale
        with torch.inference_mode(False):
            packed = torch.empty(plan.nz * n_rows, plan.nkb * KB, dtype=torch.bfloat16, device=weight.device)
        job = L.PackJob()
        job.w, job.out, job.psched = weight.data_ptr(), packed.data_ptr(), psched.data_ptr()
        job.mode, job.cout, job.cin_total, job.ntaps = mode, weight.shape[0], weight.shape[1], _KIND_TAPS[kind]
        job.nkb, job.n_rows, job.n_off, job.n_count = plan.nkb, n_rows, n_off, n_count
        job.total = plan.nz * n_rows * plan.nkb * KB
        if ent is not None:               # same storage, another tensor object: rebind the slot
            self.jobs[ent[3]] = job
            idx = ent[3]
        else:
            self.jobs.append(job)
            idx = len(self.jobs) - 1
        self.entries[key] = [packed, weakref.ref(weight), weight._version, idx, psched]
        self.table = None                 # rebuilt by the next begin_step
        return packed, True

    def begin_step(self, device):
        """One ccdm_pack_multi launch over every known weight (none on the very first step: the layers register themselves)."""
        dead = [k for k, e in self.entries.items() if e[1]() is None]
        if dead:                          # a model went away: forget its weights (pointers in the table would dangle)
            self.entries, self.jobs, self.table = {}, [], None
            return
        if not self.jobs:
            return
        if self.table is None or self.table_len != len(self.jobs):
            raw = b"".join(bytes(j) for j in self.jobs)
            with torch.inference_mode(False):
                self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
            self.table_len = len(self.jobs)
        L.check(L.lib().ccdm_pack_multi(self.table.data_ptr(), len(self.jobs), _stream()), "pack_multi")
        for e in self.entries.values():
            w = e[1]()
            if w is not None:
                e[2] = w._version


PACKS = PackCache()


@functools.lru_cache(maxsize=None)
def _plan_cached_stem(cout: int) -> ConvPlan:
    return plan_conv("stem7", (64,), cout)


def _sm_count(dev) -> int:
    return torch.cuda.get_device_properties(dev).multi_processor_count if dev.type == "cuda" else 148


def wgrad_slots(plan: ConvPlan, tile, gw: int, gh: int, gb: int, cout: int, dev) -> int:
    """Position slices (split-K CTAs per unit) of one weight-gradient launch: about one wave of CTAs, as the library's own
    default -- computed here because in PARTIAL mode (ccdm_wgrad_args.slots) the caller owns one gradient slot per slice."""
    tw, th, tb = tile
    tiles_m = -(-gw // tw) * -(-gh // th) * -(-gb // tb)
    if plan.halo:
        n_tiles = -(-cout // 96)
        n_tile = (-(-cout // n_tiles) + 15) // 16 * 16
        units = plan.ngroups * -(-cout // n_tile)
    else:
        units = plan.nz * plan.ngroups * -(-cout // 128)
    return max(1, min(_sm_count(dev) // units, tiles_m))


def wgrad_packed(plan: ConvPlan, tile, views: List[L.View], dz: torch.Tensor, gw: int, gh: int, sched: torch.Tensor,
                 cout: int, n_rows: int, ksplit: int = 0, timing: Optional[list] = None, partial: bool = False):
    """fp32 weight gradient in the packed layout of the forward weights (ccdm_conv_wgrad).  ``partial`` = False: ONE
    [nz*n_rows, nkb*64] accumulator (pre-zeroed arena memory, red.global.add from every position slice).  ``partial`` = True:
    returns (buffer, nslots, slot_stride): every slice STORES its partial into its own slot (nothing to zero, no atomics)
    and ccdm_unpack_wgrad_slots sums them -- the red.add throughput of L2 bounded the split-K layers."""
    dev = dz.device
    size = plan.nz * n_rows * plan.nkb * KB
    slots = wgrad_slots(plan, tile, gw, gh, dz.shape[0], cout, dev) if partial else 0
    if partial:
        stride = (size + 63) // 64 * 64
        gpacked = SCRATCH.take(slots * stride, dev)
    else:
        gpacked = ARENA.zeros(size, dev).view(plan.nz * n_rows, plan.nkb * KB)
    a = L.WgradArgs()
    a.n_src = len(views)
    for i, v in enumerate(views):
        a.src[i] = v
    a.dz = dz.data_ptr()
    (a.dsW, a.dsH, a.dsB), ooff = _out_geometry(dz, plan.out_parity)
    for i in range(L.MAX_Z):
        a.doff[i] = ooff[i]
    a.gW, a.gH, a.gB = gw, gh, dz.shape[0]
    a.tw, a.th, a.tb = tile
    a.nz, a.ngroups, a.R = plan.nz, plan.ngroups, plan.R
    a.sched = sched.data_ptr()
    a.N, a.n_rows = cout, n_rows
    a.wgrad_packed = gpacked.data_ptr()
    a.ksplit = ksplit
    a.slots, a.slot_stride = (slots, stride) if partial else (0, 0)
    if timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(L.lib().ccdm_conv_wgrad(C.byref(a), _stream()), f"conv_wgrad[{plan.kind}]")
    if timing is not None:
        ev1.record()
        timing.append((ev0, ev1))
    return (gpacked, slots, stride) if partial else gpacked


def conv_wgrad(kind: str, srcs: Sequence[torch.Tensor], dz: torch.Tensor, ksplit: int = 0,
               timing: Optional[list] = None, accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW [Cout, sum(cins), kh, kw] fp32 of ``conv(kind)`` given its inputs and the gradient of its output.
    With ``accumulate_into`` (a contiguous fp32 tensor of that shape, e.g. ``weight.grad``) the result is added there."""
    for s in srcs:
        _check(s, "conv_wgrad source")
    _check(dz, "conv_wgrad dz")
    b, h, w, _ = srcs[0].shape
    _, oh, ow, cout = dz.shape
    gh, gw = (h, w) if kind == "up2x3x3" else (oh, ow)
    cins = [s.shape[3] for s in srcs]
    plan, tile = _wgrad_plan_for(kind, tuple(cins), cout, gw, gh)
    dev = dz.device
    n_rows = (cout + 31) // 32 * 32
    sched, psched = _dev_i32(plan.sched, dev), _dev_i32(plan.psched, dev)
    views: List[L.View] = []
    for s in srcs:
        views += _parity_views(s) if plan.n_views == 4 else [_view(s)]
    gpacked, nslots, stride = wgrad_packed(plan, tile, views, dz, gw, gh, sched, cout, n_rows, ksplit, timing, partial=True)
    k = int(math.isqrt(_KIND_TAPS[kind]))
    dw = accumulate_into if accumulate_into is not None else torch.empty(cout, sum(cins), k, k, dtype=torch.float32, device=dev)
    L.check(L.lib().ccdm_unpack_wgrad_slots(gpacked.data_ptr(), nslots, stride, dw.data_ptr(), cout, sum(cins),
                                            _KIND_TAPS[kind], psched.data_ptr(), plan.nz, plan.nkb, n_rows, None, 1.0,
                                            1 if accumulate_into is not None else 0, _stream()), "unpack_wgrad")
    return dw


def block_backward(dy: torch.Tensor, z: torch.Tensor, gain: torch.Tensor, scale_shift: Optional[torch.Tensor] = None,
                   ss_off: int = 0, silu: bool = True, dgain_into: Optional[torch.Tensor] = None,
                   dbias_into: Optional[torch.Tensor] = None):
    """Backward of ``silu(rmsnorm(z) * (1+scale) + shift)`` (Block.forward, unet.py:143-152).

    ``scale_shift`` is the fp32 [B, ld] buffer the forward read (scale at [ss_off, ss_off+C), shift right after).
    Returns (dz bf16, d_scale_shift fp32 [B, ld] or None, dgain [C], dbias [C]); ``dgain_into`` / ``dbias_into``
    (fp32, C elements, e.g. the parameters' ``.grad``) receive ``+=`` instead and None is returned in their place."""
    _check(dy, "block_backward dy")
    _check(z, "block_backward z")
    b, h, w, c = z.shape
    dev = z.device
    flags = (L.EPI_SILU if silu else 0) | (L.EPI_SS if scale_shift is not None else 0)
    dz = torch.empty_like(z)
    sums = ARENA.zeros(3 * b * c, dev)
    dgain = dbias = None
    if dgain_into is None or dbias_into is None:          # returned to autograd: never arena memory
        zbuf = torch.zeros(2 * c, dtype=torch.float32, device=dev)
        dgain, dbias = zbuf[:c], zbuf[c:]
    ld = scale_shift.shape[1] if scale_shift is not None else 0
    gm = math.sqrt(c)
    g = gain.reshape(-1)
    L.check(L.lib().ccdm_block_bwd(dy.data_ptr(), z.data_ptr(), dz.data_ptr(), b * h * w, c, h * w, g.data_ptr(), gm,
                                   L.ptr(scale_shift), ld, ss_off, sums.data_ptr(), flags, _stream()), "block_bwd")
    d_ss = None
    if scale_shift is not None:             # every element is written when the buffer is exactly [scale | shift]
        d_ss = torch.empty_like(scale_shift) if (ss_off == 0 and ld == 2 * c) else torch.zeros_like(scale_shift)
    L.check(L.lib().ccdm_block_bwd_finish(sums.data_ptr(), b, c, g.data_ptr(), gm, L.ptr(scale_shift), ld, ss_off,
                                          L.ptr(d_ss), (dgain_into if dgain_into is not None else dgain).data_ptr(),
                                          (dbias_into if dbias_into is not None else dbias).data_ptr(), _stream()),
            "block_bwd_finish")
    return dz, d_ss, (None if dgain_into is not None else dgain), (None if dbias_into is not None else dbias)
