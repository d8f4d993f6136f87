"""Execution engine of the B200 UNet: turns a :class:`ccdm_b200.unet.Unet` into a flat program of C-ABI calls.

The program is built once per (batch, resolution, mode): every intermediate lives in a persistent bf16 NHWC
buffer, every kernel argument struct is pre-filled, so a forward is a list of ``lib.fn(args, stream)`` calls with
no allocation and no host sync -- which is what makes the whole denoising step CUDA-graph capturable.

Layer -> kernel mapping (reference lines in CCDM_unified/models/unet.py):
  Block            (:143-152)  one tap-GEMM: conv3x3 + bias + RMSNorm + [(1+scale),shift] + SiLU
  ResnetBlock      (:167-187)  all tc_mlp Linear layers of the network as ONE row-GEMM, then block1, [res_conv],
                               block2 with the residual add fused; torch.cat inputs become dual-source K loops
  LinearAttention  (:202-216)  qkv tap-GEMM (PreNorm folded: g*sqrt(C) into the weights, 1/|x| per row; q softmax in
                               the epilogue) -> context kernel -> fold context into to_out -> tap-GEMM over q with
                               per-sample weights + bias + RMSNorm + residual
  Attention        (:228-240)  qkv tap-GEMM -> small softmax-attention kernel -> to_out tap-GEMM + residual
  Downsample       (:80-81)    4x4/s2 conv as 16 taps over four strided parity views (no im2col, no gather pass)
  Upsample         (:74-78)    nearest-2x folded into the conv: four output-parity 2x2 convs with summed taps
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .plan import ConvPlan, plan_conv, tile_box, can_reuse_rows, n_tiling, halo_ok, HALO_TILE, KB

_SYNC_EACH = os.environ.get("CCDM_SYNC_EACH") == "1"


# --------------------------------------------------------------------------------------------- records

@dataclass
class ViewRec:
    base: torch.Tensor          # bf16 storage the view lives in
    off: int                    # element offset of the first element
    C: int
    W: int
    H: int
    B: int
    sW: int
    sH: int
    sB: int


@dataclass
class PackRec:
    name: str
    weight: torch.Tensor        # fp32 [Cout, Cin_total, taps...] parameter (or [out, in] linear weight)
    plan: ConvPlan
    n_rows: int
    packed: torch.Tensor        # bf16 [nz*n_rows, nkb*64]
    sched: torch.Tensor         # int32 device [nz*nkb, 4]
    psched: torch.Tensor
    cin_gain: Optional[torch.Tensor] = None
    gain_mul: float = 1.0
    row_off: int = 0            # first packed row this weight occupies (tc_mlp layers share one matrix)
    rows_alloc: int = 0
    nkb_total: int = 0          # > 0: the packed matrix has this many K blocks per row and this weight starts at block kb0
    kb0: int = 0                #      (a conv and its block's shortcut share one matrix, CCDM_EPI_RESACC)


@dataclass
class TapGemmRec:
    name: str
    plan: ConvPlan
    views: List[ViewRec]
    gW: int
    gH: int
    gB: int
    tile: Tuple[int, int, int]
    pack: Optional[PackRec]             # None when the weights are produced on the fly (linattn fold)
    wpacked: torch.Tensor
    sched: torch.Tensor
    n_rows: int
    N: int
    n_tile: int
    flags: int
    out: torch.Tensor
    out_strides: Tuple[int, int, int]   # (sW, sH, sB) elements
    ooff: Tuple[int, int, int, int] = (0, 0, 0, 0)
    bias: Optional[torch.Tensor] = None
    rowss: Optional[torch.Tensor] = None
    gain: Optional[torch.Tensor] = None
    gain_mul: float = 1.0
    ss: Optional[torch.Tensor] = None
    ss_ld: int = 0
    ss_off: int = 0
    resid: Optional[torch.Tensor] = None
    resid_strides: Tuple[int, int, int] = (0, 0, 0)
    out_rowss: Optional[torch.Tensor] = None
    q_scale: float = 0.0
    q_cols: int = 0
    w_batch_rows: int = 0
    algo_flops: Optional[float] = None  # overrides tapgemm_flops (GEMMs that are not a reference convolution)
    head: Optional[tuple] = None        # CCDM_EPI_HEAD: (weight [k][N] fp32, bias [k], out fp32 NCHW [B][k][H][W])
    n_res: int = 0                      # CCDM_EPI_RESACC: shortcut load groups (sched rows after the main ones)
    res_sched: Optional[list] = None    # [(view index, 0, 0, c0)]
    res_bias: Optional[torch.Tensor] = None
    res_cin: int = 0                    # input channels of a real res_conv (0 for the identity shortcut): FLOP accounting


@dataclass
class KernelRec:
    """Any non-tap-GEMM call: ``kind`` names the C entry point (without the ccdm_ prefix), ``a`` its operands."""
    kind: str
    a: dict


# --------------------------------------------------------------------------------------------- helpers

def nhwc_view(t: torch.Tensor, c_off: int = 0, C: Optional[int] = None) -> ViewRec:
    b, h, w, ct = t.shape
    return ViewRec(t, c_off, ct - c_off if C is None else C, w, h, b, ct, w * ct, h * w * ct)


def parity_views(t: torch.Tensor) -> List[ViewRec]:
    """x[:, pr::2, pq::2, :] for (pr, pq) in row-major order -- the four planes of a stride-2 conv."""
    b, h, w, ct = t.shape
    out = []
    for pr in range(2):
        for pq in range(2):
            out.append(ViewRec(t, (pr * w + pq) * ct, ct, (w - pq + 1) // 2, (h - pr + 1) // 2, b,
                               2 * ct, 2 * w * ct, h * w * ct))
    return out


def fused_linattn_ok(training: bool, heads: int, C: int, n: int) -> bool:
    """Shapes the fused inference kernels of csrc/linattn_fused.cu cover: a pure function of the layer (never of the batch
    size), so the choice cannot break batch-shard invariance."""
    return (not training) and heads == 4 and C <= 128 and C % 8 == 0 and n >= 256 and n % 128 == 0


class Program:
    """A flat list of records plus their pre-filled ctypes argument structs."""

    def __init__(self, device, precision: str = "bf16"):
        self.device = device
        self.precision = precision          # which build of the library runs this program (ccdm_b200._lib.LIB_PATHS)
        self.recs: List[object] = []
        self.calls: List[tuple] = []        # (fn, args tuple without the stream)
        self.bufs: Dict[str, torch.Tensor] = {}
        self._keep: List[object] = []

    def buf(self, name, shape, dtype):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        assert name not in self.bufs, name
        self.bufs[name] = t
        return t

    def finalize(self):
        lib = L.lib(self.precision)
        self.calls = [_make_call(lib, r, self._keep) for r in self.recs]

    def run(self, stream: int):
        if _SYNC_EACH:                                     # debugging aid: attribute an asynchronous fault to its record
            for fn, args, name in self.calls:
                rc = fn(*args, stream)
                if rc != 0:
                    L.check(rc, name, self.precision)
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    raise RuntimeError(f"device fault in record '{name}': {e}") from e
            return
        for fn, args, name in self.calls:
            rc = fn(*args, stream)
            if rc != 0:
                L.check(rc, name, self.precision)

    def run_timed(self):
        """Eager run on torch's current stream with a CUDA event between launches.
        Returns [(record, milliseconds)] -- the per-kernel durations bench.py builds its roofline from."""
        stream = torch.cuda.current_stream().cuda_stream
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.calls) + 1)]
        evs[0].record()
        for i, (fn, args, name) in enumerate(self.calls):
            rc = fn(*args, stream)
            if rc != 0:
                L.check(rc, name, self.precision)
            evs[i + 1].record()
        torch.cuda.synchronize()
        return [(self.recs[i], evs[i].elapsed_time(evs[i + 1])) for i in range(len(self.calls))]


def tapgemm_flops(r: "TapGemmRec") -> float:
    """Algorithmic FLOPs (2*MAC) of the reference convolution this record computes:
    2 * B * Hout * Wout * Cout * Cin * kh * kw -- the nearest-2x fold and padded channels are NOT credited."""
    if r.algo_flops is not None:
        return r.algo_flops
    cin = sum(r.plan.cins)
    if r.plan.out_parity:
        pos = r.gB * (2 * r.gH) * (2 * r.gW)
    else:
        pos = r.gB * r.gH * r.gW
    return 2.0 * pos * r.N * (cin * r.plan.ntaps + r.res_cin)


KSTEPS_ENABLED = os.environ.get("CCDM_KSTEPS", "1") != "0"      # A/B switch: skip the empty K steps of partial channel blocks
_KSTEPS_DEV: Dict[tuple, tuple] = {}


def _ksteps_dev(table, device) -> torch.Tensor:
    """One device copy per (K-step table of a plan, device); the table object is kept alive with its copy."""
    key = (id(table), str(device))
    hit = _KSTEPS_DEV.get(key)
    if hit is None or hit[0] is not table:
        with torch.inference_mode(False):
            hit = (table, torch.tensor(table, dtype=torch.int32, device=device).contiguous())
        _KSTEPS_DEV[key] = hit
    return hit[1]


def _fill_tapgemm(r: TapGemmRec) -> L.TapGemmArgs:
    a = L.TapGemmArgs()
    a.n_src = len(r.views)
    for i, v in enumerate(r.views):
        a.src[i] = L.View(v.base.data_ptr() + 2 * v.off, v.C, v.W, v.H, v.B, v.sW, v.sH, v.sB)
    a.gW, a.gH, a.gB = r.gW, r.gH, r.gB
    a.tw, a.th, a.tb = r.tile
    a.nz, a.ngroups, a.R = r.plan.nz, r.plan.ngroups, r.plan.R
    a.halo = int(getattr(r.plan, "halo", False))
    ks = getattr(r.plan, "ksteps", None) if KSTEPS_ENABLED else None
    a.ksteps = _ksteps_dev(ks, r.sched.device).data_ptr() if ks is not None else None
    a.sched, a.wpacked = r.sched.data_ptr(), r.wpacked.data_ptr()
    a.n_rows, a.w_batch_rows, a.N, a.n_tile, a.flags = r.n_rows, r.w_batch_rows, r.N, r.n_tile, r.flags
    a.bias, a.rowss, a.gain, a.gain_mul = L.ptr(r.bias), L.ptr(r.rowss), L.ptr(r.gain), r.gain_mul
    a.scale_shift, a.ss_ld, a.ss_off = L.ptr(r.ss), r.ss_ld, r.ss_off
    a.resid = L.ptr(r.resid)
    a.rsW, a.rsH, a.rsB = r.resid_strides
    a.out = L.ptr(r.out)
    a.osW, a.osH, a.osB = r.out_strides
    if r.head is not None:
        hw_, hb_, ho_ = r.head
        a.head_n, a.head_w, a.head_b, a.head_out = hw_.shape[0], hw_.data_ptr(), hb_.data_ptr(), ho_.data_ptr()
        a.hsC, a.hsB = ho_.shape[2] * ho_.shape[3], ho_.shape[1] * ho_.shape[2] * ho_.shape[3]
    for i in range(L.MAX_Z):
        a.ooff[i] = r.ooff[i]
    a.out_rowss, a.q_scale, a.q_cols = L.ptr(r.out_rowss), r.q_scale, r.q_cols
    a.n_res, a.res_bias = r.n_res, L.ptr(r.res_bias)
    return a


def _make_call(lib, r, keep):
    import ctypes as C
    if isinstance(r, TapGemmRec):
        args = _fill_tapgemm(r)
        keep.append(args)
        return lib.ccdm_tapgemm, (C.byref(args),), r.name
    if isinstance(r, PackRec):
        w = r.weight
        cout, cin_total = w.shape[0], w.shape[1]
        ntaps = w.numel() // (cout * cin_total)
        dst = r.packed.data_ptr() + 2 * r.row_off * (r.nkb_total or r.plan.nkb) * KB
        if r.nkb_total:
            return lib.ccdm_pack_weights_at, (w.data_ptr(), cout, cin_total, ntaps, r.psched.data_ptr(), r.plan.nz,
                                              r.plan.nkb, r.n_rows, L.ptr(r.cin_gain), r.gain_mul, dst, r.nkb_total,
                                              r.kb0), "pack:" + r.name
        return lib.ccdm_pack_weights, (w.data_ptr(), cout, cin_total, ntaps, r.psched.data_ptr(), r.plan.nz, r.plan.nkb,
                                       r.n_rows, L.ptr(r.cin_gain), r.gain_mul, dst), "pack:" + r.name
    k, a = r.kind, r.a
    p = L.ptr
    if k == "stem_im2row":
        return lib.ccdm_stem_im2row, (p(a["x"]), p(a["out"]), a["B"], a["Cin"], a["H"], a["W"]), k
    if k == "stem_pack":
        return lib.ccdm_stem_pack, (p(a["w"]), p(a["wpacked"]), a["Cout"], a["Cin"], a["n_rows"]), k
    if k == "rmsnorm_act":
        return lib.ccdm_rmsnorm_act, (p(a["z"]), p(a["out"]), a["rows"], a["C"], a["rows_per_sample"], p(a["gain"]),
                                      a["gain_mul"], p(a.get("ss")), a.get("ss_ld", 0), a.get("ss_off", 0),
                                      p(a.get("resid")), p(a.get("out_rowss")), a["flags"]), k
    if k == "head_conv1":
        return lib.ccdm_head_conv1, (p(a["x"]), p(a["w"]), p(a["bias"]), p(a["out"]), a["B"], a["H"], a["W"], a["Cin"],
                                     a["Cout"]), k
    if k == "kexp_bound":
        return lib.ccdm_kexp_bound, (p(a["wpacked"]), a["n_rows"], a["K"], a["lo"], a["hi"], p(a["bias"])), k
    if k == "linattn_context":
        return lib.ccdm_linattn_context, (p(a["qkv"]), p(a["ctx"]), None, a["B"], a["n"], a["heads"], p(a["w_out"]),
                                          p(a["wfold"]), a["C"], a["n_rows"]), k
    if k == "linattn_pack_blockdiag":
        return lib.ccdm_linattn_pack_blockdiag, (p(a["m"]), p(a.get("row_div")), a["transpose"], p(a["w"]), a["B"]), k
    if k == "linattn_fold":
        return lib.ccdm_linattn_fold, (p(a["w_out"]), p(a["ctx"]), p(a["wfold"]), a["B"], a["C"], a["n_rows"],
                                       a["heads"]), k
    if k == "linattn_kv_partials":
        return lib.ccdm_linattn_kv_partials, (p(a["x"]), a["B"], a["n"], a["C"], p(a["rowss"]), p(a["wqkv"]), p(a["kbias"]),
                                              p(a["part"]), p(a["psum"])), k
    if k == "linattn_fold_partials":
        return lib.ccdm_linattn_fold_partials, (p(a["part"]), p(a["psum"]), a["B"], a["ups"], p(a["w_out"]), a["C"],
                                                a["n_rows"], p(a["wfold"])), k
    if k == "linattn_q_out":
        return lib.ccdm_linattn_q_out, (p(a["x"]), a["B"], a["n"], a["C"], p(a["rowss"]), p(a["wqkv"]), p(a["wfold"]),
                                        a["n_rows"], p(a["bias"]), p(a["gain"]), a["gain_mul"], a["q_scale"], p(a["out"])), k
    if k == "attention_small":
        return lib.ccdm_attention_small, (p(a["qkv"]), p(a["out"]), a["B"], a["n"], a["heads"], a["dim_head"],
                                          a["scale"]), k
    if k == "linear_small":
        bn = a.get("bn")
        return lib.ccdm_linear_small, (p(a["x"]), a["B"], a["in_dim"], p(a["w"]), p(a["bias"]), a["out_dim"],
                                       p(bn.weight) if bn else None, p(bn.bias) if bn else None,
                                       p(bn.running_mean) if bn else None, p(bn.running_var) if bn else None,
                                       int(a.get("bn_train", 0)), a["act"], p(a["y"]), a["out_dim"]), k
    if k == "time_features":
        return lib.ccdm_time_features, (p(a["t"]), a["B"], a["dim"], p(a["out"])), k
    if k == "select_null":
        return lib.ccdm_select_null, (p(a["c"]), p(a["keep"]), 0, p(a["null_emb"]), a["B"], a["dim"]), k
    if k == "silu_concat_bf16":
        return lib.ccdm_silu_concat_bf16, (p(a["t_emb"]), a["dt"], p(a["c_emb"]), a["dc"], a["B"], p(a["out"])), k
    # ---- vanilla (GroupNorm) UNet, ccdm_b200/vanilla_unet.py
    if k == "channel_stats":
        return lib.ccdm_channel_stats, (p(a["x"]), a["B"], a["rows"], a["C"], p(a["sums"]), a["ld"], a["c_off"],
                                        a["zero_first"]), k
    if k == "groupnorm_coef":
        return lib.ccdm_groupnorm_coef, (p(a["sums"]), a["B"], a["Ctot"], a["groups"], a["rows"], a["eps"], p(a["gamma"]),
                                         p(a["beta"]), p(a.get("ss")), a.get("ss_ld", 0), a.get("ss_off", 0), a["C0"],
                                         p(a["coef"])), k
    if k == "affine_act":
        return lib.ccdm_affine_act, (p(a["x"]), p(a["out"]), a["rows"], a["C"], a["rows_per_sample"], p(a["ss"]),
                                     a["ss_ld"], a["ss_off"], a["act"]), k
    if k == "attention_tokens":
        return lib.ccdm_attention_tokens, (p(a["qkv"]), p(a["out"]), a["B"], a["n"], a["heads"], a["dim_head"],
                                           a["scale"], a["head_major"]), k
    if k == "time_features_adm":
        return lib.ccdm_time_features_adm, (p(a["t"]), a["B"], a["dim"], a["max_period"], p(a["out"])), k
    raise ValueError(k)


# --------------------------------------------------------------------------------------------- weights

class WeightStore:
    """bf16 K-blocked copies of the conv / linear weights, re-derived whenever a parameter changes.

    The fp32 ``nn.Parameter`` tensors stay the source of truth (optimizer, EMA ``lerp_``, ``load_state_dict`` all
    write them in place and bump ``_version``), so the packed copies are a cache keyed on those versions."""

    def __init__(self, device, precision: str = "bf16"):
        self.device = device
        self.packs: Dict[str, PackRec] = {}
        self.program = Program(device, precision)
        self._stamp = None
        self._params: List[torch.Tensor] = []

    def add(self, name, weight, plan: ConvPlan, n_rows, cin_gain=None, gain_mul=1.0, shared=None, row_off=0,
            nkb_total=0, kb0=0) -> PackRec:
        if name in self.packs:
            return self.packs[name]
        dev = self.device
        if shared is None:
            packed = torch.zeros(plan.nz * n_rows, (nkb_total or plan.nkb) * KB, dtype=torch.bfloat16, device=dev)
        else:
            packed = shared
        sched = torch.tensor(plan.sched, dtype=torch.int32, device=dev).contiguous()
        psched = torch.tensor(plan.psched, dtype=torch.int32, device=dev).contiguous()
        rec = PackRec(name, weight, plan, n_rows, packed, sched, psched, cin_gain, gain_mul, row_off, nkb_total=nkb_total,
                      kb0=kb0)
        self.packs[name] = rec
        self.program.recs.append(rec)
        self._params.append(weight)
        if cin_gain is not None:
            self._params.append(cin_gain)
        return rec

    def identity_weight(self, c: int) -> torch.Tensor:
        """[c, c, 1, 1] identity "1x1 conv" (the shortcut of a ResnetBlock whose channel count does not change, unet.py:165)."""
        key = f"eye:{c}"
        if key not in self.packs:
            self.packs[key] = torch.eye(c, dtype=torch.float32, device=self.device).view(c, c, 1, 1).contiguous()
        return self.packs[key]

    def kexp_bias(self, pack: PackRec, lo: int, hi: int) -> torch.Tensor:
        """Bias vector holding the softmax shift of the k columns of a to_qkv weight (refreshed with the pack)."""
        key = "kexp:" + pack.name
        if key not in self.packs:
            bias = torch.zeros(pack.n_rows, dtype=torch.float32, device=self.device)
            rec = KernelRec("kexp_bound", dict(wpacked=pack.packed, n_rows=pack.n_rows, K=pack.plan.nkb * KB, lo=lo,
                                               hi=hi, bias=bias))
            self.packs[key] = rec
            self.program.recs.append(rec)
        return self.packs[key].a["bias"]

    def stem_pack(self, weight: torch.Tensor, n_rows: int) -> torch.Tensor:
        """bf16 [n_rows][256] packing of the 7x7 stem weight for the im2row formulation (ccdm_stem_pack)."""
        key = "stem_pack"
        if key not in self.packs:
            packed = torch.zeros(n_rows, 256, dtype=torch.bfloat16, device=self.device)
            rec = KernelRec("stem_pack", dict(w=weight, wpacked=packed, Cout=weight.shape[0], Cin=weight.shape[1],
                                              n_rows=n_rows))
            self.packs[key] = rec
            self.program.recs.append(rec)
            self._params.append(weight)
        return self.packs[key].a["wpacked"]

    def stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self._params)

    def refresh(self, stream: int):
        st = self.stamp()
        if st != self._stamp:
            if len(self.program.calls) != len(self.program.recs):
                self.program.finalize()
            self.program.run(stream)
            self._stamp = st


# --------------------------------------------------------------------------------------------- program builder

RESACC_ENABLED = os.environ.get("CCDM_RESACC", "1") != "0"      # A/B switch for profiles; the fused shortcut is the default
HALO_ENABLED = os.environ.get("CCDM_HALO", "1") != "0"          # A/B switch: one halo box per 3x3 tile instead of three boxes


class UnetEngine:
    def __init__(self, net):
        self.net = net
        self.device = net.init_conv.weight.device
        self.precision = getattr(net, "precision", "bf16")
        self.weights = WeightStore(self.device, self.precision)
        self.programs: Dict[tuple, "UnetProgram"] = {}
        self._ptr_stamp = None

    def _param_ptrs(self):
        return tuple(p.data_ptr() for p in self.net.parameters())

    def program(self, B, x_batch, H, W, training) -> "UnetProgram":
        ptrs = self._param_ptrs()
        if ptrs != self._ptr_stamp:        # parameters were re-allocated (.to(), load with assign): rebuild all
            self.programs.clear()
            self.weights = WeightStore(self.device, self.precision)
            self._ptr_stamp = ptrs
        key = (B, x_batch, H, W, bool(training))
        prog = self.programs.get(key)
        if prog is None:
            with torch.inference_mode(False):      # persistent buffers must be normal tensors (reused outside)
                prog = UnetProgram(self.net, self.weights, B, x_batch, H, W, training)
            self.programs[key] = prog
        return prog

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _run(self, prog, x, t, emb_rows, keep_rows):
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, model on {self.device}")
        prog.x_in.copy_(x)
        prog.t_in.copy_(t.reshape(-1))
        prog.emb_in.copy_(emb_rows)
        prog.keep.copy_(keep_rows)
        stream = self._stream()
        self.weights.refresh(stream)
        prog.run(stream)
        if self.net.training:
            for bn in (self.net.cond_mlp_1[1], self.net.cond_mlp_2[1]):
                bn.num_batches_tracked += 1
        return prog.out

    def forward(self, x, t, labels_emb, keep_mask):
        """One UNet evaluation.  keep_mask: bool [B] (True = conditional row) or None for all-conditional."""
        B, _, H, W = x.shape
        prog = self.program(B, B, H, W, self.net.training)
        keep = torch.ones(B, dtype=torch.uint8, device=self.device) if keep_mask is None else keep_mask.to(torch.uint8)
        return self._run(prog, x, t, labels_emb, keep).clone()

    def forward_pair(self, x, t, labels_emb):
        """Conditional and unconditional evaluations as one 2B batch (eval mode only).  Returns views of the
        persistent output buffer: (cond [B,...], null [B,...])."""
        B, _, H, W = x.shape
        prog = self.program(2 * B, B, H, W, False)
        if prog.pair_keep is None:
            with torch.inference_mode(False):
                prog.pair_keep = torch.cat([torch.ones(B, dtype=torch.uint8),
                                            torch.zeros(B, dtype=torch.uint8)]).to(self.device)
        t2 = torch.cat([t.reshape(-1), t.reshape(-1)])
        emb2 = torch.cat([labels_emb, labels_emb])
        out = self._run(prog, x, t2, emb2, prog.pair_keep)
        return out[:B], out[B:]

    def cfg_combine(self, cond, null, cond_scale, rescaled_phi, remove_parallel=True, keep_parallel_frac=0.0):
        B = cond.shape[0]
        chw = cond[0].numel()
        out = torch.empty_like(cond)
        L.check(L.lib().ccdm_cfg_combine(cond.data_ptr(), null.data_ptr(), out.data_ptr(), B, chw, float(cond_scale),
                                         float(rescaled_phi), int(bool(remove_parallel)), float(keep_parallel_frac),
                                         self._stream()), "cfg_combine")
        return out


class UnetProgram(Program):
    def __init__(self, net, weights: WeightStore, B, x_batch, H, W, training):
        super().__init__(net.init_conv.weight.device, getattr(weights.program, "precision", "bf16"))
        self.net, self.weights = net, weights
        self.B, self.x_batch, self.H, self.W, self.training = B, x_batch, H, W, training
        self.pair_keep = None
        self._uid = 0
        self._build()
        self.finalize()

    # ------------------------------------------------------------------ small builders
    def act(self, name, h, w, c):
        return self.buf(name, (self.B, h, w, c), torch.bfloat16)

    def kernel(self, kind, **a):
        self.recs.append(KernelRec(kind, a))

    def conv(self, name, kind, srcs: List[torch.Tensor], conv_mod, out: torch.Tensor, flags=0, *, gain=None,
             ss_off=None, resid=None, out_rowss=None, rowss=None, cin_gain=None, cin_gain_mul=1.0, q=None,
             views: Optional[List[ViewRec]] = None, head=None, res=None):
        """Append one tap-GEMM over NHWC sources.  ``conv_mod`` owns .weight / .bias (nn.Conv2d).  ``head`` = (1x1 conv module,
        fp32 NCHW output): the tile is not stored but projected by the 1x1 conv in the epilogue (CCDM_EPI_HEAD); ``out`` is then
        only the (h, w, channels) geometry."""
        cins = [s.shape[3] for s in srcs] if views is None else [v.C for v in views]
        cout = conv_mod.weight.shape[0]
        out_t = out
        if head is not None:
            out_t, out = None, torch.empty((self.B,) + tuple(out), device="meta")
            flags |= L.EPI_HEAD
        gh, gw = out.shape[1], out.shape[2]
        if kind == "up2x3x3":
            gh, gw = gh // 2, gw // 2
        tile = tile_box(gw, gh, square=(kind != "1x1"))
        reuse = kind != "1x1" and can_reuse_rows(tile)
        plan = plan_conv(kind, cins, cout, reuse_rows=reuse)
        # 3x3 layers whose weights stay resident in shared memory read ALL nine taps out of one halo box per 64-channel block
        # (a third of the TMA fills and L2 -> SM bytes of the three shifted boxes): the <= 64-wide layers of the 64x64 / 32x32
        # levels, i.e. exactly the ones bound by the shared-memory port
        if HALO_ENABLED and views is None and halo_ok(kind, gw, gh) and cout <= 128:
            n_res_kb = sum(-(-t.shape[3] // KB) for t in res[0]) if res is not None else 0
            nkb_tot = 9 * sum(-(-c // KB) for c in cins) + n_res_kb
            pad = (cout + 31) // 32 * 32
            resident = nkb_tot * pad * 128 // (2 if nkb_tot >= 16 else 1)
            # (measured: with >= 2 shortcut groups AND an output staging buffer the 23 KB stages leave too little lookahead --
            # ups.4.x block2 202 -> 210-224 us -- while the head-fused final block, which stages nothing, gains 12 us)
            crowded = n_res_kb >= 2 and head is None
            if resident <= 96 * 1024 and not crowded:
                plan, tile = plan_conv(kind, cins, cout, halo=True), HALO_TILE
        full_row = bool(flags & (L.EPI_RMSNORM | L.EPI_SUMSQ_OUT))
        # A fused channel norm needs the whole row in one CTA (<= 512 TMEM columns).  Wider layers (dim-72 models:
        # 576) and layers with too few pixel tiles to fill the GPU (4x4 / 8x8 levels) run the GEMM split over output
        # channels with a plain bias epilogue and finish with the standalone norm kernel.
        m_tiles = ((gw + tile[0] - 1) // tile[0]) * ((gh + tile[1] - 1) // tile[1]) * ((self.B + tile[2] - 1) // tile[2])
        split = full_row and (cout > 512 or (cout >= 256 and m_tiles < 100))
        if split:
            return self._conv_split(name, kind, plan, tile, srcs, views, conv_mod, out, flags, gain, ss_off, resid,
                                    out_rowss, gw, gh)
        n_rows, n_tile = n_tiling(cout, full_row, plan.nkb >= 16)
        if views is None:
            views = []
            for s in srcs:
                views += parity_views(s) if plan.n_views == 4 else [nhwc_view(s)]
        sched_t, n_res, res_sched, res_bias, res_cin = None, 0, None, None, 0
        if res is not None:
            # the block's shortcut (res_conv over the block's inputs, or the identity) as extra 1x1 load groups of THIS
            # launch, accumulated in a second TMEM accumulator (CCDM_EPI_RESACC): no res_conv launch, no residual tensor
            res_srcs, res_mod = res
            rplan = plan_conv("1x1", [t.shape[3] for t in res_srcs], cout)
            n_res, nkb_total = rplan.ngroups, plan.nkb + rplan.ngroups
            pack = self.weights.add(f"{name}/R{plan.R}+res", conv_mod.weight, plan, n_rows, cin_gain, cin_gain_mul,
                                    nkb_total=nkb_total)
            rw = res_mod.weight if res_mod is not None else self.weights.identity_weight(cout)
            self.weights.add(f"{name}/res", rw, rplan, n_rows, shared=pack.packed, nkb_total=nkb_total, kb0=plan.nkb)
            res_sched = [(len(views) + src, 0, 0, c0) for (src, _, _, c0) in rplan.sched]
            views = views + [nhwc_view(t) for t in res_srcs]
            assert len(views) <= L.MAX_SRC
            sched_t = torch.tensor(list(plan.sched) + res_sched, dtype=torch.int32, device=self.device).contiguous()
            res_bias = res_mod.bias if res_mod is not None else None
            res_cin = sum(rplan.cins) if res_mod is not None else 0
            flags |= L.EPI_RESACC
        else:
            pack = self.weights.add(f"{name}/R{plan.R}", conv_mod.weight, plan, n_rows, cin_gain, cin_gain_mul)
        co = out.shape[3]
        if plan.out_parity:
            ostr = (2 * co, 2 * out.shape[2] * co, out.shape[1] * out.shape[2] * co)
            ooff = tuple((pa * out.shape[2] + pb) * co for pa in range(2) for pb in range(2))
        else:
            ostr = (co, out.shape[2] * co, out.shape[1] * out.shape[2] * co)
            ooff = (0, 0, 0, 0)
        if conv_mod.bias is not None:
            flags |= L.EPI_BIAS
        rec = TapGemmRec(name, plan, views, gw, gh, self.B, tile, pack, pack.packed,
                         pack.sched if sched_t is None else sched_t, n_rows,
                         cout, n_tile, flags, out_t, ostr, ooff, bias=conv_mod.bias, rowss=rowss, gain=gain,
                         gain_mul=math.sqrt(cout) if gain is not None else 1.0, out_rowss=out_rowss, n_res=n_res,
                         res_sched=res_sched, res_bias=res_bias, res_cin=res_cin)
        if head is not None:
            hc, hout = head
            rec.head = (hc.weight.view(hc.weight.shape[0], -1), hc.bias, hout)
        if ss_off is not None:
            rec.ss, rec.ss_ld, rec.ss_off = self.bufs["ss_all"], self.bufs["ss_all"].shape[1], ss_off
        if resid is not None:
            rc = resid.shape[3]
            rec.resid, rec.resid_strides = resid, (rc, resid.shape[2] * rc, resid.shape[1] * resid.shape[2] * rc)
        if q is not None:
            rec.q_scale, rec.q_cols = q
        self.recs.append(rec)
        return rec

    def _conv_split(self, name, kind, plan, tile, srcs, views, conv_mod, out, flags, gain, ss_off, resid, out_rowss,
                    gw, gh):
        """conv + bias as a channel-split tap-GEMM into a bf16 scratch, then ccdm_rmsnorm_act for the tail."""
        cout = conv_mod.weight.shape[0]
        n_rows, n_tile = n_tiling(cout, False, plan.nkb >= 16)
        pack = self.weights.add(f"{name}/R{plan.R}/split", conv_mod.weight, plan, n_rows)
        if views is None:
            views = []
            for s_ in srcs:
                views += parity_views(s_) if plan.n_views == 4 else [nhwc_view(s_)]
        z = self.buf(name + ".z", tuple(out.shape), torch.bfloat16)
        ostr = (cout, out.shape[2] * cout, out.shape[1] * out.shape[2] * cout)
        self.recs.append(TapGemmRec(name, plan, views, gw, gh, self.B, tile, pack, pack.packed, pack.sched, n_rows,
                                    cout, n_tile, L.EPI_BIAS if conv_mod.bias is not None else 0, z, ostr,
                                    bias=conv_mod.bias))
        a = dict(z=z, out=out, rows=self.B * gh * gw, C=cout, rows_per_sample=gh * gw, gain=gain,
                 gain_mul=math.sqrt(cout), flags=flags & (L.EPI_SS | L.EPI_SILU | L.EPI_RESID | L.EPI_SUMSQ_OUT))
        if ss_off is not None:
            a.update(ss=self.bufs["ss_all"], ss_ld=self.bufs["ss_all"].shape[1], ss_off=ss_off)
        if resid is not None:
            a["resid"] = resid
        if out_rowss is not None:
            a["out_rowss"] = out_rowss
        self.kernel("rmsnorm_act", **a)
        return None

    # ------------------------------------------------------------------ network pieces
    def resblock(self, name, mod, srcs, h, w, ss_off, want_rowss=False, head=None):
        cout = mod.dim_out
        h1 = self.act(name + ".h1", h, w, cout)
        self.conv(name + ".block1.proj", "3x3", srcs, mod.block1.proj, h1,
                  L.EPI_RMSNORM | L.EPI_SS | L.EPI_SILU, gain=mod.block1.norm.g, ss_off=ss_off)
        has_conv = isinstance(mod.res_conv, torch.nn.Conv2d)
        # shortcut inside block 2's launch when its channels fit one <= 128-wide tile with the fused norm (second TMEM
        # accumulator); wider / split layers keep the separate res_conv launch and the residual read in the epilogue
        tile = tile_box(w, h, square=True)
        m_tiles = ((w + tile[0] - 1) // tile[0]) * ((h + tile[1] - 1) // tile[1]) * ((self.B + tile[2] - 1) // tile[2])
        fuse_res = (RESACC_ENABLED and cout <= 128 and not (cout >= 256 and m_tiles < 100) and
                    len(srcs) + 1 <= L.MAX_SRC and sum(-(-t.shape[3] // KB) for t in srcs) <= 8)
        res_arg = None
        if fuse_res:
            res, res_arg = None, (srcs, mod.res_conv if has_conv else None)
        elif has_conv:
            res = self.act(name + ".res", h, w, cout)
            self.conv(name + ".res_conv", "1x1", srcs, mod.res_conv, res)
        else:
            assert len(srcs) == 1
            res = srcs[0]
        # head: the block's output feeds only the 1x1 final_conv, which then runs in block 2's epilogue (no 64-channel tensor)
        out = self.act(name + ".out", h, w, cout) if head is None else (h, w, cout)
        rowss = self.buf(name + ".rowss", (self.B * h * w,), torch.float32) if want_rowss else None
        self.conv(name + ".block2.proj", "3x3", [h1], mod.block2.proj, out,
                  L.EPI_RMSNORM | L.EPI_SILU | (0 if fuse_res else L.EPI_RESID) | (L.EPI_SUMSQ_OUT if want_rowss else 0),
                  gain=mod.block2.norm.g, resid=res, out_rowss=rowss, head=head, res=res_arg)
        return out, rowss

    def linear_attention(self, name, mod, x, rowss, h, w):
        """mod = Residual(PreNorm(LinearAttention)); x NHWC [B,h,w,C] with its per-pixel sum of squares."""
        pre, att = mod.fn, mod.fn.fn
        C = x.shape[3]
        heads, hid = att.heads, att.heads * att.dim_head
        assert att.dim_head == 32, "linear attention kernels are written for dim_head == 32 (unet.py:190)"
        n = h * w
        if fused_linattn_ok(self.training, heads, C, n):
            return self._linear_attention_fused(name, pre, att, x, rowss, h, w)
        qkv = self.act(name + ".qkv", h, w, 3 * hid)
        rec = self.conv(name + ".to_qkv", "1x1", [x], att.to_qkv, qkv, L.EPI_ROWSCALE | L.EPI_QSOFTMAX | L.EPI_KEXP,
                        rowss=rowss, cin_gain=pre.norm.g, cin_gain_mul=math.sqrt(C), q=(att.scale, hid))
        # k columns: p = exp(k - bound), bound = ||W'_d|| from the packed weights (no max pass over the tokens)
        rec.bias = self.weights.kexp_bias(rec.pack, hid, 2 * hid)
        rec.flags |= L.EPI_BIAS
        ctx = self.buf(name + ".ctx", (self.B, heads, 32, 32), torch.float32)
        conv_out, norm_out = att.to_out[0], att.to_out[1]
        wide = C > 512                                      # norm cannot be fused: split channels + standalone norm
        n_rows, n_tile = n_tiling(C, not wide)
        wfold = self.buf(name + ".wfold", (self.B * n_rows, hid), torch.bfloat16)
        wfold.zero_()                                       # rows [C, n_rows) are padding and must stay zero
        # context, then its fold into to_out's weights: fused into the context kernel's epilogue while that is cheap
        # (C <= 64), a separate whole-GPU launch for the wider levels
        fuse = C <= 64 and heads == 4
        self.kernel("linattn_context", qkv=qkv, ctx=ctx, B=self.B, n=n, heads=heads, w_out=conv_out.weight,
                    wfold=wfold if fuse else None, C=C, n_rows=n_rows)
        if not fuse and heads == 4:
            # fold on tensor cores: wfold[b][c][hd] = sum_k Wout[c][k] * blockdiag(ctx_b)[hd][k] is a tap-GEMM whose
            # "image" is the (packed, bf16) to_out weight itself -- C pixels x 128 channels, the same for every sample
            # (batch stride 0) -- and whose per-sample weights are the block-diagonal contexts
            wb = self.buf(name + ".ctx_bd", (self.B * 128, 128), torch.bfloat16)
            wb.zero_()                                      # off-diagonal blocks stay zero: only the diagonal is rewritten
            self.kernel("linattn_pack_blockdiag", m=ctx, row_div=None, transpose=1 | 2, w=wb, B=self.B)
            fplan = plan_conv("1x1", [hid], 128)
            wpack = self.weights.add(name + ".to_out.w/R1", conv_out.weight, plan_conv("1x1", [hid], C),
                                     n_tiling(C, False)[0])
            wview = ViewRec(wpack.packed, 0, hid, C, 1, self.B, hid, C * hid, 0)
            fsched = torch.tensor(fplan.sched, dtype=torch.int32, device=self.device)
            self.recs.append(TapGemmRec(name + ".fold", fplan, [wview], C, 1, self.B, tile_box(C, 1, force_tb1=True),
                                        None, wb, fsched, 128, 128, 128, 0, wfold, (hid, C * hid, n_rows * hid),
                                        w_batch_rows=128, algo_flops=2.0 * self.B * C * hid * 32))
        elif not fuse:
            self.kernel("linattn_fold", w_out=conv_out.weight, ctx=ctx, wfold=wfold, B=self.B, C=C, n_rows=n_rows,
                        heads=heads)
        out = self.act(name + ".out", h, w, C)
        plan = plan_conv("1x1", [hid], C)
        sched = torch.tensor(plan.sched, dtype=torch.int32, device=self.device)
        qview = ViewRec(qkv, 0, hid, w, h, self.B, 3 * hid, w * 3 * hid, h * w * 3 * hid)
        dst = self.act(name + ".z", h, w, C) if wide else out
        flags = L.EPI_BIAS if wide else (L.EPI_BIAS | L.EPI_RMSNORM | L.EPI_RESID)
        rec = TapGemmRec(name + ".to_out", plan, [qview], w, h, self.B, tile_box(w, h, force_tb1=True), None, wfold,
                         sched, n_rows, C, n_tile, flags, dst, (C, w * C, h * w * C), bias=conv_out.bias,
                         w_batch_rows=n_rows)
        if not wide:
            rec.gain, rec.gain_mul = norm_out.g, math.sqrt(C)
            rec.resid, rec.resid_strides = x, (C, w * C, h * w * C)
        self.recs.append(rec)
        if wide:
            self.kernel("rmsnorm_act", z=dst, out=out, rows=self.B * h * w, C=C, rows_per_sample=h * w,
                        gain=norm_out.g, gain_mul=math.sqrt(C), resid=x, flags=L.EPI_RESID)
        return out

    def _linear_attention_fused(self, name, pre, att, x, rowss, h, w):
        """Inference path for C <= 128 and n % 128 == 0 (csrc/linattn_fused.cu): x is read twice, q | k | v never reach HBM.
        kv-kernel -> per-unit context partials -> fold into to_out's weight -> q-kernel with the output projection, the
        RMSNorm and the residual in its epilogue."""
        C, n, hid = x.shape[3], h * w, att.heads * att.dim_head
        plan = plan_conv("1x1", [C], 3 * hid)
        n_rows_w, _ = n_tiling(3 * hid, False)
        pack = self.weights.add(f"{name}.to_qkv/R{plan.R}", att.to_qkv.weight, plan, n_rows_w, pre.norm.g, math.sqrt(C))
        kbias = self.weights.kexp_bias(pack, hid, 2 * hid)
        ups = L.lib().ccdm_linattn_fused_units(n)
        assert ups > 0
        part = self.buf(name + ".ctx_part", (self.B * ups, 128, 32), torch.float32)
        psum = self.buf(name + ".ctx_psum", (self.B * ups, 128), torch.float32)
        conv_out, norm_out = att.to_out[0], att.to_out[1]
        n_rows, _ = n_tiling(C, True)
        wfold = self.buf(name + ".wfold", (self.B * n_rows, hid), torch.bfloat16)
        wfold.zero_()                                       # rows [C, n_rows) are padding and must stay zero
        out = self.act(name + ".out", h, w, C)
        self.kernel("linattn_kv_partials", x=x, B=self.B, n=n, C=C, rowss=rowss, wqkv=pack.packed, kbias=kbias, part=part,
                    psum=psum)
        self.kernel("linattn_fold_partials", part=part, psum=psum, B=self.B, ups=ups, w_out=conv_out.weight, C=C,
                    n_rows=n_rows, wfold=wfold)
        self.kernel("linattn_q_out", x=x, B=self.B, n=n, C=C, rowss=rowss, wqkv=pack.packed, wfold=wfold, n_rows=n_rows,
                    bias=conv_out.bias, gain=norm_out.g, gain_mul=math.sqrt(C), q_scale=att.scale, out=out)
        return out

    def mid_attention(self, name, mod, x, rowss, h, w):
        pre, att = mod.fn, mod.fn.fn
        C = x.shape[3]
        hid = att.heads * att.dim_head
        qkv = self.act(name + ".qkv", h, w, 3 * hid)
        self.conv(name + ".to_qkv", "1x1", [x], att.to_qkv, qkv, L.EPI_ROWSCALE, rowss=rowss, cin_gain=pre.norm.g,
                  cin_gain_mul=math.sqrt(C))
        ao = self.act(name + ".attn", h, w, hid)
        self.kernel("attention_small", qkv=qkv, out=ao, B=self.B, n=h * w, heads=att.heads, dim_head=att.dim_head,
                    scale=att.scale)
        out = self.act(name + ".out", h, w, C)
        self.conv(name + ".to_out", "1x1", [ao], att.to_out, out, L.EPI_RESID, resid=x)
        return out

    # ------------------------------------------------------------------ whole network
    def _build(self):
        net, B, H, W, dev = self.net, self.B, self.H, self.W, self.device
        dim, emb = net.dim, net.dim * 4
        self.x_in = self.buf("x_in", (self.x_batch, net.in_channels, H, W), torch.float32)
        self.t_in = self.buf("t_in", (B,), torch.int64)
        self.emb_in = self.buf("emb_in", (B, net.embed_input_dim), torch.float32)
        self.keep = self.buf("keep", (B,), torch.uint8)
        self.out = self.buf("out", (B, net.out_dim, H, W), torch.float32)

        # ---- embeddings (unet.py:397-421)
        bn1, bn2 = net.cond_mlp_1[1], net.cond_mlp_2[1]
        c1 = self.buf("c1", (B, dim), torch.float32)
        self.kernel("linear_small", x=self.emb_in, B=B, in_dim=net.embed_input_dim, w=net.cond_mlp_1[0].weight,
                    bias=net.cond_mlp_1[0].bias, out_dim=dim, bn=bn1, bn_train=self.training, act=L.ACT_RELU, y=c1)
        self.kernel("select_null", c=c1, keep=self.keep, null_emb=net.null_cond_emb, B=B, dim=dim)
        cemb = self.buf("c_emb", (B, emb), torch.float32)
        self.kernel("linear_small", x=c1, B=B, in_dim=dim, w=net.cond_mlp_2[0].weight, bias=net.cond_mlp_2[0].bias,
                    out_dim=emb, bn=bn2, bn_train=self.training, act=L.ACT_RELU, y=cemb)
        tf = self.buf("t_feat", (B, dim), torch.float32)
        self.kernel("time_features", t=self.t_in, B=B, dim=dim, out=tf)
        t1 = self.buf("t_hidden", (B, emb), torch.float32)
        self.kernel("linear_small", x=tf, B=B, in_dim=dim, w=net.time_mlp[1].weight, bias=net.time_mlp[1].bias,
                    out_dim=emb, act=L.ACT_GELU, y=t1)
        temb = self.buf("t_emb", (B, emb), torch.float32)
        self.kernel("linear_small", x=t1, B=B, in_dim=emb, w=net.time_mlp[3].weight, bias=net.time_mlp[3].bias,
                    out_dim=emb, act=L.ACT_NONE, y=temb)
        tc = self.buf("tc_silu", (1, 1, B, 2 * emb), torch.bfloat16)        # "image" of B pixels in one row
        self.kernel("silu_concat_bf16", t_emb=temb, dt=emb, c_emb=cemb, dc=emb, B=B, out=tc)

        # ---- every ResnetBlock's tc_mlp Linear as one row-GEMM into ss_all[B, sum 2*Cout]
        blocks = self._resblocks()
        ss_offs, tot = {}, 0
        for nm, mod in blocks:
            ss_offs[nm] = tot
            tot += 2 * mod.dim_out
        n_rows = (tot + 127) // 128 * 128
        ss_all = self.buf("ss_all", (B, n_rows), torch.float32)
        plan = plan_conv("1x1", [2 * emb], n_rows)
        shared = self.weights.packs["tc_mlp.0"].packed if "tc_mlp.0" in self.weights.packs else torch.zeros(
            n_rows, plan.nkb * KB, dtype=torch.bfloat16, device=dev)
        if "tc_bias" not in self.weights.__dict__:
            self.weights.tc_bias = torch.zeros(n_rows, dtype=torch.float32, device=dev)
        first = None
        for i, (nm, mod) in enumerate(blocks):
            lin = mod.tc_mlp[1]
            rec = self.weights.add(f"tc_mlp.{i}", lin.weight, plan_conv("1x1", [2 * emb], 2 * mod.dim_out),
                                   2 * mod.dim_out, shared=shared, row_off=ss_offs[nm])
            first = first or rec
        self._tc_bias_srcs = [(ss_offs[nm], mod.tc_mlp[1].bias) for nm, mod in blocks]
        sched = torch.tensor(plan.sched, dtype=torch.int32, device=dev)
        self.recs.append(TapGemmRec("tc_mlp", plan, [nhwc_view(tc)], B, 1, 1, (128, 1, 1), None, shared, sched, n_rows,
                                    n_rows, 128, L.EPI_BIAS | L.EPI_OUT_F32, ss_all, (n_rows, 0, 0),
                                    bias=self.weights.tc_bias))

        # ---- stem (unet.py:418-419): 7x7 conv as a 4-tap tap-GEMM over a 64-channel im2row tensor.  The stem sees no
        # conditioning, so the cond / null halves of a 2B batch share it: computed per x_batch slab.
        stem = self.act("stem", H, W, net.init_dim)
        xb = self.x_batch
        rowimg = self.buf("stem_rows", (xb, H + 1, W, 64), torch.bfloat16)
        self.kernel("stem_im2row", x=self.x_in, out=rowimg, B=xb, Cin=net.in_channels, H=H, W=W)
        splan = plan_conv("stem7", [64], net.init_dim)
        s_rows, s_tile = n_tiling(net.init_dim, False)
        spack = self.weights.stem_pack(net.init_conv.weight, s_rows)
        s_sched = torch.tensor(splan.sched, dtype=torch.int32, device=dev)
        co = net.init_dim
        for rep in range(B // xb):
            self.recs.append(TapGemmRec(f"init_conv.{rep}", splan, [nhwc_view(rowimg)], W, H, xb, tile_box(W, H), None,
                                        spack, s_sched, s_rows, co, s_tile, L.EPI_BIAS, stem[rep * xb:(rep + 1) * xb],
                                        (co, W * co, H * W * co), bias=net.init_conv.bias,
                                        # the reference conv is 7x7 over in_channels (3), not over the 64-wide im2row K
                                        algo_flops=2.0 * xb * H * W * co * net.in_channels * 49))

        # ---- down path (unet.py:425-433)
        x, h, w = stem, H, W
        skips = []
        nlev = len(net.downs)
        for k, (b1, b2, attn, down) in enumerate(net.downs):
            x, _ = self.resblock(f"downs.{k}.0", b1, [x], h, w, ss_offs[f"downs.{k}.0"])
            skips.append(x)
            x, rss = self.resblock(f"downs.{k}.1", b2, [x], h, w, ss_offs[f"downs.{k}.1"], want_rowss=True)
            x = self.linear_attention(f"downs.{k}.2", attn, x, rss, h, w)
            skips.append(x)
            if k < nlev - 1:
                h, w = h // 2, w // 2
                y = self.act(f"downs.{k}.3.out", h, w, down.weight.shape[0])
                self.conv(f"downs.{k}.3", "down4x4s2", [x], down, y)
            else:
                y = self.act(f"downs.{k}.3.out", h, w, down.weight.shape[0])
                self.conv(f"downs.{k}.3", "3x3", [x], down, y)
            x = y

        # ---- bottleneck (unet.py:435-439)
        x, rss = self.resblock("mid_block1", net.mid_block1, [x], h, w, ss_offs["mid_block1"], want_rowss=True)
        x = self.mid_attention("mid_attn", net.mid_attn, x, rss, h, w)
        x, _ = self.resblock("mid_block2", net.mid_block2, [x], h, w, ss_offs["mid_block2"])

        # ---- up path (unet.py:441-449): torch.cat((x, skip)) == two K-loop sources
        for k, (b1, b2, attn, up) in enumerate(net.ups):
            x, _ = self.resblock(f"ups.{k}.0", b1, [x, skips.pop()], h, w, ss_offs[f"ups.{k}.0"])
            x, rss = self.resblock(f"ups.{k}.1", b2, [x, skips.pop()], h, w, ss_offs[f"ups.{k}.1"], want_rowss=True)
            x = self.linear_attention(f"ups.{k}.2", attn, x, rss, h, w)
            if k < nlev - 1:
                conv = up[1]
                h, w = h * 2, w * 2
                y = self.act(f"ups.{k}.3.out", h, w, conv.weight.shape[0])
                self.conv(f"ups.{k}.3.1", "up2x3x3", [x], conv, y)
            else:
                y = self.act(f"ups.{k}.3.out", h, w, up.weight.shape[0])
                self.conv(f"ups.{k}.3", "3x3", [x], up, y)
            x = y

        # ---- head (unet.py:451-455)
        # final_conv (1x1 to out_dim channels, fp32 NCHW) lives in the final block's epilogue when the block's channels fit one
        # tile and the channel norm is fused (otherwise: the standalone head kernel over the stored bf16 tensor)
        fuse_head = net.out_dim <= 4 and net.init_dim <= 128
        if fuse_head:
            self.resblock("final_res_block", net.final_res_block, [x, stem], h, w, ss_offs["final_res_block"],
                          head=(net.final_conv, self.out))
        else:
            x, _ = self.resblock("final_res_block", net.final_res_block, [x, stem], h, w, ss_offs["final_res_block"])
            self.kernel("head_conv1", x=x, w=net.final_conv.weight, bias=net.final_conv.bias, out=self.out, B=B, H=H, W=W,
                        Cin=net.init_dim, Cout=net.out_dim)

    def _resblocks(self):
        net = self.net
        out = []
        for k, lvl in enumerate(net.downs):
            out += [(f"downs.{k}.0", lvl[0]), (f"downs.{k}.1", lvl[1])]
        out += [("mid_block1", net.mid_block1), ("mid_block2", net.mid_block2)]
        for k, lvl in enumerate(net.ups):
            out += [(f"ups.{k}.0", lvl[0]), (f"ups.{k}.1", lvl[1])]
        out.append(("final_res_block", net.final_res_block))
        return out

    def run(self, stream: int):
        # tc_mlp biases are gathered into one vector; cheap device copies, refreshed with the weights
        st = tuple(b._version for _, b in self._tc_bias_srcs)
        if getattr(self.weights, "_tc_bias_stamp", None) != st:
            for off, b in self._tc_bias_srcs:
                self.weights.tc_bias[off:off + b.numel()].copy_(b.detach())
            self.weights._tc_bias_stamp = st
        super().run(stream)
