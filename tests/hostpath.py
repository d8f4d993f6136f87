"""TEST INFRASTRUCTURE: run the PRODUCT training / loss code of ccdm_b200 on CPU tensors.

  * every CUDA-core kernel comes from a host build of its own source (tests/hostsim);
  * the tcgen05 pieces are restated in torch with bf16 storage: the tap-GEMM based helpers of ccdm_b200.backward
    (conv forward / data gradient / weight gradient, per-sample 128x128 products, the stem's tap-GEMM and packed weight
    gradient) and the two tensor-core linear-attention entry points (ccdm_linattn_context, ccdm_linattn_dcontext);
  * ccdm_b200 itself is untouched: `install(monkeypatch)` swaps `_lib.lib`, a few `backward.*` functions and the stream getter.
"""
import ctypes as C

import torch
import torch.nn.functional as F

import ccdm_b200.backward as K
from ccdm_b200 import _lib as L
from ccdm_b200.engine import TapGemmRec, ViewRec
from ccdm_b200.plan import KB
from tests.emu import shifted
from tests.emu_engine import _view_tensor, run_tapgemm
from tests.hostsim.build import build, build_extract


def _bf16(ptr, n):
    return torch.frombuffer((C.c_uint16 * n).from_address(ptr), dtype=torch.bfloat16)


def _f32(ptr, n):
    return torch.frombuffer((C.c_float * n).from_address(ptr), dtype=torch.float32)


class HostLib:
    """libccdm_b200.so stand-in for CPU tensors."""

    def __init__(self):
        handles = [C.CDLL(build(f)) for f in ("kernels.cu", "train_kernels.cu", "backward.cu", "sampler.cu", "groupnorm.cu",
                                              "optim.cu")]
        handles.append(C.CDLL(build_extract("wgrad.cu", ["unpack_wgrad_kernel", "pack_weights_t_kernel"],
                                            ["ccdm_unpack_wgrad", "ccdm_pack_weights_t"])))
        for name, (res, args) in L.SIGNATURES.items():
            for h in handles:
                fn = getattr(h, name, None)
                if fn is not None:
                    fn.restype, fn.argtypes = res, args
                    setattr(self, name, fn)
                    break
        handles[0].hostsim_last_error.restype = C.c_char_p
        self.ccdm_last_error = handles[0].hostsim_last_error

    # ---- tcgen05 entry points, restated (linattn.cu)
    @staticmethod
    def ccdm_linattn_context(qkv, ctx, colsum, b, n, heads, w_out, wfold, c, n_rows, stream):
        """ctx[b,h,d,e] = sum_n p[n,hd] v[n,he] / S[hd],  S = sum_n p  (p = exp(k - max) written by ccdm_linattn_prep)."""
        assert not w_out and not wfold
        q = _bf16(qkv, b * n * 3 * heads * 32).float().reshape(b, n, 3, heads, 32)
        p, v = q[:, :, 1], q[:, :, 2]
        s = p.sum(1)                                                         # [b, heads, 32]
        _f32(ctx, b * heads * 32 * 32).copy_((torch.einsum("bnhd,bnhe->bhde", p, v) / s[..., None]).reshape(-1))
        if colsum:
            _f32(colsum, b * heads * 32).copy_(s.reshape(-1))
        return 0

    @staticmethod
    def ccdm_linattn_dcontext(qkv, dout, dctx, b, n, stream):
        q = _bf16(qkv, b * n * 384).float().reshape(b, n, 3, 4, 32)[:, :, 0]
        g = _bf16(dout, b * n * 128).float().reshape(b, n, 4, 32)
        _f32(dctx, b * 4 * 32 * 32).copy_(torch.einsum("bnhd,bnhe->bhde", q, g).reshape(-1))
        return 0

    @staticmethod
    def ccdm_affine_act(x, out, rows, c, rps, ss, ss_ld, ss_off, act, stream):      # kept for parity with older callers
        raise AssertionError("ccdm_affine_act is provided by the host build of kernels.cu")


# ------------------------------------------------------------------------------------------------- ccdm_b200.backward stand-ins

def _conv(kind, x, w, b=None):
    if kind == "up2x3x3":
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    stride = 2 if kind in ("down3x3s2", "down4x4s2") else 1
    return F.conv2d(x, w, b, stride=stride, padding=0 if kind == "1x1" else 1)


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def conv_forward(kind, srcs, weight, bias=None, resid=None):
    y = _conv(kind, torch.cat([_nchw(s) for s in srcs], 1), weight.detach().to(torch.bfloat16).float(),
              bias.detach() if bias is not None else None).permute(0, 2, 3, 1)
    if resid is not None:
        y = y + resid.float()
    return y.to(torch.bfloat16).contiguous()


def conv_dgrad(kind, dy, weight, cins):
    b, oh, ow, _ = dy.shape
    h, w = {"1x1": (oh, ow), "3x3": (oh, ow), "down3x3s2": (2 * oh, 2 * ow), "down4x4s2": (2 * oh, 2 * ow),
            "up2x3x3": (oh // 2, ow // 2)}[kind]
    x = torch.zeros(b, sum(cins), h, w, requires_grad=True)
    with torch.enable_grad():
        y = _conv(kind, x, weight.detach().to(torch.bfloat16).float())
    (dx,) = torch.autograd.grad(y, x, _nchw(dy))
    return [t.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous() for t in dx.split(list(cins), 1)]


def conv_wgrad(kind, srcs, dz, ksplit=0, timing=None, accumulate_into=None):
    x = torch.cat([_nchw(s) for s in srcs], 1)
    k = {"1x1": 1, "down4x4s2": 4}.get(kind, 3)
    w = torch.zeros(dz.shape[3], x.shape[1], k, k, requires_grad=True)
    with torch.enable_grad():
        y = _conv(kind, x, w)
    (dw,) = torch.autograd.grad(y, w, _nchw(dz))
    if accumulate_into is not None:
        accumulate_into.add_(dw.view_as(accumulate_into))
        return accumulate_into
    return dw


def colsum(x, accumulate_into=None):
    s = x.float().reshape(-1, x.shape[-1]).sum(0)
    if accumulate_into is not None:
        accumulate_into.add_(s)
        return accumulate_into
    return s


def per_sample_linear(src, c_off, w_batch, out, out_c_off):
    b = src.shape[0]
    x = src[..., c_off:c_off + 128].float()
    w = w_batch.float().reshape(b, 128, 128)
    out[..., out_c_off:out_c_off + 128] = torch.einsum("bhwk,bnk->bhwn", x, w).to(torch.bfloat16)
    return out


def _view(t, c_off=0, c=None):
    b, h, w, ct = t.shape
    return ViewRec(t, c_off, ct - c_off if c is None else c, w, h, b, ct, w * ct, h * w * ct)


def _launch_tapgemm(plan, tile, views, gw, gh, gb, wpacked, sched, n_rows, n, n_tile, out, ostr, ooff, bias=None, resid=None,
                    w_batch_rows=0, out_c_off=0, flags=0, ss=None):
    assert out_c_off == 0 and ss is None
    fl = flags | (L.EPI_BIAS if bias is not None else 0) | (L.EPI_RESID if resid is not None else 0)
    rec = TapGemmRec("emu", plan, list(views), gw, gh, gb, tile, None, wpacked, sched, n_rows, n, n_tile, fl, out, ostr, ooff,
                     bias=bias, w_batch_rows=w_batch_rows)
    if resid is not None:
        rc = resid.shape[3]
        rec.resid, rec.resid_strides = resid, (rc, resid.shape[2] * rc, resid.shape[1] * resid.shape[2] * rc)
    run_tapgemm(rec)


def wgrad_packed(plan, tile, views, dz, gw, gh, sched, cout, n_rows, ksplit=0, timing=None):
    """fp32 [nz*n_rows, nkb*64]: for K block (group g, tap r): dz^T . (source view shifted by the block's tap)."""
    assert plan.nz == 1
    g = torch.zeros(n_rows, plan.nkb * KB)
    d = dz.float()
    vt = [_view_tensor(v) for v in views]
    for kb in range(plan.nkb):
        grp, r = divmod(kb, plan.R)
        src, dw, dh0, c0 = plan.sched[grp]
        a = shifted(vt[src], dh0 + r, dw, gh, gw, c0)                        # [B, gh, gw, 64]
        g[:cout, kb * KB:(kb + 1) * KB] = torch.einsum("bhwn,bhwk->nk", d, a)
    return g


def install(monkeypatch):
    lib = HostLib()
    monkeypatch.setattr(L, "lib", lambda: lib)
    monkeypatch.setattr(K, "_stream", lambda: None)
    monkeypatch.setattr(K, "_check", lambda t, what: None)
    for name, fn in (("conv_forward", conv_forward), ("conv_dgrad", conv_dgrad), ("conv_wgrad", conv_wgrad), ("colsum", colsum),
                     ("per_sample_linear", per_sample_linear), ("_view", _view), ("_launch_tapgemm", _launch_tapgemm),
                     ("wgrad_packed", wgrad_packed)):
        monkeypatch.setattr(K, name, fn)
    from ccdm_b200.diffusion import GaussianDiffusion
    monkeypatch.setattr(GaussianDiffusion, "_stream", staticmethod(lambda: None))
    return lib


# ------------------------------------------------------------------------------------------------- C-ABI level ccdm_tapgemm

class _Plan:
    """The few ConvPlan fields tests.emu_engine.run_tapgemm reads, rebuilt from a ccdm_tapgemm_args struct."""

    def __init__(self, a):
        self.nz, self.ngroups, self.R = a.nz, a.ngroups, a.R
        self.nkb = a.ngroups * a.R
        n = a.nz * a.ngroups * 4
        s = torch.frombuffer((C.c_int32 * n).from_address(a.sched), dtype=torch.int32).reshape(-1, 4).tolist()
        self.sched = [tuple(r) for r in s]


def _flat(ptr, n, dtype):
    ctype, size = {torch.bfloat16: (C.c_uint16, 2), torch.float32: (C.c_float, 4)}[dtype]
    return torch.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


def tapgemm_abi(argref, stream):
    """ccdm_tapgemm(const ccdm_tapgemm_args*, stream) on host pointers: decodes the argument struct exactly as the library
    would (views, strides, packed weights, epilogue operands) and evaluates it with the documented semantics."""
    a = argref._obj if hasattr(argref, "_obj") else argref.contents
    views = []
    for i in range(a.n_src):
        v = a.src[i]
        span = (v.B - 1) * v.sB + (v.H - 1) * v.sH + (v.W - 1) * v.sW + v.C
        views.append(ViewRec(_flat(v.ptr, span, torch.bfloat16), 0, v.C, v.W, v.H, v.B, v.sW, v.sH, v.sB))
    plan = _Plan(a)
    K_ = plan.nkb * KB
    rows = (a.gB * a.w_batch_rows) if a.w_batch_rows else a.nz * a.n_rows
    wpacked = _flat(a.wpacked, rows * K_, torch.bfloat16).reshape(rows, K_)
    out_dtype = torch.float32 if a.flags & L.EPI_OUT_F32 else torch.bfloat16
    ooff = tuple(a.ooff[i] for i in range(L.MAX_Z))
    span = max(ooff[:a.nz]) + (a.gB - 1) * a.osB + (a.gH - 1) * a.osH + (a.gW - 1) * a.osW + a.N
    out = _flat(a.out, span, out_dtype)
    rec = TapGemmRec("abi", plan, views, a.gW, a.gH, a.gB, (a.tw, a.th, a.tb), None, wpacked, None, a.n_rows, a.N, a.n_tile,
                     a.flags, out, (a.osW, a.osH, a.osB), ooff, gain_mul=a.gain_mul, ss_ld=a.ss_ld, ss_off=a.ss_off,
                     q_scale=a.q_scale, q_cols=a.q_cols, w_batch_rows=a.w_batch_rows)
    npix = a.gB * a.gH * a.gW
    if a.bias:
        rec.bias = _flat(a.bias, a.N, torch.float32)
    if a.rowss:
        rec.rowss = _flat(a.rowss, npix, torch.float32)
    if a.gain:
        rec.gain = _flat(a.gain, a.N, torch.float32)
    if a.scale_shift:
        rec.ss = _flat(a.scale_shift, a.gB * a.ss_ld, torch.float32).reshape(a.gB, a.ss_ld)
    if a.resid:
        rspan = (a.gB - 1) * a.rsB + (a.gH - 1) * a.rsH + (a.gW - 1) * a.rsW + a.N
        rec.resid, rec.resid_strides = _flat(a.resid, rspan, torch.bfloat16), (a.rsW, a.rsH, a.rsB)
    if a.out_rowss:
        rec.out_rowss = _flat(a.out_rowss, npix, torch.float32)
    with torch.no_grad():
        run_tapgemm(rec)
    return 0


def linattn_context_abi(qkv, ctx, colsum, b, n, heads, w_out, wfold, c, n_rows, stream):
    """ccdm_linattn_context incl. the fused fold into to_out's weights (linattn.cu)."""
    q = _flat(qkv, b * n * 3 * heads * 32, torch.bfloat16).float().reshape(b, n, 3, heads, 32)
    p, v = q[:, :, 1], q[:, :, 2]
    s = p.sum(1)
    cm = torch.einsum("bnhd,bnhe->bhde", p, v) / s[..., None]
    if ctx:
        _flat(ctx, b * heads * 32 * 32, torch.float32).copy_(cm.reshape(-1))
    if colsum:
        _flat(colsum, b * heads * 32, torch.float32).copy_(s.reshape(-1))
    if wfold:
        w = _flat(w_out, c * heads * 32, torch.float32).reshape(c, heads, 32)
        wf = torch.einsum("che,bhde->bchd", w, cm).reshape(b, c, heads * 32)
        dst = _flat(wfold, b * n_rows * heads * 32, torch.bfloat16).reshape(b, n_rows, heads * 32)
        dst.zero_()
        dst[:, :c] = wf.to(torch.bfloat16)
    return 0


def install_engine(monkeypatch):
    """Everything `install` does, plus what the INFERENCE engine needs: ccdm_tapgemm / ccdm_linattn_context at the C-ABI
    level, the weight-pack kernels from source, eager sampler steps instead of CUDA-graph capture, and CPU devices accepted."""
    import ccdm_b200.diffusion as DM
    import ccdm_b200.engine as E
    import ccdm_b200.unet as U
    lib = install(monkeypatch)
    for cu, kernels, entries in (("tapgemm.cu", ["pack_weights_kernel"], ["ccdm_pack_weights"]),
                                 ("linattn.cu", ["kexp_bound_kernel"], ["ccdm_kexp_bound"])):
        h = C.CDLL(build_extract(cu, kernels, entries))
        for name in entries:
            fn = getattr(h, name)
            fn.restype, fn.argtypes = L.SIGNATURES[name]
            setattr(lib, name, fn)
    lib.ccdm_tapgemm = tapgemm_abi
    lib.ccdm_linattn_context = linattn_context_abi
    monkeypatch.setattr(E.UnetEngine, "_stream", staticmethod(lambda: None))
    monkeypatch.setattr(DM._SamplerState, "step", lambda self: self._launch(None))

    def engine(self):
        if self._engine is None:
            self._engine = E.UnetEngine(self)
        return self._engine
    monkeypatch.setattr(U.Unet, "engine", engine)

    import ccdm_b200.vanilla_unet as VU

    def v_engine(self):
        if self._engine is None:
            self._engine = VU.VanillaEngine(self)
        return self._engine
    monkeypatch.setattr(VU.VanillaUnet, "engine", v_engine)
    monkeypatch.setattr(VU, "_current_stream", lambda: None)
    monkeypatch.setattr(VU, "_require_cuda", lambda x: None)
    return lib
