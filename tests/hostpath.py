"""TEST INFRASTRUCTURE: run the PRODUCT code of ccdm_b200 (inference engine, training nodes, losses) on CPU tensors.

  * every CUDA-core kernel comes from a host build of its own source (tests/hostsim);
  * the tcgen05 entry points are evaluated at the C-ABI level from their decoded argument structs / pointers, in torch with
    bf16 storage: ccdm_tapgemm (forward convs, data gradients, per-sample products, stem), ccdm_conv_wgrad, and the two
    tensor-core linear-attention GEMMs (ccdm_linattn_context, ccdm_linattn_dcontext);
  * ccdm_b200 itself is untouched -- its plans, weight packing and argument marshalling all run: `install(monkeypatch)` swaps
    only `_lib.lib`, the "is this a CUDA tensor" check and the stream getters.
"""
import ctypes as C

import torch

import ccdm_b200.backward as K
from ccdm_b200 import _lib as L
from ccdm_b200.engine import TapGemmRec, ViewRec
from ccdm_b200.plan import KB
from tests.emu import shifted
from tests.emu_engine import (_view_tensor, run_tapgemm, linattn_fused_units, linattn_kv_partials_emu,
                              linattn_q_out_emu)
from tests.hostsim.build import build, build_extract


def _flat(ptr, n, dtype):
    ctype = {torch.bfloat16: C.c_uint16, torch.float32: C.c_float}[dtype]
    return torch.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


class HostLib:
    """libccdm_b200.so stand-in for CPU tensors: host builds of every CUDA-core source file."""

    def __init__(self):
        handles = [C.CDLL(build(f)) for f in ("kernels.cu", "train_kernels.cu", "backward.cu", "sampler.cu", "groupnorm.cu",
                                              "optim.cu", "packmulti.cu")]
        handles.append(C.CDLL(build_extract("wgrad.cu", ["unpack_wgrad_kernel", "pack_weights_t_kernel"],
                                            ["ccdm_unpack_wgrad_slots", "ccdm_unpack_wgrad", "ccdm_pack_weights_t"])))
        handles.append(C.CDLL(build_extract("tapgemm.cu", ["pack_weights_kernel"], ["ccdm_pack_weights_at", "ccdm_pack_weights"])))
        handles.append(C.CDLL(build_extract("linattn.cu", ["kexp_bound_kernel"], ["ccdm_kexp_bound"])))
        handles.append(C.CDLL(build_extract("linattn_fused.cu", ["linattn_fold_parts_kernel"], ["ccdm_linattn_fold_partials"])))
        for name, (res, args) in L.SIGNATURES.items():
            for h in handles:
                fn = getattr(h, name, None)
                if fn is not None:
                    fn.restype, fn.argtypes = res, args
                    setattr(self, name, fn)
                    break
        handles[0].hostsim_last_error.restype = C.c_char_p
        self.ccdm_last_error = handles[0].hostsim_last_error
        self.ccdm_tapgemm = tapgemm_abi
        self.ccdm_conv_wgrad = wgrad_abi
        self.ccdm_linattn_context = linattn_context_abi
        self.ccdm_linattn_dcontext = linattn_dcontext_abi
        self.ccdm_linattn_fused_units = linattn_fused_units
        self.ccdm_linattn_kv_partials = linattn_kv_partials_abi
        self.ccdm_linattn_q_out = linattn_q_out_abi


# ------------------------------------------------------------------------------------------------- C-ABI level ccdm_tapgemm

class _Plan:
    """The few ConvPlan fields tests.emu_engine.run_tapgemm reads, rebuilt from a ccdm_tapgemm_args struct."""

    def __init__(self, a):
        self.nz, self.ngroups, self.R = a.nz, a.ngroups, a.R
        self.halo = bool(getattr(a, "halo", 0))
        self.nkb = a.ngroups * a.R
        n_res = a.n_res if (getattr(a, "flags", 0) & L.EPI_RESACC) else 0
        n = (a.nz * a.ngroups + n_res) * 4
        s = torch.frombuffer((C.c_int32 * n).from_address(a.sched), dtype=torch.int32).reshape(-1, 4).tolist()
        self.sched = [tuple(r) for r in s[: a.nz * a.ngroups]]
        self.res_sched = [tuple(r) for r in s[a.nz * a.ngroups:]]


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
    K_ = (plan.nkb + len(plan.res_sched)) * KB
    rows = (a.gB * a.w_batch_rows) if a.w_batch_rows else a.nz * a.n_rows
    wpacked = _flat(a.wpacked, rows * K_, torch.bfloat16).reshape(rows, K_)
    out_dtype = torch.float32 if a.flags & L.EPI_OUT_F32 else torch.bfloat16
    ooff = tuple(a.ooff[i] for i in range(L.MAX_Z))
    span = max(ooff[:a.nz]) + (a.gB - 1) * a.osB + (a.gH - 1) * a.osH + (a.gW - 1) * a.osW + a.N
    out = _flat(a.out, span, out_dtype) if a.out else None
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
    if a.flags & L.EPI_RESACC:
        rec.n_res, rec.res_sched = len(plan.res_sched), plan.res_sched
        if a.res_bias:
            rec.res_bias = _flat(a.res_bias, a.N, torch.float32)
    if a.flags & L.EPI_HEAD:
        assert a.hsC == a.gH * a.gW and a.hsB == a.head_n * a.hsC
        rec.head = (_flat(a.head_w, a.head_n * a.N, torch.float32).reshape(a.head_n, a.N), _flat(a.head_b, a.head_n, torch.float32),
                    _flat(a.head_out, a.gB * a.hsB, torch.float32).reshape(a.gB, a.head_n, a.gH, a.gW))
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


def linattn_kv_partials_abi(x, b, n, c, rowss, wqkv, kbias, part, psum, stream):
    """ccdm_linattn_kv_partials (tcgen05, linattn_fused.cu) from its raw pointers."""
    nkb = (c + 63) // 64
    ups = linattn_fused_units(n)
    pt, ps = linattn_kv_partials_emu(_flat(x, b * n * c, torch.bfloat16), _flat(rowss, b * n, torch.float32),
                                     _flat(wqkv, 384 * nkb * 64, torch.bfloat16).reshape(384, nkb * 64),
                                     _flat(kbias, 384, torch.float32), b, n, c)
    _flat(part, b * ups * 128 * 32, torch.float32).copy_(pt.reshape(-1))
    _flat(psum, b * ups * 128, torch.float32).copy_(ps.reshape(-1))
    return 0


def linattn_q_out_abi(x, b, n, c, rowss, wqkv, wfold, n_rows, bias, gain, gain_mul, q_scale, out, stream):
    """ccdm_linattn_q_out (tcgen05, linattn_fused.cu) from its raw pointers."""
    nkb = (c + 63) // 64
    o = linattn_q_out_emu(_flat(x, b * n * c, torch.bfloat16), _flat(rowss, b * n, torch.float32),
                          _flat(wqkv, 384 * nkb * 64, torch.bfloat16).reshape(384, nkb * 64),
                          _flat(wfold, b * n_rows * 128, torch.bfloat16), b, n, c, n_rows, _flat(bias, c, torch.float32),
                          _flat(gain, c, torch.float32), gain_mul, q_scale)
    _flat(out, b * n * c, torch.bfloat16).copy_(o.reshape(-1))
    return 0


def wgrad_abi(argref, stream):
    """ccdm_conv_wgrad(const ccdm_wgrad_args*, stream): packed fp32 gradient += dz^T . (source view shifted by the block's tap)
    for every K block (group g, tap r) of the FORWARD schedule, per sub-problem z."""
    a = argref._obj if hasattr(argref, "_obj") else argref.contents
    plan = _Plan(a)
    views = []
    for i in range(a.n_src):
        v = a.src[i]
        span = (v.B - 1) * v.sB + (v.H - 1) * v.sH + (v.W - 1) * v.sW + v.C
        views.append(_view_tensor(ViewRec(_flat(v.ptr, span, torch.bfloat16), 0, v.C, v.W, v.H, v.B, v.sW, v.sH, v.sB)))
    doff = [a.doff[i] for i in range(L.MAX_Z)]
    span = max(doff[:a.nz]) + (a.gB - 1) * a.dsB + (a.gH - 1) * a.dsH + (a.gW - 1) * a.dsW + a.N
    dz_flat = _flat(a.dz, span, torch.bfloat16)
    size = a.nz * a.n_rows * plan.nkb * KB
    if a.slots > 0:          # PARTIAL mode: slice s stores at s*slot_stride (no pre-zeroed buffer): slice 0 gets the sum here
        allb = _flat(a.wgrad_packed, (a.slots - 1) * a.slot_stride + size, torch.float32)
        for sl in range(a.slots):
            allb[sl * a.slot_stride: sl * a.slot_stride + size].zero_()
    g = _flat(a.wgrad_packed, size, torch.float32).reshape(a.nz * a.n_rows, plan.nkb * KB)
    for z in range(a.nz):
        d = torch.as_strided(dz_flat, (a.gB, a.gH, a.gW, a.N), (a.dsB, a.dsH, a.dsW, 1), doff[z]).float()
        for kb in range(plan.nkb):
            grp, r = divmod(kb, plan.R)
            src, dw, dh0, c0 = plan.sched[z * plan.ngroups + grp]
            q = 0
            if plan.R == 9:                                             # halo plan: K block g*9 + r*3 + q = tap (r, q)
                r, q = divmod(r, 3)
            x = shifted(views[src], dh0 + r, dw + q, a.gH, a.gW, c0)[: a.gB]
            g[z * a.n_rows: z * a.n_rows + a.N, kb * KB:(kb + 1) * KB] += torch.einsum("bhwn,bhwk->nk", d, x)
    return 0


def linattn_dcontext_abi(qkv, dout, dctx, b, n, stream):
    """ccdm_linattn_dcontext: dctx[b,h,d,e] = sum_n q_sm[n,hd] dout[n,he]  (tcgen05 in linattn.cu)."""
    q = _flat(qkv, b * n * 384, torch.bfloat16).float().reshape(b, n, 3, 4, 32)[:, :, 0]
    g = _flat(dout, b * n * 128, torch.bfloat16).float().reshape(b, n, 4, 32)
    _flat(dctx, b * 4 * 32 * 32, torch.float32).copy_(torch.einsum("bnhd,bnhe->bhde", q, g).reshape(-1))
    return 0


def install(monkeypatch):
    """ccdm_b200.backward / train / diffusion run UNCHANGED (plans, weight packing, argument structs); only the library handle is
    swapped: CUDA-core kernels from source, the tcgen05 entry points at the C-ABI level."""
    lib = HostLib()
    monkeypatch.setattr(L, "lib", lambda precision="bf16": lib)
    monkeypatch.setattr(K, "_stream", lambda: None)
    monkeypatch.setattr(K, "_check", lambda t, what: None)             # "expected a CUDA tensor": host pointers are fine here
    from ccdm_b200.diffusion import GaussianDiffusion
    monkeypatch.setattr(GaussianDiffusion, "_stream", staticmethod(lambda: None))
    import ccdm_b200.train as T
    import ccdm_b200.vanilla_unet as VU
    monkeypatch.setattr(T, "_require_cuda", lambda x: None)             # the "no CPU fallback" guards of the training forwards
    monkeypatch.setattr(VU, "_require_cuda", lambda x: None)
    import ccdm_b200.label_embedding as LE
    monkeypatch.setattr(LE, "_require_cuda", lambda x: None)
    monkeypatch.setattr(LE, "_stream", lambda: None)
    return lib


def install_engine(monkeypatch):
    """Everything `install` does, plus what the INFERENCE engine needs: eager sampler steps instead of CUDA-graph capture and
    CPU devices accepted by the engines."""
    import ccdm_b200.diffusion as DM
    import ccdm_b200.engine as E
    import ccdm_b200.unet as U
    lib = install(monkeypatch)
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

    import ccdm_b200.sngan as SG
    monkeypatch.setattr(SG, "_require_cuda", lambda z: None)
    monkeypatch.setattr(SG, "_stream", lambda: None)
    return lib
