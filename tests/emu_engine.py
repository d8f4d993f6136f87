"""CPU interpreter for ccdm_b200.engine programs (test infrastructure).

Executes the *records* of a UnetProgram with torch ops, mirroring the kernels' documented semantics (bf16
storage, fp32 accumulation, zero fill outside views).  It validates the host-side wiring -- buffer connectivity,
schedules, strides, epilogue flags, scale/shift offsets -- against the oracle without a GPU, and predicts the
bf16 error the device path should show.
"""
import math

import torch
import torch.nn.functional as F

from ccdm_b200 import _lib as L
from ccdm_b200.engine import TapGemmRec, PackRec, KernelRec, ViewRec
from ccdm_b200.plan import KB
from tests.emu import pack_weights_emu, shifted


def _view_tensor(v: ViewRec):
    flat = v.base.reshape(-1)
    return torch.as_strided(flat, (v.B, v.H, v.W, v.C), (v.sB, v.sH, v.sW, 1), v.off).float()


def run_pack(r: PackRec):
    w = r.weight.detach().float()
    w4 = w.reshape(w.shape[0], w.shape[1], -1, 1)
    g = r.cin_gain.detach().reshape(-1).float() if r.cin_gain is not None else None
    packed = pack_weights_emu(r.plan, w4, r.n_rows, g, r.gain_mul)          # [nz, n_rows, K]
    K = (r.nkb_total or r.plan.nkb) * KB
    dst = r.packed.reshape(-1, K)
    c0 = r.kb0 * KB
    for z in range(r.plan.nz):
        dst[r.row_off + z * r.n_rows: r.row_off + (z + 1) * r.n_rows, c0: c0 + r.plan.nkb * KB] = packed[z].to(torch.bfloat16)


def run_tapgemm(r: TapGemmRec):
    gB, gH, gW = r.gB, r.gH, r.gW
    K = (r.plan.nkb + r.n_res) * KB
    wp = r.wpacked.reshape(-1, K).float()
    views = [_view_tensor(v) for v in r.views]
    out_flat = r.out.reshape(-1) if r.out is not None else None
    for z in range(r.plan.nz):
        acc = torch.zeros(gB, gH, gW, r.N)
        for kb in range(r.plan.nkb):
            g, rr = divmod(kb, r.plan.R)
            src, dw, dh0, c0 = r.plan.sched[z * r.plan.ngroups + g]
            if getattr(r.plan, "halo", False):                                # nine taps of one halo box
                a = shifted(views[src], dh0 + rr // 3, dw + rr % 3, gH, gW, c0)
            else:
                a = shifted(views[src], dh0 + rr, dw, gH, gW, c0)            # [vB, gH, gW, 64]
            a = a[:gB]
            if r.w_batch_rows:
                wrows = torch.stack([wp[b * r.w_batch_rows: b * r.w_batch_rows + r.N, kb * KB:(kb + 1) * KB]
                                     for b in range(gB)])                   # [B, N, 64]
                acc += torch.einsum("bhwk,bnk->bhwn", a, wrows)
            else:
                acc += a @ wp[z * r.n_rows: z * r.n_rows + r.N, kb * KB:(kb + 1) * KB].t()
        v = acc
        if r.flags & L.EPI_ROWSCALE:
            rs = 1.0 / r.rowss.reshape(gB, gH, gW, 1).sqrt().clamp_min(1e-12)
            v = v * rs
        if r.flags & L.EPI_BIAS:
            v = v + r.bias.detach().float()[: r.N]
        if r.flags & L.EPI_RMSNORM:
            inv = 1.0 / v.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12)
            v = v * inv * (r.gain.detach().reshape(-1).float() * r.gain_mul)
        if r.flags & L.EPI_SS:
            ss = r.ss[:gB]
            sc = ss[:, r.ss_off: r.ss_off + r.N].reshape(gB, 1, 1, r.N)
            sh = ss[:, r.ss_off + r.N: r.ss_off + 2 * r.N].reshape(gB, 1, 1, r.N)
            v = v * (1 + sc) + sh
        if r.flags & L.EPI_SILU:
            v = F.silu(v)
        if r.flags & L.EPI_RELU:
            v = F.relu(v)
        if r.flags & L.EPI_TANH:
            v = torch.tanh(v)
        if r.flags & L.EPI_QSOFTMAX:
            qc = r.q_cols
            q = v[..., :qc].reshape(gB, gH, gW, qc // 32, 32).softmax(-1) * r.q_scale
            v = torch.cat([q.reshape(gB, gH, gW, qc), v[..., qc:]], -1)
        if r.flags & L.EPI_KEXP:
            qc = r.q_cols
            v = torch.cat([v[..., :qc], v[..., qc:2 * qc].exp(), v[..., 2 * qc:]], -1)
        if r.flags & L.EPI_RESACC:                                           # shortcut: 1x1 groups in a second accumulator
            acc2 = torch.zeros(gB, gH, gW, r.N)
            for g2, (src, _, _, c0) in enumerate(r.res_sched):
                a2 = shifted(views[src], 0, 0, gH, gW, c0)[:gB]
                kb = r.plan.nkb + g2
                acc2 += a2 @ wp[: r.N, kb * KB:(kb + 1) * KB].t()
            if r.res_bias is not None:
                acc2 = acc2 + r.res_bias.detach().float()[: r.N]
            v = v + acc2
        if r.flags & L.EPI_RESID:
            rs_ = r.resid_strides
            res = torch.as_strided(r.resid.reshape(-1), (gB, gH, gW, r.N), (rs_[2], rs_[1], rs_[0], 1), 0).float()
            v = v + res
        if r.flags & L.EPI_HEAD:                                              # fused 1x1 head: nothing else is stored
            hw_, hb_, ho_ = r.head
            ho_[:gB].copy_((v @ hw_.detach().float().t() + hb_.detach().float()).permute(0, 3, 1, 2))
            continue
        os_ = r.out_strides
        dst = torch.as_strided(out_flat, (gB, gH, gW, r.N), (os_[2], os_[1], os_[0], 1),
                               out_flat.storage_offset() + r.ooff[z])
        dst.copy_(v.to(r.out.dtype))
        if r.flags & L.EPI_SUMSQ_OUT:
            r.out_rowss.copy_(dst.float().pow(2).sum(-1).reshape(-1))


def run_kernel(r: KernelRec):
    k, a = r.kind, r.a
    if k == "stem_im2row":
        x, Cin, H, W = a["x"], a["Cin"], a["H"], a["W"]
        out = torch.zeros(a["B"], H + 1, W, 64)
        xp = F.pad(x[: a["B"]], (3, 3, 1, 1))                                # cols -3..W+2, rows -1..H
        for dr in range(2):
            for sx in range(7):
                for c in range(Cin):
                    out[..., dr * 7 * Cin + sx * Cin + c] = xp[:, c, dr:dr + H + 1, sx:sx + W]
        a["out"].copy_(out.to(torch.bfloat16))
    elif k == "stem_pack":
        w, Cin, Cout = a["w"].detach(), a["Cin"], a["Cout"]
        packed = torch.zeros(a["n_rows"], 4, 64)
        for g in range(4):
            for dr in range(2):
                r = 2 * g + dr
                if r < 7:
                    for sx in range(7):
                        for c in range(Cin):
                            packed[:Cout, g, dr * 7 * Cin + sx * Cin + c] = w[:, c, r, sx]
        a["wpacked"].copy_(packed.reshape(a["n_rows"], 256).to(torch.bfloat16))
    elif k == "rmsnorm_act":
        C, rps = a["C"], a["rows_per_sample"]
        z = a["z"].float().reshape(-1, C)
        v = z / z.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12) * (a["gain"].detach().reshape(-1).float() * a["gain_mul"])
        fl = a["flags"]
        if fl & L.EPI_SS:
            b = torch.arange(z.shape[0]) // rps
            ss = a["ss"][b]
            v = v * (1 + ss[:, a["ss_off"]: a["ss_off"] + C]) + ss[:, a["ss_off"] + C: a["ss_off"] + 2 * C]
        if fl & L.EPI_SILU:
            v = F.silu(v)
        if fl & L.EPI_RESID:
            v = v + a["resid"].float().reshape(-1, C)
        o = v.to(torch.bfloat16)
        a["out"].copy_(o.reshape(a["out"].shape))
        if fl & L.EPI_SUMSQ_OUT:
            a["out_rowss"].copy_(o.float().pow(2).sum(-1))
    elif k == "head_conv1":
        x = a["x"].float().permute(0, 3, 1, 2)
        a["out"].copy_(F.conv2d(x, a["w"].detach(), a["bias"].detach()))
    elif k == "linattn_context":
        B, n, heads = a["B"], a["n"], a["heads"]
        qkv = a["qkv"].float().reshape(B, n, 3, heads, 32)
        pp, vv = qkv[:, :, 1], qkv[:, :, 2]                                  # [B, n, heads, 32]; pp = exp(k - bound)
        a["ctx"].copy_(torch.einsum("bnhd,bnhe->bhde", pp, vv) / pp.sum(1).permute(0, 1, 2)[..., None])
        if a.get("wfold") is not None:
            run_kernel(KernelRec("linattn_fold", dict(w_out=a["w_out"], ctx=a["ctx"], wfold=a["wfold"], B=B, C=a["C"],
                                                     n_rows=a["n_rows"], heads=heads)))
    elif k in ("linattn_kv_partials", "linattn_fold_partials", "linattn_q_out"):
        run_linattn_fused(k, a)
    elif k == "kexp_bound":
        w = a["wpacked"].float().reshape(a["n_rows"], a["K"])
        bias = a["bias"]
        bias.zero_()
        bias[a["lo"]:a["hi"]] = -1.01 * w[a["lo"]:a["hi"]].norm(dim=1) - 1e-3
    elif k == "linattn_pack_blockdiag":
        B = a["B"]
        m = a["m"].float().reshape(B, 4, 32, 32)                             # [b, h, d, e]
        if a.get("row_div") is not None:
            m = m / a["row_div"].reshape(B, 4, 32, 1)
        w = torch.zeros(B, 128, 128)
        for h in range(4):
            blk = m[:, h] if (a["transpose"] & 1) else m[:, h].transpose(1, 2)   # rows d (transpose) or rows e
            w[:, h * 32:(h + 1) * 32, h * 32:(h + 1) * 32] = blk
        a["w"].reshape(B, 128, 128).copy_(w.to(torch.bfloat16))
    elif k == "linattn_fold":
        B, C, heads = a["B"], a["C"], a["heads"]
        w = a["w_out"].detach().reshape(C, heads, 32)                        # [c, h, e]
        wf = torch.einsum("che,bhde->bchd", w, a["ctx"]).reshape(B, C, heads * 32)
        dst = a["wfold"].reshape(B, a["n_rows"], heads * 32)
        dst.zero_()
        dst[:, :C] = wf.to(torch.bfloat16)
    elif k == "attention_small":
        B, n, heads, dh = a["B"], a["n"], a["heads"], a["dim_head"]
        qkv = a["qkv"].float().reshape(B, n, 3, heads, dh)
        q, kk, vv = qkv[:, :, 0] * a["scale"], qkv[:, :, 1], qkv[:, :, 2]
        att = torch.einsum("bihd,bjhd->bhij", q, kk).softmax(-1)
        o = torch.einsum("bhij,bjhd->bihd", att, vv).reshape(B, n, heads * dh)
        a["out"].copy_(o.reshape(a["out"].shape).to(torch.bfloat16))
    elif k == "linear_small":
        y = F.linear(a["x"], a["w"].detach(), a["bias"].detach())
        bn = a.get("bn")
        if bn is not None:
            if a.get("bn_train"):
                mean, var = y.mean(0), y.var(0, unbiased=False)
                n = y.shape[0]
                bn.running_mean.mul_(0.9).add_(0.1 * mean)
                bn.running_var.mul_(0.9).add_(0.1 * var * n / max(n - 1, 1))
            else:
                mean, var = bn.running_mean, bn.running_var
            y = (y - mean) * torch.rsqrt(var + 1e-5) * bn.weight.detach() + bn.bias.detach()
        act = a["act"]
        y = {L.ACT_NONE: lambda v: v, L.ACT_RELU: F.relu, L.ACT_GELU: F.gelu, L.ACT_SILU: F.silu}[act](y)
        a["y"].copy_(y)
    elif k == "time_features":
        dim = a["dim"]
        half = dim // 2
        kf = math.log(10000) / (half - 1)
        ang = a["t"].float()[:, None] * torch.exp(torch.arange(half) * -kf)[None]
        a["out"].copy_(torch.cat([ang.sin(), ang.cos()], -1))
    elif k == "select_null":
        c = a["c"]
        c.copy_(torch.where(a["keep"].bool()[:, None], c, a["null_emb"].detach()[None].expand_as(c)))
    elif k == "silu_concat_bf16":
        v = F.silu(torch.cat([a["t_emb"], a["c_emb"]], 1))
        a["out"].copy_(v.reshape(a["out"].shape).to(torch.bfloat16))
    elif k == "channel_stats":
        x = a["x"].float().reshape(a["B"], a["rows"], a["C"])
        sums = a["sums"]                                                     # [B, 2, ld]
        if a["zero_first"]:
            sums.zero_()
        sums[:, 0, a["c_off"]: a["c_off"] + a["C"]] += x.sum(1)
        sums[:, 1, a["c_off"]: a["c_off"] + a["C"]] += x.pow(2).sum(1)
    elif k == "groupnorm_coef":
        B, Ctot, G, C0 = a["B"], a["Ctot"], a["groups"], a["C0"]
        cg = Ctot // G
        cnt = a["rows"] * cg
        s = a["sums"].reshape(B, 2, G, cg).sum(-1)                           # [B, 2, G]
        mean = s[:, 0] / cnt
        var = (s[:, 1] / cnt - mean * mean).clamp_min(0)
        rstd = torch.rsqrt(var + a["eps"])
        mean_c, rstd_c = mean.repeat_interleave(cg, 1), rstd.repeat_interleave(cg, 1)
        aa = rstd_c * a["gamma"].detach()[None]
        tt = a["beta"].detach()[None] - mean_c * aa
        if a.get("ss") is not None:
            ss, off = a["ss"][:B], a["ss_off"]
            sc = 1 + ss[:, off: off + Ctot]
            aa, tt = aa * sc, tt * sc + ss[:, off + Ctot: off + 2 * Ctot]
        coef = a["coef"]
        coef[:, :C0], coef[:, C0: 2 * C0] = aa[:, :C0] - 1, tt[:, :C0]
        if C0 < Ctot:
            C1 = Ctot - C0
            coef[:, 2 * C0: 2 * C0 + C1], coef[:, 2 * C0 + C1:] = aa[:, C0:] - 1, tt[:, C0:]
    elif k == "affine_act":
        C, rps = a["C"], a["rows_per_sample"]
        x = a["x"].float().reshape(-1, rps, C)
        ss, off = a["ss"], a["ss_off"]
        v = x * (1 + ss[: x.shape[0], None, off: off + C]) + ss[: x.shape[0], None, off + C: off + 2 * C]
        v = {0: lambda u: u, 1: F.relu, 2: F.silu}[a["act"]](v)
        a["out"].copy_(v.reshape(a["out"].shape).to(torch.bfloat16))
    elif k == "attention_tokens":
        B, n, heads, dh = a["B"], a["n"], a["heads"], a["dim_head"]
        flat = a["qkv"].float().reshape(B, n, 3 * heads * dh)
        if a["head_major"]:
            qkv = flat.reshape(B, n, heads, 3, dh).permute(0, 1, 3, 2, 4)     # -> [B, n, 3, heads, dh]
        else:
            qkv = flat.reshape(B, n, 3, heads, dh)
        q, kk, vv = qkv[:, :, 0] * a["scale"], qkv[:, :, 1], qkv[:, :, 2]
        att = torch.einsum("bihd,bjhd->bhij", q, kk).softmax(-1)
        o = torch.einsum("bhij,bjhd->bihd", att, vv).reshape(B, n, heads * dh)
        a["out"].copy_(o.reshape(a["out"].shape).to(torch.bfloat16))
    elif k == "time_features_adm":
        half = a["dim"] // 2
        f = torch.exp(-math.log(a["max_period"]) * torch.arange(half, dtype=torch.float32) / half)
        ang = a["t"].float()[:, None] * f[None]
        a["out"].copy_(torch.cat([ang.cos(), ang.sin()], -1))
    else:
        raise ValueError(k)


def run_program(prog, weights):
    with torch.no_grad():
        if hasattr(prog, "glue_in"):                                         # VanillaProgram: boundary layout copies
            prog.glue_in()
        for r in weights.program.recs:
            if isinstance(r, PackRec):
                run_pack(r)
            else:
                run_kernel(r)
        for off, b in prog._tc_bias_srcs:
            weights.tc_bias[off:off + b.numel()].copy_(b.detach())
        if hasattr(prog, "_head_bias_src"):                                  # VanillaProgram: padded bias of the output conv
            weights.head_bias[: prog._head_bias_src.numel()].copy_(prog._head_bias_src.detach())
        for r in prog.recs:
            if isinstance(r, TapGemmRec):
                run_tapgemm(r)
            else:
                run_kernel(r)
    if hasattr(prog, "glue_out"):
        prog.glue_out()
    return prog.out


def linattn_fused_units(n: int) -> int:
    """ccdm_linattn_fused_units: units per sample = tiles / (largest divisor of the tile count that is <= 8)."""
    if n <= 0 or n % 128:
        return 0
    tps = n // 128
    g = next(g for g in range(8, 0, -1) if tps % g == 0)
    return tps // g


def linattn_kv_partials_emu(x, rowss, wqkv, kbias, B, n, C):
    """x bf16 [B, n, C] (any shape with that many elements), wqkv bf16 [384, K] -> (part [B*ups,128,32], psum [B*ups,128])."""
    xf = x.float().reshape(B, n, C)
    K = wqkv.shape[1]
    w = wqkv.float()[:, :C]
    rs = 1.0 / rowss.reshape(B, n, 1).sqrt().clamp_min(1e-12)
    k = xf @ w[128:256].t()
    v = xf @ w[256:384].t()
    p = torch.exp2((k * (rs * 1.4426950408889634)) + kbias[128:256].float() * 1.4426950408889634)
    p = p.to(torch.bfloat16).float()
    v = (v * rs).to(torch.bfloat16).float()
    ups = linattn_fused_units(n)
    pu = p.reshape(B, ups, n // ups, 4, 32)
    vu = v.reshape(B, ups, n // ups, 4, 32)
    part = torch.einsum("buthd,buthe->buhde", pu, vu).reshape(B * ups, 128, 32)
    psum = pu.sum(2).reshape(B * ups, 128)
    return part, psum


def linattn_fold_partials_emu(part, psum, B, ups, w_out, C):
    ctx = part.reshape(B, ups, 4, 32, 32).sum(1) / psum.reshape(B, ups, 4, 32).sum(1)[..., None]     # [b, h, d, e]
    w = w_out.detach().float().reshape(C, 4, 32)
    return torch.einsum("che,bhde->bchd", w, ctx).reshape(B, C, 128).to(torch.bfloat16)


def linattn_q_out_emu(x, rowss, wqkv, wfold, B, n, C, n_rows, bias, gain, gain_mul, q_scale):
    xf = x.float().reshape(B, n, C)
    w = wqkv.float()[:128, :C]
    rs = 1.0 / rowss.reshape(B, n, 1).sqrt().clamp_min(1e-12)
    q = ((xf @ w.t()) * rs).reshape(B, n, 4, 32).softmax(-1) * q_scale
    q = q.reshape(B, n, 128).to(torch.bfloat16).float()
    wf = wfold.float().reshape(B, n_rows, 128)[:, :C]
    y = torch.einsum("bnk,bck->bnc", q, wf) + bias.detach().float()
    y = y / y.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12) * (gain.detach().reshape(-1).float() * gain_mul)
    return (y + xf).to(torch.bfloat16)


def run_linattn_fused(k, a):
    if k == "linattn_kv_partials":
        part, psum = linattn_kv_partials_emu(a["x"], a["rowss"], a["wqkv"], a["kbias"], a["B"], a["n"], a["C"])
        a["part"].copy_(part)
        a["psum"].copy_(psum)
    elif k == "linattn_fold_partials":
        wf = linattn_fold_partials_emu(a["part"], a["psum"], a["B"], a["ups"], a["w_out"], a["C"])
        a["wfold"].reshape(a["B"], a["n_rows"], 128)[:, :a["C"]] = wf
    else:
        o = linattn_q_out_emu(a["x"], a["rowss"], a["wqkv"], a["wfold"], a["B"], a["n"], a["C"], a["n_rows"], a["bias"],
                              a["gain"], a["gain_mul"], a["q_scale"])
        a["out"].copy_(o.reshape(a["out"].shape))
