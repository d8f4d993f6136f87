"""CPU emulation of the tap-GEMM *semantics* (test infrastructure): checks ccdm_b200.plan schedules and the
weight-packing rule against torch convolutions without a GPU.  Mirrors what tapgemm.cu / pack_weights_kernel do
with TMA boxes (out-of-range reads are zero) -- not how they do it."""
import torch

from ccdm_b200.plan import KB, ConvPlan


def pack_weights_emu(plan: ConvPlan, w: torch.Tensor, n_rows: int, cin_gain=None, gain_mul: float = 1.0, n_off: int = 0,
                     n_count=None):
    """w: [Cout, Cin_total, kh, kw] fp32 -> [nz, n_rows, nkb*64] fp32 (the kernel stores bf16).
    Transposed plans (data gradient) read W[k_channel][n_off + n][tap] for n < n_count."""
    if plan.transposed:
        wt = w.reshape(w.shape[0], w.shape[1], -1).transpose(0, 1)          # [Cin_total(n), Cout(k), taps]
        wt = wt[n_off:n_off + (n_count if n_count is not None else wt.shape[0] - n_off)]
        cout, cin_total = wt.shape[0], wt.shape[1]
    else:
        cout, cin_total = w.shape[0], w.shape[1]
        wt = w.reshape(cout, cin_total, -1)
    out = torch.zeros(plan.nz, n_rows, plan.nkb * KB)
    for z in range(plan.nz):
        for kb in range(plan.nkb):
            cin0, nvalid, mask, _ = plan.psched[z * plan.nkb + kb]
            taps = [t for t in range(plan.ntaps) if mask >> t & 1]
            blk = wt[:, cin0:cin0 + nvalid, :][:, :, taps].sum(-1) * gain_mul
            if cin_gain is not None:
                blk = blk * cin_gain[cin0:cin0 + nvalid][None, :]
            out[z, :cout, kb * KB:kb * KB + nvalid] = blk
    return out


def make_views(plan: ConvPlan, srcs):
    """srcs: list of NHWC tensors, one per concatenated source -> list of views in schedule order."""
    if plan.n_views == 1:
        return list(srcs)
    views = []
    for s in srcs:
        for pr in range(2):
            for pq in range(2):
                views.append(s[:, pr::2, pq::2, :])
    return views


def shifted(view, dh, dw, gh, gw, c0):
    """view[b, h+dh, w+dw, c0:c0+64] for h<gh, w<gw with zero fill outside the view (TMA semantics)."""
    b, vh, vw, vc = view.shape
    out = torch.zeros(b, gh, gw, KB)
    h_lo, h_hi = max(0, -dh), min(gh, vh - dh)
    w_lo, w_hi = max(0, -dw), min(gw, vw - dw)
    nc = min(KB, vc - c0)
    if h_hi > h_lo and w_hi > w_lo and nc > 0:
        out[:, h_lo:h_hi, w_lo:w_hi, :nc] = view[:, h_lo + dh:h_hi + dh, w_lo + dw:w_hi + dw, c0:c0 + nc]
    return out


def tapgemm_emu(plan: ConvPlan, srcs, w, gh, gw, n_rows=None, cin_gain=None, n_off=0, n_count=None):
    """Returns [nz, B, gh, gw, N] fp32 accumulators (no epilogue)."""
    nout = (n_count if n_count is not None else w.shape[1] - n_off) if plan.transposed else w.shape[0]
    n_rows = n_rows or nout
    packed = pack_weights_emu(plan, w, n_rows, cin_gain, n_off=n_off, n_count=n_count)
    views = make_views(plan, srcs)
    outs = []
    for z in range(plan.nz):
        acc = 0
        for g in range(plan.ngroups):
            src, dw, dh0, c0 = plan.sched[z * plan.ngroups + g]
            for r in range(plan.R):
                kb = g * plan.R + r
                if getattr(plan, "halo", False):                      # nine taps of one halo box: (r, q) = divmod(tap, 3)
                    a = shifted(views[src], dh0 + r // 3, dw + r % 3, gh, gw, c0)
                else:
                    a = shifted(views[src], dh0 + r, dw, gh, gw, c0)
                acc = acc + a @ packed[z, :nout, kb * KB:(kb + 1) * KB].t()
        outs.append(acc)
    return torch.stack(outs)


def assemble_parity(out4):
    """[4, B, h, w, C] parity planes (pa, pb) -> [B, 2h, 2w, C]."""
    _, b, h, w, c = out4.shape
    full = torch.zeros(b, 2 * h, 2 * w, c)
    for pa in range(2):
        for pb in range(2):
            full[:, pa::2, pb::2, :] = out4[pa * 2 + pb]
    return full
