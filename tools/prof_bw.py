"""Achieved HBM GB/s of the bandwidth-bound kernels the north star names (sampler step, q_sample, vicinal loss, GroupNorm
statistics / apply, the trainer's gather + augmentation), each timed alone with CUDA events over L2-busting rotations of its
buffers, against MEASURED_PEAKS.json: hbm_gbs.  Algorithmic bytes = every input read once + every output written once.

    python tools/prof_bw.py            # writes gpurun_out/prof_bw.json
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccdm_b200 import _lib as L  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
ROT = 6                                            # buffer sets rotated so that no launch finds its inputs in the 126 MB L2


def timed(fns, n=30):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib, dev = L.lib(), torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = []

    def report(kernel, shape, nbytes, ms, note=""):
        rec = dict(kernel=kernel, shape=shape, algorithmic_mb=round(nbytes / 2 ** 20, 1), us=round(ms * 1e3, 1),
                   gbps=round(nbytes / ms / 1e6, 1), frac_of_hbm_peak=round(nbytes / ms / 1e6 / PEAK, 3), note=note)
        print(json.dumps(rec), flush=True)
        out.append(rec)

    # ---- sampler step: guided DDIM update (CFG orthogonal projection + std rescale + pred_x0 -> x_{t-1}), bench shape
    for B, size in ((200, 64), (64, 128)):
        chw = 3 * size * size
        coef = torch.rand(4, L.STEP_NCOEF, device=dev) + 0.5
        coef[:, 6] = 0
        coef[:, 7] = 0
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        sets = []
        for _ in range(ROT):
            o = torch.randn(2 * B, chw, device=dev)
            x = torch.randn(B, chw, device=dev)
            a = L.StepArgs()
            a.out_cond, a.out_null, a.x, a.noise = o.data_ptr(), o[B:].data_ptr(), x.data_ptr(), x.data_ptr()
            a.pred_noise = a.pred_x0 = None
            a.B, a.chw, a.cond_scale, a.rescaled_phi, a.keep_parallel_frac = B, chw, 1.5, 0.7, 0.0
            a.remove_parallel, a.objective, a.clip_x0, a.cfg_plus_plus, a.sampler = 1, L.OBJ["pred_x0"], 1, 0, 0
            a.coef, a.step_counter, a.advance, a.t_rows = coef.data_ptr(), counter.data_ptr(), 0, None
            sets.append((a, o, x))
        fns = [(lambda a=a: L.check(lib.ccdm_sampler_step(ctypes.byref(a), st))) for a, _, _ in sets]
        report("sampler_step (guided DDIM, pred_x0)", f"B={B} {size}x{size}x3 fp32", 4 * B * chw * 4, timed(fns),
               "cond + null + x_t read, x_{t-1} written")

    # ---- q_sample + vicinal loss at the training shape (batch 128, 64x64), Hy on
    B, chw = 128, 3 * 64 * 64
    sa, s1 = torch.rand(1000, device=dev), torch.rand(1000, device=dev)
    lw = torch.rand(1000, device=dev)
    qs, ls = [], []
    for _ in range(ROT):
        t = {k: torch.rand(B, chw, device=dev) for k in ("img", "noise", "cov", "x0", "nout", "xt", "mo", "grad")}
        keep = torch.ones(B, dtype=torch.uint8, device=dev)
        tt = torch.randint(0, 1000, (B,), device=dev)
        qa = L.QSampleArgs()
        qa.img01, qa.noise, qa.noise2, qa.cov = t["img"].data_ptr(), t["noise"].data_ptr(), None, t["cov"].data_ptr()
        qa.keep, qa.t, qa.sqrt_acp, qa.sqrt_1m_acp = keep.data_ptr(), tt.data_ptr(), sa.data_ptr(), s1.data_ptr()
        qa.x0, qa.noise_out, qa.x_t, qa.B, qa.chw, qa.normalize = t["x0"].data_ptr(), t["nout"].data_ptr(), t["xt"].data_ptr(), B, chw, 1
        per, loss = torch.empty(B, device=dev), torch.empty(1, device=dev)
        la = L.LossArgs()
        la.model_out, la.x0, la.noise, la.cov = t["mo"].data_ptr(), t["x0"].data_ptr(), t["nout"].data_ptr(), t["cov"].data_ptr()
        la.keep, la.t, la.sqrt_acp, la.sqrt_1m_acp = keep.data_ptr(), tt.data_ptr(), sa.data_ptr(), s1.data_ptr()
        la.loss_weight, la.row_weight, la.per_sample, la.loss = lw.data_ptr(), None, per.data_ptr(), loss.data_ptr()
        la.grad_out, la.B, la.chw, la.objective = t["grad"].data_ptr(), B, chw, L.OBJ["pred_x0"]
        qs.append((qa, t, keep, tt))
        ls.append((la, per, loss))
    report("q_sample (normalize + Hy noise)", f"B={B} 64x64x3 fp32", 6 * B * chw * 4,
           timed([(lambda qa=qa: L.check(lib.ccdm_q_sample(ctypes.byref(qa), st))) for qa, *_ in qs]),
           "img, noise, cov read; x0, noise_out, x_t written")
    report("vicinal_loss rows + final + grad (3 launches)", f"B={B} 64x64x3 fp32", (2 * 3 + 1) * B * chw * 4,
           timed([(lambda la=la: L.check(lib.ccdm_vicinal_loss(ctypes.byref(la), st))) for la, *_ in ls]),
           "model_out, x0, cov read by the row pass and again by the gradient pass, gradient written")

    # ---- GroupNorm pieces of the vanilla UNet at its largest level (batch 128, 64x64, 64 channels, bf16)
    B, hw, C = 128, 64, 64
    rows = B * hw * hw
    xs = [torch.randn(B, hw, hw, C, device=dev).bfloat16() for _ in range(ROT)]
    os_ = [torch.empty_like(x) for x in xs]
    sums = torch.zeros(B, 2, C, device=dev)
    ss = 0.1 * torch.randn(B, 2 * C, device=dev)
    nb = xs[0].numel() * 2
    report("channel_stats (GroupNorm sums)", f"B={B} {hw}x{hw}x{C} bf16", nb,
           timed([(lambda x=x: L.check(lib.ccdm_channel_stats(x.data_ptr(), B, hw * hw, C, sums.data_ptr(), C, 0, 1, st))) for x in xs]),
           "one read (+ a 64 KB memset node)")
    report("affine_act (GroupNorm apply + SiLU)", f"B={B} {hw}x{hw}x{C} bf16", 2 * nb,
           timed([(lambda x=x, o=o: L.check(lib.ccdm_affine_act(x.data_ptr(), o.data_ptr(), rows, C, hw * hw, ss.data_ptr(), 2 * C, 0, 2, st)))
                  for x, o in zip(xs, os_)]), "one read, one write")

    # ---- trainer batch construction: uint8 dataset gather + flip + /255 -> fp32 NCHW (batch 128 of 64x64x3)
    N = 60000
    imgs = torch.randint(0, 256, (N, 3, 64, 64), dtype=torch.uint8, device=dev)
    idx = torch.randint(0, N, (128,), device=dev)
    aug = (torch.rand(128, device=dev) > 0.5).to(torch.uint8) << 2
    o = torch.empty(128, 3, 64, 64, device=dev)
    report("gather_augment_u8", "128 of 60000 x 3x64x64 u8 -> fp32", 128 * 3 * 64 * 64 * 5,
           timed([lambda: L.check(lib.ccdm_gather_augment_u8(imgs.data_ptr(), N, idx.data_ptr(), aug.data_ptr(), o.data_ptr(),
                                                            128, 3, 64, 64, st))]), "1.5 MB read, 6 MB written: launch-latency sized")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/prof_bw.json", "w"), indent=1)


if __name__ == "__main__":
    main()
