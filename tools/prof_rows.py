"""Kernel-only timing of the HBM-bound row kernels (ccdm_rmsnorm_act, ccdm_block_bwd) at training shapes:
achieved GB/s = algorithmic bytes (one read of each input, one write of the output) / CUDA-event time, against the
measured HBM peak in MEASURED_PEAKS.json.  Inputs are larger than L2 (126 MB) at the big shapes; the small ones are
reported as such."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccdm_b200 import _lib as L  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(os.environ.get("B", 128))
    lib = L.lib()
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for c, hw in [(64, 64), (72, 64), (128, 32), (144, 32), (256, 16), (288, 16), (512, 8), (576, 8)]:
        rows = B * hw * hw
        z = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
        dy = torch.randn_like(z)
        res = torch.randn_like(z)
        o = torch.empty_like(z)
        gain = torch.ones(c, device="cuda")
        ss = 0.1 * torch.randn(B, 2 * c, device="cuda")
        sums = torch.zeros(3, B, c, device="cuda")
        gm = math.sqrt(c)
        f_fwd = lambda: L.check(lib.ccdm_rmsnorm_act(z.data_ptr(), o.data_ptr(), rows, c, hw * hw, gain.data_ptr(), gm,
                                                     ss.data_ptr(), 2 * c, 0, None, None, L.EPI_SS | L.EPI_SILU, st))
        f_res = lambda: L.check(lib.ccdm_rmsnorm_act(z.data_ptr(), o.data_ptr(), rows, c, hw * hw, gain.data_ptr(), gm,
                                                     None, 0, 0, res.data_ptr(), None, L.EPI_SILU | L.EPI_RESID, st))
        f_bwd = lambda: L.check(lib.ccdm_block_bwd(dy.data_ptr(), z.data_ptr(), o.data_ptr(), rows, c, hw * hw,
                                                   gain.data_ptr(), gm, ss.data_ptr(), 2 * c, 0, sums.data_ptr(),
                                                   L.EPI_SS | L.EPI_SILU, st))
        nbytes = z.numel() * 2
        for name, fn, k in (("rmsnorm_act ss+silu", f_fwd, 2), ("rmsnorm_act silu+resid", f_res, 3), ("block_bwd", f_bwd, 3)):
            t = timed(fn)
            rec = dict(kernel=name, C=c, hw=hw, B=B, tensor_mb=round(nbytes / 2 ** 20, 1), us=round(t * 1e3, 1),
                       gbps=round(k * nbytes / t / 1e6, 1), frac_of_hbm_peak=round(k * nbytes / t / 1e6 / PEAK, 3))
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/prof_rows.json", "w"), indent=1)


if __name__ == "__main__":
    main()
