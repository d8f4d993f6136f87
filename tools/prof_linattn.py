"""Isolated timing of the fused linear-attention kernels (csrc/linattn_fused.cu) at the bench shapes, for CUDA events / ncu.

    python tools/prof_linattn.py [--B 400] [--n 4096] [--C 64] [--iters 5]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ccdm_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=400)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--C", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    B, n, C = a.B, a.n, a.C
    dev = "cuda"
    lib = L.lib()
    torch.manual_seed(0)
    nkb = (C + 63) // 64
    x = torch.randn(B, n, C, device=dev).to(torch.bfloat16)
    rowss = x.float().pow(2).sum(-1).reshape(-1).contiguous()
    wqkv = torch.zeros(384, nkb * 64, device=dev)
    wqkv[:, :C] = torch.randn(384, C, device=dev) * 1.5
    wqkv = wqkv.to(torch.bfloat16)
    kbias = torch.zeros(384, device=dev)
    kbias[128:256] = -1.01 * wqkv.float()[128:256].norm(dim=1) - 1e-3
    w_out = (torch.randn(C, 128, device=dev) / math.sqrt(128)).contiguous()
    bias = torch.randn(C, device=dev) * 0.1
    gain = torch.ones(C, device=dev)
    ups = lib.ccdm_linattn_fused_units(n)
    n_rows = (C + 31) // 32 * 32
    part = torch.empty(B * ups, 128, 32, device=dev)
    psum = torch.empty(B * ups, 128, device=dev)
    wfold = torch.zeros(B * n_rows, 128, dtype=torch.bfloat16, device=dev)
    out = torch.empty(B, n, C, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = L.ptr
    calls = [
        ("kv_partials", lambda: lib.ccdm_linattn_kv_partials(p(x), B, n, C, p(rowss), p(wqkv), p(kbias), p(part), p(psum), st)),
        ("fold_partials", lambda: lib.ccdm_linattn_fold_partials(p(part), p(psum), B, ups, p(w_out), C, n_rows, p(wfold), st)),
        ("q_out", lambda: lib.ccdm_linattn_q_out(p(x), B, n, C, p(rowss), p(wqkv), p(wfold), n_rows, p(bias), p(gain),
                                                 math.sqrt(C), 32 ** -0.5, p(out), st)),
    ]
    res = {}
    for name, fn in calls:
        L.check(fn(), name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            L.check(fn(), name)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.iters * 1e3
    xbytes = B * n * C * 2
    print(json.dumps({"B": B, "n": n, "C": C, "us": res,
                      "kv_gbs": xbytes / res["kv_partials"] / 1e3, "qout_gbs": 2 * xbytes / res["q_out"] / 1e3}))


if __name__ == "__main__":
    main()
