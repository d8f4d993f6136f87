"""One guided UNet evaluation of the bench workload (RC-49 64x64, 2x200 batch), eager, for ncu.

    python tools/prof_forward.py [--batch 200] [--iters 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ccdm_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=200)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(111)
    net = ccdm_b200.Unet(dim=64, dim_mults=(1, 2, 2, 4, 8), cond_drop_prob=0.1).to(dev).eval()
    eng = net.engine()
    B = a.batch
    x = torch.randn(B, 3, 64, 64, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)
    emb = ccdm_b200.LabelEmbed(y2h_type="sinusoidal", h_dim=128, device=dev).fn_y2h(torch.linspace(0, 1, B, device=dev))
    for _ in range(a.iters):
        eng.forward_pair(x, t, emb)
    torch.cuda.synchronize()
    # record names in launch order (one launch per record): join with an ncu launch list via tools/ncu_layers.py
    import json
    from ccdm_b200.engine import TapGemmRec, tapgemm_flops
    prog = eng.program(2 * B, B, 64, 64, False)
    recs = []
    for r in prog.recs:
        if isinstance(r, TapGemmRec):
            recs.append({"name": r.name, "kind": r.plan.kind, "flops": tapgemm_flops(r), "grid": [r.gB, r.gH, r.gW],
                         "cin": sum(r.plan.cins), "N": r.N})
        else:
            recs.append({"name": r.kind, "kind": "kernel"})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(recs, open(os.path.join(ROOT, "gpurun_out", "forward_recs.json"), "w"))
    print("ok", len(recs), "records per forward")


if __name__ == "__main__":
    main()
