"""Golden vectors at the HEADLINE widths: the reference's own Unet (dim 64, mults 1-2-2-4-8, the RC-49 64x64 script
configuration) at 64x64, run by the reference's own modules on CPU.

    python tests/golden/make_golden_rc64.py          # build container only; writes tests/golden/rc64.pt (~200 KB)

Weights come from ``oracle.make_state_dict`` (pure function of key / shape / seed), inputs from seeded generators.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import torch  # noqa: E402

from tests.golden.make_golden import import_reference, build_ref_unet  # noqa: E402
from oracle.unet_ref import UnetSpec  # noqa: E402

RC64 = UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4)
SEED, B, SIZE = 7, 2, 64


def rc64_inputs():
    g = torch.Generator().manual_seed(640)
    x = torch.randn(B, 3, SIZE, SIZE, generator=g)
    t = torch.tensor([37, 912], dtype=torch.long)
    half = 64
    import math
    f = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    a = torch.tensor([0.21, 0.83])[:, None] * f[None]
    emb = (torch.cat([torch.cos(a), torch.sin(a)], -1) + 1) / 2
    return x, t, emb


def main():
    torch.set_num_threads(8)
    Unet, _, _ = import_reference()
    net = build_ref_unet(Unet, RC64, SEED).eval()
    x, t, emb = rc64_inputs()
    with torch.no_grad():
        cond = net(x, t, emb, cond_drop_prob=0.0)
        null = net(x, t, emb, cond_drop_prob=1.0)
        guided, null2 = net.forward_with_cond_scale(x, t, emb, cond_scale=1.5, rescaled_phi=0.7)
    assert torch.equal(null, null2)
    torch.save({"cond": cond.clone(), "null": null.clone(), "guided": guided.clone()}, os.path.join(HERE, "rc64.pt"))
    print({k: float(v.abs().mean()) for k, v in (("cond", cond), ("null", null), ("guided", guided))})


if __name__ == "__main__":
    main()
