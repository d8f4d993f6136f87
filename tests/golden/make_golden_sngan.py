"""Golden outputs of the reference's OWN one-step generator (CCDM_unified/models/sngan.py), CPU, build container only.

    python tests/golden/make_golden_sngan.py        # needs /root/reference; writes tests/golden/sngan.pt

Weights come from oracle.sngan_ref.make_state_dict (loaded with strict=True, which also pins key names and shapes);
inputs from seeded generators.  Only the outputs are stored.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import importlib.util  # noqa: E402

import torch  # noqa: E402

from oracle.sngan_ref import make_state_dict  # noqa: E402
from tests.golden.sngan_cases import GEN_CASES, GEN_SPECS, gen_inputs  # noqa: E402


def main():
    spec_mod = importlib.util.spec_from_file_location("ref_sngan", "/root/reference/CCDM_unified/models/sngan.py")
    ref = importlib.util.module_from_spec(spec_mod)
    spec_mod.loader.exec_module(ref)
    out = {}
    for name, (sname, seed, batch) in GEN_CASES.items():
        s = GEN_SPECS[sname]
        net = ref.sngan_generator(dim_z=s.dim_z, dim_embed=s.dim_embed, nc=s.nc, img_size=s.img_size, gene_ch=s.gene_ch)
        net.load_state_dict(make_state_dict(s, seed), strict=True)
        net.eval()
        z, y = gen_inputs(s, batch)
        with torch.no_grad():
            out[name] = {"out": net(z, y).clone(), "keys": list(net.state_dict().keys())}
        print(name, tuple(out[name]["out"].shape), float(out[name]["out"].abs().mean()))
        if name == "g64":
            # training mode (batch-statistics BatchNorm2d, sngan.py:19-36): output and the updated running statistics
            z, y = gen_inputs(s, 6, seed=51)
            net.train()
            with torch.no_grad():
                o = net(z, y).clone()
            out["g64_train"] = {"out": o, "stats": {k: v.clone() for k, v in net.state_dict().items()
                                                     if "running_" in k or "num_batches" in k}}
            print("g64_train", tuple(o.shape), float(o.abs().mean()))
    torch.save(out, os.path.join(HERE, "sngan.pt"))


if __name__ == "__main__":
    main()
