"""Generator cases shared by tests/golden/make_golden_sngan.py (reference side) and the tests."""
import torch

from oracle.sngan_ref import GenSpec

GEN_SPECS = {
    "g64": GenSpec(dim_z=32, dim_embed=16, nc=3, img_size=64, gene_ch=8),
    "g128": GenSpec(dim_z=64, dim_embed=128, nc=3, img_size=128, gene_ch=8),
    "g192_mono": GenSpec(dim_z=48, dim_embed=32, nc=1, img_size=192, gene_ch=8),
}
GEN_CASES = {"g64": ("g64", 1, 3), "g128": ("g128", 2, 2), "g192_mono": ("g192_mono", 3, 2)}   # spec, weight seed, batch


def gen_inputs(spec, batch, seed=50):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(batch, spec.dim_z, generator=g)
    y = torch.rand(batch, spec.dim_embed, generator=g)           # embedded labels live in [0, 1] (sinusoid mapped to [0,1])
    return z, y
