"""Case table shared by the golden generator (reference side) and the tests (oracle / CUDA side)."""
import torch

from oracle.unet_ref import UnetSpec

SPECS = {
    # small copies of the RC-49 / Cell-200 shapes: every layer kind appears (4x4/s2 down, 3x3 last-level
    # resample, nearest-2x up, concat blocks with and without res_conv, linear + softmax attention)
    "tiny": UnetSpec(dim=32, dim_mults=(1, 2, 2), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4),
    "cell": UnetSpec(dim=32, dim_mults=(1, 2), in_channels=1, embed_input_dim=128, attn_dim_head=16, attn_heads=2),
    # channel counts that are multiples of 64, as the sm_100a path wants them (RC-49 widths, shallow)
    "rc_small": UnetSpec(dim=64, dim_mults=(1, 2), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4),
}
# not in the golden tables (oracle-only comparisons): 256-wide bottleneck -> channel-split GEMM + standalone norm;
# dim 72 = the UTKFace-64 widths (72 / 144 / 288 / 576: not multiples of 64, deepest level wider than 512)
SPECS["wide"] = UnetSpec(dim=32, dim_mults=(1, 2, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4)
SPECS["uk64"] = UnetSpec(dim=72, dim_mults=(1, 2, 4, 4, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32,
                         attn_heads=4)
SIZES = {"tiny": 16, "cell": 8, "rc_small": 16, "wide": 16, "uk64": 32}
BATCH = {"tiny": 3, "cell": 4, "rc_small": 2, "wide": 2, "uk64": 2}


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def scalar_sinusoid(labels, dim):
    import math
    half = dim // 2
    freq = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    ang = labels.reshape(-1)[:, None].float() * freq[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def unet_inputs(spec_name):
    spec, size, b = SPECS[spec_name], SIZES[spec_name], BATCH[spec_name]
    x = torch.randn(b, spec.in_channels, size, size, generator=_gen(100))
    t = torch.tensor([5, 500, 999, 42, 731][:b], dtype=torch.long)
    emb = (scalar_sinusoid(torch.linspace(0.1, 0.9, b), spec.embed_input_dim) + 1) / 2
    return x, t, emb


# name -> (spec, weight seed, "eval"|"train", cond_drop_prob, torch seed used for the in-UNet mask or None)
UNET_CASES = {
    "tiny_eval_cond": ("tiny", 1, "eval", 0.0, None),
    "tiny_eval_null": ("tiny", 1, "eval", 1.0, None),
    "tiny_eval_mixed": ("tiny", 1, "eval", 0.5, 3),
    "tiny_train_cond": ("tiny", 2, "train", 0.0, None),
    "cell_eval_cond": ("cell", 3, "eval", 0.0, None),
    "cell_train_mixed": ("cell", 3, "train", 0.4, 9),
    "rc_small_eval_cond": ("rc_small", 4, "eval", 0.0, None),
    "rc_small_eval_null": ("rc_small", 4, "eval", 1.0, None),
}

# name -> (spec, weight seed, cond_scale, rescaled_phi)
CFG_CASES = {
    "tiny_s1.5_phi0.7": ("tiny", 1, 1.5, 0.7),
    "tiny_s2.0_phi0": ("tiny", 1, 2.0, 0.0),
    "rc_small_s1.5_phi0.7": ("rc_small", 4, 1.5, 0.7),
}

SAMPLER_CASES = {
    "ddim_x0": dict(kind="ddim", spec="tiny", size=16, seed=1, B=2, T=1000, S=5, objective="pred_x0",
                    eta=0.0, use_Hy=False, scale=1.5, rng=11),
    "ddim_eps_eta": dict(kind="ddim", spec="tiny", size=16, seed=1, B=2, T=1000, S=4, objective="pred_noise",
                         eta=0.5, use_Hy=False, scale=1.5, rng=12),
    "ddim_x0_Hy": dict(kind="ddim", spec="tiny", size=16, seed=1, B=2, T=1000, S=4, objective="pred_x0",
                       eta=0.0, use_Hy=True, scale=1.5, rng=13),
    "ddim_v": dict(kind="ddim", spec="cell", size=8, seed=3, B=3, T=200, S=3, objective="pred_v",
                   eta=0.0, use_Hy=False, scale=2.0, rng=14),
    "ddpm_eps": dict(kind="ddpm", spec="tiny", size=16, seed=1, B=2, T=1000, S=4, objective="pred_noise",
                     eta=0.0, use_Hy=False, scale=2.0, rng=15),
    "ddpm_x0_Hy": dict(kind="ddpm", spec="cell", size=8, seed=3, B=2, T=1000, S=3, objective="pred_x0",
                       eta=0.0, use_Hy=True, scale=1.5, rng=16),
    "ddim_rc_small": dict(kind="ddim", spec="rc_small", size=16, seed=4, B=2, T=1000, S=6, objective="pred_x0",
                          eta=0.0, use_Hy=False, scale=1.5, rng=17),
}

_L = dict(spec="tiny", size=16, seed=2, B=8, p_drop=0.3, kappa=0.12, use_Hy=False, label_dim=1)
LOSS_CASES = {
    "hv_x0": dict(_L, objective="pred_x0", vic="hv", rng=21),
    "hv_x0_Hy": dict(_L, objective="pred_x0", vic="hv", use_Hy=True, rng=22),
    "sv_eps": dict(_L, objective="pred_noise", vic="sv", kappa=0.3, rng=23),
    "shv_scalar_v": dict(_L, objective="pred_v", vic="shv", rng=24),
    "shv_multi": dict(_L, objective="pred_x0", vic="shv", label_dim=3, kappa=0.25, nproj=2, rng=25),
    "ssv_multi_Hy": dict(_L, objective="pred_noise", vic="ssv", label_dim=3, kappa=0.4, nproj=3, use_Hy=True, rng=26),
    "hv_multi": dict(_L, objective="pred_x0", vic="hv", label_dim=3, kappa=0.35, rng=27),
    "hv_col_labels": dict(_L, objective="pred_x0", vic="hv", label_dim=-1, rng=28),   # labels shaped [B,1]
    "novic_eps": dict(_L, objective="pred_noise", vic=None, rng=29),
    "hv_rc_small": dict(_L, spec="rc_small", seed=4, B=4, objective="pred_x0", vic="hv", use_Hy=True, rng=30),
}


def loss_inputs(c):
    """Images in [0,1], labels in [0,1] ([B], [B,1] or [B,D]) and the tensor handed to fn_y2h."""
    spec = SPECS[c["spec"]]
    g = _gen(1000 + c["rng"])
    img = torch.rand(c["B"], spec.in_channels, c["size"], c["size"], generator=g)
    d = c["label_dim"]
    if d == 1:
        labels = torch.rand(c["B"], generator=g)
    elif d == -1:
        labels = torch.rand(c["B"], 1, generator=g)
    else:
        labels = torch.rand(c["B"], d, generator=g)
    return img, labels, labels
