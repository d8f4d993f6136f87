"""Vanilla (GroupNorm) UNet cases shared by tests/golden/make_golden_vanilla.py (reference side) and the tests."""
import torch

from oracle.vanilla_unet_ref import VanillaSpec

V_SPECS = {
    # group boundaries straddle the two concatenated sources in the up path (96 channels / 4 groups = 24)
    "v_tiny": VanillaSpec(model_channels=32, channel_mult=(1, 2), num_res_blocks=1, attention_resolutions=(2,),
                          num_heads=2, num_groups=4, embed_input_dim=16),
    # attention at every level: 256 / 64 / 16 tokens, head widths 16 / 32 / 32
    "v_attn": VanillaSpec(model_channels=32, channel_mult=(1, 2, 2), num_res_blocks=1, attention_resolutions=(1, 2, 4),
                          num_heads=2, num_groups=8, embed_input_dim=32),
    # the RC-49 64x64 script configuration (V/scripts/run_train_ccdm.sh): attention only in the middle block
    # (64 tokens, 4 heads of 128)
    "v_rc": VanillaSpec(model_channels=64, channel_mult=(1, 2, 4, 8), num_res_blocks=2, attention_resolutions=(16, 32),
                        num_heads=4, num_groups=8, embed_input_dim=128),
}
V_SIZES = {"v_tiny": 16, "v_attn": 16, "v_rc": 64}
V_BATCH = {"v_tiny": 3, "v_attn": 4, "v_rc": 2}
# name -> (spec, weight seed, mode, keep mask kind)
V_CASES = {
    "v_tiny_eval_mixed": ("v_tiny", 1, "eval", "mixed"),
    "v_tiny_train_cond": ("v_tiny", 2, "train", "cond"),
    "v_attn_eval_mixed": ("v_attn", 3, "eval", "mixed"),
    "v_attn_eval_null": ("v_attn", 3, "eval", "null"),
    "v_rc_eval_mixed": ("v_rc", 4, "eval", "mixed"),
}
# name -> (spec, weight seed, cond_scale, rescaled_phi)
V_CFG_CASES = {"v_tiny_cfg": ("v_tiny", 1, 1.5, 0.7), "v_attn_cfg_plain": ("v_attn", 3, 2.0, 0.0)}


def keep_mask(kind, b):
    return {"cond": torch.ones(b, dtype=torch.bool), "null": torch.zeros(b, dtype=torch.bool),
            "mixed": torch.tensor([True, False, True, False, True, True][:b])}[kind]


def vanilla_inputs(spec_name, seed=70):
    spec, size, b = V_SPECS[spec_name], V_SIZES[spec_name], V_BATCH[spec_name]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, spec.in_channels, size, size, generator=g)
    t = torch.randint(0, 1000, (b,), generator=g)
    classes = torch.rand(b, spec.embed_input_dim, generator=g)
    return x, t, classes

# sampling through the reference's own GaussianDiffusion (V/diffusion.py); name -> settings
V_SAMPLER_CASES = {
    "v_ddim_x0": dict(kind="ddim", spec="v_tiny", seed=1, B=2, T=1000, S=5, objective="pred_x0", eta=0.0, scale=1.5,
                      phi=0.3, rng=41),          # phi is ignored by the reference's ddim_sample (V:335)
    "v_ddim_eps_eta": dict(kind="ddim", spec="v_tiny", seed=1, B=3, T=1000, S=4, objective="pred_noise", eta=1.0,
                           scale=2.0, phi=0.7, rng=42),
    "v_ddim_v_attn": dict(kind="ddim", spec="v_attn", seed=3, B=2, T=200, S=3, objective="pred_v", eta=0.5, scale=1.5,
                          phi=0.7, rng=43),
    "v_ddpm_x0": dict(kind="ddpm", spec="v_tiny", seed=1, B=2, T=1000, S=4, objective="pred_x0", eta=1.0, scale=1.5,
                      phi=0.4, rng=44),
    "v_ddpm_eps_unguided": dict(kind="ddpm", spec="v_attn", seed=3, B=2, T=1000, S=3, objective="pred_noise", eta=1.0,
                                scale=1.0, phi=0.7, rng=45),
}


def sampler_classes(c):
    g = torch.Generator().manual_seed(c["rng"] + 1000)
    return torch.rand(c["B"], V_SPECS[c["spec"]].embed_input_dim, generator=g)

# training loss through the reference's own GaussianDiffusion.p_losses + backward; name -> settings
V_LOSS_CASES = {
    "v_loss_x0_vic": dict(spec="v_tiny", seed=2, objective="pred_x0", vic=True, kind="mixed", rng=51),
    "v_loss_eps_plain": dict(spec="v_tiny", seed=2, objective="pred_noise", vic=False, kind="cond", rng=52),
    "v_loss_v_vic_attn": dict(spec="v_attn", seed=3, objective="pred_v", vic=True, kind="mixed", rng=53),
}
V_LOSS_GRAD_KEYS = ("out.2.bias", "down_blocks.0.0.weight", "middle_block.0.tc_mlp.1.weight", "out.0.weight")


def loss_inputs(c):
    spec, size, b = V_SPECS[c["spec"]], V_SIZES[c["spec"]], V_BATCH[c["spec"]]
    g = torch.Generator().manual_seed(c["rng"])
    x0 = torch.rand(b, spec.in_channels, size, size, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (b,), generator=g)
    classes = torch.rand(b, spec.embed_input_dim, generator=g)
    noise = torch.randn(b, spec.in_channels, size, size, generator=g)
    weights = (0.2 + torch.rand(b, generator=g)) if c["vic"] else None
    return x0, t, classes, noise, weights
