"""``GaussianDiffusion`` of the vanilla CCDM (CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/diffusion.py:82-370, ``V/`` below)
over :class:`ccdm_b200.VanillaUnet`: the reference's constructor and sampling signatures, driven by the same CUDA-graph
step loop as :class:`ccdm_b200.GaussianDiffusion` (UNet program + fused guidance / prediction / DDIM-or-DDPM update).

Differences from the unified wrapper that this class encodes (all from the reference text):
  * guidance is plain classifier-free guidance + std rescale (V:34-56): no orthogonal-update projection;
  * ``ddim_sample`` calls ``model_predictions`` WITHOUT its ``rescaled_phi`` argument (V:335), so DDIM always uses the
    default 0.7 whatever the caller passes; ``p_sample`` forwards it (V:262-276);
  * ``sample`` / ``p_sample_loop`` run all ``num_timesteps`` DDPM steps unless ``preset_sampling_timesteps`` is given
    (V:293-296), and ``ddim_sampling_eta`` defaults to 1;
  * labels reach the network as ``classes`` (the embedded label); there is no covariance embedding (no ``use_Hy``).
Training (``p_losses`` / ``forward``, V:388-484) needs the GroupNorm UNet's backward, which is not built: both raise.
"""
from __future__ import annotations

import torch

from .diffusion import GaussianDiffusion


class VanillaGaussianDiffusion(GaussianDiffusion):
    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, objective="pred_noise",
                 beta_schedule="cosine", ddim_sampling_eta=1.0, offset_noise_strength=0.0, min_snr_loss_weight=False,
                 min_snr_gamma=5):
        super().__init__(model, image_size=image_size, use_Hy=False, fn_y2cov=None, timesteps=timesteps,
                         sampling_timesteps=sampling_timesteps, objective=objective, beta_schedule=beta_schedule,
                         ddim_sampling_eta=ddim_sampling_eta, offset_noise_strength=offset_noise_strength,
                         min_snr_loss_weight=min_snr_loss_weight, min_snr_gamma=min_snr_gamma, use_cfg_plus_plus=False,
                         vicinity_type=None)

    # ------------------------------------------------------------------ sampling (V:283-368)
    def _with(self, sampling_timesteps=None, eta=None):
        class _Ctx:
            def __enter__(ctx):
                ctx.saved = (self.sampling_timesteps, self.ddim_sampling_eta)
                if sampling_timesteps is not None:
                    self.sampling_timesteps = int(sampling_timesteps)
                if eta is not None:
                    self.ddim_sampling_eta = eta

            def __exit__(ctx, *exc):
                self.sampling_timesteps, self.ddim_sampling_eta = ctx.saved
        return _Ctx()

    @torch.no_grad()
    def p_sample_loop(self, classes, shape, cond_scale=6.0, rescaled_phi=0.7, save_intermediate=False,
                      preset_sampling_timesteps=None, x_init=None):
        if save_intermediate:
            raise NotImplementedError("save_intermediate copies every step to the host (V:302-304): a debugging aid only")
        steps = preset_sampling_timesteps if preset_sampling_timesteps else self.num_timesteps
        with self._with(sampling_timesteps=steps):
            return self._loop("ddpm", classes, None, shape, cond_scale, rescaled_phi, True, x_init=x_init)

    @torch.no_grad()
    def ddim_sample(self, classes, shape, cond_scale=6.0, rescaled_phi=0.7, clip_denoised=True,
                    preset_sampling_timesteps=None, preset_ddim_sampling_eta=None, save_intermediate=False, trace=None,
                    x_init=None):
        if save_intermediate:
            raise NotImplementedError("save_intermediate copies every step to the host (V:357-359): a debugging aid only")
        with self._with(preset_sampling_timesteps, preset_ddim_sampling_eta):
            # V:335 does not forward rescaled_phi: model_predictions' default 0.7 is what the reference computes with
            return self._loop("ddim", classes, None, shape, cond_scale, 0.7, clip_denoised, trace, x_init)

    @torch.no_grad()
    def sample(self, classes, cond_scale=6.0, rescaled_phi=0.7, save_intermediate=False, preset_sampling_timesteps=None):
        b = classes.shape[0]
        return self.p_sample_loop(classes, (b, self.channels, self.image_size, self.image_size), cond_scale, rescaled_phi,
                                  save_intermediate, preset_sampling_timesteps=preset_sampling_timesteps)

    # ------------------------------------------------------------------ training (not built)
    def p_losses(self, *args, **kwargs):
        raise NotImplementedError("ccdm_b200.VanillaGaussianDiffusion: training the GroupNorm UNet (V/diffusion.py:388-476) "
                                  "needs its backward, which is not built")

    def forward(self, *args, **kwargs):
        return self.p_losses(*args, **kwargs)
