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
Training: ``p_losses`` / ``forward`` (V:388-424,478-484) run ccdm_q_sample -> VanillaUnet (autograd nodes of
ccdm_b200/vanilla_train.py) -> ccdm_vicinal_loss with the trainer-supplied per-sample vicinal weights; the auxiliary-
regressor penalty (V:446-474) is out of scope.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from .diffusion import GaussianDiffusion, _LossGrad


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

    # ------------------------------------------------------------------ training loss (V:388-424; forward V:478-484 inherited)
    def p_losses(self, x_start, t, *, classes, noise=None, vicinal_weights=None, aux_info=None):
        """q_sample -> UNet (draws its own label-drop mask) -> weighted MSE, on ccdm_q_sample / ccdm_vicinal_loss.
        ``vicinal_weights`` [B] come from the trainer (V/trainer.py:234-289); rows the UNet dropped the label of get
        weight 1, IN PLACE, as the reference does (V:398-399).  RNG draws follow the reference order (noise, then mask)."""
        if aux_info is not None:
            raise NotImplementedError("the auxiliary-regressor penalty (V:446-474) is outside the hot path")
        b, c, h, w = x_start.shape
        dev, chw = x_start.device, c * h * w
        lib, stream = L.lib(), self._stream()
        noise = torch.randn_like(x_start) if noise is None else noise
        x0 = x_start.contiguous().float()
        noise = noise.contiguous().float()
        tt = t.to(torch.int64).contiguous()
        ones = torch.ones(b, dtype=torch.uint8, device=dev)
        x_t, noise_used, x0n = torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(x0)
        qa = L.QSampleArgs()
        qa.img01, qa.noise, qa.noise2, qa.cov, qa.normalize = x0.data_ptr(), noise.data_ptr(), None, None, 0
        qa.keep, qa.t = ones.data_ptr(), tt.data_ptr()
        qa.sqrt_acp, qa.sqrt_1m_acp = self.sqrt_alphas_cumprod.data_ptr(), self.sqrt_one_minus_alphas_cumprod.data_ptr()
        qa.x0, qa.noise_out, qa.x_t, qa.B, qa.chw = x0n.data_ptr(), noise_used.data_ptr(), x_t.data_ptr(), b, chw
        L.check(lib.ccdm_q_sample(ctypes.byref(qa), stream), "q_sample")

        row_w = None
        if vicinal_weights is not None:
            model_out, null_indx = self.unet(x_t, tt, classes, return_null_indx=True)
            vicinal_weights[null_indx] = 1                              # do not weight the unconditional rows (V:399)
            assert len(vicinal_weights) == b
            row_w = vicinal_weights.reshape(-1).to(dev, torch.float32).contiguous()
        else:
            model_out = self.unet(x_t, tt, classes, return_null_indx=False)
        model_out = model_out.contiguous()
        per = torch.empty(b, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        la = L.LossArgs()
        la.model_out, la.x0, la.noise, la.cov = model_out.data_ptr(), x0.data_ptr(), noise_used.data_ptr(), None
        la.keep, la.t = ones.data_ptr(), tt.data_ptr()
        la.sqrt_acp, la.sqrt_1m_acp = qa.sqrt_acp, qa.sqrt_1m_acp
        la.loss_weight, la.row_weight = self.loss_weight.data_ptr(), L.ptr(row_w)
        want_grad = torch.is_grad_enabled() and model_out.requires_grad
        grad_out = torch.empty(b, chw, dtype=torch.float32, device=dev) if want_grad else None
        la.per_sample, la.loss, la.grad_out = per.data_ptr(), loss.data_ptr(), L.ptr(grad_out)
        la.B, la.chw, la.objective = b, chw, L.OBJ[self.objective]
        L.check(lib.ccdm_vicinal_loss(ctypes.byref(la), stream), "vicinal_loss")
        out = loss[0]
        if want_grad:
            out = _LossGrad.apply(model_out, out, grad_out.view_as(model_out))
        return out
