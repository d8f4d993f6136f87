"""fp32 restatement of the vanilla CCDM sampling loops (test oracle; TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Reference: ``CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/diffusion.py`` (``V/`` below).  The schedule tables are the
unified ones (same cosine / linear formulas; pinned bit-exactly in tests/golden/schedules.pt) -- what differs is the
guidance (plain CFG, V:34-56) and three call-site details restated below.  Pinned by tests/golden/vanilla_sampler.pt
(outputs of the reference's own GaussianDiffusion, tests/golden/make_golden_vanilla.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .diffusion_ref import Schedule, _at, _eps_from_x0, _x0_from_eps, _x0_from_v, ddim_time_pairs

Tensor = torch.Tensor
Guided = Callable[[Tensor, Tensor, Tensor, float, float], Tensor]      # (x, t, classes, cond_scale, rescaled_phi) -> output


def v_model_predictions(sch: Schedule, guided: Guided, x, t, classes, cond_scale=6.0, rescaled_phi=0.7,
                        clip_x_start=False):
    """V/diffusion.py:237-257.  Returns (pred_noise, pred_x_start)."""
    out = guided(x, t, classes, cond_scale, rescaled_phi)
    clip = (lambda v: v.clamp(-1.0, 1.0)) if clip_x_start else (lambda v: v)
    if sch.objective == "pred_noise":
        return out, clip(_x0_from_eps(sch, x, t, out))
    if sch.objective == "pred_x0":
        x0 = clip(out)
        return _eps_from_x0(sch, x, t, x0), x0
    x0 = clip(_x0_from_v(sch, x, t, out))
    return _eps_from_x0(sch, x, t, x0), x0


@torch.no_grad()
def v_ddim_sample(sch: Schedule, guided: Guided, classes, shape, *, sampling_timesteps, cond_scale=6.0, eta=1.0,
                  clip_denoised=True, trace: Optional[list] = None):
    """V/diffusion.py:312-361.  ``model_predictions`` is called without ``rescaled_phi`` (V:335): always 0.7."""
    dev = classes.device
    img = torch.randn(shape, device=dev)
    for time, time_next in ddim_time_pairs(sch.num_timesteps, sampling_timesteps):
        tt = torch.full((shape[0],), time, device=dev, dtype=torch.long)
        eps, x0 = v_model_predictions(sch, guided, img, tt, classes, cond_scale, 0.7, clip_x_start=clip_denoised)
        if trace is not None:
            trace.append((eps.clone(), x0.clone()))
        if time_next < 0:
            img = x0
            continue
        a, an = sch.alphas_cumprod[time], sch.alphas_cumprod[time_next]
        sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        img = x0 * an.sqrt() + c * eps + sigma * torch.randn_like(img)
    return (img + 1) * 0.5


@torch.no_grad()
def v_ddpm_sample(sch: Schedule, guided: Guided, classes, shape, *, steps: Optional[int] = None, cond_scale=6.0,
                  rescaled_phi=0.7):
    """sample -> p_sample_loop -> p_sample, V/diffusion.py:262-309,363-368: ``steps`` = preset_sampling_timesteps or all
    num_timesteps; x0 is always clamped (clip_denoised=True, V:265-266)."""
    dev = classes.device
    img = torch.randn(shape, device=dev)
    for t in reversed(range(0, steps if steps else sch.num_timesteps)):
        tt = torch.full((shape[0],), t, device=dev, dtype=torch.long)
        _, x0 = v_model_predictions(sch, guided, img, tt, classes, cond_scale, rescaled_phi, clip_x_start=False)
        x0 = x0.clamp(-1.0, 1.0)
        mean = _at(sch.posterior_mean_coef1, tt, img) * x0 + _at(sch.posterior_mean_coef2, tt, img) * img
        logvar = _at(sch.posterior_log_variance_clipped, tt, img)
        noise = torch.randn_like(img) if t > 0 else 0.0
        img = mean + (0.5 * logvar).exp() * noise
    return (img + 1) * 0.5


def v_p_losses(sch: Schedule, net_train: Callable[[Tensor, Tensor, Tensor], Tensor], x0: Tensor, t: Tensor, *, classes: Tensor,
               noise: Tensor, keep: Tensor, vicinal_weights: Optional[Tensor] = None) -> Tensor:
    """V/diffusion.py:388-424 (without the auxiliary-regressor penalty).  ``net_train(x, t, classes)`` is the training-mode
    UNet evaluated with the label-drop mask ``keep`` (bool [B]) it would have drawn; rows with keep == False get vicinal
    weight 1 (V:398-399).  The squared error is NOT averaged per sample first: einops ``reduce(loss, 'b ... -> b (...)',
    'mean')`` with the same axes on both sides only flattens (V:414)."""
    from .diffusion_ref import q_sample
    b, c, h, w = x0.shape
    out = net_train(q_sample(sch, x0, t, noise), t, classes)
    if sch.objective == "pred_noise":
        target = noise
    elif sch.objective == "pred_x0":
        target = x0
    else:
        target = _at(sch.sqrt_alphas_cumprod, t, x0) * noise - _at(sch.sqrt_one_minus_alphas_cumprod, t, x0) * x0
    loss = ((out - target) ** 2).flatten(1)
    loss = loss * _at(sch.loss_weight, t, loss)
    if vicinal_weights is None:
        return loss.mean()
    wts = vicinal_weights.clone().to(loss.dtype)
    wts[~keep] = 1.0
    return torch.sum(wts * loss.sum(1)) / (b * c * h * w)
