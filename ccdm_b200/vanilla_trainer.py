"""Trainer of the vanilla CCDM tree (CCDM_vanilla/RC-49/RC-49_64x64/CCGM/CCDM/trainer.py:33-419, ``V/`` below): constructor
keywords, ``train(net_y2h)``, ``sample_given_labels(...)``, ``save`` / ``load``.  No ``accelerate``: one process per GPU.

``train`` builds every batch like the reference (V:221-289) -- target labels = labels of the data set + N(0, kernel_sigma),
one real image drawn uniformly from the hard (|y - t| <= kappa) or soft ((y - t)^2 <= -log(threshold) / kappa) vicinity of
each target, targets whose vicinity is empty re-drawn until it is not -- but as ONE [B, N] device mask per attempt instead of
a per-sample numpy loop; the soft weights ``exp(-kappa (y - t)^2)`` (V:282) or ones go to ``p_losses`` as
``vicinal_weights``.  The loss, the backward and the optimizer step run on the CUDA path (vanilla_diffusion / vanilla_train,
FusedAdam); gradients are averaged over the ranks when launched under torchrun.  The auxiliary-regressor penalty
(``net_aux``, ``lambda_aux``) and the periodic sample-grid PNG dump (V:311-339) are outside the hot path.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import numpy as np
import torch
from torch.optim import Adam

from . import dist as ccdm_dist
from .ema import EMA
from .optim import FusedAdam
from .utils import divisible_by


class VanillaTrainer(object):
    def __init__(self, diffusion_model, train_images, train_labels, vicinal_params, *, train_batch_size=16,
                 gradient_accumulate_every=1, train_lr=1e-4, train_num_steps=100000, ema_update_after_step=1e30,
                 ema_update_every=10, ema_decay=0.995, adam_betas=(0.9, 0.99), sample_every=1000, save_every=1000,
                 results_folder="./results", amp=False, mixed_precision_type="fp16", split_batches=True, max_grad_norm=1.0,
                 y_visual=None, net_aux=None, lambda_aux=0, aux_start_step=0):
        if lambda_aux > 0:
            raise NotImplementedError("the auxiliary-regressor penalty (V/diffusion.py:446-474) is outside the hot path")
        self.train_images, self.train_labels = train_images, train_labels
        if train_images is not None:
            assert train_images.max() > 1.0                        # V:69: images arrive un-normalised
            assert train_labels.min() >= 0 and train_labels.max() <= 1.0
        self.kernel_sigma = vicinal_params["kernel_sigma"]
        self.kappa = vicinal_params["kappa"]
        self.threshold_type = vicinal_params["threshold_type"]
        self.nonzero_soft_weight_threshold = vicinal_params["nonzero_soft_weight_threshold"]
        self.y_visual = y_visual
        self.model = diffusion_model
        self.channels = diffusion_model.channels
        self.image_size = diffusion_model.image_size
        self.sample_every, self.save_every = sample_every, save_every
        self.batch_size, self.gradient_accumulate_every = train_batch_size, gradient_accumulate_every
        assert (train_batch_size * gradient_accumulate_every) >= 16, \
            "your effective batch size (train_batch_size x gradient_accumulate_every) should be at least 16 or above"
        self.train_num_steps, self.max_grad_norm = train_num_steps, max_grad_norm
        params = [p for p in diffusion_model.parameters() if p.requires_grad]
        if params and params[0].is_cuda:
            self.opt = FusedAdam(params, lr=train_lr, betas=adam_betas, max_grad_norm=max_grad_norm)
        else:
            self.opt = Adam(params, lr=train_lr, betas=adam_betas)
        self._params = params
        self.ema = EMA(diffusion_model, update_after_step=ema_update_after_step, beta=ema_decay,
                       update_every=ema_update_every)
        self.ema.to(self.device)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(exist_ok=True)
        self.step = 0
        self._labels_dev = self._unique_dev = None

    @property
    def device(self):
        return self.model.device

    # ------------------------------------------------------------------ checkpoints (V:140-177 layout)
    def save(self, milestone):
        data = {"step": self.step, "model": self.model.state_dict(), "opt": self.opt.state_dict(),
                "ema": self.ema.state_dict(), "scaler": None}
        torch.save(data, str(self.results_folder / f"model-{milestone}.pt"))

    def load(self, milestone, return_ema=False):
        data = torch.load(str(self.results_folder / f"model-{milestone}.pt"), map_location=self.device, weights_only=True)
        self.model.load_state_dict(data["model"])
        self.step = data["step"]
        self.opt.load_state_dict(data["opt"])
        self.ema.load_state_dict(data["ema"])
        if return_ema:
            return self.ema

    # ------------------------------------------------------------------ batch construction (V:221-289)
    def _labels(self):
        if self._labels_dev is None:
            self._labels_dev = torch.from_numpy(np.asarray(self.train_labels)).float().to(self.device).reshape(-1)
            self._unique_dev = torch.unique(self._labels_dev)                 # sorted, like np.sort(set(labels)) (V:67)
        return self._labels_dev, self._unique_dev

    def _vicinity_mask(self, targets: torch.Tensor) -> torch.Tensor:
        lab, _ = self._labels()
        d = lab[None, :] - targets[:, None]                                   # [B, N]
        if self.threshold_type == "hard":
            return d.abs() <= self.kappa                                      # V:247
        return d * d <= -math.log(self.nonzero_soft_weight_threshold) / self.kappa      # V:250 (soft: kappa is 1/sigma^2-like)

    def draw_batch(self):
        """(real indices [B], target labels [B] or None, vicinal weights [B] or None) for one micro-batch."""
        lab, uniq = self._labels()
        B, dev = self.batch_size, self.device
        if self.threshold_type == "hard" and self.kappa == 0:                 # no vicinity (V:221-233)
            return torch.randint(0, len(lab), (B,), device=dev), None, None
        base = uniq[torch.randint(0, len(uniq), (B,), device=dev)]
        targets = base + torch.randn(B, device=dev) * self.kernel_sigma
        mask = self._vicinity_mask(targets)
        empty = ~mask.any(1)
        while bool(empty.any()):                                              # re-draw the noise of those targets (V:253-263)
            targets = torch.where(empty, base + torch.randn(B, device=dev) * self.kernel_sigma, targets)
            mask = self._vicinity_mask(targets)
            empty = ~mask.any(1)
        idx = (torch.rand(mask.shape, device=dev) * mask).argmax(1)           # uniform over the vicinity members
        if self.threshold_type == "soft":
            weights = torch.exp(-self.kappa * (lab[idx] - targets) ** 2)      # V:282
        else:
            weights = torch.ones(B, dtype=torch.float32, device=dev)
        return idx, targets, weights

    def process_images(self, idx: torch.Tensor) -> torch.Tensor:
        imgs = torch.from_numpy(np.asarray(self.train_images)[idx.cpu().numpy()]).to(self.device).float()
        return imgs / 255.0                                                   # normalize_images(to_neg_one_to_one=False)

    # ------------------------------------------------------------------ training loop (V:179-349)
    def train(self, net_y2h):
        if isinstance(net_y2h, torch.nn.Module):
            net_y2h = net_y2h.to(self.device).eval()
        log = os.path.join(self.results_folder, f"log_loss_niters{self.train_num_steps}.txt")
        lab, _ = self._labels()
        while self.step < self.train_num_steps:
            total = 0.0
            for _ in range(self.gradient_accumulate_every):
                idx, targets, weights = self.draw_batch()
                images = self.process_images(idx)
                with torch.no_grad():
                    classes = net_y2h(lab[idx] if targets is None else targets)
                loss = self.model(images, classes=classes, vicinal_weights=weights, aux_info=None)
                loss = loss / self.gradient_accumulate_every
                total += loss.item()
                loss.backward()
            if isinstance(self.opt, FusedAdam):
                self.opt.all_reduce_gradients()                               # clip is fused into the optimizer kernel
            else:
                ccdm_dist.all_reduce_gradients(self._params)
                torch.nn.utils.clip_grad_norm_(self._params, self.max_grad_norm)
            if self.step % 500 == 0:
                with open(log, "a") as f:
                    f.write(f"\r Step: {self.step}, Loss: {total:.4f}.")
            self.opt.step()
            self.opt.zero_grad()
            self.step += 1
            self.ema.update()
            if self.step != 0 and divisible_by(self.step, self.save_every):
                self.ema.ema_model.eval()
                self.save(self.step)

    # ------------------------------------------------------------------ sampling (V:353-419)
    def sample_given_labels(self, given_labels, net_y2h, batch_size, denorm=True, to_numpy=False, verbose=False,
                            sampler="ddpm", cond_scale=6.0, sample_timesteps=1000, ddim_eta=0):
        assert given_labels.min() >= 0 and given_labels.max() <= 1.0
        nfake = len(given_labels)
        batch_size = min(batch_size, nfake)
        assert nfake % batch_size == 0
        if isinstance(net_y2h, torch.nn.Module):
            net_y2h = net_y2h.to(self.device).eval()
        model = self.ema.ema_model
        model.eval()
        out = []
        for lo in range(0, nfake, batch_size):
            y = torch.from_numpy(np.asarray(given_labels[lo:lo + batch_size])).float().view(-1).to(self.device)
            with torch.inference_mode():
                if sampler == "ddpm":
                    img = model.sample(classes=net_y2h(y), cond_scale=cond_scale, preset_sampling_timesteps=sample_timesteps)
                elif sampler == "ddim":
                    img = model.ddim_sample(classes=net_y2h(y), shape=(y.shape[0], self.channels, self.image_size, self.image_size),
                                            cond_scale=cond_scale, preset_sampling_timesteps=sample_timesteps,
                                            preset_ddim_sampling_eta=ddim_eta)
                else:
                    raise ValueError(sampler)
                if denorm:
                    if img.min() < 0 or img.max() > 1:
                        print("\r Generated images are out of range. (min={}, max={})".format(img.min(), img.max()))
                    img = (torch.clip(img, 0, 1) * 255.0).type(torch.uint8)
            out.append(img.cpu())
            if verbose:
                print("\r {}/{} complete...".format(lo + batch_size, nfake))
        fake = torch.cat(out, dim=0)[0:nfake]
        return (fake.numpy() if to_numpy else fake), given_labels
