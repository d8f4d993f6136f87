"""Trainer with the reference's surface (CCDM_unified/trainer.py:43-871): constructor keywords, ``train(fn_y2h)``,
``sample_given_labels(...)``, ``save`` / ``load``.  No ``accelerate``: one process per GPU (``ccdm_b200.dist``).

``sample_given_labels`` is the hot sampling caller.  ``train`` builds batches like the reference (vicinity search
vectorised on the device instead of a per-sample Python loop with a ``.cpu()`` sync each), evaluates the fused loss,
back-propagates through the CUDA graph of ccdm_b200/train.py, averages gradients over the ranks when launched under
torchrun (``ccdm_b200.dist.all_reduce_gradients``) and steps torch's Adam + the EMA like the reference.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch
from torch.optim import Adam

from . import dist as ccdm_dist
from .ema import EMA
from .optim import FusedAdam
from .diffusion import generate_random_vectors
from .utils import divisible_by


class Trainer(object):
    def __init__(self, data_name, diffusion_model, train_images, train_labels, vicinal_params, *, train_batch_size=16,
                 gradient_accumulate_every=1, train_lr=1e-4, train_num_steps=100000, ema_update_after_step=1e30,
                 ema_update_every=10, ema_decay=0.995, adam_betas=(0.9, 0.99), sample_every=1000, save_every=1000,
                 results_folder="./results", amp=False, mixed_precision_type="fp16", split_batches=True,
                 max_grad_norm=1.0, y_visual=None, nrow_visual=6, cond_scale_visual=1.5, vicinity_type="shv",
                 kappa=None, sigma_delta=None, vector_type="gaussian", num_projections=1, distance="l2", label_dim=1,
                 adaptive_slicing=False, hyperparameter="rule_of_thumb", percentile=5.0):
        self.data_name = data_name
        self.train_images, self.train_labels = train_images, train_labels
        if train_images is not None:
            assert train_images.max() > 1.0                       # trainer.py:89: images arrive un-normalised
            assert train_labels.min() >= 0 and train_labels.max() <= 1.0
        self.kernel_sigma = vicinal_params["kernel_sigma"]
        self.kappa = vicinal_params["kappa"]
        self.nonzero_soft_weight_threshold = vicinal_params["nonzero_soft_weight_threshold"]
        self.y_visual, self.cond_scale_visual, self.nrow_visual = y_visual, cond_scale_visual, nrow_visual
        self.model = diffusion_model
        self.channels = diffusion_model.channels
        self.image_size = diffusion_model.image_size
        self.sample_every, self.save_every = sample_every, save_every
        self.batch_size, self.gradient_accumulate_every = train_batch_size, gradient_accumulate_every
        assert (train_batch_size * gradient_accumulate_every) >= 16, \
            "your effective batch size (train_batch_size x gradient_accumulate_every) should be at least 16 or above"
        self.train_num_steps, self.max_grad_norm = train_num_steps, max_grad_norm
        params = list(diffusion_model.parameters())
        if params and params[0].is_cuda:
            # clip_grad_norm_ + Adam fused over flat buffers (ccdm_b200/optim.py); same arithmetic as torch's Adam
            self.opt = FusedAdam(params, lr=train_lr, betas=adam_betas, max_grad_norm=max_grad_norm)
        else:
            self.opt = Adam(params, lr=train_lr, betas=adam_betas)
        self.ema = EMA(diffusion_model, update_after_step=ema_update_after_step, beta=ema_decay,
                       update_every=ema_update_every)
        self.ema.to(self.device)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(exist_ok=True)
        self.step = 0
        self.vicinity_type, self.vector_type, self.num_projections = vicinity_type, vector_type, num_projections
        self.distance, self.label_dim = distance, label_dim
        self.adaptive_slicing, self.hyperparameter, self.percentile = adaptive_slicing, hyperparameter, percentile
        if kappa is not None:
            self.kappa = kappa
        self.sigma_delta = sigma_delta if sigma_delta is not None else self.kernel_sigma
        self._labels_dev = None

    @property
    def device(self):
        return self.model.device

    # ------------------------------------------------------------------ checkpoints (trainer.py:488-535 layout)
    def save(self, milestone):
        data = {"step": self.step, "model": self.model.state_dict(), "opt": self.opt.state_dict(),
                "ema": self.ema.state_dict(), "scaler": None}
        torch.save(data, str(self.results_folder / f"model-{milestone}.pt"))

    def load(self, milestone, return_ema=False, return_unet=False):
        data = torch.load(str(self.results_folder / f"model-{milestone}.pt"), map_location=self.device,
                          weights_only=True)
        self.model.load_state_dict(data["model"])
        self.step = data["step"]
        self.opt.load_state_dict(data["opt"])
        self.ema.load_state_dict(data["ema"])
        if return_ema:
            return self.ema
        if return_unet:
            return self.model.model

    # ------------------------------------------------------------------ batch construction (trainer.py:308-482)
    def _train_labels_dev(self):
        if self._labels_dev is None:
            t = torch.from_numpy(np.asarray(self.train_labels)).float().to(self.device)
            self._labels_dev = t.view(len(t), -1)
        return self._labels_dev

    def sample_real_indices(self, target_labels: torch.Tensor) -> torch.Tensor:
        """One real sample per target label, uniformly among those whose label lies in the hard vicinity (or, for
        the sliced types, whose projection on a random direction does); nearest neighbour when the vicinity is
        empty.  Vectorised restatement of trainer.py:317-459 (one [B, N] mask instead of B host round trips)."""
        lab = self._train_labels_dev()                                    # [N, D]
        tgt = target_labels.view(len(target_labels), -1).to(lab)          # [B, D]
        if self.vicinity_type in ("shv", "ssv") and lab.shape[1] > 1:
            v = generate_random_vectors(self.vector_type, lab.shape[1], self.num_projections, lab.device).float()
            vn = torch.nn.functional.normalize(v, dim=1)
            d = ((tgt @ vn.t())[:, None, :] - (lab @ vn.t())[None, :, :]).abs()          # [B, N, P]
            mask = (d <= (self.kappa * v.norm(dim=1))[None, None, :]).any(-1)
        elif self.vicinity_type == "sv":
            mask = torch.ones(len(tgt), len(lab), dtype=torch.bool, device=lab.device)
        else:
            mask = torch.cdist(tgt, lab) <= self.kappa
        nearest = torch.cdist(tgt, lab).argmin(1)
        score = torch.rand(mask.shape, device=lab.device) * mask
        pick = score.argmax(1)
        return torch.where(mask.any(1), pick, nearest)

    def process_images(self, idx: torch.Tensor) -> torch.Tensor:
        imgs = torch.from_numpy(np.asarray(self.train_images)[idx.cpu().numpy()]).to(self.device).float()
        if self.data_name in ("UTKFace", "Cell200"):
            flip = torch.rand(len(imgs), device=self.device) < 0.5
            imgs = torch.where(flip[:, None, None, None], imgs.flip(3), imgs)
        return imgs / 255.0                                               # normalize_images(to_neg_one_to_one=False)

    # ------------------------------------------------------------------ training loop (trainer.py:537-780)
    def train(self, fn_y2h):
        uniq = torch.unique(self._train_labels_dev(), dim=0)
        log = os.path.join(self.results_folder, f"log_loss_niters{self.train_num_steps}.txt")
        while self.step < self.train_num_steps:
            total = 0.0
            for _ in range(self.gradient_accumulate_every):
                if self.vicinity_type in ("shv", "ssv", "hv", "sv"):
                    tgt = uniq[torch.randint(0, len(uniq), (self.batch_size,), device=self.device)]
                    tgt = tgt + torch.randn_like(tgt) * float(np.mean(self.sigma_delta))
                    idx = self.sample_real_indices(tgt)
                    weights = torch.ones(self.batch_size, device=self.device)
                else:
                    idx = torch.randint(0, len(self.train_images), (self.batch_size,), device=self.device)
                    weights = None
                images = self.process_images(idx)
                labels = self._train_labels_dev()[idx]
                labels = labels.view(-1) if labels.shape[1] == 1 else labels
                kw = {}
                if self.vicinity_type in ("shv", "ssv"):
                    kw = dict(vicinity_type=self.vicinity_type, kappa=self.kappa, vector_type=self.vector_type,
                              num_projections=self.num_projections)
                loss = self.model(images, labels_emb=fn_y2h(labels), labels=labels, vicinal_weights=weights, **kw)
                loss = loss / self.gradient_accumulate_every
                total += loss.item()
                loss.backward()
            if isinstance(self.opt, FusedAdam):
                self.opt.all_reduce_gradients()                   # one collective over the flat gradient buffer; clip is fused
            else:
                ccdm_dist.all_reduce_gradients(list(self.model.parameters()))
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
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

    # ------------------------------------------------------------------ sampling (trainer.py:782-869)
    def sample_given_labels(self, given_labels, fn_y2h, batch_size, denorm=True, to_numpy=True, verbose=False,
                            sampler="ddpm", cond_scale=6.0, sample_timesteps=1000, ddim_eta=0):
        assert given_labels.min() >= 0 and given_labels.max() <= 1.0
        nfake = len(given_labels)
        batch_size = min(batch_size, nfake)
        assert nfake % batch_size == 0
        model = self.ema.ema_model
        model.eval()
        out = []
        for lo in range(0, nfake, batch_size):
            y = torch.from_numpy(np.asarray(given_labels[lo:lo + batch_size])).float().view(-1).to(self.device)
            with torch.inference_mode():
                if sampler == "ddpm":
                    img = model.sample(labels_emb=fn_y2h(y), labels=y, cond_scale=cond_scale)
                elif sampler == "ddim":
                    img = model.ddim_sample(labels_emb=fn_y2h(y), labels=y,
                                            shape=(y.shape[0], self.channels, self.image_size, self.image_size),
                                            cond_scale=cond_scale)
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
