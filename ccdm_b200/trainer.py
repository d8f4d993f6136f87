"""Trainer with the reference's surface (CCDM_unified/trainer.py:43-871): constructor keywords, ``train(fn_y2h)``,
``sample_given_labels(...)``, ``save`` / ``load``.  No ``accelerate``: one process per GPU (``ccdm_b200.dist``).

What is different from the reference is WHERE the work runs, not what is computed:

* the dataset is uploaded once as uint8 and stays on the device; a micro-batch is built by device ops only -- target labels,
  the vicinity search as ONE [B, N] mask / count matrix instead of B host round trips (trainer.py:317-459), the per-sample
  vicinal weights (trainer.py:663-690), and ``ccdm_gather_augment_u8`` (index gather + flips / quarter turns + /255,
  trainer.py:461-482) -- so there is no ``.cpu()``, no ``.item()`` and no host gather in the step;
* on a GPU the whole optimizer step (``gradient_accumulate_every`` micro-batches: batch construction, q_sample, UNet forward,
  vicinal loss, backward; then gradient all-reduce, clip, Adam) is ONE CUDA graph (``train_graph.GraphedTrainStep``); the loss is
  accumulated on the device and read every 500 steps, when the reference writes its log line;
* data-parallel runs (torchrun): replicas start from rank 0's weights, gradients are averaged over NCCL inside the captured
  step, and -- as in the reference (trainer.py:488-491,727-779) -- only the main process updates the EMA, samples previews and
  writes checkpoints / logs.

Deliberate deviations, all in how ties / randomness are consumed (the reference's own choices are RNG-stream dependent):
the sliced search picks uniformly among the (up to) ten training samples matched by the most projections with ties broken by
the lower index (the reference sorts with an unstable ``argsort``, trainer.py:387-397); numpy's host RNG is replaced by the
torch CUDA generator; ``amp`` / ``mixed_precision_type`` are accepted and ignored (the CUDA path always computes in bf16 with
fp32 master weights); preview grids need torchvision (skipped with a warning when it is missing).
"""
from __future__ import annotations

import os
import warnings
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import Adam

from . import _lib as L
from . import dist as ccdm_dist
from .ema import EMA
from .optim import FusedAdam
from .diffusion import generate_random_vectors, compute_distance
from .utils import divisible_by


class Trainer(object):
    def __init__(self, data_name, diffusion_model, train_images, train_labels, vicinal_params, *, train_batch_size=16,
                 gradient_accumulate_every=1, train_lr=1e-4, train_num_steps=100000, ema_update_after_step=1e30,
                 ema_update_every=10, ema_decay=0.995, adam_betas=(0.9, 0.99), sample_every=1000, save_every=1000,
                 results_folder="./results", amp=False, mixed_precision_type="fp16", split_batches=True,
                 max_grad_norm=1.0, y_visual=None, nrow_visual=6, cond_scale_visual=1.5, vicinity_type="shv",
                 kappa=None, sigma_delta=None, vector_type="gaussian", num_projections=1, distance="l2", label_dim=1,
                 adaptive_slicing=False, hyperparameter="rule_of_thumb", percentile=5.0, use_cuda_graph=True):
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
        self.rank, _, self.world = ccdm_dist.env_world()
        self.is_main = self.rank == 0
        ccdm_dist.broadcast_parameters(diffusion_model)           # replicas start from rank 0 (accelerate.prepare / DDP)
        params = list(diffusion_model.parameters())
        self.on_cuda = bool(params) and params[0].is_cuda
        if self.on_cuda:
            # clip_grad_norm_ + Adam fused over flat buffers (ccdm_b200/optim.py); same arithmetic as torch's Adam
            self.opt = FusedAdam(params, lr=train_lr, betas=adam_betas, max_grad_norm=max_grad_norm,
                                 early_params=ccdm_dist.early_gradient_params(diffusion_model))
            ccdm_dist.wire_overlap(diffusion_model, self.opt)
        else:
            self.opt = Adam(params, lr=train_lr, betas=adam_betas)
        self.ema = EMA(diffusion_model, update_after_step=ema_update_after_step, beta=ema_decay,
                       update_every=ema_update_every)
        self.ema.to(self.device)
        self.results_folder = Path(results_folder)
        if self.is_main:
            self.results_folder.mkdir(exist_ok=True)
        self.step = 0
        self.vicinity_type, self.vector_type, self.num_projections = vicinity_type, vector_type, num_projections
        self.distance, self.label_dim = distance, label_dim
        self.adaptive_slicing, self.hyperparameter, self.percentile = adaptive_slicing, hyperparameter, percentile
        # trainer.py:159-171: the keyword arguments win over vicinal_params; when either is missing (and the per-batch
        # adaptive mode is off) both are derived from the training labels
        if kappa is not None:
            self.kappa = kappa
        self.sigma_delta = sigma_delta
        if not adaptive_slicing and (kappa is None or sigma_delta is None) and train_labels is not None:
            self.sigma_delta, self.kappa = self.compute_hyperparameters()
        if self.sigma_delta is None:
            self.sigma_delta = self.kernel_sigma
        self.use_cuda_graph = use_cuda_graph
        self._labels_dev = self._images_dev = self._uniq = None
        self._graph_step = None

    @property
    def device(self):
        return self.model.device

    # ------------------------------------------------------------------ hyper-parameters (trainer.py:173-307)
    def _labels_2d(self):
        lab = np.asarray(self.train_labels, dtype=np.float64)
        return lab.reshape(len(lab), -1)

    def compute_hyperparameters(self):
        """sigma_delta, kappa from the training labels: rule of thumb (1.06 std N^-1/5; largest gap between consecutive
        sorted unique labels) or a percentile of the pairwise distances (trainer.py:173-252), vectorised."""
        lab = self._labels_2d()
        hard = self.vicinity_type in ("hv", "shv")
        if self.hyperparameter == "rule_of_thumb":
            std = np.std(lab, axis=0) if lab.shape[1] > 1 else np.std(lab)
            sigma_delta = 1.06 * std * len(lab) ** (-1 / 5)
            uniq = np.unique(lab, axis=0)
            if len(uniq) > 1:
                order = np.lexsort([uniq[:, i] for i in range(uniq.shape[1] - 1, -1, -1)])
                srt = uniq[order]
                kappa_base = float(np.linalg.norm(srt[1:] - srt[:-1], axis=1).max())
                kappa = kappa_base if hard else 1 / kappa_base ** 2
            else:
                kappa = 0.01 if hard else 10000
        else:
            t = torch.from_numpy(lab)
            iu = torch.triu_indices(len(t), len(t), 1)
            d = compute_distance(t[iu[0]], t[iu[1]], self.distance).numpy()
            kappa = float(np.percentile(d, self.percentile))
            sigma_delta = kappa / 3
            if not hard:
                kappa = 1 / kappa ** 2
        return sigma_delta, kappa

    def compute_adaptive_params(self, batch_labels: torch.Tensor):
        """Per-batch sigma_delta / kappa (trainer.py:254-306) as device tensors: no host round trip, graph-capturable."""
        y = batch_labels.view(len(batch_labels), -1).float()
        hard = self.vicinity_type in ("hv", "shv")
        n = len(y)
        iu = torch.triu_indices(n, n, 1, device=y.device)
        if self.hyperparameter == "rule_of_thumb":
            sigma_delta = 1.06 * y.std(dim=0, unbiased=False) * n ** (-1 / 5)
            kappa_base = torch.sqrt(((y[iu[0]] - y[iu[1]]) ** 2).sum(-1)).min()
            kappa = kappa_base if hard else 1 / kappa_base ** 2
        else:
            d = compute_distance(y[iu[0]], y[iu[1]], self.distance)
            kappa = torch.quantile(d, self.percentile / 100.0)
            sigma_delta = kappa / 3
            if not hard:
                kappa = 1 / kappa ** 2
        return sigma_delta, kappa

    # ------------------------------------------------------------------ checkpoints (trainer.py:488-535 layout)
    def save(self, milestone):
        """Main process only (trainer.py:489-490), with a barrier so that no replica races ahead of the file."""
        if self.is_main:
            data = {"step": self.step, "model": self.model.state_dict(), "opt": self.opt.state_dict(),
                    "ema": self.ema.state_dict(), "scaler": None}
            torch.save(data, str(self.results_folder / f"model-{milestone}.pt"))
        ccdm_dist.barrier()

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

    # ------------------------------------------------------------------ device-resident data
    def _train_labels_dev(self):
        if self._labels_dev is None:
            t = torch.from_numpy(np.asarray(self.train_labels)).float().to(self.device)
            self._labels_dev = t.view(len(t), -1)
        return self._labels_dev

    def _unique_labels_dev(self):
        if self._uniq is None:
            self._uniq = torch.unique(self._train_labels_dev(), dim=0)     # trainer.py:313 (once: np.unique per batch there)
        return self._uniq

    def _train_images_dev(self):
        """The whole dataset as uint8 on the device (RC-49 64x64: 0.54 GB; UTKFace 192x192: 1.6 GB of 180 GB)."""
        if self._images_dev is None:
            arr = np.asarray(self.train_images)
            t = torch.from_numpy(np.ascontiguousarray(arr))
            if t.dtype != torch.uint8:
                t = t.round().clamp(0, 255).to(torch.uint8)
            self._images_dev = t.to(self.device)
        return self._images_dev

    # ------------------------------------------------------------------ batch construction (trainer.py:308-482,560-700)
    def _sigma_tensor(self, sigma):
        if torch.is_tensor(sigma):
            return sigma.to(self.device).float()
        # uploaded once: a pageable host->device copy is not allowed inside a CUDA-graph capture
        arr = np.asarray(sigma, dtype=np.float32)
        key = arr.tobytes()
        cache = self.__dict__.setdefault("_sigma_cache", {})
        if key not in cache:
            cache[key] = torch.from_numpy(arr.copy()).to(self.device)
        return cache[key]

    def sample_target_labels(self, batch_size):
        """trainer.py:308-315 + :575-580 / :630-636: unique training labels drawn with replacement, plus N(0, sigma_delta)."""
        uniq = self._unique_labels_dev()
        base = uniq[torch.randint(0, len(uniq), (batch_size,), device=self.device)]
        sigma, kappa = self.sigma_delta, self.kappa
        if self.adaptive_slicing and self.vicinity_type in ("shv", "ssv"):
            sigma, kappa = self.compute_adaptive_params(base)              # trainer.py:577-581
        return base + torch.randn_like(base) * self._sigma_tensor(sigma), kappa

    def sample_real_indices(self, target_labels: torch.Tensor, kappa=None) -> torch.Tensor:
        """One real sample per target label.  hv / sv (trainer.py:420-459): uniformly among the samples whose label is within
        ``kappa`` (``distance``) of the target (sv: among all), uniformly at random over the whole set when that vicinity is
        empty.  shv / ssv (trainer.py:317-418): samples whose projection on a random direction lies within
        ``kappa*|v|``; among the (up to) ten samples matched by the most projections one is drawn uniformly; nearest
        neighbour when nothing matches.  One [B, N] matrix instead of B host round trips; no host synchronisation."""
        kappa = self.kappa if kappa is None else kappa
        lab = self._train_labels_dev()                                    # [N, D]
        tgt = target_labels.view(len(target_labels), -1).to(lab)          # [B, D]
        B, N = len(tgt), len(lab)
        if self.vicinity_type in ("shv", "ssv"):
            v = generate_random_vectors(self.vector_type, lab.shape[1], self.num_projections, lab.device).float()
            vn = F.normalize(v, dim=1)
            d = ((tgt @ vn.t())[:, None, :] - (lab @ vn.t())[None, :, :]).abs()          # [B, N, P]
            cnt = (d <= (kappa * v.norm(dim=1))[None, None, :]).sum(-1)                  # projections matched
            nmatch = (cnt > 0).sum(1)
            key = cnt.to(torch.int64) * (N + 1) - torch.arange(N, device=lab.device)[None, :]
            k = min(10, N)
            topi = torch.topk(key, k, dim=1).indices                                     # most matches first, then lower index
            kvalid = nmatch.clamp(max=k)
            choice = (torch.rand(B, device=lab.device) * kvalid).long().clamp(min=0)
            choice = torch.minimum(choice, (kvalid - 1).clamp(min=0))
            pick = topi.gather(1, choice[:, None]).squeeze(1)
            if lab.shape[1] > 1:
                nearest = torch.cdist(tgt, lab).argmin(1)
            else:
                nearest = (tgt - lab.t()).abs().argmin(1)
            return torch.where(nmatch > 0, pick, nearest)
        if self.vicinity_type == "hv":
            mask = compute_distance(lab[None, :, :], tgt[:, None, :], self.distance) <= kappa
            mask = mask | ~mask.any(1, keepdim=True)                      # empty vicinity: random over the whole set
        else:                                                             # sv: every sample is a candidate
            mask = torch.ones(B, N, dtype=torch.bool, device=lab.device)
        score = torch.rand(B, N, device=lab.device).masked_fill(~mask, -1.0)
        return score.argmax(1)

    def vicinal_weights(self, batch_labels: torch.Tensor, target_labels: torch.Tensor, kappa=None) -> torch.Tensor:
        """trainer.py:610-616 (sliced types: ones; the weighting happens inside p_losses) and :663-690 (hv: 1 inside the
        vicinity of the TARGET label, else 0; sv: exp(-nu dist^2), nu = 1/kappa^2)."""
        kappa = self.kappa if kappa is None else kappa
        B = len(batch_labels)
        if self.vicinity_type in ("shv", "ssv"):
            return torch.ones(B, device=self.device)
        d = compute_distance(batch_labels.view(B, -1), target_labels.view(B, -1), self.distance)
        if self.vicinity_type == "hv":
            return (d <= kappa).float()
        nu = 1.0 / (kappa ** 2)
        return torch.exp(-nu * d ** 2)

    def augmentation_bits(self, batch_size) -> torch.Tensor | None:
        """Per-sample augmentation code of ccdm_gather_augment_u8 (trainer.py:468-474, utils.py:164-211): UTKFace: horizontal
        flip with p = 1/2; Cell200: quarter turns k ~ U{0..3}, then horizontal flip, then vertical flip (p = 1/2 each)."""
        dev = self.device
        if self.data_name == "UTKFace":
            return ((torch.rand(batch_size, device=dev) > 0.5).to(torch.uint8) << 2)
        if self.data_name == "Cell200":
            k = torch.randint(0, 4, (batch_size,), device=dev, dtype=torch.uint8)
            hf = (torch.rand(batch_size, device=dev) > 0.5).to(torch.uint8) << 2
            vf = (torch.rand(batch_size, device=dev) < 0.5).to(torch.uint8) << 3
            return k | hf | vf
        return None

    def process_images(self, idx: torch.Tensor) -> torch.Tensor:
        """images[idx] -> augmentation -> /255, fp32 NCHW in [0, 1] (trainer.py:461-482).  CUDA: one kernel over the
        device-resident uint8 dataset; CPU (host-logic tests): the same arithmetic in torch."""
        imgs = self._train_images_dev()
        aug = self.augmentation_bits(len(idx))
        n, c, h, w = imgs.shape
        if imgs.is_cuda:
            out = torch.empty(len(idx), c, h, w, dtype=torch.float32, device=imgs.device)
            L.check(L.lib().ccdm_gather_augment_u8(imgs.data_ptr(), n, idx.contiguous().data_ptr(), L.ptr(aug), out.data_ptr(),
                                                   len(idx), c, h, w, torch.cuda.current_stream().cuda_stream),
                    "gather_augment_u8")
            return out
        x = imgs[idx.clamp(0, n - 1)].float()
        if aug is not None:
            k = (aug & 3)
            for q in (1, 2, 3):
                sel = (k == q)[:, None, None, None]
                x = torch.where(sel, torch.rot90(x, q, dims=(2, 3)), x)
            x = torch.where(((aug >> 2) & 1).bool()[:, None, None, None], x.flip(3), x)
            x = torch.where(((aug >> 3) & 1).bool()[:, None, None, None], x.flip(2), x)
        return x / 255.0

    def device_batch(self, fn_y2h):
        """One micro-batch, entirely from device ops (capturable): (images, labels, labels_emb, vicinal_weights, loss kwargs)."""
        B = self.batch_size
        kw = {}
        if self.vicinity_type in ("shv", "ssv", "hv", "sv"):
            tgt, kappa = self.sample_target_labels(B)
            idx = self.sample_real_indices(tgt, kappa)
            lab2 = self._train_labels_dev()[idx]
            weights = self.vicinal_weights(lab2, tgt, kappa)
            if self.vicinity_type in ("shv", "ssv"):
                kw = dict(vicinity_type=self.vicinity_type, kappa=kappa, vector_type=self.vector_type,
                          num_projections=self.num_projections)
        else:                                                             # trainer.py:701-706
            idx = torch.randint(0, len(self._train_labels_dev()), (B,), device=self.device)
            lab2 = self._train_labels_dev()[idx]
            weights = None
        images = self.process_images(idx)
        labels = lab2.view(-1) if np.asarray(self.train_labels).ndim == 1 else lab2
        return images, labels, fn_y2h(labels), weights, kw

    # ------------------------------------------------------------------ training loop (trainer.py:537-780)
    def _log_loss(self, total):
        if self.is_main and self.step % 500 == 0:
            log = os.path.join(self.results_folder, f"log_loss_niters{self.train_num_steps}.txt")
            with open(log, "a") as f:
                f.write(f"\r Step: {self.step}, Loss: {float(total):.4f}.")

    def _after_step(self, fn_y2h):
        self.step += 1
        if not self.is_main:
            return
        self.ema.update()
        if self.step != 0 and self.y_visual is not None and divisible_by(self.step, self.sample_every):
            self.save_preview(fn_y2h)
        if self.step != 0 and divisible_by(self.step, self.save_every):
            self.ema.ema_model.eval()

    def save_preview(self, fn_y2h):
        """trainer.py:742-770: a DDIM sample grid of the EMA model for the fixed visualisation labels."""
        self.ema.ema_model.eval()
        y = torch.as_tensor(np.asarray(self.y_visual), dtype=torch.float32, device=self.device)
        with torch.inference_mode():
            img = self.ema.ema_model.ddim_sample(labels_emb=fn_y2h(y), labels=y,
                                                 shape=(y.shape[0], self.channels, self.image_size, self.image_size),
                                                 cond_scale=self.cond_scale_visual)
        img = img.detach().cpu()
        if img.min() < 0 or img.max() > 1:
            print(f"\r Generated images are out of range. (min={img.min()}, max={img.max()})")
        img = torch.clip(img, 0, 1)
        try:
            from torchvision import utils as tvu
        except Exception:                                                 # pragma: no cover
            warnings.warn("torchvision is not importable: preview grid not written")
            return img
        tvu.save_image(img, str(self.results_folder / f"sample_{self.step}.png"), nrow=self.nrow_visual, normalize=False,
                       padding=1)
        return img

    def train(self, fn_y2h):
        if self.is_main:
            log = os.path.join(self.results_folder, f"log_loss_niters{self.train_num_steps}.txt")
            with open(log, "a") as f:
                f.write("\n===================================================================================================")
        # capturing costs a few eager steps: worth it only for a run of some length (the capture's warm-up steps ARE optimizer
        # steps on real batches and are counted)
        graphed = (self.on_cuda and self.use_cuda_graph and isinstance(self.opt, FusedAdam)
                   and (self._graph_step is not None or self.train_num_steps - self.step >= 8))
        if graphed and self._graph_step is None:
            from .train_graph import GraphedTrainStep
            self._graph_step = GraphedTrainStep(self.model, self.opt, batch_fn=lambda: self.device_batch(fn_y2h),
                                                accumulate=self.gradient_accumulate_every,
                                                max_grad_norm=self.max_grad_norm)
            # the capture warm-up ran real optimizer steps on real batches: count them
            for _ in range(self._graph_step.warmup_steps):
                if self.step < self.train_num_steps:
                    self._after_step(fn_y2h)
                    self._save_if_due()
        while self.step < self.train_num_steps:
            if graphed:
                total = self._graph_step.replay()                          # device scalar; read only when it is logged
            else:
                total = self._eager_step(fn_y2h)
            if self.step % 500 == 0:
                self._log_loss(total)
            self._after_step(fn_y2h)
            self._save_if_due()

    def _save_if_due(self):
        if self.step != 0 and divisible_by(self.step, self.save_every):
            self.save(self.step)

    def _eager_step(self, fn_y2h):
        total = torch.zeros((), device=self.device)
        for mb in range(self.gradient_accumulate_every):
            if isinstance(self.opt, FusedAdam):
                self.opt.arm_early_bucket(mb == self.gradient_accumulate_every - 1)
            images, labels, emb, weights, kw = self.device_batch(fn_y2h)
            loss = self.model(images, labels_emb=emb, labels=labels, vicinal_weights=weights, **kw)
            loss = loss / self.gradient_accumulate_every
            total += loss.detach()
            loss.backward()
        if isinstance(self.opt, FusedAdam):
            self.opt.all_reduce_gradients()                       # one collective over the flat gradient buffer; clip is fused
        else:
            ccdm_dist.all_reduce_gradients(list(self.model.parameters()))
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
        self.opt.step()
        self.opt.zero_grad()
        return total

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
