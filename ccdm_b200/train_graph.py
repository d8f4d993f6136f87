"""One whole optimizer step -- q_sample, UNet forward, vicinal loss, backward, gradient all-reduce, clip, Adam --
captured once into a CUDA graph and replayed (CCDM_unified/trainer.py:557-740 is the loop it replaces).

The eager training path (ccdm_b200/train.py) issues ~1000 small launches per step from Python; at B200 speeds the step
is then bound by the host.  Captured, the step costs one graph launch.  Requirements that make the capture legal:

* static shapes: the batch is copied into fixed device buffers before every replay;
* no host synchronisation inside the step: ``GaussianDiffusion.graph_safe_rng`` draws the null-row noise for the whole
  batch instead of first asking the host how many null rows there are (diffusion.py:532-536 does a ``torch.where`` +
  ``len``); the noise is then distributed identically but consumes the RNG stream differently from the reference;
* the optimizer keeps its step count on the device (``capturable=True``).

Random draws (timesteps, masks, noise) use PyTorch's graph-safe Philox generator, so every replay sees fresh numbers.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import dist as ccdm_dist
from .optim import FusedAdam


class GraphedTrainStep:
    def __init__(self, diffusion, optimizer: torch.optim.Optimizer, images: Optional[torch.Tensor] = None,
                 labels: Optional[torch.Tensor] = None, labels_emb: Optional[torch.Tensor] = None, *,
                 loss_kwargs: Optional[dict] = None, max_grad_norm: Optional[float] = 1.0, warmup: int = 3,
                 batch_fn: Optional[Callable] = None, accumulate: int = 1):
        """Two ways to feed the step:

        * ``images`` / ``labels`` / ``labels_emb``: example batches (their shapes are frozen); ``__call__`` copies a new batch
          into the static buffers and replays;
        * ``batch_fn() -> (images, labels, labels_emb, vicinal_weights, loss_kwargs)``: the batch is BUILT inside the captured
          step from device ops only (``Trainer.device_batch``: target labels, vicinity search, gather + augmentation), so a
          training step is one ``replay()`` with no host work at all; ``accumulate`` micro-batches are captured back to back
          (``gradient_accumulate_every`` of trainer.py:560-720), their losses summed into ``self.loss``.
        """
        self.gd, self.opt = diffusion, optimizer
        self.kw = dict(loss_kwargs or {})
        self.max_grad_norm = max_grad_norm
        self.batch_fn, self.accumulate = batch_fn, int(accumulate)
        self.params = [p for p in diffusion.parameters() if p.requires_grad]
        self.fused = isinstance(optimizer, FusedAdam)
        dev = self.params[0].device
        if self.fused:
            optimizer.max_grad_norm = max_grad_norm               # clipping is part of the fused step
            ccdm_dist.wire_overlap(diffusion, optimizer)          # early-bucket all-reduce under the encoder's backward
        else:
            for grp in optimizer.param_groups:
                grp["capturable"] = True
                if "foreach" in grp and grp["foreach"] is None:
                    grp["foreach"] = True
            for st in optimizer.state.values():
                if "step" in st and torch.is_tensor(st["step"]) and not st["step"].is_cuda:
                    st["step"] = st["step"].to(dev)
        if batch_fn is None:
            self.images = images.detach().clone()
            self.labels = labels.detach().clone()
            self.labels_emb = labels_emb.detach().clone()
            self.weights = torch.ones(images.shape[0], device=dev)
        self.loss = torch.zeros((), device=dev)
        self.graph = torch.cuda.CUDAGraph()
        diffusion.graph_safe_rng = True
        diffusion.train()

        self.warmup_steps = max(warmup, 1)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup_steps):   # allocator, schedule tables, optimizer state, autograd threads
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self._step()
        torch.cuda.synchronize()

    def _step(self):
        self.opt.zero_grad(set_to_none=True)
        total = None
        for mb in range(self.accumulate):
            if self.fused:
                self.opt.arm_early_bucket(mb == self.accumulate - 1)
            if self.batch_fn is not None:
                images, labels, emb, weights, kw = self.batch_fn()
                kw = dict(self.kw, **kw)
            else:
                images, labels, emb, weights, kw = self.images, self.labels, self.labels_emb, self.weights, self.kw
            loss = self.gd(images, labels_emb=emb, labels=labels, vicinal_weights=weights, **kw)
            if self.accumulate > 1:
                loss = loss / self.accumulate
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if self.fused:
            self.opt.all_reduce_gradients()
        else:
            ccdm_dist.all_reduce_gradients(self.params)
            if self.max_grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()
        self.loss.copy_(total)

    def replay(self) -> torch.Tensor:
        """One optimizer step on a batch built inside the graph (``batch_fn`` mode); returns the device loss scalar."""
        self.graph.replay()
        # the replayed optimizer kernel wrote the parameters without any Python-side version bump: the eval engines'
        # packed-weight caches (keyed on _version) must see the change
        torch._C._increment_version(self.params)
        return self.loss

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, labels_emb: torch.Tensor) -> torch.Tensor:
        """Runs one optimizer step on the given batch; returns the (device, 0-dim) loss of that step."""
        assert self.batch_fn is None, "this step builds its own batches: call replay()"
        self.images.copy_(images, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.labels_emb.copy_(labels_emb, non_blocking=True)
        return self.replay()
