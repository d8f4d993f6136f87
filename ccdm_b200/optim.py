"""Fused optimizer step: ``FusedAdam`` = ``clip_grad_norm_`` + ``torch.optim.Adam.step``
of CCDM_unified/trainer.py:137,724,733-734 in three kernel launches.

Gradients and both Adam moments live in flat fp32 buffers; every ``p.grad`` is a view into the gradient buffer, so

* ``zero_grad`` is one memset, the data-parallel exchange is ONE all-reduce over the flat buffer (no bucketing
  copies), the global gradient norm is one reduction, and
* all pointers are stable from step to step, which is what lets ``train_graph.GraphedTrainStep`` capture the step.

Arithmetic follows ``torch.optim.Adam`` (no amsgrad): bias corrections from a device-resident step count,
``p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)``; clipping follows ``torch.nn.utils.clip_grad_norm_``
(``coef = min(1, max_norm / (norm + 1e-6))``).  ``state_dict`` / ``load_state_dict`` use torch Adam's layout, so the
"opt" entry of a reference checkpoint (trainer.py:488-535) loads.  No fallback: parameters must be CUDA fp32.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _lib as L

_CHUNK = 1 << 16


def _chunk_tables(tensors, offsets, device):
    ptrs, offs, ns = [], [], []
    for t, off in zip(tensors, offsets):
        n = t.numel()
        for c0 in range(0, n, _CHUNK):
            ptrs.append(t.data_ptr() + 4 * c0)
            offs.append(off + c0)
            ns.append(min(_CHUNK, n - c0))
    return (torch.tensor(ptrs, dtype=torch.int64, device=device), torch.tensor(offs, dtype=torch.int64, device=device),
            torch.tensor(ns, dtype=torch.int32, device=device))


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None,
                 early_params: Optional[Iterable[torch.nn.Parameter]] = None):
        """``early_params``: parameters whose gradients are complete early in the backward pass (the UNet's decoder half,
        ``dist.early_gradient_params``).  They are laid out at the END of the flat buffers, so that their all-reduce is one
        contiguous NCCL call that ``launch_early_bucket`` starts on a side stream while the encoder's backward still runs
        (overlap of the gradient exchange with the backward)."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam keeps one flat buffer: pass a single parameter group")
        self.max_grad_norm = max_grad_norm
        all_ps = list(self.param_groups[0]["params"])
        # torch.optim.Adam numbers its state by position in the FULL parameter group (frozen parameters included)
        self._index = [i for i, p in enumerate(all_ps) if p.requires_grad]
        ps = [p for p in all_ps if p.requires_grad]
        if not ps:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = ps[0].device
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.device == dev and p.is_contiguous()):
                raise RuntimeError("FusedAdam needs contiguous CUDA fp32 parameters on one device (no CPU fallback)")
        self._ps = ps
        early = {id(p) for p in early_params} if early_params is not None else set()
        self._offsets, total = [0] * len(ps), 0
        for want_early in (False, True):                           # flat layout: [late ... | early ...]
            if want_early:
                self.early_start = total
            for i, p in enumerate(ps):
                if (id(p) in early) == want_early:
                    self._offsets[i] = total
                    total += (p.numel() + 3) // 4 * 4              # 16-byte aligned views
        self._side = None
        self._early_armed = self._early_launched = False
        with torch.inference_mode(False):
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
            self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
            self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
            self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
            self._tab = _chunk_tables(ps, self._offsets, dev)
        self._ptrs = [p.data_ptr() for p in ps]
        self._attach()

    # ------------------------------------------------------------------ gradient views
    def _view(self, i):
        p, off = self._ps[i], self._offsets[i]
        return self.flat_grad[off:off + p.numel()].view_as(p)

    def _attach(self):
        for i, p in enumerate(self._ps):
            v = self._view(i)
            if p.grad is None:
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)                                    # someone replaced .grad: keep its value
                p.grad = v

    def zero_grad(self, set_to_none: bool = False):
        """One memset; ``set_to_none`` is ignored on purpose (the views must stay attached)."""
        self.flat_grad.zero_()
        self._attach()

    def arm_early_bucket(self, on: bool = True):
        """Called by the step driver before the backward of the LAST micro-batch of an optimizer step."""
        self._early_armed = bool(on) and self.early_start < self.flat_grad.numel()

    def launch_early_bucket(self):
        """Runs inside the backward pass (a tensor hook on the decoder's input, train.py) once every early parameter's
        gradient is in the flat buffer: their all-reduce starts on a side stream and overlaps the rest of the backward."""
        if not (self._early_armed and dist.is_initialized() and dist.get_world_size() > 1):
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.flat_grad[self.early_start:], op=dist.ReduceOp.SUM)
        self._early_armed, self._early_launched = False, True

    def all_reduce_gradients(self):
        """Data-parallel exchange: NCCL all-reduce over the flat buffer (one call, or the late segment
        here + the early segment already in flight on the side stream), then the mean."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            if self._early_launched:
                dist.all_reduce(self.flat_grad[:self.early_start], op=dist.ReduceOp.SUM)
                torch.cuda.current_stream().wait_stream(self._side)
                self._early_launched = False
            else:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.flat_grad.div_(dist.get_world_size())
        self._early_armed = False

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the gradients seen by the last ``step`` (before clipping); device tensor."""
        return self.grad_sumsq.sqrt().float()

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        ptrs = [p.data_ptr() for p in self._ps]
        if ptrs != self._ptrs:                                     # a parameter's storage moved (.to(), load with assign)
            for p in self._ps:
                if not (p.is_cuda and p.device == self.flat_grad.device and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam: a parameter left the device / dtype the optimizer was built for")
            with torch.inference_mode(False):
                self._tab = _chunk_tables(self._ps, self._offsets, self.flat_grad.device)
            self._ptrs = ptrs
        self._attach()
        g = self.param_groups[0]
        ptrs, offs, ns = self._tab
        clip = self.max_grad_norm is not None
        L.check(L.lib().ccdm_fused_adam(ptrs.data_ptr(), offs.data_ptr(), ns.data_ptr(), ns.numel(),
                                        self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        self.flat_grad.numel(), self.step_count.data_ptr(),
                                        self.grad_sumsq.data_ptr() if clip else None,
                                        float(self.max_grad_norm or 0.0), float(g["lr"]), float(g["betas"][0]),
                                        float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                                        torch.cuda.current_stream().cuda_stream), "fused_adam")
        self.mark_params_modified()
        return loss

    def mark_params_modified(self):
        """The kernel writes the parameters through raw pointers, which autograd's version counters do not see; the
        engines' packed bf16 weight caches are keyed on ``(data_ptr, _version)`` (engine.WeightStore.stamp), so bump
        them here.  ``GraphedTrainStep`` calls this after every graph replay (no Python ``step`` runs there)."""
        torch._C._increment_version(self._ps)

    # ------------------------------------------------------------------ checkpoints in torch.optim.Adam's layout
    def state_dict(self):
        state = {}
        for j, p in enumerate(self._ps):
            i, off, n = self._index[j], self._offsets[j], p.numel()
            state[i] = {"step": self.step_count[0].clone(), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        grp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        grp["params"] = list(range(len(self.param_groups[0]["params"])))
        return {"state": state, "param_groups": [grp]}

    def load_state_dict(self, sd):
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]
        for j, p in enumerate(self._ps):
            i = self._index[j]
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            off, n = self._offsets[j], p.numel()
            if st["exp_avg"].numel() != n or st["exp_avg_sq"].numel() != n:
                raise ValueError(f"FusedAdam.load_state_dict: state {i} has {st['exp_avg'].numel()} elements, the parameter {n}")
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count.fill_(float(st["step"]))


class MultiLerp:
    """``dst[i].lerp_(src[i], w)`` for lists of CUDA fp32 tensors in one launch (the EMA update)."""

    def __init__(self, dst, src):
        dev = dst[0].device
        dp, sp, ns = [], [], []
        for d, s in zip(dst, src):
            if not (d.is_cuda and s.is_cuda and d.dtype == s.dtype == torch.float32 and d.is_contiguous()
                    and s.is_contiguous() and d.numel() == s.numel()):
                raise RuntimeError("MultiLerp needs matching contiguous CUDA fp32 tensors")
            for c0 in range(0, d.numel(), _CHUNK):
                dp.append(d.data_ptr() + 4 * c0)
                sp.append(s.data_ptr() + 4 * c0)
                ns.append(min(_CHUNK, d.numel() - c0))
        self.key = tuple(dp) + tuple(sp)
        self._dst = list(dst)
        with torch.inference_mode(False):
            self.dp = torch.tensor(dp, dtype=torch.int64, device=dev)
            self.sp = torch.tensor(sp, dtype=torch.int64, device=dev)
            self.ns = torch.tensor(ns, dtype=torch.int32, device=dev)
            self.w = torch.zeros(1, dtype=torch.float32, device=dev)

    def __call__(self, weight: float):
        self.w.fill_(float(weight))
        L.check(L.lib().ccdm_multi_lerp(self.dp.data_ptr(), self.sp.data_ptr(), self.ns.data_ptr(), self.ns.numel(),
                                        self.w.data_ptr(), torch.cuda.current_stream().cuda_stream), "multi_lerp")
        # raw-pointer writes are invisible to autograd's version counters: invalidate the packed-weight caches
        torch._C._increment_version(self._dst)
