"""Exponential moving average of a model, with the interface Trainer uses (CCDM_unified/ema_pytorch.py:18-181):
``EMA(model, beta, update_after_step, update_every)``, ``.ema_model``, ``.update()``, ``state_dict()``.

The lerp over all parameters is one launch of ``ccdm_multi_lerp`` over a pointer table (HBM-bound, every
``update_every`` steps); CPU models (tests) use ``torch._foreach_lerp_``."""
import copy

import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, ema_model=None, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0,
                 power=2 / 3, min_value=0.0):
        super().__init__()
        self.beta = beta
        self.online_model = model
        self.ema_model = ema_model if ema_model is not None else copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.update_every, self.update_after_step = update_every, update_after_step
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.register_buffer("initted", torch.tensor([False]))
        self.register_buffer("step", torch.tensor([0]))
        # host shadows of the two buffers: update() must not read the device every training step (a .item() is a full
        # host-device synchronisation); they are re-read after load_state_dict
        self._step_host = None
        self._initted_host = None

    def _sync_host_state(self):
        if self._step_host is None:
            self._step_host = int(self.step.item())
            self._initted_host = bool(self.initted.item())

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._step_host = None
        return out

    def _pairs(self):
        on = dict(self.online_model.named_parameters())
        return [(p, on[n]) for n, p in self.ema_model.named_parameters() if p.dtype.is_floating_point]

    def _buffer_pairs(self):
        on = dict(self.online_model.named_buffers())
        return [(b, on[n]) for n, b in self.ema_model.named_buffers()]

    def copy_params_from_model_to_ema(self):
        with torch.no_grad():
            for e, o in self._pairs() + self._buffer_pairs():
                e.copy_(o)

    def get_current_decay(self):                     # ema_pytorch.py:124-131
        self._sync_host_state()
        epoch = max(self._step_host - self.update_after_step - 1, 0)
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        return 0.0 if epoch <= 0 else min(max(value, self.min_value), self.beta)

    @torch.no_grad()
    def update(self):                                # ema_pytorch.py:133-178
        self._sync_host_state()
        step = self._step_host
        self._step_host += 1
        self.step += 1
        if (step % self.update_every) != 0:
            return
        if step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if not self._initted_host:
            self.copy_params_from_model_to_ema()
            self.initted.fill_(True)
            self._initted_host = True
        w = 1.0 - self.get_current_decay()
        pairs = self._pairs()
        if pairs and all(e.is_cuda and e.dtype == torch.float32 for e, _ in pairs):
            # one launch over a pointer table (ccdm_multi_lerp); rebuilt only if a tensor moved
            from .optim import MultiLerp
            ml = getattr(self, "_ml", None)
            ptrs = tuple(e.data_ptr() for e, _ in pairs) + tuple(o.data_ptr() for _, o in pairs)
            if ml is None or self._ml_ptrs != ptrs:
                ml = MultiLerp([e for e, _ in pairs], [o for _, o in pairs])
                object.__setattr__(self, "_ml", ml)
                object.__setattr__(self, "_ml_ptrs", ptrs)
            ml(w)
        else:
            torch._foreach_lerp_([e for e, _ in pairs], [o for _, o in pairs], w)
        for e, o in self._buffer_pairs():
            if e.dtype.is_floating_point:
                e.lerp_(o.to(e.dtype), w)
            else:
                e.copy_(o)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
