"""Conditional UNet denoiser with the reference's constructor, call signatures and checkpoint layout.

Drop-in for ``CCDM_unified/models/unet.py:244-455`` (``Unet``): same constructor keywords, same ``forward`` /
``forward_with_cond_scale`` signatures, same ``state_dict`` keys, shapes and default initialisation order (so a
shared ``torch.manual_seed`` yields identical random weights).  The sub-modules below only *own parameters*; the
arithmetic runs in hand-written sm_100a kernels driven by :mod:`ccdm_b200.engine`.  There is no PyTorch fallback:
calling the model on a CPU tensor raises.
"""
from __future__ import annotations

import torch
from torch import nn

LIN_ATTN_HEADS, LIN_ATTN_DIM_HEAD = 4, 32      # unet.py:190 defaults, never overridden (:325,:340)


def _holder_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} only holds parameters; run the whole ccdm_b200.Unet instead")


class RMSNorm(nn.Module):                       # unet.py:83-89
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
    forward = _holder_forward


class Block(nn.Module):                         # unet.py:136-152
    def __init__(self, dim, dim_out):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = RMSNorm(dim_out)
    forward = _holder_forward


class ResnetBlock(nn.Module):                   # unet.py:154-187
    def __init__(self, dim, dim_out, *, time_emb_dim, cond_emb_dim=0):
        super().__init__()
        self.dim, self.dim_out = dim, dim_out
        self.tc_mlp = nn.Sequential(nn.SiLU(), nn.Linear(int(time_emb_dim) + int(cond_emb_dim), dim_out * 2))
        self.block1 = Block(dim, dim_out)
        self.block2 = Block(dim_out, dim_out)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    forward = _holder_forward


class LinearAttention(nn.Module):               # unet.py:189-216
    def __init__(self, dim, heads=LIN_ATTN_HEADS, dim_head=LIN_ATTN_DIM_HEAD):
        super().__init__()
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), RMSNorm(dim))
    forward = _holder_forward


class Attention(nn.Module):                     # unet.py:218-240
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)
    forward = _holder_forward


class PreNorm(nn.Module):                       # unet.py:91-99
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = RMSNorm(dim)
    forward = _holder_forward


class Residual(nn.Module):                      # unet.py:66-72
    def __init__(self, fn):
        super().__init__()
        self.fn = fn
    forward = _holder_forward


class SinusoidalPosEmb(nn.Module):              # unet.py:102-115 (parameter-free; evaluated by ccdm_time_features)
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
    forward = _holder_forward


class Unet(nn.Module):
    def __init__(self, dim, embed_input_dim=128, cond_drop_prob=0.5, init_dim=None, out_dim=None,
                 dim_mults=(1, 2, 4, 8), in_channels=3, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, attn_dim_head=32, attn_heads=4,
                 precision="bf16"):
        """``precision`` (beyond the reference's signature) selects the inference tier: "bf16" (default) or "fp16" -- IEEE
        binary16 storage of activations and kernel weights, i.e. TF32's 10-bit mantissa on the same tcgen05 kernels with fp32
        accumulation (BASELINE.json north_star: the fp32/TF32 tolerance tier).  Training always uses the bf16 kernels."""
        super().__init__()
        if precision not in ("bf16", "fp16"):
            raise ValueError(f"precision must be 'bf16' or 'fp16' (got {precision!r})")
        self.precision = precision
        if learned_sinusoidal_cond or random_fourier_features:
            # GaussianDiffusion refuses such a model anyway (diffusion.py:135)
            raise NotImplementedError("random / learned sinusoidal time embeddings are outside the hot path")
        self.dim, self.dim_mults = dim, tuple(dim_mults)
        self.embed_input_dim = embed_input_dim
        self.in_channels = in_channels
        self.cond_drop_prob = cond_drop_prob
        self.random_or_learned_sinusoidal_cond = False
        self.attn_heads, self.attn_dim_head = attn_heads, attn_dim_head

        init_dim = dim if init_dim is None else init_dim
        self.init_dim = init_dim
        self.init_conv = nn.Conv2d(in_channels, init_dim, 7, padding=3)
        widths = [init_dim] + [dim * m for m in dim_mults]
        self.in_out = list(zip(widths[:-1], widths[1:]))
        emb = dim * 4
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, emb), nn.GELU(), nn.Linear(emb, emb))
        self.cond_mlp_1 = nn.Sequential(nn.Linear(embed_input_dim, dim), nn.BatchNorm1d(dim), nn.ReLU())
        self.null_cond_emb = nn.Parameter(-1 * torch.abs(torch.randn(dim)), requires_grad=True)
        self.cond_mlp_2 = nn.Sequential(nn.Linear(dim, emb), nn.BatchNorm1d(emb), nn.ReLU())

        rb = lambda i, o: ResnetBlock(i, o, time_emb_dim=emb, cond_emb_dim=emb)
        self.downs, self.ups = nn.ModuleList([]), nn.ModuleList([])
        n = len(self.in_out)
        for k, (ci, co) in enumerate(self.in_out):
            last = k >= n - 1
            self.downs.append(nn.ModuleList([
                rb(ci, ci), rb(ci, ci), Residual(PreNorm(ci, LinearAttention(ci))),
                nn.Conv2d(ci, co, 4, 2, 1) if not last else nn.Conv2d(ci, co, 3, padding=1)]))
        mid = widths[-1]
        self.mid_block1 = rb(mid, mid)
        self.mid_attn = Residual(PreNorm(mid, Attention(mid, dim_head=attn_dim_head, heads=attn_heads)))
        self.mid_block2 = rb(mid, mid)
        for k, (ci, co) in enumerate(reversed(self.in_out)):
            last = k == n - 1
            up = (nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(co, ci, 3, padding=1))
                  if not last else nn.Conv2d(co, ci, 3, padding=1))
            self.ups.append(nn.ModuleList([rb(co + ci, co), rb(co + ci, co),
                                           Residual(PreNorm(co, LinearAttention(co))), up]))
        self.out_dim = out_dim if out_dim is not None else in_channels * (1 if not learned_variance else 2)
        self.final_res_block = rb(init_dim * 2, init_dim)
        self.final_conv = nn.Conv2d(init_dim, self.out_dim, 1)
        self._engine = None                     # lazily built ccdm_b200.engine.UnetEngine (never copied / saved)

    # ------------------------------------------------------------------ engine plumbing
    def __deepcopy__(self, memo):               # EMA deep-copies the diffusion wrapper (ema_pytorch.py:69-74)
        eng, self._engine = self._engine, None
        try:
            cls = type(self)
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        return st

    def engine(self):
        from .engine import UnetEngine
        dev = self.init_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ccdm_b200.Unet runs on sm_100a only: move the model to a CUDA device "
                               "(there is no CPU fallback)")
        if self._engine is None or self._engine.device != dev:
            self._engine = UnetEngine(self)
        return self._engine

    # ------------------------------------------------------------------ reference API
    def forward(self, x, timesteps, labels_emb, cond_drop_prob=None, keep_mask=None, return_bottleneck=False):
        """unet.py:382-455.  ``keep_mask`` is accepted and ignored exactly like the reference (the branch that
        would use it is unreachable there): a Bernoulli(1-p) mask is drawn here instead."""
        if return_bottleneck:
            raise NotImplementedError("return_bottleneck is only used by the out-of-scope DMD trainer")
        p = self.cond_drop_prob if cond_drop_prob is None else cond_drop_prob
        b = x.shape[0]
        mask = None
        if p > 0:
            keep = 1 - p
            if keep == 1:
                mask = torch.ones(b, device=x.device, dtype=torch.bool)
            elif keep == 0:
                mask = torch.zeros(b, device=x.device, dtype=torch.bool)
            else:
                mask = torch.zeros((b,), device=x.device).float().uniform_(0, 1) < keep
            self.keep_mask = mask
        if self.training and torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()):
            # training step: autograd graph over the CUDA kernels (ccdm_b200/train.py); BatchNorm1d uses batch statistics
            from .train import unet_train_forward
            return unet_train_forward(self, x, timesteps, labels_emb, mask)
        return self.engine().forward(x, timesteps, labels_emb, mask)

    def forward_with_cond_scale(self, *args, cond_scale=1.0, rescaled_phi=0.0, remove_parallel_component=True,
                                keep_parallel_frac=0.0, **kwargs):
        """unet.py:350-380.  Returns ``(guided, null)``; for ``cond_scale == 1`` the bare conditional output, as
        the reference does."""
        x, t, labels_emb = args[0], args[1], args[2]
        eng = self.engine()
        if cond_scale == 1:
            return eng.forward(x, t, labels_emb, None)
        if self.training:
            cond = eng.forward(x, t, labels_emb, None).clone()
            null = eng.forward(x, t, labels_emb, torch.zeros(x.shape[0], device=x.device, dtype=torch.bool)).clone()
        else:
            cond, null = eng.forward_pair(x, t, labels_emb)     # one 2B batch: BatchNorm1d uses running stats
        guided = eng.cfg_combine(cond, null, cond_scale, rescaled_phi, remove_parallel_component, keep_parallel_frac)
        return guided, null
