"""One-process-per-GPU plumbing for the batch-sharded sampler and the benchmark.

Sampling shards by sample: every rank owns a contiguous slice of the label batch, a full UNet replica and its own
CUDA graph; there is NO data-path collective.  torch.distributed is used only for the launch barrier, the
max-over-ranks timing and an optional gather of finished uint8 images.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init(backend: str | None = None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for this rank (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, rank: int, world: int):
    lo, hi = shard_bounds(len(t), rank, world)
    return t[lo:hi]


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's value -- every multi-GPU timing is reported as the max over ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_images_u8(img_u8: torch.Tensor, dst: int = 0):
    """Optional epilogue of a sharded sampling job: concatenate per-rank uint8 images on rank ``dst``."""
    if not dist.is_initialized():
        return img_u8
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=img_u8.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([img_u8.shape[0]], dtype=torch.int64, device=img_u8.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(img_u8.shape[1:]), dtype=img_u8.dtype, device=img_u8.device)
    pad[: img_u8.shape[0]] = img_u8
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if dist.get_rank() != dst:
        return None
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)])


# ------------------------------------------------------------------------------------------- data-parallel training
# training shards by micro-batch; the one exchange step is the gradient all-reduce.  The reference
# gets it from accelerate/DDP (trainer.py:121-128,724); here it is an explicit bucketed all-reduce over NCCL (NVLS
# reduces inside the NVSwitch), issued after the backward of the last micro-batch.

def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every replica start from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def all_reduce_gradients(params, bucket_bytes: int = 64 << 20):
    """Average ``p.grad`` over the ranks in flat fp32 buckets (one collective per ~64 MiB instead of one per tensor)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def early_gradient_params(model):
    """Parameters of the UNet's decoder half (``ups``, ``final_res_block``, ``final_conv``; unet.py:330-348) except the
    ``tc_mlp`` Linears, whose gradients come out of the batched conditioning GEMM at the very end of the backward.  Their
    weight gradients are complete when the backward pass reaches the bottleneck -- ``FusedAdam(early_params=...)`` puts them
    in one contiguous segment and ``wire_overlap`` starts that segment's all-reduce from a hook at that point."""
    net = getattr(model, "unet", model)
    net = getattr(net, "module", net)
    out, skip = [], set()
    for name in ("ups", "final_res_block", "final_conv"):
        mod = getattr(net, name, None)
        if mod is None:
            return []
        for n, p in mod.named_parameters():
            if "tc_mlp" in n or id(p) in skip:
                continue
            skip.add(id(p))
            out.append(p)
    return out


def wire_overlap(model, opt) -> bool:
    """Lets the training forward (train.unet_train_forward) call ``opt.launch_early_bucket`` from a gradient hook on the
    decoder's input.  Returns whether the overlap is active (fused optimizer with an early segment, world size > 1)."""
    net = getattr(model, "unet", model)
    net = getattr(net, "module", net)
    ok = hasattr(opt, "launch_early_bucket") and getattr(opt, "early_start", 0) < opt.flat_grad.numel() and dist.is_initialized() and dist.get_world_size() > 1
    net._early_grad_hook = opt.launch_early_bucket if ok else None
    return ok
