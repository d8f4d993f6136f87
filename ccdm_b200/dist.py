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
