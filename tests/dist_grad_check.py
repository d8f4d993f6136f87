"""torchrun worker (world size 2, NCCL) for tests/test_gpu_dist.py: data-parallel gradient parity on hardware.

SURVEY.md section 8e semantics: every rank runs the training forward / backward on ITS half of the batch with local
BatchNorm1d statistics, then the flat gradient buffer is averaged with one NCCL all-reduce (FusedAdam.all_reduce_gradients,
what GraphedTrainStep captures).  Checks, on rank 0:
  1. the all-reduced buffer == mean of the two ranks' local gradients (gathered bit-exactly) to 1e-6 (the collective itself);
  2. it equals the mean of the two half-batch gradients recomputed by ONE process (rank 0 runs both halves itself) up to the
     run-to-run noise of the fp32 atomics in the weight-gradient kernels (cosine >= 0.9999, relative error <= 2e-3);
  3. after one fused Adam step the two replicas hold bit-identical parameters.
Prints "DIST_OK ..." on success.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import ccdm_b200
from ccdm_b200 import dist as D
from ccdm_b200.optim import FusedAdam
from ccdm_b200.train import unet_train_forward


def main():
    rank, local_rank, world = D.init("nccl")
    assert world == 2, world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(1234)                                     # identical initial weights on both ranks
    net = ccdm_b200.Unet(dim=64, dim_mults=(1, 2, 2), cond_drop_prob=0.1).to(dev).train()
    D.broadcast_parameters(net)
    opt = FusedAdam(net.parameters(), lr=1e-3, betas=(0.9, 0.99), max_grad_norm=1.0,
                    early_params=D.early_gradient_params(net))
    assert 0 < opt.early_start < opt.flat_grad.numel() and D.wire_overlap(net, opt)

    g = torch.Generator().manual_seed(7)                        # the GLOBAL batch, identical on both ranks
    B = 8
    x = torch.randn(2 * B, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (2 * B,), generator=g)
    emb = torch.rand(2 * B, 128, generator=g)
    keep = torch.rand(2 * B, generator=g) > 0.2
    dout = torch.randn(2 * B, 3, 32, 32, generator=g)

    def half_grad(h, overlap=False):
        sl = slice(h * B, (h + 1) * B)
        opt.zero_grad()                                          # (BatchNorm1d uses batch statistics in training mode)
        opt.arm_early_bucket(overlap)                            # overlap: the decoder segment is all-reduced DURING backward
        unet_train_forward(net, x[sl].to(dev), t[sl].to(dev), emb[sl].to(dev), keep[sl].to(dev)).backward(dout[sl].to(dev))
        return opt.flat_grad.clone()

    local = half_grad(rank)                                      # this rank's shard
    opt.flat_grad.copy_(local)
    opt.all_reduce_gradients()
    reduced = opt.flat_grad.clone()
    both = [torch.empty_like(local) for _ in range(2)]
    dist.all_gather(both, local)
    # 1b. the overlapped exchange (early segment launched from the gradient hook on a side stream, late segment afterwards)
    #     gives the same mean up to the run-to-run noise of the fp32 atomics of a second backward pass
    half_grad(rank, overlap=True)
    assert opt._early_launched, "the early bucket was not launched from the backward hook"
    opt.all_reduce_gradients()
    torch.cuda.synchronize()
    e_ov = ((opt.flat_grad - reduced).norm() / reduced.norm()).item()
    ok = True
    msg = ""
    if rank == 0:
        mean = (both[0] + both[1]) / 2
        e1 = ((reduced - mean).norm() / mean.norm()).item()
        ref = (half_grad(0) + half_grad(1)) / 2                  # one process, both halves
        e2 = ((reduced - ref).norm() / ref.norm()).item()
        cos = (reduced.double() @ ref.double() / (reduced.double().norm() * ref.double().norm())).item()
        ok = e1 <= 1e-6 and e2 <= 2e-3 and cos >= 0.9999 and e_ov <= 2e-3
        msg = f"allreduce_vs_gathered_mean={e1:.2e} dp_vs_single_process={e2:.2e} cos={cos:.6f} overlapped_vs_plain={e_ov:.2e}"
    opt.flat_grad.copy_(reduced)
    opt.step()
    flat_p = torch.cat([p.detach().flatten() for p in net.parameters()])
    ps = [torch.empty_like(flat_p) for _ in range(2)]
    dist.all_gather(ps, flat_p)
    same = torch.equal(ps[0], ps[1])
    if rank == 0:
        print(("DIST_OK " if ok and same else "DIST_FAIL ") + msg + f" replicas_identical={same}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
