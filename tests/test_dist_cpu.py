"""world_size-2 gloo test of the N>1 host logic (sharding, max-over-ranks timing, image gather) on CPU."""
import os
import socket

import torch
import torch.multiprocessing as mp

from ccdm_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = D.init("gloo")
    labels = torch.linspace(0, 1, 7)
    mine = D.shard(labels, r, w)
    D.barrier()
    slowest = D.max_over_ranks(1.0 + r)                     # rank 1 is "slower"
    img = (mine * 255).to(torch.uint8).view(-1, 1, 1, 1).repeat(1, 3, 2, 2)
    allimg = D.gather_images_u8(img)
    # data-parallel training exchange: replicas start equal, gradients are averaged in flat buckets
    torch.manual_seed(r)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    D.broadcast_parameters(lin)
    w0 = lin[0].weight.detach().clone()
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(r + 1) * (i + 1))
    D.all_reduce_gradients(list(lin.parameters()), bucket_bytes=64)      # tiny buckets: several collectives
    gsum = [p.grad.flatten()[0].item() for p in lin.parameters()]
    q.put((r, mine.tolist(), slowest, None if allimg is None else allimg[:, 0, 0, 0].tolist(), w0.sum().item(), gsum))
    torch.distributed.destroy_process_group()


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 200, 256):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    labels = torch.linspace(0, 1, 7)
    assert res[0][1] + res[1][1] == labels.tolist()         # shards partition the batch in order
    assert res[0][2] == res[1][2] == 2.0                    # max over ranks
    assert res[0][3] == (labels * 255).to(torch.uint8).tolist() and res[1][3] is None
    assert res[0][4] == res[1][4]                           # parameters broadcast from rank 0
    assert res[0][5] == res[1][5] == [1.5 * (i + 1) for i in range(4)]   # mean of (1, 2) * (i + 1)
