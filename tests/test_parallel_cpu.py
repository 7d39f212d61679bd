"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths — batch sharding for sampling, the gradient mean
all-reduce of data-parallel training, the optional global Langevin norms, and gather for logging."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from indm_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        total = 11
        a, b = parallel.shard_range(total)
        x = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
        mine = parallel.shard_batch(x)
        # data-parallel gradient exchange: local "gradient" = mean over the local shard; global mean of shards weighted equally
        g = torch.full((5,), float(rank + 1))
        parallel.allreduce_mean_(g)
        # global Langevin norms from local sums
        local = x[a:b]
        sums = torch.stack([local.norm(dim=1).sum(), (2 * local).norm(dim=1).sum()])
        means = parallel.allreduce_langevin_norms(sums, local.shape[0])
        allx = parallel.gather_cat(mine)
        q.put((rank, a, b, mine.clone(), g.clone(), means.clone(), allx.clone()))
    finally:
        dist.destroy_process_group()


def test_sharding_allreduce_and_gather_world2():
    ws = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = torch.arange(33, dtype=torch.float32).reshape(11, 3)
    (r0, a0, b0, m0, g0, n0, all0), (r1, a1, b1, m1, g1, n1, all1) = res
    assert (a0, b0, a1, b1) == (0, 6, 6, 11)                       # contiguous, remainder to the earlier rank
    assert torch.equal(torch.cat([m0, m1]), x)
    assert torch.allclose(g0, torch.full((5,), 1.5)) and torch.equal(g0, g1)
    want = torch.stack([x.norm(dim=1).mean(), (2 * x).norm(dim=1).mean()])
    assert torch.allclose(n0, want, rtol=1e-6) and torch.allclose(n1, want, rtol=1e-6)
    assert torch.equal(all0, x) and torch.equal(all1, x)


def test_shard_range_covers_everything_without_overlap():
    for total in (1, 7, 128, 1024):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
