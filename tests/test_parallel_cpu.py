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
        sums = torch.stack([local.norm(dim=1).sum(), (2 * local).norm(dim=1).sum(), torch.tensor(float(local.shape[0]))])
        parallel.allreduce_langevin_sums_(sums)
        means = sums[:2] / sums[2]
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


def _bpd_worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from indm_b200 import configs, datasets, evaluation
        cfg = configs.get_config('vp/CIFAR10/indm_nll')
        cfg.flow.model = 'identity'
        cfg.device = torch.device('cpu')
        cfg.data.image_size = 4
        cfg.eval.batch_size = 250
        cfg.eval.num_nelbo = 1
        cfg.eval.skip_nll_wrong = True
        # every image carries its own id in pixel (0, 0, 0): the per-sample "bpd" below is that id, so the global mean is known
        ds = []
        for b in range(4):
            x = torch.zeros(250, 3, 4, 4)
            x[:, 0, 0, 0] = (torch.arange(250) % 256) / 255.
            ds.append(x)
        seen = []

        def sample_id(batch):
            v = (batch[:, 0, 0, 0] + 1.) / 2. * 256.           # scaled + dequantised pixel -> floor recovers the 8-bit value
            return torch.floor(v)

        def nelbo_fn(model, flow, batch, logdet):
            seen.append(batch.shape[0])
            return sample_id(batch), sample_id(batch) * 2

        def nll_fn(model, flow, batch, logdet, residual=True, eps_bpd=1e-5):
            return sample_id(batch), None, 10

        res = evaluation.get_bpd(cfg, ds, datasets.get_data_scaler(cfg), nelbo_fn, nll_fn, None, None, step=0, eval=False)
        q.put((rank, seen, res))
    finally:
        dist.destroy_process_group()


def test_get_bpd_shards_every_batch_and_reduces_the_means_world2():
    """NLL / NELBO evaluation shards by image; the figures a rank reports are means over ALL ranks' samples."""
    ws = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bpd_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(ws)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = float(torch.mean((torch.arange(250) % 256).float()))       # every batch holds ids 0..249
    for rank, seen, res in outs:
        assert seen == [125] * 40                                     # 10000 / 250 batches, half of each per rank
        assert abs(res['nelbo'] - want) < 1e-9 and abs(res['nelbo_residual'] - 2 * want) < 1e-9
        assert abs(res['nll'] - want) < 1e-9 and res['nll_wrong'] is None
    # rank 0 saw ids 0..124, rank 1 ids 125..249: only the reduction makes both report the global mean
    assert outs[0][2]['nll'] == outs[1][2]['nll']
