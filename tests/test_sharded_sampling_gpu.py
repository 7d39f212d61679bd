"""GPU, world_size 2: batch-sharded VE PC sampling with the Langevin corrector.  The reference takes the corrector's gradient /
noise norms as means over the WHOLE batch (sampling.py:286-288); with `sampling.global_langevin_norms = True` every rank
all-reduces three floats per corrector step and the sharded trajectory equals the single-batch one of the live reference
(tests/golden/pc_tiny_ve.npz, batch 3 split 2 + 1).  Both ranks share cuda:0 over gloo here (the test box has one GPU; NCCL
refuses two ranks on one device) — on a multi-GPU node the same call runs over NCCL inside the sampler's CUDA graph (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q, global_norms):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from helpers import load_npz, tiny
        from indm_b200 import configs, sde_lib, sampling, parallel
        from indm_b200.models import utils as mutils
        from oracle import ncsnpp as oncsnpp
        g = load_npz('pc_tiny_ve.npz')
        cfg = configs.get_config('ve/CIFAR10/indm')
        tiny(cfg)
        cfg.sampling.method, cfg.sampling.predictor, cfg.sampling.corrector = 'pc', 'reverse_diffusion', 'langevin'
        cfg.sampling.num_scales = int(g['num_scales'])
        cfg.sampling.global_langevin_norms = global_norms
        cfg.flow.model = 'identity'
        cfg.device = torch.device('cuda:0')
        model = mutils.create_model(cfg)
        model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, 11).items()})
        model.eval()
        model.module.compute_mode = 'tf32'
        sde = sde_lib.get_sde(cfg)
        a, b = parallel.shard_range(g['prior'].shape[0])
        S = cfg.data.image_size
        fn = sampling.get_sampling_fn(cfg, sde, (b - a, 3, S, S), lambda v: v, float(g['eps']))
        prior = torch.from_numpy(g['prior'][a:b]) * cfg.model.sigma_max
        before, _, _ = fn(model, None, prior=prior, noise=[torch.from_numpy(n[a:b]) for n in g['noises']])
        torch.cuda.synchronize()
        q.put((rank, a, b, before.cpu().numpy()))
    finally:
        dist.destroy_process_group()


def _run(global_norms):
    ws = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q, global_norms)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return np.concatenate([r[3] for r in res], axis=0)


def test_sharded_langevin_trajectory_equals_single_batch_reference():
    from helpers import load_npz, rel_l2
    want = load_npz('pc_tiny_ve.npz')['out']
    got = _run(True)
    e_on = rel_l2(got, want)
    e_off = rel_l2(_run(False), want)
    print(f'sharded 2 + 1 VE PC + Langevin vs the single-batch reference: global norms rel-L2 {e_on:.2e}; per-rank statistics rel-L2 {e_off:.2e}')
    assert e_on < 1e-3                    # the TF32-mode tolerance of the unsharded trajectory test
    assert e_off > 10 * e_on              # without the exchange the shards follow different (statistically equivalent) trajectories
