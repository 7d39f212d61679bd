#!/usr/bin/env python
"""Development probe: device-synchronised wall time of the phases of one JOINT training step (vp/CIFAR10/indm_nll, batch 128)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, losses, sde_lib, _lib as L  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.models.ema import ExponentialMovingAverage  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402
from indm_b200.flow_models.wolf import FlowEngine  # noqa: E402

T = {}


def timed(name, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n0 = L.launches
        r = fn(*a, **k)
        torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        T[name + ' #'] = T.get(name + ' #', 0) + L.launches - n0
        return r
    return wrap


def main():
    dev = torch.device('cuda:0')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.device = dev
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    flow = fm.create_flow_model(cfg)
    sde = sde_lib.get_sde(cfg)
    opt = losses.get_optimizer(cfg, model.parameters())
    state = dict(optimizer=opt, model=model, ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
    fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
    flow_state = dict(optimizer=fopt, model=flow, ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    B = 128
    batch = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    FlowEngine.precapture_backward = lambda self: None      # the phase timers below synchronise inside the backward: no capture here
    FlowEngine.load_weights = timed('flow repack', FlowEngine.load_weights)
    FlowEngine.train_posterior = timed('flow encoder+posterior+KL', FlowEngine.train_posterior)
    FlowEngine.forward_logdet = timed('flow blocks fwd+logdet', FlowEngine.forward_logdet)
    FlowEngine.train_backward = timed('flow backward', FlowEngine.train_backward)
    from indm_b200.flow_models import wolf_backward as wb, wolf_encoder_train as we
    wb.FlowBackward.run = timed('  blocks bwd', wb.FlowBackward.run)
    wb.PosteriorBackward.run = timed('  prior/posterior bwd', wb.PosteriorBackward.run)
    we.EncoderTrain.backward = timed('  encoder bwd', we.EncoderTrain.backward)
    for it in range(14):
        T.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_fn(state, flow_state, batch)
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
        if it >= 1:
            print(f'step {tot * 1e3:.1f} ms | ' + ' | '.join(f'{k} {v * 1e3:.1f} ms ({T[k + " #"]} launches)' for k, v in T.items() if not k.endswith('#')),
                  '| flow VJPs', flow.module.engine(B).vjp_count, flush=True)
    print('max memory GiB', torch.cuda.max_memory_allocated() / 2**30)
    if '--kineto' in sys.argv:
        # per-kernel device time of one step (CUPTI through torch.profiler; a development view, not a bench number)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_fn(state, flow_state, batch)
            torch.cuda.synchronize()
        rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
        tot = sum(r[2] for r in rows)
        print(f'device time of one step: {tot:.1f} ms over {sum(r[1] for r in rows)} kernels')
        for k, c, t in sorted(rows, key=lambda r: -r[2])[:28]:
            print(f'  {t:8.2f} ms {100 * t / tot:5.1f}%  x{c:5d}  {k[:110]}')


if __name__ == '__main__':
    main()
