#!/usr/bin/env python
"""Development probe: wall-clock / device time of the phases of one training step (vp/CIFAR10/indm_nll, batch 128)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, losses, sde_lib, _lib as L  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.models.ema import ExponentialMovingAverage  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.device = dev
    cfg.training.freeze_flow = True
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    flow = fm.create_flow_model(cfg)
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    B = 128
    batch = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    loss_fn = losses.get_sde_loss_fn(cfg, sde, train=True)
    optimize_fn = losses.optimization_manager(cfg)
    net = model.module

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(4):
        t = {}
        t0 = sync()
        with torch.no_grad():
            latent, lf = fm.flow_forward(cfg, flow, batch, reverse=False, estimator='train')
        t['flow fwd + logdet'] = sync() - t0
        t0 = sync()
        opt.zero_grad()
        eng = net.engine(B)
        if eng._weights_version != eng.weights_version():
            eng.load_weights()
        t['repack weights'] = sync() - t0
        t0 = sync()
        ls = loss_fn(model, latent)
        t['score fwd + loss'] = sync() - t0
        t0 = sync()
        torch.mean(ls).backward()
        t['score bwd (dgrad+wgrad)'] = sync() - t0
        t0 = sync()
        optimize_fn(opt, model.parameters(), step=it)
        ema.update(model.parameters())
        t['clip+adamw+ema'] = sync() - t0
        t0 = sync()
        _ = ls.detach().cpu()
        t['loss readback'] = sync() - t0
        if it >= 1:
            print(' | '.join(f'{k} {v * 1e3:.1f} ms' for k, v in t.items()), '| total', f'{sum(t.values()) * 1e3:.1f} ms',
                  '| flow VJPs', flow.module.engine(B).vjp_count, '| launches so far', L.launches)


if __name__ == '__main__':
    main()
