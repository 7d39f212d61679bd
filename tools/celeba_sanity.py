#!/usr/bin/env python
"""Development probe: BASELINE configs 4 / 5 (CelebA 64x64, num_res_blocks = 8) run end to end: VE PC sampling with the Langevin
corrector + wolf flow inverse (flow.squeeze), and a VP training step (frozen flow)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, sde_lib, sampling, losses  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.models.ema import ExponentialMovingAverage  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)

cfg = configs.get_config('ve/CELEBA/indm')
cfg.model.num_res_blocks = 8
cfg.device = dev
cfg.sampling.num_scales = 20
model = mutils.create_model(cfg)
flow = fm.create_flow_model(cfg)
flow.eval()
sde = sde_lib.get_sde(cfg)
print('VE CelebA params', sum(p.numel() for p in model.parameters()), 'flow', sum(p.numel() for p in flow.parameters()))
fn = sampling.get_sampling_fn(cfg, sde, (B, 3, 64, 64), lambda v: v, cfg.sampling.truncation_time)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    before, after, nfe = fn(model, flow, seed=it)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'VE CelebA PC+Langevin 20 steps (40 NFE) batch {B}: {dt * 1e3:.0f} ms -> {B / (dt * 50):.2f} img/s at 1000 steps; finite {bool(torch.isfinite(after).all())}')
del model, flow
torch.cuda.empty_cache()

cfg = configs.get_config('vp/CELEBA/indm_nll')
cfg.model.num_res_blocks = 8
cfg.device = dev
cfg.training.freeze_flow = True
model = mutils.create_model(cfg)
flow = fm.create_flow_model(cfg)
flow.eval()
sde = sde_lib.get_sde(cfg)
opt = losses.get_optimizer(cfg, model.parameters())
state = dict(optimizer=opt, model=model, ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
batch = torch.rand(B, 3, 64, 64, device=dev) * 2 - 1
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = step_fn(state, dict(model=flow, step=0), batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'VP CelebA nres=8 train step batch {B}: {dt * 1e3:.0f} ms -> {B / dt:.1f} samples/s; loss {float(res[0].mean()):.3f} finite {bool(torch.isfinite(res[0]).all())}')
