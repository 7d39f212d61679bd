#!/usr/bin/env python
"""Development repro: flow reverse (sampling leg, TF32 engine) followed by joint training steps with the TF32 blocks policy."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, losses, sde_lib, precision
from indm_b200.models import utils as mutils
from indm_b200.models.ema import ExponentialMovingAverage
from indm_b200.flow_models import flow_model as fm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda:0')
cfg = configs.get_config('vp/CIFAR10/indm_nll')
if len(sys.argv) > 2:
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 128, (1, 2), 1, (16,)
cfg.device = dev
torch.manual_seed(0)
model = mutils.create_model(cfg)
flow = fm.create_flow_model(cfg)
flow.eval()
sde = sde_lib.get_sde(cfg)
z = torch.randn(B, 3, 32, 32, device=dev)
for _ in range(3):
    fm.flow_forward(cfg, flow, z, log_det=None, reverse=True)
torch.cuda.synchronize()
print('reverse ok', flush=True)
if os.environ.get('EMPTY_CACHE'):
    torch.cuda.empty_cache(); torch.cuda.synchronize(); print('cache emptied', flush=True)
opt = losses.get_optimizer(cfg, model.parameters())
state = dict(optimizer=opt, model=model, ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
flow_state = dict(optimizer=fopt, model=flow, ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
batch = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
for i in range(0 if os.environ.get('SKIP_BF16') else 4):
    step_fn(state, flow_state, batch)
    torch.cuda.synchronize()
    print('bf16 step', i, flush=True)
precision.set_policy('flow', 'training', 'tf32')
for i in range(int(os.environ.get('TF32_STEPS', '5'))):
    step_fn(state, flow_state, batch)
    torch.cuda.synchronize()
    if os.environ.get('EMPTY_CACHE'):
        torch.cuda.empty_cache()
    print('tf32 step', i, flush=True)
print('done')
