#!/usr/bin/env python
"""Development probe: one wolf-flow forward with the training log-det series (batch 128, CIFAR flow), for ncu launch lists."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402

dev = torch.device('cuda:0')
cfg = configs.get_config('vp/CIFAR10/indm_nll')
cfg.device = dev
torch.manual_seed(0)
flow = fm.create_flow_model(cfg)
flow.eval()
x = torch.rand(128, 3, 32, 32, device=dev) * 2 - 1
ns = np.zeros(32, dtype=np.int64)          # K = 2 terms + the Neumann VJP: 3 VJPs per block
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    z, l = fm.flow_forward(cfg, flow, x, reverse=False, estimator='train', n_terms=ns)
    torch.cuda.synchronize()
    print(f'flow forward (96 VJPs): {(time.perf_counter() - t0) * 1e3:.1f} ms')
