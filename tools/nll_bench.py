#!/usr/bin/env python
"""BASELINE configs[2], second half: probability-flow ODE NLL with the Hutchinson trace (`likelihood.get_likelihood_fn`,
vp/CIFAR10/indm_nll, wolf flow forward + log-det inside every call), batch 128 on one GPU, with both integrators:
  method='RK45'         SciPy on the host like the reference (likelihood.py:116): full float64 state host <-> device per RHS
  method='RK45-device'  indm_b200/ode.py: same Dormand-Prince controller, state resident on the GPU
Same data / Hutchinson probe / noise for both, so bpd and nfe must agree.  One JSON line per method + one comparison line.
Weights: the modules' own initialisers under manual_seed(0) — the head conv is ~0-initialised, so the ODE is
smooth and NFE is small; the per-NFE cost (1 forward + 1 input-VJP) is what the line measures.
    python tools/nll_bench.py [--batch 128] [--reps 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, sde_lib, likelihood, _lib as L  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--flow', default='wolf', choices=['wolf', 'identity'])
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.device = dev
    if a.flow == 'identity':
        cfg.flow.model = 'identity'
    model = mutils.create_model(cfg)
    model.eval()
    flow = fm.create_flow_model(cfg) if a.flow == 'wolf' else None
    if flow is not None:
        flow.eval()
    sde = sde_lib.get_sde(cfg)
    B = a.batch
    g = torch.Generator(device='cpu').manual_seed(1)
    data = (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).to(dev)
    eps = (torch.randint(0, 2, data.shape, generator=g).float() * 2 - 1).to(dev)
    noise = torch.randn(data.shape, generator=g).to(dev)
    rn = (torch.randn(data.shape, generator=g).to(dev), torch.randn(data.shape, generator=g).to(dev))
    res = {}
    for method in ('RK45-device', 'RK45'):
        fn = likelihood.get_likelihood_fn(cfg, sde, lambda v: (v + 1.) / 2., method=method)
        out = None
        times = []
        for r in range(a.reps + 1):
            # pin the flow's own draws so both integrators see the same latent: posterior noise / Hutchinson probes come from the
            # in-kernel Philox stream at an offset counted per call, the Poisson series lengths from numpy's global generator
            torch.manual_seed(7)
            np.random.seed(7)
            if flow is not None:
                for m_ in flow.modules():
                    if hasattr(m_, '_draws'):
                        m_._draws = 1000
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            l0 = L.launches
            out = fn(model, flow, data, epsilon=eps, noise=noise, residual_noise=rn)
            bpd_host = out[0].cpu()                   # the result the caller reads
            torch.cuda.synchronize()
            if r > 0:
                times.append(time.perf_counter() - t0)
            launches = L.launches - l0
        ms = 1e3 * sum(times) / len(times)
        res[method] = (bpd_host, out[1].cpu(), out[2])
        print(json.dumps({"leg": f"vp/CIFAR10/indm_nll likelihood_fn(method='{method}'), flow={a.flow}", "batch": B, "nfe": int(out[2]),
                          "ms_per_call": ms, "images_per_sec": B / (ms * 1e-3), "ms_per_nfe": ms / max(int(out[2]), 1),
                          "gpu_launches_per_call": launches, "bpd_mean": float(bpd_host.mean()),
                          "finite": bool(torch.isfinite(bpd_host).all())}), flush=True)
    (b0, z0, n0), (b1, z1, n1) = res['RK45'], res['RK45-device']
    print(json.dumps({"compare": "RK45-device vs RK45 (SciPy)", "nfe": [int(n0), int(n1)], "max_abs_bpd_diff": float((b0 - b1).abs().max()),
                      "latent_rel_l2": float((z0 - z1).norm() / z0.norm())}), flush=True)


if __name__ == '__main__':
    main()
