#!/usr/bin/env python
"""Cost of the compensated (3xTF32) mode next to BF16, leg by leg, at the benched sizes (vp/CIFAR10/indm_nll, batch 128):
flow reverse, flow eval forward + log-det, flow training forward + backward, score forward, score input-VJP — and the
cross-mode differences of the results (the TF32 mode is the one pinned against the reference at 1e-3 / 1e-4 / 0.01 bpd).
One JSON line per leg.     python tools/precision_probe.py [--batch 128] [--legs reverse,eval,train,score]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs, _lib as L  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402


def timeit(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--legs', default='reverse,eval,train,score')
    ap.add_argument('--modes', default='bf16,tf32')
    a = ap.parse_args()
    legs, modes = a.legs.split(','), a.modes.split(',')
    dev = torch.device('cuda:0')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.device = dev
    B = a.batch
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).to(dev)
    z = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    eps64 = torch.randn(B, 64, generator=g).to(dev)
    res = {}
    if any(l in legs for l in ('reverse', 'eval', 'train')):
        flow = fm.create_flow_model(cfg)
        core = flow.module
        nblk = len(core.blocks())
        ns = np.random.RandomState(3).poisson(2.0, size=nblk)
        varepss = []
        for (s, b, m) in core.blocks():
            f = 2 ** s
            varepss.append(torch.randn(B, m.channels, 32 // f, 32 // f, generator=g).to(dev))
        for mode in modes:
            core.compute_mode = mode
            if 'reverse' in legs:
                flow.eval()
                ms, out = timeit(lambda: core(z, reverse=True, eps=eps64))
                res[('reverse', mode)] = out.float().cpu()
                print(json.dumps({"leg": "flow reverse (32 fixed-point inverses)", "mode": mode, "batch": B, "ms": ms,
                                  "iterations": int(sum(core.engine(B).iterations))}), flush=True)
            if 'eval' in legs:
                flow.eval()
                ms, out = timeit(lambda: core(x, reverse=False, eps=eps64, vareps=varepss, n_terms=ns), reps=2, warm=2)
                res[('eval', mode)] = (out[0].float().cpu(), out[1].float().cpu())
                print(json.dumps({"leg": "flow eval forward + (20+n)-term log-det + KL", "mode": mode, "batch": B, "ms": ms,
                                  "vjps": int(core.engine(B).vjp_count)}), flush=True)
            if 'train' in legs:
                flow.train()
                Gz = torch.randn(B, 3, 32, 32, generator=g).to(dev)

                def step():
                    for p in core.parameters():
                        if p.grad is not None:
                            p.grad.zero_()
                    zz, ld = core(x, reverse=False, eps=eps64, vareps=varepss, n_terms=ns)
                    ((zz * Gz).sum() + ld.sum()).backward()
                    return zz.detach(), ld.detach()
                ms, out = timeit(step, reps=3, warm=3)
                gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in core.parameters() if p.grad is not None)))
                res[('train', mode)] = (out[0].float().cpu(), out[1].float().cpu())
                print(json.dumps({"leg": "flow training forward (Neumann series) + full backward", "mode": mode, "batch": B, "ms": ms,
                                  "grad_norm": gn}), flush=True)
        for leg in ('reverse', 'eval', 'train'):
            if (leg, 'bf16') in res and (leg, 'tf32') in res:
                a_, b_ = res[(leg, 'bf16')], res[(leg, 'tf32')]
                if leg == 'reverse':
                    print(json.dumps({"compare": leg, "max_abs_diff_x": float((a_ - b_).abs().max())}), flush=True)
                else:
                    print(json.dumps({"compare": leg, "max_abs_diff_z": float((a_[0] - b_[0]).abs().max()),
                                      "logdet_rel_diff_max": float(((a_[1] - b_[1]).abs() / b_[1].abs().clamp_min(1e-6)).max()),
                                      "logdet_rel_diff_vs_maxabs": float((a_[1] - b_[1]).abs().max() / b_[1].abs().max())}), flush=True)
        del flow, core
    if 'score' in legs:
        model = mutils.create_model(cfg)
        net = model.module
        with torch.no_grad():
            for n_, p_ in net.named_parameters():
                if p_.dim() > 1 and float(p_.abs().max()) < 1e-6:
                    fan = p_[0].numel() + p_.shape[0] * (p_[0, 0].numel() if p_.dim() > 2 else 1)
                    p_.uniform_(-1, 1).mul_((6.0 / fan) ** 0.5)
        net.eval()
        t = torch.rand(B, generator=g).to(dev) * 999
        v = (torch.randint(0, 2, (B, 3, 32, 32), generator=g).float() * 2 - 1).to(dev)
        for mode in modes:
            eng = net.engine(B, mode)
            ms_f, out = timeit(lambda: eng.forward(z, t), reps=5)
            ms_b, gx = timeit(lambda: eng.vjp(v), reps=5)
            res[('score', mode)] = (out.clone().cpu(), gx.clone().cpu())
            print(json.dumps({"leg": "score forward / input-VJP (one PF-ODE right-hand side = both)", "mode": mode, "batch": B, "ms_forward": ms_f,
                              "ms_vjp": ms_b}), flush=True)
        if ('score', 'bf16') in res and ('score', 'tf32') in res:
            a_, b_ = res[('score', 'bf16')], res[('score', 'tf32')]
            rl = lambda p, q: float((p - q).norm() / q.norm())
            print(json.dumps({"compare": "score", "rel_l2_out": rl(a_[0], b_[0]), "rel_l2_vjp": rl(a_[1], b_[1])}), flush=True)


if __name__ == '__main__':
    main()
