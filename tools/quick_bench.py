#!/usr/bin/env python
"""Development probe: time one score-network forward (VP CIFAR-10 DDPM++, batch 128) eagerly, per launch, and as a
CUDA graph.  Not the benchmark of record (that is bench.py)."""
import argparse
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import configs  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--config', default='vp/CIFAR10/indm_fid')
    ap.add_argument('--mode', default='bf16')
    ap.add_argument('--per-op', action='store_true')
    ap.add_argument('--nres', type=int, default=0, help='override model.num_res_blocks (CelebA bench legs use 8)')
    ap.add_argument('--infer', action='store_true', help='the forward-only plan of the samplers (padded-pixel operands on small maps)')
    a = ap.parse_args()
    cfg = configs.get_config(a.config)
    cfg.device = torch.device('cuda:0')
    if a.nres:
        cfg.model.num_res_blocks = a.nres
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    net = model.module
    net.compute_mode = a.mode
    net.eval()
    t0 = time.time()
    eng = net.engine(a.batch, infer=a.infer)
    torch.cuda.synchronize()
    print(f'engine build {time.time() - t0:.2f}s, {eng.num_launches} launches/forward, '
          f'{torch.cuda.memory_allocated() / 2**30:.2f} GiB allocated')
    eng.x_in.normal_()
    eng.time_cond.fill_(500.0)
    for _ in range(3):
        eng.launch()
    torch.cuda.synchronize()
    if a.infer:
        base = net.engine(a.batch)
        base.x_in.copy_(eng.x_in); base.time_cond.copy_(eng.time_cond)
        base.launch()
        torch.cuda.synchronize()
        d = (eng.out.float() - base.out.float()).norm() / base.out.float().norm()
        print(f'infer plan vs default plan: output rel-L2 {float(d):.3e}')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        eng.launch()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    gflop = 21.69 * a.batch
    print(f'eager: {ms:.3f} ms/forward  -> {gflop / ms:.1f} TFLOP/s algorithmic')
    if a.per_op:
        names = []
        for op in eng.ops:
            n = getattr(op, '__name__', 'op')
            cl = op.__closure__
            tag = n
            if cl:
                for c in cl:
                    v = c.cell_contents
                    if isinstance(v, str):
                        tag = v
            names.append(tag)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.ops) + 1)]
        tot = collections.defaultdict(float)
        cnt = collections.Counter()
        evs[0].record()
        for i, op in enumerate(eng.ops):
            op()
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            tot[n] += evs[i].elapsed_time(evs[i + 1])
            cnt[n] += 1
        for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f'  {n:28s} x{cnt[n]:4d}  {v:8.3f} ms')
        # per-shape igemm table
        shp = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for i, op in enumerate(eng.ops):
            d = None
            for c in (op.__closure__ or ()):
                if c.cell_contents.__class__.__name__ == 'IgemmDesc':
                    d = c.cell_contents
            if d is None:
                continue
            key = (d.H, d.W, d.Cin, d.Cin2 if d.a2 else 0, d.Cout, d.taps, d.batched_b)
            fl = 2.0 * d.N * d.H * d.W * d.Cout * (d.Cin * d.taps + (d.Cin2 if d.a2 else 0))
            r = shp[key]
            r[0] += 1; r[1] += evs[i].elapsed_time(evs[i + 1]); r[2] += fl
        print('  igemm shapes (H,W,Cin,Cin2,Cout,taps,batched): count  ms  TFLOP/s')
        for k, r in sorted(shp.items(), key=lambda kv: -kv[1][1]):
            print(f'   {str(k):40s} x{r[0]:3d} {r[1]:8.3f} ms  {r[2] / r[1] / 1e9:8.1f}')
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.launch()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        eng.launch()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        g.replay()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print(f'graph: {ms:.3f} ms/forward  -> {gflop / ms:.1f} TFLOP/s algorithmic ({a.batch / ms * 1e3 / 1000:.2f} img/s at 1000 NFE)')


if __name__ == '__main__':
    main()
