#!/usr/bin/env python
"""Development probe: indm_gn_apply on the score network's shapes (batch 128) in isolation, as one CUDA graph of launches over
rotating buffers larger than L2.  A/B switches: INDM_GN_STREAM=0 (register kernel), INDM_GN_STAGES, INDM_GN_CHUNK_KB,
INDM_GN_CTAS_PER_SM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import _lib as L

dev = torch.device('cuda:0')


def bench(N, S, C, in_bf16, want_raw, reps=12, nbuf=6):
    tin = torch.bfloat16 if in_bf16 else torch.float32
    xs = [torch.randn(N, S, S, C, device=dev).to(tin) for _ in range(nbuf)]
    outs = [torch.empty(N, S, S, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    raws = [torch.empty(N, S, S, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)] if want_raw else None
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    part = torch.zeros(N, 32, 2, device=dev)
    L.call('indm_gn_stats', L.ptr(xs[0]), C, None, 0, L.DTYPE_BF16 if in_bf16 else L.DTYPE_F32, N, S * S, 32, L.ptr(part))

    def go(i):
        L.call('indm_gn_apply', L.ptr(xs[i]), C, None, 0, L.DTYPE_BF16 if in_bf16 else L.DTYPE_F32, N, S, S, 32, L.ptr(part), L.ptr(gamma),
               L.ptr(beta), 1e-6, 1, 0, L.ptr(outs[i]), L.ptr(raws[i]) if want_raw else None, L.DTYPE_BF16)
    go(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            go(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / 5
    nbytes = N * S * S * C * ((2 if in_bf16 else 4) + 2 + (2 if want_raw else 0))
    print(f'gn_apply N={N} {S:2d}x{S:<2d} C={C:3d} {"bf16" if in_bf16 else "fp32"}->bf16{" +raw" if want_raw else "     "}: {us:7.1f} us '
          f'{nbytes / us / 1e6:6.2f} TB/s', flush=True)


for S, C in ((32, 128), (16, 256), (8, 256)):
    bench(128, S, C, False, False)
    bench(128, S, C, False, True)
    bench(128, S, C, True, False)
