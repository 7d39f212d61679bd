#!/usr/bin/env python
"""Development probe: the FIR resampling launches of one VE score-network forward (batch 128) in isolation, GB/s of in + out bytes."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import _lib as L
dev = torch.device('cuda:0')
k1 = (np.array([1, 3, 3, 1], np.float32) / 8)
kp = k1.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
cases = [(32, 128, 2, 'f32'), (32, 128, 2, 'bf16'), (16, 256, 2, 'f32'), (16, 256, 1, 'bf16'), (16, 256, 1, 'f32'), (8, 256, 1, 'bf16'), (32, 64, 3, 'bf16'), (16, 128, 3, 'f32')]
for S, C, mode, dt in cases:
    N = 128
    tin = torch.float32 if dt == 'f32' else torch.bfloat16
    So = {1: 2 * S, 2: S // 2, 3: S + 1}[mode]
    xs = [torch.randn(N, S, S, C, device=dev).to(tin) for _ in range(6)]
    ys = [torch.empty(N, So, So, C, device=dev, dtype=torch.bfloat16) for _ in range(6)]
    din = L.DTYPE_F32 if dt == 'f32' else L.DTYPE_BF16
    def run(i):
        L.call('indm_fir_nhwc', L.ptr(xs[i % 6]), L.ptr(ys[i % 6]), din, L.DTYPE_BF16, N, S, S, C, kp, mode)
    for i in range(6): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(24): run(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 24
    by = xs[0].numel() * xs[0].element_size() + ys[0].numel() * 2
    print(f'mode {mode} {S}x{S}x{C} {dt}: {us:6.1f} us {by / us / 1e3:7.0f} GB/s', flush=True)
