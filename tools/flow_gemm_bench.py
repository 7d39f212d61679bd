#!/usr/bin/env python
"""Development probe: the GEMM shapes of one iResBlock branch (batch 128, 32x32, idim 512) in isolation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import _lib as L

dev = torch.device('cuda:0')
N, H, W, idim = 128, 32, 32, 512
P = N * H * W
bf = torch.bfloat16
a64 = torch.randn(N, H, W, 64, device=dev).to(bf)
a512 = torch.randn(N, H, W, idim, device=dev).to(bf)
w1 = (torch.randn(idim, 64, device=dev) * 0.05).to(bf)
w2 = (torch.randn(idim, idim, device=dev) * 0.02).to(bf)
w3 = (torch.randn(27, idim, device=dev) * 0.02).to(bf)
o512, aux, mul = (torch.empty(N, H, W, idim, device=dev, dtype=bf) for _ in range(3))
mul.fill_(0.5)
o9 = torch.empty(N, H, W, 28, device=dev)
bias = torch.zeros(idim, device=dev)
rowb = torch.zeros(N, idim, device=dev)
geo = dict(dtype=L.DTYPE_BF16, N=N, H=H, W=W, taps=1)


def timeit(name, flops, **kw):
    for _ in range(3):
        L.igemm(**geo, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.igemm(**geo, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f'{name:44s} {us:8.1f} us  {flops / us / 1e6:8.1f} TFLOP/s', flush=True)


f1, f2, f3 = 2.0 * P * 64 * idim, 2.0 * P * idim * idim, 2.0 * P * idim * 27
timeit('64->512 plain bf16 out', f1, a=a64, Cin=64, b=w1, Cout=idim, out_bf16=o512, out_ld=idim)
timeit('64->512 +bias +Sin +aux_cos', f1, a=a64, Cin=64, b=w1, Cout=idim, bias=bias, act=1, aux_cos=aux, out_bf16=o512, out_ld=idim)
timeit('64->512 *mul', f1, a=a64, Cin=64, b=w1, Cout=idim, mul=mul, mul_ld=idim, out_bf16=o512, out_ld=idim)
timeit('512->512 plain bf16 out', f2, a=a512, Cin=idim, b=w2, Cout=idim, out_bf16=o512, out_ld=idim)
timeit('512->512 +rowbias +Sin +aux_cos', f2, a=a512, Cin=idim, b=w2, Cout=idim, rowbias=rowb, rowbias_ld=idim, act=1, aux_cos=aux, out_bf16=o512, out_ld=idim)
timeit('512->512 *mul', f2, a=a512, Cin=idim, b=w2, Cout=idim, mul=mul, mul_ld=idim, out_bf16=o512, out_ld=idim)
for bn in (0, 32, 64, 128):
    timeit(f'512->27 f32 out ld 28, block_n={bn}', f3, a=a512, Cin=idim, b=w3, Cout=27, out_f32=o9, out_ld=28, block_n=bn)
