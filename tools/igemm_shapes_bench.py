#!/usr/bin/env python
"""Development probe: the score network's dominant convolution shapes (batch 128) through indm_igemm in isolation, with the
epilogues the engine uses (Conv_0: bf16 out + bias + per-image row bias + GroupNorm statistics; Conv_1: fp32 out + bias + residual).
Environment switches of the kernel can be A/B-ed from outside: INDM_IGEMM_TSTORE=0, INDM_IGEMM_CTA2=0, INDM_IGEMM_DBG=1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indm_b200 import _lib as L

dev = torch.device('cuda:0')
bf = torch.bfloat16


PP = bool(os.environ.get('PP'))      # operands in the padded-pixel layout (indm_igemm_t.a_pp)


def bench(N, S, Cin, Cout, kind, taps=9, reps=20, nbuf=6):
    if PP and S <= 16:
        xs = [torch.randn((N * (S + 1) + 1) * (S + 2), Cin, device=dev).to(bf) for _ in range(nbuf)]
    else:
        xs = [torch.randn(N, S, S, Cin, device=dev).to(bf) for _ in range(nbuf)]
    w = (torch.randn(taps, Cout, Cin, device=dev) / (taps * Cin) ** 0.5).to(bf)
    bias = torch.randn(Cout, device=dev)
    rowb = torch.randn(N, Cout, device=dev)
    part = torch.zeros(N, 32, 2, device=dev)
    kws = []
    for i in range(nbuf):
        kw = dict(dtype=L.DTYPE_BF16, a=xs[i], N=N, H=S, W=S, Cin=Cin, b=w, Cout=Cout, taps=taps, bias=bias, out_ld=Cout)
        if PP and S <= 16:
            kw['a_pp'] = 1
        if S * S < 32:
            part = None
        if kind == 'conv0':
            kw.update(out_bf16=torch.empty(N, S, S, Cout, device=dev, dtype=bf), rowbias=rowb, rowbias_ld=Cout)
            if part is not None:
                kw.update(gn_partial=part, gn_cpg=Cout // 32, gn_groups=32)
        else:
            kw.update(out_f32=torch.empty(N, S, S, Cout, device=dev), residual=torch.randn(N, S, S, Cout, device=dev), res_ld=Cout,
                      res_scale=0.7071, scale=0.7071)
            if part is not None:
                kw.update(gn_partial=part, gn_cpg=Cout // 32, gn_groups=32)
        kws.append(kw)
    for i in range(nbuf):
        L.igemm(**kws[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # one CUDA graph of `reps` launches: short kernels are otherwise timed at the host's launch rate
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            L.igemm(**kws[i % nbuf])
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps / 5
    fl = 2.0 * N * S * S * Cout * taps * Cin
    print(f'{kind:6s} N={N} {S:2d}x{S:<2d} {Cin:3d}->{Cout:3d} taps {taps}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s', flush=True)


SIZES = [int(v) for v in os.environ.get('SIZES', '32,16,8,4').split(',')]
for kind in ('conv0', 'conv1'):
    for S in SIZES:
        if S == 32:
            bench(128, 32, 128, 128, kind)
            bench(128, 32, 256, 128, kind)
        else:
            bench(128, S, 256, 256, kind)
            bench(128, S, 512, 256, kind)
