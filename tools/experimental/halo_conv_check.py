#!/usr/bin/env python
"""Round-2 starting point (never run on hardware yet): build and check the EXPERIMENTAL padded-pixel 3x3 convolution
(indm_b200/csrc/experimental/igemm_halo.cu) against torch's bf16 convolution and against the production igemm kernel's timing.
    gpurun -- 'make -C indm_b200/csrc experimental && python tools/experimental/halo_conv_check.py'
The first thing this answers is whether tcgen05 accepts a 128B-swizzled operand whose start address is offset by whole 128-byte
rows inside the 1024-byte swizzle atom (descriptor base-offset field); if the small cases fail with a structured error (every
output row wrong except taps whose row offset is a multiple of 8), that assumption is what to revisit."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "indm_b200", "libindm_experimental.so"))
    fn = lib.indm_exp_conv3x3_halo_bf16
    fn.restype = ctypes.c_int
    lib.indm_last_error.restype = ctypes.c_char_p
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for (N, H, W, Cin, Cout) in [(1, 8, 8, 64, 128), (2, 16, 16, 64, 128), (2, 32, 32, 128, 128), (128, 32, 32, 128, 128),
                                 (128, 32, 32, 256, 128), (128, 16, 16, 256, 256)]:
        x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)).to(torch.bfloat16)
        b = torch.randn(Cout, device=dev)
        wpack = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
        out = torch.full((N, H, W, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = fn(vp(x), vp(wpack), vp(b), vp(out), N, H, W, Cin, Cout, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            print("error:", lib.indm_last_error().decode())
            return 1
        torch.cuda.synchronize()
        want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, padding=1).permute(0, 2, 3, 1)
        err = float((out.float() - want).norm() / want.norm())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            fn(vp(x), vp(wpack), vp(b), vp(out), N, H, W, Cin, Cout, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        ev[0].record()
        for _ in range(10):
            fn(vp(x), vp(wpack), vp(b), vp(out), N, H, W, Cin, Cout, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) * 100
        tf = 2.0 * N * H * W * Cout * 9 * Cin / (us * 1e-6) / 1e12
        print(f"N={N} {H}x{W} {Cin}->{Cout}: rel-L2 {err:.2e} (bf16 output: expect ~3e-3), {us:.1f} us, {tf:.0f} TFLOP/s "
              f"(production igemm, CTA pairs: 62 us / 623 TFLOP/s at 128x32x32 128->128)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
