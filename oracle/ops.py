"""CPU restatement of the reference's two native operators (FP32/FP64, numpy).

TEST INFRASTRUCTURE ONLY — only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this.

Pinned against the live reference (`op/upfirdn2d.py:159-200` `upfirdn2d_native`,
`op/fused_act.py:86-94` CPU branch) by `tests/golden/make_golden.py` →
`tests/golden/ops.npz`, checked in `tests/test_oracle_golden.py`.
"""
import numpy as np


def upfirdn2d(x, k, up=1, down=1, pad=(0, 0)):
    """Up-FIR-down resampler, restating `upfirdn2d_native` (op/upfirdn2d.py:159-200).

    x: [N, C, H, W]; k: [kh, kw]; same up/down/pad on both axes as the public
    wrapper `upfirdn2d(input, kernel, up, down, pad)` (op/upfirdn2d.py:145-156).
      1. zero-insert: xu[y*up, x*up] = x[y, x]            (:168-170)
      2. pad p0 before / p1 after, negative pad crops     (:172-180)
      3. correlate with the flipped kernel                (:186-187)
      4. keep every `down`-th sample                      (:195)
    Output size (H*up + p0 + p1 - kh)//down + 1           (:197-198).
    """
    return upfirdn2d_xy(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_xy(x, k, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    x = np.asarray(x)
    k = np.asarray(k, dtype=x.dtype)
    n, c, h, w = x.shape
    kh, kw = k.shape
    xu = np.zeros((n, c, h * up_y, w * up_x), dtype=x.dtype)
    xu[:, :, ::up_y, ::up_x] = x
    xp = np.pad(xu, ((0, 0), (0, 0), (max(py0, 0), max(py1, 0)), (max(px0, 0), max(px1, 0))))
    xp = xp[:, :, max(-py0, 0): xp.shape[2] - max(-py1, 0), max(-px0, 0): xp.shape[3] - max(-px1, 0)]
    oh_full = h * up_y + py0 + py1 - kh + 1
    ow_full = w * up_x + px0 + px1 - kw + 1
    out = np.zeros((n, c, max(oh_full, 0), max(ow_full, 0)), dtype=x.dtype)
    kf = k[::-1, ::-1]
    for i in range(kh):
        for j in range(kw):
            out += xp[:, :, i:i + oh_full, j:j + ow_full] * kf[i, j]
    return np.ascontiguousarray(out[:, :, ::down_y, ::down_x])


def upfirdn2d_backward(gy, k, up, down, pad, in_hw):
    """Gradient of `upfirdn2d` w.r.t. its input, restating `UpFirDn2d.backward`
    (op/upfirdn2d.py:111-114 for g_pad, :19-44 for the transposed launch):
    the same operator with up<->down, flipped kernel and pads g_pad."""
    k = np.asarray(k)
    kh, kw = k.shape
    in_h, in_w = in_hw
    p0, p1 = pad
    out_h = (in_h * up + p0 + p1 - kh) // down + 1
    out_w = (in_w * up + p0 + p1 - kw) // down + 1
    gx0 = kw - p0 - 1
    gy0 = kh - p0 - 1
    gx1 = in_w * up - out_w * down + p0 - up + 1
    gy1 = in_h * up - out_h * down + p0 - up + 1
    return upfirdn2d_xy(gy, k[::-1, ::-1], down, down, up, up, gx0, gx1, gy0, gy1)


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """`scale * leaky_relu(x + bias[None,:,None,...], slope)`.

    Restates the CUDA semantics (`op/fused_bias_act_kernel.cu:36-45`, act=3 grad=0) which honour
    `negative_slope`; the reference's *CPU* branch hard-codes slope 0.2 (`op/fused_act.py:91`),
    identical at the default.  SURVEY.md §7 hard part 7."""
    x = np.asarray(x)
    b = np.asarray(bias, dtype=x.dtype).reshape((1, -1) + (1,) * (x.ndim - 2))
    y = x + b
    return (np.where(y > 0, y, y * np.asarray(negative_slope, dtype=x.dtype)) * np.asarray(scale, dtype=x.dtype)).astype(x.dtype)


def fused_leaky_relu_backward(gy, out, negative_slope=0.2, scale=2 ** 0.5):
    """grad wrt input (and bias = sum over all but dim 1), restating act=3 grad=1
    (`op/fused_bias_act_kernel.cu:40`, `op/fused_act.py:20-41`): sign taken from the forward output."""
    gy = np.asarray(gy)
    gx = (np.where(np.asarray(out) > 0, gy, gy * np.asarray(negative_slope, dtype=gy.dtype)) * np.asarray(scale, dtype=gy.dtype)).astype(gy.dtype)
    axes = (0,) + tuple(range(2, gy.ndim))
    return gx, gx.sum(axis=axes)
