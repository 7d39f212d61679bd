"""`upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))` — the reference's op/upfirdn2d.py:145-156 on the C-ABI kernel.

Autograd structure follows the reference (op/upfirdn2d.py:19-142): the gradient w.r.t. the input is the same operator with
up<->down, the flipped kernel and the `g_pad` paddings; its own gradient is the forward operator again, so double
backward works.  The FIR kernel receives no gradient (op/upfirdn2d.py:142).  CUDA tensors only: there is no CPU path
(the reference's `upfirdn2d_native` lives in `oracle/ops.py` as the checker).
"""
import torch
from torch.autograd import Function

from .. import _lib as L


def _launch(x, kernel, up, down, pad):
    """x [major, H, W] fp32 contiguous -> [major, H', W'] through indm_upfirdn2d_f32 (op/upfirdn2d_kernel.cu:209-369)."""
    if not x.is_cuda:
        raise RuntimeError('indm_b200.op.upfirdn2d needs CUDA tensors: there is no CPU / PyTorch fallback path')
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    major, in_h, in_w = x.shape
    kh, kw = kernel.shape
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1
    if out_h <= 0 or out_w <= 0:
        raise RuntimeError(f'upfirdn2d: empty output {out_h}x{out_w}')
    x = x.contiguous().float()
    k = kernel.to(device=x.device, dtype=torch.float32).contiguous()
    y = torch.empty((major, out_h, out_w), device=x.device, dtype=torch.float32)
    L.call('indm_upfirdn2d_f32', L.ptr(x), L.ptr(k), L.ptr(y), major, in_h, in_w, kh, kw, up_x, up_y, down_x, down_y,
           px0, px1, py0, py1)
    return y


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        g = grad_output.reshape(-1, out_size[0], out_size[1])
        gi = _launch(g, grad_kernel, down, up, g_pad)
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad, in_size, out_size)
        return gi.view(in_size[0], in_size[1], in_size[2], in_size[3])

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        gg = _launch(gradgrad_input.reshape(-1, in_size[2], in_size[3]), kernel, up, down, pad)
        return gg.view(in_size[0], in_size[1], out_size[0], out_size[1]), None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        batch, channel, in_h, in_w = input.shape
        out = _launch(input.reshape(-1, in_h, in_w), kernel, up, down, pad)
        out_h, out_w = out.shape[1], out.shape[2]
        ctx.in_size = tuple(input.shape)
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        # paddings of the transposed operator (op/upfirdn2d.py:111-114)
        ctx.g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                     kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        gi = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad, ctx.in_size, ctx.out_size)
        return gi, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d.py:145-156.  input [N,C,H,W], kernel [kh,kw] -> [N,C,H',W'], H' = (H*up + pad0 + pad1 - kh)//down + 1."""
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
