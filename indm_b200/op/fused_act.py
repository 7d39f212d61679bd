"""`fused_leaky_relu` / `FusedLeakyReLU` — the reference's op/fused_act.py:20-97 on the C-ABI kernel `indm_bias_act_f32`
(which has the signature of the reference's pybind `fused_bias_act`, op/fused_bias_act.cpp:11-17).

y = leaky_relu(x + bias[channel], negative_slope) * scale, bias broadcast over dim 1.  Differentiable to second order
like the reference: the first derivative reuses the kernel with grad=1 and the forward OUTPUT as sign reference
(op/fused_act.py:27-29), the second derivative is the same call again (:44-46).  CUDA tensors only — no CPU path; note
the reference's CPU branch hard-codes slope 0.2 (op/fused_act.py:91), a quirk the oracle (`oracle/ops.py`) reproduces.
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _lib as L


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    if not x.is_cuda:
        raise RuntimeError('indm_b200.op.fused_leaky_relu needs CUDA tensors: there is no CPU / PyTorch fallback path')
    x = x.contiguous().float()
    y = torch.empty_like(x)
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.shape[i]
    b = bias.contiguous().float() if bias is not None and bias.numel() else None
    r = ref.contiguous().float() if ref is not None and ref.numel() else None
    L.call('indm_bias_act_f32', L.ptr(x), L.ptr(b), L.ptr(r), L.ptr(y), x.numel(), b.numel() if b is not None else 0, step_b,
           act, grad, float(alpha), float(scale))
    return y


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = _bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        dim = [0] + list(range(2, grad_input.ndim))
        grad_bias = grad_input.sum(dim).detach()
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = _bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):
    """op/fused_act.py:74-83"""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:86-97 (CUDA branch)."""
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
