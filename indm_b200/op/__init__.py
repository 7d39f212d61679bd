"""Drop-in for the reference's `op` package (op/__init__.py): same names, same signatures, same autograd behaviour
(differentiable to second order w.r.t. `input`), executed by the C-ABI kernels `indm_upfirdn2d_f32` / `indm_bias_act_f32`."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d

__all__ = ['FusedLeakyReLU', 'fused_leaky_relu', 'upfirdn2d']
