"""Parameter containers and initialisers of the NCSN++ score network.

The modules here own parameters under exactly the names the reference uses (state-dict / checkpoint compatible,
SURVEY.md appendix B) but carry no PyTorch compute: the forward pass is executed by `indm_b200.models.engine`
through the C-ABI CUDA library.  Initialisers restate models/layers.py:53-91 (`variance_scaling`, `default_init`).
"""
import numpy as np
import torch
import torch.nn as nn


def variance_scaling(scale, mode, distribution, in_axis=1, out_axis=0, dtype=torch.float32, device='cpu'):
    """JAX-style variance scaling (models/layers.py:53-85)."""

    def _compute_fans(shape):
        receptive = np.prod(shape) / shape[in_axis] / shape[out_axis]
        return shape[in_axis] * receptive, shape[out_axis] * receptive

    def init(shape, dtype=dtype, device=device):
        fan_in, fan_out = _compute_fans(shape)
        denom = {'fan_in': fan_in, 'fan_out': fan_out, 'fan_avg': (fan_in + fan_out) / 2}[mode]
        variance = scale / denom
        if distribution == 'normal':
            return torch.randn(*shape, dtype=dtype, device=device) * np.sqrt(variance)
        if distribution == 'uniform':
            return (torch.rand(*shape, dtype=dtype, device=device) * 2. - 1.) * np.sqrt(3 * variance)
        raise ValueError('invalid distribution for variance scaling initializer')

    return init


def default_init(scale=1.):
    """DDPM initialiser; scale 0 means 1e-10 (models/layers.py:88-91)."""
    scale = 1e-10 if scale == 0 else scale
    return variance_scaling(scale, 'fan_avg', 'uniform')


class Conv2dParams(nn.Module):
    """weight [out, in, k, k] + bias [out], DDPM init (models/layers.py:100-124 ddpm_conv1x1 / ddpm_conv3x3)."""

    def __init__(self, in_ch, out_ch, kernel, init_scale=1., stride=1, padding=None):
        super().__init__()
        self.in_ch, self.out_ch, self.kernel, self.stride = in_ch, out_ch, kernel, stride
        self.padding = kernel // 2 if padding is None else padding
        self.weight = nn.Parameter(default_init(init_scale)((out_ch, in_ch, kernel, kernel)))
        self.bias = nn.Parameter(torch.zeros(out_ch))


def conv3x3(in_ch, out_ch, init_scale=1.):
    return Conv2dParams(in_ch, out_ch, 3, init_scale)


def conv1x1(in_ch, out_ch, init_scale=1.):
    return Conv2dParams(in_ch, out_ch, 1, init_scale)


class LinearParams(nn.Module):
    """weight [out, in] + bias [out] with default_init and zero bias (models/ncsnpp.py:90-95, layerspp.py:240-242)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.weight = nn.Parameter(default_init()((out_dim, in_dim)))
        self.bias = nn.Parameter(torch.zeros(out_dim))


class GroupNormParams(nn.Module):
    """nn.GroupNorm(num_groups=min(C // 4, 32), eps=1e-6) parameters (models/layerspp.py:232)."""

    def __init__(self, channels):
        super().__init__()
        self.num_groups = min(channels // 4, 32)
        self.num_channels = channels
        self.eps = 1e-6
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class NIN(nn.Module):
    """W [in, out], b [out] (models/layers.py:546-555)."""

    def __init__(self, in_dim, num_units, init_scale=0.1):
        super().__init__()
        self.W = nn.Parameter(default_init(scale=init_scale)((in_dim, num_units)))
        self.b = nn.Parameter(torch.zeros(num_units))
