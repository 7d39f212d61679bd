"""NCSN++ blocks as parameter containers (models/layerspp.py of the reference); compute lives in engine.py."""
import torch
import torch.nn as nn

from .layers import Conv2dParams, GroupNormParams, LinearParams, NIN, conv1x1, conv3x3, default_init


class GaussianFourierProjection(nn.Module):
    """Frozen Gaussian Fourier features for noise levels (models/layerspp.py:45-54)."""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)


class AttnBlockpp(nn.Module):
    """models/layerspp.py:75-104"""

    def __init__(self, channels, skip_rescale=False, init_scale=0.):
        super().__init__()
        self.GroupNorm_0 = GroupNormParams(channels)
        self.NIN_0 = NIN(channels, channels)
        self.NIN_1 = NIN(channels, channels)
        self.NIN_2 = NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=init_scale)
        self.skip_rescale = skip_rescale
        self.channels = channels


class ResnetBlockBigGANpp(nn.Module):
    """models/layerspp.py:225-287"""

    def __init__(self, act, in_ch, out_ch=None, temb_dim=None, up=False, down=False, dropout=0.1, fir=False,
                 fir_kernel=(1, 3, 3, 1), skip_rescale=True, init_scale=0.):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        self.GroupNorm_0 = GroupNormParams(in_ch)
        self.up, self.down, self.fir, self.fir_kernel = up, down, fir, tuple(fir_kernel)
        self.Conv_0 = conv3x3(in_ch, out_ch)
        if temb_dim is not None:
            self.Dense_0 = LinearParams(temb_dim, out_ch)
        self.GroupNorm_1 = GroupNormParams(out_ch)
        self.dropout = dropout
        self.Conv_1 = conv3x3(out_ch, out_ch, init_scale=init_scale)
        if in_ch != out_ch or up or down:
            self.Conv_2 = conv1x1(in_ch, out_ch)
        self.skip_rescale = skip_rescale
        self.act = act
        self.in_ch, self.out_ch = in_ch, out_ch


class Downsample(nn.Module):
    """FIR + strided-conv input-pyramid downsampler, the only Downsample variant the INDM configs build
    (models/layerspp.py:142-176 with fir=True, with_conv=True -> up_or_down_sampling.Conv2d(down=True), :23-56)."""

    def __init__(self, in_ch, out_ch, fir_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.Conv2d_0 = nn.Module()
        self.Conv2d_0.weight = nn.Parameter(default_init()((out_ch, in_ch, 3, 3)))
        self.Conv2d_0.bias = nn.Parameter(torch.zeros(out_ch))
        self.fir_kernel = tuple(fir_kernel)
        self.in_ch, self.out_ch = in_ch, out_ch
