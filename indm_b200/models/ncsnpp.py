"""NCSN++ / DDPM++ score network, drop-in for the reference's `@register_model(name='ncsnpp')` class
(models/ncsnpp.py:34-414): constructed as `NCSNpp(config)`, same parameter names and order (`all_modules`
ModuleList, state-dict compatible), `forward(x, time_cond)` with NCHW FP32 tensors in and out.

The forward pass runs on the hand-written sm_100a kernels through `ScoreEngine` (no PyTorch compute, no CPU path).
Supported option set = what the INDM configs use: resblock_type='biggan', progressive='none',
progressive_input in {'none', 'residual'}, embedding_type in {'positional', 'fourier'}, conditional=True.
"""
import torch
import torch.nn as nn

from . import layers, layerspp, utils
from .. import precision
from .engine import ScoreEngine


class _EngineFunction(torch.autograd.Function):
    """Bridges the hand-written engine into torch.autograd so the reference's call sites keep working unchanged:
    `torch.autograd.grad(fn_eps, x)` of likelihood.py:32-35 and `.backward()` of losses.py:250 reach `ScoreEngine.vjp`,
    which replays the explicit backward plan over the activations the forward left in HBM.  In training the parameter
    gradients are accumulated by the plan straight into `param.grad` (the `anchor` input only makes autograd call us)."""

    @staticmethod
    def forward(ctx, x, time_cond, scale, net, train, anchor):
        eng = net.engine(x.shape[0])
        out = eng.forward(x.float(), time_cond.float(), scale, train=train).clone()
        eng.precapture_backward(train)
        ctx.eng, ctx.token, ctx.train, ctx.need_x = eng, eng.forward_count, train, x.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.eng
        if eng.forward_count != ctx.token:
            raise RuntimeError('indm_b200: the engine ran another forward before this backward; activations were overwritten '
                               '(call backward right after the forward it belongs to)')
        gx = eng.vjp(grad_out.contiguous().float(), train=ctx.train)
        return (gx.clone() if ctx.need_x else None), None, None, None, None, None


@utils.register_model(name='ncsnpp')
class NCSNpp(nn.Module):
    """NCSN++ model"""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        self.register_buffer('sigmas', torch.tensor(utils.get_sigmas(config)))
        self.nf = nf = m.nf
        ch_mult = m.ch_mult
        self.num_res_blocks = num_res_blocks = m.num_res_blocks
        self.attn_resolutions = attn_resolutions = m.attn_resolutions
        self.attention = attention = m.attention
        dropout = m.dropout
        self.num_resolutions = num_resolutions = len(ch_mult)
        self.all_resolutions = all_resolutions = [config.data.image_size // (2 ** i) for i in range(num_resolutions)]
        self.conditional = m.conditional
        fir, fir_kernel = m.fir, m.fir_kernel
        self.skip_rescale = skip_rescale = m.skip_rescale
        self.resblock_type = m.resblock_type.lower()
        self.auxiliary_resblock = m.auxiliary_resblock
        self.progressive = m.progressive.lower()
        self.progressive_input = progressive_input = m.progressive_input.lower()
        self.embedding_type = embedding_type = m.embedding_type.lower()
        init_scale = m.init_scale
        if self.resblock_type != 'biggan' or self.progressive != 'none' or progressive_input not in ('none', 'residual') \
                or not m.conditional or m.fourier_feature or not m.auxiliary_resblock or m.nonlinearity.lower() != 'swish':
            raise NotImplementedError('indm_b200 NCSNpp covers the option set of the INDM configs only '
                                      '(biggan res-blocks, swish, progressive=none, progressive_input none/residual)')
        act = 'swish'
        modules = []
        if embedding_type == 'fourier':
            assert config.training.continuous, "Fourier features are only used for continuous training."
            modules.append(layerspp.GaussianFourierProjection(embedding_size=nf, scale=m.fourier_scale))
            embed_dim = 2 * nf
        elif embedding_type == 'positional':
            embed_dim = nf
        else:
            raise ValueError(f'embedding type {embedding_type} unknown.')
        modules.append(layers.LinearParams(embed_dim, nf * 4))
        modules.append(layers.LinearParams(nf * 4, nf * 4))

        def ResnetBlock(**kw):
            return layerspp.ResnetBlockBigGANpp(act=act, dropout=dropout, fir=fir, fir_kernel=fir_kernel, init_scale=init_scale,
                                                skip_rescale=skip_rescale, temb_dim=nf * 4, **kw)

        def AttnBlock(channels):
            return layerspp.AttnBlockpp(channels=channels, init_scale=init_scale, skip_rescale=skip_rescale)

        channels = config.data.num_channels
        input_pyramid_ch = channels
        modules.append(layers.conv3x3(channels, nf))
        hs_c = [nf]
        in_ch = nf
        for i_level in range(num_resolutions):
            for _ in range(num_res_blocks):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch, out_ch=out_ch))
                in_ch = out_ch
                if all_resolutions[i_level] in attn_resolutions and attention:
                    modules.append(AttnBlock(channels=in_ch))
                hs_c.append(in_ch)
            if i_level != num_resolutions - 1:
                modules.append(ResnetBlock(down=True, in_ch=in_ch))
                if progressive_input == 'residual':
                    modules.append(layerspp.Downsample(in_ch=input_pyramid_ch, out_ch=in_ch, fir_kernel=fir_kernel))
                    input_pyramid_ch = in_ch
                hs_c.append(in_ch)
        in_ch = hs_c[-1]
        modules.append(ResnetBlock(in_ch=in_ch))
        modules.append(AttnBlock(channels=in_ch))
        modules.append(ResnetBlock(in_ch=in_ch))
        for i_level in reversed(range(num_resolutions)):
            for _ in range(num_res_blocks + 1):
                out_ch = nf * ch_mult[i_level]
                modules.append(ResnetBlock(in_ch=in_ch + hs_c.pop(), out_ch=out_ch))
                in_ch = out_ch
            if all_resolutions[i_level] in attn_resolutions and attention:
                modules.append(AttnBlock(channels=in_ch))
            if i_level != 0:
                modules.append(ResnetBlock(in_ch=in_ch, up=True))
        assert not hs_c
        modules.append(layers.GroupNormParams(in_ch))
        modules.append(layers.conv3x3(in_ch, channels, init_scale=init_scale))
        self.all_modules = nn.ModuleList(modules)
        self._engines = {}
        self.compute_mode = 'auto'   # 'auto' (indm_b200/precision.py policy), or 'bf16' / 'tf32' to force one arithmetic

    # ---------------------------------------------------------------------------------------------------------------
    def engine(self, batch, mode=None, infer=False):
        """infer=True: the forward-only plan (ScoreEngine(pp=True)) the samplers run; the default plan also serves the backward"""
        mode = precision.resolve('score', mode or self.compute_mode, 'training' if self.training else 'sampling')
        dev = next(self.parameters()).device
        infer = bool(infer) and mode == 'bf16'
        key = (int(batch), mode, str(dev), infer)
        eng = self._engines.get(key)
        if eng is None:
            eng = ScoreEngine(self, batch, mode=mode, device=dev, pp=infer)
            self._engines[key] = eng
        return eng

    def forward(self, x, time_cond):
        if not x.is_cuda:
            raise RuntimeError('indm_b200.NCSNpp.forward needs CUDA tensors: there is no CPU / PyTorch fallback path')
        anchor = None
        need_grad = False
        if torch.is_grad_enabled():
            anchor = next((p for p in self.parameters() if p.requires_grad), None) if self.training else None
            need_grad = x.requires_grad or anchor is not None
        eng = self.engine(x.shape[0], infer=not need_grad)     # nothing to differentiate: the forward-only plan
        scale = None
        if self.config.model.scale_by_sigma:
            if self.embedding_type == 'fourier':
                used_sigmas = time_cond
            else:
                used_sigmas = self.sigmas[time_cond.long()].float()
            scale = 1.0 / used_sigmas.float()
        if need_grad:
            return _EngineFunction.apply(x, time_cond, scale, self, anchor is not None, anchor)
        out = eng.forward(x.float(), time_cond.float(), scale, train=self.training)
        return out.clone()
