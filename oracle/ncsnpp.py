"""CPU restatement of the NCSN++ / DDPM++ score network forward (torch-CPU FP32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows `NCSNpp.__init__` / `NCSNpp.forward` (models/ncsnpp.py:38-414) for the option set the BASELINE configs
use: resblock_type='biggan', progressive='none', progressive_input in {'none','residual'},
embedding_type in {'positional','fourier'}, fir in {False, True}, conditional=True, skip_rescale=True.
Parameters come in as a flat dict keyed exactly like the reference's state-dict without the DataParallel
`module.` prefix.  Pinned by tests/golden/ncsnpp_*.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops as _ops


def module_plan(config):
    """Ordered list of module descriptors, restating the constructor's `modules.append` sequence
    (models/ncsnpp.py:72-249).  Each entry: (kind, dict)."""
    m = config.model
    nf, ch_mult, nres = m.nf, tuple(m.ch_mult), m.num_res_blocks
    nlev = len(ch_mult)
    all_res = [config.data.image_size // (2 ** i) for i in range(nlev)]
    attn_res = tuple(m.attn_resolutions)
    pin = m.progressive_input.lower()
    assert m.resblock_type.lower() == 'biggan' and m.progressive.lower() == 'none' and pin in ('none', 'residual')
    assert m.conditional and not m.fourier_feature and m.auxiliary_resblock
    plan = []
    if m.embedding_type.lower() == 'fourier':
        plan.append(('fourier', dict(size=nf)))
        embed_dim = 2 * nf
    else:
        embed_dim = nf
    plan.append(('linear', dict(cin=embed_dim, cout=nf * 4)))
    plan.append(('linear', dict(cin=nf * 4, cout=nf * 4)))
    ch = config.data.num_channels
    plan.append(('conv3x3', dict(cin=ch, cout=nf)))
    hs_c = [nf]
    in_ch = nf
    pyr_ch = ch
    for lv in range(nlev):
        for _ in range(nres):
            out_ch = nf * ch_mult[lv]
            plan.append(('res', dict(cin=in_ch, cout=out_ch, up=False, down=False)))
            in_ch = out_ch
            if all_res[lv] in attn_res and m.attention:
                plan.append(('attn', dict(c=in_ch)))
            hs_c.append(in_ch)
        if lv != nlev - 1:
            plan.append(('res', dict(cin=in_ch, cout=in_ch, up=False, down=True)))
            if pin == 'residual':
                plan.append(('pyr_down', dict(cin=pyr_ch, cout=in_ch)))
                pyr_ch = in_ch
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    plan.append(('res', dict(cin=in_ch, cout=in_ch, up=False, down=False)))
    plan.append(('attn', dict(c=in_ch)))
    plan.append(('res', dict(cin=in_ch, cout=in_ch, up=False, down=False)))
    for lv in reversed(range(nlev)):
        for _ in range(nres + 1):
            out_ch = nf * ch_mult[lv]
            plan.append(('res', dict(cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False)))
            in_ch = out_ch
        if all_res[lv] in attn_res and m.attention:
            plan.append(('attn', dict(c=in_ch)))
        if lv != 0:
            plan.append(('res', dict(cin=in_ch, cout=in_ch, up=True, down=False)))
    assert not hs_c
    plan.append(('gn', dict(c=in_ch)))
    plan.append(('conv3x3', dict(cin=in_ch, cout=ch)))
    return plan


def param_shapes(config):
    """[(state-dict key, shape)] in the reference's state_dict order (buffers included)."""
    out = [('sigmas', (config.model.num_scales,))]
    for i, (kind, a) in enumerate(module_plan(config)):
        p = f'all_modules.{i}.'
        if kind == 'fourier':
            out.append((p + 'W', (a['size'],)))
        elif kind == 'linear':
            out += [(p + 'weight', (a['cout'], a['cin'])), (p + 'bias', (a['cout'],))]
        elif kind == 'conv3x3':
            out += [(p + 'weight', (a['cout'], a['cin'], 3, 3)), (p + 'bias', (a['cout'],))]
        elif kind == 'gn':
            out += [(p + 'weight', (a['c'],)), (p + 'bias', (a['c'],))]
        elif kind == 'pyr_down':
            out += [(p + 'Conv2d_0.weight', (a['cout'], a['cin'], 3, 3)), (p + 'Conv2d_0.bias', (a['cout'],))]
        elif kind == 'attn':
            c = a['c']
            out += [(p + 'GroupNorm_0.weight', (c,)), (p + 'GroupNorm_0.bias', (c,))]
            for j in range(4):
                out += [(p + f'NIN_{j}.W', (c, c)), (p + f'NIN_{j}.b', (c,))]
        elif kind == 'res':
            ci, co = a['cin'], a['cout']
            out += [(p + 'GroupNorm_0.weight', (ci,)), (p + 'GroupNorm_0.bias', (ci,)),
                    (p + 'Conv_0.weight', (co, ci, 3, 3)), (p + 'Conv_0.bias', (co,)),
                    (p + 'Dense_0.weight', (co, 4 * config.model.nf)), (p + 'Dense_0.bias', (co,)),
                    (p + 'GroupNorm_1.weight', (co,)), (p + 'GroupNorm_1.bias', (co,)),
                    (p + 'Conv_1.weight', (co, co, 3, 3)), (p + 'Conv_1.bias', (co,))]
            if ci != co or a['up'] or a['down']:
                out += [(p + 'Conv_2.weight', (co, ci, 1, 1)), (p + 'Conv_2.bias', (co,))]
    return out


def synth_params(config, seed=0, dtype=np.float32):
    """Deterministic, platform-independent synthetic weights (numpy PCG64), used by golden vectors, parity tests
    and the benchmark ("random-init weights of that architecture").  Unlike the reference initialiser, the
    ≈0-initialised tensors (Conv_1, NIN_3, head conv: init_scale=0 → 1e-10, models/layers.py:88-91) get ordinary
    fan-avg magnitudes so that residual branches are visible to parity checks.
    GroupNorm affine and biases are randomised too."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in param_shapes(config):
        leaf = name.split('.')[-1]
        if name == 'sigmas':
            # models/utils.py:46-57 get_sigmas
            v = np.exp(np.linspace(np.log(config.model.sigma_max), np.log(config.model.sigma_min), config.model.num_scales))
            sd[name] = v.astype(np.float64)
            continue
        if name.endswith('.W') and len(shape) == 1:   # GaussianFourierProjection, models/layerspp.py:50
            v = rng.standard_normal(shape) * config.model.fourier_scale
        elif len(shape) == 1:
            # 1-D `weight` only occurs on GroupNorm modules (incl. the bare head GroupNorm `all_modules.N.weight`)
            if leaf == 'weight':
                v = 1.0 + 0.1 * rng.standard_normal(shape)
            else:
                v = 0.05 * rng.standard_normal(shape)
        else:
            if len(shape) == 4:
                fan_in, fan_out = shape[1] * shape[2] * shape[3], shape[0] * shape[2] * shape[3]
            elif leaf == 'W':
                fan_in, fan_out = shape[0], shape[1]
            else:
                fan_in, fan_out = shape[1], shape[0]
            lim = math.sqrt(3.0 * 1.0 / ((fan_in + fan_out) / 2))
            v = rng.uniform(-lim, lim, size=shape)
        sd[name] = v.astype(dtype)
    return sd


def damp_zero_init(sd, factor=0.1):
    """Scale the tensors the reference initialises to ~0 (init_scale=0: every res-block's Conv_1, every attention block's NIN_3
    and the head conv, models/layers.py:88-91, models/ncsnpp.py:230-232) by `factor`, in place.  `synth_params` gives them ordinary
    fan-avg magnitudes so residual branches are visible to parity checks; for the probability-flow ODE that makes the random-weight
    network so rough that the reference's own NLL needs 1658 RK45 evaluations — with the damping the fixture behaves like a
    network near its initialisation (the regime INDM trains from), and parity at 0.01 bpd is a statement about arithmetic, not
    about which way a chaotic trajectory happened to branch."""
    last = max(int(k.split('.')[1]) for k in sd if k.startswith('all_modules.'))
    for k in sd:
        if k.endswith('Conv_1.weight') or k.endswith('NIN_3.W') or k == f'all_modules.{last}.weight':
            sd[k] = (sd[k] * factor).astype(sd[k].dtype)
    return sd


def _gn(x, w, b, c):
    return F.group_norm(x, min(c // 4, 32), w, b, eps=1e-6)   # models/layerspp.py:232 etc.


def _nin(x, W, b):   # models/layers.py:552-555
    return torch.einsum('bchw,cd->bdhw', x, W) + b[None, :, None, None]


def _fir_k(k, gain=1.0, factor=1):   # models/up_or_down_sampling.py:181-188
    k = np.asarray(k, dtype=np.float32)
    k = np.outer(k, k)
    k /= np.sum(k)
    return k * (gain * factor ** 2)


def _upfirdn(x, k, **kw):
    return torch.from_numpy(_ops.upfirdn2d(x.numpy(), k, **kw))


def _resample(x, up, down, fir, fir_kernel):
    """models/layerspp.py:258-271 + models/up_or_down_sampling.py:59-69,195-257"""
    n, c, h, w = x.shape
    if up:
        if fir:
            return _upfirdn(x, _fir_k(fir_kernel, factor=2), up=2, pad=(2, 1))
        return x.reshape(n, c, h, 1, w, 1).repeat(1, 1, 1, 2, 1, 2).reshape(n, c, h * 2, w * 2)
    if down:
        if fir:
            return _upfirdn(x, _fir_k(fir_kernel), down=2, pad=(1, 1))
        return x.reshape(n, c, h // 2, 2, w // 2, 2).mean(dim=(3, 5))
    return x


def forward(config, params, x, time_cond, taps=None):
    """Restates NCSNpp.forward (models/ncsnpp.py:251-414).  `params` maps state-dict keys to torch CPU tensors.
    `taps` (optional dict) receives named intermediate activations for layer-wise kernel debugging."""
    m = config.model
    P = params
    plan = module_plan(config)
    act = F.silu
    fir, firk = bool(m.fir), list(m.fir_kernel)
    idx = 0

    def g(i, name):
        return P[f'all_modules.{i}.{name}']

    # --- time embedding (:255-274)
    if m.embedding_type.lower() == 'fourier':
        used_sigmas = time_cond
        xp = torch.log(used_sigmas)[:, None] * g(idx, 'W')[None, :] * 2 * np.pi   # models/layerspp.py:52-54
        temb = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
        idx += 1
    else:
        used_sigmas = P['sigmas'][time_cond.long()]
        half = m.nf // 2                                       # models/layers.py:515-529
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
        e = time_cond.float()[:, None] * e[None, :]
        temb = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    temb = F.linear(temb, g(idx, 'weight'), g(idx, 'bias')); idx += 1
    temb = F.linear(act(temb), g(idx, 'weight'), g(idx, 'bias')); idx += 1
    if taps is not None:
        taps['temb'] = temb
    if not config.data.centered:
        x = 2 * x - 1.
    input_pyramid = x if m.progressive_input.lower() != 'none' else None

    def res(i, a, x, temb):   # models/layerspp.py:255-287
        p = f'all_modules.{i}.'
        h = act(_gn(x, P[p + 'GroupNorm_0.weight'], P[p + 'GroupNorm_0.bias'], a['cin']))
        h = _resample(h, a['up'], a['down'], fir, firk)
        x = _resample(x, a['up'], a['down'], fir, firk)
        h = F.conv2d(h, P[p + 'Conv_0.weight'], P[p + 'Conv_0.bias'], padding=1)
        h = h + F.linear(act(temb), P[p + 'Dense_0.weight'], P[p + 'Dense_0.bias'])[:, :, None, None]
        h = act(_gn(h, P[p + 'GroupNorm_1.weight'], P[p + 'GroupNorm_1.bias'], a['cout']))
        # Dropout_0: identity in eval mode (sampling / likelihood)
        h = F.conv2d(h, P[p + 'Conv_1.weight'], P[p + 'Conv_1.bias'], padding=1)
        if (p + 'Conv_2.weight') in P:
            x = F.conv2d(x, P[p + 'Conv_2.weight'], P[p + 'Conv_2.bias'])
        return (x + h) / np.sqrt(2.)

    def attn(i, a, x):   # models/layerspp.py:88-104
        p = f'all_modules.{i}.'
        B, C, H, W = x.shape
        h = _gn(x, P[p + 'GroupNorm_0.weight'], P[p + 'GroupNorm_0.bias'], C)
        q = _nin(h, P[p + 'NIN_0.W'], P[p + 'NIN_0.b'])
        k = _nin(h, P[p + 'NIN_1.W'], P[p + 'NIN_1.b'])
        v = _nin(h, P[p + 'NIN_2.W'], P[p + 'NIN_2.b'])
        w = torch.einsum('bchw,bcij->bhwij', q, k) * (int(C) ** (-0.5))
        w = F.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
        h = torch.einsum('bhwij,bcij->bchw', w, v)
        h = _nin(h, P[p + 'NIN_3.W'], P[p + 'NIN_3.b'])
        return (x + h) / np.sqrt(2.)

    def run(i, x, temb=None):
        kind, a = plan[i]
        if kind == 'res':
            y = res(i, a, x, temb)
        elif kind == 'attn':
            y = attn(i, a, x)
        elif kind == 'conv3x3':
            y = F.conv2d(x, g(i, 'weight'), g(i, 'bias'), padding=1)
        elif kind == 'gn':
            y = _gn(x, g(i, 'weight'), g(i, 'bias'), a['c'])
        elif kind == 'pyr_down':
            # layerspp.Downsample(fir=True, with_conv=True) → up_or_down_sampling.Conv2d(down=True) (:44-55)
            # → conv_downsample_2d (:144-178): FIR pad (2,2) then stride-2 VALID conv, + bias
            xk = _upfirdn(x, _fir_k(firk), pad=(2, 2))
            y = F.conv2d(xk, g(i, 'Conv2d_0.weight'), stride=2) + g(i, 'Conv2d_0.bias').reshape(1, -1, 1, 1)
        else:
            raise ValueError(kind)
        if taps is not None:
            taps[f'm{i}'] = y
        return y

    nlev = len(m.ch_mult)
    attn_res = tuple(m.attn_resolutions)
    hs = [run(idx, x)]; idx += 1
    for lv in range(nlev):
        for _ in range(m.num_res_blocks):
            h = run(idx, hs[-1], temb); idx += 1
            if h.shape[-1] in attn_res and m.attention:
                h = run(idx, h); idx += 1
            hs.append(h)
        if lv != nlev - 1:
            h = run(idx, hs[-1], temb); idx += 1
            if m.progressive_input.lower() == 'residual':
                input_pyramid = run(idx, input_pyramid); idx += 1
                input_pyramid = (input_pyramid + h) / np.sqrt(2.)
                h = input_pyramid
            hs.append(h)
    h = hs[-1]
    h = run(idx, h, temb); idx += 1
    h = run(idx, h); idx += 1
    h = run(idx, h, temb); idx += 1
    for lv in reversed(range(nlev)):
        for _ in range(m.num_res_blocks + 1):
            h = run(idx, torch.cat([h, hs.pop()], dim=1), temb); idx += 1
        if h.shape[-1] in attn_res and m.attention:
            h = run(idx, h); idx += 1
        if lv != 0:
            h = run(idx, h, temb); idx += 1
    assert not hs
    h = act(run(idx, h)); idx += 1
    h = run(idx, h); idx += 1
    assert idx == len(plan)
    if m.scale_by_sigma:
        h = h / used_sigmas.reshape(-1, 1, 1, 1)
    return h


def score_fn(config, sde, params, x, t):
    """models/utils.py:get_score_fn (:140-197), continuous=True, eval mode."""
    from . import sde as _sde
    if isinstance(sde, _sde.VP):
        labels = t * 999
        out = forward(config, params, x, labels)
        std = sde.marginal_prob(torch.zeros_like(x), t)[1]
        if config.training.ddpm_score:
            out = -out / std[:, None, None, None]
        return out
    labels = sde.marginal_prob(torch.zeros_like(x), t)[1]
    return forward(config, params, x, labels)


def to_torch(params_np):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params_np.items()}
