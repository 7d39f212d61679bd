"""Wolf flow (conditional Residual Flow + Gaussian posterior with a latent flow prior) for the INDM hot path.

`WolfCore` is a parameter container whose state-dict is key-for-key the reference's (flow_models/wolf/wolf.py:18-145,
modules/generators/generator.py:88-109, flows/resflow/resflow_.py:20-518, layers/iresblock.py:14-60,
layers/base/lipschitz.py:321-441, modules/discriminators/gaussian.py:14-21, modules/encoders/global_encoder.py:12-31,
nnet/resnets/resnet_batchnorm.py:18-47, modules/discriminators/priors/flow.py:16-217).
Compute runs on the C-ABI kernels through `FlowEngine` (no PyTorch compute, no CPU path):

  reverse (sampling, wolf.py:82-89):  h = prior-flow(eps) in ONE launch (indm_prior_flow), then for each of the 32
      iResBlocks the fixed-point inverse x <- y - g(x; h) where g is three tensor-core implicit GEMMs
      (3x3 c->512 with Sin epilogue, 1x1 512->512 with the per-sample conditioning bias W2(Ah+a)+b2 as row bias and Sin
      epilogue, 3x3 512->c whose epilogue writes y - g directly in NCHW) and the stop rule is a device-side reduction.
  forward without log-det (flow_model.py:57-60 with log_det != 0): y = x + g(x; h) with the same three GEMMs.

The Lipschitz-normalised weights W / max(1, |row|_1 / 0.98) are computed once per weight version at pack time, not
inside every call as the reference does (lipschitz.py:350-363).
"""
import ctypes
import os
import threading
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .. import precision

COEFF = 0.98


def _poisson_1mcdf(lamb, k, offset):
    """P(N >= k - offset) for N ~ Poisson(lamb), 1 for k <= offset (iresblock.py:306-318)"""
    if k <= offset:
        return 1.
    k = k - offset
    acc = 1.
    for i in range(1, k):
        acc += lamb ** i / math.factorial(i)
    return 1 - math.exp(-lamb) * acc


def series_coefficients(n, training, lamb=2.0, n_exact_terms=2):
    """Russian-roulette power-series weights of iResBlock._logdetgrad (iresblock.py:114-132) for one Poisson draw n:
    K = n + offset terms, c_k = 1 / P(N >= k - offset); offset = n_exact_terms in training, 20 in eval."""
    offset = n_exact_terms if training else 20
    K = n + offset
    return K, [1.0 / _poisson_1mcdf(lamb, k, offset) for k in range(1, K + 1)]


# ------------------------------------------------------------------------------------------------ parameter containers
class _LopConv(nn.Module):
    def __init__(self, cin, cout, k, cond_dim=None):
        super().__init__()
        bound = 1.0 / math.sqrt(cin * k * k)
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        self.register_buffer('scale', torch.tensor(0.))
        if cond_dim is not None:
            self.h_net = nn.Module()
            self.h_net.net = nn.Module()
            b2 = 1.0 / math.sqrt(cond_dim)
            self.h_net.net.weight = nn.Parameter(torch.empty(cin, cond_dim).uniform_(-b2, b2))
            self.h_net.net.bias = nn.Parameter(torch.empty(cin).uniform_(-b2, b2))


class _Placeholder(nn.Module):
    """parameter-free chain element (Sin activation / SqueezeLayer): keeps ModuleList indices aligned with the reference"""

    def __init__(self, what):
        super().__init__()
        self.what = what


class _IResBlock(nn.Module):
    def __init__(self, c, idim, first):
        super().__init__()
        self.geom_p = nn.Parameter(torch.tensor(np.log(0.5) - np.log(1. - 0.5), dtype=torch.float32))
        self.lamb = nn.Parameter(torch.tensor(2.))
        # no loss term reaches these two (iresblock.py:36-42: learn_p is off, lamb only parameterises the Poisson draw): the
        # reference's AdamW finds .grad None and skips them — no weight decay either (indm_b200/losses.py:_stepped)
        self.geom_p._indm_no_grad = True
        self.lamb._indm_no_grad = True
        self.register_buffer('last_n_samples', torch.zeros(1))
        self.register_buffer('last_firmom', torch.zeros(1))
        self.register_buffer('last_secmom', torch.zeros(1))
        net = [] if first else [_Placeholder('sin')]
        net += [_LopConv(c, idim, 3), _Placeholder('sin'), _LopConv(idim, idim, 1, cond_dim=64), _Placeholder('sin'), _LopConv(idim, c, 3)]
        self.nnet = nn.ModuleList(net)
        self.first = first
        self.channels = c

    def convs(self):
        return [m for m in self.nnet if isinstance(m, _LopConv)]


class _Stack(nn.Module):
    def __init__(self, c, idim, n_blocks, squeeze, first_scale):
        super().__init__()
        chain = [_IResBlock(c, idim, first_scale and i == 0) for i in range(n_blocks)]
        if squeeze:
            chain.append(_Placeholder('squeeze'))
        self.chain = nn.ModuleList(chain)


class _ResidualFlow(nn.Module):
    def __init__(self, config, input_shape):
        super().__init__()
        nb = [int(v) for v in config.flow.nblocks.split('-')]
        c, h, w = input_shape
        stacks = []
        for i, n in enumerate(nb):
            stacks.append(_Stack(c, config.flow.intermediate_dim, n, i < len(nb) - 1, i == 0))
            c *= 4
        self.transforms = nn.ModuleList(stacks)
        self.n_blocks = nb


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))


class _ConvNoBias(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        bound = 1.0 / math.sqrt(cin * k * k)
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k).uniform_(-bound, bound))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _ConvNoBias(cin, cout, 3)
        self.bn1 = _BN(cout)
        self.conv2 = _ConvNoBias(cout, cout, 3)
        self.bn2 = _BN(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.ModuleList([_ConvNoBias(cin, cout, 1), _BN(cout)])
        self.stride = stride


class _Encoder(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.net = nn.Module()
        inp = p['in_planes']
        for lv, hid in enumerate(p['hidden_planes']):
            r = nn.Module()
            r.main = nn.ModuleList([_EncBlock(inp, hid, 1), _EncBlock(hid, hid, 2)])
            setattr(self.net, f'resnet{lv}', r)
            inp = hid
        top = nn.Module()
        bound = 1.0 / math.sqrt(inp)
        top.weight = nn.Parameter(torch.empty(p['out_planes'], inp, 1, 1).uniform_(-bound, bound))
        top.bias = nn.Parameter(torch.zeros(p['out_planes']))
        self.net.top = top


class _WNLinear(nn.Module):
    """legacy nn.utils.weight_norm layout: bias, weight_g [out,1], weight_v [out,in] (nnet/weight_norm.py:8-40)"""

    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Module()
        self.linear.bias = nn.Parameter(torch.zeros(cout))
        v = torch.randn(cout, cin) * 0.05
        self.linear.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True))
        self.linear.weight_v = nn.Parameter(v)


class _Lin(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        bound = 1.0 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.empty(cout, cin).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(cout))


class _Coupling(nn.Module):
    def __init__(self, d, hf):
        super().__init__()
        self.net = nn.Module()
        self.net.fc1 = _Lin(d // 2, hf)
        self.net.fc2 = _Lin(hf, hf)
        self.net.fc3 = _WNLinear(hf, d)


class _ActNorm1d(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.log_scale = nn.Parameter(torch.randn(d) * 0.05)
        self.bias = nn.Parameter(torch.zeros(d))


class _InvLinear(nn.Module):
    def __init__(self, d):
        super().__init__()
        w = torch.empty(d, d)
        nn.init.orthogonal_(w)
        self.weight = nn.Parameter(w)
        self.register_buffer('weight_inv', w.inverse().clone())

    def sync(self):
        self.weight_inv.copy_(self.weight.data.inverse())


class _PriorStep(nn.Module):
    def __init__(self, d, hf):
        super().__init__()
        self.actnorm = _ActNorm1d(d)
        self.linear = _InvLinear(d)
        self.unit = nn.Module()
        self.unit.coupling1_up = _Coupling(d, hf)
        self.unit.coupling1_dn = _Coupling(d, hf)
        self.unit.actnorm = _ActNorm1d(d)
        self.unit.coupling2_up = _Coupling(d, hf)
        self.unit.coupling2_dn = _Coupling(d, hf)


class WolfCore(nn.Module):
    """flow_models/wolf/wolf.py:18 — generator (ResidualFlow) + discriminator (encoder, fc, flow prior)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        wp = config.flow.wolf_params
        c, s = config.data.num_channels, config.data.image_size
        self.input_shape = (c * 4, s // 2, s // 2) if config.flow.squeeze else (c, s, s)
        self.generator = nn.Module()
        self.generator.flow = _ResidualFlow(config, self.input_shape)
        d = wp['discriminator']
        self.discriminator = nn.Module()
        self.discriminator.encoder = _Encoder(d['encoder'])
        self.discriminator.fc = _WNLinear(d['in_dim'], 2 * d['dim'])
        self.discriminator.prior = nn.Module()
        self.discriminator.prior.flow = nn.Module()
        pr = d['prior']
        self.discriminator.prior.flow.steps = nn.ModuleList([_PriorStep(pr['in_features'], pr['hidden_features'])
                                                             for _ in range(pr['num_steps'])])
        self.latent_dim = d['dim']
        self._engines = {}
        self.compute_mode = 'auto'    # indm_b200/precision.py policy; 'bf16' / 'tf32' force one arithmetic
        self._draws = 0               # number of prior draws made with the in-kernel generator (advances the Philox stream)

    @classmethod
    def from_params(cls, params, config=None):
        """WolfCore.from_params (wolf.py:132-145); `params` is the JSON dict (must equal config.flow.wolf_params)."""
        return cls(config)

    def add_config(self, config):
        self.config = config

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._lamb_epoch = getattr(self, '_lamb_epoch', 0) + 1      # engines re-read the iResBlocks' lamb on next use

    def blocks(self):
        """[(scale, block index, module)] in forward order"""
        out = []
        for s, st in enumerate(self.generator.flow.transforms):
            for b, m in enumerate(st.chain):
                if isinstance(m, _IResBlock):
                    out.append((s, b, m))
        return out

    def engine(self, batch, mode=None, leg=None):
        if leg is None:
            leg = 'training' if self.training else 'eval'
        mode = precision.resolve('flow', mode or self.compute_mode, leg)
        dev = next(self.parameters()).device
        key = (int(batch), mode, str(dev))
        e = self._engines.get(key)
        if e is None:
            e = FlowEngine(self, batch, mode, dev)
            self._engines[key] = e
        return e

    def forward(self, data, y=None, n_bits=8, nsamples=1, reverse=False, eval_logdet=True, *, eps=None, h=None, seed=0,
                atol=1e-5, rtol=1e-5, vareps=None, n_terms=None, estimator=None):
        """wolf.py:81-130.  reverse=True: sample h from the prior (or use `eps` / `h` if given) and invert the flow.
        reverse=False: h ~ q(h|x) (posterior encoder; `eps` = the reparameterisation noise, `h` overrides it), then the
        residual-flow forward; with eval_logdet the power-series log-det of every block and the KL term are returned as
        `-logdet_total - KL` exactly like the reference (`vareps` = per-block probe tensors, `n_terms` = per-block Poisson
        draws, both optional: defaults are the in-kernel Philox generator and numpy's global RNG, the generator the
        reference uses, iresblock.py:306)."""
        if not data.is_cuda:
            raise RuntimeError('indm_b200 WolfCore needs CUDA tensors: there is no CPU / PyTorch fallback path')
        eng = self.engine(data.shape[0], leg='reverse' if reverse else None)
        if reverse:
            if eps is None and h is None:
                self._draws += 1      # a fresh h ~ prior per call, like discriminator.sample_from_prior (wolf.py:83)
            return eng.reverse(data, eps=eps, h=h, seed=seed, offset=self._draws, atol=atol, rtol=rtol)
        if nsamples != 1:
            raise NotImplementedError('flow.train_k = 1 in every INDM config')
        if self.training:
            # joint training (losses.py:258-320): batch-statistics BatchNorm in the posterior encoder, the Neumann-series
            # log-det (n + 2 terms), and an explicit backward plan reached through torch.autograd
            if h is not None:
                raise NotImplementedError('training-mode flow forward samples h from the posterior')
            self._draws += 1
            if not eval_logdet:
                # losses.py:381-383: the latent under the just-updated flow, without log-det (still batch-statistics BatchNorm)
                with torch.no_grad():
                    hh, _ = eng.train_posterior(data, eps=eps, seed=seed, offset=self._draws)
                    return eng.forward_map(data, hh)
            anchor = next(p for p in self.parameters() if p.requires_grad)
            return _FlowTrainFunction.apply(data, anchor, self, dict(eps=eps, vareps=vareps, n_terms=n_terms, seed=seed, offset=self._draws))
        kl = None
        if h is None:
            self._draws += 1
            h, kl = eng.posterior(data, eps=eps, seed=seed, offset=self._draws)
        if not eval_logdet:
            return eng.forward_map(data, h)
        z, logpx = eng.forward_logdet(data, h, vareps=vareps, n_terms=n_terms, training=(estimator == 'train'), seed=seed,
                                      offset=self._draws)
        loss = -logpx                       # wolf.py:126-128: loss = -logdet - kl with logdet = logpx = -(sum of block log-dets)
        if kl is not None:
            loss = loss - kl
        return z, loss


class _FlowTrainFunction(torch.autograd.Function):
    """Bridges the flow engine's explicit training backward into torch.autograd, so that the reference's call site
    `torch.mean(losses).backward()` (losses.py:304) reaches it: forward returns (z, logdet - KL); backward accumulates every
    flow parameter gradient into `param.grad` (the `anchor` parameter only makes autograd call us)."""

    @staticmethod
    def forward(ctx, data, anchor, core, kw):
        eng = core.engine(data.shape[0])
        z, loss = eng.train_forward(data, **kw)
        eng.precapture_backward()
        ctx.eng, ctx.token = eng, eng.train_token
        return z, loss

    @staticmethod
    def backward(ctx, gz, gloss):
        eng = ctx.eng
        if eng.train_token != ctx.token:
            raise RuntimeError('indm_b200: the flow engine ran another training forward before this backward')
        hook = getattr(eng.core, '_before_backward', None)
        if hook is not None:
            hook()
        if gz is None:
            gz = torch.zeros((eng.N,) + tuple(eng.core.input_shape), device=eng.dev)
        if gloss is None:
            gloss = torch.zeros((eng.N,), device=eng.dev)
        eng.train_backward(gz.contiguous().float(), gloss.contiguous().float())
        return None, None, None, None


# ------------------------------------------------------------------------------------------------ engine
class FlowEngine:
    def __init__(self, core, batch, mode, dev):
        if dev.type != 'cuda':
            raise RuntimeError('FlowEngine needs a CUDA device')
        L.lib()
        self.core, self.N, self.mode, self.dev = core, int(batch), mode, dev
        self.dt = L.DTYPE_BF16 if mode == 'bf16' else L.DTYPE_TF32
        self.tdtype = torch.bfloat16 if mode == 'bf16' else torch.float32
        self.kchunk = 64 if mode == 'bf16' else 32
        self.idim = core.config.flow.intermediate_dim
        self.blocks = core.blocks()
        self.nb = core.generator.flow.n_blocks
        self._version = None
        self._version_blk = None
        self.iterations = []          # fixed-point iterations of the last reverse() call, per block
        N, idim = self.N, self.idim
        c0, h0, w0 = core.input_shape
        # operand / activation buffers, sized for the largest scale (scale 0) and reused by every block
        # tap packing of the few-channel 3x3 convs (csrc/flow_kernels.cu): K (first conv) / N (last conv) = 9 c packed values
        self.cs = [c0 * 4 ** s for s in range(len(self.nb))]
        self.kp = [((9 * c + self.kchunk - 1) // self.kchunk) * self.kchunk for c in self.cs]      # im2col row length
        # col2im row stride = the N extent of the idim -> 9c GEMMs, padded to whole 32-column slabs (27 -> 32, 108 -> 128) with zero
        # weight rows: a ragged N (27) took the generic scalar epilogue and ran the 134 MB-input GEMM at a fifth of the HBM roofline
        self.ld9 = [((9 * c + 31) // 32) * 32 for c in self.cs]
        self.a0 = [torch.empty((N, h0 >> s, w0 >> s, self.kp[s]), device=dev, dtype=self.tdtype) for s in range(len(self.nb))]
        self.o9 = [torch.empty((N, h0 >> s, w0 >> s, self.ld9[s]), device=dev) for s in range(len(self.nb))]
        self.u1 = torch.empty((N, h0, w0, idim), device=dev, dtype=self.tdtype)
        self.u2 = torch.empty((N, h0, w0, idim), device=dev, dtype=self.tdtype)
        self.flag = torch.zeros((1,), device=dev)
        self.h = torch.empty((N, core.latent_dim), device=dev)
        self.eps = torch.empty((N, core.latent_dim), device=dev)
        self.cond = torch.empty((N, len(self.blocks) * idim), device=dev)
        self.w = {}
        self.w2f = {}
        self._saved = None
        self._graphs, self._pool, self.use_graphs = {}, None, not os.environ.get('INDM_NO_GRAPHS')
        self._lamb_vers = -1
        self.prior_winvT = []
        self._ops, self._bufs = {}, {}
        self._alloc_weights()

    # ---- weights
    def _alloc_weights(self):
        dev, idim = self.dev, self.idim
        for s, b, m in self.blocks:
            c = m.channels
            kp = self.kp[s]
            z = lambda *shape: torch.zeros(shape, device=dev, dtype=self.tdtype)
            self.w[(s, b)] = dict(w1=z(idim, kp), b1=torch.empty((idim,), device=dev), w2=z(idim, idim), w3=z(self.ld9[s], idim),
                                  b3=torch.empty((c,), device=dev),
                                  w3v=z(idim, kp), w2d=z(idim, idim), w1v=z(self.ld9[s], idim))      # transposed packs for the VJP chain
        nblk = len(self.blocks)
        self.cond_w = torch.empty((nblk * idim, self.core.latent_dim), device=dev)
        self.cond_b = torch.empty((nblk * idim,), device=dev)
        # prior flow program
        self.prior_params = None
        self.prior_ops = {}

    def _round(self, w):
        if self.mode == 'bf16':
            return w.to(torch.bfloat16)
        return w.float().contiguous()     # TF32 mode keeps full fp32 operands: the GEMM kernel splits hi/lo itself (3xTF32)

    @staticmethod
    def _lop(w):
        """LopConv2d.compute_weight, domain = codomain = inf (lipschitz.py:350-359): rows scaled to L1 norm <= 0.98"""
        scale = w.abs().reshape(w.shape[0], -1).sum(dim=1)
        return w / torch.clamp(scale / COEFF, min=1.0).reshape(-1, 1, 1, 1), scale.max()

    def version(self):
        return sum(p._version for p in self.core.parameters()) + sum(b._version for b in self.core.buffers()) + L.param_epoch

    def _graphed(self, key, fn):
        """run `fn` (a fixed sequence of launches over engine-owned buffers) through a CUDA graph: eager on first use (lazy
        launch lists / buffers get built), captured on the second, replayed afterwards.  The flow passes issue thousands of
        small launches per step; replaying them removes the per-launch host cost that bounded the training step."""
        st = self._graphs.get(key)
        if st is None or torch.cuda.is_current_stream_capturing():
            fn()
            if st is None:
                self._graphs[key] = 1
            return
        if st == 1:
            # stream capture from inside an autograd worker thread gets invalidated (observed: cudaErrorStreamCaptureInvalidated
            # at capture_end when the flow backward is reached through loss.backward()); those calls stay eager — the flow
            # backward is GPU-bound (76 ms eager == 76 ms replayed at batch 128), so nothing is lost
            if not self.use_graphs or threading.current_thread() is not threading.main_thread():
                fn()
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            L.debug_sync(f'pre-capture {key[:2]}')
            # every graph owns its private memory pool: a pool handle shared by the engine's graphs tripped the caching allocator
            # (use_count assert at capture_begin / freed-pool memory replayed) once graphs of the same engine had been dropped
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # may run inside an autograd worker thread
                fn()
            self._graphs[key] = st = g
            L.debug_sync(f'capture of flow graph {key[:2]}')
        st.replay()
        L.debug_sync(f'replay of flow graph {key[:2]}')

    def _repack_blocks(self):
        """device part of the iResBlock weight (re)pack: every result is written in place, so launch lists and graphs stay valid"""
        dev = self.dev
        for i, (s, b, m) in enumerate(self.blocks):
            cv1, cv2, cv3 = m.convs()
            W = self.w[(s, b)]
            w1, sc1 = self._lop(cv1.weight.detach().to(dev, torch.float32))
            w2, sc2 = self._lop(cv2.weight.detach().to(dev, torch.float32))
            w3, sc3 = self._lop(cv3.weight.detach().to(dev, torch.float32))
            cv1.scale.copy_(sc1); cv2.scale.copy_(sc2); cv3.scale.copy_(sc3)      # lipschitz.py:353-354
            co, ci = w1.shape[:2]                                         # idim, c
            # packed index t * c + ch with t = ky * 3 + kx (what im2col / col2im produce)
            W['w1'][:, :9 * ci].copy_(self._round(w1.permute(0, 2, 3, 1).reshape(co, 9 * ci)))          # [o][t*c+ch]
            W['b1'].copy_(cv1.bias.detach())
            w2m = w2.reshape(w2.shape[0], w2.shape[1])
            W['w2'].copy_(self._round(w2m))
            W['w3'][:9 * ci].copy_(self._round(w3.permute(2, 3, 0, 1).reshape(9 * ci, co)))             # [t*c+ch][k]
            W['b3'].copy_(cv3.bias.detach())
            # VJP chain: conv3^T = im2col(flip) . w3v^T ; conv2^T ; conv1^T = col2im(flip)(. w1v^T)
            W['w3v'][:, :9 * ci].copy_(self._round(w3.permute(1, 2, 3, 0).reshape(co, 9 * ci)))         # [k][t*c+ch]
            W['w2d'].copy_(self._round(w2m.t()))
            W['w1v'][:9 * ci].copy_(self._round(w1.permute(2, 3, 1, 0).reshape(9 * ci, co)))            # [t*c+ch][k]
            # conditioning: conv1x1(u + (A h + a)) + b2 = conv1x1(u) + (W2 A) h + (W2 a + b2)   (lipschitz.py:431-435)
            A = cv2.h_net.net.weight.detach().to(dev, torch.float32)
            a = cv2.h_net.net.bias.detach().to(dev, torch.float32)
            # use the operand-rounded W2 so the folded bias matches what the tensor cores apply to u
            if i not in self.w2f:
                if not hasattr(self, 'w2f_all'):
                    self.w2f_all = torch.empty((len(self.blocks), self.idim, self.idim), device=dev)
                self.w2f[i] = self.w2f_all[i]
            self.w2f[i].copy_(W['w2'])      # fp32 copy of the operand-rounded W2: the conditioning path's backward uses it
            w2r = self.w2f[i]
            self.cond_w[i * self.idim:(i + 1) * self.idim].copy_(w2r @ A)
            self.cond_b[i * self.idim:(i + 1) * self.idim].copy_(w2r @ a + cv2.bias.detach().to(dev, torch.float32))

    def _repack_head(self):
        """prior flow + posterior encoder / fc packs (what the encoder legs need)"""
        self._pack_prior_device()
        if hasattr(self, 'enc'):
            for job in self.enc['jobs']:
                job()

    def load_weights(self, blocks=True, head=True):
        """(re)pack after a parameter update.  The iResBlock packs (Lipschitz normalisation of 96 convolutions: ~1300 small launches
        inside one graph, ~5 ms) are only refreshed for an engine whose block passes are about to run: the engine that only serves
        the posterior-encoder legs of a mixed-precision training step (`_enc_engine`) skips them."""
        with torch.no_grad():
            ver = self.version()
            # parameters re-homed (losses.FusedAdamW moves them into flat storage) or replaced since the packs last ran: the repack
            # graphs start over — one eager pass over the new storages, captured on the following use
            sig = hash(tuple(p.data_ptr() for p in self.core.parameters()))
            if sig != getattr(self, '_param_sig', None):
                self._param_sig = sig
                for key in [k for k in self._graphs if k[0] in ('repack_head', 'repack_blocks')]:
                    del self._graphs[key]
            # lamb (the Poisson rate of the series length) is never trained (iresblock.py:40; no loss reaches it, the optimiser skips
            # it): ONE batched read-back per state-dict load, not a host sync per optimiser step
            vers = getattr(self.core, '_lamb_epoch', 0)
            if vers != self._lamb_vers:
                vals = torch.stack([m.lamb.detach().reshape(()) for (_, _, m) in self.blocks]).cpu().tolist()
                self.lamb = [float(v) for v in vals]
                self._lamb_vers = vers
            if head and self._version != ver:
                self._graphed(('repack_head', hasattr(self, 'enc')), self._repack_head)
                self._pack_prior_host()
                self._version = ver
            if blocks and self._version_blk != ver:
                self._graphed(('repack_blocks',), self._repack_blocks)
                self._version_blk = ver

    def _prior_layout(self):
        """(getter list, per-step record of offsets) of the flat prior-parameter buffer; the offsets never change"""
        steps = self.core.discriminator.prior.flow.steps
        getters, off = [], 0

        def put(fn, n):
            nonlocal off
            o = off
            getters.append((o, n, fn))
            off += n
            return o

        def plain(t):
            return put(lambda t=t: t.detach().reshape(-1), t.numel())

        def coupling(cp):
            n = cp.net
            lin = n.fc3.linear

            def w3(lin=lin):
                v, g = lin.weight_v.detach(), lin.weight_g.detach()
                return (g * v / v.norm(dim=1, keepdim=True)).reshape(-1)
            return [plain(n.fc1.weight), plain(n.fc1.bias), plain(n.fc2.weight), plain(n.fc2.bias), put(w3, lin.weight_v.numel()),
                    plain(lin.bias)]

        rec = []
        for st in steps:
            rec.append(dict(an=[plain(st.actnorm.log_scale), plain(st.actnorm.bias)], W=plain(st.linear.weight), Winv=plain(st.linear.weight_inv),
                            c1u=coupling(st.unit.coupling1_up), c1d=coupling(st.unit.coupling1_dn),
                            uan=[plain(st.unit.actnorm.log_scale), plain(st.unit.actnorm.bias)],
                            c2u=coupling(st.unit.coupling2_up), c2d=coupling(st.unit.coupling2_dn)))
        return getters, rec, off

    def _pack_prior_device(self):
        dev = self.dev
        if self.prior_params is None:
            self._prior_getters, self._prior_rec, total = self._prior_layout()
            self.prior_params = torch.empty((total,), device=dev)
        for o, n, fn in self._prior_getters:
            self.prior_params[o:o + n].copy_(fn().to(dev, torch.float32))

    def _pack_prior_host(self):
        """log|det| of the invertible linears (a per-call constant of the prior-flow kernel) and W^-T (the gradient of that
        log-det, permutation.py:107-112): 64x64 factorizations in fp64 on the host, one read-back for all of them"""
        dev = self.dev
        steps = self.core.discriminator.prior.flow.steps
        rec = self._prior_rec
        mats = torch.stack([torch.stack([st.linear.weight.detach(), st.linear.weight_inv.detach()]) for st in steps]).double().cpu()
        ld = torch.linalg.slogdet(mats)[1]
        winvT = torch.linalg.inv(mats[:, 0]).transpose(1, 2).contiguous().float()
        if not self.prior_winvT:
            self.prior_winvT = [torch.empty((mats.shape[-1], mats.shape[-1]), device=dev) for _ in steps]
        for j in range(len(steps)):
            self.prior_winvT[j].copy_(winvT[j])
            rec[j]['ld_w'], rec[j]['ld_winv'] = float(ld[j, 0]), float(ld[j, 1])

        def op(kind, backward, offs, skip=0, up=0):
            o = L.FlowOp()
            o.kind, o.backward, o.split_skip, o.up = kind, backward, skip, up
            for i, v in enumerate(offs):
                o.off[i] = v
            return o

        # backward program = FlowPrior.sample (priors/flow.py:226-230 -> PriorFlow.backward :181-189)
        bw, ld_b = [], 0.0
        for r in reversed(rec):
            bw += [op(2, 1, r['c2d'], 1, 0), op(2, 1, r['c2u'], 1, 1), op(0, 1, r['uan']), op(2, 1, r['c1d'], 0, 0), op(2, 1, r['c1u'], 0, 1),
                   op(1, 1, [r['Winv']]), op(0, 1, r['an'])]
            ld_b += r['ld_winv']
        fw, ld_f = [], 0.0
        for r in rec:
            fw += [op(0, 0, r['an']), op(1, 0, [r['W']]), op(2, 0, r['c1u'], 0, 1), op(2, 0, r['c1d'], 0, 0), op(0, 0, r['uan']),
                   op(2, 0, r['c2u'], 1, 1), op(2, 0, r['c2d'], 1, 0)]
            ld_f += r['ld_w']
        for name, prog, ldc in (('backward', bw, ld_b), ('forward', fw, ld_f)):
            if name in self.prior_ops:                       # the op program (offsets) is fixed; only the constant moves
                self.prior_ops[name] = (self.prior_ops[name][0], len(prog), ldc)
            else:
                arr = (L.FlowOp * len(prog))(*prog)
                buf = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
                self.prior_ops[name] = (buf, len(prog), ldc)

    # ---- building blocks
    def prior_flow(self, z, direction, want_logdet=False, kl_base=None):
        """kl_base = log q(h|x): the second result becomes KL = log q - log p(h) (FlowPrior.calcKL) instead of the log-det"""
        buf, n, ld = self.prior_ops[direction]
        out = torch.empty_like(z)
        logdet = torch.empty((z.shape[0],), device=self.dev) if (want_logdet or kl_base is not None) else None
        L.call('indm_prior_flow', L.ptr(z), L.ptr(out), L.ptr(logdet), L.ptr(self.prior_params), L.ptr(buf), n, ctypes.c_float(ld),
               L.ptr(kl_base), z.shape[0])
        return (out, logdet) if logdet is not None else out

    def _cond_table(self, h):
        nblk = len(self.blocks)
        L.call('indm_linear_f32', L.ptr(h), L.ptr(self.cond_w), L.ptr(self.cond_b), L.ptr(self.cond), self.N, self.core.latent_dim,
               nblk * self.idim, 0, 0, L.DTYPE_F32)

    # ---- prebuilt launch lists.  The flow passes issue ~1000 small launches per call; building a ctypes descriptor per launch
    # from Python (~100 us each) cost more than the GPU work, so every (block, buffers) combination is built once and replayed.
    def _mk_call(self, name, *args):
        fn = getattr(L.lib(), name)
        cargs = [ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a for a in args]
        keep = [a for a in args if isinstance(a, torch.Tensor)]

        def run(fn=fn, cargs=cargs, keep=keep, name=name):
            L.check(fn(*cargs, L._stream()), name)
        return run

    def _mk_igemm(self, **kw):
        d = L.IgemmDesc()
        d.scale = 1.0
        d.res_scale = 1.0
        keep = []
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                keep.append(v)
                v = v.data_ptr()
            setattr(d, k, v)
        lib = L.lib()

        def run(d=d, lib=lib, keep=keep):
            L.check(lib.indm_igemm(ctypes.byref(d), L._stream()), 'igemm')
        return run

    def _replay(self, key, builder):
        ops = self._ops.get(key)
        if ops is None:
            ops = builder()
            self._ops[key] = ops
        for op in ops:
            op()

    def _static(self, name, like):
        """engine-owned buffer with a stable address (launch lists are keyed by data pointers)"""
        t = self._bufs.get(name)
        if t is None or t.shape != like.shape:
            t = torch.empty_like(like)
            self._bufs[name] = t
        return t

    def _g(self, i, s, m, x_nchw, out, residual, scale):
        """out = scale * g(x; h) + residual, all NCHW fp32 [N, c, H, W] at scale s; block index i selects the cond bias."""
        self._g_impl(i, s, m, x_nchw, out, residual, scale, None, None)

    def _g_impl(self, i, s, m, x_nchw, out, residual, scale, d1, d2):
        key = ('g', i, x_nchw.data_ptr(), out.data_ptr(), residual.data_ptr() if residual is not None else 0, float(scale),
               d1.data_ptr() if d1 is not None else 0)

        def build():
            N, idim = self.N, self.idim
            c = m.channels
            _, h0, w0 = self.core.input_shape
            H, Wd = h0 >> s, w0 >> s
            W = self.w[(self.blocks[i][0], self.blocks[i][1])]
            a0, o9 = self.a0[s], self.o9[s]
            n_el = N * H * Wd * idim
            u1 = self.u1.view(-1)[:n_el].view(N, H, Wd, idim)
            u2 = self.u2.view(-1)[:n_el].view(N, H, Wd, idim)
            ob = (lambda t: dict(out_bf16=t)) if self.mode == 'bf16' else (lambda t: dict(out_f32=t))
            return [
                self._mk_call('indm_im2col3x3_nchw', x_nchw, a0, ctypes.c_int64(N), c, H, Wd, self.kp[s], 0, 0 if m.first else 1, self.dt),
                self._mk_igemm(dtype=self.dt, a=a0, N=N, H=H, W=Wd, Cin=self.kp[s], b=W['w1'], Cout=idim, taps=1, bias=W['b1'], act=1,
                               out_ld=idim, aux_cos=d1, **ob(u1)),
                self._mk_igemm(dtype=self.dt, a=u1, N=N, H=H, W=Wd, Cin=idim, b=W['w2'], Cout=idim, taps=1, rowbias=self.cond[:, i * idim:],
                               rowbias_ld=self.cond.shape[1], act=1, out_ld=idim, aux_cos=d2, **ob(u2)),
                self._mk_igemm(dtype=self.dt, a=u2, N=N, H=H, W=Wd, Cin=idim, b=W['w3'], Cout=self.ld9[s], taps=1, out_f32=o9, out_ld=self.ld9[s]),
                self._mk_call('indm_col2im3x3_nchw', o9, ctypes.c_int64(self.ld9[s]), W['b3'], residual, None, ctypes.c_float(scale), out,
                              ctypes.c_int64(N), c, H, Wd, 0),
            ]
        self._replay(key, build)

    def _ensure(self, blocks=True, head=True):
        """head = prior flow + posterior encoder packs (one host read-back per refresh for the 64 x 64 log-dets): the block passes of
        a training step (forward_logdet / forward_map / the backward) never touch them"""
        ver = self.version()
        if (head and self._version != ver) or (blocks and self._version_blk != ver):
            self.load_weights(blocks=blocks, head=head)

    # ---- public passes
    def reverse(self, z, eps=None, h=None, seed=0, offset=0, atol=1e-5, rtol=1e-5, max_iter=1000):
        """WolfCore.forward(reverse=True) (wolf.py:82-89) on the flow's own input layout (flow_forward has already applied
        SqueezeLayer when flow.squeeze): h ~ prior, then ResidualFlow.bwdpass(z, h).  Returns x with z's shape."""
        self._ensure()
        cfg = self.core.config
        N = self.N
        z = z.float().contiguous()
        shape = z.shape
        if h is None:
            if eps is None:
                L.call('indm_randn_f32', L.ptr(self.eps), self.eps.numel(), seed, 0x7F100000 + offset)
                eps = self.eps
            h = self.prior_flow(eps.float().contiguous(), 'backward')
        self.h.copy_(h)
        self._cond_table(self.h)
        nb = self.nb
        x = z
        if len(nb) > 1:   # resflow_.py:331-333
            x = x.view(x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2, 2, 2).permute(0, 1, 5, 2, 3, 4) \
                 .reshape(x.shape[0], x.shape[1], x.shape[2], 2, x.shape[3] // 2).permute(0, 1, 3, 2, 4).reshape(x.shape)
        c0, h0, w0 = self.core.input_shape
        k = len(nb) - 1
        x = x.reshape(N, c0 * 4 ** k, h0 >> k, w0 >> k).contiguous()
        self.iterations = []
        idx_of = {(s, b): i for i, (s, b, _) in enumerate(self.blocks)}
        for s in reversed(range(len(nb))):
            if s < len(nb) - 1:
                x = torch.nn.functional.pixel_shuffle(x, 2).contiguous()      # SqueezeLayer.inverse: a pure permutation
            blks = [(b, m) for (ss, b, m) in self.blocks if ss == s]
            for b, m in reversed(blks):
                i = idx_of[(s, b)]
                y = self._static(f'inv_y{s}', x)
                y.copy_(x)
                ping = [self._static(f'inv_a{s}', x), self._static(f'inv_b{s}', x)]
                cur, nxt = y, ping[0]
                it = 0
                while True:
                    self._g(i, s, m, cur, nxt, y, -1.0)                       # nxt = y - g(cur)
                    L.call('indm_fixed_point_check', L.ptr(nxt), L.ptr(cur), L.ptr(y), y.numel(), ctypes.c_float(atol), ctypes.c_float(rtol),
                           L.ptr(self.flag))
                    conv = float(self.flag.item()) < 1.0
                    cur, nxt = nxt, (ping[1] if cur is y else cur)
                    if conv:
                        break
                    it += 1
                    if it > max_iter:
                        break
                self.iterations.append(it)
                x = cur.clone()
        return x.reshape(shape)

    # ---- posterior q(h|x): BN-ResNet encoder -> weight-normed linear -> reparameterisation -> KL against the flow prior
    def _enc_engine(self):
        """The engine that runs the posterior encoder / fc / KL legs.  The encoder is < 2 % of the flow's work but its output decides
        h and the KL term (20-35 nats against a log-det of ~0.03 here): BF16 operands there put 4e-3 into h and 0.03-0.1 nats into
        the KL — the whole error of the training-mode (log-det - KL).  POLICY['flow']['encoder'] (default 'tf32') therefore runs
        these legs on the compensated-TF32 engine of the same batch even when the iResBlocks run in BF16."""
        want = precision.POLICY['flow'].get('encoder', self.mode) if self.core.compute_mode in (None, 'auto') else self.mode
        return self if want == self.mode else self.core.engine(self.N, mode=want)

    def _build_encoder(self):
        core, N, dev = self.core, self.N, self.dev
        enc = core.config.flow.wolf_params['discriminator']['encoder']
        kc = self.kchunk
        cp = lambda c: ((c + kc - 1) // kc) * kc
        c0, S, _ = core.input_shape
        E = dict(ops=[], jobs=[])
        self.enc = E
        E['x_in'] = torch.zeros((N, S, S, cp(c0)), device=dev, dtype=self.tdtype)
        net = core.discriminator.encoder.net

        def fold(conv, bn, cin_pad, cout):
            """conv weight with the eval-mode BatchNorm folded in: W * s[o], bias = beta - mean * s (nn.BatchNorm2d, eps 1e-5)"""
            k = conv.weight.shape[-1]
            w = torch.zeros((k * k, cout, cin_pad), device=dev, dtype=self.tdtype)
            b = torch.zeros((cout,), device=dev)

            def job():
                W = conv.weight.detach().to(dev, torch.float32)
                if bn is not None:
                    sc = bn.weight.detach().to(dev, torch.float32) / torch.sqrt(bn.running_var.detach().to(dev, torch.float32) + 1e-5)
                    W = W * sc[:, None, None, None]
                    b.copy_(bn.bias.detach().to(dev, torch.float32) - bn.running_mean.detach().to(dev, torch.float32) * sc)
                else:
                    b.copy_(conv.bias.detach().to(dev, torch.float32))
                co, ci = W.shape[:2]
                w.zero_()
                w[:, :, :ci].copy_(self._round(W.permute(2, 3, 0, 1).reshape(k * k, co, ci)))
            E['jobs'].append(job)
            return w, b

        def conv(a, Hin, cin, w, b, cout, stride, taps, act, residual=None, want_f32=False, nchw=False):
            Ho = Hin // stride
            kw = dict(dtype=self.dt, a=a, N=N, H=Ho, W=Ho, Cin=cp(cin), b=w, Cout=cout, taps=taps, bias=b, act=act)
            if stride == 2:
                kw.update(stride=2, a_H=Hin, a_W=Hin, pad=1 if taps == 9 else 0)
            if residual is not None:
                kw.update(residual=residual, res_ld=cp(cout), res_scale=1.0)
            out_op = out_f = None
            if nchw:
                out_f = torch.zeros((N, cout, Ho, Ho), device=dev)
                kw.update(out_mode=1, out_f32=out_f)
            else:
                out_op = torch.zeros((N, Ho, Ho, cp(cout)), device=dev, dtype=self.tdtype)
                kw.update(out_ld=cp(cout))
                if self.mode == 'bf16':
                    kw['out_bf16'] = out_op
                    if want_f32:
                        out_f = torch.zeros((N, Ho, Ho, cp(cout)), device=dev)
                        kw['out_f32'] = out_f
                else:
                    kw['out_f32'] = out_op
                    out_f = out_op
            E['ops'].append(kw)
            return out_op, out_f, Ho

        x, xf, H, inp = E['x_in'], None, S, c0
        for lv, hid in enumerate(enc['hidden_planes']):
            res = getattr(net, f'resnet{lv}')
            for m, stride in enumerate((1, 2)):
                blk = res.main[m]
                ci = inp if m == 0 else hid
                w1, b1 = fold(blk.conv1, blk.bn1, cp(ci), hid)
                t1, _, Ho = conv(x, H, ci, w1, b1, hid, stride, 9, 2)
                if hasattr(blk, 'downsample'):
                    wd, bd = fold(blk.downsample[0], blk.downsample[1], cp(ci), hid)
                    _, r, _ = conv(x, H, ci, wd, bd, hid, stride, 1, 0, want_f32=True)
                else:
                    r = xf                         # identity shortcut: the fp32 copy of the block input
                w2, b2 = fold(blk.conv2, blk.bn2, cp(hid), hid)
                x, xf, _ = conv(t1, Ho, hid, w2, b2, hid, 1, 9, 2, residual=r, want_f32=True)
                H = Ho
            inp = hid
        wt, bt = fold(net.top, None, cp(inp), enc['out_planes'])
        _, top, _ = conv(x, H, inp, wt, bt, enc['out_planes'], 1, 1, 2, nchw=True)
        E['top'] = top                                                   # [N, out_planes, h, w] == flattened [N, in_dim]
        d = core.latent_dim
        E['fc_w'] = torch.empty((2 * d, top[0].numel()), device=dev)
        E['fc_b'] = torch.empty((2 * d,), device=dev)
        fc = core.discriminator.fc.linear

        def job_fc():
            v, g = fc.weight_v.detach().to(dev, torch.float32), fc.weight_g.detach().to(dev, torch.float32)
            E['fc_w'].copy_(g * v / v.norm(dim=1, keepdim=True))           # legacy weight_norm, dim 0 (nnet/weight_norm.py:8-40)
            E['fc_b'].copy_(fc.bias.detach().to(dev, torch.float32))
        E['jobs'].append(job_fc)
        E['c'] = torch.empty((N, 2 * d), device=dev)
        E['logq'] = torch.empty((N,), device=dev)
        E['c0'] = c0

    def posterior(self, x, eps=None, seed=0, offset=0):
        """GaussianDiscriminator.sampling_and_KL (gaussian.py:67-76) with nsamples = 1: returns (h [N,64], KL [N])."""
        enc = self._enc_engine()
        if enc is not self:
            return enc.posterior(x, eps=eps, seed=seed, offset=offset)
        self._ensure(blocks=False)
        if not hasattr(self, 'enc'):          # built on first use: sampling-only callers never need the posterior encoder
            self._build_encoder()
            with torch.no_grad():
                for job in self.enc['jobs']:
                    job()
        N, E = self.N, self.enc
        c0, S, _ = self.core.input_shape
        x = x.float().contiguous()
        L.call('indm_prep_input', L.ptr(x), L.ptr(E['x_in']), N, c0, S, S, E['x_in'].shape[-1], ctypes.c_float(1.0), ctypes.c_float(0.0), 0, self.dt)
        for kw in E['ops']:
            L.igemm(**kw)
        top = E['top']
        L.call('indm_linear_f32', L.ptr(top), L.ptr(E['fc_w']), L.ptr(E['fc_b']), L.ptr(E['c']), N, top[0].numel(), E['c'].shape[1], 0, 0, L.DTYPE_F32)
        if eps is None:
            L.call('indm_randn_f32', L.ptr(self.eps), self.eps.numel(), seed, 0x7F200000 + offset)
            eps = self.eps
        h = torch.empty((N, self.core.latent_dim), device=self.dev)
        L.call('indm_posterior_sample', L.ptr(E['c']), L.ptr(eps.float().contiguous()), L.ptr(h), L.ptr(E['logq']), N)
        _, kl = self.prior_flow(h, 'forward', kl_base=E['logq'])
        return h, kl

    # ---- joint training: forward with everything the explicit backward needs, and that backward
    train_token = 0

    def train_forward(self, x, eps=None, vareps=None, n_terms=None, seed=0, offset=0):
        """WolfCore.forward(reverse=False) in training mode (wolf.py:90-128): returns (z, logdet - KL)."""
        x = x.detach().float().contiguous()
        h, kl = self.train_posterior(x, eps=eps, seed=seed, offset=offset)
        _, c, eps_s, enc_out = self._train_saved
        z, logpx = self.forward_logdet(x, h, vareps=vareps, n_terms=n_terms, training=True, seed=seed, offset=offset, save=True)
        FlowEngine.train_token += 1
        self.train_token = FlowEngine.train_token
        return z, -logpx - kl

    def train_posterior(self, x, eps=None, seed=0, offset=0):
        """GaussianDiscriminator.sampling_and_KL in training mode: batch-statistics encoder -> fc -> h ~ q(h|x), KL"""
        from .wolf_encoder_train import EncoderTrain
        enc = self._enc_engine()
        if enc is not self:
            h, kl = enc.train_posterior(x, eps=eps, seed=seed, offset=offset)
            self._train_saved = enc._train_saved
            return h, kl
        self._ensure(blocks=False)
        if not hasattr(self, 'enc'):
            self._build_encoder()
            with torch.no_grad():
                for job in self.enc['jobs']:
                    job()
        if not hasattr(self, 'enc_train'):
            self.enc_train = EncoderTrain(self)
        N, E = self.N, self.enc
        xs = self._static('tr_x', x)
        xs.copy_(x)
        if eps is None:
            L.call('indm_randn_f32', L.ptr(self.eps), self.eps.numel(), seed, 0x7F200000 + offset)
            eps = self.eps
        eps_s = self._static('tr_eps', self.eps)
        eps_s.copy_(eps)
        c = self._static('tr_c', E['c'])
        h = self._static('tr_h', self.h)
        self.enc_train._ensure()
        out = {}

        def body():
            enc_out = self.enc_train.forward(xs)
            out['enc_out'] = enc_out
            L.call('indm_linear_f32', L.ptr(enc_out), L.ptr(E['fc_w']), L.ptr(E['fc_b']), L.ptr(c), N, enc_out.shape[1], c.shape[1], 0, 0, L.DTYPE_F32)
            L.call('indm_posterior_sample', L.ptr(c), L.ptr(eps_s), L.ptr(h), L.ptr(E['logq']), N)
        self._graphed(('posterior',), body)
        enc_out = out.get('enc_out', self.enc_train.top.view(N, -1))
        _, kl = self.prior_flow(h, 'forward', kl_base=E['logq'])      # carries log|det W| as a host constant: stays out of the graph
        self._train_saved = (h, c, eps_s, enc_out)
        return h, kl

    def _bwd_key(self):
        anchor = next(p for p in self.core.parameters() if p.requires_grad)
        return ('bwd', anchor.grad.data_ptr() if anchor.grad is not None else 0, tuple(x[0] for x in self._saved))

    def _bwd_body(self):
        from .wolf_backward import FlowBackward, PosteriorBackward
        enc = self._enc_engine()
        if not hasattr(self, '_fbw'):
            self._fbw, self._pbw = FlowBackward(self), PosteriorBackward(enc)
        h, c, eps, enc_out = self._train_saved
        gz_s, gl_s = self._bufs['bw_gz'], self._bufs['bw_gl']
        _, gh_blocks = self._fbw.run(gz_s, gl_s)
        g_enc = self._pbw.run(h, gh_blocks, -gl_s, c, eps, enc_out)
        enc.enc_train.backward(g_enc)

    def train_backward(self, gz, gloss):
        """gz = d L / d z, gloss [N] = d L / d (logdet - KL): runs the residual-flow, KL / posterior-head and encoder backward plans.
        The whole flow backward is a fixed launch sequence (~1600 launches) over engine-owned buffers: one CUDA graph, keyed by
        where the gradients live.  autograd calls this from its worker thread, where stream capture gets invalidated — the graph
        is therefore captured ahead of time on the main thread (`precapture_backward`, called at the end of the training forward once
        the eager first run has built the launch lists) and only REPLAYED here."""
        gz_s, gl_s = self._static('bw_gz', gz), self._static('bw_gl', gloss)
        gz_s.copy_(gz)
        gl_s.copy_(gloss)
        self._graphed(self._bwd_key(), self._bwd_body)

    def precapture_backward(self):
        """main thread, after train_forward: capture (without running) the backward graph for the buffers of this forward, if the
        plan has run eagerly once and no graph exists yet"""
        if not self.use_graphs or threading.current_thread() is not threading.main_thread() or torch.cuda.is_current_stream_capturing():
            return
        if 'bw_gz' not in self._bufs or not self._saved:
            return
        key = self._bwd_key()
        if self._graphs.get(key) != 1:
            return
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._bwd_body()
        self._graphs[key] = g

    # ---- power-series log-det (iresblock.py:90-174): VJP chain of g on the tensor cores
    def _dviews(self, s, m, x_nchw):
        """views of the cos-factor buffers (d0: input Sin', None for the flow's first block; d1, d2: the two hidden Sin') for a block
        at scale s"""
        N, idim = self.N, self.idim
        _, h0, w0 = self.core.input_shape
        H, Wd = h0 >> s, w0 >> s
        n_el = N * H * Wd * idim
        d1, d2 = self.d1.view(-1)[:n_el].view(N, H, Wd, idim), self.d2.view(-1)[:n_el].view(N, H, Wd, idim)
        d0 = None if m.first else self.d0.view(-1)[:x_nchw.numel()].view(x_nchw.shape)
        return d0, d1, d2

    def _g_store(self, i, s, m, x_nchw, out):
        """out = x + g(x; h) (NCHW fp32) keeping cos(2 pi .) of the three Sin pre-activations for the VJP chain"""
        d0, d1, d2 = self._dviews(s, m, x_nchw)
        if not m.first:
            L.call('indm_cos2pi_f32', L.ptr(x_nchw), L.ptr(d0), x_nchw.numel())
        self._g_impl(i, s, m, x_nchw, out, x_nchw, 1.0, d1, d2)
        return d0, d1, d2

    def _g_vjp(self, i, s, m, v, out, d0, d1, d2):
        """out = J_g(x)^T v for the block whose cos factors are (d0, d1, d2); v, out NCHW fp32"""
        key = ('vjp', i, v.data_ptr(), out.data_ptr(), d0.data_ptr() if d0 is not None else 0, d1.data_ptr(), d2.data_ptr())

        def build():
            N, idim = self.N, self.idim
            c = m.channels
            _, h0, w0 = self.core.input_shape
            H, Wd = h0 >> s, w0 >> s
            W = self.w[(self.blocks[i][0], self.blocks[i][1])]
            a0, o9 = self.a0[s], self.o9[s]
            n_el = N * H * Wd * idim
            t1, t2 = self.u1.view(-1)[:n_el].view(N, H, Wd, idim), self.u2.view(-1)[:n_el].view(N, H, Wd, idim)
            ob = (lambda t: dict(out_bf16=t)) if self.mode == 'bf16' else (lambda t: dict(out_f32=t))
            return [
                self._mk_call('indm_im2col3x3_nchw', v, a0, ctypes.c_int64(N), c, H, Wd, self.kp[s], 1, 0, self.dt),
                self._mk_igemm(dtype=self.dt, a=a0, N=N, H=H, W=Wd, Cin=self.kp[s], b=W['w3v'], Cout=idim, taps=1, out_ld=idim, mul=d2,
                               mul_ld=idim, **ob(t2)),
                self._mk_igemm(dtype=self.dt, a=t2, N=N, H=H, W=Wd, Cin=idim, b=W['w2d'], Cout=idim, taps=1, out_ld=idim, mul=d1,
                               mul_ld=idim, **ob(t1)),
                self._mk_igemm(dtype=self.dt, a=t1, N=N, H=H, W=Wd, Cin=idim, b=W['w1v'], Cout=self.ld9[s], taps=1, out_f32=o9, out_ld=self.ld9[s]),
                self._mk_call('indm_col2im3x3_nchw', o9, ctypes.c_int64(self.ld9[s]), None, None, d0, ctypes.c_float(1.0), out,
                              ctypes.c_int64(N), c, H, Wd, 1),
            ]
        self._replay(key, build)

    def _series_rng(self):
        """Generator of the Poisson series lengths (iresblock.py:306 draws them from numpy's global generator).  The reference is ONE
        process (nn.DataParallel): every replica evaluates the same n for a block.  With one process per GPU the ranks must agree
        too — otherwise each step costs the slowest rank's draw (the series length decides 5 - 15 % of a block's work) and the ranks
        estimate different truncations of one mini-batch's log-det: rank 0 draws a seed from numpy's global generator once and
        broadcasts it; all ranks then draw from a private RandomState(seed), in lock step.  Single process: numpy's global generator,
        exactly like the reference."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return np.random
        rng = getattr(self.core, '_shared_series_rng', None)
        if rng is None:
            seed = torch.tensor([int(np.random.randint(0, 2 ** 31 - 1))], dtype=torch.int64, device=self.dev)
            dist.broadcast(seed, src=0)
            rng = self.core._shared_series_rng = np.random.RandomState(int(seed.item()))
        return rng

    def forward_logdet(self, x, h, vareps=None, n_terms=None, training=False, seed=0, offset=0, save=False):
        """ResidualFlow.fwdpass(x, h, eval_logdet=True) (resflow_.py:310-324): returns (z, logpx [N]) with
        logpx = -sum over blocks of the power-series log-det estimate: basic estimator with 20 exact terms in eval mode
        (iresblock.py:127-132,253-261), Neumann estimator value with 2 exact terms in training mode (:114-121,264-273)."""
        self._ensure(head=False)
        N = self.N
        x = x.float().contiguous()
        shape = x.shape
        self.h.copy_(h)
        self._cond_table(self.h)
        if not hasattr(self, 'd1'):
            self.d0 = torch.empty((x.numel(),), device=self.dev)
            self.d1, self.d2 = torch.empty_like(self.u1), torch.empty_like(self.u2)
        logpx = self._static('logpx', torch.empty((N,), device=self.dev))
        logpx.zero_()
        nb = self.nb
        bi = 0
        self.vjp_count = 0
        self.vjp_per_block = []       # VJP chains evaluated per block in this pass (bench.py counts algorithmic FLOPs from it)
        if save and not training:
            raise RuntimeError('the flow backward belongs to the training-mode (Neumann) estimator')
        self._saved = [] if save else None
        for s in range(len(nb)):
            # engine-owned buffers with stable addresses: the prebuilt launch lists are keyed by data pointers
            xs = [self._static(f'fx_a{s}', x), self._static(f'fx_b{s}', x)]
            ve = self._static(f've{s}', x)
            neumann = self._static(f'neu{s}', x)
            bufs = [self._static(f'vj_a{s}', x), self._static(f'vj_b{s}', x)]
            xs[0].copy_(x)
            cur_x = 0
            D = x[0].numel()
            for i, (ss, b, m) in enumerate(self.blocks):
                if ss != s:
                    continue
                n = int(n_terms[bi]) if n_terms is not None else int(self._series_rng().poisson(self.lamb[i], 1)[0])     # iresblock.py:306
                K, coef = series_coefficients(n, training, self.lamb[i])
                if vareps is not None:
                    ve.copy_(vareps[bi])
                else:
                    L.call('indm_randn_f32', L.ptr(ve), ve.numel(), seed, 0x7F300000 + (offset << 8) + bi)
                xin, xout = xs[cur_x], xs[1 - cur_x]

                if training:
                    # Neumann estimator (value): w = eps + sum_k (-1)^k c_k J^k^T eps ; logdet = <J^T w, eps>.  THREE graphs per block
                    # — start (branch + cos factors, w = cur = eps), one series term (cur <- J^T cur ; w += coef[k] cur ; k += 1; the
                    # coefficient comes from a device table indexed by a device counter) and finish — replayed 1 + K + 1 times, so no
                    # series length ever triggers a new capture (one graph per (block, K) meant ~100 ms capture stalls in random
                    # steps for the first few hundred steps of a run).
                    curb = self._static(f'ser_cur{s}', x)
                    tab = self._static('ser_coef', torch.empty((64,), device=self.dev))
                    kctr = self._static('ser_k', torch.zeros((1,), dtype=torch.int32, device=self.dev))
                    if K > tab.numel():
                        raise RuntimeError(f'series of {K} terms exceeds the coefficient table')
                    # pageable source: the runtime stages it before returning, so the host list can be rebuilt for the next block
                    tab[:K].copy_(torch.tensor([((-1) ** k) * coef[k - 1] for k in range(1, K + 1)], dtype=torch.float32))
                    kctr.zero_()
                    sx, sv, sw = (self._static(f'sv_{nm}{i}', xin) for nm in ('x', 'e', 'w')) if save else (None, None, None)

                    dv = self._dviews(s, m, xin)

                    def start(i=i, s=s, m=m, xin=xin, xout=xout):
                        self._g_store(i, s, m, xin, xout)
                        neumann.copy_(ve)
                        curb.copy_(ve)

                    def term(i=i, s=s, m=m, dv=dv):
                        d0, d1, d2 = dv
                        nxt = bufs[0]
                        self._g_vjp(i, s, m, curb, nxt, d0, d1, d2)
                        L.call('indm_series_step_f32', L.ptr(neumann), L.ptr(curb), L.ptr(nxt), L.ptr(tab), L.ptr(kctr), neumann.numel())
                        L.call('indm_advance_step', L.ptr(kctr))

                    def finish(i=i, s=s, m=m, xin=xin, dv=dv):
                        d0, d1, d2 = dv
                        nxt = bufs[1]
                        self._g_vjp(i, s, m, neumann, nxt, d0, d1, d2)
                        L.call('indm_rowdot_f32', L.ptr(nxt), L.ptr(ve), L.ptr(logpx), N, D, ctypes.c_float(-1.0), 1)
                        if save:
                            # what the block's backward needs: its input, the probe and the (constant) Neumann vector
                            sx.copy_(xin); sv.copy_(ve); sw.copy_(neumann)
                    self._graphed(('blk_start', i, xin.data_ptr(), xout.data_ptr()), start)
                    for _k in range(K):
                        self._graphed(('blk_term', i, xin.data_ptr()), term)
                    self._graphed(('blk_finish', i, xin.data_ptr(), bool(save)), finish)
                else:
                    def body(i=i, s=s, m=m, xin=xin, xout=xout, K=K, coef=coef):
                        d0, d1, d2 = self._g_store(i, s, m, xin, xout)
                        cur = ve
                        # basic estimator: sum_k (-1)^(k+1)/k c_k <J^k^T eps, eps>
                        for k in range(1, K + 1):
                            nxt = bufs[k & 1]
                            self._g_vjp(i, s, m, cur, nxt, d0, d1, d2)
                            L.call('indm_rowdot_f32', L.ptr(nxt), L.ptr(ve), L.ptr(logpx), N, D, ctypes.c_float(-((-1) ** (k + 1)) / k * coef[k - 1]), 1)
                            cur = nxt
                    # one CUDA graph per (block, series length): the chain is ~5 K small launches
                    self._graphed(('blk', i, K, False, False, float(self.lamb[i])), body)
                self.vjp_count += K if not training else K + 1
                self.vjp_per_block.append(K if not training else K + 1)
                if training and save:
                    self._saved.append((i, s, m) + tuple(self._static(f'sv_{nm}{i}', xin) for nm in ('x', 'e', 'w')))
                cur_x = 1 - cur_x
                bi += 1
            x = xs[cur_x]
            if s < len(nb) - 1:
                x = _squeeze2(x).contiguous()
        out = x.reshape(N, -1)
        if len(nb) > 1:
            out = out.view(shape[0], shape[1], 2, 2, shape[2] // 2, shape[3] // 2).permute(0, 1, 4, 2, 5, 3).reshape(shape)
        else:
            out = out.view(shape)
        return out.clone(), logpx.clone()              # the chain lives in engine-owned buffers: hand out copies

    def forward_map(self, x, h):
        """ResidualFlow.fwdpass(x, h, eval_logdet=False) on the flow's own input layout (resflow_.py:310-324)."""
        self._ensure(head=False)
        N = self.N
        x = x.float().contiguous()
        shape = x.shape
        self.h.copy_(h)
        self._cond_table(self.h)
        nb = self.nb
        for s in range(len(nb)):
            xs = [self._static(f'fx_a{s}', x), self._static(f'fx_b{s}', x)]
            xs[0].copy_(x)
            cur_x = 0
            for i, (ss, b, m) in enumerate(self.blocks):
                if ss != s:
                    continue
                self._g(i, s, m, xs[cur_x], xs[1 - cur_x], xs[cur_x], 1.0)
                cur_x = 1 - cur_x
            x = xs[cur_x]
            if s < len(nb) - 1:
                x = _squeeze2(x).contiguous()
        out = x.reshape(N, -1)
        if len(nb) > 1:
            out = out.view(shape[0], shape[1], 2, 2, shape[2] // 2, shape[3] // 2).permute(0, 1, 4, 2, 5, 3).reshape(shape)
        else:
            out = out.view(shape)
        return out.clone()


def _squeeze2(x):
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 4, h // 2, w // 2)
