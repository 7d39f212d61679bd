"""CPU restatement of the wolf / residual-flow side of the INDM hot path (torch-CPU FP32, functional).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned by tests/golden/flow_*.npz (live reference outputs).

Covers: the conditional Residual Flow (flow_models/wolf/flows/resflow/resflow_.py:20-518, layers/iresblock.py:14-179,
layers/base/lipschitz.py:321-441, layers/base/activations.py:7-12, layers/squeeze.py:7-45), the 64-d latent prior flow
(flow_models/wolf/modules/discriminators/priors/flow.py:16-286 with flows/normalization.py:13-112,
flows/permutation.py:75-149, flows/couplings/coupling.py:13-177, transform.py:49-81, blocks.py:11-48),
`WolfCore.forward(reverse=True)` (flow_models/wolf/wolf.py:81-89) and the `flow_forward` wrapper
(flow_models/flow_model.py:53-67).  Parameters: flat dict keyed like the reference state-dict without `module.`.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

COEFF = 0.98   # resflow_.py:40 (coeff) -> LopConv2d soft normalisation


def n_blocks(config):
    return [int(v) for v in config.flow.nblocks.split('-')]


def flow_input_shape(config):
    """generator.py:96-101"""
    c, s = config.data.num_channels, config.data.image_size
    return (c * 4, s // 2, s // 2) if config.flow.squeeze else (c, s, s)


def block_layout(config):
    """[(scale, block, channels, first)] in forward order; nnet indices are 0/2/4 for the first block of scale 0
    (no leading Sin, resflow_.py:442-444) and 1/3/5 elsewhere."""
    c, _, _ = flow_input_shape(config)
    out = []
    for s, nb in enumerate(n_blocks(config)):
        for b in range(nb):
            out.append((s, b, c, s == 0 and b == 0))
        c *= 4
    return out


def param_shapes(config):
    """[(key, shape)] of the full wolf flow state-dict, reference order (probed: 687 entries for CIFAR)."""
    idim = config.flow.intermediate_dim
    out = []
    for s, b, c, first in block_layout(config):
        p = f'generator.flow.transforms.{s}.chain.{b}.'
        out += [(p + 'geom_p', ()), (p + 'lamb', ()), (p + 'last_n_samples', (1,)), (p + 'last_firmom', (1,)), (p + 'last_secmom', (1,))]
        j0 = 0 if first else 1
        out += [(p + f'nnet.{j0}.weight', (idim, c, 3, 3)), (p + f'nnet.{j0}.bias', (idim,)), (p + f'nnet.{j0}.scale', ())]
        out += [(p + f'nnet.{j0 + 2}.weight', (idim, idim, 1, 1)), (p + f'nnet.{j0 + 2}.bias', (idim,)), (p + f'nnet.{j0 + 2}.scale', ()),
                (p + f'nnet.{j0 + 2}.h_net.net.weight', (idim, 64)), (p + f'nnet.{j0 + 2}.h_net.net.bias', (idim,))]
        out += [(p + f'nnet.{j0 + 4}.weight', (c, idim, 3, 3)), (p + f'nnet.{j0 + 4}.bias', (c,)), (p + f'nnet.{j0 + 4}.scale', ())]
    enc = config.flow.wolf_params['discriminator']['encoder']
    inp = enc['in_planes']
    for lv, hid in enumerate(enc['hidden_planes']):
        for m, (ci, stride) in enumerate(((inp, 1), (hid, 2))):
            p = f'discriminator.encoder.net.resnet{lv}.main.{m}.'
            out += [(p + 'conv1.weight', (hid, ci, 3, 3))]
            for bn in ('bn1',):
                out += [(p + f'{bn}.weight', (hid,)), (p + f'{bn}.bias', (hid,)), (p + f'{bn}.running_mean', (hid,)),
                        (p + f'{bn}.running_var', (hid,)), (p + f'{bn}.num_batches_tracked', ())]
            out += [(p + 'conv2.weight', (hid, hid, 3, 3))]
            out += [(p + 'bn2.weight', (hid,)), (p + 'bn2.bias', (hid,)), (p + 'bn2.running_mean', (hid,)),
                    (p + 'bn2.running_var', (hid,)), (p + 'bn2.num_batches_tracked', ())]
            if stride != 1 or ci != hid:
                out += [(p + 'downsample.0.weight', (hid, ci, 1, 1)), (p + 'downsample.1.weight', (hid,)), (p + 'downsample.1.bias', (hid,)),
                        (p + 'downsample.1.running_mean', (hid,)), (p + 'downsample.1.running_var', (hid,)),
                        (p + 'downsample.1.num_batches_tracked', ())]
        inp = hid
    dsc = config.flow.wolf_params['discriminator']
    out += [('discriminator.encoder.net.top.weight', (enc['out_planes'], inp, 1, 1)), ('discriminator.encoder.net.top.bias', (enc['out_planes'],))]
    out += [('discriminator.fc.linear.bias', (2 * dsc['dim'],)), ('discriminator.fc.linear.weight_g', (2 * dsc['dim'], 1)),
            ('discriminator.fc.linear.weight_v', (2 * dsc['dim'], dsc['in_dim']))]
    pr = dsc['prior']
    d, hf = pr['in_features'], pr['hidden_features']
    for t in range(pr['num_steps']):
        p = f'discriminator.prior.flow.steps.{t}.'
        out += [(p + 'actnorm.log_scale', (d,)), (p + 'actnorm.bias', (d,)), (p + 'linear.weight', (d, d)), (p + 'linear.weight_inv', (d, d))]
        for cp in ('coupling1_up', 'coupling1_dn'):
            out += _coupling_shapes(p + f'unit.{cp}.net.', d, hf)
        out += [(p + 'unit.actnorm.log_scale', (d,)), (p + 'unit.actnorm.bias', (d,))]
        for cp in ('coupling2_up', 'coupling2_dn'):
            out += _coupling_shapes(p + f'unit.{cp}.net.', d, hf)
    return out


def _coupling_shapes(p, d, hf):
    return [(p + 'fc1.weight', (hf, d // 2)), (p + 'fc1.bias', (hf,)), (p + 'fc2.weight', (hf, hf)), (p + 'fc2.bias', (hf,)),
            (p + 'fc3.linear.bias', (d,)), (p + 'fc3.linear.weight_g', (d, 1)), (p + 'fc3.linear.weight_v', (d, hf))]


def synth_params(config, seed=0):
    """Deterministic synthetic flow weights (numpy PCG64).  Conv weights get PyTorch-default magnitudes
    (U(+-1/sqrt(fan_in))), for which every row L1 norm exceeds 0.98, so the Lipschitz soft-normalisation is active and
    the residual branch is a genuine contraction (~0.98^3): the fixed-point inverse needs several iterations.
    `linear.weight_inv` is deliberately NOT the inverse of `linear.weight` (the reference never re-syncs it after
    training, SURVEY.md §7 hard part 6), so a product that 'fixes' the quirk fails parity."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in param_shapes(config):
        leaf = name.split('.')[-1]
        if leaf == 'geom_p':
            v = np.asarray(0.0)                       # log(0.5) - log(0.5), iresblock.py:42
        elif leaf == 'lamb':
            v = np.asarray(2.0)
        elif leaf in ('last_n_samples', 'last_firmom', 'last_secmom', 'scale'):
            v = np.zeros(shape)
        elif leaf == 'num_batches_tracked':
            sd[name] = np.asarray(3, dtype=np.int64)
            continue
        elif leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == 'running_mean':
            v = 0.1 * rng.standard_normal(shape)
        elif leaf == 'log_scale':
            v = 0.05 * rng.standard_normal(shape)
        elif leaf == 'weight_g':
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf in ('weight', 'weight_inv') and len(shape) == 2 and shape[0] == shape[1] and 'linear' in name:
            q, _ = np.linalg.qr(rng.standard_normal(shape))
            v = q if leaf == 'weight' else np.linalg.inv(q + 0.05 * rng.standard_normal(shape))
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-b, b, size=shape)
        elif leaf == 'weight':                        # BatchNorm scale
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        else:                                         # biases
            v = 0.05 * rng.standard_normal(shape)
        sd[name] = np.asarray(v, dtype=np.float32)
    return sd


def to_torch(params_np):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params_np.items()}


# ------------------------------------------------------------------------------------------------ residual flow
def sin_act(x):
    """activations.py:11-12"""
    return torch.sin(2. * math.pi * x) / math.pi * 0.5


def lop_weight(w):
    """LopConv2d.compute_weight (lipschitz.py:350-359) for domain = codomain = inf: per-output-channel L1 norm."""
    scale = w.abs().reshape(w.shape[0], -1).sum(dim=1).reshape(-1, 1, 1, 1)
    return w / torch.max(torch.ones(1), scale / COEFF)


def g_branch(P, s, b, first, x, h):
    """iResBlock.nnet_forward (iresblock.py:55-61) for the 3-1-3 Lipschitz conv stack (resflow_.py:432-470)."""
    p = f'generator.flow.transforms.{s}.chain.{b}.nnet.'
    j0 = 0 if first else 1
    if not first:
        x = sin_act(x)
    u = F.conv2d(x, lop_weight(P[p + f'{j0}.weight']), P[p + f'{j0}.bias'], padding=1)
    u = sin_act(u)
    cond = F.linear(h, P[p + f'{j0 + 2}.h_net.net.weight'], P[p + f'{j0 + 2}.h_net.net.bias'])     # lipschitz.py:431
    u = F.conv2d(u + cond[:, :, None, None], lop_weight(P[p + f'{j0 + 2}.weight']), P[p + f'{j0 + 2}.bias'])
    u = sin_act(u)
    return F.conv2d(u, lop_weight(P[p + f'{j0 + 4}.weight']), P[p + f'{j0 + 4}.bias'], padding=1)


def inverse_fixed_point(gfn, y, atol=1e-5, rtol=1e-5, max_iter=1000):
    """iResBlock._inverse_fixed_point (iresblock.py:78-88).  Returns (x, iterations)."""
    x, x_prev = y - gfn(y), y
    i = 0
    tol = atol + y.abs() * rtol
    while not torch.all((x - x_prev) ** 2 / tol < 1):
        x, x_prev = y - gfn(x), x
        i += 1
        if i > max_iter:
            break
    return x, i


def squeeze2(x):
    """squeeze.py:32-45"""
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 4, h // 2, w // 2)


def unsqueeze2(x):
    return F.pixel_shuffle(x, 2)


def resflow_forward(config, P, x, h):
    """ResidualFlow.fwdpass(x, h, eval_logdet=False) (resflow_.py:205-235, 310-324)."""
    nb = n_blocks(config)
    shape = x.shape
    for s, n in enumerate(nb):
        for b in range(n):
            x = x + g_branch(P, s, b, s == 0 and b == 0, x, h)
        if s < len(nb) - 1:
            x = squeeze2(x)
    out = x.reshape(shape[0], -1)
    if len(nb) > 1:
        out = out.view(shape[0], shape[1], 2, 2, shape[2] // 2, shape[3] // 2).permute(0, 1, 4, 2, 5, 3).reshape(shape)
    else:
        out = out.view(shape)
    return out


def resflow_inverse(config, P, z, h, atol=1e-5, rtol=1e-5):
    """ResidualFlow.bwdpass(z, h) (resflow_.py:326-335 -> inverse :237-267).  Returns (x, [iterations per block])."""
    nb = n_blocks(config)
    x = z
    if len(nb) > 1:
        x = x.view(x.shape[0], x.shape[1], x.shape[2] // 2, x.shape[3] // 2, 2, 2).permute(0, 1, 5, 2, 3, 4) \
             .reshape(x.shape[0], x.shape[1], x.shape[2], 2, x.shape[3] // 2).permute(0, 1, 3, 2, 4).reshape(x.shape)
    c0, h0, w0 = flow_input_shape(config)
    k = len(nb) - 1
    x = x.reshape(x.shape[0], c0 * 4 ** k, h0 // 2 ** k, w0 // 2 ** k)    # self.dims[-1], resflow_.py:262
    iters = []
    for s in reversed(range(len(nb))):
        if s < len(nb) - 1:
            x = unsqueeze2(x)          # SqueezeLayer.inverse is the last element of the chain of scale s
        for b in reversed(range(nb[s])):
            first = (s == 0 and b == 0)
            x, it = inverse_fixed_point(lambda v: g_branch(P, s, b, first, v, h), x, atol, rtol)
            iters.append(it)
    return x, iters


# ------------------------------------------------------------------------------------------------ latent prior flow
def _wn(P, p):
    v, g = P[p + 'weight_v'], P[p + 'weight_g']
    return g * v / v.norm(dim=1, keepdim=True)      # nn.utils.weight_norm, dim=0


def _mlp(P, p, z):
    """NICEMLPBlock.forward (blocks.py:27-34), ELU."""
    out = F.elu(F.linear(z, P[p + 'fc1.weight'], P[p + 'fc1.bias']))
    out = F.elu(F.linear(out, P[p + 'fc2.weight'], P[p + 'fc2.bias']))
    return F.linear(out, _wn(P, p + 'fc3.linear.'), P[p + 'fc3.linear.bias'])


def _coupling(P, p, z, skip, up, backward):
    """NICE1d.forward / backward_analytic (coupling.py:87-145) with Affine (transform.py:56-76)."""
    d = z.shape[1]
    if skip:
        z1, z2 = z[:, 0::2], z[:, 1::2]
    else:
        z1, z2 = z[:, :d // 2], z[:, d // 2:]
    zc, zp = (z1, z2) if up else (z2, z1)
    mu, ls = _mlp(P, p + 'net.', zc).chunk(2, dim=1)
    scale = torch.sigmoid(ls + 2.0) + 1e-3
    if backward:
        zp = (zp - mu) / (scale + 1e-12)
        logdet = -scale.log().sum(dim=1)
    else:
        zp = scale * zp + mu
        logdet = scale.log().sum(dim=1)
    z1, z2 = (zc, zp) if up else (zp, zc)
    if skip:
        out = torch.stack([z1, z2], dim=2).reshape(z.shape[0], d)
    else:
        out = torch.cat([z1, z2], dim=1)
    return out, logdet


def _actnorm(P, p, z, backward):
    ls, b = P[p + 'log_scale'], P[p + 'bias']
    if backward:
        return (z - b) / (ls.exp() + 1e-8), -ls.sum() * torch.ones(z.shape[0])
    return z * ls.exp() + b, ls.sum() * torch.ones(z.shape[0])


def prior_flow(config, P, z, backward):
    """PriorFlow.forward / backward (priors/flow.py:172-189).  Returns (out, logdet)."""
    T = config.flow.wolf_params['discriminator']['prior']['num_steps']
    ld = torch.zeros(z.shape[0])
    steps = range(T)
    for t in (reversed(steps) if backward else steps):
        p = f'discriminator.prior.flow.steps.{t}.'
        if not backward:
            z, l = _actnorm(P, p + 'actnorm.', z, False); ld = ld + l
            z = F.linear(z, P[p + 'linear.weight']); ld = ld + torch.slogdet(P[p + 'linear.weight'])[1]
            for name, skip, up in (('coupling1_up', False, True), ('coupling1_dn', False, False)):
                z, l = _coupling(P, p + f'unit.{name}.', z, skip, up, False); ld = ld + l
            z, l = _actnorm(P, p + 'unit.actnorm.', z, False); ld = ld + l
            for name, skip, up in (('coupling2_up', True, True), ('coupling2_dn', True, False)):
                z, l = _coupling(P, p + f'unit.{name}.', z, skip, up, False); ld = ld + l
        else:
            for name, skip, up in (('coupling2_dn', True, False), ('coupling2_up', True, True)):
                z, l = _coupling(P, p + f'unit.{name}.', z, skip, up, True); ld = ld + l
            z, l = _actnorm(P, p + 'unit.actnorm.', z, True); ld = ld + l
            for name, skip, up in (('coupling1_dn', False, False), ('coupling1_up', False, True)):
                z, l = _coupling(P, p + f'unit.{name}.', z, skip, up, True); ld = ld + l
            z = F.linear(z, P[p + 'linear.weight_inv']); ld = ld + torch.slogdet(P[p + 'linear.weight_inv'])[1]
            z, l = _actnorm(P, p + 'actnorm.', z, True); ld = ld + l
    return z, ld


def prior_sample(config, P, eps):
    """FlowPrior.sample (priors/flow.py:226-230): the flow is built with inverse=True, so fwdpass = backward()."""
    return prior_flow(config, P, eps, backward=True)[0]


def wolf_reverse(config, P, z, eps, atol=1e-5, rtol=1e-5):
    """flow_forward(config, flow, z, reverse=True) (flow_model.py:53-67) -> WolfCore.forward(reverse=True)
    (wolf.py:82-89).  `eps` [B,64] is the standard-normal draw of FlowPrior.sample.  Returns (x, h, iters)."""
    if config.flow.squeeze:
        z = squeeze2(z)
    h = prior_sample(config, P, eps)
    x, iters = resflow_inverse(config, P, z, h, atol, rtol)
    x = x.reshape(z.shape)
    if config.flow.squeeze:
        x = unsqueeze2(x)
    return x, h, iters


# ------------------------------------------------------------------------------------------------ posterior q(h|x) and KL
def _bn_eval(P, p, x, eps=1e-5):
    """nn.BatchNorm2d in eval mode (running statistics)."""
    sc = P[p + 'weight'] / torch.sqrt(P[p + 'running_var'] + eps)
    return x * sc[None, :, None, None] + (P[p + 'bias'] - P[p + 'running_mean'] * sc)[None, :, None, None]


def _bn_train(P, p, x, eps=1e-5):
    """nn.BatchNorm2d in training mode: batch statistics over (N, H, W), biased variance (running buffers not touched here)."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * P[p + 'weight'][None, :, None, None] + P[p + 'bias'][None, :, None, None]


def encoder_forward(config, P, x, training=False):
    """GlobalResNetEncoderBatchNorm.forward (modules/encoders/global_encoder.py:12-36) over ResNetBlockBatchNorm
    (nnet/resnets/resnet_batchnorm.py:18-76), ELU activation; batch norm with running statistics (eval) or batch
    statistics (training=True, what the joint training step runs).  Returns [B, out_planes*h*w]."""
    enc = config.flow.wolf_params['discriminator']['encoder']
    _bn = _bn_train if training else _bn_eval
    for lv, hid in enumerate(enc['hidden_planes']):
        for m, stride in enumerate((1, 2)):
            p = f'discriminator.encoder.net.resnet{lv}.main.{m}.'
            out = F.conv2d(x, P[p + 'conv1.weight'], stride=stride, padding=1)
            out = F.elu(_bn(P, p + 'bn1.', out))
            out = _bn(P, p + 'bn2.', F.conv2d(out, P[p + 'conv2.weight'], padding=1))
            if (p + 'downsample.0.weight') in P:
                x = _bn(P, p + 'downsample.1.', F.conv2d(x, P[p + 'downsample.0.weight'], stride=stride))
            x = F.elu(out + x)
    x = F.elu(F.conv2d(x, P['discriminator.encoder.net.top.weight'], P['discriminator.encoder.net.top.bias']))
    return x.reshape(x.shape[0], -1)


def posterior_sample_and_kl(config, P, x, eps, training=False):
    """GaussianDiscriminator.sampling_and_KL (modules/discriminators/gaussian.py:67-76) with nsamples = 1 and the
    reparameterisation noise `eps` [B, 64] supplied; FlowPrior.calcKL (priors/flow.py:233-253).  Returns (h, KL, mu, logvar)."""
    c = encoder_forward(config, P, x, training)
    c = F.linear(c, _wn(P, 'discriminator.fc.linear.'), P['discriminator.fc.linear.bias'])
    mu, logvar = c.chunk(2, dim=1)
    h = eps * torch.exp(0.5 * logvar) + mu
    dim = h.shape[1]
    cc = math.log(math.pi * 2.)
    log_posterior = -0.5 * ((logvar + eps ** 2).sum(dim=1) + cc * dim)
    e, logdet = prior_flow(config, P, h, backward=False)           # flow built with inverse=True: bwdpass = forward()
    log_prior = -0.5 * ((e * e).sum(dim=1) + cc * dim) + logdet
    return h, log_posterior - log_prior, mu, logvar


# ------------------------------------------------------------------------------------------------ log-det estimators
def poisson_1mcdf(lamb, k, offset):
    """iresblock.py:306-318: P(n >= k - offset), 1 for k <= offset."""
    if k <= offset:
        return 1.
    k = k - offset
    s = 1.
    for i in range(1, k):
        s += lamb ** i / math.factorial(i)
    return 1 - math.exp(-lamb) * s


def series_coefficients(n, training, lamb=2.0, n_exact_terms=2):
    """(K, [c_1..c_K]) of iResBlock._logdetgrad (iresblock.py:114-132) for one Poisson draw n:
    K = n + offset terms, c_k = 1[n >= k - offset] / P(N >= k - offset), offset = n_exact_terms (train) or 20 (eval)."""
    offset = n_exact_terms if training else 20
    K = n + offset
    return K, [1.0 / poisson_1mcdf(lamb, k, offset) * (1.0 if n >= k - offset else 0.0) for k in range(1, K + 1)]


def block_logdet(gfn, x, n, vareps, training, differentiable=False):
    """g(x) and the power-series log-det estimate of one iResBlock (iresblock.py:90-174): `basic_logdet_estimator`
    (:253-261) in eval mode, `neumann_logdet_estimator` (:264-273) in training mode.  Uses autograd for the VJPs, like the
    reference.  Returns (g, logdet [B]).  differentiable=True keeps the autograd graph exactly like the reference's training step
    (:264-273: the Neumann vector is a constant, the last VJP is taken with create_graph) so that the gradient of the estimate
    w.r.t. parameters, x and h — second order in g — can be taken by the caller."""
    K, coef = series_coefficients(n, training)
    if differentiable and not training:
        raise ValueError('the differentiable estimator is the training-mode one')
    with torch.enable_grad():
        if differentiable:
            if not x.requires_grad:
                x = x.requires_grad_(True)
        else:
            x = x.detach().requires_grad_(True)
        g = gfn(x)
        B = x.shape[0]
        if training:
            vjp, neumann = vareps, vareps
            with torch.no_grad():
                for k in range(1, K + 1):
                    vjp = torch.autograd.grad(g, x, vjp, retain_graph=True)[0]
                    neumann = neumann + (-1) ** k * coef[k - 1] * vjp
            vj = torch.autograd.grad(g, x, neumann, retain_graph=True, create_graph=differentiable)[0]
            ld = (vj.reshape(B, -1) * vareps.reshape(B, -1)).sum(1)
            if differentiable:
                return g, ld
        else:
            vjp, ld = vareps, torch.zeros(B)
            for k in range(1, K + 1):
                vjp = torch.autograd.grad(g, x, vjp, retain_graph=True)[0]
                ld = ld + (-1) ** (k + 1) / k * coef[k - 1] * (vjp.reshape(B, -1) * vareps.reshape(B, -1)).sum(1)
    return g.detach(), ld.detach()


def resflow_forward_logdet(config, P, x, h, ns, varepss, training=False, differentiable=False):
    """ResidualFlow.fwdpass(x, h, eval_logdet=True) (resflow_.py:310-324): returns (z in image layout, logpx [B]) with
    logpx = -sum_blocks logdet (iresblock.py:63-69 with logpx starting at 0).  ns / varepss: per-block Poisson draws and
    Gaussian probe tensors, in forward block order."""
    nb = n_blocks(config)
    shape = x.shape
    logpx = torch.zeros(x.shape[0])
    i = 0
    for s, n_s in enumerate(nb):
        for b in range(n_s):
            first = (s == 0 and b == 0)
            g, ld = block_logdet(lambda v: g_branch(P, s, b, first, v, h), x, int(ns[i]), varepss[i], training, differentiable)
            x = x + g
            logpx = logpx - ld
            i += 1
        if s < len(nb) - 1:
            x = squeeze2(x)
    out = x.reshape(shape[0], -1)
    if len(nb) > 1:
        out = out.view(shape[0], shape[1], 2, 2, shape[2] // 2, shape[3] // 2).permute(0, 1, 4, 2, 5, 3).reshape(shape)
    else:
        out = out.view(shape)
    return out, logpx


def wolf_forward(config, P, x, eps_post, ns, varepss, training=False):
    """flow_forward(config, flow, x, reverse=False) (flow_model.py:53-67) -> WolfCore.forward (wolf.py:90-130):
    returns (z, logdet - KL) where logdet = sum of block log-dets (`fwdpass` returns logpx = -sum, wolf.py:126 negates)."""
    if config.flow.squeeze:
        x = squeeze2(x)
    h, kl, _, _ = posterior_sample_and_kl(config, P, x, eps_post)
    z, logpx = resflow_forward_logdet(config, P, x, h, ns, varepss, training)
    if config.flow.squeeze:
        z = unsqueeze2(z)
    return z, -logpx - kl, h, kl


def wolf_train_forward(config, P, x, eps_post, ns, varepss):
    """WolfCore.forward(reverse=False) in TRAINING mode (wolf.py:90-128), differentiable exactly like the reference's graph:
    batch-statistics encoder -> h ~ q(h|x), KL -> residual flow with the Neumann log-det series (n + 2 terms, constant Neumann
    vector, last VJP with create_graph).  Returns (z, logdet - KL, h, KL); torch.autograd of any scalar of (z, logdet - KL) w.r.t.
    the entries of P is the gradient the joint training step applies (losses.py:300-304)."""
    if config.flow.squeeze:
        x = squeeze2(x)
    h, kl, _, _ = posterior_sample_and_kl(config, P, x, eps_post, training=True)
    z, logpx = resflow_forward_logdet(config, P, x, h, ns, varepss, training=True, differentiable=True)
    if config.flow.squeeze:
        z = unsqueeze2(z)
    return z, -logpx - kl, h, kl


# ------------------------------------------------------------------------------------------------ replayed draws (fixtures)
def replay_draws(config, seed, B, training=False):
    """Every random draw one wolf forward consumes (flow_model.py:53-67 -> wolf.py:90-130), regenerated from one numpy PCG64
    seed in a fixed order, so the full-size fixtures (tests/golden/make_golden.py:make_fullflow / make_fulljoint /
    make_fulllikelihood) only need to store the seed: x ~ U(-1, 1) image batch, eps_post (reparameterisation noise,
    gaussian.py:67-76), ns (Poisson(2) series lengths, iresblock.py:306), varepss (one Gaussian probe per iResBlock,
    iresblock.py:107), z_rev / eps_rev (latent and prior noise of a reverse pass), Gz / cl (cotangents of a training backward)."""
    rng = np.random.default_rng(seed)
    S, C = config.data.image_size, config.data.num_channels
    layout = block_layout(config)
    c0, h0, w0 = flow_input_shape(config)
    d = dict(x=rng.uniform(-1, 1, size=(B, C, S, S)).astype(np.float32),
             eps_post=rng.standard_normal((B, 64)).astype(np.float32),
             ns=rng.poisson(2.0, size=len(layout)).astype(np.int64))
    d['varepss'] = [rng.standard_normal((B, c, h0 >> s, w0 >> s)).astype(np.float32) for (s, b, c, first) in layout]
    d['z_rev'] = rng.standard_normal((B, C, S, S)).astype(np.float32)
    d['eps_rev'] = rng.standard_normal((B, 64)).astype(np.float32)
    d['Gz'] = rng.standard_normal((B, C, S, S)).astype(np.float32)
    d['cl'] = rng.uniform(0.5, 1.5, size=(B,)).astype(np.float32)
    return d


def grad_digest(name, g, seed=97, n_sub=256):
    """What a full-size fixture keeps of one parameter gradient (13 M flow + 63 M score parameters do not fit a committed file):
    its L2 norm, its projection on a seeded Gaussian direction (every element takes part) and a strided sub-sample."""
    import zlib
    g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
    r = np.random.default_rng([seed, zlib.crc32(name.encode())]).standard_normal(g.size).astype(np.float32)
    step = max(1, g.size // n_sub)
    return np.float64(np.linalg.norm(g.astype(np.float64))), np.float64(np.dot(g.astype(np.float64), r.astype(np.float64)) / np.sqrt(g.size)), \
        g[::step][:n_sub].copy()
