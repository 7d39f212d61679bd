#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs (small, committed): tests/golden/{configs.json, ops.npz, sde.npz, ncsnpp_*.npz, shapes_*.json, pc_*.npz}

Everything the oracle (`oracle/`) claims is pinned by these files: they hold seeded inputs and the outputs the
reference's own code produced for them.  Weights are not stored: they are regenerated bit-identically from
`oracle.ncsnpp.synth_params(config, seed)` (numpy PCG64) and loaded into the reference model's state-dict.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
from oracle import ncsnpp as oncsnpp  # noqa: E402


def tiny(cfg, nf=128):
    """Shrink a reference config to a fast test size (same code paths)."""
    cfg.model.nf = nf
    cfg.model.ch_mult = (1, 2)
    cfg.model.num_res_blocks = 1
    cfg.model.attn_resolutions = (8,)
    cfg.data.image_size = 16
    return cfg


def plain(cfg):
    out = {}
    for k, v in cfg.items():
        if hasattr(v, 'items'):
            out[k] = plain(v)
        elif isinstance(v, torch.device):
            out[k] = str(v)
        elif isinstance(v, tuple):
            out[k] = list(v)
        else:
            out[k] = v
    return out


def make_configs():
    names = ['ve/CIFAR10/indm', 've/CELEBA/indm', 'vp/CIFAR10/indm_fid', 'vp/CIFAR10/indm_nll',
             'vp/CELEBA/indm_fid', 'vp/CELEBA/indm_nll']
    out = {}
    for n in names:
        c = plain(rl.get_config(f'configs/{n}.py'))
        c.pop('device')
        out[n] = c
    with open(os.path.join(HERE, 'configs.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


def make_ops():
    op = rl.load('op')
    rng = np.random.default_rng(1)
    out = {}
    k4 = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32)
    k4 /= k4.sum()
    cases = {
        'up2':   dict(shape=(2, 5, 8, 8), k=k4 * 4, up=2, down=1, pad=(2, 1)),      # upsample_2d
        'down2': dict(shape=(2, 5, 8, 8), k=k4, up=1, down=2, pad=(1, 1)),          # downsample_2d
        'pyr':   dict(shape=(2, 3, 8, 8), k=k4, up=1, down=1, pad=(2, 2)),          # conv_downsample_2d FIR stage
        'k3':    dict(shape=(1, 2, 7, 9), k=rng.standard_normal((3, 3)).astype(np.float32), up=1, down=1, pad=(1, 1)),
        'crop':  dict(shape=(1, 2, 9, 9), k=rng.standard_normal((2, 2)).astype(np.float32), up=2, down=3, pad=(-1, 2)),
        'gen':   dict(shape=(2, 3, 6, 5), k=rng.standard_normal((5, 5)).astype(np.float32), up=3, down=2, pad=(3, 1)),
        'up2_32': dict(shape=(1, 4, 32, 32), k=k4 * 4, up=2, down=1, pad=(2, 1)),
    }
    for name, c in cases.items():
        x = rng.standard_normal(c['shape']).astype(np.float32)
        y = op.upfirdn2d(torch.from_numpy(x), torch.from_numpy(c['k']), up=c['up'], down=c['down'], pad=c['pad'])
        out[f'upfirdn_{name}_x'] = x
        out[f'upfirdn_{name}_k'] = c['k']
        out[f'upfirdn_{name}_args'] = np.array([c['up'], c['down'], c['pad'][0], c['pad'][1]], dtype=np.int64)
        out[f'upfirdn_{name}_y'] = y.numpy()
        # gradient wrt input via autograd through the native path (the reference CUDA op's backward must equal it)
        xt = torch.from_numpy(x).requires_grad_(True)
        yy = op.upfirdn2d(xt, torch.from_numpy(c['k']), up=c['up'], down=c['down'], pad=c['pad'])
        gy = rng.standard_normal(tuple(yy.shape)).astype(np.float32)
        yy.backward(torch.from_numpy(gy))
        out[f'upfirdn_{name}_gy'] = gy
        out[f'upfirdn_{name}_gx'] = xt.grad.numpy()
    for name, shape in {'4d': (2, 6, 5, 7), '2d': (4, 6), '3d': (2, 6, 9)}.items():
        x = rng.standard_normal(shape).astype(np.float32)
        b = rng.standard_normal((shape[1],)).astype(np.float32)
        y = op.fused_leaky_relu(torch.from_numpy(x), torch.from_numpy(b))   # defaults: slope 0.2, scale sqrt(2)
        out[f'lrelu_{name}_x'] = x
        out[f'lrelu_{name}_b'] = b
        out[f'lrelu_{name}_y'] = y.numpy()
        xt = torch.from_numpy(x).requires_grad_(True)
        bt = torch.from_numpy(b).requires_grad_(True)
        yy = op.fused_leaky_relu(xt, bt)
        gy = rng.standard_normal(shape).astype(np.float32)
        yy.backward(torch.from_numpy(gy))
        out[f'lrelu_{name}_gy'] = gy
        out[f'lrelu_{name}_gx'] = xt.grad.numpy()
        out[f'lrelu_{name}_gb'] = bt.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)


def make_sde():
    sde_lib = rl.load('sde_lib')
    out = {}
    t = torch.tensor([1e-5, 1e-3, 0.02, 0.25, 0.5, 0.77, 0.999, 1.0])
    x = torch.from_numpy(np.random.default_rng(2).standard_normal((8, 3, 4, 4)).astype(np.float32))
    u = torch.from_numpy(np.random.default_rng(3).uniform(size=8).astype(np.float32))
    out['t'] = t.numpy(); out['x'] = x.numpy(); out['u'] = u.numpy()
    for tag, sde in (('vp', sde_lib.VPSDE(truncation_time=1e-5, beta_min=0.1, beta_max=20., N=1000)),
                     ('ve', sde_lib.VESDE(truncation_time=1e-5, sigma_min=0.01, sigma_max=50, N=1000)),
                     ('ve90', sde_lib.VESDE(truncation_time=1e-5, sigma_min=0.01, sigma_max=90., N=1000))):
        d, g = sde.sde(x, t)
        out[f'{tag}_drift'] = d.numpy(); out[f'{tag}_diff'] = g.numpy()
        mean, std = sde.marginal_prob(x, t)
        out[f'{tag}_mean'] = mean.numpy(); out[f'{tag}_std'] = std.numpy()
        f, G = sde.discretize(x, t, None)
        out[f'{tag}_disc_f'] = f.numpy(); out[f'{tag}_disc_G'] = G.numpy()
        nt = t * 0.9
        f, G = sde.discretize(x, t, nt)
        out[f'{tag}_disc2_f'] = f.numpy(); out[f'{tag}_disc2_G'] = G.numpy()
        out[f'{tag}_prior_logp'] = sde.prior_logp(x).numpy()
        Z = sde.normalizing_constant(1e-5)
        out[f'{tag}_Z'] = np.asarray(float(Z))

        class _C:  # get_diffusion_time reads config.training.importance_sampling only when arg is None
            pass
        torch.manual_seed(7)
        u_ref = torch.rand(8)
        torch.manual_seed(7)
        tt, ZZ = sde.get_diffusion_time(None, 8, 'cpu', 1e-5, importance_sampling=True)
        out[f'{tag}_is_u'] = u_ref.numpy(); out[f'{tag}_is_t'] = tt.numpy()
    np.savez_compressed(os.path.join(HERE, 'sde.npz'), **out)


def ref_model(cfg, seed):
    mutils, _ = rl.load('models.utils', 'models.ncsnpp')
    model = mutils.create_model(cfg)        # DataParallel wrapper, CPU
    ref_sd = model.module.state_dict()
    shapes = [(k, list(v.shape)) for k, v in ref_sd.items()]
    sd = oncsnpp.synth_params(cfg, seed)
    assert [(k, list(v.shape)) for k, v in sd.items()] == shapes, 'oracle.param_shapes != reference state_dict'
    model.module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model.eval()
    return model, shapes


def make_ncsnpp(only=None):
    mutils, sde_lib = rl.load('models.utils', 'sde_lib')
    specs = {
        'tiny_vp': ('configs/vp/CIFAR10/indm_fid.py', True, 3),
        'tiny_ve': ('configs/ve/CIFAR10/indm.py', True, 3),
        'vp_cifar': ('configs/vp/CIFAR10/indm_fid.py', False, 2),
        've_cifar': ('configs/ve/CIFAR10/indm.py', False, 2),
        # BASELINE configs 4 / 5: 3x64x64, the deep variant (model.num_res_blocks = 8), one image
        've_celeba': ('configs/ve/CELEBA/indm.py', False, 1),
        'vp_celeba': ('configs/vp/CELEBA/indm_nll.py', False, 1),
    }
    for tag, (path, is_tiny, B) in specs.items():
        if only is not None and tag not in only:
            continue
        cfg = rl.get_config(path)
        if tag.endswith('celeba'):
            cfg.model.num_res_blocks = 8
        if is_tiny:
            tiny(cfg)
        model, shapes = ref_model(cfg, seed=11)
        with open(os.path.join(HERE, f'shapes_{tag}.json'), 'w') as f:
            json.dump(shapes, f)
        S = cfg.data.image_size
        rng = np.random.default_rng(5)
        x = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        t = np.array([0.9, 0.31, 0.02][:B], dtype=np.float32)
        sde = sde_lib.get_sde(cfg)
        if cfg.training.sde == 'vesde':
            x = x * np.asarray(sde.marginal_prob(torch.zeros(B), torch.from_numpy(t))[1].numpy()).reshape(B, 1, 1, 1).astype(np.float32)
        score_fn = mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)
        with torch.no_grad():
            s = score_fn(torch.from_numpy(x), torch.from_numpy(t))
            if cfg.training.sde == 'vpsde':
                raw = model(torch.from_numpy(x), torch.from_numpy(t) * 999)
            else:
                raw = model(torch.from_numpy(x), sde.marginal_prob(torch.zeros(B), torch.from_numpy(t))[1])
        np.savez_compressed(os.path.join(HERE, f'ncsnpp_{tag}.npz'), x=x, t=t, score=s.numpy(), raw=raw.numpy(),
                            seed=np.asarray(11))
        print(tag, 'score rms', float(s.pow(2).mean().sqrt()), 'raw rms', float(raw.pow(2).mean().sqrt()))


def make_flowfwd():
    """Wolf flow FORWARD with log-det in eval mode (posterior encoder + reparameterisation + prior-flow KL + 20+n term
    power-series log-det of every iResBlock) through the reference's flow_forward, with every random draw replayed:
    torch.randn (posterior eps), torch.randn_like (Hutchinson probes), poisson_sample (series lengths)."""
    fm = rl.load('flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    for tag, path, squeeze in (('tiny', 'configs/vp/CIFAR10/indm_nll.py', False), ('tiny_sq', 'configs/vp/CELEBA/indm_nll.py', True)):
        # the posterior encoder needs the full image extent (3 stride-2 levels -> 4x4x8 = in_dim 128): keep S, shrink depth
        cfg = rl.get_config(path)
        tiny_flow(cfg, squeeze)
        S = 64 if squeeze else 32
        cfg.data.image_size = cfg.flow.image_size = S
        with rl.reference_cwd():
            flow = fm.create_flow_model(cfg)
        pcfg = pconfigs.get_config(path.replace('configs/', '').replace('.py', ''))
        tiny_flow(pcfg, squeeze)
        pcfg.data.image_size = pcfg.flow.image_size = S
        sd = oflow.synth_params(pcfg, 21)
        flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        flow.eval()
        B = 2
        rng = np.random.default_rng(41)
        x = rng.uniform(-1, 1, size=(B, 3, S, S)).astype(np.float32)
        eps_post = rng.standard_normal((B, 64)).astype(np.float32)
        layout = oflow.block_layout(pcfg)
        c0, h0, w0 = oflow.flow_input_shape(pcfg)
        ns = rng.poisson(2.0, size=len(layout)).astype(np.int64)
        varepss = [rng.standard_normal((B, c, h0 >> s, w0 >> s)).astype(np.float32) for (s, b, c, first) in layout]
        q_eps, q_n = [torch.from_numpy(v) for v in varepss], list(ns)
        real = (torch.randn, torch.randn_like, irb.poisson_sample)
        torch.randn = lambda *a, **k: torch.from_numpy(eps_post).reshape(B, 1, 64)
        torch.randn_like = lambda t, **k: q_eps.pop(0)
        irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
        try:
            z, ldkl = fm.flow_forward(cfg, flow, torch.from_numpy(x), reverse=False)
        finally:
            torch.randn, torch.randn_like, irb.poisson_sample = real
        assert not q_eps and not q_n
        out = dict(x=x, eps_post=eps_post, ns=ns, z=z.detach().numpy(), ldkl=ldkl.detach().numpy(), seed=np.asarray(21))
        for i, v in enumerate(varepss):
            out[f'vareps_{i}'] = v
        np.savez_compressed(os.path.join(HERE, f'flowfwd_{tag}.npz'), **out)
        print('flowfwd', tag, 'ns', ns, 'ldkl', ldkl.detach().numpy())


def small_joint(cfg):
    """Small score net + small wolf flow on 32x32 images (the posterior encoder needs the full extent): same code paths as
    configs/vp/CIFAR10/indm_nll.py."""
    cfg.model.nf = 128
    cfg.model.ch_mult = (1, 2)
    cfg.model.num_res_blocks = 1
    cfg.model.attn_resolutions = (16,)
    cfg.flow.nblocks = '2-2'
    cfg.flow.intermediate_dim = 128
    return cfg


def make_likelihood():
    """likelihood.get_likelihood_fn (PF-ODE NLL, RK45 rtol = atol = 1e-3) and likelihood.get_elbo_fn of the live reference on a
    small INDM-VP model, every random draw replayed (torch.randn / randn_like / randint_like / rand, poisson_sample)."""
    mutils, sde_lib, likelihood, fm = rl.load('models.utils', 'sde_lib', 'likelihood', 'flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path = 'configs/vp/CIFAR10/indm_nll.py'
    cfg = small_joint(rl.get_config(path))
    model, _ = ref_model(cfg, seed=11)
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = small_joint(pconfigs.get_config('vp/CIFAR10/indm_nll'))
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, 21).items()})
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    B, S = 2, 32
    rng = np.random.default_rng(51)
    data = rng.uniform(-1, 1, size=(B, 3, S, S)).astype(np.float32)
    layout = oflow.block_layout(pcfg)
    c0, h0, w0 = oflow.flow_input_shape(pcfg)
    inverse_scaler = lambda v: (v + 1.) / 2.
    out = dict(data=data, seed_score=np.asarray(11), seed_flow=np.asarray(21))
    for which in ('nll', 'elbo'):
        eps_post = rng.standard_normal((B, 64)).astype(np.float32)
        ns = rng.poisson(2.0, size=len(layout)).astype(np.int64)
        varepss = [rng.standard_normal((B, c, h0 >> s, w0 >> s)).astype(np.float32) for (s, b, c, first) in layout]
        rad = (rng.integers(0, 2, size=(B, 3, S, S)).astype(np.float32))         # randint_like result in {0, 1}
        gauss = [rng.standard_normal((B, 3, S, S)).astype(np.float32) for _ in range(4)]
        u = rng.uniform(size=(B,)).astype(np.float32)
        q_like = [torch.from_numpy(v) for v in varepss] + [torch.from_numpy(v) for v in gauss]
        q_n = list(ns)
        real = (torch.randn, torch.randn_like, torch.randint_like, torch.rand, irb.poisson_sample)
        torch.randn = lambda *a, **k: torch.from_numpy(eps_post).reshape(B, 1, 64)
        torch.randn_like = lambda t, **k: q_like.pop(0)
        torch.randint_like = lambda t, **k: torch.from_numpy(rad)
        torch.rand = lambda *a, **k: torch.from_numpy(u)
        irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
        try:
            if which == 'nll':
                fn = likelihood.get_likelihood_fn(cfg, sde, inverse_scaler, rtol=1e-3, atol=1e-3)
                bpd, z, nfe = fn(model, flow, torch.from_numpy(data), eps_bpd=1e-5)
                out.update(nll_bpd=bpd.detach().numpy(), nll_z=z.detach().numpy(), nll_nfe=np.asarray(nfe))
                used = 4 + 3         # vareps x4, perturbation z, residual x2
            else:
                fn = likelihood.get_elbo_fn(cfg, sde, inverse_scaler)
                a, b = fn(model, flow, torch.from_numpy(data))
                out.update(elbo_bpd=a.detach().numpy(), elbo_bpd_residual=b.detach().numpy())
                used = 4 + 4         # vareps x4, z, lp_z, residual x2
        finally:
            torch.randn, torch.randn_like, torch.randint_like, torch.rand, irb.poisson_sample = real
        assert len(q_like) == 8 - used and not q_n, (len(q_like), q_n)
        out.update({f'{which}_eps_post': eps_post, f'{which}_ns': ns, f'{which}_rad': rad, f'{which}_u': u})
        for i, v in enumerate(varepss):
            out[f'{which}_vareps_{i}'] = v
        for i, v in enumerate(gauss):
            out[f'{which}_gauss_{i}'] = v
        print('likelihood', which, {k: v for k, v in out.items() if k.endswith('bpd') or k.endswith('nfe') or k.endswith('residual')})
    np.savez_compressed(os.path.join(HERE, 'likelihood_small_vp.npz'), **out)


def _sub(a, limit=4096):
    """flattened strided subsample (<= limit elements) of a tensor: keeps the fixtures small; tests index the same way"""
    f = np.ascontiguousarray(a).reshape(-1)
    return f[::max(1, (f.size + limit - 1) // limit)].copy()


def make_train():
    """One optimisation step of the score network through the reference's losses.get_step_fn (flow.model = 'identity',
    dropout = 0 so that no mask has to be replayed): raw gradients of every parameter (their norms + a few full tensors),
    the per-sample losses, and the parameters / EMA after clip + AdamW."""
    mutils, sde_lib, losses, ema_mod = rl.load('models.utils', 'sde_lib', 'losses', 'models.ema')
    for tag, path in (('tiny_vp', 'configs/vp/CIFAR10/indm_fid.py'), ('tiny_ve', 'configs/ve/CIFAR10/indm.py')):
        cfg = rl.get_config(path)
        tiny(cfg)
        cfg.model.dropout = 0.0
        cfg.flow.model = 'identity'
        cfg.training.importance_sampling = True
        model, _ = ref_model(cfg, seed=11)
        model.train()
        sde = sde_lib.get_sde(cfg)
        B, S = 4, cfg.data.image_size
        rng = np.random.default_rng(61)
        batch = rng.uniform(-1, 1, size=(B, 3, S, S)).astype(np.float32)
        u = rng.uniform(size=(B,)).astype(np.float32)
        z = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        real = (torch.rand, torch.randn_like)

        def patched():
            torch.rand = lambda *a, **k: torch.from_numpy(u)
            torch.randn_like = lambda t, **k: torch.from_numpy(z)

        def restore():
            torch.rand, torch.randn_like = real
        # (1) raw gradients
        loss_fn = losses.get_sde_loss_fn(cfg, sde, train=True)
        patched()
        try:
            ls = loss_fn(model, torch.from_numpy(batch))
        finally:
            restore()
        torch.mean(ls).backward()
        names = [n for n, p in model.named_parameters() if p.requires_grad]
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
        gnorm = np.array([float(grads[n].norm()) for n in names], dtype=np.float64)
        keep = [n for n in names if any(k in n for k in ('all_modules.0.', 'all_modules.1.', 'all_modules.2.', 'all_modules.3.Conv_0.weight',
                                                          'all_modules.3.Dense_0', 'all_modules.3.GroupNorm_1', 'all_modules.4.NIN_1.W',
                                                          'all_modules.4.NIN_3', 'all_modules.5.Conv_2.weight'))]
        out = dict(batch=batch, u=u, z=z, losses_raw=ls.detach().numpy(), grad_norms=gnorm, names=np.array(names), keep=np.array(keep),
                   total_norm=np.asarray(float(np.sqrt((gnorm ** 2).sum()))), seed=np.asarray(11))
        for n in keep:
            out['grad::' + n] = _sub(grads[n].numpy())
        # (2) the full step on a fresh model
        model, _ = ref_model(cfg, seed=11)
        opt = losses.get_optimizer(cfg, model.parameters())
        ema = ema_mod.ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
        state = dict(optimizer=opt, model=model, ema=ema, step=0)
        step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
        patched()
        try:
            res = step_fn(state, None, torch.from_numpy(batch))
        finally:
            restore()
        out['losses_step'] = res[0].numpy()
        sd = dict(model.named_parameters())
        for n in keep:
            out['param::' + n] = _sub(sd[n].detach().numpy())
        idx = {n: i for i, n in enumerate(names)}
        for n in keep:
            out['ema::' + n] = _sub(ema.shadow_params[idx[n]].numpy())
        np.savez_compressed(os.path.join(HERE, f'train_{tag}.npz'), **out)
        print('train', tag, 'losses', out['losses_step'], 'total grad norm', float(out['total_norm']), 'kept', len(keep))


def make_vjp():
    """Input vector-Jacobian products of the reference score function (what likelihood.get_div_fn builds through autograd,
    likelihood.py:27-38): J^T eps with Rademacher eps, and the Hutchinson contraction eps^T J eps from the reference's own div_fn."""
    mutils, sde_lib, likelihood = rl.load('models.utils', 'sde_lib', 'likelihood')
    for tag, (path, is_tiny, B) in {'tiny_vp': ('configs/vp/CIFAR10/indm_fid.py', True, 3),
                                    'vp_cifar': ('configs/vp/CIFAR10/indm_nll.py', False, 2),
                                    'tiny_ve': ('configs/ve/CIFAR10/indm.py', True, 3)}.items():
        cfg = rl.get_config(path)
        if is_tiny:
            tiny(cfg)
        model, _ = ref_model(cfg, seed=11)
        S = cfg.data.image_size
        rng = np.random.default_rng(31)
        x = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        t = np.array([0.6, 0.6, 0.6][:B], dtype=np.float32)      # the ODE evaluates one t for the whole batch (likelihood.py:96)
        eps = (rng.integers(0, 2, size=(B, 3, S, S)).astype(np.float32) * 2 - 1)
        sde = sde_lib.get_sde(cfg)
        if cfg.training.sde == 'vesde':
            x = x * float(sde.marginal_prob(torch.zeros(1), torch.tensor([0.6]))[1])
        score_fn = mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)
        xt = torch.from_numpy(x).requires_grad_(True)
        sc = score_fn(xt, torch.from_numpy(t))
        vjp, = torch.autograd.grad((sc * torch.from_numpy(eps)).sum(), xt)
        div = likelihood.get_div_fn(lambda xx, tt: score_fn(xx, tt))(torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(eps))
        np.savez_compressed(os.path.join(HERE, f'vjp_{tag}.npz'), x=x, t=t, eps=eps, score=sc.detach().numpy(), vjp=vjp.numpy(),
                            div=div.detach().numpy(), seed=np.asarray(11))
        print('vjp', tag, 'vjp rms', float(vjp.pow(2).mean().sqrt()), 'div', div.detach().numpy())


def make_pc():
    """Short PC trajectories through the reference's own get_pc_sampler, with randn_like / randn replayed."""
    mutils, sde_lib, sampling = rl.load('models.utils', 'sde_lib', 'sampling')
    import tempfile
    for tag, path, corr in (('tiny_vp', 'configs/vp/CIFAR10/indm_fid.py', 'none'),
                            ('tiny_ve', 'configs/ve/CIFAR10/indm.py', 'langevin')):
        cfg = rl.get_config(path)
        tiny(cfg)
        cfg.sampling.method = 'pc'
        cfg.sampling.predictor = 'reverse_diffusion'
        cfg.sampling.corrector = corr
        cfg.sampling.num_scales = 6
        cfg.flow.model = 'identity'    # isolate the PC loop; the flow inverse has its own fixtures
        model, _ = ref_model(cfg, seed=11)
        sde = sde_lib.get_sde(cfg)
        B, S = 3, cfg.data.image_size
        rng = np.random.default_rng(9)
        n_draws = cfg.sampling.num_scales * (2 if corr == 'langevin' else 1)
        prior = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        noises = rng.standard_normal((n_draws, B, 3, S, S)).astype(np.float32)
        q = [torch.from_numpy(n) for n in noises]
        real_randn, real_randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *shape, **kw: torch.from_numpy(prior)
        torch.randn_like = lambda x, **kw: q.pop(0)
        try:
            fn = sampling.get_sampling_fn(cfg, sde, (B, 3, S, S), lambda v: v, cfg.sampling.truncation_time)
            with tempfile.TemporaryDirectory() as d:
                before, after, nfe = fn(model, None, sample_dir=d, r=0)
        finally:
            torch.randn, torch.randn_like = real_randn, real_randn_like
        assert not q
        np.savez_compressed(os.path.join(HERE, f'pc_{tag}.npz'), prior=prior, noises=noises, out=before.numpy(),
                            nfe=np.asarray(nfe), num_scales=np.asarray(cfg.sampling.num_scales),
                            eps=np.asarray(cfg.sampling.truncation_time), snr=np.asarray(cfg.sampling.snr))
        print('pc', tag, 'rms', float(before.pow(2).mean().sqrt()), 'nfe', nfe)


def make_samplers():
    """The other sampling components of sampling.py through the reference's own get_sampling_fn, noise replayed:
    euler_maruyama / ancestral_sampling predictors, the annealed-Langevin corrector, pc_sampler_search (sampling.pc_denoise) and
    the black-box ODE sampler (the default of configs/vp/*/indm_fid.py)."""
    mutils, sde_lib, sampling = rl.load('models.utils', 'sde_lib', 'sampling')
    import tempfile
    out = {}
    specs = [('em_vp', 'configs/vp/CIFAR10/indm_fid.py', dict(method='pc', predictor='euler_maruyama', corrector='none', num_scales=6)),
             # ancestral_sampling cannot be pinned: the reference's pc_sampler calls update_fn(x, t, next_t) but
             # AncestralSamplingPredictor.update_fn takes (x, t) only (sampling.py:245 vs :351) -> TypeError in the reference itself
             ('ald_ve', 'configs/ve/CIFAR10/indm.py', dict(method='pc', predictor='reverse_diffusion', corrector='ald', num_scales=6)),
             ('search_ve', 'configs/ve/CIFAR10/indm.py', dict(method='pc', predictor='reverse_diffusion', corrector='langevin', pc_denoise=True)),
             ('ode_vp', 'configs/vp/CIFAR10/indm_fid.py', dict(method='ode'))]
    for tag, path, over in specs:
        cfg = rl.get_config(path)
        tiny(cfg)
        cfg.flow.model = 'identity'
        for k, v in over.items():
            setattr(cfg.sampling, k, v)
        if tag == 'search_ve':
            cfg.model.num_scales = 8                      # sde.N = 8: pc_sampler_search runs N - 1 steps + the denoising step
        if tag == 'ode_vp':
            cfg.eval.rtol = cfg.eval.atol = 1e-3
        model, _ = ref_model(cfg, seed=11)
        sde = sde_lib.get_sde(cfg)
        B, S = 3, cfg.data.image_size
        rng = np.random.default_rng(71)
        prior = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        noises = rng.standard_normal((40, B, 3, S, S)).astype(np.float32)
        q = [torch.from_numpy(n) for n in noises]
        real_randn, real_randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *shape, **kw: torch.from_numpy(prior)
        torch.randn_like = lambda x, **kw: q.pop(0)
        try:
            fn = sampling.get_sampling_fn(cfg, sde, (B, 3, S, S), lambda v: v, cfg.sampling.truncation_time)
            with tempfile.TemporaryDirectory() as d:
                before, after, nfe = fn(model, None, sample_dir=d, r=0)
        finally:
            torch.randn, torch.randn_like = real_randn, real_randn_like
        used = 40 - len(q)
        out.update({f'{tag}_prior': prior, f'{tag}_noises': noises[:used], f'{tag}_out': before.numpy(), f'{tag}_nfe': np.asarray(nfe)})
        print('sampler', tag, 'rms', float(before.pow(2).mean().sqrt()), 'nfe', nfe, 'noise draws', used)
    np.savez_compressed(os.path.join(HERE, 'samplers_tiny.npz'), **out)


def make_flowtrain():
    """Wolf flow in TRAINING mode through the reference's flow_forward (posterior encoder with batch-statistics BatchNorm,
    reparameterisation, prior-flow KL, Neumann-series log-det with n + 2 terms whose gradient is second order in g), then
    backward of L = <z, Gz> + <logdet - KL, cl> with fixed cotangents: records z, logdet - KL, h, and the gradient of EVERY
    flow parameter.  Every random draw is replayed (torch.randn, torch.randn_like, poisson_sample)."""
    fm = rl.load('flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path = 'configs/vp/CIFAR10/indm_nll.py'
    cfg = rl.get_config(path)
    tiny_flow(cfg, False)
    S = 32
    cfg.data.image_size = cfg.flow.image_size = S
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = pconfigs.get_config('vp/CIFAR10/indm_nll')
    tiny_flow(pcfg, False)
    pcfg.data.image_size = pcfg.flow.image_size = S
    sd = oflow.synth_params(pcfg, 23)
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    flow.train()
    B = 4
    rng = np.random.default_rng(43)
    x = rng.uniform(-1, 1, size=(B, 3, S, S)).astype(np.float32)
    eps_post = rng.standard_normal((B, 64)).astype(np.float32)
    layout = oflow.block_layout(pcfg)
    c0, h0, w0 = oflow.flow_input_shape(pcfg)
    ns = np.array([1, 0, 2, 3][:len(layout)], dtype=np.int64)
    varepss = [rng.standard_normal((B, c, h0 >> s, w0 >> s)).astype(np.float32) for (s, b, c, first) in layout]
    Gz = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    cl = rng.uniform(0.5, 1.5, size=(B,)).astype(np.float32)
    q_eps, q_n = [torch.from_numpy(v) for v in varepss], list(ns)
    real = (torch.randn, torch.randn_like, irb.poisson_sample)
    torch.randn = lambda *a, **k: torch.from_numpy(eps_post).reshape(B, 1, 64)
    torch.randn_like = lambda t, **k: q_eps.pop(0)
    irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
    cap = {}
    disc = flow.module.discriminator
    real_skl = disc.sampling_and_KL

    def skl(xx, y=None, nsamples=1):
        h, kl = real_skl(xx, y=y, nsamples=nsamples)
        h.retain_grad()
        cap['h'], cap['kl'] = h, kl
        return h, kl
    disc.sampling_and_KL = skl

    def keep(name):
        def hook(mod, inp, outp):
            outp.retain_grad()
            cap[name] = outp
        return hook
    hooks = [disc.fc.register_forward_hook(keep('fc_out')), disc.encoder.register_forward_hook(keep('enc_out'))]
    # x must NOT require grad up front: the first block would then differentiate g through the encoder path (h = enc(x)) too,
    # which the real training step (data batch without grad, requires_grad_ set inside _logdetgrad after the encoder ran) never does
    xt = torch.from_numpy(x)
    try:
        z, ldkl = fm.flow_forward(cfg, flow, xt, reverse=False)
        loss = (z * torch.from_numpy(Gz)).sum() + (ldkl * torch.from_numpy(cl)).sum()
        loss.backward(retain_graph=True)
        gh_total = cap['h'].grad.clone()       # read now: the retain_grad hook fires again in the autograd.grad call below
        gh_kl = torch.autograd.grad(-(cap['kl'].reshape(B) * torch.from_numpy(cl)).sum(), cap['h'])[0]
    finally:
        torch.randn, torch.randn_like, irb.poisson_sample = real
        disc.sampling_and_KL = real_skl
        for hk in hooks:
            hk.remove()
    assert not q_eps and not q_n
    out = dict(x=x, eps_post=eps_post, ns=ns, Gz=Gz, cl=cl, z=z.detach().numpy(), ldkl=ldkl.detach().numpy(),
               h=cap['h'].detach().numpy().reshape(B, 64), kl=cap['kl'].detach().numpy().reshape(B),
               gh=gh_total.numpy().reshape(B, 64), gh_kl=gh_kl.numpy().reshape(B, 64), fc_out=cap['fc_out'].detach().numpy(),
               g_fc_out=cap['fc_out'].grad.numpy(), enc_out=cap['enc_out'].detach().numpy(), g_enc_out=cap['enc_out'].grad.numpy(),
               seed=np.asarray(23))
    for i, v in enumerate(varepss):
        out[f'vareps_{i}'] = v
    nograd = []
    for k, p_ in flow.module.named_parameters():
        if p_.grad is None:
            nograd.append(k)
        else:
            out['grad.' + k] = p_.grad.numpy()
    for k, b_ in flow.module.named_buffers():          # BatchNorm running statistics after the step, Lipschitz scales
        out['buf.' + k] = b_.detach().numpy()
    np.savez_compressed(os.path.join(HERE, 'flowtrain_tiny.npz'), **out)
    print('flowtrain ns', ns, 'ldkl', ldkl.detach().numpy(), 'kl', out['kl'], 'params without grad:', nograd)
    print('gh', np.linalg.norm(out['gh']), 'gh_kl', np.linalg.norm(out['gh_kl']), 'g_fc_out', np.linalg.norm(out['g_fc_out']))


def make_jointtrain():
    """One JOINT optimisation step of the flow and the score network through the reference's own losses.get_step_fn ->
    flow_step_fn_nll (losses.py:258-320) on the small INDM-VP model (dropout 0, so no mask to replay), every random draw replayed:
    the four loss vectors, and parameters / EMA of both networks after clip + AdamW (a sample of tensors, sub-sampled)."""
    mutils, sde_lib, losses, ema_mod, fm = rl.load('models.utils', 'sde_lib', 'losses', 'models.ema', 'flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path = 'configs/vp/CIFAR10/indm_nll.py'
    cfg = small_joint(rl.get_config(path))
    cfg.model.dropout = 0.0
    model, _ = ref_model(cfg, seed=11)
    model.train()
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = small_joint(pconfigs.get_config('vp/CIFAR10/indm_nll'))
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, 21).items()})
    sde = sde_lib.get_sde(cfg)
    B, S = 4, 32
    rng = np.random.default_rng(71)
    batch = rng.uniform(-1, 1, size=(B, 3, S, S)).astype(np.float32)
    layout = oflow.block_layout(pcfg)
    c0, h0, w0 = oflow.flow_input_shape(pcfg)
    eps_post = rng.standard_normal((B, 64)).astype(np.float32)
    ns = np.array([2, 0, 1, 3][:len(layout)], dtype=np.int64)
    varepss = [rng.standard_normal((B, c, h0 >> s, w0 >> s)).astype(np.float32) for (s, b, c, first) in layout]
    z = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    logp_noise = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    u = rng.uniform(size=(B,)).astype(np.float32)
    q_like = [torch.from_numpy(v) for v in varepss] + [torch.from_numpy(z), torch.from_numpy(logp_noise)]
    q_n = list(ns)
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ema_mod.ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    state = dict(optimizer=opt, model=model, ema=ema, step=0)
    fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
    fema = ema_mod.ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate)
    flow_state = dict(optimizer=fopt, model=flow, ema=fema, step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    before = {'s::' + n: p.detach().clone() for n, p in model.named_parameters()}
    before.update({'f::' + n: p.detach().clone() for n, p in flow.named_parameters()})
    real = (torch.randn, torch.randn_like, torch.rand, irb.poisson_sample)
    torch.randn = lambda *a, **k: torch.from_numpy(eps_post).reshape(B, 1, 64)
    torch.randn_like = lambda t, **k: q_like.pop(0)
    torch.rand = lambda *a, **k: torch.from_numpy(u)
    irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
    try:
        res = step_fn(state, flow_state, torch.from_numpy(batch))
    finally:
        torch.randn, torch.randn_like, torch.rand, irb.poisson_sample = real
    assert not q_like and not q_n, (len(q_like), q_n)
    out = dict(batch=batch, eps_post=eps_post, ns=ns, z=z, logp_noise=logp_noise, u=u, seed_score=np.asarray(11), seed_flow=np.asarray(21),
               losses=res[0].numpy(), losses_score=res[1].numpy(), losses_flow=res[2].numpy(), losses_logp=res[3].numpy())
    for i, v in enumerate(varepss):
        out[f'vareps_{i}'] = v
    keep_s = ('all_modules.0.weight', 'all_modules.2.weight', 'all_modules.3.Conv_0.weight', 'all_modules.3.GroupNorm_1.bias')
    keep_f = ('transforms.0.chain.0.nnet.0.weight', 'transforms.0.chain.1.nnet.3.weight', 'transforms.1.chain.1.nnet.5.weight',
              'chain.1.nnet.3.h_net.net.weight', 'resnet0.main.0.conv1.weight', 'resnet1.main.1.bn2.weight', 'resnet2.main.1.conv2.weight',
              'encoder.net.top.weight', 'fc.linear.weight_v', 'fc.linear.weight_g', 'steps.0.linear.weight', 'steps.1.actnorm.log_scale',
              'steps.1.unit.coupling2_dn.net.fc2.weight', 'steps.0.unit.coupling1_up.net.fc3.linear.weight_v')
    names, shadow = [], {}
    for tag, net, em, keep in (('s', model, ema, keep_s), ('f', flow, fema, keep_f)):
        plist = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        for i, (n, p) in enumerate(plist):
            if any(k in n for k in keep):
                key = f'{tag}::{n}'
                names.append(key)
                out['param::' + key] = _sub(p.detach().numpy())
                out['step::' + key] = _sub((p.detach() - before[key]).numpy())
                out['ema::' + key] = _sub(em.shadow_params[i].numpy())
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'jointtrain_small_vp.npz'), **out)
    print('jointtrain losses', res[0].numpy(), 'score', res[1].numpy(), 'flow', res[2].numpy(), 'logp', res[3].numpy(), 'kept', len(names))


def tiny_flow(cfg, squeeze):
    """Small wolf flow with the same code paths: 2+2 iResBlocks, 128 hidden channels, 16x16 images."""
    cfg.flow.nblocks = '2-2'
    cfg.flow.intermediate_dim = 128
    cfg.data.image_size = 16
    cfg.flow.image_size = 16
    cfg.flow.squeeze = squeeze
    return cfg


def make_flow():
    """Wolf flow reverse pass (prior sample of h + fixed-point inverse of every iResBlock) and the residual-flow
    forward (no log-det) with h given, through the reference's own flow_forward / fwdpass."""
    from oracle import flow as oflow
    fm = rl.load('flow_models.flow_model')
    for tag, path, squeeze in (('tiny', 'configs/vp/CIFAR10/indm_fid.py', False), ('tiny_sq', 'configs/vp/CELEBA/indm_fid.py', True)):
        cfg = rl.get_config(path)
        tiny_flow(cfg, squeeze)
        with rl.reference_cwd():
            flow = fm.create_flow_model(cfg)
        ref_sd = flow.module.state_dict()
        shapes = [(k, list(v.shape)) for k, v in ref_sd.items()]
        # the oracle's shape table is driven by the restated JSON content carried in indm_b200.configs
        from indm_b200 import configs as pconfigs
        pcfg = pconfigs.get_config(path)
        tiny_flow(pcfg, squeeze)
        sd = oflow.synth_params(pcfg, 21)
        assert [(k, list(v.shape)) for k, v in sd.items()] == shapes, 'oracle.flow.param_shapes != reference state_dict'
        flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        flow.eval()
        with open(os.path.join(HERE, f'shapes_flow_{tag}.json'), 'w') as f:
            json.dump(shapes, f)
        B, S = 3, 16
        rng = np.random.default_rng(23)
        z = rng.standard_normal((B, 3, S, S)).astype(np.float32)
        eps = rng.standard_normal((B, 64)).astype(np.float32)
        real_randn = torch.randn
        torch.randn = lambda *a, **k: torch.from_numpy(eps)
        try:
            with torch.no_grad():
                x, _ = fm.flow_forward(cfg, flow, torch.from_numpy(z), log_det=None, reverse=True)
                h = flow.module.discriminator.sample_from_prior(B, torch.device('cpu'))
        finally:
            torch.randn = real_randn
        # forward (no logdet) of the generator flow with that h, on the flow's own input layout
        from flow_models.resflow.layers.squeeze import SqueezeLayer
        xin = torch.from_numpy(rng.standard_normal((B, 3, S, S)).astype(np.float32))
        xin_f = SqueezeLayer(2).forward(xin) if squeeze else xin
        with torch.no_grad():
            zf = flow.module.generator.flow.fwdpass(xin_f, h, eval_logdet=False)
        np.savez_compressed(os.path.join(HERE, f'flow_{tag}.npz'), z=z, eps=eps, x=x.numpy(), h=h.numpy(), xin=xin.numpy(),
                            zf=zf.numpy(), seed=np.asarray(21))
        print('flow', tag, 'x rms', float(x.pow(2).mean().sqrt()), 'max|x-z|', float((x - torch.from_numpy(z)).abs().max()),
              'zf rms', float(zf.pow(2).mean().sqrt()))


# ------------------------------------------------------------------------------------------------ full-size fixtures (round 2)
FULL = {'cifar': ('configs/vp/CIFAR10/indm_nll.py', 'vp/CIFAR10/indm_nll'), 'celeba': ('configs/vp/CELEBA/indm_nll.py', 'vp/CELEBA/indm_nll')}


def _full_flow(tag, seed=21):
    fm = rl.load('flow_models.flow_model')
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path, pname = FULL[tag]
    cfg = rl.get_config(path)
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = pconfigs.get_config(pname)
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, seed).items()})
    return fm, cfg, pcfg, flow


def make_ncsnpp_celeba():
    make_ncsnpp(only=('ve_celeba', 'vp_celeba'))


def make_fullflow():
    """The wolf flow AT THE BENCHED SIZE — flow.nblocks = '16-16', flow.intermediate_dim = 512 (configs/ve/CIFAR10/indm.py:68-69;
    CIFAR 3x32x32 and CelebA 3x64x64 squeezed to 12x32x32) — through the live reference's flow_forward, batch 2, every random draw
    replayed from ONE seed (oracle.flow.replay_draws): (1) the reverse pass (prior sample of h + fixed-point inverse of all 32
    blocks), (2) the eval-mode forward with the (20 + n)-term power-series log-det and the KL, (3) the training-mode forward
    (batch-statistics encoder, Neumann series) and the gradient of EVERY flow parameter for fixed cotangents — kept as
    norm / seeded projection / sub-sample per tensor (oracle.flow.grad_digest)."""
    import time
    from oracle import flow as oflow
    for tag in ('cifar', 'celeba'):
        fm, cfg, pcfg, flow = _full_flow(tag)
        import flow_models.wolf.flows.resflow.layers.iresblock as irb
        B, draw_seed = 2, 141
        d = oflow.replay_draws(pcfg, draw_seed, B)
        out = dict(seed=np.asarray(21), draw_seed=np.asarray(draw_seed), B=np.asarray(B))
        real = (torch.randn, torch.randn_like, irb.poisson_sample)
        # (1) reverse
        flow.eval()
        torch.randn = lambda *a, **k: torch.from_numpy(d['eps_rev'])
        t0 = time.time()
        try:
            with torch.no_grad():
                x_rev, _ = fm.flow_forward(cfg, flow, torch.from_numpy(d['z_rev']), log_det=None, reverse=True)
        finally:
            torch.randn = real[0]
        out['x_rev'] = x_rev.numpy()
        print('fullflow', tag, 'reverse', round(time.time() - t0, 1), 's  max|x - z|', float((x_rev - torch.from_numpy(d['z_rev'])).abs().max()))
        # the posterior sample and its KL are captured too, so a log-det error and a KL error can be told apart
        cap = {}
        disc = flow.module.discriminator
        real_skl = disc.sampling_and_KL

        def skl(xx, y=None, nsamples=1):
            h_, kl_ = real_skl(xx, y=y, nsamples=nsamples)
            cap['h'], cap['kl'] = h_.detach().clone(), kl_.detach().clone()
            return h_, kl_
        disc.sampling_and_KL = skl
        # (2) eval-mode forward with log-det and KL
        q_eps, q_n = [torch.from_numpy(v) for v in d['varepss']], list(d['ns'])
        torch.randn = lambda *a, **k: torch.from_numpy(d['eps_post']).reshape(B, 1, 64)
        torch.randn_like = lambda t, **k: q_eps.pop(0)
        irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
        t0 = time.time()
        try:
            z, ldkl = fm.flow_forward(cfg, flow, torch.from_numpy(d['x']), reverse=False)
        finally:
            torch.randn, torch.randn_like, irb.poisson_sample = real
        assert not q_eps and not q_n
        out.update(z_eval=z.detach().numpy(), ldkl_eval=ldkl.detach().numpy(), h_eval=cap['h'].numpy().reshape(B, 64),
                   kl_eval=cap['kl'].numpy().reshape(B))
        print('fullflow', tag, 'eval forward', round(time.time() - t0, 1), 's  ldkl', ldkl.detach().numpy())
        # (3) training-mode forward + backward
        flow.train()
        q_eps, q_n = [torch.from_numpy(v) for v in d['varepss']], list(d['ns'])
        torch.randn = lambda *a, **k: torch.from_numpy(d['eps_post']).reshape(B, 1, 64)
        torch.randn_like = lambda t, **k: q_eps.pop(0)
        irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
        t0 = time.time()
        try:
            z, ldkl = fm.flow_forward(cfg, flow, torch.from_numpy(d['x']), reverse=False)
            loss = (z * torch.from_numpy(d['Gz'])).sum() + (ldkl * torch.from_numpy(d['cl'])).sum()
            loss.backward()
        finally:
            torch.randn, torch.randn_like, irb.poisson_sample = real
        assert not q_eps and not q_n
        out.update(z_train=z.detach().numpy(), ldkl_train=ldkl.detach().numpy(), h_train=cap['h'].numpy().reshape(B, 64),
                   kl_train=cap['kl'].numpy().reshape(B))
        disc.sampling_and_KL = real_skl
        names, norms, projs = [], [], []
        for k, p_ in flow.module.named_parameters():
            if p_.grad is None:
                continue
            n_, pr_, sub_ = oflow.grad_digest(k, p_.grad.numpy())
            names.append(k); norms.append(n_); projs.append(pr_)
            out['gsub.' + k] = sub_
        out.update(grad_names=np.array(names), grad_norms=np.array(norms), grad_projs=np.array(projs))
        print('fullflow', tag, 'train fwd+bwd', round(time.time() - t0, 1), 's  ldkl', ldkl.detach().numpy(), 'grads', len(names),
              'total norm', float(np.sqrt((np.array(norms) ** 2).sum())))
        np.savez_compressed(os.path.join(HERE, f'flowfull_{tag}.npz'), **out)


def make_fulljoint():
    """One JOINT optimisation step (losses.get_step_fn -> flow_step_fn_nll, losses.py:258-320) of the live reference at the
    BENCHED sizes: configs/vp/CIFAR10/indm_nll.py with the full DDPM++ (nres = 4, nf = 128) and the full wolf flow (16-16 / 512),
    batch 2, dropout 0 (no mask to replay), every random draw replayed: the four loss vectors, and for EVERY parameter of both
    networks the applied update (p_after - p_before) as norm / projection / sub-sample."""
    import time
    mutils, sde_lib, losses, ema_mod, fm = rl.load('models.utils', 'sde_lib', 'losses', 'models.ema', 'flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path, pname = FULL['cifar']
    cfg = rl.get_config(path)
    cfg.model.dropout = 0.0
    model, _ = ref_model(cfg, seed=11)
    model.train()
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = pconfigs.get_config(pname)
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, 21).items()})
    sde = sde_lib.get_sde(cfg)
    B, draw_seed = 2, 171
    d = oflow.replay_draws(pcfg, draw_seed, B)
    rng = np.random.default_rng(draw_seed + 1)
    S = cfg.data.image_size
    z = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    logp_noise = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    u = rng.uniform(size=(B,)).astype(np.float32)
    q_like = [torch.from_numpy(v) for v in d['varepss']] + [torch.from_numpy(z), torch.from_numpy(logp_noise)]
    q_n = list(d['ns'])
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ema_mod.ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    state = dict(optimizer=opt, model=model, ema=ema, step=0)
    fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
    fema = ema_mod.ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate)
    flow_state = dict(optimizer=fopt, model=flow, ema=fema, step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    before = {'s::' + n: p.detach().clone() for n, p in model.named_parameters()}
    before.update({'f::' + n: p.detach().clone() for n, p in flow.named_parameters()})
    real = (torch.randn, torch.randn_like, torch.rand, irb.poisson_sample)
    torch.randn = lambda *a, **k: torch.from_numpy(d['eps_post']).reshape(B, 1, 64)
    torch.randn_like = lambda t, **k: q_like.pop(0)
    torch.rand = lambda *a, **k: torch.from_numpy(u)
    irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
    t0 = time.time()
    try:
        res = step_fn(state, flow_state, torch.from_numpy(d['x']))
    finally:
        torch.randn, torch.randn_like, torch.rand, irb.poisson_sample = real
    assert not q_like and not q_n, (len(q_like), q_n)
    out = dict(draw_seed=np.asarray(draw_seed), B=np.asarray(B), z=z, logp_noise=logp_noise, u=u, seed_score=np.asarray(11), seed_flow=np.asarray(21),
               losses=res[0].numpy(), losses_score=res[1].numpy(), losses_flow=res[2].numpy(), losses_logp=res[3].numpy())
    names, norms, projs = [], [], []
    for tag, net in (('s', model), ('f', flow)):
        for n, p in net.named_parameters():
            if not p.requires_grad:
                continue
            key = f'{tag}::{n}'
            n_, pr_, sub_ = oflow.grad_digest(key, (p.detach() - before[key]).numpy())
            names.append(key); norms.append(n_); projs.append(pr_)
            out['usub.' + key] = sub_
    out.update(names=np.array(names), upd_norms=np.array(norms), upd_projs=np.array(projs))
    np.savez_compressed(os.path.join(HERE, 'jointfull_vp.npz'), **out)
    print('fulljoint', round(time.time() - t0, 1), 's losses', res[0].numpy(), 'score', res[1].numpy(), 'flow', res[2].numpy(), 'logp', res[3].numpy(),
          'tensors', len(names))


def make_fulllikelihood():
    """likelihood.get_likelihood_fn (PF-ODE NLL, RK45 at the reference's default rtol = atol = 1e-5) and likelihood.get_elbo_fn of
    the live reference AT THE BENCHED SIZES (configs/vp/CIFAR10/indm_nll.py: full DDPM++ + full wolf flow), batch 2, every random
    draw replayed."""
    import time
    mutils, sde_lib, likelihood, fm = rl.load('models.utils', 'sde_lib', 'likelihood', 'flow_models.flow_model')
    import flow_models.wolf.flows.resflow.layers.iresblock as irb
    from oracle import flow as oflow
    from indm_b200 import configs as pconfigs
    path, pname = FULL['cifar']
    cfg = rl.get_config(path)
    model, _ = ref_model(cfg, seed=11)
    smooth = bool(os.environ.get('INDM_GOLDEN_SMOOTH'))
    if smooth:
        # the same weights with the reference's ~0-initialised tensors damped x0.1 (oracle.ncsnpp.damp_zero_init): a well-conditioned ODE
        sd = oncsnpp.damp_zero_init(oncsnpp.synth_params(cfg, 11))
        model.module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    with rl.reference_cwd():
        flow = fm.create_flow_model(cfg)
    pcfg = pconfigs.get_config(pname)
    flow.module.load_state_dict({k: torch.from_numpy(v) for k, v in oflow.synth_params(pcfg, 21).items()})
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    B, S = 2, 32
    inverse_scaler = lambda v: (v + 1.) / 2.
    out = dict(seed_score=np.asarray(11), seed_flow=np.asarray(21), B=np.asarray(B), damp=np.asarray(0.1 if smooth else 1.0))
    fname = 'likelihood_full_vp_smooth.npz' if smooth else 'likelihood_full_vp.npz'
    for which, draw_seed in (('nll', 151), ('elbo', 152)):
        d = oflow.replay_draws(pcfg, draw_seed, B)
        rng = np.random.default_rng(draw_seed + 1000)
        rad = (rng.integers(0, 2, size=(B, 3, S, S)).astype(np.float32))         # randint_like result in {0, 1}
        gauss = [rng.standard_normal((B, 3, S, S)).astype(np.float32) for _ in range(4)]
        u = rng.uniform(size=(B,)).astype(np.float32)
        q_like = [torch.from_numpy(v) for v in d['varepss']] + [torch.from_numpy(v) for v in gauss]
        q_n = list(d['ns'])
        real = (torch.randn, torch.randn_like, torch.randint_like, torch.rand, irb.poisson_sample)
        torch.randn = lambda *a, **k: torch.from_numpy(d['eps_post']).reshape(B, 1, 64)
        torch.randn_like = lambda t, **k: q_like.pop(0)
        torch.randint_like = lambda t, **k: torch.from_numpy(rad)
        torch.rand = lambda *a, **k: torch.from_numpy(u)
        irb.poisson_sample = lambda lamb, m: np.array([q_n.pop(0)])
        t0 = time.time()
        try:
            if which == 'nll':
                fn = likelihood.get_likelihood_fn(cfg, sde, inverse_scaler)
                bpd, z, nfe = fn(model, flow, torch.from_numpy(d['x']), eps_bpd=1e-5)
                out.update(nll_bpd=bpd.detach().numpy(), nll_z=z.detach().numpy(), nll_nfe=np.asarray(nfe))
                used = 3             # perturbation z, residual x2
            else:
                fn = likelihood.get_elbo_fn(cfg, sde, inverse_scaler)
                a, b = fn(model, flow, torch.from_numpy(d['x']))
                out.update(elbo_bpd=a.detach().numpy(), elbo_bpd_residual=b.detach().numpy())
                used = 4             # z, lp_z, residual x2
        finally:
            torch.randn, torch.randn_like, torch.randint_like, torch.rand, irb.poisson_sample = real
        assert len(q_like) == 4 - used and not q_n, (len(q_like), q_n)
        out.update({f'{which}_draw_seed': np.asarray(draw_seed), f'{which}_rad': rad.astype(np.int8), f'{which}_u': u})
        print('fulllikelihood', which, round(time.time() - t0, 1), 's', {k: v for k, v in out.items() if k.endswith('bpd') or k.endswith('nfe') or k.endswith('residual')})
    if os.environ.get('INDM_GOLDEN_SENSITIVITY'):
        # noise floor of the fixture: the same reference run with another intra-op thread count (fp32 summation order in the
        # convolutions changes at the 1e-7 level) — how far the reference's OWN NLL moves says how tight a parity bound can be
        old = dict(np.load(os.path.join(HERE, fname)))
        print('sensitivity: threads', torch.get_num_threads(), {k: (out[k] - old[k]).tolist() for k in ('nll_bpd', 'elbo_bpd', 'elbo_bpd_residual')},
              'nfe', int(out['nll_nfe']), 'vs', int(old['nll_nfe']),
              'latent rel-L2', float(np.linalg.norm(out['nll_z'] - old['nll_z']) / np.linalg.norm(old['nll_z'])))
        return
    np.savez_compressed(os.path.join(HERE, fname), **out)


if __name__ == '__main__':
    if os.environ.get('INDM_GOLDEN_THREADS'):
        torch.set_num_threads(int(os.environ['INDM_GOLDEN_THREADS']))
    which = sys.argv[1:] or ['configs', 'ops', 'sde', 'ncsnpp', 'pc', 'flow', 'vjp', 'flowfwd', 'likelihood', 'train', 'samplers', 'flowtrain', 'jointtrain']
    for w in which:
        globals()['make_' + w]()
        print('made', w)
