"""CPU: the oracle (oracle/) against the golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

from helpers import load_npz, load_json, tiny, rel_l2
from indm_b200 import configs
from oracle import ops as oops, sde as osde, ncsnpp as oncsnpp, sampler as osampler

UPFIRDN_CASES = ['up2', 'down2', 'pyr', 'k3', 'crop', 'gen', 'up2_32']


@pytest.mark.parametrize("case", UPFIRDN_CASES)
def test_upfirdn2d_oracle_matches_reference(case):
    g = load_npz('ops.npz')
    up, down, p0, p1 = [int(v) for v in g[f'upfirdn_{case}_args']]
    y = oops.upfirdn2d(g[f'upfirdn_{case}_x'], g[f'upfirdn_{case}_k'], up=up, down=down, pad=(p0, p1))
    assert y.shape == g[f'upfirdn_{case}_y'].shape
    np.testing.assert_allclose(y, g[f'upfirdn_{case}_y'], rtol=1e-5, atol=1e-6)
    x = g[f'upfirdn_{case}_x']
    gx = oops.upfirdn2d_backward(g[f'upfirdn_{case}_gy'], g[f'upfirdn_{case}_k'], up, down, (p0, p1), x.shape[2:])
    assert gx.shape == x.shape
    np.testing.assert_allclose(gx, g[f'upfirdn_{case}_gx'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", ['4d', '2d', '3d'])
def test_fused_leaky_relu_oracle_matches_reference(case):
    g = load_npz('ops.npz')
    y = oops.fused_leaky_relu(g[f'lrelu_{case}_x'], g[f'lrelu_{case}_b'])
    np.testing.assert_allclose(y, g[f'lrelu_{case}_y'], rtol=1e-6, atol=1e-7)
    gx, gb = oops.fused_leaky_relu_backward(g[f'lrelu_{case}_gy'], y)
    np.testing.assert_allclose(gx, g[f'lrelu_{case}_gx'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gb, g[f'lrelu_{case}_gb'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ['vp', 've', 've90'])
def test_sde_oracle_matches_reference(tag):
    g = load_npz('sde.npz')
    t, x, u = torch.from_numpy(g['t']), torch.from_numpy(g['x']), torch.from_numpy(g['is_u' if False else f'{tag}_is_u'])
    sde = {'vp': osde.VP(), 've': osde.VE(sigma_max=50), 've90': osde.VE(sigma_max=90.)}[tag]
    d, gg = sde.sde(x, t)
    np.testing.assert_allclose(d.numpy(), g[f'{tag}_drift'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gg.numpy(), g[f'{tag}_diff'], rtol=1e-6)
    mean, std = sde.marginal_prob(x, t)
    np.testing.assert_allclose(mean.numpy(), g[f'{tag}_mean'], rtol=1e-6)
    np.testing.assert_allclose(std.numpy(), g[f'{tag}_std'], rtol=1e-6)
    f, G = sde.discretize(x, t)
    np.testing.assert_allclose(f.numpy(), g[f'{tag}_disc_f'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(G.numpy(), g[f'{tag}_disc_G'], rtol=1e-6)
    f, G = sde.discretize(x, t, t * 0.9)
    np.testing.assert_allclose(f.numpy(), g[f'{tag}_disc2_f'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(G.numpy(), g[f'{tag}_disc2_G'], rtol=1e-6)
    np.testing.assert_allclose(sde.prior_logp(x).numpy(), g[f'{tag}_prior_logp'], rtol=1e-6)
    tt, Z = sde.importance_time(u, 1e-5)
    np.testing.assert_allclose(float(Z), float(g[f'{tag}_Z']), rtol=1e-6)
    np.testing.assert_allclose(tt.numpy(), g[f'{tag}_is_t'], rtol=1e-5, atol=1e-7)


def _cfg(tag):
    base = {'tiny_vp': 'vp/CIFAR10/indm_fid', 'tiny_ve': 've/CIFAR10/indm',
            'vp_cifar': 'vp/CIFAR10/indm_fid', 've_cifar': 've/CIFAR10/indm'}[tag]
    cfg = configs.get_config(base)
    if tag.startswith('tiny'):
        tiny(cfg)
    return cfg


@pytest.mark.parametrize("tag", ['tiny_vp', 'tiny_ve', 'vp_cifar', 've_cifar'])
def test_param_shapes_match_reference_state_dict(tag):
    want = [(k, tuple(s)) for k, s in load_json(f'shapes_{tag}.json')]
    got = [(k, tuple(s)) for k, s in oncsnpp.param_shapes(_cfg(tag))]
    assert got == want


@pytest.mark.parametrize("tag", ['tiny_vp', 'tiny_ve', 'vp_cifar', 've_cifar'])
def test_ncsnpp_oracle_matches_reference(tag):
    g = load_npz(f'ncsnpp_{tag}.npz')
    cfg = _cfg(tag)
    P = oncsnpp.to_torch(oncsnpp.synth_params(cfg, int(g['seed'])))
    sde = osde.get_sde(cfg)
    with torch.no_grad():
        s = oncsnpp.score_fn(cfg, sde, P, torch.from_numpy(g['x']), torch.from_numpy(g['t']))
    assert rel_l2(s.numpy(), g['score']) < 2e-5, rel_l2(s.numpy(), g['score'])


@pytest.mark.parametrize("tag,corr", [('tiny_vp', 'none'), ('tiny_ve', 'langevin')])
def test_pc_sampler_oracle_matches_reference(tag, corr):
    g = load_npz(f'pc_{tag}.npz')
    cfg = _cfg(tag)
    P = oncsnpp.to_torch(oncsnpp.synth_params(cfg, 11))
    sde = osde.get_sde(cfg)
    sf = lambda x, t: oncsnpp.score_fn(cfg, sde, P, x, t)
    prior = torch.from_numpy(g['prior'])
    if tag.endswith('ve'):
        prior = prior * sde.sigma_max      # sde_lib.py:293
    with torch.no_grad():
        out = osampler.pc_sampler(sde, sf, prior, [torch.from_numpy(n) for n in g['noises']],
                                  int(g['num_scales']), float(g['eps']), float(g['snr']), corrector=corr)
    assert rel_l2(out.numpy(), g['out']) < 1e-4, rel_l2(out.numpy(), g['out'])
