"""indm_b200/sde_lib.py (pure per-sample scalar glue: runs on any device) against the golden vectors taken from the live
reference's sde_lib (tests/golden/sde.npz, made by tests/golden/make_golden.py), plus the reverse process and the per-step
scalars the fused sampler kernels are fed with."""
import numpy as np
import pytest
import torch

from helpers import load_npz
from indm_b200 import configs, sde_lib


def _make(tag):
    return {'vp': sde_lib.VPSDE(), 've': sde_lib.VESDE(sigma_max=50), 've90': sde_lib.VESDE(sigma_max=90.)}[tag]


@pytest.mark.parametrize("tag", ['vp', 've', 've90'])
def test_sde_lib_matches_reference_golden(tag):
    g = load_npz('sde.npz')
    t, x, u = torch.from_numpy(g['t']), torch.from_numpy(g['x']), torch.from_numpy(g[f'{tag}_is_u'])
    sde = _make(tag)
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
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    tt, Z = sde.get_diffusion_time(cfg, u.shape[0], u.device, 1e-5, importance_sampling=True, u=u)
    np.testing.assert_allclose(float(Z), float(g[f'{tag}_Z']), rtol=1e-6)
    np.testing.assert_allclose(tt.numpy(), g[f'{tag}_is_t'], rtol=1e-5, atol=1e-7)
    assert sde.T == 1 and sde.N == 1000 and sde.eps == 1e-5


@pytest.mark.parametrize("tag", ['vp', 've'])
@pytest.mark.parametrize("pf", [False, True])
def test_reverse_process_follows_the_reference_formulas(tag, pf):
    """sde_lib.py:74-120: drift - g^2 score (x 1/2 for the probability flow), zero diffusion for the ODE; the last discretised step
    onto t = 0 has no forward drift and G = g(t) sqrt(t - next_t)."""
    sde = _make(tag)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(4, 3, 8, 8, generator=gen)
    t = torch.rand(4, generator=gen) * 0.9 + 0.05
    score = lambda a, b: torch.tanh(a) * (1 + b[:, None, None, None])
    rev = sde.reverse(score, probability_flow=pf)
    assert rev.N == sde.N and rev.T == sde.T and rev.probability_flow == pf
    w = 0.5 if pf else 1.
    f, g = sde.sde(x, t)
    rf, rg = rev.sde(x, t)
    assert torch.equal(rf, f - g[:, None, None, None] ** 2 * score(x, t) * w)
    assert (rg == 0.) if pf else torch.equal(rg, g)
    for nt in (None, t - 1e-3):          # a 1e-3 gap keeps the VP step variance (t - next_t) beta(t) below 1
        f, G = sde.discretize(x, t, nt)
        rf, rG = rev.discretize(x, t, nt)
        assert torch.equal(rf, f - G[:, None, None, None] ** 2 * score(x, t) * w)
        assert torch.equal(rG, torch.zeros_like(G) if pf else G)
    rf, rG = rev.discretize(x, t, torch.zeros_like(t))
    G0 = sde.sde(x, t)[1] * torch.sqrt(t)
    assert torch.equal(rf, -G0[:, None, None, None] ** 2 * score(x, t) * w)
    assert torch.equal(rG, torch.zeros_like(G0) if pf else G0)


def test_sampler_scalars_are_consistent_with_discretize():
    """x_mean = a x + c score, x = x_mean + d z must reproduce ReverseDiffusionPredictor (sampling.py:205-210) built on
    rsde.discretize: x_mean = x - (f - G^2 score), x = x_mean + G z."""
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(5, 3, 4, 4, generator=gen)
    s = torch.randn(5, 3, 4, 4, generator=gen)
    t = torch.tensor([1.0, 0.7, 0.31, 0.002, 1e-5])
    for sde in (sde_lib.VPSDE(), sde_lib.VESDE(sigma_max=50)):
        a, c, d = sde.reverse_diffusion_coef(t)
        f, G = sde.reverse(lambda *_: s).discretize(x, t)
        want = x - f
        got = a[:, None, None, None] * x + c[:, None, None, None] * s
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
        assert torch.allclose(d, G, rtol=1e-7, atol=0)
    vp = sde_lib.VPSDE()
    assert torch.allclose(vp.score_scale(t), -1.0 / vp.marginal_prob(x, t)[1])
    assert torch.equal(vp.time_cond(t), t * 999) and torch.equal(vp.langevin_alpha(t), vp.alphas[(t * 999).long()])
    ve = sde_lib.VESDE(sigma_max=50)
    assert torch.equal(ve.time_cond(t), ve.marginal_prob(x, t)[1]) and torch.equal(ve.score_scale(t), torch.ones_like(t))


def test_get_sde_and_soft_truncation():
    for name, cls in (('vp/CIFAR10/indm_nll', sde_lib.VPSDE), ('ve/CELEBA/indm', sde_lib.VESDE)):
        cfg = configs.get_config(name)
        sde = sde_lib.get_sde(cfg)
        assert isinstance(sde, cls) and sde.N == cfg.model.num_scales and sde.eps == cfg.training.truncation_time
        assert sde.get_t_min(cfg) == sde.eps
        np.random.seed(3)
        r = np.random.rand()
        np.random.seed(3)
        k = cfg.training.k
        want = sde.eps ** (1. - r) if k == 1.0 else sde.eps / (1. - r * (1 - sde.eps ** (k - 1))) ** (1. / (k - 1))
        assert sde.get_t_min(cfg, st=True) == want
        tt, Z = sde.get_diffusion_time(cfg, 6, 'cpu', 1e-3, importance_sampling=False)
        assert Z == 1 and tt.shape == (6,) and float(tt.min()) >= 1e-3 and float(tt.max()) <= 1.
    assert sde_lib.get_sde(configs.get_config('ve/CELEBA/indm')).sigma_max == 90
    bad = configs.get_config('vp/CIFAR10/indm_nll')
    bad.training.sde = 'subvpsde'
    with pytest.raises(NotImplementedError):
        sde_lib.get_sde(bad)
