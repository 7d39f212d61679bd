"""GPU: the remaining sampling components of sampling.py on the fused update kernels — Euler-Maruyama and ancestral predictors,
the annealed-Langevin corrector, `pc_sampler_search` (sampling.pc_denoise, the README's VE evaluation path) and the black-box ODE
sampler — against trajectories of the live reference with replayed noise (tests/golden/samplers_tiny.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, tiny, rel_l2  # noqa: E402
from indm_b200 import configs, sde_lib, sampling  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from oracle import ncsnpp as oncsnpp  # noqa: E402

SPECS = {
    'em_vp': ('vp/CIFAR10/indm_fid', dict(method='pc', predictor='euler_maruyama', corrector='none', num_scales=6)),
    'ald_ve': ('ve/CIFAR10/indm', dict(method='pc', predictor='reverse_diffusion', corrector='ald', num_scales=6)),
    'search_ve': ('ve/CIFAR10/indm', dict(method='pc', predictor='reverse_diffusion', corrector='langevin', pc_denoise=True)),
    'ode_vp': ('vp/CIFAR10/indm_fid', dict(method='ode')),
}


def _setup(tag, mode):
    base, over = SPECS[tag]
    cfg = configs.get_config(base)
    tiny(cfg)
    cfg.flow.model = 'identity'
    for k, v in over.items():
        setattr(cfg.sampling, k, v)
    if tag == 'search_ve':
        cfg.model.num_scales = 8
    if tag == 'ode_vp':
        cfg.eval.rtol = cfg.eval.atol = 1e-3
    cfg.device = torch.device('cuda:0')
    model = mutils.create_model(cfg)
    model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, 11).items()})
    model.eval()
    model.module.compute_mode = mode
    return cfg, model, sde_lib.get_sde(cfg)


@pytest.mark.parametrize("mode,tol", [('tf32', 1e-3), ('bf16', 3e-2)])
@pytest.mark.parametrize("tag", ['em_vp', 'ald_ve', 'search_ve'])
def test_pc_variants_match_reference(tag, mode, tol):
    g = load_npz('samplers_tiny.npz')
    cfg, model, sde = _setup(tag, mode)
    prior = torch.from_numpy(g[f'{tag}_prior']) * (cfg.model.sigma_max if tag.endswith('ve') else 1.0)
    fn = sampling.get_sampling_fn(cfg, sde, tuple(prior.shape), lambda v: v, cfg.sampling.truncation_time)
    before, after, nfe = fn(model, None, prior=prior, noise=[torch.from_numpy(n) for n in g[f'{tag}_noises']])
    torch.cuda.synchronize()
    err = rel_l2(before.cpu().numpy(), g[f'{tag}_out'])
    print(f'{tag} {mode}: rel-L2 {err:.3e}')
    assert nfe == int(g[f'{tag}_nfe'])
    assert err < tol


@pytest.mark.parametrize("mode,tol", [('tf32', 2e-2), ('bf16', 0.3)])
def test_ode_sampler_matches_reference(mode, tol):
    g = load_npz('samplers_tiny.npz')
    cfg, model, sde = _setup('ode_vp', mode)
    prior = torch.from_numpy(g['ode_vp_prior'])
    fn = sampling.get_sampling_fn(cfg, sde, tuple(prior.shape), lambda v: v, cfg.sampling.truncation_time)
    before, after, nfe = fn(model, None, prior=prior)
    torch.cuda.synchronize()
    err = rel_l2(before.cpu().numpy(), g['ode_vp_out'])
    print(f'ode {mode}: rel-L2 {err:.3e}, nfe {nfe} (reference {int(g["ode_vp_nfe"])})')
    assert err < tol          # an adaptive RK45 trajectory through a random-weight network amplifies rounding differences


@pytest.mark.parametrize("base", ['vp/CIFAR10/indm_fid', 've/CIFAR10/indm'])
def test_ancestral_predictor_follows_the_reference_formulas(base):
    """AncestralSamplingPredictor cannot run inside the reference's own pc_sampler (its update_fn lacks the next_t argument the
    sampler passes, sampling.py:245 vs :351), so it is checked against the formulas of sampling.py:224-243 directly."""
    cfg = configs.get_config(base)
    tiny(cfg)
    cfg.device = torch.device('cuda:0')
    model = mutils.create_model(cfg)
    model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, 11).items()})
    model.eval()
    model.module.compute_mode = 'tf32'
    sde = sde_lib.get_sde(cfg)
    score_fn = mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 16, 16, generator=gen).cuda()
    z = torch.randn(2, 3, 16, 16, generator=gen).cuda()
    t = torch.full((2,), 0.37, device='cuda')
    with torch.no_grad():
        got, got_mean = sampling.AncestralSamplingPredictor(sde, score_fn).update_fn(x, t, noise=z)
        s = score_fn(x, t)
        ts = (t * (sde.N - 1) / sde.T).long().cpu()
        if isinstance(sde, sde_lib.VPSDE):
            beta = sde.discrete_betas[ts].cuda()[:, None, None, None]
            want_mean = (x + beta * s) / torch.sqrt(1. - beta)
            want = want_mean + torch.sqrt(beta) * z
        else:
            sigma, adj = sde.discrete_sigmas[ts].cuda()[:, None, None, None], sde.discrete_sigmas[ts - 1].cuda()[:, None, None, None]
            want_mean = x + s * (sigma ** 2 - adj ** 2)
            want = want_mean + torch.sqrt(adj ** 2 * (sigma ** 2 - adj ** 2) / sigma ** 2) * z
    assert rel_l2(got_mean.cpu().numpy(), want_mean.cpu().numpy()) < 1e-5
    assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("tag", ['em_vp', 'search_ve'])
def test_plain_forward_after_sampling_is_conditioned_on_t(tag):
    """The sampler's schedule-reading time embedding must not leak into ordinary `model(x, t)` calls on the same cached engine
    (same batch size): a forward after `pc_sampler` equals the forward of a model that never sampled."""
    cfg, model, sde = _setup(tag, 'tf32')
    cfg2, fresh, _ = _setup(tag, 'tf32')
    g = load_npz('samplers_tiny.npz')
    prior = torch.from_numpy(g[f'{tag}_prior']) * (cfg.model.sigma_max if tag.endswith('ve') else 1.0)
    fn = sampling.get_sampling_fn(cfg, sde, tuple(prior.shape), lambda v: v, cfg.sampling.truncation_time)
    fn(model, None, prior=prior)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(tuple(prior.shape), generator=gen).cuda()
    t = (torch.rand(prior.shape[0], generator=gen) * 0.8 + 0.1).cuda()
    with torch.no_grad():
        a = mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)(x, t).cpu().numpy()
        b = mutils.get_score_fn(cfg2, sde, fresh, train=False, continuous=True)(x, t).cpu().numpy()
    assert np.isfinite(a).all()
    assert rel_l2(a, b) < 1e-5          # fp32 atomics in the fused GroupNorm statistics reorder roundings from run to run


def test_unseeded_sampler_calls_use_fresh_noise_and_manual_seed_controls_it():
    """Reference-signature callers pass no seed (sampling_lib.get_samples): every call must draw its own noise path, yet
    torch.manual_seed must make a run reproducible; an explicit seed= pins the path."""
    cfg, model, sde = _setup('em_vp', 'bf16')
    shape = (2, 3, 16, 16)
    fn = sampling.get_sampling_fn(cfg, sde, shape, lambda v: v, cfg.sampling.truncation_time)
    prior = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(7)
    a1 = fn(model, None, prior=prior)[0].cpu().numpy()
    a2 = fn(model, None, prior=prior)[0].cpu().numpy()
    torch.manual_seed(7)
    b1 = fn(model, None, prior=prior)[0].cpu().numpy()
    assert rel_l2(a2, a1) > 1e-2                  # two rounds: different Brownian paths
    assert rel_l2(a1, b1) < 1e-3                  # same global seed: same path (not bitwise: fp32 atomics in the GroupNorm statistics)
    s1 = fn(model, None, prior=prior, seed=5)[0].cpu().numpy()
    s2 = fn(model, None, prior=prior, seed=5)[0].cpu().numpy()
    assert rel_l2(s1, s2) < 1e-3 and rel_l2(s1, a1) > 1e-2
