"""Device-resident RK45 (indm_b200/ode.py) against SciPy's solve_ivp itself — the integrator behind the reference's
likelihood.py:116 and sampling.py:603 — on CPU tensors: same accepted/rejected step sequence, same nfev, same result."""
import numpy as np
import pytest
import torch
from scipy import integrate

from indm_b200.ode import solve_ivp_rk45


def _np_rhs(kind, n):
    rng = np.random.default_rng(3)
    if kind == "linear":                       # stiff-ish linear system with rotation: rejected steps occur
        A = rng.standard_normal((n, n)) / np.sqrt(n) - 1.5 * np.eye(n)
        return lambda t, y: A @ y + np.sin(5 * t)
    if kind == "nonlinear":                    # smooth nonlinear, time dependent (shape of a PF-ODE drift: -0.5 beta(t) (y + s(y,t)))
        W = rng.standard_normal((n, n)) / np.sqrt(n)
        return lambda t, y: -0.5 * (0.1 + 19.9 * t) * (y - np.tanh(W @ y) / np.sqrt(1.0 - np.exp(-0.1 * t - 9.95 * t * t) + 1e-3))
    raise KeyError(kind)


@pytest.mark.parametrize("kind,span,rtol,atol", [
    ("linear", (0.0, 2.0), 1e-5, 1e-5),
    ("linear", (0.0, 2.0), 1e-3, 1e-6),
    ("nonlinear", (1e-5, 1.0), 1e-5, 1e-5),        # the reference's likelihood tolerances and time span
    ("nonlinear", (1.0, 1e-3), 1e-5, 1e-5),        # backward in time: the ODE sampler's direction (sampling.py:603)
])
def test_rk45_matches_scipy_step_for_step(kind, span, rtol, atol):
    n = 96
    f = _np_rhs(kind, n)
    y0 = np.random.default_rng(0).standard_normal(n)
    want = integrate.solve_ivp(f, span, y0, rtol=rtol, atol=atol, method="RK45")
    calls = []

    def f_t(t, y):
        calls.append(t)
        assert y.dtype == torch.float64 and y.dim() == 1
        return torch.from_numpy(f(t, y.numpy()))

    got = solve_ivp_rk45(f_t, span, torch.from_numpy(y0), rtol=rtol, atol=atol)
    assert got.status == 0 and want.status == 0
    assert got.nfev == want.nfev == len(calls)
    assert got.n_steps == len(want.t) - 1
    assert got.t == want.t[-1] == span[1]
    np.testing.assert_allclose(got.y_final.numpy(), want.y[:, -1], rtol=1e-11, atol=1e-13)
    assert got.y.shape == (n, 1)


def test_rk45_float32_rhs_is_promoted_like_scipy():
    """The reference's drift is float32 (likelihood.py:101 `.type(torch.float32)`), stored into SciPy's float64 stages."""
    n = 64
    W = np.random.default_rng(1).standard_normal((n, n)).astype(np.float32) / 8

    def f_np(t, y):
        return (-(y.astype(np.float32)) + np.tanh(W @ y.astype(np.float32)) * np.float32(1 + t)).astype(np.float32)

    y0 = np.random.default_rng(2).standard_normal(n)
    want = integrate.solve_ivp(f_np, (1e-5, 1.0), y0, rtol=1e-5, atol=1e-5, method="RK45")
    got = solve_ivp_rk45(lambda t, y: torch.from_numpy(f_np(t, y.numpy())), (1e-5, 1.0), torch.from_numpy(y0), rtol=1e-5, atol=1e-5)
    assert got.nfev == want.nfev
    # a float32 right-hand side is a step function of its float64 argument at the 6e-8 level: last-bit differences in the stage
    # sums (BLAS gemv vs sequential axpy) flip float32 roundings, so agreement is bounded by float32 eps, far inside rtol = 1e-5
    np.testing.assert_allclose(got.y_final.numpy(), want.y[:, -1], rtol=2e-5, atol=1e-6)


def test_rk45_edge_cases():
    # zero-length interval: SciPy returns immediately with the initial state after the two set-up evaluations' first one
    y0 = torch.ones(4, dtype=torch.float64)
    got = solve_ivp_rk45(lambda t, y: -y, (0.5, 0.5), y0)
    want = integrate.solve_ivp(lambda t, y: -y, (0.5, 0.5), np.ones(4), method="RK45")
    assert got.status == 0 and got.nfev == want.nfev and torch.equal(got.y_final, y0)
    # zero right-hand side: error norm 0 -> MAX_FACTOR growth, few steps
    got = solve_ivp_rk45(lambda t, y: torch.zeros_like(y), (0.0, 1.0), y0, rtol=1e-5, atol=1e-5)
    want = integrate.solve_ivp(lambda t, y: np.zeros_like(y), (0.0, 1.0), np.ones(4), rtol=1e-5, atol=1e-5, method="RK45")
    assert got.nfev == want.nfev and got.n_steps == len(want.t) - 1 and torch.equal(got.y_final, y0)
    # empty state
    got = solve_ivp_rk45(lambda t, y: y, (0.0, 1.0), torch.zeros(0, dtype=torch.float64))
    assert got.status == 0 and got.y_final.numel() == 0


class _ToyScoreNet(torch.nn.Module):
    """A small differentiable stand-in for the score network (plain torch, CPU): the wiring under test is the integrator's, not
    the network's."""

    def __init__(self, ve=False):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = torch.nn.Parameter(torch.randn(3, 3, 3, 3, generator=g) * 0.2)
        self.ve = ve

    def forward(self, x, labels):
        h = torch.tanh(torch.nn.functional.conv2d(x, self.w, padding=1))
        if self.ve:         # labels = sigma(t); output is the score itself: ~ -x / (1 + sigma^2) keeps the VE ODE well posed
            return (0.1 * h - x) / (1. + labels[:, None, None, None] ** 2)
        return h * (1 + labels[:, None, None, None] / 999.) - x          # labels = 999 t; output is the noise prediction


@pytest.mark.parametrize("sde_name", ["vp/CIFAR10/indm_nll", "ve/CIFAR10/indm"])
def test_likelihood_fn_device_integrator_equals_scipy_path(sde_name):
    """likelihood.get_likelihood_fn(method='RK45-device') against method='RK45' (SciPy, the reference's likelihood.py:116) with the
    same right-hand side: same nfe, bpd and latent to float32-drift noise."""
    from indm_b200 import configs, likelihood, sde_lib
    cfg = configs.get_config(sde_name)
    cfg.flow.model = "identity"
    cfg.device = torch.device("cpu")
    sde = sde_lib.get_sde(cfg)
    model = _ToyScoreNet(ve=sde_name.startswith("ve"))
    g = torch.Generator().manual_seed(1)
    data = torch.rand(3, 3, 8, 8, generator=g) * 2 - 1
    eps = (torch.randint(0, 2, data.shape, generator=g).float() * 2 - 1)
    noise = torch.randn(data.shape, generator=g)
    rn = (torch.randn(data.shape, generator=g), torch.randn(data.shape, generator=g))
    out = {}
    for method in ("RK45", "RK45-device"):
        fn = likelihood.get_likelihood_fn(cfg, sde, lambda v: (v + 1.) / 2., method=method)
        out[method] = fn(model, None, data, epsilon=eps, noise=noise, residual_noise=rn)
    (b0, z0, n0), (b1, z1, n1) = out["RK45"], out["RK45-device"]
    assert n0 == n1 and n0 > 0
    assert z1.dtype == torch.float32 and z1.shape == data.shape and b1.shape == (3,)
    assert float((b0 - b1).abs().max()) < 1e-3                       # bits/dim; north star: 0.01
    assert float((z0 - z1).norm() / z0.norm()) < 1e-4


def test_ode_sampler_device_integrator_equals_scipy_path():
    """sampling.get_ode_sampler(method='RK45-device') against the SciPy path (reference sampling.py:596-606)."""
    from indm_b200 import configs, sampling, sde_lib
    cfg = configs.get_config("vp/CIFAR10/indm_fid")
    cfg.flow.model = "identity"
    cfg.device = torch.device("cpu")
    sde = sde_lib.get_sde(cfg)
    model = _ToyScoreNet()
    shape = (2, 3, 8, 8)
    prior = torch.randn(shape, generator=torch.Generator().manual_seed(5))
    res = {}
    for method in ("RK45", "RK45-device"):
        fn = sampling.get_ode_sampler(cfg, sde, shape, lambda v: (v + 1.) / 2., denoise=False, method=method, eps=1e-3, device="cpu")
        res[method] = fn(model, None, prior=prior)
    (x0, _, n0), (x1, _, n1) = res["RK45"], res["RK45-device"]
    assert n0 == n1
    assert float((x0 - x1).norm() / x0.norm()) < 1e-4
