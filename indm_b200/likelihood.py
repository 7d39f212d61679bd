"""Likelihood evaluation with the reference's interface (likelihood.py): `get_div_fn` :27-38, `get_likelihood_fn` :41-140,
`get_elbo_fn` :142-238, `get_likelihood_residual_fn` :241-283.

Execution differences, arithmetic unchanged:
  * the Hutchinson divergence eps^T d(drift)/dx eps comes from the score engine's explicit backward plan
    (`ScoreEngine.vjp`) reached through `torch.autograd.grad`, exactly the call the reference makes;
  * one ODE right-hand side costs 1 network forward + 1 input-VJP: drift and divergence share the forward
    (the reference evaluates the drift twice, likelihood.py:97-98 — 2 forwards + 1 backward);
  * the flow's log-determinant uses the tensor-core VJP chain (flow_models/wolf.py).
The black-box integrator is `scipy.integrate.solve_ivp` (RK45) on the host like the reference, stepping the whole batch
vector with one adaptive step size (likelihood.py:111-116).  `method='RK45-device'` selects the same Dormand-Prince 5(4)
algorithm with the float64 state and stage derivatives resident on the GPU (`indm_b200/ode.py`, pinned step for step
against SciPy): one scalar crosses to the host per step instead of 12 full-state copies.
"""
import numpy as np
import torch
from scipy import integrate

import functools

from . import precision
from .flow_models.flow_model import flow_forward
from .models import utils as mutils
from .ode import solve_ivp_rk45

DEVICE_METHOD = 'RK45-device'


def _likelihood_leg(fn):
    """every network call made by `fn` runs in the precision the policy gives the likelihood leg (indm_b200/precision.py:
    compensated 3xTF32 for the score forward + Hutchinson VJP and the flow log-det, so NLL / NELBO meet 0.01 bpd)"""
    @functools.wraps(fn)
    def run(*a, **k):
        with precision.purpose('likelihood'):
            return fn(*a, **k)
    return run


def get_div_fn(fn):
    """Divergence of `fn` by the Hutchinson-Skilling estimator (likelihood.py:27-38)."""

    def div_fn(x, t, eps):
        with torch.enable_grad():
            x.requires_grad_(True)
            fn_eps = torch.sum(fn(x, t) * eps)
            grad_fn_eps = torch.autograd.grad(fn_eps, x)[0]
        x.requires_grad_(False)
        return torch.sum(grad_fn_eps * eps, dim=tuple(range(1, len(x.shape))))

    return div_fn


def _hutchinson_noise(like, hutchinson_type):
    if hutchinson_type == 'Gaussian':
        return torch.randn_like(like)
    if hutchinson_type == 'Rademacher':
        return torch.randint_like(like, low=0, high=2).float() * 2 - 1.
    raise NotImplementedError(f"Hutchinson type {hutchinson_type} unknown.")


def get_likelihood_fn(config, sde, inverse_scaler, hutchinson_type='Rademacher', rtol=1e-5, atol=1e-5, method='RK45'):
    """likelihood.py:41-140.  Returns `likelihood_fn(model, flow_model, data, logdet=None, residual=True, eps_bpd=1e-5)
    -> (bpd [B], z, nfe)`.  Keyword-only extras pin the random draws for parity tests: `epsilon=` (Hutchinson probe),
    `noise=` (the perturbation at eps_bpd), `residual_noise=` (two tensors for the residual term), `flow_kw=` (passed to
    flow_forward)."""

    def drift_fn(model, x, t):
        """probability-flow drift f - g^2 score / 2 (sde_lib.py:96-103)"""
        score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
        rsde = sde.reverse(score_fn, probability_flow=True)
        return rsde.sde(x, t)[0]

    def drift_and_div(model, x, t, eps):
        """one forward + one input-VJP for both ODE components"""
        with torch.enable_grad():
            x = x.detach().requires_grad_(True)
            drift = drift_fn(model, x, t)
            grad_fn_eps = torch.autograd.grad(torch.sum(drift * eps), x)[0]
        return drift.detach(), torch.sum(grad_fn_eps * eps, dim=tuple(range(1, len(x.shape))))

    def likelihood_fn(model, flow_model, data, logdet=None, residual=True, eps_bpd=1e-5, *, epsilon=None, noise=None,
                      residual_noise=None, flow_kw=None):
        with torch.no_grad():
            score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
            shape = data.shape
            if epsilon is None:
                epsilon = _hutchinson_noise(data, hutchinson_type)

            def ode_func(t, x):
                sample = mutils.from_flattened_numpy(x[:-shape[0]], shape).to(data.device).type(torch.float32)
                vec_t = torch.ones(sample.shape[0], device=sample.device) * t
                drift, logp_grad = drift_and_div(model, sample, vec_t, epsilon)
                return np.concatenate([mutils.to_flattened_numpy(drift), mutils.to_flattened_numpy(logp_grad)], axis=0)

            if config.flow.model != 'identity':
                data, log_jacob = flow_forward(config, flow_model, data, reverse=False, **(flow_kw or {}))
            else:
                log_jacob = torch.zeros(data.shape[0], device=data.device)

            if residual:
                z = torch.randn_like(data) if noise is None else noise
                mean, std = sde.marginal_prob(data, torch.ones(data.shape[0], device=data.device) * eps_bpd)
                start = mean + std[:, None, None, None] * z
            else:
                start = data

            if method == DEVICE_METHOD:
                def ode_func_device(t, y):
                    sample = y[:-shape[0]].reshape(shape).to(torch.float32)
                    vec_t = torch.ones(shape[0], device=sample.device) * t
                    drift, logp_grad = drift_and_div(model, sample, vec_t, epsilon)
                    return torch.cat([drift.reshape(-1), logp_grad.reshape(-1)])

                init_dev = torch.cat([start.reshape(-1).to(torch.float64),
                                      torch.zeros(shape[0], dtype=torch.float64, device=data.device)])
                solution = solve_ivp_rk45(ode_func_device, (eps_bpd, sde.T), init_dev, rtol=rtol, atol=atol)
                nfe = solution.nfev
                z = solution.y_final[:-shape[0]].reshape(shape).to(torch.float32)
                delta_logp = solution.y_final[-shape[0]:].to(torch.float32)
            else:
                init = np.concatenate([mutils.to_flattened_numpy(start), np.zeros((shape[0],))], axis=0)
                solution = integrate.solve_ivp(ode_func, (eps_bpd, sde.T), init, rtol=rtol, atol=atol, method=method)
                nfe = solution.nfev
                zp = solution.y[:, -1]
                z = mutils.from_flattened_numpy(zp[:-shape[0]], shape).to(data.device).type(torch.float32)
                delta_logp = mutils.from_flattened_numpy(zp[-shape[0]:], (shape[0],)).to(data.device).type(torch.float32)
            prior_logp = sde.prior_logp(z)
            if residual:
                residual_fn = get_likelihood_residual_fn(config, sde, score_fn, eps_bpd=eps_bpd)
                delta_logp = delta_logp - residual_fn(data, noise=residual_noise)
            if logdet is None:
                logdet = torch.zeros(data.shape[0], device=data.device)
            assert prior_logp.shape == delta_logp.shape == logdet.shape == log_jacob.shape == torch.Size([data.shape[0]])
            bpd = -(prior_logp + delta_logp + logdet + log_jacob) / np.log(2)
            bpd = bpd / np.prod(shape[1:])
            offset = 7. - inverse_scaler(-1.)          # converts nats of the [-1, 1]-scaled data to bits/dim of 8-bit data
            return bpd + offset, z, nfe

    return _likelihood_leg(likelihood_fn)


def get_elbo_fn(config, sde, inverse_scaler=None, hutchinson_type='Rademacher'):
    """likelihood.py:142-238.  Returns `loss_fn(model, flow_model, batch, logdet=None) -> (nelbo_bpd, nelbo_bpd_residual)`.
    Keyword-only `draws=` is a dict of pre-drawn tensors {u, z, epsilon, lp_z, residual_noise} and `flow_kw=`."""

    @torch.enable_grad()
    def loss_fn(model, flow_model, batch, logdet=None, *, draws=None, flow_kw=None):
        draws = draws or {}
        if config.flow.model != 'identity':
            with torch.no_grad():
                batch, log_jacob = flow_forward(config, flow_model, batch, reverse=False, **(flow_kw or {}))
            log_jacob = log_jacob.squeeze()
        else:
            log_jacob = torch.zeros(batch.shape[0], device=batch.device)
        if logdet is None:
            logdet = torch.zeros(batch.shape[0], device=batch.device)
        score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
        t, Z = sde.get_diffusion_time(config, batch.shape[0], batch.device, sde.eps, importance_sampling=True, u=draws.get('u'))
        qt = 1 / sde.T
        z = draws['z'] if 'z' in draws else torch.randn_like(batch)
        mean, std = sde.marginal_prob(batch, t)
        perturbed_data = (mean + std[:, None, None, None] * z).detach().requires_grad_()
        score = score_fn(perturbed_data, t)
        f, g = sde.sde(perturbed_data, t)
        a = std[:, None, None, None] * score
        mu = (std[:, None, None, None] ** 2) * score - (std[:, None, None, None] ** 2) / (g[:, None, None, None] ** 2) * f
        epsilon = draws['epsilon'] if 'epsilon' in draws else _hutchinson_noise(batch, hutchinson_type)
        Mu = -(torch.autograd.grad(mu, perturbed_data, epsilon, create_graph=False)[0] * epsilon
               ).reshape(batch.size(0), -1).sum(1, keepdim=False) * Z / qt
        Nu = -(a.detach() ** 2).reshape(batch.size(0), -1).sum(1, keepdim=False) * Z / 2 / qt
        with torch.no_grad():
            lp_t = torch.ones_like(t) * sde.T
            lp_z = draws['lp_z'] if 'lp_z' in draws else torch.randn_like(batch)
            lp_mean, lp_std = sde.marginal_prob(batch, lp_t)
            lp = sde.prior_logp(lp_mean + lp_std[:, None, None, None] * lp_z)
            elbos = lp + Mu.detach() + Nu.detach() + log_jacob
            residual_fn = get_likelihood_residual_fn(config, sde, score_fn, eps_bpd=config.training.truncation_time)
            elbos_residual = elbos - residual_fn(batch, noise=draws.get('residual_noise'))
            n = np.prod(list(batch.shape[1:]))
            off = 7. - inverse_scaler(-1.)
            return -(elbos + logdet) / n / np.log(2) + off, -(elbos_residual + logdet) / n / np.log(2) + off

    return _likelihood_leg(loss_fn)


def get_likelihood_residual_fn(config, sde, score_fn, variance='scoreflow', eps_bpd=1e-5):
    """Gaussian-decoder correction between t = 0 and t = eps_bpd (likelihood.py:241-283)."""

    def likelihood_residual_fn(batch, noise=None):
        z1, z2 = noise if noise is not None else (torch.randn_like(batch), torch.randn_like(batch))
        eps_vec = torch.ones((batch.shape[0]), device=batch.device) * config.training.truncation_time
        mean, std = sde.marginal_prob(batch, eps_vec)
        perturbed_data = mean + std[:, None, None, None] * z1
        with torch.no_grad():
            score = score_fn(perturbed_data, eps_vec)
        noise_pred = -std[:, None, None, None] * score

        eps_vec = torch.ones((batch.shape[0]), device=batch.device) * eps_bpd
        mean, std = sde.marginal_prob(batch, eps_vec)
        perturbed_data = mean + std[:, None, None, None] * z2
        alpha, beta = sde.marginal_prob(torch.ones_like(batch), eps_vec)
        q_mean = perturbed_data / alpha - beta[:, None, None, None] * noise_pred / alpha
        if variance == 'ddpm':
            q_std = beta
        elif variance == 'scoreflow':
            q_std = beta / torch.mean(alpha, axis=(1, 2, 3))
        else:
            raise ValueError(variance)
        n_dim = np.prod(batch.shape[1:])
        p_entropy = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(std) + 1.)
        q_recon = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(q_std)) + 0.5 / (q_std ** 2) * torch.square(batch - q_mean).sum(axis=(1, 2, 3))
        assert q_recon.shape == p_entropy.shape == torch.Size([batch.shape[0]])
        return q_recon - p_entropy

    return likelihood_residual_fn
