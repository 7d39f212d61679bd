"""CPU restatement of the reference's VP / VE SDE math (sde_lib.py), torch-CPU FP32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the reference lines it follows.
Pinned by tests/golden/sde.npz (generated from the live reference by tests/golden/make_golden.py).
"""
import numpy as np
import torch


class VP:
    """sde_lib.VPSDE (sde_lib.py:123-215)."""

    def __init__(self, beta_min=0.1, beta_max=20., N=1000, truncation_time=1e-5):
        self.beta_0, self.beta_1, self.N, self.eps, self.T = beta_min, beta_max, N, truncation_time, 1
        # :137-141
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    def sde(self, x, t):  # :147-151
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None, None, None] * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):  # :153-157
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(lmc[:, None, None, None]) * x, torch.sqrt(1. - torch.exp(2. * lmc))

    def prior_logp(self, z):  # :165-169
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

    def discretize(self, x, t, next_t=None):  # :171-184
        if next_t is None:
            ts = (t * (self.N - 1) / self.T).long()
            beta = self.discrete_betas[ts]
            alpha = self.alphas[ts]
            f = torch.sqrt(alpha)[:, None, None, None] * x - x
            G = torch.sqrt(beta)
        else:
            G = torch.sqrt((t - next_t) * (self.beta_0 + (self.beta_1 - self.beta_0) * t))
            f = torch.sqrt(1. - G ** 2)[:, None, None, None] * x - x
        return f, G

    def integral_beta(self, t):  # :186-187
        return 0.5 * t ** 2 * (self.beta_1 - self.beta_0) + t * self.beta_0

    def antiderivative(self, t):  # :189-192
        t = torch.as_tensor(t, dtype=torch.float32)
        return torch.log(1. - torch.exp(-self.integral_beta(t))) + self.integral_beta(t)

    def normalizing_constant(self, t_min):  # :194-195
        return self.antiderivative(self.T) - self.antiderivative(t_min)

    def importance_time(self, u, t_min):  # :197-204 (u ~ U(0,1) passed in)
        Z = self.normalizing_constant(t_min)
        t = (-self.beta_0 + torch.sqrt(self.beta_0 ** 2 + 2 * (self.beta_1 - self.beta_0) *
                                       torch.log(1. + torch.exp(Z * u + self.antiderivative(t_min))))) / (self.beta_1 - self.beta_0)
        return t, Z

    def alpha_for_corrector(self, t):  # sampling.py:277-279
        return self.alphas[(t * (self.N - 1) / self.T).long()]

    def score_labels_and_std(self, t):  # models/utils.py:167-171 (continuous): labels = 999 t, std = marginal std
        return t * 999, self.marginal_prob(torch.zeros(t.shape[0], 1, 1, 1), t)[1]


class VE:
    """sde_lib.VESDE (sde_lib.py:257-350)."""

    def __init__(self, sigma_min=0.01, sigma_max=50., N=1000, truncation_time=1e-5):
        self.sigma_min, self.sigma_max, self.N, self.eps, self.T = sigma_min, sigma_max, N, truncation_time, 1
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))  # :270

    def sde(self, x, t):  # :277-282
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        g = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min))))
        return torch.zeros_like(x), g

    def marginal_prob(self, x, t):  # :284-287
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_logp(self, z):  # :295-298
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)

    def discretize(self, x, t, next_t=None):  # :310-323
        if next_t is None:
            ts = (t * (self.N - 1) / self.T).long()
            sigma = self.discrete_sigmas[ts]
            adj = torch.where(ts == 0, torch.zeros_like(t), self.discrete_sigmas[ts - 1])
            G = torch.sqrt(sigma ** 2 - adj ** 2)
        else:
            G = torch.sqrt(self.marginal_prob(x, t)[1] ** 2 - self.marginal_prob(x, next_t)[1] ** 2)
        return torch.zeros_like(x), G

    def antiderivative(self, t):  # :325-328
        t = torch.as_tensor(t, dtype=torch.float32)
        return 2. * torch.log(self.sigma_min * (self.sigma_max / self.sigma_min) ** t)

    def normalizing_constant(self, t_min):  # :330-331
        return self.antiderivative(self.T) - self.antiderivative(t_min)

    def importance_time(self, u, t_min):  # :333-339
        Z = self.normalizing_constant(t_min)
        return t_min + ((Z * u) / (2. * (np.log(self.sigma_max) - np.log(self.sigma_min)))), Z

    def alpha_for_corrector(self, t):  # sampling.py:280-281
        return torch.ones_like(t)


def get_sde(config):
    """sde_lib.get_sde (sde_lib.py:469-481), VP and VE only (the BASELINE configs)."""
    name = config.training.sde.lower()
    if name == 'vpsde':
        return VP(config.model.beta_min, config.model.beta_max, config.model.num_scales, config.training.truncation_time)
    if name == 'vesde':
        return VE(config.model.sigma_min, config.model.sigma_max, config.model.num_scales, config.training.truncation_time)
    raise NotImplementedError(name)
