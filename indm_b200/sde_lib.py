"""VP / VE SDEs with the reference's interface (sde_lib.py: SDE :7-120, VPSDE :123-215, VESDE :257-350,
get_sde :469-481).  subVP / geometric-VP are out of scope (no BASELINE config uses them).

These classes are the per-sample *scalar* glue (drift / diffusion coefficients, marginal std, discretisation tables,
importance-sampled times).  The tensor-sized work that consumes them on the hot path — state updates, perturbation,
loss reductions — runs in the fused CUDA kernels; `reverse_diffusion_table` / `langevin_table` below precompute the
per-step scalars of the PC sampler into a table those kernels index on the device.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """SDE abstract class (sde_lib.py:7-72)."""

    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        pass

    @abc.abstractmethod
    def sde(self, x, t):
        pass

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        pass

    @abc.abstractmethod
    def prior_sampling(self, shape, data_mean=None):
        pass

    @abc.abstractmethod
    def prior_logp(self, z):
        pass

    def discretize(self, x, t, next_t=None):
        """Euler-Maruyama default (sde_lib.py:54-72)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        f = drift * dt
        G = diffusion * torch.sqrt(torch.tensor(dt, device=t.device))
        return f, G

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE / ODE (sde_lib.py:74-120)."""
        N = self.N
        T = self.T
        sde_fn = self.sde
        discretize_fn = self.discretize

        class RSDE(self.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, t):
                drift, diffusion = sde_fn(x, t)
                score = score_fn(x, t)
                drift = drift - diffusion[:, None, None, None] ** 2 * score * (0.5 if self.probability_flow else 1.)
                diffusion = 0. if self.probability_flow else diffusion
                return drift, diffusion

            def discretize(self, x, t, next_t=None):
                if next_t is None:
                    f, G = discretize_fn(x, t, next_t)
                else:
                    if next_t[0].item() > 0:
                        f, G = discretize_fn(x, t, next_t)
                    else:
                        f = torch.zeros(x.shape, device=x.device)
                        _, G = sde_fn(x, t)
                        G = G * torch.sqrt(t - next_t)
                rev_f = f - G[:, None, None, None] ** 2 * score_fn(x, t) * (0.5 if self.probability_flow else 1.)
                rev_G = torch.zeros_like(G) if self.probability_flow else G
                return rev_f, rev_G

        return RSDE()


class VPSDE(SDE):
    def __init__(self, truncation_time=1e-5, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self.eps = truncation_time
        self.N = N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        drift = -0.5 * beta_t[:, None, None, None] * x
        diffusion = torch.sqrt(beta_t)
        return drift, diffusion

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        mean = torch.exp(log_mean_coeff[:, None, None, None]) * x
        # 1 - exp(-1e-6) at t = 1e-5 is ~17 float32 ulps of 1: a one-ulp difference between the CPU's and the GPU's expf moves
        # std by 3 % (measured: 1.00662e-3 vs 1.03580e-3) and the PF-ODE latent by 6 %.  Evaluate the fp32 argument's exponential
        # in fp64 and round once, i.e. the correctly rounded fp32 exp - what the reference computes when it runs on the host.
        std = torch.sqrt(1. - torch.exp((2. * log_mean_coeff).double()).to(log_mean_coeff.dtype))
        return mean, std

    def prior_sampling(self, shape, data_mean=None):
        if data_mean is None:
            data_mean = 0.
        return torch.randn(*shape) + data_mean

    def prior_logp(self, z):
        shape = z.shape
        N = np.prod(shape[1:])
        return -N / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

    def discretize(self, x, t, next_t=None):
        """DDPM discretization (sde_lib.py:171-184)."""
        if next_t is None:
            timestep = (t * (self.N - 1) / self.T).long()
            beta = self.discrete_betas.to(x.device)[timestep]
            alpha = self.alphas.to(x.device)[timestep]
            sqrt_beta = torch.sqrt(beta)
            f = torch.sqrt(alpha)[:, None, None, None] * x - x
            G = sqrt_beta
        else:
            G = torch.sqrt((t - next_t) * (self.beta_0 + (self.beta_1 - self.beta_0) * t))
            f = torch.sqrt(1. - G ** 2)[:, None, None, None] * x - x
        return f, G

    def integral_beta(self, t):
        return 0.5 * t ** 2 * (self.beta_1 - self.beta_0) + t * self.beta_0

    def antiderivative(self, t, stabilizing_constant=0.):
        if isinstance(t, float) or isinstance(t, int):
            t = torch.tensor(t).float()
        return torch.log(1. - torch.exp(- self.integral_beta(t)) + stabilizing_constant) + self.integral_beta(t)

    def normalizing_constant(self, t_min):
        return self.antiderivative(self.T) - self.antiderivative(t_min)

    def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=None, u=None):
        """`u` (optional): the uniform draw, supplied by parity tests instead of torch.rand"""
        if importance_sampling is None:
            importance_sampling = config.training.importance_sampling
        if importance_sampling:
            Z = self.normalizing_constant(t_min)
            if u is None:
                u = torch.rand(batch_size, device=batch_device)
            return (-self.beta_0 + torch.sqrt(self.beta_0 ** 2 + 2 * (self.beta_1 - self.beta_0) *
                    torch.log(1. + torch.exp(Z * u + self.antiderivative(t_min))))) / (self.beta_1 - self.beta_0), Z.detach()
        return torch.rand(batch_size, device=batch_device) * (self.T - t_min) + t_min, 1

    def get_t_min(self, config, st=False):
        if st:
            if config.training.k == 1.0:
                return self.eps ** (1. - np.random.rand())
            return self.eps / (1. - np.random.rand() * (1 - self.eps ** (config.training.k - 1))) ** (1. / (config.training.k - 1))
        return self.eps

    # ---- scalar tables for the fused sampler kernels -------------------------------------------------------------
    def score_scale(self, t):
        """score = net_out * score_scale(t): -1/std(t) (models/utils.py:171-177, ddpm_score, continuous)."""
        return -1.0 / self.marginal_prob(torch.zeros(1, 1, 1, 1), t)[1]

    def time_cond(self, t):
        return t * 999

    def reverse_diffusion_coef(self, t):
        """(a, c, d) with x_mean = a x + c score, x = x_mean + d z, next_t=None (sde_lib.py:105-118,171-179;
        sampling.py:205-210): f = (sqrt(alpha) - 1) x, G = sqrt(beta)  =>  a = 2 - sqrt(alpha), c = beta, d = sqrt(beta)."""
        ts = (t * (self.N - 1) / self.T).long()
        beta, alpha = self.discrete_betas[ts], self.alphas[ts]
        return 2. - torch.sqrt(alpha), torch.sqrt(beta) ** 2, torch.sqrt(beta)

    def langevin_alpha(self, t):
        """sampling.py:277-279"""
        return self.alphas[(t * (self.N - 1) / self.T).long()]


class VESDE(SDE):
    def __init__(self, truncation_time=1e-5, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(N)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.eps = truncation_time
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(self.sigma_min), np.log(self.sigma_max), N))
        self.N = N

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        drift = torch.zeros_like(x)
        diffusion = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)), device=t.device))
        return drift, diffusion

    def marginal_prob(self, x, t):
        std = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        mean = x
        return mean, std

    def prior_sampling(self, shape, data_mean=None):
        if data_mean is None:
            data_mean = 0.
        return torch.randn(*shape) * self.sigma_max + data_mean

    def prior_logp(self, z):
        shape = z.shape
        N = np.prod(shape[1:])
        return -N / 2. * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)

    def discretize(self, x, t, next_t=None):
        """sde_lib.py:310-323"""
        if next_t is None:
            timestep = (t * (self.N - 1) / self.T).long()
            sigma = self.discrete_sigmas.to(t.device)[timestep]
            adjacent_sigma = torch.where(timestep == 0, torch.zeros_like(t), self.discrete_sigmas[timestep - 1].to(t.device))
            f = torch.zeros_like(x)
            G = torch.sqrt(sigma ** 2 - adjacent_sigma ** 2)
        else:
            _, std_t = self.marginal_prob(x, t)
            _, std_next_t = self.marginal_prob(x, next_t)
            f = torch.zeros_like(x)
            G = torch.sqrt(std_t ** 2 - std_next_t ** 2)
        return f, G

    def antiderivative(self, t):
        if isinstance(t, float) or isinstance(t, int):
            t = torch.tensor(t).float()
        return 2. * torch.log(self.sigma_min * (self.sigma_max / self.sigma_min) ** t)

    def normalizing_constant(self, t_min):
        return self.antiderivative(self.T) - self.antiderivative(t_min)

    def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=None, u=None):
        if importance_sampling is None:
            importance_sampling = config.training.importance_sampling
        if importance_sampling:
            Z = self.normalizing_constant(t_min)
            if u is None:
                u = torch.rand(batch_size, device=batch_device)
            return t_min + ((Z * u) / (2. * (np.log(self.sigma_max) - np.log(self.sigma_min)))), Z.detach()
        return torch.rand(batch_size, device=batch_device) * (self.T - t_min) + t_min, 1

    def get_t_min(self, config, st=False):
        if st:
            if config.training.k == 1.0:
                return self.eps ** (1. - np.random.rand())
            return self.eps / (1. - np.random.rand() * (1 - self.eps ** (config.training.k - 1))) ** (1. / (config.training.k - 1))
        return self.eps

    # ---- scalar tables for the fused sampler kernels -------------------------------------------------------------
    def score_scale(self, t):
        """VE: the network (scale_by_sigma) already returns the score (models/utils.py:182-192)."""
        return torch.ones_like(t)

    def time_cond(self, t):
        return self.marginal_prob(None, t)[1]

    def reverse_diffusion_coef(self, t):
        """f = 0, G = sqrt(sigma_i^2 - sigma_{i-1}^2) (sde_lib.py:311-317)  =>  a = 1, c = G^2, d = G."""
        ts = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas[ts]
        adj = torch.where(ts == 0, torch.zeros_like(t), self.discrete_sigmas[ts - 1])
        G = torch.sqrt(sigma ** 2 - adj ** 2)
        return torch.ones_like(t), G ** 2, G

    def langevin_alpha(self, t):
        return torch.ones_like(t)


def get_sde(config):
    """sde_lib.py:469-481"""
    name = config.training.sde.lower()
    if name == 'vpsde':
        return VPSDE(truncation_time=config.training.truncation_time, beta_min=config.model.beta_min,
                     beta_max=config.model.beta_max, N=config.model.num_scales)
    if name == 'vesde':
        return VESDE(truncation_time=config.training.truncation_time, sigma_min=config.model.sigma_min,
                     sigma_max=config.model.sigma_max, N=config.model.num_scales)
    raise NotImplementedError(f"SDE {config.training.sde} unknown (subVP / gVP are outside the INDM hot-path scope).")
