"""VP / VE SDEs with the reference's interface (sde_lib.py: SDE :7-120, VPSDE :123-215, VESDE :257-350,
get_sde :469-481).  subVP / geometric-VP are out of scope (no BASELINE config uses them).

These classes are the per-sample *scalar* glue (drift / diffusion coefficients, marginal std, discretisation tables,
importance-sampled times).  The tensor-sized work that consumes them on the hot path — state updates, perturbation,
loss reductions — runs in the fused CUDA kernels; `reverse_diffusion_coef` / `langevin_alpha` / `score_scale` / `time_cond`
below give the per-step scalars of the PC sampler, which `sampling._GraphedPC` tabulates for those kernels to index on the device.

Layout of this module: one `_ReverseProcess` serves both SDEs (the reference builds a nested subclass per call); what VP and VE
share — the importance-sampling normaliser, the truncation-time draw, the timestep index — lives once in `SDE`.  Every formula
keeps the reference's operation order, so float32 results are bit-identical to the reference's on the same device
(tests/test_sde_cpu.py pins them to the golden vectors taken from the live reference).
"""
import abc

import numpy as np
import torch


def _bc(v):
    """[B] -> [B, 1, 1, 1]: per-sample scalars against NCHW tensors"""
    return v[:, None, None, None]


def _as_tensor(t):
    return torch.tensor(t).float() if isinstance(t, (float, int)) else t


class _ReverseProcess:
    """What `SDE.reverse(score_fn, probability_flow)` returns (sde_lib.py:74-120): drift / diffusion of the reverse-time SDE, or of
    the probability-flow ODE (half the score term, no diffusion), in continuous form (`sde`) and per discretisation step
    (`discretize`)."""

    def __init__(self, forward, score_fn, probability_flow):
        self._fwd, self._score, self.probability_flow = forward, score_fn, probability_flow
        self.N = forward.N

    @property
    def T(self):
        return self._fwd.T

    def _weight(self):
        return 0.5 if self.probability_flow else 1.

    def sde(self, x, t):
        f, g = self._fwd.sde(x, t)
        rev = f - _bc(g) ** 2 * self._score(x, t) * self._weight()
        return rev, (0. if self.probability_flow else g)

    def discretize(self, x, t, next_t=None):
        if next_t is None or next_t[0].item() > 0:
            f, G = self._fwd.discretize(x, t, next_t)
        else:
            # last step onto t = 0 (sde_lib.py:110-113): no drift, the diffusion coefficient integrated over the gap
            f = torch.zeros(x.shape, device=x.device)
            G = self._fwd.sde(x, t)[1] * torch.sqrt(t - next_t)
        rev_f = f - _bc(G) ** 2 * self._score(x, t) * self._weight()
        return rev_f, (torch.zeros_like(G) if self.probability_flow else G)


class SDE(abc.ABC):
    """SDE abstract class (sde_lib.py:7-72) plus what the two concrete SDEs share."""

    def __init__(self, N, truncation_time):
        super().__init__()
        self.N = N
        self.eps = truncation_time

    @property
    def T(self):
        return 1

    @abc.abstractmethod
    def sde(self, x, t):
        """(drift [B,C,H,W], diffusion [B])"""

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        """(mean, std [B]) of p_t(x_t | x_0 = x)"""

    @abc.abstractmethod
    def prior_sampling(self, shape, data_mean=None):
        pass

    @abc.abstractmethod
    def prior_logp(self, z):
        pass

    @abc.abstractmethod
    def antiderivative(self, t):
        pass

    def discretize(self, x, t, next_t=None):
        """Euler-Maruyama default (sde_lib.py:54-72)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE / ODE (sde_lib.py:74-120)."""
        return _ReverseProcess(self, score_fn, probability_flow)

    def _index(self, t):
        """position of t in the N-entry discretisation tables (sde_lib.py:173,312)"""
        return (t * (self.N - 1) / self.T).long()

    def normalizing_constant(self, t_min):
        """Z of the importance-sampling density over [t_min, T] (sde_lib.py:194-195, 330-331)"""
        return self.antiderivative(self.T) - self.antiderivative(t_min)

    def _uniform_time(self, batch_size, batch_device, t_min):
        return torch.rand(batch_size, device=batch_device) * (self.T - t_min) + t_min, 1

    def get_t_min(self, config, st=False):
        """soft truncation (sde_lib.py:208-215, 343-350): a random smallest diffusion time, numpy's global generator"""
        if not st:
            return self.eps
        k = config.training.k
        if k == 1.0:
            return self.eps ** (1. - np.random.rand())
        return self.eps / (1. - np.random.rand() * (1 - self.eps ** (k - 1))) ** (1. / (k - 1))


class VPSDE(SDE):
    """dx = -beta(t) x / 2 dt + sqrt(beta(t)) dw, beta linear in t (sde_lib.py:123-215)."""

    def __init__(self, truncation_time=1e-5, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N, truncation_time)
        self.beta_0, self.beta_1 = beta_min, beta_max
        # DDPM tables of the N-step discretisation
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    def _beta(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def sde(self, x, t):
        b = self._beta(t)
        return -0.5 * _bc(b) * x, torch.sqrt(b)

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        # 1 - exp(-1e-6) at t = 1e-5 is ~17 float32 ulps of 1: a one-ulp difference between the CPU's and the GPU's expf moves
        # std by 3 % (measured: 1.00662e-3 vs 1.03580e-3) and the PF-ODE latent by 6 %.  Evaluate the fp32 argument's exponential
        # in fp64 and round once, i.e. the correctly rounded fp32 exp - what the reference computes when it runs on the host.
        decay = torch.exp((2. * log_mean_coeff).double()).to(log_mean_coeff.dtype)
        return torch.exp(_bc(log_mean_coeff)) * x, torch.sqrt(1. - decay)

    def prior_sampling(self, shape, data_mean=None):
        return torch.randn(*shape) + (0. if data_mean is None else data_mean)

    def prior_logp(self, z):
        D = np.prod(z.shape[1:])
        return -D / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

    def discretize(self, x, t, next_t=None):
        """DDPM discretization (sde_lib.py:171-184); with `next_t` the step covers the continuous gap t - next_t."""
        if next_t is None:
            i = self._index(t)
            G = torch.sqrt(self.discrete_betas.to(x.device)[i])
            keep = torch.sqrt(self.alphas.to(x.device)[i])
        else:
            G = torch.sqrt((t - next_t) * (self.beta_0 + (self.beta_1 - self.beta_0) * t))
            keep = torch.sqrt(1. - G ** 2)
        return _bc(keep) * x - x, G

    def integral_beta(self, t):
        return 0.5 * t ** 2 * (self.beta_1 - self.beta_0) + t * self.beta_0

    def antiderivative(self, t, stabilizing_constant=0.):
        t = _as_tensor(t)
        B = self.integral_beta(t)
        return torch.log(1. - torch.exp(- B) + stabilizing_constant) + B

    def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=None, u=None):
        """sde_lib.py:197-206.  `u` (optional): the uniform draw, supplied by parity tests instead of torch.rand"""
        if importance_sampling is None:
            importance_sampling = config.training.importance_sampling
        if not importance_sampling:
            return self._uniform_time(batch_size, batch_device, t_min)
        Z = self.normalizing_constant(t_min)
        if u is None:
            u = torch.rand(batch_size, device=batch_device)
        slope = self.beta_1 - self.beta_0
        # inverse CDF of the likelihood-weighting density g^2 / std^2 on [t_min, T]
        root = torch.sqrt(self.beta_0 ** 2 + 2 * slope * torch.log(1. + torch.exp(Z * u + self.antiderivative(t_min))))
        return (-self.beta_0 + root) / slope, Z.detach()

    # ---- per-step scalars of the fused sampler kernels -----------------------------------------------------------
    def score_scale(self, t):
        """score = net_out * score_scale(t): -1/std(t) (models/utils.py:171-177, ddpm_score, continuous)."""
        return -1.0 / self.marginal_prob(torch.zeros(1, 1, 1, 1), t)[1]

    def time_cond(self, t):
        return t * 999

    def reverse_diffusion_coef(self, t):
        """(a, c, d) with x_mean = a x + c score, x = x_mean + d z, next_t=None (sde_lib.py:105-118,171-179;
        sampling.py:205-210): f = (sqrt(alpha) - 1) x, G = sqrt(beta)  =>  a = 2 - sqrt(alpha), c = beta, d = sqrt(beta)."""
        i = self._index(t)
        G = torch.sqrt(self.discrete_betas[i])
        return 2. - torch.sqrt(self.alphas[i]), G ** 2, G

    def langevin_alpha(self, t):
        """sampling.py:277-279"""
        return self.alphas[self._index(t)]


class VESDE(SDE):
    """dx = sigma(t) sqrt(2 log(sigma_max / sigma_min)) dw, sigma geometric in t (sde_lib.py:257-350)."""

    def __init__(self, truncation_time=1e-5, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(N, truncation_time)
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(self.sigma_min), np.log(self.sigma_max), N))

    def _sigma(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def _log_ratio(self):
        return np.log(self.sigma_max) - np.log(self.sigma_min)

    def sde(self, x, t):
        g = self._sigma(t) * torch.sqrt(torch.tensor(2 * self._log_ratio(), device=t.device))
        return torch.zeros_like(x), g

    def marginal_prob(self, x, t):
        return x, self._sigma(t)

    def prior_sampling(self, shape, data_mean=None):
        return torch.randn(*shape) * self.sigma_max + (0. if data_mean is None else data_mean)

    def prior_logp(self, z):
        D = np.prod(z.shape[1:])
        return -D / 2. * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)

    def _step_std(self, t, device=None):
        """sqrt(sigma_i^2 - sigma_{i-1}^2) of the table step that contains t, sigma_{-1} = 0 (sde_lib.py:311-317)"""
        i = self._index(t)
        table = self.discrete_sigmas if device is None else self.discrete_sigmas.to(device)
        below = table[i - 1] if device is None else self.discrete_sigmas[i - 1].to(device)
        prev = torch.where(i == 0, torch.zeros_like(t), below)
        return torch.sqrt(table[i] ** 2 - prev ** 2)

    def discretize(self, x, t, next_t=None):
        """sde_lib.py:310-323"""
        if next_t is None:
            G = self._step_std(t, t.device)
        else:
            G = torch.sqrt(self._sigma(t) ** 2 - self._sigma(next_t) ** 2)
        return torch.zeros_like(x), G

    def antiderivative(self, t):
        return 2. * torch.log(self._sigma(_as_tensor(t)))

    def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=None, u=None):
        """sde_lib.py:333-341"""
        if importance_sampling is None:
            importance_sampling = config.training.importance_sampling
        if not importance_sampling:
            return self._uniform_time(batch_size, batch_device, t_min)
        Z = self.normalizing_constant(t_min)
        if u is None:
            u = torch.rand(batch_size, device=batch_device)
        return t_min + ((Z * u) / (2. * self._log_ratio())), Z.detach()

    # ---- per-step scalars of the fused sampler kernels -----------------------------------------------------------
    def score_scale(self, t):
        """VE: the network (scale_by_sigma) already returns the score (models/utils.py:182-192)."""
        return torch.ones_like(t)

    def time_cond(self, t):
        return self._sigma(t)

    def reverse_diffusion_coef(self, t):
        """f = 0, G = sqrt(sigma_i^2 - sigma_{i-1}^2) (sde_lib.py:311-317)  =>  a = 1, c = G^2, d = G."""
        G = self._step_std(t)
        return torch.ones_like(t), G ** 2, G

    def langevin_alpha(self, t):
        return torch.ones_like(t)


def get_sde(config):
    """sde_lib.py:469-481"""
    kind = config.training.sde.lower()
    common = dict(truncation_time=config.training.truncation_time, N=config.model.num_scales)
    if kind == 'vpsde':
        return VPSDE(beta_min=config.model.beta_min, beta_max=config.model.beta_max, **common)
    if kind == 'vesde':
        return VESDE(sigma_min=config.model.sigma_min, sigma_max=config.model.sigma_max, **common)
    raise NotImplementedError(f"SDE {config.training.sde} unknown (subVP / gVP are outside the INDM hot-path scope).")
