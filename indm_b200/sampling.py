"""Predictor-corrector sampling with the reference's interface (sampling.py: registries :36-83, get_sampling_fn
:86-133, Predictor/Corrector :136-182, ReverseDiffusionPredictor :200-210, LangevinCorrector :263-292,
NoneCorrector/NonePredictor :252-260,:332-340, get_pc_sampler :365-545 → pc_sampler :410-456).

B200-native execution of `pc_sampler`: one CUDA graph holds a whole sampling step — time embedding read from a device
schedule table, the score network (ScoreEngine launch plan), the fused predictor (and Langevin corrector) update with
in-register Philox noise, and the device step counter increment — and is replayed `num_scales` times with no host
work or host<->device traffic in the loop.  Noise can alternatively be supplied (`noise=` iterator) for bit-comparable
trajectories against a recorded reference run (SURVEY.md §7 hard part 3).
"""
import abc
import io
import os

import numpy as np
import torch

from . import _lib as L
from . import parallel
from . import sde_lib
from .models import utils as mutils

_CORRECTORS = {}
_PREDICTORS = {}


def register_predictor(cls=None, *, name=None):
    """A decorator for registering predictor classes (sampling.py:40-57)."""

    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _PREDICTORS:
            raise ValueError(f'Already registered model with name: {local_name}')
        _PREDICTORS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def register_corrector(cls=None, *, name=None):
    """A decorator for registering corrector classes (sampling.py:60-77)."""

    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _CORRECTORS:
            raise ValueError(f'Already registered model with name: {local_name}')
        _CORRECTORS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


def get_sampling_fn(config, sde, shape, inverse_scaler, eps):
    """sampling.py:86-133"""
    sampler_name = config.sampling.method
    if sampler_name.lower() == 'pc':
        predictor = get_predictor(config.sampling.predictor.lower())
        corrector = get_corrector(config.sampling.corrector.lower())
        return get_pc_sampler(config=config, sde=sde, shape=shape, predictor=predictor, corrector=corrector,
                              inverse_scaler=inverse_scaler, snr=config.sampling.snr, n_steps=config.sampling.n_steps_each,
                              probability_flow=config.sampling.probability_flow, continuous=config.training.continuous,
                              denoise=config.sampling.noise_removal, eps=eps, device=config.device)
    if sampler_name.lower() == 'ode':
        return get_ode_sampler(config=config, sde=sde, shape=shape, inverse_scaler=inverse_scaler, denoise=config.sampling.noise_removal,
                               eps=eps, rtol=config.eval.rtol, atol=config.eval.atol, device=config.device)
    raise ValueError(f"Sampler name {sampler_name} unknown.")


class Predictor(abc.ABC):
    """sampling.py:136-158"""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t, next_t=None):
        pass


class Corrector(abc.ABC):
    """sampling.py:161-182"""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.sde = sde
        self.score_fn = score_fn
        self.snr = snr
        self.n_steps = n_steps

    @abc.abstractmethod
    def update_fn(self, x, t):
        pass


def _coef_table(rows, device):
    return torch.tensor(rows, dtype=torch.float32, device=device)


def predictor_coefficients(sde, name, t, next_t=None, probability_flow=False):
    """Per-step scalars (a, c, d) of a predictor written as  x_mean = a x + c score,  x = x_mean + d z  — the form the fused update
    kernel evaluates.  Every INDM predictor has this form because VP / VE drifts are linear in x.  Derived from the SDE's own
    methods on a unit probe, so the discretisation rules live in one place (sde_lib).  t, next_t: 1-D CPU tensors.
      reverse_diffusion  (sampling.py:200-210, sde_lib.py:105-118):  a = 1 - phi, c = G^2 (1/2 if PF), d = G (0 if PF)
      euler_maruyama     (sampling.py:186-197, sde_lib.py:96-103):   dt = -1/N
      ancestral_sampling (sampling.py:213-249)"""
    t = t.detach().float().cpu().reshape(-1)
    one = torch.ones((t.shape[0], 1, 1, 1))
    pfac = 0.5 if probability_flow else 1.0
    if name == 'reverse_diffusion':
        if next_t is None:
            f, G = sde.discretize(one, t, None)
        else:
            nt = next_t.detach().float().cpu().reshape(-1)
            pos = nt > 0
            f1, G1 = sde.discretize(one, t, torch.where(pos, nt, t * 0.5))
            G0 = sde.sde(one, t)[1] * torch.sqrt(t - nt)             # RSDE.discretize with next_t == 0 (sde_lib.py:109-113)
            f = torch.where(pos[:, None, None, None], f1, torch.zeros_like(f1))
            G = torch.where(pos, G1, G0)
        phi = f.reshape(-1)
        a, c, d = 1.0 - phi, G ** 2 * pfac, G.clone()
    elif name == 'euler_maruyama':
        drift, g = sde.sde(one, t)
        dt = -1.0 / sde.N
        a, c, d = 1.0 + drift.reshape(-1) * dt, -(g ** 2) * pfac * dt, g * float(np.sqrt(-dt))
    elif name == 'ancestral_sampling':
        if probability_flow:
            raise AssertionError('Probability flow not supported by ancestral sampling')
        ts = (t * (sde.N - 1) / sde.T).long()
        if isinstance(sde, sde_lib.VPSDE):
            beta = sde.discrete_betas[ts]
            a, c, d = 1.0 / torch.sqrt(1. - beta), beta / torch.sqrt(1. - beta), torch.sqrt(beta)
        elif isinstance(sde, sde_lib.VESDE):
            sigma = sde.discrete_sigmas[ts]
            adj = torch.where(ts == 0, torch.zeros_like(t), sde.discrete_sigmas[torch.clamp(ts - 1, min=0)])
            a, c = torch.ones_like(t), sigma ** 2 - adj ** 2
            d = torch.sqrt(adj ** 2 * (sigma ** 2 - adj ** 2) / sigma ** 2)
        else:
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
    elif name == 'none':
        a, c, d = torch.ones_like(t), torch.zeros_like(t), torch.zeros_like(t)
    else:
        raise ValueError(f'unknown predictor {name!r}')
    if probability_flow:
        d = torch.zeros_like(d)
    return a.float(), c.float(), d.float()


def _fused_predictor_step(x, score, coef3, noise, seed, offset):
    coef = _coef_table([[float(coef3[0]), float(coef3[1]), float(coef3[2]), 0.0]], x.device)
    x = x.contiguous().clone()
    x_mean = torch.empty_like(x)
    N, D = x.shape[0], x[0].numel()
    L.call('indm_pc_predictor_update', L.ptr(x), L.ptr(score), L.ptr(noise.contiguous()) if noise is not None else None,
           L.ptr(x_mean), L.ptr(coef), 4, None, N, D, seed, None, offset)
    return x, x_mean


@register_predictor(name='reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
    """x_mean = x - rev_f, x = x_mean + G z (sampling.py:200-210) — one fused kernel launch."""

    name = 'reverse_diffusion'

    def update_fn(self, x, t, next_t=None, noise=None, seed=0, offset=0):
        a, c, d = predictor_coefficients(self.sde, 'reverse_diffusion', t, next_t, self.probability_flow)
        if not (bool((a == a[0]).all()) and bool((c == c[0]).all())):
            raise NotImplementedError('per-sample time steps in one batch')
        return _fused_predictor_step(x, self.score_fn(x, t).contiguous(), (a[0], c[0], d[0]), noise, seed, offset)


@register_predictor(name='euler_maruyama')
class EulerMaruyamaPredictor(Predictor):
    """sampling.py:186-197"""
    name = 'euler_maruyama'

    def update_fn(self, x, t, next_t=None, noise=None, seed=0, offset=0):
        a, c, d = predictor_coefficients(self.sde, 'euler_maruyama', t, None, self.probability_flow)
        if not (bool((a == a[0]).all()) and bool((c == c[0]).all())):
            raise NotImplementedError('per-sample time steps in one batch')
        return _fused_predictor_step(x, self.score_fn(x, t).contiguous(), (a[0], c[0], d[0]), noise, seed, offset)


@register_predictor(name='ancestral_sampling')
class AncestralSamplingPredictor(Predictor):
    """sampling.py:213-249 (VE / VP)"""
    name = 'ancestral_sampling'

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        assert not probability_flow, "Probability flow not supported by ancestral sampling"

    def update_fn(self, x, t, next_t=None, noise=None, seed=0, offset=0):
        a, c, d = predictor_coefficients(self.sde, 'ancestral_sampling', t)
        if not (bool((a == a[0]).all()) and bool((c == c[0]).all())):
            raise NotImplementedError('per-sample time steps in one batch')
        return _fused_predictor_step(x, self.score_fn(x, t).contiguous(), (a[0], c[0], d[0]), noise, seed, offset)


@register_predictor(name='none')
class NonePredictor(Predictor):
    def __init__(self, sde, score_fn, probability_flow=False):
        pass

    def update_fn(self, x, t, next_t=None):
        return x, x


@register_corrector(name='langevin')
class LangevinCorrector(Corrector):
    """sampling.py:263-292: batch-mean gradient / noise norms -> step size -> update; 2 fused launches per inner step."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    def update_fn(self, x, t, noise=None, seed=0, offset=1):
        alpha = float(self.sde.langevin_alpha(t.detach().float().cpu())[0])
        coef = _coef_table([[alpha, float(self.snr)]], x.device)
        x = x.contiguous().clone()
        x_mean = torch.empty_like(x)
        N, D = x.shape[0], x[0].numel()
        norms = torch.empty((N, 2), device=x.device)
        for i in range(self.n_steps):
            grad = self.score_fn(x, t).contiguous()
            z = noise.contiguous() if noise is not None else None
            L.call('indm_langevin_norms', L.ptr(grad), L.ptr(z), L.ptr(norms), None, N, D, seed, None, offset + i)
            L.call('indm_langevin_update', L.ptr(x), L.ptr(grad), L.ptr(z), L.ptr(x_mean), L.ptr(norms), L.ptr(coef), 2, None,
                   N, D, seed, None, offset + i)
        return x, x_mean


@register_corrector(name='ald')
class AnnealedLangevinDynamics(Corrector):
    """sampling.py:295-329: step = (snr std(t))^2 2 alpha;  x_mean = x + step s;  x = x_mean + sqrt(2 step) z"""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

    @staticmethod
    def coefficients(sde, snr, t):
        t = t.detach().float().cpu().reshape(-1)
        std = sde.marginal_prob(torch.zeros((t.shape[0], 1, 1, 1)), t)[1]
        step = (snr * std) ** 2 * 2 * sde.langevin_alpha(t)
        return torch.ones_like(step), step, torch.sqrt(step * 2)

    def update_fn(self, x, t, noise=None, seed=0, offset=1):
        a, c, d = self.coefficients(self.sde, self.snr, t)
        x_mean = x
        for i in range(self.n_steps):
            x, x_mean = _fused_predictor_step(x, self.score_fn(x, t).contiguous(), (a[0], c[0], d[0]), noise, seed, offset + i)
        return x, x_mean


@register_corrector(name='none')
class NoneCorrector(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, x, t):
        return x, x


def shared_predictor_update_fn(x, t, next_t, sde, model, predictor, probability_flow, continuous, config):
    """sampling.py:343-351"""
    score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=continuous)
    predictor_obj = NonePredictor(sde, score_fn, probability_flow) if predictor is None else predictor(sde, score_fn, probability_flow)
    return predictor_obj.update_fn(x, t, next_t)


def shared_corrector_update_fn(x, t, sde, model, corrector, continuous, snr, n_steps, config):
    """sampling.py:354-362"""
    score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=continuous)
    corrector_obj = NoneCorrector(sde, score_fn, snr, n_steps) if corrector is None else corrector(sde, score_fn, snr, n_steps)
    return corrector_obj.update_fn(x, t)


class _GraphedPC:
    """One sampling step = one CUDA graph (built once per (model, batch, corrector) and cached on the model)."""

    LD = 12   # schedule row: [time_cond, out_scale, a, c, d, alpha, snr, 0, ca, cc, cd, 0]

    def __init__(self, config, sde, net, batch, corrector_name, n_steps, probability_flow, seed, predictor_name='reverse_diffusion'):
        self.sde, self.net, self.N = sde, net, batch
        self.eng = net.engine(batch, infer=True)      # forward-only plan: padded-pixel operands on the small feature maps
        self.dev = self.eng.dev
        self.langevin = corrector_name == 'langevin'
        self.ald = corrector_name == 'ald'
        self.predictor_name = predictor_name
        self.n_steps = n_steps
        self.pf = probability_flow
        self.seed = seed
        eng = self.eng
        D = eng.ch * eng.S * eng.S
        self.D = D
        self.x = eng.x_in                                # the sampler state IS the network input buffer (no copies)
        self.x_mean = torch.zeros_like(self.x)
        self.norms = torch.zeros((batch, 2), device=self.dev)
        # sampling.global_langevin_norms: batch-sharded sampling with the Langevin step size taken from the GLOBAL batch means of
        # the gradient / noise norms (sampling.py:286-288 takes .mean() over the whole batch): three floats all-reduced per
        # corrector step on the sampling stream (NCCL; inside the captured graph).  Off = per-rank statistics, no communication.
        self.global_norms = bool(getattr(config.sampling, 'global_langevin_norms', False)) and \
            torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        self.gsums = torch.zeros((3,), device=self.dev)
        self.step = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.seed_dev = torch.zeros((1,), dtype=torch.int64, device=self.dev)   # read by the kernels: graph-replay safe
        self.sched = None
        self.graph = None

    def set_schedule(self, timesteps, snr_per_step, next_timesteps=None):
        """per-step scalars of the sampling loop (sampling.py:425-435, 470-478) as one device table, one row per step:
        the network's conditioning value and output scale, the predictor's (a, c, d), the Langevin (alpha, snr) and the annealed
        Langevin (1, step, sqrt(2 step)) coefficients"""
        t = timesteps.detach().float().cpu()
        a, c, d = predictor_coefficients(self.sde, self.predictor_name, t, next_timesteps, self.pf)
        sscale = self.sde.score_scale(t).float().reshape(-1)
        if isinstance(self.sde, sde_lib.VPSDE) and not self.net.config.training.ddpm_score:
            sscale = torch.ones_like(sscale)
        snr_t = torch.as_tensor(snr_per_step, dtype=torch.float32)
        z = torch.zeros_like(t)
        ca, cc, cd = (z + 1, z, z)
        if self.ald:
            std = self.sde.marginal_prob(torch.zeros((t.shape[0], 1, 1, 1)), t)[1]
            stepsz = (snr_t * std) ** 2 * 2 * self.sde.langevin_alpha(t)
            cc, cd = stepsz, torch.sqrt(stepsz * 2)
        rows = torch.stack([self.sde.time_cond(t).float(), sscale, a.float(), c.float(), d.float(), self.sde.langevin_alpha(t).float(),
                            snr_t, z, ca.float(), cc.float(), cd.float(), z], dim=1).contiguous()
        if self.sched is not None and tuple(self.sched.shape) == tuple(rows.shape):
            self.sched.copy_(rows)          # same buffer: the captured graph stays valid
        else:
            self.sched = rows.to(self.dev)
            self._temb_op = self.eng._bind_time_source(self.sched, self.step, self.LD, 0)
            self.graph = None
        self.n_rows = rows.shape[0]

    def _one_step(self, noise_c, noise_p):
        eng, N, D = self.eng, self.N, self.D
        if self.langevin:
            for i in range(self.n_steps):
                self._fill_scale()
                eng.launch(self._temb_op)
                L.call('indm_langevin_norms', L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.norms), L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
                if self.global_norms:
                    L.call('indm_langevin_norm_sums', L.ptr(self.norms), L.ptr(self.gsums), N)
                    parallel.allreduce_langevin_sums_(self.gsums)
                    L.call('indm_langevin_update_global', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.x_mean), L.ptr(self.gsums),
                           L.ptr(self.sched[:, 5:]), self.LD, L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
                    continue
                L.call('indm_langevin_update', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.x_mean), L.ptr(self.norms),
                       L.ptr(self.sched[:, 5:]), self.LD, L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
        elif self.ald:
            for i in range(self.n_steps):
                self._fill_scale()
                eng.launch(self._temb_op)
                L.call('indm_pc_predictor_update', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.x_mean),
                       L.ptr(self.sched[:, 8:]), self.LD, L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
        if self.predictor_name != 'none':
            self._fill_scale()
            eng.launch(self._temb_op)
            L.call('indm_pc_predictor_update', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_p), L.ptr(self.x_mean), L.ptr(self.sched[:, 2:]),
                   self.LD, L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 0)
        L.call('indm_advance_step', L.ptr(self.step))

    def _fill_scale(self):
        # engine.out_scale[n] <- sched[*step][1] (/ sigma if scale_by_sigma; time_cond column = sigma for VE)
        L.call('indm_sched_broadcast', L.ptr(self.eng.out_scale), self.N, L.ptr(self.sched), self.LD, 1,
               0 if self.net.config.model.scale_by_sigma else -1, L.ptr(self.step))

    def run(self, x_init, num_steps, noises=None):
        """x_init: device tensor [N,C,S,S].  Returns (x, x_mean) after num_steps replays."""
        if x_init is not None:
            self.x.copy_(x_init)
            self.step.zero_()
        if self.eng._weights_version != self.eng.weights_version():
            self.eng.load_weights()
        if noises is not None:
            it = noises if hasattr(noises, '__next__') else iter(noises)
            for _ in range(num_steps):
                zc = next(it).to(self.dev).contiguous() if (self.langevin or self.ald) else None
                if (self.langevin or self.ald) and self.n_steps != 1:
                    raise NotImplementedError('supplied noise with n_steps_each > 1')
                zp = next(it).to(self.dev).contiguous() if self.predictor_name != 'none' else None
                self._one_step(zc, zp)
            return self.x, self.x_mean
        if self.graph is None:
            # warm-up on a side stream, then capture
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._one_step(None, None)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            launches_before = L.launches
            with torch.cuda.graph(g):
                self._one_step(None, None)
            self.launches_per_step = L.launches - launches_before
            self.graph = g
            # warm-up advanced the state by one step: restore (capture itself executes nothing)
            if x_init is not None:
                self.x.copy_(x_init)
                self.step.zero_()
            else:
                raise RuntimeError('graph must be built on a fresh run')
        for _ in range(num_steps):
            self.graph.replay()
        L.launches += self.launches_per_step * num_steps
        return self.x, self.x_mean


def get_pc_sampler(config, sde, shape, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False,
                   continuous=False, denoise=True, eps=1e-3, device='cuda'):
    """Create a Predictor-Corrector (PC) sampler (sampling.py:365-456).  Returns
    `pc_sampler(model, flow_model, temperature=1., data_mean=None, final_time=0., before_data=None, sample_dir=None,
    r=None) -> (sample_before_flow, sample_after_flow, nfe)`; extra keyword-only arguments `noise=` (iterator of
    tensors replayed in the reference's draw order) and `seed=` select recorded vs in-kernel Philox noise."""
    if not continuous:
        raise NotImplementedError('INDM configs are continuous-time (training.continuous=True)')
    pred_name = 'none' if predictor is None else {ReverseDiffusionPredictor: 'reverse_diffusion', EulerMaruyamaPredictor: 'euler_maruyama',
                                                  AncestralSamplingPredictor: 'ancestral_sampling', NonePredictor: 'none'}.get(predictor)
    corr_name = 'none' if corrector is None else {LangevinCorrector: 'langevin', AnnealedLangevinDynamics: 'ald', NoneCorrector: 'none'}.get(corrector)
    if pred_name is None or corr_name is None:
        raise NotImplementedError(f'predictor {predictor} / corrector {corrector} are not INDM sampling components')

    def _sde_key():
        """everything of the SDE the schedule table depends on: a sampler built later for the same net with another SDE (other N,
        beta / sigma range, VE vs VP) must not reuse a cached graph's stale `sde`"""
        vals = tuple(float(getattr(sde, a)) for a in ('beta_0', 'beta_1', 'sigma_min', 'sigma_max') if hasattr(sde, a))
        return (type(sde).__name__, int(sde.N), float(sde.T)) + vals

    def _fresh_seed(seed):
        """seed=None (the reference-signature callers: sampling_lib.get_samples, utils.get_loss_fns): a fresh Philox key per call
        drawn from torch's global generator (so torch.manual_seed still controls it) with the rank mixed in — every sampling
        round and every rank gets its own noise path, like the reference's torch.randn_like draws (sampling.py:207,290)."""
        if seed is not None:
            return int(seed)
        s = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64).item())
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            s = (s + 0x9E3779B97F4A7C15 * (torch.distributed.get_rank() + 1)) % (2 ** 62)
        return s

    def _graph(net, seed):
        key = (shape[0], pred_name, corr_name, n_steps, probability_flow, _sde_key(), bool(getattr(config.sampling, 'global_langevin_norms', False)))
        cache = net.__dict__.setdefault('_pc_graphs', {})
        g = cache.get(key)
        if g is None:
            g = _GraphedPC(config, sde, net, shape[0], corr_name, n_steps, probability_flow, seed, predictor_name=pred_name)
            cache[key] = g
        g.sde = sde
        g.seed_dev.fill_(_fresh_seed(seed))
        return g

    def _finish(sample_before_flow, flow_model, temperature):
        from .flow_models.flow_model import flow_forward
        if config.flow.model != 'identity':
            sample_after_flow, _ = flow_forward(config, flow_model, sample_before_flow * temperature, log_det=None, reverse=True)
        else:
            sample_after_flow = sample_before_flow
        return inverse_scaler(sample_before_flow), inverse_scaler(sample_after_flow), sde.N * (n_steps + 1)

    def denoise_update_fn(model, x, final_time):
        """sampling.py:402-408: one noise-free probability-flow reverse-diffusion step from eps to final_time"""
        score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
        predictor_obj = ReverseDiffusionPredictor(sde, score_fn, probability_flow=True)
        vec_eps = torch.ones(x.shape[0], device=x.device) * eps
        _, x = predictor_obj.update_fn(x, vec_eps, torch.ones_like(vec_eps) * final_time)
        return x

    def pc_sampler_search(model, flow_model, temperature=1., data_mean=None, final_time=0., before_data=None, sample_dir=None, r=None,
                          *, noise=None, seed=None, prior=None):
        """sampling.py:458-493 (selected by sampling.pc_denoise): sde.N - 1 steps on the continuous-gap discretisation
        (explicit next_t, sde_lib.py:180-183,318-322), then one denoising step unless sampling.need_sample."""
        net = model.module if hasattr(model, 'module') else model
        net.eval()
        with torch.no_grad():
            if before_data is None:
                x0 = (sde.prior_sampling(shape, data_mean) if prior is None else prior).to(device)
                timesteps = torch.linspace(sde.T, eps, sde.N)
                g = _graph(net, seed if noise is None else (seed or 0))     # replayed noise: no generator draw
                g.set_schedule(timesteps[:-1], [config.sampling.snr] * (sde.N - 1), next_timesteps=timesteps[1:])
                x, x_mean = g.run(x0, sde.N - 1, noises=noise)
                x, x_mean = x.clone(), x_mean.clone()
            else:
                x_mean = x = before_data.to(device)
            if not config.sampling.need_sample:
                x_mean = x = denoise_update_fn(model, x_mean if denoise else x, final_time)
            return _finish((x_mean if denoise else x).clone(), flow_model, temperature)

    def pc_sampler_more_step(*a, **k):
        raise NotImplementedError('sampling.more_step: the reference variant indexes timesteps[i+1] out of range on its last step '
                                  '(sampling.py:509-513) and cannot run; not reproduced')

    def pc_sampler(model, flow_model, temperature=1., data_mean=None, final_time=0., before_data=None, sample_dir=None, r=None,
                   *, noise=None, seed=None, prior=None):
        net = model.module if hasattr(model, 'module') else model
        net.eval()
        with torch.no_grad():
            num_scales = config.sampling.num_scales if config.sampling.num_scales != sde.N else sde.N
            # prior draw on the host generator, like the reference (sde_lib.py:163,293), unless supplied
            x0 = sde.prior_sampling(shape, data_mean) if prior is None else prior
            x0 = x0.to(device)
            timesteps = torch.linspace(sde.T, eps, num_scales)
            if config.sampling.snr_scheduling == 'none':
                snrs = [config.sampling.snr] * num_scales
            else:
                snrs = [config.sampling.begin_snr + (config.sampling.end_snr - config.sampling.begin_snr) * i / num_scales
                        for i in range(num_scales)]
            g = _graph(net, seed if noise is None else (seed or 0))     # replayed noise: no generator draw
            g.set_schedule(timesteps, snrs)
            if sample_dir is not None and num_scales >= 2:
                # side effect of the reference loop at i == num_scales-2 (sampling.py:436-445): x_mean of that step
                x, x_mean = g.run(x0, num_scales - 1, noises=noise)
                samples = (inverse_scaler(x_mean).permute(0, 2, 3, 1).cpu().numpy() * 255.)
                samples = samples.reshape((-1, config.data.image_size, config.data.image_size, config.data.num_channels))
                with open(os.path.join(sample_dir, f"samples_{r}_before_flow_for_search.npz"), "wb") as fout:
                    io_buffer = io.BytesIO()
                    np.savez_compressed(io_buffer, samples=samples)
                    fout.write(io_buffer.getvalue())
                x, x_mean = g.run(None, 1, noises=noise)
            else:
                x, x_mean = g.run(x0, num_scales, noises=noise)
            return _finish((x_mean if denoise else x).clone(), flow_model, temperature)

    if getattr(config.sampling, 'pc_denoise', False):
        return pc_sampler_search
    if getattr(config.sampling, 'more_step', False):
        return pc_sampler_more_step
    return pc_sampler


def get_ode_sampler(config, sde, shape, inverse_scaler, denoise=False, rtol=1e-5, atol=1e-5, method='RK45', eps=1e-3, device='cuda'):
    """Probability-flow ODE sampler with the black-box solver (sampling.py:547-621; the default of configs/vp/*/indm_fid.py).
    The right-hand side f - g^2 score / 2 is one score-network forward on the engine; SciPy RK45 steps the whole batch on the host
    like the reference, or `method='RK45-device'` keeps the float64 state on the GPU (indm_b200/ode.py).  Returns `ode_sampler(model, flow_model, temperature, ...) -> (before_flow, after_flow, nfe)`."""
    from scipy import integrate

    def denoise_update_fn(model, x):
        score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
        predictor_obj = ReverseDiffusionPredictor(sde, score_fn, probability_flow=False)
        vec_eps = torch.ones(x.shape[0], device=x.device) * eps
        _, x = predictor_obj.update_fn(x, vec_eps, torch.zeros_like(vec_eps))
        return x

    def drift_fn(model, x, t):
        score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
        rsde = sde.reverse(score_fn, probability_flow=True)
        return rsde.sde(x, t)[0]

    def ode_sampler(model, flow_model, temperature=1., data_mean=None, final_time=0., before_data=None, sample_dir=None, r=None, *,
                    prior=None):
        from .flow_models.flow_model import flow_forward
        with torch.no_grad():
            x = (sde.prior_sampling(shape, data_mean) if prior is None else prior).to(device)

            def ode_func(t, x):
                x = mutils.from_flattened_numpy(x, shape).to(device).type(torch.float32)
                vec_t = torch.ones(shape[0], device=x.device) * t
                return mutils.to_flattened_numpy(drift_fn(model, x, vec_t))

            if method == 'RK45-device':
                # same Dormand-Prince 5(4) algorithm, float64 state resident on the GPU (indm_b200/ode.py)
                from .ode import solve_ivp_rk45

                def ode_func_device(t, y):
                    xs = y.reshape(shape).to(torch.float32)
                    vec_t = torch.ones(shape[0], device=xs.device) * t
                    return drift_fn(model, xs, vec_t).reshape(-1)

                solution = solve_ivp_rk45(ode_func_device, (sde.T, eps), x.reshape(-1).to(torch.float64), rtol=rtol, atol=atol)
                nfe = solution.nfev
                x = solution.y_final.reshape(shape).to(torch.float32)
            else:
                solution = integrate.solve_ivp(ode_func, (sde.T, eps), mutils.to_flattened_numpy(x), rtol=rtol, atol=atol, method=method)
                nfe = solution.nfev
                x = torch.tensor(solution.y[:, -1]).reshape(shape).to(device).type(torch.float32)
            sample_before_flow = denoise_update_fn(model, x) if denoise else x
            if config.flow.model != 'identity':
                sample_after_flow, _ = flow_forward(config, flow_model, sample_before_flow * temperature, log_det=None, reverse=True)
            else:
                sample_after_flow = sample_before_flow
            return inverse_scaler(sample_before_flow), inverse_scaler(sample_after_flow), nfe

    return ode_sampler
