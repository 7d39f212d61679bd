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
import functools
import io
import os

import numpy as np
import torch

from . import _lib as L
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
    """sampling.py:86-133 (the 'ode' black-box sampler is a later row of SURVEY.md §8f)."""
    sampler_name = config.sampling.method
    if sampler_name.lower() == 'pc':
        predictor = get_predictor(config.sampling.predictor.lower())
        corrector = get_corrector(config.sampling.corrector.lower())
        return get_pc_sampler(config=config, sde=sde, shape=shape, predictor=predictor, corrector=corrector,
                              inverse_scaler=inverse_scaler, snr=config.sampling.snr, n_steps=config.sampling.n_steps_each,
                              probability_flow=config.sampling.probability_flow, continuous=config.training.continuous,
                              denoise=config.sampling.noise_removal, eps=eps, device=config.device)
    raise ValueError(f"Sampler name {sampler_name} not available in indm_b200 (only 'pc' is on the hot path).")


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


@register_predictor(name='reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
    """x_mean = x - rev_f, x = x_mean + G z (sampling.py:200-210) — one fused kernel launch."""

    def update_fn(self, x, t, next_t=None, noise=None, seed=0, offset=0):
        if next_t is not None:
            raise NotImplementedError('explicit next_t (pc_sampler_search) is a later row of SURVEY.md §8f')
        score = self.score_fn(x, t).contiguous()
        a, c, d = self.sde.reverse_diffusion_coef(t.detach().float().cpu())
        if not (bool((a == a[0]).all()) and bool((c == c[0]).all())):
            raise NotImplementedError('per-sample time steps in one batch')
        if self.probability_flow:
            c, d = c * 0.5, d * 0.0
        coef = _coef_table([[float(a[0]), float(c[0]), float(d[0]), 0.0]], x.device)
        x = x.contiguous().clone()
        x_mean = torch.empty_like(x)
        N, D = x.shape[0], x[0].numel()
        L.call('indm_pc_predictor_update', L.ptr(x), L.ptr(score), L.ptr(noise.contiguous()) if noise is not None else None,
               L.ptr(x_mean), L.ptr(coef), 4, None, N, D, seed, None, offset)
        return x, x_mean


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

    def __init__(self, config, sde, net, batch, corrector_name, n_steps, probability_flow, seed):
        self.sde, self.net, self.N = sde, net, batch
        self.eng = net.engine(batch)
        self.dev = self.eng.dev
        self.langevin = corrector_name == 'langevin'
        self.n_steps = n_steps
        self.pf = probability_flow
        self.seed = seed
        eng = self.eng
        D = eng.ch * eng.S * eng.S
        self.D = D
        self.x = eng.x_in                                # the sampler state IS the network input buffer (no copies)
        self.x_mean = torch.zeros_like(self.x)
        self.norms = torch.zeros((batch, 2), device=self.dev)
        self.step = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.seed_dev = torch.zeros((1,), dtype=torch.int64, device=self.dev)   # read by the kernels: graph-replay safe
        self.sched = None
        self.graph = None

    def set_schedule(self, timesteps, snr_per_step):
        """per-step scalars (sampling.py:425-435): row = [time_cond, out_scale, a, c, d, alpha, snr, 0]"""
        t = timesteps.detach().float().cpu()
        a, c, d = self.sde.reverse_diffusion_coef(t)
        if self.pf:
            c, d = c * 0.5, d * 0.0
        sscale = self.sde.score_scale(t).float().reshape(-1)
        if isinstance(self.sde, sde_lib.VPSDE) and not self.net.config.training.ddpm_score:
            sscale = torch.ones_like(sscale)
        rows = torch.stack([self.sde.time_cond(t).float(), sscale, a.float(), c.float(),
                            d.float(), self.sde.langevin_alpha(t).float(), torch.as_tensor(snr_per_step, dtype=torch.float32),
                            torch.zeros_like(t)], dim=1).contiguous()
        if self.sched is not None and tuple(self.sched.shape) == tuple(rows.shape):
            self.sched.copy_(rows)          # same buffer: the captured graph stays valid
        else:
            self.sched = rows.to(self.dev)
            self.eng._bind_time_source(self.sched, self.step, 8, 0)
            self.graph = None
        self.n_rows = rows.shape[0]

    def _one_step(self, noise_c, noise_p):
        eng, N, D = self.eng, self.N, self.D
        if self.langevin:
            for i in range(self.n_steps):
                self._fill_scale()
                eng.launch()
                L.call('indm_langevin_norms', L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.norms), L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
                L.call('indm_langevin_update', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_c), L.ptr(self.x_mean), L.ptr(self.norms),
                       L.ptr(self.sched[:, 5:]), 8, L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 1 + i)
        self._fill_scale()
        eng.launch()
        L.call('indm_pc_predictor_update', L.ptr(self.x), L.ptr(eng.out), L.ptr(noise_p), L.ptr(self.x_mean), L.ptr(self.sched[:, 2:]), 8,
               L.ptr(self.step), N, D, 0, L.ptr(self.seed_dev), 0)
        L.call('indm_advance_step', L.ptr(self.step))

    def _fill_scale(self):
        # engine.out_scale[n] <- sched[*step][1] (/ sigma if scale_by_sigma; time_cond column = sigma for VE)
        L.call('indm_sched_broadcast', L.ptr(self.eng.out_scale), self.N, L.ptr(self.sched), 8, 1,
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
                zc = next(it).to(self.dev).contiguous() if self.langevin else None
                if self.langevin and self.n_steps != 1:
                    raise NotImplementedError('supplied noise with n_steps_each > 1')
                zp = next(it).to(self.dev).contiguous()
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
    if predictor is not ReverseDiffusionPredictor:
        raise NotImplementedError('pc_sampler hot path: reverse_diffusion predictor (others are SURVEY.md §8f rows)')
    if corrector not in (LangevinCorrector, NoneCorrector, None):
        raise NotImplementedError('pc_sampler hot path: langevin / none correctors')
    if not continuous:
        raise NotImplementedError('INDM configs are continuous-time (training.continuous=True)')
    corr_name = 'langevin' if corrector is LangevinCorrector else 'none'

    def pc_sampler(model, flow_model, temperature=1., data_mean=None, final_time=0., before_data=None, sample_dir=None, r=None,
                   *, noise=None, seed=0, prior=None):
        from .flow_models.flow_model import flow_forward
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
            key = (shape[0], corr_name, n_steps, probability_flow)
            cache = net.__dict__.setdefault('_pc_graphs', {})
            g = cache.get(key)
            if g is None:
                g = _GraphedPC(config, sde, net, shape[0], corr_name, n_steps, probability_flow, seed)
                cache[key] = g
            g.seed_dev.fill_(int(seed))
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
            sample_before_flow = (x_mean if denoise else x).clone()
            if config.flow.model != 'identity':
                sample_after_flow, _ = flow_forward(config, flow_model, sample_before_flow * temperature, log_det=None, reverse=True)
            else:
                sample_after_flow = sample_before_flow
            sample_before_flow = inverse_scaler(sample_before_flow)
            sample_after_flow = inverse_scaler(sample_after_flow)
            return sample_before_flow, sample_after_flow, sde.N * (n_steps + 1)

    return pc_sampler
