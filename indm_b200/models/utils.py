"""Model registry, `create_model`, `get_model_fn`, `get_score_fn` — same names, arguments and behaviour as the
reference's models/utils.py (register_model :27-43, create_model :88-94, get_model_fn :96-125,
get_score_fn :140-197), VP and VE SDEs (the BASELINE configs)."""
import numpy as np
import torch

from .. import sde_lib

_MODELS = {}


def register_model(cls=None, *, name=None):
    """A decorator for registering model classes (models/utils.py:27-43)."""

    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _MODELS:
            raise ValueError(f'Already registered model with name: {local_name}')
        _MODELS[local_name] = cls
        return cls

    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]


def get_sigmas(config):
    """models/utils.py:46-57"""
    return np.exp(np.linspace(np.log(config.model.sigma_max), np.log(config.model.sigma_min), config.model.num_scales))


class SingleDeviceParallel(torch.nn.Module):
    """Stands where the reference puts `torch.nn.DataParallel` (models/utils.py:93): exposes `.module` and prefixes
    state-dict keys with `module.` so checkpoints stay interchangeable.  Parallelism here is one process per GPU
    (batch sharding for sampling, NCCL gradient all-reduce for training), not per-call replication."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def create_model(config):
    """Create the score model (models/utils.py:88-94)."""
    score_model = get_model(config.model.name)(config)
    score_model = score_model.to(config.device)
    return SingleDeviceParallel(score_model)


def get_model_fn(model, train=False):
    """models/utils.py:96-125"""

    def model_fn(x, labels):
        if not train:
            model.eval()
            return model(x, labels)
        model.train()
        return model(x, labels)

    return model_fn


def get_score_fn(config, sde, model, gamma_t=None, train=False, continuous=False):
    """models/utils.py:140-197 (VDM / unbounded parametrisation / subVP branches are out of scope: no BASELINE
    config selects them)."""
    model_fn = get_model_fn(model, train=train)

    if isinstance(sde, sde_lib.VPSDE):
        def score_fn(x, t):
            if continuous:
                labels = t * 999
                score = model_fn(x, labels)
                std = sde.marginal_prob(torch.zeros_like(t)[:, None, None, None], t)[1]
            else:
                labels = t * (sde.N - 1)
                score = model_fn(x, labels)
                std = sde.sqrt_1m_alphas_cumprod.to(labels.device)[labels.long()]
            if config.training.ddpm_score:
                score = -score / std[:, None, None, None]
            return score
    elif isinstance(sde, sde_lib.VESDE):
        def score_fn(x, t):
            if continuous:
                labels = sde.marginal_prob(torch.zeros_like(t)[:, None, None, None], t)[1]
            else:
                labels = sde.T - t
                labels *= sde.N - 1
                labels = torch.round(labels).long()
            return model_fn(x, labels)
    else:
        raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
    return score_fn


def to_flattened_numpy(x):
    return x.detach().cpu().numpy().reshape((-1,))


def from_flattened_numpy(x, shape):
    return torch.from_numpy(x.reshape(shape))
