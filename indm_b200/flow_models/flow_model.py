"""`flow_forward` / `create_flow_model` with the reference's interface (flow_models/flow_model.py:7-111), wolf branch."""
import json
import os

import torch

from ..models.utils import SingleDeviceParallel
from .wolf import WolfCore


def squeeze2(x):
    """SqueezeLayer(2).forward (flow_models/resflow/layers/squeeze.py:32-45): a pure permutation, no arithmetic."""
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 4, h // 2, w // 2)


def unsqueeze2(x):
    """SqueezeLayer(2).inverse (squeeze.py:19-30)."""
    return torch.nn.functional.pixel_shuffle(x, 2)


def flow_forward(config, flow_model, x, log_det=0, reverse=False, **kw):
    """flow_models/flow_model.py:7-69 (wolf branch :53-67).  Extra keyword arguments (eps=, h=, seed=, vareps=, n_terms=)
    are handed to `WolfCore.forward`: they pin the random draws for parity tests."""
    if config.flow.model == 'identity':
        return flow_model(x, reverse=reverse) if flow_model is not None else (x, -1)
    if config.flow.model != 'wolf':
        raise NotImplementedError(f"flow.model={config.flow.model!r}: only 'wolf' (all INDM configs) and 'identity'")
    if config.flow.squeeze:
        x = squeeze2(x)
    if not reverse:
        if log_det == 0:
            z, logdet_kl = flow_model(x, y=None, n_bits=config.flow.n_bits, nsamples=config.flow.train_k, reverse=False,
                                      eval_logdet=True, **kw)
        else:
            z = flow_model(x, y=None, n_bits=config.flow.n_bits, nsamples=config.flow.train_k, reverse=False, eval_logdet=False, **kw)
            logdet_kl = -1
    else:
        z = flow_model(x, reverse=True, **kw).view(x.shape)
        logdet_kl = -1
    if config.flow.squeeze:
        z = unsqueeze2(z)
    return z, logdet_kl


def create_flow_model(config):
    """flow_models/flow_model.py:86-111, wolf branch: WolfCore.from_params(json, config) on config.device, wrapped where
    the reference puts nn.DataParallel (state-dict keys keep the `module.` prefix)."""
    if config.flow.model != 'wolf':
        raise NotImplementedError(f"flow.model={config.flow.model!r}: only 'wolf' is used by the INDM configs")
    params = config.flow.get('wolf_params') if hasattr(config.flow, 'get') else None
    if params is None:
        path = config.flow.model_config
        if not os.path.exists(path):
            raise FileNotFoundError(f'flow.model_config {path!r} not found and config.flow.wolf_params not set')
        with open(path) as f:
            params = json.load(f)
    flow_model = WolfCore.from_params(params, config)
    flow_model.add_config(config)
    flow_model = flow_model.to(config.device)
    return SingleDeviceParallel(flow_model)
