"""`flow_forward` / `create_flow_model` with the reference's interface (flow_models/flow_model.py:7-111), wolf branch."""
import torch


def flow_forward(config, flow_model, x, log_det=0, reverse=False):
    """flow_models/flow_model.py:7-69 (wolf branch :53-67)."""
    if config.flow.model == 'identity':
        return flow_model(x, reverse=reverse) if flow_model is not None else (x, -1)
    if config.flow.model != 'wolf':
        raise NotImplementedError(f"flow.model={config.flow.model!r}: only 'wolf' (all INDM configs) and 'identity'")
    raise NotImplementedError('wolf flow: CUDA path lands in the next milestone')


def create_flow_model(config):
    raise NotImplementedError('wolf flow: CUDA path lands in the next milestone')
