"""Checkpoint wire format of the reference (utils.py:14-48): `checkpoint_N.pth` / `flow_checkpoint_N.pth` are `torch.save`d dicts
`{'optimizer': AdamW.state_dict(), 'model': DataParallel state dict ('module.'-prefixed keys), 'ema': {'decay', 'num_updates',
'shadow_params'}, 'step': int}`.  Same function names, arguments and behaviour; the VDM-only entries are out of scope.
`FusedAdamW` reads and writes torch.optim.AdamW's per-parameter layout (losses.pack_adamw_state), the model containers keep the
reference's state-dict keys (tests/test_model_structure.py), so files move between the two code bases in both directions."""
import logging
import os

import torch


_STATEFUL = ('optimizer', 'model', 'ema')          # entries restored through load_state_dict / saved through state_dict


def restore_checkpoint(config, ckpt_dir, state, device):
    """utils.py:14-34.  Missing file: warn, create the parent directory, return `state` unchanged.  VE-SDE runs do not restore the
    optimizer (utils.py:23-24); the model is loaded with strict=False like the reference."""
    if not os.path.exists(ckpt_dir):
        os.makedirs(os.path.dirname(ckpt_dir) or '.', exist_ok=True)
        logging.warning(f"No checkpoint found at {ckpt_dir}. Returned the same state as input")
        return state
    logging.info(ckpt_dir + ' loaded ...')
    loaded = torch.load(ckpt_dir, map_location=device, weights_only=False)
    for key in _STATEFUL:
        if key == 'optimizer' and config.training.sde == 'vesde':
            continue
        if key == 'model':
            state[key].load_state_dict(loaded[key], strict=False)
        else:
            state[key].load_state_dict(loaded[key])
    state['step'] = loaded['step']
    from . import _lib as L
    L.param_epoch += 1                 # engines repack their operand copies of the weights on next use
    return state


def save_checkpoint(config, ckpt_dir, state):
    """utils.py:37-48"""
    payload = {key: state[key].state_dict() for key in _STATEFUL}
    payload['step'] = state['step']
    torch.save(payload, ckpt_dir)


def create_name(prefix, name, ext):
    """utils.py:50-59: `checkpoint_12.pth` from 12 / '12', `<prefix>_<stem>.<ext>` from a path-like name."""
    try:
        name = f'{prefix}_{int(name)}.{ext}'
    except (TypeError, ValueError):
        if len(name.split('.')) == 1:
            name = f'{prefix}_{name}.{ext}'
        else:
            name = name.split('/')[-1]
            name = f'{prefix}_{name.split(".")[0]}.{ext}'
    return name


def get_loss_fns(config, sde, inverse_scaler, train=True, scaler=None):
    """utils.py:132-140: the four per-step callables the reference's train / eval loops use —
    `(train_step_fn, nll_fn, nelbo_fn, sampling_fn)` — built from this package's hot-path factories."""
    from . import likelihood, losses, sampling
    optimize_fn = losses.optimization_manager(config)
    train_step_fn = losses.get_step_fn(config, sde, train=train, optimize_fn=optimize_fn, scaler=scaler)
    nll_fn = likelihood.get_likelihood_fn(config, sde, inverse_scaler, rtol=config.eval.rtol, atol=config.eval.atol)
    nelbo_fn = likelihood.get_elbo_fn(config, sde, inverse_scaler=inverse_scaler)
    sampling_shape = (config.sampling.batch_size, config.data.num_channels, config.data.image_size, config.data.image_size)
    sampling_fn = sampling.get_sampling_fn(config, sde, sampling_shape, inverse_scaler, config.sampling.truncation_time)
    return train_step_fn, nll_fn, nelbo_fn, sampling_fn
