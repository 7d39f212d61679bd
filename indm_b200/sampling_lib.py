"""Sample cache layout of the reference (sampling_lib.py:30-110): per round `r`
  <sample_dir>/samples_{r}_before_flow.npz      float64 NHWC, latent-side samples x 255 (unclipped, not rounded)
  <this_sample_dir>/samples_{r}.npz             uint8 NHWC, clip(after_flow x 255, 0, 255)
both `np.savez_compressed(samples=...)`; a round whose before-flow file exists is not re-sampled."""
import logging
import os

import numpy as np
import torch


def _nhwc(config, x, scale=255.):
    s = x.permute(0, 2, 3, 1).cpu().numpy() * scale
    s = s.reshape((-1, config.data.image_size, config.data.image_size, config.data.num_channels))
    assert s.shape == (s.shape[0], config.data.image_size, config.data.image_size, config.data.num_channels)
    return s


def _save(path, samples):
    with open(path, 'wb') as fout:
        np.savez_compressed(fout, samples=samples)


def _load_before(path):
    return torch.tensor(np.load(path)['samples']).permute(0, 3, 1, 2) / 255.


def get_samples(config, score_model, flow_model, sampling_fn, step, r, sample_dir, temperature=1., inverse_scaler=None,
                this_sample_dir=None, scaler=None, data_mean=None):
    """sampling_lib.py:30-110.  Runs `sampling_fn(score_model, flow_model, temperature, data_mean, sample_dir=, r=)` unless round `r`
    is cached, writes the two npz files, and with `sampling.pc_denoise` runs the extra denoising pass from the cached latent-side
    samples (`final_time=sampling.pc_denoise_time`, `before_data=scaler(samples)`), cached as `samples_{r}_denoise_{time}.npz` and
    `samples_{r}_before_flow_denoise_{time}.npz`.  Returns the uint8 NHWC samples of the round."""
    logging.info("sampling -- ckpt step: %d, round: %d" % (step, r))
    os.makedirs(sample_dir, exist_ok=True)
    os.makedirs(this_sample_dir, exist_ok=True)
    before_path = os.path.join(sample_dir, f'samples_{r}_before_flow.npz')
    final_path = os.path.join(this_sample_dir, f'samples_{r}.npz')
    if not os.path.exists(before_path):
        before, after, n = sampling_fn(score_model, flow_model, temperature, data_mean, sample_dir=sample_dir, r=r)
        logging.info(f'nfe: {n}')
        _save(before_path, _nhwc(config, before))
        samples = np.clip(_nhwc(config, after), 0., 255.).astype(np.uint8)
        _save(final_path, samples)
    else:
        samples = np.load(final_path)['samples'] if os.path.exists(final_path) else None
    if config.sampling.pc_denoise:
        t = config.sampling.pc_denoise_time
        den_path = os.path.join(this_sample_dir, f'samples_{r}_denoise_{t}.npz')
        den_before_path = os.path.join(sample_dir, f'samples_{r}_before_flow_denoise_{t}.npz')
        if not os.path.exists(den_path):
            if not os.path.exists(den_before_path):
                logging.info(f'denoise for pc with round {r} and final time {t}')
                src = os.path.join(sample_dir, f'samples_{r}_before_flow_for_search.npz') if config.training.sde == 'vesde' else before_path
                before = _load_before(src)
                before, after, n = sampling_fn(score_model, flow_model, temperature, data_mean, final_time=t,
                                               before_data=scaler(before))
                _save(den_before_path, _nhwc(config, before))
            else:
                from .flow_models.flow_model import flow_forward
                before = _load_before(den_before_path)
                with torch.no_grad():
                    x = scaler(before).to(config.device).float()
                    if config.flow.model != 'identity':
                        after, _ = flow_forward(config, flow_model, x * temperature, log_det=None, reverse=True)
                    else:
                        after = x
                    after = inverse_scaler(after)
            samples = np.clip(_nhwc(config, after), 0., 255.).astype(np.uint8)
            _save(den_path, samples)
        else:
            samples = np.load(den_path)['samples']
    return samples
