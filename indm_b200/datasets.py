"""The tensor-side half of the reference's data path that sits directly on the hot path's input: scalers (datasets.py:56-71),
uniform dequantisation (run_lib.py:85-86, evaluation.py:404), batch fetch with epoch wrap-around (datasets.py:106-128).  Dataset
readers (TFDS / torchvision) are out of scope."""
import logging

import torch


def get_data_scaler(config):
    """Data normalizer. Assume data are always in [0, 1] (datasets.py:56-62)."""
    if config.data.centered:
        return lambda x: x * 2. - 1.
    return lambda x: x


def get_data_inverse_scaler(config):
    """Inverse data normalizer (datasets.py:65-71)."""
    if config.data.centered:
        return lambda x: (x + 1.) / 2.
    return lambda x: x


def dequantize(batch, noise=None):
    """(255 x + u) / 256 with u ~ U[0, 1) (run_lib.py:86): 8-bit images in [0, 1] -> continuous density on [0, 1)."""
    if noise is None:
        noise = torch.rand_like(batch)
    return (255. * batch + noise) / 256.


def get_batch(config, data_iter, data):
    """datasets.py:106-128 for iterables that yield NCHW float tensors (or (tensor, label) pairs): wraps to a new epoch when the
    iterator is exhausted, checks the shape, moves to `config.device`."""
    try:
        batch = next(data_iter)
    except StopIteration:
        logging.info('New Epoch Start')
        data_iter = iter(data)
        batch = next(data_iter)
    if isinstance(batch, (tuple, list)):
        batch = batch[0]
    assert batch.shape == (batch.shape[0], config.data.num_channels, config.data.image_size, config.data.image_size)
    return batch.to(config.device), data_iter
