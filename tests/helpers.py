"""Shared helpers for the test-suite (config shrinkers, error metrics)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def tiny(cfg, nf=128):
    """Same shrink as tests/golden/make_golden.py:tiny."""
    cfg.model.nf = nf
    cfg.model.ch_mult = (1, 2)
    cfg.model.num_res_blocks = 1
    cfg.model.attn_resolutions = (8,)
    cfg.data.image_size = 16
    return cfg


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor * max(np.max(np.abs(b)), 1e-30) + 1e-30)))


def tiny_flow(cfg, squeeze):
    """Same shrink as tests/golden/make_golden.py:tiny_flow."""
    cfg.flow.nblocks = '2-2'
    cfg.flow.intermediate_dim = 128
    cfg.data.image_size = 16
    cfg.flow.image_size = 16
    cfg.flow.squeeze = squeeze
    return cfg
