"""CPU: oracle/flow.py against golden vectors produced by the live reference (wolf reverse pass, resflow forward)."""
import numpy as np
import pytest
import torch

from helpers import load_npz, load_json, tiny_flow, rel_l2
from indm_b200 import configs
from oracle import flow as oflow


def _cfg(tag):
    cfg = configs.get_config('vp/CELEBA/indm_fid' if tag == 'tiny_sq' else 'vp/CIFAR10/indm_fid')
    return tiny_flow(cfg, tag == 'tiny_sq')


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_flow_param_shapes_match_reference(tag):
    want = [(k, tuple(s)) for k, s in load_json(f'shapes_flow_{tag}.json')]
    assert [(k, tuple(s)) for k, s in oflow.param_shapes(_cfg(tag))] == want


def test_full_size_flow_has_the_probed_entry_count():
    # 687 state-dict entries for the CIFAR wolf flow
    assert len(oflow.param_shapes(configs.get_config('vp/CIFAR10/indm_fid'))) == 687


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_wolf_reverse_matches_reference(tag):
    g = load_npz(f'flow_{tag}.npz')
    cfg = _cfg(tag)
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    with torch.no_grad():
        x, h, iters = oflow.wolf_reverse(cfg, P, torch.from_numpy(g['z']), torch.from_numpy(g['eps']))
    assert rel_l2(h.numpy(), g['h']) < 1e-5
    assert float(np.abs(x.numpy() - g['x']).max()) < 1e-5
    assert max(iters) >= 1      # x0 = y - g(y) plus at least one more sweep under the reference stop rule


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_resflow_forward_matches_reference_and_round_trips(tag):
    g = load_npz(f'flow_{tag}.npz')
    cfg = _cfg(tag)
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    xin, h = torch.from_numpy(g['xin']), torch.from_numpy(g['h'])
    xf = oflow.squeeze2(xin) if cfg.flow.squeeze else xin
    with torch.no_grad():
        zf = oflow.resflow_forward(cfg, P, xf, h)
        assert rel_l2(zf.numpy(), g['zf']) < 1e-5
        # inverse round trip with the same h (SURVEY.md §8c caveat i), tight stop rule
        back, _ = oflow.resflow_inverse(cfg, P, zf, h, atol=1e-10, rtol=1e-10)
    assert float((back.reshape(xf.shape) - xf).abs().max()) < 1e-4
