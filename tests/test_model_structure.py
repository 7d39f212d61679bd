"""CPU: the product's NCSNpp parameter container is state-dict compatible with the reference's (names, order, shapes),
the registry / create_model surface behaves like models/utils.py, and the CUDA-only contract is enforced."""
import pytest
import torch

from helpers import load_json, tiny
from indm_b200 import configs
from indm_b200.models import utils as mutils


def _cfg(tag):
    base = {'tiny_vp': 'vp/CIFAR10/indm_fid', 'tiny_ve': 've/CIFAR10/indm',
            'vp_cifar': 'vp/CIFAR10/indm_fid', 've_cifar': 've/CIFAR10/indm'}[tag]
    cfg = configs.get_config(base)
    if tag.startswith('tiny'):
        tiny(cfg)
    cfg.device = torch.device('cpu')
    return cfg


@pytest.mark.parametrize("tag", ['tiny_vp', 'tiny_ve', 'vp_cifar', 've_cifar'])
def test_state_dict_matches_reference(tag):
    model = mutils.create_model(_cfg(tag))
    want = [('module.' + k, tuple(s)) for k, s in load_json(f'shapes_{tag}.json')]
    got = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert got == want
    assert hasattr(model, 'module')


def test_registry_like_reference():
    assert mutils.get_model('ncsnpp').__name__ == 'NCSNpp'
    with pytest.raises(ValueError):
        mutils.register_model(name='ncsnpp')(type('X', (), {}))


def test_forward_refuses_cpu_tensors():
    model = mutils.create_model(_cfg('tiny_vp'))
    with torch.no_grad(), pytest.raises(RuntimeError, match='no CPU'):
        model(torch.zeros(1, 3, 16, 16), torch.zeros(1))
