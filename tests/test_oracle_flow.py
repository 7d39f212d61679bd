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


def _cfg_fwd(tag):
    cfg = configs.get_config('vp/CELEBA/indm_nll' if tag == 'tiny_sq' else 'vp/CIFAR10/indm_nll')
    tiny_flow(cfg, tag == 'tiny_sq')
    cfg.data.image_size = cfg.flow.image_size = 64 if tag == 'tiny_sq' else 32
    return cfg


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_wolf_forward_logdet_kl_matches_reference(tag):
    """flow_forward(reverse=False) in eval mode: posterior encoder, reparameterisation, prior-flow KL and the (20 + n)-term
    power-series log-det of every iResBlock, all random draws replayed from the reference run."""
    g = load_npz(f'flowfwd_{tag}.npz')
    cfg = _cfg_fwd(tag)
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    nblk = len(oflow.block_layout(cfg))
    varepss = [torch.from_numpy(g[f'vareps_{i}']) for i in range(nblk)]
    with torch.no_grad():
        z, ldkl, h, kl = oflow.wolf_forward(cfg, P, torch.from_numpy(g['x']), torch.from_numpy(g['eps_post']), g['ns'], varepss)
    assert float(np.abs(z.numpy() - g['z']).max()) < 1e-5
    assert float(np.abs(ldkl.numpy() - g['ldkl']).max()) < 1e-3 * float(np.abs(g['ldkl']).max())


def test_series_coefficients_follow_the_reference_rule():
    # eval: 20 exact terms then Russian-roulette reweighting by 1 / P(N >= k - 20); train: 2 exact terms
    K, c = oflow.series_coefficients(3, training=False)
    assert K == 23 and c[:20] == [1.0] * 20 and c[20] == pytest.approx(1.0 / (1 - np.exp(-2.0)))
    K, c = oflow.series_coefficients(0, training=True)
    assert K == 2 and c == [1.0, 1.0]


def test_training_gradients_of_the_oracle_match_reference():
    """The oracle's differentiable training forward (batch-statistics encoder, KL, Neumann log-det series) against the live
    reference's autograd: loss values and the gradient of EVERY flow parameter (tests/golden/flowtrain_tiny.npz).  This pins the
    checker the GPU flow-backward tests are judged by."""
    g = load_npz('flowtrain_tiny.npz')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    tiny_flow(cfg, False)
    cfg.data.image_size = cfg.flow.image_size = 32
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    for k, v in P.items():
        if v.dtype.is_floating_point and ('running_' not in k) and not k.endswith('weight_inv') and not k.endswith('.scale'):
            v.requires_grad_(True)
    nblk = len(oflow.block_layout(cfg))
    varepss = [torch.from_numpy(g[f'vareps_{i}']) for i in range(nblk)]
    z, ldkl, h, kl = oflow.wolf_train_forward(cfg, P, torch.from_numpy(g['x']), torch.from_numpy(g['eps_post']), g['ns'], varepss)
    assert float(np.abs(z.detach().numpy() - g['z']).max()) < 1e-5
    assert float(np.abs(ldkl.detach().numpy() - g['ldkl']).max()) < 1e-4 * float(np.abs(g['ldkl']).max())
    assert rel_l2(h.detach().numpy(), g['h']) < 1e-5
    loss = (z * torch.from_numpy(g['Gz'])).sum() + (ldkl * torch.from_numpy(g['cl'])).sum()
    loss.backward()
    worst, n = 0.0, 0
    for k in list(g.keys()):
        if not k.startswith('grad.'):
            continue
        got = P[k[5:]].grad
        assert got is not None, k
        worst = max(worst, rel_l2(got.numpy(), g[k]))
        n += 1
    assert n > 150 and worst < 2e-4, (n, worst)
