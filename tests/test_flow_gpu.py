"""GPU: the wolf flow (latent prior flow + conditional residual flow) through the C-ABI kernels against golden outputs of
the live reference (tests/golden/flow_*.npz, made by tests/golden/make_golden.py:make_flow) and the oracle.

Tolerances (BASELINE.json north_star): inverse round trip within 1e-4 max-abs (TF32 validation mode, h held fixed,
tight stop rule — SURVEY.md §8c caveats i/ii); BF16 production mode is checked at the BF16 operand resolution.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, tiny_flow, rel_l2  # noqa: E402
from indm_b200 import configs  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402
from oracle import flow as oflow  # noqa: E402


def _cfg(tag):
    cfg = configs.get_config('vp/CELEBA/indm_fid' if tag == 'tiny_sq' else 'vp/CIFAR10/indm_fid')
    tiny_flow(cfg, tag == 'tiny_sq')
    cfg.device = torch.device('cuda:0')
    return cfg


def _flow(cfg, seed, mode):
    flow = fm.create_flow_model(cfg)
    flow.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oflow.synth_params(cfg, seed).items()})
    flow.eval()
    flow.module.compute_mode = mode
    return flow


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_state_dict_keys_match_reference(tag):
    import json, os
    from helpers import GOLDEN
    cfg = _cfg(tag)
    flow = fm.create_flow_model(cfg)
    with open(os.path.join(GOLDEN, f'shapes_flow_{tag}.json')) as f:
        want = [(k, tuple(s)) for k, s in json.load(f)]
    got = [(k, tuple(v.shape)) for k, v in flow.module.state_dict().items()]
    assert got == want


@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_prior_flow_sample_matches_reference(tag):
    """FlowPrior.sample (priors/flow.py:226-230) in one launch, FP32: h = backward-program(eps)."""
    g = load_npz(f'flow_{tag}.npz')
    cfg = _cfg(tag)
    flow = _flow(cfg, int(g['seed']), 'bf16')
    eng = flow.module.engine(g['eps'].shape[0])
    eng._ensure()
    eps = torch.from_numpy(g['eps']).cuda()
    h, ld_b = eng.prior_flow(eps, 'backward', want_logdet=True)
    assert rel_l2(h.cpu().numpy(), g['h']) < 1e-5
    # both directions and their log-determinants against the oracle restatement (the synthetic weight_inv buffer is an
    # independent random matrix, like a stale buffer would be — SURVEY.md §7 hard part 6 — so the two programs are not
    # inverses of each other here and are checked separately)
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    h_ref, ldb_ref = oflow.prior_flow(cfg, P, torch.from_numpy(g['eps']), backward=True)
    assert float((ld_b.cpu() - ldb_ref).abs().max()) < 1e-3 * max(1.0, float(ldb_ref.abs().max()))
    fwd, ld_f = eng.prior_flow(h, 'forward', want_logdet=True)
    fwd_ref, ldf_ref = oflow.prior_flow(cfg, P, h_ref, backward=False)
    assert rel_l2(fwd.cpu().numpy(), fwd_ref.numpy()) < 1e-4
    assert float((ld_f.cpu() - ldf_ref).abs().max()) < 1e-3 * max(1.0, float(ldf_ref.abs().max()))


# 'auto' = the default precision policy (indm_b200/precision.py: flow reverse / eval legs in compensated TF32) = what users and bench.py
# run: held to north_star's numbers.  'bf16' forces BF16 everywhere and is bounded at BF16 operand resolution.
@pytest.mark.parametrize("mode,tol", [('auto', 2e-4), ('tf32', 2e-4), ('bf16', 5e-3)])
@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_wolf_reverse_matches_reference(tag, mode, tol):
    """flow_forward(config, flow, z, reverse=True): prior sample of h + fixed-point inverse of every iResBlock."""
    g = load_npz(f'flow_{tag}.npz')
    cfg = _cfg(tag)
    flow = _flow(cfg, int(g['seed']), mode)
    z, eps = torch.from_numpy(g['z']).cuda(), torch.from_numpy(g['eps']).cuda()
    x, ld = fm.flow_forward(cfg, flow, z, log_det=None, reverse=True, eps=eps)
    torch.cuda.synchronize()
    err = float(np.abs(x.cpu().numpy() - g['x']).max())
    print(f'{tag} {mode}: reverse max-abs err {err:.3e}; iterations {flow.module.engine(z.shape[0]).iterations}')
    assert ld == -1 and x.shape == z.shape
    assert err < tol


@pytest.mark.parametrize("mode,tol,rt_tol", [('auto', 2e-4, 1e-4), ('tf32', 2e-4, 1e-4), ('bf16', 5e-3, 5e-3)])
@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_resflow_forward_and_round_trip(tag, mode, tol, rt_tol):
    """ResidualFlow.fwdpass(x, h, eval_logdet=False) against the reference, then bwdpass with the same h."""
    g = load_npz(f'flow_{tag}.npz')
    cfg = _cfg(tag)
    flow = _flow(cfg, int(g['seed']), mode)
    core = flow.module
    xin, h = torch.from_numpy(g['xin']).cuda(), torch.from_numpy(g['h']).cuda()
    xf = fm.squeeze2(xin).contiguous() if cfg.flow.squeeze else xin
    zf = core(xf, reverse=False, eval_logdet=False, h=h)
    err = float(np.abs(zf.cpu().numpy() - g['zf']).max())
    back = core(zf, reverse=True, h=h, atol=1e-10, rtol=1e-10)
    rt = float((back - xf).abs().max())
    print(f'{tag} {mode}: forward max-abs err {err:.3e}, round trip {rt:.3e}')
    assert err < tol
    assert rt < rt_tol


def _cfg_fwd(tag):
    cfg = configs.get_config('vp/CELEBA/indm_nll' if tag == 'tiny_sq' else 'vp/CIFAR10/indm_nll')
    tiny_flow(cfg, tag == 'tiny_sq')
    cfg.data.image_size = cfg.flow.image_size = 64 if tag == 'tiny_sq' else 32
    cfg.device = torch.device('cuda:0')
    return cfg


@pytest.mark.parametrize("mode,tol_z,tol_ld", [('auto', 1e-4, 1e-3), ('tf32', 1e-4, 1e-3), ('bf16', 5e-3, 5e-2)])
@pytest.mark.parametrize("tag", ['tiny', 'tiny_sq'])
def test_wolf_forward_logdet_kl_matches_reference(tag, mode, tol_z, tol_ld):
    """flow_forward(config, flow, x, reverse=False) in eval mode — posterior encoder (BN folded, ELU), reparameterisation,
    prior-flow KL, and the (20 + n)-term power-series log-det of every iResBlock through the tensor-core VJP chain —
    against the live reference with every random draw replayed (north_star: flow logdet within 1e-3 relative)."""
    g = load_npz(f'flowfwd_{tag}.npz')
    cfg = _cfg_fwd(tag)
    flow = _flow(cfg, int(g['seed']), mode)
    nblk = len(oflow.block_layout(cfg))
    varepss = [torch.from_numpy(g[f'vareps_{i}']).cuda() for i in range(nblk)]
    x = torch.from_numpy(g['x']).cuda()
    z, ldkl = fm.flow_forward(cfg, flow, x, reverse=False, eps=torch.from_numpy(g['eps_post']).cuda(), vareps=varepss, n_terms=g['ns'])
    torch.cuda.synchronize()
    e_z = float(np.abs(z.cpu().numpy() - g['z']).max())
    e_ld = float(np.abs(ldkl.cpu().numpy() - g['ldkl']).max() / np.abs(g['ldkl']).max())
    eng = flow.module.engine(x.shape[0])
    print(f'{tag} {mode}: z max-abs err {e_z:.3e}, (logdet - KL) rel err {e_ld:.3e}, {eng.vjp_count} VJPs; ref {g["ldkl"]} got {ldkl.cpu().numpy()}')
    assert z.shape == x.shape and e_z < tol_z and e_ld < tol_ld


@pytest.mark.parametrize("tag", ['tiny'])
def test_posterior_and_kl_match_oracle(tag):
    g = load_npz(f'flowfwd_{tag}.npz')
    cfg = _cfg_fwd(tag)
    flow = _flow(cfg, int(g['seed']), 'tf32')
    P = oflow.to_torch(oflow.synth_params(cfg, int(g['seed'])))
    x, eps = torch.from_numpy(g['x']), torch.from_numpy(g['eps_post'])
    with torch.no_grad():
        h_ref, kl_ref, _, _ = oflow.posterior_sample_and_kl(cfg, P, x, eps)
    h, kl = flow.module.engine(x.shape[0]).posterior(x.cuda(), eps=eps.cuda())
    assert rel_l2(h.cpu().numpy(), h_ref.numpy()) < 1e-4
    assert float((kl.cpu() - kl_ref).abs().max()) < 1e-3 * float(kl_ref.abs().max())
