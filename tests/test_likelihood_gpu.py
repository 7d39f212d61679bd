"""GPU: likelihood.get_likelihood_fn (probability-flow ODE NLL with Hutchinson trace, RK45) and likelihood.get_elbo_fn on a
small INDM-VP model (score net + wolf flow) against the live reference, every random draw replayed
(tests/golden/likelihood_small_vp.npz from tests/golden/make_golden.py:make_likelihood).
Tolerance (north_star): NELBO / NLL within 0.01 bpd."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, rel_l2  # noqa: E402
from indm_b200 import configs, sde_lib, likelihood  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402
from oracle import flow as oflow, ncsnpp as oncsnpp  # noqa: E402


def _setup(mode):
    g = load_npz('likelihood_small_vp.npz')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 128, (1, 2), 1, (16,)
    cfg.flow.nblocks, cfg.flow.intermediate_dim = '2-2', 128
    cfg.device = torch.device('cuda:0')
    model = mutils.create_model(cfg)
    model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, int(g['seed_score'])).items()})
    model.eval()
    flow = fm.create_flow_model(cfg)
    flow.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oflow.synth_params(cfg, int(g['seed_flow'])).items()})
    flow.eval()
    model.module.compute_mode = flow.module.compute_mode = mode
    return g, cfg, model, flow, sde_lib.get_sde(cfg)


def _draws(g, which, nblk):
    cu = lambda a: torch.from_numpy(a).cuda()
    flow_kw = dict(eps=cu(g[f'{which}_eps_post']), vareps=[cu(g[f'{which}_vareps_{i}']) for i in range(nblk)], n_terms=g[f'{which}_ns'])
    return flow_kw, cu(g[f'{which}_rad']) * 2 - 1., [cu(g[f'{which}_gauss_{i}']) for i in range(4)], cu(g[f'{which}_u'])


# 'auto' = the default precision policy (likelihood leg: score forward + Hutchinson VJP and flow log-det in compensated TF32): 0.01 bpd
@pytest.mark.parametrize("mode,tol", [('auto', 0.01), ('tf32', 0.01), ('bf16', 0.1)])
def test_pf_ode_nll_matches_reference(mode, tol):
    g, cfg, model, flow, sde = _setup(mode)
    flow_kw, rad, gauss, _ = _draws(g, 'nll', len(oflow.block_layout(cfg)))
    fn = likelihood.get_likelihood_fn(cfg, sde, lambda v: (v + 1.) / 2., rtol=1e-3, atol=1e-3)
    bpd, z, nfe = fn(model, flow, torch.from_numpy(g['data']).cuda(), eps_bpd=1e-5, epsilon=rad, noise=gauss[0],
                     residual_noise=(gauss[1], gauss[2]), flow_kw=flow_kw)
    torch.cuda.synchronize()
    err = float(np.abs(bpd.cpu().numpy() - g['nll_bpd']).max())
    print(f'NLL {mode}: bpd {bpd.cpu().numpy()} ref {g["nll_bpd"]} |err| {err:.2e}; nfe {nfe} ref {int(g["nll_nfe"])}')
    e_z = rel_l2(z.cpu().numpy(), g['nll_z'])
    print(f'latent z rel-L2 {e_z:.2e}')
    assert err < tol          # 0.01 bpd in the validation precision (north_star); BF16 is limited by the BF16 Hutchinson VJP
    # The PF-ODE of this random-weight network amplifies perturbations ~100x (measured on the live reference: 1e-5 relative
    # weight noise moves z by 1e-3), so the BF16 latent (score error ~1e-2 per evaluation) is only sanity-bounded; the
    # validation precision is held to 2e-2 (observed 6e-4).
    assert e_z < (0.6 if mode == 'bf16' else 2e-2)


def test_pf_ode_nll_device_integrator_matches_reference():
    """method='RK45-device' (indm_b200/ode.py: SciPy's Dormand-Prince controller with the float64 state resident on the GPU) against
    the live reference's SciPy run: same golden NLL within 0.01 bpd, same number of function evaluations as the SciPy path."""
    g, cfg, model, flow, sde = _setup('tf32')
    flow_kw, rad, gauss, _ = _draws(g, 'nll', len(oflow.block_layout(cfg)))
    data = torch.from_numpy(g['data']).cuda()
    out = {}
    for method in ('RK45', 'RK45-device'):
        fn = likelihood.get_likelihood_fn(cfg, sde, lambda v: (v + 1.) / 2., rtol=1e-3, atol=1e-3, method=method)
        out[method] = fn(model, flow, data, eps_bpd=1e-5, epsilon=rad, noise=gauss[0], residual_noise=(gauss[1], gauss[2]),
                         flow_kw=flow_kw)
    torch.cuda.synchronize()
    bpd, z, nfe = out['RK45-device']
    err = float(np.abs(bpd.cpu().numpy() - g['nll_bpd']).max())
    d = float((bpd - out['RK45'][0]).abs().max())
    print(f'NLL device integrator: |err| vs reference {err:.2e}, vs SciPy path {d:.2e}; nfe {nfe} / SciPy path {out["RK45"][2]} / ref {int(g["nll_nfe"])}')
    assert err < 0.01 and d < 0.01
    assert abs(nfe - out['RK45'][2]) <= 12          # same controller; fp32-atomics noise in the RHS may move one step decision
    assert rel_l2(z.cpu().numpy(), g['nll_z']) < 2e-2


# 0.01 bpd (north_star) is held in the validation precision (observed 1e-4).  The BF16 Hutchinson term carries the BF16 rounding
# of one forward + one VJP (score / VJP rel-L2 1e-2): observed 0.039 - 0.052 bpd from run to run (fp32 atomics in the fused
# GroupNorm statistics reorder the roundings), so the production precision is bounded at 0.08.
@pytest.mark.parametrize("mode,tol", [('auto', 0.01), ('tf32', 0.01), ('bf16', 0.08)])
def test_nelbo_matches_reference(mode, tol):
    g, cfg, model, flow, sde = _setup(mode)
    flow_kw, rad, gauss, u = _draws(g, 'elbo', len(oflow.block_layout(cfg)))
    fn = likelihood.get_elbo_fn(cfg, sde, lambda v: (v + 1.) / 2.)
    a, b = fn(model, flow, torch.from_numpy(g['data']).cuda(),
              draws=dict(u=u, z=gauss[0], epsilon=rad, lp_z=gauss[1], residual_noise=(gauss[2], gauss[3])), flow_kw=flow_kw)
    torch.cuda.synchronize()
    e_a = float(np.abs(a.cpu().numpy() - g['elbo_bpd']).max())
    e_b = float(np.abs(b.cpu().numpy() - g['elbo_bpd_residual']).max())
    print(f'NELBO {mode}: {a.cpu().numpy()} ref {g["elbo_bpd"]} |err| {e_a:.2e}; with residual |err| {e_b:.2e}')
    assert e_a < tol and e_b < tol
