"""GPU: gradients of the wolf flow's TRAINING forward (joint flow + score training, losses.py:258-320) on the explicit backward
plan (indm_b200/flow_models/wolf_backward.py) against the live reference's autograd (tests/golden/flowtrain_tiny.npz, made by
make_golden.py:make_flowtrain): L = <z, Gz> + <logdet - KL, cl>, every random draw replayed, every flow parameter compared.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, tiny_flow, rel_l2  # noqa: E402
from indm_b200 import configs  # noqa: E402
from indm_b200.flow_models import flow_model as fm  # noqa: E402
from indm_b200.flow_models.wolf_backward import FlowBackward  # noqa: E402
from oracle import flow as oflow  # noqa: E402


def _setup(mode):
    g = load_npz('flowtrain_tiny.npz')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    tiny_flow(cfg, False)
    cfg.data.image_size = cfg.flow.image_size = 32
    cfg.device = torch.device('cuda:0')
    flow = fm.create_flow_model(cfg)
    flow.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oflow.synth_params(cfg, int(g['seed'])).items()})
    flow.module.compute_mode = mode
    return g, cfg, flow


@pytest.mark.parametrize("mode,tol", [('tf32', 2e-3), ('bf16', 6e-2)])
def test_resflow_block_gradients_match_reference(mode, tol):
    """Stage A: the residual-flow blocks with the conditioning latent h given — forward value of the training (Neumann) series,
    then every conv / conditioning-layer parameter gradient (first order through g + second order through the log-det
    estimator + the Lipschitz normalisation) and d L / d h of the blocks."""
    g, cfg, flow = _setup(mode)
    core = flow.module
    N = g['x'].shape[0]
    eng = core.engine(N)
    nblk = len(oflow.block_layout(cfg))
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    varepss = [cu(g[f'vareps_{i}']) for i in range(nblk)]
    z, logpx = eng.forward_logdet(cu(g['x']), cu(g['h']), vareps=varepss, n_terms=g['ns'], training=True, save=True)
    e_z = float(np.abs(z.cpu().numpy() - g['z']).max())
    ld = (-logpx).cpu().numpy()
    e_ld = float(np.abs(ld - (g['ldkl'] + g['kl'])).max() / np.abs(g['ldkl'] + g['kl']).max())
    print(f'{mode}: z max-abs err {e_z:.3e}; training log-det rel err {e_ld:.3e} ({ld} vs {g["ldkl"] + g["kl"]})')
    assert e_z < (1e-4 if mode == 'tf32' else 5e-3) and e_ld < (1e-3 if mode == 'tf32' else 5e-2)
    for p in core.parameters():
        p.grad = None
    bw = FlowBackward(eng)
    gx, gh = bw.run(cu(g['Gz']), cu(g['cl']))
    torch.cuda.synchronize()
    worst = 0.0
    for k, p in core.named_parameters():
        if not k.startswith('generator.') or k.endswith('lamb') or k.endswith('geom_p'):
            continue
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu().numpy(), g['grad.' + k])
        worst = max(worst, e)
        print(f'   grad {k}: rel-L2 {e:.2e}')
    e_h = rel_l2(gh.cpu().numpy(), g['gh'] - g['gh_kl'])
    print(f'{mode}: worst parameter-gradient rel-L2 {worst:.2e}; d L / d h (blocks) rel-L2 {e_h:.2e}')
    assert worst < tol and e_h < tol


@pytest.mark.parametrize("mode,tol", [('tf32', 2e-3), ('bf16', 6e-2)])
def test_kl_and_posterior_head_gradients_match_reference(mode, tol):
    """Stage B: from the (training-mode) encoder output on — weight-normed fc, reparameterisation, prior-flow KL — forward values,
    then d L / d (fc output), d L / d (encoder output) and every prior-flow / fc parameter gradient."""
    import ctypes
    from indm_b200 import _lib as L
    from indm_b200.flow_models.wolf_backward import PosteriorBackward
    g, cfg, flow = _setup(mode)
    core = flow.module
    N = g['x'].shape[0]
    eng = core.engine(N)
    eng._ensure()
    if not hasattr(eng, 'enc'):
        eng._build_encoder()
        for job in eng.enc['jobs']:
            job()
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    enc_out, eps = cu(g['enc_out'].reshape(N, -1)), cu(g['eps_post'])
    c = torch.empty((N, 128), device='cuda')
    L.call('indm_linear_f32', L.ptr(enc_out), L.ptr(eng.enc['fc_w']), L.ptr(eng.enc['fc_b']), L.ptr(c), N, enc_out.shape[1], 128, 0, 0, L.DTYPE_F32)
    h, logq = torch.empty((N, 64), device='cuda'), torch.empty((N,), device='cuda')
    L.call('indm_posterior_sample', L.ptr(c), L.ptr(eps), L.ptr(h), L.ptr(logq), N)
    _, kl = eng.prior_flow(h, 'forward', kl_base=logq)
    print(f'fc out rel-L2 {rel_l2(c.cpu().numpy(), g["fc_out"]):.2e}; h rel-L2 {rel_l2(h.cpu().numpy(), g["h"]):.2e}; '
          f'KL max-abs err {float(np.abs(kl.cpu().numpy() - g["kl"]).max()):.2e}')
    assert rel_l2(h.cpu().numpy(), g['h']) < 1e-5 and float(np.abs(kl.cpu().numpy() - g['kl']).max()) < 1e-3
    nblk = len(oflow.block_layout(cfg))
    varepss = [cu(g[f'vareps_{i}']) for i in range(nblk)]
    eng.forward_logdet(cu(g['x']), h, vareps=varepss, n_terms=g['ns'], training=True, save=True)
    for p in core.parameters():
        p.grad = None
    _, gh_blocks = FlowBackward(eng).run(cu(g['Gz']), cu(g['cl']))
    pb = PosteriorBackward(eng)
    g_enc = pb.run(h, gh_blocks, -cu(g['cl']), c, eps, enc_out)
    torch.cuda.synchronize()
    e_c = rel_l2(pb.gc.cpu().numpy(), g['g_fc_out'])
    e_e = rel_l2(g_enc.cpu().numpy(), g['g_enc_out'].reshape(N, -1))
    worst, worst_k = 0.0, ''
    for k, p in core.named_parameters():
        if not (k.startswith('discriminator.prior') or k.startswith('discriminator.fc')):
            continue
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu().numpy(), g['grad.' + k])
        if e > worst:
            worst, worst_k = e, k
    print(f'{mode}: d/d fc_out rel-L2 {e_c:.2e}; d/d enc_out rel-L2 {e_e:.2e}; worst prior / fc parameter gradient {worst:.2e} ({worst_k})')
    assert e_c < tol and e_e < tol and worst < tol


@pytest.mark.parametrize("mode,tol_f,tol", [('tf32', 1e-4, 2e-3), ('bf16', 2e-2, 8e-2)])
def test_training_mode_encoder_forward_backward(mode, tol_f, tol):
    """Stage C: posterior encoder with batch-statistics BatchNorm — forward output, BatchNorm running buffers after the step,
    and every encoder parameter gradient given the reference's cotangent at the encoder output."""
    from indm_b200.flow_models.wolf_encoder_train import EncoderTrain
    g, cfg, flow = _setup(mode)
    core = flow.module
    N = g['x'].shape[0]
    eng = core.engine(N)
    eng._ensure()
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    enc = EncoderTrain(eng)
    for p in core.parameters():
        p.grad = None
    out = enc.forward(cu(g['x']))
    e_f = rel_l2(out.cpu().numpy(), g['enc_out'].reshape(N, -1))
    enc.backward(cu(g['g_enc_out'].reshape(N, -1)))
    torch.cuda.synchronize()
    worst, worst_k = 0.0, ''
    for k, p in core.named_parameters():
        if not k.startswith('discriminator.encoder'):
            continue
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu().numpy(), g['grad.' + k])
        if e > worst:
            worst, worst_k = e, k
    e_b = max(rel_l2(b.cpu().numpy(), g['buf.' + k]) for k, b in core.named_buffers()
              if k.startswith('discriminator.encoder') and not k.endswith('num_batches_tracked'))
    print(f'{mode}: encoder output rel-L2 {e_f:.2e}; worst parameter gradient {worst:.2e} ({worst_k}); running buffers {e_b:.2e}')
    assert e_f < tol_f and worst < tol and e_b < tol_f
