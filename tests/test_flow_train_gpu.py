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


def _sub(a, limit=4096):
    f = np.ascontiguousarray(a).reshape(-1)
    return f[::max(1, (f.size + limit - 1) // limit)].copy()


@pytest.mark.parametrize("mode,tol,step_tol,ema_tol", [('tf32', 2e-3, 3e-2, 1e-4), ('bf16', 5e-2, 0.4, 5e-4)])
def test_joint_flow_score_step_matches_reference(mode, tol, step_tol, ema_tol):
    """The reference's flow_step_fn_nll (losses.py:258-320) end to end: flow forward in training mode -> latent -> score loss +
    flow loss + prior log-p -> one backward through both networks -> clip + AdamW + EMA on both.  Loss vectors and the applied
    parameter updates of both networks against the live reference (first AdamW step ~ lr * sign(grad): the update is compared)."""
    from indm_b200 import losses, sde_lib
    from indm_b200.models import utils as mutils
    from indm_b200.models.ema import ExponentialMovingAverage
    from oracle import ncsnpp as oncsnpp
    g = load_npz('jointtrain_small_vp.npz')
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 128, (1, 2), 1, (16,)
    cfg.flow.nblocks, cfg.flow.intermediate_dim = '2-2', 128
    cfg.model.dropout = 0.0
    cfg.device = torch.device('cuda:0')
    model = mutils.create_model(cfg)
    model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, int(g['seed_score'])).items()})
    flow = fm.create_flow_model(cfg)
    flow.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oflow.synth_params(cfg, int(g['seed_flow'])).items()})
    model.module.compute_mode = flow.module.compute_mode = mode
    model.train()
    sde = sde_lib.get_sde(cfg)
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    state = dict(optimizer=opt, model=model, ema=ema, step=0)
    fopt = losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr)
    fema = ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate)
    flow_state = dict(optimizer=fopt, model=flow, ema=fema, step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    nets = {'s': (model, ema), 'f': (flow, fema)}
    before = {f'{t}::{n}': p.detach().clone() for t, (net, _) in nets.items() for n, p in net.named_parameters()}
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nblk = len(oflow.block_layout(cfg))
    flow_kw = dict(eps=cu(g['eps_post']), vareps=[cu(g[f'vareps_{i}']) for i in range(nblk)], n_terms=g['ns'])
    res = step_fn(state, flow_state, cu(g['batch']), draws=dict(u=cu(g['u']), z=cu(g['z'])), flow_kw=flow_kw, logp_noise=cu(g['logp_noise']))
    torch.cuda.synchronize()
    assert state['step'] == 1 and flow_state['step'] == 1
    for i, key in enumerate(('losses', 'losses_score', 'losses_flow', 'losses_logp')):
        e = float(np.abs(res[i].numpy() - g[key]).max() / np.abs(g[key]).max())
        print(f'{mode}: {key} rel err {e:.2e}  ({res[i].numpy()} vs {g[key]})')
        assert e < tol, key
    for key in [str(k) for k in g['names']]:
        t, n = key.split('::')
        net, em = nets[t]
        plist = [(nn_, p) for nn_, p in net.named_parameters() if p.requires_grad]
        idx = [nn_ for nn_, _ in plist].index(n)
        p = plist[idx][1]
        got = _sub(p.detach().cpu().numpy()) - _sub(before[key].cpu().numpy())
        ref = g['step::' + key]
        e = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))
        e_ema = float(np.abs(_sub(em.shadow_params[idx].cpu().numpy()) - g['ema::' + key]).max())
        print(f'   update {key}: rel-L2 of the step {e:.2e}; ema max-abs err {e_ema:.2e}')
        assert e < step_tol, key
        # one sign flip of a ~zero-gradient element moves the parameter by 2 lr, and the EMA by 0.82 of that on the first update
        # (models/ema.py:40: decay = min(decay, (1 + n) / (10 + n)) = 2 / 11); the rel-L2 of the step above bounds how many flip
        lr_ = cfg.optim.lr if t == 's' else cfg.flow.lr
        assert e_ema < max(ema_tol, 2.0 * lr_), key


def test_training_after_sampling_on_the_same_flow_engines_at_batch_128():
    """Regression (round 2): sampling (flow reverse on the TF32 engine) followed, in the same process, by joint training steps — first
    with the default policy (BF16 blocks, TF32 encoder legs on that same engine), then with the TF32-blocks policy.  With one
    memory-pool handle shared by an engine's CUDA graphs this sequence faulted (illegal address / caching-allocator assert) at batch
    128 once graphs had been dropped; every graph owns its pool now.  Small score network: the flow is what is exercised."""
    from indm_b200 import losses, sde_lib, precision
    from indm_b200.models import utils as mutils
    from indm_b200.models.ema import ExponentialMovingAverage
    B = 128
    cfg = configs.get_config('vp/CIFAR10/indm_nll')
    cfg.model.nf, cfg.model.ch_mult, cfg.model.num_res_blocks, cfg.model.attn_resolutions = 128, (1, 2), 1, (16,)
    cfg.device = torch.device('cuda:0')
    torch.manual_seed(0)
    model = mutils.create_model(cfg)
    flow = fm.create_flow_model(cfg)
    flow.eval()
    sde = sde_lib.get_sde(cfg)
    z = torch.randn(B, 3, 32, 32, device='cuda')
    for _ in range(2):
        x, _ = fm.flow_forward(cfg, flow, z, log_det=None, reverse=True)
    assert torch.isfinite(x).all()
    state = dict(optimizer=losses.get_optimizer(cfg, model.parameters()), model=model,
                 ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
    flow_state = dict(optimizer=losses.get_optimizer(cfg, flow.parameters(), lr=cfg.flow.lr), model=flow,
                      ema=ExponentialMovingAverage(flow.parameters(), decay=cfg.flow.ema_rate), step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    batch = torch.rand(B, 3, 32, 32, device='cuda') * 2 - 1
    try:
        for _ in range(3):
            res = step_fn(state, flow_state, batch)
            assert torch.isfinite(res[0]).all()
        precision.set_policy('flow', 'training', 'tf32')
        for _ in range(3):
            res = step_fn(state, flow_state, batch)
            assert torch.isfinite(res[0]).all()
        flow.eval()
        x, _ = fm.flow_forward(cfg, flow, z, log_det=None, reverse=True)      # and back to sampling on the updated weights
        torch.cuda.synchronize()
        assert torch.isfinite(x).all()
    finally:
        precision.set_policy('flow', 'training', 'bf16')
