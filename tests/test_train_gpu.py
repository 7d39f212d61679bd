"""GPU: one optimisation step of the score network through losses.get_step_fn (flow.model='identity') against the live
reference (tests/golden/train_tiny_vp.npz from tests/golden/make_golden.py:make_train): per-sample losses, the gradient of
every parameter (norms for all 100+ tensors, sampled elements for one tensor of each kind), and the parameters / EMA after
global-norm clip + AdamW."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, tiny, rel_l2  # noqa: E402
from indm_b200 import configs, sde_lib, losses  # noqa: E402
from indm_b200.models import utils as mutils  # noqa: E402
from indm_b200.models.ema import ExponentialMovingAverage  # noqa: E402
from oracle import ncsnpp as oncsnpp  # noqa: E402


def _sub(a, limit=4096):
    f = np.ascontiguousarray(a).reshape(-1)
    return f[::max(1, (f.size + limit - 1) // limit)]


def _setup(mode, tag='tiny_vp'):
    g = load_npz(f'train_{tag}.npz')
    cfg = configs.get_config('vp/CIFAR10/indm_fid' if tag == 'tiny_vp' else 've/CIFAR10/indm')
    tiny(cfg)
    cfg.model.dropout = 0.0
    cfg.flow.model = 'identity'
    cfg.training.importance_sampling = True
    cfg.device = torch.device('cuda:0')
    model = mutils.create_model(cfg)
    model.load_state_dict({'module.' + k: torch.from_numpy(v) for k, v in oncsnpp.synth_params(cfg, int(g['seed'])).items()})
    model.module.compute_mode = mode
    return g, cfg, model, sde_lib.get_sde(cfg)


@pytest.mark.parametrize("tag", ['tiny_vp', 'tiny_ve'])
@pytest.mark.parametrize("mode,tol", [('tf32', 2e-3), ('bf16', 5e-2)])
def test_parameter_gradients_match_reference(mode, tol, tag):
    g, cfg, model, sde = _setup(mode, tag)
    model.train()
    opt = losses.get_optimizer(cfg, model.parameters())          # re-homes parameters / gradients into flat storage
    opt.zero_grad()
    loss_fn = losses.get_sde_loss_fn(cfg, sde, train=True)
    cu = lambda a: torch.from_numpy(a).cuda()
    ls = loss_fn(model, cu(g['batch']), draws=dict(u=cu(g['u']), z=cu(g['z'])))
    torch.mean(ls).backward()
    torch.cuda.synchronize()
    e_l = float(np.abs(ls.detach().cpu().numpy() - g['losses_raw']).max() / np.abs(g['losses_raw']).max())
    named = dict(model.named_parameters())
    names = [str(n) for n in g['names']]
    got_norm = np.array([float(named[n].grad.norm()) for n in names])
    rel = np.abs(got_norm - g['grad_norms']) / (g['grad_norms'] + 1e-3 * g['grad_norms'].max())
    worst = int(np.argmax(rel))
    tot = float(np.sqrt((got_norm ** 2).sum()))
    print(f'{mode}: losses rel {e_l:.2e}; total grad norm {tot:.4f} ref {float(g["total_norm"]):.4f}; worst per-tensor norm rel '
          f'{rel[worst]:.2e} at {names[worst]}')
    assert e_l < tol
    assert abs(tot - float(g['total_norm'])) / float(g['total_norm']) < tol
    assert rel.max() < 5 * tol, (names[worst], got_norm[worst], g['grad_norms'][worst])
    for n in [str(k) for k in g['keep']]:
        e = rel_l2(_sub(named[n].grad.detach().cpu().numpy()), g['grad::' + n])
        print(f'   grad {n}: rel-L2 {e:.2e}')
        assert e < 5 * tol, n


@pytest.mark.parametrize("mode,tol,step_tol,ema_tol", [('tf32', 2e-3, 2e-2, 1e-4), ('bf16', 5e-2, 0.3, 5e-4)])
def test_step_fn_updates_match_reference(mode, tol, step_tol, ema_tol):
    g, cfg, model, sde = _setup(mode)
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    state = dict(optimizer=opt, model=model, ema=ema, step=0)
    step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    cu = lambda a: torch.from_numpy(a).cuda()
    res = step_fn(state, None, cu(g['batch']), draws=dict(u=cu(g['u']), z=cu(g['z'])))
    torch.cuda.synchronize()
    assert state['step'] == 1 and len(res) == 5
    assert float(np.abs(res[0].numpy() - g['losses_step']).max() / np.abs(g['losses_step']).max()) < tol
    named = dict(model.named_parameters())
    idx = {n: i for i, (n, p) in enumerate(model.named_parameters())}
    for n in [str(k) for k in g['keep']]:
        ref_delta = g['param::' + n] - _sub(before[n].cpu().numpy())
        got_delta = _sub(named[n].detach().cpu().numpy()) - _sub(before[n].cpu().numpy())
        # the first AdamW step moves every weight by ~lr * sign(grad): compare the update, not the (dominant) old value.  Elements
        # whose gradient is ~0 can flip sign under rounding-level gradient differences, each contributing 2 lr to the error.
        e = float(np.linalg.norm(got_delta - ref_delta) / max(np.linalg.norm(ref_delta), 1e-30))
        e_ema = float(np.abs(_sub(ema.shadow_params[idx[n]].cpu().numpy()) - g['ema::' + n]).max())
        print(f'   update {n}: rel-L2 of the step {e:.2e}; ema max-abs err {e_ema:.2e}')
        assert e < step_tol, n
        assert e_ema < ema_tol


def test_fused_adamw_state_dict_interoperates_with_torch_adamw():
    """Checkpoint wire format (reference utils.py:37-43 saves `optimizer.state_dict()` of torch.optim.AdamW): FusedAdamW emits and
    accepts that layout, so optimisation continues identically after moving the state in either direction.  The torch.optim.AdamW
    side runs on the CPU (state dicts are deep-copied across, as `torch.save` / `torch.load` would)."""
    import copy
    from indm_b200.losses import FusedAdamW
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    shapes = [(8, 4, 3, 3), (8,), (16, 8), (16,)]
    init = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(3)]
    kw = dict(lr=2e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)

    def step(opt, ps, gs):
        for p, gr in zip(ps, gs):
            if p.grad is None:
                p.grad = gr.to(p.device).clone()
            else:
                p.grad.copy_(gr.to(p.device))
        opt.step()

    def to_cpu(sd):
        sd = copy.deepcopy(sd)
        for st in sd['state'].values():
            for k, v in st.items():
                if torch.is_tensor(v):
                    st[k] = v.detach().cpu().clone()
        return sd

    pf = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
    pt = [torch.nn.Parameter(t.clone()) for t in init]
    fo, to = FusedAdamW(pf, **kw), torch.optim.AdamW(pt, **kw)
    for k in range(2):
        step(fo, pf, grads[k])
        step(to, pt, grads[k])
    torch.cuda.synchronize()
    sd_f, sd_t = to_cpu(fo.state_dict()), to_cpu(to.state_dict())
    assert sorted(sd_f.keys()) == ['param_groups', 'state'] and sorted(sd_f['state'].keys()) == [0, 1, 2, 3]
    assert sd_f['param_groups'][0]['params'] == [0, 1, 2, 3] and float(sd_f['state'][0]['step']) == 2.0
    for i in range(4):
        assert torch.allclose(pf[i].detach().cpu(), pt[i].detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(sd_f['state'][i]['exp_avg'], sd_t['state'][i]['exp_avg'], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sd_f['state'][i]['exp_avg_sq'], sd_t['state'][i]['exp_avg_sq'], rtol=1e-5, atol=1e-9)
    # fused -> torch and torch -> fused, then one more step everywhere
    pt2 = [torch.nn.Parameter(p.detach().cpu().clone()) for p in pf]
    to2 = torch.optim.AdamW(pt2, lr=1.0)
    to2.load_state_dict(sd_f)
    pf2 = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in pt]
    fo2 = FusedAdamW(pf2, lr=1.0)
    fo2.load_state_dict(sd_t)
    assert fo2.steps == 2 and fo2.param_groups[0]['lr'] == kw['lr'] and tuple(fo2.param_groups[0]['betas']) == (0.9, 0.99)
    for opt, ps in ((fo, pf), (to, pt), (to2, pt2), (fo2, pf2)):
        step(opt, ps, grads[2])
    torch.cuda.synchronize()
    for i in range(4):
        for name, other in (('fused', pf), ('torch<-fused', pt2), ('fused<-torch', pf2)):
            d = float((other[i].detach().cpu() - pt[i].detach()).abs().max())
            assert d < 2e-6, (name, i, d)
