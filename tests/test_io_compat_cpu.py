"""Host-side format compatibility: checkpoint dict layout (reference utils.py:14-48) incl. the
torch.optim.AdamW state-dict layout, npz sample cache (sampling_lib.py:30-110), dequantisation + scalers (run_lib.py:85-87,
datasets.py:56-71) and the get_bpd loop (evaluation.py:388-495).  CPU only: the functions are generic over anything with
state_dict / load_state_dict, so torch's own AdamW and a toy module stand in for the CUDA-only classes."""
import os

import numpy as np
import pytest
import torch

from indm_b200 import configs, datasets, evaluation, sampling_lib, utils
from indm_b200.losses import pack_adamw_state, unpack_adamw_state


def _toy_params():
    g = torch.Generator().manual_seed(0)
    return [torch.nn.Parameter(torch.randn(s, generator=g)) for s in [(4, 3, 3, 3), (4,), (5, 4), (5,)]]


def _stepped_adamw(params, n=3):
    opt = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for _ in range(n):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    return opt


def test_adamw_state_layout_round_trip_with_torch_adamw():
    params = _toy_params()
    opt = _stepped_adamw(params)
    sd = opt.state_dict()
    total = sum(p.numel() for p in params)
    m, v = torch.zeros(total), torch.zeros(total)
    groups = [dict(lr=1.0, betas=(0., 0.), eps=0., weight_decay=0., params=params)]
    steps = unpack_adamw_state(sd, params, m, v, groups)
    assert steps == 3 and groups[0]['lr'] == 2e-4 and tuple(groups[0]['betas']) == (0.9, 0.99)
    off = 0
    for i, p in enumerate(params):
        assert torch.equal(m[off:off + p.numel()].view_as(p), sd['state'][i]['exp_avg'])
        assert torch.equal(v[off:off + p.numel()].view_as(p), sd['state'][i]['exp_avg_sq'])
        off += p.numel()
    # and back: what FusedAdamW.state_dict() emits is accepted by torch.optim.AdamW and continues identically
    packed = pack_adamw_state(params, steps, m, v, groups)
    assert sorted(packed['state'].keys()) == [0, 1, 2, 3] and packed['param_groups'][0]['params'] == [0, 1, 2, 3]
    p2 = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt2 = torch.optim.AdamW(p2, lr=1.0)
    opt2.load_state_dict(packed)
    gs = [torch.randn(p.shape, generator=torch.Generator().manual_seed(7 + i)) for i, p in enumerate(params)]
    for p, q, g in zip(params, p2, gs):
        p.grad, q.grad = g.clone(), g.clone()
    opt.step()
    opt2.step()
    for p, q in zip(params, p2):
        assert torch.allclose(p, q, rtol=0, atol=1e-7)


def test_adamw_state_accepts_int_steps_and_fresh_optimizer_and_rejects_mismatch():
    params = _toy_params()
    sd = _stepped_adamw(params).state_dict()
    for st in sd['state'].values():
        st['step'] = 3                                   # torch 1.7.1 (the reference's pin) stores a python int
    total = sum(p.numel() for p in params)
    m, v = torch.ones(total), torch.ones(total)
    groups = [dict(lr=1.0, params=params)]
    assert unpack_adamw_state(sd, params, m, v, groups) == 3
    fresh = torch.optim.AdamW(_toy_params(), lr=1e-3).state_dict()        # saved before the first step: empty state
    assert unpack_adamw_state(fresh, params, m, v, groups) == 0 and float(m.abs().sum()) == 0 and float(v.abs().sum()) == 0
    assert pack_adamw_state(params, 0, m, v, groups)['state'] == {}
    with pytest.raises(ValueError):
        unpack_adamw_state(sd, params[:-1], m, v, [dict(lr=1.0, params=params[:-1])])


class _Ema:
    def __init__(self, params):
        self.decay, self.num_updates, self.shadow_params = 0.9999, 0, [p.detach().clone() for p in params]

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, sd):
        self.decay, self.num_updates = sd['decay'], sd['num_updates']
        self.shadow_params = [s.clone() for s in sd['shadow_params']]


def test_checkpoint_save_restore_layout(tmp_path):
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    model = torch.nn.DataParallel(torch.nn.Conv2d(3, 4, 3))        # 'module.'-prefixed keys, as the reference saves them
    opt = _stepped_adamw(list(model.parameters()), n=2)
    ema = _Ema(model.parameters())
    ema.num_updates = 2
    state = dict(optimizer=opt, model=model, ema=ema, step=2)
    path = str(tmp_path / "checkpoints" / utils.create_name("checkpoint", 12, "pth"))
    assert path.endswith("checkpoint_12.pth")
    # restoring from a missing file returns the state unchanged and creates the directory (utils.py:15-19)
    assert utils.restore_checkpoint(cfg, path, state, "cpu") is state and os.path.isdir(os.path.dirname(path))
    utils.save_checkpoint(cfg, path, state)
    raw = torch.load(path, weights_only=False)
    assert sorted(raw.keys()) == ['ema', 'model', 'optimizer', 'step']
    assert sorted(raw['model'].keys()) == ['module.bias', 'module.weight']
    assert sorted(raw['ema'].keys()) == ['decay', 'num_updates', 'shadow_params']
    model2 = torch.nn.DataParallel(torch.nn.Conv2d(3, 4, 3))
    state2 = dict(optimizer=torch.optim.AdamW(model2.parameters(), lr=1.0), model=model2, ema=_Ema(model2.parameters()), step=0)
    utils.restore_checkpoint(cfg, path, state2, "cpu")
    assert state2['step'] == 2 and state2['ema'].num_updates == 2
    assert torch.equal(model2.module.weight, model.module.weight)
    assert state2['optimizer'].param_groups[0]['lr'] == 2e-4
    # VE-SDE runs keep their optimizer (utils.py:23-24)
    cfg_ve = configs.get_config("ve/CIFAR10/indm")
    opt3 = torch.optim.AdamW(torch.nn.Conv2d(3, 4, 3).parameters(), lr=1.0)
    state3 = dict(optimizer=opt3, model=torch.nn.DataParallel(torch.nn.Conv2d(3, 4, 3)), ema=_Ema(model2.parameters()), step=0)
    utils.restore_checkpoint(cfg_ve, path, state3, "cpu")
    assert opt3.param_groups[0]['lr'] == 1.0 and state3['step'] == 2


def test_create_name_variants():
    assert utils.create_name("checkpoint", "7", "pth") == "checkpoint_7.pth"
    assert utils.create_name("flow_checkpoint", "best", "pth") == "flow_checkpoint_best.pth"
    assert utils.create_name("checkpoint", "a/b/ckpt_3.pth", "pth") == "checkpoint_ckpt_3.pth"


def test_scalers_and_dequantisation():
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    assert cfg.data.centered
    x = torch.tensor([0., 0.5, 1.])
    assert torch.equal(datasets.get_data_scaler(cfg)(x), torch.tensor([-1., 0., 1.]))
    assert torch.equal(datasets.get_data_inverse_scaler(cfg)(datasets.get_data_scaler(cfg)(x)), x)
    ve = configs.get_config("ve/CIFAR10/indm")
    assert not ve.data.centered and torch.equal(datasets.get_data_scaler(ve)(x), x)
    img = torch.randint(0, 256, (2, 3, 4, 4)).float() / 255.
    u = torch.rand(img.shape)
    d = datasets.dequantize(img, u)
    assert torch.equal(d, (255. * img + u) / 256.) and float(d.min()) >= 0 and float(d.max()) < 1
    assert torch.equal(torch.floor(d * 256.), torch.round(img * 255.))       # the 8-bit value is recoverable


def test_sample_cache_layout(tmp_path):
    cfg = configs.get_config("vp/CIFAR10/indm_fid")
    cfg.data.image_size = 8
    calls = []
    g = torch.Generator().manual_seed(0)
    before = torch.rand(4, 3, 8, 8, generator=g) * 1.2 - 0.1
    after = torch.rand(4, 3, 8, 8, generator=g) * 1.2 - 0.1

    def sampling_fn(score_model, flow_model, temperature, data_mean, sample_dir=None, r=None):
        calls.append(r)
        return before, after, 1000

    sd, td = str(tmp_path / "samples"), str(tmp_path / "samples" / "ckpt_1")
    out = sampling_lib.get_samples(cfg, None, None, sampling_fn, 1, 0, sd, this_sample_dir=td)
    b = np.load(os.path.join(sd, "samples_0_before_flow.npz"))["samples"]
    a = np.load(os.path.join(td, "samples_0.npz"))["samples"]
    assert b.shape == (4, 8, 8, 3) and a.shape == (4, 8, 8, 3) and a.dtype == np.uint8
    np.testing.assert_array_equal(b, before.permute(0, 2, 3, 1).numpy() * 255.)            # unclipped, unrounded
    np.testing.assert_array_equal(a, np.clip(after.permute(0, 2, 3, 1).numpy() * 255., 0, 255).astype(np.uint8))
    np.testing.assert_array_equal(out, a)
    # cached round: the sampler is not called again
    out2 = sampling_lib.get_samples(cfg, None, None, sampling_fn, 1, 0, sd, this_sample_dir=td)
    assert calls == [0]
    np.testing.assert_array_equal(out2, a)


def test_get_bpd_loop_counts_and_means():
    cfg = configs.get_config("vp/CIFAR10/indm_nll")
    cfg.flow.model = "identity"
    cfg.device = torch.device("cpu")
    cfg.data.image_size = 4
    cfg.eval.batch_size = 500
    cfg.eval.num_nelbo = 2
    ds = [torch.full((500, 3, 4, 4), i / 255.) for i in range(4)]          # 4 batches per epoch: the loop must wrap around
    seen = dict(nelbo=0, nll=[])

    def nelbo_fn(model, flow, batch, logdet):
        assert float(batch.min()) >= -1 and float(batch.max()) < 1           # dequantised, then scaled to [-1, 1)
        seen['nelbo'] += 1
        return torch.full((batch.shape[0],), 3.0), torch.full((batch.shape[0],), 2.5)

    def nll_fn(model, flow, batch, logdet, residual=True, eps_bpd=1e-5):
        seen['nll'].append((residual, eps_bpd, batch.shape[0]))
        return torch.full((batch.shape[0],), 2.0 if residual else 2.2), None, 100

    res = evaluation.get_bpd(cfg, ds, datasets.get_data_scaler(cfg), nelbo_fn, nll_fn, None, None, step=5, eval=False)
    assert seen['nelbo'] == 2 * 20                                          # 10000 samples / 500 per batch, num_nelbo passes
    n_nll = (1000 - 1) // 500 + 1                                           # NLL on num_data // 10 outside eval mode
    assert seen['nll'][:n_nll] == [(False, 1e-5, 500)] * n_nll and seen['nll'][n_nll:2 * n_nll] == [(True, 1e-5, 500)] * n_nll
    assert res['nelbo'] == 3.0 and res['nelbo_residual'] == 2.5
    assert res['nll'] == pytest.approx(2.0) and res['nll_wrong'] == pytest.approx(2.2)
    tt = cfg.training.truncation_time
    assert (res['nll_train_eps'] is None) == (tt == 1e-5)
    if tt != 1e-5:
        assert seen['nll'][-1] == (True, tt, 500)


@pytest.mark.parametrize("name", ["vp/CIFAR10/indm_nll", "vp/CIFAR10/indm_fid", "ve/CIFAR10/indm"])
def test_get_loss_fns_builds_the_four_step_callables(name):
    """reference utils.py:132-140"""
    from indm_b200 import sde_lib
    cfg = configs.get_config(name)
    cfg.device = torch.device("cpu")
    sde = sde_lib.get_sde(cfg)
    fns = utils.get_loss_fns(cfg, sde, datasets.get_data_inverse_scaler(cfg), scaler=datasets.get_data_scaler(cfg))
    assert len(fns) == 4 and all(callable(f) for f in fns)


def test_sample_cache_pc_denoise_branch(tmp_path):
    """sampling_lib.py:60-105: with sampling.pc_denoise the cached latent-side samples are denoised by a second sampler call
    (`final_time=`, `before_data=scaler(cached)`), cached under `*_denoise_{time}.npz`; a cached denoised latent is only pushed
    through the flow inverse again."""
    cfg = configs.get_config("vp/CIFAR10/indm_fid")
    cfg.data.image_size = 8
    cfg.flow.model = "identity"
    cfg.device = torch.device("cpu")
    cfg.sampling.pc_denoise = True
    cfg.sampling.pc_denoise_time = 0.001
    scaler, inv = datasets.get_data_scaler(cfg), datasets.get_data_inverse_scaler(cfg)
    g = torch.Generator().manual_seed(0)
    before = torch.rand(4, 3, 8, 8, generator=g)
    after = torch.rand(4, 3, 8, 8, generator=g)
    calls = []

    def sampling_fn(score_model, flow_model, temperature, data_mean, sample_dir=None, r=None, final_time=0., before_data=None):
        calls.append((final_time, None if before_data is None else before_data.clone()))
        if before_data is None:
            return before, after, 1000
        return inv(before_data) * 0.5, inv(before_data) * 0.25, 10      # a recognisable "denoised" result

    sd, td = str(tmp_path / "s"), str(tmp_path / "s" / "ckpt")
    out = sampling_lib.get_samples(cfg, None, None, sampling_fn, 1, 3, sd, inverse_scaler=inv, this_sample_dir=td, scaler=scaler)
    assert [c[0] for c in calls] == [0., 0.001]
    cached = np.load(os.path.join(sd, "samples_3_before_flow.npz"))["samples"]
    want_in = scaler(torch.tensor(cached).permute(0, 3, 1, 2) / 255.)
    assert torch.allclose(calls[1][1].double(), want_in.double(), atol=1e-6)
    den = np.load(os.path.join(td, "samples_3_denoise_0.001.npz"))["samples"]
    np.testing.assert_array_equal(out, den)
    assert np.abs(den.astype(np.int64) - np.clip(cached * 0.25, 0., 255.).astype(np.uint8).astype(np.int64)).max() <= 1
    assert den.dtype == np.uint8 and den.shape == (4, 8, 8, 3)
    assert os.path.exists(os.path.join(sd, "samples_3_before_flow_denoise_0.001.npz"))
    # second call: everything cached, the sampler is not invoked
    out2 = sampling_lib.get_samples(cfg, None, None, sampling_fn, 1, 3, sd, inverse_scaler=inv, this_sample_dir=td, scaler=scaler)
    assert len(calls) == 2
    np.testing.assert_array_equal(out2, den)
    # denoised latent cached but final file missing: only the (identity) flow inverse runs
    os.remove(os.path.join(td, "samples_3_denoise_0.001.npz"))
    out3 = sampling_lib.get_samples(cfg, None, None, sampling_fn, 1, 3, sd, inverse_scaler=inv, this_sample_dir=td, scaler=scaler)
    assert len(calls) == 2
    lat = np.load(os.path.join(sd, "samples_3_before_flow_denoise_0.001.npz"))["samples"]
    np.testing.assert_array_equal(out3, np.clip(lat, 0., 255.).astype(np.uint8))
