"""Training losses and step functions with the reference's interface (losses.py): `get_optimizer` :30-45,
`optimization_manager` :48-62, `get_sde_loss_fn` :65-144, `get_step_fn` :194-420.

What runs where: the score network's forward (train mode) and its full backward — input gradient and every parameter gradient —
run on the engine's explicit plans (models/engine.py); the DSM loss and its gradient are one kernel; gradient-norm clipping +
AdamW run as one pass over flat parameter storage; under torch.distributed the flat gradient buffer is all-reduced with NCCL
before the clip (data-parallel training, one process per GPU, where the reference uses nn.DataParallel).

Scope of this module: `step_fn` (score network only), `flow_step_fn_nll` and `flow_step_fn_fid` (JOINT flow + score steps,
both phases of the FID variant) are complete: the flow's gradients — first order through g and the posterior encoder, second
order through the Neumann log-det estimator (iresblock.py:264-273) — come from the explicit backward plans of
flow_models/wolf_backward.py, reached through torch.autograd.  Joint training is the default, like the reference;
`config.training.freeze_flow=True` opts into a variant that keeps the flow parameters fixed.
"""
import logging
import os

import numpy as np
import torch

from . import _lib as L
from .flow_models.flow_model import flow_forward
from .models import utils as mutils


# ------------------------------------------------------------------------------------------------ optimizer
def _stepped(p):
    """True for the parameters torch.optim.AdamW would update: trainable AND reached by the loss.  torch skips a parameter whose
    `.grad` is None entirely — no moment update and NO WEIGHT DECAY (torch/optim/adamw.py) — which is what happens to the
    iResBlocks' `lamb` / `geom_p` in the reference (nn.Parameters no loss term touches, iresblock.py:36-42).  With flat storage
    every slot has a gradient, so those parameters are marked `_indm_no_grad` by their module: they stay inside the flat buffers
    (the parameters must tile one contiguous buffer for the fused EMA pass) and `FusedAdamW.step` restores their values."""
    return p.requires_grad and not getattr(p, '_indm_no_grad', False)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, amsgrad off) as ONE kernel over flat storage,
    with `clip_grad_norm_` folded in (the clip coefficient is computed on the device from the global gradient norm: no host
    sync) and the data-parallel gradient all-reduce issued on the same flat buffer.  Parameters and their `.grad` are re-homed
    into two flat buffers at construction; `zero_grad` zeroes in place so the engine's backward plan keeps valid pointers."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError('FusedAdamW steps one parameter group (the reference builds one, losses.py:33-41)')
        ps = [p for g in self.param_groups for p in g['params'] if p.requires_grad]
        if not ps or not ps[0].is_cuda:
            raise RuntimeError('indm_b200 FusedAdamW needs CUDA parameters: there is no CPU path')
        if not decoupled:
            raise NotImplementedError("optim.optimizer='Adam' (coupled weight decay) is not used by any INDM config")
        dev, total = ps[0].device, sum(p.numel() for p in ps)
        self._params = ps
        self.flat_p = torch.empty((total,), dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self._sumsq = torch.zeros((1,), dtype=torch.float32, device=dev)
        off, skip = 0, []
        with torch.no_grad():
            for p in ps:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
                p.grad = self.flat_g[off:off + n].view_as(p)
                if not _stepped(p):
                    skip += list(range(off, off + n))
                off += n
        # flat offsets of the parameters torch.optim.AdamW would skip (see _stepped): their values are carried across the kernel
        self._skip_idx = torch.tensor(skip, dtype=torch.int64, device=dev) if skip else None
        self.steps = 0
        self.max_norm = -1.0          # set by optimize_fn (losses.py:58-59); < 0 disables clipping
        self.fused_ema = None         # an ExponentialMovingAverage over the same parameters: its update rides in the AdamW kernel
        L.param_epoch += 1

    def attach_ema(self, ema):
        """The step functions call `optimize_fn(...)` and then `ema.update(params)` (losses.py:311-316): with the EMA attached, the
        AdamW kernel applies shadow -= (1 - decay) (shadow - p) to the freshly updated parameter in the same pass (one read of p less,
        one launch less) and the following `ema.update` only advances its counter."""
        ok = ema is not None and getattr(ema, '_flat', None) is not None and ema._flat.numel() == self.flat_p.numel() and ema._flat.is_cuda
        self.fused_ema = ema if ok else None

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        self._reduced = False

    def start_allreduce(self):
        """Data-parallel exchange of this optimiser's flat gradient buffer, started EARLY on a side stream: the joint step calls it
        when the score network's backward is complete and the flow's backward is about to be queued, so the NCCL all-reduce of the
        62.8 M score gradients runs under the flow backward instead of after it.  `step` then only waits for the side stream."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
            return
        if os.environ.get('INDM_EARLY_ALLREDUCE', '1') == '0':
            return
        if getattr(self, '_comm', None) is None:
            self._comm = torch.cuda.Stream()
        self._comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm):
            torch.distributed.all_reduce(self.flat_g)
            self.flat_g.div_(torch.distributed.get_world_size())
        self._reduced = True

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        if getattr(self, '_reduced', False):
            torch.cuda.current_stream().wait_stream(self._comm)       # the exchange was started early (start_allreduce)
            self._reduced = False
        elif torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.flat_g)                 # NCCL over NVLink: SUM, then the mean
            self.flat_g.div_(torch.distributed.get_world_size())
        self.steps += 1
        sumsq = None
        if self.max_norm >= 0:
            self._sumsq.zero_()
            L.call('indm_sumsq_f32', L.ptr(self.flat_g), self.flat_g.numel(), L.ptr(self._sumsq))
            sumsq = self._sumsq
        b1, b2 = g['betas']
        keep = self.flat_p.index_select(0, self._skip_idx) if self._skip_idx is not None else None
        ema, ema_ptr, decay = self.fused_ema, None, 0.0
        if ema is not None:
            decay = ema.next_decay()
            ema_ptr = L.ptr(ema._flat)
            keep_ema = ema._flat.index_select(0, self._skip_idx) if self._skip_idx is not None else None
        L.call('indm_adamw_ema_f32', L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), ema_ptr,
               self.flat_p.numel(), float(g['lr']), float(b1), float(b2), float(g['eps']), float(g['weight_decay']), self.steps,
               L.ptr(sumsq), float(self.max_norm), float(decay))
        if keep is not None:          # zero gradient, zero moments: the kernel only applied the weight decay; undo it
            self.flat_p.index_copy_(0, self._skip_idx, keep)
            if ema is not None:       # their shadow is their (constant) value: restore what the pass blended with the decayed value
                ema._flat.index_copy_(0, self._skip_idx, keep_ema)
        if ema is not None:
            ema.fused_update_done()
        L.param_epoch += 1            # engines repack their operand copies of the weights on next use

    def state_dict(self):
        """`torch.optim.AdamW.state_dict()` layout — the wire format of the reference's `checkpoint_N.pth['optimizer']`
        (utils.py:37-43 saves `state['optimizer'].state_dict()`): per-parameter `step` / `exp_avg` / `exp_avg_sq` keyed by the
        parameter's index, views into the flat moment buffers."""
        return pack_adamw_state(self._params, self.steps, self.exp_avg, self.exp_avg_sq, self.param_groups)

    def load_state_dict(self, sd):
        """Accepts a `torch.optim.AdamW` state dict (reference checkpoints, utils.py:24) and this class's own."""
        self.steps = unpack_adamw_state(sd, self._params, self.exp_avg, self.exp_avg_sq, self.param_groups)


def pack_adamw_state(params, steps, exp_avg_flat, exp_avg_sq_flat, param_groups):
    """Flat moment buffers -> the dict `torch.optim.AdamW.state_dict()` returns (state empty before the first step, like torch)."""
    state, off = {}, 0
    every = [p for g in param_groups for p in g['params']]      # torch indexes state by position in the FULL group list:
    for i, p in enumerate(every):                               # a frozen parameter keeps its index and simply has no state
        if not p.requires_grad:
            continue
        n = p.numel()
        if steps > 0 and _stepped(p):
            state[i] = dict(step=torch.tensor(float(steps)), exp_avg=exp_avg_flat[off:off + n].view_as(p),
                            exp_avg_sq=exp_avg_sq_flat[off:off + n].view_as(p))
        off += n
    groups, base = [], 0
    for g in param_groups:
        d = {k: v for k, v in g.items() if k != 'params'}
        d.setdefault('amsgrad', False)          # torch 1.7.1's AdamW (the reference's pin) reads it from the loaded group
        d['params'] = list(range(base, base + len(g['params'])))
        base += len(g['params'])
        groups.append(d)
    return dict(state=state, param_groups=groups)


def unpack_adamw_state(sd, params, exp_avg_flat, exp_avg_sq_flat, param_groups):
    """Inverse of `pack_adamw_state`; `step` may be an int (torch 1.7.1, the reference's pin) or a tensor.  Returns the step count.
    Parameters without saved state (a checkpoint written before the first step) get zero moments."""
    if 'state' not in sd:                      # round-1 layout of this class (flat tensors)
        exp_avg_flat.copy_(sd['exp_avg'])
        exp_avg_sq_flat.copy_(sd['exp_avg_sq'])
        for g, s in zip(param_groups, sd['param_groups']):
            g.update({k: v for k, v in s.items() if k != 'params'})
        return int(sd['steps'])
    n_saved = sum(len(g['params']) for g in sd['param_groups'])
    every = [p for g in param_groups for p in g['params']]
    if n_saved != len(every):
        raise ValueError(f"loaded state dict contains a parameter group that doesn't match the size of optimizer's group "
                         f"({n_saved} vs {len(every)} parameters)")
    ids = [i for g in sd['param_groups'] for i in g['params']]
    steps, off = 0, 0
    with torch.no_grad():
        for i, p in zip(ids, every):
            if not p.requires_grad:       # frozen: holds an index, no state (torch.optim.AdamW never steps it)
                continue
            n = p.numel()
            st = sd['state'].get(i) if _stepped(p) else None
            if st is None:
                exp_avg_flat[off:off + n].zero_()
                exp_avg_sq_flat[off:off + n].zero_()
            else:
                if tuple(st['exp_avg'].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state of parameter {i} has shape {tuple(st['exp_avg'].shape)}, expected {tuple(p.shape)}")
                exp_avg_flat[off:off + n].copy_(st['exp_avg'].reshape(-1))
                exp_avg_sq_flat[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
                steps = max(steps, int(float(st['step'])))
            off += n
    for g, s in zip(param_groups, sd['param_groups']):
        g.update({k: v for k, v in s.items() if k != 'params'})
    return steps


def get_optimizer(config, params, lr=None, beta1=None, eps=None, weight_decay=None):
    """losses.py:30-45"""
    if lr is None: lr = config.optim.lr
    if beta1 is None: beta1 = config.optim.beta1
    if eps is None: eps = config.optim.eps
    if weight_decay is None: weight_decay = config.optim.weight_decay
    if config.optim.amsgrad:
        raise NotImplementedError('amsgrad is off in every INDM config')
    if config.optim.optimizer == 'AdamW':
        return FusedAdamW(params, lr=lr, betas=(beta1, 0.99), eps=eps, weight_decay=weight_decay)
    if config.optim.optimizer == 'Adam':
        return FusedAdamW(params, lr=lr, betas=(beta1, 0.999), eps=eps, weight_decay=weight_decay, decoupled=False)
    raise NotImplementedError(f'Optimizer {config.optim.optimizer} not supported yet!')


def optimization_manager(config):
    """Returns an optimize_fn based on `config` (losses.py:48-62)."""

    def optimize_fn(optimizer, params, step, lr=config.optim.lr, warmup=config.optim.warmup, grad_clip=config.optim.grad_clip):
        """Optimizes with warmup and gradient clipping (disabled if negative)."""
        if warmup > 0:
            for g in optimizer.param_groups:
                g['lr'] = lr * np.minimum(step / warmup, 1.0)
        if isinstance(optimizer, FusedAdamW):
            optimizer.max_norm = float(grad_clip)
        elif grad_clip >= 0:
            torch.nn.utils.clip_grad_norm_(params, max_norm=grad_clip)
        optimizer.step()

    return optimize_fn


# ------------------------------------------------------------------------------------------------ loss
class _DSMLoss(torch.autograd.Function):
    """losses[n] = 0.5 * w[n] * norm * sum((score * std[n] + z)^2) and d losses / d score in the same kernel pass"""

    @staticmethod
    def forward(ctx, score, z, std, w, norm):
        N, D = score.shape[0], score[0].numel()
        score, z = score.contiguous().float(), z.contiguous().float()
        std, w = std.contiguous().float(), w.contiguous().float()
        losses = torch.empty((N,), device=score.device)
        dscore = torch.empty_like(score)
        L.call('indm_dsm_loss_f32', L.ptr(score), L.ptr(z), L.ptr(std), L.ptr(w), L.ptr(losses), L.ptr(dscore), N, D, float(norm), 1.0)
        ctx.save_for_backward(dscore)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        dscore, = ctx.saved_tensors
        return dscore * grad_losses.reshape(-1, 1, 1, 1), None, None, None, None


def get_sde_loss_fn(config, sde, train, variance='scoreflow'):
    """losses.py:65-144: `loss_fn(model, batch, st=False, recon_loss=None, importance_sampling=None) -> losses [B]`.
    Keyword-only `draws=` = dict(u=, z=) pins the random draws for parity tests."""
    reduce_mean = bool(config.training.reduce_mean)

    def loss_fn(model, batch, st=False, recon_loss=None, importance_sampling=None, *, draws=None):
        draws = draws or {}
        if recon_loss is None:
            recon_loss = config.training.reconstruction_loss
        if importance_sampling is None:
            importance_sampling = config.training.importance_sampling
        if recon_loss:
            raise NotImplementedError('training.reconstruction_loss=True is not used by any INDM config')
        t_min = sde.get_t_min(config, st)
        if importance_sampling:
            t, Z = sde.get_diffusion_time(config, batch.shape[0], batch.device, t_min, importance_sampling=True, u=draws.get('u'))
        else:
            u = draws.get('u')
            t = (u if u is not None else torch.rand(batch.shape[0], device=batch.device)) * (sde.T - t_min) + t_min
            Z = 1
        score_fn = mutils.get_score_fn(config, sde, model, None, train=train, continuous=config.training.continuous)
        z = draws['z'] if 'z' in draws else torch.randn_like(batch)
        N, D = batch.shape[0], batch[0].numel()
        # x_t = mean + std z with mean = a(t) x (sde.marginal_prob): one fused pass
        a = sde.marginal_prob(torch.ones((N, 1, 1, 1), device=batch.device), t)[0].reshape(N).contiguous().float()
        std = sde.marginal_prob(torch.zeros((N, 1, 1, 1), device=batch.device), t)[1].contiguous().float()
        if batch.requires_grad:
            perturbed_data = a[:, None, None, None] * batch + std[:, None, None, None] * z      # keeps the graph to the flow
        else:
            perturbed_data = torch.empty_like(batch, dtype=torch.float32)
            L.call('indm_perturb_f32', L.ptr(batch.contiguous().float()), L.ptr(z.contiguous().float()), L.ptr(a), L.ptr(std),
                   L.ptr(perturbed_data), N, D)
        score = score_fn(perturbed_data, t)
        Zt = torch.as_tensor(Z, dtype=torch.float32, device=batch.device)
        if importance_sampling or not config.training.likelihood_weighting:
            w = Zt.expand(N) if Zt.dim() == 0 else Zt
        else:
            g2 = sde.sde(torch.zeros((N, 1, 1, 1), device=batch.device), t)[1] ** 2
            w = Zt * g2 / std ** 2
        return _DSMLoss.apply(score, z, std, w.contiguous(), (1.0 / D) if reduce_mean else 1.0)

    return loss_fn


# ------------------------------------------------------------------------------------------------ step functions
def get_step_fn(config, sde, train, optimize_fn=None, scaler=None):
    """losses.py:194-420."""
    if not config.training.continuous:
        raise NotImplementedError('INDM configs are continuous-time (training.continuous=True)')
    loss_fn = get_sde_loss_fn(config, sde, train)

    def calculate_logp(batch):
        """losses.py:219-225"""
        Ts = torch.ones(batch.shape[0], device=batch.device) * sde.T
        meanT, stdT = sde.marginal_prob(batch, Ts)
        yT = meanT + stdT[:, None, None, None] * torch.randn_like(batch)
        return sde.prior_logp(yT)

    def _attach(optimizer, st):
        if isinstance(optimizer, FusedAdamW):
            optimizer.attach_ema(st.get('ema'))

    def step_fn(state, flow_state, batch, **kw):
        """losses.py:227-256: one optimisation step of the score network (flow.model == 'identity')."""
        model, optimizer = state['model'], state['optimizer']
        optimizer.zero_grad()
        batch_size = batch.shape[0]
        nmb = config.optim.num_micro_batch
        losses_ = torch.zeros(batch_size)
        for k in range(nmb):
            sl = slice(batch_size // nmb * k, batch_size // nmb * (k + 1))
            losses = loss_fn(model, batch[sl], **kw)
            if train:
                (torch.mean(losses) / 1.0).backward()
            losses_[sl] = losses.detach().cpu()
        if train:
            _attach(optimizer, state)
            optimize_fn(optimizer, model.parameters(), step=state['step'])
            state['step'] += 1
            state['ema'].update(model.parameters())
        return losses_, None, None, None, None

    def _frozen_flow_step(state, flow_state, batch, fid_variant, **kw):
        """score-network half of flow_step_fn_nll (:258-320) / flow_step_fn_fid (:322-406) with the flow frozen (eval mode):
        latent = flow(x) without gradient, score loss + prior log-p + flow log-det / KL are all evaluated and reported, only the
        score network is updated."""
        model, flow_model, optimizer = state['model'], flow_state['model'], state['optimizer']
        batch_size = batch.shape[0]
        nmb = config.optim.num_micro_batch
        losses_, losses_score_, losses_flow_, losses_logp_ = (torch.zeros(batch_size) for _ in range(4))
        optimizer.zero_grad()
        D = float(np.prod(batch.shape[1:]))
        flow_model.eval()
        for k in range(nmb):
            sl = slice(batch_size // nmb * k, batch_size // nmb * (k + 1))
            with torch.no_grad():
                # the series the reference's training step evaluates (n + 2 terms, Neumann form, iresblock.py:114-121)
                latent, losses_flow = flow_forward(config, flow_model, batch[sl], reverse=False, estimator='train')
            if fid_variant:
                losses_score = loss_fn(model, latent, st=config.training.st, recon_loss=False, **kw)
            else:
                losses_score = loss_fn(model, latent, st=config.training.st, **kw)
            with torch.no_grad():
                losses_logp = calculate_logp(latent)
                if config.training.reduce_mean:
                    losses_flow, losses_logp = -losses_flow / D, -losses_logp / D
                else:
                    losses_flow, losses_logp = -losses_flow, -losses_logp
            if train:
                torch.mean(losses_score).backward()
            losses_[sl] = (losses_score.detach() + losses_flow + losses_logp).cpu()
            losses_score_[sl], losses_flow_[sl], losses_logp_[sl] = losses_score.detach().cpu(), losses_flow.cpu(), losses_logp.cpu()
        if train:
            optimize_fn(optimizer, model.parameters(), step=state['step'])
            state['step'] += 1
            state['ema'].update(model.parameters())
            flow_state['step'] += 1
        return losses_, losses_score_, losses_flow_, losses_logp_

    def _flow_losses(model, flow_model, mb, flow_kw, logp_noise, D, **lkw):
        """flow forward (training mode, differentiable) + score loss on the latent + prior log-p of the diffused latent"""
        latent, losses_flow = flow_forward(config, flow_model, mb, reverse=False, **flow_kw)
        losses_score = loss_fn(model, latent, **lkw)
        Ts = torch.ones(latent.shape[0], device=latent.device) * sde.T
        meanT, stdT = sde.marginal_prob(latent, Ts)
        noise = logp_noise if logp_noise is not None else torch.randn_like(latent)
        losses_logp = sde.prior_logp(meanT + stdT[:, None, None, None] * noise)          # calculate_logp, losses.py:219-225
        if config.training.reduce_mean:
            losses_flow, losses_logp = -losses_flow / D, -losses_logp / D
        else:
            losses_flow, losses_logp = -losses_flow, -losses_logp
        return latent, losses_score, losses_flow, losses_logp

    def _joint_step(state, flow_state, batch, fid_variant, **kw):
        """flow_step_fn_nll (losses.py:258-320) / flow_step_fn_fid (:322-406): joint optimisation of the flow and the score
        network.  The flow forward runs in training mode (batch-statistics BatchNorm, Neumann log-det series) and its explicit
        backward plan is reached through autograd from `torch.mean(losses).backward()`, like the score network's.
        Keyword-only test hooks: flow_kw= (draws of the flow forward), logp_noise=, draws= / draws2= (loss_fn draws of the two phases)."""
        model, flow_model = state['model'], flow_state['model']
        optimizer, flow_optimizer = state['optimizer'], flow_state['optimizer']
        batch_size = batch.shape[0]
        nmb = config.optim.num_micro_batch
        losses_, losses_score_, losses_flow_, losses_logp_ = (torch.zeros(batch_size) for _ in range(4))
        optimizer.zero_grad()
        flow_optimizer.zero_grad()
        D = float(np.prod(batch.shape[1:]))
        flow_kw = kw.pop('flow_kw', None) or {}
        logp_noise = kw.pop('logp_noise', None)
        draws2 = kw.pop('draws2', None)
        mbs = [slice(batch_size // nmb * k, batch_size // nmb * (k + 1)) for k in range(nmb)]
        if not fid_variant:
            if train:
                flow_model.train()
                dev_rows = []         # per-micro-batch loss vectors stay on the device: ONE read-back per step, after the optimiser
                core = flow_model.module if hasattr(flow_model, 'module') else flow_model
                for k_mb, sl in enumerate(mbs):   # launches are queued (the reference syncs four times per micro-batch, losses.py:306-309)
                    _, losses_score, losses_flow, losses_logp = _flow_losses(model, flow_model, batch[sl], flow_kw, logp_noise, D,
                                                                             st=config.training.st, **kw)
                    losses = losses_score + losses_flow + losses_logp
                    # last micro-batch: when autograd reaches the flow, the score network's gradients are final -> start their
                    # all-reduce on a side stream so that it overlaps the flow backward
                    last = k_mb == len(mbs) - 1 and isinstance(optimizer, FusedAdamW)
                    core._before_backward = optimizer.start_allreduce if last else None
                    torch.mean(losses).backward()
                    core._before_backward = None
                    dev_rows.append(torch.stack([losses.detach(), losses_score.detach(), losses_flow.detach(), losses_logp.detach()]))
                _attach(optimizer, state)
                _attach(flow_optimizer, flow_state)
                optimize_fn(optimizer, model.parameters(), step=state['step'])
                optimize_fn(flow_optimizer, flow_model.parameters(), step=flow_state['step'])
                host = torch.cat(dev_rows, dim=1).cpu()
                losses_, losses_score_, losses_flow_, losses_logp_ = host[0].clone(), host[1].clone(), host[2].clone(), host[3].clone()
            # update_lipschitz(flow_model) (losses.py:313) only touches spectral / induced-norm layers: a no-op for LopConv2d
            state['step'] += 1
            state['ema'].update(model.parameters())
            flow_state['step'] += 1
            flow_state['ema'].update(flow_model.parameters())
            return losses_, losses_score_, losses_flow_, losses_logp_
        if train:
            flow_model.train()
            latents = []
            # phase 1 (:354-374): the flow is trained on all three losses (score loss importance-sampled)
            for sl in mbs:
                latent, losses_score, losses_flow, losses_logp = _flow_losses(model, flow_model, batch[sl], flow_kw, logp_noise, D,
                                                                              importance_sampling=True, **kw)
                losses = losses_score + losses_flow + losses_logp
                torch.mean(losses).backward()
                latents.append(latent.detach())
                losses_[sl] = losses.detach().cpu()
                losses_flow_[sl], losses_logp_[sl] = losses_flow.detach().cpu(), losses_logp.detach().cpu()
            optimize_fn(flow_optimizer, flow_model.parameters(), step=flow_state['step'])
            flow_state['ema'].update(flow_model.parameters())
            # phase 2 (:376-400): the score network on the (re-evaluated unless st) latent
            if not config.training.st:
                optimizer.zero_grad()
            kw2 = dict(kw)
            if draws2 is not None:
                kw2['draws'] = draws2
            for k, sl in enumerate(mbs):
                if not config.training.st:
                    with torch.no_grad():
                        latent, _ = flow_forward(config, flow_model, batch[sl], log_det=None, reverse=False,
                                                 **{a: b for a, b in flow_kw.items() if a in ('eps', 'seed')})
                else:
                    latent = latents[k]
                losses_add = loss_fn(model, latent.detach(), st=config.training.st, recon_loss=False, **kw2)
                if config.training.st:
                    const_adj = (losses_add.mean() / losses_score.mean()).detach()
                    for p in model.parameters():
                        if p.grad is not None:
                            p.grad.mul_(const_adj)
                torch.mean(losses_add).backward()
                losses_score_[sl] = losses_add.detach().cpu()
            optimize_fn(optimizer, model.parameters(), step=state['step'])
            state['ema'].update(model.parameters())
            state['step'] += 1
            flow_state['step'] += 1
        return losses_, losses_score_, losses_flow_, losses_logp_

    def _frozen(config):
        return bool(getattr(config.training, 'freeze_flow', False))

    def flow_step_fn_nll(state, flow_state, batch, **kw):
        if _frozen(config):
            return _frozen_flow_step(state, flow_state, batch, False, **kw)
        return _joint_step(state, flow_state, batch, False, **kw)

    def flow_step_fn_fid(state, flow_state, batch, **kw):
        if _frozen(config):
            return _frozen_flow_step(state, flow_state, batch, True, **kw)
        return _joint_step(state, flow_state, batch, True, **kw)

    if config.flow.model == 'identity':
        logging.info('Train only the score network.')
        return step_fn
    if not config.training.likelihood_weighting:
        logging.info('Train score network with FID-favorable setting (weighting function = variance weighting).')
        return flow_step_fn_fid
    logging.info('Train score network with NLL-favorable setting (weighting function = likelihood weighting).')
    return flow_step_fn_nll
