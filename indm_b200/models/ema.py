"""`ExponentialMovingAverage` with the reference's interface (models/ema.py:10-97).  The update runs as one fused kernel over
flat storage when the parameters tile one contiguous buffer (they do once `losses.get_optimizer` has re-homed them), else one
launch per tensor."""
import torch

from .. import _lib as L


def flat_view(tensors):
    """the 1-D tensor that `tensors` tile back to back inside one storage, or None"""
    tensors = list(tensors)
    if not tensors:
        return None
    t0 = tensors[0]
    ptr, total = t0.data_ptr(), 0
    for t in tensors:
        if not t.is_contiguous() or t.dtype != torch.float32 or t.data_ptr() != ptr + 4 * total or t.device != t0.device:
            return None
        if t.untyped_storage().data_ptr() != t0.untyped_storage().data_ptr():
            return None
        total += t.numel()
    off = (ptr - t0.untyped_storage().data_ptr()) // 4
    return torch.empty(0, dtype=torch.float32, device=t0.device).set_(t0.untyped_storage(), off, (total,), (1,))


class ExponentialMovingAverage:
    """Maintains (exponential) moving average of a set of parameters (models/ema.py:10)."""

    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError('Decay must be between 0 and 1')
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        params = [p for p in parameters if p.requires_grad]
        total = sum(p.numel() for p in params)
        dev = params[0].device if params else torch.device('cpu')
        self._flat = torch.empty((total,), dtype=torch.float32, device=dev)
        self.shadow_params, off = [], 0
        for p in params:
            s = self._flat[off:off + p.numel()].view_as(p)
            s.copy_(p.detach())
            self.shadow_params.append(s)
            off += p.numel()
        self.collected_params = []
        self._fused_done = False

    def next_decay(self):
        """the decay the next `update` call will use (models/ema.py:38-41)"""
        if self.num_updates is None:
            return self.decay
        n = self.num_updates + 1
        return min(self.decay, (1 + n) / (10 + n))

    def fused_update_done(self):
        """the optimiser kernel already applied the next update (losses.FusedAdamW.attach_ema): the next `update` only counts it"""
        self._fused_done = True

    def update(self, parameters):
        """models/ema.py:32-51: shadow -= (1 - decay) * (shadow - param), decay = min(decay, (1 + n) / (10 + n))"""
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        if self._fused_done:
            self._fused_done = False
            return
        params = [p for p in parameters if p.requires_grad]
        if not params:
            return
        if not params[0].is_cuda:
            raise RuntimeError('indm_b200 ExponentialMovingAverage needs CUDA parameters: there is no CPU path')
        flat = flat_view([p.data for p in params])
        if flat is not None and flat.numel() == self._flat.numel():
            L.call('indm_ema_f32', L.ptr(self._flat), L.ptr(flat), flat.numel(), float(decay))
        else:
            for s, p in zip(self.shadow_params, params):
                L.call('indm_ema_f32', L.ptr(s), L.ptr(p.data.contiguous()), p.numel(), float(decay))

    def copy_to(self, parameters):
        parameters = [p for p in parameters if p.requires_grad]
        for s_param, param in zip(self.shadow_params, parameters):
            param.data.copy_(s_param.data)
        L.param_epoch += 1

    def store(self, parameters):
        self.collected_params = [param.clone() for param in parameters]

    def restore(self, parameters):
        for c_param, param in zip(self.collected_params, parameters):
            param.data.copy_(c_param.data)
        L.param_epoch += 1

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.num_updates = state_dict['num_updates']
        for s, src in zip(self.shadow_params, state_dict['shadow_params']):
            s.copy_(src)
