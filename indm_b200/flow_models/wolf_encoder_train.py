"""Posterior encoder of the wolf flow in TRAINING mode (modules/encoders/global_encoder.py:12-44, nnet/resnets/
resnet_batchnorm.py:18-76): nn.BatchNorm2d with batch statistics, forward and explicit backward on the C-ABI kernels.

Forward per ResNet block:  y1 = conv1(x) -> t1 = ELU(BN1(y1)) -> y2 = conv2(t1);  r = BNd(convd(x)) or x;  out = ELU(BN2(y2) + r).
Convolutions are indm_igemm launches (TMA-strided for the stride-2 ones); BatchNorm statistics / apply / backward are the
indm_bn_* kernels.  The backward of a stride-2 convolution is expressed on the stride-1 tensor-core kernels by inserting zeros
between the output-gradient pixels: with dz[2o] = dy[o], dgrad = conv3x3(dz, flipped W) and wgrad = sum_p dz[p] (x) x[p + tap],
exactly the stride-1 forms (4x the work on layers that carry < 1 % of the step).
"""
import ctypes

import torch

from .. import _lib as L


def _i64(v):
    return ctypes.c_int64(int(v))


def _f(v):
    return ctypes.c_float(float(v))


class EncoderTrain:
    def __init__(self, eng):
        self.eng = eng
        e = eng
        core, N, dev = e.core, e.N, e.dev
        p = core.config.flow.wolf_params['discriminator']['encoder']
        kc = e.kchunk
        self.cp = cp = lambda c: ((c + kc - 1) // kc) * kc
        c0, S, _ = core.input_shape
        self.c0, self.S = c0, S
        self.op_dt = L.DTYPE_BF16 if e.mode == 'bf16' else L.DTYPE_F32
        self.x_in = torch.zeros((N, S, S, cp(c0)), device=dev, dtype=e.tdtype)
        net = core.discriminator.encoder.net
        self.blocks = []
        self.pack_jobs = []
        nbn = 0
        H, inp = S, c0
        first = True
        for lv, hid in enumerate(p['hidden_planes']):
            res = getattr(net, f'resnet{lv}')
            for m, stride in enumerate((1, 2)):
                blk = res.main[m]
                ci = inp if m == 0 else hid
                Ho = H // stride
                B = dict(blk=blk, ci=ci, co=hid, stride=stride, H=H, Ho=Ho, first=first, ds=hasattr(blk, 'downsample'))
                B['c1'] = self._conv(blk.conv1, ci, hid, 3, need_dgrad=not first)
                B['c2'] = self._conv(blk.conv2, hid, hid, 3, need_dgrad=True)
                if B['ds']:
                    B['cd'] = self._conv(blk.downsample[0], ci, hid, 1, need_dgrad=not first)
                f32 = lambda h, c: torch.zeros((N, h, h, cp(c)), device=dev)
                opt = lambda h, c: torch.zeros((N, h, h, cp(c)), device=dev, dtype=e.tdtype)
                B.update(y1=f32(Ho, hid), t1=opt(Ho, hid), y2=f32(Ho, hid), out=opt(Ho, hid), out_f=f32(Ho, hid),
                         dy1=opt(Ho, hid), dy2=opt(Ho, hid), g_t1=f32(Ho, hid), gs=f32(Ho, hid), gx=None if first else f32(H, ci))
                if B['ds']:
                    B.update(yd=f32(Ho, hid), r=f32(Ho, hid), dyd=opt(Ho, hid))
                if stride == 2:
                    B.update(dz1=opt(H, hid), dzd=opt(H, hid))
                B['bn_ids'] = (nbn, nbn + 1, nbn + 2 if B['ds'] else None)
                nbn += 3 if B['ds'] else 2
                self.blocks.append(B)
                H, first = Ho, False
            inp = hid
        self.nbn = nbn
        self.cmax = max(p['hidden_planes'])
        self.sums = torch.zeros((2, nbn, 2 * self.cmax), device=dev)         # [0]: forward (sum, sumsq); [1]: backward (sum gs, sum gs xhat)
        # top 1x1 conv + ELU -> NCHW fp32 [N, out_planes, h, w] == the flattened encoder output
        self.out_planes = p['out_planes']
        self.Hl, self.Cl = H, inp
        top = net.top
        self.top_w = torch.zeros((1, self.out_planes, cp(inp)), device=dev, dtype=e.tdtype)
        self.top_wd = torch.zeros((1, inp, cp(self.out_planes)), device=dev, dtype=e.tdtype)
        self.top_b = torch.zeros((self.out_planes,), device=dev)

        def job_top():
            W = top.weight.detach().to(dev, torch.float32)[:, :, 0, 0]
            self.top_w.zero_(); self.top_wd.zero_()
            self.top_w[0, :, :inp].copy_(e._round(W))
            self.top_wd[0, :, :self.out_planes].copy_(e._round(W.t()))
            self.top_b.copy_(top.bias.detach())
        self.pack_jobs.append(job_top)
        self.top = torch.zeros((N, self.out_planes, H, H), device=dev)
        self.g_top = torch.zeros((N, H, H, cp(self.out_planes)), device=dev, dtype=e.tdtype)
        self.g_last = torch.zeros((N, H, H, cp(inp)), device=dev)
        self._version = None

    # ---- weights
    def _conv(self, conv, ci, co, k, need_dgrad):
        e, dev, cp = self.eng, self.eng.dev, self.cp
        w = torch.zeros((k * k, co, cp(ci)), device=dev, dtype=e.tdtype)
        wd = torch.zeros((k * k, ci, cp(co)), device=dev, dtype=e.tdtype) if need_dgrad else None

        def job():
            W = conv.weight.detach().to(dev, torch.float32)
            w.zero_()
            w[:, :, :ci].copy_(e._round(W.permute(2, 3, 0, 1).reshape(k * k, co, ci)))
            if wd is not None:
                # transposed convolution: taps flipped, channels swapped
                wd.zero_()
                wd[:, :, :co].copy_(e._round(W.flip(2, 3).permute(2, 3, 1, 0).reshape(k * k, ci, co)))
        self.pack_jobs.append(job)
        return dict(conv=conv, w=w, wd=wd, k=k, ci=ci, co=co)

    def _ensure(self):
        v = self.eng.version()
        if self._version != v:
            def body():
                for job in self.pack_jobs:
                    job()
            with torch.no_grad():
                self.eng._graphed(('enc_train_pack',), body)
            self._version = v

    # ---- launches
    def _fwd_conv(self, a, Hin, c, stride, out_f32):
        e, cp = self.eng, self.cp
        Ho = Hin // stride
        kw = dict(dtype=e.dt, a=a, N=e.N, H=Ho, W=Ho, Cin=cp(c['ci']), b=c['w'], Cout=c['co'], taps=c['k'] * c['k'], out_f32=out_f32,
                  out_ld=cp(c['co']))
        if stride == 2:
            kw.update(stride=2, a_H=Hin, a_W=Hin, pad=1 if c['k'] == 3 else 0)
        L.igemm(**kw)

    def _dgrad(self, dy, H, c, out_f32, residual=None):
        """out [N,H,H,cp(ci)] = conv_transpose(dy) (+ residual); dy [N,H,H,cp(co)] (already zero-inserted for stride 2)"""
        e, cp = self.eng, self.cp
        kw = dict(dtype=e.dt, a=dy, N=e.N, H=H, W=H, Cin=cp(c['co']), b=c['wd'], Cout=c['ci'], taps=c['k'] * c['k'], out_f32=out_f32,
                  out_ld=cp(c['ci']))
        if residual is not None:
            kw.update(residual=residual, res_ld=cp(c['ci']), res_scale=1.0)
        L.igemm(**kw)

    def _wgrad(self, dy, x, H, c):
        conv = c['conv']
        if not conv.weight.requires_grad:
            return
        if conv.weight.grad is None:
            conv.weight.grad = torch.zeros_like(conv.weight)
        cp, taps = self.cp, c['k'] * c['k']
        L.call('indm_conv_wgrad', L.ptr(dy), _i64(cp(c['co'])), L.ptr(x), _i64(cp(c['ci'])), self.op_dt, self.eng.N, H, H, c['co'], c['ci'], taps,
               L.ptr(conv.weight.grad), _i64(c['ci'] * taps), _i64(taps), _i64(1 if taps == 9 else 0), _f(1.0))

    def _bn_fwd(self, idx, bn, y, P, C, residual, act, out_op, out_f32, update_running):
        s = self.sums[0, idx]
        L.call('indm_bn_stats', L.ptr(y), _i64(P), C, self.cp(C), L.ptr(s))
        L.call('indm_bn_apply', L.ptr(y), L.ptr(s), L.ptr(bn.weight.detach()), L.ptr(bn.bias.detach()), _i64(P), C, self.cp(C), L.ptr(residual), act,
               L.ptr(out_op), L.ptr(out_f32), self.op_dt)
        if update_running:
            # nn.BatchNorm2d buffers, momentum 0.1, unbiased variance (torch/nn/modules/batchnorm.py)
            with torch.no_grad():
                mean = s[:C] / P
                var = (s[C:2 * C] / P - mean * mean).clamp_(min=0) * (P / max(P - 1, 1))
                bn.running_mean.mul_(0.9).add_(mean, alpha=0.1)
                bn.running_var.mul_(0.9).add_(var, alpha=0.1)
                bn.num_batches_tracked.add_(1)

    def _bn_bwd(self, idx, bn, g, o, y, P, C, dy, gs_out):
        s, b = self.sums[0, idx], self.sums[1, idx]
        L.call('indm_bn_bwd_stats', L.ptr(g), L.ptr(o), L.ptr(y), L.ptr(s), _i64(P), C, self.cp(C), L.ptr(b), self.op_dt)
        L.call('indm_bn_bwd_apply', L.ptr(g), L.ptr(o), L.ptr(y), L.ptr(s), L.ptr(b), L.ptr(bn.weight.detach()), _i64(P), C, self.cp(C), L.ptr(dy),
               L.ptr(gs_out), self.op_dt)
        for prm, val in ((bn.weight, b[C:2 * C]), (bn.bias, b[:C])):
            if prm.requires_grad:
                if prm.grad is None:
                    prm.grad = torch.zeros_like(prm)
                prm.grad.add_(val)

    # ---- passes
    def forward(self, x, update_running=True):
        """x [N, c0, S, S] fp32 NCHW -> encoder output [N, out_planes * h * w] (a view of an engine-owned buffer)"""
        e, N = self.eng, self.eng.N
        self._ensure()
        self.sums.zero_()
        x = x.float().contiguous()
        L.call('indm_prep_input', L.ptr(x), L.ptr(self.x_in), N, self.c0, self.S, self.S, self.x_in.shape[-1], _f(1.0), _f(0.0), 0, e.dt)
        a, af = self.x_in, None
        for B in self.blocks:
            blk, hid, Ho = B['blk'], B['co'], B['Ho']
            P = N * Ho * Ho
            i1, i2, idd = B['bn_ids']
            B['x'] = a
            self._fwd_conv(a, B['H'], B['c1'], B['stride'], B['y1'])
            self._bn_fwd(i1, blk.bn1, B['y1'], P, hid, None, 2, B['t1'], None, update_running)
            self._fwd_conv(B['t1'], Ho, B['c2'], 1, B['y2'])
            if B['ds']:
                self._fwd_conv(a, B['H'], B['cd'], B['stride'], B['yd'])
                self._bn_fwd(idd, blk.downsample[1], B['yd'], P, hid, None, 0, None, B['r'], update_running)
                r = B['r']
            else:
                r = af
            self._bn_fwd(i2, blk.bn2, B['y2'], P, hid, r, 2, B['out'], B['out_f'], update_running)
            a, af = B['out'], B['out_f']
        self.x_last = a
        H = self.Hl
        L.igemm(dtype=e.dt, a=a, N=N, H=H, W=H, Cin=self.cp(self.Cl), b=self.top_w, Cout=self.out_planes, taps=1, bias=self.top_b, act=2,
                out_mode=1, out_f32=self.top)
        return self.top.view(N, -1)

    def backward(self, g_out):
        """g_out [N, out_planes * h * w]: gradient w.r.t. forward()'s result; accumulates every encoder parameter gradient"""
        e, N, cp = self.eng, self.eng.N, self.cp
        top = e.core.discriminator.encoder.net.top
        H = self.Hl
        # top: ELU' on the tiny [N, 8, 4, 4] head, then 1x1 dgrad / wgrad
        gp = (g_out.view_as(self.top) * torch.where(self.top > 0, torch.ones_like(self.top), self.top + 1.0)).contiguous()
        L.call('indm_nchw_to_nhwc', L.ptr(gp), None, L.ptr(self.g_top), _i64(N), self.out_planes, H, H, self.g_top.shape[-1], _f(1.0), self.op_dt)
        ctop = dict(conv=top, wd=self.top_wd, k=1, ci=self.Cl, co=self.out_planes)
        self._dgrad(self.g_top, H, ctop, self.g_last)
        self._wgrad(self.g_top, self.x_last, H, ctop)
        if top.bias.requires_grad:
            if top.bias.grad is None:
                top.bias.grad = torch.zeros_like(top.bias)
            top.bias.grad.add_(gp.sum(dim=(0, 2, 3)))
        g = self.g_last
        for B in reversed(self.blocks):
            blk, hid, Ho, Hin, s = B['blk'], B['co'], B['Ho'], B['H'], B['stride']
            P = N * Ho * Ho
            i1, i2, idd = B['bn_ids']
            self._bn_bwd(i2, blk.bn2, g, B['out'], B['y2'], P, hid, B['dy2'], B['gs'])
            self._wgrad(B['dy2'], B['t1'], Ho, B['c2'])
            self._dgrad(B['dy2'], Ho, B['c2'], B['g_t1'])
            self._bn_bwd(i1, blk.bn1, B['g_t1'], B['t1'], B['y1'], P, hid, B['dy1'], None)
            d1 = B['dy1']
            if s == 2:
                B['dz1'][:, ::2, ::2].copy_(B['dy1'])
                d1 = B['dz1']
            self._wgrad(d1, B['x'], Hin, B['c1'])
            if B['ds']:
                self._bn_bwd(idd, blk.downsample[1], B['gs'], None, B['yd'], P, hid, B['dyd'], None)
                dd = B['dyd']
                if s == 2:
                    B['dzd'][:, ::2, ::2].copy_(B['dyd'])
                    dd = B['dzd']
                self._wgrad(dd, B['x'], Hin, B['cd'])
            if B['first']:
                break                                         # the data batch needs no gradient
            if B['ds']:
                self._dgrad(dd, Hin, B['cd'], B['gx'])
                self._dgrad(d1, Hin, B['c1'], B['gx'], residual=B['gx'])
            else:
                self._dgrad(d1, Hin, B['c1'], B['gx'], residual=B['gs'])      # identity shortcut
            g = B['gx']
