"""Explicit backward pass of the wolf flow's training forward (what `torch.mean(losses).backward()` does to the flow in
losses.py:300-304 of the reference), on the C-ABI kernels.

Per iResBlock y = x + g(x; h), g = W3 phi(W2 (phi(W1 psi(x) + b1) + A h + a) + b2) + b3 with phi = Sin, psi = Sin (identity in the
first block of the flow), the training forward evaluates S = w^T J_g eps with w the (constant) Neumann vector
(iresblock.py:264-273).  With the per-sample loss weight folded into w, D_i = diag(phi'(a_i)), and phi'' = -4 pi^2 phi:

    reverse chain of w     r3 = W3^T w,  q2 = D2 r3,  r2 = W2^T q2,  q1 = D1 r2,  r1 = W1^T q1
    forward mode of eps    e0 = D0 eps,  u1 = W1 e0,  t1 = D1 u1,    u2 = W2 t1,  t2 = D2 u2
    pre-activation grads   G2 = D2 W3^T gy - 4 pi^2 r3 u2 s2,   G1 = D1 W2^T G2 - 4 pi^2 r2 u1 s1
    input grad             gx = gy + D0 W1^T G1 - 4 pi^2 r1 eps psi(x)
    weight grads           dW3 = gy (x) s2 + w (x) t2,  dW2 = G2 (x) (s1 + c) + q2 (x) t1,  dW1 = G1 (x) psi(x) + q1 (x) e0

(checked against autograd to 1e-16 in fp64).  Every contraction is an indm_igemm / indm_conv_wgrad launch over the same
tap-packed operands the forward uses; the normalised-weight gradients then go through the Lipschitz normalisation's own
backward (indm_lop_bwd_f32) into `param.grad`.
"""
import ctypes
import math

import torch

from .. import _lib as L

K2 = -4.0 * math.pi ** 2
COEFF = 0.98


def _i64(v):
    return ctypes.c_int64(int(v))


def _f(v):
    return ctypes.c_float(float(v))


class FlowBackward:
    """Work buffers + launch lists of the backward of `FlowEngine.forward_logdet(training=True, save=True)`."""

    def __init__(self, eng):
        self.eng = eng
        e = eng
        N, idim, dev = e.N, e.idim, e.dev
        c0, h0, w0 = e.core.input_shape
        big = lambda: torch.empty((N, h0, w0, idim), device=dev, dtype=e.tdtype)
        names = ('S1', 'S2', 'D1', 'D2', 'R3', 'Q2', 'R2', 'Q1', 'U1F', 'T1F', 'U2F', 'T2F', 'GA', 'G2', 'G1')
        self.big = {n: big() for n in names}
        # tap-packed operands (im2col of x, w, e0, gy) per scale
        self.A = [{n: torch.zeros_like(e.a0[s]) for n in ('x', 'w', 'e', 'g')} for s in range(len(e.nb))]
        self.o9 = [torch.empty_like(e.o9[s]) for s in range(len(e.nb))]
        self.grads = {}
        hd = e.core.latent_dim
        nblk = len(e.blocks)
        z = lambda *shape: torch.zeros(shape, device=dev)
        # conditioning-path gradients of all blocks stacked, so their small products run as ONE batched launch each
        self.gw2_all, self.gA_all, self.ga_all, self.gb2_all = z(nblk, idim, idim), z(nblk, idim, hd), z(nblk, idim), z(nblk, idim)
        self.gb2_img_all = z(nblk, N, idim)
        self.gc_all = torch.empty((nblk, N, idim), device=dev)
        self.cvec_all = torch.empty((nblk, N, idim), device=dev)
        self.ghb_all = torch.empty((nblk, N, hd), device=dev)
        for i, (s, b, m) in enumerate(e.blocks):
            c, kp = m.channels, e.kp[s]
            # packed like the forward operands: w1 [idim][kp], w3 [kp][idim] (columns / rows >= 9c stay zero)
            self.grads[i] = dict(w1=z(idim, kp), w2=self.gw2_all[i], w3=z(kp, idim), b1=z(idim), b2=self.gb2_all[i], b3=z(c),
                                 A=self.gA_all[i], a=self.ga_all[i])
        self.gh = torch.zeros((N, hd), device=dev)
        self.ones = {}
        self._ops = {}

    # ---- launch-list helpers (same memoisation as FlowEngine._replay)
    def _view(self, name, s):
        e = self.eng
        _, h0, w0 = e.core.input_shape
        H, Wd = h0 >> s, w0 >> s
        n_el = e.N * H * Wd * e.idim
        return self.big[name].view(-1)[:n_el].view(e.N, H, Wd, e.idim)

    def _block_ops(self, i, s, m, xin, ve, wS, gy, gx, tmp):
        e = self.eng
        N, idim, dt = e.N, e.idim, e.dt
        c = m.channels
        _, h0, w0 = e.core.input_shape
        H, Wd = h0 >> s, w0 >> s
        kp, ld9 = e.kp[s], e.ld9[s]
        Wt = e.w[(e.blocks[i][0], e.blocks[i][1])]
        G = self.grads[i]
        V = lambda n: self._view(n, s)
        S1, S2, D1, D2, R3, Q2, R2, Q1 = (V(n) for n in ('S1', 'S2', 'D1', 'D2', 'R3', 'Q2', 'R2', 'Q1'))
        U1F, T1F, U2F, T2F, GA, G2, G1 = (V(n) for n in ('U1F', 'T1F', 'U2F', 'T2F', 'GA', 'G2', 'G1'))
        Ax, Aw, Ae, Ag = (self.A[s][n] for n in ('x', 'w', 'e', 'g'))
        o9 = self.o9[s]
        ob = (lambda t: dict(out_bf16=t)) if e.mode == 'bf16' else (lambda t: dict(out_f32=t))
        n_big = S1.numel()
        n_x = xin.numel()
        wdt = L.DTYPE_BF16 if e.mode == 'bf16' else L.DTYPE_F32
        call, ig = e._mk_call, e._mk_igemm
        geo = dict(dtype=dt, N=N, H=H, W=Wd, taps=1)
        first = m.first
        d0 = e._static(f'bw_d0_{s}', xin) if not first else None
        x0 = e._static(f'bw_x0_{s}', xin) if not first else None
        e0 = e._static(f'bw_e0_{s}', xin) if not first else ve
        r1 = e._static(f'bw_r1_{s}', xin)
        ops = []
        if not first:
            ops += [call('indm_cos2pi_f32', xin, d0, _i64(n_x)), call('indm_sin2pi_f32', xin, x0, _i64(n_x)),
                    call('indm_mul_op', d0, ve, e0, _i64(n_x), L.DTYPE_F32)]
        # 1. recompute the branch, keeping post-activations (S) and Sin' factors (D)
        ops += [
            call('indm_im2col3x3_nchw', xin, Ax, _i64(N), c, H, Wd, kp, 0, 0 if first else 1, dt),
            ig(a=Ax, Cin=kp, b=Wt['w1'], Cout=idim, bias=Wt['b1'], act=1, out_ld=idim, aux_cos=D1, **geo, **ob(S1)),
            ig(a=S1, Cin=idim, b=Wt['w2'], Cout=idim, rowbias=e.cond[:, i * idim:], rowbias_ld=e.cond.shape[1], act=1, out_ld=idim,
               aux_cos=D2, **geo, **ob(S2)),
        ]
        # 2. reverse chain of the (loss-weighted) Neumann vector
        ops += [
            call('indm_im2col3x3_nchw', wS, Aw, _i64(N), c, H, Wd, kp, 1, 0, dt),
            ig(a=Aw, Cin=kp, b=Wt['w3v'], Cout=idim, out_ld=idim, **geo, **ob(R3)),
            call('indm_mul_op', R3, D2, Q2, _i64(n_big), dt),
            ig(a=Q2, Cin=idim, b=Wt['w2d'], Cout=idim, out_ld=idim, **geo, **ob(R2)),
            call('indm_mul_op', R2, D1, Q1, _i64(n_big), dt),
            ig(a=Q1, Cin=idim, b=Wt['w1v'], Cout=ld9, out_f32=o9, out_ld=ld9, **geo),
            call('indm_col2im3x3_nchw', o9, _i64(ld9), None, None, None, _f(1.0), r1, _i64(N), c, H, Wd, 1),
        ]
        # 3. forward mode of the probe
        ops += [
            call('indm_im2col3x3_nchw', e0, Ae, _i64(N), c, H, Wd, kp, 0, 0, dt),
            ig(a=Ae, Cin=kp, b=Wt['w1'], Cout=idim, out_ld=idim, **geo, **ob(U1F)),
            call('indm_mul_op', U1F, D1, T1F, _i64(n_big), dt),
            ig(a=T1F, Cin=idim, b=Wt['w2'], Cout=idim, out_ld=idim, **geo, **ob(U2F)),
            call('indm_mul_op', U2F, D2, T2F, _i64(n_big), dt),
        ]
        # 4. reverse chain of the incoming gradient with the second-order terms
        ops += [
            call('indm_im2col3x3_nchw', gy, Ag, _i64(N), c, H, Wd, kp, 1, 0, dt),
            ig(a=Ag, Cin=kp, b=Wt['w3v'], Cout=idim, out_ld=idim, mul=D2, mul_ld=idim, **geo, **ob(GA)),
            call('indm_fma3_op', GA, None, R3, U2F, S2, G2, _i64(n_big), _f(K2), dt),
            ig(a=G2, Cin=idim, b=Wt['w2d'], Cout=idim, out_ld=idim, mul=D1, mul_ld=idim, **geo, **ob(GA)),
            call('indm_fma3_op', GA, None, R2, U1F, S1, G1, _i64(n_big), _f(K2), dt),
            ig(a=G1, Cin=idim, b=Wt['w1v'], Cout=ld9, out_f32=o9, out_ld=ld9, **geo),
            call('indm_col2im3x3_nchw', o9, _i64(ld9), None, None, d0, _f(1.0), tmp, _i64(N), c, H, Wd, 1),
        ]
        if first:
            ops += [call('indm_axpy_f32', tmp, gy, _f(1.0), _i64(n_x))]          # gx = gy + W1^T G1
            gx_src = tmp
        else:
            ops += [call('indm_fma3_op', gy, tmp, r1, ve, x0, gx, _i64(n_x), _f(K2), L.DTYPE_F32)]
            gx_src = gx
        # 5. weight gradients (fp32, accumulated by the wgrad kernel's split-K atomics)
        wg = lambda dy, dy_ld, Co, x, x_ld, Ci, dw, so: call('indm_conv_wgrad', dy, _i64(dy_ld), x, _i64(x_ld), wdt, N, H, Wd, Co, Ci, 1, dw,
                                                            _i64(so), _i64(1), _i64(0), _f(1.0))
        ops += [
            wg(Ag, kp, kp, S2, idim, idim, G['w3'], idim), wg(Aw, kp, kp, T2F, idim, idim, G['w3'], idim),
            wg(G2, idim, idim, S1, idim, idim, G['w2'], idim), wg(Q2, idim, idim, T1F, idim, idim, G['w2'], idim),
            wg(G1, idim, idim, Ax, kp, kp, G['w1'], kp), wg(Q1, idim, idim, Ae, kp, kp, G['w1'], kp),
            call('indm_colsum', G1, wdt, _i64(N), _i64(H * Wd), idim, _i64(idim), None, _i64(0), G['b1'], _f(1.0)),
            call('indm_colsum', G2, wdt, _i64(N), _i64(H * Wd), idim, _i64(idim), self.gb2_img_all[i], _i64(idim), None, _f(1.0)),
        ]
        return ops, gx_src

    def block(self, i, s, m, xin, ve, wS, gy, gx, tmp):
        """gx <- gradient w.r.t. the block input; accumulates the block's normalised-weight / bias / conditioning gradients"""
        e = self.eng
        key = ('bwd', i, xin.data_ptr(), ve.data_ptr(), wS.data_ptr(), gy.data_ptr(), gx.data_ptr(), tmp.data_ptr())
        ent = self._ops.get(key)
        if ent is None:
            ent = self._block_ops(i, s, m, xin, ve, wS, gy, gx, tmp)
            self._ops[key] = ent
        ops, gx_src = ent
        for op in ops:
            op()
        # db3 = sum over images and pixels of gy (NCHW): row sums against ones, then over images
        c = m.channels
        HW = xin.shape[2] * xin.shape[3]
        rows = torch.empty((e.N * c,), device=e.dev)
        ones = self.ones.get(s)
        if ones is None:
            ones = self.ones[s] = torch.ones_like(gy)
        L.call('indm_rowdot_f32', L.ptr(gy), L.ptr(ones), L.ptr(rows), e.N * c, HW, _f(1.0), 0)
        self.grads[i]['b3'].add_(rows.view(e.N, c).sum(0))
        return gx_src

    def _conditioning_path(self):
        """a2 = W2 (s1 + A h + a) + b2 (lipschitz.py:431-435) for all blocks at once: with gb2 = per-image pixel sums of G2,
        gc = gb2 W2;  dA = gc^T h;  da = sum_n gc;  dh += gc A;  dW2 += gb2^T (A h + a);  db2 = sum_n gb2.
        Per block these are [N x 512] x [512 x 512]-sized products on a handful of CTAs (5 launches x 32 blocks took 10 ms);
        batched over the blocks they are 4 launches."""
        e = self.eng
        idim, hd, N, nblk = e.idim, e.core.latent_dim, e.N, len(e.blocks)
        A_all = torch.stack([m.convs()[1].h_net.net.weight.detach() for (_, _, m) in e.blocks]).contiguous()      # [nblk, idim, hd]
        a_all = torch.stack([m.convs()[1].h_net.net.bias.detach() for (_, _, m) in e.blocks])                      # [nblk, idim]

        def bg(ta, tb, M, Nn, K, A, lda, sA, B, ldb, sB, beta, C, ldc, sC):
            L.call('indm_sgemm_batched_f32', ta, tb, M, Nn, K, _f(1.0), L.ptr(A), _i64(lda), _i64(sA), L.ptr(B), _i64(ldb), _i64(sB), _f(beta),
                   L.ptr(C), _i64(ldc), _i64(sC), nblk)
        gb = self.gb2_img_all
        bg(0, 0, N, idim, idim, gb, idim, N * idim, e.w2f_all, idim, idim * idim, 0.0, self.gc_all, idim, N * idim)        # gc = gb2 W2
        bg(1, 0, idim, hd, N, self.gc_all, idim, N * idim, e.h, hd, 0, 0.0, self.gA_all, hd, idim * hd)                    # dA = gc^T h
        self.ga_all.copy_(self.gc_all.sum(dim=1))
        bg(0, 0, N, hd, idim, self.gc_all, idim, N * idim, A_all, hd, idim * hd, 0.0, self.ghb_all, hd, N * hd)           # gc A per block
        self.gh.add_(self.ghb_all.sum(dim=0))
        bg(0, 1, N, idim, hd, e.h, hd, 0, A_all, hd, idim * hd, 0.0, self.cvec_all, idim, N * idim)                        # h A^T
        self.cvec_all.add_(a_all[:, None, :])
        bg(1, 0, idim, idim, N, gb, idim, N * idim, self.cvec_all, idim, N * idim, 1.0, self.gw2_all, idim, idim * idim)  # dW2 += gb2^T c
        self.gb2_all.copy_(gb.sum(dim=1))

    def run(self, gz, cS):
        """gz: gradient w.r.t. the flow output (the flow's own input layout [N, c0, H, W]); cS [N]: d loss / d (sum of block
        log-dets) per sample.  Returns (gx, gh): gradients w.r.t. the flow input and the conditioning latent h; the parameter
        gradients are accumulated into `param.grad` of the residual-flow convolutions and conditioning layers."""
        e = self.eng
        if not e._saved:
            raise RuntimeError('flow backward without a saved training forward (forward_logdet(training=True, save=True))')
        from .wolf import _squeeze2
        N = e.N
        for G in self.grads.values():
            for t in G.values():
                t.zero_()
        self.gh.zero_()
        self.gb2_img_all.zero_()
        cS = cS.float().contiguous()
        g = gz.float().contiguous()
        nscales = len(e.nb)
        if nscales > 1:
            for _ in range(nscales - 1):
                g = _squeeze2(g).contiguous()
        cur_scale = nscales - 1
        for (i, s, m, sx, sv, sw) in reversed(e._saved):
            while s < cur_scale:                     # leave a squeezed scale: inverse of SqueezeLayer(2) (squeeze.py:19-30)
                n_, c_, h_, w_ = g.shape
                g = g.view(n_, c_ // 4, 2, 2, h_, w_).permute(0, 1, 4, 2, 5, 3).reshape(n_, c_ // 4, 2 * h_, 2 * w_).contiguous()
                cur_scale -= 1
            gy = e._static(f'bw_gy_{s}', sx)
            gy.copy_(g)
            wS = e._static(f'bw_ws_{s}', sx)
            L.call('indm_rowscale_f32', L.ptr(sw), L.ptr(cS), L.ptr(wS), N, sw[0].numel())
            gx = e._static(f'bw_gx_{s}', sx)
            tmp = e._static(f'bw_tmp_{s}', sx)
            g = self.block(i, s, m, sx, sv, wS, gy, gx, tmp)
        while cur_scale > 0:
            n_, c_, h_, w_ = g.shape
            g = g.view(n_, c_ // 4, 2, 2, h_, w_).permute(0, 1, 4, 2, 5, 3).reshape(n_, c_ // 4, 2 * h_, 2 * w_).contiguous()
            cur_scale -= 1
        self._conditioning_path()
        self._to_param_grads()
        return g.clone(), self.gh.clone()

    def _to_param_grads(self):
        """normalised-weight gradients -> parameter gradients through LopConv2d.compute_weight's backward; accumulates"""
        e = self.eng
        idim = e.idim

        def acc(p, g):
            if not p.requires_grad:
                return
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            p.grad.add_(g.view_as(p))

        def lop(p, gn):
            if not p.requires_grad:
                return
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            rows = p.shape[0]
            raw = p.detach().contiguous()
            gn = gn.contiguous()
            L.call('indm_lop_bwd_f32', L.ptr(raw), L.ptr(gn), L.ptr(p.grad), rows, raw.numel() // rows, _f(COEFF), 1)

        for i, (s, b, m) in enumerate(e.blocks):
            cv1, cv2, cv3 = m.convs()
            G = self.grads[i]
            c = m.channels
            lop(cv1.weight, G['w1'][:, :9 * c].reshape(idim, 9, c).permute(0, 2, 1).reshape(idim, c, 3, 3))
            lop(cv2.weight, G['w2'].view(idim, idim, 1, 1))
            lop(cv3.weight, G['w3'][:9 * c].reshape(9, c, idim).permute(1, 2, 0).reshape(c, idim, 3, 3))
            acc(cv1.bias, G['b1']); acc(cv2.bias, G['b2']); acc(cv3.bias, G['b3'])
            acc(cv2.h_net.net.weight, G['A']); acc(cv2.h_net.net.bias, G['a'])


def _acc(p, g):
    if not p.requires_grad:
        return
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    p.grad.add_(g.view_as(p))


def _wn_backward(lin, dW):
    """legacy weight_norm, dim 0 (nnet/weight_norm.py:8-40): w = g v / |v| per output row -> (d g, d v) from d w.  Parameter
    re-parameterisation on [<=128, 256] matrices: plain elementwise torch ops on the device."""
    v, g = lin.weight_v.detach(), lin.weight_g.detach()
    nrm = v.norm(dim=1, keepdim=True)
    vh = v / nrm
    dot = (dW * vh).sum(dim=1, keepdim=True)
    _acc(lin.weight_g, dot)
    _acc(lin.weight_v, g / nrm * (dW - dot * vh))


class PosteriorBackward:
    """Backward of the KL term and of the posterior head: prior flow (indm_prior_flow_bwd) -> reparameterisation
    (indm_posterior_bwd) -> weight-normed fc.  Produces d L / d (encoder output) and every prior / fc parameter gradient."""

    CPL = 32 + 4 * 256 + 64

    def __init__(self, eng):
        self.eng = eng
        N, dev = eng.N, eng.dev
        steps = eng.core.discriminator.prior.flow.steps
        self.ns = len(steps)
        self.ws_c = torch.empty((4 * self.ns, N * self.CPL), device=dev)
        self.ws_a = torch.empty((2 * self.ns, 2, N, 64), device=dev)
        self.ws_l = torch.empty((self.ns, 2, N, 64), device=dev)
        self.gh_prior = torch.empty((N, 64), device=dev)
        self.gc = torch.empty((N, 128), device=dev)
        self.g_enc = torch.empty((N, eng.enc['fc_w'].shape[1]), device=dev)

    @staticmethod
    def _sg(ta, tb, M, Nn, K, A, lda, B, ldb, beta, C, ldc):
        L.call('indm_sgemm_f32', ta, tb, M, Nn, K, _f(1.0), L.ptr(A), _i64(lda), L.ptr(B), _i64(ldb), _f(beta), L.ptr(C), _i64(ldc))

    @staticmethod
    def _colsum(x, rows, dim, out):
        L.call('indm_colsum', L.ptr(x), L.DTYPE_F32, _i64(1), _i64(rows), dim, _i64(dim), None, _i64(0), L.ptr(out), _f(1.0))

    def run(self, h, gh_blocks, cK, c, eps, enc_out):
        """h [N,64] the posterior sample, gh_blocks = d L / d h of the residual flow, cK [N] = d L / d KL, c [N,128] = fc output,
        eps the reparameterisation noise, enc_out [N,128] the encoder output.  Returns d L / d enc_out."""
        e = self.eng
        N = e.N
        cK = cK.float().contiguous()
        buf, n_ops, _ = e.prior_ops['forward']
        L.call('indm_prior_flow_bwd', L.ptr(h), L.ptr(e.prior_params), L.ptr(buf), n_ops, L.ptr(cK), L.ptr(self.ws_c), L.ptr(self.ws_a),
               L.ptr(self.ws_l), L.ptr(self.gh_prior), N)
        gh = gh_blocks + self.gh_prior
        L.call('indm_posterior_bwd', L.ptr(c), L.ptr(eps.float().contiguous()), L.ptr(gh), L.ptr(cK), L.ptr(self.gc), N)
        # ---- prior parameters
        steps = e.core.discriminator.prior.flow.steps
        dev = e.dev
        sum_ck = cK.sum()
        for j, st in enumerate(steps):
            for slot, an in ((2 * j, st.actnorm), (2 * j + 1, st.unit.actnorm)):
                g_ls, g_b = torch.zeros((64,), device=dev), torch.zeros((64,), device=dev)
                self._colsum(self.ws_a[slot, 0], N, 64, g_ls)
                self._colsum(self.ws_a[slot, 1], N, 64, g_b)
                _acc(an.log_scale, g_ls)
                _acc(an.bias, g_b)
            dW = torch.empty((64, 64), device=dev)
            self._sg(1, 0, 64, 64, N, self.ws_l[j, 1], 64, self.ws_l[j, 0], 64, 0.0, dW, 64)
            _acc(st.linear.weight, dW - sum_ck * e.prior_winvT[j])          # d (-ck log|det W|) / d W = -ck W^-T
            for q, cp in enumerate((st.unit.coupling1_up, st.unit.coupling1_dn, st.unit.coupling2_up, st.unit.coupling2_dn)):
                base = self.ws_c[4 * j + q]
                o = 0
                zin = base[o:o + N * 32].view(N, 32); o += N * 32
                ha = base[o:o + N * 256].view(N, 256); o += N * 256
                hb = base[o:o + N * 256].view(N, 256); o += N * 256
                d1 = base[o:o + N * 256].view(N, 256); o += N * 256
                d2 = base[o:o + N * 256].view(N, 256); o += N * 256
                d3 = base[o:o + N * 64].view(N, 64)
                net = cp.net
                dW1, dW2, dW3 = torch.empty((256, 32), device=dev), torch.empty((256, 256), device=dev), torch.empty((64, 256), device=dev)
                self._sg(1, 0, 256, 32, N, d1, 256, zin, 32, 0.0, dW1, 32)
                self._sg(1, 0, 256, 256, N, d2, 256, ha, 256, 0.0, dW2, 256)
                self._sg(1, 0, 64, 256, N, d3, 64, hb, 256, 0.0, dW3, 256)
                b1, b2, b3 = torch.zeros((256,), device=dev), torch.zeros((256,), device=dev), torch.zeros((64,), device=dev)
                self._colsum(d1, N, 256, b1); self._colsum(d2, N, 256, b2); self._colsum(d3, N, 64, b3)
                _acc(net.fc1.weight, dW1); _acc(net.fc1.bias, b1)
                _acc(net.fc2.weight, dW2); _acc(net.fc2.bias, b2)
                _wn_backward(net.fc3.linear, dW3); _acc(net.fc3.linear.bias, b3)
        # ---- posterior head: c = enc_out W^T + b with the weight-normed W (gaussian.py:21-25)
        fc = e.core.discriminator.fc.linear
        K = enc_out.shape[1]
        dWfc = torch.empty((128, K), device=dev)
        self._sg(1, 0, 128, K, N, self.gc, 128, enc_out, K, 0.0, dWfc, K)
        bfc = torch.zeros((128,), device=dev)
        self._colsum(self.gc, N, 128, bfc)
        _wn_backward(fc, dWfc)
        _acc(fc.bias, bfc)
        self._sg(0, 0, N, K, 128, self.gc, 128, e.enc['fc_w'], K, 0.0, self.g_enc, K)
        return self.g_enc
