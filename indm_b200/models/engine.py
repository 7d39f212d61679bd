"""Execution engine of the NCSN++ / DDPM++ score network on B200.

`ScoreEngine` turns an `NCSNpp` parameter container into a static launch plan over the C-ABI kernels
(include/indm_b200.h): activations live in NHWC (FP32 residual stream, BF16 — or tf32-rounded FP32 in validation mode
— tensor-core operands), weights are repacked once to [tap][Cout][Cin], every buffer is allocated up front so the plan
has fixed addresses and can be captured into a CUDA graph and replayed by the sampler.

Per res-block (models/layerspp.py:255-287) the plan is 5-6 launches instead of the reference's ~20:
    gn_stats(x)  ->  gn_apply(+SiLU, +nearest-up / mean-down, +raw bf16 copy)          [HBM-bound]
    igemm 3x3 (+bias, +Dense_0(SiLU(temb)) row bias, +fused GroupNorm_1 statistics)    [tcgen05]
    gn_apply(+SiLU)                                                                    [HBM-bound]
    igemm 3x3 (+ fused 1x1 skip conv as extra K iterations, + residual, * 1/sqrt(2))   [tcgen05]
All 44 Dense_0 layers are evaluated at once (one [N,4nf] x [sum Cout, 4nf] product, temb is shared).
"""
import ctypes
import math
import os

import numpy as np
import torch

from .. import _lib as L


class _Buf:
    """bump allocation record (all buffers are plain torch tensors kept alive by the engine)"""
    pass


class ScoreEngine:
    def __init__(self, model, batch, mode='bf16', device=None, pp=False):
        cfg = model.config
        # pp: inference-only plan whose small-feature-map residual blocks keep their convolution operands in the padded-pixel
        # layout (indm_igemm_t.a_pp): the 3x3 convolutions then read every activation once instead of once per tap
        self.infer = bool(pp) and mode == 'bf16'
        self.pp = self.infer and not os.environ.get('INDM_NO_PP')
        self.fused_attn = self.infer and not os.environ.get('INDM_NO_FUSED_ATTN')   # indm_attention_fwd (keeps no probabilities for a backward)
        self.fused_attn_blocks = 0
        self.pp_max_w = int(os.environ.get('INDM_PP_MAX_W', '4'))   # measured: faster on 4x4 maps, slower on 8x8 / 16x16 (border rows)
        self.pp_convs = 0         # convolutions of this plan that read padded-pixel operands
        self.model = model
        self.cfg = cfg
        self.N = int(batch)
        self.mode = mode
        assert mode in ('bf16', 'tf32')
        self.dt = L.DTYPE_BF16 if mode == 'bf16' else L.DTYPE_TF32
        self.tdtype = torch.bfloat16 if mode == 'bf16' else torch.float32
        self.kchunk = 64 if mode == 'bf16' else 32
        self.dev = device if device is not None else next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise RuntimeError('ScoreEngine needs a CUDA device: indm_b200 has no CPU path')
        L.lib()
        self.S = cfg.data.image_size
        self.ch = cfg.data.num_channels
        self.nf = cfg.model.nf
        self.fir = bool(cfg.model.fir)
        # split-K (indm_igemm_t.splitk_ws) is available for launches with few tiles but measured SLOWER on the 4x4 / 8x8 layers at
        # batch 128 (two launch floors + 18 MB of partial sums vs one 36-iteration CTA per tile): left off.  Set a workspace
        # tensor here (e.g. torch.empty(16 << 20, device=...)) to enable it.
        self.splitk_ws = None
        self.keep = []            # every tensor the plan points into
        self.ops = []             # list of zero-arg callables (forward plan)
        self.bops = None          # current backward plan, built on first use by build_backward()
        self._plans = {}          # {train flag: plan}
        self._train = False
        self.tape = []            # closures recorded by the forward builders; replayed in reverse to emit the backward plan
        self._cur = self.ops      # list the emit helpers append to
        self._producers = {}      # data_ptr of an fp32 NHWC tensor -> igemm descriptor that writes it
        self.pack_jobs = []       # (fn) re-run by load_weights()
        self.gn_slots = 0
        self._weights_version = None
        self._repack_graph, self._repack_key = None, None
        self.forward_count = 0    # bumped by every forward(): lets a pending backward detect overwritten activations
        self._drop_on = False
        self._bwd_graphs, self._bwd_ran = {}, set()   # {train flag: (CUDA graph of the backward plan, key)}; plans that ran eagerly once
        self._build()
        self.load_weights()

    # ------------------------------------------------------------------ helpers
    def _alloc(self, shape, dtype=torch.float32, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, device=self.dev, dtype=dtype)
        self.keep.append(t)
        return t

    def _op_t(self, shape):
        """tensor-core operand buffer (bf16, or fp32 in tf32 mode)"""
        return self._alloc(shape, self.tdtype)

    def _gn_slot(self):
        i = self.gn_slots
        self.gn_slots += 1
        return i

    def _round_op(self, w):
        """fp32 tensor -> operand dtype (bf16 / tf32-rounded fp32)"""
        if self.mode == 'bf16':
            return w.to(torch.bfloat16)
        return w.float().contiguous()     # TF32 mode keeps full fp32 operands: the GEMM kernel splits hi/lo itself (3xTF32)

    def _pack_conv(self, dst, conv_w, cin_pad=None):
        """[Cout, Cin, k, k] parameter -> dst [k*k][Cout][Cin_pad]"""
        def job():
            w = conv_w.detach().to(self.dev, torch.float32)
            co, ci, kh, kw = w.shape
            w = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
            if cin_pad is not None and cin_pad != ci:
                dst.zero_()
                dst[:, :, :ci].copy_(self._round_op(w))
            else:
                dst.copy_(self._round_op(w))
        self.pack_jobs.append(job)

    def _pack_f32(self, dst, srcs, transform=None):
        def job():
            parts = [s.detach().to(self.dev, torch.float32) for s in srcs]
            v = torch.cat([p.reshape(-1) if transform is None else transform(p).reshape(-1) for p in parts])
            dst.view(-1).copy_(v)
        self.pack_jobs.append(job)

    def load_weights(self):
        """(Re)pack all parameters into the engine's device buffers: one pass over ~62 M parameters.  The ~700 small packing jobs
        are captured into a CUDA graph the first time they run against a given set of parameter storages, so the per-step repack
        of training (parameters change every optimizer step) costs one graph replay instead of ~11 ms of Python."""
        key = (tuple(p.data_ptr() for p in self.model.parameters()), len(self.pack_jobs))
        if self._repack_graph is not None and self._repack_key == key:
            self._repack_graph.replay()
            L.debug_sync('replay of the score repack graph')
        else:
            with torch.no_grad():
                for job in self.pack_jobs:
                    job()
            self._repack_graph = None
            if not torch.cuda.is_current_stream_capturing():
                try:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        with torch.no_grad():
                            for job in self.pack_jobs:
                                job()
                    self._repack_graph, self._repack_key = g, key
                    L.debug_sync('capture of the score repack graph')
                except Exception:          # capture is an optimisation only: fall back to eager packing
                    self._repack_graph = None
        self._weights_version = self.weights_version()

    def weights_version(self):
        return sum(p._version for p in self.model.parameters()) + L.param_epoch

    # ------------------------------------------------------------------ plan construction
    def _igemm(self, **kw):
        d = L.IgemmDesc()
        d.scale = 1.0
        d.res_scale = 1.0
        d.dtype = self.dt
        if self.splitk_ws is not None:
            d.splitk_ws, d.splitk_ws_bytes = self.splitk_ws.data_ptr(), self.splitk_ws.numel() * 4
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                v = v.data_ptr()
            setattr(d, k, v)
        lib = L.lib()

        def run():
            L.check(lib.indm_igemm(ctypes.byref(d), L._stream()), 'igemm')
        self._cur.append(run)
        # fp32 NHWC outputs of the forward plan can accumulate the GroupNorm statistics of their consumers in the epilogue
        # (patched into this descriptor later by _gn): needs whole 32-column slabs and one image per epilogue warp
        out = kw.get('out_f32')
        if (self._cur is self.ops and isinstance(out, torch.Tensor) and kw.get('out_mode', 0) == 0 and not kw.get('batched_b')
                and kw.get('out_bf16') is None and kw['Cout'] % 32 == 0 and self.mode == 'bf16'
                and (kw['H'] * kw['W'] >= 32 or (kw.get('a_pp') and (kw['H'] + 1) * (kw['W'] + 2) >= 16))):
            self._producers[out.data_ptr()] = d
        return d

    def _fuse_stats_into_producers(self, xa, Ca, xb, Cb, G, part):
        """True if the producers of xa (and xb) now accumulate this GroupNorm's statistics in their epilogues"""
        C = Ca + Cb
        cpg = C // G
        if cpg not in (4, 8, 16, 32) or Ca % cpg != 0:
            return False
        prods = [(self._producers.get(xa.data_ptr()), 0)]
        if xb is not None:
            prods.append((self._producers.get(xb.data_ptr()), Ca // cpg))
        for d, _ in prods:
            if d is None or (d.gn_partial and d.gn2_partial):
                return False
        for d, goff in prods:
            if not d.gn_partial:
                d.gn_partial, d.gn_cpg, d.gn_groups, d.gn_goff = part.data_ptr(), cpg, G, goff
            else:
                d.gn2_partial, d.gn2_cpg, d.gn2_groups, d.gn2_goff = part.data_ptr(), cpg, G, goff
        return True

    def _call(self, name, *args):
        fn = getattr(L.lib(), name)
        cargs = [ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a for a in args]

        def run():
            L.check(fn(*cargs, L._stream()), name)
        self._cur.append(run)

    def _call_unless_dropping(self, name, args, drop_name, drop_args):
        """`name(*args)` when dropout is off for this forward (eval / sampling: the kernels instantiated without the Philox path
        keep 79 instead of 120 registers), `drop_name(*drop_args)` when forward(train=True) switched the masks on"""
        fn, fn_drop = getattr(L.lib(), name), getattr(L.lib(), drop_name)
        conv = lambda seq: [ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a for a in seq]
        cargs, dargs = conv(args), conv(drop_args)

        def run():
            if self._drop_on:
                L.check(fn_drop(*dargs, L._stream()), drop_name)
            else:
                L.check(fn(*cargs, L._stream()), name)
        self._cur.append(run)

    def _gn(self, xa, Ca, xb, Cb, in_dt, H, W, gparams, act, resample, want_raw, slot=None, stats_done=False, dropout=0.0, pp=False):
        """GroupNorm(+SiLU)(+resample) of concat(xa, xb) -> operand tensor (and optional raw copy of the input).
        Emits the statistics launch unless a producer already accumulated them into `slot`."""
        N = self.N
        C = Ca + Cb
        G = gparams.num_groups
        gamma, beta = self._alloc((C,)), self._alloc((C,))
        self._pack_f32(gamma, [gparams.weight])
        self._pack_f32(beta, [gparams.bias])
        if slot is None:
            slot = self._gn_slot()
        part = self.gn_part[slot]
        if not stats_done and not (in_dt == L.DTYPE_F32 and self._fuse_stats_into_producers(xa, Ca, xb, Cb, G, part)):
            self._call('indm_gn_stats', xa, Ca, xb, Cb, in_dt, ctypes.c_int64(N), ctypes.c_int64(H * W), G, part)
        Ho, Wo = (2 * H, 2 * W) if resample == 1 else ((H // 2, W // 2) if resample == 2 else (H, W))
        odt_ = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32
        if pp:
            # zero-bordered padded-pixel buffers: only interior rows are ever written, so the borders stay zero across forwards
            assert resample == 0
            rows = (N * (H + 1) + 1) * (W + 2)
            out = self._alloc((rows, C), self.tdtype, zero=True)
            raw = self._alloc((rows, C), self.tdtype, zero=True) if want_raw else None
            head = (xa, Ca, xb, Cb, in_dt, ctypes.c_int64(N), H, W, G, part, gamma, beta, ctypes.c_float(1e-6), act, out, raw, odt_)
            if dropout > 0.0:
                self._call_unless_dropping('indm_gn_apply_pp', head + (ctypes.c_float(0.0), None, ctypes.c_uint32(slot)),
                                           'indm_gn_apply_pp', head + (ctypes.c_float(dropout), self.drop_ctl, ctypes.c_uint32(slot)))
            else:
                self._call('indm_gn_apply_pp', *head, ctypes.c_float(0.0), None, ctypes.c_uint32(slot))
            self._last_gn = None
            return out, raw
        out = self._op_t((N, Ho, Wo, C))
        raw = self._op_t((N, Ho, Wo, C)) if want_raw else None
        if dropout > 0.0:
            assert resample == 0 and not want_raw
            head = (xa, Ca, xb, Cb, in_dt, ctypes.c_int64(N), H, W, G, part, gamma, beta, ctypes.c_float(1e-6), act)
            self._call_unless_dropping('indm_gn_apply', head + (0, out, None, odt_),
                                       'indm_gn_apply_dropout', head + (out, odt_, ctypes.c_float(dropout), self.drop_ctl, ctypes.c_uint32(slot)))
        else:
            self._call('indm_gn_apply', xa, Ca, xb, Cb, in_dt, ctypes.c_int64(N), H, W, G, part, gamma, beta, ctypes.c_float(1e-6),
                       act, resample, out, raw, odt_)
        self._last_gn = dict(xa=xa, Ca=Ca, xb=xb, Cb=Cb, in_dt=in_dt, H=H, W=W, G=G, part=part, gamma=gamma, beta=beta, act=act,
                             resample=resample, params=gparams, drop_p=float(dropout), drop_stream=int(slot))
        return out, raw

    def _build(self):
        cfg, m = self.cfg, self.model
        N, S, nf = self.N, self.S, self.nf
        mods = list(m.all_modules)
        # GroupNorm statistics: one [slots][N][32][2] buffer zeroed by a single memset per forward
        MAX_SLOTS = 2 * len(mods) + 8
        self.gn_part_all = self._alloc((MAX_SLOTS, N, 32, 2), zero=True)
        self.drop_ctl = self._alloc((2,), torch.int64, zero=True)     # {Philox seed, enabled}: read by the dropout kernels
        self._drop_seed = 0x5EED
        self.gn_part = [self.gn_part_all[i] for i in range(MAX_SLOTS)]
        part_all = self.gn_part_all

        def zero_stats():
            part_all.zero_()    # cudaMemsetAsync on the current stream
        self.ops.append(zero_stats)

        idx = 0
        # ---------------- time embedding + MLP + all Dense_0 at once
        self.time_cond = self._alloc((N,))
        self.sched = None          # optional device schedule table / step counter installed by the sampler
        emb_type = cfg.model.embedding_type.lower()
        if emb_type == 'fourier':
            fw = mods[idx]; idx += 1
            freqs = self._alloc((nf,))
            self._pack_f32(freqs, [fw.W])
            emb_dim, kind = 2 * nf, 1
        else:
            freqs, emb_dim, kind = None, nf, 0
        self.emb = self._alloc((N, emb_dim))
        self._temb_args = dict(freqs=freqs, kind=kind, emb_dim=emb_dim)
        self._temb_op_index = len(self.ops)
        self.ops.append(None)   # placeholder, bound in _bind_time_source()
        lin0, lin1 = mods[idx], mods[idx + 1]; idx += 2
        w0, b0 = self._alloc((4 * nf, emb_dim)), self._alloc((4 * nf,))
        w1, b1 = self._alloc((4 * nf, 4 * nf)), self._alloc((4 * nf,))
        self._pack_f32(w0, [lin0.weight]); self._pack_f32(b0, [lin0.bias])
        self._pack_f32(w1, [lin1.weight]); self._pack_f32(b1, [lin1.bias])
        # temb = Linear(SiLU(Linear(emb))) (models/ncsnpp.py:270-274); only SiLU(temb) is ever consumed (Dense_0), so the
        # second layer stores SiLU(temb) directly, in the tensor-core operand dtype
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32
        t0, temb_act = self._alloc((N, 4 * nf)), self._op_t((N, 4 * nf))
        self._call('indm_linear_f32', self.emb, w0, b0, t0, ctypes.c_int64(N), emb_dim, 4 * nf, 0, 1, L.DTYPE_F32)
        self._call('indm_linear_f32', t0, w1, b1, temb_act, ctypes.c_int64(N), 4 * nf, 4 * nf, 0, 1, op_dt)
        res_blocks = [mm for mm in mods if mm.__class__.__name__ == 'ResnetBlockBigGANpp']
        dense_total = sum(rb.out_ch for rb in res_blocks)
        self.dense_total = dense_total
        wd, bd = self._op_t((dense_total, 4 * nf)), self._alloc((dense_total,))

        def job_wd(wd=wd, res_blocks=res_blocks):
            wd.copy_(self._round_op(torch.cat([rb.Dense_0.weight.detach().to(self.dev, torch.float32) for rb in res_blocks], dim=0)))
        self.pack_jobs.append(job_wd)
        self._pack_f32(bd, [rb.Dense_0.bias for rb in res_blocks])
        self.dense_tab = self._alloc((N, dense_total))
        # all Dense_0 layers of the network in one tensor-core GEMM: [N, 4nf] x [sum Cout, 4nf]^T
        self._igemm(a=temb_act, N=1, H=1, W=N, Cin=4 * nf, b=wd, Cout=dense_total, taps=1, bias=bd, out_f32=self.dense_tab,
                    out_ld=dense_total)
        self.tape.append(('temb', dict(lin0=lin0, lin1=lin1, w0=w0, b0=b0, w1=w1, b1=b1, t0=t0, res_blocks=res_blocks, emb_dim=emb_dim)))
        dense_off = {}
        off = 0
        for rb in res_blocks:
            dense_off[id(rb)] = off
            off += rb.out_ch

        # ---------------- stem
        self.x_in = self._alloc((N, self.ch, S, S))
        cpad = self.kchunk
        x_nhwc = self._op_t((N, S, S, cpad))
        centered = bool(cfg.data.centered)
        self._call('indm_prep_input', self.x_in, x_nhwc, ctypes.c_int64(N), self.ch, S, S, cpad,
                   ctypes.c_float(1.0 if centered else 2.0), ctypes.c_float(0.0 if centered else -1.0), 0,
                   L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32)
        stem = mods[idx]; idx += 1
        wst = self._op_t((9, nf, cpad))
        self._pack_conv(wst, stem.weight, cin_pad=cpad)
        bst = self._alloc((nf,)); self._pack_f32(bst, [stem.bias])
        h0 = self._alloc((N, S, S, nf))
        self._igemm(a=x_nhwc, N=N, H=S, W=S, Cin=cpad, b=wst, Cout=nf, taps=9, bias=bst, out_f32=h0, out_ld=nf)
        self.tape.append(('stem', dict(conv=stem, h0=h0, mul=1.0 if centered else 2.0, x_nhwc=x_nhwc, cpad=cpad)))

        inv_sqrt2 = 1.0 / math.sqrt(2.0)

        def res_block(rb, xa, Ca, xb, Cb, H, W):
            """returns (out fp32 [N,H',W',Cout], H', W')"""
            Cin, Cout = rb.in_ch, rb.out_ch
            assert Cin == Ca + Cb, (Cin, Ca, Cb)
            has_skip = hasattr(rb, 'Conv_2')
            resample = 1 if rb.up else (2 if rb.down else 0)
            use_fir = self.fir and resample != 0
            use_pp = (self.pp and resample == 0 and W <= self.pp_max_w and Cin % 64 == 0 and Cout % 128 == 0)
            h1, raw = self._gn(xa, Ca, xb, Cb, L.DTYPE_F32, H, W, rb.GroupNorm_0, 1, 0 if use_fir else resample,
                               has_skip and not use_fir, pp=use_pp)
            gn0 = self._last_gn
            Ho, Wo = (2 * H, 2 * W) if resample == 1 else ((H // 2, W // 2) if resample == 2 else (H, W))
            if use_fir:
                assert xb is None, 'FIR resampling blocks never take a concatenated input'
                k1 = np.asarray(rb.fir_kernel, dtype=np.float32)
                k1 = k1 / k1.sum() * (2.0 if resample == 1 else 1.0)
                self.keep.append(k1)
                kptr = k1.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
                h1r = self._op_t((N, Ho, Wo, Cin))
                raw = self._op_t((N, Ho, Wo, Cin))
                odt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32
                self._call('indm_fir_nhwc', h1, h1r, L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32, odt, ctypes.c_int64(N), H, W,
                           Cin, kptr, resample)
                self._call('indm_fir_nhwc', xa, raw, L.DTYPE_F32, odt, ctypes.c_int64(N), H, W, Cin, kptr, resample)
                h1 = h1r
            # conv0 (+bias +temb row bias) -> h2 (operand dtype), GroupNorm_1 statistics fused when the tile allows
            w0 = self._op_t((9, Cout, Cin)); self._pack_conv(w0, rb.Conv_0.weight)
            b0 = self._alloc((Cout,)); self._pack_f32(b0, [rb.Conv_0.bias])
            h2 = self._op_t((N, Ho, Wo, Cout))
            G1 = rb.GroupNorm_1.num_groups
            cpg1 = Cout // G1
            px = Ho * Wo
            fuse_stats = ((32 % cpg1 == 0) and (Cout % 32 == 0) and self.mode == 'bf16'
                          and (px >= 32 or (use_pp and (Ho + 1) * (Wo + 2) >= 16)))      # the padded-pixel kernel also covers 4x4 maps
            slot1 = self._gn_slot()
            kw = dict(a=h1, N=N, H=Ho, W=Wo, Cin=Cin, b=w0, Cout=Cout, taps=9, bias=b0,
                      rowbias=self.dense_tab[:, dense_off[id(rb)]:], rowbias_ld=self.dense_total, out_ld=Cout)
            if self.mode == 'bf16':
                kw['out_bf16'] = h2
            else:
                kw['out_f32'] = h2
            if fuse_stats:
                kw.update(gn_partial=self.gn_part[slot1], gn_cpg=cpg1, gn_groups=G1)
            if use_pp:
                kw['a_pp'] = 1
                self.pp_convs += 2
            self._igemm(**kw)
            in_dt1 = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
            h3, _ = self._gn(h2, Cout, None, 0, in_dt1, Ho, Wo, rb.GroupNorm_1, 1, 0, False, slot=slot1, stats_done=fuse_stats,
                             dropout=float(rb.dropout), pp=use_pp)
            gn1 = self._last_gn
            # conv1 (+ fused skip 1x1 | + residual), * 1/sqrt(2)
            w1 = self._op_t((9, Cout, Cout)); self._pack_conv(w1, rb.Conv_1.weight)
            b1 = self._alloc((Cout,))
            out = self._alloc((N, Ho, Wo, Cout))
            kw = dict(a=h3, N=N, H=Ho, W=Wo, Cin=Cout, b=w1, Cout=Cout, taps=9, bias=b1, scale=inv_sqrt2 if rb.skip_rescale else 1.0,
                      out_f32=out, out_ld=Cout)
            if has_skip:
                w2 = self._op_t((Cout, Cin)); self._pack_conv(w2.view(1, Cout, Cin), rb.Conv_2.weight)

                def job(b1=b1, rb=rb):   # one bias vector for both K segments
                    b1.copy_((rb.Conv_1.bias.detach() + rb.Conv_2.bias.detach()).to(self.dev, torch.float32))
                self.pack_jobs.append(job)
                kw.update(a2=raw, Cin2=Cin, b2=w2)
            else:
                self._pack_f32(b1, [rb.Conv_1.bias])
                assert xb is None and Ca == Cout and resample == 0
                kw.update(residual=xa, res_ld=Cout, res_scale=inv_sqrt2 if rb.skip_rescale else 1.0)
            if use_pp:
                kw['a_pp'] = 1
            self._igemm(**kw)
            self.tape.append(('res_block', dict(rb=rb, gn0=gn0, gn1=gn1, out=out, Cin=Cin, Cout=Cout, Ho=Ho, Wo=Wo, has_skip=has_skip,
                                                h1=h1, h3=h3, raw=raw, dense_off=dense_off[id(rb)], H=H, W=W, resample=resample,
                                                fir_k1=(k1 if use_fir else None),
                                                use_fir=use_fir, s=inv_sqrt2 if rb.skip_rescale else 1.0)))
            return out, Ho, Wo

        def attn_block(ab, x, C, H, W):
            Lq = H * W
            h, _ = self._gn(x, C, None, 0, L.DTYPE_F32, H, W, ab.GroupNorm_0, 0, 0, False)
            gna = self._last_gn
            wqkv = self._op_t((3 * C, C))
            bqkv = self._alloc((3 * C,))

            def job(wqkv=wqkv, bqkv=bqkv, ab=ab):
                ws = [getattr(ab, f'NIN_{j}').W.detach().to(self.dev, torch.float32).t() for j in range(3)]
                wqkv.copy_(self._round_op(torch.cat(ws, dim=0)))
                bqkv.copy_(torch.cat([getattr(ab, f'NIN_{j}').b.detach().to(self.dev, torch.float32) for j in range(3)]))
            self.pack_jobs.append(job)
            # q | k | v rows in one GEMM (N = 3C); V^T (the K-major B operand of P.V) by a batched transpose of the v columns
            qkv = self._op_t((N, Lq, 3 * C))
            self._igemm(a=h, N=N, H=H, W=W, Cin=C, b=wqkv, Cout=3 * C, taps=1, bias=bqkv, **self._okw(qkv, 3 * C))
            if self.fused_attn and Lq == 256 and C == 256:
                # forward-only plan: scores in TMEM, probabilities in shared memory, V read in place — one launch
                o = self._op_t((N, Lq, C))
                self._call('indm_attention_fwd', qkv, o, ctypes.c_int64(N), Lq, C, ctypes.c_float(float(int(C) ** (-0.5))), L.DTYPE_BF16)
                self.fused_attn_blocks += 1
                return attn_out(ab, gna, x, C, H, W, qkv, None, h, o)
            qk = qkv                                    # q at columns [0, C), k at [C, 2C): row stride 3C
            vt = self._op_t((N, C, Lq))
            op_dt_ = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
            self._call('indm_transpose_batched', qkv[:, :, 2 * C:], vt, ctypes.c_int64(N), Lq, C, ctypes.c_int64(3 * C),
                       ctypes.c_int64(Lq * 3 * C), op_dt_)
            s = self._alloc((N, Lq, Lq))
            self._igemm(a=qk, a_ld=3 * C, a_img_stride=Lq * 3 * C, N=N, H=1, W=Lq, Cin=C, b=qk[:, :, C:], b_ld=3 * C,
                        b_tap_stride=Lq * 3 * C, Cout=Lq, taps=1, batched_b=1, scale=float(int(C) ** (-0.5)), out_f32=s, out_ld=Lq)
            p = self._op_t((N, Lq, Lq))
            self._call('indm_softmax_rows', s, p, ctypes.c_int64(N * Lq), Lq, L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32)
            o = self._op_t((N, Lq, C))
            kw = dict(a=p, N=N, H=1, W=Lq, Cin=Lq, b=vt, Cout=C, taps=1, batched_b=1, out_ld=C)
            if self.mode == 'bf16':
                kw['out_bf16'] = o
            else:
                kw.update(out_f32=o, round_tf32_out=1)
            self._igemm(**kw)
            return attn_out(ab, gna, x, C, H, W, qkv, p, h, o)

        def attn_out(ab, gna, x, C, H, W, qkv, p, h, o):
            w3 = self._op_t((C, C))
            b3 = self._alloc((C,))

            def job3(w3=w3, b3=b3, ab=ab):
                w3.copy_(self._round_op(ab.NIN_3.W.detach().to(self.dev, torch.float32).t().contiguous()))
                b3.copy_(ab.NIN_3.b.detach().to(self.dev, torch.float32))
            self.pack_jobs.append(job3)
            out = self._alloc((N, H, W, C))
            self._igemm(a=o, N=N, H=H, W=W, Cin=C, b=w3, Cout=C, taps=1, bias=b3, residual=x, res_ld=C,
                        scale=inv_sqrt2 if ab.skip_rescale else 1.0, res_scale=inv_sqrt2 if ab.skip_rescale else 1.0,
                        out_f32=out, out_ld=C)
            self.tape.append(('attn', dict(ab=ab, gn=gna, x=x, C=C, H=H, W=W, qkv=qkv, p=p, out=out, h=h, o=o,
                                           s=inv_sqrt2 if ab.skip_rescale else 1.0)))
            return out

        # ---------------- down path
        nlev = len(cfg.model.ch_mult)
        attn_res = tuple(cfg.model.attn_resolutions)
        pin = cfg.model.progressive_input.lower()
        odt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_TF32

        def pyramid_block(ds, pyr, pyr_c, pyr_dt, Hi, Wi, h, Cout):
            """input pyramid of progressive_input='residual' (models/ncsnpp.py:319-326): FIR pad-(2,2) filter
            -> 3x3 stride-2 VALID conv (+bias) -> (pyramid + h) / sqrt(2), written as the new fp32 residual stream."""
            k1 = np.asarray(ds.fir_kernel, dtype=np.float32)
            k1 = k1 / k1.sum()
            self.keep.append(k1)
            fir_out = self._op_t((N, Hi + 1, Wi + 1, pyr_c))
            self._call('indm_fir_nhwc', pyr, fir_out, pyr_dt, odt, ctypes.c_int64(N), Hi, Wi, pyr_c,
                       k1.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 3)
            wp = self._op_t((9, Cout, pyr_c)); self._pack_conv(wp, ds.Conv2d_0.weight, cin_pad=pyr_c)
            bp = self._alloc((Cout,)); self._pack_f32(bp, [ds.Conv2d_0.bias])
            out = self._alloc((N, Hi // 2, Wi // 2, Cout))
            sc = inv_sqrt2 if m.skip_rescale else 1.0
            self._igemm(a=fir_out, N=N, H=Hi // 2, W=Wi // 2, Cin=pyr_c, b=wp, Cout=Cout, taps=9, stride=2, bias=bp, scale=sc,
                        residual=h, res_ld=Cout, res_scale=sc, out_f32=out, out_ld=Cout)
            self.tape.append(('pyramid', dict(ds=ds, pyr=pyr, pyr_c=pyr_c, pyr_dt=pyr_dt, Hi=Hi, Wi=Wi, h=h, Cout=Cout, out=out, sc=sc,
                                              fir_out=fir_out, k1=k1, first=(pyr is x_nhwc))))
            return out

        # the pyramid starts from the network input (after the 2x-1 affine): the padded NHWC operand copy of it
        pyr, pyr_c, pyr_dt = x_nhwc, cpad, (L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32)
        hs = [(h0, nf)]
        H = W = S
        for lv in range(nlev):
            for _ in range(cfg.model.num_res_blocks):
                x, Cx = hs[-1]
                rb = mods[idx]; idx += 1
                h, H, W = res_block(rb, x, Cx, None, 0, H, W)
                Ch = rb.out_ch
                if H in attn_res and cfg.model.attention:
                    h = attn_block(mods[idx], h, Ch, H, W); idx += 1
                hs.append((h, Ch))
            if lv != nlev - 1:
                x, Cx = hs[-1]
                rb = mods[idx]; idx += 1
                Hi, Wi = H, W
                h, H, W = res_block(rb, x, Cx, None, 0, H, W)
                if pin == 'residual':
                    ds = mods[idx]; idx += 1
                    h = pyramid_block(ds, pyr, pyr_c, pyr_dt, Hi, Wi, h, rb.out_ch)
                    pyr, pyr_c, pyr_dt = h, rb.out_ch, L.DTYPE_F32
                elif pin != 'none':
                    raise NotImplementedError(f'progressive_input={pin!r} is not used by any INDM config')
                hs.append((h, rb.out_ch))
        h, Ch = hs[-1]
        rb = mods[idx]; idx += 1
        h, H, W = res_block(rb, h, Ch, None, 0, H, W)
        h = attn_block(mods[idx], h, Ch, H, W); idx += 1
        rb = mods[idx]; idx += 1
        h, H, W = res_block(rb, h, Ch, None, 0, H, W)
        # ---------------- up path
        for lv in reversed(range(nlev)):
            for _ in range(cfg.model.num_res_blocks + 1):
                skip, Cs = hs.pop()
                rb = mods[idx]; idx += 1
                h, H, W = res_block(rb, h, Ch, skip, Cs, H, W)
                Ch = rb.out_ch
            if H in attn_res and cfg.model.attention:
                h = attn_block(mods[idx], h, Ch, H, W); idx += 1
            if lv != 0:
                rb = mods[idx]; idx += 1
                h, H, W = res_block(rb, h, Ch, None, 0, H, W)
                Ch = rb.out_ch
        assert not hs
        # ---------------- head: GroupNorm + SiLU + conv3x3 -> NCHW fp32, optional per-sample output scale
        gnh = mods[idx]; idx += 1
        hh, _ = self._gn(h, Ch, None, 0, L.DTYPE_F32, H, W, gnh, 1, 0, False)
        gn_head = self._last_gn
        head = mods[idx]; idx += 1
        assert idx == len(mods)
        wh = self._op_t((9, self.ch, Ch)); self._pack_conv(wh, head.weight)
        bh = self._alloc((self.ch,)); self._pack_f32(bh, [head.bias])
        self.out = self._alloc((N, self.ch, S, S))
        self.out_scale = self._alloc((N,))
        self.out_scale.fill_(1.0)
        self._igemm(a=hh, N=N, H=H, W=W, Cin=Ch, b=wh, Cout=self.ch, taps=9, bias=bh, rowscale=self.out_scale, out_mode=1,
                    out_f32=self.out)
        self.tape.append(('head', dict(conv=head, gn=gn_head, Ch=Ch, H=H, W=W, hh=hh)))
        self._bind_time_source(None, None, 0, 0)

    def _bind_time_source(self, sched, step, sched_ld, sched_col):
        """Build the time-embedding launch that reads `self.time_cond[n]` (sched is None: installed as the plan's default) or a
        device schedule table row `*step`.  The schedule-reading launch is only RETURNED: the sampler passes it to
        `launch(temb_op=...)` for the launches it captures, so an ordinary `forward(x, t)` on the same (cached) engine is always
        conditioned on `t` (it used to stay bound to the sampler's table after any PC sampling call)."""
        a = self._temb_args
        fn = L.lib().indm_time_embedding
        args = [ctypes.c_void_p(self.time_cond.data_ptr()),
                ctypes.c_void_p(sched.data_ptr()) if sched is not None else None,
                ctypes.c_void_p(step.data_ptr()) if step is not None else None, sched_ld, sched_col,
                ctypes.c_void_p(a['freqs'].data_ptr()) if a['freqs'] is not None else None, a['kind'],
                ctypes.c_int64(self.N), a['emb_dim'], ctypes.c_void_p(self.emb.data_ptr())]

        def run():
            L.check(fn(*args, L._stream()), 'indm_time_embedding')
        self.keep.append((sched, step))
        if sched is None:
            self.ops[self._temb_op_index] = run
        return run


    # ------------------------------------------------------------------ backward (input-VJP) plan
    # Built lazily from `self.tape`, in reverse.  Gradients of the fp32 residual stream are fp32 buffers (first writer
    # overwrites, later writers accumulate — decided here, statically); gradients that feed a dgrad GEMM are produced
    # directly in the tensor-core operand dtype.  Replaces the autograd graph of likelihood.py:27-38 / losses.py:250.
    def _grad_of(self, t):
        """(fp32 gradient buffer of residual-stream tensor t, accumulate flag); marks it written."""
        k = t.data_ptr()
        g = self._grads.get(k)
        if g is None:
            g = self._alloc(tuple(t.shape))
            self._grads[k] = g
        acc = 1 if k in self._gwritten else 0
        self._gwritten.add(k)
        return g, acc

    def _pack_dgrad(self, conv_w, kpad=None):
        """[Cout, Cin, k, k] parameter -> dgrad operand [k*k (taps flipped)][Cin][Cout (padded to kpad)]"""
        co, ci, kh, kw = conv_w.shape
        dst = self._op_t((kh * kw, ci, kpad or co))

        def job():
            w = conv_w.detach().to(self.dev, torch.float32)
            w = w.flip(2, 3).permute(2, 3, 1, 0).reshape(kh * kw, ci, co)
            if kpad and kpad != co:
                dst.zero_()
                dst[:, :, :co].copy_(self._round_op(w))
            else:
                dst.copy_(self._round_op(w))
        self.pack_jobs.append(job)
        job_now = job
        with torch.no_grad():
            job_now()
        return dst

    # parameter gradients (training plan only): accumulated straight into `param.grad` storage, in the parameter's own layout
    def _pgrad(self, param):
        if param.grad is None:
            param.grad = torch.zeros_like(param)
        g = param.grad
        if not g.is_contiguous() or g.dtype != torch.float32 or g.device != self.dev:
            raise RuntimeError('indm_b200: parameter gradients must be contiguous fp32 tensors on the engine device')
        self._pgrad_ptrs.append((param, g.data_ptr()))
        return g

    def _wgrad(self, dy, dy_ld, x, x_ld, H, W, Cout, Cin, taps, weight, strides=None):
        if not self._train or not weight.requires_grad:
            return
        g = self._pgrad(weight)
        so, sc, st = strides if strides is not None else (Cin * taps, taps, 1)
        self._call('indm_conv_wgrad', dy, ctypes.c_int64(dy_ld), x, ctypes.c_int64(x_ld), L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32,
                   self.N, H, W, Cout, Cin, taps, g, ctypes.c_int64(so), ctypes.c_int64(sc), ctypes.c_int64(st), ctypes.c_float(1.0))

    def _bgrad(self, dy, P, C, ld, biases=(), img_out=None, img_ld=0):
        """bias gradients (sum over images and pixels of dy) and / or the per-image sums that feed the time-embedding path"""
        if not self._train:
            return
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
        outs = [self._pgrad(b) for b in biases if b.requires_grad]
        first = outs[0] if outs else None
        if first is not None or img_out is not None:
            self._call('indm_colsum', dy, op_dt, ctypes.c_int64(self.N), ctypes.c_int64(P), C, ctypes.c_int64(ld), img_out, ctypes.c_int64(img_ld),
                       first, ctypes.c_float(1.0))
        for extra in outs[1:]:
            self._call('indm_colsum', dy, op_dt, ctypes.c_int64(self.N), ctypes.c_int64(P), C, ctypes.c_int64(ld), None, ctypes.c_int64(0),
                       extra, ctypes.c_float(1.0))

    def _gn_bwd(self, gn, dy, extra_post=None, extra_pre=None, extra_scale=0.0, to_operand=False):
        """emit stats + apply of the GroupNorm(+SiLU)(+resample) backward; returns the operand-dtype gradient if to_operand"""
        N = self.N
        slot = self._gnb_slots
        self._gnb_slots += 1
        assert slot < self.gnb_part_all.shape[0]
        pb = self.gnb_part_all[slot]
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
        x_dt = gn['in_dt']
        if to_operand:
            dxa, acc_a, dxb, acc_b = self._op_t(tuple(gn['xa'].shape)), 0, None, 0
            out_dt = op_dt
        else:
            dxa, acc_a = self._grad_of(gn['xa'])
            dxb, acc_b = self._grad_of(gn['xb']) if gn['xb'] is not None else (None, 0)
            out_dt = L.DTYPE_F32
        common = (dy, op_dt, gn['xa'], gn['Ca'], gn['xb'], gn['Cb'], x_dt, ctypes.c_int64(N), gn['H'], gn['W'], gn['G'], gn['part'],
                  gn['gamma'], gn['beta'], ctypes.c_float(1e-6), gn['act'], gn['resample'])
        dgam = dbet = None
        if self._train and gn['params'].weight.requires_grad:
            dgam, dbet = self._pgrad(gn['params'].weight), self._pgrad(gn['params'].bias)
        drop = (ctypes.c_float(gn['drop_p']), self.drop_ctl if gn['drop_p'] > 0 else None, ctypes.c_uint32(gn['drop_stream']))
        self._call('indm_gn_bwd_stats', *common, pb, dgam, dbet, out_dt, *drop)
        self._call('indm_gn_bwd_apply', *common, pb, extra_post, extra_pre, ctypes.c_float(extra_scale), dxa, acc_a, dxb, acc_b, out_dt, *drop)
        return dxa

    def _cast_grad(self, t, scale):
        """operand-dtype copy of scale * grad(t) (t must already have its full gradient)"""
        g = self._grads[t.data_ptr()]
        out = self._op_t(tuple(t.shape))
        self._call('indm_cast_scale', g, out, ctypes.c_int64(g.numel()), ctypes.c_float(scale),
                   L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32)
        return out

    def _okw(self, t, ld):
        return dict(out_bf16=t, out_ld=ld) if self.mode == 'bf16' else dict(out_f32=t, out_ld=ld)

    def _bwd_res_block(self, r):
        N, rb = self.N, r['rb']
        Cin, Cout, Ho, Wo = r['Cin'], r['Cout'], r['Ho'], r['Wo']
        g = self._cast_grad(r['out'], r['s'])                                   # d(out) * 1/sqrt(2), operand dtype
        w1d = self._pack_dgrad(rb.Conv_1.weight)
        d_h3 = self._op_t((N, Ho, Wo, Cout))
        self._igemm(a=g, N=N, H=Ho, W=Wo, Cin=Cout, b=w1d, Cout=Cout, taps=9, **self._okw(d_h3, Cout))
        self._wgrad(g, Cout, r['h3'], Cout, Ho, Wo, Cout, Cout, 9, rb.Conv_1.weight)
        self._bgrad(g, Ho * Wo, Cout, Cout, biases=[rb.Conv_1.bias] + ([rb.Conv_2.bias] if r['has_skip'] else []))
        d_h2 = self._gn_bwd(r['gn1'], d_h3, to_operand=True)
        self._wgrad(d_h2, Cout, r['h1'], Cin, Ho, Wo, Cout, Cin, 9, rb.Conv_0.weight)
        self._bgrad(d_h2, Ho * Wo, Cout, Cout, biases=[rb.Conv_0.bias],
                    img_out=self.d_dense_tab[:, r['dense_off']:] if self._train else None, img_ld=self.dense_total)
        w0d = self._pack_dgrad(rb.Conv_0.weight)
        d_h1 = self._op_t((N, Ho, Wo, Cin))
        self._igemm(a=d_h2, N=N, H=Ho, W=Wo, Cin=Cout, b=w0d, Cout=Cin, taps=9, **self._okw(d_h1, Cin))
        if r['use_fir']:
            # FIR-resampled block (model.fir=True, models/layerspp.py:256-271): both branches pass through upfirdn2d, whose
            # transpose is upfirdn2d with up <-> down and the same (symmetric) taps (op/upfirdn2d.py:111-114)
            op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
            H, W = r['H'], r['W']
            tmode = 2 if r['resample'] == 1 else 1
            kptr = r['fir_k1'].ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            d_h1s = self._op_t((N, H, W, Cin))
            self._call('indm_fir_nhwc', d_h1, d_h1s, op_dt, op_dt, ctypes.c_int64(N), Ho, Wo, Cin, kptr, tmode)
            w2d = self._pack_dgrad(rb.Conv_2.weight)
            d_raw = self._alloc((N, Ho, Wo, Cin))
            self._igemm(a=g, N=N, H=Ho, W=Wo, Cin=Cout, b=w2d, Cout=Cin, taps=1, out_f32=d_raw, out_ld=Cin)
            self._wgrad(g, Cout, r['raw'], Cin, Ho, Wo, Cout, Cin, 1, rb.Conv_2.weight, strides=(Cin, 1, 0))
            d_xs = self._alloc((N, H, W, Cin))
            self._call('indm_fir_nhwc', d_raw, d_xs, L.DTYPE_F32, L.DTYPE_F32, ctypes.c_int64(N), Ho, Wo, Cin, kptr, tmode)
            self._gn_bwd(r['gn0'], d_h1s, extra_pre=d_xs, extra_scale=1.0)
        elif r['has_skip']:
            w2d = self._pack_dgrad(rb.Conv_2.weight)
            d_raw = self._alloc((N, Ho, Wo, Cin))
            self._igemm(a=g, N=N, H=Ho, W=Wo, Cin=Cout, b=w2d, Cout=Cin, taps=1, out_f32=d_raw, out_ld=Cin)
            self._wgrad(g, Cout, r['raw'], Cin, Ho, Wo, Cout, Cin, 1, rb.Conv_2.weight, strides=(Cin, 1, 0))
            self._gn_bwd(r['gn0'], d_h1, extra_post=d_raw)
        else:
            self._gn_bwd(r['gn0'], d_h1, extra_pre=self._grads[r['out'].data_ptr()], extra_scale=r['s'])

    def _bwd_attn(self, r):
        N, ab, C, H, W = self.N, r['ab'], r['C'], r['H'], r['W']
        Lq = H * W
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
        g = self._cast_grad(r['out'], r['s'])
        w3d = self._op_t((C, C))

        def job3(w3d=w3d, ab=ab):          # NIN_3.W is [in, out]: exactly the dgrad operand [N = in][K = out]
            w3d.copy_(self._round_op(ab.NIN_3.W.detach().to(self.dev, torch.float32).contiguous()))
        self.pack_jobs.append(job3)
        with torch.no_grad():
            job3()
        d_o = self._op_t((N, Lq, C))
        self._igemm(a=g, N=N, H=H, W=W, Cin=C, b=w3d, Cout=C, taps=1, **self._okw(d_o, C))
        # NIN weights are [in, out]: dW[in][out] = sum_p x[p][in] dy[p][out]  ->  strides (o: 1, c: out)
        self._wgrad(g, C, r['o'], C, H, W, C, C, 1, ab.NIN_3.W, strides=(1, C, 0))
        self._bgrad(g, Lq, C, C, biases=[ab.NIN_3.b])
        # row-major V from the forward; transposed Q, K (and V, unused) from the fused q|k|v rows
        qkv = r['qkv']
        qkvT = self._op_t((N, 3 * C, Lq))
        self._call('indm_transpose_batched', qkv, qkvT, ctypes.c_int64(N), Lq, 3 * C, ctypes.c_int64(0), ctypes.c_int64(0), op_dt)
        qkT = qkvT[:, :2 * C, :]
        d_p = self._alloc((N, Lq, Lq))
        self._igemm(a=d_o, N=N, H=1, W=Lq, Cin=C, b=qkv[:, :, 2 * C:], b_ld=3 * C, b_tap_stride=Lq * 3 * C, Cout=Lq, taps=1, batched_b=1,
                    out_f32=d_p, out_ld=Lq)
        d_s = self._op_t((N, Lq, Lq))
        self._call('indm_softmax_bwd_rows', d_p, r['p'], d_s, ctypes.c_int64(N * Lq), Lq, ctypes.c_float(float(int(C) ** (-0.5))), op_dt)
        d_qkv = self._op_t((N, Lq, 3 * C))
        # dQ = dS K
        self._igemm(a=d_s, N=N, H=1, W=Lq, Cin=Lq, b=qkT[:, C:, :], b_ld=Lq, b_tap_stride=3 * C * Lq, Cout=C, taps=1, batched_b=1,
                    **self._okw(d_qkv, 3 * C))
        # dK = dS^T Q
        d_sT = self._op_t((N, Lq, Lq))
        self._call('indm_transpose_batched', d_s, d_sT, ctypes.c_int64(N), Lq, Lq, ctypes.c_int64(0), ctypes.c_int64(0), op_dt)
        self._igemm(a=d_sT, N=N, H=1, W=Lq, Cin=Lq, b=qkT, b_ld=Lq, b_tap_stride=3 * C * Lq, Cout=C, taps=1, batched_b=1,
                    **self._okw(d_qkv[:, :, C:], 3 * C))
        # dV = P^T dO
        pT = self._op_t((N, Lq, Lq))
        self._call('indm_transpose_batched', r['p'], pT, ctypes.c_int64(N), Lq, Lq, ctypes.c_int64(0), ctypes.c_int64(0), op_dt)
        d_oT = self._op_t((N, C, Lq))
        self._call('indm_transpose_batched', d_o, d_oT, ctypes.c_int64(N), Lq, C, ctypes.c_int64(0), ctypes.c_int64(0), op_dt)
        self._igemm(a=pT, N=N, H=1, W=Lq, Cin=Lq, b=d_oT, b_ld=Lq, b_tap_stride=C * Lq, Cout=C, taps=1, batched_b=1,
                    **self._okw(d_qkv[:, :, 2 * C:], 3 * C))
        # back through the fused q/k/v projection
        wqkvd = self._op_t((C, 3 * C))

        def jobq(wqkvd=wqkvd, ab=ab):
            ws = [getattr(ab, f'NIN_{j}').W.detach().to(self.dev, torch.float32) for j in range(3)]    # [in, out] each
            wqkvd.copy_(self._round_op(torch.cat(ws, dim=1)))
        self.pack_jobs.append(jobq)
        with torch.no_grad():
            jobq()
        d_h = self._op_t((N, Lq, C))
        self._igemm(a=d_qkv, N=N, H=H, W=W, Cin=3 * C, b=wqkvd, Cout=C, taps=1, **self._okw(d_h, C))
        for jq in range(3):
            nin = getattr(ab, f'NIN_{jq}')
            self._wgrad(d_qkv[:, :, jq * C:], 3 * C, r['h'], C, H, W, C, C, 1, nin.W, strides=(1, C, 0))
            self._bgrad(d_qkv[:, :, jq * C:], Lq, C, 3 * C, biases=[nin.b])
        self._gn_bwd(r['gn'], d_h, extra_pre=self._grads[r['out'].data_ptr()], extra_scale=r['s'])

    def _bwd_head(self, r):
        N, S, Ch = self.N, self.S, r['Ch']
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
        cpad = self.kchunk
        self.gout = self._alloc((N, self.ch, S, S), zero=True)          # cotangent of the network output (NCHW fp32)
        g = self._op_t((N, S, S, cpad))
        self._call('indm_nchw_to_nhwc', self.gout, self.out_scale, g, ctypes.c_int64(N), self.ch, S, S, cpad, ctypes.c_float(1.0), op_dt)
        whd = self._pack_dgrad(r['conv'].weight, kpad=cpad)
        d_hh = self._op_t((N, S, S, Ch))
        self._igemm(a=g, N=N, H=S, W=S, Cin=cpad, b=whd, Cout=Ch, taps=9, **self._okw(d_hh, Ch))
        self._wgrad(g, cpad, r['hh'], Ch, S, S, self.ch, Ch, 9, r['conv'].weight)
        if self._train and r['conv'].bias.requires_grad:
            tmp = self._alloc((cpad,), zero=True)      # colsum works on whole channel quads: sum the padded row, keep the first `ch`
            self._scratch_zero.append(tmp)
            self._call('indm_colsum', g, L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32, ctypes.c_int64(N), ctypes.c_int64(S * S), cpad,
                       ctypes.c_int64(cpad), None, ctypes.c_int64(0), tmp, ctypes.c_float(1.0))
            self._call('indm_axpy_f32', self._pgrad(r['conv'].bias), tmp, ctypes.c_float(1.0), ctypes.c_int64(self.ch))
        self._gn_bwd(r['gn'], d_hh)

    def _bwd_stem(self, r):
        N, S = self.N, self.S
        g = self._cast_grad(r['h0'], 1.0)
        wsd = self._pack_dgrad(r['conv'].weight)
        self.gx = self._alloc((N, self.ch, S, S))
        kw = {}
        if getattr(self, '_pyr_gx', None) is not None:
            # VE: the input pyramid also reads the (affinely rescaled) network input
            pg = self._alloc((N, self.ch, S, S))
            self._call('indm_nhwc_to_nchw_f32', self._pyr_gx, ctypes.c_int64(r['cpad']), pg, ctypes.c_int64(N), self.ch, S, S, ctypes.c_float(r['mul']))
            kw = dict(residual=pg, res_scale=1.0)
        self._igemm(a=g, N=N, H=S, W=S, Cin=self.nf, b=wsd, Cout=self.ch, taps=9, scale=r['mul'], out_mode=1, out_f32=self.gx, **kw)
        self._wgrad(g, self.nf, r['x_nhwc'], r['cpad'], S, S, self.nf, self.ch, 9, r['conv'].weight)
        self._bgrad(g, S * S, self.nf, self.nf, biases=[r['conv'].bias])

    def _bwd_temb(self, r):
        """time-embedding MLP and the 44 Dense_0 layers (models/ncsnpp.py:270-274, models/layerspp.py:276): small fp32 GEMMs"""
        if not self._train:
            return
        N, nf4, tot = self.N, 4 * self.nf, self.dense_total
        f32 = lambda *shape: self._alloc(shape)
        # fp32 recomputation of the pre-activations (the forward keeps only SiLU outputs, in the operand dtype)
        pre0, pre1, tact = f32(N, nf4), f32(N, nf4), f32(N, nf4)
        self._call('indm_linear_f32', self.emb, r['w0'], r['b0'], pre0, ctypes.c_int64(N), r['emb_dim'], nf4, 0, 0, L.DTYPE_F32)
        self._call('indm_linear_f32', r['t0'], r['w1'], r['b1'], pre1, ctypes.c_int64(N), nf4, nf4, 0, 0, L.DTYPE_F32)
        self._call('indm_linear_f32', r['t0'], r['w1'], r['b1'], tact, ctypes.c_int64(N), nf4, nf4, 0, 1, L.DTYPE_F32)
        wd32 = f32(tot, nf4)

        def job(wd32=wd32, blocks=r['res_blocks']):
            wd32.copy_(torch.cat([rb.Dense_0.weight.detach().to(self.dev, torch.float32) for rb in blocks], dim=0))
        self.pack_jobs.append(job)
        with torch.no_grad():
            job()
        sg = lambda *a: self._call('indm_sgemm_f32', *a)
        off = 0
        for rb in r['res_blocks']:
            co = rb.out_ch
            dsl = self.d_dense_tab[:, off:]
            if rb.Dense_0.weight.requires_grad:
                # dW[out][k] += sum_n d_tab[n][out] * SiLU(temb)[n][k]
                sg(1, 0, co, nf4, N, ctypes.c_float(1.0), dsl, ctypes.c_int64(tot), tact, ctypes.c_int64(nf4), ctypes.c_float(1.0),
                   self._pgrad(rb.Dense_0.weight), ctypes.c_int64(nf4))
                self._call('indm_colsum', dsl, L.DTYPE_F32, ctypes.c_int64(1), ctypes.c_int64(N), co, ctypes.c_int64(tot), None, ctypes.c_int64(0),
                           self._pgrad(rb.Dense_0.bias), ctypes.c_float(1.0))
            off += co
        d_tact, d_pre1, d_t0, d_pre0 = f32(N, nf4), f32(N, nf4), f32(N, nf4), f32(N, nf4)
        sg(0, 0, N, nf4, tot, ctypes.c_float(1.0), self.d_dense_tab, ctypes.c_int64(tot), wd32, ctypes.c_int64(nf4), ctypes.c_float(0.0),
           d_tact, ctypes.c_int64(nf4))
        self._call('indm_silu_bwd_f32', d_tact, pre1, d_pre1, ctypes.c_int64(N * nf4))
        lin0, lin1 = r['lin0'], r['lin1']
        sg(1, 0, nf4, nf4, N, ctypes.c_float(1.0), d_pre1, ctypes.c_int64(nf4), r['t0'], ctypes.c_int64(nf4), ctypes.c_float(1.0),
           self._pgrad(lin1.weight), ctypes.c_int64(nf4))
        self._call('indm_colsum', d_pre1, L.DTYPE_F32, ctypes.c_int64(1), ctypes.c_int64(N), nf4, ctypes.c_int64(nf4), None, ctypes.c_int64(0),
                   self._pgrad(lin1.bias), ctypes.c_float(1.0))
        sg(0, 0, N, nf4, nf4, ctypes.c_float(1.0), d_pre1, ctypes.c_int64(nf4), r['w1'], ctypes.c_int64(nf4), ctypes.c_float(0.0),
           d_t0, ctypes.c_int64(nf4))
        self._call('indm_silu_bwd_f32', d_t0, pre0, d_pre0, ctypes.c_int64(N * nf4))
        ed = r['emb_dim']
        sg(1, 0, nf4, ed, N, ctypes.c_float(1.0), d_pre0, ctypes.c_int64(nf4), self.emb, ctypes.c_int64(ed), ctypes.c_float(1.0),
           self._pgrad(lin0.weight), ctypes.c_int64(ed))
        self._call('indm_colsum', d_pre0, L.DTYPE_F32, ctypes.c_int64(1), ctypes.c_int64(N), nf4, ctypes.c_int64(nf4), None, ctypes.c_int64(0),
                   self._pgrad(lin0.bias), ctypes.c_float(1.0))

    def build_backward(self, train=False):
        """emit the backward plan: input-VJP only (train=False: likelihood / Hutchinson) or input-VJP + every parameter
        gradient (train=True: losses.get_step_fn)"""
        if self.infer:
            raise RuntimeError('ScoreEngine(pp=True) is an inference-only plan (padded-pixel operands / attention probabilities are not kept for the backward): '
                               'use NCSNpp.engine(batch) for vector-Jacobian products')
        if self._plans.get(train) is not None:
            self.bops = self._plans[train]['ops']
            self.gout, self.gx = self._plans[train]['gout'], self._plans[train]['gx']
            return
        self._train = bool(train)
        self._grads, self._gwritten, self._gnb_slots, self._pgrad_ptrs, self._scratch_zero = {}, set(), 0, [], []
        self._pyr_gx = None
        self.gnb_part_all = self._alloc((2 * len(self.tape) + 4, self.N, 32, 2), zero=True)
        if train:
            self.d_dense_tab = self._alloc((self.N, self.dense_total), zero=True)
        self._cur = bops = []
        part, scratch, ddt = self.gnb_part_all, self._scratch_zero, (self.d_dense_tab if train else None)

        def zero_bwd_state():
            part.zero_()
            if ddt is not None:
                ddt.zero_()
            for t in scratch:
                t.zero_()
        bops.append(zero_bwd_state)
        try:
            for kind, rec in reversed(self.tape):
                getattr(self, '_bwd_' + kind)(rec)
        finally:
            self._cur = self.ops
        self._plans[train] = dict(ops=bops, gout=self.gout, gx=self.gx, pgrads=self._pgrad_ptrs)
        self.bops = bops
        self._weights_version = None      # new weight packs were registered: repack on next use

    def _accumulate_into(self, target, src, scale):
        """grad(target) (+)= scale * src  (fp32 tensors of target's shape)"""
        g, acc = self._grad_of(target)
        if acc:
            self._call('indm_axpy_f32', g, src, ctypes.c_float(scale), ctypes.c_int64(g.numel()))
        else:
            self._call('indm_cast_scale', src, g, ctypes.c_int64(g.numel()), ctypes.c_float(scale), L.DTYPE_F32)

    def _bwd_pyramid(self, r):
        """backward of the progressive_input='residual' combine (models/ncsnpp.py:319-326): out = (conv_s2(FIR(pyr)) + b + h) * sc"""
        N, Hi, Wi, Cout, pc = self.N, r['Hi'], r['Wi'], r['Cout'], r['pyr_c']
        Ho, Wo = Hi // 2, Wi // 2
        ds, sc = r['ds'], r['sc']
        op_dt = L.DTYPE_BF16 if self.mode == 'bf16' else L.DTYPE_F32
        gout = self._grads[r['out'].data_ptr()]
        self._accumulate_into(r['h'], gout, sc)                       # the res-block branch
        g = self._cast_grad(r['out'], sc)
        cin = ds.Conv2d_0.weight.shape[1]
        if self._train and ds.Conv2d_0.weight.requires_grad:
            self._call('indm_conv_s2_wgrad', g, r['fir_out'], self._pgrad(ds.Conv2d_0.weight), op_dt, ctypes.c_int64(N), Ho, Wo, Cout, cin, pc)
        self._bgrad(g, Ho * Wo, Cout, Cout, biases=[ds.Conv2d_0.bias])
        w32 = self._alloc((Cout, cin, 3, 3))
        self._pack_f32(w32, [ds.Conv2d_0.weight])
        with torch.no_grad():
            self.pack_jobs[-1]()
        d_fir = self._alloc((N, Hi + 1, Wi + 1, pc))
        self._call('indm_conv_s2_dgrad', g, w32, d_fir, op_dt, ctypes.c_int64(N), Ho, Wo, Cout, cin, pc)
        d_src = self._alloc((N, Hi, Wi, pc))
        self._call('indm_fir_nhwc', d_fir, d_src, L.DTYPE_F32, L.DTYPE_F32, ctypes.c_int64(N), Hi + 1, Wi + 1, pc,
                   r['k1'].ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 4)
        if r['first']:
            self._pyr_gx = d_src          # gradient w.r.t. the NHWC-padded network input copy: folded in by _bwd_stem
        else:
            self._accumulate_into(r['pyr'], d_src, 1.0)

    def vjp(self, v, train=False):
        """v^T d(out)/d(x_in) for the activations of the last forward()/launch(): [N,C,S,S] fp32 in and out.
        (`out` includes the per-sample output scale that forward() was given.)  train=True additionally ACCUMULATES the
        gradient of <v, out> w.r.t. every parameter into `param.grad` (what `.backward()` does in losses.py:250)."""
        self.build_backward(train)
        if train:
            for prm, ptr in self._plans[True]['pgrads']:
                if prm.grad is None or prm.grad.data_ptr() != ptr:
                    raise RuntimeError('indm_b200: a parameter .grad tensor was replaced after the training plan was built '
                                       '(use optimizer.zero_grad(set_to_none=False) / the fused optimizer of indm_b200.losses)')
        if self._weights_version != self.weights_version():
            self.load_weights()
        self.gout.copy_(v)
        g = self._bwd_graphs.get(train)
        if g is not None and g[1] == self._bwd_graph_key(train):
            g[0].replay()             # captured ahead of time on the main thread (precapture_backward); autograd only replays
        else:
            for op in self.bops:
                op()
            self._bwd_ran.add(train)
        return self.gx

    def _bwd_graph_key(self, train):
        return tuple(ptr for _, ptr in self._plans[train]['pgrads']) if train else ()

    def precapture_backward(self, train=False):
        """Called on the main thread right after a forward whose backward will be asked for (models/ncsnpp.py:_EngineFunction): the
        backward plan (~800 launches) becomes one CUDA graph once it has run eagerly (that first run builds weight packs and lazily
        sized buffers).  torch.autograd invokes `vjp` from its worker thread, where a capture would be invalidated and ~800 eager
        launches per call bound the step on the host; a replay has neither problem."""
        import threading
        if train not in self._bwd_ran or threading.current_thread() is not threading.main_thread() or torch.cuda.is_current_stream_capturing():
            return
        key = self._bwd_graph_key(train)
        g = self._bwd_graphs.get(train)
        if g is not None and g[1] == key:
            return
        if self._weights_version != self.weights_version():
            self.load_weights()
        self.build_backward(train)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for op in self.bops:
                op()
        self._bwd_graphs[train] = (graph, key)

    # ------------------------------------------------------------------ execution
    def launch(self, temb_op=None):
        """enqueue the whole forward on the current stream (inputs: self.x_in, self.time_cond, self.out_scale); `temb_op` (from
        `_bind_time_source(sched, ...)`) replaces the time-embedding launch for this call only (sampler graphs)"""
        for i, op in enumerate(self.ops):
            if temb_op is not None and i == self._temb_op_index:
                temb_op()
            else:
                op()

    def forward(self, x, time_cond, out_scale=None, train=False, seed=None):
        """train=True enables the dropout masks (a fresh Philox seed per call unless `seed` is given)"""
        if train:
            self._drop_seed = (self._drop_seed * 6364136223846793005 + 1442695040888963407) % (1 << 63) if seed is None else int(seed)
            self.drop_ctl.copy_(torch.tensor([self._drop_seed, 1], dtype=torch.int64))
        elif self._drop_on:
            self.drop_ctl.zero_()
        self._drop_on = bool(train)
        if x.shape[0] != self.N:
            raise ValueError(f'engine was built for batch {self.N}, got {x.shape[0]}')
        if self._weights_version != self.weights_version():
            self.load_weights()
        self.x_in.copy_(x)
        self.time_cond.copy_(time_cond)
        if out_scale is None:
            self.out_scale.fill_(1.0)
        else:
            self.out_scale.copy_(out_scale)
        self.forward_count += 1
        self.launch()
        return self.out

    @property
    def num_launches(self):
        return len(self.ops)
