"""GPU: every C-ABI kernel against a CPU FP32/FP64 computation of the same op on the same seeded inputs.

Tolerances: BF16 mode — inputs are pre-rounded to bf16 so the only differences are fp32 accumulation order and the
output cast (2^-9 relative when the output is bf16); TF32 mode — inputs pre-rounded to tf32, fp32 accumulate.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from indm_b200 import _lib as L  # noqa: E402

DEV = 'cuda'


_KEEP = []


def D(t):
    """tensor -> device, kept alive until the test module is torn down (the C ABI sees raw pointers only)"""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    t = t.to(DEV)
    _KEEP.append(t)
    return t


def P(t):
    return L.ptr(D(t))


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def tf32r(t):
    # round-to-nearest-away on the 13 dropped mantissa bits == cvt.rna.tf32.f32
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def rel_l2(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev_op(t, dtype):
    """operand on device in the kernel's element type"""
    return (t.to(torch.bfloat16) if dtype == L.DTYPE_BF16 else t.float()).contiguous().to(DEV)


def pack_w(w, dtype):
    """[Cout, Cin, kh, kw] -> [taps][Cout][Cin]"""
    co, ci, kh, kw = w.shape
    return dev_op(w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci), dtype)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def round_in(t, dtype):
    return bf16r(t) if dtype == L.DTYPE_BF16 else tf32r(t)


@pytest.mark.parametrize("dtype", [L.DTYPE_BF16, L.DTYPE_TF32])
@pytest.mark.parametrize("M,K,Nn,bn", [(300, 192, 200, 0), (128, 64, 32, 32), (1000, 512, 768, 256), (257, 128, 128, 64)])
def test_igemm_plain_gemm(dtype, M, K, Nn, bn):
    a = round_in(rnd(M, K, seed=1), dtype)
    b = round_in(rnd(Nn, K, seed=2) / math.sqrt(K), dtype)
    bias = rnd(Nn, seed=3)
    out = torch.full((M, Nn), float('nan'), device=DEV)
    L.igemm(dtype=dtype, a=dev_op(a, dtype), N=1, H=1, W=M, Cin=K, b=dev_op(b, dtype), Cout=Nn, taps=1,
            bias=bias.to(DEV), out_f32=out, out_ld=Nn, block_n=bn)
    torch.cuda.synchronize()
    want = a.double() @ b.double().t() + bias.double()
    assert rel_l2(out.cpu(), want) < 2e-5


@pytest.mark.parametrize("dtype", [L.DTYPE_BF16, L.DTYPE_TF32])
@pytest.mark.parametrize("N,S,Cin,Cout,bn", [(2, 16, 128, 256, 0), (1, 32, 64, 128, 128), (3, 8, 128, 128, 64),
                                             (3, 4, 256, 256, 256), (9, 4, 64, 64, 0), (1, 64, 64, 64, 0)])
def test_igemm_conv3x3(dtype, N, S, Cin, Cout, bn):
    x = round_in(rnd(N, Cin, S, S, seed=4), dtype)
    w = round_in(rnd(Cout, Cin, 3, 3, seed=5) / math.sqrt(9 * Cin), dtype)
    bias = rnd(Cout, seed=6)
    rowb = rnd(N, Cout + 8, seed=7)
    res = rnd(N, S, S, Cout, seed=8)
    o32 = torch.full((N, S, S, Cout), float('nan'), device=DEV)
    o16 = torch.zeros((N, S, S, Cout), device=DEV, dtype=torch.bfloat16)
    L.igemm(dtype=dtype, a=dev_op(nhwc(x), dtype), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dtype), Cout=Cout, taps=9,
            bias=bias.to(DEV), rowbias=rowb.to(DEV), rowbias_ld=Cout + 8, residual=res.to(DEV), res_ld=Cout,
            scale=0.7071, res_scale=0.7071, out_f32=o32, out_bf16=o16, out_ld=Cout, block_n=bn)
    torch.cuda.synchronize()
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    want = (nhwc(want) + rowb[:, None, None, :Cout].double() + res.double()) * 0.7071
    assert rel_l2(o32.cpu(), want) < 2e-5
    assert rel_l2(o16.float().cpu(), want) < 4e-3


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("N,S,Cin,Cin2,Cout", [(20, 32, 128, 0, 128), (38, 16, 256, 0, 256), (302, 8, 128, 0, 128), (80, 16, 128, 256, 256),
                                               (75, 16, 64, 0, 512)])
def test_igemm_cta_pair_path(kind, N, S, Cin, Cin2, Cout):
    """Launches that fill the chip (>= 148 output tiles) with the plain BF16 epilogues take the cta_group::2 path: pairs of M tiles
    issued as one M = 256 MMA, B split across the pair.  Covers 128- and 256-wide tiles, two N tiles, an odd number of M tiles
    (302 images at 8x8: 151 tiles, the last pair has one all-out-of-bounds tile), the fused second K segment, fused GroupNorm
    statistics, against fp64 convolution."""
    dtype = L.DTYPE_BF16
    x = round_in(rnd(N, Cin, S, S, seed=40), dtype)
    w = round_in(rnd(Cout, Cin, 3, 3, seed=41) / math.sqrt(9 * Cin), dtype)
    bias = rnd(Cout, seed=42)
    kw = dict(dtype=dtype, a=dev_op(nhwc(x), dtype), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dtype), Cout=Cout, taps=9, bias=bias.to(DEV),
              scale=0.7071, out_ld=Cout)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    if Cin2:
        xs = round_in(rnd(N, Cin2, S, S, seed=43), dtype)
        w2 = round_in(rnd(Cout, Cin2, 1, 1, seed=44) / math.sqrt(Cin2), dtype)
        kw.update(a2=dev_op(nhwc(xs), dtype), Cin2=Cin2, b2=dev_op(w2.reshape(Cout, Cin2), dtype))
        want = want + F.conv2d(xs.double(), w2.double())
    want = nhwc(want)
    part = torch.zeros((N, 32, 2), device=DEV)
    if kind == 1:
        rowb = rnd(N, Cout, seed=45)
        out = torch.zeros((N, S, S, Cout), device=DEV, dtype=torch.bfloat16)
        kw.update(rowbias=rowb.to(DEV), rowbias_ld=Cout, out_bf16=out, gn_partial=part, gn_cpg=Cout // 32, gn_groups=32)
        want = (want + rowb[:, None, None, :].double()) * 0.7071
        tol = 4e-3
    else:
        res = rnd(N, S, S, Cout, seed=46)
        out = torch.full((N, S, S, Cout), float('nan'), device=DEV)
        kw.update(residual=res.to(DEV), res_ld=Cout, res_scale=0.7071, out_f32=out, gn_partial=part, gn_cpg=Cout // 32, gn_groups=32)
        want = want * 0.7071 + res.double() * 0.7071
        tol = 2e-5
    L.igemm(**kw)
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), want) < tol
    g = want.reshape(N, S * S, 32, Cout // 32)
    ws, wq = g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))
    assert rel_l2(part[..., 0].cpu(), ws) < 2e-3 and rel_l2(part[..., 1].cpu(), wq) < 2e-3


@pytest.mark.parametrize("dtype", [L.DTYPE_BF16, L.DTYPE_TF32])
def test_igemm_fused_skip_segment(dtype):
    N, S, C1, C2, Cout = 2, 16, 128, 384, 256
    h = round_in(rnd(N, C1, S, S, seed=9), dtype)
    xs = round_in(rnd(N, C2, S, S, seed=10), dtype)
    w1 = round_in(rnd(Cout, C1, 3, 3, seed=11) / math.sqrt(9 * C1), dtype)
    w2 = round_in(rnd(Cout, C2, 1, 1, seed=12) / math.sqrt(C2), dtype)
    o32 = torch.empty((N, S, S, Cout), device=DEV)
    L.igemm(dtype=dtype, a=dev_op(nhwc(h), dtype), N=N, H=S, W=S, Cin=C1, b=pack_w(w1, dtype), Cout=Cout, taps=9,
            a2=dev_op(nhwc(xs), dtype), Cin2=C2, b2=dev_op(w2.reshape(Cout, C2), dtype), out_f32=o32, out_ld=Cout)
    torch.cuda.synchronize()
    want = nhwc(F.conv2d(h.double(), w1.double(), padding=1) + F.conv2d(xs.double(), w2.double()))
    assert rel_l2(o32.cpu(), want) < 2e-5


@pytest.mark.parametrize("L_", [256, 64, 16])
def test_igemm_batched_attention_products(L_):
    """S = Q K^T / sqrt(C) and O = P V with per-image B operands, exactly the two einsums of AttnBlockpp."""
    B, Cc = 3, 256
    dt = L.DTYPE_BF16
    qkv = bf16r(rnd(B, L_, 3 * Cc, seed=13))
    qkv_d = dev_op(qkv, dt)
    s = torch.empty((B, L_, L_), device=DEV)
    L.igemm(dtype=dt, a=qkv_d, a_ld=3 * Cc, a_img_stride=L_ * 3 * Cc, N=B, H=1, W=L_, Cin=Cc,
            b=qkv_d[:, :, Cc:], b_ld=3 * Cc, b_tap_stride=L_ * 3 * Cc, Cout=L_, taps=1, batched_b=1,
            scale=Cc ** -0.5, out_f32=s, out_ld=L_)
    torch.cuda.synchronize()
    q, k, v = qkv[..., :Cc].double(), qkv[..., Cc:2 * Cc].double(), qkv[..., 2 * Cc:].double()
    want_s = torch.einsum('bic,bjc->bij', q, k) * Cc ** -0.5
    assert rel_l2(s.cpu(), want_s) < 2e-5
    p = bf16r(torch.softmax(want_s.float(), dim=-1))
    vt = dev_op(qkv[..., 2 * Cc:].transpose(1, 2).contiguous(), dt)     # [B, C, L]
    o = torch.empty((B, L_, Cc), device=DEV, dtype=torch.bfloat16)
    L.igemm(dtype=dt, a=dev_op(p, dt), N=B, H=1, W=L_, Cin=L_, b=vt, Cout=Cc, taps=1, batched_b=1,
            out_bf16=o, out_ld=Cc)
    torch.cuda.synchronize()
    want_o = torch.einsum('bij,bjc->bic', p.double(), v)
    assert rel_l2(o.float().cpu(), want_o) < 4e-3


def test_igemm_head_nchw_and_transposed_modes():
    dt = L.DTYPE_BF16
    N, S, Cin = 2, 32, 128
    x = bf16r(rnd(N, Cin, S, S, seed=14))
    w = bf16r(rnd(3, Cin, 3, 3, seed=15) / math.sqrt(9 * Cin))
    bias = rnd(3, seed=16)
    rs = torch.tensor([0.5, -2.0])
    out = torch.full((N, 3, S, S), float('nan'), device=DEV)
    L.igemm(dtype=dt, a=dev_op(nhwc(x), dt), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dt), Cout=3, taps=9,
            bias=bias.to(DEV), rowscale=rs.to(DEV), out_mode=1, out_f32=out)
    torch.cuda.synchronize()
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1) * rs.double()[:, None, None, None]
    assert rel_l2(out.cpu(), want) < 2e-5
    # mode 2: q,k rows + V transposed
    B, Lq, Cc = 2, 256, 256
    hh = bf16r(rnd(B, Lq, Cc, seed=17))
    wq = bf16r(rnd(3 * Cc, Cc, seed=18) / math.sqrt(Cc))
    bq = rnd(3 * Cc, seed=19)
    qk = torch.zeros((B, Lq, 2 * Cc), device=DEV, dtype=torch.bfloat16)
    vt = torch.zeros((B, Cc, Lq), device=DEV, dtype=torch.bfloat16)
    L.igemm(dtype=dt, a=dev_op(hh, dt), N=B, H=16, W=16, Cin=Cc, b=dev_op(wq, dt), Cout=3 * Cc, taps=1, bias=bq.to(DEV),
            out_mode=2, out_bf16=qk, out_ld=2 * Cc, tcol0=2 * Cc, out_t=vt)
    torch.cuda.synchronize()
    want = hh.double() @ wq.double().t() + bq.double()
    assert rel_l2(qk.float().cpu(), want[..., :2 * Cc]) < 4e-3
    assert rel_l2(vt.float().cpu(), want[..., 2 * Cc:].transpose(1, 2)) < 4e-3


@pytest.mark.parametrize("S,N", [(16, 2), (8, 3), (32, 1)])
def test_igemm_fused_groupnorm_statistics(S, N):
    dt = L.DTYPE_BF16
    Cin, Cout, G = 128, 256, 32
    x = bf16r(rnd(N, Cin, S, S, seed=20))
    w = bf16r(rnd(Cout, Cin, 3, 3, seed=21) / math.sqrt(9 * Cin))
    o16 = torch.zeros((N, S, S, Cout), device=DEV, dtype=torch.bfloat16)
    part = torch.zeros((N, G, 2), device=DEV)
    L.igemm(dtype=dt, a=dev_op(nhwc(x), dt), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dt), Cout=Cout, taps=9,
            out_bf16=o16, out_ld=Cout, gn_partial=part, gn_cpg=Cout // G, gn_groups=G)
    torch.cuda.synchronize()
    # the statistics are those of the fp32 accumulator values (the stored copy is their BF16 rounding: zero-mean noise)
    y = nhwc(F.conv2d(x.double(), w.double(), padding=1)).reshape(N, S * S, G, Cout // G)
    want = torch.stack([y.sum(dim=(1, 3)), (y * y).sum(dim=(1, 3))], dim=-1)
    assert rel_l2(part.cpu(), want) < 1e-5
    # and they describe the stored tensor to BF16-rounding accuracy
    y16 = o16.float().cpu().reshape(N, S * S, G, Cout // G).double()
    want16 = torch.stack([y16.sum(dim=(1, 3)), (y16 * y16).sum(dim=(1, 3))], dim=-1)
    assert rel_l2(part.cpu(), want16) < 1e-3


@pytest.mark.parametrize("in_dt,out_dt", [(L.DTYPE_F32, L.DTYPE_BF16), (L.DTYPE_BF16, L.DTYPE_BF16), (L.DTYPE_F32, L.DTYPE_TF32)])
@pytest.mark.parametrize("Ca,Cb,S,res,act", [(128, 0, 16, 0, 1), (256, 128, 8, 0, 1), (256, 256, 4, 0, 1), (128, 0, 16, 2, 1),
                                              (256, 0, 8, 1, 1), (256, 0, 16, 0, 0), (512, 0, 4, 0, 1), (128, 0, 32, 0, 1)])
def test_groupnorm_silu_resample(in_dt, out_dt, Ca, Cb, S, res, act):
    N = 3
    C = Ca + Cb
    G = min(C // 4, 32)
    xa = rnd(N, S, S, Ca, seed=22) * 1.5 + 0.3
    xb = rnd(N, S, S, Cb, seed=23) if Cb else None
    if in_dt == L.DTYPE_BF16:
        xa = bf16r(xa); xb = bf16r(xb) if Cb else None
    gamma, beta = 1 + 0.1 * rnd(C, seed=24), 0.1 * rnd(C, seed=25)
    tin = torch.bfloat16 if in_dt == L.DTYPE_BF16 else torch.float32
    xa_d = xa.to(tin).to(DEV)
    xb_d = xb.to(tin).to(DEV) if Cb else None
    part = torch.zeros((N, G, 2), device=DEV)
    L.call('indm_gn_stats', L.ptr(xa_d), Ca, L.ptr(xb_d), Cb, in_dt, N, S * S, G, L.ptr(part))
    So = S * 2 if res == 1 else (S // 2 if res == 2 else S)
    tout = torch.bfloat16 if out_dt == L.DTYPE_BF16 else torch.float32
    out = torch.zeros((N, So, So, C), device=DEV, dtype=tout)
    raw = torch.zeros((N, So, So, C), device=DEV, dtype=tout)
    L.call('indm_gn_apply', L.ptr(xa_d), Ca, L.ptr(xb_d), Cb, in_dt, N, S, S, G, L.ptr(part), P(gamma),
           P(beta), 1e-6, act, res, L.ptr(out), L.ptr(raw), out_dt)
    torch.cuda.synchronize()
    x = torch.cat([xa, xb], dim=-1) if Cb else xa
    xc = x.permute(0, 3, 1, 2).double()
    y = F.group_norm(xc, G, gamma.double(), beta.double(), eps=1e-6)
    if act:
        y = F.silu(y)

    def resample(t):
        if res == 1:
            return t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        if res == 2:
            return F.avg_pool2d(t, 2)
        return t
    tol = 4e-3 if out_dt == L.DTYPE_BF16 else 5e-4
    assert rel_l2(out.float().cpu(), nhwc(resample(y))) < tol
    assert rel_l2(raw.float().cpu(), nhwc(resample(xc))) < tol


@pytest.mark.parametrize("cols", [16, 64, 256])
def test_softmax_rows(cols):
    s = rnd(37, cols, seed=26) * 3
    out = torch.zeros((37, cols), device=DEV, dtype=torch.bfloat16)
    L.call('indm_softmax_rows', P(s), L.ptr(out), 37, cols, L.DTYPE_BF16)
    out32 = torch.zeros((37, cols), device=DEV)
    L.call('indm_softmax_rows', P(s), L.ptr(out32), 37, cols, L.DTYPE_F32)
    torch.cuda.synchronize()
    want = torch.softmax(s.double(), dim=-1)
    assert rel_l2(out.float().cpu(), want) < 4e-3
    assert rel_l2(out32.cpu(), want) < 1e-5


def test_prep_input_and_time_embedding_and_linear():
    x = rnd(3, 3, 8, 8, seed=27)
    out = torch.full((3, 8, 8, 64), 7.0, device=DEV, dtype=torch.bfloat16)
    L.call('indm_prep_input', P(x), L.ptr(out), 3, 3, 8, 8, 64, 2.0, -1.0, 0, L.DTYPE_BF16)
    torch.cuda.synchronize()
    want = torch.zeros(3, 8, 8, 64)
    want[..., :3] = nhwc(2 * x - 1)
    assert rel_l2(out.float().cpu(), want) < 4e-3
    assert float(out[..., 3:].abs().max()) == 0.0
    # positional embedding, models/layers.py:515-529
    t = torch.tensor([0.0, 1.7, 333.3, 999.0])
    emb = torch.zeros((4, 128), device=DEV)
    L.call('indm_time_embedding', P(t), None, None, 0, 0, None, 0, 4, 128, L.ptr(emb))
    half = 64
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * e[None, :]
    want = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    torch.cuda.synchronize()
    assert float((emb.cpu() - want).abs().max()) < 2e-4     # fp32 sin/cos at arguments up to ~1e3
    # Gaussian Fourier, models/layerspp.py:52-54
    sig = torch.tensor([0.01, 0.5, 3.0, 50.0])
    Wf = rnd(64, seed=28) * 16
    emb = torch.zeros((4, 128), device=DEV)
    L.call('indm_time_embedding', P(sig), None, None, 0, 0, P(Wf), 1, 4, 128, L.ptr(emb))
    xp = torch.log(sig)[:, None] * Wf[None, :] * 2 * np.pi
    want = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
    torch.cuda.synchronize()
    assert float((emb.cpu() - want).abs().max()) < 5e-4
    # dense
    inp, w, b = rnd(5, 512, seed=29), rnd(300, 512, seed=30) / 22, rnd(300, seed=31)
    o = torch.zeros((5, 300), device=DEV)
    L.call('indm_linear_f32', P(inp), P(w), P(b), L.ptr(o), 5, 512, 300, 1, 0, L.DTYPE_F32)
    torch.cuda.synchronize()
    want = F.linear(F.silu(inp.double()), w.double(), b.double())
    assert rel_l2(o.cpu(), want) < 1e-5
    o16 = torch.zeros((5, 300), device=DEV, dtype=torch.bfloat16)
    L.call('indm_linear_f32', P(inp), P(w), P(b), L.ptr(o16), 5, 512, 300, 0, 1, L.DTYPE_BF16)
    torch.cuda.synchronize()
    assert rel_l2(o16.float().cpu(), F.silu(F.linear(inp.double(), w.double(), b.double()))) < 4e-3


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_fir_nhwc_matches_oracle_upfirdn2d(mode):
    from oracle import ops as oops
    N, S, Cc = 2, 8, 16
    x = rnd(N, Cc, S, S, seed=32)
    k1 = np.array([1, 3, 3, 1], dtype=np.float32)
    k1 = k1 / k1.sum()
    k2 = np.outer(k1, k1)
    args = {1: dict(up=2, down=1, pad=(2, 1)), 2: dict(up=1, down=2, pad=(1, 1)), 3: dict(up=1, down=1, pad=(2, 2))}[mode]
    want = oops.upfirdn2d(x.numpy(), k2 * (4 if mode == 1 else 1), **args)
    kk = (k1 * (2 if mode == 1 else 1)).astype(np.float32)
    So = want.shape[-1]
    out = torch.zeros((N, So, So, Cc), device=DEV)
    import ctypes
    L.call('indm_fir_nhwc', P(nhwc(x)), L.ptr(out), L.DTYPE_F32, L.DTYPE_F32, N, S, S, Cc,
           kk.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), mode)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), nhwc(torch.from_numpy(want))) < 1e-6


def test_pc_update_kernels_match_oracle():
    from oracle import sampler as osampler, sde as osde
    N, D = 4, 3 * 16 * 16
    x, s, z = rnd(N, 3, 16, 16, seed=33), rnd(N, 3, 16, 16, seed=34), rnd(N, 3, 16, 16, seed=35)
    sde = osde.VP()
    t = torch.full((N,), 0.37)
    want_x, want_mean = osampler.reverse_diffusion_update(sde, s, x, t, z)
    ts = int(0.37 * 999)
    beta, alpha = float(sde.discrete_betas[ts]), float(sde.alphas[ts])
    coef = torch.tensor([[2 - math.sqrt(alpha), beta, math.sqrt(beta), 0.0]])
    xd, xm = x.clone().to(DEV), torch.zeros_like(x).to(DEV)
    L.call('indm_pc_predictor_update', L.ptr(xd), P(s), P(z), L.ptr(xm), P(coef), 4, None,
           N, D, 0, None, 0)
    torch.cuda.synchronize()
    assert rel_l2(xd.cpu(), want_x) < 1e-6 and rel_l2(xm.cpu(), want_mean) < 1e-6
    # Langevin, VE (alpha = 1)
    ve = osde.VE()
    want_x, want_mean = osampler.langevin_update(ve, s, x, t, z, 0.16)
    norms = torch.zeros((N, 2), device=DEV)
    L.call('indm_langevin_norms', P(s), P(z), L.ptr(norms), None, N, D, 0, None, 0)
    coef = torch.tensor([[1.0, 0.16]])
    xd, xm = x.clone().to(DEV), torch.zeros_like(x).to(DEV)
    L.call('indm_langevin_update', L.ptr(xd), P(s), P(z), L.ptr(xm), L.ptr(norms), P(coef), 2,
           None, N, D, 0, None, 0)
    torch.cuda.synchronize()
    assert rel_l2(xd.cpu(), want_x) < 1e-5 and rel_l2(xm.cpu(), want_mean) < 1e-5


def test_philox_normal_stream_statistics_and_replay():
    n = 1 << 20
    a, b = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    L.call('indm_randn_f32', L.ptr(a), n, 1234, 5)
    L.call('indm_randn_f32', L.ptr(b), n, 1234, 5)
    c = torch.zeros(n, device=DEV)
    L.call('indm_randn_f32', L.ptr(c), n, 1234, 6)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3) < 0.05
    assert abs(float((a * c).mean())) < 5e-3
    # in-kernel noise: langevin_norms(z=None) sees the same draw as langevin_update(z=None)
    N, D = 2, 4096
    s = rnd(N, D, seed=36).to(DEV)
    norms = torch.zeros((N, 2), device=DEV)
    L.call('indm_langevin_norms', L.ptr(s), None, L.ptr(norms), None, N, D, 77, None, 1)
    x0 = torch.zeros((N, D), device=DEV)
    coef = torch.tensor([[1.0, 0.16]], device=DEV)
    L.call('indm_langevin_update', L.ptr(x0), P(torch.zeros_like(s)), None, None, L.ptr(norms), L.ptr(coef), 2, None, N, D, 77, None, 1)
    torch.cuda.synchronize()
    # with s = 0 in the update, x = sqrt(2 eps) z  =>  |x_n|^2 / (2 eps) == |z_n|^2 from the norms kernel
    r = 0.16 * norms[:, 1].sqrt().mean() / norms[:, 0].sqrt().mean()
    eps = float(2 * r * r)
    got = (x0 ** 2).sum(dim=1) / (2 * eps)
    assert rel_l2(got.cpu(), norms[:, 1].cpu()) < 1e-4


UPFIRDN_CASES = ['up2', 'down2', 'pyr', 'k3', 'crop', 'gen', 'up2_32']


@pytest.mark.parametrize("case", UPFIRDN_CASES)
def test_upfirdn2d_c_abi_against_reference_golden(case):
    from helpers import load_npz
    g = load_npz('ops.npz')
    up, down, p0, p1 = [int(v) for v in g[f'upfirdn_{case}_args']]
    x, k, want = g[f'upfirdn_{case}_x'], g[f'upfirdn_{case}_k'], g[f'upfirdn_{case}_y']
    n, c, h, w = x.shape
    y = torch.full(want.shape, float('nan'), device=DEV)
    L.call('indm_upfirdn2d_f32', P(x), P(k), L.ptr(y), n * c, h, w,
           k.shape[0], k.shape[1], up, up, down, down, p0, p1, p0, p1)
    torch.cuda.synchronize()
    # north_star: upfirdn2d within 1e-6 relative in FP32
    assert rel_l2(y.cpu(), torch.from_numpy(want)) < 1e-6


@pytest.mark.parametrize("case", ['4d', '2d', '3d'])
def test_bias_act_c_abi_against_reference_golden(case):
    from helpers import load_npz
    g = load_npz('ops.npz')
    x, b, want = g[f'lrelu_{case}_x'], g[f'lrelu_{case}_b'], g[f'lrelu_{case}_y']
    step = int(np.prod(x.shape[2:]))
    y = torch.zeros(x.shape, device=DEV)
    L.call('indm_bias_act_f32', P(x), P(b), None, L.ptr(y), x.size,
           b.shape[0], step, 3, 0, 0.2, 2 ** 0.5)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu(), torch.from_numpy(want)) < 1e-6
    # backward: grad=1 with ref = forward output
    gy, want_gx = g[f'lrelu_{case}_gy'], g[f'lrelu_{case}_gx']
    gx = torch.zeros(x.shape, device=DEV)
    L.call('indm_bias_act_f32', P(gy), None, L.ptr(y), L.ptr(gx), x.size, 1, 1, 3, 1, 0.2, 2 ** 0.5)
    torch.cuda.synchronize()
    assert rel_l2(gx.cpu(), torch.from_numpy(want_gx)) < 1e-6


@pytest.mark.parametrize("mode", ['bf16', 'tf32'])
@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 256, 9), (8, 4, 4, 256, 128, 9), (2, 32, 32, 64, 192, 9), (2, 16, 16, 128, 320, 1),
                                   (1, 1, 256, 256, 64, 1)])
def test_conv_wgrad_matches_autograd(shape, mode):
    """indm_conv_wgrad (tcgen05 with MN-major operands read straight from the NHWC forward boxes) against the weight gradient
    autograd computes for F.conv2d (3x3 pad 1) / a 1x1 product, accumulated on top of a non-zero initial gradient."""
    N, H, W, Cout, Cin, taps = shape
    dt = L.DTYPE_BF16 if mode == 'bf16' else L.DTYPE_TF32
    rd = bf16r if mode == 'bf16' else tf32r
    x = rd(rnd(N, H, W, Cin, seed=1))
    dy = rd(rnd(N, H, W, Cout, seed=2))
    k = 3 if taps == 9 else 1
    w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=k // 2)
    want, = torch.autograd.grad(y, w, dy.permute(0, 3, 1, 2).double())
    init = rnd(Cout, Cin, k, k, seed=3)
    dw = D(init.clone())
    tdt = torch.bfloat16 if mode == 'bf16' else torch.float32
    L.call('indm_conv_wgrad', P(dy.to(tdt)), 0, P(x.to(tdt)), 0, dt, N, H, W, Cout, Cin, taps, L.ptr(dw), Cin * taps, taps, 1, 0.5)
    torch.cuda.synchronize()
    got = (dw.cpu() - init) / 0.5
    e = rel_l2(got, want.float())
    print(f'wgrad {shape} {mode}: rel-L2 {e:.3e}')
    assert e < (2e-3 if mode == 'tf32' else 1e-4)


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("N,S,Cin,Cin2,Cout", [(20, 32, 128, 0, 128), (19, 32, 64, 0, 128), (18, 32, 128, 256, 128), (18, 32, 256, 0, 256),
                                               (4, 64, 128, 0, 128)])
def test_igemm_halo_padded_pixel_path(kind, N, S, Cin, Cin2, Cout):
    """3x3 convolutions of 32- and 64-wide maps take the padded-pixel kernel (csrc/igemm_halo.cu): one activation box per K chunk,
    nine descriptor offsets, tiles of 128 padded pixels that start anywhere inside an image row, CTA pairs over two images.  Covers an
    odd image count (the last pair's second image is out of bounds), the fused 1x1 skip segment, 256-wide tiles, GroupNorm statistics
    of two consumers, against fp64 convolution; and bit-for-bit insensitivity to what the previous launch left in shared memory."""
    dtype = L.DTYPE_BF16
    x = round_in(rnd(N, Cin, S, S, seed=50), dtype)
    w = round_in(rnd(Cout, Cin, 3, 3, seed=51) / math.sqrt(9 * Cin), dtype)
    bias = rnd(Cout, seed=52)
    kw = dict(dtype=dtype, a=dev_op(nhwc(x), dtype), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dtype), Cout=Cout, taps=9, bias=bias.to(DEV),
              scale=0.7071, out_ld=Cout)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    if Cin2:
        xs = round_in(rnd(N, Cin2, S, S, seed=53), dtype)
        w2 = round_in(rnd(Cout, Cin2, 1, 1, seed=54) / math.sqrt(Cin2), dtype)
        kw.update(a2=dev_op(nhwc(xs), dtype), Cin2=Cin2, b2=dev_op(w2.reshape(Cout, Cin2), dtype))
        want = want + F.conv2d(xs.double(), w2.double())
    want = nhwc(want)
    part = torch.zeros((N, 32, 2), device=DEV)
    part2 = torch.zeros((N, 16, 2), device=DEV)
    gn = dict(gn_partial=part, gn_cpg=Cout // 32, gn_groups=32, gn2_partial=part2, gn2_cpg=Cout // 16, gn2_groups=16)
    if kind == 1:
        rowb = rnd(N, Cout, seed=55)
        out = torch.full((N, S, S, Cout), float('nan'), device=DEV, dtype=torch.bfloat16)
        kw.update(rowbias=rowb.to(DEV), rowbias_ld=Cout, out_bf16=out, **gn)
        want = (want + rowb[:, None, None, :].double()) * 0.7071
        tol = 4e-3
    else:
        res = rnd(N, S, S, Cout, seed=56)
        out = torch.full((N, S, S, Cout), float('nan'), device=DEV)
        kw.update(residual=res.to(DEV), res_ld=Cout, res_scale=0.7071, out_f32=out, **gn)
        want = want * 0.7071 + res.double() * 0.7071
        tol = 2e-5
    L.igemm(**kw)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float().cpu(), want) < tol
    for pt, G in ((part, 32), (part2, 16)):
        g = want.reshape(N, S * S, G, Cout // G)
        ws, wq = g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))
        assert rel_l2(pt[..., 0].cpu(), ws) < 2e-3 and rel_l2(pt[..., 1].cpu(), wq) < 2e-3
    first = out.clone()
    L.igemm(**kw)
    torch.cuda.synchronize()
    assert torch.equal(first, out)


@pytest.mark.parametrize("N", [1, 5, 128])
def test_fused_attention_forward_matches_softmax_attention(N):
    """indm_attention_fwd (scores in TMEM, probabilities in shared memory, V read in place as an MN-major operand) against
    softmax(q k^T C^-1/2) v in fp64 on the BF16-rounded q | k | v (models/layerspp.py:94-99), L = 256, C = 256; peaked and flat
    rows both occur (q, k scaled so the logits span ~[-8, 8])."""
    Lq, Cc = 256, 256
    qkv = (rnd(N, Lq, 3 * Cc, seed=90) * torch.tensor([1.5] * Cc + [1.5] * Cc + [1.0] * Cc)).to(torch.bfloat16)
    out = torch.full((N, Lq, Cc), float('nan'), device=DEV, dtype=torch.bfloat16)
    dq = qkv.to(DEV)
    L.call('indm_attention_fwd', L.ptr(dq), L.ptr(out), N, Lq, Cc, Cc ** -0.5, L.DTYPE_BF16)
    torch.cuda.synchronize()
    q, k, v = (t.double() for t in qkv.split(Cc, dim=2))
    want = torch.softmax(q @ k.transpose(1, 2) * Cc ** -0.5, dim=-1) @ v
    assert torch.isfinite(out.float()).all()
    e = rel_l2(out.float().cpu(), want)
    print(f'fused attention N={N}: rel-L2 {e:.2e}')
    assert e < 4e-3                                   # BF16 probabilities and output
    first = out.clone()
    L.call('indm_attention_fwd', L.ptr(dq), L.ptr(out), N, Lq, Cc, Cc ** -0.5, L.DTYPE_BF16)
    torch.cuda.synchronize()
    assert torch.equal(first, out)


def test_fused_attention_rejects_other_shapes():
    x = torch.zeros((2, 64, 3 * 128), device=DEV, dtype=torch.bfloat16)
    o = torch.zeros((2, 64, 128), device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        L.call('indm_attention_fwd', L.ptr(x), L.ptr(o), 2, 64, 128, 128 ** -0.5, L.DTYPE_BF16)


@pytest.mark.parametrize("N,S,C,Cb", [(3, 8, 128, 0), (2, 16, 256, 0), (5, 4, 128, 128)])
def test_gn_apply_dropout_mask_is_the_same_in_every_kernel_variant(N, S, C, Cb):
    """nn.Dropout inside ResnetBlockBigGANpp (models/layerspp.py:278): the mask is a pure function of (seed, stream, element quad), so
    the 8-channel BF16 kernel, the TF32-output kernel (whose indexing the backward kernels share) and the padded-pixel variant must
    drop exactly the same elements and scale the kept ones by 1 / (1 - p); enabled = 0 in the control buffer switches it off."""
    p_drop, seed, stream_id = 0.1, 0x1234, 7
    xa = rnd(N, S, S, C, seed=95).to(DEV)
    xb = rnd(N, S, S, Cb, seed=96).to(DEV) if Cb else None
    Ct = C + Cb
    gamma, beta = (1 + 0.1 * rnd(Ct, seed=97)).to(DEV), (0.5 + rnd(Ct, seed=98)).to(DEV)
    part = torch.zeros((N, 32, 2), device=DEV)
    L.call('indm_gn_stats', L.ptr(xa), C, L.ptr(xb) if Cb else None, Cb, L.DTYPE_F32, N, S * S, 32, L.ptr(part))
    ctl = torch.tensor([seed, 1], dtype=torch.int64, device=DEV)
    head = (L.ptr(xa), C, L.ptr(xb) if Cb else None, Cb, L.DTYPE_F32, N, S, S, 32, L.ptr(part), L.ptr(gamma), L.ptr(beta), 1e-6, 1)
    ob = torch.empty((N, S, S, Ct), device=DEV, dtype=torch.bfloat16)
    of = torch.empty((N, S, S, Ct), device=DEV)
    plain = torch.empty((N, S, S, Ct), device=DEV)
    L.call('indm_gn_apply_dropout', *head, L.ptr(ob), L.DTYPE_BF16, p_drop, L.ptr(ctl), stream_id)
    L.call('indm_gn_apply_dropout', *head, L.ptr(of), L.DTYPE_TF32, p_drop, L.ptr(ctl), stream_id)
    L.call('indm_gn_apply', *head, 0, L.ptr(plain), None, L.DTYPE_TF32)
    pp = torch.zeros(((N * (S + 1) + 1) * (S + 2), Ct), device=DEV, dtype=torch.bfloat16)
    L.call('indm_gn_apply_pp', *head, L.ptr(pp), None, L.DTYPE_BF16, p_drop, L.ptr(ctl), stream_id)
    torch.cuda.synchronize()
    zb, zf = (ob == 0).cpu(), (of == 0).cpu()
    assert torch.equal(zb, zf)                                           # same elements dropped
    frac = float(zf.float().mean())
    assert abs(frac - p_drop) < 0.02, frac
    kept = ~zf
    assert rel_l2(of.cpu()[kept], plain.cpu()[kept] / (1 - p_drop)) < 1e-6     # kept elements scaled by 1 / (1 - p)
    assert rel_l2(ob.float().cpu(), of.cpu()) < 4e-3
    assert torch.equal(pp.cpu(), to_pp(ob.cpu()))
    ctl.zero_()                                                          # enabled = 0: the same launch is a plain GroupNorm apply
    L.call('indm_gn_apply_dropout', *head, L.ptr(ob), L.DTYPE_BF16, p_drop, L.ptr(ctl), stream_id)
    torch.cuda.synchronize()
    assert rel_l2(ob.float().cpu(), plain.cpu()) < 4e-3 and float((ob == 0).float().mean()) < 1e-3


def to_pp(x_nhwc):
    """[N,H,W,C] -> the padded-pixel buffer of indm_igemm_t.a_pp: [(N (H + 1) + 1)(W + 2), C], zero borders"""
    N, H, W, C = x_nhwc.shape
    buf = torch.zeros((N * (H + 1) + 1, W + 2, C), dtype=x_nhwc.dtype)
    v = buf[:N * (H + 1)].view(N, H + 1, W + 2, C)
    v[:, 1:, 1:W + 1] = x_nhwc
    return buf.reshape(-1, C)


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("N,S,Cin,Cin2,Cout", [(37, 8, 256, 0, 256), (128, 8, 256, 0, 256), (21, 4, 256, 0, 256), (128, 4, 256, 0, 256),
                                               (9, 8, 128, 256, 128), (7, 16, 128, 0, 256), (3, 8, 64, 0, 128), (1, 4, 64, 0, 128)])
def test_igemm_whole_batch_padded_pixel_operands(kind, N, S, Cin, Cin2, Cout):
    """indm_igemm_t.a_pp: 3x3 convolutions of small maps whose operands live in the zero-bordered padded-pixel buffer
    (indm_gn_apply_pp): the whole batch is one sequence of padded pixels cut into 128-pixel tiles that cross image boundaries,
    every tap a row offset into one shared-memory box.  Odd batch sizes (ragged last tile / idle second CTA of the last pair),
    the fused 1x1 skip segment, fused GroupNorm statistics of two consumers where a warp's rows span two images; against fp64."""
    dtype = L.DTYPE_BF16
    x = round_in(rnd(N, Cin, S, S, seed=70), dtype)
    w = round_in(rnd(Cout, Cin, 3, 3, seed=71) / math.sqrt(9 * Cin), dtype)
    bias = rnd(Cout, seed=72)
    kw = dict(dtype=dtype, a=to_pp(nhwc(x)).to(torch.bfloat16).to(DEV), a_pp=1, N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dtype), Cout=Cout,
              taps=9, bias=bias.to(DEV), scale=0.7071, out_ld=Cout)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    if Cin2:
        xs = round_in(rnd(N, Cin2, S, S, seed=73), dtype)
        w2 = round_in(rnd(Cout, Cin2, 1, 1, seed=74) / math.sqrt(Cin2), dtype)
        kw.update(a2=to_pp(nhwc(xs)).to(torch.bfloat16).to(DEV), Cin2=Cin2, b2=dev_op(w2.reshape(Cout, Cin2), dtype))
        want = want + F.conv2d(xs.double(), w2.double())
    want = nhwc(want)
    stats = (S + 1) * (S + 2) >= 16           # 4x4 maps: a warp's 32 rows touch up to three images
    part = torch.zeros((N, 32, 2), device=DEV)
    part2 = torch.zeros((N, 16, 2), device=DEV)
    gn = dict(gn_partial=part, gn_cpg=Cout // 32, gn_groups=32, gn2_partial=part2, gn2_cpg=Cout // 16, gn2_groups=16) if stats else {}
    if kind == 1:
        rowb = rnd(N, Cout, seed=75)
        out = torch.full((N, S, S, Cout), float('nan'), device=DEV, dtype=torch.bfloat16)
        kw.update(rowbias=rowb.to(DEV), rowbias_ld=Cout, out_bf16=out, **gn)
        want = (want + rowb[:, None, None, :].double()) * 0.7071
        tol = 4e-3
    else:
        res = rnd(N, S, S, Cout, seed=76)
        out = torch.full((N, S, S, Cout), float('nan'), device=DEV)
        kw.update(residual=res.to(DEV), res_ld=Cout, res_scale=0.7071, out_f32=out, **gn)
        want = want * 0.7071 + res.double() * 0.7071
        tol = 2e-5
    L.igemm(**kw)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    e = rel_l2(out.float().cpu(), want)
    print(f'a_pp kind {kind} N={N} S={S} Cin={Cin}+{Cin2} Cout={Cout}: rel-L2 {e:.2e}')
    assert e < tol
    if stats:
        for pt, G in ((part, 32), (part2, 16)):
            g = want.reshape(N, S * S, G, Cout // G)
            ws, wq = g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))
            assert rel_l2(pt[..., 0].cpu(), ws) < 2e-3 and rel_l2(pt[..., 1].cpu(), wq) < 2e-3
    first = out.clone()
    L.igemm(**kw)
    torch.cuda.synchronize()
    assert torch.equal(first, out)


def test_igemm_padded_pixel_rejects_what_it_cannot_run():
    x = torch.zeros(((2 * 5 + 1) * 6, 64), device=DEV, dtype=torch.bfloat16)
    w = torch.zeros((9, 128, 64), device=DEV, dtype=torch.bfloat16)
    out = torch.zeros((2, 4, 4, 128), device=DEV)
    part = torch.zeros((2, 32, 2), device=DEV)
    x2 = torch.zeros(((2 * 3 + 1) * 4, 64), device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):       # 2x2 maps: a warp's 32 rows can span four images, no fused statistics
        L.igemm(dtype=L.DTYPE_BF16, a=x2, a_pp=1, N=2, H=2, W=2, Cin=64, b=w, Cout=128, taps=9, out_f32=out, out_ld=128,
                gn_partial=part, gn_cpg=4, gn_groups=32)
    with pytest.raises(RuntimeError):       # 1x1 products have no padded-pixel form
        L.igemm(dtype=L.DTYPE_BF16, a=x, a_pp=1, N=2, H=4, W=4, Cin=64, b=w, Cout=128, taps=1, out_f32=out, out_ld=128)


@pytest.mark.parametrize("N,S,C,Cb", [(5, 8, 256, 0), (3, 4, 256, 256), (2, 16, 128, 0)])
def test_gn_apply_pp_is_gn_apply_scattered_into_the_padded_buffer(N, S, C, Cb):
    """indm_gn_apply_pp writes exactly what indm_gn_apply writes, at the padded-pixel rows, and never touches the border rows."""
    xa = rnd(N, S, S, C, seed=80).to(DEV)
    xb = rnd(N, S, S, Cb, seed=81).to(DEV) if Cb else None
    Ct = C + Cb
    gamma, beta = (1 + 0.1 * rnd(Ct, seed=82)).to(DEV), rnd(Ct, seed=83).to(DEV)
    part = torch.zeros((N, 32, 2), device=DEV)
    L.call('indm_gn_stats', L.ptr(xa), C, L.ptr(xb) if Cb else None, Cb, L.DTYPE_F32, N, S * S, 32, L.ptr(part))
    dense = torch.empty((N, S, S, Ct), device=DEV, dtype=torch.bfloat16)
    raw = torch.empty_like(dense)
    L.call('indm_gn_apply', L.ptr(xa), C, L.ptr(xb) if Cb else None, Cb, L.DTYPE_F32, N, S, S, 32, L.ptr(part), L.ptr(gamma), L.ptr(beta),
           1e-6, 1, 0, L.ptr(dense), L.ptr(raw), L.DTYPE_BF16)
    rows = (N * (S + 1) + 1) * (S + 2)
    pp = torch.zeros((rows, Ct), device=DEV, dtype=torch.bfloat16)
    pr = torch.zeros_like(pp)
    L.call('indm_gn_apply_pp', L.ptr(xa), C, L.ptr(xb) if Cb else None, Cb, L.DTYPE_F32, N, S, S, 32, L.ptr(part), L.ptr(gamma), L.ptr(beta),
           1e-6, 1, L.ptr(pp), L.ptr(pr), L.DTYPE_BF16, 0.0, None, 0)
    torch.cuda.synchronize()
    assert torch.equal(pp.cpu(), to_pp(dense.cpu())) and torch.equal(pr.cpu(), to_pp(raw.cpu()))


@pytest.mark.parametrize("N,S,Cin,Cout", [(16, 8, 256, 256), (32, 4, 256, 256)])
def test_igemm_split_k_path(N, S, Cin, Cout):
    """Launches with too few output tiles to fill the chip can split K across CTAs (indm_igemm_t.splitk_ws: raw partial sums in a
    caller workspace + a finish kernel that applies the epilogue).  The engine leaves it off (measured slower than one CTA per tile
    on the 4x4 / 8x8 layers at batch 128), but the path stays tested: bias + row bias + scale + fused statistics against fp64."""
    dtype = L.DTYPE_BF16
    x = round_in(rnd(N, Cin, S, S, seed=60), dtype)
    w = round_in(rnd(Cout, Cin, 3, 3, seed=61) / math.sqrt(9 * Cin), dtype)
    bias, rowb = rnd(Cout, seed=62), rnd(N, Cout, seed=63)
    ws = torch.empty(8 << 20, device=DEV)
    out = torch.zeros((N, S, S, Cout), device=DEV, dtype=torch.bfloat16)
    part = torch.zeros((N, 32, 2), device=DEV)
    L.igemm(dtype=dtype, a=dev_op(nhwc(x), dtype), N=N, H=S, W=S, Cin=Cin, b=pack_w(w, dtype), Cout=Cout, taps=9, bias=bias.to(DEV),
            rowbias=rowb.to(DEV), rowbias_ld=Cout, scale=0.7071, out_bf16=out, out_ld=Cout, splitk_ws=ws, splitk_ws_bytes=ws.numel() * 4,
            **(dict(gn_partial=part, gn_cpg=Cout // 32, gn_groups=32) if S * S >= 32 else {}))
    torch.cuda.synchronize()
    want = (nhwc(F.conv2d(x.double(), w.double(), bias.double(), padding=1)) + rowb[:, None, None, :].double()) * 0.7071
    assert rel_l2(out.float().cpu(), want) < 4e-3
    if S * S >= 32:
        g = want.reshape(N, S * S, 32, Cout // 32)
        assert rel_l2(part[..., 0].cpu(), g.sum(dim=(1, 3))) < 2e-3 and rel_l2(part[..., 1].cpu(), (g * g).sum(dim=(1, 3))) < 2e-3
