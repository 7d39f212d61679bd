"""GPU: the reference's public `op` API (op.upfirdn2d, op.fused_leaky_relu, op.FusedLeakyReLU) on the C-ABI kernels —
forward, first derivative and second derivative — against golden vectors produced by the live reference
(tests/golden/ops.npz).  Tolerance (north_star): 1e-6 relative in FP32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_npz, rel_l2  # noqa: E402
from indm_b200 import op  # noqa: E402

CASES = ['up2', 'down2', 'pyr', 'k3', 'crop', 'gen', 'up2_32']


@pytest.mark.parametrize("case", CASES)
def test_upfirdn2d_forward_backward_double_backward(case):
    g = load_npz('ops.npz')
    up, down, p0, p1 = [int(v) for v in g[f'upfirdn_{case}_args']]
    x = torch.from_numpy(g[f'upfirdn_{case}_x']).cuda().requires_grad_(True)
    k = torch.from_numpy(g[f'upfirdn_{case}_k']).cuda()
    y = op.upfirdn2d(x, k, up=up, down=down, pad=(p0, p1))
    assert rel_l2(y.detach().cpu().numpy(), g[f'upfirdn_{case}_y']) < 1e-6
    gy = torch.from_numpy(g[f'upfirdn_{case}_gy']).cuda().requires_grad_(True)
    gx, = torch.autograd.grad(y, x, gy, create_graph=True)
    assert rel_l2(gx.detach().cpu().numpy(), g[f'upfirdn_{case}_gx']) < 1e-6
    # the operator is linear in x: d<gx, v>/d gy = upfirdn2d(v) — the double-backward path of op/upfirdn2d.py:64-85
    v = torch.randn_like(x)
    ggy, = torch.autograd.grad(gx, gy, v)
    want = op.upfirdn2d(v, k, up=up, down=down, pad=(p0, p1))
    assert rel_l2(ggy.cpu().numpy(), want.detach().cpu().numpy()) < 1e-6


@pytest.mark.parametrize("case", ['4d', '2d', '3d'])
def test_fused_leaky_relu_forward_backward(case):
    g = load_npz('ops.npz')
    x = torch.from_numpy(g[f'lrelu_{case}_x']).cuda().requires_grad_(True)
    b = torch.from_numpy(g[f'lrelu_{case}_b']).cuda().requires_grad_(True)
    y = op.fused_leaky_relu(x, b)
    assert rel_l2(y.detach().cpu().numpy(), g[f'lrelu_{case}_y']) < 1e-6
    gy = torch.from_numpy(g[f'lrelu_{case}_gy']).cuda().requires_grad_(True)
    gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
    assert rel_l2(gx.detach().cpu().numpy(), g[f'lrelu_{case}_gx']) < 1e-6
    assert rel_l2(gb.detach().cpu().numpy(), g[f'lrelu_{case}_gb']) < 1e-5
    # second derivative w.r.t. gy: the map gy -> gx is linear with the sign pattern of y
    v = torch.randn_like(x)
    ggy, = torch.autograd.grad(gx, gy, v)
    slope = torch.where(y.detach() >= 0, torch.ones_like(y), torch.full_like(y, 0.2)) * 2 ** 0.5
    assert rel_l2(ggy.cpu().numpy(), (v * slope).cpu().numpy()) < 1e-6


def test_fused_leaky_relu_module_and_cpu_refusal():
    m = op.FusedLeakyReLU(6).cuda()
    x = torch.randn(2, 6, 5, 7, device='cuda')
    y = m(x)
    want = torch.nn.functional.leaky_relu(x, 0.2) * 2 ** 0.5
    assert rel_l2(y.detach().cpu().numpy(), want.cpu().numpy()) < 1e-6
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.zeros(1, 2, 4), torch.zeros(2))
