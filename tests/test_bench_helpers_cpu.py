"""bench.py host-side helpers that cannot be exercised without a GPU otherwise: the launch-closure introspection behind the
per-kernel roofline accounting (algorithmic bytes of a GroupNorm-apply launch), argument parsing of the profiling flags."""
import ctypes
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _engine_style_call(name, *args):
    """same closure shape as ScoreEngine._call (indm_b200/models/engine.py)"""
    fn = object()
    cargs = list(args)

    def run():
        return fn, cargs, name
    return run


def test_gn_apply_bytes_from_launch_closure():
    from indm_b200 import _lib as L
    N, H, W, C = 128, 32, 32, 128
    vp = ctypes.c_void_p(0x1000)
    common = (vp, C, None, 0, L.DTYPE_F32, ctypes.c_int64(N), H, W, 32, vp, vp, vp, ctypes.c_float(1e-6), 1)
    op = _engine_style_call('indm_gn_apply', *common, 0, vp, None, L.DTYPE_BF16)
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + N * H * W * C * 2
    op = _engine_style_call('indm_gn_apply', *common, 0, vp, vp, L.DTYPE_BF16)            # + raw operand copy
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + 2 * N * H * W * C * 2
    op = _engine_style_call('indm_gn_apply', *common, 1, vp, None, L.DTYPE_BF16)          # nearest up x2: 4x the output
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + 4 * N * H * W * C * 2
    op = _engine_style_call('indm_gn_apply', *common, 2, vp, None, L.DTYPE_BF16)          # mean down x2
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + N * H * W * C * 2 // 4
    bf = (vp, 64, vp, 64, L.DTYPE_BF16, ctypes.c_int64(N), H, W, 32, vp, vp, vp, ctypes.c_float(1e-6), 1)   # concat input, bf16
    op = _engine_style_call('indm_gn_apply', *bf, 0, vp, None, L.DTYPE_BF16)
    assert bench._gn_apply_bytes(op) == N * H * W * C * 2 * 2
    op = _engine_style_call('indm_gn_apply_dropout', *bf, vp, L.DTYPE_BF16, ctypes.c_float(0.1), vp, ctypes.c_uint32(3))
    assert bench._gn_apply_bytes(op) == N * H * W * C * 2 * 2
    # padded-pixel variant of the forward-only plan: the zero borders are never written, so the bytes are the dense ones
    op = _engine_style_call('indm_gn_apply_pp', *common, vp, vp, L.DTYPE_BF16, ctypes.c_float(0.0), None, ctypes.c_uint32(3))
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + 2 * N * H * W * C * 2
    op = _engine_style_call('indm_gn_apply_pp', *common, vp, None, L.DTYPE_BF16, ctypes.c_float(0.0), None, ctypes.c_uint32(3))
    assert bench._gn_apply_bytes(op) == N * H * W * C * 4 + N * H * W * C * 2
    assert bench._gn_apply_bytes(_engine_style_call('indm_gn_stats', vp)) is None
    assert bench._gn_apply_bytes(lambda: None) is None


def test_gn_apply_bytes_sees_the_dropout_switch_closure():
    """ScoreEngine._call_unless_dropping: one closure holding both the plain and the dropout launch; the accounting reads the
    plain one (`name` / `cargs`), which is what runs whenever the masks are off (sampling, bench)"""
    import torch  # noqa: F401  (engine import below needs it)
    from indm_b200 import _lib as L
    from indm_b200.models import engine as E

    class _FakeLib:
        def __getattr__(self, k):
            return lambda *a: 0

    class _Eng:
        _drop_on = False
        _cur = []
        _call_unless_dropping = E.ScoreEngine._call_unless_dropping

    real = L.lib
    L.lib = lambda: _FakeLib()
    try:
        N, H, W, C = 128, 16, 16, 256
        vp = ctypes.c_void_p(0x1000)
        head = (vp, C, None, 0, L.DTYPE_BF16, ctypes.c_int64(N), H, W, 32, vp, vp, vp, ctypes.c_float(1e-6), 1)
        e = _Eng()
        e._call_unless_dropping('indm_gn_apply', head + (0, vp, None, L.DTYPE_BF16),
                                'indm_gn_apply_dropout', head + (vp, L.DTYPE_BF16, ctypes.c_float(0.1), vp, ctypes.c_uint32(3)))
        op = e._cur[-1]
        assert bench._gn_apply_bytes(op) == N * H * W * C * 2 * 2
    finally:
        L.lib = real


def test_workload_config_slice_keeps_the_sde():
    cfg = bench.workload_config("cpu", 4)
    assert cfg.sampling.num_scales == 4 and cfg.model.num_scales == 1000
    assert cfg.sampling.predictor == "reverse_diffusion" and cfg.sampling.corrector == "none"
    assert bench.workload_config("cpu").sampling.num_scales == 1000
