"""ctypes binding of the C-ABI library `libindm_b200.so` (declared in include/indm_b200.h).

There is NO fallback: if the library is missing, or a call returns a non-zero status, a RuntimeError is raised.
PyTorch is used only as plumbing around it (device memory, streams): every wrapper passes raw device pointers and
the current CUDA stream handle.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libindm_b200.so")

DTYPE_BF16, DTYPE_TF32, DTYPE_F32 = 0, 1, 2

_lib = None
launches = 0  # number of kernel-launching C-ABI calls issued through this module (bench.py reports it)
param_epoch = 0  # bumped whenever parameters are updated through raw pointers (fused optimizer, EMA copy): engines repack


class IgemmDesc(C.Structure):
    """Mirror of `indm_igemm_t` (include/indm_b200.h)."""
    _fields_ = [
        ("dtype", C.c_int32),
        ("a", C.c_void_p), ("a_ld", C.c_int64), ("a_img_stride", C.c_int64),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("b", C.c_void_p), ("b_ld", C.c_int64), ("b_tap_stride", C.c_int64),
        ("Cout", C.c_int32), ("taps", C.c_int32), ("batched_b", C.c_int32),
        ("a2", C.c_void_p), ("a2_ld", C.c_int64), ("Cin2", C.c_int32),
        ("b2", C.c_void_p), ("b2_ld", C.c_int64),
        ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rowbias_ld", C.c_int64),
        ("residual", C.c_void_p), ("res_ld", C.c_int64), ("res_scale", C.c_float), ("act", C.c_int32),
        ("rowscale", C.c_void_p), ("scale", C.c_float),
        ("out_mode", C.c_int32), ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("out_ld", C.c_int64),
        ("tcol0", C.c_int32), ("out_t", C.c_void_p), ("round_tf32_out", C.c_int32),
        ("gn_partial", C.c_void_p), ("gn_cpg", C.c_int32), ("gn_groups", C.c_int32),
        ("block_n", C.c_int32), ("stride", C.c_int32),
        ("a_H", C.c_int32), ("a_W", C.c_int32), ("pad", C.c_int32),
        ("mul", C.c_void_p), ("mul_ld", C.c_int64), ("aux_cos", C.c_void_p),
        ("gn_goff", C.c_int32), ("gn2_partial", C.c_void_p), ("gn2_cpg", C.c_int32), ("gn2_groups", C.c_int32), ("gn2_goff", C.c_int32),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64),
        ("a_pp", C.c_int32),
    ]


class FlowOp(C.Structure):
    """Mirror of `indm_flow_op_t`."""
    _fields_ = [("kind", C.c_int32), ("backward", C.c_int32), ("split_skip", C.c_int32), ("up", C.c_int32), ("off", C.c_int64 * 6)]


_i32, _i64, _u64, _f32, _vp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p
_SIGS = {
    "indm_upfirdn2d_f32": [_vp, _vp, _vp, _i64] + [C.c_int] * 12 + [_vp],
    "indm_bias_act_f32": [_vp, _vp, _vp, _vp, _i64, C.c_int, _i64, C.c_int, C.c_int, _f32, _f32, _vp],
    "indm_igemm": [C.POINTER(IgemmDesc), _vp],
    "indm_gn_stats": [_vp, C.c_int, _vp, C.c_int, C.c_int, _i64, _i64, C.c_int, _vp, _vp],
    "indm_gn_apply": [_vp, C.c_int, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _f32, C.c_int,
                      C.c_int, _vp, _vp, C.c_int, _vp],
    "indm_softmax_rows": [_vp, _vp, _i64, C.c_int, C.c_int, _vp],
    "indm_prep_input": [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _f32, _f32, C.c_int, C.c_int, _vp],
    "indm_time_embedding": [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _i64, C.c_int, _vp, _vp],
    "indm_linear_f32": [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "indm_fir_nhwc": [_vp, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, _vp],
    "indm_pc_predictor_update": [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _i64, _u64, _vp, _u64, _vp],
    "indm_langevin_norms": [_vp, _vp, _vp, _vp, _i64, _i64, _u64, _vp, _u64, _vp],
    "indm_langevin_update": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _i64, _u64, _vp, _u64, _vp],
    "indm_langevin_norm_sums": [_vp, _vp, _i64, _vp],
    "indm_langevin_update_global": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _i64, _u64, _vp, _u64, _vp],
    "indm_advance_step": [_vp, _vp],
    "indm_randn_f32": [_vp, _i64, _u64, _u64, _vp],
    "indm_prior_flow": [_vp, _vp, _vp, _vp, _vp, C.c_int, _f32, _vp, _i64, _vp],
    "indm_posterior_sample": [_vp, _vp, _vp, _vp, _i64, _vp],
    "indm_axpy_f32": [_vp, _vp, _f32, _i64, _vp],
    "indm_im2col3x3_nchw": [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "indm_col2im3x3_nchw": [_vp, _i64, _vp, _vp, _vp, _f32, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "indm_series_step_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "indm_cos2pi_f32": [_vp, _vp, _i64, _vp],
    "indm_fixed_point_check": [_vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp],
    "indm_sched_broadcast": [_vp, _i64, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp],
    "indm_gn_bwd_stats": [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _f32,
                          C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _f32, _vp, C.c_uint32, _vp],
    "indm_gn_bwd_apply": [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _f32,
                          C.c_int, C.c_int, _vp, _vp, _vp, _f32, _vp, C.c_int, _vp, C.c_int, C.c_int, _f32, _vp, C.c_uint32, _vp],
    "indm_gn_apply_dropout": [_vp, C.c_int, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _f32, C.c_int,
                              _vp, C.c_int, _f32, _vp, C.c_uint32, _vp],
    "indm_gn_apply_pp": [_vp, C.c_int, _vp, C.c_int, C.c_int, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _f32, C.c_int,
                         _vp, _vp, C.c_int, _f32, _vp, C.c_uint32, _vp],
    "indm_attention_fwd": [_vp, _vp, _i64, C.c_int, C.c_int, _f32, C.c_int, _vp],
    "indm_cast_scale": [_vp, _vp, _i64, _f32, C.c_int, _vp],
    "indm_colsum": [_vp, C.c_int, _i64, _i64, C.c_int, _i64, _vp, _i64, _vp, _f32, _vp],
    "indm_sgemm_batched_f32": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32, _vp, _i64, _i64, _vp, _i64, _i64, _f32, _vp, _i64, _i64,
                               C.c_int, _vp],
    "indm_sgemm_f32": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32, _vp, _i64, _vp, _i64, _f32, _vp, _i64, _vp],
    "indm_silu_bwd_f32": [_vp, _vp, _vp, _i64, _vp],
    "indm_perturb_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "indm_dsm_loss_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _vp],
    "indm_sumsq_f32": [_vp, _i64, _vp, _vp],
    "indm_adamw_ema_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i64, _vp, _f32, _f32, _vp],
    "indm_ema_f32": [_vp, _vp, _i64, _f32, _vp],
    "indm_conv_wgrad": [_vp, _i64, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _i64, _i64, _i64,
                        _f32, _vp],
    "indm_softmax_bwd_rows": [_vp, _vp, _vp, _i64, C.c_int, _f32, C.c_int, _vp],
    "indm_transpose_batched": [_vp, _vp, _i64, C.c_int, C.c_int, _i64, _i64, C.c_int, _vp],
    "indm_nchw_to_nhwc": [_vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _f32, C.c_int, _vp],
    "indm_rowdot_f32": [_vp, _vp, _vp, _i64, _i64, _f32, C.c_int, _vp],
    "indm_nhwc_to_nchw_f32": [_vp, _i64, _vp, _i64, C.c_int, C.c_int, C.c_int, _f32, _vp],
    "indm_conv_s2_dgrad": [_vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "indm_mul_op": [_vp, _vp, _vp, _i64, C.c_int, _vp],
    "indm_fma3_op": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, C.c_int, _vp],
    "indm_sin2pi_f32": [_vp, _vp, _i64, _vp],
    "indm_rowscale_f32": [_vp, _vp, _vp, _i64, _i64, _vp],
    "indm_lop_bwd_f32": [_vp, _vp, _vp, C.c_int, C.c_int, _f32, C.c_int, _vp],
    "indm_prior_flow_bwd": [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "indm_posterior_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "indm_bn_stats": [_vp, _i64, C.c_int, C.c_int, _vp, _vp],
    "indm_bn_apply": [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp],
    "indm_bn_bwd_stats": [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, C.c_int, _vp],
    "indm_bn_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp],
    "indm_conv_s2_wgrad": [_vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
}
EXPORTS = ["indm_version", "indm_last_error"] + list(_SIGS)


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C indm_b200/csrc`). indm_b200 has no CPU / PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        h.indm_version.restype = C.c_char_p
        h.indm_last_error.restype = C.c_char_p
        for name, args in _SIGS.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = h
    return _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("indm_b200: expected a CUDA tensor (there is no CPU path)")
    return C.c_void_p(t.data_ptr())


_SYNC_DEBUG = bool(os.environ.get("INDM_SYNC_DEBUG"))     # development: synchronise after every C-ABI launch and name the one that faults


def check(rc, what):
    global launches
    if rc != 0:
        raise RuntimeError(f"indm_b200.{what} failed (status {rc}): {lib().indm_last_error().decode()}")
    launches += 1
    if _SYNC_DEBUG and not torch.cuda.is_current_stream_capturing():
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise RuntimeError(f"indm_b200.{what}: device fault surfaced after this launch: {e}") from e


def debug_sync(what):
    """development (INDM_SYNC_DEBUG): synchronise and name `what` if a device fault surfaces here (graph replays, captures)"""
    if _SYNC_DEBUG and not torch.cuda.is_current_stream_capturing():
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise RuntimeError(f"indm_b200: device fault surfaced after {what}: {e}") from e


def call(name, *args):
    check(getattr(lib(), name)(*args, _stream()), name)


def igemm(**kw):
    d = IgemmDesc()
    d.scale = 1.0
    d.res_scale = 1.0
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(d, k, v)
    check(lib().indm_igemm(C.byref(d), _stream()), "igemm")
