"""Precision policy of the hot path: which arithmetic each leg runs in BY DEFAULT, chosen so that the mode that is timed is the
mode that meets BASELINE.json's stated tolerance for that leg.

Two arithmetic modes exist in the kernels (csrc/igemm.cu): 'bf16' (BF16 operands, FP32 accumulate — the tensor-core production
rate) and 'tf32' (error-compensated 3xTF32: FP32 operands split hi/lo inside the GEMM, FP32-class results).

| leg                                             | default | stated tolerance it has to meet (north_star)             |
|-------------------------------------------------|---------|----------------------------------------------------------|
| score net, sampling (PC / ODE sampler)          | bf16    | score output 2e-2 rel-L2 in BF16                         |
| score net, training step                        | bf16    | same                                                     |
| score net, likelihood (PF-ODE drift + VJP, ELBO) | tf32   | NLL / NELBO within 0.01 bpd                              |
| flow reverse (sampling), forward map            | tf32    | inverse round trip 1e-4 max-abs                          |
| flow eval forward with log-det (NLL / NELBO)    | tf32    | log-det 1e-3 relative                                    |
| flow training: posterior encoder, fc, KL        | tf32    | (log-det - KL) 1e-3 relative (the KL carries the value)  |
| flow training: iResBlocks forward + backward    | bf16    | see note                                                 |

Note on the training-mode iResBlocks.  Measured at the benched size against the live reference (tests/test_fullsize_gpu.py): with
the encoder legs in TF32 the loss term (log-det - KL) is within 1e-5 relative in either block precision; the block log-det ALONE
(a Hutchinson estimate of ~0.03 nats on the fixtures' weights) is within 3e-5 relative in TF32 and 3e-3 - 1e-2 relative in BF16.
BF16 blocks are the training default because the TF32 blocks cost 5.5x (600 ms vs 109 ms per joint step at batch 128 on B200,
tools/precision_probe.py); `set_policy('flow', 'training', 'tf32')` selects them, and bench.py reports both.

A net whose `compute_mode` attribute is 'bf16' or 'tf32' ignores the policy (tests, side-by-side benchmarks);
`compute_mode = 'auto'` (the default) follows it.  `set_policy(...)` changes a leg globally, e.g. to time BF16 everywhere.
"""
import contextlib
import threading

MODES = ('bf16', 'tf32')

POLICY = {
    'score': {'sampling': 'bf16', 'training': 'bf16', 'likelihood': 'tf32'},
    'flow': {'reverse': 'tf32', 'eval': 'tf32', 'training': 'bf16', 'encoder': 'tf32'},
}

_tl = threading.local()


def set_policy(kind, leg, mode):
    if mode not in MODES:
        raise ValueError(f'unknown compute mode {mode!r}')
    POLICY[kind][leg] = mode


@contextlib.contextmanager
def purpose(name):
    """marks the calls made inside as belonging to leg `name` ('likelihood'): read by `resolve` on the calling thread"""
    old = getattr(_tl, 'purpose', None)
    _tl.purpose = name
    try:
        yield
    finally:
        _tl.purpose = old


def current_purpose():
    return getattr(_tl, 'purpose', None)


def resolve(kind, requested, leg):
    """requested: the net's compute_mode ('auto' / 'bf16' / 'tf32'); leg: what the call is ('sampling', 'training', 'reverse', ...).
    An enclosing `purpose('likelihood')` wins over the leg for the score net and selects the eval leg's precision for the flow."""
    if requested in MODES:
        return requested
    if requested not in (None, 'auto'):
        raise ValueError(f'unknown compute mode {requested!r}')
    p = current_purpose()
    if p == 'likelihood':
        return POLICY[kind]['likelihood' if kind == 'score' else 'eval']
    return POLICY[kind][leg]
