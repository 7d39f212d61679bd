"""Device-resident adaptive Runge-Kutta integrator.

The reference integrates the probability-flow ODE with `scipy.integrate.solve_ivp(..., method='RK45')` over one flattened
float64 numpy vector (likelihood.py:111-119, sampling.py:596-606): every right-hand-side evaluation moves the whole
state host -> device and the drift device -> host, six times per step.  This module restates SciPy's explicit
Dormand-Prince 5(4) pair — the third-party algorithm behind those two call sites, SciPy `integrate/_ivp/rk.py` (`RK45`,
`rk_step`, `RungeKutta._step_impl`) and `integrate/_ivp/common.py` (`select_initial_step`, `norm`), unpinned in the
reference's requirements.txt:12, 1.18.1 in this image — with the state, the seven stage derivatives and the error
estimate kept as float64 tensors on the device of `y0`.  Only ONE scalar (the RMS error norm of the step) crosses to the
host per attempted step, where the step-size controller runs exactly as in SciPy, so step sequences, `nfev` and results
agree with `solve_ivp` to float64 rounding (tests/test_ode_cpu.py pins this against SciPy itself on CPU tensors).

Only what the two call sites use is provided: `t_span`, `y0`, `rtol`, `atol`, no dense output, no events; the result
carries `.y_final` (the reference reads `solution.y[:, -1]`), `.t`, `.nfev`, `.status`, `.message`.
"""
import math

import numpy as np
import torch

# Dormand-Prince 5(4) tableau (SciPy rk.py, class RK45)
_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
_A = ((),
      (1 / 5,),
      (3 / 40, 9 / 40),
      (44 / 45, -56 / 15, 32 / 9),
      (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
      (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
_B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
_E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
_ORDER = 5
_ERROR_ESTIMATOR_ORDER = 4
_N_STAGES = 6
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0      # SciPy rk.py module constants
_EPS = float(np.finfo(np.float64).eps)


class OdeResult:
    """The fields of SciPy's `OdeResult` that the reference reads."""

    def __init__(self, t, y_final, nfev, status, message, n_steps, n_rejected):
        self.t, self.y_final, self.nfev, self.status, self.message = t, y_final, nfev, status, message
        self.success = status >= 0
        self.n_steps, self.n_rejected = n_steps, n_rejected

    @property
    def y(self):
        """`solution.y[:, -1]` compatibility: a [n, 1] view of the final state."""
        return self.y_final[:, None]


def _rms(x):
    """common.py `norm`: RMS norm, returned as a host float (the one device -> host scalar per use)."""
    return float(torch.linalg.vector_norm(x)) / math.sqrt(x.numel())


def _combine(K, coeffs, n):
    """sum_j coeffs[j] * K[j] over the first n stage rows (SciPy: np.dot(K[:n].T, coeffs[:n])); zero coefficients are skipped
    only where the tableau has structural zeros, which changes nothing in float64 (0 * finite = 0)."""
    out = None
    for j in range(n):
        c = coeffs[j]
        if c == 0.0:
            continue
        out = K[j] * c if out is None else out.add_(K[j], alpha=c)
    return out if out is not None else torch.zeros_like(K[0])


def _select_initial_step(fun, t0, y0, t_bound, f0, direction, rtol, atol, max_step=math.inf):
    """common.py `select_initial_step` (Hairer, Norsett & Wanner, Sec. II.4).  One extra right-hand-side evaluation."""
    if y0.numel() == 0:
        return math.inf
    interval_length = abs(t_bound - t0)
    if interval_length == 0.0:
        return 0.0
    scale = atol + y0.abs() * rtol
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval_length)
    y1 = y0 + h0 * direction * f0
    f1 = fun(t0 + h0 * direction, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / (_ERROR_ESTIMATOR_ORDER + 1))
    return min(100 * h0, h1, interval_length, max_step)


def solve_ivp_rk45(fun, t_span, y0, rtol=1e-3, atol=1e-6, max_step=math.inf, first_step=None):
    """`scipy.integrate.solve_ivp(fun, t_span, y0, method='RK45', rtol=, atol=)` with `y0` and every intermediate a float64
    tensor on `y0.device`.  `fun(t: float, y: Tensor[n] float64) -> Tensor[n]` (any float dtype; promoted to float64 like
    SciPy promotes the reference's float32 drifts when it stores them into its float64 stage array)."""
    t0, t_bound = float(t_span[0]), float(t_span[1])
    y = y0.detach().to(torch.float64).reshape(-1).clone()
    n = y.numel()
    rtol = max(float(rtol), 100 * _EPS)                      # common.py validate_tol
    atol = float(atol)
    direction = float(np.sign(t_bound - t0)) if t_bound != t0 else 1.0
    nfev = 0

    def f64(t, v):
        nonlocal nfev
        nfev += 1
        return fun(t, v).detach().to(torch.float64).reshape(-1)

    t = t0
    f = f64(t, y)
    if first_step is None:
        h_abs = _select_initial_step(f64, t, y, t_bound, f, direction, rtol, atol, max_step)
    else:
        h_abs = float(first_step)
    K = torch.empty((_N_STAGES + 1, n), dtype=torch.float64, device=y.device)
    error_exponent = -1.0 / (_ERROR_ESTIMATOR_ORDER + 1)
    n_steps = n_rejected = 0
    status, message = None, None

    while status is None:
        if n == 0 or t == t_bound:
            status, message = 0, "The solver successfully reached the end of the integration interval."
            break
        # ---- RungeKutta._step_impl
        min_step = 10 * abs(float(np.nextafter(t, direction * np.inf)) - t)
        if h_abs > max_step:
            h_abs = max_step
        elif h_abs < min_step:
            h_abs = min_step
        step_accepted = step_rejected = False
        failed = False
        while not step_accepted:
            if h_abs < min_step:
                failed = True
                break
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            # ---- rk_step
            K[0] = f
            for s in range(1, _N_STAGES):
                dy = _combine(K, _A[s], s).mul_(h)
                K[s] = f64(t + _C[s] * h, y + dy)
            y_new = y + _combine(K, _B, _N_STAGES).mul_(h)
            f_new = f64(t + h, y_new)
            K[_N_STAGES] = f_new
            # ---- error norm (RungeKutta._estimate_error_norm)
            scale = torch.maximum(y.abs(), y_new.abs()).mul_(rtol).add_(atol)
            error_norm = _rms(_combine(K, _E, _N_STAGES + 1).mul_(h).div_(scale))
            if error_norm < 1:
                factor = MAX_FACTOR if error_norm == 0 else min(MAX_FACTOR, SAFETY * error_norm ** error_exponent)
                if step_rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                step_accepted = True
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * error_norm ** error_exponent)
                step_rejected = True
                n_rejected += 1
        if failed:
            status, message = -1, "Required step size is less than spacing between numbers."
            break
        n_steps += 1
        t, y, f = t_new, y_new, f_new
        if direction * (t - t_bound) >= 0:
            status, message = 0, "The solver successfully reached the end of the integration interval."
    return OdeResult(t, y, nfev, status, message, n_steps, n_rejected)
