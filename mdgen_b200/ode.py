"""Adaptive Dormand-Prince 5(4) integrator with dense output: the `torchdiffeq.odeint(func, y0, t,
method='dopri5', rtol=, atol=)` call of mdgen/transport/integrators.py:106-113, which is
the reference's DEFAULT sampler (`--sampling_method dopri5`, mdgen/parsing.py:102; rtol 1e-3, atol 1e-6,
mdgen/transport/transport.py:411-414).

torchdiffeq (third-party, unpinned in the reference's README.md:17, absent from this image) is restated
from its published algorithm (torchdiffeq 0.2.x `_impl/dopri5.py`, `_impl/rk_common.py`, `_impl/misc.py`):
  * Dormand-Prince-Shampine tableau, FSAL (the 7th stage of an accepted step is f0 of the next),
  * mixed error tolerance  atol + rtol * max(|y0|, |y1|)  under the RMS norm over the WHOLE state tensor
    (one step size for the whole batch),
  * Hairer's initial step (order 4 for the selection, as torchdiffeq passes `order - 1`),
  * step controller  h <- h * min(10, max(0.9 / ratio^(1/5), 0.2))  (0.2 -> 1 when the step is accepted),
  * NO clipping of steps to the output times: the solver steps past every requested time and evaluates
    the quartic Hermite-type interpolant fitted through (y0, y_mid, y1, f0, f1) - so the right-hand side
    IS evaluated at t > 1 during the last step, exactly as under the reference.
Step sizes and times are kept as Python floats (torchdiffeq keeps them as fp32 tensors): accept/reject
decisions can therefore differ from a live torchdiffeq at rounding level; parity of this row is
"unpinned"  and is tested by replaying the accepted step sequence through the oracle.

The right-hand side is the fused CUDA forward (`LatentMDGenModel.forward_inference` -> mdgen_forward). With an
`engine` (the sampler passes the model's) every stage / solution / error / dense-output combination is ONE pass of
`lincomb_kernel` (mdgen_lincomb) and the mixed-tolerance RMS error ratio is `mdgen_rk_error_ratio`; without one (the
CPU tests of the integrator itself) the same formulas run as torch ops.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch

# Dormand-Prince-Shampine tableau (torchdiffeq _impl/dopri5.py)
ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
BETA = (
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
C_SOL = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0)
C_ERROR = (
    35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0,
)
C_MID = (
    6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
)
SAFETY, IFACTOR, DFACTOR, ORDER = 0.9, 10.0, 0.2, 5


def rms_norm(x: torch.Tensor) -> float:
    return float(x.float().pow(2).mean().sqrt())


def _lincomb(y: Optional[torch.Tensor], h: float, coeffs: Sequence[float], ks: Sequence[torch.Tensor], engine=None
             ) -> torch.Tensor:
    """(y or 0) + h * sum_i coeffs[i] * ks[i]."""
    if engine is not None and ks[0].is_cuda:
        return engine.lincomb(y, h, coeffs, ks)
    out = y.clone() if y is not None else torch.zeros_like(ks[0])
    for c, k in zip(coeffs, ks):
        if c != 0.0:
            out.add_(k, alpha=h * c)
    return out


def _error_ratio(err, y0, y1, rtol: float, atol: float, engine=None) -> float:
    """sqrt(mean((err / (atol + rtol * max(|y0|, |y1|)))^2)) over the whole state (torchdiffeq's mixed norm)."""
    if engine is not None and err.is_cuda:
        return engine.rk_error_ratio(err, y0, y1, rtol, atol)
    return rms_norm(err / (atol + rtol * torch.maximum(y0.abs(), y1.abs())))


def rk_step(func: Callable, t0: float, h: float, y0: torch.Tensor, f0: torch.Tensor, engine=None):
    """One Dormand-Prince step from (t0, y0) with f0 = func(t0, y0).
    Returns y1 (5th order), f1 = func(t0+h, y1), the error estimate, y_mid and the number of func calls (6)."""
    ks: List[torch.Tensor] = [f0]
    for a, row in zip(ALPHA, BETA):
        yi = _lincomb(y0, h, row, ks, engine)
        ks.append(func(t0 + a * h, yi))
    # the last stage is evaluated at y1 itself (BETA[-1] == C_SOL[:6]): FSAL
    y1 = _lincomb(y0, h, C_SOL, ks, engine)
    f1 = ks[-1]
    err = _lincomb(None, h, C_ERROR, ks, engine)
    y_mid = _lincomb(y0, h, C_MID, ks, engine)
    return y1, f1, err, y_mid, 6


def interp_fit(y0, y1, y_mid, f0, f1, h: float, engine=None):
    """Coefficients (a, b, c, d, e) of the quartic p(x), x = (t - t0)/h, with p(0)=y0, p(1/2)=y_mid,
    p(1)=y1, p'(0)=h f0, p'(1)=h f1  (torchdiffeq _impl/interp.py)."""
    if engine is not None and y0.is_cuda:
        ks = (f0, f1, y0, y1, y_mid)
        a = engine.lincomb(None, 1.0, (-2 * h, 2 * h, -8.0, -8.0, 16.0), ks)
        b = engine.lincomb(None, 1.0, (5 * h, -3 * h, 18.0, 14.0, -32.0), ks)
        c = engine.lincomb(None, 1.0, (-4 * h, h, -11.0, -5.0, 16.0), ks)
        d = engine.lincomb(None, h, (1.0,), (f0,))
        return a, b, c, d, y0
    a = 2 * h * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = h * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = h * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = h * f0
    e = y0
    return a, b, c, d, e


def interp_evaluate(coeffs, t0: float, t1: float, t: float, engine=None) -> torch.Tensor:
    a, b, c, d, e = coeffs
    x = (t - t0) / (t1 - t0)
    if engine is not None and e.is_cuda:
        return engine.lincomb(e, 1.0, (x ** 4, x ** 3, x ** 2, x), (a, b, c, d))
    return e + x * (d + x * (c + x * (b + x * a)))


def select_initial_step(func: Callable, t0: float, y0: torch.Tensor, f0: torch.Tensor, rtol: float,
                        atol: float, order: int = ORDER - 1, engine=None):
    """Hairer, Norsett & Wanner I, II.4 as in torchdiffeq _impl/misc.py:_select_initial_step. Returns (h, nfe).
    (scale = atol + |y0| rtol is the mixed tolerance with y1 = y0.)"""
    d0 = _error_ratio(y0, y0, y0, rtol, atol, engine)
    d1 = _error_ratio(f0, y0, y0, rtol, atol, engine)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    y1 = _lincomb(y0, h0, (1.0,), (f0,), engine)
    f1 = func(t0 + h0, y1)
    d2 = _error_ratio(_lincomb(None, 1.0, (1.0, -1.0), (f1, f0), engine), y0, y0, rtol, atol, engine) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return min(100 * h0, h1), 1


def optimal_step_size(last_step: float, error_ratio: float) -> float:
    """torchdiffeq _impl/misc.py:_optimal_step_size with safety 0.9, ifactor 10, dfactor 0.2, order 5."""
    if error_ratio == 0:
        return last_step * IFACTOR
    dfactor = 1.0 if error_ratio < 1 else DFACTOR
    factor = min(IFACTOR, max(SAFETY / error_ratio ** (1.0 / ORDER), dfactor))
    return last_step * factor


def dopri5_integrate(func: Callable[[float, torch.Tensor], torch.Tensor], y0: torch.Tensor,
                     t_grid: Sequence[float], rtol: float = 1e-3, atol: float = 1e-6,
                     last_only: bool = True, stats: Optional[dict] = None, max_steps: int = 100000, engine=None):
    """Solution of y' = func(t, y), y(t_grid[0]) = y0 at the times `t_grid` (increasing).
    Returns the state at t_grid[-1] (last_only) or the stacked states at every grid time (what the reference's
    odeint returns; its caller only uses [-1], wrapper.py:444-447).
    `stats`, when given, receives nfe, accepted, rejected and `steps` = [(t0, h)] of the accepted steps."""
    ts = [float(t) for t in t_grid]
    if len(ts) < 2 or any(b <= a for a, b in zip(ts, ts[1:])):
        raise ValueError("t_grid must hold at least two strictly increasing times")
    t0 = ts[0]
    f0 = func(t0, y0)
    nfe = 1
    h, n = select_initial_step(func, t0, y0, f0, rtol, atol, engine=engine)
    nfe += n
    # state of the solver: the last accepted step [t_lo, t_hi] with its interpolant
    t_lo = t_hi = t0
    y_hi, f_hi = y0, f0
    coeffs = (torch.zeros_like(y0),) * 4 + (y0,)
    outs = [y0] if not last_only else None
    accepted = rejected = 0
    steps = []
    for t_next in ts[1:]:
        while t_next > t_hi:
            if accepted + rejected >= max_steps:
                raise RuntimeError("dopri5: max_steps exceeded")
            y1, f1, err, y_mid, n = rk_step(func, t_hi, h, y_hi, f_hi, engine)
            nfe += n
            ratio = _error_ratio(err, y_hi, y1, rtol, atol, engine)
            if not math.isfinite(ratio):
                raise FloatingPointError("dopri5: non-finite error estimate")
            if ratio <= 1:
                coeffs = interp_fit(y_hi, y1, y_mid, f_hi, f1, h, engine)
                steps.append((t_hi, h))
                t_lo, t_hi = t_hi, t_hi + h
                y_hi, f_hi = y1, f1
                accepted += 1
            else:
                rejected += 1
            h = optimal_step_size(h, ratio)
        if outs is not None:
            outs.append(interp_evaluate(coeffs, t_lo, t_hi, t_next, engine))
    if stats is not None:
        stats.update(nfe=nfe, accepted=accepted, rejected=rejected, steps=steps)
    if last_only:
        return interp_evaluate(coeffs, t_lo, t_hi, ts[-1], engine)
    return torch.stack(outs)
