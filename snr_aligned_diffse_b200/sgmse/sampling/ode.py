"""Explicit Runge-Kutta 5(4) (Dormand-Prince) integrator with the state resident on the GPU.

The reference's `ode_sampler` (sgmse-bbed/sgmse/sampling/__init__.py:149-161) hands a flattened numpy copy of the
state to `scipy.integrate.solve_ivp(method='RK45')` and moves it host<->device at every right-hand-side evaluation.
Here the state, the seven stage derivatives and every linear combination stay on the device
(`snrse_rk_combine`, `snrse_rk_scaled_sqnorm`); one scalar (the scaled error norm) crosses to the host per step
attempt.  Step-size control restates scipy's published algorithm (scipy 1.8 `integrate/_ivp/rk.py`: `RK45`,
`rk_step`, `RungeKutta._step_impl`; `_ivp/common.py: select_initial_step`), the third-party solver the reference
pins, so the accepted step sequence is the same up to float32-vs-complex128 state rounding.
"""
import math

import numpy as np
import torch

from ... import ops

C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
A = ((),
     (1 / 5,),
     (3 / 40, 9 / 40),
     (44 / 45, -56 / 15, 32 / 9),
     (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
     (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
SAFETY, MIN_FACTOR, MAX_FACTOR, ERR_EXP = 0.9, 0.2, 10.0, -1.0 / 5.0


class DeviceRK45Result:
    def __init__(self, y, t, nfev, n_steps, n_rejected, status):
        self.y, self.t, self.nfev, self.n_steps, self.n_rejected, self.status = y, t, nfev, n_steps, n_rejected, status


def _initial_step(fun, t0, y0, f0, K, direction, t_bound, rtol, atol, ops=ops):
    interval = abs(t_bound - t0)
    one = (1.0,)
    d0 = ops.rk_scaled_norm(y0.reshape((1,) + y0.shape), one, 1.0, y0, None, atol, rtol)
    d1 = ops.rk_scaled_norm(K[0:1], one, 1.0, y0, None, atol, rtol)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval)
    y1 = ops.rk_combine(y0, K[0:1], one, h0 * direction)
    K[1].copy_(fun(t0 + h0 * direction, y1))
    d2 = ops.rk_scaled_norm(K[0:2], (-1.0, 1.0), 1.0, y0, None, atol, rtol) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    return min(100 * h0, h1, interval)


def rk45_integrate(fun, t0, y0, t_bound, rtol=1e-5, atol=1e-5, max_step=math.inf, max_nfev=None, _kernels=None):
    """Integrate dy/dt = fun(t, y) from t0 to t_bound.  `fun(t: float, y: complex64 cuda tensor) -> tensor` of the same
    shape.  Returns DeviceRK45Result with the state at t_bound (status 0) or where the solver stopped (status -1).
    `_kernels` (tests only): an object with `rk_combine` / `rk_scaled_norm` standing in for the CUDA operators, so the
    step controller can be checked against scipy on the CPU; the product path always uses the sm_100a library."""
    ops = _kernels if _kernels is not None else globals()["ops"]
    if _kernels is None and not (y0.is_cuda and y0.dtype == torch.complex64):
        raise ValueError("rk45_integrate: the state must be a complex64 CUDA tensor")
    y = y0.contiguous().clone()
    K = torch.empty((7,) + tuple(y.shape), dtype=y.dtype, device=y.device)
    direction = float(np.sign(t_bound - t0)) if t_bound != t0 else 1.0
    t = float(t0)
    K[0].copy_(fun(t, y))
    nfev = 1
    h_abs = _initial_step(fun, t, y, K[0], K, direction, t_bound, rtol, atol, ops)
    nfev += 1
    n_steps = n_rej = 0
    status = 0
    while direction * (t - t_bound) < 0:
        min_step = 10 * abs(np.nextafter(t, direction * np.inf) - t)
        if h_abs > max_step:
            h_abs = max_step
        elif h_abs < min_step:
            h_abs = min_step
        accepted = rejected = False
        while not accepted:
            if h_abs < min_step or (max_nfev is not None and nfev >= max_nfev):
                status = -1
                break
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            for s in range(1, 6):
                ys = ops.rk_combine(y, K[0:s], A[s], h)
                K[s].copy_(fun(t + C[s] * h, ys))
            y_new = ops.rk_combine(y, K[0:6], B, h)
            K[6].copy_(fun(t + h, y_new))
            nfev += 6
            err = ops.rk_scaled_norm(K, E, h, y, y_new, atol, rtol)
            if err < 1:
                factor = MAX_FACTOR if err == 0 else min(MAX_FACTOR, SAFETY * err ** ERR_EXP)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                accepted = True
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * err ** ERR_EXP)
                rejected = True
                n_rej += 1
        if status != 0:
            break
        t, y = t_new, y_new
        K[0].copy_(K[6])                      # first-same-as-last
        n_steps += 1
    return DeviceRK45Result(y, t, nfev, n_steps, n_rej, status)
