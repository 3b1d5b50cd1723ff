"""Stand-in for torchdiffeq.odeint (reference call site: mdgen/transport/integrators.py:3,
106-113). Only the fixed-grid explicit Euler method is restated: the grid is the `t` tensor
passed by the caller, y_{i+1} = y_i + (t_{i+1}-t_i) * f(t_i, y_i), and the stacked solution at
every grid point is returned (torchdiffeq's FixedGridODESolver with step_size=None)."""
import torch


def odeint(func, y0, t, *, method="euler", rtol=None, atol=None, options=None, **_):
    if method != "euler":
        raise NotImplementedError(f"torchdiffeq stand-in implements 'euler' only, got {method!r}")
    ys = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        y = y + dt * func(t0, y)
        ys.append(y)
    return torch.stack(ys, 0)
