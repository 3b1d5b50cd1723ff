"""Host-side adaptive sampler (the reference's default `--sampling_method dopri5`,
mdgen/transport/integrators.py:106-113 -> torchdiffeq.odeint) - CPU tests of mdgen_b200/ode.py and of its
oracle restatement. torchdiffeq is absent from this image, so the pins are: scipy's independent RK45
tableau, the order conditions of the embedded pair, the defining conditions of the dense output, analytic
solutions, and product-vs-oracle agreement."""
import functools
import math

import numpy as np
import pytest
import torch

from mdgen_b200 import ode
from oracle import mdgen_oracle as O


def test_tableau_matches_scipy_rk45():
    from scipy.integrate._ivp.rk import RK45
    assert np.allclose(RK45.C[1:], ode.ALPHA[:5], rtol=0, atol=1e-15)
    for i, row in enumerate(ode.BETA[:5]):                      # scipy's A holds stages 2..6
        assert np.allclose(RK45.A[i + 1, :len(row)], row, rtol=0, atol=1e-15)
    assert np.allclose(RK45.B, ode.C_SOL[:6], rtol=0, atol=1e-15)
    assert np.allclose(ode.BETA[5], ode.C_SOL[:6])              # FSAL: last stage is evaluated at y1
    assert ode.ALPHA[5] == 1.0 and ode.C_SOL[6] == 0.0
    # the oracle's independent copy
    assert np.allclose(O._DP_BETA.numpy()[:5, :5], RK45.A[1:6, :5], atol=1e-15)
    assert np.allclose(O._DP_CSOL.numpy()[:6], RK45.B, atol=1e-15)
    assert np.allclose(O._DP_CERR.numpy(), ode.C_ERROR, atol=0) and np.allclose(O._DP_CMID.numpy(), ode.C_MID, atol=0)


def _full_tableau():
    c = np.array([0.0] + list(ode.ALPHA))
    a = np.zeros((7, 7))
    for i, row in enumerate(ode.BETA):
        a[i + 1, :len(row)] = row
    return c, a


def test_embedded_pair_satisfies_order_conditions():
    """c_sol is 5th order, c_sol - c_error (Shampine's embedded weights) is 4th order."""
    c, a = _full_tableau()
    b5 = np.array(ode.C_SOL)
    b4 = b5 - np.array(ode.C_ERROR)
    ac = a @ c
    conds4 = [(lambda b: b.sum(), 1), (lambda b: b @ c, 1 / 2), (lambda b: b @ c**2, 1 / 3), (lambda b: b @ ac, 1 / 6),
              (lambda b: b @ c**3, 1 / 4), (lambda b: (b * c) @ ac, 1 / 8), (lambda b: b @ (a @ c**2), 1 / 12),
              (lambda b: b @ (a @ ac), 1 / 24)]
    for fn, val in conds4:
        assert abs(fn(b5) - val) < 1e-14
        assert abs(fn(b4) - val) < 1e-12
    conds5 = [(lambda b: b @ c**4, 1 / 5), (lambda b: (b * c**2) @ ac, 1 / 10), (lambda b: (b * c) @ (a @ c**2), 1 / 15),
              (lambda b: (b * c) @ (a @ ac), 1 / 30), (lambda b: b @ (ac * ac), 1 / 20), (lambda b: b @ (a @ c**3), 1 / 20),
              (lambda b: b @ (a @ (c * ac)), 1 / 40), (lambda b: b @ (a @ (a @ c**2)), 1 / 60),
              (lambda b: b @ (a @ (a @ ac)), 1 / 120)]
    for fn, val in conds5:
        assert abs(fn(b5) - val) < 1e-14
    assert abs(sum(ode.C_ERROR)) < 1e-15
    assert any(abs(fn(b4) - val) > 1e-6 for fn, val in conds5)      # the embedded solution really is of lower order


def test_midpoint_weights_are_fourth_order_quadrature():
    c, _ = _full_tableau()
    m = np.array(ode.C_MID)
    for k in range(4):                                              # h * sum m_i f(c_i h) == int_0^{h/2} t^k dt
        assert abs(m @ c**k - 0.5 ** (k + 1) / (k + 1)) < 1e-12


def test_dense_output_meets_its_defining_conditions():
    g = torch.Generator().manual_seed(0)
    y0, y1, ym, f0, f1 = (torch.randn(3, 5, generator=g, dtype=torch.float64) for _ in range(5))
    h = 0.37
    co = ode.interp_fit(y0, y1, ym, f0, f1, h)
    p = lambda x: ode.interp_evaluate(co, 2.0, 2.0 + h, 2.0 + x * h)
    assert torch.allclose(p(0.0), y0, atol=1e-12) and torch.allclose(p(1.0), y1, atol=1e-12)
    assert torch.allclose(p(0.5), ym, atol=1e-12)
    e = 1e-6
    assert torch.allclose((p(e) - p(-e)) / (2 * e * h), f0, atol=1e-6)
    assert torch.allclose((p(1 + e) - p(1 - e)) / (2 * e * h), f1, atol=1e-6)
    assert torch.allclose(O._dp_interp(y0, y1, ym, f0, f1, h, 0.3), p(0.3), atol=1e-12)


def test_step_controller():
    assert ode.optimal_step_size(0.1, 0.0) == pytest.approx(1.0)               # ifactor 10
    assert ode.optimal_step_size(0.1, 1e-12) == pytest.approx(1.0)             # capped at 10x
    assert ode.optimal_step_size(0.1, 0.5) == pytest.approx(0.1 * 0.9 / 0.5 ** 0.2)
    assert ode.optimal_step_size(0.1, 0.9) == pytest.approx(0.1)               # accepted: never shrinks (dfactor -> 1)
    assert ode.optimal_step_size(0.1, 2.0) == pytest.approx(0.1 * 0.9 / 2.0 ** 0.2)
    assert ode.optimal_step_size(0.1, 1e9) == pytest.approx(0.02)              # rejected: at most 5x smaller


def _rotation():
    A = torch.tensor([[0.0, 1.0], [-1.0, 0.0]], dtype=torch.float64)
    return (lambda t, y: (1 + t) * (y @ A.T)), torch.tensor([[1.0, 0.0], [0.0, 2.0]], dtype=torch.float64)


def test_solves_analytic_problem_within_tolerance():
    f, y0 = _rotation()
    th = 1.5                                                        # int_0^1 (1 + t) dt
    exact = torch.tensor([[math.cos(th), -math.sin(th)], [2 * math.sin(th), 2 * math.cos(th)]], dtype=torch.float64)
    errs = []
    for rtol in (1e-3, 1e-5, 1e-7):
        st = {}
        y = ode.dopri5_integrate(f, y0, torch.linspace(0, 1, 50).tolist(), rtol=rtol, atol=rtol * 1e-3, stats=st)
        errs.append(float((y - exact).abs().max()))
        assert st["nfe"] == 2 + 6 * (st["accepted"] + st["rejected"])           # f0 + initial-step probe + 6 per step
        assert errs[-1] < 20 * rtol
    assert errs[0] > errs[1] > errs[2]


def test_dense_outputs_and_stepping_past_the_end():
    f, y0 = _rotation()
    seen = []

    def g(t, y):
        seen.append(t)
        return f(t, y)
    grid = torch.linspace(0, 1, 50).tolist()
    ys = ode.dopri5_integrate(g, y0, grid, last_only=False)
    assert ys.shape == (50, 2, 2) and torch.equal(ys[0], y0)
    for i in (7, 23, 49):
        th = grid[i] + grid[i] ** 2 / 2
        assert abs(float(ys[i, 0, 0]) - math.cos(th)) < 5e-3
    assert max(seen) > 1.0          # like torchdiffeq: no clipping to the end time, the last step overshoots t = 1
    assert torch.equal(ys[-1], ode.dopri5_integrate(f, y0, grid))


def test_product_and_oracle_agree_and_replay():
    f, y0 = _rotation()
    st = {}
    yp = ode.dopri5_integrate(f, y0, torch.linspace(0, 1, 50).tolist(), stats=st)
    yo, steps, nfe = O.dopri5_solve(f, y0, 1.0)
    assert nfe == st["nfe"] and len(steps) == st["accepted"]
    assert np.allclose(np.array(steps), np.array(st["steps"]), rtol=1e-12)
    assert float((yp - yo).abs().max()) < 1e-9
    assert float((O.dopri5_replay(f, y0, st["steps"], 1.0) - yp).abs().max()) < 1e-12


def test_sampler_surface_runs_dopri5_around_forward_inference():
    """Sampler.sample_ode('dopri5') keeps the reference's calling convention (transport.py:408-451) and
    calls forward_inference with t * ones(B) (integrators.py:98-101)."""
    from mdgen_b200.transport import Sampler

    class Owner:
        def __init__(self):
            self.calls = []

        def sample_euler(self, *a, **k):
            raise AssertionError("not the Euler path")

        def forward_inference(self, x, t, mask=None):
            assert t.shape == (x.shape[0],) and t.dtype == torch.float32
            self.calls.append(float(t[0]))
            return -x * mask

    own = Owner()
    sampler = Sampler()
    fn = sampler.sample_ode(sampling_method="dopri5")               # rtol 1e-3, atol 1e-6: transport.py:411-414
    x0 = torch.ones(2, 3, 4, 5)
    out = fn(x0, functools.partial(own.forward_inference, mask=torch.ones(())))[-1]
    assert torch.allclose(out, x0 * math.exp(-1.0), rtol=2e-3)
    assert sampler.last_stats["nfe"] == len(own.calls) == 2 + 6 * (sampler.last_stats["accepted"] + sampler.last_stats["rejected"])
    with pytest.raises(NotImplementedError):
        sampler.sample_ode(sampling_method="heun")
