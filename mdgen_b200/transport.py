"""Sampler surface of mdgen/transport/transport.py:279-451 for the one configuration the hot
path uses (velocity prediction, GVP/Linear path => t in [0,1], mdgen/transport/transport.py:95-124,
:560-566). `sample_ode('euler', num_steps=K+1)` returns a callable with the reference's calling
convention `sample_fn(zs, model_fn, **model_kwargs)`; when `model_fn` is (a functools.partial of)
our `LatentMDGenModel.forward_inference`, the whole fixed-grid Euler loop
(mdgen/transport/integrators.py:90-113) runs natively in libmdgen_b200 and only the final state
is produced (the reference stacks every step and the caller takes [-1], wrapper.py:444-447).
`sample_ode('dopri5')` - the reference's default - drives the same fused forward (mdgen_forward) from the
adaptive Dormand-Prince integrator of mdgen_b200/ode.py."""
from __future__ import annotations

import functools

import torch


class _LastOnly:
    """Indexable like the reference's stacked samples; only [-1] exists."""

    def __init__(self, last):
        self._last = last

    def __getitem__(self, i):
        if i != -1:
            raise IndexError("mdgen_b200 sampler keeps only the final ODE state (index -1)")
        return self._last


class Sampler:
    def __init__(self, transport=None):
        self.transport = transport
        self.last_stats = {}      # nfe / accepted / rejected / steps of the last dopri5 call

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3,
                   reverse=False):
        if sampling_method not in ("euler", "dopri5"):
            raise NotImplementedError(
                f"mdgen_b200: sampling_method {sampling_method!r} is not implemented (euler: native loop, "
                "dopri5: adaptive integrator around the native forward)")
        if reverse:
            raise NotImplementedError("reverse-time sampling is not used by the hot path")
        t_grid = torch.linspace(0.0, 1.0, num_steps)          # integrators.py:90 (t0=0, t1=1)

        def sample(x, model, **model_kwargs):
            fn, kw = model, dict(model_kwargs)
            if isinstance(model, functools.partial):
                fn = model.func
                kw = {**model.keywords, **kw}
            owner = getattr(fn, "__self__", None)
            if owner is None or not hasattr(owner, "sample_euler"):
                raise NotImplementedError("sample_ode needs LatentMDGenModel.forward_inference")
            if sampling_method == "euler":
                return _LastOnly(owner.sample_euler(x, t_grid, **kw))
            from .ode import dopri5_integrate

            def rhs(t, y):                                     # integrators.py:98-101: t * ones(B)
                tb = torch.full((y.shape[0],), float(t), dtype=torch.float32, device=y.device)
                return owner.forward_inference(y, tb, **kw)

            self.last_stats = {}
            return _LastOnly(dopri5_integrate(rhs, x, t_grid.tolist(), rtol=rtol, atol=atol,
                                              last_only=True, stats=self.last_stats))

        return sample


class Transport:
    """Placeholder for `create_transport` (training losses are a 'next' row)."""

    def __init__(self, args=None):
        self.args = args
        self.train_eps = 0
        self.sample_eps = 0

    def training_losses(self, *a, **k):
        raise NotImplementedError("mdgen_b200: training path not implemented (SURVEY.md §8f-3)")


def create_transport(args, path_type="GVP", prediction="velocity", loss_weight=None,
                     train_eps=None, sample_eps=None):
    if prediction != "velocity" or path_type not in ("GVP", "Linear"):
        raise NotImplementedError("only velocity prediction with GVP/Linear paths is supported")
    return Transport(args)
