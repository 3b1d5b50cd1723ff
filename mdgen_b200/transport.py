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
            eng = owner.engine() if (hasattr(owner, "engine") and x.is_cuda) else None
            if eng is None:          # (host-side tests of the sampler surface drive a plain callable)
                return _LastOnly(dopri5_integrate(rhs, x, t_grid.tolist(), rtol=rtol, atol=atol,
                                                  last_only=True, stats=self.last_stats))
            first = [True]

            def rhs_cached(t, y):
                # the conditioning embedding does not depend on (t, y): built by the first stage, kept by the others
                out = rhs(t, y)
                if first[0]:
                    first[0] = False
                    eng.set_option("reuse_cond", 1)
                return out

            try:
                with torch.cuda.device(x.device):
                    return _LastOnly(dopri5_integrate(rhs_cached, x, t_grid.tolist(), rtol=rtol, atol=atol,
                                                      last_only=True, stats=self.last_stats, engine=eng))
            finally:
                eng.set_option("reuse_cond", 0)

        return sample


class Transport:
    """Flow-matching transport of mdgen/transport/transport.py:77-258 for the one configuration the reference trains
    (velocity prediction, GVP or Linear interpolant => train_eps = sample_eps = 0, transport.py:95-124).
    `training_losses` is the FORWARD half of the training step (what `validation_step` and the loss of
    `training_step` need): noise / time sampling in host code (same RNG calls as the reference, transport.py:126-136),
    interpolant plan and masked MSE as CUDA kernels behind the C ABI (mdgen_flow_plan, mdgen_masked_mse), the
    denoiser through mdgen_forward. The backward pass is a 'next' row."""

    def __init__(self, args=None, path_type="GVP"):
        self.args = args
        self.path_type = path_type
        self.train_eps = 0
        self.sample_eps = 0

    def sample(self, x1):
        """== transport.py:126-136: x0 ~ N(0, I) like x1, t ~ U(0, 1) per sample (drawn on the host, then moved)."""
        x0 = torch.randn_like(x1)
        t = torch.rand((x1.shape[0],)).to(x1)
        return t, x0, x1

    def training_losses(self, model, x1, aatype1=None, mask=None, model_kwargs=None, t=None, x0=None):
        """== transport.py:138-223 (non-design): returns {'t', 'pred', 'loss'} with loss [B] = mean_flat((v - ut)^2, mask).
        `t` / `x0` are optional overrides of the random draws (parity tests)."""
        if getattr(self.args, "design", False):
            raise NotImplementedError("mdgen_b200: design-mode Dirichlet flow matching is a 'next' row (SURVEY.md §8f-4)")
        model_kwargs = model_kwargs or {}
        if t is None or x0 is None:
            t_s, x0_s, _ = self.sample(x1)
            t = t_s if t is None else t
            x0 = x0_s if x0 is None else x0
        eng = model.engine()
        with torch.cuda.device(x1.device):
            xt, ut = eng.flow_plan(x1, x0, t, self.path_type)                    # path.py:132-135
            pred = model(xt, t, **model_kwargs)                                  # transport.py:174
            loss = eng.masked_mse(pred, ut, mask if mask is not None else torch.ones_like(pred))   # :189
        return {"t": t, "pred": pred, "loss": loss}


def create_transport(args, path_type="GVP", prediction="velocity", loss_weight=None,
                     train_eps=None, sample_eps=None):
    if prediction != "velocity" or path_type not in ("GVP", "Linear"):
        raise NotImplementedError("only velocity prediction with GVP/Linear paths is supported")
    return Transport(args, path_type)
