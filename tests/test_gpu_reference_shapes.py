"""GPU parity at BASELINE.json's own shapes, where the CPU oracle is too slow to be the checker: the checker
runs ON THE GPU in plain fp32 (allow_tf32 off) beside the CUDA path.

Checker = the UNMODIFIED reference (`oracle/_ref`, the sha256-pinned copy made by oracle/vendor_reference.py:
its own `prep_batch`, `forward_inference` and `sample_ode('euler')`), or - when that copy is absent - the
device-agnostic oracle port (`oracle/mdgen_oracle.py`, itself pinned to the reference's golden vectors).

  c2 shape  T = 1000 frames x crop 4 (11 key tiles with a 41-key tail, 8 query tiles per sequence): forward
            velocity at two times and a 2-step Euler state on `euler_time_grid(2)` (dt = 0.5, full-size moves)
  c5 shape  T = 250 x crop 256 with 16 padded residues: forward velocity (both attentions on tcgen05, key padding)
  stress    the same c2 shape with the "trained-like" stress weights (sharp softmax rows, O(1) gates)
Tolerance: north_star, 1e-3 relative fp32.
"""
import pytest
import torch

from mdgen_b200.config import config_from_args, default_args
from mdgen_b200.synthetic import (euler_time_grid, synthetic_batch, synthetic_noise,
                                  synthetic_state_dict)
from tests.helpers import max_rel, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _ours(args, sd, attn_variant=None):
    from mdgen_b200.wrapper import NewMDGenWrapper
    m = NewMDGenWrapper(args)
    m.model.load_state_dict(sd)
    m = m.eval().to("cuda")
    if attn_variant is not None:
        m.model.engine().set_option("attn_variant", attn_variant)
    return m


class _Checker:
    """forward(x, t) / euler(zs, grid) of the reference (or the oracle port) on the GPU in strict fp32."""

    def __init__(self, args_kw, sd, batch):
        from oracle import ref_loader
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.kind = "reference" if ref_loader.reference_available() else "port"
        if self.kind == "reference":
            from functools import partial
            args = ref_loader.make_args(**args_kw)
            m = ref_loader.reference_wrapper(args, sd, "cpu")
            with torch.no_grad():
                prep = m.prep_batch(batch)                      # the reference's own featurisation (CPU LAPACK eigh)
            self.latents = prep["latents"]
            kw = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in prep["model_kwargs"].items()}
            self.m = m.cuda()
            self.f = partial(self.m.model.forward_inference, **kw)
        else:
            from oracle import mdgen_oracle as O
            self.O = O
            self.cfg = config_from_args(default_args(**args_kw))
            op = O.prep_batch(self.cfg, batch)
            self.latents = op["latents"]
            self.sd = {k: v.cuda() for k, v in sd.items()}
            self.kw = dict(mask=op["mask"].cuda(), start=tuple(t.cuda() for t in op["start"]),
                           end=tuple(t.cuda() for t in op["end"]), x_cond=op["x_cond"].cuda(),
                           x_cond_mask=op["x_cond_mask"].cuda(), aatype=op["aatype"].cuda())

    @torch.no_grad()
    def forward(self, x, t):
        if self.kind == "reference":
            return self.f(x, t)
        return self.O.forward(self.sd, self.cfg, x, t, **self.kw)

    @torch.no_grad()
    def euler(self, zs, K):
        if self.kind == "reference":
            fn = self.m.transport_sampler.sample_ode(sampling_method="euler", num_steps=K + 1)
            return fn(zs, self.f)[-1]
        return self.O.sample_euler(self.sd, self.cfg, zs, euler_time_grid(K).cuda(), **self.kw)


def _case(B, T, L, stress=False, pad_last=0, seed=21):
    args_kw = dict(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T,
                   sampling_method="euler")
    cfg = config_from_args(default_args(**args_kw))
    sd = synthetic_state_dict(cfg, seed=0, stress=stress)
    batch = synthetic_batch(B, T, L, seed=seed, vary_frames=True, pad_last=pad_last)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=seed + 1)
    return args_kw, cfg, sd, batch, zs


@pytest.mark.parametrize("stress", [False, True], ids=["bench_weights", "stress_weights"])
def test_c2_shape_forward_and_euler_match_reference_on_gpu(stress):
    """BASELINE configs[1] sequence shape (T = 1000, crop 4) at B = 2."""
    B, T, L = 2, 1000, 4
    args_kw, cfg, sd, batch, zs = _case(B, T, L, stress=stress)
    chk = _Checker(args_kw, sd, batch)
    m = _ours(default_args(**args_kw), sd)
    db = {k: v.cuda() for k, v in batch.items()}
    prep = m.prep_batch(db)
    kw = prep["model_kwargs"]
    assert max_rel(prep["latents"].cpu(), chk.latents) < 2e-5
    zs = zs.cuda()
    t = torch.tensor([0.15, 0.8]).cuda()
    v = m.model.forward_inference(zs, t, **kw)
    vr = chk.forward(zs, t)
    assert max_rel(v.cpu(), vr.cpu()) < TOL, (chk.kind, max_rel(v.cpu(), vr.cpu()))
    assert rel_l2(v.cpu(), vr.cpu()) < TOL
    # two Euler steps with dt = 0.5: the state moves by O(|v|), so velocity errors are not hidden by a small dt
    x = m.model.sample_euler(zs, euler_time_grid(2), **kw)
    xr = chk.euler(zs, 2)
    assert float((x - zs).abs().max()) > 0.1 * float(vr.abs().max())
    assert max_rel(x.cpu(), xr.cpu()) < TOL, (chk.kind, max_rel(x.cpu(), xr.cpu()))
    assert rel_l2((x - zs).cpu(), (xr - zs).cpu()) < 2 * TOL           # the displacement itself
    engine_launches = m.model.engine().launch_count
    assert engine_launches > 0


def test_c5_shape_forward_matches_reference_on_gpu():
    """BASELINE configs[4] per-GPU shape: ATLAS forward-sim T = 250, crop 256 (64,000 tokens), 16 padded residues."""
    B, T, L = 1, 250, 256
    args_kw, cfg, sd, batch, zs = _case(B, T, L, pad_last=16, seed=31)
    chk = _Checker(args_kw, sd, batch)
    m = _ours(default_args(**args_kw), sd)
    kw = m.prep_batch({k: v.cuda() for k, v in batch.items()})["model_kwargs"]
    zs = zs.cuda()
    t = torch.tensor([0.6]).cuda()
    v = m.model.forward_inference(zs, t, **kw)
    vr = chk.forward(zs, t)
    real = batch["mask"][0].bool()
    assert max_rel(v[:, :, real].cpu(), vr[:, :, real].cpu()) < TOL, (chk.kind, max_rel(v.cpu(), vr.cpu()))
    assert rel_l2(v[:, :, real].cpu(), vr[:, :, real].cpu()) < TOL


def test_weight_reload_does_not_leak_device_memory():
    """mdgen_finalize_weights releases the previous generation of packed weights: reloading the
    state dict ten times must not grow the device footprint (one generation is ~0.45 GB)."""
    args_kw, cfg, sd, batch, zs = _case(1, 16, 4)
    m = _ours(default_args(**args_kw), sd)
    eng = m.model.engine()
    dsd = {k: v.cuda() for k, v in sd.items()}
    eng.load_state_dict(dsd)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(10):
        eng.load_state_dict(dsd)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, (free0 - free1) / 2 ** 20
    kw = m.prep_batch({k: v.cuda() for k, v in batch.items()})["model_kwargs"]
    v = m.model.forward_inference(zs.cuda(), torch.tensor([0.5]).cuda(), **kw)
    assert torch.isfinite(v).all()
