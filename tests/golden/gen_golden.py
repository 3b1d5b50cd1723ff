"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run HERE (build container): `python tests/golden/gen_golden.py`. The reference is imported
through oracle/ref_loader.py (stand-in third-party modules only), the deterministic synthetic
weights of mdgen_b200.synthetic are loaded into the reference's own `NewMDGenWrapper`
(`load_state_dict(strict=True)` — which also pins the state-dict schema), and the reference's own
public entry points are called:

  prep      NewMDGenWrapper.prep_batch                      (mdgen/wrapper.py:283-365)
  forward   model.forward_inference(x, t, **model_kwargs)    (mdgen/model/latent_model.py:263-269)
  euler     transport_sampler.sample_ode('euler', K+1)(zs, f)[-1]
                                                            (mdgen/transport/transport.py:408-451)
  infer     NewMDGenWrapper.inference(batch) with torch.randn patched to return the seeded zs
            and args.sampling_method='euler' (49 steps: wrapper.py:441-447, D3 in SURVEY.md)

Two-trunk cases (tps, inpaint): the reference leaves LAPACK's arbitrary eigenvector sign on the relative
quaternions of run_ipa's tps branch (latent_model.py:194-195), which makes its output backend dependent.
Both conventions are recorded:
  v, x_euler, x49, atom14      with mdgen.rigid_utils.rot_to_quat wrapped so that the sign is canonical
                               (w >= 0): the product's default `quat_sign_mode="canonical"`
  v_eigh, x_euler_eigh         from the UNPATCHED reference (this container's CPU LAPACK), plus
  quat_sign [2,B,L]            the signs it used (row 0: end^-1 o start -> latent_to_emb_r, row 1: start^-1 o end
                               -> latent_to_emb_f), read off the reference's own to_tensor_7 outputs: the
                               product's `quat_sign_mode="eigh"` fed with these signs must reproduce v_eigh.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference, make_args  # noqa: E402
from mdgen_b200.config import config_from_args  # noqa: E402
from mdgen_b200.synthetic import (euler_time_grid, synthetic_batch, synthetic_noise,  # noqa: E402
                                  synthetic_state_dict)
from tests.golden.cases import CASES  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(8)
    wrapper_mod = load_reference()
    import mdgen.rigid_utils as ru
    orig_r2q = ru.rot_to_quat

    def canonical_r2q(rot):
        q = orig_r2q(rot)
        return q * torch.where(q[..., 0:1] < 0, -1.0, 1.0)

    for name, case in CASES.items():
        args = make_args(**case["args"])
        cfg = config_from_args(args)
        ru.rot_to_quat = canonical_r2q if case.get("canonical_quat") else orig_r2q
        torch.manual_seed(0)
        m = wrapper_mod.NewMDGenWrapper(args).eval()
        sd = synthetic_state_dict(cfg, seed=0, stress=bool(case.get("stress")))
        missing = m.model.load_state_dict(sd, strict=True)
        B, T, L, K = case["B"], case["T"], case["L"], case["K"]
        batch = synthetic_batch(B, T, L, seed=1, **case.get("batch", {}))
        zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=2)
        t_fwd = torch.tensor(case["t_fwd"], dtype=torch.float32)
        with torch.no_grad():
            prep = m.prep_batch(batch)
            kw = prep["model_kwargs"]
            v = m.model.forward_inference(zs, t_fwd, **kw)
            f = partial(m.model.forward_inference, **kw)
            xk = m.transport_sampler.sample_ode(sampling_method="euler", num_steps=K + 1)(zs, f)[-1]
            # public API: inference() with the seeded noise (49 Euler steps)
            args.sampling_method = "euler"
            real_randn = torch.randn
            torch.randn = lambda *a, **k: zs.clone()
            try:
                atom14, aa_out = m.inference(batch)
            finally:
                torch.randn = real_randn
            x49 = m.transport_sampler.sample_ode(sampling_method="euler", num_steps=50)(zs, f)[-1]
        extra = {}
        if case.get("canonical_quat"):
            ru.rot_to_quat = orig_r2q                      # the unmodified reference
            with torch.no_grad():
                s_f, e_f = kw["start_frames"], kw["end_frames"]
                q_r = e_f.invert().compose(s_f).to_tensor_7()[..., 0]
                q_f = s_f.invert().compose(e_f).to_tensor_7()[..., 0]
                extra["quat_sign"] = torch.stack([torch.where(q_r < 0, -1.0, 1.0),
                                                  torch.where(q_f < 0, -1.0, 1.0)]).float().numpy()
                extra["v_eigh"] = m.model.forward_inference(zs, t_fwd, **kw).numpy()
                extra["x_euler_eigh"] = m.transport_sampler.sample_ode(
                    sampling_method="euler", num_steps=K + 1)(zs, f)[-1].numpy()
            print(name, "unpatched reference: negative-w fraction",
                  float((extra["quat_sign"] < 0).mean()), "max |v_eigh - v| / max|v|",
                  float(np.abs(extra["v_eigh"] - v.numpy()).max() / np.abs(v.numpy()).max()))
        wsum = float(sum(t.double().abs().sum() for t in sd.values()))
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            latents=prep["latents"].numpy(), x_cond=kw["x_cond"].numpy(),
            x_cond_mask=kw["x_cond_mask"].numpy(), v=v.numpy(), x_euler=xk.numpy(),
            x49=x49.numpy(), atom14=atom14.numpy(), aa_out=aa_out.numpy(),
            weight_abs_sum=np.float64(wsum), zs_abs_sum=np.float64(zs.double().abs().sum()), **extra,
        )
        print(name, "latent_dim", cfg.latent_dim, "v rms", float(v.pow(2).mean().sqrt()),
              "x_euler rms", float(xk.pow(2).mean().sqrt()), "atom14", tuple(atom14.shape),
              "finite", bool(torch.isfinite(atom14).all()))
    ru.rot_to_quat = orig_r2q


if __name__ == "__main__":
    main()
