"""Generate tests/golden/train_loss.npz by running the UNMODIFIED reference's validation / training-loss path.

Run HERE (build container): `python tests/golden/gen_train_golden.py`. For every golden case the reference's own
`NewMDGenWrapper.general_step(batch, stage='val')` (mdgen/wrapper.py:367-403 -> Transport.training_losses,
mdgen/transport/transport.py:138-223) is executed under torch.manual_seed(7); the random draws it makes
(`th.randn_like(x1)`, `th.rand((B,))`, transport.py:126-136) are recorded by wrapping the two torch functions, so the
parity tests can feed identical (t, x0) to the oracle and to the CUDA path regardless of the device's RNG stream.
Stored per case: t [B], x0 [B,T,L,D], per-sample loss [B] (out_dict['loss']) and the value general_step returns.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference, make_args  # noqa: E402
from mdgen_b200.config import config_from_args  # noqa: E402
from mdgen_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402
from tests.golden.cases import CASES  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TRAIN_CASES = ["sim_c1", "atlas_small", "upsampling", "tps", "inpaint", "stress"]


def main():
    torch.set_num_threads(8)
    wrapper_mod = load_reference()
    import mdgen.rigid_utils as ru
    orig_r2q = ru.rot_to_quat

    def canonical_r2q(rot):
        q = orig_r2q(rot)
        return q * torch.where(q[..., 0:1] < 0, -1.0, 1.0)

    out = {}
    for name in TRAIN_CASES:
        case = CASES[name]
        args = make_args(**case["args"])
        cfg = config_from_args(args)
        ru.rot_to_quat = canonical_r2q if case.get("canonical_quat") else orig_r2q
        torch.manual_seed(0)
        m = wrapper_mod.NewMDGenWrapper(args).eval()
        m.model.load_state_dict(synthetic_state_dict(cfg, seed=0, stress=bool(case.get("stress"))), strict=True)
        batch = synthetic_batch(case["B"], case["T"], case["L"], seed=1, **case.get("batch", {}))
        rec = {}
        real_randn_like, real_rand = torch.randn_like, torch.rand

        def randn_like(x, *a, **k):
            r = real_randn_like(x, *a, **k)
            rec["x0"] = r.clone()
            return r

        def rand(*a, **k):
            r = real_rand(*a, **k)
            rec["t"] = r.clone()
            return r

        captured = {}
        orig_tl = m.transport.training_losses

        def training_losses(*a, **k):
            o = orig_tl(*a, **k)
            captured["loss"] = o["loss"].detach().clone()
            return o

        m.transport.training_losses = training_losses
        torch.manual_seed(7)
        torch.randn_like, torch.rand = randn_like, rand
        try:
            with torch.no_grad():
                mean_loss = m.general_step(batch, stage="val")
        finally:
            torch.randn_like, torch.rand = real_randn_like, real_rand
        out[f"{name}/t"] = rec["t"].float().numpy()
        out[f"{name}/x0"] = rec["x0"].numpy()
        out[f"{name}/loss"] = captured["loss"].numpy()
        out[f"{name}/loss_mean"] = np.float32(mean_loss.item())
        print(name, "t", rec["t"].tolist(), "loss", captured["loss"].tolist(), "mean", float(mean_loss))
    ru.rot_to_quat = orig_r2q
    np.savez_compressed(os.path.join(OUT, "train_loss.npz"), **out)


if __name__ == "__main__":
    main()
