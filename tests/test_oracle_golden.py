"""Pins the CPU oracle (oracle/mdgen_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/gen_golden.py). CPU only."""
import pytest
import torch

from oracle import mdgen_oracle as O
from mdgen_b200.synthetic import euler_time_grid
from tests.golden.cases import CASES
from tests.helpers import load_case, max_rel, rel_l2

# fp32 re-association noise only (same algorithm, same op order up to einsum/matmul kernels)
TOL_PREP = 2e-5
TOL_FWD = 2e-5
TOL_EULER = 5e-5


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(8)
    case, args, cfg, sd, batch, zs, g = load_case(name)
    prep = O.prep_batch(cfg, batch)
    assert max_rel(prep["latents"], g["latents"]) < TOL_PREP
    assert max_rel(prep["x_cond"], g["x_cond"]) < TOL_PREP
    assert (prep["x_cond_mask"].numpy() == g["x_cond_mask"]).all()
    kw = dict(mask=prep["mask"], start=prep["start"], end=prep["end"], x_cond=prep["x_cond"],
              x_cond_mask=prep["x_cond_mask"], aatype=prep["aatype"])
    with torch.no_grad():
        v = O.forward(sd, cfg, zs, torch.tensor(case["t_fwd"]), **kw)
        assert max_rel(v, g["v"]) < TOL_FWD, max_rel(v, g["v"])
        xk = O.sample_euler(sd, cfg, zs, euler_time_grid(case["K"]), **kw)
        assert max_rel(xk, g["x_euler"]) < TOL_EULER
        assert rel_l2(xk, g["x_euler"]) < TOL_EULER
        # decode tail on the reference's own 49-step state -> reference atom14
        atom14 = O.decode_atom14(cfg, torch.from_numpy(g["x49"]), batch["rots"][:, 0],
                                 batch["trans"][:, 0], batch["seqres"])
        assert max_rel(atom14, g["atom14"]) < 2e-5
    assert (g["aa_out"] == batch["seqres"][:, None].expand(-1, case["T"], -1).numpy()).all()


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c.get("canonical_quat")])
def test_oracle_follows_unpatched_reference_eigh_sign(name):
    """Two-trunk configs: the UNPATCHED reference keeps torch.linalg.eigh's eigenvector sign on the relative
    quaternions (latent_model.py:194-195). With CANONICAL_TPS_QUAT off the oracle makes the same LAPACK call
    and must reproduce the unpatched goldens; the product's sign helper must return the signs the reference
    used. (The canonical goldens of the same case differ from these by up to 86 % of max|v|: the sign matters.)"""
    from mdgen_b200.rigid import eigh_quat_sign
    torch.set_num_threads(8)
    case, args, cfg, sd, batch, zs, g = load_case(name)
    prep = O.prep_batch(cfg, batch)
    sign = eigh_quat_sign(prep["start"], prep["end"], device="cpu")
    assert (sign.numpy() == g["quat_sign"]).all(), "torch.linalg.eigh sign differs from golden generation"
    assert (g["quat_sign"] < 0).any() and (g["quat_sign"] > 0).any()
    kw = dict(mask=prep["mask"], start=prep["start"], end=prep["end"], x_cond=prep["x_cond"],
              x_cond_mask=prep["x_cond_mask"], aatype=prep["aatype"])
    O.CANONICAL_TPS_QUAT = False
    try:
        with torch.no_grad():
            v = O.forward(sd, cfg, zs, torch.tensor(case["t_fwd"]), **kw)
            xk = O.sample_euler(sd, cfg, zs, euler_time_grid(case["K"]), **kw)
    finally:
        O.CANONICAL_TPS_QUAT = True
    assert max_rel(v, g["v_eigh"]) < TOL_FWD, max_rel(v, g["v_eigh"])
    assert max_rel(xk, g["x_euler_eigh"]) < TOL_EULER
    assert max_rel(torch.from_numpy(g["v_eigh"]), g["v"]) > 1e-2      # the two conventions really differ


def test_rope_shim_matches_hf_esm_port():
    """The fair-esm rotary embedding restated in oracle/ref_shims is bit-identical to the
    independent HF transformers port (the only offline cross-check for this un-vendored dep)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "oracle", "ref_shims"))
    try:
        from transformers.models.esm.modeling_esm import RotaryEmbedding as HF
    except Exception as e:  # pragma: no cover
        pytest.skip(f"transformers ESM port unavailable: {e}")
    from esm.rotary_embedding import RotaryEmbedding as Shim
    torch.manual_seed(0)
    q = torch.randn(3, 16, 7, 24)
    k = torch.randn(3, 16, 8, 24)
    hq, hk = HF(24)(q, k)
    sq, sk = Shim(24)(q.reshape(48, 7, 24), k.reshape(48, 8, 24))
    assert torch.equal(hq.reshape(48, 7, 24), sq) and torch.equal(hk.reshape(48, 8, 24), sk)
    # and the oracle's table form
    cos, sin = O.rope_tables(8, Shim(24).inv_freq)
    ok = k * cos + O.rotate_half(k) * sin
    assert torch.allclose(ok, hk, atol=0, rtol=0)


def test_featurize_oracle_matches_reference_golden(golden_dir):
    """Rollout re-featurisation: oracle vs the reference's atom14_to_frames /
    atom37_to_torsions outputs on the last frame of its own inference() results."""
    import os

    import numpy as np
    from mdgen_b200.synthetic import synthetic_batch
    f = dict(np.load(os.path.join(golden_dir, "featurize.npz")))
    for name in ("sim_c1", "atlas_small", "tps"):
        case = CASES[name]
        g = np.load(os.path.join(golden_dir, f"{name}.npz"))
        batch = synthetic_batch(case["B"], case["T"], case["L"], seed=1, **case.get("batch", {}))
        a14 = torch.from_numpy(g["atom14"])[:, -1]
        R, t, sc, m = O.featurize_atom14(a14, batch["seqres"])
        assert (R - torch.from_numpy(f[f"{name}/rots"])).abs().max() < 1e-5
        assert (t - torch.from_numpy(f[f"{name}/trans"])).abs().max() < 1e-6
        assert (sc - torch.from_numpy(f[f"{name}/torsions"])).abs().max() < 2e-5
        assert (m.numpy() == f[f"{name}/torsion_mask"]).all()


TRAIN_CASES = ["sim_c1", "atlas_small", "upsampling", "tps", "inpaint", "stress"]


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_training_loss_matches_reference_general_step(name, golden_dir):
    """Forward half of the training step: the oracle's restatement of Transport.training_losses /
    mean_flat / the GVP plan against the per-sample losses of the reference's own general_step(stage='val')
    (tests/golden/gen_train_golden.py), fed with the (t, x0) the reference drew."""
    import os

    import numpy as np
    torch.set_num_threads(8)
    g = np.load(os.path.join(golden_dir, "train_loss.npz"))
    case, args, cfg, sd, batch, zs, _ = load_case(name)
    prep = O.prep_batch(cfg, batch)
    kw = dict(mask=prep["mask"], start=prep["start"], end=prep["end"], x_cond=prep["x_cond"],
              x_cond_mask=prep["x_cond_mask"], aatype=prep["aatype"])
    with torch.no_grad():
        loss, _ = O.training_losses(sd, cfg, prep["latents"], prep["loss_mask"], torch.from_numpy(g[f"{name}/t"]),
                                    torch.from_numpy(g[f"{name}/x0"]), **kw)
    assert max_rel(loss, g[f"{name}/loss"]) < 5e-5, (loss, g[f"{name}/loss"])
    assert abs(float(loss.mean()) - float(g[f"{name}/loss_mean"])) < 5e-5 * float(g[f"{name}/loss_mean"])
