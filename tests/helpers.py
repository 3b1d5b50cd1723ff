"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from mdgen_b200.config import config_from_args, default_args
from mdgen_b200.synthetic import synthetic_batch, synthetic_noise, synthetic_state_dict
from tests.golden.cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    """Returns (case, args, cfg, sd, batch, zs, golden-dict) for a golden case; verifies that the
    synthetic weights / noise regenerated here are bit-identical to those used at generation."""
    case = CASES[name]
    args = default_args(**case["args"])
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0, stress=bool(case.get("stress")))
    batch = synthetic_batch(case["B"], case["T"], case["L"], seed=1, **case.get("batch", {}))
    zs = synthetic_noise(case["B"], case["T"], case["L"], cfg.latent_dim, seed=2)
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    wsum = float(sum(t.double().abs().sum() for t in sd.values()))
    # (summation order of the fp64 reduction depends on the thread count: compare to 1e-12 relative)
    assert abs(wsum - float(g["weight_abs_sum"])) <= 1e-12 * wsum, "synthetic weights differ from golden generation"
    assert abs(float(zs.double().abs().sum()) - float(g["zs_abs_sum"])) <= 1e-12 * float(g["zs_abs_sum"]), "synthetic noise differs"
    return case, args, cfg, sd, batch, zs, g


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    """max |a-b| / max |b| — the 'relative fp32' bound of BASELINE.json's north_star."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
