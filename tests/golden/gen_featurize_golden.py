"""Golden vectors for the rollout re-featurisation row, from the UNMODIFIED reference:
frames = mdgen.geometry.atom14_to_frames(atom14[:, -1]) and torsions =
atom37_to_torsions(atom14_to_atom37(...)) exactly as sim_inference.py:91-96 calls them, applied to the
last frame of the reference's own `inference()` outputs stored in the other golden files.
Run HERE: python tests/golden/gen_featurize_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402
from mdgen_b200.synthetic import synthetic_batch  # noqa: E402
from tests.golden.cases import CASES  # noqa: E402

load_reference()
from mdgen.geometry import atom14_to_atom37, atom14_to_frames, atom37_to_torsions  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
out = {}
for name in ("sim_c1", "atlas_small", "tps"):
    case = CASES[name]
    g = np.load(os.path.join(OUT, f"{name}.npz"))
    batch = synthetic_batch(case["B"], case["T"], case["L"], seed=1, **case.get("batch", {}))
    a14 = torch.from_numpy(g["atom14"])[:, -1]
    fr = atom14_to_frames(a14)
    tors, masks = [], []
    for i in range(case["B"]):
        a37 = atom14_to_atom37(a14[i], batch["seqres"][i])
        t, m = atom37_to_torsions(a37, batch["seqres"][i])
        tors.append(t)
        masks.append(m)
    out[f"{name}/rots"] = fr._rots._rot_mats.numpy()
    out[f"{name}/trans"] = fr._trans.numpy()
    out[f"{name}/torsions"] = torch.stack(tors).numpy()
    out[f"{name}/torsion_mask"] = torch.stack(masks).numpy()
np.savez_compressed(os.path.join(OUT, "featurize.npz"), **out)
print({k: v.shape for k, v in out.items()})
