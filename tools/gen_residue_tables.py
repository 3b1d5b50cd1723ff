"""Generate mdgen_b200/data/residue_tables.npz from the reference's residue constants.

Run HERE (build container; needs /root/reference). The four arrays are AlphaFold's public
stereochemistry tables as materialised by the reference at import time
(mdgen/residue_constants.py:1124-1130 and the fill loop below them); they are *data*, reused
as constants ("data only — reuse as constants"). Both the product (uploaded to
the device by mdgen_create) and the oracle read this file, so the GPU box never needs the
reference.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
import mdgen.residue_constants as rc  # noqa: E402

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "mdgen_b200", "data", "residue_tables.npz")
from mdgen.geometry import get_chi_atom_indices  # noqa: E402

# rollout re-featurisation (atom14 -> torsions, mdgen/geometry.py:82-202): chi atoms as atom14 indices,
# their existence mask (RESTYPE_ATOM37_MASK, residue_constants.py:1477) and chi_angles_mask (:81-102)
chi37 = np.asarray(get_chi_atom_indices(), dtype=np.int64)                      # [21,4,4] atom37 indices
to14 = np.asarray(rc.RESTYPE_ATOM37_TO_ATOM14)                                  # [21,37]
m37 = np.asarray(rc.RESTYPE_ATOM37_MASK, dtype=np.float32)                      # [21,37]
chi14 = np.take_along_axis(to14[:, None, :].repeat(4, 1), chi37, axis=2).astype(np.int32)
chi_atom_mask = np.take_along_axis(m37[:, None, :].repeat(4, 1), chi37, axis=2).astype(np.float32)
chi_mask = np.asarray(list(rc.chi_angles_mask) + [[0.0, 0.0, 0.0, 0.0]], dtype=np.float32)   # [21,4]
bb_mask = m37[:, [0, 1, 2, 4]].astype(np.float32)                               # N, CA, C, O  (atom14 0..3)
assert (to14[:20, [0, 1, 2, 4]] == np.array([0, 1, 2, 3])).all()

np.savez_compressed(
    out,
    chi_atom14_idx=chi14, chi_atom_mask=chi_atom_mask, chi_mask=chi_mask, bb_mask=bb_mask,
    default_frame=np.asarray(rc.restype_rigid_group_default_frame, dtype=np.float32),   # [21,8,4,4]
    atom14_group_pos=np.asarray(rc.restype_atom14_rigid_group_positions, dtype=np.float32),  # [21,14,3]
    atom14_to_group=np.asarray(rc.restype_atom14_to_rigid_group, dtype=np.int32),       # [21,14]
    atom14_mask=np.asarray(rc.restype_atom14_mask, dtype=np.float32),                   # [21,14]
)
print("wrote", out, os.path.getsize(out), "bytes")
