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
np.savez_compressed(
    out,
    default_frame=np.asarray(rc.restype_rigid_group_default_frame, dtype=np.float32),   # [21,8,4,4]
    atom14_group_pos=np.asarray(rc.restype_atom14_rigid_group_positions, dtype=np.float32),  # [21,14,3]
    atom14_to_group=np.asarray(rc.restype_atom14_to_rigid_group, dtype=np.int32),       # [21,14]
    atom14_mask=np.asarray(rc.restype_atom14_mask, dtype=np.float32),                   # [21,14]
)
print("wrote", out, os.path.getsize(out), "bytes")
