"""Minimal rigid-frame holder with the accessor names the reference's callers use
(`Rigid.get_trans()`, `Rigid.get_rots().get_rot_mats()`, indexing, `.shape`;
mdgen/rigid_utils.py:813-1391). It carries tensors only — all SE(3) arithmetic of the hot path
runs inside libmdgen_b200 (prep / IPA / decode kernels)."""
from __future__ import annotations

import torch


class Rotation:
    def __init__(self, rot_mats: torch.Tensor):
        self._rot_mats = rot_mats

    def get_rot_mats(self) -> torch.Tensor:
        return self._rot_mats


class Rigid:
    def __init__(self, rots, trans: torch.Tensor):
        self._rots = rots if isinstance(rots, Rotation) else Rotation(rots)
        self._trans = trans

    @property
    def shape(self):
        return self._trans.shape[:-1]

    @property
    def device(self):
        return self._trans.device

    def get_rots(self) -> Rotation:
        return self._rots

    def get_trans(self) -> torch.Tensor:
        return self._trans

    def __getitem__(self, idx) -> "Rigid":
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Rigid(self._rots.get_rot_mats()[idx + (slice(None), slice(None))],
                     self._trans[idx + (slice(None),)])


def as_rot_trans(frames):
    """Accepts this module's Rigid, the reference's Rigid (duck-typed) or a (rot, trans) tuple."""
    if frames is None:
        return None
    if isinstance(frames, (tuple, list)):
        return frames[0], frames[1]
    return frames.get_rots().get_rot_mats(), frames.get_trans()


def eigh_quat_sign(start, end, device=None):
    """Sign (+1 / -1) that `torch.linalg.eigh` leaves on the relative quaternions of the reference's two-trunk
    branch: rot_to_quat(R)[..., 0] for R = end^-1 o start (row 0) and R = start^-1 o end (row 1)
    (mdgen/model/latent_model.py:194-195 -> Rigid.to_tensor_7 -> mdgen/rigid_utils.py:191-210). The sign of an
    eigenvector is a property of the LAPACK / cuSOLVER build, so it can only be reproduced by making the same
    library call; `device` selects where (None: where the frames live, as the reference would). The 4x4 matrices
    are assembled exactly as the reference does; only the sign of the top eigenvector's first component is kept.
    start / end: (rot [B,L,3,3], trans) tuples. Returns float32 [2,B,L] on the frames' device."""
    Rs, Re = start[0], end[0]
    out_dev = Rs.device
    if device is not None:
        Rs, Re = Rs.to(device), Re.to(device)

    def sign_of(rot):
        xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
        yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
        zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
        k = [[xx + yy + zz, zy - yz, xz - zx, yx - xy],
             [zy - yz, xx - yy - zz, xy + yx, xz + zx],
             [xz - zx, xy + yx, yy - xx - zz, yz + zy],
             [yx - xy, xz + zx, yz + zy, zz - xx - yy]]
        k = (1.0 / 3.0) * torch.stack([torch.stack(r, dim=-1) for r in k], dim=-2)
        _, vec = torch.linalg.eigh(k)
        return torch.where(vec[..., 0, -1] < 0, -1.0, 1.0)

    rel_r = Re.transpose(-1, -2) @ Rs        # end^-1 o start   -> latent_to_emb_r
    rel_f = Rs.transpose(-1, -2) @ Re        # start^-1 o end   -> latent_to_emb_f
    return torch.stack([sign_of(rel_r), sign_of(rel_f)]).to(torch.float32).to(out_dev)
