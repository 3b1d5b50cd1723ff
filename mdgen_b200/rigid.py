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
