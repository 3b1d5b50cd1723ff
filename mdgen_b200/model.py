"""`LatentMDGenModel` — same constructor, parameter names and `forward_inference` signature as
the reference module (mdgen/model/latent_model.py:43-317), but the arithmetic runs in
libmdgen_b200 (hand-written sm_100a CUDA behind the C ABI). The nn.Module only *holds* the
parameters so checkpoints (`state_dict` keys) load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import Engine, MDGenError
from .config import (MDGenConfig, config_from_args, is_buffer, model_schema)
from .rigid import as_rot_trans, eigh_quat_sign
from .synthetic import sincos_pos_embed


class _Holder(nn.Module):
    """Parameter container addressed by dotted names."""


def _assign(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            setattr(mod, p, _Holder())
        mod = getattr(mod, p)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class LatentMDGenModel(nn.Module):
    def __init__(self, args, latent_dim):
        super().__init__()
        self.args = args
        self.cfg: MDGenConfig = config_from_args(args)
        if latent_dim != self.cfg.latent_dim:
            raise NotImplementedError(f"latent_dim {latent_dim} unsupported")
        for name, shape in model_schema(self.cfg).items():
            if name == "pos_embed":
                t = torch.from_numpy(sincos_pos_embed(shape[-1], shape[1]))[None]
            elif name.endswith("inv_freq"):
                t = 1.0 / (10000 ** (torch.arange(0, 24, 2).float() / 24))
            else:
                t = torch.zeros(shape)
            _assign(self, name, t, is_buffer(name))
        self._engine = None
        self._engine_key = None
        # Two-trunk (tps / inpainting) models only: sign convention of the relative start<->end quaternions.
        #   "canonical": w >= 0 (deterministic; default)
        #   "eigh" / "eigh_cpu": whatever torch.linalg.eigh returns on the frames' device / on the CPU, i.e. what
        #   the reference feeds latent_to_emb_f/r (latent_model.py:194-195) on that backend - use the one the
        #   checkpoint was trained with ("Eigenvector sign").
        self.quat_sign_mode = "canonical"

    # -- engine management ---------------------------------------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict().values())

    def engine(self) -> Engine:
        """Creates the device handle on first use and (re)packs weights when they changed."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise MDGenError("LatentMDGenModel must live on a CUDA device (.to('cuda')); "
                             "mdgen_b200 has no CPU fallback")
        with torch.cuda.device(dev):
            if self._engine is None:
                self._engine = Engine(self.cfg)
            key = self._weights_key()
            if key != self._engine_key:
                self._engine.load_state_dict(self.state_dict())
                self._engine_key = key
        return self._engine

    def _quat_sign(self, start, end):
        if torch.is_tensor(self.quat_sign_mode):          # explicit signs [2,B,L] (tests: the golden's own signs)
            return self.quat_sign_mode.to(start[0].device) if self.cfg.two_trunks else None
        if not self.cfg.two_trunks or self.quat_sign_mode == "canonical" or end is None:
            return None
        if self.quat_sign_mode not in ("eigh", "eigh_cpu"):
            raise ValueError(f"quat_sign_mode {self.quat_sign_mode!r}")
        return eigh_quat_sign(start, end, device="cpu" if self.quat_sign_mode == "eigh_cpu" else None)

    # -- reference surface ---------------------------------------------------------------------
    @torch.no_grad()
    def forward_inference(self, x, t, mask, start_frames=None, end_frames=None, x_cond=None,
                          x_cond_mask=None, aatype=None):
        """== mdgen/model/latent_model.py:263-269 (non-design)."""
        eng = self.engine()
        start, end = as_rot_trans(start_frames), as_rot_trans(end_frames)
        with torch.cuda.device(x.device):
            return eng.forward(x, t, mask, start, end, x_cond, x_cond_mask, aatype,
                               quat_sign=self._quat_sign(start, end))

    def forward(self, *a, **k):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "mdgen_b200: the training (backward) path is a 'next' row of SURVEY.md §8f; "
                "only the no-grad sampling path is implemented")
        return self.forward_inference(*a, **k)

    @torch.no_grad()
    def sample_euler(self, zs, t_grid, mask, start_frames=None, end_frames=None, x_cond=None,
                     x_cond_mask=None, aatype=None):
        """All Euler steps inside the native library (no Python per step)."""
        eng = self.engine()
        start, end = as_rot_trans(start_frames), as_rot_trans(end_frames)
        with torch.cuda.device(zs.device):
            return eng.sample_euler(zs, t_grid, mask, start, end, x_cond, x_cond_mask, aatype,
                                    quat_sign=self._quat_sign(start, end))
