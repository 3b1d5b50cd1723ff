"""Deterministic synthetic weights and batches (no dataset / checkpoint is available offline).

Everything is drawn from numpy PCG64 streams keyed by (seed, tensor name) so the build
container (golden-vector generation against the live reference) and the GPU box (parity tests,
bench) see identical bits. Shapes follow SURVEY.md §8d:
  weights  seed 0 — every tensor non-zero (the reference's own init zeroes the output layers,
           which would make the network output identically 0: mdgen/model/latent_model.py:140-173)
  batch    seed 1 — unit-quaternion frames, 3.8 Å random-walk CA trace, unit (sin,cos) torsions
  noise    seed 2 — zs ~ N(0,1), time grid linspace(0,1,K+1) float32
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

from .config import EMBED_DIM, HEAD_DIM, MDGenConfig, model_schema


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def sincos_pos_embed(embed_dim: int, n: int) -> np.ndarray:
    """1-D sin/cos table: omega_i = 10000^(-i/(D/2)), [sin | cos], float64 then cast
    (same table as mdgen/model/latent_model.py:22-40,151-153)."""
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    out = np.arange(n, dtype=np.float64)[:, None] * omega[None, :]
    return np.concatenate([np.sin(out), np.cos(out)], axis=1).astype(np.float32)


def synthetic_state_dict(cfg: MDGenConfig, seed: int = 0, stress: bool = False
                         ) -> "OrderedDict[str, torch.Tensor]":
    """Full `LatentMDGenModel` state dict (keys without the wrapper's `model.` prefix).

    stress=True makes the weights "trained-like" in the ways that matter numerically: the query
    projections of every token attention are scaled x6 (logit std ~6: sharp, near one-hot softmax rows),
    the adaLN gates are ~1 instead of ~0.02 (every branch contributes at O(1) to the residual stream)
    and the final projection is xavier-scaled (O(1) velocities)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in model_schema(cfg).items():
        g = _rng(seed, name)
        if name == "pos_embed":
            arr = sincos_pos_embed(EMBED_DIM, cfg.crop)[None]
        elif name.endswith("rot_emb.inv_freq"):
            arr = (1.0 / (10000 ** (np.arange(0, HEAD_DIM, 2, dtype=np.float32) / HEAD_DIM))
                   ).astype(np.float32)
        elif name.endswith("head_weights"):
            arr = (0.5413 + 0.1 * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)
        elif name.endswith("ipa_norm.weight"):
            arr = (1.0 + 0.05 * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)
        elif name.endswith(("bias_k", "bias_v")):
            arr = math.sqrt(2.0 / (2 * EMBED_DIM)) * g.standard_normal(shape, dtype=np.float32)
        elif name in ("mask_to_emb.weight", "aatype_to_emb.weight"):
            arr = g.standard_normal(shape, dtype=np.float32)
        elif name.endswith(".bias"):
            arr = 0.02 * g.standard_normal(shape, dtype=np.float32)
        elif name.startswith("t_embedder") or "adaLN_modulation" in name \
                or name.startswith("emb_to_latent.linear"):
            arr = 0.02 * g.standard_normal(shape, dtype=np.float32)
        else:  # Linear weights [out, in]: xavier-normal scale
            fan_out, fan_in = shape
            arr = math.sqrt(2.0 / (fan_in + fan_out)) * g.standard_normal(shape, dtype=np.float32)
        if stress:
            if name.startswith("layers.") and name.endswith(("attn.q_proj.weight", "attn.q_proj.bias")):
                arr = 6.0 * arr
            elif name.endswith("adaLN_modulation.1.bias"):
                arr = arr.copy()
                nchunk = arr.shape[0] // EMBED_DIM          # 6 (IPA layers) / 9 (main layers) / 2 (final)
                for c in range(2, nchunk, 3):               # gate chunks: every third, starting at index 2
                    arr[c * EMBED_DIM:(c + 1) * EMBED_DIM] += 1.0
            elif name == "emb_to_latent.linear.weight":
                fan_out, fan_in = shape
                arr = math.sqrt(2.0 / (fan_in + fan_out)) * g.standard_normal(shape, dtype=np.float32)
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def _quat_to_rot(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r = np.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z,
    ], axis=-1)
    return r.reshape(q.shape[:-1] + (3, 3))


def synthetic_batch(B: int, T: int, L: int, seed: int = 1, vary_frames: bool = True,
                    pad_last: int = 0, cond_interval: int = 0) -> Dict[str, torch.Tensor]:
    """The batch dict `NewMDGenWrapper.inference` consumes (mdgen/wrapper.py:405; keys built by
    sim_inference.py:52-59 / upsampling_inference.py:53-64 / dataset.py:70-100).

    vary_frames=False reproduces the forward-sim rollout input (frame 0 tiled over T);
    vary_frames=True gives every frame its own pose so prep_batch parity is non-trivial.
    pad_last>0 marks the last residues as ATLAS-style padding (mask 0, identity frames).
    cond_interval>0 blanks non-key frames as upsampling_inference.py:53-64 does.
    """
    g = _rng(seed, f"batch/{B}/{T}/{L}")
    Tq = T if vary_frames else 1
    q = g.standard_normal((B, Tq, L, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    rots = _quat_to_rot(q).astype(np.float32)
    steps = g.standard_normal((B, 1, L, 3)).astype(np.float32)
    steps /= np.linalg.norm(steps, axis=-1, keepdims=True)
    trans = np.cumsum(3.8 * steps, axis=2)
    if vary_frames:
        trans = trans + 0.5 * g.standard_normal((B, T, L, 3)).astype(np.float32)
    tors = g.standard_normal((B, Tq, L, 7, 2)).astype(np.float32)
    tors /= np.linalg.norm(tors, axis=-1, keepdims=True)
    if not vary_frames:
        rots = np.broadcast_to(rots, (B, T, L, 3, 3)).copy()
        trans = np.broadcast_to(trans, (B, T, L, 3)).copy()
        tors = np.broadcast_to(tors, (B, T, L, 7, 2)).copy()
    seqres = g.integers(0, 20, size=(B, L)).astype(np.int64)
    mask = np.ones((B, L), np.float32)
    tmask = np.ones((B, L, 7), np.float32)
    if pad_last:
        mask[:, L - pad_last:] = 0
        tmask[:, L - pad_last:] = 0
        rots[:, :, L - pad_last:] = np.eye(3, dtype=np.float32)
        trans[:, :, L - pad_last:] = 0
        tors[:, :, L - pad_last:] = 0
        seqres[:, L - pad_last:] = 0
    if cond_interval:
        keep = np.zeros(T, bool)
        keep[::cond_interval] = True
        rots[:, ~keep] = np.eye(3, dtype=np.float32)
        trans[:, ~keep] = 0
        tors[:, ~keep] = 0
    return {
        "torsions": torch.from_numpy(np.ascontiguousarray(tors, np.float32)),
        "torsion_mask": torch.from_numpy(tmask),
        "trans": torch.from_numpy(np.ascontiguousarray(trans, np.float32)),
        "rots": torch.from_numpy(np.ascontiguousarray(rots, np.float32)),
        "seqres": torch.from_numpy(seqres),
        "mask": torch.from_numpy(mask),
    }


def synthetic_noise(B: int, T: int, L: int, D: int, seed: int = 2) -> torch.Tensor:
    g = _rng(seed, f"noise/{B}/{T}/{L}/{D}")
    return torch.from_numpy(g.standard_normal((B, T, L, D), dtype=np.float32))


def euler_time_grid(num_steps: int) -> torch.Tensor:
    """float32 linspace(0,1,K+1): the grid mdgen/transport/integrators.py:90 builds for
    sample_ode(num_steps=K+1)."""
    return torch.linspace(0.0, 1.0, num_steps + 1, dtype=torch.float32)
