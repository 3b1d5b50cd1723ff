"""Model configuration and state-dict schema of the MDGen denoiser hot path.

Mirrors what the reference derives from its argparse namespace:
  * latent_dim rule            — mdgen/wrapper.py:196-200
  * module/parameter layout    — mdgen/model/latent_model.py:44-128, mdgen/model/mha.py:111-130,
                                 mdgen/model/ipa.py:66-87, mdgen/model/layers.py:23-27,63-68
The state-dict key names are a checkpoint-compatibility contract.
"""
from __future__ import annotations

import argparse
import dataclasses
from collections import OrderedDict
from typing import Dict, Tuple

# Fixed architecture constants the sm_100a kernels are specialised for
# (defaults of mdgen/parsing.py:87-93; every published checkpoint uses them).
EMBED_DIM = 384
MHA_HEADS = 16
HEAD_DIM = 24
FFN_DIM = 1536
IPA_HEADS = 4
IPA_HEAD_DIM = 32
IPA_QK_POINTS = 8
IPA_V_POINTS = 8
T_FREQ_DIM = 256

#: attributes back-filled with False when missing (mdgen/wrapper.py:178-194,215-216)
BACKFILL_FALSE = (
    "inpainting", "no_torsion", "hyena", "no_aa_emb", "supervise_all_torsions",
    "supervise_no_torsions", "design_key_frames", "no_design_torsion", "cond_interval", "mpnn",
    "dynamic_mpnn", "no_offsets", "no_frames", "ema",
)


def default_args(**overrides) -> argparse.Namespace:
    """A namespace carrying every flag of mdgen/parsing.py:9-120 at its default."""
    d = dict(
        ckpt=None, validate=False, num_workers=4,
        epochs=100, overfit=False, overfit_peptide=None, overfit_frame=False,
        train_batches=None, val_batches=None, val_repeat=1, inference_batches=0,
        batch_size=8, val_freq=None, val_epoch_freq=1, no_validate=False, designability_freq=1,
        print_freq=100, ckpt_freq=1, wandb=False, run_name="default",
        accumulate_grad=1, grad_clip=1.0, check_grad=False, grad_checkpointing=False,
        adamW=False, ema=False, ema_decay=0.999, lr=1e-4, precision="32-true",
        train_split=None, val_split=None, data_dir=None, num_frames=50, crop=256, suffix="",
        atlas=False, copy_frames=False, no_pad=False, short_md=False,
        design_key_frames=False, no_aa_emb=False, no_torsion=False, no_design_torsion=False,
        supervise_no_torsions=False, supervise_all_torsions=False,
        no_offsets=False, no_frames=False,
        hyena=False, no_rope=False, dropout=0.0, scale_factor=1.0, interleave_ipa=False,
        prepend_ipa=False, oracle=False, num_layers=5, embed_dim=384, mha_heads=16,
        ipa_heads=4, ipa_head_dim=32, ipa_qk=8, ipa_v=8, time_multiplier=100.0,
        abs_pos_emb=False, abs_time_emb=False,
        path_type="GVP", prediction="velocity", sampling_method="dopri5", alpha_max=8,
        discrete_loss_weight=0.5, dirichlet_flow_temp=1.0, allow_nan_cfactor=False,
        tps_condition=False, design=False, design_from_traj=False, sim_condition=False,
        inpainting=False, dynamic_mpnn=False, mpnn=False, frame_interval=None,
        cond_interval=None,
    )
    d.update(overrides)
    return argparse.Namespace(**d)


def backfill_args(args) -> None:
    for key in BACKFILL_FALSE:
        if not hasattr(args, key):
            setattr(args, key, False)


@dataclasses.dataclass(frozen=True)
class MDGenConfig:
    """What the device library needs to know (mirrors `mdgen_config` in include/mdgen_b200.h)."""
    latent_dim: int
    num_layers: int
    crop: int                 # L of pos_embed when abs_pos_emb
    abs_pos_emb: bool
    two_trunks: bool          # tps_condition / inpainting: IPA trunk run from both key frames
    use_aa_emb: bool
    time_multiplier: float
    sim_condition: bool
    tps_condition: bool
    inpainting: bool
    cond_interval: int        # 0 = none
    no_torsion: bool = False  # --no_torsion: torsion channels of the latent zeroed (wrapper.py:320-321)

    @property
    def cond_dim(self) -> int:
        return self.latent_dim


UNSUPPORTED_FLAGS = (
    "design", "hyena", "no_rope", "interleave_ipa", "abs_time_emb", "dynamic_mpnn", "mpnn",
    "no_frames", "no_offsets", "design_key_frames",
    "oracle",          # skips the torsion normalisation of inference() (wrapper.py:474-476); decode_kernel normalises
)


def config_from_args(args) -> MDGenConfig:
    """Derives the device configuration; raises for flag combinations that SURVEY.md §8 marks
    out of scope / 'next' (design-mode Dirichlet flow, Hyena, ablations)."""
    backfill_args(args)
    for flag in UNSUPPORTED_FLAGS:
        if getattr(args, flag, False):
            raise NotImplementedError(
                f"mdgen_b200: --{flag} is outside the B200 hot-path scope (SURVEY.md §8f)")
    if not args.prepend_ipa:
        raise NotImplementedError("mdgen_b200: only --prepend_ipa models are supported")
    if (args.embed_dim, args.mha_heads, args.ipa_heads, args.ipa_head_dim, args.ipa_qk,
            args.ipa_v) != (EMBED_DIM, MHA_HEADS, IPA_HEADS, IPA_HEAD_DIM, IPA_QK_POINTS,
                            IPA_V_POINTS):
        raise NotImplementedError(
            "mdgen_b200 kernels are specialised for embed_dim=384, mha_heads=16, "
            "ipa_heads=4, ipa_head_dim=32, ipa_qk=ipa_v=8 (mdgen/parsing.py:87-93 defaults)")
    if args.dropout != 0.0:
        raise NotImplementedError("dropout is not supported on the sampling path")
    two = bool(args.tps_condition or args.inpainting)
    if not (args.sim_condition or two):
        raise NotImplementedError("need --sim_condition, --tps_condition or --inpainting")
    latent_dim = 28 if two else 21          # mdgen/wrapper.py:196
    return MDGenConfig(
        latent_dim=latent_dim, num_layers=int(args.num_layers), crop=int(args.crop),
        abs_pos_emb=bool(args.abs_pos_emb), two_trunks=two and not args.sim_condition,
        use_aa_emb=not args.no_aa_emb, time_multiplier=float(args.time_multiplier),
        sim_condition=bool(args.sim_condition), tps_condition=bool(args.tps_condition),
        inpainting=bool(args.inpainting),
        cond_interval=int(args.cond_interval) if args.cond_interval else 0,
        no_torsion=bool(args.no_torsion),
    )


def _mha_schema(prefix: str, out: "OrderedDict[str, Tuple[int, ...]]") -> None:
    C = EMBED_DIM
    out[prefix + "attn.bias_k"] = (1, 1, C)
    out[prefix + "attn.bias_v"] = (1, 1, C)
    for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
        out[prefix + f"attn.{p}.weight"] = (C, C)
        out[prefix + f"attn.{p}.bias"] = (C,)
    out[prefix + "attn.rot_emb.inv_freq"] = (HEAD_DIM // 2,)


def model_schema(cfg: MDGenConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Key → shape of `LatentMDGenModel.state_dict()` for the supported configurations,
    in the reference's registration order."""
    C, D, F = EMBED_DIM, cfg.latent_dim, FFN_DIM
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    if cfg.abs_pos_emb:
        s["pos_embed"] = (1, cfg.crop, C)
    s["latent_to_emb.weight"] = (C, D)
    s["latent_to_emb.bias"] = (C,)
    if cfg.tps_condition or cfg.inpainting:
        for n in ("latent_to_emb_f", "latent_to_emb_r"):
            s[n + ".weight"] = (C, 7)
            s[n + ".bias"] = (C,)
    s["cond_to_emb.weight"] = (C, D)
    s["cond_to_emb.bias"] = (C,)
    s["mask_to_emb.weight"] = (2, C)
    if cfg.use_aa_emb:
        s["aatype_to_emb.weight"] = (21, C)
    hc = IPA_HEADS * IPA_HEAD_DIM
    for i in range(cfg.num_layers):
        p = f"ipa_layers.{i}."
        s[p + "adaLN_modulation.1.weight"] = (6 * C, C)
        s[p + "adaLN_modulation.1.bias"] = (6 * C,)
        s[p + "ipa_norm.weight"] = (C,)
        s[p + "ipa_norm.bias"] = (C,)
        s[p + "ipa.head_weights"] = (IPA_HEADS,)
        s[p + "ipa.linear_q.weight"] = (hc, C)
        s[p + "ipa.linear_q.bias"] = (hc,)
        s[p + "ipa.linear_kv.weight"] = (2 * hc, C)
        s[p + "ipa.linear_kv.bias"] = (2 * hc,)
        s[p + "ipa.linear_q_points.weight"] = (IPA_HEADS * IPA_QK_POINTS * 3, C)
        s[p + "ipa.linear_q_points.bias"] = (IPA_HEADS * IPA_QK_POINTS * 3,)
        nkv = IPA_HEADS * (IPA_QK_POINTS + IPA_V_POINTS) * 3
        s[p + "ipa.linear_kv_points.weight"] = (nkv, C)
        s[p + "ipa.linear_kv_points.bias"] = (nkv,)
        cat = IPA_HEADS * (IPA_HEAD_DIM + IPA_V_POINTS * 4)
        s[p + "ipa.linear_out.weight"] = (C, cat)
        s[p + "ipa.linear_out.bias"] = (C,)
        _mha_schema(p + "mha_l.", s)
        s[p + "fc1.weight"] = (F, C)
        s[p + "fc1.bias"] = (F,)
        s[p + "fc2.weight"] = (C, F)
        s[p + "fc2.bias"] = (C,)
    for i in range(cfg.num_layers):
        p = f"layers.{i}."
        s[p + "adaLN_modulation.1.weight"] = (9 * C, C)
        s[p + "adaLN_modulation.1.bias"] = (9 * C,)
        _mha_schema(p + "mha_t.", s)
        _mha_schema(p + "mha_l.", s)
        s[p + "fc1.weight"] = (F, C)
        s[p + "fc1.bias"] = (F,)
        s[p + "fc2.weight"] = (C, F)
        s[p + "fc2.bias"] = (C,)
    s["emb_to_latent.linear.weight"] = (D, C)
    s["emb_to_latent.linear.bias"] = (D,)
    s["emb_to_latent.adaLN_modulation.1.weight"] = (2 * C, C)
    s["emb_to_latent.adaLN_modulation.1.bias"] = (2 * C,)
    s["t_embedder.mlp.0.weight"] = (C, T_FREQ_DIM)
    s["t_embedder.mlp.0.bias"] = (C,)
    s["t_embedder.mlp.2.weight"] = (C, C)
    s["t_embedder.mlp.2.bias"] = (C,)
    return s


BUFFER_SUFFIXES = ("pos_embed", "rot_emb.inv_freq")


def is_buffer(key: str) -> bool:
    return key.endswith(BUFFER_SUFFIXES)


def num_parameters(schema: Dict[str, Tuple[int, ...]]) -> int:
    n = 0
    for k, shp in schema.items():
        if is_buffer(k):
            continue
        m = 1
        for d in shp:
            m *= d
        n += m
    return n
