"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE. /root/reference does not exist on the GPU box, so nothing that runs
there may call into this file; it is used by tests/golden/gen_golden.py (here) and by the
`-m "not gpu"` tests that cross-check the oracle against the live reference when present.

Recipe: put the stand-in third-party modules (oracle/ref_shims) and
/root/reference on sys.path, set MODEL_DIR to a scratch dir (mdgen/logger.py opens
$MODEL_DIR/log.out at import), and hide `flash_attn`/`deepspeed` from
importlib.util.find_spec while mdgen/model/primitives.py:20-31 is imported (the installed
flash-attn 2.8 no longer exports flash_attn_unpadded_kvpacked_func).
"""
import argparse
import importlib
import importlib.util
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("MDGEN_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mdgen", "wrapper.py"))


_loaded = None


def load_reference():
    """Returns the imported reference module `mdgen.wrapper` (with mdgen.* in sys.modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    os.environ.setdefault("MODEL_DIR", tempfile.mkdtemp(prefix="mdgen_ref_"))
    os.environ.setdefault("WANDB_MODE", "disabled")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    orig_find_spec = importlib.util.find_spec

    def masked(name, *a, **k):
        if name.split(".")[0] in ("flash_attn", "deepspeed"):
            return None
        return orig_find_spec(name, *a, **k)

    importlib.util.find_spec = masked
    try:
        wrapper = importlib.import_module("mdgen.wrapper")
    finally:
        importlib.util.find_spec = orig_find_spec
    _loaded = wrapper
    return wrapper


def make_args(**overrides) -> argparse.Namespace:
    """Every attribute the reference reads from `args`, at the defaults of
    mdgen/parsing.py:9-120 (restated as data, not copied code)."""
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
