"""Import the UNMODIFIED reference: from /root/reference in the build container, else from the
byte-identical git-ignored copy `oracle/_ref` made by `oracle/vendor_reference.py` (the copy travels
to the GPU box with the snapshot; its MANIFEST.json pins every file's sha256).

TEST / BENCH INFRASTRUCTURE. Used by tests/golden/gen_golden.py (here), by the tests that
cross-check against the live reference, and by bench.py's reference legs. Never by the product.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "ref_shims")


def _resolve_root():
    for root in (os.environ.get("MDGEN_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(root, "mdgen", "wrapper.py")):
            return root
    return None


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return REFERENCE_ROOT is not None


def reference_kind() -> str:
    """'live' (/root/reference), 'vendored' (oracle/_ref, sha256-verified) or 'absent'."""
    if REFERENCE_ROOT is None:
        return "absent"
    if os.path.abspath(REFERENCE_ROOT) == os.path.join(_HERE, "_ref"):
        from . import vendor_reference
        if not vendor_reference.verify():
            raise RuntimeError("oracle/_ref does not match its MANIFEST.json (not the unmodified reference)")
        return "vendored"
    return "live"


_loaded = None


def load_reference():
    """Returns the imported reference module `mdgen.wrapper` (with mdgen.* in sys.modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not found (neither /root/reference nor oracle/_ref)")
    reference_kind()
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


def chunked_eigh_patch(chunk=8192):
    """Context manager: torch.linalg.eigh evaluated in batches of `chunk` matrices. cuSOLVER's batched
    syevj refuses the 256,000 4x4 matrices of `rot_to_quat` (mdgen/rigid_utils.py:191-210) at BASELINE
    configs[1]; splitting the batch changes no value. A patch on torch, not on the reference."""
    import contextlib

    import torch

    @contextlib.contextmanager
    def cm():
        orig = torch.linalg.eigh

        def eigh(a, *args, **kw):
            if a.dim() < 3 or a.shape[:-2].numel() <= chunk:
                return orig(a, *args, **kw)
            flat = a.reshape(-1, a.shape[-2], a.shape[-1])
            ws, vs = [], []
            for i in range(0, flat.shape[0], chunk):
                w, v = orig(flat[i:i + chunk], *args, **kw)
                ws.append(w); vs.append(v)
            return (torch.cat(ws).reshape(*a.shape[:-2], a.shape[-1]),
                    torch.cat(vs).reshape(a.shape))

        torch.linalg.eigh = eigh
        try:
            yield
        finally:
            torch.linalg.eigh = orig

    return cm()


def reference_wrapper(args, sd, device="cpu"):
    """The reference's own NewMDGenWrapper with `sd` loaded strictly into its model, in eval mode."""
    import torch
    wrapper_mod = load_reference()
    torch.manual_seed(0)
    m = wrapper_mod.NewMDGenWrapper(args).eval()
    m.model.load_state_dict(sd, strict=True)
    return m.to(device)
