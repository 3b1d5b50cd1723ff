"""Recipe: make the UNMODIFIED reference importable on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY. `/root/reference` exists in the build container but not on the
GPU box. This script copies the reference's Python package (`/root/reference/mdgen/**.py`, 328 KB,
byte for byte) into `oracle/_ref/mdgen/` and writes `oracle/_ref/MANIFEST.json` with the sha256 of
every file. `oracle/_ref/` is git-ignored (no reference source ever enters history) but NOT
gpurun-ignored, so it travels with the snapshot like the built `.so`. `oracle/ref_loader.py`
imports from `/root/reference` when present, else from `oracle/_ref`.

Consumers (all checkers / baselines, never the product path):
  * `bench.py --impl reference`  : the reference's own `NewMDGenWrapper.inference` (mdgen/wrapper.py:405-484)
  * `bench.py` `torch_gpu_reference` leg and the `-m gpu` parity tests at BASELINE shapes: the
    reference's `forward_inference` / `sample_ode('euler')` on the same B200, fp32 (allow_tf32 off)

    python -m oracle.vendor_reference          # run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("MDGEN_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")


def vendor(force: bool = False) -> str | None:
    """Copies the package; returns the destination root, or None when the source is absent."""
    src = os.path.join(SRC_ROOT, "mdgen")
    if not os.path.isfile(os.path.join(src, "wrapper.py")):
        return DST_ROOT if os.path.isfile(os.path.join(DST_ROOT, "mdgen", "wrapper.py")) else None
    manifest = {}
    for dirpath, _, files in os.walk(src):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            sp = os.path.join(dirpath, f)
            rel = os.path.relpath(sp, SRC_ROOT)
            dp = os.path.join(DST_ROOT, rel)
            data = open(sp, "rb").read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            if not force and os.path.isfile(dp) and open(dp, "rb").read() == data:
                continue
            os.makedirs(os.path.dirname(dp), exist_ok=True)
            shutil.copyfile(sp, dp)
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST_ROOT


def verify() -> bool:
    """True when every vendored file still matches the manifest (i.e. is the unmodified reference)."""
    mpath = os.path.join(DST_ROOT, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    files = json.load(open(mpath))["files"]
    for rel, digest in files.items():
        p = os.path.join(DST_ROOT, rel)
        if not os.path.isfile(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != digest:
            return False
    return bool(files)


if __name__ == "__main__":
    print(vendor(), "verified" if verify() else "NOT verified")
