"""oracle/ — TEST INFRASTRUCTURE, not product code.

CPU restatement (plain PyTorch fp32 on the host) of the MDGen hot path named in
BASELINE.json's north_star, plus the loader that imports the *unmodified* reference from
/root/reference in the build container to pin the restatement with golden vectors.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import from
here; `mdgen_b200/` (the product) must never do so.
"""
