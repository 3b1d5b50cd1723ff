#!/usr/bin/env python
"""A/B of the tcgen05 attention variants (option `attn_variant`, see csrc/attention_tc.cuh) on one GPU.

For every variant: device time of a K-step sampling call at the BASELINE configs[1] shape, the per-launch
time of the mha_t family, the deviation of the sampled state from variant 0, and (at B=1) the deviation
from the exact-fp32 SIMT path of the same library. Prints one JSON line per variant.

    python tools/ab_attention.py [--variants 0,1,3,6,14,19] [--batch 64] [--euler-steps 4]

Variant bits: 1 bf16 P.V, 2 staged pre-pass, 4 persistent kernel, 8 persistent with 12 softmax warps, 16 pre-pass
only (timing aid, output undefined), 32/64/96 debug modes of the persistent kernel (no MUFU / no P.V MMAs / no TMEM
traffic in the probability loop; output undefined), 128 bound-adopted first softmax reference (experiment).
`--l4 1` switches the S = 4 residue attention to the shared-memory exchange kernel (must be bit-identical:
rel_vs_variant0 == 0 when run with the same attn_variant).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1,3,131,6,14,19")
    ap.add_argument("--l4", type=int, default=0, help="l4_variant (1 = shared-memory exchange kernel for S = 4)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--residues", type=int, default=4)
    ap.add_argument("--euler-steps", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()

    import torch
    from mdgen_b200.config import default_args
    from mdgen_b200.synthetic import euler_time_grid, synthetic_batch, synthetic_noise, synthetic_state_dict
    from mdgen_b200.wrapper import NewMDGenWrapper

    dev = torch.device("cuda", 0)
    B, T, L, K = a.batch, a.frames, a.residues, a.euler_steps
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T,
                        sampling_method="euler")
    m = NewMDGenWrapper(args)
    m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0))
    m = m.eval().to(dev)
    eng = m.model.engine()
    D = m.latent_dim
    grid = euler_time_grid(K)

    def setup(nb):
        batch = {k: v.to(dev) for k, v in synthetic_batch(nb, T, L, seed=1, vary_frames=False).items()}
        zs = synthetic_noise(nb, T, L, D, seed=2).to(dev)
        return zs, m.prep_batch(batch)["model_kwargs"]

    def rel(x, ref):
        return float(((x - ref).abs().max() / ref.abs().max()).item())

    # exact-fp32 reference of the library itself at B = 1
    zs1, kw1 = setup(1)
    eng.set_option("use_tc", 0)
    ref1 = m.model.sample_euler(zs1, grid, **kw1).clone()
    eng.set_option("use_tc", 1)
    zsB, kwB = setup(B)

    base = None
    for v in [int(x) for x in a.variants.split(",")]:
        rec = {"variant": v}
        try:
            eng.set_option("attn_variant", v)
            eng.set_option("l4_variant", a.l4)
            out1 = m.model.sample_euler(zs1, grid, **kw1)
            rec["rel_vs_fp32_simt_B1"] = rel(out1, ref1)
            out = m.model.sample_euler(zsB, grid, **kwB)       # warm-up + result
            torch.cuda.synchronize()
            rec["finite"] = bool(torch.isfinite(out).all().item())
            if base is None:
                base = out.clone()
            rec["rel_vs_variant0"] = rel(out, base)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                m.model.sample_euler(zsB, grid, **kwB)
            e1.record()
            torch.cuda.synchronize()
            rec["ms_per_euler_step"] = e0.elapsed_time(e1) / a.reps / K
            eng.set_option("profile", 1)
            m.model.sample_euler(zsB, grid, **kwB)
            prof = eng.profile_dump()
            eng.set_option("profile", 0)
            if "mha_t" in prof:
                rec["mha_t_ms_per_launch"] = prof["mha_t"][0] / prof["mha_t"][1]
            if "mha_l" in prof:
                rec["mha_l_ms_per_launch"] = prof["mha_l"][0] / prof["mha_l"][1]
        except Exception as ex:  # keep going: the remaining variants are independent
            rec["error"] = repr(ex)[:300]
        print(json.dumps(rec), flush=True)
    eng.set_option("attn_variant", 3)
    eng.set_option("l4_variant", 0)


if __name__ == "__main__":
    main()
