#!/usr/bin/env python
"""CPU study for the next attention iteration: how loose is the Cauchy-Schwarz reference
m_i = |q_i| max_j |k_j| against the true row maximum of the frame-attention scores, and what would fp16 storage
of the shifted scores s - m cost in the probabilities? Runs the oracle forward once (synthetic weights,
B = 1, T frames, L = 4) and analyses every mha_t call. Scores are in log2 units like in the kernel."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--t", type=float, default=0.5)
    a = ap.parse_args()
    from mdgen_b200.config import config_from_args, default_args
    from mdgen_b200.synthetic import synthetic_batch, synthetic_noise, synthetic_state_dict
    from oracle import mdgen_oracle as O
    T, L = a.frames, 4
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=L, num_frames=T)
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(1, T, L, seed=1, vary_frames=False)
    zs = synthetic_noise(1, T, L, cfg.latent_dim, seed=2)
    op = O.prep_batch(cfg, batch)
    kw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
              x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
    calls = []
    orig = O.mha

    def spy(sd_, prefix, x, mask):
        if "mha_t" in prefix:
            calls.append((prefix + "attn.", x.detach().clone()))
        return orig(sd_, prefix, x, mask)
    O.mha = spy
    with torch.no_grad():
        O.forward(sd, cfg, zs, torch.full((1,), a.t), **kw)
    O.mha = orig
    LOG2E = 1.4426950408889634
    for prefix, x in calls:
        Bq, S, C = x.shape
        q = torch.nn.functional.linear(x, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"]) * (O.HD ** -0.5)
        k = torch.nn.functional.linear(x, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
        k = torch.cat([k, sd[prefix + "bias_k"].reshape(1, 1, C).expand(Bq, 1, C)], 1)
        q = q.reshape(Bq, S, O.H, O.HD).permute(0, 2, 1, 3)
        k = k.reshape(Bq, S + 1, O.H, O.HD).permute(0, 2, 1, 3)
        cos, sin = O.rope_tables(S + 1, sd[prefix + "rot_emb.inv_freq"])
        rot = lambda t, n: t * cos[:n] + O.rotate_half(t) * sin[:n]
        q, k = rot(q, S), rot(k, S + 1)
        s = (q @ k.transpose(-1, -2)) * LOG2E                     # [Bq, H, S, S+1] in log2 units
        smax = s.max(-1).values
        bound = q.norm(dim=-1) * LOG2E * k.norm(dim=-1).max(-1, keepdim=True).values
        gap = bound - smax
        p = torch.exp2(s - smax[..., None])
        l = p.sum(-1)
        # fp16 storage of the shifted score with the bound as reference: abs error 2^-11 |s - bound|
        sh = (s - bound[..., None]).to(torch.float16).to(torch.float32)
        p16 = torch.exp2(sh + (bound - smax)[..., None])
        rel = ((p16 - p).abs().sum(-1) / l)
        print(f"{prefix:32s} |s|max {float(s.abs().max()):6.1f}  row max {float(smax.mean()):6.2f}  "
              f"bound-max: mean {float(gap.mean()):5.1f} p99 {float(gap.flatten().quantile(0.99)):5.1f} max {float(gap.max()):5.1f}  "
              f"fp16(s-bound): sum|dp|/l mean {float(rel.mean()):.2e} max {float(rel.max()):.2e}  "
              f"n_eff {float((l * l / (p * p).sum(-1)).mean()):6.1f}")


if __name__ == "__main__":
    main()
