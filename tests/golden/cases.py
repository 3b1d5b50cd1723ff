"""Golden-vector cases shared by the generator (runs the live reference in the build container)
and the parity tests (run anywhere). Flags mirror the README commands of the reference
(README.md:46-61); sizes are small enough for the CPU suite."""

_SIM = dict(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4)

CASES = {
    # BASELINE.json configs[0]: tetrapeptide forward-sim T=64 crop=4, 10 Euler steps (B=2 here)
    "sim_c1": dict(args=dict(_SIM, num_frames=64), B=2, T=64, L=4, K=10, t_fwd=[0.3, 0.7]),
    # ATLAS-shaped (no abs_pos_emb, longer chain, padded residues -> key-padding path)
    "atlas_small": dict(args=dict(sim_condition=True, prepend_ipa=True, crop=24, num_frames=12),
                        B=2, T=12, L=24, K=4, t_fwd=[0.55, 0.05], batch=dict(pad_last=5)),
    # upsampling: conditioning every cond_interval frames
    "upsampling": dict(args=dict(_SIM, num_frames=33, cond_interval=8), B=1, T=33, L=4, K=5,
                       t_fwd=[0.9], batch=dict(cond_interval=8)),
    # transition-path sampling: both end frames conditioned, two IPA trunks, latent_dim 28
    "tps": dict(args=dict(tps_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=4,
                          num_frames=16), B=2, T=16, L=4, K=4, t_fwd=[0.2, 0.6],
                canonical_quat=True),
    # inpainting without the design head (design-mode Dirichlet flow is a "next" row)
    "inpaint": dict(args=dict(inpainting=True, prepend_ipa=True, abs_pos_emb=True, crop=4,
                              num_frames=10, no_aa_emb=True, no_torsion=True),
                    B=2, T=10, L=4, K=3, t_fwd=[0.4, 0.8], canonical_quat=True),
    # "trained-like" stress weights (synthetic_state_dict(stress=True): sharp softmax rows, O(1) gates and
    # velocities) on a time axis long enough for the tcgen05 attention (3 key tiles, ragged tail)
    "stress": dict(args=dict(_SIM, num_frames=200), B=1, T=200, L=4, K=3, t_fwd=[0.45], stress=True),
    # the same weights on an ATLAS-shaped chain (residue attention over 80 residues with padding)
    "stress_atlas": dict(args=dict(sim_condition=True, prepend_ipa=True, crop=80, num_frames=10),
                         B=1, T=10, L=80, K=2, t_fwd=[0.7], batch=dict(pad_last=7), stress=True),
}
