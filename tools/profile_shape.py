"""Per-family device time of one K-step sampling call at a given shape: python tools/profile_shape.py B T L [K] [option=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgen_b200.config import default_args
from mdgen_b200.synthetic import euler_time_grid, synthetic_batch, synthetic_noise, synthetic_state_dict
from mdgen_b200.wrapper import NewMDGenWrapper
B, T, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K = int(sys.argv[4]) if len(sys.argv) > 4 and "=" not in sys.argv[4] else 4
OPTS = dict(a.split("=") for a in sys.argv[4:] if "=" in a)          # engine options, e.g. gemm_dbg=1
args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T, sampling_method="euler")
m = NewMDGenWrapper(args); m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0)); m = m.eval().cuda()
eng = m.model.engine()
kw = m.prep_batch({k: v.cuda() for k, v in synthetic_batch(B, T, L, seed=1, vary_frames=False).items()})["model_kwargs"]
zs = synthetic_noise(B, T, L, m.latent_dim, seed=2).cuda()
grid = euler_time_grid(100)[: K + 1]
for k, v in OPTS.items():
    eng.set_option(k, int(v))
m.model.sample_euler(zs, grid, **kw)
eng.set_option("profile", 1)
m.model.sample_euler(zs, grid, **kw)
prof = eng.profile_dump()
tot = sum(v[0] for k, v in prof.items() if k not in ("ipa_gemm", "ipa_mha"))
print(f"B={B} T={T} L={L} K={K}: {tot:.2f} ms total, {tot / K:.2f} ms per Euler step")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:12s} {v[0]:9.3f} ms  {v[1]:4d} calls  {v[0] / v[1]:8.4f} ms/call  {100 * v[0] / tot:5.1f}%")
