"""One Euler step at BASELINE configs[1] with options from the command line (key=value ...): for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgen_b200.config import default_args
from mdgen_b200.synthetic import euler_time_grid, synthetic_batch, synthetic_noise, synthetic_state_dict
from mdgen_b200.wrapper import NewMDGenWrapper
opts = dict(kv.split("=") for kv in sys.argv[1:])
B, T, L = int(opts.pop("B", 64)), int(opts.pop("T", 1000)), int(opts.pop("L", 4))
args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T, sampling_method="euler")
m = NewMDGenWrapper(args); m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0)); m = m.eval().cuda()
eng = m.model.engine()
for k, v in opts.items():
    eng.set_option(k, int(v))
kw = m.prep_batch({k: v.cuda() for k, v in synthetic_batch(B, T, L, seed=1, vary_frames=False).items()})["model_kwargs"]
zs = synthetic_noise(B, T, L, m.latent_dim, seed=2).cuda()
m.model.sample_euler(zs, euler_time_grid(100)[:2], **kw)
torch.cuda.synchronize()
# second call between cudaProfilerStart/Stop: `ncu --profile-from-start off` then sees exactly one sampling call
torch.cuda.profiler.start()
m.model.sample_euler(zs, euler_time_grid(100)[:2], **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
