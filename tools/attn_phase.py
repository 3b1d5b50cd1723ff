"""Debug aid: per-phase cycle counters of the generation-8 attention softmax warps (attn_variant 258) at BASELINE configs[1]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgen_b200.config import default_args
from mdgen_b200.synthetic import euler_time_grid, synthetic_batch, synthetic_noise, synthetic_state_dict
from mdgen_b200.wrapper import NewMDGenWrapper
B, T, L = 64, 1000, 4
args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=True, crop=L, num_frames=T, sampling_method="euler")
m = NewMDGenWrapper(args); m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0)); m = m.eval().cuda()
eng = m.model.engine()
kw = m.prep_batch({k: v.cuda() for k, v in synthetic_batch(B, T, L, seed=1, vary_frames=False).items()})["model_kwargs"]
zs = synthetic_noise(B, T, L, m.latent_dim, seed=2).cuda()
m.model.sample_euler(zs, euler_time_grid(100)[:2], **kw)
eng.set_option("attn_variant", int(sys.argv[1]) if len(sys.argv) > 1 else 258)
m.model.sample_euler(zs, euler_time_grid(100)[:2], **kw)
torch.cuda.synchronize()
