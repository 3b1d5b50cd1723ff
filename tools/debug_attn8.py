"""Debug aid: forward velocity of attn_variant 256 / 257 vs the exact-fp32 SIMT path on several shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgen_b200.config import default_args
from mdgen_b200.synthetic import synthetic_batch, synthetic_noise, synthetic_state_dict
from mdgen_b200.wrapper import NewMDGenWrapper

def run(B, T, L, pad):
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T, sampling_method="euler")
    m = NewMDGenWrapper(args); m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0)); m = m.eval().cuda()
    eng = m.model.engine()
    eng.set_option("tc_min_rows", 65)
    kw = m.prep_batch({k: v.cuda() for k, v in synthetic_batch(B, T, L, seed=3, pad_last=pad).items()})["model_kwargs"]
    zs = synthetic_noise(B, T, L, m.latent_dim, seed=4).cuda()
    t = torch.full((B,), 0.4).cuda()
    eng.set_option("use_tc", 0)
    ref = m.model.forward_inference(zs, t, **kw).clone()
    eng.set_option("use_tc", 1)
    out = {}
    for v in (3, 256, 257):
        eng.set_option("attn_variant", v)
        errs = []
        for rep in range(3):
            o = m.model.forward_inference(zs, t, **kw)
            errs.append(float((o - ref).abs().max() / ref.abs().max()))
        out[v] = ["%.2e" % e for e in errs]
    print((B, T, L, pad), out, flush=True)

shapes = [(2, 40, 70, 0), (2, 40, 70, 5), (1, 40, 96, 0), (1, 40, 130, 3)] if len(sys.argv) > 1 else [(2, 40, 70, 0), (2, 40, 70, 5), (1, 300, 4, 0), (2, 1000, 4, 0), (1, 100, 4, 0), (1, 40, 96, 0), (1, 40, 130, 3)]
for shp in shapes:
    run(*shp)
