"""Precision diagnostics on the GPU: error of forward / K-step Euler / 49-step inference vs golden
for each case under different kernel-path options."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_case, max_rel, rel_l2
from tests.golden.cases import CASES
from mdgen_b200.synthetic import euler_time_grid
from mdgen_b200.wrapper import NewMDGenWrapper

cases = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)
configs = [
    dict(use_tc=1, tc_min_rows=65, gemm_bf16=0),                 # TF32 token GEMMs
    dict(use_tc=1, tc_min_rows=65, gemm_bf16=1),                 # production: bf16 token GEMMs
    dict(use_tc=1, tc_min_rows=65, gemm_bf16=1, emu_bf16=4),     # + q,k,v rounded to bf16 (emulated)
]
for name in cases:
    case, args, cfg, sd, batch, zs, g = load_case(name)
    args.sampling_method = "euler"
    for conf in configs:
        m = NewMDGenWrapper(args); m.model.load_state_dict(sd); m = m.eval().cuda()
        eng = m.model.engine()
        for k, v in conf.items():
            eng.set_option(k, v)
        db = {k: v.cuda() for k, v in batch.items()}
        prep = m.prep_batch(db); kw = prep["model_kwargs"]
        v = m.model.forward_inference(zs.cuda(), torch.tensor(case["t_fwd"]).cuda(), **kw)
        xk = m.model.sample_euler(zs.cuda(), euler_time_grid(case["K"]), **kw)
        x49 = m.model.sample_euler(zs.cuda(), euler_time_grid(49), **kw)
        atom14, _ = m.inference(db, zs=zs.cuda())
        print(f"{name:12s} {str(conf):60s} fwd {max_rel(v.cpu(), g['v']):.2e} eulerK {max_rel(xk.cpu(), g['x_euler']):.2e} "
              f"x49 max {max_rel(x49.cpu(), g['x49']):.2e} l2 {rel_l2(x49.cpu(), g['x49']):.2e} "
              f"atom14 max {max_rel(atom14.cpu(), g['atom14']):.2e} l2 {rel_l2(atom14.cpu(), g['atom14']):.2e}", flush=True)
