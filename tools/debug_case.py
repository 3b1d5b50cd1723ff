"""Run one golden case through the CUDA path (for compute-sanitizer / debugging)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_case, max_rel
from mdgen_b200.synthetic import euler_time_grid
from mdgen_b200.wrapper import NewMDGenWrapper

name = sys.argv[1]
use_tc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
case, args, cfg, sd, batch, zs, g = load_case(name)
args.sampling_method = "euler"
m = NewMDGenWrapper(args); m.model.load_state_dict(sd); m = m.eval().cuda()
eng = m.model.engine(); eng.set_option("use_tc", use_tc)
if use_tc: eng.set_option("tc_min_rows", 1)
db = {k: v.cuda() for k, v in batch.items()}
prep = m.prep_batch(db); kw = prep["model_kwargs"]
v = m.model.forward_inference(zs.cuda(), torch.tensor(case["t_fwd"]).cuda(), **kw)
torch.cuda.synchronize()
print("forward max_rel", max_rel(v.cpu(), g["v"]))
xk = m.model.sample_euler(zs.cuda(), euler_time_grid(case["K"]), **kw)
print("euler max_rel", max_rel(xk.cpu(), g["x_euler"]))
