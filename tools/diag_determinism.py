"""Run the stress forward several times per option set and report bitwise repeatability and the distance to the golden
velocity. Used under compute-sanitizer as well (which perturbs kernel timing): any run-to-run difference is a race."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import load_case, max_rel
from test_gpu_parity import _wrapper, _dev

case, args, cfg, sd, batch, zs, gold = load_case(sys.argv[1] if len(sys.argv) > 1 else "stress")
m = _wrapper(args, sd, "fp16")
kw = m.prep_batch(_dev(batch))["model_kwargs"]
t = torch.tensor(case["t_fwd"]).cuda()
eng = m.model.engine()
for name, opts in [("default", {}), ("fuse_resid_ln", {"fuse_resid_ln": 1}), ("v7", {"attn_variant": 6}),
                   ("v8_mufu", {"attn_variant": 258})]:
    for k, v in opts.items():
        eng.set_option(k, v)
    outs = [m.model.forward_inference(zs.cuda(), t, **kw).cpu() for _ in range(4)]
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(name, "bitwise_repeatable", same, "vs_gold", ["%.3e" % max_rel(o, gold["v"]) for o in outs],
          "vs_run0", ["%.3e" % max_rel(o, outs[0]) for o in outs[1:]], flush=True)
    for k in opts:
        eng.set_option(k, 0 if k != "attn_variant" else 256)
