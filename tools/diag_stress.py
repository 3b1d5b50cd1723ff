"""Precision of every kernel path on the "trained-like" stress weights (tests/golden/stress.npz, produced by the
unmodified reference): max-rel / rel-L2 error of the forward velocity and of the K-step Euler state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgen_b200.synthetic import euler_time_grid
from mdgen_b200.wrapper import NewMDGenWrapper
from tests.helpers import load_case, max_rel, rel_l2

names = sys.argv[1:] or ["stress", "stress_atlas", "sim_c1"]
for name in names:
    case, args, cfg, sd, batch, zs, g = load_case(name)
    args.sampling_method = "euler"
    m = NewMDGenWrapper(args)
    m.model.load_state_dict(sd)
    m = m.eval().to("cuda")
    eng = m.model.engine()
    kw = m.prep_batch({k: v.cuda() for k, v in batch.items()})["model_kwargs"]
    t = torch.tensor(case["t_fwd"]).cuda()
    for label, opts in [("simt fp32", dict(use_tc=0)),
                        ("tf32 gemm + tf32 attn(v0)", dict(use_tc=1, gemm_bf16=0, attn_variant=0)),
                        ("tf32 gemm + attn v3", dict(use_tc=1, gemm_bf16=0, attn_variant=3)),
                        ("fp16 gemm + attn v8 (default)", dict(use_tc=1, gemm_bf16=2, attn_variant=256)),
                        ("tf32 gemm + attn v8", dict(use_tc=1, gemm_bf16=0, attn_variant=256)),
                        ("fp16 gemm + attn v0", dict(use_tc=1, gemm_bf16=2, attn_variant=0)),
                        ("fp16 gemm + attn v3", dict(use_tc=1, gemm_bf16=2, attn_variant=3)),
                        ("fp16 gemm, simt attention", dict(use_tc=1, gemm_bf16=2, use_tc_attn=0)),
                        ("bf16 gemm + attn v0", dict(use_tc=1, gemm_bf16=1, attn_variant=0)),
                        ("bf16 gemm + attn v3 (default)", dict(use_tc=1, gemm_bf16=1, attn_variant=3)),
                        ("bf16 gemm, simt attention", dict(use_tc=1, gemm_bf16=1, use_tc_attn=0)),
                        ("bf16 MLP only (emu)", dict(use_tc=1, gemm_bf16=0, emu_bf16=1)),
                        ("bf16 attn proj only (emu)", dict(use_tc=1, gemm_bf16=0, emu_bf16=2)),
                        ("bf16 qkv store only (emu)", dict(use_tc=1, gemm_bf16=0, emu_bf16=4)),
                        ]:
        for k_, v_ in dict(use_tc=1, gemm_bf16=2, attn_variant=3, use_tc_attn=1, emu_bf16=0).items():
            eng.set_option(k_, v_)
        eng.set_option("tc_min_rows", 65)
        for k_, v_ in opts.items():
            eng.set_option(k_, v_)
        v = m.model.forward_inference(zs.cuda(), t, **kw)
        xk = m.model.sample_euler(zs.cuda(), euler_time_grid(case["K"]), **kw)
        print(f"{name:13s} {label:32s} v max-rel {max_rel(v.cpu(), g['v']):.2e} l2 {rel_l2(v.cpu(), g['v']):.2e} | "
              f"x_euler max-rel {max_rel(xk.cpu(), g['x_euler']):.2e} l2 {rel_l2(xk.cpu(), g['x_euler']):.2e}")
