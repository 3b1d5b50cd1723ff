#!/usr/bin/env python
"""bench.py — sampled frames/sec of the MDGen Euler sampling hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N>1 it
is launched under torch.distributed.run, one rank per GPU. One JSON line on stdout (rank 0).

A "step" is one pass of the hot path over one batch: a complete `--euler-steps`-step Euler
sampling call (mdgen/transport/integrators.py:90-113) over B trajectories of T frames x L residues
(BASELINE.json configs[1]: T=1000, crop=4, 100 Euler steps, batch 64 on one B200).
  value : B_total*T / time with the state, conditioning and weights already resident in HBM
          (mdgen_sample_euler through the C ABI; final ODE state out).
  e2e   : same metric through the public API NewMDGenWrapper.inference(batch) with HOST (pinned)
          buffers: H2D of the batch, featurisation, noise, sampling, decode to atom14, D2H.
Batches shard across ranks with no data-path collective: "scaling": "weak".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sampled frames/sec (1000x4-token seq, 100 Euler steps)"
C, H, FF = 384, 16, 1536


def flops_forward(N, T, L, layers=5):
    """Algorithmic FLOPs of one denoiser forward."""
    return N * layers * (32 * C * C + 4 * C * (T + 1) + 4 * C * (L + 1))


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=64, help="trajectories per GPU")
    p.add_argument("--frames", type=int, default=1000)
    p.add_argument("--residues", type=int, default=4)
    p.add_argument("--euler-steps", type=int, default=100)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-torch-gpu", action="store_true", help="skip timing the stock-ATen port on the GPU")
    p.add_argument("--use-tc", type=int, default=-1, help="-1 = library default")
    p.add_argument("--use-graph", type=int, default=-1, help="-1 = library default (CUDA-graph step replay for small workloads)")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_port_run(B, T, L, k_sample, euler_steps, threads):
    """Times the oracle port (CPU restatement of the reference path) on a bounded sample:
    `k_sample` of the `euler_steps` Euler steps for B trajectories; returns extrapolated frames/s."""
    import torch
    from mdgen_b200.config import config_from_args, default_args
    from mdgen_b200.synthetic import (euler_time_grid, synthetic_batch, synthetic_noise,
                                      synthetic_state_dict)
    from oracle import mdgen_oracle as O
    torch.set_num_threads(threads)
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T)
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(B, T, L, seed=1, vary_frames=False)
    zs = synthetic_noise(B, T, L, cfg.latent_dim, seed=2)
    grid = euler_time_grid(euler_steps)[: k_sample + 1]
    with torch.no_grad():
        t0 = time.perf_counter()
        op = O.prep_batch(cfg, batch)
        kw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
                  x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
        t1 = time.perf_counter()
        O.sample_euler(sd, cfg, zs, grid, **kw)
        t2 = time.perf_counter()
    per_step = (t2 - t1) / k_sample
    full = per_step * euler_steps
    return B * T / full, {"prep_s": t1 - t0, "sec_per_euler_step": per_step}


def best_cpu_threads(T, L):
    """torch CPU ops on this path stop scaling (and regress) well before 128 threads: pick the
    thread count that maximises the port's forward throughput on this host."""
    import torch
    from mdgen_b200.config import config_from_args, default_args
    from mdgen_b200.synthetic import synthetic_batch, synthetic_noise, synthetic_state_dict
    from oracle import mdgen_oracle as O
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T)
    cfg = config_from_args(args)
    sd = synthetic_state_dict(cfg, seed=0)
    batch = synthetic_batch(1, T, L, seed=1, vary_frames=False)
    zs = synthetic_noise(1, T, L, cfg.latent_dim, seed=2)
    op = O.prep_batch(cfg, batch)
    kw = dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
              x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            O.forward(sd, cfg, zs, torch.zeros(1), **kw)   # warm
            t0 = time.perf_counter()
            O.forward(sd, cfg, zs, torch.zeros(1), **kw)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(a, rank, world):
    """--impl reference: the reference's CPU implementation of the path. The reference is Python
    and cannot travel to the GPU box, so this is the pinned oracle port (oracle/mdgen_oracle.py),
    all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    threads = best_cpu_threads(a.frames, a.residues)
    kb, ks = 1, 2   # 1 trajectory, 2 of the 100 Euler steps per bench step (~3-4 s of CPU work)
    vals = []
    for i in range(a.warmup + a.steps):
        v, _ = cpu_port_run(kb, a.frames, a.residues, ks, a.euler_steps, threads)
        if i >= a.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    sample = (f"oracle port, B={kb} trajectory x T={a.frames} x L={a.residues}, {ks} of {a.euler_steps} "
              f"Euler steps timed, extrapolated x{a.euler_steps // ks}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * kb * a.frames / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"tetrapeptide forward-sim T={a.frames} L={a.residues} "
                               f"{a.euler_steps} Euler steps (CPU sample)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mdgen_b200.config import default_args
    from mdgen_b200.dist import gather_counts, max_over_ranks, rank_seed
    from mdgen_b200.synthetic import (euler_time_grid, synthetic_batch, synthetic_noise,
                                      synthetic_state_dict)
    from mdgen_b200.wrapper import NewMDGenWrapper

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, T, L, K = a.batch, a.frames, a.residues, a.euler_steps
    args = default_args(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L,
                        num_frames=T, sampling_method="euler")
    m = NewMDGenWrapper(args)
    m.model.load_state_dict(synthetic_state_dict(m.cfg, seed=0))
    m = m.eval().to(dev)
    eng = m.model.engine()
    if a.use_tc >= 0:
        eng.set_option("use_tc", a.use_tc)
    if a.use_graph >= 0:
        eng.set_option("use_graph", a.use_graph)
    D = m.latent_dim
    # every rank samples its own shard of independent trajectories (different seeds per rank)
    hbatch = synthetic_batch(B, T, L, seed=rank_seed(1, rank), vary_frames=False)
    hbatch = {k: v.pin_memory() for k, v in hbatch.items()}
    dbatch = {k: v.to(dev, non_blocking=True) for k, v in hbatch.items()}
    zs = synthetic_noise(B, T, L, D, seed=rank_seed(2, rank)).to(dev)
    grid = euler_time_grid(K)
    prep = m.prep_batch(dbatch)
    kw = prep["model_kwargs"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` calls bracketed by barrier+synchronize; CUDA-event time, max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
        sync_all()
        return ms

    def hot():
        return m.model.sample_euler(zs, grid, **kw)

    def e2e_call():
        # host (pinned) batch -> device, featurise, noise, sample, decode, atom14 -> host
        db = {k: v.to(dev, non_blocking=True) for k, v in hbatch.items()}
        atom14, _ = m.inference(db, num_steps=K + 1)
        return atom14.cpu()

    for _ in range(a.warmup):
        hot()
    l0 = eng.launch_count
    clocks = ClockSampler(local_rank)
    ms = timed(hot, a.steps)
    clk = clocks.stop()
    launches = eng.launch_count - l0
    ms_per_step = ms / a.steps
    total_traj = sum(gather_counts(B, device=dev))          # trajectories sampled per step, all ranks
    value = total_traj * T / (ms_per_step / 1e3)

    e2e = None
    if not a.no_e2e:
        e2e_call()
        ms_e = timed(e2e_call, max(1, min(a.steps, 2))) / max(1, min(a.steps, 2))
        h2d = sum(v.numel() * v.element_size() for v in hbatch.values())
        d2h = B * T * L * 14 * 3 * 4
        e2e = {"value": total_traj * T / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e,
               "api": "NewMDGenWrapper.inference(batch) on pinned host tensors -> atom14.cpu()"}

    # ---- per-kernel device time of one profiled sampling call (CUDA events on the launch stream,
    #      recorded inside the library around every kernel family)
    eng.set_option("profile", 1)
    hot()
    prof = eng.profile_dump()
    eng.set_option("profile", 0)
    peaks = load_peaks()
    N = B * T * L
    fam_flops = {  # algorithmic FLOPs per launch of each tensor-bound family
        "gemm_qkv": 2 * N * 1152 * C, "gemm_out": 2 * N * C * C, "gemm_fc1": 2 * N * FF * C,
        "gemm_fc2": 2 * N * C * FF, "mha_t": 4 * N * H * 24 * (T + 1), "mha_l": 4 * N * H * 24 * (L + 1),
    }
    total_ms = sum(v[0] for k, v in prof.items() if k != "ipa_gemm" and k != "ipa_mha") or 1.0   # (nested in ipa_trunk)
    shares = {k: round(v[0] / total_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    dom = max((k for k in prof if k in fam_flops), key=lambda k: prof[k][0])
    dom_ms = prof[dom][0] / prof[dom][1]
    achieved = fam_flops[dom] / (dom_ms / 1e3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    use_tc = eng.get_option("use_tc")
    fam_tflops = {k: round(fam_flops[k] / (prof[k][0] / prof[k][1] / 1e3) / 1e12, 1) for k in fam_flops if k in prof}
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_attention_ncu.md:
    # attn_tc_kernel<PV16> 1.14 GB read + 0.19 GB write, attn_prep2_kernel 0.40 + 0.60 GB) — valid for this workload only
    traffic = 2.33e9 if (dom == "mha_t" and (B, T, L) == (64, 1000, 4)) else None
    roofline = {
        "kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic, "family_tflops": fam_tflops,
        "note": "mha_t = attn_prep2_kernel + attn_tc_kernel; at head_dim 24 it is bound by TMEM read bandwidth "
                "(every fp32 score is read back once: 64 B/clk/SM) and the MUFU ex2 pipe (96 MMA FLOP per "
                "exponential), not by the tensor pipe - see profiles/r1_attention_ncu.md; the GEMM families' "
                "achieved TFLOP/s are in family_tflops",
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); "
                       + ("token GEMMs and attention P.V: bf16 operands (kind::f16), attention Q.K^T: TF32 operands"
                          if use_tc else "fp32 SIMT FMA (validation path), not the tensor pipe"),
        "avg_launch_ms": dom_ms, "launches": prof[dom][1], "time_shares": shares,
        # the resource that actually limits the dominant kernel, from the committed ncu capture of this workload
        "limiter": ({"resource": "TMEM read bandwidth (tcgen05.ld of the fp32 score tiles)", "achieved": 44.8,
                     "peak": 64.0, "unit": "B/clk/SM", "frac": 0.70,
                     "source": "profiles/r1_attention_ncu.md: 10.2 M LDTM.x16 x 2 KB in the 1.665 ms attn_tc_kernel; "
                               "peak from the microarchitecture notes (TMEM read 64 B/clk/SM); MUFU pipe 65 %"}
                    if (dom == "mha_t" and (B, T, L) == (64, 1000, 4)) else None),
        "whole_step_tflops": flops_forward(N, T, L) * K / (ms_per_step / 1e3) / 1e12,
    }

    # ---- the north-star's denominator: the reference's stock-ATen formulation (materialised fp32
    #      attention, no TF32) on the SAME B200, here through the pinned oracle port (it omits the
    #      reference's wasted head-mean of the attention weights, mha.py:399-405, so it is if anything
    #      faster than the real reference). Bounded sample: 2 of the K Euler steps, extrapolated.
    torch_gpu = None
    if rank == 0 and not a.no_torch_gpu:
        try:
            from oracle import mdgen_oracle as O
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            sd_dev = {k_: v_.to(dev) for k_, v_ in synthetic_state_dict(m.cfg, seed=0).items()}
            with torch.no_grad():
                # featurisation taken from our prep kernel (the reference's batched 4x4 eigh,
                # rigid_utils.py:191-210, fails in cuSOLVER at 256,000 matrices on this stack)
                okw = dict(mask=kw["mask"].float(), start=(dbatch["rots"][:, 0], dbatch["trans"][:, 0]),
                           end=(dbatch["rots"][:, -1], dbatch["trans"][:, -1]), x_cond=kw["x_cond"],
                           x_cond_mask=kw["x_cond_mask"], aatype=kw["aatype"])
                ks = 2
                g2 = grid[: ks + 1]
                O.sample_euler(sd_dev, m.cfg, zs, grid[:2], **okw)      # warm-up (1 step)
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                xr = O.sample_euler(sd_dev, m.cfg, zs, g2, **okw)
                t1.record()
                torch.cuda.synchronize()
                ms_ref = t0.elapsed_time(t1) / ks
                # same-step agreement of our path with the port on this very input (sanity, not the parity gate)
                xo = m.model.sample_euler(zs, g2, **kw)
                agree = float((xo - xr).abs().max() / xr.abs().max())
            torch_gpu = {"value": B * T / (ms_ref * K / 1e3), "unit": "frames/s", "ms_per_euler_step": ms_ref,
                         "sample": f"oracle port (stock ATen, fp32, allow_tf32=False) on this B200: B={B}, "
                                   f"{ks} of {K} Euler steps timed, extrapolated", "max_rel_diff_vs_ours_2_steps": agree}
            del sd_dev, xr, xo, okw
            torch.cuda.empty_cache()
        except Exception as e:  # e.g. out of memory for the materialised score tensors
            torch_gpu = {"unavailable": repr(e)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = best_cpu_threads(T, L)
        v, extra = cpu_port_run(1, T, L, 4, K, threads)
        cpu_baseline = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                        "sample": f"oracle port: B=1 x T={T} x L={L}, 4 of {K} Euler steps timed "
                                  f"({extra['sec_per_euler_step']:.2f} s/step, best of 8/16/32/64/all threads), "
                                  f"extrapolated x{K / 4:g}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("bf16" if eng.get_option("gemm_bf16") else "tf32") if use_tc else "f32", "data": "synthetic",
            "config": {"workload": f"tetrapeptide forward-sim num_frames={T} crop={L}, {K} Euler steps, "
                                   f"batch {B} per GPU (BASELINE.json configs[1])",
                       "tokens_per_forward": N, "parallelism": f"dp{world} (independent trajectories, "
                       "no data-path collective)", "l2": "working set (>4 GB activations per forward) "
                       "far exceeds the 126 MB L2; no explicit flush needed",
                       "graph_replays": eng.get_option("graph_replays"),
                       "gemm_path": ("tcgen05 " + ("bf16 operands (token GEMMs, attention P.V) / TF32 (attention Q.K^T), fp32 accumulate; "
                                     "IPA key-frame trunk fp32" if eng.get_option("gemm_bf16") else "TF32"))
                                    if use_tc else "fp32 SIMT"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "torch_gpu_reference": torch_gpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
