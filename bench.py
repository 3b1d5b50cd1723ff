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

`--impl reference` runs the UNMODIFIED reference (oracle/_ref, the sha256-pinned copy of /root/reference/mdgen made by
oracle/vendor_reference.py) through its own sampler API on the SAME GPU in strict fp32 - the denominator of the
north-star's ">= 10x the reference single-GPU PyTorch" - on a bounded sample (16 of the 64 trajectories per step, all 100
Euler steps: measured time, nothing extrapolated), and reports the reference's CPU path beside it (`cpu_baseline`).
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
REF_BATCH = 16     # trajectories per step of the reference arm (bounded sample of the 64-trajectory workload)


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
    p.add_argument("--no-torch-gpu", action="store_true", help="skip timing the reference on the GPU")
    p.add_argument("--no-extra", action="store_true", help="skip the extra configurations (ATLAS-shaped, upsampling, strong scaling)")
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


def load_profile_facts():
    """ncu-derived facts of the dominant kernel (DRAM bytes per launch, limiter), committed under profiles/ by
    tools/ncu_summary.py from the raw export of the same build; bench.py never hard-codes them."""
    path = os.path.join(ROOT, "profiles", "roofline_facts.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


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
# reference legs (the only places that touch oracle/)
def _bench_args_kw(T, L, **extra):
    kw = dict(sim_condition=True, prepend_ipa=True, abs_pos_emb=(L == 4), crop=L, num_frames=T,
              sampling_method="euler")
    kw.update(extra)
    return kw


class RefModel:
    """The reference's own NewMDGenWrapper (oracle/_ref) or, when that copy is absent, the oracle port; strict fp32."""

    def __init__(self, T, L, device, stress=False):
        import torch
        from mdgen_b200.config import config_from_args, default_args
        from mdgen_b200.synthetic import synthetic_state_dict
        from oracle import ref_loader
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.device = device
        self.cfg = config_from_args(default_args(**_bench_args_kw(T, L)))
        sd = synthetic_state_dict(self.cfg, seed=0, stress=stress)
        self.kind = "reference" if ref_loader.reference_available() else "port"
        if self.kind == "reference":
            self.source = ref_loader.reference_kind()
            self.m = ref_loader.reference_wrapper(ref_loader.make_args(**_bench_args_kw(T, L)), sd, device)
            self.eigh = ref_loader.chunked_eigh_patch
        else:
            from oracle import mdgen_oracle as O
            self.O = O
            self.sd = {k: v.to(device) for k, v in sd.items()}

    def prep(self, batch):
        """batch on self.device -> opaque conditioning (the reference's own prep_batch)."""
        import torch
        with torch.no_grad():
            if self.kind == "reference":
                with self.eigh():
                    return self.m.prep_batch(batch)["model_kwargs"]
            op = self.O.prep_batch(self.cfg, batch)
            return dict(mask=op["mask"], start=op["start"], end=op["end"], x_cond=op["x_cond"],
                        x_cond_mask=op["x_cond_mask"], aatype=op["aatype"])

    def forward(self, x, t, kw):
        import torch
        with torch.no_grad():
            if self.kind == "reference":
                return self.m.model.forward_inference(x, t, **kw)
            return self.O.forward(self.sd, self.cfg, x, t, **kw)

    def sample(self, zs, K, kw, k_run=None):
        """Fixed-grid Euler through the reference's sampler API; k_run < K integrates only the first k_run steps."""
        import torch
        from functools import partial
        from mdgen_b200.synthetic import euler_time_grid
        with torch.no_grad():
            if self.kind == "reference" and (k_run is None or k_run == K):
                fn = self.m.transport_sampler.sample_ode(sampling_method="euler", num_steps=K + 1)
                return fn(zs, partial(self.m.model.forward_inference, **kw))[-1]
            grid = euler_time_grid(K)[: (k_run or K) + 1].to(zs.device)
            if self.kind == "reference":     # same Euler recurrence the torchdiffeq stand-in runs, on a grid prefix
                f = partial(self.m.model.forward_inference, **kw)
                x = zs
                for i in range(len(grid) - 1):
                    x = x + (grid[i + 1] - grid[i]) * f(x, grid[i] * torch.ones(x.shape[0], device=x.device))
                return x
            return self.O.sample_euler(self.sd, self.cfg, zs, grid, **kw)


def cpu_reference_sample(T, L, K, k_sample=2):
    """The reference's CPU path on a bounded sample: 1 trajectory, k_sample of the K Euler steps, all host threads
    that help (torch CPU ops on this path stop scaling well before 128 threads). Returns (frames/s, cores, text)."""
    import torch
    from mdgen_b200.synthetic import synthetic_batch, synthetic_noise
    ncpu = os.cpu_count() or 1
    ref = RefModel(T, L, "cpu")
    batch = synthetic_batch(1, T, L, seed=1, vary_frames=False)
    zs = synthetic_noise(1, T, L, ref.cfg.latent_dim, seed=2)
    kw = ref.prep(batch)
    best, best_t = None, float("inf")
    for c in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(c)
        ref.forward(zs, torch.zeros(1), kw)
        t0 = time.perf_counter()
        ref.forward(zs, torch.zeros(1), kw)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    t0 = time.perf_counter()
    ref.sample(zs, K, kw, k_run=k_sample)
    per_step = (time.perf_counter() - t0) / k_sample
    text = (f"{'unmodified reference (oracle/_ref)' if ref.kind == 'reference' else 'oracle port'} on the host CPU: "
            f"B=1 x T={T} x L={L}, {k_sample} of {K} Euler steps timed ({per_step:.2f} s/step, best of "
            f"8/16/32/64/all threads), scaled to {K} steps")
    return T / (per_step * K), best, ref.kind, text


def run_reference(a, rank, world):
    """--impl reference (rank 0 only): the unmodified reference on this GPU, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from mdgen_b200.synthetic import synthetic_batch, synthetic_noise
    T, L, K = a.frames, a.residues, a.euler_steps
    line = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0}
    cpu_v, cores, kind, cpu_text = cpu_reference_sample(T, L, K)
    line["cpu_baseline"] = {"value": cpu_v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": cpu_text}
    if torch.cuda.is_available():
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        ref = RefModel(T, L, dev)
        Bs = min(REF_BATCH, a.batch)
        batch = {k: v.to(dev) for k, v in synthetic_batch(Bs, T, L, seed=1, vary_frames=False).items()}
        zs = synthetic_noise(Bs, T, L, ref.cfg.latent_dim, seed=2).to(dev)
        kw = ref.prep(batch)
        clocks = ClockSampler(dev.index or 0)
        for _ in range(a.warmup):
            ref.sample(zs, K, kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            ref.sample(zs, K, kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        value = Bs * T / (ms / 1e3)
        sample = (f"{'unmodified reference (oracle/_ref: ' + ref.source + ')' if ref.kind == 'reference' else 'oracle port'}"
                  f" on cuda:{dev.index or 0}, strict fp32 (allow_tf32 off): {Bs} of the {a.batch} trajectories per step, "
                  f"all {K} Euler steps through its own sample_ode('euler') - measured, not extrapolated")
        line.update({"value": value, "ms_per_step": ms, "clocks": clocks.stop(),
                     "config": {"workload": f"tetrapeptide forward-sim num_frames={T} crop={L}, {K} Euler steps "
                                            f"(BASELINE.json configs[1])", "sample": sample, "device": "cuda",
                                "reference_kind": ref.kind},
                     "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:   # no GPU: the CPU sample is all there is
        line.update({"value": cpu_v, "ms_per_step": 1e3 * T / cpu_v,
                     "config": {"workload": f"tetrapeptide forward-sim num_frames={T} crop={L}, {K} Euler steps "
                                            f"(BASELINE.json configs[1])", "sample": cpu_text, "device": "cpu",
                                "reference_kind": kind},
                     "e2e": {"value": cpu_v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
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
    from mdgen_b200.dist import gather_counts, max_over_ranks, rank_seed, shard_range
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

    def make_model(T_, L_, **extra):
        m_ = NewMDGenWrapper(default_args(**_bench_args_kw(T_, L_, **extra)))
        m_.model.load_state_dict(synthetic_state_dict(m_.cfg, seed=0))
        return m_.eval().to(dev)

    m = make_model(T, L)
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

    # ---- per-kernel device time of one profiled sampling call + one profiled public-API call (CUDA events on the
    #      launch stream, recorded inside the library around every kernel family)
    eng.set_option("profile", 1)
    hot()
    prof = eng.profile_dump()
    db = {k: v.to(dev, non_blocking=True) for k, v in hbatch.items()}
    atom14, _ = m.inference(db, num_steps=2)                 # prep + decode families (2 Euler steps)
    eng.featurize_atom14(atom14.reshape(B * T, L, 14, 3), dbatch["seqres"].repeat_interleave(T, 0))
    prof_api = eng.profile_dump()
    eng.set_option("profile", 0)
    del atom14
    peaks = load_peaks()
    facts = load_profile_facts()
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
    gdt = eng.get_option("gemm_bf16")
    fam_tflops = {k: round(fam_flops[k] / (prof[k][0] / prof[k][1] / 1e3) / 1e12, 1) for k in fam_flops if k in prof}
    # HBM-bound side kernels: algorithmic bytes per token (or residue) per launch / measured time, against the copy peak
    side_bytes = {"ln_mod": 4 * C + 2 * C, "embed": 4 * D + 4 * C + 4 * C, "final": 4 * C + 2 * 4 * D,
                  "mha_l": 2 * 3 * C + 2 * C, "prep": 104 + 2 * 4 * D + 8, "decode": 4 * D + 168,
                  "featurize": 168 + 104 + 28}
    side = {}
    for name, bpt in side_bytes.items():
        src = prof if name in prof else prof_api
        if name in src and src[name][0] > 0:
            gbs = bpt * N / (src[name][0] / src[name][1] / 1e3) / 1e9
            side[name] = {"gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 3),
                          "bytes_per_token": bpt, "ms": round(src[name][0] / src[name][1], 4)}
    fact = facts.get(f"{dom}:{B}x{T}x{L}", {})
    roofline = {
        "kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": fact.get("traffic"), "family_tflops": fam_tflops,
        "note": "mha_t = attn8_prep_kernel + attn8_kernel (fp16 QK^T and P.V on tcgen05); at head_dim 24 there are only "
                "96 MMA FLOP per exponential, so its real limiter is the MUFU ex2 pipe (see `limiter`), not the tensor "
                "pipe; the GEMM families' achieved TFLOP/s are in family_tflops",
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); fp16 operands run at the bf16 rate",
        "avg_launch_ms": dom_ms, "launches": prof[dom][1], "time_shares": shares,
        "limiter": fact.get("limiter"), "facts_source": fact.get("source"),
        "side_kernels": {"peak_gbs": peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})",
                         **side},
        "whole_step_tflops": flops_forward(N, T, L) * K / (ms_per_step / 1e3) / 1e12,
    }

    # ---- extra configurations (SURVEY.md §8d / BASELINE.json configs[2], [4]) and a strong-scaling point, timed after
    #      the headline with one warm-up and one timed sampling call each (all ranks; max over ranks)
    extra = None
    if not a.no_extra:
        extra = {}

        def run_cfg(name, Bx, Tx, Lx, total_traj_x, note, **model_kw):
            try:
                mx = make_model(Tx, Lx, **model_kw)
                bx = {k: v.to(dev) for k, v in synthetic_batch(Bx, Tx, Lx, seed=rank_seed(11, rank), vary_frames=False,
                                                               cond_interval=model_kw.get("cond_interval", 0) or 0).items()}
                zx = synthetic_noise(Bx, Tx, Lx, mx.latent_dim, seed=rank_seed(12, rank)).to(dev)
                kwx = mx.prep_batch(bx)["model_kwargs"]
                fn = lambda: mx.model.sample_euler(zx, grid, **kwx)
                fn()
                msx = timed(fn, 1)
                extra[name] = {"value": total_traj_x * Tx / (msx / 1e3), "unit": "frames/s", "ms_per_step": msx,
                               "config": note, "tflops": flops_forward(Bx * Tx * Lx, Tx, Lx) * K * world / (msx / 1e3) / 1e12}
                del mx, bx, zx, kwx
                torch.cuda.empty_cache()
            except Exception as ex:
                extra[name] = {"error": repr(ex)[:200]}

        run_cfg("atlas_shaped", 1, 250, 256, world, f"ATLAS forward-sim num_frames=250 crop=256 (64,000 tokens), 1 protein per "
                f"GPU, {K} Euler steps (BASELINE.json configs[4] shape), weak")
        run_cfg("upsampling", 16, 1000, 4, 16 * world, f"tetrapeptide upsampling num_frames=1000 cond_interval=100, 16 per GPU, "
                f"{K} Euler steps (BASELINE.json configs[2] shape), weak", cond_interval=100)
        s0, s1 = shard_range(64, rank, world)
        run_cfg("strong_scaling_c2", max(1, s1 - s0), T, L, 64, f"BASELINE.json configs[1] with 64 trajectories TOTAL split over "
                f"{world} rank(s) ({max(1, s1 - s0)} on this rank): strong scaling")

    # ---- the north-star's denominator measured beside ours: the unmodified reference on the SAME GPU (strict fp32),
    #      B trajectories, the first 5 of the K Euler steps of its own recurrence timed; and the parity spot check
    #      on this very batch: forward velocities of the two implementations on identical inputs
    torch_gpu = None
    if rank == 0 and not a.no_torch_gpu:
        try:
            ref = RefModel(T, L, dev)
            rkw = ref.prep(dbatch)
            tq = torch.full((B,), 0.3, device=dev)
            v_ref = ref.forward(zs, tq, rkw)
            v_our = m.model.forward_inference(zs, tq, **kw)
            vel_max = float((v_our - v_ref).abs().max() / v_ref.abs().max())
            vel_l2 = float((v_our - v_ref).norm() / v_ref.norm())
            del v_ref, v_our
            ks = 5
            ref.sample(zs, K, rkw, k_run=1)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            ref.sample(zs, K, rkw, k_run=ks)
            t1.record()
            torch.cuda.synchronize()
            ms_ref = t0.elapsed_time(t1) / ks
            torch_gpu = {"value": B * T / (ms_ref * K / 1e3), "unit": "frames/s", "ms_per_euler_step": ms_ref,
                         "kind": ref.kind,
                         "sample": f"{'unmodified reference (oracle/_ref)' if ref.kind == 'reference' else 'oracle port'} on this "
                                   f"B200, strict fp32: B={B}, {ks} of {K} Euler steps timed, scaled to {K}",
                         "velocity_max_rel_diff_vs_ours": vel_max, "velocity_rel_l2_diff_vs_ours": vel_l2,
                         "velocity_check": f"forward_inference at t=0.3 on the bench batch (B={B}); north-star tolerance 1e-3"}
            del ref, rkw
            torch.cuda.empty_cache()
        except Exception as e:  # e.g. out of memory for the materialised score tensors
            torch_gpu = {"unavailable": repr(e)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cores, kind, text = cpu_reference_sample(T, L, K, k_sample=4)
        cpu_baseline = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": text}

    if rank == 0:
        dtype = {0: "tf32", 1: "bf16", 2: "f16"}.get(gdt, "f16") if use_tc else "f32"
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"tetrapeptide forward-sim num_frames={T} crop={L}, {K} Euler steps, "
                                   f"batch {B} per GPU (BASELINE.json configs[1])",
                       "tokens_per_forward": N, "parallelism": f"dp{world} (independent trajectories, "
                       "no data-path collective)", "l2": "working set (>3 GB activations per forward) "
                       "far exceeds the 126 MB L2; no explicit flush needed",
                       "graph_replays": eng.get_option("graph_replays"),
                       "gemm_path": ("tcgen05, fp32 accumulate: token GEMMs " + {0: "TF32", 1: "bf16", 2: "fp16"}.get(gdt, "?")
                                     + " operands, attention (generation 8) fp16 Q.K^T / P.V; IPA key-frame trunk fp32")
                                    if use_tc else "fp32 SIMT"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "torch_gpu_reference": torch_gpu, "extra_configs": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
