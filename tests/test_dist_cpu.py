"""N>1 host logic on CPU: world_size-2 gloo processes (the GPU path uses the same helpers with NCCL)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdgen_b200.dist import gather_counts, max_over_ranks, rank_seed, shard_range


def test_shard_range_partitions():
    for total in (1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            covered = []
            for a, b in spans:
                covered += list(range(a, b))
            assert covered == list(range(total))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert len({rank_seed(5, r) for r in range(8)}) == 8


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_range(65, rank, world)
        # each rank "samples" its shard with its own seed; results must be disjoint and complete
        torch.manual_seed(rank_seed(2, rank))
        local_ms = 10.0 + 5.0 * rank            # rank 1 is the slow one
        slow = max_over_ranks(local_ms)
        counts = gather_counts(b - a)
        q.put((rank, a, b, slow, counts))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_max_time():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, s0, c0), (r1, a1, b1, s1, c1) = res
    assert (a0, b0, a1, b1) == (0, 33, 33, 65)
    assert s0 == s1 == 15.0                     # max over ranks, identical on both
    assert c0 == c1 == [33, 32]


def test_reference_arm_under_torchrun_prints_one_line():
    """bench.py --impl reference launched the way the driver launches it for N > 1: rank 0 alone runs the reference
    (here, without a GPU, its CPU path on a bounded sample) and prints ONE JSON line with the contract's keys; the
    other rank exits 0 silently."""
    import json
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), "bench.py", "--impl", "reference",
           "--gpus", "2", "--steps", "1", "--warmup", "0", "--frames", "24", "--euler-steps", "4"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "frames/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["device"] == "cpu" and d["ms_per_step"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
