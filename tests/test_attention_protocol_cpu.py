"""mbarrier protocol of the tcgen05 attention kernels under randomised timings (tests/protocol_sim.py):
no parity wait may be released for a phase that has not completed, nothing deadlocks, S / P / O hazards hold."""
import pytest

from tests.protocol_sim import Violation, attention_cta


@pytest.mark.parametrize("nkt", [1, 2, 3, 4, 11])
def test_single_item_kernel_protocol(nkt):
    for seed in range(40):
        assert attention_cta(seed, nkt, n_items=1, slow_warp=seed % 8) == []


@pytest.mark.parametrize("nkt,items", [(1, 5), (2, 4), (3, 4), (5, 3), (11, 2)])
def test_persistent_kernel_protocol(nkt, items):
    for seed in range(25):
        assert attention_cta(100 + seed, nkt, n_items=items, slow_warp=seed % 8) == []
        assert attention_cta(200 + seed, nkt, n_items=items, nsw=12, slow_warp=seed % 12) == []


def test_single_o_valid_barrier_is_caught():
    """Negative control: with ONE 'O valid' barrier advancing once per key tile (the first version of the
    kernel) a softmax warp that runs a full tile ahead of the slowest one is released too early or blocks."""
    found = 0
    for seed in range(60):
        try:
            v = attention_cta(seed, 6, n_items=2, single_odone=True, slow_warp=seed % 8)
        except Violation:
            v = ["deadlock"]
        found += bool(v)
    assert found > 0


@pytest.mark.parametrize("nkt,items,nq", [(1, 4, 2), (2, 3, 2), (3, 3, 4), (5, 2, 4), (11, 2, 2), (21, 2, 4)])
def test_generation8_kernel_protocol(nkt, items, nq):
    """attn8_kernel<NQ>: per-tile issuer threads, shared K/V ring, per-parity P.V-done barriers."""
    from tests.protocol_sim import attention8_cta
    for seed in range(12):
        slow = (seed % nq, seed % 4) if seed % 3 else None
        assert attention8_cta(300 + seed, nkt, n_items=items, nq=nq, slow_warp=slow) == []
        assert attention8_cta(400 + seed, nkt, n_items=items, nq=nq, stages=3, slow_warp=slow) == []


def test_generation8_single_pv_done_barrier_is_caught():
    """Negative control: the race found on hardware (deviations of 1e-4 .. 1e-2 in ~1 % of the items). With S double
    buffered a softmax warp finishes tile c before P.V(c-1) retires; one P.V-done barrier per query tile advancing once
    per key tile then releases a parity wait two phases early (or blocks it forever)."""
    from tests.protocol_sim import attention8_cta
    found = 0
    for seed in range(60):
        try:
            v = attention8_cta(seed, 6, n_items=2, nq=2, single_pvdone=True, slow_warp=(seed % 2, seed % 4))
        except Violation:
            v = ["deadlock"]
        found += bool(v)
    assert found > 0


@pytest.mark.parametrize("tiles,kblocks,stages,resident", [(1, 6, 4, False), (5, 6, 4, False), (7, 24, 4, False),
                                                           (6, 6, 3, True), (9, 1, 3, True), (4, 3, 2, True)])
def test_gemm_pipeline_protocol(tiles, kblocks, stages, resident):
    """gemm_tc_kernel (tile streaming) and gemm_tc_ws_kernel (weight stationary): operand ring, TMEM double buffer with the
    accumulator handed back right after the last tcgen05.ld."""
    from tests.protocol_sim import gemm_cta
    for seed in range(15):
        assert gemm_cta(500 + seed, tiles, kblocks, stages=stages, resident_b=resident) == []
        assert gemm_cta(600 + seed, tiles, kblocks, stages=stages, resident_b=resident, early_release=False) == []


def test_gemm_wrong_tempty_count_is_caught():
    """Negative control: if the 'accumulator drained' barrier expected fewer arrivals than there are epilogue warps, the MMA
    thread would overwrite a buffer that some warp has not read yet."""
    from tests.protocol_sim import gemm_cta
    found = 0
    for seed in range(40):
        try:
            v = gemm_cta(seed, 8, 2, stages=4, n_epi=12, tempty_count=6)
        except (Violation, AssertionError):
            v = ["deadlock"]
        found += bool(v)
    assert found > 0
