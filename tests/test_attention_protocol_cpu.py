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
