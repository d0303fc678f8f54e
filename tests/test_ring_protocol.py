"""The tile ring of the fused FC kernel (fc_tcgen05.cu: kFullBars "tile landed" barriers per stage) checked on a discrete-event
model with arbitrary latencies: no parity wait may pass before its tile has landed, and every tile completes.  CPU only.

The rule the kernel states: a shape with `slots` tile slots on `stages` ring stages needs at least slots / gcd(slots, stages)
barriers per stage (the latest same-stage tile that is certain to have landed when the issuer reaches tile j is that many ring
rounds back).  kFullBars = 8 covers every shape (slots <= 8)."""
import re
from math import gcd
from pathlib import Path

import pytest

from ring_protocol_sim import Violation, simulate

# (warpgroups, slots per warpgroup, ring stages): the shapes the plans launch
SHIPPED = [(3, 2, 6), (3, 2, 5),   # width-64 models; their gather launches (staging buffers cost a stage)
           (4, 1, 5),              # four-warpgroup form (2bitsym-96)
           (3, 1, 2), (2, 1, 2),   # shared-memory-activation form (Binary-160: three slots, two stages), 256-wide
           (2, 1, 4), (3, 1, 4),   # wide models on TMEM activations
           (2, 2, 4)]              # float-input launches
FAST_CHAIN = dict(mma=(0.01, 0.05), epi=(0.01, 0.08), load_lat=(0.3, 4.0), slow_load=(0.05, 20.0))   # tiny model, slow loads


def kernel_full_bars():
    src = (Path(__file__).resolve().parents[1] / "bitnetmcu_b200" / "csrc" / "fc_tcgen05.cu").read_text()
    return int(re.search(r"constexpr uint32_t kFullBars = (\d+);", src).group(1))


@pytest.mark.parametrize("shape", SHIPPED)
def test_shipped_shapes_are_safe_with_the_kernels_barrier_count(shape):
    n_wg, slots, stages = shape
    bars = kernel_full_bars()
    assert bars >= n_wg * slots // gcd(n_wg * slots, stages)
    for seed in range(60):
        simulate(n_wg, slots, stages, full_bars=bars, seed=seed)
        simulate(n_wg, slots, stages, full_bars=bars, seed=seed, **FAST_CHAIN)


def test_two_barriers_per_stage_are_only_safe_while_slots_do_not_outnumber_stages():
    # slots <= stages: the wait for tile i + 2 stages comes after its slot's previous tile, whose load was requested after tile i landed
    for seed in range(60):
        simulate(3, 2, 6, full_bars=2, seed=seed, **FAST_CHAIN)
    # the development shape that trapped on the GPU (eight slots on six stages, tiny model, overlapped launches = slow loads)
    failures = 0
    for seed in range(60):
        try:
            simulate(4, 2, 6, full_bars=2, seed=seed, **FAST_CHAIN)
        except Violation:
            failures += 1
    assert failures > 0
    # ... and what the gather launches and the shared-memory-activation form would have risked with a very late load
    with pytest.raises(Violation):
        for seed in range(400):
            simulate(3, 1, 2, full_bars=2, seed=seed, **FAST_CHAIN)


def test_barriers_needed_is_slots_over_gcd():
    slow = dict(FAST_CHAIN, slow_load=(0.05, 40.0))
    for n_wg, slots, stages in [(3, 2, 5), (4, 1, 5), (3, 2, 4), (4, 2, 7), (4, 2, 6), (3, 2, 3)]:
        need = n_wg * slots // gcd(n_wg * slots, stages)
        for seed in range(40):   # enough barriers: never a violation
            simulate(n_wg, slots, stages, full_bars=need, seed=seed, n_tiles=90, **slow)
        if need > 1:             # too few: the adversarial latencies find it
            with pytest.raises(Violation):
                for seed in range(300):
                    simulate(n_wg, slots, stages, full_bars=max(1, need // 2), seed=seed, n_tiles=90, **slow)
