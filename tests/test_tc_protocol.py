"""Synchronisation protocol of the tcgen05 encoder GEMM (csrc/gemm_tc.cu), model-checked on the CPU.

The kernel's hazards are timing-dependent and the build container has no GPU, so the wait rules of its splitter groups
are checked against a discrete-event model of the kernel's mbarriers (tools/tc_protocol_sim.py) under random schedules
with rare very long stalls: no read before the data has landed, no overwrite under a reader, no surplus arrival, and
termination.  The model is also shown to FIND the two bugs the kernel has had (parity aliasing, round 1; a late
warp lapped by the full barrier on refill duty, round 2), so a green run means something."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("tc_protocol_sim", os.path.join(ROOT, "tools", "tc_protocol_sim.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


def _failures(rule, mode, trials, **kw):
    bad = []
    for seed in range(trials):
        try:
            sim.trial(seed, rule, mode, **kw)
        except sim.ProtocolError as e:
            bad.append(str(e))
    return bad


@pytest.mark.parametrize("mode", ["none", "all", "head"])
def test_kernel_rule_survives_every_schedule(mode):
    # own_only is what gemm_tc.cu does in CTAs on refill duty (mode all = weight gradient's self tiles, head = forward);
    # it is also sound for TMA-only CTAs (mode none), which keep the hardware-proven observe-all rule
    assert _failures("own_only", mode, 400, stall_p=0.05) == []


def test_kernel_rule_at_the_bench_shape():
    # layer 1 of the Reddit-shape bench: forward 38 chunks (19 gathered), weight gradient <= 32 chunks per split
    for seed in range(100):
        s = sim.Sim(38, lambda c: c < 19, "own_only", sim.random.Random(seed), stall_p=0.05)
        s.start()
        s.run()
        s = sim.Sim(28, lambda c: True, "own_only", sim.random.Random(seed), stall_p=0.05)
        s.start()
        s.run()


def test_observe_all_is_sound_without_refill_duty():
    assert _failures("observe_all", "none", 300) == []


@pytest.mark.parametrize("mode", ["all", "head"])
def test_model_finds_the_lapping_deadlock_of_round_2(mode):
    bad = _failures("observe_all", mode, 200)
    assert bad and all("deadlock" in b for b in bad)


def test_model_finds_the_parity_aliasing_of_round_1():
    bad = _failures("own_naive", "none", 200)
    assert bad and any("reads X" in b for b in bad)
