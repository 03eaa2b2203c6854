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


# ---- every interleaving, not a sample (tools/tc_protocol_exhaustive.py: breadth-first over the reachable states, no time)
_spec2 = importlib.util.spec_from_file_location("tc_protocol_exhaustive", os.path.join(ROOT, "tools", "tc_protocol_exhaustive.py"))
exh = importlib.util.module_from_spec(_spec2)
_spec2.loader.exec_module(exh)


@pytest.mark.parametrize("mode,n,head", [("all", 8, None), ("head", 8, 3), ("head", 7, 1), ("none", 8, None), ("all", 1, None),
                                         ("all", 2, None), ("all", 3, None), ("head", 4, 2)])
def test_kernel_rules_hold_under_every_interleaving(mode, n, head):
    r = exh.explore(n, exh.mode_fn(mode, n, head), "own_only")
    assert r["ok"] is True, r
    # "kernel" = what gemm_tc.cu does: own_only on refill duty; TMA-only CTAs keep observe_all, which the timeless model
    # cannot prove (it lets a warp idle between two register instructions for thousands of cycles)
    if mode != "none":
        assert exh.explore(n, exh.mode_fn(mode, n, head), "kernel")["ok"] is True


def test_two_warps_per_group_named_barrier_and_skew():
    assert exh.explore(3, exh.mode_fn("all", 3), "kernel", wpg=2)["ok"] is True        # 4+ chunks: tools/, minutes


def test_exhaustive_search_finds_the_round_2_deadlock_with_a_shortest_schedule():
    r = exh.explore(4, exh.mode_fn("all", 4), "observe_all")
    assert r["ok"] is False and "deadlock" in r["why"]
    # the late group is the one that has to issue TWO refills before its first wait (chunks 0 and 2): it is lapped on
    # stage 0, whose next phase (chunk 3) completes without it
    assert any("('wait', 'F', 0, 0)" in w for w in [r["why"]])


def test_exhaustive_search_finds_the_round_1_aliasing():
    r = exh.explore(6, exh.mode_fn("none", 6), "own_naive")
    assert r["ok"] is False and "reads X" in r["why"]


def test_model_reproduces_the_four_stage_history(monkeypatch):
    """Round 1: with FOUR stages each group always meets the same stages, so waiting for own chunks only was sound (0 / 60
    bad runs on the B200) -- it is the odd ring that aliased.  Round 2: four stages + refill duty + observe-everything
    hung on the B200.  The model says the same three things."""
    monkeypatch.setattr(exh, "S", 4)
    assert exh.explore(9, exh.mode_fn("none", 9), "own_naive")["ok"] is True
    assert exh.explore(9, exh.mode_fn("all", 9), "observe_all")["ok"] is False
    assert exh.explore(9, exh.mode_fn("none", 9), "own_only")["ok"] is True
