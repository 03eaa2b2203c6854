"""Cross-GPU protocol of the fused all-reduce + SGD kernel (csrc/peer.cu), model-checked on the CPU: per-(rank, CTA)
flags, staging reused every second epoch, stream-ordered launches, remote stores that land late (tools/peer_protocol_sim.py).
The 2- and 8-GPU parity runs (tests/multigpu_dp_check.py) need a multi-GPU box; this runs everywhere."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("peer_protocol_sim", os.path.join(ROOT, "tools", "peer_protocol_sim.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


def _failures(trials, **kw):
    bad = []
    for seed in range(trials):
        try:
            sim.trial(seed, **kw)
        except sim.ProtocolError as e:
            bad.append(str(e))
    return bad


def test_two_phase_exchange_reads_only_this_epochs_data_and_terminates():
    assert _failures(250) == []


def test_eight_ranks_many_epochs():
    for seed in range(10):
        sim.PeerSim(8, 5, 12, sim.random.Random(seed)).run()


@pytest.mark.parametrize("mutation", ["no_flag_b", "early_flag_a"])
def test_model_finds_a_missing_flag(mutation):
    assert _failures(60, mutate=mutation)
