"""Host logic of SupervisedGraphSage.train_step's prefetch protocol (model.py of the drop-in package; the unit is the
reference's loop body, graphsage/model.py:245-250) against a fake engine: whatever the caller announces -- the right
next batches, wrong ones, none -- every call must train exactly the batch it was given, each batch is staged once when
the announcements were right, and a wrong announcement costs a re-stage, never a wrong step."""
import numpy as np
import pytest


class _FakeEngine:
    depth, trainable_table = 3, False

    def __init__(self):
        self.queue, self.trained, self.staged, self.resets, self.pushes = [], [], None, 0, 0

    def stage(self, nodes, labels, step):
        self.staged = (list(map(int, nodes)), list(map(int, labels)), step)
        return len(nodes)

    def train_step(self, b, lr, allreduce=None):
        assert not self.queue
        self.trained.append(self.staged)

    def read_loss(self):
        return float(len(self.trained))

    read_loss_async = read_loss

    def reset_pipeline(self):
        self.resets += 1
        self.queue = []

    def drop_queued(self, keep):
        del self.queue[keep:]

    def push(self, nodes, labels, step, **kw):
        assert len(self.queue) < self.depth
        self.pushes += 1
        self.queue.append({"ids": np.array(nodes, dtype=np.int64), "labels": list(map(int, labels)), "step": step, "b": len(nodes)})

    def step_pipelined(self, lr, allreduce=None):
        e = self.queue.pop(0)
        self.trained.append((e["ids"].tolist(), e["labels"], e["step"]))


@pytest.fixture
def step_fn(monkeypatch):
    from graphsage import engine as E
    from graphsage import model as M
    from graphsage import sampling
    eng = _FakeEngine()
    monkeypatch.setattr(E, "engine_for", lambda model, batch: eng)
    sampling.seed(1)
    me = type("Mdl", (), {"grad_allreduce": None})()
    return (lambda nodes, labels, **kw: M.SupervisedGraphSage.train_step(me, nodes, labels, lr=0.5, **kw)), eng


def _batches(n, size=3):
    return [(np.arange(size) + 10 * i, np.full(size, i)) for i in range(n)]


def _ids(trained):
    return [t[0] for t in trained]


def test_plain_calls_train_their_own_batch(step_fn):
    step, eng = step_fn
    bs = _batches(5)
    for nodes, labels in bs:
        step(nodes, labels)
    assert _ids(eng.trained) == [b[0].tolist() for b in bs] and eng.pushes == 0


@pytest.mark.parametrize("ahead", [1, 2])
def test_right_announcements_stage_every_batch_once(step_fn, ahead):
    step, eng = step_fn
    bs = _batches(9)
    for i, (nodes, labels) in enumerate(bs):
        step(nodes, labels, prefetch=bs[i + 1:i + 1 + ahead], sync=False)
    assert _ids(eng.trained) == [b[0].tolist() for b in bs]
    assert [t[1] for t in eng.trained] == [b[1].tolist() for b in bs]
    assert eng.pushes == len(bs) and eng.resets <= 1
    assert not eng.queue                                           # the tail drained: nothing announced, nothing queued
    steps = [t[2] for t in eng.trained]
    assert steps == sorted(steps) and len(set(steps)) == len(steps)  # one sampler step per batch, in order


def test_wrong_announcements_never_train_a_wrong_batch(step_fn):
    step, eng = step_fn
    bs = _batches(8)
    decoy = (np.array([777, 778, 779]), np.zeros(3, dtype=np.int64))
    for i, (nodes, labels) in enumerate(bs):
        nxt = bs[i + 1:i + 3]
        if i in (2, 5) and nxt:
            nxt = [decoy] + nxt[1:]                                 # announces a batch that never comes
        if i == 4:
            nxt = nxt[::-1]                                         # announces the right batches in the wrong order
        step(nodes, labels, prefetch=nxt, sync=False)
    assert _ids(eng.trained) == [b[0].tolist() for b in bs]
    assert [t[1] for t in eng.trained] == [b[1].tolist() for b in bs]


def test_a_single_pair_is_accepted_as_prefetch(step_fn):
    step, eng = step_fn
    bs = _batches(4)
    for i, (nodes, labels) in enumerate(bs):
        step(nodes, labels, prefetch=bs[i + 1] if i + 1 < len(bs) else None)
    assert _ids(eng.trained) == [b[0].tolist() for b in bs]
