"""Host logic of model.StreamTrainer (the e2e API of bench.py: feed(host ids, host labels) per minibatch, k steps per
graph launch, two pinned buffer pairs used alternately) against a fake engine whose launches complete LAZILY -- only
when somebody synchronises on their event, i.e. a device that is always slower than the host.  Checks the order and
count of the returned losses for stream lengths around the launch size, and that the losses of a launch nobody has
asked for yet survive the reuse of its pinned loss buffer two launches later (model.py:245-252 reports every step's
loss; the reference's loop blocks on each)."""
import numpy as np
import pytest
import torch


class _Clock:
    """Launch log of the fake device: closures run in submission order up to the event somebody waits for."""

    def __init__(self):
        self.pending, self.done = [], 0

    def submit(self, fn):
        self.pending.append(fn)
        return len(self.pending)

    def run_until(self, n):
        while self.done < n:
            self.pending[self.done]()
            self.done += 1


class _FakeEvent:
    clock = None

    def __init__(self, *a, **k):
        self.mark = None

    def record(self, *a):
        self.mark = len(self.clock.pending)

    def synchronize(self):
        self.clock.run_until(self.mark)

    def query(self):
        return self.clock.done >= self.mark


class _Pending:
    def __init__(self, cell, event):
        self.cell, self.event = cell, event

    def ready(self):
        return self.event.query()

    def __float__(self):
        self.event.synchronize()
        return float(self.cell[0])


class _FakeEngine:
    B, slots, trainable_table = 4, 4, False
    nap_after_launches = ()               # launch numbers after which the host is descheduled and the device finishes everything
    launches = 0

    def __init__(self, clock):
        self.clock, self.queue, self.staged = clock, [], []
        self.sets = [type("FS", (), {"stage_dev": torch.zeros(16, dtype=torch.uint8)})()]
        self.trained = []

    @staticmethod
    def loss_of(batch_id):
        return 1000.0 + batch_id

    def reset_pipeline(self):
        self.queue, self.staged = [], []

    def flush_update(self):
        pass

    def pack_stage(self, nodes, labels, step):
        blk = torch.zeros(16, dtype=torch.uint8)
        blk[:8].view(torch.int64)[0] = int(nodes[0])
        return blk

    def push(self, nodes, labels, step, packed=None):
        self.queue.append({"b": packed[1]})
        bid = int(packed[0][:8].view(torch.int64)[0])      # the staging copy is ordered behind the work enqueued so far
        self.clock.submit(lambda: self.staged.append(bid))

    def step_pipelined(self, lr, allreduce=None):
        self.queue.pop(0)
        cell = [None]

        def step():
            bid = self.staged.pop(0)
            self.trained.append(bid)
            cell[0] = self.loss_of(bid)
        self.clock.submit(step)
        self._last = cell

    def read_loss_async(self):
        ev = _FakeEvent()
        ev.record()
        return _Pending(self._last, ev)

    def run_host_queue(self, blocks, losses, lr):
        assert len(self.queue) == 2
        k = blocks.shape[0]

        def launch():                                  # runs when the fake device gets there: reads the pinned blocks NOW
            for j in range(k):
                bid = self.staged.pop(0)
                self.trained.append(bid)
                losses[j] = self.loss_of(bid)
                self.staged.append(int(blocks[j][:8].view(torch.int64)[0]))
        n = self.clock.submit(launch)
        if self.launches in self.nap_after_launches:
            self.clock.run_until(n)
        self.launches += 1


@pytest.fixture
def trainer(monkeypatch):
    from graphsage import model as M
    from graphsage import engine as E
    clock = _Clock()
    _FakeEvent.clock = clock
    eng = _FakeEngine(clock)
    monkeypatch.setattr(E, "engine_for", lambda model, batch: eng)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    fake_model = type("Mdl", (), {"grad_allreduce": None})()
    return M.StreamTrainer(fake_model, lr=0.1, steps_per_launch=4), eng


@pytest.mark.parametrize("count", [1, 2, 3, 4, 6, 7, 8, 11, 12, 15, 16, 23, 40])
def test_losses_come_back_once_each_in_feed_order(trainer, count):
    tr, eng = trainer
    got = []
    for i in range(count):
        got += tr.feed(np.full(4, i, dtype=np.int64), np.zeros(4, dtype=np.int64))
    got += tr.finish()
    assert got == [eng.loss_of(i) for i in range(count)]
    assert eng.trained == list(range(count))


def test_a_pinned_block_is_not_refilled_before_the_launch_that_reads_it_ran(trainer):
    """The fake launch reads its staging blocks when it RUNS; the trainer may only overwrite a buffer pair after
    waiting for the launch that used it -- otherwise the device would train on the wrong batches."""
    tr, eng = trainer
    for i in range(31):
        tr.feed(np.full(4, i, dtype=np.int64), np.zeros(4, dtype=np.int64))
    tr.finish()
    assert eng.trained == list(range(31))


def test_unconsumed_losses_survive_the_reuse_of_their_pinned_buffer(trainer):
    """Device slow while the host feeds (no loss is ready when feed() looks), then the host is descheduled right after a
    launch and the device runs through it: the launch has overwritten the loss buffer of the launch two before it, whose
    losses nobody had read yet.  They must have been read out when the trainer waited for that older launch."""
    tr, eng = trainer
    eng.nap_after_launches = (2, 5)
    got = []
    for i in range(27):
        got += tr.feed(np.full(4, i, dtype=np.int64), np.zeros(4, dtype=np.int64))
    got += tr.finish()
    assert got == [eng.loss_of(i) for i in range(27)]
