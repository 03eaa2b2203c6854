"""Stream / event choreography of the pipelined engine (engine.TrainEngine.push, step_pipelined, run_device_queue,
reset_pipeline), checked on the CPU as a happens-before graph.

torch.cuda's streams and events are replaced by fakes that record every enqueued operation as a node and every
wait_stream / wait_event / fork / join as an edge; the engine's launch chains are replaced by loggers.  The test then
asks the questions a race on the GPU would answer only occasionally: for every minibatch, is staging ordered before
its sampler chain, that before its gather, that before its compute chain?  Is a frontier set restaged only after the
compute chain that last read it?  Are the sampler chains (one shared dedup scratch), the compute chains and the weight
updates totally ordered, each forward after the previous update?  Three minibatches are in flight in four frontier
sets on four streams; the reference's loop (graphsage/model.py:245-250) has none of this to get wrong."""
import contextlib

import pytest
import torch


class _World:
    def __init__(self):
        self.preds = []                 # node id -> list of predecessor node ids
        self.info = []
        self.stack = []

    def node(self, preds, info=None):
        self.preds.append([p for p in preds if p is not None])
        self.info.append(info)
        return len(self.preds) - 1

    def before(self, a, b):
        """a happens-before b"""
        seen, todo = set(), [b]
        while todo:
            n = todo.pop()
            if n == a:
                return True
            for p in self.preds[n]:
                if p not in seen and p >= a:          # node ids grow along every edge
                    seen.add(p)
                    todo.append(p)
        return False


W = None


class _Stream:
    def __init__(self, *a, **k):
        self.last = None

    def enqueue(self, info):
        self.last = W.node([self.last], info)
        return self.last

    def wait_stream(self, other):
        self.last = W.node([self.last, other.last])

    def wait_event(self, ev):
        self.last = W.node([self.last, ev.node])

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        self.node = None

    def record(self, stream=None):
        self.node = (stream or W.stack[-1]).last

    def synchronize(self):
        pass

    def query(self):
        return True


@contextlib.contextmanager
def _stream_ctx(s):
    W.stack.append(s)
    try:
        yield
    finally:
        W.stack.pop()


class _Set:
    def __init__(self, i):
        self.i = i
        self.stage_dev = torch.zeros(16, dtype=torch.uint8)


@pytest.fixture
def engine(monkeypatch):
    global W
    from graphsage import engine as E
    W = _World()
    main = _Stream()
    W.stack.append(main)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: W.stack[-1])
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "stream", _stream_ctx)
    log = []

    def op(kind, slot):
        n = W.stack[-1].enqueue((kind, slot))
        log.append((kind, slot, n))

    class Eng(E.TrainEngine):
        def __init__(self):                      # no device buffers: only what the pipeline's host logic touches
            self.sets = [_Set(i) for i in range(self.slots)]
            self.cur, self.queue, self.B, self.dev = 0, [], 8, "cpu"
            self.use_graphs, self._pending_lr, self.peer = False, None, None
            self._side = self._gstream = self._sstream = self._cstream = None
            self._graphs, self._warm, self._launch_count = {}, set(), {}

        def _sample_chain(self, fs, b):
            op("sample", fs.i)

        def _gather(self, fs, b):
            op("gather", fs.i)

        def _compute_chain(self, fs, b):
            op("compute", fs.i)

        def _update(self, lr):
            op("update", None)

        def stage_packed(self, block, b, slot=None):
            op("stage", self.cur if slot is None else slot)
            return b

    monkeypatch.setattr(E.ops, "stage_next", lambda pool, cursor, dst: op("stage", [s.i for s in eng.sets if s.stage_dev is dst][0]))
    eng = Eng()
    return eng, log


def _batches(log, slots=4):
    """Group the logged operations by minibatch: the j-th stage / sample / gather / compute on a slot belong together."""
    per = {}
    count = {}
    for kind, slot, n in log:
        if kind == "update":
            continue
        j = count.get((kind, slot), 0)
        count[(kind, slot)] = j + 1
        per.setdefault((slot, j), {})[kind] = n
    return per


def _check(log, expect_batches):
    per = _batches(log)
    done = {k: v for k, v in per.items() if "compute" in v}
    assert len(done) == expect_batches
    for key, ops_ in done.items():
        assert set(ops_) == {"stage", "sample", "gather", "compute"}, (key, ops_)
        assert W.before(ops_["stage"], ops_["sample"]) and W.before(ops_["sample"], ops_["gather"]) \
            and W.before(ops_["gather"], ops_["compute"]), key
    for (slot, j), ops_ in per.items():              # a set is restaged only after the compute chain that last read it
        if j > 0 and "stage" in ops_ and (slot, j - 1) in done:
            assert W.before(done[(slot, j - 1)]["compute"], ops_["stage"]), (slot, j)
        if j > 0 and "stage" in ops_ and (slot, j - 1) in per and "gather" in per[(slot, j - 1)]:
            assert W.before(per[(slot, j - 1)]["gather"], ops_["stage"]), (slot, j)
    for kind in ("sample", "compute"):               # shared scratch / shared activations: totally ordered
        seq = [n for k, s, n in log if k == kind]
        assert all(W.before(a, b) for a, b in zip(seq, seq[1:])), kind
    # every forward sees the weights of all earlier updates, and an update follows the backward whose gradients it applies
    seq = [(k, n) for k, s, n in log if k in ("compute", "update")]
    assert all(W.before(a[1], b[1]) for a, b in zip(seq, seq[1:]))
    computes = sum(1 for k, n in seq if k == "compute")
    updates = sum(1 for k, n in seq if k == "update")
    assert computes == updates == expect_batches
    kinds = [k for k, n in seq]
    for i, k in enumerate(kinds):                    # never two forwards without the update between them
        if k == "compute" and i + 1 < len(kinds):
            assert kinds[i + 1] == "update"


def _feed(eng, n):
    """bench.py's device_step loop: two batches queued, then push one + step per iteration, then the drain."""
    blk = torch.zeros(16, dtype=torch.uint8)
    eng.push(None, None, None, packed=(blk, 8))
    eng.push(None, None, None, packed=(blk, 8))
    for _ in range(n - 2):
        eng.push(None, None, None, packed=(blk, 8))
        eng.step_pipelined(0.1)
    while eng.queue:
        eng.step_pipelined(0.1)
    eng.flush_update()


@pytest.mark.parametrize("n", [2, 3, 4, 5, 9, 14])
def test_one_launch_per_step_pipeline(engine, n):
    eng, log = engine
    _feed(eng, n)
    _check(log, n)


def test_k_steps_per_launch_between_single_steps(engine):
    eng, log = engine
    blk = torch.zeros(16, dtype=torch.uint8)
    pool = torch.zeros((32, 16), dtype=torch.uint8)
    cursor = torch.zeros(1, dtype=torch.int64)
    eng.push(None, None, None, packed=(blk, 8))
    eng.push(None, None, None, packed=(blk, 8))
    for _ in range(3):                               # reach the steady state with single steps
        eng.push(None, None, None, packed=(blk, 8))
        eng.step_pipelined(0.1)
    eng.run_device_queue(pool, cursor, 8, 0.1)       # 8 steps as one launch, batches staged by gs_stage_next
    eng.run_device_queue(pool, cursor, 4, 0.1)
    for _ in range(2):                               # back to single steps: pushes after a multi-step launch
        eng.push(None, None, None, packed=(blk, 8))
        eng.step_pipelined(0.1)
    while eng.queue:
        eng.step_pipelined(0.1)
    eng.flush_update()
    _check(log, 3 + 8 + 4 + 2 + 2)


def test_reset_drops_queued_batches_without_racing_their_staging_copies(engine):
    eng, log = engine
    blk = torch.zeros(16, dtype=torch.uint8)
    for _ in range(3):
        eng.push(None, None, None, packed=(blk, 8))
    eng.step_pipelined(0.1)
    mark = len(log)
    eng.reset_pipeline()                             # two batches were staged / sampled / gathered ahead and are dropped
    _feed(eng, 5)
    # everything enqueued after the reset is ordered after everything enqueued before it (the dropped batches' copies,
    # sampler chains and gathers included): the sets are reused from slot 0 on
    before = [n for k, s, n in log[:mark]]
    after_first = {}
    for k, s, n in log[mark:]:
        after_first.setdefault((k, s), n)
    for (k, s), n in after_first.items():
        if k in ("stage", "sample", "gather"):
            assert all(W.before(b, n) for b in before), (k, s)
