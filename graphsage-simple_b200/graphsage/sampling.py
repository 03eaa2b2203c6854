"""Host-side state of the counter-based sampler.

The reference seeds CPython's global Mersenne Twister once (graphsage/model.py:193) and
every ``random.sample`` call advances it.  The device sampler is stateless: a draw is a pure
function of (seed, step, tag, node id) -- ``seed`` set here, ``step`` advanced once per
top-level forward (one minibatch), ``tag`` identifying which aggregator call of that forward
is drawing (the reference makes three independent draws per SAGE forward)."""

_state = {"seed": 1, "step": 0, "depth": 0, "uid": 0, "step_dev": None}


def seed(s):
    """Analogue of ``random.seed(seed)`` at model.py:193."""
    _state["seed"] = int(s)
    _state["step"] = 0


def get_seed():
    return _state["seed"]


def get_step():
    return _state["step"]


def set_step(step):
    _state["step"] = int(step)


def next_uid():
    _state["uid"] += 1
    return _state["uid"]


class top_level_call:
    """Context manager used by Encoder.forward: the outermost encoder call of a forward pass
    advances the step, nested calls (layer 1 reached through the ``features`` closure) do not."""

    def __enter__(self):
        if _state["depth"] == 0:
            _state["step"] += 1
        _state["depth"] += 1
        return _state["step"]

    def __exit__(self, *exc):
        _state["depth"] -= 1
        return False


def call_tag(uid, call_index):
    return ((int(uid) & 0xFFFFFF) << 8) | (int(call_index) & 0xFF)


class static_step:
    """While active, samplers read the step from the int64 device scalar ``step_dev`` instead of the host counter, so a
    captured CUDA graph of a whole train step draws fresh neighbours on every replay (model.GraphedStep)."""

    def __init__(self, step_dev):
        self.step_dev = step_dev

    def __enter__(self):
        self.prev, _state["step_dev"] = _state["step_dev"], self.step_dev
        return self.step_dev

    def __exit__(self, *exc):
        _state["step_dev"] = self.prev
        return False


def get_step_dev():
    return _state["step_dev"]
