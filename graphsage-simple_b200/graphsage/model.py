"""Drop-in for the hot-path classes of graphsage/model.py of zjzijielu/graphsage-simple:
``SupervisedGraphSage`` (model.py:52-69).  ``forward`` returns ``[n, num_classes]`` scores,
``loss`` the mean cross-entropy; both run on the CUDA kernels.  When the module tree is the
canonical 2-layer wiring of model.py:214-227 over a frozen feature table, ``loss`` routes the
whole sample -> aggregate -> update step through the fused engine (engine.py) instead of the
op-by-op autograd path; both produce the same numbers."""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from . import ops
from .aggregators import _device
from .functional import EncoderGemm, SoftmaxXent


def build_sage(features, feat_dim, hidden, adj_lists, fanouts, num_classes, gcn=False, agg_gcn=False,
               initializer="None", feature_dim=100, num_nodes=100):
    """Wire an L-layer supervised GraphSAGE exactly the way the reference's driver does for two
    layers (model.py:218-227): layer l's ``features`` is the closure ``lambda nodes: enc_{l-1}(nodes).t()``.

    hidden   -- output width per layer, innermost first (model.py:219, 221: [identity_dim, 128])
    fanouts  -- ``num_sample`` per layer, innermost first (the reference's nominal (num_sample1,
                num_sample2), model.py:188-189, which its driver sets on an attribute nobody reads)
    adj_lists-- mapping int -> set (model.py:303-310), a CSRGraph, or a sharded.ShardedCSR
    Returns (SupervisedGraphSage, [enc_1 .. enc_L])."""
    from .aggregators import MeanAggregator
    from .encoders import Encoder
    encs = []
    prev_feats, prev_dim = features, feat_dim
    for layer, (dim, k) in enumerate(zip(hidden, fanouts)):
        if layer == 0:
            agg = MeanAggregator(prev_feats, initializer, cuda=True, gcn=agg_gcn, feature_dim=feature_dim,
                                 num_nodes=num_nodes)
            enc = Encoder(prev_feats, prev_dim, dim, adj_lists, agg, num_sample=k, initializer=initializer, gcn=gcn,
                          cuda=True)
        else:
            below = encs[-1]
            feats = (lambda b: (lambda nodes: b(nodes).t()))(below)          # model.py:220-221
            agg = MeanAggregator(feats, cuda=True, gcn=agg_gcn)
            enc = Encoder(feats, below.embed_dim, dim, adj_lists, agg, num_sample=k, base_model=below, gcn=gcn, cuda=True)
        encs.append(enc)
    return SupervisedGraphSage(num_classes, encs[-1]), encs


class SGD:
    """``torch.optim.SGD(params, lr)`` of model.py:237 (no momentum, no weight decay) on gs_sgd_step: one launch per
    parameter, ``p = p - lr * grad`` with the same two roundings as ``p.add_(grad, alpha=-lr)``."""

    def __init__(self, params, lr=0.7):
        self.params = [p for p in params if p.requires_grad]
        self.lr = float(lr)

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        for p in self.params:
            if p.grad is None:
                continue
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if p.data.is_contiguous():
                ops.sgd_step(p.data.view(-1), g.view(-1), self.lr)
            else:                       # a view into a padded parameter block: update a packed copy, write it back
                w = p.data.contiguous()
                ops.sgd_step(w.view(-1), g.view(-1), self.lr)
                p.data.copy_(w)


class SupervisedGraphSage(nn.Module):

    def __init__(self, num_classes, enc):
        super().__init__()
        self.enc = enc
        self.xent = nn.CrossEntropyLoss()
        self.weight = nn.Parameter(torch.empty(num_classes, enc.embed_dim, device=_device()))
        init.xavier_uniform_(self.weight)                          # model.py:59-60

    def forward(self, nodes):
        embeds = self.enc(nodes)                                   # [d, n]
        return EncoderGemm.apply(embeds.t(), self.weight, ops.ACT_NONE)   # == weight.mm(embeds).t()

    def _labels(self, labels):
        if isinstance(labels, torch.Tensor):
            return labels.to(device=_device(), dtype=torch.int64).reshape(-1).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(labels, dtype=np.int64)).reshape(-1)).to(_device())

    use_engine = None       # None: automatic (fused engine when the wiring is canonical); False: never

    def loss(self, nodes, labels):
        """model.py:67-69.  With the canonical 2-layer wiring and autograd enabled the whole
        step runs through the fused engine (engine.py) -- same kernels, same sampler draws as
        the op-by-op path below -- and the returned scalar carries the gradients."""
        if self.use_engine is not False and torch.is_grad_enabled():
            from .engine import engine_for, _EngineLoss
            eng = engine_for(self, len(nodes))
            if eng is not None:
                from . import sampling
                with sampling.top_level_call() as step:
                    b = eng.stage(nodes, labels, step)
                    eng.forward_backward(b)
                enc2 = self.enc
                params = [self.weight, enc2.weight, enc2.base_model.weight]
                grads = [eng.gwc, eng.gw2, eng.gw1]
                if eng.trainable_table:                                # 1hot / node_degree: aggregators.py:30-31
                    params.append(enc2.base_model.aggregator.embed.weight)
                    grads.append(eng.gembed)
                return _EngineLoss.apply(eng.loss, len(params), *params, *grads)
        embeds = self.enc(nodes)
        return SoftmaxXent.apply(embeds.t(), self.weight, self._labels(labels))

    grad_allreduce = None   # data parallel: callable(flat_grads) run between backward and the SGD step

    def stream_trainer(self, lr=0.7, steps_per_launch=4):
        """A ``StreamTrainer`` bound to this model: ``feed(nodes, labels)`` per minibatch, ``finish()`` at the end."""
        return StreamTrainer(self, lr, steps_per_launch)

    def train_step(self, nodes, labels, lr=0.7, prefetch=None, sync=True):
        """The reference's timed unit (model.py:246-250: zero_grad, loss, backward, SGD step) as
        one fused call; returns the loss as a Python float (one 4-byte device->host read).  With
        ``sync=False`` the read is enqueued and an ``engine.PendingLoss`` is returned: ``float()`` of it
        waits for that step only, so the host can launch the next step before consuming this loss
        (at most 8 may be outstanding).

        ``prefetch`` names the minibatches that FOLLOW this one -- ``(nodes, labels)`` or a list of
        up to two such pairs -- so that their neighbour sampling and feature gather run on side
        streams while this batch is in its GEMM/backward chain (the data-loader style overlap the
        reference's single Python thread cannot do).  Pass the same batches, in order, as
        ``nodes, labels`` of the following calls; results equal plain sequential steps.  While batches
        are announced the engine is in streaming mode: the SGD update of a step is enqueued at the head
        of the next step, so read the weights after the last batch (no ``prefetch``) or after
        ``model._engine.flush_update()``."""
        from .engine import engine_for
        from . import sampling
        upcoming = [] if prefetch is None else ([prefetch] if isinstance(prefetch, tuple) else list(prefetch))
        eng = engine_for(self, max([len(nodes)] + [len(u[0]) for u in upcoming]))
        if eng is None:
            raise RuntimeError("train_step needs the canonical 2-layer wiring (model.py:214-227)")
        if eng.trainable_table:
            upcoming = []          # the gather reads weights the previous step updates: no gather-ahead, steps run in order
        with sampling.top_level_call() as step:
            if not upcoming and not eng.queue:
                b = eng.stage(nodes, labels, step)
                eng.train_step(b, lr, self.grad_allreduce)
                return eng.read_loss() if sync else eng.read_loss_async()
            same = lambda entry, ids: entry["ids"] is not None and len(entry["ids"]) == len(ids) and \
                np.array_equal(entry["ids"], np.asarray(ids, dtype=np.int64))
            if not eng.queue or not same(eng.queue[0], nodes):
                eng.reset_pipeline()
                eng.push(nodes, labels, step)
            for j, (nn, ll) in enumerate(upcoming[:eng.depth - 1]):
                if len(eng.queue) > j + 1 and not same(eng.queue[j + 1], nn):
                    eng.drop_queued(j + 1)                   # a different batch was announced earlier
                if len(eng.queue) <= j + 1:
                    eng.push(nn, ll, step + 1 + j)
            if len(eng.queue) > len(upcoming) + 1:
                eng.drop_queued(len(upcoming) + 1)
            eng.step_pipelined(lr, self.grad_allreduce)
        return eng.read_loss() if sync else eng.read_loss_async()


class _HostLoss:
    """Loss of one step of a multi-step launch: a slot of the launch's pinned loss buffer, valid once its event fired."""

    def __init__(self, event, buf, j):
        self._event, self._buf, self._j, self._value = event, buf, j, None

    def ready(self):
        return self._value is not None or self._event.query()

    def __float__(self):
        if self._value is None:
            self._event.synchronize()
            self._value = float(self._buf[self._j])
        return self._value


class StreamTrainer:
    """The reference's training loop (model.py:241-252) as a stream: ``feed(batch_nodes, labels)`` once per minibatch,
    ``finish()`` at the end; both return the losses (floats, in order) of the steps that have completed since the last
    call.  Every batch still goes through one ``zero_grad / loss / backward / step`` (model.py:246-250) -- its ids and
    labels travel host -> device and its loss device -> host -- but ``steps_per_launch`` of them are submitted as ONE
    CUDA-graph replay (``TrainEngine.run_host_queue``: the per-step host -> device copies out of pinned memory and the
    loss read-backs are nodes of the graph), so the Python / launch cost per step is paid once per launch.  Knowing the
    batches in advance is what a data loader gives; the reference's loop (one blocking step per iteration) cannot use it.
    Canonical 2-layer wiring (fused engine) with equal batch sizes; anything else falls back to ``train_step`` per batch."""

    def __init__(self, model, lr=0.7, steps_per_launch=4):
        self.model, self.lr = model, float(lr)
        self.k = max(4, int(steps_per_launch) - int(steps_per_launch) % 4)
        self.eng, self.b = None, None
        self.pending, self.fifo, self.fed = [], [], 0
        self.variants, self.busy, self.turn = None, None, 0

    # ---- helpers
    def _collect(self, wait):
        out = []
        while self.fifo and (wait or self.fifo[0].ready()):
            out.append(float(self.fifo.pop(0)))
        return out

    def _single(self, block):
        eng = self.eng
        eng.push(None, None, None, packed=(block, self.b))
        if len(eng.queue) == 3:
            eng.step_pipelined(self.lr, self.model.grad_allreduce)
            self.fifo.append(eng.read_loss_async())

    def _submit(self):
        eng, k = self.eng, self.k
        if self.variants is None:
            # two (staging blocks, losses) pinned buffer pairs, used alternately; they are baked into the captured graphs,
            # so they live on the engine and are shared by every stream over it
            cache = eng.__dict__.setdefault("_stream_bufs", {})
            if k not in cache:
                nb = eng.sets[0].stage_dev.numel()
                cache[k] = ([(torch.empty((k, nb), dtype=torch.uint8).pin_memory(),
                              torch.zeros(k, dtype=torch.float32).pin_memory()) for _ in range(2)], [None, None])
            self.variants, self.busy = cache[k]
        v = self.turn
        self.turn ^= 1
        blocks, losses = self.variants[v]
        if self.busy[v] is not None:
            self.busy[v].synchronize()             # the launch that last read these pinned buffers has finished
            for h in self.fifo:                    # its losses that nobody has asked for yet: read them out before the
                if isinstance(h, _HostLoss) and h._buf is losses and h._value is None:      # new launch overwrites them
                    h._value = float(losses[h._j])
        for j, blk in enumerate(self.pending[:k]):
            blocks[j].copy_(blk)
        del self.pending[:k]
        eng.run_host_queue(blocks, losses, self.lr)
        ev = torch.cuda.Event()
        ev.record()
        self.busy[v] = ev
        self.fifo.extend(_HostLoss(ev, losses, j) for j in range(k))

    # ---- public
    def feed(self, nodes, labels):
        from . import sampling
        from .engine import engine_for
        eng = engine_for(self.model, len(nodes)) if self.eng is None else self.eng
        plain = eng is None or eng.trainable_table or self.model.grad_allreduce is not None
        if not plain and self.b is not None and len(nodes) != self.b:
            done = self.finish()                  # a different batch size: drain, then start a new stream
            return done + self.feed(nodes, labels)
        if plain:
            return [self.model.train_step(nodes, labels, lr=self.lr)]
        if self.eng is None:
            self.eng, self.b = eng, len(nodes)
            eng.reset_pipeline()
        with sampling.top_level_call() as step:
            block = eng.pack_stage(nodes, labels, step)
        self.fed += 1
        if self.fed <= 3:                         # two batches queued + one single step = the pipeline's steady state
            self._single(block)
        else:
            self.pending.append(block)
            if len(self.pending) >= self.k:
                self._submit()
        return self._collect(wait=False)

    def finish(self):
        if self.eng is None:
            return []
        eng = self.eng
        for blk in self.pending:
            self._single(blk)
        self.pending = []
        while eng.queue:                          # drain the batches still in flight
            eng.step_pipelined(self.lr, self.model.grad_allreduce)
            self.fifo.append(eng.read_loss_async())
        eng.flush_update()
        out = self._collect(wait=True)
        self.eng, self.b, self.fed = None, None, 0
        return out


class GraphedStep:
    """The reference's timed unit -- ``optimizer.zero_grad(); loss = graphsage.loss(batch_nodes, labels);
    loss.backward(); optimizer.step()`` (model.py:245-250) -- for ANY model wired from the drop-in modules (any depth
    through the closure recursion of model.py:220-221, local or partitioned table / CSR with peer-memory lookups),
    captured ONCE as a CUDA graph and replayed per minibatch.

    The fused engine covers the canonical 2-layer wiring; deeper models (BASELINE config 5: 3 layers on a partitioned
    graph) run through the op-by-op autograd path, which costs ~150 kernel launches and, until round 2, one host
    read per aggregator call.  Here the aggregators run in ``static_shapes`` mode (no size ever leaves the device) and
    the sampler step lives in device memory (``sampling.static_step``), so forward + backward + SGD is a fixed launch
    list: the first two calls run eagerly (they size every scratch buffer), the third captures, later calls replay.
    Data parallel (``world > 1``): the gradient all-reduce (NCCL) and the SGD step follow the replay eagerly.
    Batches must have exactly ``batch_size`` targets."""

    def __init__(self, model, batch_size, lr=0.7, world=1, n_global=None, group=None, sampler_step=0):
        from . import sampling
        self.model, self.lr, self.world, self.group = model, float(lr), int(world), group
        self.n_local, self.n_global = int(batch_size), int(n_global if n_global is not None else batch_size * world)
        dev = _device()
        self.ids = torch.zeros(batch_size, dtype=torch.int32, device=dev)
        self.labels = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self.step_dev = torch.full((1,), int(sampler_step if sampler_step else sampling.get_step()), dtype=torch.int64, device=dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = SGD(self.params, lr)
        self.graph, self.loss, self._calls, self._stream = None, None, 0, None
        model.use_engine = False

    def _body(self):
        from . import sampling
        from .aggregators import static_shapes
        ops.advance_step(self.step_dev)
        self.opt.zero_grad()
        with sampling.static_step(self.step_dev), static_shapes():
            loss = self.model.loss(self.ids, self.labels)
        loss.backward()
        if self.world == 1:
            self.opt.step()
        return loss

    def __call__(self, nodes, labels):
        """One train step on host (or device) ids / labels; returns the loss as a device scalar."""
        if len(nodes) != self.ids.shape[0]:
            raise ValueError("GraphedStep was built for batches of %d targets" % self.ids.shape[0])
        self.ids.copy_(ops.as_ids(nodes, self.ids.device), non_blocking=True)
        self.labels.copy_(self.model._labels(labels), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        elif self._calls < 2:
            # eager warm-up on the side stream the capture will use (never the legacy default stream: autograd's
            # AccumulateGrad nodes remember the stream they were created on)
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            before = ops.LAUNCHES[0]
            with torch.cuda.stream(self._stream):
                self.loss = self._body()
            torch.cuda.current_stream().wait_stream(self._stream)
            self.launches_per_step = ops.LAUNCHES[0] - before      # kernels of libgsage_sm100.so in one step
        else:
            self.loss = None                     # drop the last eager autograd graph before capturing a new one
            self.opt.zero_grad()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self.loss = self._body()
            self.graph = g
            g.replay()
        self._calls += 1
        if self.world > 1:
            from . import sharded
            sharded.allreduce_grads(self.params, self.world, self.n_local, self.n_global, self.group)
            self.opt.step()
        return self.loss


# ------------------------------------------------------------------------------------------------
# Driver: the drop-in for ``python -m graphsage.model`` (graphsage/model.py:184-259, 539-567)
# ------------------------------------------------------------------------------------------------
def run_model(dataset, initializer, seed, epochs, classify="node", batch_size=128, feature_dim=100, identity_dim=50,
              data_root=".", num_samples=None, as_run=False, lr=0.7, gcn=True, data=None, verbose=True):
    """Train + validate a 2-layer supervised GraphSAGE on cora / citeseer / pubmed: same flow, same
    constants and same printed lines as the reference's ``run_model`` (model.py:184-259).

    The reference's driver has three accidents which are NOT reproduced by default:
    it sets ``enc.num_samples`` (an attribute nobody reads, model.py:223-224) so the effective fan-out is
    the constructor default 10/10; it slices ``train[batch:max(train_num, batch+batch_size)]``
    (model.py:244), so every "batch" is the rest of the epoch; and for initialisers whose table width
    differs from ``feature_dim`` (pagerank, deepwalk) the weight shape does not match.  Here the nominal
    per-dataset fan-outs (model.py:188-189) are applied to ``num_sample``, batches are
    ``train[batch:batch+batch_size]``, and the table width is taken from the table.  ``as_run=True``
    restores the reference's EFFECTIVE behaviour (fan-out 10/10, rest-of-epoch batches) for like-for-like
    F1 / timing comparisons.  Extra arguments: ``data_root`` (the reference reads from the CWD),
    ``num_samples=(inner, outer)``, ``lr``, ``gcn`` (the reference hard-codes gcn=True encoders),
    ``data=(feat_data, labels, adj_lists)`` to skip the loader.  Returns a dict with the numbers printed."""
    import random
    import time
    from . import data as D, sampling
    from .engine import engine_for
    spec = D.DATASETS[dataset]
    np.random.seed(seed)                                                     # model.py:192-193
    random.seed(seed)
    sampling.seed(seed)                                                      # the device sampler's stream
    num_nodes, num_classes = spec["num_nodes"], spec["num_classes"]
    feat_data, labels, adj_lists = data if data is not None else D.load_dataset(dataset, feature_dim, initializer, data_root)
    num_nodes = feat_data.shape[0]
    if initializer != "None":
        feature_dim = feat_data.shape[1]                                     # model.py:209-212, for every table
    else:
        feature_dim = feat_data.shape[1] if data is not None else spec["attr_dim"]
    if verbose:
        print("feature dim is", feature_dim)
    dev = _device()
    features = nn.Embedding(num_nodes, feature_dim, device="meta")
    table = ops.empty_rows(num_nodes, feature_dim, dev, zero=True)
    table.copy_(torch.from_numpy(np.asarray(feat_data, dtype=np.float32)))
    features.weight = nn.Parameter(table, requires_grad=False)               # model.py:214-215
    k1, k2 = (10, 10) if as_run else (num_samples if num_samples is not None else spec["num_samples"])
    model, (enc1, enc2) = build_sage(features, feature_dim, [identity_dim, 128], adj_lists, [k1, k2], num_classes,
                                     gcn=gcn, initializer=initializer, feature_dim=feature_dim, num_nodes=num_nodes)
    rand_indices = np.random.permutation(num_nodes)                          # model.py:229-235
    test_end, val_end = int(0.1 * num_nodes), int(0.2 * num_nodes)
    test, val, train = rand_indices[:test_end], rand_indices[test_end:val_end], list(rand_indices[val_end:])
    train_num = len(train)
    fused = engine_for(model, min(batch_size, train_num) if not as_run else train_num) is not None
    optimizer = None if fused else SGD(model.parameters(), lr=lr)             # model.py:237 on gs_sgd_step
    times, losses = [], []
    for epoch in range(epochs):
        random.shuffle(train)                                                # model.py:242
        for batch in range(0, train_num, batch_size):
            stop = max(train_num, batch + batch_size) if as_run else batch + batch_size
            batch_nodes = train[batch:stop]
            batch_labels = labels[np.array(batch_nodes)]
            torch.cuda.synchronize()
            start = time.time()
            if fused:                                                        # zero_grad + loss + backward + step, fused
                loss = model.train_step(batch_nodes, batch_labels, lr=lr)
            else:                                                            # model.py:246-250 verbatim
                optimizer.zero_grad()
                loss_t = model.loss(batch_nodes, torch.LongTensor(batch_labels))
                loss_t.backward()
                optimizer.step()
                loss = loss_t.item()
            torch.cuda.synchronize()
            times.append(time.time() - start)
            losses.append(loss)
    from sklearn.metrics import f1_score
    val_output = model.forward(val)                                          # model.py:256
    pred = val_output.detach().cpu().numpy().argmax(axis=1)
    out = {"f1_micro": f1_score(labels[val], pred, average="micro"), "f1_macro": f1_score(labels[val], pred, average="macro"),
           "avg_batch_time": float(np.mean(times)) if times else 0.0, "losses": losses, "fused_engine": fused,
           "num_sample": (k1, k2), "test_nodes": test, "model": model}
    if verbose:
        print("Validation F1 micro:", out["f1_micro"])
        print("Validation F1 macro:", out["f1_macro"])
        print("Average batch time:", out["avg_batch_time"])
    return out


def main(argv=None):
    """Same flags as the reference's CLI (model.py:539-567) plus the knobs run_model adds."""
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--initializer", type=str, default="None", help="node feature initialiation method")
    parser.add_argument("--identity_dim", type=int, default=50, help="node embedding dimension")
    parser.add_argument("--feature_dim", type=int, default=100, help="node feature dimension")
    parser.add_argument("--seed", type=int, default=1, help="random seed for initialization")
    parser.add_argument("--epochs", type=int, default=5, help="number of epochs")
    parser.add_argument("--dataset", type=str, default="cora", help="dataset used")
    parser.add_argument("--classify", type=str, default="node", help="classify task")
    parser.add_argument("--batch_size", type=int, default=128)
    parser.add_argument("--data_root", type=str, default=".", help="directory holding cora/, citeseer/, pubmed-data/")
    parser.add_argument("--num_sample1", type=int, default=None, help="fan-out of the inner layer")
    parser.add_argument("--num_sample2", type=int, default=None, help="fan-out of the outer layer")
    parser.add_argument("--lr", type=float, default=0.7)
    parser.add_argument("--as_run", action="store_true",
                        help="reproduce the reference driver's effective behaviour (fan-out 10/10, rest-of-epoch batches)")
    parser.add_argument("--synthetic", action="store_true",
                        help="write a synthetic dataset in the reference's file formats into --data_root first")
    args = parser.parse_args(argv)
    if args.synthetic:
        from . import data as D
        D.write_synthetic_dataset(args.dataset, args.data_root, seed=args.seed)
    ns = None
    if args.num_sample1 is not None or args.num_sample2 is not None:
        from . import data as D
        d1, d2 = D.DATASETS[args.dataset]["num_samples"]
        ns = (args.num_sample1 or d1, args.num_sample2 or d2)
    return run_model(args.dataset, args.initializer, args.seed, args.epochs, classify=args.classify,
                     batch_size=args.batch_size, feature_dim=args.feature_dim, identity_dim=args.identity_dim,
                     data_root=args.data_root, num_samples=ns, as_run=args.as_run, lr=args.lr)


if __name__ == "__main__":
    main()
