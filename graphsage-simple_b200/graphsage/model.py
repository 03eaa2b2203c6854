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
                return _EngineLoss.apply(eng.loss, self.weight, enc2.weight, enc2.base_model.weight,
                                         eng.gwc, eng.gw2, eng.gw1)
        embeds = self.enc(nodes)
        return SoftmaxXent.apply(embeds.t(), self.weight, self._labels(labels))

    grad_allreduce = None   # data parallel: callable(flat_grads) run between backward and the SGD step

    def train_step(self, nodes, labels, lr=0.7, prefetch=None):
        """The reference's timed unit (model.py:246-250: zero_grad, loss, backward, SGD step) as
        one fused call; returns the loss as a Python float (one 4-byte device->host read).

        ``prefetch`` names the minibatches that FOLLOW this one -- ``(nodes, labels)`` or a list of
        up to two such pairs -- so that their neighbour sampling and feature gather run on side
        streams while this batch is in its GEMM/backward chain (the data-loader style overlap the
        reference's single Python thread cannot do).  Pass the same batches, in order, as
        ``nodes, labels`` of the following calls; results equal plain sequential steps."""
        from .engine import engine_for
        from . import sampling
        upcoming = [] if prefetch is None else ([prefetch] if isinstance(prefetch, tuple) else list(prefetch))
        eng = engine_for(self, max([len(nodes)] + [len(u[0]) for u in upcoming]))
        if eng is None:
            raise RuntimeError("train_step needs the canonical 2-layer wiring (model.py:214-227)")
        with sampling.top_level_call() as step:
            if not upcoming and not eng.queue:
                b = eng.stage(nodes, labels, step)
                eng.train_step(b, lr, self.grad_allreduce)
                return eng.read_loss()
            same = lambda entry, ids: entry["ids"] is not None and len(entry["ids"]) == len(ids) and \
                np.array_equal(entry["ids"], np.asarray(ids, dtype=np.int64))
            if not eng.queue or not same(eng.queue[0], nodes):
                eng.reset_pipeline()
                eng.push(nodes, labels, step)
            for j, (nn, ll) in enumerate(upcoming[:eng.depth - 1]):
                if len(eng.queue) > j + 1 and not same(eng.queue[j + 1], nn):
                    del eng.queue[j + 1:]                    # a different batch was announced earlier
                if len(eng.queue) <= j + 1:
                    eng.push(nn, ll, step + 1 + j)
            del eng.queue[len(upcoming) + 1:]
            eng.step_pipelined(lr, self.grad_allreduce)
        return eng.read_loss()
