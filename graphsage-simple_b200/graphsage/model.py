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

        ``prefetch=(next_nodes, next_labels)`` starts the NEXT minibatch's neighbour sampling and
        feature gather on a side stream while this one is in its GEMM/backward chain (the
        data-loader style overlap the reference's single Python thread cannot do); pass the same
        batch as ``nodes, labels`` of the following call."""
        from .engine import engine_for
        from . import sampling
        eng = engine_for(self, max(len(nodes), len(prefetch[0]) if prefetch is not None else 0))
        if eng is None:
            raise RuntimeError("train_step needs the canonical 2-layer wiring (model.py:214-227)")
        with sampling.top_level_call() as step:
            if prefetch is None and getattr(self, "_primed", None) is None:
                b = eng.stage(nodes, labels, step)
                eng.train_step(b, lr, self.grad_allreduce)
                return eng.read_loss()
            primed = getattr(self, "_primed", None)
            if primed is None or len(primed) != len(nodes) or not np.array_equal(primed, np.asarray(nodes)):
                b = eng.stage(nodes, labels, step)          # not prefetched: do its gather chain now
                eng.prime(b)
            b = len(nodes)
            if prefetch is not None:
                eng.enable_pipeline()
                nb = eng.stage(prefetch[0], prefetch[1], step + 1, slot=1 - eng.cur)
                eng.train_step_pipelined(b, lr, nb, self.grad_allreduce)
                self._primed = np.array(prefetch[0], copy=True)
            else:                                           # last batch of a pipelined run
                eng._run(("tail", b, eng.cur), lambda: eng._compute_chain(eng.sets[eng.cur], b))
                if self.grad_allreduce is not None:
                    self.grad_allreduce(eng.flat_g)
                eng.update(lr)
                self._primed = None
        return eng.read_loss()
