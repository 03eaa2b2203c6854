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

    def loss(self, nodes, labels):
        embeds = self.enc(nodes)
        return SoftmaxXent.apply(embeds.t(), self.weight, self._labels(labels))   # model.py:67-69
