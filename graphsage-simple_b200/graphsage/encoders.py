"""Drop-in for graphsage/encoders.py of zjzijielu/graphsage-simple (Encoder, encoders.py:8-62):
same constructor, same public attributes, ``forward(nodes)`` returns ``[embed_dim, len(nodes)]``.
The adjacency mapping is converted once to a device CSR (graph.py); sampling, the mean, the
``W . combined^T`` contraction and the activation run in CUDA kernels."""
import torch
import torch.nn as nn
from torch.nn import init

from . import ops, sampling
from .aggregators import MeanAggregator, _device
from .functional import EncoderGemm, TableLookup
from .graph import CSRGraph, graph_of

SIGMOID_INITIALIZERS = ("node_degree", "shared", "pagerank")      # encoders.py:58


class Encoder(nn.Module):
    """Encodes a node using the 'convolutional' GraphSage approach."""

    def __init__(self, features, feature_dim, embed_dim, adj_lists, aggregator,
                 num_sample=10, initializer="None", base_model=None, gcn=False, cuda=False,
                 feature_transform=False, verbose=False):
        super().__init__()
        dev = _device()
        if isinstance(features, nn.Module):
            features.to(dev)
        self.features = features
        self.feat_dim = feature_dim
        self.adj_lists = adj_lists
        self.aggregator = aggregator
        self.num_sample = num_sample
        if base_model is not None:
            self.base_model = base_model                           # encoders.py:24-25
        self.gcn = gcn
        self.embed_dim = embed_dim
        self.cuda = cuda
        self.aggregator.cuda = cuda                                # encoders.py:30
        self.weight = nn.Parameter(torch.empty(
            embed_dim, self.feat_dim if self.gcn else 2 * self.feat_dim, device=dev))
        self.initializer = initializer
        init.xavier_uniform_(self.weight)                          # encoders.py:36
        if verbose:
            print("feat dim:", self.feat_dim, "embed_dim:", self.embed_dim)   # encoders.py:38
        self._graph = adj_lists if hasattr(adj_lists, "sample") else None      # CSRGraph / ShardedCSR

    @property
    def graph(self):
        """Device CSR of ``adj_lists`` (built on first use, shared between encoders that were
        given the same mapping object)."""
        if self._graph is None:
            # rows of the feature table (when known): isolated trailing nodes get a CSR row too
            n = getattr(self.features, "num_embeddings", None)
            if n is None and getattr(self, "base_model", None) is not None and self.base_model.adj_lists is self.adj_lists:
                n = self.base_model.graph.num_nodes
            self._graph = graph_of(self.adj_lists, _device(), num_nodes=n)
        return self._graph

    @property
    def activation(self):
        return ops.ACT_SIGMOID if self.initializer in SIGMOID_INITIALIZERS else ops.ACT_RELU

    def forward(self, nodes):
        """Generates embeddings for a batch of nodes -> FloatTensor [embed_dim, len(nodes)]."""
        dev = _device()
        with sampling.top_level_call():
            ids = ops.as_ids(nodes, dev)
            if isinstance(self.aggregator, MeanAggregator):
                neigh_feats = self.aggregator.forward(ids, None, self.num_sample,
                                                      initializer=self.initializer, graph=self.graph)
            else:   # foreign aggregator: the reference's calling convention (encoders.py:47)
                neigh_feats = self.aggregator.forward(
                    nodes, [self.adj_lists[int(node)] for node in nodes], self.num_sample,
                    initializer=self.initializer)
            if not self.gcn:
                if isinstance(self.features, nn.Embedding):        # encoders.py:53
                    self_feats = TableLookup.apply(self.features.weight, ids)
                else:
                    self_feats = self.features(ids.long())
                combined = torch.cat([self_feats, neigh_feats], dim=1)
            else:
                combined = neigh_feats
            h = EncoderGemm.apply(combined, self.weight, self.activation)
        return h.t()
