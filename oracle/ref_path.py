"""Torch-CPU restatement of the reference hot path (TEST INFRASTRUCTURE -- see
oracle/__init__.py; never imported by the product package).

Follows zjzijielu/graphsage-simple:
  * neighbour sampling ............ graphsage/aggregators.py:42-48
  * (intended) GCN self-loop union  graphsage/aggregators.py:50-51
  * dedup + dense row-normalised mask + ``mask.mm`` .. aggregators.py:52-61, 74
  * feature lookup through the ``features`` callable .. aggregators.py:62-65
  * trainable-table remap for 1hot/node_degree ........ aggregators.py:30-31, 68-71
  * self lookup + concat + ``W.mm(combined.t())`` + relu/sigmoid .. encoders.py:47-61
  * classifier + cross-entropy ...... graphsage/model.py:59-69
  * SGD step (lr 0.7, no momentum) .. graphsage/model.py:237, 246-250

It deliberately keeps the reference's *formulation* (a dense B x U mask multiplied into
the gathered unique rows, Python-level sampling and dedup) because it doubles as the
"reference CPU path" timed by ``bench.py --impl reference``; the arithmetic is the same
ATen CPU ops the reference calls, driven through autograd for the backward.

PARITY PIN: tests/test_oracle_golden.py checks every function below against outputs of
the unmodified reference imported from /root/reference (tests/golden/make_golden.py).
"""
import random as _random

import numpy as np
import torch

SIGMOID_INITIALIZERS = ("node_degree", "shared", "pagerank")   # encoders.py:58
TABLE_INITIALIZERS = ("1hot", "node_degree")                   # aggregators.py:30, 68


def draw_neighbours(to_neighs, num_sample, rng=_random):
    """aggregators.py:42-48.  ``random.sample`` on a set is what CPython <= 3.10 did by
    converting the set to a tuple first (Lib/random.py); the conversion is spelt out here
    so the function runs on Python >= 3.11 with the same draws."""
    if num_sample is None:
        return list(to_neighs)
    out = []
    for neigh in to_neighs:
        if len(neigh) >= num_sample:
            out.append(set(rng.sample(tuple(neigh), num_sample)))
        else:
            out.append(neigh)
    return out


def add_self_loops(nodes, samp_neighs):
    """What aggregators.py:50-51 means (as written ``set + set`` raises TypeError):
    union each sampled set with its own node, set semantics (no double count)."""
    return [set(s) | {int(nodes[i])} for i, s in enumerate(samp_neighs)]


def mean_of_sampled(samp_neighs, lookup, table_embed=None):
    """aggregators.py:52-74: dedup, dense 0/1 mask, row-normalise, gather the distinct
    rows through ``lookup`` and multiply.  ``table_embed`` (an nn.Embedding) reproduces the
    1hot/node_degree branch at :68-71.  Returns (to_feats, unique_nodes_list)."""
    distinct = list(set.union(*samp_neighs))
    column_of = {n: i for i, n in enumerate(distinct)}
    mask = torch.zeros(len(samp_neighs), len(distinct))
    cols = [column_of[n] for s in samp_neighs for n in s]
    rows = [i for i, s in enumerate(samp_neighs) for _ in range(len(s))]
    mask[rows, cols] = 1
    mask = mask.div(mask.sum(1, keepdim=True))
    rows_of_distinct = lookup(torch.LongTensor(distinct))
    if table_embed is not None:
        hot = [int(np.where(r == 1)[0][0]) for r in rows_of_distinct.detach().numpy()]
        rows_of_distinct = table_embed(torch.LongTensor(hot))
    return mask.mm(rows_of_distinct), distinct


class Layer:
    """One reference Encoder + its MeanAggregator (encoders.py:8-62, aggregators.py:12-76)
    as a plain object.  ``lookup`` maps a LongTensor of ids to rows [n, feat_dim]."""

    def __init__(self, lookup, feat_dim, embed_dim, adj_lists, num_sample=10,
                 gcn=False, agg_gcn=False, initializer="None", table_embed=None,
                 weight=None, rng=_random):
        self.lookup = lookup
        self.feat_dim = feat_dim
        self.embed_dim = embed_dim
        self.adj_lists = adj_lists
        self.num_sample = num_sample
        self.gcn = gcn
        self.agg_gcn = agg_gcn
        self.initializer = initializer
        self.table_embed = table_embed if initializer in TABLE_INITIALIZERS else None
        self.rng = rng
        k_in = feat_dim if gcn else 2 * feat_dim                     # encoders.py:31-32
        if weight is None:
            weight = torch.empty(embed_dim, k_in)
            torch.nn.init.xavier_uniform_(weight)                    # encoders.py:36
        self.weight = weight.clone().requires_grad_(True)
        self.trace = []          # (nodes, samp_neighs) per aggregator call, in call order

    def parameters(self):
        ps = [self.weight]
        if self.table_embed is not None:
            ps += list(self.table_embed.parameters())
        return ps

    def aggregate(self, nodes):
        to_neighs = [self.adj_lists[int(v)] for v in nodes]          # encoders.py:47
        samp = draw_neighbours(to_neighs, self.num_sample, self.rng)
        if self.agg_gcn:
            samp = add_self_loops(nodes, samp)
        self.trace.append(([int(v) for v in nodes], [set(s) for s in samp]))
        feats, _ = mean_of_sampled(samp, self.lookup, self.table_embed)
        return feats

    def __call__(self, nodes):
        """encoders.py:40-62; returns [embed_dim, len(nodes)] like the reference."""
        neigh = self.aggregate(nodes)
        if not self.gcn:
            own = self.lookup(torch.LongTensor([int(v) for v in nodes]))   # encoders.py:53
            combined = torch.cat([own, neigh], dim=1)
        else:
            combined = neigh
        pre = self.weight.mm(combined.t())
        if self.initializer in SIGMOID_INITIALIZERS:
            return torch.sigmoid(pre)
        return torch.relu(pre)


class TwoLayerModel:
    """The wiring of model.py:214-227 (+ SupervisedGraphSage 52-69) with every knob
    explicit: feature table -> Layer 1 -> closure ``lambda n: enc1(n).t()`` -> Layer 2 ->
    classifier.  ``gcn`` is the Encoder flag (True = as the fork runs it, no self term;
    False = GraphSAGE concat)."""

    def __init__(self, table, adj1, adj2, d1, d2, num_classes, k1, k2, gcn=False,
                 agg_gcn=False, initializer="None", w1=None, w2=None, wc=None,
                 table_embed=None, rng=_random):
        self.table = table                                   # frozen fp32 [N, F]
        first = lambda ids: table[ids]                       # nn.Embedding lookup, model.py:214-215
        self.enc1 = Layer(first, table.shape[1] if table_embed is None else table_embed.embedding_dim,
                          d1, adj1, k1, gcn, agg_gcn, initializer, table_embed, w1, rng)
        if table_embed is not None:
            # the self lookup of a non-gcn layer 1 sees raw one-hot rows (encoders.py:53);
            # the fork only ever runs table initialisers with gcn=True (model.py:219).
            assert gcn, "table initialisers are only defined for gcn encoders in the reference"
        second = lambda ids: self.enc1(ids).t()              # model.py:220-221
        self.enc2 = Layer(second, d1, d2, adj2, k2, gcn, agg_gcn, "None", None, w2, rng)
        if wc is None:
            wc = torch.empty(num_classes, d2)
            torch.nn.init.xavier_uniform_(wc)                # model.py:59-60
        self.weight = wc.clone().requires_grad_(True)

    def parameters(self):
        return [self.weight] + self.enc2.parameters() + self.enc1.parameters()

    def forward(self, nodes):
        embeds = self.enc2(nodes)                            # model.py:63
        return self.weight.mm(embeds).t()                    # model.py:64-65

    def loss(self, nodes, labels):
        scores = self.forward(nodes)
        labels = torch.as_tensor(np.asarray(labels), dtype=torch.long).reshape(-1)
        return torch.nn.functional.cross_entropy(scores, labels)    # model.py:57, 69

    def train_step(self, nodes, labels, lr=0.7):
        """model.py:246-250: zero_grad, loss, backward, plain SGD."""
        for p in self.parameters():
            p.grad = None
        loss = self.loss(nodes, labels)
        loss.backward()
        with torch.no_grad():
            for p in self.parameters():
                if p.grad is not None:
                    p.add_(p.grad, alpha=-lr)
        return loss.detach()


class StackedModel:
    """The same wiring for ANY depth (BASELINE config 5 is three layers deep): layer l's lookup is the closure
    ``lambda ids: layer_{l-1}(ids).t()`` (model.py:220-221 applied repeatedly), the classifier sits on the last
    layer (model.py:52-69).  ``adjs``, ``dims`` and ``ks`` are per layer, innermost first."""

    def __init__(self, table, adjs, dims, num_classes, ks, gcn=False, weights=None, wc=None, rng=_random):
        self.table = table
        self.layers = []
        lookup, feat_dim = (lambda ids: table[ids]), table.shape[1]
        for l, (adj, dim, k) in enumerate(zip(adjs, dims, ks)):
            layer = Layer(lookup, feat_dim, dim, adj, k, gcn, False, "None", None,
                          None if weights is None else weights[l], rng)
            self.layers.append(layer)
            lookup, feat_dim = (lambda ids, below=layer: below(ids).t()), dim
        if wc is None:
            wc = torch.empty(num_classes, dims[-1])
            torch.nn.init.xavier_uniform_(wc)
        self.weight = wc.clone().requires_grad_(True)

    def parameters(self):
        return [self.weight] + [p for layer in reversed(self.layers) for p in layer.parameters()]

    def forward(self, nodes):
        return self.weight.mm(self.layers[-1](nodes)).t()

    def loss(self, nodes, labels):
        labels = torch.as_tensor(np.asarray(labels), dtype=torch.long).reshape(-1)
        return torch.nn.functional.cross_entropy(self.forward(nodes), labels)

    def train_step(self, nodes, labels, lr=0.7):
        """model.py:246-250: zero_grad, loss, backward, plain SGD."""
        for p in self.parameters():
            p.grad = None
        loss = self.loss(nodes, labels)
        loss.backward()
        with torch.no_grad():
            for p in self.parameters():
                if p.grad is not None:
                    p.add_(p.grad, alpha=-lr)
        return loss.detach()


def adj_from_csr(rowptr, col):
    """CSR -> the reference's ``adj_lists`` mapping (model.py:303-310 builds the same
    structure from the edge file)."""
    rowptr = np.asarray(rowptr)
    col = np.asarray(col)
    return {v: set(int(c) for c in col[rowptr[v]:rowptr[v + 1]]) for v in range(len(rowptr) - 1)}


def adj_from_tiles(nodes, idx, cnt):
    """Pre-sampled fixed-width tiles -> a replay mapping for ``num_sample=None``
    (SURVEY.md s8c replay mechanism)."""
    return {int(v): set(int(c) for c in idx[i, :cnt[i]]) for i, v in enumerate(nodes)}
