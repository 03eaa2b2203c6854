#!/usr/bin/env python
"""Golden vectors for a THREE-layer supervised GraphSAGE-mean (the depth of BASELINE config 5) from the UNMODIFIED
reference: its Encoder / MeanAggregator / SupervisedGraphSage stacked by closure recursion exactly like
model.py:218-227 does for two layers, replaying fixed sampled neighbour lists (per-layer dicts with
``num_sample=None``).  Writes tests/golden/model_3layer.npz.

    python tests/golden/make_golden_3layer.py        # build container only (needs /root/reference)
"""
import os
import random
import sys
import warnings

import numpy as np

REF = os.environ.get("GSAGE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

_orig = random.sample
random.sample = lambda pop, k: _orig(tuple(pop) if isinstance(pop, (set, frozenset)) else pop, k)

from graphsage.aggregators import MeanAggregator  # noqa: E402
from graphsage.encoders import Encoder  # noqa: E402
from graphsage.model import SupervisedGraphSage  # noqa: E402

assert os.path.realpath(sys.modules["graphsage.aggregators"].__file__).startswith(os.path.realpath(REF))


def quiet(fn, *a, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def sampled_dict(rng, adj, n, k):
    """One fixed draw per node: min(k, deg) neighbours (what aggregators.py:42-46 would produce once)."""
    out = {}
    for v in range(n):
        nb = sorted(adj[v])
        out[v] = set(nb if len(nb) <= k else rng.choice(nb, k, replace=False).tolist())
    return out


def tiles(sd, n, width):
    idx = np.full((n, width), -1, dtype=np.int32)
    cnt = np.zeros(n, dtype=np.int32)
    for v in range(n):
        row = sorted(sd[v])
        idx[v, :len(row)] = row
        cnt[v] = len(row)
    return idx, cnt


def main():
    rng = np.random.default_rng(33)
    torch.manual_seed(33)
    n, f, dims, c, fan = 240, 14, [12, 10, 9], 5, [3, 4, 5]
    adj = {v: set() for v in range(n)}
    for a, b in rng.integers(0, n, (n * 4, 2)):
        if a != b:
            adj[int(a)].add(int(b)); adj[int(b)].add(int(a))
    for v in range(n):
        adj[v].add((v + 1) % n); adj[(v + 1) % n].add(v)
    table = rng.standard_normal((n, f)).astype(np.float32)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    S = [sampled_dict(rng, adj, n, k) for k in fan]                      # innermost layer first
    features = nn.Embedding(n, f)
    features.weight = nn.Parameter(torch.FloatTensor(table), requires_grad=False)
    agg1 = MeanAggregator(features, cuda=False)
    enc1 = quiet(Encoder, features, f, dims[0], S[0], agg1, num_sample=None, gcn=False, cuda=False)
    agg2 = MeanAggregator(lambda nodes: enc1(nodes).t(), cuda=False)
    enc2 = quiet(Encoder, lambda nodes: enc1(nodes).t(), enc1.embed_dim, dims[1], S[1], agg2, num_sample=None,
                 base_model=enc1, gcn=False, cuda=False)
    agg3 = MeanAggregator(lambda nodes: enc2(nodes).t(), cuda=False)
    enc3 = quiet(Encoder, lambda nodes: enc2(nodes).t(), enc2.embed_dim, dims[2], S[2], agg3, num_sample=None,
                 base_model=enc2, gcn=False, cuda=False)
    model = SupervisedGraphSage(c, enc3)
    w = {"w1": enc1.weight.detach().numpy().copy(), "w2": enc2.weight.detach().numpy().copy(),
         "w3": enc3.weight.detach().numpy().copy(), "wc": model.weight.detach().numpy().copy()}
    nodes = rng.permutation(n)[:40]
    scores = model.forward(list(nodes)).detach().numpy().copy()
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=0.7)
    opt.zero_grad()
    loss = model.loss(list(nodes), torch.LongTensor(labels[nodes]))
    loss.backward()
    grads = {"gw1": enc1.weight.grad.numpy().copy(), "gw2": enc2.weight.grad.numpy().copy(),
             "gw3": enc3.weight.grad.numpy().copy(), "gwc": model.weight.grad.numpy().copy()}
    opt.step()
    new = {"w1_new": enc1.weight.detach().numpy().copy(), "w3_new": enc3.weight.detach().numpy().copy(),
           "wc_new": model.weight.detach().numpy().copy()}
    out = dict(table=table, labels=labels, nodes=nodes.astype(np.int64), scores=scores, loss=np.float32(loss.item()),
               fan=np.array(fan), **w, **grads, **new)
    for l, (sd, k) in enumerate(zip(S, fan), 1):
        out["idx%d" % l], out["cnt%d" % l] = tiles(sd, n, k)
    np.savez_compressed(os.path.join(HERE, "model_3layer.npz"), **out)
    print("loss", loss.item(), "scores", scores.shape, {k: v.shape for k, v in grads.items()})


if __name__ == "__main__":
    main()
