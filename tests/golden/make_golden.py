#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports ``graphsage.aggregators / encoders / model`` from /root/reference (never this
repo's drop-in package), applies the one compatibility shim SURVEY.md s8c documents --
``random.sample(set, k)`` -> ``random.sample(tuple(set), k)``, which is what CPython <= 3.10
did internally -- and records inputs + outputs of the reference for small seeded cases.
The oracle (oracle/ref_path.py) and the CUDA path are both checked against these files.
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

_orig_sample = random.sample
SAMPLE_LOG = []


def _sample_shim(pop, k):
    res = _orig_sample(tuple(pop) if isinstance(pop, (set, frozenset)) else pop, k)
    SAMPLE_LOG.append(list(res))
    return res


random.sample = _sample_shim

from graphsage.aggregators import MeanAggregator  # noqa: E402
from graphsage.encoders import Encoder  # noqa: E402
from graphsage.model import SupervisedGraphSage  # noqa: E402

assert os.path.realpath(sys.modules["graphsage.aggregators"].__file__).startswith(os.path.realpath(REF))


def quiet(fn, *a, **kw):
    """Encoder.__init__ prints (encoders.py:38)."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def random_graph(rng, n, avg_deg, min_deg=1):
    adj = {v: set() for v in range(n)}
    m = n * avg_deg // 2
    for a, b in rng.integers(0, n, (m, 2)):
        adj[int(a)].add(int(b))
        adj[int(b)].add(int(a))
    for v in range(n):
        while len(adj[v]) < min_deg:
            u = int(rng.integers(0, n))
            if u != v:
                adj[v].add(u)
                adj[u].add(v)
    return adj


def to_csr(adj, n):
    rowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for v in range(n):
        nb = sorted(adj[v])
        cols.extend(nb)
        rowptr[v + 1] = rowptr[v] + len(nb)
    return rowptr, np.asarray(cols, dtype=np.int32)


def canonical_adj(rowptr, col):
    """Rebuild the sets by inserting each row's ids in ascending order, so that anyone
    holding the CSR can reconstruct sets with the SAME CPython iteration order (which
    ``random.sample(tuple(set))`` depends on, SURVEY.md s8c caveat (i))."""
    return {v: set(int(c) for c in col[rowptr[v]:rowptr[v + 1]]) for v in range(len(rowptr) - 1)}


def citeseer_graph():
    """Citeseer topology in the loader's index space (citeseer.cites.parsed is produced by
    citeseer/parse_citeseer_link.py with the same node_map rule as model.py:113-119);
    symmetrised the way model.py:176-181 does."""
    n = 3312
    adj = {v: set() for v in range(n)}
    with open(os.path.join(REF, "citeseer", "citeseer.cites.parsed")) as fp:
        for line in fp:
            a, b = map(int, line.split())
            adj[a].add(b)
            adj[b].add(a)
    for v in range(n):      # parsed file leaves no isolated node; keep the guarantee explicit
        assert adj[v]
    return adj, n


def subset_tiles(rng, adj, nodes, k):
    """Pick <=k neighbours per node (all when deg<=k) with numpy -- any subsets will do,
    they are fed to BOTH sides verbatim (num_sample=None replay)."""
    idx = np.full((len(nodes), k), -1, dtype=np.int32)
    cnt = np.zeros(len(nodes), dtype=np.int32)
    for i, v in enumerate(nodes):
        nb = sorted(adj[int(v)])
        if len(nb) > k:
            nb = sorted(rng.choice(nb, size=k, replace=False).tolist())
        idx[i, :len(nb)] = nb
        cnt[i] = len(nb)
    return idx, cnt


def tiles_to_adj(nodes, idx, cnt):
    return {int(v): set(int(c) for c in idx[i, :cnt[i]]) for i, v in enumerate(nodes)}


def embedding_of(table):
    emb = nn.Embedding(*table.shape)
    emb.weight = nn.Parameter(torch.FloatTensor(table), requires_grad=False)   # model.py:214-215
    return emb


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


# --------------------------------------------------------------------------- aggregator
def case_aggregator():
    rng = np.random.default_rng(101)
    n, f = 120, 20
    adj = random_graph(rng, n, 6)
    rowptr, col = to_csr(adj, n)
    adj = canonical_adj(rowptr, col)
    table = rng.standard_normal((n, f)).astype(np.float32)
    nodes = rng.permutation(n)[:33].astype(np.int64)
    agg = MeanAggregator(embedding_of(table), cuda=False, gcn=False)
    # (a) full neighbourhoods, num_sample=None (aggregators.py:47-48)
    full = agg.forward(list(nodes), [adj[int(v)] for v in nodes], None)
    # (b) Python-RNG sampling, num_sample=4 (aggregators.py:42-46)
    random.seed(11)
    SAMPLE_LOG.clear()
    sampled = agg.forward(list(nodes), [adj[int(v)] for v in nodes], 4)
    log = np.asarray([sorted(s) for s in SAMPLE_LOG], dtype=np.int32)
    # (c) replay of given subsets
    idx, cnt = subset_tiles(rng, adj, nodes, 3)
    rep = agg.forward(list(nodes), [set(idx[i, :cnt[i]].tolist()) for i in range(len(nodes))], None)
    save("aggregator", rowptr=rowptr, col=col, table=table, nodes=nodes,
         out_full=full.detach().numpy(), out_sampled=sampled.detach().numpy(), sample_log=log,
         sample_seed=np.int64(11), sample_k=np.int64(4),
         rep_idx=idx, rep_cnt=cnt, out_replay=rep.detach().numpy())


# --------------------------------------------------------------------------- encoder
def case_encoder():
    rng = np.random.default_rng(202)
    n, f, d = 150, 18, 10
    adj = random_graph(rng, n, 5)
    rowptr, col = to_csr(adj, n)
    table = rng.standard_normal((n, f)).astype(np.float32)
    nodes = rng.permutation(n)[:41].astype(np.int64)
    idx, cnt = subset_tiles(rng, adj, nodes, 4)
    rep = tiles_to_adj(nodes, idx, cnt)
    out = dict(rowptr=rowptr, col=col, table=table, nodes=nodes, idx=idx, cnt=cnt)
    for tag, gcn, init in (("sage_relu", False, "None"), ("gcn_relu", True, "None"),
                           ("sage_sigmoid", False, "shared"), ("gcn_sigmoid", True, "pagerank")):
        emb = embedding_of(table)
        agg = MeanAggregator(emb, cuda=False)
        enc = quiet(Encoder, emb, f, d, rep, agg, num_sample=None, gcn=gcn, cuda=False, initializer=init)
        torch.manual_seed(7)
        w = torch.empty_like(enc.weight.data)
        nn.init.xavier_uniform_(w)
        enc.weight.data.copy_(w)
        h = enc(list(nodes))                      # [d, n] (encoders.py:61)
        g = torch.from_numpy(rng.standard_normal(tuple(h.shape)).astype(np.float32))
        (h * g).sum().backward()
        out["w_" + tag] = w.numpy()
        out["h_" + tag] = h.detach().numpy()
        out["gout_" + tag] = g.numpy()
        out["gw_" + tag] = enc.weight.grad.numpy().copy()
    save("encoder", **out)


# --------------------------------------------------------------------------- 2-layer model
def build_reference_model(table, adj1, adj2, d1, d2, c, k1, k2, gcn, seed):
    f = table.shape[1]
    emb = embedding_of(table)
    agg1 = MeanAggregator(emb, cuda=False)
    enc1 = quiet(Encoder, emb, f, d1, adj1, agg1, num_sample=k1, gcn=gcn, cuda=False)
    agg2 = MeanAggregator(lambda nodes: enc1(nodes).t(), cuda=False)           # model.py:220
    enc2 = quiet(Encoder, lambda nodes: enc1(nodes).t(), enc1.embed_dim, d2, adj2, agg2,
                 num_sample=k2, base_model=enc1, gcn=gcn, cuda=False)          # model.py:221-222
    model = SupervisedGraphSage(c, enc2)
    torch.manual_seed(seed)
    for p in (model.weight, enc2.weight, enc1.weight):
        w = torch.empty_like(p.data)
        nn.init.xavier_uniform_(w)
        p.data.copy_(w)
    return model, enc1, enc2


def run_step(model, enc1, enc2, nodes, labels, lr=0.7):
    w0 = dict(wc=model.weight.data.numpy().copy(), w2=enc2.weight.data.numpy().copy(),
              w1=enc1.weight.data.numpy().copy())
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)   # model.py:237
    opt.zero_grad()
    scores = model.forward(list(nodes)).detach().numpy().copy()   # same replay -> same values
    loss = model.loss(list(nodes), torch.LongTensor(labels))       # model.py:247-248
    loss.backward()
    grads = dict(gwc=model.weight.grad.numpy().copy(), gw2=enc2.weight.grad.numpy().copy(),
                 gw1=enc1.weight.grad.numpy().copy())
    opt.step()
    w1 = dict(wc_new=model.weight.data.numpy().copy(), w2_new=enc2.weight.data.numpy().copy(),
              w1_new=enc1.weight.data.numpy().copy())
    return dict(scores=scores, loss=np.float32(loss.item()), **w0, **grads, **w1)


def case_model(name, gcn, graph="citeseer"):
    rng = np.random.default_rng(303 if gcn else 304)
    if graph == "citeseer":
        adj, n = citeseer_graph()
    else:
        n = 400
        adj = random_graph(rng, n, 8)
    rowptr, col = to_csr(adj, n)
    f, d1, d2, c, k1, k2, b = 12, 16, 12, 6, 3, 4, 48
    table = rng.standard_normal((n, f)).astype(np.float32)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    nodes = rng.permutation(n)[:b].astype(np.int64)
    # replayed samples: layer-2 tiles over the targets, layer-1 tiles over every node that
    # layer 1 is evaluated on (hop-1 uniques + the targets themselves in SAGE mode)
    idx2, cnt2 = subset_tiles(rng, adj, nodes, k2)
    hop1 = sorted(set(int(x) for i in range(b) for x in idx2[i, :cnt2[i]]) | set(int(v) for v in nodes))
    hop1 = np.asarray(hop1, dtype=np.int64)
    idx1, cnt1 = subset_tiles(rng, adj, hop1, k1)
    model, enc1, enc2 = build_reference_model(table, tiles_to_adj(hop1, idx1, cnt1),
                                              tiles_to_adj(nodes, idx2, cnt2), d1, d2, c, None, None, gcn, 9)
    res = run_step(model, enc1, enc2, nodes, labels[nodes])
    save(name, rowptr=rowptr, col=col, table=table, labels=labels, nodes=nodes,
         idx2=idx2, cnt2=cnt2, hop1=hop1, idx1=idx1, cnt1=cnt1, gcn=np.bool_(gcn), **res)


def case_model_live():
    """Live Python-RNG sampling through the whole 2-layer model (pins the oracle's
    ``random`` call order: agg1(U1) -> agg2(B) -> agg1(B))."""
    rng = np.random.default_rng(404)
    n, f, d1, d2, c, k1, k2, b = 300, 10, 8, 8, 4, 3, 5, 32
    adj = random_graph(rng, n, 9)
    rowptr, col = to_csr(adj, n)
    adj = canonical_adj(rowptr, col)
    table = rng.standard_normal((n, f)).astype(np.float32)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    nodes = rng.permutation(n)[:b].astype(np.int64)
    out = dict(rowptr=rowptr, col=col, table=table, labels=labels, nodes=nodes,
               k1=np.int64(k1), k2=np.int64(k2), seed=np.int64(5))
    for tag, gcn in (("sage", False), ("gcn", True)):
        model, enc1, enc2 = build_reference_model(table, adj, adj, d1, d2, c, k1, k2, gcn, 13)
        random.seed(5)
        res = run_step_live(model, enc1, enc2, nodes, labels[nodes])
        out.update({f"{tag}_{k}": v for k, v in res.items()})
    save("model_live", **out)


def run_step_live(model, enc1, enc2, nodes, labels, lr=0.7):
    w0 = dict(wc=model.weight.data.numpy().copy(), w2=enc2.weight.data.numpy().copy(),
              w1=enc1.weight.data.numpy().copy())
    opt = torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)
    opt.zero_grad()
    loss = model.loss(list(nodes), torch.LongTensor(labels))
    loss.backward()
    grads = dict(gwc=model.weight.grad.numpy().copy(), gw2=enc2.weight.grad.numpy().copy(),
                 gw1=enc1.weight.grad.numpy().copy())
    opt.step()
    return dict(loss=np.float32(loss.item()), **w0, **grads, wc_new=model.weight.data.numpy().copy())


# --------------------------------------------------------------------------- trainable table
def case_table_initialisers():
    """1hot / node_degree: the aggregator owns a trainable nn.Embedding indexed by the
    position of the 1 in the looked-up row (aggregators.py:30-31, 68-71); wired with
    gcn=True encoders as model.py:218-222 does."""
    rng = np.random.default_rng(505)
    n, fd, d1, d2, c, b = 60, 7, 6, 5, 3, 16
    adj = random_graph(rng, n, 5)
    rowptr, col = to_csr(adj, n)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    nodes = rng.permutation(n)[:b].astype(np.int64)
    idx2, cnt2 = subset_tiles(rng, adj, nodes, 3)
    hop1 = np.asarray(sorted(set(int(x) for i in range(b) for x in idx2[i, :cnt2[i]])), dtype=np.int64)
    idx1, cnt1 = subset_tiles(rng, adj, hop1, 3)
    out = dict(rowptr=rowptr, col=col, labels=labels, nodes=nodes, idx2=idx2, cnt2=cnt2,
               hop1=hop1, idx1=idx1, cnt1=cnt1)
    deg = np.asarray([len(adj[v]) for v in range(n)])
    for init in ("1hot", "node_degree"):
        if init == "1hot":
            table = np.eye(n, dtype=np.float32)                    # model.py:289-290 style
            num_rows, in_dim = n, n
        else:
            table = np.zeros((n, int(deg.max()) + 1), dtype=np.float32)   # one-hot of the degree
            table[np.arange(n), deg] = 1
            num_rows, in_dim = n, table.shape[1]
        emb = embedding_of(table)
        agg1 = MeanAggregator(emb, cuda=False, feature_dim=fd, num_nodes=num_rows, initializer=init)
        enc1 = quiet(Encoder, emb, fd, d1, tiles_to_adj(hop1, idx1, cnt1), agg1, num_sample=None,
                     gcn=True, cuda=False, initializer=init)
        agg2 = MeanAggregator(lambda x: enc1(x).t(), n, cuda=False)              # model.py:220 (int in the initializer slot)
        enc2 = quiet(Encoder, lambda x: enc1(x).t(), d1, d2, tiles_to_adj(nodes, idx2, cnt2), agg2,
                     num_sample=None, base_model=enc1, gcn=True, cuda=False)
        model = SupervisedGraphSage(c, enc2)
        torch.manual_seed(21)
        for p in (model.weight, enc2.weight, enc1.weight):
            w = torch.empty_like(p.data)
            nn.init.xavier_uniform_(w)
            p.data.copy_(w)
        e0 = torch.randn(num_rows, fd)
        agg1.embed.weight.data.copy_(e0)
        res = run_step(model, enc1, enc2, nodes, labels[nodes])
        res["embed"] = e0.numpy().copy()
        res["gembed"] = agg1.embed.weight.grad.numpy().copy()
        res["embed_new"] = agg1.embed.weight.data.numpy().copy()
        res["table"] = table
        out.update({f"{init}_{k}": v for k, v in res.items()})
    save("table_init", **out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    case_aggregator()
    case_encoder()
    case_model("model_sage", gcn=False)
    case_model("model_gcn", gcn=True)
    case_model_live()
    case_table_initialisers()
