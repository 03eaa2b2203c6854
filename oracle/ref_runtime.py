"""Run the UNMODIFIED reference (``oracle/_ref``, made by ``oracle/build_ref.py``) -- TEST INFRASTRUCTURE.

``load()`` returns the reference's own modules (``graphsage.aggregators.MeanAggregator``,
``graphsage.encoders.Encoder``, ``graphsage.model.SupervisedGraphSage``) imported from ``oracle/_ref`` under a
private package name, with the ``random.sample`` shim of SURVEY.md s8c installed: on Python >= 3.11
``random.sample(set, k)`` (graphsage/aggregators.py:44) raises TypeError; CPython <= 3.10 converted the set to a
tuple first (Lib/random.py), which is what the shim does.  The copied files are not touched.

``build_two_layer`` wires the model exactly as the reference driver does (graphsage/model.py:214-227, with the
SAGE-concat options the benchmark config names) and ``train_step`` is its timed unit (model.py:245-250).
"""
import importlib
import os
import random
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_loaded = None


def available():
    from . import build_ref
    return build_ref.verify()


def _install_shim():
    if getattr(random.sample, "_gsage_set_shim", False):
        return
    orig = random.sample

    def sample(population, k, **kw):
        if isinstance(population, (set, frozenset)):
            population = tuple(population)          # Lib/random.py of CPython <= 3.10 did exactly this
        return orig(population, k, **kw)
    sample._gsage_set_shim = True
    random.sample = sample


def load():
    """(aggregators, encoders, model) modules of the unmodified reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is missing or does not match its manifest: run oracle/build_ref.py "
                           "in the build container (it copies the reference's sources from /root/reference)")
    _install_shim()
    # the reference's modules import each other as ``graphsage.*``; our product package has the same name,
    # so the reference is imported with oracle/_ref FIRST on sys.path and then re-registered under a private
    # name, leaving ``graphsage`` free for (or restored to) the product package.
    saved = {k: v for k, v in sys.modules.items() if k == "graphsage" or k.startswith("graphsage.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            agg = importlib.import_module("graphsage.aggregators")
            enc = importlib.import_module("graphsage.encoders")
            mod = importlib.import_module("graphsage.model")
    finally:
        sys.path.remove(REF_DIR)
        for k in [k for k in sys.modules if k == "graphsage" or k.startswith("graphsage.")]:
            sys.modules["_gsage_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    assert os.path.realpath(agg.__file__).startswith(os.path.realpath(REF_DIR))
    _loaded = (agg, enc, mod)
    return _loaded


def build_two_layer(table, adj_lists, feat_dim, d1, d2, num_classes, k1, k2, gcn=False, weights=None):
    """model.py:214-227 with the reference's own classes: frozen nn.Embedding table, two Encoders joined by the
    closure ``lambda nodes: enc1(nodes).t()``, SupervisedGraphSage on top.  ``weights`` = (w1, w2, wc) to copy in."""
    import torch
    import torch.nn as nn
    agg_m, enc_m, model_m = load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        features = nn.Embedding(table.shape[0], feat_dim)
        features.weight = nn.Parameter(torch.as_tensor(table, dtype=torch.float32), requires_grad=False)
        agg1 = agg_m.MeanAggregator(features, cuda=False)
        enc1 = enc_m.Encoder(features, feat_dim, d1, adj_lists, agg1, num_sample=k1, gcn=gcn, cuda=False)
        agg2 = agg_m.MeanAggregator(lambda nodes: enc1(nodes).t(), cuda=False)
        enc2 = enc_m.Encoder(lambda nodes: enc1(nodes).t(), enc1.embed_dim, d2, adj_lists, agg2, num_sample=k2,
                             base_model=enc1, gcn=gcn, cuda=False)
        model = model_m.SupervisedGraphSage(num_classes, enc2)
    if weights is not None:
        with torch.no_grad():
            enc1.weight.copy_(weights[0]); enc2.weight.copy_(weights[1]); model.weight.copy_(weights[2])
    return model


def build_stack(table, adj_lists, feat_dim, hidden, fanouts, num_classes, gcn=False):
    """Any depth, wired the way model.py:218-222 wires two layers: layer l's ``features`` is the closure
    ``lambda nodes: enc_{l-1}(nodes).t()``.  hidden / fanouts: innermost layer first."""
    import torch
    import torch.nn as nn
    agg_m, enc_m, model_m = load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        features = nn.Embedding(table.shape[0], feat_dim)
        features.weight = nn.Parameter(torch.as_tensor(table, dtype=torch.float32), requires_grad=False)
        encs = []
        for layer, (dim, k) in enumerate(zip(hidden, fanouts)):
            if layer == 0:
                enc = enc_m.Encoder(features, feat_dim, dim, adj_lists, agg_m.MeanAggregator(features, cuda=False),
                                    num_sample=k, gcn=gcn, cuda=False)
            else:
                below = encs[-1]
                feats = (lambda b: (lambda nodes: b(nodes).t()))(below)
                enc = enc_m.Encoder(feats, below.embed_dim, dim, adj_lists, agg_m.MeanAggregator(feats, cuda=False),
                                    num_sample=k, base_model=below, gcn=gcn, cuda=False)
            encs.append(enc)
        return model_m.SupervisedGraphSage(num_classes, encs[-1])


def make_optimizer(model, lr):
    import torch
    return torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)      # model.py:237


def train_step(model, optimizer, nodes, labels):
    """model.py:245-250 verbatim: zero_grad, loss, backward, step.  Returns the loss tensor."""
    import torch
    optimizer.zero_grad()
    loss = model.loss(nodes, torch.LongTensor(labels))
    loss.backward()
    optimizer.step()
    return loss
