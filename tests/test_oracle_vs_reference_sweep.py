"""The oracle port (oracle/ref_path.py) against the UNMODIFIED reference modules (oracle/_ref, copied byte for byte from
/root/reference by oracle/build_ref.py) over a sweep of configurations, bit for bit: same Python RNG state in, same loss
every step and same weights after SGD out.  This is the pin that makes the port a valid checker for the GPU path at
shapes the committed golden files do not hold -- SAGE concat and GCN encoders, fan-outs below / at / above the degrees,
the un-sampled neighbourhood (num_sample=None), repeated targets in a batch, two and three layers.
Skipped where oracle/_ref did not travel."""
import contextlib
import io
import random

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref, ref_runtime as RR
    if not build_ref.verify():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return RR


def _graph(rng, n, pairs, isolated=0):
    """random pairs + a ring over the non-isolated nodes (every one of them has a neighbour); the last ``isolated`` none"""
    adj = {i: set() for i in range(n)}
    m = n - isolated
    ring = np.stack([np.arange(m), (np.arange(m) + 1) % m], axis=1)
    for a, b in np.concatenate([rng.integers(0, m, (pairs, 2)), ring]):
        adj[int(a)].add(int(b))
        adj[int(b)].add(int(a))
    return adj


def _chain(model):
    encs = [model.enc]
    while hasattr(encs[-1], "base_model"):
        encs.append(encs[-1].base_model)
    return encs[::-1]                                    # innermost first


@pytest.mark.parametrize("gcn,f,d1,d2,c,k1,k2,pairs", [
    (False, 20, 16, 16, 5, 4, 6, 3000),                  # the shape of the existing pin
    (True, 33, 8, 12, 3, 5, 5, 2500),                    # as the fork runs it: no self term
    (False, 7, 128, 128, 41, 10, 25, 6000),              # BASELINE widths / fan-outs, most rows sampled
    (False, 12, 9, 5, 2, 50, 60, 900),                   # fan-outs above every degree: take-all through the sampler branch
    (True, 5, 6, 6, 4, None, None, 1200),                # num_sample=None (aggregators.py:42): full neighbourhoods
    (False, 16, 4, 4, 7, 1, 1, 2000),                    # one neighbour per node
])
def test_two_layer_port_equals_reference_bit_for_bit(ref, gcn, f, d1, d2, c, k1, k2, pairs):
    from oracle import ref_path as R
    rng = np.random.default_rng(pairs + f)
    n = 300
    adj = _graph(rng, n, pairs)
    table = torch.randn(n, f, generator=torch.Generator().manual_seed(f))
    k_in1, k_in2 = (f, d1) if gcn else (2 * f, 2 * d1)
    ws = [torch.randn(s, generator=torch.Generator().manual_seed(i)) * 0.1
          for i, s in enumerate([(d1, k_in1), (d2, k_in2), (c, d2)])]
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.build_two_layer(table, adj, f, d1, d2, c, k1, k2, gcn=gcn, weights=ws)
    opt = ref.make_optimizer(m, 0.7)
    o = R.TwoLayerModel(table, adj, adj, d1, d2, c, k1, k2, gcn=gcn, w1=ws[0].clone(), w2=ws[1].clone(), wc=ws[2].clone())
    for step in range(3):
        nodes = list(rng.integers(0, n, 48))
        if step == 1:
            nodes[5] = nodes[0]                            # a target that appears twice in the batch
        random.seed(11 + step)
        l_ref = float(ref.train_step(m, opt, nodes, labels[np.array(nodes)]).detach())
        random.seed(11 + step)
        l_port = float(o.train_step(nodes, labels[np.array(nodes)], lr=0.7))
        assert l_ref == l_port, (step, l_ref, l_port)
    e1, e2 = _chain(m)
    assert torch.equal(e1.weight.detach(), o.enc1.weight.detach())
    assert torch.equal(e2.weight.detach(), o.enc2.weight.detach())
    assert torch.equal(m.weight.detach(), o.weight.detach())


@pytest.mark.parametrize("gcn,dims,ks", [(False, (8, 8, 6), (5, 4, 3)), (True, (16, 4, 4), (3, 3, 10)),
                                        (False, (6, 6, 6, 6), (2, 2, 2, 2))])
def test_deeper_stacks_port_equals_reference_bit_for_bit(ref, gcn, dims, ks):
    """BASELINE config 5 is three layers deep (model.py:218-227 extended by the closure recursion); four for good measure."""
    from oracle import ref_path as R
    rng = np.random.default_rng(len(dims) + int(gcn))
    n, f, c = 200, 10, 4
    adj = _graph(rng, n, 1500, isolated=0)
    table = torch.randn(n, f, generator=torch.Generator().manual_seed(2))
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.build_stack(table, adj, f, list(dims), list(ks), c, gcn=gcn)
    encs = _chain(m)
    assert len(encs) == len(dims)
    o = R.StackedModel(table, [adj] * len(dims), list(dims), c, list(ks), gcn=gcn,
                       weights=[e.weight.detach().clone() for e in encs], wc=m.weight.detach().clone())
    opt = ref.make_optimizer(m, 0.5)
    for step in range(2):
        nodes = list(rng.integers(0, n, 24))
        random.seed(step)
        l_ref = float(ref.train_step(m, opt, nodes, labels[np.array(nodes)]).detach())
        random.seed(step)
        l_port = float(o.train_step(nodes, labels[np.array(nodes)], lr=0.5))
        assert l_ref == l_port, (step, l_ref, l_port)
    for e, layer in zip(encs, o.layers):
        assert torch.equal(e.weight.detach(), layer.weight.detach())
    assert torch.equal(m.weight.detach(), o.weight.detach())


def test_isolated_targets_are_nan_in_both(ref):
    """aggregators.py:60-61 divides a zero row by its zero sum: a node without neighbours poisons the loss with NaN in
    the reference, and the port restates exactly that (the GPU path's documented deviation is zeros)."""
    from oracle import ref_path as R
    rng = np.random.default_rng(3)
    n, f = 60, 6
    adj = _graph(rng, n, 200, isolated=2)
    table = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    ws = [torch.randn(s, generator=torch.Generator().manual_seed(i)) * 0.1 for i, s in enumerate([(4, 2 * f), (4, 8), (3, 4)])]
    labels = rng.integers(0, 3, (n, 1)).astype(np.int64)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.build_two_layer(table, adj, f, 4, 4, 3, 3, 3, weights=ws)
    o = R.TwoLayerModel(table, adj, adj, 4, 4, 3, 3, 3, w1=ws[0].clone(), w2=ws[1].clone(), wc=ws[2].clone())
    nodes = [0, 1, n - 1, 5]                               # n - 1 has no neighbours
    random.seed(0)
    l_ref = float(m.loss(nodes, torch.LongTensor(labels[np.array(nodes)])).detach())
    random.seed(0)
    l_port = float(o.loss(nodes, labels[np.array(nodes)]).detach())
    assert np.isnan(l_ref) and np.isnan(l_port)
    clean = [0, 1, 2, 5]
    random.seed(0)
    a = float(m.loss(clean, torch.LongTensor(labels[np.array(clean)])).detach())
    random.seed(0)
    b = float(o.loss(clean, labels[np.array(clean)]).detach())
    assert a == b and np.isfinite(a)
