"""Seeded inputs of the BASELINE-shape golden cases, shared by the generator
(tests/golden/make_golden_baseline_shapes.py, build container, runs the UNMODIFIED reference) and the GPU parity
test (tests/test_gpu_baseline_shapes.py, GPU box).  Only the reference's OUTPUTS and the replayed neighbour tiles
are stored in the .npz; the big inputs (feature table, initial weights, labels) are regenerated from these seeds on
both sides, which keeps the fixtures small.

Cases (BASELINE.json configs at their real widths, on a 3000-node graph so that the reference finishes in seconds):
  bs_reddit    F=602,  hidden 128/128, 41 classes, fan-out 25 (targets) / 10 (hop 1), SAGE concat   (config 4)
  bs_pubmed    F=500,  hidden 128/128,  3 classes, fan-out 10 (targets) / 25 (hop 1), SAGE concat   (config 2, k swapped)
  bs_citeseer  F=3703, hidden 128/128,  6 classes, fan-out 5 / 5, gcn=True encoders, sparse 0/1 bag-of-words rows (config 3)
"""
import numpy as np

CASES = {
    "bs_reddit": dict(n=3000, f=602, d1=128, d2=128, c=41, k1=10, k2=25, b=64, gcn=False, seed=602, binary=False),
    "bs_pubmed": dict(n=3000, f=500, d1=128, d2=128, c=3, k1=25, k2=10, b=64, gcn=False, seed=500, binary=False),
    "bs_citeseer": dict(n=3312, f=3703, d1=128, d2=128, c=6, k1=5, k2=5, b=64, gcn=True, seed=3703, binary=True),
}


def xavier(rng, rows, cols):
    a = np.sqrt(6.0 / (rows + cols))              # torch.nn.init.xavier_uniform_ bound (encoders.py:36)
    return rng.uniform(-a, a, (rows, cols)).astype(np.float32)


def inputs(name):
    """table [n, f], labels [n, 1] int64, nodes [b] int64, w1, w2, wc -- all from the case's seed."""
    p = CASES[name]
    rng = np.random.default_rng(p["seed"])
    if p["binary"]:
        table = (rng.random((p["n"], p["f"])) < 0.01).astype(np.float32)      # ~37 words per document
    else:
        table = rng.standard_normal((p["n"], p["f"])).astype(np.float32)
    labels = rng.integers(0, p["c"], (p["n"], 1)).astype(np.int64)
    nodes = rng.permutation(p["n"])[:p["b"]].astype(np.int64)
    k1_in = p["f"] if p["gcn"] else 2 * p["f"]
    k2_in = p["d1"] if p["gcn"] else 2 * p["d1"]
    w1 = xavier(rng, p["d1"], k1_in)
    w2 = xavier(rng, p["d2"], k2_in)
    wc = xavier(rng, p["c"], p["d2"])
    return dict(table=table, labels=labels, nodes=nodes, w1=w1, w2=w2, wc=wc)


GW1_COLS = 4          # bs_citeseer stores every 4th column of the layer-1 weight gradient (128 x 3703 otherwise)
