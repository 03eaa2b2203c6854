#!/usr/bin/env python
"""Golden digests for the dataset loaders / initialisers: the UNMODIFIED reference loaders
(/root/reference/graphsage/model.py: load_cora 261-346, load_citeseer 88-182, load_pubmed 404-494) run
on synthetic datasets written in the reference's own text formats by
graphsage.data.write_synthetic_dataset (the real .content blobs are missing from the checkout,
.MISSING_LARGE_BLOBS).  Writes tests/golden/loaders.json.

    python tests/golden/make_golden_loaders.py        # in the build container only (needs /root/reference)
"""
import hashlib
import json
import os
import random
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "graphsage-simple_b200"))
warnings.filterwarnings("ignore")

CASES = [("cora", i) for i in ("None", "1hot", "random_normal", "shared", "node_degree", "pagerank", "deepwalk")] + \
        [("citeseer", i) for i in ("None", "node_degree", "pagerank", "deepwalk")] + \
        [("pubmed", i) for i in ("None", "node_degree", "random_normal")]
SEEDS = {"cora": 11, "citeseer": 12, "pubmed": 13}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(feat, labels, adj, initializer):
    edges = np.array(sorted((u, v) for u, nb in adj.items() for v in nb), dtype=np.int64).reshape(-1, 2)
    d = {"shape": list(feat.shape), "labels_sha": sha(labels.astype(np.int64)), "adj_sha": sha(edges),
         "num_adj_entries": int(edges.shape[0]), "label_hist": np.bincount(labels.ravel()).tolist()}
    if initializer == "pagerank":
        d["feat_values"] = [float(x) for x in feat.ravel()]          # iterative floats: compared with a tolerance
    else:
        d["feat_sha"] = sha(feat.astype(np.float64))
        d["feat_sum"] = float(feat.sum())
    return d


def main():
    from graphsage import data as D
    sys.path.insert(0, "/root/reference")
    _orig = random.sample
    random.sample = lambda pop, k: _orig(tuple(pop) if isinstance(pop, (set, frozenset)) else pop, k)
    import graphsage.model as ref_pkg_guard  # noqa: F401  (our drop-in, already imported through D)
    # import the REFERENCE model module under another name (our package shadows "graphsage")
    import importlib.util
    for name in ("aggregators", "encoders"):
        spec = importlib.util.spec_from_file_location("graphsage_ref." + name, "/root/reference/graphsage/%s.py" % name)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["graphsage_ref." + name] = mod
        spec.loader.exec_module(mod)
    src = open("/root/reference/graphsage/model.py").read()
    src = src.replace("from graphsage.encoders", "from graphsage_ref.encoders").replace(
        "from graphsage.aggregators", "from graphsage_ref.aggregators")
    ref = type(sys)("graphsage_ref.model")
    exec(compile(src, "/root/reference/graphsage/model.py", "exec"), ref.__dict__)    # executed, not copied
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for ds in ("cora", "citeseer", "pubmed"):
            D.write_synthetic_dataset(ds, tmp, seed=SEEDS[ds])
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            for ds, init in CASES:
                np.random.seed(5)
                feat, labels, adj = getattr(ref, "load_" + ds)(100, init)
                out["%s/%s" % (ds, init)] = digest(np.asarray(feat), labels, adj, init)
                print(ds, init, out["%s/%s" % (ds, init)]["shape"])
        finally:
            os.chdir(cwd)
    json.dump({"seeds": SEEDS, "cases": out}, open(os.path.join(HERE, "loaders.json"), "w"))


if __name__ == "__main__":
    main()
