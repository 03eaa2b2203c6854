"""Host-side loaders / initialisers (graphsage/data.py) against digests of the UNMODIFIED reference
loaders (tests/golden/loaders.json, made by tests/golden/make_golden_loaders.py on the same synthetic
files, which graphsage.data.write_synthetic_dataset regenerates here bit-for-bit), and the
adj_lists -> CSR conversion that feeds the device sampler."""
import hashlib
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loaders.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def datasets(tmp_path_factory):
    from graphsage import data as D
    g = json.load(open(GOLDEN))
    root = str(tmp_path_factory.mktemp("datasets"))
    for ds, seed in g["seeds"].items():
        D.write_synthetic_dataset(ds, root, seed=seed)
    return root, g["cases"]


def test_loaders_match_reference_digests(datasets):
    from graphsage import data as D
    root, cases = datasets
    assert len(cases) >= 14
    for key, want in cases.items():
        ds, init = key.split("/")
        np.random.seed(5)
        feat, labels, adj = D.load_dataset(ds, 100, init, root=root)
        assert list(feat.shape) == want["shape"], key
        assert labels.shape == (D.DATASETS[ds]["num_nodes"], 1) and labels.dtype == np.int64
        assert sha(labels.astype(np.int64)) == want["labels_sha"], key
        edges = np.array(sorted((u, v) for u, nb in adj.items() for v in nb), dtype=np.int64).reshape(-1, 2)
        assert edges.shape[0] == want["num_adj_entries"] and sha(edges) == want["adj_sha"], key
        if "feat_values" in want:
            np.testing.assert_allclose(feat.ravel(), np.array(want["feat_values"]), rtol=1e-9, atol=1e-15, err_msg=key)
        else:
            assert abs(float(feat.sum()) - want["feat_sum"]) <= 1e-9 * max(1.0, abs(want["feat_sum"])), key
            assert sha(feat.astype(np.float64)) == want["feat_sha"], key


def test_named_loaders_and_errors(datasets, tmp_path):
    from graphsage import data as D
    root, _ = datasets
    f1, l1, a1 = D.load_cora(root=root)
    f2, l2, a2 = D.load_dataset("cora", root=root)
    assert np.array_equal(f1, f2) and np.array_equal(l1, l2) and a1 == a2
    with pytest.raises(FileNotFoundError, match="MISSING_LARGE_BLOBS"):
        D.load_pubmed(100, "None", root=str(tmp_path))
    with pytest.raises(ValueError, match="unknown initializer"):
        D.load_dataset("cora", 100, "bogus", root=root)


def test_eigen_decomposition_initialiser_is_spectral(tmp_path):
    """model.py:326-344: rows of feat_data are the node's coordinates in the leading eigenvectors."""
    from graphsage import data as D
    root = str(tmp_path)
    D.write_synthetic_dataset("cora", root, seed=3)
    feat, _, adj = D.load_dataset("cora", 8, "eigen_decomposition", root=root)
    assert feat.shape == (2708, 8) and os.path.exists(os.path.join(root, "cora", "cora_eigenvector.npy"))
    a = np.zeros((2708, 2708))
    for u, nb in adj.items():
        a[u, list(nb)] = 1
    lam = np.array([feat[:, j] @ a @ feat[:, j] / (feat[:, j] @ feat[:, j]) for j in range(8)])
    for j in range(8):                                   # A v = lambda v
        assert np.abs(a @ feat[:, j] - lam[j] * feat[:, j]).max() < 1e-6 * max(1.0, abs(lam[j]))
    assert np.all(np.diff(lam) <= 1e-8)                 # descending eigenvalues
    again, _, _ = D.load_dataset("cora", 8, "eigen_decomposition", root=root)     # from the .npy cache
    assert np.array_equal(again, feat)


def test_adj_lists_to_csr_round_trip(datasets):
    """graph.CSRGraph.from_adj_lists: rows sorted ascending, symmetric, same sets back."""
    from graphsage import data as D
    from graphsage.graph import CSRGraph
    root, _ = datasets
    _, _, adj = D.load_dataset("citeseer", 100, "shared", root=root)
    g = CSRGraph.from_adj_lists(adj, num_nodes=3312, device="cpu")
    assert g.num_nodes == 3312 and g.num_entries == sum(len(v) for v in adj.values())
    col, rp = g.col.numpy(), g.rowptr_host
    for v in (0, 1, 17, 3311):
        row = col[rp[v]:rp[v + 1]]
        assert np.all(np.diff(row) > 0) and set(row.tolist()) == adj.get(v, set())
    back = g.to_adj_lists()
    assert all(back[v] == adj.get(v, set()) for v in range(3312))
    assert g.max_degree == max(len(v) for v in adj.values())
