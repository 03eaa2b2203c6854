"""Dataset loaders and node-feature initialisers of zjzijielu/graphsage-simple, host side.

Same inputs, same outputs as the reference's ``load_cora`` / ``load_citeseer`` / ``load_pubmed``
(graphsage/model.py:261-346, 88-182, 404-494) and ``extract_deepwalk_embeddings`` (model.py:71-86):

    feat_data  float64 ndarray [N, F]        (F depends on the initialiser, see ``init_features``)
    labels     int64   ndarray [N, 1]
    adj_lists  defaultdict(set), symmetric   (model.py:303-310)

but written once, table-driven, instead of three near-identical functions.  These run once per
process on the host; the hot path only ever sees their result after ``graph.CSRGraph.from_adj_lists``
(adjacency -> device CSR) and the padded fp32 upload of ``feat_data``.

Text formats (cora/README:19-29 of the reference; Pubmed-Diabetes tab files):
  <name>.content   one node per line:  <paper_id> <f_1> ... <f_F> <class_label>     (whitespace)
  <name>.cites     one edge per line:  <cited_id> <citing_id>
  Pubmed NODE tab  line 1 header, line 2 "cat=..\\tnumeric:<word>:0.0\\t...", then
                   <id>\\tlabel=<1..3>\\t<word>=<tfidf>\\t...\\tsummary=...
  Pubmed cites tab two header lines, then  <n>\\tpaper:<id>\\t|\\tpaper:<id>
  <name>.embeddings  word2vec text: "N D" header, then "<id> v_1 .. v_D"
"""
import os
from collections import defaultdict

import numpy as np

# per-dataset constants the reference hard-codes in run_model (model.py:186-190) and in the loaders
DATASETS = {
    "cora": dict(num_nodes=2708, num_classes=7, attr_dim=1433, num_samples=(5, 5), kind="content",
                 content="cora/cora.content", cites="cora/cora.cites", embeddings="cora/cora.embeddings",
                 eigen_cache="cora/cora_eigenvector.npy", skip_unknown=False, embedding_ids="raw"),
    "citeseer": dict(num_nodes=3312, num_classes=6, attr_dim=3703, num_samples=(5, 5), kind="content",
                     content="citeseer/citeseer.content", cites="citeseer/citeseer.cites",
                     embeddings="citeseer/citeseer.embeddings", eigen_cache="citeseer/citeseer_eigenvector.npy",
                     skip_unknown=True, embedding_ids="index"),
    "pubmed": dict(num_nodes=19717, num_classes=3, attr_dim=500, num_samples=(10, 25), kind="pubmed",
                   content="pubmed-data/Pubmed-Diabetes.NODE.paper.tab",
                   cites="pubmed-data/Pubmed-Diabetes.DIRECTED.cites.tab", embeddings="pubmed-data/Pubmed.embeddings",
                   eigen_cache="pubmed-data/pubmed_eigenvector.npy", skip_unknown=False, embedding_ids="raw"),
}

INITIALIZERS = ("None", "1hot", "random_normal", "shared", "node_degree", "pagerank", "eigen_decomposition", "deepwalk")


def read_embeddings(filename, node_map, ids="raw"):
    """word2vec text format -> [N, D] float64 (model.py:71-86).  ``ids="raw"`` maps the first column
    through ``node_map`` (Cora, Pubmed), ``"index"`` takes it as the row index (Citeseer)."""
    feat = None
    with open(filename) as f:
        for i, line in enumerate(f):
            tok = line.split()
            if i == 0:
                feat = np.zeros((int(tok[0]), int(tok[1])))
                continue
            row = node_map[tok[0]] if ids == "raw" else int(tok[0])
            feat[row, :] = [float(x) for x in tok[1:]]
    return feat


def _read_content(path, num_nodes, num_feats, want_features):
    """Planetoid-style .content file (model.py:270-287): line order defines node ids, first appearance
    defines class ids."""
    feat = np.zeros((num_nodes, num_feats)) if want_features else None
    labels = np.empty((num_nodes, 1), dtype=np.int64)
    node_map, label_map = {}, {}
    with open(path) as fp:
        for i, line in enumerate(fp):
            tok = line.split()
            if want_features:
                feat[i, :] = [float(x) for x in tok[1:-1]]
            node_map[tok[0]] = i
            labels[i] = label_map.setdefault(tok[-1], len(label_map))
    return feat, labels, node_map


def _read_pubmed_nodes(path, num_nodes, num_feats, want_features):
    """Pubmed-Diabetes.NODE.paper.tab (model.py:414-433)."""
    feat = np.zeros((num_nodes, num_feats)) if want_features else None
    labels = np.empty((num_nodes, 1), dtype=np.int64)
    node_map = {}
    with open(path) as fp:
        fp.readline()
        header = fp.readline().split("\t")
        feat_map = {entry.split(":")[1]: i - 1 for i, entry in enumerate(header)} if want_features else None
        for i, line in enumerate(fp):
            tok = line.split("\t")
            node_map[tok[0]] = i
            labels[i] = int(tok[1].split("=")[1]) - 1
            if want_features:
                for item in tok[2:-1]:
                    word, val = item.split("=")
                    feat[i][feat_map[word]] = float(val)
    return feat, labels, node_map


def _read_edges(spec, path, node_map):
    """Symmetric adjacency (model.py:303-310, 136-147, 447-455) plus the (u, v) list in file order."""
    adj = defaultdict(set)
    edges = []
    with open(path) as fp:
        if spec["kind"] == "pubmed":
            fp.readline()
            fp.readline()
        for line in fp:
            if spec["kind"] == "pubmed":
                tok = line.strip().split("\t")
                a, b = tok[1].split(":")[1], tok[-1].split(":")[1]
            else:
                tok = line.split()
                a, b = tok[0], tok[1]
            if spec["skip_unknown"] and (a not in node_map or b not in node_map):
                continue                                    # citeseer.cites names papers without content
            u, v = node_map[a], node_map[b]
            adj[u].add(v)
            adj[v].add(u)
            edges.append((u, v))
    return adj, edges


def init_features(initializer, num_nodes, feature_dim, adj_lists, edges, node_map, spec, root):
    """The fork's "alternative node-feature initialisations" (model.py:289-301, 315-344):
      1hot                 identity [N, N]
      random_normal        np.random.normal(0, 1, [N, feature_dim])   (global numpy RNG, seeded by run_model)
      shared               ones [N, feature_dim]
      node_degree          one-hot of the degree, [N, max_degree + 1]
      pagerank             networkx.pagerank value, [N, 1]
      eigen_decomposition  leading ``feature_dim`` eigenvectors of the adjacency matrix (cached .npy)
      deepwalk             rows of the .embeddings file
    """
    if initializer == "1hot":
        return np.eye(num_nodes)
    if initializer == "random_normal":
        return np.random.normal(0, 1, (num_nodes, feature_dim))
    if initializer == "shared":
        return np.ones((num_nodes, feature_dim))
    if initializer == "node_degree":
        width = max(len(v) for v in adj_lists.values()) + 1
        feat = np.zeros((num_nodes, width))
        for k, v in adj_lists.items():
            feat[k, len(v)] = 1
        return feat
    if initializer == "deepwalk":
        return read_embeddings(os.path.join(root, spec["embeddings"]), node_map, spec["embedding_ids"])
    if initializer in ("pagerank", "eigen_decomposition"):
        import networkx as nx
        g = nx.Graph()
        g.add_nodes_from(node_map.values())
        g.add_edges_from(edges)
        if initializer == "pagerank":
            feat = np.zeros((num_nodes, 1))
            for k, v in nx.pagerank(g).items():
                feat[k, 0] = v
            return feat
        cache = os.path.join(root, spec["eigen_cache"])
        if os.path.exists(cache):
            vecs = np.load(cache)
        else:
            w, v = np.linalg.eig(nx.to_numpy_array(g))
            vecs = v.transpose()[np.argsort(w)[::-1]][:1000]            # top 1000 kept (model.py:333-336)
            np.save(cache, vecs)
        assert feature_dim <= 1000
        return np.ascontiguousarray(np.real(vecs[:feature_dim]).T[:num_nodes])
    raise ValueError("unknown initializer %r (one of %s)" % (initializer, ", ".join(INITIALIZERS)))


def load_dataset(name, feature_dim=100, initializer="None", root="."):
    """(feat_data, labels, adj_lists) of ``name`` in {"cora", "citeseer", "pubmed"} -- the merged
    equivalent of load_cora / load_citeseer / load_pubmed.  ``root`` is the directory that holds the
    reference's ``cora/``, ``citeseer/`` and ``pubmed-data/`` folders (the reference uses the CWD)."""
    spec = DATASETS[name]
    n = spec["num_nodes"]
    plain = initializer == "None"
    reader = _read_pubmed_nodes if spec["kind"] == "pubmed" else _read_content
    path = os.path.join(root, spec["content"])
    if not os.path.exists(path):
        raise FileNotFoundError("%s not found: the node-content files are large blobs that are not part of every "
                                "checkout of the reference (.MISSING_LARGE_BLOBS); pass --data_root" % path)
    feat, labels, node_map = reader(path, n, spec["attr_dim"], plain)
    adj, edges = _read_edges(spec, os.path.join(root, spec["cites"]), node_map)
    if not plain:
        feat = init_features(initializer, n, feature_dim, adj, edges, node_map, spec, root)
    return feat, labels, adj


def load_cora(feature_dim=100, initializer="None", root="."):
    return load_dataset("cora", feature_dim, initializer, root)


def load_citeseer(feature_dim=100, initializer="None", root="."):
    return load_dataset("citeseer", feature_dim, initializer, root)


def load_pubmed(feature_dim=100, initializer="None", root="."):
    return load_dataset("pubmed", feature_dim, initializer, root)


# ---- synthetic data in the reference's file formats (tests, demos: the real .content blobs are absent) ----
def write_synthetic_dataset(name, root, seed=0, avg_degree=4.0, num_nodes=None, attr_dim=None, planted=True):
    """Write a random dataset with the node count / widths the reference hard-codes for ``name`` into
    ``root`` using the reference's text formats, so that both the reference's loaders and ours can read
    it.  Labels follow a planted partition (edges mostly inside a class, features carry a noisy class
    signature) so that a trained model's F1 is far from chance.  Returns the class of every node."""
    spec = DATASETS[name]
    n = num_nodes or spec["num_nodes"]
    f = attr_dim or spec["attr_dim"]
    c = spec["num_classes"]
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, c, n)
    ids = rng.permutation(np.arange(10 * n))[:n] + 1000                      # arbitrary raw paper ids
    os.makedirs(os.path.dirname(os.path.join(root, spec["content"])), exist_ok=True)
    m = int(n * avg_degree / 2)
    src = rng.integers(0, n, m)
    same = rng.random(m) < (0.8 if planted else 0.0)
    dst = np.where(same, 0, rng.integers(0, n, m))
    by_cls = [np.nonzero(cls == k)[0] for k in range(c)]
    for k in range(c):
        sel = np.nonzero(same & (cls[src] == k))[0]
        dst[sel] = by_cls[k][rng.integers(0, len(by_cls[k]), sel.size)]
    ring = np.arange(n)                                                      # every node gets an edge
    src, dst = np.concatenate([src, ring]), np.concatenate([dst, (ring + 1) % n])
    sig = rng.random((c, f)) < 0.08                                          # class signatures
    if spec["kind"] == "content":
        names = ["Class_%d" % k for k in range(c)]
        with open(os.path.join(root, spec["content"]), "w") as fp:
            for i in range(n):
                row = (rng.random(f) < 0.01) | (sig[cls[i]] & (rng.random(f) < 0.5))
                fp.write("%d\t%s\t%s\n" % (ids[i], "\t".join("1" if b else "0" for b in row), names[cls[i]]))
        with open(os.path.join(root, spec["cites"]), "w") as fp:
            for a, b in zip(src, dst):
                fp.write("%d\t%d\n" % (ids[a], ids[b]))
            if spec["skip_unknown"]:
                fp.write("unknown_paper\t%d\n%d\tghost\n" % (ids[0], ids[1]))      # ids without content rows
    else:
        words = ["w-%d" % j for j in range(f)]
        with open(os.path.join(root, spec["content"]), "w") as fp:
            fp.write("NODE\tpaper\n")
            fp.write("cat=1,2,3:label\t" + "\t".join("numeric:%s:0.0" % w for w in words) + "\tstring:summary\n")
            for i in range(n):
                on = np.nonzero((rng.random(f) < 0.01) | (sig[cls[i]] & (rng.random(f) < 0.5)))[0]
                vals = rng.random(on.size) * 0.1
                fp.write("%d\tlabel=%d\t%s\tsummary=%s\n" % (
                    ids[i], cls[i] + 1, "\t".join("%s=%.6f" % (words[j], v) for j, v in zip(on, vals)),
                    ",".join(words[j] for j in on)))
        with open(os.path.join(root, spec["cites"]), "w") as fp:
            fp.write("DIRECTED\tcites\nNO_FEATURES\n")
            for e, (a, b) in enumerate(zip(src, dst)):
                fp.write("%d\tpaper:%d\t|\tpaper:%d\n" % (e, ids[a], ids[b]))
    with open(os.path.join(root, spec["embeddings"]), "w") as fp:              # DeepWalk-style text
        d = 16
        fp.write("%d %d\n" % (n, d))
        emb = rng.standard_normal((n, d))
        for i in rng.permutation(n):
            key = ids[i] if spec["embedding_ids"] == "raw" else i
            fp.write("%d %s\n" % (key, " ".join("%.6f" % x for x in emb[i])))
    return cls
