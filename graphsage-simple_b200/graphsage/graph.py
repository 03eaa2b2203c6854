"""CSR graph container resident in HBM.

The reference keeps the graph as ``adj_lists: defaultdict(set)`` built by its loaders
(graphsage/model.py:303-310) and walks it per node in Python (graphsage/encoders.py:47).
Here the same adjacency is converted ONCE to CSR -- ``rowptr`` int64 [N+1], ``col`` int32
[sum deg], every row sorted ascending -- uploaded, and sampled from on the device
(gs_sample_csr).  A host copy of rowptr is kept for sizing take-all tiles."""
import numpy as np
import torch


class CSRGraph:
    """Whole graph on one device.  ``sample`` is the interface the aggregator uses; the partitioned
    variant (sharded.ShardedCSR) answers the same call through an all-to-all exchange."""

    def __init__(self, rowptr, col, device="cuda"):
        rowptr = np.ascontiguousarray(np.asarray(rowptr, dtype=np.int64))
        col = np.ascontiguousarray(np.asarray(col, dtype=np.int32))
        assert rowptr.ndim == 1 and rowptr[0] == 0 and rowptr[-1] == col.shape[0]
        self.num_nodes = rowptr.shape[0] - 1
        self.num_entries = int(col.shape[0])
        self.rowptr_host = rowptr
        deg = np.diff(rowptr)
        self.max_degree = int(deg.max()) if self.num_nodes else 0
        self.min_degree = int(deg.min()) if self.num_nodes else 0
        self.device = torch.device(device)
        self.rowptr = torch.from_numpy(rowptr).to(self.device)
        self.col = torch.from_numpy(col if col.size else np.zeros(1, np.int32)).to(self.device)

    def sample(self, ids, k, add_self=False, seed=0, step=0, tag=0, width=None, step_dev=None):
        """Fixed-width tile (idx [n, width], cnt [n]) of sampled neighbours of ``ids`` (int32 CUDA)."""
        from . import ops
        return ops.sample_csr(self.rowptr, self.col, self.num_nodes, ids, k, add_self=add_self, seed=seed,
                              step=step, tag_head=tag, width=width, step_dev=step_dev)

    # ---- constructors ------------------------------------------------------------------
    @classmethod
    def from_adj_lists(cls, adj_lists, num_nodes=None, device="cuda"):
        """``adj_lists``: mapping int -> iterable of int (the reference's defaultdict(set))."""
        keys = [int(k) for k in adj_lists.keys()]
        hi = max(keys) if keys else -1
        for k in keys:
            nb = adj_lists[k]
            if len(nb):
                hi = max(hi, max(int(x) for x in nb))
        n = max(int(num_nodes), hi + 1) if num_nodes is not None else hi + 1
        deg = np.zeros(n + 1, dtype=np.int64)
        for k in keys:
            deg[k + 1] = len(adj_lists[k])
        rowptr = np.cumsum(deg)
        col = np.empty(int(rowptr[-1]), dtype=np.int32)
        for k in keys:
            nb = sorted(int(x) for x in adj_lists[k])
            col[rowptr[k]:rowptr[k] + len(nb)] = nb
        return cls(rowptr, col, device)

    @classmethod
    def from_edges(cls, src, dst, num_nodes, symmetric=True, device="cuda"):
        """Edge arrays -> deduplicated, row-sorted CSR (symmetrised like model.py:308-310)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        if symmetric:
            src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
        key = np.unique(src * np.int64(num_nodes) + dst)
        s = key // num_nodes
        d = (key - s * num_nodes).astype(np.int32)
        rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(s, minlength=num_nodes), out=rowptr[1:])
        return cls(rowptr, d, device)

    def to_adj_lists(self):
        col = self.col.cpu().numpy()
        rp = self.rowptr_host
        return {v: set(int(c) for c in col[rp[v]:rp[v + 1]]) for v in range(self.num_nodes)}

    def max_degree_of(self, nodes_host):
        nodes_host = np.asarray(nodes_host, dtype=np.int64)
        if nodes_host.size == 0:
            return 0
        return int((self.rowptr_host[nodes_host + 1] - self.rowptr_host[nodes_host]).max())


_GRAPH_CACHE = {}


def graph_of(adj_lists, device="cuda", num_nodes=None):
    """CSRGraph for an ``adj_lists`` mapping, converted once per object (the reference passes
    the same dict to every Encoder, model.py:219-221).  Mutating the mapping afterwards is
    not tracked -- call ``forget(adj_lists)`` first.  ``num_nodes`` (rows of the feature table): the CSR gets a row
    for every node of the table, also for isolated ones that appear in no edge (the reference's defaultdict(set)
    answers those with an empty set, model.py:303); without it rows stop at the largest id seen in an edge, and the
    kernels treat any id past that as isolated."""
    if isinstance(adj_lists, CSRGraph):
        return adj_lists
    hit = _GRAPH_CACHE.get(id(adj_lists))
    if hit is not None and hit[0] is adj_lists and (num_nodes is None or hit[1].num_nodes >= int(num_nodes)):
        return hit[1]
    g = CSRGraph.from_adj_lists(adj_lists, num_nodes=num_nodes, device=device)
    _GRAPH_CACHE[id(adj_lists)] = (adj_lists, g)
    return g


def forget(adj_lists=None):
    if adj_lists is None:
        _GRAPH_CACHE.clear()
    else:
        _GRAPH_CACHE.pop(id(adj_lists), None)
