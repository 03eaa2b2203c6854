"""Drop-in for graphsage/aggregators.py of zjzijielu/graphsage-simple: same class, same
constructor and ``forward`` signature (aggregators.py:16, 34), CUDA-only implementation.

What changed underneath: ``random.sample`` over Python sets -> the CSR
counter-based sampler kernel; ``set.union`` + id->column dict -> the dedup kernel; the dense
row-normalised ``B x U`` mask and ``mask.mm(embed_matrix)`` -> the fused gather-mean kernel;
its autograd backward -> the scatter-add kernel."""
import numpy as np
import torch
import torch.nn as nn

from . import ops, sampling
from .functional import GatherMean, RaggedGatherMean, TableLookup
from .graph import CSRGraph

TABLE_INITIALIZERS = ("1hot", "node_degree")          # aggregators.py:30, 68
RAGGED_MIN_DEGREE = 64      # un-sampled neighbourhoods wider than this use ragged tiles instead of [n, max_degree]


_static = {"on": False}


class static_shapes:
    """While active, ``MeanAggregator.forward`` never reads a size back from the device: the distinct-id list keeps
    its upper-bound length (``min(n * width, num_nodes)``, padded with id 0, which no tile entry refers to), so every
    tensor of a train step has a shape known on the host and the step can be captured as one CUDA graph
    (model.GraphedStep).  The padded rows are looked up / encoded and ignored."""

    def __enter__(self):
        self.prev, _static["on"] = _static["on"], True

    def __exit__(self, *exc):
        _static["on"] = self.prev
        return False


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("graphsage (B200 build) needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class MeanAggregator(nn.Module):
    """Aggregates a node's embeddings using the mean of its (sampled) neighbours' embeddings.

    Accepts both constructor forms: the fork's
    ``(features, initializer="None", cuda=False, gcn=False, feature_dim=100, num_nodes=100)``
    (aggregators.py:16; the fork itself passes an int in the ``initializer`` slot,
    model.py:220) and upstream's ``(features, cuda=False, gcn=False)``.  The ``cuda`` flag is
    stored (``Encoder.__init__`` overwrites it, encoders.py:30) but placement is always the
    current CUDA device."""

    def __init__(self, features, *args, **kwargs):
        super().__init__()
        if args and isinstance(args[0], bool):
            names = ("cuda", "gcn")                                   # upstream form
        else:
            names = ("initializer", "cuda", "gcn", "feature_dim", "num_nodes")
        if len(args) > len(names):
            raise TypeError("MeanAggregator: too many positional arguments")
        opts = dict(initializer="None", cuda=False, gcn=False, feature_dim=100, num_nodes=100)
        opts.update(dict(zip(names, args)))
        for k, v in kwargs.items():
            if k not in opts:
                raise TypeError("MeanAggregator: unexpected keyword %r" % k)
            opts[k] = v
        if not isinstance(opts["initializer"], str):
            opts["initializer"] = "None"                              # stray int, model.py:220
        dev = _device()
        if isinstance(features, nn.Module):
            features.to(dev)                                          # "# features.cuda()", model.py:216
        self.features = features
        self.cuda = opts["cuda"]
        self.gcn = opts["gcn"]
        self.initializer = opts["initializer"]
        if self.initializer in TABLE_INITIALIZERS:
            self.embed = nn.Embedding(opts["num_nodes"], opts["feature_dim"]).to(dev)   # aggregators.py:31
        self.uid = sampling.next_uid()
        self._calls = (-1, 0)            # (step, calls made in that step) -> RNG tag
        self._scratch = None
        self._hot = None
        self._aligned = None

    # ---- sampler tag bookkeeping -----------------------------------------------------------
    def _next_tag(self):
        step = sampling.get_step()
        last, calls = self._calls
        calls = calls + 1 if last == step else 0
        self._calls = (step, calls)
        return sampling.call_tag(self.uid, calls)

    def _dedup_scratch(self, num_ids, dev):
        if self._scratch is None or self._scratch.num_nodes < num_ids:
            self._scratch = ops.DedupScratch(num_ids, dev)
        return self._scratch

    def forward(self, nodes, to_neighs=None, num_sample=10, initializer="None", graph=None):
        """
        nodes      -- list / ndarray / LongTensor of node ids in the batch
        to_neighs  -- list of sets (reference form, aggregators.py:36-37), or None together
                      with ``graph`` (a CSRGraph) to sample straight from the device CSR
        num_sample -- neighbours to sample per node; no sampling if None (aggregators.py:47-48)
        returns    -- FloatTensor [len(nodes), D] (CUDA), differentiable w.r.t. ``features``
        """
        dev = _device()
        if to_neighs is not None:
            if self.gcn:
                if num_sample is None:
                    to_neighs = [set(s) | {int(nodes[i])} for i, s in enumerate(to_neighs)]
                    idx, cnt, num_ids = self._tile_plain(nodes, to_neighs, None, dev)
                else:
                    idx, cnt, num_ids = self._tile_plain(nodes, to_neighs, num_sample, dev, add_self=True)
            else:
                idx, cnt, num_ids = self._tile_plain(nodes, to_neighs, num_sample, dev)
        else:
            if not hasattr(graph, "sample"):
                raise TypeError("MeanAggregator.forward needs to_neighs (list of sets) or graph=CSRGraph")
            ids = ops.as_ids(nodes, dev)
            if num_sample is None and isinstance(graph, CSRGraph) and graph.max_degree > RAGGED_MIN_DEGREE:
                return self._forward_ragged(ids, graph, initializer, dev)
            width = None
            if num_sample is None:
                width = graph.max_degree + (1 if self.gcn else 0)
            kw = {"step_dev": sampling.get_step_dev()} if sampling.get_step_dev() is not None else {}
            idx, cnt = graph.sample(ids, num_sample, add_self=self.gcn, seed=sampling.get_seed(),
                                    step=sampling.get_step(), tag=self._next_tag(), width=width, **kw)
            num_ids = graph.num_nodes
        # dedup (aggregators.py:52-53) and lookup of the distinct rows (aggregators.py:62-65)
        direct = self._direct_table(initializer)
        if direct is not None:
            # Innermost layer over a FROZEN table (model.py:214-215): nothing upstream needs a gradient, so the dedup +
            # lookup of the distinct rows (aggregators.py:52-65) -- which exist to avoid fetching a row twice through the
            # dense mask -- are skipped and the mean reads the sampled rows straight from the table (K2): same rows, same
            # summation order, same bits, without materialising ``embed_matrix``.
            n = idx.shape[0]
            if direct[0] == "peer":
                f = direct[1]
                out = ops.empty_rows(n, f.dim, dev)
                ops.gather_mean_fwd_peer(f.table_ptrs, f.ex.world, f.ld, f.dim, idx, cnt, out, neigh_off=0)
            else:
                t = direct[1]
                out = ops.empty_rows(n, t.shape[1], dev)
                ops.gather_mean_fwd(t, t.shape[1], idx, cnt, out, neigh_off=0)
            return out
        if _static["on"]:                # no host read of the distinct count: padded, fixed-size id list
            uniq = torch.zeros(max(min(idx.shape[0] * idx.shape[1], num_ids), 1), device=dev, dtype=torch.int32)
            ops.dedup_remap(idx, cnt, self._dedup_scratch(num_ids, dev), uniq=uniq)
            return GatherMean.apply(self._lookup(uniq, initializer), idx, cnt)
        uniq, n_total = ops.dedup_remap(idx, cnt, self._dedup_scratch(num_ids, dev))
        return GatherMean.apply(self._lookup(uniq[:int(n_total.item())], initializer), idx, cnt)

    def _direct_table(self, initializer):
        """("local", table [N, F]) / ("peer", ShardedFeatures) when ``features`` is a frozen table the gather-mean kernel
        can read by global node id, else None."""
        if initializer in TABLE_INITIALIZERS:
            return None
        f = self.features
        if isinstance(f, nn.Embedding) and not f.weight.requires_grad and f.weight.is_cuda:
            key = (f.weight.data_ptr(), f.weight._version, tuple(f.weight.shape))
            if self._aligned is None or self._aligned[0] != key:       # padded copy only if the rows are not 16-B aligned
                self._aligned = (key, ops.aligned_rows(f.weight.data))
            return ("local", self._aligned[1])
        if getattr(f, "table_ptrs", None) is not None:                 # sharded.ShardedFeatures(peer=True)
            return ("peer", f)
        return None

    def _lookup(self, unique_ids, initializer):
        """``embed_matrix`` of aggregators.py:62-71 for the distinct ids (int32 CUDA): rows of the frozen table, of
        the trainable table ``self.embed`` (1hot / node_degree), or whatever the ``features`` callable returns."""
        frozen = isinstance(self.features, nn.Embedding) and not self.features.weight.requires_grad
        if initializer in TABLE_INITIALIZERS:                          # aggregators.py:68-71
            if frozen:                                                 # position of the 1: once per table, not per row
                hot = ops.remap_ids(unique_ids.clone().view(-1, 1), None, self.hot_map()).view(-1)
            else:
                hot = self.features(unique_ids.long()).argmax(dim=1).to(torch.int32)
            return TableLookup.apply(self.embed.weight, hot)
        if isinstance(self.features, nn.Embedding):
            return TableLookup.apply(self.features.weight, unique_ids)
        return self.features(unique_ids.long())

    def _forward_ragged(self, ids, graph, initializer, dev):
        """num_sample=None (aggregators.py:47-48) on a graph with large degrees: the whole neighbourhoods as a
        ragged tile (offsets + flat ids) instead of a [n, max_degree] one; same dedup / lookup / mean."""
        off, flat = ops.take_all_csr(graph.rowptr, graph.col, ids, add_self=self.gcn)
        uniq, n_total = ops.dedup_remap(flat.view(-1, 1), None, self._dedup_scratch(graph.num_nodes, dev))
        return RaggedGatherMean.apply(self._lookup(uniq[:int(n_total.item())], initializer), off, flat)

    def _tile_plain(self, nodes, to_neighs, num_sample, dev, add_self=False):
        n = len(to_neighs)
        lens = np.fromiter((len(s) for s in to_neighs), dtype=np.int64, count=n)
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=rowptr[1:])
        total = int(rowptr[-1])
        col = np.zeros(max(total, 1), dtype=np.int32)
        for i, s in enumerate(to_neighs):
            col[rowptr[i]:rowptr[i + 1]] = sorted(int(x) for x in s)
        node_ids = np.asarray([int(v) for v in nodes], dtype=np.int32)
        hi = max(int(col[:total].max()) if total else 0, int(node_ids.max()) if n else 0)
        width = None
        if num_sample is None:
            width = int(lens.max()) if n else 1
        idx, cnt = ops.sample_csr(torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev), n,
                                  torch.arange(n, device=dev, dtype=torch.int32), num_sample,
                                  add_self=False, seed=sampling.get_seed(), step=sampling.get_step(),
                                  tag_head=self._next_tag(), width=(width if not add_self else num_sample + 1))
        if add_self:      # union with the node itself after sampling (set semantics: no double count)
            ids = torch.from_numpy(node_ids).to(dev)
            has = (idx == ids[:, None]).any(dim=1)
            rows = torch.nonzero(~has).flatten()
            idx[rows, cnt[rows].long()] = ids[rows]
            cnt[rows] += 1
        return idx, cnt, hi + 1

    def hot_map(self):
        """int32 [num_nodes]: position of the 1 in each row of the frozen one-hot feature table (aggregators.py:69) --
        the node id itself for ``1hot``, its degree for ``node_degree`` -- computed once per table, not per batch."""
        f = self.features
        key = (f.weight.data_ptr(), f.weight._version)
        if self._hot is None or self._hot[0] != key:
            self._hot = (key, f.weight.argmax(dim=1).to(torch.int32).contiguous())
        return self._hot[1]
