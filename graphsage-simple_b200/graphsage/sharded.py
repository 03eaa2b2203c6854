"""Partitioned feature table and CSR for graphs that are spread over the GPUs of one box
(BASELINE config 5): node ``v`` is owned by rank ``v % world``.

The reference looks features and adjacency up in local Python containers --
``features(LongTensor(unique_nodes_list))`` (graphsage/aggregators.py:62-65) and
``adj_lists[int(node)]`` (graphsage/encoders.py:47).  Here both lookups become one request/reply
round trip over NCCL all-to-all (NVLink / NVSwitch):

    bucket ids by owner (gs_bucket_by_owner) -> all_to_all(counts) -> all_to_all(ids)
      -> the owner's local kernel (gs_gather_rows / gs_sample_csr)
      -> all_to_all(answers) -> un-permute (gs_gather_rows with the permutation as ids)

``ShardedFeatures`` is a drop-in for the ``features`` callable the reference passes to
MeanAggregator / Encoder; ``ShardedCSR`` is a drop-in for ``adj_lists`` / CSRGraph.  Sampling is
counter-based on the GLOBAL node id, so a partitioned run draws exactly the neighbours a
single-GPU run draws (bit-exact; tests/test_sharded_gloo.py, tests/multigpu_sharded_check.py).
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


def _round4(x):
    return (int(x) + 3) // 4 * 4


def peer_mapped(t, group=None):
    """Put tensor ``t`` (this rank's shard; sizes may differ between ranks) into symmetric memory and
    return ``(local, ptrs)``: ``local`` aliases the symmetric buffer with t's shape and values, ``ptrs`` is an
    int64 CUDA tensor [world] holding every rank's buffer as mapped into THIS process -- what the
    ``*_peer`` kernels take.  Collective: every rank of the group must call it.  Without a process
    group (world 1) the tensor itself is used."""
    on = dist.is_available() and dist.is_initialized()
    t = t.contiguous()
    if not on or dist.get_world_size(group) == 1:
        return t, torch.tensor([t.data_ptr()], dtype=torch.int64, device=t.device)
    import torch.distributed._symmetric_memory as symm
    group = group if group is not None else dist.group.WORLD
    n = torch.tensor([t.numel()], device=t.device, dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
    buf = symm.empty(max(int(n.item()), 1), dtype=t.dtype, device=t.device)
    buf[:t.numel()].copy_(t.reshape(-1))
    handle = symm.rendezvous(buf, group)
    ptrs = torch.tensor([int(p) for p in handle.buffer_ptrs], dtype=torch.int64, device=t.device)
    torch.cuda.synchronize(t.device)
    dist.barrier(group)                      # every shard is complete before anyone reads it remotely
    local = buf[:t.numel()].view(t.shape)
    local._gs_symm_handle = handle           # keep the mapping alive as long as the view
    return local, ptrs


class _Plan:
    __slots__ = ("perm", "send_splits", "recv_splits", "recv_ids", "n")

    def __init__(self, perm, send_splits, recv_splits, recv_ids, n):
        self.perm, self.send_splits, self.recv_splits, self.recv_ids, self.n = perm, send_splits, recv_splits, recv_ids, n


class OwnerExchange:
    """Request/reply plumbing between the ranks that ask about node ids and the ranks that own them.
    ``route`` ships each id to its owner; ``reply`` ships one row of answer per id back and restores
    the caller's order.  The two device primitives are methods so that the world-size-2 gloo test can
    run the same choreography on CPU tensors with the oracle's restatement of the kernels."""

    def __init__(self, rank=None, world=None, group=None):
        on = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = int(rank if rank is not None else (dist.get_rank(group) if on else 0))
        self.world = int(world if world is not None else (dist.get_world_size(group) if on else 1))
        self.bytes_sent = 0          # payload bytes this rank put on the wire (excludes the self bucket)

    # ---- device primitives -------------------------------------------------------------------
    def _bucket(self, ids, emit_local):
        return ops.bucket_by_owner(ids, self.world, emit_local=emit_local)

    def _take_rows(self, src, index):
        """out[i, :] = src[index[i], :] for a contiguous fp32 [m, ld] buffer (ld % 4 == 0)."""
        out = torch.empty((max(index.shape[0], 1), src.shape[1]), device=src.device, dtype=torch.float32)[:index.shape[0]]
        ops.gather_rows(src, src.shape[1], index, out)
        return out

    # ---- collectives ---------------------------------------------------------------------------
    def _a2a(self, out, inp, out_splits=None, in_splits=None):
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
        return out

    def route(self, ids, emit_local=False):
        """ids: int32 [n] global node ids on this rank's device.  Returns a plan whose ``recv_ids`` are the
        ids (global, or local row = id // world with ``emit_local``) this rank must answer, grouped by
        asking rank."""
        n = ids.shape[0]
        send, perm, counts = self._bucket(ids, emit_local)
        counts64 = counts.to(torch.int64)
        theirs = torch.empty_like(counts64)
        self._a2a(theirs, counts64)
        both = torch.stack([counts64, theirs]).cpu()          # the one host sync of a round trip
        send_splits, recv_splits = both[0].tolist(), both[1].tolist()
        recv = torch.empty(max(sum(recv_splits), 1), device=ids.device, dtype=torch.int32)[:sum(recv_splits)]
        self._a2a(recv, send, recv_splits, send_splits)
        self.bytes_sent += 4 * (n - send_splits[self.rank])
        return _Plan(perm, send_splits, recv_splits, recv, n)

    def reply(self, plan, payload):
        """payload: contiguous fp32 [len(plan.recv_ids), ld] (ld % 4 == 0), row i answering recv_ids[i].
        Returns fp32 [n, ld] in the order of the ids given to ``route``."""
        ld = payload.shape[1]
        back = torch.empty((max(plan.n, 1), ld), device=payload.device, dtype=torch.float32)[:plan.n]
        self._a2a(back, payload, plan.send_splits, plan.recv_splits)
        self.bytes_sent += 4 * ld * (payload.shape[0] - plan.recv_splits[self.rank])
        return self._take_rows(back, plan.perm)


class ShardedFeatures(nn.Module):
    """``features`` callable (aggregators.py:20: "function mapping LongTensor of node ids to
    FloatTensor of feature values") over a table whose row ``v`` lives on rank ``v % world`` as
    local row ``v // world``.  Frozen, like the reference's table (model.py:214-215)."""

    def __init__(self, local_rows, num_nodes, rank=None, world=None, group=None, exchange=None, peer=False):
        """``peer=True``: the shards are mapped into every rank (symmetric memory) and a lookup is ONE kernel
        that reads remote rows over NVLink (gs_gather_rows_peer) instead of the all-to-all round trip."""
        super().__init__()
        self.ex = exchange if exchange is not None else OwnerExchange(rank, world, group)
        self.peer = bool(peer)
        self.num_nodes = int(num_nodes)
        self.dim = int(local_rows.shape[1])
        self.ld = _round4(self.dim)
        t = local_rows if isinstance(local_rows, torch.Tensor) else torch.from_numpy(np.asarray(local_rows, np.float32))
        if t.shape[1] != self.ld or not t.is_contiguous():
            buf = torch.zeros((t.shape[0], self.ld), dtype=torch.float32, device=t.device)
            buf[:, :self.dim] = t
            t = buf
        self.table_ptrs = None
        if self.peer:
            t, self.table_ptrs = peer_mapped(t, self.ex.group)
        self.register_buffer("table", t, persistent=False)

    @staticmethod
    def shard_of(table, rank, world):
        """Rows owned by ``rank`` of a full [N, F] table, in local order."""
        return table[rank::world]

    def _local_rows(self, local_ids):
        out = torch.empty((max(local_ids.shape[0], 1), self.ld), device=self.table.device,
                          dtype=torch.float32)[:local_ids.shape[0]]
        ops.gather_rows(self.table, self.ld, local_ids, out)
        return out

    def forward(self, ids):
        ids32 = ops.as_ids(ids, self.table.device)
        if self.peer:
            n = ids32.shape[0]
            out = torch.empty((max(n, 1), self.ld), device=self.table.device, dtype=torch.float32)[:n]
            ops.gather_rows_peer(self.table_ptrs, self.ex.world, self.ld, self.ld, ids32, out)
            return out[:, :self.dim]
        plan = self.ex.route(ids32, emit_local=True)
        rows = self.ex.reply(plan, self._local_rows(plan.recv_ids))
        return rows[:, :self.dim]


class ShardedCSR:
    """Adjacency partitioned by owner: this rank keeps the neighbour lists of its own nodes only
    (``col``); ``rowptr`` stays full length (8 B per node, rows of other ranks are empty) so the sampler
    kernel is indexed by global id unchanged.  ``sample`` answers the same call as CSRGraph.sample."""

    def __init__(self, rowptr_local, col_local, num_nodes, max_degree, exchange, device="cuda", peer=False):
        """``peer=True``: every rank's (compact rowptr, col) is mapped into every rank and sampling is ONE
        kernel reading remote adjacency rows over NVLink (gs_sample_csr_peer)."""
        self.ex = exchange
        self.peer = bool(peer)
        self.num_nodes = int(num_nodes)
        self.max_degree = int(max_degree)
        self.device = torch.device(device)
        self.rowptr = torch.as_tensor(np.ascontiguousarray(rowptr_local, dtype=np.int64)).to(self.device)
        col = np.ascontiguousarray(col_local, dtype=np.int32)
        self.col = torch.as_tensor(col if col.size else np.zeros(1, np.int32)).to(self.device)
        self.num_entries = int(col.shape[0])
        self.rowptr_ptrs = self.col_ptrs = None
        if self.peer:
            rp = np.ascontiguousarray(rowptr_local, dtype=np.int64)
            w, r = self.ex.world, self.ex.rank
            compact = np.concatenate([rp[r:self.num_nodes:w], rp[-1:]])          # owned rows only: local row = v // world
            self._rp_local, self.rowptr_ptrs = peer_mapped(torch.from_numpy(compact).to(self.device), self.ex.group)
            self._col_local, self.col_ptrs = peer_mapped(self.col, self.ex.group)

    @classmethod
    def from_global(cls, rowptr, col, rank, world, device="cuda", exchange=None, peer=False):
        """Keep the rows v % world == rank of a full CSR (host arrays)."""
        rowptr = np.asarray(rowptr, dtype=np.int64)
        col = np.asarray(col)
        n = rowptr.shape[0] - 1
        deg = np.diff(rowptr)
        mine = (np.arange(n) % world) == rank
        deg_local = np.where(mine, deg, 0)
        rp = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(deg_local, out=rp[1:])
        keep = np.repeat(mine, deg)
        ex = exchange if exchange is not None else OwnerExchange(rank, world)
        return cls(rp, col[keep].astype(np.int32), n, int(deg.max()) if n else 0, ex, device, peer=peer)

    def _local_sample(self, ids, k, add_self, seed, step, tag, width, step_dev=None):
        return ops.sample_csr(self.rowptr, self.col, self.num_nodes, ids, k, add_self=add_self, seed=seed,
                              step=step, tag_head=tag, width=width, step_dev=step_dev)

    def sample(self, ids, k, add_self=False, seed=0, step=0, tag=0, width=None, step_dev=None):
        if width is None:
            width = (k if k is not None else self.max_degree) + (1 if add_self else 0)
        width = max(int(width), 1)
        if self.peer:
            return ops.sample_csr_peer(self.rowptr_ptrs, self.col_ptrs, self.ex.world, self.num_nodes, ids, k,
                                       add_self=add_self, seed=seed, step=step, tag_head=tag, width=width,
                                       step_dev=step_dev)
        plan = self.ex.route(ids, emit_local=False)
        idx, cnt = self._local_sample(plan.recv_ids, k, add_self, seed, step, tag, width, step_dev)
        ld = _round4(width + 1)
        tile = torch.zeros((max(idx.shape[0], 1), ld), device=idx.device, dtype=torch.int32)[:idx.shape[0]]
        tile[:, :width] = idx
        tile[:, width] = cnt
        back = self.ex.reply(plan, tile.view(torch.float32)).view(torch.int32)      # bit-exact 4-byte words
        return back[:, :width].contiguous(), back[:, width].contiguous()


def allreduce_grads(params, world, n_local, n_global, group=None):
    """Data-parallel gradient step for the op-by-op path: sums the gradients of ``params`` over the
    ranks so that the result is the gradient of the mean loss over the GLOBAL batch (model.py:57, 69
    uses mean reduction): grad = sum_r (n_r / N) grad_r."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params]) * (float(n_local) / float(n_global))
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        k = p.numel()
        p.grad.copy_(flat[off:off + k].view_as(p.grad))
        off += k
