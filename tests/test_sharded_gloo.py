"""World-size-2 (and 3) gloo tests (CPU) of the partitioned feature table / CSR host logic in
graphsage/sharded.py.  The request/reply choreography -- bucket by owner,
all_to_all of counts / ids / answers, un-permute -- runs exactly as on the GPUs; only the two device
primitives (gs_bucket_by_owner, gs_gather_rows) and the owner's local kernels are replaced by the
oracle's CPU restatements.  Pins:
  * ShardedFeatures(ids) == table[ids]                       (aggregators.py:62-65 semantics)
  * ShardedCSR.sample(ids) == the single-process sampler over the whole CSR, bit-exact
  * data-parallel gradient sync of the op-by-op path == gradient of the global-batch mean loss
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import exchange_port as XP
from oracle import sampler_port as SP


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(rng, n, avg):
    deg = rng.poisson(avg, n).astype(np.int64)
    deg[::17] = 0
    deg = np.minimum(deg, n - 1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg] + [np.zeros(0, np.int64)]).astype(np.int32)
    return rowptr, col


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "graphsage-simple_b200"))
    sys.path.insert(0, root)
    from graphsage import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)

    class CpuExchange(sharded.OwnerExchange):          # device primitives -> oracle restatements
        def _bucket(self, ids, emit_local):
            s, p, c = XP.bucket_by_owner(ids.numpy(), self.world, emit_local)
            return torch.from_numpy(s), torch.from_numpy(p), torch.from_numpy(c)

        def _take_rows(self, src, index):
            return src[index.long()]

    class CpuFeatures(sharded.ShardedFeatures):
        def _local_rows(self, local_ids):
            return self.table[local_ids.long()]

    class CpuCSR(sharded.ShardedCSR):
        def _local_sample(self, ids, k, add_self, seed, step, tag, width, step_dev=None):
            i, c = SP.sample_csr(self.rowptr.numpy(), self.col.numpy(), ids.numpy(), -1 if k is None else k, seed, step,
                                 tag, add_self=add_self, width=width)
            return torch.from_numpy(i), torch.from_numpy(c)

    rng = np.random.default_rng(4)                     # same graph / table on every rank
    n, f = 997, 10
    table = rng.standard_normal((n, f)).astype(np.float32)
    rowptr, col = _graph(rng, n, 9)
    ex = CpuExchange(rank, world)
    feats = CpuFeatures(torch.from_numpy(XP.shard_rows(table, rank, world).copy()), n, exchange=ex)
    graph = CpuCSR.from_global(rowptr, col, rank, world, device="cpu", exchange=ex)
    my = np.random.default_rng(100 + rank)
    res = {}
    for trial, m in enumerate((257 + 31 * rank, 1, 0, 64)):        # ragged, single, EMPTY request, dup-heavy
        ids = my.integers(0, n, m).astype(np.int32) if trial != 3 else my.integers(0, 5, m).astype(np.int32)
        rows = feats(torch.from_numpy(ids))
        res["rows%d" % trial] = rows.numpy().copy()
        res["want_rows%d" % trial] = table[ids.astype(np.int64)].reshape(-1, f)
        for k, add_self in ((5, False), (None, True)):
            idx, cnt = graph.sample(torch.from_numpy(ids), k, add_self=add_self, seed=77, step=3 + trial, tag=9)
            width = idx.shape[1]
            ridx, rcnt = SP.sample_csr(rowptr, col, ids, -1 if k is None else k, 77, 3 + trial, 9, add_self=add_self,
                                       width=width)
            key = "%d_%s" % (trial, k)
            res["idx" + key], res["cnt" + key] = idx.numpy().copy(), cnt.numpy().copy()
            res["want_idx" + key], res["want_cnt" + key] = ridx, rcnt
    # gradient sync: rank r holds n_r targets with a per-rank "mean loss" gradient g_r -> global mean gradient
    n_local = 3 + 2 * rank
    n_global = sum(3 + 2 * r for r in range(world))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.full((4,), float(rank + 1))
    sharded.allreduce_grads([p], world, n_local, n_global)
    res["grad"] = p.grad.numpy().copy()
    res["want_grad"] = np.full(4, sum((3 + 2 * r) * (r + 1) for r in range(world)) / n_global, dtype=np.float32)
    res["sent"] = np.array([ex.bytes_sent])
    np.savez(out % rank, **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_lookups_equal_local_lookups(tmp_path, world):
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        r = np.load(out % rank)
        keys = [k for k in r.files if k.startswith("want_")]
        assert len(keys) >= 13
        for k in keys:
            got, want = r[k[5:]], r[k]
            assert got.shape == want.shape, (k, got.shape, want.shape)
            if got.dtype.kind == "f":
                assert np.array_equal(got, want) or np.allclose(got, want, rtol=1e-6), k
            else:
                assert np.array_equal(got, want), k
        assert r["sent"][0] > 0


def test_oracle_bucket_is_a_stable_partition():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 1000, 500).astype(np.int32)
    send, perm, counts = XP.bucket_by_owner(ids, 8, emit_local=True)
    assert counts.sum() == 500 and np.array_equal(send[perm], ids // 8)
    off = np.concatenate([[0], np.cumsum(counts)])
    full, _, _ = XP.bucket_by_owner(ids, 8)
    for o in range(8):
        seg = full[off[o]:off[o + 1]]
        assert np.all(seg % 8 == o)
        assert np.array_equal(seg, ids[ids % 8 == o])          # stable: original order inside a bucket
    shards = [XP.shard_rows(np.arange(1000)[:, None], r, 8) for r in range(8)]
    assert np.array_equal(XP.partitioned_lookup(shards, ids, 8)[:, 0], ids)
