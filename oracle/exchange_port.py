"""CPU restatement of the owner-bucketing specification of the partitioned table / CSR
(TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference has no distributed code; its lookups are local (``features(LongTensor(ids))``,
graphsage/aggregators.py:62-65; ``adj_lists[int(node)]``, graphsage/encoders.py:47).  With rows
partitioned by owner = id % world the CUDA path buckets the requested ids by owner
(include/gsage.h: gs_bucket_by_owner): a STABLE counting sort, i.e.

    order      = stable argsort of (ids % world)
    send_ids   = ids[order]              (or ids[order] // world with emit_local)
    perm[i]    = position of ids[i] in send_ids
    counts[o]  = number of ids owned by rank o

``partitioned_lookup`` is the whole round trip written as plain indexing: whatever the exchange does,
the result must equal ``table[ids]``.
"""
import numpy as np


def bucket_by_owner(ids, world, emit_local=False):
    ids = np.asarray(ids, dtype=np.int32)
    owner = ids.astype(np.int64) % world
    order = np.argsort(owner, kind="stable")
    send = ids[order] // world if emit_local else ids[order]
    perm = np.empty(ids.shape[0], dtype=np.int32)
    perm[order] = np.arange(ids.shape[0], dtype=np.int32)
    counts = np.bincount(owner, minlength=world).astype(np.int32)
    return send.astype(np.int32), perm, counts


def shard_rows(table, rank, world):
    """Rows owned by ``rank`` (v % world == rank) in local order (local row = v // world)."""
    return np.asarray(table)[rank::world]


def partitioned_lookup(shards, ids, world):
    """table[ids] computed through the shards: row v is shards[v % world][v // world]."""
    ids = np.asarray(ids, dtype=np.int64)
    return np.stack([shards[int(v % world)][int(v // world)] for v in ids]) if ids.size else \
        np.zeros((0,) + shards[0].shape[1:], dtype=shards[0].dtype)
