"""CPU restatement of the CSR neighbour sampler / frontier dedup specification
(TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference samples with ``random.sample`` over Python sets
(graphsage/aggregators.py:42-46): for every node, *all* neighbours when
``len(neigh) < num_sample`` and a uniformly random ``num_sample``-subset otherwise (for
``len == num_sample`` the subset is the whole set), ``num_sample=None`` meaning "no
sampling" (:47-48).  The intended GCN self-loop union is at :50-51.  Deduplication of the
sampled ids into ``unique_nodes_list`` is at :52-53.

The CUDA sampler (include/gsage.h: gs_sample_csr) keeps those semantics but draws from a
counter-based generator so it can run on the device without state:

  for node v with sorted CSR row  col[rowptr[v] : rowptr[v+1]],  deg = row length
    if k < 0 (no sampling) or deg <= k:  take the whole row, cnt = deg
    else Floyd's k-subset algorithm over positions 0..deg-1:
         for m, j in enumerate(range(deg-k, deg)):
             r = word (m % 4) of Philox4x32-10(counter=(v, m//4, step_lo, tag), key=(seed_lo, seed_hi))
             t = (r * (j+1)) >> 32
             chosen.add(j if t in chosen else t)
         positions sorted ascending, cnt = k
    if add_self and v not in the tile: append v, cnt += 1       (aggregators.py:50-51, set semantics)
    unused tile slots are -1

Every function here is a plain-loop / numpy statement of exactly that.
"""
import numpy as np

from .philox import philox4x32_10


def sample_csr(rowptr, col, nodes, k, seed, step, tags, add_self=False, width=None):
    """Returns (idx[n, width] int32 filled with -1, cnt[n] int32).

    ``tags`` is a scalar or a per-row uint32 array (the call tag distinguishes the
    reference's three independent aggregator calls of one forward).
    ``k < 0`` means take-all (reference ``num_sample=None``).
    """
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int32)
    nodes = np.asarray(nodes, dtype=np.int64)
    n = nodes.shape[0]
    tags = np.broadcast_to(np.asarray(tags, dtype=np.uint32), (n,))
    deg = rowptr[nodes + 1] - rowptr[nodes]
    if width is None:
        width = (int(deg.max()) if n and k < 0 else max(k, 0)) + (1 if add_self else 0)
    idx = np.full((n, width), -1, dtype=np.int32)
    cnt = np.zeros(n, dtype=np.int32)
    seed = int(seed)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    step_lo = int(step) & 0xFFFFFFFF
    # Pre-draw the Philox words for every row that needs them (vectorised).
    need = (deg > k) & (k >= 0)
    nblk = (max(k, 0) + 3) // 4
    words = None
    if need.any() and nblk:
        c0 = nodes.astype(np.uint32)[:, None]
        c1 = np.arange(nblk, dtype=np.uint32)[None, :]
        w = philox4x32_10(c0, c1, np.uint32(step_lo), tags[:, None], k0, k1)
        words = np.stack(w, axis=-1).reshape(n, nblk * 4)
    for i in range(n):
        v = int(nodes[i])
        base, d = int(rowptr[v]), int(deg[i])
        if k < 0 or d <= k:
            row = col[base:base + d]
            c = d
        else:
            chosen = []
            for m, j in enumerate(range(d - k, d)):
                r = int(words[i, m])
                t = (r * (j + 1)) >> 32
                chosen.append(j if t in chosen else t)
            chosen.sort()
            row = col[base + np.asarray(chosen, dtype=np.int64)]
            c = k
        idx[i, :c] = row
        if add_self and v not in row:
            idx[i, c] = v
            c += 1
        cnt[i] = c
    return idx, cnt


def dedup_remap(idx, cnt, slot_base=0):
    """aggregators.py:52-56 restated for fixed-width tiles: the distinct sampled ids (sorted
    ascending -- the reference's order is Python set order, which only permutes mask
    columns) and the tile rewritten as positions into that list (+slot_base)."""
    idx = np.asarray(idx)
    n, width = idx.shape
    valid = np.arange(width)[None, :] < np.asarray(cnt)[:, None]
    uniq = np.unique(idx[valid]).astype(np.int32)
    out = np.full_like(idx, -1)
    out[valid] = (np.searchsorted(uniq, idx[valid]) + slot_base).astype(np.int32)
    return uniq, out
