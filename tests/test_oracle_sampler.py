"""Oracle self-checks for the sampler port: Philox known-answer vectors (Random123
kat_vectors) and the reference's sampling semantics (aggregators.py:42-48)."""
import numpy as np

from oracle import sampler_port as SP
from oracle.philox import philox4x32_10


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(x) for x in philox4x32_10(*ctr, *key)) == out


def _graph(rng, n, avg):
    deg = rng.poisson(avg, n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg]).astype(np.int32)
    return rowptr, col


def test_semantics_and_determinism():
    rng = np.random.default_rng(0)
    rowptr, col = _graph(rng, 400, 9)
    nodes = np.arange(400)
    idx, cnt = SP.sample_csr(rowptr, col, nodes, 6, seed=5, step=3, tags=1)
    idx_b, cnt_b = SP.sample_csr(rowptr, col, nodes, 6, seed=5, step=3, tags=1)
    assert np.array_equal(idx, idx_b) and np.array_equal(cnt, cnt_b)
    deg = np.diff(rowptr)
    assert np.array_equal(cnt, np.minimum(deg, 6))
    for v in nodes:
        row = idx[v, :cnt[v]]
        assert len(set(row.tolist())) == cnt[v]
        assert set(row.tolist()) <= set(col[rowptr[v]:rowptr[v + 1]].tolist())
        assert (np.diff(row) > 0).all()
    other, _ = SP.sample_csr(rowptr, col, nodes, 6, seed=5, step=4, tags=1)
    assert (other != idx).any()
    # take-all and self-loop union (set semantics)
    full, cf = SP.sample_csr(rowptr, col, nodes, -1, 5, 3, 1, add_self=True)
    for v in nodes:
        assert set(full[v, :cf[v]].tolist()) == set(col[rowptr[v]:rowptr[v + 1]].tolist()) | {int(v)}


def test_uniform_marginals():
    deg, k, steps = 30, 7, 3000
    rowptr = np.array([0, deg])
    col = np.arange(deg, dtype=np.int32)
    counts = np.zeros(deg)
    for s in range(steps):
        idx, _ = SP.sample_csr(rowptr, col, [0], k, seed=11, step=s, tags=2)
        counts[idx[0]] += 1
    expect = steps * k / deg
    chi2 = ((counts - expect) ** 2 / (expect * (1 - k / deg))).sum()
    assert chi2 < 65.0, chi2


def test_dedup_port():
    idx = np.array([[5, 9, -1], [9, 2, 7], [-1, -1, -1]], dtype=np.int32)
    cnt = np.array([2, 3, 0], dtype=np.int32)
    uniq, out = SP.dedup_remap(idx, cnt, slot_base=10)
    assert uniq.tolist() == [2, 5, 7, 9]
    assert out.tolist() == [[11, 13, -1], [13, 10, 12], [-1, -1, -1]]
