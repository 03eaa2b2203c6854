"""Kernel-level parity (through the C ABI) against the CPU oracle on seeded inputs.
Integer/index work is bit-exact; fp32 work is checked norm-wise at 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import sampler_port as SP

pytestmark = pytest.mark.gpu

REL = 1e-5          # north_star: fp32 outputs within 1e-5 relative (norm-wise)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def random_csr(rng, n, avg_deg, hub=0):
    deg = rng.poisson(avg_deg, n).astype(np.int64)
    deg[rng.integers(0, n, max(n // 50, 1))] = 0          # some isolated nodes
    if hub:
        deg[rng.integers(0, n, 3)] = hub
    deg = np.minimum(deg, n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(n, size=d, replace=False)) for d in deg] + [np.zeros(0, np.int64)])
    return rowptr, col.astype(np.int32)


@pytest.fixture(scope="module")
def ops():
    from graphsage import ops as o
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("k,add_self", [(5, False), (10, False), (25, True), (1, False), (None, False), (None, True),
                                        (32, False), (40, True), (64, False)])
def test_sampler_bit_exact(ops, k, add_self):
    rng = np.random.default_rng(3)
    n = 3000
    rowptr, col = random_csr(rng, n, 12, hub=200)
    nodes = rng.integers(0, n, 1777).astype(np.int32)
    tags = np.where(np.arange(nodes.size) < 700, 7, 9).astype(np.uint32)
    width = None if k is not None else int(np.diff(rowptr).max()) + (1 if add_self else 0)
    idx, cnt = ops.sample_csr(dev(rowptr), dev(col), n, dev(nodes), k, add_self=add_self, seed=0x1234567890ABCDEF,
                              step=41, tag_head=7, tag_tail=9, n_head=700, width=width)
    ridx, rcnt = SP.sample_csr(rowptr, col, nodes, -1 if k is None else k, 0x1234567890ABCDEF, 41, tags,
                               add_self=add_self, width=width)
    assert np.array_equal(cnt.cpu().numpy(), rcnt)
    assert np.array_equal(idx.cpu().numpy(), ridx)


def test_sampler_semantics_match_reference(ops):
    """aggregators.py:42-48: all neighbours when deg < k (== k gives the same set), exactly k
    distinct members of the adjacency otherwise; different steps/tags give different draws."""
    rng = np.random.default_rng(5)
    n, k = 2000, 10
    rowptr, col = random_csr(rng, n, 14)
    nodes = np.arange(n, dtype=np.int32)
    idx, cnt = ops.sample_csr(dev(rowptr), dev(col), n, dev(nodes), k, seed=9, step=1, tag_head=1)
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    deg = np.diff(rowptr)
    assert np.array_equal(cnt, np.minimum(deg, k))
    for v in range(n):
        row = idx[v, :cnt[v]]
        adj = col[rowptr[v]:rowptr[v + 1]]
        assert len(set(row.tolist())) == cnt[v] and np.isin(row, adj).all()
        if deg[v] <= k:
            assert np.array_equal(row, adj)
        assert (idx[v, cnt[v]:] == -1).all()
    idx2, _ = ops.sample_csr(dev(rowptr), dev(col), n, dev(nodes), k, seed=9, step=2, tag_head=1)
    assert (idx2.cpu().numpy() != idx).any()


def test_sampler_uniform_marginals(ops):
    """chi-square on the inclusion counts of one 40-neighbour row over 4000 steps."""
    deg, k, steps = 40, 10, 4000
    rowptr = np.array([0, deg], dtype=np.int64)
    col = np.arange(100, 100 + deg, dtype=np.int32)
    counts = np.zeros(deg)
    nodes = dev(np.zeros(1, dtype=np.int32))
    rp, cl = dev(rowptr), dev(col)
    for s in range(steps):
        idx, _ = ops.sample_csr(rp, cl, 1, nodes, k, seed=77, step=s, tag_head=3)
        counts[idx.cpu().numpy()[0] - 100] += 1
    expect = steps * k / deg
    chi2 = ((counts - expect) ** 2 / (expect * (1 - k / deg))).sum()
    assert chi2 < 80.0, chi2            # 39 dof: P(chi2 > 80) ~ 1e-4


@pytest.mark.parametrize("n_rows,width,num_nodes", [(1000, 25, 5000), (37, 3, 50), (4096, 10, 233000), (1, 1, 3),
                                                    (16384, 11, 2400000), (300, 7, 65537), (5, 2, 31)])
def test_dedup_bit_exact(ops, n_rows, width, num_nodes):
    rng = np.random.default_rng(8)
    scratch = ops.DedupScratch(num_nodes, "cuda")
    for rep in range(3):                                   # the scratch (node bitmap) is left clean by every call:
        cnt = rng.integers(0, width + 1, n_rows).astype(np.int32)      # a DIFFERENT tile each time must not see stale bits
        idx = rng.integers(0, num_nodes, (n_rows, width)).astype(np.int32)
        if rep == 1:
            idx[0, :] = num_nodes - 1                      # last id of the last bitmap word
            idx[-1, :] = 0
        idx[np.arange(width)[None, :] >= cnt[:, None]] = -1
        d_idx = dev(idx)
        uniq, total = ops.dedup_remap(d_idx, dev(cnt), scratch, slot_base=17)
        ru, ridx = SP.dedup_remap(idx, cnt, slot_base=17)
        u = int(total.item()) - 17
        assert u == ru.size
        assert np.array_equal(uniq.cpu().numpy()[:u], ru)
        assert np.array_equal(d_idx.cpu().numpy(), ridx)
        assert int(scratch.slot_of[:(num_nodes + 31) // 32].abs().sum().item()) == 0      # bitmap zero again


def test_dedup_device_row_count(ops):
    rng = np.random.default_rng(9)
    idx = rng.integers(0, 100, (64, 4)).astype(np.int32)
    cnt = np.full(64, 4, dtype=np.int32)
    d_idx = dev(idx)
    n_dev = torch.tensor([20], dtype=torch.int32, device="cuda")
    uniq, total = ops.dedup_remap(d_idx, dev(cnt), ops.DedupScratch(100, "cuda"), n_dev=n_dev)
    ru, ridx = SP.dedup_remap(idx[:20], cnt[:20])
    assert int(total.item()) == ru.size
    assert np.array_equal(d_idx.cpu().numpy()[:20], ridx)
    assert np.array_equal(d_idx.cpu().numpy()[20:], idx[20:])      # untouched beyond *n_dev


def make_tile(rng, n, width, rows, empty_rows=True):
    cnt = rng.integers(1, width + 1, n).astype(np.int32)
    if empty_rows:
        cnt[rng.integers(0, n, max(n // 20, 1))] = 0
    idx = rng.integers(0, rows, (n, width)).astype(np.int32)
    idx[np.arange(width)[None, :] >= cnt[:, None]] = -1
    return idx, cnt


def ref_gather_mean(table, idx, cnt):
    out = np.zeros((idx.shape[0], table.shape[1]), dtype=np.float64)
    for i in range(idx.shape[0]):
        if cnt[i]:
            out[i] = table[idx[i, :cnt[i]]].astype(np.float64).sum(0) / cnt[i]
    return out


@pytest.mark.parametrize("dim,width,with_self", [(602, 10, True), (128, 25, True), (50, 5, False), (1433, 10, False),
                                                 (3703, 5, True), (7, 40, False), (100, 15, True)])
def test_gather_mean_forward(ops, dim, width, with_self):
    rng = np.random.default_rng(dim)
    rows, n = 5000, 777
    table = rng.standard_normal((rows, dim)).astype(np.float32)
    idx, cnt = make_tile(rng, n, width, rows)
    self_ids = rng.integers(0, rows, n).astype(np.int32)
    t = ops.aligned_rows(dev(table))
    off = dim if with_self else 0
    out = ops.empty_rows(n, off + dim, "cuda", zero=True)
    ops.gather_mean_fwd(t, dim, dev(idx), dev(cnt), out, neigh_off=off, self_ids=dev(self_ids) if with_self else None)
    got = out.cpu().numpy()
    ref = ref_gather_mean(table, idx, cnt)
    assert relerr(got[:, off:], ref) < REL
    if with_self:
        assert np.array_equal(got[:, :dim], table[self_ids])       # gathered rows: bit-exact copy


def test_gather_rows_bit_exact(ops):
    rng = np.random.default_rng(12)
    table = rng.standard_normal((1000, 602)).astype(np.float32)
    ids = rng.integers(0, 1000, 333).astype(np.int32)
    out = ops.empty_rows(333, 602, "cuda")
    ops.gather_rows(ops.aligned_rows(dev(table)), 602, dev(ids), out)
    assert np.array_equal(out.cpu().numpy(), table[ids])


@pytest.mark.parametrize("dim,width,with_self,off", [(128, 25, True, 128), (50, 5, False, 0), (602, 10, True, 602), (30, 7, True, 30)])
def test_scatter_mean_backward(ops, dim, width, with_self, off):
    rng = np.random.default_rng(dim + 1)
    rows, n = 900, 1500               # many duplicates -> contention on the atomics
    idx, cnt = make_tile(rng, n, width, rows)
    self_ids = rng.integers(0, rows, n).astype(np.int32)
    gout = rng.standard_normal((n, off + dim)).astype(np.float32)
    gtable = ops.empty_rows(rows, dim, "cuda", zero=True)
    ops.scatter_mean_bwd(ops.aligned_rows(dev(gout)), dim, dev(idx), dev(cnt), gtable, neigh_off=off,
                         self_ids=dev(self_ids) if with_self else None)
    ref = np.zeros((rows, dim), dtype=np.float64)
    for i in range(n):
        if cnt[i]:
            np.add.at(ref, idx[i, :cnt[i]], gout[i, off:off + dim].astype(np.float64) / cnt[i])
        if with_self:
            ref[self_ids[i]] += gout[i, :dim]
    assert relerr(gtable.cpu().numpy(), ref) < REL


@pytest.mark.parametrize("n,k_in,d_out,act", [(1000, 1204, 128, 1), (333, 256, 128, 1), (77, 1433, 50, 2),
                                              (5000, 100, 128, 1), (64, 36, 12, 0), (1, 8, 4, 1), (2049, 602, 128, 2)])
def test_encoder_forward_backward(ops, n, k_in, d_out, act):
    rng = np.random.default_rng(n + k_in)
    x = rng.standard_normal((n, k_in)).astype(np.float32)
    w = (rng.standard_normal((d_out, k_in)) / np.sqrt(k_in)).astype(np.float32)
    gh = rng.standard_normal((n, d_out)).astype(np.float32)
    xd, wd, ghd = ops.aligned_rows(dev(x)), ops.aligned_rows(dev(w)), ops.aligned_rows(dev(gh))
    h = ops.empty_rows(n, d_out, "cuda")
    ops.encoder_fwd(xd, wd, act, h)
    tx = torch.from_numpy(x).double().requires_grad_(True)
    tw = torch.from_numpy(w).double().requires_grad_(True)
    pre = tx @ tw.t()
    th = {0: pre, 1: torch.relu(pre), 2: torch.sigmoid(pre)}[act]
    assert relerr(h.cpu().numpy(), th.detach().numpy()) < REL
    th.backward(torch.from_numpy(gh).double())
    gw = ops.empty_rows(d_out, k_in, "cuda")
    gx = ops.empty_rows(n, k_in, "cuda")
    ops.encoder_bwd(xd, wd, h, ghd, act, gw, gx)
    assert relerr(gw.cpu().numpy(), tw.grad.numpy()) < REL
    assert relerr(gx.cpu().numpy(), tx.grad.numpy()) < REL


def test_encoder_device_row_count(ops):
    rng = np.random.default_rng(4)
    n_max, n, k_in, d_out = 600, 417, 64, 32
    x = rng.standard_normal((n_max, k_in)).astype(np.float32)
    w = rng.standard_normal((d_out, k_in)).astype(np.float32)
    gh = rng.standard_normal((n_max, d_out)).astype(np.float32)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    xd, wd = dev(x), dev(w)
    h = torch.full((n_max, d_out), 7.0, device="cuda")
    ops.encoder_fwd(xd, wd, 1, h, n_dev=n_dev)
    ref = np.maximum(x[:n].astype(np.float64) @ w.T.astype(np.float64), 0)
    assert relerr(h.cpu().numpy()[:n], ref) < REL
    assert (h.cpu().numpy()[n:] == 7.0).all()
    gw = ops.empty_rows(d_out, k_in, "cuda")
    ops.encoder_bwd(xd, wd, h, dev(gh), 1, gw, None, n_dev=n_dev)
    dz = gh[:n].astype(np.float64) * (ref > 0)
    assert relerr(gw.cpu().numpy(), dz.T @ x[:n].astype(np.float64)) < REL


@pytest.mark.parametrize("n,d,c", [(1024, 128, 41), (48, 12, 6), (5, 128, 3), (300, 128, 47), (100, 64, 100)])
def test_classifier_xent(ops, n, d, c):
    rng = np.random.default_rng(n + c)
    h = rng.standard_normal((n, d)).astype(np.float32)
    wc = (rng.standard_normal((c, d)) * 0.3).astype(np.float32)
    y = rng.integers(0, c, n).astype(np.int64)
    hd, wd = ops.aligned_rows(dev(h)), ops.aligned_rows(dev(wc))
    logits = ops.empty_rows(n, c, "cuda")
    loss = torch.zeros(1, device="cuda")
    gh = ops.empty_rows(n, d, "cuda")
    gwc = ops.empty_rows(c, d, "cuda")
    ops.classifier_xent(hd, wd, dev(y), 1.0, logits, loss, gh, gwc)
    th = torch.from_numpy(h).double().requires_grad_(True)
    tw = torch.from_numpy(wc).double().requires_grad_(True)
    tl = torch.nn.functional.cross_entropy(th @ tw.t(), torch.from_numpy(y))
    tl.backward()
    assert relerr(logits.cpu().numpy(), (th @ tw.t()).detach().numpy()) < REL
    assert abs(float(loss.item()) - float(tl)) / abs(float(tl)) < REL
    assert relerr(gh.cpu().numpy(), th.grad.numpy()) < REL
    assert relerr(gwc.cpu().numpy(), tw.grad.numpy()) < REL


def test_sgd_step_bit_exact(ops):
    rng = np.random.default_rng(1)
    p = rng.standard_normal(100003).astype(np.float32)
    g = rng.standard_normal(100003).astype(np.float32)
    pd = dev(p)
    ops.sgd_step(pd, dev(g), 0.7)
    ref = torch.from_numpy(p).add_(torch.from_numpy(g), alpha=-0.7).numpy()      # model.py:237, 250
    assert np.array_equal(pd.cpu().numpy(), ref)


def test_cpu_tensors_are_rejected(ops):
    with pytest.raises(RuntimeError):
        ops.gather_rows(torch.zeros(4, 4), 4, torch.zeros(2, dtype=torch.int32), torch.zeros(2, 4))


@pytest.mark.parametrize("n,k_in,act", [(1000, 1204, 1), (128, 32, 0), (333, 256, 1), (26000, 1204, 1), (2049, 600, 2),
                                        (77, 1000, 1)])
def test_encoder_tensor_core_path(ops, n, k_in, act):
    """tcgen05 3xTF32 encoder GEMMs against an fp64 reference, same 1e-5 bar as the fp32 path."""
    d_out = 128
    assert ops.encoder_tc_supported(k_in, d_out)
    g = torch.Generator(device="cuda").manual_seed(n + k_in)
    x = ops.empty_rows(n, k_in, "cuda")
    x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
    w = torch.randn(d_out, k_in, device="cuda", generator=g) / k_in ** 0.5
    gh = torch.randn(n, d_out, device="cuda", generator=g)
    h = torch.full((n, d_out), float("nan"), device="cuda")
    ops.encoder_fwd_tc(x, w, act, h)
    pre = x.double() @ w.double().t()
    ref = {0: pre, 1: torch.relu(pre), 2: torch.sigmoid(pre)}[act]
    assert relerr(h.cpu().numpy(), ref.cpu().numpy()) < REL
    gw = torch.full((d_out, k_in), float("nan"), device="cuda")
    ops.encoder_wgrad_tc(x, h, gh, act, gw)
    hd = h.double()
    dz = gh.double() * {0: torch.ones_like(hd), 1: (hd > 0).double(), 2: hd * (1 - hd)}[act]
    assert relerr(gw.cpu().numpy(), (dz.t() @ x.double()).cpu().numpy()) < REL


@pytest.mark.parametrize("n,feat,act,use_n_dev", [(3000, 602, 1, False), (700, 50, 2, True), (129, 128, 1, False),
                                                  (26000, 602, 1, True), (40, 500, 1, False)])
def test_sage_encoder_tensor_core_gathers_self_rows_in_place(ops, n, feat, act, use_n_dev):
    """gs_sage_encoder_fwd_tc / _wgrad_tc: X = [table[self_ids] | mean] with the self half gathered from the feature
    table inside the GEMM (never materialised), against fp64 and against the plain tcgen05 path fed the concatenated
    tile (encoders.py:53-61 and its MmBackward)."""
    g = torch.Generator(device="cuda").manual_seed(n + feat)
    num_nodes, d = 5000, 128
    table = ops.empty_rows(num_nodes, feat, "cuda", zero=True)
    table.copy_(torch.randn(num_nodes, feat, device="cuda", generator=g))
    n_alloc = n + 300 if use_n_dev else n
    ids = torch.randint(0, num_nodes, (n_alloc,), device="cuda", generator=g, dtype=torch.int32)
    ids[:3] = torch.tensor([0, num_nodes - 1, 0], dtype=torch.int32)             # first / last table row, a repeat
    mean = ops.empty_rows(n_alloc, feat, "cuda", zero=True)
    mean.copy_(torch.randn(n_alloc, feat, device="cuda", generator=g))
    w = torch.randn(d, 2 * feat, device="cuda", generator=g) / (2 * feat) ** 0.5
    gh = torch.randn(n_alloc, d, device="cuda", generator=g)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda") if use_n_dev else None
    h = torch.full((n_alloc, d), 7.0, device="cuda")
    ops.sage_encoder_fwd_tc(table, ids, feat, mean, w, act, h, n_dev=n_dev)
    x = torch.cat([table[ids[:n].long()], mean[:n]], dim=1).double()
    z = x @ w.double().t()
    ref = torch.relu(z) if act == 1 else torch.sigmoid(z)
    assert relerr(h[:n].cpu().numpy(), ref.cpu().numpy()) < REL
    if use_n_dev:
        assert (h[(n + 127) // 128 * 128:] == 7.0).all()
    gw = torch.full((d, 2 * feat), float("nan"), device="cuda")
    ops.sage_encoder_wgrad_tc(table, ids, feat, mean, h, gh, act, gw, n_dev=n_dev)
    hd = h[:n].double()
    dz = gh[:n].double() * ((hd > 0).double() if act == 1 else hd * (1 - hd))
    assert relerr(gw.cpu().numpy(), (dz.t() @ x).cpu().numpy()) < REL
    # same numbers as the plain kernels on the materialised tile, up to the summation order of the K chunks
    comb = ops.empty_rows(n, 2 * feat, "cuda", zero=True)
    comb.copy_(x.float())
    h2 = torch.empty((n, d), device="cuda")
    ops.encoder_fwd_tc(comb, w, act, h2)
    assert relerr(h2.cpu().numpy(), h[:n].cpu().numpy()) < REL


def test_encoder_tensor_core_device_row_count(ops):
    n_max, n, k_in, d_out = 900, 517, 256, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n_max, k_in, device="cuda", generator=g)
    w = torch.randn(d_out, k_in, device="cuda", generator=g) / 16
    gh = torch.randn(n_max, d_out, device="cuda", generator=g)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    h = torch.full((n_max, d_out), 7.0, device="cuda")
    ops.encoder_fwd_tc(x, w, 1, h, n_dev=n_dev)
    ref = torch.relu(x[:n].double() @ w.double().t())
    assert relerr(h[:n].cpu().numpy(), ref.cpu().numpy()) < REL
    assert (h[(n + 127) // 128 * 128:] == 7.0).all()          # tiles past *n_dev are not touched
    gw = torch.empty(d_out, k_in, device="cuda")
    ops.encoder_wgrad_tc(x, h, gh, 1, gw, n_dev=n_dev)
    dz = gh[:n].double() * (ref > 0)
    assert relerr(gw.cpu().numpy(), (dz.t() @ x[:n].double()).cpu().numpy()) < REL


@pytest.mark.parametrize("n,sage,c,act,width", [(1024, True, 41, 1, 25), (37, True, 7, 2, 10), (300, False, 128, 1, 5),
                                                (8, False, 3, 1, 1), (513, True, 47, 1, 26)])
def test_fused_head_matches_fp64_reference(ops, n, sage, c, act, width):
    """gs_head_fwd_bwd (outer layer + classifier + loss + backward in two launches) against an fp64
    torch restatement of aggregators.py:54-74, encoders.py:49-61, model.py:57-69 and autograd."""
    rng = np.random.default_rng(11)
    d1 = d2 = 128
    k2 = 2 * d1 if sage else d1
    m = n + 3 * n + 5                                   # rows of h1: [targets | hop-1 slots]
    h1 = rng.standard_normal((m, d1)).astype(np.float32)
    cnt = rng.integers(0, width + 1, n).astype(np.int32)
    cnt[0] = width
    idx = np.full((n, width), -1, dtype=np.int32)
    for i in range(n):                                  # first half of the targets share hub slot n
        if cnt[i] and i < n // 2:
            idx[i, :cnt[i]] = np.concatenate([[n], np.sort(rng.choice(np.arange(n + 1, m), cnt[i] - 1, replace=False))])
        else:
            idx[i, :cnt[i]] = np.sort(rng.choice(np.arange(n, m), cnt[i], replace=False))
    w2 = (rng.standard_normal((d2, k2)) / np.sqrt(k2)).astype(np.float32)
    wc = (rng.standard_normal((c, d2)) / np.sqrt(d2)).astype(np.float32)
    labels = rng.integers(0, c, n).astype(np.int64)
    scale = 0.75
    assert ops.head_supported(d1, k2, d2, c)
    dv = lambda a: dev(a)
    comb2 = torch.zeros((n, k2), device="cuda"); h2 = torch.zeros((n, d2), device="cuda")
    logits = ops.empty_rows(n, c, "cuda", zero=True); loss = torch.zeros(1, device="cuda")
    gh1 = torch.zeros((m, d1), device="cuda"); gw2 = torch.zeros((d2, k2), device="cuda")
    gwc = ops.empty_rows(c, d2, "cuda", zero=True)
    ws = ops.head_ws(n, k2, c, "cuda")
    self_slots = dv(np.arange(n, dtype=np.int32)) if sage else None
    for rep in range(2):                                # second call: re-armed tickets, gh1 re-zeroed
        gh1.zero_()
        ops.head_fwd_bwd(dv(h1), d1, dv(idx), dv(cnt), self_slots, dv(w2), act, dv(wc), dv(labels), scale,
                         comb2, h2, logits, loss, gh1, gw2, gwc, ws)
    # fp64 reference
    H1 = torch.tensor(h1, dtype=torch.float64, requires_grad=True)
    W2 = torch.tensor(w2, dtype=torch.float64, requires_grad=True)
    WC = torch.tensor(wc, dtype=torch.float64, requires_grad=True)
    mask = torch.zeros((n, m), dtype=torch.float64)
    for i in range(n):
        mask[i, idx[i, :cnt[i]]] = 1.0
    mean = (mask / mask.sum(1, keepdim=True).clamp(min=1)).mm(H1)
    comb = torch.cat([H1[:n], mean], 1) if sage else mean
    z = comb.mm(W2.t())
    hh = torch.relu(z) if act == 1 else torch.sigmoid(z)
    sc = hh.mm(WC.t())
    ref = torch.nn.functional.cross_entropy(sc, torch.tensor(labels)) * scale
    ref.backward()
    assert abs(float(loss.item()) * scale - float(ref)) / abs(float(ref)) < REL
    assert relerr(comb2.cpu().numpy(), comb.detach().numpy()) < REL
    assert relerr(h2.cpu().numpy(), hh.detach().numpy()) < REL
    assert relerr(logits.cpu().numpy(), sc.detach().numpy()) < REL
    assert relerr(gwc.cpu().numpy(), WC.grad.numpy()) < REL
    assert relerr(gw2.cpu().numpy(), W2.grad.numpy()) < REL
    assert relerr(gh1.cpu().numpy(), H1.grad.numpy()) < REL


@pytest.mark.parametrize("n_max,n_small,c", [(1024, 414, 3), (128, 100, 7), (512, 119, 41)])
def test_fused_head_workspace_serves_smaller_batches(ops, n_max, n_small, c):
    """ONE head workspace sized for the largest batch (what engine.py allocates) is used for B, then b < B
    (the last partial batch of an epoch), then B again: the weight gradients of every call must equal those of a
    call with a fresh exactly-sized workspace (round-1 bug: the re-arming tickets sat at an n-dependent offset)."""
    rng = np.random.default_rng(5)
    d1 = d2 = 128
    k2, width = 2 * d1, 10
    m = 4 * n_max
    h1 = dev(rng.standard_normal((m, d1)).astype(np.float32))
    w2 = dev((rng.standard_normal((d2, k2)) / np.sqrt(k2)).astype(np.float32))
    wc = dev((rng.standard_normal((c, d2)) / np.sqrt(d2)).astype(np.float32))
    shared = ops.head_ws(n_max, k2, c, "cuda")

    def run(n, ws, seed):
        r = np.random.default_rng(seed)
        cnt = r.integers(1, width + 1, n).astype(np.int32)
        idx = np.full((n, width), -1, dtype=np.int32)
        for i in range(n):
            idx[i, :cnt[i]] = np.sort(r.choice(np.arange(n, m), cnt[i], replace=False))
        labels = r.integers(0, c, n).astype(np.int64)
        comb2 = torch.zeros((n, k2), device="cuda"); h2 = torch.zeros((n, d2), device="cuda")
        loss = torch.zeros(1, device="cuda"); gh1 = torch.zeros((m, d1), device="cuda")
        gw2 = torch.full((d2, k2), 9.0, device="cuda"); gwc = ops.empty_rows(c, d2, "cuda", zero=True)
        gwc.fill_(9.0)
        ops.head_fwd_bwd(h1, d1, dev(idx), dev(cnt), dev(np.arange(n, dtype=np.int32)), w2, 1, wc, dev(labels), 1.0,
                         comb2, h2, None, loss, gh1, gw2, gwc, ws)
        return float(loss.item()), gw2.cpu().numpy().copy(), gwc.cpu().numpy().copy()

    for seed, n in enumerate([n_max, n_small, n_max, n_small, 8, n_max]):
        got = run(n, shared, seed)
        want = run(n, ops.head_ws(n, k2, c, "cuda"), seed)
        assert got[0] == want[0]
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), (n, seed)


@pytest.mark.parametrize("n,world,local", [(100000, 8, True), (1023, 2, False), (1025, 3, True), (1, 4, False), (5000, 1, True),
                                           (4096, 16, False)])
def test_bucket_by_owner_bit_exact(ops, n, world, local):
    """gs_bucket_by_owner == stable sort of the ids by (id % world)."""
    rng = np.random.default_rng(n + world)
    ids = rng.integers(0, 2_400_000, n).astype(np.int32)
    send, perm, counts = ops.bucket_by_owner(dev(ids), world, emit_local=local)
    order = np.argsort(ids % world, kind="stable")
    want = ids[order] // world if local else ids[order]
    assert np.array_equal(counts.cpu().numpy(), np.bincount(ids % world, minlength=world))
    assert np.array_equal(send.cpu().numpy(), want)
    inv = np.empty(n, dtype=np.int64); inv[order] = np.arange(n)
    assert np.array_equal(perm.cpu().numpy(), inv)


@pytest.mark.parametrize("k", [3, 40, None])
def test_ids_past_the_last_csr_row_are_isolated_nodes(ops, k):
    """A node that appears in no edge and whose id is above every id of the CSR (a trailing isolated node of the
    feature table; the reference's defaultdict(set) answers set(), model.py:303) gets an empty tile row instead of
    an out-of-bounds rowptr read -- warp kernel (k <= 32), thread kernel (k > 32), take-all and the ragged tiles."""
    rng = np.random.default_rng(3)
    n = 50
    rowptr, col = random_csr(rng, n, 6)
    nodes = np.array([0, n, 7, n + 100, 2 ** 31 - 1, 3], dtype=np.int32)
    width = (k if k is not None else int(np.diff(rowptr).max())) + 1
    idx, cnt = ops.sample_csr(dev(rowptr), dev(col), n, dev(nodes), k, add_self=True, seed=1, step=1, tag_head=1, width=width)
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    for i, v in enumerate(nodes):
        if v >= n:
            assert cnt[i] == 1 and idx[i, 0] == v and (idx[i, 1:] == -1).all()      # only the self loop
    idx, cnt = ops.sample_csr(dev(rowptr), dev(col), n, dev(nodes), k, add_self=False, seed=1, step=1, tag_head=1, width=width)
    assert (cnt.cpu().numpy()[[1, 3, 4]] == 0).all() and (idx.cpu().numpy()[[1, 3, 4]] == -1).all()
    off, flat = ops.take_all_csr(dev(rowptr), dev(col), dev(nodes), add_self=False)
    lens = np.diff(off.cpu().numpy())
    assert (lens[[1, 3, 4]] == 0).all() and lens[0] == rowptr[1] - rowptr[0]


def test_encoder_graph_covers_isolated_trailing_nodes():
    """Encoder.graph sizes the CSR by the feature table, not by the largest id seen in an edge."""
    import torch.nn as nn
    from graphsage.aggregators import MeanAggregator
    from graphsage.encoders import Encoder
    from collections import defaultdict
    adj = defaultdict(set)
    for a, b in [(0, 1), (1, 2), (2, 3)]:
        adj[a].add(b); adj[b].add(a)
    emb = nn.Embedding(8, 4)
    emb.weight = nn.Parameter(torch.arange(32, dtype=torch.float32).view(8, 4), requires_grad=False)
    enc = Encoder(emb, 4, 3, adj, MeanAggregator(emb, cuda=True), num_sample=2, gcn=True, cuda=True)
    assert enc.graph.num_nodes == 8
    out = enc(torch.tensor([7, 3, 5]))                   # 5 and 7 are isolated: zero neighbour mean -> act(0)
    assert out.shape == (3, 3) and torch.isfinite(out).all()
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, 2].abs().max()) == 0.0


def test_encoder_tensor_core_repeatable(ops):
    """Race regression: the 26 000-row shape (two waves of CTAs, 29-stage split-K partials) run 25 times must
    give the same bits every time and stay inside the 1e-5 bar.  (A parity-aliasing race between the two
    splitter groups once made half of these runs return garbage with a 3-stage ring, profiles/README.md s10.)"""
    n, k_in, d_out, act = 26000, 1204, 128, 1
    g = torch.Generator(device="cuda").manual_seed(7)
    x = ops.empty_rows(n, k_in, "cuda")
    x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
    w = torch.randn(d_out, k_in, device="cuda", generator=g) / k_in ** 0.5
    gh = torch.randn(n, d_out, device="cuda", generator=g)
    h0 = torch.empty((n, d_out), device="cuda")
    ops.encoder_fwd_tc(x, w, act, h0)
    gw0 = torch.empty((d_out, k_in), device="cuda")
    ops.encoder_wgrad_tc(x, h0, gh, act, gw0)
    hd = h0.double()
    assert relerr(h0.cpu().numpy(), torch.relu(x.double() @ w.double().t()).cpu().numpy()) < REL
    assert relerr(gw0.cpu().numpy(), ((gh.double() * (hd > 0).double()).t() @ x.double()).cpu().numpy()) < REL
    for _ in range(25):
        h = torch.full((n, d_out), float("nan"), device="cuda")
        ops.encoder_fwd_tc(x, w, act, h)
        gw = torch.full((d_out, k_in), float("nan"), device="cuda")
        ops.encoder_wgrad_tc(x, h0, gh, act, gw)
        assert torch.equal(h, h0) and torch.equal(gw, gw0)


def test_sampler_rejects_unsupported_fanout(ops):
    """k > 64 is outside the kernel's register-resident Floyd set: an error code, not a wrong answer."""
    rowptr = dev(np.array([0, 2], dtype=np.int64)); col = dev(np.array([0, 0], dtype=np.int32))
    with pytest.raises(RuntimeError, match="not supported"):
        ops.sample_csr(rowptr, col, 1, dev(np.zeros(1, np.int32)), 65)


def test_degenerate_batches(ops):
    """Edge cases the reference hits with tiny graphs: a single target, rows without neighbours (the reference
    divides 0/0 -> NaN, aggregators.py:60-61; documented deviation: zeros), an empty frontier."""
    table = ops.empty_rows(5, 12, "cuda"); table.copy_(torch.arange(60, device="cuda", dtype=torch.float32).view(5, 12))
    idx = dev(np.array([[3, -1, -1]], dtype=np.int32)); cnt = dev(np.array([1], dtype=np.int32))
    out = ops.empty_rows(1, 24, "cuda", zero=True)
    ops.gather_mean_fwd(table, 12, idx, cnt, out, neigh_off=12, self_ids=dev(np.array([4], dtype=np.int32)))
    assert torch.equal(out[0, :12], table[4]) and torch.equal(out[0, 12:], table[3])
    cnt0 = dev(np.array([0], dtype=np.int32))
    ops.gather_mean_fwd(table, 12, idx, cnt0, out, neigh_off=12, self_ids=None)
    assert torch.equal(out[0, 12:], torch.zeros(12, device="cuda"))
    empty_ids = torch.zeros(0, dtype=torch.int32, device="cuda")
    i0, c0 = ops.sample_csr(dev(np.array([0, 1], dtype=np.int64)), dev(np.array([0], dtype=np.int32)), 1, empty_ids, 5)
    assert i0.shape == (0, 5) and c0.shape == (0,)


@pytest.mark.parametrize("add_self", [False, True])
def test_ragged_full_neighbourhood(ops, add_self):
    """num_sample=None over ragged tiles (gs_take_all_*, gs_gather_mean_ragged, gs_scatter_mean_ragged): bit-exact
    index work against the CSR, mean / backward against fp64, on a graph with a 5000-neighbour hub, isolated nodes
    and self loops."""
    rng = np.random.default_rng(21)
    n, dim = 6000, 37
    rowptr, col = random_csr(rng, n, 8, hub=5000)
    nodes = np.concatenate([rng.integers(0, n, 700), np.argsort(-np.diff(rowptr))[:3]]).astype(np.int32)
    off, flat = ops.take_all_csr(dev(rowptr), dev(col), dev(nodes), add_self=add_self)
    off_h, flat_h = off.cpu().numpy(), flat.cpu().numpy()
    want = []
    for v in nodes:
        row = col[rowptr[v]:rowptr[v + 1]].tolist()
        if add_self and v not in row:
            row = row + [int(v)]
        want.append(row)
    assert np.array_equal(np.diff(off_h), [len(r) for r in want]) and off_h[0] == 0
    assert np.array_equal(flat_h, np.concatenate([np.array(r, dtype=np.int32) for r in want]))
    assert max(len(r) for r in want) >= 5000
    table = rng.standard_normal((n, dim)).astype(np.float32)
    t = ops.empty_rows(n, dim, "cuda"); t.copy_(dev(table))
    out = ops.empty_rows(len(nodes), dim, "cuda", zero=True)
    ops.gather_mean_ragged(t, dim, off, flat, out)
    ref = np.stack([table[r].astype(np.float64).mean(0) if r else np.zeros(dim) for r in want])
    assert relerr(out.cpu().numpy(), ref) < REL
    gout = rng.standard_normal((len(nodes), dim)).astype(np.float32)
    g = ops.empty_rows(len(nodes), dim, "cuda"); g.copy_(dev(gout))
    gt = ops.empty_rows(n, dim, "cuda", zero=True)
    ops.scatter_mean_ragged(g, dim, off, flat, gt)
    gref = np.zeros((n, dim))
    for i, r in enumerate(want):
        if r:
            np.add.at(gref, r, gout[i].astype(np.float64) / len(r))
    assert relerr(gt.cpu().numpy(), gref) < REL
    # flat dedup (cnt = NULL): distinct ids ascending, flat rewritten to their ranks
    scratch = ops.DedupScratch(n, "cuda")
    f2 = flat.clone()
    uniq, tot = ops.dedup_remap(f2.view(-1, 1), None, scratch)
    u = np.unique(flat_h)
    assert int(tot.item()) == u.size and np.array_equal(uniq[:u.size].cpu().numpy(), u)
    assert np.array_equal(u[f2.cpu().numpy()], flat_h)
