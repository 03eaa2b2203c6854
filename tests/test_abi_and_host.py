"""CPU-side checks: the C-ABI library loads without a GPU and exports every symbol that
include/gsage.h declares; the ctypes table matches the header; host-side containers."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gsage.h")).read()
    return re.findall(r"GS_API\s+[\w\s\*]+?\b(gs_\w+)\s*\(", text)


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as entry
    return entry.build()


def test_library_exports_every_declared_symbol(lib_path):
    names = header_symbols()
    assert len(names) >= 14
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), n
    lib.gs_abi_version.restype = ctypes.c_int
    assert lib.gs_abi_version() == 1


def test_ctypes_table_matches_header(lib_path):
    from graphsage import _native
    assert sorted(_native.SIGNATURES) == sorted(header_symbols())
    text = open(os.path.join(ROOT, "include", "gsage.h")).read()
    for name, (_, args) in _native.SIGNATURES.items():
        m = re.search(r"GS_API[^;(]*\b%s\s*\(([^;]*?)\)\s*;" % name, text, re.S)
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))
    assert _native.load().gs_strerror(-2).decode().startswith("gsage:")


def test_argument_errors_are_codes_not_crashes(lib_path):
    from graphsage import _native
    lib = _native.load()
    assert lib.gs_sgd_step(None, None, 0.1, 10, None) == -1
    assert lib.gs_gather_rows(None, 0, 4, None, 1, None, None, 0, None) == -1
    assert lib.gs_encoder_bwd_ws_floats(1024, 256, 128) > 0
    assert lib.gs_dedup_scratch_ints(233000) == (233000 + 31) // 32 // 2048 + 2      # bitmap words / 2048 per block + ticket


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graphsage-simple_b200", "graphsage")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_modules_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import torch.nn as nn
    from graphsage.aggregators import MeanAggregator
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MeanAggregator(nn.Embedding(4, 4))


def test_bench_graph_generators_are_consistent():
    """Host logic of bench.py: the per-rank partitioned graph builder (config 5) must produce exactly the rows
    v % world == rank of the single-rank build, symmetric, deduplicated, sorted; the R-MAT variant is symmetric,
    heavy-tailed and leaves no node without an edge."""
    import bench
    n, pairs = 3001, 20000
    rp1, col1 = bench.build_partitioned_graph(n, pairs, 0, 1, seed=5, chunk=4096)
    deg = np.diff(rp1)
    edges = set(zip(np.repeat(np.arange(n), deg).tolist(), col1.tolist()))
    assert all((b, a) in edges for a, b in edges) and len(edges) == col1.size
    assert all(np.all(np.diff(col1[rp1[v]:rp1[v + 1]]) > 0) for v in range(n))
    world = 3
    total = 0
    for rank in range(world):
        rp, col = bench.build_partitioned_graph(n, pairs, rank, world, seed=5, chunk=4096)
        assert rp.shape == (n + 1,)
        for v in range(n):
            row = col[rp[v]:rp[v + 1]]
            if v % world == rank:
                assert np.array_equal(row, col1[rp1[v]:rp1[v + 1]])
            else:
                assert row.size == 0
        total += col.size
    assert total == col1.size
    rp, col = bench.build_graph_arrays(4000, 40000, "rmat")
    d = np.diff(rp)
    e = set(zip(np.repeat(np.arange(4000), d).tolist(), col.tolist()))
    assert d.min() >= 1 and d.max() > 20 * d.mean() and all((b, a) in e for a, b in e)
    rp_u, col_u = bench.build_graph_arrays(4000, 40000)
    assert np.diff(rp_u).max() < 5 * np.diff(rp_u).mean()


def test_every_exported_symbol_is_documented():
    """INTEGRATION.md names every entry point of include/gsage.h next to the reference lines it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in header_symbols() if n not in doc]
    assert not missing, missing


def test_host_helpers():
    """Pure host logic of the drop-in package (no CUDA): sampler tags, data-parallel scaling, CSR builders."""
    from graphsage import dist as gd, sampling
    from graphsage.graph import CSRGraph
    sampling.seed(7)
    assert sampling.get_seed() == 7 and sampling.get_step() == 0
    with sampling.top_level_call() as s1:
        with sampling.top_level_call() as inner:            # nested encoder call: same minibatch, same step
            assert inner == s1 == 1
    with sampling.top_level_call() as s2:
        assert s2 == 2
    assert sampling.call_tag(3, 1) != sampling.call_tag(3, 0) != sampling.call_tag(4, 0)
    nodes, labels = np.arange(10), np.arange(10)
    parts = [gd.shard_batch(nodes, labels, r, 3)[0] for r in range(3)]
    assert np.array_equal(np.concatenate(parts), nodes) and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert abs(sum(gd.local_grad_scale(len(p), 10, 3) for p in parts) / 3 - 1.0) < 1e-12
    assert gd.dp_lr(0.7, 8) == 0.7 / 8
    g = CSRGraph.from_edges([0, 0, 2, 2], [1, 1, 0, 2], 4, device="cpu")       # duplicate edge, self loop, isolated node 3
    assert g.rowptr_host.tolist() == [0, 2, 3, 5, 5] and g.col.tolist() == [1, 2, 0, 0, 2]
    assert g.max_degree == 2 and g.min_degree == 0 and g.to_adj_lists() == {0: {1, 2}, 1: {0}, 2: {0, 2}, 3: set()}
    h = CSRGraph.from_adj_lists({0: {2, 1}, 2: {0}}, num_nodes=4, device="cpu")
    assert h.rowptr_host.tolist() == [0, 2, 2, 3, 3] and h.col.tolist() == [1, 2, 0]


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm: the unmodified reference from oracle/_ref, or the oracle port of its
    dense-mask path, on the host cores) prints exactly one JSON line with the contract's keys, also when launched as rank 1 of a torchrun (no output)."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--nodes", "4000", "--pairs", "40000", "--feat", "20", "--hidden", "16", "--classes", "5", "--cpu-batch", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    from oracle import build_ref
    kind = "reference" if build_ref.verify() else "port"        # oracle/_ref present (built here from /root/reference)?
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] == kind
    assert j["config"]["batch_per_gpu"] == 32 and j["config"]["global_batch"] == 32     # what the arm actually ran
    assert j["cpu_baseline"]["cores"] >= 1 and j["e2e"]["h2d_bytes_per_step"] == 0 and j["vs_baseline"] is None
    if kind == "reference":                                      # the port stays available as a fallback
        rp = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, GSAGE_REFERENCE_PORT="1"))
        assert rp.returncode == 0 and json.loads(rp.stdout.strip())["cpu_baseline"]["kind"] == "port"
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_reference_copy_matches_oracle_port_bit_for_bit():
    """oracle/_ref (the UNMODIFIED reference modules, oracle/build_ref.py) and the oracle port give the same loss and
    the same post-SGD weights, bit for bit, for the same Python RNG state -- the pin of the port that also holds on
    the GPU box, where /root/reference does not exist.  Skipped when oracle/_ref did not travel."""
    import random
    import numpy as np
    import pytest
    import torch
    from oracle import build_ref, ref_path as R, ref_runtime as RR
    if not build_ref.verify():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(0)
    n, f = 400, 20
    adj = {i: set() for i in range(n)}
    for a, b in rng.integers(0, n, (3000, 2)):
        adj[int(a)].add(int(b)); adj[int(b)].add(int(a))
    table = torch.randn(n, f, generator=torch.Generator().manual_seed(0))
    ws = [torch.randn(s, generator=torch.Generator().manual_seed(i)) * 0.1 for i, s in enumerate([(16, 2 * f), (16, 32), (5, 16)])]
    labels = rng.integers(0, 5, (n, 1)).astype(np.int64)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = RR.build_two_layer(table, adj, f, 16, 16, 5, 4, 6, weights=ws)
    opt = RR.make_optimizer(m, 0.7)
    o = R.TwoLayerModel(table, adj, adj, 16, 16, 5, 4, 6, gcn=False, w1=ws[0].clone(), w2=ws[1].clone(), wc=ws[2].clone())
    for step in range(3):
        nodes = list(rng.integers(0, n, 64))
        random.seed(3 + step)
        l_ref = float(RR.train_step(m, opt, nodes, labels[np.array(nodes)]).detach())
        random.seed(3 + step)
        l_port = float(o.train_step(nodes, labels[np.array(nodes)], lr=0.7))
        assert l_ref == l_port
    assert torch.equal(m.enc.base_model.weight.detach(), o.enc1.weight.detach())
    assert torch.equal(m.enc.weight.detach(), o.enc2.weight.detach()) and torch.equal(m.weight.detach(), o.weight.detach())
    import graphsage                                            # the product package is still the one importable as `graphsage`
    assert "graphsage-simple_b200" in graphsage.__file__
