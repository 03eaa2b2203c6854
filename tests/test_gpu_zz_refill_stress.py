"""Stress of the tcgen05 GEMMs' refill duty (csrc/gemm_tc.cu) under the conditions of the pipelined step: the SAGE
in-place forward + weight gradient at the bench's layer-1 shape, 60 times, WHILE gather-mean launches of another batch
stream the feature table on a side stream (the co-resident kernel that keeps every SM's L1 request queue full).  The
splitter groups then spend memory-latency-bound time between two barrier waits; with the wait rule of round 2's first
version a late warp could be lapped by its full barrier and the kernel deadlocked (profiles/README.md R2.8,
tools/tc_protocol_sim.py).  Every repeat must give the bits of the run that was made alone, inside the 1e-5 bar.

The file sorts last on purpose: it is the newest check of the most timing-sensitive kernel and must not mask the rest
of the suite (conftest.py ends the session when a GPU test sits in the driver for 10 minutes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-5


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("n,feat,num_nodes,k", [(25154, 602, 233000, 10), (6000, 100, 50000, 15)])
def test_refill_duty_gemms_under_a_concurrent_gather(n, feat, num_nodes, k):
    from graphsage import ops
    d, act = 128, 1
    g = torch.Generator(device="cuda").manual_seed(n + feat)
    table = ops.empty_rows(num_nodes, feat, "cuda", zero=True)
    table.copy_(torch.randn(num_nodes, feat, device="cuda", generator=g))
    ids = torch.randint(0, num_nodes, (n,), device="cuda", generator=g, dtype=torch.int32)
    mean = ops.empty_rows(n, feat, "cuda", zero=True)
    mean.copy_(torch.randn(n, feat, device="cuda", generator=g))
    w = torch.randn(d, 2 * feat, device="cuda", generator=g) / (2 * feat) ** 0.5
    gh = torch.randn(n, d, device="cuda", generator=g)
    ws_f = torch.empty(ops.encoder_fwd_tc_ws_floats(2 * feat, d), device="cuda")
    ws_g = torch.empty(ops.encoder_wgrad_tc_ws_floats(n, 2 * feat, d), device="cuda")

    # alone: the reference bits, checked against fp64
    h0 = torch.empty((n, d), device="cuda")
    ops.sage_encoder_fwd_tc(table, ids, feat, mean, w, act, h0, ws=ws_f)
    gw0 = torch.empty((d, 2 * feat), device="cuda")
    ops.sage_encoder_wgrad_tc(table, ids, feat, mean, h0, gh, act, gw0, ws=ws_g)
    x = torch.cat([table[ids.long()], mean], dim=1).double()
    assert relerr(h0.cpu().numpy(), torch.relu(x @ w.double().t()).cpu().numpy()) < REL
    dz = gh.double() * (h0.double() > 0).double()
    assert relerr(gw0.cpu().numpy(), (dz.t() @ x).cpu().numpy()) < REL
    del x, dz

    # the co-resident traffic: layer-1 gather-mean of another batch (k sampled neighbour rows per row)
    idx = torch.sort(torch.randint(0, num_nodes, (n, k), device="cuda", generator=g, dtype=torch.int32), dim=1).values
    cnt = torch.full((n,), k, device="cuda", dtype=torch.int32)
    other = ops.empty_rows(n, feat, "cuda", zero=True)
    side = torch.cuda.Stream(priority=-1)
    bad = torch.zeros((), dtype=torch.bool, device="cuda")
    h = torch.empty((n, d), device="cuda")
    gw = torch.empty((d, 2 * feat), device="cuda")
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    for _ in range(60):
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.gather_mean_fwd(table, feat, idx, cnt, other)
            ops.gather_mean_fwd(table, feat, idx, cnt, other)
        h.fill_(float("nan"))
        gw.fill_(float("nan"))
        ops.sage_encoder_fwd_tc(table, ids, feat, mean, w, act, h, ws=ws_f)
        ops.sage_encoder_wgrad_tc(table, ids, feat, mean, h0, gh, act, gw, ws=ws_g)
        bad |= (h != h0).any() | (gw != gw0).any()
    main.wait_stream(side)
    torch.cuda.synchronize()
    assert not bool(bad.item())
