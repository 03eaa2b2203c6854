"""Stand-alone timings of the layer-1 kernels at the bench shape (n1 ~ 25 154 rows, K = 1204, d = 128), with the
tcgen05 GEMM's experiment switches (GSAGE_TC_DEBUG: 1 = no MMA, 2 = no hi/lo split, 8 = no TMA loads), to see which
stage of the GEMM pipeline bounds it.  python tools/exp_kernels.py   (GPU box)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "graphsage-simple_b200"))
import torch
from graphsage import ops

os.environ.setdefault("GSAGE_TC_TRACE", "0")   # the library decides once whether a trace is wanted at all; 0 = not yet

n, k_in, d = 25154, 1204, 128
g = torch.Generator(device="cuda").manual_seed(1)
xs = []
for _ in range(3):                      # rotate 3 x 122 MB inputs: no launch finds its operand in the 126 MB L2
    x = ops.empty_rows(n, k_in, "cuda")
    x.copy_(torch.randn(n, k_in, device="cuda", generator=g))
    xs.append(x)
w = torch.randn(d, k_in, device="cuda", generator=g) / k_in ** 0.5
gh = torch.randn(n, d, device="cuda", generator=g)
h = torch.empty((n, d), device="cuda")
gw = torch.empty((d, k_in), device="cuda")
ws = torch.empty(max(ops.encoder_fwd_tc_ws_floats(k_in, d), ops.encoder_wgrad_tc_ws_floats(n, k_in, d)), device="cuda")


def timed(fn, iters=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


table = ops.empty_rows(233000, 602, "cuda", zero=True)
table.copy_(torch.randn(233000, 602, device="cuda", generator=g))
ids = torch.randint(0, 233000, (n,), device="cuda", generator=g, dtype=torch.int32)
means = []
for _ in range(3):
    m = ops.empty_rows(n, 602, "cuda")
    m.copy_(torch.randn(n, 602, device="cuda", generator=g))
    means.append(m)
for pf in (0,):
    tf = timed(lambda i: ops.sage_encoder_fwd_tc(table, ids, 602, means[i % 3], w, 1, h, ws=ws))
    tw = timed(lambda i: ops.sage_encoder_wgrad_tc(table, ids, 602, means[i % 3], h, gh, 1, gw, ws=ws))
    print("SAGE in-place concat, L2 prefetch %2d chunks ahead: fwd %.1f us   wgrad %.1f us" % (pf, tf, tw))
for dbg in (0,):
    os.environ["GSAGE_TC_DEBUG"] = str(dbg)
    tf = timed(lambda i: ops.encoder_fwd_tc(xs[i % 3], w, 1, h, ws=ws))
    tw = timed(lambda i: ops.encoder_wgrad_tc(xs[i % 3], h, gh, 1, gw, ws=ws))
    print("debug %2d (%s): fwd %.1f us   wgrad %.1f us" % (
        dbg, ",".join(s for b, s in ((1, "no-mma"), (2, "no-split"), (8, "no-tma")) if dbg & b) or "production", tf, tw))
os.environ["GSAGE_TC_DEBUG"] = "0"

# in-kernel clock64 trace of CTA 0 (forward): events per chunk
#   0 producer: stage free   1 producer: TMA issued   4 splitter(q=2): tile landed   5 splitter: A slot free
#   6 splitter: tcgen05.st done   7 MMA warp: operands ready   8 MMA warp: MMAs issued
trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
os.environ["GSAGE_TC_TRACE"] = str(trace.data_ptr())
ops.sage_encoder_fwd_tc(table, ids, 602, means[0], w, 1, h, ws=ws)
torch.cuda.synchronize()
os.environ["GSAGE_TC_TRACE"] = "0"
t = trace.view(64, 16).cpu().numpy()
t0 = t[0, 0]
print("chunk  stage_free tma_issued | landed slot_free st_done | mma_ready mma_issued   (cycles since the first TMA wait)")
for c in range(38):
    r = t[c]
    print("%3d   %8d %8d | %8d %8d %8d | %8d %8d" % ((c,) + tuple(int(r[e] - t0) if r[e] else -1 for e in (0, 1, 4, 5, 6, 7, 8))))
