#!/usr/bin/env python
"""Latency of the fused all-reduce + SGD kernel (gs_allreduce_sgd, NVLink peer memory) against NCCL
all_reduce + gs_sgd_step for the gradient block of the Reddit-shape model (192 128 floats), under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/peer_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphsage-simple_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from graphsage import dist as gdist, ops

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 128 * 1204 + 128 * 256 + 41 * 128
w = torch.zeros(n, device=dev)
g = torch.randn(n, device=dev)
peer = gdist.PeerAllreduceSGD(n, dev)


def timed(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def nccl():
    dist.all_reduce(g)
    ops.sgd_step(w, g, 0.0)


t_peer = timed(lambda: peer.step(w, g, 1e-6))
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    peer.step(w, g, 1e-6)
t_peer_graph = timed(gr.replay)
t_nccl = timed(nccl)
t_sgd = timed(lambda: ops.sgd_step(w, g, 0.0))
if rank == 0:
    print("world %d: fused peer all-reduce+SGD %.1f us eager, %.1f us graph-replayed; NCCL all_reduce + SGD %.1f us; "
          "local SGD alone %.1f us" % (world, t_peer, t_peer_graph, t_nccl, t_sgd))
dist.destroy_process_group()
