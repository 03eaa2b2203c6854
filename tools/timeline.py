#!/usr/bin/env python
"""Per-kernel timeline of the graph-replayed, software-pipelined train step (bench.py workload),
taken with torch.profiler (CUPTI activity records: start, duration, stream of every kernel inside the
replayed CUDA graphs).  Not a benchmark: numbers under a profiler are never reported as bench values;
this shows WHERE the step's time goes (which chain is critical, how much the chains overlap).

    python tools/timeline.py [--steps 6] [--out gpurun_out/timeline.txt]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphsage-simple_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.nn as nn

import bench
from graphsage import ops, sampling
from graphsage.aggregators import MeanAggregator
from graphsage.encoders import Encoder
from graphsage.engine import engine_for
from graphsage.graph import CSRGraph
from graphsage.model import SupervisedGraphSage

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.txt"))
a = ap.parse_args()


class A:
    nodes = 233000; pairs = 5800000; feat = 602; hidden = 128; classes = 41; k1 = 10; k2 = 25


dev = torch.device("cuda")
rowptr, col = bench.build_graph_arrays(A.nodes, A.pairs)
graph = CSRGraph(rowptr, col, dev)
table = ops.empty_rows(A.nodes, A.feat, dev, zero=True)
table.copy_(torch.randn(A.nodes, A.feat, device=dev))
emb = nn.Embedding(A.nodes, A.feat, device="meta")
emb.weight = nn.Parameter(table, requires_grad=False)


def build_model():      # inside a function: the engine recognises the wiring through the lambdas' closure cells
    agg1 = MeanAggregator(emb, cuda=True)
    enc1 = Encoder(emb, A.feat, 128, graph, agg1, num_sample=A.k1, gcn=False, cuda=True)
    agg2 = MeanAggregator(lambda n: enc1(n).t(), cuda=True)
    enc2 = Encoder(lambda n: enc1(n).t(), 128, 128, graph, agg2, num_sample=A.k2, base_model=enc1, gcn=False, cuda=True)
    return SupervisedGraphSage(A.classes, enc2)


model = build_model()
sampling.seed(1)
B = a.batch
eng = engine_for(model, B)
rng = np.random.default_rng(0)
pool = 32
nodes = torch.from_numpy(rng.integers(0, A.nodes, (pool, B)).astype(np.int32)).to(dev)
labels = torch.from_numpy(rng.integers(0, A.classes, (pool, B))).to(dev)
eng.reset_pipeline()
eng.push(nodes[0], labels[0], 1, on_device=True)
eng.push(nodes[1], labels[1], 2, on_device=True)


def step(i):
    eng.push(nodes[(i + 2) % pool], labels[(i + 2) % pool], i + 3, on_device=True)
    eng.step_pipelined(0.01)


for i in range(10):
    step(i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(a.steps):
        step(10 + i)
    torch.cuda.synchronize()
import json
trace_path = a.out + ".trace.json"
os.makedirs(os.path.dirname(a.out), exist_ok=True)
prof.export_chrome_trace(trace_path)
tr = json.load(open(trace_path))
evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
evs.sort(key=lambda e: e["ts"])
t0 = evs[0]["ts"]
lines = []
for e in evs:
    lines.append("%9.1f %9.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["ts"] + e["dur"] - t0, e["dur"],
                                                 e.get("args", {}).get("stream", "?"), e["name"][:90]))
with open(a.out, "w") as f:
    f.write("# start_us end_us dur_us stream kernel   (%d pipelined steps, B=%d, under torch.profiler)\n" % (a.steps, B))
    f.write("\n".join(lines) + "\n")
os.remove(trace_path)
print("\n".join(lines[-80:]))
