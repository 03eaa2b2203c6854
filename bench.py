#!/usr/bin/env python
"""Benchmark of the sample -> aggregate -> update hot path (BASELINE.json metric:
"target nodes/sec, 2-layer SAGE-mean fwd+bwd, 1/2/4/8 B200; gather HBM GB/s").

    python bench.py --gpus 1 --steps 50 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1

Workload (SURVEY.md s8d, config 4 of BASELINE.json): synthetic Reddit-shape graph -- 233 000
nodes, 5.8 M random undirected pairs (CSR ~11.6 M entries), 602-d fp32 features, 41 classes,
2-layer SAGE-mean (concat), hidden 128/128, fan-out 25 at the targets / 10 at hop 1, B targets
per GPU per step, SGD.  One "step" = sampling + forward + backward + SGD update of one batch
(the unit timed at graphsage/model.py:245-252 of the reference).  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

# ONE JSON line on stdout: NCCL writes its banner / INFO lines to stdout by default, so whatever NCCL_DEBUG level the
# caller asked for is kept but routed to stderr (or to the NCCL_DEBUG_FILE the caller named)
if os.environ.get("GSAGE_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = os.environ["GSAGE_NCCL_DEBUG"]
if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "graphsage-simple_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "target nodes/sec, 2-layer SAGE-mean fwd+bwd"
UNIT = "target nodes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="targets per GPU per step")
    ap.add_argument("--nodes", type=int, default=233000)
    ap.add_argument("--pairs", type=int, default=5800000)
    ap.add_argument("--feat", type=int, default=602)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--classes", type=int, default=41)
    ap.add_argument("--k1", type=int, default=10, help="fan-out at hop 1 (inner layer)")
    ap.add_argument("--k2", type=int, default=25, help="fan-out at the targets (outer layer)")
    ap.add_argument("--lr", type=float, default=0.01,
                    help="SGD lr (reference uses 0.7; 0.01 keeps long runs on random labels finite)")
    ap.add_argument("--cpu-batch", type=int, default=256, help="targets per step of the CPU baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--workload", default="reddit", choices=["reddit", "products"],
                    help="reddit: BASELINE config 4 (default, the judged line); products: config 5 -- 3-layer model, "
                         "feature table + CSR partitioned over the ranks, NCCL all-to-all per lookup")
    ap.add_argument("--graph", default="uniform", choices=["uniform", "rmat"],
                    help="uniform: default_rng(1) pairs (SURVEY s8d, the judged line); rmat: heavy-tailed variant "
                         "(a,b,c = 0.57,0.19,0.19, same node count and pair count, every node >= 1 edge) for hub stress")
    ap.add_argument("--partitioned", action="store_true",
                    help="reddit workload on N GPUs with the feature table AND the CSR partitioned by owner = id %% N and read "
                         "through NVLink peer memory by the fused engine's gather / sampler kernels (instead of replicas)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "a2a"],
                    help="products workload: peer = remote rows read over NVLink inside the kernels (symmetric memory); "
                         "a2a = NCCL all-to-all round trip per lookup")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------ workload
def rmat_pairs(n, pairs, rng, a=0.57, b=0.19, c=0.19):
    """R-MAT edge generator (Chakrabarti et al.): one quadrant choice per bit, ids folded into [0, n)."""
    bits = int(np.ceil(np.log2(n)))
    src = np.zeros(pairs, dtype=np.int64)
    dst = np.zeros(pairs, dtype=np.int64)
    for _ in range(bits):
        r = rng.random(pairs)
        src = (src << 1) | (r >= a + b)
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c))
    perm = rng.permutation(1 << bits)                   # scatter the hubs over the id space
    return np.stack([perm[src] % n, perm[dst] % n])


def build_graph_arrays(n, pairs, kind="uniform"):
    """Reddit-shape synthetic graph of SURVEY.md s8d: default_rng(1) pairs, symmetrised,
    deduplicated, rows sorted; every node has degree >= 1 with these sizes (rmat: a ring is added)."""
    rng = np.random.default_rng(1)
    if kind == "rmat":
        e = rmat_pairs(n, pairs, rng)
        ring = np.arange(n, dtype=np.int64)
        e = np.concatenate([e, np.stack([ring, (ring + 1) % n])], axis=1)
    else:
        e = rng.integers(0, n, (2, pairs), dtype=np.int64)
    src = np.concatenate([e[0], e[1]])
    dst = np.concatenate([e[1], e[0]])
    key = np.unique(src * np.int64(n) + dst)
    s = key // n
    d = (key - s * n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(s, minlength=n), out=rowptr[1:])
    return rowptr, d


class LazyAdj(dict):
    """``adj_lists`` view of a CSR for the CPU oracle: builds each node's set on first access
    (inserting ids in ascending order) instead of materialising 11.6 M Python ints up front."""

    def __init__(self, rowptr, col):
        super().__init__()
        self.rowptr, self.col = rowptr, col

    def __missing__(self, v):
        s = set(self.col[self.rowptr[v]:self.rowptr[v + 1]].tolist())
        self[v] = s
        return s


def cpu_reference_rate(args, rowptr, col, table, labels, w1, w2, wc, batch, steps, warmup=1):
    """Time the reference's CPU path on this box's host cores: the UNMODIFIED reference modules from oracle/_ref
    (oracle/build_ref.py; kind "reference") driven through its own loop (model.py:245-250), or, when oracle/_ref did
    not travel, the oracle port of the same formulation (kind "port").  Returns (targets/s, cores, kind, description)."""
    import contextlib
    import random
    import torch
    from oracle import ref_path as R
    from oracle import ref_runtime as RR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    adj = LazyAdj(rowptr, col)
    kind = "port"
    if RR.available() and not os.environ.get("GSAGE_REFERENCE_PORT"):
        try:
            with contextlib.redirect_stdout(sys.stderr):          # the reference's Encoder prints its dimensions
                ref = RR.build_two_layer(table, adj, args.feat, args.hidden, args.hidden, args.classes, args.k1, args.k2,
                                         gcn=False, weights=(w1, w2, wc))
                opt = RR.make_optimizer(ref, args.lr)
            step = lambda nodes: RR.train_step(ref, opt, list(nodes), labels[nodes])
            kind = "reference"
        except Exception as e:                                    # e.g. a dependency of the reference's model.py is absent
            print("oracle/_ref not usable (%s); timing the oracle port" % str(e).splitlines()[0], file=sys.stderr)
    if kind == "port":
        model = R.TwoLayerModel(table, adj, adj, args.hidden, args.hidden, args.classes, args.k1, args.k2,
                                gcn=False, w1=w1, w2=w2, wc=wc)
        step = lambda nodes: model.train_step(list(nodes), labels[nodes], lr=args.lr)
    random.seed(1)
    rng = np.random.default_rng(7)
    times = []
    for it in range(warmup + steps):
        nodes = rng.integers(0, args.nodes, batch)
        t0 = time.perf_counter()
        step(nodes)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = float(np.median(times))
    what = ("the unmodified reference modules (oracle/_ref: graphsage.aggregators / encoders / model.SupervisedGraphSage, "
            "loop of model.py:245-250)" if kind == "reference" else "oracle port of the reference's dense-mask formulation")
    return batch / med, cores, kind, ("%s; bounded sample of the workload: B=%d targets/step (the dense B x U mask of the "
                                      "reference's formulation is 4.75 GB at 512), %d warm-up + %d timed steps, median "
                                      "%.3f s/step, %d torch threads" % (what, batch, warmup, steps, med, cores))


def cpu_reference_rate_stack(rowptr, col, table, labels, feat, hidden, fanouts, classes, lr, batch, steps, warmup=1):
    """CPU arm of config 5: the UNMODIFIED reference classes stacked three deep by the closure recursion of
    model.py:220-221 (oracle/ref_runtime.build_stack), or the oracle port (R.StackedModel) when oracle/_ref did not
    travel, on a bounded batch (the innermost dense mask is batch*16*11 x batch*16*11*6 floats: 0.76 GB at 32)."""
    import contextlib
    import random
    import torch
    from oracle import ref_path as R
    from oracle import ref_runtime as RR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    adj = LazyAdj(rowptr, col)
    if RR.available() and not os.environ.get("GSAGE_REFERENCE_PORT"):
        with contextlib.redirect_stdout(sys.stderr):
            ref = RR.build_stack(table, adj, feat, hidden, fanouts, classes, gcn=False)
            opt = RR.make_optimizer(ref, lr)
        step = lambda nodes: RR.train_step(ref, opt, list(nodes), labels[nodes])
        kind = "reference"
    else:
        torch.manual_seed(2)
        model = R.StackedModel(table, [adj] * len(hidden), hidden, classes, fanouts, gcn=False)
        step = lambda nodes: model.train_step(list(nodes), labels[nodes], lr=lr)
        kind = "port"
    random.seed(2)
    rng = np.random.default_rng(7)
    times = []
    for it in range(warmup + steps):
        nodes = rng.integers(0, table.shape[0], batch)
        t0 = time.perf_counter()
        step(nodes)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return batch / med, cores, kind, ("%s, %d layers; bounded sample: B=%d targets/step, %d warm-up + %d timed steps, "
                                      "median %.3f s/step, %d torch threads" % (
                                          "unmodified reference classes (oracle/_ref) stacked by closure recursion" if kind == "reference"
                                          else "oracle port (StackedModel)", len(hidden), batch, warmup, steps, med, cores))


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)      # the timed regions are tens of milliseconds long

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.dp_mode = "n/a"
    import torch
    rowptr, col = build_graph_arrays(args.nodes, args.pairs, args.graph)
    torch.manual_seed(1)
    table = torch.randn(args.nodes, args.feat)
    labels = np.random.default_rng(1).integers(0, args.classes, (args.nodes, 1)).astype(np.int64)
    ws = []
    for shape in ((args.hidden, 2 * args.feat), (args.hidden, 2 * args.hidden), (args.classes, args.hidden)):
        w = torch.empty(shape)
        torch.nn.init.xavier_uniform_(w)
        ws.append(w)
    t0 = time.perf_counter()
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    rate, cores, kind, sample = cpu_reference_rate(args, rowptr, col, table, labels, *ws, batch=args.cpu_batch,
                                                   steps=steps, warmup=warmup)
    # the config states what THIS arm ran: the reference's dense mask does not fit beyond B = 512, so its batch is
    # --cpu-batch (256), not the 1024 of the B200 arm; one process on the host cores whatever --gpus says
    cfg = workload_config(args, args.cpu_batch)
    cfg["global_batch"] = args.cpu_batch
    cfg["parallelism"] = "one CPU process, %d torch threads (rank 0 only; --gpus %d ignored)" % (cores, args.gpus)
    cfg["b200_arm_batch_per_gpu"] = args.batch
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * args.cpu_batch / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def workload_config(args, batch):
    return {"workload": ("" if getattr(args, "graph", "uniform") == "uniform" else "[R-MAT heavy-tailed variant] ") +
                        "synthetic Reddit-shape graph: %d nodes, %d undirected pairs (CSR ~2x), %d-d fp32 "
                        "features, %d classes, 2-layer SAGE-mean concat, hidden %d/%d, fan-out %d (targets) / "
                        "%d (hop-1), SGD" % (args.nodes, args.pairs, args.feat, args.classes, args.hidden,
                                            args.hidden, args.k2, args.k1),
            "batch_per_gpu": batch, "global_batch": batch * max(args.gpus, 1),
            "parallelism": "dp%d (%s, all-reduce of weight grads: %s)" % (
                args.gpus, "feature table + CSR PARTITIONED by owner = id %% world, remote rows read over NVLink peer memory "
                "inside the gather / sampler kernels" if getattr(args, "partitioned", False) else "graph + features replicated",
                getattr(args, "dp_mode", "none")),
            "l2_policy": "inputs larger than L2: 561 MB feature table, fresh random targets every step",
            "device_loop": getattr(args, "device_loop", None),
            "lr": args.lr}


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from graphsage import ops, sampling
    from graphsage.aggregators import MeanAggregator
    from graphsage.encoders import Encoder
    from graphsage.graph import CSRGraph
    from graphsage.model import SupervisedGraphSage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rowptr, col = build_graph_arrays(args.nodes, args.pairs, args.graph)
    graph = CSRGraph(rowptr, col, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    table = ops.empty_rows(args.nodes, args.feat, dev, zero=True)
    table.copy_(torch.randn(args.nodes, args.feat, device=dev, generator=gen))
    labels_np = np.random.default_rng(1).integers(0, args.classes, (args.nodes, 1)).astype(np.int64)

    emb = nn.Embedding(args.nodes, args.feat, device="meta")
    emb.weight = nn.Parameter(table, requires_grad=False)                  # model.py:214-215
    if args.partitioned:
        from graphsage import sharded
        ex = sharded.OwnerExchange(rank, world)
        emb = sharded.ShardedFeatures(table[rank::world, :args.feat].contiguous(), args.nodes, exchange=ex, peer=True)
        graph = sharded.ShardedCSR.from_global(rowptr, col, rank, world, device=dev, exchange=ex, peer=True)
        if rank != 0 or args.no_cpu_baseline:
            del table                                                      # only the shard stays resident
            torch.cuda.empty_cache()
    torch.manual_seed(1)
    agg1 = MeanAggregator(emb, cuda=True)
    enc1 = Encoder(emb, args.feat, args.hidden, graph, agg1, num_sample=args.k1, gcn=False, cuda=True)
    agg2 = MeanAggregator(lambda nodes: enc1(nodes).t(), cuda=True)
    enc2 = Encoder(lambda nodes: enc1(nodes).t(), enc1.embed_dim, args.hidden, graph, agg2, num_sample=args.k2,
                   base_model=enc1, gcn=False, cuda=True)
    model = SupervisedGraphSage(args.classes, enc2)
    sampling.seed(1)
    w0 = [p.detach().cpu().clone() for p in (enc1.weight, enc2.weight, model.weight)]

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    rng = np.random.default_rng(100 + rank)
    pool = K + W + 8
    pool_nodes = rng.integers(0, args.nodes, (pool, B)).astype(np.int32)
    pool_labels = labels_np[pool_nodes.reshape(-1)].reshape(pool, B)
    from graphsage.engine import engine_for
    eng = engine_for(model, B)
    assert eng is not None, "canonical wiring not recognised"
    allreduce = None
    lr = args.lr
    dp_mode = "none"
    if world > 1:
        from graphsage import dist as gdist
        lr = gdist.dp_lr(args.lr, world)   # mean over the global batch = sum of rank means / world
        eng.grad_scale = gdist.local_grad_scale(B, B * world, world)
        dp_mode = os.environ.get("GSAGE_DP", "peer")
        if dp_mode == "none":            # diagnosis only: no gradient exchange (ranks diverge); never a bench line
            pass
        elif dp_mode == "peer":
            # all-reduce fused into the SGD kernel over NVLink peer memory: stays inside the step's CUDA graph
            try:
                eng.peer = gdist.PeerAllreduceSGD(eng.flat_w.numel(), dev)
            except Exception as e:            # no symmetric-memory support on this box: NCCL between graphs
                print("peer all-reduce unavailable (%s); falling back to NCCL" % str(e).splitlines()[0], file=sys.stderr)
                dp_mode = "nccl"
        ok = torch.tensor([1 if dp_mode in ("peer", "none") else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            dp_mode, eng.peer = "nccl", None
        if dp_mode == "nccl":
            allreduce = gdist.make_allreduce()

    args.dp_mode = {"peer": "fused into the SGD kernel over NVLink peer memory", "nccl": "NCCL between graphs",
                    "none": "n/a" if world == 1 else "DISABLED (diagnostic run, not a result)"}[dp_mode]
    # ---- (1) device-resident throughput: inputs already in HBM, one graph replay per step.
    # Steps are software-pipelined three deep: while batch i is in its GEMM/backward chain, batch
    # i+1 is in its feature gather and batch i+2 in its sampler chain on side streams (all inside
    # one captured graph), so every timed step still contains exactly one sampler chain, one
    # gather and one compute chain.
    d_nodes = torch.from_numpy(pool_nodes).to(dev)
    d_labels = torch.from_numpy(pool_labels).to(dev)
    model.grad_allreduce = allreduce
    eng.reset_pipeline()
    # device-resident inputs: one pre-packed staging block [sampler step | labels | targets] per batch of the
    # pool, already in HBM; a step's staging is ONE device-to-device copy of 12 KB
    warm = W + 10              # every graph variant (4 frontier-set rotations x eager, capture) is replaying before the timed loop
    total = warm + 4 * max(K, 4) + 8          # + three warm-up launches of up to K steps each
    d_blocks = torch.stack([eng.pack_stage(pool_nodes[i % pool], pool_labels[i % pool], i + 1)
                            for i in range(total)]).to(dev)
    eng.push(None, None, None, packed=(d_blocks[0], B))
    eng.push(None, None, None, packed=(d_blocks[1], B))

    step_done = {}

    def device_step(i):
        # the host stays at most two steps ahead of the device (as any loop that consumes a per-step result does):
        # an unthrottled launch loop of 8 ranks on one box was measured SLOWER than the e2e loop (0.383 vs 0.327 ms)
        old = step_done.pop(i - 2, None)
        if old is not None:
            old.synchronize()
        eng.push(None, None, None, packed=(d_blocks[(i + 2) % total], B))
        eng.step_pipelined(lr, allreduce)
        step_done[i] = torch.cuda.Event()
        step_done[i].record()

    # Device-resident loop, K_MULTI steps per graph launch: the staging of a step's inputs is gs_stage_next (block index
    # read from a device cursor), so K_MULTI whole pipelined steps -- each still one stage + sample chain, one gather,
    # one compute chain, one SGD / all-reduce -- replay as ONE captured graph; the per-step host work (two launches
    # per rank, 8 ranks on 16 host cores, each waiting for the slowest rank's flags every step) is paid once per
    # K_MULTI steps (default: as many of the K timed steps as fit a multiple of the four frontier sets, at most 64).
    # GSAGE_BENCH_MULTI=0 restores one launch per step.
    k_multi = int(os.environ.get("GSAGE_BENCH_MULTI", "64")) if allreduce is None else 0
    k_multi = min(k_multi, K)
    k_multi -= k_multi % eng.slots
    cursor = torch.zeros(1, dtype=torch.int64, device=dev)
    multi_done = []

    def device_multi(i):
        """steps i .. i + k_multi - 1"""
        if len(multi_done) >= 2:
            multi_done.pop(0).synchronize()
        eng.run_device_queue(d_blocks, cursor, k_multi, lr)
        ev = torch.cuda.Event()
        ev.record()
        multi_done.append(ev)

    n_single = K % k_multi if k_multi else K
    n_multi = (K - n_single) // k_multi if k_multi else 0
    for i in range(warm):                  # includes the eager + capture iterations of every rotation
        device_step(i)
    i_next = warm
    if n_multi:
        torch.cuda.synchronize()
        step_done.clear()
        for _ in range(3):                 # eager run, capture + replay, replay
            cursor.fill_((i_next + 2) % total)
            device_multi(i_next)
            i_next += k_multi
    torch.cuda.synchronize()
    clocks = ClockSampler(local)           # NVML init takes milliseconds and varies per rank: BEFORE the barrier, or the
    clocks.start()                         # ranks enter the timed region skewed and the fast ones wait inside it
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    multi_done.clear()
    if n_multi:
        cursor.fill_((i_next + 2) % total)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev0.record()
    for _ in range(n_multi):
        device_multi(i_next)
        i_next += k_multi
    for _ in range(n_single):
        device_step(i_next)
        i_next += 1
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    args.device_loop = ("%d graph launches of %d pipelined steps each (inputs staged from a device-resident pool by "
                        "gs_stage_next) + %d single-step launches" % (n_multi, k_multi, n_single)) if n_multi else \
        "one graph launch per step"
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total / 1e3)
    done = eng.sets[(eng.cur - 1) % eng.slots]           # the set whose compute chain ran last
    n1 = int(done.n1_dev.item())
    s1 = int(done.cnt1[:n1].sum().item())
    s2 = int(done.cnt2[:B].sum().item())
    loss_dev = float(eng.loss.item())

    # ---- (2) end to end through the public API with host buffers: per step the NEXT batch's ids +
    # labels go host -> device (pinned staging) and this batch's loss comes back device -> host
    eng.reset_pipeline()
    nxt = lambda i: [(pool_nodes[(i + j) % pool], pool_labels[(i + j) % pool]) for j in (1, 2)]

    def e2e_step(i):
        return model.train_step(pool_nodes[i % pool], pool_labels[i % pool], lr=lr, prefetch=nxt(i), sync=False)

    # Public API of the e2e loop: model.stream_trainer -- feed(host ids, host labels) per minibatch, finish() at the end;
    # every batch's ids + labels go host -> device out of pinned memory and every step's loss comes back device -> host,
    # k_e2e steps per graph launch (the copies are nodes of the graph).  GSAGE_E2E_MULTI=0: one train_step call per step
    # (prefetch of the next two batches, loss of step i consumed after step i+1 was launched), the round-1 loop.
    k_e2e = int(os.environ.get("GSAGE_E2E_MULTI", "4")) if allreduce is None else 0
    e2e_api = ("SupervisedGraphSage.train_step(host ids, host labels, prefetch=[next two host batches], sync=False) -> loss "
               "handle, float() of it one step later")
    if k_e2e:
        e2e_api = ("SupervisedGraphSage.stream_trainer(lr, steps_per_launch=%d): feed(host ids, host labels) per minibatch -> "
                   "losses of the completed steps, finish(); per-step H2D of ids + labels and D2H of the loss are nodes of "
                   "the %d-step graph" % (k_e2e, k_e2e))

        def e2e_run(count, first):
            tr = model.stream_trainer(lr=lr, steps_per_launch=k_e2e)
            got = []
            for i in range(count):
                got += tr.feed(pool_nodes[(first + i) % pool], pool_labels[(first + i) % pool])
            got += tr.finish()
            assert len(got) == count
            return got

        # warm-up streams with the timed run's shape (same head of single steps, same tail): every graph the engine
        # uses is met three times -- eager, capture, replay -- before the timed region, for both pinned-buffer variants
        for _ in range(3):
            e2e_run(3 + 2 * k_e2e + (K - 3) % k_e2e if K > 3 else K, 0)
    else:
        for i in range(3):
            float(e2e_step(i))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0.record()
    if k_e2e:
        e2e_loss = e2e_run(K, 3)[-1]
    else:
        pending = None
        for i in range(K):                      # every step's loss is read back; the read of step i is consumed
            h = e2e_step(3 + i)                 # after step i+1 has been launched (one step of host run-ahead)
            if pending is not None:
                e2e_loss = float(pending)
            pending = h
        e2e_loss = float(pending)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t.item()) / 1e3)
    clocks.stop_flag = True
    clocks.join()
    eng.reset_pipeline()

    # ---- (3) drop-in API exactly as the reference's loop writes it (model.py:245-250)
    api_value = None
    if world == 1:
        opt = torch.optim.SGD(filter(lambda p: p.requires_grad, model.parameters()), lr=lr)
        lab_t = [torch.LongTensor(pool_labels[i]) for i in range(pool)]
        node_l = [list(pool_nodes[i]) for i in range(pool)]
        for i in range(3):
            opt.zero_grad(); loss = model.loss(node_l[i], lab_t[i]); loss.backward(); opt.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            opt.zero_grad()
            loss = model.loss(node_l[(3 + i) % pool], lab_t[(3 + i) % pool])
            loss.backward()
            opt.step()
            loss.item()
        torch.cuda.synchronize()
        api_value = B * K / (time.perf_counter() - t0)

    # ---- (4) per-kernel times (eager launches, CUDA events around each C-ABI call)
    kernels, roofline = None, None
    if rank == 0 and not args.no_kernel_profile:
        peer, eng.peer = eng.peer, None          # rank 0 alone from here on: plain local SGD
        kernels = profile_kernels(eng, B, lr, d_nodes, d_labels)
        eng.peer = peer
        import json as _j
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, which = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            peak, which = float(_j.load(open(peaks_path))["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
        # gather-1 read = (s1 + n1) * F * 4 B (neighbour rows + self rows) + the index read; the
        # combined tile the kernel WRITES is an intermediate and is not algorithmic traffic (reported separately)
        in_place = bool(getattr(eng, "split_self", False))
        if in_place:
            # the engine's default: the self rows are gathered inside the encoder GEMMs, this kernel reads only the s1
            # sampled neighbour rows (+ tile indices and counts) and writes the mean half of the tile
            g1_bytes = s1 * args.feat * 4 + (s1 + n1) * 4
            g1_write = n1 * args.feat * 4
        else:
            g1_bytes = (s1 + n1) * args.feat * 4 + (s1 + 2 * n1) * 4
            g1_write = n1 * 2 * args.feat * 4
        g1_ms = kernels["gather_mean_fwd[layer1]"]
        ach = g1_bytes / (g1_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and (args.nodes, args.feat, args.batch, args.k1, args.k2, args.graph) == (233000, 602, 1024, 10, 25, "uniform"):
            tj = _j.load(open(tpath))           # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            if bool(tj.get("in_place_concat", False)) == in_place:
                traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
                traffic_src = tj["source"]
        if args.partitioned and world > 1:      # the gather's rows come over NVLink: bound by the 900 GB/s/direction ingress
            remote = (s1 + n1) * args.feat * 4 * (world - 1) / world
            nv = remote / (g1_ms * 1e-3) / 1e9
            roofline = {"kernel": "gather_mean_kernel<PEER> (layer 1; %d/%d of the rows read from peers over NVLink)" % (world - 1, world),
                        "bound": "nvlink", "achieved": nv, "peak": 900.0, "unit": "GB/s", "frac": nv / 900.0,
                        "peak_source": "NVLink 5 per-direction bandwidth per GPU (B200_PROFILING.md)", "traffic": None,
                        "algorithmic_remote_bytes_per_launch": remote, "rows_read": s1 + n1, "n1": n1, "s1": s1, "s2": s2,
                        "avg_launch_ms": g1_ms, "step_share": g1_ms / sum(kernels.values()),
                        "timing": "rank 0's kernel launched alone (eager, CUDA events), peers idle"}
        else:
            roofline = {"kernel": ("gather_mean_kernel (layer 1: mean of k1 sampled neighbour rows -> mean half of the tile; the self "
                                   "rows are gathered inside the tcgen05 encoder GEMMs)") if in_place else
                                  "gather_mean_kernel (layer 1: self row + mean of k1 neighbour rows -> comb1)",
                        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "peak_source": which, "traffic": traffic, "traffic_source": traffic_src,
                        "timing": "kernel launched alone (eager, CUDA events on its stream, 5 launches), same launch "
                                  "configuration as inside the pipelined step",
                        "algorithmic_bytes_per_launch": g1_bytes,
                        "algorithmic_bytes": ("SURVEY s8(d) for the rows THIS kernel reads: s1 * F * 4 neighbour-row bytes + (s1 + n1) * 4 "
                                              "index bytes (the n1 self rows of s8(d)'s gather-1 are read by the GEMMs' producers, not here); "
                                              "the intermediate tile the kernel writes is NOT counted") if in_place else
                                             ("SURVEY s8(d): (s1 + n1) * F * 4 row bytes + (s1 + 2 * n1) * 4 index bytes; "
                                              "the intermediate tile the kernel writes is NOT counted"),
                        "intermediate_write_bytes_per_launch": g1_write,
                        "achieved_with_intermediate_write": (g1_bytes + g1_write) / (g1_ms * 1e-3) / 1e9,
                        "frac_with_intermediate_write": (g1_bytes + g1_write) / (g1_ms * 1e-3) / 1e9 / peak,
                        "in_step_frac": g1_bytes / (ms_total / K * 1e-3) / 1e9 / peak,
                        "in_step_note": "same algorithmic bytes over the WHOLE pipelined step (the gather of batch t+1 "
                                        "runs concurrently with the compute chain of batch t and may stretch to the step)",
                        "rows_read": s1 + n1, "n1": n1, "s1": s1, "s2": s2, "avg_launch_ms": g1_ms,
                        "step_share": g1_ms / sum(kernels.values())}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        rate, cores, kind, sample = cpu_reference_rate(args, rowptr, col, table[:, :args.feat].cpu().contiguous(), labels_np,
                                                       *w0, batch=args.cpu_batch, steps=args.cpu_steps)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(args, B),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 + 12 * B, "d2h_bytes_per_step": 4,
                        "api": e2e_api,
                        "reference_loop_api": api_value},
                "gpu_launches": eng.launches_per_step * K if hasattr(eng, "launches_per_step") else None,
                "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu, "kernels_ms": kernels,
                "loss": loss_dev, "e2e_last_loss": e2e_loss}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def profile_kernels(eng, B, lr, d_nodes, d_labels, iters=5):
    """Average device time of every C-ABI call of one step, measured with CUDA events on the
    launching stream (eager launches, not under a profiler)."""
    import torch
    from graphsage import ops
    names, totals = [], {}
    orig = {}
    stack = []

    def wrap(name):
        fn = getattr(ops, name)
        orig[name] = fn

        def timed(*a, **kw):
            label = name[:-5] if name.endswith("_peer") else name
            if label in ("gather_mean_fwd", "encoder_fwd", "encoder_bwd", "sample_csr", "encoder_fwd_tc",
                         "encoder_wgrad_tc", "sage_encoder_fwd_tc", "sage_encoder_wgrad_tc"):
                label += "[layer1]" if (kw.get("n_dev") is not None) else "[layer2]"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            stack.append((label, e0, e1))
            return out
        setattr(ops, name, timed)

    for nm in ("sample_csr", "sample_csr_peer", "gather_mean_fwd_peer", "dedup_remap", "gather_mean_fwd", "encoder_fwd", "encoder_fwd_tc",
               "sage_encoder_fwd_tc", "sage_encoder_wgrad_tc", "classifier_xent",
               "encoder_bwd", "encoder_wgrad_tc", "encoder_dgrad", "scatter_mean_bwd", "head_rows", "head_wgrad", "sgd_step"):
        wrap(nm)
    try:
        for it in range(iters + 1):
            eng.stage_device(d_nodes[it], d_labels[it], 1000 + it)
            stack.clear()
            eng._forward_backward(B)
            eng._update(lr)
            torch.cuda.synchronize()
            if it == 0:
                continue
            for label, e0, e1 in stack:
                totals[label] = totals.get(label, 0.0) + e0.elapsed_time(e1)
    finally:
        for nm, fn in orig.items():
            setattr(ops, nm, fn)
    return {k: v / iters for k, v in totals.items()}


# ------------------------------------------------------------------------------ config 5
def build_partitioned_graph(n, pairs, rank, world, seed=2, chunk=1 << 24):
    """Rows v % world == rank of the symmetrised, deduplicated random graph of SURVEY.md s8d
    (products-shape: default_rng(2) pairs).  Every rank draws the same pair stream and keeps its rows,
    so no rank ever materialises the whole 124 M-entry CSR."""
    rng = np.random.default_rng(seed)
    keys = []
    left = pairs
    while left > 0:
        m = min(chunk, left)
        e = rng.integers(0, n, (2, m), dtype=np.int64)
        for a, b in ((e[0], e[1]), (e[1], e[0])):
            sel = (a % world) == rank
            keys.append(a[sel] * np.int64(n) + b[sel])
        left -= m
    key = np.unique(np.concatenate(keys))
    del keys
    s = key // n
    col = (key - s * n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(s, minlength=n), out=rowptr[1:])
    return rowptr, col


def profile_steps(step_fn, n_steps, out_path, title):
    """Per-kernel-name device time of ``n_steps`` calls of ``step_fn`` under torch.profiler (CUPTI activity records, also
    inside replayed CUDA graphs), written as a sorted table.  A diagnostic (shares, launch counts): never a bench value."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_steps):
            step_fn()
        torch.cuda.synchronize()
    tmp = out_path + ".trace.json"
    prof.export_chrome_trace(tmp)
    tr = json.load(open(tmp))
    os.remove(tmp)
    evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    if not evs:
        return
    agg = {}
    for e in evs:
        a = agg.setdefault(e["name"][:110], [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    span = max(e["ts"] + e["dur"] for e in evs) - min(e["ts"] for e in evs)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    with open(out_path, "w") as f:
        f.write("# %s\n# %d steps under torch.profiler: wall span %.1f us/step, sum of kernel time %.1f us/step, %d launches/step\n" % (
            title, n_steps, span / n_steps, sum(v[1] for v in agg.values()) / n_steps, len(evs) // n_steps))
        f.write("# us/step  launches/step  kernel\n")
        for name, (cnt, dur) in rows:
            f.write("%9.1f %6.1f  %s\n" % (dur / n_steps, cnt / n_steps, name))


def gstep_launches(gstep):
    return getattr(gstep, "launches_per_step", 0)


def run_products(args):
    """BASELINE config 5: synthetic ogbn-products-shape graph, 3-layer SAGE-mean (fan-out 15/10/5 from the
    targets outward), feature table AND CSR partitioned by owner = id % world, every lookup an NCCL
    all-to-all round trip (graphsage/sharded.py), weight gradients all-reduced.  Op-by-op autograd path."""
    import torch
    import torch.distributed as dist
    from graphsage import ops, sampling, sharded
    from graphsage.model import build_sage
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.nodes if args.nodes != 233000 else 2400000
    pairs = args.pairs if args.pairs != 5800000 else 62000000
    feat = args.feat if args.feat != 602 else 100
    classes = args.classes if args.classes != 41 else 47
    fan = [5, 10, 15]                                    # innermost first
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    t0 = time.perf_counter()
    rowptr, col = build_partitioned_graph(n, pairs, rank, world)
    deg_max = torch.tensor([int(np.diff(rowptr).max())], device=dev)
    if world > 1:
        dist.all_reduce(deg_max, op=dist.ReduceOp.MAX)
    ex = sharded.OwnerExchange(rank, world)
    peer = args.exchange == "peer"
    graph = sharded.ShardedCSR(rowptr, col, n, int(deg_max.item()), ex, dev, peer=peer)
    entries_local = int(col.shape[0])
    host_csr = (rowptr, col) if (world == 1 and not args.no_cpu_baseline) else None      # N = 1: the CPU baseline's graph
    del rowptr, col
    gen = torch.Generator(device=dev)
    gen.manual_seed(2)
    full = torch.randn(n, feat, device=dev, generator=gen)      # 0.96 GB transient; each rank keeps its rows
    feats = sharded.ShardedFeatures(full[rank::world].contiguous(), n, exchange=ex, peer=peer)
    host_table = full.cpu() if host_csr is not None else None
    del full
    labels_np = np.random.default_rng(2).integers(0, classes, (n, 1)).astype(np.int64)
    torch.manual_seed(2)
    model, encs = build_sage(feats, feat, [args.hidden] * 3, graph, fan, classes)
    sampling.seed(2)
    build_s = time.perf_counter() - t0
    rng = np.random.default_rng(200 + rank)
    params = list(model.parameters())
    graphed = peer and not os.environ.get("GSAGE_PRODUCTS_EAGER")
    if graphed:
        # forward + backward (+ SGD on one GPU) of the whole 3-layer step captured as ONE CUDA graph: aggregators in
        # static-shape mode (no size read back from the device), sampler step in device memory (model.GraphedStep)
        from graphsage.model import GraphedStep
        gstep = GraphedStep(model, B, lr=args.lr, world=world, n_global=B * world)

        def step():
            nodes = rng.integers(0, n, B)
            return gstep(nodes, labels_np[nodes])
    else:
        opt = torch.optim.SGD(model.parameters(), lr=args.lr)

        def step():
            nodes = rng.integers(0, n, B)
            opt.zero_grad()
            loss = model.loss(nodes, labels_np[nodes])
            loss.backward()
            sharded.allreduce_grads(params, world, B, B * world)
            opt.step()
            return loss

    for _ in range(W + (3 if graphed else 0)):           # graphed: two eager steps, the capture, then replays
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)           # before the barrier (NVML init skews the ranks)
    clocks.start()
    sent0, launches0 = ex.bytes_sent, ops.LAUNCHES[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(K):
        loss = step()
        loss_host = loss.item()                                 # D2H read of the step's result
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks.stop_flag = True
    clocks.join()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total / 1e3)
    sent = (ex.bytes_sent - sent0) / K
    if rank == 0 and os.environ.get("GSAGE_PROFILE_OUT") and world == 1:
        profile_steps(step, 3, os.environ["GSAGE_PROFILE_OUT"], "products-shape 3-layer step, %d GPU(s), B=%d" % (world, B))
    # roofline of the partitioned path's dominant kernel: the innermost feature lookup, gs_gather_rows_peer over the
    # padded hop-3 frontier (B * 16 * 11 * 5 ids of 100 floats), (world-1)/world of whose rows cross NVLink; rank 0's
    # launch timed alone with CUDA events while the peers idle
    roof = None
    if peer:
        n_rows = min(B * 16 * 11 * 5, n)
        ids = torch.randint(0, n, (n_rows,), device=dev, dtype=torch.int32)
        out = torch.empty((n_rows, feats.ld), device=dev)
        if rank == 0:
            for _ in range(2):
                ops.gather_rows_peer(feats.table_ptrs, world, feats.ld, feats.ld, ids, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gather_rows_peer(feats.table_ptrs, world, feats.ld, feats.ld, ids, out)
            e1.record()
            torch.cuda.synchronize()
            g_ms = e0.elapsed_time(e1) / 5
            row_bytes = n_rows * feat * 4
            if world > 1:
                remote = row_bytes * (world - 1) / world
                ach = remote / (g_ms * 1e-3) / 1e9
                roof = {"kernel": "gather_rows_kernel<PEER> (innermost feature lookup, %d/%d of the rows read from peers over NVLink)" % (world - 1, world),
                        "bound": "nvlink", "achieved": ach, "peak": 770.0, "unit": "GB/s", "frac": ach / 770.0,
                        "peak_source": "measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md; 900 nominal)",
                        "traffic": None, "algorithmic_remote_bytes_per_launch": remote, "rows": n_rows, "avg_launch_ms": g_ms,
                        "timing": "rank 0's kernel launched alone (eager, CUDA events), peers idle"}
            else:
                peak = 6650.0
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
                    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
                ach = (row_bytes + n_rows * 4) / (g_ms * 1e-3) / 1e9
                roof = {"kernel": "gather_rows_kernel (innermost feature lookup, one GPU: all rows local)", "bound": "hbm",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "algorithmic_bytes_per_launch": row_bytes + n_rows * 4, "rows": n_rows, "avg_launch_ms": g_ms,
                        "timing": "kernel launched alone (eager, CUDA events); rows read once, written once (write not counted)"}
        if world > 1:
            dist.barrier()
    cpu = None
    if rank == 0 and host_csr is not None:
        rate, cores, kind, sample = cpu_reference_rate_stack(host_csr[0], host_csr[1], host_table, labels_np, feat,
                                                             [args.hidden] * 3, fan, classes, args.lr,
                                                             batch=min(args.cpu_batch, 32), steps=args.cpu_steps)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if rank == 0:
        wire = sent / (ms_total / K * 1e-3) / 1e9
        line = {"metric": METRIC.replace("2-layer", "3-layer"), "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "synthetic ogbn-products-shape graph: %d nodes, %d undirected pairs (CSR ~2x), "
                                       "%d-d fp32 features, %d classes, 3-layer SAGE-mean concat, hidden %d, fan-out "
                                       "15/10/5 from the targets outward, SGD" % (n, pairs, feat, classes, args.hidden),
                           "batch_per_gpu": B, "global_batch": B * world,
                           "parallelism": "feature table + CSR partitioned by owner = id %% %d; %s; all-reduce of weight grads" % (
                               world, "remote rows read over NVLink peer memory inside the gather / sampler kernels"
                               if peer else "NCCL all-to-all per lookup"),
                           "l2_policy": "inputs larger than L2: fresh random targets every step", "lr": args.lr,
                           "csr_entries_rank0": entries_local, "build_s": build_s},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 4,
                        "api": ("model.GraphedStep(model, B)(host ids, host labels) -> loss, .item() of it every step: "
                                "the reference's loop body (model.py:245-250) captured as one CUDA graph") if graphed else
                               "the reference's loop (model.py:245-250) on the drop-in modules with host ids/labels"},
                "gpu_launches": (gstep_launches(gstep) * K) if graphed else ops.LAUNCHES[0] - launches0,
                "clocks": clocks.summary(),
                "roofline": roof if peer else {"kernel": "all-to-all exchange (ids + feature rows + sampled tiles), rank 0 send side",
                             "bound": "nvlink", "achieved": wire, "peak": 900.0, "unit": "GB/s", "frac": wire / 900.0,
                             "traffic": None, "bytes_sent_per_step_rank0": sent,
                             "note": "op-by-op path: the step is launch/sync-bound, not wire-bound (one host sync per lookup)"},
                "cpu_baseline": cpu, "loss": loss_host}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def start_watchdog():
    """A run that stops making progress (a hung kernel keeps every later CUDA call of this rank, and through the peer
    all-reduce every other rank, waiting forever) ends with a message on stderr and exit code 3 instead of sitting in
    cudaStreamSynchronize until some outer limit kills it.  GSAGE_BENCH_WATCHDOG_S=0 disables it."""
    limit = float(os.environ.get("GSAGE_BENCH_WATCHDOG_S", "1500"))
    if limit <= 0:
        return

    def fire():
        print("bench.py: no result after %.0f s (rank %s): giving up -- a kernel or a peer rank hangs"
              % (limit, os.environ.get("RANK", "0")), file=sys.stderr, flush=True)
        os._exit(3)
    t = threading.Timer(limit, fire)
    t.daemon = True
    t.start()


if __name__ == "__main__":
    a = parse()
    start_watchdog()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "products":
        run_products(a)
    else:
        run_b200(a)
