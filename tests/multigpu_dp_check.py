#!/usr/bin/env python
"""Multi-GPU parity check of the data-parallel path ("8-GPU gradients == 1-GPU gradients
on the concatenated batch"), run under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29534 tests/multigpu_dp_check.py

Every rank trains the fused engine on its slice of each global batch with the gradient all-reduce fused
into the SGD kernel over NVLink peer memory (gs_allreduce_sgd, CUDA-graph replayed, pipelined and
not), then repeats the same steps alone on the whole batches; the weights must agree to 1e-5 (norm-wise)
and be BIT-identical across ranks.  Prints "DP-CHECK OK" from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphsage-simple_b200")):
    sys.path.insert(0, p)


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def main():
    from graphsage import dist as gdist, sampling
    from graphsage.engine import engine_for
    from graphsage.graph import CSRGraph
    from graphsage.model import build_sage
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(8)
    n, f, c, k1, k2 = 6000, 64, 41, 5, 7
    src, dst = rng.integers(0, n, (2, 8 * n))
    ring = np.arange(n)
    graph = CSRGraph.from_edges(np.concatenate([src, ring]), np.concatenate([dst, (ring + 1) % n]), n, device=dev)
    table = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32)).to(dev)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)
    gb = 64 * world
    batches = [np.random.default_rng(70 + s).permutation(n)[:gb] for s in range(7)]
    lr = 0.4

    from graphsage import sharded
    ex = sharded.OwnerExchange(rank, world)
    sh_feats = sharded.ShardedFeatures(table[rank::world].contiguous(), n, exchange=ex, peer=True)
    sh_graph = sharded.ShardedCSR.from_global(graph.rowptr_host, graph.col.cpu().numpy(), rank, world, device=dev,
                                              exchange=ex, peer=True)

    def fresh(partitioned=False):
        torch.manual_seed(5)
        if partitioned:        # table + CSR partitioned over the ranks, read through NVLink peer memory by the fused engine
            model, encs = build_sage(sh_feats, f, [128, 128], sh_graph, [k1, k2], c)
        else:
            emb = torch.nn.Embedding(n, f, device="meta")
            emb.weight = torch.nn.Parameter(table, requires_grad=False)
            model, encs = build_sage(emb, f, [128, 128], graph, [k1, k2], c)
        for i, e in enumerate(encs):
            e.aggregator.uid = 300 + i
        sampling.seed(13)
        return model

    results = {}
    for mode in ("dp", "dp_pipelined", "dp_partitioned", "single"):
        model = fresh(partitioned=(mode == "dp_partitioned"))
        eng = engine_for(model, gb)
        assert eng is not None and eng.head and (eng.table_peer is not None) == (mode == "dp_partitioned")
        if mode != "single":
            eng.peer = gdist.PeerAllreduceSGD(eng.flat_w.numel(), dev)
            eng.grad_scale = gdist.local_grad_scale(gb // world, gb, world)
        losses = []
        for i, nodes in enumerate(batches):
            mine = nodes if mode == "single" else nodes[rank::world]
            step_lr = lr if mode == "single" else gdist.dp_lr(lr, world)
            pre = None
            if mode in ("dp_pipelined", "dp_partitioned"):
                pre = [(batches[j][rank::world], labels[batches[j][rank::world]]) for j in (i + 1, i + 2) if j < len(batches)] or None
            losses.append(model.train_step(mine, labels[mine], lr=step_lr, prefetch=pre))
        torch.cuda.synchronize()
        results[mode] = [p.detach().clone() for p in (model.weight, model.enc.weight, model.enc.base_model.weight)]
        dist.barrier()
    for mode in ("dp", "dp_pipelined", "dp_partitioned"):
        for a, b in zip(results[mode], results["single"]):
            e = relerr(a, b)
            assert e < 1e-5, "%s weights differ from the single-rank run: %g" % (mode, e)
        for a in results[mode]:                      # identical bits on every rank
            ref = a.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(a, ref), "ranks diverged (%s)" % mode
    for a, b in zip(results["dp"], results["dp_pipelined"]):
        assert relerr(a, b) < 1e-5
    if rank == 0:
        print("DP-CHECK OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
