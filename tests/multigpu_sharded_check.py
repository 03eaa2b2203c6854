#!/usr/bin/env python
"""Multi-GPU parity check of the partitioned feature table / CSR, run
under torchrun on N GPUs of one box (NCCL over NVLink):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_sharded_check.py

Also runs as a single process (world 1).  Checks, with the real CUDA kernels and NCCL collectives:
  1. ShardedFeatures(ids) is bit-identical to table[ids]           (aggregators.py:62-65)
  2. ShardedCSR.sample(ids) is bit-identical to the local sampler over the whole CSR
  3. three SGD steps of a 3-layer SAGE-mean model on the partitioned graph, global batch split over
     the ranks, end on the same weights as the same steps on ONE rank holding everything
     (norm-wise 1e-5: indices identical, fp32 sums re-associated by the gradient all-reduce).
Prints "SHARDED-CHECK OK" from rank 0 and exits 0, or raises."""
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
    from graphsage import sharded
    from graphsage.graph import CSRGraph
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rng = np.random.default_rng(12)                       # identical on every rank
    n, f, c = 20011, 100, 47
    src, dst = rng.integers(0, n, (2, 12 * n))
    ring = np.arange(n)
    full = CSRGraph.from_edges(np.concatenate([src, ring]), np.concatenate([dst, (ring + 1) % n]), n, device=dev)
    rowptr, col = full.rowptr_host, full.col.cpu().numpy()
    table = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32)).to(dev)
    labels = rng.integers(0, c, (n, 1)).astype(np.int64)

    for peer in (False, True):
        check_mode(peer, sharded, full, rowptr, col, table, labels, n, f, c, rank, world, dev)
    if world > 1:
        dist.barrier()
    if rank == 0:
        print("SHARDED-CHECK OK world=%d (all-to-all and peer-memory lookups)" % world)
    if world > 1:
        dist.destroy_process_group()


def check_mode(peer, sharded, full, rowptr, col, table, labels, n, f, c, rank, world, dev):
    """peer=False: NCCL all-to-all round trips; peer=True: remote rows read over NVLink inside the kernels."""
    from graphsage import sampling
    from graphsage.model import build_sage
    ex = sharded.OwnerExchange(rank, world)
    feats = sharded.ShardedFeatures(sharded.ShardedFeatures.shard_of(table, rank, world).contiguous(), n, exchange=ex,
                                    peer=peer)
    graph = sharded.ShardedCSR.from_global(rowptr, col, rank, world, device=dev, exchange=ex, peer=peer)

    # 1 + 2: lookups (every rank asks for different, duplicate-containing ids; one rank asks for nothing)
    my = np.random.default_rng(50 + rank)
    for m in (5000 + 17 * rank, 0 if rank == world - 1 else 3, 1):
        ids = torch.from_numpy(my.integers(0, n, m).astype(np.int32)).to(dev)
        got = feats(ids)
        assert torch.equal(got, table[ids.long()]), "feature rows differ"
        for k, add_self in ((10, False), (5, True), (None, False)):
            idx, cnt = graph.sample(ids, k, add_self=add_self, seed=5, step=7, tag=3)
            ridx, rcnt = full.sample(ids, k, add_self=add_self, seed=5, step=7, tag=3, width=idx.shape[1])
            assert torch.equal(cnt, rcnt) and torch.equal(idx, ridx), "sampled tiles differ"

    # 3: 3-layer training, partitioned vs everything-on-one-rank
    hidden, fan = [32, 48, 40], [3, 4, 5]
    gb = 96 * world
    batches = [np.random.default_rng(900 + s).permutation(n)[:gb] for s in range(3)]

    def run(model, encs, split):
        sampling.seed(21)
        for i, e in enumerate(encs):
            e.aggregator.uid = 500 + i
        opt = torch.optim.SGD(model.parameters(), lr=0.5)
        losses = []
        for nodes in batches:
            mine = nodes[rank::world] if split else nodes
            opt.zero_grad()
            loss = model.loss(list(mine), torch.LongTensor(labels[mine]))
            loss.backward()
            if split:
                sharded.allreduce_grads(list(model.parameters()), world, len(mine), len(nodes))
            opt.step()
            losses.append(loss.item())
        return losses

    torch.manual_seed(3)
    model_s, encs_s = build_sage(feats, f, hidden, graph, fan, c)
    w0 = [p.detach().clone() for p in model_s.parameters()]
    run(model_s, encs_s, split=True)
    emb = torch.nn.Embedding(n, f, device="meta")
    emb.weight = torch.nn.Parameter(table, requires_grad=False)
    model_l, encs_l = build_sage(emb, f, hidden, full, fan, c)
    model_l.use_engine = False
    with torch.no_grad():
        for p, w in zip(model_l.parameters(), w0):
            p.copy_(w)
    run(model_l, encs_l, split=False)
    for ps, pl in zip(model_s.parameters(), model_l.parameters()):
        e = relerr(ps.detach(), pl.detach())
        assert e < 1e-5, "weights after 3 partitioned steps differ from the single-rank run: %g (peer=%s)" % (e, peer)
    if not peer:
        assert world == 1 or ex.bytes_sent > 0


if __name__ == "__main__":
    main()
