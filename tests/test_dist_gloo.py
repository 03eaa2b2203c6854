"""World-size-2 gloo test (CPU) of the data-parallel host logic in graphsage/dist.py: two ranks
compute gradients of their shard of a global batch with the oracle, all-reduce the flat gradient
block, apply SGD with lr/world -- and must land on the weights of the single-process step over
the whole batch (SURVEY.md s8e parity: N-rank gradients == 1-rank gradients on the concatenation)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_path as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(g, uneven):
    return R.TwoLayerModel(torch.from_numpy(g["table"]), R.adj_from_tiles(g["hop1"], g["idx1"], g["cnt1"]),
                           R.adj_from_tiles(g["nodes"], g["idx2"], g["cnt2"]), g["w1"].shape[0], g["w2"].shape[0],
                           g["wc"].shape[0], None, None, gcn=bool(g["gcn"]), w1=torch.from_numpy(g["w1"]),
                           w2=torch.from_numpy(g["w2"]), wc=torch.from_numpy(g["wc"]))


def _worker(rank, world, port, uneven, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "graphsage-simple_b200"))
    from graphsage import dist as gd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = dict(np.load(os.path.join(GOLDEN, "model_sage.npz")))
    nodes = g["nodes"][:-1] if uneven else g["nodes"]
    labels = g["labels"][nodes]
    m = _model(g, uneven)
    my_nodes, my_labels = gd.shard_batch(nodes, labels, rank, world)
    loss = m.loss(list(my_nodes), my_labels) * gd.local_grad_scale(len(my_nodes), len(nodes), world)
    for p in m.parameters():
        p.grad = None
    loss.backward()
    params = m.parameters()
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    gd.make_allreduce()(flat)
    gd.sgd_from_summed(params, flat, 0.7, world)
    if rank == 0:
        np.savez(out, *[p.detach().numpy() for p in params])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("uneven", [False, True])
def test_two_rank_data_parallel_equals_single_process(tmp_path, uneven):
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, _free_port(), uneven, out), nprocs=2, join=True)
    got = np.load(out)
    g = dict(np.load(os.path.join(GOLDEN, "model_sage.npz")))
    m = _model(g, uneven)
    nodes = g["nodes"][:-1] if uneven else g["nodes"]
    m.train_step(list(nodes), g["labels"][nodes], lr=0.7)
    for i, p in enumerate(m.parameters()):
        ref = p.detach().numpy()
        assert np.abs(got["arr_%d" % i] - ref).max() <= 1e-5 * np.abs(ref).max()
    if not uneven:      # the even case is the reference's own golden step
        np.testing.assert_allclose(got["arr_0"], g["wc_new"], rtol=1e-5, atol=1e-7)
