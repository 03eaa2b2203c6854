"""Data-parallel plumbing for small graphs: every rank holds the whole CSR and
feature table, takes an equal slice of each global batch of targets, and the flattened weight
gradients are summed with one all-reduce per step (NCCL on GPUs; gloo in the CPU tests).  The
reference has no distributed code at all -- this is the one real exchange step the path has.

The global loss is the mean over ALL targets of the global batch (graphsage/model.py:57, 69 uses
CrossEntropyLoss's mean reduction), so with rank r holding n_r of N targets

    grad = sum_r (n_r / N) * grad_r          (grad_r = gradient of rank r's local mean loss)

``local_grad_scale`` gives the factor n_r * world / N to apply locally so that the all-reduced SUM
divided by ``world`` (folded into the learning rate by ``dp_lr``) is exactly that."""
import torch
import torch.distributed as dist


def shard_batch(nodes, labels, rank, world):
    """Contiguous slice of the global batch for ``rank`` (sizes differ by at most one)."""
    n = len(nodes)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return nodes[lo:hi], labels[lo:hi]


def local_grad_scale(n_local, n_global, world):
    return float(n_local) * world / float(n_global)


def dp_lr(lr, world):
    return lr / world


def make_allreduce(group=None):
    """callable(flat_grads) summing the flat gradient block over the ranks in place."""
    def allreduce(flat):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return flat
    return allreduce


def sgd_from_summed(params, flat_sum, lr, world):
    """Reference SGD step (model.py:237, 250) from an all-reduced gradient sum: p -= lr/world * sum."""
    with torch.no_grad():
        off = 0
        for p in params:
            k = p.numel()
            p.add_(flat_sum[off:off + k].view_as(p), alpha=-dp_lr(lr, world))
            off += k
