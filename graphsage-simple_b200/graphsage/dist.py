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


class PeerAllreduceSGD:
    """Gradient all-reduce fused with the SGD update over NVLink peer memory (gs_allreduce_sgd): one
    kernel per step, no NCCL call and no host synchronisation, so it sits inside the step's CUDA graph.
    The staging buffer and flag pad live in symmetric memory (torch.distributed._symmetric_memory:
    one allocation per rank, mapped into every peer of the box)."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _native as N
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = int(n_floats)
        assert self.n % 4 == 0
        blocks = N.load().gs_allreduce_sgd_blocks(self.n)
        flag_words = 2 * self.world * blocks                 # flag A (published) + flag B (slice reduced) per (peer, CTA)
        total = 4 * self.n + (flag_words + 3) // 4 * 4       # in[2][n] | out[2][n] | flags
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.stage_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.flag_ptrs = torch.tensor([p + 4 * 4 * self.n for p in ptrs], dtype=torch.int64, device=device)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every pad is zero before the first remote flag lands

    def step(self, flat_w, flat_g, lr):
        """flat_w -= lr * sum_r flat_g_r   (pass lr / world for the global-batch mean, see dp_lr)."""
        from . import ops
        assert flat_w.numel() == self.n and flat_g.numel() == self.n
        return ops.allreduce_sgd(flat_w, flat_g, lr, self.stage_ptrs, self.flag_ptrs, self.rank, self.world, self.state)
