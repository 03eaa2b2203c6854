"""Tensor-level wrappers over the C ABI (include/gsage.h).  Each function takes CUDA tensors,
passes raw device pointers + sizes + the current stream to libgsage_sm100.so and raises on a
non-zero return code.  Nothing here computes on the host."""
import torch

from . import _native as N

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2

LAUNCHES = [0]        # kernels of libgsage_sm100.so enqueued through this module (bench.py reports it)

TAG_AGG2 = 2          # layer-2 aggregator over the targets        (RNG draw #2)
TAG_AGG1_HOP = 1      # layer-1 aggregator over the hop-1 uniques  (RNG draw #1)
TAG_AGG1_SELF = 3     # layer-1 aggregator over the batch nodes    (RNG draw #3)


def round4(x):
    return (int(x) + 3) // 4 * 4


def empty_rows(n, d, device, dtype=torch.float32, zero=False):
    """[n, d] fp32 view of a buffer whose leading dimension is a multiple of 4 floats."""
    ld = round4(d)
    buf = (torch.zeros if zero else torch.empty)((max(int(n), 1), ld), device=device, dtype=dtype)
    return buf[:n, :d]


def aligned_rows(x):
    """Return x itself if it is a row-major fp32 CUDA matrix usable by the kernels (unit inner
    stride, ld % 4 == 0, 16-B aligned base) else a padded copy."""
    N.require_cuda(x)
    if (x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0
            and x.stride(0) >= x.shape[1] and x.data_ptr() % 16 == 0):
        return x
    out = empty_rows(x.shape[0], x.shape[1], x.device)
    out.copy_(x)
    return out


def as_ids(nodes, device):
    """list / ndarray / LongTensor of node ids -> int32 CUDA tensor."""
    if isinstance(nodes, torch.Tensor):
        return nodes.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(np.asarray(nodes, dtype=np.int32))).to(device, non_blocking=True)


def sample_csr(rowptr, col, num_nodes, nodes, k, add_self=False, seed=0, step=0, tag_head=0,
               tag_tail=None, n_head=None, n_dev=None, step_dev=None, width=None, idx=None, cnt=None):
    """K1 -- see gs_sample_csr.  k=None means take-all (reference num_sample=None)."""
    lib = N.load()
    N.require_cuda(rowptr, col, nodes)
    n_max = nodes.shape[0]
    kk = -1 if k is None else int(k)
    if width is None:
        if kk < 0:
            raise ValueError("width must be given for take-all sampling")
        width = kk + (1 if add_self else 0)
    width = max(int(width), 1)
    if idx is None:
        idx = torch.empty((max(n_max, 1), width), device=nodes.device, dtype=torch.int32)[:n_max]
    if cnt is None:
        cnt = torch.empty((max(n_max, 1),), device=nodes.device, dtype=torch.int32)[:n_max]
    if tag_tail is None:
        tag_tail = tag_head
    if n_head is None:
        n_head = n_max
    N.check(lib.gs_sample_csr(N.ptr(rowptr), N.ptr(col), int(num_nodes), N.ptr(nodes), n_max, N.ptr(n_dev),
                              kk, width, int(bool(add_self)), int(seed) & (2 ** 64 - 1), int(step), N.ptr(step_dev),
                              int(tag_head), int(tag_tail), int(n_head), N.ptr(idx), N.ptr(cnt), N.stream()),
            "gs_sample_csr")
    LAUNCHES[0] += 1
    return idx, cnt


class DedupScratch:
    """Reusable scratch for gs_dedup_remap over a graph with num_nodes ids."""

    def __init__(self, num_nodes, device):
        lib = N.load()
        self.num_nodes = int(num_nodes)
        # node bitmap + per-word ranks (gs_dedup_remap): zero on entry, left zero by every call
        self.slot_of = torch.zeros(max(self.num_nodes, 64), device=device, dtype=torch.int32)
        self.block_counts = torch.zeros(lib.gs_dedup_scratch_ints(self.num_nodes), device=device, dtype=torch.int32)


def dedup_remap(idx, cnt, scratch, slot_base=0, n_dev=None, uniq=None, n_total=None):
    """Frontier dedup -- see gs_dedup_remap.  Rewrites idx in place; returns (uniq, n_total_dev)."""
    lib = N.load()
    N.require_cuda(idx, cnt)
    n_max, width = idx.shape
    if uniq is None:
        uniq = torch.empty(max(min(n_max * width, scratch.num_nodes), 1), device=idx.device, dtype=torch.int32)
    if n_total is None:
        n_total = torch.zeros(1, device=idx.device, dtype=torch.int32)
    N.check(lib.gs_dedup_remap(N.ptr(idx), N.ptr(cnt), n_max, N.ptr(n_dev), width, scratch.num_nodes,
                               N.ptr(scratch.slot_of), N.ptr(scratch.block_counts), int(slot_base),
                               N.ptr(uniq), N.ptr(n_total), N.stream()), "gs_dedup_remap")
    LAUNCHES[0] += 5
    return uniq, n_total


def gather_mean_fwd(table, dim, idx, cnt, out, neigh_off=0, self_ids=None, n_dev=None):
    lib = N.load()
    N.require_cuda(table, idx, cnt, out)
    n_max, width = idx.shape
    N.check(lib.gs_gather_mean_fwd(N.ptr(table), table.stride(0), int(dim), N.ptr(idx), N.ptr(cnt), width,
                                   N.ptr(self_ids), n_max, N.ptr(n_dev), N.ptr(out), out.stride(0),
                                   int(neigh_off), N.stream()), "gs_gather_mean_fwd")
    LAUNCHES[0] += 1
    return out


def scatter_mean_bwd(gout, dim, idx, cnt, gtable, neigh_off=0, self_ids=None, n_dev=None):
    lib = N.load()
    N.require_cuda(gout, idx, cnt, gtable)
    n_max, width = idx.shape
    N.check(lib.gs_scatter_mean_bwd(N.ptr(gout), gout.stride(0), int(neigh_off), int(dim), N.ptr(idx), N.ptr(cnt),
                                    width, N.ptr(self_ids), n_max, N.ptr(n_dev), N.ptr(gtable), gtable.stride(0),
                                    N.stream()), "gs_scatter_mean_bwd")
    LAUNCHES[0] += 1
    return gtable


def gather_rows(table, dim, ids, out, n_dev=None):
    lib = N.load()
    N.require_cuda(table, ids, out)
    N.check(lib.gs_gather_rows(N.ptr(table), table.stride(0), int(dim), N.ptr(ids), ids.shape[0], N.ptr(n_dev),
                               N.ptr(out), out.stride(0), N.stream()), "gs_gather_rows")
    LAUNCHES[0] += 1
    return out


def encoder_fwd(x, w, act, h, n_dev=None):
    lib = N.load()
    N.require_cuda(x, w, h)
    n_max, k_in = x.shape
    d_out = w.shape[0]
    N.check(lib.gs_encoder_fwd(N.ptr(x), x.stride(0), N.ptr(w), w.stride(0), k_in, d_out, int(act), n_max,
                               N.ptr(n_dev), N.ptr(h), h.stride(0), N.stream()), "gs_encoder_fwd")
    LAUNCHES[0] += 1
    return h


def encoder_bwd_ws_floats(n_max, k_in, d_out):
    return int(N.load().gs_encoder_bwd_ws_floats(int(n_max), int(k_in), int(d_out)))


def encoder_bwd(x, w, h, gh, act, gw, gx=None, dz=None, ws=None, n_dev=None):
    lib = N.load()
    N.require_cuda(x, w, h, gh, gw)
    n_max, k_in = x.shape
    d_out = w.shape[0]
    if dz is None:
        dz = torch.empty((max(n_max, 1), round4(d_out)), device=x.device, dtype=torch.float32)
    if ws is None:
        ws = torch.empty(max(encoder_bwd_ws_floats(n_max, k_in, d_out), 4), device=x.device, dtype=torch.float32)
    N.check(lib.gs_encoder_bwd(N.ptr(x), x.stride(0), N.ptr(w), w.stride(0), N.ptr(h), h.stride(0),
                               N.ptr(gh), gh.stride(0), k_in, d_out, int(act), n_max, N.ptr(n_dev),
                               N.ptr(dz), N.ptr(gw), gw.stride(0), N.ptr(gx), gx.stride(0) if gx is not None else 0,
                               N.ptr(ws), N.stream()), "gs_encoder_bwd")
    LAUNCHES[0] += 2 + (1 if gx is not None else 0) + (1 if encoder_bwd_ws_floats(n_max, k_in, d_out) > d_out * round4(k_in) else 0)
    return gw, gx


def encoder_dgrad(w, h, gh, act, gx, dz=None, n_dev=None):
    """gx = (gh * act'(h)) . w -- the input gradient of K3 alone."""
    lib = N.load()
    N.require_cuda(w, h, gh, gx)
    n_max, d_out = h.shape
    k_in = w.shape[1]
    if dz is None:
        dz = torch.empty((max(n_max, 1), round4(d_out)), device=h.device, dtype=torch.float32)
    N.check(lib.gs_encoder_dgrad(N.ptr(w), w.stride(0), N.ptr(h), h.stride(0), N.ptr(gh), gh.stride(0), k_in, d_out,
                                 int(act), n_max, N.ptr(n_dev), N.ptr(dz), N.ptr(gx), gx.stride(0), N.stream()),
            "gs_encoder_dgrad")
    LAUNCHES[0] += 2
    return gx


def encoder_tc_supported(k_in, d_out):
    return bool(N.load().gs_encoder_tc_supported(int(k_in), int(d_out)))


def encoder_fwd_tc_ws_floats(k_in, d_out):
    return int(N.load().gs_encoder_fwd_tc_ws_floats(int(k_in), int(d_out)))


def encoder_wgrad_tc_ws_floats(n_max, k_in, d_out):
    return int(N.load().gs_encoder_wgrad_tc_ws_floats(int(n_max), int(k_in), int(d_out)))


def encoder_fwd_tc(x, w, act, h, ws=None, n_dev=None):
    """K3 forward on tcgen05 (3xTF32) -- same contract as encoder_fwd."""
    lib = N.load()
    N.require_cuda(x, w, h)
    n_max, k_in = x.shape
    d_out = w.shape[0]
    if ws is None:
        ws = torch.empty(encoder_fwd_tc_ws_floats(k_in, d_out), device=x.device, dtype=torch.float32)
    N.check(lib.gs_encoder_fwd_tc(N.ptr(x), x.stride(0), N.ptr(w), w.stride(0), k_in, d_out, int(act), n_max,
                                  N.ptr(n_dev), N.ptr(h), h.stride(0), N.ptr(ws), N.stream()), "gs_encoder_fwd_tc")
    LAUNCHES[0] += 2
    return h


def encoder_wgrad_tc(x, h, gh, act, gw, ws=None, n_dev=None):
    """Weight gradient of K3 on tcgen05 (3xTF32): gw = (gh * act'(h))^T . x."""
    lib = N.load()
    N.require_cuda(x, h, gh, gw)
    n_max, k_in = x.shape
    d_out = h.shape[1]
    if ws is None:
        ws = torch.empty(encoder_wgrad_tc_ws_floats(n_max, k_in, d_out), device=x.device, dtype=torch.float32)
    N.check(lib.gs_encoder_wgrad_tc(N.ptr(x), x.stride(0), N.ptr(h), h.stride(0), N.ptr(gh), gh.stride(0), k_in, d_out,
                                    int(act), n_max, N.ptr(n_dev), N.ptr(gw), gw.stride(0), N.ptr(ws), N.stream()),
            "gs_encoder_wgrad_tc")
    LAUNCHES[0] += 3
    return gw


def sage_encoder_fwd_tc(table, self_ids, feat_dim, mean, w, act, h, ws=None, n_dev=None):
    """h = act([table[self_ids] | mean] . w^T) on tcgen05; the self rows are gathered from the table inside the
    GEMM (the self half of the combined tile is never materialised) -- see gs_sage_encoder_fwd_tc."""
    lib = N.load()
    N.require_cuda(table, self_ids, mean, w, h)
    n_max, d_out = mean.shape[0], w.shape[0]
    if ws is None:
        ws = torch.empty(encoder_fwd_tc_ws_floats(2 * feat_dim, d_out), device=mean.device, dtype=torch.float32)
    N.check(lib.gs_sage_encoder_fwd_tc(N.ptr(table), table.stride(0), N.ptr(self_ids), int(feat_dim), N.ptr(mean),
                                       mean.stride(0), N.ptr(w), w.stride(0), d_out, int(act), n_max, N.ptr(n_dev),
                                       N.ptr(h), h.stride(0), N.ptr(ws), N.stream()), "gs_sage_encoder_fwd_tc")
    LAUNCHES[0] += 2
    return h


def sage_encoder_wgrad_tc(table, self_ids, feat_dim, mean, h, gh, act, gw, ws=None, n_dev=None):
    """gw [d_out, 2 F] = (gh * act'(h))^T . [table[self_ids] | mean] on tcgen05 -- see gs_sage_encoder_wgrad_tc."""
    lib = N.load()
    N.require_cuda(table, self_ids, mean, h, gh, gw)
    n_max, d_out = mean.shape[0], h.shape[1]
    if ws is None:
        ws = torch.empty(encoder_wgrad_tc_ws_floats(n_max, 2 * feat_dim, d_out), device=mean.device, dtype=torch.float32)
    N.check(lib.gs_sage_encoder_wgrad_tc(N.ptr(table), table.stride(0), N.ptr(self_ids), int(feat_dim), N.ptr(mean),
                                         mean.stride(0), N.ptr(h), h.stride(0), N.ptr(gh), gh.stride(0), d_out,
                                         int(act), n_max, N.ptr(n_dev), N.ptr(gw), gw.stride(0), N.ptr(ws), N.stream()),
            "gs_sage_encoder_wgrad_tc")
    LAUNCHES[0] += 3
    return gw


def classifier_ws_floats(n, d, c):
    return int(N.load().gs_classifier_ws_floats(int(n), int(d), int(c)))


def classifier_xent(h, wc, labels, grad_scale=1.0, logits=None, loss=None, gh=None, gwc=None, ws=None):
    lib = N.load()
    N.require_cuda(h, wc, labels)
    n, d = h.shape
    c = wc.shape[0]
    if ws is None:
        ws = torch.empty(classifier_ws_floats(n, d, c), device=h.device, dtype=torch.float32)
    N.check(lib.gs_classifier_xent(N.ptr(h), h.stride(0), N.ptr(wc), wc.stride(0), N.ptr(labels), d, c, n,
                                   float(grad_scale), N.ptr(logits), logits.stride(0) if logits is not None else 0,
                                   N.ptr(loss), N.ptr(gh), gh.stride(0) if gh is not None else 0,
                                   N.ptr(gwc), gwc.stride(0) if gwc is not None else 0, N.ptr(ws), N.stream()),
            "gs_classifier_xent")
    LAUNCHES[0] += 3 if gwc is not None else 2
    return loss


def head_supported(d1, k2_in, d2, num_classes):
    return bool(N.load().gs_head_supported(int(d1), int(k2_in), int(d2), int(num_classes)))


def head_ws(n, k2_in, num_classes, device):
    """Zeroed workspace for head_fwd_bwd (holds its re-arming reduction tickets)."""
    return torch.zeros(int(N.load().gs_head_ws_floats(int(n), int(k2_in), int(num_classes))), device=device)


def head_fwd_bwd(h1, d1, idx, cnt, self_slots, w2, act2, wc, labels, grad_scale, comb2, h2, logits, loss,
                 gh1, gw2, gwc, ws):
    """Fused outer layer + classifier + loss + backward -- see gs_head_fwd_bwd."""
    lib = N.load()
    N.require_cuda(h1, idx, cnt, w2, wc, labels, comb2, h2, gh1, gw2, gwc, ws)
    n, width = idx.shape
    N.check(lib.gs_head_fwd_bwd(N.ptr(h1), h1.stride(0), int(d1), N.ptr(idx), N.ptr(cnt), width, N.ptr(self_slots),
                                N.ptr(w2), w2.stride(0), w2.shape[0], int(act2), N.ptr(wc), wc.stride(0), wc.shape[0],
                                N.ptr(labels), n, float(grad_scale), N.ptr(comb2), comb2.stride(0),
                                N.ptr(h2), h2.stride(0), N.ptr(logits), logits.stride(0) if logits is not None else 0,
                                N.ptr(loss), N.ptr(gh1), gh1.stride(0), N.ptr(gw2), gw2.stride(0),
                                N.ptr(gwc), gwc.stride(0), N.ptr(ws), N.stream()), "gs_head_fwd_bwd")
    LAUNCHES[0] += 2
    return loss


def head_rows(h1, d1, idx, cnt, self_slots, w2, act2, wc, labels, grad_scale, comb2, h2, logits, gh1, ws):
    """First launch of the fused head (everything row-local) -- see gs_head_rows."""
    lib = N.load()
    N.require_cuda(h1, idx, cnt, w2, wc, labels, comb2, h2, gh1, ws)
    n, width = idx.shape
    N.check(lib.gs_head_rows(N.ptr(h1), h1.stride(0), int(d1), N.ptr(idx), N.ptr(cnt), width, N.ptr(self_slots),
                             N.ptr(w2), w2.stride(0), w2.shape[0], int(act2), N.ptr(wc), wc.stride(0), wc.shape[0],
                             N.ptr(labels), n, float(grad_scale), N.ptr(comb2), comb2.stride(0),
                             N.ptr(h2), h2.stride(0), N.ptr(logits), logits.stride(0) if logits is not None else 0,
                             N.ptr(gh1), gh1.stride(0), N.ptr(ws), N.stream()), "gs_head_rows")
    LAUNCHES[0] += 1


def head_wgrad(comb2, h2, d1, num_classes, sage, loss, gw2, gwc, ws):
    """Second launch of the fused head (weight gradients + mean loss) -- see gs_head_wgrad."""
    lib = N.load()
    N.require_cuda(comb2, h2, gw2, gwc, ws)
    N.check(lib.gs_head_wgrad(N.ptr(comb2), comb2.stride(0), N.ptr(h2), h2.stride(0), int(d1), h2.shape[1],
                              int(num_classes), comb2.shape[0], int(bool(sage)), N.ptr(loss), N.ptr(gw2), gw2.stride(0),
                              N.ptr(gwc), gwc.stride(0), N.ptr(ws), N.stream()), "gs_head_wgrad")
    LAUNCHES[0] += 1


def sgd_step(p, g, lr):
    lib = N.load()
    N.require_cuda(p, g)
    assert p.is_contiguous() and g.is_contiguous() and p.numel() == g.numel()
    N.check(lib.gs_sgd_step(N.ptr(p), N.ptr(g), float(lr), p.numel(), N.stream()), "gs_sgd_step")
    LAUNCHES[0] += 1
    return p


def bucket_by_owner(ids, world, emit_local=False):
    """Stable counting sort of int32 ids by owner = id % world -- see gs_bucket_by_owner.
    Returns (send_ids [n], perm [n], counts [world]) on the device."""
    lib = N.load()
    N.require_cuda(ids)
    n = ids.shape[0]
    dev = ids.device
    scratch = torch.empty(max(lib.gs_bucket_scratch_ints(n, int(world)), 1), device=dev, dtype=torch.int32)
    send_ids = torch.empty(max(n, 1), device=dev, dtype=torch.int32)[:n]
    perm = torch.empty(max(n, 1), device=dev, dtype=torch.int32)[:n]
    counts = torch.empty(int(world), device=dev, dtype=torch.int32)
    N.check(lib.gs_bucket_by_owner(N.ptr(ids), n, None, int(world), int(bool(emit_local)), N.ptr(scratch),
                                   N.ptr(send_ids), N.ptr(perm), N.ptr(counts), N.stream()), "gs_bucket_by_owner")
    LAUNCHES[0] += 3 if n else 0
    return send_ids, perm, counts


def allreduce_sgd(p, g, lr, stage_ptrs, flag_ptrs, rank, world, state):
    """p -= lr * sum over ranks of g, through peer-mapped staging buffers -- see gs_allreduce_sgd."""
    lib = N.load()
    N.require_cuda(p, g, stage_ptrs, flag_ptrs, state)
    N.check(lib.gs_allreduce_sgd(N.ptr(p), N.ptr(g), p.numel(), float(lr), N.ptr(stage_ptrs), N.ptr(flag_ptrs),
                                 int(rank), int(world), N.ptr(state), N.stream()), "gs_allreduce_sgd")
    LAUNCHES[0] += 1
    return p


def gather_rows_peer(table_ptrs, world, ld, dim, ids, out, n_dev=None):
    """gs_gather_rows over a table partitioned by owner = id % world, shards read through peer memory."""
    lib = N.load()
    N.require_cuda(table_ptrs, ids, out)
    N.check(lib.gs_gather_rows_peer(N.ptr(table_ptrs), int(world), int(ld), int(dim), N.ptr(ids), ids.shape[0], N.ptr(n_dev),
                                    N.ptr(out), out.stride(0), N.stream()), "gs_gather_rows_peer")
    LAUNCHES[0] += 1
    return out


def gather_mean_fwd_peer(table_ptrs, world, ld, dim, idx, cnt, out, neigh_off=0, self_ids=None, n_dev=None):
    lib = N.load()
    N.require_cuda(table_ptrs, idx, cnt, out)
    n_max, width = idx.shape
    N.check(lib.gs_gather_mean_fwd_peer(N.ptr(table_ptrs), int(world), int(ld), int(dim), N.ptr(idx), N.ptr(cnt), width,
                                        N.ptr(self_ids), n_max, N.ptr(n_dev), N.ptr(out), out.stride(0), int(neigh_off),
                                        N.stream()), "gs_gather_mean_fwd_peer")
    LAUNCHES[0] += 1
    return out


def sample_csr_peer(rowptr_ptrs, col_ptrs, world, num_nodes, nodes, k, add_self=False, seed=0, step=0, tag_head=0,
                    tag_tail=None, n_head=None, n_dev=None, step_dev=None, width=None, idx=None, cnt=None):
    """gs_sample_csr over a CSR partitioned by owner = id % world, rows read through peer memory."""
    lib = N.load()
    N.require_cuda(rowptr_ptrs, col_ptrs, nodes)
    n_max = nodes.shape[0]
    kk = -1 if k is None else int(k)
    if width is None:
        if kk < 0:
            raise ValueError("width must be given for take-all sampling")
        width = kk + (1 if add_self else 0)
    width = max(int(width), 1)
    if idx is None:
        idx = torch.empty((max(n_max, 1), width), device=nodes.device, dtype=torch.int32)[:n_max]
    if cnt is None:
        cnt = torch.empty((max(n_max, 1),), device=nodes.device, dtype=torch.int32)[:n_max]
    if tag_tail is None:
        tag_tail = tag_head
    if n_head is None:
        n_head = n_max
    N.check(lib.gs_sample_csr_peer(N.ptr(rowptr_ptrs), N.ptr(col_ptrs), int(world), int(num_nodes), N.ptr(nodes), n_max,
                                   N.ptr(n_dev), kk, width, int(bool(add_self)), int(seed) & (2 ** 64 - 1), int(step),
                                   N.ptr(step_dev), int(tag_head), int(tag_tail), int(n_head), N.ptr(idx), N.ptr(cnt),
                                   N.stream()), "gs_sample_csr_peer")
    LAUNCHES[0] += 1
    return idx, cnt


def take_all_csr(rowptr, col, nodes, add_self=False, num_nodes=None):
    """Ragged full-neighbourhood tile of ``nodes`` -- see gs_take_all_count / gs_take_all_fill.
    Returns (off int32 [n+1], flat int32 [total]); one host read of the total."""
    lib = N.load()
    N.require_cuda(rowptr, col, nodes)
    n = nodes.shape[0]
    dev = nodes.device
    num_nodes = rowptr.shape[0] - 1 if num_nodes is None else int(num_nodes)
    length = torch.empty(max(n, 1), device=dev, dtype=torch.int32)
    off = torch.empty(n + 1, device=dev, dtype=torch.int32)
    N.check(lib.gs_take_all_count(N.ptr(rowptr), N.ptr(col), num_nodes, N.ptr(nodes), n, int(bool(add_self)),
                                  N.ptr(length), N.ptr(off), N.stream()), "gs_take_all_count")
    total = int(off[n].item())
    flat = torch.empty(max(total, 1), device=dev, dtype=torch.int32)[:total]
    N.check(lib.gs_take_all_fill(N.ptr(rowptr), N.ptr(col), num_nodes, N.ptr(nodes), n, N.ptr(off), N.ptr(flat),
                                 N.stream()), "gs_take_all_fill")
    LAUNCHES[0] += 3
    return off, flat


def gather_mean_ragged(table, dim, off, flat, out):
    lib = N.load()
    N.require_cuda(table, off, flat, out)
    N.check(lib.gs_gather_mean_ragged(N.ptr(table), table.stride(0), int(dim), N.ptr(off), N.ptr(flat), off.shape[0] - 1,
                                      N.ptr(out), out.stride(0), N.stream()), "gs_gather_mean_ragged")
    LAUNCHES[0] += 1
    return out


def scatter_mean_ragged(gout, dim, off, flat, gtable):
    lib = N.load()
    N.require_cuda(gout, off, flat, gtable)
    N.check(lib.gs_scatter_mean_ragged(N.ptr(gout), gout.stride(0), int(dim), N.ptr(off), N.ptr(flat), off.shape[0] - 1,
                                       N.ptr(gtable), gtable.stride(0), N.stream()), "gs_scatter_mean_ragged")
    LAUNCHES[0] += 1
    return gtable


def advance_step(step_dev):
    N.check(N.load().gs_advance_step(N.ptr(step_dev), N.stream()), "gs_advance_step")
    LAUNCHES[0] += 1


def stage_next(pool, cursor, dst):
    """dst <- pool[*cursor % len(pool)]; ++*cursor  (device-side batch queue, see gs_stage_next)."""
    N.require_cuda(pool, cursor, dst)
    assert pool.dim() == 2 and pool.is_contiguous() and pool.dtype == torch.uint8 and dst.numel() == pool.shape[1]
    N.check(N.load().gs_stage_next(N.ptr(pool), pool.shape[1], pool.shape[0], N.ptr(cursor), N.ptr(dst), N.stream()),
            "gs_stage_next")
    LAUNCHES[0] += 1


def remap_ids(idx, cnt, id_map, n_dev=None):
    """idx[i, j] = id_map[idx[i, j]] in place for the valid entries -- see gs_remap_ids."""
    N.require_cuda(idx, id_map)
    n_max, width = idx.shape
    N.check(N.load().gs_remap_ids(N.ptr(idx), N.ptr(cnt), n_max, N.ptr(n_dev), width, N.ptr(id_map), N.stream()),
            "gs_remap_ids")
    LAUNCHES[0] += 1
    return idx
