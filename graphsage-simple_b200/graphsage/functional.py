"""autograd.Function wrappers that make the CUDA kernels participate in ``loss.backward()``
the way the reference's ATen ops do (graphsage/model.py:249)."""
import torch

from . import ops


class GatherMean(torch.autograd.Function):
    """to_feats = mask.mm(embed_matrix) of aggregators.py:54-74 without the mask:
    out[i] = mean_j table[idx[i, j]].  Backward is the scatter-add mask^T . g."""

    @staticmethod
    def forward(ctx, table, idx, cnt):
        t = ops.aligned_rows(table)
        n, dim = idx.shape[0], table.shape[1]
        out = ops.empty_rows(n, dim, table.device)
        ops.gather_mean_fwd(t, dim, idx, cnt, out, neigh_off=0)
        ctx.save_for_backward(idx, cnt)
        ctx.rows, ctx.dim = table.shape[0], dim
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, cnt = ctx.saved_tensors
        gtable = None
        if ctx.needs_input_grad[0]:
            gtable = ops.empty_rows(ctx.rows, ctx.dim, gout.device, zero=True)
            ops.scatter_mean_bwd(ops.aligned_rows(gout), ctx.dim, idx, cnt, gtable, neigh_off=0)
        return gtable, None, None


class TableLookup(torch.autograd.Function):
    """rows = weight[ids]: the ``nn.Embedding`` lookups of the reference -- ``features(LongTensor(unique_nodes_list))``
    (aggregators.py:65), ``self.embed(indices)`` (aggregators.py:71), ``self.features(nodes)`` (encoders.py:53) -- as
    gs_gather_rows (bit-exact row copies).  Backward is EmbeddingDenseBackward (model.py:249): a dense gradient of
    the table's shape, rows scatter-added with 128-bit reductions (gs_scatter_mean_bwd with one-entry tile rows)."""

    @staticmethod
    def forward(ctx, weight, ids):
        w = ops.aligned_rows(weight)
        ids = ids.to(torch.int32).contiguous()
        out = ops.empty_rows(ids.shape[0], weight.shape[1], weight.device)
        ops.gather_rows(w, weight.shape[1], ids, out)
        ctx.save_for_backward(ids)
        ctx.rows, ctx.dim = weight.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        (ids,) = ctx.saved_tensors
        gw = None
        if ctx.needs_input_grad[0]:
            gw = ops.empty_rows(ctx.rows, ctx.dim, gout.device, zero=True)
            ones = torch.ones(ids.shape[0], device=gout.device, dtype=torch.int32)
            ops.scatter_mean_bwd(ops.aligned_rows(gout), ctx.dim, ids.view(-1, 1), ones, gw, neigh_off=0)
        return gw, None


class RaggedGatherMean(torch.autograd.Function):
    """The same mean for un-sampled neighbourhoods of any size (num_sample=None, aggregators.py:47-48):
    row i averages table[flat[off[i] : off[i+1]]]."""

    @staticmethod
    def forward(ctx, table, off, flat):
        t = ops.aligned_rows(table)
        n, dim = off.shape[0] - 1, table.shape[1]
        out = ops.empty_rows(n, dim, table.device)
        ops.gather_mean_ragged(t, dim, off, flat, out)
        ctx.save_for_backward(off, flat)
        ctx.rows, ctx.dim = table.shape[0], dim
        return out

    @staticmethod
    def backward(ctx, gout):
        off, flat = ctx.saved_tensors
        gtable = None
        if ctx.needs_input_grad[0]:
            gtable = ops.empty_rows(ctx.rows, ctx.dim, gout.device, zero=True)
            ops.scatter_mean_ragged(ops.aligned_rows(gout), ctx.dim, off, flat, gtable)
        return gtable, None, None


class EncoderGemm(torch.autograd.Function):
    """h[n, d_out] = act(x . w^T): ``F.relu(self.weight.mm(combined.t()))`` of encoders.py:58-61
    in row-major form (the module returns the transposed view)."""

    TC_MIN_ROWS = 512        # below this the tcgen05 kernels' fixed cost (tensor maps, TMEM allocation) is not worth it

    @staticmethod
    def forward(ctx, x, w, act):
        xa, wa = ops.aligned_rows(x), ops.aligned_rows(w)
        h = ops.empty_rows(x.shape[0], w.shape[0], x.device)
        # width-128 layers on enough rows take the tcgen05 3xTF32 kernels (same 1e-5 bar), the rest the fp32 SIMT ones
        ctx.tc = x.shape[0] >= EncoderGemm.TC_MIN_ROWS and ops.encoder_tc_supported(x.shape[1], w.shape[0])
        if ctx.tc:
            ops.encoder_fwd_tc(xa, wa, act, h)
        else:
            ops.encoder_fwd(xa, wa, act, h)
        ctx.save_for_backward(xa, wa, h)
        ctx.act = act
        return h

    @staticmethod
    def backward(ctx, gh):
        xa, wa, h = ctx.saved_tensors
        gw = ops.empty_rows(wa.shape[0], wa.shape[1], gh.device)
        gx = ops.empty_rows(xa.shape[0], xa.shape[1], gh.device) if ctx.needs_input_grad[0] else None
        gha = ops.aligned_rows(gh)
        if ctx.tc:
            ops.encoder_wgrad_tc(xa, h, gha, ctx.act, gw)
            if gx is not None:
                ops.encoder_dgrad(wa, h, gha, ctx.act, gx)
        else:
            ops.encoder_bwd(xa, wa, h, gha, ctx.act, gw, gx)
        return gx, gw, None


class SoftmaxXent(torch.autograd.Function):
    """loss = CrossEntropyLoss()(h . wc^T, labels) (model.py:57, 64-69); the backward
    quantities are produced by the same launch and rescaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, h, wc, labels):
        ha, wa = ops.aligned_rows(h), ops.aligned_rows(wc)
        loss = torch.empty(1, device=h.device, dtype=torch.float32)
        gh = ops.empty_rows(h.shape[0], h.shape[1], h.device)
        gwc = ops.empty_rows(wc.shape[0], wc.shape[1], h.device)
        ops.classifier_xent(ha, wa, labels, 1.0, None, loss, gh, gwc)
        ctx.save_for_backward(gh, gwc)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gh, gwc = ctx.saved_tensors
        return gh * g, gwc * g, None
