"""Fused 2-layer train step: sample -> aggregate -> update with no host round trip.

This is the unit the reference times at graphsage/model.py:245-252 (zero_grad, loss,
backward, SGD step) for the wiring of model.py:214-227: a frozen feature table, layer-1
Encoder, layer-2 Encoder reached through the closure ``lambda nodes: enc1(nodes).t()``, and
the SupervisedGraphSage classifier.  Instead of recursing through Python closures and
autograd, the engine lays the whole step out over statically allocated HBM buffers:

  targets[B] --K1--> tile2[B,k2] --dedup--> frontier1 = [targets | distinct hop-1 ids]
  frontier1[n1] --K1--> tile1[n1,k1] --K2--> comb1[n1, 2F] = [x_v | mean x_u] --K3--> h1[n1,d1]
  tile2 (slots into h1) --K2--> comb2[B, 2d1] = [h1_t | mean h1_u] --K3--> h2[B,d2]
  h2 --K5--> loss, dh2, dWc --K3'--> dW2, dcomb2 --K4--> dh1 --K3'--> dW1 --(NCCL)--> K6 SGD

``n1`` (the number of layer-1 evaluations, B + |distinct hop-1 ids|) only exists in device
memory; every kernel takes (n_max, n_dev), so the sequence is a fixed list of launches and is
captured once per batch size into a CUDA graph.  All arithmetic is in the kernels of
libgsage_sm100.so; torch supplies memory, streams, graphs and NCCL.
"""
import numpy as np
import torch

from . import ops, sampling

_SIGMOID = ("node_degree", "shared", "pagerank")


class _FrontierSet:
    """Everything one minibatch's sample -> gather chain produces, plus its staged inputs.  Three
    sets exist when pipelining: batch t+2 is being sampled and batch t+1 gathered while batch t
    is in its GEMM/backward chain (``TrainEngine.step_pipelined``)."""

    def __init__(self, eng):
        dev, B, n1_max = eng.dev, eng.B, eng.n1_max
        i32 = dict(device=dev, dtype=torch.int32)
        # per-step inputs: one pinned staging block [step | labels | targets] -> one H2D copy
        self.stage_host = torch.empty(8 + 8 * B + 4 * B, dtype=torch.uint8).pin_memory()
        self.stage_dev = torch.zeros(8 + 8 * B + 4 * B, dtype=torch.uint8, device=dev)
        self.stage_event = None
        self.step_dev = self.stage_dev[:8].view(torch.int64)
        self.labels = self.stage_dev[8:8 + 8 * B].view(torch.int64)
        self.targets = self.stage_dev[8 + 8 * B:].view(torch.int32)
        self.idx2 = torch.empty((B, eng.w2_width), **i32)
        self.cnt2 = torch.empty(B, **i32)
        self.frontier1 = torch.zeros(n1_max, **i32)
        self.n1_dev = torch.zeros(1, **i32)
        self.idx1 = torch.empty((n1_max, eng.w1_width), **i32)
        self.cnt1 = torch.empty(n1_max, **i32)
        # SAGE on the tensor-core path: only the neighbour-mean half [n1, F] exists; the self half of the combined
        # tile is gathered from the feature table inside the GEMMs (gs_sage_encoder_*_tc)
        self.comb1 = ops.empty_rows(n1_max, eng.F if eng.split_self else eng.K1, dev, zero=True)


class TrainEngine:
    def __init__(self, graph1, graph2, table, feat_dim, d1, d2, num_classes, k1, k2, max_batch,
                 gcn=False, agg_gcn1=False, agg_gcn2=False, act1=ops.ACT_RELU, act2=ops.ACT_RELU,
                 uid1=1, uid2=2, device=None, use_graphs=True, embed=None, hot_map=None):
        self.dev = torch.device(device) if device is not None else (
            table.table.device if hasattr(table, "table_ptrs") else table.device)
        self.g1, self.g2 = graph1, graph2
        # ``table`` is the [N, F] feature matrix, or a sharded.ShardedFeatures(peer=True): the table partitioned over
        # the ranks and read through NVLink peer memory by the gather kernel itself
        self.table_peer = table if hasattr(table, "table_ptrs") else None
        self.table = table.table if self.table_peer is not None else ops.aligned_rows(table)
        # 1hot / node_degree initialisers (aggregators.py:30-31, 68-71): layer 1 aggregates rows of the TRAINABLE table
        # ``embed`` [rows, feat_dim], reached through ``hot_map`` (node id -> position of the 1 in its one-hot feature
        # row); the table joins the flat parameter block, its gradient is the scatter-add K4 and SGD updates it densely
        self.trainable_table = embed is not None
        if self.trainable_table:
            if not gcn or hot_map is None or self.table_peer is not None:
                raise ValueError("trainable table: gcn encoders on a local table only (the reference driver's wiring)")
            self.hot_map = hot_map.to(torch.int32).contiguous()
            feat_dim = embed.shape[1]
        self.F, self.d1, self.d2, self.C = int(feat_dim), int(d1), int(d2), int(num_classes)
        self.k1, self.k2 = k1, k2
        self.gcn, self.agg_gcn1, self.agg_gcn2 = bool(gcn), bool(agg_gcn1), bool(agg_gcn2)
        self.act1, self.act2 = act1, act2
        self.uid1, self.uid2 = uid1, uid2
        self.use_graphs = use_graphs
        self.B = int(max_batch)
        dev = self.dev
        B = self.B
        self.w2_width = (k2 if k2 is not None else graph2.max_degree) + (1 if agg_gcn2 else 0)
        self.w1_width = (k1 if k1 is not None else graph1.max_degree) + (1 if agg_gcn1 else 0)
        self.w2_width, self.w1_width = max(self.w2_width, 1), max(self.w1_width, 1)
        self.slot_base_of = (lambda b: 0) if self.gcn else (lambda b: b)
        n1_max = min(B * self.w2_width, graph2.num_nodes) + (0 if self.gcn else B)
        self.n1_max = n1_max
        self.K1 = self.F if self.gcn else 2 * self.F
        self.K2 = self.d1 if self.gcn else 2 * self.d1
        # layer 1 runs on the tcgen05 path when the shape qualifies (d1 == 128, K1 >= 32)
        self.tc1 = ops.encoder_tc_supported(self.K1, self.d1)
        # ... and, in SAGE mode on a local table, consumes the concat in place: only the neighbour-mean half of the
        # layer-1 tile is written, the self rows `self.features(nodes)` (encoders.py:53) are gathered from the feature
        # table inside the GEMMs (gs_sage_encoder_*_tc).  -122 MB of HBM traffic per Reddit-shape step, gather 0.183 ->
        # 0.135 ms, step 0.244 -> 0.230 ms (profiles/README.md R2.3).  GSAGE_SPLIT_SELF=0 keeps the [self | mean] tile.
        self.split_self = (self.tc1 and not self.gcn and self.table_peer is None and
                           ops.encoder_tc_supported(self.F, self.d1) and
                           bool(int(__import__('os').environ.get('GSAGE_SPLIT_SELF', '1'))))
        self.sets = [_FrontierSet(self)]           # sets 2 and 3 are created on first pipelined use
        self.cur = 0
        self.loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
        self.self2 = torch.arange(B, device=dev, dtype=torch.int32)
        self.scratch = ops.DedupScratch(graph2.num_nodes, dev)
        # activations (row-major, ld % 4 == 0)
        self.h1 = ops.empty_rows(n1_max, self.d1, dev, zero=True)
        self.comb2 = ops.empty_rows(B, self.K2, dev, zero=True)
        self.h2 = ops.empty_rows(B, self.d2, dev, zero=True)
        self.logits = ops.empty_rows(B, self.C, dev, zero=True)
        self.loss = torch.zeros(1, device=dev)
        # gradients of activations
        self.gh2 = ops.empty_rows(B, self.d2, dev, zero=True)
        self.gcomb2 = ops.empty_rows(B, self.K2, dev, zero=True)
        self.gh1 = ops.empty_rows(n1_max, self.d1, dev, zero=True)
        self.dz2 = torch.empty((B, ops.round4(self.d2)), device=dev)
        self.dz1 = torch.empty((n1_max, ops.round4(self.d1)), device=dev)
        ws = max(ops.encoder_bwd_ws_floats(n1_max, self.K1, self.d1),
                 ops.encoder_bwd_ws_floats(B, self.K2, self.d2), 4)
        self.ws = torch.empty(ws, device=dev)
        self.xent_ws = torch.empty(ops.classifier_ws_floats(B, self.d2, self.C), device=dev)
        self.tc2 = ops.encoder_tc_supported(self.K2, self.d2) and B >= 512
        need = [4]
        if self.tc1:
            need += [ops.encoder_fwd_tc_ws_floats(self.K1, self.d1), ops.encoder_wgrad_tc_ws_floats(n1_max, self.K1, self.d1)]
        if self.tc2:
            need += [ops.encoder_fwd_tc_ws_floats(self.K2, self.d2), ops.encoder_wgrad_tc_ws_floats(B, self.K2, self.d2)]
        self.tc_ws = torch.empty(max(need), device=dev)
        # the outer layer + classifier + loss + backward as two launches when the shape qualifies
        self.head = ops.head_supported(self.d1, self.K2, self.d2, self.C)
        self.head_ws = ops.head_ws(B, self.K2, self.C, dev) if self.head else None
        self._aux = None
        # parameters + gradients: one flat block each (padded rows), module params alias into it
        shapes = [(self.d1, self.K1), (self.d2, self.K2), (self.C, self.d2)]
        if self.trainable_table:
            shapes.append((embed.shape[0], self.F))
        sizes = [r * ops.round4(c) for r, c in shapes]
        self.flat_w = torch.zeros(sum(sizes), device=dev)
        self.flat_g = torch.zeros(sum(sizes), device=dev)
        views_w, views_g, off = [], [], 0
        for (r, c), sz in zip(shapes, sizes):
            views_w.append(self.flat_w[off:off + sz].view(r, ops.round4(c))[:, :c])
            views_g.append(self.flat_g[off:off + sz].view(r, ops.round4(c))[:, :c])
            off += sz
        self.w1, self.w2, self.wc = views_w[:3]
        self.gw1, self.gw2, self.gwc = views_g[:3]
        if self.trainable_table:
            self.embed, self.gembed = views_w[3], views_g[3]
            self.table = self.embed                      # what the layer-1 gather reads
            self.gcomb1 = ops.empty_rows(n1_max, self.K1, dev, zero=True)
        self._graphs = {}
        self._warm = set()
        self._launch_count = {}
        self._side = self._gstream = self._sstream = self._cstream = None
        self._loss_ring = None
        self._pending_lr = None
        self.queue = []            # batches in flight: {slot, b, state (0 staged, 1 sampled, 2 gathered), ids}
        self.grad_scale = 1.0      # data parallel: n_local * world / n_global (dist.local_grad_scale)

    # convenience views of the current frontier set (tests / bench read these)
    @property
    def n1_dev(self):
        return self.sets[self.cur].n1_dev

    @property
    def cnt1(self):
        return self.sets[self.cur].cnt1

    @property
    def cnt2(self):
        return self.sets[self.cur].cnt2

    # ------------------------------------------------------------------ the launch sequence
    def _n1_max(self, b):
        return min(b * self.w2_width, self.g2.num_nodes) + self.slot_base_of(b)

    def _gather_chain(self, fs, b):
        """sample -> dedup -> sample -> gather for the ``b`` staged targets of frontier set ``fs``:
        everything of a step that does not depend on the weights."""
        self._sample_chain(fs, b)
        self._gather(fs, b)

    def _sample_chain(self, fs, b):
        """Stage 1 of the pipeline: both sampled tiles and the layer-1 frontier of a staged batch
        (latency-bound integer work, no feature bytes)."""
        seed = sampling.get_seed()
        base = self.slot_base_of(b)
        targets = fs.targets[:b]
        idx2, cnt2 = fs.idx2[:b], fs.cnt2[:b]
        # layer-2 tile over the targets (aggregators.py:42-48; RNG draw of agg2)
        self._sample(self.g2, targets, self.k2, add_self=self.agg_gcn2,
                       seed=seed, step_dev=fs.step_dev, tag_head=sampling.call_tag(self.uid2, 0),
                       width=self.w2_width, idx=idx2, cnt=cnt2)
        # frontier of layer 1: [targets (SAGE self pass) | distinct hop-1 ids] (aggregators.py:52-56)
        if not self.gcn:
            fs.frontier1[:b].copy_(targets)
        ops.dedup_remap(idx2, cnt2, self.scratch, slot_base=base, uniq=fs.frontier1[base:], n_total=fs.n1_dev)
        n1_max = self._n1_max(b)
        fr = fs.frontier1[:n1_max]
        idx1, cnt1 = fs.idx1[:n1_max], fs.cnt1[:n1_max]
        # layer-1 tiles: rows < b are the self pass over the batch nodes (independent draw,
        # call index 1), the rest the hop-1 pass (call index 0)  -- SURVEY.md s3.2
        self._sample(self.g1, fr, self.k1, add_self=self.agg_gcn1,
                       seed=seed, step_dev=fs.step_dev, tag_head=sampling.call_tag(self.uid1, 1),
                       tag_tail=sampling.call_tag(self.uid1, 0), n_head=base, n_dev=fs.n1_dev,
                       width=self.w1_width, idx=idx1, cnt=cnt1)
        if self.trainable_table:          # node ids -> rows of the trainable table (aggregators.py:68-71)
            ops.remap_ids(idx1, cnt1, self.hot_map, n_dev=fs.n1_dev)

    @staticmethod
    def _sample(g, nodes, k, **kw):
        """Local CSR, or a sharded.ShardedCSR(peer=True): adjacency rows of other ranks read over NVLink."""
        if getattr(g, "rowptr_ptrs", None) is not None:
            return ops.sample_csr_peer(g.rowptr_ptrs, g.col_ptrs, g.ex.world, g.num_nodes, nodes, k, **kw)
        return ops.sample_csr(g.rowptr, g.col, g.num_nodes, nodes, k, **kw)

    def _gather(self, fs, b):
        """Stage 2: the HBM-bound (partitioned: NVLink-bound) layer-1 gather-mean over the sampled tile -> comb1."""
        n1_max = self._n1_max(b)
        plain = self.gcn or self.split_self                # no self half in the tile this kernel writes
        kw = dict(neigh_off=0 if plain else self.F, self_ids=None if plain else fs.frontier1[:n1_max],
                  n_dev=fs.n1_dev)
        if self.table_peer is not None:
            tp = self.table_peer
            ops.gather_mean_fwd_peer(tp.table_ptrs, tp.ex.world, tp.ld, self.F, fs.idx1[:n1_max], fs.cnt1[:n1_max],
                                     fs.comb1[:n1_max], **kw)
        else:
            ops.gather_mean_fwd(self.table, self.F, fs.idx1[:n1_max], fs.cnt1[:n1_max], fs.comb1[:n1_max], **kw)

    def _compute_chain(self, fs, b):
        """Encoder GEMMs, classifier/loss and the whole backward for frontier set ``fs``."""
        n1_max = self._n1_max(b)
        labels = fs.labels[:b]
        idx2, cnt2 = fs.idx2[:b], fs.cnt2[:b]
        comb1, h1 = fs.comb1[:n1_max], self.h1[:n1_max]
        gh1 = self.gh1[:n1_max]
        if self.head:
            # gh1 is cleared on a forked branch, concurrently with the layer-1 GEMM
            main = torch.cuda.current_stream()
            if self._aux is None:
                self._aux = torch.cuda.Stream(device=self.dev)
            self._aux.wait_stream(main)
            with torch.cuda.stream(self._aux):
                gh1.zero_()
        if self.split_self:
            ops.sage_encoder_fwd_tc(self.table, fs.frontier1[:n1_max], self.F, comb1, self.w1, self.act1, h1,
                                    ws=self.tc_ws, n_dev=fs.n1_dev)
        elif self.tc1:
            ops.encoder_fwd_tc(comb1, self.w1, self.act1, h1, ws=self.tc_ws, n_dev=fs.n1_dev)
        else:
            ops.encoder_fwd(comb1, self.w1, self.act1, h1, n_dev=fs.n1_dev)
        comb2, h2 = self.comb2[:b], self.h2[:b]
        if self.head:
            main.wait_stream(self._aux)
            ops.head_rows(self.h1, self.d1, idx2, cnt2, None if self.gcn else self.self2[:b], self.w2, self.act2,
                          self.wc, labels, self.grad_scale, comb2, h2, self.logits[:b], self.gh1, self.head_ws)
            # the head's weight gradients only feed the update: forked, concurrent with the layer-1 backward
            self._aux.wait_stream(main)
            with torch.cuda.stream(self._aux):
                ops.head_wgrad(comb2, h2, self.d1, self.C, not self.gcn, self.loss, self.gw2, self.gwc, self.head_ws)
            self._wgrad1(fs, comb1, h1, gh1)
            main.wait_stream(self._aux)
            return
        ops.gather_mean_fwd(self.h1, self.d1, idx2, cnt2, comb2, neigh_off=0 if self.gcn else self.d1,
                            self_ids=None if self.gcn else self.self2[:b])
        if self.tc2:
            ops.encoder_fwd_tc(comb2, self.w2, self.act2, h2, ws=self.tc_ws)
        else:
            ops.encoder_fwd(comb2, self.w2, self.act2, h2)
        ops.classifier_xent(h2, self.wc, labels, self.grad_scale, self.logits[:b], self.loss, self.gh2[:b], self.gwc,
                            ws=self.xent_ws)
        if self.tc2:
            ops.encoder_wgrad_tc(comb2, h2, self.gh2[:b], self.act2, self.gw2, ws=self.tc_ws)
            ops.encoder_dgrad(self.w2, h2, self.gh2[:b], self.act2, self.gcomb2[:b], dz=self.dz2)
        else:
            ops.encoder_bwd(comb2, self.w2, h2, self.gh2[:b], self.act2, self.gw2, self.gcomb2[:b], dz=self.dz2,
                            ws=self.ws)
        gh1.zero_()
        ops.scatter_mean_bwd(self.gcomb2[:b], self.d1, idx2, cnt2, self.gh1, neigh_off=0 if self.gcn else self.d1,
                             self_ids=None if self.gcn else self.self2[:b])
        self._wgrad1(fs, comb1, h1, gh1)

    def _wgrad1(self, fs, comb1, h1, gh1):
        if self.trainable_table:
            # gradient of the trainable table: d comb1 = dz1 . W1, then the scatter-add of the mean (K4) into the dense
            # gradient block EmbeddingDenseBackward would produce (model.py:249)
            n1_max = comb1.shape[0]
            ops.encoder_dgrad(self.w1, h1, gh1, self.act1, self.gcomb1[:n1_max], dz=self.dz1, n_dev=fs.n1_dev)
            self.gembed.zero_()
            ops.scatter_mean_bwd(self.gcomb1[:n1_max], self.F, fs.idx1[:n1_max], fs.cnt1[:n1_max], self.gembed,
                                 neigh_off=0, n_dev=fs.n1_dev)
        if self.split_self:
            ops.sage_encoder_wgrad_tc(self.table, fs.frontier1[:comb1.shape[0]], self.F, comb1, h1, gh1, self.act1,
                                      self.gw1, ws=self.tc_ws, n_dev=fs.n1_dev)
        elif self.tc1:
            ops.encoder_wgrad_tc(comb1, h1, gh1, self.act1, self.gw1, ws=self.tc_ws, n_dev=fs.n1_dev)
        else:
            ops.encoder_bwd(comb1, self.w1, h1, gh1, self.act1, self.gw1, None, dz=self.dz1, ws=self.ws,
                            n_dev=fs.n1_dev)

    def _forward_backward(self, b):
        """Enqueue one fwd+bwd over the first ``b`` staged targets on the current stream."""
        fs = self.sets[self.cur]
        self._gather_chain(fs, b)
        self._compute_chain(fs, b)

    # streaming mode defers a step's SGD (or all-reduce + SGD) to the head of the next step's graph: measured
    # 0.279 vs 0.297 ms/step on 1 GPU, 0.284 vs 0.371 on 2 (profiles/README.md); GSAGE_DEFER_UPDATE=0 disables
    defer_update = bool(int(__import__('os').environ.get('GSAGE_DEFER_UPDATE', '1')))
    peer = None        # dist.PeerAllreduceSGD: data parallel with the all-reduce fused into the update kernel

    def _update(self, lr):
        if self.peer is not None:
            self.peer.step(self.flat_w, self.flat_g, lr)
        else:
            ops.sgd_step(self.flat_w, self.flat_g, lr)

    def _overlapped(self, p, b0, b1, b2, lr, lead_lr=None, stage=None):
        """One pipelined step as a fork/join over three streams (captured as ONE CUDA graph):
        compute chain of set p on the current stream || gather of set p+1 || sample chain of set
        p+2.  b1 / b2 are None when the queue is shorter (tail of a run).  ``lead_lr``: apply the
        PREVIOUS step's deferred update first (data parallel, see step_pipelined); ``lr=None``: leave
        this step's update to the next call.  ``stage``: callable(frontier set) that stages the inputs of the batch
        sampled in this step from inside the graph (run_device_queue / run_host_queue)."""
        main = torch.cuda.current_stream()
        if b1 is not None:
            self._gstream.wait_stream(main)
            with torch.cuda.stream(self._gstream):
                self._gather(self.sets[(p + 1) % self.slots], b1)
        if b2 is not None:
            self._sstream.wait_stream(main)
            with torch.cuda.stream(self._sstream):
                if stage is not None:
                    stage(self.sets[(p + 2) % self.slots])
                self._sample_chain(self.sets[(p + 2) % self.slots], b2)
        if lead_lr is not None:
            self._update(lead_lr)
        self._compute_chain(self.sets[p], b0)
        if lr is not None:
            self._update(lr)
        if b1 is not None:
            main.wait_stream(self._gstream)
        if b2 is not None:
            main.wait_stream(self._sstream)

    # ------------------------------------------------------------------ graph capture / replay
    def _run(self, key, fn):
        if not self.use_graphs:
            fn()
            return
        key = key + (sampling.get_seed(),)         # the sampler seed is an immediate kernel argument baked into a capture
        g = self._graphs.get(key)
        if g is None:
            if key not in self._warm:          # first call eager: loads kernels, sets func attributes
                before = ops.LAUNCHES[0]
                fn()
                self._launch_count[key] = ops.LAUNCHES[0] - before
                self._warm.add(key)
                return
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = g
        g.replay()

    @property
    def launches_per_step(self):
        """Kernels of libgsage_sm100.so in one train step (fwd+bwd+SGD), counted when enqueued."""
        c = self._launch_count
        whole = [v for k, v in c.items() if k[0] in ("step", "pipe")] + \
            [v // k[3] for k, v in c.items() if k[0] in ("multi", "hostq")]
        if whole:
            return max(whole)
        parts = sum(max([v for k, v in c.items() if k[0] == nm] or [0]) for nm in ("s", "g", "cchain"))
        return max(max([v for k, v in c.items() if k[0] == "fb"] or [0]), parts) + \
            max([v for k, v in c.items() if k[0] == "sgd"] or [0])

    def stage(self, nodes, labels, step, slot=None):
        """Host -> device copy of one minibatch's inputs (ids, labels, sampler step) into frontier
        set ``slot`` (default: the current one)."""
        b = len(nodes)
        if b > self.B:
            raise ValueError("batch of %d exceeds the engine's max_batch %d" % (b, self.B))
        fs = self.sets[self.cur if slot is None else slot]
        h = fs.stage_host
        if fs.stage_event is not None:
            fs.stage_event.synchronize()          # the previous H2D copy out of this pinned block is done
        h[:8].view(torch.int64)[0] = int(step)
        if isinstance(labels, torch.Tensor):
            h[8:8 + 8 * b].view(torch.int64).copy_(labels.reshape(-1))
        else:
            h[8:8 + 8 * b].view(torch.int64).copy_(torch.from_numpy(np.asarray(labels, dtype=np.int64).reshape(-1)))
        if isinstance(nodes, torch.Tensor):
            h[8 + 8 * self.B:8 + 8 * self.B + 4 * b].view(torch.int32).copy_(nodes.reshape(-1))
        else:
            h[8 + 8 * self.B:8 + 8 * self.B + 4 * b].view(torch.int32).copy_(
                torch.from_numpy(np.asarray(nodes, dtype=np.int32)))
        fs.stage_dev.copy_(h, non_blocking=True)
        if fs.stage_event is None:
            fs.stage_event = torch.cuda.Event()
        fs.stage_event.record()
        return b

    def stage_device(self, d_nodes, d_labels, step, slot=None):
        """Same as ``stage`` for inputs that already live in HBM (device-to-device copies)."""
        fs = self.sets[self.cur if slot is None else slot]
        b = d_nodes.shape[0]
        fs.targets[:b].copy_(d_nodes)
        fs.labels[:b].copy_(d_labels)
        fs.step_dev.fill_(int(step))
        return b

    def pack_stage(self, nodes, labels, step):
        """Host-side: one staging block [step | labels | targets] (uint8 tensor) for ``stage_packed``."""
        blk = torch.zeros(8 + 12 * self.B, dtype=torch.uint8)
        b = len(nodes)
        blk[:8].view(torch.int64)[0] = int(step)
        blk[8:8 + 8 * b].view(torch.int64).copy_(torch.as_tensor(np.asarray(labels, dtype=np.int64).reshape(-1)))
        blk[8 + 8 * self.B:8 + 8 * self.B + 4 * b].view(torch.int32).copy_(torch.as_tensor(np.asarray(nodes, dtype=np.int32)))
        return blk

    def stage_packed(self, block, b, slot=None):
        """Stage a pre-packed block (``pack_stage``; device- or pinned-host-resident) with ONE copy."""
        fs = self.sets[self.cur if slot is None else slot]
        fs.stage_dev.copy_(block, non_blocking=True)
        return b

    def forward_backward(self, b):
        self.flush_update()
        self._run(("fb", b, self.cur), lambda: self._forward_backward(b))

    def update(self, lr):
        self._run(("sgd", float(lr)), lambda: self._update(lr))

    def train_step(self, b, lr, allreduce=None):
        """fwd + bwd (+ gradient all-reduce) + SGD on the staged batch; returns nothing (the
        loss stays in ``self.loss`` on the device)."""
        self.flush_update()
        if allreduce is None:
            self._run(("step", b, float(lr), self.cur), lambda: (self._forward_backward(b), self._update(lr)))
        else:
            self.forward_backward(b)
            allreduce(self.flat_g)
            self.update(lr)

    # ---- software pipelining across minibatches ---------------------------------------------
    # Three stages, one frontier set each, all in flight at once:
    #   S  sample chain of batch t+2   (integer work, latency-bound; high-priority stream)
    #   G  layer-1 gather of batch t+1 (the HBM-bound kernel: gathers run back to back)
    #   C  GEMMs / loss / backward / SGD of batch t  (tensor pipe + L2)
    # so the HBM pipe never idles behind the sampler and the tensor pipe never behind the gather.
    depth = 3
    # One frontier set more than stages: the set that batch t+3 is staged into is then not in use by any stage of
    # step t, so its 12 KB staging copy runs on a copy stream DURING step t instead of between two graphs.
    slots = 4

    def enable_pipeline(self):
        while len(self.sets) < self.slots:
            self.sets.append(_FrontierSet(self))
        if self._cstream is None:
            self._cstream = torch.cuda.Stream(device=self.dev, priority=-1)
            self._slot_done = [None] * self.slots
        if self._gstream is None:
            # the side streams outrank the main stream: their short kernels must not queue behind the
            # second wave of a 200+-CTA GEMM grid (measured: a 35 us stall per step, profiles/README.md)
            self._gstream = torch.cuda.Stream(device=self.dev, priority=-1)
            self._sstream = torch.cuda.Stream(device=self.dev, priority=-1)
        self._side = self._gstream

    def flush_update(self):
        """Apply a deferred weight update (data-parallel pipelining defers the all-reduce + SGD of a step
        to the head of the next step's graph); call before reading the weights mid-run."""
        if self._pending_lr is not None:
            lr, self._pending_lr = self._pending_lr, None
            self.update(lr)

    def reset_pipeline(self):
        self.flush_update()
        if self._cstream is not None:                 # staging copies of dropped batches must not outlive the reset
            torch.cuda.current_stream().wait_stream(self._cstream)
            self._slot_done = [None] * self.slots     # next pushes order themselves after everything enqueued so far
        self.queue = []
        self.cur = 0

    def drop_queued(self, keep):
        """Forget the queued batches after the first ``keep`` (a different batch was announced than the one
        that was prefetched); their sets may have been sampled / gathered by the last step, so the next
        staging copy into them waits for all work enqueued so far instead of the per-set event."""
        for e in self.queue[keep:]:
            self._slot_done[e["slot"]] = None
        del self.queue[keep:]

    def push(self, nodes, labels, step, on_device=False, packed=None):
        """Stage the next minibatch (host ids/labels, device tensors with ``on_device``, or a block made
        by ``pack_stage`` with ``packed=(block, b)``) into the next free frontier set and append it to the
        pipeline queue."""
        self.enable_pipeline()
        if len(self.queue) >= self.depth:
            raise RuntimeError("pipeline queue is full (%d batches in flight)" % self.depth)
        slot = (self.cur + len(self.queue)) % self.slots
        # the copy runs on the copy stream, after the last graph that used this set (as its compute set) has finished
        main = torch.cuda.current_stream()
        if self._slot_done[slot] is not None:
            self._cstream.wait_event(self._slot_done[slot])
        else:
            self._cstream.wait_stream(main)
        with torch.cuda.stream(self._cstream):
            if packed is not None:
                b = self.stage_packed(packed[0], packed[1], slot=slot)
            else:
                b = (self.stage_device if on_device else self.stage)(nodes, labels, step, slot=slot)
            staged = torch.cuda.Event()
            staged.record()
        self.queue.append({"slot": slot, "b": b, "state": 0, "staged": staged,
                           "ids": None if (on_device or packed is not None) else np.array(nodes, dtype=np.int64, copy=True)})
        return b

    def step_pipelined(self, lr, allreduce=None):
        """Train on the oldest queued batch while the next one is gathered and the one after is
        sampled.  Stages a batch has not been through yet (start of a run) are caught up first."""
        q = self.queue
        if not q:
            raise RuntimeError("step_pipelined: no batch queued (call push first)")
        p = self.cur
        assert q[0]["slot"] == p
        main_stream = torch.cuda.current_stream()
        for e in q:                                   # inputs staged on the copy stream are visible to this step
            if e["staged"] is not None:
                main_stream.wait_event(e["staged"])
                e["staged"] = None
        if q[0]["state"] < 1:
            self._run(("s", q[0]["b"], p), lambda: self._sample_chain(self.sets[p], q[0]["b"]))
            q[0]["state"] = 1
        if q[0]["state"] < 2:
            self._run(("g", q[0]["b"], p), lambda: self._gather(self.sets[p], q[0]["b"]))
            q[0]["state"] = 2
        if len(q) > 1 and q[1]["state"] < 1:
            s1 = q[1]["slot"]
            self._run(("s", q[1]["b"], s1), lambda: self._sample_chain(self.sets[s1], q[1]["b"]))
            q[1]["state"] = 1
        b0 = q[0]["b"]
        b1 = q[1]["b"] if len(q) > 1 else None
        b2 = q[2]["b"] if len(q) > 2 else None
        if allreduce is None:
            # The update of this step (SGD, or all-reduce + SGD with self.peer) is deferred to the head of
            # the NEXT step's graph: it (and, data parallel, the wait for the slowest rank) then overlaps
            # the next batch's gather instead of holding the fork/join of this graph open.  Every forward
            # sees the same weights as without deferral; the last step of a run (nothing queued behind
            # it) updates immediately, flush_update() does so on demand.
            lead, self._pending_lr = self._pending_lr, None
            defer = (self.peer is not None or self.defer_update) and b1 is not None
            self._run(("pipe", p, b0, b1, b2, float(lr), lead, defer),
                      lambda: self._overlapped(p, b0, b1, b2, None if defer else lr, lead))
            if defer:
                self._pending_lr = float(lr)
        else:
            # data parallel: the NCCL all-reduce is enqueued eagerly (capturing a collective inside a
            # forked graph deadlocked across ranks), so each stage is its own graph on its own stream:
            # gather and sample run on the side streams during compute -> all-reduce -> SGD.
            main = torch.cuda.current_stream()
            if b1 is not None:
                self._gstream.wait_stream(main)
                with torch.cuda.stream(self._gstream):
                    self._run(("g", b1, (p + 1) % self.slots), lambda: self._gather(self.sets[(p + 1) % self.slots], b1))
            if b2 is not None:
                self._sstream.wait_stream(main)
                with torch.cuda.stream(self._sstream):
                    self._run(("s", b2, (p + 2) % self.slots),
                              lambda: self._sample_chain(self.sets[(p + 2) % self.slots], b2))
            self._run(("cchain", b0, p), lambda: self._compute_chain(self.sets[p], b0))
            allreduce(self.flat_g)
            self.update(lr)
            if b1 is not None:
                main.wait_stream(self._gstream)
            if b2 is not None:
                main.wait_stream(self._sstream)
        if b1 is not None:
            q[1]["state"] = 2
        if b2 is not None:
            q[2]["state"] = 1
        q.pop(0)
        done = torch.cuda.Event()
        done.record()                                 # set p may be restaged once this step has finished
        self._slot_done[p] = done
        self.cur = (p + 1) % self.slots

    def run_device_queue(self, pool, cursor, k, lr):
        """``k`` pipelined steps (a multiple of ``slots``) as ONE graph replay, for inputs that already live in HBM:
        ``pool`` [n_blocks, stage bytes] uint8 holds pre-packed staging blocks (``pack_stage``), ``cursor`` (int64
        device scalar) the index of the next block to stage.  Every step is exactly a ``step_pipelined`` step in its
        steady state -- SGD / all-reduce of the previous batch, compute chain of batch t || feature gather of t+1 ||
        [stage + sample chain] of t+2 -- only the per-step host work (a staging copy and a graph launch per step,
        from every rank of a data-parallel job) is paid once per ``k`` steps.  Needs the steady state of the
        streaming mode: two full batches queued (gathered, sampled) and the previous update deferred."""
        q = self.queue
        if k <= 0 or k % self.slots:
            raise ValueError("run_device_queue: k must be a positive multiple of %d" % self.slots)
        if len(q) != 2 or q[0]["state"] != 2 or q[1]["state"] != 1 or q[0]["b"] != q[1]["b"] or self._pending_lr is None:
            raise RuntimeError("run_device_queue needs the pipeline's steady state (run step_pipelined first)")
        if float(lr) != self._pending_lr:
            raise ValueError("run_device_queue: lr differs from the deferred update's lr")
        b, p0 = q[0]["b"], self.cur
        assert q[0]["slot"] == p0 and pool.shape[1] == self.sets[0].stage_dev.numel()
        main_stream = torch.cuda.current_stream()
        for e in q:
            if e["staged"] is not None:
                main_stream.wait_event(e["staged"])
                e["staged"] = None
        if self._cstream is not None:
            main_stream.wait_stream(self._cstream)

        stage_fn = lambda fs: ops.stage_next(pool, cursor, fs.stage_dev)

        def body():
            # one fork/join per step, as in step_pipelined.  (Round 2 also tried a dataflow capture -- each stage waiting
            # only for its true producer, gathers free-running back to back across step boundaries with the four frontier
            # sets as slack: 0.275 instead of 0.243 ms/step on one GPU, 0.300 instead of 0.259 on two; a gather that
            # never pauses slows the compute chain, which is the critical path, more than the slack buys.)
            for j in range(k):
                self._overlapped((p0 + j) % self.slots, b, b, b, None, float(lr), stage=stage_fn)
        self._run(("multi", p0, b, k, float(lr), pool.data_ptr(), cursor.data_ptr()), body)
        self._slot_done = [None] * self.slots         # later pushes order themselves after everything enqueued so far

    def run_host_queue(self, blocks_host, losses_host, lr):
        """The end-to-end twin of ``run_device_queue``: ``k = len(blocks_host)`` pipelined steps as ONE graph replay whose
        inputs come from HOST memory.  ``blocks_host`` [k, stage bytes] uint8 PINNED holds the pre-packed staging blocks
        of the k batches that get staged during these steps (captured host->device copies, one per step, on the sampler
        stream); ``losses_host`` [k] fp32 PINNED receives every step's loss (captured device->host copies).  The caller
        fills ``blocks_host`` before the launch and reads ``losses_host`` after the graph's completion event; both
        buffers are baked into the capture, so callers alternate between a few fixed (blocks, losses) pairs.  Same
        preconditions as ``run_device_queue``."""
        q = self.queue
        k = int(blocks_host.shape[0])
        if k <= 0 or k % self.slots:
            raise ValueError("run_host_queue: the number of steps must be a positive multiple of %d" % self.slots)
        if len(q) != 2 or q[0]["state"] != 2 or q[1]["state"] != 1 or q[0]["b"] != q[1]["b"] or self._pending_lr is None:
            raise RuntimeError("run_host_queue needs the pipeline's steady state (run step_pipelined first)")
        if float(lr) != self._pending_lr:
            raise ValueError("run_host_queue: lr differs from the deferred update's lr")
        if not (blocks_host.is_pinned() and losses_host.is_pinned()):
            raise ValueError("run_host_queue: blocks_host / losses_host must be pinned host tensors")
        b, p0 = q[0]["b"], self.cur
        assert q[0]["slot"] == p0 and blocks_host.shape[1] == self.sets[0].stage_dev.numel()
        main_stream = torch.cuda.current_stream()
        for e in q:
            if e["staged"] is not None:
                main_stream.wait_event(e["staged"])
                e["staged"] = None
        if self._cstream is not None:
            main_stream.wait_stream(self._cstream)

        def body():
            for j in range(k):
                self._overlapped((p0 + j) % self.slots, b, b, b, None, float(lr),
                                 stage=lambda fs, j=j: fs.stage_dev.copy_(blocks_host[j], non_blocking=True))
                losses_host[j:j + 1].copy_(self.loss, non_blocking=True)
        self._run(("hostq", p0, b, k, float(lr), blocks_host.data_ptr(), losses_host.data_ptr()), body)
        self._slot_done = [None] * self.slots

    def read_loss(self):
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host[0])

    def read_loss_async(self):
        """Enqueue the device->host copy of this step's loss and return a handle; ``float(handle)`` waits
        for that copy only, so the host can launch the next step before reading this one's result."""
        if self._loss_ring is None:
            self._loss_ring = [(torch.empty(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            self._loss_slot = 0
        buf, ev = self._loss_ring[self._loss_slot]
        self._loss_slot = (self._loss_slot + 1) % len(self._loss_ring)
        buf.copy_(self.loss, non_blocking=True)
        ev.record()
        return PendingLoss(buf, ev)


class PendingLoss:
    """Loss of a step whose device->host copy is in flight (``train_step(..., sync=False)``)."""

    def __init__(self, buf, event):
        self._buf, self._event, self._value = buf, event, None

    def ready(self):
        return self._value is not None or self._event.query()

    def __float__(self):
        if self._value is None:
            self._event.synchronize()
            self._value = float(self._buf[0])
        return self._value

    item = __float__


class _EngineLoss(torch.autograd.Function):
    """Hands the engine's already-computed gradients to autograd, so that the reference's
    ``loss.backward(); optimizer.step()`` (model.py:249-250) keeps working unchanged."""

    @staticmethod
    def forward(ctx, loss_buf, n_params, *params_and_grads):
        ctx.n_params = n_params
        ctx.save_for_backward(*params_and_grads[n_params:])
        return loss_buf[0].clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.saved_tensors
        return (None, None) + tuple(gr * g for gr in grads) + (None,) * ctx.n_params


def _closure_reaches(fn, target):
    cells = getattr(fn, "__closure__", None) or ()
    for c in cells:
        try:
            if c.cell_contents is target:
                return True
        except ValueError:
            pass
    return False


def engine_for(model, batch):
    """Return a TrainEngine bound to ``model`` (a SupervisedGraphSage) if its module tree is the
    canonical 2-layer wiring over a frozen nn.Embedding table, else None.  The module
    parameters are re-pointed at the engine's flat parameter block (same values)."""
    import torch.nn as nn
    from .aggregators import MeanAggregator, TABLE_INITIALIZERS
    from .encoders import Encoder
    eng = getattr(model, "_engine", None)
    if eng is not None and eng.B >= batch and eng.k1 == model.enc.base_model.num_sample \
            and eng.k2 == model.enc.num_sample:
        return eng
    if eng is not None:
        # a larger batch / different fan-out replaces the engine: apply its deferred update and drop its queue first,
        # so no weight update is lost (the new engine copies the weights out of the old flat block below)
        eng.reset_pipeline()
    enc2 = model.enc
    enc1 = getattr(enc2, "base_model", None)
    if not (isinstance(enc2, Encoder) and isinstance(enc1, Encoder)):
        return None
    agg1, agg2 = enc1.aggregator, enc2.aggregator
    if not (isinstance(agg1, MeanAggregator) and isinstance(agg2, MeanAggregator)):
        return None
    emb = enc1.features
    peer_table = getattr(emb, "table_ptrs", None) is not None            # sharded.ShardedFeatures(peer=True)
    if not ((peer_table or (isinstance(emb, nn.Embedding) and not emb.weight.requires_grad)) and agg1.features is emb):
        return None
    table_init = enc1.initializer in TABLE_INITIALIZERS
    if enc2.initializer in TABLE_INITIALIZERS:
        return None
    if table_init and not (enc1.gcn and not peer_table and hasattr(agg1, "embed")):
        return None                      # SAGE encoders mix raw self rows with trainable-table means: op-by-op path
    from .graph import CSRGraph
    fusable = lambda g: isinstance(g, CSRGraph) or getattr(g, "rowptr_ptrs", None) is not None
    if not (fusable(enc1.graph) and fusable(enc2.graph)):
        return None                      # partitioned graph answered by all-to-all (sharded.ShardedCSR): op-by-op path
    if not (_closure_reaches(enc2.features, enc1) and _closure_reaches(agg2.features, enc1)):
        return None
    if enc1.gcn != enc2.gcn or getattr(enc1, "base_model", None) is not None:
        return None
    act = lambda e: ops.ACT_SIGMOID if e.initializer in _SIGMOID else ops.ACT_RELU
    eng = TrainEngine(enc1.graph, enc2.graph, emb if peer_table else emb.weight.data, enc1.feat_dim, enc1.embed_dim, enc2.embed_dim,
                      model.weight.shape[0], enc1.num_sample, enc2.num_sample, max(batch, 1), gcn=enc1.gcn,
                      agg_gcn1=agg1.gcn, agg_gcn2=agg2.gcn, act1=act(enc1), act2=act(enc2),
                      uid1=agg1.uid, uid2=agg2.uid, embed=agg1.embed.weight.data if table_init else None,
                      hot_map=agg1.hot_map() if table_init else None)
    pairs = [(eng.w1, enc1.weight), (eng.w2, enc2.weight), (eng.wc, model.weight)]
    if table_init:
        pairs.append((eng.embed, agg1.embed.weight))
    with torch.no_grad():
        for view, p in pairs:
            view.copy_(p.data)
            p.data = view
    model._engine = eng
    return eng
