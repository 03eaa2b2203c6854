/*
 * gsage.h -- C ABI of libgsage_sm100.so: the B200 (sm_100a) implementation of GraphSAGE's
 * sample -> aggregate -> update minibatch hot path.
 *
 * Every entry point below replaces one piece of zjzijielu/graphsage-simple's Python hot
 * path (file:line cited per function, paths relative to the reference checkout).  The
 * reference has no FFI of its own (it is pure Python/PyTorch); the binding a maintainer
 * would add is a ctypes stub, shown in INTEGRATION.md and shipped as
 * graphsage-simple_b200/graphsage/_native.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*
 *   - `stream` is a cudaStream_t passed as void*; nothing here allocates, frees,
 *     synchronises or touches the default stream, so every call is CUDA-graph capturable
 *   - return value: 0 = ok, < 0 = argument error (GS_E*), > 0 = cudaError_t of the launch
 *   - row counts come as (n_max, n_dev): n_max sizes the grid and the buffers; if n_dev is
 *     non-NULL the kernels read the actual count (<= n_max) from device memory, so that a
 *     captured graph can process data-dependent frontier sizes without a host round trip
 *   - node ids and tile entries are int32, CSR row pointers int64, features/weights fp32
 *   - all fp32 matrices are row-major with a leading dimension `ld*` in floats; `ld % 4 == 0`
 *     and 16-byte aligned bases are required (128-bit loads)
 *   - a "tile" is the fixed-width sampled neighbourhood: idx[n, width] (unused slots -1)
 *     plus cnt[n]
 */
#ifndef GSAGE_H_
#define GSAGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 1

#if defined(__GNUC__)
#define GS_API __attribute__((visibility("default")))
#else
#define GS_API
#endif

#define GS_OK        0
#define GS_EINVAL   -1   /* bad size / null pointer                       */
#define GS_EALIGN   -2   /* pointer or leading dimension not 16-B aligned */
#define GS_ENOSUP   -3   /* shape outside what the kernel supports        */

#define GS_ACT_NONE    0
#define GS_ACT_RELU    1  /* encoders.py:61 */
#define GS_ACT_SIGMOID 2  /* encoders.py:59 (initializer in node_degree/shared/pagerank) */

GS_API int gs_abi_version(void);
GS_API const char* gs_strerror(int code);

/* ---- K1: neighbour sampling ------------------------------------------------------------
 * Replaces `random.sample(to_neigh, num_sample)` over adj_lists sets,
 * graphsage/aggregators.py:42-48, the adjacency lookup graphsage/encoders.py:47 and the
 * (intended) self-loop union aggregators.py:50-51.
 * For row i (node v = nodes[i]): all neighbours when k < 0 or deg <= k, else a uniform
 * k-subset drawn with Floyd's algorithm from Philox4x32-10(counter = (v, m/4, step, tag),
 * key = seed); positions sorted, so tile entries ascend.  Rows i < n_head use tag_head,
 * the rest tag_tail (the reference's independent aggregator calls of one forward).
 * If step_dev != NULL the step is read from device memory (graph replay), else `step`.
 * An id outside [0, num_nodes) is an isolated node: cnt = 0 (+ the node itself with add_self),
 * as the reference's defaultdict(set) answers an empty set (model.py:303).
 * The exact specification is restated on the CPU in oracle/sampler_port.py.            */
GS_API int gs_sample_csr(const int64_t* rowptr, const int32_t* col, int32_t num_nodes,
                  const int32_t* nodes, int32_t n_max, const int32_t* n_dev,
                  int32_t k, int32_t width, int32_t add_self,
                  uint64_t seed, int64_t step, const int64_t* step_dev,
                  uint32_t tag_head, uint32_t tag_tail, int32_t n_head,
                  int32_t* idx, int32_t* cnt, void* stream);

/* ---- frontier dedup ---------------------------------------------------------------------
 * Replaces `unique_nodes_list = list(set.union(*samp_neighs))` and the id->column dict,
 * aggregators.py:52-56.  Distinct ids of the tile, ascending, are written to uniq[0..U);
 * the tile is rewritten in place as positions slot_base + rank; *n_total_dev = slot_base + U.
 * Scratch: slot_of[max(num_nodes, 64)] ints holding a node BITMAP (one bit per id) and one rank per bitmap word --
 * ZEROED ONCE by the caller; every call leaves it zero again (only the words it touched are cleared, through the ids
 * it found), so per call the work is the frontier's entries + a popcount scan over num_nodes / 32 words, not a clear
 * and a scan of one 4-byte slot per node -- and block_counts[gs_dedup_scratch_ints(num_nodes)] (zeroed once).
 * cnt == NULL: every entry of idx is valid (flat ragged index array: n_max = entries, width = 1). */
GS_API int32_t gs_dedup_scratch_ints(int32_t num_nodes);
GS_API int gs_dedup_remap(int32_t* idx, const int32_t* cnt, int32_t n_max, const int32_t* n_dev,
                   int32_t width, int32_t num_nodes, int32_t* slot_of, int32_t* block_counts,
                   int32_t slot_base, int32_t* uniq, int32_t* n_total_dev, void* stream);

/* ---- K2: gather-mean forward ------------------------------------------------------------
 * Replaces the dense mask build, row-normalisation, feature lookup and `mask.mm(embed)`,
 * aggregators.py:54-65, 74, plus the self lookup + torch.cat of encoders.py:49-54.
 *   out[i, neigh_off : neigh_off+dim] = (1/cnt[i]) * sum_j table[idx[i,j], 0:dim]
 *   out[i, 0:dim]                     = table[self_ids[i], 0:dim]      (if self_ids != NULL)
 * Rows with cnt == 0 produce zeros (the reference produces NaN from 0/0).               */
GS_API int gs_gather_mean_fwd(const float* table, int64_t ld_table, int32_t dim,
                       const int32_t* idx, const int32_t* cnt, int32_t width,
                       const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                       float* out, int64_t ld_out, int32_t neigh_off, void* stream);

/* ---- K4: scatter-add backward of the mean ------------------------------------------------
 * Replaces autograd's MmBackward of `mask.mm(embed_matrix)` (mask^T . g) and the
 * CatBackward split, reached from loss.backward() at graphsage/model.py:249.
 *   gtable[idx[i,j], 0:dim] += gout[i, neigh_off : neigh_off+dim] / cnt[i]
 *   gtable[self_ids[i], 0:dim] += gout[i, 0:dim]                      (if self_ids != NULL)
 * Accumulates with vectorised fp32 atomics into a caller-zeroed gtable.                 */
GS_API int gs_scatter_mean_bwd(const float* gout, int64_t ld_gout, int32_t neigh_off, int32_t dim,
                        const int32_t* idx, const int32_t* cnt, int32_t width,
                        const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                        float* gtable, int64_t ld_gtable, void* stream);

/* ---- full neighbourhood (num_sample=None) over ragged tiles ----------------------------------------
 * Replaces the un-sampled path of MeanAggregator.forward (aggregators.py:47-48 with the lookup of
 * encoders.py:47 and the self-loop union of aggregators.py:50-51) and its mean / backward
 * (aggregators.py:54-74, model.py:249) without a tile as wide as the largest degree:
 *   gs_take_all_count: len[i] = deg(nodes[i]) (+1 if add_self and the node is not its own neighbour),
 *                      off[0..n] = exclusive prefix sums (off[n] = total entries)
 *   gs_take_all_fill:  flat[off[i] .. off[i+1]) = the row (ascending) [+ the node itself]
 *   gs_gather_mean_ragged:  out[i, :] = mean_e table[flat[e], :]        (zeros for empty rows)
 *   gs_scatter_mean_ragged: gtable[flat[e], :] += gout[i, :] / len(i)   (caller-zeroed gtable)        */
GS_API int gs_take_all_count(const int64_t* rowptr, const int32_t* col, int32_t num_nodes, const int32_t* nodes,
                      int32_t n, int32_t add_self, int32_t* len, int32_t* off, void* stream);
GS_API int gs_take_all_fill(const int64_t* rowptr, const int32_t* col, int32_t num_nodes, const int32_t* nodes,
                     int32_t n, const int32_t* off, int32_t* flat, void* stream);
GS_API int gs_gather_mean_ragged(const float* table, int64_t ld_table, int32_t dim, const int32_t* off,
                          const int32_t* flat, int32_t n, float* out, int64_t ld_out, void* stream);
GS_API int gs_scatter_mean_ragged(const float* gout, int64_t ld_gout, int32_t dim, const int32_t* off,
                           const int32_t* flat, int32_t n, float* gtable, int64_t ld_gtable, void* stream);

/* ---- K3: encoder GEMM + activation --------------------------------------------------------
 * Replaces `F.relu(self.weight.mm(combined.t()))` / sigmoid, encoders.py:58-61.
 *   h[i, 0:d_out] = act( sum_k x[i,k] * w[o,k] ),  x = combined [n, k_in], w [d_out, k_in].
 * Output is row-major [n, d_out]; the reference's [d_out, n] is its transposed view.    */
GS_API int gs_encoder_fwd(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                   int32_t k_in, int32_t d_out, int32_t act,
                   int32_t n_max, const int32_t* n_dev,
                   float* h, int64_t ld_h, void* stream);

/* Backward of the above (autograd ThresholdBackward/SigmoidBackward + MmBackward):
 *   dz = gh * act'(h)                      (written to dz [n_max, round_up(d_out,4)])      
 *   gw[d_out, k_in]  = dz^T . x            (overwritten; split-K partials in ws)
 *   gx[n, k_in]      = dz . w              (only if gx != NULL)
 * ws must hold gs_encoder_bwd_ws_floats(n_max, k_in, d_out) floats.                      */
GS_API int64_t gs_encoder_bwd_ws_floats(int32_t n_max, int32_t k_in, int32_t d_out);
GS_API int gs_encoder_bwd(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                   const float* h, int64_t ld_h, const float* gh, int64_t ld_gh,
                   int32_t k_in, int32_t d_out, int32_t act,
                   int32_t n_max, const int32_t* n_dev,
                   float* dz, float* gw, int64_t ld_gw, float* gx, int64_t ld_gx,
                   float* ws, void* stream);

/* Input gradient alone: dz = gh * act'(h) (dz [n_max, round_up(d_out,4)]), gx[n, k_in] = dz . w.   */
GS_API int gs_encoder_dgrad(const float* w, int64_t ld_w, const float* h, int64_t ld_h,
                     const float* gh, int64_t ld_gh, int32_t k_in, int32_t d_out, int32_t act,
                     int32_t n_max, const int32_t* n_dev, float* dz, float* gx, int64_t ld_gx,
                     void* stream);

/* ---- K3 on tcgen05 tensor cores (3xTF32, fp32 accumulation in TMEM) ------------------------
 * Same contracts as gs_encoder_fwd and the gw part of gs_encoder_bwd, for the shapes
 * gs_encoder_tc_supported() accepts (d_out == 128, k_in >= 32).  The three-term hi/lo
 * TF32 split keeps results within 1e-5 (norm-wise) of the fp32 reference path.
 * ws: gs_encoder_fwd_tc_ws_floats / gs_encoder_wgrad_tc_ws_floats floats, 16-B aligned.    */
GS_API int gs_encoder_tc_supported(int32_t k_in, int32_t d_out);
GS_API int64_t gs_encoder_fwd_tc_ws_floats(int32_t k_in, int32_t d_out);
GS_API int gs_encoder_fwd_tc(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                      int32_t k_in, int32_t d_out, int32_t act,
                      int32_t n_max, const int32_t* n_dev,
                      float* h, int64_t ld_h, float* ws, void* stream);
/* SAGE concat consumed in place (encoders.py:49-61 as ONE op for the rows of the feature table):
 *   h = act([table[self_ids] | mean] . w^T)          w [d_out, 2 * feat_dim] = [W_self | W_neigh]
 * `mean` [n, feat_dim] is the neighbour mean gs_gather_mean_fwd wrote (self_ids == NULL, neigh_off == 0); the self
 * half of the combined tile is NOT materialised: the GEMM's splitter warps gather the rows `self.features(nodes)`
 * (encoders.py:53) straight from the table into the shared-memory stage ring (cp.async into the stage region they
 * just freed), and the tcgen05 MMAs consume them from there.  gs_sage_encoder_wgrad_tc is the matching weight gradient gw [d_out, 2 * feat_dim] = dz^T . [table[self_ids] | mean].
 * ws sizes: gs_encoder_fwd_tc_ws_floats(2 * feat_dim, d_out), gs_encoder_wgrad_tc_ws_floats(n_max, 2 * feat_dim, d_out). */
GS_API int gs_sage_encoder_fwd_tc(const float* table, int64_t ld_table, const int32_t* self_ids, int32_t feat_dim,
                           const float* mean, int64_t ld_mean, const float* w, int64_t ld_w,
                           int32_t d_out, int32_t act, int32_t n_max, const int32_t* n_dev,
                           float* h, int64_t ld_h, float* ws, void* stream);
GS_API int gs_sage_encoder_wgrad_tc(const float* table, int64_t ld_table, const int32_t* self_ids, int32_t feat_dim,
                             const float* mean, int64_t ld_mean, const float* h, int64_t ld_h,
                             const float* gh, int64_t ld_gh, int32_t d_out, int32_t act,
                             int32_t n_max, const int32_t* n_dev, float* gw, int64_t ld_gw, float* ws, void* stream);
GS_API int64_t gs_encoder_wgrad_tc_ws_floats(int32_t n_max, int32_t k_in, int32_t d_out);
GS_API int gs_encoder_wgrad_tc(const float* x, int64_t ld_x, const float* h, int64_t ld_h,
                        const float* gh, int64_t ld_gh, int32_t k_in, int32_t d_out, int32_t act,
                        int32_t n_max, const int32_t* n_dev,
                        float* gw, int64_t ld_gw, float* ws, void* stream);

/* ---- K5: classifier + softmax cross-entropy, forward and backward -------------------------
 * Replaces `scores = self.weight.mm(embeds).t()` and nn.CrossEntropyLoss (mean reduction),
 * graphsage/model.py:57, 62-69, and their autograd backward.
 *   logits[i, c] = sum_d h[i,d] * wc[c,d]            (written if logits != NULL)
 *   loss[0]      = mean_i ( logsumexp(logits[i]) - logits[i, labels[i]] )
 *   gh[i, :]     = grad_scale/n * (softmax(logits[i]) - onehot) . wc      (if gh  != NULL)
 *   gwc[c, :]    = grad_scale/n * sum_i (softmax - onehot)[i,c] * h[i,:]  (if gwc != NULL)
 * ws must hold gs_classifier_ws_floats(n, d, num_classes) floats.                          */
GS_API int64_t gs_classifier_ws_floats(int32_t n, int32_t d, int32_t num_classes);
GS_API int gs_classifier_xent(const float* h, int64_t ld_h, const float* wc, int64_t ld_wc,
                       const int64_t* labels, int32_t d, int32_t num_classes, int32_t n,
                       float grad_scale, float* logits, int64_t ld_logits, float* loss,
                       float* gh, int64_t ld_gh, float* gwc, int64_t ld_gwc,
                       float* ws, void* stream);

/* ---- fused head: outer layer + classifier + loss + backward, two launches ---------------------
 * For the outer (last) layer of the 2-layer model, whose input rows are the previous layer's
 * outputs h1 (positions, not node ids): replaces aggregators.py:54-74 (mean over the sampled
 * tile), encoders.py:49-61 (self | neigh concat, W2 GEMM, activation), model.py:57-69 (classifier,
 * mean cross-entropy) and their autograd backward (model.py:249) for n targets at once:
 *   comb2[i] = [h1[self_slots[i]] | mean_j h1[idx[i,j]]]   (no self half if self_slots == NULL)
 *   h2 = act2(comb2 . w2^T);  logits = h2 . wc^T;  loss[0] = mean CE(logits, labels)
 *   gw2, gwc = weight gradients (overwritten; deterministic two-stage reduction)
 *   gh1 += d loss / d h1   (128-bit reductions into a caller-zeroed [*, d1] buffer)
 * with every gradient scaled by grad_scale.  Supported shapes: gs_head_supported() (d1 == d2 ==
 * 128, num_classes <= 128).  ws: gs_head_ws_floats(n_max, ...) floats, 16-B aligned, ZEROED ONCE by
 * the caller before first use (it holds re-arming tickets, at an offset that does not depend on n):
 * one workspace sized for n_max serves every call with n <= n_max.  logits may be NULL.          */
GS_API int gs_head_supported(int32_t d1, int32_t k2_in, int32_t d2, int32_t num_classes);
GS_API int64_t gs_head_ws_floats(int32_t n, int32_t k2_in, int32_t num_classes);
GS_API int gs_head_fwd_bwd(const float* h1, int64_t ld_h1, int32_t d1,
                    const int32_t* idx, const int32_t* cnt, int32_t width, const int32_t* self_slots,
                    const float* w2, int64_t ld_w2, int32_t d2, int32_t act2,
                    const float* wc, int64_t ld_wc, int32_t num_classes,
                    const int64_t* labels, int32_t n, float grad_scale,
                    float* comb2, int64_t ld_comb2, float* h2, int64_t ld_h2,
                    float* logits, int64_t ld_logits, float* loss,
                    float* gh1, int64_t ld_gh1, float* gw2, int64_t ld_gw2, float* gwc, int64_t ld_gwc,
                    float* ws, void* stream);

/* The two launches of gs_head_fwd_bwd separately, so that a caller can run the weight gradients
 * (gs_head_wgrad: gw2, gwc, loss; reads what gs_head_rows left in comb2, h2 and ws) on another
 * stream, concurrently with the inner layer's backward, which only needs gh1.               */
GS_API int gs_head_rows(const float* h1, int64_t ld_h1, int32_t d1,
                 const int32_t* idx, const int32_t* cnt, int32_t width, const int32_t* self_slots,
                 const float* w2, int64_t ld_w2, int32_t d2, int32_t act2,
                 const float* wc, int64_t ld_wc, int32_t num_classes,
                 const int64_t* labels, int32_t n, float grad_scale,
                 float* comb2, int64_t ld_comb2, float* h2, int64_t ld_h2,
                 float* logits, int64_t ld_logits, float* gh1, int64_t ld_gh1,
                 float* ws, void* stream);
GS_API int gs_head_wgrad(const float* comb2, int64_t ld_comb2, const float* h2, int64_t ld_h2,
                  int32_t d1, int32_t d2, int32_t num_classes, int32_t n, int32_t sage,
                  float* loss, float* gw2, int64_t ld_gw2, float* gwc, int64_t ld_gwc,
                  float* ws, void* stream);

/* ---- K6: SGD -----------------------------------------------------------------------------
 * Replaces torch.optim.SGD(lr=0.7).step(), graphsage/model.py:237, 250: p -= lr * g.      */
GS_API int gs_sgd_step(float* p, const float* g, float lr, int64_t n, void* stream);

/* ---- data parallel: gradient all-reduce fused with SGD over NVLink peer memory -----------------
 * Replaces `optimizer.step()` (model.py:237, 250) when the global batch is split over `world`
 * ranks:  p -= lr * sum_r g_r, the sum taken in rank order (identical bits on every rank).
 * Two phases over peer memory: every rank publishes its gradient, rank r sums slice r of the block over all ranks
 * and writes the sum into every rank's buffer, every rank updates from the broadcast sums -- 2 (world-1)/world of the
 * block crosses NVLink per rank.
 * stage_ptrs / flag_ptrs: DEVICE arrays of `world` pointers, entry q = rank q's staging buffer
 * (4 * n floats: in[2][n] | out[2][n], double-buffered by step parity) / flag pad (2 * world *
 * gs_allreduce_sgd_blocks(n) uint32, zeroed once) as mapped into THIS process (CUDA IPC / symmetric memory).  state: 2 zeroed uint32 in
 * local device memory (epoch, ticket).  n % 4 == 0.  Every rank must launch it once per step;
 * no host or NCCL synchronisation is involved, so the launch can sit in a captured graph.      */
GS_API int32_t gs_allreduce_sgd_blocks(int64_t n);
GS_API int gs_allreduce_sgd(float* p, const float* g, int64_t n, float lr,
                     float* const* stage_ptrs, uint32_t* const* flag_ptrs,
                     int32_t rank, int32_t world, uint32_t* state, void* stream);

/* Row gather without the mean (feature lookup aggregators.py:63-65 as a bit-exact copy):
 * out[i, 0:dim] = table[ids[i], 0:dim].                                                  */
GS_API int gs_gather_rows(const float* table, int64_t ld_table, int32_t dim, const int32_t* ids,
                   int32_t n_max, const int32_t* n_dev, float* out, int64_t ld_out,
                   void* stream);

/* ---- partitioned table / CSR: bucket ids by owner ---------------------------------------------
 * No reference counterpart (the reference is single-process); this is the first step of the
 * exchange that replaces the local lookups `features(LongTensor(unique_nodes_list))`
 * (aggregators.py:62-65) and `adj_lists[int(node)]` (encoders.py:47) when rows are partitioned
 * by owner = id % world.  Stable counting sort of ids[0..n) by owner:
 *   send_ids[pos] = ids[i] (or ids[i] / world if emit_local), perm[i] = pos, counts[o] = |bucket o|
 * with bucket o occupying send_ids[sum(counts[:o]) ...).  scratch: gs_bucket_scratch_ints ints.  */
GS_API int32_t gs_bucket_scratch_ints(int32_t n_max, int32_t world);
GS_API int gs_bucket_by_owner(const int32_t* ids, int32_t n_max, const int32_t* n_dev, int32_t world,
                       int32_t emit_local, int32_t* scratch, int32_t* send_ids, int32_t* perm,
                       int32_t* counts, void* stream);

/* ---- partitioned table / CSR read through peer memory (NVLink) -----------------------------------
 * Same contracts as gs_gather_rows, gs_gather_mean_fwd and gs_sample_csr (k <= 32 or take-all) for
 * a feature table / CSR partitioned by owner = id % world: node v is local row v / world of rank
 * v % world.  tables / rowptrs / cols are DEVICE arrays of `world` pointers, entry q = rank q's shard
 * as mapped into this process (CUDA IPC / symmetric memory); ld_table is common to all shards and
 * rank q's rowptr has (its row count + 1) entries into its own col.  The exchange of the partitioned
 * path is thereby fused into the kernels: remote rows are read over NVLink by the
 * warp that consumes them; no all-to-all, no staging, no host synchronisation.               */
GS_API int gs_gather_rows_peer(const float* const* tables, int32_t world, int64_t ld_table, int32_t dim,
                        const int32_t* ids, int32_t n_max, const int32_t* n_dev, float* out, int64_t ld_out,
                        void* stream);
GS_API int gs_gather_mean_fwd_peer(const float* const* tables, int32_t world, int64_t ld_table, int32_t dim,
                            const int32_t* idx, const int32_t* cnt, int32_t width,
                            const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                            float* out, int64_t ld_out, int32_t neigh_off, void* stream);
GS_API int gs_sample_csr_peer(const int64_t* const* rowptrs, const int32_t* const* cols, int32_t world,
                       int32_t num_nodes, const int32_t* nodes, int32_t n_max, const int32_t* n_dev,
                       int32_t k, int32_t width, int32_t add_self,
                       uint64_t seed, int64_t step, const int64_t* step_dev,
                       uint32_t tag_head, uint32_t tag_tail, int32_t n_head,
                       int32_t* idx, int32_t* cnt, void* stream);

/* idx[i, j] = map[idx[i, j]] for the valid entries (j < cnt[i], or all when cnt == NULL) of the rows below *n_dev.
 * Replaces the per-row Python loop `indices = [np.where(row == 1)[0][0] for row in embed_matrix]`
 * (aggregators.py:68-70) of the 1hot / node_degree initialisers: `map` holds, per node id, the position of the 1 in
 * its one-hot feature row (the node id itself for 1hot, its degree for node_degree), computed once per table, so a
 * sampled tile of node ids becomes a tile of rows of the trainable table `self.embed` (aggregators.py:30-31, 71). */
GS_API int gs_remap_ids(int32_t* idx, const int32_t* cnt, int32_t n_max, const int32_t* n_dev, int32_t width,
                 const int32_t* map, void* stream);

/* Small device-side helpers used to keep a training step free of host round trips.       */
GS_API int gs_advance_step(int64_t* step_dev, void* stream);                    /* ++*step_dev   */
/* Device-side batch queue for inputs that already live in HBM: dst[0..block_bytes) = block (*cursor % n_blocks) of
 * `pool` (n_blocks blocks of block_bytes, a multiple of 4, 4-B aligned), then ++*cursor.  One staging block holds what
 * the reference's loop fetches per minibatch (model.py:243-248: the batch's node ids and labels) plus the sampler step;
 * reading the block index from device memory lets several train steps replay as ONE captured graph.               */
GS_API int gs_stage_next(const void* pool, int64_t block_bytes, int64_t n_blocks, int64_t* cursor, void* dst,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSAGE_H_ */
