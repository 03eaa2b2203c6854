// Full-neighbourhood aggregation over ragged tiles -- the reference's `num_sample=None` mode
// (graphsage/aggregators.py:47-48: "samp_neighs = to_neighs"), used by its validation forward
// (graphsage/model.py:256) when sampling is switched off.  The fixed-width tile of the sampled path would have
// to be as wide as the largest degree (30 000+ on heavy-tailed graphs), so here a frontier's neighbourhoods
// are kept ragged: off[n+1] (exclusive prefix of the row lengths) + one flat index array.
//   gs_take_all_csr        adjacency lookup encoders.py:47 + self-loop union aggregators.py:50-51
//   gs_gather_mean_ragged  mask build / normalise / mask.mm, aggregators.py:54-61, 74
//   gs_scatter_mean_ragged its autograd backward (model.py:249)
// Dedup of the flat array (aggregators.py:52-56) is gs_dedup_remap with width = 1 and cnt = NULL.
#include "gs_common.cuh"

namespace {

constexpr int kWarps = 8;

// pass 1: row lengths (degree, +1 if the node itself has to be appended)
__global__ void __launch_bounds__(kWarps * 32)
ragged_count_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int num_nodes,
                    const int32_t* __restrict__ nodes, int n, int add_self, int32_t* __restrict__ len) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const int32_t v = nodes[i];
    const bool known = (unsigned)v < (unsigned)num_nodes;      // unknown id = isolated node (empty set in the reference)
    const int64_t base = known ? rowptr[v] : 0;
    const int deg = known ? (int)(rowptr[v + 1] - base) : 0;
    bool has = false;
    if (add_self)
        for (int j = lane; j < deg; j += 32) has |= (col[base + j] == v);
    has = __any_sync(0xffffffffu, has);
    if (lane == 0) len[i] = deg + ((add_self && !has) ? 1 : 0);
}

// pass 2: off = exclusive scan of len (one block; n is a frontier size)
__global__ void __launch_bounds__(1024)
ragged_scan_kernel(const int32_t* __restrict__ len, int n, int32_t* __restrict__ off) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int v = c < n ? len[c] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            int s = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            s_warp[lane] = s;
        }
        __syncthreads();
        const int carry = s_carry;
        if (c < n) off[c] = carry + (w ? s_warp[w - 1] : 0) + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = s_carry;
}

// pass 3: copy the rows (ascending, as stored) and append the node itself where needed
__global__ void __launch_bounds__(kWarps * 32)
ragged_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int num_nodes,
                   const int32_t* __restrict__ nodes, int n, const int32_t* __restrict__ off,
                   int32_t* __restrict__ flat) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const int32_t v = nodes[i];
    const bool known = (unsigned)v < (unsigned)num_nodes;
    const int64_t base = known ? rowptr[v] : 0;
    const int deg = known ? (int)(rowptr[v + 1] - base) : 0;
    const int o = off[i], len = off[i + 1] - o;
    for (int j = lane; j < deg; j += 32) flat[o + j] = col[base + j];
    if (lane == 0 && len > deg) flat[o + deg] = v;
}

__device__ __forceinline__ float4 ld_chunk(const float* __restrict__ row, int c4, int dim) {
    const int col = c4 * 4;
    if (col + 4 <= dim) return gs_ldg_stream(reinterpret_cast<const float4*>(row) + c4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < dim) v.x = __ldg(row + col);
    if (col + 1 < dim) v.y = __ldg(row + col + 1);
    if (col + 2 < dim) v.z = __ldg(row + col + 2);
    return v;
}

// out[i, :] = mean over flat[off[i] .. off[i+1]) of table rows; warp per row, lane owns float4 chunks l, l+32, ...
__global__ void __launch_bounds__(kWarps * 32)
gather_mean_ragged_kernel(const float* __restrict__ table, int64_t ld_table, int dim,
                          const int32_t* __restrict__ off, const int32_t* __restrict__ flat, int n,
                          float* __restrict__ out, int64_t ld_out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int o = off[row], c = off[row + 1] - o;
    const float inv = c > 0 ? 1.f / (float)c : 0.f;
    const int nchunks = (dim + 3) >> 2;
    const int32_t* irow = flat + o;
    float* orow = out + (int64_t)row * ld_out;
    for (int c4 = lane; c4 < nchunks; c4 += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = 0;
        for (; j + 4 <= c; j += 4) {                      // 4 independent row loads in flight per lane
            float4 v[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] = ld_chunk(table + (int64_t)irow[j + b] * ld_table, c4, dim);
#pragma unroll
            for (int b = 0; b < 4; ++b) { acc.x += v[b].x; acc.y += v[b].y; acc.z += v[b].z; acc.w += v[b].w; }
        }
        for (; j < c; ++j) {
            const float4 v = ld_chunk(table + (int64_t)irow[j] * ld_table, c4, dim);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const int col = c4 * 4;
        if (col + 4 <= dim) *reinterpret_cast<float4*>(orow + col) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        else {
            if (col < dim) orow[col] = acc.x * inv;
            if (col + 1 < dim) orow[col + 1] = acc.y * inv;
            if (col + 2 < dim) orow[col + 2] = acc.z * inv;
        }
    }
}

// gtable[flat[e], :] += gout[i, :] / len(i) for e in row i
__global__ void __launch_bounds__(kWarps * 32)
scatter_mean_ragged_kernel(const float* __restrict__ gout, int64_t ld_gout, int dim,
                           const int32_t* __restrict__ off, const int32_t* __restrict__ flat, int n,
                           float* __restrict__ gtable, int64_t ld_gtable) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    const int o = off[row], c = off[row + 1] - o;
    const float inv = c > 0 ? 1.f / (float)c : 0.f;
    const int nchunks = (dim + 3) >> 2;
    const int32_t* irow = flat + o;
    const float* grow = gout + (int64_t)row * ld_gout;
    for (int c4 = lane; c4 < nchunks; c4 += 32) {
        const int col = c4 * 4;
        const bool full = col + 4 <= dim;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full) g = *reinterpret_cast<const float4*>(grow + col);
        else { if (col < dim) g.x = grow[col]; if (col + 1 < dim) g.y = grow[col + 1]; if (col + 2 < dim) g.z = grow[col + 2]; }
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
        for (int j = 0; j < c; ++j) {
            float* dst = gtable + (int64_t)irow[j] * ld_gtable + col;
            if (full) gs_red_add_v4(dst, g);
            else { if (col < dim) atomicAdd(dst, g.x); if (col + 1 < dim) atomicAdd(dst + 1, g.y);
                   if (col + 2 < dim) atomicAdd(dst + 2, g.z); }
        }
    }
}

}  // namespace

extern "C" int gs_take_all_count(const int64_t* rowptr, const int32_t* col, int32_t num_nodes, const int32_t* nodes,
                                 int32_t n, int32_t add_self, int32_t* len, int32_t* off, void* stream) {
    if (n < 0 || !off || num_nodes < 0) return GS_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        if (!rowptr || !col || !nodes || !len) return GS_EINVAL;
        ragged_count_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(rowptr, col, num_nodes, nodes, n, add_self, len);
        GS_LAUNCH_CHECK();
    }
    ragged_scan_kernel<<<1, 1024, 0, s>>>(len, n, off);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_take_all_fill(const int64_t* rowptr, const int32_t* col, int32_t num_nodes, const int32_t* nodes,
                                int32_t n, const int32_t* off, int32_t* flat, void* stream) {
    if (n == 0) return GS_OK;
    if (!rowptr || !col || !nodes || !off || !flat || n < 0 || num_nodes < 0) return GS_EINVAL;
    ragged_fill_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(rowptr, col, num_nodes, nodes, n, off, flat);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_gather_mean_ragged(const float* table, int64_t ld_table, int32_t dim, const int32_t* off,
                                     const int32_t* flat, int32_t n, float* out, int64_t ld_out, void* stream) {
    if (n == 0) return GS_OK;
    if (!table || !off || !flat || !out || dim <= 0 || n < 0) return GS_EINVAL;
    if (!gs_aligned16(table) || !gs_aligned16(out) || (ld_table & 3) || (ld_out & 3)) return GS_EALIGN;
    if (ld_table < dim || ld_out < dim) return GS_EINVAL;
    gather_mean_ragged_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(
        table, ld_table, dim, off, flat, n, out, ld_out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_scatter_mean_ragged(const float* gout, int64_t ld_gout, int32_t dim, const int32_t* off,
                                      const int32_t* flat, int32_t n, float* gtable, int64_t ld_gtable, void* stream) {
    if (n == 0) return GS_OK;
    if (!gout || !off || !flat || !gtable || dim <= 0 || n < 0) return GS_EINVAL;
    if (!gs_aligned16(gout) || !gs_aligned16(gtable) || (ld_gout & 3) || (ld_gtable & 3)) return GS_EALIGN;
    if (ld_gout < dim || ld_gtable < dim) return GS_EINVAL;
    scatter_mean_ragged_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(
        gout, ld_gout, dim, off, flat, n, gtable, ld_gtable);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
