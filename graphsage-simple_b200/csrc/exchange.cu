// Owner bucketing for the partitioned feature table / CSR (BASELINE config 5).
// The reference has no distributed code: `features(LongTensor(unique_nodes_list))`
// (graphsage/aggregators.py:62-65) and `adj_lists[int(node)]` (graphsage/encoders.py:47) are
// local dictionary lookups.  With the table and the CSR rows partitioned by owner = id % world,
// each lookup becomes: bucket the ids by owner (this file) -> all-to-all of ids -> the owner's
// local kernel (gs_gather_rows / gs_sample_csr) -> all-to-all of the answers -> un-permute
// (gs_gather_rows with the permutation as ids).
#include "gs_common.cuh"

namespace {

constexpr int kBlock = 1024;          // ids per block (one per thread)
constexpr int kMaxWorld = 16;

// pass 1: per-block histogram over owners -> block_hist[owner][block]
__global__ void __launch_bounds__(kBlock)
owner_hist_kernel(const int32_t* __restrict__ ids, int n_max, const int32_t* __restrict__ n_dev, int world,
                  int nb, int32_t* __restrict__ block_hist) {
    __shared__ int s_cnt[kMaxWorld];
    if (threadIdx.x < kMaxWorld) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int n = gs_row_count(n_max, n_dev);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const int o = i < n ? (int)((uint32_t)ids[i] % (uint32_t)world) : -1;
    // warp-aggregated: one shared atomic per owner per warp
    for (int w = 0; w < world; ++w) {
        const unsigned m = __ballot_sync(0xffffffffu, o == w);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[w], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x < world) block_hist[threadIdx.x * nb + blockIdx.x] = s_cnt[threadIdx.x];
}

// pass 2 (one block): exclusive scan of block_hist in (owner, block) order -> start offset of
// every (owner, block) cell in the send buffer; counts[owner] = bucket sizes
__global__ void __launch_bounds__(1024)
owner_scan_kernel(int32_t* __restrict__ block_hist, int total_cells, int nb, int world, int32_t* __restrict__ counts) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int c0 = 0; c0 < total_cells; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int v = c < total_cells ? block_hist[c] : 0;
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
        const int ex = carry + (w ? s_warp[w - 1] : 0) + inc - v;
        if (c < total_cells) block_hist[c] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    // bucket sizes from the offsets of each owner's first cell
    if (threadIdx.x < world) {
        const int lo = block_hist[threadIdx.x * nb];
        const int hi = threadIdx.x + 1 < world ? block_hist[(threadIdx.x + 1) * nb] : s_carry;
        counts[threadIdx.x] = hi - lo;
    }
}

// pass 3: stable scatter.  perm[i] = position of ids[i] in the send buffer.
__global__ void __launch_bounds__(kBlock)
owner_scatter_kernel(const int32_t* __restrict__ ids, int n_max, const int32_t* __restrict__ n_dev, int world,
                     int nb, const int32_t* __restrict__ block_hist, int emit_local,
                     int32_t* __restrict__ send_ids, int32_t* __restrict__ perm) {
    __shared__ int s_warp[kMaxWorld][32];
    const int n = gs_row_count(n_max, n_dev);
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t v = i < n ? ids[i] : 0;
    const int o = i < n ? (int)((uint32_t)v % (uint32_t)world) : -1;
    int rank_in_warp = 0;
    for (int q = 0; q < world; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, o == q);
        if (o == q) rank_in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[q][w] = __popc(m);
    }
    __syncthreads();
    if (w < world) {                 // warp q scans the 32 per-warp counts of owner q
        int c = s_warp[w][lane], inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        s_warp[w][lane] = inc - c;
    }
    __syncthreads();
    if (i < n) {
        const int pos = block_hist[o * nb + blockIdx.x] + s_warp[o][w] + rank_in_warp;
        send_ids[pos] = emit_local ? (int32_t)((uint32_t)v / (uint32_t)world) : v;
        perm[i] = pos;
    }
}

}  // namespace

extern "C" int32_t gs_bucket_scratch_ints(int32_t n_max, int32_t world) {
    const int nb = (n_max + kBlock - 1) / kBlock;
    return (nb > 0 ? nb : 1) * world;
}

extern "C" int gs_bucket_by_owner(const int32_t* ids, int32_t n_max, const int32_t* n_dev, int32_t world,
                                  int32_t emit_local, int32_t* scratch, int32_t* send_ids, int32_t* perm,
                                  int32_t* counts, void* stream) {
    if (!counts || n_max < 0) return GS_EINVAL;
    if (world < 1 || world > kMaxWorld) return GS_ENOSUP;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_max > 0 && (!ids || !scratch || !send_ids || !perm)) return GS_EINVAL;
    if (n_max == 0) {                   // an empty request is legal (a rank may ask for nothing)
        cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * world, s);
        return e == cudaSuccess ? GS_OK : (int)e;
    }
    const int nb = (n_max + kBlock - 1) / kBlock;
    owner_hist_kernel<<<nb, kBlock, 0, s>>>(ids, n_max, n_dev, world, nb, scratch);
    GS_LAUNCH_CHECK();
    owner_scan_kernel<<<1, 1024, 0, s>>>(scratch, nb * world, nb, world, counts);
    GS_LAUNCH_CHECK();
    owner_scatter_kernel<<<nb, kBlock, 0, s>>>(ids, n_max, n_dev, world, nb, scratch, emit_local, send_ids, perm);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
