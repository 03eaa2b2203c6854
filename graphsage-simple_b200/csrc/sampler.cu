// K1: CSR-resident counter-based neighbour sampler + frontier dedup.
// Replaces graphsage/aggregators.py:42-56 and the adjacency lookup at encoders.py:47 of
// the reference.  Specification restated on the CPU in oracle/sampler_port.py.
#include "gs_common.cuh"

namespace {

constexpr int kMaxK = 64;          // largest sampled fan-out (Floyd set lives in registers/local)

struct Philox {
    uint32_t c[4];
};

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox p; p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
    return p;
}

// One thread per frontier row.  Work per row is O(k^2) integer ops and k random 4-byte CSR
// reads -- three orders of magnitude below the feature gather it feeds, so the simple
// mapping is the right one; the grid is sized from n_max and trimmed by *n_dev.
__global__ void __launch_bounds__(128)
sample_csr_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int num_nodes,
                  const int32_t* __restrict__ nodes, int n_max, const int32_t* __restrict__ n_dev,
                  int k, int width, int add_self, uint32_t seed_lo, uint32_t seed_hi,
                  int64_t step_imm, const int64_t* __restrict__ step_dev,
                  uint32_t tag_head, uint32_t tag_tail, int n_head,
                  int32_t* __restrict__ idx, int32_t* __restrict__ cnt) {
    const int n = gs_row_count(n_max, n_dev);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    int32_t* out = idx + (int64_t)i * width;
    if (i >= n) {                      // keep unused rows well defined for downstream kernels
        for (int j = 0; j < width; ++j) out[j] = -1;
        cnt[i] = 0;
        return;
    }
    const int32_t v = nodes[i];
    // ids outside [0, num_nodes) are isolated nodes (the reference's defaultdict(set) answers an empty set)
    const bool known = (unsigned)v < (unsigned)num_nodes;
    const int64_t base = known ? rowptr[v] : 0;
    const int deg = known ? (int)(rowptr[v + 1] - base) : 0;
    const uint32_t step = (uint32_t)(step_dev ? *step_dev : step_imm);
    const uint32_t tag = i < n_head ? tag_head : tag_tail;
    int c = 0;
    bool has_self = false;
    if (k < 0 || deg <= k) {
        const int take = deg < width ? deg : width;
        for (int j = 0; j < take; ++j) {
            int32_t u = col[base + j];
            has_self |= (u == v);
            out[j] = u;
        }
        c = take;
    } else {
        int32_t pos[kMaxK];
        Philox rnd;
        for (int m = 0; m < k; ++m) {
            if ((m & 3) == 0) rnd = philox4x32_10((uint32_t)v, (uint32_t)(m >> 2), step, tag, seed_lo, seed_hi);
            const int j = deg - k + m;
            const uint32_t r = rnd.c[m & 3];
            int t = (int)(((uint64_t)r * (uint64_t)(j + 1)) >> 32);
            bool dup = false;
            for (int q = 0; q < m; ++q) dup |= (pos[q] == t);
            if (dup) t = j;
            // insertion keeps pos[] ascending
            int q = m;
            while (q > 0 && pos[q - 1] > t) { pos[q] = pos[q - 1]; --q; }
            pos[q] = t;
        }
        for (int m = 0; m < k; ++m) {
            int32_t u = col[base + pos[m]];
            has_self |= (u == v);
            out[m] = u;
        }
        c = k;
    }
    if (add_self && !has_self && c < width) out[c++] = v;
    for (int j = c; j < width; ++j) out[j] = -1;
    cnt[i] = c;
}

// Warp-per-row variant for k <= 32 (the common fan-outs): every lane draws its own Philox word
// up front, Floyd's sequential membership test is one ballot per draw, the ascending order
// comes from a rank count over shuffles, and the CSR reads/tile writes are coalesced.  Same
// specification (and bit-identical output) as the thread-per-row kernel above.
constexpr int kWarpRows = 8;      // warps (rows) per block

// Partitioned CSR: node v's neighbour list lives on rank v % world as local row v / world of that
// rank's (rowptr, col); `rowptrs[q]` / `cols[q]` are rank q's arrays mapped into this process (symmetric memory), so
// the sampling warp reads a remote adjacency row straight over NVLink.  Draws are keyed on the GLOBAL id.
struct CsrPeers {
    const int64_t* const* rowptrs;
    const int32_t* const* cols;
    int world, shift;
};

template <bool PEER>
__global__ void __launch_bounds__(kWarpRows * 32)
sample_csr_warp_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const CsrPeers peers,
                       int num_nodes, const int32_t* __restrict__ nodes, int n_max, const int32_t* __restrict__ n_dev,
                       int k, int width, int add_self, uint32_t seed_lo, uint32_t seed_hi,
                       int64_t step_imm, const int64_t* __restrict__ step_dev,
                       uint32_t tag_head, uint32_t tag_tail, int n_head,
                       int32_t* __restrict__ idx, int32_t* __restrict__ cnt) {
    const int n = gs_row_count(n_max, n_dev);
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kWarpRows + (threadIdx.x >> 5);
    if (i >= n_max) return;
    int32_t* out = idx + (int64_t)i * width;
    if (i >= n) {
        for (int j = lane; j < width; j += 32) out[j] = -1;
        if (lane == 0) cnt[i] = 0;
        return;
    }
    const int32_t v = nodes[i];
    int r = v;
    if (PEER) {
        const unsigned u = (unsigned)v;
        const unsigned owner = peers.shift >= 0 ? (u & (unsigned)(peers.world - 1)) : (u % (unsigned)peers.world);
        r = (int)(peers.shift >= 0 ? (u >> peers.shift) : (u / (unsigned)peers.world));
        rowptr = peers.rowptrs[owner];
        col = peers.cols[owner];
    }
    const bool known = (unsigned)v < (unsigned)num_nodes;     // unknown id = isolated node: empty neighbour set
    const int64_t base = known ? rowptr[r] : 0;
    const int deg = known ? (int)(rowptr[r + 1] - base) : 0;
    int c;
    bool has_self = false;
    if (k < 0 || deg <= k) {
        c = deg < width ? deg : width;
        for (int j = lane; j < c; j += 32) {
            const int32_t u = col[base + j];
            has_self |= (u == v);
            out[j] = u;
        }
    } else {
        const uint32_t step = (uint32_t)(step_dev ? *step_dev : step_imm);
        const uint32_t tag = i < n_head ? tag_head : tag_tail;
        uint32_t r = 0;
        if (lane < k) r = philox4x32_10((uint32_t)v, (uint32_t)(lane >> 2), step, tag, seed_lo, seed_hi).c[lane & 3];
        int mine = -1;                               // lane m holds the m-th chosen position
        for (int m = 0; m < k; ++m) {
            const int j = deg - k + m;
            const uint32_t rm = __shfl_sync(0xffffffffu, r, m);
            int t = (int)(((uint64_t)rm * (uint64_t)(j + 1)) >> 32);
            const unsigned dup = __ballot_sync(0xffffffffu, lane < m && mine == t);
            if (dup) t = j;
            if (lane == m) mine = t;
        }
        int rank = 0;
        for (int l = 0; l < k; ++l) rank += (__shfl_sync(0xffffffffu, mine, l) < mine) ? 1 : 0;
        if (lane < k) {
            const int32_t u = col[base + mine];
            has_self = (u == v);
            out[rank] = u;
        }
        c = k;
    }
    has_self = __any_sync(0xffffffffu, has_self);
    if (add_self && !has_self && c < width) {
        if (lane == 0) out[c] = v;
        ++c;
    }
    for (int j = c + lane; j < width; j += 32) out[j] = -1;
    if (lane == 0) cnt[i] = c;
}

// ---- dedup over a node BITMAP: mark -> count (+ scan of block sums by the last block) -> compact -> remap -> clear ----
// One bit per node id (N / 32 words) plus one running-rank word per bitmap word, both inside the caller's scratch.
// Work per call: the frontier's entries (mark, remap, clear) + a popcount scan over N / 32 words (0.3 MB at 2.4 M
// nodes; round 1 cleared and scanned a 4-byte slot per node: 9.6 MB + 2 x 9.6 MB).  The bitmap is zero on entry and
// zero on exit: only the words a call touched are cleared again, by the ids it found.
constexpr int kWordsPerThread = 8;
constexpr int kScanThreads = 256;
constexpr int kWordsPerBlock = kWordsPerThread * kScanThreads;     // 2048 words = 65 536 node ids per block

__global__ void dedup_mark_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                  int n_max, const int32_t* __restrict__ n_dev, int width,
                                  uint32_t* __restrict__ bitmap) {
    const int n = gs_row_count(n_max, n_dev);
    const int64_t total = (int64_t)n * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / width), j = (int)(e - (int64_t)i * width);
        if (cnt == nullptr || j < cnt[i]) {
            const uint32_t id = (uint32_t)idx[e];
            atomicOr(&bitmap[id >> 5], 1u << (id & 31u));
        }
    }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total_out) {
    // kScanThreads-wide exclusive scan (warp shuffles + one smem hop)
    __shared__ int warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int before = w ? warp_sums[w - 1] : 0;
    if (total_out) *total_out = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return before + inc - v;
}

// block_counts layout: [0..nb) per-block counts -> exclusive offsets, [nb] ticket counter
__global__ void __launch_bounds__(kScanThreads)
dedup_count_kernel(const uint32_t* __restrict__ bitmap, int nwords, int nb,
                   int32_t* __restrict__ block_counts, int slot_base, int32_t* __restrict__ n_total_dev) {
    __shared__ int s_last;
    const int first = blockIdx.x * kWordsPerBlock + threadIdx.x * kWordsPerThread;
    int mine = 0;
#pragma unroll
    for (int t = 0; t < kWordsPerThread; ++t)
        if (first + t < nwords) mine += __popc(bitmap[first + t]);
    int total;
    block_exclusive_scan(mine, &total);
    if (threadIdx.x == 0) {
        block_counts[blockIdx.x] = total;
        __threadfence();
        const int ticket = atomicAdd(&block_counts[nb], 1);
        s_last = (ticket == nb - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: exclusive scan over the nb block sums (nb is small: num_nodes / 65 536)
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += kScanThreads) {
        const int b = b0 + threadIdx.x;
        const int v = b < nb ? ((volatile int32_t*)block_counts)[b] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, &tot);
        if (b < nb) block_counts[b] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        *n_total_dev = slot_base + carry;
        block_counts[nb] = 0;                       // re-arm the ticket for the next call
    }
}

// uniq[rank] = id in ascending id order; word_rank[w] = slot of the first set bit of word w
__global__ void __launch_bounds__(kScanThreads)
dedup_compact_kernel(const uint32_t* __restrict__ bitmap, int nwords, const int32_t* __restrict__ block_counts,
                     int slot_base, int32_t* __restrict__ word_rank, int32_t* __restrict__ uniq) {
    const int first = blockIdx.x * kWordsPerBlock + threadIdx.x * kWordsPerThread;   // consecutive words per thread keep order
    uint32_t words[kWordsPerThread];
    int mine = 0;
#pragma unroll
    for (int t = 0; t < kWordsPerThread; ++t) {
        words[t] = first + t < nwords ? bitmap[first + t] : 0u;
        mine += __popc(words[t]);
    }
    int rank = block_counts[blockIdx.x] + block_exclusive_scan(mine, nullptr);
#pragma unroll
    for (int t = 0; t < kWordsPerThread; ++t) {
        uint32_t w = words[t];
        if (w == 0u) continue;
        word_rank[first + t] = slot_base + rank;
        while (w) {
            const int bit = __ffs(w) - 1;
            uniq[rank++] = ((first + t) << 5) + bit;
            w &= w - 1u;
        }
    }
}

__global__ void dedup_remap_kernel(int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                   int n_max, const int32_t* __restrict__ n_dev, int width,
                                   const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank) {
    const int n = gs_row_count(n_max, n_dev);
    const int64_t total = (int64_t)n * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / width), j = (int)(e - (int64_t)i * width);
        if (cnt == nullptr || j < cnt[i]) {
            const uint32_t id = (uint32_t)idx[e];
            idx[e] = word_rank[id >> 5] + __popc(bitmap[id >> 5] & ((1u << (id & 31u)) - 1u));
        }
    }
}

// zero exactly the bitmap words this call set (idempotent plain stores), found through the ids it produced
__global__ void dedup_clear_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_total_dev, int slot_base,
                                   int cap, uint32_t* __restrict__ bitmap) {
    const int n = min(*n_total_dev - slot_base, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        bitmap[(uint32_t)uniq[i] >> 5] = 0u;
}

// plain table lookup of ids (gs_remap_ids)
__global__ void remap_ids_kernel(int32_t* __restrict__ idx, const int32_t* __restrict__ cnt,
                                 int n_max, const int32_t* __restrict__ n_dev, int width,
                                 const int32_t* __restrict__ map) {
    const int n = gs_row_count(n_max, n_dev);
    const int64_t total = (int64_t)n * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / width), j = (int)(e - (int64_t)i * width);
        if (cnt == nullptr || j < cnt[i]) idx[e] = map[idx[e]];
    }
}

__global__ void advance_step_kernel(int64_t* step) { *step += 1; }

// Device-side batch queue: copy staging block (*cursor % n_blocks) of a pool that already lives in HBM into the
// frontier set's staging area and advance the cursor -- what lets several train steps sit in ONE captured graph
// (the source of a captured memcpy is fixed at capture time; this kernel reads it from device memory).
__global__ void __launch_bounds__(256)
stage_next_kernel(const uint32_t* __restrict__ pool, int64_t block_words, int64_t n_blocks, int64_t* cursor,
                  uint32_t* __restrict__ dst) {
    const int64_t cur = *reinterpret_cast<volatile int64_t*>(cursor);
    const uint32_t* src = pool + (cur % n_blocks) * block_words;
    for (int64_t i = threadIdx.x; i < block_words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) *cursor = cur + 1;
}

}  // namespace

extern "C" int gs_sample_csr(const int64_t* rowptr, const int32_t* col, int32_t num_nodes,
                             const int32_t* nodes, int32_t n_max, const int32_t* n_dev,
                             int32_t k, int32_t width, int32_t add_self,
                             uint64_t seed, int64_t step, const int64_t* step_dev,
                             uint32_t tag_head, uint32_t tag_tail, int32_t n_head,
                             int32_t* idx, int32_t* cnt, void* stream) {
    if (n_max == 0) return GS_OK;       // empty frontier (e.g. nothing requested from this owner)
    if (!rowptr || !col || !nodes || !idx || !cnt || n_max < 0 || width <= 0 || num_nodes < 0) return GS_EINVAL;
    if (k > kMaxK) return GS_ENOSUP;
    if (k >= 0 && width < k + (add_self ? 1 : 0)) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    if (k <= 32) {
        GS_PREFER_SMEM(sample_csr_warp_kernel<false>);
        sample_csr_warp_kernel<false><<<(n_max + kWarpRows - 1) / kWarpRows, kWarpRows * 32, 0, (cudaStream_t)stream>>>(
            rowptr, col, CsrPeers{nullptr, nullptr, 1, 0}, num_nodes, nodes, n_max, n_dev, k, width, add_self, (uint32_t)seed,
            (uint32_t)(seed >> 32), step, step_dev, tag_head, tag_tail, n_head, idx, cnt);
    } else {
        const int threads = 128;
        sample_csr_kernel<<<(n_max + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
            rowptr, col, num_nodes, nodes, n_max, n_dev, k, width, add_self, (uint32_t)seed, (uint32_t)(seed >> 32),
            step, step_dev, tag_head, tag_tail, n_head, idx, cnt);
    }
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_sample_csr_peer(const int64_t* const* rowptrs, const int32_t* const* cols, int32_t world,
                                  int32_t num_nodes, const int32_t* nodes, int32_t n_max, const int32_t* n_dev,
                                  int32_t k, int32_t width, int32_t add_self,
                                  uint64_t seed, int64_t step, const int64_t* step_dev,
                                  uint32_t tag_head, uint32_t tag_tail, int32_t n_head,
                                  int32_t* idx, int32_t* cnt, void* stream) {
    if (n_max == 0) return GS_OK;
    if (!rowptrs || !cols || !nodes || !idx || !cnt || n_max < 0 || width <= 0 || num_nodes < 0) return GS_EINVAL;
    if (world < 1 || world > 16 || k > 32) return GS_ENOSUP;
    if (k >= 0 && width < k + (add_self ? 1 : 0)) return GS_EINVAL;
    int shift = -1;
    for (int b = 0; b < 5; ++b) if ((1 << b) == world) shift = b;
    GS_PREFER_SMEM(sample_csr_warp_kernel<true>);
    sample_csr_warp_kernel<true><<<(n_max + kWarpRows - 1) / kWarpRows, kWarpRows * 32, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, CsrPeers{rowptrs, cols, world, shift}, num_nodes, nodes, n_max, n_dev, k, width, add_self, (uint32_t)seed,
        (uint32_t)(seed >> 32), step, step_dev, tag_head, tag_tail, n_head, idx, cnt);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int32_t gs_dedup_scratch_ints(int32_t num_nodes) {
    const int nwords = (num_nodes + 31) / 32;
    return (nwords + kWordsPerBlock - 1) / kWordsPerBlock + 1;
}

extern "C" int gs_dedup_remap(int32_t* idx, const int32_t* cnt, int32_t n_max, const int32_t* n_dev,
                              int32_t width, int32_t num_nodes, int32_t* slot_of, int32_t* block_counts,
                              int32_t slot_base, int32_t* uniq, int32_t* n_total_dev, void* stream) {
    if (!idx || !slot_of || !block_counts || !uniq || !n_total_dev || num_nodes <= 0 || width <= 0 ||
        n_max < 0)
        return GS_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    // scratch `slot_of` (num_nodes ints, ZERO on entry and on exit): [0, nwords) the node bitmap, [nwords, 2 nwords)
    // the rank of each word's first set bit
    const int nwords = (num_nodes + 31) / 32;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(slot_of);
    int32_t* word_rank = slot_of + nwords;
    const int nb = (nwords + kWordsPerBlock - 1) / kWordsPerBlock;
    const int64_t entries = (int64_t)n_max * width;
    int eb = (int)((entries + 255) / 256);
    if (eb > GS_NUM_SMS * 8) eb = GS_NUM_SMS * 8;
    if (eb < 1) eb = 1;
    const int64_t cap64 = entries < num_nodes ? entries : num_nodes;
    const int cap = (int)cap64;
    GS_PREFER_SMEM(dedup_mark_kernel);
    GS_PREFER_SMEM(dedup_count_kernel);
    GS_PREFER_SMEM(dedup_compact_kernel);
    GS_PREFER_SMEM(dedup_remap_kernel);
    GS_PREFER_SMEM(dedup_clear_kernel);
    dedup_mark_kernel<<<eb, 256, 0, s>>>(idx, cnt, n_max, n_dev, width, bitmap);
    GS_LAUNCH_CHECK();
    dedup_count_kernel<<<nb, kScanThreads, 0, s>>>(bitmap, nwords, nb, block_counts, slot_base, n_total_dev);
    GS_LAUNCH_CHECK();
    dedup_compact_kernel<<<nb, kScanThreads, 0, s>>>(bitmap, nwords, block_counts, slot_base, word_rank, uniq);
    GS_LAUNCH_CHECK();
    dedup_remap_kernel<<<eb, 256, 0, s>>>(idx, cnt, n_max, n_dev, width, bitmap, word_rank);
    GS_LAUNCH_CHECK();
    int cb = (cap + 255) / 256;
    if (cb > GS_NUM_SMS * 4) cb = GS_NUM_SMS * 4;
    if (cb < 1) cb = 1;
    dedup_clear_kernel<<<cb, 256, 0, s>>>(uniq, n_total_dev, slot_base, cap, bitmap);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_remap_ids(int32_t* idx, const int32_t* cnt, int32_t n_max, const int32_t* n_dev, int32_t width,
                            const int32_t* map, void* stream) {
    if (n_max == 0) return GS_OK;
    if (!idx || !map || n_max < 0 || width <= 0) return GS_EINVAL;
    const int64_t entries = (int64_t)n_max * width;
    int eb = (int)((entries + 255) / 256);
    if (eb > GS_NUM_SMS * 8) eb = GS_NUM_SMS * 8;
    GS_PREFER_SMEM(remap_ids_kernel);
    remap_ids_kernel<<<eb, 256, 0, (cudaStream_t)stream>>>(idx, cnt, n_max, n_dev, width, map);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_stage_next(const void* pool, int64_t block_bytes, int64_t n_blocks, int64_t* cursor, void* dst,
                             void* stream) {
    if (!pool || !cursor || !dst || block_bytes <= 0 || n_blocks <= 0) return GS_EINVAL;
    if ((block_bytes & 3) || (reinterpret_cast<uintptr_t>(pool) & 3) || (reinterpret_cast<uintptr_t>(dst) & 3)) return GS_EALIGN;
    GS_PREFER_SMEM(stage_next_kernel);
    stage_next_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(pool), block_bytes / 4, n_blocks,
                                                          cursor, reinterpret_cast<uint32_t*>(dst));
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_advance_step(int64_t* step_dev, void* stream) {
    if (!step_dev) return GS_EINVAL;
    advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
