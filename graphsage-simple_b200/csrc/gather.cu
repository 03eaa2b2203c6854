// K2 gather-mean forward, K4 scatter-add backward, row gather, SGD.
// Replaces graphsage/aggregators.py:54-65,74 (dense mask + mask.mm), encoders.py:49-54
// (self lookup + cat), the autograd backward of mask.mm (model.py:249) and the SGD step
// (model.py:250) of the reference.
#include <cstdlib>
#include "gs_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

// Where feature row `id` lives.  Plain: one table.  Partitioned: node v is owned by rank
// v % world as local row v / world, and `peers[q]` is rank q's shard mapped into this process (symmetric
// memory), so a remote row is read straight over NVLink by the gathering warp -- the exchange step of the
// partitioned path fused into the gather itself, no all-to-all, no staging buffer.
struct TableRef {
    const float* table;             // world == 1
    const float* const* peers;      // world > 1
    int world, shift;               // shift = log2(world) when world is a power of two, else -1
};

template <bool PEER>
__device__ __forceinline__ const float* row_ptr(const TableRef& t, int id, int64_t ld) {
    if (!PEER) return t.table + (int64_t)id * ld;
    const unsigned u = (unsigned)id;
    const unsigned owner = t.shift >= 0 ? (u & (unsigned)(t.world - 1)) : (u % (unsigned)t.world);
    const unsigned local = t.shift >= 0 ? (u >> t.shift) : (u / (unsigned)t.world);
    return t.peers[owner] + (int64_t)local * ld;       // 8-byte pointer table, L1-resident
}

// L2 eviction policies (createpolicy): the feature rows stream through once per launch (evict_first), the tile the
// kernel writes is read again by the encoder GEMMs right after (evict_last), so that the 126 MB L2 keeps as much of the
// 122 MB tile as it can instead of the 650 MB of rows that pass through it.  mode 0: both evict_normal (the default
// behaviour of plain accesses).
__device__ __forceinline__ uint64_t l2_policy(int which) {       // 0 normal, 1 evict_first, 2 evict_last
    uint64_t p;
    if (which == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (which == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_stream_hint(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint_v4(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_v2(float* p, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

// Load 4 consecutive floats of a row whose base is 16-B aligned; columns >= dim read as 0.
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ row, int c4, int dim, uint64_t pol) {
    const int col = c4 * 4;
    if (col + 4 <= dim) return ldg_stream_hint(reinterpret_cast<const float4*>(row) + c4, pol);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < dim) v.x = __ldg(row + col);
    if (col + 1 < dim) v.y = __ldg(row + col + 1);
    if (col + 2 < dim) v.z = __ldg(row + col + 2);
    return v;
}

__device__ __forceinline__ float4 load_chunk(const float* __restrict__ row, int c4, int dim) {
    const int col = c4 * 4;
    if (col + 4 <= dim) return gs_ldg_stream(reinterpret_cast<const float4*>(row) + c4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < dim) v.x = __ldg(row + col);
    if (col + 1 < dim) v.y = __ldg(row + col + 1);
    if (col + 2 < dim) v.z = __ldg(row + col + 2);
    return v;
}

// Store 4 floats at dst (alignment `al` floats: 4, 2 or 1), only columns < dim.
__device__ __forceinline__ void store_chunk(float* __restrict__ dst_row, int c4, int dim, int al, float4 v, uint64_t pol) {
    const int col = c4 * 4;
    float* d = dst_row + col;
    if (col + 4 <= dim) {
        if (al == 4) { st_hint_v4(d, v, pol); }
        else if (al == 2) {
            st_hint_v2(d, make_float2(v.x, v.y), pol);
            st_hint_v2(d + 2, make_float2(v.z, v.w), pol);
        } else { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
    } else {
        if (col < dim) d[0] = v.x;
        if (col + 1 < dim) d[1] = v.y;
        if (col + 2 < dim) d[2] = v.z;
    }
}

__device__ __forceinline__ void store_chunk(float* __restrict__ dst_row, int c4, int dim, int al, float4 v) {
    const int col = c4 * 4;
    float* d = dst_row + col;
    if (col + 4 <= dim) {
        if (al == 4) { *reinterpret_cast<float4*>(d) = v; }
        else if (al == 2) {
            *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
            *reinterpret_cast<float2*>(d + 2) = make_float2(v.z, v.w);
        } else { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
    } else {
        if (col < dim) d[0] = v.x;
        if (col + 1 < dim) d[1] = v.y;
        if (col + 2 < dim) d[2] = v.z;
    }
}

// One warp per target row.  Lane l owns float4 chunks l, l+32, ... of the feature row, so a
// warp-wide load is one contiguous 512-byte run of the neighbour's row (4 full 128-B lines).
// CH chunks per lane are kept in registers and NB neighbours are in flight at once, i.e.
// CH*NB independent 128-bit loads per lane, which is what hides HBM latency here.
template <int CH, int NB, bool PEER>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
gather_mean_kernel(const TableRef table, int64_t ld_table, int dim,
                   const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt, int width,
                   const int32_t* __restrict__ self_ids, int n_max, const int32_t* __restrict__ n_dev,
                   float* __restrict__ out, int64_t ld_out, int neigh_off, int out_align, int l2_mode) {
    const int n = gs_row_count(n_max, n_dev);
    const int lane = threadIdx.x & 31;
    const int nchunks = (dim + 3) >> 2;
    const uint64_t pol_ld = l2_policy(l2_mode ? 1 : 0), pol_st = l2_policy(l2_mode ? 2 : 0);
    // grid-stride over rows: the grid is capped at a fixed number of blocks per SM so that the
    // kernel leaves room for a co-resident tensor-core CTA (engine.py pipelining)
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += gridDim.x * wpb) {
    const int c = min(cnt[row], width);
    const float inv = c > 0 ? 1.f / (float)c : 0.f;
    float* orow = out + (int64_t)row * ld_out;
    const int32_t* irow = idx + (int64_t)row * width;

    if (self_ids != nullptr) {          // bit-exact copy of the node's own row (encoders.py:53)
        const float* srow = row_ptr<PEER>(table, self_ids[row], ld_table);
        for (int c4 = lane; c4 < nchunks; c4 += 32) store_chunk(orow, c4, dim, 4, load_chunk(srow, c4, dim, pol_ld), pol_st);
    }
    for (int c0 = 0; c0 < nchunks; c0 += 32 * CH) {
        float4 acc[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < c; j0 += 32) {
            const int my = (j0 + lane < c) ? irow[j0 + lane] : 0;
            const int lim = min(32, c - j0);
            int j = 0;
            for (; j + NB <= lim; j += NB) {
                float4 v[NB][CH];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float* nrow = row_ptr<PEER>(table, __shfl_sync(0xffffffffu, my, j + b), ld_table);
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const int c4 = c0 + u * 32 + lane;
                        v[b][u] = c4 < nchunks ? load_chunk(nrow, c4, dim, pol_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        acc[u].x += v[b][u].x; acc[u].y += v[b][u].y;
                        acc[u].z += v[b][u].z; acc[u].w += v[b][u].w;
                    }
            }
            for (; j < lim; ++j) {
                const float* nrow = row_ptr<PEER>(table, __shfl_sync(0xffffffffu, my, j), ld_table);
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int c4 = c0 + u * 32 + lane;
                    if (c4 < nchunks) {
                        float4 t = load_chunk(nrow, c4, dim, pol_ld);
                        acc[u].x += t.x; acc[u].y += t.y; acc[u].z += t.z; acc[u].w += t.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int c4 = c0 + u * 32 + lane;
            if (c4 < nchunks) {
                float4 m = make_float4(acc[u].x * inv, acc[u].y * inv, acc[u].z * inv, acc[u].w * inv);
                store_chunk(orow + neigh_off, c4, dim, out_align, m, pol_st);
            }
        }
    }
    }
}

template <bool PEER>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gather_rows_kernel(const TableRef table, int64_t ld_table, int dim,
                   const int32_t* __restrict__ ids, int n_max, const int32_t* __restrict__ n_dev,
                   float* __restrict__ out, int64_t ld_out) {
    const int n = gs_row_count(n_max, n_dev);
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n) return;
    const int nchunks = (dim + 3) >> 2;
    const float* srow = row_ptr<PEER>(table, ids[row], ld_table);
    float* orow = out + (int64_t)row * ld_out;
    for (int c4 = lane; c4 < nchunks; c4 += 32) store_chunk(orow, c4, dim, 4, load_chunk(srow, c4, dim));
}

// Backward of the mean: one warp per target row; the row's gradient chunk is read once,
// scaled by 1/cnt and pushed to every sampled neighbour's row with 128-bit reductions
// (RED.ADD.F32x4 resolves at L2 -- no return value, no read-modify-write round trip).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scatter_mean_kernel(const float* __restrict__ gout, int64_t ld_gout, int neigh_off, int dim,
                    const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt, int width,
                    const int32_t* __restrict__ self_ids, int n_max, const int32_t* __restrict__ n_dev,
                    float* __restrict__ gtable, int64_t ld_gtable, int in_align) {
    const int n = gs_row_count(n_max, n_dev);
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= n) return;
    const int nchunks = (dim + 3) >> 2;
    const int c = min(cnt[row], width);
    const float inv = c > 0 ? 1.f / (float)c : 0.f;
    const float* grow = gout + (int64_t)row * ld_gout;
    const int32_t* irow = idx + (int64_t)row * width;
    for (int c4 = lane; c4 < nchunks; c4 += 32) {
        const int col = c4 * 4;
        const bool full = col + 4 <= dim;
        if (self_ids != nullptr) {
            float4 g = load_chunk(grow, c4, dim);
            float* dst = gtable + (int64_t)self_ids[row] * ld_gtable + col;
            if (full) gs_red_add_v4(dst, g);
            else { if (col < dim) atomicAdd(dst, g.x); if (col + 1 < dim) atomicAdd(dst + 1, g.y);
                   if (col + 2 < dim) atomicAdd(dst + 2, g.z); }
        }
        float4 g;
        const float* src = grow + neigh_off;
        if (in_align == 4) g = load_chunk(src, c4, dim);
        else {
            g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < dim) g.x = src[col];
            if (col + 1 < dim) g.y = src[col + 1];
            if (col + 2 < dim) g.z = src[col + 2];
            if (col + 3 < dim) g.w = src[col + 3];
        }
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
        for (int j = 0; j < c; ++j) {
            float* dst = gtable + (int64_t)irow[j] * ld_gtable + col;
            if (full) gs_red_add_v4(dst, g);
            else { if (col < dim) atomicAdd(dst, g.x); if (col + 1 < dim) atomicAdd(dst + 1, g.y);
                   if (col + 2 < dim) atomicAdd(dst + 2, g.z); }
        }
    }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float lr, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = p[i] - lr * g[i];          // same two roundings as torch's add_(g, alpha=-lr)
}

}  // namespace

namespace {

TableRef peer_ref(const float* const* tables, int world) {
    int shift = -1;
    for (int b = 0; b < 5; ++b) if ((1 << b) == world) shift = b;
    return TableRef{nullptr, tables, world, shift};
}

int launch_gather_mean(const TableRef& table, bool peer, int64_t ld_table, int32_t dim,
                       const int32_t* idx, const int32_t* cnt, int32_t width,
                       const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                       float* out, int64_t ld_out, int32_t neigh_off, void* stream) {
    const int align = (neigh_off & 3) == 0 ? 4 : ((neigh_off & 1) == 0 ? 2 : 1);
    const int nchunks = (dim + 3) / 4;
    // Register budget of an SM (64 K): the gather must leave room for a co-resident tcgen05 GEMM CTA
    // (320 threads x 72 regs = 23 K) or the fused head (256 x 96 = 24.6 K) AND a sampler block (8 K):
    // 3 blocks of 4 warps x 80 regs = 30.7 K, i.e. 12 warps x 10 independent 128-bit loads per lane
    // = 61 KB in flight per SM (Little: 6.5 TB/s x ~800 ns / 148 SMs = 35 KB).  Tunable for experiments.
    // (Round 2: a column-split variant -- 2 warps per row, 3 chunks x 3 neighbours per lane in 64 registers, 16 warps
    // per SM = 72 KB in flight -- measured the SAME 0.183 ms: with the 164 KB shared / 64 KB L1 split the co-resident
    // GEMM needs, the bound is the L1's outstanding-request capacity, not warps x registers; profiles/README.md.)
    static int bps = 0, wpb = 0, carve = 0, l2_mode = 0;
    if (bps == 0) {
        l2_mode = getenv("GSAGE_GATHER_L2") ? atoi(getenv("GSAGE_GATHER_L2")) : 0;
        // L1/shared split preference (percent of shared; 1 = max shared, 0 = driver default).  The split can only
        // change on an idle SM, so a resident gather with the default (L1-heavy) split keeps the 145 KB GEMM CTAs
        // and the 175 KB head CTAs out; max-shared starves the gather's outstanding loads of L1.  72 % = 164 KB
        // shared / 92 KB L1 serves both (0.255 ms/step vs 0.275 max-shared, 0.30+ default; profiles/README.md s4).
        carve = getenv("GSAGE_GATHER_CARVEOUT") ? atoi(getenv("GSAGE_GATHER_CARVEOUT")) : 72;
        const char* e = getenv("GSAGE_GATHER_BPS");
        bps = e ? atoi(e) : 3;
        if (bps < 1) bps = 1;
        e = getenv("GSAGE_GATHER_WPB");
        wpb = e ? atoi(e) : 4;
        if (wpb < 1 || wpb > kWarpsPerBlock) wpb = 4;
    }
    int nblocks = (n_max + wpb - 1) / wpb;
    if (nblocks > GS_NUM_SMS * bps) nblocks = GS_NUM_SMS * bps;
    const dim3 grid(nblocks), block(wpb * 32);
    cudaStream_t s = (cudaStream_t)stream;
#define GS_GM1(CH, NB, P) do { \
        static bool d__ = false; \
        if (!d__ && carve) { cudaFuncSetAttribute((gather_mean_kernel<CH, NB, P>), cudaFuncAttributePreferredSharedMemoryCarveout, \
                                                  carve == 1 ? (int)cudaSharedmemCarveoutMaxShared : carve); d__ = true; } \
        gather_mean_kernel<CH, NB, P><<<grid, block, 0, s>>>(table, ld_table, dim, idx, cnt, width, \
            self_ids, n_max, n_dev, out, ld_out, neigh_off, align, l2_mode); } while (0)
#define GS_GM(CH, NB) do { if (peer) GS_GM1(CH, NB, true); else GS_GM1(CH, NB, false); } while (0)
    if (nchunks <= 32) { GS_GM(1, 8); }
    else if (nchunks <= 64) { GS_GM(2, 4); }
    else if (nchunks <= 96) { GS_GM(3, 4); }
    else if (nchunks <= 128) { GS_GM(4, 2); }
    else { GS_GM(5, 2); }
#undef GS_GM
#undef GS_GM1
    GS_LAUNCH_CHECK();
    return GS_OK;
}

}  // namespace

extern "C" int gs_gather_mean_fwd(const float* table, int64_t ld_table, int32_t dim,
                                  const int32_t* idx, const int32_t* cnt, int32_t width,
                                  const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                                  float* out, int64_t ld_out, int32_t neigh_off, void* stream) {
    if (!table || !idx || !cnt || !out || dim <= 0 || width <= 0 || n_max < 0 || neigh_off < 0) return GS_EINVAL;
    if (!gs_aligned16(table) || !gs_aligned16(out) || (ld_table & 3) || (ld_out & 3)) return GS_EALIGN;
    if (ld_table < dim || ld_out < neigh_off + dim) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    return launch_gather_mean(TableRef{table, nullptr, 1, 0}, false, ld_table, dim, idx, cnt, width, self_ids, n_max, n_dev,
                              out, ld_out, neigh_off, stream);
}

extern "C" int gs_gather_mean_fwd_peer(const float* const* tables, int32_t world, int64_t ld_table, int32_t dim,
                                       const int32_t* idx, const int32_t* cnt, int32_t width,
                                       const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                                       float* out, int64_t ld_out, int32_t neigh_off, void* stream) {
    if (!tables || !idx || !cnt || !out || dim <= 0 || width <= 0 || n_max < 0 || neigh_off < 0) return GS_EINVAL;
    if (world < 1 || world > 16) return GS_ENOSUP;
    if (!gs_aligned16(out) || (ld_table & 3) || (ld_out & 3)) return GS_EALIGN;
    if (ld_table < dim || ld_out < neigh_off + dim) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    return launch_gather_mean(peer_ref(tables, world), true, ld_table, dim, idx, cnt, width, self_ids, n_max, n_dev,
                              out, ld_out, neigh_off, stream);
}

extern "C" int gs_gather_rows(const float* table, int64_t ld_table, int32_t dim, const int32_t* ids,
                              int32_t n_max, const int32_t* n_dev, float* out, int64_t ld_out, void* stream) {
    if (n_max == 0) return GS_OK;
    if (!table || !ids || !out || dim <= 0 || n_max < 0) return GS_EINVAL;
    if (!gs_aligned16(table) || !gs_aligned16(out) || (ld_table & 3) || (ld_out & 3)) return GS_EALIGN;
    if (ld_table < dim || ld_out < dim) return GS_EINVAL;
    GS_PREFER_SMEM(gather_rows_kernel<false>);
    gather_rows_kernel<false><<<(n_max + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        TableRef{table, nullptr, 1, 0}, ld_table, dim, ids, n_max, n_dev, out, ld_out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_gather_rows_peer(const float* const* tables, int32_t world, int64_t ld_table, int32_t dim,
                                   const int32_t* ids, int32_t n_max, const int32_t* n_dev, float* out, int64_t ld_out,
                                   void* stream) {
    if (n_max == 0) return GS_OK;
    if (!tables || !ids || !out || dim <= 0 || n_max < 0) return GS_EINVAL;
    if (world < 1 || world > 16) return GS_ENOSUP;
    if (!gs_aligned16(out) || (ld_table & 3) || (ld_out & 3)) return GS_EALIGN;
    if (ld_table < dim || ld_out < dim) return GS_EINVAL;
    GS_PREFER_SMEM(gather_rows_kernel<true>);
    gather_rows_kernel<true><<<(n_max + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        peer_ref(tables, world), ld_table, dim, ids, n_max, n_dev, out, ld_out);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_scatter_mean_bwd(const float* gout, int64_t ld_gout, int32_t neigh_off, int32_t dim,
                                   const int32_t* idx, const int32_t* cnt, int32_t width,
                                   const int32_t* self_ids, int32_t n_max, const int32_t* n_dev,
                                   float* gtable, int64_t ld_gtable, void* stream) {
    if (!gout || !idx || !cnt || !gtable || dim <= 0 || width <= 0 || n_max < 0 || neigh_off < 0) return GS_EINVAL;
    if (!gs_aligned16(gout) || !gs_aligned16(gtable) || (ld_gout & 3) || (ld_gtable & 3)) return GS_EALIGN;
    if (ld_gtable < dim || ld_gout < neigh_off + dim) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    const int align = (neigh_off & 3) == 0 ? 4 : 1;
    GS_PREFER_SMEM(scatter_mean_kernel);
    scatter_mean_kernel<<<(n_max + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        gout, ld_gout, neigh_off, dim, idx, cnt, width, self_ids, n_max, n_dev, gtable, ld_gtable, align);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_sgd_step(float* p, const float* g, float lr, int64_t n, void* stream) {
    if (!p || !g || n < 0) return GS_EINVAL;
    if (n == 0) return GS_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > GS_NUM_SMS * 8) blocks = GS_NUM_SMS * 8;
    GS_PREFER_SMEM(sgd_kernel);
    sgd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, lr, n);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
