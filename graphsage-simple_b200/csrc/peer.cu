// Data-parallel gradient all-reduce fused with the SGD update, over NVLink peer memory.
//
// The reference has no distributed code; its update is `optimizer.step()` on local gradients
// (graphsage/model.py:237, 250).  With one process per GPU each rank holds the gradient of its share
// of the global batch; the update every rank must apply is  p -= lr * sum_r g_r.  Instead of an NCCL
// all-reduce followed by an SGD kernel (two launches that cannot sit inside the step's CUDA graph
// without a collective in the capture), ONE kernel does both in two phases over peer-mapped memory: every CTA
// publishes its pieces of the local gradient in the rank's staging buffer and raises a per-(rank, CTA) flag in every
// peer's flag pad (st.release.sys over NVLink); rank r then owns slice r of the block: its CTAs wait for the peers'
// flags (relaxed polls, one acquire), pull their piece of that slice from every rank with 128-bit loads through
// NVLink / NVSwitch, add the `world` values in rank order and WRITE the sum into every rank's buffer (second flag);
// finally every CTA applies the update from the broadcast sums.  Each rank moves 2 (world-1)/world of the block over
// NVLink (round 1: every rank pulled the whole block from every peer, (world-1) x, measured ~14 us per peer); every
// element is summed by exactly one rank in a fixed order, so the weights stay bit-identical on all ranks.  No
// grid-wide or host synchronisation; the epoch lives in device memory, so the launch is CUDA-graph replayable.
// Staging is double-buffered by epoch parity: a rank can only reach epoch e+2 after every peer finished epoch e
// (see the argument in DESIGN.md s5).
#include <cstdlib>
#include "gs_common.cuh"

namespace {

constexpr int kMaxPeers = 16;
constexpr int kPeerThreads = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ float4 ld_peer_v4(const float4* p) {
    float4 r;      // system-scope relaxed load: never served from a stale L1 line (the addresses are reused every 2 epochs)
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ void wait_flag(const uint32_t* f, uint32_t epoch) {
    // RELAXED polls, ONE acquire after the flag was seen: an acquire load per poll compiles to LDG + CCTL.IVALL, i.e.
    // every poll of every waiting CTA invalidates the L1 of an SM the co-resident gather of the next batch streams through
    while ((int32_t)(ld_relaxed_sys(f) - epoch) < 0) __nanosleep(100);
    fence_acq_rel_sys();
}

// Per-rank symmetric buffer (floats): in[2][n] | out[2][n] | flagsA[world * G] | flagsB[world * G]   (G = gridDim.x)
// Two-phase exchange, so that each rank moves 2 * (world-1)/world of the block over NVLink instead of (world-1) x:
//   1. publish    every CTA copies its pieces of the local gradient into in[parity]             flag A -> all peers
//   2. reduce     rank r owns slice r of the block: CTA b pulls piece (r, b) of every rank's `in`, adds the
//                 `world` values in rank order and writes the sum into EVERY rank's out[parity]  flag B -> all peers
//   3. update     every CTA applies p -= lr * out[parity] to its pieces of all slices
// CTA b of every rank handles piece b of every slice, so the only cross-rank dependencies are (peer q, CTA b) flags.
// Each element is summed by exactly one rank in a fixed order and broadcast: the weights stay bit-identical on all ranks.
// (Round 1 had every rank pull the whole block from every peer: measured ~14 us per peer, 100 us of the step at 8 GPUs.)
// state[0] = last completed epoch, state[1] = CTA ticket of the running launch
__global__ void __launch_bounds__(kPeerThreads)
allreduce_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n4, float lr,
                     float* const* __restrict__ stage, uint32_t* const* __restrict__ flags,
                     int rank, int world, uint32_t* __restrict__ state) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(state) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const int G = gridDim.x, b = blockIdx.x;
    const int64_t per_rank = (n4 + world - 1) / world;         // float4 per slice
    const int64_t per_cta = (per_rank + G - 1) / G;            // float4 per piece
    const int64_t in_off = (int64_t)(epoch & 1u) * n4, out_off = (2 + (int64_t)(epoch & 1u)) * n4;
    const int64_t piece_lo = (int64_t)b * per_cta;
    const int64_t piece_n = max((int64_t)0, min(per_cta, per_rank - piece_lo));
    const int64_t flag_b = (int64_t)world * G;                 // flagsB follows flagsA

    // 1. publish my pieces of every slice
    float4* my_in = reinterpret_cast<float4*>(stage[rank]) + in_off;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int s = 0; s < world; ++s)
        for (int64_t i = threadIdx.x; i < piece_n; i += kPeerThreads) {
            const int64_t e = s * per_rank + piece_lo + i;
            if (e < n4) my_in[e] = g4[e];
        }
    __syncthreads();
    // the release of a flag store is cumulative over the CTA's writes ordered before it by the barrier
    if (threadIdx.x < world) {
        st_release_sys(flags[threadIdx.x] + (int64_t)rank * G + b, epoch);
        wait_flag(flags[rank] + (int64_t)threadIdx.x * G + b, epoch);
    }
    __syncthreads();
    // 2. reduce piece b of MY slice over all ranks (rank order), broadcast the sum
    for (int64_t i = threadIdx.x; i < piece_n; i += kPeerThreads) {
        const int64_t e = (int64_t)rank * per_rank + piece_lo + i;
        if (e >= n4) break;
        float4 v[kMaxPeers];                                   // all peers' loads in flight before the first is used
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < world) v[q] = ld_peer_v4(reinterpret_cast<const float4*>(stage[q]) + in_off + e);
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < world) { sum.x += v[q].x; sum.y += v[q].y; sum.z += v[q].z; sum.w += v[q].w; }
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < world) reinterpret_cast<float4*>(stage[q])[out_off + e] = sum;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        st_release_sys(flags[threadIdx.x] + flag_b + (int64_t)rank * G + b, epoch);
        wait_flag(flags[rank] + flag_b + (int64_t)threadIdx.x * G + b, epoch);
    }
    __syncthreads();
    // 3. update my pieces of every slice from the broadcast sums (written by the peers: read past the L1)
    const float4* my_out = reinterpret_cast<const float4*>(stage[rank]) + out_off;
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int s = 0; s < world; ++s)
        for (int64_t i = threadIdx.x; i < piece_n; i += kPeerThreads) {
            const int64_t e = s * per_rank + piece_lo + i;
            if (e >= n4) break;
            const float4 sum = ld_peer_v4(my_out + e);
            float4 w = p4[e];
            w.x = w.x - lr * sum.x; w.y = w.y - lr * sum.y; w.z = w.z - lr * sum.z; w.w = w.w - lr * sum.w;
            p4[e] = w;
        }
    // 4. the last CTA closes the epoch
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(&state[1], 1u);
        if (t == gridDim.x - 1) {
            state[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(state) = epoch;
        }
    }
}

}  // namespace

extern "C" int32_t gs_allreduce_sgd_blocks(int64_t n) {
    // one float4 per thread per pass and at most one CTA per SM: the pull is NVLink-latency bound (a few
    // hundred KB in total), so width, not depth; the CTAs spin only while the slowest peer catches up
    int64_t b = (n / 4 + kPeerThreads - 1) / kPeerThreads;
    if (b < 1) b = 1;
    static int cap = 0;                       // GSAGE_AR_BLOCKS: experiment knob (every rank must use the same value)
    if (cap == 0) { cap = getenv("GSAGE_AR_BLOCKS") ? atoi(getenv("GSAGE_AR_BLOCKS")) : GS_NUM_SMS; if (cap < 1) cap = 1; }
    if (b > cap) b = cap;
    return (int32_t)b;
}

extern "C" int gs_allreduce_sgd(float* p, const float* g, int64_t n, float lr,
                                float* const* stage_ptrs, uint32_t* const* flag_ptrs,
                                int32_t rank, int32_t world, uint32_t* state, void* stream) {
    if (!p || !g || !stage_ptrs || !flag_ptrs || !state || n <= 0 || (n & 3)) return GS_EINVAL;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return GS_EINVAL;
    if (!gs_aligned16(p) || !gs_aligned16(g)) return GS_EALIGN;
    GS_PREFER_SMEM(allreduce_sgd_kernel);
    allreduce_sgd_kernel<<<gs_allreduce_sgd_blocks(n), kPeerThreads, 0, (cudaStream_t)stream>>>(
        p, g, n / 4, lr, stage_ptrs, flag_ptrs, rank, world, state);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
