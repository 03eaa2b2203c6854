// Data-parallel gradient all-reduce fused with the SGD update, over NVLink peer memory.
//
// The reference has no distributed code; its update is `optimizer.step()` on local gradients
// (graphsage/model.py:237, 250).  With one process per GPU each rank holds the gradient of its share
// of the global batch; the update every rank must apply is  p -= lr * sum_r g_r.  Instead of an NCCL
// all-reduce followed by an SGD kernel (two launches that cannot sit inside the step's CUDA graph
// without a collective in the capture), ONE kernel does both: every CTA publishes its slice of the
// local gradient in a peer-mapped staging buffer, raises a per-slice flag in every peer's flag pad
// (st.release.sys over NVLink), waits for the peers' flags for the SAME slice (ld.acquire.sys), pulls the
// peers' slices with 128-bit loads through NVLink / NVSwitch, adds them in rank order (bit-identical on
// every rank) and applies the update.  No grid-wide or host synchronisation; the epoch lives in device
// memory, so the launch is CUDA-graph replayable.  Staging is double-buffered by epoch parity: a rank
// can only reach epoch e+2 after every peer finished reading epoch e (see the argument in DESIGN.md s5).
#include "gs_common.cuh"

namespace {

constexpr int kMaxPeers = 16;
constexpr int kPeerThreads = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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

// state[0] = last completed epoch, state[1] = CTA ticket of the running launch
__global__ void __launch_bounds__(kPeerThreads)
allreduce_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n4, float lr,
                     float* const* __restrict__ stage, uint32_t* const* __restrict__ flags,
                     int rank, int world, uint32_t* __restrict__ state) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(state) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = lo + per < n4 ? lo + per : n4;
    const int64_t half = (int64_t)(epoch & 1u) * n4;

    // 1. publish my slice
    float4* mine = reinterpret_cast<float4*>(stage[rank]) + half;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) mine[i] = g4[i];
    __syncthreads();
    // 2. raise this slice's flag at every peer, then wait for every peer's flag for the same slice.
    // The release of the flag store is cumulative over the CTA's copies ordered before it by the barrier, so only the
    // `world` flag-writing threads fence at system scope (all 256 did in round 1).  The wait polls with RELAXED loads
    // and acquires ONCE after it saw the flag: an acquire load per poll compiles to LDG + CCTL.IVALL, i.e. every poll
    // of every waiting CTA invalidated the L1 of an SM that the co-resident gather of the next batch is streaming
    // through -- measured on 8 GPUs as ~14 us of step time per peer (profiles/README.md, round 2).
    if (threadIdx.x < world) {
        st_release_sys(flags[threadIdx.x] + (int64_t)rank * gridDim.x + blockIdx.x, epoch);
        const uint32_t* f = flags[rank] + (int64_t)threadIdx.x * gridDim.x + blockIdx.x;
        while ((int32_t)(ld_relaxed_sys(f) - epoch) < 0) __nanosleep(100);
        fence_acq_rel_sys();
    }
    __syncthreads();
    // 3. pull, add in rank order, update
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) {
        // all peers' loads are issued before the first is consumed: one NVLink round trip, not `world`
        float4 v[kMaxPeers];
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < world) v[q] = ld_peer_v4(reinterpret_cast<const float4*>(stage[q]) + half + i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < world) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
        float4 w = p4[i];
        w.x = w.x - lr * s.x; w.y = w.y - lr * s.y; w.z = w.z - lr * s.z; w.w = w.w - lr * s.w;
        p4[i] = w;
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
    if (b > GS_NUM_SMS) b = GS_NUM_SMS;
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
