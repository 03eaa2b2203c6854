// Shared device/host helpers for libgsage_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gsage.h"

// SM count of the current device (B200: 148 = 2 dies x 74), queried once; 148 when no device is visible (the sizing
// helpers of the C ABI are callable without a GPU)
static inline int gs_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else { (void)cudaGetLastError(); n = 148; }
    }
    return n;
}
#define GS_NUM_SMS gs_num_sms()

#define GS_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

// Kernels that may share an SM with a tcgen05 GEMM CTA (198 KB of shared memory) or the fused head
// (175 KB) must not pull the SM's L1/shared split towards L1: the split can only change on an idle
// SM, so a resident zero-smem kernel with the default carveout keeps the big-smem CTAs out until it
// drains (measured: the layer-1 GEMM waited 130 us for the gather, profiles/README.md).
#define GS_PREFER_SMEM(kernel)                                                                        \
    do {                                                                                              \
        static bool done__ = false;                                                                   \
        if (!done__) {                                                                                \
            cudaFuncSetAttribute((kernel), cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                 (int)cudaSharedmemCarveoutMaxShared);                                \
            done__ = true;                                                                            \
        }                                                                                             \
    } while (0)

static inline bool gs_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ int gs_row_count(int n_max, const int32_t* n_dev) {
    if (n_dev == nullptr) return n_max;
    int n = __ldg(n_dev);
    return n < n_max ? n : n_max;
}

// 128-bit read-only load that does not allocate in L1 (streaming feature rows).
__device__ __forceinline__ float4 gs_ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void gs_red_add_v4(float* addr, float4 v) {
    // sm_90+: vectorised fp32 reduction, one L2 atomic transaction per 16 bytes
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float gs_apply_act(float x, int act) {
    if (act == GS_ACT_RELU) return x > 0.f ? x : 0.f;
    if (act == GS_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
    return x;
}

// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float gs_act_grad(float y, int act) {
    if (act == GS_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == GS_ACT_SIGMOID) return y * (1.f - y);
    return 1.f;
}
