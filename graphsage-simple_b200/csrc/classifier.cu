// K5: classifier + softmax cross-entropy, forward and backward in two launches.
// Replaces `scores = self.weight.mm(embeds).t()` and nn.CrossEntropyLoss()(scores, labels)
// (graphsage/model.py:57, 62-69) and their autograd backward (model.py:249).
#include "gs_common.cuh"

namespace {

constexpr int kRowsPerBlock = 16;     // one warp per target row
constexpr int kMaxClsPerLane = 4;     // num_classes <= 128

// ws layout: dl[n, C] (scaled softmax - onehot), then loss_i[n]
// One warp per target row, 16 rows per block.  The classifier weight is staged once per block in
// shared memory with row stride d+1 (conflict-free when each lane walks its own class row).
__global__ void __launch_bounds__(kRowsPerBlock * 32)
xent_rows_kernel(const float* __restrict__ h, int64_t ld_h, const float* __restrict__ wc, int64_t ld_wc,
                 const int64_t* __restrict__ labels, int d, int C, int n, float gscale,
                 float* __restrict__ logits, int64_t ld_logits, float* __restrict__ gh, int64_t ld_gh,
                 float* __restrict__ ws) {
    extern __shared__ float smem[];
    float* s_wc = smem;                               // [C][d+1]
    float* s_h = smem + (size_t)C * (d + 1);          // [rows][d]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if ((d & 3) == 0 && (ld_wc & 3) == 0) {           // 128-bit global reads of the weight
        const int d4 = d >> 2;
        for (int e = threadIdx.x; e < C * d4; e += blockDim.x) {
            const int c = e / d4, k = (e - c * d4) * 4;
            const float4 v = *reinterpret_cast<const float4*>(wc + (int64_t)c * ld_wc + k);
            float* dst = s_wc + c * (d + 1) + k;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    } else {
        for (int e = threadIdx.x; e < C * d; e += blockDim.x) {
            const int c = e / d, k = e - c * d;
            s_wc[c * (d + 1) + k] = wc[(int64_t)c * ld_wc + k];
        }
    }
    const int row = blockIdx.x * kRowsPerBlock + w;
    float* hrow = s_h + w * d;
    if (row < n)
        for (int k = lane; k < d; k += 32) hrow[k] = h[(int64_t)row * ld_h + k];
    __syncthreads();
    if (row >= n) return;

    float z[kMaxClsPerLane];
    float zmax = -INFINITY;
#pragma unroll
    for (int q = 0; q < kMaxClsPerLane; ++q) {
        const int c = lane + 32 * q;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;          // 4 independent chains
        if (c < C) {
            const float* wr = s_wc + c * (d + 1);
            int k = 0;
            for (; k + 4 <= d; k += 4) {
                a0 = fmaf(hrow[k], wr[k], a0); a1 = fmaf(hrow[k + 1], wr[k + 1], a1);
                a2 = fmaf(hrow[k + 2], wr[k + 2], a2); a3 = fmaf(hrow[k + 3], wr[k + 3], a3);
            }
            for (; k < d; ++k) a0 = fmaf(hrow[k], wr[k], a0);
        }
        const float acc = (a0 + a1) + (a2 + a3);
        if (c < C) {
            zmax = fmaxf(zmax, acc);
            if (logits) logits[(int64_t)row * ld_logits + c] = acc;
        }
        z[q] = acc;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxClsPerLane; ++q)
        if (lane + 32 * q < C) sum += expf(z[q] - zmax);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = zmax + logf(sum);
    const int y = (int)labels[row];
    float picked = 0.f;
    float dl[kMaxClsPerLane];
#pragma unroll
    for (int q = 0; q < kMaxClsPerLane; ++q) {
        const int c = lane + 32 * q;
        dl[q] = 0.f;
        if (c < C) {
            const float p = expf(z[q] - lse);
            dl[q] = (p - (c == y ? 1.f : 0.f)) * gscale;
            if (c == y) picked = z[q];
            ws[(int64_t)row * C + c] = dl[q];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) picked += __shfl_xor_sync(0xffffffffu, picked, o);
    if (lane == 0) ws[(int64_t)n * C + row] = lse - picked;
    if (gh != nullptr) {
        // gh[row, k] = sum_c dl[c] * wc[c, k]: lanes across k, classes broadcast by shuffle
        for (int k0 = 0; k0 < d; k0 += 128) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < kMaxClsPerLane; ++q) {
                const int cmax = min(32, C - 32 * q);
                for (int l = 0; l < cmax; ++l) {
                    const float dv = __shfl_sync(0xffffffffu, dl[q], l);
                    const float* wr = s_wc + (l + 32 * q) * (d + 1) + k0 + lane;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (k0 + lane + 32 * u < d) acc[u] = fmaf(dv, wr[32 * u], acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + lane + 32 * u < d) gh[(int64_t)row * ld_gh + k0 + lane + 32 * u] = acc[u];
        }
    }
}

// gwc[c, :] = sum_i dl[i, c] * h[i, :] in two fixed-order stages: grid (C, S) partial sums over
// row slices, then one pass adding the S partials (which also reduces the per-row losses).
constexpr int kWgradSplits = 16;

__global__ void __launch_bounds__(128)
xent_wgrad_partial_kernel(const float* __restrict__ h, int64_t ld_h, int d, int C, int n,
                          const float* __restrict__ ws, float* __restrict__ part) {
    const int c = blockIdx.x, z = blockIdx.y;
    const int per = (n + kWgradSplits - 1) / kWgradSplits;
    const int i0 = z * per, i1 = min(n, i0 + per);
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float acc = 0.f;
        int i = i0;
        for (; i + 4 <= i1; i += 4) {
            const float w0 = ws[(int64_t)i * C + c], w1 = ws[(int64_t)(i + 1) * C + c];
            const float w2 = ws[(int64_t)(i + 2) * C + c], w3 = ws[(int64_t)(i + 3) * C + c];
            const float h0 = h[(int64_t)i * ld_h + k], h1 = h[(int64_t)(i + 1) * ld_h + k];
            const float h2 = h[(int64_t)(i + 2) * ld_h + k], h3 = h[(int64_t)(i + 3) * ld_h + k];
            acc = fmaf(w0, h0, acc); acc = fmaf(w1, h1, acc); acc = fmaf(w2, h2, acc); acc = fmaf(w3, h3, acc);
        }
        for (; i < i1; ++i) acc = fmaf(ws[(int64_t)i * C + c], h[(int64_t)i * ld_h + k], acc);
        part[((int64_t)z * C + c) * d + k] = acc;
    }
}

__global__ void __launch_bounds__(256)
xent_finish_kernel(const float* __restrict__ part, const float* __restrict__ ws, int d, int C, int n,
                   float* __restrict__ gwc, int64_t ld_gwc, float* __restrict__ loss) {
    if (gwc != nullptr) {
        const int e = blockIdx.x * blockDim.x + threadIdx.x;
        if (e < C * d) {
            float s = 0.f;
#pragma unroll
            for (int z = 0; z < kWgradSplits; ++z) s += part[(int64_t)z * C * d + e];
            const int c = e / d, k = e - c * d;
            gwc[(int64_t)c * ld_gwc + k] = s;
        }
    }
    if (blockIdx.x == 0 && loss != nullptr) {
        __shared__ float red[256];
        float s = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += ws[(int64_t)n * C + i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) loss[0] = red[0] / (float)n;
    }
}

}  // namespace

extern "C" int64_t gs_classifier_ws_floats(int32_t n, int32_t d, int32_t num_classes) {
    return (int64_t)n * num_classes + n + (int64_t)kWgradSplits * num_classes * d;
}

extern "C" int gs_classifier_xent(const float* h, int64_t ld_h, const float* wc, int64_t ld_wc,
                                  const int64_t* labels, int32_t d, int32_t num_classes, int32_t n,
                                  float grad_scale, float* logits, int64_t ld_logits, float* loss,
                                  float* gh, int64_t ld_gh, float* gwc, int64_t ld_gwc,
                                  float* ws, void* stream) {
    if (!h || !wc || !labels || !ws || d <= 0 || num_classes <= 0 || n <= 0) return GS_EINVAL;
    if (num_classes > 32 * kMaxClsPerLane) return GS_ENOSUP;
    const size_t smem = ((size_t)num_classes * (d + 1) + (size_t)kRowsPerBlock * d) * sizeof(float);
    if (smem > 200 * 1024) return GS_ENOSUP;
    cudaStream_t s = (cudaStream_t)stream;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(xent_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    xent_rows_kernel<<<(n + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock * 32, smem, s>>>(
        h, ld_h, wc, ld_wc, labels, d, num_classes, n, grad_scale / (float)n, logits, ld_logits, gh, ld_gh, ws);
    GS_LAUNCH_CHECK();
    float* part = ws + (int64_t)n * num_classes + n;
    if (gwc != nullptr) {
        xent_wgrad_partial_kernel<<<dim3(num_classes, kWgradSplits), 128, 0, s>>>(h, ld_h, d, num_classes, n, ws, part);
        GS_LAUNCH_CHECK();
    }
    xent_finish_kernel<<<(num_classes * d + 255) / 256, 256, 0, s>>>(part, ws, d, num_classes, n, gwc, ld_gwc, loss);
    GS_LAUNCH_CHECK();
    return GS_OK;
}
