// fp32 SIMT GEMM family used by the encoder/classifier where the shapes are too small or
// too ragged for the tcgen05 path, and as its exact-fp32 fallback.
// Replaces `self.weight.mm(combined.t())` + relu/sigmoid (graphsage/encoders.py:58-61) and
// the autograd MmBackward pairs reached from loss.backward() (graphsage/model.py:249).
#include "gs_common.cuh"

namespace {

constexpr int BK = 16;

struct GemmArgs {
    const float* a; int64_t lda;     // A(m,k): A_MCONTIG ? a[k*lda + m] : a[m*lda + k]
    const float* b; int64_t ldb;     // B(k,n): B_NCONTIG ? b[k*ldb + n] : b[n*ldb + k]
    float* c; int64_t ldc;           // C[m*ldc + n]  (+ z * c_split_stride for split-K partials)
    int M, N, K;
    const int32_t* n_dev;            // device override (<= the host value) ...
    int dev_dim;                     // ... of M (0) or K (1); ignored when n_dev == nullptr
    int act;
    int k_per_split;                 // multiple of BK
    int64_t c_split_stride;
};

template <int BM, int BN, int TM, int TN, bool A_MCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(GemmArgs g) {
    constexpr int THREADS = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    int M = g.M, K = g.K;
    if (g.n_dev != nullptr) {
        const int nd = __ldg(g.n_dev);
        if (g.dev_dim == 0) M = min(M, nd); else K = min(K, nd);
    }
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= M) return;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(K, k_begin + g.k_per_split);
    float* cbase = g.c + (int64_t)blockIdx.z * g.c_split_stride;

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    // ---- global -> register staging (each thread moves A_LD + B_LD float4 per k-step)
    constexpr int A_LD = BM * BK / 4 / THREADS, B_LD = BN * BK / 4 / THREADS;
    static_assert(A_LD >= 1 && B_LD >= 1, "tile too small for the thread count");
    float4 ra[A_LD], rb[B_LD];

    auto load_a = [&](int k0) {
#pragma unroll
        for (int u = 0; u < A_LD; ++u) {
            const int f = tid + u * THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_MCONTIG) {
                const int kk = f / (BM / 4), m = m0 + (f % (BM / 4)) * 4, k = k0 + kk;
                if (k < k_end) {
                    const float* p = g.a + (int64_t)k * g.lda + m;
                    if (m + 4 <= M) v = *reinterpret_cast<const float4*>(p);
                    else { if (m < M) v.x = p[0]; if (m + 1 < M) v.y = p[1]; if (m + 2 < M) v.z = p[2]; }
                }
            } else {
                const int r = f / (BK / 4), m = m0 + r, k = k0 + (f % (BK / 4)) * 4;
                if (m < M) {
                    const float* p = g.a + (int64_t)m * g.lda + k;
                    if (k + 4 <= k_end) v = *reinterpret_cast<const float4*>(p);
                    else { if (k < k_end) v.x = p[0]; if (k + 1 < k_end) v.y = p[1]; if (k + 2 < k_end) v.z = p[2]; }
                }
            }
            ra[u] = v;
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int u = 0; u < B_LD; ++u) {
            const int f = tid + u * THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_NCONTIG) {
                const int kk = f / (BN / 4), n = n0 + (f % (BN / 4)) * 4, k = k0 + kk;
                if (k < k_end) {
                    const float* p = g.b + (int64_t)k * g.ldb + n;
                    if (n + 4 <= g.N) v = *reinterpret_cast<const float4*>(p);
                    else { if (n < g.N) v.x = p[0]; if (n + 1 < g.N) v.y = p[1]; if (n + 2 < g.N) v.z = p[2]; }
                }
            } else {
                const int r = f / (BK / 4), n = n0 + r, k = k0 + (f % (BK / 4)) * 4;
                if (n < g.N) {
                    const float* p = g.b + (int64_t)n * g.ldb + k;
                    if (k + 4 <= k_end) v = *reinterpret_cast<const float4*>(p);
                    else { if (k < k_end) v.x = p[0]; if (k + 1 < k_end) v.y = p[1]; if (k + 2 < k_end) v.z = p[2]; }
                }
            }
            rb[u] = v;
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int u = 0; u < A_LD; ++u) {
            const int f = tid + u * THREADS;
            if (A_MCONTIG) {
                const int kk = f / (BM / 4), m = (f % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[buf][kk][m]) = ra[u];
            } else {
                const int r = f / (BK / 4), kk = (f % (BK / 4)) * 4;
                As[buf][kk][r] = ra[u].x; As[buf][kk + 1][r] = ra[u].y;
                As[buf][kk + 2][r] = ra[u].z; As[buf][kk + 3][r] = ra[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < B_LD; ++u) {
            const int f = tid + u * THREADS;
            if (B_NCONTIG) {
                const int kk = f / (BN / 4), n = (f % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][kk][n]) = rb[u];
            } else {
                const int r = f / (BK / 4), kk = (f % (BK / 4)) * 4;
                Bs[buf][kk][r] = rb[u].x; Bs[buf][kk + 1][r] = rb[u].y;
                Bs[buf][kk + 2][r] = rb[u].z; Bs[buf][kk + 3][r] = rb[u].w;
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // thread micro-tile: rows ty*4+{0..3} (+BM/2 when TM==8), cols tx*4+{0..3} (+BN/2 when TN==8)
    if (k_begin < k_end) {
        load_a(k_begin); load_b(k_begin);
        stage(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool more = k0 + BK < k_end;
            if (more) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float av[TM], bv[TN];
#pragma unroll
                for (int h = 0; h < TM / 4; ++h) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][h * (BM / 2) + ty * 4]);
                    av[h * 4] = t.x; av[h * 4 + 1] = t.y; av[h * 4 + 2] = t.z; av[h * 4 + 3] = t.w;
                }
#pragma unroll
                for (int h = 0; h < TN / 4; ++h) {
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][h * (BN / 2) + tx * 4]);
                    bv[h * 4] = t.x; bv[h * 4 + 1] = t.y; bv[h * 4 + 2] = t.z; bv[h * 4 + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (more) {
                stage(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i % 4);
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int n = n0 + h * (BN / 2) + tx * 4;
            float4 v = make_float4(gs_apply_act(acc[i][h * 4], g.act), gs_apply_act(acc[i][h * 4 + 1], g.act),
                                   gs_apply_act(acc[i][h * 4 + 2], g.act), gs_apply_act(acc[i][h * 4 + 3], g.act));
            float* p = cbase + (int64_t)m * g.ldc + n;
            if (n + 4 <= g.N) *reinterpret_cast<float4*>(p) = v;
            else { if (n < g.N) p[0] = v.x; if (n + 1 < g.N) p[1] = v.y; if (n + 2 < g.N) p[2] = v.z; }
        }
    }
}

// dz = gh * act'(h); rows >= n are written as zeros so later reductions over n_max are safe
__global__ void act_grad_kernel(const float* __restrict__ h, int64_t ld_h, const float* __restrict__ gh, int64_t ld_gh,
                                int d, int act, int n_max, const int32_t* __restrict__ n_dev, float* __restrict__ dz,
                                int ld_dz) {
    const int n = gs_row_count(n_max, n_dev);
    const int64_t total = (int64_t)n_max * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e - (int64_t)i * d);
        dz[(int64_t)i * ld_dz + j] = i < n ? gh[(int64_t)i * ld_gh + j] * gs_act_grad(h[(int64_t)i * ld_h + j], act) : 0.f;
    }
}

// out[m, n] = sum_z ws[z][m][n]   (fixed order -> deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t stride, int64_t ld_ws,
                                     int M, int N, float* __restrict__ out, int64_t ld_out) {
    const int64_t total = (int64_t)M * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / N), n = (int)(e - (int64_t)m * N);
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * stride + (int64_t)m * ld_ws + n];
        out[(int64_t)m * ld_out + n] = s;
    }
}

template <bool A_MCONTIG, bool B_NCONTIG>
int launch_gemm(GemmArgs g, int splits, cudaStream_t s) {
    // big tiles when both extents fill them, small tiles otherwise
    if (g.M > 64 && g.N > 64 && (int64_t)g.M * g.N >= 128 * 128 * 32) {
        dim3 grid((g.N + 127) / 128, (g.M + 127) / 128, splits);
        gemm_kernel<128, 128, 8, 8, A_MCONTIG, B_NCONTIG><<<grid, 256, 0, s>>>(g);
    } else {
        dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, splits);
        gemm_kernel<64, 64, 4, 4, A_MCONTIG, B_NCONTIG><<<grid, 256, 0, s>>>(g);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GS_OK : (int)e;
}

int bwd_splits(int n_max, int k_in, int d_out) {
    const int tiles = ((k_in + 127) / 128) * ((d_out + 127) / 128);
    int s = (2 * GS_NUM_SMS + tiles - 1) / tiles;
    const int cap = (n_max + 4 * BK - 1) / (4 * BK);
    if (s > cap) s = cap;
    if (s < 1) s = 1;
    return s;
}

}  // namespace

extern "C" int gs_encoder_fwd(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                              int32_t k_in, int32_t d_out, int32_t act,
                              int32_t n_max, const int32_t* n_dev,
                              float* h, int64_t ld_h, void* stream) {
    if (!x || !w || !h || k_in <= 0 || d_out <= 0 || n_max < 0) return GS_EINVAL;
    if (!gs_aligned16(x) || !gs_aligned16(w) || !gs_aligned16(h) || (ld_x & 3) || (ld_w & 3) || (ld_h & 3)) return GS_EALIGN;
    if (ld_x < k_in || ld_w < k_in || ld_h < d_out) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    GemmArgs g{x, ld_x, w, ld_w, h, ld_h, n_max, d_out, k_in, n_dev, 0, act, ((k_in + BK - 1) / BK) * BK, 0};
    return launch_gemm<false, false>(g, 1, (cudaStream_t)stream);
}

extern "C" int64_t gs_encoder_bwd_ws_floats(int32_t n_max, int32_t k_in, int32_t d_out) {
    return (int64_t)bwd_splits(n_max, k_in, d_out) * d_out * ((k_in + 3) & ~3);
}

extern "C" int gs_encoder_bwd(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                              const float* h, int64_t ld_h, const float* gh, int64_t ld_gh,
                              int32_t k_in, int32_t d_out, int32_t act,
                              int32_t n_max, const int32_t* n_dev,
                              float* dz, float* gw, int64_t ld_gw, float* gx, int64_t ld_gx,
                              float* ws, void* stream) {
    if (!x || !w || !h || !gh || !dz || !gw || !ws || k_in <= 0 || d_out <= 0 || n_max < 0) return GS_EINVAL;
    if (!gs_aligned16(x) || !gs_aligned16(w) || !gs_aligned16(dz) || !gs_aligned16(ws) || (ld_x & 3) || (ld_w & 3) ||
        (gx && (!gs_aligned16(gx) || (ld_gx & 3))))
        return GS_EALIGN;
    const int ld_dz = (d_out + 3) & ~3;
    if (n_max == 0) return GS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    {
        const int64_t total = (int64_t)n_max * d_out;
        int64_t blocks = (total + 255) / 256;
        if (blocks > GS_NUM_SMS * 8) blocks = GS_NUM_SMS * 8;
        act_grad_kernel<<<(int)blocks, 256, 0, s>>>(h, ld_h, gh, ld_gh, d_out, act, n_max, n_dev, dz, ld_dz);
        GS_LAUNCH_CHECK();
    }
    // gw[d_out, k_in] = dz^T . x   (reduction over the n rows, split-K with ordered partials)
    const int splits = bwd_splits(n_max, k_in, d_out);
    int kper = (n_max + splits - 1) / splits;
    kper = ((kper + BK - 1) / BK) * BK;
    const bool direct = splits == 1 && (ld_gw & 3) == 0 && gs_aligned16(gw);
    const int64_t ld_ws = (k_in + 3) & ~3;          // keeps every partial row 16-B aligned
    GemmArgs g{dz, ld_dz, x, ld_x, direct ? gw : ws, direct ? ld_gw : ld_ws, d_out, k_in, n_max, n_dev, 1,
               GS_ACT_NONE, kper, (int64_t)d_out * ld_ws};
    int rc = launch_gemm<true, true>(g, splits, s);
    if (rc) return rc;
    if (!direct) {
        const int64_t total = (int64_t)d_out * k_in;
        int64_t blocks = (total + 255) / 256;
        if (blocks > GS_NUM_SMS * 8) blocks = GS_NUM_SMS * 8;
        splitk_reduce_kernel<<<(int)blocks, 256, 0, s>>>(ws, splits, (int64_t)d_out * ld_ws, ld_ws, d_out, k_in, gw, ld_gw);
        GS_LAUNCH_CHECK();
    }
    if (gx) {   // gx[n, k_in] = dz . w
        GemmArgs gd{dz, ld_dz, w, ld_w, gx, ld_gx, n_max, k_in, d_out, n_dev, 0, GS_ACT_NONE, ((d_out + BK - 1) / BK) * BK, 0};
        rc = launch_gemm<false, true>(gd, 1, s);
        if (rc) return rc;
    }
    return GS_OK;
}

// Input gradient only: gx[n, k_in] = (gh * act'(h)) . w   (used when the weight gradient runs on tcgen05)
extern "C" int gs_encoder_dgrad(const float* w, int64_t ld_w, const float* h, int64_t ld_h,
                                const float* gh, int64_t ld_gh, int32_t k_in, int32_t d_out, int32_t act,
                                int32_t n_max, const int32_t* n_dev, float* dz, float* gx, int64_t ld_gx,
                                void* stream) {
    if (!w || !h || !gh || !dz || !gx || k_in <= 0 || d_out <= 0 || n_max < 0) return GS_EINVAL;
    if (!gs_aligned16(w) || !gs_aligned16(dz) || !gs_aligned16(gx) || (ld_w & 3) || (ld_gx & 3)) return GS_EALIGN;
    if (n_max == 0) return GS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int ld_dz = (d_out + 3) & ~3;
    const int64_t total = (int64_t)n_max * d_out;
    int64_t blocks = (total + 255) / 256;
    if (blocks > GS_NUM_SMS * 8) blocks = GS_NUM_SMS * 8;
    act_grad_kernel<<<(int)blocks, 256, 0, s>>>(h, ld_h, gh, ld_gh, d_out, act, n_max, n_dev, dz, ld_dz);
    GS_LAUNCH_CHECK();
    GemmArgs gd{dz, ld_dz, w, ld_w, gx, ld_gx, n_max, k_in, d_out, n_dev, 0, GS_ACT_NONE, ((d_out + BK - 1) / BK) * BK, 0};
    return launch_gemm<false, true>(gd, 1, s);
}
