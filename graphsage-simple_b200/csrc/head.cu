// Fused "head" of the 2-layer step: layer-2 aggregation + encoder, classifier, softmax
// cross-entropy and the whole backward down to the gradient of the layer-1 outputs, in TWO
// launches instead of fifteen.  Per target this is ~230 kFLOP on ~2 KB of rows -- pure launch
// latency when spread over separate kernels (profiles/README.md, r01 timeline: 136 us of a
// 297 us step), so everything that is row-local runs in one CTA pass with both weight matrices
// resident in shared memory, and the two weight gradients (reductions over the batch) in a second.
//
// Replaces, for the outer layer of the reference (graphsage-simple):
//   MeanAggregator.forward  mask.mm(embed_matrix)                 graphsage/aggregators.py:54-74
//   Encoder.forward         cat([self, neigh]); relu(W.mm(c.t())) graphsage/encoders.py:49-61
//   SupervisedGraphSage     weight.mm(embeds).t(), CrossEntropy   graphsage/model.py:57-69
//   loss.backward()         their autograd backward               graphsage/model.py:249
// Same arithmetic (fp32 FMA) as gs_gather_mean_fwd + gs_encoder_fwd + gs_classifier_xent +
// gs_encoder_bwd + gs_scatter_mean_bwd composed; summation orders differ.
#include "gs_common.cuh"

namespace {

constexpr int kRows = 8;            // targets per CTA (one warp per row in the row-wise phases)
constexpr int kThreads = 256;
constexpr int kD2 = 128;            // layer-2 width this kernel is specialised for
constexpr int kMaxClsPerLane = 4;   // num_classes <= 128

struct HeadArgs {
    const float* h1; int64_t ld_h1; int d1;
    const int32_t* idx; const int32_t* cnt; int width; const int32_t* self_slots;
    const float* w2; int64_t ld_w2; int act2;
    const float* wc; int64_t ld_wc; int C;
    const int64_t* labels; int n; float gscale;          // gscale = grad_scale / n
    float* comb2; int64_t ld_comb2;                       // [n, K2]   (stage B operand)
    float* h2; int64_t ld_h2;                             // [n, 128]  (stage B operand, public output)
    float* logits; int64_t ld_logits;                     // optional
    float* gh1; int64_t ld_gh1;                           // accumulated with vector reductions
    float* dz2;                                           // ws: [n, 128]
    float* dl;                                            // ws: [n, C]
    float* loss_rows;                                     // ws: [n]
};

// ---------------------------------------------------------------------------------------------
// Stage A: one CTA = 8 targets.
//   0  W2, Wc -> shared memory; warp r gathers comb2[r] = [h1[self] | mean_j h1[idx[r,j]]]
//   1  z2 = W2 . comb2, h2 = act(z2)              thread = (output o, K half), 8 rows in registers
//   2  logits, softmax, loss, dl                  warp = row
//   3  gh2 = dl . Wc, dz2 = gh2 * act'(h2)        thread = (o, row half)
//   4  gcomb2 = dz2 . W2                          thread = input column k
//   5  gh1[self] += gcomb2[:, :d1]; gh1[idx[r,j]] += gcomb2[:, d1:] / cnt   warp = row
template <int K2>
__global__ void __maxnreg__(96)
head_rows_kernel(HeadArgs a) {
    constexpr int kPitchW = K2 + 4;                       // conflict-free LDS.128 with one W2 row per lane
    constexpr int kPitchC = kD2 + 1;                      // conflict-free LDS.32 with one Wc row per lane
    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;                                    // [128][K2+4]
    float* sX = sW2 + kD2 * kPitchW;                      // [8][K2]   comb2 tile, later gcomb2 tile
    float* sH = sX + kRows * K2;                          // [8][128]  h2
    float* sDZ = sH + kRows * kD2;                        // [8][128]  K-half partials, later dz2
    float* sDL = sDZ + kRows * kD2;                       // [8][128]  dl (classes)
    float* sWc = sDL + kRows * 128;                       // [C][129]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * kRows;
    const int d1 = a.d1;
    const int neigh_off = K2 - d1;                        // 0 in GCN mode (K2 == d1), d1 in SAGE mode

    // ---- phase 0a: weights -> shared memory (128-bit global reads)
    for (int e = tid; e < kD2 * (K2 / 4); e += kThreads) {
        const int o = e / (K2 / 4), k4 = e - o * (K2 / 4);
        const float4 v = *reinterpret_cast<const float4*>(a.w2 + (int64_t)o * a.ld_w2 + 4 * k4);
        *reinterpret_cast<float4*>(sW2 + o * kPitchW + 4 * k4) = v;
    }
    for (int e = tid; e < a.C * (kD2 / 4); e += kThreads) {
        const int c = e / (kD2 / 4), k4 = e - c * (kD2 / 4);
        const float4 v = *reinterpret_cast<const float4*>(a.wc + (int64_t)c * a.ld_wc + 4 * k4);
        float* dst = sWc + c * kPitchC + 4 * k4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    // ---- phase 0b: warp r builds row r of the comb2 tile (lane owns one float4 of a 128-wide row)
    {
        const int r = warp, t = row0 + r;
        float4 self = make_float4(0.f, 0.f, 0.f, 0.f), acc = self;
        if (t < a.n) {
            const int c = min(a.cnt[t], a.width);
            const float inv = c > 0 ? 1.f / (float)c : 0.f;
            const int32_t* irow = a.idx + (int64_t)t * a.width;
            if (neigh_off)
                self = *reinterpret_cast<const float4*>(a.h1 + (int64_t)a.self_slots[t] * a.ld_h1 + 4 * lane);
            for (int j0 = 0; j0 < c; j0 += 32) {              // d1 == 128: every lane owns one float4
                const int my = (j0 + lane < c) ? irow[j0 + lane] : 0;
                const int lim = min(32, c - j0);
                int j = 0;
                for (; j + 4 <= lim; j += 4) {                // 4 independent row loads in flight
                    float4 v[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        v[b] = *reinterpret_cast<const float4*>(
                            a.h1 + (int64_t)__shfl_sync(0xffffffffu, my, j + b) * a.ld_h1 + 4 * lane);
#pragma unroll
                    for (int b = 0; b < 4; ++b) { acc.x += v[b].x; acc.y += v[b].y; acc.z += v[b].z; acc.w += v[b].w; }
                }
                for (; j < lim; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(
                        a.h1 + (int64_t)__shfl_sync(0xffffffffu, my, j) * a.ld_h1 + 4 * lane);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        }
        if (4 * lane < d1) {
            if (neigh_off) *reinterpret_cast<float4*>(sX + r * K2 + 4 * lane) = self;
            *reinterpret_cast<float4*>(sX + r * K2 + neigh_off + 4 * lane) = acc;
            if (t < a.n) {
                float* crow = a.comb2 + (int64_t)t * a.ld_comb2;
                if (neigh_off) *reinterpret_cast<float4*>(crow + 4 * lane) = self;
                *reinterpret_cast<float4*>(crow + neigh_off + 4 * lane) = acc;
            }
        }
    }
    __syncthreads();

    // ---- phase 1: z2[r][o] = sum_k W2[o][k] * comb2[r][k]
    {
        const int o = tid & (kD2 - 1), kh = tid >> 7;         // two K halves
        constexpr int kHalf = K2 / 2;
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
        const float* wrow = sW2 + o * kPitchW + kh * kHalf;
        const float* xbase = sX + kh * kHalf;
#pragma unroll 4
        for (int k = 0; k < kHalf; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(xbase + r * K2 + k);
                acc[r] = fmaf(w.x, x.x, acc[r]); acc[r] = fmaf(w.y, x.y, acc[r]);
                acc[r] = fmaf(w.z, x.z, acc[r]); acc[r] = fmaf(w.w, x.w, acc[r]);
            }
        }
        if (kh == 1) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) sDZ[r * kD2 + o] = acc[r];
        }
        __syncthreads();
        if (kh == 0) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float h = gs_apply_act(acc[r] + sDZ[r * kD2 + o], a.act2);
                sH[r * kD2 + o] = h;
                if (row0 + r < a.n) a.h2[(int64_t)(row0 + r) * a.ld_h2 + o] = h;
            }
        }
        __syncthreads();
    }

    // ---- phase 2: classifier + softmax cross-entropy, warp = row (same formulas as xent_rows_kernel)
    {
        const int r = warp, t = row0 + r;
        const float* hrow = sH + r * kD2;
        float z[kMaxClsPerLane];
        float zmax = -INFINITY;
#pragma unroll
        for (int q = 0; q < kMaxClsPerLane; ++q) {
            const int c = lane + 32 * q;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (c < a.C) {
                const float* wr = sWc + c * kPitchC;
#pragma unroll 8
                for (int k = 0; k < kD2; k += 4) {
                    a0 = fmaf(hrow[k], wr[k], a0); a1 = fmaf(hrow[k + 1], wr[k + 1], a1);
                    a2 = fmaf(hrow[k + 2], wr[k + 2], a2); a3 = fmaf(hrow[k + 3], wr[k + 3], a3);
                }
            }
            const float acc = (a0 + a1) + (a2 + a3);
            if (c < a.C) {
                zmax = fmaxf(zmax, acc);
                if (a.logits && t < a.n) a.logits[(int64_t)t * a.ld_logits + c] = acc;
            }
            z[q] = acc;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxClsPerLane; ++q)
            if (lane + 32 * q < a.C) sum += expf(z[q] - zmax);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float lse = zmax + logf(sum);
        const int y = t < a.n ? (int)a.labels[t] : -1;
        float picked = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxClsPerLane; ++q) {
            const int c = lane + 32 * q;
            if (c < a.C) {
                const float p = expf(z[q] - lse);
                const float d = t < a.n ? (p - (c == y ? 1.f : 0.f)) * a.gscale : 0.f;
                if (c == y) picked = z[q];
                sDL[r * 128 + c] = d;
                if (t < a.n) a.dl[(int64_t)t * a.C + c] = d;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) picked += __shfl_xor_sync(0xffffffffu, picked, o);
        if (lane == 0 && t < a.n) a.loss_rows[t] = lse - picked;
    }
    __syncthreads();

    // ---- phase 3: gh2[r][o] = sum_c dl[r][c] * Wc[c][o];  dz2 = gh2 * act'(h2)
    {
        const int o = tid & (kD2 - 1), rh = tid >> 7;         // rows rh*4 .. rh*4+3
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < a.C; ++c) {
            const float w = sWc[c * kPitchC + o];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmaf(sDL[(rh * 4 + u) * 128 + c], w, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = rh * 4 + u;
            const float dz = acc[u] * gs_act_grad(sH[r * kD2 + o], a.act2);
            sDZ[r * kD2 + o] = dz;
            if (row0 + r < a.n) a.dz2[(int64_t)(row0 + r) * kD2 + o] = dz;
        }
    }
    __syncthreads();

    // ---- phase 4: gcomb2[r][k] = sum_o dz2[r][o] * W2[o][k]  -> sX
    {
        constexpr int kGroups = kThreads / K2;                // 1 (K2 = 256) or 2 (K2 = 128)
        constexpr int kRpt = kRows / kGroups;                 // rows per thread
        const int k = tid % K2, rg = tid / K2;
        float acc[kRpt];
#pragma unroll
        for (int u = 0; u < kRpt; ++u) acc[u] = 0.f;
#pragma unroll 2
        for (int o = 0; o < kD2; o += 4) {
            const float w0 = sW2[(o + 0) * kPitchW + k], w1 = sW2[(o + 1) * kPitchW + k];
            const float w2 = sW2[(o + 2) * kPitchW + k], w3 = sW2[(o + 3) * kPitchW + k];
#pragma unroll
            for (int u = 0; u < kRpt; ++u) {
                const float4 dz = *reinterpret_cast<const float4*>(sDZ + (rg * kRpt + u) * kD2 + o);
                acc[u] = fmaf(dz.x, w0, acc[u]); acc[u] = fmaf(dz.y, w1, acc[u]);
                acc[u] = fmaf(dz.z, w2, acc[u]); acc[u] = fmaf(dz.w, w3, acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kRpt; ++u) sX[(rg * kRpt + u) * K2 + k] = acc[u];
    }
    __syncthreads();

    // ---- phase 5: scatter-add into gh1 (128-bit reductions at L2), warp = row
    {
        const int r = warp, t = row0 + r;
        if (t < a.n && 4 * lane < d1) {
            const int c = min(a.cnt[t], a.width);
            const float inv = c > 0 ? 1.f / (float)c : 0.f;
            if (neigh_off) {
                const float4 g = *reinterpret_cast<const float4*>(sX + r * K2 + 4 * lane);
                gs_red_add_v4(a.gh1 + (int64_t)a.self_slots[t] * a.ld_gh1 + 4 * lane, g);
            }
            float4 g = *reinterpret_cast<const float4*>(sX + r * K2 + neigh_off + 4 * lane);
            g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
            const int32_t* irow = a.idx + (int64_t)t * a.width;
            for (int j = 0; j < c; ++j) gs_red_add_v4(a.gh1 + (int64_t)irow[j] * a.ld_gh1 + 4 * lane, g);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stage B: the two weight gradients, reductions over the n targets:
//   gw2[o][k] = sum_t dz2[t][o] * comb2[t][k]     (128 x K2)
//   gwc[c][o] = sum_t dl[t][c]  * h2[t][o]        (C x 128)
// Output tiles of 32 x 64, the batch split kSplits ways; every CTA writes its partial tile, the
// last CTA of a tile (ticket) adds the kSplits partials in fixed order -> deterministic.
constexpr int kSplits = 8;
constexpr int kTM = 32, kTN = 64, kTT = 32;               // tile rows, tile cols, batch rows per smem stage

struct WgradArgs {
    const float* dz2; const float* comb2; int64_t ld_comb2; int K2;
    const float* dl; int C; const float* h2; int64_t ld_h2;
    int n;
    float* gw2; int64_t ld_gw2; float* gwc; int64_t ld_gwc;
    float* part;                  // [tiles][kSplits][32*64]
    int32_t* tickets;             // [tiles], zero on entry, re-armed on exit
    const float* loss_rows; float* loss;
    int tiles_w2;                 // number of (32 x 64) tiles of gw2
    int tn_w2;                    // tiles along k of gw2
};

__global__ void __launch_bounds__(256)
head_wgrad_kernel(WgradArgs a) {
    __shared__ __align__(16) float sA[kTT][kTM];
    __shared__ __align__(16) float sB[kTT][kTN];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, split = blockIdx.y;
    // which product / tile
    const float* A; const float* Bm; int64_t lda, ldb; int M, N; float* out; int64_t ldo; int m0, n0;
    if (tile < a.tiles_w2) {
        A = a.dz2; lda = kD2; M = kD2; Bm = a.comb2; ldb = a.ld_comb2; N = a.K2; out = a.gw2; ldo = a.ld_gw2;
        m0 = (tile / a.tn_w2) * kTM; n0 = (tile % a.tn_w2) * kTN;
    } else {
        const int tt = tile - a.tiles_w2;
        A = a.dl; lda = a.C; M = a.C; Bm = a.h2; ldb = a.ld_h2; N = kD2; out = a.gwc; ldo = a.ld_gwc;
        m0 = (tt / (kD2 / kTN)) * kTM; n0 = (tt % (kD2 / kTN)) * kTN;
    }
    const int per = (a.n + kSplits - 1) / kSplits;
    const int t0 = split * per, t1 = min(a.n, t0 + per);
    // thread tile: 2 (m) x 4 (n)
    const int tm = (tid >> 4) * 2, tn = (tid & 15) * 4;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int tb = t0; tb < t1; tb += kTT) {
        // stage kTT batch rows of both operands (zero-filled outside the matrices)
        for (int e = tid; e < kTT * kTM; e += 256) {
            const int tt = e / kTM, m = e - tt * kTM;
            const int t = tb + tt;
            sA[tt][m] = (t < t1 && m0 + m < M) ? A[(int64_t)t * lda + m0 + m] : 0.f;
        }
        for (int e = tid; e < kTT * kTN / 4; e += 256) {
            const int tt = e / (kTN / 4), c4 = e - tt * (kTN / 4);
            const int t = tb + tt;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < t1 && n0 + 4 * c4 < N) v = *reinterpret_cast<const float4*>(Bm + (int64_t)t * ldb + n0 + 4 * c4);
            *reinterpret_cast<float4*>(&sB[tt][4 * c4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int tt = 0; tt < kTT; ++tt) {
            const float2 av = *reinterpret_cast<const float2*>(&sA[tt][tm]);
            const float4 bv = *reinterpret_cast<const float4*>(&sB[tt][tn]);
            acc[0][0] = fmaf(av.x, bv.x, acc[0][0]); acc[0][1] = fmaf(av.x, bv.y, acc[0][1]);
            acc[0][2] = fmaf(av.x, bv.z, acc[0][2]); acc[0][3] = fmaf(av.x, bv.w, acc[0][3]);
            acc[1][0] = fmaf(av.y, bv.x, acc[1][0]); acc[1][1] = fmaf(av.y, bv.y, acc[1][1]);
            acc[1][2] = fmaf(av.y, bv.z, acc[1][2]); acc[1][3] = fmaf(av.y, bv.w, acc[1][3]);
        }
        __syncthreads();
    }
    float* mypart = a.part + ((int64_t)tile * kSplits + split) * (kTM * kTN);
    *reinterpret_cast<float4*>(mypart + tm * kTN + tn) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
    *reinterpret_cast<float4*>(mypart + (tm + 1) * kTN + tn) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int ticket = atomicAdd(&a.tickets[tile], 1);
        s_last = (ticket == kSplits - 1);
        if (s_last) a.tickets[tile] = 0;                  // re-arm for the next launch
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float* base = a.part + (int64_t)tile * kSplits * (kTM * kTN);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int z = 0; z < kSplits; ++z) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(base + z * (kTM * kTN) + (tm + u) * kTN + tn));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            const int m = m0 + tm + u;
            if (m < M) {
                float* dst = out + (int64_t)m * ldo + n0 + tn;
                if (n0 + tn + 0 < N) dst[0] = s.x;
                if (n0 + tn + 1 < N) dst[1] = s.y;
                if (n0 + tn + 2 < N) dst[2] = s.z;
                if (n0 + tn + 3 < N) dst[3] = s.w;
            }
        }
    }
    // mean loss: fixed-order tree over the per-row losses (block (0,0) only)
    if (tile == 0 && split == 0 && a.loss != nullptr) {
        __shared__ float red[256];
        float s = 0.f;
        for (int i = tid; i < a.n; i += 256) s += a.loss_rows[i];
        red[tid] = s;
        __syncthreads();
        for (int o = 128; o; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) a.loss[0] = red[0] / (float)a.n;
    }
}

size_t head_smem_bytes(int K2, int C) {
    return sizeof(float) * ((size_t)kD2 * (K2 + 4) + (size_t)kRows * K2 + 3 * (size_t)kRows * 128 + (size_t)C * (kD2 + 1));
}

int wgrad_tiles(int K2, int C, int* tiles_w2, int* tn_w2) {
    *tn_w2 = (K2 + kTN - 1) / kTN;
    *tiles_w2 = (kD2 / kTM) * *tn_w2;
    return *tiles_w2 + ((C + kTM - 1) / kTM) * (kD2 / kTN);
}

}  // namespace

extern "C" int gs_head_supported(int32_t d1, int32_t k2_in, int32_t d2, int32_t num_classes) {
    if (d2 != kD2 || d1 != 128) return 0;
    if (k2_in != d1 && k2_in != 2 * d1) return 0;
    if (num_classes < 1 || num_classes > 32 * kMaxClsPerLane) return 0;
    return head_smem_bytes(k2_in, num_classes) <= 227 * 1024 ? 1 : 0;
}

// ws layout (floats): tickets[tiles] (int32, zeroed once by the caller, re-armed by every launch) | part[tiles*kSplits*2048] |
// dz2[n*128] | dl[n*C] | loss_rows[n].  The tickets and the partial tiles come FIRST, at offsets that depend only on
// (K2, C): a workspace sized for a capacity n_max serves every call with n <= n_max, and the tickets a launch with
// n = n_max re-armed are the ones a later launch with a smaller n reads.
extern "C" int64_t gs_head_ws_floats(int32_t n, int32_t k2_in, int32_t num_classes) {
    int tw, tn;
    const int tiles = wgrad_tiles(k2_in, num_classes, &tw, &tn);
    const int64_t nn = n > 0 ? n : 1;
    return (((int64_t)tiles + 3) & ~(int64_t)3) + (int64_t)tiles * kSplits * (kTM * kTN) +
           ((nn * kD2 + nn * num_classes + nn + 3) & ~(int64_t)3);
}

namespace {

struct HeadPlan {
    int K2, tiles, tiles_w2, tn_w2;
    float *dz2, *dl, *loss_rows, *part;
    int32_t* tickets;
};

int head_plan(int32_t d1, int32_t d2, int32_t num_classes, int32_t n, bool sage, float* ws, HeadPlan* hp) {
    hp->K2 = sage ? 2 * d1 : d1;
    if (!gs_head_supported(d1, hp->K2, d2, num_classes)) return GS_ENOSUP;
    hp->tiles = wgrad_tiles(hp->K2, num_classes, &hp->tiles_w2, &hp->tn_w2);
    hp->tickets = reinterpret_cast<int32_t*>(ws);                         // n-independent offsets first (see above)
    hp->part = ws + (((int64_t)hp->tiles + 3) & ~(int64_t)3);
    hp->dz2 = hp->part + (int64_t)hp->tiles * kSplits * (kTM * kTN);
    hp->dl = hp->dz2 + (int64_t)n * kD2;
    hp->loss_rows = hp->dl + (int64_t)n * num_classes;
    return GS_OK;
}

}  // namespace

extern "C" int gs_head_rows(const float* h1, int64_t ld_h1, int32_t d1,
                            const int32_t* idx, const int32_t* cnt, int32_t width, const int32_t* self_slots,
                            const float* w2, int64_t ld_w2, int32_t d2, int32_t act2,
                            const float* wc, int64_t ld_wc, int32_t num_classes,
                            const int64_t* labels, int32_t n, float grad_scale,
                            float* comb2, int64_t ld_comb2, float* h2, int64_t ld_h2,
                            float* logits, int64_t ld_logits, float* gh1, int64_t ld_gh1,
                            float* ws, void* stream) {
    if (!h1 || !idx || !cnt || !w2 || !wc || !labels || !comb2 || !h2 || !gh1 || !ws || n <= 0 || width <= 0)
        return GS_EINVAL;
    HeadPlan hp;
    int rc = head_plan(d1, d2, num_classes, n, self_slots != nullptr, ws, &hp);
    if (rc) return rc;
    if (!gs_aligned16(h1) || !gs_aligned16(w2) || !gs_aligned16(wc) || !gs_aligned16(comb2) || !gs_aligned16(gh1) ||
        !gs_aligned16(ws) || !gs_aligned16(h2) || (ld_h1 & 3) || (ld_w2 & 3) || (ld_wc & 3) || (ld_comb2 & 3) ||
        (ld_gh1 & 3) || (ld_h2 & 3))
        return GS_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    HeadArgs a{h1, ld_h1, d1, idx, cnt, width, self_slots, w2, ld_w2, act2, wc, ld_wc, num_classes, labels, n,
               grad_scale / (float)n, comb2, ld_comb2, h2, ld_h2, logits, ld_logits, gh1, ld_gh1, hp.dz2, hp.dl,
               hp.loss_rows};
    const size_t smem = head_smem_bytes(hp.K2, num_classes);
    static size_t attr256 = 0, attr128 = 0;
    const int blocks = (n + kRows - 1) / kRows;
    if (hp.K2 == 256) {
        if (smem > attr256) {
            cudaError_t e = cudaFuncSetAttribute(head_rows_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr256 = smem;
        }
        head_rows_kernel<256><<<blocks, kThreads, smem, s>>>(a);
    } else {
        if (smem > attr128) {
            cudaError_t e = cudaFuncSetAttribute(head_rows_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            attr128 = smem;
        }
        head_rows_kernel<128><<<blocks, kThreads, smem, s>>>(a);
    }
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_head_wgrad(const float* comb2, int64_t ld_comb2, const float* h2, int64_t ld_h2,
                             int32_t d1, int32_t d2, int32_t num_classes, int32_t n, int32_t sage,
                             float* loss, float* gw2, int64_t ld_gw2, float* gwc, int64_t ld_gwc,
                             float* ws, void* stream) {
    if (!comb2 || !h2 || !gw2 || !gwc || !ws || n <= 0) return GS_EINVAL;
    HeadPlan hp;
    int rc = head_plan(d1, d2, num_classes, n, sage != 0, ws, &hp);
    if (rc) return rc;
    if (!gs_aligned16(comb2) || !gs_aligned16(h2) || !gs_aligned16(ws) || (ld_comb2 & 3) || (ld_h2 & 3)) return GS_EALIGN;
    WgradArgs b{hp.dz2, comb2, ld_comb2, hp.K2, hp.dl, num_classes, h2, ld_h2, n, gw2, ld_gw2, gwc, ld_gwc, hp.part,
                hp.tickets, hp.loss_rows, loss, hp.tiles_w2, hp.tn_w2};
    GS_PREFER_SMEM(head_wgrad_kernel);
    head_wgrad_kernel<<<dim3(hp.tiles, kSplits), 256, 0, (cudaStream_t)stream>>>(b);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

extern "C" int gs_head_fwd_bwd(const float* h1, int64_t ld_h1, int32_t d1,
                               const int32_t* idx, const int32_t* cnt, int32_t width, const int32_t* self_slots,
                               const float* w2, int64_t ld_w2, int32_t d2, int32_t act2,
                               const float* wc, int64_t ld_wc, int32_t num_classes,
                               const int64_t* labels, int32_t n, float grad_scale,
                               float* comb2, int64_t ld_comb2, float* h2, int64_t ld_h2,
                               float* logits, int64_t ld_logits, float* loss,
                               float* gh1, int64_t ld_gh1, float* gw2, int64_t ld_gw2, float* gwc, int64_t ld_gwc,
                               float* ws, void* stream) {
    int rc = gs_head_rows(h1, ld_h1, d1, idx, cnt, width, self_slots, w2, ld_w2, d2, act2, wc, ld_wc, num_classes,
                          labels, n, grad_scale, comb2, ld_comb2, h2, ld_h2, logits, ld_logits, gh1, ld_gh1, ws, stream);
    if (rc) return rc;
    return gs_head_wgrad(comb2, ld_comb2, h2, ld_h2, d1, d2, num_classes, n, self_slots != nullptr, loss, gw2, ld_gw2,
                         gwc, ld_gwc, ws, stream);
}
