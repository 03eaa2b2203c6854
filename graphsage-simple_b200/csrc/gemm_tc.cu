// K3 on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with a 3-term hi/lo operand split
// ("3xTF32": a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, fp32 accumulation in TMEM), which keeps
// the encoder within the 1e-5 fp32 parity bar that a single TF32 pass misses.
//
//   forward  (NT):  H[n, 128]      = act( X[n, K] . W[128, K]^T )          X = combined tile
//   backward (TN):  dW[128, K]^T   = X[n, K]^T . dZ[n, 128]                split over n
//
// Replaces `F.relu(self.weight.mm(combined.t()))` (graphsage/encoders.py:58-61) and the
// MmBackward that produces `enc.weight.grad` (graphsage/model.py:249) of the reference.
//
// In both kernels the A operand is derived from X (the combined tile, which needs the hi/lo
// split and therefore has to pass through registers anyway) and is staged in TENSOR MEMORY:
// the splitter warps read their row (NT) / column (TN) of the raw TMA tile from shared memory,
// split it, and tcgen05.st the hi and lo halves into a TMEM slot; the MMA then takes A from
// TMEM and only B -- the pre-split W or dZ tiles TMA wrote -- from shared memory.  That halves
// the shared-memory read traffic of the MMAs, which (not the tensor pipe) bounded the first
// version of this kernel (profiles/README.md, r01 tc_gemm v1 vs v2).
//
// SAGE concat in place (graphsage/encoders.py:53-56 `cat([self.features(nodes), neigh_feats])`): when X is
// [table[self_ids] | mean], only the mean half exists in memory; the self rows are gathered straight from the feature
// table with coalesced 16-B cp.async (LDGSTS) into the same swizzled stage ring -- by the splitter group that just
// freed the stage's X region ("refill duty", below) -- so the self half of the combined tile is never written or
// re-read through HBM (gs_sage_encoder_fwd_tc / gs_sage_encoder_wgrad_tc; the engine's default in SAGE mode).
//
// Per CTA (320 threads, 1 CTA/SM): 3-stage smem ring of 48 KB (X raw, Y_hi, Y_lo), TMEM =
// 2 main accumulators + 1 correction accumulator (3 x 128 columns) + 2 A slots (2 x 64 columns).
//   warp 0      TMA producer (W / dZ tiles, X tiles that exist in memory)
//   warps 2..9  two splitter groups (smem -> registers -> TMEM A slot; gathered X tiles: cp.async refill), later the epilogue
//   warp 1      MMA issuer (one lane): per stage 4 k-steps x 3 tcgen05.mma, tcgen05.commit
#include <cuda.h>
#include <cstdlib>
#include "gs_common.cuh"

namespace {

// 3 stages (145 KB) instead of 4 (193 KB): alone the GEMMs lose ~9 % (fwd 73 -> 80 us, wgrad 69 -> 75 us), but the
// SM can then run with a 164 KB shared / 92 KB L1 split, which the co-resident gather needs for its
// outstanding loads: whole step 0.265 -> 0.255 ms (profiles/README.md s4).  -DGS_TC_STAGES=4 restores 4.
#ifndef GS_TC_STAGES
#define GS_TC_STAGES 3
#endif
constexpr int kStages = GS_TC_STAGES;
static_assert(kStages == 3, "the splitter groups' wait rules and refill duty are verified (GPU tests, protocol model) for a "
                            "3-stage ring only");
constexpr int kTile = 128;            // M and N of the UMMA tile (d_out == 128)
constexpr int kChunk = 32;            // reduction elements per stage: 32 fp32 = one 128-B swizzle row
constexpr int kOperandBytes = kTile * kChunk * 4;          // 16 KB
constexpr int kXBytes = kOperandBytes;
constexpr int kStageBytes = kXBytes + 2 * kOperandBytes;   // X raw, Y_hi, Y_lo (pre-split in global memory)
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers + tmem slot*/ + 512 /*NT: table rows of the tile*/;
constexpr int kThreads = 320;          // TMA, MMA, 2 x 4 X-splitter/epilogue warps
// The tensor core rounds toward zero when it adds a k-step into the fp32 accumulator: measured
// bias ~1 ulp per tcgen05.mma (profiles/README.md), i.e. ~1e-5 relative after the 450 MMAs of a
// K = 1204 row.  Spreading the hi.hi products over kMainAccs accumulators and keeping the two
// 2^-11-sized correction products in their own accumulator cuts the number of roundings that
// touch a large partial sum by 9x; the epilogue adds the accumulators in fp32 (round-to-nearest).
constexpr int kMainAccs = 2;
constexpr int kAccs = kMainAccs + 1;  // 3 x 128 accumulator columns
constexpr int kASlots = 2;            // TMEM staging slots for the A operand: 32 hi + 32 lo columns each
constexpr int kASlotCols = 2 * kChunk;
constexpr int kTmemCols = 512;        // 384 accumulator + 128 A-staging columns = the SM's whole TMEM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
// Gathered feature rows: 16-B cp.async (LDGSTS, L2-only) straight into the stage, completion counted on the stage's
// full barrier; src_bytes = 0 zero-fills the 16 bytes.  Measured (profiles/README.md R2.3): one issuing warp pays
// ~45-50 cycles per LDGSTS (32 per chunk = the whole chunk period) and ~55 cycles per 1-D bulk copy (UBLKCP; 128 per
// chunk made the forward 2.9x slower), so the copies are issued by the 128 threads of a splitter group, 8 each.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {      // arrives once this thread's copies have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 format (version 1).
//   K-major : rows of 128 B, 8-row groups SBO = 1024 B apart (LBO unused for one swizzle atom)
//   MN-major: tf32 operands must use SWIZZLE_128B_BASE32B (32-B swizzle atoms, the layout TMA's
//             SWIZZLE_128B_ATOM_32B writes): 128-B rows of 32 M/N elements, 4-deep k groups SBO = 512 B
//             apart, 32-element blocks along M/N LBO apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 2u /* SWIZZLE_128B */) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B (16-B atoms), 1 = SWIZZLE_128B_BASE32B (32-B atoms)
    return d;
}

// Instruction descriptor: D fp32, A/B tf32, M = N = 128, dense.
__host__ __device__ constexpr uint32_t make_idesc(bool b_mn_major) {
    // A comes from tensor memory (row per lane, K along columns -> K-major); B is K-major (W) or N-major (dZ)
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    // hi = x rounded to nearest tf32 (10 explicit mantissa bits); lo = x - hi is exact in fp32
    const uint32_t b = __float_as_uint(x);
    hi = __uint_as_float((b + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

struct TcArgs {
    int n_max;                 // rows of X / dZ the grid was sized for
    const int32_t* n_dev;      // optional device row count
    int k_in;                  // columns of X
    int act;
    float* out;                // NT: h [n, 128]; TN: partials [splits][128][ld_out]
    int64_t ld_out;
    int rows_per_split;        // TN only (multiple of kChunk)
    int64_t split_stride;      // TN only
    // SAGE concat consumed in place: X = [table[self_ids] | mean].  The self half is never materialised: the rows
    // are gathered straight from the feature table (graphsage/encoders.py:53 `self.features(nodes)` + the cat of :56)
    // by the splitter groups' refill duty.  NT: the first self_units K chunks; TN: the first self_units column tiles.
    const float* table;        // nullptr: X is one dense matrix (map_x), self_units == 0
    int64_t ld_table;
    const int32_t* self_ids;   // [n_max] table row of every X row
    int self_units;
    int self_cols;             // feature width F (columns of the self half; the mean half follows at column F of W / dW)
    int debug;                 // experiment switches (GSAGE_TC_DEBUG): read by the DBG instantiation only
    long long* trace;          // optional [64 chunks][16 events] clock64 trace of block 0 (GSAGE_TC_TRACE)
};
// experiment hooks (stage switches, clock64 trace) exist only in the DBG instantiation; production launches the other one
#define TC_TRACE(ev, c) do { if constexpr (DBG) { if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && (c) < 64 && lane == 0) g.trace[(c) * 16 + (ev)] = clock64(); } } while (0)

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t acc) {
    // A from tensor memory, B from shared memory
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// barrier slots (8 bytes each) after the stage ring
constexpr int kBarFull = 0;                       // [kStages]  TMA bytes landed
constexpr int kBarEmpty = kStages;                // [kStages]  stage's MMAs retired (tcgen05.commit)
constexpr int kBarAReady = 2 * kStages;           // [kASlots]  splitter filled the TMEM A slot
constexpr int kBarAFree = 2 * kStages + kASlots;  // [kASlots]  MMAs that read the slot retired
constexpr int kBarAccum = 2 * kStages + 2 * kASlots;
constexpr int kBarYReady = kBarAccum + 1;         // [kStages]  Y tile split into hi/lo in shared memory

// 72 registers: 320 x 72 = 23 K leaves room for the gather blocks that share the SM (gather.cu)
//   map_x          dense part of X: the whole X, or its mean half [n, k_in] when the self half is gathered
//   map_yhi/ylo    NT: W_hi / W_lo columns of the dense part;  TN: dZ_hi / dZ_lo
//   map_shi/slo    NT only: W_hi / W_lo columns of the self half (columns [0, self_cols) of W)
template <bool TN, bool DBG>
__global__ void __maxnreg__(72)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_yhi,
               const __grid_constant__ CUtensorMap map_ylo, const __grid_constant__ CUtensorMap map_shi,
               const __grid_constant__ CUtensorMap map_slo, TcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kStages * kStageBytes;
    const uint32_t tmem_slot = bars + 224;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n = gs_row_count(g.n_max, g.n_dev);
    // NT: chunks [0, self_chunks) take X from the table (gathered rows), the rest from map_x.
    // TN: column tiles [0, self_units) take X from the table, the rest from map_x; chunks run over rows.
    int chunk_begin, chunk_end, x_fixed, self_chunks = 0;
    bool self_tile = false;             // TN: this CTA's column tile lies in the self half
    int out_col0 = 0, col_limit = g.k_in;
    if (!TN) {
        x_fixed = blockIdx.x * kTile;                          // first row of this tile
        if (x_fixed >= n) return;
        self_chunks = g.self_units;
        chunk_begin = 0;
        chunk_end = self_chunks + (g.k_in + kChunk - 1) / kChunk;   // over columns of [self | dense]
    } else {
        self_tile = (int)blockIdx.x < g.self_units;
        x_fixed = (self_tile ? blockIdx.x : blockIdx.x - g.self_units) * kTile;   // first column of this tile (in its half)
        if (self_tile) col_limit = g.self_cols; else out_col0 = g.table ? g.self_cols : 0;
        const int r0 = blockIdx.y * g.rows_per_split;
        const int r1 = min(n, r0 + g.rows_per_split);
        chunk_begin = r0 / kChunk;
        chunk_end = r1 > r0 ? (r1 + kChunk - 1) / kChunk : chunk_begin;   // over rows of X
    }
    const int nchunks = chunk_end - chunk_begin;
    const bool gathers = TN ? self_tile : self_chunks > 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            // producer's expect_tx arrive; in a CTA that gathers rows, every chunk additionally gets one arrival from each
            // of the 128 threads of the splitter group on refill duty (its cp.async group, or a plain arrive for a TMA chunk)
            mbar_init(bars + 8 * (kBarFull + s), gathers ? 1 + 128 : 1);
            mbar_init(bars + 8 * (kBarEmpty + s), 1);          // tcgen05.commit
        }
        for (int a = 0; a < kASlots; ++a) {
            mbar_init(bars + 8 * (kBarAReady + a), 4);         // one arrive per warp of the slot's splitter group
            mbar_init(bars + 8 * (kBarAFree + a), 1);          // tcgen05.commit
        }
        mbar_init(bars + 8 * kBarAccum, 1);
        for (int st = 0; st < kStages; ++st) mbar_init(bars + 8 * (kBarYReady + st), 4);   // the chunk's 4 X-splitter warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int32_t* tile_ids = reinterpret_cast<int32_t*>(gen_base + kStages * kStageBytes + 256);
    if (!TN && gathers && threadIdx.x < kTile) {               // table row of every row of this tile (rows past n: any valid row)
        const int r = x_fixed + (int)threadIdx.x;
        tile_ids[threadIdx.x] = __ldg(g.self_ids + (r < n ? r : n - 1));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + kStages * kStageBytes + 224);
    const uint32_t tmem_a = tmem + kAccs * kTile;              // first A-staging column

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        // W / dZ tiles always; the X tile unless this chunk's rows are gathered from the feature table (then the
        // splitter groups load it, see "refill duty" below).
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages, it = c / kStages;
                mbar_wait(bars + 8 * (kBarEmpty + s), (it & 1) ^ 1);
                TC_TRACE(0, c);
                const uint32_t st = base + s * kStageBytes;
                const uint32_t full = bars + 8 * (kBarFull + s);
                if (DBG && (g.debug & 8) && !gathers) { mbar_arrive(full); continue; }
                if (!TN) {
                    if (c < self_chunks) {
                        const int col = c * kChunk;
                        mbar_expect_tx(full, 2 * kOperandBytes);
                        tma_load_2d(st + kXBytes, &map_shi, col, 0, full);                 // W_hi, self columns
                        tma_load_2d(st + kXBytes + kOperandBytes, &map_slo, col, 0, full); // W_lo
                    } else {
                        const int kc = (c - self_chunks) * kChunk;
                        mbar_expect_tx(full, 3 * kOperandBytes);
                        tma_load_2d(st, &map_x, kc, x_fixed, full);                        // X rows (SW128)
                        tma_load_2d(st + kXBytes, &map_yhi, kc, 0, full);                  // W_hi (SW128, K-major)
                        tma_load_2d(st + kXBytes + kOperandBytes, &map_ylo, kc, 0, full);  // W_lo
                    }
                } else {
                    const int kc = (chunk_begin + c) * kChunk;
                    mbar_expect_tx(full, (self_tile ? 2 : 3) * kOperandBytes);
                    if (!self_tile) tma_load_2d(st, &map_x, x_fixed, kc, full);            // X [32 rows][128 cols], linear
#pragma unroll
                    for (int b = 0; b < 4; ++b) {                                          // dZ in 32-column blocks
                        tma_load_2d(st + kXBytes + b * 4096, &map_yhi, 32 * b, kc, full);
                        tma_load_2d(st + kXBytes + kOperandBytes + b * 4096, &map_ylo, 32 * b, kc, full);
                    }
                }
                TC_TRACE(1, c);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        // The whole warp runs the loop (waits are warp-uniform); one elected lane issues.  Descriptors
        // are base + small offset so that each tcgen05.mma costs a couple of uniform-datapath adds.
        constexpr uint32_t idesc = make_idesc(TN);
        const uint64_t yh_base = TN ? make_desc(base + kXBytes, 4096, 512, 1u) : make_desc(base + kXBytes, 16, 1024);
        const uint32_t d_corr = tmem + (uint32_t)kMainAccs * kTile;
        const bool leader = elect_one();
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % kStages, it = c / kStages;
            const int a = c % kASlots;
            mbar_wait(bars + 8 * (kBarFull + s), it & 1);      // Y_hi / Y_lo tiles landed in shared memory
            mbar_wait(bars + 8 * (kBarYReady + s), it & 1);    // A slot in tensor memory written
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_TRACE(7, c);
            if (leader) {
                const uint32_t a_hi = tmem_a + a * kASlotCols, a_lo = a_hi + kChunk;
                const uint64_t yh0 = yh_base + (uint64_t)((s * kStageBytes) >> 4);
                const uint64_t yl0 = yh0 + (uint64_t)(kOperandBytes >> 4);
#pragma unroll
                for (int j = 0; j < kChunk / 8; ++j) {         // UMMA_K = 8 for tf32
                    // K-major W: advance 32 B inside the swizzle row; N-major dZ: one 8-deep k group (1024 B)
                    const uint64_t dyh = yh0 + (uint64_t)((TN ? 1024 * j : 32 * j) >> 4);
                    const uint64_t dyl = yl0 + (uint64_t)((TN ? 1024 * j : 32 * j) >> 4);
                    const int ks = c * (kChunk / 8) + j;       // k-step index within this CTA
                    const uint32_t d_main = tmem + (uint32_t)(ks % kMainAccs) * kTile;
                    const uint32_t acc_main = ks >= kMainAccs ? 1u : 0u, acc_corr = ks ? 1u : 0u;
                    if (!(DBG && (g.debug & 1))) {
                    umma_tf32_ts(d_corr, a_lo + 8 * j, dyh, idesc, acc_corr);
                    umma_tf32_ts(d_corr, a_hi + 8 * j, dyl, idesc, 1u);
                    umma_tf32_ts(d_main, a_hi + 8 * j, dyh, idesc, acc_main);
                    }
                }
                umma_commit(bars + 8 * (kBarEmpty + s));       // smem stage AND TMEM A slot reusable once these retire
                if (c == nchunks - 1) umma_commit(bars + 8 * kBarAccum);
                TC_TRACE(8, c);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- X splitters (2 groups) + epilogue
        // group 0 (warps 2..5) takes the even chunks and TMEM A slot 0, group 1 (warps 6..9) the odd
        // chunks and slot 1: each group has two chunk periods for its load -> split -> tcgen05.st chain
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;                         // A/accumulator row handled by this thread
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t slot = tmem_a + lane_base + grp * kASlotCols;
        // ---- refill duty (CTAs that gather rows of the feature table).  The X region of a stage is free the moment the
        // group that split chunk c has read it (the MMAs only read the W region), so that group loads the X tile of chunk
        // c + kStages into it right away: 128 threads x 8 coalesced 16-B cp.async instead of one producer warp x 32
        // (measured: ~45 cycles per LDGSTS from one warp made the gathered chunk the whole chunk period).  Every chunk
        // gets exactly one duty -- cp.async + arrive-on-completion for a gathered chunk, a plain arrive for a TMA chunk --
        // by group (c + kStages) % 2; chunks < kStages are served before the loop.
        const int tg = (warp - 2 - 4 * grp) * 32 + lane;       // thread index within the group, 0..127
        const int64_t ld_t = g.ld_table;
        auto refill = [&](int c) {
            const uint32_t st = base + (c % kStages) * kStageBytes;
            const uint32_t full = bars + 8 * (kBarFull + c % kStages);
            if (!TN) {
                if (c >= self_chunks) { mbar_arrive(full); return; }
                // 128 table rows x 128 B in the SWIZZLE_128B layout of the TMA tiles: piece p of row r at r * 128 + ((p ^ (r % 8)) * 16)
                const int col = c * kChunk, piece = tg & 7, sw = (tg >> 3) & 7;
                const uint32_t src_bytes = (col + 4 * piece < (int)ld_t) ? 16u : 0u;       // past the row: zero-fill
                const float* src0 = g.table + (src_bytes ? col + 4 * piece : 0);
                const uint32_t dst0 = st + (uint32_t)((tg >> 3) * 128 + ((piece ^ sw) << 4));
#pragma unroll
                for (int i = 0; i < 8; ++i)                                                // tile row 16 i + tg / 8
                    cp_async16(dst0 + (uint32_t)(16 * i * 128), src0 + (int64_t)tile_ids[16 * i + (tg >> 3)] * ld_t, src_bytes);
            } else {
                // rows kc .. kc + 31 of X = table rows self_ids[kc ..], columns [x_fixed, x_fixed + 128): linear [32][128] tile
                const int kc = (chunk_begin + c) * kChunk, piece = tg & 31;
                const uint32_t src_bytes = (x_fixed + 4 * piece < (int)ld_t) ? 16u : 0u;
                const float* src0 = g.table + (src_bytes ? x_fixed + 4 * piece : 0);
                int ids[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = kc + 4 * i + (tg >> 5);
                    ids[i] = __ldg(g.self_ids + (r < n ? r : n - 1));      // rows past n meet dZ rows of 0
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    cp_async16(st + (uint32_t)((4 * i + (tg >> 5)) * (kTile * 4) + piece * 16), src0 + (int64_t)ids[i] * ld_t, src_bytes);
            }
            cp_async_arrive(full);
        };
        if (gathers)
            for (int c = 0; c < kStages && c < nchunks; ++c)
                if ((c + kStages) % kASlots == grp) refill(c);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % kStages, it = c / kStages;
            // A parity wait only tells phase k from phase k+1, and with an odd stage count the two groups alternate on a
            // stage, so a group that asks for its own chunks' phases only could see "phase k-1 done" as "phase k+1 done"
            // and split a tile that has not landed (measured: 3 stages, 50 % of the 26 000-row weight-gradient runs
            // returned garbage).  Two rules keep the phase unambiguous (tools/tc_protocol_sim.py, tc_protocol_exhaustive.py):
            //  * CTAs whose X tiles all come by TMA: EVERY chunk's full barrier is observed, also the other group's.
            //    Between two waits such a warp only runs register / shared-memory / TMEM work, so it cannot fall a
            //    whole ring cycle behind.
            //  * CTAs on refill duty spend memory-latency-bound time between two waits (id loads + cp.async behind the
            //    co-resident gather's requests), and the phase after the one a late warp is about to ask for completes
            //    WITHOUT that warp (other group's refill + producer): lapped, its parity wait turns into a wait for a
            //    phase that needs its own work -- a deadlock (an 8-GPU bench run hung in round 2).  They wait for their
            //    own chunks only, and first for the stage's EMPTY barrier of chunk c - kStages: the MMAs of that chunk
            //    waited for its full-barrier phase, and the phase after chunk c needs this warp's split of chunk c, so
            //    the full barrier is in the phase of chunk c, pending or complete, and the parity wait is exact.  (The
            //    producer issues chunk c only after the same empty barrier: the extra wait never delays anything.)
            if (!gathers) mbar_wait(bars + 8 * (kBarFull + s), it & 1);
            if (c % kASlots != grp) continue;
            if (gathers) {
                if (c >= kStages) mbar_wait(bars + 8 * (kBarEmpty + s), (it - 1) & 1);
                mbar_wait(bars + 8 * (kBarFull + s), it & 1);
            }
            if (q == 2) TC_TRACE(4, c);
            const uint8_t* xs = gen_base + s * kStageBytes;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t hi[16], lo[16];
                if (DBG && (g.debug & 2)) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) { hi[k] = 0; lo[k] = 0; }
                } else if (!TN) {
                    // row `row` of the SWIZZLE_128B tile (TMA-written or gathered): 16-B chunk j sits at position j ^ (row % 8)
                    const float4* xr = reinterpret_cast<const float4*>(xs + row * 128);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = xr[(4 * half + j) ^ (row & 7)];
                        float h, l;
                        split_tf32(v.x, h, l); hi[4 * j] = __float_as_uint(h); lo[4 * j] = __float_as_uint(l);
                        split_tf32(v.y, h, l); hi[4 * j + 1] = __float_as_uint(h); lo[4 * j + 1] = __float_as_uint(l);
                        split_tf32(v.z, h, l); hi[4 * j + 2] = __float_as_uint(h); lo[4 * j + 2] = __float_as_uint(l);
                        split_tf32(v.w, h, l); hi[4 * j + 3] = __float_as_uint(h); lo[4 * j + 3] = __float_as_uint(l);
                    }
                } else {
                    // column `row` of the linear [32 rows][128 cols] tile (lanes read consecutive words)
                    const float* xc = reinterpret_cast<const float*>(xs) + row + (16 * half) * kTile;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        float h, l;
                        split_tf32(xc[k * kTile], h, l);
                        hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(l);
                    }
                }
                if (half == 0 && c >= kASlots) {
                    // the slot was last read by the MMAs of chunk c-2, whose retirement is what frees that
                    // chunk's smem stage -- wait on the same barrier instead of a second tcgen05.commit
                    const int cp = c - kASlots;
                    mbar_wait(bars + 8 * (kBarEmpty + cp % kStages), (cp / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (q == 2) TC_TRACE(5, c);
                }
                tmem_st16(slot + 16 * half, hi);
                tmem_st16(slot + kChunk + 16 * half, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (q == 2) TC_TRACE(6, c);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (kBarYReady + s));   // same per-stage barrier as the Y splitter
            if (gathers) {
                // all four warps of the group are done with this stage's X region: refill it for chunk c + kStages
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                if (c + kStages < nchunks) refill(c + kStages);
            }
        }
        if (nchunks > 0) {
            mbar_wait(bars + 8 * kBarAccum, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // epilogue: group 0 drains accumulator columns [0, 64), group 1 columns [64, 128)
#pragma unroll 1
        for (int cb = grp * 4; cb < grp * 4 + 4; ++cb) {       // 16-column blocks
            uint32_t r[16];
            if (nchunks > 0) {                                 // sum the accumulators in fp32 (RN)
                tmem_ld16(tmem + lane_base + cb * 16, r);
#pragma unroll 1
                for (int acc = 1; acc < kAccs; ++acc) {
                    uint32_t t2[16];
                    tmem_ld16(tmem + lane_base + acc * kTile + cb * 16, t2);
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t2[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0u;
            }
            if (!TN) {
                const int grow = x_fixed + row;
                if (grow < n) {
                    float* dst = g.out + (int64_t)grow * g.ld_out + cb * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 v = make_float4(gs_apply_act(__uint_as_float(r[i]), g.act),
                                               gs_apply_act(__uint_as_float(r[i + 1]), g.act),
                                               gs_apply_act(__uint_as_float(r[i + 2]), g.act),
                                               gs_apply_act(__uint_as_float(r[i + 3]), g.act));
                        *reinterpret_cast<float4*>(dst + i) = v;
                    }
                }
            } else {
                // thread = column (x_fixed + row) of its half of X, registers = 16 d_out rows of dW: transposed
                // store, consecutive lanes write consecutive columns -> one 128-B line per d_out row per warp
                const int col = x_fixed + row;
                if (col < col_limit) {
                    float* dst = g.out + (int64_t)blockIdx.y * g.split_stride + (int64_t)(cb * 16) * g.ld_out + out_col0 + col;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[(int64_t)i * g.ld_out] = __uint_as_float(r[i]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
    }
}

// W -> (W_hi, W_lo): the pre-split B operand of the forward GEMM (d_out x k_in, tiny)
// columns [0, first) land at [0, first), the rest at [second_off, ...): the two 16-B aligned halves of a SAGE weight
__global__ void split_rows_kernel(const float* __restrict__ src, int64_t ld_src, int rows, int cols, int first, int second_off,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ld_dst) {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols), c = (int)(e - (int64_t)r * cols);
        const int d = c < first ? c : second_off + (c - first);
        float h, l;
        split_tf32(src[(int64_t)r * ld_src + c], h, l);
        hi[(int64_t)r * ld_dst + d] = h;
        lo[(int64_t)r * ld_dst + d] = l;
    }
}

// dz = gh * act'(h) -> (dz_hi, dz_lo) for the rows below *n_dev, zeros above (d % 4 == 0, 128-bit accesses)
__global__ void act_grad_rows_kernel(const float* __restrict__ h, int64_t ld_h, const float* __restrict__ gh, int64_t ld_gh,
                                     int d, int act, int n_max, const int32_t* __restrict__ n_dev,
                                     float* __restrict__ hi, float* __restrict__ lo) {
    const int n = gs_row_count(n_max, n_dev);
    const int d4 = d >> 2;
    const int64_t total = (int64_t)n_max * d4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d4), j = (int)(e - (int64_t)i * d4) * 4;
        float4 oh = make_float4(0.f, 0.f, 0.f, 0.f), ol = oh;
        if (i < n) {
            const float4 a = *reinterpret_cast<const float4*>(h + (int64_t)i * ld_h + j);
            const float4 b = *reinterpret_cast<const float4*>(gh + (int64_t)i * ld_gh + j);
            split_tf32(b.x * gs_act_grad(a.x, act), oh.x, ol.x);
            split_tf32(b.y * gs_act_grad(a.y, act), oh.y, ol.y);
            split_tf32(b.z * gs_act_grad(a.z, act), oh.z, ol.z);
            split_tf32(b.w * gs_act_grad(a.w, act), oh.w, ol.w);
        }
        *reinterpret_cast<float4*>(hi + (int64_t)i * d + j) = oh;
        *reinterpret_cast<float4*>(lo + (int64_t)i * d + j) = ol;
    }
}

__global__ void tc_reduce_kernel(const float* __restrict__ ws, int splits, int64_t stride, int64_t ld_ws,
                                 int M, int N, float* __restrict__ out, int64_t ld_out) {
    const int64_t total = (int64_t)M * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / N), c = (int)(e - (int64_t)m * N);
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * stride + (int64_t)m * ld_ws + c];
        out[(int64_t)m * ld_out + c] = s;
    }
}

// ---- host: tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// row-major fp32 [rows, cols] with leading dimension ld; box = box_cols x box_rows, SWIZZLE_128B
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return GS_ENOSUP;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GS_OK : GS_EINVAL;
}

int tn_splits(int n_max, int tiles) {
    // Enough splits that (a) no CTA accumulates more than kMaxChunksPerSplit stages (bounds the
    // round-toward-zero accumulation bias, see kMainAccs) and (b) tiles x splits fills whole
    // waves of 148 CTAs (1 CTA/SM): the split count is rounded up to the end of the last wave.
    constexpr int kMaxChunksPerSplit = 32;
    int need = (n_max + kMaxChunksPerSplit * kChunk - 1) / (kMaxChunksPerSplit * kChunk);
    if (need < 1) need = 1;
    const int waves = (tiles * need + GS_NUM_SMS - 1) / GS_NUM_SMS;
    int s = waves * GS_NUM_SMS / tiles;
    if (s < need) s = need;
    const int cap = (n_max + 4 * kChunk - 1) / (4 * kChunk);       // at least 4 stages of work per CTA
    if (s > cap) s = cap;
    return s < 1 ? 1 : s;
}

// experiment switches: read once per process, zero / null in production
int tc_debug() {
    static int v = -1;
    if (v < 0) v = getenv("GSAGE_TC_DEBUG") ? atoi(getenv("GSAGE_TC_DEBUG")) : 0;
    return v;
}
long long* tc_trace() {
    static int asked = -1;       // whether the variable exists is decided once; the experiment tool may move the buffer
    if (asked < 0) asked = getenv("GSAGE_TC_TRACE") ? 1 : 0;
    if (!asked) return nullptr;
    const char* e = getenv("GSAGE_TC_TRACE");
    return e ? (long long*)strtoull(e, nullptr, 10) : nullptr;
}

int grid1d(int64_t total) {
    int64_t b = (total + 255) / 256;
    if (b > GS_NUM_SMS * 8) b = GS_NUM_SMS * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int gs_encoder_tc_supported(int32_t k_in, int32_t d_out) {
    // any k_in >= one chunk: the tensor maps carry the true column count, so the tail of the last K chunk (and a k_in
    // that is not a multiple of 4, e.g. Citeseer's 3703 words) is zero-filled by TMA on both operands
    return (d_out == kTile && k_in >= kChunk) ? 1 : 0;
}

namespace {

inline int64_t round4(int64_t v) { return (v + 3) & ~(int64_t)3; }

// Where X comes from: one dense matrix, or [table[self_ids] | dense] with the self half gathered in the kernel.
struct XSource {
    const float* x; int64_t ld_x; int32_t k_dense;             // dense part (the whole X, or the mean half)
    const float* table; int64_t ld_table; const int32_t* self_ids; int32_t self_cols;   // table == nullptr: no self half
};

int check_xsource(const XSource& xs) {
    if (!xs.x || xs.k_dense <= 0) return GS_EINVAL;
    if (!gs_aligned16(xs.x) || (xs.ld_x & 3) || xs.ld_x < xs.k_dense) return GS_EALIGN;
    if (xs.table) {
        if (!xs.self_ids || xs.self_cols <= 0 || xs.ld_table < xs.self_cols) return GS_EINVAL;
        if (!gs_aligned16(xs.table) || (xs.ld_table & 3)) return GS_EALIGN;
    }
    return GS_OK;
}

// W [d_out, self_cols + k_dense] -> W_hi / W_lo in ws, each [d_out, ldw] with the self columns at [0, self_cols) and the
// dense columns at [round4(self_cols), ...): both halves start 16-B aligned, which their TMA maps need
int launch_fwd(const XSource& xs, const float* w, int64_t ld_w, int32_t d_out, int32_t act, int32_t n_max,
               const int32_t* n_dev, float* h, int64_t ld_h, float* ws, cudaStream_t s) {
    const int32_t sc = xs.table ? xs.self_cols : 0;
    const int64_t off = round4(sc), ldw = off + round4(xs.k_dense);
    float* w_hi = ws;
    float* w_lo = ws + (int64_t)d_out * ldw;
    GS_PREFER_SMEM(split_rows_kernel);
    split_rows_kernel<<<grid1d((int64_t)d_out * (sc + xs.k_dense)), 256, 0, s>>>(w, ld_w, d_out, sc + xs.k_dense, sc, (int)off,
                                                                                 w_hi, w_lo, ldw);
    GS_LAUNCH_CHECK();
    CUtensorMap mx, mh, ml, msh, msl;
    int rc;
    if ((rc = make_map(&mx, xs.x, n_max, xs.k_dense, xs.ld_x, kChunk, kTile))) return rc;
    if ((rc = make_map(&mh, w_hi + off, d_out, xs.k_dense, ldw, kChunk, kTile))) return rc;
    if ((rc = make_map(&ml, w_lo + off, d_out, xs.k_dense, ldw, kChunk, kTile))) return rc;
    msh = mh; msl = ml;
    if (sc) {
        if ((rc = make_map(&msh, w_hi, d_out, sc, ldw, kChunk, kTile))) return rc;
        if ((rc = make_map(&msl, w_lo, d_out, sc, ldw, kChunk, kTile))) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    TcArgs g{n_max, n_dev, xs.k_dense, act, h, ld_h, 0, 0, xs.table, xs.ld_table, xs.self_ids,
             sc ? (sc + kChunk - 1) / kChunk : 0, sc, tc_debug(), tc_trace()};
    if (g.debug || g.trace) tc_gemm_kernel<false, true><<<(n_max + kTile - 1) / kTile, kThreads, kSmemBytes, s>>>(mx, mh, ml, msh, msl, g);
    else tc_gemm_kernel<false, false><<<(n_max + kTile - 1) / kTile, kThreads, kSmemBytes, s>>>(mx, mh, ml, msh, msl, g);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

int launch_wgrad(const XSource& xs, const float* h, int64_t ld_h, const float* gh, int64_t ld_gh, int32_t d_out,
                 int32_t act, int32_t n_max, const int32_t* n_dev, float* gw, int64_t ld_gw, float* ws, cudaStream_t s) {
    const int32_t sc = xs.table ? xs.self_cols : 0;
    const int32_t k_all = sc + xs.k_dense;
    const int64_t ldw = round4(k_all);
    float* dz_hi = ws;
    float* dz_lo = ws + (int64_t)n_max * d_out;
    float* part = ws + 2 * (int64_t)n_max * d_out;
    GS_PREFER_SMEM(act_grad_rows_kernel);
    GS_PREFER_SMEM(tc_reduce_kernel);
    act_grad_rows_kernel<<<grid1d((int64_t)n_max * (d_out / 4)), 256, 0, s>>>(h, ld_h, gh, ld_gh, d_out, act, n_max, n_dev,
                                                                             dz_hi, dz_lo);
    GS_LAUNCH_CHECK();
    const int self_tiles = sc ? (sc + kTile - 1) / kTile : 0;
    const int tiles = self_tiles + (xs.k_dense + kTile - 1) / kTile;
    const int splits = tn_splits(n_max, tiles);
    int rps = (n_max + splits - 1) / splits;
    rps = ((rps + kChunk - 1) / kChunk) * kChunk;
    CUtensorMap mx, mh, ml;
    int rc;
    const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;    // MN-major tf32 operand layout
    if ((rc = make_map(&mx, xs.x, n_max, xs.k_dense, xs.ld_x, kTile, kChunk, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    if ((rc = make_map(&mh, dz_hi, n_max, d_out, d_out, 32, kChunk, swz))) return rc;
    if ((rc = make_map(&ml, dz_lo, n_max, d_out, d_out, 32, kChunk, swz))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    TcArgs g{n_max, n_dev, xs.k_dense, GS_ACT_NONE, part, ldw, rps, (int64_t)d_out * ldw, xs.table, xs.ld_table, xs.self_ids,
             self_tiles, sc, tc_debug(), nullptr};
    dim3 grid(tiles, splits);
    if (g.debug) tc_gemm_kernel<true, true><<<grid, kThreads, kSmemBytes, s>>>(mx, mh, ml, mh, ml, g);
    else tc_gemm_kernel<true, false><<<grid, kThreads, kSmemBytes, s>>>(mx, mh, ml, mh, ml, g);
    GS_LAUNCH_CHECK();
    tc_reduce_kernel<<<grid1d((int64_t)d_out * k_all), 256, 0, s>>>(part, splits, (int64_t)d_out * ldw, ldw, d_out, k_all,
                                                                   gw, ld_gw);
    GS_LAUNCH_CHECK();
    return GS_OK;
}

}  // namespace

// floats of workspace for the forward: W_hi + W_lo (k_in = all columns of W; room for the 16-B aligned start of a second half)
extern "C" int64_t gs_encoder_fwd_tc_ws_floats(int32_t k_in, int32_t d_out) {
    return 2 * (int64_t)d_out * (round4(k_in) + 4);
}

extern "C" int gs_encoder_fwd_tc(const float* x, int64_t ld_x, const float* w, int64_t ld_w,
                                 int32_t k_in, int32_t d_out, int32_t act,
                                 int32_t n_max, const int32_t* n_dev,
                                 float* h, int64_t ld_h, float* ws, void* stream) {
    if (!x || !w || !h || !ws || n_max < 0) return GS_EINVAL;
    if (!gs_encoder_tc_supported(k_in, d_out)) return GS_ENOSUP;
    if (!gs_aligned16(x) || !gs_aligned16(w) || !gs_aligned16(h) || !gs_aligned16(ws) || (ld_x & 3) || (ld_w & 3) || (ld_h & 3))
        return GS_EALIGN;
    if (n_max == 0) return GS_OK;
    return launch_fwd(XSource{x, ld_x, k_in, nullptr, 0, nullptr, 0}, w, ld_w, d_out, act, n_max, n_dev, h, ld_h, ws,
                      (cudaStream_t)stream);
}

extern "C" int gs_sage_encoder_fwd_tc(const float* table, int64_t ld_table, const int32_t* self_ids, int32_t feat_dim,
                                      const float* mean, int64_t ld_mean, const float* w, int64_t ld_w,
                                      int32_t d_out, int32_t act, int32_t n_max, const int32_t* n_dev,
                                      float* h, int64_t ld_h, float* ws, void* stream) {
    if (!table || !self_ids || !mean || !w || !h || !ws || n_max < 0 || feat_dim <= 0) return GS_EINVAL;
    if (!gs_encoder_tc_supported(feat_dim, d_out)) return GS_ENOSUP;
    XSource xs{mean, ld_mean, feat_dim, table, ld_table, self_ids, feat_dim};
    int rc = check_xsource(xs);
    if (rc) return rc;
    if (!gs_aligned16(w) || !gs_aligned16(h) || !gs_aligned16(ws) || (ld_w & 3) || (ld_h & 3)) return GS_EALIGN;
    if (ld_w < 2 * (int64_t)feat_dim) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    return launch_fwd(xs, w, ld_w, d_out, act, n_max, n_dev, h, ld_h, ws, (cudaStream_t)stream);
}

// floats of workspace for the weight gradient: dz_hi + dz_lo + split-K partials
extern "C" int64_t gs_encoder_wgrad_tc_ws_floats(int32_t n_max, int32_t k_in, int32_t d_out) {
    const int64_t ldw = round4(k_in);
    // one column tile more than ceil(k_in / 128): a SAGE call tiles its two halves separately
    const int tiles = (k_in + kTile - 1) / kTile + 1;
    int splits = tn_splits(n_max, tiles);
    const int s2 = tn_splits(n_max, tiles - 1);
    if (s2 > splits) splits = s2;
    return 2 * (int64_t)(n_max > 0 ? n_max : 1) * d_out + (int64_t)splits * d_out * ldw;
}

extern "C" int gs_encoder_wgrad_tc(const float* x, int64_t ld_x, const float* h, int64_t ld_h,
                                   const float* gh, int64_t ld_gh, int32_t k_in, int32_t d_out, int32_t act,
                                   int32_t n_max, const int32_t* n_dev,
                                   float* gw, int64_t ld_gw, float* ws, void* stream) {
    if (!x || !h || !gh || !gw || !ws || n_max < 0) return GS_EINVAL;
    if (!gs_encoder_tc_supported(k_in, d_out)) return GS_ENOSUP;
    if (!gs_aligned16(x) || !gs_aligned16(ws) || !gs_aligned16(h) || !gs_aligned16(gh) || (ld_x & 3) || (ld_h & 3) || (ld_gh & 3))
        return GS_EALIGN;
    if (n_max == 0) return GS_OK;
    return launch_wgrad(XSource{x, ld_x, k_in, nullptr, 0, nullptr, 0}, h, ld_h, gh, ld_gh, d_out, act, n_max, n_dev, gw,
                        ld_gw, ws, (cudaStream_t)stream);
}

extern "C" int gs_sage_encoder_wgrad_tc(const float* table, int64_t ld_table, const int32_t* self_ids, int32_t feat_dim,
                                        const float* mean, int64_t ld_mean, const float* h, int64_t ld_h,
                                        const float* gh, int64_t ld_gh, int32_t d_out, int32_t act,
                                        int32_t n_max, const int32_t* n_dev, float* gw, int64_t ld_gw, float* ws,
                                        void* stream) {
    if (!table || !self_ids || !mean || !h || !gh || !gw || !ws || n_max < 0 || feat_dim <= 0) return GS_EINVAL;
    if (!gs_encoder_tc_supported(feat_dim, d_out)) return GS_ENOSUP;
    XSource xs{mean, ld_mean, feat_dim, table, ld_table, self_ids, feat_dim};
    int rc = check_xsource(xs);
    if (rc) return rc;
    if (!gs_aligned16(ws) || !gs_aligned16(h) || !gs_aligned16(gh) || (ld_h & 3) || (ld_gh & 3)) return GS_EALIGN;
    if (ld_gw < 2 * (int64_t)feat_dim) return GS_EINVAL;
    if (n_max == 0) return GS_OK;
    return launch_wgrad(xs, h, ld_h, gh, ld_gh, d_out, act, n_max, n_dev, gw, ld_gw, ws, (cudaStream_t)stream);
}
