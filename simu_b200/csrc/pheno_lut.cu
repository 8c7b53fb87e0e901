// Kernel (b), FAST mode: y[:,k] += sum_j t_jk[g_ij] with 4-SNP lookup tables.
//
// Why tables: at the measured 6.4 TB/s a B200 SM has to retire ~90-130
// genotypes per clock, i.e. about one issue slot per genotype; "shift, mask,
// select, add" per genotype per trait is 3-5x over that budget.  The
// contribution of SNP j to a sample is a 4-valued function of the raw 2-bit
// field v (00 -> x(2-2p), 01 missing -> 0, 10 -> x(1-2p), 11 -> x(-2p)), so
// four SNP rows are combined into one byte index e and one table of 256
// entries: one shared-memory load + one add per FOUR genotypes per trait.
//
// This file holds, in this order:
//   lut_build_kernel        the tables (FP64 -> FP32), once per run
//   pheno_lut_kernel        ROWS panels (row-major, optional row list): the layout described below
//   pheno_lut_g4_kernel     GROUP4 panels (the product layout): the byte in HBM IS the table index, so the
//                           bit transposition and three of the four row-word loads disappear
//   pheno_fused_g4_kernel   EXPERIMENTAL counts + lookups in one persistent kernel
//   lut_reduce_kernel, the launchers, and find_freq_pheno built from the standard kernels
//
// Layout of one step (64 causal rows x 4096 samples) in a CTA (ROWS panels):
//   * tile: 64 rows x 1024 B in shared memory, filled by one TMA bulk copy
//     (cp.async.bulk) per row issued by 8 producer warps (8 rows each);
//   * table: 16 groups (4 rows each) x 256 entries, 32 KB, precomputed once
//     per run by lut_build_kernel in FP64 -> FP32, fetched with one bulk copy;
//   * 8 consumer warps, 512 samples each; a lane owns one 32-bit column word
//     (16 samples) and keeps its 16*K partial sums in registers.
// Bank conflicts are designed out: lane l reads word l of every row (bank l),
// and at sub-step t it works on group (l+t) mod 16, whose table column sits in
// bank (l+t) mod 16 (+16 for the upper half-warp, which uses a replica of the
// table; for K=2 the 8-byte entries make each half-warp its own wavefront).
// The 4 row words are bit-transposed to 16 byte indices with 16 LOP3/SHF.
//
// Work is split over min(#SM, steps) persistent CTAs as contiguous ranges of the
// (sample-tile major, row-block minor) step list, so the load is balanced to
// within one step; a CTA flushes its FP32 partial sums to a private FP64
// slot every 64 blocks (1024 terms per accumulator, relative rounding error
// ~1e-6) and at tile boundaries; lut_reduce_kernel adds the slots in a fixed
// order, so the result is deterministic.
//
// Algorithmic bytes: Mc*(ceil(N/4) + 12 + 8K) + 8NK.
#include <stdlib.h>

#include <float.h>

#include <algorithm>

#include "counts_g4.cuh"
#include "ctx.cuh"
#include "pipeline.cuh"

namespace {

constexpr int kBlockRows = 64;                       // causal rows per step
constexpr int kGroups = kBlockRows / 4;              // 16 four-row groups
constexpr int kWarps = 8;                            // consumer warps
constexpr int kTileSamples = kWarps * 512;           // 4096
constexpr int kTileBytes = kTileSamples / 4;         // 1024 per row
constexpr int kTableBytes = 256 * 128;               // 32 KB
constexpr int kStageBytes = kBlockRows * kTileBytes + kTableBytes;   // 96 KB
constexpr int kStages = 2;
constexpr int kFlushBlocks = 64;
#ifndef LUT_PRODUCERS
#define LUT_PRODUCERS 8
#endif
#ifndef LUT_UNROLL
#define LUT_UNROLL 2
#endif
constexpr int kProducers = LUT_PRODUCERS;            // TMA-issuing warps
constexpr int kUnroll = LUT_UNROLL;                  // sub-steps unrolled in the consumer loop
constexpr int kRowsPerProducer = kBlockRows / kProducers;
constexpr int kThreads = (kWarps + kProducers) * 32;
constexpr int kSmemBytes = kStages * kStageBytes + 64 + kProducers * 32 * 8;

// (a & m) | (b & ~m) in one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// ---------------------------------------------------------------- LUT build

// lut[block][e][128 B]:  K=1: 32 floats, slot = replica*16 + group
//                        K=2: 16 float2, slot = group -> (trait1, trait2)
template <int K>
__global__ void __launch_bounds__(256)
lut_build_kernel(const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2,
                 int64_t mc, float *__restrict__ lut)
{
    __shared__ double T[kBlockRows][4][K];
    const int64_t b = blockIdx.x;
    if (threadIdx.x < kBlockRows) {
        const int64_t j = b * kBlockRows + threadIdx.x;
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t0 = 0.0, t2 = 0.0, t3 = 0.0;
            if (j < mc) {
                const double avg = 2.0 * freq[j];
                const double x = (k == 0) ? x1[j] : x2[j];
                t0 = (2.0 - avg) * x;      // bits 00: two A1 copies
                t2 = (1.0 - avg) * x;      // bits 10: one copy
                t3 = (0.0 - avg) * x;      // bits 11: none
            }
            T[threadIdx.x][0][k] = t0;
            T[threadIdx.x][1][k] = 0.0;    // bits 01: missing contributes 0 (source.cc:991)
            T[threadIdx.x][2][k] = t2;
            T[threadIdx.x][3][k] = t3;
        }
    }
    __syncthreads();
    const int e = threadIdx.x;
    float out[32];
#pragma unroll
    for (int g = 0; g < kGroups; g++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double s = T[4 * g + 0][e & 3][k] + T[4 * g + 1][(e >> 2) & 3][k];
            s += T[4 * g + 2][(e >> 4) & 3][k] + T[4 * g + 3][(e >> 6) & 3][k];
            if (K == 1) { out[g] = (float)s; out[16 + g] = (float)s; }
            else out[2 * g + k] = (float)s;
        }
    }
    float4 *dst = reinterpret_cast<float4 *>(lut + (b * 256 + e) * 32);
#pragma unroll
    for (int q = 0; q < 8; q++) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

// ---------------------------------------------------------------- main kernel

struct LutParams {
    const uint8_t *panel;
    int64_t row_stride;
    const int32_t *rows;
    int64_t mc;
    int64_t n_samples;
    const float *lut;
    double *ypart;          // [slots][K][npad]
    int64_t npad;           // tiles * kTileSamples
    int64_t tiles;
    int64_t nblocks;
    int64_t steps;          // tiles * nblocks
};

// Accumulators are float2 so that every add is one packed FADD2 (add.f32x2, sm_100):
//   K=2: acc[s]       = (trait1, trait2) of sample s            (16 registers pairs)
//   K=1: acc[2*q + h] = samples (q + 8h, q + 8h + 4), q = 0..3  ( 8 register pairs)
// The table address e*128 + base comes from ONE dot-product instruction
// (IDP.4A: sum_i z.byte_i * sel.byte_i + base, sel = 0x80 << 8m) instead of PRMT + IMAD.
template <int K>
__device__ __forceinline__ void lookup4(float2 (&acc)[8 * K], const uint32_t z, const uint32_t tbl, const int q)
{
    // bytes 0..3 of z are the table indices of samples q, q+4, q+8, q+12
    uint32_t a[4];
#pragma unroll
    for (int m = 0; m < 4; m++) a[m] = __dp4a(z, 0x80u << (8 * m), tbl);
    if (K == 1) {
        float v[4];
#pragma unroll
        for (int m = 0; m < 4; m++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[m]) : "r"(a[m]));
        acc[2 * q + 0] = __fadd2_rn(acc[2 * q + 0], make_float2(v[0], v[1]));
        acc[2 * q + 1] = __fadd2_rn(acc[2 * q + 1], make_float2(v[2], v[3]));
    } else {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a[m]));
            acc[q + 4 * m] = __fadd2_rn(acc[q + 4 * m], v);
        }
    }
}

// value of (sample s, trait k) in the accumulator layout above
template <int K>
__device__ __forceinline__ float &acc_at(float2 (&acc)[8 * K], int s, int k)
{
    if (K == 2) return k == 0 ? acc[s].x : acc[s].y;
    const int q = s & 3, m = s >> 2;          // s = q + 4m
    float2 &r = acc[2 * q + (m >> 1)];
    return (m & 1) ? r.y : r.x;
}

template <int K>
__device__ __forceinline__ void flush_acc(float2 (&acc)[8 * K], double *__restrict__ ypart, int64_t npad,
                                          int64_t slot, int64_t sample0)
{
#pragma unroll
    for (int k = 0; k < K; k++) {
        double *dst = ypart + (slot * K + k) * npad + sample0;
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            double2 v = *reinterpret_cast<double2 *>(dst + s);
            v.x += (double)acc_at<K>(acc, s, k);
            v.y += (double)acc_at<K>(acc, s + 1, k);
            *reinterpret_cast<double2 *>(dst + s) = v;
        }
    }
#pragma unroll
    for (int i = 0; i < 8 * K; i++) acc[i] = make_float2(0.f, 0.f);
}

template <int K>
__global__ void __launch_bounds__(kThreads, 1)
pheno_lut_kernel(const LutParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t grid = gridDim.x, cta = blockIdx.x;
    StepWalk walk(p.steps, p.nblocks, cta, grid);
    const int64_t len = walk.len;
    typedef StepWalk::Cursor Cursor;
    const Cursor first = walk.first;
    auto advance = [&](Cursor &c) { walk.advance(c); };

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    // full[s] at bar_base + 8*s, empty[s] at bar_base + 16 + 8*s
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(bar_base + 8 * s, kProducers);
            mbar_init(bar_base + 16 + 8 * s, kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kWarps) {
        // ============ producer warps: TMA bulk copies, 32 rows of the step each ============
        // Row sources are looked up one step ahead (the rows[] gather load is off the critical
        // path) and published to shared memory; one lane then issues the copies back to back.
        const int pw = warp - kWarps;
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        const uint8_t **ptrs = reinterpret_cast<const uint8_t **>(smem + kStages * kStageBytes + 64) + pw * 32;
        auto row_src = [&](const Cursor &c) -> const uint8_t * {
            const int64_t j = c.blk * kBlockRows + pw * kRowsPerProducer + lane;
            if (lane >= kRowsPerProducer || j >= p.mc) return nullptr;
            const int64_t row = p.rows ? (int64_t)p.rows[j] : j;
            return p.panel + row * p.row_stride + c.tile * kTileBytes;
        };
        Cursor cn = first;                              // cursor of the step being looked up ahead
        const uint8_t *nxt = (len > 0) ? row_src(cn) : nullptr;
        for (int64_t it = 0; it < len; it++) {
            const int64_t tile = cn.tile, blk = cn.blk;
            const uint8_t *cur = nxt;
            if (it + 1 < len) {
                // look up the next step's row and pull its 1 KB slice into L2 now, one whole
                // step before its bulk copy can be issued (the stage is still being consumed)
                advance(cn);
                nxt = row_src(cn);
                if (nxt) {
                    const int nlines = (int)(min((int64_t)kTileBytes, p.row_stride - cn.tile * kTileBytes) >> 7);
                    for (int q = 0; q < nlines; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + q * 128));
                }
            }
            const int stage = (int)(it & 1);
            const uint32_t parity = (uint32_t)((it >> 1) & 1);
            const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
            const uint32_t stage_base = smem_base + stage * kStageBytes;
            const uint32_t bytes = (uint32_t)min((int64_t)kTileBytes, p.row_stride - tile * kTileBytes);
            const int nvalid = (int)max((int64_t)0, min((int64_t)kRowsPerProducer, p.mc - blk * kBlockRows - pw * kRowsPerProducer));
            ptrs[lane] = cur;
            // one warp polls the mbarrier; the other producer warps sleep in a named hardware
            // barrier until it gets through (bar.sync also orders the ptrs[] stores)
            if (pw == 0) mbar_wait_backoff(empty, parity ^ 1u);
            asm volatile("bar.sync 1, %0;" ::"n"(kProducers * 32) : "memory");
            if (lane == 0) {
                mbar_arrive_expect_tx(full, (uint32_t)nvalid * bytes + (pw == 0 ? (uint32_t)kTableBytes : 0u));
                if (pw == 0)
                    bulk_g2s(stage_base + kBlockRows * kTileBytes, p.lut + blk * (kTableBytes / 4), kTableBytes, full,
                             pol_keep);
                const uint32_t dst0 = stage_base + pw * kRowsPerProducer * kTileBytes;
#pragma unroll 8
                for (int r = 0; r < nvalid; r++) bulk_g2s(dst0 + r * kTileBytes, ptrs[r], bytes, full, pol_stream);
            }
            __syncwarp();
        }
        return;
    }

    // ========================= consumer warps: lookups =========================
    float2 acc[8 * K];
#pragma unroll
    for (int i = 0; i < 8 * K; i++) acc[i] = make_float2(0.f, 0.f);

    int64_t cur_tile = -1, slot = 0;
    int since_flush = 0;
    Cursor cc = first;
    for (int64_t it = 0; it < len; it++, advance(cc)) {
        const int stage = (int)(it & 1);
        const uint32_t parity = (uint32_t)((it >> 1) & 1);
        const int64_t tile = cc.tile;
        if (tile != cur_tile || since_flush == kFlushBlocks) {
            if (cur_tile >= 0)
                flush_acc<K>(acc, p.ypart, p.npad, slot, cur_tile * kTileSamples + warp * 512 + lane * 16);
            if (tile != cur_tile) {
                cur_tile = tile;
                slot = cta - cta_of_step(tile * p.nblocks, grid, p.steps);
            }
            since_flush = 0;
        }
        since_flush++;

        const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
        const uint32_t stage_base = smem_base + stage * kStageBytes;
        const uint32_t words = stage_base + warp * 128 + lane * 4;        // this lane's column word, row 0
        const uint32_t table = stage_base + kBlockRows * kTileBytes;
        mbar_wait(full, parity);

#pragma unroll kUnroll
        for (int t = 0; t < kGroups; t++) {
            const uint32_t gi = (uint32_t)(lane + t) & 15u;
            const uint32_t wa = words + gi * (4u * kTileBytes);
            uint32_t A, B, C, D;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(A) : "r"(wa));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(B) : "r"(wa + kTileBytes));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(C) : "r"(wa + 2 * kTileBytes));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(D) : "r"(wa + 3 * kTileBytes));
            // 4 rows x 16 fields  ->  16 bytes (A | B<<2 | C<<4 | D<<6 per sample)
            const uint32_t X0 = bitsel(A, B << 2, 0x33333333u);     // even samples: nibble = A | B<<2
            const uint32_t X1 = bitsel(A >> 2, B, 0x33333333u);     // odd samples
            const uint32_t Y0 = bitsel(C, D << 2, 0x33333333u);
            const uint32_t Y1 = bitsel(C >> 2, D, 0x33333333u);
            const uint32_t Z00 = bitsel(X0, Y0 << 4, 0x0F0F0F0Fu);  // samples 0,4,8,12
            const uint32_t Z01 = bitsel(X0 >> 4, Y0, 0x0F0F0F0Fu);  // samples 2,6,10,14
            const uint32_t Z10 = bitsel(X1, Y1 << 4, 0x0F0F0F0Fu);  // samples 1,5,9,13
            const uint32_t Z11 = bitsel(X1 >> 4, Y1, 0x0F0F0F0Fu);  // samples 3,7,11,15
            const uint32_t tbl = table + (K == 1 ? (gi + (lane & 16u)) * 4u : gi * 8u);
            lookup4<K>(acc, Z00, tbl, 0);
            lookup4<K>(acc, Z10, tbl, 1);
            lookup4<K>(acc, Z01, tbl, 2);
            lookup4<K>(acc, Z11, tbl, 3);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty);
    }
    if (cur_tile >= 0)
        flush_acc<K>(acc, p.ypart, p.npad, slot, cur_tile * kTileSamples + warp * 512 + lane * 16);
}

// ---------------------------------------------------------------- GROUP4 panel layout
//
// In a GROUP4 panel four consecutive SNP rows are stored sample-interleaved: byte i of group g
// holds the 2-bit fields of sample i for rows 4g..4g+3 (row 4g in bits 0-1) -- which IS the
// table index e.  The kernel then needs no bit transposition at all: a lane reads its 16
// samples of a group with ONE 16-byte shared-memory load and goes straight to the lookups
// (44 instead of ~70 instructions per 64 genotypes), and a step is 16 bulk copies of 4 KB
// instead of 64 of 1 KB.  SPLIT = 2 runs two consumer warps per 512-sample column set, each
// taking every other sub-step (16 consumer warps: more latency hiding when the shared-memory
// pipe is not yet the limit, K = 1), flushing into separate FP64 slots.
constexpr int kG4TileBytes = kTileSamples;           // 4096: one byte per sample per group

template <int SPLIT> struct G4Shape {
    static constexpr int kConsumers = kWarps * SPLIT;
    static constexpr int kThreads = (kConsumers + 1) * 32;      // + one producer warp
    static constexpr int kSmemBytes = kStages * kStageBytes + 64;
};

template <int K>
__device__ __forceinline__ void lookup4_g4(float2 (&acc)[8 * K], const uint32_t z, const uint32_t tbl, const int w)
{
    // bytes 0..3 of z are the table indices of samples 4w .. 4w+3
    uint32_t a[4];
#pragma unroll
    for (int m = 0; m < 4; m++) a[m] = __dp4a(z, 0x80u << (8 * m), tbl);
    if (K == 1) {
        float v[4];
#pragma unroll
        for (int m = 0; m < 4; m++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[m]) : "r"(a[m]));
        acc[2 * w + 0] = __fadd2_rn(acc[2 * w + 0], make_float2(v[0], v[1]));
        acc[2 * w + 1] = __fadd2_rn(acc[2 * w + 1], make_float2(v[2], v[3]));
    } else {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a[m]));
            acc[4 * w + m] = __fadd2_rn(acc[4 * w + m], v);
        }
    }
}

template <int K>
__device__ __forceinline__ void flush_acc_g4(float2 (&acc)[8 * K], double *__restrict__ ypart, int64_t npad,
                                             int64_t slot, int64_t sample0)
{
#pragma unroll
    for (int k = 0; k < K; k++) {
        double *dst = ypart + (slot * K + k) * npad + sample0;
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            double2 v = *reinterpret_cast<double2 *>(dst + s);
            if (K == 1) { v.x += (double)acc[s >> 1].x; v.y += (double)acc[s >> 1].y; }
            else { v.x += (double)(k == 0 ? acc[s].x : acc[s].y); v.y += (double)(k == 0 ? acc[s + 1].x : acc[s + 1].y); }
            *reinterpret_cast<double2 *>(dst + s) = v;
        }
    }
#pragma unroll
    for (int i = 0; i < 8 * K; i++) acc[i] = make_float2(0.f, 0.f);
}

template <int K, int SPLIT>
__global__ void __launch_bounds__(G4Shape<SPLIT>::kThreads, 1)
pheno_lut_g4_kernel(const LutParams p)
{
    constexpr int kConsumers = G4Shape<SPLIT>::kConsumers;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t grid = gridDim.x, cta = blockIdx.x;
    StepWalk walk(p.steps, p.nblocks, cta, grid);
    const int64_t len = walk.len;
    typedef StepWalk::Cursor Cursor;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages * kStageBytes;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 16 + 8 * s, kConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t gstride = p.row_stride;            // bytes per 4-row group (>= N, multiple of 128)
    const int64_t ngroups_total = (p.mc + 3) >> 2;

    if (warp == kConsumers) {
        // ============ producer warp: lane g copies group g's 4 KB slice, lane 16 the table ============
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        Cursor c = walk.first;
        for (int64_t it = 0; it < len; it++) {
            const int stage = (int)(it & 1);
            const uint32_t parity = (uint32_t)((it >> 1) & 1);
            const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
            const uint32_t stage_base = smem_base + stage * kStageBytes;
            const int64_t tile = c.tile, blk = c.blk;
            const uint32_t bytes = (uint32_t)min((int64_t)kG4TileBytes, gstride - tile * kG4TileBytes);
            const int ng = (int)min((int64_t)kGroups, ngroups_total - blk * kGroups);
            const uint8_t *src = p.panel + (blk * kGroups + (lane & 15)) * gstride + tile * kG4TileBytes;
            walk.advance(c);
            if (it + 1 < len) {
                // pull the next step's slices into L2 one whole step ahead (2 lanes per group)
                const int ngn = (int)min((int64_t)kGroups, ngroups_total - c.blk * kGroups);
                if ((lane & 15) < ngn) {
                    const uint8_t *nx = p.panel + (c.blk * kGroups + (lane & 15)) * gstride + c.tile * kG4TileBytes;
                    const int nlines = (int)(min((int64_t)kG4TileBytes, gstride - c.tile * kG4TileBytes) >> 7);
                    for (int q = (lane >> 4); q < nlines; q += 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + q * 128));
                }
            }
            if (lane == 0) {
                mbar_wait_backoff(empty, parity ^ 1u);
                mbar_arrive_expect_tx(full, (uint32_t)ng * bytes + (uint32_t)kTableBytes);
            }
            __syncwarp();
            if (lane < ng) bulk_g2s(stage_base + lane * kG4TileBytes, src, bytes, full, pol_stream);
            else if (lane == 16)
                bulk_g2s(stage_base + kGroups * kG4TileBytes, p.lut + blk * (kTableBytes / 4), kTableBytes, full, pol_keep);
        }
        return;
    }

    // ========================= consumer warps: lookups =========================
    const int cw = warp % kWarps, half = warp / kWarps;       // column set, sub-step parity
    float2 acc[8 * K];
#pragma unroll
    for (int i = 0; i < 8 * K; i++) acc[i] = make_float2(0.f, 0.f);

    int64_t cur_tile = -1, slot = 0;
    int since_flush = 0;
    Cursor cc = walk.first;
    for (int64_t it = 0; it < len; it++, walk.advance(cc)) {
        const int stage = (int)(it & 1);
        const uint32_t parity = (uint32_t)((it >> 1) & 1);
        const int64_t tile = cc.tile;
        if (tile != cur_tile || since_flush == kFlushBlocks * SPLIT) {
            if (cur_tile >= 0)
                flush_acc_g4<K>(acc, p.ypart, p.npad, slot, cur_tile * kTileSamples + cw * 512 + lane * 16);
            if (tile != cur_tile) {
                cur_tile = tile;
                slot = (cta - cta_of_step(tile * p.nblocks, grid, p.steps)) * SPLIT + half;
            }
            since_flush = 0;
        }
        since_flush++;

        const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
        const uint32_t stage_base = smem_base + stage * kStageBytes;
        const uint32_t mine = stage_base + cw * 512 + lane * 16;            // this lane's 16 samples, group 0
        const uint32_t table = stage_base + kGroups * kG4TileBytes;
        mbar_wait(full, parity);

#pragma unroll kUnroll
        for (int i = 0; i < kGroups / SPLIT; i++) {
            const int t = i * SPLIT + half;
            const uint32_t gi = (uint32_t)(lane + t) & 15u;
            uint32_t z0, z1, z2, z3;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(z0), "=r"(z1), "=r"(z2), "=r"(z3)
                         : "r"(mine + gi * kG4TileBytes));
            const uint32_t tbl = table + (K == 1 ? (gi + (lane & 16u)) * 4u : gi * 8u);
            lookup4_g4<K>(acc, z0, tbl, 0);
            lookup4_g4<K>(acc, z1, tbl, 1);
            lookup4_g4<K>(acc, z2, tbl, 2);
            lookup4_g4<K>(acc, z3, tbl, 3);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty);
    }
    if (cur_tile >= 0)
        flush_acc_g4<K>(acc, p.ypart, p.npad, slot, cur_tile * kTileSamples + cw * 512 + lane * 16);
}

// ================================================================ fused one-pass kernel (EXPERIMENTAL)
//
// STATUS (round 1): results are parity-green (tests/test_gpu_fused.py) but it is SLOWER than the two
// passes it replaces -- config 2: 2.4 ms against 0.36 + 0.70 ms, config 4: 19.8 against 4.8 + 6.5 ms --
// so nothing uses it by default (bench.py --fused, simu_b200_find_freq_pheno).  What was measured on
// the way (SIMU_B200_FUSED_DEBUG=1 prints the cycle counters) is in DESIGN.md section 8; in short:
// the lookup side runs at full speed on this schedule (0.79 ms with prebuilt tables), the counting
// side is latency-bound per warp (2400 clk per 4 KB unit, 6 counter warps = 3200 clk per step, above
// the 2048-3072 clk the consumers need), and a block's table is ready only ~0.6 steps before its
// consumers need it whatever the look-ahead, because 2 of the ~49 CTAs of a front are always busy
// building (16K clk per table) and deliver their counts late.
//
// find_freq and find_pheno read every causal row once each: counts_g4 is HBM-bound with the SM
// pipes idle, pheno_lut_g4 shared-memory-bound with HBM half idle.  This kernel does both in ONE
// pass over HBM ("a fused count+accumulate design is credited against Mc*B").
//
//   * Schedule: T sample tiles (nW x 512 samples) x k block fronts = T*k <= #SM co-resident CTAs;
//     CTA (t, f) walks the row blocks f*L .. f*L+L-1 of tile t in order, so all T CTAs of a front are on
//     the same row block at the same local time -- no skew, one FP64 slot per front.
//   * 4 counter warps per CTA run D steps ahead of the consumers: at local time i the CTAs of a
//     front count row block f*L+i+D (group rows split over the CTAs of a 4..7-tile slab, so a lane
//     sees >= 64 words per group and the carry-save tree stays deep).  The loads (L1 no-allocate)
//     leave the block in L2, where the producers' bulk copies find it D steps later: HBM is read once.
//   * A counting unit is (group row, half of the slab's sample range).  Its partial counts go to
//     global memory as ONE 64-bit atomic per SNP row that carries the three counts AND an arrival
//     tick -- [arrivals:7 | lo&hi:19 | hi:19 | lo:19] -- so no fence is needed between "data" and
//     "flag", and the returned old value tells the warp whether it completed the row; the check of
//     the returned value is deferred by one unit, so the atomic's round trip overlaps the next
//     unit's loads.  (First version: 12 32-bit atomics + __threadfence + arrival atomic per unit =
//     ~14 us of serial latency per step, 7x slower than two passes.)  N < 2^19 for the packing.
//   * The warp that completes the last group of a row block turns the counts into freq, scales
//     the raw effects by het^S / sqrt(het) (source.cc:1245-1262, :1272-1286 -- host work between
//     find_freq and find_pheno in the reference), builds the block's 32 KB table exactly as
//     lut_build_kernel does and raises ready[b]; a producer waits for ready[b] before it copies
//     table b.  Deadlock-free: counters wait only for their own CTA's consumers (window D), the
//     slowest CTA of a front always finds its block counted.
constexpr int kCounterWarps = 6;                    // + 1 builder warp
constexpr int kLookahead = 10;
constexpr int kCountBits = 19;                      // per-field width of the packed counters

struct FusedParams {
    const uint8_t *panel;
    int64_t gstride, mc, n_samples;
    int nW, T, k;                    // column sets per CTA, tiles, fronts
    int64_t nb, L;                   // row blocks, blocks per front
    float *lut;
    double *ypart;
    int64_t npad;                    // T * nW * 512
    unsigned long long *cnt;         // [nb][16 groups][4 rows]: arrivals:7 | both:19 | hi:19 | lo:19
    uint32_t *groups_done, *ready;   // [nb], [nb]
    int32_t *counts;                 // outputs (may be NULL)
    double *freq;
    const double *raw1, *raw2;       // effect draws (per causal variant)
    double *x1, *x2;                 // scaled effects out (may alias raw)
    int op;                          // -1: none, 0: *= sqrt(pow(het,S)), 1: /= sqrt(het)
    const int32_t *comp;
    const double *spow;              // [2][ncomp]
    int ncomp;
    unsigned long long *bad;
    int fused;                       // 0: tables are prebuilt, counters idle (debugging / A-B)
    unsigned long long *dbg;         // optional cycle counters (SIMU_B200_FUSED_DEBUG)
    int lookahead;                   // D: steps the counters may run ahead of the consumers
};

template <int SPLIT> struct FusedShape {
    static constexpr int kMaxThreads = (kWarps * SPLIT + 2 + kCounterWarps) * 32;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Table build, shared by the counter warps of a CTA.  The warp that completed a row block owns the
// build; the other counter warps of the CTA notice the request between two counting units and grab
// work items with shared-memory atomics (a single warp needs ~50-100 us for a table -- far more
// than the look-ahead window -- all of them together a few us):
//   phase 0  per (row, trait): counts -> freq -> het -> scaled effect -> the row's 4 table terms
//   phase 1  per group: the 16+16 pair sums T0+T1 and T2+T3
//   phase 2  per 32 entries: entry e of group g = pair01[e & 15] + pair23[e >> 4]  (the association
//            (T0+T1)+(T2+T3) of lut_build_kernel, so both build the same floats), 4 KB coalesced
template <int K> struct BuildShared {
    double T[kBlockRows][4][K];          // [row][raw bits][trait]
    double P[kGroups][32][K];            // [group][0..15: rows 0,1 | 16..31: rows 2,3][trait]
    long long blk;
    int lock, active, inside;
    int next[3], done[3];
};

template <int K>
__device__ __forceinline__ void build_work(const FusedParams &p, BuildShared<K> *bs, const int lane)
{
    volatile BuildShared<K> *vs = bs;
    const int64_t b = vs->blk;
    constexpr int kItems0 = kBlockRows * K / 32;
    constexpr unsigned long long kMask = (1ull << kCountBits) - 1ull;
    // ---- phase 0
    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&bs->next[0], 1);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
        if (it >= kItems0) break;
        const int idx = it * 32 + lane, r = idx / K, k = idx % K;
        const int64_t j = b * kBlockRows + r;
        double t0 = 0.0, t2 = 0.0, t3 = 0.0;
        if (j < p.mc) {
            const unsigned long long packed = __ldcg(p.cnt + b * kBlockRows + r);
            const uint32_t lo = (uint32_t)(packed & kMask), hi = (uint32_t)((packed >> kCountBits) & kMask);
            const uint32_t both = (uint32_t)((packed >> (2 * kCountBits)) & kMask);
            const int n2 = (int)both, n3 = (int)(lo - both), n1 = (int)(hi - both);
            const int n0 = (int)p.n_samples - n1 - n2 - n3;
            const int nm = n2 + n1 + n0;                    // source.cc:874
            const double f = (nm == 0) ? 0.0 : static_cast<double>(2 * n0 + 1 * n1) / static_cast<double>(2 * nm);
            if (k == 0) {
                if (p.counts) reinterpret_cast<int4 *>(p.counts)[j] = make_int4(n0, n1, n2, n3);
                if (p.freq) p.freq[j] = f;
            }
            const double het = fabs(__dmul_rn(__dmul_rn(2.0, f), __dsub_rn(1.0, f)));
            const bool zero_het = p.op >= 0 && het < (double)FLT_EPSILON;      // :1251, :1278 (the host throws)
            if (zero_het && k == 0) atomicMin(p.bad, (unsigned long long)j);
            double x = (k == 0) ? p.raw1[j] : p.raw2[j];
            if (zero_het) x = 0.0;
            else if (p.op == 0) x = __dmul_rn(x, sqrt(pow(het, p.spow[k * p.ncomp + (p.comp ? p.comp[j] : 0)])));
            else if (p.op == 1) x = __ddiv_rn(x, sqrt(het));
            double *xo = (k == 0) ? p.x1 : p.x2;
            if (xo) xo[j] = x;
            const double avg = 2.0 * f;
            t0 = (2.0 - avg) * x; t2 = (1.0 - avg) * x; t3 = (0.0 - avg) * x;   // as lut_build_kernel
        }
        bs->T[r][0][k] = t0; bs->T[r][1][k] = 0.0; bs->T[r][2][k] = t2; bs->T[r][3][k] = t3;
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&bs->done[0], 1); }
    }
    while (vs->done[0] < kItems0) __nanosleep(20);
    __syncwarp();
    // ---- phase 1
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(&bs->next[1], 1);
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        if (g >= kGroups) break;
        const int q = lane & 15, r0 = 4 * g + ((lane >> 4) << 1);
#pragma unroll
        for (int k = 0; k < K; k++) bs->P[g][lane][k] = vs->T[r0][q & 3][k] + vs->T[r0 + 1][q >> 2][k];
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&bs->done[1], 1); }
    }
    while (vs->done[1] < kGroups) __nanosleep(20);
    __syncwarp();
    // ---- phase 2
    float *lut = p.lut + b * (kTableBytes / 4);
    for (;;) {
        int ib = 0;
        if (lane == 0) ib = atomicAdd(&bs->next[2], 1);
        ib = __shfl_sync(0xFFFFFFFFu, ib, 0);
        if (ib >= 8) break;
        const int e = ib * 32 + lane;
        float out[32];
#pragma unroll
        for (int g = 0; g < kGroups; g++) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double sum = vs->P[g][e & 15][k] + vs->P[g][16 + (e >> 4)][k];
                if (K == 1) { out[g] = (float)sum; out[16 + g] = (float)sum; }
                else out[2 * g + k] = (float)sum;
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(lut + e * 32);
#pragma unroll
        for (int q = 0; q < 8; q++) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(&bs->done[2], 1); }   // the owner's gpu-scope fence + release is cumulative
    }
}

// a counter warp between two units: lend a hand if a table build is in progress in this CTA
template <int K>
__device__ __forceinline__ void build_help(const FusedParams &p, BuildShared<K> *bs, const int lane)
{
    volatile BuildShared<K> *vs = bs;
    if (!vs->active) return;
    int in = 0;
    if (lane == 0) { atomicAdd(&bs->inside, 1); __threadfence_block(); in = vs->active; }
    in = __shfl_sync(0xFFFFFFFFu, in, 0);
    if (in) build_work<K>(p, bs, lane);
    __syncwarp();
    if (lane == 0) atomicSub(&bs->inside, 1);
}

// the builder warp: run the build of row block b (with whoever helps) and publish the table
template <int K>
__device__ __forceinline__ void build_own(const FusedParams &p, BuildShared<K> *bs, const int64_t b, const int lane)
{
    volatile BuildShared<K> *vs = bs;
    if (lane == 0) {
        vs->blk = b;
#pragma unroll
        for (int i = 0; i < 3; i++) { vs->next[i] = 0; vs->done[i] = 0; }
        __threadfence_block();
        vs->active = 1;
    }
    __syncwarp();
    build_work<K>(p, bs, lane);
    if (lane == 0) {
        while (vs->done[2] < 8) __nanosleep(20);
        __threadfence();
        st_release_gpu(p.ready + b, 1u);
        vs->active = 0;
        __threadfence_block();
        while (vs->inside > 0) __nanosleep(20);       // helpers have left: the state may be reused
    }
    __syncwarp();
}

template <int K, int SPLIT>
__global__ void __launch_bounds__(FusedShape<SPLIT>::kMaxThreads, 1)
pheno_fused_g4_kernel(const FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nW = p.nW, nC = nW * SPLIT;
    const uint32_t W = (uint32_t)nW * 512u;                  // bytes per group slice = samples per tile
    const uint32_t stage_bytes = kGroups * W + kTableBytes;
    const int t = (int)(blockIdx.x % (unsigned)p.T), f = (int)(blockIdx.x / (unsigned)p.T);
    const int64_t b0 = (int64_t)f * p.L;
    const int64_t nsteps = max((int64_t)0, min(p.L, p.nb - b0));
    const int64_t ngroups_total = (p.mc + 3) >> 2;

    // slab of this tile: balanced partition of the T tiles into runs of 4..7 tiles whose CTAs share the
    // counting of the slab's sample range (CTA j of a slab of S tiles counts the groups g = j mod S)
    struct { int t_lo, t_hi, S, j, nslabs; } slab;
    slab.nslabs = p.T >= 4 ? p.T / 4 : 1;
    {
        int q = 0;
        while (q + 1 < slab.nslabs && (int)(((int64_t)(q + 1) * p.T) / slab.nslabs) <= t) q++;
        slab.t_lo = (int)(((int64_t)q * p.T) / slab.nslabs);
        slab.t_hi = (int)(((int64_t)(q + 1) * p.T) / slab.nslabs);
        slab.S = slab.t_hi - slab.t_lo;
        slab.j = t - slab.t_lo;
    }

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages * stage_bytes;
    volatile int *progress = reinterpret_cast<volatile int *>(smem + kStages * stage_bytes + 32);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 16 + 8 * s, nC);
        }
        *progress = 0;
        BuildShared<K> *bs0 = reinterpret_cast<BuildShared<K> *>(smem + kStages * stage_bytes + 64);
        bs0->lock = 0; bs0->active = 0; bs0->inside = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == nC) {
        // ================= producer warp =================
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        const uint32_t bytes = (uint32_t)min((int64_t)W, p.gstride - (int64_t)t * W);
        const int pf_S = slab.S, pf_j = slab.j;
        const int64_t pf_lo = (int64_t)slab.t_lo * W, pf_bytes = min(p.gstride, (int64_t)slab.t_hi * W) - pf_lo;
        for (int64_t it = 0; it < nsteps; it++) {
            const int64_t b = b0 + it;
            const int stage = (int)(it & 1);
            const uint32_t parity = (uint32_t)((it >> 1) & 1);
            const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
            const uint32_t stage_base = smem_base + stage * stage_bytes;
            const int ng = (int)min((int64_t)kGroups, ngroups_total - b * kGroups);
            const uint8_t *src = p.panel + (b * kGroups + (lane & 15)) * p.gstride + (int64_t)t * W;
            if (!p.fused && it + 1 < nsteps) {      // unfused: pull the next step's slices into L2 (fused: the counters did)
                const int ngn = (int)min((int64_t)kGroups, ngroups_total - (b + 1) * kGroups);
                if ((lane & 15) < ngn)
                    for (uint32_t q = (uint32_t)(lane >> 4); q < (bytes >> 7); q += 2)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + kGroups * p.gstride + q * 128));
            }
            if (p.fused && it + p.lookahead + 2 < nsteps) {
                // the counters of this CTA will read their slab of block b + D + 2: pull it from HBM into L2 now
                const int64_t bp = b + p.lookahead + 2;
                const int ngp = (int)min((int64_t)kGroups, ngroups_total - bp * kGroups);
                for (int g = pf_j; g < ngp; g += pf_S) {
                    const uint8_t *row = p.panel + (bp * kGroups + g) * p.gstride + pf_lo;
                    for (int64_t o = (int64_t)lane * 128; o < pf_bytes; o += 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                }
            }
            if (lane == 0) {
                mbar_wait_backoff(empty, parity ^ 1u);
                if (p.fused) {                       // the table of block b exists once all its rows are counted
                    const long long t0 = clock64();
                    while (ld_acquire_gpu(p.ready + b) == 0u) {
                        __nanosleep(200);
                        if (clock64() - t0 > 4000000000ll) __trap();
                    }
                    if (p.dbg) atomicAdd(p.dbg + 0, (unsigned long long)(clock64() - t0));
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (other SMs) -> async-proxy reads
                }
                mbar_arrive_expect_tx(full, (uint32_t)ng * bytes + (uint32_t)kTableBytes);
            }
            __syncwarp();
            if (lane < ng) bulk_g2s(stage_base + lane * W, src, bytes, full, pol_stream);
            else if (lane == 16)
                bulk_g2s(stage_base + kGroups * W, p.lut + b * (kTableBytes / 4), kTableBytes, full, pol_keep);
        }
        return;
    }

    if (warp == nC + 1 + kCounterWarps) {
        // ================= builder warp: tables of the row blocks i = t, t + T, ... of this front =================
        // (static assignment: the CTA that happens to deliver a block's last count must not be the
        // one that builds its table -- it would fall behind with its counting, deliver the last
        // count of the next block too, and serialise all builds of the front)
        if (!p.fused) return;
        BuildShared<K> *bs = reinterpret_cast<BuildShared<K> *>(smem + kStages * stage_bytes + 64);
        for (int64_t i = t; i < nsteps; i += p.T) {
            const int64_t b = b0 + i;
            const uint32_t ngb = (uint32_t)min((int64_t)kGroups, ngroups_total - b * kGroups);
            const long long t0 = clock64();
            while (ld_acquire_gpu(p.groups_done + b) < ngb) {
                __nanosleep(100);
                if (clock64() - t0 > 4000000000ll) __trap();
            }
            const long long tb0 = clock64();
            const int prog0 = *progress;
            build_own<K>(p, bs, b, lane);
            if (p.dbg && lane == 0) {
                atomicAdd(p.dbg + 5, (unsigned long long)(clock64() - tb0));
                atomicAdd(p.dbg + 6, 1ull);
                atomicAdd(p.dbg + 7, (unsigned long long)(i - (int64_t)*progress + 64));
                atomicAdd(p.dbg + 8, (unsigned long long)(tb0 - t0));
                atomicAdd(p.dbg + 9, (unsigned long long)(i - (int64_t)prog0 + 64));
            }
        }
        return;
    }

    if (warp > nC) {
        // ================= counter warps: allele counts of the block D steps ahead =================
        if (!p.fused) return;
        const int w = warp - nC - 1;
        BuildShared<K> *bs = reinterpret_cast<BuildShared<K> *>(smem + kStages * stage_bytes + 64);
        const int S = slab.S, j = slab.j;
        const int nvec_g = (int)((p.n_samples + 15) >> 4), nfull_g = nvec_g - 1;
        const int64_t tail_bits = 8 * (p.n_samples - 16 * (int64_t)nfull_g);
        const int v_begin = (int)(((int64_t)slab.t_lo * W) >> 4);
        const int v_end = (int)min((int64_t)nvec_g, ((int64_t)slab.t_hi * W) >> 4);
        const int v_mid = v_begin + (((v_end - v_begin) / 2 + 31) & ~31);      // halves: whole 32-vector strides first
        const int gmax = (kGroups - j + S - 1) / S;             // groups of a full block that are this CTA's
        const int ups = 2 * gmax;                               // counting units per step
        const uint32_t units_per_row = (uint32_t)(2 * slab.nslabs);
        // stream of units (step, group, half), dealt round-robin to the counter warps; the completion
        // check of a unit is deferred until the next unit's loads are in flight
        unsigned long long pend_old = 0;
        int64_t pend_b = -1;
        auto resolve = [&]() {
            if (pend_b < 0) return;
            const uint32_t arrivals = (uint32_t)(__shfl_sync(0xFFFFFFFFu, pend_old, 0) >> (3 * kCountBits));
            if (arrivals + 1u == units_per_row && lane == 0) {  // this unit completed its group (all four rows alike)
                __threadfence();
                atomicAdd(p.groups_done + pend_b, 1u);          // the block's builder polls this
            }
            pend_b = -1;
        };
        for (int64_t U = w; U < nsteps * ups; U += kCounterWarps) {
            const int64_t ic = U / ups;
            const int u = (int)(U - ic * ups);
            if (ic > (int64_t)*progress + p.lookahead) {
                resolve();                                      // do not sit on a completion while throttled
                const long long t0 = clock64();
                while (ic > (int64_t)*progress + p.lookahead) {
                    build_help<K>(p, bs, lane);
                    __nanosleep(100);
                    if (clock64() - t0 > 4000000000ll) __trap();
                }
                if (p.dbg && lane == 0) atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - t0));
            }
            const int64_t b = b0 + ic;
            const int ngb = (int)min((int64_t)kGroups, ngroups_total - b * kGroups);
            const int g = j + (u >> 1) * S;
            if (g >= ngb) { continue; }
            const int half = u & 1;
            const int vb = half ? v_mid : v_begin, ve = half ? v_end : min(v_mid, v_end);
            uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, both[4] = {0, 0, 0, 0};
            const uint4 *row = reinterpret_cast<const uint4 *>(p.panel + (b * kGroups + g) * p.gstride);
            const long long tc0 = clock64();
            if (ve > vb) g4::count_vectors(row + vb, lane, 32, ve - vb, nfull_g - vb, tail_bits, lo, hi, both);
            const long long tc1 = clock64();
            resolve();                                          // previous unit's atomics have long returned
            const long long tc2 = clock64();
            if (p.dbg && lane == 0) {
                atomicAdd(p.dbg + 2, (unsigned long long)(tc1 - tc0));
                atomicAdd(p.dbg + 3, (unsigned long long)(tc2 - tc1));
                atomicAdd(p.dbg + 4, 1ull);
            }
            unsigned long long mine = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const unsigned long long a = __reduce_add_sync(0xFFFFFFFFu, lo[r]);
                const unsigned long long h = __reduce_add_sync(0xFFFFFFFFu, hi[r]);
                const unsigned long long c = __reduce_add_sync(0xFFFFFFFFu, both[r]);
                if (lane == r) mine = a | (h << kCountBits) | (c << (2 * kCountBits)) | (1ull << (3 * kCountBits));
            }
            if (lane < 4) pend_old = atomicAdd(p.cnt + (b * kGroups + g) * 4 + lane, mine);
            pend_b = b;
        }
        resolve();
        return;
    }

    // ================= consumer warps: lookups (as pheno_lut_g4_kernel) =================
    const int cw = warp % nW, half = warp / nW;
    float2 acc[8 * K];
#pragma unroll
    for (int i = 0; i < 8 * K; i++) acc[i] = make_float2(0.f, 0.f);
    const int64_t slot = (int64_t)f * SPLIT + half;
    const int64_t sample0 = (int64_t)t * W + cw * 512 + lane * 16;
    int since_flush = 0;
    for (int64_t it = 0; it < nsteps; it++) {
        const int stage = (int)(it & 1);
        const uint32_t parity = (uint32_t)((it >> 1) & 1);
        if (since_flush == kFlushBlocks * SPLIT) {
            flush_acc_g4<K>(acc, p.ypart, p.npad, slot, sample0);
            since_flush = 0;
        }
        since_flush++;
        const uint32_t full = bar_base + 8 * stage, empty = bar_base + 16 + 8 * stage;
        const uint32_t stage_base = smem_base + stage * stage_bytes;
        const uint32_t mine = stage_base + cw * 512 + lane * 16;
        const uint32_t table = stage_base + kGroups * W;
        mbar_wait(full, parity);
#pragma unroll kUnroll
        for (int i = 0; i < kGroups / SPLIT; i++) {
            const int tt = i * SPLIT + half;
            const uint32_t gi = (uint32_t)(lane + tt) & 15u;
            uint32_t z0, z1, z2, z3;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(z0), "=r"(z1), "=r"(z2), "=r"(z3)
                         : "r"(mine + gi * W));
            const uint32_t tbl = table + (K == 1 ? (gi + (lane & 16u)) * 4u : gi * 8u);
            lookup4_g4<K>(acc, z0, tbl, 0);
            lookup4_g4<K>(acc, z1, tbl, 1);
            lookup4_g4<K>(acc, z2, tbl, 2);
            lookup4_g4<K>(acc, z3, tbl, 3);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty);
            if (warp == 0) *progress = (int)(it + 1);
        }
    }
    if (nsteps > 0) flush_acc_g4<K>(acc, p.ypart, p.npad, slot, sample0);
}

// y[k][i] = sum over the slots of the tile that holds sample i, in slot order
template <int K>
__global__ void __launch_bounds__(256)
lut_reduce_kernel(const double *__restrict__ ypart, int64_t npad, int64_t n, int slots,
                  double *__restrict__ y1, double *__restrict__ y2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = 0.0;
        for (int q = 0; q < slots; q++) s += ypart[((int64_t)q * K + k) * npad + i];
        if (k == 0) y1[i] = s; else y2[i] = s;
    }
}

template <int K>
int run_lut(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq, const double *d_x1,
            const double *d_x2, double *d_y1, double *d_y2)
{
    const int64_t n = ctx->n_samples;
    const int64_t tiles = (n + kTileSamples - 1) / kTileSamples;
    const int64_t nblocks = (mc + kBlockRows - 1) / kBlockRows;
    const int64_t steps = tiles * nblocks;
    const int64_t grid = steps < ctx->num_sms ? steps : ctx->num_sms;
    const int slots1 = (int)((grid + tiles - 1) / tiles) + 1;
    const int64_t npad = tiles * kTileSamples;

    float *lut = (float *)ctx_scratch(ctx, SCR_LUT, (size_t)nblocks * kTableBytes);

    // GROUP4 panels: K = 1 runs two consumer warps per column set (measured on config 4: 8.30 ms
    // row-major -> 6.96 ms GROUP4 -> 6.66 ms with the split); K = 2 sits on the shared-memory pipe
    // either way and keeps 8 consumer warps.  SIMU_B200_LUT_SPLIT=1|2 overrides for A/B runs.
    const bool g4 = ctx->layout == SIMU_B200_LAYOUT_GROUP4;
    int split = (g4 && K == 1) ? 2 : 1;
    if (const char *e = getenv("SIMU_B200_LUT_SPLIT")) { if (g4 && K == 1) split = atoi(e) == 1 ? 1 : 2; }
    const int slots = slots1 * split;
    const size_t ypart_bytes = (size_t)slots * K * npad * sizeof(double);
    double *ypart = (double *)ctx_scratch(ctx, SCR_YPART, ypart_bytes);
    if (!lut || !ypart) return SIMU_B200_ENOMEM;
    if (!ctx->lut_attr_set[K]) {   // per context: the attribute is per device
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_g4_kernel<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G4Shape<1>::kSmemBytes));
        if (K == 1)
            CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_g4_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G4Shape<2>::kSmemBytes));
        ctx->lut_attr_set[K] = true;
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ypart, 0, ypart_bytes, ctx->stream));
    int ps = prof_begin(ctx, PROF_LUT_BUILD);
    lut_build_kernel<K><<<(unsigned)nblocks, 256, 0, ctx->stream>>>(d_freq, d_x1, d_x2, mc, lut);
    prof_end(ctx, ps);

    LutParams p;
    p.panel = ctx->panel; p.row_stride = g4 ? ctx->group_stride : ctx->row_stride; p.rows = d_rows; p.mc = mc;
    p.n_samples = n; p.lut = lut; p.ypart = ypart; p.npad = npad; p.tiles = tiles; p.nblocks = nblocks; p.steps = steps;
    ps = prof_begin(ctx, PROF_LUT_MAIN);
    if (g4) {
        if (K == 1 && split == 2)
            pheno_lut_g4_kernel<1, 2><<<(unsigned)grid, G4Shape<2>::kThreads, G4Shape<2>::kSmemBytes, ctx->stream>>>(p);
        else
            pheno_lut_g4_kernel<K, 1><<<(unsigned)grid, G4Shape<1>::kThreads, G4Shape<1>::kSmemBytes, ctx->stream>>>(p);
    } else {
        pheno_lut_kernel<K><<<(unsigned)grid, kThreads, kSmemBytes, ctx->stream>>>(p);
    }
    prof_end(ctx, ps);
    ps = prof_begin(ctx, PROF_LUT_REDUCE);
    lut_reduce_kernel<K><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ypart, npad, n, slots, d_y1, d_y2);
    prof_end(ctx, ps);
    ctx->launches += 3;
    return ctx_cuda(ctx, cudaGetLastError(), "pheno_lut launch");
}


// wavefronts of the shared-memory data pipe one CTA spends per step (the model DESIGN.md validates
// against ncu), used to pick the tile width and the number of fronts
static double fused_step_cost(int nW, int K)
{
    const double W = nW * 512.0;
    return nW * 16.0 * (4 + 16 * K) + (16 * W + kTableBytes) / 128.0 + 16 * W / 128.0;
}

template <int K>
int run_fused(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1, const double *d_raw2, int op, const int32_t *d_comp,
              const double *d_spow, int ncomp, int32_t *d_counts, double *d_freq, double *d_x1, double *d_x2,
              double *d_y1, double *d_y2, unsigned long long *d_bad, int fused, const double *d_freq_in)
{
    constexpr int SPLIT = 1;      // 8 consumer + 1 producer + 6 counter + 1 builder warps = 512 threads, 128 registers each
    const int64_t n = ctx->n_samples;
    const int64_t nb = (mc + kBlockRows - 1) / kBlockRows;
    // schedule: tile width nW*512 and k fronts with T*k <= #SM, minimising (blocks per front) x (cost per step)
    int best_nW = 0, best_k = 0, best_T = 0;
    double best = 0;
    for (int nW = 2; nW <= kWarps; nW++) {
        const int64_t T = (n + nW * 512 - 1) / (nW * 512);
        if (T > ctx->num_sms) continue;
        int64_t k = ctx->num_sms / T;
        if (k > nb) k = nb;
        const int64_t L = (nb + k - 1) / k;
        const double cost = (double)L * fused_step_cost(nW, K);
        if (!best_nW || cost < best) { best = cost; best_nW = nW; best_k = (int)k; best_T = (int)T; }
    }
    if (!best_nW) return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: %lld samples need more than %d tiles of 4096",
                                  (long long)n, ctx->num_sms);
    if (const char *e = getenv("SIMU_B200_FUSED_NW")) {          // A/B override
        const int nW = atoi(e);
        const int64_t T = (n + nW * 512 - 1) / (nW * 512);
        if (nW >= 1 && nW <= kWarps && T <= ctx->num_sms) { best_nW = nW; best_T = (int)T; best_k = (int)std::min<int64_t>(ctx->num_sms / T, nb); }
    }
    const int nW = best_nW, T = best_T, k = best_k;
    const int64_t L = (nb + k - 1) / k;
    const int64_t npad = (int64_t)T * nW * 512;
    const int slots = k * SPLIT;

    float *lut = (float *)ctx_scratch(ctx, SCR_LUT, (size_t)nb * kTableBytes);
    const size_t ypart_bytes = (size_t)slots * K * npad * sizeof(double);
    double *ypart = (double *)ctx_scratch(ctx, SCR_YPART, ypart_bytes);
    if (n >= (1ll << kCountBits))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: packed counters hold fewer than 2^%d samples", kCountBits);
    const size_t sync_words = (size_t)nb * (kBlockRows * 2 + 2);             // 64-bit counters, then groups_done, ready
    uint32_t *sync = (uint32_t *)ctx_scratch(ctx, SCR_FUSED, sync_words * 4);
    if (!lut || !ypart || !sync) return SIMU_B200_ENOMEM;
    const size_t smem = (size_t)kStages * (kGroups * nW * 512 + kTableBytes) + 64 + sizeof(BuildShared<K>);
    if (!ctx->fused_attr_set[K]) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_fused_g4_kernel<K, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kStages * kStageBytes + 64 + (int)sizeof(BuildShared<K>)));
        ctx->fused_attr_set[K] = true;
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ypart, 0, ypart_bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(sync, 0, sync_words * 4, ctx->stream));
    if (!fused) {      // A/B: tables from given frequencies, counters idle
        lut_build_kernel<K><<<(unsigned)nb, 256, 0, ctx->stream>>>(d_freq_in, d_raw1, d_raw2, mc, lut);
        ctx->launches++;
    }
    FusedParams p;
    p.panel = ctx->panel; p.gstride = ctx->group_stride; p.mc = mc; p.n_samples = n;
    p.nW = nW; p.T = T; p.k = k; p.nb = nb; p.L = L; p.lut = lut; p.ypart = ypart; p.npad = npad;
    p.cnt = reinterpret_cast<unsigned long long *>(sync);
    p.groups_done = sync + (size_t)nb * kBlockRows * 2; p.ready = p.groups_done + nb;
    p.counts = d_counts; p.freq = d_freq; p.raw1 = d_raw1; p.raw2 = d_raw2; p.x1 = d_x1; p.x2 = d_x2;
    p.op = op; p.comp = d_comp; p.spow = d_spow; p.ncomp = ncomp; p.bad = d_bad; p.fused = fused;
    p.lookahead = kLookahead;
    if (const char *e = getenv("SIMU_B200_FUSED_LOOKAHEAD")) p.lookahead = std::max(1, atoi(e));
    p.dbg = nullptr;
    if (getenv("SIMU_B200_FUSED_DEBUG")) {
        p.dbg = (unsigned long long *)ctx_scratch(ctx, SCR_DEBUG, 256);
        CUDA_TRY(ctx, cudaMemsetAsync(p.dbg, 0, 128, ctx->stream));
    }
    const int threads = (nW * SPLIT + 2 + kCounterWarps) * 32;
    const int ps = prof_begin(ctx, PROF_FUSED);
    pheno_fused_g4_kernel<K, SPLIT><<<(unsigned)(T * k), threads, smem, ctx->stream>>>(p);
    prof_end(ctx, ps);
    const int pr = prof_begin(ctx, PROF_LUT_REDUCE);
    lut_reduce_kernel<K><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ypart, npad, n, slots, d_y1, d_y2);
    prof_end(ctx, pr);
    ctx->launches += 2;
    if (p.dbg) {
        unsigned long long h[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "fused dbg: grid %d x steps %lld nW %d | producer ready-wait %.0f clk/CTA-step | counter throttle %.0f clk/warp-step"
                " | count %.0f clk/unit | resolve %.0f clk/unit | units %llu\n", T * k, (long long)L, nW,
                (double)h[0] / ((double)T * k * L), (double)h[1] / ((double)T * k * L * kCounterWarps), (double)h[2] / (double)(h[4] ? h[4] : 1),
                (double)h[3] / (double)(h[4] ? h[4] : 1), h[4]);
        fprintf(stderr, "fused dbg: builds %llu, %.0f clk each, block ready %.2f steps ahead of the builder's consumers\n", h[6],
                (double)h[5] / (double)(h[6] ? h[6] : 1), (double)h[7] / (double)(h[6] ? h[6] : 1) - 64.0);
        fprintf(stderr, "fused dbg: builder polled %.0f clk per build; block complete %.2f steps ahead\n",
                (double)h[8] / (double)(h[6] ? h[6] : 1), (double)h[9] / (double)(h[6] ? h[6] : 1) - 64.0);
    }
    return ctx_cuda(ctx, cudaGetLastError(), "pheno_fused launch");
}

// ------------------------------------------------- find_freq_pheno from the standard kernels
//
// counts -> effect scaling -> table build -> lookups, all on the device, so that one call replaces
// find_freq + the host's effect scaling + find_pheno without a host round trip in between.
// SIMU_B200_OVERLAP_MB=<chunk size> turns it into a chunked two-stream pipeline (stream B does
// counts -> scaling -> tables of chunk c+1 while stream A looks up chunk c; the lookups of all chunks
// accumulate into the same FP64 slots, reduced once).  The idea: counts_g4 is HBM-bound with idle SM
// pipes, pheno_lut_g4 shared-memory-bound with ~45 % of HBM idle, and a counts CTA (no shared memory,
// 20 K registers) fits on an SM beside the resident lookup CTA.  MEASURED (config 2, 1.124 ms as two
// passes): 3.78 / 2.46 / 1.50 / 1.19 ms with 32 / 64 / 256 / 1300 MB chunks -- ~34 us of launch work per
// chunk on the CPU, and with few chunks the counts kernel squeezed to one CTA per SM (32 KB in flight
// instead of 96) runs at a third of its bandwidth and becomes the bottleneck.  So the default is ONE
// chunk (= the plain sequence); the pipeline stays as an experiment switch.
template <int K>
int run_overlap(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1, const double *d_raw2, int op, const int32_t *d_comp,
                const double *d_spow, int ncomp, int32_t *d_counts, double *d_freq, double *d_x1, double *d_x2,
                double *d_y1, double *d_y2, unsigned long long *d_bad)
{
    const int64_t n = ctx->n_samples;
    const int64_t tiles = (n + kTileSamples - 1) / kTileSamples;
    const int64_t npad = tiles * kTileSamples;
    int64_t chunk_mb = 1 << 24;                        // default: one chunk
    if (const char *e = getenv("SIMU_B200_OVERLAP_MB")) chunk_mb = std::max(1, atoi(e));
    int64_t chunk_rows = (chunk_mb << 20) / std::max<int64_t>(1, ctx->group_stride / 4);
    chunk_rows = std::max<int64_t>(kBlockRows, chunk_rows / kBlockRows * kBlockRows);
    const int64_t nchunks = (mc + chunk_rows - 1) / chunk_rows;
    const int64_t nblocks_all = (mc + kBlockRows - 1) / kBlockRows;
    // pipelined: 8 consumer warps, which leaves registers for a counts CTA on the SM; one chunk: the best kernel
    const int split = (K == 1 && nchunks == 1) ? 2 : 1;
    const int64_t steps_max = tiles * ((std::min(mc, chunk_rows) + kBlockRows - 1) / kBlockRows);
    const int64_t grid_max = std::min<int64_t>(steps_max, ctx->num_sms);
    const int slots = ((int)((grid_max + tiles - 1) / tiles) + 1) * split;

    float *lut = (float *)ctx_scratch(ctx, SCR_LUT, (size_t)nblocks_all * kTableBytes);
    const size_t ypart_bytes = (size_t)slots * K * npad * sizeof(double);
    double *ypart = (double *)ctx_scratch(ctx, SCR_YPART, ypart_bytes);
    if (!d_freq) d_freq = (double *)ctx_scratch(ctx, SCR_FREQ, (size_t)mc * 8);
    if (!d_x1) d_x1 = (double *)ctx_scratch(ctx, SCR_X1, (size_t)mc * 8);
    if (K == 2 && !d_x2) d_x2 = (double *)ctx_scratch(ctx, SCR_X2, (size_t)mc * 8);
    if (!lut || !ypart || !d_freq || !d_x1 || (K == 2 && !d_x2)) return SIMU_B200_ENOMEM;
    if (!ctx->lut_attr_set[K]) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_g4_kernel<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G4Shape<1>::kSmemBytes));
        if (K == 1)
            CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_lut_g4_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G4Shape<2>::kSmemBytes));
        ctx->lut_attr_set[K] = true;
    }
    if (!ctx->aux_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    while ((int64_t)ctx->chunk_events.size() < nchunks + 1) {
        cudaEvent_t e;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_events.push_back(e);
    }
    cudaStream_t sA = ctx->stream, sB = nchunks == 1 ? ctx->stream : ctx->aux_stream;   // one chunk: one stream
    const int ps = prof_begin(ctx, PROF_FUSED);
    CUDA_TRY(ctx, cudaMemsetAsync(ypart, 0, ypart_bytes, sA));
    if (sB != sA) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_events[nchunks], sA));      // fork: B starts after everything queued on A
        CUDA_TRY(ctx, cudaStreamWaitEvent(sB, ctx->chunk_events[nchunks], 0));
    }
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t r0 = c * chunk_rows, mcc = std::min(chunk_rows, mc - r0);
        const int64_t b0 = r0 / kBlockRows, nb = (mcc + kBlockRows - 1) / kBlockRows;
        // ---- stream B: counts -> effect scaling -> tables of chunk c
        int rc = launch_counts_g4_range(ctx, sB, r0 / 4, mcc, d_counts ? d_counts + 4 * r0 : nullptr, d_freq + r0, 256);
        if (rc) return rc;
        if (d_x1 != d_raw1) CUDA_TRY(ctx, cudaMemcpyAsync(d_x1 + r0, d_raw1 + r0, (size_t)mcc * 8, cudaMemcpyDeviceToDevice, sB));
        if (K == 2 && d_x2 != d_raw2)
            CUDA_TRY(ctx, cudaMemcpyAsync(d_x2 + r0, d_raw2 + r0, (size_t)mcc * 8, cudaMemcpyDeviceToDevice, sB));
        if (op >= 0 && (rc = launch_scale_effects_on(ctx, sB, d_freq + r0, d_x1 + r0, K == 2 ? d_x2 + r0 : nullptr, mcc,
                                                     d_comp ? d_comp + r0 : nullptr, d_spow, ncomp, op, d_bad, r0)))
            return rc;
        lut_build_kernel<K><<<(unsigned)nb, 256, 0, sB>>>(d_freq + r0, d_x1 + r0, K == 2 ? d_x2 + r0 : nullptr, mcc,
                                                          lut + b0 * (kTableBytes / 4));
        // ---- stream A: lookups of chunk c into the shared slots
        if (sB != sA) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_events[c], sB));
            CUDA_TRY(ctx, cudaStreamWaitEvent(sA, ctx->chunk_events[c], 0));
        }
        LutParams p;
        p.panel = ctx->panel + (r0 / 4) * ctx->group_stride; p.row_stride = ctx->group_stride; p.rows = nullptr; p.mc = mcc;
        p.n_samples = n; p.lut = lut + b0 * (kTableBytes / 4); p.ypart = ypart; p.npad = npad; p.tiles = tiles;
        p.nblocks = nb; p.steps = tiles * nb;
        const int64_t grid = std::min<int64_t>(p.steps, ctx->num_sms);
        if (K == 1 && split == 2)
            pheno_lut_g4_kernel<1, 2><<<(unsigned)grid, G4Shape<2>::kThreads, G4Shape<2>::kSmemBytes, sA>>>(p);
        else
            pheno_lut_g4_kernel<K, 1><<<(unsigned)grid, G4Shape<1>::kThreads, G4Shape<1>::kSmemBytes, sA>>>(p);
        ctx->launches += 2;
    }
    lut_reduce_kernel<K><<<(unsigned)((n + 255) / 256), 256, 0, sA>>>(ypart, npad, n, slots, d_y1, d_y2);
    prof_end(ctx, ps);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "find_freq_pheno (overlap) launch");
}

}  // namespace

int launch_pheno_lut(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                     const double *d_x1, const double *d_x2, double *d_y1, double *d_y2)
{
    if (mc <= 0) {   // source.cc:962-966: phenotypes start at zero
        CUDA_TRY(ctx, cudaMemsetAsync(d_y1, 0, (size_t)ctx->n_samples * sizeof(double), ctx->stream));
        if (d_y2) CUDA_TRY(ctx, cudaMemsetAsync(d_y2, 0, (size_t)ctx->n_samples * sizeof(double), ctx->stream));
        return SIMU_B200_OK;
    }
    if (d_x2 && d_y2) return run_lut<2>(ctx, d_rows, mc, d_freq, d_x1, d_x2, d_y1, d_y2);
    return run_lut<1>(ctx, d_rows, mc, d_freq, d_x1, nullptr, d_y1, nullptr);
}

int launch_pheno_fused(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1, const double *d_raw2, int op,
                       const int32_t *d_comp, const double *d_spow, int ncomp, int32_t *d_counts, double *d_freq,
                       double *d_x1, double *d_x2, double *d_y1, double *d_y2, unsigned long long *d_bad, int fused,
                       const double *d_freq_in)
{
    if (d_raw2 && d_y2)
        return run_fused<2>(ctx, mc, d_raw1, d_raw2, op, d_comp, d_spow, ncomp, d_counts, d_freq, d_x1, d_x2, d_y1, d_y2, d_bad,
                            fused, d_freq_in);
    return run_fused<1>(ctx, mc, d_raw1, nullptr, op, d_comp, d_spow, ncomp, d_counts, d_freq, d_x1, nullptr, d_y1, nullptr,
                        d_bad, fused, d_freq_in);
}

int launch_pheno_overlap(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1, const double *d_raw2, int op,
                         const int32_t *d_comp, const double *d_spow, int ncomp, int32_t *d_counts, double *d_freq,
                         double *d_x1, double *d_x2, double *d_y1, double *d_y2, unsigned long long *d_bad)
{
    if (d_raw2 && d_y2)
        return run_overlap<2>(ctx, mc, d_raw1, d_raw2, op, d_comp, d_spow, ncomp, d_counts, d_freq, d_x1, d_x2, d_y1, d_y2, d_bad);
    return run_overlap<1>(ctx, mc, d_raw1, nullptr, op, d_comp, d_spow, ncomp, d_counts, d_freq, d_x1, nullptr, d_y1, nullptr, d_bad);
}
