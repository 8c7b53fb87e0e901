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
//   lut_reduce_kernel and the launcher
// (SAMPLE16 panels take the tensor-core kernel of pheno_tc.cu instead.)
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
