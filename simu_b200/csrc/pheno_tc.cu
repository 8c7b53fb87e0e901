// Kernel (b), FAST mode on SAMPLE16 panels: find_pheno (src/source.cc:951-1006) as an exact INTEGER
// matrix product on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// Why tensor cores for a K <= 2 path.  The 4-SNP lookup-table kernel (pheno_lut.cu) moves 2 bytes through
// the shared-memory data pipe per genotype and trait pair; ncu showed that pipe 94 % busy at 0.55 of the
// HBM roofline for two traits (profiles/r01_g4_config2_ncu.txt): the path is NOT memory-bound on the CUDA
// cores.  Here the genotypes never go back through shared memory after their one 4-byte load per 16 SNPs:
// a thread expands its word in registers, stores the bytes to TENSOR MEMORY, and the MMA reads them there.
//
// The arithmetic.  With the panel's dosage code c in {0, 1, 2, 3 = missing} (stage.cu) and m = [c == 3],
//     contribution_j = (a - 2 f_j) x_j [c != 3],  a = 2 - c      (source.cc:986-996)
//                    = x_j (2 - 2 f_j)  +  (-x_j) c  +  x_j (1 + 2 f_j) m
// so   y_i = C + sum_j Wp_j c_ij + sum_j Wm_j m_ij,   C = sum_j x_j (2 - 2 f_j), Wp = -x, Wm = x (1 + 2 f).
// Wp, Wm are rounded to 40-bit fixed point relative to the largest |W| of the trait (quantum 2^e, a power
// of two: error <= 2^-41 max|W| per weight, ~1e-12 of max|y| in total, against the stated 1e-5) and split
// into six signed base-256 digits.  The MMA is tcgen05.mma.kind::i8, M = 128 samples, N = 16 columns
// (trait x digit), K = 32 bytes = 16 SNPs x {c plane, m plane}:
//     A (u8, from TMEM)   byte = c << 2s  resp.  m << 2s : the word ANDed with a byte mask, no shifts
//     B (s8, from smem)   digit_d( W << (6 - 2s) ), canonical K-major no-swizzle core matrices
//     D (s32, in TMEM)    64 * sum_j c_ij * digit_d(Wp_j) + ... : EXACT integer sums, order-independent
// Every <= 256 steps (16384 SNPs: |D| < 2^29) the owners of the TMEM lanes read D back, recombine the
// digits in 64-bit integers and add the result, scaled by 2^(e-6), to an FP64 register.
//
// One step of a CTA = 64 SNPs (4 groups) x 512 samples:
//   producer warp   TMA bulk copies: 4 x 2 KB of genotype words + the 2 KB weight tile of the 64-SNP block
//                   into a 12-stage shared-memory ring (mbarrier full / empty)
//   16 expander     sub-tile t = warp / 4 owns samples 128 t .. 128 t + 127 = TMEM lanes 0..127 and is an
//   warps           independent pipeline: a thread loads its 4 words (LDS.32, conflict-free, one step ahead),
//                   forms 8 masked words per raw word (9 LOP3/SHF per 16 genotypes), writes them with ONE
//                   tcgen05.st.32x32b.x32 into one of two A buffers; after a 128-thread named barrier the
//                   sub-tile's first warp issues 4 x tcgen05.mma (8 clk each at N = 16) from one elected lane,
//                   and tcgen05.commit releases the A buffer, the stage's weight tile and, at drain points, D
//                   (a single MMA-issuing warp for all sub-tiles was 2.4x slower: five dependent mbarrier
//                   waits of ~90 clk per step on one warp)
// Work is split over min(#SM, steps) persistent CTAs exactly as in pheno_lut.cu (contiguous ranges of the
// (tile, block) step list, rotated start, per-(CTA, tile) FP64 slots added in a fixed order).
//
// The m plane is SPARSE (missing genotypes only).  When the panel's missing-genotype index is available
// (miss_index.cu: a list of the missing entries built once per panel) the kernel runs WITHOUT the m plane --
// K = 32 bytes then cover the c planes of two groups, half the tensor-memory traffic and half the MMAs -- and
// miss_correct_kernel adds sum_j Wm_j m_ij from the list as exact 64-bit integers in the same quantum: the same
// fixed-point sum, equal to FP64 rounding (config 2: 0.547 -> 0.375 ms, the list pass runs beside it).  DENSE = true keeps the plane (index not
// applicable or more than 0.4 % missing).
//
// Algorithmic bytes: Mc (ceil(N/4) + 12 + 8K) + 8NK.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "ctx.cuh"
#include "pipeline.cuh"

namespace {

constexpr int kSub = 6;                             // 128-sample sub-tiles (one TMEM lane set each)
constexpr int kTile = 128 * kSub;                   // 768 samples per CTA tile
constexpr int kStepGroups = 4;                      // 16-SNP groups per tcgen05.st (32 TMEM columns)
constexpr int kHalves = 2;                          // ... and per shared-memory stage / A buffer
constexpr int kStageGroups = kHalves * kStepGroups; // 8 groups = 128 SNPs per stage
constexpr int kStepSnps = 16 * kStepGroups;         // 64: one weight tile
constexpr int kStageSnps = 16 * kStageGroups;       // 128
constexpr int kExpWarps = 4 * kSub;                 // 24
constexpr int kMmaWarp0 = kExpWarps;                // one MMA-issuing warp per sub-tile
constexpr int kProducerWarp = kExpWarps + kSub;
constexpr int kThreads = (kExpWarps + kSub + 1) * 32;      // 31 warps
constexpr int kBoxSamples = kTile / 2;              // a tensor copy: 384 samples (192 8-byte elements <= 256) x 8 groups
constexpr int kBoxRowBytes = kBoxSamples * 4;       // 1536
constexpr int kBoxBytes = kStageGroups * kBoxRowBytes;      // 12 KB
constexpr int kWTileBytes = kStepGroups * 512;      // weights of 64 SNPs: 4 x [N = 16][K = 32] s8
constexpr int kStageBytes = 2 * kBoxBytes + kHalves * kWTileBytes;     // 24 KB + 4 KB
constexpr int kStages = 5;                          // 140 KB: also keeps a second CTA (and its TMEM request) off the SM
constexpr int kFlushSteps = 128;                    // stages between drains: 16384 SNPs, |D| < 2^29
constexpr int kDigits = 6;                          // signed base-256 digits per weight (of 8 columns per trait)
constexpr int kFracBits = 40;                       // |W| / 2^e < 2^40
constexpr int kTmemCols = 512;
constexpr int kColD = 0;                            // D: kSub x 16 columns
constexpr int kColA = 16 * kSub;                    // A: kSub x 64 columns (one buffer per sub-tile = one stage)
static_assert(kColA + 64 * kSub <= kTmemCols, "TMEM columns");
// barriers (8 bytes each) after the ring
constexpr int kBarFull = 0, kBarEmpty = kBarFull + kStages, kBarAEmpty = kBarEmpty + kStages,
              kBarDFull = kBarAEmpty + kSub, kNumBars = kBarDFull + kSub;
constexpr int kSmemBytes = kStages * kStageBytes + kNumBars * 8 + 16;

// ------------------------------------------------------------------ tcgen05 wrappers

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem], issued by one elected lane of a converged warp (operands warp-uniform)
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], bd, %4, {%6, %7, %8, %9}, p;\n\t}"
        ::"r"(d), "r"(a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
// keep a value in a register: without this the compiler re-derives shared-memory addresses (S2UR + ULEA + UIADD3 ...)
// at every use, which was a quarter of all issued instructions
__device__ __forceinline__ uint32_t pin(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Wait with a counted bound (a protocol bug must trap, not hang the GPU): try_wait itself suspends the warp for
// a hardware-defined time, so 2^22 failed polls are far beyond any legitimate wait.  (Letting one lane poll and
// the warp reconverge behind it was measured 50 % SLOWER than all lanes polling.)
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity)
{
#pragma unroll 1
    for (int i = 0; i < (1 << 22); i++)
        if (mbar_test(bar, parity)) return;
    __trap();
}

// shared-memory matrix descriptor of one weight sub-tile: [N = 16][K = 32] s8, K-major, no swizzle, as 8 x 16-byte
// core matrices: K chunks 256 B apart (leading-dimension byte offset), 8-column groups 128 B apart (stride)
__device__ __forceinline__ uint32_t wtile_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((256u >> 4) << 16); }
constexpr uint32_t kWtileDescHi = (128u >> 4) | (1u << 14);
// one box of the panel's tensor map (256 x 8-byte elements = 512 samples, kStageGroups rows) -> shared memory
__device__ __forceinline__ void tensor_g2s(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar), "l"(policy) : "memory");
}
// instruction descriptor: D = s32, A = u8, B = s8, both K-major, N = 16, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

// ------------------------------------------------------------------ weights

struct TcScale {            // per trait, written by the weight kernels
    double constant;        // C = sum_j x_j (2 - 2 f_j)
    int exponent;           // e: weights are integers in units of 2^e
    int pad;
};

__device__ __forceinline__ void weights_of(double f, double x, double &wp, double &wm, double &c)
{
    wp = -x;
    wm = x * (1.0 + 2.0 * f);
    c = x * (2.0 - 2.0 * f);
}

// the weight tiles.  One work item (blk, t), t < 8 * K, converts the 16 SNPs of one (64-SNP block, group q,
// plane, trait) to fixed point ONCE and writes all eight digit columns of that trait (128 bytes).
template <int K>
__device__ __forceinline__ void weight_item(const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2,
                                            int64_t mc, const double (&inv_quantum)[2], uint8_t *__restrict__ wt, int64_t blk, int t,
                                            long long *__restrict__ wm_fix /* sparse m plane: [mc][K] weights in quanta, else NULL */)
{
    const int q = t & 3, plane = (t >> 2) & 1, k = t >> 3;
    if (wm_fix && plane == 1) {           // the m plane is a list: its weights stay whole numbers of quanta
#pragma unroll
        for (int f = 0; f < 16; f++) {
            const int64_t j = blk * kStepSnps + 16 * q + f;
            if (j < mc) {
                double wp, wm, c;
                weights_of(freq[j], k == 0 ? x1[j] : x2[j], wp, wm, c);
                wm_fix[j * K + k] = __double2ll_rn(wm * inv_quantum[k]);
            }
        }
        return;
    }
    uint32_t out[8][4];
#pragma unroll
    for (int d = 0; d < 8; d++)
#pragma unroll
        for (int i = 0; i < 4; i++) out[d][i] = 0;
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
        const int s = kk >> 2, b = kk & 3;                         // byte b of the word masked for field s
        const int64_t j = blk * kStepSnps + 16 * q + 4 * b + s;   // SNP 4b + s of the group
        if (j < mc) {
            double wp, wm, c;
            weights_of(freq[j], k == 0 ? x1[j] : x2[j], wp, wm, c);
            long long v = __double2ll_rn((plane ? wm : wp) * inv_quantum[k]);   // |v| <= 2^40 (the quantum is a power of two)
            v <<= 6 - 2 * s;                                        // the A byte carries the factor 4^s
#pragma unroll
            for (int d = 0; d < kDigits; d++) {
                const long long digit = (long long)(signed char)(v & 0xFF);
                v = (v - digit) >> 8;
                out[d][kk >> 2] |= (uint32_t)(uint8_t)digit << (8 * (kk & 3));
            }
        }
    }
    // sub-tile q at q*512; element (n, kappa): (kappa / 16) * 256 + (n / 8) * 128 + (n % 8) * 16 + kappa % 16, n = 8 k + d
    // sparse m plane: an MMA's 32 K bytes are the c planes of TWO groups -- group G = 4 (blk & 1) + q of the 128-SNP
    // stage goes to sub-tile G / 2, K half G % 2 (2 KB of tiles per stage instead of 4 KB)
    const int G = 4 * (int)(blk & 1) + q;
    uint4 *dst = wm_fix ? reinterpret_cast<uint4 *>(wt + (blk >> 1) * (kHalves * kWTileBytes / 2) + (G >> 1) * 512 + (G & 1) * 256 + k * 128)
                        : reinterpret_cast<uint4 *>(wt + blk * kWTileBytes + q * 512 + plane * 256 + k * 128);
#pragma unroll
    for (int d = 0; d < 8; d++) dst[d] = make_uint4(out[d][0], out[d][1], out[d][2], out[d][3]);
    if (K == 1) {                                                   // the second trait's columns are zero
#pragma unroll
        for (int d = 0; d < 8; d++) dst[8 + d] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// Pass 1 (largest |W| per trait -> the power-of-two quantum; the constant term, summed in a fixed order) and pass 2 (the
// tiles) in ONE cooperative launch (causal sets above kSmallMc): every block sums / maximises its strided share, the blocks
// meet at a grid barrier, EVERY block reduces the per-block partials in the same fixed order (so all of them hold the same
// exponents without another round trip) and goes on to its share of the weight tiles.  It replaced memset + a prepare kernel
// + a weights kernel (29 us of a 0.81 ms step on config 2, 28 us of 0.17 ms on config 3: two launches, a serial
// last-block loop over 128 partials and a ticket).  The barrier cleans up after itself: the last block to leave zeroes both
// counters, so the launch parameters never change and the step still replays as one CUDA graph.
constexpr int kFusedThreads = 256;
template <int K>
__global__ void __launch_bounds__(kFusedThreads)
tc_weights_fused_kernel(const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2, int64_t mc,
                        int64_t ntiles, double *__restrict__ partial /* [K][grid] sums, then [K][grid] maxima */,
                        unsigned int *__restrict__ bar, TcScale *__restrict__ scale, uint8_t *__restrict__ wt,
                        long long *__restrict__ wm_fix)
{
    __shared__ double ssum[K][kFusedThreads], smax[K][kFusedThreads];
    const int G = (int)gridDim.x;                                  // <= kFusedThreads (host)
    double sum[K], mx[K];
#pragma unroll
    for (int k = 0; k < K; k++) { sum[k] = 0.0; mx[k] = 0.0; }
    for (int64_t j = (int64_t)blockIdx.x * kFusedThreads + threadIdx.x; j < mc; j += (int64_t)G * kFusedThreads) {
        const double f = freq[j];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double wp, wm, c;
            weights_of(f, k == 0 ? x1[j] : x2[j], wp, wm, c);
            sum[k] += c;
            mx[k] = fmax(mx[k], fmax(fabs(wp), fabs(wm)));
        }
    }
    auto tree = [&]() {                                            // fixed shape: deterministic for a given grid
        __syncthreads();
        for (int s = kFusedThreads / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    ssum[k][threadIdx.x] += ssum[k][threadIdx.x + s];
                    smax[k][threadIdx.x] = fmax(smax[k][threadIdx.x], smax[k][threadIdx.x + s]);
                }
            }
            __syncthreads();
        }
    };
#pragma unroll
    for (int k = 0; k < K; k++) { ssum[k][threadIdx.x] = sum[k]; smax[k][threadIdx.x] = mx[k]; }
    tree();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            partial[k * G + blockIdx.x] = ssum[k][0];
            partial[(K + k) * G + blockIdx.x] = smax[k][0];
        }
    }
    grid_barrier_once(bar);
#pragma unroll
    for (int k = 0; k < K; k++) {
        const bool in = (int)threadIdx.x < G;
        ssum[k][threadIdx.x] = in ? __ldcg(partial + k * G + threadIdx.x) : 0.0;
        smax[k][threadIdx.x] = in ? __ldcg(partial + (K + k) * G + threadIdx.x) : 0.0;
    }
    tree();
    int e[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double m = smax[k][0];
        // |W| < 2^(ilogb(m) + 1)  ->  |W| / 2^e < 2^kFracBits
        e[k] = (m > 0.0 && isfinite(m)) ? ilogb(m) + 1 - kFracBits : 0;
    }
    if (blockIdx.x == 0 && threadIdx.x < K) {
        const int k = threadIdx.x;
        scale[k].constant = ssum[k][0];
        scale[k].exponent = e[k];
        scale[k].pad = 0;
    }
    const double inv_quantum[2] = {ldexp(1.0, -e[0]), K == 2 ? ldexp(1.0, -e[1]) : 0.0};
    for (int64_t item = (int64_t)blockIdx.x * kFusedThreads + threadIdx.x; item < ntiles * (8 * K); item += (int64_t)G * kFusedThreads)
        weight_item<K>(freq, x1, x2, mc, inv_quantum, wt, item / (8 * K), (int)(item % (8 * K)), wm_fix);
}

// both passes in ONE CTA for small causal sets (the sparse configs are launch-bound: two dependent launches and a
// grid-wide ticket cost more than the arithmetic).  Same summation order as nothing else: the constant is a sum
// over a fixed strided partition followed by a fixed tree, so it is deterministic for a given mc.
constexpr int kSmallThreads = 1024, kSmallMc = 4096;
template <int K>
__global__ void __launch_bounds__(kSmallThreads)
tc_small_prepare_kernel(const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2, int64_t mc,
                        int64_t ntiles, TcScale *__restrict__ scale, uint8_t *__restrict__ wt, long long *__restrict__ wm_fix)
{
    __shared__ double ssum[K][kSmallThreads], smax[K][kSmallThreads];
    __shared__ int sexp[2];
    double sum[K], mx[K];
#pragma unroll
    for (int k = 0; k < K; k++) { sum[k] = 0.0; mx[k] = 0.0; }
    for (int64_t j = threadIdx.x; j < mc; j += kSmallThreads) {
        const double f = freq[j];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double wp, wm, c;
            weights_of(f, k == 0 ? x1[j] : x2[j], wp, wm, c);
            sum[k] += c;
            mx[k] = fmax(mx[k], fmax(fabs(wp), fabs(wm)));
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) { ssum[k][threadIdx.x] = sum[k]; smax[k][threadIdx.x] = mx[k]; }
    __syncthreads();
    for (int s = kSmallThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                ssum[k][threadIdx.x] += ssum[k][threadIdx.x + s];
                smax[k][threadIdx.x] = fmax(smax[k][threadIdx.x], smax[k][threadIdx.x + s]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        const double m = smax[k][0];
        const int e = (m > 0.0 && isfinite(m)) ? ilogb(m) + 1 - kFracBits : 0;
        scale[k].constant = ssum[k][0];
        scale[k].exponent = e;
        scale[k].pad = 0;
        sexp[k] = e;
    }
    __syncthreads();
    const double inv_quantum[2] = {ldexp(1.0, -sexp[0]), K == 2 ? ldexp(1.0, -sexp[1]) : 0.0};
    for (int64_t item = threadIdx.x; item < ntiles * (8 * K); item += kSmallThreads)
        weight_item<K>(freq, x1, x2, mc, inv_quantum, wt, item / (8 * K), (int)(item % (8 * K)), wm_fix);
}

// ------------------------------------------------------------------ main kernel

struct TcParams {
    const uint8_t *panel;
    int64_t s16_stride;
    int64_t mc;
    int64_t n_samples;
    const uint8_t *wt;
    const TcScale *scale;
    double *ypart;          // [slots][K][npad]
    int64_t npad;           // tiles * kTile
    int64_t tiles;
    int64_t nblocks;        // 128-SNP stage blocks
    int64_t steps;          // tiles * nblocks
};

// The CTA's contiguous range of the (sample-tile major, SNP-block minor) step list, walked in the rotated
// order of StepWalk (pipeline.cuh) with 32-bit state: every warp of the CTA steps through it on its own, and a
// single warp executes ~1 dependent instruction per 5 clocks, so the 64-bit divisions of a naive cursor
// (it % stages, it / stages, ...) made the producer warp the bottleneck (measured: 1039 clk per step).
struct Walk32 {
    int len, nb, tile_b, blk_b;     // range length, blocks per tile, (tile, block) of the range start
    int tile, blk, left;            // cursor; steps left before the walk wraps to the range start
    __device__ __forceinline__ Walk32(int64_t steps, int64_t nblocks, int64_t cta, int64_t grid)
    {
        const int64_t s_begin = (cta * steps) / grid, s_end = ((cta + 1) * steps) / grid;
        len = (int)(s_end - s_begin);
        nb = (int)nblocks;
        tile_b = (int)(s_begin / nblocks);
        blk_b = (int)(s_begin - (int64_t)tile_b * nblocks);
        const int rot = len > 0 ? (len - blk_b % len) % len : 0;
        tile = tile_b + (blk_b + rot) / nb;
        blk = (blk_b + rot) - (tile - tile_b) * nb;
        left = len - rot;
    }
};

// A maximal run of stages that needs no bookkeeping inside: one sample tile, consecutive SNP blocks, no drain.
// Every run ends with a drain of the accumulators (tile end, wrap of the rotated walk, or kFlushSteps stages).
struct Segment { int tile, blk, n; };

struct SegWalk {
    Walk32 w;
    int done = 0;
    __device__ __forceinline__ SegWalk(int64_t steps, int64_t nblocks, int64_t cta, int64_t grid) : w(steps, nblocks, cta, grid) {}
    __device__ __forceinline__ bool next(Segment &s)
    {
        if (done == w.len) return false;
        s.tile = w.tile; s.blk = w.blk;
        s.n = min(min(w.nb - w.blk, w.left), min(kFlushSteps, w.len - done));
        w.blk += s.n; w.left -= s.n; done += s.n;
        if (w.blk == w.nb) { w.blk = 0; w.tile++; }
        if (w.left == 0) { w.tile = w.tile_b; w.blk = w.blk_b; w.left = w.len; }
        return true;
    }
};

struct Ring {                       // position in the shared-memory ring
    int stage = 0;
    uint32_t parity = 0;
    __device__ __forceinline__ void next() { if (++stage == kStages) { stage = 0; parity ^= 1u; } }
};

template <int K, bool DENSE>
__global__ void __launch_bounds__(kThreads, 1)
pheno_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap panel_map)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
    const int64_t grid = gridDim.x, cta = blockIdx.x;
    SegWalk walk(p.steps, p.nblocks, cta, grid);
    const int len = walk.w.len;

    const uint32_t smem_base = pin(smem_u32(smem));
    const uint32_t bars = pin(smem_base + kStages * kStageBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kStages * kStageBytes + kNumBars * 8);
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(bar(kBarFull + s), 1); mbar_init(bar(kBarEmpty + s), kExpWarps + kSub); }
        for (int i = 0; i < kSub; i++) { mbar_init(bar(kBarAEmpty + i), 1); mbar_init(bar(kBarDFull + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarp) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = pin(*tmem_slot);
    Segment seg;
    constexpr int kWBytes = DENSE ? kHalves * kWTileBytes : kHalves * kWTileBytes / 2;     // weight tiles per stage
    constexpr int kStageMmas = DENSE ? kStageGroups : kStageGroups / 2;                    // K = 32 bytes: one group's c and m planes, or two groups' c planes

    if (warp == kProducerWarp) {
        // ============ producer warp: per stage two tensor copies (384 samples x 8 groups of words each; rows and
        // columns beyond the panel arrive as zeros) and one 4 KB bulk copy of the stage's weight tiles.  A bulk
        // copy costs its issuing warp ~200 clk whatever its size (measured: five 2 KB copies per 64 SNPs made
        // this warp the bottleneck at 1024 clk per step), so the copies are few and large.
        uint64_t pol_stream, pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&panel_map) : "memory");
            Ring ring;
            while (walk.next(seg)) {
                const int x = seg.tile * (kTile / 2);
                for (int i = 0; i < seg.n; i++, ring.next()) {
                    const uint32_t stage_base = smem_base + ring.stage * kStageBytes, full = bar(kBarFull + ring.stage);
                    const int blk = seg.blk + i;
                    mbar_wait_backoff(bar(kBarEmpty + ring.stage), ring.parity ^ 1u);
                    mbar_arrive_expect_tx(full, (uint32_t)(2 * kBoxBytes + kWBytes));
                    tensor_g2s(stage_base, &panel_map, x, blk * kStageGroups, full, pol_stream);
                    tensor_g2s(stage_base + kBoxBytes, &panel_map, x + kBoxSamples / 2, blk * kStageGroups, full, pol_stream);
                    bulk_g2s(stage_base + 2 * kBoxBytes, p.wt + (int64_t)blk * kWBytes, kWBytes, full, pol_keep);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kMmaWarp0) {
        // ============ MMA warps: warp kMmaWarp0 + t serves sub-tile t.  The whole warp runs the loop (uniform
        // operands -> uniform registers), one elected lane issues.  The hand-over from the expanders is a named
        // barrier of 160 threads on which the 128 expander threads only ARRIVE: they never wait for the issue.
        const int t = warp - kMmaWarp0;
        const uint32_t d_tmem = pin(tmem + kColD + 16 * t);
        const uint32_t a_tmem = pin(tmem + kColA + 64 * t);
        const uint32_t bar_aempty = pin(bar(kBarAEmpty + t)), bar_dfull = pin(bar(kBarDFull + t)), bar_id = pin(1 + t);
        const uint32_t wt0 = pin(smem_base + 2 * kBoxBytes);
        Ring ring;
        while (walk.next(seg)) {
            for (int i = 0; i < seg.n; i++, ring.next()) {
                tc_wait(bar(kBarFull + ring.stage), ring.parity);       // the stage's weight tiles have landed
                const uint32_t desc = wtile_desc_lo(wt0 + ring.stage * kStageBytes);
                asm volatile("bar.sync %0, 160;" ::"r"(bar_id) : "memory");    // A buffer written, D drained
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < kStageMmas; q++)        // the first MMA of a segment overwrites D
                    mma_i8_ts(d_tmem, a_tmem + 8 * q, desc + q * (512 >> 4), kWtileDescHi, kIdesc, (i == 0 && q == 0) ? 0u : 1u);
                mma_commit(bar_aempty);                      // A buffer free once these MMAs have read it
                mma_commit(bar(kBarEmpty + ring.stage));    // this sub-tile is done with the stage's weight tiles
            }
            mma_commit(bar_dfull);                           // the segment's accumulators are complete
        }
    } else {
        // ============ expander warps: sub-tile t = 4 warps = one independent pipeline ============
        const int t = warp >> 2;                                // sub-tile
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;         // this warp's quarter of the TMEM lanes
        const int my = t * 128 + (warp & 3) * 32 + lane;        // sample within the tile
        const uint32_t d_tmem = pin(tmem + kColD + 16 * t + lane_field);
        const uint32_t a_tmem = pin(tmem + kColA + 64 * t + lane_field);   // the sub-tile's A buffer: 64 columns = one stage
        const uint32_t bar_aempty = pin(bar(kBarAEmpty + t)), bar_dfull = pin(bar(kBarDFull + t)), bar_id = pin(1 + t);
        const uint32_t bar_full0 = pin(bar(kBarFull)), bar_empty0 = pin(bar(kBarEmpty));
        // this thread's word in a stage: box (384 samples) x group row (1536 B) x sample
        const uint32_t my_off = pin(smem_base + (uint32_t)((my / kBoxSamples) * kBoxBytes + (my % kBoxSamples) * 4));
        double y[K];
#pragma unroll
        for (int k = 0; k < K; k++) y[k] = 0.0;
        int escale[K];
#pragma unroll
        for (int k = 0; k < K; k++) escale[k] = p.scale[k].exponent - 6;
        uint32_t dphase = 0, aphase = 0;

        // read D back (complete once the MMA warp's commit has fired), recombine the digits exactly, add to y
        auto drain = [&]() {
            tc_wait(bar_dfull, dphase);
            dphase ^= 1u;
            tc_fence_after();
            uint32_t r[16];
            tmem_ld16(d_tmem, r);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const long long lo = (long long)(int)r[8 * k] + ((long long)(int)r[8 * k + 1] << 8) + ((long long)(int)r[8 * k + 2] << 16);
                const long long hi = (long long)(int)r[8 * k + 3] + ((long long)(int)r[8 * k + 4] << 8) + ((long long)(int)r[8 * k + 5] << 16);
                y[k] += ldexp((double)hi * 16777216.0 + (double)lo, escale[k]);
            }
        };
        auto store = [&](int tile) {
            const int64_t slot = cta - cta_of_step((int64_t)tile * p.nblocks, grid, p.steps);
#pragma unroll
            for (int k = 0; k < K; k++) {
                // +=: the rotated walk may visit the tile of its starting point twice (the slots start at 0)
                p.ypart[(slot * K + k) * p.npad + (int64_t)tile * kTile + my] += y[k];
                y[k] = 0.0;
            }
        };
        // this thread's 4 words of one half of a stage -> registers
        auto load_words = [&](int stage, int half, uint32_t (&w)[kStepGroups]) {
            const uint32_t src = my_off + stage * kStageBytes + half * (kStepGroups * kBoxRowBytes);
#pragma unroll
            for (int q = 0; q < kStepGroups; q++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[q]) : "r"(src + q * kBoxRowBytes));
        };
        // expand 4 words (64 SNPs) into 32 (c and m planes) or 16 (c planes only) columns of the A buffer
        auto expand_store = [&](const uint32_t (&w)[kStepGroups], int half) {
            uint32_t r[DENSE ? 32 : 16];
#pragma unroll
            for (int q = 0; q < kStepGroups; q++) {
                const uint32_t x = w[q];
                if (DENSE) {
                    const uint32_t hb = __umulhi(x, 0x80000000u);              // x >> 1 on the FMA pipe (the ALU pipe is the limit)
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        r[8 * q + s] = x & (0x03030303u << (2 * s));           // c << 2s
                        // [c == 3] << 2s = x & (x >> 1) & mask in ONE three-input LOP3
                        asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r[8 * q + 4 + s]) : "r"(x), "r"(hb), "r"(0x01010101u << (2 * s)));
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) r[4 * q + s] = x & (0x03030303u << (2 * s));
                }
            }
            if (half == 0) {
                tc_wait(bar_aempty, aphase ^ 1u);                // the MMAs of the previous stage have read the buffer
                tc_fence_after();
            }
            if constexpr (DENSE) tmem_st32(a_tmem + 32 * half, r);
            else tmem_st16(a_tmem + 16 * half, r);
        };

        Ring ring;
        uint32_t w0[kStepGroups], w1[kStepGroups];
        int remaining = len, prev_tile = -1;
        tc_wait(bar(kBarFull), 0);
        load_words(0, 0, w0);
        while (walk.next(seg)) {
            if (prev_tile >= 0) {                               // the previous segment's sums
                drain();
                if (seg.tile != prev_tile) store(prev_tile);
            }
            prev_tile = seg.tile;
            for (int i = 0; i < seg.n; i++) {
                load_words(ring.stage, 1, w1);                  // the stage's words are then all in registers
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty0 + 8 * ring.stage);
                expand_store(w0, 0);
                // the next stage's first half travels while the second half is expanded
                ring.next();
                if (--remaining > 0) {
                    tc_wait(bar_full0 + 8 * ring.stage, ring.parity);
                    load_words(ring.stage, 0, w0);
                }
                expand_store(w1, 1);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                asm volatile("bar.arrive %0, 160;" ::"r"(bar_id) : "memory");  // 128 rows x 128 SNPs are in TMEM (and D was drained)
                aphase ^= 1u;
            }
        }
        drain();
        store(prev_tile);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarp) tmem_dealloc(tmem, kTmemCols);
}

// y[k][i] = C_k + sum over the slots of the tile that holds sample i, in slot order
template <int K>
__global__ void __launch_bounds__(256)
tc_reduce_kernel(const double *__restrict__ ypart, int64_t npad, int64_t n, int slots, const TcScale *__restrict__ scale,
                 double *__restrict__ y1, double *__restrict__ y2, int accumulate,
                 const unsigned long long *__restrict__ corr /* [K][corr_stride] missing-genotype sums in quanta, or NULL */,
                 int64_t corr_stride)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = 0.0;
        for (int q = 0; q < slots; q++) s += ypart[((int64_t)q * K + k) * npad + i];
        s += scale[k].constant;
        if (corr) {
            s += ldexp((double)(long long)corr[k * corr_stride + i], scale[k].exponent);
        }
        if (accumulate) s += (k == 0) ? y1[i] : y2[i];
        if (k == 0) y1[i] = s; else y2[i] = s;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the panel as a 2-D tensor for the TMA: dimension 0 = 8-byte elements (2 samples) of a group, dimension 1 = groups
int make_panel_map(simu_b200_ctx *ctx, CUtensorMap *map)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
            return ctx_fail(ctx, SIMU_B200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)(ctx->s16_stride / 8), (cuuint64_t)((ctx->panel_rows + 15) / 16)};
    const cuuint64_t strides[1] = {(cuuint64_t)ctx->s16_stride};
    const cuuint32_t box[2] = {(cuuint32_t)(kBoxSamples / 2), (cuuint32_t)kStageGroups};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, ctx->panel, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ctx_fail(ctx, SIMU_B200_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SIMU_B200_OK;
}

template <int K>
int run_tc(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *d_x1, const double *d_x2, double *d_y1,
           double *d_y2, int accumulate)
{
    const int64_t n = ctx->n_samples;
    const int64_t tiles = (n + kTile - 1) / kTile;
    const int64_t nblocks = (mc + kStageSnps - 1) / kStageSnps;
    const int64_t steps = tiles * nblocks;
    const int64_t grid = steps < ctx->num_sms ? steps : ctx->num_sms;
    const int slots = (int)((grid + tiles - 1) / tiles) + 1;
    const int64_t npad = tiles * kTile;

    // the m plane as a list of the missing entries (miss_index.cu) when the panel's index is available
    const int sparse = miss_index_ensure(ctx);
    if (sparse < 0) return -sparse;
    long long *wm_fix = nullptr;
    unsigned long long *corr = nullptr;
    const int64_t corr_stride = sparse ? miss_corr_stride(ctx) : 0;
    if (sparse) {
        wm_fix = (long long *)ctx_scratch(ctx, SCR_WMFIX, (size_t)mc * K * sizeof(long long));
        const size_t corr_bytes = (size_t)K * corr_stride * sizeof(unsigned long long);
        corr = (unsigned long long *)ctx_scratch(ctx, SCR_CORR, corr_bytes);
        if (!wm_fix || !corr) return SIMU_B200_ENOMEM;
        // one warp per sample stores that sample's sum; only when a sample has more missing genotypes than one warp
        // takes do several warps add into a zeroed buffer
        if (miss_correct_accumulates(ctx)) CUDA_TRY(ctx, cudaMemsetAsync(corr, 0, corr_bytes, ctx->stream));
    }

    // weight tiles, then: [TcScale x 2][pad][partials 4 x kFusedThreads doubles]
    const size_t wt_bytes = (size_t)nblocks * kHalves * kWTileBytes;
    const size_t aux_off = (wt_bytes + 255) & ~(size_t)255;
    uint8_t *wt = (uint8_t *)ctx_scratch(ctx, SCR_TCW, aux_off + 64 + 4 * kFusedThreads * sizeof(double));
    const size_t ypart_bytes = (size_t)slots * K * npad * sizeof(double);
    double *ypart = (double *)ctx_scratch(ctx, SCR_YPART, ypart_bytes);
    if (!wt || !ypart) return SIMU_B200_ENOMEM;
    TcScale *scale = reinterpret_cast<TcScale *>(wt + aux_off);
    double *partial = reinterpret_cast<double *>(wt + aux_off + 64);
    // the fused weight kernel's barrier counters: zero between launches (the kernel resets them), cleared when first allocated
    const bool fbar_fresh = ctx->scratch_bytes[SCR_FUSED] < 64;
    unsigned int *fbar = (unsigned int *)ctx_scratch(ctx, SCR_FUSED, 64);
    if (!fbar) return SIMU_B200_ENOMEM;
    if (fbar_fresh) CUDA_TRY(ctx, cudaMemsetAsync(fbar, 0, 64, ctx->stream));
    if (!ctx->tc_attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(pheno_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        ctx->tc_attr_set = true;
    }
    CUtensorMap map;
    int rc = make_panel_map(ctx, &map);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(ypart, 0, ypart_bytes, ctx->stream));
    int ps = prof_begin(ctx, PROF_LUT_BUILD);
    if (mc <= kSmallMc) {
        tc_small_prepare_kernel<K><<<1, kSmallThreads, 0, ctx->stream>>>(d_freq, d_x1, d_x2, mc, nblocks * kHalves, scale, wt, wm_fix);
    } else {
        const int64_t items = nblocks * kHalves * 8 * K, most = items > mc ? items : mc;
        int64_t fgrid = (most + kFusedThreads - 1) / kFusedThreads;
        const int64_t fmax_grid = ctx->num_sms < kFusedThreads ? ctx->num_sms : kFusedThreads;     // co-resident, and one partial per thread
        if (fgrid > fmax_grid) fgrid = fmax_grid;
        const int64_t ntiles = nblocks * kHalves;
        const double *a_freq = d_freq, *a_x1 = d_x1, *a_x2 = d_x2;
        int64_t a_mc = mc;
        void *args[] = {(void *)&a_freq, (void *)&a_x1, (void *)&a_x2, (void *)&a_mc, (void *)&ntiles, (void *)&partial,
                        (void *)&fbar, (void *)&scale, (void *)&wt, (void *)&wm_fix};
        CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void *)tc_weights_fused_kernel<K>, dim3((unsigned)fgrid), dim3(kFusedThreads), args, 0, ctx->stream));
    }
    prof_end(ctx, ps);

    TcParams p;
    p.panel = ctx->panel; p.s16_stride = ctx->s16_stride; p.mc = mc; p.n_samples = n; p.wt = wt; p.scale = scale;
    p.ypart = ypart; p.npad = npad; p.tiles = tiles; p.nblocks = nblocks; p.steps = steps;
    // The pass over the missing entries (L2-bound gathers, 40 MB of DRAM traffic on config 2) runs on the side stream
    // BESIDE the main kernel (HBM-bound, one CTA per SM with room for a 256-thread CTA next to it): fork after the weights,
    // join before the reduce.  Launched after the main kernel so that its CTAs fill the gaps instead of the SMs.
    if (sparse) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    ps = prof_begin(ctx, PROF_LUT_MAIN);
    if (sparse) pheno_tc_kernel<K, false><<<(unsigned)grid, kThreads, kSmemBytes, ctx->stream>>>(p, map);
    else pheno_tc_kernel<K, true><<<(unsigned)grid, kThreads, kSmemBytes, ctx->stream>>>(p, map);
    prof_end(ctx, ps);
    if (sparse) {
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        if ((rc = launch_miss_correct(ctx, ctx->copy_stream, mc, K, wm_fix, corr))) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->copy_stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    ps = prof_begin(ctx, PROF_LUT_REDUCE);
    tc_reduce_kernel<K><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ypart, npad, n, slots, scale, d_y1, d_y2, accumulate, corr, corr_stride);
    prof_end(ctx, ps);
    ctx->launches += 3;
    return ctx_cuda(ctx, cudaGetLastError(), "pheno_tc launch");
}

}  // namespace

int launch_pheno_tc(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *d_x1, const double *d_x2,
                    double *d_y1, double *d_y2, int accumulate)
{
    if (mc <= 0 && accumulate) return SIMU_B200_OK;
    if (mc <= 0) {   // source.cc:962-966: phenotypes start at zero
        CUDA_TRY(ctx, cudaMemsetAsync(d_y1, 0, (size_t)ctx->n_samples * sizeof(double), ctx->stream));
        if (d_y2) CUDA_TRY(ctx, cudaMemsetAsync(d_y2, 0, (size_t)ctx->n_samples * sizeof(double), ctx->stream));
        return SIMU_B200_OK;
    }
    if (d_x2 && d_y2) return run_tc<2>(ctx, mc, d_freq, d_x1, d_x2, d_y1, d_y2, accumulate);
    return run_tc<1>(ctx, mc, d_freq, d_x1, nullptr, d_y1, nullptr, accumulate);
}
