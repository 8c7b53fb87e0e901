// Kernel (b), EXACT mode: y[i] += ((2 - code_ij) - 2 f_j) * x_jk in FP64, per
// sample in ascending causal-SNP order with separate subtract / multiply / add
// -- the arithmetic of /root/reference/src/source.cc:986-996 (an x86-64 build
// without -march has no FMA contraction), so the result is bit-identical to
// the reference on one GPU.
//
// One thread owns one packed byte column = 4 samples and walks every causal
// row in order.  The per-SNP products (a - avg) * x for a in {2,1,0} are the
// same three doubles for every sample, so each block computes them once per
// 64-SNP chunk into shared memory (index = raw 2-bit field: 00 -> a=2,
// 01 -> missing -> +0.0, 10 -> a=1, 11 -> a=0) and the inner loop is one
// shared-memory gather plus one DADD per update.  Adding +0.0 for a missing
// genotype equals the reference's `continue` because the accumulator can
// never be -0.0 (it starts at +0.0, and x + (-x) rounds to +0.0).
//
// This mode is FP64-pipe / shared-memory bound, not HBM bound; it is the
// parity anchor and the CLI's default.  The FAST mode (pheno_tc.cu, pheno_lut.cu)
// is the throughput path.  The SAMPLE16 kernels (the product layout) are at the end.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <new>

#include "ctx.cuh"

namespace {

constexpr int kChunk = 64;      // SNPs per shared-memory table refill
constexpr int kThreads = 128;   // byte columns per block (512 samples)
constexpr int kPrefetch = 8;    // row bytes fetched ahead per thread

// SPT = samples per thread (1, 2 or 4: a quarter, half or whole packed byte).  Fewer samples
// per thread means more warps in flight to hide the shared-memory gather -> DADD dependency.
template <int K, int SPT>
__global__ void __launch_bounds__(kThreads)
pheno_exact_kernel(const uint8_t *__restrict__ panel, int64_t row_stride, const int32_t *__restrict__ rows,
                   int64_t mc, int64_t n_samples, const double *__restrict__ freq,
                   const double *__restrict__ x1, const double *__restrict__ x2,
                   double *__restrict__ y1, double *__restrict__ y2)
{
    __shared__ double tbl[kChunk][4][K];      // [snp][raw bits][trait]
    __shared__ int64_t row_off[kChunk];

    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SPT;   // first sample of this thread
    const bool active = first < n_samples;
    const uint8_t *base = panel + (active ? (first >> 2) : 0);
    const int shift0 = (int)(first & 3) * 2;

    double acc[SPT][K];
#pragma unroll
    for (int s = 0; s < SPT; s++)
#pragma unroll
        for (int k = 0; k < K; k++) acc[s][k] = 0.0;

    for (int64_t j0 = 0; j0 < mc; j0 += kChunk) {
        const int nj = (int)min((int64_t)kChunk, mc - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nj; t += blockDim.x) {
            const int64_t j = j0 + t;
            const double avg = __dmul_rn(2.0, freq[j]);          // source.cc:986
            row_off[t] = (rows ? (int64_t)rows[j] : j) * row_stride;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double x = (k == 0) ? x1[j] : x2[j];
                tbl[t][0][k] = __dmul_rn(__dsub_rn(2.0, avg), x);   // bits 00: two A1 copies
                tbl[t][1][k] = 0.0;                                 // bits 01: missing (:991)
                tbl[t][2][k] = __dmul_rn(__dsub_rn(1.0, avg), x);   // bits 10: one copy
                tbl[t][3][k] = __dmul_rn(__dsub_rn(0.0, avg), x);   // bits 11: none
            }
        }
        __syncthreads();
        for (int jj = 0; jj < nj; jj += kPrefetch) {
            uint32_t g[kPrefetch];
#pragma unroll
            for (int u = 0; u < kPrefetch; u++)
                g[u] = (jj + u < nj) ? ((uint32_t)__ldg(base + row_off[jj + u]) >> shift0) : 0u;
#pragma unroll
            for (int u = 0; u < kPrefetch; u++) {
                if (jj + u < nj) {
#pragma unroll
                    for (int s = 0; s < SPT; s++) {
                        const uint32_t v = (g[u] >> (2 * s)) & 3u;
#pragma unroll
                        for (int k = 0; k < K; k++)
                            acc[s][k] = __dadd_rn(acc[s][k], tbl[jj + u][v][k]);   // source.cc:994-995
                    }
                }
            }
        }
    }

    if (active) {
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            const int64_t i = first + s;
            if (i < n_samples) {
                y1[i] = acc[s][0];
                if (K == 2) y2[i] = acc[s][K - 1];
            }
        }
    }
}

// GROUP4 layout: byte i of group g holds the 2-bit fields of sample i for rows 4g..4g+3, so a
// thread reads ONE byte per sample per four SNPs and adds the four contributions in ascending
// SNP order -- the same per-sample order as above, hence the same bits.
template <int K, int SPT>
__global__ void __launch_bounds__(kThreads)
pheno_exact_g4_kernel(const uint8_t *__restrict__ panel, int64_t group_stride, int64_t mc, int64_t n_samples,
                      const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2,
                      double *__restrict__ y1, double *__restrict__ y2)
{
    __shared__ double tbl[kChunk][4][K];      // [snp][raw bits][trait]
    static_assert(SPT == 1 || SPT == 2 || SPT == 4, "SPT");
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * SPT;
    const bool active = first < n_samples;
    const uint8_t *base = panel + (active ? first : 0);    // group_stride >= round_up(N, 128): SPT bytes are readable

    double acc[SPT][K];
#pragma unroll
    for (int s = 0; s < SPT; s++)
#pragma unroll
        for (int k = 0; k < K; k++) acc[s][k] = 0.0;

    for (int64_t j0 = 0; j0 < mc; j0 += kChunk) {
        const int nj = (int)min((int64_t)kChunk, mc - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nj; t += blockDim.x) {
            const int64_t j = j0 + t;
            const double avg = __dmul_rn(2.0, freq[j]);          // source.cc:986
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double x = (k == 0) ? x1[j] : x2[j];
                tbl[t][0][k] = __dmul_rn(__dsub_rn(2.0, avg), x);   // bits 00: two A1 copies
                tbl[t][1][k] = 0.0;                                 // bits 01: missing (:991)
                tbl[t][2][k] = __dmul_rn(__dsub_rn(1.0, avg), x);   // bits 10: one copy
                tbl[t][3][k] = __dmul_rn(__dsub_rn(0.0, avg), x);   // bits 11: none
            }
        }
        __syncthreads();
        const int ng = (nj + 3) >> 2;                              // groups in this chunk (kChunk % 4 == 0)
        const uint8_t *gp = base + (j0 >> 2) * group_stride;
        constexpr int kAhead = 8;
        for (int gg = 0; gg < ng; gg += kAhead) {
            uint32_t b[kAhead];
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                b[u] = 0u;
                if (gg + u < ng) {
                    const uint8_t *q = gp + (int64_t)(gg + u) * group_stride;
                    if (SPT == 1) b[u] = __ldg(q);
                    else if (SPT == 2) b[u] = __ldg(reinterpret_cast<const uint16_t *>(q));
                    else b[u] = __ldg(reinterpret_cast<const uint32_t *>(q));
                }
            }
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int jj = 4 * (gg + u) + r;
                    if (jj < nj) {
#pragma unroll
                        for (int s = 0; s < SPT; s++) {
                            const uint32_t v = (b[u] >> (8 * s + 2 * r)) & 3u;
#pragma unroll
                            for (int k = 0; k < K; k++)
                                acc[s][k] = __dadd_rn(acc[s][k], tbl[jj][v][k]);   // source.cc:994-995
                        }
                    }
                }
            }
        }
    }

    if (active) {
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            const int64_t i = first + s;
            if (i < n_samples) {
                y1[i] = acc[s][0];
                if (K == 2) y2[i] = acc[s][K - 1];
            }
        }
    }
}

// SAMPLE16 layout: word i of group g holds sample i's dosage codes (0 hom A1, 1 het, 2 hom A2, 3 missing)
// for rows 16g..16g+15, so a thread owns ONE sample, reads one coalesced 32-bit word per 16 SNPs and adds
// the 16 contributions in ascending SNP order -- the same per-sample order and operations as above, hence
// the same bits.  The table is indexed by the dosage code: entry 3 (missing) is +0.0.
template <int K>
__global__ void __launch_bounds__(kThreads)
pheno_exact_s16_kernel(const uint8_t *__restrict__ panel, int64_t s16_stride, int64_t mc, int64_t n_samples,
                       const double *__restrict__ freq, const double *__restrict__ x1, const double *__restrict__ x2,
                       double *__restrict__ y1, double *__restrict__ y2, int accumulate)
{
    __shared__ __align__(16) double tbl[kChunk][4][K];      // [snp][dosage code][trait]
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n_samples;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(panel) + (active ? i : 0);
    const int64_t gw = s16_stride >> 2;                        // words per group
    // accumulate: y holds the sums over the causal rows of earlier panels (batched runs): the per-sample chain of
    // additions simply continues, so the result is bit-identical to one pass over all rows
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = (accumulate && active) ? (k == 0 ? y1[i] : y2[i]) : 0.0;

    uint32_t w[4], wn[4];
#pragma unroll
    for (int q = 0; q < 4; q++) w[q] = (16 * q < mc) ? __ldg(base + q * gw) : 0u;
    for (int64_t j0 = 0; j0 < mc; j0 += kChunk) {
        const int nj = (int)min((int64_t)kChunk, mc - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nj; t += blockDim.x) {
            const int64_t j = j0 + t;
            const double avg = __dmul_rn(2.0, freq[j]);          // source.cc:986
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double x = (k == 0) ? x1[j] : x2[j];
                tbl[t][0][k] = __dmul_rn(__dsub_rn(2.0, avg), x);   // code 0: two A1 copies
                tbl[t][1][k] = __dmul_rn(__dsub_rn(1.0, avg), x);   // code 1: one copy
                tbl[t][2][k] = __dmul_rn(__dsub_rn(0.0, avg), x);   // code 2: none
                tbl[t][3][k] = 0.0;                                 // code 3: missing (:991)
            }
        }
        // the next chunk's words travel while this chunk is added
#pragma unroll
        for (int q = 0; q < 4; q++) wn[q] = (j0 + kChunk + 16 * q < mc) ? __ldg(base + ((j0 + kChunk) / 16 + q) * gw) : 0u;
        __syncthreads();
        const char *tb = reinterpret_cast<const char *>(&tbl[0][0][0]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int f = 0; f < 16; f++) {
                const int jj = 16 * q + f;
                if (jj < nj) {
                    const uint32_t off = ((w[q] >> (2 * f)) & 3u) * (8u * K);
                    const double *e = reinterpret_cast<const double *>(tb + jj * (32 * K) + off);
#pragma unroll
                    for (int k = 0; k < K; k++) acc[k] = __dadd_rn(acc[k], e[k]);   // source.cc:994-995
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = wn[q];
    }
    if (active) {
        y1[i] = acc[0];
        if (K == 2) y2[i] = acc[K - 1];
    }
}


// SAMPLE16, second form (the default): the EFFECTS come from the kernel-parameter constant bank, TWO samples per thread.
//
// The kernel above gathers the finished addend (a - 2f) x_k per trait: 16 bytes per sample and SNP for two
// traits through the 128 B/clk shared-memory pipe -- 2e10 gathers are 4.4 ms on config 2 whatever else happens
// (measured 7.4 ms).  Here the table holds only t = a - 2f (8 bytes, the same for both traits) and the
// multiplication by x_k is done per sample with x_k as a CONSTANT-BANK operand: a warp-uniform value costs no
// shared-memory or register-file bandwidth there (SASS: LDCU.64 + DMUL R, R, UR).  The arithmetic is the reference's,
// operation for operation (source.cc:986-996: `(a - avg) * x` then `+=`): __dsub_rn once per SNP, __dmul_rn and
// __dadd_rn per sample -- bit-identical.  A missing genotype has t = +0.0: the product is +-0.0 and adding it leaves
// the accumulator unchanged (it can never be -0.0), which equals the reference's `continue` for finite effects.
// The effects of a slice of SNPs travel as a __grid_constant__ kernel parameter (24 KB of the 32 KB limit), so
// a call is a chain of launches, each continuing every sample's sum from y (stream order keeps the SNP order).
//
// ncu on the first version of this form (one sample per thread, 64-thread blocks; profiles/r02f_exact_s16p_ncu.txt):
// 10.8 issued instructions per 32 samples and SNP at 71 % of the issue slots, FP64 pipe 51 %, shared-memory pipe 52 % -- issue-bound before either data pipe.  Per 64
// samples and SNP this form issues 2 LDS.64 + 2 PRMT (the table offsets: the codes of a word are multiplied by 8 four fields
// at a time, 8 shifts + 8 LOP3 per 16 SNPs, and a byte permute picks one -- 3.98 -> 3.81 ms against a shift + mask per SNP)
// + 2 LDCU.64 (the traits' effects, adjacent in the parameter) + 4 DMUL + 4 DADD:
// the uniform loads, the loop bookkeeping and the table build are shared by twice the samples, and a warp carries
// four independent addition chains instead of two.  Same operations in the same order per sample -> same bits.
// One warp per block keeps the block count at N/64 (1563 for 100K samples = 10.6 per SM; 64-thread blocks of two
// samples per thread would be 5.3 per SM, a 14 % tail).
template <int K> struct alignas(16) ExactEffectsI { static constexpr int kSnps = (K == 2) ? 1536 : 3072; double x[kSnps][K]; };

template <int K>
__global__ void __launch_bounds__(64)   // launched with 32: ptxas gives up the uniform datapath (LDCU, DMUL R, R, UR) for kernels bounded to one warp
pheno_exact_s16d_kernel(const uint8_t *__restrict__ panel, int64_t s16_stride, int mc, int64_t n_samples,
                        const double *__restrict__ freq, double *__restrict__ y1, double *__restrict__ y2, int accumulate, int pdl,
                        const __grid_constant__ ExactEffectsI<K> eff)
{
    __shared__ __align__(16) double tbl[2][kChunk][4];        // [buffer][snp][dosage code] -> a - 2f, missing -> +0.0
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;        // samples i0, i0 + 1 (the pitch covers round_up(N, 128))
    const bool act0 = i0 < n_samples, act1 = i0 + 1 < n_samples;
    const uint2 *base = reinterpret_cast<const uint2 *>(panel) + (act0 ? (i0 >> 1) : 0);
    const int64_t gw = s16_stride >> 3;                       // 8-byte words per group
    // the frequencies of a chunk are loaded one chunk ahead (two per thread of a 32-thread block) and turned into the
    // table at the END of the chunk before: ncu showed the load -> first use latency of a build at the top of a chunk
    // as the kernel's top stall (14 % of the samples) with 2.6 warps per scheduler
    double fr[2];
    auto load_freq = [&](int j0) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int t = threadIdx.x + u * blockDim.x;
            fr[u] = (t < kChunk && j0 + t < mc) ? __ldg(freq + j0 + t) : 0.0;
        }
    };
    auto build = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int t = threadIdx.x + u * blockDim.x;
            if (t < kChunk) {
                const double avg = __dmul_rn(2.0, fr[u]);     // source.cc:986
                *reinterpret_cast<double2 *>(&tbl[buf][t][0]) = make_double2(__dsub_rn(2.0, avg), __dsub_rn(1.0, avg));
                *reinterpret_cast<double2 *>(&tbl[buf][t][2]) = make_double2(__dsub_rn(0.0, avg), 0.0);
            }
        }
    };
    uint2 w[4], wn[4];
#pragma unroll
    for (int q = 0; q < 4; q++) w[q] = (16 * q < mc) ? __ldg(base + q * gw) : make_uint2(0u, 0u);
    load_freq(0);
    build(0);
    // launched with programmatic stream serialisation behind the previous slice of the same call (pdl != 0): the launch
    // itself, the frequency loads and the first table overlap that slice's tail (ptxas sinks the word loads below the
    // wait); the previous slice's sums are visible after the wait
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    double acc[2][K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        acc[0][k] = (accumulate && act0) ? (k == 0 ? y1[i0] : y2[i0]) : 0.0;
        acc[1][k] = (accumulate && act1) ? (k == 0 ? y1[i0 + 1] : y2[i0 + 1]) : 0.0;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int buf = 0;
    for (int j0 = 0; j0 < mc; j0 += kChunk, buf ^= 1) {
        const int nj = min(kChunk, mc - j0);
        __syncthreads();                                        // table `buf` is complete, table `buf ^ 1` is no longer read
        load_freq(j0 + kChunk);
#pragma unroll
        for (int q = 0; q < 4; q++)
            wn[q] = (j0 + kChunk + 16 * q < mc) ? __ldg(base + ((j0 + kChunk) / 16 + q) * gw) : make_uint2(0u, 0u);
        const char *tb = reinterpret_cast<const char *>(&tbl[buf][0][0]);
        if (nj == kChunk) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                // the codes times 8, four fields per word: byte b of u[k] is the table offset of SNP 4b + k
                const uint32_t u0[4] = {(w[q].x << 3) & 0x18181818u, (w[q].x << 1) & 0x18181818u, (w[q].x >> 1) & 0x18181818u, (w[q].x >> 3) & 0x18181818u};
                const uint32_t u1[4] = {(w[q].y << 3) & 0x18181818u, (w[q].y << 1) & 0x18181818u, (w[q].y >> 1) & 0x18181818u, (w[q].y >> 3) & 0x18181818u};
#pragma unroll
                for (int f = 0; f < 16; f++) {
                    const int jj = 16 * q + f;
                    // byte offset of the table entry: the 2-bit code times 8
                    const uint32_t o0 = __byte_perm(u0[f & 3], 0u, 0x4440u + (f >> 2));
                    const uint32_t o1 = __byte_perm(u1[f & 3], 0u, 0x4440u + (f >> 2));
                    const double t0 = *reinterpret_cast<const double *>(tb + jj * 32 + o0);
                    const double t1 = *reinterpret_cast<const double *>(tb + jj * 32 + o1);
                    double x[2];
                    if (K == 2) *reinterpret_cast<double2 *>(x) = *reinterpret_cast<const double2 *>(&eff.x[j0 + jj][0]);   // adjacent in the parameter (ptxas: two LDCU.64)
                    else x[0] = eff.x[j0 + jj][0];
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        acc[0][k] = __dadd_rn(acc[0][k], __dmul_rn(t0, x[k]));   // source.cc:994-995
                        acc[1][k] = __dadd_rn(acc[1][k], __dmul_rn(t1, x[k]));
                    }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                for (int f = 0; f < 16 && 16 * q + f < nj; f++) {       // w[q] with a static index: the words stay in registers
                    const int jj = 16 * q + f;
                    const uint32_t o0 = ((w[q].x >> (2 * f)) & 3u) * 8u, o1 = ((w[q].y >> (2 * f)) & 3u) * 8u;
                    const double t0 = *reinterpret_cast<const double *>(tb + jj * 32 + o0);
                    const double t1 = *reinterpret_cast<const double *>(tb + jj * 32 + o1);
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const double x = eff.x[j0 + jj][k];
                        acc[0][k] = __dadd_rn(acc[0][k], __dmul_rn(t0, x));
                        acc[1][k] = __dadd_rn(acc[1][k], __dmul_rn(t1, x));
                    }
                }
            }
        }
        build(buf ^ 1);
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = wn[q];
    }
    if (act0) { y1[i0] = acc[0][0]; if (K == 2) y2[i0] = acc[0][K - 1]; }
    if (act1) { y1[i0 + 1] = acc[1][0]; if (K == 2) y2[i0 + 1] = acc[1][K - 1]; }
}

template <int K, int SPT>
void launch_exact_t(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq, const double *d_x1,
                    const double *d_x2, double *d_y1, double *d_y2)
{
    // small blocks: the block count should be several times the SM count, or the last wave is
    // half empty (N = 100K with 128-thread blocks gives 196 blocks on 148 SMs: 2x the time)
    int threads = 128;
    if (const char *e = getenv("SIMU_B200_EXACT_THREADS")) threads = atoi(e);
    const int64_t per_block = (int64_t)threads * SPT;
    const int64_t blocks = (ctx->n_samples + per_block - 1) / per_block;
    if (ctx->layout == SIMU_B200_LAYOUT_GROUP4)
        pheno_exact_g4_kernel<K, SPT><<<(unsigned)blocks, threads, 0, ctx->stream>>>(
            ctx->panel, ctx->group_stride, mc, ctx->n_samples, d_freq, d_x1, d_x2, d_y1, d_y2);
    else
        pheno_exact_kernel<K, SPT><<<(unsigned)blocks, threads, 0, ctx->stream>>>(
            ctx->panel, ctx->row_stride, d_rows, mc, ctx->n_samples, d_freq, d_x1, d_x2, d_y1, d_y2);
}

template <int K>
int launch_exact_s16d(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *hx1, const double *hx2, double *d_y1,
                      double *d_y2, int accumulate)
{
    using Eff = ExactEffectsI<K>;
    const int64_t blocks = (ctx->n_samples + 63) / 64;
    std::unique_ptr<Eff> effp(new (std::nothrow) Eff);        // host staging of one launch's parameters
    if (!effp) return ctx_fail(ctx, SIMU_B200_ENOMEM, "find_pheno: out of host memory");
    Eff &eff = *effp;
    memset(&eff, 0, sizeof(eff));
    static const bool use_pdl = getenv("SIMU_B200_EXACT_NOPDL") == nullptr;     // A/B runs
    for (int64_t j0 = 0; j0 < mc; j0 += Eff::kSnps) {
        const int n = (int)std::min<int64_t>(Eff::kSnps, mc - j0);
        for (int j = 0; j < n; j++) {
            eff.x[j][0] = hx1[j0 + j];
            if (K == 2) eff.x[j][K - 1] = hx2[j0 + j];
        }
        // every slice after the first overlaps its prologue (genotype words, frequencies, first table) with the tail of
        // the slice before it; the first one follows kernels whose output it reads in that prologue and is launched plainly
        const int pdl = (j0 > 0 && use_pdl) ? 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, pheno_exact_s16d_kernel<K>, (const uint8_t *)(ctx->panel + (j0 / 16) * ctx->s16_stride),
                                         (int64_t)ctx->s16_stride, n, (int64_t)ctx->n_samples, (const double *)(d_freq + j0), d_y1, d_y2,
                                         (accumulate || j0 > 0) ? 1 : 0, pdl, eff));
        ctx->launches++;
    }
    return SIMU_B200_OK;
}

}  // namespace

int launch_pheno_exact(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                       const double *d_x1, const double *d_x2, double *d_y1, double *d_y2)
{
    // measured on B200 (config 2, K=2): 13.6-17 ms for every combination of 1/2/4 samples per
    // thread and 32/64/128 threads per block -- the kernel is bound by its per-update
    // shared-memory gather + FP64 add chain, not by occupancy.  SIMU_B200_EXACT_SPT /
    // SIMU_B200_EXACT_THREADS override for experiments.
    int spt = 2;
    if (const char *e = getenv("SIMU_B200_EXACT_SPT")) spt = atoi(e);
    const bool k2 = d_x2 && d_y2;
    const int ps = prof_begin(ctx, PROF_PHENO_EXACT);
    if (k2) {
        if (spt == 4) launch_exact_t<2, 4>(ctx, d_rows, mc, d_freq, d_x1, d_x2, d_y1, d_y2);
        else if (spt == 2) launch_exact_t<2, 2>(ctx, d_rows, mc, d_freq, d_x1, d_x2, d_y1, d_y2);
        else launch_exact_t<2, 1>(ctx, d_rows, mc, d_freq, d_x1, d_x2, d_y1, d_y2);
    } else {
        if (spt == 4) launch_exact_t<1, 4>(ctx, d_rows, mc, d_freq, d_x1, nullptr, d_y1, nullptr);
        else if (spt == 2) launch_exact_t<1, 2>(ctx, d_rows, mc, d_freq, d_x1, nullptr, d_y1, nullptr);
        else launch_exact_t<1, 1>(ctx, d_rows, mc, d_freq, d_x1, nullptr, d_y1, nullptr);
    }
    prof_end(ctx, ps);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "pheno_exact_kernel launch");
}

int launch_pheno_exact_s16(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *d_x1, const double *d_x2,
                           double *d_y1, double *d_y2, int accumulate)
{
    const bool k2 = d_x2 && d_y2;
    // the parameter-bank kernel needs the effects on the HOST at launch time: the host-buffer entry points leave
    // them in ctx->h_x1 / h_x2; a device-resident caller's effects are fetched once per call (Mc x K doubles), which
    // needs a stream synchronisation and is therefore not possible while the stream is being captured into a graph
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cap);
    static const bool table_kernel = getenv("SIMU_B200_EXACT_TABLE") != nullptr;      // A/B runs: the 16-byte gather kernel
    const double *hx1 = ctx->h_x1, *hx2 = ctx->h_x2;
    bool param_kernel = !table_kernel && mc > 0 && (hx1 || cap == cudaStreamCaptureStatusNone);
    if (param_kernel && !hx1) {
        const size_t bytes = (size_t)mc * sizeof(double);
        if (ctx->h_eff_bytes < 2 * bytes) {
            if (ctx->h_eff) cudaFreeHost(ctx->h_eff);
            ctx->h_eff = nullptr; ctx->h_eff_bytes = 0;
            CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->h_eff, 2 * bytes, cudaHostAllocDefault));
            ctx->h_eff_bytes = 2 * bytes;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_eff, d_x1, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (k2) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_eff + mc, d_x2, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        hx1 = ctx->h_eff; hx2 = k2 ? ctx->h_eff + mc : nullptr;
    }
    const int ps = prof_begin(ctx, PROF_PHENO_EXACT);
    if (param_kernel) {
        const int rc = k2 ? launch_exact_s16d<2>(ctx, mc, d_freq, hx1, hx2, d_y1, d_y2, accumulate)
                          : launch_exact_s16d<1>(ctx, mc, d_freq, hx1, nullptr, d_y1, nullptr, accumulate);
        if (rc) return rc;
    } else {
        const int64_t blocks = (ctx->n_samples + kThreads - 1) / kThreads;
        if (k2)
            pheno_exact_s16_kernel<2><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(ctx->panel, ctx->s16_stride, mc, ctx->n_samples,
                                                                                      d_freq, d_x1, d_x2, d_y1, d_y2, accumulate);
        else
            pheno_exact_s16_kernel<1><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(ctx->panel, ctx->s16_stride, mc, ctx->n_samples,
                                                                                      d_freq, d_x1, nullptr, d_y1, nullptr, accumulate);
        ctx->launches++;
    }
    prof_end(ctx, ps);
    return ctx_cuda(ctx, cudaGetLastError(), "pheno_exact_s16_kernel launch");
}
