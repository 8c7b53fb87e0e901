// Missing-genotype index of a SAMPLE16 panel, and the correction pass that uses it.
//
// find_pheno's tensor-core kernel (pheno_tc.cu) evaluates  y_i = C + sum_j Wp_j c_ij + sum_j Wm_j [c_ij == 3]
// (source.cc:986-996 restated over the dosage code c).  The second sum is over the MISSING genotypes only -- 0.1 %
// of the entries of the SURVEY.md 8(d) panel -- yet as a dense "m plane" it doubled the bytes through tensor memory
// and the number of MMAs, and held the kernel at 0.72 of the HBM roofline (measured: 0.547 ms with, 0.375 ms without
// the plane on config 2).  The index below lists the missing entries once per panel, as part of staging: which
// entries are missing depends on the panel only, not on effects or frequencies.  Each call then adds
//     corr_i = sum over sample i's missing entries j of round(Wm_j / quantum)          (64-bit integers)
// in a separate pass over the list (1.6 % of the panel's bytes at a 0.1 % missing rate).  Integer sums are exact
// and order-independent, so the result is the same exact fixed-point sum the dense plane produces; the two forms
// differ only in where integers become FP64 (per drained segment there, once per sample here): equal to FP64
// rounding, ~1e-16 of max|y|.
//
// Layout: compressed rows BY SAMPLE: offsets[i] .. offsets[i + 1] delimit sample i's entries, an entry is the
// panel row of a missing genotype (4 bytes, in no particular order).  The correction pass gives every sample a
// warp: coalesced reads of the segment, one gather of the row's weight per entry, sums in registers, a shuffle
// reduction and ONE plain store -- no atomics (shared-memory atomics sustain ~2.4 per clock and SM, measured: a
// bucketed version of this pass took 87 us for config 2's 1e7 entries).  Panels with more than kMaxRate missing
// entries keep the dense m plane (the index is then marked "dense").  The index is built lazily, at the SECOND
// FAST call after a write to the panel (two scans of the panel: count per sample, host prefix sum, fill; the first
// call runs with the dense plane), and never while the stream is being captured.
#include <stdlib.h>

#include <algorithm>

#include "ctx.cuh"

namespace {

constexpr double kMaxRate = 0.004;             // above this the dense plane is cheaper than the list
constexpr int kSeg = 2048;                     // entries per warp of the correction pass (longer segments are split)

__device__ __forceinline__ uint32_t missing_bits(uint32_t x) { return x & (x >> 1) & 0x55555555u; }   // code 3 = both bits

// pass 1: missing entries per sample.  grid.x covers the samples (256 per block), grid.y strides over the groups
__global__ void __launch_bounds__(256)
miss_count_kernel(const uint32_t *__restrict__ panel, int64_t words_per_group, int64_t ngroups, unsigned long long *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= words_per_group) return;
    unsigned n = 0;
    int64_t g = blockIdx.y;
    for (; g + 3 * (int64_t)gridDim.y < ngroups; g += 4 * (int64_t)gridDim.y) {           // four loads in flight
        uint32_t x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) x[u] = __ldg(panel + (g + u * (int64_t)gridDim.y) * words_per_group + i);
#pragma unroll
        for (int u = 0; u < 4; u++) n += __popc(missing_bits(x[u]));
    }
    for (; g < ngroups; g += gridDim.y) n += __popc(missing_bits(__ldg(panel + g * words_per_group + i)));
    if (n) atomicAdd(&counts[i], (unsigned long long)n);
}

// pass 2: the entries, appended to their sample's segment
__global__ void __launch_bounds__(256)
miss_fill_kernel(const uint32_t *__restrict__ panel, int64_t words_per_group, int64_t ngroups, unsigned long long *__restrict__ cursor,
                 uint32_t *__restrict__ entries)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= words_per_group) return;
    for (int64_t g0 = blockIdx.y; g0 < ngroups; g0 += 4 * (int64_t)gridDim.y) {
        uint32_t xs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {                                                      // four loads in flight
            const int64_t gu = g0 + u * (int64_t)gridDim.y;
            xs[u] = gu < ngroups ? __ldg(panel + gu * words_per_group + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t mm = missing_bits(xs[u]);
            if (!mm) continue;
            const int64_t g = g0 + u * (int64_t)gridDim.y;
            unsigned long long pos = atomicAdd(&cursor[i], (unsigned long long)__popc(mm));
            while (mm) {
                const int f = __ffs(mm) - 1;
                mm &= mm - 1;
                entries[pos++] = (uint32_t)(16 * g + (f >> 1));
            }
        }
    }
}

// corr[k][i] (+)= sum over sample i's missing entries j < mc of wm[j][k]: warp (i, c) takes entries [c kSeg, (c+1) kSeg)
// of sample i.  SPLIT = false: every segment fits one warp, which stores its sum; SPLIT = true (a sample with more
// than kSeg missing genotypes exists): corr is zero on entry and the warps add to it.
template <int K, bool SPLIT>
__global__ void __launch_bounds__(256)
miss_correct_kernel(const uint32_t *__restrict__ entries, const unsigned long long *__restrict__ offsets, int64_t n_samples, int64_t mc,
                    const long long *__restrict__ wm, unsigned long long *__restrict__ corr, int64_t corr_stride)
{
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_samples) return;
    unsigned long long lo = offsets[i];
    const unsigned long long end = offsets[i + 1];
    if (SPLIT) {
        lo += (unsigned long long)blockIdx.y * kSeg;
        if (lo >= end) return;
    }
    const unsigned long long hi = SPLIT ? (lo + kSeg < end ? lo + kSeg : end) : end;
    long long sum[K];
#pragma unroll
    for (int k = 0; k < K; k++) sum[k] = 0;
    for (unsigned long long t = lo + lane; t < hi; t += 64) {              // two entries per lane in flight
        const uint32_t j0 = __ldg(entries + t);
        const uint32_t j1 = t + 32 < hi ? __ldg(entries + t + 32) : 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const long long a = (int64_t)j0 < mc ? __ldg(wm + (int64_t)j0 * K + k) : 0ll;
            const long long b = (int64_t)j1 < mc ? __ldg(wm + (int64_t)j1 * K + k) : 0ll;
            sum[k] += a + b;
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], d);
        if (lane == 0) {
            if (SPLIT) { if (sum[k]) atomicAdd(&corr[k * corr_stride + i], (unsigned long long)sum[k]); }
            else corr[k * corr_stride + i] = (unsigned long long)sum[k];
        }
    }
}

}  // namespace

void miss_index_invalidate(simu_b200_ctx *ctx) { ctx->miss_state = 0; ctx->miss_uses = 0; }

// 1: the sparse index is ready; 0: use the dense m plane (index not applicable, too many missing entries, or it
// cannot be built right now); < 0: error (status code negated)
int miss_index_ensure(simu_b200_ctx *ctx)
{
    if (ctx->miss_state == 1) return 1;
    if (ctx->miss_state == 2) return 0;
    if (ctx->layout != SIMU_B200_LAYOUT_SAMPLE16 || !ctx->panel) return 0;
    if (getenv("SIMU_B200_TC_DENSE")) { ctx->miss_state = 2; return 0; }
    // The build costs two scans of the panel, a call with the dense plane a fraction of one: the FIRST call on a panel
    // runs dense, the index is built when the same panel is used again (several simulations over one resident panel;
    // a drop-in run that stages, simulates once and exits never pays for it).  Both forms give the same bits.
    if (++ctx->miss_uses < 2) return 0;
    // small panels are launch-bound (config 1: 7 launches of a few microseconds): one more kernel costs more than the plane
    if ((double)ctx->panel_rows * (double)ctx->n_samples < 2.5e8 && !getenv("SIMU_B200_TC_SPARSE")) { ctx->miss_state = 2; return 0; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) return 0;                      // the build synchronises: not inside a capture

    const int64_t wpg = ctx->s16_stride >> 2, ngroups = (ctx->panel_rows + 15) >> 4;       // words (samples, padded) per group
    unsigned long long *d_cnt = (unsigned long long *)ctx_scratch(ctx, SCR_MISS_CNT, (size_t)wpg * 8);
    unsigned long long *d_off = (unsigned long long *)ctx_scratch(ctx, SCR_MISS_OFF, (size_t)(wpg + 1) * 8);
    // no HBM left for the index (a panel that fills the GPU): the dense plane needs none
    if (!d_cnt || !d_off) { cudaGetLastError(); ctx->err.clear(); ctx->miss_state = 2; return 0; }
    auto fail = [&](cudaError_t e, const char *what) { return -ctx_cuda(ctx, e, what); };
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_cnt, 0, (size_t)wpg * 8, ctx->stream)) != cudaSuccess) return fail(e, "miss index memset");
    const int64_t bx = (wpg + 255) / 256;
    const dim3 grid((unsigned)bx, (unsigned)std::min<int64_t>(std::min<int64_t>(ngroups, 65535), std::max<int64_t>(1, 24 * ctx->num_sms / bx)));
    miss_count_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(ctx->panel), wpg, ngroups, d_cnt);
    ctx->launches++;
    std::vector<unsigned long long> off((size_t)wpg + 1);
    if ((e = cudaMemcpyAsync(off.data() + 1, d_cnt, (size_t)wpg * 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return fail(e, "miss index D2H");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "miss index sync");
    off[0] = 0;
    unsigned long long most = 0;
    for (int64_t i = 0; i < wpg; i++) { most = std::max(most, off[i + 1]); off[i + 1] += off[i]; }
    const unsigned long long total = off[wpg];
    if ((double)total > kMaxRate * (double)ctx->panel_rows * (double)ctx->n_samples + 1024.0) { ctx->miss_state = 2; return 0; }
    uint32_t *d_ent = (uint32_t *)ctx_scratch(ctx, SCR_MISS_ENT, (size_t)std::max<unsigned long long>(total, 1) * 4);
    if (!d_ent) { cudaGetLastError(); ctx->err.clear(); ctx->miss_state = 2; return 0; }
    // offsets, and the fill cursors (= the segment starts) in the count array
    if ((e = cudaMemcpyAsync(d_off, off.data(), (size_t)(wpg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_cnt, d_off, (size_t)wpg * 8, cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess)
        return fail(e, "miss index H2D");
    if (total) {
        miss_fill_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(ctx->panel), wpg, ngroups, d_cnt, d_ent);
        ctx->launches++;
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "miss index fill");   // `off` is a host vector
    ctx->miss_total = (int64_t)total;
    ctx->miss_chunks = (int)std::max<unsigned long long>(1, (most + kSeg - 1) / kSeg);
    ctx->miss_state = 1;
    return 1;
}

int64_t miss_corr_stride(const simu_b200_ctx *ctx) { return ctx->s16_stride >> 2; }

// true when launch_miss_correct adds to corr (which must then be zero on entry) instead of storing every sample's sum
bool miss_correct_accumulates(const simu_b200_ctx *ctx) { return ctx->miss_chunks > 1; }

// queue the correction pass: corr[k][i] = the sum of sample i's missing entries' weights, i < n_samples
int launch_miss_correct(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int traits, const long long *d_wm, unsigned long long *d_corr)
{
    const int64_t n = ctx->n_samples, stride = miss_corr_stride(ctx);
    const uint32_t *ent = (const uint32_t *)ctx->scratch[SCR_MISS_ENT];
    const unsigned long long *off = (const unsigned long long *)ctx->scratch[SCR_MISS_OFF];
    const unsigned bx = (unsigned)((n + 7) / 8);
    if (ctx->miss_chunks > 1) {
        const dim3 grid(bx, (unsigned)ctx->miss_chunks);
        if (traits == 2) miss_correct_kernel<2, true><<<grid, 256, 0, stream>>>(ent, off, n, mc, d_wm, d_corr, stride);
        else miss_correct_kernel<1, true><<<grid, 256, 0, stream>>>(ent, off, n, mc, d_wm, d_corr, stride);
    } else {
        if (traits == 2) miss_correct_kernel<2, false><<<bx, 256, 0, stream>>>(ent, off, n, mc, d_wm, d_corr, stride);
        else miss_correct_kernel<1, false><<<bx, 256, 0, stream>>>(ent, off, n, mc, d_wm, d_corr, stride);
    }
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "miss_correct_kernel launch");
}
