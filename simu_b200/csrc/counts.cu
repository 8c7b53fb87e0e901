// Kernel (a): per-SNP allele counts and A1 frequency.
//
// Replaces find_freq's inner loop (/root/reference/src/source.cc:869-875) and
// the bed_read_row + unpack_snps it sits on (libplinkio/src/bed.c:91-114,
// :321-358).  One warp owns one causal SNP row and streams it with coalesced
// 16-byte vector loads.  Nothing is unpacked: with lo = even bits, hi = odd
// bits of the packed row,
//     n3 (missing, bits 01) = popc(lo) - popc(lo&hi)
//     n1 (het,     bits 10) = popc(hi) - popc(lo&hi)
//     n2 (hom A2,  bits 11) = popc(lo&hi)
//     n0 (hom A1,  bits 00) = N - n1 - n2 - n3
// so row padding (which decodes as 00) never has to be counted.  POPC issues
// at 16 lanes/clk/SM, so the raw words go through a Harley-Seal carry-save
// tree (full-rate LOP3) and only the "fours" word is popcounted: 4 POPC per
// 64 genotypes instead of 12.  Lane partials are combined with
// __reduce_add_sync.
//
// Algorithmic bytes per SNP: ceil(N/4) read + 16 (counts) + 8 (freq) written.
#include "counts_g4.cuh"
#include "ctx.cuh"

namespace {

constexpr uint32_t kLoMask = 0x55555555u;

using g4::ldg_stream;
using g4::word_mask;

struct LaneCounts {
    uint32_t ones = 0, twos = 0;      // carry-save state of the raw words
    uint32_t four_all = 0, four_lo = 0;  // popcounts of the weight-4 words
    uint32_t both = 0;                // popc(lo & hi)

    __device__ __forceinline__ void add(const uint4 v)
    {
        // carry-save adders: (a,b,c) -> sum a^b^c (weight 1), maj(a,b,c) (weight 2)
        uint32_t t1 = (ones & v.x) | (ones & v.y) | (v.x & v.y);
        ones = ones ^ v.x ^ v.y;
        uint32_t t2 = (ones & v.z) | (ones & v.w) | (v.z & v.w);
        ones = ones ^ v.z ^ v.w;
        uint32_t f = (twos & t1) | (twos & t2) | (t1 & t2);
        twos = twos ^ t1 ^ t2;
        four_all += __popc(f);
        four_lo += __popc(f & kLoMask);
        // lo&hi plane of two words packed into one (even bits: x, odd bits: y)
        uint32_t p1 = (v.x & (v.x >> 1) & kLoMask) | (v.y & (v.y << 1) & ~kLoMask);
        uint32_t p2 = (v.z & (v.z >> 1) & kLoMask) | (v.w & (v.w << 1) & ~kLoMask);
        both += __popc(p1) + __popc(p2);
    }
};

__global__ void __launch_bounds__(256)
counts_kernel(const uint8_t *__restrict__ panel, int64_t row_stride, const int32_t *__restrict__ rows,
              int64_t mc, int64_t n_samples, int32_t *__restrict__ counts, double *__restrict__ freq)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row_bytes = (n_samples + 3) >> 2;
    const int64_t nvec = (row_bytes + 15) >> 4;           // 16-byte vectors holding data
    const int64_t nfull = nvec - 1;                       // all but the last need no mask
    const int64_t tail_bits = 2 * n_samples - 128 * nfull;  // valid bits in the last vector

    for (int64_t j = warp_id; j < mc; j += warps_total) {
        const int64_t r = rows ? (int64_t)rows[j] : j;
        const uint4 *p = reinterpret_cast<const uint4 *>(panel + r * row_stride);
        LaneCounts c;
        int64_t i = lane;
        for (; i + 96 < nfull; i += 128) {
            uint4 v0 = ldg_stream(p + i), v1 = ldg_stream(p + i + 32);
            uint4 v2 = ldg_stream(p + i + 64), v3 = ldg_stream(p + i + 96);
            c.add(v0); c.add(v1); c.add(v2); c.add(v3);
        }
        for (; i < nfull; i += 32) c.add(ldg_stream(p + i));
        if (lane == (int)(nfull & 31)) {                  // masked last vector
            uint4 v = ldg_stream(p + nfull);
            v.x &= word_mask(tail_bits, 0); v.y &= word_mask(tail_bits, 1);
            v.z &= word_mask(tail_bits, 2); v.w &= word_mask(tail_bits, 3);
            c.add(v);
        }
        uint32_t all = 4u * c.four_all + 2u * __popc(c.twos) + __popc(c.ones);
        uint32_t lo = 4u * c.four_lo + 2u * __popc(c.twos & kLoMask) + __popc(c.ones & kLoMask);
        all = __reduce_add_sync(0xFFFFFFFFu, all);
        lo = __reduce_add_sync(0xFFFFFFFFu, lo);
        uint32_t both = __reduce_add_sync(0xFFFFFFFFu, c.both);
        if (lane == 0) {
            const int n2 = (int)both;
            const int n3 = (int)(lo - both);
            const int n1 = (int)(all - lo - both);
            const int n0 = (int)n_samples - n1 - n2 - n3;
            if (counts) reinterpret_cast<int4 *>(counts)[j] = make_int4(n0, n1, n2, n3);
            if (freq) {
                const int nm = n2 + n1 + n0;               // source.cc:874
                freq[j] = (nm == 0) ? 0.0                  // source.cc:875
                                    : static_cast<double>(2 * n0 + 1 * n1) / static_cast<double>(2 * nm);
            }
        }
    }
}

// ------------------------------------------------------------ GROUP4 layout
//
// One warp per 4-row group (counts_g4.cuh has the carry-save counting itself); SPLIT warps of a
// CTA share one group when there are fewer groups than resident warps.

template <int SPLIT>
__global__ void __launch_bounds__(256)
counts_g4_kernel(const uint8_t *__restrict__ panel, int64_t group_stride, int64_t mc, int64_t n_samples,
                 int32_t *__restrict__ counts, double *__restrict__ freq)
{
    __shared__ uint32_t part[8][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kGroupsPerCta = (int)(blockDim.x >> 5) / SPLIT;   // 8 warps per CTA (4 when launched beside a lookup CTA)
    constexpr int kStride = 32 * SPLIT;                    // vectors between two loads of a lane
    const int sub = warp % SPLIT, gslot = warp / SPLIT;
    const int64_t ngroups = (mc + 3) >> 2;
    const int nvec = (int)((n_samples + 15) >> 4);         // 16-byte vectors holding samples
    const int nfull = nvec - 1;
    const int64_t tail_bits = 8 * (n_samples - 16 * (int64_t)nfull);   // valid bits in the last vector
    const int64_t iters = (ngroups + (int64_t)gridDim.x * kGroupsPerCta - 1) / ((int64_t)gridDim.x * kGroupsPerCta);

    for (int64_t it = 0; it < iters; it++) {
        const int64_t g = (it * gridDim.x + blockIdx.x) * kGroupsPerCta + gslot;
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, both[4] = {0, 0, 0, 0};
        if (g < ngroups)
            g4::count_vectors(reinterpret_cast<const uint4 *>(panel + g * group_stride), sub * 32 + lane, kStride, nvec,
                              nfull, tail_bits, lo, hi, both);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            lo[r] = __reduce_add_sync(0xFFFFFFFFu, lo[r]);
            hi[r] = __reduce_add_sync(0xFFFFFFFFu, hi[r]);
            both[r] = __reduce_add_sync(0xFFFFFFFFu, both[r]);
        }
        if (SPLIT > 1) {
            __syncthreads();
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < 4; r++) { part[warp][r] = lo[r]; part[warp][4 + r] = hi[r]; part[warp][8 + r] = both[r]; }
            }
            __syncthreads();
        }
        if (sub == 0 && lane < 4 && g < ngroups) {
            const int r = lane;
            uint32_t l = 0, h = 0, b = 0;
            if (SPLIT > 1) {
                for (int q = 0; q < SPLIT; q++) { l += part[warp + q][r]; h += part[warp + q][4 + r]; b += part[warp + q][8 + r]; }
            } else {
                l = r == 0 ? lo[0] : r == 1 ? lo[1] : r == 2 ? lo[2] : lo[3];
                h = r == 0 ? hi[0] : r == 1 ? hi[1] : r == 2 ? hi[2] : hi[3];
                b = r == 0 ? both[0] : r == 1 ? both[1] : r == 2 ? both[2] : both[3];
            }
            const int64_t j = 4 * g + r;
            if (j < mc) {
                const int n2 = (int)b, n3 = (int)(l - b), n1 = (int)(h - b);
                const int n0 = (int)n_samples - n1 - n2 - n3;
                if (counts) reinterpret_cast<int4 *>(counts)[j] = make_int4(n0, n1, n2, n3);
                if (freq) {
                    const int nm = n2 + n1 + n0;               // source.cc:874
                    freq[j] = (nm == 0) ? 0.0                  // source.cc:875
                                        : static_cast<double>(2 * n0 + 1 * n1) / static_cast<double>(2 * nm);
                }
            }
        }
    }
}

}  // namespace

// GROUP4 counts of mc rows that start at `first_group` (rows 4*first_group ..), on an explicit stream
int launch_counts_g4_range(simu_b200_ctx *ctx, cudaStream_t stream, int64_t first_group, int64_t mc, int32_t *d_counts,
                           double *d_freq, int threads)
{
    if (mc <= 0) return SIMU_B200_OK;
    const int warps_per_block = threads / 32;               // 8, or 4 to squeeze in beside a resident lookup CTA
    const int64_t max_blocks = (int64_t)ctx->num_sms * 8;   // 64 resident warps per SM
    // one warp per 4-row group; when the groups do not fill the resident warps, 2/4/8 warps
    // of a CTA share a group (more loads in flight per group row)
    const int64_t groups = (mc + 3) >> 2;
    const int64_t resident = max_blocks * 8;
    int split = 1;
    while (split < warps_per_block && groups * split * 2 <= resident) split *= 2;
    int64_t blocks = (groups * split + warps_per_block - 1) / warps_per_block;
    if (blocks > max_blocks) blocks = max_blocks;
    const uint8_t *base = ctx->panel + first_group * ctx->group_stride;
#define LAUNCH_G4(S) counts_g4_kernel<S><<<(unsigned)blocks, warps_per_block * 32, 0, stream>>>( \
        base, ctx->group_stride, mc, ctx->n_samples, d_counts, d_freq)
    if (split == 1) LAUNCH_G4(1); else if (split == 2) LAUNCH_G4(2); else if (split == 4) LAUNCH_G4(4); else LAUNCH_G4(8);
#undef LAUNCH_G4
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "counts_g4_kernel launch");
}

int launch_counts(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, int32_t *d_counts, double *d_freq)
{
    if (mc <= 0) return SIMU_B200_OK;
    const int ps = prof_begin(ctx, PROF_COUNTS);
    int rc = SIMU_B200_OK;
    if (ctx->layout == SIMU_B200_LAYOUT_SAMPLE16) {
        rc = launch_counts_s16(ctx, ctx->stream, mc, d_counts, d_freq);
    } else if (ctx->layout == SIMU_B200_LAYOUT_GROUP4) {
        rc = launch_counts_g4_range(ctx, ctx->stream, 0, mc, d_counts, d_freq, 256);
    } else {
        const int warps_per_block = 8;
        const int64_t max_blocks = (int64_t)ctx->num_sms * 8;   // 64 resident warps per SM
        int64_t blocks = (mc + warps_per_block - 1) / warps_per_block;
        if (blocks > max_blocks) blocks = max_blocks;
        counts_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, ctx->stream>>>(
            ctx->panel, ctx->row_stride, d_rows, mc, ctx->n_samples, d_counts, d_freq);
        ctx->launches++;
        rc = ctx_cuda(ctx, cudaGetLastError(), "counts_kernel launch");
    }
    prof_end(ctx, ps);
    return rc;
}
