// Kernel (a) on SAMPLE16 panels: per-SNP allele counts + frequency (src/source.cc:843-885).
//
// In a SAMPLE16 panel a 32-bit word holds ONE sample's genotypes for 16 consecutive SNP rows
// (dosage coding: 00 hom A1, 01 het, 10 hom A2, 11 missing), so the counts of a SNP are sums over
// WORDS of one bit position -- popcount does not apply.  The kernel counts "vertically": the words
// of a group go through carry-save adder trees whose state is a BIT-SLICED counter (level l holds
// bit l of the running count of every bit position), at the tree's inherent 2 LOP3 per word:
//
//   R tree  input w               -> per SNP f: #(low bit set) at position 2f, #(high bit set) at 2f+1
//   A tree  input w & (w << 1)    -> #(both bits set = missing) at position 2f+1 (the shift is an
//                                    IMAD on the FMA pipe; even positions carry garbage that no
//                                    adder ever moves to another position)
//
//   n3 = both, n1 = low - both, n2 = high - both, n0 = N - n1 - n2 - n3 (padding samples are 0 = never
//   counted), freq with the reference's own double expression (source.cc:874-875) -> bit-exact.
//
// Schedule per lane: a round = 8 x 16-byte loads (32 words, all in flight together) folded through
// levels 0-4 into one weight-32 word per tree; 8 rounds fold through levels 5-7 into a weight-256 word
// that is ripple-added into levels 8+ (2 ops per level per 256 words).  ~5 ALU-pipe instructions per
// word = 0.31 per genotype, the same budget as the GROUP4 popcount kernel.  At the end the lanes'
// counters are added with a shuffle butterfly of bit-sliced full adders and lane p extracts position p.
//
// One warp per 16-row group; 2/4/8 warps of a CTA share a group (sample ranges) when there are fewer
// groups than resident warps.  Algorithmic bytes: ceil(N/4) + 24 per SNP.
#include "ctx.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxLevels = 29;     // counts up to 2^29 - 1 >= the largest N simu_b200_open accepts

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"     // not volatile: the scheduler may hoist the next round's loads
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void csa(uint32_t &sum, uint32_t &carry, uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t s, m;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(m) : "r"(a), "r"(b), "r"(c));   // majority
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(c));   // parity
    sum = s; carry = m;
}

template <int kLevels>
struct Sliced {
    uint32_t lvl[kLevels];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int l = 0; l < kLevels; l++) lvl[l] = 0;
    }
    // fold the 2^K words w[OFF .. OFF + 2^K) through levels BASE .. BASE+K-1 (depth first: a half's carry
    // is formed before the other half is touched); returns the carry of weight 2^(BASE+K)
    template <int BASE, int K, int OFF = 0, int NW>
    __device__ __forceinline__ uint32_t fold(const uint32_t (&w)[NW])
    {
        uint32_t c;
        if constexpr (K == 1) {
            csa(lvl[BASE], c, lvl[BASE], w[OFF], w[OFF + 1]);
        } else {
            const uint32_t a = fold<BASE, K - 1, OFF>(w);
            const uint32_t b = fold<BASE, K - 1, OFF + (1 << (K - 1))>(w);
            csa(lvl[BASE + K - 1], c, lvl[BASE + K - 1], a, b);
        }
        return c;
    }
    // add a word of weight 2^FROM (binary ripple; nlev = levels that can be non-zero for this N)
    template <int FROM>
    __device__ __forceinline__ void ripple(uint32_t x, int nlev)
    {
#pragma unroll
        for (int l = FROM; l < kLevels; l++) {
            if (l < nlev) {
                const uint32_t t = lvl[l] & x;
                lvl[l] ^= x;
                x = t;
            }
        }
    }
    // this += the same counter of lane (lane ^ off)
    __device__ __forceinline__ void add_lane(int off, int nlev)
    {
        uint32_t carry = 0;
#pragma unroll
        for (int l = 0; l < kLevels; l++) {
            if (l < nlev) {           // warp-uniform
                const uint32_t o = __shfl_xor_sync(0xffffffffu, lvl[l], off);
                csa(lvl[l], carry, lvl[l], o, carry);
            }
        }
    }
    __device__ __forceinline__ uint32_t extract(int pos, int nlev) const
    {
        uint32_t c = 0;
#pragma unroll
        for (int l = 0; l < kLevels; l++)
            if (l < nlev) c |= ((lvl[l] >> pos) & 1u) << l;
        return c;
    }
};

// one round: vectors i0 + q*stride, q < 8 (zero beyond nvec when CHECKED) -> weight-32 words of both trees
template <bool CHECKED, int NLEV>
__device__ __forceinline__ void round32(const uint4 *__restrict__ p, int64_t i0, int64_t stride, int64_t nvec, Sliced<NLEV> &R,
                                        Sliced<NLEV> &A, uint32_t &r32, uint32_t &a32)
{
    uint4 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int64_t i = i0 + q * stride;
        if (!CHECKED || i < nvec) v[q] = ldg_stream(p + i);
        else v[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t w[32], b[32];
#pragma unroll
    for (int q = 0; q < 8; q++) { w[4 * q] = v[q].x; w[4 * q + 1] = v[q].y; w[4 * q + 2] = v[q].z; w[4 * q + 3] = v[q].w; }
#pragma unroll
    for (int q = 0; q < 32; q++) b[q] = w[q] & (w[q] * 2u);
    r32 = R.template fold<0, 5, 0>(w);
    a32 = A.template fold<0, 5, 0>(b);
}

// NLEV: levels of the bit-sliced counters = bits of the largest count, N < 2^NLEV (20 covers a million samples and
// leaves the registers the scheduler needs to keep the next round's loads in flight)
template <int SPLIT, int NLEV>
__global__ void __launch_bounds__(kWarps * 32, 2)
counts_s16_kernel(const uint8_t *__restrict__ panel, int64_t s16_stride, int64_t mc, int64_t n_samples,
                  int32_t *__restrict__ counts, double *__restrict__ freq)
{
    __shared__ uint32_t part[kWarps / SPLIT][16][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = warp / SPLIT, sub = warp % SPLIT;               // group slot in the CTA, sample split
    const int64_t ngroups = (mc + 15) >> 4;
    const int64_t nvec = (n_samples + 3) >> 2;                     // 16-byte vectors (4 samples) with data
    const int64_t stride = 32 * SPLIT;
    int nlev = 1;
    while (nlev < NLEV && (n_samples >> nlev) != 0) nlev++;        // counts <= N < 2^nlev
    constexpr int kPerCta = kWarps / SPLIT;

    for (int64_t gb = (int64_t)blockIdx.x * kPerCta; gb < ngroups; gb += (int64_t)gridDim.x * kPerCta) {
        const int64_t g = gb + gl;
        if (SPLIT > 1) {
            for (int i = threadIdx.x; i < kPerCta * 48; i += kWarps * 32) (&part[0][0][0])[i] = 0;
            __syncthreads();
        }
        uint32_t lo = 0, hi = 0, both = 0;
        if (g < ngroups) {
            const uint4 *p = reinterpret_cast<const uint4 *>(panel + g * s16_stride);
            Sliced<NLEV> R, A;
            R.clear(); A.clear();
            int64_t i0 = sub * 32 + lane;
            // super-rounds of 8 rounds = 256 words per lane
            for (; i0 + 63 * stride < nvec; i0 += 64 * stride) {
                uint32_t r32[8], a32[8];
#pragma unroll
                for (int r = 0; r < 8; r++) round32<false, NLEV>(p, i0 + 8 * r * stride, stride, nvec, R, A, r32[r], a32[r]);
                R.template ripple<8>(R.template fold<5, 3, 0>(r32), nlev);
                A.template ripple<8>(A.template fold<5, 3, 0>(a32), nlev);
            }
            for (; i0 < nvec; i0 += 8 * stride) {
                uint32_t r32, a32;
                if (i0 + 7 * stride < nvec) round32<false, NLEV>(p, i0, stride, nvec, R, A, r32, a32);
                else round32<true, NLEV>(p, i0, stride, nvec, R, A, r32, a32);
                R.template ripple<5>(r32, nlev);
                A.template ripple<5>(a32, nlev);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { R.add_lane(off, nlev); A.add_lane(off, nlev); }
            const uint32_t mine = R.extract(lane, nlev);            // even lane 2f: low-bit count, odd 2f+1: high-bit count
            const uint32_t bth = A.extract(lane, nlev);             // odd lane 2f+1: both-bits count
            const uint32_t low = __shfl_sync(0xffffffffu, mine, lane & ~1);
            lo = low; hi = mine; both = bth;                        // valid on odd lanes
            if (SPLIT > 1 && (lane & 1)) {
                atomicAdd(&part[gl][lane >> 1][0], lo);
                atomicAdd(&part[gl][lane >> 1][1], hi);
                atomicAdd(&part[gl][lane >> 1][2], both);
            }
        }
        if (SPLIT > 1) {
            __syncthreads();
            if (g < ngroups && (lane & 1)) { lo = part[gl][lane >> 1][0]; hi = part[gl][lane >> 1][1]; both = part[gl][lane >> 1][2]; }
            __syncthreads();
        }
        if (g < ngroups && sub == 0 && (lane & 1)) {
            const int64_t j = 16 * g + (lane >> 1);
            if (j < mc) {
                const int n3 = (int)both, n1 = (int)(lo - both), n2 = (int)(hi - both);
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

int launch_counts_s16(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int32_t *d_counts, double *d_freq)
{
    if (mc <= 0) return SIMU_B200_OK;
    const int64_t groups = (mc + 15) >> 4;
    const int64_t max_blocks = (int64_t)ctx->num_sms * 2;            // 2 CTAs of 8 warps per SM
    const int64_t resident = max_blocks * kWarps;
    // static grid-stride assignment: aim at >= 8 (group, sample range) tasks per resident warp so that the last,
    // partly filled round costs a few per cent, not a third (6250 groups on 2368 warps: 2.6 rounds)
    int split = 1;
    while (split < kWarps && groups * split < 8 * resident) split *= 2;
    int64_t blocks = (groups * split + kWarps - 1) / kWarps;
    if (blocks > max_blocks) blocks = max_blocks;
#define LAUNCH_S16(S, L) counts_s16_kernel<S, L><<<(unsigned)blocks, kWarps * 32, 0, stream>>>( \
        ctx->panel, ctx->s16_stride, mc, ctx->n_samples, d_counts, d_freq)
#define LAUNCH_S16_L(L) do { if (split == 1) LAUNCH_S16(1, L); else if (split == 2) LAUNCH_S16(2, L); else if (split == 4) LAUNCH_S16(4, L); \
                             else LAUNCH_S16(8, L); } while (0)
    if (ctx->n_samples < (1 << 20)) LAUNCH_S16_L(20); else LAUNCH_S16_L(kMaxLevels);
#undef LAUNCH_S16_L
#undef LAUNCH_S16
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "counts_s16_kernel launch");
}
