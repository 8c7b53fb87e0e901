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
#include <stdlib.h>

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


// ------------------------------------------------------------------------------------------------
// Second form: the words arrive through a shared-memory ring filled by TMA bulk copies.
//
// The kernel above keeps 8 x 16 bytes per lane in flight in REGISTERS; at 128 registers per thread that is 16 warps and
// 64 KB per SM, and the warps alternate between waiting for their loads and folding them (ncu: long-scoreboard-bound,
// 0.89-0.94 of the HBM roofline on config 2).  Here every warp owns a ring of kRing 8 KB slots: one lane issues
// cp.async.bulk copies two slots ahead, the warp folds a slot out of shared memory (conflict-free LDS.128) and refills
// it -- 128 KB in flight per SM whatever the register allocation, and no producer warp or cross-warp barrier.
//
// Work split: the (group, 8 KB chunk) list in group-major order is cut into one CONTIGUOUS range per warp, equal to
// within one chunk.  A warp that covers a whole group writes its counts directly; groups that straddle ranges are
// combined through global atomics and a ticket per group, and the last piece to arrive finalises (and zeroes the
// accumulators again for the next call).
constexpr int kRing = 3;
constexpr int kSlotBytes = 8192;                  // 512 vectors = 2048 samples of one group
constexpr int kSlotVecs = kSlotBytes / 16;
constexpr int kWarpsT = 8;
constexpr int kSmemT = kWarpsT * kRing * kSlotBytes + kWarpsT * kRing * 8;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// two rounds (16 vectors per lane) of one slot -> two weight-32 words per tree; vectors >= nv read as zero (FULL: the
// slot holds all 512 vectors and the loads carry no predicate -- 16 compares and 32 register clears less per slot)
template <int NLEV, bool FULL>
__device__ __forceinline__ void slot_rounds(uint32_t slot, int lane, int nv, Sliced<NLEV> &R, Sliced<NLEV> &A, uint32_t (&r32)[2], uint32_t (&a32)[2])
{
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t w[32], b[32];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int v = (8 * h + q) * 32 + lane;
            uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
            if (FULL || v < nv) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(slot + 16u * (uint32_t)v));
            w[4 * q] = x0; w[4 * q + 1] = x1; w[4 * q + 2] = x2; w[4 * q + 3] = x3;
        }
#pragma unroll
        for (int q = 0; q < 32; q++) b[q] = w[q] & (w[q] * 2u);
        r32[h] = R.template fold<0, 5, 0>(w);
        a32[h] = A.template fold<0, 5, 0>(b);
    }
}

struct S16Tma {
    const uint8_t *panel;
    int64_t s16_stride, mc, n_samples;
    int32_t *counts;
    double *freq;
    unsigned int *acc;          // [groups][48] partial sums + [groups] tickets after them; all zero between calls
    int64_t ngroups;
    int cpg;                    // chunks per group
    int64_t total;              // ngroups * cpg
    int spread;                 // range -> warp assignment (see the kernel)
};

__device__ __forceinline__ int64_t warp_of_chunk(int64_t x, int64_t warps, int64_t total) { return ((x + 1) * warps - 1) / total; }

template <int NLEV>
__global__ void __launch_bounds__(kWarpsT * 32, 1)
counts_s16_tma_kernel(const S16Tma p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // participating warps: never more than chunks, so that no range is empty (the ticket of a straddling group counts
    // the warps between its first and its last owner)
    const int64_t launched = (int64_t)gridDim.x * kWarpsT, warps = launched < p.total ? launched : p.total;
    // which contiguous range this warp takes: ranges that are neighbours in the panel go to DIFFERENT SMs (an SM's eight
    // warps read eight regions spread over the panel instead of one 17 MB run).  ncu, config 2, active cycles per SM
    // min / avg / max: 599K / 665K / 755K with neighbouring ranges on one SM, 610K / 664K / 721K spread (the kernel ends
    // with its slowest warp: 397 -> 383 us, config 5 1.94 -> 1.87 ms, config 3 58.6 -> 55.0 us on the same box).  The rest
    // of the spread is per-warp; a dynamic tail that balances it was measured and does not shorten the kernel.
    const int64_t me = p.spread ? (int64_t)warp * gridDim.x + blockIdx.x : (int64_t)blockIdx.x * kWarpsT + warp;
    const int64_t c_begin = me < warps ? (me * p.total) / warps : 0, c_end = me < warps ? ((me + 1) * p.total) / warps : 0;
    const uint32_t ring = smem_addr(smem) + (uint32_t)(warp * kRing * kSlotBytes);
    const uint32_t bars = smem_addr(smem) + (uint32_t)(kWarpsT * kRing * kSlotBytes + warp * kRing * 8);
    if (lane == 0) {
        for (int s = 0; s < kRing; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (c_begin >= c_end) return;
    const int64_t row_bytes = p.s16_stride;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    auto chunk_bytes = [&](int c) { const int64_t left = row_bytes - (int64_t)c * kSlotBytes; return (uint32_t)(left < kSlotBytes ? left : kSlotBytes); };
    // producer cursor (lane 0 issues): chunk id -> (group, chunk in group)
    int64_t pid = c_begin;
    int64_t pg = c_begin / p.cpg;
    int pc = (int)(c_begin - pg * p.cpg), pslot = 0;
    auto issue = [&]() {
        if (pid < c_end) {
            if (lane == 0) {
                const uint32_t bytes = chunk_bytes(pc), bar = bars + 8u * pslot;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(ring + (uint32_t)(pslot * kSlotBytes)), "l"(p.panel + pg * row_bytes + (int64_t)pc * kSlotBytes), "r"(bytes), "r"(bar), "l"(policy)
                             : "memory");
            }
            pid++;
            if (++pc == p.cpg) { pc = 0; pg++; }
            if (++pslot == kRing) pslot = 0;
        }
    };
#pragma unroll
    for (int s = 0; s < kRing; s++) issue();

    int nlev = 1;
    while (nlev < NLEV && (p.n_samples >> nlev) != 0) nlev++;          // counts <= N < 2^nlev
    int64_t g = c_begin / p.cpg;
    int c = (int)(c_begin - g * p.cpg), slot = 0;
    uint32_t phase = 0;
    int64_t id = c_begin;
    while (id < c_end) {
        // one piece: the chunks [c, c_last] of group g that lie in this warp's range
        const int c_stop = (int)min((int64_t)p.cpg, (int64_t)c + (c_end - id));       // exclusive
        Sliced<NLEV> R, A;
        R.clear(); A.clear();
        auto wait_slot = [&]() {
            const uint32_t bar = bars + 8u * slot;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        };
        auto next_slot = [&]() {
            __syncwarp();                       // every lane has its words in registers: the slot may be overwritten
            issue();
            if (++slot == kRing) { slot = 0; phase ^= 1u; }
            id++; c++;
        };
        // One slot per iteration, NOT unrolled: the two rounds' weight-32 words meet in one carry-save adder at level 5
        // and the carry is ripple-added from level 6 up (0.7 ALU instructions per word more than folding 8 rounds
        // through levels 5-7 first, which unrolled to 40 KB of code: with 8 warps per SM the instruction fetch then
        // showed as the top stall, `no_instruction` 2.5 per issue).
#pragma unroll 1
        while (c < c_stop) {
            uint32_t r2[2], a2[2], cr, ca;
            const int nv = (int)(chunk_bytes(c) >> 4);
            wait_slot();
            if (nv == kSlotVecs) slot_rounds<NLEV, true>(ring + (uint32_t)(slot * kSlotBytes), lane, nv, R, A, r2, a2);
            else slot_rounds<NLEV, false>(ring + (uint32_t)(slot * kSlotBytes), lane, nv, R, A, r2, a2);
            next_slot();
            csa(R.lvl[5], cr, R.lvl[5], r2[0], r2[1]);
            csa(A.lvl[5], ca, A.lvl[5], a2[0], a2[1]);
            // all NLEV levels, unconditionally: skipping the levels N cannot reach costs a select per level and tree
            // (61 SEL + 25 ISETP per slot against the 12 LOP3 it saved for N = 100K)
            R.template ripple<6>(cr, NLEV);
            A.template ripple<6>(ca, NLEV);
        }
        // the piece's counts: lanes' counters added, lane p extracts bit position p
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { R.add_lane(off, nlev); A.add_lane(off, nlev); }
        const uint32_t mine = R.extract(lane, nlev);            // even lane 2f: low-bit count, odd 2f+1: high-bit count
        uint32_t both = A.extract(lane, nlev);                  // odd lane 2f+1: both-bits count
        uint32_t lo = __shfl_sync(0xffffffffu, mine, lane & ~1), hi = mine;        // valid on odd lanes
        // pieces of this group = warps whose ranges touch it
        const int64_t first_w = warp_of_chunk(g * p.cpg, warps, p.total), last_w = warp_of_chunk((g + 1) * p.cpg - 1, warps, p.total);
        bool finalise = true;
        if (first_w != last_w) {
            unsigned int *acc = p.acc + g * 48, *ticket = p.acc + p.ngroups * 48 + g;
            if (lane & 1) {
                atomicAdd(acc + 3 * (lane >> 1), lo);
                atomicAdd(acc + 3 * (lane >> 1) + 1, hi);
                atomicAdd(acc + 3 * (lane >> 1) + 2, both);
            }
            __threadfence();
            __syncwarp();
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            finalise = t == (unsigned int)(last_w - first_w);
            if (finalise) {
                __threadfence();
                if (lane & 1) {
                    lo = atomicExch(acc + 3 * (lane >> 1), 0u);
                    hi = atomicExch(acc + 3 * (lane >> 1) + 1, 0u);
                    both = atomicExch(acc + 3 * (lane >> 1) + 2, 0u);
                }
                if (lane == 0) *ticket = 0u;
            }
        }
        if (finalise && (lane & 1)) {
            const int64_t j = 16 * g + (lane >> 1);
            if (j < p.mc) {
                const int n3 = (int)both, n1 = (int)(lo - both), n2 = (int)(hi - both);
                const int n0 = (int)p.n_samples - n1 - n2 - n3;
                if (p.counts) reinterpret_cast<int4 *>(p.counts)[j] = make_int4(n0, n1, n2, n3);
                if (p.freq) {
                    const int nm = n2 + n1 + n0;               // source.cc:874
                    p.freq[j] = (nm == 0) ? 0.0                // source.cc:875
                                          : static_cast<double>(2 * n0 + 1 * n1) / static_cast<double>(2 * nm);
                }
            }
        }
        if (c == p.cpg) { c = 0; g++; }
    }
}

}  // namespace

static int launch_counts_s16_tma(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int32_t *d_counts, double *d_freq)
{
    S16Tma p;
    p.panel = ctx->panel; p.s16_stride = ctx->s16_stride; p.mc = mc; p.n_samples = ctx->n_samples; p.counts = d_counts; p.freq = d_freq;
    p.ngroups = (mc + 15) >> 4;
    p.cpg = (int)((ctx->s16_stride + kSlotBytes - 1) / kSlotBytes);
    p.total = p.ngroups * p.cpg;
    { const char *e = getenv("SIMU_B200_COUNTS_SPREAD"); p.spread = e ? atoi(e) : 1; }
    // partial sums of the groups that straddle two warps' ranges + their tickets: zero between calls (the finalising
    // warp clears what it reads), so only a new or larger buffer is cleared here
    const size_t acc_bytes = (size_t)p.ngroups * 49 * sizeof(unsigned int);
    const bool fresh = ctx->scratch_bytes[SCR_CNTACC] < acc_bytes;
    p.acc = (unsigned int *)ctx_scratch(ctx, SCR_CNTACC, acc_bytes);
    if (!p.acc) return SIMU_B200_ENOMEM;
    if (fresh) CUDA_TRY(ctx, cudaMemsetAsync(p.acc, 0, ctx->scratch_bytes[SCR_CNTACC], stream));
    if (!ctx->cnt_attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(counts_s16_tma_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT));
        CUDA_TRY(ctx, cudaFuncSetAttribute(counts_s16_tma_kernel<kMaxLevels>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT));
        ctx->cnt_attr_set = true;
    }
    int64_t blocks = (p.total + kWarpsT - 1) / kWarpsT;
    if (blocks > ctx->num_sms) blocks = ctx->num_sms;
    if (ctx->n_samples < (1 << 20)) counts_s16_tma_kernel<20><<<(unsigned)blocks, kWarpsT * 32, kSmemT, stream>>>(p);
    else counts_s16_tma_kernel<kMaxLevels><<<(unsigned)blocks, kWarpsT * 32, kSmemT, stream>>>(p);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "counts_s16_tma_kernel launch");
}

int launch_counts_s16(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int32_t *d_counts, double *d_freq)
{
    if (mc <= 0) return SIMU_B200_OK;
    const int64_t groups = (mc + 15) >> 4;
    const int64_t max_blocks = (int64_t)ctx->num_sms * 2;            // 2 CTAs of 8 warps per SM
    // Which kernel: with at least 8 whole groups per resident warp the register-fed kernel needs no sample split and
    // no cross-warp combine and is the faster one (config 4: 4.95 against 5.12 ms); below that the TMA-fed ring with
    // its balanced contiguous ranges wins (config 2: 0.375 against 0.428 ms, config 3: 0.055 against 0.067 ms).
    // SIMU_B200_COUNTS=reg|tma forces one of them (A/B runs).
    const char *force = getenv("SIMU_B200_COUNTS");
    // ... and a panel of fewer chunks than two per warp of the ring kernel is launch-bound either way; the register-fed kernel
    // has the shorter prologue there (config 1: 13.5 against 16 us)
    const int64_t chunks = groups * ((ctx->s16_stride + kSlotBytes - 1) / kSlotBytes);
    const bool tma = force ? force[0] == 't' : (groups < 8 * max_blocks * kWarps && chunks >= 2 * (int64_t)ctx->num_sms * kWarpsT);
    if (tma) return launch_counts_s16_tma(ctx, stream, mc, d_counts, d_freq);
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
