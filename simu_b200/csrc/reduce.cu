// Kernel group (c): the scalar reductions that follow find_pheno.
//
//   sumsq            sum_i y_i^2                    source.cc:1009-1012
//   heritability     gen/env coefficients + y = y*gen + noise*env
//                                                   source.cc:1018-1027
//   argsort_f64      the sorted order that apply_liability_threshold takes
//                    with std::sort                 source.cc:1032-1034
//
// All three are O(N) / O(N log N) on N <= 500K doubles: latency bound, a few
// launches each.  They are deterministic (fixed reduction tree, stable LSD
// radix sort) so repeated runs give identical bits.
#include <cooperative_groups.h>
#include <float.h>

#include <algorithm>

#include "ctx.cuh"

namespace cg = cooperative_groups;

namespace {

// ------------------------------------------------------------------ sumsq

constexpr int kRedThreads = 256;
constexpr int kRedItems = 16;   // doubles per thread per block

__device__ __forceinline__ double block_sum(double v, double *smem)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    }
    return v;   // valid in thread 0
}

__global__ void __launch_bounds__(kRedThreads)
sumsq_partial_kernel(const double *__restrict__ y, int64_t n, double *__restrict__ partial)
{
    __shared__ double smem[32];
    const int64_t base = (int64_t)blockIdx.x * kRedThreads * kRedItems;
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < kRedItems; u++) {
        const int64_t i = base + (int64_t)u * kRedThreads + threadIdx.x;
        if (i < n) { const double v = y[i]; s += v * v; }
    }
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kRedThreads)
sum_final_kernel(const double *__restrict__ partial, int64_t np, double *__restrict__ out)
{
    __shared__ double smem[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < np; i += kRedThreads) s += partial[i];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) *out = s;
}

// ----------------------------------------------------------- heritability

__global__ void __launch_bounds__(256)
heritability_kernel(float hsq, double *__restrict__ y, int64_t n, const double *__restrict__ noise,
                    const double *__restrict__ sumsq, double *__restrict__ coef)
{
    double pheno_var = *sumsq / static_cast<double>(n);                       // source.cc:1012
    double gen_coef, env_coef;
    if ((fabs((double)hsq) <= (double)FLT_EPSILON) || (fabs(pheno_var) <= (double)FLT_EPSILON)) {
        gen_coef = 0; env_coef = 1;                                           // :1019-1021
    } else {
        gen_coef = sqrt(hsq / pheno_var);                                     // :1023
        env_coef = sqrt(1.0 - hsq);                                           // :1024
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && coef) { coef[0] = gen_coef; coef[1] = env_coef; }
    if (i < n)                                                                // :1027, no FMA contraction
        y[i] = __dadd_rn(__dmul_rn(y[i], gen_coef), __dmul_rn(noise[i], env_coef));
}

// apply_heritability in ONE cooperative launch: per-block sums of squares, a grid barrier, then every block adds the
// partials in the same fixed order (all blocks hold the same variance without another round trip), forms the coefficients
// exactly as source.cc:1012-1025 and scales its share.  Replaces sumsq_partial + sum_final + heritability_kernel
// (11.5 us per trait in the launch list of a 0.81 ms step; the sparse configs are a few launches of microseconds each).
constexpr int kHerThreads = 256;
__global__ void __launch_bounds__(kHerThreads)
heritability_fused_kernel(float hsq, double *__restrict__ y, int64_t n, const double *__restrict__ noise,
                          double *__restrict__ partial /* [grid] */, unsigned int *__restrict__ bar, double *__restrict__ coef)
{
    __shared__ double smem[32];
    __shared__ double total;
    const int G = (int)gridDim.x;                                  // <= kHerThreads (host)
    // a block owns a CONTIGUOUS range, a thread keeps its values in registers across the barrier (up to kHerKeep of them)
    constexpr int kHerKeep = 4;
    const int64_t per = (n + G - 1) / G, lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    double keep[kHerKeep];
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < kHerKeep; u++) {
        const int64_t i = lo + (int64_t)u * kHerThreads + threadIdx.x;
        keep[u] = i < hi ? y[i] : 0.0;
        s += keep[u] * keep[u];
    }
    for (int64_t i = lo + (int64_t)kHerKeep * kHerThreads + threadIdx.x; i < hi; i += kHerThreads) { const double v = y[i]; s += v * v; }
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    grid_barrier_once(bar);
    s = (int)threadIdx.x < G ? __ldcg(partial + threadIdx.x) : 0.0;
    s = block_sum(s, smem);
    if (threadIdx.x == 0) total = s;
    __syncthreads();
    const double pheno_var = total / static_cast<double>(n);                  // source.cc:1012
    double gen_coef, env_coef;
    if ((fabs((double)hsq) <= (double)FLT_EPSILON) || (fabs(pheno_var) <= (double)FLT_EPSILON)) {
        gen_coef = 0; env_coef = 1;                                           // :1019-1021
    } else {
        gen_coef = sqrt(hsq / pheno_var);                                     // :1023
        env_coef = sqrt(1.0 - hsq);                                           // :1024
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && coef) { coef[0] = gen_coef; coef[1] = env_coef; }
#pragma unroll
    for (int u = 0; u < kHerKeep; u++) {
        const int64_t i = lo + (int64_t)u * kHerThreads + threadIdx.x;
        if (i < hi) y[i] = __dadd_rn(__dmul_rn(keep[u], gen_coef), __dmul_rn(noise[i], env_coef));      // :1027, no FMA contraction
    }
    for (int64_t i = lo + (int64_t)kHerKeep * kHerThreads + threadIdx.x; i < hi; i += kHerThreads)
        y[i] = __dadd_rn(__dmul_rn(y[i], gen_coef), __dmul_rn(noise[i], env_coef));
}

// ------------------------------------------------------- effect-size scaling

// Per causal variant (source.cc:1245-1262, :1272-1286, :1089-1094):
//   het = fabs(2 f (1 - f));   het < FLT_EPSILON -> *bad = min(index)   (the host throws)
//   op 0:  x *= sqrt(pow(het, S_k[component]))      --gcta-sigma (S = -1) / --traitK-s-pow
//   op 1:  x /= sqrt(het)                            --norm-effect on effects read from a file
//   op 2:  x *= sqrt(het)                            --norm-effect on output (.causals)
// sqrt and the division are correctly rounded on the device, so ops 1 and 2 give the host's
// bits; pow() is within 2 ulp of glibc's (op 0 is tolerance-matched, 1e-14).
__global__ void __launch_bounds__(256)
scale_effects_kernel(const double *__restrict__ freq, double *__restrict__ x1, double *__restrict__ x2, int64_t mc,
                     const int32_t *__restrict__ comp, const double *__restrict__ spow, int ncomp, int op,
                     unsigned long long *__restrict__ bad, int64_t index_base)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mc) return;
    const double f = freq[j];
    const double het = fabs(__dmul_rn(__dmul_rn(2.0, f), __dsub_rn(1.0, f)));      // :1250
    if (op != 2 && het < (double)FLT_EPSILON) { atomicMin(bad, (unsigned long long)(index_base + j)); return; }   // :1251, :1278; none at :1092
    const int c = comp ? comp[j] : 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double *x = k == 0 ? x1 : x2;
        if (!x) continue;
        double v = x[j];
        if (op == 0) v = __dmul_rn(v, sqrt(pow(het, spow[k * ncomp + c])));
        else if (op == 1) v = __ddiv_rn(v, sqrt(het));
        else v = __dmul_rn(v, sqrt(het));
        x[j] = v;
    }
}

// ---------------------------------------------------------------- argsort

constexpr int kSortTile = 1024;     // elements per warp
constexpr int kSortWarps = 4;       // warps per block

__device__ __forceinline__ uint64_t f64_sort_key(double v)
{
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // ascending order as unsigned
}

__global__ void sort_init_kernel(const double *__restrict__ y, int64_t n, uint64_t *__restrict__ keys,
                                 int32_t *__restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = f64_sort_key(y[i]); vals[i] = (int32_t)i; }
}

// hist[digit * tiles + tile] = number of keys of this tile with this digit
__global__ void __launch_bounds__(kSortWarps * 32)
sort_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, int64_t tiles,
                 uint32_t *__restrict__ hist)
{
    __shared__ uint32_t cnt[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * kSortWarps + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (tile < tiles) {
        const int64_t lo = tile * kSortTile, hi = min(n, lo + kSortTile);
        for (int64_t i = lo + lane; i < hi; i += 32)
            atomicAdd(&cnt[warp][(keys[i] >> shift) & 0xFF], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(int64_t)d * tiles + tile] = cnt[warp][d];
    }
}

// exclusive scan of hist (length len) in place, one block
__global__ void __launch_bounds__(1024)
sort_scan_kernel(uint32_t *__restrict__ hist, int64_t len)
{
    __shared__ uint32_t warp_sum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (len + 1023) / 1024;
    const int64_t lo = min(len, per * threadIdx.x), hi = min(len, lo + per);
    uint32_t s = 0;
    for (int64_t i = lo; i < hi; i++) s += hist[i];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sum[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sum[warp] + (incl - s);
    for (int64_t i = lo; i < hi; i++) { uint32_t v = hist[i]; hist[i] = run; run += v; }
}

// stable scatter: a warp walks its tile 32 keys at a time
__global__ void __launch_bounds__(kSortWarps * 32)
sort_scatter_kernel(const uint64_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n,
                    int shift, int64_t tiles, const uint32_t *__restrict__ offs,
                    uint64_t *__restrict__ keys_out, int32_t *__restrict__ vals_out)
{
    __shared__ uint32_t pos[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * kSortWarps + warp;
    if (tile >= tiles) return;
    for (int d = lane; d < 256; d += 32) pos[warp][d] = offs[(int64_t)d * tiles + tile];
    __syncwarp();
    const int64_t lo = tile * kSortTile, hi = min(n, lo + kSortTile);
    for (int64_t i0 = lo; i0 < hi; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool ok = i < hi;
        const uint64_t k = ok ? keys_in[i] : 0;
        const int32_t v = ok ? vals_in[i] : 0;
        const uint32_t d = ok ? (uint32_t)((k >> shift) & 0xFF) : 256u + lane;   // inactive lanes never match
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        if (ok) dst = pos[warp][d] + rank;
        __syncwarp();
        if (ok && rank == 0) pos[warp][d] += __popc(peers);
        __syncwarp();
        if (ok) { keys_out[dst] = k; vals_out[dst] = v; }
    }
}


// ---- the same sort as ONE cooperative launch: 8 passes x (histogram | scan | scatter) separated
// by grid-wide barriers instead of 25 kernel launches.  For N <= 500K the sort is pure launch
// latency (each kernel runs a few microseconds), so this is what a --cc step actually waits on.
constexpr int kCoopWarps = 8;
// kCoopTile = keys per warp-tile (all loaded into registers before they are used): 256 for
// N <= 200K (more warps, shorter dependent chains), 1024 above (fewer histogram columns to scan)

// ------------------------------------------------------ liability threshold (order statistic)
//
// apply_liability_threshold (source.cc:1031-1055) sorts the samples by liability and calls ranks
// [0, max_ncon) the control pool, the rest the case pool.  With the default --ncas / --ncon every
// sample of a pool is labelled (the shuffles only consume RNG state), so all that matters is on
// which side of the max_ncon-th order statistic a sample falls.  This kernel finds that statistic
// with an MSB-first radix select on the order-preserving 64-bit keys (8 passes, each a 256-bin
// histogram of the keys that still match the prefix) and marks the pools; ties at the threshold go
// to the controls in ascending sample index -- the order the stable argsort above defines.  One cooperative launch, 9 grid syncs, ~G x 256 global atomics per
// pass, against 25 grid syncs and 8 scatter passes for the full argsort.
constexpr int kSelThreads = 1024;
constexpr int kSelKeys = 4;        // keys a thread keeps in registers across the 8 passes (CACHED variant)

// Grid-wide barrier for a cooperative launch (all blocks co-resident): one arrival counter, monotonically
// growing targets.  ~1.5 us against ~8 us measured for cooperative_groups' grid.sync() on this part.
__device__ __forceinline__ void grid_barrier(uint32_t *counter, uint32_t target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

template <bool CACHED>
__global__ void __launch_bounds__(kSelThreads)
liability_split_kernel(const double *__restrict__ y, int64_t n, int64_t k, uint32_t *ghist /* [8][256] + barrier word, zeroed */,
                       uint32_t *geq /* [gridDim.x] */, uint8_t *__restrict__ is_case)
{
    uint32_t *barrier = ghist + 8 * 256;
    uint32_t phase = 0;
    __shared__ uint32_t hist[256];
    __shared__ uint32_t warp_sums[kSelThreads / 32];
    __shared__ uint64_t s_prefix;
    __shared__ int64_t s_rank;
    __shared__ uint32_t s_count;
    const int lane = threadIdx.x & 31;
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= 0 || k >= n) {                           // empty control or case pool: no threshold
        for (int64_t i = gtid; i < n; i += gthreads) is_case[i] = k <= 0;
        return;
    }
    uint64_t prefix = 0;                              // the threshold key T = key of rank k, found digit by digit
    int64_t rank = k;                                 // rank of T among the keys that match the prefix
    // CACHED: the grid covers all keys with kSelKeys per thread, loaded once (n <= grid * 1024 * kSelKeys)
    uint64_t mykeys[kSelKeys];
    if (CACHED) {
#pragma unroll
        for (int u = 0; u < kSelKeys; u++) {
            const int64_t i = ((int64_t)blockIdx.x * kSelKeys + u) * blockDim.x + threadIdx.x;
            mykeys[u] = i < n ? f64_sort_key(y[i]) : 0;
        }
    }
    auto tally = [&](const bool ok, const uint64_t key, const int pass, const int shift) {
        const bool match = ok && (pass == 7 || (key >> (shift + 8)) == (prefix >> (shift + 8)));
        const uint32_t d = match ? (uint32_t)((key >> shift) & 0xFF) : 256u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);               // one atomic per distinct digit and warp
        if (match && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&hist[d], (uint32_t)__popc(peers));
    };
    for (int pass = 7; pass >= 0; pass--) {
        const int shift = 8 * pass;
        for (int d = threadIdx.x; d < 256; d += blockDim.x) hist[d] = 0;
        __syncthreads();
        if (CACHED) {
#pragma unroll
            for (int u = 0; u < kSelKeys; u++)
                tally(((int64_t)blockIdx.x * kSelKeys + u) * blockDim.x + threadIdx.x < n, mykeys[u], pass, shift);
        } else {
            for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += gthreads) {
                const int64_t i = i0 + threadIdx.x;
                tally(i < n, i < n ? f64_sort_key(y[i]) : 0, pass, shift);
            }
        }
        __syncthreads();
        for (int d = threadIdx.x; d < 256; d += blockDim.x)
            if (hist[d]) atomicAdd(&ghist[pass * 256 + d], hist[d]);
        grid_barrier(barrier, ++phase * gridDim.x);
        if (threadIdx.x < 32) {                       // every block picks the digit for itself
            int64_t carry = 0, found_rank = -1;
            uint32_t found_d = 0, found_count = 0;
            uint32_t bins[8];
#pragma unroll
            for (int q = 0; q < 8; q++) bins[q] = __ldcg(&ghist[pass * 256 + 32 * q + lane]);   // one L2 round trip
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int d0 = 32 * q;
                const uint32_t v = bins[q];
                uint32_t incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += u;
                }
                const bool here = found_rank < 0 && carry + incl > rank;      // first bin whose cumulative count exceeds rank
                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, here);
                if (found_rank < 0 && ballot) {
                    const int src = __ffs(ballot) - 1;
                    const uint32_t excl = __shfl_sync(0xFFFFFFFFu, incl - v, src);
                    found_d = (uint32_t)(d0 + src);
                    found_rank = rank - (carry + excl);
                    found_count = __shfl_sync(0xFFFFFFFFu, v, src);
                }
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (lane == 0) { s_prefix = prefix | ((uint64_t)found_d << shift); s_rank = found_rank; s_count = found_count; }
        }
        __syncthreads();
        prefix = s_prefix;
        rank = s_rank;                                // = number of keys equal so far that sort before rank k
        const uint32_t candidates = s_count;
        __syncthreads();
        // ONE key is left under this prefix: it is T, and its lower digits are not needed -- every other key differs
        // from it in the digits found so far, so comparing with the prefix (lower digits zero) classifies them, and T
        // itself (>= prefix, rank 0 among its equals) falls on the case side as it must.  Continuous liabilities get
        // here after 3 passes instead of 8; ties run all passes.  (Every block sees the same histogram: uniform exit.)
        if (candidates == 1) break;
    }
    // prefix = T; rank = how many of the keys equal to T belong to the controls (lowest sample indices first).
    // Each block owns a contiguous range of samples, each thread a contiguous sub-range.
    const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per_block, hi = min(n, lo + per_block);
    const int64_t per_thread = (per_block + blockDim.x - 1) / blockDim.x;
    const int64_t tlo = min(hi, lo + (int64_t)threadIdx.x * per_thread), thi = min(hi, tlo + per_thread);
    uint32_t eq = 0;
    for (int64_t i = tlo; i < thi; i++) eq += f64_sort_key(y[i]) == prefix;
    // block-wide exclusive scan of eq
    uint32_t incl = eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t v = warp_sums[lane];
        uint32_t w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= o) w += u;
        }
        warp_sums[lane] = w - v;
        if (lane == 31) geq[blockIdx.x] = w;
    }
    __syncthreads();
    const uint32_t before_in_block = warp_sums[threadIdx.x >> 5] + incl - eq;
    grid_barrier(barrier, ++phase * gridDim.x);
    int64_t before = before_in_block;
    for (unsigned b = 0; b < blockIdx.x; b++) before += __ldcg(&geq[b]);
    for (int64_t i = tlo; i < thi; i++) {
        const uint64_t key = f64_sort_key(y[i]);
        bool control = key < prefix;
        if (key == prefix) { control = before < rank; before++; }
        is_case[i] = !control;
    }
}

template <int kCoopTile>
__global__ void __launch_bounds__(kCoopWarps * 32)
sort_coop_kernel(const double *__restrict__ y, int64_t n, uint64_t *k0, uint64_t *k1, int32_t *v0, int32_t *v1,
                 uint32_t *hist, int64_t tiles, int32_t *out_index, uint32_t *barrier /* zeroed */)
{
    uint32_t phase = 0;
    constexpr int kCoopPer = kCoopTile / 32;
    __shared__ uint32_t cnt[kCoopWarps][256];
    __shared__ uint32_t digit_base[256];
    uint32_t *digit_total = hist + tiles * 256;       // 256 extra words behind the histograms
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gwarp = (int64_t)blockIdx.x * kCoopWarps + warp, gwarps = (int64_t)gridDim.x * kCoopWarps;

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gthreads) {
        k0[i] = f64_sort_key(y[i]);
        v0[i] = (int32_t)i;
    }
    grid_barrier(barrier, ++phase * gridDim.x);

    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        // ---- per-tile digit histograms
        for (int64_t tile = gwarp; tile < tiles; tile += gwarps) {
            for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
            __syncwarp();
            const int64_t lo = tile * kCoopTile, hi = min(n, lo + kCoopTile);
            uint64_t kk[kCoopPer];
#pragma unroll
            for (int u = 0; u < kCoopPer; u++) { const int64_t i = lo + 32 * u + lane; kk[u] = (i < hi) ? k0[i] : 0; }
#pragma unroll
            for (int u = 0; u < kCoopPer; u++)
                if (lo + 32 * u + lane < hi) atomicAdd(&cnt[warp][(kk[u] >> shift) & 0xFF], 1u);
            __syncwarp();
            for (int d = lane; d < 256; d += 32) hist[(int64_t)d * tiles + tile] = cnt[warp][d];
            __syncwarp();
        }
        grid_barrier(barrier, ++phase * gridDim.x);
        // ---- exclusive scan over (digit major, tile minor): every warp of the grid scans the tiles
        // of some digits (coalesced loads + shuffles) and records the digit totals ...
        for (int64_t d = gwarp; d < 256; d += gwarps) {
            uint32_t carry = 0;
            for (int64_t t0 = 0; t0 < tiles; t0 += 32) {
                const int64_t t = t0 + lane;
                const uint32_t v = (t < tiles) ? hist[d * tiles + t] : 0u;
                uint32_t incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += u;
                }
                if (t < tiles) hist[d * tiles + t] = carry + incl - v;
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (lane == 0) digit_total[d] = carry;
        }
        grid_barrier(barrier, ++phase * gridDim.x);
        // ... which every block scans for itself (256 values) before scattering
        if (warp == 0) {
            uint32_t carry = 0;
            for (int d0 = 0; d0 < 256; d0 += 32) {
                const uint32_t v = digit_total[d0 + lane];
                uint32_t incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += u;
                }
                digit_base[d0 + lane] = carry + incl - v;
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
        __syncthreads();
        // ---- stable scatter, one warp per tile, 32 keys at a time
        int32_t *vout = (pass == 7) ? out_index : v1;
        for (int64_t tile = gwarp; tile < tiles; tile += gwarps) {
            for (int d = lane; d < 256; d += 32) cnt[warp][d] = hist[(int64_t)d * tiles + tile] + digit_base[d];
            __syncwarp();
            const int64_t lo = tile * kCoopTile, hi = min(n, lo + kCoopTile);
            uint64_t kk[kCoopPer];
            int32_t vv[kCoopPer];
#pragma unroll
            for (int u = 0; u < kCoopPer; u++) {
                const int64_t i = lo + 32 * u + lane;
                kk[u] = (i < hi) ? k0[i] : 0;
                vv[u] = (i < hi) ? v0[i] : 0;
            }
#pragma unroll
            for (int u = 0; u < kCoopPer; u++) {
                const bool ok = lo + 32 * u + lane < hi;
                const uint64_t k = kk[u];
                const int32_t v = vv[u];
                const uint32_t d = ok ? (uint32_t)((k >> shift) & 0xFF) : 256u + lane;
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                uint32_t dst = 0;
                if (ok) dst = cnt[warp][d] + rank;
                __syncwarp();
                if (ok && rank == 0) cnt[warp][d] += __popc(peers);
                __syncwarp();
                if (ok) { k1[dst] = k; vout[dst] = v; }
            }
            __syncwarp();
        }
        grid_barrier(barrier, ++phase * gridDim.x);
        uint64_t *tk = k0; k0 = k1; k1 = tk;
        int32_t *tv = v0; v0 = v1; v1 = tv;
    }
}

}  // namespace

int launch_sumsq(simu_b200_ctx *ctx, const double *d_y, int64_t n, double *d_out)
{
    if (n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "sumsq: n must be positive");
    const int64_t per_block = (int64_t)kRedThreads * kRedItems;
    const int64_t blocks = (n + per_block - 1) / per_block;
    double *partial = (double *)ctx_scratch(ctx, SCR_MISC, (size_t)blocks * sizeof(double));
    if (!partial) return SIMU_B200_ENOMEM;
    sumsq_partial_kernel<<<(unsigned)blocks, kRedThreads, 0, ctx->stream>>>(d_y, n, partial);
    sum_final_kernel<<<1, kRedThreads, 0, ctx->stream>>>(partial, blocks, d_out);
    ctx->launches += 2;
    return ctx_cuda(ctx, cudaGetLastError(), "sumsq launch");
}

int launch_heritability(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise,
                        const double *d_sumsq, double *d_coef)
{
    if (n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "heritability: n must be positive");
    heritability_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(hsq, d_y, n, d_noise, d_sumsq, d_coef);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "heritability launch");
}

int launch_apply_heritability(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise, double *d_coef)
{
    if (n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "heritability: n must be positive");
    int64_t grid = (n + 4 * kHerThreads - 1) / (4 * kHerThreads);
    const int64_t most = ctx->num_sms < kHerThreads ? ctx->num_sms : kHerThreads;      // co-resident, one partial per thread
    if (grid > most) grid = most;
    const bool fresh = ctx->scratch_bytes[SCR_HERIT] < 4096;
    uint8_t *scr = (uint8_t *)ctx_scratch(ctx, SCR_HERIT, 4096);
    if (!scr) return SIMU_B200_ENOMEM;
    if (fresh) CUDA_TRY(ctx, cudaMemsetAsync(scr, 0, 4096, ctx->stream));      // the barrier counters: the kernel leaves them zero
    unsigned int *bar = (unsigned int *)scr;
    double *partial = (double *)(scr + 64);
    void *args[] = {(void *)&hsq, (void *)&d_y, (void *)&n, (void *)&d_noise, (void *)&partial, (void *)&bar, (void *)&d_coef};
    const int ps = prof_begin(ctx, PROF_REDUCTIONS);
    cudaError_t e = cudaLaunchCooperativeKernel((void *)heritability_fused_kernel, dim3((unsigned)grid), dim3(kHerThreads), args, 0, ctx->stream);
    prof_end(ctx, ps);
    ctx->launches++;
    return ctx_cuda(ctx, e, "heritability_fused_kernel launch");
}

int launch_scale_effects_on(simu_b200_ctx *ctx, cudaStream_t stream, const double *d_freq, double *d_x1, double *d_x2,
                            int64_t mc, const int32_t *d_comp, const double *d_spow, int ncomp, int op,
                            unsigned long long *d_bad, int64_t index_base)
{
    if (mc <= 0) return SIMU_B200_OK;
    scale_effects_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(d_freq, d_x1, d_x2, mc, d_comp, d_spow, ncomp, op, d_bad,
                                                                         index_base);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "scale_effects_kernel launch");
}

int launch_scale_effects(simu_b200_ctx *ctx, const double *d_freq, double *d_x1, double *d_x2, int64_t mc,
                         const int32_t *d_comp, const double *d_spow, int ncomp, int op, unsigned long long *d_bad)
{
    return launch_scale_effects_on(ctx, ctx->stream, d_freq, d_x1, d_x2, mc, d_comp, d_spow, ncomp, op, d_bad, 0);
}

int launch_liability_split(simu_b200_ctx *ctx, const double *d_pheno, int64_t n, int64_t max_ncon, uint8_t *d_is_case)
{
    if (n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "liability_split: n must be positive");
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (!coop) return ctx_fail(ctx, SIMU_B200_ECUDA, "liability_split: device lacks cooperative launch");
    const int64_t per_block = (int64_t)kSelThreads * kSelKeys;
    const bool cached = (n + per_block - 1) / per_block <= ctx->num_sms;      // every key in a register of some thread
    const int64_t blocks = cached ? (n + per_block - 1) / per_block : ctx->num_sms;
    uint32_t *scr = (uint32_t *)ctx_scratch(ctx, SCR_SORT, (size_t)(8 * 256 + 32 + ctx->num_sms) * 4);
    if (!scr) return SIMU_B200_ENOMEM;
    CUDA_TRY(ctx, cudaMemsetAsync(scr, 0, (8 * 256 + 32) * 4, ctx->stream));      // histograms + the barrier word
    uint32_t *ghist = scr, *geq = scr + 8 * 256 + 32;
    void *args[] = {(void *)&d_pheno, (void *)&n, (void *)&max_ncon, (void *)&ghist, (void *)&geq, (void *)&d_is_case};
    const int ps = prof_begin(ctx, PROF_SORT);
    cudaError_t e = cudaLaunchCooperativeKernel(cached ? (void *)liability_split_kernel<true> : (void *)liability_split_kernel<false>,
                                                dim3((unsigned)blocks), dim3(kSelThreads), args, 0, ctx->stream);
    prof_end(ctx, ps);
    ctx->launches++;
    return ctx_cuda(ctx, e, "liability_split launch");
}

int launch_argsort_f64(simu_b200_ctx *ctx, const double *d_key, int64_t n, int32_t *d_index)
{
    if (n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "argsort: n must be positive");
    const int64_t tiles = (n + kSortTile - 1) / kSortTile;
    const size_t kb = ((size_t)n * 8 + 255) & ~(size_t)255, vb = ((size_t)n * 4 + 255) & ~(size_t)255;
    const size_t hb = (((size_t)(n + 255) / 256 + 1) * 256 * 4 + 255) & ~(size_t)255;   // >= every tile layout
    uint8_t *s = (uint8_t *)ctx_scratch(ctx, SCR_SORT, 2 * kb + 2 * vb + hb);
    if (!s) return SIMU_B200_ENOMEM;
    uint64_t *k0 = (uint64_t *)s, *k1 = (uint64_t *)(s + kb);
    int32_t *v0 = (int32_t *)(s + 2 * kb), *v1 = (int32_t *)(s + 2 * kb + vb);
    uint32_t *hist = (uint32_t *)(s + 2 * kb + 2 * vb);
    if (ctx->coop_sort_blocks < 0) {      // how many blocks of the cooperative kernel fit on this device
        int per_sm = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sort_coop_kernel<1024>, kCoopWarps * 32, 0) == cudaSuccess)
            ctx->coop_sort_blocks = per_sm * ctx->num_sms;
        else
            ctx->coop_sort_blocks = 0;
        cudaGetLastError();
    }
    if (ctx->coop_sort_blocks > 0) {
        const int tile = n <= 200000 ? 256 : 1024;
        int64_t ctiles = (n + tile - 1) / tile;
        int64_t blocks = (ctiles + kCoopWarps - 1) / kCoopWarps;
        blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, ctx->coop_sort_blocks));
        uint32_t *barrier = (uint32_t *)ctx_scratch(ctx, SCR_SYNC, 256);
        if (!barrier) return SIMU_B200_ENOMEM;
        CUDA_TRY(ctx, cudaMemsetAsync(barrier, 0, 4, ctx->stream));
        void *args[] = {(void *)&d_key, (void *)&n, (void *)&k0, (void *)&k1, (void *)&v0, (void *)&v1,
                        (void *)&hist, (void *)&ctiles, (void *)&d_index, (void *)&barrier};
        const int ps = prof_begin(ctx, PROF_SORT);
        void *kern = tile == 256 ? (void *)sort_coop_kernel<256> : (void *)sort_coop_kernel<1024>;
        cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(kCoopWarps * 32), args, 0, ctx->stream);
        prof_end(ctx, ps);
        ctx->launches++;
        return ctx_cuda(ctx, e, "cooperative argsort launch");
    }
    // devices without cooperative launch: the same passes as separate kernels
    const unsigned tb = (unsigned)((tiles + kSortWarps - 1) / kSortWarps);
    sort_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_key, n, k0, v0);
    ctx->launches++;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        int32_t *vout = (pass == 7) ? d_index : v1;
        sort_hist_kernel<<<tb, kSortWarps * 32, 0, ctx->stream>>>(k0, n, shift, tiles, hist);
        sort_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, tiles * 256);
        sort_scatter_kernel<<<tb, kSortWarps * 32, 0, ctx->stream>>>(k0, v0, n, shift, tiles, hist, k1, vout);
        ctx->launches += 3;
        uint64_t *tk = k0; k0 = k1; k1 = tk;
        int32_t *tv = v0; v0 = v1; v1 = tv;
    }
    return ctx_cuda(ctx, cudaGetLastError(), "argsort launch");
}
