// tcgen05 bring-up probe for the tensor-core phenotype kernel (sm_100a only).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
//
// Checks, against a host computation, that
//   T1  one tcgen05.mma.kind::i8 (M=128, N=16, K=32, A = u8 from TMEM, B = s8 from shared memory in the
//       canonical K-major no-swizzle layout, D = s32 in TMEM) gives D = A * B^T for the TMEM/descriptor
//       conventions the kernel assumes (A(m,k) at lane m, column k/4, byte k%4; LBO = K-chunk stride,
//       SBO = 8-row-group stride);
//   T2  four accumulating MMAs over K = 128;
// and measures
//   T3  back-to-back MMA issue rate and tcgen05.st rate on one SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    while (!mbar_test(bar, parity)) if (clock64() - t0 > 2000000000ll) __trap();
}

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline uint64_t make_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t make_idesc_i8(int m, int n, int a_signed, int b_signed)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A: [128][kbytes] u8, B: [n][kbytes] s8, D: [128][n] s32; kbytes = 32 * nmma
__global__ void __launch_bounds__(128) probe_mma(const uint8_t *A, const int8_t *B, int32_t *D, int n, int nmma, int swap_lbo_sbo)
{
    __shared__ __align__(1024) uint8_t bs[4 * 1024];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int t = threadIdx.x, warp = t >> 5;
    const int ngroups = n / 8;
    const uint32_t chunk = (uint32_t)ngroups * 128;          // bytes of one 16-byte K chunk over all N rows
    // B sub-tile q (32 K-bytes) = 2 chunks; element (r, k): q*2*chunk + (k/16)*chunk + (r/8)*128 + (r%8)*16 + k%16
    for (int i = t; i < n * 32 * nmma; i += 128) {
        const int r = i / (32 * nmma), k = i % (32 * nmma);
        const int q = k / 32, kk = k % 32;
        bs[q * 2 * chunk + (kk / 16) * chunk + (r / 8) * 128 + (r % 8) * 16 + kk % 16] = (uint8_t)B[r * 32 * nmma + k];
    }
    if (t == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 64);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tbase;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    // A row t -> TMEM lane t, columns 32 .. 32 + 8*nmma
    for (int q = 0; q < nmma; q++) {
        uint32_t r[8];
        for (int c = 0; c < 8; c++) {
            const uint8_t *p = A + (size_t)t * 32 * nmma + q * 32 + c * 4;
            r[c] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        }
        st8(lane_base + 32 + q * 8, r);
    }
    wait_st();
    fence_before();
    __syncthreads();
    fence_after();
    if (t == 0) {
        const uint32_t idesc = make_idesc_i8(128, n, 0, 1);
        for (int q = 0; q < nmma; q++) {
            const uint32_t lbo = swap_lbo_sbo ? 128 : chunk, sbo = swap_lbo_sbo ? chunk : 128;
            mma_i8_ts(tm, tm + 32 + q * 8, make_bdesc(smem_u32(bs) + q * 2 * chunk, lbo, sbo), idesc, q > 0);
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    fence_after();
    uint32_t d[16];
    ld16(lane_base, d);
    for (int j = 0; j < n; j++) D[t * n + j] = (int32_t)d[j];
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

// throughput: out[0] = clocks for `iters` back-to-back MMAs issued by one thread, rotating over `nacc`
// accumulators; out[1] = clocks for `iters` tcgen05.st.x32 per warp with blockDim/32 warps in parallel
__global__ void __launch_bounds__(512) probe_rate(long long *out, int iters, int n, int nacc)
{
    __shared__ __align__(1024) uint8_t bs[8192];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 8192; i += blockDim.x) bs[i] = (uint8_t)(i & 3);
    if (t == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tbase;
    const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; c++) r[c] = t * 131 + c;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 2) {
        st32(lane_base + 256 + (warp >> 2) * 64, r);
        st32(lane_base + 256 + (warp >> 2) * 64 + 32, r);
    }
    wait_st();
    long long t1 = clock64();
    if (t == 0) out[1] = t1 - t0;
    fence_before();
    __syncthreads();
    fence_after();
    if (t == 0) {
        const uint32_t idesc = make_idesc_i8(128, n, 0, 1);
        const uint64_t bd = make_bdesc(smem_u32(bs), (uint32_t)(n / 8) * 128, 128);
        t0 = clock64();
        int a = 0;
        for (int i = 0; i < iters; i++) {
            mma_i8_ts(tm + a * 32, tm + 256 + (i & 15) * 8, bd, idesc, 1);
            if (++a == nacc) a = 0;
        }
        mma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        out[0] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}


// MMA issue rate with `issuers` warps each issuing `iters` MMAs (unrolled x4, operands precomputed) into its own
// accumulator: separates "one thread cannot issue faster" from "the tensor pipe cannot retire faster"
__global__ void __launch_bounds__(256) probe_issue(long long *out, int iters, int n, int issuers, int same_acc)
{
    __shared__ __align__(1024) uint8_t bs[8192];
    __shared__ __align__(8) uint64_t bar[8];
    __shared__ uint32_t tbase;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 8192; i += blockDim.x) bs[i] = (uint8_t)(i & 3);
    if (t == 0) { for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tbase;
    const long long t0 = clock64();
    if (warp < issuers && lane == 0) {
        const uint32_t idesc = make_idesc_i8(128, n, 0, 1);
        const uint64_t bd = make_bdesc(smem_u32(bs), (uint32_t)(n / 8) * 128, 128);
        const uint32_t d = tm + (same_acc ? 0 : warp * 64), a = tm + 256 + warp * 32;
        for (int i = 0; i < iters; i += 4) {
            mma_i8_ts(d, a, bd, idesc, 1);
            mma_i8_ts(d, a + 8, bd, idesc, 1);
            mma_i8_ts(d, a + 16, bd, idesc, 1);
            mma_i8_ts(d, a + 24, bd, idesc, 1);
        }
        mma_commit(smem_u32(&bar[warp]));
        out[2 + warp] = clock64() - t0;          // issue time alone
    }
    if (t == 0) {
        for (int w = 0; w < issuers; w++) mbar_wait(smem_u32(&bar[w]), 0);
        out[0] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}


// as probe_issue, but the whole warp runs the issue loop with warp-uniform operands and one elected lane
// issues (the CUTLASS pattern): the descriptors can then live in uniform registers
__device__ __forceinline__ void mma_i8_ts_elect(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__global__ void __launch_bounds__(256) probe_issue_u(long long *out, int iters, int n, int issuers, int same_acc)
{
    __shared__ __align__(1024) uint8_t bs[8192];
    __shared__ __align__(8) uint64_t bar[8];
    __shared__ uint32_t tbase;
    const int t = threadIdx.x, lane = t & 31;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    for (int i = t; i < 8192; i += blockDim.x) bs[i] = (uint8_t)(i & 3);
    if (t == 0) { for (int i = 0; i < 8; i++) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tbase;
    const long long t0 = clock64();
    if (warp < issuers) {
        const uint32_t idesc = make_idesc_i8(128, n, 0, 1);
        const uint64_t bd = make_bdesc(smem_u32(bs), (uint32_t)(n / 8) * 128, 128);
        const uint32_t d = tm + (same_acc ? 0 : warp * 64), a = tm + 256 + warp * 32;
        for (int i = 0; i < iters; i += 4) {
            mma_i8_ts_elect(d, a, bd, idesc, 1);
            mma_i8_ts_elect(d, a + 8, bd + 32, idesc, 1);
            mma_i8_ts_elect(d, a + 16, bd + 64, idesc, 1);
            mma_i8_ts_elect(d, a + 24, bd, idesc, 1);
        }
        if (lane == 0) {
            mma_commit(smem_u32(&bar[warp]));
            out[2 + warp] = clock64() - t0;          // issue time alone
        }
    }
    if (t == 0) {
        for (int w = 0; w < issuers; w++) mbar_wait(smem_u32(&bar[w]), 0);
        out[0] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}


// latencies of the synchronisation primitives the kernel is made of, one warp, 512 repetitions each (clk per op)
__global__ void __launch_bounds__(128) probe_lat(long long *out)
{
    __shared__ __align__(1024) uint8_t bs[8192];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tbase;
    const int t = threadIdx.x, lane = t & 31;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    for (int i = t; i < 8192; i += blockDim.x) bs[i] = (uint8_t)(i & 3);
    if (t == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tbase;
    const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int R = 512;
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; c++) r[c] = t * 131 + c;
    long long t0, t1;
    // (0) try_wait on a barrier whose phase already completed (parity 1 of a fresh barrier)
    t0 = clock64();
    for (int i = 0; i < R; i++) mbar_wait(smem_u32(&bar[1]), 1);
    t1 = clock64();
    if (t == 0) out[0] = (t1 - t0) / R;
    // (1) st.x32 + wait::st
    t0 = clock64();
    for (int i = 0; i < R; i++) { st32(lane_base + 256, r); wait_st(); }
    t1 = clock64();
    if (t == 0) out[1] = (t1 - t0) / R;
    // (2) fence::before + fence::after pair
    t0 = clock64();
    for (int i = 0; i < R; i++) { fence_before(); fence_after(); }
    t1 = clock64();
    if (t == 0) out[2] = (t1 - t0) / R;
    // (3) bar.sync of 128 threads
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < R; i++) asm volatile("bar.sync 1, 128;" ::: "memory");
    t1 = clock64();
    if (t == 0) out[3] = (t1 - t0) / R;
    __syncthreads();
    if (warp == 0) {
        const uint32_t idesc = make_idesc_i8(128, 16, 0, 1);
        const uint64_t bd = make_bdesc(smem_u32(bs), 256, 128);
        // (4) commit alone -> wait for its arrival
        uint32_t ph = 0;
        t0 = clock64();
        for (int i = 0; i < R; i++) {
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar[0])) : "memory");
            mbar_wait(smem_u32(&bar[0]), ph); ph ^= 1;
        }
        t1 = clock64();
        if (lane == 0) out[4] = (t1 - t0) / R;
        // (5) 4 MMAs + commit -> wait for the arrival (issue + execution + signalling latency)
        t0 = clock64();
        for (int i = 0; i < R; i++) {
            mma_i8_ts_elect(tm, tm + 256, bd, idesc, 1); mma_i8_ts_elect(tm, tm + 264, bd + 32, idesc, 1);
            mma_i8_ts_elect(tm, tm + 272, bd + 64, idesc, 1); mma_i8_ts_elect(tm, tm + 280, bd + 96, idesc, 1);
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar[0])) : "memory");
            mbar_wait(smem_u32(&bar[0]), ph); ph ^= 1;
        }
        t1 = clock64();
        if (lane == 0) out[5] = (t1 - t0) / R;
        // (6) issue cost alone: 4 MMAs + commit, no wait (one wait at the end)
        t0 = clock64();
        for (int i = 0; i < R; i++) {
            mma_i8_ts_elect(tm, tm + 256, bd, idesc, 1); mma_i8_ts_elect(tm, tm + 264, bd + 32, idesc, 1);
            mma_i8_ts_elect(tm, tm + 272, bd + 64, idesc, 1); mma_i8_ts_elect(tm, tm + 280, bd + 96, idesc, 1);
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar[2])) : "memory");
        }
        t1 = clock64();
        if (lane == 0) out[6] = (t1 - t0) / R;
        // (7) mbarrier.arrive (count 1) + try_wait on it: software signalling round trip inside a warp
        t0 = clock64();
        uint32_t ph3 = 0;
        for (int i = 0; i < R; i++) {
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[3])) : "memory");
            mbar_wait(smem_u32(&bar[3]), ph3); ph3 ^= 1;
        }
        t1 = clock64();
        if (lane == 0) out[7] = (t1 - t0) / R;
        // (8) 4 x ld.shared.u32 dependent use
        t0 = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < R; i++) {
            uint32_t a, b, c, d;
            const uint32_t src = smem_u32(bs) + ((lane * 4 + (acc & 3) * 128) & 8191);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(src));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b) : "r"(src + 2048));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(src + 4096));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(src + 6144));
            acc += a + b + c + d;
        }
        t1 = clock64();
        if (lane == 0) { out[8] = (t1 - t0) / R; out[15] = acc; }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

static int run_case(int n, int nmma, int swap)
{
    const int kb = 32 * nmma;
    uint8_t *hA = (uint8_t *)malloc(128 * kb);
    int8_t *hB = (int8_t *)malloc(n * kb);
    int32_t *hD = (int32_t *)malloc(128 * n * 4), *ref = (int32_t *)malloc(128 * n * 4);
    uint32_t s = 12345u + n * 7 + nmma;
    for (int i = 0; i < 128 * kb; i++) { s = s * 1664525u + 1013904223u; hA[i] = (uint8_t)(s >> 24); }
    for (int i = 0; i < n * kb; i++) { s = s * 1664525u + 1013904223u; hB[i] = (int8_t)(s >> 24); }
    for (int m = 0; m < 128; m++)
        for (int j = 0; j < n; j++) {
            int32_t acc = 0;
            for (int k = 0; k < kb; k++) acc += (int32_t)hA[m * kb + k] * (int32_t)hB[j * kb + k];
            ref[m * n + j] = acc;
        }
    uint8_t *dA; int8_t *dB; int32_t *dD;
    CK(cudaMalloc(&dA, 128 * kb)); CK(cudaMalloc(&dB, n * kb)); CK(cudaMalloc(&dD, 128 * n * 4));
    CK(cudaMemcpy(dA, hA, 128 * kb, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, n * kb, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xEE, 128 * n * 4));
    probe_mma<<<1, 128>>>(dA, dB, dD, n, nmma, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d nmma=%d swap=%d: kernel failed: %s\n", n, nmma, swap, cudaGetErrorString(e)); return -1; }
    CK(cudaMemcpy(hD, dD, 128 * n * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 128 * n; i++) bad += hD[i] != ref[i];
    printf("N=%d nmma=%d swap_lbo_sbo=%d: %d / %d mismatches%s  (D[0][0..3] = %d %d %d %d, ref %d %d %d %d; D[77][5] = %d ref %d)\n", n, nmma, swap,
           bad, 128 * n, bad ? "" : "  OK", hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3], hD[77 * n + 5], ref[77 * n + 5]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    free(hA); free(hB); free(hD); free(ref);
    return bad;
}

int main(int argc, char **argv)
{
    const char *what = argc > 1 ? argv[1] : "all";
    if (!strcmp(what, "mma") || !strcmp(what, "all")) {
        const int swap = argc > 2 ? atoi(argv[2]) : 0;
        run_case(16, 1, swap);
        run_case(16, 4, swap);
        run_case(8, 1, swap);
        run_case(8, 4, swap);
        run_case(32, 4, swap);
    }
    if (!strcmp(what, "rate") || !strcmp(what, "all")) {
        long long *d, h[2];
        CK(cudaMalloc(&d, 16));
        for (int warps = 4; warps <= 16; warps *= 2)
            for (int nacc = 1; nacc <= 8; nacc *= 2)
                for (int n = 8; n <= 64; n *= 2) {
                    if (warps > 4 && (nacc > 1 || n != 16)) continue;
                    probe_rate<<<1, warps * 32>>>(d, 4096, n, nacc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("rate N=%d failed: %s\n", n, cudaGetErrorString(e)); return 1; }
                    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                    printf("rate N=%d nacc=%d warps=%d: %.2f clk per MMA (M=128, K=32, A in TMEM), %.2f clk per tcgen05.st.32x32b.x32 per warp\n",
                           n, nacc, warps, h[0] / 4096.0, h[1] / 4096.0);
                }
    }
    if (!strcmp(what, "issue") || !strcmp(what, "all")) {
        long long *d, h[8];
        CK(cudaMalloc(&d, 64));
        for (int same = 0; same <= 1; same++)
            for (int issuers = 1; issuers <= 4; issuers *= 2)
                for (int n = 8; n <= 128; n *= 4) {
                    probe_issue<<<1, 256>>>(d, 4096, n, issuers, same);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("issue N=%d failed: %s\n", n, cudaGetErrorString(e)); return 1; }
                    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
                    printf("issue N=%d issuers=%d same_acc=%d: %.2f clk per MMA overall (%.2f clk per MMA per issuer to issue)\n", n, issuers, same,
                           h[0] / (4096.0 * issuers), h[2] / 4096.0);
                }
    }
    if (!strcmp(what, "lat") || !strcmp(what, "all")) {
        long long *d, h[16];
        CK(cudaMalloc(&d, 128));
        probe_lat<<<1, 128>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lat failed: %s\n", cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
        printf("latency (clk): try_wait(complete) %lld | st.x32+wait::st %lld | fence before+after %lld | bar.sync(128) %lld | commit->arrival %lld | "
               "4 MMA+commit->arrival %lld | 4 MMA+commit issue only %lld | arrive+try_wait %lld | 4 LDS+use %lld\n",
               h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
    }
    if (!strcmp(what, "issueu") || !strcmp(what, "all")) {
        long long *d, h[8];
        CK(cudaMalloc(&d, 64));
        for (int same = 0; same <= 1; same++)
            for (int issuers = 1; issuers <= 4; issuers *= 2)
                for (int n = 8; n <= 128; n *= 4) {
                    probe_issue_u<<<1, 256>>>(d, 4096, n, issuers, same);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("issueu N=%d failed: %s\n", n, cudaGetErrorString(e)); return 1; }
                    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
                    printf("issue(uniform) N=%d issuers=%d same_acc=%d: %.2f clk per MMA overall (%.2f clk per MMA per issuer to issue)\n", n, issuers, same,
                           h[0] / (4096.0 * issuers), h[2] / 4096.0);
                }
    }
    return 0;
}
