// The path's one exchange step over NVLink peer memory instead of NCCL: every rank
// owns a slice of the causal SNPs and a partial N x K phenotype; the sum over ranks is 0.8-8 MB,
// i.e. latency-bound.  One-shot all-reduce: each rank publishes its partial in a symmetric buffer
// (CUDA IPC, one process per GPU), raises a flag in every peer's memory, waits for the peers' flags
// and sums all partials straight out of the peers' HBM through NVSwitch, in rank order -- the
// result is identical on all ranks and independent of timing.  Two buffers (epoch parity) make the
// exchange safe without a trailing barrier: a rank can only write buffer e & 1 again after the
// all-reduce of epoch e+1, which needs every rank's flag for e+1, which is raised after that
// rank's reads of epoch e.
//
// With more than two ranks the one-shot form reads (world - 1) x the payload per rank (8 ranks, 1.6 MB:
// 35 us against NCCL's 30).  The TWO-SHOT form below moves 2 (world - 1) / world of it: every rank
// sums ITS 1/world slice of the samples out of the peers' partials (reduce-scatter), publishes the
// reduced slice, and gathers the peers' slices (all-gather) -- still one cooperative launch, two
// flag rounds, sums in rank order, the same bits on every rank.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "ctx.cuh"

namespace {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
    const double *buf[kMaxPeers];      // each rank's symmetric buffer: partials [2][count], then reduced slices [2][count]
    uint32_t *flags[kMaxPeers];        // each rank's flag block:       [2 rounds][2][kMaxPeers]
};

__device__ __forceinline__ void grid_barrier_sys(uint32_t *barrier, uint32_t target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd(barrier, 1u);
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(barrier) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t ld_state(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// raise flag `round` of this rank's epoch in every peer's memory, then wait for every peer's
__device__ __forceinline__ void flag_round(const PeerPtrs &pp, int round, int rank, int world, int parity, uint32_t epoch)
{
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[threadIdx.x] + (round * 2 + parity) * kMaxPeers + rank), "r"(epoch)
                     : "memory");
    }
    if (threadIdx.x < world) {                      // my flag block is written by the peers over NVLink
        const uint32_t *f = pp.flags[rank] + (round * 2 + parity) * kMaxPeers + threadIdx.x;
        uint32_t v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (clock64() - t0 > 20000000000ll) __trap();      // a peer died: fail, do not hang
        } while (v < epoch);
    }
    __syncthreads();
}

// publish -> grid barrier -> flag the peers -> wait for their flags -> sum, as ONE cooperative launch
// (three separate launches cost ~16 us of launch latency for ~5 us of work)
constexpr int kPeerThreads = 1024;

__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(PeerPtrs pp, int rank, int world, int64_t count, double *__restrict__ y, uint32_t *state)
{
    // The epoch lives in DEVICE memory (state[1]; state[0] is the grid-barrier counter, monotonic): a launch derives its
    // epoch, buffer parity and barrier target from it and advances it after its last grid barrier, when every block
    // has read it.  Nothing in the launch parameters changes from call to call, so the step that contains the exchange
    // can be captured into a CUDA graph and replayed.
    const uint32_t epoch = ld_state(state + 1) + 1u;
    const int parity = (int)(epoch & 1u);
    double *mine = const_cast<double *>(pp.buf[rank]) + (int64_t)parity * count;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = first; i < count; i += stride) mine[i] = y[i];
    // grid barrier: every block's part of the partial is written (system scope: the peers will read it)
    grid_barrier_sys(state, epoch * gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) st_state(state + 1, epoch);
    flag_round(pp, 0, rank, world, parity, epoch);
    // pairs of doubles (the buffers are 16-byte aligned when count is even), all peers' loads in flight at once;
    // rank order of the adds: the same bits on every rank
    const int64_t pairs = ((count & 1) || (reinterpret_cast<uintptr_t>(y) & 15)) ? 0 : count / 2;
    for (int64_t i = first; i < pairs; i += stride) {
        double s0 = 0.0, s1 = 0.0;
        for (int r0 = 0; r0 < world; r0 += 8) {     // 8 peers' loads in flight at a time
            double a[8], b[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (r0 + q < world)
                    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(a[q]), "=d"(b[q])
                                 : "l"(pp.buf[r0 + q] + (int64_t)parity * count + 2 * i));
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (r0 + q < world) { s0 += a[q]; s1 += b[q]; }
        }
        reinterpret_cast<double2 *>(y)[i] = make_double2(s0, s1);
    }
    for (int64_t i = 2 * pairs + first; i < count; i += stride) {
        double s = 0.0;
        for (int r = 0; r < world; r++) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(pp.buf[r] + (int64_t)parity * count + i));
            s += v;
        }
        y[i] = s;
    }
}

// two-shot: publish | barrier | flags A | reduce my slice from all partials | barrier | flags B | gather the slices.
// count2 = count / 2 pairs of doubles (the host routes odd counts and unaligned y to the one-shot kernel).
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce2_kernel(PeerPtrs pp, int rank, int world, int64_t count, double *__restrict__ y, uint32_t *state)
{
    const uint32_t epoch = ld_state(state + 1) + 1u;             // see peer_allreduce_kernel
    const int parity = (int)(epoch & 1u);
    double *mine = const_cast<double *>(pp.buf[rank]) + (int64_t)parity * count;
    double *my_slices = const_cast<double *>(pp.buf[rank]) + (int64_t)(2 + parity) * count;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pairs = count / 2;
    for (int64_t i = first; i < pairs; i += stride) reinterpret_cast<double2 *>(mine)[i] = reinterpret_cast<const double2 *>(y)[i];
    grid_barrier_sys(state, (2u * epoch - 1u) * gridDim.x);
    flag_round(pp, 0, rank, world, parity, epoch);
    // reduce-scatter: pairs [lo, hi) are this rank's; every peer's partial is read for that slice only
    const int64_t lo = pairs * rank / world, hi = pairs * (rank + 1) / world;
    for (int64_t i = lo + first; i < hi; i += stride) {
        double s0 = 0.0, s1 = 0.0;
        for (int r0 = 0; r0 < world; r0 += 8) {
            double a[8], b[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (r0 + q < world)
                    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(a[q]), "=d"(b[q])
                                 : "l"(pp.buf[r0 + q] + (int64_t)parity * count + 2 * i));
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (r0 + q < world) { s0 += a[q]; s1 += b[q]; }
        }
        reinterpret_cast<double2 *>(my_slices)[i] = make_double2(s0, s1);
        reinterpret_cast<double2 *>(y)[i] = make_double2(s0, s1);
    }
    grid_barrier_sys(state, 2u * epoch * gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) st_state(state + 1, epoch);
    flag_round(pp, 1, rank, world, parity, epoch);
    // all-gather: the peers' reduced slices, straight into y
    for (int64_t i = first; i < pairs; i += stride) {
        if (i >= lo && i < hi) continue;
        int owner = (int)(((i + 1) * world - 1) / pairs);          // largest r with pairs * r / world <= i
        double a, b;
        asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b)
                     : "l"(pp.buf[owner] + (int64_t)(2 + parity) * count + 2 * i));
        reinterpret_cast<double2 *>(y)[i] = make_double2(a, b);
    }
}

}  // namespace

struct simu_b200_peer {
    int rank = 0, world = 1;
    int64_t count = 0;
    double *sym = nullptr;             // [2][count] then the flag block
    void *peer_base[kMaxPeers] = {nullptr};
    PeerPtrs pp;
    uint32_t *barrier = nullptr;       // device state: [0] grid-barrier counter (monotonic), [1] epoch of the last all-reduce
    bool connected = false;
    int form = -1;                     // -1 not used yet, 0 one-shot, 1 two-shot
};

static size_t sym_bytes(int64_t count) { return (size_t)4 * count * sizeof(double) + 4 * kMaxPeers * sizeof(uint32_t); }

extern "C" {

int simu_b200_peer_init(simu_b200_ctx *ctx, int rank, int world, int64_t count, void *handle_out)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || count <= 0 || !handle_out)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "peer_init: bad arguments (world <= %d)", kMaxPeers);
    if (ctx->peer) return ctx_fail(ctx, SIMU_B200_EINVAL, "peer_init: already initialised");
    simu_b200_peer *p = new (std::nothrow) simu_b200_peer();
    if (!p) return SIMU_B200_ENOMEM;
    p->rank = rank; p->world = world; p->count = count;
    cudaError_t e = cudaMalloc((void **)&p->sym, sym_bytes(count));
    if (e != cudaSuccess) { delete p; return ctx_cuda(ctx, e, "peer_init cudaMalloc"); }
    cudaMemset(p->sym, 0, sym_bytes(count));
    e = cudaMalloc((void **)&p->barrier, 256);
    if (e != cudaSuccess) { cudaFree(p->sym); delete p; return ctx_cuda(ctx, e, "peer_init cudaMalloc"); }
    cudaMemset(p->barrier, 0, 256);
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p->sym);
    if (e != cudaSuccess) { cudaFree(p->sym); delete p; return ctx_cuda(ctx, e, "cudaIpcGetMemHandle"); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    ctx->peer = p;
    return SIMU_B200_OK;
}

int simu_b200_peer_connect(simu_b200_ctx *ctx, const void *handles)
{
    if (!ctx || !ctx->peer || !handles) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    simu_b200_peer *p = ctx->peer;
    for (int r = 0; r < p->world; r++) {
        void *base = p->sym;
        if (r != p->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)handles + 64 * r, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return ctx_cuda(ctx, e, "cudaIpcOpenMemHandle (peer access over NVLink)");
            p->peer_base[r] = base;
        }
        p->pp.buf[r] = (const double *)base;
        p->pp.flags[r] = (uint32_t *)((char *)base + (size_t)4 * p->count * sizeof(double));
    }
    p->connected = true;
    return SIMU_B200_OK;
}

int simu_b200_peer_allreduce_dev(simu_b200_ctx *ctx, double *d_y, int64_t count)
{
    if (!ctx || !ctx->peer || !ctx->peer->connected) return ctx_fail(ctx, SIMU_B200_EINVAL, "peer_allreduce: not connected");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    simu_b200_peer *p = ctx->peer;
    if (!d_y || count != p->count) return ctx_fail(ctx, SIMU_B200_EINVAL, "peer_allreduce: count differs from peer_init");
    int blocks = (int)std::min<int64_t>((count + 2 * kPeerThreads - 1) / (2 * kPeerThreads), (int64_t)ctx->num_sms);   // co-resident
    // more than two ranks: two-shot (reduce-scatter + all-gather), which moves 2 (world-1)/world of the payload per
    // rank instead of (world-1) x; SIMU_B200_PEER_ONESHOT=1 forces the one-shot form (A/B runs).  The form must not
    // change between calls on one context: the grid-barrier counter advances by one or two rounds per epoch.
    const bool two_shot = p->world > 2 && !(count & 1) && !(reinterpret_cast<uintptr_t>(d_y) & 15) && !getenv("SIMU_B200_PEER_ONESHOT");
    if (p->form >= 0 && p->form != (int)two_shot)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "peer_allreduce: buffer alignment changed between calls");
    p->form = (int)two_shot;
    void *args[] = {(void *)&p->pp, (void *)&p->rank, (void *)&p->world, (void *)&count, (void *)&d_y, (void *)&p->barrier};
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel(two_shot ? (void *)peer_allreduce2_kernel : (void *)peer_allreduce_kernel,
                                              dim3((unsigned)blocks), dim3(kPeerThreads), args, 0, ctx->stream));
    ctx->launches += 1;
    return ctx_cuda(ctx, cudaGetLastError(), "peer_allreduce launch");
}

void simu_b200_peer_close(simu_b200_ctx *ctx)
{
    if (!ctx || !ctx->peer) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    simu_b200_peer *p = ctx->peer;
    for (int r = 0; r < p->world; r++)
        if (p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    if (p->sym) cudaFree(p->sym);
    if (p->barrier) cudaFree(p->barrier);
    delete p;
    ctx->peer = nullptr;
}

}  // extern "C"
