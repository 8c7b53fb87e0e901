// C-ABI layer: context, HBM panel, pinned-host chunked staging, and the
// entry points declared in include/simu_b200.h.  No torch types, no CPU
// fallback: without a B200 every call returns SIMU_B200_ECUDA.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "ctx.cuh"

static thread_local std::string g_open_error;

int ctx_fail(simu_b200_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_open_error = buf;
    return code;
}

int ctx_cuda(simu_b200_ctx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return SIMU_B200_OK;
    return ctx_fail(ctx, e == cudaErrorMemoryAllocation ? SIMU_B200_ENOMEM : SIMU_B200_ECUDA,
                    "CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

void *ctx_scratch(simu_b200_ctx *ctx, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 256;
    if (ctx->scratch_bytes[slot] >= bytes) return ctx->scratch[slot];
    if (ctx->scratch[slot]) { cudaFree(ctx->scratch[slot]); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
    size_t want = bytes + bytes / 4;   // grow with slack
    want = (want + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
    if (e != cudaSuccess) { ctx_cuda(ctx, e, "scratch cudaMalloc"); return nullptr; }
    ctx->scratch_bytes[slot] = want;
    return ctx->scratch[slot];
}

int prof_begin(simu_b200_ctx *ctx, int id)
{
    if (!ctx->prof_on || ctx->prof_n >= simu_b200_ctx::kProfCap) return -1;
    if (!ctx->prof_alloc) {
        for (int i = 0; i < simu_b200_ctx::kProfCap; i++) {
            if (cudaEventCreate(&ctx->prof_ev[i][0]) != cudaSuccess || cudaEventCreate(&ctx->prof_ev[i][1]) != cudaSuccess)
                return -1;
        }
        ctx->prof_alloc = true;
    }
    const int slot = ctx->prof_n++;
    ctx->prof_id[slot] = id;
    cudaEventRecord(ctx->prof_ev[slot][0], ctx->stream);
    return slot;
}

void prof_end(simu_b200_ctx *ctx, int slot)
{
    if (slot >= 0) cudaEventRecord(ctx->prof_ev[slot][1], ctx->stream);
}

// ---------------------------------------------------------------- context

extern "C" {

int simu_b200_version(void) { return SIMU_B200_VERSION; }

int simu_b200_device_count(int *n)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (n) *n = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess || c == 0)
        return ctx_fail(nullptr, SIMU_B200_ECUDA, "no CUDA device available (%s); simu_b200 has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return SIMU_B200_OK;
}

const char *simu_b200_last_error(const simu_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_open_error.c_str(); }

int64_t simu_b200_row_bytes(int64_t num_samples) { return (num_samples + 3) / 4; }   // bed_header.c:219-223
int64_t simu_b200_row_stride(const simu_b200_ctx *ctx) { return ctx ? ctx->row_stride : 0; }
int64_t simu_b200_group_stride(const simu_b200_ctx *ctx)
{
    return !ctx ? 0 : ctx->layout == SIMU_B200_LAYOUT_SAMPLE16 ? ctx->s16_stride : ctx->group_stride;
}
int simu_b200_panel_layout(const simu_b200_ctx *ctx) { return ctx ? ctx->layout : SIMU_B200_LAYOUT_ROWS; }
int64_t simu_b200_panel_rows(const simu_b200_ctx *ctx) { return ctx ? ctx->panel_rows : 0; }
void *simu_b200_panel_device_ptr(const simu_b200_ctx *ctx) { return ctx ? ctx->panel : nullptr; }
int64_t simu_b200_launch_count(const simu_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int simu_b200_open(simu_b200_ctx **out, int device, int64_t num_samples)
{
    if (!out) return ctx_fail(nullptr, SIMU_B200_EINVAL, "open: ctx pointer is NULL");
    *out = nullptr;
    // one packed row (at its 128-byte pitch) must fit a staging buffer: N <= 268,434,944
    if (num_samples <= 0 || (num_samples + 3) / 4 + SIMU_ROW_ALIGN > (int64_t)SIMU_STAGE_BYTES)
        return ctx_fail(nullptr, SIMU_B200_EINVAL, "open: num_samples %lld out of range (1 .. %lld)", (long long)num_samples,
                        (long long)(4 * ((int64_t)SIMU_STAGE_BYTES - SIMU_ROW_ALIGN)));
    int n = 0;
    int rc = simu_b200_device_count(&n);
    if (rc != SIMU_B200_OK) return rc;
    if (device < 0 || device >= n)
        return ctx_fail(nullptr, SIMU_B200_EINVAL, "open: device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return ctx_cuda(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return ctx_fail(nullptr, SIMU_B200_ECUDA, "device %d (%s, sm_%d%d) is not a B200-class sm_100 GPU; "
                        "simu_b200 ships sm_100a kernels only and has no fallback", device, prop.name, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return ctx_cuda(nullptr, e, "cudaSetDevice");

    simu_b200_ctx *ctx = new (std::nothrow) simu_b200_ctx();
    if (!ctx) return ctx_fail(nullptr, SIMU_B200_ENOMEM, "open: out of host memory");
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->n_samples = num_samples;
    ctx->row_bytes = simu_b200_row_bytes(num_samples);
    ctx->row_stride = (ctx->row_bytes + SIMU_ROW_ALIGN - 1) / SIMU_ROW_ALIGN * SIMU_ROW_ALIGN;
    ctx->group_stride = (num_samples + SIMU_ROW_ALIGN - 1) / SIMU_ROW_ALIGN * SIMU_ROW_ALIGN;
    ctx->s16_stride = 4 * ctx->group_stride;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev_a)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_b)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) {
        rc = ctx_cuda(nullptr, e, "stream/event creation");
        simu_b200_close(ctx);
        return rc;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return SIMU_B200_OK;
}

void simu_b200_close(simu_b200_ctx *ctx)
{
    if (!ctx) return;
    simu_b200_peer_close(ctx);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->panel) cudaFree(ctx->panel);
    for (int i = 0; i < 32; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    for (int i = 0; i < 2; i++) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        if (i == 0 && ctx->h_eff) cudaFreeHost(ctx->h_eff);
        if (ctx->dstage[i]) cudaFree(ctx->dstage[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_repitch[i]) cudaEventDestroy(ctx->ev_repitch[i]);
        if (ctx->ev_stage[i]) cudaEventDestroy(ctx->ev_stage[i]);
    }
    if (ctx->prof_alloc)
        for (int i = 0; i < simu_b200_ctx::kProfCap; i++) { cudaEventDestroy(ctx->prof_ev[i][0]); cudaEventDestroy(ctx->prof_ev[i][1]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int simu_b200_set_stream(simu_b200_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SIMU_B200_EINVAL;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SIMU_B200_OK;
}

int simu_b200_sync(simu_b200_ctx *ctx)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SIMU_B200_OK;
}

int simu_b200_profile_enable(simu_b200_ctx *ctx, int on)
{
    if (!ctx) return SIMU_B200_EINVAL;
    ctx->prof_on = on != 0;
    ctx->prof_n = 0;
    return SIMU_B200_OK;
}

int simu_b200_profile_read(simu_b200_ctx *ctx, int kernel_id, float *total_ms, int *launches)
{
    if (!ctx || !total_ms || !launches) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float tot = 0.f;
    int n = 0;
    for (int i = 0; i < ctx->prof_n; i++) {
        if (ctx->prof_id[i] != kernel_id) continue;
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[i][0], ctx->prof_ev[i][1]));
        tot += ms;
        n++;
    }
    *total_ms = tot;
    *launches = n;
    return SIMU_B200_OK;
}

int simu_b200_missing_index(const simu_b200_ctx *ctx, int *state, int64_t *entries)
{
    if (!ctx) return SIMU_B200_EINVAL;
    if (state) *state = ctx->miss_state;
    if (entries) *entries = ctx->miss_state == 1 ? ctx->miss_total : 0;
    return SIMU_B200_OK;
}

int simu_b200_last_kernel_ms(simu_b200_ctx *ctx, int which, float *ms)
{
    if (!ctx || !ms || which < 0 || which > 1) return SIMU_B200_EINVAL;
    *ms = ctx->last_ms[which];
    return SIMU_B200_OK;
}

// ------------------------------------------------------------------ panel

int simu_b200_panel_free(simu_b200_ctx *ctx)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->panel) { CUDA_TRY(ctx, cudaFree(ctx->panel)); }
    ctx->panel = nullptr;
    ctx->panel_rows = 0;
    ctx->panel_bytes = 0;
    miss_index_invalidate(ctx);
    return SIMU_B200_OK;
}

int simu_b200_panel_alloc_layout(simu_b200_ctx *ctx, int64_t num_rows, int layout)
{
    if (!ctx) return SIMU_B200_EINVAL;
    if (num_rows <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_alloc: num_rows must be positive");
    if (layout != SIMU_B200_LAYOUT_ROWS && layout != SIMU_B200_LAYOUT_GROUP4 && layout != SIMU_B200_LAYOUT_SAMPLE16)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_alloc: unknown layout %d", layout);
    int rc = simu_b200_panel_free(ctx);
    if (rc) return rc;
    const size_t bytes = layout == SIMU_B200_LAYOUT_SAMPLE16 ? (size_t)((num_rows + 15) / 16) * (size_t)ctx->s16_stride
                         : layout == SIMU_B200_LAYOUT_GROUP4 ? (size_t)((num_rows + 3) / 4) * (size_t)ctx->group_stride
                                                             : (size_t)num_rows * (size_t)ctx->row_stride;
    cudaError_t e = cudaMalloc((void **)&ctx->panel, bytes);
    if (e != cudaSuccess)
        return ctx_fail(ctx, SIMU_B200_ENOMEM, "panel_alloc: cannot allocate %zu bytes of HBM (%s)", bytes,
                        cudaGetErrorString(e));
    ctx->panel_rows = num_rows;
    ctx->panel_bytes = bytes;
    ctx->layout = layout;
    // pitch padding is never interpreted by the kernels (they mask to N); zero it once so that
    // panel_get_rows and debuggers see defined bytes.  GROUP4: rows of a group that have not been
    // written yet read as 00 until they are.
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->panel, 0, bytes, ctx->stream));
    return SIMU_B200_OK;
}

int simu_b200_panel_alloc(simu_b200_ctx *ctx, int64_t num_rows)
{
    return simu_b200_panel_alloc_layout(ctx, num_rows, SIMU_B200_LAYOUT_ROWS);
}

// Rows per staging chunk: as many as the 64 MB bounce buffers hold, a multiple of 16 when possible so that
// GROUP4 / SAMPLE16 chunk boundaries fall on group boundaries (the interleave kernels read-modify-write
// boundary groups, so any count is correct).  simu_b200_open rejects sample counts whose single row
// exceeds a buffer.
static int64_t stage_rows(int64_t pitch)
{
    const int64_t fit = std::max<int64_t>(1, (int64_t)SIMU_STAGE_BYTES / pitch);
    return fit >= 16 ? (fit & ~(int64_t)15) : fit;
}

static int ensure_pinned(simu_b200_ctx *ctx)
{
    for (int i = 0; i < 2; i++)
        if (!ctx->pinned[i]) {
            cudaError_t e = cudaHostAlloc((void **)&ctx->pinned[i], SIMU_STAGE_BYTES, cudaHostAllocDefault);
            if (e != cudaSuccess) return ctx_cuda(ctx, e, "cudaHostAlloc(staging)");
        }
    return SIMU_B200_OK;
}

static int ensure_dstage(simu_b200_ctx *ctx)
{
    for (int i = 0; i < 2; i++) {
        if (!ctx->dstage[i]) {
            cudaError_t e = cudaMalloc((void **)&ctx->dstage[i], SIMU_STAGE_BYTES);
            if (e != cudaSuccess) return ctx_cuda(ctx, e, "cudaMalloc(device staging)");
        }
        if (!ctx->ev_copy[i]) {
            cudaError_t e = cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_repitch[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return ctx_cuda(ctx, e, "staging events");
        }
    }
    return SIMU_B200_OK;
}

// One chunk of packed rows in page-locked host memory (pitch src_stride) -> panel rows: DMA on the
// copy stream into bounce buffer `buf` (rows contiguous there: one 1-D copy when the source is
// contiguous too), then the re-pitch / interleave kernel on the compute stream.
static int stage_chunk(simu_b200_ctx *ctx, const uint8_t *pinned_src, int64_t src_stride, int64_t nr, int64_t dst_row,
                       int buf)
{
    if (ctx->dstage_used[buf]) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_repitch[buf], 0));
    if (src_stride == ctx->row_bytes)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dstage[buf], pinned_src, (size_t)(nr * ctx->row_bytes), cudaMemcpyHostToDevice,
                                      ctx->copy_stream));
    else
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->dstage[buf], (size_t)ctx->row_bytes, pinned_src, (size_t)src_stride,
                                        (size_t)ctx->row_bytes, (size_t)nr, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
    int rc = launch_repitch(ctx, ctx->dstage[buf], ctx->row_bytes, nr, dst_row, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_repitch[buf], ctx->stream));
    ctx->dstage_used[buf] = true;
    return SIMU_B200_OK;
}

static int check_rows(simu_b200_ctx *ctx, const char *who, int64_t first, int64_t n)
{
    if (!ctx->panel) return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: no panel allocated", who);
    if (n < 0 || first < 0 || first + n > ctx->panel_rows)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: rows [%lld, %lld) outside panel of %lld rows", who,
                        (long long)first, (long long)(first + n), (long long)ctx->panel_rows);
    return SIMU_B200_OK;
}

// make the copy stream wait for everything queued on the compute stream (panel memset etc.)
static int fork_copy_stream(simu_b200_ctx *ctx)
{
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
    return SIMU_B200_OK;
}

int simu_b200_panel_put_rows(simu_b200_ctx *ctx, int64_t dst_row, int64_t n_rows, const void *host_rows,
                             int64_t host_stride)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = check_rows(ctx, "panel_put_rows", dst_row, n_rows);
    if (rc) return rc;
    if (n_rows == 0) return SIMU_B200_OK;
    if (!host_rows || host_stride < ctx->row_bytes)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_put_rows: bad host buffer (stride %lld < row bytes %lld)",
                        (long long)host_stride, (long long)ctx->row_bytes);
    if ((rc = fork_copy_stream(ctx))) return rc;

    cudaPointerAttributes attr;
    bool is_pinned = (cudaPointerGetAttributes(&attr, host_rows) == cudaSuccess) && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const uint8_t *src = (const uint8_t *)host_rows;
    if ((rc = ensure_dstage(ctx))) return rc;
    const int64_t rows_per = stage_rows(ctx->row_bytes);
    int buf = 0;
    if (is_pinned) {
        // page-locked: chunked DMA straight from the caller's buffer into the device bounce
        // buffers, re-pitched / interleaved on the compute stream while the next chunk copies
        for (int64_t r = 0; r < n_rows; r += rows_per, buf ^= 1)
            if ((rc = stage_chunk(ctx, src + r * host_stride, host_stride, std::min(rows_per, n_rows - r), dst_row + r, buf)))
                return rc;
        return SIMU_B200_OK;      // the compute stream is already ordered behind the last kernel
    }
    if ((rc = ensure_pinned(ctx))) return rc;
    for (int64_t r = 0; r < n_rows; r += rows_per, buf ^= 1) {
        const int64_t nr = std::min(rows_per, n_rows - r);
        if (ctx->dstage_used[buf]) CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_copy[buf]));   // DMA out of pinned[buf] done
        if (host_stride == ctx->row_bytes) {
            memcpy(ctx->pinned[buf], src + r * host_stride, (size_t)(nr * ctx->row_bytes));
        } else {
            for (int64_t q = 0; q < nr; q++)
                memcpy(ctx->pinned[buf] + q * ctx->row_bytes, src + (r + q) * host_stride, (size_t)ctx->row_bytes);
        }
        if ((rc = stage_chunk(ctx, ctx->pinned[buf], ctx->row_bytes, nr, dst_row + r, buf))) return rc;
    }
    // the caller may free or reuse host_rows on return; pinned[] is ours, nothing to wait for
    return SIMU_B200_OK;
}

int simu_b200_panel_get_rows(simu_b200_ctx *ctx, int64_t src_row, int64_t n_rows, void *host_rows,
                             int64_t host_stride)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = check_rows(ctx, "panel_get_rows", src_row, n_rows);
    if (rc) return rc;
    if (n_rows == 0) return SIMU_B200_OK;
    if (!host_rows || host_stride < ctx->row_bytes)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_get_rows: bad host buffer");
    if ((rc = ensure_dstage(ctx))) return rc;
    // gather the rows into a bounce buffer (contiguous, pitch row_bytes) and copy that out
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    const int64_t rows_per = std::max<int64_t>(1, (int64_t)SIMU_STAGE_BYTES / ctx->row_bytes);
    for (int64_t r = 0; r < n_rows; r += rows_per) {
        const int64_t nr = std::min(rows_per, n_rows - r);
        if ((rc = launch_unpanel(ctx, src_row + r, nr, ctx->dstage[0], ctx->stream))) return rc;
        CUDA_TRY(ctx, cudaMemcpy2DAsync((uint8_t *)host_rows + r * host_stride, (size_t)host_stride, ctx->dstage[0],
                                        (size_t)ctx->row_bytes, (size_t)ctx->row_bytes, (size_t)nr,
                                        cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SIMU_B200_OK;
}

int simu_b200_panel_load_bed(simu_b200_ctx *ctx, const char *bed_path, int64_t num_loci, const int64_t *file_rows,
                             int64_t n_rows, int64_t dst_row)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = check_rows(ctx, "panel_load_bed", dst_row, n_rows);
    if (rc) return rc;
    if (!bed_path) return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_load_bed: path is NULL");
    int fd = open(bed_path, O_RDONLY);
    if (fd < 0) return ctx_fail(ctx, SIMU_B200_ERROR, "ERROR: Could not open %s (%s)", bed_path, strerror(errno));
    unsigned char hdr[3] = {0, 0, 0};
    if (pread(fd, hdr, 3, 0) != 3) {                       // parse_header, bed.c:52-58
        close(fd);
        return ctx_fail(ctx, SIMU_B200_ERROR, "ERROR: Could not open %s (file shorter than a BED header)", bed_path);
    }
    // bed_header.c:80-102 (get_version_and_order) and :56-71 (get_data_offset):
    //   v1.00: magic 6c 1b, third byte = order (01: one locus per row), data at offset 3;
    //   v0.99: first byte 00 or 01 = order, data at offset 1;
    //   older: one sample per row, data at offset 0.
    // simu runs only files with one locus per row (source.cc:196-200).
    int64_t data_off = -1;
    if (hdr[0] == 0x6c && hdr[1] == 0x1b) { if (hdr[2] == 0x01) data_off = 3; }
    else if ((hdr[0] & ~0x01) == 0) { if (hdr[0] == 0x01) data_off = 1; }
    if (data_off < 0) {
        close(fd);
        return ctx_fail(ctx, SIMU_B200_EFORMAT,
                        "ERROR: Unsupported format in %s, snps must be rows, samples must be columns", bed_path);
    }
    if (n_rows == 0) { close(fd); return SIMU_B200_OK; }
    if ((rc = ensure_pinned(ctx)) || (rc = ensure_dstage(ctx)) || (rc = fork_copy_stream(ctx))) { close(fd); return rc; }

    const int64_t B = ctx->row_bytes;
    const int64_t cap_rows = stage_rows(B);
    int io_threads = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char *e = getenv("SIMU_B200_IO_THREADS")) io_threads = std::max(1, atoi(e));
    int buf = 0;
    int64_t done = 0, prev = -1;
    while (done < n_rows) {
        const int64_t nr = std::min(cap_rows, n_rows - done);
        if (ctx->dstage_used[buf]) {
            cudaError_t e = cudaEventSynchronize(ctx->ev_copy[buf]);      // DMA out of pinned[buf] done
            if (e != cudaSuccess) { close(fd); return ctx_cuda(ctx, e, "staging event sync"); }
        }
        // coalesce runs of adjacent causal rows into single preads (pio_skip_row defers the
        // seek the same way, bed.c:333-341); long runs are cut into <= 4 MB pieces so that the
        // reader threads stay balanced
        struct Piece { int64_t file_off, bytes, dst_off; };
        std::vector<Piece> pieces;
        int64_t q = 0;
        while (q < nr) {
            const int64_t r0 = file_rows ? file_rows[done + q] : done + q;
            if (r0 <= prev || r0 >= num_loci) {
                close(fd);
                if (r0 >= num_loci)                     // source.cc:855-858
                    return ctx_fail(ctx, SIMU_B200_END, "ERROR: expect to find %lld variants across input files; "
                                    "row %lld requested. Are input Plink files corrupted?", (long long)num_loci, (long long)r0);
                return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_load_bed: file_rows must be strictly ascending");
            }
            int64_t run = 1;
            while (q + run < nr && (file_rows ? file_rows[done + q + run] : r0 + run) == r0 + run) run++;
            const int64_t total = run * B, cut = 4 << 20;
            for (int64_t o = 0; o < total; o += cut)
                pieces.push_back({data_off + r0 * B + o, std::min(cut, total - o), q * B + o});
            prev = r0 + run - 1;
            q += run;
        }
        std::atomic<int> failed{0};
        uint8_t *dst_base = ctx->pinned[buf];
        auto reader = [&](int t) {
            for (size_t i = (size_t)t; i < pieces.size() && !failed.load(std::memory_order_relaxed); i += (size_t)io_threads) {
                int64_t got = 0;
                while (got < pieces[i].bytes) {
                    ssize_t k = pread(fd, dst_base + pieces[i].dst_off + got, (size_t)(pieces[i].bytes - got),
                                      (off_t)(pieces[i].file_off + got));
                    if (k <= 0) { failed.store(1); return; }     // short read: bed.c:349-352
                    got += k;
                }
            }
        };
        if (io_threads <= 1 || pieces.size() < 2) {
            reader(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 1; t < io_threads; t++) pool.emplace_back(reader, t);
            reader(0);
            for (auto &th : pool) th.join();
        }
        if (failed.load()) {                            // -> source.cc:226-229
            close(fd);
            return ctx_fail(ctx, SIMU_B200_ERROR, "ERROR: Error reading from %s", bed_path);
        }
        if ((rc = stage_chunk(ctx, ctx->pinned[buf], B, nr, dst_row + done, buf))) { close(fd); return rc; }
        buf ^= 1;
        done += nr;
    }
    close(fd);
    // the pinned buffers are reused by the next call: drain the DMA queue before returning
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return SIMU_B200_OK;
}

// rows first_row + r, or row_ids[r] (host list), r < n_rows -> panel rows dst_row + r
static int synth_into_panel(simu_b200_ctx *ctx, uint64_t seed, int64_t first_row, const int64_t *row_ids, int64_t n_rows,
                            int64_t dst_row)
{
    int64_t *d_ids = nullptr;
    if (row_ids) {
        d_ids = (int64_t *)ctx_scratch(ctx, SCR_MISC, (size_t)n_rows * 8);
        if (!d_ids) return SIMU_B200_ENOMEM;
        CUDA_TRY(ctx, cudaMemcpyAsync(d_ids, row_ids, (size_t)n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // row_ids may be pageable
    }
    if (ctx->layout == SIMU_B200_LAYOUT_ROWS)
        return launch_synth(ctx, seed, first_row, d_ids, n_rows, ctx->panel + dst_row * ctx->row_stride, ctx->row_stride);
    // GROUP4: generate packed rows chunk by chunk into a bounce buffer, then interleave
    int rc = ensure_dstage(ctx);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    const int64_t pitch = ctx->row_stride;
    const int64_t rows_per = stage_rows(pitch);
    int64_t r = 0;
    while (r < n_rows) {
        const int64_t nr = std::min(rows_per, n_rows - r);
        if ((rc = launch_synth(ctx, seed, first_row + r, d_ids ? d_ids + r : nullptr, nr, ctx->dstage[0], pitch))) return rc;
        if ((rc = launch_repitch(ctx, ctx->dstage[0], pitch, nr, dst_row + r, ctx->stream))) return rc;
        r += nr;
    }
    return SIMU_B200_OK;
}

int simu_b200_panel_synth(simu_b200_ctx *ctx, uint64_t panel_seed, int64_t first_row, int64_t n_rows,
                          int64_t dst_row)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = check_rows(ctx, "panel_synth", dst_row, n_rows);
    if (rc) return rc;
    return synth_into_panel(ctx, panel_seed, first_row, nullptr, n_rows, dst_row);
}

int simu_b200_panel_synth_rows(simu_b200_ctx *ctx, uint64_t panel_seed, const int64_t *row_ids, int64_t n_rows,
                               int64_t dst_row)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = check_rows(ctx, "panel_synth_rows", dst_row, n_rows);
    if (rc) return rc;
    if (n_rows > 0 && !row_ids) return ctx_fail(ctx, SIMU_B200_EINVAL, "panel_synth_rows: row_ids is NULL");
    return synth_into_panel(ctx, panel_seed, 0, row_ids, n_rows, dst_row);
}

// ------------------------------------------------ device-resident variants

static int check_hot(simu_b200_ctx *ctx, const char *who, int64_t mc, const void *rows)
{
    if (!ctx) return SIMU_B200_EINVAL;
    if (!ctx->panel) return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: no panel allocated", who);
    if (mc < 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: negative causal count", who);
    if (rows && ctx->layout != SIMU_B200_LAYOUT_ROWS)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: a %s panel is compact (stage the causal rows only, in causal "
                        "order); row lists need a ROWS panel", who,
                        ctx->layout == SIMU_B200_LAYOUT_GROUP4 ? "GROUP4" : "SAMPLE16");
    cudaError_t e = cudaSetDevice(ctx->device);
    return ctx_cuda(ctx, e, "cudaSetDevice");
}

int simu_b200_find_freq_dev(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, int32_t *d_counts,
                            double *d_freq)
{
    int rc = check_hot(ctx, "find_freq", mc, d_rows);
    if (rc) return rc;
    if (!d_rows && mc > ctx->panel_rows)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq: %lld causal rows but panel holds %lld", (long long)mc,
                        (long long)ctx->panel_rows);
    return launch_counts(ctx, d_rows, mc, d_counts, d_freq);
}

static int find_pheno_dev_impl(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                               const double *d_effect1, const double *d_effect2, double *d_pheno1,
                               double *d_pheno2, int mode, int accumulate)
{
    int rc = check_hot(ctx, "find_pheno", mc, d_rows);
    if (rc) return rc;
    if (!d_rows && mc > ctx->panel_rows)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno: %lld causal rows but panel holds %lld", (long long)mc,
                        (long long)ctx->panel_rows);
    if (!d_freq || !d_effect1 || !d_pheno1 || ((d_effect2 == nullptr) != (d_pheno2 == nullptr)))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno: NULL argument (effect2/pheno2 must be both set or both NULL)");
    if (ctx->layout == SIMU_B200_LAYOUT_SAMPLE16) {
        if (mode == SIMU_B200_MODE_EXACT) {
            if (accumulate && mc == 0) return SIMU_B200_OK;
            return launch_pheno_exact_s16(ctx, mc, d_freq, d_effect1, d_effect2, d_pheno1, d_pheno2, accumulate);
        }
        if (mode == SIMU_B200_MODE_FAST) return launch_pheno_tc(ctx, mc, d_freq, d_effect1, d_effect2, d_pheno1, d_pheno2, accumulate);
    }
    if (accumulate)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno_accumulate: needs a SAMPLE16 panel");
    if (mode == SIMU_B200_MODE_EXACT)
        return launch_pheno_exact(ctx, d_rows, mc, d_freq, d_effect1, d_effect2, d_pheno1, d_pheno2);
    if (mode == SIMU_B200_MODE_FAST)
        return launch_pheno_lut(ctx, d_rows, mc, d_freq, d_effect1, d_effect2, d_pheno1, d_pheno2);
    return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno: unknown mode %d", mode);
}

int simu_b200_find_pheno_dev(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                             const double *d_effect1, const double *d_effect2, double *d_pheno1,
                             double *d_pheno2, int mode)
{
    return find_pheno_dev_impl(ctx, d_rows, mc, d_freq, d_effect1, d_effect2, d_pheno1, d_pheno2, mode, 0);
}

int simu_b200_mem_info(simu_b200_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f + (ctx->panel ? (int64_t)ctx->panel_bytes : 0);    // the current panel would be replaced
    if (total_bytes) *total_bytes = (int64_t)t;
    return SIMU_B200_OK;
}

int64_t simu_b200_panel_bytes(const simu_b200_ctx *ctx, int64_t num_rows, int layout)
{
    if (!ctx || num_rows < 0) return -1;
    if (layout == SIMU_B200_LAYOUT_SAMPLE16) return (num_rows + 15) / 16 * ctx->s16_stride;
    if (layout == SIMU_B200_LAYOUT_GROUP4) return (num_rows + 3) / 4 * ctx->group_stride;
    if (layout == SIMU_B200_LAYOUT_ROWS) return num_rows * ctx->row_stride;
    return -1;
}

int simu_b200_sumsq_dev(simu_b200_ctx *ctx, const double *d_y, int64_t n, double *d_out)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return launch_sumsq(ctx, d_y, n, d_out);
}

int simu_b200_apply_heritability_dev(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise, double *d_coef)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!d_y || !d_noise || n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "apply_heritability_dev: bad arguments");
    return launch_apply_heritability(ctx, hsq, d_y, n, d_noise, d_coef);
}

int simu_b200_heritability_dev(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise,
                               const double *d_sumsq, double *d_coef)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return launch_heritability(ctx, hsq, d_y, n, d_noise, d_sumsq, d_coef);
}

int simu_b200_liability_split_dev(simu_b200_ctx *ctx, const double *d_pheno, int64_t n, int64_t max_ncon, uint8_t *d_is_case)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!d_pheno || !d_is_case || max_ncon < 0 || max_ncon > n)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "liability_split: bad arguments");
    return launch_liability_split(ctx, d_pheno, n, max_ncon, d_is_case);
}

int simu_b200_liability_order_dev(simu_b200_ctx *ctx, const double *d_pheno, int64_t n, int32_t *d_index)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return launch_argsort_f64(ctx, d_pheno, n, d_index);
}

// --------------------------------------------------- host-buffer entry points

static int h2d(simu_b200_ctx *ctx, int slot, const void *src, size_t bytes, void **dst)
{
    void *d = ctx_scratch(ctx, slot, bytes);
    if (!d) return SIMU_B200_ENOMEM;
    CUDA_TRY(ctx, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dst = d;
    return SIMU_B200_OK;
}

static int check_row_list(simu_b200_ctx *ctx, const char *who, const int32_t *rows, int64_t mc)
{
    if (!rows) {
        if (mc > ctx->panel_rows)
            return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: %lld causal rows but panel holds %lld", who, (long long)mc,
                            (long long)ctx->panel_rows);
        return SIMU_B200_OK;
    }
    for (int64_t j = 0; j < mc; j++)
        if (rows[j] < 0 || rows[j] >= ctx->panel_rows)
            return ctx_fail(ctx, SIMU_B200_EINVAL, "%s: rows[%lld] = %d outside panel of %lld rows", who,
                            (long long)j, rows[j], (long long)ctx->panel_rows);
    return SIMU_B200_OK;
}

int simu_b200_find_freq(simu_b200_ctx *ctx, const int32_t *rows, int64_t mc, int32_t *counts, double *freq)
{
    int rc = check_hot(ctx, "find_freq", mc, rows);
    if (rc || (rc = check_row_list(ctx, "find_freq", rows, mc))) return rc;
    if (mc == 0) return SIMU_B200_OK;
    void *d_rows = nullptr;
    if (rows && (rc = h2d(ctx, SCR_ROWS, rows, (size_t)mc * 4, &d_rows))) return rc;
    int32_t *d_counts = (int32_t *)ctx_scratch(ctx, SCR_COUNTS, (size_t)mc * 16);
    double *d_freq = (double *)ctx_scratch(ctx, SCR_FREQ, (size_t)mc * 8);
    if (!d_counts || !d_freq) return SIMU_B200_ENOMEM;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    if ((rc = launch_counts(ctx, (const int32_t *)d_rows, mc, d_counts, d_freq))) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    if (counts) CUDA_TRY(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)mc * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (freq) CUDA_TRY(ctx, cudaMemcpyAsync(freq, d_freq, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev_a, ctx->ev_b));
    return SIMU_B200_OK;
}

static int find_pheno_host_impl(simu_b200_ctx *ctx, const int32_t *rows, int64_t mc, const double *freq,
                                const double *effect1, const double *effect2, double *pheno1, double *pheno2, int mode,
                                int accumulate)
{
    int rc = check_hot(ctx, "find_pheno", mc, rows);
    if (rc || (rc = check_row_list(ctx, "find_pheno", rows, mc))) return rc;
    if (!pheno1 || ((effect2 == nullptr) != (pheno2 == nullptr)) || (mc > 0 && (!freq || !effect1)))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno: NULL argument (effect2/pheno2 must be both set or both NULL)");
    if (mode != SIMU_B200_MODE_EXACT && mode != SIMU_B200_MODE_FAST)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno: unknown mode %d", mode);
    if (accumulate && ctx->layout != SIMU_B200_LAYOUT_SAMPLE16)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_pheno_accumulate: needs a SAMPLE16 panel");
    const int64_t n = ctx->n_samples;
    if (mc == 0) {                                           // source.cc:962-966: phenotypes start at 0
        if (!accumulate) for (int64_t i = 0; i < n; i++) { pheno1[i] = 0; if (pheno2) pheno2[i] = 0; }
        return SIMU_B200_OK;
    }
    void *d_rows = nullptr, *d_freq = nullptr, *d_x1 = nullptr, *d_x2 = nullptr;
    if (rows && (rc = h2d(ctx, SCR_ROWS, rows, (size_t)mc * 4, &d_rows))) return rc;
    if ((rc = h2d(ctx, SCR_FREQ, freq, (size_t)mc * 8, &d_freq))) return rc;
    if ((rc = h2d(ctx, SCR_X1, effect1, (size_t)mc * 8, &d_x1))) return rc;
    if (effect2 && (rc = h2d(ctx, SCR_X2, effect2, (size_t)mc * 8, &d_x2))) return rc;
    double *d_y1 = (double *)ctx_scratch(ctx, SCR_Y1, (size_t)n * 8);
    double *d_y2 = effect2 ? (double *)ctx_scratch(ctx, SCR_Y2, (size_t)n * 8) : nullptr;
    if (!d_y1 || (effect2 && !d_y2)) return SIMU_B200_ENOMEM;
    if (accumulate) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_y1, pheno1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (pheno2) CUDA_TRY(ctx, cudaMemcpyAsync(d_y2, pheno2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    ctx->h_x1 = effect1; ctx->h_x2 = effect2;              // EXACT on SAMPLE16 passes them as kernel parameters
    rc = find_pheno_dev_impl(ctx, (const int32_t *)d_rows, mc, (const double *)d_freq, (const double *)d_x1,
                             (const double *)d_x2, d_y1, d_y2, mode, accumulate);
    ctx->h_x1 = ctx->h_x2 = nullptr;
    if (rc) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(pheno1, d_y1, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pheno2) CUDA_TRY(ctx, cudaMemcpyAsync(pheno2, d_y2, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev_a, ctx->ev_b));
    return SIMU_B200_OK;
}

int simu_b200_find_pheno(simu_b200_ctx *ctx, const int32_t *rows, int64_t mc, const double *freq,
                         const double *effect1, const double *effect2, double *pheno1, double *pheno2, int mode)
{
    return find_pheno_host_impl(ctx, rows, mc, freq, effect1, effect2, pheno1, pheno2, mode, 0);
}

int simu_b200_find_pheno_accumulate(simu_b200_ctx *ctx, int64_t mc, const double *freq, const double *effect1,
                                    const double *effect2, double *pheno1, double *pheno2, int mode)
{
    return find_pheno_host_impl(ctx, nullptr, mc, freq, effect1, effect2, pheno1, pheno2, mode, 1);
}

int simu_b200_find_freq_pheno_dev(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1, const double *d_raw2,
                                  int scale_op, const int32_t *d_component, const double *d_s_pow, int num_components,
                                  int32_t *d_counts, double *d_freq, double *d_effect1, double *d_effect2,
                                  double *d_pheno1, double *d_pheno2, uint64_t *d_bad_index)
{
    int rc = check_hot(ctx, "find_freq_pheno", mc, nullptr);
    if (rc) return rc;
    if (ctx->layout == SIMU_B200_LAYOUT_ROWS)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: needs a compact panel (GROUP4 or SAMPLE16)");
    if (mc > ctx->panel_rows)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: %lld causal rows but panel holds %lld", (long long)mc,
                        (long long)ctx->panel_rows);
    if (!d_raw1 || !d_pheno1 || !d_bad_index || ((d_raw2 == nullptr) != (d_pheno2 == nullptr)) || scale_op < -1 ||
        scale_op > 1 || num_components < 1 || (scale_op == 0 && !d_s_pow))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: bad arguments");
    CUDA_TRY(ctx, cudaMemsetAsync(d_bad_index, 0xFF, 8, ctx->stream));
    if (mc == 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_pheno1, 0, (size_t)ctx->n_samples * 8, ctx->stream));
        if (d_pheno2) CUDA_TRY(ctx, cudaMemsetAsync(d_pheno2, 0, (size_t)ctx->n_samples * 8, ctx->stream));
        return SIMU_B200_OK;
    }
    // counts -> effect scaling (the reference's host work between the passes, source.cc:1245-1262) -> FAST
    // phenotypes, queued on one stream without a host round trip.  (A single-kernel version that read HBM once
    // was built and measured in round 1: 2.4 ms against 1.06 ms for this sequence on config 2 -- the scaled
    // effects of a SNP depend on its frequency over ALL samples -- and has been removed.)
    if (!d_freq) d_freq = (double *)ctx_scratch(ctx, SCR_FREQ, (size_t)mc * 8);
    if (!d_effect1) d_effect1 = (double *)ctx_scratch(ctx, SCR_X1, (size_t)mc * 8);
    if (d_raw2 && !d_effect2) d_effect2 = (double *)ctx_scratch(ctx, SCR_X2, (size_t)mc * 8);
    if (!d_freq || !d_effect1 || (d_raw2 && !d_effect2)) return SIMU_B200_ENOMEM;
    if ((rc = launch_counts(ctx, nullptr, mc, d_counts, d_freq))) return rc;
    if (d_effect1 != d_raw1) CUDA_TRY(ctx, cudaMemcpyAsync(d_effect1, d_raw1, (size_t)mc * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_raw2 && d_effect2 != d_raw2)
        CUDA_TRY(ctx, cudaMemcpyAsync(d_effect2, d_raw2, (size_t)mc * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (scale_op >= 0 && (rc = launch_scale_effects(ctx, d_freq, d_effect1, d_raw2 ? d_effect2 : nullptr, mc, d_component, d_s_pow,
                                                    num_components, scale_op, (unsigned long long *)d_bad_index)))
        return rc;
    return simu_b200_find_pheno_dev(ctx, nullptr, mc, d_freq, d_effect1, d_raw2 ? d_effect2 : nullptr, d_pheno1, d_pheno2,
                                    SIMU_B200_MODE_FAST);
}

int simu_b200_find_freq_pheno(simu_b200_ctx *ctx, int64_t mc, const double *raw1, const double *raw2, int scale_op,
                              const int32_t *component, const double *s_pow1, const double *s_pow2, int num_components,
                              int32_t *counts, double *freq, double *effect1, double *effect2, double *pheno1,
                              double *pheno2, int64_t *bad_index)
{
    int rc = check_hot(ctx, "find_freq_pheno", mc, nullptr);
    if (rc) return rc;
    if (bad_index) *bad_index = -1;
    if (!pheno1 || ((raw2 == nullptr) != (pheno2 == nullptr)) || (mc > 0 && !raw1) || num_components < 1 ||
        (scale_op == 0 && (!s_pow1 || (raw2 && !s_pow2))))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: bad arguments");
    const int64_t n = ctx->n_samples;
    if (mc == 0) {
        for (int64_t i = 0; i < n; i++) { pheno1[i] = 0; if (pheno2) pheno2[i] = 0; }
        return SIMU_B200_OK;
    }
    if (component)
        for (int64_t j = 0; j < mc; j++)
            if (component[j] < 0 || component[j] >= num_components)
                return ctx_fail(ctx, SIMU_B200_EINVAL, "find_freq_pheno: component[%lld] = %d out of range", (long long)j, component[j]);
    void *d_r1 = nullptr, *d_r2 = nullptr, *d_comp = nullptr;
    if ((rc = h2d(ctx, SCR_RAW1, raw1, (size_t)mc * 8, &d_r1))) return rc;
    if (raw2 && (rc = h2d(ctx, SCR_RAW2, raw2, (size_t)mc * 8, &d_r2))) return rc;
    if (component && (rc = h2d(ctx, SCR_ROWS, component, (size_t)mc * 4, &d_comp))) return rc;
    std::vector<double> sp(2 * (size_t)num_components + 1, 0.0);
    for (int c = 0; c < num_components; c++) {
        if (s_pow1) sp[c] = s_pow1[c];
        if (s_pow2) sp[num_components + c] = s_pow2[c];
    }
    double *d_sp = (double *)ctx_scratch(ctx, SCR_SPOW, sp.size() * 8);
    int32_t *d_counts = (int32_t *)ctx_scratch(ctx, SCR_COUNTS, (size_t)mc * 16);
    double *d_freq = (double *)ctx_scratch(ctx, SCR_FREQ, (size_t)mc * 8);
    double *d_x1 = (double *)ctx_scratch(ctx, SCR_X1, (size_t)mc * 8);
    double *d_x2 = raw2 ? (double *)ctx_scratch(ctx, SCR_X2, (size_t)mc * 8) : nullptr;
    double *d_y1 = (double *)ctx_scratch(ctx, SCR_Y1, (size_t)n * 8);
    double *d_y2 = raw2 ? (double *)ctx_scratch(ctx, SCR_Y2, (size_t)n * 8) : nullptr;
    if (!d_sp || !d_counts || !d_freq || !d_x1 || (raw2 && !d_x2) || !d_y1 || (raw2 && !d_y2)) return SIMU_B200_ENOMEM;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_sp, sp.data(), sp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    rc = simu_b200_find_freq_pheno_dev(ctx, mc, (const double *)d_r1, (const double *)d_r2, scale_op, (const int32_t *)d_comp,
                                       d_sp, num_components, d_counts, d_freq, d_x1, d_x2, d_y1, d_y2,
                                       (uint64_t *)(d_sp + 2 * (size_t)num_components));
    if (rc) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    unsigned long long bad = ~0ull;
    CUDA_TRY(ctx, cudaMemcpyAsync(&bad, d_sp + 2 * (size_t)num_components, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) CUDA_TRY(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)mc * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (freq) CUDA_TRY(ctx, cudaMemcpyAsync(freq, d_freq, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (effect1) CUDA_TRY(ctx, cudaMemcpyAsync(effect1, d_x1, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (effect2 && d_x2) CUDA_TRY(ctx, cudaMemcpyAsync(effect2, d_x2, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(pheno1, d_y1, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pheno2) CUDA_TRY(ctx, cudaMemcpyAsync(pheno2, d_y2, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev_a, ctx->ev_b));
    if (bad != ~0ull) {
        if (bad_index) *bad_index = (int64_t)bad;
        return ctx_fail(ctx, SIMU_B200_ERROR, "find_freq_pheno: causal variant %llu has zero frequency", bad);
    }
    return SIMU_B200_OK;
}

int simu_b200_apply_heritability(simu_b200_ctx *ctx, float hsq, double *pheno, int64_t n, const double *noise,
                                 double *effect, int64_t m, double *gen_coef, double *env_coef)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!pheno || !noise || n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "apply_heritability: bad arguments");
    void *d_y = nullptr, *d_noise = nullptr;
    int rc;
    if ((rc = h2d(ctx, SCR_Y1, pheno, (size_t)n * 8, &d_y))) return rc;
    if ((rc = h2d(ctx, SCR_NOISE, noise, (size_t)n * 8, &d_noise))) return rc;
    double *d_misc = (double *)ctx_scratch(ctx, SCR_COUNTS, 64);
    if (!d_misc) return SIMU_B200_ENOMEM;
    if ((rc = launch_apply_heritability(ctx, hsq, (double *)d_y, n, (const double *)d_noise, d_misc + 1))) return rc;
    double coef[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(pheno, d_y, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(coef, d_misc + 1, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (effect) for (int64_t i = 0; i < m; i++) effect[i] *= coef[0];   // source.cc:1028
    if (gen_coef) *gen_coef = coef[0];
    if (env_coef) *env_coef = coef[1];
    return SIMU_B200_OK;
}

int simu_b200_scale_effects(simu_b200_ctx *ctx, int op, const double *freq, double *effect1, double *effect2, int64_t mc,
                            const int32_t *component, const double *s_pow1, const double *s_pow2, int num_components,
                            int64_t *bad_index)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (bad_index) *bad_index = -1;
    if (op < 0 || op > 2 || mc < 0 || (mc > 0 && (!freq || !effect1)) || num_components < 1 ||
        (op == 0 && (!s_pow1 || (effect2 && !s_pow2))))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "scale_effects: bad arguments");
    if (mc == 0) return SIMU_B200_OK;
    if (component)
        for (int64_t j = 0; j < mc; j++)
            if (component[j] < 0 || component[j] >= num_components)
                return ctx_fail(ctx, SIMU_B200_EINVAL, "scale_effects: component[%lld] = %d out of range", (long long)j, component[j]);
    void *d_freq = nullptr, *d_x1 = nullptr, *d_x2 = nullptr, *d_comp = nullptr;
    int rc;
    if ((rc = h2d(ctx, SCR_FREQ, freq, (size_t)mc * 8, &d_freq))) return rc;
    if ((rc = h2d(ctx, SCR_X1, effect1, (size_t)mc * 8, &d_x1))) return rc;
    if (effect2 && (rc = h2d(ctx, SCR_X2, effect2, (size_t)mc * 8, &d_x2))) return rc;
    if (component && (rc = h2d(ctx, SCR_ROWS, component, (size_t)mc * 4, &d_comp))) return rc;
    // [2][ncomp] exponents followed by the "first bad variant" word
    std::vector<double> sp(2 * (size_t)num_components + 1, 0.0);
    for (int c = 0; c < num_components; c++) {
        if (s_pow1) sp[c] = s_pow1[c];
        if (s_pow2) sp[num_components + c] = s_pow2[c];
    }
    const unsigned long long none = ~0ull;
    memcpy(&sp[2 * (size_t)num_components], &none, 8);
    double *d_misc = (double *)ctx_scratch(ctx, SCR_MISC, sp.size() * 8);
    if (!d_misc) return SIMU_B200_ENOMEM;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_misc, sp.data(), sp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long *d_bad = (unsigned long long *)(d_misc + 2 * (size_t)num_components);
    if ((rc = launch_scale_effects(ctx, (const double *)d_freq, (double *)d_x1, (double *)d_x2, mc, (const int32_t *)d_comp,
                                   d_misc, num_components, op, d_bad)))
        return rc;
    unsigned long long bad = none;
    CUDA_TRY(ctx, cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(effect1, d_x1, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (effect2) CUDA_TRY(ctx, cudaMemcpyAsync(effect2, d_x2, (size_t)mc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad != none) {
        if (bad_index) *bad_index = (int64_t)bad;
        return ctx_fail(ctx, SIMU_B200_ERROR, "scale_effects: causal variant %llu has zero frequency", bad);
    }
    return SIMU_B200_OK;
}

int simu_b200_scale_effects_dev(simu_b200_ctx *ctx, int op, const double *d_freq, double *d_effect1, double *d_effect2,
                                int64_t mc, const int32_t *d_component, const double *d_s_pow, int num_components,
                                uint64_t *d_bad_index)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (op < 0 || op > 2 || mc < 0 || !d_freq || !d_effect1 || !d_bad_index || num_components < 1 || (op == 0 && !d_s_pow))
        return ctx_fail(ctx, SIMU_B200_EINVAL, "scale_effects_dev: bad arguments");
    return launch_scale_effects(ctx, d_freq, d_effect1, d_effect2, mc, d_component, d_s_pow, num_components, op,
                                (unsigned long long *)d_bad_index);
}

int simu_b200_liability_split(simu_b200_ctx *ctx, const double *pheno, int64_t n, int64_t max_ncon, uint8_t *is_case)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!pheno || !is_case || n <= 0 || max_ncon < 0 || max_ncon > n)
        return ctx_fail(ctx, SIMU_B200_EINVAL, "liability_split: bad arguments");
    void *d_y = nullptr;
    int rc;
    if ((rc = h2d(ctx, SCR_Y1, pheno, (size_t)n * 8, &d_y))) return rc;
    uint8_t *d_flag = (uint8_t *)ctx_scratch(ctx, SCR_COUNTS, (size_t)n);
    if (!d_flag) return SIMU_B200_ENOMEM;
    if ((rc = launch_liability_split(ctx, (const double *)d_y, n, max_ncon, d_flag))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(is_case, d_flag, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SIMU_B200_OK;
}

int simu_b200_liability_order(simu_b200_ctx *ctx, const double *pheno, int64_t n, int32_t *index)
{
    if (!ctx) return SIMU_B200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!pheno || !index || n <= 0) return ctx_fail(ctx, SIMU_B200_EINVAL, "liability_order: bad arguments");
    void *d_y = nullptr;
    int rc;
    if ((rc = h2d(ctx, SCR_Y1, pheno, (size_t)n * 8, &d_y))) return rc;
    int32_t *d_idx = (int32_t *)ctx_scratch(ctx, SCR_COUNTS, (size_t)n * 4);
    if (!d_idx) return SIMU_B200_ENOMEM;
    if ((rc = launch_argsort_f64(ctx, (const double *)d_y, n, d_idx))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(index, d_idx, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SIMU_B200_OK;
}

}  // extern "C"
