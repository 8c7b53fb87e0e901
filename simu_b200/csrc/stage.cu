// Staging helper: packed rows arrive from the host as ONE contiguous 1-D DMA per chunk
// (55 GB/s over PCIe 5 x16, against 52 GB/s for a pitched 2-D copy of 25 KB rows) into a
// device bounce buffer; this kernel then re-bases them to the panel's 128-byte row pitch.
// It is the device half of what bed_read_row's fread does in the reference
// (/root/reference/libplinkio/src/bed.c:343-346): the .bed's 3-byte header and the
// ceil(N/4) row length make file offsets unaligned, the panel rows are aligned.
#include "ctx.cuh"

namespace {

__global__ void __launch_bounds__(256)
repitch_kernel(const uint8_t *__restrict__ src, int64_t row_bytes, uint8_t *__restrict__ dst, int64_t row_stride,
               int64_t n_rows)
{
    const int64_t words = (row_bytes + 3) >> 2;           // destination 32-bit words holding data
    const int64_t total = n_rows * words;
    const bool aligned = ((row_bytes & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / words, w = t - r * words;
        const uint8_t *s = src + r * row_bytes + 4 * w;
        uint32_t v;
        if (aligned) {
            v = *reinterpret_cast<const uint32_t *>(s);
        } else {
            const int64_t left = row_bytes - 4 * w;       // 1..4 valid bytes in this word
            v = s[0];
            if (left > 1) v |= (uint32_t)s[1] << 8;
            if (left > 2) v |= (uint32_t)s[2] << 16;
            if (left > 3) v |= (uint32_t)s[3] << 24;
        }
        reinterpret_cast<uint32_t *>(dst + r * row_stride)[w] = v;
    }
}

}  // namespace

int launch_repitch(simu_b200_ctx *ctx, const uint8_t *d_src, int64_t n_rows, int64_t dst_row, cudaStream_t stream)
{
    if (n_rows <= 0) return SIMU_B200_OK;
    const int64_t total = n_rows * ((ctx->row_bytes + 3) >> 2);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    repitch_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_src, ctx->row_bytes, ctx->panel + dst_row * ctx->row_stride,
                                                         ctx->row_stride, n_rows);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "repitch_kernel launch");
}
