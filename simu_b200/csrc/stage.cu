// Staging helpers: packed rows arrive from the host as ONE contiguous 1-D DMA per chunk
// (55 GB/s over PCIe 5 x16, against 52 GB/s for a pitched 2-D copy of 25 KB rows) into a
// device bounce buffer; these kernels then move them into the panel.  They are the device
// half of what bed_read_row's fread does in the reference
// (/root/reference/libplinkio/src/bed.c:343-346): the .bed's 3-byte header and the
// ceil(N/4) row length make file offsets unaligned, the panel rows are aligned.
//
//   ROWS layout:   repitch_kernel re-bases each row to the panel's 128-byte row pitch.
//   SAMPLE16:      interleave16_kernel stores sixteen consecutive rows sample-major (one 32-bit word per
//                  sample and group) and re-codes the bit pairs -- see below.
//   GROUP4 layout: interleave_kernel stores four consecutive rows sample-interleaved -- byte i of
//                  group g = the 2-bit fields of sample i for rows 4g..4g+3, row 4g in bits 0-1 --
//                  so that find_pheno's FAST kernel reads its 4-SNP table index without any bit
//                  transposition.  The pass over the data is the one the re-pitch needs anyway.
#include "ctx.cuh"

namespace {

__global__ void __launch_bounds__(256)
repitch_kernel(const uint8_t *__restrict__ src, int64_t src_pitch, int64_t row_bytes, uint8_t *__restrict__ dst,
               int64_t row_stride, int64_t n_rows)
{
    const int64_t words = (row_bytes + 3) >> 2;           // destination 32-bit words holding data
    const int64_t total = n_rows * words;
    const bool aligned = ((row_bytes & 3) == 0) && ((src_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / words, w = t - r * words;
        const uint8_t *s = src + r * src_pitch + 4 * w;
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

// 4 x 4 transpose of 2-bit elements inside a 32-bit word: element (r, s) at bits 8r+2s -> 8s+2r
__device__ __forceinline__ uint32_t transpose4x4_2bit(uint32_t x)
{
    uint32_t t = ((x >> 6) ^ x) & 0x00CC00CCu;
    x ^= t | (t << 6);
    t = ((x >> 12) ^ x) & 0x0000F0F0u;
    x ^= t | (t << 12);
    return x;
}

// One thread per destination word = 4 samples x 4 rows of one group.  Rows of a boundary group
// that lie outside [dst_row, dst_row + n_rows) keep what the panel already holds, so that
// consecutive calls may append rows at any (not only 4-aligned) offset.
__global__ void __launch_bounds__(256)
interleave_kernel(const uint8_t *__restrict__ src, int64_t src_pitch, int64_t row_bytes, uint8_t *__restrict__ panel,
                  int64_t group_stride, int64_t dst_row, int64_t n_rows)
{
    const int64_t g0 = dst_row >> 2, g1 = (dst_row + n_rows - 1) >> 2;
    const int64_t total = (g1 - g0 + 1) * row_bytes;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gl = t / row_bytes, w = t - gl * row_bytes;
        const int64_t g = g0 + gl;
        uint32_t x = 0, keep = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int64_t q = 4 * g + r - dst_row;                 // source row
            if (q >= 0 && q < n_rows) x |= (uint32_t)src[q * src_pitch + w] << (8 * r);
            else keep |= 0x03030303u << (2 * r);
        }
        x = transpose4x4_2bit(x);
        uint32_t *d = reinterpret_cast<uint32_t *>(panel + g * group_stride) + w;
        if (keep) x |= *d & keep;
        *d = x;
    }
}

// GROUP4 panel rows [src_row, src_row + n_rows) -> contiguous packed rows of row_bytes
__global__ void __launch_bounds__(256)
deinterleave_kernel(const uint8_t *__restrict__ panel, int64_t group_stride, int64_t row_bytes, int64_t src_row,
                    int64_t n_rows, uint8_t *__restrict__ dst)
{
    const int64_t total = n_rows * row_bytes;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / row_bytes, w = t - q * row_bytes;
        const int64_t row = src_row + q;
        const uint32_t x = reinterpret_cast<const uint32_t *>(panel + (row >> 2) * group_stride)[w] >> (2 * (int)(row & 3));
        dst[t] = (uint8_t)((x & 3u) | ((x >> 6) & 0xCu) | ((x >> 12) & 0x30u) | ((x >> 18) & 0xC0u));
    }
}

// ROWS panel rows -> contiguous packed rows of row_bytes
__global__ void __launch_bounds__(256)
unpitch_kernel(const uint8_t *__restrict__ panel, int64_t row_stride, int64_t row_bytes, int64_t src_row, int64_t n_rows,
               uint8_t *__restrict__ dst)
{
    const int64_t total = n_rows * row_bytes;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / row_bytes, w = t - q * row_bytes;
        dst[t] = panel[(src_row + q) * row_stride + w];
    }
}

// ---------------------------------------------------------------- SAMPLE16 panel layout
//
// word i of group g = the genotypes of sample i for rows 16g..16g+15 (row 16g in bits 0-1), re-coded
// from the .bed's bit pairs to the A2-allele dosage with 3 = missing:
//     .bed (hi,lo): 00 hom A1, 10 het, 11 hom A2, 01 missing  ->  new_hi = lo, new_lo = hi ^ lo.
// One thread takes one source byte column (4 samples) of a group: 16 byte loads (a warp reads 32
// contiguous bytes of each row), four 4x4 transposes of 2-bit fields (rows <-> samples inside a
// word), a 4x4 byte transpose, the re-code, and one 16-byte store (a warp writes 512 contiguous
// bytes).  Samples >= N are written as 0, so no kernel downstream needs a tail mask.

__device__ __forceinline__ uint32_t recode_to_dosage(uint32_t w)
{
    const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
    return (lo << 1) | (hi ^ lo);
}

__device__ __forceinline__ uint32_t recode_to_bed(uint32_t w)
{
    const uint32_t nl = w & 0x55555555u, nh = (w >> 1) & 0x55555555u;
    return ((nl ^ nh) << 1) | nh;
}

__global__ void __launch_bounds__(256)
interleave16_kernel(const uint8_t *__restrict__ src, int64_t src_pitch, int64_t row_bytes, int64_t n_samples,
                    uint8_t *__restrict__ panel, int64_t s16_stride, int64_t dst_row, int64_t n_rows)
{
    const int64_t g0 = dst_row >> 4, g1 = (dst_row + n_rows - 1) >> 4;
    const int64_t cols = s16_stride >> 4;                          // 4-sample columns per group, padding included
    const int64_t total = (g1 - g0 + 1) * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gl = t / cols, w = t - gl * cols;
        const int64_t g = g0 + gl;
        uint32_t x[4] = {0, 0, 0, 0}, keep = 0;
#pragma unroll
        for (int f = 0; f < 16; f++) {
            const int64_t q = 16 * g + f - dst_row;                // source row
            if (q >= 0 && q < n_rows) { if (w < row_bytes) x[f >> 2] |= (uint32_t)src[q * src_pitch + w] << (8 * (f & 3)); }
            else keep |= 3u << (2 * f);
        }
        // x[b]: byte r = row 4b + r, fields = samples  ->  byte s = sample s, fields = rows 4b .. 4b+3
#pragma unroll
        for (int b = 0; b < 4; b++) x[b] = transpose4x4_2bit(x[b]);
        // word of sample s = bytes s of x[0..3]
        const uint32_t a0 = __byte_perm(x[0], x[1], 0x5140), a1 = __byte_perm(x[0], x[1], 0x7362);
        const uint32_t b0 = __byte_perm(x[2], x[3], 0x5140), b1 = __byte_perm(x[2], x[3], 0x7362);
        uint4 o;
        o.x = __byte_perm(a0, b0, 0x5410); o.y = __byte_perm(a0, b0, 0x7632);
        o.z = __byte_perm(a1, b1, 0x5410); o.w = __byte_perm(a1, b1, 0x7632);
        o.x = recode_to_dosage(o.x); o.y = recode_to_dosage(o.y); o.z = recode_to_dosage(o.z); o.w = recode_to_dosage(o.w);
        const int64_t i0 = 4 * w;
        if (i0 + 0 >= n_samples) o.x = 0;
        if (i0 + 1 >= n_samples) o.y = 0;
        if (i0 + 2 >= n_samples) o.z = 0;
        if (i0 + 3 >= n_samples) o.w = 0;
        uint4 *d = reinterpret_cast<uint4 *>(panel + g * s16_stride) + w;
        if (keep) {
            const uint4 old = *d;
            o.x = (o.x & ~keep) | (old.x & keep); o.y = (o.y & ~keep) | (old.y & keep);
            o.z = (o.z & ~keep) | (old.z & keep); o.w = (o.w & ~keep) | (old.w & keep);
        }
        *d = o;
    }
}

// SAMPLE16 panel rows [src_row, src_row + n_rows) -> contiguous packed .bed rows of row_bytes
__global__ void __launch_bounds__(256)
deinterleave16_kernel(const uint8_t *__restrict__ panel, int64_t s16_stride, int64_t row_bytes, int64_t src_row,
                      int64_t n_rows, uint8_t *__restrict__ dst)
{
    const int64_t total = n_rows * row_bytes;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / row_bytes, w = t - q * row_bytes;
        const int64_t row = src_row + q;
        const int sh = 2 * (int)(row & 15);
        const uint4 v = reinterpret_cast<const uint4 *>(panel + (row >> 4) * s16_stride)[w];
        const uint32_t x = ((v.x >> sh) & 3u) | (((v.y >> sh) & 3u) << 2) | (((v.z >> sh) & 3u) << 4) | (((v.w >> sh) & 3u) << 6);
        dst[t] = (uint8_t)recode_to_bed(x);
    }
}

unsigned grid_for(const simu_b200_ctx *ctx, int64_t total)
{
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

int launch_repitch(simu_b200_ctx *ctx, const uint8_t *d_src, int64_t src_pitch, int64_t n_rows, int64_t dst_row,
                   cudaStream_t stream)
{
    if (n_rows <= 0) return SIMU_B200_OK;
    miss_index_invalidate(ctx);                   // the panel changes
    if (ctx->layout == SIMU_B200_LAYOUT_SAMPLE16) {
        const int64_t groups = ((dst_row + n_rows - 1) >> 4) - (dst_row >> 4) + 1;
        interleave16_kernel<<<grid_for(ctx, groups * (ctx->s16_stride >> 4)), 256, 0, stream>>>(
            d_src, src_pitch, ctx->row_bytes, ctx->n_samples, ctx->panel, ctx->s16_stride, dst_row, n_rows);
    } else if (ctx->layout == SIMU_B200_LAYOUT_GROUP4) {
        const int64_t groups = ((dst_row + n_rows - 1) >> 2) - (dst_row >> 2) + 1;
        interleave_kernel<<<grid_for(ctx, groups * ctx->row_bytes), 256, 0, stream>>>(
            d_src, src_pitch, ctx->row_bytes, ctx->panel, ctx->group_stride, dst_row, n_rows);
    } else {
        repitch_kernel<<<grid_for(ctx, n_rows * ((ctx->row_bytes + 3) >> 2)), 256, 0, stream>>>(
            d_src, src_pitch, ctx->row_bytes, ctx->panel + dst_row * ctx->row_stride, ctx->row_stride, n_rows);
    }
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "repitch/interleave kernel launch");
}

int launch_unpanel(simu_b200_ctx *ctx, int64_t src_row, int64_t n_rows, uint8_t *d_dst, cudaStream_t stream)
{
    if (n_rows <= 0) return SIMU_B200_OK;
    const unsigned grid = grid_for(ctx, n_rows * ctx->row_bytes);
    if (ctx->layout == SIMU_B200_LAYOUT_SAMPLE16)
        deinterleave16_kernel<<<grid, 256, 0, stream>>>(ctx->panel, ctx->s16_stride, ctx->row_bytes, src_row, n_rows, d_dst);
    else if (ctx->layout == SIMU_B200_LAYOUT_GROUP4)
        deinterleave_kernel<<<grid, 256, 0, stream>>>(ctx->panel, ctx->group_stride, ctx->row_bytes, src_row, n_rows, d_dst);
    else
        unpitch_kernel<<<grid, 256, 0, stream>>>(ctx->panel, ctx->row_stride, ctx->row_bytes, src_row, n_rows, d_dst);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "unpanel kernel launch");
}
