// Device-side synthetic panel generator (DESIGN.md "Synthetic
// panel").  Integer-only, so it is byte-identical to the host generator the
// tests use: per SNP j,  p_j ~ U(0.01, 0.5);  A1 count ~ Binomial(2, p_j);
// each genotype missing with probability 0.001.
//
//   key_j   = mix64(seed ^ mix64(j))
//   thr_j   = 10486 + floor(hi32(key_j) * 513802 / 2^32)        (20-bit scale)
//   h       = mix64(key_j + i)
//   a1      = (h[0:20) < thr_j) + (h[20:40) < thr_j)
//   missing = h[40:64) < 16777                                   (24-bit scale)
//   bits    = missing ? 01 : a1==2 ? 00 : a1==1 ? 10 : 11
//
// The 275 GB / 125 GB panels of configs 4-5 cannot exist on disk; they are
// generated straight into HBM with this kernel.
#include "ctx.cuh"

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
synth_kernel(uint8_t *__restrict__ dst, int64_t row_stride, int64_t n_samples, uint64_t seed,
             int64_t first_row, const int64_t *__restrict__ row_ids, int64_t n_rows)
{
    const int64_t words_per_row = row_stride >> 2;
    const int64_t total = n_rows * words_per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / words_per_row, w = t - r * words_per_row;
        const uint64_t key = mix64(seed ^ mix64((uint64_t)(row_ids ? row_ids[r] : first_row + r)));
        const uint32_t thr = 10486u + (uint32_t)(((key >> 32) * 513802ull) >> 32);
        uint32_t word = 0;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int64_t i = 16 * w + s;
            if (i < n_samples) {
                const uint64_t h = mix64(key + (uint64_t)i);
                const uint32_t u1 = (uint32_t)(h & 0xFFFFF), u2 = (uint32_t)((h >> 20) & 0xFFFFF);
                const uint32_t u3 = (uint32_t)(h >> 40);
                const uint32_t a1 = (u1 < thr) + (u2 < thr);
                const uint32_t bits = (u3 < 16777u) ? 1u : (a1 == 2 ? 0u : (a1 == 1 ? 2u : 3u));
                word |= bits << (2 * s);
            }
        }
        reinterpret_cast<uint32_t *>(dst + r * row_stride)[w] = word;
    }
}

}  // namespace

int launch_synth(simu_b200_ctx *ctx, uint64_t seed, int64_t first_row, const int64_t *d_row_ids, int64_t n_rows,
                 uint8_t *dst, int64_t dst_pitch)
{
    if (n_rows <= 0) return SIMU_B200_OK;
    const int64_t total = n_rows * (dst_pitch >> 2);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(dst, dst_pitch, ctx->n_samples, seed, first_row, d_row_ids,
                                                            n_rows);
    ctx->launches++;
    return ctx_cuda(ctx, cudaGetLastError(), "synth_kernel launch");
}
