// Pieces shared by the persistent producer/consumer kernels (pheno_lut.cu, pheno_tc.cu): mbarrier and
// TMA bulk-copy PTX wrappers, and the walk over a CTA's contiguous range of (sample tile, row block) steps.
#pragma once

#include <stdint.h>

namespace {

// ------------------------------------------------------------- PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_test(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_test(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// Producer-side wait: poll with back-off so that the waiting warp does not burn issue slots
// the consumer warps need (the bare try_wait loop was 35 % of all executed instructions).
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity)
{
    if (mbar_test(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_test(bar, parity)) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}

// Every CTA walks its contiguous range of the (sample-tile major, row-block minor) step list in
// ROTATED order: at local time it, it is at the step whose row block b satisfies b == it (mod len).
// All CTAs (one per sample tile for any given block) therefore need the same ~nblocks/len table
// blocks at the same time, so a table is fetched from HBM once and served to the other tiles' CTAs
// from L2.  (tile, block) of the current step are carried along incrementally: 64-bit divisions
// cost ~100 instructions each and used to be a quarter of all issued instructions.
struct StepWalk {
    struct Cursor { int64_t s, tile, blk; };
    int64_t s_begin, s_end, len, nb, tile_b, blk_b;
    Cursor first;
    __device__ __forceinline__ StepWalk(int64_t steps, int64_t nblocks, int64_t cta, int64_t grid)
    {
        s_begin = (cta * steps) / grid; s_end = ((cta + 1) * steps) / grid;
        len = s_end - s_begin;
        nb = nblocks;
        const int64_t rot = len > 0 ? (len - (s_begin - (s_begin / nb) * nb) % len) % len : 0;
        tile_b = s_begin / nb; blk_b = s_begin - tile_b * nb;
        first.s = s_begin + rot;
        first.tile = tile_b + (blk_b + rot) / nb;
        first.blk = (blk_b + rot) - (first.tile - tile_b) * nb;
    }
    __device__ __forceinline__ void advance(Cursor &c) const
    {
        c.s++; c.blk++;
        if (c.blk == nb) { c.blk = 0; c.tile++; }
        if (c.s == s_end) { c.s = s_begin; c.tile = tile_b; c.blk = blk_b; }
    }
};

__device__ __forceinline__ int64_t cta_of_step(int64_t s, int64_t grid, int64_t steps)
{
    return ((s + 1) * grid - 1) / steps;   // largest c with (c*steps)/grid <= s
}

}  // namespace
