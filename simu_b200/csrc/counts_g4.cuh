// GROUP4 allele counting (counts_g4_kernel, counts.cu).
//
// In a GROUP4 panel a 32-bit word holds 4 samples x 4 SNP rows (row r of the group in bits
// 8s+2r, 8s+2r+1), so ONE pass over a group yields the counts of its four SNPs: the words go
// through a carry-save adder tree, only the final POPCs are taken under per-row masks.
// LOP3 / SHF issue on the half-rate ALU pipe (64 lanes/clk/SM), which makes this code
// ALU-bound long before it is HBM-bound unless the logic per word is kept near the tree's
// inherent 2 LOP3 per word and plane:
//   * the tree is 5 levels deep (ones .. sixteens, POPC of the weight-32 carry word), so the 12
//     masked POPCs (4 rows x {lo, hi, lo&hi}) are paid once per 32 words = 512 genotypes;
//   * the lo&hi plane is fed into a second tree as w & (w << 1) -- the shift is an IMAD.SHL on
//     the FMA pipe; the even bits of that word are garbage, but carry-save adders are bitwise,
//     so garbage never reaches the odd bits the masks select;
//   * the tail vector (samples >= N masked off) and ragged batch ends run in a separate,
//     predicated copy of the batch code, so the steady-state loop carries no predicates.
#pragma once

#include <stdint.h>

#include <type_traits>

namespace g4 {

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t word_mask(int64_t valid_bits, int w)
{
    int64_t b = valid_bits - 32 * w;
    if (b <= 0) return 0u;
    if (b >= 32) return 0xFFFFFFFFu;
    return (1u << (int)b) - 1u;
}

__device__ __forceinline__ void csa(uint32_t &sum, uint32_t &carry, uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t s, m;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(m) : "r"(a), "r"(b), "r"(c));   // majority
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(a), "r"(b), "r"(c));   // parity
    sum = s; carry = m;
}

// ones/twos/fours/eights accumulators; add() folds 16 words (clobbered) into them and returns the
// weight-16 carry word
struct Tree16 {
    uint32_t acc[4] = {0, 0, 0, 0};
    __device__ __forceinline__ uint32_t add(uint32_t (&w)[16])
    {
#pragma unroll
        for (int l = 0; l < 4; l++) {
#pragma unroll
            for (int i = 0; i < (8 >> l); i++) csa(acc[l], w[i], acc[l], w[2 * i], w[2 * i + 1]);
        }
        return w[0];
    }
};

struct PlaneCounts {
    Tree16 tree;
    uint32_t sixteens = 0;
    // two weight-16 carries -> one weight-32 carry
    __device__ __forceinline__ uint32_t pair(uint32_t a, uint32_t b)
    {
        uint32_t c;
        csa(sixteens, c, sixteens, a, b);
        return c;
    }
    __device__ __forceinline__ uint32_t residual(uint32_t m) const
    {
        return 16u * __popc(sixteens & m) + 8u * __popc(tree.acc[3] & m) + 4u * __popc(tree.acc[2] & m) +
               2u * __popc(tree.acc[1] & m) + __popc(tree.acc[0] & m);
    }
};

// Lane-partial counts over the 16-byte vectors i0, i0 + stride, i0 + 2 stride, ... < nvec of ONE
// group row p: lo[r] / hi[r] / both[r] = number of samples whose field for row r has its low bit /
// high bit / both bits set.  Vector index `nfull` (if this lane meets it) is masked to tail_bits
// valid bits: samples >= N are never counted.
__device__ __forceinline__ void count_vectors(const uint4 *__restrict__ p, int i0, const int stride, const int nvec,
                                              const int nfull, const int64_t tail_bits, uint32_t (&lo)[4],
                                              uint32_t (&hi)[4], uint32_t (&both)[4])
{
    PlaneCounts raw, bth;
    uint32_t l32[4] = {0, 0, 0, 0}, h32[4] = {0, 0, 0, 0}, b32[4] = {0, 0, 0, 0};
    // one batch = 8 vectors = 32 words of this lane; CHECKED: ragged end + masked tail vector
    auto batch = [&](int j0, auto checked) {
        constexpr bool kChecked = decltype(checked)::value;
        uint4 v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = j0 + q * stride;
            if (!kChecked || i < nvec) v[q] = ldg_stream(p + i);
            else v[q] = make_uint4(0u, 0u, 0u, 0u);
            if (kChecked && i == nfull) {
                v[q].x &= word_mask(tail_bits, 0); v[q].y &= word_mask(tail_bits, 1);
                v[q].z &= word_mask(tail_bits, 2); v[q].w &= word_mask(tail_bits, 3);
            }
        }
        uint32_t s16[2], b16[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t w[16], b[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[4 * q] = v[4 * h + q].x; w[4 * q + 1] = v[4 * h + q].y;
                w[4 * q + 2] = v[4 * h + q].z; w[4 * q + 3] = v[4 * h + q].w;
            }
#pragma unroll
            for (int q = 0; q < 16; q++) b[q] = w[q] & (w[q] << 1);
            s16[h] = raw.tree.add(w);
            b16[h] = bth.tree.add(b);
        }
        const uint32_t s = raw.pair(s16[0], s16[1]), sb = bth.pair(b16[0], b16[1]);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            l32[r] += __popc(s & (0x01010101u << (2 * r)));
            h32[r] += __popc(s & (0x02020202u << (2 * r)));
            b32[r] += __popc(sb & (0x02020202u << (2 * r)));
        }
    };
    for (; i0 + 7 * stride < nfull && i0 + 7 * stride < nvec; i0 += 8 * stride) batch(i0, std::false_type());
    for (; i0 < nvec; i0 += 8 * stride) batch(i0, std::true_type());
#pragma unroll
    for (int r = 0; r < 4; r++) {
        lo[r] += 32u * l32[r] + raw.residual(0x01010101u << (2 * r));
        hi[r] += 32u * h32[r] + raw.residual(0x02020202u << (2 * r));
        both[r] += 32u * b32[r] + bth.residual(0x02020202u << (2 * r));
    }
}

}  // namespace g4
