// Internal declarations shared by the kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/simu_b200.h"

#define SIMU_B200_VERSION 100
#define SIMU_ROW_ALIGN 128            // HBM row pitch granularity (bytes)
#define SIMU_STAGE_BYTES (64u << 20)  // pinned staging chunk (2 of them)
#define SIMU_NUM_SMS_DEFAULT 148

struct simu_b200_peer;

struct simu_b200_ctx {
    int device = 0;
    int num_sms = SIMU_NUM_SMS_DEFAULT;
    int64_t n_samples = 0;
    int64_t row_bytes = 0;    // ceil(N/4)
    int64_t row_stride = 0;   // pitch in HBM
    uint8_t *panel = nullptr;
    int64_t panel_rows = 0;
    size_t panel_bytes = 0;
    int layout = SIMU_B200_LAYOUT_ROWS;   // how the panel is laid out in HBM (simu_b200_layout)
    int64_t group_stride = 0; // GROUP4: bytes per 4-row group = round_up(N, 128), one byte per sample
    int64_t s16_stride = 0;   // SAMPLE16: bytes per 16-row group = 4 * round_up(N, 128), one word per sample

    cudaStream_t own_stream = nullptr;   // created by open
    cudaStream_t stream = nullptr;       // stream in use (own or external)
    cudaStream_t copy_stream = nullptr;  // side stream for staging
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint8_t *pinned[2] = {nullptr, nullptr};
    uint8_t *dstage[2] = {nullptr, nullptr};            // device bounce buffers for 1-D staging
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_repitch[2] = {nullptr, nullptr};
    bool dstage_used[2] = {false, false};

    // grow-on-demand device scratch
    void *scratch[32] = {nullptr};
    size_t scratch_bytes[32] = {0};

    // optional per-kernel device timing (CUDA events on the launching stream)
    bool prof_on = false;
    static constexpr int kProfIds = 8, kProfCap = 512;
    cudaEvent_t prof_ev[kProfCap][2];
    int prof_id[kProfCap];
    int prof_n = 0;
    bool prof_alloc = false;

    bool lut_attr_set[3] = {false, false, false};
    bool tc_attr_set = false, cnt_attr_set = false;
    int coop_sort_blocks = -1;           // co-resident blocks of the cooperative sort (-1: not queried)
    int64_t launches = 0;
    float last_ms[2] = {0.f, 0.f};
    // EXACT mode on SAMPLE16 panels takes the effects from the kernel-parameter bank: host copies of the current call's
    // effects (set by the host-buffer entry points for the duration of the call) or a pinned fetch buffer
    const double *h_x1 = nullptr, *h_x2 = nullptr;
    double *h_eff = nullptr;
    size_t h_eff_bytes = 0;
    // missing-genotype index of the SAMPLE16 panel (miss_index.cu): 0 = not built, 1 = sparse list ready, 2 = dense m plane
    int miss_state = 0, miss_chunks = 1, miss_uses = 0;
    int64_t miss_total = 0;
    simu_b200_peer *peer = nullptr;          // multi-GPU exchange over NVLink peer memory (peer.cu)
    std::string err;
};

enum ScratchSlot {
    SCR_ROWS = 0, SCR_COUNTS, SCR_FREQ, SCR_X1, SCR_X2, SCR_Y1, SCR_Y2, SCR_NOISE,
    SCR_LUT, SCR_YPART, SCR_SORT, SCR_MISC, SCR_FUSED, SCR_RAW1, SCR_RAW2, SCR_SPOW, SCR_SYNC, SCR_DEBUG, SCR_TCW,
    SCR_MISS_CNT, SCR_MISS_ENT, SCR_MISS_OFF, SCR_WMFIX, SCR_CORR, SCR_CNTACC, SCR_HERIT
};

int ctx_fail(simu_b200_ctx *ctx, int code, const char *fmt, ...);
int ctx_cuda(simu_b200_ctx *ctx, cudaError_t e, const char *what);
void *ctx_scratch(simu_b200_ctx *ctx, int slot, size_t bytes);  // nullptr on failure (error set)

#define CUDA_TRY(ctx, expr)                                          \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return ctx_cuda((ctx), _e, #expr);    \
    } while (0)

enum ProfId { PROF_COUNTS = 0, PROF_PHENO_EXACT = 1, PROF_LUT_MAIN = 2, PROF_LUT_BUILD = 3, PROF_LUT_REDUCE = 4,
              PROF_REDUCTIONS = 5, PROF_SORT = 6 };
// record a begin / end event pair around a kernel when profiling is enabled
int prof_begin(simu_b200_ctx *ctx, int id);      // returns slot or -1
void prof_end(simu_b200_ctx *ctx, int slot);

// ---- kernel launchers (each returns a simu_b200_status) --------------------

// kernel (a): per-SNP allele counts + frequency.  counts.cu
int launch_counts(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, int32_t *d_counts, double *d_freq);

int launch_counts_g4_range(simu_b200_ctx *ctx, cudaStream_t stream, int64_t first_group, int64_t mc, int32_t *d_counts,
                           double *d_freq, int threads);

// SAMPLE16 panels: bit-sliced vertical counting.  counts_s16.cu
int launch_counts_s16(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int32_t *d_counts, double *d_freq);

// kernel (b) on SAMPLE16 panels: tcgen05 integer MMA (FAST) and ordered FP64 (EXACT).  pheno_tc.cu, pheno_exact.cu
// accumulate != 0: d_y holds the sums over earlier panels' rows and this panel's rows are added to them
int launch_pheno_tc(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *d_x1, const double *d_x2,
                    double *d_y1, double *d_y2, int accumulate);
int launch_pheno_exact_s16(simu_b200_ctx *ctx, int64_t mc, const double *d_freq, const double *d_x1, const double *d_x2,
                           double *d_y1, double *d_y2, int accumulate);

// missing-genotype index of a SAMPLE16 panel and the correction pass over it.  miss_index.cu
void miss_index_invalidate(simu_b200_ctx *ctx);
int miss_index_ensure(simu_b200_ctx *ctx);          // 1 sparse index ready, 0 use the dense plane, < 0 -status
int64_t miss_corr_stride(const simu_b200_ctx *ctx);
bool miss_correct_accumulates(const simu_b200_ctx *ctx);
int launch_miss_correct(simu_b200_ctx *ctx, cudaStream_t stream, int64_t mc, int traits, const long long *d_wm, unsigned long long *d_corr);

// kernel (b), exact FP64 ordered mode.  pheno_exact.cu
int launch_pheno_exact(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                       const double *d_x1, const double *d_x2, double *d_y1, double *d_y2);

// kernel (b), fast FP32 4-SNP LUT mode.  pheno_lut.cu
int launch_pheno_lut(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc, const double *d_freq,
                     const double *d_x1, const double *d_x2, double *d_y1, double *d_y2);

// kernel (c): reductions.  reduce.cu
int launch_sumsq(simu_b200_ctx *ctx, const double *d_y, int64_t n, double *d_out);
// sum of squares + coefficients + scaling in one cooperative launch (coefficients to d_coef[0..1] when not NULL)
int launch_apply_heritability(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise, double *d_coef);
int launch_heritability(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n, const double *d_noise,
                        const double *d_sumsq, double *d_coef);
int launch_argsort_f64(simu_b200_ctx *ctx, const double *d_key, int64_t n, int32_t *d_index);
// is_case[i] = 0 for the max_ncon samples of lowest liability (ties: lowest index first), 1 for the rest
int launch_liability_split(simu_b200_ctx *ctx, const double *d_pheno, int64_t n, int64_t max_ncon, uint8_t *d_is_case);
// effect-size scaling by het^S / sqrt(het).  d_spow: [2][ncomp]; *d_bad: first variant with zero het
int launch_scale_effects(simu_b200_ctx *ctx, const double *d_freq, double *d_x1, double *d_x2, int64_t mc,
                         const int32_t *d_comp, const double *d_spow, int ncomp, int op, unsigned long long *d_bad);
// the same on an explicit stream, with `index_base` added to the variant index reported in *d_bad
int launch_scale_effects_on(simu_b200_ctx *ctx, cudaStream_t stream, const double *d_freq, double *d_x1, double *d_x2,
                            int64_t mc, const int32_t *d_comp, const double *d_spow, int ncomp, int op,
                            unsigned long long *d_bad, int64_t index_base);

// staging: move n_rows packed rows (pitch src_pitch) from a device bounce buffer into panel rows
// [dst_row, dst_row + n_rows) -- re-pitched (ROWS) or sample-interleaved (GROUP4).  stage.cu
int launch_repitch(simu_b200_ctx *ctx, const uint8_t *d_src, int64_t src_pitch, int64_t n_rows, int64_t dst_row,
                   cudaStream_t stream);
// the inverse, for panel_get_rows: panel rows -> contiguous packed rows (pitch row_bytes).  stage.cu
int launch_unpanel(simu_b200_ctx *ctx, int64_t src_row, int64_t n_rows, uint8_t *d_dst, cudaStream_t stream);

// synthetic panel generator: rows first_row + r (d_row_ids == NULL) or d_row_ids[r], r < n_rows, written
// as packed rows at dst + r * dst_pitch (dst_pitch a multiple of 4).  synth.cu
int launch_synth(simu_b200_ctx *ctx, uint64_t seed, int64_t first_row, const int64_t *d_row_ids, int64_t n_rows,
                 uint8_t *dst, int64_t dst_pitch);

// One grid-wide barrier inside a cooperative launch (all blocks co-resident) that cleans up after itself: the last block to
// leave zeroes both counters, so a kernel's launch parameters never change and it replays inside a CUDA graph.
__device__ __forceinline__ void grid_barrier_once(unsigned int *bar /* [2], zero between launches */)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&bar[0], 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < gridDim.x);
        if (atomicAdd(&bar[1], 1u) == gridDim.x - 1u) {      // everyone has seen the full count: reset for the next launch
            bar[0] = 0u;
            bar[1] = 0u;
            __threadfence();
        }
    }
    __syncthreads();
}

