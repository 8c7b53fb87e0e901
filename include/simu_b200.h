/*
 * simu_b200.h -- C ABI of the B200-native phenotype-synthesis hot path
 * (y = Gx + e over a 2-bit packed PLINK .bed) that replaces, inside
 * precimed/simu, the two row-by-row passes over the .bed and the scalar
 * reductions that follow them.
 *
 * Reference seams this ABI sits behind (paths relative to /root/reference):
 *   find_freq()                   src/source.cc:843-885
 *   find_pheno()                  src/source.cc:951-1006
 *   apply_heritability()          src/source.cc:1008-1029
 *   apply_liability_threshold()   src/source.cc:1031-1055
 *   PioFile/PioFiles row cursor   src/source.cc:187-361
 *   pio_open / pio_next_row / pio_skip_row / pio_reset_row / pio_close
 *                                 libplinkio/src/plinkio/plinkio.h:58-214
 *   bed_read_row + unpack_snps    libplinkio/src/bed.c:91-114, :321-358
 *
 * Conventions (carried over from libplinkio's C ABI, plinkio/status.h:16-47):
 * integer status returns, caller-owned output buffers, no exceptions across
 * the boundary, one context per run and per GPU, calls made from one host
 * thread.  There is NO CPU fallback: every entry point fails with
 * SIMU_B200_ECUDA when no sm_100 device / driver is available.
 *
 * Genotype coding (libplinkio/src/bed.c:75-85): packed bit pair, low bits
 * first inside a byte: 00 = hom A1 (code 0, two A1 copies), 10 = het (code 1),
 * 11 = hom A2 (code 2), 01 = missing (code 3).
 */
#ifndef SIMU_B200_H
#define SIMU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SIMU_B200_API __attribute__((visibility("default")))
#else
#define SIMU_B200_API
#endif

/* Status codes.  0..2 mirror pio_status_t (PIO_OK, PIO_END, PIO_ERROR). */
enum simu_b200_status {
    SIMU_B200_OK = 0,
    SIMU_B200_END = 1,      /* fewer rows in the .bed than announced (source.cc:855-858) */
    SIMU_B200_ERROR = 2,    /* generic / read error (bed.c:349-352) */
    SIMU_B200_EINVAL = 3,   /* bad argument */
    SIMU_B200_ECUDA = 4,    /* CUDA runtime / driver / device error, or no GPU */
    SIMU_B200_ENOMEM = 5,   /* host or device allocation failed */
    SIMU_B200_EFORMAT = 6   /* not a v1.00 SNP-major .bed (source.cc:196-200) */
};

/* find_pheno accumulation modes. */
enum simu_b200_mode {
    /* FP64, per-sample sum in ascending causal-SNP order, separate
     * subtract/multiply/add: bit-identical to source.cc:986-996 on one GPU. */
    SIMU_B200_MODE_EXACT = 0,
    /* Throughput mode, |dy| <= 1e-5 * max|y| (tests state the bound).
     * SAMPLE16 panels: tcgen05 integer MMA -- per-SNP weights quantised to 40-bit fixed point
     * (relative to the largest weight of the trait) and split into signed 8-bit digits, genotype
     * planes as unsigned bytes, exact 32-bit integer accumulation in TMEM, FP64 only for the final
     * recombination: the result is the exact sum of the quantised weights (error ~1e-12 of max|y|).
     * ROWS / GROUP4 panels: 4-SNP lookup tables in FP32 (entries rounded from FP64), FP32 partial
     * sums flushed to FP64 (~1e-7); entries and 1024-term partial sums must fit FP32. */
    SIMU_B200_MODE_FAST = 1
};

/* How the panel is laid out in HBM. */
enum simu_b200_layout {
    /* One SNP row after the other, exactly as in the .bed, at a 128-byte pitch.  Any row of a
     * resident panel can be addressed through the `rows` lists of find_freq / find_pheno
     * (several simulations with different causal sets over one resident panel). */
    SIMU_B200_LAYOUT_ROWS = 0,
    /* Four consecutive rows stored sample-interleaved: byte i of group g holds the 2-bit fields of
     * sample i for rows 4g..4g+3 (row 4g in bits 0-1).  For COMPACT panels -- only the causal rows
     * staged, in causal order, which is what the reference reads (source.cc:852-863) -- so `rows`
     * must be NULL.  find_pheno's FAST kernel then needs no bit transposition (config 4: 8.3 ->
     * 6.7 ms); the interleave rides on the staging pass, which rewrites every row anyway. */
    SIMU_B200_LAYOUT_GROUP4 = 1,
    /* Sixteen consecutive rows stored sample-major: 32-bit word i of group g holds the genotypes of
     * sample i for rows 16g..16g+15 (row 16g in bits 0-1), RE-CODED on the way in as the A2-allele
     * dosage: 0 = hom A1, 1 = het, 2 = hom A2, 3 = missing (.bed bits 00/10/11/01 -> 00/01/10/11), so
     * that a genotype's contribution is affine in the code except for the one AND-detectable missing
     * state.  Compact panels only (`rows` must be NULL), pitch = round_up(N, 128) words per group;
     * samples >= N and unwritten rows hold 0.  This is the layout of the tensor-core phenotype kernel
     * (MODE_FAST): a thread owns one sample = one TMEM lane and expands its 16-SNP words into the
     * K-dimension of a tcgen05 integer MMA.  panel_put_rows / load_bed / get_rows convert to and from
     * .bed rows, so nothing outside the library sees the coding. */
    SIMU_B200_LAYOUT_SAMPLE16 = 2
};

typedef struct simu_b200_ctx simu_b200_ctx;

/* ------------------------------------------------------------------ context */

SIMU_B200_API int simu_b200_version(void);

/* Number of CUDA devices; *n = 0 and SIMU_B200_ECUDA when there is none. */
SIMU_B200_API int simu_b200_device_count(int *n);

/* One context per run and GPU.  num_samples = N of the .fam
 * (pio_num_samples, plinkio.h:133).  Fails on a device that is not sm_100. */
SIMU_B200_API int simu_b200_open(simu_b200_ctx **ctx, int device, int64_t num_samples);
SIMU_B200_API void simu_b200_close(simu_b200_ctx *ctx);            /* pio_close */

/* Text of the last error on this context (ctx == NULL: last open error). */
SIMU_B200_API const char *simu_b200_last_error(const simu_b200_ctx *ctx);

/* Run all subsequent work on an external CUDA stream (cudaStream_t), e.g.
 * torch's current stream; NULL restores the context's own stream. */
SIMU_B200_API int simu_b200_set_stream(simu_b200_ctx *ctx, void *cuda_stream);
SIMU_B200_API int simu_b200_sync(simu_b200_ctx *ctx);

/* Number of kernels this context has launched so far. */
SIMU_B200_API int64_t simu_b200_launch_count(const simu_b200_ctx *ctx);

/* -------------------------------------------------------------------- panel
 * The panel is the HBM-resident packed genotype matrix: `rows` SNP rows of
 * row_bytes = ceil(N/4) bytes (bed_header.c:219-223) at a pitch of
 * row_stride bytes (multiple of 128).  It replaces the FILE* + one-row
 * read_buffer of pio_bed_file_t (plinkio/bed.h:30-56). */

SIMU_B200_API int64_t simu_b200_row_bytes(int64_t num_samples);
SIMU_B200_API int64_t simu_b200_row_stride(const simu_b200_ctx *ctx);
SIMU_B200_API int64_t simu_b200_panel_rows(const simu_b200_ctx *ctx);
SIMU_B200_API void *simu_b200_panel_device_ptr(const simu_b200_ctx *ctx);

/* panel_alloc = panel_alloc_layout(SIMU_B200_LAYOUT_ROWS).  The panel is zero-filled (all 00). */
SIMU_B200_API int simu_b200_panel_alloc(simu_b200_ctx *ctx, int64_t num_rows);
SIMU_B200_API int simu_b200_panel_alloc_layout(simu_b200_ctx *ctx, int64_t num_rows, int layout);
SIMU_B200_API int simu_b200_panel_layout(const simu_b200_ctx *ctx);
/* GROUP4: bytes between consecutive 4-row groups = round_up(N, 128); SAMPLE16: bytes between consecutive
 * 16-row groups = 4 * round_up(N, 128). */
SIMU_B200_API int64_t simu_b200_group_stride(const simu_b200_ctx *ctx);
SIMU_B200_API int simu_b200_panel_free(simu_b200_ctx *ctx);

/* Copy n_rows packed rows from host memory (pitch host_stride >= row_bytes)
 * into panel rows [dst_row, dst_row+n_rows): pinned-host chunked staging on a
 * side stream (one contiguous DMA per 64 MB chunk + a device re-pitch / interleave
 * kernel).  Rows may be appended at any offset, also inside a GROUP4 group.
 * Replaces the fread of bed_read_row (bed.c:343-346).  Pageable sources are
 * copied through the context's own pinned buffers and may be reused on return;
 * a PAGE-LOCKED source is read by DMA asynchronously: leave it untouched until
 * simu_b200_sync() or the next host-buffer hot-path call returns. */
SIMU_B200_API int simu_b200_panel_put_rows(simu_b200_ctx *ctx, int64_t dst_row, int64_t n_rows,
                                           const void *host_rows, int64_t host_stride);

/* Read rows of a .bed file into the panel.  file_rows (ascending row numbers
 * inside this file, NULL = 0..n_rows-1) are the causal rows; all others are
 * seeked past, as pio_skip_row does (bed.c:360-371); runs of adjacent rows are
 * read with single preads by up to 8 reader threads (SIMU_B200_IO_THREADS).
 * num_loci is the .bim line count.  Validates the 3-byte header (bed_header.c:80-102): returns
 * SIMU_B200_EFORMAT unless v1.00 SNP-major (source.cc:196-200), SIMU_B200_END
 * if a requested row is >= num_loci (the stream ends early, source.cc:855-858)
 * and SIMU_B200_ERROR on a short read (bed.c:349-352 -> source.cc:226-229). */
SIMU_B200_API int simu_b200_panel_load_bed(simu_b200_ctx *ctx, const char *bed_path, int64_t num_loci,
                                           const int64_t *file_rows, int64_t n_rows, int64_t dst_row);

/* Fill panel rows [dst_row, dst_row+n_rows) with the synthetic panel of
 * SURVEY.md 8(d) / DESIGN.md (global SNP rows first_row..): device-side
 * generator, byte-identical to the host generator used by the tests. */
SIMU_B200_API int simu_b200_panel_synth(simu_b200_ctx *ctx, uint64_t panel_seed, int64_t first_row,
                                        int64_t n_rows, int64_t dst_row);

/* Same generator for a scattered list of global SNP rows (HOST array row_ids, e.g. the causal
 * rows of a sparse configuration) -> panel rows dst_row + r. */
SIMU_B200_API int simu_b200_panel_synth_rows(simu_b200_ctx *ctx, uint64_t panel_seed, const int64_t *row_ids,
                                             int64_t n_rows, int64_t dst_row);

/* Copy panel rows back to the host as packed .bed rows, whatever the layout (tests, e2e
 * staging of synthetic data). */
SIMU_B200_API int simu_b200_panel_get_rows(simu_b200_ctx *ctx, int64_t src_row, int64_t n_rows,
                                           void *host_rows, int64_t host_stride);

/* ------------------------------------------- hot path, host-buffer entry points
 * These are the calls the reference's main() makes where it calls the four
 * functions today; all pointers are HOST pointers. */

/* find_freq (source.cc:843-885) over the causal SNPs.  rows[j] = panel row of
 * the j-th causal SNP in ascending variant order (NULL = 0..mc-1; must be NULL
 * for a GROUP4 panel).
 * counts: mc x 4 = genocount[0..3] (source.cc:869), freq: mc doubles with the
 * exact expression of source.cc:874-875.  Either may be NULL. */
SIMU_B200_API int simu_b200_find_freq(simu_b200_ctx *ctx, const int32_t *rows, int64_t mc,
                                      int32_t *counts, double *freq);

/* find_pheno (source.cc:951-1006).  freq/effect1/effect2 have one entry per
 * causal SNP (same order as rows); effect2/pheno2 NULL for one trait.
 * pheno arrays (N doubles) are overwritten. */
SIMU_B200_API int simu_b200_find_pheno(simu_b200_ctx *ctx, const int32_t *rows, int64_t mc,
                                       const double *freq, const double *effect1, const double *effect2,
                                       double *pheno1, double *pheno2, int mode);

/* find_pheno over the rows of the CURRENT panel, ADDED to the sums pheno1/pheno2 already hold: for causal sets
 * larger than HBM the host loads them panel by panel, in ascending order (the reference streams the same rows
 * from disk, source.cc:968-984), and calls this once per panel after zeroing pheno.  EXACT mode continues every
 * sample's chain of FP64 additions, so the result is bit-identical to a single pass (and to the reference);
 * FAST mode adds each panel's exact fixed-point sum.  SAMPLE16 panels only; rows 0..mc-1. */
SIMU_B200_API int simu_b200_find_pheno_accumulate(simu_b200_ctx *ctx, int64_t mc, const double *freq,
                                                  const double *effect1, const double *effect2, double *pheno1,
                                                  double *pheno2, int mode);

/* Free and total HBM of the context's device, in bytes; `free` counts the current panel as free (a new
 * panel_alloc replaces it).  panel_bytes: what panel_alloc_layout(num_rows, layout) allocates. */
SIMU_B200_API int simu_b200_mem_info(simu_b200_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes);
SIMU_B200_API int64_t simu_b200_panel_bytes(const simu_b200_ctx *ctx, int64_t num_rows, int layout);

/* find_freq + effect scaling + find_pheno (FAST mode) behind ONE call, everything on the device
 * (no host round trip between the passes): the standard kernels queued back to back on the context's
 * stream.  Compact panels (GROUP4 or SAMPLE16); rows 0..mc-1.
 * The call replaces this stretch of main():
 *     find_freq(...)                                   source.cc:1235
 *     draw / read effect sizes, scale them by het      source.cc:1238-1290
 *     find_pheno(...)                                  source.cc:1295
 * raw1/raw2: the effect sizes BEFORE the het scaling (the RNG draws of :887-949, or the file's
 * values) -- the draws do not depend on the frequencies, so the host makes them first, in the
 * reference's RNG order.  scale_op: -1 none; 0: *= sqrt(pow(het, S)) (--gcta-sigma S = -1,
 * --traitK-s-pow); 1: /= sqrt(het) (--norm-effect with --causal-variants effects), as in
 * simu_b200_scale_effects.  Outputs: counts / freq (as find_freq), effect1/effect2 = the SCALED
 * effects (what .causals reports after apply_heritability), pheno1/pheno2.  Any output except
 * pheno may be NULL.  Returns SIMU_B200_ERROR with *bad_index set when a causal variant has
 * het < FLT_EPSILON under scale_op 0/1 (the reference throws, :1251-1255). */
SIMU_B200_API int simu_b200_find_freq_pheno(simu_b200_ctx *ctx, int64_t mc, const double *raw1, const double *raw2,
                                            int scale_op, const int32_t *component, const double *s_pow1,
                                            const double *s_pow2, int num_components, int32_t *counts,
                                            double *freq, double *effect1, double *effect2, double *pheno1,
                                            double *pheno2, int64_t *bad_index);

/* apply_heritability (source.cc:1008-1029) with the N noise draws made by the
 * caller's RNG.  pheno updated in place; effect (m entries, may be NULL)
 * scaled by gen_coef on the host side of the call. */
SIMU_B200_API int simu_b200_apply_heritability(simu_b200_ctx *ctx, float hsq, double *pheno, int64_t n,
                                               const double *noise, double *effect, int64_t m,
                                               double *gen_coef, double *env_coef);

/* Effect-size scaling between find_freq and find_pheno (source.cc:1245-1262, :1272-1286) and on
 * output (:1089-1094), per causal variant with het = fabs(2 f (1 - f)):
 *   op 0: effect *= sqrt(pow(het, S))   S = s_powK[component[j]]  (--gcta-sigma: S = -1; --traitK-s-pow)
 *   op 1: effect /= sqrt(het)           (--norm-effect on effects read from --causal-variants)
 *   op 2: effect *= sqrt(het)           (--norm-effect when writing .causals)
 * effect2 / s_pow2 / component may be NULL (one trait / one component).  If a variant has
 * het < FLT_EPSILON the call returns SIMU_B200_ERROR with its index in *bad_index (the reference
 * throws "Causal variant ... has zero frequency", :1251-1255); effects are then unspecified.
 * Ops 1 and 2 reproduce the host's bits (sqrt and division are correctly rounded on the device);
 * op 0 is within 2 ulp of glibc's pow. */
SIMU_B200_API int simu_b200_scale_effects(simu_b200_ctx *ctx, int op, const double *freq, double *effect1,
                                          double *effect2, int64_t mc, const int32_t *component,
                                          const double *s_pow1, const double *s_pow2, int num_components,
                                          int64_t *bad_index);

/* Sorted order of apply_liability_threshold (source.cc:1032-1034):
 * index[r] = sample of rank r, ascending pheno, ties by sample index. */
SIMU_B200_API int simu_b200_liability_order(simu_b200_ctx *ctx, const double *pheno, int64_t n,
                                            int32_t *index);

/* The two pools of apply_liability_threshold without the full sort: is_case[i] = 0 for the
 * max_ncon samples of lowest liability -- ranks [0, max_ncon) of the order above, the control
 * pool of source.cc:1042 -- and 1 for the case pool (:1043).  Enough whenever every sample of a
 * pool is labelled (the default --ncas / --ncon: source.cc:613-628), because the shuffles of
 * :1046-1047 then only consume RNG state; with fewer cases / controls requested the labels depend
 * on the order itself and simu_b200_liability_order is needed.  An 8-pass radix select on the
 * device instead of an 8-pass radix sort. */
SIMU_B200_API int simu_b200_liability_split(simu_b200_ctx *ctx, const double *pheno, int64_t n, int64_t max_ncon,
                                            uint8_t *is_case);

/* ------------------------------------------ hot path, device-resident variants
 * Same operations with DEVICE pointers, asynchronous on the context stream:
 * for callers that keep x, freq and y in HBM (multi-GPU partial sums followed
 * by an NCCL all-reduce, bench.py's device-resident timing). */

SIMU_B200_API int simu_b200_find_freq_dev(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc,
                                          int32_t *d_counts, double *d_freq);
SIMU_B200_API int simu_b200_find_pheno_dev(simu_b200_ctx *ctx, const int32_t *d_rows, int64_t mc,
                                           const double *d_freq, const double *d_effect1,
                                           const double *d_effect2, double *d_pheno1, double *d_pheno2,
                                           int mode);
/* d_s_pow: [2][num_components]; *d_bad_index (device) is set to UINT64_MAX or the first zero-het variant. */
SIMU_B200_API int simu_b200_find_freq_pheno_dev(simu_b200_ctx *ctx, int64_t mc, const double *d_raw1,
                                                const double *d_raw2, int scale_op, const int32_t *d_component,
                                                const double *d_s_pow, int num_components, int32_t *d_counts,
                                                double *d_freq, double *d_effect1, double *d_effect2,
                                                double *d_pheno1, double *d_pheno2, uint64_t *d_bad_index);
/* *d_out = sum(y[i]^2)  (source.cc:1011). */
SIMU_B200_API int simu_b200_sumsq_dev(simu_b200_ctx *ctx, const double *d_y, int64_t n, double *d_out);
/* gen/env from *d_sumsq exactly as source.cc:1012-1025, written to
 * d_coef[0..1]; then y = y*gen + noise*env (source.cc:1027). */
SIMU_B200_API int simu_b200_heritability_dev(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n,
                                             const double *d_noise, const double *d_sumsq, double *d_coef);
/* apply_heritability (source.cc:1008-1029) on device-resident data in ONE launch: sum of squares,
 * coefficients (written to d_coef[0..1] when not NULL) and y = y*gen + noise*env.  Same result as
 * simu_b200_sumsq_dev followed by simu_b200_heritability_dev up to the order of the sum of squares. */
SIMU_B200_API int simu_b200_apply_heritability_dev(simu_b200_ctx *ctx, float hsq, double *d_y, int64_t n,
                                                   const double *d_noise, double *d_coef);
SIMU_B200_API int simu_b200_liability_order_dev(simu_b200_ctx *ctx, const double *d_pheno, int64_t n,
                                                int32_t *d_index);
SIMU_B200_API int simu_b200_liability_split_dev(simu_b200_ctx *ctx, const double *d_pheno, int64_t n,
                                                int64_t max_ncon, uint8_t *d_is_case);
/* d_s_pow: [2][num_components] exponents (trait 1 then trait 2); *d_bad_index must hold
 * UINT64_MAX on entry and receives the smallest index with het < FLT_EPSILON, if any. */
SIMU_B200_API int simu_b200_scale_effects_dev(simu_b200_ctx *ctx, int op, const double *d_freq, double *d_effect1,
                                              double *d_effect2, int64_t mc, const int32_t *d_component,
                                              const double *d_s_pow, int num_components, uint64_t *d_bad_index);

/* ------------------------------------------------ multi-GPU exchange (one process per GPU)
 * The path's one collective -- the sum of the per-rank partial N x K phenotypes -- as
 * a one-shot all-reduce over NVLink peer memory: every rank publishes its partial in a symmetric
 * buffer, flags the peers, and sums all partials straight out of the peers' HBM in rank order
 * (identical bits on all ranks).  Latency-bound (0.8-8 MB): one cooperative kernel instead of a NCCL call.
 *   peer_init      allocates the symmetric buffer for `count` doubles (K * N) and returns its 64-byte
 *                  CUDA IPC handle, which the caller exchanges between the ranks by any means
 *                  (torch.distributed.all_gather_object, MPI, a file ...);
 *   peer_connect   maps the other ranks' buffers (handles: world x 64 bytes in rank order);
 *   peer_allreduce_dev  d_y[i] = sum over ranks of d_y[i], in place, asynchronous on the context
 *                  stream; every rank must call it the same number of times. */
SIMU_B200_API int simu_b200_peer_init(simu_b200_ctx *ctx, int rank, int world, int64_t count, void *handle_out);
SIMU_B200_API int simu_b200_peer_connect(simu_b200_ctx *ctx, const void *handles);
SIMU_B200_API int simu_b200_peer_allreduce_dev(simu_b200_ctx *ctx, double *d_y, int64_t count);
SIMU_B200_API void simu_b200_peer_close(simu_b200_ctx *ctx);

/* The missing-genotype index of a SAMPLE16 panel.  find_pheno's FAST kernel treats the missing genotypes -- the
 * reference's `continue` (source.cc:991) -- as a sparse correction when a panel is used more than once: the list
 * of missing entries (compressed by sample) is built at the SECOND FAST call after the panel changed (two scans of
 * the panel; it depends on the panel only, like the layout), the first call runs with the dense plane of the MMA;
 * from then on each call adds the entries' weights as exact 64-bit integers beside the main kernel.  *state: 0 = not
 * built (fewer than two FAST calls since the panel changed), 1 = sparse list in use (*entries = its length), 2 =
 * dense plane in use (more than 0.4 % missing genotypes, a panel of fewer than 2.5e8 genotypes, or SIMU_B200_TC_DENSE set).  Both forms
 * are the exact sum of the same quantised weights and agree to FP64 rounding (~1e-16 of max|y|). */
SIMU_B200_API int simu_b200_missing_index(const simu_b200_ctx *ctx, int *state, int64_t *entries);

/* Device time in milliseconds of the last find_freq / find_pheno call on this
 * context, measured with CUDA events on the launching stream (kernels only;
 * 0 if none).  which: 0 = find_freq, 1 = find_pheno. */
SIMU_B200_API int simu_b200_last_kernel_ms(simu_b200_ctx *ctx, int which, float *ms);

/* Per-kernel device timing for bench.py's roofline: when enabled, every launch of the hot
 * kernels is bracketed by a CUDA event pair on the launching stream (up to 512 pairs, then
 * recording stops).  kernel_id: 0 = allele counts, 1 = find_pheno EXACT, 2 = find_pheno FAST
 * main kernel (tensor-core kernel on SAMPLE16 panels, lookup-table kernel otherwise), 3 = its table / weight
 * build, 4 = its slot reduce, 5 = sumsq/heritability, 6 = sort.
 * profile_read synchronises the stream and returns the summed milliseconds and the number
 * of recorded launches; profile_enable resets the record. */
SIMU_B200_API int simu_b200_profile_enable(simu_b200_ctx *ctx, int on);
SIMU_B200_API int simu_b200_profile_read(simu_b200_ctx *ctx, int kernel_id, float *total_ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* SIMU_B200_H */
