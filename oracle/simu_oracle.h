/*
 * simu_oracle.h -- CPU restatement of the precimed/simu phenotype-synthesis
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library.  The product path
 * (simu_b200/, include/simu_b200.h) never links, imports or calls it.
 *
 * Parity status: the 2-bit decode and BED header handling are pinned against
 * the reference's own known-answer vectors (libplinkio/tests/bed_test.c) and
 * against the reference's compiled libplinkio (oracle/_ref, see Makefile).
 * find_freq / find_pheno / apply_* have NO golden vectors upstream (the
 * reference ships no tests for src/source.cc); they are pinned instead against
 * the reference ITSELF: the unmodified src/source.cc compiled against
 * oracle/boost_shim/ into oracle/_ref/libsimu_refsrc.so, whose literal functions
 * tests/test_reference_source.py compares with this restatement bit for bit.
 * What stays unpinned is Boost.Random / program_options behaviour, which only a
 * real Boost build could confirm.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef SIMU_ORACLE_H
#define SIMU_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* libplinkio/src/plinkio/status.h:16-47 */
enum { ORACLE_OK = 0, ORACLE_END = 1, ORACLE_ERROR = 2 };

/* BED header parse.  libplinkio/src/bed_header.c:80-102 (get_version_and_order),
 * :56-71 (get_data_offset).
 * version: 100 = v1.00, 99 = v0.99, 98 = pre-0.99.
 * snp_order: 1 = one locus per row (SNP-major), 0 = one sample per row,
 *            -1 = unknown.
 * Returns the data offset in bytes (3, 1 or 0). */
int oracle_bed_header(const unsigned char header[3], int *version, int *snp_order);

/* Packed bytes per row.  libplinkio/src/bed_header.c:219-223. */
size_t oracle_bed_row_bytes(size_t num_samples);

/* 2-bit decode of one packed row into num_cols one-byte codes.
 * libplinkio/src/bed.c:91-114 + plinkio/snp_lookup_little.h:9-266:
 * bits 00->0 (hom A1), 01->3 (missing), 10->1 (het), 11->2 (hom A2),
 * lowest bit pair first. */
void oracle_unpack_snps(const unsigned char *packed, unsigned char *unpacked, size_t num_cols);

/* Inverse of the above.  libplinkio/src/bed.c:124-150 + snp_lookup.h:27. */
void oracle_pack_snps(const unsigned char *unpacked, unsigned char *packed, size_t num_cols);

/* find_freq over `mc` causal rows.  src/source.cc:843-885.
 * rows:      base of packed rows, row r at rows + r*row_stride.
 * row_index: optional gather list (NULL = rows 0..mc-1).
 * counts:    mc x 4 ints, order genocount[0..3] (source.cc:869).
 * freq:      mc doubles, source.cc:874-875.  Either output may be NULL. */
void oracle_find_freq(const unsigned char *rows, size_t row_stride,
                      const int64_t *row_index, size_t mc, size_t num_samples,
                      int32_t *counts, double *freq);

/* find_pheno over `mc` causal rows in ascending order.  src/source.cc:951-1006.
 * effect2 / pheno2 may be NULL (single trait).  pheno arrays are cleared first
 * (source.cc:962-966). */
void oracle_find_pheno(const unsigned char *rows, size_t row_stride,
                       const int64_t *row_index, size_t mc, size_t num_samples,
                       const double *freq, const double *effect1, const double *effect2,
                       double *pheno1, double *pheno2);

/* find_pheno restricted to samples [s0, s1), s0 % 4 == 0 (same per-sample operation order, hence the
 * same bits; used to split the sample loop over worker threads and for column subsets). */
void oracle_find_pheno_range(const unsigned char *rows, size_t row_stride,
                             const int64_t *row_index, size_t mc, size_t s0, size_t s1,
                             const double *freq, const double *effect1, const double *effect2,
                             double *pheno1, double *pheno2);

/* apply_heritability with the N noise draws supplied by the caller
 * (draws stay in the host RNG module).  src/source.cc:1008-1029.
 * Returns gen_coef; *env_out receives env_coef. */
double oracle_apply_heritability(float hsq, double *pheno, size_t n, const double *noise,
                                 double *effect, size_t m, double *env_out);

/* Sorted order used by apply_liability_threshold.  src/source.cc:1032-1034.
 * index[r] = sample with rank r (ascending pheno).  Ties are ordered by
 * sample index here; the reference's std::sort leaves tie order unspecified. */
void oracle_liability_order(const double *pheno, size_t n, int32_t *index);

/* Label assignment of apply_liability_threshold given the sorted order and
 * the two already-shuffled rank pools.  src/source.cc:1037-1054.
 * con_inds has max_ncon entries, cas_inds has n-max_ncon entries.
 * Returns 0, or -1/-2 for the "too many controls/cases" errors (:1049-1050). */
int oracle_liability_labels(float k, int ncas, int ncon, size_t n, const int32_t *index,
                            const int32_t *con_inds, const int32_t *cas_inds, double *pheno);

/* max_ncas / max_ncon split.  src/source.cc:1037-1039. */
void oracle_liability_split(float k, size_t n, int *max_ncas, int *max_ncon);

/* Effect-size scaling by het^S / sqrt(het).  src/source.cc:1245-1262 (op 0), :1272-1286 (op 1),
 * :1089-1094 (op 2).  Returns -1 or the first variant whose het < FLT_EPSILON (ops 0, 1). */
long oracle_scale_effects(int op, const double *freq, double *effect1, double *effect2, size_t mc,
                          const int32_t *component, const double *s_pow1, const double *s_pow2);

/* ---- synthetic panel generator; integer-only so that the
 * CUDA generator in simu_b200/csrc/synth.cu produces identical bytes. ---- */
void oracle_synth_rows(uint64_t panel_seed, int64_t first_row, size_t n_rows,
                       size_t num_samples, unsigned char *rows, size_t row_stride);

/* Sample columns [s0, s1) (s0 % 4 == 0) of the same rows, as a panel of s1 - s0 samples. */
void oracle_synth_cols(uint64_t panel_seed, int64_t first_row, const int64_t *row_ids, size_t n_rows,
                       size_t s0, size_t s1, unsigned char *rows, size_t row_stride);

#ifdef __cplusplus
}
#endif
#endif
