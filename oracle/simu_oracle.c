/*
 * simu_oracle.c -- CPU restatement of the precimed/simu phenotype-synthesis
 * hot path.  TEST INFRASTRUCTURE ONLY (see simu_oracle.h for the rules and the
 * parity status of each function).
 *
 * The arithmetic follows /root/reference/src/source.cc statement by statement:
 * same loop order, same operand types, separate subtract / multiply / add in
 * IEEE double (the reference builds with -std=c++11 and no -march, so there is
 * no FMA contraction on x86-64; this file must be compiled with
 * -ffp-contract=off to keep that property on any host).
 */
#include "simu_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ header */

/* libplinkio/src/bed_header.c:14-15 */
#define ORACLE_BED_MAGIC1 0x6c
#define ORACLE_BED_MAGIC2 0x1b

/* libplinkio/src/bed_header.c:31-46 */
static int oracle_snp_order_of(unsigned char order)
{
    if (order == 0x01) return 1;
    if (order == 0) return 0;
    return -1;
}

int oracle_bed_header(const unsigned char header[3], int *version, int *snp_order)
{
    /* libplinkio/src/bed_header.c:80-102 */
    if (header[0] == ORACLE_BED_MAGIC1 && header[1] == ORACLE_BED_MAGIC2) {
        *version = 100;
        *snp_order = oracle_snp_order_of(header[2]);
        return 3; /* bed_header.c:56-60 */
    }
    if ((header[0] & ~0x01) == 0) {
        *version = 99;
        *snp_order = oracle_snp_order_of(header[0]);
        return 1; /* bed_header.c:61-64 */
    }
    *version = 98;
    *snp_order = 0;
    return 0;
}

size_t oracle_bed_row_bytes(size_t num_samples)
{
    return (num_samples + 3) / 4; /* bed_header.c:219-223 */
}

/* ------------------------------------------------------------------ decode */

/* plinkio/snp_lookup_little.h:9-266 written as a rule instead of a table:
 * packed bit pair -> unpacked code. */
static const unsigned char oracle_bits_to_snp[4] = { 0, 3, 1, 2 };
/* plinkio/snp_lookup.h:27 */
static const unsigned char oracle_snp_to_bits[4] = { 0, 2, 3, 1 };

void oracle_unpack_snps(const unsigned char *packed, unsigned char *unpacked, size_t num_cols)
{
    /* bed.c:91-114: whole bytes first, then the trailing num_cols % 4 codes;
     * inside a byte the SNPs are read "from right to left" (bed.c:73-74). */
    size_t full = num_cols / 4, left = num_cols % 4, i, k;
    for (i = 0; i < full; i++) {
        unsigned char b = packed[i];
        for (k = 0; k < 4; k++)
            unpacked[4 * i + k] = oracle_bits_to_snp[(b >> (2 * k)) & 3];
    }
    for (k = 0; k < left; k++)
        unpacked[4 * full + k] = oracle_bits_to_snp[(packed[full] >> (2 * k)) & 3];
}

void oracle_pack_snps(const unsigned char *unpacked, unsigned char *packed, size_t num_cols)
{
    /* bed.c:124-150 */
    size_t nbytes = oracle_bed_row_bytes(num_cols), i;
    memset(packed, 0, nbytes);
    for (i = 0; i < num_cols; i++)
        packed[i / 4] |= (unsigned char)(oracle_snp_to_bits[unpacked[i] & 3] << (2 * (i % 4)));
}

/* --------------------------------------------------------------- find_freq */

void oracle_find_freq(const unsigned char *rows, size_t row_stride,
                      const int64_t *row_index, size_t mc, size_t num_samples,
                      int32_t *counts, double *freq)
{
    unsigned char *buffer = (unsigned char *)malloc(num_samples ? num_samples : 1);
    size_t j, s;
    for (j = 0; j < mc; j++) {
        const unsigned char *row = rows + (size_t)(row_index ? row_index[j] : (int64_t)j) * row_stride;
        int genocount[4] = { 0, 0, 0, 0 };       /* source.cc:869 */
        int genocount_non_missing;
        oracle_unpack_snps(row, buffer, num_samples); /* bed.c:355 via source.cc:863 */
        for (s = 0; s < num_samples; s++)        /* source.cc:870-872 */
            genocount[(int)buffer[s]]++;
        genocount_non_missing = genocount[2] + genocount[1] + genocount[0]; /* :874 */
        if (counts) {
            counts[4 * j + 0] = genocount[0];
            counts[4 * j + 1] = genocount[1];
            counts[4 * j + 2] = genocount[2];
            counts[4 * j + 3] = genocount[3];
        }
        if (freq)                                /* source.cc:875 */
            freq[j] = (genocount_non_missing == 0)
                          ? 0.0
                          : (double)(2 * genocount[0] + 1 * genocount[1]) /
                                (double)(2 * genocount_non_missing);
    }
    free(buffer);
}

/* -------------------------------------------------------------- find_pheno */

void oracle_find_pheno(const unsigned char *rows, size_t row_stride,
                       const int64_t *row_index, size_t mc, size_t num_samples,
                       const double *freq, const double *effect1, const double *effect2,
                       double *pheno1, double *pheno2)
{
    const int bivariate = (effect2 != NULL && pheno2 != NULL); /* source.cc:959 */
    unsigned char *buffer = (unsigned char *)malloc(num_samples ? num_samples : 1);
    size_t j, s;
    for (s = 0; s < num_samples; s++) {          /* source.cc:962-966 */
        pheno1[s] = 0;
        if (bivariate) pheno2[s] = 0;
    }
    for (j = 0; j < mc; j++) {                   /* causal rows, ascending: :968-984 */
        const unsigned char *row = rows + (size_t)(row_index ? row_index[j] : (int64_t)j) * row_stride;
        double average_A1_dosage = 2 * freq[j];  /* source.cc:986 */
        oracle_unpack_snps(row, buffer, num_samples);
        for (s = 0; s < num_samples; s++) {      /* source.cc:987-996 */
            double num_of_A1_copies;
            if (buffer[s] == 3) continue;        /* :991 */
            num_of_A1_copies = (double)(2 - buffer[s]); /* :992 */
            pheno1[s] += (num_of_A1_copies - average_A1_dosage) * effect1[j];        /* :994 */
            if (bivariate)
                pheno2[s] += (num_of_A1_copies - average_A1_dosage) * effect2[j];    /* :995 */
        }
    }
    free(buffer);
}

/* The same loop restricted to samples [s0, s1), s0 a multiple of 4: every sample's sum is formed in
 * the same ascending-SNP order with the same operations as source.cc:986-996, so the values are
 * bit-identical to oracle_find_pheno's; only the sample loop is split (for worker threads and for
 * column subsets of panels too large to process whole).  pheno1/pheno2 are indexed by absolute sample. */
void oracle_find_pheno_range(const unsigned char *rows, size_t row_stride,
                             const int64_t *row_index, size_t mc, size_t s0, size_t s1,
                             const double *freq, const double *effect1, const double *effect2,
                             double *pheno1, double *pheno2)
{
    const int bivariate = (effect2 != NULL && pheno2 != NULL); /* source.cc:959 */
    const size_t ns = s1 > s0 ? s1 - s0 : 0;
    unsigned char *buffer = (unsigned char *)malloc(ns ? ns : 1);
    size_t j, s;
    for (s = s0; s < s1; s++) {                  /* source.cc:962-966 */
        pheno1[s] = 0;
        if (bivariate) pheno2[s] = 0;
    }
    for (j = 0; j < mc; j++) {                   /* causal rows, ascending: :968-984 */
        const unsigned char *row = rows + (size_t)(row_index ? row_index[j] : (int64_t)j) * row_stride;
        double average_A1_dosage = 2 * freq[j];  /* source.cc:986 */
        oracle_unpack_snps(row + s0 / 4, buffer, ns);
        for (s = 0; s < ns; s++) {               /* source.cc:987-996 */
            double num_of_A1_copies;
            if (buffer[s] == 3) continue;        /* :991 */
            num_of_A1_copies = (double)(2 - buffer[s]); /* :992 */
            pheno1[s0 + s] += (num_of_A1_copies - average_A1_dosage) * effect1[j];        /* :994 */
            if (bivariate)
                pheno2[s0 + s] += (num_of_A1_copies - average_A1_dosage) * effect2[j];    /* :995 */
        }
    }
    free(buffer);
}

/* ------------------------------------------------------ apply_heritability */

double oracle_apply_heritability(float hsq, double *pheno, size_t n, const double *noise,
                                 double *effect, size_t m, double *env_out)
{
    double pheno_var = 0.0;                      /* source.cc:1009 */
    double nn = (double)n;                       /* :1010 */
    double gen_coef = 0, env_coef = 0;           /* :1018 */
    size_t i;
    for (i = 0; i < n; i++) pheno_var += (pheno[i] * pheno[i]); /* :1011 */
    pheno_var = pheno_var / nn;                  /* :1012 */
    if ((fabs(hsq) <= FLT_EPSILON) || (fabs(pheno_var) <= FLT_EPSILON)) { /* :1019 */
        gen_coef = 0; env_coef = 1;              /* :1021 */
    } else {
        gen_coef = sqrt(hsq / pheno_var);        /* :1023, float promoted to double */
        env_coef = sqrt(1.0 - hsq);              /* :1024 */
    }
    for (i = 0; i < n; i++) pheno[i] = pheno[i] * gen_coef + noise[i] * env_coef; /* :1027 */
    if (effect) for (i = 0; i < m; i++) effect[i] *= gen_coef;                    /* :1028 */
    if (env_out) *env_out = env_coef;
    return gen_coef;
}

/* ----------------------------------------------- apply_liability_threshold */

typedef struct { double key; int32_t idx; } oracle_sort_item;

static int oracle_sort_cmp(const void *a, const void *b)
{
    const oracle_sort_item *x = (const oracle_sort_item *)a, *y = (const oracle_sort_item *)b;
    if (x->key < y->key) return -1;              /* source.cc:1034 comparator */
    if (y->key < x->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* ties: by sample index */
}

void oracle_liability_order(const double *pheno, size_t n, int32_t *index)
{
    oracle_sort_item *items = (oracle_sort_item *)malloc((n ? n : 1) * sizeof(*items));
    size_t i;
    for (i = 0; i < n; i++) { items[i].key = pheno[i]; items[i].idx = (int32_t)i; } /* :1032-1033 */
    qsort(items, n, sizeof(*items), oracle_sort_cmp);                                /* :1034 */
    for (i = 0; i < n; i++) index[i] = items[i].idx;
    free(items);
}

void oracle_liability_split(float k, size_t n, int *max_ncas, int *max_ncon)
{
    int nn = (int)n;                             /* source.cc:1037 */
    int cas = (int)floor(k * nn);                /* :1038  float * int -> float, then floor() */
    *max_ncas = cas;
    *max_ncon = nn - cas;                        /* :1039 */
}

int oracle_liability_labels(float k, int ncas, int ncon, size_t n, const int32_t *index,
                            const int32_t *con_inds, const int32_t *cas_inds, double *pheno)
{
    int max_ncas, max_ncon, i;
    oracle_liability_split(k, n, &max_ncas, &max_ncon);
    if (max_ncon < ncon) return -1;              /* :1049 */
    if ((int)n - max_ncon < ncas) return -2;     /* :1050 */
    for (i = 0; i < (int)n; i++) pheno[i] = -9;  /* :1052 */
    for (i = 0; i < ncon; i++) pheno[index[con_inds[i]]] = 1; /* :1053 */
    for (i = 0; i < ncas; i++) pheno[index[cas_inds[i]]] = 2; /* :1054 */
    return 0;
}

/* ------------------------------------------------------- effect-size scaling */

/* The three per-variant effect scalings main() applies around find_pheno:
 * op 0  source.cc:1245-1262  effect *= sqrt(pow(het, S)), S = -1 under --gcta-sigma, else
 *                            --traitK-s-pow of the variant's component;
 * op 1  source.cc:1272-1286  effect /= sqrt(het)   (--norm-effect, effects from a file);
 * op 2  source.cc:1089-1094  effect *= sqrt(het)   (--norm-effect, .causals output);
 * het = fabs(2 f (1-f)) (:1250, :1277, :1092).  Returns -1, or the index of the first variant
 * with het < FLT_EPSILON (:1251-1255, :1278-1282: the reference throws there; op 2 has no check). */
long oracle_scale_effects(int op, const double *freq, double *effect1, double *effect2, size_t mc,
                          const int32_t *component, const double *s_pow1, const double *s_pow2)
{
    size_t j;
    for (j = 0; j < mc; j++) {
        const double f = freq[j];
        const double het = fabs(2 * f * (1 - f));
        const int c = component ? component[j] : 0;
        if (op != 2 && het < FLT_EPSILON) return (long)j;
        if (op == 0) {
            effect1[j] *= sqrt(pow(het, s_pow1[c]));             /* :1258 */
            if (effect2) effect2[j] *= sqrt(pow(het, s_pow2[c])); /* :1261 */
        } else if (op == 1) {
            effect1[j] /= sqrt(het);                             /* :1284 */
            if (effect2) effect2[j] /= sqrt(het);                /* :1285 */
        } else {
            effect1[j] *= sqrt(het);                             /* :1093 */
            if (effect2) effect2[j] *= sqrt(het);
        }
    }
    return -1;
}

/* ------------------------------------------------------ synthetic panel gen */

/* splitmix64 finaliser */
static uint64_t oracle_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_synth_rows(uint64_t panel_seed, int64_t first_row, size_t n_rows,
                       size_t num_samples, unsigned char *rows, size_t row_stride)
{
    /* p_j ~ U(0.01, 0.5); A1 count ~ Binomial(2, p_j);
     * missing with probability 0.001.  Integer thresholds only:
     *   thr_j   = 10486 + floor(hi32(key_j) * 513802 / 2^32)   (20-bit scale)
     *   h       = mix64(key_j + i)
     *   a1      = (h[0:20) < thr_j) + (h[20:40) < thr_j)
     *   missing = h[40:64) < 16777                              (24-bit scale)
     * Row padding bits are 00 as PLINK writes them. */
    const size_t row_bytes = oracle_bed_row_bytes(num_samples);
    long long r;
    for (r = 0; r < (long long)n_rows; r++) {
        unsigned char *row = rows + (size_t)r * row_stride;
        uint64_t key = oracle_mix64(panel_seed ^ oracle_mix64((uint64_t)(first_row + r)));
        uint32_t thr = 10486u + (uint32_t)(((key >> 32) * 513802ull) >> 32);
        size_t i;
        memset(row, 0, row_bytes);
        for (i = 0; i < num_samples; i++) {
            uint64_t h = oracle_mix64(key + (uint64_t)i);
            uint32_t u1 = (uint32_t)(h & 0xFFFFF), u2 = (uint32_t)((h >> 20) & 0xFFFFF);
            uint32_t u3 = (uint32_t)(h >> 40);
            unsigned a1 = (u1 < thr) + (u2 < thr);
            /* packed bits: 2 copies -> 00, 1 -> 10, 0 -> 11, missing -> 01 */
            unsigned bits = (u3 < 16777u) ? 1u : (a1 == 2 ? 0u : (a1 == 1 ? 2u : 3u));
            row[i >> 2] |= (unsigned char)(bits << (2 * (i & 3)));
        }
    }
}

/* Columns [s0, s1) (s0 % 4 == 0) of the same synthetic rows, packed as rows of a panel that holds only
 * those samples: byte b of an output row equals byte s0/4 + b of the full row (its last byte masked to
 * s1).  row_ids (optional) lists the SNP rows, otherwise first_row + r. */
void oracle_synth_cols(uint64_t panel_seed, int64_t first_row, const int64_t *row_ids, size_t n_rows,
                       size_t s0, size_t s1, unsigned char *rows, size_t row_stride)
{
    const size_t ns = s1 - s0, row_bytes = oracle_bed_row_bytes(ns);
    long long r;
    for (r = 0; r < (long long)n_rows; r++) {
        unsigned char *row = rows + (size_t)r * row_stride;
        uint64_t key = oracle_mix64(panel_seed ^ oracle_mix64((uint64_t)(row_ids ? row_ids[r] : first_row + r)));
        uint32_t thr = 10486u + (uint32_t)(((key >> 32) * 513802ull) >> 32);
        size_t i;
        memset(row, 0, row_bytes);
        for (i = 0; i < ns; i++) {
            uint64_t h = oracle_mix64(key + (uint64_t)(s0 + i));
            uint32_t u1 = (uint32_t)(h & 0xFFFFF), u2 = (uint32_t)((h >> 20) & 0xFFFFF);
            uint32_t u3 = (uint32_t)(h >> 40);
            unsigned a1 = (u1 < thr) + (u2 < thr);
            unsigned bits = (u3 < 16777u) ? 1u : (a1 == 2 ? 0u : (a1 == 1 ? 2u : 3u));
            row[i >> 2] |= (unsigned char)(bits << (2 * (i & 3)));
        }
    }
}
