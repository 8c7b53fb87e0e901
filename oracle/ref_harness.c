/*
 * ref_harness.c -- thin driver around the REFERENCE's own libplinkio,
 * compiled from /root/reference/libplinkio where it lies (see Makefile).
 * TEST INFRASTRUCTURE ONLY: used to validate oracle/simu_oracle.c and as the
 * I/O path of the CPU baseline.  Nothing here is product code and nothing in
 * simu_b200/ links it.
 *
 * The loops below are the caller side of libplinkio as src/source.cc uses it
 * (PioFile / PioFiles, src/source.cc:187-361): pio_open, pio_skip_row for
 * non-causal rows, pio_next_row for causal rows, pio_reset_row at the end.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <plinkio/plinkio.h>

/* Number of samples / loci of a bfile, or -1.  (src/source.cc:189-210) */
int ref_dims(const char *prefix, int64_t *num_samples, int64_t *num_loci, int *locus_per_row)
{
    struct pio_file_t f;
    if (pio_open(&f, prefix) != PIO_OK) return -1;
    *num_samples = (int64_t)pio_num_samples(&f);
    *num_loci = (int64_t)pio_num_loci(&f);
    *locus_per_row = pio_one_locus_per_row(&f);
    pio_close(&f);
    return 0;
}

/* Unpack every row with pio_next_row into out[num_loci][num_samples]. */
int ref_read_all(const char *prefix, unsigned char *out)
{
    struct pio_file_t f;
    size_t n, r = 0;
    if (pio_open(&f, prefix) != PIO_OK) return -1;
    n = pio_row_size(&f);
    while (pio_next_row(&f, out + r * n) == PIO_OK) r++;
    pio_close(&f);
    return (int)r;
}

/* find_freq + find_pheno exactly as src/source.cc:843-885 and :951-1006 drive
 * libplinkio over ONE bfile: a row cursor that skips non-causal rows and
 * unpacks causal rows.  component[v] == -1 marks a non-causal variant.
 * freq/effect arrays have one entry per variant (as in the reference).
 * pass: 1 = find_freq only, 2 = find_pheno only, 3 = both.
 * Returns 0, or -1 open error, -2 stream shorter than num_variants,
 * -3 stream longer than num_variants, -4 read error. */
int ref_freq_pheno(const char *prefix, int64_t num_variants, const int32_t *component,
                   int32_t *counts /* num_variants x 4 or NULL */, double *freq,
                   const double *effect1, const double *effect2,
                   double *pheno1, double *pheno2, int pass)
{
    struct pio_file_t f;
    snp_t *buffer;
    int64_t v, s, ns;
    int rc = 0;
    if (pio_open(&f, prefix) != PIO_OK) return -1;
    ns = (int64_t)pio_num_samples(&f);
    buffer = (snp_t *)malloc(pio_row_size(&f) ? pio_row_size(&f) : 1);

    if (pass & 1) {
        for (v = 0; v < num_variants; v++) {
            int genocount[4] = { 0, 0, 0, 0 }, nm;
            pio_status_t st;
            freq[v] = 0.0 / 0.0; /* NaN, source.cc:850 */
            if (component[v] == -1) {
                if (pio_skip_row(&f) == PIO_END) { rc = -2; goto done; }
                continue;
            }
            st = pio_next_row(&f, buffer);
            if (st == PIO_END) { rc = -2; goto done; }
            if (st == PIO_ERROR) { rc = -4; goto done; }
            for (s = 0; s < ns; s++) genocount[(int)buffer[s]]++;
            nm = genocount[2] + genocount[1] + genocount[0];
            freq[v] = (nm == 0) ? 0.0 : (double)(2 * genocount[0] + 1 * genocount[1]) / (double)(2 * nm);
            if (counts) memcpy(counts + 4 * v, genocount, sizeof(genocount));
        }
        if (pio_skip_row(&f) != PIO_END) { rc = -3; goto done; }
        pio_reset_row(&f);
    }

    if (pass & 2) {
        const int bivariate = (effect2 != NULL && pheno2 != NULL);
        for (s = 0; s < ns; s++) { pheno1[s] = 0; if (bivariate) pheno2[s] = 0; }
        for (v = 0; v < num_variants; v++) {
            double avg;
            pio_status_t st;
            if (component[v] == -1) {
                if (pio_skip_row(&f) == PIO_END) { rc = -2; goto done; }
                continue;
            }
            st = pio_next_row(&f, buffer);
            if (st == PIO_END) { rc = -2; goto done; }
            if (st == PIO_ERROR) { rc = -4; goto done; }
            avg = 2 * freq[v];
            for (s = 0; s < ns; s++) {
                double a;
                if (buffer[s] == 3) continue;
                a = (double)(2 - buffer[s]);
                pheno1[s] += (a - avg) * effect1[v];
                if (bivariate) pheno2[s] += (a - avg) * effect2[v];
            }
        }
        if (pio_skip_row(&f) != PIO_END) { rc = -3; goto done; }
        pio_reset_row(&f);
    }

done:
    free(buffer);
    pio_close(&f);
    return rc;
}

/* Locus / sample accessors so tests can check the .bim/.fam parse of the
 * reference against the product's own parser. */
int ref_locus(const char *prefix, int64_t id, int *chromosome, char *name, int name_cap,
              long long *bp, char *a1, char *a2, int allele_cap)
{
    struct pio_file_t f;
    struct pio_locus_t *l;
    if (pio_open(&f, prefix) != PIO_OK) return -1;
    l = pio_get_locus(&f, (size_t)id);
    if (!l) { pio_close(&f); return -2; }
    *chromosome = (int)l->chromosome;
    *bp = l->bp_position;
    strncpy(name, l->name, (size_t)name_cap - 1); name[name_cap - 1] = 0;
    strncpy(a1, l->allele1, (size_t)allele_cap - 1); a1[allele_cap - 1] = 0;
    strncpy(a2, l->allele2, (size_t)allele_cap - 1); a2[allele_cap - 1] = 0;
    pio_close(&f);
    return 0;
}

int ref_sample(const char *prefix, int64_t id, char *fid, char *iid, int cap)
{
    struct pio_file_t f;
    struct pio_sample_t *s;
    if (pio_open(&f, prefix) != PIO_OK) return -1;
    s = pio_get_sample(&f, (size_t)id);
    if (!s) { pio_close(&f); return -2; }
    strncpy(fid, s->fid, (size_t)cap - 1); fid[cap - 1] = 0;
    strncpy(iid, s->iid, (size_t)cap - 1); iid[cap - 1] = 0;
    pio_close(&f);
    return 0;
}
