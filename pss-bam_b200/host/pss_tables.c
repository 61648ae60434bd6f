/* pss_tables.c -- see pss_tables.h */
#include "pss_tables.h"

#include <limits.h>
#include <string.h>

/* column of the numerator of each printed rate and, implicitly (col & 3),
 * the reference base whose column total is the denominator */
static const int kRateCol[12] = { 1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14 };

void pss_sub_rates(const uint64_t *counts, int R, double *rates)
{
    for (int i = 0; i < R; i++) {
        const uint64_t *row = counts + (size_t)(i + 2) * 16;
        double *out = rates + (size_t)i * 12;
        double  per_ref[4];
        int     any_zero = 0;
        for (int ref = 0; ref < 4; ref++) {
            per_ref[ref] = (double)(row[ref] + row[4 + ref] + row[8 + ref] + row[12 + ref]);
            any_zero |= (per_ref[ref] == 0.0);
        }
        for (int j = 0; j < 12; j++)
            out[j] = any_zero ? 0.0 : (double)row[kRateCol[j]] / per_ref[kRateCol[j] & 3];   /* a row lacking any reference base stays 0 */
    }
}

static void put_count_row(FILE *fp, int label, const uint64_t *row)
{
    fprintf(fp, "%d\t", label);
    for (int j = 0; j < 16; j++) fprintf(fp, "%lu\t", (unsigned long)row[j]);
    fputc('\n', fp);
}

int pss_write_counts(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                     const uint64_t *fwd, const uint64_t *rev, int R)
{
    char fn[4096];
    snprintf(fn, sizeof fn, "%s.pss.counts.txt", out_prefix);
    FILE *fp = fopen(fn, "w");
    if (!fp) { fprintf(stderr, "ERROR: Cannot write to file %s\n.", fn); return 1; }
    fprintf(fp, "### pss-bam.c v" PSS_VERSION ":\n### FASTA: %s\n### BAM: %s\n### OUT: %s\n", fasta_fn, bam_fn, fn);
    fputs("### Format of table:\n"
          "### Counts of how often a read base and genome base were seen at\n"
          "### each position in the aligned reads.\n"
          "### First base is what was seen in the read.\n"
          "### Second base is what was in the genome at that position.\n"
          "### POS AA AC AG AT CA CC CG CT GA GC GG GT TA TC TG TT\n"
          "### Forward read substitution counts and base context\n", fp);
    for (int i = -2; i < R; i++) put_count_row(fp, i, fwd + (size_t)(i + 2) * 16);
    fputs("\n\n### Reverse read substitution counts and base context\n", fp);   /* blank lines: gnuplot block separator */
    for (int i = R - 1; i >= 0; i--) put_count_row(fp, i, rev + (size_t)(i + 2) * 16);
    put_count_row(fp, 1, rev + 16);     /* context base adjacent to the 3' end */
    put_count_row(fp, 2, rev);          /* context base two away */
    fclose(fp);
    return 0;
}

int pss_write_rates(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                    const double *fr, const double *rr, int R)
{
    char fn[4096];
    snprintf(fn, sizeof fn, "%s.pss.rates.txt", out_prefix);
    FILE *fp = fopen(fn, "w");
    if (!fp) { fprintf(stderr, "ERROR: Cannot write to file %s\n.", fn); return 1; }
    fprintf(fp, "### pss-bam.c v" PSS_VERSION "\n### FASTA: %s\n### BAM: %s\n### OUT: %s\n", fasta_fn, bam_fn, fn);
    fputs("### Format of table:\n"
          "### Substitution rates for all possible nucleotide substitutions at\n"
          "### each position in the aligned reads.\n"
          "### First base is what was seen in the read.\n"
          "### Second base is what was in the genome at that position.\n"
          "### POS AC AG AT CA CG CT GA GC GT TA TC TG\n"
          "### Forward read substitution rates\n", fp);
    for (int i = 0; i < R; i++) {
        fprintf(fp, "%d\t", i);
        for (int j = 0; j < 12; j++) fprintf(fp, "%.5e\t", fr[(size_t)i * 12 + j]);
        fputc('\n', fp);
    }
    fputs("\n\n### Reverse read substitution rates\n", fp);
    for (int i = R - 1; i >= 0; i--) {
        fprintf(fp, "%d\t", i);
        for (int j = 0; j < 12; j++) fprintf(fp, "%.5e\t", rr[(size_t)i * 12 + j]);
        fputc('\n', fp);
    }
    fclose(fp);
    return 0;
}

static void kmer_text(uint64_t inx, int k, char *out)
{
    out[k] = '\0';
    for (int i = k - 1; i >= 0; i--, inx >>= 2) out[i] = "ACGT"[inx & 3];
}
static unsigned int sat32(uint64_t v) { return v > UINT_MAX ? UINT_MAX : (unsigned int)v; }

int pss_write_fragkon(FILE *out, const char *fasta_fn, const char *bam_fn, int klen, const uint64_t *fp, const uint64_t *tp)
{
    char km[64];
    const uint64_t n = 1ull << (2 * klen);
    fprintf(out, "### fragkon.c v0.3\n### %s\n### %s\n", fasta_fn, bam_fn);
    fprintf(out, "# KMER\t5' CONTEXT COUNTS\t3' CONTEXT COUNTS\n");
    for (uint64_t i = 0; i < n; i++) {
        kmer_text(i, klen, km);
        fprintf(out, "%s\t%u\t%u\n", km, sat32(fp[i]), sat32(tp[i]));
    }
    return 0;
}

int pss_write_spectrum(FILE *out, size_t n_seqs, int k, const uint64_t *counts)
{
    char km[64];
    const uint64_t n = 1ull << (2 * k);
    fprintf(out, "Parsed input genome. Found %lu sequences.\n", (unsigned long)n_seqs);
    for (uint64_t i = 0; i < n; i++) {
        kmer_text(i, k, km);
        fprintf(out, "%s\t%u\n", km, sat32(counts[i]));
    }
    return 0;
}
