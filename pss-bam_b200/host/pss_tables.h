/* pss_tables.h -- host-side table arithmetic and writers.
 *
 * The GPU hands back plain counter arrays; everything the plotting scripts
 * read (pss-bam-plot.py, pss-bam-gnuplot-template.gp) is produced here, byte
 * for byte in the reference's format:
 *   <prefix>.pss.counts.txt   pss-bam.c:538-586 (print_counts)
 *   <prefix>.pss.rates.txt    pss-bam.c:595-633 (print_rates), rates from :504-529
 *   fragkon stdout table      fragkon.c:367-369, :231-249
 *   genome-kmer-count stdout  genome-kmer-count.c:52-53, :61-64
 */
#ifndef PSS_TABLES_H
#define PSS_TABLES_H

#include <stdint.h>
#include <stdio.h>

#define PSS_VERSION "1.2.1"

/* counts: (R+2) x 16 row-major; rates: R x 12, order AC AG AT CA CG CT GA GC GT TA TC TG */
void pss_sub_rates(const uint64_t *counts, int region_len, double *rates);
int  pss_write_counts(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                      const uint64_t *fwd, const uint64_t *rev, int region_len);
int  pss_write_rates(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                     const double *fwd_rates, const double *rev_rates, int region_len);
/* 4^k rows "KMER\t5'\t3'", counts clamped to UINT_MAX like the reference's unsigned int counters */
int  pss_write_fragkon(FILE *out, const char *fasta_fn, const char *bam_fn, int klen,
                       const uint64_t *fp, const uint64_t *tp);
int  pss_write_spectrum(FILE *out, size_t n_seqs, int k, const uint64_t *counts);

#endif
