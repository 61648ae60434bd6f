/* oracle_pss.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the algorithms on the pss-bam hot path, used only
 * as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product path (pss-bam_b200/csrc, libpssgpu.so) never
 * links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py runs this restatement
 * against the unmodified reference binaries (oracle/_ref, built by
 * `make -C oracle ref` from /root/reference) where /root/reference exists, and
 * against the committed golden fixtures (tests/golden/, generated from those
 * binaries by tests/golden/make_golden.py) everywhere else.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef ORACLE_PSS_H
#define ORACLE_PSS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- genome (fasta-genome-io.h:13-32) ---------------------------------- */
typedef struct ora_contig {
    char  *id;   /* chars after '>' up to first isspace()            */
    char  *seq;  /* every non-space char, upper-cased, NUL-terminated */
    size_t len;
} ora_contig;

typedef struct ora_genome {
    ora_contig *ctg;   /* sorted by strcmp(id) (fasta-genome-io.c:236) */
    size_t      n;
} ora_genome;

/* fasta-genome-io.c:221-238 init_genome + :105-200 read_fasta/gzread_fasta */
ora_genome *ora_genome_load(const char *fasta_path);
/* Same parse over an in-memory FASTA text. */
ora_genome *ora_genome_parse(const char *text, size_t len);
/* Same from contigs already split out of their FASTA text (upper-cases, sorts). */
ora_genome *ora_genome_from_contigs(const char *const *ids, const char *const *seqs, const size_t *lens, size_t n);
void        ora_genome_free(ora_genome *g);
/* fasta-genome-io.c:202-213 find_seq; returns index or -1 */
long        ora_find_contig(const ora_genome *g, const char *id);

/* ---- pss-bam (pss-bam.c) ------------------------------------------------ */
typedef struct ora_pss_params {
    int           region_len;   /* -r REGION_LEN    pss-bam.c:12  */
    unsigned long min_len;      /* -l MIN_READ_LEN  pss-bam.c:13  */
    unsigned long max_len;      /* -L MAX_READ_LEN  pss-bam.c:14  */
    int           min_mq;       /* -q MIN_MQ        pss-bam.c:15  */
    const char   *up_ctx;       /* -U UP_CTX        pss-bam.c:16  */
    const char   *down_ctx;     /* -D DOWN_CTX      pss-bam.c:17  */
    unsigned int  merged_only;  /* -m MERGED_ONLY   pss-bam.c:18  */
} ora_pss_params;

void ora_pss_default_params(ora_pss_params *p);

/* per-line outcome codes written to `status` (optional, one int8 per line) */
enum {
    ORA_COUNTED     = 0,   /* process_aln returned 0                          */
    ORA_NO_CONTIG   = 1,   /* process_aln returned 1  (pss-bam.c:393-396)      */
    ORA_FILTERED    = -1,  /* process_aln returned -1 / fragkon 2 or -1        */
    ORA_PARSE_FAIL  = -2,  /* line2saml returned 1    (sam-parse.c:88-90)      */
    ORA_UNDEFINED   = -3   /* reference behaviour depends on stale buffers
                              (paired read with strlen(SEQ) < |TLEN|): dropped */
};

typedef struct ora_stats {
    uint64_t lines;
    uint64_t counted;
    uint64_t no_contig;
    uint64_t filtered;
    uint64_t parse_fail;
    uint64_t undefined;
} ora_stats;

/* The fgets/line2saml/process_aln loop, pss-bam.c:764-783.
 * fwd/rev: (region_len+2)*16 counters each, row-major, ZEROED BY CALLER
 * (accumulates, so several SAM blocks can be tallied into one table).
 * status: NULL or array with room for one entry per line.
 * Returns number of lines seen. */
uint64_t ora_pss_tally(const ora_genome *g, const char *sam, size_t sam_len,
                       const ora_pss_params *p,
                       uint64_t *fwd, uint64_t *rev,
                       int8_t *status, size_t status_cap, ora_stats *st);

/* pss-bam.c:504-529 find_sub_rates; rates: region_len*12 doubles, zeroed here */
void ora_pss_rates(const uint64_t *counts, int region_len, double *rates);

/* pss-bam.c:538-586 print_counts / :595-633 print_rates. Return 0 ok. */
int ora_pss_write_counts(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                         const uint64_t *fwd, const uint64_t *rev, int region_len);
int ora_pss_write_rates(const char *fasta_fn, const char *bam_fn, const char *out_prefix,
                        const double *fwd_rates, const double *rev_rates, int region_len);

/* Test hook: glibc's conversions of one NUL-terminated line (sam-parse.c:36-50).
 * 0 ok / 1 line2saml fails / 2 undefined (a token would overflow a field).
 * rname/cigar/seq: caller buffers of 2048 bytes. */
int ora_parse_line(const char *line, unsigned int *flag, unsigned long *pos, unsigned int *mapq,
                   int *isize, char *rname, char *cigar, char *seq);

/* ---- fragkon (fragkon.c) ------------------------------------------------ */
typedef struct ora_fk_params {
    int           klen;         /* -k KLEN         fragkon.c:14 */
    unsigned long min_len;      /* -l              fragkon.c:16 */
    unsigned long max_len;      /* -L              fragkon.c:17 */
    int           min_mq;       /* -q              fragkon.c:15 */
    int           merged_only;  /* -m              fragkon.c:18 */
} ora_fk_params;

void ora_fk_default_params(ora_fk_params *p);

/* fragkon.c:342-363 loop + :122-216 process_aln.  fp/tp: 4^k counters each,
 * unsaturated u64 (the reference's UINT_MAX clamp, kmer.c:102-104, is applied
 * by the writer).  ZEROED BY CALLER. */
uint64_t ora_fragkon_tally(const ora_genome *g, const char *sam, size_t sam_len,
                           const ora_fk_params *p,
                           uint64_t *fp, uint64_t *tp,
                           int8_t *status, size_t status_cap, ora_stats *st);

/* fragkon.c:367-369 + :231-249 output to an open stream. */
int ora_fragkon_write(void *FILE_out, const char *fasta_fn, const char *bam_fn,
                      int klen, const uint64_t *fp, const uint64_t *tp);

/* ---- genome-kmer-count (genome-kmer-count.c, kmer.c) --------------------- */
/* kmer.c:184-214 kmer2inx: returns 1 valid / 0 invalid (inverted convention) */
int  ora_kmer2inx(const char *kmer, size_t klen, uint64_t *inx);
/* genome-kmer-count.c:68-79 count_kmers over every contig; counts: 4^k u64,
 * ZEROED BY CALLER, unsaturated. */
void ora_kmer_spectrum(const ora_genome *g, int k, uint64_t *counts);
/* genome-kmer-count.c:52-53,61-64 */
int  ora_kmer_spectrum_write(void *FILE_out, size_t n_seqs, int k, const uint64_t *counts);

#ifdef __cplusplus
}
#endif
#endif
