/* pss-bam on a B200: same command line and the same two output tables as the
 * reference program (pss-bam.c:650-805), with its per-read loop
 * (pss-bam.c:764-783) replaced by the GPU tally behind include/pssgpu.h.
 *
 *   pss-bam -F genome.fa -B reads.bam -o prefix [-r 15] [-l 0] [-L 250000000]
 *           [-q 0] [-R read_group] [-U ACGT] [-D ACGT] [-m]
 *
 * The host keeps what the reference's main keeps: options, FASTA loading,
 * the `samtools view` pipe, rate arithmetic and the table writers.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "pss_host.h"
#include "pss_tables.h"

static void usage(void)
{
    fputs("pss-bam v" PSS_VERSION ": Program for describing base context and counting\n"
          "the number of matches/mismatches in aligned reads to a genome.\n"
          "-F <reference FASTA (required)>\n"
          "-B <input BAM (required)>\n"
          "-o <output filename prefix (required)>\n"
          "-r <length in basepairs into the interior of alignments to report on (default: 15)>\n"
          "-l <minimum length of read to report (default: 0)>\n"
          "-L <maximum length of read to report (default: 250000000)>\n"
          "-q <map quality filter of read to report (default: 0)>\n"
          "-R <read group name to restrict analysis to (default: all reads)>\n"
          "-U <upstream context base filter; first base before alignment must be one of these (default: ACGT)>\n"
          "-D <downstream context base filter; first base before alignment must be one of these (default: ACGT)>\n"
          "-m <only consider merged reads>\n", stderr);
    exit(1);
}

int main(int argc, char *argv[])
{
    const char *fasta_fn = NULL, *bam_fn = NULL, *out_prefix = NULL, *read_group = NULL;
    pssgpu_pss_params par;
    int opt;

    pssgpu_pss_default_params(&par);
    while ((opt = getopt(argc, argv, ":F:B:o:R:r:l:L:q:U:D:m")) != -1) {
        switch (opt) {
        case 'F': fasta_fn = optarg; break;
        case 'B': bam_fn = optarg; break;
        case 'o': out_prefix = optarg; break;
        case 'R': read_group = optarg; break;
        case 'r': par.region_len = atoi(optarg); break;
        case 'l': par.min_len = strtoul(optarg, NULL, 10); break;
        case 'L': par.max_len = strtoul(optarg, NULL, 10); break;
        case 'q': par.min_mq = atoi(optarg); break;
        case 'U': par.up_ctx = optarg; break;
        case 'D': par.down_ctx = optarg; break;
        case 'm': par.merged_only = 1; break;
        case ':':
            fprintf(stderr, "Please enter required argument for option -%c.\n", optopt);
            exit(0);
        case '?':
            if (isprint(optopt)) fprintf(stderr, "Unknown option -%c.\n", optopt);
            else fprintf(stderr, "Unknown option character \\x%x.\n", optopt);
            break;
        default:
            fprintf(stderr, "Error parsing command-line options.\n");
            exit(0);
        }
    }
    for (int i = optind; i < argc; i++) fprintf(stderr, "Non-option argument %s\n", argv[i]);
    if (!fasta_fn || !bam_fn || !out_prefix) usage();

    fprintf(stderr, "Full command: %s -F %s -B %s -o %s -r %d -l %lu -L %lu -q %d", argv[0], fasta_fn, bam_fn, out_prefix,
            par.region_len, par.min_len, par.max_len, par.min_mq);
    if (read_group) fprintf(stderr, " -R %s", read_group);
    fprintf(stderr, " -U %s -D %s%s\n", par.up_ctx, par.down_ctx, par.merged_only ? " -m" : "");

    pssgpu_group *gpu = pss_open_devices();

    fprintf(stderr, "Reading genome sequence from:\n%s\n", fasta_fn);
    pss_resident_genome_group(gpu, fasta_fn, NULL);
    fprintf(stderr, "Finished loading genome.\nCounting matches/mismatches from:\n%s\n", bam_fn);

    if (pssgpu_group_pss_begin(gpu, &par) != PSSGPU_OK) pss_die_group(gpu, "pss_begin");
    if (pss_stream_input_group(gpu, bam_fn, read_group) != PSSGPU_OK) pss_die_group(gpu, "tally");

    const int    R = par.region_len;
    const size_t cells = (size_t)(R + 2) * 16;
    uint64_t *fwd = (uint64_t *)calloc(cells, sizeof *fwd), *rev = (uint64_t *)calloc(cells, sizeof *rev);
    double   *fwd_rates = (double *)calloc((size_t)(R > 0 ? R : 1) * 12, sizeof(double));
    double   *rev_rates = (double *)calloc((size_t)(R > 0 ? R : 1) * 12, sizeof(double));
    if (pssgpu_group_pss_finish(gpu, fwd, rev) != PSSGPU_OK) pss_die_group(gpu, "pss_finish");
    pss_sub_rates(fwd, R, fwd_rates);
    pss_sub_rates(rev, R, rev_rates);
    int rc = pss_write_counts(fasta_fn, bam_fn, out_prefix, fwd, rev, R);
    rc |= pss_write_rates(fasta_fn, bam_fn, out_prefix, fwd_rates, rev_rates, R);

    if (getenv("PSSGPU_VERBOSE")) {
        pssgpu_stats st;
        if (pssgpu_group_get_stats(gpu, &st, 0) == PSSGPU_OK)
            fprintf(stderr, "lines %llu counted %llu no_contig %llu filtered %llu unparsable %llu undefined %llu\n",
                    (unsigned long long)st.lines, (unsigned long long)st.counted, (unsigned long long)st.no_contig,
                    (unsigned long long)st.filtered, (unsigned long long)st.parse_fail, (unsigned long long)st.undefined);
    }
    free(fwd); free(rev); free(fwd_rates); free(rev_rates);
    pssgpu_group_destroy(gpu);
    fprintf(stderr, "Done.\n");
    return rc;
}
