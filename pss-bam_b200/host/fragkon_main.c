/* fragkon on a B200: command line and stdout table of the reference program
 * (fragkon.c:253-386); its per-read loop (fragkon.c:342-363) is the GPU
 * end-context histogram behind include/pssgpu.h.
 *
 *   fragkon -F genome.fa -B reads.bam [-k 8] [-l 0] [-L 250000000] [-q 0] [-m]
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
    fputs("fragkon: Program for describing kmer-based genomic sequence\n"
          "contexts around the fragmentation points of aligned reads.\n"
          "-F <reference FASTA (required)>\n"
          "-B <input BAM (required)>\n"
          "-k <kmer length (default: 8)>\n"
          "-l <minimum length of read to report (default: 0)>\n"
          "-L <maximum length of read to report (default: 250000000)>\n"
          "-q <map quality filter of read to report (default: 0)>\n"
          "-m <only consider merged reads>\n", stderr);
    exit(1);
}

int main(int argc, char *argv[])
{
    const char *fasta_fn = NULL, *bam_fn = NULL;
    pssgpu_fragkon_params par;
    int opt;

    pssgpu_fragkon_default_params(&par);
    while ((opt = getopt(argc, argv, ":F:B:k:l:L:q:m")) != -1) {
        switch (opt) {
        case 'F': fasta_fn = optarg; break;
        case 'B': bam_fn = optarg; break;
        case 'k': par.klen = atoi(optarg); break;
        case 'l': par.min_len = strtoul(optarg, NULL, 10); break;
        case 'L': par.max_len = strtoul(optarg, NULL, 10); break;
        case 'q': par.min_mq = atoi(optarg); break;
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
    if (!fasta_fn || !bam_fn) usage();

    fputs("# Entered command:", stderr);
    for (int i = 0; i < argc; i++) fprintf(stderr, " %s", argv[i]);
    fprintf(stderr, " \nInput kmer length = %d.\n", par.klen);
    if (par.klen & 1)
        fprintf(stderr, "    *** k is odd - counting %d bases outside %d bases inside of alignment.\n", par.klen / 2, par.klen / 2 + 1);

    pssgpu_group *gpu = pss_open_devices();
    fprintf(stderr, "Reading genome sequence from: %s\n", fasta_fn);
    pss_resident_genome_group(gpu, fasta_fn, NULL);
    fprintf(stderr, "Finished loading genome.\nCounting kmer contexts for: %s\n", bam_fn);

    if (pssgpu_group_fragkon_begin(gpu, &par) != PSSGPU_OK) pss_die_group(gpu, "fragkon_begin");
    if (pss_stream_input_group(gpu, bam_fn, NULL) != PSSGPU_OK) pss_die_group(gpu, "tally");

    const size_t bins = (size_t)1 << (2 * par.klen);
    uint64_t *fp = (uint64_t *)calloc(bins, sizeof *fp), *tp = (uint64_t *)calloc(bins, sizeof *tp);
    if (pssgpu_group_fragkon_finish(gpu, fp, tp) != PSSGPU_OK) pss_die_group(gpu, "fragkon_finish");
    pss_write_fragkon(stdout, fasta_fn, bam_fn, par.klen, fp, tp);

    free(fp); free(tp);
    pssgpu_group_destroy(gpu);
    fprintf(stderr, "Done.\n");
    return 0;
}
