/* genome-kmer-count on a B200: command line and stdout of the reference
 * program (genome-kmer-count.c:23-66); the per-base insert loop
 * (genome-kmer-count.c:56-58, :68-79) is the GPU spectrum kernel behind
 * include/pssgpu.h.
 *
 *   genome-kmer-count -f genome.fa [-k 4]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "pss_host.h"
#include "pss_tables.h"

static void help(void)
{
    fputs("genome-kmer-count -f <fasta genome file>\n"
          "                  -k <kmer size; default = 4>\n"
          "This program reports the number of observed number\n"
          "of all possible kmers of the given length in the\n"
          "input genome.\n", stdout);
    exit(0);
}

int main(int argc, char *argv[])
{
    const char *fa_in = NULL;
    int k = 4, ich;
    while ((ich = getopt(argc, argv, "f:k:")) != -1) {
        switch (ich) {
        case 'f': fa_in = optarg; break;
        case 'k': k = atoi(optarg); break;
        default:  help();
        }
    }
    if (!fa_in || !*fa_in) help();

    pssgpu_group *gpu = pss_open_devices();
    unsigned long n_seqs = 0;
    pss_resident_genome_group(gpu, fa_in, &n_seqs);
    if (k < 1 || k > 14) { fprintf(stderr, "ERROR: k must be in [1,14] on this build\n"); return 1; }

    const size_t bins = (size_t)1 << (2 * k);
    uint64_t *counts = (uint64_t *)calloc(bins, sizeof *counts);
    /* one GPU: the whole genome; several: GPU i counts slice i, the 4^k counters are summed over NVLink */
    if (pssgpu_group_kmer_spectrum(gpu, k, counts) != PSSGPU_OK) pss_die_group(gpu, "kmer_spectrum");
    if (getenv("PSSGPU_VERBOSE") && pssgpu_group_size(gpu) > 1)
        fprintf(stderr, "spectrum summed over %d GPUs (%s) in %.3f ms\n", pssgpu_group_size(gpu), pssgpu_group_reduce_backend(gpu),
                pssgpu_group_last_reduce_ms(gpu));
    pss_write_spectrum(stdout, n_seqs, k, counts);
    free(counts);
    pssgpu_group_destroy(gpu);
    exit(0);
}
