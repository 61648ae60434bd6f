/* pss_host.c -- see pss_host.h */
#include "pss_host.h"

#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#define PSS_PIPE_CHUNK (32u << 20)

void pss_die(pssgpu_ctx *ctx, const char *what)
{
    fprintf(stderr, "ERROR: %s: %s\n", what, pssgpu_last_error(ctx));
    exit(1);
}

pssgpu_ctx *pss_open_device(void)
{
    const char *env = getenv("PSSGPU_DEVICE");
    pssgpu_ctx *ctx = NULL;
    if (pssgpu_init(env ? atoi(env) : 0, &ctx) != PSSGPU_OK) {
        fprintf(stderr, "ERROR: no usable B200: %s\n(this build has no CPU path)\n", pssgpu_last_error(NULL));
        exit(1);
    }
    return ctx;
}

int pss_upload_genome(pssgpu_ctx *ctx, const Genome *genome)
{
    pssgpu_contig *c = (pssgpu_contig *)malloc((genome->n_seqs ? genome->n_seqs : 1) * sizeof *c);
    int rc;
    for (size_t i = 0; i < genome->n_seqs; i++) {
        c[i].id = genome->seqs[i]->id;
        c[i].seq = genome->seqs[i]->seq;
        c[i].len = genome->seqs[i]->len;
    }
    rc = pssgpu_genome_upload(ctx, c, genome->n_seqs);
    free(c);
    return rc;
}

void pss_resident_genome(pssgpu_ctx *ctx, const char *fasta_fn, unsigned long *n_seqs)
{
    const char *env = getenv("PSSGPU_GENOME_CACHE");
    char        cache[2 * MAX_FN_LEN + 16];
    cache[0] = 0;
    if (env && *env) {
        if (strcmp(env, "1") == 0) snprintf(cache, sizeof cache, "%s.pssgpu", fasta_fn);
        else {
            const char *base = strrchr(fasta_fn, '/');
            snprintf(cache, sizeof cache, "%s/%s.pssgpu", env, base ? base + 1 : fasta_fn);
        }
        struct stat sf, sc;
        if (stat(fasta_fn, &sf) == 0 && stat(cache, &sc) == 0 && sc.st_mtime >= sf.st_mtime) {
            uint64_t nc = 0;
            if (pssgpu_genome_load(ctx, cache) == PSSGPU_OK && pssgpu_genome_info(ctx, &nc, NULL, NULL) == PSSGPU_OK) {
                if (n_seqs) *n_seqs = (unsigned long)nc;
                return;
            }
            fprintf(stderr, "WARNING: ignoring genome cache %s: %s\n", cache, pssgpu_last_error(ctx));
        }
    }
    Genome *genome = init_genome(fasta_fn);
    if (!genome) { fprintf(stderr, "ERROR: cannot read %s\n", fasta_fn); exit(1); }
    if (pss_upload_genome(ctx, genome) != PSSGPU_OK) pss_die(ctx, "genome upload");
    if (n_seqs) *n_seqs = (unsigned long)genome->n_seqs;
    destroy_genome(genome);
    if (cache[0] && pssgpu_genome_save(ctx, cache) != PSSGPU_OK)
        fprintf(stderr, "WARNING: could not write genome cache %s: %s\n", cache, pssgpu_last_error(ctx));
}

FILE *pss_bam_to_sam(const char *bam_fn, const char *read_group)
{
    char  cmd[2 * MAX_FN_LEN + 64];
    FILE *p;
    if (read_group) snprintf(cmd, sizeof cmd, "samtools view -r %s %s", read_group, bam_fn);
    else snprintf(cmd, sizeof cmd, "samtools view %s", bam_fn);
    p = popen(cmd, "r");
    if (!p) {
        fprintf(stderr, "Error: Unable to open %s with samtools view.\n", bam_fn);
        exit(1);
    }
    return p;
}

int pss_stream_sam(pssgpu_ctx *ctx, FILE *sam)
{
    char  *buf = (char *)pssgpu_host_alloc(PSS_PIPE_CHUNK);
    size_t got;
    int    rc = PSSGPU_OK;
    if (!buf) return PSSGPU_ENOMEM;
    while (rc == PSSGPU_OK && (got = fread(buf, 1, PSS_PIPE_CHUNK, sam)) > 0) rc = pssgpu_feed(ctx, buf, got, 0);
    if (rc == PSSGPU_OK) rc = pssgpu_feed(ctx, buf, 0, 1);      /* a last line without '\n' */
    pssgpu_host_free(buf);
    return rc;
}
