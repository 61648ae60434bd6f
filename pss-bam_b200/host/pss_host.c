/* pss_host.c -- see pss_host.h */
#define _GNU_SOURCE
#include "pss_host.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

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

/* FNV-1a, 64 bit */
static uint64_t fnv1a(const char *s)
{
    uint64_t h = 1469598103934665603ull;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 1099511628211ull; }
    return h;
}

/* What a packed-genome cache was built from: size and mtime of the FASTA and a hash of its canonical path.  The tag
 * goes into the cache header (pssgpu_genome_save_tagged) and must match on load; its path hash also keys the file
 * name, so /a/hg19/genome.fa and /b/hg38/genome.fa never share a cache in one cache directory. */
static int fasta_tag(const char *fasta_fn, pssgpu_genome_tag *tag)
{
    struct stat sf;
    char        real[PATH_MAX];
    if (stat(fasta_fn, &sf) != 0) return -1;
    memset(tag, 0, sizeof *tag);
    tag->source_size = (uint64_t)sf.st_size;
    tag->source_mtime_ns = (int64_t)sf.st_mtim.tv_sec * 1000000000ll + (int64_t)sf.st_mtim.tv_nsec;
    tag->source_path_hash = fnv1a(realpath(fasta_fn, real) ? real : fasta_fn);
    return 0;
}

void pss_resident_genome(pssgpu_ctx *ctx, const char *fasta_fn, unsigned long *n_seqs)
{
    const char *env = getenv("PSSGPU_GENOME_CACHE");
    char        cache[2 * MAX_FN_LEN + 64];
    pssgpu_genome_tag tag;
    cache[0] = 0;
    if (env && *env && fasta_tag(fasta_fn, &tag) == 0) {
        if (strcmp(env, "1") == 0) snprintf(cache, sizeof cache, "%s.pssgpu", fasta_fn);
        else {
            const char *base = strrchr(fasta_fn, '/');
            snprintf(cache, sizeof cache, "%s/%s.%016llx.pssgpu", env, base ? base + 1 : fasta_fn,
                     (unsigned long long)tag.source_path_hash);
        }
        if (access(cache, R_OK) == 0) {
            uint64_t nc = 0;
            if (pssgpu_genome_load_tagged(ctx, cache, &tag) == PSSGPU_OK && pssgpu_genome_info(ctx, &nc, NULL, NULL) == PSSGPU_OK) {
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
    if (cache[0] && pssgpu_genome_save_tagged(ctx, cache, &tag) != PSSGPU_OK)
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

/* The pump: a reader thread fills one of two pinned buffers from the pipe while the calling thread hands the other
 * to pssgpu_feed (which returns when its bytes are on the device; the tally of a chunk runs on while the next is
 * copied).  Pipe read, H2D copy and kernel thus overlap instead of alternating. */
typedef struct pump {
    FILE           *in;
    char           *buf[2];
    size_t          got[2];
    int             full[2], eof;
    pthread_mutex_t mu;
    pthread_cond_t  cv;
} pump;

static void *pump_reader(void *arg)
{
    pump *p = (pump *)arg;
    int   k = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->full[k]) pthread_cond_wait(&p->cv, &p->mu);
        pthread_mutex_unlock(&p->mu);
        const size_t got = fread(p->buf[k], 1, PSS_PIPE_CHUNK, p->in);
        pthread_mutex_lock(&p->mu);
        p->got[k] = got;
        p->full[k] = 1;
        if (got == 0) p->eof = 1;
        pthread_cond_broadcast(&p->cv);
        pthread_mutex_unlock(&p->mu);
        if (got == 0) return NULL;
        k ^= 1;
    }
}

static int pump_run(pssgpu_ctx *ctx, FILE *in, int as_bam);

int pss_stream_sam(pssgpu_ctx *ctx, FILE *sam) { return pump_run(ctx, sam, 0); }

int pss_is_bgzf(const char *fn)
{
    unsigned char h[16];
    FILE *f = fopen(fn, "rb");
    size_t n;
    if (!f) return 0;
    n = fread(h, 1, sizeof h, f);
    fclose(f);
    return n == sizeof h && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
}

int pss_stream_input(pssgpu_ctx *ctx, const char *bam_fn, const char *read_group)
{
    int rc;
    if (pss_is_bgzf(bam_fn) && !getenv("PSSGPU_USE_SAMTOOLS")) {
        /* the file's own bytes go to the GPU: BGZF inflate, BAM decoding and the -R filter run there */
        FILE *f = fopen(bam_fn, "rb");
        if (!f) { fprintf(stderr, "Error: Unable to open %s.\n", bam_fn); exit(1); }
        rc = pssgpu_bam_read_group(ctx, read_group);
        if (rc == PSSGPU_OK) rc = pump_run(ctx, f, 1);
        fclose(f);
        return rc;
    }
    {
        FILE *sam = pss_bam_to_sam(bam_fn, read_group);
        rc = pump_run(ctx, sam, 0);
        pclose(sam);
        return rc;
    }
}

static int pump_run(pssgpu_ctx *ctx, FILE *sam, int as_bam)
{
    pump      p;
    pthread_t th;
    int       rc = PSSGPU_OK, k = 0;
    memset(&p, 0, sizeof p);
    p.in = sam;
    p.buf[0] = (char *)pssgpu_host_alloc(PSS_PIPE_CHUNK);
    p.buf[1] = (char *)pssgpu_host_alloc(PSS_PIPE_CHUNK);
    if (!p.buf[0] || !p.buf[1]) { pssgpu_host_free(p.buf[0]); pssgpu_host_free(p.buf[1]); return PSSGPU_ENOMEM; }
    pthread_mutex_init(&p.mu, NULL);
    pthread_cond_init(&p.cv, NULL);
    if (pthread_create(&th, NULL, pump_reader, &p) != 0) {
        pssgpu_host_free(p.buf[0]); pssgpu_host_free(p.buf[1]);
        return PSSGPU_ENOMEM;
    }
    for (;;) {
        pthread_mutex_lock(&p.mu);
        while (!p.full[k]) pthread_cond_wait(&p.cv, &p.mu);
        const size_t got = p.got[k];
        pthread_mutex_unlock(&p.mu);
        if (got == 0) break;
        if (rc == PSSGPU_OK)                                               /* after an error the pipe is still drained */
            rc = as_bam ? pssgpu_feed_bam(ctx, p.buf[k], got, 0) : pssgpu_feed(ctx, p.buf[k], got, 0);
        pthread_mutex_lock(&p.mu);
        p.full[k] = 0;
        pthread_cond_broadcast(&p.cv);
        pthread_mutex_unlock(&p.mu);
        k ^= 1;
    }
    pthread_join(th, NULL);
    if (rc == PSSGPU_OK)                                                  /* a last line without '\n' / the end of the BAM */
        rc = as_bam ? pssgpu_feed_bam(ctx, p.buf[0], 0, 1) : pssgpu_feed(ctx, p.buf[0], 0, 1);
    if (rc == PSSGPU_OK) rc = pssgpu_sync(ctx);
    pthread_mutex_destroy(&p.mu);
    pthread_cond_destroy(&p.cv);
    pssgpu_host_free(p.buf[0]);
    pssgpu_host_free(p.buf[1]);
    return rc;
}
