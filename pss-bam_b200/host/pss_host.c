/* pss_host.c -- see pss_host.h */
#define _GNU_SOURCE
#include "pss_host.h"
#include "pss_io.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#define PSS_PIPE_CHUNK (32u << 20)
/* A BAM file is read in pieces of the device batch size: a feed call forms its batches from its own bytes, and a batch
 * must hold several thousand BGZF blocks to fill the inflate kernel (one warp per block, 4 440 warps). */
#define PSS_BAM_CHUNK (256u << 20)

void pss_die(pssgpu_ctx *ctx, const char *what)
{
    fprintf(stderr, "ERROR: %s: %s\n", what, pssgpu_last_error(ctx));
    exit(1);
}

void pss_die_group(pssgpu_group *g, const char *what)
{
    fprintf(stderr, "ERROR: %s: %s\n", what, pssgpu_group_last_error(g));
    exit(1);
}

pssgpu_group *pss_open_devices(void)
{
    /* $PSSGPU_DEVICES: "all", or a comma separated list of CUDA device numbers; unset: $PSSGPU_DEVICE or device 0 */
    const char *list = getenv("PSSGPU_DEVICES"), *one = getenv("PSSGPU_DEVICE");
    int dev[64], n = 0;
    pssgpu_group *g = NULL;
    if (list && *list && strcmp(list, "all") != 0) {
        const char *p = list;
        while (*p && n < 64) {
            char *end;
            long v = strtol(p, &end, 10);
            if (end == p) { fprintf(stderr, "ERROR: cannot read PSSGPU_DEVICES=%s\n", list); exit(1); }
            dev[n++] = (int)v;
            p = *end == ',' ? end + 1 : end;
            if (*end && *end != ',') { fprintf(stderr, "ERROR: cannot read PSSGPU_DEVICES=%s\n", list); exit(1); }
        }
    } else if (!(list && *list)) {
        dev[n++] = one ? atoi(one) : 0;
    }
    if (pssgpu_group_init(n ? dev : NULL, n, &g) != PSSGPU_OK) {
        fprintf(stderr, "ERROR: no usable B200: %s\n(this build has no CPU path)\n", pssgpu_last_error(NULL));
        exit(1);
    }
    if (pssgpu_group_size(g) > 1)
        fprintf(stderr, "Using %d GPUs; tables are summed with: %s\n", pssgpu_group_size(g), pssgpu_group_reduce_backend(g));
    return g;
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

void pss_resident_genome_group(pssgpu_group *g, const char *fasta_fn, unsigned long *n_seqs)
{
    const int n = pssgpu_group_size(g);
    const char *env = getenv("PSSGPU_GENOME_CACHE");
    char        cache[2 * MAX_FN_LEN + 64];
    pssgpu_genome_tag tag;
    if (n == 1) { pss_resident_genome(pssgpu_group_ctx(g, 0), fasta_fn, n_seqs); return; }
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
            if (pssgpu_group_genome_load_tagged(g, cache, &tag) == PSSGPU_OK &&
                pssgpu_genome_info(pssgpu_group_ctx(g, 0), &nc, NULL, NULL) == PSSGPU_OK) {
                if (n_seqs) *n_seqs = (unsigned long)nc;
                return;
            }
            fprintf(stderr, "WARNING: ignoring genome cache %s: %s\n", cache, pssgpu_group_last_error(g));
        }
    }
    Genome *genome = init_genome(fasta_fn);
    if (!genome) { fprintf(stderr, "ERROR: cannot read %s\n", fasta_fn); exit(1); }
    {
        pssgpu_contig *c = (pssgpu_contig *)malloc((genome->n_seqs ? genome->n_seqs : 1) * sizeof *c);
        for (size_t i = 0; i < genome->n_seqs; i++) {
            c[i].id = genome->seqs[i]->id;
            c[i].seq = genome->seqs[i]->seq;
            c[i].len = genome->seqs[i]->len;
        }
        if (pssgpu_group_genome_upload(g, c, genome->n_seqs) != PSSGPU_OK) pss_die_group(g, "genome upload");
        free(c);
    }
    if (n_seqs) *n_seqs = (unsigned long)genome->n_seqs;
    destroy_genome(genome);
    if (cache[0] && pssgpu_genome_save_tagged(pssgpu_group_ctx(g, 0), cache, &tag) != PSSGPU_OK)
        fprintf(stderr, "WARNING: could not write genome cache %s: %s\n", cache, pssgpu_last_error(pssgpu_group_ctx(g, 0)));
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
    size_t          got[2], chunk;
    int             full[2], eof;
    int             fd, read_threads;      /* fd >= 0: a regular file, read with read_threads concurrent pread()s */
    off_t           offset;
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
        size_t got;
        if (p->fd >= 0) {
            got = pss_pread_parallel(p->fd, p->buf[k], p->chunk, p->offset, p->read_threads);
            p->offset += (off_t)got;
        } else {
            got = fread(p->buf[k], 1, p->chunk, p->in);
        }
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

static int pump_run(pssgpu_ctx *ctx, pssgpu_group *g, FILE *in, int as_bam);

int pss_stream_sam(pssgpu_ctx *ctx, FILE *sam) { return pump_run(ctx, NULL, sam, 0); }

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
        if (rc == PSSGPU_OK) rc = pump_run(ctx, NULL, f, 1);
        fclose(f);
        return rc;
    }
    {
        FILE *sam = pss_bam_to_sam(bam_fn, read_group);
        rc = pump_run(ctx, NULL, sam, 0);
        pclose(sam);
        return rc;
    }
}

/* g != NULL: a BAM file over the GPUs of a group (pssgpu_group_feed_bam); else the one context */
static int pump_feed(pssgpu_ctx *ctx, pssgpu_group *g, const char *buf, size_t n, int as_bam, int last)
{
    if (g) return pssgpu_group_feed_bam(g, buf, n, last);
    return as_bam ? pssgpu_feed_bam(ctx, buf, n, last) : pssgpu_feed(ctx, buf, n, last);
}

static int pump_run(pssgpu_ctx *ctx, pssgpu_group *g, FILE *sam, int as_bam)
{
    pump      p;
    pthread_t th;
    int       rc = PSSGPU_OK, k = 0;
    memset(&p, 0, sizeof p);
    p.in = sam;
    p.chunk = PSS_PIPE_CHUNK;
    p.fd = -1;
    if (as_bam) {                                                           /* no larger than the file */
        struct stat sb;
        p.chunk = PSS_BAM_CHUNK;
        if (getenv("PSS_BAM_CHUNK_MB") && atoi(getenv("PSS_BAM_CHUNK_MB")) > 0)      /* test switch: many small pieces */
            p.chunk = (size_t)atoi(getenv("PSS_BAM_CHUNK_MB")) << 20;
        if (fstat(fileno(sam), &sb) == 0 && S_ISREG(sb.st_mode)) {
            const char *e = getenv("PSS_READ_THREADS");
            if ((size_t)sb.st_size + 1 < p.chunk) p.chunk = (size_t)sb.st_size + 1 > (1u << 20) ? (size_t)sb.st_size + 1 : (1u << 20);
            /* the file itself, by several pread()s at once: one thread copies out of the page cache at a third of
             * the rate the GPU inflates at */
            p.read_threads = e ? atoi(e) : 4;
            if (p.read_threads > 0) { p.fd = fileno(sam); p.offset = ftello(sam); if (p.offset < 0) p.fd = -1; }
        }
    }
    p.buf[0] = (char *)pssgpu_host_alloc(p.chunk);
    p.buf[1] = (char *)pssgpu_host_alloc(p.chunk);
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
            rc = pump_feed(ctx, g, p.buf[k], got, as_bam, 0);
        pthread_mutex_lock(&p.mu);
        p.full[k] = 0;
        pthread_cond_broadcast(&p.cv);
        pthread_mutex_unlock(&p.mu);
        k ^= 1;
    }
    pthread_join(th, NULL);
    if (rc == PSSGPU_OK)                                                  /* a last line without '\n' / the end of the BAM */
        rc = pump_feed(ctx, g, p.buf[0], 0, as_bam, 1);
    if (rc == PSSGPU_OK) rc = g ? pssgpu_group_sync(g) : pssgpu_sync(ctx);
    pthread_mutex_destroy(&p.mu);
    pthread_cond_destroy(&p.cv);
    pssgpu_host_free(p.buf[0]);
    pssgpu_host_free(p.buf[1]);
    return rc;
}

/* SAM text to the members of a group: whole lines, chunk by chunk in turn.  Every GPU has two pinned buffers; a chunk
 * is read straight into one of them (behind the unfinished line left over from the previous chunk), cut after its
 * last newline and handed to pssgpu_feed_async, which does not wait for the copy -- so the PCIe links of all GPUs work
 * at once while this thread reads on.  A buffer is reused two rounds later, after pssgpu_feed_wait on its GPU. */
static int deal_sam(pssgpu_group *g, FILE *in)
{
    const int n = pssgpu_group_size(g);
    char    **buf = (char **)calloc((size_t)2 * n, sizeof *buf);
    char     *carry = (char *)malloc(PSS_PIPE_CHUNK);
    size_t    carry_len = 0, k = 0;
    int       rc = PSSGPU_OK, single = -1;          /* single >= 0: a line longer than a chunk -- the rest goes to that GPU */
    for (int i = 0; i < 2 * n; i++) {
        buf[i] = (char *)pssgpu_host_alloc(PSS_PIPE_CHUNK);
        if (!buf[i]) rc = PSSGPU_ENOMEM;
    }
    while (rc == PSSGPU_OK) {
        const int   gpu = single >= 0 ? single : (int)(k % (size_t)n), slot = (int)((k / (size_t)n) & 1u);
        pssgpu_ctx *c = pssgpu_group_ctx(g, gpu);
        char       *b = buf[2 * gpu + slot];
        if (k >= (size_t)2 * n || single >= 0) rc = pssgpu_feed_wait(c);
        if (rc != PSSGPU_OK) break;
        memcpy(b, carry, carry_len);
        const size_t got = fread(b + carry_len, 1, PSS_PIPE_CHUNK - carry_len, in);
        const size_t total = carry_len + got;
        if (got == 0) {
            if (total) rc = pssgpu_feed_async(c, b, total, 1);          /* a last line without a newline */
            break;
        }
        if (single >= 0) {                          /* the member's own carry logic takes partial lines */
            rc = pssgpu_feed_async(c, b, total, 0);
            carry_len = 0;
            k++;
            continue;
        }
        char *nl = (char *)memrchr(b, '\n', total);
        if (!nl) {
            if (total < PSS_PIPE_CHUNK) { memcpy(carry, b, total); carry_len = total; continue; }   /* short read: keep reading */
            single = gpu;
            rc = pssgpu_feed_async(c, b, total, 0);
            carry_len = 0;
            k++;
            continue;
        }
        const size_t cut = (size_t)(nl - b) + 1;
        carry_len = total - cut;
        memcpy(carry, b + cut, carry_len);
        rc = pssgpu_feed_async(c, b, cut, 0);
        k++;
    }
    for (int i = 0; i < n; i++) {
        pssgpu_ctx *c = pssgpu_group_ctx(g, i);
        int r2 = pssgpu_feed_wait(c);
        if (r2 == PSSGPU_OK && single == i) r2 = pssgpu_feed(c, buf[0], 0, 1);       /* flush that member's carried line */
        if (r2 == PSSGPU_OK) r2 = pssgpu_sync(c);
        if (rc == PSSGPU_OK && r2 != PSSGPU_OK) { rc = r2; fprintf(stderr, "ERROR: GPU member %d: %s\n", i, pssgpu_last_error(c)); }
    }
    for (int i = 0; i < 2 * n; i++) pssgpu_host_free(buf[i]);
    free(buf);
    free(carry);
    return rc;
}

int pss_stream_input_group(pssgpu_group *g, const char *bam_fn, const char *read_group)
{
    const int n = pssgpu_group_size(g);
    int rc;
    if (n == 1) return pss_stream_input(pssgpu_group_ctx(g, 0), bam_fn, read_group);
    if (pss_is_bgzf(bam_fn) && !getenv("PSSGPU_USE_SAMTOOLS")) {
        /* BAM records run across BGZF blocks and carry no sync marks: the byte stream of one file cannot be dealt like
         * text.  The inflate can: batches of BGZF blocks go to the members in turn, the first member fetches the inflated
         * bytes and frames, renders and tallies them (pssgpu_group_feed_bam); the others add zeros to the tables. */
        FILE *f = fopen(bam_fn, "rb");
        if (!f) { fprintf(stderr, "Error: Unable to open %s.\n", bam_fn); exit(1); }
        rc = pssgpu_group_bam_read_group(g, read_group);
        if (rc == PSSGPU_OK) rc = pump_run(pssgpu_group_ctx(g, 0), g, f, 1);
        fclose(f);
        return rc;
    }
    {
        FILE *sam = pss_bam_to_sam(bam_fn, read_group);
        rc = deal_sam(g, sam);
        pclose(sam);
    }
    return rc;
}
