/* pssgpu.h -- C ABI of the B200 (sm_100a) hot path of pss-bam / fragkon /
 * genome-kmer-count.
 *
 * The reference has no plugin or FFI layer: its three programs call static
 * functions in a per-read / per-base loop.  This header is the boundary a
 * maintainer binds instead of those loops (see INTEGRATION.md for the patch
 * to the reference's own mains):
 *
 *   reference loop replaced                       entry points here
 *   -------------------------------------------   ---------------------------
 *   init_genome() result made device resident     pssgpu_genome_upload
 *     (fasta-genome-io.c:221-238, Seq/Genome
 *      fasta-genome-io.h:13-24)
 *   pss-bam.c:764-783  fgets + line2saml +        pssgpu_pss_begin / pssgpu_feed /
 *     process_aln (sam-parse.c:10-91,               pssgpu_pss_finish
 *     pss-bam.c:390-496, :169-326)
 *   fragkon.c:342-363  fgets + line2saml +        pssgpu_fragkon_begin / pssgpu_feed /
 *     process_aln (fragkon.c:122-216,               pssgpu_fragkon_finish
 *     kmer.c:43-110)
 *   genome-kmer-count.c:56-58 count_kmers         pssgpu_kmer_spectrum
 *     (:68-79, kmer.c:43-110)
 *
 * Conventions: plain C, plain pointers and sizes, no CUDA or torch types.
 * Every call returns 0 on success or a negative PSSGPU_E* code;
 * pssgpu_last_error() gives the text.  A context is bound to one CUDA device
 * and must be used from one host thread at a time (the reference is
 * single-threaded and keeps its configuration in file-scope statics; here the
 * configuration lives in the context).  There is NO CPU fallback: if the
 * device or the kernels are unavailable the call fails.
 */
#ifndef PSSGPU_H
#define PSSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSSGPU_ABI_VERSION 1

enum {
    PSSGPU_OK        = 0,
    PSSGPU_EINVAL    = -1,   /* bad argument / call out of order            */
    PSSGPU_ECUDA     = -2,   /* CUDA runtime error (text in last_error)     */
    PSSGPU_ENOMEM    = -3,   /* host or device allocation failed            */
    PSSGPU_ENOGENOME = -4,   /* tally requested before pssgpu_genome_upload */
    PSSGPU_EUNSUPP   = -5    /* parameter outside the supported range       */
};

typedef struct pssgpu_ctx pssgpu_ctx;

/* ---- lifetime ------------------------------------------------------------ */
int         pssgpu_abi_version(void);
/* Number of visible CUDA devices (<0 on error). */
int         pssgpu_device_count(void);
/* Create a context on CUDA device `device` (one process per GPU: pass
 * LOCAL_RANK).  Fails unless the device is sm_100 class. */
int         pssgpu_init(int device, pssgpu_ctx **out);
void        pssgpu_destroy(pssgpu_ctx *ctx);
const char *pssgpu_last_error(const pssgpu_ctx *ctx);   /* ctx may be NULL: last init error */
/* The cudaStream_t (as void*) every kernel and copy of this context is issued
 * on, so a caller can bracket calls with its own CUDA events. */
void       *pssgpu_cuda_stream(pssgpu_ctx *ctx);

/* Pinned host memory for SAM staging (pssgpu_feed copies straight from it). */
void       *pssgpu_host_alloc(size_t bytes);
void        pssgpu_host_free(void *p);

/* ---- genome residency ------------------------------------------------------
 * One entry per FASTA record, exactly the three fields of the reference's
 * `Seq` (fasta-genome-io.h:13-17): id (NUL terminated), seq (upper-cased
 * ASCII as produced by read_fasta, fasta-genome-io.c:120-131; need not be NUL
 * terminated) and len.  Pointers are borrowed for the duration of the call.
 * The contigs may be given in any order; lookup by name reproduces
 * find_seq()/chr_cmp() (fasta-genome-io.c:202-219): exact byte equality.
 * The genome is packed on the device to 4 bits per base (2-bit base code +
 * 2-bit class) and stays resident until the next upload or
 * pssgpu_destroy. */
typedef struct pssgpu_contig {
    const char *id;
    const char *seq;
    uint64_t    len;
} pssgpu_contig;

int pssgpu_genome_upload(pssgpu_ctx *ctx, const pssgpu_contig *contigs, uint64_t n_contigs);
/* Same, but each seq pointer is DEVICE memory holding the ASCII contig. */
int pssgpu_genome_upload_device(pssgpu_ctx *ctx, const pssgpu_contig *contigs, uint64_t n_contigs);
/* Packed-genome cache (takes the FASTA parse -- read_fasta's
 * fgetc loop, fasta-genome-io.c:105-148, ~150 s for 3.1 Gb -- and the pack off
 * the critical path of repeated runs).  save writes the resident genome (packed
 * groups, contig table, names, exception list) to `path`; load makes the
 * genome of such a file resident instead of pssgpu_genome_upload.  The file is
 * tied to this library's packed layout (magic + layout version, sizes checked
 * against the file length); a mismatch is PSSGPU_EINVAL and leaves no genome
 * resident. */
int pssgpu_genome_save(pssgpu_ctx *ctx, const char *path);
int pssgpu_genome_load(pssgpu_ctx *ctx, const char *path);
/* The same with a fingerprint of what the genome was built from, so that a
 * cache of another FASTA (or of an older version of this one) is never taken
 * for the right one: save stores the tag in the file, load compares all four
 * fields with `expect` and fails with PSSGPU_EINVAL when any differs.  The
 * host programs fill it from stat() and realpath() of the FASTA
 * (host/pss_host.c).  On load, every contig / hash / exception-list entry is
 * range checked before the genome becomes resident. */
typedef struct pssgpu_genome_tag {
    uint64_t source_size;        /* st_size of the FASTA                       */
    int64_t  source_mtime_ns;    /* st_mtim of the FASTA, nanoseconds          */
    uint64_t source_path_hash;   /* FNV-1a of realpath(FASTA)                  */
    uint64_t user;               /* caller defined (0 in the host programs)    */
} pssgpu_genome_tag;
int pssgpu_genome_save_tagged(pssgpu_ctx *ctx, const char *path, const pssgpu_genome_tag *tag);
int pssgpu_genome_load_tagged(pssgpu_ctx *ctx, const char *path, const pssgpu_genome_tag *expect);
/* Total bases resident / bytes of HBM used by the packed genome. */
int pssgpu_genome_info(const pssgpu_ctx *ctx, uint64_t *n_contigs, uint64_t *n_bases, uint64_t *hbm_bytes);

/* ---- pss-bam tally -----------------------------------------------------------
 * Mirrors the file-scope options of pss-bam.c:12-18 (same defaults via
 * pssgpu_pss_default_params). */
typedef struct pssgpu_pss_params {
    int           region_len;   /* -r  REGION_LEN   (default 15); 0..2045 */
    unsigned long min_len;      /* -l  MIN_READ_LEN (default 0)           */
    unsigned long max_len;      /* -L  MAX_READ_LEN (default 250000000)   */
    int           min_mq;       /* -q  MIN_MQ       (default 0)           */
    const char   *up_ctx;       /* -U  UP_CTX       (default "ACGT")      */
    const char   *down_ctx;     /* -D  DOWN_CTX     (default "ACGT")      */
    unsigned int  merged_only;  /* -m  MERGED_ONLY  (default 0)           */
} pssgpu_pss_params;

void pssgpu_pss_default_params(pssgpu_pss_params *p);

/* Per-record outcome counters, in the vocabulary of process_aln's return
 * codes (pss-bam.c:387-389, fragkon.c:119-121). */
typedef struct pssgpu_stats {
    uint64_t lines;        /* fgets() lines seen                                   */
    uint64_t counted;      /* process_aln returned 0                               */
    uint64_t no_contig;    /* returned 1: RNAME not in the genome                  */
    uint64_t filtered;     /* returned -1 (pss-bam) / 2 or -1 (fragkon)            */
    uint64_t parse_fail;   /* line2saml returned 1 (sam-parse.c:88-90)             */
    uint64_t undefined;    /* dropped because the reference's result depends on
                              stale buffer contents (see DESIGN.md "domain")       */
} pssgpu_stats;

/* Start a tally: zeroes the device tables.  Needs a resident genome. */
int pssgpu_pss_begin(pssgpu_ctx *ctx, const pssgpu_pss_params *p);

/* Feed SAM text (what `samtools view` writes: pss-bam.c:148-162), any
 * chunking; partial trailing lines are carried to the next call; `last` != 0
 * flushes a final line without '\n'.  Host memory; pinned memory from
 * pssgpu_host_alloc is copied without an extra staging pass.  Returns once the
 * bytes have been COPIED to the device (the caller may overwrite `sam_bytes`
 * right away, also after an error return); the tally kernels may still be
 * running -- they are waited for by pssgpu_sync, *_finish and get_stats, and a
 * kernel fault is reported there.  Lines of any length are accepted: one that
 * does not fit the staging buffer is tallied in the 200000-byte stretches
 * fgets() would hand to line2saml (pss-bam.c:761-764). */
int pssgpu_feed(pssgpu_ctx *ctx, const char *sam_bytes, size_t len, int last);

/* The same without waiting for the copies: `sam_bytes` (pinned memory from
 * pssgpu_host_alloc) must stay untouched until pssgpu_feed_wait(ctx) returns.
 * This is what lets one host thread keep the PCIe links of several GPUs busy at
 * once (host/pss_host.c deals its input chunks to the members of a group). */
int pssgpu_feed_async(pssgpu_ctx *ctx, const char *sam_bytes, size_t len, int last);
int pssgpu_feed_wait(pssgpu_ctx *ctx);

/* Same for SAM text already resident in device memory.  `d_sam` must be
 * 16-byte aligned and hold WHOLE lines (last byte '\n'); it must stay valid
 * until the next pssgpu_*_finish / pssgpu_sync.  No copy is made.  The bulk
 * copies of the kernel move 16-byte units: up to 15 bytes past `len` are
 * read (never interpreted) -- inside the allocation granule of any buffer that
 * came from cudaMalloc. */
int pssgpu_feed_device(pssgpu_ctx *ctx, const void *d_sam, size_t len);

/* ---- BAM input ----------------------------------------------------------------
 * The reference reads its BAM through popen("samtools view [-r RG] <bam>")
 * (pss-bam.c:148-162, fragkon.c:84-93).  pssgpu_feed_bam takes the BYTES OF THE BAM
 * FILE instead (BGZF blocks, any chunking, partial blocks are carried to the
 * next call; `last` != 0 marks the end of the file) and does on the device what
 * the samtools child did: BGZF inflate, BAM record framing, and the rendering of
 * every alignment as the SAM line the reference's fgets loop would have seen --
 * reduced to the fields line2saml / process_aln read (sam-parse.c:36-50) -- which
 * then goes through the same tally kernel as pssgpu_feed's text.  Tables are
 * identical to feeding `samtools view` output of the same file.  Returns once
 * the bytes have been copied to the device, like pssgpu_feed.  Malformed input
 * (bad BGZF / BAM structure, a file that ends inside a record) is reported by
 * pssgpu_sync / *_finish as PSSGPU_EINVAL; so is a block whose CRC32 does not
 * match its data (htslib checks it too; $PSSGPU_BAM_CRC=0 switches the check
 * off).  Do not mix with pssgpu_feed within one tally. */
int pssgpu_feed_bam(pssgpu_ctx *ctx, const void *bgzf_bytes, size_t len, int last);
/* `samtools view -r RG`, natively (pss-bam.c:153-155 `-R`): only alignments
 * whose RG:Z tag equals `read_group` are tallied; NULL switches the filter off.
 * Applies to pssgpu_feed_bam; sticky across tallies. */
int pssgpu_bam_read_group(pssgpu_ctx *ctx, const char *read_group);
typedef struct pssgpu_bam_stats {
    uint64_t records;                 /* alignment records seen since *_begin      */
    uint64_t dropped_by_read_group;   /* ... of which the -R filter dropped         */
    uint64_t references;              /* n_ref of the BAM header                   */
    uint64_t batches;                 /* device batches                            */
    uint64_t blocks_rewalked;         /* BGZF blocks whose first-record guess had to be corrected */
} pssgpu_bam_stats;
int pssgpu_bam_info(pssgpu_ctx *ctx, pssgpu_bam_stats *out);

/* Wait for all fed bytes to be tallied. */
int pssgpu_sync(pssgpu_ctx *ctx);

/* Finish: waits, then copies the tables out.  fwd/rev: (region_len+2)*16
 * counters each, row-major, same layout as the reference's count matrices
 * (pss-bam.c:24-35): row 0/1 = context bases 2/1 away, row i+2 = position i,
 * column 4*code(read)+code(ref), A=0 C=1 G=2 T=3.  May be called repeatedly;
 * the tally stays open for more pssgpu_feed calls until the next *_begin. */
int pssgpu_pss_finish(pssgpu_ctx *ctx, uint64_t *fwd, uint64_t *rev);
/* Same into DEVICE memory: d_tables receives 2*(region_len+2)*16 u64 (fwd
 * then rev), e.g. the buffer handed to ncclAllReduce(ncclSum, ncclUint64). */
int pssgpu_pss_finish_device(pssgpu_ctx *ctx, void *d_tables);

int pssgpu_get_stats(pssgpu_ctx *ctx, pssgpu_stats *out);


/* ---- fragkon ----------------------------------------------------------------
 * Options of fragkon.c:14-18. */
typedef struct pssgpu_fragkon_params {
    int           klen;         /* -k KLEN (default 8); 1 <= klen <= 14 */
    unsigned long min_len;      /* -l */
    unsigned long max_len;      /* -L */
    int           min_mq;       /* -q */
    int           merged_only;  /* -m */
} pssgpu_fragkon_params;

void pssgpu_fragkon_default_params(pssgpu_fragkon_params *p);
int  pssgpu_fragkon_begin(pssgpu_ctx *ctx, const pssgpu_fragkon_params *p);
/* (feed with pssgpu_feed / pssgpu_feed_device) */
/* fp/tp: 4^klen counters each, index = 2-bit MSB-first k-mer code
 * (kmer.c:184-214), UNSATURATED u64; the reference's `unsigned int` clamp
 * (kmer.c:102-104) is applied by the table writer. */
int  pssgpu_fragkon_finish(pssgpu_ctx *ctx, uint64_t *fp, uint64_t *tp);
int  pssgpu_fragkon_finish_device(pssgpu_ctx *ctx, void *d_tables /* 2*4^k u64: 5' then 3' */);

/* ---- both tallies from one scan of the text --------------------------------------
 * pss-bam and fragkon read the same SAM stream with the same parser and differ
 * only in filters and in what they count (pss-bam.c:390-496 vs fragkon.c:122-216);
 * running both over one `samtools view` pipe is the usual workflow.  After
 * pssgpu_both_begin every fed byte updates both table sets; read them with
 * pssgpu_pss_finish* and pssgpu_fragkon_finish*.  pssgpu_get_stats reports
 * pss-bam's outcomes, pssgpu_get_fragkon_stats fragkon's. */
int pssgpu_both_begin(pssgpu_ctx *ctx, const pssgpu_pss_params *pss, const pssgpu_fragkon_params *fragkon);
int pssgpu_get_fragkon_stats(pssgpu_ctx *ctx, pssgpu_stats *out);

/* ---- genome-kmer-count ---------------------------------------------------------
 * Whole-genome forward-strand k-mer spectrum over every contig
 * (genome-kmer-count.c:56-58,68-79).  counts: 4^k u64, unsaturated.
 * shard/n_shards split the packed genome into contiguous ranges (k-mers are
 * attributed to the shard holding their first base; contig boundaries are
 * respected by construction); summing the shards' outputs gives the full
 * spectrum.  1 <= k <= 14. */
int pssgpu_kmer_spectrum(pssgpu_ctx *ctx, int k, uint64_t *counts);
int pssgpu_kmer_spectrum_shard(pssgpu_ctx *ctx, int k, int shard, int n_shards, uint64_t *counts);
int pssgpu_kmer_spectrum_shard_device(pssgpu_ctx *ctx, int k, int shard, int n_shards, void *d_counts);

/* ---- several GPUs of one box ---------------------------------------------------
 * reads shard, the genome replicates, the tables are summed once at
 * the end.  A group is one context per GPU in ONE process (the host programs
 * use it when $PSSGPU_DEVICES names more than one GPU; bench.py instead runs
 * one process per GPU under torchrun and sums with torch.distributed).  The sum
 * is an ncclAllReduce(ncclSum, ncclUint64) over NVLink -- NCCL is loaded at run
 * time -- or, without NCCL / with PSSGPU_GROUP_REDUCE=peer, peer copies to the
 * first GPU and an add kernel there.  devices == NULL or n <= 0: all visible
 * GPUs.  Group calls report through pssgpu_group_last_error. */
typedef struct pssgpu_group pssgpu_group;
int         pssgpu_group_init(const int *devices, int n, pssgpu_group **out);
void        pssgpu_group_destroy(pssgpu_group *g);
int         pssgpu_group_size(const pssgpu_group *g);
pssgpu_ctx *pssgpu_group_ctx(pssgpu_group *g, int i);          /* member i, e.g. for pssgpu_feed_async / pssgpu_feed_bam */
const char *pssgpu_group_last_error(const pssgpu_group *g);
const char *pssgpu_group_reduce_backend(const pssgpu_group *g);
double      pssgpu_group_last_reduce_ms(const pssgpu_group *g); /* device time of the last table sum */
int pssgpu_group_genome_upload(pssgpu_group *g, const pssgpu_contig *contigs, uint64_t n_contigs);
int pssgpu_group_genome_load_tagged(pssgpu_group *g, const char *path, const pssgpu_genome_tag *expect);
int pssgpu_group_pss_begin(pssgpu_group *g, const pssgpu_pss_params *p);
int pssgpu_group_fragkon_begin(pssgpu_group *g, const pssgpu_fragkon_params *p);
int pssgpu_group_both_begin(pssgpu_group *g, const pssgpu_pss_params *pss, const pssgpu_fragkon_params *fragkon);
/* Whole lines (a piece ends with '\n' unless `last`); pieces go to the members in turn. */
int pssgpu_group_feed(pssgpu_group *g, const char *sam_bytes, size_t len, int last);
int pssgpu_group_sync(pssgpu_group *g);
/* A BAM FILE over the GPUs of a group (pssgpu_feed_bam's contract: the file's
 * bytes in any chunking, `last` marks the end).  Records run across BGZF blocks,
 * so the stream cannot be dealt like text; the inflate -- five sixths of the
 * ingest -- can: batches of BGZF blocks go to the GPU with the fewest in flight,
 * each inflates its batches, member 0 fetches the inflated bytes with a peer
 * copy and frames, renders and tallies them in file order (the other members'
 * tables stay zero; the sums are the same).  Replaces the same samtools child
 * (pss-bam.c:148-162, fragkon.c:84-93).
 * batches_per_member: may be NULL; one entry per member otherwise. */
int pssgpu_group_feed_bam(pssgpu_group *g, const void *bgzf_bytes, size_t len, int last);
int pssgpu_group_bam_read_group(pssgpu_group *g, const char *read_group);
int pssgpu_group_bam_info(pssgpu_group *g, pssgpu_bam_stats *out, uint64_t *batches_per_member);
int pssgpu_group_pss_finish(pssgpu_group *g, uint64_t *fwd, uint64_t *rev);
int pssgpu_group_fragkon_finish(pssgpu_group *g, uint64_t *fp, uint64_t *tp);
int pssgpu_group_get_stats(pssgpu_group *g, pssgpu_stats *out, int fragkon);
/* genome-kmer-count.c:56-58 genome-sharded: member i counts slice i, then the 4^k counters are summed. */
int pssgpu_group_kmer_spectrum(pssgpu_group *g, int k, uint64_t *counts);

/* ---- measurement hooks ------------------------------------------------------------
 * CUDA-event timing of the kernels this library launches, on the stream it
 * launches them on (a caller's torch.cuda.Event only sees torch's stream).
 * pssgpu_timing_reset zeroes the accumulators; every tally / spectrum / pack
 * launch afterwards is bracketed by events. */
typedef struct pssgpu_timing {
    uint64_t launches;        /* kernel launches since reset               */
    double   kernel_ms;       /* sum of their event-measured durations     */
    uint64_t bytes_scanned;   /* SAM bytes (or packed-genome bytes) they read */
    uint64_t h2d_bytes;       /* bytes copied host->device by pssgpu_feed  */
    uint64_t d2h_bytes;       /* bytes copied device->host by *_finish     */
} pssgpu_timing;

int pssgpu_timing_reset(pssgpu_ctx *ctx, int enable);
int pssgpu_timing_get(pssgpu_ctx *ctx, pssgpu_timing *out);

/* ---- test hook ------------------------------------------------------------------------
 * Per-record outcomes for parity debugging: after pssgpu_debug_status(ctx,1)
 * every processed line appends (byte offset of the line within everything fed
 * since *_begin, outcome code) to a device log; pssgpu_debug_fetch copies up
 * to `cap` entries out (unordered) and returns the count in *n.
 * Outcome codes: 0 counted, 1 no contig, -1 filtered, -2 parse fail, -3 undefined. */
int pssgpu_debug_status(pssgpu_ctx *ctx, int enable);
int pssgpu_debug_fetch(pssgpu_ctx *ctx, uint64_t *offsets, int8_t *codes, uint64_t cap, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* PSSGPU_H */
