/* pss_host.h -- glue between the C API types and the GPU library. */
#ifndef PSS_HOST_H
#define PSS_HOST_H

#include <stdio.h>

#include "../../include/pssgpu.h"
#include "pss_fasta.h"

/* Open the device named by $PSSGPU_DEVICE (default 0); exits with a message when there is none. */
pssgpu_ctx *pss_open_device(void);
/* Make init_genome()'s result device resident (pssgpu_genome_upload). */
int pss_upload_genome(pssgpu_ctx *ctx, const Genome *genome);
/* init_genome(fasta_fn) + pss_upload_genome(), or -- when $PSSGPU_GENOME_CACHE is set -- the packed-genome cache:
 * "1" keeps it beside the FASTA as <fasta>.pssgpu, any other value names a directory for <basename>.pssgpu.  A cache
 * at least as new as the FASTA is loaded instead of parsing the FASTA (read_fasta's fgetc loop is the reference's
 * ~150 s for 3.1 Gb, fasta-genome-io.c:105-148); otherwise it is (re)written after the upload.  *n_seqs receives
 * the number of FASTA records.  Exits with a message when the FASTA cannot be read. */
void pss_resident_genome(pssgpu_ctx *ctx, const char *fasta_fn, unsigned long *n_seqs);
/* popen("samtools view [-r RG] <bam>") -- the reference's only source of SAM text (pss-bam.c:148-162). */
FILE *pss_bam_to_sam(const char *bam_fn, const char *read_group);
/* Pump a SAM text stream into the open tally: fread into pinned memory, pssgpu_feed, flush at EOF.
 * Replaces the fgets + line2saml + process_aln loop (pss-bam.c:764-783, fragkon.c:342-363). */
int pss_stream_sam(pssgpu_ctx *ctx, FILE *sam);
/* The -B argument, whatever it is: a BGZF file (a real BAM) is read as it is and decoded on the GPU
 * (pssgpu_feed_bam; -R through pssgpu_bam_read_group) -- no samtools needed; anything else goes through
 * `samtools view [-r RG]` like in the reference (pss-bam.c:148-162).  $PSSGPU_USE_SAMTOOLS forces the pipe. */
int pss_stream_input(pssgpu_ctx *ctx, const char *bam_fn, const char *read_group);
int pss_is_bgzf(const char *fn);
void pss_die(pssgpu_ctx *ctx, const char *what);

/* ---- several GPUs ($PSSGPU_DEVICES = "all" | "0,1,..."; unset: one GPU, $PSSGPU_DEVICE or 0) ---- */
pssgpu_group *pss_open_devices(void);
void pss_resident_genome_group(pssgpu_group *g, const char *fasta_fn, unsigned long *n_seqs);
/* SAM text is dealt to the members line-wise, chunk by chunk (reads shard); a BAM file goes to member 0. */
int  pss_stream_input_group(pssgpu_group *g, const char *bam_fn, const char *read_group);
void pss_die_group(pssgpu_group *g, const char *what);

#endif
