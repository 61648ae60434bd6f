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
/* popen("samtools view [-r RG] <bam>") -- the reference's only source of SAM text (pss-bam.c:148-162). */
FILE *pss_bam_to_sam(const char *bam_fn, const char *read_group);
/* Pump a SAM text stream into the open tally: fread into pinned memory, pssgpu_feed, flush at EOF.
 * Replaces the fgets + line2saml + process_aln loop (pss-bam.c:764-783, fragkon.c:342-363). */
int pss_stream_sam(pssgpu_ctx *ctx, FILE *sam);
void pss_die(pssgpu_ctx *ctx, const char *what);

#endif
