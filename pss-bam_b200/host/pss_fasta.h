/* pss_fasta.h -- FASTA genome loader and contig lookup of the host side.
 *
 * API and struct layouts of the reference's fasta-genome-io.h:13-46
 * (`fasta-genome-io.h` here includes this file), new implementation: the
 * file is read in 4 MiB blocks (fread / gzread) through a small byte source
 * instead of one fgetc()/gzgetc() call per character
 * (fasta-genome-io.c:105-200), which is what makes a 3.1 Gb load take
 * minutes in the reference.  Record semantics are unchanged:
 *   id   = bytes after '>' up to the first isspace() byte (at most MAX_ID_LEN)
 *   seq  = every non-space byte up to the next '>' or the end, upper-cased
 *   contigs sorted by strcmp(id); find_seq = binary search.
 */
#ifndef PSS_FASTA_H
#define PSS_FASTA_H

#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#ifndef MAX_FN_LEN
#define MAX_FN_LEN (2047)
#endif
#define MAX_ID_LEN      (511)
#define MAX_SEQ_LEN     (536870911)    /* longer contigs are truncated (with a warning), as in the reference */
#define MAX_GENOME_SEQS (1000000)

typedef struct seq {
    char   id[MAX_ID_LEN + 1];
    char  *seq;       /* upper-cased, NUL terminated, exactly len + 1 bytes */
    size_t len;
} Seq;

typedef struct genome {
    Seq  **seqs;      /* n_seqs pointers, ordered by strcmp(id) once init_genome returns */
    Seq   *dummy;     /* scratch key for find_seq (makes find_seq non re-entrant, as in the reference) */
    size_t n_seqs;
} Genome;

typedef struct fa_src {
    char   fn[MAX_FN_LEN + 1];
    char  *seq_buffer;   /* here: the state of the block reader */
    int    is_gz;
    gzFile fagz;
    FILE  *fafp;
    size_t n;            /* records handed out so far */
} Fa_Src;

Genome *init_genome(const char fn[]);               /* NULL if the file cannot be opened */
Fa_Src *init_fasta_src(const char fn[]);
Seq    *get_next_fa(Fa_Src *fa_source, Genome *genome);   /* NULL at the end of the file */
int     read_fasta(FILE *fp, Seq *seq, char *seq_buffer);        /* 0 ok, -1 end of file */
int     gzread_fasta(gzFile gzfp, Seq *seq, char *seq_buffer);
Seq    *find_seq(Genome *genome, const char id[]);  /* borrowed pointer or NULL */
int     is_gz(const char *fn);
FILE   *fileOpen(const char *name, char access_mode[]);
int     close_fasta_src(Fa_Src *);
int     chr_cmp(const void *v1, const void *v2);
int     destroy_seq(Seq *seq);
int     destroy_genome(Genome *genome);

#endif
