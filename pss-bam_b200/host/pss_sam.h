/* pss_sam.h -- SAM record type and line parser of the host side.
 *
 * Same API and the same struct layout as the reference's sam-parse.h:20-61
 * (a program written against that header compiles and links unchanged against
 * this one; `sam-parse.h` in this directory simply includes this file).  The
 * parser itself is new: instead of one unbounded sscanf into fixed 2048-byte
 * fields (sam-parse.c:36-48) it runs a bounded scanner that applies glibc's
 * conversion rules for "%s %u %s %lu %u %s %s %u %i %s %s" by hand and
 * refuses tokens that would not fit a field.
 *
 * The GPU path does not call line2saml(); it parses records on the device
 * (csrc/pss_record.h) with the same rules.  line2saml stays for callers of
 * the C API and for the host-side tools.
 */
#ifndef PSS_SAM_H
#define PSS_SAM_H

#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define MAX_LINE_LEN    (200000)   /* longest line the programs hand to the parser */
#define MAX_FN_LEN      (2047)
#define MAX_FIELD_WIDTH (2047)     /* longest SAM field kept */
/* scoring constants of the reference API (unused by the three programs) */
#define MATCH    (1)
#define MISMATCH (4)
#define GAP_OPEN (6)
#define GAP_EXT  (1)

#define PSS_SAM_TEXT(name) char name[MAX_FIELD_WIDTH + 1]

/* One alignment.  Field order and types follow sam-parse.h:20-56 so that the
 * layout (sizeof == 20536 on LP64) is interchangeable. */
typedef struct saml {
    PSS_SAM_TEXT(qname);
    unsigned int flag;
    /* FLAG bits 0x1 .. 0x800, unpacked by line2saml */
    unsigned int paired : 1, proper_pair : 1, unmap : 1, munmap : 1, reverse : 1, mreverse : 1;
    unsigned int read1 : 1, read2 : 1, secondary : 1, qc_failed : 1, duplicate : 1, supplementary : 1;
    PSS_SAM_TEXT(rname);
    unsigned long pos;             /* 1-based leftmost position */
    unsigned int mapq;
    PSS_SAM_TEXT(cigar);
    PSS_SAM_TEXT(mrnm);
    unsigned int mpos;
    int isize;                     /* TLEN; for unpaired records: strlen(seq) */
    int seq_len;
    PSS_SAM_TEXT(seq);
    PSS_SAM_TEXT(qual);
    PSS_SAM_TEXT(tags);            /* everything after the 11th tab */
    PSS_SAM_TEXT(BC);
    PSS_SAM_TEXT(RG);
    PSS_SAM_TEXT(opt_tags);
    int aln_seq_len;
    int NM, AS, XM, XO, XG;        /* never filled by line2saml (as in the reference) */
} Saml;

/* 0: *sp filled; 1: not a usable alignment line (fewer than 11 fields, a
 * number that does not convert, SEQ and QUAL of different length, or a field
 * longer than MAX_FIELD_WIDTH). */
int line2saml(const char *line, Saml *sp);
/* non-zero when the line starts with '@' */
int is_header(const char *line);
/* sum of the lengths of the M operations of a CIGAR string */
int aln_seq_len(const char *cigar);
/* 1 unless sp->AS > 0 and AS < m * seq_len + b */
int good_score(Saml *sp, float m, float b);

#endif
