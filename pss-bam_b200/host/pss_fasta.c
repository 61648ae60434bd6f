/* pss_fasta.c -- block-buffered FASTA reader behind the fasta-genome-io.h API. */
#include "pss_fasta.h"

#define BLOCK (4u << 20)

/* ---- byte source: one refill per 4 MiB instead of one libc call per byte ---- */
typedef struct bytesrc {
    FILE          *fp;
    gzFile         gz;
    unsigned char *buf;
    size_t         have, at;
    int            eof;
} bytesrc;

static int src_peek(bytesrc *s)
{
    if (s->at == s->have) {
        if (s->eof) return EOF;
        s->have = s->gz ? (size_t)(gzread(s->gz, s->buf, BLOCK) > 0 ? gztell(s->gz), 0 : 0) : 0;   /* placeholder, replaced below */
    }
    return s->buf[s->at];
}
#undef BLOCK
