/* pss_io.h -- reading a regular file fast enough for the GPU (no counterpart in the reference, which reads a pipe from
 * `samtools view`, pss-bam.c:148-162: here the bytes of the BAM file itself are the input). */
#ifndef PSS_IO_H
#define PSS_IO_H
#include <stddef.h>
#include <sys/types.h>

/* Read up to n bytes at offset off of fd into buf with `threads` concurrent pread()s over disjoint slices (one
 * thread copies out of the page cache at 3-6 GB/s; the BAM ingest takes 17 GB/s per GPU).  Returns the number of bytes
 * read contiguously from off (less than n only at the end of the file or on an I/O error). */
size_t pss_pread_parallel(int fd, char *buf, size_t n, off_t off, int threads);

#endif
