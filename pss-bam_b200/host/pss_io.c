/* pss_io.c -- see pss_io.h */
#define _GNU_SOURCE
#include "pss_io.h"

#include <errno.h>
#include <pthread.h>
#include <unistd.h>

#define PSS_IO_MAX_THREADS 16

typedef struct slice {
    int    fd;
    char  *buf;
    size_t n, got;
    off_t  off;
} slice;

static void *slice_reader(void *arg)
{
    slice *s = (slice *)arg;
    size_t done = 0;
    while (done < s->n) {
        const ssize_t r = pread(s->fd, s->buf + done, s->n - done, s->off + (off_t)done);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) break;                                   /* end of file, or an error: the caller sees a short count */
        done += (size_t)r;
    }
    s->got = done;
    return NULL;
}

size_t pss_pread_parallel(int fd, char *buf, size_t n, off_t off, int threads)
{
    slice     s[PSS_IO_MAX_THREADS];
    pthread_t th[PSS_IO_MAX_THREADS];
    int       started[PSS_IO_MAX_THREADS];
    size_t    per, at = 0, total = 0;
    int       k = 0, i;
    if (n == 0) return 0;
    if (threads < 1) threads = 1;
    if (threads > PSS_IO_MAX_THREADS) threads = PSS_IO_MAX_THREADS;
    per = (n + (size_t)threads - 1) / (size_t)threads;
    per = (per + 0xfffffu) & ~(size_t)0xfffffu;             /* slices of whole MiB */
    while (at < n && k < threads) {
        s[k].fd = fd; s[k].buf = buf + at; s[k].off = off + (off_t)at; s[k].got = 0;
        s[k].n = n - at < per ? n - at : per;
        at += s[k].n;
        k++;
    }
    for (i = 1; i < k; i++) started[i] = pthread_create(&th[i], NULL, slice_reader, &s[i]) == 0;
    slice_reader(&s[0]);
    for (i = 1; i < k; i++) {
        if (started[i]) pthread_join(th[i], NULL);
        else slice_reader(&s[i]);                            /* no thread to be had: read the slice here */
    }
    for (i = 0; i < k; i++) {                                /* the contiguous prefix */
        total += s[i].got;
        if (s[i].got < s[i].n) break;
    }
    return total;
}
