/*
 * oracle/ref_glue_jfif.c — TEST INFRASTRUCTURE.  Buffer-level driver around the REFERENCE's vendored baseline-JPEG
 * writer, stbi_write_jpg_to_func (Algorithms/sequential/JPEG/stb_image_write.h:1607, core :1398-1605).  The header is
 * #included from where it lies under /root/reference (REF_STBW, set by oracle/build.py: a temporary copy in which the
 * one line `subsample = quality <= 90 ? 1 : 0;` reads a global override first, deleted after compilation) — no
 * reference source is stored in this repository.  Output: oracle/_ref/libref_jfif.so (git-ignored).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

__thread int ljb_force_subsample = -1; /* read by the patched line; -1 keeps stb's own rule */

#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STBI_WRITE_NO_STDIO
#define STB_IMAGE_WRITE_STATIC
#include REF_STBW

typedef struct {
    uint8_t *out;
    size_t cap, n;
} mem_sink;

static void sink_write(void *ctx, void *data, int size)
{
    mem_sink *s = (mem_sink *)ctx;
    if (s->out && s->n + (size_t)size <= s->cap) memcpy(s->out + s->n, data, (size_t)size);
    s->n += (size_t)size;
}

/* px: h rows of w pixels of `comp` bytes, tightly packed (stb has no stride parameter). */
int ref_jfif_encode(const uint8_t *px, int w, int h, int comp, int quality, int force_subsample, uint8_t *out, size_t cap,
                    size_t *out_len)
{
    mem_sink s = {out, cap, 0};
    ljb_force_subsample = force_subsample;
    const int ok = stbi_write_jpg_to_func(sink_write, &s, w, h, comp, px, quality);
    ljb_force_subsample = -1;
    if (out_len) *out_len = s.n;
    if (!ok) return -1;
    return s.n <= cap ? 0 : -3;
}

/* CPU-baseline timing: `nthreads` host threads each encode the whole image `reps` times into a counting sink
 * (stb's encoder is one sequential bit stream per image, so images are the unit of host parallelism).
 * Returns seconds of wall time around the encode loops only. */
typedef struct {
    const uint8_t *px;
    int w, h, comp, quality, force, reps;
    size_t bytes;
} mt_arg;

static void *mt_worker(void *p)
{
    mt_arg *a = (mt_arg *)p;
    ljb_force_subsample = a->force;
    for (int r = 0; r < a->reps; ++r) {
        mem_sink s = {NULL, 0, 0};
        stbi_write_jpg_to_func(sink_write, &s, a->w, a->h, a->comp, a->px, a->quality);
        a->bytes = s.n;
    }
    return NULL;
}

double ref_jfif_time_mt(const uint8_t *px, int w, int h, int comp, int quality, int force_subsample, int nthreads, int reps,
                        size_t *bytes_per_image)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mt_arg *args = (mt_arg *)malloc(sizeof(mt_arg) * (size_t)nthreads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < nthreads; ++i) {
        args[i] = (mt_arg){px, w, h, comp, quality, force_subsample, reps, 0};
        pthread_create(&th[i], NULL, mt_worker, &args[i]);
    }
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (bytes_per_image) *bytes_per_image = args[0].bytes;
    free(th);
    free(args);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
