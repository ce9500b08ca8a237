/*
 * oracle/ref_glue_jpeg_par.c — TEST INFRASTRUCTURE.  CPU-baseline driver around the reference's PARALLEL build: the fused
 * per-group thread body process() (Algorithms/parallel/JPEG/JPEG.c:1103-1252: forward chain AND inverse chain), #included
 * from where it lies under /root/reference (REF_SRC) and compiled against oracle/shim/windows.h.
 * The reference starts one OS thread per 8x8 group (:1297-1302) and hands each a COPY of its group (args->block = blocks[i],
 * :1300), so the results are discarded — reproduced as is: this times what the reference's JPEG_par.exe computes, from a pool
 * of `nthreads` workers instead of a thread per group (4.2 M threads at 16384 x 16384 are not runnable).
 */
#define _GNU_SOURCE
#define main ref_jpegpar_main
#include REF_SRC
#undef main
#include <time.h>

typedef struct {
    PixelGroup *blocks;
    size_t total;
    int tid, nthreads;
} jpar_arg;

static void *jpar_worker(void *p)
{
    jpar_arg *a = (jpar_arg *)p;
    for (size_t i = (size_t)a->tid; i < a->total; i += (size_t)a->nthreads) {
        parallel_args *args = malloc(sizeof(parallel_args)); /* P-JPG:1299-1300 */
        args->block = a->blocks[i];
        process(args);
        free(args->block.lum_coefficients); /* the reference leaks these; a benchmark loop cannot */
        free(args->block.r_coefficients);
        free(args->block.b_coefficients);
        free(args);
    }
    return NULL;
}

int ref_jpegpar_time_groups(const uint8_t *rgba, int w, int h, size_t stride, int nthreads, double *seconds)
{
    if (nthreads < 1) nthreads = 1;
    ImageData im;
    im.height = h;
    im.width = w;
    im.pixel_count = (size_t)h * w;
    im.pixels = malloc(sizeof(Pixel *) * (size_t)h);
    for (int y = 0; y < h; y++) {
        im.pixels[y] = malloc(sizeof(Pixel) * (size_t)w);
        memcpy(im.pixels[y], rgba + (size_t)y * stride, sizeof(Pixel) * (size_t)w);
    }
    uint8_t **ym, **rm, **bm;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    build_luminance_matrix(im, &ym); /* P-JPG:1265-1289 */
    build_rChrominance_matrix(im, &rm);
    build_bChrominance_matrix(im, &bm);
    chroma_subsample(&bm, im);
    chroma_subsample(&rm, im);
    size_t total_blocks = (size_t)ceil((double)im.pixel_count / 64);
    PixelGroup *blocks = divide_image(ym, rm, bm, im, 8);
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    jpar_arg *args = calloc((size_t)nthreads, sizeof(jpar_arg));
    for (int t = 0; t < nthreads; t++) {
        args[t] = (jpar_arg){blocks, total_blocks, t, nthreads};
        pthread_create(&th[t], NULL, jpar_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(th);
    free(args);
    free(blocks);
    for (int y = 0; y < h; y++) {
        free(ym[y]);
        free(rm[y]);
        free(bm[y]);
        free(im.pixels[y]);
    }
    free(ym);
    free(rm);
    free(bm);
    free(im.pixels);
    return 0;
}
