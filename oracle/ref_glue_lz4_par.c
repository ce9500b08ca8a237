/*
 * oracle/ref_glue_lz4_par.c — TEST INFRASTRUCTURE.  CPU-baseline driver around the reference's PARALLEL build:
 * parallel_block_encode() (Algorithms/parallel/LZ4/LZ4.c:518-628) with its global add_seq / block_plus locks
 * (:28-31, :584-587, :612-616, :508-510), #included from where it lies under /root/reference (REF_SRC; the bounded
 * variant is a patched temporary copy, see oracle/build.py) and compiled against oracle/shim/windows.h.
 * The reference starts one OS thread per block (:724-749); at 65 536 blocks that is not runnable, so the same thread
 * body is called from a pool of `nthreads` workers — stated in bench.py's cpu_baseline sample text.
 */
#define _GNU_SOURCE
#define main ref_lz4par_main
#include REF_SRC
#undef main
#include <time.h>

typedef struct {
    const uint8_t *in;
    size_t n, block_len, block_count;
    int tid, nthreads;
    LZ4Frame *frame;
} par_arg;

static void *par_worker(void *p)
{
    par_arg *a = (par_arg *)p;
    for (size_t i = (size_t)a->tid; i < a->block_count; i += (size_t)a->nthreads) {
        size_t len = (i == a->block_count - 1) ? a->n - i * a->block_len : a->block_len;
        char *blk = (char *)malloc(len); /* exact-size copy as divide_input does (P-LZ4:123-177) */
        memcpy(blk, a->in + i * a->block_len, len);
        BlockEncodeArgs *args = malloc(sizeof(BlockEncodeArgs)); /* as parallel_LZ4_encode prepares them, P-LZ4:728-739 */
        args->block_entry = blk;
        args->block_length = len;
        args->block = malloc(sizeof(LZ4Block));
        args->log_file = NULL;
        args->output_file = NULL;
        args->frame = a->frame;
        args->index = i;
        LZ4Block *keep = args->block;
        parallel_block_encode(args); /* frees args (P-LZ4:625) */
        free(keep);
        free(blk);
    }
    return NULL;
}

/* returns the wall time of the block loop and the sum of the blocks' byte sizes (a checksum the caller compares with the
 * sequential build's) */
int ref_lz4par_time_blocks(const uint8_t *in, size_t n, size_t block_len, int nthreads, double *seconds, uint64_t *bytes_out)
{
    if (nthreads < 1) nthreads = 1;
    size_t block_count = (n + block_len - 1) / block_len;
    LZ4Frame frame;
    frame.blocks = 0;
    frame.frame_blocks = malloc(sizeof(LZ4Block) * block_count); /* pre-sized, P-LZ4:708 */
    InitializeCriticalSection(&block_plus); /* main() does this, P-LZ4:1229-1231 */
    InitializeCriticalSection(&add_seq);
    InitializeCriticalSection(&seq_decode);
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    par_arg *args = calloc((size_t)nthreads, sizeof(par_arg));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        args[t] = (par_arg){in, n, block_len, block_count, t, nthreads, &frame};
        pthread_create(&th[t], NULL, par_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    uint64_t total = 0;
    for (size_t i = 0; i < block_count; i++) {
        total += frame.frame_blocks[i].byte_size;
        free(frame.frame_blocks[i].sequences);
    }
    if (bytes_out) *bytes_out = total;
    free(frame.frame_blocks);
    free(th);
    free(args);
    return frame.blocks == block_count ? 0 : -1;
}
