/*
 * oracle/ref_glue_lz4.c — TEST INFRASTRUCTURE.  Buffer-level driver around the REFERENCE's own
 * block_encode()/write_block() (Algorithms/sequential/LZ4/LZ4.c:506, :415).  The reference source is
 * #included from where it lies under /root/reference (REF_SRC, set by oracle/build.py; for the
 * "bounded" build a patched temporary copy that is deleted after compilation) — no reference source
 * is stored in this repository.  Output: oracle/_ref/libref_lz4*.so (git-ignored).
 */
#define _GNU_SOURCE
#define main ref_lz4_main
#include REF_SRC
#undef main
#include <pthread.h>

/* Encode in[0..n) exactly as lz4_encode() does (S-LZ4:699-727): divide_input -> block_encode per block
 * -> write_output layout, but into memory and recording true block offsets. */
int ref_lz4_compress(const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                     uint64_t *block_offsets, size_t *out_len)
{
    size_t block_count = 0;
    char **blocks = divide_input(in, n, block_len, &block_count); /* exact-size mallocs, S-LZ4:123 */
    LZ4Frame frame;
    frame.blocks = 0;
    frame.frame_blocks = NULL;
    for (size_t i = 0; i < block_count; i++) {
        LZ4Block cur = {0};
        size_t len = (i == block_count - 1) ? n - i * block_len : block_len;
        block_encode(blocks[i], len, &cur, NULL, NULL, &frame);
    }
    char *mem = NULL;
    size_t memlen = 0;
    FILE *f = open_memstream(&mem, &memlen);
    fwrite(&frame.blocks, sizeof(uint8_t), 1, f); /* S-LZ4:429 */
    for (size_t i = 0; i < frame.blocks; i++) {
        fflush(f);
        if (block_offsets) block_offsets[i] = (uint64_t)ftell(f);
        write_block(&frame.frame_blocks[i], f); /* S-LZ4:433 */
        free(frame.frame_blocks[i].sequences);
    }
    fflush(f);
    if (block_offsets) block_offsets[frame.blocks] = (uint64_t)ftell(f);
    fclose(f);
    free(frame.frame_blocks);
    for (size_t i = 0; i < block_count; i++) free(blocks[i]);
    free(blocks);
    int rc = 0;
    if (memlen > out_cap) rc = -1;
    else memcpy(out, mem, memlen);
    if (out_len) *out_len = memlen;
    free(mem);
    return rc;
}

/* CPU-baseline timing: the reference's per-block hot path (block_encode) over all blocks on
 * `nthreads` host threads (static round-robin), wall time around the block loop only. */
typedef struct {
    const uint8_t *in;
    size_t n, block_len, block_count;
    int tid, nthreads;
    uint64_t bytes_out;
} ref_mt_arg;
static void *ref_mt_worker(void *p)
{
    ref_mt_arg *a = (ref_mt_arg *)p;
    for (size_t i = (size_t)a->tid; i < a->block_count; i += (size_t)a->nthreads) {
        size_t len = (i == a->block_count - 1) ? a->n - i * a->block_len : a->block_len;
        char *blk = (char *)malloc(len); /* exact-size copy as divide_input does (S-LZ4:156-171) */
        memcpy(blk, a->in + i * a->block_len, len);
        LZ4Block cur = {0};
        LZ4Frame frame;
        frame.blocks = 0;
        frame.frame_blocks = NULL;
        block_encode(blk, len, &cur, NULL, NULL, &frame);
        a->bytes_out += frame.frame_blocks[0].byte_size;
        free(frame.frame_blocks[0].sequences);
        free(frame.frame_blocks);
        free(blk);
    }
    return NULL;
}
int ref_lz4_time_blocks(const uint8_t *in, size_t n, size_t block_len, int nthreads, double *seconds,
                        uint64_t *bytes_out)
{
    if (nthreads < 1) nthreads = 1;
    size_t block_count = (n + block_len - 1) / block_len;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    ref_mt_arg *args = calloc((size_t)nthreads, sizeof(ref_mt_arg));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        args[t] = (ref_mt_arg){in, n, block_len, block_count, t, nthreads, 0};
        pthread_create(&th[t], NULL, ref_mt_worker, &args[t]);
    }
    uint64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += args[t].bytes_out;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (bytes_out) *bytes_out = total;
    free(th);
    free(args);
    return 0;
}
