/* compat/test_compat.c — a program written against the reference's LZ4 entry points (Algorithms/sequential/LZ4/LZ4.c) the way
 * its own lz4_encode() uses them: divide_input -> block_encode per block -> write_output.  Linked against libljb_compat.so.
 *
 *   test_compat blocks  <input> <block_len> <out.bin>   the frame through block_encode()/write_output()
 *   test_compat par     <input> <block_len> <out.bin>   the same through the parallel build's thread body (stored by index)
 *   test_compat match   <input> <block_len>             find_longest_match() at every position of the first block against the
 *                                                        exhaustive scan of LZ4.c:290-323 (bounded at the block end); prints mismatches
 *   test_compat files                                    lz4_encode() + LZ4_decode() on the reference's fixed relative paths
 *   test_compat process <samples.bin> <out.bin>          process() (Algorithms/parallel/JPEG/JPEG.c:1103) on every 128-byte group of the
 *                                                        file (lum 64 | b 32 | r 32); writes per group the 128 reconstructed samples and
 *                                                        the 128 dequantised coefficients (doubles: lum 64 | r 32 | b 32)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ljb_compat.h"

static uint8_t *read_all(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = malloc((size_t)sz + 1);
    if (fread(b, 1, (size_t)sz, f) != (size_t)sz) exit(2);
    fclose(f);
    *n = (size_t)sz;
    return b;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && strcmp(argv[1], "files") == 0) {
        FILE *c = fopen("../Output-Input/out/compressed.bin", "wb"); /* clear_files, LZ4.c:204-213 */
        if (c) fclose(c);
        lz4_encode();
        LZ4_decode("../Output-Input/out/compressed.bin", "../Output-Input/log/encoding_log.txt");
        return 0;
    }
    if (argc >= 4 && strcmp(argv[1], "process") == 0) {
        size_t n = 0;
        uint8_t *smp = read_all(argv[2], &n);
        FILE *out = fopen(argv[3], "wb");
        if (!out) return 2;
        for (size_t g = 0; g < n / 128; ++g) {
            parallel_args *args = malloc(sizeof(parallel_args)); /* as the reference's main() prepares them, P-JPG:1299-1300 */
            memset(args, 0, sizeof *args);
            memcpy(args->block.lum_values, smp + g * 128, 64);
            memcpy(args->block.b_values, smp + g * 128 + 64, 32);
            memcpy(args->block.r_values, smp + g * 128 + 96, 32);
            process(args);
            fwrite(args->block.lum_values, 1, 64, out);
            fwrite(args->block.b_values, 1, 32, out);
            fwrite(args->block.r_values, 1, 32, out);
            fwrite(args->block.lum_coefficients, sizeof(double), 64, out);
            fwrite(args->block.r_coefficients, sizeof(double), 32, out);
            fwrite(args->block.b_coefficients, sizeof(double), 32, out);
            free(args->block.lum_coefficients);
            free(args->block.r_coefficients);
            free(args->block.b_coefficients);
            free(args);
        }
        fclose(out);
        return 0;
    }
    if (argc < 4) return 2;
    size_t n = 0;
    uint8_t *input = read_all(argv[2], &n);
    const size_t block_size = (size_t)atol(argv[3]);
    ljb_compat_set_block_length(block_size);
    if (strcmp(argv[1], "match") == 0) {
        const size_t len = n < block_size ? n : block_size;
        ljb_compat_bind_block(input, len);
        size_t bad = 0;
        for (size_t cur = 0; cur < len; ++cur) {
            size_t best = 0, best_dist = 0;
            for (size_t i = 0; i < cur; ++i) { /* LZ4.c:297-311, extension bounded at the block end */
                size_t l = 0;
                while (l < 1024 && cur + l < len && input[i + l] == input[cur + l]) ++l;
                if (l > best) {
                    best = l;
                    best_dist = cur - i;
                }
            }
            uint16_t dist = 0;
            const uint8_t got = find_longest_match(input, cur, &dist);
            const uint8_t want = best >= 4 ? (uint8_t)best : 0;
            if (got != want || (want && best != 1024 && dist != (uint16_t)best_dist)) {
                if (bad < 5) printf("position %zu: got (%u, %u), want (%u, %zu)\n", cur, got, dist, want, best_dist);
                ++bad;
            }
        }
        printf("%zu positions, %zu mismatches\n", len, bad);
        return bad ? 1 : 0;
    }
    const int par = strcmp(argv[1], "par") == 0;
    size_t block_count = 0;
    char **blocks = divide_input(input, n, block_size, &block_count); /* LZ4.c:704 */
    LZ4Frame frame;
    frame.blocks = 0;
    frame.frame_blocks = par ? malloc(sizeof(LZ4Block) * block_count) : NULL; /* pre-sized in the parallel build, P-LZ4:708 */
    for (size_t i = 0; i < block_count; i++) {
        LZ4Block currentBlock = {0};
        size_t current_block_size = (i == block_count - 1) ? n - i * block_size : block_size;
        if (par) parallel_block_encode_at(blocks[i], current_block_size, &currentBlock, &frame, i);
        else block_encode(blocks[i], current_block_size, &currentBlock, NULL, NULL, &frame); /* LZ4.c:721 */
    }
    size_t nseq = 0;
    for (size_t i = 0; i < frame.blocks; i++) nseq += frame.frame_blocks[i].sequences_count;
    FILE *out = fopen(argv[4], "wb");
    if (!out) {
        perror(argv[4]);
        return 2;
    }
    write_output(&frame, out); /* LZ4.c:733 */
    fclose(out);
    printf("%zu blocks, %zu sequences\n", block_count, nseq);
    for (size_t i = 0; i < block_count; i++) free(blocks[i]);
    free(blocks);
    free(input);
    return 0;
}
