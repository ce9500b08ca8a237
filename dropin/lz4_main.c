/*
 * dropin/lz4_main.c — drop-in for the reference's LZ4_seq.exe / LZ4_par.exe.
 *
 * Same contract as the reference programs (Algorithms/sequential/LZ4/LZ4.c:1125-1136,
 * Algorithms/parallel/LZ4/LZ4.c:1227-1250), which the unmodified timing harnesses spawn with popen()
 * (Experiment/LZ4_sequential_experiment.c:102, Experiment/LZ4_parallel_experiment.c:102): no arguments, fixed
 * relative paths (LZ4.c:24-28), nothing printed on success, exit status 0.
 *
 *   reads   ../Output-Input/input/input.txt
 *   writes  ../Output-Input/out/compressed.bin   (same bytes as the reference encoder)
 *           ../Output-Input/out/compressed.txt   (hex dump, "%02X " per byte, LZ4.c:75-107)
 *           ../Output-Input/out/uncompressed.txt (decoded text; non-printables as 0xNN, LZ4.c:1021-1032)
 *           ../Output-Input/log/encoding_log.txt (truncated, LZ4.c:204-213)
 *
 * The block loop (divide_input + block_encode + write_output, LZ4.c:704-733) and the decoder (LZ4.c:1038-1121) are
 * one call each into liblz4jpeg_b200; everything else here is file I/O.  Deliberately not replicated:
 * ensure_directories() exits when it manages to create a directory (LZ4.c:196-201) — the harness never relies on it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include "lz4jpeg_b200.h"

#ifndef DEFAULT_BLOCK_LENGTH
#define DEFAULT_BLOCK_LENGTH 300 /* LZ4.c:23 */
#endif
#define DEFAULT_LOG_FILE "../Output-Input/log/encoding_log.txt"
#define DEFAULT_COMPRESSED_FILE "../Output-Input/out/compressed.bin"
#define DEFAULT_UNCOMPRESSED_FILE "../Output-Input/out/uncompressed.txt"
#define DEFAULT_INPUT_FILE "../Output-Input/input/input.txt"
#define DEFAULT_HEX_FILE "../Output-Input/out/compressed.txt"

static FILE *safe_open(const char *file_name, const char *mode) /* LZ4.c:109-120 */
{
    FILE *file = fopen(file_name, mode);
    if (file == NULL) {
        perror("Error: Unable to open file");
        exit(1);
    }
    return file;
}

static void die(int rc, const char *what)
{
    fprintf(stderr, "%s: %s", what, ljb_strerror(rc));
    if (rc == LJB_E_CUDA) fprintf(stderr, " (%s)", ljb_last_cuda_error());
    fprintf(stderr, "\n");
    exit(1);
}

int main(void)
{
    size_t block_length = DEFAULT_BLOCK_LENGTH;
    const char *env = getenv("LJB_BLOCK_LENGTH"); /* the reference's knob is a compile-time #define */
    if (env && atol(env) > 0) block_length = (size_t)atol(env);
    if (block_length == 500) { /* LZ4.c:672-677 */
        printf("Error: block length cannot have the value 500");
        exit(1);
    }
    /* clear_files, LZ4.c:204-213 */
    fclose(safe_open(DEFAULT_COMPRESSED_FILE, "wb"));
    fclose(safe_open(DEFAULT_LOG_FILE, "w"));

    /* lz4_encode, LZ4.c:670-742 */
    FILE *log_file = safe_open(DEFAULT_LOG_FILE, "a");
    FILE *input_file = safe_open(DEFAULT_INPUT_FILE, "r");
    FILE *output_file = safe_open(DEFAULT_COMPRESSED_FILE, "ab");
    fseek(input_file, 0, SEEK_END);
    long file_size = ftell(input_file);
    fseek(input_file, 0, SEEK_SET);
    if (file_size < (long)block_length) { /* extract_uncompressed_file, LZ4.c:632-637 */
        printf("Error: default block length is too high, please reduce it before proceding.");
        exit(1);
    }
    uint8_t *input = malloc((size_t)file_size + 1);
    if (!input || fread(input, 1, (size_t)file_size, input_file) != (size_t)file_size) {
        perror("Error reading input file");
        exit(1);
    }
    ljb_ctx *ctx = NULL;
    int rc = ljb_ctx_create(0, &ctx);
    if (rc != LJB_OK) die(rc, "ljb_ctx_create");
    const size_t n = (size_t)file_size;
    const size_t nblocks = ljb_lz4_block_count(n, block_length);
    const size_t cap = ljb_lz4_bound(n, block_length);
    uint8_t *stream = malloc(cap);
    uint64_t *offsets = malloc((nblocks + 1) * sizeof *offsets);
    size_t out_len = 0;
    if (!stream || !offsets) {
        perror("malloc");
        exit(1);
    }
    rc = ljb_lz4_compress(ctx, input, n, block_length, stream, cap, offsets, &out_len, NULL);
    if (rc != LJB_OK) die(rc, "ljb_lz4_compress");
    fwrite(stream, 1, out_len, output_file); /* the bytes write_output() produces, LZ4.c:427-441 */
    fclose(log_file);
    fclose(input_file);
    fclose(output_file);
    { /* dump_to_hex_file, LZ4.c:75-107 */
        FILE *hex = fopen(DEFAULT_HEX_FILE, "w");
        if (hex == NULL) {
            perror("Error opening output file");
        } else {
            for (size_t i = 0; i < out_len; i++) fprintf(hex, "%02X ", stream[i]);
            fclose(hex);
        }
    }

    /* LZ4_decode + interpret_frame, LZ4.c:1038-1121, :984-1036 */
    uint8_t *decoded = malloc(nblocks * block_length + 1);
    size_t decoded_len = 0;
    if (!decoded) {
        perror("malloc");
        exit(1);
    }
    rc = ljb_lz4_decompress(ctx, stream, out_len, offsets, nblocks, block_length, decoded, nblocks * block_length, &decoded_len);
    if (rc != LJB_OK) die(rc, "ljb_lz4_decompress");
    FILE *uncompressed_file = safe_open(DEFAULT_UNCOMPRESSED_FILE, "wb");
    for (size_t i = 0; i < decoded_len; i++) {
        if (decoded[i] >= 32 && decoded[i] <= 126)
            fprintf(uncompressed_file, "%c", decoded[i]);
        else
            fprintf(uncompressed_file, "0x%02X", decoded[i]);
    }
    fclose(uncompressed_file);
    ljb_ctx_destroy(ctx);
    free(decoded);
    free(offsets);
    free(stream);
    free(input);
    return 0;
}
