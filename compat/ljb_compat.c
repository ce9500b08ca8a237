/* compat/ljb_compat.c — the reference's C entry points (names, signatures, file contract) over liblz4jpeg_b200.
 * See include/ljb_compat.h.  Host code only: every byte of codec work happens in the CUDA library; what is here is file I/O,
 * (de)serialisation of the reference's structs and the reference's error convention (message + exit(1), LZ4.c:113-118). */
#include "ljb_compat.h"

#include <stdlib.h>
#include <string.h>

#include "lz4jpeg_b200.h"

#define DEFAULT_LOG_FILE "../Output-Input/log/encoding_log.txt"           /* LZ4.c:24-28 */
#define DEFAULT_COMPRESSED_FILE "../Output-Input/out/compressed.bin"
#define DEFAULT_UNCOMPRESSED_FILE "../Output-Input/out/uncompressed.txt"
#define DEFAULT_INPUT_FILE "../Output-Input/input/input.txt"
#define DEFAULT_OUTPUT_HEX_FILE "../Output-Input/out/compressed.txt"

static size_t g_block_length = LJB_LZ4_REF_BLOCK; /* DEFAULT_BLOCK_LENGTH 300, LZ4.c:23 */
static int g_device = 0;
static ljb_ctx *g_ctx = NULL;
static const uint8_t *g_bound_block = NULL;
static size_t g_bound_len = 0;
static uint16_t *g_match_len = NULL, *g_match_dist = NULL; /* per-position matches of the bound block */
static int g_match_valid = 0;

void ljb_compat_set_block_length(size_t block_length) { g_block_length = block_length; }
size_t ljb_compat_block_length(void) { return g_block_length; }
void ljb_compat_set_device(int device) { g_device = device; }
void ljb_compat_bind_block(const uint8_t *block, size_t block_length)
{
    g_bound_block = block;
    g_bound_len = block_length;
    g_match_valid = 0;
}

static void die(int rc, const char *what)
{
    fprintf(stderr, "%s: %s", what, ljb_strerror(rc));
    if (rc == LJB_E_CUDA) fprintf(stderr, " (%s)", ljb_last_cuda_error());
    fprintf(stderr, "\n");
    exit(1);
}

static ljb_ctx *ctx(void)
{
    if (!g_ctx) {
        int rc = ljb_ctx_create(g_device, &g_ctx);
        if (rc != LJB_OK) die(rc, "ljb_ctx_create");
    }
    return g_ctx;
}

static FILE *safe_open_(const char *file_name, const char *mode) /* LZ4.c:109-120 */
{
    FILE *file = fopen(file_name, mode);
    if (file == NULL) {
        perror("Error: Unable to open file");
        exit(1);
    }
    return file;
}

/* ---- struct plumbing, LZ4.c:443-504, :347-363 ---------------------------------------------------- */
void add_sequence_to_block(LZ4Sequence seq, LZ4Block *block)
{
    if (block->sequences_count == 0) block->sequences = malloc(sizeof(LZ4Sequence));
    else block->sequences = realloc(block->sequences, sizeof(LZ4Sequence) * (block->sequences_count + 1));
    block->sequences[block->sequences_count] = seq;
    block->sequences_count += 1;
    block->byte_size += seq.byte_size;
}

void add_block_to_frame(LZ4Frame *frame, LZ4Block block)
{
    if (frame->blocks == 0) frame->frame_blocks = malloc(sizeof(LZ4Block));
    else frame->frame_blocks = realloc(frame->frame_blocks, sizeof(LZ4Block) * (frame->blocks + 1));
    if (frame->frame_blocks == NULL) {
        fprintf(stderr, "Error: Memory allocation failed for frame_blocks\n");
        exit(1);
    }
    frame->frame_blocks[frame->blocks] = block;
    frame->blocks++;
}

void free_frame(LZ4Frame *frame)
{
    if (!frame) return;
    if (frame->frame_blocks) {
        for (size_t i = 0; i < frame->blocks; i++) free(frame->frame_blocks[i].sequences);
        free(frame->frame_blocks);
    }
    frame->frame_blocks = NULL;
    frame->blocks = 0;
}

char **divide_input(const uint8_t *input_data, size_t input_size, size_t block_size, size_t *block_count) /* LZ4.c:123-177 */
{
    *block_count = ljb_lz4_block_count(input_size, block_size);
    char **blocks = malloc(*block_count * sizeof(char *));
    if (!blocks) {
        perror("malloc");
        exit(1);
    }
    for (size_t i = 0; i < *block_count; i++) {
        size_t len = (i == *block_count - 1) ? input_size - i * block_size : block_size;
        blocks[i] = malloc(len); /* exact size, no padding (LZ4.c:156) */
        if (!blocks[i]) {
            perror("malloc");
            exit(1);
        }
        memcpy(blocks[i], input_data + i * block_size, len);
    }
    return blocks;
}

/* ---- serialisation, LZ4.c:365-441 (the wire format) -------------------------------- */
void write_sequence(LZ4Sequence sequence, FILE *file)
{
    fwrite(&sequence.token, sizeof(uint8_t), 1, file);
    fwrite(&sequence.byte_size, sizeof(uint16_t), 1, file); /* low 16 bits (little-endian host, as the reference assumes) */
    if (sequence.literals_count >= 15) {
        uint8_t remaining = (uint8_t)(sequence.literals_count - 15); /* uint8_t, LZ4.c:374 */
        while (remaining >= 255) {
            uint8_t to_write = 255;
            fwrite(&to_write, sizeof(uint8_t), 1, file);
            remaining -= 255;
        }
        fwrite(&remaining, sizeof(uint8_t), 1, file);
    }
    fwrite(sequence.literals, sizeof(uint8_t), sequence.literals_count, file);
    fwrite(&sequence.match_offset, sizeof(uint16_t), 1, file);
    if (sequence.match_length >= 4) {
        uint8_t adjusted = (uint8_t)(sequence.match_length - 4);
        if (adjusted >= 15) {
            uint8_t remaining = (uint8_t)(adjusted - 15);
            while (remaining >= 255) {
                uint8_t to_write = 255;
                fwrite(&to_write, sizeof(uint8_t), 1, file);
                remaining -= 255;
            }
            fwrite(&remaining, sizeof(uint8_t), 1, file);
        }
    }
}

void write_block(LZ4Block *block, FILE *output_file)
{
    fwrite(&block->token, sizeof(uint8_t), 1, output_file);
    fwrite(&block->byte_size, sizeof(uint16_t), 1, output_file);
    for (size_t i = 0; i < block->sequences_count; i++) write_sequence(block->sequences[i], output_file);
}

void write_output(LZ4Frame *frame, FILE *output_file)
{
    fwrite(&frame->blocks, sizeof(uint8_t), 1, output_file);
    for (size_t i = 0; i < frame->blocks; i++) write_block(&frame->frame_blocks[i], output_file);
    free(frame->frame_blocks); /* as the reference does (LZ4.c:436-440): the sequences themselves leak there too */
    frame->frame_blocks = NULL;
    frame->blocks = 0;
}

/* ---- block_encode: one block through the GPU encoder, its bytes read back into the reference's structs ---- */
/* A sequence in the stream: token, u16 size, [literal-length bytes], literals, u16 offset, [match-length byte].
 * Sequences the format cannot represent (match lengths 257..259 mod 256) are written by the reference with a
 * size field one larger than their bytes and a token of 0xFD..0xFF; they are recognised by exactly that. */
static void parse_block(const uint8_t *s, size_t len, const uint8_t *block_entry, LZ4Block *block)
{
    size_t q = 3; /* block header */
    size_t in_pos = 0;
    block->sequences_count = 0;
    block->sequences = NULL;
    block->byte_size = 0;
    while (q + 5 <= len) {
        LZ4Sequence seq;
        memset(&seq, 0, sizeof seq);
        const uint8_t token = s[q];
        const size_t size16 = (size_t)s[q + 1] | ((size_t)s[q + 2] << 8);
        const size_t mtok = token & 15;
        size_t lit = token >> 4, next = 0, mext = mtok == 15 ? 1 : 0;
        int normal = 1;
        if (lit == 15) { /* literal-length byte(s) follow; the true count is recovered from the size field (it wraps mod 256) */
            next = s[q + 3] == 255 ? 2 : 1;
            size_t sz = size16;
            if (sz < 5 + next + mext + 15) sz += 65536; /* the u16 field wrapped (>= 65531 literals) */
            lit = sz - (5 + next + mext);
            normal = ((lit - 15) & 0xFF) == (next == 2 ? 255u : s[q + 3]) && q + 3 + next + lit + 2 + mext <= len;
        } else {
            normal = size16 == lit + 5 + mext;
        }
        if (!normal) { /* a match of 1..3 mod 256: token 0xFD..0xFF, one counted-but-unwritten match-length byte */
            mext = 1;
            next = 0;
            lit = size16 - 6;
            if (lit >= 15) {
                next = s[q + 3] == 255 ? 2 : 1;
                lit -= next;
            }
            seq.match_length = (size_t)token + 4 - 256; /* (uint8_t)(match_length - 4) == token's value 253..255 */
        }
        size_t p = q + 3 + next + lit;
        seq.token = token;
        seq.byte_size = lit + 5 + next + mext; /* as block_encode computes it (LZ4.c:546-575), before the u16 truncation */
        seq.literals = (uint8_t *)block_entry + in_pos;
        seq.literals_count = lit;
        seq.match_offset = (uint16_t)(s[p] | (s[p + 1] << 8));
        p += 2;
        if (normal) {
            if (seq.match_offset != 0) seq.match_length = mtok == 15 ? (size_t)s[p++] + 19 : mtok + 4;
            else seq.match_length = 0;
        }
        in_pos += lit + seq.match_length;
        add_sequence_to_block(seq, block);
        q = p;
    }
}

static void encode_block_into(const char *block_entry, size_t block_length, LZ4Block *block)
{
    const size_t cap = ljb_lz4_bound(block_length, block_length);
    uint8_t *stream = malloc(cap);
    uint64_t offs[2];
    size_t out_len = 0;
    if (!stream) {
        perror("malloc");
        exit(1);
    }
    int rc = ljb_lz4_compress(ctx(), (const uint8_t *)block_entry, block_length, block_length, stream, cap, offs, &out_len, NULL);
    if (rc != LJB_OK) die(rc, "ljb_lz4_compress");
    parse_block(stream + offs[0], (size_t)(offs[1] - offs[0]), (const uint8_t *)block_entry, block);
    block->token = (uint8_t)block->sequences_count; /* LZ4.c:615 */
    block->byte_size += 3;                          /* LZ4.c:617 */
    free(stream);
}

void block_encode(const char *block_entry, size_t block_length, LZ4Block *block, FILE *log_file, FILE *output_file, LZ4Frame *frame)
{
    (void)log_file;
    (void)output_file;
    encode_block_into(block_entry, block_length, block);
    add_block_to_frame(frame, *block); /* LZ4.c:619 */
}

void parallel_block_encode_at(const char *block_entry, size_t block_length, LZ4Block *block, LZ4Frame *frame, size_t index)
{
    encode_block_into(block_entry, block_length, block);
    frame->frame_blocks[index] = *block; /* parallel_add_block_to_frame, Algorithms/parallel/LZ4/LZ4.c:504-510 */
    frame->blocks++;
}

uint8_t find_longest_match(uint8_t *input, size_t current_index, uint16_t *match_distance)
{
    if (input != g_bound_block) ljb_compat_bind_block(input, g_block_length);
    if (current_index >= g_bound_len || g_bound_len > LJB_LZ4_MAX_BLOCK) return 0;
    if (!g_match_valid) {
        g_match_len = realloc(g_match_len, sizeof(uint16_t) * LJB_LZ4_MAX_BLOCK);
        g_match_dist = realloc(g_match_dist, sizeof(uint16_t) * LJB_LZ4_MAX_BLOCK);
        if (!g_match_len || !g_match_dist) {
            perror("malloc");
            exit(1);
        }
        int rc = ljb_lz4_block_matches(ctx(), g_bound_block, g_bound_len, g_match_len, g_match_dist);
        if (rc != LJB_OK) die(rc, "ljb_lz4_block_matches");
        g_match_valid = 1;
    }
    if (g_match_len[current_index] >= 4) { /* LZ4.c:314-321 */
        *match_distance = g_match_dist[current_index];
        return (uint8_t)g_match_len[current_index];
    }
    return 0;
}

/* ---- the file-level drivers, LZ4.c:670-742 and :1038-1121 ----------------------------------------- */
void lz4_encode(void)
{
    if (g_block_length == 500) { /* LZ4.c:672-677 */
        printf("Error: block length cannot have the value 500");
        exit(1);
    }
    FILE *log_file = safe_open_(DEFAULT_LOG_FILE, "a");
    FILE *input_file = safe_open_(DEFAULT_INPUT_FILE, "r");
    FILE *output_file = safe_open_(DEFAULT_COMPRESSED_FILE, "ab");
    fseek(input_file, 0, SEEK_END);
    long file_size = ftell(input_file);
    fseek(input_file, 0, SEEK_SET);
    if (file_size < (long)g_block_length) { /* extract_uncompressed_file, LZ4.c:632-637 */
        printf("Error: default block length is too high, please reduce it before proceding.");
        exit(1);
    }
    uint8_t *input = malloc((size_t)file_size + 1);
    if (!input || fread(input, 1, (size_t)file_size, input_file) != (size_t)file_size) {
        perror("Error reading input file");
        exit(1);
    }
    const size_t n = (size_t)file_size, cap = ljb_lz4_bound(n, g_block_length);
    uint8_t *stream = malloc(cap);
    size_t out_len = 0;
    if (!stream) {
        perror("malloc");
        exit(1);
    }
    int rc = ljb_lz4_compress(ctx(), input, n, g_block_length, stream, cap, NULL, &out_len, NULL);
    if (rc != LJB_OK) die(rc, "ljb_lz4_compress");
    fwrite(stream, 1, out_len, output_file); /* the bytes write_output() produces, LZ4.c:427-441 */
    fclose(log_file);
    fclose(input_file);
    fclose(output_file);
    FILE *hex = fopen(DEFAULT_OUTPUT_HEX_FILE, "w"); /* dump_to_hex_file, LZ4.c:75-107 */
    if (hex == NULL) {
        perror("Error opening output file");
    } else {
        for (size_t i = 0; i < out_len; i++) fprintf(hex, "%02X ", stream[i]);
        fclose(hex);
    }
    free(stream);
    free(input);
}

void parallel_LZ4_encode(void) { lz4_encode(); }

void LZ4_decode(char *input_bin_file, char *log)
{
    if (g_block_length == 500) { /* LZ4.c:1040-1045 */
        printf("Error: block length cannot have the value 500");
        exit(1);
    }
    FILE *input_file = safe_open_(input_bin_file, "rb");
    FILE *log_file = safe_open_(log, "a");
    fseek(input_file, 0, SEEK_END);
    long sz = ftell(input_file);
    fseek(input_file, 0, SEEK_SET);
    uint8_t *comp = malloc((size_t)sz + 8);
    if (!comp || sz < 1 || fread(comp, 1, (size_t)sz, input_file) != (size_t)sz) {
        perror("Error reading compressed file");
        exit(1);
    }
    fclose(input_file);
    fclose(log_file);
    /* The frame carries no offset table: the blocks are delimited by their 16-bit size fields (LZ4.c:419), which is exact as
     * long as no block exceeds 65535 bytes — true for the reference's own block length and anything near it. */
    size_t cap_blocks = 16, nblocks = 0;
    uint64_t *offs = malloc(sizeof(uint64_t) * (cap_blocks + 1));
    size_t q = 1;
    while (q + 3 <= (size_t)sz) {
        size_t bs = (size_t)comp[q + 1] | ((size_t)comp[q + 2] << 8);
        if (bs < 3 || q + bs > (size_t)sz) {
            fprintf(stderr, "Error: inconsistent block size in %s\n", input_bin_file);
            exit(1);
        }
        if (nblocks == cap_blocks) {
            cap_blocks *= 2;
            offs = realloc(offs, sizeof(uint64_t) * (cap_blocks + 1));
        }
        offs[nblocks++] = q;
        q += bs;
    }
    offs[nblocks] = q;
    if (nblocks == 0) {
        fprintf(stderr, "Error: no block in %s\n", input_bin_file);
        exit(1);
    }
    uint8_t *decoded = malloc(nblocks * g_block_length + 1);
    size_t decoded_len = 0;
    int rc = ljb_lz4_decompress(ctx(), comp, (size_t)sz, offs, nblocks, g_block_length, decoded, nblocks * g_block_length, &decoded_len);
    if (rc != LJB_OK) die(rc, "ljb_lz4_decompress");
    FILE *uncompressed_file = safe_open_(DEFAULT_UNCOMPRESSED_FILE, "wb"); /* interpret_frame, LZ4.c:1021-1032 */
    for (size_t i = 0; i < decoded_len; i++) {
        if (decoded[i] >= 32 && decoded[i] <= 126) fprintf(uncompressed_file, "%c", decoded[i]);
        else fprintf(uncompressed_file, "0x%02X", decoded[i]);
    }
    fclose(uncompressed_file);
    free(decoded);
    free(offs);
    free(comp);
}

void parallel_LZ4_decode(char *input_bin_file, char *log) { LZ4_decode(input_bin_file, log); }

/* ---- JPEG: process(), Algorithms/parallel/JPEG/JPEG.c:1103-1252 ------------------------------------ */
static const double kLumTable[64] = {8,  6,  6,  8,  10, 14, 18, 22, 6,  6,  7,  9,  12, 20, 22, 20, 6,  7,  8,  10, 14, 22,
                                     25, 22, 8,  9,  10, 14, 18, 28, 27, 22, 10, 12, 14, 18, 22, 35, 33, 26, 14, 18, 22, 22,
                                     27, 33, 36, 30, 18, 22, 26, 28, 33, 40, 40, 34, 22, 26, 28, 30, 36, 34, 35, 33}; /* JPEG.c:12-20 */
static const double kChrTable[32] = {17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99,
                                     66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}; /* JPEG.c:22-27 */

int ljb_process_groups(PixelGroup *groups, size_t n)
{
    if (!groups || n == 0) return LJB_E_ARG;
    uint8_t *samples = malloc(n * 128);
    int16_t *coefs = malloc(n * 128 * sizeof(int16_t));
    if (!samples || !coefs) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    for (size_t i = 0; i < n; i++) {
        memcpy(samples + i * 128, groups[i].lum_values, 64);
        memcpy(samples + i * 128 + 64, groups[i].b_values, 32);
        memcpy(samples + i * 128 + 96, groups[i].r_values, 32);
    }
    int rc = ljb_jpeg_process_groups(ctx(), samples, n, coefs);
    if (rc != LJB_OK) die(rc, "ljb_jpeg_process_groups");
    for (size_t i = 0; i < n; i++) {
        PixelGroup *g = &groups[i];
        memcpy(g->lum_values, samples + i * 128, 64);
        memcpy(g->b_values, samples + i * 128 + 64, 32);
        memcpy(g->r_values, samples + i * 128 + 96, 32);
        /* the reference leaves the DEQUANTISED coefficients in the arrays discrete_cosine_transform malloc'd (JPEG.c:453, Inverse_quantize
         * P-JPG:1243-1245); the RLE arrays are never read after process() and are left untouched */
        g->lum_coefficients = malloc(sizeof(double) * 64);
        g->r_coefficients = malloc(sizeof(double) * 32);
        g->b_coefficients = malloc(sizeof(double) * 32);
        const int16_t *c = coefs + i * 128;
        for (int k = 0; k < 64; k++) g->lum_coefficients[k] = (double)c[k] * kLumTable[k];
        for (int k = 0; k < 32; k++) {
            g->r_coefficients[k] = (double)c[64 + k] * kChrTable[k];
            g->b_coefficients[k] = (double)c[96 + k] * kChrTable[k];
        }
    }
    free(samples);
    free(coefs);
    return 0;
}

void *process(void *lpParam)
{
    parallel_args *args = (parallel_args *)lpParam;
    ljb_process_groups(&args->block, 1);
    return NULL;
}
