/* include/ljb_compat.h — the reference's own C entry points, exported with their names and signatures from
 * compat/libljb_compat.so ("Reference C entry points to keep exported").
 *
 * A program written against CyrilMorel42/LZ4-JPEG's LZ4.c / JPEG.c keeps its calls; every function below does its compute
 * in liblz4jpeg_b200 (CUDA, sm_100a) and keeps the reference's types, file contract and error behaviour.  The struct
 * layouts repeat Algorithms/sequential/LZ4/LZ4.c:30-58 and Algorithms/parallel/JPEG/JPEG.c:29-60 because they ARE the
 * interface.  There is no CPU path behind these functions: without a usable B200 they print the error and exit(1), the
 * reference's own failure convention (LZ4.c:113-118).
 */
#ifndef LJB_COMPAT_H
#define LJB_COMPAT_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- LZ4 (Algorithms/sequential/LZ4/LZ4.c) ------------------------------------------------------ */
typedef struct { /* LZ4.c:30-38 */
    uint8_t token;
    size_t byte_size;
    uint8_t *literals;
    size_t literals_count;
    uint16_t match_offset;
    size_t match_length;
} LZ4Sequence;

typedef struct { /* LZ4.c:40-46 */
    uint8_t token;
    size_t byte_size;
    size_t sequences_count;
    LZ4Sequence *sequences;
} LZ4Block;

typedef struct { /* LZ4.c:48-52 */
    size_t blocks;
    LZ4Block *frame_blocks;
} LZ4Frame;

typedef struct { /* LZ4.c:54-58 */
    uint8_t *input_data;
    size_t input_size;
} LZ4Context;

/* LZ4.c:670 — reads ../Output-Input/input/input.txt, appends the frame to ../Output-Input/out/compressed.bin, writes the
 * hex dump ../Output-Input/out/compressed.txt.  Block length: ljb_compat_set_block_length (default 300, LZ4.c:23). */
void lz4_encode(void);
/* Algorithms/parallel/LZ4/LZ4.c:680 — the same frame (the GPU path is always block-parallel). */
void parallel_LZ4_encode(void);
/* LZ4.c:1038 / Algorithms/parallel/LZ4/LZ4.c:1105 — decodes input_bin_file to ../Output-Input/out/uncompressed.txt
 * (printable bytes as they are, the others as 0xNN, LZ4.c:1021-1032). */
void LZ4_decode(char *input_bin_file, char *log);
void parallel_LZ4_decode(char *input_bin_file, char *log);
/* LZ4.c:506 — encodes one block into *block (sequences malloc'd, literals pointing into block_entry, LZ4.c:525) and appends
 * it to *frame (LZ4.c:619).  log_file and output_file are unused, as in the reference. */
void block_encode(const char *block_entry, size_t block_length, LZ4Block *block, FILE *log_file, FILE *output_file, LZ4Frame *frame);
/* Algorithms/parallel/LZ4/LZ4.c:518 — the thread body of the parallel build; same result, stored at frame_blocks[index]. */
void parallel_block_encode_at(const char *block_entry, size_t block_length, LZ4Block *block, LZ4Frame *frame, size_t index);
/* LZ4.c:290 — longest earlier match of input[current_index..] (earliest among the longest, cap 1024), (uint8_t) of the
 * length, 0 below 4.  The reference passes no block extent (it reads past the block); the extent searched
 * here is the block bound with ljb_compat_bind_block, else ljb_compat block length bytes from `input`. */
uint8_t find_longest_match(uint8_t *input, size_t current_index, uint16_t *match_distance);
/* LZ4.c:427 / :415 / :365 — serialise a frame, a block, a sequence. */
void write_output(LZ4Frame *frame, FILE *output_file);
void write_block(LZ4Block *block, FILE *output_file);
void write_sequence(LZ4Sequence sequence, FILE *file);
/* LZ4.c:461, :443, :347 */
void add_block_to_frame(LZ4Frame *frame, LZ4Block block);
void add_sequence_to_block(LZ4Sequence seq, LZ4Block *block);
void free_frame(LZ4Frame *frame);
/* LZ4.c:123 — ceil(n / block_size) exact-size copies. */
char **divide_input(const uint8_t *input_data, size_t input_size, size_t block_size, size_t *block_count);

/* knobs that are compile-time #defines in the reference */
void ljb_compat_set_block_length(size_t block_length); /* DEFAULT_BLOCK_LENGTH, LZ4.c:23 */
size_t ljb_compat_block_length(void);
void ljb_compat_bind_block(const uint8_t *block, size_t block_length);
void ljb_compat_set_device(int device);

/* ---- JPEG (Algorithms/parallel/JPEG/JPEG.c) ------------------------------------------------------ */
typedef struct { /* JPEG.c:29-32 */
    unsigned char r, g, b, a;
} Pixel;

typedef struct { /* JPEG.c:42-55 */
    uint8_t lum_values[64];
    uint8_t b_values[32];
    uint8_t r_values[32];
    size_t index;
    double *lum_coefficients;
    double *r_coefficients;
    double *b_coefficients;
    int *RLE_encoded_lum;
    int *RLE_encoded_r;
    int *RLE_encoded_b;
} PixelGroup;

typedef struct { /* JPEG.c:57-60 */
    PixelGroup block;
} parallel_args;

/* Algorithms/parallel/JPEG/JPEG.c:1103 — the fused per-group pipeline, DWORD WINAPI process(LPVOID) there: forward chain
 * (DCT, Quantize, zig-zag, RLE, Huffman) and inverse chain in place.  On return, as in the reference, the three
 * *_coefficients arrays (malloc'd, JPEG.c:453) hold the dequantised coefficients and the *_values arrays the reconstructed
 * samples.  ljb_process_groups does the same for n groups in one launch. */
void *process(void *lpParam);
int ljb_process_groups(PixelGroup *groups, size_t n);

#ifdef __cplusplus
}
#endif
#endif
