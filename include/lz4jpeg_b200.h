/*
 * include/lz4jpeg_b200.h — C ABI of the B200-native LZ4 / JPEG block codecs.
 *
 * Drop-in boundary for the two per-block hot paths of CyrilMorel42/LZ4-JPEG.  The reference has no
 * library boundary of its own: each codec is one C file with a main(), and the timing harnesses
 * (Experiment/LZ4_*_experiment.c, JPEG_*_experiment.c) popen() an executable.  The functions below are the buffer-level
 * operations those programs are made of; each cites the reference interface it replaces.  Host code
 * stays plain C: include this header, link liblz4jpeg_b200.so (INTEGRATION.md shows the exact edits).
 *
 * Conventions
 *   - every function returns LJB_OK (0) or a negative LJB_E_* code and never calls exit() (the reference
 *     perror()+exit(1)s: Algorithms/sequential/LZ4/LZ4.c:113-118, :632-637);
 *   - pointers named d_* are CUDA device pointers, all others are host pointers;
 *   - the library is CUDA-only: there is no CPU fallback; without a usable device every entry point
 *     returns LJB_E_CUDA.
 */
#ifndef LZ4JPEG_B200_H
#define LZ4JPEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LJB_OK 0
#define LJB_E_ARG (-1)         /* invalid argument (NULL, zero size, block_len > 65536, odd image width ...) */
#define LJB_E_CUDA (-2)        /* CUDA runtime error or no device; ljb_last_cuda_error() has the text */
#define LJB_E_CAPACITY (-3)    /* output buffer too small; *out_len holds a lower bound of the required size (the bytes produced up to
                                  and including the chunk that did not fit); ljb_*_bound() is always enough */
#define LJB_E_FORMAT (-4)      /* decoder: stream is inconsistent / ambiguous (SURVEY.md A.3-b phantom sequences) */
#define LJB_E_UNSUPPORTED (-5) /* input outside the reference's defined behaviour (e.g. code longer than 31 bits) */

#define LJB_LZ4_MAX_BLOCK 65536u /* uint16 offsets / WINDOW_SIZE 65535 (LZ4.c:22) bound a block to 64 KiB */
#define LJB_LZ4_REF_BLOCK 300u   /* DEFAULT_BLOCK_LENGTH (LZ4.c:23) */

typedef struct ljb_ctx ljb_ctx; /* one per (process, GPU): device id, stream, persistent scratch */

/* Context: replaces the implicit process-global state of the reference programs. */
int ljb_ctx_create(int device, ljb_ctx **out);
void ljb_ctx_destroy(ljb_ctx *ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
void *ljb_ctx_stream(ljb_ctx *ctx);
const char *ljb_strerror(int code);
const char *ljb_last_cuda_error(void);
/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
uint64_t ljb_ctx_launch_count(const ljb_ctx *ctx);
/* Duration in ms of the most recent ljb_*_dev call's dominant kernel (CUDA events on the ctx stream); after a host-buffer
 * call (which launches one kernel per chunk): the sum over its chunks. */
float ljb_ctx_last_kernel_ms(const ljb_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * LZ4 (reference dialect)
 *
 * ljb_lz4_compress replaces lz4_encode()'s compute: divide_input -> block_encode per block ->
 * write_output (Algorithms/sequential/LZ4/LZ4.c:123, :506, :427; thread-per-block form
 * parallel_LZ4_encode, Algorithms/parallel/LZ4/LZ4.c:680).  Output bytes are identical to the
 * reference's compressed.bin for the same input and block_len (match extension bounded at the block
 * end):  u8 nblocks_lo8 | { u8 nseq_lo8, u16 size, sequences }*.
 *
 *   in, n          input bytes (n >= 1)
 *   block_len      1..65536; the reference's compile-time DEFAULT_BLOCK_LENGTH
 *   out, out_cap   receives the stream; ljb_lz4_bound(n, block_len) is always enough
 *   block_offsets  optional, nblocks+1 entries: byte offset of every block in out, and the end.  The
 *                  in-band size fields are 8/16-bit and wrap at this block size, so this table is the
 *                  authoritative framing.
 *   out_len        receives the stream length
 *   phantom        optional: number of sequences whose size field counts a byte that is not written
 *                  (match lengths 257..259 mod 256) — such streams are not decodable,
 *                  by the reference or by anyone
 * ------------------------------------------------------------------------------------------------ */
size_t ljb_lz4_bound(size_t n, size_t block_len);
size_t ljb_lz4_block_count(size_t n, size_t block_len);

int ljb_lz4_compress(ljb_ctx *ctx, const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                     uint64_t *block_offsets, size_t *out_len, uint64_t *phantom);

/* Same operation on device-resident buffers, asynchronous on the context stream.
 *   d_block_offsets  nblocks+1 device entries (required)
 *   d_result         3 device uint64: [0] stream length, [1] phantom count, [2] error flags (0 = ok,
 *                    bit0 = out_cap exceeded)
 *   first_block / frame_blocks: this call encodes a shard of a larger frame — blocks are numbered from
 *   first_block and the leading frame byte (low 8 bits of frame_blocks, LZ4.c:429) is written only when
 *   first_block == 0.  Single-GPU callers pass 0 and the total block count. */
int ljb_lz4_compress_dev(ljb_ctx *ctx, const uint8_t *d_in, size_t n, size_t block_len, uint8_t *d_out, size_t out_cap,
                         uint64_t *d_block_offsets, uint64_t *d_result, size_t first_block, size_t frame_blocks);

/* Per-position longest match of ONE block (n <= 65536): len[p] in {0,4..1024}, dist[p] = p - earliest
 * position.  Exposes the kernel stage that replaces find_longest_match (LZ4.c:290-323) for parity tests. */
int ljb_lz4_block_matches(ljb_ctx *ctx, const uint8_t *in, size_t n, uint16_t *len, uint16_t *dist);

/* Format-level decoder (replaces LZ4_decode's compute, LZ4.c:1038; block-parallel like
 * parallel_LZ4_decode, Algorithms/parallel/LZ4/LZ4.c:1105).  Needs the out-of-band block_offsets. */
int ljb_lz4_decompress(ljb_ctx *ctx, const uint8_t *comp, size_t comp_len, const uint64_t *block_offsets,
                       size_t nblocks, size_t block_len, uint8_t *out, size_t out_cap, size_t *out_len);

/* The same on device-resident buffers, asynchronous on the context stream.
 *   d_block_out_len  nblocks device uint32: decoded length of every block (scratch the caller provides)
 *   d_result         3 device uint64: [0] decoded bytes, [2] error flags (bit0 = out_cap exceeded,
 *                    bit1 = malformed stream or offset table; a short block that is not the last one is malformed) */
int ljb_lz4_decompress_dev(ljb_ctx *ctx, const uint8_t *d_comp, size_t comp_len, const uint64_t *d_block_offsets,
                           size_t nblocks, size_t block_len, uint8_t *d_out, size_t out_cap, uint32_t *d_block_out_len,
                           uint64_t *d_result);

/* ------------------------------------------------------------------------------------------------
 * JPEG-like encoder
 *
 * ljb_jpeg_encode_rgba replaces the encode half of the reference's main()/process():
 * build_{luminance,rChrominance,bChrominance}_matrix -> chroma_subsample -> divide_image ->
 * discrete_cosine_transform -> Quantize -> zigzag_pattern -> RLE -> encode_huffman ->
 * generate_encoded_sequence (Algorithms/sequential/JPEG/JPEG.c:114-185, :302, :496, :451, :621, :693,
 * :767, :1035, :993; fused per-block form Algorithms/parallel/JPEG/JPEG.c:1103-1252).
 *
 *   rgba, w, h, stride  8-bit RGBA pixels as the reference's Pixel (JPEG.c:29-32); w must be even
 *   first_group, ngroups  which 8x8 groups to encode (row-major over ceil(w/8) columns); the reference
 *                       processes groups [0, ceil(w*h/64)) (JPEG.c:1131) = ljb_jpeg_group_count(w,h).
 *                       Shards of one image pass sub-ranges; rgba always points at the whole image.
 *   out, out_cap        packed code bits: per group the luma, Cr, Cb strings (reference order,
 *                       JPEG.c:1242-1321) concatenated MSB-first and zero-padded to a byte boundary
 *   group_offsets       optional, ngroups+1 byte offsets into out
 *   group_bits          optional, 3 uint16 per group: bit lengths of the three strings
 *   coefs               optional, 128 int16 per group: quantised coefficients before zig-zag,
 *                       lum[64] (u*8+v), Cr[32], Cb[32] (u*4+v)
 *   out_len             receives the stream length in bytes
 * Returns LJB_E_UNSUPPORTED if some code is longer than 31 bits or a string is longer than the
 * reference's fixed buffers allow (1023 / 511 chars): the reference itself overflows there.
 * ------------------------------------------------------------------------------------------------ */
size_t ljb_jpeg_group_count(int w, int h);
size_t ljb_jpeg_bound(size_t ngroups);

int ljb_jpeg_encode_rgba(ljb_ctx *ctx, const uint8_t *rgba, int w, int h, size_t stride, size_t first_group,
                         size_t ngroups, uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits,
                         int16_t *coefs, size_t *out_len);

/* Device-resident form, asynchronous on the context stream.  d_result: 3 device uint64:
 * [0] stream length, [1] groups outside the reference's defined behaviour, [2] error flags. */
int ljb_jpeg_encode_rgba_dev(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t first_group,
                             size_t ngroups, uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets,
                             uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result);

/* The same two calls for pixels of THREE bytes (r g b), the layout stbi_load(path, &w, &h, &n, 3) returns and the one the reference's
 * Pixel rows (JPEG.c:29-40, three uint8_t) have: stride >= 3 * w.  A quarter less to upload for the same result, byte for byte. */
int ljb_jpeg_encode_rgb(ljb_ctx *ctx, const uint8_t *rgb, int w, int h, size_t stride, size_t first_group, size_t ngroups, uint8_t *out,
                        size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, int16_t *coefs, size_t *out_len);
int ljb_jpeg_encode_rgb_dev(ljb_ctx *ctx, const uint8_t *d_rgb, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                            uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                            uint64_t *d_result);

/* Batch of nimages equal-sized images (BASELINE.json configs[4]: 8192 x 1920x1080): ONE launch, one ticket counter, one
 * look-back for the whole batch — the per-image loop of the reference's harness (Experiment/JPEG_sequential_experiment.c:57-144)
 * costs a launch, a memset and a synchronisation per image, which dominates a 50 us frame.
 *   image i lies at rgba + i * image_stride and holds groups [i*G, (i+1)*G), G = ljb_jpeg_group_count(w, h);
 *   group_offsets has nimages*G + 1 entries, group_bits 3 per group; image i's stream is
 *   out[group_offsets[i*G], group_offsets[(i+1)*G]) — byte for byte what ljb_jpeg_encode_rgba gives for that image alone. */
int ljb_jpeg_encode_batch(ljb_ctx *ctx, const uint8_t *rgba, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                          uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len);
int ljb_jpeg_encode_batch_dev(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                              uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                              uint64_t *d_result);
/* The batch calls for r g b pixels of three bytes (stride >= 3 * w). */
int ljb_jpeg_encode_batch_rgb(ljb_ctx *ctx, const uint8_t *rgb, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                              uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len);
int ljb_jpeg_encode_batch_rgb_dev(ljb_ctx *ctx, const uint8_t *d_rgb, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                                  uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                                  uint64_t *d_result);

/* process() of the reference's parallel build (Algorithms/parallel/JPEG/JPEG.c:1103-1252) on groups given by their samples
 * (PixelGroup, JPEG.c:42-46: lum_values[64], b_values[32], r_values[32] = 128 bytes per group):
 *   ljb_jpeg_encode_groups_dev   the forward chain from the DCT on (device buffers; outputs as ljb_jpeg_encode_rgba_dev)
 *   ljb_jpeg_decode_groups_dev   Inverse_quantize + inverse DCT of given coefficients -> samples in the same 128-byte layout
 *   ljb_jpeg_process_groups      both, host buffers: samples in -> reconstructed samples out, plus the quantised coefficients */
int ljb_jpeg_encode_groups_dev(ljb_ctx *ctx, const uint8_t *d_samples, size_t ngroups, uint8_t *d_out, size_t out_cap,
                               uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result);
int ljb_jpeg_decode_groups_dev(ljb_ctx *ctx, const int16_t *d_coefs, size_t ngroups, uint8_t *d_samples, uint64_t *d_result);
int ljb_jpeg_process_groups(ljb_ctx *ctx, uint8_t *samples, size_t ngroups, int16_t *coefs);

/* The entropy half of the inverse chain: decode_huffman (JPEG.c:1009-1033) -> inverse_RLE (JPEG.c:811-842) ->
 * reverse_zigzag_pattern (JPEG.c:729-764).  The reference decodes while the Huffman tree is still in memory and never
 * serialises it; here the tree of every (group, channel) is serialised at LJB_JPEG_TREE_BYTES per group:
 *     luma at +0, Cr at +512, Cb at +768, each:  u8 k | i16 value[k] (leaves, first-appearance order, node ids 0..k-1)
 *                                                | { u8 left, u8 right }[k-1] (internal nodes k..2k-2 in creation order; root last)
 *   ljb_jpeg_trees           quantised coefficients (as ljb_jpeg_encode_rgba returns them) -> trees, by the reference's
 *                            calculate_frequency / build_heap / build_huffman_tree (JPEG.c:864-961)
 *   ljb_jpeg_entropy_decode  packed stream + group_offsets (relative to `stream`) + group_bits + trees -> coefficients;
 *                            LJB_E_FORMAT if an offset, a tree or a bit string is malformed
 * The _dev forms work on device-resident buffers, asynchronously on the context stream; d_result[2] bit1 = malformed;
 * offs_base is subtracted from every offset (the offsets of a shard are stream-global). */
#define LJB_JPEG_TREE_BYTES 1024
int ljb_jpeg_trees(ljb_ctx *ctx, const int16_t *coefs, size_t ngroups, uint8_t *trees);
int ljb_jpeg_trees_dev(ljb_ctx *ctx, const int16_t *d_coefs, size_t ngroups, uint8_t *d_trees);
int ljb_jpeg_entropy_decode(ljb_ctx *ctx, const uint8_t *stream, size_t stream_len, const uint64_t *group_offsets,
                            const uint16_t *group_bits, const uint8_t *trees, size_t ngroups, int16_t *coefs);
int ljb_jpeg_entropy_decode_dev(ljb_ctx *ctx, const uint8_t *d_stream, size_t stream_len, const uint64_t *d_group_offsets,
                                const uint16_t *d_group_bits, const uint8_t *d_trees, size_t ngroups, uint64_t offs_base,
                                int16_t *d_coefs, uint64_t *d_result);

/* Decode half of the reference's main() (JPEG.c:1408-1428): Inverse_quantize (JPEG.c:631) ->
 * inverse_discrete_cosine_transform (JPEG.c:399) -> assemble_image (JPEG.c:552, YCbCr -> RGB), bit-exact.
 *   coefs       128 int16 per group for all ljb_jpeg_group_count(w, h) groups, as ljb_jpeg_encode_rgba returns
 *               them or as ljb_jpeg_entropy_decode recovers them from the bit stream
 *   orig_rgba   the original image, or NULL.  When w or h is not a multiple of 8 the reference leaves the last
 *               tiled groups unprocessed (JPEG.c:1131) and its reconstructed.png shows their
 *               colour-converted original samples; NULL is an error (LJB_E_ARG) for such sizes.
 *   out_rgba    receives h rows of w RGBA pixels (a = 255) */
int ljb_jpeg_decode_coefs(ljb_ctx *ctx, const int16_t *coefs, int w, int h, const uint8_t *orig_rgba, size_t orig_stride,
                          uint8_t *out_rgba, size_t out_stride);
/* Device-resident form, asynchronous on the context stream.  d_result: 3 device uint64, [2] = error flags
 * (bit0: unprocessed groups exist and d_orig_rgba is NULL). */
int ljb_jpeg_decode_coefs_dev(ljb_ctx *ctx, const int16_t *d_coefs, int w, int h, const uint8_t *d_orig_rgba,
                              size_t orig_stride, uint8_t *d_out_rgba, size_t out_stride, uint64_t *d_result);

/* ------------------------------------------------------------------------------------------------
 * Baseline JPEG / JFIF
 *
 * ljb_jfif_encode replaces stbi_write_jpg_to_func / stbi_write_jpg_core of the stb_image_write.h the reference
 * vendors (Algorithms/sequential/JPEG/stb_image_write.h:1607, :1398-1605; the reference programs never call it):
 * a complete .jpg file — SOI, JFIF APP0, DQT, SOF0, DHT, SOS, entropy-coded segment with 0xFF stuffing, EOI —
 * byte-identical to what stb writes for the same pixels and quality.
 *
 *   pixels, w, h, comp, stride  8-bit samples, comp = 1 (grey), 2 (grey + alpha), 3 (RGB) or 4 (RGBA), as stb's
 *                       `comp`; stride is the row pitch in bytes (stb: w * comp)
 *   quality             1..100, 0 = 90 (stb_image_write.h:1479)
 *   subsample           -1 = stb's rule: 4:2:0 (16x16 MCUs) when quality <= 90, else 4:4:4 (:1480);
 *                       0 / 1 force 4:4:4 / 4:2:0 (BASELINE.json words its workload as "quality 75, 4:4:4")
 *   out, out_cap        receives the file; ljb_jfif_bound(w, h) is always enough, far less usually is
 *   out_len             receives the file length (on LJB_E_CAPACITY: a lower bound of the size needed)
 * ------------------------------------------------------------------------------------------------ */
size_t ljb_jfif_bound(int w, int h);

int ljb_jfif_encode(ljb_ctx *ctx, const uint8_t *pixels, int w, int h, int comp, size_t stride, int quality, int subsample,
                    uint8_t *out, size_t out_cap, size_t *out_len);

/* Device-resident form, asynchronous on the context stream.  d_result: 3 device uint64: [0] file length,
 * [1] unused, [2] error flags (bit0 = out_cap exceeded, bit1 = internal scratch exceeded — both mean a larger
 * out_cap is needed).  d_coefs (optional, for tests): 64 int16 per data unit in stream order, zig-zag order. */
int ljb_jfif_encode_dev(ljb_ctx *ctx, const uint8_t *d_pixels, int w, int h, int comp, size_t stride, int quality,
                        int subsample, uint8_t *d_out, size_t out_cap, uint64_t *d_result, int16_t *d_coefs);

/* ------------------------------------------------------------------------------------------------
 * Workload generators (host code): the reference's Experiment/random_extract.c:8-71 and
 * Experiment/random_image.c:58-77 with an explicit seed (the reference uses time() / unseeded rand()).
 * ------------------------------------------------------------------------------------------------ */
void ljb_synth_text(const uint8_t *corpus, size_t corpus_len, uint64_t seed, size_t passage, uint8_t *out, size_t n);
void ljb_synth_image(uint64_t seed, int w, int h, uint8_t *rgba);

#ifdef __cplusplus
}
#endif
#endif /* LZ4JPEG_B200_H */
