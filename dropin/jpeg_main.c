/*
 * dropin/jpeg_main.c — drop-in for the reference's JPEG_seq.exe / JPEG_par.exe.
 *
 * Same contract as the reference programs (Algorithms/sequential/JPEG/JPEG.c:1099-1460,
 * Algorithms/parallel/JPEG/JPEG.c:1254-1355), spawned without arguments by the unmodified harnesses
 * (Experiment/JPEG_sequential_experiment.c:99, Experiment/JPEG_parallel_experiment.c:99):
 *
 *   reads   ../Assets/Images/rand_8X8.png                                        (JPEG.c:9, :1102)
 *   writes  ../Output-Input/Images/{original,bChrominance,rChrominance,luminance,reconstructed}.png
 *                                                                                (JPEG.c:10, :1107, :1121-1123, :1428)
 *
 * The per-group encode chain (colour -> subsample -> tile -> DCT -> quantise -> zig-zag -> RLE -> Huffman) and the
 * decode chain (dequantise -> IDCT -> assemble) are one call each into liblz4jpeg_b200.  The three colour-plane
 * pictures are debugging output of the reference's driver, produced here on the host with the reference's own
 * expressions.  PNG decoding / encoding is the stb single-header library, as in the reference (third-party code,
 * compiled from where the reference vendors it: see dropin/build.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"

#include "lz4jpeg_b200.h"

#define IMAGES_DIRECTORY "../Assets/Images/"
#define OUTPUT_DIRECTORY "../Output-Input/Images/"

static void write_png(const char *name, int w, int h, const unsigned char *rgba) /* create_png_image, JPEG.c:187-214 */
{
    char path[512];
    snprintf(path, sizeof path, "%s%s", OUTPUT_DIRECTORY, name);
    if (stbi_write_png(path, w, h, 4, rgba, w * 4) == 0) printf("Error: Failed to write the PNG image.\n");
}

static unsigned char clamp(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); } /* JPEG.c:132-139 */

static void die(int rc, const char *what)
{
    fprintf(stderr, "%s: %s", what, ljb_strerror(rc));
    if (rc == LJB_E_CUDA) fprintf(stderr, " (%s)", ljb_last_cuda_error());
    fprintf(stderr, "\n");
    exit(EXIT_FAILURE);
}

int main(void)
{
    int w, h, channels;
    unsigned char *data = stbi_load(IMAGES_DIRECTORY "rand_8X8.png", &w, &h, &channels, 0); /* read_image, JPEG.c:66-103 */
    if (!data) {
        fprintf(stderr, "Error: Could not load image %s\n", IMAGES_DIRECTORY "rand_8X8.png");
        exit(EXIT_FAILURE);
    }
    const size_t npx = (size_t)w * (size_t)h;
    unsigned char *rgba = malloc(npx * 4), *plane = malloc(npx * 4);
    if (!rgba || !plane) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    for (size_t i = 0; i < npx; i++) { /* Pixel{r,g,b,a}: a = 255 unless the file has 4 channels (JPEG.c:85-96) */
        rgba[4 * i + 0] = data[channels * i + 0];
        rgba[4 * i + 1] = channels >= 3 ? data[channels * i + 1] : data[channels * i + 0];
        rgba[4 * i + 2] = channels >= 3 ? data[channels * i + 2] : data[channels * i + 0];
        rgba[4 * i + 3] = channels == 4 ? data[channels * i + 3] : 255;
    }
    stbi_image_free(data);
    write_png("original.png", w, h, rgba); /* JPEG.c:1107 */

    /* debugging pictures of the colour planes, JPEG.c:1110-1123 with :114-185 and :216-300 */
    for (size_t i = 0; i < npx; i++) { /* bChrominance.png */
        const unsigned char r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
        const unsigned char v = clamp((int)(-0.148 * r - 0.291 * g + 0.439 * b + 128));
        plane[4 * i + 0] = 128 + 1.402 * (128 - 128);
        plane[4 * i + 1] = 128 - 0.344 * (v - 128) - 0.714 * (128 - 128);
        plane[4 * i + 2] = 128 + 1.772 * (v - 128);
        plane[4 * i + 3] = 255;
    }
    write_png("bChrominance.png", w, h, plane);
    for (size_t i = 0; i < npx; i++) { /* rChrominance.png */
        const unsigned char r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
        const unsigned char v = clamp((int)(0.439 * r - 0.368 * g - 0.071 * b + 128));
        plane[4 * i + 0] = 128 + 1.402 * (v - 128);
        plane[4 * i + 1] = 128 - 0.344 * (128 - 128) - 0.714 * (v - 128);
        plane[4 * i + 2] = 128 + 1.772 * (128 - 128);
        plane[4 * i + 3] = 255;
    }
    write_png("rChrominance.png", w, h, plane);
    for (size_t i = 0; i < npx; i++) { /* luminance.png */
        const unsigned char r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
        const unsigned char v = (unsigned char)(0.299 * r + 0.587 * g + 0.114 * b);
        plane[4 * i + 0] = plane[4 * i + 1] = plane[4 * i + 2] = v;
        plane[4 * i + 3] = 255;
    }
    write_png("luminance.png", w, h, plane);

    /* encode (JPEG.c:1126-1249) and decode (JPEG.c:1253-1425) on the GPU */
    if (w & 1) { /* the reference reads past its subsampled rows for odd widths (JPEG.c:543 with :314) */
        fprintf(stderr, "Error: odd image widths are outside the reference's defined behaviour\n");
        exit(EXIT_FAILURE);
    }
    ljb_ctx *ctx = NULL;
    int rc = ljb_ctx_create(0, &ctx);
    if (rc != LJB_OK) die(rc, "ljb_ctx_create");
    const size_t ng = ljb_jpeg_group_count(w, h);
    const size_t cap = ljb_jpeg_bound(ng);
    uint8_t *bits = malloc(cap);
    int16_t *coefs = malloc(ng * 128 * sizeof *coefs);
    size_t out_len = 0;
    if (!bits || !coefs) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    rc = ljb_jpeg_encode_rgba(ctx, rgba, w, h, (size_t)w * 4, 0, ng, bits, cap, NULL, NULL, coefs, &out_len);
    if (rc != LJB_OK) die(rc, "ljb_jpeg_encode_rgba");
    rc = ljb_jpeg_decode_coefs(ctx, coefs, w, h, rgba, (size_t)w * 4, plane, (size_t)w * 4);
    if (rc != LJB_OK) die(rc, "ljb_jpeg_decode_coefs");
    write_png("reconstructed.png", w, h, plane); /* JPEG.c:1428 */
    ljb_ctx_destroy(ctx);
    free(coefs);
    free(bits);
    free(plane);
    free(rgba);
    return 0;
}
