/* compat/test_comm.c — a C host driving N GPUs through include/ljb_comm.h (one process, NCCL all-gather of shard totals):
 * the N-GPU stream and tables must equal the single-GPU ones byte for byte.
 *   test_comm <ngpus> <text-file> <block_len> <w> <h>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ljb_comm.h"

static void die(const char *what, int rc)
{
    fprintf(stderr, "%s: %s (%s)\n", what, ljb_strerror(rc), ljb_last_cuda_error());
    exit(1);
}

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    const int ngpus = atoi(argv[1]);
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *in = malloc(n);
    if (fread(in, 1, n, f) != n) return 2;
    fclose(f);
    const size_t bl = (size_t)atol(argv[3]);
    const int w = atoi(argv[4]), h = atoi(argv[5]);
    ljb_comm *comm = NULL;
    int rc = ljb_comm_create(ngpus, &comm);
    if (rc != LJB_OK) die("ljb_comm_create", rc);
    /* LZ4 */
    const size_t nb = ljb_lz4_block_count(n, bl), cap = ljb_lz4_bound(n, bl);
    uint8_t *a = malloc(cap), *b = malloc(cap);
    uint64_t *ao = malloc((nb + 1) * 8), *bo = malloc((nb + 1) * 8), pa = 0, pb = 0;
    size_t la = 0, lb = 0;
    if ((rc = ljb_comm_lz4_compress(comm, in, n, bl, a, cap, ao, &la, &pa)) != LJB_OK) die("ljb_comm_lz4_compress", rc);
    if ((rc = ljb_lz4_compress(ljb_comm_ctx(comm, 0), in, n, bl, b, cap, bo, &lb, &pb)) != LJB_OK) die("ljb_lz4_compress", rc);
    int bad = la != lb || pa != pb || memcmp(a, b, la) != 0 || memcmp(ao, bo, (nb + 1) * 8) != 0;
    printf("lz4: %d GPUs %zu -> %zu bytes, %zu blocks, %s\n", ngpus, n, la, nb, bad ? "MISMATCH" : "identical to 1 GPU");
    /* JPEG: a deterministic w x h RGBA image */
    uint8_t *img = malloc((size_t)w * h * 4);
    ljb_synth_image(42, w, h, img);
    const size_t ng = ljb_jpeg_group_count(w, h), jcap = ljb_jpeg_bound(ng);
    uint8_t *ja = malloc(jcap), *jb = malloc(jcap);
    uint64_t *jao = malloc((ng + 1) * 8), *jbo = malloc((ng + 1) * 8);
    uint16_t *jab = malloc(ng * 6), *jbb = malloc(ng * 6);
    size_t jla = 0, jlb = 0;
    if ((rc = ljb_comm_jpeg_encode_rgba(comm, img, w, h, (size_t)w * 4, ja, jcap, jao, jab, &jla)) != LJB_OK) die("ljb_comm_jpeg_encode_rgba", rc);
    if ((rc = ljb_jpeg_encode_rgba(ljb_comm_ctx(comm, 0), img, w, h, (size_t)w * 4, 0, ng, jb, jcap, jbo, jbb, NULL, &jlb)) != LJB_OK)
        die("ljb_jpeg_encode_rgba", rc);
    int jbad = jla != jlb || memcmp(ja, jb, jla) != 0 || memcmp(jao, jbo, (ng + 1) * 8) != 0 || memcmp(jab, jbb, ng * 6) != 0;
    printf("jpeg: %d GPUs %dx%d -> %zu bytes, %zu groups, %s\n", ngpus, w, h, jla, ng, jbad ? "MISMATCH" : "identical to 1 GPU");
    /* the same image as r g b, three bytes per pixel: the same stream */
    uint8_t *rgb = malloc((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; ++i) memcpy(rgb + 3 * i, img + 4 * i, 3);
    size_t jlc = 0;
    if ((rc = ljb_comm_jpeg_encode_rgb(comm, rgb, w, h, (size_t)w * 3, ja, jcap, jao, jab, &jlc)) != LJB_OK) die("ljb_comm_jpeg_encode_rgb", rc);
    int rbad = jlc != jlb || memcmp(ja, jb, jlc) != 0 || memcmp(jao, jbo, (ng + 1) * 8) != 0 || memcmp(jab, jbb, ng * 6) != 0;
    printf("jpeg rgb: %d GPUs -> %zu bytes, %s\n", ngpus, jlc, rbad ? "MISMATCH" : "identical to 1 GPU");
    ljb_comm_destroy(comm);
    return bad || jbad || rbad;
}
