/*
 * oracle/jpeg_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's "JPEG-like" per-8x8-group encoder
 * (/root/reference/Algorithms/sequential/JPEG/JPEG.c, cited below as S-JPG:line; the fused per-block
 * form is `process` in Algorithms/parallel/JPEG/JPEG.c:1103-1252).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's shared object.  The product path never links or calls it.
 *
 * Parity status: PINNED against outputs of the reference itself run in this container: the
 * reference's own stage functions compiled from /root/reference into oracle/_ref/libref_jpeg.so
 * (oracle/build.py; gcc x86-64 SSE2, -O2 -ffp-contract=off, glibc libm) are compared stage by stage
 * with this file in tests/test_oracle_jpeg.py, and the committed fixtures under tests/golden/jpeg_*
 * were generated from that build by tests/golden/make_golden.py.  The reference repository holds no
 * known-answer vector for DCT/quantise/RLE/Huffman; its colour-conversion PNGs
 * (Output-Input/Images/{b,r}Chrominance.png) are reproduced exactly.
 *
 * All floating point is IEEE double, evaluated left to right with no contraction, like the
 * reference's gcc -O2 x86-64 build (compile this file with -ffp-contract=off).
 *
 * Bit-string packing (new; the reference keeps '0'/'1' C strings in memory and never serialises them,
 * S-JPG:1248-1249): per 8x8 group the luma, Cr and Cb code strings are concatenated in the
 * reference's emission order lum, r, b (S-JPG:1242-1321) MSB-first into bytes; the group record is
 * zero-padded to a byte boundary.  group_bits[3*g+{0,1,2}] hold the three string lengths.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define ORJ_PI 3.14159265358979323846 /* S-JPG:11 */

/* S-JPG:12-20 */
static const unsigned orj_qlum[64] = {8,  6,  6,  8,  10, 14, 18, 22, 6,  6,  7,  9,  12, 20, 22, 20, 6,  7,  8,  10, 14, 22,
                                      25, 22, 8,  9,  10, 14, 18, 28, 27, 22, 10, 12, 14, 18, 22, 35, 33, 26, 14, 18, 22, 22,
                                      27, 33, 36, 30, 18, 22, 26, 28, 33, 40, 40, 34, 22, 26, 28, 30, 36, 34, 35, 33};
/* S-JPG:22-27; consumed as 8 rows (u) x 4 cols (v), SURVEY.md B.5 */
static const unsigned orj_qchr[32] = {17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99,
                                      66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* libm cos through a volatile pointer so the compiler cannot fold it at build time with MPFR
 * (the reference evaluates cos() at run time inside the MAC loop, S-JPG:481-482). */
static double (*volatile orj_cos)(double) = cos;
static double (*volatile orj_sqrt)(double) = sqrt;

/* ---- S-JPG:114-185 colour conversion ----------------------------------------------------------- */
static inline uint8_t orj_clamp(int v) { return v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v); } /* S-JPG:132-139 */
static inline uint8_t orj_luma(unsigned r, unsigned g, unsigned b)
{
    double y = 0.299 * r + 0.587 * g + 0.114 * b; /* S-JPG:127, implicit double->uint8 truncation */
    return (uint8_t)y;
}
static inline uint8_t orj_cr(unsigned r, unsigned g, unsigned b)
{
    int v = (int)(0.439 * r - 0.368 * g - 0.071 * b + 128); /* S-JPG:157 */
    return orj_clamp(v);
}
static inline uint8_t orj_cb(unsigned r, unsigned g, unsigned b)
{
    int v = (int)(-0.148 * r - 0.291 * g + 0.439 * b + 128); /* S-JPG:180 */
    return orj_clamp(v);
}

void oracle_jpeg_planes(const uint8_t *rgba, int w, int h, size_t stride, uint8_t *Y, uint8_t *Cr, uint8_t *Cb)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t *p = rgba + (size_t)y * stride + 4 * (size_t)x;
            Y[(size_t)y * w + x] = orj_luma(p[0], p[1], p[2]);
            Cr[(size_t)y * w + x] = orj_cr(p[0], p[1], p[2]);
            Cb[(size_t)y * w + x] = orj_cb(p[0], p[1], p[2]);
        }
}

/* ---- S-JPG:302-375 chroma_subsample + S-JPG:496-550 divide_image -------------------------------
 * Group g = (row/8)*ceil(w/8) + col/8.  Luma 8x8 row-major; chroma 8 rows x 4 cols: for an even local
 * column c the sample is the subsampled plane at [row][col/2], i.e. the ORIGINAL chroma at column
 * col+1 (the subsample keeps odd columns).  Pixels outside the image stay 0.  Width must be even:
 * for odd widths the reference reads one element past its w/2-wide row (S-JPG:543 with :314).     */
static void orj_gather_group(const uint8_t *rgba, int w, int h, size_t stride, size_t g, uint8_t lum[64],
                             uint8_t r[32], uint8_t b[32])
{
    size_t bpr = ((size_t)w + 7) / 8;
    size_t brow = g / bpr, bcol = g % bpr;
    memset(lum, 0, 64);
    memset(r, 0, 32);
    memset(b, 0, 32);
    for (int lr = 0; lr < 8; ++lr) {
        size_t row = brow * 8 + lr;
        if (row >= (size_t)h) break;
        for (int lc = 0; lc < 8; ++lc) {
            size_t col = bcol * 8 + lc;
            if (col >= (size_t)w) break;
            const uint8_t *p = rgba + row * stride + 4 * col;
            lum[lr * 8 + lc] = orj_luma(p[0], p[1], p[2]);
            if ((lc & 1) == 0) {
                /* subsampled[row][col/2] == original chroma at column 2*(col/2)+1 == col+1 */
                const uint8_t *q = rgba + row * stride + 4 * (col + 1);
                r[lr * 4 + lc / 2] = orj_cr(q[0], q[1], q[2]);
                b[lr * 4 + lc / 2] = orj_cb(q[0], q[1], q[2]);
            }
        }
    }
}

/* ---- S-JPG:451-494 discrete_cosine_transform --------------------------------------------------- */
static void orj_dct(const uint8_t *data, size_t width, size_t height, double *coef)
{
    int corr[64];
    for (size_t i = 0; i < width * height; ++i) corr[i] = (int)data[i] - 128;
    for (size_t u = 0; u < height; ++u)
        for (size_t v = 0; v < width; ++v) {
            double sum = 0.0;
            for (size_t x = 0; x < height; ++x)
                for (size_t y = 0; y < width; ++y) {
                    double cos_x = orj_cos((ORJ_PI * (2 * x + 1) * u) / (2.0 * height));
                    double cos_y = orj_cos((ORJ_PI * (2 * y + 1) * v) / (2.0 * width));
                    sum += corr[x * width + y] * cos_x * cos_y;
                }
            double alpha_u = (u == 0) ? orj_sqrt(1.0 / height) : orj_sqrt(2.0 / height);
            double alpha_v = (v == 0) ? orj_sqrt(1.0 / width) : orj_sqrt(2.0 / width);
            coef[u * width + v] = alpha_u * alpha_v * sum;
        }
}

/* ---- S-JPG:621-629 Quantize: divide, truncate toward zero -------------------------------------- */
static void orj_quantize(const double *coef, const unsigned *table, size_t n, int *q)
{
    for (size_t i = 0; i < n; ++i) {
        double c = coef[i] / (double)table[i];
        q[i] = (int)c;
    }
}

/* ---- S-JPG:693-727 zigzag_pattern (generic width x height anti-diagonal walk) ------------------- */
static void orj_zigzag(size_t width, size_t height, const int *in, int *out)
{
    size_t index = 0;
    for (size_t sum = 0; sum < width + height - 1; ++sum) {
        size_t start_row = (sum < width) ? 0 : sum - width + 1;
        size_t end_row = (sum < height) ? sum : height - 1;
        if (sum % 2 == 0) {
            for (size_t row = end_row; row >= start_row && row < height; --row) { /* row-- wraps at 0 */
                size_t col = sum - row;
                if (col < width) out[index++] = in[row * width + col];
            }
        } else {
            for (size_t row = start_row; row <= end_row; ++row) {
                size_t col = sum - row;
                if (col < width) out[index++] = in[row * width + col];
            }
        }
    }
}

/* ---- S-JPG:767-809 RLE over all values: (count, value) pairs ----------------------------------- */
static size_t orj_rle(const int *in, size_t n, int *out)
{
    size_t o = 0, count = 1;
    int cur = in[0];
    for (size_t i = 1; i <= n; ++i) {
        if (i < n && in[i] == cur) {
            ++count;
        } else {
            out[o++] = (int)count;
            out[o++] = cur;
            if (i < n) {
                cur = in[i];
                count = 1;
            }
        }
    }
    return o;
}

/* ---- S-JPG:864-1097 per-(group,channel) adaptive Huffman code ----------------------------------
 * Replicates the reference procedure, including the array "heap" whose insert never sifts up
 * (S-JPG:956-958), because the emitted codes depend on it.                                        */
typedef struct {
    int count;
    int value; /* symbol + 1000, or -1 for an internal node (S-JPG:950) */
    int left, right; /* indices into the node pool, -1 for leaves */
} orj_node;

static void orj_heapify(orj_node *heap, size_t size, size_t i) /* S-JPG:894-911 */
{
    for (;;) {
        size_t smallest = i, l = 2 * i + 1, r = 2 * i + 2;
        if (l < size && heap[l].count < heap[smallest].count) smallest = l;
        if (r < size && heap[r].count < heap[smallest].count) smallest = r;
        if (smallest == i) return;
        orj_node t = heap[i];
        heap[i] = heap[smallest];
        heap[smallest] = t;
        i = smallest;
    }
}

typedef struct {
    int nsym;
    int value[128];       /* symbol value (without the +1000 bias), in DFS (codes[]) order */
    unsigned char len[128];
    unsigned char bits[128][128]; /* code as '0'/'1' values; the reference's char code[32] overflows at >= 32 */
} orj_codebook;

/* pool: leaves/internals popped from the heap are copied into a pool (the reference mallocs copies,
 * S-JPG:951-954) so child links stay valid while the heap array is shuffled. */
static void orj_assign(const orj_node *pool, const orj_node *node, unsigned char *code, int depth, orj_codebook *cb)
{ /* S-JPG:963-982 */
    if (node->value != -1) {
        cb->value[cb->nsym] = node->value - 1000;
        cb->len[cb->nsym] = (unsigned char)depth;
        memcpy(cb->bits[cb->nsym], code, (size_t)depth);
        cb->nsym++;
        return;
    }
    code[depth] = 0;
    orj_assign(pool, &pool[node->left], code, depth + 1, cb);
    code[depth] = 1;
    orj_assign(pool, &pool[node->right], code, depth + 1, cb);
}

static void orj_huffman(const int *in, size_t n, orj_codebook *cb)
{
    /* S-JPG:864-885 calculate_frequency: first-appearance order, linear search */
    orj_node heap[128];
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        int v = in[i] + 1000, found = 0;
        for (size_t j = 0; j < k; ++j)
            if (heap[j].value == v) {
                heap[j].count++;
                found = 1;
                break;
            }
        if (!found) {
            heap[k].value = v;
            heap[k].count = 1;
            heap[k].left = heap[k].right = -1;
            ++k;
        }
    }
    /* S-JPG:913-934 build_heap */
    for (int i = (int)(k / 2) - 1; i >= 0; --i) orj_heapify(heap, k, (size_t)i);
    /* S-JPG:936-961 build_huffman_tree */
    orj_node pool[256];
    int np = 0;
    size_t size = k;
    while (size > 1) {
        orj_node left = heap[0];
        heap[0] = heap[--size];
        orj_heapify(heap, size, 0);
        orj_node right = heap[0];
        heap[0] = heap[--size];
        orj_heapify(heap, size, 0);
        pool[np] = left;
        pool[np + 1] = right;
        orj_node nn;
        nn.count = left.count + right.count;
        nn.value = -1;
        nn.left = np;
        nn.right = np + 1;
        np += 2;
        heap[size] = nn;
        ++size;
        orj_heapify(heap, size, size - 1); /* no-op: a leaf index has no children (S-JPG:958) */
    }
    cb->nsym = 0;
    unsigned char code[130];
    orj_assign(pool, &heap[0], code, 0, cb);
}

/* S-JPG:993-1007 generate_encoded_sequence, appended to a bit writer (MSB first) */
typedef struct {
    uint8_t *buf;
    size_t cap;
    size_t bitpos;
    int overflow;
} orj_bw;
static void orj_put(orj_bw *w, unsigned bit)
{
    size_t byte = w->bitpos >> 3;
    if (byte >= w->cap) {
        w->overflow = 1;
        return;
    }
    if ((w->bitpos & 7) == 0) w->buf[byte] = 0;
    if (bit) w->buf[byte] |= (uint8_t)(0x80u >> (w->bitpos & 7));
    w->bitpos++;
}
static size_t orj_emit(const int *in, size_t n, const orj_codebook *cb, orj_bw *w, int *maxlen)
{
    size_t bits = 0;
    for (size_t i = 0; i < n; ++i)
        for (int j = 0; j < cb->nsym; ++j)
            if (cb->value[j] == in[i]) {
                for (int t = 0; t < cb->len[j]; ++t) orj_put(w, cb->bits[j][t]);
                bits += cb->len[j];
                if (cb->len[j] > *maxlen) *maxlen = cb->len[j];
                break;
            }
    return bits;
}

/* One channel of one group: DCT -> quantise -> zig-zag -> RLE -> Huffman -> bits. */
static size_t orj_channel(const uint8_t *samples, size_t width, size_t height, const unsigned *qt, int16_t *qout,
                          orj_bw *w, int *maxlen, int *rle_out, size_t *rle_len)
{
    double coef[64];
    int q[64], zz[64], rle[128];
    size_t n = width * height;
    orj_dct(samples, width, height, coef);
    orj_quantize(coef, qt, n, q);
    if (qout)
        for (size_t i = 0; i < n; ++i) qout[i] = (int16_t)q[i];
    orj_zigzag(width, height, q, zz);
    size_t m = orj_rle(zz, n, rle);
    if (rle_out) {
        memcpy(rle_out, rle, m * sizeof(int));
        *rle_len = m;
    }
    orj_codebook *cb = malloc(sizeof(orj_codebook));
    orj_huffman(rle, m, cb);
    size_t bits = orj_emit(rle, m, cb, w, maxlen);
    free(cb);
    return bits;
}

/* Number of groups the reference processes: ceil(w*h/64) (S-JPG:1131), which is < the number of
 * tiled groups when a dimension is not a multiple of 8. */
size_t oracle_jpeg_group_count(int w, int h)
{
    return ((size_t)w * (size_t)h + 63) / 64;
}

/*
 * Encode groups [g0, g1) of an RGBA image.
 *   coefs       : optional, 128 int16 per group: lum[64] row-major (u*8+v), r[32], b[32] (u*4+v), quantised
 *   out         : packed bit stream of the groups, group records byte aligned, in order
 *   group_off   : optional, (g1-g0)+1 byte offsets into out
 *   group_bits  : optional, 3 uint16 per group: bit lengths of the lum, r, b strings
 *   max_code_len: optional, longest code emitted (the reference's char code[32] holds <= 31)
 * Returns 0; -1 bad args / odd width; -2 output overflow.
 */
int oracle_jpeg_encode(const uint8_t *rgba, int w, int h, size_t stride, size_t g0, size_t g1, int16_t *coefs,
                       uint8_t *out, size_t out_cap, uint64_t *group_off, uint16_t *group_bits, size_t *out_len,
                       int *max_code_len)
{
    if (w <= 0 || h <= 0 || (w & 1)) return -1;
    if (g1 > oracle_jpeg_group_count(w, h) || g0 > g1) return -1;
    size_t o = 0;
    int maxlen = 0;
    for (size_t g = g0; g < g1; ++g) {
        uint8_t lum[64], r[32], b[32];
        orj_gather_group(rgba, w, h, stride, g, lum, r, b);
        orj_bw bw = {out + o, out_cap - o, 0, 0};
        int16_t *c = coefs ? coefs + 128 * (g - g0) : NULL;
        size_t bl = orj_channel(lum, 8, 8, orj_qlum, c, &bw, &maxlen, NULL, NULL);
        size_t br = orj_channel(r, 4, 8, orj_qchr, c ? c + 64 : NULL, &bw, &maxlen, NULL, NULL);
        size_t bb = orj_channel(b, 4, 8, orj_qchr, c ? c + 96 : NULL, &bw, &maxlen, NULL, NULL);
        if (bw.overflow) return -2;
        if (group_off) group_off[g - g0] = o;
        if (group_bits) {
            group_bits[3 * (g - g0) + 0] = (uint16_t)bl;
            group_bits[3 * (g - g0) + 1] = (uint16_t)br;
            group_bits[3 * (g - g0) + 2] = (uint16_t)bb;
        }
        o += (bw.bitpos + 7) / 8;
    }
    if (group_off) group_off[g1 - g0] = o;
    if (out_len) *out_len = o;
    if (max_code_len) *max_code_len = maxlen;
    return 0;
}

/* Stage-level access for tests: one group's samples, unquantised coefficients, RLE arrays. */
int oracle_jpeg_group_stages(const uint8_t *rgba, int w, int h, size_t stride, size_t g, uint8_t *samples /*128*/,
                             double *coef /*128*/, int *rle /*3*128*/, size_t *rle_len /*3*/)
{
    if (w <= 0 || h <= 0 || (w & 1)) return -1;
    uint8_t *lum = samples, *r = samples + 64, *b = samples + 96;
    orj_gather_group(rgba, w, h, stride, g, lum, r, b);
    orj_dct(lum, 8, 8, coef);
    orj_dct(r, 4, 8, coef + 64);
    orj_dct(b, 4, 8, coef + 96);
    uint8_t scratch[4096];
    int maxlen = 0;
    orj_bw bw = {scratch, sizeof scratch, 0, 0};
    orj_channel(lum, 8, 8, orj_qlum, NULL, &bw, &maxlen, rle, &rle_len[0]);
    orj_channel(r, 4, 8, orj_qchr, NULL, &bw, &maxlen, rle + 128, &rle_len[1]);
    orj_channel(b, 4, 8, orj_qchr, NULL, &bw, &maxlen, rle + 256, &rle_len[2]);
    return 0;
}

/* ---- decode half: S-JPG:631-638 Inverse_quantize, S-JPG:399-448 inverse_discrete_cosine_transform ---------- */
static void orj_idct(const int16_t *q, const unsigned *table, size_t width, size_t height, uint8_t *values)
{
    double coef[64];
    for (size_t i = 0; i < width * height; ++i) {
        coef[i] = (double)q[i];   /* the reference keeps the quantised integers in doubles (S-JPG:627) */
        coef[i] *= table[i];      /* S-JPG:636 */
    }
    for (size_t x = 0; x < height; ++x)
        for (size_t y = 0; y < width; ++y) {
            double sum = 0.0;
            for (size_t u = 0; u < height; ++u)
                for (size_t v = 0; v < width; ++v) {
                    double alpha_u = (u == 0) ? orj_sqrt(1.0 / height) : orj_sqrt(2.0 / height);
                    double alpha_v = (v == 0) ? orj_sqrt(1.0 / width) : orj_sqrt(2.0 / width);
                    double cos_x = orj_cos((ORJ_PI * (2 * x + 1) * u) / (2.0 * height));
                    double cos_y = orj_cos((ORJ_PI * (2 * y + 1) * v) / (2.0 * width));
                    sum += alpha_u * alpha_v * coef[u * width + v] * cos_x * cos_y; /* S-JPG:430 */
                }
            int value = (int)round(sum + 128.0); /* S-JPG:441 */
            values[x * width + y] = (value < 0) ? 0 : (value > 255) ? 255 : (uint8_t)value;
        }
}

/* Decode: coefs (128 int16 per group, groups [0, ceil(w*h/64))) -> RGBA image, following the tail of the
 * reference's main() (S-JPG:1408-1428) and assemble_image (S-JPG:552-619).  Tiled groups beyond that count are
 * never processed by the reference (S-JPG:1131): they still hold divide_image's samples of the original, so the
 * original image is needed for them (orig may be NULL when w and h are multiples of 8).  Returns 0, -1 bad args. */
int oracle_jpeg_decode(const int16_t *coefs, int w, int h, const uint8_t *orig, size_t orig_stride, uint8_t *out, size_t out_stride)
{
    if (w <= 0 || h <= 0 || (w & 1) || !coefs || !out) return -1;
    size_t bpr = ((size_t)w + 7) / 8, bpc = ((size_t)h + 7) / 8;
    size_t total = oracle_jpeg_group_count(w, h);
    for (size_t g = 0; g < bpr * bpc; ++g) {
        uint8_t lum[64], r[32], b[32];
        if (g < total) {
            orj_idct(coefs + 128 * g, orj_qlum, 8, 8, lum);
            orj_idct(coefs + 128 * g + 96, orj_qchr, 4, 8, b); /* order lum, b, r (S-JPG:1418-1420): irrelevant */
            orj_idct(coefs + 128 * g + 64, orj_qchr, 4, 8, r);
        } else {
            if (!orig) return -1;
            orj_gather_group(orig, w, h, orig_stride, g, lum, r, b);
        }
        size_t brow = g / bpr, bcol = g % bpr;
        for (size_t lr = 0; lr < 8; ++lr)
            for (size_t lc = 0; lc < 8; ++lc) {
                size_t row = brow * 8 + lr, col = bcol * 8 + lc;
                if (row >= (size_t)h || col >= (size_t)w) continue;
                uint8_t Y = lum[lr * 8 + lc];
                uint8_t Cb = b[lr * 4 + lc / 2], Cr = r[lr * 4 + lc / 2];
                int R = (int)Y + (int)(1.402 * (Cr - 128)); /* S-JPG:598-600 */
                int G = (int)Y - (int)(0.344136 * (Cb - 128)) - (int)(0.714136 * (Cr - 128));
                int B = (int)Y + (int)(1.772 * (Cb - 128));
                R = R < 0 ? 0 : (R > 255 ? 255 : R);
                G = G < 0 ? 0 : (G > 255 ? 255 : G);
                B = B < 0 ? 0 : (B > 255 ? 255 : B);
                uint8_t *p = out + row * out_stride + 4 * col;
                p[0] = (uint8_t)R;
                p[1] = (uint8_t)G;
                p[2] = (uint8_t)B;
                p[3] = 255;
            }
    }
    return 0;
}

/* The DCT basis values the reference's cos()/sqrt() calls produce on this libm, for checking the
 * constants embedded in the CUDA source: cos8[x*8+u], cos4[y*4+v], alpha8[2], alpha4[2]. */
void oracle_jpeg_basis(double *cos8, double *cos4, double *alpha8, double *alpha4)
{
    for (size_t x = 0; x < 8; ++x)
        for (size_t u = 0; u < 8; ++u) cos8[x * 8 + u] = orj_cos((ORJ_PI * (2 * x + 1) * u) / (2.0 * 8));
    for (size_t y = 0; y < 4; ++y)
        for (size_t v = 0; v < 4; ++v) cos4[y * 4 + v] = orj_cos((ORJ_PI * (2 * y + 1) * v) / (2.0 * 4));
    alpha8[0] = orj_sqrt(1.0 / 8);
    alpha8[1] = orj_sqrt(2.0 / 8);
    alpha4[0] = orj_sqrt(1.0 / 4);
    alpha4[1] = orj_sqrt(2.0 / 4);
}

/* random_image-style noise (Experiment/random_image.c:58-77): r,g,b iid uniform bytes, a = 255.
 * The reference calls rand() unseeded; a splitmix64 stream with an explicit seed is used instead.
 * (Duplicated on purpose in lz4-jpeg_b200/csrc/synth.c: the product must not link the oracle.)   */
void oracle_synth_image(uint64_t seed, int w, int h, uint8_t *rgba)
{
    uint64_t s = seed;
    size_t npx = (size_t)w * h;
    for (size_t i = 0; i < npx; i += 2) {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        /* 6 random bytes -> two pixels */
        rgba[4 * i + 0] = (uint8_t)(z);
        rgba[4 * i + 1] = (uint8_t)(z >> 8);
        rgba[4 * i + 2] = (uint8_t)(z >> 16);
        rgba[4 * i + 3] = 255;
        if (i + 1 < npx) {
            rgba[4 * i + 4] = (uint8_t)(z >> 24);
            rgba[4 * i + 5] = (uint8_t)(z >> 32);
            rgba[4 * i + 6] = (uint8_t)(z >> 40);
            rgba[4 * i + 7] = 255;
        }
    }
}
