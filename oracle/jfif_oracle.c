/*
 * oracle/jfif_oracle.c — TEST INFRASTRUCTURE (never linked into or called by the product library).
 *
 * CPU restatement of the only true baseline-JPEG encoder under /root/reference: the vendored, never-called
 * stbi_write_jpg_core (Algorithms/sequential/JPEG/stb_image_write.h v1.16, :1398-1605; helpers :1250-1396).
 * SURVEY.md §8f rank 4 ("true baseline JFIF mode").  Parity is PINNED: tests/test_oracle_jfif.py compares this file
 * byte for byte with that header compiled from where it lies (oracle/ref_glue_jfif.c -> oracle/_ref/libref_jfif.so).
 *
 * What the encoder is (each step cites the stb line it restates):
 *   tables      Annex-K luminance/chrominance quantisers scaled by the libjpeg quality rule, clamped to 1..255,
 *               stored in zig-zag order (:1472-1487); float divisors 1/(q * aan[row] * aan[col]) (:1489-1494)
 *   header      SOI, JFIF APP0, DQT x2, SOF0 (3 components, 2x2 luma sampling when subsampled), DHT x4 (Annex K
 *               code lengths + values), SOS (:1497-1519) — 607 bytes
 *   pixels      float YCbCr, Y level-shifted by -128 (:1541-1543, :1572-1574); edge MCUs repeat the last row/column
 *   subsample   quality <= 90 -> 16x16 MCUs, chroma = mean of 2x2 (:1480, :1553-1561); else 8x8 MCUs, 4:4:4
 *   DCT         AAN 8-point float flow graph on rows then columns (:1270-1316, :1336-1343)
 *   quantise    v = coef * divisor; (int)(v < 0 ? v - 0.5f : v + 0.5f), written in zig-zag order (:1345-1355)
 *   entropy     DC difference category + AC (run,size) symbols with the Annex-K Huffman codes, ZRL for runs >= 16,
 *               EOB unless the last coefficient is non-zero (:1357-1395); one continuous bit stream, 0xFF followed by
 *               a stuffed 0x00 (:1250-1268); 7 one-bits of padding, whole bytes only, then EOI (:1586-1591)
 *
 * `force_subsample` (-1 = stb's rule, 0 = 4:4:4, 1 = 4:2:0) is this repo's one extension: BASELINE.json words the
 * workload as "quality 75, 4:4:4", which stb never produces (4:4:4 needs quality > 90); the reference build for that
 * case is the same header with the `subsample = quality <= 90` line made overridable on a temporary copy.
 *
 * The Huffman code words are derived here from the Annex-K (BITS, HUFFVAL) lists by the canonical procedure of
 * ITU-T T.81 Annex C instead of being stored as tables; the parity test shows they are the ones stb hard-codes.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* ---- ITU-T T.81 Annex K tables ------------------------------------------------------------------------------- */
static const uint8_t K1_LUMA_Q[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                      14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                      18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                      49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t K2_CHROMA_Q[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                        99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                        99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
/* BITS[1..16] and HUFFVAL of Tables K.3 - K.6 */
static const uint8_t K3_DC_LUMA_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t K4_DC_CHROMA_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t K_DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t K5_AC_LUMA_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t K5_AC_LUMA_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t K6_AC_CHROMA_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t K6_AC_CHROMA_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

/* natural (row-major) index -> position in the zig-zag scan (T.81 Figure 5) */
static void zigzag_rank(uint8_t rank[64])
{
    int r = 0, c = 0;
    for (int k = 0; k < 64; ++k) {
        rank[r * 8 + c] = (uint8_t)k;
        if (((r + c) & 1) == 0) { /* moving up-right */
            if (c == 7) ++r;
            else if (r == 0) ++c;
            else { --r; ++c; }
        } else { /* moving down-left */
            if (r == 7) ++c;
            else if (c == 0) ++r;
            else { ++r; --c; }
        }
    }
}

typedef struct {
    uint16_t code[256];
    uint8_t len[256];
} huff_t;

/* T.81 Annex C: codes of one length are consecutive integers; moving to the next length doubles the counter. */
static void canonical_codes(const uint8_t bits[16], const uint8_t *vals, huff_t *h)
{
    memset(h, 0, sizeof *h);
    unsigned code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k) {
            h->code[vals[k]] = (uint16_t)code++;
            h->len[vals[k]] = (uint8_t)len;
        }
        code <<= 1;
    }
}

typedef struct {
    uint8_t qy[64], quv[64];   /* quantisers in zig-zag order, as the DQT segment carries them */
    float dy[64], duv[64];     /* per-coefficient multipliers in natural order */
    huff_t dc_y, ac_y, dc_c, ac_c;
    uint8_t rank[64];
} jfif_tables;

static void make_tables(int quality, jfif_tables *t)
{
    /* stb :1479-1482 — quality 0 means 90; scale = 5000/q below 50, 200-2q from 50 up */
    quality = quality ? quality : 90;
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    zigzag_rank(t->rank);
    for (int i = 0; i < 64; ++i) {
        int y = (K1_LUMA_Q[i] * scale + 50) / 100, c = (K2_CHROMA_Q[i] * scale + 50) / 100;
        t->qy[t->rank[i]] = (uint8_t)(y < 1 ? 1 : y > 255 ? 255 : y);
        t->quv[t->rank[i]] = (uint8_t)(c < 1 ? 1 : c > 255 ? 255 : c);
    }
    /* AAN scale factors times 2*sqrt(2); every product is rounded to float, as the constant folding of stb's
     * `1.387039845f * 2.828427125f` is (:1466-1467) */
    static const float aan[8] = {1.0f, 1.387039845f, 1.306562965f, 1.175875602f, 1.0f, 0.785694958f, 0.541196100f, 0.275899379f};
    float s[8];
    for (int i = 0; i < 8; ++i) s[i] = aan[i] * 2.828427125f;
    for (int r = 0, k = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c, ++k) {
            t->dy[k] = 1 / (t->qy[t->rank[k]] * s[r] * s[c]);
            t->duv[k] = 1 / (t->quv[t->rank[k]] * s[r] * s[c]);
        }
    canonical_codes(K3_DC_LUMA_BITS, K_DC_VALS, &t->dc_y);
    canonical_codes(K5_AC_LUMA_BITS, K5_AC_LUMA_VALS, &t->ac_y);
    canonical_codes(K4_DC_CHROMA_BITS, K_DC_VALS, &t->dc_c);
    canonical_codes(K6_AC_CHROMA_BITS, K6_AC_CHROMA_VALS, &t->ac_c);
}

/* ---- byte sink + bit writer ---------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *out;
    size_t cap, n; /* n keeps counting past cap so that the caller learns the needed size */
    uint64_t acc;  /* pending bits, right-aligned */
    int nacc;
} sink_t;

static void put_byte(sink_t *s, unsigned b)
{
    if (s->n < s->cap) s->out[s->n] = (uint8_t)b;
    s->n++;
}
static void put_bytes(sink_t *s, const void *p, size_t n)
{
    for (size_t i = 0; i < n; ++i) put_byte(s, ((const uint8_t *)p)[i]);
}
static void put_bits(sink_t *s, unsigned value, int nbits)
{
    s->acc = (s->acc << nbits) | (value & ((1u << nbits) - 1u));
    s->nacc += nbits;
    while (s->nacc >= 8) {
        unsigned b = (unsigned)(s->acc >> (s->nacc - 8)) & 0xffu;
        put_byte(s, b);
        if (b == 0xff) put_byte(s, 0); /* stb :1258-1262 */
        s->nacc -= 8;
    }
}

/* ---- transform ------------------------------------------------------------------------------------------------ */
/* One 8-point AAN pass over v[0], v[st], ..., v[7*st], in place; every operation is a separately rounded float
 * operation in the order of stb :1270-1316 (the build uses -ffp-contract=off). */
static void aan_pass(float *v, int st)
{
    const float a0 = v[0], a1 = v[st], a2 = v[2 * st], a3 = v[3 * st], a4 = v[4 * st], a5 = v[5 * st], a6 = v[6 * st], a7 = v[7 * st];
    const float s07 = a0 + a7, d07 = a0 - a7, s16 = a1 + a6, d16 = a1 - a6;
    const float s25 = a2 + a5, d25 = a2 - a5, s34 = a3 + a4, d34 = a3 - a4;
    /* even half */
    const float e0 = s07 + s34, e3 = s07 - s34, e1 = s16 + s25, e2 = s16 - s25;
    v[0] = e0 + e1;
    v[4 * st] = e0 - e1;
    const float r = (e2 + e3) * 0.707106781f;
    v[2 * st] = e3 + r;
    v[6 * st] = e3 - r;
    /* odd half */
    const float o0 = d34 + d25, o1 = d25 + d16, o2 = d16 + d07;
    const float z5 = (o0 - o2) * 0.382683433f;
    const float z2 = o0 * 0.541196100f + z5;
    const float z4 = o2 * 1.306562965f + z5;
    const float z3 = o1 * 0.707106781f;
    const float p = d07 + z3, m = d07 - z3;
    v[5 * st] = m + z2;
    v[3 * st] = m - z2;
    v[st] = p + z4;
    v[7 * st] = p - z4;
}

/* 8x8 samples at s[y*st + x] -> 64 quantised coefficients in zig-zag order */
static void transform_unit(float *s, int st, const float *mult, const uint8_t rank[64], int q[64])
{
    for (int y = 0; y < 8; ++y) aan_pass(s + y * st, 1);
    for (int x = 0; x < 8; ++x) aan_pass(s + x, st);
    for (int y = 0, k = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x, ++k) {
            const float v = s[y * st + x] * mult[k];
            q[rank[k]] = (int)(v < 0 ? v - 0.5f : v + 0.5f);
        }
}

/* magnitude category and the "additional bits" of T.81 F.1.2.1 (stb :1318-1326) */
static int category(int v, unsigned *extra)
{
    int a = v < 0 ? -v : v, n = 0;
    while (a) { ++n; a >>= 1; }
    if (n == 0) n = 1; /* stb's loop starts at 1; only reached with v == 0 by callers that never pass 0 */
    *extra = (unsigned)(v < 0 ? v - 1 : v) & ((1u << n) - 1u);
    return n;
}

static int entropy_unit(sink_t *s, const int q[64], int pred, const huff_t *dc, const huff_t *ac)
{
    unsigned extra;
    const int diff = q[0] - pred;
    if (diff == 0) {
        put_bits(s, dc->code[0], dc->len[0]);
    } else {
        const int n = category(diff, &extra);
        put_bits(s, dc->code[n], dc->len[n]);
        put_bits(s, extra, n);
    }
    int last = 63;
    while (last > 0 && q[last] == 0) --last;
    int run = 0;
    for (int k = 1; k <= last; ++k) {
        if (q[k] == 0) {
            ++run;
            continue;
        }
        for (; run >= 16; run -= 16) put_bits(s, ac->code[0xf0], ac->len[0xf0]);
        const int n = category(q[k], &extra);
        put_bits(s, ac->code[run * 16 + n], ac->len[run * 16 + n]);
        put_bits(s, extra, n);
        run = 0;
    }
    if (last != 63) put_bits(s, ac->code[0], ac->len[0]);
    return q[0];
}

/* ---- header ---------------------------------------------------------------------------------------------------- */
static void put_u16(sink_t *s, unsigned v)
{
    put_byte(s, v >> 8);
    put_byte(s, v & 255);
}
static void put_dht(sink_t *s, int cls_id, const uint8_t bits[16], const uint8_t *vals, int nvals)
{
    put_byte(s, cls_id);
    put_bytes(s, bits, 16);
    put_bytes(s, vals, nvals);
}
static void write_header(sink_t *s, const jfif_tables *t, int w, int h, int subsample)
{
    put_u16(s, 0xFFD8);
    put_u16(s, 0xFFE0); /* APP0: JFIF 1.1, aspect 1:1, no thumbnail */
    put_u16(s, 16);
    put_bytes(s, "JFIF", 5);
    put_u16(s, 0x0101);
    put_byte(s, 0);
    put_u16(s, 1);
    put_u16(s, 1);
    put_u16(s, 0);
    put_u16(s, 0xFFDB); /* DQT, both tables in one segment */
    put_u16(s, 2 + 65 + 65);
    put_byte(s, 0);
    put_bytes(s, t->qy, 64);
    put_byte(s, 1);
    put_bytes(s, t->quv, 64);
    put_u16(s, 0xFFC0); /* SOF0 */
    put_u16(s, 17);
    put_byte(s, 8);
    put_u16(s, (unsigned)h & 0xffff);
    put_u16(s, (unsigned)w & 0xffff);
    put_byte(s, 3);
    put_byte(s, 1); put_byte(s, subsample ? 0x22 : 0x11); put_byte(s, 0);
    put_byte(s, 2); put_byte(s, 0x11); put_byte(s, 1);
    put_byte(s, 3); put_byte(s, 0x11); put_byte(s, 1);
    put_u16(s, 0xFFC4); /* DHT, four tables in one segment */
    put_u16(s, 2 + 2 * (17 + 12) + 2 * (17 + 162));
    put_dht(s, 0x00, K3_DC_LUMA_BITS, K_DC_VALS, 12);
    put_dht(s, 0x10, K5_AC_LUMA_BITS, K5_AC_LUMA_VALS, 162);
    put_dht(s, 0x01, K4_DC_CHROMA_BITS, K_DC_VALS, 12);
    put_dht(s, 0x11, K6_AC_CHROMA_BITS, K6_AC_CHROMA_VALS, 162);
    put_u16(s, 0xFFDA); /* SOS */
    put_u16(s, 12);
    put_byte(s, 3);
    put_byte(s, 1); put_byte(s, 0x00);
    put_byte(s, 2); put_byte(s, 0x11);
    put_byte(s, 3); put_byte(s, 0x11);
    put_byte(s, 0); put_byte(s, 63); put_byte(s, 0);
}

/* ---- pixels ---------------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *px;
    int w, h, comp;
    size_t stride;
} image_t;

/* float Y-128, Cb, Cr of the pixel at (x, y), coordinates clamped to the image (stb :1533-1543) */
static void ycc_at(const image_t *im, int x, int y, float *Y, float *U, float *V)
{
    if (x >= im->w) x = im->w - 1;
    if (y >= im->h) y = im->h - 1;
    const uint8_t *p = im->px + (size_t)y * im->stride + (size_t)x * (size_t)im->comp;
    const float r = p[0], g = p[im->comp > 2 ? 1 : 0], b = p[im->comp > 2 ? 2 : 0];
    *Y = +0.29900f * r + 0.58700f * g + 0.11400f * b - 128;
    *U = -0.16874f * r - 0.33126f * g + 0.50000f * b;
    *V = +0.50000f * r - 0.41869f * g - 0.08131f * b;
}

/* ---- the encoder ----------------------------------------------------------------------------------------------- */
/* coefs (optional): quantised coefficients of every data unit in stream order, 64 int16 each, zig-zag order. */
int oracle_jfif_encode(const uint8_t *px, int w, int h, int comp, size_t stride, int quality, int force_subsample,
                       uint8_t *out, size_t cap, size_t *out_len, int16_t *coefs)
{
    if (!px || w <= 0 || h <= 0 || comp < 1 || comp > 4) return -1;
    jfif_tables t;
    make_tables(quality, &t);
    const int q0 = quality ? quality : 90;
    const int subsample = force_subsample < 0 ? (q0 <= 90) : (force_subsample != 0);
    sink_t s = {out, cap, 0, 0, 0};
    write_header(&s, &t, w, h, subsample);
    const image_t im = {px, w, h, comp, stride};
    int py = 0, pu = 0, pv = 0, q[64];
    size_t unit = 0;
#define EMIT(samples, st, mult, pred, dc, ac)                                   \
    do {                                                                        \
        transform_unit(samples, st, mult, t.rank, q);                           \
        if (coefs) {                                                            \
            for (int i_ = 0; i_ < 64; ++i_) coefs[unit * 64 + i_] = (int16_t)q[i_]; \
        }                                                                       \
        ++unit;                                                                 \
        pred = entropy_unit(&s, q, pred, dc, ac);                               \
    } while (0)
    if (subsample) {
        for (int y0 = 0; y0 < h; y0 += 16)
            for (int x0 = 0; x0 < w; x0 += 16) {
                float Y[256], U[256], V[256], cu[64], cv[64];
                for (int y = 0; y < 16; ++y)
                    for (int x = 0; x < 16; ++x) ycc_at(&im, x0 + x, y0 + y, &Y[y * 16 + x], &U[y * 16 + x], &V[y * 16 + x]);
                EMIT(Y, 16, t.dy, py, &t.dc_y, &t.ac_y);
                EMIT(Y + 8, 16, t.dy, py, &t.dc_y, &t.ac_y);
                EMIT(Y + 128, 16, t.dy, py, &t.dc_y, &t.ac_y);
                EMIT(Y + 136, 16, t.dy, py, &t.dc_y, &t.ac_y);
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) {
                        const int j = y * 32 + x * 2;
                        cu[y * 8 + x] = (U[j] + U[j + 1] + U[j + 16] + U[j + 17]) * 0.25f;
                        cv[y * 8 + x] = (V[j] + V[j + 1] + V[j + 16] + V[j + 17]) * 0.25f;
                    }
                EMIT(cu, 8, t.duv, pu, &t.dc_c, &t.ac_c);
                EMIT(cv, 8, t.duv, pv, &t.dc_c, &t.ac_c);
            }
    } else {
        for (int y0 = 0; y0 < h; y0 += 8)
            for (int x0 = 0; x0 < w; x0 += 8) {
                float Y[64], U[64], V[64];
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) ycc_at(&im, x0 + x, y0 + y, &Y[y * 8 + x], &U[y * 8 + x], &V[y * 8 + x]);
                EMIT(Y, 8, t.dy, py, &t.dc_y, &t.ac_y);
                EMIT(U, 8, t.duv, pu, &t.dc_c, &t.ac_c);
                EMIT(V, 8, t.duv, pv, &t.dc_c, &t.ac_c);
            }
    }
#undef EMIT
    put_bits(&s, 0x7f, 7); /* pad with ones; a trailing partial byte is dropped (stb :1586) */
    put_byte(&s, 0xFF);
    put_byte(&s, 0xD9);
    if (out_len) *out_len = s.n;
    return s.n <= cap ? 0 : -3;
}

/* Number of data units oracle_jfif_encode produces (for sizing `coefs`). */
size_t oracle_jfif_unit_count(int w, int h, int quality, int force_subsample)
{
    const int q0 = quality ? quality : 90;
    const int subsample = force_subsample < 0 ? (q0 <= 90) : (force_subsample != 0);
    if (subsample) return (size_t)((w + 15) / 16) * (size_t)((h + 15) / 16) * 6;
    return (size_t)((w + 7) / 8) * (size_t)((h + 7) / 8) * 3;
}

/* Tables for inspection by the tests: quantisers (zig-zag order), multipliers (natural order), zig-zag ranks. */
void oracle_jfif_tables(int quality, uint8_t qy[64], uint8_t quv[64], float dy[64], float duv[64], uint8_t rank[64])
{
    jfif_tables t;
    make_tables(quality, &t);
    memcpy(qy, t.qy, 64);
    memcpy(quv, t.quv, 64);
    memcpy(dy, t.dy, sizeof t.dy);
    memcpy(duv, t.duv, sizeof t.duv);
    memcpy(rank, t.rank, 64);
}

/* which: 0 DC luma, 1 AC luma, 2 DC chroma, 3 AC chroma */
void oracle_jfif_huffman(int which, uint16_t code[256], uint8_t len[256])
{
    jfif_tables t;
    make_tables(75, &t);
    const huff_t *h = which == 0 ? &t.dc_y : which == 1 ? &t.ac_y : which == 2 ? &t.dc_c : &t.ac_c;
    memcpy(code, h->code, sizeof h->code);
    memcpy(len, h->len, sizeof h->len);
}
