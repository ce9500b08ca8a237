// lz4-jpeg_b200/csrc/jpeg_encode.cu — "JPEG-like" per-8x8-group encoder (reference dialect) for sm_100a.
//
// One fused kernel replaces the encode half of the reference's main()/process()
// (Algorithms/sequential/JPEG/JPEG.c:1109-1249, Algorithms/parallel/JPEG/JPEG.c:1103-1252):
//   build_luminance_matrix / build_rChrominance_matrix / build_bChrominance_matrix   JPEG.c:114-185
//   chroma_subsample (keeps odd columns) / divide_image (zero padded 8x8 / 8x4 tiles)  JPEG.c:302, :496
//   discrete_cosine_transform / Quantize (truncating)                                  JPEG.c:451, :621
//   zigzag_pattern / RLE over all values                                               JPEG.c:693, :767
//   encode_huffman (array heap without sift-up) / generate_encoded_sequence            JPEG.c:864-1097, :993
// plus what the reference leaves implicit: packing the '0'/'1' strings into bytes and concatenating
// the per-group records (device-wide decoupled look-back over record sizes).
//
// Work decomposition: the reference spawns one thread per 8x8 group (P-JPG:1297-1302); here one CUDA
// thread owns one group; every WARP of the one persistent 512-thread CTA per SM pulls 32-group tiles from a
// ticket counter on its own (no CTA-wide barrier in the kernel).  Every per-thread array lives in shared memory WORD-INTERLEAVED across the 32 lanes of its warp
// (word w of lane l at warp_base + (w*32 + l)*4): whatever data-dependent index a lane uses, it stays in
// its own bank, so the divergent heap / table walks of the Huffman construction are bank-conflict free.
// 81 words per thread (symbol table 33 | counts / heap / codes 32 | child links 16) let 22 warps share the 227 KB of one SM.
//
// Exactness (results are bit-identical to the reference built with gcc x86-64 SSE2 -O2 -ffp-contract=off):
//   colour   the reference evaluates 0.299*r+0.587*g+0.114*b (etc.) in double and truncates.  The exact
//            rational value is S/1000 with S an integer; unless S is a multiple of 1000 it is >= 1e-3 away
//            from every integer while the double evaluation errs by < 1e-12, so floor(S/1000) is the
//            answer; for S % 1000 == 0 (one pixel in a thousand) the double expression itself is evaluated
//            with explicit round-to-nearest mul/add in the reference's left-to-right order (no FMA).
//   DCT      a separable double-precision evaluation with even/odd butterflies gives every quotient to
//            ~1e-12; a coefficient whose quotient lies within 1e-6 of a non-zero integer (where truncation
//            could go either way) is re-evaluated in the reference's own summation order (x outer, y
//            inner, (corr*cos_x)*cos_y, no FMA) with the exact cos()/sqrt() doubles glibc returns
//            (jpeg_tables.inc).
//   Huffman  the reference's array-heap procedure is simulated step by step per (group, channel).
// Channels outside the fast path's limits (a quantised value outside int8, more than 32 distinct symbols,
// a code longer than 26 bits) are redone by a general per-thread routine on local-memory arrays.
#include "common.cuh"

#include <vector>

#include <stdlib.h>

namespace jpgk {

#include "jpeg_tables.inc"

constexpr int THREADS = 704;       // 22 warps, each working on its own 32-group tile
constexpr int NWARPS = THREADS / 32;
constexpr int REC_BYTES = 256;     // max packed record: (1023 + 511 + 511) bits (JPEG.c:1248, :1286, :1320)
constexpr int REC_WORDS = REC_BYTES / 4;

// per-thread workspace, in 32-bit words (interleaved across the warp)
constexpr int W_LUT = 0;           // u8[129]: symbol + 64 -> (epoch << 5) | slot, symbols -64 .. 64 (33 words)
constexpr int W_CNT = 33;          // u32[32] counts -> u16 heap -> u16[63] code words (see entropy_fast)
constexpr int W_PAR = 65;          // u16[31]: children of the internal nodes; int8 coefficients while the DCT runs
constexpr int WS_WORDS = 81;
constexpr int SYM_MIN = -64, SYM_MAX = 64; // table range; a DC value outside it is carried in a register (see entropy_fast)
// DCT phase view of the same words
constexpr int W_T = 0;             // double[16]: row-pass results (8 rows x 2 columns), 64-bit interleaved = words 0..31
constexpr int W_SMP = 33;          // u8[128]: samples lum[64] | Cr[32] | Cb[32]
constexpr int FAST_MAXSYM = 32;
constexpr int FAST_MAXLEN = 12;

constexpr int SM_WS = THREADS * WS_WORDS * 4; // 228096
constexpr int SM_TOTAL = SM_WS;
static_assert(SM_TOTAL <= 227 * 1024, "exceeds B200 shared memory per CTA");

struct Params {
    const uint8_t *rgba;
    int w, h;
    size_t stride;
    size_t first_group, ngroups;
    uint8_t *out;
    size_t out_cap;
    uint64_t *group_offsets; // ngroups + 1
    uint16_t *group_bits;    // optional, 3 per group
    int16_t *coefs;          // optional, 128 per group
    uint64_t *result;        // [0] length, [1] groups outside the reference's defined behaviour, [2] error flags
    uint64_t *status;        // [0] ticket, [1..] look-back words (one per 32-group tile)
    uint32_t *scratch;       // per warp: REC_WORDS x 32 words of record staging, word-major
    uint32_t ntiles;
    uint64_t offs_bias;      // added to every group_offsets entry (base of this shard in a larger stream)
    int force_slow;          // test hook: route every channel through the general routine
    uint32_t img_groups;     // batch: groups per image (0 = one image); group g belongs to image g / img_groups
    size_t img_stride;       // batch: bytes between consecutive images
    int bpp;                 // bytes per pixel of the image: 4 (r g b a) or 3 (r g b)
    const uint8_t *planar;   // or: the samples of every group as the reference's PixelGroup holds them (JPEG.c:42-46):
                             // lum_values[64], b_values[32], r_values[32] = 128 bytes per group; rgba is then unused
};

// JPEG.c:12-27 as doubles (the chroma table is consumed as 8 rows x 4 columns)
__constant__ double kQLum[64] = {
    8.0, 6.0, 6.0, 8.0, 10.0, 14.0, 18.0, 22.0,
    6.0, 6.0, 7.0, 9.0, 12.0, 20.0, 22.0, 20.0,
    6.0, 7.0, 8.0, 10.0, 14.0, 22.0, 25.0, 22.0,
    8.0, 9.0, 10.0, 14.0, 18.0, 28.0, 27.0, 22.0,
    10.0, 12.0, 14.0, 18.0, 22.0, 35.0, 33.0, 26.0,
    14.0, 18.0, 22.0, 22.0, 27.0, 33.0, 36.0, 30.0,
    18.0, 22.0, 26.0, 28.0, 33.0, 40.0, 40.0, 34.0,
    22.0, 26.0, 28.0, 30.0, 36.0, 34.0, 35.0, 33.0,
};
__constant__ double kQChr[32] = {
    17.0, 18.0, 24.0, 47.0, 18.0, 21.0, 26.0, 66.0,
    24.0, 26.0, 56.0, 99.0, 47.0, 66.0, 99.0, 99.0,
    66.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0,
    99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0,
};
// zig-zag position of row-major index (inverse of the walk in JPEG.c:693-727), evaluated at compile time
template <int W, int H>
struct ZigZag {
    int pos[W * H];
    constexpr ZigZag() : pos()
    {
        int index = 0;
        for (int sum = 0; sum < W + H - 1; ++sum) {
            int start_row = (sum < W) ? 0 : sum - W + 1;
            int end_row = (sum < H) ? sum : H - 1;
            if (sum % 2 == 0) {
                for (int row = end_row; row >= start_row; --row) {
                    int col = sum - row;
                    if (col < W) pos[row * W + col] = index++;
                }
            } else {
                for (int row = start_row; row <= end_row; ++row) {
                    int col = sum - row;
                    if (col < W) pos[row * W + col] = index++;
                }
            }
        }
    }
};
__constant__ const ZigZag<8, 8> kZZ8 = ZigZag<8, 8>(); // dynamically indexed copies for the general routine
__constant__ const ZigZag<4, 8> kZZ4 = ZigZag<4, 8>();

// ---- interleaved workspace accessors: wl = warp workspace base + lane ------------------------------------
__device__ __forceinline__ uint32_t &ws_w(uint32_t *wl, int word) { return wl[word * 32]; }
__device__ __forceinline__ uint8_t &ws_b(uint32_t *wl, int word0, int byte)
{
    return reinterpret_cast<uint8_t *>(wl + (word0 + (byte >> 2)) * 32)[byte & 3];
}
__device__ __forceinline__ uint16_t &ws_h(uint32_t *wl, int word0, int half)
{
    return reinterpret_cast<uint16_t *>(wl + (word0 + (half >> 1)) * 32)[half & 1];
}
// 64-bit interleaved view of words [0, 64): double d of lane l at warp_base + (d*32 + l)*8
__device__ __forceinline__ double &ws_d(uint32_t *wbase, int lane, int d) { return reinterpret_cast<double *>(wbase)[d * 32 + lane]; }

#include "jpeg_colour.cuh"

// ---- the reference's own summation, JPEG.c:471-490, for one coefficient ------------------------------------
// smp(i) returns sample i of the channel (row-major, W columns)
template <int W, class F>
__device__ __forceinline__ double exact_coef_t(F smp, int u, int v)
{
    double sum = 0.0;
#pragma unroll 1
    for (int x = 0; x < 8; ++x) {
        const double cx = kCos8[x * 8 + u];
#pragma unroll 1
        for (int y = 0; y < W; ++y) {
            const double cy = (W == 8) ? kCos8[y * 8 + v] : kCos4[y * 4 + v];
            const double corr = (double)(smp(x * W + y) - 128);
            sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(corr, cx), cy));
        }
    }
    const double au = u == 0 ? kAlpha8[0] : kAlpha8[1];
    const double av = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) : (v == 0 ? kAlpha4[0] : kAlpha4[1]);
    return __dmul_rn(__dmul_rn(au, av), sum);
}
template <int W>
__device__ __noinline__ int exact_quant_ws(uint32_t *wl, int byte0, int u, int v)
{
    const double ce = exact_coef_t<W>([&](int i) { return (int)ws_b(wl, W_SMP, byte0 + i); }, u, v);
    return (int)__ddiv_rn(ce, (W == 8) ? kQLum[u * W + v] : kQChr[u * W + v]);
}

// Fast-path quantisation constants: alpha_u * alpha_v / Q[u][v] * 2^20 (the quotient in 12.20 fixed point)
template <int W>
struct QuantScale {
    double m[W * 8];
    constexpr QuantScale(const int *q) : m()
    {
        for (int u = 0; u < 8; ++u)
            for (int v = 0; v < W; ++v) {
                const double au = u == 0 ? 0.35355339059327376220 : 0.5;
                const double av = W == 8 ? (v == 0 ? 0.35355339059327376220 : 0.5) : (v == 0 ? 0.5 : 0.70710678118654752440);
                m[u * W + v] = au * av * 1048576.0 / (double)q[u * W + v];
            }
    }
};
constexpr int kQLumI[64] = {8,  6,  6,  8,  10, 14, 18, 22, 6,  6,  7,  9,  12, 20, 22, 20, 6,  7,  8,  10, 14, 22,
                            25, 22, 8,  9,  10, 14, 18, 28, 27, 22, 10, 12, 14, 18, 22, 35, 33, 26, 14, 18, 22, 22,
                            27, 33, 36, 30, 18, 22, 26, 28, 33, 40, 40, 34, 22, 26, 28, 30, 36, 34, 35, 33}; // JPEG.c:12-20
constexpr int kQChrI[32] = {17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99,
                            66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}; // JPEG.c:22-27 as 8 rows x 4
__constant__ const QuantScale<8> kMLum = QuantScale<8>(kQLumI);
__constant__ const QuantScale<4> kMChr = QuantScale<4>(kQChrI);

// Quantise one fast-path coefficient: qs = quotient * 2^20 (accurate to ~1e-11 of the quotient).  Returns false
// when the quotient lies within 2^-19 of a non-zero integer, where truncation could go either way.
__device__ __forceinline__ bool quant_fast(double qs, int &t)
{
    const int fx = __double2int_rz(qs);
    const unsigned a = (unsigned)abs(fx), ti = a >> 20, fr = a & 0xFFFFFu;
    t = fx < 0 ? -(int)ti : (int)ti; // truncation toward zero, JPEG.c:627
    return !((fr >= 0xFFFFEu) || (fr <= 1u && ti >= 1u));
}

// Column pass + quantise + zig-zag of ONE column of row-pass results, shared by luma (W = 8) and chroma (W = 4).
// T holds 8 rows x 2 columns of doubles, this call consumes column vi (0 or 1), which is frequency v of the channel.
// Results go to byte out0 + zigzag(u, v) of the W_PAR words.  Kept out of line and rolled: the kernel's hot code
// has to stay well inside the instruction cache because the warps of an SM run desynchronised.
__device__ __noinline__ unsigned col_pass(uint32_t *wl, uint32_t *wbase, int lane, int vi, int v, int W, int byte0, int out0,
                                          int16_t *co)
{
    unsigned wide = 0;
    double s[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double a = ws_d(wbase, lane, j * 2 + vi), b = ws_d(wbase, lane, (7 - j) * 2 + vi);
        s[j] = a + b; // cos8[7-x][u] = (-1)^u cos8[x][u]
        d[j] = a - b;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { // unrolled: the 32 cosine operands become immediates instead of indexed constant loads
#pragma unroll
        for (int odd = 0; odd < 2; ++odd) {
            const int u = 2 * k + odd;
            double acc = (odd ? d[0] : s[0]) * kCos8[0 * 8 + u];
#pragma unroll
            for (int j = 1; j < 4; ++j) acc = fma(odd ? d[j] : s[j], kCos8[j * 8 + u], acc);
            const int idx = u * W + v;
            int t;
            const double qs = acc * (W == 8 ? kMLum.m[idx] : kMChr.m[idx]);
            if (!quant_fast(qs, t)) {
                // Luma DC is sum(corr) / 64 exactly: "near an integer" means it IS that integer n, and the reference's
                // (alpha0 * alpha0) * sum / 8 with alpha0^2 = 0.125 (1 + 2^-52) never falls below |n|, so it truncates to n.
                if (W == 8 && u == 0 && v == 0) t = __double2int_rn(qs * (1.0 / 1048576.0));
                else t = W == 8 ? exact_quant_ws<8>(wl, byte0, u, v) : exact_quant_ws<4>(wl, byte0, u, v);
            }
            // the DC may be any int8 (it gets a register-held table slot), the others have to fit the symbol table
            wide |= (u == 0 && v == 0) ? ((t < -128) | (t > 127)) : ((t < SYM_MIN) | (t > SYM_MAX));
            ws_b(wl, W_PAR, out0 + (W == 8 ? kZZ8.pos[idx] : kZZ4.pos[idx])) = (uint8_t)t;
            if (co) co[idx] = (int16_t)t;
        }
    }
    return wide;
}

// Luma: 8x8 DCT + quantise + zig-zag.  The 64 int8 results go to bytes [0, 64) of the W_PAR words (zig-zag order);
// returns non-zero if some value does not fit the fast path.  Two output columns per pass: the row-pass results of
// a pass (8 x 2 doubles) are all that fits beside the samples in the 81-word workspace.
__device__ __forceinline__ unsigned dct_luma(uint32_t *wl, uint32_t *wbase, int lane, int16_t *co)
{
    unsigned wide = 0;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) { // columns v = 2 vi + par, vi = 2 (pass >> 1) + {0, 1}, par = pass & 1
        const int par = pass & 1, vi0 = 2 * (pass >> 1);
        double c[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) c[q][j] = kCos8[j * 8 + 2 * (vi0 + q) + par];
        // row pass: T[x][q] = sum_y corr[x][y] * cos8[y][v], using cos8[7-y][v] = (-1)^v cos8[y][v]
#pragma unroll 1
        for (int x = 0; x < 8; ++x) {
            const uint32_t w0 = ws_w(wl, W_SMP + 2 * x), w1 = ws_w(wl, W_SMP + 2 * x + 1);
            double e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int a = (w0 >> (8 * j)) & 0xFF, b = (w1 >> (8 * (3 - j))) & 0xFF; // y = j and y = 7 - j
                e[j] = (double)(par == 0 ? a + b - 256 : a - b);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                double acc = e[0] * c[q][0];
#pragma unroll
                for (int j = 1; j < 4; ++j) acc = fma(e[j], c[q][j], acc);
                ws_d(wbase, lane, x * 2 + q) = acc;
            }
        }
#pragma unroll 1
        for (int q = 0; q < 2; ++q) wide |= col_pass(wl, wbase, lane, q, 2 * (vi0 + q) + par, 8, 0, 0, co);
    }
    return wide;
}

// Chroma: 8 rows x 4 columns; the 32 int8 results go to bytes [out0, out0 + 32) of the W_PAR words
__device__ __forceinline__ unsigned dct_chroma(uint32_t *wl, uint32_t *wbase, int lane, int byte0, int out0, int16_t *co)
{
    unsigned wide = 0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) { // columns v = 2 half + {0, 1}
#pragma unroll 1
        for (int x = 0; x < 8; ++x) {
            const uint32_t w0 = ws_w(wl, W_SMP + (byte0 >> 2) + x);
            const int c0 = w0 & 0xFF, c1 = (w0 >> 8) & 0xFF, c2 = (w0 >> 16) & 0xFF, c3 = w0 >> 24;
            const double s0 = (double)(c0 + c3 - 256), s1 = (double)(c1 + c2 - 256), d0 = (double)(c0 - c3), d1 = (double)(c1 - c2);
            // v even uses the sums, v odd the differences (cos4[3-y][v] = (-1)^v cos4[y][v])
            ws_d(wbase, lane, x * 2 + 0) = fma(s1, kCos4[1 * 4 + 2 * half], s0 * kCos4[0 * 4 + 2 * half]);
            ws_d(wbase, lane, x * 2 + 1) = fma(d1, kCos4[1 * 4 + 2 * half + 1], d0 * kCos4[0 * 4 + 2 * half + 1]);
        }
#pragma unroll 1
        for (int q = 0; q < 2; ++q) wide |= col_pass(wl, wbase, lane, q, 2 * half + q, 4, byte0, out0, co);
    }
    return wide;
}

// ---- bit packing: records are staged word-major ([word][thread]) so that a warp's stores coalesce ----------
struct BitWriter {
    unsigned long long acc;
    int nbits;          // bits pending in acc (< 32 between calls)
    uint32_t *dst;      // this thread's column of the CTA's staging area
    int wpos;
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc = (acc << len) | code;
        nbits += len;
        if (nbits >= 32) {
            const uint32_t wv = (uint32_t)(acc >> (nbits - 32));
            if (wpos < REC_WORDS) dst[wpos * 32] = __byte_perm(wv, 0, 0x0123); // MSB-first bytes
            ++wpos;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nbits > 0) {
            const uint32_t wv = (uint32_t)(acc << (32 - nbits));
            if (wpos < REC_WORDS) dst[wpos * 32] = __byte_perm(wv, 0, 0x0123);
        }
    }
};

// ---- general routine (local-memory arrays): any values, any number of symbols -------------------------------
// Recomputes the channel from the pixels.  Follows JPEG.c:864-1007 literally (linear-search symbol table).
__device__ __noinline__ int slow_channel(const uint8_t *rgba, int w, int h, size_t stride, int bpp, size_t g, int ch, int16_t *co,
                                         BitWriter *bwp, int max_bits, int *bad, const uint8_t *planar_group)
{
    const int W = ch == 0 ? 8 : 4, N = 8 * W;
    uint8_t smp[64];
    int16_t z[64];
    if (planar_group) {
        const uint8_t *src = planar_group + (ch == 0 ? 0 : (ch == 1 ? 96 : 64)); // lum | b | r (JPEG.c:44-46); channel 1 is Cr
        for (int i = 0; i < N; ++i) smp[i] = src[i];
    } else {
        const size_t bpr = ((size_t)w + 7) / 8;
        const size_t brow = g / bpr, bcol = g % bpr;
        for (int lr = 0; lr < 8; ++lr)
            for (int c = 0; c < W; ++c) {
                const size_t row = brow * 8 + lr;
                const size_t col = bcol * 8 + (ch == 0 ? c : 2 * c + 1); // chroma keeps the odd columns (JPEG.c:302-375)
                const size_t gate = bcol * 8 + (ch == 0 ? c : 2 * c);    // ... filed under the even column (JPEG.c:543)
                int s = 0;
                if (row < (size_t)h && gate < (size_t)w && col < (size_t)w) {
                    const uint8_t *q = rgba + row * stride + col * (size_t)bpp;
                    s = ch == 0 ? luma_of(q[0], q[1], q[2]) : (ch == 1 ? cr_of(q[0], q[1], q[2]) : cb_of(q[0], q[1], q[2]));
                }
                smp[lr * W + c] = (uint8_t)s;
            }
    }
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < W; ++v) {
            // separable evaluation is not worth its code here: evaluate the double sum with FMAs, then fall back like the fast path
            double acc = 0.0;
            for (int x = 0; x < 8; ++x) {
                double row = 0.0;
                for (int y = 0; y < W; ++y)
                    row = fma((double)((int)smp[x * W + y] - 128), (W == 8) ? kCos8[y * 8 + v] : kCos4[y * 4 + v], row);
                acc = fma(row, kCos8[x * 8 + u], acc);
            }
            const double au = u == 0 ? kAlpha8[0] : kAlpha8[1];
            const double av = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) : (v == 0 ? kAlpha4[0] : kAlpha4[1]);
            int t;
            (void)au;
            (void)av;
            if (!quant_fast(acc * ((W == 8) ? kMLum.m[u * 8 + v] : kMChr.m[u * 4 + v]), t)) {
                const double ce = (W == 8) ? exact_coef_t<8>([&](int i) { return (int)smp[i]; }, u, v)
                                           : exact_coef_t<4>([&](int i) { return (int)smp[i]; }, u, v);
                t = (int)__ddiv_rn(ce, (W == 8) ? kQLum[u * 8 + v] : kQChr[u * 4 + v]);
            }
            z[(W == 8) ? kZZ8.pos[u * 8 + v] : kZZ4.pos[u * 4 + v]] = (int16_t)t;
            if (co) co[u * W + v] = (int16_t)t;
        }
    // RLE (JPEG.c:767-809) + calculate_frequency (JPEG.c:864-885)
    int16_t sym[130];
    uint8_t cnt[260], heap[130], par[260], rbit[260], seq[130];
    int k = 0, m = 0;
    for (int i = 0; i < N;) {
        const int v = z[i];
        int j = i + 1;
        while (j < N && z[j] == v) ++j;
        const int two[2] = {j - i, v};
        for (int s = 0; s < 2; ++s) {
            int slot = -1;
            for (int q = 0; q < k; ++q)
                if (sym[q] == two[s]) {
                    slot = q;
                    break;
                }
            if (slot < 0) {
                slot = k++;
                sym[slot] = (int16_t)two[s];
                cnt[slot] = 0;
            }
            cnt[slot]++;
            seq[m++] = (uint8_t)slot;
        }
        i = j;
    }
    auto heapify = [&](int size, int i) { // JPEG.c:894-911
        for (;;) {
            int smallest = i, l = 2 * i + 1, r = 2 * i + 2;
            if (l < size && cnt[heap[l]] < cnt[heap[smallest]]) smallest = l;
            if (r < size && cnt[heap[r]] < cnt[heap[smallest]]) smallest = r;
            if (smallest == i) return;
            const uint8_t t = heap[i];
            heap[i] = heap[smallest];
            heap[smallest] = t;
            i = smallest;
        }
    };
    for (int i = 0; i < k; ++i) heap[i] = (uint8_t)i;
    for (int i = k / 2 - 1; i >= 0; --i) heapify(k, i); // JPEG.c:913-934
    int size = k, next = k;
    while (size > 1) { // JPEG.c:936-961: pop two, append their parent at the END of the array (no sift-up)
        const int left = heap[0];
        heap[0] = heap[--size];
        heapify(size, 0);
        const int right = heap[0];
        heap[0] = heap[--size];
        heapify(size, 0);
        cnt[next] = (uint8_t)(cnt[left] + cnt[right]);
        par[left] = (uint8_t)next;
        par[right] = (uint8_t)next;
        rbit[left] = 0;
        rbit[right] = 1;
        heap[size++] = (uint8_t)next;
        ++next;
    }
    const int root = heap[0];
    // assign_codes (JPEG.c:963-982) read bottom-up, then generate_encoded_sequence (JPEG.c:993-1007)
    BitWriter bw = *bwp;
    int bits = 0;
    for (int i = 0; i < m; ++i) {
        uint32_t c = 0;
        int l = 0, node = seq[i];
        while (node != root) {
            if (l < 32) c |= (uint32_t)rbit[node] << l;
            ++l;
            node = par[node];
        }
        if (l > 31) { // HuffmanCode.code is char[32] (JPEG.c:861)
            *bad = 1;
            l = 31;
        }
        if (l > 16) {
            bw.put(c >> 16, l - 16);
            bw.put(c & 0xFFFF, 16);
        } else {
            bw.put(c, l);
        }
        bits += l;
    }
    if (bits > max_bits) *bad = 1; // char encoded_sequence[1024] / [512] (JPEG.c:1248, :1286)
    *bwp = bw;
    return bits;
}

// ---- fast path: symbol table + array heap in the interleaved workspace --------------------------------------
// Words 64..95 hold, in turn: CNT u32[32] (occurrences per slot, while scanning) -> HEAP u16[1..k] (the 1-based
// array heap of (count << 8) | node id, built in place from CNT) -> CW u16[63] ((length << 12) | code per node).
// Words 96..111 hold CHILD u16[31]: left | right << 8 of internal node k + i.
struct Ent {
    uint8_t *lb;    // byte address of this lane's word 0
    uint32_t epoch; // pre-shifted: epoch << 5
    uint32_t k;
    uint32_t over;
    uint32_t dc_slot; // slot of a DC value outside the table range (it can only be the value of the channel's first run)
};
__device__ __forceinline__ bool in_table(int sym) { return (unsigned)(sym - SYM_MIN) <= (unsigned)(SYM_MAX - SYM_MIN); }
__device__ __forceinline__ uint8_t *lut_addr(uint8_t *lb, int sym)
{
    const uint32_t idx = (uint32_t)(sym - SYM_MIN);
    return lb + ((idx & 0xFCu) << 5) + (idx & 3u); // word idx >> 2 of the lane, byte idx & 3
}
__device__ __forceinline__ void sym_count(Ent &E, int sym) // calculate_frequency, JPEG.c:864-885 (branch free)
{
    uint8_t *la = lut_addr(E.lb, sym);
    const uint32_t e = *la;
    const bool hit = (e & 0xE0u) == E.epoch;
    E.over |= (!hit && E.k >= (uint32_t)FAST_MAXSYM) ? 1u : 0u;
    const uint32_t slot = hit ? (e & 31u) : min(E.k, (uint32_t)FAST_MAXSYM - 1u);
    *la = (uint8_t)(E.epoch | slot);
    uint32_t *ca = reinterpret_cast<uint32_t *>(E.lb + (W_CNT * 128) + slot * 128);
    const uint32_t c = *ca;
    *ca = hit ? c + 1u : 1u;
    E.k += hit ? 0u : 1u;
}
__device__ __forceinline__ uint32_t sym_code(const Ent &E, int sym)
{
    const uint32_t slot = *lut_addr(E.lb, sym) & 31u;
    return *reinterpret_cast<const uint16_t *>(E.lb + (W_CNT * 128) + ((slot >> 1) << 7) + ((slot & 1u) << 1));
}
// one finished run (count, value) for calculate_frequency; `first` = it is the channel's first run
__device__ __forceinline__ void pair_count(Ent &E, int c, int v, bool first)
{
    sym_count(E, c);
    if (in_table(v)) {
        sym_count(E, v);
    } else if (first && E.k < (uint32_t)FAST_MAXSYM) { // the DC: a symbol of its own, next slot in first-appearance order
        E.dc_slot = E.k;
        *reinterpret_cast<uint32_t *>(E.lb + (W_CNT * 128) + E.k * 128) = 1u;
        ++E.k;
    } else {
        E.over = 1;
    }
}
// generate_encoded_sequence, JPEG.c:993-1007: the codes of one (count, value) pair, at most 12 + 12 bits
__device__ __forceinline__ void pair_emit(const Ent &E, BitWriter &bw, int c, int v)
{
    const uint32_t cc = sym_code(E, c);
    const uint32_t cv = in_table(v) ? sym_code(E, v)
                                    : *reinterpret_cast<const uint16_t *>(E.lb + (W_CNT * 128) + ((E.dc_slot >> 1) << 7) + ((E.dc_slot & 1u) << 1));
    const uint32_t lv = cv >> 12;
    bw.put(((cc & 0xFFFu) << lv) | (cv & 0xFFFu), (int)((cc >> 12) + lv));
}
// heap entry j (1-based) is the u16 at byte hb + (j >> 1) * 128 + (j & 1) * 2
__device__ __forceinline__ uint16_t *heap_at(uint8_t *hb, uint32_t j)
{
    return reinterpret_cast<uint16_t *>(hb + ((j >> 1) << 7) + ((j & 1u) << 1));
}
// sift entry e down from 1-based position j in a heap of n entries (heapify, JPEG.c:894-911)
__device__ __forceinline__ void sift(uint8_t *hb, uint32_t n, uint32_t j, uint32_t e)
{
    const uint32_t ec = e >> 8;
    uint16_t *self = heap_at(hb, j);
    while (2 * j <= n) {
        uint8_t *pa = hb + (j << 7);                                   // children 2j (low half) and 2j+1 (high half)
        const uint32_t pair = *reinterpret_cast<const uint32_t *>(pa);
        const uint32_t lo = pair & 0xFFFFu, hi = pair >> 16;
        const uint32_t lc = lo >> 8, hc = (2 * j + 1 <= n) ? (hi >> 8) : 0x1FFu;
        // smallest = i; if (l.count < smallest.count) smallest = l; if (r.count < smallest.count) smallest = r;
        const bool tl = lc < ec;
        const uint32_t sc = tl ? lc : ec;
        const bool tr = hc < sc;
        if (!(tl | tr)) break;
        *self = (uint16_t)(tr ? hi : lo);
        self = reinterpret_cast<uint16_t *>(pa + (tr ? 2 : 0));
        j = 2 * j + (tr ? 1u : 0u);
    }
    *self = (uint16_t)e;
}

// The 128 int8 coefficients of a group live in 32 registers: luma cz[0..16), Cr cz[16..24), Cb cz[24..32).
// Registers cannot be indexed dynamically, so the scans below pull 16 coefficients at a time into a 4-register
// window (a select chain) and shift the window by one byte per step: the loop body exists once.
__device__ __forceinline__ void window_load(uint32_t (&w)[4], const uint32_t (&cz)[32], int q8)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t v = cz[i];
#pragma unroll
        for (int q = 1; q < 8; ++q) v = (q8 == q) ? cz[4 * q + i] : v;
        w[i] = v;
    }
}
__device__ __forceinline__ int window_pop(uint32_t (&w)[4])
{
    const int v = (int)(int8_t)(w[0] & 0xFFu);
    w[0] = __funnelshift_r(w[0], w[1], 8);
    w[1] = __funnelshift_r(w[1], w[2], 8);
    w[2] = __funnelshift_r(w[2], w[3], 8);
    w[3] >>= 8;
    return v;
}

// One channel through RLE + Huffman + emission.  q0 = index of the channel's first 16-coefficient window
// (luma 0, Cr 4, Cb 6), nq = number of windows (4 or 2).
// Returns the number of bits, or -1 if the channel has to go through slow_channel.
__device__ __forceinline__ int entropy_fast(uint32_t *wl, uint32_t epoch, const uint32_t (&cz)[32], int q0, int nq, BitWriter &bw)
{
    Ent E;
    E.lb = reinterpret_cast<uint8_t *>(wl);
    E.epoch = epoch << 5;
    E.k = 0;
    E.over = 0;
    E.dc_slot = 0;
    uint8_t *hb = E.lb + W_CNT * 128;
    const int n = 16 * nq;
    uint32_t w[4];
    // pass 1: RLE (JPEG.c:767-809) feeding the symbol table in first-appearance order
    window_load(w, cz, q0);
    int prev = (int)(int8_t)(w[0] & 0xFFu), run = 0;
    bool first = true;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        if (i && (i & 15) == 0) window_load(w, cz, q0 + (i >> 4));
        const int v = window_pop(w);
        if (v != prev) {
            pair_count(E, run, prev, first);
            first = false;
            prev = v;
            run = 1;
        } else {
            ++run;
        }
    }
    pair_count(E, run, prev, first);
    if (E.over) return -1;
    const uint32_t k = E.k;
    // the heap array in the reference's initial (first-appearance) order, in place over CNT: entry j sits in word
    // j >> 1 <= j - 1, which has been consumed by the time it is overwritten
    for (uint32_t j = 1; j <= k; ++j) {
        const uint32_t c = *reinterpret_cast<const uint32_t *>(hb + ((j - 1) << 7));
        *heap_at(hb, j) = (uint16_t)((c << 8) | (j - 1));
    }
    // build_heap (JPEG.c:913-934)
    for (uint32_t j = k >> 1; j >= 1; --j) sift(hb, k, j, *heap_at(hb, j));
    // build_huffman_tree (JPEG.c:936-961): pop two, append their parent at the END of the array (no sift-up)
    uint32_t hn = k, next = k;
    int bits = 0;
    uint8_t *chb = E.lb + W_PAR * 128;
    while (hn > 1) {
        uint32_t pop[2];
#pragma unroll 1
        for (int t = 0; t < 2; ++t) { // heap[0] = heap[--size]; heapify(0)
            pop[t] = *heap_at(hb, 1);
            const uint32_t e = *heap_at(hb, hn);
            --hn;
            if (hn >= 1) sift(hb, hn, 1, e);
        }
        const uint32_t c = (pop[0] >> 8) + (pop[1] >> 8);
        bits += (int)c; // total code length = sum of the internal nodes' counts
        *heap_at(chb, next - k) = (uint16_t)((pop[0] & 0xFFu) | ((pop[1] & 0xFFu) << 8));
        ++hn;
        *heap_at(hb, hn) = (uint16_t)((c << 8) | next);
        ++next;
    }
    // assign_codes (JPEG.c:963-982), top down: left = '0', right = '1'.  Children always have smaller ids than
    // their parent, so one descending sweep over the internal nodes reaches every node after its parent.
    uint32_t deep = 0;
    *heap_at(hb, k > 1 ? next - 1 : 0) = 0; // root: empty code (the heap array is dead: CW reuses its words)
    for (uint32_t t = next; t-- > k;) {
        const uint32_t cw = *heap_at(hb, t);
        const uint32_t ch = *heap_at(chb, t - k);
        const uint32_t len = (cw >> 12) + 1u, code = (cw & 0xFFFu) << 1;
        deep |= len > (uint32_t)FAST_MAXLEN ? 1u : 0u;
        const uint32_t down = ((len & 15u) << 12) | (code & 0xFFFu);
        *heap_at(hb, ch & 0xFFu) = (uint16_t)down;
        *heap_at(hb, ch >> 8) = (uint16_t)(down | 1u);
    }
    if (deep) return -1;
    // pass 2: generate_encoded_sequence (JPEG.c:993-1007)
    window_load(w, cz, q0);
    prev = (int)(int8_t)(w[0] & 0xFFu);
    run = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        if (i && (i & 15) == 0) window_load(w, cz, q0 + (i >> 4));
        const int v = window_pop(w);
        if (v != prev) {
            pair_emit(E, bw, run, prev);
            prev = v;
            run = 1;
        } else {
            ++run;
        }
    }
    pair_emit(E, bw, run, prev);
    return bits;
}

// BPP: bytes per pixel of the image, 4 (r g b a) or 3 (r g b) — a template parameter because the kernel's hot code has to stay small
template <int BPP>
__global__ void __launch_bounds__(THREADS, 1) jpeg_encode_kernel(Params P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t *wbase = reinterpret_cast<uint32_t *>(smem) + warp * (WS_WORDS * 32);
    uint32_t *wl = wbase + lane;
    // each warp is its own producer: 32 groups per tile, no CTA-wide barrier anywhere, so warps drift apart and
    // their DCT (fp64 pipe), Huffman (integer / shared memory) and copy-out (L2 latency) phases overlap
    uint32_t *stage2 = P.scratch + ((size_t)blockIdx.x * NWARPS + warp) * (2 * REC_WORDS * 32) + lane; // two staging buffers
    const size_t bpr = ((size_t)P.w + 7) / 8;
    constexpr bool rgb = BPP == 3; // three bytes per pixel: a group row is 24 bytes, read as three 8-byte words
    const bool aligned = ((reinterpret_cast<uintptr_t>(P.rgba) | P.stride | P.img_stride) & (rgb ? 7 : 15)) == 0;
    // the tile computed in the previous iteration: its records wait in the other staging buffer until its offset
    // is fetched, one tile later, when its predecessors have (almost always) published theirs
    bool pend = false;
    long long p_tile = 0;
    unsigned int p_off = 0, p_bytes = 0, p_total = 0, p_bits = 0;
    int buf = 0;

    for (;;) {
        long long tile = 0;
        if (lane == 0) tile = (long long)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const bool have = tile < (long long)P.ntiles;
        uint32_t *stage = stage2 + buf * (REC_WORDS * 32);
        const size_t gl = (size_t)tile * 32 + lane; // group index inside this call
        const bool active = have && gl < P.ngroups;
        unsigned int rec_bytes = 0;
        int bl = 0, br = 0, bb = 0;
        if (active) {
            size_t g = P.first_group + gl;
            const uint8_t *img = P.rgba;
            if (P.img_groups) { // batch: image g / img_groups, group g % img_groups of it
                const uint32_t im = (uint32_t)g / P.img_groups;
                g -= (size_t)im * P.img_groups;
                img += (size_t)im * P.img_stride;
            }
            const uint8_t *const pgroup = P.planar ? P.planar + (P.first_group + gl) * 128 : nullptr;
            const size_t brow = g / bpr, bcol = g % bpr;
            const size_t col0 = bcol * 8;
            // ---- colour conversion + 4:2:2 point subsampling + tiling -> samples u8[64 | 32 | 32]
            const bool full = brow * 8 + 8 <= (size_t)P.h && col0 + 8 <= (size_t)P.w;
            if (pgroup) { // the samples are given (PixelGroup order lum | b | r): no colour conversion, no tiling
                const uint32_t *src = reinterpret_cast<const uint32_t *>(pgroup);
#pragma unroll
                for (int i = 0; i < 16; ++i) ws_w(wl, W_SMP + i) = src[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ws_w(wl, W_SMP + 16 + i) = src[24 + i]; // r_values
                    ws_w(wl, W_SMP + 24 + i) = src[16 + i]; // b_values
                }
            } else if (full && aligned) {
                const uint8_t *rp = img + brow * 8 * P.stride + col0 * (size_t)BPP;
                // one row of the group: 8 x 4 bytes as two 16-byte words, or 8 x 3 bytes as three 8-byte words (streamed once)
                auto load_row = [&](const uint8_t *p, uint4 &a, uint4 &b) {
                    if (rgb) {
                        const uint2 t0 = __ldcs(reinterpret_cast<const uint2 *>(p)), t1 = __ldcs(reinterpret_cast<const uint2 *>(p) + 1);
                        const uint2 t2 = __ldcs(reinterpret_cast<const uint2 *>(p) + 2);
                        a = make_uint4(t0.x, t0.y, t1.x, t1.y);
                        b = make_uint4(t2.x, t2.y, 0u, 0u);
                    } else {
                        a = __ldcs(reinterpret_cast<const uint4 *>(p));
                        b = __ldcs(reinterpret_cast<const uint4 *>(p) + 1);
                    }
                };
                uint4 na, nb;
                load_row(rp, na, nb);
#pragma unroll 1
                for (int lr = 0; lr < 8; ++lr) {
                    uint32_t px[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
                    if (rgb) { // pixel j starts at byte 3 j: r | g << 8 | b << 16 in the low three bytes, like the four-byte form
                        const uint32_t w6[7] = {na.x, na.y, na.z, na.w, nb.x, nb.y, 0u};
#pragma unroll
                        for (int j = 0; j < 8; ++j) px[j] = __funnelshift_r(w6[(3 * j) >> 2], w6[((3 * j) >> 2) + 1], 8 * ((3 * j) & 3));
                    }
                    if (lr < 7) load_row(rp + (lr + 1) * P.stride, na, nb); // next row's pixels are in flight while this row is converted
                    uint32_t yy[2] = {0, 0}, cr = 0, cb = 0;
#pragma unroll 1
                    for (int it = 0; it < 4; ++it) { // two pixels per step; the odd one also gives the chroma sample
                        const int r0 = px[0] & 0xFF, g0 = (px[0] >> 8) & 0xFF, b0 = (px[0] >> 16) & 0xFF;
                        const int r1 = px[1] & 0xFF, g1 = (px[1] >> 8) & 0xFF, b1 = (px[1] >> 16) & 0xFF;
                        const uint32_t pair = (uint32_t)luma_of(r0, g0, b0) | ((uint32_t)luma_of(r1, g1, b1) << 8);
                        const uint32_t sh = pair << (16 * (it & 1));
                        yy[0] |= it < 2 ? sh : 0u;
                        yy[1] |= it < 2 ? 0u : sh;
                        cr |= (uint32_t)cr_of(r1, g1, b1) << (8 * it); // chroma of local column 2 it = original chroma at column 2 it + 1
                        cb |= (uint32_t)cb_of(r1, g1, b1) << (8 * it);
#pragma unroll
                        for (int q = 0; q < 6; ++q) px[q] = px[q + 2];
                    }
                    ws_w(wl, W_SMP + 2 * lr) = yy[0];
                    ws_w(wl, W_SMP + 2 * lr + 1) = yy[1];
                    ws_w(wl, W_SMP + 16 + lr) = cr;
                    ws_w(wl, W_SMP + 24 + lr) = cb;
                }
            } else {
#pragma unroll 1
                for (int lr = 0; lr < 8; ++lr) {
                    const size_t row = brow * 8 + lr;
                    uint32_t y0 = 0, y1 = 0, cr = 0, cb = 0;
#pragma unroll 1
                    for (int lc = 0; lc < 8; ++lc) {
                        if (row < (size_t)P.h && col0 + lc < (size_t)P.w) {
                            const uint8_t *q = img + row * P.stride + (col0 + lc) * (size_t)BPP;
                            const int r = q[0], gg = q[1], bq = q[2];
                            const uint32_t y = (uint32_t)luma_of(r, gg, bq);
                            if (lc < 4) y0 |= y << (8 * lc);
                            else y1 |= y << (8 * (lc - 4));
                            if (lc & 1) {
                                cr |= (uint32_t)cr_of(r, gg, bq) << (8 * (lc >> 1));
                                cb |= (uint32_t)cb_of(r, gg, bq) << (8 * (lc >> 1));
                            }
                        }
                    }
                    ws_w(wl, W_SMP + 2 * lr) = y0;
                    ws_w(wl, W_SMP + 2 * lr + 1) = y1;
                    ws_w(wl, W_SMP + 16 + lr) = cr;
                    ws_w(wl, W_SMP + 24 + lr) = cb;
                }
            }
            // ---- DCT + quantise + zig-zag of the three channels -> int8 coefficients in registers
            uint32_t cz[32]; // luma [0,16), Cr [16,24), Cb [24,32)
            int16_t *co = P.coefs ? P.coefs + gl * 128 : nullptr;
            unsigned widemask = P.force_slow ? 7u : 0u;
            if (dct_luma(wl, wbase, lane, co)) widemask |= 1u;
#pragma unroll
            for (int i = 0; i < 16; ++i) cz[i] = ws_w(wl, W_PAR + i);
#pragma unroll 1
            for (int c = 0; c < 2; ++c)
                if (dct_chroma(wl, wbase, lane, 64 + 32 * c, 32 * c, co ? co + 64 + 32 * c : nullptr)) widemask |= 2u << c;
#pragma unroll
            for (int i = 0; i < 16; ++i) cz[16 + i] = ws_w(wl, W_PAR + i);
            // ---- entropy coding of lum, r, b (reference order JPEG.c:1242, :1284, :1318)
            // The symbol table shares its words with the DCT's row-pass results: clear it once per group; the three
            // channels then tag their entries with epochs 1, 2, 3.
#pragma unroll 8
            for (int i = 0; i < W_CNT - W_LUT; ++i) ws_w(wl, W_LUT + i) = 0;
            BitWriter bw;
            bw.acc = 0;
            bw.nbits = 0;
            bw.dst = stage;
            bw.wpos = 0;
            int bad = 0;
#pragma unroll 1
            for (int ch = 0; ch < 3; ++ch) {
                const int max_bits = ch == 0 ? 1023 : 511;
                int bits = -1;
                if (!((widemask >> ch) & 1u)) bits = entropy_fast(wl, (uint32_t)ch + 1u, cz, ch == 0 ? 0 : 2 + 2 * ch, ch == 0 ? 4 : 2, bw);
                if (bits < 0) {
                    BitWriter tmp = bw;
                    bits = slow_channel(img, P.w, P.h, P.stride, BPP, g, ch, nullptr, &tmp, max_bits, &bad, pgroup);
                    bw = tmp;
                } else if (bits > max_bits) {
                    bad = 1; // char encoded_sequence[1024] / [512] (JPEG.c:1248, :1286)
                }
                if (ch == 0) bl = bits;
                else if (ch == 1) br = bits;
                else bb = bits;
            }
            bw.finish();
            if (bad) {
                atomicAdd((unsigned long long *)&P.result[1], 1ull);
                atomicOr((unsigned long long *)&P.result[2], 2ull);
            }
            const int total_bits = bl + br + bb;
            rec_bytes = (unsigned int)min((total_bits + 7) >> 3, REC_BYTES);
        }
        // ---- warp-wide exclusive scan of record sizes
        unsigned int inc = rec_bytes;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const unsigned int tile_total = __shfl_sync(0xffffffffu, inc, 31);
        const unsigned int my_off = inc - rec_bytes;
        if (have) ljb_lookback_publish(P.status + 1, tile, tile_total, 0);
        // ---- place the PREVIOUS tile in the output stream and move its records there
        if (pend) {
            const unsigned long long base = ljb_lookback_resolve(P.status + 1, p_tile, p_total, 0);
            const bool emit_ok = base + p_total <= P.out_cap;
            if (lane == 0) {
                if (!emit_ok) atomicOr((unsigned long long *)&P.result[2], 1ull);
                if (p_tile == (long long)P.ntiles - 1) {
                    P.result[0] = base + p_total;
                    P.group_offsets[P.ngroups] = base + p_total + P.offs_bias;
                }
            }
            const size_t pgl = (size_t)p_tile * 32 + lane;
            if (pgl < P.ngroups) {
                P.group_offsets[pgl] = base + p_off + P.offs_bias;
                if (P.group_bits) {
                    P.group_bits[3 * pgl + 0] = (uint16_t)(p_bits & 0x3FFu);
                    P.group_bits[3 * pgl + 1] = (uint16_t)((p_bits >> 10) & 0x3FFu);
                    P.group_bits[3 * pgl + 2] = (uint16_t)(p_bits >> 20);
                }
                // head bytes, aligned words, tail bytes
                if (emit_ok && p_bytes) {
                    const uint32_t *src = stage2 + (buf ^ 1) * (REC_WORDS * 32);
                    uint8_t *dst = P.out + base + p_off;
                    const unsigned int head = min((unsigned int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), p_bytes);
                    uint32_t w0 = src[0];
                    for (unsigned int i = 0; i < head; ++i) dst[i] = (uint8_t)(w0 >> (8 * i));
                    const unsigned int nwords = (p_bytes - head) >> 2;
                    uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
                    unsigned int j = 0;
                    for (; j + 4 <= nwords; j += 4) { // four staged words in flight per step (they come from L2)
                        uint32_t wv[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) wv[q] = (j + 1 + q < REC_WORDS) ? src[(j + 1 + q) * 32] : 0u;
                        d4[j] = __funnelshift_r(w0, wv[0], 8 * head);
                        d4[j + 1] = __funnelshift_r(wv[0], wv[1], 8 * head);
                        d4[j + 2] = __funnelshift_r(wv[1], wv[2], 8 * head);
                        d4[j + 3] = __funnelshift_r(wv[2], wv[3], 8 * head);
                        w0 = wv[3];
                    }
                    for (; j < nwords; ++j) {
                        const uint32_t w1 = (j + 1 < REC_WORDS) ? src[(j + 1) * 32] : 0u;
                        d4[j] = __funnelshift_r(w0, w1, 8 * head);
                        w0 = w1;
                    }
                    const unsigned int done = head + 4 * nwords;
                    if (done < p_bytes) {
                        const uint32_t w1 = (j + 1 < REC_WORDS) ? src[(j + 1) * 32] : 0u;
                        const uint32_t t = __funnelshift_r(w0, w1, 8 * head);
                        for (unsigned int i = 0; done + i < p_bytes; ++i) dst[done + i] = (uint8_t)(t >> (8 * i));
                    }
                }
            }
        }
        if (!have) break;
        pend = true;
        p_tile = tile;
        p_off = my_off;
        p_bytes = rec_bytes;
        p_total = tile_total;
        p_bits = (unsigned)bl | ((unsigned)br << 10) | ((unsigned)bb << 20);
        buf ^= 1;
    }
}

} // namespace jpgk

extern "C" size_t ljb_jpeg_group_count(int w, int h)
{
    if (w <= 0 || h <= 0) return 0;
    return ((size_t)w * (size_t)h + 63) / 64; // JPEG.c:1131
}

extern "C" size_t ljb_jpeg_bound(size_t ngroups) { return ngroups * (size_t)jpgk::REC_BYTES + 64; }

struct BatchMode {
    size_t nimages = 0, img_stride = 0; // nimages > 0: groups [0, nimages * group_count(w, h)) of a batch of equal-sized images
    const uint8_t *planar = nullptr;    // samples given per group (128 bytes each): w, h, stride and d_rgba are not used
    int bpp = 4;                        // bytes per pixel: 4 (r g b a) or 3 (r g b)
};
static int jpeg_launch(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                       uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                       uint64_t *d_result, uint64_t offs_bias, const BatchMode &bm = BatchMode())
{
    using namespace jpgk;
    if (!ctx || !d_out || !d_group_offsets || !d_result) return LJB_E_ARG;
    if (bm.planar) { // samples given: the "image" is a column of groups
        w = 8;
        h = 8;
        stride = 32;
    } else if (!d_rgba || w <= 0 || h <= 0 || (w & 1) || (bm.bpp != 3 && bm.bpp != 4) || stride < (size_t)w * (size_t)bm.bpp) {
        return LJB_E_ARG; // odd widths make the reference read past its subsampled rows (JPEG.c:543 with :314)
    }
    const size_t per_image = ljb_jpeg_group_count(w, h);
    const size_t total = bm.planar ? ngroups : (bm.nimages ? bm.nimages * per_image : per_image);
    if (ngroups == 0 || first_group + ngroups > total || total > 0xFFFFFFFFull) return LJB_E_ARG;
    if (bm.nimages && bm.img_stride < stride * (size_t)h) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    const size_t ntiles = (ngroups + 31) / 32;
    if (ntiles > 0x7fffffffull) return LJB_E_ARG;
    const size_t want = (ntiles + NWARPS - 1) / NWARPS;
    const int grid = (int)((want < (size_t)ctx->num_sms) ? want : (size_t)ctx->num_sms);
    int rc;
    if ((rc = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, (size_t)ctx->num_sms * THREADS * REC_BYTES * 2)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_status, &ctx->status_bytes, (ntiles + 2) * sizeof(uint64_t))) != 0) return rc;
    LJB_CUDA(cudaMemsetAsync(ctx->d_status, 0, (ntiles + 2) * sizeof(uint64_t), ctx->stream));
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    Params P;
    P.rgba = d_rgba;
    P.w = w;
    P.h = h;
    P.stride = stride;
    P.first_group = first_group;
    P.ngroups = ngroups;
    P.out = d_out;
    P.out_cap = out_cap;
    P.group_offsets = d_group_offsets;
    P.group_bits = d_group_bits;
    P.coefs = d_coefs;
    P.result = d_result;
    P.status = (uint64_t *)ctx->d_status;
    P.scratch = (uint32_t *)ctx->d_scratch;
    P.ntiles = (uint32_t)ntiles;
    P.offs_bias = offs_bias;
    P.force_slow = getenv("LJB_JPEG_FORCE_SLOW") ? 1 : 0; // test hook
    P.img_groups = bm.nimages ? (uint32_t)per_image : 0u;
    P.img_stride = bm.nimages ? bm.img_stride : 0;
    P.planar = bm.planar;
    P.bpp = bm.bpp;
    if (!(ctx->attr_mask & LJB_ATTR_JPEG)) { // per device (context), not per process
        LJB_CUDA(cudaFuncSetAttribute(jpeg_encode_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        LJB_CUDA(cudaFuncSetAttribute(jpeg_encode_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        ctx->attr_mask |= LJB_ATTR_JPEG;
    }
    ctx->kernel_ms_summed = 0;
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (bm.bpp == 3) jpeg_encode_kernel<3><<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    else jpeg_encode_kernel<4><<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

extern "C" int ljb_jpeg_encode_rgba_dev(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t first_group,
                                        size_t ngroups, uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets,
                                        uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result)
{
    return jpeg_launch(ctx, d_rgba, w, h, stride, first_group, ngroups, d_out, out_cap, d_group_offsets, d_group_bits, d_coefs,
                       d_result, 0);
}

// The same with three bytes per pixel (r g b), as stbi_load(..., 3) and the reference's Pixel rows hold them: a quarter less to upload.
extern "C" int ljb_jpeg_encode_rgb_dev(ljb_ctx *ctx, const uint8_t *d_rgb, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                                       uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                                       uint64_t *d_result)
{
    BatchMode bm;
    bm.bpp = 3;
    return jpeg_launch(ctx, d_rgb, w, h, stride, first_group, ngroups, d_out, out_cap, d_group_offsets, d_group_bits, d_coefs, d_result, 0, bm);
}

// Batch of equal-sized images in ONE launch (one ticket counter, one look-back): image i holds groups
// [i * G, (i + 1) * G) of the numbering, G = ljb_jpeg_group_count(w, h); its stream is out[group_offsets[i*G], group_offsets[(i+1)*G]).
static int jpeg_batch_dev(ljb_ctx *ctx, const uint8_t *d_px, int bpp, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                          uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result)
{
    if (nimages == 0) return LJB_E_ARG;
    BatchMode bm;
    bm.nimages = nimages;
    bm.img_stride = image_stride;
    bm.bpp = bpp;
    return jpeg_launch(ctx, d_px, w, h, stride, 0, nimages * ljb_jpeg_group_count(w, h), d_out, out_cap, d_group_offsets, d_group_bits,
                       d_coefs, d_result, 0, bm);
}
extern "C" int ljb_jpeg_encode_batch_dev(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                                         uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                                         uint64_t *d_result)
{
    return jpeg_batch_dev(ctx, d_rgba, 4, w, h, stride, image_stride, nimages, d_out, out_cap, d_group_offsets, d_group_bits, d_coefs, d_result);
}
extern "C" int ljb_jpeg_encode_batch_rgb_dev(ljb_ctx *ctx, const uint8_t *d_rgb, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                                             uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs,
                                             uint64_t *d_result)
{
    return jpeg_batch_dev(ctx, d_rgb, 3, w, h, stride, image_stride, nimages, d_out, out_cap, d_group_offsets, d_group_bits, d_coefs, d_result);
}

// Groups given by their samples, as the reference's process() receives them (PixelGroup, JPEG.c:42-46: lum_values[64],
// b_values[32], r_values[32] = 128 bytes per group): everything from the DCT on.
extern "C" int ljb_jpeg_encode_groups_dev(ljb_ctx *ctx, const uint8_t *d_samples, size_t ngroups, uint8_t *d_out, size_t out_cap,
                                          uint64_t *d_group_offsets, uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result)
{
    if (!d_samples || (reinterpret_cast<uintptr_t>(d_samples) & 3)) return LJB_E_ARG;
    BatchMode bm;
    bm.planar = d_samples;
    return jpeg_launch(ctx, nullptr, 8, 8, 32, 0, ngroups, d_out, out_cap, d_group_offsets, d_group_bits, d_coefs, d_result, 0, bm);
}

// Host-buffer entry point: bands of whole group rows go through a three-stream pipeline (upload of band k+1 and
// download of band k-1 overlap the kernel of band k), like ljb_lz4_compress.
static int jpeg_encode_host(ljb_ctx *ctx, const uint8_t *rgba, int bpp, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                            uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, int16_t *coefs, size_t *out_len)
{
    if (!ctx || !rgba || !out || w <= 0 || h <= 0 || (w & 1) || stride < (size_t)w * (size_t)bpp) return LJB_E_ARG;
    BatchMode bm;
    bm.bpp = bpp;
    const size_t total = ljb_jpeg_group_count(w, h);
    if (ngroups == 0 || first_group + ngroups > total) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // device layout of a band: image rows packed at a 16-byte aligned stride
    const size_t dstride = ((size_t)w * (size_t)bpp + 15) & ~(size_t)15;
    const size_t bpr = ((size_t)w + 7) / 8;
    const size_t r_begin = first_group / bpr, r_end = (first_group + ngroups + bpr - 1) / bpr; // group rows touched
    // (3/8 of a pipeline chunk = 48 MiB of pixels per band: what is not hidden is the first band's upload and the last one's kernel and
    // download; measured 17.3 / 16.3 / 16.0 / 16.1 / 17.0 ms at 128 / 64 / 48 / 32 / 16 MiB for 16384 x 16384 r g b)
    size_t rows_per_band = ljb_pipe_chunk() * 3 / 8 / (dstride * 8);
    if (rows_per_band == 0) rows_per_band = 1;
    const size_t nbands = (r_end - r_begin + rows_per_band - 1) / rows_per_band;
    const size_t band_groups = rows_per_band * bpr < ngroups + bpr ? rows_per_band * bpr : ngroups + bpr;
    size_t bcap = ljb_jpeg_bound(band_groups);
    if (out_cap < bcap) bcap = out_cap;
    if ((rc = ljb_pipe_init(ctx, nbands)) != 0) return rc;
    for (int i = 0; i < (nbands > 1 ? 2 : 1); ++i) {
        if ((rc = ljb_ensure(&ctx->d_pin[i], &ctx->pin_bytes[i], dstride * 8 * rows_per_band + 64)) != 0) return rc;
        if ((rc = ljb_ensure(&ctx->d_pout[i], &ctx->pout_bytes[i], bcap + 64)) != 0) return rc;
    }
    const size_t offs_words = ngroups + nbands + 4;
    const size_t small = (offs_words + 3 * nbands + 8) * sizeof(uint64_t) + ngroups * 3 * sizeof(uint16_t) + 64 +
                         (coefs ? ngroups * 128 * sizeof(int16_t) : 0);
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, small)) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;              // band k: its groups + 1 entries, at (first group of band - first_group) + k
    uint64_t *d_res = d_offs + offs_words;                    // 3 per band
    int16_t *d_coefs = coefs ? (int16_t *)(d_res + 3 * nbands + 4) : nullptr;
    uint16_t *d_bits = (uint16_t *)((uint8_t *)(d_res + 3 * nbands + 4) + (coefs ? ngroups * 128 * sizeof(int16_t) : 0));
    uint64_t *h_res = ctx->h_res;
    // band k covers group rows [r0, r1) and, of those, the requested groups [g0, g1)
    auto band = [&](size_t k, size_t &r0, size_t &r1, size_t &g0, size_t &g1) {
        r0 = r_begin + k * rows_per_band;
        r1 = r0 + rows_per_band < r_end ? r0 + rows_per_band : r_end;
        g0 = r0 * bpr > first_group ? r0 * bpr : first_group;
        g1 = r1 * bpr < first_group + ngroups ? r1 * bpr : first_group + ngroups;
    };
    auto upload = [&](size_t k, int b) -> cudaError_t {
        size_t r0, r1, g0, g1;
        band(k, r0, r1, g0, g1);
        const size_t y0 = r0 * 8, y1 = r1 * 8 < (size_t)h ? r1 * 8 : (size_t)h;
        if (y1 <= y0) return cudaSuccess;
        return cudaMemcpy2DAsync(ctx->d_pin[b], dstride, rgba + y0 * stride, stride, (size_t)w * (size_t)bpp, y1 - y0, cudaMemcpyHostToDevice,
                                 ctx->s_in);
    };
    size_t running = 0;
    int status = LJB_OK;
    bool unsupported = false;
    cudaError_t e;
#define PIPE(x)                                                                 \
    do {                                                                        \
        e = (x);                                                                \
        if (e != cudaSuccess) {                                                 \
            status = ljb_set_cuda_error(e, #x, __LINE__);                       \
            goto done;                                                          \
        }                                                                       \
    } while (0)
    PIPE(upload(0, 0));
    PIPE(cudaEventRecord(ctx->ev_h2d[0], ctx->s_in));
    for (size_t k = 0; k < nbands; ++k) {
        const int b = (int)(k & 1);
        size_t r0, r1, g0, g1;
        band(k, r0, r1, g0, g1);
        const size_t gi = g0 - first_group; // index of the band's first group inside this call
        PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        if (k >= 2) PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));
        const size_t cap_k = out_cap - running < bcap ? out_cap - running : bcap;
        // the kernel addresses absolute image rows: bias the band buffer so that row r0*8 is its first row
        const uint8_t *biased = (const uint8_t *)ctx->d_pin[b] - r0 * 8 * dstride;
        // `running` is known here (band k-1 has been waited for), so the kernel writes stream-global offsets itself
        rc = jpeg_launch(ctx, biased, w, h, dstride, g0, g1 - g0, (uint8_t *)ctx->d_pout[b], cap_k, d_offs + gi + k, d_bits + 3 * gi,
                         d_coefs ? d_coefs + 128 * gi : nullptr, d_res + 3 * k, running, bm);
        if (rc != 0) {
            status = rc;
            goto done;
        }
        PIPE(cudaMemcpyAsync(h_res + 3 * k, d_res + 3 * k, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        PIPE(cudaEventRecord(ctx->ev_kern[b], ctx->stream));
        if (k + 1 < nbands) {
            if (k >= 1) PIPE(cudaStreamWaitEvent(ctx->s_in, ctx->ev_kern[b ^ 1], 0));
            PIPE(upload(k + 1, b ^ 1));
            PIPE(cudaEventRecord(ctx->ev_h2d[b ^ 1], ctx->s_in));
        }
        PIPE(cudaEventSynchronize(ctx->ev_kern[b]));
        const uint64_t len_k = h_res[3 * k + 0];
        if (h_res[3 * k + 2] & 2) unsupported = true;
        if (h_res[3 * k + 2] & 1) {
            running += (size_t)len_k;
            status = LJB_E_CAPACITY;
            goto done;
        }
        PIPE(cudaStreamWaitEvent(ctx->s_out, ctx->ev_kern[b], 0));
        PIPE(cudaMemcpyAsync(out + running, ctx->d_pout[b], (size_t)len_k, cudaMemcpyDeviceToHost, ctx->s_out));
        if (group_offsets) // consecutive bands overlap in one entry (end of k == start of k+1): the values agree
            PIPE(cudaMemcpyAsync(group_offsets + gi, d_offs + gi + k, (g1 - g0 + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 ctx->s_out));
        PIPE(cudaEventRecord(ctx->ev_d2h[b], ctx->s_out));
        running += (size_t)len_k;
    }
    if (group_bits)
        PIPE(cudaMemcpyAsync(group_bits, d_bits, ngroups * 3 * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->s_out));
    if (coefs) PIPE(cudaMemcpyAsync(coefs, d_coefs, ngroups * 128 * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->s_out));
    PIPE(cudaStreamSynchronize(ctx->s_out));
done:
#undef PIPE
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    if (out_len) *out_len = running;
    if (status == LJB_OK && unsupported) return LJB_E_UNSUPPORTED;
    return status;
}

extern "C" int ljb_jpeg_encode_rgba(ljb_ctx *ctx, const uint8_t *rgba, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                                    uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, int16_t *coefs, size_t *out_len)
{
    return jpeg_encode_host(ctx, rgba, 4, w, h, stride, first_group, ngroups, out, out_cap, group_offsets, group_bits, coefs, out_len);
}
extern "C" int ljb_jpeg_encode_rgb(ljb_ctx *ctx, const uint8_t *rgb, int w, int h, size_t stride, size_t first_group, size_t ngroups,
                                   uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, int16_t *coefs, size_t *out_len)
{
    return jpeg_encode_host(ctx, rgb, 3, w, h, stride, first_group, ngroups, out, out_cap, group_offsets, group_bits, coefs, out_len);
}

// Host-buffer batch: equal-sized images stored one after the other.  Images whose sides are multiples of 8 and that lie back to
// back are, group for group, one tall image (8 | h: no group straddles two images; ceil(w*h/64) == tiles): they go through the
// pipelined single-image path in one call.  Anything else is encoded image by image and the offsets are rebased.
static int jpeg_batch_host(ljb_ctx *ctx, const uint8_t *rgba, int bpp, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                           uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    if (!ctx || !rgba || !out || w <= 0 || h <= 0 || (w & 1) || nimages == 0 || stride < (size_t)w * (size_t)bpp || image_stride < stride * (size_t)h)
        return LJB_E_ARG;
    const size_t G = ljb_jpeg_group_count(w, h);
    if ((w % 8) == 0 && (h % 8) == 0 && image_stride == stride * (size_t)h && (size_t)h * nimages <= 0x7FFFFFFFull)
        return jpeg_encode_host(ctx, rgba, bpp, w, (int)((size_t)h * nimages), stride, 0, G * nimages, out, out_cap, group_offsets, group_bits,
                                nullptr, out_len);
    size_t running = 0;
    std::vector<uint64_t> offs(G + 1);
    for (size_t i = 0; i < nimages; ++i) {
        size_t len = 0;
        const int rc = jpeg_encode_host(ctx, rgba + i * image_stride, bpp, w, h, stride, 0, G, out + running, out_cap - running,
                                        group_offsets ? offs.data() : nullptr, group_bits ? group_bits + 3 * i * G : nullptr, nullptr, &len);
        if (rc != LJB_OK) {
            if (out_len) *out_len = running + len;
            return rc;
        }
        if (group_offsets)
            for (size_t k = 0; k <= G; ++k) group_offsets[i * G + k] = offs[k] + running;
        running += len;
    }
    if (out_len) *out_len = running;
    return LJB_OK;
}

extern "C" int ljb_jpeg_encode_batch(ljb_ctx *ctx, const uint8_t *rgba, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                                     uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    return jpeg_batch_host(ctx, rgba, 4, w, h, stride, image_stride, nimages, out, out_cap, group_offsets, group_bits, out_len);
}
extern "C" int ljb_jpeg_encode_batch_rgb(ljb_ctx *ctx, const uint8_t *rgb, int w, int h, size_t stride, size_t image_stride, size_t nimages,
                                         uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    return jpeg_batch_host(ctx, rgb, 3, w, h, stride, image_stride, nimages, out, out_cap, group_offsets, group_bits, out_len);
}
