// lz4-jpeg_b200/csrc/jfif_encode.cu — true baseline JPEG (JFIF) encoder for sm_100a.
//
// Replaces the only real baseline-JPEG encoder under the reference tree, the vendored stbi_write_jpg_core
// (Algorithms/sequential/JPEG/stb_image_write.h:1398-1605; stbiw__jpg_DCT :1270, stbiw__jpg_processDU :1328,
// stbiw__jpg_writeBits :1250), byte for byte: same header, same float arithmetic in the same order (explicit
// round-to-nearest mul/add, never fused), same Annex-K Huffman codes, one continuous bit stream with 0xFF stuffing.
//
// The sequential program is one bit stream whose every data unit depends on its predecessor twice (DC prediction,
// bit position).  Here it is two kernels on one stream:
//
//   jfif_encode_kernel   persistent; every WARP pulls tiles of R "rounds" from a ticket counter.  A round is 30 data
//                        units = 5 MCUs of 4:2:0 (4 Y + Cb + Cr) or 10 MCUs of 4:4:4; each data unit owns a 64-word block
//                        of the warp's shared memory that holds, in turn, its pixels, its samples and its coefficients.
//                          stage      cp.async of the round's pixels into the blocks (16 B per lane and instruction)
//                          convert    float YCbCr.  4:4:4: the three lanes of an MCU read the same 64 pixels and keep one
//                                     component each, straight in registers, through one branch-free form for the three
//                                     components (no divergence between the lanes of an MCU).  4:2:0: in place, a rolled
//                                     loop over rows; the Y lanes also write the 2x2 chroma means into the chroma lanes' blocks
//                          transform  one data unit per lane: 64 samples in registers, AAN DCT rows/columns, quantise,
//                                     DC difference against the neighbouring lane's DC
//                          halo       before a tile: the DCs of the three data units in front of it, recomputed by the whole
//                                     warp from the 64 pixels of the preceding MCU (shuffle trees in the DCT's own order)
//                          entropy    every lane codes ITS data unit into a lane-private area of the blocks (the bits grow
//                                     from word 0 over the coefficients already read), one warp scan of the bit counts, every
//                                     lane copies its string to its bit offset of the warp's bit buffer (funnel shifts; an
//                                     atomicOr only for the two words it shares with its neighbours)
//                        A finished tile publishes its bit count and its last seven bits and keeps its bits in one of two
//                        bit buffers; one tile later (every predecessor has published by then) the warp resolves the
//                        decoupled look-back, funnel-shifts the buffer to the global bit offset and writes aligned 32-bit
//                        words to the unstuffed stream.  The byte two tiles share is completed by the later one.
//   jfif_stuff_kernel    persistent; 4 KiB chunks of the unstuffed stream: count 0xFF, block scan, look-back for the
//                        output offset, expand in shared memory, coalesced copy-out; the first chunk also writes the
//                        607-byte header (passed by value), the last one the EOI marker and the length.  The host-buffer
//                        entry point launches it once per band of uploaded rows, over the chunks the band's tiles completed,
//                        and sends the finished part of the file down while later bands come up.
//
// Code size matters here: sixteen warps per SM run in different phases, so only the 8x8 transform and the 4:4:4 conversion
// are unrolled (they must be: their 64 values live in registers); 5 900 instructions, of which the out-of-line copies the
// compiler makes for diverged shuffles are never fetched.
//
// Algorithmic bytes: comp*W*H read + N_out written; the unstuffed stream is one extra write + read of ~N_out.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace jfk {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int MAX_ROUNDS_PER_TILE = 8;
constexpr int UNITS_PER_ROUND = 30;
constexpr int CAP_WORDS = 640; // one bit buffer: 2.5 KB; a warp has two
constexpr int CAP_BITS = CAP_WORDS * 32;
constexpr int SPILL_WORDS = CAP_WORDS + 4; // a spilled buffer + its bit count
// data-unit blocks of one round.  4:2:0, per MCU: a 16 x 16 luminance area (row stride 16 words, the four Y units are its
// quadrants) + one 8 x 8 block each for Cb and Cr (row stride 8) + 4 words so that MCUs start in different banks.
// 4:4:4, per MCU: three 8 x 8 blocks (Y, Cb, Cr; the pixels are staged into the first) + 4 words.
constexpr int MCU_WORDS_420 = 256 + 128 + 4, MCU_WORDS_444 = 192 + 4;
constexpr int BLOCK_WORDS = 10 * MCU_WORDS_444;
static_assert(5 * MCU_WORDS_420 <= BLOCK_WORDS, "4:2:0 blocks fit the 4:4:4 area");
constexpr int WARP_WORDS = 2 * CAP_WORDS + BLOCK_WORDS;
// entropy coding: per lane 53 words; a data unit has at most 20 + 63 * 26 = 1658 bits = 52 words
constexpr int AREA_WORDS = 53, COEF_AT = AREA_WORDS - 32;
static_assert(32 * AREA_WORDS <= BLOCK_WORDS, "the lanes' entropy areas fit the blocks");
// word COEF_AT + m (coefficients 2m, 2m+1) is read when the DC and the coefficients up to 2m-1 have been coded: the complete
// words written by then, at most (20 + 26 (2m - 1)) / 32, must end at or below it
constexpr bool area_is_safe()
{
    for (int m = 1; m < 32; ++m)
        if ((20 + 26 * (2 * m - 1)) / 32 > COEF_AT + m) return false;
    return (20 + 26 * 63 + 31) / 32 <= AREA_WORDS;
}
static_assert(area_is_safe(), "the bits of a data unit never overwrite a coefficient that is still to be read");

// table block (32-bit words): built on the host, copied to shared memory by every CTA
constexpr int T_AC_Y = 0;      // 256 x ((code << 8) | len), index run*16 + size
constexpr int T_AC_C = 256;
constexpr int T_DC_Y = 512;    // 16
constexpr int T_DC_C = 528;
constexpr int T_WORDS = 544;
constexpr int S_MULT = T_WORDS;       // 64 floats luminance multipliers, natural order
constexpr int S_MULT_C = S_MULT + 65; // chroma table one bank further, so that mixed warps do not conflict
constexpr int S_FIXED = S_MULT_C + 64 + 3;
static_assert(S_FIXED % 4 == 0 && WARP_WORDS % 4 == 0 && CAP_WORDS % 4 == 0, "16-byte alignment of the blocks");
constexpr int SM_BYTES = (S_FIXED + NWARPS * WARP_WORDS) * 4;
static_assert((SM_BYTES + 1024) * 2 <= 228 * 1024, "two CTAs per SM must fit");

constexpr int HEADER_BYTES = 607;
constexpr int STUFF_THREADS = 256;
constexpr int STUFF_CHUNK = STUFF_THREADS * 16;

// tail word of a tile: bit 8 = published, bits 6..0 = the last seven bits of the tile's bit string
constexpr uint32_t TAIL_VALID = 0x100u;

struct Params {
    const uint8_t *px;
    int w, h, comp;
    size_t stride;
    int fast_ok;        // comp == 4, base and stride 16-byte aligned: cp.async straight from the image
    int mcux;           // MCUs per row
    int rounds_per_tile;
    uint32_t nmcu, nrounds, ntiles;
    uint32_t tile_begin, tile_end; // the tiles of this launch (a band of the image whose pixels are on the device)
    unsigned long long *ticket;    // this launch's ticket counter
    uint8_t *ustream;   // unstuffed entropy-coded bytes
    size_t ucap;
    uint64_t *status;   // [0] ticket, [1 + t] look-back word of tile t (bits)
    uint32_t *tailw;    // [t]
    uint32_t *spill;    // per warp of the grid: spill_slots x SPILL_WORDS words, for tiles denser than the bit buffer
    int spill_slots;
    uint64_t *total_bits;
    uint64_t *result;   // [0] length, [1] tiles that spilled, [2] flags (bit1: scratch capacity exceeded)
    int16_t *coefs;     // optional: 64 per data unit, zig-zag order
    const uint32_t *tables;
    float mult[128];    // [0..63] luminance, [64..127] chrominance
};

struct ZigZag {
    int nat[64]; // zig-zag position -> row-major index (T.81 Figure 5)
};
constexpr ZigZag make_zigzag()
{
    ZigZag z{};
    int r = 0, c = 0;
    for (int k = 0; k < 64; ++k) {
        z.nat[k] = r * 8 + c;
        if (((r + c) & 1) == 0) {
            if (c == 7) ++r;
            else if (r == 0) ++c;
            else { --r; ++c; }
        } else {
            if (r == 7) ++c;
            else if (c == 0) ++r;
            else { ++r; --c; }
        }
    }
    return z;
}
__device__ constexpr ZigZag kZZ = make_zigzag();
__constant__ ZigZag cZZ = make_zigzag(); // the same table for run-time indices

// ---- arithmetic with the reference's rounding: one IEEE operation per C operator, no contraction ---------------
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FM(a, b) __fmul_rn((a), (b))

// stbiw__jpg_DCT (stb_image_write.h:1270-1316): 8-point AAN flow graph
__device__ __forceinline__ void aan8(float &v0, float &v1, float &v2, float &v3, float &v4, float &v5, float &v6, float &v7)
{
    const float s07 = FA(v0, v7), d07 = FS(v0, v7), s16 = FA(v1, v6), d16 = FS(v1, v6);
    const float s25 = FA(v2, v5), d25 = FS(v2, v5), s34 = FA(v3, v4), d34 = FS(v3, v4);
    const float e0 = FA(s07, s34), e3 = FS(s07, s34), e1 = FA(s16, s25), e2 = FS(s16, s25);
    v0 = FA(e0, e1);
    v4 = FS(e0, e1);
    const float r = FM(FA(e2, e3), 0.707106781f);
    v2 = FA(e3, r);
    v6 = FS(e3, r);
    const float o0 = FA(d34, d25), o1 = FA(d25, d16), o2 = FA(d16, d07);
    const float z5 = FM(FS(o0, o2), 0.382683433f);
    const float z2 = FA(FM(o0, 0.541196100f), z5);
    const float z4 = FA(FM(o2, 1.306562965f), z5);
    const float z3 = FM(o1, 0.707106781f);
    const float p = FA(d07, z3), m = FS(d07, z3);
    v5 = FA(m, z2);
    v3 = FS(m, z2);
    v1 = FA(p, z4);
    v7 = FS(p, z4);
}
// (int)(v < 0 ? v - 0.5f : v + 0.5f)  (stb_image_write.h:1353)
__device__ __forceinline__ int quantise(float coef, float mult)
{
    const float v = FM(coef, mult);
    return __float2int_rz(v < 0.f ? FS(v, 0.5f) : FA(v, 0.5f));
}

// stb_image_write.h:1541-1543; p = r | g << 8 | b << 16.  All three components are ((cr*r + cg*g) + cb*b) + off with one IEEE
// operation per C operator: a subtraction in the source is the addition of the exactly negated product, "- 128.f" is
// "+ (-128.f)", and off = -0.0f is the identity for Cb and Cr.  One form for the three, so that lanes working on different
// components of an MCU do not diverge.
// byte k of p as a float: 2^23 + byte assembled with one byte permute, minus 2^23 (exact)
// (measured: the permute form is faster where one lane converts one component — 4:4:4 — and the integer conversion where a lane
// converts all three — 4:2:0)
__device__ __forceinline__ float byte_f(uint32_t p, uint32_t sel) { return FS(__uint_as_float(__byte_perm(p, 0x4B000000u, sel)), 8388608.f); }
__device__ __forceinline__ float to_c(uint32_t p, float cr, float cg, float cb, float off)
{
    const float r = byte_f(p, 0x7540u), g = byte_f(p, 0x7541u), b = byte_f(p, 0x7542u);
    return FA(FA(FA(FM(cr, r), FM(cg, g)), FM(cb, b)), off);
}
__device__ __forceinline__ float to_y(uint32_t p)
{
    const float r = (float)(p & 255u), g = (float)((p >> 8) & 255u), b = (float)((p >> 16) & 255u);
    return FS(FA(FA(FM(0.29900f, r), FM(0.58700f, g)), FM(0.11400f, b)), 128.f);
}
__device__ __forceinline__ float to_u(uint32_t p)
{
    const float r = (float)(p & 255u), g = (float)((p >> 8) & 255u), b = (float)((p >> 16) & 255u);
    return FA(FS(FM(-0.16874f, r), FM(0.33126f, g)), FM(0.50000f, b));
}
__device__ __forceinline__ float to_v(uint32_t p)
{
    const float r = (float)(p & 255u), g = (float)((p >> 8) & 255u), b = (float)((p >> 16) & 255u);
    return FS(FS(FM(0.50000f, r), FM(0.41869f, g)), FM(0.08131f, b));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Stage the pixels of `count` consecutive MCUs starting at MCU m0 into the warp's blocks, 4 pixels (16 B) per chunk, as
// words r | g << 8 | b << 16 (| a << 24).  Chunks that lie inside the image horizontally, of 16-byte aligned RGBA rows,
// are copied with cp.async; all others pixel by pixel, coordinates beyond the image repeating the last row / column
// (stb_image_write.h:1533-1540).
template <bool SUB>
__device__ __noinline__ void stage_pixels(const Params &P, uint32_t *blocks, uint32_t m0, int count)
{
    constexpr int MSZ = SUB ? 16 : 8;  // MCU edge in pixels
    constexpr int CPR = MSZ / 4;       // chunks per MCU row (also: words per row / 4)
    constexpr int CPM = MSZ * CPR;     // chunks per MCU: 64 / 16
    constexpr int MCU_WORDS = SUB ? MCU_WORDS_420 : MCU_WORDS_444;
    const int lane = threadIdx.x & 31;
    int mx = (int)(m0 % (uint32_t)P.mcux), my = (int)(m0 / (uint32_t)P.mcux);
    int cur = 0; // MCU (within this call) that (mx, my) refers to
    const int nchunks = count * CPM;
    for (int c = lane; c < nchunks; c += 32) {
        const int mcu = c / CPM, within = c % CPM;
        while (cur < mcu) {
            ++cur;
            if (++mx == P.mcux) {
                mx = 0;
                ++my;
            }
        }
        const int row = within / CPR, part = within % CPR;
        const int x = mx * MSZ + part * 4, y = my * MSZ + row;
        const int yy = y < P.h ? y : P.h - 1;
        uint32_t *dst = blocks + mcu * MCU_WORDS + within * 4; // row stride = MSZ words in both layouts
        const uint8_t *rowp = P.px + (size_t)yy * P.stride;
        if (P.fast_ok && x + 4 <= P.w) {
            cp_async16(dst, rowp + (size_t)x * 4);
        } else {
            const int og = P.comp > 2 ? 1 : 0, ob = P.comp > 2 ? 2 : 0;
            uint32_t t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = x + i < P.w ? x + i : P.w - 1;
                const uint8_t *q = rowp + (size_t)xx * (size_t)P.comp;
                t[i] = (uint32_t)q[0] | ((uint32_t)q[og] << 8) | ((uint32_t)q[ob] << 16);
            }
            *reinterpret_cast<uint4 *>(dst) = make_uint4(t[0], t[1], t[2], t[3]);
        }
    }
}

__device__ __forceinline__ int bitlen(int v) { return 32 - __clz(v < 0 ? -v : v); }
__device__ __forceinline__ uint32_t extra_bits(int v, int n) { return (uint32_t)(v + (v >> 31)) & ((1u << n) - 1u); }

// OR a left-aligned 64-bit string (hi:lo, bits beyond its length zero) into the big-endian bit buffer at bit offset `off`.
__device__ __forceinline__ void or_left64(uint32_t *buf, uint32_t off, uint32_t hi, uint32_t lo)
{
    const int sh = (int)(off & 31);
    uint32_t *w = buf + (off >> 5);
    const uint32_t w0 = hi >> sh, w1 = __funnelshift_r(lo, hi, sh), w2 = __funnelshift_lc(0u, lo, 32 - sh);
    if (w0) atomicOr(w, w0);
    if (w1) atomicOr(w + 1, w1);
    if (w2) atomicOr(w + 2, w2);
}
// the low `len` (<= 32) bits of `val`
__device__ __forceinline__ void or_bits(uint32_t *buf, uint32_t off, uint32_t val, int len)
{
    or_left64(buf, off, __funnelshift_lc(0u, val, 32 - len), 0u);
}
// the last min(7, nbits) bits of a big-endian bit buffer holding nbits bits
__device__ __forceinline__ uint32_t last_bits(const uint32_t *buf, uint32_t nbits, int &count)
{
    count = nbits < 7 ? (int)nbits : 7;
    if (count == 0) return 0;
    const uint32_t pos = nbits - (uint32_t)count, wi = pos >> 5;
    const uint64_t two = ((uint64_t)buf[wi] << 32) | ((wi + 1) * 32 < nbits ? buf[wi + 1] : 0u);
    return (uint32_t)(two >> (64 - (pos & 31) - count)) & ((1u << count) - 1u);
}

// What a warp remembers about the tile it is writing out (one group normally; several when the tile spilled).
struct TileOut {
    uint64_t g_tile;    // global bit offset of the tile, valid once `resolved`
    uint64_t flushed;   // bits of the tile already placed
    uint32_t last7;     // the last seven bits placed so far
    uint32_t resolved;
};

// Write `nbits` bits of the big-endian bit buffer `src` (shared memory; cleared afterwards) to the unstuffed stream as the
// next group of tile `tile`: decoupled look-back over the per-tile bit counts the first time, then the buffer
// funnel-shifted to the global bit offset and stored as aligned 32-bit words.  The byte two groups share is completed by
// the later one from the earlier one's last bits.  final: the tile's last group (the tile has published its size and its
// last bits itself, as soon as it was finished; here it publishes its inclusive prefix).
__device__ __noinline__ void place_group(const Params &P, uint32_t *src, uint32_t tile, uint32_t nbits, bool final, TileOut &t)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (!t.resolved) {
        uint64_t g = 0;
        if (tile > 0) { // exclusive prefix over the tiles before this one (all claimed by running warps)
            long long at = (long long)tile - 1;
            for (;;) {
                const long long j = at - lane;
                uint64_t wd;
                if (j < 0) {
                    wd = LJB_ST_INC;
                } else {
                    for (unsigned ns = 100; ((wd = ljb_ld_volatile(&P.status[1 + j])) & LJB_ST_MASK) == 0; ns = ns < 1600 ? ns * 2 : ns)
                        __nanosleep(ns);
                }
                const unsigned inc = __ballot_sync(0xffffffffu, (wd & LJB_ST_MASK) == LJB_ST_INC);
                uint64_t v = wd & ~LJB_ST_MASK;
                if (inc) {
                    const int first = __ffs(inc) - 1;
                    if (lane > first) v = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                g += v;
                if (inc) break;
                at -= 32;
            }
        }
        t.g_tile = g;
        t.resolved = 1;
        if ((g & 7) && tile > 0) { // the byte shared with the previous tile: its last bits were published with its size
            uint32_t tw = 0;
            if (lane == 0) {
                for (unsigned ns = 100; ((tw = *(volatile uint32_t *)&P.tailw[tile - 1]) & TAIL_VALID) == 0; ns = ns < 1600 ? ns * 2 : ns)
                    __nanosleep(ns);
            }
            t.last7 = __shfl_sync(0xffffffffu, tw, 0) & 0x7fu;
        }
    }
    const uint64_t G = t.g_tile + t.flushed, E = G + nbits;
    const int k = (int)(G & 7);
    const uint32_t prev_tail = k ? (t.last7 & ((1u << k) - 1u)) << (8 - k) : 0u; // the k bits before G, as the high bits of a byte
    int cnt;
    const uint32_t lb = last_bits(src, nbits, cnt);
    const uint32_t last7 = ((t.last7 << cnt) | lb) & 0x7fu;
    if (final && lane == 0) {
        ljb_st_volatile(&P.status[1 + tile], LJB_ST_INC | E); // walkers stop here
        if (tile + 1 == P.ntiles) *P.total_bits = E;
    }
    const uint64_t blo = G >> 3, bhi = E >> 3; // bytes completed by this group (the first may start in the previous one)
    const bool fits = ((E + 7) >> 3) <= P.ucap;
    if (!fits && lane == 0) atomicOr((unsigned long long *)&P.result[2], 2ull);
    const int sh = (int)(G & 31);
    const uint64_t w0 = G >> 5;
    const int nsrc = (int)((nbits + 31) >> 5);
    const int nw = (int)((sh + nbits + 31) >> 5); // output words touched
    for (int j = lane; j < nw; j += 32) {
        const uint32_t cur = j < nsrc ? src[j] : 0u;
        const uint32_t prev = j > 0 ? src[j - 1] : 0u;
        uint32_t be = sh ? __funnelshift_r(cur, prev, sh) : cur;            // big-endian bit order
        if (j == 0 && k) be |= prev_tail << (24 - 8 * (int)((G >> 3) & 3)); // complete the shared byte
        const uint64_t b0 = (w0 + (uint64_t)j) << 2;                       // first global byte of this word
        if (fits) {
            if (b0 >= blo && b0 + 4 <= bhi) {
                reinterpret_cast<uint32_t *>(P.ustream)[w0 + j] = __byte_perm(be, 0, 0x0123);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint64_t gb = b0 + b;
                    if (gb >= blo && gb < bhi) P.ustream[gb] = (uint8_t)(be >> (24 - 8 * b));
                }
            }
        }
    }
    __syncwarp();
    for (int j = lane; j < nsrc; j += 32) src[j] = 0;
    t.last7 = last7;
    t.flushed += nbits;
    __syncwarp();
}

template <bool SUB>
__global__ void __launch_bounds__(THREADS, 2) jfif_encode_kernel(const __grid_constant__ Params P)
{
    constexpr int DPM = SUB ? 6 : 3;  // data units per MCU
    constexpr int MPR = SUB ? 5 : 10; // MCUs per round
    constexpr int MCU_WORDS = SUB ? MCU_WORDS_420 : MCU_WORDS_444;
    extern __shared__ __align__(16) uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < T_WORDS; i += THREADS) sm[i] = P.tables[i];
    for (int i = threadIdx.x; i < 64; i += THREADS) {
        sm[S_MULT + i] = __float_as_uint(P.mult[i]);
        sm[S_MULT_C + i] = __float_as_uint(P.mult[64 + i]);
    }
    uint32_t *buf = sm + S_FIXED + warp * WARP_WORDS; // the bit buffer being filled
    uint32_t *pend = buf + CAP_WORDS;                 // the other one: the previous tile's bits, until placed
    uint32_t *blocks = buf + 2 * CAP_WORDS;
    for (int i = lane; i < 2 * CAP_WORDS; i += 32) buf[i] = 0;
    __syncthreads();

    const int mi = lane / DPM, d = lane - mi * DPM; // MCU within the round, data unit within the MCU
    const bool lane_used = lane < UNITS_PER_ROUND;
    const bool is_luma = SUB ? d < 4 : d == 0;
    const float *mult = reinterpret_cast<const float *>(sm + (is_luma ? S_MULT : S_MULT_C));
    const int R = P.rounds_per_tile;
    // this lane's data-unit block and its row stride
    const int my_rs = (SUB && d < 4) ? 16 : 8;
    uint32_t *my_blk = blocks + mi * MCU_WORDS + (SUB ? (d < 4 ? (d >> 1) * 128 + (d & 1) * 8 : 256 + (d - 4) * 64) : d * 64);
    // entropy phase: a lane-private area of the blocks (odd stride: no bank conflicts between lanes at the same index): the data
    // unit's bits grow from word 0 while its 64 coefficients (int16, zig-zag order) are read from words COEF_AT.. — a coefficient
    // yields at most 26 bits, the DC 20, so the bits never catch up with the coefficients still to be read
    uint32_t *const area = blocks + lane * AREA_WORDS;
    const uint32_t *const tab_ac = sm + (is_luma ? T_AC_Y : T_AC_C), *const tab_dc = sm + (is_luma ? T_DC_Y : T_DC_C);

    uint32_t pend_tile = 0, pend_bits = 0; // finished tile waiting in `pend` for its offset
    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = P.tile_begin + (uint32_t)atomicAdd(P.ticket, 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const bool have = tile < P.tile_end;
        const uint32_t r0 = tile * (uint32_t)R;
        const uint32_t r1 = !have ? r0 : r0 + (uint32_t)R < P.nrounds ? r0 + (uint32_t)R : P.nrounds;
        int carry_y = 0, carry_u = 0, carry_v = 0; // DC of the last Y / Cb / Cr data unit before the current round
        uint32_t running = 0;                      // bits in the buffer
        auto place_pending = [&]() {
            if (pend_bits) {
                TileOut po = {0, 0, 0, 0};
                place_group(P, pend, pend_tile, pend_bits, true, po);
                pend_bits = 0;
            }
        };
        // The buffer is full in the middle of a tile (denser data than the rounds-per-tile choice expects): move it to this
        // warp's spill area in global memory and go on; the tile is then written out from there when it is finished.
        uint32_t *spill = P.spill + (size_t)(blockIdx.x * NWARPS + warp) * (size_t)P.spill_slots * SPILL_WORDS;
        int nspill = 0;
        auto spill_buffer = [&]() {
            __syncwarp();
            uint32_t *slot = spill + (size_t)nspill * SPILL_WORDS;
            const int used = (int)((running + 31) >> 5);
            for (int j = lane; j < used; j += 32) {
                slot[j] = buf[j];
                buf[j] = 0;
            }
            if (lane == 0) slot[CAP_WORDS] = running;
            ++nspill;
            running = 0;
            __syncwarp();
        };

        // ---- the halo: the DC values of the last Y, Cb and Cr data units before the tile (they predict the tile's first ones),
        // recomputed from the pixels of MCU r0 * MPR - 1.  A DC is output 0 of the row passes, then of the column pass over them:
        // ((v0 + v7) + (v3 + v4)) + ((v1 + v6) + (v2 + v5)) over the rows, the same over the eight row sums.  Lane 4 r + j holds
        // columns j and 7 - j of row r; the three levels of each tree are shuffles (IEEE addition commutes, so both partners get
        // the same sum).  The conversions are the ones of the main path, operation for operation.
        if (have && r0 > 0) {
            const uint32_t mh = r0 * MPR - 1;
            const int hx = (int)(mh % (uint32_t)P.mcux) * (SUB ? 16 : 8), hy = (int)(mh / (uint32_t)P.mcux) * (SUB ? 16 : 8);
            const int og = P.comp > 2 ? 1 : 0, ob = P.comp > 2 ? 2 : 0;
            auto pixel = [&](int x, int y) -> uint32_t { // coordinates beyond the image repeat the last column / row (stb :1533-1540)
                const int xx = x < P.w ? x : P.w - 1, yy = y < P.h ? y : P.h - 1;
                const uint8_t *q = P.px + (size_t)yy * P.stride + (size_t)xx * (size_t)P.comp;
                return (uint32_t)q[0] | ((uint32_t)q[og] << 8) | ((uint32_t)q[ob] << 16);
            };
            auto tree = [&](float a, float b) -> float {
                float t = FA(a, b);
                t = FA(t, __shfl_xor_sync(0xffffffffu, t, 3));
                t = FA(t, __shfl_xor_sync(0xffffffffu, t, 1));
                t = FA(t, __shfl_xor_sync(0xffffffffu, t, 28));
                t = FA(t, __shfl_xor_sync(0xffffffffu, t, 12));
                return FA(t, __shfl_xor_sync(0xffffffffu, t, 4));
            };
            const int hr = lane >> 2, hj = lane & 3;
            const float *mult_y = reinterpret_cast<const float *>(sm + S_MULT), *mult_c = reinterpret_cast<const float *>(sm + S_MULT_C);
            float ya, yb, ua, ub, va, vb;
            if (SUB) { // the last Y unit is the bottom-right one; a chroma sample is the mean of 2 x 2 pixels (stb :1557-1558)
                const uint32_t pa = pixel(hx + 8 + hj, hy + 8 + hr), pb = pixel(hx + 8 + 7 - hj, hy + 8 + hr);
                ya = to_c(pa, 0.29900f, 0.58700f, 0.11400f, -128.f);
                yb = to_c(pb, 0.29900f, 0.58700f, 0.11400f, -128.f);
                float uu[2], vv[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int cx = k ? 7 - hj : hj;
                    const uint32_t tl = pixel(hx + 2 * cx, hy + 2 * hr), tr = pixel(hx + 2 * cx + 1, hy + 2 * hr);
                    const uint32_t bl = pixel(hx + 2 * cx, hy + 2 * hr + 1), br = pixel(hx + 2 * cx + 1, hy + 2 * hr + 1);
                    uu[k] = FM(FA(FA(FA(to_c(tl, -0.16874f, -0.33126f, 0.50000f, -0.f), to_c(tr, -0.16874f, -0.33126f, 0.50000f, -0.f)),
                                     to_c(bl, -0.16874f, -0.33126f, 0.50000f, -0.f)), to_c(br, -0.16874f, -0.33126f, 0.50000f, -0.f)), 0.25f);
                    vv[k] = FM(FA(FA(FA(to_c(tl, 0.50000f, -0.41869f, -0.08131f, -0.f), to_c(tr, 0.50000f, -0.41869f, -0.08131f, -0.f)),
                                     to_c(bl, 0.50000f, -0.41869f, -0.08131f, -0.f)), to_c(br, 0.50000f, -0.41869f, -0.08131f, -0.f)), 0.25f);
                }
                ua = uu[0];
                ub = uu[1];
                va = vv[0];
                vb = vv[1];
            } else {
                const uint32_t pa = pixel(hx + hj, hy + hr), pb = pixel(hx + 7 - hj, hy + hr);
                ya = to_c(pa, 0.29900f, 0.58700f, 0.11400f, -128.f);
                yb = to_c(pb, 0.29900f, 0.58700f, 0.11400f, -128.f);
                ua = to_c(pa, -0.16874f, -0.33126f, 0.50000f, -0.f);
                ub = to_c(pb, -0.16874f, -0.33126f, 0.50000f, -0.f);
                va = to_c(pa, 0.50000f, -0.41869f, -0.08131f, -0.f);
                vb = to_c(pb, 0.50000f, -0.41869f, -0.08131f, -0.f);
            }
            carry_y = quantise(tree(ya, yb), mult_y[0]);
            carry_u = quantise(tree(ua, ub), mult_c[0]);
            carry_v = quantise(tree(va, vb), mult_c[0]);
        }
#pragma unroll 1
        for (long long rr = (long long)r0; rr < (long long)r1; ++rr) {
            const uint32_t m0 = (uint32_t)rr * MPR;
            const uint32_t left = P.nmcu - m0;
            const int nmcus = left < (uint32_t)MPR ? (int)left : MPR;
            const bool valid = lane_used && mi < nmcus;
            const uint32_t m = m0 + (uint32_t)mi;
            {
                // common case: the round's MCUs lie in one MCU row, completely inside an aligned RGBA image: every lane has a
                // fixed (row, 16-byte part) of the MCU and copies it for one MCU after the other
                const int mx0 = (int)(m0 % (uint32_t)P.mcux), my0 = (int)(m0 / (uint32_t)P.mcux);
                constexpr int MSZ = SUB ? 16 : 8;
                if (P.fast_ok && mx0 + nmcus <= P.mcux && (mx0 + nmcus) * MSZ <= P.w) {
                    if (SUB) {
                        const int part = lane & 3;
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            const int row = (lane >> 2) + 8 * hb, y = my0 * 16 + row;
                            const uint8_t *src = P.px + (size_t)(y < P.h ? y : P.h - 1) * P.stride + (size_t)mx0 * 64 + part * 16;
                            uint32_t *dst = blocks + row * 16 + part * 4;
                            for (int i = 0; i < nmcus; ++i) cp_async16(dst + i * MCU_WORDS, src + i * 64);
                        }
                    } else {
                        const int row = (lane & 15) >> 1, part = lane & 1, y = my0 * 8 + row;
                        const uint8_t *src = P.px + (size_t)(y < P.h ? y : P.h - 1) * P.stride + (size_t)mx0 * 32 + part * 16;
                        uint32_t *dst = blocks + row * 8 + part * 4;
                        for (int i = lane >> 4; i < nmcus; i += 2) cp_async16(dst + i * MCU_WORDS, src + i * 32);
                    }
                } else {
                    stage_pixels<SUB>(P, blocks, m0, nmcus);
                }
            }
            cp_async_wait_all();
            __syncwarp();

            // ---- colour conversion, in place ----
            if (SUB) {
                if (valid && d < 4) {
                    float *cu = reinterpret_cast<float *>(blocks + mi * MCU_WORDS + 256 + ((d >> 1) * 4) * 8 + (d & 1) * 4);
#pragma unroll 1
                    for (int yp = 0; yp < 4; ++yp) { // two pixel rows -> 16 Y samples + one row of 4 chroma means each
                        uint4 *row0 = reinterpret_cast<uint4 *>(my_blk + (2 * yp) * 16), *row1 = reinterpret_cast<uint4 *>(my_blk + (2 * yp + 1) * 16);
                        const uint4 a0 = row0[0], a1 = row0[1], b0 = row1[0], b1 = row1[1];
                        const uint32_t t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const uint32_t bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        float y0[8], y1[8], u[4], v[4];
#pragma unroll
                        for (int x = 0; x < 8; ++x) {
                            y0[x] = to_y(t[x]);
                            y1[x] = to_y(bt[x]);
                        }
#pragma unroll
                        for (int x = 0; x < 4; ++x) { // (top-left + top-right + bottom-left + bottom-right) * 0.25f  (stb :1557-1558)
                            u[x] = FM(FA(FA(FA(to_u(t[2 * x]), to_u(t[2 * x + 1])), to_u(bt[2 * x])), to_u(bt[2 * x + 1])), 0.25f);
                            v[x] = FM(FA(FA(FA(to_v(t[2 * x]), to_v(t[2 * x + 1])), to_v(bt[2 * x])), to_v(bt[2 * x + 1])), 0.25f);
                        }
                        reinterpret_cast<float4 *>(row0)[0] = make_float4(y0[0], y0[1], y0[2], y0[3]);
                        reinterpret_cast<float4 *>(row0)[1] = make_float4(y0[4], y0[5], y0[6], y0[7]);
                        reinterpret_cast<float4 *>(row1)[0] = make_float4(y1[0], y1[1], y1[2], y1[3]);
                        reinterpret_cast<float4 *>(row1)[1] = make_float4(y1[4], y1[5], y1[6], y1[7]);
                        *reinterpret_cast<float4 *>(cu + yp * 8) = make_float4(u[0], u[1], u[2], u[3]);
                        *reinterpret_cast<float4 *>(cu + 64 + yp * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
                __syncwarp();
            }

            // ---- the lane's 64 samples to registers ----
            float s[64];
            if (SUB) {
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                    if (valid) {
                        a = *reinterpret_cast<const float4 *>(my_blk + y * my_rs);
                        b = *reinterpret_cast<const float4 *>(my_blk + y * my_rs + 4);
                    }
                    s[y * 8 + 0] = a.x; s[y * 8 + 1] = a.y; s[y * 8 + 2] = a.z; s[y * 8 + 3] = a.w;
                    s[y * 8 + 4] = b.x; s[y * 8 + 5] = b.y; s[y * 8 + 6] = b.z; s[y * 8 + 7] = b.w;
                }
            } else {
                // 4:4:4: the three lanes of an MCU read the same 64 pixels and keep one component each, straight in registers
                const uint4 *src = reinterpret_cast<const uint4 *>(blocks + mi * MCU_WORDS);
                const float cr = d == 0 ? 0.29900f : d == 1 ? -0.16874f : 0.50000f, cg = d == 0 ? 0.58700f : d == 1 ? -0.33126f : -0.41869f;
                const float cb = d == 0 ? 0.11400f : d == 1 ? 0.50000f : -0.08131f, coff = d == 0 ? -128.f : -0.f;
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
                    if (valid) {
                        a0 = src[2 * y];
                        a1 = src[2 * y + 1];
                    }
                    s[y * 8 + 0] = to_c(a0.x, cr, cg, cb, coff); s[y * 8 + 1] = to_c(a0.y, cr, cg, cb, coff);
                    s[y * 8 + 2] = to_c(a0.z, cr, cg, cb, coff); s[y * 8 + 3] = to_c(a0.w, cr, cg, cb, coff);
                    s[y * 8 + 4] = to_c(a1.x, cr, cg, cb, coff); s[y * 8 + 5] = to_c(a1.y, cr, cg, cb, coff);
                    s[y * 8 + 6] = to_c(a1.z, cr, cg, cb, coff); s[y * 8 + 7] = to_c(a1.w, cr, cg, cb, coff);
                }
                __syncwarp(); // every lane holds its samples: the pixel words may be overwritten with coefficients
            }

            // ---- DCT, quantisation, DC difference; 64 int32 back into the block ----
#pragma unroll
            for (int y = 0; y < 8; ++y)
                aan8(s[y * 8], s[y * 8 + 1], s[y * 8 + 2], s[y * 8 + 3], s[y * 8 + 4], s[y * 8 + 5], s[y * 8 + 6], s[y * 8 + 7]);
#pragma unroll
            for (int x = 0; x < 8; ++x) aan8(s[x], s[8 + x], s[16 + x], s[24 + x], s[32 + x], s[40 + x], s[48 + x], s[56 + x]);
            int q[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) q[i] = quantise(s[i], mult[i]);
            const int dc = valid ? q[0] : 0;
            int pred;
            {
                const int back = SUB ? (d == 0 ? 3 : d < 4 ? 1 : 6) : 3;
                const bool from_carry = SUB ? (mi == 0 && (d == 0 || d >= 4)) : mi == 0;
                const int srcl = lane - back;
                const int up = __shfl_sync(0xffffffffu, dc, srcl < 0 ? 0 : srcl);
                const int cy = SUB ? (d == 0 ? carry_y : d == 4 ? carry_u : carry_v) : (d == 0 ? carry_y : d == 1 ? carry_u : carry_v);
                pred = from_carry ? cy : up;
            }
            if (valid) {
                if (P.coefs) {
                    int16_t *gdst = P.coefs + ((size_t)m * DPM + d) * 64;
#pragma unroll
                    for (int kk = 0; kk < 64; ++kk) gdst[kk] = (int16_t)q[kZZ.nat[kk]];
                }
                q[0] = dc - pred; // position 0 carries the DC DIFFERENCE: that is what gets coded
            }
            {
                const int base = (nmcus - 1) * DPM;
                carry_y = __shfl_sync(0xffffffffu, dc, base + (SUB ? 3 : 0));
                carry_u = __shfl_sync(0xffffffffu, dc, base + (SUB ? 4 : 1));
                carry_v = __shfl_sync(0xffffffffu, dc, base + (SUB ? 5 : 2));
            }
            __syncwarp();

            // ---- entropy coding: every lane codes its own data unit into its area, then the strings are concatenated ----
            // (stb_image_write.h:1357-1395: DC category + bits, (run, size) codes + bits, ZRL for runs of 16, EOB)
            uint32_t nbits = 0;
            if (valid) {
#pragma unroll
                for (int m2 = 0; m2 < 32; ++m2)
                    area[COEF_AT + m2] = __byte_perm((uint32_t)q[kZZ.nat[2 * m2]], (uint32_t)q[kZZ.nat[2 * m2 + 1]], 0x5410);
                uint32_t cur = 0;  // the word being filled, left-aligned
                int nacc = 0;      // bits in it
                uint32_t *wp = area;
                auto put = [&](uint32_t code, int len) { // 1 <= len <= 27, code right-aligned
                    const uint32_t v = code << (32 - len);
                    cur |= v >> nacc;
                    const int tot = nacc + len;
                    if (tot >= 32) { // (then nacc > 0)
                        *wp++ = cur;
                        cur = v << (32 - nacc);
                    }
                    nacc = tot & 31;
                };
                {
                    const int n = bitlen(q[0]);
                    const uint32_t e = tab_dc[n];
                    put(((e >> 8) << n) | extra_bits(q[0], n), (int)(e & 255u) + n);
                }
                const uint32_t zrl = tab_ac[0xF0], eob = tab_ac[0];
                int run = 0;
                uint32_t w = area[COEF_AT];
#pragma unroll 2
                for (int k = 1; k < 64; ++k) {
                    if ((k & 1) == 0) w = area[COEF_AT + (k >> 1)];
                    const int c = (k & 1) ? (int)w >> 16 : (int)(w << 16) >> 16;
                    if (c == 0) {
                        ++run;
                        continue;
                    }
                    while (run >= 16) {
                        put(zrl >> 8, (int)(zrl & 255u));
                        run -= 16;
                    }
                    const int n = bitlen(c);
                    const uint32_t e = tab_ac[run * 16 + n];
                    put(((e >> 8) << n) | extra_bits(c, n), (int)(e & 255u) + n);
                    run = 0;
                }
                if (run) put(eob >> 8, (int)(eob & 255u));
                if (nacc) *wp++ = cur;
                nbits = (uint32_t)(wp - area) * 32u - (nacc ? 32u - (uint32_t)nacc : 0u);
            }
            __syncwarp();
            {
                uint32_t incl = nbits;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t tt = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += tt;
                }
                // normally the whole round fits the buffer; otherwise as many leading data units as fit, a spill, and on
                uint32_t done = 0; // bits of the round already copied
                unsigned todo = __ballot_sync(0xffffffffu, nbits != 0);
                while (todo) {
                    const bool fits = nbits != 0 && ((todo >> lane) & 1u) && running + (incl - done) + 7 <= (uint32_t)CAP_BITS;
                    const unsigned now = __ballot_sync(0xffffffffu, fits);
                    if (now == 0) { // not even the next data unit: empty the buffer (a data unit always fits an empty one)
                        spill_buffer();
                        continue;
                    }
                    if (fits) {
                        const uint32_t D = running + (incl - nbits - done);
                        const int sh = (int)(D & 31u);
                        uint32_t *dst = buf + (D >> 5);
                        const int nsrc = (int)((nbits + 31u) >> 5), nout = (int)(((uint32_t)sh + nbits + 31u) >> 5);
                        uint32_t prev = 0;
                        for (int j = 0; j < nout; ++j) {
                            const uint32_t cw = j < nsrc ? area[j] : 0u;
                            const uint32_t o = __funnelshift_r(cw, prev, sh); // big-endian bit order
                            if (j == 0 || j == nout - 1) atomicOr(dst + j, o); // the words shared with the neighbours
                            else dst[j] = o;
                            prev = cw;
                        }
                    }
                    const int last = 31 - __clz(now); // `now` is a run of lanes: the sum up to its last one is in
                    const uint32_t upto = __shfl_sync(0xffffffffu, incl, last);
                    running += upto - done;
                    done = upto;
                    todo &= ~now;
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        // the previous tile: another tile's worth of time has passed, its predecessors have long published
        place_pending();
        if (!have) break;
        if (r1 == P.nrounds) { // padding of the EOI marker: seven one-bits (stb :1586)
            if (lane == 0) or_bits(buf, running, 0x7Fu, 7);
            running += 7;
            __syncwarp();
        }
        if (nspill == 0) { // publish size and last bits now; the offset is fetched after the next tile
            if (lane == 0) {
                int cnt;
                const uint32_t lb = last_bits(buf, running, cnt); // a tile has at least 12 bits: cnt == 7
                *(volatile uint32_t *)&P.tailw[tile] = TAIL_VALID | lb;
                ljb_st_volatile(&P.status[1 + tile], (tile == 0 ? LJB_ST_INC : LJB_ST_AGG) | (uint64_t)running);
            }
            uint32_t *t2 = buf;
            buf = pend;
            pend = t2;
            pend_tile = tile;
            pend_bits = running;
        } else { // a spilled tile: publish, then write it out at once, group by group
            uint64_t total = running;
            uint32_t last7 = 0;
            for (int i = 0; i < nspill; ++i) {
                const uint32_t *slot = spill + (size_t)i * SPILL_WORDS;
                int cnt;
                const uint32_t nb = slot[CAP_WORDS], lb = last_bits(slot, nb, cnt);
                last7 = ((last7 << cnt) | lb) & 0x7fu;
                total += nb;
            }
            {
                int cnt;
                const uint32_t lb = last_bits(buf, running, cnt);
                last7 = ((last7 << cnt) | lb) & 0x7fu;
            }
            if (lane == 0) {
                atomicAdd((unsigned long long *)&P.result[1], 1ull);
                *(volatile uint32_t *)&P.tailw[tile] = TAIL_VALID | last7;
                ljb_st_volatile(&P.status[1 + tile], (tile == 0 ? LJB_ST_INC : LJB_ST_AGG) | total);
            }
            TileOut tout = {0, 0, 0, 0};
            for (int i = 0; i < nspill; ++i) {
                uint32_t *slot = spill + (size_t)i * SPILL_WORDS;
                place_group(P, slot, tile, slot[CAP_WORDS], false, tout);
            }
            place_group(P, buf, tile, running, true, tout);
        }
    }
}

struct Header {
    uint8_t b[HEADER_BYTES + 1];
};

struct StuffParams {
    const uint8_t *ustream;
    size_t ucap;
    const uint64_t *total_bits;
    uint64_t *status; // [0] ticket, [1 + c] look-back word of chunk c (output bytes)
    uint8_t *out;
    size_t out_cap;
    uint64_t *result;
    // One launch per band of the host-buffer entry point, so that the finished part of the file can go down while later bands
    // still come up: a launch stuffs the complete 4 KiB chunks below the bytes the tiles [0, upto_tile) have produced and
    // were not yet below those of [0, prev_tile); the last launch (final) takes what is left, EOI and the length.
    const uint64_t *tile_status; // the encode kernel's look-back words: inclusive bit count of the tiles placed so far
    uint32_t prev_tile, upto_tile;
    int final;
    unsigned long long *ticket; // this launch's chunk counter
    uint64_t *band_len;         // out: [0] file bytes complete after this launch (header included); [-1] the launch before (if prev_tile)
    Header hdr;
};

// Byte stuffing (stb_image_write.h:1258-1262): every 0xFF of the entropy-coded segment is followed by 0x00.
__global__ void __launch_bounds__(STUFF_THREADS) jfif_stuff_kernel(const __grid_constant__ StuffParams P)
{
    __shared__ uint32_t stage[STUFF_CHUNK * 2 / 4 + 2];
    __shared__ uint32_t warp_sum[STUFF_THREADS / 32];
    __shared__ uint64_t s_base;
    __shared__ uint32_t s_chunk;
    uint8_t *stage8 = reinterpret_cast<uint8_t *>(stage);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // whole bytes only: a trailing partial byte is dropped (stb :1586); before the last launch, the bytes of the tiles placed so far
    const uint64_t nbytes = P.final ? *P.total_bits >> 3 : (P.tile_status[P.upto_tile] & ~LJB_ST_MASK) >> 3;
    const uint64_t before_bytes = P.prev_tile ? (P.tile_status[P.prev_tile] & ~LJB_ST_MASK) >> 3 : 0;
    const uint64_t chunk_begin = before_bytes / STUFF_CHUNK; // the earlier launches did the complete chunks below their bytes
    const uint64_t nchunks = P.final ? (nbytes + STUFF_CHUNK - 1) / STUFF_CHUNK : nbytes / STUFF_CHUNK;
    const uint64_t chunk_end = P.final ? (nchunks ? nchunks : 1) : nchunks; // an empty stream still runs chunk 0 (header + EOI)
    if (nbytes > P.ucap) { // the encode kernel ran out of scratch (= the caller's buffer is too small): report a lower bound
        if (blockIdx.x == 0 && tid == 0) {
            if (P.final) P.result[0] = HEADER_BYTES + nbytes + 2;
            atomicOr((unsigned long long *)&P.result[2], 1ull);
            P.band_len[0] = P.prev_tile ? P.band_len[-1] : 0;
        }
        return;
    }
    if (chunk_end <= chunk_begin) { // nothing new is complete
        if (blockIdx.x == 0 && tid == 0) P.band_len[0] = P.prev_tile ? P.band_len[-1] : 0;
        return;
    }
    for (;;) {
        if (tid == 0) s_chunk = (uint32_t)atomicAdd(P.ticket, 1ull);
        __syncthreads();
        const uint64_t chunk = chunk_begin + s_chunk;
        if (chunk >= chunk_end) break;
        const uint64_t at = chunk * STUFF_CHUNK + (uint64_t)tid * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        int nvalid = 0;
        if (at < nbytes) {
            v = *reinterpret_cast<const uint4 *>(P.ustream + at);
            nvalid = nbytes - at >= 16 ? 16 : (int)(nbytes - at);
        }
        uint32_t wds[4] = {v.x, v.y, v.z, v.w};
        int nff = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t eq = __vcmpeq4(wds[i], 0xffffffffu); // 0xff in every byte that is 0xff
            const int nb = nvalid - 4 * i;
            if (nb < 4) eq &= nb <= 0 ? 0u : (0xffffffffu >> (32 - 8 * nb));
            nff += __popc(eq) >> 3;
        }
        // block exclusive scan of the expanded sizes
        const uint32_t mine = (uint32_t)(nvalid + nff);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < STUFF_THREADS / 32; ++i) {
            const uint32_t t = warp_sum[i];
            if (i < warp) before += t;
            total += t;
        }
        if (warp == 0) {
            const uint64_t excl = ljb_lookback(P.status + 1, (long long)chunk, total, 0);
            if (lane == 0) s_base = excl;
        }
        // expand into shared memory
        uint32_t o = before + incl - mine;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < nvalid) {
                const uint8_t b = (uint8_t)(wds[i >> 2] >> (8 * (i & 3)));
                stage8[o++] = b;
                if (b == 0xff) stage8[o++] = 0;
            }
        }
        __syncthreads();
        const uint64_t base = HEADER_BYTES + s_base;
        const bool is_last = P.final && chunk + 1 >= chunk_end;
        const uint64_t end = base + total + (is_last ? 2 : 0);
        if (is_last && tid == 0) {
            stage8[total] = 0xFF; // EOI
            stage8[total + 1] = 0xD9;
            P.result[0] = end;
            if (end > P.out_cap) atomicOr((unsigned long long *)&P.result[2], 1ull);
        }
        if (chunk + 1 == chunk_end && tid == 0) P.band_len[0] = end <= P.out_cap ? end : 0;
        __syncthreads();
        if (end <= P.out_cap) {
            // copy stage8[0, n) to out + base: byte stores up to the first aligned word, then aligned 32-bit stores
            const uint32_t n = total + (is_last ? 2u : 0u);
            uint8_t *dst = P.out + base;
            const uint32_t lead = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
            const uint32_t nlead = lead < n ? lead : n;
            if ((uint32_t)tid < nlead) dst[tid] = stage8[tid];
            const uint32_t nwords = (n - nlead) >> 2;
            for (uint32_t k = tid; k < nwords; k += STUFF_THREADS) {
                const uint32_t so = nlead + 4 * k; // source byte offset, any alignment
                const uint32_t lo = stage[so >> 2], hi = stage[(so >> 2) + 1];
                reinterpret_cast<uint32_t *>(dst + nlead)[k] = __funnelshift_r(lo, hi, 8 * (so & 3));
            }
            const uint32_t done = nlead + 4 * nwords;
            if ((uint32_t)tid < n - done) dst[done + tid] = stage8[done + tid];
            if (chunk == 0) {
                for (int i = tid; i < HEADER_BYTES; i += STUFF_THREADS) P.out[i] = P.hdr.b[i];
            }
        }
        __syncthreads();
    }
}

// ---- host side: ITU-T T.81 Annex K tables, canonical Huffman codes, header -------------------------------------
static const uint8_t kLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                   14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                   18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                   49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                     99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// T.81 Annex C: canonical code words from (BITS, HUFFVAL), packed (code << 8) | length
static void canonical(const uint8_t bits[16], const uint8_t *vals, uint32_t *packed, int nsyms)
{
    for (int i = 0; i < nsyms; ++i) packed[i] = 0;
    unsigned code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k) packed[vals[k]] = (code++ << 8) | (unsigned)len;
        code <<= 1;
    }
}

static void build_tables(uint32_t *t)
{
    memset(t, 0, T_WORDS * sizeof(uint32_t));
    canonical(kAcLumaBits, kAcLumaVals, t + T_AC_Y, 256);
    canonical(kAcChromaBits, kAcChromaVals, t + T_AC_C, 256);
    uint32_t dc[256];
    canonical(kDcLumaBits, kDcVals, dc, 256);
    memcpy(t + T_DC_Y, dc, 16 * sizeof(uint32_t));
    canonical(kDcChromaBits, kDcVals, dc, 256);
    memcpy(t + T_DC_C, dc, 16 * sizeof(uint32_t));
}

struct Plan {
    int subsample;
    uint8_t qy[64], quv[64]; // zig-zag order
    float mult[128];
    Header hdr;
};

static int zz_rank(int natural)
{
    constexpr ZigZag z = make_zigzag();
    for (int k = 0; k < 64; ++k)
        if (z.nat[k] == natural) return k;
    return 0;
}

static void put16(uint8_t *&p, unsigned v)
{
    *p++ = (uint8_t)(v >> 8);
    *p++ = (uint8_t)v;
}

// stb_image_write.h:1466-1519
static void make_plan(int w, int h, int quality, int subsample_mode, Plan *pl)
{
    const int q0 = quality ? quality : 90;
    pl->subsample = subsample_mode < 0 ? (q0 <= 90) : (subsample_mode != 0);
    int q = q0 < 1 ? 1 : q0 > 100 ? 100 : q0;
    q = q < 50 ? 5000 / q : 200 - q * 2;
    for (int i = 0; i < 64; ++i) {
        const int y = (kLumaQ[i] * q + 50) / 100, c = (kChromaQ[i] * q + 50) / 100;
        pl->qy[zz_rank(i)] = (uint8_t)(y < 1 ? 1 : y > 255 ? 255 : y);
        pl->quv[zz_rank(i)] = (uint8_t)(c < 1 ? 1 : c > 255 ? 255 : c);
    }
    static const float aan[8] = {1.0f, 1.387039845f, 1.306562965f, 1.175875602f, 1.0f, 0.785694958f, 0.541196100f, 0.275899379f};
    volatile float sc[8]; // every product rounded to float, as the reference's constant folding does
    for (int i = 0; i < 8; ++i) sc[i] = aan[i] * 2.828427125f;
    for (int r = 0, k = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c, ++k) {
            volatile float dy = (float)pl->qy[zz_rank(k)] * sc[r];
            dy = dy * sc[c];
            volatile float dc = (float)pl->quv[zz_rank(k)] * sc[r];
            dc = dc * sc[c];
            pl->mult[k] = 1.0f / dy;
            pl->mult[64 + k] = 1.0f / dc;
        }
    uint8_t *p = pl->hdr.b;
    put16(p, 0xFFD8);
    put16(p, 0xFFE0);
    put16(p, 16);
    memcpy(p, "JFIF", 5);
    p += 5;
    put16(p, 0x0101);
    *p++ = 0;
    put16(p, 1);
    put16(p, 1);
    put16(p, 0);
    put16(p, 0xFFDB);
    put16(p, 2 + 65 + 65);
    *p++ = 0;
    memcpy(p, pl->qy, 64);
    p += 64;
    *p++ = 1;
    memcpy(p, pl->quv, 64);
    p += 64;
    put16(p, 0xFFC0);
    put16(p, 17);
    *p++ = 8;
    put16(p, (unsigned)h & 0xffff);
    put16(p, (unsigned)w & 0xffff);
    *p++ = 3;
    *p++ = 1; *p++ = (uint8_t)(pl->subsample ? 0x22 : 0x11); *p++ = 0;
    *p++ = 2; *p++ = 0x11; *p++ = 1;
    *p++ = 3; *p++ = 0x11; *p++ = 1;
    put16(p, 0xFFC4);
    put16(p, 2 + 2 * (17 + 12) + 2 * (17 + 162));
    *p++ = 0x00; memcpy(p, kDcLumaBits, 16); p += 16; memcpy(p, kDcVals, 12); p += 12;
    *p++ = 0x10; memcpy(p, kAcLumaBits, 16); p += 16; memcpy(p, kAcLumaVals, 162); p += 162;
    *p++ = 0x01; memcpy(p, kDcChromaBits, 16); p += 16; memcpy(p, kDcVals, 12); p += 12;
    *p++ = 0x11; memcpy(p, kAcChromaBits, 16); p += 16; memcpy(p, kAcChromaVals, 162); p += 162;
    put16(p, 0xFFDA);
    put16(p, 12);
    *p++ = 3;
    *p++ = 1; *p++ = 0x00;
    *p++ = 2; *p++ = 0x11;
    *p++ = 3; *p++ = 0x11;
    *p++ = 0; *p++ = 63; *p++ = 0;
}

} // namespace jfk

static size_t jfif_units(int w, int h, int subsample)
{
    if (subsample) return (size_t)((w + 15) / 16) * (size_t)((h + 15) / 16) * 6;
    return (size_t)((w + 7) / 8) * (size_t)((h + 7) / 8) * 3;
}

extern "C" size_t ljb_jfif_bound(int w, int h)
{
    if (w <= 0 || h <= 0) return 0;
    // 27 bits per coefficient, every byte stuffed; 4:4:4 has the most data units
    return (size_t)jfk::HEADER_BYTES + 2 + jfif_units(w, h, 0) * 216 * 2 + 64;
}

// One image: the encode kernel once per band of pixel rows (band b: rows below band_end[b]; launched after band_ready[b]
// on the context stream), then the stuffing kernel.  nbands == 1 with band_ready == nullptr is the device-resident call.
static int jfif_run(ljb_ctx *ctx, const uint8_t *d_pixels, int w, int h, int comp, size_t stride, int quality, int subsample,
                    uint8_t *d_out, size_t out_cap, uint64_t *d_result, int16_t *d_coefs, int nbands, const int *band_end,
                    const cudaEvent_t *band_ready, uint64_t *band_len_host = nullptr, const cudaEvent_t *band_done = nullptr)
{
    using namespace jfk;
    if (!ctx || !d_pixels || !d_out || !d_result || w <= 0 || h <= 0 || comp < 1 || comp > 4 || stride < (size_t)w * (size_t)comp ||
        subsample < -1 || subsample > 1 || out_cap < (size_t)HEADER_BYTES + 2)
        return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    Plan pl;
    make_plan(w, h, quality, subsample, &pl);
    const int msz = pl.subsample ? 16 : 8;
    const int mcux = (w + msz - 1) / msz, mcuy = (h + msz - 1) / msz;
    const uint64_t nmcu64 = (uint64_t)mcux * (uint64_t)mcuy;
    if (nmcu64 > 0x7fffffffull) return LJB_E_ARG;
    const uint32_t nmcu = (uint32_t)nmcu64;
    const uint32_t mpr = pl.subsample ? 5 : 10;
    const uint32_t nrounds = (nmcu + mpr - 1) / mpr;
    // rounds per tile: as many as fit the 2.5 KB bit buffer on uniform noise (the densest realistic input) at this quality;
    // denser tiles spill their bit buffer to global memory and are written out from there: correct, only slower
    const int q0 = quality ? quality : 90;
    int R = q0 <= 50 ? 4 : q0 <= 80 ? 3 : q0 <= 90 ? 2 : 1;
    if (const char *e = getenv("LJB_JFIF_ROUNDS")) { // test / tuning hook
        const int v = atoi(e);
        if (v >= 1 && v <= MAX_ROUNDS_PER_TILE) R = v;
    }
    const uint32_t ntiles = (nrounds + (uint32_t)R - 1) / (uint32_t)R;
    // unstuffed stream: never longer than the worst case, nor (usefully) than the caller's output buffer
    size_t ucap = jfif_units(w, h, pl.subsample) * 216 + 8;
    if (ucap > out_cap) ucap = out_cap;
    const size_t nchunks_max = (ucap + STUFF_CHUNK - 1) / STUFF_CHUNK + 1;
    int rc;
    if ((rc = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, ucap + STUFF_CHUNK + 64)) != 0) return rc;
    // status block: tables | total_bits | ticket + tile status | ticket + chunk status | tile tail words
    const size_t o_tab = 0;
    const size_t o_total = o_tab + T_WORDS * 4;
    const size_t o_st1 = o_total + 8;
    const size_t o_st2 = o_st1 + ((size_t)ntiles + 1) * 8;
    const size_t o_tailw = o_st2 + (nchunks_max + 1) * 8;
    const size_t o_tick = (o_tailw + (size_t)ntiles * 4 + 7) & ~(size_t)7; // per band: ticket of the encode launch, of the stuff launch, file bytes
    const size_t o_end = o_tick + (size_t)(nbands + 1) * 3 * 8;
    // spill area: a tile is at most R rounds of 30 data units of 216 bytes
    const int spill_slots = (R * UNITS_PER_ROUND * 216 * 8) / (CAP_BITS - UNITS_PER_ROUND * 216 * 8 / 15) + 2;
    const size_t o_spill = (o_end + 15) & ~(size_t)15;
    const size_t spill_bytes = (size_t)ctx->num_sms * 2 * NWARPS * (size_t)spill_slots * SPILL_WORDS * 4;
    if ((rc = ljb_ensure(&ctx->d_status, &ctx->status_bytes, o_spill + spill_bytes)) != 0) return rc;
    uint8_t *sb = (uint8_t *)ctx->d_status;
    uint32_t tables[T_WORDS];
    build_tables(tables);
    LJB_CUDA(cudaMemsetAsync(sb + o_total, 0, o_end - o_total, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(sb + o_tab, tables, sizeof tables, cudaMemcpyHostToDevice, ctx->stream));
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));

    Params P;
    memset(&P, 0, sizeof P);
    P.px = d_pixels;
    P.w = w;
    P.h = h;
    P.comp = comp;
    P.stride = stride;
    P.fast_ok = comp == 4 && ((uintptr_t)d_pixels & 15) == 0 && (stride & 15) == 0;
    P.mcux = mcux;
    P.nmcu = nmcu;
    P.nrounds = nrounds;
    P.ntiles = ntiles;
    P.ustream = (uint8_t *)ctx->d_scratch;
    P.ucap = ucap;
    P.status = (uint64_t *)(sb + o_st1);
    P.tailw = (uint32_t *)(sb + o_tailw);
    P.rounds_per_tile = R;
    P.spill = (uint32_t *)(sb + o_spill);
    P.spill_slots = spill_slots;
    P.total_bits = (uint64_t *)(sb + o_total);
    P.result = d_result;
    P.coefs = d_coefs;
    P.tables = (const uint32_t *)(sb + o_tab);
    memcpy(P.mult, pl.mult, sizeof P.mult);

    if (!(ctx->attr_mask & LJB_ATTR_JFIF)) { // per device (context), not per process
        LJB_CUDA(cudaFuncSetAttribute(jfif_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES));
        LJB_CUDA(cudaFuncSetAttribute(jfif_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES));
        ctx->attr_mask |= LJB_ATTR_JFIF;
    }
    ctx->kernel_ms_summed = 0;
    const size_t full = (size_t)ctx->num_sms * 2;
    StuffParams S;
    S.ustream = P.ustream;
    S.ucap = ucap;
    S.total_bits = P.total_bits;
    S.status = (uint64_t *)(sb + o_st2);
    S.out = d_out;
    S.out_cap = out_cap;
    S.result = d_result;
    S.tile_status = P.status; // word t = look-back word of tile t - 1 (its inclusive bit count once the tile is placed)
    S.hdr = pl.hdr;
    const size_t sfull = (size_t)ctx->num_sms * 8;
    const int sgrid = (int)(nchunks_max < sfull ? nchunks_max : sfull);
    uint32_t t_begin = 0;
    for (int b = 0; b < nbands; ++b) {
        // tiles whose MCUs (and the one before the first: the DC halo) lie in the rows that are on the device by now
        uint32_t t_end = ntiles;
        if (b + 1 < nbands) {
            const uint64_t mcus = (uint64_t)(band_end[b] / msz) * (uint64_t)mcux;
            const uint64_t tiles = mcus / mpr / (uint64_t)R;
            t_end = (uint32_t)(tiles < ntiles ? tiles : ntiles);
        }
        if (band_ready) LJB_CUDA(cudaStreamWaitEvent(ctx->stream, band_ready[b], 0));
        if (b == 0) LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        if (t_end > t_begin) {
            P.tile_begin = t_begin;
            P.tile_end = t_end;
            P.ticket = (unsigned long long *)(sb + o_tick) + b;
            const size_t want = ((size_t)(t_end - t_begin) + NWARPS - 1) / NWARPS;
            const int grid = (int)(want < full ? want : full);
            if (pl.subsample) jfif_encode_kernel<true><<<grid, THREADS, SM_BYTES, ctx->stream>>>(P);
            else jfif_encode_kernel<false><<<grid, THREADS, SM_BYTES, ctx->stream>>>(P);
            LJB_CUDA(cudaGetLastError());
            ctx->launches += 1;
        }
        if (b + 1 == nbands) LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        // stuffing of what this band completed (everything left, in the last band)
        const bool final = b + 1 == nbands;
        if (final || t_end > t_begin || band_len_host) {
            S.prev_tile = t_begin;
            S.upto_tile = t_end;
            S.final = final ? 1 : 0;
            S.ticket = (unsigned long long *)(sb + o_tick) + (nbands + 1) + b;
            S.band_len = (uint64_t *)(sb + o_tick) + 2 * (nbands + 1) + 1 + b; // ([-1] of band 0 is never read)
            jfif_stuff_kernel<<<sgrid, STUFF_THREADS, 0, ctx->stream>>>(S);
            LJB_CUDA(cudaGetLastError());
            ctx->launches += 1;
            if (band_len_host) {
                LJB_CUDA(cudaMemcpyAsync(band_len_host + b, S.band_len, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
                LJB_CUDA(cudaEventRecord(band_done[b], ctx->stream));
            }
        }
        t_begin = t_end;
    }
    return LJB_OK;
}

extern "C" int ljb_jfif_encode_dev(ljb_ctx *ctx, const uint8_t *d_pixels, int w, int h, int comp, size_t stride, int quality,
                                   int subsample, uint8_t *d_out, size_t out_cap, uint64_t *d_result, int16_t *d_coefs)
{
    if (!ctx || !d_pixels || !d_out || !d_result || w <= 0 || h <= 0 || comp < 1 || comp > 4 || stride < (size_t)w * (size_t)comp ||
        subsample < -1 || subsample > 1 || out_cap < (size_t)jfk::HEADER_BYTES + 2)
        return LJB_E_ARG;
    return jfif_run(ctx, d_pixels, w, h, comp, stride, quality, subsample, d_out, out_cap, d_result, d_coefs, 1, nullptr, nullptr);
}

// Host-buffer entry point.  The image is one bit stream, but its tiles only look BACK (bit offsets, the DC of the data unit
// before the tile): the rows are uploaded in bands on a second stream and the tiles of a band are encoded as soon as the
// band has arrived, so the encode kernel hides behind the upload; every band is followed by a stuffing launch over the 4 KiB
// chunks its tiles completed, and that part of the file goes down on the third stream while later bands still come up.
extern "C" int ljb_jfif_encode(ljb_ctx *ctx, const uint8_t *pixels, int w, int h, int comp, size_t stride, int quality, int subsample,
                               uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!ctx || !pixels || !out || !out_len || w <= 0 || h <= 0 || comp < 1 || comp > 4 || stride < (size_t)w * (size_t)comp ||
        subsample < -1 || subsample > 1)
        return LJB_E_ARG;
    if (out_cap < (size_t)jfk::HEADER_BYTES + 2) return LJB_E_CAPACITY;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t rowbytes = (size_t)w * (size_t)comp;
    const size_t dstride = (rowbytes + 15) & ~(size_t)15;
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], dstride * (size_t)h + 64)) != 0) return rc;
    size_t dcap = out_cap;
    const size_t bound = ljb_jfif_bound(w, h);
    if (dcap > bound) dcap = bound;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], dcap + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, 64)) != 0) return rc;
    // bands of whole 16-row MCU rows, 3/8 of a pipeline chunk of pixels each (48 MiB: what follows the last upload — its tiles, their
    // stuffing and download — is a band's worth; measured 18.1 / 16.7 / 16.3 / 16.4 ms at 128 / 64 / 48 / 32 MiB for 16384 x 16384 r g b),
    // at most MAXBANDS
    constexpr int MAXBANDS = 24;
    int band_end[MAXBANDS];
    cudaEvent_t ready[2 * MAXBANDS]; // per band: its rows are on the device | its part of the file is complete
    int nbands = 0;
    {
        size_t rows = ljb_pipe_chunk() * 3 / 8 / (dstride ? dstride : 1);
        rows = (rows + 15) & ~(size_t)15;
        if (rows < 16) rows = 16;
        if (rows * MAXBANDS < (size_t)h) rows = (((size_t)h + MAXBANDS - 1) / MAXBANDS + 15) & ~(size_t)15;
        for (size_t y = 0; y < (size_t)h; y += rows) band_end[nbands++] = (int)(y + rows < (size_t)h ? y + rows : (size_t)h);
    }
    int made = 0;
    int status = LJB_OK;
    if ((rc = ljb_pipe_init(ctx, (size_t)nbands)) != 0) return rc;
    for (; made < 2 * nbands; ++made)
        if (cudaEventCreateWithFlags(&ready[made], cudaEventDisableTiming) != cudaSuccess) break;
    if (made < 2 * nbands) status = ljb_set_cuda_error(cudaGetLastError(), "cudaEventCreateWithFlags", __LINE__);
    const cudaEvent_t *done = ready + nbands;
    uint64_t *band_len = ctx->h_res; // pinned: the file bytes complete after each band
    if (status == LJB_OK) {
        // the upload stream must not overtake work still reading the buffer from an earlier call on the context stream
        cudaError_t e = cudaEventRecord(ctx->ev_kern[0], ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_in, ctx->ev_kern[0], 0);
        for (int b = 0, y0 = 0; b < nbands && e == cudaSuccess; y0 = band_end[b], ++b) {
            e = cudaMemcpy2DAsync((uint8_t *)ctx->d_pin[0] + (size_t)y0 * dstride, dstride, pixels + (size_t)y0 * stride, stride, rowbytes,
                                  (size_t)(band_end[b] - y0), cudaMemcpyHostToDevice, ctx->s_in);
            if (e == cudaSuccess) e = cudaEventRecord(ready[b], ctx->s_in);
        }
        if (e != cudaSuccess) status = ljb_set_cuda_error(e, "band upload", __LINE__);
    }
    if (status == LJB_OK)
        status = jfif_run(ctx, (const uint8_t *)ctx->d_pin[0], w, h, comp, dstride, quality, subsample, (uint8_t *)ctx->d_pout[0], dcap,
                          (uint64_t *)ctx->d_small, nullptr, nbands, band_end, ready, band_len, done);
    uint64_t res[3] = {0, 0, 0};
    // the part of the file every band completed goes down on the third stream while later bands come up and are encoded
    size_t sent = 0;
    if (status == LJB_OK) {
        cudaError_t e = cudaSuccess;
        for (int b = 0; b < nbands && e == cudaSuccess; ++b) {
            e = cudaEventSynchronize(done[b]);
            const size_t upto = e == cudaSuccess ? (size_t)band_len[b] : 0;
            if (e == cudaSuccess && upto > sent && upto <= out_cap) {
                e = cudaMemcpyAsync(out + sent, (const uint8_t *)ctx->d_pout[0] + sent, upto - sent, cudaMemcpyDeviceToHost, ctx->s_out);
                sent = upto;
            }
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(res, ctx->d_small, sizeof res, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_out);
        if (e != cudaSuccess) status = ljb_set_cuda_error(e, "file download", __LINE__);
    }
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    for (int i = 0; i < made; ++i) cudaEventDestroy(ready[i]);
    if (status != LJB_OK) return status;
    *out_len = (size_t)res[0];
    if ((res[2] & 3) || sent != (size_t)res[0]) return LJB_E_CAPACITY;
    return LJB_OK;
}
