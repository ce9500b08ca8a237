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
//                        units = 5 MCUs of 4:2:0 (4 Y + Cb + Cr) or 10 MCUs of 4:4:4.
//                          transform, one data unit per lane: pixels (prefetched into registers during the previous
//                            round's entropy phase) -> float YCbCr (the Y lanes of 4:2:0 also produce the 2x2 chroma
//                            means for the chroma lanes of their MCU, through shared memory) -> AAN DCT rows/columns in
//                            registers -> quantise -> DC difference against the neighbouring lane's DC (the data unit
//                            before the tile is recomputed, DC only) -> 64 int16 in zig-zag order to shared memory;
//                          entropy, one data unit per iteration, TWO coefficients per lane: zero runs from ballots,
//                            (run, size) code + extra bits per lane, warp scan of the lengths, every lane ORs its bits at
//                            its exact offset into the warp's bit buffer in shared memory.  No divergence, no unrolling.
//                        At the end of the tile a decoupled look-back over per-tile bit counts gives the global bit
//                        offset; the buffer is funnel-shifted to it and written to the unstuffed stream as aligned
//                        32-bit words.  The byte two tiles share is completed by the later one (the earlier tile
//                        publishes its trailing bits).
//   jfif_stuff_kernel    persistent; 4 KiB chunks of the unstuffed stream: count 0xFF, block scan, look-back for the
//                        output offset, expand in shared memory, coalesced copy-out; the first chunk also writes the
//                        607-byte header (passed by value), the last one the EOI marker and the length.
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
constexpr int CAP_WORDS = 1024; // per-warp bit buffer: 4 KB
constexpr int CAP_BITS = CAP_WORDS * 32;
constexpr int UNIT_MAX_BITS = 64 * 27; // 16-bit code + 11 extra bits per coefficient
constexpr int COEF_STRIDE = 33;       // 32 words (64 int16) per data unit + 1: lane j writes unit j conflict-free
constexpr int COEF_WORDS = UNITS_PER_ROUND * COEF_STRIDE + 2;
constexpr int CHROMA_SLOT = 65;       // 64 floats + 1: the chroma lanes read conflict-free
static_assert(10 * CHROMA_SLOT <= COEF_WORDS, "the chroma exchange aliases the coefficient buffer");
// pixel staging (cp.async): 4:2:0 5 MCUs x 16 rows x 64 B, 4:4:4 10 MCUs x 8 rows x 32 B; each MCU padded by 16 B so
// that the 128-bit reads of the lanes spread over the banks
constexpr int PIX_MCU_WORDS_420 = 16 * 16 + 4, PIX_MCU_WORDS_444 = 8 * 8 + 4;
constexpr int PIX_WORDS = 5 * PIX_MCU_WORDS_420;
static_assert(10 * PIX_MCU_WORDS_444 <= PIX_WORDS, "4:4:4 staging fits the 4:2:0 buffer");
constexpr int WARP_WORDS = CAP_WORDS + COEF_WORDS + PIX_WORDS;

// table block (32-bit words): built on the host, copied to shared memory by every CTA
constexpr int T_AC_Y = 0;      // 256 x ((code << 8) | len), index run*16 + size
constexpr int T_AC_C = 256;
constexpr int T_DC_Y = 512;    // 16
constexpr int T_DC_C = 528;
constexpr int T_WORDS = 544;
constexpr int S_MULT = T_WORDS;       // 64 floats luminance multipliers, natural order
constexpr int S_MULT_C = S_MULT + 65; // chroma table one bank further, so that mixed warps do not conflict
constexpr int S_FIXED = S_MULT_C + 64 + 3;
constexpr int SM_BYTES = (S_FIXED + NWARPS * WARP_WORDS) * 4;
static_assert(SM_BYTES * 2 <= 227 * 1024, "two CTAs per SM must fit");

constexpr int HEADER_BYTES = 607;
constexpr int STUFF_THREADS = 256;
constexpr int STUFF_CHUNK = STUFF_THREADS * 16;

// tail word of a tile: bit 8 = published, bits 7..0 = its last, partial byte (high bits valid, rest zero)
constexpr uint32_t TAIL_VALID = 0x100u;

struct Params {
    const uint8_t *px;
    int w, h, comp;
    size_t stride;
    int fast_ok;        // comp == 4, base and stride 16-byte aligned: interior units use two 128-bit loads per row
    int mcux;           // MCUs per row
    int rounds_per_tile;
    uint32_t nmcu, nrounds, ntiles;
    uint8_t *ustream;   // unstuffed entropy-coded bytes
    size_t ucap;
    uint64_t *status;   // [0] ticket, [1 + t] look-back word of tile t (bits)
    uint32_t *tailw;    // [t]
    uint64_t *total_bits;
    uint64_t *result;   // [0] length, [1] tiles that had to flush early, [2] flags (bit1: scratch capacity exceeded)
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
// output 0 of the same graph
__device__ __forceinline__ float aan8_dc(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7)
{
    return FA(FA(FA(v0, v7), FA(v3, v4)), FA(FA(v1, v6), FA(v2, v5)));
}
// (int)(v < 0 ? v - 0.5f : v + 0.5f)  (stb_image_write.h:1353)
__device__ __forceinline__ int quantise(float coef, float mult)
{
    const float v = FM(coef, mult);
    return __float2int_rz(v < 0.f ? FS(v, 0.5f) : FA(v, 0.5f));
}

// stb_image_write.h:1541-1543
__device__ __forceinline__ float to_y(float r, float g, float b) { return FS(FA(FA(FM(0.29900f, r), FM(0.58700f, g)), FM(0.11400f, b)), 128.f); }
__device__ __forceinline__ float to_u(float r, float g, float b) { return FA(FS(FM(-0.16874f, r), FM(0.33126f, g)), FM(0.50000f, b)); }
__device__ __forceinline__ float to_v(float r, float g, float b) { return FS(FS(FM(0.50000f, r), FM(0.41869f, g)), FM(0.08131f, b)); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Stage the pixels of `count` consecutive MCUs starting at MCU m0 into the warp's pixel buffer, 4 pixels (16 B) per
// chunk, as words r | g << 8 | b << 16 (| a << 24).  MCUs that lie inside the image horizontally, of 16-byte aligned RGBA
// rows, are copied with cp.async; all others pixel by pixel, coordinates beyond the image repeating the last row /
// column (stb_image_write.h:1533-1540).
template <bool SUB>
__device__ __forceinline__ void stage_pixels(const Params &P, uint32_t *pix, uint32_t m0, int count, int lane)
{
    constexpr int MSZ = SUB ? 16 : 8;                 // MCU edge in pixels
    constexpr int CPR = MSZ / 4;                      // chunks per MCU row
    constexpr int CPM = MSZ * CPR;                    // chunks per MCU: 64 / 16
    constexpr int MCU_WORDS = SUB ? PIX_MCU_WORDS_420 : PIX_MCU_WORDS_444;
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
        uint32_t *dst = pix + mcu * MCU_WORDS + within * 4;
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

// OR the low `len` bits of `val` (len <= 64 - 31) into the big-endian bit buffer at bit offset `off`.
__device__ __forceinline__ void or_bits(uint32_t *buf, uint32_t off, uint64_t val, int len)
{
    if (len == 0) return;
    const uint64_t x = val << (64 - len); // left aligned
    const int sh = (int)(off & 31);
    uint32_t *w = buf + (off >> 5);
    const uint32_t w0 = (uint32_t)(x >> (32 + sh)), w1 = (uint32_t)(x >> sh);
    const uint32_t w2 = (uint32_t)(((x & 0xffffffffull) << 32) >> sh);
    if (w0) atomicOr(w, w0);
    if (w1) atomicOr(w + 1, w1);
    if (w2) atomicOr(w + 2, w2);
}

template <bool SUB>
__global__ void __launch_bounds__(THREADS, 2) jfif_encode_kernel(const __grid_constant__ Params P)
{
    constexpr int DPM = SUB ? 6 : 3;  // data units per MCU
    constexpr int MPR = SUB ? 5 : 10; // MCUs per round
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < T_WORDS; i += THREADS) sm[i] = P.tables[i];
    for (int i = threadIdx.x; i < 64; i += THREADS) {
        sm[S_MULT + i] = __float_as_uint(P.mult[i]);
        sm[S_MULT_C + i] = __float_as_uint(P.mult[64 + i]);
    }
    uint32_t *buf = sm + S_FIXED + warp * WARP_WORDS;
    uint32_t *coef = buf + CAP_WORDS;
    uint32_t *pix = coef + COEF_WORDS;
    float *chroma = reinterpret_cast<float *>(coef); // 4:2:0 exchange of the 2x2 means; dead before the coefficients are stored
    for (int i = lane; i < CAP_WORDS; i += 32) buf[i] = 0;
    __syncthreads();

    const int mi = lane / DPM, d = lane - mi * DPM; // MCU within the round, data unit within the MCU
    const bool lane_used = lane < UNITS_PER_ROUND;
    const bool is_luma = SUB ? d < 4 : d == 0;
    const float *mult = reinterpret_cast<const float *>(sm + (is_luma ? S_MULT : S_MULT_C));
    const uint32_t below = (1u << lane) - 1u;
    const int R = P.rounds_per_tile;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = (uint32_t)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= P.ntiles) break;
        const uint32_t r0 = tile * (uint32_t)R;
        const uint32_t r1 = r0 + (uint32_t)R < P.nrounds ? r0 + (uint32_t)R : P.nrounds;
        int carry_y = 0, carry_u = 0, carry_v = 0; // DC of the last Y / Cb / Cr data unit before the current round
        uint32_t running = 0;                      // bits in the buffer
        bool resolved = false;                     // global bit offset of the tile known (an early flush happened)
        uint64_t g_tile = 0, flushed = 0;          // bit offset of the tile; bits of the tile already written
        uint32_t prev_tail = 0;                    // partial byte left by the previous group (high bits)

        // Write the buffer to the unstuffed stream.  final: this is the tile's last group.
        auto flush = [&](bool final) {
            __syncwarp();
            if (!resolved) {
                if (final && lane == 0 && tile > 0) ljb_st_volatile(&P.status[1 + tile], LJB_ST_AGG | (uint64_t)running);
                if (tile > 0) { // exclusive prefix over the tiles before this one (all claimed by running warps)
                    long long at = (long long)tile - 1;
                    for (;;) {
                        const long long j = at - lane;
                        uint64_t wd;
                        if (j < 0) {
                            wd = LJB_ST_INC;
                        } else {
                            while (((wd = ljb_ld_volatile(&P.status[1 + j])) & LJB_ST_MASK) == 0) __nanosleep(200);
                        }
                        const unsigned inc = __ballot_sync(0xffffffffu, (wd & LJB_ST_MASK) == LJB_ST_INC);
                        uint64_t v = wd & ~LJB_ST_MASK;
                        if (inc) {
                            const int first = __ffs(inc) - 1;
                            if (lane > first) v = 0;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        g_tile += v;
                        if (inc) break;
                        at -= 32;
                    }
                }
                resolved = true;
                if (!final && lane == 0) atomicAdd((unsigned long long *)&P.result[1], 1ull);
            }
            const uint64_t G = g_tile + flushed, E = G + running;
            if (final && lane == 0) {
                ljb_st_volatile(&P.status[1 + tile], LJB_ST_INC | E);
                if (tile + 1 == P.ntiles) *P.total_bits = E;
                // The tile's trailing partial byte depends only on its own bits and offset (a tile has at least 12 bits):
                // publish it before waiting for the predecessor's, so that tiles never wait in a chain.
                uint32_t tb = 0;
                const int k = (int)(E & 7);
                if (k && running >= (uint32_t)k) { // the last k bits of the buffer, as the high bits of a byte
                    const uint32_t pos = running - (uint32_t)k;
                    const uint32_t wi = pos >> 5;
                    const uint64_t two = ((uint64_t)buf[wi] << 32) | (wi + 1 < (uint32_t)CAP_WORDS ? buf[wi + 1] : 0u);
                    tb = (uint32_t)(two >> (56 - (pos & 31))) & (0xff00u >> k) & 0xffu;
                } else if (k) { // a last group shorter than that (only after an early flush): it continues this warp's own byte
                    tb = (prev_tail | ((buf[0] >> 24) >> (int)(G & 7))) & (0xff00u >> k) & 0xffu;
                }
                *(volatile uint32_t *)&P.tailw[tile] = TAIL_VALID | tb;
            }
            if (flushed == 0 && tile > 0 && (G & 7)) { // first group of the tile: the byte shared with the previous tile
                uint32_t tw = 0;
                if (lane == 0) {
                    while (((tw = *(volatile uint32_t *)&P.tailw[tile - 1]) & TAIL_VALID) == 0) __nanosleep(200);
                }
                prev_tail = __shfl_sync(0xffffffffu, tw, 0) & 0xffu;
            }
            const uint64_t blo = G >> 3, bhi = E >> 3; // bytes completed by this group (the first may start in the previous one)
            const bool fits = ((E + 7) >> 3) <= P.ucap;
            if (!fits && lane == 0) atomicOr((unsigned long long *)&P.result[2], 2ull);
            const int sh = (int)(G & 31);
            const uint64_t w0 = G >> 5;
            const int nw = (int)((sh + running + 31) >> 5); // output words touched
            uint32_t tail_here = 0;
            for (int j = lane; j < nw; j += 32) {
                const uint32_t cur = j < CAP_WORDS ? buf[j] : 0u; // sh + running can reach one word past the buffer
                const uint32_t prev = j > 0 ? buf[j - 1] : 0u;
                uint32_t be = sh ? __funnelshift_r(cur, prev, sh) : cur; // big-endian bit order
                if (j == 0 && (G & 7)) be |= prev_tail << (24 - 8 * (int)((G >> 3) & 3)); // complete the shared byte
                const uint64_t b0 = (w0 + (uint64_t)j) << 2;                             // first global byte of this word
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
                if ((E & 7) && b0 <= bhi && bhi < b0 + 4) tail_here = 0x100u | ((be >> (24 - 8 * (int)(bhi & 3))) & 0xffu);
            }
            // after an early flush the trailing partial byte is completed by this warp's next group
            const unsigned who = __ballot_sync(0xffffffffu, tail_here != 0);
            prev_tail = who ? (__shfl_sync(0xffffffffu, tail_here, __ffs(who) - 1) & 0xffu) : 0u;
            __syncwarp();
            const int used = (int)((running + 31) >> 5) + 1;
            for (int j = lane; j < used && j < CAP_WORDS; j += 32) buf[j] = 0;
            __syncwarp();
            flushed += running;
            running = 0;
        };
        // lane's data unit in round `rr` (rr == r0 - 1: the halo MCU just before the tile, DC values only)
        auto unit_of = [&](long long rr, uint32_t &m, bool &valid) {
            const bool halo = rr < (long long)r0;
            m = halo ? r0 * MPR - 1 : (uint32_t)rr * MPR + mi;
            valid = lane_used && m < P.nmcu && (!halo || mi == 0);
        };
        // asynchronous copy of the round's pixels into shared memory
        auto fetch = [&](long long rr) {
            if (rr < (long long)r0) {
                stage_pixels<SUB>(P, pix, r0 * MPR - 1, 1, lane);
            } else {
                const uint32_t m0 = (uint32_t)rr * MPR, left = P.nmcu - m0;
                stage_pixels<SUB>(P, pix, m0, left < (uint32_t)MPR ? (int)left : MPR, lane);
            }
        };

        const long long first = r0 == 0 ? 0 : (long long)r0 - 1;
        fetch(first);
        for (long long rr = first; rr < (long long)r1; ++rr) {
            const bool halo = rr < (long long)r0;
            uint32_t m;
            bool valid;
            unit_of(rr, m, valid);
            cp_async_wait_all();
            __syncwarp();
            float s[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) s[i] = 0.f;

            // ---- colour conversion ----
            if (SUB) {
                float *slot = chroma + (mi * 2) * CHROMA_SLOT;
                if (valid && d < 4) {
                    const int qx = d & 1, qy = d >> 1;
                    const uint32_t *blk = pix + mi * PIX_MCU_WORDS_420 + qy * 128 + qx * 8;
                    float pu[8], pv[8];
#pragma unroll
                    for (int y = 0; y < 8; ++y) {
                        float cu[8], cv[8];
                        const uint4 pa = *reinterpret_cast<const uint4 *>(blk + y * 16), pb = *reinterpret_cast<const uint4 *>(blk + y * 16 + 4);
                        const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                        for (int x = 0; x < 8; ++x) {
                            const uint32_t p = pw[x];
                            const float r = (float)(p & 255u), g = (float)((p >> 8) & 255u), b = (float)((p >> 16) & 255u);
                            s[y * 8 + x] = to_y(r, g, b);
                            cu[x] = to_u(r, g, b);
                            cv[x] = to_v(r, g, b);
                        }
                        if (y & 1) { // (top-left + top-right + bottom-left + bottom-right) * 0.25f  (stb :1557-1558)
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const int at = (qy * 4 + (y >> 1)) * 8 + qx * 4 + x;
                                slot[at] = FM(FA(FA(FA(pu[2 * x], pu[2 * x + 1]), cu[2 * x]), cu[2 * x + 1]), 0.25f);
                                slot[CHROMA_SLOT + at] = FM(FA(FA(FA(pv[2 * x], pv[2 * x + 1]), cv[2 * x]), cv[2 * x + 1]), 0.25f);
                            }
                        } else {
#pragma unroll
                            for (int x = 0; x < 8; ++x) {
                                pu[x] = cu[x];
                                pv[x] = cv[x];
                            }
                        }
                    }
                }
                __syncwarp();
                if (valid && d >= 4) {
                    const float *src = slot + (d - 4) * CHROMA_SLOT;
#pragma unroll
                    for (int i = 0; i < 64; ++i) s[i] = src[i];
                }
                __syncwarp();
            } else {
                if (valid) {
                    const uint32_t *blk = pix + mi * PIX_MCU_WORDS_444;
#pragma unroll
                    for (int y = 0; y < 8; ++y) {
                        const uint4 pa = *reinterpret_cast<const uint4 *>(blk + y * 8), pb = *reinterpret_cast<const uint4 *>(blk + y * 8 + 4);
                        const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                        for (int x = 0; x < 8; ++x) {
                            const uint32_t p = pw[x];
                            const float r = (float)(p & 255u), g = (float)((p >> 8) & 255u), b = (float)((p >> 16) & 255u);
                            s[y * 8 + x] = d == 0 ? to_y(r, g, b) : d == 1 ? to_u(r, g, b) : to_v(r, g, b);
                        }
                    }
                }
            }

            if (halo) { // DC only: output 0 of the row passes, then of the column pass over them
                float rs[8];
#pragma unroll
                for (int y = 0; y < 8; ++y)
                    rs[y] = aan8_dc(s[y * 8], s[y * 8 + 1], s[y * 8 + 2], s[y * 8 + 3], s[y * 8 + 4], s[y * 8 + 5], s[y * 8 + 6], s[y * 8 + 7]);
                const int dc = quantise(aan8_dc(rs[0], rs[1], rs[2], rs[3], rs[4], rs[5], rs[6], rs[7]), mult[0]);
                carry_y = __shfl_sync(0xffffffffu, dc, SUB ? 3 : 0);
                carry_u = __shfl_sync(0xffffffffu, dc, SUB ? 4 : 1);
                carry_v = __shfl_sync(0xffffffffu, dc, SUB ? 5 : 2);
                __syncwarp();
                fetch(rr + 1);
                continue;
            }
            const uint32_t r = (uint32_t)rr;

            // ---- DCT, quantisation, DC difference; 64 int16 in zig-zag order to shared memory ----
            int dc = 0;
            if (valid) {
#pragma unroll
                for (int y = 0; y < 8; ++y)
                    aan8(s[y * 8], s[y * 8 + 1], s[y * 8 + 2], s[y * 8 + 3], s[y * 8 + 4], s[y * 8 + 5], s[y * 8 + 6], s[y * 8 + 7]);
#pragma unroll
                for (int x = 0; x < 8; ++x) aan8(s[x], s[8 + x], s[16 + x], s[24 + x], s[32 + x], s[40 + x], s[48 + x], s[56 + x]);
                dc = quantise(s[0], mult[0]);
            }
            int pred;
            {
                const int back = SUB ? (d == 0 ? 3 : d < 4 ? 1 : 6) : 3;
                const bool from_carry = SUB ? (mi == 0 && (d == 0 || d >= 4)) : mi == 0;
                const int src = lane - back;
                const int up = __shfl_sync(0xffffffffu, dc, src < 0 ? 0 : src);
                const int cy = SUB ? (d == 0 ? carry_y : d == 4 ? carry_u : carry_v) : (d == 0 ? carry_y : d == 1 ? carry_u : carry_v);
                pred = from_carry ? cy : up;
            }
            if (valid) {
                uint32_t *dst = coef + lane * COEF_STRIDE;
                int16_t *gdst = P.coefs ? P.coefs + ((size_t)m * DPM + d) * 64 : nullptr;
#pragma unroll
                for (int k = 0; k < 64; k += 2) {
                    const int a = k == 0 ? dc : quantise(s[kZZ.nat[k]], mult[kZZ.nat[k]]);
                    const int b = quantise(s[kZZ.nat[k + 1]], mult[kZZ.nat[k + 1]]);
                    if (gdst) {
                        gdst[k] = (int16_t)a;
                        gdst[k + 1] = (int16_t)b;
                    }
                    // position 0 carries the DC DIFFERENCE: that is what gets coded
                    dst[k >> 1] = (uint32_t)((k == 0 ? dc - pred : a) & 0xffff) | ((uint32_t)b << 16);
                }
            }
            {
                const uint32_t left = P.nmcu - r * MPR;
                const int nv = left < (uint32_t)MPR ? (int)left : MPR;
                const int base = (nv - 1) * DPM;
                carry_y = __shfl_sync(0xffffffffu, dc, base + (SUB ? 3 : 0));
                carry_u = __shfl_sync(0xffffffffu, dc, base + (SUB ? 4 : 1));
                carry_v = __shfl_sync(0xffffffffu, dc, base + (SUB ? 5 : 2));
            }
            __syncwarp();
            if (rr + 1 < (long long)r1) fetch(rr + 1); // in flight during the entropy phase

            // ---- entropy coding: one data unit per iteration, lane l codes zig-zag positions 2l and 2l+1 ----
            {
                const uint32_t left = P.nmcu - r * MPR;
                const int nunits = (left < (uint32_t)MPR ? (int)left : MPR) * DPM;
                int dd = 0; // data unit within its MCU
                for (int j = 0; j < nunits; ++j) {
                    const bool luma = SUB ? dd < 4 : dd == 0;
                    dd = dd + 1 == DPM ? 0 : dd + 1;
                    const uint32_t *tab_ac = sm + (luma ? T_AC_Y : T_AC_C);
                    const uint32_t *tab_dc = sm + (luma ? T_DC_Y : T_DC_C);
                    const uint32_t wd = coef[j * COEF_STRIDE + lane];
                    const int c0 = (int)(short)(wd & 0xffffu), c1 = (int)wd >> 16;
                    const uint32_t nz_e = __ballot_sync(0xffffffffu, c0 != 0) | 1u; // position 0 (DC) always bounds a run
                    const uint32_t nz_o = __ballot_sync(0xffffffffu, c1 != 0);
                    // nearest coded position before 2l: even ones are 2*i, odd ones 2*i+1
                    const uint32_t pe = nz_e & below, po = nz_o & below;
                    const int prev_e = max(62 - 2 * __clz(pe), 63 - 2 * __clz(po)); // -2 / -1 when empty (lane 0 only)
                    const int run0 = 2 * lane - 1 - prev_e;
                    const int run1 = (c0 != 0 || lane == 0) ? 0 : run0 + 1;
                    uint64_t v0 = 0, v1 = 0;
                    int l0 = 0, l1 = 0;
                    const int n0 = bitlen(c0), n1 = bitlen(c1);
                    if (lane == 0) {
                        const uint32_t e = tab_dc[n0];
                        v0 = ((uint64_t)(e >> 8) << n0) | (n0 ? extra_bits(c0, n0) : 0u);
                        l0 = (int)(e & 255u) + n0;
                    } else if (c0 != 0) {
                        const uint32_t e = tab_ac[(run0 & 15) * 16 + n0];
                        v0 = ((uint64_t)(e >> 8) << n0) | extra_bits(c0, n0);
                        l0 = (int)(e & 255u) + n0;
                    }
                    if (c1 != 0) {
                        const uint32_t e = tab_ac[(run1 & 15) * 16 + n1];
                        v1 = ((uint64_t)(e >> 8) << n1) | extra_bits(c1, n1);
                        l1 = (int)(e & 255u) + n1;
                    } else if (lane == 31) { // position 63 is zero: end of block
                        v1 = tab_ac[0] >> 8;
                        l1 = (int)(tab_ac[0] & 255u);
                    }
                    // runs of 16 or more zeros: ZRL codes in front (rare; at most 3 per symbol)
                    const bool z0 = lane != 0 && c0 != 0 && run0 >= 16, z1 = c1 != 0 && run1 >= 16;
                    const bool any_zrl = __any_sync(0xffffffffu, z0 || z1);
                    if (any_zrl) {
                        const uint32_t zrl = tab_ac[0xF0];
                        const int zl = (int)(zrl & 255u);
                        if (z0) {
                            for (int i = 0; i < (run0 >> 4); ++i) v0 |= (uint64_t)(zrl >> 8) << (l0 + i * zl);
                            l0 += (run0 >> 4) * zl;
                        }
                        if (z1) {
                            for (int i = 0; i < (run1 >> 4); ++i) v1 |= (uint64_t)(zrl >> 8) << (l1 + i * zl);
                            l1 += (run1 >> 4) * zl;
                        }
                    }
                    const uint32_t len = (uint32_t)(l0 + l1);
                    uint32_t incl = len;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const uint32_t unit_bits = __shfl_sync(0xffffffffu, incl, 31);
                    if (running + unit_bits + 7 > (uint32_t)CAP_BITS) flush(false);
                    const uint32_t off = running + incl - len;
                    if (any_zrl) { // up to 59 bits each
                        or_bits(buf, off, v0, l0);
                        or_bits(buf, off + (uint32_t)l0, v1, l1);
                    } else { // at most 2 x 27 bits: one string
                        or_bits(buf, off, (v0 << l1) | v1, (int)len);
                    }
                    running += unit_bits;
                }
            }
            __syncwarp();
        }
        if (r1 == P.nrounds) { // padding of the EOI marker: seven one-bits (stb :1586)
            if (lane == 0) or_bits(buf, running, 0x7Fu, 7);
            running += 7;
        }
        flush(true);
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
    const uint64_t nbytes = *P.total_bits >> 3; // whole bytes only: a trailing partial byte is dropped (stb :1586)
    const uint64_t nchunks = (nbytes + STUFF_CHUNK - 1) / STUFF_CHUNK;
    if (nbytes > P.ucap) { // the encode kernel ran out of scratch (= the caller's buffer is too small): report a lower bound
        if (blockIdx.x == 0 && tid == 0) {
            P.result[0] = HEADER_BYTES + nbytes + 2;
            atomicOr((unsigned long long *)&P.result[2], 1ull);
        }
        return;
    }
    for (;;) {
        if (tid == 0) s_chunk = (uint32_t)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        __syncthreads();
        const uint64_t chunk = s_chunk;
        if (chunk >= (nchunks ? nchunks : 1)) break; // an empty stream still runs chunk 0 (header + EOI)
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
        const bool is_last = chunk + 1 >= nchunks;
        const uint64_t end = base + total + (is_last ? 2 : 0);
        if (is_last && tid == 0) {
            stage8[total] = 0xFF; // EOI
            stage8[total + 1] = 0xD9;
            P.result[0] = end;
            if (end > P.out_cap) atomicOr((unsigned long long *)&P.result[2], 1ull);
        }
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

extern "C" int ljb_jfif_encode_dev(ljb_ctx *ctx, const uint8_t *d_pixels, int w, int h, int comp, size_t stride, int quality,
                                   int subsample, uint8_t *d_out, size_t out_cap, uint64_t *d_result, int16_t *d_coefs)
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
    // rounds per tile: as many as fit the 4 KB bit buffer on uniform noise (the densest realistic input) at this quality;
    // denser tiles still encode correctly through early flushes, only slower
    const int q0 = quality ? quality : 90;
    int R = q0 <= 60 ? 6 : q0 <= 80 ? 4 : q0 <= 90 ? 3 : q0 <= 95 ? 2 : 1;
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
    const size_t o_end = o_tailw + (size_t)ntiles * 4;
    if ((rc = ljb_ensure(&ctx->d_status, &ctx->status_bytes, o_end)) != 0) return rc;
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
    P.total_bits = (uint64_t *)(sb + o_total);
    P.result = d_result;
    P.coefs = d_coefs;
    P.tables = (const uint32_t *)(sb + o_tab);
    memcpy(P.mult, pl.mult, sizeof P.mult);

    static bool attr_done = false;
    if (!attr_done) {
        LJB_CUDA(cudaFuncSetAttribute(jfif_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES));
        LJB_CUDA(cudaFuncSetAttribute(jfif_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES));
        attr_done = true;
    }
    const size_t want = ((size_t)ntiles + NWARPS - 1) / NWARPS;
    const size_t full = (size_t)ctx->num_sms * 2;
    const int grid = (int)(want < full ? want : full);
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (pl.subsample) jfif_encode_kernel<true><<<grid, THREADS, SM_BYTES, ctx->stream>>>(P);
    else jfif_encode_kernel<false><<<grid, THREADS, SM_BYTES, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    StuffParams S;
    S.ustream = P.ustream;
    S.ucap = ucap;
    S.total_bits = P.total_bits;
    S.status = (uint64_t *)(sb + o_st2);
    S.out = d_out;
    S.out_cap = out_cap;
    S.result = d_result;
    S.hdr = pl.hdr;
    const size_t sfull = (size_t)ctx->num_sms * 8;
    const int sgrid = (int)(nchunks_max < sfull ? nchunks_max : sfull);
    jfif_stuff_kernel<<<sgrid, STUFF_THREADS, 0, ctx->stream>>>(S);
    LJB_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return LJB_OK;
}

// Host-buffer entry point: upload, encode, download.  (One bit stream per image: no band pipeline as in
// ljb_jpeg_encode_rgba; the whole image is uploaded before the kernels start.)
extern "C" int ljb_jfif_encode(ljb_ctx *ctx, const uint8_t *pixels, int w, int h, int comp, size_t stride, int quality, int subsample,
                               uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!ctx || !pixels || !out || !out_len || w <= 0 || h <= 0 || comp < 1 || comp > 4 || stride < (size_t)w * (size_t)comp) return LJB_E_ARG;
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
    LJB_CUDA(cudaMemcpy2DAsync(ctx->d_pin[0], dstride, pixels, stride, rowbytes, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    rc = ljb_jfif_encode_dev(ctx, (const uint8_t *)ctx->d_pin[0], w, h, comp, dstride, quality, subsample, (uint8_t *)ctx->d_pout[0], dcap,
                             (uint64_t *)ctx->d_small, nullptr);
    if (rc != 0) return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, ctx->d_small, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_len) *out_len = (size_t)res[0];
    if (res[2] & 3) return LJB_E_CAPACITY;
    LJB_CUDA(cudaMemcpyAsync(out, ctx->d_pout[0], (size_t)res[0], cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}
