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
// thread owns one group, 192 groups per CTA tile, persistent CTAs pull tiles from a ticket counter.
//
// Exactness: colour conversion is evaluated in IEEE double with explicit round-to-nearest mul/add in
// the reference's left-to-right order (no FMA contraction).  The DCT is evaluated twice at most: a
// separable double-precision fast path, and — only for a coefficient whose quotient lies within 1e-6
// of a non-zero integer, where truncation could go either way — the reference's own summation order
// (x outer, y inner, (corr*cos_x)*cos_y, no FMA) with the exact cos()/sqrt() doubles glibc returns
// (jpeg_tables.inc).  The result is bit-identical quantised coefficients, hence identical bit strings.
#include "common.cuh"

namespace jpgk {

#include "jpeg_tables.inc"

constexpr int THREADS = 192;       // groups per tile
constexpr int REC_BYTES = 256;     // max packed record: (1023 + 511 + 511) bits (JPEG.c:1248, :1286, :1320)
constexpr int MAXSYM = 72;         // <= 64 distinct values + <= 10 distinct run lengths
constexpr int MAXNODE = 2 * MAXSYM;

// per-thread scratch in shared memory (bytes); the stride is an odd number of words so that threads
// touching the same logical index fall into different banks
constexpr int OFF_COEF = 0;        // int16[128]: zig-zag ordered quantised lum[64], r[32], b[32]
constexpr int OFF_LUT = 256;       // uint8[320]: symbol value+160 -> slot   (aliased by samples u8[128] before entropy)
constexpr int OFF_X = 576;         // 288 B: while building: cnt u8[144] | heap u8[72]; afterwards code u32[72]
constexpr int OFF_PAR = 864;       // uint8[144] parent node
constexpr int OFF_PBIT = 1008;     // uint32[5]  bit of each node under its parent
constexpr int OFF_LEN = 1028;      // uint8[72]  code length per slot
constexpr int STRIDE = 1100;
static_assert((STRIDE / 4) % 2 == 1 && STRIDE % 4 == 0, "stride must be an odd number of words");
constexpr int SM_THREADS = THREADS * STRIDE;
constexpr int SM_COS8 = SM_THREADS;            // double[64]
constexpr int SM_COS4 = SM_COS8 + 512;         // double[16]
constexpr int SM_MISC = SM_COS4 + 128;
constexpr int SM_TOTAL = SM_MISC + 2048;
static_assert(SM_TOTAL <= 227 * 1024, "exceeds B200 shared memory per CTA");

struct Misc {
    unsigned int rec_off[THREADS + 1]; // exclusive scan of record sizes inside the tile
    unsigned int warp_sum[THREADS / 32];
    long long ticket;
    unsigned long long base;
    int emit_ok;
};

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
    uint64_t *status;        // [0] ticket, [1..] look-back words (one per tile)
    uint32_t *scratch;       // per CTA: THREADS * 64 words of record staging
    uint32_t ntiles;
};

// JPEG.c:12-27 as doubles, and their reciprocals for the fast path (the chroma table is consumed as
// 8 rows x 4 columns)
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
__constant__ double kRLum[64] = {
    1.0 / 8.0, 1.0 / 6.0, 1.0 / 6.0, 1.0 / 8.0, 1.0 / 10.0, 1.0 / 14.0, 1.0 / 18.0, 1.0 / 22.0,
    1.0 / 6.0, 1.0 / 6.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 12.0, 1.0 / 20.0, 1.0 / 22.0, 1.0 / 20.0,
    1.0 / 6.0, 1.0 / 7.0, 1.0 / 8.0, 1.0 / 10.0, 1.0 / 14.0, 1.0 / 22.0, 1.0 / 25.0, 1.0 / 22.0,
    1.0 / 8.0, 1.0 / 9.0, 1.0 / 10.0, 1.0 / 14.0, 1.0 / 18.0, 1.0 / 28.0, 1.0 / 27.0, 1.0 / 22.0,
    1.0 / 10.0, 1.0 / 12.0, 1.0 / 14.0, 1.0 / 18.0, 1.0 / 22.0, 1.0 / 35.0, 1.0 / 33.0, 1.0 / 26.0,
    1.0 / 14.0, 1.0 / 18.0, 1.0 / 22.0, 1.0 / 22.0, 1.0 / 27.0, 1.0 / 33.0, 1.0 / 36.0, 1.0 / 30.0,
    1.0 / 18.0, 1.0 / 22.0, 1.0 / 26.0, 1.0 / 28.0, 1.0 / 33.0, 1.0 / 40.0, 1.0 / 40.0, 1.0 / 34.0,
    1.0 / 22.0, 1.0 / 26.0, 1.0 / 28.0, 1.0 / 30.0, 1.0 / 36.0, 1.0 / 34.0, 1.0 / 35.0, 1.0 / 33.0,
};
__constant__ double kQChr[32] = {
    17.0, 18.0, 24.0, 47.0, 18.0, 21.0, 26.0, 66.0,
    24.0, 26.0, 56.0, 99.0, 47.0, 66.0, 99.0, 99.0,
    66.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0,
    99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0, 99.0,
};
__constant__ double kRChr[32] = {
    1.0 / 17.0, 1.0 / 18.0, 1.0 / 24.0, 1.0 / 47.0, 1.0 / 18.0, 1.0 / 21.0, 1.0 / 26.0, 1.0 / 66.0,
    1.0 / 24.0, 1.0 / 26.0, 1.0 / 56.0, 1.0 / 99.0, 1.0 / 47.0, 1.0 / 66.0, 1.0 / 99.0, 1.0 / 99.0,
    1.0 / 66.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0,
    1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0, 1.0 / 99.0,
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

// ---- colour conversion, JPEG.c:127, :157, :180 (double, left to right, no contraction) ---------------
__device__ __forceinline__ int luma_of(int r, int g, int b)
{
    double y = __dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)r), __dmul_rn(0.587, (double)g)), __dmul_rn(0.114, (double)b));
    return (int)y & 0xFF; // implicit double -> uint8_t conversion of a value in [0, 255]
}
__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
__device__ __forceinline__ int cr_of(int r, int g, int b)
{
    double v = __dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(0.439, (double)r), __dmul_rn(0.368, (double)g)), __dmul_rn(0.071, (double)b)), 128.0);
    return clamp255((int)v);
}
__device__ __forceinline__ int cb_of(int r, int g, int b)
{
    double v = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(-0.148, (double)r), __dmul_rn(0.291, (double)g)), __dmul_rn(0.439, (double)b)), 128.0);
    return clamp255((int)v);
}

// ---- the reference's own summation, JPEG.c:471-490, for one coefficient ------------------------------
template <int W>
__device__ __noinline__ double exact_coef(const uint8_t *smp, int u, int v, const double *c8, const double *c4)
{
    double sum = 0.0;
    for (int x = 0; x < 8; ++x) {
        const double cx = c8[x * 8 + u];
        for (int y = 0; y < W; ++y) {
            const double cy = (W == 8) ? c8[y * 8 + v] : c4[y * 4 + v];
            const double corr = (double)((int)smp[x * W + y] - 128);
            sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(corr, cx), cy));
        }
    }
    const double au = u == 0 ? kAlpha8[0] : kAlpha8[1];
    const double av = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) : (v == 0 ? kAlpha4[0] : kAlpha4[1]);
    return __dmul_rn(__dmul_rn(au, av), sum);
}

// DCT + quantise + zig-zag of one channel (W = 8 luma, W = 4 chroma; 8 rows).
template <int W>
__device__ __forceinline__ void transform_channel(const uint8_t *smp, int16_t *cz, int16_t *coefs_out, const double *c8s,
                                                  const double *c4s)
{
    constexpr ZigZag<W, 8> zz;
    double T[8][W];
    // row pass: T[x][v] = sum_y corr[x][y] * cos_y[y][v]
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        double c[W];
#pragma unroll
        for (int y = 0; y < W; ++y) c[y] = (double)((int)smp[x * W + y] - 128);
#pragma unroll
        for (int v = 0; v < W; ++v) {
            double acc = 0.0;
#pragma unroll
            for (int y = 0; y < W; ++y) acc = fma(c[y], (W == 8) ? kCos8[y * 8 + v] : kCos4[y * 4 + v], acc);
            T[x][v] = acc;
        }
    }
    // column pass + quantise
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int v = 0; v < W; ++v) {
            double acc = 0.0;
#pragma unroll
            for (int x = 0; x < 8; ++x) acc = fma(T[x][v], kCos8[x * 8 + u], acc);
            const double au = u == 0 ? kAlpha8[0] : kAlpha8[1];
            const double av = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) : (v == 0 ? kAlpha4[0] : kAlpha4[1]);
            const double q = (W == 8) ? kQLum[u * W + v] : kQChr[u * W + v];
            const double qf = ((au * av) * acc) * ((W == 8) ? kRLum[u * W + v] : kRChr[u * W + v]); // within 2 ulp of the quotient
            int t = (int)qf; // truncation toward zero, JPEG.c:627
            // |qf| < 1 - 1e-6 truncates to 0 whatever the last bits are; otherwise a quotient within 1e-6
            // of an integer is re-evaluated in the reference's own order (fast-path error is < 1e-9)
            if (fabs(qf) >= 0.999999 && fabs(qf - rint(qf)) < 1e-6) {
                const double ce = exact_coef<W>(smp, u, v, c8s, c4s);
                t = (int)__ddiv_rn(ce, q);
            }
            cz[zz.pos[u * W + v]] = (int16_t)t;
            if (coefs_out) coefs_out[u * W + v] = (int16_t)t;
        }
    }
}

// ---- per-channel adaptive Huffman code + emission -----------------------------------------------------
struct BitWriter {
    unsigned long long acc;
    int nbits;          // bits pending in acc (< 32 between calls)
    uint32_t *dst;      // record staging, REC_BYTES / 4 words
    int wpos;
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc = (acc << len) | code;
        nbits += len;
        if (nbits >= 32) {
            uint32_t wv = (uint32_t)(acc >> (nbits - 32));
            if (wpos < REC_BYTES / 4) dst[wpos] = __byte_perm(wv, 0, 0x0123); // MSB-first bytes
            ++wpos;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nbits > 0) {
            uint32_t wv = (uint32_t)(acc << (32 - nbits));
            if (wpos < REC_BYTES / 4) dst[wpos] = __byte_perm(wv, 0, 0x0123);
        }
    }
};

__device__ __forceinline__ void heapify(uint8_t *heap, const uint8_t *cnt, int size, int i) // JPEG.c:894-911
{
    for (;;) {
        int smallest = i, l = 2 * i + 1, r = 2 * i + 2;
        if (l < size && cnt[heap[l]] < cnt[heap[smallest]]) smallest = l;
        if (r < size && cnt[heap[r]] < cnt[heap[smallest]]) smallest = r;
        if (smallest == i) return;
        uint8_t t = heap[i];
        heap[i] = heap[smallest];
        heap[smallest] = t;
        i = smallest;
    }
}

// Returns the number of bits emitted; sets *bad when the reference itself would overflow.
template <int N>
__device__ __noinline__ int entropy_channel(uint8_t *ts, const int16_t *cz, BitWriter &bw, int max_bits, int *bad)
{
    uint8_t *lut = ts + OFF_LUT;
    uint8_t *cnt = ts + OFF_X;
    uint8_t *heap = ts + OFF_X + MAXNODE;
    uint32_t *code = reinterpret_cast<uint32_t *>(ts + OFF_X);
    uint8_t *par = ts + OFF_PAR;
    uint32_t *pbit = reinterpret_cast<uint32_t *>(ts + OFF_PBIT);
    uint8_t *len = ts + OFF_LEN;

    // calculate_frequency (JPEG.c:864-885): symbols in first-appearance order; RLE emits (count, value)
    int k = 0;
    for (int i = 0; i < N;) {
        const int v = cz[i];
        int j = i + 1;
        while (j < N && cz[j] == v) ++j;
        const int sym[2] = {j - i, v};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int idx = sym[s] + 160;
            const int slot = lut[idx];
            if (slot == 0xFF) {
                lut[idx] = (uint8_t)k;
                cnt[k] = 1;
                ++k;
            } else {
                cnt[slot]++;
            }
        }
        i = j;
    }
    // build_heap (JPEG.c:913-934)
    for (int i = 0; i < k; ++i) heap[i] = (uint8_t)i;
    for (int i = k / 2 - 1; i >= 0; --i) heapify(heap, cnt, k, i);
    // build_huffman_tree (JPEG.c:936-961): pop two, append their parent at the END of the array (no sift-up)
#pragma unroll
    for (int i = 0; i < 5; ++i) pbit[i] = 0;
    int size = k, next = k;
    while (size > 1) {
        const int left = heap[0];
        heap[0] = heap[--size];
        heapify(heap, cnt, size, 0);
        const int right = heap[0];
        heap[0] = heap[--size];
        heapify(heap, cnt, size, 0);
        cnt[next] = (uint8_t)(cnt[left] + cnt[right]);
        par[left] = (uint8_t)next;
        par[right] = (uint8_t)next;
        pbit[right >> 5] |= 1u << (right & 31);
        heap[size++] = (uint8_t)next;
        ++next;
    }
    const int root = heap[0];
    // assign_codes (JPEG.c:963-982): left = '0', right = '1'; a leaf's code is read off its path to the root.
    // (cnt/heap are dead from here on; code[] reuses their storage.)
    uint32_t codes_tmp;
    for (int s = 0; s < k; ++s) {
        uint32_t c = 0;
        int l = 0, node = s;
        while (node != root) {
            if (l < 32) c |= ((pbit[node >> 5] >> (node & 31)) & 1u) << l;
            ++l;
            node = par[node];
        }
        if (l > 31) *bad = 1; // HuffmanCode.code is char[32] (JPEG.c:861)
        len[s] = (uint8_t)(l > 31 ? 31 : l);
        codes_tmp = c;
        // code[] aliases cnt[]/heap[]: slots >= s*4 bytes may still hold parents' cnt, which are no longer read
        code[s] = codes_tmp;
    }
    // generate_encoded_sequence (JPEG.c:993-1007)
    int bits = 0;
    for (int i = 0; i < N;) {
        const int v = cz[i];
        int j = i + 1;
        while (j < N && cz[j] == v) ++j;
        const int s0 = lut[(j - i) + 160], s1 = lut[v + 160];
        bw.put(code[s0], len[s0]);
        bw.put(code[s1], len[s1]);
        bits += len[s0] + len[s1];
        i = j;
    }
    if (bits > max_bits) *bad = 1; // char encoded_sequence[1024] / [512] (JPEG.c:1248, :1286)
    // reset the lookup table for the next channel
    for (int i = 0; i < N;) {
        const int v = cz[i];
        int j = i + 1;
        while (j < N && cz[j] == v) ++j;
        lut[(j - i) + 160] = 0xFF;
        lut[v + 160] = 0xFF;
        i = j;
    }
    return bits;
}

__global__ void __launch_bounds__(THREADS, 1) jpeg_encode_kernel(Params P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    uint8_t *ts = smem + tid * STRIDE;
    double *c8s = reinterpret_cast<double *>(smem + SM_COS8);
    double *c4s = reinterpret_cast<double *>(smem + SM_COS4);
    Misc &M = *reinterpret_cast<Misc *>(smem + SM_MISC);
    for (int i = tid; i < 64; i += THREADS) c8s[i] = kCos8[i];
    for (int i = tid; i < 16; i += THREADS) c4s[i] = kCos4[i];
    uint32_t *stage = P.scratch + ((size_t)blockIdx.x * THREADS + tid) * (REC_BYTES / 4);
    const size_t bpr = ((size_t)P.w + 7) / 8;
    const bool aligned = ((reinterpret_cast<uintptr_t>(P.rgba) | P.stride) & 15) == 0;

    for (;;) {
        if (tid == 0) M.ticket = (long long)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        __syncthreads();
        const long long tile = M.ticket;
        if (tile >= (long long)P.ntiles) break;
        const size_t gl = (size_t)tile * THREADS + tid; // group index inside this call
        const bool active = gl < P.ngroups;
        unsigned int rec_bytes = 0;
        int bl = 0, br = 0, bb = 0;
        if (active) {
            const size_t g = P.first_group + gl;
            const size_t brow = g / bpr, bcol = g % bpr;
            // ---- colour conversion + 4:2:2 point subsampling + tiling -> samples u8[64 | 32 | 32]
            uint8_t *smp = ts + OFF_LUT;
            for (int lr = 0; lr < 8; ++lr) {
                const size_t row = brow * 8 + lr;
                uint32_t px[8];
                const size_t col0 = bcol * 8;
                if (row < (size_t)P.h) {
                    const uint8_t *rp = P.rgba + row * P.stride + col0 * 4;
                    if (aligned && col0 + 8 <= (size_t)P.w) {
                        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rp));
                        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(rp) + 1);
                        px[0] = a.x; px[1] = a.y; px[2] = a.z; px[3] = a.w;
                        px[4] = b.x; px[5] = b.y; px[6] = b.z; px[7] = b.w;
                    } else {
#pragma unroll
                        for (int lc = 0; lc < 8; ++lc) {
                            px[lc] = 0xFFFFFFFFu; // marks "outside the image": sample stays 0
                            if (col0 + lc < (size_t)P.w) {
                                const uint8_t *q = rp + 4 * lc;
                                px[lc] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int lc = 0; lc < 8; ++lc) px[lc] = 0xFFFFFFFFu;
                }
                const bool full = row < (size_t)P.h && col0 + 8 <= (size_t)P.w;
#pragma unroll
                for (int lc = 0; lc < 8; ++lc) {
                    const bool inside = full || (row < (size_t)P.h && col0 + lc < (size_t)P.w);
                    const int r = px[lc] & 0xFF, gg = (px[lc] >> 8) & 0xFF, b = (px[lc] >> 16) & 0xFF;
                    smp[lr * 8 + lc] = inside ? (uint8_t)luma_of(r, gg, b) : 0;
                    if (lc & 1) { // chroma sample of local column lc-1 is the original chroma at column lc
                        smp[64 + lr * 4 + (lc >> 1)] = inside ? (uint8_t)cr_of(r, gg, b) : 0;
                        smp[96 + lr * 4 + (lc >> 1)] = inside ? (uint8_t)cb_of(r, gg, b) : 0;
                    }
                }
            }
            // ---- DCT + quantise + zig-zag
            int16_t *cz = reinterpret_cast<int16_t *>(ts + OFF_COEF);
            int16_t *co = P.coefs ? P.coefs + gl * 128 : nullptr;
            transform_channel<8>(smp, cz, co, c8s, c4s);
            transform_channel<4>(smp + 64, cz + 64, co ? co + 64 : nullptr, c8s, c4s);
            transform_channel<4>(smp + 96, cz + 96, co ? co + 96 : nullptr, c8s, c4s);
            // ---- entropy coding of lum, r, b (reference order JPEG.c:1242, :1284, :1318)
            uint32_t *lutw = reinterpret_cast<uint32_t *>(ts + OFF_LUT);
#pragma unroll 4
            for (int i = 0; i < 80; ++i) lutw[i] = 0xFFFFFFFFu;
            BitWriter bw;
            bw.acc = 0;
            bw.nbits = 0;
            bw.dst = stage;
            bw.wpos = 0;
            int bad = 0;
            bl = entropy_channel<64>(ts, cz, bw, 1023, &bad);
            br = entropy_channel<32>(ts, cz + 64, bw, 511, &bad);
            bb = entropy_channel<32>(ts, cz + 96, bw, 511, &bad);
            bw.finish();
            if (bad) {
                atomicAdd((unsigned long long *)&P.result[1], 1ull);
                atomicOr((unsigned long long *)&P.result[2], 2ull);
            }
            const int total_bits = bl + br + bb;
            rec_bytes = (unsigned int)min((total_bits + 7) >> 3, REC_BYTES);
        }
        // ---- tile-wide exclusive scan of record sizes
        unsigned int inc = rec_bytes;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) M.warp_sum[warp] = inc;
        __syncthreads();
        unsigned int wbase = 0, tile_total = 0;
        for (int k = 0; k < THREADS / 32; ++k) {
            if (k < warp) wbase += M.warp_sum[k];
            tile_total += M.warp_sum[k];
        }
        M.rec_off[tid] = wbase + inc - rec_bytes;
        if (tid == 0) M.rec_off[THREADS] = tile_total;
        // ---- place the tile in the output stream
        if (warp == 0) {
            unsigned long long base = ljb_lookback(P.status + 1, tile, tile_total, 0);
            if (lane == 0) {
                M.base = base;
                M.emit_ok = (base + tile_total <= P.out_cap) ? 1 : 0;
                if (!M.emit_ok) atomicOr((unsigned long long *)&P.result[2], 1ull);
                if (tile == (long long)P.ntiles - 1) {
                    P.result[0] = base + tile_total;
                    P.group_offsets[P.ngroups] = base + tile_total;
                }
            }
        }
        __syncthreads();
        const unsigned long long base = M.base;
        if (active) {
            P.group_offsets[gl] = base + M.rec_off[tid];
            if (P.group_bits) {
                P.group_bits[3 * gl + 0] = (uint16_t)bl;
                P.group_bits[3 * gl + 1] = (uint16_t)br;
                P.group_bits[3 * gl + 2] = (uint16_t)bb;
            }
        }
        // ---- gather the staged records into the stream: one warp per record, coalesced byte runs
        if (M.emit_ok) {
            const uint8_t *tile_stage = reinterpret_cast<const uint8_t *>(P.scratch + (size_t)blockIdx.x * THREADS * (REC_BYTES / 4));
            for (int r = warp; r < THREADS; r += THREADS / 32) {
                const unsigned int o0 = M.rec_off[r], o1 = M.rec_off[r + 1];
                const uint8_t *srcp = tile_stage + (size_t)r * REC_BYTES;
                uint8_t *dstp = P.out + base + o0;
                for (unsigned int k = lane; k < o1 - o0; k += 32) dstp[k] = srcp[k];
            }
        }
        __syncthreads();
    }
}

} // namespace jpgk

extern "C" size_t ljb_jpeg_group_count(int w, int h)
{
    if (w <= 0 || h <= 0) return 0;
    return ((size_t)w * (size_t)h + 63) / 64; // JPEG.c:1131
}

extern "C" size_t ljb_jpeg_bound(size_t ngroups) { return ngroups * (size_t)jpgk::REC_BYTES + 64; }

extern "C" int ljb_jpeg_encode_rgba_dev(ljb_ctx *ctx, const uint8_t *d_rgba, int w, int h, size_t stride, size_t first_group,
                                        size_t ngroups, uint8_t *d_out, size_t out_cap, uint64_t *d_group_offsets,
                                        uint16_t *d_group_bits, int16_t *d_coefs, uint64_t *d_result)
{
    using namespace jpgk;
    if (!ctx || !d_rgba || !d_out || !d_group_offsets || !d_result || w <= 0 || h <= 0 || (w & 1) || stride < (size_t)w * 4)
        return LJB_E_ARG; // odd widths make the reference read past its subsampled rows (JPEG.c:543 with :314)
    const size_t total = ljb_jpeg_group_count(w, h);
    if (ngroups == 0 || first_group + ngroups > total) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    const size_t ntiles = (ngroups + THREADS - 1) / THREADS;
    if (ntiles > 0x7fffffffull) return LJB_E_ARG;
    const int grid = (int)((ntiles < (size_t)ctx->num_sms) ? ntiles : (size_t)ctx->num_sms);
    int rc;
    if ((rc = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, (size_t)ctx->num_sms * THREADS * REC_BYTES)) != 0) return rc;
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
    static bool attr_done = false;
    if (!attr_done) {
        LJB_CUDA(cudaFuncSetAttribute(jpeg_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_done = true;
    }
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    jpeg_encode_kernel<<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

extern "C" int ljb_jpeg_encode_rgba(ljb_ctx *ctx, const uint8_t *rgba, int w, int h, size_t stride, size_t first_group,
                                    size_t ngroups, uint8_t *out, size_t out_cap, uint64_t *group_offsets, uint16_t *group_bits,
                                    int16_t *coefs, size_t *out_len)
{
    if (!ctx || !rgba || !out || w <= 0 || h <= 0 || (w & 1) || stride < (size_t)w * 4) return LJB_E_ARG;
    const size_t total = ljb_jpeg_group_count(w, h);
    if (ngroups == 0 || first_group + ngroups > total) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // device layout of the staging area: image rows packed at a 16-byte aligned stride
    const size_t dstride = ((size_t)w * 4 + 15) & ~(size_t)15;
    size_t dcap = ljb_jpeg_bound(ngroups);
    if (out_cap < dcap) dcap = out_cap;
    if ((rc = ljb_ensure(&ctx->d_stage_in, &ctx->stage_in_bytes, dstride * (size_t)h + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_stage_out, &ctx->stage_out_bytes, dcap + 64)) != 0) return rc;
    const size_t small = (ngroups + 1 + 3) * sizeof(uint64_t) + ngroups * 3 * sizeof(uint16_t) + 64 + (coefs ? ngroups * 128 * sizeof(int16_t) : 0);
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, small)) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;
    uint64_t *d_res = d_offs + ngroups + 1;
    int16_t *d_coefs = coefs ? (int16_t *)(d_res + 3) : nullptr;
    uint16_t *d_bits = (uint16_t *)((uint8_t *)(d_res + 3) + (coefs ? ngroups * 128 * sizeof(int16_t) : 0));
    LJB_CUDA(cudaMemcpy2DAsync(ctx->d_stage_in, dstride, rgba, stride, (size_t)w * 4, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    rc = ljb_jpeg_encode_rgba_dev(ctx, (const uint8_t *)ctx->d_stage_in, w, h, dstride, first_group, ngroups,
                                  (uint8_t *)ctx->d_stage_out, dcap, d_offs, d_bits, d_coefs, d_res);
    if (rc != 0) return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_len) *out_len = (size_t)res[0];
    if (res[2] & 1) return LJB_E_CAPACITY;
    LJB_CUDA(cudaMemcpyAsync(out, ctx->d_stage_out, (size_t)res[0], cudaMemcpyDeviceToHost, ctx->stream));
    if (group_offsets)
        LJB_CUDA(cudaMemcpyAsync(group_offsets, d_offs, (ngroups + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (group_bits)
        LJB_CUDA(cudaMemcpyAsync(group_bits, d_bits, ngroups * 3 * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (coefs)
        LJB_CUDA(cudaMemcpyAsync(coefs, d_coefs, ngroups * 128 * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res[2] & 2) return LJB_E_UNSUPPORTED;
    return LJB_OK;
}
