// lz4-jpeg_b200/csrc/lz4_encode.cu — LZ4 (reference dialect) block encoder for sm_100a.
//
// Replaces, per 64 KiB block, the reference's block_encode() -> find_longest_match() -> write_block()
// (Algorithms/sequential/LZ4/LZ4.c:506-620, :290-323, :365-425; thread-per-block form
// parallel_block_encode, Algorithms/parallel/LZ4/LZ4.c:518-628) and, across blocks, the serial
// concatenation of write_output() (LZ4.c:427-441).
//
// One persistent CTA (1024 threads) per SM pulls blocks from a ticket counter and runs, per block:
//   P1 stage    : the block is copied once from HBM into shared memory (128-bit coalesced loads)
//   P2 index    : counting sort of every position by a 13-bit hash of its 4-gram (shared-memory atomics:
//                 histogram, exclusive scan, scatter in 64 position-ordered rounds)
//   P3 search   : exact longest-previous-match for EVERY position — equivalent to the reference's
//                 exhaustive scan because a match of length >= 4 shares its 4-gram with the current
//                 position; ties resolved to the earliest position (strict '>' in LZ4.c:307), length
//                 capped at min(1024, block end).  Two phases:
//                 A  threads walk the index in SORTED order, so the lanes of a warp sit in the same
//                    bucket: equal trip counts, and the candidate's bytes are one broadcast load.  Only
//                    the first 8 bytes are compared: matches shorter than 8 are final here.
//                 B  positions that have an 8-byte match are resolved from a second index keyed by the
//                    8-gram (tiny buckets), in position order, each thread inheriting the previous
//                    position's match along its diagonal instead of re-comparing up to 1024 bytes.
//   P4 parse    : the greedy chain 0 -> p+step[p] is resolved in parallel with per-segment exit tables,
//                 then sequences are sized with the reference's uint8/uint16 wrap rules
//   P5 place    : decoupled look-back over the per-block byte counts gives the block's output offset
//   P6 emit     : sequences are serialised straight to their final position in the output stream
// HBM traffic per block is therefore N_in + N_out (+ a per-CTA scratch that stays in L2).
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

namespace lz4k {

constexpr int THREADS = 1024;
constexpr int NWARPS = THREADS / 32;
constexpr int MAXB = 65536;          // LJB_LZ4_MAX_BLOCK
constexpr int HASH_BITS = 13;
constexpr int NBUCKET = 1 << HASH_BITS;
constexpr int SEG = 68;              // parse segment: 17 words, so per-thread segment walks are bank-conflict free
constexpr int MAXSEG = (MAXB + SEG - 1) / SEG; // 964
constexpr int MAX_MATCH = 1024;      // LZ4.c:20
constexpr int REGION = MAXB / NWARPS; // 2048 positions per warp in the sequence passes

// ---- shared memory map (bytes) ---------------------------------------------------------------------
constexpr int SM_DATA = 0;                         // 65536 + 64 pad
constexpr int SM_B = MAXB + 64;                    // region B
// index / search view of region B
constexpr int SM_S = SM_B;                         // u16 S[65536]          sorted positions
constexpr int SM_DIR = SM_S + 2 * MAXB;            // u32 dirw[4096 + 1]    packed u16 bucket ends
// parse view of region B
constexpr int SM_STEP = SM_B;                      // u8 step[65536 + 64]
constexpr int SM_FLAG = SM_STEP + MAXB + 64;       // u8 x1/flag[65536 + 64]
constexpr int SM_ENTRY = SM_FLAG + MAXB + 64;      // u8 entry[1024]
constexpr int SM_LONG = SM_DIR + 4 * (NBUCKET / 2 + 4); // u32 longbits[2048]: positions that have an >= 8 byte match
constexpr int SM_MISC = SM_LONG + MAXB / 8;
constexpr int SM_TOTAL = SM_MISC + 1024;
static_assert(SM_ENTRY + 1024 <= SM_LONG, "parse view must fit inside region B");
static_assert(SM_TOTAL <= 227 * 1024, "exceeds B200 shared memory per CTA");

struct Misc {
    unsigned long long warp_bytes[NWARPS];   // payload bytes per warp region
    unsigned long long warp_sizes[NWARPS];   // sum of byte_size fields per warp region (header arithmetic)
    unsigned int warp_nseq[NWARPS];
    unsigned int warp_phantom[NWARPS];
    unsigned int warp_last_end[NWARPS];      // end of the last match in the region (0 = none)
    unsigned int scan_tmp[NWARPS];
    long long ticket;
    unsigned long long base;                 // output offset of this block
    int emit_ok;
};

struct Params {
    const uint8_t *in;
    size_t n;
    uint32_t block_len;
    uint32_t nblocks;      // blocks in this call (shard)
    uint8_t *out;
    size_t out_cap;
    uint64_t *block_offsets; // nblocks + 1
    uint64_t *result;        // [0] length, [1] phantom, [2] error flags
    uint64_t *status;        // [0] ticket, [1..] look-back words
    uint32_t *scratch;       // per CTA: MAXB u32 match records
    uint32_t lead;           // 1 if this shard writes the frame byte
    uint32_t frame_byte;
    uint16_t *dump_len;      // optional single-block stage dump
    uint16_t *dump_dist;
    unsigned long long *phase_cycles; // optional: per-phase SM cycles summed over CTAs (profiling aid)
};

__device__ __forceinline__ uint32_t load32u(const uint32_t *w, uint32_t a)
{
    uint32_t i = a >> 2, s = (a & 3) * 8;
    return __funnelshift_r(w[i], w[i + 1], s);
}
__device__ __forceinline__ uint32_t hash4(uint32_t key) { return (key * 2654435761u) >> (32 - HASH_BITS); }
__device__ __forceinline__ uint32_t hash8(uint32_t k0, uint32_t k1)
{
    return ((k0 * 2654435761u) ^ (k1 * 2246822519u) ^ ((k1 * 3266489917u) >> 15)) >> (32 - HASH_BITS);
}
template <int GRAM>
__device__ __forceinline__ uint32_t hash_at(const uint32_t *w, uint32_t p)
{
    if (GRAM == 4) return hash4(load32u(w, p));
    return hash8(load32u(w, p), load32u(w, p + 4));
}

// Longest common prefix of data[c..] and data[p..], p-side first 8 bytes given; result clamped to cap.
__device__ __forceinline__ uint32_t lcp_from(const uint32_t *w, uint32_t c, uint32_t p, uint32_t P0, uint32_t P1,
                                             uint32_t cap)
{
    uint32_t ci = c >> 2, cs = (c & 3) * 8;
    uint32_t a0 = w[ci], a1 = w[ci + 1], a2 = w[ci + 2];
    uint32_t x = __funnelshift_r(a0, a1, cs) ^ P0;
    if (x) return min((uint32_t)(__ffs(x) - 1) >> 3, cap);
    x = __funnelshift_r(a1, a2, cs) ^ P1;
    if (x) return min(4u + ((uint32_t)(__ffs(x) - 1) >> 3), cap);
    uint32_t l = 8;
    uint32_t pi = p >> 2, ps = (p & 3) * 8;
    uint32_t cw = a2, pw = w[pi + 2]; // words straddling offset 8 on either side
    while (l < cap) {
        const uint32_t c1 = w[ci + (l >> 2) + 1], c2 = w[ci + (l >> 2) + 2];
        const uint32_t p1 = w[pi + (l >> 2) + 1], p2 = w[pi + (l >> 2) + 2];
        x = __funnelshift_r(cw, c1, cs) ^ __funnelshift_r(pw, p1, ps);
        if (x) return min(l + ((uint32_t)(__ffs(x) - 1) >> 3), cap);
        x = __funnelshift_r(c1, c2, cs) ^ __funnelshift_r(p1, p2, ps);
        if (x) return min(l + 4 + ((uint32_t)(__ffs(x) - 1) >> 3), cap);
        cw = c2;
        pw = p2;
        l += 8;
    }
    return cap;
}

// ---- sequence sizing, SURVEY.md A.3 / LZ4.c:540-575 -------------------------------------------------
struct SeqSize {
    uint32_t byte_size; // what the reference stores in the size field (before u16 truncation)
    uint32_t payload;   // bytes actually written
};
__device__ __forceinline__ uint32_t lit_ext_count(uint32_t lit) { return lit < 15 ? 0u : ((((lit - 15) & 0xFF) == 255) ? 2u : 1u); }
__device__ __forceinline__ SeqSize seq_size(uint32_t lit, uint32_t ml)
{
    SeqSize s;
    uint32_t adj = (ml - 4) & 0xFF;              // uint8_t adjusted_match_length, LZ4.c:562
    uint32_t mext = adj >= 15 ? 1u : 0u;         // counted in byte_size, LZ4.c:564-575
    uint32_t written = (ml >= 4 && mext) ? 1u : 0u; // actually written, LZ4.c:393-411
    s.byte_size = lit + 5 + lit_ext_count(lit) + mext;
    s.payload = s.byte_size - (mext - written);
    return s;
}

template <int TEAM>
__global__ void __launch_bounds__(THREADS, 1) lz4_encode_kernel(Params P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *data = smem + SM_DATA;
    const uint32_t *dataw = reinterpret_cast<const uint32_t *>(data);
    uint16_t *S = reinterpret_cast<uint16_t *>(smem + SM_S);
    uint32_t *dirw = reinterpret_cast<uint32_t *>(smem + SM_DIR);
    const uint16_t *dir16 = reinterpret_cast<const uint16_t *>(dirw);
    uint8_t *step = smem + SM_STEP;
    uint8_t *flag = smem + SM_FLAG;
    uint8_t *entry = smem + SM_ENTRY;
    Misc &M = *reinterpret_cast<Misc *>(smem + SM_MISC);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint32_t *R = P.scratch + (size_t)blockIdx.x * MAXB;

    long long t_prev = clock64();
#define LJB_PHASE(idx)                                                                      \
    do {                                                                                    \
        if (P.phase_cycles && tid == 0) {                                                   \
            long long t_now = clock64();                                                    \
            atomicAdd(&P.phase_cycles[idx], (unsigned long long)(t_now - t_prev));          \
            t_prev = t_now;                                                                 \
        }                                                                                   \
    } while (0)
    for (;;) {
        // ---------------- ticket ----------------
        if (tid == 0) M.ticket = (long long)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        __syncthreads();
        const long long b = M.ticket;
        if (b >= (long long)P.nblocks) break;
        const size_t boff = (size_t)b * P.block_len;
        const uint32_t nb = (uint32_t)min((size_t)P.block_len, P.n - boff);
        const uint8_t *src = P.in + boff;

        // ---------------- P1: stage block in shared memory ----------------
        {
            uint32_t head = 0;
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
                uint4 *d4 = reinterpret_cast<uint4 *>(data);
                const uint32_t n16 = nb >> 4;
                for (uint32_t i = tid; i < n16; i += THREADS) d4[i] = __ldg(&s4[i]);
                head = n16 << 4;
            }
            for (uint32_t i = head + tid; i < nb; i += THREADS) data[i] = __ldg(&src[i]);
            for (uint32_t i = nb + tid; i < ((nb + 63) & ~15u) + 16; i += THREADS) data[i] = 0; // defined bytes past the end
            for (int i = tid; i < NBUCKET / 2 + 1; i += THREADS) dirw[i] = 0;
        }
        __syncthreads();
        LJB_PHASE(0); // stage

        const uint32_t npos = nb >= 4 ? nb - 3 : 0; // positions that still have a 4-gram inside the block

        // ---------------- P2/P3 ----------------
        uint32_t *longbits = reinterpret_cast<uint32_t *>(smem + SM_LONG);
        for (int i = tid; i < MAXB / 32; i += THREADS) longbits[i] = 0;

        // index = counting sort of positions [0, cnt) by the hash of their GRAM-gram.  Afterwards
        // dir16[h] = end of bucket h = start of bucket h+1; inside a bucket, entries of an earlier
        // 1024-position chunk come first (the scatter runs in position-ordered rounds).
        auto build_index = [&](auto gram_tag, uint32_t cnt) {
            constexpr int GRAM = decltype(gram_tag)::value;
            for (uint32_t p = tid; p < cnt; p += THREADS) {
                uint32_t h = hash_at<GRAM>(dataw, p);
                atomicAdd(&dirw[h >> 1], (h & 1) ? 0x10000u : 1u);
            }
            __syncthreads();
            {
                // exclusive scan of 8192 u16 counts; thread t owns buckets 8t .. 8t+7 (4 packed words)
                uint32_t c[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t wv = dirw[tid * 4 + k];
                    c[2 * k] = wv & 0xFFFF;
                    c[2 * k + 1] = wv >> 16;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) sum += c[k];
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) M.scan_tmp[warp] = inc;
                __syncthreads();
                uint32_t wbase = 0;
                for (int k = 0; k < warp; ++k) wbase += M.scan_tmp[k];
                uint32_t run = wbase + inc - sum;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t lo = run;
                    run += c[2 * k];
                    uint32_t hi = run;
                    run += c[2 * k + 1];
                    dirw[tid * 4 + k] = (lo & 0xFFFF) | (hi << 16);
                }
            }
            __syncthreads();
            for (uint32_t base = 0; base < cnt; base += THREADS) {
                uint32_t p = base + tid;
                if (p < cnt) {
                    uint32_t h = hash_at<GRAM>(dataw, p);
                    uint32_t old = atomicAdd(&dirw[h >> 1], (h & 1) ? 0x10000u : 1u);
                    uint32_t slot = (h & 1) ? (old >> 16) : (old & 0xFFFF);
                    S[slot] = (uint16_t)p;
                }
                __syncthreads();
            }
        };

        build_index(std::integral_constant<int, 4>{}, npos);
        LJB_PHASE(1); // index (4-gram)

        // ---- phase A: sorted-order walk, first 8 bytes only
        for (uint32_t j = tid; j < npos; j += THREADS) {
            const uint32_t p = S[j];
            const uint32_t P0 = load32u(dataw, p), P1 = load32u(dataw, p + 4);
            const uint32_t h = hash4(P0);
            const uint32_t lo = h ? dir16[h - 1] : 0u, hi = dir16[h];
            const uint32_t cap8 = min(8u, nb - p);
            const uint32_t pch = p >> 10;
            uint32_t best = 0;
            for (uint32_t i = lo; i < hi; ++i) {
                const uint32_t c = S[i];
                const uint32_t cch = c >> 10;
                if (cch > pch) break; // only later positions from here on
                // an 8-byte (or block-end) match from an earlier chunk cannot be beaten by later positions
                if ((best >> 16) == cap8 && cch > ((0xFFFFu - (best & 0xFFFFu)) >> 10)) break;
                if (c < p) {
                    const uint32_t ci = c >> 2, cs = (c & 3) * 8;
                    const uint32_t a0 = dataw[ci], a1 = dataw[ci + 1];
                    if (__funnelshift_r(a0, a1, cs) == P0) {
                        const uint32_t x = __funnelshift_r(a1, dataw[ci + 2], cs) ^ P1;
                        uint32_t l = x ? 4u + ((uint32_t)(__ffs(x) - 1) >> 3) : 8u;
                        l = min(l, cap8);
                        best = max(best, (l << 16) | (0xFFFFu - c));
                    }
                }
            }
            const uint32_t bl = best >> 16, bp = 0xFFFFu - (best & 0xFFFFu);
            R[p] = bl ? ((bl << 16) | bp) : 0u;
            if (bl == 8 && nb - p > 8) atomicOr(&longbits[p >> 5], 1u << (p & 31)); // may be longer: phase B decides
        }
        for (uint32_t p = npos + tid; p < nb; p += THREADS) R[p] = 0; // the last 3 positions cannot start a match
        __syncthreads();
        LJB_PHASE(2); // search phase A

        // ---- phase B: positions with an >= 8 byte match, from the 8-gram index
        {
            uint32_t any = 0;
            for (int i = tid; i < MAXB / 32; i += THREADS) any |= longbits[i];
            any = __syncthreads_or(any != 0);
            if (any) {
                for (int i = tid; i < NBUCKET / 2 + 1; i += THREADS) dirw[i] = 0;
                __syncthreads();
                const uint32_t npos8 = nb >= 8 ? nb - 7 : 0;
                build_index(std::integral_constant<int, 8>{}, npos8);
                LJB_PHASE(3); // index (8-gram)
                // thread t owns the 32-position chunks t, t + 1024: one word of longbits each
                for (uint32_t ch = tid; ch * 32 < nb; ch += THREADS) {
                    uint32_t bits = longbits[ch];
                    uint32_t prev_len = 0, prev_pos = 0, prev_p = 0xFFFFFFFFu;
                    bool prev_capped = false;
                    while (bits) {
                        const uint32_t p = ch * 32 + (__ffs(bits) - 1);
                        bits &= bits - 1;
                        const uint32_t cap = min((uint32_t)MAX_MATCH, nb - p);
                        const uint32_t P0 = load32u(dataw, p), P1 = load32u(dataw, p + 4);
                        uint32_t bestkey = 0;
                        if (prev_p + 1 == p && prev_len >= 5) { // inherited candidate on the previous best's diagonal
                            const uint32_t c0 = prev_pos + 1;
                            uint32_t l0 = prev_len - 1;
                            if (prev_capped) // the previous run was cut by the cap, not by a mismatch: it may go on
                                while (l0 < cap && data[c0 + l0] == data[p + l0]) ++l0;
                            l0 = min(l0, cap);
                            bestkey = (l0 << 16) | (0xFFFFu - c0);
                        }
                        // a capped 1024 match becomes a literal step ((uint8_t)1024 == 0, LZ4.c:317) whose distance
                        // is never used: no need to look for an earlier one
                        if ((bestkey >> 16) != (uint32_t)MAX_MATCH) {
                            const uint32_t h = hash8(P0, P1);
                            const uint32_t lo = h ? dir16[h - 1] : 0u, hi = dir16[h];
                            const uint32_t pch = p >> 10;
                            for (uint32_t k = lo; k < hi; ++k) {
                                const uint32_t c = S[k];
                                const uint32_t cch = c >> 10;
                                if (cch > pch) break;
                                const uint32_t bl = bestkey >> 16, bp = 0xFFFFu - (bestkey & 0xFFFFu);
                                if (bl == cap && cch > (bp >> 10)) break; // later positions cannot win a tie
                                if (c < p) {
                                    // a candidate matters only if it beats the best length, or ties it from an earlier position
                                    const uint32_t need = bl == 0 ? 8u : (c < bp ? bl : bl + 1);
                                    if (need <= cap) {
                                        bool ok = true;
                                        if (bl != 0) ok = data[c + need - 1] == data[p + need - 1];
                                        if (ok) {
                                            const uint32_t l = lcp_from(dataw, c, p, P0, P1, cap);
                                            if (l >= 8) bestkey = max(bestkey, (l << 16) | (0xFFFFu - c));
                                            if (l == (uint32_t)MAX_MATCH) break;
                                        }
                                    }
                                }
                            }
                        }
                        const uint32_t bl = bestkey >> 16, bp = 0xFFFFu - (bestkey & 0xFFFFu);
                        prev_len = bl;
                        prev_pos = bp;
                        prev_p = p;
                        prev_capped = bl == cap;
                        R[p] = (bl << 16) | bp; // bl >= 8 here: phase A proved an 8-byte match exists
                    }
                }
            }
        }
        __syncthreads();
        LJB_PHASE(4); // search phase B
        if (P.dump_len) { // stage dump for parity tests of the search (single block calls only)
            for (uint32_t p = tid; p < nb; p += THREADS) {
                uint32_t r = R[p];
                P.dump_len[p] = (uint16_t)(r >> 16);
                P.dump_dist[p] = (r >> 16) ? (uint16_t)(p - (r & 0xFFFF)) : 0;
            }
        }

        // ---------------- P4: parse ----------------
        // step[p] = (uint8_t) best if best >= 4 else 0 (LZ4.c:314-321): 0 means "literal step"
        for (uint32_t p = tid; p < nb; p += THREADS) step[p] = (uint8_t)(R[p] >> 16);
        for (int i = tid; i < 1024; i += THREADS) entry[i] = 0xFF;
        __syncthreads();
        const uint32_t nseg = (nb + SEG - 1) / SEG;
        // pass A: x1[p] = (first chain position >= segment end) - segment end, for every p (walk descending)
        if ((uint32_t)tid < nseg) {
            const uint32_t s0 = tid * SEG, s1 = min(s0 + SEG, nb);
            for (uint32_t p = s1; p-- > s0;) {
                uint32_t st = step[p];
                uint32_t t = p + (st ? st : 1u);
                flag[p] = (uint8_t)(t >= s1 ? t - s1 : flag[t]);
            }
        }
        __syncthreads();
        // pass C: one thread hops segment to segment and records where the chain enters each one
        if (tid == 0) {
            uint32_t pos = 0;
            while (pos < nb) {
                uint32_t s = pos / SEG;
                entry[s] = (uint8_t)(pos - s * SEG);
                uint32_t s1 = min((s + 1) * SEG, nb);
                pos = s1 + flag[pos];
            }
        }
        __syncthreads();
        // pass E: mark the chain inside each segment: 0 not visited, 1 literal step, 2 match start
        if ((uint32_t)tid < nseg) {
            const uint32_t s0 = tid * SEG, s1 = min(s0 + SEG, nb);
            for (uint32_t p = s0; p < s1; ++p) flag[p] = 0;
            if (entry[tid] != 0xFF) {
                uint32_t p = s0 + entry[tid];
                while (p < s1) {
                    uint32_t st = step[p];
                    flag[p] = st ? 2 : 1;
                    p += st ? st : 1u;
                }
            }
        }
        __syncthreads();

        LJB_PHASE(5); // parse: chain resolution
        // sequence pass 1: last match end per warp region
        const uint32_t r0 = warp * REGION;
        {
            uint32_t le = 0;
            for (uint32_t p = r0 + lane; p < min(r0 + REGION, nb); p += 32)
                if (flag[p] == 2) le = max(le, p + step[p]);
            le = __reduce_max_sync(0xffffffffu, le);
            if (lane == 0) M.warp_last_end[warp] = le;
        }
        __syncthreads();
        uint32_t start_end = 0; // end of the last match before this warp's region
        for (int k = 0; k < warp; ++k) start_end = max(start_end, M.warp_last_end[k]);

        // sequence pass 2 (sizes) and pass 3 (emit) share one walker
        auto walk = [&](bool emit, unsigned long long out_base) {
            uint32_t carry_end = start_end;
            unsigned long long bytes = 0, sizes = 0;
            uint32_t nseq = 0, phantom = 0;
            const uint32_t rend = min(r0 + REGION, nb);
            for (uint32_t q = r0; q < rend; q += 32) {
                const uint32_t p = q + lane;
                const bool isM = (p < rend) && flag[p] == 2;
                const unsigned m = __ballot_sync(0xffffffffu, isM);
                if (m == 0) continue;
                const uint32_t ml = isM ? step[p] : 0;
                const uint32_t end = p + ml;
                const unsigned lower = m & ((1u << lane) - 1u);
                const int srcl = lower ? 31 - __clz(lower) : 0;
                const uint32_t pe = __shfl_sync(0xffffffffu, end, srcl);
                const uint32_t prev_end = lower ? pe : carry_end;
                const uint32_t lit = isM ? p - prev_end : 0;
                SeqSize sz = {0, 0};
                if (isM) sz = seq_size(lit, ml);
                uint32_t inc = sz.payload;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
                if (emit) {
                    unsigned lits_long = __ballot_sync(0xffffffffu, isM && lit > 16);
                    uint8_t *dst = P.out + out_base + bytes + (inc - sz.payload);
                    uint32_t lit_dst_off = 0;
                    if (isM) {
                        const uint32_t tok_lit = lit >= 15 ? 15u : lit;
                        const uint32_t tok_m = ml >= 19 ? 15u : ((ml - 4) & 0xFF);
                        uint32_t o = 0;
                        dst[o++] = (uint8_t)((tok_lit << 4) | tok_m);
                        dst[o++] = (uint8_t)(sz.byte_size & 0xFF);
                        dst[o++] = (uint8_t)((sz.byte_size >> 8) & 0xFF);
                        if (lit >= 15) {
                            uint32_t rem = (lit - 15) & 0xFF;
                            if (rem == 255) { dst[o++] = 255; rem = 0; }
                            dst[o++] = (uint8_t)rem;
                        }
                        lit_dst_off = o;
                        if (lit <= 16)
                            for (uint32_t k = 0; k < lit; ++k) dst[o + k] = data[prev_end + k];
                        o += lit;
                        const uint32_t dist = p - (R[p] & 0xFFFF);
                        dst[o++] = (uint8_t)(dist & 0xFF);
                        dst[o++] = (uint8_t)(dist >> 8);
                        if (ml >= 19) dst[o++] = (uint8_t)(ml - 19);
                    }
                    while (lits_long) { // long literal runs: the whole warp copies them
                        const int L = __ffs(lits_long) - 1;
                        lits_long &= lits_long - 1;
                        const uint32_t n_l = __shfl_sync(0xffffffffu, lit, L);
                        const uint32_t s_l = __shfl_sync(0xffffffffu, prev_end, L);
                        const unsigned long long d_l =
                            __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(dst + lit_dst_off), L);
                        uint8_t *dp = reinterpret_cast<uint8_t *>((uintptr_t)d_l);
                        for (uint32_t k = lane; k < n_l; k += 32) dp[k] = data[s_l + k];
                    }
                }
                bytes += tot;
                sizes += __reduce_add_sync(0xffffffffu, sz.byte_size);
                nseq += __popc(m);
                phantom += __popc(__ballot_sync(0xffffffffu, isM && sz.byte_size != sz.payload));
                carry_end = __shfl_sync(0xffffffffu, end, 31 - __clz(m));
            }
            if (!emit && lane == 0) {
                M.warp_bytes[warp] = bytes;
                M.warp_sizes[warp] = sizes;
                M.warp_nseq[warp] = nseq;
                M.warp_phantom[warp] = phantom;
            }
        };
        walk(false, 0);
        __syncthreads();

        // block totals (every thread computes them redundantly from 32 warp entries)
        unsigned long long pay = 3, sizes = 3, my_prefix = 3;
        uint32_t nseq = 0, phantom = 0, last_end = 0;
        for (int k = 0; k < NWARPS; ++k) {
            if (k == warp) my_prefix = pay;
            pay += M.warp_bytes[k];
            sizes += M.warp_sizes[k];
            nseq += M.warp_nseq[k];
            phantom += M.warp_phantom[k];
            last_end = max(last_end, M.warp_last_end[k]);
        }
        // trailing literals (LZ4.c:585-613); literal_counter is uint16_t (LZ4.c:514) so 65536 wraps to "none"
        const uint32_t tlit = (nb - last_end) & 0xFFFF;
        const unsigned long long trail_off = pay;
        uint32_t tsize = 0;
        if (tlit) {
            tsize = tlit + 5 + lit_ext_count(tlit);
            pay += tsize;
            sizes += tsize;
            nseq += 1;
        }

        LJB_PHASE(6); // sequence sizing
        // ---------------- P5: place (decoupled look-back) ----------------
        if (warp == 0) {
            unsigned long long base = ljb_lookback(P.status + 1, b, pay, P.lead);
            if (lane == 0) {
                M.base = base;
                M.emit_ok = (base + pay <= P.out_cap) ? 1 : 0;
                P.block_offsets[b] = base;
                if (phantom) atomicAdd((unsigned long long *)&P.result[1], (unsigned long long)phantom);
                if (b == (long long)P.nblocks - 1) {
                    P.block_offsets[P.nblocks] = base + pay;
                    P.result[0] = base + pay;
                }
                if (!M.emit_ok) atomicOr((unsigned long long *)&P.result[2], 1ull);
            }
        }
        __syncthreads();

        LJB_PHASE(7); // look-back
        // ---------------- P6: emit ----------------
        if (M.emit_ok) {
            const unsigned long long base = M.base;
            walk(true, base + my_prefix);
            if (tid == 0) {
                uint8_t *hdr = P.out + base;
                hdr[0] = (uint8_t)(nseq & 0xFF);          // LZ4.c:615, :417
                hdr[1] = (uint8_t)(sizes & 0xFF);         // LZ4.c:617, :419 (low 16 bits)
                hdr[2] = (uint8_t)((sizes >> 8) & 0xFF);
                if (b == 0 && P.lead) P.out[0] = (uint8_t)P.frame_byte; // LZ4.c:429
            }
            if (tlit) {
                uint8_t *dst = P.out + base + trail_off;
                uint32_t hdrlen = 3 + lit_ext_count(tlit);
                if (tid == 0) {
                    uint32_t o = 0;
                    dst[o++] = (uint8_t)((tlit >= 15 ? 15u : tlit) << 4);
                    dst[o++] = (uint8_t)(tsize & 0xFF);
                    dst[o++] = (uint8_t)((tsize >> 8) & 0xFF);
                    if (tlit >= 15) {
                        uint32_t rem = (tlit - 15) & 0xFF;
                        if (rem == 255) { dst[o++] = 255; rem = 0; }
                        dst[o++] = (uint8_t)rem;
                    }
                    dst[hdrlen + tlit] = 0;     // match_offset = 0 (LZ4.c:587)
                    dst[hdrlen + tlit + 1] = 0;
                }
                // literals start where the counter was last reset; after a uint16 wrap that is the wrap point
                const uint32_t lsrc = nb - tlit;
                for (uint32_t k = tid; k < tlit; k += THREADS) dst[hdrlen + k] = data[lsrc + k];
            }
        }
        __syncthreads(); // region B and data are reused by the next block
        LJB_PHASE(8); // emit
    }
#undef LJB_PHASE
}

} // namespace lz4k

// ---- host side ---------------------------------------------------------------------------------------
extern "C" size_t ljb_lz4_block_count(size_t n, size_t block_len) { return block_len ? (n + block_len - 1) / block_len : 0; }

extern "C" size_t ljb_lz4_bound(size_t n, size_t block_len)
{
    // worst case of the dialect: a 5-byte sequence header per input byte (ml = 1 phantom matches) + literals
    return 1 + 3 * ljb_lz4_block_count(n, block_len) + 6 * n + 64;
}

static int lz4_launch(ljb_ctx *ctx, const uint8_t *d_in, size_t n, size_t block_len, uint8_t *d_out, size_t out_cap,
                      uint64_t *d_block_offsets, uint64_t *d_result, size_t first_block, size_t frame_blocks,
                      uint16_t *d_dump_len, uint16_t *d_dump_dist)
{
    using namespace lz4k;
    if (!ctx || !d_in || !d_out || !d_block_offsets || !d_result || n == 0 || block_len == 0 || block_len > MAXB)
        return LJB_E_ARG;
    const size_t nblocks = ljb_lz4_block_count(n, block_len);
    if (nblocks > 0x7fffffffull) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    const int grid = (int)((nblocks < (size_t)ctx->num_sms) ? nblocks : (size_t)ctx->num_sms);
    int rc;
    if ((rc = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, (size_t)ctx->num_sms * MAXB * sizeof(uint32_t))) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_status, &ctx->status_bytes, (nblocks + 2 + 16) * sizeof(uint64_t))) != 0) return rc;
    LJB_CUDA(cudaMemsetAsync(ctx->d_status, 0, (nblocks + 2 + 16) * sizeof(uint64_t), ctx->stream));
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    Params P;
    P.in = d_in;
    P.n = n;
    P.block_len = (uint32_t)block_len;
    P.nblocks = (uint32_t)nblocks;
    P.out = d_out;
    P.out_cap = out_cap;
    P.block_offsets = d_block_offsets;
    P.result = d_result;
    P.status = (uint64_t *)ctx->d_status;
    P.scratch = (uint32_t *)ctx->d_scratch;
    P.lead = first_block == 0 ? 1u : 0u;
    P.frame_byte = (uint32_t)(frame_blocks & 0xFF);
    P.dump_len = d_dump_len;
    P.dump_dist = d_dump_dist;
    P.phase_cycles = nullptr;
    if (getenv("LJB_LZ4_PHASES")) { // profiling aid: per-phase cycle counters behind the status words
        P.phase_cycles = (unsigned long long *)ctx->d_status + (nblocks + 2);
    }
    static bool attr_done = false;
    if (!attr_done) {
        LJB_CUDA(cudaFuncSetAttribute(lz4_encode_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        attr_done = true;
    }
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    lz4_encode_kernel<8><<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    if (P.phase_cycles) {
        unsigned long long ph[16];
        LJB_CUDA(cudaMemcpyAsync(ph, P.phase_cycles, sizeof ph, cudaMemcpyDeviceToHost, ctx->stream));
        LJB_CUDA(cudaStreamSynchronize(ctx->stream));
        static const char *names[9] = {"stage", "index4", "phaseA", "index8", "phaseB", "parse", "sizing", "lookback", "emit"};
        unsigned long long tot = 0;
        for (int i = 0; i < 9; ++i) tot += ph[i];
        fprintf(stderr, "[ljb lz4 phases] cycles per block:");
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %s=%.0f(%.0f%%)", names[i], (double)ph[i] / (double)nblocks, 100.0 * (double)ph[i] / (double)tot);
        fprintf(stderr, " total=%.0f\n", (double)tot / (double)nblocks);
    }
    return LJB_OK;
}

extern "C" int ljb_lz4_compress_dev(ljb_ctx *ctx, const uint8_t *d_in, size_t n, size_t block_len, uint8_t *d_out,
                                    size_t out_cap, uint64_t *d_block_offsets, uint64_t *d_result, size_t first_block,
                                    size_t frame_blocks)
{
    return lz4_launch(ctx, d_in, n, block_len, d_out, out_cap, d_block_offsets, d_result, first_block, frame_blocks, nullptr,
                      nullptr);
}

extern "C" int ljb_lz4_compress(ljb_ctx *ctx, const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                                uint64_t *block_offsets, size_t *out_len, uint64_t *phantom)
{
    if (!ctx || !in || !out || n == 0 || block_len == 0 || block_len > lz4k::MAXB) return LJB_E_ARG;
    const size_t nblocks = ljb_lz4_block_count(n, block_len);
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // device capacity: never more than the caller can take, never more than the dialect can produce
    size_t dcap = ljb_lz4_bound(n, block_len);
    if (out_cap < dcap) dcap = out_cap;
    if ((rc = ljb_ensure(&ctx->d_stage_in, &ctx->stage_in_bytes, n + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_stage_out, &ctx->stage_out_bytes, dcap + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, (nblocks + 1 + 3) * sizeof(uint64_t))) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;
    uint64_t *d_res = d_offs + nblocks + 1;
    LJB_CUDA(cudaMemcpyAsync(ctx->d_stage_in, in, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = ljb_lz4_compress_dev(ctx, (const uint8_t *)ctx->d_stage_in, n, block_len, (uint8_t *)ctx->d_stage_out, dcap, d_offs,
                              d_res, 0, nblocks);
    if (rc != 0) return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, d_res, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    LJB_CUDA(cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    if (out_len) *out_len = (size_t)res[0];
    if (phantom) *phantom = res[1];
    if (res[2] & 1) return LJB_E_CAPACITY;
    LJB_CUDA(cudaMemcpyAsync(out, ctx->d_stage_out, (size_t)res[0], cudaMemcpyDeviceToHost, ctx->stream));
    if (block_offsets)
        LJB_CUDA(cudaMemcpyAsync(block_offsets, d_offs, (nblocks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}

extern "C" int ljb_lz4_block_matches(ljb_ctx *ctx, const uint8_t *in, size_t n, uint16_t *len, uint16_t *dist)
{
    if (!ctx || !in || !len || !dist || n == 0 || n > lz4k::MAXB) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t dcap = ljb_lz4_bound(n, n);
    if ((rc = ljb_ensure(&ctx->d_stage_in, &ctx->stage_in_bytes, n + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_stage_out, &ctx->stage_out_bytes, dcap + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, 8 * sizeof(uint64_t) + 4 * lz4k::MAXB)) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;
    uint64_t *d_res = d_offs + 2;
    uint16_t *d_len = (uint16_t *)(d_offs + 8);
    uint16_t *d_dist = d_len + lz4k::MAXB;
    LJB_CUDA(cudaMemcpyAsync(ctx->d_stage_in, in, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = lz4_launch(ctx, (const uint8_t *)ctx->d_stage_in, n, n, (uint8_t *)ctx->d_stage_out, dcap, d_offs, d_res, 0, 1, d_len,
                    d_dist);
    if (rc != 0) return rc;
    LJB_CUDA(cudaMemcpyAsync(len, d_len, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(dist, d_dist, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}
