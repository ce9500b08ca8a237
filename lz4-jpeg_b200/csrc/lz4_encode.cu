// lz4-jpeg_b200/csrc/lz4_encode.cu — LZ4 (reference dialect) block encoder for sm_100a.
//
// Replaces, per 64 KiB block, the reference's block_encode() -> find_longest_match() -> write_block()
// (Algorithms/sequential/LZ4/LZ4.c:506-620, :290-323, :365-425; thread-per-block form
// parallel_block_encode, Algorithms/parallel/LZ4/LZ4.c:518-628) and, across blocks, the serial
// concatenation of write_output() (LZ4.c:427-441).
//
// One persistent CTA (1024 threads) per SM pulls blocks from a ticket counter and runs, per block:
//   stage    : the block is copied once from HBM into shared memory (128-bit coalesced loads)
//   search   : exact longest-previous-match for EVERY position — equivalent to the reference's exhaustive scan
//              because a match of length >= 4 shares its 4-gram with the current position; ties resolved to the
//              earliest position (strict '>' in LZ4.c:307), length capped at min(1024, block end).
//              ladder  k = 8 .. 4: first occurrence of every k-gram by atomicMin into a hashed table; a position
//                      with an earlier occurrence of its k-gram and none of its (k+1)-gram is final
//              index   the positions with an >= 8 byte match and the first occurrences of their 8-grams, counting-
//                      sorted by an exact group id (rank of the first occurrence)
//              B1      sorted-order walk of the groups comparing bytes 8 .. 16: matches below 16 are final
//              B2      position order, each lane inheriting the previous position's match along its diagonal
//   parse    : the greedy chain 0 -> p + step[p] through per-segment exit tables
//   size     : match positions parked per segment, sequences sized by eight lanes per segment with the reference's
//              uint8/uint16 wrap rules, two block scans
//   publish  : decoupled look-back over the per-block byte counts
//   emit     : sequences serialised into shared memory, copied to a staging buffer with 16-byte stores
//   place    : one block later (in the shadow of the next block's chain hop) the block moves to its offset in the stream
// HBM traffic per block is N_in + N_out (+ per-CTA scratch that stays in L2: match records, group ids, staging).
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace lz4k {

constexpr int THREADS = 1024;
constexpr int NWARPS = THREADS / 32;
constexpr int MAXB = 65536;          // LJB_LZ4_MAX_BLOCK
constexpr int HASH_BITS = 13;
constexpr int NBUCKET = 1 << HASH_BITS;
constexpr int SEG = 132;             // parse segment: 33 words, so per-thread segment walks are bank-conflict free
constexpr int MAXSEG = (MAXB + SEG - 1) / SEG; // 497
constexpr int SUP = 16;              // segments per super-segment of the two-level chain hop (32 super-segments at most)
constexpr int SLOTS = 27;            // match positions a segment parks in shared memory for the eight-lane passes (the rest: its own thread)
constexpr int MAX_MATCH = 1024;      // LZ4.c:20
constexpr int CHUNK = 16;            // positions a lane of B2 works through in sequence (a chain of carried pairs never crosses a chunk)
constexpr int CH = 12;               // the index keeps the entries of a group ordered by 4096-position chunk (one scatter round each)
constexpr int B1_BUDGET = 512;       // positions inside a chain leave groups larger than this to B2

// ---- shared memory map (bytes) ---------------------------------------------------------------------
constexpr int SM_DATA = 0;                         // 65536 + 64 pad
constexpr int SM_B = MAXB + 64;                    // region B
// index / search view of region B
constexpr int SM_S = SM_B;                         // u16 S[65536]          sorted positions
constexpr int SM_DIR = SM_S + 2 * MAXB;            // u32 dirw[4096 + 1]    packed u16 bucket ends
// parse view of region B
constexpr int SM_STEP = SM_B;                      // u8 step[65536 + 64]
constexpr int SM_ENTRY = SM_STEP + MAXB + 64;      // u8 entry[1024]
constexpr int SM_FLAG = SM_ENTRY + 1024;           // u8 exit table[65536 + 64]; dead once the chain's segment entries are known, then:
constexpr int SM_EXIT2 = SM_FLAG + MAXB + 64;       // u8 exit2[32][256] + u16 supentry[32]: behind the exit table, parse only
constexpr int SM_SEGINFO = SM_FLAG;                // uint4 seginfo[512]            per segment: previous match end, byte offset, payload, size sum
constexpr int SM_SEGCNT = SM_SEGINFO + 16 * 512;   // u8 segcnt[512], u8 segph[512]
constexpr int SM_SLOT = SM_SEGCNT + 1024;          // u16 slots[MAXSEG * SLOTS]     positions of the segment's first matches
constexpr int SM_OUT = (SM_SLOT + 2 * 512 * SLOTS + 15) & ~15; // the encoded block, up to SOUT_CAP bytes (to the end of region B)
constexpr int SM_LONG = SM_DIR + 4 * (NBUCKET / 2 + 4); // u32 longbits[2048]: positions that have an >= 8 byte match
constexpr int SM_FIRST = SM_LONG + MAXB / 8;             // u32 firstbits[2048]: first occurrences of a repeated 8-gram
constexpr int SM_MISC = SM_FIRST + MAXB / 8;
constexpr int SM_PREF = SM_LONG - (2 * (MAXB / 32) + 16); // u16 pref[2048]: first occurrences before each 32-position word
constexpr int GIDX_BYTES = SM_PREF - SM_S;             // what S and the group directory share
constexpr int SM_TOTAL = SM_MISC + 1024;
constexpr int SOUT_CAP = ((SM_MISC - SM_OUT) & ~15) - 16; // larger blocks are encoded straight into the global staging buffer
static_assert(SM_EXIT2 + 32 * 256 + 64 <= SM_MISC, "parse view must fit inside region B");
static_assert((MAXSEG + SUP - 1) / SUP <= 32 && SUP * SEG > 255, "one lane per super-segment; a hop cannot skip one");
static_assert(SOUT_CAP >= 56 * 1024, "the shared staging area should hold a typical encoded block");
static_assert(SM_SEGINFO % 16 == 0, "seginfo is a uint4 array");
static_assert(SEG <= 254 && MAXSEG <= 512, "entry offsets are bytes; one segment per thread");
static_assert(SM_TOTAL <= 227 * 1024, "exceeds B200 shared memory per CTA");

struct Misc {
    unsigned long long warp_x[NWARPS];       // per warp of segments: payload bytes | sum of byte_size fields << 32
    unsigned long long warp_y[NWARPS];       // per warp of segments: sequences | phantom sequences << 32
    unsigned int scan_tmp[NWARPS];
    long long ticket;
    unsigned long long base;                 // output offset of this block
    int emit_ok;
    unsigned int wdone[NWARPS];              // chain search: ticket + 1 of the block whose first walks warp w has finished
};
static_assert(sizeof(Misc) <= 1024, "Misc must fit its area");

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
    uint32_t *idx8;          // per CTA: NBUCKET u32 + MAXB u16: the 8-gram index of low-entropy blocks (lz4_lazy.cuh)
    uint16_t *gids;          // per CTA: MAXB u16 group ids (dense rank of the 8-gram's first occurrence)
    uint8_t *staging;        // per CTA: two buffers of stage_stride bytes holding the encoded block until its offset is known
    size_t stage_stride;
    uint64_t offs_bias;      // added to every block_offsets entry (base of this shard in a larger stream)
    uint32_t lead;           // 1 if this shard writes the frame byte
    uint32_t frame_byte;
    uint16_t *dump_len;      // optional single-block stage dump
    uint16_t *dump_dist;
    uint32_t tune;           // experiment switches (LJB_LZ4_TUNE), 0 in production
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

#include "lz4_lazy.cuh"
#include "lz4_small.cuh"

template <int MODE> // 0: every position is searched (ladder / group index / B1 / B2); 1: only the positions the greedy parse visits (lz4_lazy.cuh)
__global__ void __launch_bounds__(THREADS, 1) lz4_encode_kernel(Params P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *data = smem + SM_DATA;
    const uint32_t *dataw = reinterpret_cast<const uint32_t *>(data);
    uint16_t *S = reinterpret_cast<uint16_t *>(smem + SM_S);
    uint8_t *step = smem + SM_STEP;
    uint8_t *flag = smem + SM_FLAG;
    uint8_t *entry = smem + SM_ENTRY;
    Misc &M = *reinterpret_cast<Misc *>(smem + SM_MISC);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint32_t *R = P.scratch + (size_t)blockIdx.x * MAXB; // match records, one per position
    // The encoded block is written to a staging buffer first; its offset in the stream is fetched (decoupled
    // look-back) only after the NEXT block has been encoded, when the predecessors have long published their
    // sizes, so the look-back practically never waits.  (Fetching it right away cost 12 % of the kernel in
    // barrier stalls: every CTA waited for the slowest of the 147 blocks in flight before it.)
    uint8_t *const stage0 = P.staging + (size_t)blockIdx.x * 2 * P.stage_stride;
    int buf = 0;
    bool pend = false;
    long long pend_b = 0;
    unsigned long long pend_pay = 0;
    // moves the pending block from its staging buffer to its place in the stream.  Done by warps fw .. 31: by all of them
    // (fw = 0) after the last block, by warps 1 .. 31 (a named barrier among them) while thread 0 hops along the parse
    // chain of the next block — the one stretch of the kernel in which the other 1023 threads would only wait.
    auto flush_pending = [&](int fw) {
        const int nthr = THREADS - 32 * fw, t = tid - 32 * fw;
        auto group_sync = [&]() {
            if (fw == 0) __syncthreads();
#ifdef LJB_EMU_BUILD
            else emu_named_barrier((unsigned)nthr);
#else
            else asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
#endif
        };
        if (warp == fw) {
            const unsigned long long base = ljb_lookback_resolve(P.status + 1, pend_b, pend_pay, P.lead);
            if (lane == 0) {
                M.base = base;
                M.emit_ok = (base + pend_pay <= P.out_cap) ? 1 : 0;
                P.block_offsets[pend_b] = base + P.offs_bias;
                if (pend_b == (long long)P.nblocks - 1) {
                    P.block_offsets[P.nblocks] = base + pend_pay + P.offs_bias;
                    P.result[0] = base + pend_pay;
                }
                if (!M.emit_ok) atomicOr((unsigned long long *)&P.result[2], 1ull);
            }
        }
        group_sync();
        if (M.emit_ok) {
            const uint8_t *src = stage0 + (size_t)(buf ^ 1) * P.stage_stride;
            const uint32_t *srcw = reinterpret_cast<const uint32_t *>(src);
            uint8_t *dst = P.out + M.base;
            const uint32_t n = (uint32_t)pend_pay;
            const uint32_t head = min((uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), n);
            if ((uint32_t)t < head) dst[t] = src[t];
            const uint32_t nwords = (n - head) >> 2;
            uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
            for (uint32_t j = t; j < nwords; j += nthr) __stcs(&d4[j], __funnelshift_r(srcw[j], srcw[j + 1], 8 * head)); // streaming store
            const uint32_t done = head + 4 * nwords;
            if ((uint32_t)t < n - done) dst[done + t] = src[done + t];
        }
        if (fw == 0) __syncthreads(); // M.base / M.emit_ok are reused (after a partial flush: the barrier behind the chain hop)
    };

    if (tid < NWARPS) M.wdone[tid] = 0; // (read after the first block's barriers; a block's own value is its ticket + 1, never 0)
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
        if (b >= (long long)P.nblocks) {
            if (pend) flush_pending(0);
            break;
        }
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
                for (uint32_t i = tid; i < n16; i += THREADS) d4[i] = __ldcs(&s4[i]); // streamed once: do not displace the scratch in L2
                head = n16 << 4;
            }
            for (uint32_t i = head + tid; i < nb; i += THREADS) data[i] = __ldg(&src[i]);
            for (uint32_t i = nb + tid; i < ((nb + 63) & ~15u) + 16; i += THREADS) data[i] = 0; // defined bytes past the end
        }
        __syncthreads();
        LJB_PHASE(0); // stage

        const uint32_t npos = nb >= 4 ? nb - 3 : 0; // positions that still have a 4-gram inside the block
        const uint32_t nseg = (nb + SEG - 1) / SEG;
        const uint32_t s0 = (uint32_t)tid * SEG, s1 = min(s0 + SEG, nb); // this thread's segment (tid < nseg)

        if constexpr (MODE == 1) {
            // ---------------- search + parse, only along the greedy chain (lz4_lazy.cuh) ----------------
            lazy_search(smem, nb, R, P, M, [&](int idx) { LJB_PHASE(idx); }, [&]() { if (pend) flush_pending(0); });
        } else {
        // ---------------- P2/P3 ----------------
        uint32_t *longbits = reinterpret_cast<uint32_t *>(smem + SM_LONG);
        uint32_t *firstbits = reinterpret_cast<uint32_t *>(smem + SM_FIRST);
        for (int i = tid; i < MAXB / 32; i += THREADS) {
            longbits[i] = 0;
            firstbits[i] = 0;
        }

        // ---- phase A: first-occurrence ladder, k = 8, 7, 6, 5, 4
        // Level k puts every still-unresolved position into a table keyed by a hash of its k-gram that keeps
        // the MINIMUM position (atomicMin).  A position whose slot holds an earlier position with the same
        // k-gram has found the first occurrence F_k(p) of that gram: the longest match is >= k, and because
        // no earlier position shares its (k+1)-gram (it would have been resolved one level up) every earlier
        // occurrence matches exactly k bytes, so (k, F_k(p)) IS the reference's answer (earliest of the
        // longest).  Resolved positions are never a first occurrence of any shorter gram either, so they
        // leave the ladder.  Slots owned by a different gram ("losers") are re-hashed in further rounds; all
        // occurrences of a gram win or lose together, which keeps the minimum exact.  Level 8 only flags:
        // matches of 8+ bytes get their true length in phase B.
        {
            uint32_t *T = reinterpret_cast<uint32_t *>(smem + SM_S); // 32768 slots
            constexpr int TBITS = 15;
            constexpr int MAXR = 12;
            unsigned long long unres = 0;
            // bit i of a thread's masks is position tid + 1024 * i.  (Giving every lane its own shared-memory bank instead —
            // position 2048 * warp + 128 * (i / 4) + 4 * lane + i % 4 — was measured 3 % slower: the position loads are not
            // what the ladder waits for, and the records are then stored 16 bytes apart.)
            auto pos_of = [&](int i) -> uint32_t { return (uint32_t)tid + 1024u * (uint32_t)i; };
            auto gram_at = [&](uint32_t q, uint32_t mask1, uint32_t &g0, uint32_t &g1) { // the 8 bytes at q (second word masked)
                const uint32_t qi = q >> 2, qs = (q & 3) * 8;
                const uint32_t w0 = dataw[qi], w1 = dataw[qi + 1], w2 = dataw[qi + 2];
                g0 = __funnelshift_r(w0, w1, qs);
                g1 = __funnelshift_r(w1, w2, qs) & mask1;
            };
            // positions tid + 1024 * i below a limit: the low ceil((limit - tid) / 1024) bits
            auto below = [&](uint32_t limit) -> unsigned long long {
                if (limit <= (uint32_t)tid) return 0ull;
                const uint32_t cnt = (limit - (uint32_t)tid + 1023u) >> 10;
                return cnt >= 64u ? ~0ull : ((1ull << cnt) - 1ull);
            };
            unres = below(nb);
            {
                // the table is filled once per block: a slot holds (generation << 16) | position with generations counting
                // DOWN, so an entry of the current round always beats what earlier rounds left behind (no clearing, one
                // barrier less per round)
                uint4 *T4 = reinterpret_cast<uint4 *>(T);
                const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                for (int i = tid; i < (1 << TBITS) / 4; i += THREADS) T4[i] = ff;
            }
            uint32_t gen = 0xFFFEu;
            __syncthreads();
#pragma unroll 1
            for (int k = 8; k >= 4; --k) {
                const uint32_t npk = nb >= (uint32_t)k ? nb - k + 1 : 0;
                const uint32_t m1 = k >= 8 ? 0xFFFFFFFFu : (k == 4 ? 0u : ((1u << (8 * (k - 4))) - 1u));
                unsigned long long ins = unres & below(npk);
                int round = 0;
                for (;;) {
                    const uint32_t tag = gen << 16; // (the barrier that ended the previous round separates its reads from these writes)
                    --gen;
                    const uint32_t A = 2654435761u + 0x9E3779B1u * (uint32_t)round * 2u;
                    const uint32_t B = 2246822519u + 0x85EBCA77u * (uint32_t)round * 2u;
                    if (P.phase_cycles) { // (one atomic per warp: per-thread atomics here tripled the ladder's time)
                        const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)__popcll(ins));
                        if (lane == 0) atomicAdd(&P.phase_cycles[21], (unsigned long long)tot);
                        if (tid == 0) atomicAdd(&P.phase_cycles[22], 1ull);
                    }
                    for (unsigned long long m = ins; m;) {
                        const int i = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const uint32_t p = pos_of(i);
                        uint32_t P0, P1;
                        gram_at(p, m1, P0, P1);
                        uint32_t h = (P0 * A) ^ (P1 * B);
                        h = ((h ^ (h >> 15)) * 2246822519u) >> (32 - TBITS);
                        atomicMin(&T[h], tag | p);
                    }
                    __syncthreads();
                    unsigned long long next = 0;
                    for (unsigned long long m = ins; m;) {
                        const int i = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const uint32_t p = pos_of(i);
                        uint32_t P0, P1;
                        gram_at(p, m1, P0, P1);
                        uint32_t h = (P0 * A) ^ (P1 * B);
                        h = ((h ^ (h >> 15)) * 2246822519u) >> (32 - TBITS);
                        const uint32_t q = T[h] & 0xFFFFu;
                        if (q != p) { // q < p: the earliest position that hashes here
                            uint32_t Q0, Q1;
                            gram_at(q, m1, Q0, Q1);
                            if (Q0 == P0 && Q1 == P1) {
                                R[p] = ((uint32_t)k << 16) | q;
                                unres &= ~(1ull << i);
                                if (k == 8 && nb - p > 8) {
                                    atomicOr(&longbits[p >> 5], 1u << (p & 31));
                                    atomicOr(&firstbits[q >> 5], 1u << (q & 31));
                                }
                            } else {
                                next |= 1ull << i; // slot owned by another gram: try again with another hash
                            }
                        }
                    }
                    ins = next;
                    ++round;
                    if (!__syncthreads_or(ins != 0)) break;
                    if (round >= MAXR) break;
                }
                // practically unreachable: grams that kept colliding through MAXR independent hashes
                for (unsigned long long m = ins; m;) {
                    const int i = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const uint32_t p = pos_of(i);
                    const uint32_t P0 = load32u(dataw, p), P1 = load32u(dataw, p + 4) & m1;
                    for (uint32_t q = 0; q < p; ++q) {
                        if (load32u(dataw, q) == P0 && (load32u(dataw, q + 4) & m1) == P1) {
                            R[p] = ((uint32_t)k << 16) | q;
                            unres &= ~(1ull << i);
                            if (k == 8 && nb - p > 8) {
                                atomicOr(&longbits[p >> 5], 1u << (p & 31));
                                atomicOr(&firstbits[q >> 5], 1u << (q & 31));
                            }
                            break;
                        }
                    }
                }
            }
            for (unsigned long long m = unres; m;) { // no earlier occurrence of even the 4-gram: literal
                const int i = __ffsll((long long)m) - 1;
                m &= m - 1;
                R[pos_of(i)] = 0;
            }
        }
        __syncthreads();
        LJB_PHASE(2); // search phase A (ladder)

        // ---- phase B: positions with an >= 8 byte match, from the 8-gram index
        {
            uint32_t any = 0;
            for (int i = tid; i < MAXB / 32; i += THREADS) any |= longbits[i];
            any = __syncthreads_or(any != 0);
            if (any) {
                // ---- group index.  Level 8 of the ladder left, for every flagged position, the FIRST occurrence q of its
                // 8-gram (R[p] & 0xFFFF) and a bit for every such q: the rank of q among the set bits is an exact, dense id
                // of the 8-gram ("group").  Counting sort of the flagged positions and the first occurrences by group id:
                // afterwards S holds every group contiguously (entries of an earlier 4096-position chunk first) and
                // gdir16[g] = end of group g = start of group g+1.  A position then visits only true occurrences of its
                // own 8-gram — no hash collisions, no 8-byte compare.  Layout inside the S/dir area: S (2 bytes x entries),
                // the directory right behind it (2 bytes x groups), the rank table at the end.  Should the directory not
                // fit (only when nearly every position is flagged), neighbouring groups share a bucket (gid >> shift)
                // and candidates are filtered by comparing the 8 bytes, as a hashed index would.
                const uint32_t npos8 = nb >= 8 ? nb - 7 : 0;
                uint16_t *const pref = reinterpret_cast<uint16_t *>(smem + SM_PREF);
                uint16_t *const gids = P.gids + (size_t)blockIdx.x * MAXB;
                uint32_t ngroups, nidx;
                {
                    const uint32_t f0 = firstbits[2 * tid], f1 = firstbits[2 * tid + 1];
                    const uint32_t cf = __popc(f0) + __popc(f1);
                    const uint32_t cl = __popc(longbits[2 * tid]) + __popc(longbits[2 * tid + 1]);
                    uint32_t inc = cf;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    const uint32_t wl = __reduce_add_sync(0xffffffffu, cl);
                    if (lane == 31) {
                        M.scan_tmp[warp] = inc;
                        M.warp_y[warp] = wl;
                    }
                    __syncthreads();
                    uint32_t before = 0, tot_f = 0, tot_l = 0;
                    for (int k = 0; k < NWARPS; ++k) {
                        if (k == warp) before = tot_f;
                        tot_f += M.scan_tmp[k];
                        tot_l += (uint32_t)M.warp_y[k];
                    }
                    pref[2 * tid] = (uint16_t)(before + inc - cf);
                    pref[2 * tid + 1] = (uint16_t)(before + inc - cf + __popc(f0));
                    ngroups = tot_f;
                    nidx = tot_f + tot_l;
                }
                // S holds the flagged positions only; the first occurrence of every group sits in a table of its own (it is a
                // candidate, never a worker: in S it would idle a third of the lanes of the sorted walk)
                const uint32_t nS = nidx - ngroups;
                const uint32_t sbytes = (2 * nS + 4 + 3) & ~3u; // S, with slack for the one-ahead reads
                const uint32_t fpbytes = (2 * ngroups + 3) & ~3u;
                uint32_t shift = 0;
                while (sbytes + fpbytes + 4 * (((ngroups >> shift) + 1) / 2 + 2) > (uint32_t)GIDX_BYTES) ++shift;
                const bool exact = shift == 0;
                uint16_t *const firstpos = reinterpret_cast<uint16_t *>(smem + SM_S + sbytes); // [ngroups]
                uint32_t *const gdirw = reinterpret_cast<uint32_t *>(smem + SM_S + sbytes + fpbytes);
                const uint16_t *const gdir16 = reinterpret_cast<const uint16_t *>(gdirw);
                const uint32_t dwords = ((ngroups >> shift) + 1) / 2 + 1;
                for (uint32_t i = tid; i < dwords; i += THREADS) gdirw[i] = 0;
                __syncthreads();
                auto rank_of = [&](uint32_t q) -> uint32_t { return (uint32_t)pref[q >> 5] + __popc(firstbits[q >> 5] & ((1u << (q & 31)) - 1u)); };
                // histogram; the group id of a flagged position is kept in a per-CTA array (it is needed three more times).
                // A thread takes four consecutive positions per 4096-position chunk: one word of each bit set, one 16-byte load
                // of the records (fetched one chunk ahead), one 8-byte store of the ids.
                {
                    auto bits4 = [&](const uint32_t *bits, uint32_t q0) -> uint32_t { return q0 < (uint32_t)MAXB ? (bits[q0 >> 5] >> (q0 & 31)) & 0xFu : 0u; };
                    const uint4 *R4 = reinterpret_cast<const uint4 *>(R);
                    uint32_t q0n = 4u * (uint32_t)tid;
                    uint32_t lwn = q0n < npos8 ? bits4(longbits, q0n) : 0u;
                    uint4 rn = lwn ? R4[q0n >> 2] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
                    for (uint32_t base = 0; base < npos8; base += 4 * THREADS) {
                        const uint32_t q0 = q0n, lw = lwn;
                        const uint4 r4 = rn;
                        q0n = base + 4 * THREADS + 4u * (uint32_t)tid;
                        lwn = q0n < npos8 ? bits4(longbits, q0n) : 0u;
                        rn = lwn ? R4[q0n >> 2] : make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t fw = q0 < npos8 ? bits4(firstbits, q0) : 0u;
                        if ((lw | fw) == 0u) continue;
                        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
                        uint32_t g[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            g[j] = 0;
                            if ((lw >> j) & 1u) {
                                g[j] = rank_of(rr[j] & 0xFFFFu);
                                const uint32_t bk = g[j] >> shift;
                                atomicAdd(&gdirw[bk >> 1], (bk & 1) ? 0x10000u : 1u);
                            } else if ((fw >> j) & 1u) {
                                firstpos[rank_of(q0 + j)] = (uint16_t)(q0 + j);
                            }
                        }
                        // (ids of positions that are not flagged are never read: the four are stored together)
                        *reinterpret_cast<uint2 *>(gids + q0) = make_uint2(g[0] | (g[1] << 16), g[2] | (g[3] << 16));
                    }
                }
                __syncthreads();
                {
                    // exclusive scan of the u16 counts; every thread owns `cw` consecutive packed words
                    const uint32_t cw = (dwords + THREADS - 1) / THREADS;
                    const uint32_t w0 = min((uint32_t)tid * cw, dwords), w1 = min(w0 + cw, dwords);
                    uint32_t sum = 0;
                    for (uint32_t k = w0; k < w1; ++k) {
                        const uint32_t wv = gdirw[k];
                        sum += (wv & 0xFFFF) + (wv >> 16);
                    }
                    uint32_t inc = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    if (lane == 31) M.scan_tmp[warp] = inc; // (last read before the barrier that follows the histogram)
                    __syncthreads();
                    uint32_t run = inc - sum;
                    for (int k = 0; k < warp; ++k) run += M.scan_tmp[k];
                    for (uint32_t k = w0; k < w1; ++k) {
                        const uint32_t wv = gdirw[k];
                        const uint32_t lo = run;
                        run += wv & 0xFFFF;
                        const uint32_t hi = run;
                        run += wv >> 16;
                        gdirw[k] = (lo & 0xFFFF) | (hi << 16);
                    }
                }
                __syncthreads();
                {
                    // scatter in position-ordered rounds of 4096 positions (four consecutive ones per thread); the ids of the next
                    // round are fetched while this one runs
                    auto bits4 = [&](const uint32_t *bits, uint32_t q0) -> uint32_t { return q0 < (uint32_t)MAXB ? (bits[q0 >> 5] >> (q0 & 31)) & 0xFu : 0u; };
                    uint32_t q0n = 4u * (uint32_t)tid;
                    uint32_t lwn = q0n < npos8 ? bits4(longbits, q0n) : 0u;
                    uint2 gn = lwn ? *reinterpret_cast<const uint2 *>(gids + q0n) : make_uint2(0u, 0u);
#pragma unroll 1
                    for (uint32_t base = 0; base < npos8; base += 4 * THREADS) {
                        const uint32_t q0 = q0n, lw = lwn;
                        const uint2 g2 = gn;
                        q0n = base + 4 * THREADS + 4u * (uint32_t)tid;
                        lwn = q0n < npos8 ? bits4(longbits, q0n) : 0u;
                        gn = lwn ? *reinterpret_cast<const uint2 *>(gids + q0n) : make_uint2(0u, 0u);
                        const uint32_t gg[4] = {g2.x & 0xFFFFu, g2.x >> 16, g2.y & 0xFFFFu, g2.y >> 16};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if ((lw >> j) & 1u) {
                                const uint32_t bk = gg[j] >> shift;
                                const uint32_t old = atomicAdd(&gdirw[bk >> 1], (bk & 1) ? 0x10000u : 1u);
                                S[(bk & 1) ? (old >> 16) : (old & 0xFFFF)] = (uint16_t)(q0 + j);
                            }
                        }
                        __syncthreads(); // one round = one 4096-position chunk
                    }
                }
                LJB_PHASE(3); // index (8-gram groups)
                // ---- B1: sorted-order walk of the 8-gram index, first 16 bytes only.  The lanes of a warp sit in
                // the same bucket (equal trip counts, broadcast loads).  A position whose best candidate is
                // shorter than 16 bytes is final here; the others ("very long") go to B2.
                uint32_t *vlong = firstbits; // the first-occurrence bits are dead once the index is built
                for (int i = tid; i < MAXB / 32; i += THREADS) vlong[i] = 0;
                if (tid == 0) M.scan_tmp[1] = 0;                                // B1's ticket counter
                __syncthreads();
                {
                    uint32_t d_pos = 0, d_vis = 0, d_eq = 0;
                    const long long tb1 = clock64();
                    // Warps take 32 consecutive sorted entries at a time from a ticket counter (group sizes differ widely: a fixed
                    // assignment left warps waiting 4 % of the kernel at the barrier below); the next ticket, its entries and
                    // their group ids are fetched while the current entries are processed.
                    auto grab = [&]() -> uint32_t {
                        uint32_t t = 0;
                        if (lane == 0) t = atomicAdd(&M.scan_tmp[1], 1u);
                        return __shfl_sync(0xffffffffu, t, 0) * 32u + (uint32_t)lane;
                    };
                    uint32_t j_n = grab();
                    uint32_t p_n = j_n < nS ? (uint32_t)S[j_n] : 0u;
                    bool l_n = j_n < nS;
                    uint32_t g_n = l_n ? (uint32_t)gids[p_n] : 0u;
                    for (;;) {
                        if (j_n - (uint32_t)lane >= nS) break; // (warp-uniform)
                        const uint32_t p = p_n, gp = g_n;
                        const bool lp = l_n;
                        {
                            j_n = grab();
                            p_n = j_n < nS ? (uint32_t)S[j_n] : 0u;
                            l_n = j_n < nS;
                            g_n = l_n ? (uint32_t)gids[p_n] : 0u;
                        }
                        if (!lp) continue; // beyond the last entry
                        ++d_pos;
                        const uint32_t pi = p >> 2, ps = (p & 3) * 8;
                        const uint32_t w0 = dataw[pi], w1 = dataw[pi + 1], w2 = dataw[pi + 2], w3 = dataw[pi + 3], w4 = dataw[pi + 4];
                        const uint32_t P0 = __funnelshift_r(w0, w1, ps), P1 = __funnelshift_r(w1, w2, ps);
                        const uint32_t P2 = __funnelshift_r(w2, w3, ps), P3 = __funnelshift_r(w3, w4, ps);
                        const uint32_t bk = gp >> shift;
                        const uint32_t lo = bk ? gdir16[bk - 1] : 0u, hi = gdir16[bk];
                        const uint32_t cap16 = min(16u, nb - p);
                        const uint32_t pch = p >> CH;
                        // a chain = consecutive flagged positions inside one 32-position chunk (B2's unit of sequential work)
                        const bool chain_start = (p & (uint32_t)(CHUNK - 1)) == 0u || !((longbits[(p - 1) >> 5] >> ((p - 1) & 31)) & 1u);
                        const uint32_t prev_byte = p ? data[p - 1] : 0u;
                        uint32_t best = 0, n16 = 0, cand = 0xFFFFFFFFu;
                        const bool budgeted = !chain_start && nb - p > 16 && hi - lo > (uint32_t)B1_BUDGET;
                        // candidates: the first occurrence of the group, then its flagged members in chunk order
                        const uint32_t nvis = budgeted ? 0u : hi - lo + 1u; // (a position inside a chain does not look at a huge group at all)
                        uint32_t c_next = firstpos[gp];
                        for (uint32_t i = 0; i < nvis; ++i) {
                            const uint32_t c = c_next;
                            c_next = S[lo + i]; // one entry ahead (S has slack past the last bucket): shortens the dependent chain
                            const uint32_t cch = c >> CH;
                            if (cch > pch) break; // only later positions from here on
                            ++d_vis;
                            // Inside a chain only the (few) pairs that start a diagonal count, so nothing would end the walk of a huge
                            // group early (runs of one byte, short periods: every position is in one group): positions inside a
                            // chain whose group has more than B1_BUDGET entries are handed to B2 at once, whose own walk is pruned
                            // by the pairs it carries (hi_w == lo).  A third 16-byte candidate means B2 walks the group anyway.
                            // (that exit sits where the candidates are counted.)  Close to the block end, where a match cannot exceed
                            // 16 bytes, later positions cannot win a tie once the best pair reaches the end
                            if (nb - p <= 16 && (best >> 16) == cap16 && cch > ((0xFFFFu - (best & 0xFFFFu)) >> CH)) break;
                            if (c < p) {
                                const uint32_t ci = c >> 2, cs = (c & 3) * 8;
                                // an entry of the same group has the same 8 bytes; shared buckets need the comparison
                                if (exact || (__funnelshift_r(dataw[ci], dataw[ci + 1], cs) == P0 && __funnelshift_r(dataw[ci + 1], dataw[ci + 2], cs) == P1)) {
                                    // inside a chain B2 carries the pairs that continue a diagonal: only pairs that START one here
                                    // (different byte in front, or nothing in front) are measured
                                    if (!chain_start && c != 0u && data[c - 1] == prev_byte) continue;
                                    ++d_eq;
                                    const uint32_t a2 = dataw[ci + 2], a3 = dataw[ci + 3], a4 = dataw[ci + 4];
                                    const uint32_t x2 = __funnelshift_r(a2, a3, cs) ^ P2, x3 = __funnelshift_r(a3, a4, cs) ^ P3;
                                    uint32_t l = x2 ? 8u + ((uint32_t)(__ffs(x2) - 1) >> 3) : (x3 ? 12u + ((uint32_t)(__ffs(x3) - 1) >> 3) : 16u);
                                    l = min(l, cap16);
                                    best = max(best, (l << 16) | (0xFFFFu - c));
                                    if (l == 16) { // remember up to 2 candidates that may be longer; B2 measures only these
                                        if (n16 == 0) cand = (cand & 0xFFFF0000u) | c;
                                        if (n16 == 1) cand = (cand & 0xFFFFu) | (c << 16);
                                        ++n16;
                                        if (n16 > 2 && nb - p > 16) break; // B2 walks the group anyway
                                    }
                                }
                            }
                        }
                        const bool forced = budgeted;
                        const uint32_t bl = best >> 16, bp = 0xFFFFu - (best & 0xFFFFu);
                        if ((bl == 16 && nb - p > 16) || forced) { // may be longer: B2 decides
                            atomicOr(&vlong[p >> 5], 1u << (p & 31));
                            if (n16 > 2 || forced) { // overflow: B2 walks the group itself
                                cand = (cand & 0xFFFFu) | 0xFFFE0000u;
                            }
                            R[p] = cand; // provisional: the two candidates, in place of (length, position)
                        } else {
                            R[p] = best ? ((bl << 16) | bp) : 0u; // 0: no pair starts here (the match continues a diagonal)
                        }
                    }
                    if (P.phase_cycles) {
                        const unsigned t_pos = __reduce_add_sync(0xffffffffu, d_pos), t_vis = __reduce_add_sync(0xffffffffu, d_vis);
                        const unsigned t_eq = __reduce_add_sync(0xffffffffu, d_eq);
                        if (lane == 0) {
                            atomicAdd(&P.phase_cycles[16], (unsigned long long)t_pos);
                            atomicAdd(&P.phase_cycles[17], (unsigned long long)t_vis);
                            atomicAdd(&P.phase_cycles[18], (unsigned long long)t_eq);
                        }
                        if (lane == 0) atomicAdd(&P.phase_cycles[19], (unsigned long long)(clock64() - tb1));
                        if (tid == 0) atomicAdd(&P.phase_cycles[20], (unsigned long long)nidx);
                    }
                }
                __syncthreads();
                LJB_PHASE(1); // search phase B1
                // ---- B2: every position with an >= 8 byte match, in position order.
                // Lemma (order preservation): if (c, p) and (c', p) both match >= 9 bytes, then (c+1, p+1) and (c'+1, p+1)
                // match one byte less each, so the best pair of p stays the best among all pairs that CONTINUE a diagonal
                // into p+1.  A pair that starts a new diagonal at p+1 is "left-maximal" (data[c-1] != data[p], or c == 0),
                // and only those were measured by B1 for positions inside a chain.  Hence
                //     best(p+1) = better of { best(p) moved along its diagonal, B1's best over the left-maximal pairs of p+1 }.
                // The lemma needs true lengths: pairs cut by the cap (1024 or the block end) tie there and may separate
                // later, so ALL pairs that reach the cap are carried along (up to KEEP; beyond that, and whenever a bucket
                // walk stopped early, the next position walks its bucket again).
                constexpr int KEEP = 4;
                constexpr uint32_t SHORT = 32; // a lane compares this much on its own; longer runs are compared by the whole warp
                // Every lane works through chunks of CHUNK positions, one flagged position per step, carrying the pairs of the
                // previous position along; a lane that has finished its chunk takes the next one from a queue (the chunks in
                // block order), so all lanes of a warp have a position to work on at every step.
                if (tid == 0) M.scan_tmp[0] = 0;
                __syncthreads();
                {
                    const uint32_t nq = (nb + CHUNK - 1) / CHUNK;
                    const uint32_t cmask = CHUNK == 32 ? 0xFFFFFFFFu : ((1u << CHUNK) - 1u);
                    uint32_t ch = 0, bits = 0, vbits = 0, rem = 0; // this lane's chunk, its flagged positions, the ones still to do
                    uint32_t pc[KEEP], pl[KEEP]; // pairs carried from the previous position: candidate position, length | capped << 16
                    uint32_t pn = 0;
                    bool pinc = false;           // the carried set may be incomplete: walk the bucket
                    uint32_t dbg_pos = 0, dbg_inh = 0, dbg_walk = 0;
                    const long long trow0 = clock64();
                    // the records of the next three positions of the lane are in flight (L2 round trips) while one is processed
                    auto fetch = [&](uint32_t m) -> uint32_t { return m ? R[ch * CHUNK + (__ffs(m) - 1)] : 0u; };
                    uint32_t sl_a = 0, sl_b = 0, sl_c = 0;
                    bool drained = false; // (warp-uniform) the queue is empty
                    for (;;) {
                        while (!drained) {
                            const unsigned need = __ballot_sync(0xffffffffu, rem == 0);
                            if (!need) break;
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(&M.scan_tmp[0], (uint32_t)__popc(need));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            const uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                            if (rem == 0 && idx < nq) {
                                ch = nq - 1u - idx; // (from the end of the block: the chunks whose matches run into the block end are the slow ones)
                                const uint32_t cw = ch * CHUNK >> 5, csh = (ch * CHUNK) & 31u;
                                bits = (longbits[cw] >> csh) & cmask;
                                vbits = (vlong[cw] >> csh) & cmask;
                                rem = bits;
                                if (rem) {
                                    pn = 0;
                                    pinc = false;
                                    const uint32_t m2 = rem & (rem - 1);
                                    sl_a = fetch(rem);
                                    sl_b = fetch(m2);
                                    sl_c = fetch(m2 & (m2 - 1));
                                }
                            }
                            drained = base + (uint32_t)__popc(need) >= nq;
                        }
                        if (!__any_sync(0xffffffffu, rem != 0)) break; // (then the queue is empty as well)
                        const bool active = rem != 0;
                        const int i = active ? __ffs(rem) - 1 : 0;
                        rem &= rem - 1; // (0 stays 0)
                        const uint32_t sl = sl_a;
                        sl_a = sl_b;
                        sl_b = sl_c;
                        {
                            const uint32_t m2 = rem & (rem - 1);
                            sl_c = fetch(m2 & (m2 - 1));
                        }
                        const uint32_t p = ch * CHUNK + i;
                        const bool chained = active && i > 0 && ((bits >> (i - 1)) & 1u);
                        if (!chained) {
                            pn = 0;
                            pinc = false;
                        }
                        uint32_t cap = 0, P0 = 0, P1 = 0;
                        if (active) {
                            ++dbg_pos;
                            cap = min((uint32_t)MAX_MATCH, nb - p);
                            P0 = load32u(dataw, p);
                            P1 = load32u(dataw, p + 4);
                        }
                        const bool isv = active && ((vbits >> i) & 1u);
                        bool walk = active && (pinc || (isv && (sl >> 16) == 0xFFFEu));
                        // this position's pairs: the capped ones (all of them matter later) and the best uncapped one
                        uint32_t nc[KEEP], nl[KEEP], nn = 0;
                        uint32_t bestkey = 0; // (length << 16) | (0xFFFF - position): longest, then earliest
                        bool ninc = false;
                        auto take = [&](uint32_t l, uint32_t c) {
                            bestkey = max(bestkey, (l << 16) | (0xFFFFu - c));
                            if (l == cap) {
                                if (nn < KEEP) {
#pragma unroll
                                    for (int u = 0; u < KEEP; ++u)
                                        if ((uint32_t)u == nn) {
                                            nc[u] = c;
                                            nl[u] = l | 0x10000u;
                                        }
                                    ++nn;
                                } else {
                                    ninc = true;
                                }
                            }
                        };
                        // 1. pairs carried along their diagonals
#pragma unroll
                        for (int t = 0; t < KEEP; ++t) {
                            if (active && (uint32_t)t < pn) {
                                const uint32_t c = pc[t] + 1;
                                uint32_t l = (pl[t] & 0xFFFF) - 1;
                                if (pl[t] >> 16) // the length was cut by the cap, not by a mismatch: it may go on
                                    while (l < cap && data[c + l] == data[p + l]) ++l;
                                l = min(l, cap);
                                pc[t] = c; // now this position's candidate (duplicate test below)
                                take(l, c);
                                ++dbg_inh;
                            }
                        }
                        // 2. what B1 measured: either the best pair below 16 bytes, or up to two 16-byte candidates
                        if (active && !isv && sl != 0u && !walk) take(sl >> 16, sl & 0xFFFFu);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t c = (sl >> (q * 16)) & 0xFFFFu;
                            bool eval = isv && !walk && c != 0xFFFFu;
                            if (eval && (bestkey >> 16) == (uint32_t)MAX_MATCH) { // a carried pair already gives 1024: not measured now,
                                eval = false;                                     // but it may be the one that lasts longest
                                ninc = true;
                            }
#pragma unroll
                            for (int t = 0; t < KEEP; ++t) eval = eval && !((uint32_t)t < pn && pc[t] == c); // already carried
                            uint32_t l = 0;
                            if (eval) l = lcp_from(dataw, c, p, P0, P1, min(cap, SHORT));
                            // runs that are still matching after SHORT bytes: the warp compares 256 bytes per step
                            unsigned pend = __ballot_sync(0xffffffffu, eval && l == SHORT && cap > SHORT);
                            while (pend) {
                                const int src = __ffs(pend) - 1;
                                pend &= pend - 1;
                                const uint32_t bc = __shfl_sync(0xffffffffu, c, src), bp_ = __shfl_sync(0xffffffffu, p, src);
                                const uint32_t bcap = __shfl_sync(0xffffffffu, cap, src);
                                uint32_t res = bcap;
                                for (uint32_t base = SHORT; base < bcap; base += 256) {
                                    const uint32_t off = base + 8 * lane;
                                    uint32_t x0 = 0, x1 = 0;
                                    if (off < bcap) {
                                        const uint32_t ci = (bc + off) >> 2, cs = ((bc + off) & 3) * 8;
                                        const uint32_t pi = (bp_ + off) >> 2, ps = ((bp_ + off) & 3) * 8;
                                        const uint32_t a0 = dataw[ci], a1 = dataw[ci + 1], a2 = dataw[ci + 2];
                                        const uint32_t b0 = dataw[pi], b1 = dataw[pi + 1], b2 = dataw[pi + 2];
                                        x0 = __funnelshift_r(a0, a1, cs) ^ __funnelshift_r(b0, b1, ps);
                                        x1 = __funnelshift_r(a1, a2, cs) ^ __funnelshift_r(b1, b2, ps);
                                    }
                                    const unsigned mm = __ballot_sync(0xffffffffu, (x0 | x1) != 0);
                                    if (mm) {
                                        const int f = __ffs(mm) - 1;
                                        const uint32_t mine = off + (x0 ? ((uint32_t)(__ffs(x0) - 1) >> 3) : 4u + ((uint32_t)(__ffs(x1) - 1) >> 3));
                                        res = min(__shfl_sync(0xffffffffu, mine, f), bcap);
                                        break;
                                    }
                                }
                                if (lane == src) l = res;
                            }
                            if (eval) take(l, c);
                        }
                        // 3. too many candidates, or an incomplete carried set: walk the bucket with pruning.  A capped 1024
                        //    match becomes a literal step ((uint8_t)1024 == 0, LZ4.c:317) whose distance is never used, so
                        //    then nothing else needs to be looked at here (but the set stays marked incomplete).
                        if (walk) {
                            ++dbg_walk;
                            if ((bestkey >> 16) == (uint32_t)MAX_MATCH) {
                                ninc = true;
                            } else {
                                const uint32_t bk = (uint32_t)gids[p] >> shift;
                                const uint32_t lo = bk ? gdir16[bk - 1] : 0u, hi = gdir16[bk];
                                const uint32_t pch = p >> CH;
                                const uint32_t fq = firstpos[gids[p]];
                                for (uint32_t k = lo; k <= hi; ++k) { // the first occurrence of the group, then its flagged members
                                    const uint32_t c = k == lo ? fq : (uint32_t)S[k - 1];
                                    if ((c >> CH) > pch) break; // only later positions from here on
                                    if (c >= p) continue;
                                    const uint32_t bl = bestkey >> 16, bp = 0xFFFFu - (bestkey & 0xFFFFu);
                                    // later positions cannot win a tie (and nothing in the group lies before its first occurrence) ...
                                    if (bl == cap && ((c >> CH) > (bp >> CH) || bp == fq)) {
                                        ninc = true; // ... but may reach the cap as well
                                        break;
                                    }
                                    // a candidate matters only if it beats the best length, or ties it from an earlier position
                                    const uint32_t need = max(8u, bl == 0 ? 8u : (c < bp ? bl : bl + 1));
                                    if (need > cap) continue;
                                    bool dup = false;
#pragma unroll
                                    for (int t = 0; t < KEEP; ++t) dup |= ((uint32_t)t < pn && pc[t] == c);
                                    if (dup) continue; // already measured as a carried pair
                                    if (data[c + need - 1] != data[p + need - 1]) continue;
                                    const uint32_t l = lcp_from(dataw, c, p, P0, P1, cap);
                                    if (l < 8) continue; // a different 8-gram in the same bucket
                                    take(l, c);
                                    if (l == (uint32_t)MAX_MATCH) {
                                        ninc = true;
                                        break;
                                    }
                                }
                            }
                        }
                        if (active) {
                            const uint32_t bl = bestkey >> 16, bp = 0xFFFFu - (bestkey & 0xFFFFu);
                            R[p] = (bl << 16) | bp;
                            if (bl == cap) { // carry every pair that reached the cap
#pragma unroll
                                for (int t = 0; t < KEEP; ++t) {
                                    pc[t] = nc[t];
                                    pl[t] = nl[t];
                                }
                                pn = nn;
                            } else { // carry the best pair only (order preservation)
                                pc[0] = bp;
                                pl[0] = bl;
                                pn = bl >= 9 ? 1u : 0u;
                                ninc = false; // pairs below the best never matter again
                            }
                            pinc = ninc;
                        }
                    }
                    if (P.phase_cycles) {
                        const unsigned t_pos = __reduce_add_sync(0xffffffffu, dbg_pos), t_inh = __reduce_add_sync(0xffffffffu, dbg_inh);
                        const unsigned t_walk = __reduce_add_sync(0xffffffffu, dbg_walk);
                        if (lane == 0) {
                            atomicAdd(&P.phase_cycles[9], (unsigned long long)t_pos);
                            atomicAdd(&P.phase_cycles[13], (unsigned long long)t_inh);
                            atomicAdd(&P.phase_cycles[12], (unsigned long long)t_walk);
                            atomicAdd(&P.phase_cycles[14], (unsigned long long)(clock64() - trow0));
                        }
                    }
                }
            }
        }
        if (P.phase_cycles && lane == 0) {
            // (profiling aid) how long this warp was busy in B2, and the slowest warp
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
        // (eight records in flight per thread: the loop is bound by the L2 round trip, not by bandwidth)
#pragma unroll 1
        for (uint32_t base = 0; base < nb; base += 8 * THREADS) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t p = base + j * THREADS + tid;
                r[j] = p < nb ? R[p] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t p = base + j * THREADS + tid;
                if (p < nb) step[p] = (uint8_t)(r[j] >> 16);
            }
        }
        for (int i = tid; i < 1024; i += THREADS) entry[i] = 0xFF;
        __syncthreads();
        // pass A: x1[p] = (first chain position >= segment end) - segment end, for every p (walk descending)
        if ((uint32_t)tid < nseg) {
            for (uint32_t p = s1; p-- > s0;) {
                uint32_t st = step[p];
                uint32_t t = p + (st ? st : 1u);
                flag[p] = (uint8_t)(t >= s1 ? t - s1 : flag[t]);
            }
        }
        __syncthreads();
        LJB_PHASE(15); // (probe) step extraction + pass A
        // pass B: the same for super-segments of SUP segments: where the chain leaves one, for every offset (< 255) at which it can
        // enter it.  8160 short independent walks instead of one long dependent one.
        uint8_t *const exit2 = smem + SM_EXIT2; // [32][256]
        const uint32_t nsup = (nseg + SUP - 1) / SUP;
        for (uint32_t task = tid; task < nsup * 256u; task += THREADS) {
            const uint32_t u = task >> 8, e = task & 255u;
            const uint32_t end_u = min((u + 1) * (uint32_t)(SUP * SEG), nb);
            uint32_t pos = u * (uint32_t)(SUP * SEG) + e;
            while (pos < end_u) {
                const uint32_t x = flag[pos];
                const uint32_t sg = pos / SEG;
                pos = min((sg + 1) * SEG, nb) + x;
            }
            exit2[task] = (uint8_t)min(pos - end_u, 255u); // (only read for offsets the chain can really enter at; past the end: unused)
        }
        __syncthreads();
        // pass C: thread 0 hops from super-segment to super-segment, then one lane per super-segment records where the chain
        // enters each of its segments.  Meanwhile warps 1 .. 31 place the previous block.
        if (warp == 0) {
            uint16_t *const supentry = reinterpret_cast<uint16_t *>(smem + SM_EXIT2 + 32 * 256); // [32]
            supentry[lane] = 0xFFFFu; // (a short last super-segment may be jumped over)
            __syncwarp();
            if (lane == 0) {
                uint32_t pos = 0;
                for (uint32_t u = 0; u < nsup && pos < nb; ++u) {
                    const uint32_t e = pos - u * (uint32_t)(SUP * SEG);
                    supentry[u] = (uint16_t)e;
                    pos = min((u + 1) * (uint32_t)(SUP * SEG), nb) + exit2[u * 256u + e];
                }
            }
            __syncwarp();
            if ((uint32_t)lane < nsup && supentry[lane] != 0xFFFFu) {
                const uint32_t end_u = min(((uint32_t)lane + 1) * (uint32_t)(SUP * SEG), nb);
                uint32_t pos = (uint32_t)lane * (uint32_t)(SUP * SEG) + supentry[lane];
                while (pos < end_u) {
                    const uint32_t x = flag[pos];
                    const uint32_t sg = pos / SEG;
                    entry[sg] = (uint8_t)(pos - sg * SEG);
                    pos = min((sg + 1) * SEG, nb) + x;
                }
            }
        } else if (pend) {
            flush_pending(1); // ---------------- P7: place the PREVIOUS block, in the shadow of the hops ----------------
        }
        __syncthreads(); // the exit table is dead from here on: its memory becomes the slots and the output area
        LJB_PHASE(5); // parse: chain resolution
        } // MODE == 0

        // ---------------- sizing and emission ----------------
        // A sequence = the literals since the previous match + one match (LZ4.c:516-583).
        //   E1  one thread per segment walks the chain inside its segment and parks the match positions (the first SLOTS of
        //       them; more than that needs runs of one- to three-byte phantom matches) in shared memory
        //   S   eight lanes per segment size the parked sequences (a sequence needs its own match, the match before it, and
        //       for the first one of a segment the end of the last match before the segment: block scan 1)
        //   scan 2 gives every segment the byte offset of its first sequence
        //   E2  eight lanes per segment serialise the parked sequences; whatever a segment could not park is sized and
        //       serialised by the segment's own thread
        uint16_t *const slots = reinterpret_cast<uint16_t *>(smem + SM_SLOT);
        uint4 *const seginfo = reinterpret_cast<uint4 *>(smem + SM_SEGINFO); // x: end of the last match before, y: byte offset, z: payload, w: size sum
        uint8_t *const segcnt = smem + SM_SEGCNT;                            // matches per segment (<= SEG)
        uint8_t *const segph = segcnt + 512;                                 // phantom sequences among the parked ones
        uint16_t *const myslots = slots + tid * SLOTS;
        const bool has_seg = (uint32_t)tid < nseg && entry[tid] != 0xFF;
        uint32_t cnt = 0, seg_last_end = 0, ov_p = 0, ov_pe = 0;
        if (has_seg) {
            uint32_t p = s0 + entry[tid], pe = 0;
            while (p < s1) {
                const uint32_t st = step[p];
                if (st) {
                    if (cnt < (uint32_t)SLOTS) {
                        myslots[cnt] = (uint16_t)p;
                    } else if (cnt == (uint32_t)SLOTS) {
                        ov_p = p;
                        ov_pe = pe;
                    }
                    pe = p + st;
                    ++cnt;
                    p += st;
                } else {
                    ++p;
                }
            }
            seg_last_end = pe;
        }
        if (tid < 512) segcnt[tid] = (uint8_t)cnt;
        LJB_PHASE(10); // (probe) E1 walk
        // scan 1: end of the last match before this segment (exclusive max; ends grow along the chain)
        uint32_t prev_end_in;
        {
            uint32_t inc = seg_last_end;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = max(inc, v);
            }
            if (lane == 31) M.scan_tmp[warp] = inc;
            uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) ex = 0;
            __syncthreads();
            uint32_t wmax = 0;
            for (int k = 0; k < warp; ++k) wmax = max(wmax, M.scan_tmp[k]);
            prev_end_in = max(wmax, ex);
            if (tid < 512) seginfo[tid].x = prev_end_in;
        }
        uint32_t last_end = 0; // end of the last match of the block
        for (int k = 0; k < NWARPS; ++k) last_end = max(last_end, M.scan_tmp[k]);
        __syncthreads();
        // the sequence parked in slot i of a segment: its match, and the end of the match before it
        auto parked = [&](const uint16_t *sl, uint32_t i, uint32_t seg_prev_end, uint32_t &p, uint32_t &ml, uint32_t &pe) {
            p = sl[i];
            ml = step[p];
            if (i) {
                const uint32_t pp = sl[i - 1];
                pe = pp + step[pp];
            } else {
                pe = seg_prev_end;
            }
        };
        // S: eight lanes per segment, 128 segments per pass
        const int gj = tid & 7;
#pragma unroll 1
        for (uint32_t seg = (uint32_t)tid >> 3; seg < 512u; seg += THREADS / 8) {
            const uint32_t c = seg < nseg ? (uint32_t)segcnt[seg] : 0u, m = min(c, (uint32_t)SLOTS);
            const uint16_t *sl = slots + seg * SLOTS;
            const uint32_t spe = seginfo[seg].x;
            uint32_t gp = 0, gs = 0, gph = 0;
            for (uint32_t i = gj; i < m; i += 8) {
                uint32_t p, ml, pe;
                parked(sl, i, spe, p, ml, pe);
                const SeqSize z = seq_size(p - pe, ml);
                gp += z.payload;
                gs += z.byte_size;
                gph += (z.payload != z.byte_size) ? 1u : 0u;
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                gp += __shfl_xor_sync(0xffffffffu, gp, o);
                gs += __shfl_xor_sync(0xffffffffu, gs, o);
                gph += __shfl_xor_sync(0xffffffffu, gph, o);
            }
            if (gj == 0) {
                seginfo[seg].z = gp;
                seginfo[seg].w = gs;
                segph[seg] = (uint8_t)gph;
            }
        }
        __syncthreads();
        // scan 2: bytes, size-field sums, sequence and phantom counts before this segment
        unsigned long long my_prefix; // offset of this segment's first sequence in the encoded block
        unsigned long long pay = 3, sizes = 3;
        uint32_t nseq = 0, phantom = 0, par_pay = 0;
        {
            uint32_t seg_pay = 0, seg_sizes = 0, seg_ph = 0;
            if (tid < 512) {
                seg_pay = seginfo[tid].z;
                seg_sizes = seginfo[tid].w;
                seg_ph = segph[tid];
            }
            par_pay = seg_pay;
            if (cnt > (uint32_t)SLOTS) { // what could not be parked: sized here
                uint32_t p = ov_p, pe = ov_pe;
                while (p < s1) {
                    const uint32_t st = step[p];
                    if (st) {
                        const SeqSize z = seq_size(p - pe, st);
                        seg_pay += z.payload;
                        seg_sizes += z.byte_size;
                        seg_ph += (z.payload != z.byte_size) ? 1u : 0u;
                        pe = p + st;
                        p += st;
                    } else {
                        ++p;
                    }
                }
            }
            unsigned long long x = (unsigned long long)seg_pay | ((unsigned long long)seg_sizes << 32);
            unsigned long long y = (unsigned long long)cnt | ((unsigned long long)seg_ph << 32);
            const unsigned long long own = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long vx = __shfl_up_sync(0xffffffffu, x, o), vy = __shfl_up_sync(0xffffffffu, y, o);
                if (lane >= o) {
                    x += vx;
                    y += vy;
                }
            }
            if (lane == 31) {
                M.warp_x[warp] = x;
                M.warp_y[warp] = y;
            }
            __syncthreads();
            unsigned long long bx = 0, tx = 0, ty = 0;
            for (int k = 0; k < NWARPS; ++k) {
                if (k == warp) bx = tx;
                tx += M.warp_x[k];
                ty += M.warp_y[k];
            }
            my_prefix = 3 + ((bx + x - own) & 0xFFFFFFFFull);
            if (tid < 512) seginfo[tid].y = (uint32_t)my_prefix;
            pay += tx & 0xFFFFFFFFull;
            sizes += tx >> 32;
            nseq = (uint32_t)(ty & 0xFFFFFFFFull);
            phantom = (uint32_t)(ty >> 32);
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
        // ---------------- P5: publish this block's size; P6: emit ----------------
        if (warp == 0) {
            ljb_lookback_publish(P.status + 1, b, pay, P.lead);
            if (lane == 0 && phantom) atomicAdd((unsigned long long *)&P.result[1], (unsigned long long)phantom);
        }
        LJB_PHASE(7); // look-back
        // The block is encoded in shared memory and copied to its staging buffer with 16-byte stores; a block too large for
        // that (it expanded) is encoded straight into the staging buffer through the same generic pointer.
        uint8_t *const stage = stage0 + (size_t)buf * P.stage_stride;
        const bool in_shared = pay <= (unsigned long long)SOUT_CAP;
        uint8_t *const obase = in_shared ? smem + SM_OUT : stage;
        // one sequence: token, size field, literal-length extension, literals (short runs here, long ones by the caller),
        // match offset, match-length extension (LZ4.c:365-413).  Returns where the literals go.
        auto put_sequence = [&](uint8_t *dst, uint32_t lit, uint32_t ml, uint32_t pe, uint32_t dist, const SeqSize &sz, bool copy_all) -> uint32_t {
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
            const uint32_t lit_dst_off = o;
            if (lit <= 16 || copy_all)
                for (uint32_t k = 0; k < lit; ++k) dst[o + k] = data[pe + k];
            o += lit;
            dst[o++] = (uint8_t)(dist & 0xFF);
            dst[o++] = (uint8_t)(dist >> 8);
            if (ml >= 19) dst[o++] = (uint8_t)(ml - 19);
            return lit_dst_off;
        };
        __syncthreads(); // byte offsets of the segments are in place
        {
            // E2: eight lanes per segment, four segments per warp and pass; the records of all rounds are fetched first
            constexpr int ROUNDS = (SLOTS + 7) / 8;
#pragma unroll 1
            for (uint32_t seg = (uint32_t)tid >> 3; seg < 512u; seg += THREADS / 8) {
                const uint32_t c = seg < nseg ? (uint32_t)segcnt[seg] : 0u, m = min(c, (uint32_t)SLOTS);
                const uint32_t mmax = __reduce_max_sync(0xffffffffu, m);
                if (mmax == 0) continue;
                const uint16_t *sl = slots + seg * SLOTS;
                const uint4 info = seginfo[seg];
                uint32_t rr[ROUNDS];
#pragma unroll
                for (int k = 0; k < ROUNDS; ++k) {
                    const uint32_t i = (uint32_t)(8 * k + gj);
                    rr[k] = i < m ? R[sl[i]] : 0u;
                }
                uint32_t running = info.y;
#pragma unroll
                for (int k = 0; k < ROUNDS; ++k) {
                    if ((uint32_t)(8 * k) >= mmax) break; // warp-uniform
                    const uint32_t i = (uint32_t)(8 * k + gj);
                    const bool act = i < m;
                    uint32_t p = 0, ml = 0, pe = 0, lit = 0;
                    SeqSize sz = {0, 0};
                    if (act) {
                        parked(sl, i, info.x, p, ml, pe);
                        lit = p - pe;
                        sz = seq_size(lit, ml);
                    }
                    uint32_t inc = sz.payload;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o, 8);
                        if (gj >= o) inc += v;
                    }
                    const uint32_t gtot = __shfl_sync(0xffffffffu, inc, 7, 8);
                    uint8_t *dst = obase + running + (inc - sz.payload);
                    uint32_t lit_dst_off = 0;
                    if (act) lit_dst_off = put_sequence(dst, lit, ml, pe, (p - (rr[k] & 0xFFFFu)) & 0xFFFFu, sz, false);
                    unsigned lits_long = __ballot_sync(0xffffffffu, act && lit > 16);
                    while (lits_long) { // long literal runs: the whole warp copies them
                        const int L = __ffs(lits_long) - 1;
                        lits_long &= lits_long - 1;
                        const uint32_t n_l = __shfl_sync(0xffffffffu, lit, L);
                        const uint32_t s_l = __shfl_sync(0xffffffffu, pe, L);
                        const unsigned long long d_l =
                            __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)(dst + lit_dst_off), L);
                        uint8_t *dp = reinterpret_cast<uint8_t *>((uintptr_t)d_l);
                        for (uint32_t q = lane; q < n_l; q += 32) dp[q] = data[s_l + q];
                    }
                    running += gtot;
                }
            }
            if (cnt > (uint32_t)SLOTS) { // what could not be parked: serialised by the segment's own thread
                uint32_t p = ov_p, pe = ov_pe;
                unsigned long long off = my_prefix + par_pay;
                while (p < s1) {
                    const uint32_t st = step[p];
                    if (st) {
                        const uint32_t lit = p - pe;
                        const SeqSize sz = seq_size(lit, st);
                        put_sequence(obase + off, lit, st, pe, (p - (R[p] & 0xFFFFu)) & 0xFFFFu, sz, true);
                        off += sz.payload;
                        pe = p + st;
                        p += st;
                    } else {
                        ++p;
                    }
                }
            }
            LJB_PHASE(11); // (probe) E2 walk
            if (tid == 0) {
                uint8_t *hdr = obase;
                hdr[0] = (uint8_t)(nseq & 0xFF);          // LZ4.c:615, :417
                hdr[1] = (uint8_t)(sizes & 0xFF);         // LZ4.c:617, :419 (low 16 bits)
                hdr[2] = (uint8_t)((sizes >> 8) & 0xFF);
                if (b == 0 && P.lead && P.out_cap) P.out[0] = (uint8_t)P.frame_byte; // LZ4.c:429
            }
            if (tlit) {
                uint8_t *dst = obase + trail_off;
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
            if (in_shared) {
                __syncthreads();
                const uint4 *s4 = reinterpret_cast<const uint4 *>(smem + SM_OUT);
                uint4 *d4 = reinterpret_cast<uint4 *>(stage);
                const uint32_t n16 = (uint32_t)((pay + 15) >> 4);
                for (uint32_t i = tid; i < n16; i += THREADS) d4[i] = s4[i];
            }
        }
        __syncthreads(); // staging writes of this block are complete; region B and data are free
        LJB_PHASE(8); // emit
        pend = true;
        pend_b = b;
        pend_pay = pay;
        buf ^= 1;
    }
#undef LJB_PHASE
}

} // namespace lz4k

#ifndef LJB_EMU_BUILD
// ---- host side ---------------------------------------------------------------------------------------
extern "C" size_t ljb_lz4_block_count(size_t n, size_t block_len) { return block_len ? (n + block_len - 1) / block_len : 0; }

extern "C" size_t ljb_lz4_bound(size_t n, size_t block_len)
{
    // worst case of the dialect: a 5-byte sequence header per input byte (ml = 1 phantom matches) + literals
    return 1 + 3 * ljb_lz4_block_count(n, block_len) + 6 * n + 64;
}

static int lz4_launch(ljb_ctx *ctx, const uint8_t *d_in, size_t n, size_t block_len, uint8_t *d_out, size_t out_cap,
                      uint64_t *d_block_offsets, uint64_t *d_result, size_t first_block, size_t frame_blocks,
                      uint16_t *d_dump_len, uint16_t *d_dump_dist, uint64_t offs_bias = 0)
{
    using namespace lz4k;
    if (!ctx || !d_in || !d_out || !d_block_offsets || !d_result || n == 0 || block_len == 0 || block_len > MAXB)
        return LJB_E_ARG;
    const size_t nblocks = ljb_lz4_block_count(n, block_len);
    if (nblocks > 0x7fffffffull) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    if (!d_dump_len && block_len <= (size_t)SMALL_MAXB && !getenv("LJB_LZ4_NO_SMALL")) { // small blocks: one warp per block (lz4_small.cuh)
        const SmallGeom g = small_geometry((uint32_t)block_len);
        const size_t want = (nblocks + g.warps - 1) / g.warps;
        const int sgrid = (int)(want < (size_t)ctx->num_sms ? want : (size_t)ctx->num_sms);
        const size_t stride = (ljb_lz4_bound(block_len, block_len) + 63) & ~(size_t)63;
        int src;
        if ((src = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, (size_t)sgrid * g.warps * stride)) != 0) return src;
        if ((src = ljb_ensure(&ctx->d_status, &ctx->status_bytes, (nblocks + 2) * sizeof(uint64_t))) != 0) return src;
        LJB_CUDA(cudaMemsetAsync(ctx->d_status, 0, (nblocks + 2) * sizeof(uint64_t), ctx->stream));
        LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
        SmallParams S;
        S.in = d_in;
        S.n = n;
        S.block_len = (uint32_t)block_len;
        S.nblocks = (uint32_t)nblocks;
        S.out = d_out;
        S.out_cap = out_cap;
        S.block_offsets = d_block_offsets;
        S.result = d_result;
        S.status = (uint64_t *)ctx->d_status;
        S.staging = (uint8_t *)ctx->d_scratch;
        S.stage_stride = stride;
        S.offs_bias = offs_bias;
        S.lead = first_block == 0 ? 1u : 0u;
        S.frame_byte = (uint32_t)(frame_blocks & 0xFF);
        S.hbits = g.hbits;
        S.warp_bytes = g.warp_bytes;
        S.data_bytes = g.data_bytes;
        if (!(ctx->attr_mask & LJB_ATTR_LZ4_SMALL)) {
            LJB_CUDA(cudaFuncSetAttribute(lz4_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            ctx->attr_mask |= LJB_ATTR_LZ4_SMALL;
        }
        ctx->kernel_ms_summed = 0;
        LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        lz4_small_kernel<<<sgrid, 32 * g.warps, g.smem_bytes, ctx->stream>>>(S);
        LJB_CUDA(cudaGetLastError());
        LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->launches += 1;
        return LJB_OK;
    }
    const int grid = (int)((nblocks < (size_t)ctx->num_sms) ? nblocks : (size_t)ctx->num_sms);
    int rc;
    const size_t rec32_bytes = (size_t)ctx->num_sms * MAXB * sizeof(uint32_t);
    const size_t idx8_bytes = (size_t)ctx->num_sms * (NBUCKET + MAXB / 2) * sizeof(uint32_t); // 8-gram index of low-entropy blocks
    const size_t rec_bytes = rec32_bytes + (size_t)ctx->num_sms * MAXB * sizeof(uint16_t) + idx8_bytes; // match records + group ids / steps + 8-gram index
    const size_t stage_stride = (ljb_lz4_bound(block_len, block_len) + 16 + 255) & ~(size_t)255; // one encoded block, worst case
    if ((rc = ljb_ensure(&ctx->d_scratch, &ctx->scratch_bytes, rec_bytes + (size_t)grid * 2 * stage_stride)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_status, &ctx->status_bytes, (nblocks + 2 + 24) * sizeof(uint64_t))) != 0) return rc;
    LJB_CUDA(cudaMemsetAsync(ctx->d_status, 0, (nblocks + 2 + 24) * sizeof(uint64_t), ctx->stream));
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
    P.gids = (uint16_t *)((uint8_t *)ctx->d_scratch + rec32_bytes);
    P.idx8 = (uint32_t *)((uint8_t *)ctx->d_scratch + rec_bytes - idx8_bytes);
    P.staging = (uint8_t *)ctx->d_scratch + rec_bytes;
    P.stage_stride = stage_stride;
    P.offs_bias = offs_bias;
    P.lead = first_block == 0 ? 1u : 0u;
    P.frame_byte = (uint32_t)(frame_blocks & 0xFF);
    P.dump_len = d_dump_len;
    P.dump_dist = d_dump_dist;
    P.tune = getenv("LJB_LZ4_TUNE") ? (uint32_t)atoi(getenv("LJB_LZ4_TUNE")) : 0u;
    P.phase_cycles = nullptr;
    if (getenv("LJB_LZ4_PHASES")) { // profiling aid: per-phase cycle counters behind the status words
        P.phase_cycles = (unsigned long long *)ctx->d_status + (nblocks + 2);
    }
    // the full search (every position) serves ljb_lz4_block_matches and LJB_LZ4_SEARCH=full; the product path searches along the chain
    const char *mode_env = getenv("LJB_LZ4_SEARCH");
    const bool lazy = !d_dump_len && !(mode_env && strcmp(mode_env, "full") == 0);
    if (!(ctx->attr_mask & LJB_ATTR_LZ4)) { // per device (context), not per process
        LJB_CUDA(cudaFuncSetAttribute(lz4_encode_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        LJB_CUDA(cudaFuncSetAttribute(lz4_encode_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        ctx->attr_mask |= LJB_ATTR_LZ4;
    }
    // The per-CTA match records (256 KiB each, rewritten for every block) are the only data the kernel re-reads: pin
    // them in L2 for the duration of the launch so that the input / output streams cannot push them out to HBM.
    cudaStreamAttrValue win;
    memset(&win, 0, sizeof win);
    if (ctx->l2_persist_bytes) {
        win.accessPolicyWindow.base_ptr = ctx->d_scratch;
        win.accessPolicyWindow.num_bytes = rec_bytes < ctx->l2_window_max ? rec_bytes : ctx->l2_window_max;
        win.accessPolicyWindow.hitRatio = 1.0f;
        win.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        win.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        LJB_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &win));
    }
    ctx->kernel_ms_summed = 0;
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (lazy) lz4_encode_kernel<1><<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    else lz4_encode_kernel<0><<<grid, THREADS, SM_TOTAL, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    if (ctx->l2_persist_bytes) {
        win.accessPolicyWindow.num_bytes = 0; // later launches on this stream (JPEG, copies) run without a window
        LJB_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &win));
    }
    if (P.phase_cycles) {
        unsigned long long ph[24];
        LJB_CUDA(cudaMemcpyAsync(ph, P.phase_cycles, sizeof ph, cudaMemcpyDeviceToHost, ctx->stream));
        LJB_CUDA(cudaStreamSynchronize(ctx->stream));
        static const char *names[9] = {"stage", "phaseB1", "ladderA", "index8", "phaseB", "parse", "sizing", "lookback", "emit"};
        unsigned long long tot = 0;
        for (int i = 0; i < 9; ++i) tot += ph[i];
        fprintf(stderr, "[ljb lz4 phases] cycles per block:");
        for (int i = 0; i < 9; ++i) fprintf(stderr, " %s=%.0f(%.0f%%)", names[i], (double)ph[i] / (double)nblocks, 100.0 * (double)ph[i] / (double)tot);
        fprintf(stderr, " total=%.0f\n", (double)tot / (double)nblocks);
        if (lazy) {
            fprintf(stderr, "[ljb lz4 lazy] names: phaseB1 = walk rounds >= 1, ladderA = first walks, index8 = 4-gram index, phaseB = steps + entries | per block: searches=%.0f candidates=%.0f warp-steps=%.0f long compares=%.0f positions in capped runs=%.0f\n",
                    (double)ph[16] / nblocks, (double)ph[17] / nblocks, (double)ph[18] / nblocks, (double)ph[12] / nblocks, (double)ph[13] / nblocks);
            return LJB_OK;
        }
        fprintf(stderr, "[ljb lz4 probes] per block: T1+passA=%.0f E1=%.0f E2=%.0f\n", (double)ph[15] / nblocks, (double)ph[10] / nblocks, (double)ph[11] / nblocks);
        fprintf(stderr, "[ljb lz4 phaseB1] per block: indexed=%.0f positions=%.0f visited=%.0f equal8=%.0f warp-cycles=%.0f | ladder: inserts=%.0f rounds=%.1f\n",
                (double)ph[20] / nblocks, (double)ph[16] / nblocks, (double)ph[17] / nblocks, (double)ph[18] / nblocks, (double)ph[19] / nblocks,
                (double)ph[21] / nblocks, (double)ph[22] / nblocks);
        fprintf(stderr, "[ljb lz4 phaseB2] per block: positions=%.0f carried=%.0f bucket walks=%.0f | warp-cycles in rows=%.0f\n",
                (double)ph[9] / nblocks, (double)ph[13] / nblocks, (double)ph[12] / nblocks, (double)ph[14] / nblocks);
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

// Host-buffer entry point.  The input is cut into chunks of whole blocks; chunk k+1 is copied to the device and
// chunk k-1 is copied back while chunk k is being encoded (three streams, two buffers each way), so that the
// PCIe transfers hide behind the kernel.  Every chunk is encoded as a shard (first_block / frame_blocks): its
// stream starts at offset 0 of its own device buffer and lands at the running offset of the host stream.
extern "C" int ljb_lz4_compress(ljb_ctx *ctx, const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                                uint64_t *block_offsets, size_t *out_len, uint64_t *phantom)
{
    if (!ctx || !in || !out || n == 0 || block_len == 0 || block_len > lz4k::MAXB) return LJB_E_ARG;
    const size_t nblocks = ljb_lz4_block_count(n, block_len);
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // Chunk schedule: 2 x the base chunk (256 MiB) in the middle — every chunk is a kernel launch whose last wave leaves SMs
    // idle, which outweighs finer overlap — ramping up from 1/8 of that at the start and down to 1/8 at the end, because the
    // first upload and the last download are the only transfers nothing hides.  (4 GiB at 28 GB/s of kernel: 165.8 / 159.5 /
    // 157.8 / 159.7 / 164.1 ms with middle chunks of 1 GiB / 512 / 256 / 128 / 64 MiB.)
    size_t cblocks = 2 * ljb_pipe_chunk() / block_len; // blocks per full chunk
    if (cblocks == 0) cblocks = 1;
    std::vector<size_t> start; // first block of every chunk, plus the end
    {
        const size_t ramp[3] = {cblocks / 8, cblocks / 4, cblocks / 2};
        const size_t ramp_total = ramp[0] + ramp[1] + ramp[2];
        std::vector<size_t> sizes;
        if (ramp[0] >= 1 && nblocks >= 2 * ramp_total + cblocks) {
            for (int i = 0; i < 3; ++i) sizes.push_back(ramp[i]);
            size_t mid = nblocks - 2 * ramp_total;
            while (mid) {
                const size_t t = mid < cblocks ? mid : cblocks;
                sizes.push_back(t);
                mid -= t;
            }
            for (int i = 2; i >= 0; --i) sizes.push_back(ramp[i]);
        } else {
            // too small for the ramp: about eight chunks of whole waves of blocks (one block per SM at a time), so that uploads,
            // kernels and downloads still overlap (256 MiB: 18.1 -> 12 ms)
            const size_t sms = (size_t)(ctx->num_sms > 0 ? ctx->num_sms : 1);
            size_t c = ((nblocks + 7) / 8 + sms - 1) / sms * sms;
            if (c > cblocks) c = cblocks;
            for (size_t left = nblocks; left;) {
                const size_t t = left < c ? left : c;
                sizes.push_back(t);
                left -= t;
            }
        }
        size_t at = 0;
        for (size_t t : sizes) {
            start.push_back(at);
            at += t;
        }
        start.push_back(at);
    }
    const size_t nchunks = start.size() - 1;
    if (cblocks > nblocks) cblocks = nblocks;
    const size_t cbytes = cblocks * block_len < n ? cblocks * block_len : n; // largest chunk
    // device capacity per chunk: never more than the caller can take, never more than the dialect can produce
    size_t ccap = ljb_lz4_bound(cbytes, block_len);
    if (out_cap < ccap) ccap = out_cap;
    if ((rc = ljb_pipe_init(ctx, nchunks)) != 0) return rc;
    for (int i = 0; i < (nchunks > 1 ? 2 : 1); ++i) {
        if ((rc = ljb_ensure(&ctx->d_pin[i], &ctx->pin_bytes[i], cbytes + 64)) != 0) return rc;
        if ((rc = ljb_ensure(&ctx->d_pout[i], &ctx->pout_bytes[i], ccap + 64)) != 0) return rc;
    }
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, (nblocks + nchunks + 3 * nchunks + 8) * sizeof(uint64_t))) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;          // chunk k: its blocks + 1 entries at start[k] + k
    uint64_t *d_res = d_offs + nblocks + nchunks + 4;      // per chunk: 3 entries
    uint64_t *h_res = ctx->h_res;
    auto chunk_off = [&](size_t k) { return start[k] * block_len; };
    auto chunk_n = [&](size_t k) { return (k + 1 < nchunks ? start[k + 1] * block_len : n) - chunk_off(k); };
    size_t running = 0;
    uint64_t ph_total = 0;
    int status = LJB_OK;
    float kernel_ms = 0.f;
    cudaError_t e;
#define PIPE(x)                                                                 \
    do {                                                                        \
        e = (x);                                                                \
        if (e != cudaSuccess) {                                                 \
            status = ljb_set_cuda_error(e, #x, __LINE__);                       \
            goto done;                                                          \
        }                                                                       \
    } while (0)
    PIPE(cudaMemcpyAsync(ctx->d_pin[0], in, chunk_n(0), cudaMemcpyHostToDevice, ctx->s_in));
    PIPE(cudaEventRecord(ctx->ev_h2d[0], ctx->s_in));
    for (size_t k = 0; k < nchunks; ++k) {
        const int b = (int)(k & 1);
        const size_t kb = start[k + 1] - start[k];
        PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        if (k >= 2) PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0)); // output buffer b is free again
        size_t cap_k = out_cap - running < ccap ? out_cap - running : ccap;
        // `running` is known here (chunk k-1 has been waited for), so the kernel writes stream-global offsets itself
        rc = lz4_launch(ctx, (const uint8_t *)ctx->d_pin[b], chunk_n(k), block_len, (uint8_t *)ctx->d_pout[b], cap_k,
                        d_offs + start[k] + k, d_res + 3 * k, start[k], nblocks, nullptr, nullptr, running);
        if (rc != 0) {
            status = rc;
            goto done;
        }
        PIPE(cudaMemcpyAsync(h_res + 3 * k, d_res + 3 * k, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        PIPE(cudaEventRecord(ctx->ev_kern[b], ctx->stream));
        if (k + 1 < nchunks) { // next chunk's upload overlaps this chunk's kernel; its buffer was read by kernel k-1
            if (k >= 1) PIPE(cudaStreamWaitEvent(ctx->s_in, ctx->ev_kern[b ^ 1], 0));
            PIPE(cudaMemcpyAsync(ctx->d_pin[b ^ 1], in + chunk_off(k + 1), chunk_n(k + 1), cudaMemcpyHostToDevice, ctx->s_in));
            PIPE(cudaEventRecord(ctx->ev_h2d[b ^ 1], ctx->s_in));
        }
        PIPE(cudaEventSynchronize(ctx->ev_kern[b]));
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) kernel_ms += ms;
        }
        const uint64_t len_k = h_res[3 * k + 0];
        ph_total += h_res[3 * k + 1];
        if (h_res[3 * k + 2] & 1) {
            running += (size_t)len_k;
            status = LJB_E_CAPACITY;
            goto done;
        }
        PIPE(cudaStreamWaitEvent(ctx->s_out, ctx->ev_kern[b], 0));
        PIPE(cudaMemcpyAsync(out + running, ctx->d_pout[b], (size_t)len_k, cudaMemcpyDeviceToHost, ctx->s_out));
        if (block_offsets) // consecutive chunks overlap in one entry (end of k == start of k+1): the values agree
            PIPE(cudaMemcpyAsync(block_offsets + start[k], d_offs + start[k] + k, (kb + 1) * sizeof(uint64_t),
                                 cudaMemcpyDeviceToHost, ctx->s_out));
        PIPE(cudaEventRecord(ctx->ev_d2h[b], ctx->s_out));
        running += (size_t)len_k;
    }
    PIPE(cudaStreamSynchronize(ctx->s_out));
done:
#undef PIPE
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    ctx->last_kernel_ms = kernel_ms;
    ctx->kernel_ms_summed = 1;
    if (out_len) *out_len = running;
    if (phantom) *phantom = ph_total;
    return status;
}

extern "C" int ljb_lz4_block_matches(ljb_ctx *ctx, const uint8_t *in, size_t n, uint16_t *len, uint16_t *dist)
{
    if (!ctx || !in || !len || !dist || n == 0 || n > lz4k::MAXB) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t dcap = ljb_lz4_bound(n, n);
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], n + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], dcap + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, 8 * sizeof(uint64_t) + 4 * lz4k::MAXB)) != 0) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;
    uint64_t *d_res = d_offs + 2;
    uint16_t *d_len = (uint16_t *)(d_offs + 8);
    uint16_t *d_dist = d_len + lz4k::MAXB;
    LJB_CUDA(cudaMemcpyAsync(ctx->d_pin[0], in, n, cudaMemcpyHostToDevice, ctx->stream));
    rc = lz4_launch(ctx, (const uint8_t *)ctx->d_pin[0], n, n, (uint8_t *)ctx->d_pout[0], dcap, d_offs, d_res, 0, 1, d_len,
                    d_dist);
    if (rc != 0) return rc;
    LJB_CUDA(cudaMemcpyAsync(len, d_len, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(dist, d_dist, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}
#endif // !LJB_EMU_BUILD
