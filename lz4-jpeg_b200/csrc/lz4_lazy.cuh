// lz4-jpeg_b200/csrc/lz4_lazy.cuh — search + parse of one LZ4 block along the greedy chain only (included by lz4_encode.cu,
// inside namespace lz4k, after the shared-memory layout).
//
// block_encode() (Algorithms/sequential/LZ4/LZ4.c:516-583) asks find_longest_match() (LZ4.c:290-323) only at the positions
// its greedy parse visits: p -> p + (uint8_t)len for a match, p -> p + 1 otherwise.  On the benchmark text that is 11 k of
// the 65 k positions of a block once the runs of capped matches are skipped (below), and the candidates of those positions
// — the earlier positions with the same 4-gram — number 0.2 M per block, against 2.1 M for all positions.  The chain is
// sequential, so it is cut into walker segments of WG bytes that are parsed speculatively and stitched together:
//
//   index    counting sort of all positions by a hash of their 4-gram (8192 buckets, entries of a bucket ordered by
//            4096-position chunk): S[] holds every bucket contiguously, dir16[h] its end
//   round 0  walker s (one lane each) parses from the start of segment s to its end: at an unknown position it searches
//            (exactly: every earlier member of the bucket is compared, longest wins, earliest among the longest — the
//            strict '>' of LZ4.c:307), records (length, position) in R[p], sets known[p] and steps on
//   round r  walker s restarts from where the walk of segment s-1 left that segment, if that differs from where it started
//            before; it stops at the first known position (chains that meet stay together: the rest of the segment, and
//            its exit, are those of the walk that got there first).  Rounds repeat until no entry changes: then
//            entry(s) = exit(s-1) for all s, entry(0) = 0, which is the reference's chain.
//   Measured on the benchmark text: round 0 does 97 % of the searches, chains meet after a few steps, 2-3 rounds.
//
//   Rounds 0 and 1 have no CTA barrier between them (a walk needs one thing from another warp: the exit of the segment before
//   its own); blocks of low entropy take segments of WIDE x WG bytes, because chains of long matches rarely meet inside 66 bytes.
//
// A search is done by the warp for its 32 lanes together, three ways by the length of the candidate list (p50 = 3, p99 = 311 on
// text): short lists stay with their lane, lists up to 32 go to teams of eight lanes (four lists at a time), longer ones to the
// whole warp; every tier keeps two candidates per lane in flight (probe of the first 8 bytes, extension only for pairs that
// match them).  A bucket of more than MINFRONT entries keeps its earliest position in front: a search whose best pair reaches
// the cap there is over.
//
// Capped matches.  A match of MAX_MATCH = 1024 bytes is a literal step ((uint8_t)1024 == 0, LZ4.c:317) whose distance is
// never used, and if (c, p) match 1024 + j bytes then (c + j', p + j') match >= 1024 for j' <= j: those positions are
// literal steps too, whatever their earliest best candidate is.  The walker marks them known without searching (the
// benchmark text, 30 000-byte passages of a 118 KB corpus, has 10 k .. 30 k such positions per block).
#pragma once

#ifdef LJB_EMU // (the CPU emulation runs one fiber at a time: a spinning lane hands over)
#define LJB_SPIN() emu_spin_yield()
#define __threadfence_block() ((void)0)
#else
#define LJB_SPIN() ((void)0)
#endif
constexpr int WG = 66;                              // bytes per walker segment: two per emission segment
constexpr int NWALK_MAX = (MAXB + WG - 1) / WG;     // 993
constexpr int LCH = 12;                             // the index keeps a bucket ordered by 4096-position chunk
#ifndef LJB_TINYLIST
#define LJB_TINYLIST 4
#endif
#ifndef LJB_TINYMAX
#define LJB_TINYMAX 8
#endif
#ifndef LJB_TINYCROWD
#define LJB_TINYCROWD 4
#endif
#ifndef LJB_BIGLIST
#define LJB_BIGLIST 32
#endif
#ifndef LJB_VLONG
#define LJB_VLONG 32
#endif
constexpr int TINYMAX = LJB_TINYMAX;                          // ... up to this length when TINYCROWD or more lanes of the warp have such a list (high-entropy data: every lane has one)
constexpr int TINYCROWD = LJB_TINYCROWD;
constexpr int TINYLIST = LJB_TINYLIST;                          // candidate lists up to this length are walked by the lane that owns them
constexpr int BIGLIST = LJB_BIGLIST;                          // candidate lists this long are taken by the whole warp, shorter ones by a team of eight lanes
constexpr uint32_t HOT_BUILD = 2048;                 // a block in which some bucket holds more entries than this also gets an index by 8-gram (in L2)
constexpr uint32_t HOT_USE = 512;                    // candidate lists longer than this look at the 8-gram bucket first
#ifndef LJB_MINFRONT
#define LJB_MINFRONT 1024
#endif
#ifndef LJB_WIDE
#define LJB_WIDE 8
#endif
constexpr uint32_t WIDE = LJB_WIDE;                   // walker segments of low-entropy blocks: this many times WG (even)
static_assert(WIDE % 2 == 0 && WIDE >= 2, "a wide walker segment is a whole number of emission segments");
constexpr int MINFRONT = LJB_MINFRONT;                         // buckets with more entries than this keep their earliest position in front
constexpr uint32_t VLONG = LJB_VLONG;                      // a lane compares this much on its own; longer runs are compared by the whole warp
static_assert(2 * WG == SEG, "an emission segment is two walker segments");
static_assert(NWALK_MAX <= THREADS, "one lane per walker");
// walker state in the area the full search uses for its first-occurrence bits
constexpr int SM_XIN = SM_FIRST;                    // u16 xin[1024]   where the latest walk of the segment started (relative to the segment start)
constexpr int SM_XOUT = SM_FIRST + 2048;            // u16 xout[1024]  where it left the segment (relative to the segment start)
constexpr int SM_XCAN = SM_FIRST + 4096;            // u16 xcan[1024]  exit of the segment's first walk
constexpr int SM_SPLIT = SM_FIRST + 6144;           // u8 split[1024]  a later walk left the segment without meeting an earlier one
static_assert(SM_SPLIT + 1024 <= SM_MISC, "walker state must fit its area");

template <class Phase, class Flush>
__device__ __forceinline__ void lazy_search(uint8_t *smem, const uint32_t nb, uint32_t *R, const Params &P, Misc &M, Phase &&phase,
                                            Flush &&flush_prev)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint8_t *data = smem + SM_DATA;
    const uint32_t *dataw = reinterpret_cast<const uint32_t *>(data);
    uint16_t *S = reinterpret_cast<uint16_t *>(smem + SM_S);
    uint32_t *dirw = reinterpret_cast<uint32_t *>(smem + SM_DIR);
    const uint16_t *dir16 = reinterpret_cast<const uint16_t *>(dirw);
    uint32_t *known = reinterpret_cast<uint32_t *>(smem + SM_LONG);
    uint16_t *xin = reinterpret_cast<uint16_t *>(smem + SM_XIN);
    uint16_t *xout = reinterpret_cast<uint16_t *>(smem + SM_XOUT);
    uint8_t *const stepg = reinterpret_cast<uint8_t *>(P.gids + (size_t)blockIdx.x * MAXB); // per CTA: the step of every visited position
    uint8_t *step = smem + SM_STEP;
    uint8_t *entry = smem + SM_ENTRY;
    const uint32_t npos = nb >= 4 ? nb - 3 : 0; // positions that still have a 4-gram inside the block

    // ---------------- index: counting sort of the positions by 4-gram hash ----------------
    for (int i = tid; i < NBUCKET / 2 + 4; i += THREADS) dirw[i] = 0;
    for (int i = tid; i < MAXB / 32; i += THREADS) known[i] = 0;
    __syncthreads();
    // a thread takes four consecutive positions per 4096-position chunk (two words of the block give their four 4-grams)
    for (uint32_t base = 0; base < npos; base += 4 * THREADS) {
        const uint32_t q0 = base + 4u * (uint32_t)tid;
        if (q0 < npos) {
            const uint32_t w0 = dataw[q0 >> 2], w1 = dataw[(q0 >> 2) + 1];
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = q0 + j < npos ? hash4(__funnelshift_r(w0, w1, 8 * j)) : 0xFFFFFFFFu;
            if (h[0] == h[1] && h[1] == h[2] && h[2] == h[3]) { // a run of one byte: one add for the four (65 k adds to one counter otherwise)
                atomicAdd(&dirw[h[0] >> 1], (h[0] & 1) ? 0x40000u : 4u);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (h[j] != 0xFFFFFFFFu) atomicAdd(&dirw[h[j] >> 1], (h[j] & 1) ? 0x10000u : 1u);
            }
        }
    }
    __syncthreads();
    uint32_t maxc = 0; // the largest bucket among this thread's counters
    {
        // exclusive scan of the 8192 u16 counts (two per word); every thread owns four consecutive words
        constexpr uint32_t cw = (NBUCKET / 2) / THREADS;
        static_assert(cw * THREADS == NBUCKET / 2, "directory words divide evenly among the threads");
        const uint32_t w0 = (uint32_t)tid * cw;
        uint32_t wv[cw], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < cw; ++k) {
            wv[k] = dirw[w0 + k];
            sum += (wv[k] & 0xFFFF) + (wv[k] >> 16);
            maxc = max(maxc, max(wv[k] & 0xFFFFu, wv[k] >> 16));
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) M.scan_tmp[tid >> 5] = inc;
        __syncthreads();
        uint32_t run = inc - sum;
        for (int k = 0; k < (tid >> 5); ++k) run += M.scan_tmp[k];
#pragma unroll
        for (uint32_t k = 0; k < cw; ++k) {
            const uint32_t lo = run;
            run += wv[k] & 0xFFFF;
            const uint32_t hi = run;
            run += wv[k] >> 16;
            dirw[w0 + k] = (lo & 0xFFFF) | (hi << 16);
        }
    }
    __syncthreads();
    // scatter in position-ordered rounds of 4096 positions: afterwards dir16[h] is the END of bucket h
    for (uint32_t base = 0; base < npos; base += 4 * THREADS) {
        const uint32_t q0 = base + 4u * (uint32_t)tid;
        if (q0 < npos) {
            const uint32_t w0 = dataw[q0 >> 2], w1 = dataw[(q0 >> 2) + 1];
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = q0 + j < npos ? hash4(__funnelshift_r(w0, w1, 8 * j)) : 0xFFFFFFFFu;
            if (h[0] == h[1] && h[1] == h[2] && h[2] == h[3]) {
                const uint32_t old = atomicAdd(&dirw[h[0] >> 1], (h[0] & 1) ? 0x40000u : 4u);
                const uint32_t at = (h[0] & 1) ? (old >> 16) : (old & 0xFFFF);
#pragma unroll
                for (int j = 0; j < 4; ++j) S[at + j] = (uint16_t)(q0 + j);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (h[j] != 0xFFFFFFFFu) {
                        const uint32_t old = atomicAdd(&dirw[h[j] >> 1], (h[j] & 1) ? 0x10000u : 1u);
                        S[(h[j] & 1) ? (old >> 16) : (old & 0xFFFF)] = (uint16_t)(q0 + j);
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- big buckets: the EARLIEST position goes to the front (the entries of a chunk are in no particular order).  A search whose
    // best pair reaches the cap at that position is over: no other member of the bucket lies before it.  Blocks of one byte, or
    // of a short period, have one or two buckets of tens of thousands of entries and every search ends at their first entry.
    {
        constexpr uint32_t per_warp = NBUCKET / (THREADS / 32);
        for (uint32_t b0 = (uint32_t)(tid >> 5) * per_warp; b0 < (uint32_t)(tid >> 5) * per_warp + per_warp; b0 += 32) {
            const uint32_t h = b0 + (uint32_t)lane;
            const uint32_t lo = h ? dir16[h - 1] : 0u, hi = dir16[h];
            unsigned bigm = __ballot_sync(FULL, hi - lo > (uint32_t)MINFRONT);
            while (bigm) {
                const int src = __ffs(bigm) - 1;
                bigm &= bigm - 1;
                const uint32_t blo = __shfl_sync(FULL, lo, src), bhi = __shfl_sync(FULL, hi, src);
                const uint32_t ch0 = (uint32_t)S[blo] >> LCH;
                uint32_t mn = 0xFFFFFFFFu; // (position << 16) | index in the bucket
                for (uint32_t base = 0;; base += 32) {
                    const uint32_t i = base + (uint32_t)lane;
                    uint32_t v = 0;
                    const bool in = blo + i < bhi && ((v = S[blo + i]) >> LCH) == ch0; // the earliest position is in the bucket's first chunk
                    if (in) mn = min(mn, (v << 16) | i);
                    if (!__all_sync(FULL, in)) break;
                }
                mn = __reduce_min_sync(FULL, mn);
                if (lane == 0 && (mn & 0xFFFFu) != 0u) {
                    const uint16_t first = S[blo];
                    S[blo] = (uint16_t)(mn >> 16);
                    S[blo + (mn & 0xFFFFu)] = first;
                }
            }
        }
    }
    // ---- low-entropy blocks (a 4-gram that occurs thousands of times: two-symbol data, long runs of a short period) also get an
    // index by 8-GRAM, in L2: a position whose 4-gram list is long looks at its 8-gram bucket first — every candidate that
    // matches 8 bytes or more is in there, and if there is one the answer is among them (longer beats shorter).  Only when
    // nothing matches 8 bytes is the long 4-gram list walked.  Two-symbol random data: 128 candidates per search instead of 2000.
    uint32_t *const dir8 = P.idx8 + (size_t)blockIdx.x * (NBUCKET + MAXB / 2); // u32 dir8[NBUCKET] | u16 S8[MAXB]
    uint16_t *const S8 = reinterpret_cast<uint16_t *>(dir8 + NBUCKET);
    // (not when a bucket holds a quarter of the block: runs of one byte or of a period up to four have as few 8-grams as 4-grams)
    const int hotbits = __syncthreads_or((maxc > HOT_BUILD ? 1 : 0) | (maxc > (uint32_t)MAXB / 4 ? 2 : 0));
    const bool hot = hotbits == 1;
    if (hot) {
        const uint32_t npos8 = nb >= 8 ? nb - 7 : 0;
        for (int i = tid; i < NBUCKET; i += THREADS) dir8[i] = 0;
        __syncthreads();
        for (uint32_t base = 0; base < npos8; base += 4 * THREADS) {
            const uint32_t q0 = base + 4u * (uint32_t)tid;
            if (q0 < npos8) {
                const uint32_t w0 = dataw[q0 >> 2], w1 = dataw[(q0 >> 2) + 1], w2 = dataw[(q0 >> 2) + 2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q0 + j < npos8) atomicAdd(&dir8[hash8(__funnelshift_r(w0, w1, 8 * j), __funnelshift_r(w1, w2, 8 * j))], 1u);
            }
        }
        __syncthreads();
        {
            constexpr uint32_t per = NBUCKET / THREADS; // 8 counters per thread
            uint32_t cnt[per], sum = 0;
#pragma unroll
            for (uint32_t k = 0; k < per; ++k) {
                cnt[k] = __ldcg(&dir8[(uint32_t)tid * per + k]);
                sum += cnt[k];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) M.scan_tmp[tid >> 5] = inc;
            __syncthreads();
            uint32_t run = inc - sum;
            for (int k = 0; k < (tid >> 5); ++k) run += M.scan_tmp[k];
#pragma unroll
            for (uint32_t k = 0; k < per; ++k) {
                dir8[(uint32_t)tid * per + k] = run;
                run += cnt[k];
            }
        }
        __syncthreads();
        for (uint32_t base = 0; base < npos8; base += 4 * THREADS) {
            const uint32_t q0 = base + 4u * (uint32_t)tid;
            if (q0 < npos8) {
                const uint32_t w0 = dataw[q0 >> 2], w1 = dataw[(q0 >> 2) + 1], w2 = dataw[(q0 >> 2) + 2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (q0 + j < npos8)
                        S8[atomicAdd(&dir8[hash8(__funnelshift_r(w0, w1, 8 * j), __funnelshift_r(w1, w2, 8 * j))], 1u)] = (uint16_t)(q0 + j);
            }
        }
        __syncthreads(); // dir8[h] is now the END of bucket h
    }
    phase(3); // index

    // ---------------- walk rounds ----------------
    // Consecutive segments go to different warps (lane l of warp w parses segment 32 l + w): the expensive stretches of a block
    // (fresh text) and the cheap ones (capped runs) are then spread over all warps — with contiguous segments per warp the round
    // waited for its slowest warp for a quarter of the kernel.
    // Low-entropy blocks (the ones that got the 8-gram index) take walker segments of WIDE x WG bytes: their matches are long (15 bytes
    // on two-symbol data) and two chains that enter a 66-byte segment at different offsets rarely meet inside it, so every round
    // corrected one more segment — half of such a block's cycles went to rounds >= 1.  (tune bit 1 forces, bit 2 forbids.)
    const bool wide = ((hot && !(P.tune & 4u)) || (P.tune & 2u)) && nb > WIDE * (uint32_t)WG;
    const uint32_t wg = wide ? WIDE * (uint32_t)WG : (uint32_t)WG;
    const uint32_t nwalk = (nb + wg - 1) / wg;
    const uint32_t s = (P.tune & 1u) ? (uint32_t)tid : (uint32_t)lane * 32u + (uint32_t)(tid >> 5); // (tune bit 0: contiguous segments per warp, for measurements)
    const bool has = s < nwalk;
    const uint32_t seg0 = s * wg, segend = min(seg0 + wg, nb);
    uint32_t my_in = 0, my_can = 0; // (own lane only) where the latest walk of the segment started; exit of its first walk
    bool my_split = false;          // a later walk left the segment without meeting an earlier one
    uint32_t d_search = 0, d_cand = 0, d_steps = 0, d_vl = 0, d_run = 0;
    // One candidate in two parts, so that two candidates of a lane can be in flight together:
    //   probe   (branch-free) its first 8 bytes against the position's: the pair's key if it matches fewer than 8 bytes, or `lng`
    //   extend  (only for `lng`) up to VLONG bytes by the lane itself; pairs still matching then are left to the warp (`vl`)
    // key = (length << 16) | (0xFFFF - position): the maximum is the longest match, the earliest among equals.  A candidate of
    // 0xFFFF stands for "none" (it is never below the position).
    auto probe = [&](uint32_t c, uint32_t op, uint32_t og0, uint32_t og1, uint32_t ocap, uint32_t &a2, bool &lng) -> uint32_t {
        const uint32_t ci = c >> 2, cs = (c & 3) * 8;
        const uint32_t a0 = dataw[ci], a1 = dataw[ci + 1];
        a2 = dataw[ci + 2];
        const uint32_t x0 = __funnelshift_r(a0, a1, cs) ^ og0, x1 = __funnelshift_r(a1, a2, cs) ^ og1;
        const bool ok = c < op && x0 == 0u; // (a bucket holds several 4-grams)
        lng = ok && x1 == 0u;
        const uint32_t l = min(4u + ((uint32_t)(__ffs(x1) - 1) >> 3), ocap);
        return (ok && x1 != 0u) ? ((l << 16) | (0xFFFFu - c)) : 0u;
    };
    auto extend = [&](uint32_t c, uint32_t op, uint32_t ocap, uint32_t a2, uint32_t tb, bool &vl) -> uint32_t {
        // a pair that already reaches the cap is only beaten by an earlier position
        if ((tb >> 16) == ocap && c > 0xFFFFu - (tb & 0xFFFFu)) return 0u;
        uint32_t l = 8;
        const uint32_t lim = min(ocap, VLONG);
        // 8 bytes per step, the word straddling the step carried over on either side
        const uint32_t ci = c >> 2, cs = (c & 3) * 8, pi = op >> 2, ps = (op & 3) * 8;
        uint32_t cw = a2, pw = dataw[pi + 2];
        while (l < lim) {
            const uint32_t c1 = dataw[ci + (l >> 2) + 1], c2 = dataw[ci + (l >> 2) + 2];
            const uint32_t p1 = dataw[pi + (l >> 2) + 1], p2 = dataw[pi + (l >> 2) + 2];
            uint32_t x = __funnelshift_r(cw, c1, cs) ^ __funnelshift_r(pw, p1, ps);
            if (x) {
                l += (uint32_t)(__ffs(x) - 1) >> 3;
                break;
            }
            x = __funnelshift_r(c1, c2, cs) ^ __funnelshift_r(p1, p2, ps);
            if (x) {
                l += 4u + ((uint32_t)(__ffs(x) - 1) >> 3);
                break;
            }
            cw = c2;
            pw = p2;
            l += 8;
        }
        if (l >= VLONG && ocap > VLONG) { // still matching after VLONG bytes: the warp goes on
            vl = true;
            return 0u;
        }
        return (min(l, ocap) << 16) | (0xFFFFu - c);
    };
    auto eval = [&](uint32_t c, uint32_t op, uint32_t og0, uint32_t og1, uint32_t ocap, uint32_t tb, bool &vl) -> uint32_t {
        uint32_t a2;
        bool lng;
        uint32_t key = probe(c, op, og0, og1, ocap, a2, lng);
        if (lng) key = extend(c, op, ocap, a2, tb, vl);
        return key;
    };
    // The pair (bc, bp) matches VLONG bytes: the warp compares on, 256 bytes per step, unless the pair cannot beat btb (it must
    // exceed the best length, or tie it from an earlier position).  Warp-uniform.
    auto long_compare = [&](uint32_t bc, uint32_t bp, uint32_t bcap, uint32_t btb) -> uint32_t {
        const uint32_t bl = btb >> 16, bpos = 0xFFFFu - (btb & 0xFFFFu);
        const uint32_t needl = btb == 0u ? 0u : (bc < bpos ? bl : bl + 1u);
        if (needl > bcap || (needl > VLONG && data[bc + needl - 1] != data[bp + needl - 1])) return 0u;
        ++d_vl;
        uint32_t res = bcap;
        for (uint32_t base = VLONG; base < bcap; base += 256) {
            const uint32_t off = base + 8u * (uint32_t)lane;
            uint32_t x0 = 0, x1 = 0;
            if (off < bcap) {
                const uint32_t ci = (bc + off) >> 2, cs = ((bc + off) & 3) * 8;
                const uint32_t pi = (bp + off) >> 2, ps = ((bp + off) & 3) * 8;
                const uint32_t a0 = dataw[ci], a1 = dataw[ci + 1], a2 = dataw[ci + 2];
                const uint32_t e0 = dataw[pi], e1 = dataw[pi + 1], e2 = dataw[pi + 2];
                x0 = __funnelshift_r(a0, a1, cs) ^ __funnelshift_r(e0, e1, ps);
                x1 = __funnelshift_r(a1, a2, cs) ^ __funnelshift_r(e1, e2, ps);
            }
            const unsigned mm = __ballot_sync(FULL, (x0 | x1) != 0u);
            if (mm) {
                const int f = __ffs(mm) - 1;
                const uint32_t mine = off + (x0 ? ((uint32_t)(__ffs(x0) - 1) >> 3) : 4u + ((uint32_t)(__ffs(x1) - 1) >> 3));
                res = min(__shfl_sync(FULL, mine, f), bcap);
                break;
            }
        }
        return (res << 16) | (0xFFFFu - bc);
    };
    // Rounds 0 and 1 run without a CTA barrier between them: what a walker of round 1 needs from the others is the exit of the
    // segment before its own, and that segment belongs to the same lane of the warp before (lane - 1 of the last warp for the
    // first one).  A warp that has written its exits of round 0 says so in wdone[] and goes on as soon as the warp before it has
    // said the same: the re-walks of the fast warps fill the time they would have waited for the slowest one.  (Everything else
    // a walk touches — known bits, steps, records — belongs to its own segment.)
#ifndef LJB_FUSE_ROUNDS
#define LJB_FUSE_ROUNDS 1
#endif
    volatile unsigned int *const wdone = M.wdone;
    const unsigned int epoch = (unsigned int)M.ticket + 1u; // (tickets are unique: no flag has to be cleared between blocks)
    for (uint32_t round = 0;; ++round) {
        bool run = false;
        uint32_t p = 0;
        if (LJB_FUSE_ROUNDS && round == 1) {
            if (lane == 0)
                while (wdone[((tid >> 5) + 31) & 31] != epoch) LJB_SPIN();
            __syncwarp();
            __threadfence_block();
        }
        if (has) {
            if (round == 0) {
                run = true;
                p = seg0;
            } else if (s > 0) {
                const uint32_t e = (uint32_t)xout[s - 1] + (s - 1) * wg; // where the chain leaves the segment before this one
                if (e != my_in) {
                    run = true;
                    p = e;
                }
            }
        }
        if (!(LJB_FUSE_ROUNDS && round == 1) && !__syncthreads_or(run)) break; // (also separates the reads above from this round's writes)
        if (run) my_in = p;
        uint32_t ex = p;                   // a chain that jumps over the segment leaves it where it enters it
        bool act = run && p < segend;
        bool met = false;                  // stopped at a known position
        const bool walked = act;
        while (__any_sync(FULL, act)) {
            ++d_steps;
            bool need = false;
            if (act) {
                if ((known[p >> 5] >> (p & 31)) & 1u) {
                    met = true;
                    if (!my_split) { // every known position of the segment leads to the exit of its first walk
                        ex = my_can;
                        act = false;
                    } else { // two chains cross this segment side by side: follow the recorded steps
                        const uint32_t st = __ldcg(&stepg[p]);
                        p += st ? st : 1u;
                        if (p >= segend) {
                            ex = p;
                            act = false;
                        }
                    }
                } else {
                    need = true;
                }
            }
            // ---- the candidates of this lane's position: the members of its bucket in earlier chunks and in its own
            uint32_t g0 = 0, g1 = 0, lo = 0, n = 0;
            const uint32_t cap = min((uint32_t)MAX_MATCH, nb - p);
            if (need && p < npos) {
                g0 = load32u(dataw, p);
                g1 = load32u(dataw, p + 4);
                const uint32_t h = hash4(g0);
                lo = h ? dir16[h - 1] : 0u;
                const uint32_t hi = dir16[h];
#ifndef LJB_SHORTCUT
#define LJB_SHORTCUT TINYLIST
#endif
                if (hi - lo <= (uint32_t)(LJB_SHORTCUT)) {
                    n = hi - lo; // (later positions are turned away one by one)
                } else {
                    const uint32_t pch = p >> LCH;
                    uint32_t a = lo, b = hi;
                    while (a < b) { // first entry of a later chunk
                        const uint32_t m = (a + b) >> 1;
                        if ((uint32_t)(S[m] >> LCH) <= pch) a = m + 1;
                        else b = m;
                    }
                    n = a - lo;
                }
            }
            if (need) {
                ++d_search;
                d_cand += n;
            }
            // ---- evaluation, three ways by the length of the list (p50 = 3, p90 = 52, p99 = 311 candidates on the benchmark
            // text): up to TINYLIST every lane walks its own list; up to BIGLIST teams of eight lanes take a list each, four at a
            // time; longer ones are taken by the whole warp, one at a time.
            uint32_t best = 0;
            // (i) short lists: every lane its own
            bool taken = false; // this lane's list was dealt with in (i)
            {
                // teams finish a list of 5 .. 8 in one pass, four lists at a time; when many lanes hold one, every lane walking its own is faster
                const uint32_t tiny = __popc(__ballot_sync(FULL, need && n > (uint32_t)TINYLIST && n <= (uint32_t)TINYMAX)) >= TINYCROWD ? (uint32_t)TINYMAX : (uint32_t)TINYLIST;
                const bool mine = need && n > 0 && n <= tiny;
                taken = mine;
                const uint32_t nmax = __reduce_max_sync(FULL, mine ? n : 0u);
                for (uint32_t it = 0; it < nmax; it += 2) { // two entries per iteration: their loads are in flight together
                    const uint32_t c0 = (mine && it < n) ? (uint32_t)S[lo + it] : 0xFFFFu, c1 = (mine && it + 1 < n) ? (uint32_t)S[lo + it + 1] : 0xFFFFu;
                    uint32_t a20, a21;
                    bool lng0, lng1, vl0 = false, vl1 = false;
                    uint32_t key0 = probe(c0, p, g0, g1, cap, a20, lng0), key1 = probe(c1, p, g0, g1, cap, a21, lng1);
                    if (lng0) key0 = extend(c0, p, cap, a20, best, vl0);
                    if (lng1) key1 = extend(c1, p, cap, a21, best, vl1);
                    best = max(best, max(key0, key1));
                    if (__any_sync(FULL, vl0 || vl1)) {
                        unsigned pendv = __ballot_sync(FULL, vl0);
                        while (pendv) {
                            const int sv = __ffs(pendv) - 1;
                            pendv &= pendv - 1;
                            const uint32_t rkey = long_compare(__shfl_sync(FULL, c0, sv), __shfl_sync(FULL, p, sv), __shfl_sync(FULL, cap, sv),
                                                               __shfl_sync(FULL, best, sv));
                            if (lane == sv) best = max(best, rkey);
                        }
                        pendv = __ballot_sync(FULL, vl1);
                        while (pendv) {
                            const int sv = __ffs(pendv) - 1;
                            pendv &= pendv - 1;
                            const uint32_t rkey = long_compare(__shfl_sync(FULL, c1, sv), __shfl_sync(FULL, p, sv), __shfl_sync(FULL, cap, sv),
                                                               __shfl_sync(FULL, best, sv));
                            if (lane == sv) best = max(best, rkey);
                        }
                    }
                }
            }
            // (ii) long lists: the whole warp
            unsigned big = __ballot_sync(FULL, need && n >= (uint32_t)BIGLIST);
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const uint32_t op = __shfl_sync(FULL, p, src), og0 = __shfl_sync(FULL, g0, src), og1 = __shfl_sync(FULL, g1, src);
                const uint32_t olo = __shfl_sync(FULL, lo, src), on = __shfl_sync(FULL, n, src);
                const uint32_t ocap = min((uint32_t)MAX_MATCH, nb - op);
                uint32_t lb = 0, tb = 0; // this lane's best; what the whole warp knows (pairs at the cap, long compares)
                bool capped = false;     // tb has reached the cap
                // the bucket's earliest position as the low half of a key (only buckets of more than MINFRONT entries keep it in front)
                const uint32_t ofront = on > (uint32_t)MINFRONT ? 0xFFFFu - (uint32_t)S[olo] : 0xFFFFFFFFu;
                if (hot && on > HOT_USE && ocap >= 8u) { // the 8-gram bucket first (entries in no particular order)
                    const uint32_t h8 = hash8(og0, og1);
                    const uint32_t lo8 = h8 ? __ldcg(&dir8[h8 - 1]) : 0u, n8 = __ldcg(&dir8[h8]) - lo8;
                    if (n8 <= HOT_USE) {
                        uint32_t lb8 = 0, tb8 = 0;
                        for (uint32_t it = 0; it < n8; it += 32) {
                            const uint32_t i = it + (uint32_t)lane;
                            uint32_t key = 0, c = 0;
                            bool vl = false;
                            if (i < n8) {
                                c = __ldcg(&S8[lo8 + i]);
                                key = eval(c, op, og0, og1, ocap, tb8, vl);
                            }
                            lb8 = max(lb8, key);
                            if (__any_sync(FULL, vl || (key >> 16) == ocap)) {
                                unsigned pendv = __ballot_sync(FULL, vl);
                                while (pendv) {
                                    const int sv = __ffs(pendv) - 1;
                                    pendv &= pendv - 1;
                                    tb8 = max(tb8, long_compare(__shfl_sync(FULL, c, sv), op, ocap, tb8));
                                }
                                tb8 = max(tb8, __reduce_max_sync(FULL, key));
                            }
                        }
                        tb8 = max(tb8, __reduce_max_sync(FULL, lb8));
                        if ((tb8 >> 16) >= 8u) { // every pair that matches 8 bytes was in that bucket: this is the answer
                            if (lane == src) best = tb8;
                            continue;
                        }
                    }
                }
                for (uint32_t it = 0; it < on; it += 64) { // two entries per lane: their loads are in flight together
                    const uint32_t i0 = it + (uint32_t)lane, i1 = i0 + 32u;
                    const uint32_t c0 = i0 < on ? (uint32_t)S[olo + i0] : 0xFFFFu, c1 = i1 < on ? (uint32_t)S[olo + i1] : 0xFFFFu;
                    uint32_t a20, a21;
                    bool lng0, lng1, vl0 = false, vl1 = false;
                    uint32_t key0 = probe(c0, op, og0, og1, ocap, a20, lng0), key1 = probe(c1, op, og0, og1, ocap, a21, lng1);
                    if (lng0) key0 = extend(c0, op, ocap, a20, tb, vl0);
                    if (lng1) key1 = extend(c1, op, ocap, a21, tb, vl1);
                    const uint32_t key = max(key0, key1);
                    lb = max(lb, key);
                    if (__any_sync(FULL, vl0 || vl1 || (key >> 16) == ocap)) { // rare on text: a pair at the cap, or one for the warp
                        unsigned pendv = __ballot_sync(FULL, vl0);
                        while (pendv && tb != ((ocap << 16) | ofront)) { // (the cap at the earliest position of all: nothing beats it)
                            const int sv = __ffs(pendv) - 1;
                            pendv &= pendv - 1;
                            tb = max(tb, long_compare(__shfl_sync(FULL, c0, sv), op, ocap, tb));
                        }
                        pendv = __ballot_sync(FULL, vl1);
                        while (pendv && tb != ((ocap << 16) | ofront)) {
                            const int sv = __ffs(pendv) - 1;
                            pendv &= pendv - 1;
                            tb = max(tb, long_compare(__shfl_sync(FULL, c1, sv), op, ocap, tb));
                        }
                        tb = max(tb, __reduce_max_sync(FULL, key));
                        capped = (tb >> 16) == ocap;
                    }
                    // once the cap is reached only earlier positions matter: the entries of later chunks cannot win
                    if (capped && it + 64 < on && ((tb & 0xFFFFu) == ofront || (uint32_t)(S[olo + it + 64] >> LCH) > ((0xFFFFu - (tb & 0xFFFFu)) >> LCH))) break;
                }
                tb = max(tb, __reduce_max_sync(FULL, lb));
                if (lane == src) best = tb;
            }
            // (iii) the lists in between: teams
            unsigned rem = __ballot_sync(FULL, need && !taken && n > 0u && n < (uint32_t)BIGLIST);
            while (rem) {
                const int b0 = __ffs(rem) - 1;
                rem &= rem - 1;
                const int b1 = __ffs(rem) - 1;
                rem &= rem - 1;
                const int b2 = __ffs(rem) - 1;
                rem &= rem - 1;
                const int b3 = __ffs(rem) - 1;
                rem &= rem - 1;
                const int team = lane >> 3, tl = lane & 7;
                const int src = team == 0 ? b0 : (team == 1 ? b1 : (team == 2 ? b2 : b3));
                const int srcl = src < 0 ? 0 : src;
                const uint32_t op = __shfl_sync(FULL, p, srcl), og0 = __shfl_sync(FULL, g0, srcl), og1 = __shfl_sync(FULL, g1, srcl);
                const uint32_t olo = __shfl_sync(FULL, lo, srcl);
                uint32_t on = __shfl_sync(FULL, n, srcl);
                if (src < 0) on = 0;
                const uint32_t ocap = min((uint32_t)MAX_MATCH, nb - op);
                uint32_t lb = 0, tb = 0; // this lane's best; what the whole team knows (the same in its eight lanes)
                uint32_t nit = __reduce_max_sync(FULL, on);
                for (uint32_t it = 0; it < nit; it += 16) { // two entries per lane
                    const uint32_t i0 = it + (uint32_t)tl, i1 = i0 + 8u;
                    const uint32_t c0 = i0 < on ? (uint32_t)S[olo + i0] : 0xFFFFu, c1 = i1 < on ? (uint32_t)S[olo + i1] : 0xFFFFu;
                    uint32_t a20, a21;
                    bool lng0, lng1, vl0 = false, vl1 = false;
                    uint32_t key0 = probe(c0, op, og0, og1, ocap, a20, lng0), key1 = probe(c1, op, og0, og1, ocap, a21, lng1);
                    if (lng0) key0 = extend(c0, op, ocap, a20, tb, vl0);
                    if (lng1) key1 = extend(c1, op, ocap, a21, tb, vl1);
                    uint32_t key = max(key0, key1);
                    lb = max(lb, key);
                    if (__any_sync(FULL, vl0 || vl1 || (key >> 16) == ocap)) {
                        for (int half = 0; half < 2; ++half) {
                            unsigned pendv = __ballot_sync(FULL, half ? vl1 : vl0);
                            const uint32_t cc = half ? c1 : c0;
                            while (pendv) {
                                const int sv = __ffs(pendv) - 1;
                                pendv &= pendv - 1;
                                const uint32_t rkey = long_compare(__shfl_sync(FULL, cc, sv), __shfl_sync(FULL, op, sv), __shfl_sync(FULL, ocap, sv),
                                                                   __shfl_sync(FULL, tb, sv));
                                if ((lane >> 3) == (sv >> 3)) tb = max(tb, rkey); // known to the whole team at once
                            }
                        }
                        key = max(key, __shfl_xor_sync(FULL, key, 4));
                        key = max(key, __shfl_xor_sync(FULL, key, 2));
                        key = max(key, __shfl_xor_sync(FULL, key, 1));
                        tb = max(tb, key);
                        if (it + 16 < on && (tb >> 16) == ocap && (uint32_t)(S[olo + it + 16] >> LCH) > ((0xFFFFu - (tb & 0xFFFFu)) >> LCH)) on = 0;
                        nit = __reduce_max_sync(FULL, on); // (the branch is warp-uniform)
                    }
                }
                lb = max(lb, __shfl_xor_sync(FULL, lb, 4));
                lb = max(lb, __shfl_xor_sync(FULL, lb, 2));
                lb = max(lb, __shfl_xor_sync(FULL, lb, 1));
                tb = max(tb, lb);
                const uint32_t v0 = __shfl_sync(FULL, tb, 0), v1 = __shfl_sync(FULL, tb, 8), v2 = __shfl_sync(FULL, tb, 16), v3 = __shfl_sync(FULL, tb, 24);
                if (lane == b0) best = v0;
                if (lane == b1) best = v1;
                if (lane == b2) best = v2;
                if (lane == b3) best = v3;
            }
            // ---- record: the step (the reference's uint8_t cast of the length, LZ4.c:317; 0 = literal step) and, for a match, where
            const uint32_t blen = best >> 16, bstep = blen & 0xFFu;
            if (need) {
                stepg[p] = (uint8_t)bstep;
                if (bstep) R[p] = (blen << 16) | (0xFFFFu - (best & 0xFFFFu));
                atomicOr(&known[p >> 5], 1u << (p & 31));
            }
            // ---- capped matches: the following positions of the segment that inherit a 1024-byte match are literal steps
            uint32_t skip = 0;
            unsigned runs = __ballot_sync(FULL, need && blen == (uint32_t)MAX_MATCH);
            while (runs) {
                const int sr = __ffs(runs) - 1;
                runs &= runs - 1;
                const uint32_t rp = __shfl_sync(FULL, p, sr), rc = 0xFFFFu - (__shfl_sync(FULL, best, sr) & 0xFFFFu);
                const uint32_t rend = __shfl_sync(FULL, segend, sr);
                // position rp + j has a 1024-byte match at rc + j if the pair (rc, rp) goes on for j more bytes and rp + j + 1024 <= nb
                const uint32_t lim = min(rend - 1u - rp, nb - (uint32_t)MAX_MATCH - rp);
                uint32_t k = lim;
                for (uint32_t j0 = 0; j0 < lim; j0 += 32) {
                    const uint32_t j = j0 + (uint32_t)lane + 1u;
                    const bool mis = j <= lim && data[rc + (uint32_t)MAX_MATCH - 1u + j] != data[rp + (uint32_t)MAX_MATCH - 1u + j];
                    const unsigned mm = __ballot_sync(FULL, mis);
                    if (mm) {
                        k = j0 + (uint32_t)(__ffs(mm) - 1);
                        break;
                    }
                }
                for (uint32_t j = (uint32_t)lane + 1u; j <= k; j += 32) {
                    stepg[rp + j] = 0;
                    atomicOr(&known[(rp + j) >> 5], 1u << ((rp + j) & 31));
                }
                if (lane == sr) skip = k;
                d_run += lane == sr ? k : 0u;
            }
            // ---- step on
            if (need) {
                p += (bstep ? bstep : 1u) + skip;
                if (p >= segend) {
                    ex = p;
                    act = false;
                }
            }
        }
        if (run) { // (the predecessor's exit was read before this round's barrier)
            xout[s] = (uint16_t)(ex - seg0);
            if (round == 0) my_can = ex;
            else if (walked && !met && ex != my_can) my_split = true;
        }
        if (LJB_FUSE_ROUNDS && round == 0) { // this warp's exits of round 0 are written: the warp after it may start its round 1
            __threadfence_block();
            __syncwarp();
            if (lane == 0) wdone[tid >> 5] = epoch;
        } else {
            __syncthreads(); // the exits of this round are what the next round starts from
        }
        if (round == 0) phase(2); // first walks
    }
    phase(1); // later rounds
    if (has) xin[s] = (uint16_t)(my_in - seg0);
    __syncthreads(); // the entries are read by other threads below
    flush_prev(); // the previous block moves to its place in the stream (its predecessors have long published their sizes)

    // ---------------- what the emission passes read: step[] along the chain, entry[] per emission segment ----------------
    // (S is dead: step[] and entry[] take its place.  step[] is copied whole from the per-CTA array the walkers wrote; what it
    // holds at positions no walker visited is left over from earlier blocks and never read: the emission passes only follow the chain.)
    {
        const uint4 *g4 = reinterpret_cast<const uint4 *>(stepg);
        uint4 *z = reinterpret_cast<uint4 *>(step);
        for (uint32_t i = (uint32_t)tid; i < (nb + 15u) / 16u; i += THREADS) z[i] = __ldcg(&g4[i]);
        const uint32_t nseg = (nb + SEG - 1) / SEG;
        uint32_t e8 = 0xFFu;
        if (!wide && (uint32_t)tid < nseg) {
            const uint32_t e = (uint32_t)xin[2 * tid] + 2u * (uint32_t)tid * WG; // first chain position at or behind the segment's start
            if (e < min(((uint32_t)tid + 1u) * SEG, nb)) e8 = e - (uint32_t)tid * SEG;
        }
        entry[tid] = (uint8_t)e8;
        if (wide) { // a wide walker segment is WIDE / 2 emission segments: the entries of all but the first are found along the steps
            __syncthreads(); // step[] and the defaults of entry[] are complete
            if ((uint32_t)tid < nwalk) {
                uint32_t e = (uint32_t)xin[tid] + (uint32_t)tid * wg;
#pragma unroll 1
                for (uint32_t g = (WIDE / 2u) * (uint32_t)tid; g < (WIDE / 2u) * ((uint32_t)tid + 1u) && g < nseg; ++g) {
                    const uint32_t lo = g * SEG, hi = min(lo + (uint32_t)SEG, nb);
                    while (e < lo) e += step[e] ? (uint32_t)step[e] : 1u;
                    if (e < hi) entry[g] = (uint8_t)(e - lo);
                }
            }
        }
    }
    if (P.phase_cycles) {
        const unsigned t_s = __reduce_add_sync(FULL, d_search), t_c = __reduce_add_sync(FULL, d_cand);
        const unsigned t_r = __reduce_add_sync(FULL, d_run);
        if (lane == 0) {
            atomicAdd(&P.phase_cycles[16], (unsigned long long)t_s);
            atomicAdd(&P.phase_cycles[17], (unsigned long long)t_c);
            atomicAdd(&P.phase_cycles[18], (unsigned long long)d_steps);
            atomicAdd(&P.phase_cycles[12], (unsigned long long)d_vl);
            atomicAdd(&P.phase_cycles[13], (unsigned long long)t_r);
        }
    }
    __syncthreads();
    phase(4); // steps and entries
}
