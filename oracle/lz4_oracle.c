/*
 * oracle/lz4_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's custom "LZ4" block codec
 * (/root/reference/Algorithms/sequential/LZ4/LZ4.c, cited below as S-LZ4:line).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product path (lz4-jpeg_b200/csrc) never
 * links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_lz4.py checks this restatement against
 *   (1) the reference's own golden vector  Output-Input/input/input.txt -> out/compressed.bin
 *       (committed as tests/golden/lz4_input.txt / lz4_compressed.bin), and
 *   (2) the reference's own block_encode()/write_output() compiled from /root/reference into
 *       oracle/_ref/libref_lz4.so (bounded variant, see oracle/build.py) on seeded inputs.
 *
 * One semantic repair relative to the verbatim reference: the match extension is
 * bounded at the block end.  The verbatim reference reads past its exact-size malloc (S-LZ4:301-302,
 * S-LZ4:156), which is undefined behaviour whose result depends on heap residue.
 *
 * Everything else — including every 8/16-bit truncation — follows the reference bit for bit.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define ORC_MAX_MATCH 1024u /* S-LZ4:20 MAX_MATCH_LENGTH */
#define ORC_MIN_MATCH 4u    /* S-LZ4:21 MIN_MATCH_LENGTH */
#define ORC_WINDOW 65535u   /* S-LZ4:22 WINDOW_SIZE      */

/* ---- S-LZ4:290-323 find_longest_match (bounded variant) -----------------------------------
 * Exhaustive scan i = window_start .. cur-1, ascending; strict '>' keeps the EARLIEST position among
 * equal lengths (largest distance).  Extension capped at 1024 and (repair) at the block end.
 * Returns the TRUE longest length (0..1024) — the caller applies the reference's (uint8_t) cast.  */
static size_t orc_longest_match(const uint8_t *in, size_t n, size_t cur, size_t *dist)
{
    size_t best = 0, best_dist = 0;
    size_t start = (cur >= ORC_WINDOW) ? cur - ORC_WINDOW : 0;
    for (size_t i = start; i < cur; ++i) {
        size_t len = 0;
        while (len < ORC_MAX_MATCH && cur + len < n && in[i + len] == in[cur + len])
            ++len;
        if (len > best) {
            best = len;
            best_dist = cur - i;
        }
    }
    *dist = best_dist;
    return best;
}

/* Same contract as orc_longest_match but O(candidates) instead of O(n): a match of length >= 4 must
 * share its first four bytes with `cur`, so it suffices to visit earlier positions with the same
 * 4-gram, in ascending order (chain built by orc_build_chains).  Used only to make parity tests on
 * multi-megabyte inputs finish in seconds; tests/test_oracle_lz4.py proves it equal to the exhaustive
 * scan above on every small case. */
typedef struct {
    uint32_t *first; /* 65536-entry hash -> first position with that hash (or NONE) */
    uint32_t *next;  /* position -> next later position with same hash (or NONE)   */
    uint32_t *last;
} orc_chains;
#define ORC_NONE 0xFFFFFFFFu
static inline uint32_t orc_hash4(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return (v * 2654435761u) >> 16;
}
static void orc_build_chains(const uint8_t *in, size_t n, orc_chains *c)
{
    for (size_t h = 0; h < 65536; ++h) c->first[h] = c->last[h] = ORC_NONE;
    if (n < 4) return;
    for (size_t p = 0; p + 4 <= n; ++p) {
        uint32_t h = orc_hash4(in + p);
        c->next[p] = ORC_NONE;
        if (c->last[h] == ORC_NONE) c->first[h] = (uint32_t)p;
        else c->next[c->last[h]] = (uint32_t)p;
        c->last[h] = (uint32_t)p;
    }
}
static size_t orc_longest_match_fast(const uint8_t *in, size_t n, size_t cur, size_t *dist, const orc_chains *c)
{
    size_t best = 0, best_dist = 0;
    *dist = 0;
    if (cur + 4 > n) { /* fewer than 4 bytes left: any match is < 4 and is discarded by the caller */
        return 0;
    }
    size_t cap = n - cur;
    if (cap > ORC_MAX_MATCH) cap = ORC_MAX_MATCH;
    for (uint32_t i = c->first[orc_hash4(in + cur)]; i != ORC_NONE && i < cur; i = c->next[i]) {
        if (memcmp(in + i, in + cur, 4) != 0) continue;
        size_t len = 4;
        while (len < cap && in[i + len] == in[cur + len]) ++len;
        if (len > best) {
            best = len;
            best_dist = cur - i;
            if (best == cap) break; /* strict '>' can never fire again */
        }
    }
    *dist = best_dist;
    return best; /* 0 or >= 4 */
}

/* ---- literal / match extension byte counts in the reference's uint8_t arithmetic -------------
 * S-LZ4:548-560 (sizing) and S-LZ4:372-386 (writing) use `uint8_t remaining = literals_count - 15`,
 * so the count wraps mod 256: one byte, or two (255,0) when the wrapped value is exactly 255.     */
static inline unsigned orc_lit_ext_count(size_t lit)
{
    if (lit < 15) return 0;
    return (((lit - 15) & 0xFF) == 255) ? 2u : 1u;
}

typedef struct {
    size_t payload;     /* bytes actually serialised for this block (header + sequences)        */
    size_t header_size; /* 3 + sum(seq.byte_size) as the reference computes it (S-LZ4:617)      */
    size_t nseq;        /* true sequence count (header stores the low byte, S-LZ4:615)          */
    size_t phantom;     /* sequences whose byte_size counts a match-ext byte that is not written */
} orc_block_info;

/* ---- S-LZ4:506-620 block_encode + S-LZ4:365-425 write_sequence/write_block --------------------
 * Emits the serialised block into out (if out != NULL) and returns its description.            */
static int orc_block_encode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, orc_block_info *info,
                            const orc_chains *chains)
{
    size_t o = 3; /* block header is patched at the end */
    size_t sum_sizes = 0, nseq = 0, phantom = 0;
    size_t pos = 0;
    uint16_t literal_counter = 0; /* S-LZ4:514: uint16_t, wraps at 65536 */
    size_t lit_start = 0;         /* S-LZ4:523-526: pointer captured when the counter is 0 */
    if (out && cap < 3) return -1;

    while (pos < n) {
        size_t dist;
        size_t best = chains ? orc_longest_match_fast(in, n, pos, &dist, chains)
                             : orc_longest_match(in, n, pos, &dist);
        uint8_t ml = (best >= ORC_MIN_MATCH) ? (uint8_t)best : 0; /* S-LZ4:314-317 */
        if (ml == 0) {
            if (literal_counter == 0) lit_start = pos;
            ++pos;
            ++literal_counter;
            continue;
        }
        /* match sequence, S-LZ4:534-581 */
        size_t lit = literal_counter;
        uint8_t tok_lit = (lit >= 15) ? 15 : (uint8_t)lit;
        uint8_t tok_m = (ml >= 19) ? 15 : (uint8_t)(ml - ORC_MIN_MATCH); /* wraps to 253..255 for ml 1..3 */
        uint8_t token = (uint8_t)((tok_lit << 4) | tok_m);
        size_t byte_size = lit + 5 + orc_lit_ext_count(lit);
        uint8_t adj = (uint8_t)(ml - 4);
        if (adj >= 15) byte_size += 1; /* remaining = adj-15 <= 240 < 255: exactly one byte, S-LZ4:564-575 */
        int writes_mext = (ml >= 4) && (adj >= 15); /* S-LZ4:393-411 */
        if (adj >= 15 && !writes_mext) ++phantom;
        size_t need = byte_size - ((adj >= 15 && !writes_mext) ? 1 : 0);
        if (out) {
            if (o + need > cap) return -1;
            out[o++] = token;
            out[o++] = (uint8_t)(byte_size & 0xFF);
            out[o++] = (uint8_t)((byte_size >> 8) & 0xFF);
            if (lit >= 15) {
                uint8_t rem = (uint8_t)(lit - 15);
                if (rem == 255) { out[o++] = 255; rem = 0; }
                out[o++] = rem;
            }
            memcpy(out + o, in + lit_start, lit);
            o += lit;
            out[o++] = (uint8_t)(dist & 0xFF);
            out[o++] = (uint8_t)((dist >> 8) & 0xFF);
            if (writes_mext) out[o++] = (uint8_t)(adj - 15);
        } else {
            o += need;
        }
        sum_sizes += byte_size;
        ++nseq;
        literal_counter = 0;
        pos += ml;
    }
    if (literal_counter > 0) { /* trailing literals, S-LZ4:585-613 */
        size_t lit = literal_counter;
        uint8_t tok_lit = (lit >= 15) ? 15 : (uint8_t)lit;
        size_t byte_size = lit + 5 + orc_lit_ext_count(lit);
        if (out) {
            if (o + byte_size > cap) return -1;
            out[o++] = (uint8_t)(tok_lit << 4);
            out[o++] = (uint8_t)(byte_size & 0xFF);
            out[o++] = (uint8_t)((byte_size >> 8) & 0xFF);
            if (lit >= 15) {
                uint8_t rem = (uint8_t)(lit - 15);
                if (rem == 255) { out[o++] = 255; rem = 0; }
                out[o++] = rem;
            }
            memcpy(out + o, in + lit_start, lit);
            o += lit;
            out[o++] = 0;
            out[o++] = 0;
        } else {
            o += byte_size;
        }
        sum_sizes += byte_size;
        ++nseq;
    }
    size_t header_size = sum_sizes + 3; /* S-LZ4:617 */
    if (out) {
        out[0] = (uint8_t)(nseq & 0xFF);            /* S-LZ4:615, 417 */
        out[1] = (uint8_t)(header_size & 0xFF);     /* S-LZ4:419: low 16 bits, little endian */
        out[2] = (uint8_t)((header_size >> 8) & 0xFF);
    }
    info->payload = o;
    info->header_size = header_size;
    info->nseq = nseq;
    info->phantom = phantom;
    return 0;
}

/* ---- S-LZ4:670-742 lz4_encode (block loop) + S-LZ4:427-441 write_output ------------------------
 * in[0..n) is cut into ceil(n/block_len) blocks (S-LZ4:123-177), each encoded independently.
 * out = u8 nblocks_lo8 | block*      block_offsets[b] = byte offset of block b in out; [nblocks] = end.
 * mode: 0 = exhaustive search (the reference algorithm), 1 = 4-gram chains (same result, faster).
 * Returns 0, or -1 when out_cap is too small / arguments are invalid.                             */
int oracle_lz4_compress(const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                        uint64_t *block_offsets, size_t *out_len, uint64_t *phantom_total, int mode)
{
    if (block_len == 0 || block_len > 65536 || n == 0) return -1;
    size_t nblocks = (n + block_len - 1) / block_len;
    if (out_cap < 1) return -1;
    out[0] = (uint8_t)(nblocks & 0xFF); /* S-LZ4:429 */
    size_t o = 1;
    uint64_t ph = 0;
    orc_chains ch = {0};
    if (mode == 1) {
        ch.first = malloc(65536 * sizeof(uint32_t));
        ch.last = malloc(65536 * sizeof(uint32_t));
        ch.next = malloc((block_len + 1) * sizeof(uint32_t));
        if (!ch.first || !ch.last || !ch.next) return -1;
    }
    int rc = 0;
    for (size_t b = 0; b < nblocks; ++b) {
        const uint8_t *blk = in + b * block_len;
        size_t len = (b == nblocks - 1) ? n - b * block_len : block_len;
        orc_block_info info;
        if (mode == 1) orc_build_chains(blk, len, &ch);
        if (block_offsets) block_offsets[b] = o;
        if (orc_block_encode(blk, len, out + o, out_cap - o, &info, mode == 1 ? &ch : NULL) != 0) {
            rc = -1;
            break;
        }
        o += info.payload;
        ph += info.phantom;
    }
    if (block_offsets && rc == 0) block_offsets[nblocks] = o;
    if (mode == 1) {
        free(ch.first);
        free(ch.last);
        free(ch.next);
    }
    if (out_len) *out_len = o;
    if (phantom_total) *phantom_total = ph;
    return rc;
}

/* Per-position longest match (true length 0..1024 and earliest position), for kernel-level parity
 * tests of the match-search stage on one block.  best_len[p] < 4 is reported as 0.               */
int oracle_lz4_matches(const uint8_t *in, size_t n, uint16_t *best_len, uint16_t *best_dist, int mode)
{
    if (n == 0 || n > 65536) return -1;
    orc_chains ch = {0};
    if (mode == 1) {
        ch.first = malloc(65536 * sizeof(uint32_t));
        ch.last = malloc(65536 * sizeof(uint32_t));
        ch.next = malloc((n + 1) * sizeof(uint32_t));
        if (!ch.first || !ch.last || !ch.next) return -1;
        orc_build_chains(in, n, &ch);
    }
    for (size_t p = 0; p < n; ++p) {
        size_t dist;
        size_t best = mode == 1 ? orc_longest_match_fast(in, n, p, &dist, &ch) : orc_longest_match(in, n, p, &dist);
        if (best < ORC_MIN_MATCH) { best = 0; dist = 0; }
        best_len[p] = (uint16_t)best;
        best_dist[p] = (uint16_t)dist;
    }
    if (mode == 1) {
        free(ch.first);
        free(ch.last);
        free(ch.next);
    }
    return 0;
}

/* ---- Format-level decoder -------------------------------------------------------
 * The reference decoder (S-LZ4:744-1121) is count-driven (u8 counts wrap) and uses signed char
 * arithmetic, so it is only valid for tiny inputs.  This decoder walks the stream structurally using
 * the out-of-band block_offsets table (true block extents), recovers literal counts >= 271 from
 * seq_byte_size (A.3-c), and copies matches byte-wise with overlap (S-LZ4:956-977).
 * It cannot decode the reference's ambiguous 257..259-length "phantom" sequences (A.3-b): those are
 * reported by the encoder (phantom_total) and make this function return -2 when detected.          */
int oracle_lz4_decompress(const uint8_t *comp, size_t clen, const uint64_t *block_offsets, size_t nblocks,
                          size_t block_len, uint8_t *out, size_t out_cap, size_t *out_len)
{
    size_t o = 0;
    for (size_t b = 0; b < nblocks; ++b) {
        size_t s = (size_t)block_offsets[b], e = (size_t)block_offsets[b + 1];
        size_t blk_out0 = o;
        if (s > e || e > clen) return -2; /* offsets must be monotonic and inside the stream */
        if (e < s + 3) return -1;
        s += 3;
        while (s < e) {
            if (s + 3 > e) return -1;
            uint8_t token = comp[s];
            size_t size16 = comp[s + 1] | ((size_t)comp[s + 2] << 8);
            size_t q = s + 3;
            size_t lit = token >> 4;
            unsigned mtok = token & 15;
            if (lit == 15) {
                /* one ext byte, or (255,0); true literal count is recovered from the size field */
                unsigned next = 1;
                if (q >= e) return -1;
                if (comp[q] == 255) next = 2;
                /* size = lit + 5 + next (+1 if a match-ext byte follows the offset) */
                size_t fixed = 5 + next + (mtok == 15 ? 1 : 0);
                /* the size field is the low 16 bits of byte_size (S-LZ4:369); a sequence of >= 65531
                 * literals wraps it.  Unwrapped, a sequence with >= 15 literals has size >= fixed + 15;
                 * wrapped (true size <= 65536 + 8) the low 16 bits are <= 8: the cases exclude each other. */
                if (size16 < fixed + 15) size16 += 65536;
                lit = size16 - fixed;
                if (((lit - 15) & 0xFF) != (next == 2 ? 255u : comp[q])) return -2;
                q += next;
            }
            if (q + lit + 2 > e) return -2;
            if (o + lit > out_cap) return -1;
            memcpy(out + o, comp + q, lit);
            o += lit;
            q += lit;
            size_t off = comp[q] | ((size_t)comp[q + 1] << 8);
            q += 2;
            if (off != 0) {
                size_t mlen = mtok + 4;
                if (mtok == 15) {
                    if (q >= e) return -2;
                    mlen = (size_t)comp[q++] + 19;
                }
                if (off > o - blk_out0) return -2;
                if (o + mlen > out_cap) return -1;
                for (size_t k = 0; k < mlen; ++k, ++o) out[o] = out[o - off];
            } else if (mtok != 0) {
                return -2;
            }
            s = q;
        }
        if (o - blk_out0 > block_len) return -2;
        if (b + 1 < nblocks && o - blk_out0 != block_len) return -2; /* only the last block may be short */
    }
    if (out_len) *out_len = o;
    return 0;
}

/* ---- Synthetic text, random_extract-style (Experiment/random_extract.c:8-71) --------------------
 * Repeatedly copy a `passage`-byte slice of the corpus starting at rng() % (corpus_len - passage)
 * (X-rext:36) and turn CR/LF into spaces (X-rext:49-53).  The reference seeds rand() with time();
 * here a splitmix64 stream with an explicit seed makes CPU and GPU runs see identical bytes.
 * (Duplicated on purpose in lz4-jpeg_b200/csrc/synth.c: the product must not link the oracle.)    */
static inline uint64_t orc_splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void oracle_synth_text(const uint8_t *corpus, size_t corpus_len, uint64_t seed, size_t passage, uint8_t *out,
                       size_t n)
{
    uint64_t s = seed;
    size_t o = 0;
    while (o < n) {
        size_t start = (size_t)(orc_splitmix64(&s) % (uint64_t)(corpus_len - passage));
        size_t take = (n - o < passage) ? n - o : passage;
        for (size_t k = 0; k < take; ++k) {
            uint8_t c = corpus[start + k];
            out[o + k] = (c == '\n' || c == '\r') ? ' ' : c;
        }
        o += take;
    }
}
