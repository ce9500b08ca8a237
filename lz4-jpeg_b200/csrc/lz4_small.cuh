// lz4-jpeg_b200/csrc/lz4_small.cuh — LZ4 (reference dialect) encoder for SMALL blocks: one WARP per block (included by
// lz4_encode.cu, inside namespace lz4k).
//
// The reference's own block length is 300 bytes (DEFAULT_BLOCK_LENGTH, LZ4.c:23; every size of its harness,
// Experiment/LZ4_sequential_experiment.c:60, uses it).  A 1024-thread CTA with 226 KB of shared memory per 300-byte block — the
// shape of the 64 KiB kernel — ran at 0.5 GB/s.  Here a block belongs to a warp: up to 32 blocks are in flight per SM, each with
// its own few KB of shared memory:
//   stage    the block's bytes
//   index    counting sort of the block's positions by a hash of their 4-gram (256 or 1024 buckets)
//   parse    the greedy chain of block_encode (LZ4.c:516-583), position by position: the 32 lanes compare the bucket's earlier
//            entries with the current position (find_longest_match, LZ4.c:290-323: longest, earliest among the longest, cap
//            1024 and the block end), the match is taken with the reference's (uint8_t) cast (LZ4.c:317), and the sequence is
//            serialised at once (write_sequence, LZ4.c:365-413) into the warp's staging buffer
//   place    decoupled look-back over the blocks' byte counts, then the block is copied to its offset in the stream
#pragma once

constexpr int SMALL_MAXB = 4096; // block lengths up to this take the warp-per-block kernel

struct SmallParams {
    const uint8_t *in;
    size_t n;
    uint32_t block_len;
    uint32_t nblocks;
    uint8_t *out;
    size_t out_cap;
    uint64_t *block_offsets; // nblocks + 1
    uint64_t *result;        // [0] length, [1] phantom, [2] error flags
    uint64_t *status;        // [0] ticket, [1..] look-back words
    uint8_t *staging;        // per warp: one encoded block (worst case)
    size_t stage_stride;
    uint64_t offs_bias;
    uint32_t lead;
    uint32_t frame_byte;
    uint32_t hbits;          // log2 of the number of hash buckets
    uint32_t warp_bytes;     // shared memory per warp
    uint32_t data_bytes;     // of which: the block (padded)
};

__global__ void __launch_bounds__(THREADS, 1) lz4_small_kernel(SmallParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint8_t *const data = smem + (size_t)warp * P.warp_bytes;
    const uint32_t *const dataw = reinterpret_cast<const uint32_t *>(data);
    const uint32_t bl_pad = (P.block_len + 1u) & ~1u;
    uint16_t *const S = reinterpret_cast<uint16_t *>(data + P.data_bytes);
    uint16_t *const dir = S + bl_pad; // [nbk + 2]: bucket starts, then (after the scatter) bucket ends
    const uint32_t nbk = 1u << P.hbits;
    uint8_t *const stage = P.staging + ((size_t)blockIdx.x * nwarps + warp) * P.stage_stride;

    for (;;) {
        long long b = 0;
        if (lane == 0) b = (long long)atomicAdd((unsigned long long *)&P.status[0], 1ull);
        b = __shfl_sync(FULL, b, 0);
        if (b >= (long long)P.nblocks) break;
        const size_t boff = (size_t)b * P.block_len;
        const uint32_t nb = (uint32_t)min((size_t)P.block_len, P.n - boff);
        const uint8_t *src = P.in + boff;
        // ---- stage
        for (uint32_t i = lane; i < nb; i += 32) data[i] = __ldcs(&src[i]);
        for (uint32_t i = nb + lane; i < nb + 32 && i < P.data_bytes; i += 32) data[i] = 0; // defined bytes past the end
        for (uint32_t i = lane; i <= nbk; i += 32) dir[i] = 0;
        __syncwarp();
        // ---- index: histogram, exclusive scan, scatter (the order inside a bucket is whatever the atomics give: the search
        // takes the maximum of (length, -position) over the whole bucket, so it does not matter)
        const uint32_t npos = nb >= 4 ? nb - 3 : 0;
        unsigned short *dir_s = reinterpret_cast<unsigned short *>(dir);
        for (uint32_t p = lane; p < npos; p += 32) {
            const uint32_t h = (load32u(dataw, p) * 2654435761u) >> (32 - P.hbits);
            // u16 counters, two per word: the add goes to the right half (a bucket holds at most 4093 entries: no carry)
            atomicAdd(reinterpret_cast<uint32_t *>(dir_s) + (h >> 1), (h & 1) ? 0x10000u : 1u);
        }
        __syncwarp();
        {
            const uint32_t per = nbk / 32; // counters per lane (8 or 32)
            uint32_t sum = 0;
            for (uint32_t k = 0; k < per; ++k) sum += dir[lane * per + k];
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            uint32_t run = inc - sum;
            for (uint32_t k = 0; k < per; ++k) {
                const uint32_t cnt = dir[lane * per + k];
                dir[lane * per + k] = (uint16_t)run;
                run += cnt;
            }
        }
        __syncwarp();
        for (uint32_t p = lane; p < npos; p += 32) {
            const uint32_t h = (load32u(dataw, p) * 2654435761u) >> (32 - P.hbits);
            const uint32_t old = atomicAdd(reinterpret_cast<uint32_t *>(dir_s) + (h >> 1), (h & 1) ? 0x10000u : 1u);
            S[(h & 1) ? (old >> 16) : (old & 0xFFFFu)] = (uint16_t)p;
        }
        __syncwarp(); // dir[h] is now the END of bucket h

        // ---- parse and emit (warp-uniform control flow)
        uint32_t p = 0, pe = 0;               // current position; end of the last match (start of the pending literals)
        uint32_t out_off = 3, sizes = 3, nseq = 0, phantom = 0;
        auto emit = [&](uint32_t lit, uint32_t ml, uint32_t dist, bool trailing) { // one sequence (LZ4.c:365-413, sizes LZ4.c:540-575)
            uint32_t byte_size, payload;
            if (trailing) {
                byte_size = payload = lit + 5 + lit_ext_count(lit);
            } else {
                const SeqSize z = seq_size(lit, ml);
                byte_size = z.byte_size;
                payload = z.payload;
            }
            uint8_t *dst = stage + out_off;
            const uint32_t next = lit_ext_count(lit);
            if (lane == 0) {
                const uint32_t tok_lit = lit >= 15 ? 15u : lit;
                const uint32_t tok_m = trailing ? 0u : (ml >= 19 ? 15u : ((ml - 4) & 0xFFu));
                dst[0] = (uint8_t)((tok_lit << 4) | tok_m);
                dst[1] = (uint8_t)(byte_size & 0xFF);
                dst[2] = (uint8_t)((byte_size >> 8) & 0xFF);
                if (lit >= 15) {
                    uint32_t rem = (lit - 15) & 0xFF;
                    uint32_t o = 3;
                    if (rem == 255) {
                        dst[o++] = 255;
                        rem = 0;
                    }
                    dst[o] = (uint8_t)rem;
                }
                uint32_t o = 3 + next + lit;
                dst[o++] = (uint8_t)(dist & 0xFF);
                dst[o++] = (uint8_t)(dist >> 8);
                if (!trailing && ml >= 19) dst[o] = (uint8_t)(ml - 19);
            }
            for (uint32_t k = lane; k < lit; k += 32) dst[3 + next + k] = data[pe + k];
            out_off += payload;
            sizes += byte_size;
            phantom += payload != byte_size ? 1u : 0u;
            ++nseq;
        };
        while (p < nb) {
            uint32_t best = 0; // (length << 16) | (0xFFFF - position)
            const uint32_t cap = min((uint32_t)MAX_MATCH, nb - p);
            if (p < npos) {
                const uint32_t g0 = load32u(dataw, p), g1 = load32u(dataw, p + 4);
                const uint32_t h = (g0 * 2654435761u) >> (32 - P.hbits);
                const uint32_t lo = h ? dir[h - 1] : 0u, hi = dir[h];
                for (uint32_t i = lo; i < hi; i += 32) {
                    const uint32_t k = i + (uint32_t)lane;
                    uint32_t key = 0;
                    if (k < hi) {
                        const uint32_t c = S[k];
                        // (a pair that already reaches the cap is only beaten by an earlier position)
                        if (c < p && load32u(dataw, c) == g0 && !((best >> 16) == cap && c > 0xFFFFu - (best & 0xFFFFu)))
                            key = (lcp_from(dataw, c, p, g0, g1, cap) << 16) | (0xFFFFu - c);
                    }
                    best = max(best, __reduce_max_sync(FULL, key));
                }
            }
            const uint32_t len = best >> 16, st = len & 0xFFu; // (uint8_t) cast, LZ4.c:317
            if (len >= 4 && st) {
                emit(p - pe, st, p - (0xFFFFu - (best & 0xFFFFu)), false);
                pe = p + st;
                p += st;
            } else if (len == (uint32_t)MAX_MATCH) {
                // a capped match is a literal step, and so is every following position whose pair with the same candidate still
                // matches 1024 bytes: skip them without searching
                const uint32_t c = 0xFFFFu - (best & 0xFFFFu);
                const uint32_t lim = nb - (uint32_t)MAX_MATCH - p; // positions p + j, j <= lim, still have 1024 bytes before the end
                uint32_t k = lim;
                for (uint32_t j0 = 0; j0 < lim; j0 += 32) {
                    const uint32_t j = j0 + (uint32_t)lane + 1u;
                    const bool mis = j <= lim && data[c + (uint32_t)MAX_MATCH - 1u + j] != data[p + (uint32_t)MAX_MATCH - 1u + j];
                    const unsigned mm = __ballot_sync(FULL, mis);
                    if (mm) {
                        k = j0 + (uint32_t)(__ffs(mm) - 1);
                        break;
                    }
                }
                p += k + 1;
            } else {
                ++p;
            }
        }
        if (nb > pe) emit(nb - pe, 0, 0, true); // trailing literals (LZ4.c:585-613), match_offset 0
        if (lane == 0) {
            stage[0] = (uint8_t)(nseq & 0xFF);          // LZ4.c:615, :417
            stage[1] = (uint8_t)(sizes & 0xFF);         // LZ4.c:617, :419 (low 16 bits)
            stage[2] = (uint8_t)((sizes >> 8) & 0xFF);
            if (phantom) atomicAdd((unsigned long long *)&P.result[1], (unsigned long long)phantom);
            if (b == 0 && P.lead && P.out_cap) P.out[0] = (uint8_t)P.frame_byte; // LZ4.c:429
        }
        __syncwarp(); // the staged bytes were written by different lanes
        // ---- place
        const unsigned long long pay = out_off;
        const unsigned long long base = ljb_lookback(P.status + 1, b, pay, P.lead);
        const bool ok = base + pay <= P.out_cap;
        if (lane == 0) {
            P.block_offsets[b] = base + P.offs_bias;
            if (b == (long long)P.nblocks - 1) {
                P.block_offsets[P.nblocks] = base + pay + P.offs_bias;
                P.result[0] = base + pay;
            }
            if (!ok) atomicOr((unsigned long long *)&P.result[2], 1ull);
        }
        if (ok) {
            uint8_t *dst = P.out + base;
            for (uint32_t k = lane; k < (uint32_t)pay; k += 32) dst[k] = stage[k];
        }
        __syncwarp();
    }
}

// launch geometry of the small-block kernel for a block length (shared by the host wrapper and the emulator build)
struct SmallGeom {
    uint32_t hbits, data_bytes, warp_bytes, warps, smem_bytes;
};
static inline SmallGeom small_geometry(uint32_t block_len)
{
    SmallGeom g;
    g.hbits = block_len <= 1024 ? 8u : 10u;
    g.data_bytes = (block_len + 64u + 15u) & ~15u;
    const uint32_t s_bytes = 2u * ((block_len + 1u) & ~1u);
    const uint32_t dir_bytes = 2u * ((1u << g.hbits) + 2u);
    g.warp_bytes = (g.data_bytes + s_bytes + dir_bytes + 15u) & ~15u;
    uint32_t w = (227u * 1024u) / g.warp_bytes;
    g.warps = w > 32u ? 32u : w;
    g.smem_bytes = g.warps * g.warp_bytes;
    return g;
}
