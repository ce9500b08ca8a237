// lz4-jpeg_b200/csrc/lz4_decode.cu — format-level LZ4 (reference dialect) block decoder for sm_100a.
//
// Replaces the compute of LZ4_decode() / block_decode() / sequence_decode() / interpret_sequence()
// (Algorithms/sequential/LZ4/LZ4.c:1038, :845, :744, :937; "parallel" form parallel_LZ4_decode,
// Algorithms/parallel/LZ4/LZ4.c:1105).  The reference decoder is count-driven (its u8 counters wrap
// beyond 255 sequences / 127 blocks) and does signed-char arithmetic, so it cannot decode what its own
// encoder writes at 64 KiB blocks.  This decoder follows the wire format instead:
// blocks are delimited by the out-of-band offset table, sequences are walked structurally, literal
// counts >= 271 are recovered from the sequence size field, matches are copied with overlap semantics
// (LZ4.c:956-977).  One warp decodes one block; blocks are independent (offsets never reach before the
// block start), which is the reference's own thread-per-block decomposition.
#include "common.cuh"

namespace lz4d {

// Eight lanes decode one block: four blocks per warp, 32 per CTA, up to 512 in flight per SM.  The walk along a block's sequences
// is a chain of dependent loads (token and size, then the literals' end, then the match source, which is output this same team has
// just written): what limits the kernel is how many such chains are in flight, not how wide the copies are (a sequence of the
// benchmark text is 7 bytes).  One warp per block ran at 62 GB/s.
constexpr int TEAM = 8;
constexpr int WARPS_PER_CTA = 8;
constexpr int BLOCKS_PER_CTA = WARPS_PER_CTA * (32 / TEAM);

struct Params {
    const uint8_t *comp;
    const uint64_t *offs; // nblocks + 1
    size_t comp_len;
    uint32_t nblocks;        // blocks of this launch
    uint32_t first_block;    // their place in the stream: block b of the launch is block first_block + b
    uint32_t total_blocks;   // blocks of the whole stream (only the last of them may be short)
    uint32_t block_len;
    uint8_t *out;
    size_t out_cap;
    uint32_t *block_out_len; // per block decoded length
    uint64_t *result;        // [0] decoded bytes, [2] error flags: bit0 capacity, bit1 format
};

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) lz4_decode_kernel(Params P)
{
    const uint32_t lane = threadIdx.x & 31, tl = lane & (TEAM - 1);
    const uint32_t b = (blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * (32 / TEAM) + lane / TEAM;
    const unsigned tmask = ((1u << TEAM) - 1u) << (lane & ~(uint32_t)(TEAM - 1)); // this team's lanes (teams take different paths)
    if (b >= P.nblocks) return;
    const uint8_t *c = P.comp;
    size_t s = (size_t)P.offs[b], e = (size_t)P.offs[b + 1];
    if (s > e || e > P.comp_len) { // a table that is not monotonic or points past the stream: nothing is read through it
        if (tl == 0) {
            P.block_out_len[b] = 0;
            atomicOr((unsigned long long *)&P.result[2], 2ull);
        }
        return;
    }
    const size_t gb = (size_t)P.first_block + b;
    const size_t out0 = gb * P.block_len;
    const size_t out_lim = min(P.out_cap, out0 + (size_t)P.block_len);
    uint8_t *out = P.out;
    size_t o = out0;
    unsigned err = 0;
    if (e < s + 3) err = 2;
    s += 3; // block header: u8 nseq_lo8, u16 size (both wrap; the offset table is authoritative)
    while (!err && s < e) {
        if (s + 3 > e) { err = 2; break; }
        const uint32_t token = c[s];
        size_t size16 = (size_t)c[s + 1] | ((size_t)c[s + 2] << 8);
        size_t q = s + 3;
        size_t lit = token >> 4;
        const uint32_t mtok = token & 15;
        if (lit == 15) {
            if (q >= e) { err = 2; break; }
            const uint32_t e1 = c[q];
            const uint32_t next = e1 == 255 ? 2 : 1;
            const size_t fixed = 5 + next + (mtok == 15 ? 1 : 0);
            // u16 wrap of the size field (>= 65531 literals): an unwrapped field of a sequence with >= 15 literals is at least
            // fixed + 15, a wrapped one (true size <= 65536 + 8) has low 16 bits <= 8 — the two cases exclude each other
            if (size16 < fixed + 15) size16 += 65536;
            lit = size16 - fixed;
            if (((lit - 15) & 0xFF) != (next == 2 ? 255u : e1)) { err = 2; break; }
            q += next;
        }
        if (q + lit + 2 > e) { err = 2; break; }
        if (o + lit > out_lim) { err = (o + lit > P.out_cap) ? 1 : 2; break; }
        for (size_t k = tl; k < lit; k += TEAM) out[o + k] = c[q + k];
        o += lit;
        q += lit;
        const size_t off = (size_t)c[q] | ((size_t)c[q + 1] << 8);
        q += 2;
        if (off != 0) {
            size_t mlen = mtok + 4;
            if (mtok == 15) {
                if (q >= e) { err = 2; break; }
                mlen = (size_t)c[q++] + 19;
            }
            if (off > o - out0) { err = 2; break; }
            if (o + mlen > out_lim) { err = (o + mlen > P.out_cap) ? 1 : 2; break; }
            __syncwarp(tmask); // literals written by other lanes of the team may be match source
            // overlapping forward copy == periodic extension of the last `off` bytes
            if (off >= mlen) {
                for (size_t k = tl; k < mlen; k += TEAM) out[o + k] = out[o - off + k];
            } else {
                for (size_t k = tl; k < mlen; k += TEAM) out[o + k] = out[o - off + (k % off)];
            }
            o += mlen;
            __syncwarp(tmask);
        } else if (mtok != 0) {
            err = 2;
            break;
        }
        s = q;
    }
    if (!err && gb + 1 < P.total_blocks && o - out0 != P.block_len) err = 2; // only the last block may be short: no holes in the output
    if (tl == 0) {
        P.block_out_len[b] = (uint32_t)(o - out0);
        if (gb + 1 == P.total_blocks) P.result[0] = (uint64_t)(o); // decoded bytes (every earlier block is block_len long, checked above)
        if (err) atomicOr((unsigned long long *)&P.result[2], (unsigned long long)err);
    }
}

} // namespace lz4d

#ifndef LJB_EMU_BUILD
// One launch over blocks [first_block, first_block + nblocks) of a stream of total_blocks.  d_comp / d_out are addressed with the
// stream's own offsets (block_offsets values; gb * block_len), so a caller holding only a chunk passes biased pointers.
static int decode_launch(ljb_ctx *ctx, const uint8_t *d_comp, size_t comp_len, const uint64_t *d_offs, size_t nblocks, size_t first_block,
                         size_t total_blocks, size_t block_len, uint8_t *d_out, size_t out_cap, uint32_t *d_block_out_len, uint64_t *d_result)
{
    using namespace lz4d;
    Params P;
    P.comp = d_comp;
    P.offs = d_offs;
    P.comp_len = comp_len;
    P.nblocks = (uint32_t)nblocks;
    P.first_block = (uint32_t)first_block;
    P.total_blocks = (uint32_t)total_blocks;
    P.block_len = (uint32_t)block_len;
    P.out = d_out;
    P.out_cap = out_cap;
    P.block_out_len = d_block_out_len;
    P.result = d_result;
    const unsigned grid = (unsigned)((nblocks + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA);
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    lz4_decode_kernel<<<grid, WARPS_PER_CTA * 32, 0, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

// Device-resident form: compressed stream, offset table and output stay in HBM; asynchronous on the context's stream.
// d_result[3] (u64): [0] decoded bytes, [2] error flags (bit 0 capacity, bit 1 format).  d_block_out_len: nblocks u32.
extern "C" int ljb_lz4_decompress_dev(ljb_ctx *ctx, const uint8_t *d_comp, size_t comp_len, const uint64_t *d_block_offsets,
                                      size_t nblocks, size_t block_len, uint8_t *d_out, size_t out_cap, uint32_t *d_block_out_len,
                                      uint64_t *d_result)
{
    if (!ctx || !d_comp || !d_block_offsets || !d_out || !d_block_out_len || !d_result || nblocks == 0 || nblocks > 0x7fffffffull ||
        block_len == 0 || block_len > 65536)
        return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    ctx->kernel_ms_summed = 0;
    return decode_launch(ctx, d_comp, comp_len, d_block_offsets, nblocks, 0, nblocks, block_len, d_out, out_cap, d_block_out_len, d_result);
}

// Host-buffer entry point: like ljb_lz4_compress, the stream goes through the device in chunks of whole blocks — upload of chunk
// k+1, kernel of chunk k and download of chunk k-1 run concurrently (three streams, two buffers each way).
extern "C" int ljb_lz4_decompress(ljb_ctx *ctx, const uint8_t *comp, size_t comp_len, const uint64_t *block_offsets,
                                  size_t nblocks, size_t block_len, uint8_t *out, size_t out_cap, size_t *out_len)
{
    if (!ctx || !comp || !block_offsets || !out || nblocks == 0 || nblocks > 0x7fffffffull || block_len == 0 || block_len > 65536)
        return LJB_E_ARG;
    if (block_offsets[nblocks] > comp_len) return LJB_E_ARG;
    for (size_t i = 0; i < nblocks; ++i) // the chunks are cut along the table: it has to be monotonic (the kernel checks the rest)
        if (block_offsets[i] > block_offsets[i + 1]) return LJB_E_FORMAT;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    size_t cblocks = 2 * ljb_pipe_chunk() / block_len; // 256 MiB of output per chunk
    if (cblocks == 0) cblocks = 1;
    if (nblocks < 4 * cblocks) { // a small stream: about eight chunks, so that uploads, kernels and downloads still overlap
        const size_t c = ((nblocks + 7) / 8 + 255) / 256 * 256;
        if (c < cblocks) cblocks = c;
    }
    if (cblocks > nblocks) cblocks = nblocks;
    const size_t nchunks = (nblocks + cblocks - 1) / cblocks;
    size_t max_comp = 0;
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t b0 = k * cblocks, b1 = b0 + cblocks < nblocks ? b0 + cblocks : nblocks;
        const size_t cb = (size_t)(block_offsets[b1] - block_offsets[b0]);
        if (cb > max_comp) max_comp = cb;
    }
    const size_t need_out = nblocks * block_len;
    const size_t dcap = need_out < out_cap ? need_out : out_cap; // decoded bytes the caller can take
    if ((rc = ljb_pipe_init(ctx, nchunks)) != 0) return rc;
    for (int i = 0; i < (nchunks > 1 ? 2 : 1); ++i) {
        if ((rc = ljb_ensure(&ctx->d_pin[i], &ctx->pin_bytes[i], max_comp + 64)) != 0) return rc;
        if ((rc = ljb_ensure(&ctx->d_pout[i], &ctx->pout_bytes[i], cblocks * block_len + 64)) != 0) return rc;
    }
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, (nblocks + 1 + 3) * sizeof(uint64_t) + nblocks * sizeof(uint32_t))) != 0)
        return rc;
    uint64_t *d_offs = (uint64_t *)ctx->d_small;
    uint64_t *d_res = d_offs + nblocks + 1;
    uint32_t *d_len = (uint32_t *)(d_res + 3);
    int status = LJB_OK;
    float kernel_ms = 0.f;
    uint64_t res[3] = {0, 0, 0};
    cudaError_t e;
#define PIPE(x)                                                                 \
    do {                                                                        \
        e = (x);                                                                \
        if (e != cudaSuccess) {                                                 \
            status = ljb_set_cuda_error(e, #x, __LINE__);                       \
            goto done;                                                          \
        }                                                                       \
    } while (0)
    PIPE(cudaMemcpyAsync(d_offs, block_offsets, (nblocks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    PIPE(cudaMemsetAsync(d_res, 0, 3 * sizeof(uint64_t), ctx->stream));
    PIPE(cudaMemcpyAsync(ctx->d_pin[0], comp + block_offsets[0], (size_t)(block_offsets[cblocks < nblocks ? cblocks : nblocks] - block_offsets[0]),
                         cudaMemcpyHostToDevice, ctx->s_in));
    PIPE(cudaEventRecord(ctx->ev_h2d[0], ctx->s_in));
    for (size_t k = 0; k < nchunks; ++k) {
        const int b = (int)(k & 1);
        const size_t b0 = k * cblocks, b1 = b0 + cblocks < nblocks ? b0 + cblocks : nblocks;
        PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        if (k >= 2) PIPE(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0)); // output buffer b is free again
        // the chunk's bytes are addressed with the stream's own offsets: bias the buffers' pointers (never dereferenced outside the chunk)
        const uint8_t *c_biased = (const uint8_t *)ctx->d_pin[b] - block_offsets[b0];
        uint8_t *o_biased = (uint8_t *)ctx->d_pout[b] - b0 * block_len;
        const size_t cap_k = b1 * block_len < dcap ? b1 * block_len : dcap;
        rc = decode_launch(ctx, c_biased, (size_t)block_offsets[b1], d_offs + b0, b1 - b0, b0, nblocks, block_len, o_biased, cap_k, d_len + b0, d_res);
        if (rc != 0) {
            status = rc;
            goto done;
        }
        PIPE(cudaEventRecord(ctx->ev_kern[b], ctx->stream));
        if (k + 1 < nchunks) { // next chunk's upload overlaps this chunk's kernel; its buffer was read by kernel k-1
            const size_t n0 = b1, n1 = n0 + cblocks < nblocks ? n0 + cblocks : nblocks;
            if (k >= 1) PIPE(cudaStreamWaitEvent(ctx->s_in, ctx->ev_kern[b ^ 1], 0));
            PIPE(cudaMemcpyAsync(ctx->d_pin[b ^ 1], comp + block_offsets[n0], (size_t)(block_offsets[n1] - block_offsets[n0]),
                                 cudaMemcpyHostToDevice, ctx->s_in));
            PIPE(cudaEventRecord(ctx->ev_h2d[b ^ 1], ctx->s_in));
        }
        PIPE(cudaStreamWaitEvent(ctx->s_out, ctx->ev_kern[b], 0));
        if (k + 1 < nchunks) { // every block of a chunk that is not the last decodes to block_len bytes (or the call fails)
            if (b1 * block_len <= dcap)
                PIPE(cudaMemcpyAsync(out + b0 * block_len, ctx->d_pout[b], (b1 - b0) * block_len, cudaMemcpyDeviceToHost, ctx->s_out));
        } else { // the last chunk: its length is known once its kernel has run
            PIPE(cudaEventSynchronize(ctx->ev_kern[b]));
            PIPE(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, ctx->s_out));
            PIPE(cudaStreamSynchronize(ctx->s_out));
            if (!(res[2] & 3) && res[0] >= b0 * block_len && res[0] <= dcap)
                PIPE(cudaMemcpyAsync(out + b0 * block_len, ctx->d_pout[b], (size_t)res[0] - b0 * block_len, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        PIPE(cudaEventRecord(ctx->ev_d2h[b], ctx->s_out));
        {
            PIPE(cudaEventSynchronize(ctx->ev_kern[b])); // (ev0 / ev1 are re-recorded by the next launch)
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) kernel_ms += ms;
        }
    }
    PIPE(cudaStreamSynchronize(ctx->s_out));
done:
#undef PIPE
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    ctx->last_kernel_ms = kernel_ms;
    ctx->kernel_ms_summed = 1;
    if (status != LJB_OK) return status;
    if (res[2] & 2) return LJB_E_FORMAT;
    if (res[2] & 1) return LJB_E_CAPACITY;
    if (out_len) *out_len = (size_t)res[0];
    if (res[0] > out_cap) return LJB_E_CAPACITY;
    return LJB_OK;
}
#endif // !LJB_EMU_BUILD
