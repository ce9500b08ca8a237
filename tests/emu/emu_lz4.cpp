// tests/emu/emu_lz4.cpp — TEST INFRASTRUCTURE: the LZ4 encoder kernel (lz4-jpeg_b200/csrc/lz4_encode.cu, both search modes)
// compiled for the host under the lock-step emulator and driven with host buffers.  Used by tests/test_emu_lz4.py to check
// the kernel's logic against the oracle without a GPU.  Not a product path.
#define LJB_EMU_BUILD 1
#include "../../lz4-jpeg_b200/csrc/lz4_encode.cu"
#include "../../lz4-jpeg_b200/csrc/lz4_decode.cu"

namespace lz4k {
constexpr int EMU_SMEM = 227 * 1024;      // what a CTA can have on the B200
alignas(16) uint8_t smem[EMU_SMEM + 256]; // the kernels' `extern __shared__` array; 256 canary bytes behind it
}

extern "C" int emu_lz4_compress(const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap, uint64_t *block_offsets,
                                uint64_t *result, int mode, uint16_t *dump_len, uint16_t *dump_dist, unsigned long long *phase24)
{
    using namespace lz4k;
    if (!in || !out || n == 0 || block_len == 0 || block_len > (size_t)MAXB) return -1;
    const size_t nblocks = (n + block_len - 1) / block_len;
    std::vector<uint8_t> din(n + 64, 0);
    memcpy(din.data(), in, n);
    const size_t stage_stride = (1 + 3 + 6 * block_len + 64 + 16 + 255) & ~(size_t)255;
    std::vector<uint64_t> status(nblocks + 2 + 24, 0);
    std::vector<uint32_t> scratch(MAXB, 0xDEADBEEFu); // (stale records of "the previous block" must never be read)
    std::vector<uint16_t> gids(MAXB, 0xBEEF);
    std::vector<uint8_t> staging(2 * stage_stride + 64, 0xEE);
    Params P;
    memset(&P, 0, sizeof P);
    P.in = din.data();
    P.n = n;
    P.block_len = (uint32_t)block_len;
    P.nblocks = (uint32_t)nblocks;
    P.out = out;
    P.out_cap = out_cap;
    P.block_offsets = block_offsets;
    P.result = result;
    P.status = status.data();
    P.scratch = scratch.data();
    P.gids = gids.data();
    std::vector<uint32_t> idx8(NBUCKET + MAXB / 2, 0xDEADBEEFu);
    P.idx8 = idx8.data();
    P.staging = staging.data();
    P.stage_stride = stage_stride;
    P.offs_bias = 0;
    P.lead = 1;
    P.frame_byte = (uint32_t)(nblocks & 0xFF);
    P.dump_len = dump_len;
    P.dump_dist = dump_dist;
    P.tune = getenv("LJB_LZ4_TUNE") ? (uint32_t)atoi(getenv("LJB_LZ4_TUNE")) : 0u; // the kernel's experiment switches, as the library reads them
    P.phase_cycles = phase24; // optional: 24 counters (the kernel's LJB_LZ4_PHASES probes; 'cycles' are emulator ticks)
    if (phase24) memset(phase24, 0, 24 * sizeof *phase24);
    result[0] = result[1] = result[2] = 0;
    memset(smem, 0xA5, sizeof smem);
    if (mode == 2) { // the warp-per-block kernel for small blocks
        if (block_len > (size_t)SMALL_MAXB) return -1;
        const SmallGeom g = small_geometry((uint32_t)block_len);
        const size_t stride = (1 + 3 + 6 * block_len + 64 + 63) & ~(size_t)63;
        std::vector<uint8_t> sstage((size_t)g.warps * stride + 64, 0xEE);
        SmallParams S;
        memset(&S, 0, sizeof S);
        S.in = din.data();
        S.n = n;
        S.block_len = (uint32_t)block_len;
        S.nblocks = (uint32_t)nblocks;
        S.out = out;
        S.out_cap = out_cap;
        S.block_offsets = block_offsets;
        S.result = result;
        S.status = status.data();
        S.staging = sstage.data();
        S.stage_stride = stride;
        S.lead = 1;
        S.frame_byte = (uint32_t)(nblocks & 0xFF);
        S.hbits = g.hbits;
        S.warp_bytes = g.warp_bytes;
        S.data_bytes = g.data_bytes;
        if (g.smem_bytes > (uint32_t)EMU_SMEM) return -1;
        emu::launch(1, 32 * g.warps, [&]() { lz4_small_kernel(S); });
    } else if (mode == 1) emu::launch(1, THREADS, [&]() { lz4_encode_kernel<1>(P); });
    else emu::launch(1, THREADS, [&]() { lz4_encode_kernel<0>(P); });
    const int extent = mode == 2 ? (int)small_geometry((uint32_t)block_len).smem_bytes : SM_TOTAL;
    for (int i = extent; i < EMU_SMEM + 256; ++i)
        if (smem[i] != 0xA5) return -7; // a write past the kernel's shared-memory extent
    return 0;
}

extern "C" int emu_lz4_decompress(const uint8_t *comp, size_t comp_len, const uint64_t *block_offsets, size_t nblocks, size_t block_len,
                                  uint8_t *out, size_t out_cap, uint32_t *block_out_len, uint64_t *result)
{
    using namespace lz4d;
    Params P;
    memset(&P, 0, sizeof P);
    P.comp = comp;
    P.offs = block_offsets;
    P.comp_len = comp_len;
    P.nblocks = (uint32_t)nblocks;
    P.first_block = 0;
    P.total_blocks = (uint32_t)nblocks;
    P.block_len = (uint32_t)block_len;
    P.out = out;
    P.out_cap = out_cap;
    P.block_out_len = block_out_len;
    P.result = result;
    result[0] = result[1] = result[2] = 0;
    const unsigned grid = (unsigned)((nblocks + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA);
    emu::launch(grid, WARPS_PER_CTA * 32, [&]() { lz4_decode_kernel(P); });
    return 0;
}
