// lz4-jpeg_b200/csrc/jpeg_entropy.cu — the entropy half of the "JPEG-like" codec's inverse chain (reference dialect), sm_100a.
//
// The reference decodes what it has just encoded while the Huffman TREE is still in memory (Algorithms/sequential/JPEG/
// JPEG.c:1253-1403): decode_huffman() walks the tree bit by bit (JPEG.c:1009-1033), inverse_RLE() expands the (count, value)
// pairs (JPEG.c:811-842), reverse_zigzag_pattern() puts the values back in row-major order (JPEG.c:729-764).  It never
// serialises the tree, so its bit strings alone are not decodable.  Here the tree is serialised:
//
//   jpeg_trees_kernel    quantised coefficients -> per (group, channel) the tree calculate_frequency / build_heap /
//                        build_huffman_tree construct (JPEG.c:864-961), written as
//                            u8 k | i16 value[k] (leaves = symbols in first-appearance order, node ids 0 .. k-1)
//                                 | { u8 left, u8 right }[k-1]  (internal nodes k .. 2k-2 in creation order; root = the last)
//                        at a fixed 1024 bytes per group: luma at +0 (<= 511 B), Cr at +512 (<= 255 B), Cb at +768
//   jpeg_entropy_decode_kernel   packed bit stream + offsets + bit lengths + trees -> quantised coefficients, by exactly the
//                        reference's three functions
//
// Together with jpeg_decode.cu (Inverse_quantize, IDCT, assemble_image) this is the whole inverse chain of JPEG_seq.exe, and
// the proof that the stream the encoder packs is decodable: tests assert decode(stream, trees) == the encoder's coefficients.
// One thread per (group, channel), arrays in local memory: the general procedure, not a tuned one ("next" tier).
#include "common.cuh"

namespace jpge {

constexpr int TREE_STRIDE = LJB_JPEG_TREE_BYTES; // per group
constexpr int THREADS = 96;                      // 32 groups x 3 channels per CTA

__device__ __forceinline__ int chan_off(int ch) { return ch == 0 ? 0 : (ch == 1 ? 512 : 768); }

// zigzag_pattern's walk (JPEG.c:693-727) for a W-column, 8-row block: order[i] = row-major index of the i-th value
template <int W>
__device__ __forceinline__ void zigzag_order(uint8_t (&order)[8 * W])
{
    int index = 0;
    for (int sum = 0; sum < W + 8 - 1; ++sum) {
        const int start_row = sum < W ? 0 : sum - W + 1;
        const int end_row = sum < 8 ? sum : 7;
        if ((sum & 1) == 0) {
            for (int row = end_row; row >= start_row; --row) order[index++] = (uint8_t)(row * W + (sum - row));
        } else {
            for (int row = start_row; row <= end_row; ++row) order[index++] = (uint8_t)(row * W + (sum - row));
        }
    }
}

struct TreeParams {
    const int16_t *coefs; // 128 per group: luma u*8+v, Cr 64 + u*4+v, Cb 96 + u*4+v
    size_t ngroups;
    uint8_t *trees;
};

template <int W>
__device__ void build_tree(const int16_t *c, uint8_t *dst)
{
    constexpr int N = 8 * W;
    uint8_t order[N];
    zigzag_order<W>(order);
    int16_t sym[2 * N];
    uint16_t cnt[4 * N];
    uint8_t heap[2 * N];
    int k = 0;
    // RLE (JPEG.c:767-809) feeding calculate_frequency (JPEG.c:864-885): every int of the (count, value) array is a symbol
    for (int i = 0; i < N;) {
        const int v = c[order[i]];
        int j = i + 1;
        while (j < N && c[order[j]] == v) ++j;
        const int two[2] = {j - i, v};
        for (int s = 0; s < 2; ++s) {
            int slot = -1;
            for (int q = 0; q < k; ++q)
                if (sym[q] == two[s]) {
                    slot = q;
                    break;
                }
            if (slot < 0) {
                slot = k++;
                sym[slot] = (int16_t)two[s];
                cnt[slot] = 0;
            }
            cnt[slot]++;
        }
        i = j;
    }
    auto heapify = [&](int size, int i) { // JPEG.c:894-911
        for (;;) {
            int smallest = i;
            const int l = 2 * i + 1, r = 2 * i + 2;
            if (l < size && cnt[heap[l]] < cnt[heap[smallest]]) smallest = l;
            if (r < size && cnt[heap[r]] < cnt[heap[smallest]]) smallest = r;
            if (smallest == i) return;
            const uint8_t t = heap[i];
            heap[i] = heap[smallest];
            heap[smallest] = t;
            i = smallest;
        }
    };
    for (int i = 0; i < k; ++i) heap[i] = (uint8_t)i;
    for (int i = k / 2 - 1; i >= 0; --i) heapify(k, i); // build_heap, JPEG.c:913-934
    dst[0] = (uint8_t)k;
    for (int i = 0; i < k; ++i) {
        dst[1 + 2 * i] = (uint8_t)(sym[i] & 0xFF);
        dst[2 + 2 * i] = (uint8_t)((uint16_t)sym[i] >> 8);
    }
    uint8_t *ch = dst + 1 + 2 * k;
    int size = k, next = k;
    while (size > 1) { // build_huffman_tree, JPEG.c:936-961: pop two, append their parent at the END of the array (no sift-up)
        const int left = heap[0];
        heap[0] = heap[--size];
        heapify(size, 0);
        const int right = heap[0];
        heap[0] = heap[--size];
        heapify(size, 0);
        cnt[next] = (uint16_t)(cnt[left] + cnt[right]);
        ch[2 * (next - k)] = (uint8_t)left;
        ch[2 * (next - k) + 1] = (uint8_t)right;
        heap[size++] = (uint8_t)next;
        ++next;
    }
}

__global__ void __launch_bounds__(THREADS) jpeg_trees_kernel(TreeParams P)
{
    const size_t g = (size_t)blockIdx.x * 32 + threadIdx.x / 3;
    const int ch = threadIdx.x % 3;
    if (g >= P.ngroups) return;
    const int16_t *c = P.coefs + g * 128 + (ch == 0 ? 0 : (ch == 1 ? 64 : 96));
    uint8_t *dst = P.trees + g * TREE_STRIDE + chan_off(ch);
    if (ch == 0) build_tree<8>(c, dst);
    else build_tree<4>(c, dst);
}

struct DecodeParams {
    const uint8_t *stream;
    size_t stream_len;
    const uint64_t *group_offsets; // ngroups + 1
    const uint16_t *group_bits;    // 3 per group: luma, Cr, Cb
    const uint8_t *trees;
    size_t ngroups;
    uint64_t offs_base;            // subtracted from every offset (the offsets of a shard are stream-global)
    int16_t *coefs;                // out, 128 per group
    uint64_t *result;              // [2] error flags: bit1 = malformed (offsets, tree or bit string)
};

template <int W>
__device__ bool decode_channel(const uint8_t *s, size_t bit0, uint32_t nbits, const uint8_t *tree, int16_t *out)
{
    constexpr int N = 8 * W;
    const int k = tree[0];
    if (k < 1 || k > 2 * N) return false;
    const uint8_t *leaf = tree + 1, *ch = tree + 1 + 2 * k;
    const int root = k == 1 ? 0 : 2 * k - 2;
    // decode_huffman, JPEG.c:1009-1033: walk from the root, a leaf yields its value and sends the walk back to the root
    int16_t rle[2 * N];
    int m = 0, node = root;
    for (uint32_t i = 0; i < nbits; ++i) {
        if (node < k) return false; // (only possible with a one-symbol tree and a non-empty string)
        const size_t bp = bit0 + i;
        const int bit = (s[bp >> 3] >> (7 - (bp & 7))) & 1;
        node = ch[2 * (node - k) + bit];
        if (node >= 2 * k - 1) return false;
        if (node < k) {
            if (m >= 2 * N) return false;
            rle[m++] = (int16_t)((uint16_t)leaf[2 * node] | ((uint16_t)leaf[2 * node + 1] << 8));
            node = root;
        }
    }
    if (node != root) return false; // the string ends inside a code
    // inverse_RLE, JPEG.c:811-842 (counts clipped to the block, the rest zero)
    int16_t z[N];
    int index = 0;
    for (int i = 0; i + 1 < m; i += 2) {
        int count = rle[i];
        const int value = rle[i + 1];
        if (index + count > N) count = N - index;
        for (int j = 0; j < count; ++j) z[index++] = (int16_t)value;
    }
    while (index < N) z[index++] = 0;
    // reverse_zigzag_pattern, JPEG.c:729-764
    uint8_t order[N];
    zigzag_order<W>(order);
    for (int i = 0; i < N; ++i) out[order[i]] = z[i];
    return true;
}

__global__ void __launch_bounds__(THREADS) jpeg_entropy_decode_kernel(DecodeParams P)
{
    const size_t g = (size_t)blockIdx.x * 32 + threadIdx.x / 3;
    const int ch = threadIdx.x % 3;
    if (g >= P.ngroups) return;
    const uint64_t o0 = P.group_offsets[g], o1 = P.group_offsets[g + 1];
    const uint32_t bl = P.group_bits[3 * g], br = P.group_bits[3 * g + 1], bb = P.group_bits[3 * g + 2];
    bool ok = o0 >= P.offs_base && o1 >= o0 && o1 - P.offs_base <= P.stream_len && (uint64_t)(bl + br + bb + 7) / 8 <= o1 - o0;
    if (ok) {
        const size_t bit0 = (size_t)(o0 - P.offs_base) * 8 + (ch == 0 ? 0u : (ch == 1 ? bl : bl + br));
        const uint8_t *tree = P.trees + g * TREE_STRIDE + chan_off(ch);
        int16_t *out = P.coefs + g * 128 + (ch == 0 ? 0 : (ch == 1 ? 64 : 96));
        ok = ch == 0 ? decode_channel<8>(P.stream, bit0, bl, tree, out) : decode_channel<4>(P.stream, bit0, ch == 1 ? br : bb, tree, out);
    }
    if (!ok) atomicOr((unsigned long long *)&P.result[2], 2ull);
}

} // namespace jpge

extern "C" int ljb_jpeg_trees_dev(ljb_ctx *ctx, const int16_t *d_coefs, size_t ngroups, uint8_t *d_trees)
{
    using namespace jpge;
    if (!ctx || !d_coefs || !d_trees || ngroups == 0) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    TreeParams P;
    P.coefs = d_coefs;
    P.ngroups = ngroups;
    P.trees = d_trees;
    ctx->kernel_ms_summed = 0;
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    jpeg_trees_kernel<<<(unsigned)((ngroups + 31) / 32), THREADS, 0, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

extern "C" int ljb_jpeg_entropy_decode_dev(ljb_ctx *ctx, const uint8_t *d_stream, size_t stream_len, const uint64_t *d_group_offsets,
                                           const uint16_t *d_group_bits, const uint8_t *d_trees, size_t ngroups, uint64_t offs_base,
                                           int16_t *d_coefs, uint64_t *d_result)
{
    using namespace jpge;
    if (!ctx || !d_stream || !d_group_offsets || !d_group_bits || !d_trees || !d_coefs || !d_result || ngroups == 0) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    DecodeParams P;
    P.stream = d_stream;
    P.stream_len = stream_len;
    P.group_offsets = d_group_offsets;
    P.group_bits = d_group_bits;
    P.trees = d_trees;
    P.ngroups = ngroups;
    P.offs_base = offs_base;
    P.coefs = d_coefs;
    P.result = d_result;
    ctx->kernel_ms_summed = 0;
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    jpeg_entropy_decode_kernel<<<(unsigned)((ngroups + 31) / 32), THREADS, 0, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

// Host-buffer forms (upload, one kernel, download).
extern "C" int ljb_jpeg_trees(ljb_ctx *ctx, const int16_t *coefs, size_t ngroups, uint8_t *trees)
{
    if (!ctx || !coefs || !trees || ngroups == 0) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], ngroups * 128 * sizeof(int16_t) + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], ngroups * (size_t)LJB_JPEG_TREE_BYTES + 64)) != 0) return rc;
    LJB_CUDA(cudaMemcpyAsync(ctx->d_pin[0], coefs, ngroups * 128 * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    LJB_CUDA(cudaMemsetAsync(ctx->d_pout[0], 0, ngroups * (size_t)LJB_JPEG_TREE_BYTES, ctx->stream));
    if ((rc = ljb_jpeg_trees_dev(ctx, (const int16_t *)ctx->d_pin[0], ngroups, (uint8_t *)ctx->d_pout[0])) != 0) return rc;
    LJB_CUDA(cudaMemcpyAsync(trees, ctx->d_pout[0], ngroups * (size_t)LJB_JPEG_TREE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}

extern "C" int ljb_jpeg_entropy_decode(ljb_ctx *ctx, const uint8_t *stream, size_t stream_len, const uint64_t *group_offsets,
                                       const uint16_t *group_bits, const uint8_t *trees, size_t ngroups, int16_t *coefs)
{
    if (!ctx || !stream || !group_offsets || !group_bits || !trees || !coefs || ngroups == 0) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t o_off = (stream_len + 255) & ~(size_t)255;
    const size_t o_bits = o_off + (((ngroups + 1) * 8 + 255) & ~(size_t)255);
    const size_t o_tree = o_bits + ((ngroups * 6 + 255) & ~(size_t)255);
    const size_t in_bytes = o_tree + ngroups * (size_t)LJB_JPEG_TREE_BYTES;
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], in_bytes + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], ngroups * 128 * sizeof(int16_t) + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, 64)) != 0) return rc;
    uint8_t *d = (uint8_t *)ctx->d_pin[0];
    LJB_CUDA(cudaMemcpyAsync(d, stream, stream_len, cudaMemcpyHostToDevice, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(d + o_off, group_offsets, (ngroups + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(d + o_bits, group_bits, ngroups * 6, cudaMemcpyHostToDevice, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(d + o_tree, trees, ngroups * (size_t)LJB_JPEG_TREE_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t *d_res = (uint64_t *)ctx->d_small;
    if ((rc = ljb_jpeg_entropy_decode_dev(ctx, d, stream_len, (const uint64_t *)(d + o_off), (const uint16_t *)(d + o_bits), d + o_tree, ngroups,
                                          0, (int16_t *)ctx->d_pout[0], d_res)) != 0)
        return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(coefs, ctx->d_pout[0], ngroups * 128 * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return (res[2] & 2) ? LJB_E_FORMAT : LJB_OK;
}
