// lz4-jpeg_b200/csrc/jpeg_decode.cu — decode half of the "JPEG-like" codec (reference dialect) for sm_100a.
//
// Replaces the tail of the reference's main() (Algorithms/sequential/JPEG/JPEG.c:1408-1428):
//   Inverse_quantize                     JPEG.c:631-638   coefficient * table entry
//   inverse_discrete_cosine_transform    JPEG.c:399-448   fp64 O(N^4) sum, round(sum + 128), clamp
//   assemble_image                       JPEG.c:552-619   4:2:2 chroma lookup, YCbCr -> RGB with (int) truncations
// Input are the quantised coefficients the encoder produced (the reference's Huffman / RLE / zig-zag stages are a
// loss-free round trip in memory, JPEG.c:1253-1403; its code tables are never serialised, so a bit stream alone
// is not decodable).  Groups beyond ceil(w*h/64) are left unprocessed by the reference (JPEG.c:1131): they keep
// the colour-converted samples of the original image, which is why the original can be passed in.
//
// 8 threads per group (one output row each), 16 groups per CTA.  A separable fp64 evaluation gives every sample
// to ~1e-12; a value within 1e-9 of a rounding boundary (x.5) is re-evaluated in the reference's own summation
// order (u outer, v inner, (((au*av)*c)*cos_x)*cos_y, no FMA) with glibc's cos()/sqrt() doubles.
#include "common.cuh"

namespace jpgd {

#include "jpeg_tables.inc"
#include "jpeg_colour.cuh"

constexpr int GROUPS = 16;
constexpr int THREADS = GROUPS * 8;

__constant__ double kQL[64] = {8,  6,  6,  8,  10, 14, 18, 22, 6,  6,  7,  9,  12, 20, 22, 20, 6,  7,  8,  10, 14, 22,
                               25, 22, 8,  9,  10, 14, 18, 28, 27, 22, 10, 12, 14, 18, 22, 35, 33, 26, 14, 18, 22, 22,
                               27, 33, 36, 30, 18, 22, 26, 28, 33, 40, 40, 34, 22, 26, 28, 30, 36, 34, 35, 33}; // JPEG.c:12-20
__constant__ double kQC[32] = {17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99,
                               66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}; // JPEG.c:22-27 as 8 rows x 4

struct Params {
    const int16_t *coefs; // 128 per group, groups [0, total)
    int w, h;
    size_t total;         // ceil(w*h/64): groups the reference processes
    size_t tiled;         // ceil(w/8) * ceil(h/8)
    const uint8_t *orig;  // original RGBA (only read for groups >= total); may be null
    size_t orig_stride;
    uint8_t *out;         // RGBA image, or null when only the samples are wanted
    size_t out_stride;
    uint8_t *samples;     // optional: the reconstructed samples of every group in PixelGroup order (JPEG.c:44-46):
                          // lum_values[64], b_values[32], r_values[32] — what process() leaves in its argument (P-JPG:1247-1251)
    uint64_t *result;     // [2] error flags: bit0 = unprocessed groups exist but no original was given
};

// the reference's summation for sample (x, y) of a W-column channel; cq = dequantised coefficients (u*W + v)
template <int W>
__device__ __noinline__ double exact_sample(const double *cq, int x, int y)
{
    double sum = 0.0;
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
        const double au = u == 0 ? kAlpha8[0] : kAlpha8[1];
        const double cx = kCos8[x * 8 + u];
#pragma unroll 1
        for (int v = 0; v < W; ++v) {
            const double av = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) : (v == 0 ? kAlpha4[0] : kAlpha4[1]);
            const double cy = (W == 8) ? kCos8[y * 8 + v] : kCos4[y * 4 + v];
            sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(au, av), cq[u * W + v]), cx), cy));
        }
    }
    return sum;
}

template <int W>
__device__ __forceinline__ void idct_row(const double *cq, int x, int (&out)[W])
{
    double T[W]; // T[v] = sum_u alpha_u cos8[x][u] cq[u][v]
#pragma unroll
    for (int v = 0; v < W; ++v) T[v] = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const double f = (u == 0 ? kAlpha8[0] : kAlpha8[1]) * kCos8[x * 8 + u];
#pragma unroll
        for (int v = 0; v < W; ++v) T[v] = fma(f, cq[u * W + v], T[v]);
    }
#pragma unroll
    for (int y = 0; y < W; ++y) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < W; ++v) {
            const double f = (W == 8) ? (v == 0 ? kAlpha8[0] : kAlpha8[1]) * kCos8[y * 8 + v]
                                      : (v == 0 ? kAlpha4[0] : kAlpha4[1]) * kCos4[y * 4 + v];
            s = fma(f, T[v], s);
        }
        double t = s + 128.0;
        if (fabs((t - floor(t)) - 0.5) < 1e-9) t = __dadd_rn(exact_sample<W>(cq, x, y), 128.0); // rounding boundary
        const int v = (int)round(t); // JPEG.c:441
        out[y] = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
}

__global__ void __launch_bounds__(THREADS) jpeg_decode_kernel(Params P)
{
    __shared__ double cq[GROUPS][128];
    const int tid = threadIdx.x, gi = tid >> 3, x = tid & 7;
    const size_t g = (size_t)blockIdx.x * GROUPS + gi;
    const size_t bpr = ((size_t)P.w + 7) / 8;
    if (g < P.total) { // Inverse_quantize, JPEG.c:631-638
        const int16_t *c = P.coefs + g * 128;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = x * 16 + k;
            cq[gi][i] = __dmul_rn((double)c[i], i < 64 ? kQL[i] : kQC[(i - 64) & 31]);
        }
    }
    __syncthreads();
    if (g >= P.tiled) return;
    const size_t brow = g / bpr, bcol = g % bpr;
    const size_t row = brow * 8 + x;
    if (row >= (size_t)P.h) return;
    int Y[8], Cr[4], Cb[4];
    if (g < P.total) {
        idct_row<8>(&cq[gi][0], x, Y);
        idct_row<4>(&cq[gi][64], x, Cr);
        idct_row<4>(&cq[gi][96], x, Cb);
    } else { // unprocessed group: the samples divide_image stored (JPEG.c:496-550), zero outside the image
        if (!P.orig) {
            if (x == 0) atomicOr((unsigned long long *)&P.result[2], 1ull);
            return;
        }
#pragma unroll
        for (int lc = 0; lc < 8; ++lc) {
            const size_t col = bcol * 8 + lc;
            int r = 0, gg = 0, b = 0;
            const bool in = col < (size_t)P.w;
            if (in) {
                const uint8_t *q = P.orig + row * P.orig_stride + col * 4;
                r = q[0];
                gg = q[1];
                b = q[2];
            }
            Y[lc] = in ? luma_of(r, gg, b) : 0;
            if (lc & 1) { // the chroma sample of local column lc-1 is the original chroma at column lc
                Cr[lc >> 1] = in ? cr_of(r, gg, b) : 0;
                Cb[lc >> 1] = in ? cb_of(r, gg, b) : 0;
            }
        }
    }
    if (P.samples) {
        uint8_t *sp = P.samples + g * 128;
#pragma unroll
        for (int lc = 0; lc < 8; ++lc) sp[x * 8 + lc] = (uint8_t)Y[lc];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            sp[64 + x * 4 + c] = (uint8_t)Cb[c];
            sp[96 + x * 4 + c] = (uint8_t)Cr[c];
        }
    }
    if (!P.out) return;
    // assemble_image, JPEG.c:552-619
#pragma unroll
    for (int lc = 0; lc < 8; ++lc) {
        const size_t col = bcol * 8 + lc;
        if (col >= (size_t)P.w) break;
        const int y = Y[lc], cr = Cr[lc >> 1] - 128, cb = Cb[lc >> 1] - 128;
        int R = y + (int)__dmul_rn(1.402, (double)cr);
        int G = y - (int)__dmul_rn(0.344136, (double)cb) - (int)__dmul_rn(0.714136, (double)cr);
        int B = y + (int)__dmul_rn(1.772, (double)cb);
        R = R < 0 ? 0 : (R > 255 ? 255 : R);
        G = G < 0 ? 0 : (G > 255 ? 255 : G);
        B = B < 0 ? 0 : (B > 255 ? 255 : B);
        uint8_t *o = P.out + row * P.out_stride + col * 4;
        *reinterpret_cast<uchar4 *>(o) = make_uchar4((unsigned char)R, (unsigned char)G, (unsigned char)B, 255);
    }
}

} // namespace jpgd

extern "C" int ljb_jpeg_decode_coefs_dev(ljb_ctx *ctx, const int16_t *d_coefs, int w, int h, const uint8_t *d_orig_rgba,
                                         size_t orig_stride, uint8_t *d_out_rgba, size_t out_stride, uint64_t *d_result)
{
    using namespace jpgd;
    if (!ctx || !d_coefs || !d_out_rgba || !d_result || w <= 0 || h <= 0 || (w & 1) || out_stride < (size_t)w * 4 || (out_stride & 3) ||
        (reinterpret_cast<uintptr_t>(d_out_rgba) & 3))
        return LJB_E_ARG;
    if (d_orig_rgba && orig_stride < (size_t)w * 4) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    Params P;
    P.coefs = d_coefs;
    P.w = w;
    P.h = h;
    P.total = ljb_jpeg_group_count(w, h);
    P.tiled = (((size_t)w + 7) / 8) * (((size_t)h + 7) / 8);
    P.orig = d_orig_rgba;
    P.orig_stride = orig_stride;
    P.out = d_out_rgba;
    P.out_stride = out_stride;
    P.samples = nullptr;
    P.result = d_result;
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    const unsigned grid = (unsigned)((P.tiled + GROUPS - 1) / GROUPS);
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    jpeg_decode_kernel<<<grid, THREADS, 0, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

extern "C" int ljb_jpeg_decode_coefs(ljb_ctx *ctx, const int16_t *coefs, int w, int h, const uint8_t *orig_rgba, size_t orig_stride,
                                     uint8_t *out_rgba, size_t out_stride)
{
    if (!ctx || !coefs || !out_rgba || w <= 0 || h <= 0 || (w & 1) || out_stride < (size_t)w * 4) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    const size_t total = ljb_jpeg_group_count(w, h);
    const size_t dstride = (size_t)w * 4;
    const size_t img = dstride * (size_t)h;
    int rc;
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], total * 128 * sizeof(int16_t) + (orig_rgba ? img : 0) + 256)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], img + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_small, &ctx->small_bytes, 64)) != 0) return rc;
    int16_t *d_coefs = (int16_t *)ctx->d_pin[0];
    uint8_t *d_orig = orig_rgba ? (uint8_t *)ctx->d_pin[0] + ((total * 128 * sizeof(int16_t) + 255) & ~(size_t)255) : nullptr;
    uint64_t *d_res = (uint64_t *)ctx->d_small;
    LJB_CUDA(cudaMemcpyAsync(d_coefs, coefs, total * 128 * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    if (orig_rgba)
        LJB_CUDA(cudaMemcpy2DAsync(d_orig, dstride, orig_rgba, orig_stride, dstride, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    rc = ljb_jpeg_decode_coefs_dev(ctx, d_coefs, w, h, d_orig, dstride, (uint8_t *)ctx->d_pout[0], dstride, d_res);
    if (rc != 0) return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaMemcpy2DAsync(out_rgba, out_stride, ctx->d_pout[0], dstride, dstride, (size_t)h, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res[2] & 1) return LJB_E_ARG; // unprocessed groups exist (w or h not a multiple of 8) and no original image was given
    return LJB_OK;
}

// The inverse chain of process() (Algorithms/parallel/JPEG/JPEG.c:1242-1251) for groups given by their quantised coefficients:
// Inverse_quantize -> inverse_discrete_cosine_transform of the three channels, samples in PixelGroup order (128 bytes per group).
extern "C" int ljb_jpeg_decode_groups_dev(ljb_ctx *ctx, const int16_t *d_coefs, size_t ngroups, uint8_t *d_samples, uint64_t *d_result)
{
    using namespace jpgd;
    if (!ctx || !d_coefs || !d_samples || !d_result || ngroups == 0 || ngroups > 0x0FFFFFFFull) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    Params P;
    P.coefs = d_coefs;
    P.w = 8; // a column of groups
    P.h = (int)(8 * ngroups);
    P.total = ngroups;
    P.tiled = ngroups;
    P.orig = nullptr;
    P.orig_stride = 0;
    P.out = nullptr;
    P.out_stride = 0;
    P.samples = d_samples;
    P.result = d_result;
    LJB_CUDA(cudaMemsetAsync(d_result, 0, 3 * sizeof(uint64_t), ctx->stream));
    const unsigned grid = (unsigned)((ngroups + GROUPS - 1) / GROUPS);
    ctx->kernel_ms_summed = 0;
    LJB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    jpeg_decode_kernel<<<grid, THREADS, 0, ctx->stream>>>(P);
    LJB_CUDA(cudaGetLastError());
    LJB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->launches += 1;
    return LJB_OK;
}

// process() for n groups at once, host buffers: forward chain (DCT, Quantize, zig-zag, RLE, Huffman: the packed strings are
// produced and dropped, as the reference's are), then the inverse chain.  samples: in = the groups' lum/b/r values, out = what
// the reference leaves in them; coefs: out, the quantised coefficients (128 int16 per group).
extern "C" int ljb_jpeg_process_groups(ljb_ctx *ctx, uint8_t *samples, size_t ngroups, int16_t *coefs)
{
    if (!ctx || !samples || !coefs || ngroups == 0 || ngroups > 0x0FFFFFFFull) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const size_t cap = ljb_jpeg_bound(ngroups);
    const size_t o_offs = (cap + 255) & ~(size_t)255;
    const size_t o_bits = o_offs + (((ngroups + 1) * 8 + 255) & ~(size_t)255);
    const size_t o_coef = o_bits + ((ngroups * 6 + 255) & ~(size_t)255);
    const size_t o_res = o_coef + ngroups * 256;
    if ((rc = ljb_ensure(&ctx->d_pin[0], &ctx->pin_bytes[0], ngroups * 128 + 64)) != 0) return rc;
    if ((rc = ljb_ensure(&ctx->d_pout[0], &ctx->pout_bytes[0], o_res + 64)) != 0) return rc;
    uint8_t *d_s = (uint8_t *)ctx->d_pin[0], *d = (uint8_t *)ctx->d_pout[0];
    uint64_t *d_res = (uint64_t *)(d + o_res);
    LJB_CUDA(cudaMemcpyAsync(d_s, samples, ngroups * 128, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ljb_jpeg_encode_groups_dev(ctx, d_s, ngroups, d, cap, (uint64_t *)(d + o_offs), (uint16_t *)(d + o_bits), (int16_t *)(d + o_coef),
                                         d_res)) != 0)
        return rc;
    uint64_t res[3];
    LJB_CUDA(cudaMemcpyAsync(res, d_res, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res[2] & 2) return LJB_E_UNSUPPORTED;
    if (res[2] & 1) return LJB_E_CAPACITY;
    if ((rc = ljb_jpeg_decode_groups_dev(ctx, (const int16_t *)(d + o_coef), ngroups, d_s, d_res)) != 0) return rc;
    LJB_CUDA(cudaMemcpyAsync(coefs, d + o_coef, ngroups * 256, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaMemcpyAsync(samples, d_s, ngroups * 128, cudaMemcpyDeviceToHost, ctx->stream));
    LJB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LJB_OK;
}
