// lz4-jpeg_b200/csrc/comm.cu — multi-GPU entry points for a C host (include/ljb_comm.h): one process, a context and an NCCL
// communicator per GPU, shards of independent units, ONE ncclAllGather of per-GPU byte totals.
// Built into lz4-jpeg_b200/libljb_comm.so, the only part of the package that links NCCL.
#include "common.cuh"

#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/ljb_comm.h"

struct ljb_comm {
    int n = 0;
    std::vector<ljb_ctx *> ctx;
    std::vector<ncclComm_t> nccl;
    std::vector<uint64_t *> d_all;    // per GPU: n gathered totals
    std::vector<uint64_t *> d_mine;   // per GPU: staging for this GPU's total
    std::vector<void *> d_in, d_out, d_small;
    std::vector<size_t> in_bytes, out_bytes, small_bytes;
    uint64_t *h_all = nullptr;        // pinned: the gathered totals
};

#define COMM_CUDA(x)                                                          \
    do {                                                                      \
        cudaError_t e__ = (x);                                                \
        if (e__ != cudaSuccess) return ljb_set_cuda_error(e__, #x, __LINE__); \
    } while (0)
#define COMM_NCCL(x)                                                                    \
    do {                                                                                \
        ncclResult_t r__ = (x);                                                         \
        if (r__ != ncclSuccess) {                                                       \
            fprintf(stderr, "[ljb_comm] %s: %s\n", #x, ncclGetErrorString(r__));        \
            return LJB_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

extern "C" int ljb_comm_size(const ljb_comm *c) { return c ? c->n : 0; }
extern "C" ljb_ctx *ljb_comm_ctx(ljb_comm *c, int rank) { return (c && rank >= 0 && rank < c->n) ? c->ctx[rank] : nullptr; }

extern "C" void ljb_comm_destroy(ljb_comm *c)
{
    if (!c) return;
    for (int r = 0; r < (int)c->nccl.size(); ++r) {
        cudaSetDevice(r);
        if (c->nccl[r]) ncclCommDestroy(c->nccl[r]);
    }
    for (int r = 0; r < (int)c->ctx.size(); ++r) {
        cudaSetDevice(r);
        if (r < (int)c->d_all.size()) cudaFree(c->d_all[r]);
        if (r < (int)c->d_mine.size()) cudaFree(c->d_mine[r]);
        if (r < (int)c->d_in.size()) cudaFree(c->d_in[r]);
        if (r < (int)c->d_out.size()) cudaFree(c->d_out[r]);
        if (r < (int)c->d_small.size()) cudaFree(c->d_small[r]);
        ljb_ctx_destroy(c->ctx[r]);
    }
    if (c->h_all) cudaFreeHost(c->h_all);
    delete c;
}

extern "C" int ljb_comm_create(int ngpus, ljb_comm **out)
{
    if (!out || ngpus < 1) return LJB_E_ARG;
    *out = nullptr;
    int count = 0;
    COMM_CUDA(cudaGetDeviceCount(&count));
    if (ngpus > count) return LJB_E_ARG;
    ljb_comm *c = new ljb_comm();
    c->n = ngpus;
    c->nccl.assign(ngpus, nullptr);
    c->d_all.assign(ngpus, nullptr);
    c->d_mine.assign(ngpus, nullptr);
    c->d_in.assign(ngpus, nullptr);
    c->d_out.assign(ngpus, nullptr);
    c->d_small.assign(ngpus, nullptr);
    c->in_bytes.assign(ngpus, 0);
    c->out_bytes.assign(ngpus, 0);
    c->small_bytes.assign(ngpus, 0);
    for (int r = 0; r < ngpus; ++r) {
        ljb_ctx *x = nullptr;
        int rc = ljb_ctx_create(r, &x);
        if (rc != LJB_OK) {
            ljb_comm_destroy(c);
            return rc;
        }
        c->ctx.push_back(x);
        if (cudaMalloc(&c->d_all[r], sizeof(uint64_t) * ngpus) != cudaSuccess || cudaMalloc(&c->d_mine[r], sizeof(uint64_t)) != cudaSuccess) {
            ljb_comm_destroy(c);
            return LJB_E_CUDA;
        }
    }
    std::vector<int> devs(ngpus);
    for (int r = 0; r < ngpus; ++r) devs[r] = r;
    if (ncclCommInitAll(c->nccl.data(), ngpus, devs.data()) != ncclSuccess) {
        fprintf(stderr, "[ljb_comm] ncclCommInitAll failed\n");
        ljb_comm_destroy(c);
        return LJB_E_CUDA;
    }
    if (cudaMallocHost((void **)&c->h_all, sizeof(uint64_t) * ngpus) != cudaSuccess) {
        ljb_comm_destroy(c);
        return LJB_E_CUDA;
    }
    *out = c;
    return LJB_OK;
}

extern "C" int ljb_comm_gather_totals(ljb_comm *c, const uint64_t *const *d_totals, uint64_t *bases, uint64_t *grand)
{
    if (!c || !d_totals || !bases) return LJB_E_ARG;
    COMM_NCCL(ncclGroupStart());
    for (int r = 0; r < c->n; ++r)
        COMM_NCCL(ncclAllGather(d_totals[r], c->d_all[r], 1, ncclUint64, c->nccl[r], c->ctx[r]->stream));
    COMM_NCCL(ncclGroupEnd());
    COMM_CUDA(cudaSetDevice(0));
    COMM_CUDA(cudaMemcpyAsync(c->h_all, c->d_all[0], sizeof(uint64_t) * c->n, cudaMemcpyDeviceToHost, c->ctx[0]->stream));
    for (int r = 0; r < c->n; ++r) {
        COMM_CUDA(cudaSetDevice(r));
        COMM_CUDA(cudaStreamSynchronize(c->ctx[r]->stream));
    }
    uint64_t run = 0;
    for (int r = 0; r < c->n; ++r) {
        bases[r] = run;
        run += c->h_all[r];
    }
    if (grand) *grand = run;
    return LJB_OK;
}

static int comm_ensure(ljb_comm *c, int r, void **p, size_t *have, size_t want)
{
    cudaSetDevice(r);
    (void)c;
    return ljb_ensure(p, have, want);
}

extern "C" int ljb_comm_lz4_compress(ljb_comm *c, const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                                     uint64_t *block_offsets, size_t *out_len, uint64_t *phantom)
{
    if (!c || !in || !out || n == 0 || block_len == 0 || block_len > LJB_LZ4_MAX_BLOCK) return LJB_E_ARG;
    const size_t nblocks = ljb_lz4_block_count(n, block_len);
    const size_t per = (nblocks + c->n - 1) / c->n;
    struct Shard {
        size_t first = 0, count = 0, lo = 0, hi = 0, cap = 0;
    };
    std::vector<Shard> sh(c->n);
    int rc;
    for (int r = 0; r < c->n; ++r) {
        Shard &s = sh[r];
        s.first = std::min(nblocks, (size_t)r * per);
        s.count = std::min(nblocks, s.first + per) - s.first;
        s.lo = s.first * block_len;
        s.hi = std::min(n, (s.first + s.count) * block_len);
        if (s.count == 0) continue;
        s.cap = ljb_lz4_bound(s.hi - s.lo, block_len);
        if ((rc = comm_ensure(c, r, &c->d_in[r], &c->in_bytes[r], s.hi - s.lo + 64)) != 0) return rc;
        if ((rc = comm_ensure(c, r, &c->d_out[r], &c->out_bytes[r], s.cap + 64)) != 0) return rc;
        if ((rc = comm_ensure(c, r, &c->d_small[r], &c->small_bytes[r], (s.count + 1 + 3) * sizeof(uint64_t))) != 0) return rc;
    }
    // every GPU: upload its shard and encode it (asynchronous on its own stream: the GPUs run concurrently)
    for (int r = 0; r < c->n; ++r) {
        const Shard &s = sh[r];
        COMM_CUDA(cudaSetDevice(r));
        uint64_t *d_res = (uint64_t *)c->d_small[r];
        if (s.count == 0) {
            COMM_CUDA(cudaMemsetAsync(c->d_mine[r], 0, sizeof(uint64_t), c->ctx[r]->stream));
            continue;
        }
        uint64_t *d_offs = d_res + 3;
        COMM_CUDA(cudaMemcpyAsync(c->d_in[r], in + s.lo, s.hi - s.lo, cudaMemcpyHostToDevice, c->ctx[r]->stream));
        if ((rc = ljb_lz4_compress_dev(c->ctx[r], (const uint8_t *)c->d_in[r], s.hi - s.lo, block_len, (uint8_t *)c->d_out[r], s.cap, d_offs,
                                       d_res, s.first, nblocks)) != 0)
            return rc;
        COMM_CUDA(cudaMemcpyAsync(c->d_mine[r], d_res, sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->ctx[r]->stream));
    }
    // the one collective
    std::vector<const uint64_t *> totals(c->n);
    std::vector<uint64_t> bases(c->n);
    uint64_t grand = 0;
    for (int r = 0; r < c->n; ++r) totals[r] = c->d_mine[r];
    if ((rc = ljb_comm_gather_totals(c, totals.data(), bases.data(), &grand)) != 0) return rc;
    if (out_len) *out_len = (size_t)grand;
    if (grand > out_cap) return LJB_E_CAPACITY;
    // every shard to its place; its offset table (relative to the shard's own stream) is rebased by the shard's base
    uint64_t ph_total = 0;
    std::vector<std::vector<uint64_t>> hres(c->n, std::vector<uint64_t>(3, 0)), hoffs(c->n);
    for (int r = 0; r < c->n; ++r) {
        const Shard &s = sh[r];
        if (s.count == 0) continue;
        COMM_CUDA(cudaSetDevice(r));
        uint64_t *d_res = (uint64_t *)c->d_small[r];
        COMM_CUDA(cudaMemcpyAsync(hres[r].data(), d_res, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        COMM_CUDA(cudaMemcpyAsync(out + bases[r], c->d_out[r], (size_t)c->h_all[r], cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        if (block_offsets) {
            hoffs[r].resize(s.count + 1);
            COMM_CUDA(cudaMemcpyAsync(hoffs[r].data(), d_res + 3, (s.count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        }
    }
    for (int r = 0; r < c->n; ++r) {
        COMM_CUDA(cudaSetDevice(r));
        COMM_CUDA(cudaStreamSynchronize(c->ctx[r]->stream));
    }
    for (int r = 0; r < c->n; ++r) {
        if (block_offsets)
            for (size_t k = 0; k < hoffs[r].size(); ++k) block_offsets[sh[r].first + k] = hoffs[r][k] + bases[r];
        ph_total += hres[r][1];
        if (hres[r][2] & 1) return LJB_E_CAPACITY;
    }
    if (phantom) *phantom = ph_total;
    return LJB_OK;
}

static int comm_jpeg_encode(ljb_comm *c, const uint8_t *rgba, int bpp, int w, int h, size_t stride, uint8_t *out, size_t out_cap,
                            uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    if (!c || !rgba || !out || w <= 0 || h <= 0 || (w & 1) || stride < (size_t)w * (size_t)bpp) return LJB_E_ARG;
    const size_t bpr = ((size_t)w + 7) / 8, total = ljb_jpeg_group_count(w, h), rows = (total + bpr - 1) / bpr;
    const size_t per = (rows + c->n - 1) / c->n;
    const size_t dstride = ((size_t)w * (size_t)bpp + 15) & ~(size_t)15;
    struct Shard {
        size_t g0 = 0, g1 = 0, y0 = 0, y1 = 0, cap = 0;
    };
    std::vector<Shard> sh(c->n);
    int rc;
    for (int r = 0; r < c->n; ++r) {
        Shard &s = sh[r];
        const size_t r0 = std::min(rows, (size_t)r * per), r1 = std::min(rows, r0 + per);
        s.g0 = std::min(total, r0 * bpr);
        s.g1 = std::min(total, r1 * bpr);
        s.y0 = r0 * 8;
        s.y1 = std::min((size_t)h, r1 * 8);
        if (s.g1 == s.g0) continue;
        const size_t ng = s.g1 - s.g0;
        s.cap = ljb_jpeg_bound(ng);
        if ((rc = comm_ensure(c, r, &c->d_in[r], &c->in_bytes[r], dstride * (s.y1 - s.y0) + 64)) != 0) return rc;
        if ((rc = comm_ensure(c, r, &c->d_out[r], &c->out_bytes[r], s.cap + 64)) != 0) return rc;
        if ((rc = comm_ensure(c, r, &c->d_small[r], &c->small_bytes[r], (ng + 1 + 3) * sizeof(uint64_t) + ng * 3 * sizeof(uint16_t) + 64)) != 0)
            return rc;
    }
    for (int r = 0; r < c->n; ++r) {
        const Shard &s = sh[r];
        COMM_CUDA(cudaSetDevice(r));
        if (s.g1 == s.g0) {
            COMM_CUDA(cudaMemsetAsync(c->d_mine[r], 0, sizeof(uint64_t), c->ctx[r]->stream));
            continue;
        }
        const size_t ng = s.g1 - s.g0;
        uint64_t *d_res = (uint64_t *)c->d_small[r], *d_offs = d_res + 3;
        uint16_t *d_bits = (uint16_t *)(d_offs + ng + 1);
        COMM_CUDA(cudaMemcpy2DAsync(c->d_in[r], dstride, rgba + s.y0 * stride, stride, (size_t)w * (size_t)bpp, s.y1 - s.y0, cudaMemcpyHostToDevice,
                                    c->ctx[r]->stream));
        // the kernel addresses rows of the whole image: bias the band's pointer by the rows before it (never dereferenced there)
        const uint8_t *biased = (const uint8_t *)c->d_in[r] - s.y0 * dstride;
        rc = bpp == 3 ? ljb_jpeg_encode_rgb_dev(c->ctx[r], biased, w, h, dstride, s.g0, ng, (uint8_t *)c->d_out[r], s.cap, d_offs, d_bits, nullptr, d_res)
                      : ljb_jpeg_encode_rgba_dev(c->ctx[r], biased, w, h, dstride, s.g0, ng, (uint8_t *)c->d_out[r], s.cap, d_offs, d_bits, nullptr, d_res);
        if (rc != 0) return rc;
        COMM_CUDA(cudaMemcpyAsync(c->d_mine[r], d_res, sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->ctx[r]->stream));
    }
    std::vector<const uint64_t *> totals(c->n);
    std::vector<uint64_t> bases(c->n);
    uint64_t grand = 0;
    for (int r = 0; r < c->n; ++r) totals[r] = c->d_mine[r];
    if ((rc = ljb_comm_gather_totals(c, totals.data(), bases.data(), &grand)) != 0) return rc;
    if (out_len) *out_len = (size_t)grand;
    if (grand > out_cap) return LJB_E_CAPACITY;
    std::vector<std::vector<uint64_t>> hres(c->n, std::vector<uint64_t>(3, 0)), hoffs(c->n);
    for (int r = 0; r < c->n; ++r) {
        const Shard &s = sh[r];
        if (s.g1 == s.g0) continue;
        const size_t ng = s.g1 - s.g0;
        COMM_CUDA(cudaSetDevice(r));
        uint64_t *d_res = (uint64_t *)c->d_small[r], *d_offs = d_res + 3;
        uint16_t *d_bits = (uint16_t *)(d_offs + ng + 1);
        COMM_CUDA(cudaMemcpyAsync(hres[r].data(), d_res, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        COMM_CUDA(cudaMemcpyAsync(out + bases[r], c->d_out[r], (size_t)c->h_all[r], cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        if (group_offsets) {
            hoffs[r].resize(ng + 1);
            COMM_CUDA(cudaMemcpyAsync(hoffs[r].data(), d_offs, (ng + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->ctx[r]->stream));
        }
        if (group_bits)
            COMM_CUDA(cudaMemcpyAsync(group_bits + 3 * s.g0, d_bits, ng * 3 * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->ctx[r]->stream));
    }
    for (int r = 0; r < c->n; ++r) {
        COMM_CUDA(cudaSetDevice(r));
        COMM_CUDA(cudaStreamSynchronize(c->ctx[r]->stream));
    }
    if (group_offsets) // a shard's table is relative to its own stream: rebase by the shard's base
        for (int r = 0; r < c->n; ++r)
            for (size_t k = 0; k < hoffs[r].size(); ++k) group_offsets[sh[r].g0 + k] = hoffs[r][k] + bases[r];
    for (int r = 0; r < c->n; ++r) {
        if (hres[r][2] & 2) return LJB_E_UNSUPPORTED;
        if (hres[r][2] & 1) return LJB_E_CAPACITY;
    }
    return LJB_OK;
}

extern "C" int ljb_comm_jpeg_encode_rgba(ljb_comm *c, const uint8_t *rgba, int w, int h, size_t stride, uint8_t *out, size_t out_cap,
                                         uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    return comm_jpeg_encode(c, rgba, 4, w, h, stride, out, out_cap, group_offsets, group_bits, out_len);
}
extern "C" int ljb_comm_jpeg_encode_rgb(ljb_comm *c, const uint8_t *rgb, int w, int h, size_t stride, uint8_t *out, size_t out_cap,
                                        uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len)
{
    return comm_jpeg_encode(c, rgb, 3, w, h, stride, out, out_cap, group_offsets, group_bits, out_len);
}
