// lz4-jpeg_b200/csrc/api.cu — context management and error plumbing of the C ABI (include/lz4jpeg_b200.h).
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_cuda_err[512] = "";

int ljb_set_cuda_error(cudaError_t e, const char *what, int line)
{
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s (%s) at line %d: %s", cudaGetErrorName(e), cudaGetErrorString(e), line, what);
    return LJB_E_CUDA;
}

int ljb_ensure(void **p, size_t *have, size_t want)
{
    if (*have >= want && *p) return LJB_OK;
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    size_t sz = want + (want >> 3) + 4096;
    cudaError_t e = cudaMalloc(p, sz);
    if (e != cudaSuccess) return ljb_set_cuda_error(e, "cudaMalloc(scratch)", __LINE__);
    *have = sz;
    return LJB_OK;
}

size_t ljb_pipe_chunk(void)
{
    const char *e = getenv("LJB_PIPE_CHUNK_BYTES");
    if (e) {
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 1) return (size_t)v;
    }
    return (size_t)128 << 20;
}

// pinned host array for the per-chunk kernel results of the pipelined host-buffer entry points
int ljb_pipe_init(ljb_ctx *ctx, size_t nchunks)
{
    if (ctx->h_res_chunks >= nchunks && ctx->h_res) return LJB_OK;
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    ctx->h_res = nullptr;
    ctx->h_res_chunks = 0;
    const size_t want = nchunks + 16;
    cudaError_t e = cudaMallocHost((void **)&ctx->h_res, want * 3 * sizeof(uint64_t));
    if (e != cudaSuccess) return ljb_set_cuda_error(e, "cudaMallocHost(results)", __LINE__);
    ctx->h_res_chunks = want;
    return LJB_OK;
}

extern "C" const char *ljb_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char *ljb_strerror(int code)
{
    switch (code) {
    case LJB_OK: return "ok";
    case LJB_E_ARG: return "invalid argument";
    case LJB_E_CUDA: return "CUDA error (see ljb_last_cuda_error)";
    case LJB_E_CAPACITY: return "output buffer too small";
    case LJB_E_FORMAT: return "inconsistent or ambiguous stream";
    case LJB_E_UNSUPPORTED: return "input outside the reference's defined behaviour";
    default: return "unknown error";
    }
}

extern "C" int ljb_ctx_create(int device, ljb_ctx **out)
{
    if (!out) return LJB_E_ARG;
    *out = nullptr;
    int count = 0;
    LJB_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return LJB_E_ARG;
    LJB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LJB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { // this library carries sm_100a code only: no fallback path
        snprintf(g_cuda_err, sizeof g_cuda_err, "device %d is sm_%d%d; liblz4jpeg_b200 is built for sm_100a only", device,
                 prop.major, prop.minor);
        return LJB_E_CUDA;
    }
    ljb_ctx *c = new ljb_ctx(); // value-initialised: every member starts at zero / null
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    // a failure below releases whatever has been created so far (ljb_ctx_destroy tolerates null members)
#define CTX_TRY(x)                                                    \
    do {                                                              \
        cudaError_t e__ = (x);                                        \
        if (e__ != cudaSuccess) {                                     \
            const int rc__ = ljb_set_cuda_error(e__, #x, __LINE__);   \
            ljb_ctx_destroy(c);                                       \
            return rc__;                                              \
        }                                                             \
    } while (0)
    CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (!getenv("LJB_NO_L2_PERSIST") && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        // Device-wide setting (other users of the device in this process see it too): the previous limit is kept and put
        // back by ljb_ctx_destroy.  LJB_NO_L2_PERSIST=1 leaves the device untouched.
        size_t want = (size_t)64 << 20; // room for the 38 MB of LZ4 match records of 148 CTAs
        if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
        size_t prev = 0;
        if (cudaDeviceGetLimit(&prev, cudaLimitPersistingL2CacheSize) != cudaSuccess) {
            cudaGetLastError();
            prev = 0;
        }
        if (prev >= want) {
            c->l2_persist_bytes = prev;
            c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        } else if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            c->l2_persist_bytes = want;
            c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
            c->l2_persist_prev = prev;
            c->l2_persist_set = 1;
        } else {
            cudaGetLastError();
        }
    }
    CTX_TRY(cudaEventCreate(&c->ev0));
    CTX_TRY(cudaEventCreate(&c->ev1));
    CTX_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CTX_TRY(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&c->ev_kern[i], cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
    }
    CTX_TRY(cudaEventCreateWithFlags(&c->ev_res, cudaEventDisableTiming));
#undef CTX_TRY
    *out = c;
    return LJB_OK;
}

extern "C" void ljb_ctx_destroy(ljb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s_in) cudaStreamSynchronize(c->s_in);
    if (c->s_out) cudaStreamSynchronize(c->s_out);
    cudaFree(c->d_scratch);
    cudaFree(c->d_status);
    cudaFree(c->d_small);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_pin[i]);
        cudaFree(c->d_pout[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_kern[i]) cudaEventDestroy(c->ev_kern[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    if (c->ev_res) cudaEventDestroy(c->ev_res);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->l2_persist_set) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_prev);
    cudaGetLastError();
    delete c;
}

extern "C" void *ljb_ctx_stream(ljb_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t ljb_ctx_launch_count(const ljb_ctx *c) { return c ? c->launches : 0; }

extern "C" float ljb_ctx_last_kernel_ms(const ljb_ctx *cc)
{
    ljb_ctx *c = const_cast<ljb_ctx *>(cc);
    if (!c) return -1.f;
    if (c->kernel_ms_summed) return c->last_kernel_ms; // a host-buffer call: the sum over its chunks' kernels
    float ms = -1.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
    c->last_kernel_ms = ms;
    return ms;
}
