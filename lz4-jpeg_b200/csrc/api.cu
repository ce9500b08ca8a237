// lz4-jpeg_b200/csrc/api.cu — context management and error plumbing of the C ABI (include/lz4jpeg_b200.h).
#include "common.cuh"

#include <stdio.h>
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
    ljb_ctx *c = new ljb_ctx();
    memset(c, 0, sizeof *c);
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    LJB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    LJB_CUDA(cudaEventCreate(&c->ev0));
    LJB_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return LJB_OK;
}

extern "C" void ljb_ctx_destroy(ljb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_scratch);
    cudaFree(c->d_status);
    cudaFree(c->d_stage_in);
    cudaFree(c->d_stage_out);
    cudaFree(c->d_small);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void *ljb_ctx_stream(ljb_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t ljb_ctx_launch_count(const ljb_ctx *c) { return c ? c->launches : 0; }

extern "C" float ljb_ctx_last_kernel_ms(const ljb_ctx *cc)
{
    ljb_ctx *c = const_cast<ljb_ctx *>(cc);
    if (!c) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
    c->last_kernel_ms = ms;
    return ms;
}
