// lz4-jpeg_b200/csrc/common.cuh — shared device helpers and the context object (sm_100a only).
#pragma once
#ifdef LJB_EMU_BUILD
// tests/emu: the kernel sources compiled for the host under a lock-step emulator of the CUDA builtins (test infrastructure;
// the host-side half of every file is left out of that build)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stddef.h>

#include "../../include/lz4jpeg_b200.h"

#ifndef __CUDA_ARCH__
#define LJB_HOST_ONLY 1
#endif

#ifndef LJB_EMU_BUILD
struct ljb_ctx {
    int device;
    int num_sms;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    uint64_t launches;
    float last_kernel_ms;
    int kernel_ms_summed;    // last_kernel_ms holds the sum over the chunks of a host-buffer call (not re-read from ev0/ev1)
    uint32_t attr_mask;      // kernels whose dynamic shared-memory limit has been raised on THIS device (per context, not per process)
    size_t l2_persist_prev;  // the device's persisting-L2 limit before this context raised it (restored on destroy)
    int l2_persist_set;
    size_t l2_persist_bytes; // L2 set aside for persisting accesses (0 = unsupported / disabled by LJB_NO_L2_PERSIST)
    size_t l2_window_max;    // largest access-policy window the device accepts
    // persistent scratch (grown on demand, never shrunk)
    void *d_scratch;      // per-CTA match records etc.
    size_t scratch_bytes;
    void *d_status;       // ticket counter + decoupled look-back status words
    size_t status_bytes;
    void *d_small;        // offsets / results staging
    size_t small_bytes;
    // host-buffer entry points: chunked H2D -> kernel -> D2H pipeline (two buffers each way, three streams)
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_h2d[2], ev_kern[2], ev_d2h[2], ev_res;
    void *d_pin[2];       // input chunk buffers
    size_t pin_bytes[2];
    void *d_pout[2];      // output chunk buffers
    size_t pout_bytes[2];
    uint64_t *h_res;      // pinned: per-chunk kernel results (3 x u64 each)
    size_t h_res_chunks;
};

// Bytes of input per pipeline chunk of the host-buffer entry points (a multiple of every block / group-row size used).
// (128 MiB; LJB_PIPE_CHUNK_BYTES overrides it so that tests can drive many chunks through small inputs.)
size_t ljb_pipe_chunk(void);
int ljb_pipe_init(ljb_ctx *ctx, size_t nchunks);

// bits of ljb_ctx::attr_mask
enum { LJB_ATTR_LZ4 = 1, LJB_ATTR_JPEG = 2, LJB_ATTR_JFIF = 4, LJB_ATTR_LZ4_LAZY = 8, LJB_ATTR_LZ4_DEC = 16, LJB_ATTR_JPEG_DEC = 32, LJB_ATTR_LZ4_SMALL = 64 };

int ljb_set_cuda_error(cudaError_t e, const char *what, int line);
int ljb_ensure(void **p, size_t *have, size_t want);

#define LJB_CUDA(x)                                                        \
    do {                                                                   \
        cudaError_t e__ = (x);                                             \
        if (e__ != cudaSuccess) return ljb_set_cuda_error(e__, #x, __LINE__); \
    } while (0)
#endif // !LJB_EMU_BUILD

// ---- decoupled look-back (single-pass chained scan of per-block byte counts) -------------------------
// status word: bits 63..62 = state (0 invalid, 1 aggregate published, 2 inclusive prefix published),
// bits 61..0 = value.  One 64-bit word carries flag and value together, so no fence is needed.
#define LJB_ST_AGG (1ull << 62)
#define LJB_ST_INC (2ull << 62)
#define LJB_ST_MASK (3ull << 62)

#if defined(__CUDACC__) || defined(LJB_EMU_BUILD)
#ifdef LJB_EMU_BUILD
static inline uint64_t ljb_ld_volatile(const uint64_t *p) { return *(const volatile uint64_t *)p; }
static inline void ljb_st_volatile(uint64_t *p, uint64_t v) { *(volatile uint64_t *)p = v; }
#else
__device__ __forceinline__ uint64_t ljb_ld_volatile(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ljb_st_volatile(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// Called by one full warp.  Publishes `mine` for unit `idx` and returns the exclusive prefix
// (`lead` + sum of all earlier units).  Units must have been claimed in increasing order by CTAs that
// are resident (ticket counter), so every predecessor is running or finished: no deadlock.
__device__ __forceinline__ uint64_t ljb_lookback(uint64_t *status, long long idx, uint64_t mine, uint64_t lead)
{
    const unsigned lane = threadIdx.x & 31;
    if (idx == 0) {
        if (lane == 0) ljb_st_volatile(&status[0], LJB_ST_INC | (lead + mine));
        return lead;
    }
    if (lane == 0) ljb_st_volatile(&status[idx], LJB_ST_AGG | mine);
    uint64_t excl = 0;
    long long at = idx - 1;
    for (;;) {
        long long j = at - (long long)lane;
        uint64_t w;
        if (j < 0) {
            w = LJB_ST_INC | lead; // virtual unit -1 holds the inclusive prefix `lead`
        } else {
            do {
                w = ljb_ld_volatile(&status[j]);
            } while ((w & LJB_ST_MASK) == 0);
        }
        unsigned inc = __ballot_sync(0xffffffffu, (w & LJB_ST_MASK) == LJB_ST_INC);
        uint64_t v = w & ~LJB_ST_MASK;
        if (inc) {
            int first = __ffs(inc) - 1; // nearest predecessor holding an inclusive prefix
            if ((int)lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (inc) break;
        at -= 32;
    }
    if (lane == 0) ljb_st_volatile(&status[idx], LJB_ST_INC | (excl + mine));
    return excl;
}

// The same protocol in two steps, for producers that publish a unit's size as soon as it is known and fetch its
// offset later (after more work), when the predecessors have long been published: the walk then rarely waits.
__device__ __forceinline__ void ljb_lookback_publish(uint64_t *status, long long idx, uint64_t mine, uint64_t lead)
{
    if ((threadIdx.x & 31) == 0) ljb_st_volatile(&status[idx], idx == 0 ? (LJB_ST_INC | (lead + mine)) : (LJB_ST_AGG | mine));
}
__device__ __forceinline__ uint64_t ljb_lookback_resolve(uint64_t *status, long long idx, uint64_t mine, uint64_t lead)
{
    const unsigned lane = threadIdx.x & 31;
    if (idx == 0) return lead;
    uint64_t excl = 0;
    long long at = idx - 1;
    for (;;) {
        long long j = at - (long long)lane;
        uint64_t w;
        if (j < 0) {
            w = LJB_ST_INC | lead;
        } else {
            do {
                w = ljb_ld_volatile(&status[j]);
            } while ((w & LJB_ST_MASK) == 0);
        }
        unsigned inc = __ballot_sync(0xffffffffu, (w & LJB_ST_MASK) == LJB_ST_INC);
        uint64_t v = w & ~LJB_ST_MASK;
        if (inc) {
            int first = __ffs(inc) - 1;
            if ((int)lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (inc) break;
        at -= 32;
    }
    if (lane == 0) ljb_st_volatile(&status[idx], LJB_ST_INC | (excl + mine));
    return excl;
}
#endif
