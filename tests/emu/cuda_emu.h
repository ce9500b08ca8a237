// tests/emu/cuda_emu.h — TEST INFRASTRUCTURE: a lock-step CPU emulator for the subset of CUDA the LZ4 kernels use.
//
// The build container has no GPU.  To debug kernel LOGIC before spending GPU-minutes, the kernel source is compiled for
// the host with this header standing in for the CUDA builtins: every CUDA thread is a fiber, warp collectives
// (__shfl*_sync, __ballot_sync, __reduce_*_sync, ...) and CTA barriers are rendezvous points between fibers, atomics are
// plain operations (one OS thread runs everything).  One CTA runs at a time, so a persistent kernel is launched with a
// grid of 1 and pulls all its work itself.  This is NOT a product path and not a fallback: nothing under lz4-jpeg_b200/
// includes it; tests/emu builds a separate test binary from the kernel sources (see tests/emu/emu_lz4.cpp).
//
// What it checks: the algorithm as written (indices, barriers that every thread reaches, collectives that every lane
// reaches, shared-memory extents via canaries).  What it cannot check: data races, memory-model issues, performance.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define LJB_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__
#define __align__(x)
#define __restrict__

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

namespace emu {

enum State { RUNNABLE, WAIT_WARP, WAIT_TEAM, WAIT_CTA, WAIT_NAMED, YIELDED, DONE };

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    State state = RUNNABLE;
    unsigned tid = 0;
    unsigned par = 0; // parity of this lane's next warp collective
    unsigned team_mask = 0; // WAIT_TEAM: the lanes of the __syncwarp(mask) this lane waits in
    int cta_pred = 0;
};

struct Warp {
    uint64_t slot[2][32];
    unsigned arrived = 0;
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned nthreads = 0, bid = 0, grid = 1;
    unsigned cta_arrived = 0, cta_or = 0, cta_or_result = 0;
    unsigned named_arrived = 0, named_need = 0;
    std::function<void()> body;
    void *sched_sp = nullptr;
    Fiber *cur = nullptr;
    unsigned long long clock = 0;
};

extern Cta *g;

extern "C" void emu_switch(void **save_sp, void *load_sp);

inline void block(State st)
{
    Fiber *f = g->cur;
    f->state = st;
    emu_switch(&f->sp, g->sched_sp);
}

// post a value, wait for the whole warp, return the 32 posted values
inline const uint64_t *exchange(uint64_t v)
{
    Fiber *f = g->cur;
    Warp &w = g->warps[f->tid >> 5];
    const unsigned par = f->par;
    f->par ^= 1;
    w.slot[par][f->tid & 31] = v;
    ++w.arrived;
    block(WAIT_WARP);
    return w.slot[par];
}

inline unsigned lane_id() { return g->cur->tid & 31; }
inline unsigned warp_lanes() // lanes of the current warp that exist (the last warp of a CTA may be partial)
{
    const unsigned w = g->cur->tid >> 5;
    return std::min(32u, g->nthreads - w * 32);
}

void launch(unsigned grid, unsigned threads, std::function<void()> body);

} // namespace emu

struct EmuIdx {
    unsigned y = 0, z = 0;
};
struct EmuThreadIdx { unsigned get() const { return emu::g->cur->tid; } };
#define threadIdx (EmuDim3{emu::g->cur->tid, 0, 0})
#define blockIdx (EmuDim3{emu::g->bid, 0, 0})
#define blockDim (EmuDim3{emu::g->nthreads, 1, 1})
#define gridDim (EmuDim3{emu::g->grid, 1, 1})
struct EmuDim3 { unsigned x, y, z; };

// ---- math / bit intrinsics
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { return a < b ? a : b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { return a > b ? a : b; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline long long clock64() { return (long long)(emu::g->clock += 7); }

// ---- memory
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }

template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- warp collectives (full masks only)
#define EMU_FULL(mask) assert((mask) == 0xffffffffu)
// (a partial mask is a rendezvous of a team of lanes that takes its own path through the kernel)
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    if (mask == 0xffffffffu) {
        emu::exchange(0);
    } else {
        emu::g->cur->team_mask = mask;
        emu::block(emu::WAIT_TEAM);
    }
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    EMU_FULL(mask);
    const uint64_t *s = emu::exchange(pred ? 1 : 0);
    unsigned r = 0;
    for (unsigned l = 0; l < emu::warp_lanes(); ++l) r |= (unsigned)(s[l] & 1) << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
template <class T> static inline T emu_from(uint64_t v) { T t; memcpy(&t, &v, sizeof t); return t; }
template <class T> static inline uint64_t emu_to(T t) { uint64_t v = 0; memcpy(&v, &t, sizeof t); return v; }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    EMU_FULL(mask);
    const unsigned l = emu::lane_id();
    const uint64_t *s = emu::exchange(emu_to(v));
    const unsigned from = (l & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    return emu_from<T>(s[from]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    EMU_FULL(mask);
    const unsigned l = emu::lane_id();
    const uint64_t *s = emu::exchange(emu_to(v));
    return (l & (unsigned)(width - 1)) >= d ? emu_from<T>(s[l - d]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    EMU_FULL(mask);
    const unsigned l = emu::lane_id();
    const uint64_t *s = emu::exchange(emu_to(v));
    return (l & (unsigned)(width - 1)) + d < (unsigned)width ? emu_from<T>(s[l + d]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
    EMU_FULL(mask);
    const unsigned l = emu::lane_id();
    const uint64_t *s = emu::exchange(emu_to(v));
    return emu_from<T>(s[l ^ (unsigned)x]);
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
    EMU_FULL(mask);
    const uint64_t *s = emu::exchange(v);
    unsigned r = 0;
    for (unsigned l = 0; l < emu::warp_lanes(); ++l) r += (unsigned)s[l];
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    EMU_FULL(mask);
    const uint64_t *s = emu::exchange(v);
    unsigned r = 0;
    for (unsigned l = 0; l < emu::warp_lanes(); ++l) r = std::max(r, (unsigned)s[l]);
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    EMU_FULL(mask);
    const uint64_t *s = emu::exchange(v);
    unsigned r = 0xffffffffu;
    for (unsigned l = 0; l < emu::warp_lanes(); ++l) r = std::min(r, (unsigned)s[l]);
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    EMU_FULL(mask);
    const uint64_t *s = emu::exchange(v);
    unsigned r = 0;
    for (unsigned l = 0; l < emu::warp_lanes(); ++l) r |= (unsigned)s[l];
    return r;
}

// ---- CTA barriers
static inline void __syncthreads()
{
    ++emu::g->cta_arrived;
    emu::block(emu::WAIT_CTA);
}
static inline int __syncthreads_or(int pred)
{
    emu::g->cta_or |= pred ? 1u : 0u;
    ++emu::g->cta_arrived;
    emu::block(emu::WAIT_CTA);
    return (int)emu::g->cta_or_result;
}
// a lane that spins on a flag another warp will set: it gets its next turn after every other warp has had one
static inline void emu_spin_yield() { emu::block(emu::YIELDED); }
static inline void emu_named_barrier(unsigned nthr)
{
    emu::g->named_need = nthr;
    ++emu::g->named_arrived;
    emu::block(emu::WAIT_NAMED);
}
