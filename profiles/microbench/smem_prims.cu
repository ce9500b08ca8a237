// profiles/microbench/smem_prims.cu — throughput of the shared-memory primitives the LZ4 index build
// could be made of, measured on one B200 (cycles per warp-instruction at full-SM occupancy).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_prims smem_prims.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int THREADS = 1024;
constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode 0: atomicAdd (no return use) random words in 16 KB     mode 1: atomicAdd with returned value used
// mode 2: same address per warp (worst conflict)                mode 3: __match_any_sync on 8-bit keys
// mode 4: 8x ballot emulation of match_any                      mode 5: __reduce_max_sync
// mode 6: LDS.32 random                                         mode 7: unaligned 64-bit read = 3x LDS.32 + 2 funnelshift
// mode 8: plain STS random u16                                  mode 9: LDS.U16 random
// mode 10: atomicAdd on text-like skewed distribution (zipf-ish) mode 11: atomicMin random
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(uint32_t* out, long long* cycles)
{
    extern __shared__ uint32_t sm[];
    const int NW = 48 * 1024; // words: 192 KB
    for (int i = threadIdx.x; i < NW; i += THREADS) sm[i] = mix(i);
    __syncthreads();
    uint32_t x = mix(threadIdx.x + blockIdx.x * 7919u), acc = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        x = x * 1664525u + 1013904223u;
        uint32_t r = x >> 8;
        if (MODE == 0) { atomicAdd(&sm[r & 4095], 1u); }
        else if (MODE == 1) { acc += atomicAdd(&sm[r & 4095], 1u); }
        else if (MODE == 2) { atomicAdd(&sm[(threadIdx.x >> 5)], 1u); }
        else if (MODE == 3) { acc += __match_any_sync(0xffffffffu, r & 255); }
        else if (MODE == 4) {
            uint32_t m = 0xffffffffu, d = r & 255;
#pragma unroll
            for (int b = 0; b < 8; ++b) { uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1); m &= ((d >> b) & 1) ? bal : ~bal; }
            acc += m;
        }
        else if (MODE == 5) { acc += __reduce_max_sync(0xffffffffu, r); }
        else if (MODE == 6) { acc += sm[r % NW]; }
        else if (MODE == 7) {
            uint32_t a = r % (NW * 4 - 16); uint32_t w = a >> 2, s = (a & 3) * 8;
            uint32_t w0 = sm[w], w1 = sm[w + 1], w2 = sm[w + 2];
            acc += __funnelshift_r(w0, w1, s) ^ __funnelshift_r(w1, w2, s);
        }
        else if (MODE == 8) { ((uint16_t*)sm)[r % (NW * 2)] = (uint16_t)r; }
        else if (MODE == 9) { acc += ((uint16_t*)sm)[r % (NW * 2)]; }
        else if (MODE == 10) { uint32_t z = r & 4095; z = (z * z) >> 12; z = (z * z) >> 12; atomicAdd(&sm[z], 1u); }
        else if (MODE == 11) { atomicMin(&sm[r & 16383], r); }
    }
    long long t1 = clock64();
    out[blockIdx.x * THREADS + threadIdx.x] = acc + sm[threadIdx.x];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, uint32_t* d_out, long long* d_cyc)
{
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
    k<MODE><<<148, THREADS, 192 * 1024>>>(d_out, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, THREADS, 192 * 1024>>>(d_out, d_cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[148]; CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; ++i) avg += cyc[i]; avg /= 148;
    double warp_instr = (double)ITERS * (THREADS / 32);
    printf("%-44s %8.3f ms  %10.0f cyc/CTA  %7.2f cyc per warp-op (SM-wide)  %6.3f cyc per lane-op\n", name, ms, avg, avg / warp_instr, avg / warp_instr / 32);
    return 0;
}

int main()
{
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 148 * THREADS * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
    run<0>("atomicAdd smem random 4096 words (no ret)", d_out, d_cyc);
    run<1>("atomicAdd smem random 4096 words (ret used)", d_out, d_cyc);
    run<2>("atomicAdd smem same addr per warp", d_out, d_cyc);
    run<10>("atomicAdd smem skewed", d_out, d_cyc);
    run<11>("atomicMin smem random 16384 words", d_out, d_cyc);
    run<3>("__match_any_sync 8-bit keys", d_out, d_cyc);
    run<4>("8x ballot match emulation", d_out, d_cyc);
    run<5>("__reduce_max_sync", d_out, d_cyc);
    run<6>("LDS.32 random", d_out, d_cyc);
    run<7>("unaligned 64-bit read (3 LDS + 2 SHF)", d_out, d_cyc);
    run<8>("STS.U16 random", d_out, d_cyc);
    run<9>("LDS.U16 random", d_out, d_cyc);
    return 0;
}
