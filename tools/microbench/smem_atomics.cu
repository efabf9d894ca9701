// Microbenchmark: shared-memory vote-accumulation primitives on sm_100a.
// Measures updates/clk/SM for (a) ATOMS random bank, (b) ATOMS conflict-free,
// (c) LDS+IADD+STS conflict-free (owner-computes), (d) LDS+STS random bank,
// (e) global RED into an L2-resident table. Used to choose the vote kernel design
// ("vote kernel").  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CELLS = 32 * 1024;          // 128 KB of u32 counters
constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }

template <int MODE>
__global__ void __launch_bounds__(1024) bench(unsigned *gtable, unsigned long long *cycles, unsigned *sink) {
    extern __shared__ unsigned acc[];
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u;
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        unsigned r = lcg(s) >> 8;
        if (MODE == 0) {                 // ATOMS, fully random cell
            atomicAdd(&acc[r % CELLS], 1u);
        } else if (MODE == 1) {          // ATOMS, bank == lane (conflict-free)
            atomicAdd(&acc[((r % (CELLS / 32)) * 32) + lane], 1u);
        } else if (MODE == 2) {          // owner RMW, bank == lane, warp-private rows
            unsigned rows = CELLS / 32 / 32;                       // rows per warp
            unsigned idx = ((warp * rows + (r % rows)) * 32) + lane;
            acc[idx] = acc[idx] + 1;
        } else if (MODE == 3) {          // owner RMW, random bank inside warp-private region
            unsigned per = CELLS / 32;
            unsigned idx = warp * per + (r % per);
            acc[idx] = acc[idx] + 1;     // (racy inside a warp on purpose: cost probe only)
        } else if (MODE == 4) {          // global RED, 4 MB table (L2 resident)
            atomicAdd(&gtable[(r + blockIdx.x * 7919u) & (1024 * 1024 - 1)], 1u);
        } else if (MODE == 5) {          // ATOMS u16-packed: add 1<<16 or 1
            atomicAdd(&acc[r % CELLS], (r & 0x10000) ? 65536u : 1u);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    unsigned x = 0;
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x) x += acc[i];
    if (x == 0xdeadbeef) sink[0] = x;
}

template <int MODE>
void run(const char *name, int threads) {
    int nsm = 148;
    unsigned *gtable, *sink; unsigned long long *cyc;
    cudaMalloc(&gtable, 4 << 20); cudaMemset(gtable, 0, 4 << 20);
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, nsm * 8);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CELLS * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<MODE><<<nsm, threads, CELLS * 4>>>(gtable, cyc, sink);
    cudaEventRecord(a);
    bench<MODE><<<nsm, threads, CELLS * 4>>>(gtable, cyc, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[148]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    double ops = (double)threads * ITERS;
    printf("%-34s threads=%4d  %.3f updates/clk/SM  chip %.3e updates/s (%.3f ms)  err=%s\n", name, threads,
           ops / avg, ops * nsm / (ms * 1e-3), ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(gtable); cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int threads : {256, 1024}) {
        run<0>("ATOMS random cell", threads);
        run<1>("ATOMS bank==lane", threads);
        run<2>("LDS+STS owner, bank==lane", threads);
        run<3>("LDS+STS owner, random bank", threads);
        run<4>("global RED, 4MB table", threads);
        run<5>("ATOMS random, packed u16", threads);
    }
    return 0;
}
