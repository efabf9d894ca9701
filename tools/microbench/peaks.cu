// libppf_peaks.so -- the measured ceiling bench.py quotes the vote kernel against: shared-memory atomic
// increments per clock per SM (ATOMS.POPC.INC, 1024 threads per SM, every lane its own bank = what the grouped vote
// loop issues; and fully random cells = what the one-hit-per-pass loop issues).  Same loop as smem_atomics.cu, exposed
// as a C function so that the roofline denominator is measured in the run that reports it.  MEASUREMENT
// INFRASTRUCTURE, not product code: nothing under objective_slam_b200/ loads it.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libppf_peaks.so peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CELLS = 32 * 1024;          // 128 KB of u32 counters
constexpr int ITERS = 8192;

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }

template <int MODE>
__global__ void __launch_bounds__(1024) atoms_kernel(unsigned long long *cycles, unsigned *sink) {
    extern __shared__ unsigned acc[];
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u;
    const unsigned lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        const unsigned r = lcg(s) >> 8;
        if (MODE == 0) atomicAdd(&acc[r % CELLS], 1u);                               // random cell
        else atomicAdd(&acc[((r % (CELLS / 32)) * 32) + lane], 1u);                  // bank == lane: conflict-free
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    unsigned x = 0;
    for (int i = threadIdx.x; i < CELLS; i += blockDim.x) x += acc[i];
    if (x == 0xdeadbeef) sink[0] = x;
}

template <int MODE>
static int run(double *per_clk_per_sm) {
    int nsm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    unsigned *sink = nullptr; unsigned long long *cyc = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess || cudaMalloc(&cyc, nsm * 8) != cudaSuccess) return 1;
    cudaFuncSetAttribute(atoms_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CELLS * 4);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        atoms_kernel<MODE><<<nsm, 1024, CELLS * 4>>>(cyc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(sink); cudaFree(cyc); return 2; }
        unsigned long long h[1024];
        cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; i++) avg += (double)h[i];
        avg /= nsm;
        const double rate = 1024.0 * ITERS / avg;
        if (rate > best) best = rate;
    }
    cudaFree(sink); cudaFree(cyc);
    *per_clk_per_sm = best;
    return 0;
}

// mode 0: random cells, mode 1: conflict-free (bank == lane).  Returns 0 on success.
extern "C" int peak_smem_atomics(int mode, double *per_clk_per_sm) {
    return mode == 0 ? run<0>(per_clk_per_sm) : run<1>(per_clk_per_sm);
}
