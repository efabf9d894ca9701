// Microbenchmark for the "entries in registers, loop over hits" vote loop:
// lane = entry.  A warp holds 32 x E pre-decoded bucket entries (entry word, accumulator row address) in
// registers and loops over the H hits of the bucket piece; a hit word is warp-uniform and comes from shared
// memory with one broadcast LDS.128 per four hits.  No staging, no STS, and the cost per vote does not depend on
// how many hits the bucket has (the grouped loop needs pieces of 32 hits to reach its best rate).
// Compares bank patterns: random rows / bins (what a bucket slice gives) against rows spread over distinct banks.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o entry_regs_vote entry_regs_vote.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 640, S = ROWS + 1, BINS = 31;
constexpr int TILES_PER_WARP = 256;

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ void red_shared(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr)); }

// GUARD: 1 = per-vote guard margin (VIADDMNMX) + one test per hit, 0 = no guard arithmetic at all
// PATTERN: 0 = random entry angle, rows advance every `per_row` entries (bucket slice sorted by m_r);
//          1 = bank-aware order: the 32 lanes' (row - coarse angle) differ mod 32 (conflict-free up to the borrow)
template <int E, int GUARD, int PATTERN>
__global__ void __launch_bounds__(1024) bench(unsigned long long *cycles, unsigned *sink, int H, int per_row) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem);
    uint32_t *hits = reinterpret_cast<uint32_t *>(smem + ((BINS * S * 4 + 15) / 16 * 16)) + (threadIdx.x >> 5) * 512;
    for (int i = threadIdx.x; i < BINS * S; i += blockDim.x) acc[i] = 0;
    const unsigned lane = threadIdx.x & 31;
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u;
    for (int i = lane; i < 512; i += 32) hits[i] = lcg(s) | 0xFFFu;
    __syncthreads();
    const uint32_t acc_base = (uint32_t)__cvta_generic_to_shared(acc);
    uint32_t worst_all = 0, repairs = 0;
    long long t0 = clock64();
    for (int t = 0; t < TILES_PER_WARP; t++) {
        // "load + decode" E entries per lane (stands for the LDG of 32 E consecutive bucket entries + LOP3 / LEA)
        uint32_t e[E], a[E];
        unsigned rowbase = __shfl_sync(0xffffffffu, lcg(s) >> 12, 0);
#pragma unroll
        for (int u = 0; u < E; u++) {
            const unsigned j = u * 32 + lane;
            uint32_t row, th;
            if (PATTERN == 0) { row = (rowbase + j / per_row) % ROWS; th = lcg(s) & 0xFFFFF000u; }
            else {
                // coarse angle bin A_u and row chosen so that (row - A_u) mod 32 == lane
                const unsigned A = (lcg(s) >> 8) % 30u;
                row = (rowbase / 32 * 32 + ((lane + A) & 31u) + 32 * u) % ROWS;
                const uint32_t frac = lcg(s) % (uint32_t)((1ull << 32) / 30ull);
                th = ((uint32_t)(((unsigned long long)A << 32) / 30ull) + frac) & 0xFFFFF000u;
            }
            e[u] = th;
            a[u] = acc_base + row * 4;
        }
        // loop over the H hits of the piece, four hit words per LDS.128
        for (int h = 0; h < H; h += 4) {
            const uint4 hv = *reinterpret_cast<const uint4 *>(hits + ((h + t * 4) & 508));
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (h + q < H) {
                    uint32_t worst = 0;
#pragma unroll
                    for (int u = 0; u < E; u++) {
                        const unsigned long long p = (unsigned long long)(hw[q] - e[u]) * 30ull;
                        if (GUARD) worst = max(worst, (uint32_t)p - 0x56000u);
                        red_shared((uint32_t)(p >> 32) * (S * 4) + a[u]);
                    }
                    if (GUARD && worst >= 0xFFF72000u) repairs++;
                    worst_all |= worst;
                }
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    unsigned x = worst_all + repairs;
    for (int i = threadIdx.x; i < BINS * S; i += blockDim.x) x += acc[i];
    if (x == 0xdeadbeef) sink[0] = x;
}

template <int E, int GUARD, int PATTERN>
void run(const char *name, int H, int per_row) {
    int nsm = 148, threads = 1024;
    unsigned *sink; unsigned long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, nsm * 8);
    size_t smem = (BINS * S * 4 + 15) / 16 * 16 + 32 * 512 * 4;
    cudaFuncSetAttribute(bench<E, GUARD, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<E, GUARD, PATTERN><<<nsm, threads, smem>>>(cyc, sink, H, per_row);
    bench<E, GUARD, PATTERN><<<nsm, threads, smem>>>(cyc, sink, H, per_row);
    cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    double votes = (double)threads * TILES_PER_WARP * E * H;
    printf("%-40s E=%d H=%3d per_row=%2d  %.3f votes/clk/SM   err=%s\n", name, E, H, per_row, votes / avg,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    for (int H : {1, 2, 4, 8, 16, 32, 64, 256}) {
        run<4, 1, 0>("entries in regs, guard, random banks", H, 2);
        run<8, 1, 0>("entries in regs, guard, random banks", H, 2);
        run<8, 0, 0>("entries in regs, no guard, random banks", H, 2);
        run<8, 1, 1>("entries in regs, guard, bank-aware", H, 2);
        run<8, 0, 1>("entries in regs, no guard, bank-aware", H, 2);
    }
    for (int per_row : {1, 4, 32}) run<8, 1, 0>("entries in regs, guard, random banks", 32, per_row);
    return 0;
}
