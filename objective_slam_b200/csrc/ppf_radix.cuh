// ppf_radix.cuh -- stable LSD radix sort of (u32 key, u32 value) records over the LOW `bits` bits of the key.
// The model table build sorts the N^2 (bucket rank, pair index) records with it (ppf_model.cu; the reference sorts
// (32-bit FNV key, pair index) with thrust::sort_by_key, model.cu:53-60): a 10k-point model has 6.8k buckets = 13 rank
// bits = 2 passes of 7 / 6 bits over 8-byte records.
//
// One pass = three launches:
//   radix_hist_kernel     per tile of 4096 records: digit histogram (shared-memory atomics) -> tilehist[digit][tile]
//   scan (3 small kernels) exclusive prefix over tilehist in (digit, tile) order = where the tile's records of a digit go
//   radix_scatter_kernel  per tile: every warp ranks the records of its slice among those of the same digit, in index
//                          order (the lanes of a digit are grouped by ballots in the first pass, by match.any later;
//                          the group's leader advances the warp's count), a block scan turns the counts into offsets, the records go to a shared-memory
//                          staging area sorted by digit and leave it in coalesced runs.  Stable: a warp owns a
//                          contiguous slice of the tile; warps and tiles are ordered by the prefix sums.
// HBM traffic per pass and record: 4 B (histogram) + 8 B read + 8 B written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

namespace ppf {

// The histogram kernel of pass 0 may translate the keys through a table, in place (the model build sorts by bucket rank
// but has cell codes in memory: rank = lut[code], code 0xFFFFFFFF and table value 0xFFFFFFFF -> rank 0), and pass 0
// may take the record's index as its value (vals_in == nullptr): no separate rank pass, no iota array.
__device__ __forceinline__ uint32_t rx_key(const uint32_t *__restrict__ keys, size_t p, const uint32_t *__restrict__ lut) {
    const uint32_t c = keys[p];
    if (!lut) return c;
    const uint32_t r = (c == 0xFFFFFFFFu) ? 0u : __ldg(lut + c);
    return (r == 0xFFFFFFFFu) ? 0u : r;
}

constexpr int kRxThreads = 512;
constexpr int kRxItems   = 8;
constexpr int kRxTile    = kRxThreads * kRxItems;      // 4096 records per CTA
constexpr int kRxMaxBins = 256;
constexpr int kRxWarps   = kRxThreads / 32;

// translate != nullptr: the keys are read through the table and WRITTEN BACK translated (the first pass of the model
// build turns cell codes into bucket ranks on the way: no separate rank pass)
__global__ void __launch_bounds__(kRxThreads) radix_hist_kernel(uint32_t *keys, const uint32_t *__restrict__ translate,
                                                                size_t n, int shift, uint32_t mask, uint32_t ntiles,
                                                                uint32_t *tilehist) {
    __shared__ uint32_t hist[kRxMaxBins];
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i <= (int)mask; i += kRxThreads) hist[i] = 0;
        __syncthreads();
        const size_t base = (size_t)tile * kRxTile;
#pragma unroll
        for (int j = 0; j < kRxItems; j++) {
            const size_t p = base + (size_t)j * kRxThreads + threadIdx.x;
            if (p < n) {
                const uint32_t k = rx_key(keys, p, translate);
                if (translate) keys[p] = k;
                atomicAdd(&hist[(k >> shift) & mask], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i <= (int)mask; i += kRxThreads) tilehist[(size_t)i * ntiles + tile] = hist[i];
        __syncthreads();
    }
}

// ---- exclusive scan of a u32 array (two levels: 4096-element blocks) ----------------------------------------
constexpr int kScanThreads = 256, kScanItems = 16, kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *total, uint32_t *warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t x = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t i2 = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) i2 += u;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = i2 - x;
        if (lane == 31) *total = i2;
    }
    __syncthreads();
    return warp_sums[warp] + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const uint32_t *__restrict__ a, size_t n, uint32_t *bsum) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) if (base + j < n) s += a[base + j];
    block_exclusive_scan_256(s, &total, ws);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
// one CTA: exclusive scan of the block sums (any count: sequential over 256-element rounds)
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(uint32_t *bsum, uint32_t nb) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += kScanThreads) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? bsum[i] : 0u;
        const uint32_t ex = block_exclusive_scan_256(v, &total, ws);
        if (i < nb) bsum[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *a, size_t n, const uint32_t *__restrict__ bsum) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t total;
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems], s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { v[j] = base + j < n ? a[base + j] : 0u; s += v[j]; }
    uint32_t run = bsum[blockIdx.x] + block_exclusive_scan_256(s, &total, ws);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { if (base + j < n) a[base + j] = run; run += v[j]; }
}

// ---- scatter ---------------------------------------------------------------------------------------------------
template <bool USE_MATCH>
__global__ void __launch_bounds__(kRxThreads, 2) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                   const uint32_t *__restrict__ lut,
                                                                   const uint32_t *__restrict__ vals_in, size_t n, int shift,
                                                                   uint32_t mask, uint32_t ntiles,
                                                                   const uint32_t *__restrict__ tileoff,
                                                                   uint32_t *keys_out, uint32_t *vals_out) {
    extern __shared__ uint32_t rx_smem[];
    const uint32_t nb = mask + 2u;                                  // + one bin for the slots past the end of the array
    uint32_t *cursor = rx_smem;                                     // [kRxWarps][nb] counts, then cursors
    uint32_t *tstart = cursor + kRxWarps * nb;                      // [nb] first staged slot of a digit
    uint32_t *gbase = tstart + nb;                                  // [nb] global position of the tile's first record of a digit
    uint32_t *skey = gbase + nb;                                    // [kRxTile]
    uint32_t *sval = skey + kRxTile;                                // [kRxTile]
    __shared__ uint32_t ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int nbits = 32 - __clz((int)mask);                        // digit bits of this pass (mask = 2^nbits - 1)
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = (size_t)tile * kRxTile + (size_t)warp * (kRxItems * 32) + lane;      // warp-blocked slices
        uint32_t k[kRxItems], v[kRxItems], lpos[kRxItems];
        for (uint32_t i = threadIdx.x; i < kRxWarps * nb; i += kRxThreads) cursor[i] = 0;
#pragma unroll
        for (int j = 0; j < kRxItems; j++) {
            const size_t p = base + (size_t)j * 32;
            k[j] = p < n ? rx_key(keys_in, p, lut) : 0u;
            v[j] = p < n ? (vals_in ? vals_in[p] : (uint32_t)p) : 0u;
        }
        __syncthreads();
        // 1: every warp ranks its records among the records of the same digit in its slice, in index order:
        //    match.any groups the lanes of a digit, the group's leader advances the warp's count
        uint32_t *mine = cursor + warp * nb;
#pragma unroll
        for (int j = 0; j < kRxItems; j++) {
            const size_t p = base + (size_t)j * 32;
            const uint32_t d = p < n ? ((k[j] >> shift) & mask) : mask + 1u;
            // Lanes with the same digit.  match.any takes time proportional to the number of DISTINCT values in the warp:
            // fine once the records arrive sorted by the lower digits (few distinct higher digits per warp: 527 us for
            // the second pass of a 10k-point model against 726 us with ballots), slow on the first pass, where the
            // records come in pair order (~28 distinct digits per warp: 949 us against 781 us) -- there the lanes are
            // grouped by one ballot per digit bit (+ the bit of the past-the-end bin), a fixed cost.
            unsigned peers = 0xffffffffu;
            if constexpr (USE_MATCH) {
                peers = __match_any_sync(0xffffffffu, d);
            } else {
                for (int b = 0; b <= nbits; b++) {
                    const bool bit = (d >> b) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? bal : ~bal;
                }
            }
            const int leader = __ffs((int)peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = mine[d]; mine[d] = old + (uint32_t)__popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            lpos[j] = old + (uint32_t)__popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        // 2: counts -> cursors.  Thread d < nb owns digit d: prefix over the warps, then over the digits.
        uint32_t tot = 0;
        if (threadIdx.x < nb) {
            for (int w = 0; w < kRxWarps; w++) { const uint32_t c = cursor[w * nb + threadIdx.x]; cursor[w * nb + threadIdx.x] = tot; tot += c; }
        }
        {
            // exclusive scan of tot over the 512 threads (zeros beyond nb)
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) ws[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t x = lane < kRxWarps ? ws[lane] : 0u;
                uint32_t i2 = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, i2, o);
                    if (lane >= o) i2 += u;
                }
                if (lane < kRxWarps) ws[lane] = i2 - x;
            }
            __syncthreads();
            const uint32_t ex = ws[warp] + incl - tot;
            if (threadIdx.x < nb) {
                tstart[threadIdx.x] = ex;
                if (threadIdx.x <= mask) gbase[threadIdx.x] = tileoff[(size_t)threadIdx.x * ntiles + tile];
                for (int w = 0; w < kRxWarps; w++) cursor[w * nb + threadIdx.x] += ex;
            }
        }
        __syncthreads();
        // 3: into the staging area, sorted by digit (stable: warps and tiles are ordered by the prefix sums)
#pragma unroll
        for (int j = 0; j < kRxItems; j++) {
            const size_t p = base + (size_t)j * 32;
            const uint32_t d = p < n ? ((k[j] >> shift) & mask) : mask + 1u;
            const uint32_t slot = mine[d] + lpos[j];
            skey[slot] = k[j]; sval[slot] = v[j];
        }
        __syncthreads();
        // 4: coalesced runs out of the staging area
        const size_t tile_base = (size_t)tile * kRxTile;
        const uint32_t nvalid = (uint32_t)((n - tile_base) < (size_t)kRxTile ? (n - tile_base) : (size_t)kRxTile);
        for (uint32_t s = threadIdx.x; s < nvalid; s += kRxThreads) {
            const uint32_t key = skey[s];
            const uint32_t d = (key >> shift) & mask;
            const size_t g = (size_t)gbase[d] + (s - tstart[d]);
            keys_out[g] = key; vals_out[g] = sval[s];
        }
        __syncthreads();
    }
}

struct RadixPlan {
    int passes = 0;
    int pass_bits[4] = {0, 0, 0, 0};
    uint32_t ntiles = 0;
    size_t scratch_words = 0;            // tilehist + block sums
};
inline RadixPlan radix_plan(size_t n, int bits) {
    RadixPlan pl;
    if (bits < 1) bits = 1;
    pl.passes = (bits + 7) / 8;
    const int per = (bits + pl.passes - 1) / pl.passes;
    int left = bits;
    for (int i = 0; i < pl.passes; i++) { pl.pass_bits[i] = left < per ? left : per; left -= pl.pass_bits[i]; }
    if (const char *e = getenv("PPF_B200_RADIX_FIRST")) {          // experiment hook: bits of pass 0 (two-pass plans only)
        const int f = atoi(e);
        if (pl.passes == 2 && f >= 1 && f <= 8 && bits - f >= 1 && bits - f <= 8) { pl.pass_bits[0] = f; pl.pass_bits[1] = bits - f; }
    }
    pl.ntiles = (uint32_t)((n + kRxTile - 1) / kRxTile);
    const size_t hist = (size_t)kRxMaxBins * pl.ntiles;
    pl.scratch_words = hist + (hist + kScanBlock - 1) / kScanBlock + 16;
    return pl;
}

// Sorts (keys[0], vals[0]) by the low `bits` key bits; the buffers ping-pong, the result lands in
// keys[passes & 1] / vals[passes & 1].  `scratch` holds radix_plan(n, bits).scratch_words u32.  Returns the launches.
inline int radix_sort_pairs(uint32_t *keys[2], uint32_t *vals[2], size_t n, const RadixPlan &pl, uint32_t *scratch,
                            cudaStream_t stream, const uint32_t *lut = nullptr, bool index_values = false) {
    int launches = 0, shift = 0, cur = 0;
    uint32_t *tilehist = scratch;
    for (int pass = 0; pass < pl.passes; pass++) {
        const uint32_t mask = (1u << pl.pass_bits[pass]) - 1u;
        const size_t hist_n = (size_t)(mask + 1u) * pl.ntiles;
        uint32_t *bsum = scratch + (size_t)kRxMaxBins * pl.ntiles;
        const uint32_t nblk = (uint32_t)((hist_n + kScanBlock - 1) / kScanBlock);
        const unsigned grid = pl.ntiles < 148u * 8u ? pl.ntiles : 148u * 8u;
        const uint32_t *lut0 = nullptr;                  // (the histogram kernel of pass 0 has translated the keys in place)
        const uint32_t *vin = (pass == 0 && index_values) ? nullptr : vals[cur];
        radix_hist_kernel<<<grid, kRxThreads, 0, stream>>>(keys[cur], pass == 0 ? lut : nullptr, n, shift, mask, pl.ntiles, tilehist);
        scan_block_sums_kernel<<<nblk, kScanThreads, 0, stream>>>(tilehist, hist_n, bsum);
        scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(bsum, nblk);
        scan_apply_kernel<<<nblk, kScanThreads, 0, stream>>>(tilehist, hist_n, bsum);
        const size_t smem = ((size_t)(kRxWarps + 2) * (mask + 2u) + 2u * kRxTile) * sizeof(uint32_t);
        if (pass == 0) {
            cudaFuncSetAttribute(radix_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            radix_scatter_kernel<false><<<grid, kRxThreads, smem, stream>>>(keys[cur], lut0, vin, n, shift, mask, pl.ntiles,
                                                                            tilehist, keys[cur ^ 1], vals[cur ^ 1]);
        } else {
            cudaFuncSetAttribute(radix_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            radix_scatter_kernel<true><<<grid, kRxThreads, smem, stream>>>(keys[cur], lut0, vin, n, shift, mask, pl.ntiles,
                                                                           tilehist, keys[cur ^ 1], vals[cur ^ 1]);
        }
        launches += 5;
        shift += pl.pass_bits[pass];
        cur ^= 1;
    }
    return launches;
}

}  // namespace ppf
