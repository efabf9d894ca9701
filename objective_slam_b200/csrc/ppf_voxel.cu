// ppf_voxel.cu -- GPU voxel-grid downsample: the step right before the hot path
// (voxelGridDownsample, alignment.cpp:79-87, called for every scene with scene_leaf_size and for every
// model with its own d_dist, alignment.cpp:265-288; stand-alone tool pcl/voxel_grid/voxel_grid.cpp:17-21).
//
// Semantics restated from pcl::VoxelGrid<PointNormal>::applyFilter (PCL 1.7, not vendored in the reference
// and not installed here -- PARITY UNPINNED):
//   min_b = floor(min_xyz / leaf), div_b = floor(max_xyz / leaf) - min_b + 1,
//   cell(p) = (floor(p/leaf) - min_b) . (1, div_b.x, div_b.x*div_b.y),
//   one output point per occupied cell, in ascending cell order, = the centroid of ALL fields of the
//   cell's points: position AND normal are averaged, the normal is NOT re-normalised (this is why the
//   hot path never assumes unit normals).  Non-finite points are dropped.
// Implementation: min/max reduction -> cell ids -> radix sort (cell id, point index; stable, so the sum order
// inside a cell is ascending point index and the result is deterministic) -> segment heads -> one thread per
// cell averages its points.  Same sort + segment-offset machinery as the model table build.
#include <cub/cub.cuh>
#include <algorithm>
#include <cfloat>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

namespace ppf {

__device__ __forceinline__ void atomic_min_float(float *addr, float v) {      // valid for any sign
    if (v >= 0) atomicMin((int *)addr, __float_as_int(v));
    else atomicMax((unsigned int *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned int *)addr, __float_as_uint(v));
}

__global__ void vg_minmax_kernel(const float *xyz, int xs, int n, float *mm /* min xyz, max xyz */) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float p[3] = {xyz[(size_t)i * xs], xyz[(size_t)i * xs + 1], xyz[(size_t)i * xs + 2]};
        if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
        for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], p[c]); hi[c] = fmaxf(hi[c], p[c]); }
    }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_float(mm + c, lo[c]); atomic_max_float(mm + 3 + c, hi[c]); }
    }
}

__global__ void vg_cell_kernel(const float *xyz, int xs, int n, float inv_leaf, int3 min_b, int3 div_b,
                               unsigned long long *cell, uint32_t *idx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float x = xyz[(size_t)i * xs], y = xyz[(size_t)i * xs + 1], z = xyz[(size_t)i * xs + 2];
        unsigned long long c = ~0ull;                                  // non-finite points sort last and are dropped
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            long long ix = (long long)floorf(x * inv_leaf) - min_b.x;
            long long iy = (long long)floorf(y * inv_leaf) - min_b.y;
            long long iz = (long long)floorf(z * inv_leaf) - min_b.z;
            c = (unsigned long long)(ix + iy * (long long)div_b.x + iz * (long long)div_b.x * (long long)div_b.y);
        }
        cell[i] = c;
        idx[i] = (uint32_t)i;
    }
}

__global__ void vg_heads_kernel(const unsigned long long *cell_sorted, int n, uint32_t *heads, uint32_t *n_heads) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long c = cell_sorted[i];
        if (c != ~0ull && (i == 0 || cell_sorted[i - 1] != c)) heads[atomicAdd(n_heads, 1u)] = (uint32_t)i;
    }
}

__global__ void vg_centroid_kernel(const float *xyz, int xs, const float *nrm, int ns, const unsigned long long *cell_sorted,
                                   const uint32_t *idx_sorted, const uint32_t *heads_sorted, int n_cells, int n,
                                   float *out_xyz, float *out_nrm) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_cells; k += gridDim.x * blockDim.x) {
        uint32_t b = heads_sorted[k];
        unsigned long long c = cell_sorted[b];
        float s[6] = {0, 0, 0, 0, 0, 0};
        uint32_t cnt = 0;
        for (uint32_t p = b; p < (uint32_t)n && cell_sorted[p] == c; p++, cnt++) {
            uint32_t i = idx_sorted[p];
            s[0] += xyz[(size_t)i * xs]; s[1] += xyz[(size_t)i * xs + 1]; s[2] += xyz[(size_t)i * xs + 2];
            s[3] += nrm[(size_t)i * ns]; s[4] += nrm[(size_t)i * ns + 1]; s[5] += nrm[(size_t)i * ns + 2];
        }
        float inv = 1.0f / (float)cnt;
        for (int j = 0; j < 3; j++) { out_xyz[3 * (size_t)k + j] = s[j] * inv; out_nrm[3 * (size_t)k + j] = s[3 + j] * inv; }
    }
}

int voxel_grid_run(const float *xyz, int xs, const float *nrm, int ns, int n, int mem, float leaf, float *out_xyz,
                   float *out_nrm, int *n_out) {
    if (!xyz || !nrm || !n_out || n < 0 || xs < 3 || ns < 3 || !(leaf > 0.f)) {
        set_last_error("voxel_grid: NULL pointer, negative size, stride < 3 or leaf <= 0");
        return PPF_ERR_INVALID;
    }
    *n_out = 0;
    if (n == 0) return PPF_OK;
    Workspace ws;
    struct Release { Workspace &w; ~Release() { w.release(); } } rel{ws};
    size_t bx = ((size_t)(n - 1) * xs + 3) * 4, bn = ((size_t)(n - 1) * ns + 3) * 4;
    size_t sort_tmp = 0, sort2_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, n);
    cub::DeviceRadixSort::SortKeys(nullptr, sort2_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, n);
    int rc = ws.reserve((mem == PPF_MEM_HOST ? bx + bn : 0) + (size_t)n * (8 + 8 + 4 + 4 + 4 + 4 + 24) + std::max(sort_tmp, sort2_tmp) + 64);
    if (rc) return rc;
    const float *dx = xyz, *dn = nrm;
    if (mem == PPF_MEM_HOST) {
        float *tx = (float *)ws.take_bytes(bx), *tn = (float *)ws.take_bytes(bn);
        PPF_CUDA_TRY(cudaMemcpyAsync(tx, xyz, bx, cudaMemcpyHostToDevice, cur_stream()));
        PPF_CUDA_TRY(cudaMemcpyAsync(tn, nrm, bn, cudaMemcpyHostToDevice, cur_stream()));
        dx = tx; dn = tn;
    }
    float *mm = ws.take<float>(6);
    unsigned long long *cell = ws.take<unsigned long long>(n), *cell_s = ws.take<unsigned long long>(n);
    uint32_t *idx = ws.take<uint32_t>(n), *idx_s = ws.take<uint32_t>(n), *heads = ws.take<uint32_t>(n), *heads_s = ws.take<uint32_t>(n);
    uint32_t *d_nh = ws.take<uint32_t>(1);
    float *oxyz = ws.take<float>((size_t)3 * n), *onrm = ws.take<float>((size_t)3 * n);
    void *tmp = ws.take_bytes(std::max(sort_tmp, sort2_tmp));
    if (!mm || !cell || !cell_s || !idx || !idx_s || !heads || !heads_s || !d_nh || !oxyz || !onrm || !tmp) {
        set_last_error("voxel_grid: scratch arena too small");
        return PPF_ERR_CUDA;
    }
    const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    PPF_CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, cur_stream()));
    PPF_CUDA_TRY(cudaMemsetAsync(d_nh, 0, 4, cur_stream()));
    int grid = std::min((n + 255) / 256, 148 * 8);
    vg_minmax_kernel<<<grid, 256, 0, cur_stream()>>>(dx, xs, n, mm);
    count_launch();
    float h[6];
    PPF_CUDA_TRY(memcpy_sync(h, mm, sizeof(h), cudaMemcpyDeviceToHost));
    if (!(h[0] <= h[3])) return PPF_OK;                                 // no finite point
    const float inv_leaf = 1.0f / leaf;
    int3 min_b = make_int3((int)floorf(h[0] * inv_leaf), (int)floorf(h[1] * inv_leaf), (int)floorf(h[2] * inv_leaf));
    int3 max_b = make_int3((int)floorf(h[3] * inv_leaf), (int)floorf(h[4] * inv_leaf), (int)floorf(h[5] * inv_leaf));
    int3 div_b = make_int3(max_b.x - min_b.x + 1, max_b.y - min_b.y + 1, max_b.z - min_b.z + 1);
    vg_cell_kernel<<<grid, 256, 0, cur_stream()>>>(dx, xs, n, inv_leaf, min_b, div_b, cell, idx);
    count_launch();
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, sort_tmp, cell, cell_s, idx, idx_s, n, 0, 64, cur_stream()));
    vg_heads_kernel<<<grid, 256, 0, cur_stream()>>>(cell_s, n, heads, d_nh);
    count_launch();
    uint32_t nh = 0;
    PPF_CUDA_TRY(memcpy_sync(&nh, d_nh, 4, cudaMemcpyDeviceToHost));
    if (nh == 0) return PPF_OK;
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortKeys(tmp, sort2_tmp, heads, heads_s, (int)nh, 0, 32, cur_stream()));   // ascending cell order
    vg_centroid_kernel<<<std::min(((int)nh + 127) / 128, 148 * 8), 128, 0, cur_stream()>>>(dx, xs, dn, ns, cell_s, idx_s, heads_s, (int)nh, n, oxyz, onrm);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    cudaMemcpyKind kind = (mem == PPF_MEM_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (out_xyz) PPF_CUDA_TRY(memcpy_sync(out_xyz, oxyz, (size_t)nh * 12, kind));
    if (out_nrm) PPF_CUDA_TRY(memcpy_sync(out_nrm, onrm, (size_t)nh * 12, kind));
    *n_out = (int)nh;
    return PPF_OK;
}

}  // namespace ppf
