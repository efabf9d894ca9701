// ppf_model.cu -- cloud upload, per-point local frames, and the model hash-table
// build (model_description).  Replaces Scene::initPPFs (scene.cu:64-99),
// Model::Model (model.cu:43-82) and ParallelHashArray's constructor
// (parallel_hash_array.hpp:55-77) + histogram (util.hpp:30-52).
//
// Differences from the reference, none of which change results:
//  * no N*N float4 feature matrix: one fused kernel goes point pair -> feature bins
//    -> FNV key, 8 B written per pair (key + pair index) instead of 20 B;
//  * pair indices are u32 (N <= 46340, the reference's own int limit) so the radix
//    sort moves 8 B per pair and pass instead of 12 B;
//  * the vote payload (chunk-local m_r, alpha_m as a 20-bit binary angle) is
//    computed in bucket order once, so voting streams 4 B per vote.
#include <cub/cub.cuh>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <unordered_map>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"
#include "ppf_radix.cuh"

namespace ppf {

// ---------------------------------------------------------------------------------
// Cloud
// ---------------------------------------------------------------------------------
// Stored point p = the caller's point order[p] (identity when order == nullptr).
__global__ void pack_cloud_kernel(const float *__restrict__ xyz, int xs, const float *__restrict__ nrm,
                                  int ns, int n, const uint32_t *__restrict__ order, float4 *pos, float4 *nrmo,
                                  float4 *fy, float4 *fz) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    size_t i = order ? order[p] : (uint32_t)p;
    float px = xyz[i * xs], py = xyz[i * xs + 1], pz = xyz[i * xs + 2];
    float nx = nrm[i * ns], ny = nrm[i * ns + 1], nz = nrm[i * ns + 2];
    pos[p] = make_float4(px, py, pz, 0.f);
    nrmo[p] = make_float4(nx, ny, nz, norm3(nx, ny, nz));
    FrameYZ f = frame_yz(px, py, pz, nx, ny, nz);
    fy[p] = make_float4(f.y[0], f.y[1], f.y[2], f.y[3]);
    fz[p] = make_float4(f.z[0], f.z[1], f.z[2], f.z[3]);
}

// ---- device block cache ------------------------------------------------------------------------------
// A recognition loop creates and destroys one Scene per frame (ppf.cu:56-99).  cudaMalloc / cudaFree
// synchronise the device and were measured at 1-800 ms per Scene on a B200 box under load, so a cloud lives
// in ONE device block and destroyed clouds park their block here for the next cloud of similar size.
namespace {
struct CachedBlock { void *ptr; size_t cap; int device; };
std::mutex g_pool_mutex;
std::vector<CachedBlock> g_pool;
constexpr size_t kPoolMaxBlocks = 64;
constexpr size_t kPoolMaxBytes = (size_t)6 << 30;
}  // namespace

// device every block handed out by pool_alloc was allocated on (a handle may be destroyed while another device is
// current -- ppf_registration switches devices, Python finalizers run at any time)
static std::unordered_map<void *, int> g_block_device;

void *pool_alloc(size_t bytes, size_t *cap_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        size_t best = g_pool.size();
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i].device == dev && g_pool[i].cap >= bytes && g_pool[i].cap <= 2 * bytes + (1u << 20) &&
                (best == g_pool.size() || g_pool[i].cap < g_pool[best].cap))
                best = i;
        if (best != g_pool.size()) {
            void *p = g_pool[best].ptr;
            *cap_out = g_pool[best].cap;
            g_pool.erase(g_pool.begin() + best);
            return p;
        }
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        pool_trim();                                  // give the cached blocks back and try once more
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    }
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        g_block_device[p] = dev;
    }
    *cap_out = bytes;
    return p;
}
void pool_free(void *p, size_t cap) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        int dev = 0;
        auto it = g_block_device.find(p);
        if (it != g_block_device.end()) dev = it->second; else cudaGetDevice(&dev);
        size_t total = cap;
        for (auto &b : g_pool) total += b.cap;
        if (g_pool.size() < kPoolMaxBlocks && total <= kPoolMaxBytes) {
            g_pool.push_back({p, cap, dev});
            return;
        }
        g_block_device.erase(p);
    }
    cudaFree(p);
}
void pool_trim() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (auto &b : g_pool) { cudaFree(b.ptr); g_block_device.erase(b.ptr); }
    g_pool.clear();
}

// cudaMalloc / cudaFree look-alikes on top of the block cache (the capacity of every live block is remembered)
namespace {
std::unordered_map<void *, size_t> g_live_caps;
}
cudaError_t pooled_malloc_bytes(void **p, size_t bytes) {
    size_t cap = 0;
    *p = pool_alloc(bytes ? bytes : 16, &cap);
    if (!*p) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_live_caps[*p] = cap;
    return cudaSuccess;
}
void pooled_free(void *p) {
    if (!p) return;
    size_t cap = 0;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto it = g_live_caps.find(p);
        if (it == g_live_caps.end()) { cudaFree(p); return; }      // not ours (cannot happen)
        cap = it->second;
        g_live_caps.erase(it);
    }
    pool_free(p, cap);
}

void cloud_free(Cloud &c) {
    if (c.block) {
        pool_free(c.block, c.block_cap);              // every array of the cloud lives in the block
    } else {                                          // a cloud read by model_load: one allocation per array
        pooled_free(c.pos); pooled_free(c.nrm); pooled_free(c.fy); pooled_free(c.fz);
        pooled_free(c.order); pooled_free(c.inv); pooled_free(c.gbox_lo); pooled_free(c.gbox_hi); pooled_free(c.tbox_lo); pooled_free(c.tbox_hi);
    }
    c = Cloud();
}

// ---- spatial (Morton) order + bounding boxes of a scene cloud --------------------------------------
__device__ __forceinline__ void atomic_min_f(float *addr, float v) {
    if (v >= 0) atomicMin((int *)addr, __float_as_int(v)); else atomicMax((unsigned int *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *addr, float v) {
    if (v >= 0) atomicMax((int *)addr, __float_as_int(v)); else atomicMin((unsigned int *)addr, __float_as_uint(v));
}
__global__ void bbox_kernel(const float *xyz, int xs, int n, float *mm) {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; c++) {
            float v = xyz[(size_t)i * xs + c];
            if (fabsf(v) < 3.0e38f) { lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
        }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_f(mm + c, lo[c]); atomic_max_f(mm + 3 + c, hi[c]); }
    }
}
__device__ __forceinline__ uint32_t spread10(uint32_t v) {               // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;  v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void morton_kernel(const float *xyz, int xs, int n, const float *mm, uint32_t *key, uint32_t *idx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = 0;
        for (int c = 0; c < 3; c++) {
            float lo = mm[c], ext = mm[3 + c] - mm[c];
            float t = ext > 0.f ? (xyz[(size_t)i * xs + c] - lo) / ext : 0.f;
            int q = (t == t) ? min(1023, max(0, (int)(t * 1024.f))) : 1023;    // NaN / Inf points go last
            k |= spread10((uint32_t)q) << c;
        }
        key[i] = k; idx[i] = (uint32_t)i;
    }
}
__global__ void invert_kernel(const uint32_t *order, int n, uint32_t *inv) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) inv[order[p]] = (uint32_t)p;
}
// AABB of every `group` consecutive stored points (NaN coordinates are ignored by fminf / fmaxf).
__global__ void boxes_kernel(const float4 *pos, int n, int group, float4 *lo_out, float4 *hi_out, int n_boxes) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_boxes; b += gridDim.x * blockDim.x) {
        float3 lo = make_float3(3.0e38f, 3.0e38f, 3.0e38f), hi = make_float3(-3.0e38f, -3.0e38f, -3.0e38f);
        for (int i = b * group; i < min(n, (b + 1) * group); i++) {
            float4 p = pos[i];
            lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
            hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
        }
        lo_out[b] = make_float4(lo.x, lo.y, lo.z, 0.f); hi_out[b] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
}

int cloud_create(const float *xyz, int xs, const float *nrm, int ns, int n, int mem, Cloud &c, bool spatial_sort) {
    if (!xyz || !nrm || n < 0 || xs < 3 || ns < 3) {
        set_last_error("cloud: null pointer, negative size or stride < 3");
        return PPF_ERR_INVALID;
    }
    c.n = n;
    const size_t nn = n > 0 ? (size_t)n : 1;
    const int ng = (n + 31) / 32, nt = (n + kHitQueue - 1) / kHitQueue;
    const size_t bx = n > 0 ? ((size_t)(n - 1) * xs + 3) * sizeof(float) : 0, bn = n > 0 ? ((size_t)(n - 1) * ns + 3) * sizeof(float) : 0;
    size_t sort_tmp = 0;
    if (n > 0)
        cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (uint32_t *)nullptr, n);
    // one block: the cloud's arrays, then the scratch of this function (host staging, Morton keys, CUB storage)
    const size_t persistent = 4 * nn * sizeof(float4) + (spatial_sort ? 2 * nn * 4 + 2 * (size_t)(ng + nt + 2) * sizeof(float4) : 0);
    const size_t scratch = (mem == PPF_MEM_HOST ? bx + bn : 0) + (spatial_sort ? nn * 12 + sort_tmp + 64 : 0);
    const size_t want = persistent + scratch + 32 * 256;
    c.block = (char *)pool_alloc(want, &c.block_cap);
    if (!c.block) { set_last_error("cloud: out of device memory"); return PPF_ERR_CUDA; }
    Workspace ws;                                      // bump allocator over the block (does not own it)
    ws.base = c.block; ws.cap = c.block_cap; ws.used = 0;
    c.pos = ws.take<float4>(nn); c.nrm = ws.take<float4>(nn); c.fy = ws.take<float4>(nn); c.fz = ws.take<float4>(nn);
    if (n == 0) return PPF_OK;
    const float *dx = xyz, *dn = nrm;
    if (mem == PPF_MEM_HOST) {
        float *tx = (float *)ws.take_bytes(bx), *tn = (float *)ws.take_bytes(bn);
        PPF_CUDA_TRY(cudaMemcpyAsync(tx, xyz, bx, cudaMemcpyHostToDevice, cur_stream()));
        PPF_CUDA_TRY(cudaMemcpyAsync(tn, nrm, bn, cudaMemcpyHostToDevice, cur_stream()));
        dx = tx; dn = tn;
    }
    const int grid = std::min((n + 255) / 256, 148 * 8);
    float *mm_dev = nullptr;
    if (!spatial_sort) {
        // model clouds: the AABB alone (bounds the distance bins of the model's pairs: sizes the cell map of the build)
        float *mm = ws.take<float>(6);
        mm_dev = mm;
        const float init[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
        PPF_CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, cur_stream()));
        bbox_kernel<<<grid, 256, 0, cur_stream()>>>(dx, xs, n, mm);
        count_launch();
    }
    if (spatial_sort) {
        float *mm = ws.take<float>(6);
        mm_dev = mm;
        uint32_t *key = ws.take<uint32_t>(n), *key_s = ws.take<uint32_t>(n), *idx = ws.take<uint32_t>(n);
        void *tmp = ws.take_bytes(sort_tmp);
        c.order = ws.take<uint32_t>(n); c.inv = ws.take<uint32_t>(n);
        const float init[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
        PPF_CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, cur_stream()));
        bbox_kernel<<<grid, 256, 0, cur_stream()>>>(dx, xs, n, mm);
        count_launch();
        morton_kernel<<<grid, 256, 0, cur_stream()>>>(dx, xs, n, mm, key, idx);
        count_launch();
        PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, sort_tmp, key, key_s, idx, c.order, n, 0, 30, cur_stream()));
        invert_kernel<<<grid, 256, 0, cur_stream()>>>(c.order, n, c.inv);
        count_launch();
    }
    pack_cloud_kernel<<<(n + 255) / 256, 256, 0, cur_stream()>>>(dx, xs, dn, ns, n, c.order, c.pos, c.nrm, c.fy, c.fz);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    if (spatial_sort) {
        c.gbox_lo = ws.take<float4>(ng); c.gbox_hi = ws.take<float4>(ng);
        c.tbox_lo = ws.take<float4>(nt); c.tbox_hi = ws.take<float4>(nt);
        if (!c.gbox_lo || !c.gbox_hi || !c.tbox_lo || !c.tbox_hi) { set_last_error("cloud: block too small"); return PPF_ERR_CUDA; }
        boxes_kernel<<<std::min((ng + 127) / 128, 148 * 8), 128, 0, cur_stream()>>>(c.pos, n, 32, c.gbox_lo, c.gbox_hi, ng);
        count_launch();
        boxes_kernel<<<std::min((nt + 127) / 128, 148 * 8), 128, 0, cur_stream()>>>(c.pos, n, kHitQueue, c.tbox_lo, c.tbox_hi, nt);
        count_launch();
        PPF_CUDA_TRY(cudaGetLastError());
    }
    if (mm_dev) {                                       // host copy of the AABB: bounds the distance bins of the cloud's pairs
        float mm_h[6];
        PPF_CUDA_TRY(cudaMemcpyAsync(mm_h, mm_dev, sizeof(mm_h), cudaMemcpyDeviceToHost, cur_stream()));
        PPF_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
        for (int k = 0; k < 3; k++) { c.bb_lo[k] = mm_h[k]; c.bb_hi[k] = mm_h[3 + k]; }
    }
    PPF_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return PPF_OK;
}

__device__ __forceinline__ PointN load_point(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, int i) {
    float4 p = __ldg(pos + i), q = __ldg(nrm + i);
    PointN r;
    r.x = p.x; r.y = p.y; r.z = p.z; r.nx = q.x; r.ny = q.y; r.nz = q.z; r.nn = q.w;
    return r;
}

// ---------------------------------------------------------------------------------
// Debug / parity view: quantised features + keys of a tile (ppf_kernel + ppf_hash_kernel)
// ---------------------------------------------------------------------------------
__global__ void features_tile_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm,
                                     const uint32_t *__restrict__ inv, int n,
                                     float d_dist, float inv_d, unsigned df, int rb, int re, int ob, int oe,
                                     float4 *ppfs, uint32_t *keys) {
    size_t w = (size_t)(oe - ob), total = (size_t)(re - rb) * w;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int r = rb + (int)(t / w), o = ob + (int)(t % w);
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t key = 0;
        if (n <= 1) {
            // ppf_kernel / ppf_hash_kernel return early for count <= 1: zero-initialised memory
        } else if ((r % df) != 0 || r == o) {
            out.x = CUDART_NAN_F;                                    // kernel.cu:432-441
        } else {
            PointN a = load_point(pos, nrm, inv ? inv[r] : r), b = load_point(pos, nrm, inv ? inv[o] : o);
            FeatureBins fb = pair_feature_bins(a, b, d_dist, inv_d);
            // kd == INT_MAX marks a distance more than 4e6 bins away: no table can hold it, so only
            // this debug view needs its quantised value; use the reference's own formula there.
            if (fb.kd < 0) out.x = CUDART_NAN_F;
            else if (fb.kd == 0x7FFFFFFF) out.x = __fsub_rn(fb.f1, fmodf(fb.f1, d_dist));
            else out.x = quant_value(fb.kd, d_dist);
            out.y = __uint_as_float(angle_bits(fb.k1));
            out.z = __uint_as_float(angle_bits(fb.k2));
            out.w = __uint_as_float(angle_bits(fb.k3));
            key = (out.x != out.x) ? 0u
                                   : fnv1a_4(__float_as_uint(out.x), __float_as_uint(out.y),
                                             __float_as_uint(out.z), __float_as_uint(out.w));
        }
        if (ppfs) ppfs[t] = out;
        if (keys) keys[t] = key;
    }
}

int features_tile(const Cloud &c, float d_dist, unsigned df, int rb, int re, int ob, int oe, float *ppfs_host,
                  uint32_t *keys_host) {
    if (rb < 0 || ob < 0 || re > c.n || oe > c.n || rb > re || ob > oe || df == 0 || !(d_dist > 0.f)) {
        set_last_error("features: bad tile bounds, df == 0 or d_dist <= 0");
        return PPF_ERR_INVALID;
    }
    size_t total = (size_t)(re - rb) * (size_t)(oe - ob);
    if (total == 0) return PPF_OK;
    float4 *dp = nullptr; uint32_t *dk = nullptr;
    if (ppfs_host) PPF_CUDA_TRY(pooled_malloc(&dp, total * sizeof(float4)));
    if (keys_host) PPF_CUDA_TRY(pooled_malloc(&dk, total * sizeof(uint32_t)));
    int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    features_tile_kernel<<<blocks, 256, 0, cur_stream()>>>(c.pos, c.nrm, c.inv, c.n, d_dist, 1.0f / d_dist, df, rb, re, ob, oe, dp, dk);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    if (dp) PPF_CUDA_TRY(memcpy_sync(ppfs_host, dp, total * sizeof(float4), cudaMemcpyDeviceToHost));
    if (dk) PPF_CUDA_TRY(memcpy_sync(keys_host, dk, total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    pooled_free(dp); pooled_free(dk);
    return PPF_OK;
}

// ---------------------------------------------------------------------------------
// Model table build
// ---------------------------------------------------------------------------------
// The reference sorts all N^2 (FNV key, pair index) records with a 32-bit radix sort (Model::Model, model.cu:53-60;
// ParallelHashArray, parallel_hash_array.hpp:55-77).  A model has only a few thousand distinct keys -- one per occupied
// cell of the quantised feature space -- so the table is built by sorting the pairs by BUCKET RANK instead:
//   1. model_cells_kernel: pair -> cell code (the quantised feature itself, no hash) + a byte map of the cells that
//      occur (sized from the cloud's bounding box);                                             4 B / pair written
//   2. cell_keys_kernel: FNV key of every occupied cell (a few thousand hashes); sort + unique of those keys ->
//      hashkeys[U] (cells whose keys collide share a bucket, as in the reference, where only the key is compared);
//      cell_table_kernel: cell -> bucket rank;
//   3. own stable LSD radix sort (ppf_radix.cuh) over ceil(log2 U) rank bits (13 for a 10k-point model: 2 passes
//      instead of 4 over 32 key bits); the histogram kernel of the first pass turns the codes into ranks in place (one
//      L2-resident table read) and the pair index is the implicit payload of that pass;  per pass 4 + 8 B read, 8 B written
//   4. bucket_bounds_kernel: first / counts from the rank boundaries of the sorted run.         4 B / pair read
// The arrays are bit-identical to the reference's: ranks ascend with the keys and the sort is stable in the pair index.
constexpr uint32_t kKey0Cell = 0xFFFFFFFFu;           // cell code of a pair whose key is 0 (self pair, NaN distance)

// One thread per ordered model pair p = m_r*N + m_i (coalesced along m_i): codes[p] = kd * 17^3 + cell of the three
// angle bins (ppf_kernel + the quantiser of ppf_hash_kernel; the hash itself is taken per CELL afterwards).
// occ[cell] = 1 for every cell a pair falls into (byte map sized from the cloud's bounding box: K_bound distance bins;
// plain stores, all writers store the same value; the map is small enough to live in L1 / L2).
__global__ void __launch_bounds__(256) model_cells_kernel(const float4 *__restrict__ pos,
                                                          const float4 *__restrict__ nrm, int n, float d_dist,
                                                          float inv_d, uint32_t *codes, int *max_kd, unsigned char *occ,
                                                          int K_bound) {
    size_t total = (size_t)n * n;
    int local_max = -1;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(p / n), i = (int)(p - (size_t)r * n);
        uint32_t code = kKey0Cell;
        if (r != i) {
            PointN a = load_point(pos, nrm, r), b = load_point(pos, nrm, i);
            FeatureBins fb = pair_feature_bins(a, b, d_dist, inv_d);
            if (fb.kd >= 0) {
                local_max = max(local_max, fb.kd);
                // kd >= 65536 (d_dist absurdly small for this model) is refused by the host; keep the code in range
                code = cell_index(min(fb.kd, 65535), fb.k1, fb.k2, fb.k3);
                if (fb.kd < K_bound && !occ[code]) occ[code] = 1;          // (kd >= K_bound: refused by the host)
            }
        }
        codes[p] = code;
    }
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 16));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 8));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 4));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 2));
    local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, 1));
    if ((threadIdx.x & 31) == 0 && local_max >= 0) atomicMax(max_kd, local_max);
}

// (key, cell) of every occupied cell, in any order; slot 0 is the key-0 pseudo cell (the self pairs always exist)
__global__ void cell_keys_kernel(const unsigned char *__restrict__ occ, int K_d, float d_dist, uint32_t *keys, uint32_t *cells,
                                 uint32_t *count) {
    const int total = K_d * kCellsPerDist;
    if (blockIdx.x == 0 && threadIdx.x == 0) { keys[0] = 0u; cells[0] = kKey0Cell; }
    for (int cidx = blockIdx.x * blockDim.x + threadIdx.x; cidx < total; cidx += gridDim.x * blockDim.x) {
        if (!occ[cidx]) continue;
        int k3 = cidx % kAngleCells, t = cidx / kAngleCells;
        int k2 = t % kAngleCells; t /= kAngleCells;
        int k1 = t % kAngleCells; int kd = t / kAngleCells;
        const uint32_t slot = 1u + atomicAdd(count, 1u);
        keys[slot] = feature_key(kd, k1, k2, k3, d_dist);
        cells[slot] = (uint32_t)cidx;
    }
}

// sorted occupied-cell keys -> hashkeys (one per distinct key).  heads[i] = inclusive count of distinct keys up to i.
__global__ void key_heads_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *heads) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        heads[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void unique_keys_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ ranks1, uint32_t n,
                                   uint32_t *hashkeys) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (i == 0 || keys[i] != keys[i - 1]) hashkeys[ranks1[i] - 1u] = keys[i];
}

// codes[p] -> bucket rank (in place) + the sort's payload idx[p] = p (hashkeyToDataMap)
__global__ void __launch_bounds__(256) pair_ranks_kernel(uint32_t *codes, const uint32_t *__restrict__ cell2bucket,
                                                         size_t total, uint32_t *idx) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = codes[p];
        // key 0 (self pair, NaN distance, or a cell whose FNV key happens to be 0) is the smallest key: rank 0
        const uint32_t r = (c == kKey0Cell) ? 0u : __ldg(cell2bucket + c);
        codes[p] = (r == kNoBucket) ? 0u : r;
        if (idx) idx[p] = (uint32_t)p;
    }
}

// first[r] = where rank r starts in the sorted run; counts from the differences (every rank occurs)
__global__ void __launch_bounds__(256) bucket_bounds_kernel(const uint32_t *__restrict__ ranks_sorted, size_t total,
                                                            uint32_t *first) {
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = ranks_sorted[q];
        if (q == 0 || ranks_sorted[q - 1] != r) first[r] = (uint32_t)q;
    }
}
__global__ void bucket_counts_kernel(const uint32_t *__restrict__ first, uint32_t U, uint32_t total, uint32_t *counts) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < U; r += gridDim.x * blockDim.x)
        counts[r] = (r + 1u < U ? first[r + 1u] : total) - first[r];
}

// entries[q] = [theta : 20 | slow : 1 | m_r - chunk_base : 11] for sorted position q, where theta = 20-bit binary
// angle of u = (T_mg m_i).yz (alpha_m of Drost et al.), bit 31 of theta_code = slow.  theta is COMPUTED here from
// the pair the sort put at q (two L2-resident point loads + ~150 instructions) instead of being written per pair
// before the sort and gathered after it: that gather was a random 4-byte read per pair (a 32-byte sector each,
// 1.7 ms for 1e8 pairs) and needed an N^2 scratch array.
__global__ void __launch_bounds__(256) entries_kernel(const uint32_t *__restrict__ map, const float4 *__restrict__ pos,
                                                      const float4 *__restrict__ fy, const float4 *__restrict__ fz,
                                                      size_t total, int n, int chunk_rows, uint32_t *entries) {
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const uint32_t p = map[q];
        const uint32_t r = p / (uint32_t)n, i = p - r * (uint32_t)n;
        uint32_t th = 0;
        if (r != i) {
            const float4 y = __ldg(fy + r), z = __ldg(fz + r), b = __ldg(pos + i);
            FrameYZ f;
            f.y[0] = y.x; f.y[1] = y.y; f.y[2] = y.z; f.y[3] = y.w;
            f.z[0] = z.x; f.z[1] = z.y; f.z[2] = z.z; f.z[3] = z.w;
            float uy, uz;
            frame_apply_yz(f, b.x, b.y, b.z, uy, uz);
            th = theta_code(uy, uz);
        }
        entries[q] = pack_entry(r % (uint32_t)chunk_rows, th);
    }
}

// ranges[c][b] = slice of bucket b with m_r in [c*chunk_rows, (c+1)*chunk_rows)
__global__ void chunk_ranges_kernel(const uint32_t *__restrict__ map, const uint32_t *__restrict__ first,
                                    const uint32_t *__restrict__ counts, uint32_t U, int n, int chunk_rows,
                                    int n_chunks, uint2 *ranges) {
    size_t total = (size_t)U * n_chunks;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        uint32_t b = (uint32_t)(t % U);
        int c = (int)(t / U);
        uint32_t lo0 = first[b], hi0 = lo0 + counts[b];
        uint32_t bound[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            unsigned long long want = (unsigned long long)(c + s) * chunk_rows * (unsigned long long)n;
            uint32_t lo = lo0, hi = hi0;
            while (lo < hi) {
                uint32_t mid = lo + ((hi - lo) >> 1);
                if ((unsigned long long)map[mid] < want) lo = mid + 1; else hi = mid;
            }
            bound[s] = lo;
        }
        ranges[(size_t)c * U + b] = make_uint2(bound[0], bound[1] - bound[0]);
    }
}

// cell -> bucket: hash the cell's quantised feature the way ppf_hash_kernel would and
// binary-search it in the unique keys (ParallelHashArray::GetIndices + the hit test of
// ppf_vote_count_kernel, kernel.cu:489-497).  Key 0 is the reference's "invalid" marker.
__global__ void cell_table_kernel(const uint32_t *__restrict__ hashkeys, uint32_t U, int K_d, float d_dist,
                                  uint32_t *cell2bucket) {
    int total = K_d * kCellsPerDist;
    for (int cidx = blockIdx.x * blockDim.x + threadIdx.x; cidx < total; cidx += gridDim.x * blockDim.x) {
        int k3 = cidx % kAngleCells, t = cidx / kAngleCells;
        int k2 = t % kAngleCells; t /= kAngleCells;
        int k1 = t % kAngleCells; int kd = t / kAngleCells;
        uint32_t key = feature_key(kd, k1, k2, k3, d_dist);
        uint32_t res = kNoBucket;
        if (key != 0u) {
            uint32_t lo = 0, hi = U;
            while (lo < hi) {
                uint32_t mid = lo + ((hi - lo) >> 1);
                if (hashkeys[mid] < key) lo = mid + 1; else hi = mid;
            }
            if (lo < U && hashkeys[lo] == key) res = lo;
        }
        cell2bucket[cidx] = res;
    }
}

// Far cells: hash every cell with kd in [kd0, kd1) the way ppf_hash_kernel would and keep those whose key is a model
// key (kernel.cu:480-501 matches by key equality alone).  Expected output: #cells x U / 2^32 entries.
__global__ void far_cells_kernel(const uint32_t *__restrict__ hashkeys, uint32_t U, int kd0, int kd1, float d_dist,
                                 unsigned long long *cells, uint32_t *buckets, uint32_t cap, uint32_t *count) {
    const unsigned long long total = (unsigned long long)(kd1 - kd0) * kCellsPerDist;
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const int kd = kd0 + (int)(t / kCellsPerDist);
        int r = (int)(t % kCellsPerDist);
        const int k3 = r % kAngleCells; r /= kAngleCells;
        const int k2 = r % kAngleCells, k1 = r / kAngleCells;
        const uint32_t key = feature_key(kd, k1, k2, k3, d_dist);
        if (key == 0u) continue;
        uint32_t lo = 0, hi = U;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(hashkeys + mid) < key) lo = mid + 1; else hi = mid;
        }
        if (lo < U && __ldg(hashkeys + lo) == key) {
            const uint32_t slot = atomicAdd(count, 1u);
            if (slot < cap) { cells[slot] = (unsigned long long)kd * kCellsPerDist + (unsigned long long)(t % kCellsPerDist); buckets[slot] = lo; }
        }
    }
}

// The far cells a scene of this extent can reach (kd below the scene's longest possible pair), as device arrays
// sorted by cell id.  Hashes the missing distance bins on first use and caches them in the model handle.
int model_far_cells(const ModelTable &m, const Cloud &scene, const unsigned long long **cells, const uint32_t **buckets,
                    int *n, int *kd_min, int *kd_max) {
    *cells = nullptr; *buckets = nullptr; *n = 0; *kd_min = 0; *kd_max = -1;
    if (!m.far || m.K_d <= 0 || m.U == 0 || scene.n <= 1) return PPF_OK;
    // longest pair of the scene <= diagonal of its AABB (non-finite points form no finite pair)
    double diag2 = 0.0;
    for (int k = 0; k < 3; k++) {
        const double e = (double)scene.bb_hi[k] - (double)scene.bb_lo[k];
        if (e > 0.0) diag2 += e * e;
    }
    const double bins = std::sqrt(diag2) * 1.0001 / (double)m.d_dist + 2.0;
    const int K_scene = (int)std::min<double>(bins, (double)kMaxDistBins);
    if (K_scene <= m.K_d) return PPF_OK;
    FarCells &fc = *m.far;
    std::lock_guard<std::mutex> lock(fc.mu);
    const int K_have = std::max(fc.K_scanned, m.K_d);
    if (K_scene > K_have) {
        const int K_new = (int)std::min<long long>(kMaxDistBins, ((long long)K_scene + 63) / 64 * 64);
        uint32_t cap = 4096;
        for (int attempt = 0; attempt < 2; attempt++) {
            unsigned long long *dc = nullptr; uint32_t *db = nullptr, *dn = nullptr;
            PPF_CUDA_TRY(pooled_malloc(&dc, (size_t)cap * 8)); PPF_CUDA_TRY(pooled_malloc(&db, (size_t)cap * 4));
            PPF_CUDA_TRY(pooled_malloc(&dn, 4));
            PPF_CUDA_TRY(cudaMemsetAsync(dn, 0, 4, cur_stream()));
            const unsigned long long total = (unsigned long long)(K_new - K_have) * kCellsPerDist;
            far_cells_kernel<<<(int)std::min<unsigned long long>((total + 255) / 256, 148 * 32), 256, 0, cur_stream()>>>(
                m.hashkeys, m.U, K_have, K_new, m.d_dist, dc, db, cap, dn);
            count_launch();
            uint32_t cnt = 0;
            cudaError_t e = memcpy_sync(&cnt, dn, 4, cudaMemcpyDeviceToHost);
            std::vector<unsigned long long> hc(std::min(cnt, cap));
            std::vector<uint32_t> hb(hc.size());
            if (e == cudaSuccess && !hc.empty()) e = memcpy_sync(hc.data(), dc, hc.size() * 8, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && !hc.empty()) e = memcpy_sync(hb.data(), db, hb.size() * 4, cudaMemcpyDeviceToHost);
            pooled_free(dc); pooled_free(db); pooled_free(dn);
            PPF_CUDA_TRY(e);
            if (cnt > cap) { cap = cnt; continue; }                // rare: count, then fetch
            std::vector<size_t> order(hc.size());
            for (size_t i = 0; i < order.size(); i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return hc[a] < hc[b]; });
            for (size_t i : order) { fc.cells_h.push_back(hc[i]); fc.buckets_h.push_back(hb[i]); }   // new cells lie above the old
            break;
        }
        fc.K_scanned = K_new;
        pooled_free(fc.cells); pooled_free(fc.buckets);
        fc.cells = nullptr; fc.buckets = nullptr;
        if (!fc.cells_h.empty()) {
            PPF_CUDA_TRY(pooled_malloc(&fc.cells, fc.cells_h.size() * 8));
            PPF_CUDA_TRY(pooled_malloc(&fc.buckets, fc.buckets_h.size() * 4));
            PPF_CUDA_TRY(memcpy_sync(fc.cells, fc.cells_h.data(), fc.cells_h.size() * 8, cudaMemcpyHostToDevice));
            PPF_CUDA_TRY(memcpy_sync(fc.buckets, fc.buckets_h.data(), fc.buckets_h.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    // only the cells this scene can reach (the cache may cover a larger scene met earlier)
    const unsigned long long lim = (unsigned long long)K_scene * kCellsPerDist;
    const size_t cnt = std::lower_bound(fc.cells_h.begin(), fc.cells_h.end(), lim) - fc.cells_h.begin();
    if (cnt == 0) return PPF_OK;
    *cells = fc.cells; *buckets = fc.buckets; *n = (int)cnt;
    *kd_min = (int)(fc.cells_h[0] / kCellsPerDist);
    *kd_max = (int)(fc.cells_h[cnt - 1] / kCellsPerDist);
    return PPF_OK;
}

__global__ void fill_kernel(float *v, int n, float x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}

void model_free(ModelTable &m) {
    cloud_free(m.cloud);
    pooled_free(m.hashkeys); pooled_free(m.counts); pooled_free(m.first);
    if (m.map_cap) pool_free(m.map, m.map_cap); else pooled_free(m.map);
    if (m.entries_cap) pool_free(m.entries, m.entries_cap); else pooled_free(m.entries);
    pooled_free(m.ranges); pooled_free(m.cell2bucket); pooled_free(m.weights);
    if (m.far) { pooled_free(m.far->cells); pooled_free(m.far->buckets); delete m.far; }
    m = ModelTable();
}

std::atomic<int> g_expected_scene_points{0};
constexpr int kDenseScenePoints = 40000;

int model_build(ModelTable &m) {
    const int n = m.cloud.n;
    if (n > PPF_MAX_MODEL_POINTS) {
        set_last_error("model: more than 46340 points (N*N pair indices must stay below 2^31, as in the reference)");
        return PPF_ERR_UNSUPPORTED;
    }
    if (!(m.d_dist > 0.f)) { set_last_error("model: d_dist must be > 0"); return PPF_ERR_INVALID; }
    m.inv_d_dist = 1.0f / m.d_dist;
    m.n_chunks = 1; m.chunk_rows = 32;
    if (!m.far) m.far = new FarCells();
    PPF_CUDA_TRY(pooled_malloc(&m.weights, std::max(1, n) * sizeof(float)));
    if (n > 0) fill_kernel<<<(n + 255) / 256, 256, 0, cur_stream()>>>(m.weights, n, 1.0f);
    count_launch();
    // Every reference kernel returns early when count <= 1 (kernel.cu:406,461): a model with
    // fewer than two points has an all-zero key array, i.e. one bucket (key 0) that can never match.
    size_t total = (size_t)n * n;
    if (n <= 1) {
        m.U = (uint32_t)total;              // n==1: one key (0); n==0: none
        m.K_d = 0;
        PPF_CUDA_TRY(pooled_malloc(&m.hashkeys, 4)); PPF_CUDA_TRY(pooled_malloc(&m.counts, 4));
        PPF_CUDA_TRY(pooled_malloc(&m.first, 4)); PPF_CUDA_TRY(pooled_malloc(&m.map, 4));
        PPF_CUDA_TRY(pooled_malloc(&m.entries, 4)); PPF_CUDA_TRY(pooled_malloc(&m.ranges, 8));
        PPF_CUDA_TRY(pooled_malloc(&m.cell2bucket, 4));
        uint32_t z = 0, one = 1;
        memcpy_sync(m.hashkeys, &z, 4, cudaMemcpyHostToDevice);
        memcpy_sync(m.counts, &one, 4, cudaMemcpyHostToDevice);
        memcpy_sync(m.first, &z, 4, cudaMemcpyHostToDevice);
        memcpy_sync(m.map, &z, 4, cudaMemcpyHostToDevice);
        return PPF_OK;
    }

    // All build temporaries live in ONE scratch allocation (one cudaMalloc + one cudaFree per build:
    // allocator calls, not kernels, dominate the build time of a small model).
    size_t sort_tmp = 0;
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                                 (uint32_t *)nullptr, (uint32_t *)nullptr, total));
    {   // + the scan / sort of the occupied-cell keys (at most total + 1 records) and the scratch of the own radix sort
        size_t t1 = 0, t2 = 0;
        PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                                     (uint32_t *)nullptr, (uint32_t *)nullptr, total + 1));
        PPF_CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, t2, (uint32_t *)nullptr, (uint32_t *)nullptr, total + 1));
        sort_tmp = std::max(std::max(sort_tmp, std::max(t1, t2)), radix_plan(total, 32).scratch_words * 4) + 256;
    }
    Workspace ws;
    int rc = ws.reserve(3 * total * 4 + sort_tmp + 1024);
    if (rc) return rc;
    uint32_t *ranks = ws.take<uint32_t>(total);                      // cell codes, then bucket ranks
    uint32_t *ranks_sorted = ws.take<uint32_t>(total), *iota = ws.take<uint32_t>(total);
    uint32_t *d_nocc = ws.take<uint32_t>(1);
    int *d_maxkd = ws.take<int>(1);
    void *tmp = ws.take_bytes(sort_tmp);
    struct Release { Workspace &w; ~Release() { w.release(); } } release_on_exit{ws};
    if (!ranks || !ranks_sorted || !iota || !d_nocc || !d_maxkd || !tmp) {
        set_last_error("model: scratch arena too small");
        return PPF_ERR_CUDA;
    }
    PPF_CUDA_TRY(cudaMemsetAsync(d_maxkd, 0xFF, 4, cur_stream()));
    PPF_CUDA_TRY(cudaMemsetAsync(d_nocc, 0, 4, cur_stream()));
    // distance bins a pair of this cloud can reach: longest pair <= diagonal of the AABB of the finite points
    double diag2 = 0.0;
    for (int k = 0; k < 3; k++) {
        const double e = (double)m.cloud.bb_hi[k] - (double)m.cloud.bb_lo[k];
        if (e > 0.0) diag2 += e * e;
    }
    const double kb = std::sqrt(diag2) * 1.0001 / (double)m.d_dist + 2.0;
    if (!(kb < 65536.0)) {
        set_last_error("model: d_dist is more than 65536x smaller than the model extent");
        return PPF_ERR_UNSUPPORTED;
    }
    const int K_bound = (int)kb;
    unsigned char *occ = nullptr;
    PPF_CUDA_TRY(pooled_malloc(&occ, (size_t)K_bound * kCellsPerDist));
    struct FreeOcc { unsigned char *p; ~FreeOcc() { pooled_free(p); } } free_occ{occ};
    PPF_CUDA_TRY(cudaMemsetAsync(occ, 0, (size_t)K_bound * kCellsPerDist, cur_stream()));
    int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
    model_cells_kernel<<<grid, 256, 0, cur_stream()>>>(m.cloud.pos, m.cloud.nrm, n, m.d_dist, m.inv_d_dist, ranks, d_maxkd, occ, K_bound);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    int h_maxkd = -1;
    PPF_CUDA_TRY(memcpy_sync(&h_maxkd, d_maxkd, 4, cudaMemcpyDeviceToHost));
    if (h_maxkd >= K_bound) {
        set_last_error("model: a pair longer than the cloud's bounding box diagonal (internal error)");
        return PPF_ERR_CUDA;
    }
    m.K_d = h_maxkd + 1;

    // occupied cells -> their keys -> unique sorted keys = hashkeys
    const size_t ncell = (size_t)std::max(1, m.K_d) * kCellsPerDist;
    PPF_CUDA_TRY(pooled_malloc(&m.cell2bucket, ncell * 4));
    // the (key, cell) lists of the occupied cells: at most min(ncell, total) + 1 records, four arrays
    const size_t occ_cap = std::min(ncell, total) + 1;
    uint32_t *clists = nullptr;
    PPF_CUDA_TRY(pooled_malloc(&clists, 4 * occ_cap * sizeof(uint32_t)));
    struct FreeLists { uint32_t *p; ~FreeLists() { pooled_free(p); } } free_lists{clists};
    uint32_t *ckeys = clists, *ccells = clists + occ_cap, *ckeys_s = clists + 2 * occ_cap, *ccells_s = clists + 3 * occ_cap;
    cell_keys_kernel<<<(int)std::min<size_t>((ncell + 255) / 256, 148 * 32), 256, 0, cur_stream()>>>(
        occ, m.K_d, m.d_dist, ckeys, ccells, d_nocc);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    uint32_t n_occ = 0;
    PPF_CUDA_TRY(memcpy_sync(&n_occ, d_nocc, 4, cudaMemcpyDeviceToHost));
    n_occ += 1;                                                      // + the key-0 pseudo cell
    {
        // a few thousand records: library sort + scan (0.01% of the build's bytes)
        size_t t1 = 0, t2 = 0;
        PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, ckeys, ckeys_s, ccells, ccells_s, n_occ));
        PPF_CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, t2, ckeys, ckeys, n_occ));
        if (std::max(t1, t2) > sort_tmp) { set_last_error("model: scratch arena too small"); return PPF_ERR_CUDA; }
        PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, t1, ckeys, ckeys_s, ccells, ccells_s, n_occ, 0, 32, cur_stream()));
        const int g2 = (int)std::min<size_t>((n_occ + 255) / 256, 148 * 8);
        key_heads_kernel<<<g2, 256, 0, cur_stream()>>>(ckeys_s, n_occ, ckeys);           // ckeys reused: heads, then 1-based ranks
        PPF_CUDA_TRY(cub::DeviceScan::InclusiveSum(tmp, t2, ckeys, ckeys, n_occ, cur_stream()));
        PPF_CUDA_TRY(memcpy_sync(&m.U, ckeys + (n_occ - 1), 4, cudaMemcpyDeviceToHost));
        PPF_CUDA_TRY(pooled_malloc(&m.hashkeys, (size_t)m.U * 4));
        PPF_CUDA_TRY(pooled_malloc(&m.counts, (size_t)m.U * 4));
        PPF_CUDA_TRY(pooled_malloc(&m.first, (size_t)m.U * 4));
        unique_keys_kernel<<<g2, 256, 0, cur_stream()>>>(ckeys_s, ckeys, n_occ, m.hashkeys);
        count_launch(3);
    }
    // cell -> bucket (also for unoccupied cells whose key collides with a model key: they vote in the reference)
    cell_table_kernel<<<(int)std::min<size_t>((ncell + 255) / 256, 148 * 32), 256, 0, cur_stream()>>>(m.hashkeys, m.U, m.K_d,
                                                                                   m.d_dist, m.cell2bucket);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());

    // pairs -> ranks; stable radix sort over the rank bits only (ppf_radix.cuh); every bucket ascends in pair index
    m.map = (uint32_t *)pool_alloc(total * 4, &m.map_cap);      // the two N^2 arrays come from the block cache
    if (!m.map) { set_last_error("model: out of device memory"); return PPF_ERR_CUDA; }
    int rank_bits = 1;
    while (rank_bits < 32 && (1ull << rank_bits) < (unsigned long long)m.U) rank_bits++;
    const RadixPlan plan = radix_plan(total, rank_bits);
    const bool library_sort = getenv("PPF_B200_SORT") && !strcmp(getenv("PPF_B200_SORT"), "cub");   // A/B hook
    uint32_t *rk[2] = {ranks, ranks_sorted};
    // the sorted payload must land in m.map: with an even number of passes the payload starts there
    uint32_t *pv[2] = {iota, m.map};
    if (!library_sort && (plan.passes & 1) == 0) { pv[0] = m.map; pv[1] = iota; }
    const uint32_t *ranks_final = ranks_sorted;
    if (library_sort) {
        pair_ranks_kernel<<<grid, 256, 0, cur_stream()>>>(ranks, m.cell2bucket, total, pv[0]);
        count_launch();
        PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, sort_tmp, ranks, ranks_sorted, iota, m.map, total, 0, rank_bits, cur_stream()));
    } else {
        // the histogram kernel of pass 0 turns the codes into ranks in place (cell table read) and pass 0 takes the pair
        // index as the payload: no rank pass, no index array.  (Translating inside the SCATTER kernel was measured
        // slower: there the dependent gather sits on the critical path of a tile, 2.04 ms.)
        if (plan.scratch_words * 4 > sort_tmp) { set_last_error("model: scratch arena too small"); return PPF_ERR_CUDA; }
        count_launch(radix_sort_pairs(rk, pv, total, plan, (uint32_t *)tmp, cur_stream(), m.cell2bucket, true));
        ranks_final = rk[plan.passes & 1];
    }
    PPF_CUDA_TRY(cudaGetLastError());
    bucket_bounds_kernel<<<grid, 256, 0, cur_stream()>>>(ranks_final, total, m.first);
    bucket_counts_kernel<<<(int)std::min<size_t>(((size_t)m.U + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(m.first, m.U, (uint32_t)total, m.counts);
    count_launch(2);
    PPF_CUDA_TRY(cudaGetLastError());

    // Accumulator chunk geometry = which vote kernel serves this table.  The grouped kernel (small chunks,
    // big hit queue: ppf_vote_grouped.cu) wins when voting dominates, i.e. when buckets are long (10k-point
    // model: 14.7k entries per bucket on average, +17%); with short buckets (2k-point model: 800) hit
    // collection and sorting dominate and the one-hit-per-pass kernel with 1504-row chunks is faster.
    {
        const double avg_bucket = (double)total / (double)std::max<uint32_t>(1u, m.U);
        // dense scenes (hint from the caller, ppf_set_expected_scene_points) favour the grouped kernel even with
        // short buckets: every reference point then hits every bucket many times (2k-point model: 60k-point scene
        // 73 ms against 103 ms, 200k 198 / 287 ms, 1M 1.36 / 1.83 s; but 16k-point scene 43 / 34 ms)
        const bool dense_scene = g_expected_scene_points.load() >= kDenseScenePoints;
        m.prefer_grouped = (avg_bucket >= 5000.0 || dense_scene) && m.U < (1u << 20);   // 20-bit bucket field of a hit record
        if (const char *e = getenv("PPF_B200_VOTE")) {
            if (!strcmp(e, "classic")) m.prefer_grouped = 0;
            if (!strcmp(e, "grouped") && m.U < (1u << 20)) m.prefer_grouped = 1;
        }
        int max_rows = m.prefer_grouped ? kGroupedMaxRows : kMaxChunkRows;
        if (const char *e = getenv("PPF_B200_CHUNK_ROWS")) {        // test hook: force more / smaller chunks
            int v = atoi(e);
            if (v >= 32 && v <= kMaxChunkRows) max_rows = v / 32 * 32;
        }
        m.n_chunks = std::max(1, (n + max_rows - 1) / max_rows);
        m.chunk_rows = std::max(32, (((n + m.n_chunks - 1) / m.n_chunks) + 31) / 32 * 32);
    }
    // vote payload in bucket order, per-chunk bucket slices, cell table
    m.entries = (uint32_t *)pool_alloc(total * 4, &m.entries_cap);
    if (!m.entries) { set_last_error("model: out of device memory"); return PPF_ERR_CUDA; }
    entries_kernel<<<grid, 256, 0, cur_stream()>>>(m.map, m.cloud.pos, m.cloud.fy, m.cloud.fz, total, n, m.chunk_rows, m.entries);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    PPF_CUDA_TRY(pooled_malloc(&m.ranges, (size_t)m.U * m.n_chunks * sizeof(uint2)));
    {
        size_t t = (size_t)m.U * m.n_chunks;
        chunk_ranges_kernel<<<(int)std::min<size_t>((t + 255) / 256, 148 * 32), 256, 0, cur_stream()>>>(
            m.map, m.first, m.counts, m.U, n, m.chunk_rows, m.n_chunks, m.ranges);
        count_launch();
    }
    PPF_CUDA_TRY(cudaGetLastError());
    PPF_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return PPF_OK;
}

// ---- persistent model database (no reference equivalent: ppf.cu:63-70 rebuilds the table for
// every (scene, model) pair).  One little-endian file: header, then every device array of the table verbatim.
namespace {
struct ModelFileHeader {
    char magic[8];                     // "PPFB200\0"
    uint32_t version, n, U, K_d;
    int32_t n_chunks, chunk_rows, prefer_grouped, use_l1_norm, use_averaged_clusters;
    float d_dist, vote_count_threshold;
    uint32_t reserved[5];
};
constexpr uint32_t kModelFileVersion = 2;
constexpr size_t kIoChunk = (size_t)32 << 20;

int dev_to_file(FILE *f, const void *dev, size_t bytes, std::vector<char> &buf) {
    for (size_t off = 0; off < bytes; off += kIoChunk) {
        const size_t n = std::min(kIoChunk, bytes - off);
        PPF_CUDA_TRY(memcpy_sync(buf.data(), (const char *)dev + off, n, cudaMemcpyDeviceToHost));
        if (fwrite(buf.data(), 1, n, f) != n) { set_last_error("model save: short write"); return PPF_ERR_INVALID; }
    }
    return PPF_OK;
}
int file_to_dev(FILE *f, void **dev, size_t bytes, std::vector<char> &buf) {
    PPF_CUDA_TRY(pooled_malloc(dev, std::max<size_t>(bytes, 16)));
    for (size_t off = 0; off < bytes; off += kIoChunk) {
        const size_t n = std::min(kIoChunk, bytes - off);
        if (fread(buf.data(), 1, n, f) != n) { set_last_error("model load: file truncated"); return PPF_ERR_INVALID; }
        PPF_CUDA_TRY(memcpy_sync((char *)*dev + off, buf.data(), n, cudaMemcpyHostToDevice));
    }
    return PPF_OK;
}
struct ArraySpec { void **ptr; size_t bytes; };
std::vector<ArraySpec> model_arrays(ModelTable &m) {
    const size_t n = std::max(1, m.cloud.n), total = (size_t)m.cloud.n * m.cloud.n;
    const bool tiny = m.cloud.n <= 1;
    const size_t ncell = (size_t)std::max(1, m.K_d) * kCellsPerDist;
    return {
        {(void **)&m.cloud.pos, n * sizeof(float4)}, {(void **)&m.cloud.nrm, n * sizeof(float4)},
        {(void **)&m.cloud.fy, n * sizeof(float4)},  {(void **)&m.cloud.fz, n * sizeof(float4)},
        {(void **)&m.weights, n * sizeof(float)},
        {(void **)&m.hashkeys, tiny ? 4 : (size_t)m.U * 4}, {(void **)&m.counts, tiny ? 4 : (size_t)m.U * 4},
        {(void **)&m.first, tiny ? 4 : (size_t)m.U * 4},
        {(void **)&m.map, tiny ? 4 : total * 4}, {(void **)&m.entries, tiny ? 4 : total * 4},
        {(void **)&m.ranges, tiny ? 8 : (size_t)m.U * m.n_chunks * sizeof(uint2)},
        {(void **)&m.cell2bucket, tiny ? 4 : ncell * 4},
    };
}
}  // namespace

int model_save(const ModelTable &mc, const char *path) {
    ModelTable &m = const_cast<ModelTable &>(mc);          // model_arrays only reads the pointers here
    FILE *f = fopen(path, "wb");
    if (!f) { set_last_error(std::string("model save: cannot open ") + path); return PPF_ERR_INVALID; }
    ModelFileHeader h{};
    memcpy(h.magic, "PPFB200", 8);
    h.version = kModelFileVersion; h.n = (uint32_t)m.cloud.n; h.U = m.U; h.K_d = (uint32_t)m.K_d;
    h.n_chunks = m.n_chunks; h.chunk_rows = m.chunk_rows; h.prefer_grouped = m.prefer_grouped;
    h.use_l1_norm = m.use_l1_norm; h.use_averaged_clusters = m.use_averaged_clusters;
    h.d_dist = m.d_dist; h.vote_count_threshold = m.vote_count_threshold;
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? PPF_OK : PPF_ERR_INVALID;
    std::vector<char> buf(kIoChunk);
    for (auto &a : model_arrays(m)) {
        if (rc) break;
        rc = dev_to_file(f, *a.ptr, a.bytes, buf);
    }
    if (fclose(f) != 0 && !rc) { set_last_error("model save: close failed"); rc = PPF_ERR_INVALID; }
    return rc;
}

// A loaded table is used unchecked as indices by the vote kernels (shared-memory atomics, global gathers): reject a
// corrupt or re-written file here instead of faulting on the GPU later.  flag bits name the array that failed.
__global__ void validate_table_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ first,
                                      const uint32_t *__restrict__ hashkeys, const uint32_t *__restrict__ map,
                                      const uint32_t *__restrict__ entries, const uint2 *__restrict__ ranges,
                                      const uint32_t *__restrict__ cell2bucket, uint32_t U, size_t total, int n,
                                      int n_chunks, int chunk_rows, size_t ncell, uint32_t *flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    for (size_t i = t0; i < U; i += stride) {
        const uint32_t f = first[i], c = counts[i];
        if (c == 0 || (size_t)f + c > total) bad |= 1u;
        if (i + 1 < U ? (first[i + 1] != f + c || hashkeys[i + 1] <= hashkeys[i]) : ((size_t)f + c != total)) bad |= 1u;
        if (i == 0 && f != 0) bad |= 1u;
    }
    for (size_t q = t0; q < total; q += stride) {
        const uint32_t p = map[q];
        if ((size_t)p >= total) { bad |= 2u; continue; }
        // the entry's chunk-local row must be the row of the pair the map names
        if ((entries[q] & kLocMask) != (p / (uint32_t)n) % (uint32_t)chunk_rows) bad |= 4u;
    }
    for (size_t t = t0; t < (size_t)U * n_chunks; t += stride) {
        const uint2 r = ranges[t];
        const uint32_t b = (uint32_t)(t % U);
        if (r.x < first[b] || (size_t)r.x + r.y > (size_t)first[b] + counts[b]) bad |= 8u;
        if (r.y && (size_t)r.x + r.y <= total) {    // the slice really holds rows of chunk c only
            const int c = (int)(t / U);
            const uint32_t r0 = map[r.x] / (uint32_t)n, r1 = map[r.x + r.y - 1] / (uint32_t)n;
            if ((int)(r0 / (uint32_t)chunk_rows) != c || (int)(r1 / (uint32_t)chunk_rows) != c) bad |= 8u;
        } else if (r.y) {
            bad |= 8u;
        }
    }
    for (size_t i = t0; i < ncell; i += stride) {
        const uint32_t b = cell2bucket[i];
        if (b != kNoBucket && b >= U) bad |= 16u;
    }
    if (bad) atomicOr(flag, bad);
}

static int model_validate(const ModelTable &m) {
    if (m.cloud.n <= 1) return PPF_OK;
    const size_t total = (size_t)m.cloud.n * m.cloud.n, ncell = (size_t)std::max(1, m.K_d) * kCellsPerDist;
    uint32_t *flag = nullptr, h = 0;
    PPF_CUDA_TRY(pooled_malloc(&flag, 4));
    cudaMemsetAsync(flag, 0, 4, cur_stream());
    validate_table_kernel<<<148 * 8, 256, 0, cur_stream()>>>(m.counts, m.first, m.hashkeys, m.map, m.entries, m.ranges, m.cell2bucket, m.U,
                                            total, m.cloud.n, m.n_chunks, m.chunk_rows, ncell, flag);
    count_launch();
    cudaError_t e = memcpy_sync(&h, flag, 4, cudaMemcpyDeviceToHost);
    pooled_free(flag);
    PPF_CUDA_TRY(e);
    if (h) {
        set_last_error("model load: corrupt table payload (failed checks, bit mask " + std::to_string(h) +
                       ": 1 = keys/counts/first, 2 = map, 4 = entries, 8 = ranges, 16 = cell table)");
        return PPF_ERR_INVALID;
    }
    return PPF_OK;
}

int model_load(ModelTable &m, const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { set_last_error(std::string("model load: cannot open ") + path); return PPF_ERR_INVALID; }
    ModelFileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "PPFB200", 8) != 0 || h.version != kModelFileVersion ||
        h.n > (uint32_t)PPF_MAX_MODEL_POINTS || h.n_chunks < 1 || h.chunk_rows < 32 || h.chunk_rows > kMaxChunkRows ||
        (long long)h.n_chunks * h.chunk_rows < (long long)h.n || !(h.d_dist > 0.f) || h.K_d > 65536u ||
        (h.n > 1 && (h.U == 0 || (size_t)h.U > (size_t)h.n * h.n)) || (h.prefer_grouped && h.chunk_rows > kGroupedMaxRows)) {
        fclose(f);
        set_last_error("model load: not a ppf_b200 model file of this version (or corrupt header)");
        return PPF_ERR_INVALID;
    }
    m.cloud.n = (int)h.n; m.U = h.U; m.K_d = (int)h.K_d; m.n_chunks = h.n_chunks; m.chunk_rows = h.chunk_rows;
    m.prefer_grouped = h.prefer_grouped; m.use_l1_norm = h.use_l1_norm; m.use_averaged_clusters = h.use_averaged_clusters;
    m.d_dist = h.d_dist; m.inv_d_dist = 1.0f / h.d_dist; m.vote_count_threshold = h.vote_count_threshold;
    if (!m.far) m.far = new FarCells();
    std::vector<char> buf(kIoChunk);
    int rc = PPF_OK;
    for (auto &a : model_arrays(m)) {
        if (rc) break;
        rc = file_to_dev(f, a.ptr, a.bytes, buf);
    }
    if (!rc && fgetc(f) != EOF) { set_last_error("model load: trailing bytes"); rc = PPF_ERR_INVALID; }
    fclose(f);
    if (!rc) rc = model_validate(m);
    return rc;
}

int model_table_get(const ModelTable &m, uint32_t *hashkeys, size_t *counts, size_t *first, size_t *map) {
    size_t total = (size_t)m.cloud.n * m.cloud.n;
    std::vector<uint32_t> tmp;
    if (hashkeys && m.U) PPF_CUDA_TRY(memcpy_sync(hashkeys, m.hashkeys, (size_t)m.U * 4, cudaMemcpyDeviceToHost));
    auto widen = [&](const uint32_t *dev, size_t cnt, size_t *out) -> int {
        if (!out || !cnt) return PPF_OK;
        tmp.resize(cnt);
        PPF_CUDA_TRY(memcpy_sync(tmp.data(), dev, cnt * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; i++) out[i] = tmp[i];
        return PPF_OK;
    };
    int rc;
    if ((rc = widen(m.counts, m.U, counts))) return rc;
    if ((rc = widen(m.first, m.U, first))) return rc;
    if (m.cloud.n >= 2) { if ((rc = widen(m.map, total, map))) return rc; }
    else if (map && total) map[0] = 0;
    if (m.cloud.n <= 1 && counts && m.U) counts[0] = total;
    return PPF_OK;
}

}  // namespace ppf
