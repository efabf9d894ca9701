// ppf_internal.cuh -- device-resident objects shared by the translation units.
//
// HBM layout (all arrays are plain device allocations owned by the handle):
//
//   Cloud      pos[N]   float4 (x, y, z, 0)            -- 16 B vector loads
//              nrm[N]   float4 (nx, ny, nz, |n|)        -- |n| = sqrt.approx(n.n), hoisted out of the pair loop
//              fy[N], fz[N] float4                      -- rows y and z of the local frame T_g of each point
//                                                          (trans_model_scene, kernel.cu:310-327), computed once
//                                                          per point instead of once per vote
//   ModelTable hashkeys[U] counts[U] first[U] map[N*N]  -- the reference's ParallelHashArray contents
//                                                          (parallel_hash_array.hpp:36-46) as u32
//              entries[N*N] u32                         -- voting payload in bucket order:
//                                                          [theta_u : 20 | slow : 1 | m_r - chunk_base : 11]
//              ranges[n_chunks][U] uint2                -- (start, len) of the part of bucket b whose model
//                                                          reference points fall in chunk c (buckets ascend in
//                                                          m_r, so every chunk is one contiguous slice)
//              cell2bucket[K_d * 17^3] u32              -- quantised feature cell -> bucket index or kNoBucket;
//                                                          replaces hash + binary search on the scene side
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "ppf_math.cuh"

namespace ppf {

// Shared-memory vote accumulator geometry: 31 alpha bins x chunk_rows u32 counters.
constexpr int kMaxChunkRows = 1504;          // 31 * (1504 + 1) * 4 B = 186,620 B (+ 40,960 B hit queue)
constexpr int kHitQueue     = 2048;          // 16 B record + 4 B cursor each
constexpr int kVoteBatch    = 256;           // bucket entries per inner pass: 8 per lane, double-buffered in registers
constexpr int kVoteGrab     = 2048;          // bucket entries one warp takes per scheduler grab
// grouped vote kernel (ppf_vote_grouped.cu): a bigger hit queue (all hits of a reference point, sorted by
// bucket) in exchange for a smaller accumulator chunk
constexpr int kGroupedMaxRows = 640;         // 31 * (640 + 1) * 4 B = 79,484 B + 16,384 B entry stage + 11,264 hits x 12 B
                                             // (configs[1], 10k-point model: 384 rows 692 ms, 480 683, 576 675, 640 670,
                                             //  704 712 -- there the queue no longer holds the hits of the heaviest points)
// accumulator row stride: chunk_rows is a multiple of 32, so +1 makes bank = (bin + row) mod 32 --
// lanes that hit the same model point with different alpha bins (the common case inside a bucket,
// whose entries are sorted by m_r) fall into different banks.
__host__ __device__ constexpr int acc_stride(int chunk_rows) { return chunk_rows + 1; }

struct Cloud {
    int n = 0;
    float4 *pos = nullptr, *nrm = nullptr, *fy = nullptr, *fz = nullptr;
    // Scene clouds are stored in Morton (Z-curve) order so that 32 consecutive points are spatially
    // compact: the vote kernel culls whole warps / tiles of scene points that lie farther from the
    // reference point than any model pair is long.  order[p] = caller's index of stored point p,
    // inv[i] = stored position of the caller's point i (nullptr for model clouds: identity).
    uint32_t *order = nullptr, *inv = nullptr;
    char *block = nullptr;                             // the one device allocation all arrays above / below live in
    size_t block_cap = 0;                              // (nullptr for a cloud read by model_load)
    float4 *gbox_lo = nullptr, *gbox_hi = nullptr;     // AABB of every 32-point group
    float4 *tbox_lo = nullptr, *tbox_hi = nullptr;     // AABB of every kHitQueue-point tile
    float bb_lo[3] = {0.f, 0.f, 0.f}, bb_hi[3] = {0.f, 0.f, 0.f};   // AABB of the finite points (scene clouds; host copy)
};

// Scene-side feature cells BEYOND the model's distance range (kd >= K_d) whose 32-bit FNV key equals a model key.
// The reference matches scene pairs to model buckets by key equality alone (ppf_vote_count_kernel,
// kernel.cu:480-501), so such a far pair votes for the colliding bucket.  cell2bucket covers kd < K_d; this
// list (a handful of cells: #cells x U / 2^32) covers the rest, hashed lazily up to the extent of the largest
// scene the model has met.  cell id = kd * 17^3 + (k1 * 17 + k2) * 17 + k3.
struct FarCells {
    std::mutex mu;
    int K_scanned = 0;                                  // cells with kd in [K_d, K_scanned) have been hashed
    std::vector<unsigned long long> cells_h;            // ascending
    std::vector<uint32_t> buckets_h;
    unsigned long long *cells = nullptr;                // device copies (pooled)
    uint32_t *buckets = nullptr;
};
constexpr int kMaxDistBins = 1 << 22;                   // quant_bin is exact below 2^22 bins

struct ModelTable {
    Cloud cloud;
    float d_dist = 0.f, inv_d_dist = 0.f;
    uint32_t U = 0;                           // unique keys
    uint32_t *hashkeys = nullptr, *counts = nullptr, *first = nullptr, *map = nullptr;
    uint32_t *entries = nullptr;
    size_t map_cap = 0, entries_cap = 0;      // != 0: the array is a block of the device block cache
    int n_chunks = 0, chunk_rows = 0;
    int prefer_grouped = 0;                   // chunk geometry was chosen for the grouped vote kernel
    uint2 *ranges = nullptr;
    int K_d = 0;
    uint32_t *cell2bucket = nullptr;
    float *weights = nullptr;                 // modelPointVoteWeights, all 1.0 (model.cu:67)
    FarCells *far = nullptr;                  // lazily filled cache (owned; see model_far_cells)
    // options carried by the reference's Model object (model.h:43-46)
    float vote_count_threshold = 0.4f;
    int use_l1_norm = 0, use_averaged_clusters = 0;
};

// Grow-only device scratch arena: the per-lookup temporaries (filter output, sort buffers, CUB
// temp storage, clustering tables) are carved out of it, so a lookup performs no cudaMalloc /
// cudaFree in steady state (both synchronise the device and cost up to 100s of ms under load).
struct Workspace {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes);                 // make room for `bytes` (+ alignment slack), reset the bump pointer
    void *take_bytes(size_t bytes);            // 256-byte aligned slice, nullptr if the reservation was too small
    template <typename T> T *take(size_t n) { return static_cast<T *>(take_bytes(n * sizeof(T))); }
    void release();
};

// The exchanges between the ranks of a sharded recognition (ppf_comm.cu).  All buffers are device memory; the calls
// are issued on cur_stream() and every rank must make the same sequence of calls.
struct Comm {
    int rank = 0, world = 1;
    virtual ~Comm() {}
    virtual int allreduce_max_u32(uint32_t *dev, size_t n) = 0;
    virtual int allreduce_sum_f32(float *dev, size_t n) = 0;
    virtual int allgather_u32(const uint32_t *dev_send, uint32_t *dev_recv, size_t n) = 0;        // n words per rank
    // rank r contributes bytes[r] bytes, which land at dev_recv + offsets[r] on every rank
    virtual int allgatherv(const void *dev_send, void *dev_recv, const size_t *offsets, const size_t *bytes) = 0;
};

struct VoteResult {                           // device buffers of one ppf_lookup
    Workspace ws, ws2;                        // ws: filter / clustering scratch; ws2: merge + ordering scratch
    uint32_t cand_n = 0, local_max = 0;       // host copies of scalars[0], scalars[1] after the vote kernel
    // candidates emitted by the vote kernel (superset of the survivors)
    unsigned long long *cand_codes = nullptr;
    uint32_t *cand_counts = nullptr;
    size_t cand_cap = 0;
    // device scalars: [0]=cand_n [1]=global max count [2]=non-zero cells (num_unique_votes)
    // [3]=overflow flag ; votes_total (u64) separately
    uint32_t *scalars = nullptr;
    unsigned long long *votes_total = nullptr;
    // grouped vote kernel: work counters ([0] next reference point, [1 + r] next chunk of reference point r)
    uint32_t *sched = nullptr;
    size_t sched_cap = 0;
    uint32_t *acc_scratch = nullptr;          // accumulators of a dense scene's segments (ppf_vote_grouped.cu)
    size_t acc_scratch_cap = 0;
    uint2 *replay = nullptr;                  // deferred exact votes, [CTAs][cap] (ppf_vote_grouped.cu)
    size_t replay_cap = 0;
    // survivors, ordered (count desc, code asc) -- model.cu:155-170
    size_t K = 0;
    unsigned long long *codes = nullptr;
    uint32_t *counts = nullptr;
    float *transformations = nullptr;         // K*16
    float *weighted = nullptr;                // K
    float3 *trans = nullptr;                  // K
    float4 *rots = nullptr;                   // K
    float *scores = nullptr;                  // K  (vote_counts_out)
    uint32_t max_idx = 0;
    size_t cap_K = 0;
};

// cache of device blocks of destroyed clouds (ppf_model.cu): no cudaMalloc / cudaFree per Scene in steady state
void *pool_alloc(size_t bytes, size_t *cap_out);
void  pool_free(void *p, size_t cap);
void  pool_trim();
cudaError_t pooled_malloc_bytes(void **p, size_t bytes);
void  pooled_free(void *p);
template <typename T> inline cudaError_t pooled_malloc(T **p, size_t bytes) { return pooled_malloc_bytes((void **)p, bytes); }

// size of the scenes the next models will be matched against (0 = unknown): see ppf_set_expected_scene_points
extern std::atomic<int> g_expected_scene_points;

}  // namespace ppf
struct ppf_comm;
ppf::Comm *comm_impl(ppf_comm *c);            // ppf_comm.cu
namespace ppf {

// error plumbing (never exit(): SURVEY 8b "Errors")
void set_last_error(const std::string &msg);
#define PPF_CUDA_TRY(expr)                                                                 \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::ppf::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));     \
            return PPF_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

// The library works on its own NON-BLOCKING stream: one per (host thread, device), created on first use.  Every
// launch, async copy and CUB call of the calling thread goes to cur_stream(); an API call that hands results to the
// host synchronises it (memcpy_sync).  Nothing touches the legacy default stream (SURVEY 8b "Threading").
cudaStream_t cur_stream();
cudaError_t memcpy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);

// every launch of one of OUR kernels is counted (bench.py reports it as gpu_launches)
void count_launch(int n = 1);

// ---- entry points of the translation units -------------------------------------------
int  cloud_create(const float *xyz, int xs, const float *nrm, int ns, int n, int mem, Cloud &c, bool spatial_sort = false);
void cloud_free(Cloud &c);
int  features_tile(const Cloud &c, float d_dist, unsigned df, int rb, int re, int ob, int oe,
                   float *ppfs_host, uint32_t *keys_host);
int  model_build(ModelTable &m);
void model_free(ModelTable &m);
int  model_save(const ModelTable &m, const char *path);
int  model_load(ModelTable &m, const char *path);
int  model_far_cells(const ModelTable &m, const Cloud &scene, const unsigned long long **cells, const uint32_t **buckets,
                     int *n, int *kd_min, int *kd_max);
int  model_table_get(const ModelTable &m, uint32_t *hashkeys, size_t *counts, size_t *first, size_t *map);
int  vote_run(const ModelTable &m, const Cloud &scene, unsigned df, int shard_rank, int shard_count,
              int emit_all, VoteResult &r, unsigned long long *pairs_out, int *launches);
int  vote_finalize(const ModelTable &m, uint32_t global_max, int emit_all, VoteResult &r);
int  vote_finalize_dist(const ModelTable &m, Comm *comm, VoteResult &r, uint32_t *global_max_out);
int  cluster_dist(const ModelTable &m, Comm *comm, VoteResult &r);
int  order_survivors(VoteResult &r, size_t K, unsigned long long *codes_in, uint32_t *counts_in);
size_t order_survivors_bytes(size_t K);
int  vote_reserve_K(VoteResult &r, size_t K);
void vote_result_free(VoteResult &r);
int  poses_run(const ModelTable &m, const Cloud &scene, VoteResult &r);
int  cluster_run(const ModelTable &m, VoteResult &r, int shard = 0, int n_shards = 1);
int  cluster_finish(VoteResult &r);
int  voxel_grid_run(const float *xyz, int xs, const float *nrm, int ns, int n, int mem, float leaf, float *out_xyz,
                    float *out_nrm, int *n_out);
int  op_point_pair_feature(const float *p1, const float *n1, const float *p2, const float *n2, size_t n, float d_dist,
                           float *raw_out, float *disc_out, uint32_t *keys_out);
int  op_trans_model_scene(const float *m_r, const float *n_r_m, const float *m_i, const float *s_r, const float *n_r_s,
                          const float *s_i, size_t n, float *T_m_g, float *T_s_g, float *alpha, uint32_t *alpha_idx);

}  // namespace ppf
