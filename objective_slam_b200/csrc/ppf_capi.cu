// ppf_capi.cu -- the C ABI declared in include/ppf_b200.h (handles, error strings,
// staging of the lookup stages, and the ppf_registration drop-in boundary).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

namespace ppf {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
static std::atomic<unsigned long long> g_kernel_launches{0};
void count_launch(int n) { g_kernel_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

cudaStream_t cur_stream() {
    constexpr int kMaxDev = 64;
    static thread_local cudaStream_t streams[kMaxDev] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return cudaStreamPerThread;
    if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
        streams[dev] = nullptr;
        return cudaStreamPerThread;
    }
    return streams[dev];
}
cudaError_t memcpy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
    cudaStream_t s = cur_stream();
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, s);
    return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}
}  // namespace ppf

using namespace ppf;

struct ppf_scene { Cloud cloud; };
struct ppf_model { ModelTable table; };
struct ppf_lookup {
    VoteResult res;
    ppf_lookup_stats_t stats;
    int kernel_launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {
#define PPF_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) { set_last_error(msg); return PPF_ERR_INVALID; } \
    } while (0)

int read_vote_scalars(ppf_lookup *lk) {
    uint32_t h[4] = {0, 0, 0, 0};
    unsigned long long t[2] = {0, 0};
    if (lk->res.scalars) {
        PPF_CUDA_TRY(memcpy_sync(h, lk->res.scalars, sizeof(h), cudaMemcpyDeviceToHost));
        PPF_CUDA_TRY(memcpy_sync(t, lk->res.votes_total, sizeof(t), cudaMemcpyDeviceToHost));
    }
    lk->stats.max_vote_count = h[1];
    lk->stats.num_exact_alpha = h[3];
    lk->stats.num_nonunique_votes = t[0];
    lk->stats.num_unique_votes = t[1];
    return PPF_OK;
}
}  // namespace

extern "C" {

const char *ppf_last_error(void) { return g_last_error.c_str(); }
const char *ppf_version(void) { return "ppf_b200 0.1 (sm_100a)"; }
uint64_t ppf_kernel_launch_count(void) { return g_kernel_launches.load(); }
void ppf_release_cached_memory(void) { pool_trim(); }
void ppf_set_expected_scene_points(int n) { g_expected_scene_points.store(n > 0 ? n : 0); }
int ppf_current_stream(void **stream_out) {
    if (!stream_out) { set_last_error("current_stream: NULL argument"); return PPF_ERR_INVALID; }
    *stream_out = (void *)cur_stream();
    return PPF_OK;
}

// ---- Scene ------------------------------------------------------------------------
int ppf_scene_create(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n, int mem,
                     ppf_scene_t **out) {
    PPF_CHECK_ARG(out, "scene: out is NULL");
    *out = nullptr;
    ppf_scene *s = new ppf_scene();
    int rc = cloud_create(xyz, xyz_stride, nrm, nrm_stride, n, mem, s->cloud, /*spatial_sort=*/true);
    if (rc) { cloud_free(s->cloud); delete s; return rc; }
    *out = s;
    return PPF_OK;
}
void ppf_scene_destroy(ppf_scene_t *s) {
    if (!s) return;
    cloud_free(s->cloud);
    delete s;
}
int ppf_scene_num_points(const ppf_scene_t *s) { return s ? s->cloud.n : 0; }

int ppf_scene_features(const ppf_scene_t *s, float d_dist, unsigned df, int rb, int re, int ob, int oe,
                       float *ppfs_out, uint32_t *keys_out) {
    PPF_CHECK_ARG(s, "scene is NULL");
    return features_tile(s->cloud, d_dist, df, rb, re, ob, oe, ppfs_out, keys_out);
}

// ---- Model ------------------------------------------------------------------------
int ppf_model_create(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n, int mem,
                     float d_dist, float vote_count_threshold, int use_l1_norm, int use_averaged_clusters,
                     ppf_model_t **out) {
    PPF_CHECK_ARG(out, "model: out is NULL");
    *out = nullptr;
    ppf_model *m = new ppf_model();
    m->table.d_dist = d_dist;
    m->table.vote_count_threshold = vote_count_threshold;
    m->table.use_l1_norm = use_l1_norm;
    m->table.use_averaged_clusters = use_averaged_clusters;
    int rc = cloud_create(xyz, xyz_stride, nrm, nrm_stride, n, mem, m->table.cloud);
    if (!rc) rc = model_build(m->table);
    if (rc) { model_free(m->table); delete m; return rc; }
    *out = m;
    return PPF_OK;
}
void ppf_model_destroy(ppf_model_t *m) {
    if (!m) return;
    model_free(m->table);
    delete m;
}
int ppf_model_num_points(const ppf_model_t *m) { return m ? m->table.cloud.n : 0; }

int ppf_model_save(const ppf_model_t *m, const char *path) {
    PPF_CHECK_ARG(m && path, "model save: NULL argument");
    return model_save(m->table, path);
}
int ppf_model_load(const char *path, ppf_model_t **out) {
    PPF_CHECK_ARG(out && path, "model load: NULL argument");
    *out = nullptr;
    ppf_model *m = new ppf_model();
    int rc = model_load(m->table, path);
    if (rc) { model_free(m->table); delete m; return rc; }
    *out = m;
    return PPF_OK;
}
int ppf_model_layout(const ppf_model_t *m, int *n_chunks, int *chunk_rows, int *grouped_kernel) {
    PPF_CHECK_ARG(m, "model is NULL");
    if (n_chunks) *n_chunks = m->table.n_chunks;
    if (chunk_rows) *chunk_rows = m->table.chunk_rows;
    if (grouped_kernel) *grouped_kernel = m->table.prefer_grouped;
    return PPF_OK;
}

int ppf_model_params(const ppf_model_t *m, float *d_dist, float *vote_count_threshold, int *use_l1_norm,
                     int *use_averaged_clusters) {
    PPF_CHECK_ARG(m, "model is NULL");
    if (d_dist) *d_dist = m->table.d_dist;
    if (vote_count_threshold) *vote_count_threshold = m->table.vote_count_threshold;
    if (use_l1_norm) *use_l1_norm = m->table.use_l1_norm;
    if (use_averaged_clusters) *use_averaged_clusters = m->table.use_averaged_clusters;
    return PPF_OK;
}

int ppf_model_table_sizes(const ppf_model_t *m, size_t *U, size_t *npairs) {
    PPF_CHECK_ARG(m, "model is NULL");
    if (U) *U = m->table.U;
    if (npairs) *npairs = (size_t)m->table.cloud.n * m->table.cloud.n;
    return PPF_OK;
}
int ppf_model_table_get(const ppf_model_t *m, uint32_t *hashkeys, size_t *counts, size_t *first, size_t *map) {
    PPF_CHECK_ARG(m, "model is NULL");
    return model_table_get(m->table, hashkeys, counts, first, map);
}
int ppf_model_features(const ppf_model_t *m, int rb, int re, int ob, int oe, float *ppfs_out, uint32_t *keys_out) {
    PPF_CHECK_ARG(m, "model is NULL");
    return features_tile(m->table.cloud, m->table.d_dist, 1, rb, re, ob, oe, ppfs_out, keys_out);
}

// ---- pre-processing --------------------------------------------------------------------
int ppf_voxel_grid(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n, int mem, float leaf,
                   float *out_xyz, float *out_nrm, int *n_out) {
    return voxel_grid_run(xyz, xyz_stride, nrm, nrm_stride, n, mem, leaf, out_xyz, out_nrm, n_out);
}

// ---- operator-level entry points -----------------------------------------------------
int ppf_point_pair_feature(const float *p1, const float *n1, const float *p2, const float *n2, size_t n, float d_dist,
                           float *raw_out, float *disc_out, uint32_t *keys_out) {
    return op_point_pair_feature(p1, n1, p2, n2, n, d_dist, raw_out, disc_out, keys_out);
}
int ppf_trans_model_scene(const float *m_r, const float *n_r_m, const float *m_i, const float *s_r, const float *n_r_s,
                          const float *s_i, size_t n, float *T_m_g, float *T_s_g, float *alpha, uint32_t *alpha_idx) {
    return op_trans_model_scene(m_r, n_r_m, m_i, s_r, n_r_s, s_i, n, T_m_g, T_s_g, alpha, alpha_idx);
}

// ---- Lookup -----------------------------------------------------------------------
int ppf_lookup_create(ppf_lookup_t **out) {
    PPF_CHECK_ARG(out, "lookup: out is NULL");
    ppf_lookup *lk = new ppf_lookup();
    std::memset(&lk->stats, 0, sizeof(lk->stats));
    for (auto &e : lk->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { set_last_error("cudaEventCreate failed"); delete lk; return PPF_ERR_CUDA; }
    *out = lk;
    return PPF_OK;
}
void ppf_lookup_destroy(ppf_lookup_t *lk) {
    if (!lk) return;
    vote_result_free(lk->res);
    for (auto &e : lk->ev) if (e) cudaEventDestroy(e);
    delete lk;
}

int ppf_lookup_vote(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, int shard_rank, int shard_count,
                    ppf_lookup_t *lk) {
    PPF_CHECK_ARG(m && s && lk, "vote: NULL handle");
    std::memset(&lk->stats, 0, sizeof(lk->stats));
    unsigned long long pairs = 0;
    cudaEventRecord(lk->ev[0], cur_stream());
    int rc = vote_run(m->table, s->cloud, df, shard_rank, shard_count, 0, lk->res, &pairs, &lk->kernel_launches);
    cudaEventRecord(lk->ev[1], cur_stream());
    if (rc) return rc;
    PPF_CUDA_TRY(cudaEventSynchronize(lk->ev[1]));
    cudaEventElapsedTime(&lk->stats.ms_vote, lk->ev[0], lk->ev[1]);
    lk->stats.num_scene_pairs = pairs;
    return read_vote_scalars(lk);
}

int ppf_lookup_local_max(const ppf_lookup_t *lk, uint32_t *max_count) {
    PPF_CHECK_ARG(lk && max_count, "local_max: NULL argument");
    *max_count = lk->stats.max_vote_count;
    return PPF_OK;
}

int ppf_lookup_finalize(const ppf_model_t *m, uint32_t global_max, ppf_lookup_t *lk) {
    PPF_CHECK_ARG(m && lk, "finalize: NULL handle");
    cudaEventRecord(lk->ev[1], cur_stream());
    int rc = vote_finalize(m->table, global_max, 0, lk->res);
    cudaEventRecord(lk->ev[2], cur_stream());
    if (rc) return rc;
    PPF_CUDA_TRY(cudaEventSynchronize(lk->ev[2]));
    cudaEventElapsedTime(&lk->stats.ms_finalize, lk->ev[1], lk->ev[2]);
    lk->stats.max_vote_count = global_max;
    lk->stats.num_top_votes = (uint32_t)lk->res.K;
    return PPF_OK;
}

int ppf_lookup_survivors(const ppf_lookup_t *lk, size_t *K, const uint64_t **codes_dev, const uint32_t **counts_dev) {
    PPF_CHECK_ARG(lk && K, "survivors: NULL argument");
    *K = lk->res.K;
    if (codes_dev) *codes_dev = (const uint64_t *)lk->res.codes;
    if (counts_dev) *counts_dev = lk->res.counts;
    return PPF_OK;
}

int ppf_lookup_copy_survivors(const ppf_lookup_t *lk, uint64_t *codes_dst_dev, uint32_t *counts_dst_dev) {
    PPF_CHECK_ARG(lk && (lk->res.K == 0 || (codes_dst_dev && counts_dst_dev)), "copy_survivors: NULL argument");
    if (lk->res.K == 0) return PPF_OK;
    PPF_CUDA_TRY(memcpy_sync(codes_dst_dev, lk->res.codes, lk->res.K * 8, cudaMemcpyDeviceToDevice));
    PPF_CUDA_TRY(memcpy_sync(counts_dst_dev, lk->res.counts, lk->res.K * 4, cudaMemcpyDeviceToDevice));
    return PPF_OK;
}

int ppf_lookup_set_survivors(ppf_lookup_t *lk, const uint64_t *codes_dev, const uint32_t *counts_dev, size_t K) {
    PPF_CHECK_ARG(lk && (K == 0 || (codes_dev && counts_dev)), "set_survivors: NULL argument");
    unsigned long long *c = nullptr; uint32_t *n = nullptr;
    if (K) {
        int rc0 = lk->res.ws.reserve(K * 12 + order_survivors_bytes(K));
        if (rc0) return rc0;
        c = lk->res.ws.take<unsigned long long>(K);
        n = lk->res.ws.take<uint32_t>(K);
        PPF_CUDA_TRY(cudaMemcpyAsync(c, codes_dev, K * 8, cudaMemcpyDeviceToDevice, cur_stream()));
        PPF_CUDA_TRY(cudaMemcpyAsync(n, counts_dev, K * 4, cudaMemcpyDeviceToDevice, cur_stream()));
    }
    int rc = order_survivors(lk->res, K, c, n);
    lk->stats.num_top_votes = (uint32_t)lk->res.K;
    return rc;
}

int ppf_lookup_poses(const ppf_model_t *m, const ppf_scene_t *s, ppf_lookup_t *lk) {
    PPF_CHECK_ARG(m && s && lk, "poses: NULL handle");
    cudaEventRecord(lk->ev[2], cur_stream());
    return poses_run(m->table, s->cloud, lk->res);
}

int ppf_lookup_cluster(const ppf_model_t *m, ppf_lookup_t *lk) {
    PPF_CHECK_ARG(m && lk, "cluster: NULL handle");
    int rc = cluster_run(m->table, lk->res);
    cudaEventRecord(lk->ev[3], cur_stream());
    if (rc) return rc;
    PPF_CUDA_TRY(cudaEventSynchronize(lk->ev[3]));
    cudaEventElapsedTime(&lk->stats.ms_pose_cluster, lk->ev[2], lk->ev[3]);
    lk->stats.max_idx = lk->res.max_idx;
    return PPF_OK;
}

// Multi-GPU: the clustering of the merged survivor list is quadratic in dense cells (6 ms at K = 50k, 332 ms at
// K = 396k), so every rank scores an interleaved slice, the slices are summed (all other entries are 0: exact)
// and ppf_lookup_cluster_finish picks the winner.
int ppf_lookup_cluster_shard(const ppf_model_t *m, ppf_lookup_t *lk, int shard, int n_shards) {
    PPF_CHECK_ARG(m && lk && n_shards >= 1 && shard >= 0 && shard < n_shards, "cluster_shard: bad argument");
    PPF_CHECK_ARG(!m->table.use_averaged_clusters || n_shards == 1,
                  "cluster_shard: use_averaged_clusters needs every pose's averaged translation: cluster unsharded");
    return cluster_run(m->table, lk->res, shard, n_shards);
}
int ppf_lookup_copy_scores(const ppf_lookup_t *lk, float *scores_dst_dev) {
    PPF_CHECK_ARG(lk && (lk->res.K == 0 || scores_dst_dev), "copy_scores: NULL argument");
    if (lk->res.K) PPF_CUDA_TRY(memcpy_sync(scores_dst_dev, lk->res.scores, lk->res.K * 4, cudaMemcpyDeviceToDevice));
    return PPF_OK;
}
int ppf_lookup_set_scores(ppf_lookup_t *lk, const float *scores_src_dev) {
    PPF_CHECK_ARG(lk && (lk->res.K == 0 || scores_src_dev), "set_scores: NULL argument");
    if (lk->res.K) PPF_CUDA_TRY(memcpy_sync(lk->res.scores, scores_src_dev, lk->res.K * 4, cudaMemcpyDeviceToDevice));
    return PPF_OK;
}
int ppf_lookup_cluster_finish(ppf_lookup_t *lk) {
    PPF_CHECK_ARG(lk, "cluster_finish: NULL handle");
    int rc = cluster_finish(lk->res);
    cudaEventRecord(lk->ev[3], cur_stream());
    if (rc) return rc;
    PPF_CUDA_TRY(cudaEventSynchronize(lk->ev[3]));
    cudaEventElapsedTime(&lk->stats.ms_pose_cluster, lk->ev[2], lk->ev[3]);
    lk->stats.max_idx = lk->res.max_idx;
    return PPF_OK;
}

// Whole Model::ppf_lookup (model.cu:269-306) on the library's stream, alone or as rank comm->rank of comm->world
// ranks that share the scene (reference points sharded, everything after the vote on the merged survivor list).
// Host synchronisations: after the vote kernel (candidate overflow check), after the filter (K sizes the sorts), after
// the count exchange (sharded only), and when the winner is read.
// gpu_clustering = false: stop after the poses (the caller clusters on the host; Model::ppf_lookup with
// cpu_clustering set never runs ClusterTransformations, model.cu:284-291)
static int lookup_run(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, ppf_lookup_t *lk, bool gpu_clustering,
                      Comm *comm) {
    PPF_CHECK_ARG(m && s && lk, "lookup: NULL handle");
    const int rank = comm ? comm->rank : 0, world = comm ? comm->world : 1;
    int rc = ppf_lookup_vote(m, s, df, rank, world, lk);
    if (rc) return rc;
    cudaEventRecord(lk->ev[1], cur_stream());
    uint32_t gmax = 0;
    if ((rc = vote_finalize_dist(m->table, comm, lk->res, &gmax))) return rc;
    cudaEventRecord(lk->ev[2], cur_stream());
    lk->stats.max_vote_count = gmax;
    lk->stats.num_top_votes = (uint32_t)lk->res.K;
    if ((rc = poses_run(m->table, s->cloud, lk->res))) return rc;
    if (gpu_clustering) {
        if ((rc = cluster_dist(m->table, comm, lk->res))) return rc;
        lk->stats.max_idx = lk->res.max_idx;
    }
    cudaEventRecord(lk->ev[3], cur_stream());
    PPF_CUDA_TRY(cudaEventSynchronize(lk->ev[3]));
    cudaEventElapsedTime(&lk->stats.ms_finalize, lk->ev[1], lk->ev[2]);
    cudaEventElapsedTime(&lk->stats.ms_pose_cluster, lk->ev[2], lk->ev[3]);
    if (gmax == 0) { set_last_error("lookup: no scene pair matched the model"); return PPF_ERR_NO_VOTES; }
    return PPF_OK;
}
int ppf_model_lookup(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, ppf_lookup_t *lk) {
    return lookup_run(m, s, df, lk, true, nullptr);
}
int ppf_model_lookup_sharded(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, ppf_comm_t *comm, ppf_lookup_t *lk) {
    return lookup_run(m, s, df, lk, true, comm_impl(comm));
}

int ppf_lookup_get_stats(const ppf_lookup_t *lk, ppf_lookup_stats_t *stats) {
    PPF_CHECK_ARG(lk && stats, "stats: NULL argument");
    *stats = lk->stats;
    return PPF_OK;
}

int ppf_lookup_get(const ppf_lookup_t *lk, uint64_t *votes, uint32_t *counts, float *transformations,
                   float *weighted, float *trans, float *rots, float *scores, float *pose) {
    PPF_CHECK_ARG(lk, "get: NULL handle");
    const VoteResult &r = lk->res;
    size_t K = r.K;
    if (pose) std::memset(pose, 0, 16 * sizeof(float));
    if (K == 0) return PPF_OK;
    if (votes) PPF_CUDA_TRY(memcpy_sync(votes, r.codes, K * 8, cudaMemcpyDeviceToHost));
    if (counts) PPF_CUDA_TRY(memcpy_sync(counts, r.counts, K * 4, cudaMemcpyDeviceToHost));
    if (transformations) PPF_CUDA_TRY(memcpy_sync(transformations, r.transformations, K * 64, cudaMemcpyDeviceToHost));
    if (weighted) PPF_CUDA_TRY(memcpy_sync(weighted, r.weighted, K * 4, cudaMemcpyDeviceToHost));
    if (trans) PPF_CUDA_TRY(memcpy_sync(trans, r.trans, K * 12, cudaMemcpyDeviceToHost));
    if (rots) PPF_CUDA_TRY(memcpy_sync(rots, r.rots, K * 16, cudaMemcpyDeviceToHost));
    if (scores) PPF_CUDA_TRY(memcpy_sync(scores, r.scores, K * 4, cudaMemcpyDeviceToHost));
    if (pose) {
        // ppf.cu:80-93 -- only the winning pose crosses the bus (the reference copies all K)
        float t[3];
        PPF_CUDA_TRY(memcpy_sync(pose, r.transformations + (size_t)r.max_idx * 16, 64, cudaMemcpyDeviceToHost));
        PPF_CUDA_TRY(memcpy_sync(t, r.trans + r.max_idx, 12, cudaMemcpyDeviceToHost));
        pose[3] = t[0]; pose[7] = t[1]; pose[11] = t[2];
    }
    return PPF_OK;
}

int ppf_vote_histogram(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, uint64_t *codes_out,
                       uint32_t *counts_out, size_t capacity, size_t *n_out) {
    return ppf_vote_histogram_shard(m, s, df, 0, 1, codes_out, counts_out, capacity, n_out);
}

int ppf_vote_histogram_shard(const ppf_model_t *m, const ppf_scene_t *s, unsigned df, int shard_rank, int shard_count,
                             uint64_t *codes_out, uint32_t *counts_out, size_t capacity, size_t *n_out) {
    PPF_CHECK_ARG(m && s && n_out, "histogram: NULL argument");
    VoteResult r;
    int rc = vote_run(m->table, s->cloud, df, shard_rank, shard_count, 1, r, nullptr, nullptr);
    uint32_t h[4] = {0, 0, 0, 0};
    if (!rc && r.scalars) {
        if (memcpy_sync(h, r.scalars, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) rc = PPF_ERR_CUDA;
    }
    if (!rc) {
        // ascending code order == the order thrust::sort + histogram() leaves (model.cu:148-151)
        *n_out = h[0];
        if (h[0] && codes_out && counts_out && h[0] <= capacity) {
            std::vector<unsigned long long> c(h[0]);
            std::vector<uint32_t> n(h[0]);
            memcpy_sync(c.data(), r.cand_codes, (size_t)h[0] * 8, cudaMemcpyDeviceToHost);
            memcpy_sync(n.data(), r.cand_counts, (size_t)h[0] * 4, cudaMemcpyDeviceToHost);
            std::vector<uint32_t> order(h[0]);
            for (uint32_t i = 0; i < h[0]; i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return c[a] < c[b]; });
            for (uint32_t i = 0; i < h[0]; i++) { codes_out[i] = c[order[i]]; counts_out[i] = n[order[i]]; }
        }
    }
    vote_result_free(r);
    return rc;
}

}  // extern "C"

// ---- PCL-style greedy pose clustering on the host (cpu_clustering = true) ----------------
// Follows clusterPoses / posesWithinErrorBounds (transformation_clustering.cpp:62-137) and
// Model::ClusterTransformationsCPU (model.cu:246-266).  K is 10^2..10^5 and the algorithm
// is inherently sequential (a pose joins the first cluster whose SEED is close), so it
// runs on one host core exactly like the reference.  Eigen is not available here; the
// angle-axis angle of R1^T R2 is computed through the quaternion, as Eigen does.
// PARITY UNPINNED: the reference's variant needs Eigen + PCL (neither vendored in /root/reference nor installed), so
// this option cannot be compared with it here; tests/ only check that it agrees with the GPU clustering's winner
// within the reference's own acceptance gate (0.1 x diameter, 12 degrees) and that it is deterministic.
namespace {
struct Pose { float R[3][3]; float t[3]; unsigned votes; };

void quat_from_rot(const float R[3][3], float q[4]) {   // (x, y, z, w), Eigen's Shepperd variant
    float tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > 0.f) {
        float s = std::sqrt(tr + 1.0f);
        q[3] = 0.5f * s; s = 0.5f / s;
        q[0] = (R[2][1] - R[1][2]) * s; q[1] = (R[0][2] - R[2][0]) * s; q[2] = (R[1][0] - R[0][1]) * s;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        float s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0f);
        q[i] = 0.5f * s; s = 0.5f / s;
        q[3] = (R[k][j] - R[j][k]) * s; q[j] = (R[j][i] + R[i][j]) * s; q[k] = (R[k][i] + R[i][k]) * s;
    }
}
float rotation_angle_between(const Pose &a, const Pose &b) {
    float D[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) D[i][j] = a.R[0][i] * b.R[0][j] + a.R[1][i] * b.R[1][j] + a.R[2][i] * b.R[2][j];
    float q[4];
    quat_from_rot(D, q);
    float n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    return std::fabs(2.0f * std::atan2(n, std::fabs(q[3])));
}
bool within_bounds(const Pose &a, const Pose &b, float tt, float rt) {
    float dx = a.t[0] - b.t[0], dy = a.t[1] - b.t[1], dz = a.t[2] - b.t[2];
    return std::sqrt(dx * dx + dy * dy + dz * dz) < tt && rotation_angle_between(a, b) < rt;
}
// returns the best cluster's averaged pose (row-major 4x4)
bool cluster_poses_cpu(std::vector<Pose> &poses, float tt, float rt, float out[16]) {
    std::stable_sort(poses.begin(), poses.end(), [](const Pose &a, const Pose &b) { return a.votes > b.votes; });
    std::vector<std::vector<int>> clusters;
    std::vector<std::pair<size_t, unsigned>> cv;
    for (size_t i = 0; i < poses.size(); i++) {
        bool found = false;
        for (size_t c = 0; c < clusters.size(); c++)
            if (within_bounds(poses[i], poses[clusters[c][0]], tt, rt)) {
                clusters[c].push_back((int)i); cv[c].second += poses[i].votes; found = true; break;
            }
        if (!found) { clusters.push_back({(int)i}); cv.push_back({clusters.size() - 1, poses[i].votes}); }
    }
    if (clusters.empty()) return false;
    std::stable_sort(cv.begin(), cv.end(), [](const std::pair<size_t, unsigned> &a, const std::pair<size_t, unsigned> &b) { return a.second > b.second; });
    const std::vector<int> &best = clusters[cv[0].first];
    float ta[3] = {0, 0, 0}, qa[4] = {0, 0, 0, 0};
    for (int i : best) {
        float q[4];
        quat_from_rot(poses[i].R, q);
        for (int k = 0; k < 3; k++) ta[k] += poses[i].t[k];
        for (int k = 0; k < 4; k++) qa[k] += q[k];
    }
    float inv = 1.0f / (float)best.size();
    for (int k = 0; k < 3; k++) ta[k] *= inv;
    float qn = 0;
    for (int k = 0; k < 4; k++) { qa[k] *= inv; qn += qa[k] * qa[k]; }
    qn = std::sqrt(qn);
    float x = qa[0] / qn, y = qa[1] / qn, z = qa[2] / qn, w = qa[3] / qn;
    float Rm[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                      {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                      {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out[i * 4 + j] = Rm[i][j]; out[i * 4 + 3] = ta[i]; }
    out[12] = out[13] = out[14] = 0; out[15] = 1;
    return true;
}
}  // namespace

extern "C" int ppf_lookup_cluster_cpu(const ppf_model_t *m, ppf_lookup_t *lk, float *pose_out) {
    PPF_CHECK_ARG(m && lk && pose_out, "cluster_cpu: NULL argument");
    size_t K = lk->res.K;
    std::memset(pose_out, 0, 64);
    if (!K) return PPF_OK;
    std::vector<float> T(K * 16);
    std::vector<uint32_t> cnt(K);
    PPF_CUDA_TRY(memcpy_sync(T.data(), lk->res.transformations, K * 64, cudaMemcpyDeviceToHost));
    PPF_CUDA_TRY(memcpy_sync(cnt.data(), lk->res.counts, K * 4, cudaMemcpyDeviceToHost));
    std::vector<Pose> poses(K);
    for (size_t i = 0; i < K; i++) {
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) poses[i].R[r][c] = T[i * 16 + r * 4 + c]; poses[i].t[r] = T[i * 16 + r * 4 + 3]; }
        poses[i].votes = cnt[i];
    }
    cluster_poses_cpu(poses, m->table.d_dist, d_angle0(), pose_out);
    return PPF_OK;
}

// ---- the drop-in boundary: ppf_registration (ppf.h:9-15, ppf.cu:29-106) -------------------
static int registration_run(const ppf_cloud_t *scene_clouds, int num_scenes, const ppf_cloud_t *model_clouds,
                            int num_models, const float *model_d_dists, unsigned ref_point_downsample_factor,
                            float vote_count_threshold, int cpu_clustering, int use_l1_norm,
                            int use_averaged_clusters, int device, ppf_comm_t *comm, float *poses_out, int *status_out) {
    PPF_CHECK_ARG(scene_clouds && model_clouds && model_d_dists && poses_out && num_scenes >= 0 && num_models >= 0,
                  "registration: NULL argument");
    int ndev = 0;
    PPF_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_last_error("registration: no CUDA device"); return PPF_ERR_CUDA; }
    int caller_device = 0;
    PPF_CUDA_TRY(cudaGetDevice(&caller_device));
    // sharded: the rank's communicator lives on the device that is current; alone: the reference's rule (ppf.cu:45)
    if (!comm) PPF_CUDA_TRY(cudaSetDevice(std::min(ndev - 1, std::max(device, 0))));
    struct RestoreDevice { int d; ~RestoreDevice() { cudaSetDevice(d); } } restore_device{caller_device};
    std::memset(poses_out, 0, (size_t)num_scenes * num_models * 64);
    // The reference rebuilds Scene and Model for every (scene, model) pair (ppf.cu:63-70). The model
    // table depends only on (model, d_dist), so it is built once per model and reused across scenes;
    // the uploaded scene cloud does not depend on d_dist at all and is reused across models.
    std::vector<ppf_model_t *> models(num_models, nullptr);
    int rc = PPF_OK;
    // the scenes are known before the model tables are laid out: tell the builder how dense they are
    const int hint_before = g_expected_scene_points.load();
    int largest_scene = 0;
    for (int i = 0; i < num_scenes; i++) largest_scene = std::max(largest_scene, scene_clouds[i].n);
    g_expected_scene_points.store(largest_scene);
    struct RestoreHint { int v; ~RestoreHint() { g_expected_scene_points.store(v); } } restore_hint{hint_before};
    for (int j = 0; j < num_models && !rc; j++) {
        const ppf_cloud_t &c = model_clouds[j];
        rc = ppf_model_create(c.xyz, c.xyz_stride, c.nrm, c.nrm_stride, c.n, PPF_MEM_HOST, model_d_dists[j],
                              vote_count_threshold, use_l1_norm, use_averaged_clusters, &models[j]);
    }
    ppf_lookup_t *lk = nullptr;
    if (!rc) rc = ppf_lookup_create(&lk);
    for (int i = 0; i < num_scenes && !rc; i++) {
        const ppf_cloud_t &c = scene_clouds[i];
        ppf_scene_t *scene = nullptr;
        rc = ppf_scene_create(c.xyz, c.xyz_stride, c.nrm, c.nrm_stride, c.n, PPF_MEM_HOST, &scene);
        for (int j = 0; j < num_models && !rc; j++) {
            float *pose = poses_out + ((size_t)i * num_models + j) * 16;
            int st = lookup_run(models[j], scene, ref_point_downsample_factor, lk, !cpu_clustering, comm_impl(comm));
            if (st == PPF_OK) {
                if (cpu_clustering) st = ppf_lookup_cluster_cpu(models[j], lk, pose);      // ppf.cu:75-77
                else st = ppf_lookup_get(lk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pose);
            }
            if (status_out) status_out[(size_t)i * num_models + j] = st;
            if (st != PPF_OK && st != PPF_ERR_NO_VOTES) rc = st;
        }
        ppf_scene_destroy(scene);
    }
    ppf_lookup_destroy(lk);
    for (auto *m : models) ppf_model_destroy(m);
    return rc;
}

extern "C" int ppf_registration(const ppf_cloud_t *scene_clouds, int num_scenes, const ppf_cloud_t *model_clouds,
                                int num_models, const float *model_d_dists, unsigned ref_point_downsample_factor,
                                float vote_count_threshold, int cpu_clustering, int use_l1_norm,
                                int use_averaged_clusters, int device, const float *model_weights,
                                float *poses_out, int *status_out) {
    (void)model_weights;                                           // ignored by the reference too (ppf.cu:35)
    return registration_run(scene_clouds, num_scenes, model_clouds, num_models, model_d_dists, ref_point_downsample_factor,
                            vote_count_threshold, cpu_clustering, use_l1_norm, use_averaged_clusters, device, nullptr,
                            poses_out, status_out);
}

// The same call made by every rank of `comm` with the same arguments (one process or host thread per GPU, the rank's
// device current): the scene reference points are sharded over the ranks, every rank returns the same poses.
extern "C" int ppf_registration_sharded(const ppf_cloud_t *scene_clouds, int num_scenes, const ppf_cloud_t *model_clouds,
                                        int num_models, const float *model_d_dists,
                                        unsigned ref_point_downsample_factor, float vote_count_threshold,
                                        int cpu_clustering, int use_l1_norm, int use_averaged_clusters,
                                        ppf_comm_t *comm, float *poses_out, int *status_out) {
    PPF_CHECK_ARG(comm, "registration_sharded: comm is NULL");
    return registration_run(scene_clouds, num_scenes, model_clouds, num_models, model_d_dists, ref_point_downsample_factor,
                            vote_count_threshold, cpu_clustering, use_l1_norm, use_averaged_clusters, 0, comm,
                            poses_out, status_out);
}
