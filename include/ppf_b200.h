/* ppf_b200.h -- C ABI of libppf_b200.so: the Drost point-pair-feature recognition
 * hot path of nicolasavru/objective-slam, rebuilt for NVIDIA B200 (sm_100a).
 *
 * Every entry point replaces one piece of the reference's C++/CUDA operator
 * surface (paths relative to /root/reference/pcl/alignment):
 *
 *   ppf_registration          <- ppf_registration()              include/ppf.h:9-15, src/cuda/ppf.cu:29-106
 *   ppf_scene_create          <- Scene::Scene()                   include/scene.h:14-16, src/cuda/scene.cu:24-55
 *   ppf_scene_features        <- Scene::getModelPPFs/getHashKeys  include/scene.h:21-24 (debug/parity view)
 *   ppf_model_create          <- Model::Model()                   include/model.h:17-19, src/cuda/model.cu:43-82
 *   ppf_model_table_*         <- ParallelHashArray::Get*()        include/impl/parallel_hash_array.hpp:25-30
 *   ppf_model_lookup          <- Model::ppf_lookup()              include/model.h:35, src/cuda/model.cu:269-306
 *   ppf_lookup_vote           <- Model::ComputeUniqueVotes()      src/cuda/model.cu:95-171 (accumulate part)
 *   ppf_lookup_finalize       <- Model::ComputeUniqueVotes()      src/cuda/model.cu:155-170 (threshold + order)
 *   ppf_lookup_poses          <- Model::ComputeTransformations() + ComputeWeightedVoteCounts()
 *                                                                 src/cuda/model.cu:173-200
 *   ppf_lookup_cluster        <- Model::ClusterTransformations() + max_element
 *                                                                 src/cuda/model.cu:202-244, 293-295
 *   ppf_lookup_get            <- public members votes, voteCounts, transformations, transformation_trans,
 *                                transformation_rots, vote_counts_out, max_idx   include/model.h:63-113
 *
 * Conventions: plain pointers and sizes only; every function returns a PPF_*
 * status (never terminates the process, unlike HANDLE_ERROR, util.hpp:18-26);
 * ppf_last_error() describes the last failure on the calling thread.  A handle
 * must not be used from two host threads at once.  Work is issued on the CUDA
 * device that is current when the handle is created.
 *
 * Clouds are given as two strided float arrays: point i has its position at
 * xyz[i*xyz_stride + 0..2] and its normal at nrm[i*nrm_stride + 0..2] (strides in
 * floats).  Dense N x 3 arrays use stride 3; an array of pcl::PointNormal (48 B:
 * x y z _ nx ny nz _ curvature _ _ _) uses xyz=&p[0].x, nrm=&p[0].normal_x,
 * stride 12 for both.  mem = PPF_MEM_HOST or PPF_MEM_DEVICE says where they live.
 */
#ifndef PPF_B200_H
#define PPF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PPF_OK = 0,
    PPF_ERR_INVALID = 1,      /* bad argument */
    PPF_ERR_CUDA = 2,         /* CUDA runtime failure (see ppf_last_error) */
    PPF_ERR_UNSUPPORTED = 3,  /* size outside the supported range */
    PPF_ERR_NO_VOTES = 4      /* no scene pair matched the model: no pose */
};

enum { PPF_MEM_HOST = 0, PPF_MEM_DEVICE = 1 };

#define PPF_N_ANGLE 30        /* kernel.h:15 */
#define PPF_MAX_MODEL_POINTS 46340   /* N*N pair indices stay below 2^31, as in the reference */

typedef struct ppf_scene ppf_scene_t;
typedef struct ppf_model ppf_model_t;
typedef struct ppf_lookup ppf_lookup_t;
typedef struct ppf_comm ppf_comm_t;      /* the ranks of a sharded recognition (one GPU each) */

const char *ppf_last_error(void);
const char *ppf_version(void);
/* Number of kernels of this library launched by the process so far (diagnostic; bench.py's gpu_launches). */
uint64_t ppf_kernel_launch_count(void);
/* Destroyed scenes / models park their device block in a small cache (<= 64 blocks, <= 6 GB) so that a
 * recognition loop does not call cudaMalloc / cudaFree per frame; this returns the cached blocks to the driver. */
void ppf_release_cached_memory(void);
/* Optional hint for ppf_model_create: how many points the scenes have that the next models will be matched against
 * (0 = unknown, the default).  It only selects the layout of the table / the vote kernel (scenes of >= 40k
 * points favour the grouped kernel even for small models); results never depend on it.  ppf_registration sets
 * it from its own scene list. */
void ppf_set_expected_scene_points(int n);
/* The library issues all its work on its own non-blocking CUDA stream, one per (host thread, device), never on the
 * legacy default stream (the reference uses the default stream and cudaDeviceSynchronize throughout).  Device
 * inputs must be complete before a call; results handed to the host are synchronised by the call.  This returns
 * the calling thread's stream on the current device (a cudaStream_t) so that a host can record events on it or
 * order its own work against it. */
int ppf_current_stream(void **stream_out);

/* ---- Scene ------------------------------------------------------------------ */
int ppf_scene_create(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n,
                     int mem, ppf_scene_t **out);
void ppf_scene_destroy(ppf_scene_t *scene);
int ppf_scene_num_points(const ppf_scene_t *scene);

/* Quantised features (float4 per ordered pair) and 32-bit hash keys of the tile
 * [ref_begin, ref_end) x [other_begin, other_end) of the scene's N x N pair matrix,
 * exactly as ppf_kernel + ppf_hash_kernel (kernel.cu:404-477) would have stored
 * them at out[ref*N + other]: rows with ref % ref_point_downsample_factor != 0 and
 * self pairs hold (NaN,0,0,0) / key 0.  Outputs are host arrays of
 * (ref_end-ref_begin)*(other_end-other_begin) elements; either may be NULL. */
int ppf_scene_features(const ppf_scene_t *scene, float d_dist, unsigned ref_point_downsample_factor,
                       int ref_begin, int ref_end, int other_begin, int other_end,
                       float *ppfs_out, uint32_t *keys_out);

/* ---- Model ------------------------------------------------------------------ */
int ppf_model_create(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n,
                     int mem, float d_dist, float vote_count_threshold, int use_l1_norm,
                     int use_averaged_clusters, ppf_model_t **out);
void ppf_model_destroy(ppf_model_t *model);
int ppf_model_num_points(const ppf_model_t *model);
/* Persistent model database (the reference rebuilds the table for every (scene, model) pair,
 * ppf.cu:63-70): the built table -- cloud, frames, ParallelHashArray arrays, vote payload, cell table, options --
 * in one little-endian file.  A loaded model is indistinguishable from the one that was saved. */
int ppf_model_save(const ppf_model_t *model, const char *path);
int ppf_model_load(const char *path, ppf_model_t **out);
/* How the table is laid out for voting: accumulator chunks, model points per chunk, and which vote kernel serves
 * it (1 = grouped, ppf_vote_grouped.cu; 0 = one hit per warp pass, ppf_vote.cu).  Any output may be NULL. */
int ppf_model_layout(const ppf_model_t *model, int *n_chunks, int *chunk_rows, int *grouped_kernel);

/* The options a model handle carries (the reference's Model members, model.h:43-46), e.g. after ppf_model_load.
 * Any output may be NULL. */
int ppf_model_params(const ppf_model_t *model, float *d_dist, float *vote_count_threshold, int *use_l1_norm,
                     int *use_averaged_clusters);

/* ParallelHashArray contents: U unique sorted keys, counts, first index, and the
 * N*N-long key-ordered map of pair indices (m_r*N + m_i).  size_t-typed like the
 * reference's device_vector<std::size_t>. Any output may be NULL. */
int ppf_model_table_sizes(const ppf_model_t *model, size_t *num_unique_keys, size_t *num_pairs);
int ppf_model_table_get(const ppf_model_t *model, uint32_t *hashkeys, size_t *counts,
                        size_t *first_index, size_t *key_to_pair);
/* Model-side tile of quantised features / keys (same layout rules as ppf_scene_features, df = 1). */
int ppf_model_features(const ppf_model_t *model, int ref_begin, int ref_end, int other_begin,
                       int other_end, float *ppfs_out, uint32_t *keys_out);

/* ---- Pre-processing: voxel-grid downsample (the caller right before the hot path) ---- */
/* voxelGridDownsample (alignment.cpp:79-87) = pcl::VoxelGrid<PointNormal> with a cubic leaf: one output point
 * per occupied leaf, in ascending leaf order, = centroid of the positions AND of the (un-normalised) normals
 * of the points in it.  out_xyz / out_nrm: dense n_out x 3 arrays with room for n points, in the memory
 * space `mem` names (like the inputs). */
int ppf_voxel_grid(const float *xyz, int xyz_stride, const float *nrm, int nrm_stride, int n, int mem, float leaf,
                   float *out_xyz, float *out_nrm, int *n_out);

/* ---- Operator-level entry points (the MATLAB prototype's function names) ---------- */
/* point_pair_feature + my_discretize for n independent pairs (matlab/point_pair_feature.m:1-11,
 * my_discretize.m:3-4; compute_ppf + disc_feature, kernel.cu:94-122).  p1,n1,p2,n2: n x 3 host floats.
 * raw_out (n x 4, may be NULL) = F = (|d|, ang(n1,d), ang(n2,d), ang(n1,n2)); disc_out (n x 4, may be NULL) =
 * F - mod(F, step) with step d_dist for F1 and 2*pi/30 for F2..F4; keys_out (n, may be NULL) = the 32-bit
 * hash the table is keyed on. */
int ppf_point_pair_feature(const float *p1, const float *n1, const float *p2, const float *n2, size_t n,
                           float d_dist, float *raw_out, float *disc_out, uint32_t *keys_out);
/* trans_model_scene (matlab/trans_model_scene.m:1-41, kernel.cu:302-349) for n independent tuples (all
 * n x 3 host floats): T_m_g and T_s_g (n x 16 row-major each, may be NULL), alpha in radians (n, may be NULL)
 * and the accumulator bin alpha_idx in [0,30] (n, may be NULL). */
int ppf_trans_model_scene(const float *m_r, const float *n_r_m, const float *m_i, const float *s_r,
                          const float *n_r_s, const float *s_i, size_t n, float *T_m_g, float *T_s_g,
                          float *alpha, uint32_t *alpha_idx);

/* ---- Lookup (voting -> poses -> clustering) ------------------------------------- */
typedef struct ppf_lookup_stats {
    uint64_t num_scene_pairs;       /* R * N_s pairs processed by this call (the metric's unit)   */
    uint64_t num_nonunique_votes;   /* model.cu:121-122 */
    uint64_t num_unique_votes;      /* model.cu:152     */
    uint32_t max_vote_count;        /* temp_votecounts[0], model.cu:160-164 */
    uint32_t num_top_votes;         /* K, model.cu:165-170 */
    uint32_t max_idx;               /* model.cu:293-295 */
    uint32_t num_exact_alpha;       /* votes that took the exact-alpha path (diagnostic)          */
    float    ms_vote, ms_finalize, ms_pose_cluster;   /* CUDA-event times of the last call        */
} ppf_lookup_stats_t;

int ppf_lookup_create(ppf_lookup_t **out);
void ppf_lookup_destroy(ppf_lookup_t *lk);

/* Whole Model::ppf_lookup: vote over every reference point s_r with
 * s_r % ref_point_downsample_factor == 0, threshold, poses, clustering, argmax. */
int ppf_model_lookup(const ppf_model_t *model, const ppf_scene_t *scene,
                     unsigned ref_point_downsample_factor, ppf_lookup_t *lk);

/* Staged form (what ppf_model_lookup runs, split so that several GPUs can share a
 * scene).  shard_rank / shard_count select every shard_count-th reference point. */
int ppf_lookup_vote(const ppf_model_t *model, const ppf_scene_t *scene,
                    unsigned ref_point_downsample_factor, int shard_rank, int shard_count,
                    ppf_lookup_t *lk);
int ppf_lookup_local_max(const ppf_lookup_t *lk, uint32_t *max_count);
/* Keep votes with count > vote_count_threshold * global_max, ordered (count desc, code asc). */
int ppf_lookup_finalize(const ppf_model_t *model, uint32_t global_max, ppf_lookup_t *lk);
/* Survivor list as device pointers (for NCCL allgather) and its replacement by a
 * merged list (device pointers, K entries, already filtered; will be re-ordered). */
int ppf_lookup_survivors(const ppf_lookup_t *lk, size_t *K, const uint64_t **codes_dev,
                         const uint32_t **counts_dev);
int ppf_lookup_set_survivors(ppf_lookup_t *lk, const uint64_t *codes_dev, const uint32_t *counts_dev,
                             size_t K);
/* Device-to-device copy of the K survivors into caller-owned device buffers. */
int ppf_lookup_copy_survivors(const ppf_lookup_t *lk, uint64_t *codes_dst_dev, uint32_t *counts_dst_dev);
int ppf_lookup_poses(const ppf_model_t *model, const ppf_scene_t *scene, ppf_lookup_t *lk);
int ppf_lookup_cluster(const ppf_model_t *model, ppf_lookup_t *lk);
/* Multi-GPU clustering of a merged survivor list (Model::ClusterTransformations is quadratic in dense cells):
 * rank `shard` of `n_shards` scores the poses shard, shard + n_shards, ...; the caller sums the score arrays of
 * all ranks (entries outside a slice are 0), writes the sum back and lets ppf_lookup_cluster_finish pick
 * max_idx (model.cu:293-295).  Not available with use_averaged_clusters. */
int ppf_lookup_cluster_shard(const ppf_model_t *model, ppf_lookup_t *lk, int shard, int n_shards);
int ppf_lookup_copy_scores(const ppf_lookup_t *lk, float *scores_dst_dev);
int ppf_lookup_set_scores(ppf_lookup_t *lk, const float *scores_src_dev);
int ppf_lookup_cluster_finish(ppf_lookup_t *lk);
/* cpu_clustering = true variant: PCL-style greedy clustering of the survivors on one host
 * core (transformation_clustering.cpp:62-137, model.cu:246-266); writes the best cluster's
 * averaged pose (row-major 4x4) = cpu_transformations[0] of ppf.cu:75-77. */
int ppf_lookup_cluster_cpu(const ppf_model_t *model, ppf_lookup_t *lk, float *pose_out);

int ppf_lookup_get_stats(const ppf_lookup_t *lk, ppf_lookup_stats_t *stats);
/* Host copies; K = stats.num_top_votes. votes/counts/weighted/scores: K; transformations: 16K
 * (row-major 4x4); trans: 3K; rots: 4K (scalar part first); pose: 16 = rotation rows of
 * transformations[max_idx] with translation transformation_trans[max_idx] (ppf.cu:80-93). */
int ppf_lookup_get(const ppf_lookup_t *lk, uint64_t *votes, uint32_t *counts, float *transformations,
                   float *weighted, float *trans, float *rots, float *scores, float *pose);
/* Every non-zero accumulator cell (code, count), ascending code: the reference's
 * unique votes before thresholding (model.cu:148-152). Needs vote_count_threshold = 0 semantics,
 * so it re-runs the vote stage; meant for parity tests on small inputs. */
int ppf_vote_histogram(const ppf_model_t *model, const ppf_scene_t *scene,
                       unsigned ref_point_downsample_factor, uint64_t *codes_out,
                       uint32_t *counts_out, size_t capacity, size_t *n_out);

/* The same for one shard of the reference points only (shard_rank / shard_count as in ppf_lookup_vote: reference
 * point number shard_rank + k * shard_count of the points with s_r % ref_point_downsample_factor == 0): full-size
 * parity checks against a sample of reference points (oracle Mode B). */
int ppf_vote_histogram_shard(const ppf_model_t *model, const ppf_scene_t *scene,
                             unsigned ref_point_downsample_factor, int shard_rank, int shard_count,
                             uint64_t *codes_out, uint32_t *counts_out, size_t capacity, size_t *n_out);

typedef struct ppf_cloud {
    const float *xyz; int xyz_stride;
    const float *nrm; int nrm_stride;
    int n;
} ppf_cloud_t;

/* ---- Multi-GPU: scene reference points sharded over the ranks of one node --------------
 * One host process (or host thread) per GPU, the rank's device current.  Voting needs no data-path collective (a
 * scene reference point owns its accumulator: the high 32 bits of a vote code are s_r, model.h:61-63); what the
 * reference does GLOBALLY -- the threshold count > thr * max and the clustering of the surviving votes,
 * model.cu:160-170, 202-244 -- is done on the merged survivor list: all_reduce(MAX) of one u32, all_gather of
 * the survivor records, all_reduce(SUM) of the clustering scores of interleaved pose slices.  The collectives are
 * NCCL's (libnccl.so.2 is dlopen()ed on first use: no link-time dependency); the model table is replicated.
 *
 * A communicator is made from an NCCL unique id (rank 0 calls ppf_comm_unique_id and ships the PPF_COMM_ID_BYTES
 * bytes to the other ranks by any means, then every rank calls ppf_comm_create_nccl), or wraps an ncclComm_t the
 * host already owns.  ppf_comm_create_local makes `world` communicators for `world` host THREADS of one process on
 * one GPU (rendezvous through host memory): the single-GPU test vehicle of this path. */
#define PPF_COMM_ID_BYTES 128
int ppf_comm_unique_id(void *id_out);
int ppf_comm_create_nccl(const void *id, int rank, int world, ppf_comm_t **out);
int ppf_comm_wrap_nccl(void *nccl_comm, int rank, int world, ppf_comm_t **out);
int ppf_comm_create_local(int world, ppf_comm_t **out_array);
int ppf_comm_rank(const ppf_comm_t *comm);
int ppf_comm_size(const ppf_comm_t *comm);
void ppf_comm_destroy(ppf_comm_t *comm);
/* Model::ppf_lookup over the ranks of `comm`: every rank passes the same model (replicated) and the same scene and
 * gets the same survivors, poses, scores and winner as ppf_model_lookup on one GPU (bit for bit; with
 * use_averaged_clusters the clustering runs replicated).  The stats' vote / pair counters are the rank's own. */
int ppf_model_lookup_sharded(const ppf_model_t *model, const ppf_scene_t *scene,
                             unsigned ref_point_downsample_factor, ppf_comm_t *comm, ppf_lookup_t *lk);
/* ppf_registration called by every rank of `comm` with the same arguments: same poses on every rank. */
int ppf_registration_sharded(const ppf_cloud_t *scene_clouds, int num_scenes, const ppf_cloud_t *model_clouds,
                             int num_models, const float *model_d_dists,
                             unsigned ref_point_downsample_factor, float vote_count_threshold,
                             int cpu_clustering, int use_l1_norm, int use_averaged_clusters,
                             ppf_comm_t *comm, float *poses_out, int *status_out);

/* ---- The drop-in boundary -------------------------------------------------------- */

/* ppf_registration (ppf.h:9-15): for every scene i and model j build Scene(scene_i, d_dist_j,
 * df) and Model(model_j, d_dist_j, ...), run ppf_lookup and write the best model->scene pose to
 * poses_out[(i*num_models + j)*16 ..] (row-major 4x4).  Host clouds in, host poses out.
 * cpu_clustering selects the PCL-style greedy clustering (transformation_clustering.cpp:62-137; Eigen's conversions
 * are restated, not linked: PARITY UNPINNED for that option).
 * device follows the reference: the device used is min(device_count-1, device) (ppf.cu:45);
 * model_weights is accepted and ignored, as in the reference (ppf.cu:35). status_out (optional,
 * num_scenes*num_models) receives the per-pair status (PPF_ERR_NO_VOTES leaves a zero pose). */
int ppf_registration(const ppf_cloud_t *scene_clouds, int num_scenes, const ppf_cloud_t *model_clouds,
                     int num_models, const float *model_d_dists,
                     unsigned ref_point_downsample_factor, float vote_count_threshold,
                     int cpu_clustering, int use_l1_norm, int use_averaged_clusters, int device,
                     const float *model_weights, float *poses_out, int *status_out);

#ifdef __cplusplus
}
#endif
#endif /* PPF_B200_H */
