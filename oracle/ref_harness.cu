// oracle/ref_harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the reference's own, unmodified CUDA kernels (compiled in place from
// /root/reference/pcl/alignment/src/cuda/{kernel,vector_ops}.cu for sm_100a with
// the reference's numeric flags, see oracle/Makefile) and its own
// ParallelHashArray / histogram templates (impl/parallel_hash_array.hpp,
// impl/util.hpp, included behind the stub headers in oracle/stubs/).
//
// The reference's Scene/Model classes cannot be compiled here (they need PCL,
// Eigen and Boost), so this file replays their call sequence -- and nothing
// else; it performs no arithmetic of its own:
//   Scene::Scene            scene.cu:24-55, initPPFs scene.cu:64-99
//   Model::Model            model.cu:43-82
//   ComputeUniqueVotes      model.cu:95-171
//   ComputeTransformations  model.cu:191-200
//   ComputeWeightedVoteCounts model.cu:173-189
//   ClusterTransformations  model.cu:202-244
//   ppf_lookup              model.cu:269-306
//   final pose extraction   ppf.cu:74-93
// Every launch uses the reference's launch rule: BLOCK_SIZE threads,
// min(ceil(count/BLOCK_SIZE), MAX_NBLOCKS) blocks (kernel.h:11-12).
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
// legs may load the library built from this file.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/binary_search.h>
#include <thrust/count.h>
#include <thrust/device_vector.h>
#include <thrust/extrema.h>
#include <thrust/host_vector.h>
#include <thrust/scan.h>
#include <thrust/sort.h>

#include "impl/parallel_hash_array.hpp"   // reference header (unmodified)
#include "impl/util.hpp"                  // reference header (unmodified)
#include "kernel.h"                       // reference header (unmodified)

namespace {

int launch_blocks(size_t count) {
    return std::min(((int)count + BLOCK_SIZE - 1) / BLOCK_SIZE, MAX_NBLOCKS);
}

struct GreaterThanFloat {
    float thr;
    __host__ __device__ bool operator()(unsigned int x) const { return x > thr; }
};

// Mirrors the data members of the reference's Scene (scene.h:28-49).
struct RefScene {
    int n = 0;
    float d_dist = 0.f;
    thrust::device_vector<float3> points, normals;
    thrust::device_vector<float4> ppfs;
    thrust::device_vector<unsigned int> hashKeys;   // N*N, row-major [ref][other]
};

// Mirrors the data members of the reference's Model (model.h:41-113).
struct RefModel : RefScene {
    ParallelHashArray<unsigned int> search_array;
    thrust::device_vector<float> modelPointVoteWeights;
    // results of ppf_lookup
    unsigned long num_nonunique_votes = 0, num_unique_votes = 0;
    thrust::device_vector<unsigned long> votes;
    thrust::device_vector<unsigned int> voteCounts;
    thrust::device_vector<float> transformations;
    thrust::device_vector<float> weightedVoteCounts;
    thrust::device_vector<float3> transformation_trans;
    thrust::device_vector<float4> transformation_rots;
    thrust::device_vector<float> vote_counts_out;
    unsigned int max_idx = 0;
};

// scene.cu:24-55 + 64-99
void init_scene(RefScene &s, const float *xyz, const float *nrm, int n, float d_dist,
                unsigned int ref_point_downsample_factor) {
    thrust::host_vector<float3> hp(n), hn(n);
    for (int i = 0; i < n; i++) {
        hp[i] = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        hn[i] = make_float3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    }
    s.n = n;
    s.d_dist = d_dist;
    s.points = hp;
    s.normals = hn;
    s.ppfs = thrust::device_vector<float4>((size_t)n * n);
    int blocks = std::min((n + BLOCK_SIZE - 1) / BLOCK_SIZE, MAX_NBLOCKS);
    ppf_kernel<<<blocks, BLOCK_SIZE>>>(thrust::raw_pointer_cast(s.points.data()),
                                       thrust::raw_pointer_cast(s.normals.data()),
                                       thrust::raw_pointer_cast(s.ppfs.data()), n,
                                       ref_point_downsample_factor, d_dist);
    HANDLE_ERROR(cudaGetLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());
}

void hash_scene(RefScene &s, thrust::device_vector<unsigned int> &keys) {
    keys = thrust::device_vector<unsigned int>(s.ppfs.size());
    ppf_hash_kernel<<<launch_blocks(s.ppfs.size()), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(s.ppfs.data()), thrust::raw_pointer_cast(keys.data()),
        s.ppfs.size());
    HANDLE_ERROR(cudaPeekAtLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());
}

}  // namespace

extern "C" {

// ---- Scene ---------------------------------------------------------------
void *ref_scene_create(const float *xyz, const float *nrm, int n, float d_dist,
                       unsigned int ref_point_downsample_factor) {
    RefScene *s = new RefScene();
    init_scene(*s, xyz, nrm, n, d_dist, ref_point_downsample_factor);
    hash_scene(*s, s->hashKeys);
    return s;
}
void ref_scene_destroy(void *h) { delete (RefScene *)h; }

// Quantised features (float4 per ordered pair) and hash keys, N*N each.
void ref_scene_get(void *h, float *ppfs_out, unsigned int *keys_out) {
    RefScene *s = (RefScene *)h;
    if (ppfs_out)
        cudaMemcpy(ppfs_out, thrust::raw_pointer_cast(s->ppfs.data()),
                   s->ppfs.size() * sizeof(float4), cudaMemcpyDeviceToHost);
    if (keys_out)
        cudaMemcpy(keys_out, thrust::raw_pointer_cast(s->hashKeys.data()),
                   s->hashKeys.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
}

// ---- Model ---------------------------------------------------------------
// model.cu:43-82
void *ref_model_create(const float *xyz, const float *nrm, int n, float d_dist) {
    RefModel *m = new RefModel();
    init_scene(*m, xyz, nrm, n, d_dist, 1);
    m->modelPointVoteWeights = thrust::device_vector<float>(n, 1.0);
    thrust::device_vector<unsigned int> nonunique_hashkeys;
    hash_scene(*m, nonunique_hashkeys);
    m->search_array = ParallelHashArray<unsigned int>(nonunique_hashkeys);
    HANDLE_ERROR(cudaPeekAtLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());
    return m;
}
void ref_model_destroy(void *h) { delete (RefModel *)h; }

// sizes[0] = number of unique keys U, sizes[1] = map length (N*N)
void ref_model_table_sizes(void *h, unsigned long *sizes) {
    RefModel *m = (RefModel *)h;
    sizes[0] = m->search_array.GetHashkeys()->size();
    sizes[1] = m->search_array.GetHashkeyToDataMap()->size();
}
void ref_model_table_get(void *h, unsigned int *hashkeys, unsigned long *counts,
                         unsigned long *first, unsigned long *map) {
    RefModel *m = (RefModel *)h;
    size_t U = m->search_array.GetHashkeys()->size();
    size_t N = m->search_array.GetHashkeyToDataMap()->size();
    if (hashkeys) cudaMemcpy(hashkeys, RAW_PTR(m->search_array.GetHashkeys()), U * 4, cudaMemcpyDeviceToHost);
    if (counts) cudaMemcpy(counts, RAW_PTR(m->search_array.GetCounts()), U * 8, cudaMemcpyDeviceToHost);
    if (first) cudaMemcpy(first, RAW_PTR(m->search_array.GetFirstHashkeyIndices()), U * 8, cudaMemcpyDeviceToHost);
    if (map) cudaMemcpy(map, RAW_PTR(m->search_array.GetHashkeyToDataMap()), N * 8, cudaMemcpyDeviceToHost);
}
void ref_model_get(void *h, float *ppfs_out) {
    RefModel *m = (RefModel *)h;
    cudaMemcpy(ppfs_out, thrust::raw_pointer_cast(m->ppfs.data()),
               m->ppfs.size() * sizeof(float4), cudaMemcpyDeviceToHost);
}

// ---- Model::ppf_lookup ----------------------------------------------------
// Returns the number of surviving votes K (votes whose count > thr * max), or
// -1 when no vote was cast at all (the reference would index an empty vector).
long ref_ppf_lookup(void *hm, void *hs, float vote_count_threshold, int use_l1_norm,
                    int use_averaged_clusters) {
    RefModel *m = (RefModel *)hm;
    RefScene *scene = (RefScene *)hs;

    // ---- ComputeUniqueVotes, model.cu:95-171
    thrust::device_vector<std::size_t> *sceneIndices = m->search_array.GetIndices(scene->hashKeys);
    size_t npairs = scene->hashKeys.size();
    thrust::device_vector<unsigned long> ppf_vote_counts(npairs);
    ppf_vote_count_kernel<<<launch_blocks(npairs), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(scene->hashKeys.data()), RAW_PTR(sceneIndices),
        RAW_PTR(m->search_array.GetHashkeys()), RAW_PTR(m->search_array.GetCounts()),
        thrust::raw_pointer_cast(ppf_vote_counts.data()), npairs);
    HANDLE_ERROR(cudaPeekAtLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());

    thrust::device_vector<std::size_t> ppf_vote_indices(npairs);
    thrust::exclusive_scan(ppf_vote_counts.begin(), ppf_vote_counts.end(), ppf_vote_indices.begin());
    std::size_t num_votes = 0;
    if (npairs) {
        unsigned long last_count = ppf_vote_counts.back();
        unsigned long last_index = ppf_vote_indices.back();
        num_votes = last_count + last_index;
    }
    m->num_nonunique_votes = num_votes;
    { thrust::device_vector<unsigned long> tmp; ppf_vote_counts.swap(tmp); }

    thrust::device_vector<unsigned long> nonunique_nonempty_votes(num_votes);
    ppf_vote_kernel<<<launch_blocks(npairs), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(scene->hashKeys.data()), RAW_PTR(sceneIndices),
        RAW_PTR(m->search_array.GetHashkeys()), RAW_PTR(m->search_array.GetCounts()),
        RAW_PTR(m->search_array.GetFirstHashkeyIndices()),
        RAW_PTR(m->search_array.GetHashkeyToDataMap()),
        thrust::raw_pointer_cast(m->points.data()), thrust::raw_pointer_cast(m->normals.data()), m->n,
        thrust::raw_pointer_cast(scene->points.data()), thrust::raw_pointer_cast(scene->normals.data()),
        scene->n, thrust::raw_pointer_cast(ppf_vote_indices.data()),
        thrust::raw_pointer_cast(nonunique_nonempty_votes.data()), npairs, m->d_dist);
    HANDLE_ERROR(cudaPeekAtLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());
    delete sceneIndices;
    { thrust::device_vector<std::size_t> tmp; ppf_vote_indices.swap(tmp); }

    if (num_votes == 0) {
        m->votes.clear(); m->voteCounts.clear(); m->num_unique_votes = 0;
        return -1;
    }

    thrust::sort(nonunique_nonempty_votes.begin(), nonunique_nonempty_votes.end());
    m->votes = thrust::device_vector<unsigned long>();
    m->voteCounts = thrust::device_vector<unsigned int>();
    histogram(nonunique_nonempty_votes, m->votes, m->voteCounts);
    m->num_unique_votes = m->votes.size();
    { thrust::device_vector<unsigned long> tmp; nonunique_nonempty_votes.swap(tmp); }

    thrust::sort_by_key(m->voteCounts.begin(), m->voteCounts.end(), m->votes.begin(),
                        thrust::greater<float>());

    unsigned int top = m->voteCounts[0];
    float min_votecount = vote_count_threshold * top;
    std::size_t num_top_votes =
        thrust::count_if(m->voteCounts.begin(), m->voteCounts.end(), GreaterThanFloat{min_votecount});
    m->votes.resize(num_top_votes);
    m->voteCounts.resize(num_top_votes);
    size_t K = num_top_votes;

    // ---- ComputeTransformations, model.cu:191-200
    m->transformations = thrust::device_vector<float>(K * 16);
    trans_calc_kernel2<<<launch_blocks(K), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(m->votes.data()), thrust::raw_pointer_cast(m->points.data()),
        thrust::raw_pointer_cast(m->normals.data()), thrust::raw_pointer_cast(scene->points.data()),
        thrust::raw_pointer_cast(scene->normals.data()),
        thrust::raw_pointer_cast(m->transformations.data()), K);

    // ---- ComputeWeightedVoteCounts, model.cu:173-189
    m->weightedVoteCounts = thrust::device_vector<float>(K);
    vote_weight_kernel<<<launch_blocks(K), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(m->votes.data()), thrust::raw_pointer_cast(m->voteCounts.data()),
        thrust::raw_pointer_cast(m->modelPointVoteWeights.data()),
        thrust::raw_pointer_cast(m->weightedVoteCounts.data()), K);

    // ---- ClusterTransformations, model.cu:202-244
    m->transformation_trans = thrust::device_vector<float3>(K);
    m->transformation_rots = thrust::device_vector<float4>(K);
    int blocks = launch_blocks(K);
    mat2transquat_kernel<<<blocks, BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(m->transformations.data()),
        thrust::raw_pointer_cast(m->transformation_trans.data()),
        thrust::raw_pointer_cast(m->transformation_rots.data()), K);
    thrust::device_vector<unsigned int> nonunique_trans_hash(K);
    thrust::device_vector<unsigned int> adjacent_trans_hash(27 * K);
    trans2idx_kernel<<<blocks, BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(m->transformation_trans.data()),
        thrust::raw_pointer_cast(nonunique_trans_hash.data()),
        thrust::raw_pointer_cast(adjacent_trans_hash.data()), K, m->d_dist);
    ParallelHashArray<unsigned int> trans_search_array =
        ParallelHashArray<unsigned int>(nonunique_trans_hash);
    thrust::device_vector<std::size_t> *transIndices = trans_search_array.GetIndices(adjacent_trans_hash);
    m->vote_counts_out = thrust::device_vector<float>(K);
    rot_clustering_kernel<<<blocks, BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(m->transformation_trans.data()),
        thrust::raw_pointer_cast(m->transformation_rots.data()),
        thrust::raw_pointer_cast(m->weightedVoteCounts.data()),
        thrust::raw_pointer_cast(adjacent_trans_hash.data()), RAW_PTR(transIndices),
        RAW_PTR(trans_search_array.GetHashkeys()), RAW_PTR(trans_search_array.GetCounts()),
        RAW_PTR(trans_search_array.GetFirstHashkeyIndices()),
        RAW_PTR(trans_search_array.GetHashkeyToDataMap()),
        thrust::raw_pointer_cast(m->vote_counts_out.data()), K, m->d_dist, use_l1_norm != 0,
        use_averaged_clusters != 0);
    HANDLE_ERROR(cudaPeekAtLastError());
    HANDLE_ERROR(cudaDeviceSynchronize());
    delete transIndices;

    // ---- ppf_lookup tail, model.cu:293-295
    m->max_idx = K ? (unsigned int)(thrust::max_element(m->vote_counts_out.begin(),
                                                        m->vote_counts_out.end()) -
                                    m->vote_counts_out.begin())
                   : 0;
    return (long)K;
}

// stats[0]=num_nonunique_votes stats[1]=num_unique_votes stats[2]=K stats[3]=max_idx
void ref_lookup_stats(void *hm, unsigned long *stats) {
    RefModel *m = (RefModel *)hm;
    stats[0] = m->num_nonunique_votes;
    stats[1] = m->num_unique_votes;
    stats[2] = m->votes.size();
    stats[3] = m->max_idx;
}

// Any pointer may be NULL. Array lengths: votes/counts/weighted/scores K,
// transformations 16K, trans 3K, rots 4K, pose 16 (ppf.cu:80-93).
void ref_lookup_get(void *hm, unsigned long *votes, unsigned int *counts, float *transformations,
                    float *weighted, float *trans, float *rots, float *scores, float *pose) {
    RefModel *m = (RefModel *)hm;
    size_t K = m->votes.size();
    if (votes) cudaMemcpy(votes, thrust::raw_pointer_cast(m->votes.data()), K * 8, cudaMemcpyDeviceToHost);
    if (counts) cudaMemcpy(counts, thrust::raw_pointer_cast(m->voteCounts.data()), K * 4, cudaMemcpyDeviceToHost);
    if (transformations) cudaMemcpy(transformations, thrust::raw_pointer_cast(m->transformations.data()), K * 64, cudaMemcpyDeviceToHost);
    if (weighted) cudaMemcpy(weighted, thrust::raw_pointer_cast(m->weightedVoteCounts.data()), K * 4, cudaMemcpyDeviceToHost);
    if (trans) cudaMemcpy(trans, thrust::raw_pointer_cast(m->transformation_trans.data()), K * 12, cudaMemcpyDeviceToHost);
    if (rots) cudaMemcpy(rots, thrust::raw_pointer_cast(m->transformation_rots.data()), K * 16, cudaMemcpyDeviceToHost);
    if (scores) cudaMemcpy(scores, thrust::raw_pointer_cast(m->vote_counts_out.data()), K * 4, cudaMemcpyDeviceToHost);
    if (pose && K) {
        // ppf.cu:80-93: rotation block of transformations[max_idx], translation
        // replaced by transformation_trans[max_idx].
        thrust::host_vector<float> T(m->transformations);
        thrust::host_vector<float3> tt(m->transformation_trans);
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) pose[r * 4 + c] = T[m->max_idx * 16 + r * 4 + c];
        pose[3] = tt[m->max_idx].x;
        pose[7] = tt[m->max_idx].y;
        pose[11] = tt[m->max_idx].z;
    }
}

// Raw (unsorted-by-count) vote histogram of the whole scene: all unique vote
// codes with their counts in ascending code order, before thresholding
// (model.cu:148-152). Runs the same kernels as ref_ppf_lookup up to histogram().
long ref_vote_histogram(void *hm, void *hs, unsigned long *codes_out, unsigned int *counts_out,
                        long capacity) {
    RefModel *m = (RefModel *)hm;
    RefScene *scene = (RefScene *)hs;
    thrust::device_vector<std::size_t> *sceneIndices = m->search_array.GetIndices(scene->hashKeys);
    size_t npairs = scene->hashKeys.size();
    thrust::device_vector<unsigned long> ppf_vote_counts(npairs);
    ppf_vote_count_kernel<<<launch_blocks(npairs), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(scene->hashKeys.data()), RAW_PTR(sceneIndices),
        RAW_PTR(m->search_array.GetHashkeys()), RAW_PTR(m->search_array.GetCounts()),
        thrust::raw_pointer_cast(ppf_vote_counts.data()), npairs);
    HANDLE_ERROR(cudaDeviceSynchronize());
    thrust::device_vector<std::size_t> ppf_vote_indices(npairs);
    thrust::exclusive_scan(ppf_vote_counts.begin(), ppf_vote_counts.end(), ppf_vote_indices.begin());
    std::size_t num_votes = 0;
    if (npairs) {
        unsigned long a = ppf_vote_counts.back(), b = ppf_vote_indices.back();
        num_votes = a + b;
    }
    thrust::device_vector<unsigned long> all_votes(num_votes);
    ppf_vote_kernel<<<launch_blocks(npairs), BLOCK_SIZE>>>(
        thrust::raw_pointer_cast(scene->hashKeys.data()), RAW_PTR(sceneIndices),
        RAW_PTR(m->search_array.GetHashkeys()), RAW_PTR(m->search_array.GetCounts()),
        RAW_PTR(m->search_array.GetFirstHashkeyIndices()),
        RAW_PTR(m->search_array.GetHashkeyToDataMap()),
        thrust::raw_pointer_cast(m->points.data()), thrust::raw_pointer_cast(m->normals.data()), m->n,
        thrust::raw_pointer_cast(scene->points.data()), thrust::raw_pointer_cast(scene->normals.data()),
        scene->n, thrust::raw_pointer_cast(ppf_vote_indices.data()),
        thrust::raw_pointer_cast(all_votes.data()), npairs, m->d_dist);
    HANDLE_ERROR(cudaDeviceSynchronize());
    delete sceneIndices;
    if (num_votes == 0) return 0;
    thrust::sort(all_votes.begin(), all_votes.end());
    thrust::device_vector<unsigned long> codes;
    thrust::device_vector<unsigned int> counts;
    histogram(all_votes, codes, counts);
    long n = (long)codes.size();
    if (codes_out && counts_out && n <= capacity) {
        cudaMemcpy(codes_out, thrust::raw_pointer_cast(codes.data()), n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(counts_out, thrust::raw_pointer_cast(counts.data()), n * 4, cudaMemcpyDeviceToHost);
    }
    return n;
}

// Wall-clock-free timing hook for the "reference GPU" comparator: time of
// Scene ctor + ppf_lookup (model prebuilt), in milliseconds, CUDA events.
float ref_time_scene_lookup(void *hm, const float *xyz, const float *nrm, int n,
                            unsigned int ref_df, float thr) {
    RefModel *m = (RefModel *)hm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    RefScene *s = (RefScene *)ref_scene_create(xyz, nrm, n, m->d_dist, ref_df);
    ref_ppf_lookup(m, s, thr, 0, 0);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ref_scene_destroy(s);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
}

}  // extern "C"
