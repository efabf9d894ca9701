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

// ---- Mode B: tiled oracle for sizes the whole-kernel replay cannot hold ------------------------
// The reference materialises N_s^2 features, keys and one 8-byte code per vote (~60 N_s^2 bytes + 8 B per vote):
// a 50k-point scene does not fit, a 1M-point scene overflows its int indices.  Scene reference points are
// independent (the high 32 bits of a vote code are s_r, model.h:61-63), so Mode B replays the SAME per-pair and
// per-vote steps for a LIST of reference points, streaming the scene, into a dense (m_r, alpha) histogram per
// reference point.  All arithmetic is the reference's own device code, linked through -rdc: compute_ppf
// (kernel.cu:109-122), disc_feature (:94-100), hash (:23-30), trans_model_scene (:302-349); the model table is
// the reference's ParallelHashArray built by ref_model_create (Mode A).  Ours: the loops, lower_bound (the
// std:: semantics thrust::lower_bound implements) and the hit test of ppf_vote_count_kernel (:489-497).
// tests/test_modeb_gpu.py first shows Mode B == Mode A cell for cell on sizes both run.
}  // extern "C"

extern __device__ void trans_model_scene(float3 m_r, float3 n_r_m, float3 m_i, float3 s_r, float3 n_r_s, float3 s_i,
                                         float d_dist, unsigned int &alpha_idx);

__global__ void modeb_vote_kernel(const float3 *spts, const float3 *snrm, int ns, const int *refs, int R,
                                  const float3 *mpts, const float3 *mnrm, int nm,
                                  const unsigned int *hashKeys, const std::size_t *ppfCount,
                                  const std::size_t *firstPPFIndex, const std::size_t *key2ppfMap, std::size_t U,
                                  float d_dist, unsigned int *hist, unsigned long long *num_votes) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = (long long)R * ns;
    unsigned long long mine = 0;
    for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
        const int r = (int)(t / ns), j = (int)(t - (long long)r * ns);
        const int s_r = refs[r];
        if (j == s_r) continue;                                          // ppf_kernel: .x = NaN -> key 0 (kernel.cu:437-441)
        // ppf_kernel (kernel.cu:444-450) + ppf_hash_kernel (kernel.cu:466-471)
        float4 f = disc_feature(compute_ppf(spts[s_r], snrm[s_r], spts[j], snrm[j]), d_dist, D_ANGLE0);
        const unsigned int key = isnan(f.x) ? 0u : hash(&f, sizeof(float4));
        // GetIndices = lower_bound (parallel_hash_array.hpp:81-92); hit test of ppf_vote_count_kernel (kernel.cu:489-497)
        std::size_t lo = 0, hi = U;
        while (lo < hi) {
            const std::size_t mid = lo + (hi - lo) / 2;
            if (hashKeys[mid] < key) lo = mid + 1; else hi = mid;
        }
        if (key == 0 || lo >= U || key != hashKeys[lo]) continue;
        const std::size_t cnt = ppfCount[lo], first = firstPPFIndex[lo];
        // ppf_vote_kernel's loop over the bucket (kernel.cu:536-550), 32 entries at a time
        for (std::size_t i = lane; i < cnt; i += 32) {
            const unsigned int modelPPFIndex = (unsigned int)key2ppfMap[first + i];
            const unsigned int model_r_index = modelPPFIndex / nm;
            const unsigned int model_i_index = modelPPFIndex - model_r_index * nm;
            unsigned int alpha_idx;
            trans_model_scene(mpts[model_r_index], mnrm[model_r_index], mpts[model_i_index],
                              spts[s_r], snrm[s_r], spts[j], d_dist, alpha_idx);
            atomicAdd(&hist[((std::size_t)r * nm + model_r_index) * 64 + (alpha_idx & 63u)], 1u);
            mine++;
        }
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(num_votes, mine);
}

extern "C" {

// Votes of the scene reference points refs[0..R) (caller's indices into the scene cloud, each must be a point the
// lookup would use, i.e. s_r % ref_point_downsample_factor == 0) against model hm.  codes_out / counts_out receive
// every non-zero accumulator cell as the reference's unique-vote code [s_r : 32 | m_r : 26 | alpha : 6] and its
// count, ascending code order (what thrust::sort + histogram leave, model.cu:148-151).  Returns the number of
// cells (nothing is written when it exceeds capacity), -1 on a CUDA error; *num_votes_out = votes cast.
long ref_modeb_histogram(void *hm, const float *sxyz, const float *snrm, int ns, const int *refs, int R,
                         unsigned long *codes_out, unsigned int *counts_out, long capacity,
                         unsigned long *num_votes_out) {
    RefModel *m = (RefModel *)hm;
    if (num_votes_out) *num_votes_out = 0;
    if (R <= 0 || ns <= 1 || m->n <= 1) return 0;
    thrust::host_vector<float3> hp(ns), hn(ns);
    for (int i = 0; i < ns; i++) {
        hp[i] = make_float3(sxyz[3 * i], sxyz[3 * i + 1], sxyz[3 * i + 2]);
        hn[i] = make_float3(snrm[3 * i], snrm[3 * i + 1], snrm[3 * i + 2]);
    }
    thrust::device_vector<float3> dp = hp, dn = hn;
    thrust::device_vector<int> drefs(refs, refs + R);
    thrust::device_vector<unsigned int> hist((std::size_t)R * m->n * 64, 0u);
    thrust::device_vector<unsigned long long> nv(1, 0ull);
    modeb_vote_kernel<<<148 * 16, 256>>>(
        thrust::raw_pointer_cast(dp.data()), thrust::raw_pointer_cast(dn.data()), ns,
        thrust::raw_pointer_cast(drefs.data()), R, thrust::raw_pointer_cast(m->points.data()),
        thrust::raw_pointer_cast(m->normals.data()), m->n, RAW_PTR(m->search_array.GetHashkeys()),
        RAW_PTR(m->search_array.GetCounts()), RAW_PTR(m->search_array.GetFirstHashkeyIndices()),
        RAW_PTR(m->search_array.GetHashkeyToDataMap()), m->search_array.GetHashkeys()->size(), m->d_dist,
        thrust::raw_pointer_cast(hist.data()), thrust::raw_pointer_cast(nv.data()));
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (num_votes_out) *num_votes_out = (unsigned long)nv[0];
    thrust::host_vector<unsigned int> hh = hist;
    // reference points in ascending s_r so that the codes come out sorted
    std::vector<int> order(R);
    for (int i = 0; i < R; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return refs[a] < refs[b]; });
    long n = 0;
    for (int pass = 0; pass < 2; pass++) {
        long k = 0;
        for (int oi = 0; oi < R; oi++) {
            const int r = order[oi];
            for (int mr = 0; mr < m->n; mr++)
                for (unsigned int a = 0; a < 64; a++) {
                    const unsigned int c = hh[((std::size_t)r * m->n + mr) * 64 + a];
                    if (!c) continue;
                    if (pass == 1) {
                        codes_out[k] = (((unsigned long)(unsigned int)refs[r]) << 32) | ((unsigned long)mr << 6) | a;
                        counts_out[k] = c;
                    }
                    k++;
                }
        }
        n = k;
        if (pass == 0 && (!codes_out || !counts_out || n > capacity)) break;
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
