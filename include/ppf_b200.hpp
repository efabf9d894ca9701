// ppf_b200.hpp -- header-only C++ mirror of the reference's Scene / Model objects over the C ABI
// (include/ppf_b200.h).  Same constructor arguments and member names as
// pcl/alignment/include/scene.h:11-52 and include/model.h:14-115, minus the PCL / Thrust types:
// clouds are (xyz, normals) float arrays with a stride, results are std::vector copies.
// Errors throw std::runtime_error(ppf_last_error()) instead of exit() (util.hpp:18-26).
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "ppf_b200.h"

namespace ppf_b200 {

inline void check(int rc, bool allow_no_votes = false) {
    if (rc != PPF_OK && !(allow_no_votes && rc == PPF_ERR_NO_VOTES)) throw std::runtime_error(ppf_last_error());
}

struct CloudView {                 // e.g. {&pts[0].x, 12, &pts[0].normal_x, 12, n} for pcl::PointNormal
    const float *xyz; int xyz_stride;
    const float *nrm; int nrm_stride;
    int n;
    int mem = PPF_MEM_HOST;
};

class Scene {                      // scene.h:11-52
  public:
    Scene(const CloudView &cloud, float d_dist, unsigned int ref_point_downsample_factor = 1)
        : d_dist(d_dist), ref_point_downsample_factor(ref_point_downsample_factor) {
        check(ppf_scene_create(cloud.xyz, cloud.xyz_stride, cloud.nrm, cloud.nrm_stride, cloud.n, cloud.mem, &h_));
    }
    ~Scene() { ppf_scene_destroy(h_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    int numPoints() const { return ppf_scene_num_points(h_); }
    // getModelPPFs() / getHashKeys(): quantised features and keys, N*N, row-major [ref][other]
    void getFeatures(std::vector<float> &ppfs, std::vector<uint32_t> &keys) const {
        size_t n = (size_t)numPoints();
        ppfs.resize(n * n * 4); keys.resize(n * n);
        check(ppf_scene_features(h_, d_dist, ref_point_downsample_factor, 0, (int)n, 0, (int)n, ppfs.data(), keys.data()));
    }
    ppf_scene_t *handle() const { return h_; }
    const float d_dist;
    const unsigned int ref_point_downsample_factor;
  private:
    ppf_scene_t *h_ = nullptr;
};

class Model {                      // model.h:14-115
  public:
    Model(const CloudView &cloud, float d_dist, float vote_count_threshold, bool cpu_clustering,
          bool use_l1_norm, bool use_averaged_clusters)
        : cpu_clustering(cpu_clustering) {
        check(ppf_model_create(cloud.xyz, cloud.xyz_stride, cloud.nrm, cloud.nrm_stride, cloud.n, cloud.mem, d_dist,
                               vote_count_threshold, use_l1_norm, use_averaged_clusters, &h_));
        check(ppf_lookup_create(&lk_));
    }
    // persistent model database (no reference equivalent: ppf.cu:63-70 rebuilds the table per pair)
    Model(const char *path, bool cpu_clustering = false) : cpu_clustering(cpu_clustering) {
        check(ppf_model_load(path, &h_));
        check(ppf_lookup_create(&lk_));
    }
    void save(const char *path) const { check(ppf_model_save(h_, path)); }
    ~Model() { ppf_lookup_destroy(lk_); ppf_model_destroy(h_); }
    Model(const Model &) = delete;
    Model &operator=(const Model &) = delete;

    // Model::ppf_lookup (model.cu:269-306); fills the public members below. Returns false when no
    // scene pair matched the model (the reference would read an empty vector).
    bool ppf_lookup(const Scene *scene) {
        int rc = ppf_model_lookup(h_, scene->handle(), scene->ref_point_downsample_factor, lk_);
        check(rc, true);
        ppf_lookup_stats_t st;
        check(ppf_lookup_get_stats(lk_, &st));
        stats = st;
        size_t K = st.num_top_votes;
        votes.resize(K); voteCounts.resize(K); transformations.resize(K * 16); weightedVoteCounts.resize(K);
        transformation_trans.resize(K * 3); transformation_rots.resize(K * 4); vote_counts_out.resize(K);
        check(ppf_lookup_get(lk_, votes.data(), voteCounts.data(), transformations.data(), weightedVoteCounts.data(),
                             transformation_trans.data(), transformation_rots.data(), vote_counts_out.data(), pose.data()));
        max_idx = st.max_idx;
        if (cpu_clustering && rc == PPF_OK) check(ppf_lookup_cluster_cpu(h_, lk_, pose.data()));
        return rc == PPF_OK;
    }
    // ParallelHashArray::Get* (parallel_hash_array.hpp:25-30)
    void getTable(std::vector<uint32_t> &hashkeys, std::vector<size_t> &counts, std::vector<size_t> &first,
                  std::vector<size_t> &map) const {
        size_t U = 0, N = 0;
        check(ppf_model_table_sizes(h_, &U, &N));
        hashkeys.resize(U); counts.resize(U); first.resize(U); map.resize(N);
        check(ppf_model_table_get(h_, hashkeys.data(), counts.data(), first.data(), map.data()));
    }
    const std::vector<float> &getTransformations() const { return transformations; }

    const bool cpu_clustering;
    std::vector<uint64_t> votes;                 // [scene ref:32 | model point:26 | alpha:6]  model.h:61-63
    std::vector<uint32_t> voteCounts;
    std::vector<float> transformations;          // K x 16, row-major
    std::vector<float> weightedVoteCounts, transformation_trans, transformation_rots, vote_counts_out;
    unsigned int max_idx = 0;
    std::array<float, 16> pose{};                // the pose ppf_registration returns (ppf.cu:80-93)
    ppf_lookup_stats_t stats{};
  private:
    ppf_model_t *h_ = nullptr;
    ppf_lookup_t *lk_ = nullptr;
};

}  // namespace ppf_b200
