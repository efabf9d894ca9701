/* ppf_oracle.h -- CPU restatement of the reference's PPF hot path. TEST INFRASTRUCTURE ONLY.
 * See ppf_oracle.c for the parity statement and the reference citations. */
#ifndef PPF_ORACLE_H
#define PPF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

float    oracle_d_angle0(void);
uint32_t oracle_hash(const void *bytes, int n);
void     oracle_compute_ppf(const float *p1, const float *n1, const float *p2, const float *n2, float *out4);
float    oracle_quant_downf(float x, float y);
void     oracle_disc_feature(const float *in4, float d_dist, float d_angle, float *out4);
void     oracle_rot(int axis, float theta, float *T16);
void     oracle_trans(const float *v3, float *T16);
void     oracle_mat4f_mul(const float *A16, const float *B16, float *C16);
void     oracle_mat4f_vmul(const float *A16, const float *b4, float *c4);
void     oracle_invht(const float *T16, float *Tinv16);
void     oracle_hrotmat2quat(const float *T16, float *q4);
uint32_t oracle_trans_model_scene(const float *m_r, const float *n_r_m, const float *m_i,
                                  const float *s_r, const float *n_r_s, const float *s_i);

/* Scene::Scene: quantised features (n*n*4 floats) and keys (n*n), row-major [ref][other]. */
void oracle_scene_features(const float *xyz, const float *nrm, int n, float d_dist, unsigned ref_df,
                           float *ppfs, uint32_t *keys);

/* ParallelHashArray: arrays sized npairs (only the first *U entries of hashkeys/counts/first are used). */
void oracle_hash_array(const uint32_t *keys, size_t npairs, uint32_t *hashkeys, uint64_t *counts,
                       uint64_t *first, uint64_t *map, size_t *U);

typedef struct oracle_result {
    uint64_t num_scene_pairs, num_nonunique_votes, num_unique_votes;
    uint32_t max_vote_count, K, max_idx;
    uint64_t *votes;            /* K, ordered (count desc, code asc) */
    uint32_t *counts;           /* K */
    float *transformations;     /* 16K */
    float *weighted;            /* K */
    float *trans;               /* 3K */
    float *rots;                /* 4K */
    float *scores;              /* K */
    float pose[16];
    /* full vote histogram (ascending code), filled when want_histogram != 0 */
    uint64_t hist_n;
    uint64_t *hist_codes;
    uint32_t *hist_counts;
} oracle_result_t;

/* Scene + Model + Model::ppf_lookup for one (scene, model) pair. scene_keys/model_keys may be NULL
 * (then they are computed with the CPU feature code) or hold keys produced elsewhere (e.g. by the
 * GPU), which makes every later integer stage exactly comparable. threads <= 0: all cores. */
int oracle_lookup(const float *mxyz, const float *mnrm, int nm, const float *sxyz, const float *snrm, int ns,
                  float d_dist, unsigned ref_df, float vote_count_threshold, int use_l1_norm,
                  int use_averaged_clusters, const uint32_t *model_keys, const uint32_t *scene_keys,
                  int want_histogram, int per_vote_frames, int threads, oracle_result_t *out);
void oracle_result_free(oracle_result_t *r);

/* Timing hook for bench.py: voting stage over a bounded sample of the workload -- max_refs reference
 * points spread over the scene, each paired with every scene_stride-th scene point (model table
 * built inside, timed separately); returns seconds, fills pairs/votes. */
double oracle_time_voting(const float *mxyz, const float *mnrm, int nm, const float *sxyz, const float *snrm,
                          int ns, float d_dist, unsigned ref_df, int max_refs, int scene_stride, int threads,
                          uint64_t *pairs, uint64_t *votes, double *build_seconds);

#ifdef __cplusplus
}
#endif
#endif
