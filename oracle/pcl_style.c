/* oracle/pcl_style.c -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU baseline (2) of SURVEY.md 8(d): Drost et al.'s point-pair-feature registration the way PCL's
 * PPFEstimation / PPFHashMapSearch / PPFRegistration organise it -- the algorithm the reference's CUDA path was
 * written to replace (north star: "PCL PPFEstimation/PPFRegistration").  PCL itself is not under
 * /root/reference and is not installed (only the derived clusterPoses survives in the tree,
 * pcl/alignment/src/transformation_clustering.cpp:62-137), so this is a restatement of the PUBLISHED algorithm
 * (Drost, Ulrich, Navab, Ilic: "Model Globally, Match Locally", CVPR 2010, sections 3-4) anchored on the
 * reference wherever it keeps a piece of it:
 *
 *   feature F = (|d|, angle(n1,d), angle(n2,d), angle(n1,n2))      kernel.cu:109-122 / point_pair_feature.m
 *   integer bins floor(F / step), hash multimap bins -> (m_r, m_i, alpha_m)   [PPFHashMapSearch]
 *   alpha_m precomputed per model pair, alpha_s once per SCENE pair, vote = alpha_s - alpha_m   [Drost 4.2]
 *   one (model point, alpha) accumulator per scene reference point, its peak = one pose candidate
 *   greedy clustering of the candidates, best cluster averaged     transformation_clustering.cpp:62-137
 *
 * What makes it the strongest CPU comparator: one atan2 per scene pair instead of one full frame construction per
 * vote (the MATLAB path, oracle/drost_m.c) -- a vote is an add and a table increment.
 *
 * PARITY UNPINNED (no PCL here): a timing comparator checked only by the reference's own acceptance test
 * (planted pose recovered, alignment.cpp:317-323) in tests/test_pcl_style.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N_ANGLE 30

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static float clamp1(float x) { return x > 1.f ? 1.f : (x < -1.f ? -1.f : x); }

/* local frame of a point: rotation taking its normal onto +x (Drost 4.2, fig. 3), then T = R * trans(-p) */
typedef struct { float R[9]; float p[3]; } frame_t;
static void make_frame(const float *p, const float *n, frame_t *f) {
    float len = sqrtf(dot3(n, n));
    float nx = n[0] / len, ny = n[1] / len, nz = n[2] / len;
    /* Rodrigues: axis = n x e_x = (0, nz, -ny), angle = acos(nx) */
    float s = sqrtf(ny * ny + nz * nz), c = nx;
    float *R = f->R;
    if (s < 1e-12f) {                                   /* normal already on the x axis (or opposite) */
        memset(R, 0, sizeof(float) * 9);
        R[0] = c >= 0 ? 1.f : -1.f; R[4] = 1.f; R[8] = c >= 0 ? 1.f : -1.f;
    } else {
        float ax = 0.f, ay = nz / s, az = -ny / s, t = 1.f - c;
        R[0] = c + ax * ax * t;      R[1] = ax * ay * t - az * s; R[2] = ax * az * t + ay * s;
        R[3] = ay * ax * t + az * s; R[4] = c + ay * ay * t;      R[5] = ay * az * t - ax * s;
        R[6] = az * ax * t - ay * s; R[7] = az * ay * t + ax * s; R[8] = c + az * az * t;
    }
    memcpy(f->p, p, sizeof(float) * 3);
}
/* angle of (T q) around the x axis */
static float frame_alpha(const frame_t *f, const float *q) {
    float d[3] = {q[0] - f->p[0], q[1] - f->p[1], q[2] - f->p[2]};
    float y = f->R[3] * d[0] + f->R[4] * d[1] + f->R[5] * d[2];
    float z = f->R[6] * d[0] + f->R[7] * d[1] + f->R[8] * d[2];
    return atan2f(-z, y);
}

/* feature bins; returns 0 when the pair has no feature (coincident points, zero normals) */
static int feature_bins(const float *p1, const float *n1, const float *p2, const float *n2, float d_dist, float d_angle,
                        uint64_t *key) {
    float d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    float dn = sqrtf(dot3(d, d)), l1 = sqrtf(dot3(n1, n1)), l2 = sqrtf(dot3(n2, n2));
    if (!(dn > 0.f) || !(l1 > 0.f) || !(l2 > 0.f)) return 0;
    float f1 = acosf(clamp1(dot3(n1, d) / (l1 * dn))), f2 = acosf(clamp1(dot3(n2, d) / (l2 * dn)));
    float f3 = acosf(clamp1(dot3(n1, n2) / (l1 * l2)));
    uint64_t k0 = (uint64_t)floorf(dn / d_dist), k1 = (uint64_t)floorf(f1 / d_angle), k2 = (uint64_t)floorf(f2 / d_angle),
             k3 = (uint64_t)floorf(f3 / d_angle);
    if (k0 >= (1u << 20)) return 0;
    *key = (k0 << 24) | (k1 << 16) | (k2 << 8) | k3;
    return 1;
}

/* ---- PPFHashMapSearch: open-addressing table bins -> slice of the (m_r, m_i, alpha_m) array ---- */
typedef struct { uint64_t key; uint32_t first, count, fill; int used; } cell_t;
typedef struct { uint32_t pair; float alpha_m; } entry_t;       /* pair = m_r << 16 | m_i (N < 65536) */
typedef struct { cell_t *cells; size_t cap, ncells; entry_t *entries; } table_t;

static cell_t *table_find(table_t *t, uint64_t key, int insert);
static void table_grow(table_t *t) {
    table_t big = *t;
    big.cap = t->cap * 2; big.ncells = 0;
    big.cells = calloc(big.cap, sizeof(cell_t));
    for (size_t i = 0; i < t->cap; i++)
        if (t->cells[i].used) { cell_t *c = table_find(&big, t->cells[i].key, 1); c->count = t->cells[i].count; }
    free(t->cells);
    *t = big;
}
static cell_t *table_find(table_t *t, uint64_t key, int insert) {
    for (;;) {
        size_t i = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (t->cap - 1);
        for (;;) {
            cell_t *c = &t->cells[i];
            if (c->used) {
                if (c->key == key) return c;
                i = (i + 1) & (t->cap - 1);
                continue;
            }
            if (!insert) return NULL;
            if (2 * (t->ncells + 1) > t->cap) break;
            c->used = 1; c->key = key; c->first = c->count = c->fill = 0;
            t->ncells++;
            return c;
        }
        table_grow(t);
    }
}

static void model_table(table_t *t, const float *mp, const float *mn, int n, float d_dist, float d_angle) {
    enum { ROWS = 64 };
    t->cap = 1u << 14; t->ncells = 0;
    t->cells = calloc(t->cap, sizeof(cell_t));
    t->entries = malloc(sizeof(entry_t) * ((size_t)n * n + 1));
    uint64_t *keys = malloc(sizeof(uint64_t) * (size_t)ROWS * (n ? n : 1));
    float *alphas = malloc(sizeof(float) * (size_t)ROWS * (n ? n : 1));
    for (int pass = 0; pass < 2; pass++) {
        for (int a0 = 0; a0 < n; a0 += ROWS) {
            const int a1 = a0 + ROWS < n ? a0 + ROWS : n;
#pragma omp parallel for schedule(static)
            for (int a = a0; a < a1; a++) {
                frame_t F;
                make_frame(mp + 3 * a, mn + 3 * a, &F);
                for (int b = 0; b < n; b++) {
                    size_t at = (size_t)(a - a0) * n + b;
                    uint64_t key;
                    if (a == b || !feature_bins(mp + 3 * a, mn + 3 * a, mp + 3 * b, mn + 3 * b, d_dist, d_angle, &key)) {
                        keys[at] = ~0ull;
                        continue;
                    }
                    keys[at] = key;
                    if (pass == 1) alphas[at] = frame_alpha(&F, mp + 3 * b);
                }
            }
            for (int a = a0; a < a1; a++)
                for (int b = 0; b < n; b++) {
                    size_t at = (size_t)(a - a0) * n + b;
                    if (keys[at] == ~0ull) continue;
                    cell_t *c = table_find(t, keys[at], 1);
                    if (pass == 0) c->count++;
                    else {
                        entry_t *e = &t->entries[c->first + c->fill++];
                        e->pair = ((uint32_t)a << 16) | (uint32_t)b;
                        e->alpha_m = alphas[at];
                    }
                }
        }
        if (pass == 0) {
            uint32_t run = 0;
            for (size_t i = 0; i < t->cap; i++) if (t->cells[i].used) { t->cells[i].first = run; run += t->cells[i].count; }
        }
    }
    free(keys); free(alphas);
}

typedef struct { float T[16]; unsigned votes; } pose_t;

static void rotx(float a, float R[9]) { memset(R, 0, 36); R[0] = 1; R[4] = cosf(a); R[5] = -sinf(a); R[7] = sinf(a); R[8] = cosf(a); }
/* T = T_sg^-1 * Rx(alpha) * T_mg (Drost eq. 2), row-major 4x4 */
static void compose_pose(const frame_t *Fs, float alpha, const frame_t *Fm, float T[16]) {
    float Rx[9], A[9], B[9];
    rotx(alpha, Rx);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {           /* A = Rx * Rm */
        float s = 0; for (int k = 0; k < 3; k++) s += Rx[3 * r + k] * Fm->R[3 * k + c];
        A[3 * r + c] = s;
    }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {           /* B = Rs' * A */
        float s = 0; for (int k = 0; k < 3; k++) s += Fs->R[3 * k + r] * A[3 * k + c];
        B[3 * r + c] = s;
    }
    memset(T, 0, 64); T[15] = 1;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[4 * r + c] = B[3 * r + c];
        /* x_s = Rs'(Rx Rm (x_m - p_m)) + p_s */
        T[4 * r + 3] = Fs->p[r] - (B[3 * r] * Fm->p[0] + B[3 * r + 1] * Fm->p[1] + B[3 * r + 2] * Fm->p[2]);
    }
}
static void quat_of(const float T[16], float q[4]) {                    /* (w, x, y, z), w >= 0 */
    float tr = T[0] + T[5] + T[10];
    if (tr > 0) { float s = sqrtf(tr + 1.f) * 2; q[0] = s / 4; q[1] = (T[9] - T[6]) / s; q[2] = (T[2] - T[8]) / s; q[3] = (T[4] - T[1]) / s; }
    else if (T[0] > T[5] && T[0] > T[10]) { float s = sqrtf(1.f + T[0] - T[5] - T[10]) * 2; q[0] = (T[9] - T[6]) / s; q[1] = s / 4; q[2] = (T[1] + T[4]) / s; q[3] = (T[2] + T[8]) / s; }
    else if (T[5] > T[10]) { float s = sqrtf(1.f + T[5] - T[0] - T[10]) * 2; q[0] = (T[2] - T[8]) / s; q[1] = (T[1] + T[4]) / s; q[2] = s / 4; q[3] = (T[6] + T[9]) / s; }
    else { float s = sqrtf(1.f + T[10] - T[0] - T[5]) * 2; q[0] = (T[4] - T[1]) / s; q[1] = (T[2] + T[8]) / s; q[2] = (T[6] + T[9]) / s; q[3] = s / 4; }
    if (q[0] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
}
static int cmp_votes_desc(const void *a, const void *b) {
    unsigned va = ((const pose_t *)a)->votes, vb = ((const pose_t *)b)->votes;
    return va < vb ? 1 : (va > vb ? -1 : 0);
}
/* transformation_clustering.cpp:62-137: sort by votes, join the first cluster whose SEED is within the bounds,
 * sum the votes, average translation and quaternion coefficients of the winning cluster */
static void cluster_poses(pose_t *poses, int n, float trans_thresh, float rot_thresh, double out[16]) {
    memset(out, 0, 128);
    if (n == 0) return;
    qsort(poses, n, sizeof(pose_t), cmp_votes_desc);
    int *seed = malloc(sizeof(int) * n), *member = malloc(sizeof(int) * n), ncl = 0;
    unsigned long long *cv = calloc(n, sizeof(unsigned long long));
    for (int i = 0; i < n; i++) {
        int found = -1;
        for (int c = 0; c < ncl && found < 0; c++) {
            const float *A = poses[i].T, *B = poses[seed[c]].T;
            float dt[3] = {A[3] - B[3], A[7] - B[7], A[11] - B[11]};
            float tr = 0;                                               /* trace(Ra' Rb) */
            for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) tr += A[4 * r + k] * B[4 * r + k];
            float ang = fabsf(acosf(clamp1((tr - 1.f) / 2.f)));
            if (sqrtf(dot3(dt, dt)) < trans_thresh && ang < rot_thresh) found = c;
        }
        if (found < 0) { seed[ncl] = i; found = ncl++; }
        member[i] = found; cv[found] += poses[i].votes;
    }
    int best = 0;
    for (int c = 1; c < ncl; c++) if (cv[c] > cv[best]) best = c;
    double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 0}; int cnt = 0;
    for (int i = 0; i < n; i++) if (member[i] == best) {
        float qi[4]; quat_of(poses[i].T, qi);
        for (int k = 0; k < 4; k++) q[k] += qi[k];
        t[0] += poses[i].T[3]; t[1] += poses[i].T[7]; t[2] += poses[i].T[11]; cnt++;
    }
    double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
    out[0] = 1 - 2 * (y * y + z * z); out[1] = 2 * (x * y - z * w);     out[2] = 2 * (x * z + y * w);
    out[4] = 2 * (x * y + z * w);     out[5] = 1 - 2 * (x * x + z * z); out[6] = 2 * (y * z - x * w);
    out[8] = 2 * (x * z - y * w);     out[9] = 2 * (y * z + x * w);     out[10] = 1 - 2 * (x * x + y * y);
    out[3] = t[0] / cnt; out[7] = t[1] / cnt; out[11] = t[2] / cnt; out[15] = 1;
    free(seed); free(member); free(cv);
}

/* Registration of one scene against one model.  Every `ref_rate`-th scene point is a reference point
 * (scene_reference_point_sampling_rate); max_refs / scene_stride bound the sample for timing, as in
 * oracle_time_voting.  Returns the seconds spent voting (+ clustering); *build_seconds = the model table. */
double pcl_style_run(const float *mxyz, const float *mnrm, int nm, const float *sxyz, const float *snrm, int ns,
                     float d_dist, int ref_rate, int max_refs, int scene_stride, int threads, uint64_t *pairs_out,
                     uint64_t *votes_out, double *build_seconds, double *pose_out /* 16, row-major, may be NULL */) {
    if (ref_rate < 1) ref_rate = 1;
    if (scene_stride < 1) scene_stride = 1;
    if (nm >= 65536) return -1.0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    const float d_angle = (float)(2 * M_PI / N_ANGLE);
    double t0 = now_s();
    table_t tab;
    model_table(&tab, mxyz, mnrm, nm, d_dist, d_angle);
    double t1 = now_s();
    if (build_seconds) *build_seconds = t1 - t0;

    const int R_all = ns > 0 ? (ns + ref_rate - 1) / ref_rate : 0;
    const int R = (max_refs > 0 && R_all > max_refs) ? max_refs : R_all;
    pose_t *poses = calloc(R ? R : 1, sizeof(pose_t));
    uint64_t votes = 0, pairs = 0;
#pragma omp parallel reduction(+ : votes, pairs)
    {
        uint32_t *acc = calloc((size_t)(nm ? nm : 1) * N_ANGLE, sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < R; k++) {
            const int r = (int)(((long long)k * R_all) / R) * ref_rate;
            frame_t Fs;
            make_frame(sxyz + 3 * r, snrm + 3 * r, &Fs);
            for (int i = k % scene_stride; i < ns; i += scene_stride) {
                uint64_t key;
                pairs++;
                if (i == r || !feature_bins(sxyz + 3 * r, snrm + 3 * r, sxyz + 3 * i, snrm + 3 * i, d_dist, d_angle, &key)) continue;
                const cell_t *c = table_find(&tab, key, 0);
                if (!c) continue;
                const float alpha_s = frame_alpha(&Fs, sxyz + 3 * i);          /* once per scene pair */
                const entry_t *e = tab.entries + c->first;
                for (uint32_t j = 0; j < c->count; j++) {
                    float alpha = e[j].alpha_m - alpha_s;                       /* Drost 4.2: alpha = alpha_m - alpha_s */
                    if (alpha < -(float)M_PI) alpha += 2 * (float)M_PI;
                    if (alpha >= (float)M_PI) alpha -= 2 * (float)M_PI;
                    int bin = (int)floorf((alpha + (float)M_PI) / d_angle);
                    if (bin >= N_ANGLE) bin = N_ANGLE - 1;
                    if (bin < 0) bin = 0;
                    acc[(size_t)(e[j].pair >> 16) * N_ANGLE + bin]++;
                }
                votes += c->count;
            }
            uint32_t best = 0; int br = 0, bc = 0;
            for (int a = 0; a < nm; a++) for (int b = 0; b < N_ANGLE; b++)
                if (acc[(size_t)a * N_ANGLE + b] > best) { best = acc[(size_t)a * N_ANGLE + b]; br = a; bc = b; }
            if (best > 0) {
                frame_t Fm;
                make_frame(mxyz + 3 * br, mnrm + 3 * br, &Fm);
                /* bin centre; Rx(alpha) takes the model pair, seen from its frame, onto the scene pair */
                compose_pose(&Fs, ((float)bc + 0.5f) * d_angle - (float)M_PI, &Fm, poses[k].T);
                poses[k].votes = best;
            }
            memset(acc, 0, (size_t)nm * N_ANGLE * sizeof(uint32_t));
        }
        free(acc);
    }
    int np = 0;
    for (int k = 0; k < R; k++) if (poses[k].votes) poses[np++] = poses[k];
    double pose[16];
    cluster_poses(poses, np, d_dist, d_angle, pose);
    double t2 = now_s();
    if (pose_out) memcpy(pose_out, pose, sizeof(pose));
    if (pairs_out) *pairs_out = pairs;
    if (votes_out) *votes_out = votes;
    free(poses); free(tab.cells); free(tab.entries);
    return t2 - t1;
}
