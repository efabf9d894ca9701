/* oracle/drost_m.c -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's MATLAB / Octave pipeline (matlab/drost.m and the five operators it
 * calls), in double precision with MATLAB's loop structure, as the "reference's own CPU path" baseline that
 * SURVEY.md 8(d) item (1) asks to time beside the GPU path:
 *
 *   point_pair_feature   matlab/point_pair_feature.m:1-11     F = (|d|, acos(n1.d/..), acos(n2.d/..), acos(n1.n2/..))
 *   my_discretize        matlab/my_discretize.m:3-4           x - mod(x, step)
 *   model_description    matlab/model_description.m:1-70      containers.Map: quantised feature -> [(m_r, m_i); ...]
 *   trans_model_scene    matlab/trans_model_scene.m:1-41      T_mg, T_sg, alpha
 *   voting_scheme        matlab/voting_scheme.m:1-150         dense N_m x 30 x N_s accumulator, skip = 5,
 *                                                             per-reference peaks > 0.9 * global peak
 *
 * PARITY UNPINNED: neither MATLAB, Octave nor a JVM exists in this image, so this file cannot be checked
 * against the original; it is a timing comparator and a readable spec, not a results oracle.  Two known
 * departures, both stated in SURVEY.md 8(c): the map key is the tuple of the four quantised doubles instead of
 * the first 8 bytes of their SHA-1 (DataHash needs a JVM), and real(acos(x)) for |x| > 1 is written out
 * (0 for x > 1, pi for x < -1).  An interpreter would be orders of magnitude slower than this compiled
 * restatement; the number reported is therefore a LOWER bound on the MATLAB path's run time.
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

/* real(acos(x)): MATLAB returns a complex number outside [-1, 1]; the callers keep the real part
 * (model_description.m:44, voting_scheme.m:49) */
static double real_acos(double x) { return x > 1.0 ? 0.0 : (x < -1.0 ? M_PI : acos(x)); }
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }

/* point_pair_feature.m:3-9 */
static void point_pair_feature(const double *m1, const double *n1, const double *m2, const double *n2, double F[4]) {
    double d[3] = {m2[0] - m1[0], m2[1] - m1[1], m2[2] - m1[2]};
    F[0] = norm3(d);
    F[1] = real_acos(dot3(n1, d) / (norm3(n1) * norm3(d)));
    F[2] = real_acos(dot3(n2, d) / (norm3(n2) * norm3(d)));
    F[3] = real_acos(dot3(n1, n2) / (norm3(n1) * norm3(n2)));
}
/* MATLAB mod(x, y) = x - floor(x / y) * y ; my_discretize.m:3-4 */
static double mmod(double x, double y) { return x - floor(x / y) * y; }
static void my_discretize(const double F[4], double d_dist, double d_angle, double Fd[4]) {
    Fd[0] = F[0] - mmod(F[0], d_dist);
    for (int k = 1; k < 4; k++) Fd[k] = F[k] - mmod(F[k], d_angle);
}

/* ---- containers.Map('KeyType', ..) restated as an open-addressing table keyed on the 4 quantised doubles ---- */
typedef struct { double key[4]; uint32_t first, count, fill; int used; } slot_t;
typedef struct {
    slot_t *slots; size_t cap, nkeys;   /* cap: power of two, kept at most half full */
    uint32_t *pairs;                    /* (m_r, m_i) pairs, grouped by key */
    size_t npairs;
} map_t;

static uint64_t key_hash(const double k[4]) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 4; i++) {
        uint64_t b; double v = k[i] == 0.0 ? 0.0 : k[i];          /* -0 == +0 as a MATLAB key */
        memcpy(&b, &v, 8);
        h = (h ^ b) * 1099511628211ull; h ^= h >> 29;
    }
    return h;
}
static slot_t *map_find(map_t *m, const double k[4], int insert);
static void map_grow(map_t *m) {
    map_t big = *m;
    big.cap = m->cap * 2; big.nkeys = 0;
    big.slots = calloc(big.cap, sizeof(slot_t));
    for (size_t i = 0; i < m->cap; i++)
        if (m->slots[i].used) { slot_t *s = map_find(&big, m->slots[i].key, 1); s->count = m->slots[i].count; }
    free(m->slots);
    *m = big;
}
static slot_t *map_find(map_t *m, const double k[4], int insert) {
    for (;;) {
        size_t i = key_hash(k) & (m->cap - 1);
        for (;;) {
            slot_t *s = &m->slots[i];
            if (s->used) {
                if (s->key[0] == k[0] && s->key[1] == k[1] && s->key[2] == k[2] && s->key[3] == k[3]) return s;
                i = (i + 1) & (m->cap - 1);
                continue;
            }
            if (!insert) return NULL;
            if (2 * (m->nkeys + 1) > m->cap) break;        /* new key, table half full: grow and probe again */
            s->used = 1; memcpy(s->key, k, 32); s->first = s->count = s->fill = 0;
            m->nkeys++;
            return s;
        }
        map_grow(m);
    }
}

/* model_description.m:17-68 : every ordered pair (ii, jj), ii != jj, appended to the list of its key.
 * Two passes (count, then fill) stand for MATLAB's grow-by-concatenation; row order inside a key is the
 * meshgrid order of model_description.m:17-19 (second index fastest). */
static void model_description(map_t *m, const double *mp, const double *mn, int n, double d_dist, double d_angle) {
    size_t total = (size_t)n * n;
    m->cap = 1u << 14; m->nkeys = 0;
    m->slots = calloc(m->cap, sizeof(slot_t));
    m->pairs = malloc((total ? total : 1) * 2 * sizeof(uint32_t));
    m->npairs = 0;
    /* the features of a block of rows are computed by all threads (that is where the time goes: 3 acos, 4 sqrt
     * per pair); the map itself is filled by one thread in MATLAB's order */
    enum { ROWS = 64 };
    double *buf = malloc(sizeof(double) * 4 * (size_t)ROWS * (n ? n : 1));
    for (int pass = 0; pass < 2; pass++) {
        for (int a0 = 0; a0 < n; a0 += ROWS) {
            const int a1 = a0 + ROWS < n ? a0 + ROWS : n;
#pragma omp parallel for schedule(static)
            for (int a = a0; a < a1; a++)
                for (int b = 0; b < n; b++) {
                    double F[4], *Fd = buf + 4 * ((size_t)(a - a0) * n + b);
                    if (a == b) { Fd[0] = NAN; continue; }                      /* model_description.m:37-41 */
                    point_pair_feature(mp + 3 * a, mn + 3 * a, mp + 3 * b, mn + 3 * b, F);
                    my_discretize(F, d_dist, d_angle, Fd);
                }
            for (int a = a0; a < a1; a++)
                for (int b = 0; b < n; b++) {
                    const double *Fd = buf + 4 * ((size_t)(a - a0) * n + b);
                    if (Fd[0] != Fd[0] || Fd[1] != Fd[1] || Fd[2] != Fd[2] || Fd[3] != Fd[3]) continue;   /* :57-59 */
                    slot_t *s = map_find(m, Fd, 1);
                    if (pass == 0) s->count++;
                    else {
                        uint32_t at = s->first + s->fill++;
                        m->pairs[2 * (size_t)at] = (uint32_t)a; m->pairs[2 * (size_t)at + 1] = (uint32_t)b;
                    }
                }
        }
        if (pass == 0) {
            uint32_t run = 0;
            for (size_t i = 0; i < m->cap; i++) if (m->slots[i].used) { m->slots[i].first = run; run += m->slots[i].count; }
            m->npairs = run;
        }
    }
    free(buf);
}
static void map_free(map_t *m) { free(m->slots); free(m->pairs); memset(m, 0, sizeof(*m)); }

/* 4x4 helpers in MATLAB's column-vector convention */
static void mat_mul(const double A[16], const double B[16], double C[16]) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        double s = 0; for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
        C[4 * r + c] = s;
    }
}
static void eye(double T[16]) { memset(T, 0, 128); T[0] = T[5] = T[10] = T[15] = 1; }
static void trans(const double t[3], double T[16]) { eye(T); T[3] = t[0]; T[7] = t[1]; T[11] = t[2]; }
static void roty(double a, double T[16]) { eye(T); T[0] = cos(a); T[2] = sin(a); T[8] = -sin(a); T[10] = cos(a); }
static void rotz(double a, double T[16]) { eye(T); T[0] = cos(a); T[1] = -sin(a); T[4] = sin(a); T[5] = cos(a); }
static void rotx(double a, double T[16]) { eye(T); T[5] = cos(a); T[6] = -sin(a); T[9] = sin(a); T[10] = cos(a); }
static void apply(const double T[16], const double p[3], double out[3]) {
    for (int r = 0; r < 3; r++) out[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3];
}
/* trans_model_scene.m:12-16 (and :23-27): T_g = rotz(-atan2(n'.y, n'.x)) * roty(atan2(n.z, n.x)) * trans(-p) */
static void frame(const double p[3], const double n[3], double T[16]) {
    double neg[3] = {-p[0], -p[1], -p[2]}, Tr[16], Ry[16], Rz[16], tmp[16];
    trans(neg, Tr);
    roty(atan2(n[2], n[0]), Ry);
    double nt0 = Ry[0] * n[0] + Ry[1] * n[1] + Ry[2] * n[2] + Ry[3];
    double nt1 = Ry[4] * n[0] + Ry[5] * n[1] + Ry[6] * n[2] + Ry[7];
    rotz(-atan2(nt1, nt0), Rz);
    mat_mul(Rz, Ry, tmp);
    mat_mul(tmp, Tr, T);
}
/* trans_model_scene.m:29-39 */
static double trans_model_scene(const double *m_r, const double *n_r_m, const double *m_i, const double *s_r,
                                const double *n_r_s, const double *s_i, double T_m_g[16], double T_s_g[16]) {
    frame(m_r, n_r_m, T_m_g);
    frame(s_r, n_r_s, T_s_g);
    double u[3], v[3];
    apply(T_m_g, m_i, u);
    apply(T_s_g, s_i, v);
    u[0] = 0; v[0] = 0;                                   /* u - w w'u with w = [1 0 0]' */
    double cx = u[1] * v[2] - u[2] * v[1];
    return atan2(cx, u[1] * v[1] + u[2] * v[2]);
}

/* voting_scheme.m:8-95 for the reference points k0, k0 + kstep, ... of the skip grid (every `skip`-th scene
 * point is a reference point, :10-12).  acc = one N_m x 30 slice of MATLAB's accumulator.  Returns the votes
 * cast; *peak / *peak_row / *peak_col = max of the slice (:87-92). */
static uint64_t vote_reference_point(const map_t *map, const double *mp, const double *mn, int nm, const double *sp,
                                     const double *sn, int ns, int r, double d_dist, double d_angle, uint32_t *acc,
                                     uint32_t *peak, int *peak_row, int *peak_col, int scene_stride, int scene_phase) {
    uint64_t votes = 0;
    for (int i = scene_phase; i < ns; i += scene_stride) {
        if (i == r) continue;                                                    /* :38-40 */
        double F[4], Fd[4];
        point_pair_feature(sp + 3 * r, sn + 3 * r, sp + 3 * i, sn + 3 * i, F);
        my_discretize(F, d_dist, d_angle, Fd);
        slot_t *s = map_find((map_t *)map, Fd, 0);                               /* isKey, :55 */
        if (!s) continue;
        for (uint32_t j = 0; j < s->count; j++) {                                /* :60-83 */
            const uint32_t a = map->pairs[2 * (size_t)(s->first + j)], b = map->pairs[2 * (size_t)(s->first + j) + 1];
            double Tm[16], Ts[16];
            double alpha = trans_model_scene(mp + 3 * a, mn + 3 * a, mp + 3 * b, sp + 3 * r, sn + 3 * r, sp + 3 * i, Tm, Ts);
            double alpha_disc = alpha + M_PI - mmod(alpha + M_PI, d_angle);      /* :69 */
            int alpha_ind = (int)fmin(round(alpha_disc / d_angle) + 1, N_ANGLE); /* :70, 1-based */
            acc[(size_t)a * N_ANGLE + (alpha_ind - 1)]++;
            votes++;
        }
    }
    uint32_t best = 0; int br = 0, bc = 0;
    for (int a = 0; a < nm; a++) for (int c = 0; c < N_ANGLE; c++)
        if (acc[(size_t)a * N_ANGLE + c] > best) { best = acc[(size_t)a * N_ANGLE + c]; br = a; bc = c; }
    *peak = best; *peak_row = br; *peak_col = bc;
    return votes;
}

/* drost.m:60-109 in one call: model_description, voting_scheme (skip = 5, accum_thresh = 0.9), and the pose
 * inv(T_sg) * rotx(alpha) * T_mg of the best peak (alpha = the bin's lower edge).
 * max_refs > 0 bounds the number of reference points (spread over the skip grid), scene_stride > 1 subsamples
 * the "other" scene points: both only for timing a bounded sample (bench.py cpu_baseline).
 * d_dist <= 0: model_description.m:5-13 (0.1 * max distance from the bounding-box centre).
 * Returns seconds spent voting; *build_seconds = model_description. */
double drost_m_run(const float *mxyz, const float *mnrm, int nm, const float *sxyz, const float *snrm, int ns,
                   double d_dist, int skip, int max_refs, int scene_stride, int threads, uint64_t *pairs_out,
                   uint64_t *votes_out, double *build_seconds, double *pose_out /* 16, row-major, may be NULL */) {
    if (skip < 1) skip = 5;
    if (scene_stride < 1) scene_stride = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double *mp = malloc(sizeof(double) * 3 * (nm + 1)), *mn = malloc(sizeof(double) * 3 * (nm + 1));
    double *sp = malloc(sizeof(double) * 3 * (ns + 1)), *sn = malloc(sizeof(double) * 3 * (ns + 1));
    for (int i = 0; i < 3 * nm; i++) { mp[i] = mxyz[i]; mn[i] = mnrm[i]; }
    for (int i = 0; i < 3 * ns; i++) { sp[i] = sxyz[i]; sn[i] = snrm[i]; }
    const double d_angle = 2 * M_PI / N_ANGLE;
    if (!(d_dist > 0)) {                                                          /* model_description.m:5-13 */
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, mx = 0;
        for (int i = 0; i < nm; i++) for (int c = 0; c < 3; c++) { lo[c] = fmin(lo[c], mp[3 * i + c]); hi[c] = fmax(hi[c], mp[3 * i + c]); }
        for (int i = 0; i < nm; i++) {
            double d[3] = {mp[3 * i] - (lo[0] + hi[0]) / 2, mp[3 * i + 1] - (lo[1] + hi[1]) / 2, mp[3 * i + 2] - (lo[2] + hi[2]) / 2};
            mx = fmax(mx, norm3(d));
        }
        d_dist = 0.1 * mx;
    }
    double t0 = now_s();
    map_t map;
    model_description(&map, mp, mn, nm, d_dist, d_angle);
    double t1 = now_s();
    if (build_seconds) *build_seconds = t1 - t0;

    const int R_all = ns > 0 ? (ns + skip - 1) / skip : 0;
    const int R = (max_refs > 0 && R_all > max_refs) ? max_refs : R_all;
    uint64_t votes = 0, pairs = 0;
    uint32_t g_peak = 0; int g_ref = -1, g_row = 0, g_col = 0;
#pragma omp parallel reduction(+ : votes, pairs)
    {
        uint32_t *acc = calloc((size_t)(nm ? nm : 1) * N_ANGLE, sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < R; k++) {
            const int r = (int)(((long long)k * R_all) / R) * skip;              /* r_indices = 1:skip:N, :10-11 */
            uint32_t peak; int pr, pc;
            votes += vote_reference_point(&map, mp, mn, nm, sp, sn, ns, r, d_dist, d_angle, acc, &peak, &pr, &pc,
                                          scene_stride, k % scene_stride);
            pairs += (uint64_t)((ns - (k % scene_stride) + scene_stride - 1) / scene_stride);
#pragma omp critical
            if (peak > g_peak || (peak == g_peak && g_ref >= 0 && r < g_ref)) { g_peak = peak; g_ref = r; g_row = pr; g_col = pc; }
            memset(acc, 0, (size_t)nm * N_ANGLE * sizeof(uint32_t));
        }
        free(acc);
    }
    double t2 = now_s();
    if (pose_out) {
        memset(pose_out, 0, 16 * sizeof(double));
        if (g_ref >= 0 && g_peak > 0) {
            /* drost.m:93-100: T = inv(T_s_g) * rotx(alpha) * T_m_g for the best (model point, alpha bin, reference) */
            double Tm[16], Ts[16], Rx[16], inv[16], tmp[16];
            frame(mp + 3 * g_row, mn + 3 * g_row, Tm);
            frame(sp + 3 * g_ref, sn + 3 * g_ref, Ts);
            rotx(g_col * d_angle - M_PI, Rx);
            eye(inv);                                       /* rigid inverse: R', -R't */
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) inv[4 * r + c] = Ts[4 * c + r];
            for (int r = 0; r < 3; r++) inv[4 * r + 3] = -(inv[4 * r] * Ts[3] + inv[4 * r + 1] * Ts[7] + inv[4 * r + 2] * Ts[11]);
            mat_mul(inv, Rx, tmp);
            mat_mul(tmp, Tm, pose_out);
        }
    }
    if (pairs_out) *pairs_out = pairs;
    if (votes_out) *votes_out = votes;
    map_free(&map);
    free(mp); free(mn); free(sp); free(sn);
    return t2 - t1;
}
