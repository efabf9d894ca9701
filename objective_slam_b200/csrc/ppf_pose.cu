// ppf_pose.cu -- pose from vote, vote weighting and GPU pose clustering.
// Replaces trans_calc_kernel2 (kernel.cu:605-645) with compute_rot_angles /
// compute_transforms (kernel.cu:352-401), vote_weight_kernel (kernel.cu:766-782),
// mat2transquat_kernel (kernel.cu:647-661), trans2idx_kernel (kernel.cu:663-699),
// the second ParallelHashArray + GetIndices (model.cu:222-226) and
// rot_clustering_kernel (kernel.cu:702-763), plus thrust::max_element (model.cu:293-295).
//
// K (surviving votes) is 10^2..10^5, so these kernels are latency-bound; what matters
// is that they reproduce the reference's arithmetic and its quirks:
//   * every kernel is a no-op when K <= 1 (kernel.cu:609,651,667,712,769) -> zero pose;
//   * the all-zero vote code is skipped (kernel.cu:628-631);
//   * the pose uses the lower edge of the alpha bin: Rx(alpha_idx*D_ANGLE0 - pi), an FFMA;
//   * quaternion normalised by sqrt(norm(q)) (kernel.cu:138);
//   * cells are matched by 32-bit FNV hash of (int)(quant_down(t)/d_dist), the pose's own
//     cell is excluded (hash forced to 0, kernel.cu:684-689) and hash 0 means "skip".
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

namespace ppf {

// invht (kernel.cu:254-299)
__device__ __forceinline__ void mat4_invht(const Mat4 &T, Mat4 &Ti) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Ti.m[i][j] = T.m[j][i];
    float tx = T.m[0][3], ty = T.m[1][3], tz = T.m[2][3];
#pragma unroll
    // -R' t.  nvcc rewrites the negated dot product as ((-r1*ty) - r0*tx) - r2*tz and ptxas fuses it
    // into fma(-r2, tz, fma(-r1, ty, -(r0*tx))) (reference SASS of trans_calc_kernel2): the plain
    // product is the x term here, unlike every other dot product of the reference.
    for (int i = 0; i < 3; i++)
        Ti.m[i][3] = __fmaf_rn(-Ti.m[i][2], tz, __fmaf_rn(-Ti.m[i][1], ty, -__fmul_rn(Ti.m[i][0], tx)));
    Ti.m[3][0] = 0; Ti.m[3][1] = 0; Ti.m[3][2] = 0; Ti.m[3][3] = 1;
}

__global__ void pose_kernel(const unsigned long long *__restrict__ votes, const float4 *mpos, const float4 *mnrm,
                            const float4 *spos, const float4 *snrm, const uint32_t *__restrict__ sinv, int ns,
                            float *transforms, int count) {
    if (count <= 1) return;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        unsigned long long v = votes[idx];
        uint32_t s = (uint32_t)(v >> 32), mac = (uint32_t)v, m = mac >> 6, a = mac & 63u;
        if (s == 0 && m == 0 && a == 0) continue;
        // survivors merged from other ranks are trusted to name valid points; guard against garbage anyway
        if (s >= (uint32_t)ns) continue;
        const uint32_t sp_i = sinv ? sinv[s] : s;                   // scene clouds are stored in Morton order
        float4 mn = mnrm[m], sn = snrm[sp_i], mp = mpos[m], sp = spos[sp_i];
        float m_roty, m_rotz, s_roty, s_rotz;
        frame_angles(mn.x, mn.y, mn.z, m_roty, m_rotz);
        frame_angles(sn.x, sn.y, sn.z, s_roty, s_rotz);
        Mat4 Tmg, Tsg, Rx, Tinv, T2, T;
        frame_from_angles(mp.x, mp.y, mp.z, m_roty, m_rotz, Tmg);
        frame_from_angles(sp.x, sp.y, sp.z, s_roty, s_rotz, Tsg);
        mat4_rotx(__fmaf_rn((float)a, d_angle0(), -CUDART_PI_F), Rx);
        mat4_invht(Tsg, Tinv);
        mat4_mul(Tinv, Rx, T2);
        mat4_mul(T2, Tmg, T);
        float *out = transforms + (size_t)idx * 16;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) out[i * 4 + j] = T.m[i][j];
    }
}

__global__ void weight_kernel(const unsigned long long *votes, const uint32_t *counts, const float *weights,
                              float *weighted, int count) {
    if (count <= 1) return;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        uint32_t m = ((uint32_t)votes[idx]) >> 6;
        weighted[idx] = __fmul_rn(weights[m], (float)counts[idx]);
    }
}

// mat2transquat_kernel + hrotmat2quat (kernel.cu:128-144)
__global__ void transquat_kernel(const float *T, float3 *trans, float4 *rots, int count) {
    if (count <= 1) return;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const float *M = T + (size_t)idx * 16;
        trans[idx] = make_float3(M[3], M[7], M[11]);
        float T00 = M[0], T11 = M[5], T22 = M[10];
        float t = __fadd_rn(__fadd_rn(T00, T11), T22);
        float4 q;
        q.x = __fmul_rn(0.5f, sqrt_approx_ftz(__fadd_rn(1.0f, t)));
        q.y = copysignf(__fmul_rn(0.5f, sqrt_approx_ftz(__fsub_rn(__fsub_rn(__fadd_rn(1.0f, T00), T11), T22))),
                        __fsub_rn(M[9], M[6]));
        q.z = copysignf(__fmul_rn(0.5f, sqrt_approx_ftz(__fsub_rn(__fadd_rn(__fsub_rn(1.0f, T00), T11), T22))),
                        __fsub_rn(M[2], M[8]));
        q.w = copysignf(__fmul_rn(0.5f, sqrt_approx_ftz(__fadd_rn(__fsub_rn(__fsub_rn(1.0f, T00), T11), T22))),
                        __fsub_rn(M[4], M[1]));
        float n = sqrt_approx_ftz(sqrt_approx_ftz(dot4(q.x, q.y, q.z, q.w, q.x, q.y, q.z, q.w)));
        q.x = div_full_ftz(q.x, n); q.y = div_full_ftz(q.y, n);
        q.z = div_full_ftz(q.z, n); q.w = div_full_ftz(q.w, n);
        rots[idx] = q;
    }
}

__device__ __forceinline__ int trans_cell(float x, float d) {
    float disc = __fsub_rn(x, fmodf(x, d));                 // quant_downf, kernel.cu:90-92 (x may be negative)
    return __float2int_rz(div_full_ftz(disc, d));           // (int)(disc / d_dist), kernel.cu:676-678
}

// trans2idx_kernel (kernel.cu:663-699)
__global__ void cellhash_kernel(const float3 *trans, uint32_t *cell_hash, uint32_t *adj_hash, int count, float d) {
    if (count <= 1) return;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        float3 t = trans[idx];
        int cx = trans_cell(t.x, d), cy = trans_cell(t.y, d), cz = trans_cell(t.z, d);
        cell_hash[idx] = fnv1a_3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
        int c = 0;
        for (int i = -1; i < 2; i++)
            for (int j = -1; j < 2; j++)
                for (int k = -1; k < 2; k++, c++) {
                    uint32_t h = 0;
                    if (!(i == 0 && j == 0 && k == 0))
                        h = fnv1a_3((uint32_t)(cx + i), (uint32_t)(cy + j), (uint32_t)(cz + k));
                    adj_hash[27 * (size_t)idx + c] = h;
                }
    }
}

// rot_clustering_kernel (kernel.cu:702-763). sorted_hash/sorted_idx: poses ordered by
// (cell hash, pose index) -- the ParallelHashArray of model.cu:222-223.
// One WARP per pose: the 32 lanes test 32 neighbour poses at a time (quaternion distance,
// translation distance), then the accepted weights are added to the score one by one in lane order,
// i.e. in exactly the order the reference's single thread visits them (neighbour cell 0..26, then
// ascending pose index inside the cell) -- float addition is not associative once a score passes
// 2^24, so the order is part of the result.
// The neighbour poses are read from CELL-SORTED copies (sq / st = quaternion and (translation, weight) of the pose at
// sorted position p, pose_gather_kernel): a dense cell is then one contiguous run of 16-byte records instead of three
// gathers per candidate through sorted_idx (K = 396k survivors: the gathers were most of the kernel's time).
__global__ void pose_gather_kernel(const uint32_t *__restrict__ sorted_idx, const float4 *__restrict__ quats,
                                   const float3 *__restrict__ trans, const float *__restrict__ weights, int count,
                                   float4 *sq, float4 *st) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < count; p += gridDim.x * blockDim.x) {
        const uint32_t j = sorted_idx[p];
        const float3 t = trans[j];
        sq[p] = quats[j];
        st[p] = make_float4(t.x, t.y, t.z, weights[j]);
    }
}

__global__ void cluster_kernel(const float3 *trans_in, const float4 *quats, const float4 *__restrict__ sq,
                               const float4 *__restrict__ st,
                               const uint32_t *adj_hash, const uint32_t *sorted_hash,
                               float *scores, float3 *trans_out, int count, float trans_thresh, int use_l1_norm,
                               int use_averaged_clusters, int shard, int n_shards) {
    if (count <= 1) return;
    const float rot_thresh = 2 * d_angle0();                                  // ROT_THRESH, kernel.h:17
    const float rot_thresh_sq = rot_thresh * rot_thresh;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    // poses idx = shard, shard + n_shards, ... (multi-GPU: every rank scores an interleaved slice of the merged list)
    for (int idx = shard + n_shards * ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); idx < count; idx += n_shards * warps) {
        const float3 tt = trans_in[idx];
        const float4 q = quats[idx];
        float score = 1;
        float3 to = tt;
        for (int ab = 0; ab < 27; ab++) {
            const uint32_t h = adj_hash[27 * (size_t)idx + ab];
            if (h == 0) continue;
            int lo = 0, hi = count;                                           // lower_bound over the cell hashes
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (sorted_hash[mid] < h) lo = mid + 1; else hi = mid;
            }
            for (int base = lo; base < count; base += 32) {
                const int p = base + lane;
                const bool valid = p < count && sorted_hash[p] == h;
                const unsigned run = __ballot_sync(0xffffffffu, valid);
                if (!run) break;
                bool pass = false;
                float w = 0.f;
                float3 tj = make_float3(0.f, 0.f, 0.f);
                if (valid) {
                    const float4 qj = sq[p];
                    const float qd = fabsf(__fmul_rn(8.0f, __fsub_rn(1.0f, dot4(q.x, q.y, q.z, q.w, qj.x, qj.y, qj.z, qj.w))));
                    if (qd < rot_thresh_sq) {
                        const float4 tw = st[p];
                        tj = make_float3(tw.x, tw.y, tw.z);
                        w = tw.w;
                        pass = true;
                        if (!use_l1_norm) {
                            const float nd = norm3(__fsub_rn(tt.x, tj.x), __fsub_rn(tt.y, tj.y), __fsub_rn(tt.z, tj.z));
                            pass = nd < trans_thresh;
                        }
                    }
                }
                unsigned acc = __ballot_sync(0xffffffffu, pass);
                while (acc) {                                                 // sequential, reference order
                    const int src = __ffs(acc) - 1;
                    acc &= acc - 1;
                    const float wj = __shfl_sync(0xffffffffu, w, src);
                    if (use_averaged_clusters) {                              // kernel.cu:747-752
                        const float tx = __shfl_sync(0xffffffffu, tj.x, src), ty = __shfl_sync(0xffffffffu, tj.y, src),
                                    tz = __shfl_sync(0xffffffffu, tj.z, src);
                        to.x = __fmul_rn(score, to.x); to.y = __fmul_rn(score, to.y); to.z = __fmul_rn(score, to.z);
                        to.x = __fadd_rn(to.x, __fmul_rn(wj, tx));
                        to.y = __fadd_rn(to.y, __fmul_rn(wj, ty));
                        to.z = __fadd_rn(to.z, __fmul_rn(wj, tz));
                        const float inv = div_full_ftz(1.0f, __fadd_rn(score, wj));
                        to.x = __fmul_rn(inv, to.x); to.y = __fmul_rn(inv, to.y); to.z = __fmul_rn(inv, to.z);
                    }
                    score = __fadd_rn(score, wj);
                }
                if (run != 0xffffffffu) break;                                // the run of equal hashes ended here
            }
        }
        if (lane == 0) { scores[idx] = score; trans_out[idx] = to; }
    }
}

// first index of the maximum (thrust::max_element semantics)
__global__ void argmax_kernel(const float *scores, int count, uint32_t *out) {
    __shared__ float sv[1024];
    __shared__ int si[1024];
    float best = -CUDART_INF_F; int bi = 0x7FFFFFFF;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        float s = scores[i];
        if (s > best || (s == best && i < bi)) { best = s; bi = i; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            float s = sv[threadIdx.x + o]; int i = si[threadIdx.x + o];
            if (s > sv[threadIdx.x] || (s == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = s; si[threadIdx.x] = i; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (count > 0 && si[0] != 0x7FFFFFFF) ? (uint32_t)si[0] : 0u;
}

__global__ void iota_u32_kernel(uint32_t *v, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}

// ---- operator-level kernels (MATLAB function names) ------------------------------------------------
__global__ void pair_feature_kernel(const float *p1, const float *n1, const float *p2, const float *n2, size_t n,
                                    float d_dist, float4 *raw, float4 *disc, uint32_t *keys) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        PointN a, b;
        a.x = p1[3 * i]; a.y = p1[3 * i + 1]; a.z = p1[3 * i + 2];
        a.nx = n1[3 * i]; a.ny = n1[3 * i + 1]; a.nz = n1[3 * i + 2]; a.nn = norm3(a.nx, a.ny, a.nz);
        b.x = p2[3 * i]; b.y = p2[3 * i + 1]; b.z = p2[3 * i + 2];
        b.nx = n2[3 * i]; b.ny = n2[3 * i + 1]; b.nz = n2[3 * i + 2]; b.nn = norm3(b.nx, b.ny, b.nz);
        if (raw) {                                                      // compute_ppf, kernel.cu:109-122
            float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y), dz = __fsub_rn(b.z, a.z);
            float nd = norm3(dx, dy, dz);
            raw[i] = make_float4(nd, acosf(div_full_ftz(dot3(a.nx, a.ny, a.nz, dx, dy, dz), __fmul_rn(nd, a.nn))),
                                 acosf(div_full_ftz(dot3(b.nx, b.ny, b.nz, dx, dy, dz), __fmul_rn(nd, b.nn))),
                                 acosf(div_full_ftz(dot3(a.nx, a.ny, a.nz, b.nx, b.ny, b.nz), __fmul_rn(a.nn, b.nn))));
        }
        FeatureBins fb = pair_feature_bins(a, b, d_dist, 1.0f / d_dist);
        float4 q;
        if (fb.kd < 0) q.x = CUDART_NAN_F;
        else if (fb.kd == 0x7FFFFFFF) q.x = __fsub_rn(fb.f1, fmodf(fb.f1, d_dist));
        else q.x = quant_value(fb.kd, d_dist);
        q.y = __uint_as_float(angle_bits(fb.k1)); q.z = __uint_as_float(angle_bits(fb.k2)); q.w = __uint_as_float(angle_bits(fb.k3));
        if (disc) disc[i] = q;
        if (keys) keys[i] = (q.x != q.x) ? 0u : fnv1a_4(__float_as_uint(q.x), __float_as_uint(q.y), __float_as_uint(q.z), __float_as_uint(q.w));
    }
}

__global__ void trans_model_scene_kernel(const float *m_r, const float *n_r_m, const float *m_i, const float *s_r,
                                         const float *n_r_s, const float *s_i, size_t n, float *Tmg, float *Tsg,
                                         float *alpha, uint32_t *alpha_idx) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float ry, rz;
        Mat4 A, B;
        frame_angles(n_r_m[3 * i], n_r_m[3 * i + 1], n_r_m[3 * i + 2], ry, rz);
        frame_from_angles(m_r[3 * i], m_r[3 * i + 1], m_r[3 * i + 2], ry, rz, A);
        frame_angles(n_r_s[3 * i], n_r_s[3 * i + 1], n_r_s[3 * i + 2], ry, rz);
        frame_from_angles(s_r[3 * i], s_r[3 * i + 1], s_r[3 * i + 2], ry, rz, B);
        float uy = mat4_row_apply(A.m[1], m_i[3 * i], m_i[3 * i + 1], m_i[3 * i + 2]);
        float uz = mat4_row_apply(A.m[2], m_i[3 * i], m_i[3 * i + 1], m_i[3 * i + 2]);
        float vy = mat4_row_apply(B.m[1], s_i[3 * i], s_i[3 * i + 1], s_i[3 * i + 2]);
        float vz = mat4_row_apply(B.m[2], s_i[3 * i], s_i[3 * i + 1], s_i[3 * i + 2]);
        if (Tmg) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tmg[16 * i + 4 * r + c] = A.m[r][c];
        if (Tsg) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tsg[16 * i + 4 * r + c] = B.m[r][c];
        if (alpha) alpha[i] = atan2f(__fmaf_rn(uy, vz, -__fmul_rn(uz, vy)), __fmaf_rn(uz, vz, __fmaf_rn(uy, vy, 0.0f)));
        if (alpha_idx) alpha_idx[i] = alpha_bin_exact(uy, uz, vy, vz);
    }
}

// host arrays in, host arrays out (debug / operator-level API; not a hot path)
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { pooled_free(p); }
    int up(const void *h, size_t bytes) {
        if (pooled_malloc(&p, bytes ? bytes : 4) != cudaSuccess) return PPF_ERR_CUDA;
        if (h && bytes && memcpy_sync(p, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return PPF_ERR_CUDA;
        return PPF_OK;
    }
    int down(void *h, size_t bytes) { return (h && bytes && memcpy_sync(h, p, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) ? PPF_ERR_CUDA : PPF_OK; }
};

int op_point_pair_feature(const float *p1, const float *n1, const float *p2, const float *n2, size_t n, float d_dist,
                          float *raw_out, float *disc_out, uint32_t *keys_out) {
    if (!p1 || !n1 || !p2 || !n2 || !(d_dist > 0.f)) { set_last_error("point_pair_feature: NULL input or d_dist <= 0"); return PPF_ERR_INVALID; }
    if (n == 0) return PPF_OK;
    DevBuf a, b, c, d, r, q, k;
    if (a.up(p1, n * 12) || b.up(n1, n * 12) || c.up(p2, n * 12) || d.up(n2, n * 12) || r.up(nullptr, n * 16) ||
        q.up(nullptr, n * 16) || k.up(nullptr, n * 4)) { set_last_error("point_pair_feature: device allocation/copy failed"); return PPF_ERR_CUDA; }
    pair_feature_kernel<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(
        (const float *)a.p, (const float *)b.p, (const float *)c.p, (const float *)d.p, n, d_dist,
        raw_out ? (float4 *)r.p : nullptr, disc_out ? (float4 *)q.p : nullptr, keys_out ? (uint32_t *)k.p : nullptr);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    if (r.down(raw_out, n * 16) || q.down(disc_out, n * 16) || k.down(keys_out, n * 4)) { set_last_error("point_pair_feature: copy back failed"); return PPF_ERR_CUDA; }
    return PPF_OK;
}

int op_trans_model_scene(const float *m_r, const float *n_r_m, const float *m_i, const float *s_r, const float *n_r_s,
                         const float *s_i, size_t n, float *T_m_g, float *T_s_g, float *alpha, uint32_t *alpha_idx) {
    if (!m_r || !n_r_m || !m_i || !s_r || !n_r_s || !s_i) { set_last_error("trans_model_scene: NULL input"); return PPF_ERR_INVALID; }
    if (n == 0) return PPF_OK;
    DevBuf in[6], tm, ts, al, ai;
    const float *src[6] = {m_r, n_r_m, m_i, s_r, n_r_s, s_i};
    for (int i = 0; i < 6; i++) if (in[i].up(src[i], n * 12)) { set_last_error("trans_model_scene: upload failed"); return PPF_ERR_CUDA; }
    if (tm.up(nullptr, n * 64) || ts.up(nullptr, n * 64) || al.up(nullptr, n * 4) || ai.up(nullptr, n * 4)) { set_last_error("trans_model_scene: allocation failed"); return PPF_ERR_CUDA; }
    trans_model_scene_kernel<<<(int)std::min<size_t>((n + 127) / 128, 148 * 8), 128, 0, cur_stream()>>>(
        (const float *)in[0].p, (const float *)in[1].p, (const float *)in[2].p, (const float *)in[3].p, (const float *)in[4].p,
        (const float *)in[5].p, n, T_m_g ? (float *)tm.p : nullptr, T_s_g ? (float *)ts.p : nullptr,
        alpha ? (float *)al.p : nullptr, alpha_idx ? (uint32_t *)ai.p : nullptr);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    if (tm.down(T_m_g, n * 64) || ts.down(T_s_g, n * 64) || al.down(alpha, n * 4) || ai.down(alpha_idx, n * 4)) { set_last_error("trans_model_scene: copy back failed"); return PPF_ERR_CUDA; }
    return PPF_OK;
}

static int blocks_for(size_t count) { return (int)std::min<size_t>(std::max<size_t>((count + 255) / 256, 1), 1024); }

int poses_run(const ModelTable &m, const Cloud &scene, VoteResult &r) {
    const int K = (int)r.K;
    if (K == 0) return PPF_OK;
    PPF_CUDA_TRY(cudaMemsetAsync(r.transformations, 0, (size_t)K * 64, cur_stream()));
    PPF_CUDA_TRY(cudaMemsetAsync(r.weighted, 0, (size_t)K * 4, cur_stream()));
    pose_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(r.codes, m.cloud.pos, m.cloud.nrm, scene.pos, scene.nrm, scene.inv, scene.n,
                                        r.transformations, K);
    count_launch();
    weight_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(r.codes, r.counts, m.weights, r.weighted, K);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    return PPF_OK;
}

int cluster_finish(VoteResult &r) {
    const int K = (int)r.K;
    r.max_idx = 0;
    if (K <= 1) return PPF_OK;
    uint32_t *d_arg = nullptr;
    PPF_CUDA_TRY(pooled_malloc(&d_arg, 4));
    argmax_kernel<<<1, 1024, 0, cur_stream()>>>(r.scores, K, d_arg);
    count_launch();
    cudaError_t e = memcpy_sync(&r.max_idx, d_arg, 4, cudaMemcpyDeviceToHost);
    pooled_free(d_arg);
    PPF_CUDA_TRY(e);
    return PPF_OK;
}

// Clustering of the (merged) survivor list.  With several ranks Model::ClusterTransformations is sharded: it is
// quadratic in dense cells (6 ms at K = 50k, 332 ms at K = 396k on one GPU), so every rank scores an interleaved
// slice of the poses against all of them, the slices are summed (entries outside a slice are 0: exact) and every
// rank picks the same winner.  use_averaged_clusters needs every pose's averaged translation: replicated then.
int cluster_dist(const ModelTable &m, Comm *comm, VoteResult &r) {
    const int world = comm ? comm->world : 1;
    if (world == 1 || m.use_averaged_clusters) return cluster_run(m, r);
    int rc = cluster_run(m, r, comm->rank, world);
    if (rc) return rc;
    if (r.K > 1 && (rc = comm->allreduce_sum_f32(r.scores, r.K))) return rc;
    return cluster_finish(r);
}

// shard / n_shards: score only the poses idx = shard (mod n_shards); the other scores stay 0 and max_idx is not
// computed (cluster_finish does that once the slices of all ranks have been summed).  n_shards = 1: everything.
int cluster_run(const ModelTable &m, VoteResult &r, int shard, int n_shards) {
    const int K = (int)r.K;
    r.max_idx = 0;
    if (K == 0) return PPF_OK;
    PPF_CUDA_TRY(cudaMemsetAsync(r.trans, 0, (size_t)K * sizeof(float3), cur_stream()));
    PPF_CUDA_TRY(cudaMemsetAsync(r.rots, 0, (size_t)K * sizeof(float4), cur_stream()));
    PPF_CUDA_TRY(cudaMemsetAsync(r.scores, 0, (size_t)K * 4, cur_stream()));
    if (K <= 1) return PPF_OK;
    transquat_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(r.transformations, r.trans, r.rots, K);
    count_launch();
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, K);
    int rc = r.ws.reserve((size_t)K * (4 + 27 * 4 + 4 + 4 + 4 + sizeof(float3) + 2 * sizeof(float4)) + 4 + tb + 1024);
    if (rc) return rc;
    uint32_t *cell = r.ws.take<uint32_t>(K), *adj = r.ws.take<uint32_t>((size_t)K * 27);
    uint32_t *iota = r.ws.take<uint32_t>(K), *shash = r.ws.take<uint32_t>(K), *sidx = r.ws.take<uint32_t>(K);
    float3 *tin = r.ws.take<float3>(K);
    float4 *sq = r.ws.take<float4>(K), *st = r.ws.take<float4>(K);
    uint32_t *d_arg = r.ws.take<uint32_t>(1);
    void *tmp = r.ws.take_bytes(tb);
    if (!cell || !adj || !iota || !shash || !sidx || !tin || !sq || !st || !d_arg || !tmp) {
        set_last_error("cluster: workspace too small");
        return PPF_ERR_CUDA;
    }
    cellhash_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(r.trans, cell, adj, K, m.d_dist);
    count_launch();
    iota_u32_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(iota, K);
    count_launch();
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, cell, shash, iota, sidx, K, 0, 32, cur_stream()));
    // rot_clustering_kernel updates translations in place while neighbours read them (a race in the
    // reference when use_averaged_clusters is set); we read a snapshot instead, which is deterministic.
    PPF_CUDA_TRY(cudaMemcpyAsync(tin, r.trans, (size_t)K * sizeof(float3), cudaMemcpyDeviceToDevice, cur_stream()));
    pose_gather_kernel<<<blocks_for(K), 256, 0, cur_stream()>>>(sidx, r.rots, tin, r.weighted, K, sq, st);
    count_launch();
    const size_t mine = ((size_t)K + n_shards - 1) / n_shards;
    cluster_kernel<<<(int)std::min<size_t>((mine * 32 + 255) / 256, 148 * 64), 256, 0, cur_stream()>>>(tin, r.rots, sq, st, adj, shash, r.scores, r.trans, K,
                                           m.d_dist, m.use_l1_norm, m.use_averaged_clusters, shard, n_shards);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    if (n_shards > 1) return PPF_OK;
    argmax_kernel<<<1, 1024, 0, cur_stream()>>>(r.scores, K, d_arg);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    PPF_CUDA_TRY(memcpy_sync(&r.max_idx, d_arg, 4, cudaMemcpyDeviceToHost));
    return PPF_OK;
}

}  // namespace ppf
