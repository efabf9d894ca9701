// ppf_vote.cu -- voting_scheme: scene pair generation + quantise + probe + Hough vote,
// fused in one kernel, then threshold + ordering of the surviving votes.
// Replaces Scene::Scene's ppf_kernel/ppf_hash_kernel pass (scene.cu:24-55),
// ParallelHashArray::GetIndices (parallel_hash_array.hpp:81-92), ppf_vote_count_kernel,
// ppf_vote_kernel (kernel.cu:480-554) and the sort/histogram/threshold tail of
// Model::ComputeUniqueVotes (model.cu:148-170).
//
// The reference materialises N_s^2 features, N_s^2 keys, one 8-byte code per vote, and
// sorts the codes.  Here nothing of size N_s^2 or #votes ever reaches HBM:
//   * one CTA owns one (scene reference point s_r, model chunk c) pair and keeps the
//     31 x chunk_rows accumulator of that reference point in shared memory;
//   * phase 1 streams the scene points (coalesced float4 loads), computes the feature
//     bins of (s_r, s_i), looks the cell up in the model's cell table and pushes the
//     hits (bucket slice, alpha_s) into a shared-memory queue with one warp-aggregated
//     atomic per warp;
//   * phase 2 lets each warp drain hits: 32 lanes read 32 consecutive 4-byte bucket
//     entries (one 128 B line), derive the alpha bin from two 20-bit binary angles and
//     add to the shared accumulator with ATOMS.  A vote whose angle falls within a
//     guard band of a bin edge recomputes alpha exactly as trans_model_scene does
//     (kernel.cu:302-342), which keeps every count bit-exact;
//   * phase 3 emits the cells that can still pass the reference's global threshold
//     (count > thr * max, model.cu:164-167) using a monotone lower bound of the max.
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

#include "ppf_vote_common.cuh"

namespace ppf {

template <int THREADS>
__global__ void __launch_bounds__(THREADS) vote_kernel(const VoteArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.chunk_rows, S = acc_stride(C);
    uint4 *queue = reinterpret_cast<uint4 *>(smem_raw);                               // [kHitQueue] hit records
    uint32_t *gstart = reinterpret_cast<uint32_t *>(queue + kHitQueue);               // [kHitQueue] first grab ticket of each hit
    uint32_t *acc = gstart + kHitQueue;                                               // [31][S] vote counters
    __shared__ uint32_t s_nhits, s_ticket, s_total, s_exact;
    __shared__ uint32_t s_red[32];
    __shared__ unsigned long long s_votes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x / a.ref_count;
    const int refk = blockIdx.x - chunk * a.ref_count;
    const int s_r = a.ref_start + refk * a.ref_stride;
    const int chunk_base = chunk * C;
    const uint2 *__restrict__ ranges = a.ranges + (size_t)chunk * a.U;

    for (int i = tid; i < kNAlphaBins * S; i += THREADS) acc[i] = 0;
    if (tid == 0) { s_exact = 0; s_votes = 0; }

    // reference point (registers) and its local frame; p_r = its position in the stored order
    const int p_r = (int)__ldg(a.sinv + s_r);
    PointN R;
    {
        float4 p = __ldg(a.spos + p_r), q = __ldg(a.snrm + p_r);
        R.x = p.x; R.y = p.y; R.z = p.z; R.nx = q.x; R.ny = q.y; R.nz = q.z; R.nn = q.w;
    }
    const FrameYZ FS = load_frame(a.sfy, a.sfz, p_r);
    VoteCtx ctx;
    ctx.map = a.map; ctx.mfy = a.mfy; ctx.mfz = a.mfz; ctx.mpos = a.mpos; ctx.spos = a.spos;
    ctx.nm = a.nm; ctx.chunk_base = chunk_base; ctx.stride = S; ctx.acc = acc;
    ctx.acc_addr = 0; ctx.opaque_zero = 0; ctx.rq = nullptr; ctx.rq_cap = 0; ctx.rq_count = nullptr;   // exact path inline
    unsigned long long my_votes = 0;
    uint32_t my_exact = 0;

    for (int base = 0; base < a.ns; base += kHitQueue) {
        // Tile culling (block-uniform): every scene pair (s_r, s_i) with s_i in this tile is longer than
        // any model pair -> its distance bin is outside the table -> no hit; skip the tile altogether.
        if (box_dist2(R, __ldg(a.tbox_lo + base / kHitQueue), __ldg(a.tbox_hi + base / kHitQueue)) >= a.cull_r2) continue;
        if (tid == 0) { s_nhits = 0; s_ticket = 0; }
        __syncthreads();
        // ---- phase 1: pairs (s_r, s_i) of this tile -> hit queue
#pragma unroll 1
        for (int it = 0; it < kHitQueue / THREADS; it++) {
            const int i = base + it * THREADS + tid;
            bool hit = false;
            uint4 h = make_uint4(0, 0, 0, 0);
            // warp culling: the 32 points of this warp are one Morton-compact group with a known AABB
            const bool near = (i - lane) < a.ns &&
                              box_dist2(R, __ldg(a.gbox_lo + (i >> 5)), __ldg(a.gbox_hi + (i >> 5))) < a.cull_r2;
            if (near && i < a.ns && i != p_r) {
                float4 p = __ldg(a.spos + i), q = __ldg(a.snrm + i);
                PointN O;
                O.x = p.x; O.y = p.y; O.z = p.z; O.nx = q.x; O.ny = q.y; O.nz = q.z; O.nn = q.w;
                FeatureBins fb = pair_feature_bins(R, O, a.d_dist, a.inv_d);
                const uint32_t b = probe_bucket(a, fb);
                if (b != kNoBucket) {
                    uint2 rg = __ldg(ranges + b);
                    if (rg.y != 0) {
                        float vy, vz;
                        frame_apply_yz(FS, O.x, O.y, O.z, vy, vz);
                        h = make_uint4(rg.x, rg.y, pack_hit_theta(theta_code(vy, vz)), (uint32_t)i);
                        hit = true;
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(&s_nhits, (uint32_t)__popc(m));      // one atomic per warp
                slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
                if (hit) queue[slot] = h;
            }
        }
        __syncthreads();
        // ---- phase 1.5: ticket table.  Hit h owns the grab tickets [gstart[h], gstart[h+1]): one ticket per
        // kVoteGrab of its entries (exclusive block scan of ceil(len / kVoteGrab) over the queue).
        const uint32_t nhits = s_nhits;
        {
            constexpr int ITEMS = kHitQueue / THREADS;
            uint32_t excl[ITEMS], sum = 0;
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                const uint32_t idx = tid * ITEMS + k;
                excl[k] = sum;
                sum += idx < nhits ? (queue[idx].y + kVoteGrab - 1) / kVoteGrab : 0u;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_red[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t x = lane < THREADS / 32 ? s_red[lane] : 0u;
                uint32_t inc2 = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) inc2 += t;
                }
                s_red[lane] = inc2 - x;
                if (lane == 31) s_total = inc2;
            }
            __syncthreads();
            const uint32_t base_t = s_red[warp] + incl - sum;
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                const uint32_t idx = tid * ITEMS + k;
                if (idx < nhits) gstart[idx] = base_t + excl[k];
            }
            __syncthreads();
        }
        // ---- phase 2: warps draw grab tickets; a ticket is one kVoteGrab-entry piece of one hit, so a
        // 100k-entry bucket is shared by every warp, and no warp ever polls an exhausted hit.
        const uint32_t total = s_total;
        while (true) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&s_ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) break;
            // hit of ticket t = last h with gstart[h] <= t: two-level 32-wide search over <= 2048 hits
            uint32_t hi, off;
            {
                const uint32_t i1 = (uint32_t)lane * (kHitQueue / 32);
                const unsigned m1 = __ballot_sync(0xffffffffu, i1 < nhits && gstart[i1] <= t);
                const uint32_t blk = (31u - (uint32_t)__clz((int)m1)) * (kHitQueue / 32);
                uint32_t cnt = 0, g_found = 0;
#pragma unroll
                for (int r = 0; r < kHitQueue / 1024; r++) {
                    const uint32_t i2 = blk + r * 32 + lane;
                    const uint32_t g = i2 < nhits ? gstart[i2] : 0xFFFFFFFFu;
                    const unsigned m2 = __ballot_sync(0xffffffffu, g <= t);
                    if (m2) {                                                    // the last lane with g <= t holds gstart[hi]
                        g_found = __shfl_sync(0xffffffffu, g, 31 - __clz((int)m2));
                        cnt += __popc(m2);
                    }
                }
                hi = blk + cnt - 1;
                off = (t - g_found) * kVoteGrab;
            }
            const uint4 h = queue[hi];
            const uint32_t ngrab = min((uint32_t)kVoteGrab, h.y - off);
            if (lane == 0) my_votes += ngrab;
            vote_single_hit(ctx, FS, a.entries, h.z, h.w, h.x + off, ngrab, lane, my_exact);
        }
        __syncthreads();
    }

    // ---- phase 3: block max, statistics, emission of candidate cells
    uint32_t lmax = 0, nz = 0;
    for (int i = tid; i < kNAlphaBins * S; i += THREADS) {
        uint32_t c = acc[i];
        lmax = max(lmax, c);
        nz += (c != 0);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
        my_exact += __shfl_xor_sync(0xffffffffu, my_exact, o);
    }
    if (lane == 0) {
        s_red[warp] = lmax;
        if (nz) atomicAdd(&a.totals[1], (unsigned long long)nz);
        if (my_votes) atomicAdd(&s_votes, my_votes);
        if (my_exact) atomicAdd(&s_exact, my_exact);
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t v = (lane < THREADS / 32) ? s_red[lane] : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) {
            uint32_t old = atomicMax(&a.scalars[1], v);
            s_red[0] = max(old, v);
            if (s_votes) atomicAdd(&a.totals[0], s_votes);
            if (s_exact) atomicAdd(&a.scalars[3], s_exact);
        }
    }
    __syncthreads();
    const uint32_t bound = s_red[0];                       // <= final global max
    if (bound == 0) return;
    const float min_votecount = a.emit_all ? 0.0f : a.thr * (float)bound;     // model.cu:164
    for (int i = tid; i < kNAlphaBins * S; i += THREADS) {
        uint32_t c = acc[i];
        if (c != 0 && (float)c > min_votecount) {
            uint32_t bin = i / S, loc = i - bin * S;
            uint32_t slot = atomicAdd(&a.scalars[0], 1u);
            if (slot < a.cand_cap) {
                // [scene ref : 32 | model point : 26 | alpha : 6]   (kernel.cu:548-549, model.h:61-63)
                a.cand_codes[slot] = ((unsigned long long)(uint32_t)s_r << 32) |
                                     (unsigned long long)((((uint32_t)chunk_base + loc) << 6) | bin);
                a.cand_counts[slot] = c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
__global__ void filter_kernel(const unsigned long long *codes, const uint32_t *counts, uint32_t n, float thr,
                              uint32_t gmax, int emit_all, unsigned long long *ocodes, uint32_t *ocounts,
                              uint32_t *on) {
    float min_votecount = emit_all ? 0.0f : thr * (float)gmax;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t c = counts[i];
        if ((float)c > min_votecount) {
            uint32_t s = atomicAdd(on, 1u);
            ocodes[s] = codes[i];
            ocounts[s] = c;
        }
    }
}

// The same with the candidate count and the global maximum read from device memory (scalars[0], scalars[1]: the
// maximum may just have been all-reduced in place) and the survivor count written to scalars[2]: no host round trip
// between the vote kernel, the exchange of the maximum and the filter.
__global__ void filter_dev_kernel(const unsigned long long *codes, const uint32_t *counts, uint32_t cap, float thr,
                                  uint32_t *scalars, unsigned long long *ocodes, uint32_t *ocounts) {
    const uint32_t n = min(scalars[0], cap);
    const float min_votecount = thr * (float)scalars[1];                       // model.cu:164
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = counts[i];
        if ((float)c > min_votecount) {
            const uint32_t s = atomicAdd(&scalars[2], 1u);
            ocodes[s] = codes[i];
            ocounts[s] = c;
        }
    }
}
// survivor records travel between the ranks as 12-byte triples (code lo, code hi, count)
__global__ void pack_survivors_kernel(const unsigned long long *codes, const uint32_t *counts, uint32_t n, uint32_t *rec) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long c = codes[i];
        rec[3 * i] = (uint32_t)c; rec[3 * i + 1] = (uint32_t)(c >> 32); rec[3 * i + 2] = counts[i];
    }
}
__global__ void unpack_survivors_kernel(const uint32_t *rec, uint32_t n, unsigned long long *codes, uint32_t *counts) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        codes[i] = (unsigned long long)rec[3 * i] | ((unsigned long long)rec[3 * i + 1] << 32);
        counts[i] = rec[3 * i + 2];
    }
}

int Workspace::reserve(size_t bytes) {
    bytes += 16 * 256;                                     // alignment slack for up to 16 slices
    used = 0;
    if (bytes <= cap) return PPF_OK;
    if (base) pool_free(base, cap);
    base = nullptr; cap = 0;
    size_t want = bytes + bytes / 2, got = 0;
    base = (char *)pool_alloc(want, &got);
    if (!base) { set_last_error("workspace: out of device memory"); return PPF_ERR_CUDA; }
    cap = got;
    return PPF_OK;
}
void *Workspace::take_bytes(size_t bytes) {
    size_t off = (used + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    used = off + bytes;
    return base + off;
}
void Workspace::release() { if (base) pool_free(base, cap); base = nullptr; cap = used = 0; }

static int ensure(void **p, size_t bytes) {
    if (*p) pooled_free(*p);
    *p = nullptr;
    return pooled_malloc(p, bytes ? bytes : 16) == cudaSuccess ? PPF_OK : PPF_ERR_CUDA;
}

void vote_result_free(VoteResult &r) {
    pooled_free(r.cand_codes); pooled_free(r.cand_counts); pooled_free(r.scalars); pooled_free(r.votes_total); pooled_free(r.sched); pooled_free(r.acc_scratch);
    pooled_free(r.replay);
    pooled_free(r.codes); pooled_free(r.counts); pooled_free(r.transformations); pooled_free(r.weighted);
    pooled_free(r.trans); pooled_free(r.rots); pooled_free(r.scores);
    r.ws.release(); r.ws2.release();
    r = VoteResult();
}

int vote_reserve_K(VoteResult &r, size_t K) {
    if (K <= r.cap_K && r.codes) return PPF_OK;
    size_t cap = std::max<size_t>(K, 1024);
    if (ensure((void **)&r.codes, cap * 8) || ensure((void **)&r.counts, cap * 4) ||
        ensure((void **)&r.transformations, cap * 64) || ensure((void **)&r.weighted, cap * 4) ||
        ensure((void **)&r.trans, cap * sizeof(float3)) || ensure((void **)&r.rots, cap * sizeof(float4)) ||
        ensure((void **)&r.scores, cap * 4)) {
        set_last_error("lookup: out of device memory for survivors");
        return PPF_ERR_CUDA;
    }
    r.cap_K = cap;
    return PPF_OK;
}

// Runs the fused vote kernel for the reference points of one shard.
int vote_run(const ModelTable &m, const Cloud &scene, unsigned df, int shard_rank, int shard_count, int emit_all,
             VoteResult &r, unsigned long long *pairs_out, int *launches) {
    if (df == 0 || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) {
        set_last_error("vote: ref_point_downsample_factor == 0 or bad shard");
        return PPF_ERR_INVALID;
    }
    if (!r.scalars) {
        PPF_CUDA_TRY(pooled_malloc(&r.scalars, 4 * sizeof(uint32_t)));
        PPF_CUDA_TRY(pooled_malloc(&r.votes_total, 2 * sizeof(unsigned long long)));
    }
    if (!r.cand_codes) {
        r.cand_cap = emit_all ? (size_t)1 << 24 : (size_t)1 << 22;
        if (const char *e = getenv("PPF_B200_CAND_CAP")) r.cand_cap = (size_t)std::max(1, atoi(e));   // test hook: force the overflow path
        PPF_CUDA_TRY(pooled_malloc(&r.cand_codes, r.cand_cap * 8));
        PPF_CUDA_TRY(pooled_malloc(&r.cand_counts, r.cand_cap * 4));
    }
    r.K = 0;
    const int ns = scene.n;
    if (ns > 0 && !scene.inv) { set_last_error("vote: the scene cloud has no spatial index"); return PPF_ERR_INVALID; }
    // reference points: every df-th scene point (kernel.cu:432), then every shard_count-th of those
    const int R_all = ns > 1 ? (ns + (int)df - 1) / (int)df : 0;     // ppf_kernel is a no-op for count <= 1
    const int R = R_all > shard_rank ? (R_all - shard_rank + shard_count - 1) / shard_count : 0;
    if (pairs_out) *pairs_out = (unsigned long long)R * (unsigned long long)ns;
    if (launches) *launches = 0;
    // kernel choice: made with the chunk geometry at model build time (ppf_model.cu)
    const bool use_grouped = m.prefer_grouped && vote_grouped_supported(m, ns);
    // Candidate overflow (more cells above the running threshold than the buffer holds): the buffer grows to the
    // number the first pass COUNTED and the second pass starts from the first pass's maximum, so that it emits
    // exactly the cells above the final threshold -- never more than the first pass counted: two passes at most.
    uint32_t seed_max = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        PPF_CUDA_TRY(cudaMemsetAsync(r.scalars, 0, 4 * sizeof(uint32_t), cur_stream()));
        if (seed_max && !emit_all)
            PPF_CUDA_TRY(cudaMemcpyAsync(r.scalars + 1, &seed_max, sizeof(uint32_t), cudaMemcpyHostToDevice, cur_stream()));
        PPF_CUDA_TRY(cudaMemsetAsync(r.votes_total, 0, 2 * sizeof(unsigned long long), cur_stream()));
        r.cand_n = 0; r.local_max = 0;
        if (R == 0 || m.cloud.n <= 1 || m.K_d == 0) return PPF_OK;
        VoteArgs a;
        a.spos = scene.pos; a.snrm = scene.nrm; a.sfy = scene.fy; a.sfz = scene.fz; a.ns = ns;
        a.sinv = scene.inv; a.gbox_lo = scene.gbox_lo; a.gbox_hi = scene.gbox_hi; a.tbox_lo = scene.tbox_lo; a.tbox_hi = scene.tbox_hi;
        {   // far cells: scene features beyond the model's distance range whose FNV key equals a model key still vote
            // in the reference (kernel.cu:480-501); usually there are none and the cull radius stays at the table's edge
            int rc_far = model_far_cells(m, scene, &a.far_cells, &a.far_buckets, &a.n_far, &a.far_kd_min, &a.far_kd_max);
            if (rc_far) return rc_far;
            // a pair at true distance >= (K + 1) d_dist (1 + 1e-4) has distance bin >= K even after the approximate
            // sqrt (relative error ~1e-6): with K = one past the last bin that can hit, it cannot vote.
            // Conservative: everything nearer is processed.
            const int K_hit = a.n_far ? std::max(m.K_d, a.far_kd_max + 1) : m.K_d;
            const float r = (float)(K_hit + 1) * m.d_dist * 1.0001f;
            a.cull_r2 = r * r;
        }
        a.ref_start = shard_rank * (int)df; a.ref_stride = shard_count * (int)df; a.ref_count = R;
        a.mpos = m.cloud.pos; a.mfy = m.cloud.fy; a.mfz = m.cloud.fz; a.nm = m.cloud.n;
        a.d_dist = m.d_dist; a.inv_d = m.inv_d_dist; a.K_d = m.K_d; a.U = m.U;
        a.cell2bucket = m.cell2bucket; a.ranges = m.ranges; a.entries = m.entries; a.map = m.map;
        a.n_chunks = m.n_chunks; a.chunk_rows = m.chunk_rows;
        // average entries per (bucket, chunk): >= 512 -> long slices.  PPF_B200_REST_E=4|8 forces it (A/B hook)
        a.rest_long = (double)m.cloud.n * m.cloud.n / std::max(1u, m.U) / std::max(1, m.n_chunks) >= 512.0;
        if (const char *e = getenv("PPF_B200_REST_E")) a.rest_long = atoi(e) >= 8;
        a.queue_cap = 0; a.sched = nullptr; a.acc_scratch = nullptr; a.opaque_zero = 0;
        a.replay = nullptr; a.replay_cap = 0;
        a.thr = m.vote_count_threshold; a.emit_all = emit_all;
        a.cand_codes = r.cand_codes; a.cand_counts = r.cand_counts; a.cand_cap = (uint32_t)r.cand_cap;
        a.scalars = r.scalars; a.totals = r.votes_total;
        const size_t smem = (size_t)kNAlphaBins * acc_stride(m.chunk_rows) * 4 + (size_t)kHitQueue * (sizeof(uint4) + 4);
        const long long grid = (long long)R * m.n_chunks;
        if (grid > 0x7FFFFFFFLL) { set_last_error("vote: too many (reference point, chunk) CTAs"); return PPF_ERR_UNSUPPORTED; }
        if (use_grouped) {
            // [0] next reference point, [1, R] next chunk of each, [R + 1] dense reference points registered,
            // [R + 2, 2R + 1] their list, [2R + 2] next of them to process
            const size_t sched_words = 2 * (size_t)R + 3;
            if (r.sched_cap < sched_words) {
                pooled_free(r.sched); r.sched = nullptr; r.sched_cap = 0;
                PPF_CUDA_TRY(pooled_malloc(&r.sched, sched_words * sizeof(uint32_t)));
                r.sched_cap = sched_words;
            }
            PPF_CUDA_TRY(cudaMemsetAsync(r.sched, 0, sched_words * sizeof(uint32_t), cur_stream()));
            a.sched = r.sched;
            const size_t words = vote_grouped_scratch_words(m);
            if (r.acc_scratch_cap < words) {
                pooled_free(r.acc_scratch); r.acc_scratch = nullptr; r.acc_scratch_cap = 0;
                PPF_CUDA_TRY(pooled_malloc(&r.acc_scratch, words * sizeof(uint32_t)));
                r.acc_scratch_cap = words;
            }
            a.acc_scratch = r.acc_scratch;
            const size_t rq_words = (size_t)vote_grouped_ctas() * vote_grouped_replay_cap();
            if (r.replay_cap < rq_words) {
                pooled_free(r.replay); r.replay = nullptr; r.replay_cap = 0;
                PPF_CUDA_TRY(pooled_malloc(&r.replay, rq_words * sizeof(uint2)));
                r.replay_cap = rq_words;
            }
            a.replay = getenv("PPF_B200_INLINE_EXACT") ? nullptr : r.replay;
            a.replay_cap = (uint32_t)vote_grouped_replay_cap();
            int rc = vote_grouped_launch(a, R);
            if (rc) return rc;
        } else if (smem > 113 * 1024) {
            PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            vote_kernel<1024><<<(unsigned)grid, 1024, smem, cur_stream()>>>(a);
            count_launch();
        } else {
            PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            vote_kernel<512><<<(unsigned)grid, 512, smem, cur_stream()>>>(a);
            count_launch();
        }
        PPF_CUDA_TRY(cudaGetLastError());
        if (launches) (*launches)++;
        uint32_t h[4];
        PPF_CUDA_TRY(memcpy_sync(h, r.scalars, sizeof(h), cudaMemcpyDeviceToHost));
        r.cand_n = h[0]; r.local_max = h[1];
        if (h[0] <= r.cand_cap) return PPF_OK;
        // candidate buffer too small: grow and vote again (results are deterministic)
        seed_max = h[1];
        r.cand_cap = (size_t)h[0] + 1024;
        pooled_free(r.cand_codes); pooled_free(r.cand_counts);
        r.cand_codes = nullptr; r.cand_counts = nullptr;
        PPF_CUDA_TRY(pooled_malloc(&r.cand_codes, r.cand_cap * 8));
        PPF_CUDA_TRY(pooled_malloc(&r.cand_counts, r.cand_cap * 4));
    }
    set_last_error("vote: candidate buffer kept overflowing");
    return PPF_ERR_CUDA;
}

// Threshold against the global maximum and order (count desc, code asc): the order that
// thrust::sort + histogram + sort_by_key(greater<float>) leaves (model.cu:148-158; the
// comparator sort is a stable merge sort, so equal counts stay in ascending code order).
// Workspace bytes order_survivors needs for K survivors (sort buffers + CUB temp storage).
size_t order_survivors_bytes(size_t K) {
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, K);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb2, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                              (unsigned long long *)nullptr, (unsigned long long *)nullptr, K);
    return K * 12 + std::max(tb, tb2) + 1024;
}

// codes_in / counts_in may live in r.ws; the scratch comes from r.ws2 (reserved here).
int order_survivors(VoteResult &r, size_t K, unsigned long long *codes_in, uint32_t *counts_in) {
    if (K == 0) { r.K = 0; return PPF_OK; }
    int rc = vote_reserve_K(r, K);
    if (rc) return rc;
    if ((rc = r.ws2.reserve(order_survivors_bytes(K)))) return rc;
    size_t tb = 0, tb2 = 0;
    unsigned long long *c1 = r.ws2.take<unsigned long long>(K);
    uint32_t *n1 = r.ws2.take<uint32_t>(K);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, codes_in, c1, counts_in, n1, K);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb2, n1, r.counts, c1, r.codes, K);
    void *tmp = r.ws2.take_bytes(std::max(tb, tb2));
    if (!c1 || !n1 || !tmp) { set_last_error("order_survivors: workspace too small"); return PPF_ERR_CUDA; }
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, codes_in, c1, counts_in, n1, K, 0, 64, cur_stream()));
    PPF_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(tmp, tb2, n1, r.counts, c1, r.codes, K, 0, 32, cur_stream()));
    r.K = K;
    return PPF_OK;
}

// Threshold + ordering of one rank's candidates, or of all ranks' when `comm` couples several: all_reduce(MAX) of the
// vote maximum, local filter against the GLOBAL threshold, all_gather of the survivor records, ordering of the
// merged list (identical on every rank).  Every rank makes the same collective calls whatever it has to contribute.
int vote_finalize_dist(const ModelTable &m, Comm *comm, VoteResult &r, uint32_t *global_max_out) {
    const int world = comm ? comm->world : 1;
    r.K = 0;
    if (!r.scalars) { set_last_error("finalize: no vote has run on this lookup"); return PPF_ERR_INVALID; }
    if (world > 1) { int rc = comm->allreduce_max_u32(r.scalars + 1, 1); if (rc) return rc; }
    const uint32_t n = (uint32_t)std::min<size_t>(r.cand_n, r.cand_cap);
    int rc = r.ws.reserve((size_t)n * 12 + 512 + (size_t)world * 8);
    if (rc) return rc;
    unsigned long long *fc = r.ws.take<unsigned long long>(std::max<uint32_t>(n, 1));
    uint32_t *fn = r.ws.take<uint32_t>(std::max<uint32_t>(n, 1));
    uint32_t *d_counts = r.ws.take<uint32_t>(world);
    if (!fc || !fn || !d_counts) { set_last_error("finalize: workspace too small"); return PPF_ERR_CUDA; }
    PPF_CUDA_TRY(cudaMemsetAsync(r.scalars + 2, 0, 4, cur_stream()));
    if (n) {
        filter_dev_kernel<<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(
            r.cand_codes, r.cand_counts, (uint32_t)r.cand_cap, m.vote_count_threshold, r.scalars, fc, fn);
        count_launch();
        PPF_CUDA_TRY(cudaGetLastError());
    }
    uint32_t h[4];
    std::vector<uint32_t> counts(world, 0);
    if (world > 1) {
        rc = comm->allgather_u32(r.scalars + 2, d_counts, 1);
        if (rc) return rc;
        PPF_CUDA_TRY(cudaMemcpyAsync(counts.data(), d_counts, (size_t)world * 4, cudaMemcpyDeviceToHost, cur_stream()));
    }
    PPF_CUDA_TRY(memcpy_sync(h, r.scalars, sizeof(h), cudaMemcpyDeviceToHost));
    if (global_max_out) *global_max_out = h[1];
    const size_t K_local = h[2];
    if (world == 1) return (h[1] == 0 || K_local == 0) ? PPF_OK : order_survivors(r, K_local, fc, fn);
    std::vector<size_t> off(world), bytes(world);
    size_t K = 0;
    for (int i = 0; i < world; i++) { off[i] = K * 12; bytes[i] = (size_t)counts[i] * 12; K += counts[i]; }
    // merged list: packed records in ws2; the unpacked (codes, counts) go to ws, re-reserved once fc / fn are packed
    if ((rc = r.ws2.reserve((K + K_local) * 12 + 1024))) return rc;
    uint32_t *send = r.ws2.take<uint32_t>(std::max<size_t>(K_local, 1) * 3), *recv = r.ws2.take<uint32_t>(std::max<size_t>(K, 1) * 3);
    if (!send || !recv) { set_last_error("finalize: merge workspace too small"); return PPF_ERR_CUDA; }
    if (K_local) {
        pack_survivors_kernel<<<(int)std::min<size_t>((K_local + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(fc, fn, (uint32_t)K_local, send);
        count_launch();
    }
    if ((rc = comm->allgatherv(send, recv, off.data(), bytes.data()))) return rc;
    if (K == 0) return PPF_OK;
    PPF_CUDA_TRY(cudaStreamSynchronize(cur_stream()));            // the pack kernel reads fc / fn until here
    if ((rc = r.ws.reserve(K * 12 + 512))) return rc;
    unsigned long long *mc = r.ws.take<unsigned long long>(K);
    uint32_t *mn = r.ws.take<uint32_t>(K);
    if (!mc || !mn) { set_last_error("finalize: merge workspace too small"); return PPF_ERR_CUDA; }
    unpack_survivors_kernel<<<(int)std::min<size_t>((K + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(recv, (uint32_t)K, mc, mn);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    PPF_CUDA_TRY(cudaStreamSynchronize(cur_stream()));            // order_survivors re-reserves ws2 (recv lives there)
    return order_survivors(r, K, mc, mn);
}

int vote_finalize(const ModelTable &m, uint32_t global_max, int emit_all, VoteResult &r) {
    uint32_t h[4];
    PPF_CUDA_TRY(memcpy_sync(h, r.scalars, sizeof(h), cudaMemcpyDeviceToHost));
    uint32_t n = h[0];
    r.K = 0;
    if (n == 0 || global_max == 0) return PPF_OK;
    int rc = r.ws.reserve((size_t)n * 12 + 256 + order_survivors_bytes(n));
    if (rc) return rc;
    unsigned long long *fc = r.ws.take<unsigned long long>(n);
    uint32_t *fn = r.ws.take<uint32_t>(n);
    uint32_t *d_on = r.ws.take<uint32_t>(1);
    PPF_CUDA_TRY(cudaMemsetAsync(d_on, 0, 4, cur_stream()));
    filter_kernel<<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, cur_stream()>>>(r.cand_codes, r.cand_counts, n,
                                                                         m.vote_count_threshold, global_max,
                                                                         emit_all, fc, fn, d_on);
    count_launch();
    PPF_CUDA_TRY(cudaGetLastError());
    uint32_t K = 0;
    PPF_CUDA_TRY(memcpy_sync(&K, d_on, 4, cudaMemcpyDeviceToHost));
    return order_survivors(r, K, fc, fn);
}

}  // namespace ppf
