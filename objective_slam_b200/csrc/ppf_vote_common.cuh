// ppf_vote_common.cuh -- pieces shared by the two vote kernels (ppf_vote.cu: one hit per warp pass;
// ppf_vote_grouped.cu: hits of the same bucket grouped so that one ATOMS serves one accumulator row).
#pragma once
#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

namespace ppf {

struct VoteArgs {
    // scene
    const float4 *spos, *snrm, *sfy, *sfz;           // stored (Morton) order
    const uint32_t *sinv;                            // caller's index -> stored position
    const float4 *gbox_lo, *gbox_hi, *tbox_lo, *tbox_hi;
    float cull_r2;                                   // squared distance beyond which no scene pair can hit the table
    int ns;
    int ref_start, ref_stride, ref_count;         // s_r = ref_start + k*ref_stride
    // model
    const float4 *mpos, *mfy, *mfz;
    int nm;
    float d_dist, inv_d;
    int K_d;
    uint32_t U;
    const uint32_t *cell2bucket;
    // cells beyond the model's distance range whose FNV key collides with a model key (FarCells; usually none)
    const unsigned long long *far_cells;
    const uint32_t *far_buckets;
    int n_far, far_kd_min, far_kd_max;
    const uint2 *ranges;
    const uint32_t *entries, *map;
    int n_chunks, chunk_rows;
    // grouped kernel only: hit queue capacity (records); work counters: sched[0] = next reference point,
    // sched[1 + r] = next chunk of reference point r (zeroed before the launch)
    int queue_cap;
    int rest_long;                                // grouped kernel: bucket slices are long (8 entries per lane in vote_rest, else 4)
    uint32_t opaque_zero;                         // always 0; a third add operand the compilers cannot fold (see vote_grouped)
    uint32_t *sched;
    uint32_t *acc_scratch;                        // [CTAs][n_chunks][31 x S]: accumulators parked between scene segments
    uint2 *replay;                                // [CTAs][replay_cap]: deferred exact votes (nullptr: inline)
    uint32_t replay_cap;
    // output
    float thr;
    int emit_all;                                 // 1: emit every non-zero cell (vote histogram)
    unsigned long long *cand_codes;
    uint32_t *cand_counts;
    uint32_t cand_cap;
    uint32_t *scalars;                            // [0]=cand_n [1]=max [2]=(unused) [3]=exact-alpha votes
    unsigned long long *totals;                   // [0]=votes cast [1]=non-zero cells
};

__device__ __forceinline__ FrameYZ load_frame(const float4 *__restrict__ fy, const float4 *__restrict__ fz, int i) {
    float4 y = __ldg(fy + i), z = __ldg(fz + i);
    FrameYZ f;
    f.y[0] = y.x; f.y[1] = y.y; f.y[2] = y.z; f.y[3] = y.w;
    f.z[0] = z.x; f.z[1] = z.y; f.z[2] = z.z; f.z[3] = z.w;
    return f;
}

// Probe: bucket of a quantised scene feature, or kNoBucket.  Replaces hash + lower_bound + key equality
// (ParallelHashArray::GetIndices, ppf_vote_count_kernel kernel.cu:489-497): cells inside the model's distance
// range through the cell table, cells beyond it (reachable only through a 32-bit key collision) through the short
// sorted far-cell list.
__device__ __forceinline__ uint32_t probe_bucket(const VoteArgs &a, const FeatureBins &fb) {
    if (fb.kd < 0) return kNoBucket;                               // NaN / Inf distance: key 0 never matches
    if (fb.kd < a.K_d) return __ldg(a.cell2bucket + cell_index(fb.kd, fb.k1, fb.k2, fb.k3));
    if (fb.kd < a.far_kd_min || fb.kd > a.far_kd_max) return kNoBucket;
    const unsigned long long id = (unsigned long long)fb.kd * kCellsPerDist +
                                  (unsigned long long)((fb.k1 * kAngleCells + fb.k2) * kAngleCells + fb.k3);
    int lo = 0, hi = a.n_far;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.far_cells + mid) < id) lo = mid + 1; else hi = mid;
    }
    return (lo < a.n_far && __ldg(a.far_cells + lo) == id) ? __ldg(a.far_buckets + lo) : kNoBucket;
}

// Voting context shared by the fast and the exact path.
struct VoteCtx {
    const uint32_t *map;
    const float4 *mfy, *mfz, *mpos, *spos;
    int nm, chunk_base, stride;
    uint32_t *acc;
    uint32_t acc_addr;                            // acc as a shared-window address (grouped kernel)
    uint32_t opaque_zero;
    // deferred exact votes (grouped kernel): records (table position, scene point) parked in global memory and
    // replayed by the whole CTA at the end of the chunk; rq == nullptr: the exact path runs inline
    uint2 *rq;
    uint32_t rq_cap;
    uint32_t *rq_count;                           // shared-memory counter
};

__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
    // no memory clobber: loads of staged entries may be scheduled ahead of the vote
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr));
}
// +1 on accumulator cell idx.  SH = true: explicit shared-space RED (the grouped kernel's VoteCtx travels
// through enough code that the compiler no longer proves acc is shared and would emit generic ATOM).
template <bool SH>
__device__ __forceinline__ void acc_inc(const VoteCtx &c, uint32_t idx) {
    if constexpr (SH) red_shared_inc(c.acc_addr + idx * 4u);
    else atomicAdd(&c.acc[idx], 1u);
}

// Exact alpha bin of one vote: rebuild u and v the way trans_model_scene does (kernel.cu:330-342).
// Needed by ~1e-4 of the votes (guard band around the 30 bin edges, degenerate u or v); kept out of
// line so that the hot loop stays small and free of divergence.
static __device__ __noinline__ uint32_t exact_vote_index(const VoteCtx &c, const FrameYZ &FS, uint32_t s_i, uint32_t entry,
                                                  uint32_t pos) {
    const uint32_t loc = entry & kLocMask;
    const uint32_t pidx = __ldg(c.map + pos);
    const int m_r = c.chunk_base + (int)loc;
    const int m_i = (int)(pidx - (uint32_t)m_r * (uint32_t)c.nm);
    const FrameYZ FM = load_frame(c.mfy, c.mfz, m_r);
    const float4 mi = __ldg(c.mpos + m_i);
    const float4 si = __ldg(c.spos + s_i);
    float uy, uz, vy, vz;
    frame_apply_yz(FM, mi.x, mi.y, mi.z, uy, uz);
    frame_apply_yz(FS, si.x, si.y, si.z, vy, vz);
    return alpha_bin_exact(uy, uz, vy, vz) * (uint32_t)c.stride + loc;
}

// Same from the table position alone (the row comes from the pair the map names): replay of a deferred exact vote.
static __device__ __forceinline__ uint32_t exact_vote_index_at(const VoteCtx &c, const FrameYZ &FS, uint32_t s_i, uint32_t pos) {
    const uint32_t pidx = __ldg(c.map + pos);
    const int m_r = (int)(pidx / (uint32_t)c.nm);
    const int m_i = (int)(pidx - (uint32_t)m_r * (uint32_t)c.nm);
    const FrameYZ FM = load_frame(c.mfy, c.mfz, m_r);
    const float4 mi = __ldg(c.mpos + m_i);
    const float4 si = __ldg(c.spos + s_i);
    float uy, uz, vy, vz;
    frame_apply_yz(FM, mi.x, mi.y, mi.z, uy, uz);
    frame_apply_yz(FS, si.x, si.y, si.z, vy, vz);
    return alpha_bin_exact(uy, uz, vy, vz) * (uint32_t)c.stride + (uint32_t)(m_r - c.chunk_base);
}

// +1 on the EXACT cell of one vote.  The exact path is a chain of dependent global loads (map -> frames -> points)
// run by one lane while the other 31 wait: 12% of the grouped kernel's stall samples for 1.35e-4 of the votes.  With
// a replay queue the lane only parks (position, scene point); the CTA replays the records densely (32 lanes busy,
// latencies overlapped) before the chunk's accumulator is read.  Queue full -> inline, as before.
__device__ __forceinline__ void exact_vote(const VoteCtx &c, const FrameYZ &FS, uint32_t s_i, uint32_t entry, uint32_t pos) {
    if (c.rq) {
        const uint32_t slot = atomicAdd(c.rq_count, 1u);
        if (slot < c.rq_cap) { c.rq[slot] = make_uint2(pos, s_i); return; }
    }
    atomicAdd(&c.acc[exact_vote_index(c, FS, s_i, entry, pos)], 1u);
}

// The hot loop votes OPTIMISTICALLY: every entry of a batch adds 1 to the cell its fast alpha bin
// names (always a valid cell), the guard-band margins are min-reduced and the slow flags OR-ed across
// the batch (one VIADDMNMX and half a LOP3 per vote), and only when a batch contains a vote whose fast
// bin is not provably the reference's (about one batch in 30) is it re-examined: such a vote is moved
// from the optimistic cell to the exact one (-1 / +1 by the same thread, so no other thread can
// observe a negative count, and phase 3 only reads after a barrier).  This removes the select, the
// mask bookkeeping and all branches from the per-vote path.
__device__ __forceinline__ void repair_vote(const VoteCtx &c, const FrameYZ &FS, uint32_t hit_ones, uint32_t s_i,
                                            uint32_t entry, uint32_t pos, uint32_t &n_exact) {
    uint32_t bin;
    if (alpha_bin_margin(hit_ones, entry, bin) >= kGuardSpan || (entry & kSlowBit)) {
        atomicSub(&c.acc[bin * (uint32_t)c.stride + (entry & kLocMask)], 1u);
        exact_vote(c, FS, s_i, entry, pos);
        n_exact++;
    }
}

// One full batch: E entries per lane, entry u of this lane sits at table position pos_lane + 32 u.
template <int E, bool SH = false>
__device__ __forceinline__ void vote_batch(const VoteCtx &c, const FrameYZ &FS, uint32_t hit_ones, uint32_t s_i,
                                           const uint32_t (&e)[E], uint32_t pos_lane, uint32_t &n_exact) {
    uint32_t worst = 0, flags = 0;
#pragma unroll
    for (int u = 0; u < E; u++) {
        uint32_t bin;
        worst = max(worst, alpha_bin_margin(hit_ones, e[u], bin));
        flags |= e[u];
        acc_inc<SH>(c, bin * (uint32_t)c.stride + (e[u] & kLocMask));
    }
    if (worst >= kGuardSpan || (flags & kSlowBit)) {
#pragma unroll
        for (int u = 0; u < E; u++) repair_vote(c, FS, hit_ones, s_i, e[u], pos_lane + 32 * u, n_exact);
    }
}

// All votes of ONE hit against `ngrab` consecutive table entries starting at table position pos_grab
// (one scheduler grab).  hit_word = [(theta_v + half) : 20 | slow : 1 | 0 : 11].  32 lanes read 32
// consecutive entries per load; full 256-entry batches are software-pipelined in registers.
template <bool SH = false>
__device__ __forceinline__ void vote_single_hit(const VoteCtx &ctx, const FrameYZ &FS,
                                                const uint32_t *__restrict__ entries, uint32_t hit_word, uint32_t s_i,
                                                uint32_t pos_grab, uint32_t ngrab, int lane, uint32_t &my_exact) {
    const uint32_t hit_theta = hit_word | kLowOnes;               // low 12 bits set: see alpha_bin_fast
    const uint32_t *__restrict__ ent = entries + pos_grab;
    uint32_t *acc = ctx.acc;
    const int S = ctx.stride;
    if (hit_word & kSlowBit) {
        // degenerate scene pair (v ~ 0): every vote of this hit takes the exact path
        for (uint32_t j = lane; j < ngrab; j += 32) {
            atomicAdd(&acc[exact_vote_index(ctx, FS, s_i, __ldg(ent + j), pos_grab + j)], 1u);
            my_exact++;
        }
        return;
    }
    // Full batches, software-pipelined: the 8 loads of batch b+1 are in flight while batch b votes.
    constexpr int E = kVoteBatch / 32;
    const uint32_t nfull = ngrab / kVoteBatch;
    uint32_t e0[E], e1[E];
    if (nfull) {
#pragma unroll
        for (int u = 0; u < E; u++) e0[u] = __ldg(ent + u * 32 + lane);
    }
    for (uint32_t bi = 0; bi < nfull; bi += 2) {
        if (bi + 1 < nfull) {
#pragma unroll
            for (int u = 0; u < E; u++) e1[u] = __ldg(ent + (bi + 1) * kVoteBatch + u * 32 + lane);
        }
        vote_batch<E, SH>(ctx, FS, hit_theta, s_i, e0, pos_grab + bi * kVoteBatch + lane, my_exact);
        if (bi + 1 < nfull) {
            if (bi + 2 < nfull) {
#pragma unroll
                for (int u = 0; u < E; u++) e0[u] = __ldg(ent + (bi + 2) * kVoteBatch + u * 32 + lane);
            }
            vote_batch<E, SH>(ctx, FS, hit_theta, s_i, e1, pos_grab + (bi + 1) * kVoteBatch + lane, my_exact);
        }
    }
    // tail of the grab: fewer than kVoteBatch entries, all loaded before the first vote (one L2 latency, not two)
    const uint32_t done = nfull * kVoteBatch;
    if (done < ngrab) {
        const uint32_t n = ngrab - done;
        uint32_t worst = 0, flags = 0;
#pragma unroll
        for (int u = 0; u < E; u++) {
            const uint32_t j = u * 32 + lane;
            e0[u] = (j < n) ? __ldg(ent + done + j) : 0u;
        }
#pragma unroll
        for (int u = 0; u < E; u++) {
            const uint32_t j = u * 32 + lane;
            if (j < n) {
                uint32_t bin;
                worst = max(worst, alpha_bin_margin(hit_theta, e0[u], bin));
                flags |= e0[u];
                acc_inc<SH>(ctx, bin * (uint32_t)S + (e0[u] & kLocMask));
            }
        }
        if (worst >= kGuardSpan || (flags & kSlowBit)) {
#pragma unroll
            for (int u = 0; u < E; u++) {
                const uint32_t j = u * 32 + lane;
                if (j < n) repair_vote(ctx, FS, hit_theta, s_i, e0[u], pos_grab + done + j, my_exact);
            }
        }
    }
}

// Squared distance from the reference point to an axis-aligned box (0 inside).  NaN boxes compare false
// against the cull radius, i.e. are never culled.
__device__ __forceinline__ float box_dist2(const PointN &R, float4 lo, float4 hi) {
    float dx = fmaxf(fmaxf(lo.x - R.x, R.x - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - R.y, R.y - hi.y), 0.f),
          dz = fmaxf(fmaxf(lo.z - R.z, R.z - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}


// grouped kernel (ppf_vote_grouped.cu)
bool   vote_grouped_supported(const ModelTable &m, int ns);
int    vote_grouped_launch(VoteArgs a, int ref_count);
int    vote_grouped_ctas();
size_t vote_grouped_scratch_words(const ModelTable &m);
size_t vote_grouped_replay_cap();

}  // namespace ppf
