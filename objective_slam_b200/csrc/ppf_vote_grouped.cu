// ppf_vote_grouped.cu -- voting_scheme, second formulation of the fused vote kernel.
// Same contract as vote_kernel (ppf_vote.cu: replaces Scene::Scene's ppf_kernel / ppf_hash_kernel pass
// scene.cu:24-55, ParallelHashArray::GetIndices parallel_hash_array.hpp:81-92, ppf_vote_count_kernel,
// ppf_vote_kernel kernel.cu:480-554 and the sort / histogram tail of Model::ComputeUniqueVotes
// model.cu:148-170), same accumulator cells, same bit-exact fast / exact alpha logic.
//
// What changes is WHO shares an ATOMS.  A reference point hits the same bucket many times (on the
// configs[1] workload the vote-weighted multiplicity is ~40): the classical loop lets the 32 lanes of
// a warp read 32 different entries of the bucket for ONE hit, so one ATOMS touches several accumulator
// rows (bank = (bin + row) mod 32 -> 1.8 wavefronts per ATOMS measured) and every entry is fetched and
// decoded once per hit.  Here
//   * one CTA owns one scene reference point: ALL its hits are collected once (phase 1), sorted by
//     bucket in shared memory (bitonic sort of 8-byte records), and reused for every model chunk;
//   * a bucket hit h times is cut into pieces of 32 / 16 / 8 / 4 hits (+ single hits for the rest);
//     a warp votes HG hits x (32 / HG) entries per ATOMS: the lanes of one ATOMS fall into one or two
//     accumulator rows (conflict-free by the row stride, equal cells merged by ATOMS.POPC.INC), an
//     entry is fetched once per HG hits and decoded once (staged in shared memory as
//     (entry, row address)), and a vote costs 5 instructions (IADD, IMAD.WIDE, VIADDMNMX, IMAD, ATOMS)
//     + half an LDS.128;
//   * the accumulator chunk is small (<= 640 model points x 31 bins) so that the queue holds ~11,000 hits; a
//     reference point with more hits than that (dense scenes) is handed to a second instantiation of the
//     kernel that cuts the scene into segments and parks the chunk accumulators in global scratch between them.
// tools/microbench/grouped_vote.cu: 14.2 / 11-13 / 9-11 votes/clk/SM at HG = 32 / 16 / 8 against 6.1 for
// the classical loop inside the full kernel.
#include <algorithm>
#include <cstdlib>

#include "ppf_vote_common.cuh"

namespace ppf {

constexpr uint32_t kGTile       = kHitQueue;        // scene points per phase-1 tile (the scene's tile AABBs)
#ifndef PPF_GSTAGE
#define PPF_GSTAGE 64
#endif
constexpr int      kGStage      = PPF_GSTAGE;       // staging slots per warp: 64 entries per block (two per lane) [+ group padding]
constexpr uint32_t kGGrabVotes  = 32768;            // votes per scheduler ticket of a grouped piece (8192: 699 ms, 16384: 685, 32768: 679, 65536: 689 on configs[1])
constexpr int      kGItemsMax   = 16;               // queue records per thread in the ticket scan
// hit record: [bucket : 20 | (theta_v + half) : 20 | slow : 1 | stored scene index : 23]
constexpr int      kGBucketShift = 44;
constexpr int      kGThetaShift  = 24;
constexpr uint32_t kGIndexMask   = (1u << 23) - 1u;
#ifndef PPF_GTHREADS
#define PPF_GTHREADS 1024
#endif
constexpr int      kGThreads     = PPF_GTHREADS;     // threads per CTA of the grouped kernel
constexpr size_t   kGSmemMax     = 227 * 1024 - 512;  // opt-in maximum minus the kernel's static shared memory

static size_t acc_bytes(int chunk_rows) { return ((size_t)kNAlphaBins * acc_stride(chunk_rows) * 4 + 15) & ~(size_t)15; }

// Largest hit queue (multiple of 1024 records, 12 B each) that fits next to the accumulator.
int vote_grouped_queue_cap(int chunk_rows) {
    const size_t fixed = (size_t)(kGThreads / 32) * kGStage * sizeof(uint2) + acc_bytes(chunk_rows);
    if (fixed + 1024 * 12 > kGSmemMax) return 0;
    const size_t q = (kGSmemMax - fixed) / 12 / 1024 * 1024;
    return (int)std::min<size_t>(q, (size_t)kGItemsMax * 1024);
}
size_t vote_grouped_smem(int chunk_rows) {
    return (size_t)vote_grouped_queue_cap(chunk_rows) * 12 + (size_t)(kGThreads / 32) * kGStage * sizeof(uint2) + acc_bytes(chunk_rows);
}
bool vote_grouped_supported(const ModelTable &m, int ns) {
    return m.chunk_rows <= kGroupedMaxRows && m.U < (1u << 20) && ns <= (int)kGIndexMask &&
           vote_grouped_queue_cap(m.chunk_rows) >= 2 * (int)kGTile;
}

// piece code: 0 = single hit (classical loop), c = 1..4 -> 2^(c+1) = 4 / 8 / 16 / 32 hits
__device__ __forceinline__ uint32_t piece_hits(uint32_t code) { return code ? (2u << code) : 1u; }
// guard band of the grouped loop, folded into the hit word (see vote2)
constexpr uint32_t kGGuardShift = (kGuardLoA + (uint32_t)kNAngle - 1u) / (uint32_t)kNAngle;
constexpr uint32_t kGGuardSpan  = 0u - (uint32_t)kNAngle * kGGuardShift - kGuardLoB;
constexpr uint32_t kGSingleGrab = 4096;             // entries per ticket of a single hit (1024: 681 ms, 2048: 671, 4096: 667)
constexpr uint32_t kGRestCode  = 5;                 // piece code: the 3..31 hits a bucket has beyond its pieces of 32 (vote_rest)
constexpr uint32_t kGRestGrab  = 4096;              // entries per ticket of such a piece
__device__ __forceinline__ uint32_t piece_grab(uint32_t code) {
    return code == kGRestCode ? kGRestGrab : code ? (kGGrabVotes >> (code + 1)) : kGSingleGrab;
}

struct GroupCtx {
    const unsigned long long *queue;
    uint2 *stage;                                   // this warp's kGStage staging slots
    uint32_t trash_addr;                            // pad column of the accumulator (bin 0)
};

// One ticket of a grouped piece: HG = 2 << code hits queue[i0 .. i0 + HG) x entries [pos_grab, pos_grab + ngrab).
// Lane l votes for hit l % HG and belongs to entry group g = l / HG (EG = 32 / HG groups).  A block is 64
// consecutive entries (two coalesced loads per lane); group g owns K = 2 HG of them: its position 2m holds
// entry m EG + g and position 2m + 1 entry 32 + m EG + g, so that (a) the lane that loaded entries l and
// l + 32 stages both with one STS.128, (b) a lane reads two positions per LDS.128 and (c) at every step the
// EG groups vote for EG ADJACENT entries, which mostly share m_r (a bucket is sorted by m_r): the lanes of
// one ATOMS fall into one or two accumulator rows.  HG is a run-time value on purpose: one copy of the loop
// serves every piece size (separate instantiations thrashed the instruction cache: 58% no-instruction stalls).
__device__ __forceinline__ void vote_grouped(const VoteCtx &ctx, const GroupCtx &gc, const FrameYZ &FS, uint32_t i0,
                                             uint32_t code, const uint32_t *__restrict__ entries, uint32_t pos_grab,
                                             uint32_t ngrab, int lane, uint32_t &n_exact) {
    // (pieces of 4 / 8 / 16 hits are no longer cut: the hits a bucket has beyond its pieces of 32 go through vote_rest.
    // HG stays a run-time value all the same: with lhg = 5 as a constant ptxas unrolls the 64-position loop fully and
    // the kernel slows from 591 to 628 ms on configs[1] -- instruction cache.)
    const uint32_t lhg = code + 1u;                 // log2(HG)
    const uint32_t HG = 1u << lhg, EG = 32u >> lhg, K = 2u * HG;
    const uint32_t g = (uint32_t)lane >> lhg;
    // Shared-memory banks.  Group g's K staged pairs start at g * KP (uint2 units).  KP = K puts the groups of small
    // pieces at a byte stride of 16 HG = 128 / 64 B: the 4 (HG = 8) or 8 (HG = 4) groups of one LDS.128 then read the
    // same banks (ncu: 21% of the LDS wavefronts were bank conflicts) -- one 16 B pad per group spreads them over
    // distinct banks.  And the lane that STAGES slot (g, m) is lane g HG + m, so that the 8 lanes of a quarter warp
    // write one contiguous 128 B run (the old mapping, lane = entry index, had them 64 - 256 B apart: STS.128 took
    // 11.2 wavefronts instead of 4).
    const uint32_t KP = K + ((kGStage >= 80 && HG <= 8u) ? 2u : 0u);
    const uint32_t m_slot = (uint32_t)lane & (HG - 1u);
    const uint32_t jl = m_slot * EG + g;            // this lane loads entries jl and jl + 32 of every block
    const unsigned long long rec = gc.queue[i0 + ((uint32_t)lane & (HG - 1u))];
    const uint32_t hit_ones = ((uint32_t)(rec >> kGThetaShift) << kThetaShift) | kLowOnes;
    const bool hit_slow = ((uint32_t)rec >> 23) & 1u;
    const uint32_t s_i = (uint32_t)rec & kGIndexMask;
    const uint32_t S4 = (uint32_t)ctx.stride * 4u;
    // Loop invariants that ptxas would otherwise REMATERIALISE in every block (shared-window base via S2UR +
    // ULEA + LDC + LEA ..., lane via S2R + LOP3: ~20 of the ~60 staging instructions per block): an empty asm
    // makes the value opaque, so it stays in a register for the duration of the ticket.
    uint32_t trash = gc.trash_addr, acc_base = ctx.acc_addr, l32 = jl;
    asm volatile("" : "+r"(trash), "+r"(acc_base), "+r"(l32));
    // this lane loads entries jl and jl + 32 of a block: group g, positions 2 m_slot and 2 m_slot + 1
    uint4 *my_slot = reinterpret_cast<uint4 *>(gc.stage + (g * KP + 2u * m_slot));
    const uint2 *grp = gc.stage + g * KP;
    const uint4 *src = reinterpret_cast<const uint4 *>(grp);
    const uint32_t *__restrict__ ent = entries + pos_grab + jl;

    // exact recount of the staged positions [p0, p1) of this lane's group (rare)
    auto repair = [&](uint32_t p0, uint32_t p1, uint32_t blk_pos) {
#pragma unroll 1
        for (uint32_t p = p0; p < p1; p++) {
            uint2 r = grp[p];
            r.x = 0u - r.x;                                    // staged negated
            const unsigned long long pr = (unsigned long long)(hit_ones - kGGuardShift - r.x) * (unsigned long long)kNAngle;
            const uint32_t bin = (uint32_t)(pr >> 32), margin = (uint32_t)pr;     // the cell the loop incremented
            if (r.y != trash && (margin >= kGGuardSpan || (r.x & kSlowBit) || hit_slow)) {
                const uint32_t j = (p & 1u) * 32u + (p >> 1) * EG + g;          // entry index within the block
                atomicSub(&ctx.acc[bin * (uint32_t)ctx.stride + (r.x & kLocMask)], 1u);
                exact_vote(ctx, FS, s_i, r.x, blk_pos + j);
                n_exact++;
            }
        }
    };
    // hit - entry is computed as max(hit + (-entry), zero) with the entries staged NEGATED and `zero` a kernel argument
    // the compilers cannot fold: add + max fuse into one VIADDMNMX on the ALU pipe.  A plain subtraction becomes
    // IMAD.IADD, i.e. a third op per vote (with the IMAD.WIDE and the address IMAD) on the FMA-heavy pipe, which ncu
    // shows 68% busy over the whole kernel (math-pipe throttle the #3 stall reason) against 36% for the ALU pipe.
    const uint32_t zero = ctx.opaque_zero;
    // The guard-band test costs half an instruction per vote: the hit word is lowered by a = ceil(kGuardLoA / 30), so
    // that the low word of 30 (hit - a - entry) IS the margin (low word - kGuardLoA, up to < 30 units of slack on the
    // safe side) and two margins are max-reduced by one VIMNMX3.  A vote whose low word was below 30 a lands one bin
    // lower (or in bin 29 after a wrap): always a valid cell, and exactly the votes the repair moves anyway -- the
    // repair recomputes the same shifted product to find the cell it has to decrement.
    const uint32_t hit_g = hit_ones - kGGuardShift;
    auto vote2 = [&](const uint4 q, uint32_t &worst) {
        const unsigned long long pa = (unsigned long long)max(hit_g + q.x, zero) * (unsigned long long)kNAngle;
        const unsigned long long pb = (unsigned long long)max(hit_g + q.z, zero) * (unsigned long long)kNAngle;
        red_shared_inc((uint32_t)(pa >> 32) * S4 + q.y);
        red_shared_inc((uint32_t)(pb >> 32) * S4 + q.w);
        worst = max(worst, max((uint32_t)pa, (uint32_t)pb));
    };

    // the entries of the next two blocks are in flight while a block votes
    const uint32_t nblocks = (ngrab + 63u) / 64u;
    uint32_t c0, c1, n0, n1, m0, m1;
    c0 = l32 < ngrab ? __ldg(ent) : 0u;              c1 = l32 + 32u < ngrab ? __ldg(ent + 32) : 0u;
    n0 = l32 + 64u < ngrab ? __ldg(ent + 64) : 0u;   n1 = l32 + 96u < ngrab ? __ldg(ent + 96) : 0u;
    const uint32_t *__restrict__ pre = ent + 128;      // what the loop prefetches next
    uint32_t pre_j = l32 + 128u;                       // its index within the grab
#pragma unroll 1
    for (uint32_t blk = 0; blk < nblocks; blk++) {
        const uint32_t blk0 = blk * 64u;               // first entry of this block within the grab
        m0 = pre_j < ngrab ? __ldg(pre) : 0u;
        m1 = pre_j + 32u < ngrab ? __ldg(pre + 32) : 0u;
        pre += 64; pre_j += 64u;
        const uint32_t nvalid = min(64u, ngrab - blk0);
        uint32_t a0 = acc_base + (c0 & kLocMask) * 4u, a1 = acc_base + (c1 & kLocMask) * 4u;
        bool slow = ((c0 | c1) & kSlowBit) != 0u;
        if (nvalid < 64u) {
            if (l32 >= nvalid) { a0 = trash; c0 = 0u; }
            if (l32 + 32u >= nvalid) { a1 = trash; c1 = 0u; }
            slow = ((c0 | c1) & kSlowBit) != 0u;
        }
        __syncwarp();
        *my_slot = make_uint4(0u - c0, a0, 0u - c1, a1);     // staged as (-entry, row address)
        const unsigned slowmask = __ballot_sync(0xffffffffu, slow);
        __syncwarp();
        const uint32_t worst0 = (hit_slow || slowmask) ? 0xFFFFFFFFu : 0u;
        if (nvalid == 64u) {
            // sub-batches of 8 positions (4 LDS.128), one guard-band test each; K = 8 (HG = 4) or a
            // multiple of 16
            uint32_t p = 0;
            if (K & 8u) {
                const uint4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
                uint32_t worst = worst0;
                vote2(q0, worst); vote2(q1, worst); vote2(q2, worst); vote2(q3, worst);
                if (worst >= kGGuardSpan) repair(0, 8, pos_grab + blk0);
                p = 8;
            }
#pragma unroll 1
            for (; p < K; p += 16) {
                {
                    const uint4 q0 = src[p / 2], q1 = src[p / 2 + 1], q2 = src[p / 2 + 2], q3 = src[p / 2 + 3];
                    uint32_t worst = worst0;
                    vote2(q0, worst); vote2(q1, worst); vote2(q2, worst); vote2(q3, worst);
                    if (worst >= kGGuardSpan) repair(p, p + 8, pos_grab + blk0);
                }
                {
                    const uint4 q0 = src[p / 2 + 4], q1 = src[p / 2 + 5], q2 = src[p / 2 + 6], q3 = src[p / 2 + 7];
                    uint32_t worst = worst0;
                    vote2(q0, worst); vote2(q1, worst); vote2(q2, worst); vote2(q3, worst);
                    if (worst >= kGGuardSpan) repair(p + 8, p + 16, pos_grab + blk0);
                }
            }
        } else {
            // last, partial block: the first 32 entries sit at the even positions, so pairs past
            // ceil(min(nvalid, 32) / EG) hold nothing; invalid slots inside a pair point at the pad column
            const uint32_t pairs = (min(nvalid, 32u) + EG - 1u) / EG;
            uint32_t worst = worst0;
#pragma unroll 1
            for (uint32_t k = 0; k < pairs; k++) {
                const uint4 q = src[k];
                {
                    const unsigned long long p = (unsigned long long)(hit_g + q.x) * (unsigned long long)kNAngle;
                    if (q.y != trash) worst = max(worst, (uint32_t)p);
                    red_shared_inc((uint32_t)(p >> 32) * S4 + q.y);
                }
                {
                    const unsigned long long p = (unsigned long long)(hit_g + q.z) * (unsigned long long)kNAngle;
                    if (q.w != trash) worst = max(worst, (uint32_t)p);
                    red_shared_inc((uint32_t)(p >> 32) * S4 + q.w);
                }
            }
            if (worst >= kGGuardSpan) repair(0, 2u * pairs, pos_grab + blk0);
        }
        c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
}

// The r = 3..31 hits a bucket has beyond its pieces of 32, against entries [pos_grab, pos_grab + ngrab): lane = ENTRY.
// A warp holds 32 x kRestE pre-decoded entries (negated entry word, accumulator row address) in registers and loops
// over the r hits; a hit record is warp-uniform and comes from the sorted queue with one broadcast LDS.64.  No staging,
// no STS, and ONE pass over the entries for any r -- the staged loop needs a power-of-two piece (16 / 8 / 4 hits, then
// single hits through the one-hit loop) and re-stages the entries for each: 3 passes for r = 13.  The lanes of an
// ATOMS are 32 adjacent entries here (random bins: ~1.8 wavefronts against 1.1 in the staged loop, where 32 hits share
// an entry), which is why pieces of 32 keep the staged loop (tools/microbench: 14.2 votes/clk/SM staged at 32 hits
// against 10.9 here; 9.9 / 7.9 here at 8 / 4 hits against 9.2 / ~6 staged, before the staging cost).
#ifndef PPF_REST_LONG_E
#define PPF_REST_LONG_E 8
#endif
#ifndef PPF_REST_MIN
#define PPF_REST_MIN 3
#endif
#ifndef PPF_REST_UNROLL
#define PPF_REST_UNROLL 1
#endif
constexpr int kRestUnroll = PPF_REST_UNROLL;
// kRestE entries per lane: 8 for tables with long bucket slices, 4 for short ones (a block of 32 x 8 slots would be
// mostly empty there: the 20 x 2k-point models of configs[3] leave ~200 entries per bucket and chunk, 14.5 s per scene
// with 8 against 11.8 s with 4; the 10k-point model of configs[1] ~1,150: 574 ms with 8 against 620 with 4)
template <int kRestE>
__device__ __forceinline__ void vote_rest(const VoteCtx &ctx, const GroupCtx &gc, const FrameYZ &FS, uint32_t i0, uint32_t r,
                                          const uint32_t *__restrict__ entries, uint32_t pos_grab, uint32_t ngrab, int lane,
                                          uint32_t &n_exact) {
    const uint32_t S4 = (uint32_t)ctx.stride * 4u;
    uint32_t trash = gc.trash_addr, acc_base = ctx.acc_addr;
    asm volatile("" : "+r"(trash), "+r"(acc_base));
    const uint32_t zero = ctx.opaque_zero;
    const unsigned long long *hits = gc.queue + i0;
    const uint32_t nblocks = (ngrab + 32u * kRestE - 1u) / (32u * kRestE);
    uint32_t raw[kRestE];
#pragma unroll
    for (int u = 0; u < kRestE; u++) {
        const uint32_t j = (uint32_t)u * 32u + (uint32_t)lane;
        raw[u] = j < ngrab ? __ldg(entries + pos_grab + j) : 0u;
    }
#pragma unroll 1
    for (uint32_t blk = 0; blk < nblocks; blk++) {
        const uint32_t blk0 = blk * 32u * kRestE;
        uint32_t neg[kRestE], adr[kRestE], slow_any = 0;
#pragma unroll
        for (int u = 0; u < kRestE; u++) {
            const uint32_t j = blk0 + (uint32_t)u * 32u + (uint32_t)lane;
            const bool valid = j < ngrab;
            neg[u] = valid ? 0u - raw[u] : 0u;
            adr[u] = valid ? acc_base + (raw[u] & kLocMask) * 4u : trash;
            slow_any |= valid ? raw[u] : 0u;
        }
        // the entries of the next block are in flight while this one votes
#pragma unroll
        for (int u = 0; u < kRestE; u++) {
            const uint32_t j = blk0 + 32u * kRestE + (uint32_t)u * 32u + (uint32_t)lane;
            raw[u] = j < ngrab ? __ldg(entries + pos_grab + j) : 0u;
        }
        const uint32_t worst0 = (slow_any & kSlowBit) ? 0xFFFFFFFFu : 0u;
#pragma unroll kRestUnroll
        for (uint32_t h = 0; h < r; h++) {
            const unsigned long long rec = hits[h];
            const uint32_t hit_ones = ((uint32_t)(rec >> kGThetaShift) << kThetaShift) | kLowOnes;
            const uint32_t hit_g = hit_ones - kGGuardShift;
            uint32_t worst = (((uint32_t)rec >> 23) & 1u) ? 0xFFFFFFFFu : worst0;
#pragma unroll
            for (int u = 0; u < kRestE; u += 2) {
                // (max(a + b, 0) is a + b: the DPX form keeps the subtraction one VIADDMNMX on the ALU pipe)
                const unsigned long long pa = (unsigned long long)__viaddmax_u32(hit_g, neg[u], zero) * (unsigned long long)kNAngle;
                const unsigned long long pb = (unsigned long long)__viaddmax_u32(hit_g, neg[u + 1], zero) * (unsigned long long)kNAngle;
                red_shared_inc((uint32_t)(pa >> 32) * S4 + adr[u]);
                red_shared_inc((uint32_t)(pb >> 32) * S4 + adr[u + 1]);
                worst = max(worst, max((uint32_t)pa, (uint32_t)pb));
            }
            if (worst >= kGGuardSpan) {
                // some vote of this hit is not provably in its fast bin (or the hit / an entry is flagged slow): move
                // those from the cell the loop incremented to the exact one
                const bool hit_slow = ((uint32_t)rec >> 23) & 1u;
                const uint32_t s_i = (uint32_t)rec & kGIndexMask;
#pragma unroll
                for (int u = 0; u < kRestE; u++) {
                    if (adr[u] == trash) continue;
                    const uint32_t e = 0u - neg[u];
                    const unsigned long long pr = (unsigned long long)(hit_g - e) * (unsigned long long)kNAngle;
                    if ((uint32_t)pr >= kGGuardSpan || (e & kSlowBit) || hit_slow) {
                        atomicSub(&ctx.acc[(uint32_t)(pr >> 32) * (uint32_t)ctx.stride + (e & kLocMask)], 1u);
                        exact_vote(ctx, FS, s_i, e, pos_grab + blk0 + (uint32_t)u * 32u + (uint32_t)lane);
                        n_exact++;
                    }
                }
            }
        }
    }
}

// first index i < n with (gend[i] >> 3) > t; the caller guarantees that one exists (n <= 16384)
__device__ __forceinline__ uint32_t ticket_owner(const uint32_t *gend, uint32_t n, uint32_t t, int lane) {
    const uint32_t key = (t << 3) | 7u;                          // gend[i] > key  <=>  (gend[i] >> 3) > t
    uint32_t base = 0;
    if (n > 1024u) {                                             // 32 blocks of 512, then 32 of 16, then 16
        const unsigned m = __ballot_sync(0xffffffffu, gend[min((uint32_t)lane * 512u + 511u, n - 1u)] > key);
        base = (uint32_t)(__ffs((int)m) - 1) * 512u;
        const unsigned m1 = __ballot_sync(0xffffffffu, gend[min(base + (uint32_t)lane * 16u + 15u, n - 1u)] > key);
        base += (uint32_t)(__ffs((int)m1) - 1) * 16u;
        const unsigned m2 = __ballot_sync(0xffffffffu, lane < 16 && gend[min(base + (uint32_t)lane, n - 1u)] > key);
        return base + (uint32_t)(__ffs((int)m2) - 1);
    }
    const unsigned m = __ballot_sync(0xffffffffu, gend[min((uint32_t)lane * 32u + 31u, n - 1u)] > key);
    base = (uint32_t)(__ffs((int)m) - 1) * 32u;
    const unsigned m2 = __ballot_sync(0xffffffffu, gend[min(base + (uint32_t)lane, n - 1u)] > key);
    return base + (uint32_t)(__ffs((int)m2) - 1);
}

// first index in queue[0, n) whose record is >= key
__device__ __forceinline__ uint32_t queue_lower_bound(const unsigned long long *queue, uint32_t n, unsigned long long key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (queue[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// SEGMENTS = false: the normal kernel (one hit collection per reference point).  A reference point with more
// hits than the queue holds (dense scenes) is only REGISTERED there (sched[R + 2 ...]) and is processed by the
// SEGMENTS = true instantiation, launched right after: the scene is cut into segments of tiles whose hits fit
// the queue; every segment is collected and sorted once and voted against ALL chunks, the accumulator of a
// chunk waiting in global scratch between two segments (60 KB per chunk and segment each way: noise next to
// the votes).  Two instantiations keep the segment bookkeeping out of the normal kernel's registers
// (one kernel with both paths: 104 bytes of spills, -6% on configs[1]).
template <int THREADS, bool SEGMENTS, int REST_E>
__global__ void __launch_bounds__(THREADS) vote_kernel_grouped(const VoteArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.chunk_rows, S = acc_stride(C);
    const uint32_t Q = (uint32_t)a.queue_cap;
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(smem_raw);        // [Q] hit records
    uint32_t *gend = reinterpret_cast<uint32_t *>(queue + Q);                             // [Q] ticket table
    uint2 *stage_all = reinterpret_cast<uint2 *>(gend + Q);                               // [32 warps][kGStage]
    uint32_t *acc = reinterpret_cast<uint32_t *>(stage_all + (THREADS / 32) * kGStage);   // [31][S] vote counters
    __shared__ uint32_t s_nhits, s_ticket, s_total, s_exact, s_rq;
    __shared__ uint32_t s_red[32];
    __shared__ unsigned long long s_votes;
    // reference point and its frame: only phase 1 and the exact-alpha path read them, so they live in
    // shared memory and not in 15 registers of the vote loop
    __shared__ PointN s_R;
    __shared__ FrameYZ s_FS;

    __shared__ int s_job, s_chunk;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int s_r = 0, p_r = 0;                           // the job's reference point: caller's index / stored position

    for (int i = tid; i < kNAlphaBins * S; i += THREADS) acc[i] = 0;
    if (tid == 0) { s_exact = 0; s_votes = 0; s_nhits = 0; }
    const FrameYZ &FS = s_FS;
    VoteCtx ctx;
    ctx.map = a.map; ctx.mfy = a.mfy; ctx.mfz = a.mfz; ctx.mpos = a.mpos; ctx.spos = a.spos;
    ctx.nm = a.nm; ctx.chunk_base = 0; ctx.stride = S; ctx.acc = acc;
    ctx.acc_addr = (uint32_t)__cvta_generic_to_shared(acc);
    ctx.opaque_zero = a.opaque_zero;
    ctx.rq = a.replay ? a.replay + (size_t)blockIdx.x * a.replay_cap : nullptr;
    ctx.rq_cap = a.replay_cap; ctx.rq_count = &s_rq;
    GroupCtx gc;
    gc.queue = queue; gc.stage = stage_all + warp * kGStage; gc.trash_addr = ctx.acc_addr + (uint32_t)C * 4u;
    unsigned long long my_votes = 0;
    uint32_t my_exact = 0;
    const int items = (int)((Q + THREADS - 1) / THREADS);
    unsigned long long heads = 0;                   // 4 bits per owned queue record: head << 3 | piece code
    __syncthreads();

    // ---- phase 1 for one tile of scene points: pairs (s_r, s_i) -> hit queue (tile / warp culling as in
    // vote_kernel).  Every hit whose cell exists in the table is queued, whatever the chunk.
    auto collect_tile = [&](uint32_t base) {
        const PointN R = s_R;
        const FrameYZ FS = s_FS;
        if (box_dist2(R, __ldg(a.tbox_lo + base / kGTile), __ldg(a.tbox_hi + base / kGTile)) >= a.cull_r2) return;
#pragma unroll 1
        for (uint32_t it = 0; it < kGTile / THREADS; it++) {
            const int i = (int)(base + it * THREADS) + tid;
            bool hit = false;
            unsigned long long h = 0;
            const bool near = (i - lane) < a.ns &&
                              box_dist2(R, __ldg(a.gbox_lo + (i >> 5)), __ldg(a.gbox_hi + (i >> 5))) < a.cull_r2;
            if (near && i < a.ns && i != p_r) {
                float4 p = __ldg(a.spos + i), q = __ldg(a.snrm + i);
                PointN O;
                O.x = p.x; O.y = p.y; O.z = p.z; O.nx = q.x; O.ny = q.y; O.nz = q.z; O.nn = q.w;
                FeatureBins fb = pair_feature_bins(R, O, a.d_dist, a.inv_d);
                const uint32_t b = probe_bucket(a, fb);
                if (b != kNoBucket) {
                    float vy, vz;
                    frame_apply_yz(FS, O.x, O.y, O.z, vy, vz);
                    const uint32_t tc = theta_code(vy, vz);
                    const uint32_t th = ((tc & kThetaMask) + kThetaHalf) & kThetaMask;
                    h = ((unsigned long long)b << kGBucketShift) | ((unsigned long long)th << kGThetaShift) |
                        ((unsigned long long)(tc >> 31) << 23) | (unsigned long long)(uint32_t)i;
                    hit = true;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(&s_nhits, (uint32_t)__popc(m));
                slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
                if (hit && slot < Q) queue[slot] = h;            // the first pass may overflow: it only counts then
            }
        }
    };

    // ---- sort queue[0, n) (bitonic network with the mirrored first stage of every merge: every exchange
    // is ascending, so n need not be a power of two) and cut the bucket groups into pieces.
    auto sort_and_cut = [&](uint32_t n) {
        uint32_t P = 2;
        while (P < n) P <<= 1;
        for (uint32_t k = 2; k <= P; k <<= 1) {
            const uint32_t hk = k >> 1;
            for (uint32_t t = tid; t < P / 2; t += THREADS) {
                const uint32_t blk = (t & ~(hk - 1u)) << 1, o = t & (hk - 1u);      // (t / hk) * k
                const uint32_t i = blk + o, l = blk + k - 1u - o;
                if (l < n) {
                    const unsigned long long x = queue[i], y = queue[l];
                    if (x > y) { queue[i] = y; queue[l] = x; }
                }
            }
            __syncthreads();
            for (uint32_t j = hk >> 1; j > 0; j >>= 1) {
                for (uint32_t t = tid; t < P / 2; t += THREADS) {
                    const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), l = i + j;
                    if (l < n) {
                        const unsigned long long x = queue[i], y = queue[l];
                        if (x > y) { queue[i] = y; queue[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        // A bucket hit h times: h / 32 pieces of 32 hits, then 16 / 8 / 4 by the bits of the remainder,
        // then single hits.
        heads = 0;
        for (int k = 0; k < items; k++) {
            const uint32_t idx = (uint32_t)(tid * items + k);
            if (idx >= n) break;
            const uint32_t b = (uint32_t)(queue[idx] >> kGBucketShift);
            const uint32_t lo = queue_lower_bound(queue, n, (unsigned long long)b << kGBucketShift);
            const uint32_t hi = queue_lower_bound(queue, n, (unsigned long long)(b + 1u) << kGBucketShift);
            const uint32_t h = hi - lo, o = idx - lo, full = h & ~31u;
            uint32_t hc = 0;                                       // head << 3 | code
            if (o < full) { if ((o & 31u) == 0u) hc = 8u | 4u; }
            else {
                // the rest: 3..31 hits as ONE piece (vote_rest, one pass over the entries), 1 or 2 as single hits
                const uint32_t r = h - full, q = o - full;
                if (r >= (uint32_t)PPF_REST_MIN) { if (q == 0u) hc = 8u | kGRestCode; }
                else hc = 8u;
            }
            heads |= (unsigned long long)hc << (4 * k);
        }
    };

    // ---- votes of the sorted queue[0, n) against chunk c
    auto vote_chunk = [&](uint32_t n, const uint2 *__restrict__ ranges) {
        // tickets: the head of a piece owns ceil(slice length / grab); gend = inclusive prefix << 3 | code
        uint32_t sum = 0;
        for (int k = 0; k < items; k++) {
            const uint32_t idx = (uint32_t)(tid * items + k);
            if (idx >= n) break;
            const uint32_t hc = (uint32_t)(heads >> (4 * k)) & 15u;
            if (hc & 8u) {
                const uint32_t len = __ldg(ranges + (uint32_t)(queue[idx] >> kGBucketShift)).y;
                const uint32_t grab = piece_grab(hc & 7u);
                sum += (len + grab - 1u) / grab;
            }
            gend[idx] = (sum << 3) | (hc & 7u);
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_red[warp] = incl;
        if (tid == 0) { s_ticket = 0; s_rq = 0; }
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
        const uint32_t off_t = (s_red[warp] + incl - sum) << 3;
        for (int k = 0; k < items; k++) {
            const uint32_t idx = (uint32_t)(tid * items + k);
            if (idx >= n) break;
            gend[idx] += off_t;
        }
        __syncthreads();
        const uint32_t total = s_total;
        while (true) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&s_ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total) break;
            const uint32_t i0 = ticket_owner(gend, n, t, lane);
            const uint32_t code = gend[i0] & 7u;
            const uint32_t gstart = i0 ? (gend[i0 - 1u] >> 3) : 0u;
            const unsigned long long rec = queue[i0];
            const uint2 rg = __ldg(ranges + (uint32_t)(rec >> kGBucketShift));
            const uint32_t off = (t - gstart) * piece_grab(code);
            const uint32_t ngrab = min(piece_grab(code), rg.y - off);
            const uint32_t pos_grab = rg.x + off;
            if (code == kGRestCode) {
                // hits of this piece: the records from i0 to the end of the bucket's group (at most 31)
                const unsigned long long key = rec >> kGBucketShift;
                const uint32_t idx = i0 + (uint32_t)lane;
                const unsigned same = __ballot_sync(0xffffffffu, idx < n && (queue[idx] >> kGBucketShift) == key);
                const uint32_t r = (uint32_t)__popc(same);
                if (lane == 0) my_votes += (unsigned long long)ngrab * r;
                vote_rest<REST_E>(ctx, gc, FS, i0, r, a.entries, pos_grab, ngrab, lane, my_exact);
                continue;
            }
            if (lane == 0) my_votes += (unsigned long long)ngrab * piece_hits(code);
            if (code == 0u) {
                const uint32_t hit_word = ((uint32_t)(rec >> kGThetaShift) << kThetaShift) |
                                          ((((uint32_t)rec >> 23) & 1u) << kLocBits);
                vote_single_hit<true>(ctx, FS, a.entries, hit_word, (uint32_t)rec & kGIndexMask, pos_grab, ngrab, lane, my_exact);
            } else {
                // (pieces of 32 on SHORT slices through vote_rest instead -- slices of <= 128 / 256 / 512 entries --
                // measured: 575 / 575 / 574 ms against 573.5: no gain)
                vote_grouped(ctx, gc, FS, i0, code, a.entries, pos_grab, ngrab, lane, my_exact);
            }
        }
        __syncthreads();
        // replay of the deferred exact votes of this chunk: one record per thread, all lanes busy
        if (ctx.rq) {
            const uint32_t nr = min(s_rq, ctx.rq_cap);
            for (uint32_t i = tid; i < nr; i += THREADS) {
                const uint2 r = ctx.rq[i];
                atomicAdd(&ctx.acc[exact_vote_index_at(ctx, FS, r.y, r.x)], 1u);
            }
            __syncthreads();
        }
    };

    // ---- jobs.  Persistent CTAs draw reference points from sched[0]; the chunks of reference point r are
    // drawn from sched[1 + r].  When the reference points run out, an idle CTA becomes a HELPER: it picks the
    // reference point with the most chunks left, repeats its hit collection (a few % of its work) and
    // draws chunks from the same counter, so the heaviest reference points do not leave the other SMs idle.
    uint32_t *const overflow = a.sched + 1 + a.ref_count;      // [0] = count, [1 ...] = jobs, [1 + R] = next to process
    int c_lo = 0, c_hi = a.n_chunks;                           // SEGMENTS: the chunks of this job
    while (true) {
        __syncthreads();
        int job;
        if constexpr (SEGMENTS) {
            // A job is (dense reference point, group of chunks).  Dense scenes register every reference point: one
            // group = all chunks, so that a segment is collected once for all of them.  A sparse scene registers only a
            // few outliers (configs[1] at ref_point_df = 1: a handful of points with more than 11k hits, 40 ms each on
            // ONE SM while 147 idle): their chunks are spread over the idle CTAs, each repeating the collection.
            const uint32_t n_dense = *(volatile uint32_t *)&overflow[0];
            if (n_dense == 0u) break;
            const uint32_t groups = min((uint32_t)a.n_chunks, max(1u, gridDim.x / n_dense));
            if (tid == 0) {
                const uint32_t k = atomicAdd(&overflow[1 + a.ref_count], 1u);
                s_job = k < n_dense * groups ? (int)overflow[1 + k / groups] : -1;
                s_chunk = (int)(k % groups);
            }
            __syncthreads();
            job = s_job;
            if (job < 0) break;
            const uint32_t g = (uint32_t)s_chunk;
            c_lo = (int)(g * (uint32_t)a.n_chunks / groups);
            c_hi = (int)((g + 1u) * (uint32_t)a.n_chunks / groups);
        } else {
        if (tid == 0) s_job = (int)atomicAdd(&a.sched[0], 1u);
        __syncthreads();
        job = s_job;
        if (job >= a.ref_count) {
            // helper: reference point with the most chunks not yet drawn (ties: spread by CTA)
            unsigned long long best = 0;
            for (int r = tid; r < a.ref_count; r += THREADS) {
                const uint32_t taken = *(volatile uint32_t *)(a.sched + 1 + r);
                if (taken < (uint32_t)a.n_chunks) {
                    const uint32_t left = (uint32_t)a.n_chunks - taken;
                    const uint32_t mix = ((uint32_t)r * 2654435761u + blockIdx.x * 40503u) >> 20;
                    best = max(best, ((unsigned long long)left << 44) | ((unsigned long long)mix << 32) | (uint32_t)(r + 1));
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            __syncthreads();
            if (lane == 0) reinterpret_cast<unsigned long long *>(gend)[warp] = best;
            __syncthreads();
            best = 0;
            for (int w = 0; w < THREADS / 32; w++) best = max(best, reinterpret_cast<unsigned long long *>(gend)[w]);
            if (best == 0) break;                    // nothing left anywhere
            job = (int)(uint32_t)best - 1;
        }
        }
        s_r = a.ref_start + job * a.ref_stride;
        p_r = (int)__ldg(a.sinv + s_r);
        __syncthreads();
        if (tid == 0) {
            float4 p = __ldg(a.spos + p_r), q = __ldg(a.snrm + p_r);
            PointN R;
            R.x = p.x; R.y = p.y; R.z = p.z; R.nx = q.x; R.ny = q.y; R.nz = q.z; R.nn = q.w;
            s_R = R;
            s_FS = load_frame(a.sfy, a.sfz, p_r);
            s_nhits = 0;
        }
        __syncthreads();

        uint32_t n = 0;
        int c = -1;
        bool sorted = false;
        // SEGMENTS only:
        bool more_segments = false;
        uint32_t base = 0;
        int seg = -1;
        if constexpr (!SEGMENTS) {
            // ---- first pass: all hits of the reference point, whatever the chunk
            // (every 4 tiles: stop as soon as the queue has overflowed -- on a dense scene every reference point does,
            // after a fraction of the scene, and the dense-point kernel collects it again anyway.  Block-uniform: the
            // count is read between two barriers.)
            for (uint32_t base0 = 0, k = 0; base0 < (uint32_t)a.ns; base0 += kGTile, k++) {
                collect_tile(base0);
                if ((k & 3u) == 3u) {
                    __syncthreads();
                    const uint32_t have = s_nhits;
                    __syncthreads();
                    if (have > Q) break;
                }
            }
            __syncthreads();
            n = s_nhits;
            if (n > Q) {
                // does not fit: claim every chunk (helpers must not touch it) and leave it to the segment kernel
                if (tid == 0 && atomicCAS(&a.sched[1 + job], 0u, (uint32_t)a.n_chunks) == 0u)
                    overflow[1 + atomicAdd(&overflow[0], 1u)] = (uint32_t)job;
                continue;
            }
        }

        while (true) {
            if constexpr (!SEGMENTS) {
                __syncthreads();
                if (tid == 0) s_chunk = (int)atomicAdd(&a.sched[1 + job], 1u);
                __syncthreads();
                c = s_chunk;
                if (c >= a.n_chunks) break;
                if (!sorted && n) { sort_and_cut(n); sorted = true; }
            } else {
                if (c < 0 || c == c_hi - 1) {
                    if (c >= 0 && !more_segments) break;            // the last segment has met every chunk of the job
                    seg++; c = c_lo;
                    __syncthreads();
                    if (tid == 0) s_nhits = 0;
                    __syncthreads();
                    // block-uniform loop condition: every thread reads s_nhits between two barriers, before any
                    // warp of the next collect_tile can bump it
                    while (base < (uint32_t)a.ns) {
                        const uint32_t have = s_nhits;
                        __syncthreads();
                        if (have + kGTile > Q) break;
                        collect_tile(base);
                        base += kGTile;
                        __syncthreads();
                    }
                    n = s_nhits;
                    more_segments = base < (uint32_t)a.ns;
                    if (n) sort_and_cut(n);
                } else {
                    c++;
                }
                if (seg > 0 && c_hi - c_lo > 1) {                   // (a single chunk stays in shared memory)
                    const uint32_t *src = a.acc_scratch + ((size_t)blockIdx.x * a.n_chunks + c) * ((size_t)kNAlphaBins * S);
                    for (int i = tid; i < kNAlphaBins * S; i += THREADS) acc[i] = src[i];
                }
                __syncthreads();
            }
            const uint2 *__restrict__ ranges = a.ranges + (size_t)c * a.U;
            ctx.chunk_base = c * C;
            if (n) vote_chunk(n, ranges);
            __syncthreads();
            if constexpr (SEGMENTS) {
                if (more_segments) {
                    if (c_hi - c_lo > 1) {
                        uint32_t *dst = a.acc_scratch + ((size_t)blockIdx.x * a.n_chunks + c) * ((size_t)kNAlphaBins * S);
                        for (int i = tid; i < kNAlphaBins * S; i += THREADS) { dst[i] = acc[i]; acc[i] = 0; }
                    }
                    continue;
                }
            }
        __syncthreads();
        // ---- phase 3 for this chunk: block max, statistics, emission of candidate cells, reset.
        // (the pad column is scratch: partial blocks vote into it)
        uint32_t lmax = 0, nz = 0;
        for (int i = tid; i < kNAlphaBins * S; i += THREADS) {
            uint32_t v = acc[i];
            if ((i % S) == C) v = 0;
            lmax = max(lmax, v);
            nz += (v != 0);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
            nz += __shfl_xor_sync(0xffffffffu, nz, o);
        }
        if (lane == 0) {
            s_red[warp] = lmax;
            if (nz) atomicAdd(&a.totals[1], (unsigned long long)nz);
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t v = (lane < THREADS / 32) ? s_red[lane] : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) {
                const uint32_t old = v ? atomicMax(&a.scalars[1], v) : 0u;
                s_red[0] = max(old, v);
            }
        }
        __syncthreads();
        const uint32_t bound = s_red[0];                       // <= final global max
        const float min_votecount = a.emit_all ? 0.0f : a.thr * (float)bound;     // model.cu:164
        for (int i = tid; i < kNAlphaBins * S; i += THREADS) {
            const uint32_t v = acc[i];
            if (v == 0) continue;
            acc[i] = 0;
            const uint32_t bin = i / S, loc = i - bin * S;
            if ((int)loc != C && (float)v > min_votecount) {
                const uint32_t slot = atomicAdd(&a.scalars[0], 1u);
                if (slot < a.cand_cap) {
                    // [scene ref : 32 | model point : 26 | alpha : 6]   (kernel.cu:548-549, model.h:61-63)
                    a.cand_codes[slot] = ((unsigned long long)(uint32_t)s_r << 32) |
                                         (unsigned long long)((((uint32_t)(c * C) + loc) << 6) | bin);
                    a.cand_counts[slot] = v;
                }
            }
        }
        __syncthreads();
        }   // chunks of this job
    }       // jobs

#pragma unroll
    for (int o = 16; o; o >>= 1) my_exact += __shfl_xor_sync(0xffffffffu, my_exact, o);
    if (lane == 0) {
        if (my_votes) atomicAdd(&s_votes, my_votes);
        if (my_exact) atomicAdd(&s_exact, my_exact);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_votes) atomicAdd(&a.totals[0], s_votes);
        if (s_exact) atomicAdd(&a.scalars[3], s_exact);
    }
}

int vote_grouped_ctas() {
    int n_sm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (const char *e = getenv("PPF_B200_VOTE_CTAS")) n_sm = std::max(1, atoi(e));
    return n_sm;
}
// deferred exact votes: records per CTA (8 B each); ~1.35e-4 of the votes of one (reference point, chunk) pass
size_t vote_grouped_replay_cap() { return 16384; }
size_t vote_grouped_scratch_words(const ModelTable &m) {
    return (size_t)vote_grouped_ctas() * m.n_chunks * ((size_t)kNAlphaBins * acc_stride(m.chunk_rows));
}

int vote_grouped_launch(VoteArgs a, int ref_count) {
    a.queue_cap = vote_grouped_queue_cap(a.chunk_rows);
    if (const char *e = getenv("PPF_B200_VOTE_QUEUE")) {          // test hook: force the queue-overflow path
        const int v = atoi(e) / 1024 * 1024;
        if (v >= 2 * (int)kGTile && v <= a.queue_cap) a.queue_cap = v;
    }
    // persistent CTAs (one per SM: the kernel takes all of its shared memory) draw (reference point, chunk)
    // work from the counters in a.sched (zeroed by the caller)
    const long long grid = vote_grouped_ctas();
    const size_t smem = vote_grouped_smem(a.chunk_rows);
    // entries per lane of vote_rest by the average bucket slice per chunk (see vote_rest)
    const bool long_slices = a.rest_long != 0;
    if (long_slices) {
        PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel_grouped<kGThreads, false, PPF_REST_LONG_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel_grouped<kGThreads, true, PPF_REST_LONG_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vote_kernel_grouped<kGThreads, false, PPF_REST_LONG_E><<<(unsigned)grid, kGThreads, smem, cur_stream()>>>(a);
        // reference points whose hits did not fit the queue (none on sparse scenes: the kernel then exits at once)
        vote_kernel_grouped<kGThreads, true, PPF_REST_LONG_E><<<(unsigned)grid, kGThreads, smem, cur_stream()>>>(a);
    } else {
        PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel_grouped<kGThreads, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PPF_CUDA_TRY(cudaFuncSetAttribute(vote_kernel_grouped<kGThreads, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vote_kernel_grouped<kGThreads, false, 4><<<(unsigned)grid, kGThreads, smem, cur_stream()>>>(a);
        vote_kernel_grouped<kGThreads, true, 4><<<(unsigned)grid, kGThreads, smem, cur_stream()>>>(a);
    }
    count_launch(2);
    return PPF_OK;
}

}  // namespace ppf
