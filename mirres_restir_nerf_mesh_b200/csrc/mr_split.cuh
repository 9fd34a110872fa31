// mirres-b200: ordered subtree splitting for closest-hit rays.
//
// bvh_hit_with_normal (nerf/ScreenSpaceReSTIR/utils/helperDi.slang:313-395) is a sequential fold over the leaves it
// visits: the result depends on the visit order (negative-t hits, ties, `t <= closest` keeps the LATEST of equal hits),
// so round 1 walked every such ray on one lane and the launches ended in a tail of a few ~10^2-step rays on an idle
// GPU (ncu: 7.5-9.3 of 32 lanes, 16-18 % warps active).  Here idle lanes of the ray's warp take over the OLDEST
// deferred subtree of a busy lane -- the bottom of its stack, i.e. the piece the reference would visit LAST -- and walk
// it speculatively; the result is reconstructed exactly:
//
//   * The reference visits leaf L iff every box on the path passes `min(closest, t_far) > t_near` at the time it is
//     popped.  `closest` never increases, and t_near never decreases from a box to a box inside it (exact containment +
//     monotone fp32 slab arithmetic, see mr_bvh.cuh), so the conjunction over the path collapses to the leaf's own test
//     with the value `closest` has when the walk reaches it:  visited(L)  <=>  t_far > t_near on the path  AND
//     closest_before(L) > t_near(L).
//   * A thief walks its subtree with a STALE bound (the donor's bound at the moment of the steal, which is >= every
//     value `closest` can have later in visit order), so it visits a superset of the leaves the reference visits.  It
//     does not touch the ray's state; it LOGS every triangle line-hit as (order key, t_near(L), t, L).  It may tighten
//     its own bound with a logged hit only if t >= t_near(L): if the replay later rejects that hit, then
//     closest_before(L) <= t_near(L) <= t, so the true bound was at least as tight and nothing the reference visits was
//     skipped.  (Hits with t < t_near(L) -- negative-t hits, rounding -- are logged but never used to prune.)
//   * Tasks of one ray are totally ordered like the reference's visit order: a task owns an interval [lo, hi) of an
//     order space; a steal hands the upper half [mid, hi) to the thief with the donor's bottom entry -- everything the
//     donor still owns comes earlier in visit order than what it gave away.  The first task (the ray's home) holds the
//     true state; when the last task of a ray retires, the log is replayed in key order (stable, so hits of one task stay
//     in the order they were found):  if (closest > e.t_near) { if (e.t <= closest) best = e.leaf; closest = min(e.t,
//     closest); any = true; }  -- exactly the reference's update for exactly the leaves it would have visited.
//   * A log that overflows (more than MR_SPLIT_LOG line-hits found by thieves) makes the ray start over on one lane.
//
// The same step / replay code runs in the host-check flavour under a randomised scheduler (mirres_test_closest_split)
// against the oracle; the warp plumbing around it is in wave.cu.
#pragma once
#include "mr_bvh.cuh"

namespace mr {

#define MR_SPLIT_LOG 8
#define MR_SPLIT_LANES 32

struct SplitLogEntry {
    unsigned int key; // `lo` of the task that found the hit
    float tnear;      // entry distance of the leaf's box
    float t;          // line parameter of the hit
    int leaf;         // leaf slot
};

// one set per warp; indexed by the ray's home lane.  CAP = log capacity per ray (1 for walkers that never split)
template <int CAP_>
struct SplitRecT {
    static constexpr int CAP = CAP_;
    float bound[MR_SPLIT_LANES];       // current closest distance of the ray's first task (a valid bound for every later task)
    float fin_closest[MR_SPLIT_LANES]; // state of the first task when it retired
    int fin_best[MR_SPLIT_LANES];
    int fin_any[MR_SPLIT_LANES];
    int pending[MR_SPLIT_LANES];       // live tasks of the ray
    int nlog[MR_SPLIT_LANES];          // hits appended (> MR_SPLIT_LOG: overflow)
    int overflow;                      // some task of this warp dropped a stack entry (MR_STACK)
    SplitLogEntry log[MR_SPLIT_LANES][CAP_];
};
typedef SplitRecT<MR_SPLIT_LOG> SplitRec;

struct CTask {
    Ray r;
    int slot;            // result slot of the ray
    int home;            // index of the ray's SplitRec entry
    bool first;          // holds the true state (first task in visit order)
    bool nosplit;        // restarted after a log overflow: may not be robbed
    unsigned int lo, hi; // order interval
    int sp, bot;         // live entries [bot, sp) of the task's stack (the arrays live beside the task: a struct that holds
                         // dynamically indexed arrays is placed in local memory as a whole, scalars included)
    int cur;             // reference being processed
    float cur_t;         // its entry distance
    float closest;       // first task: the ray's closest distance; others: the bound they prune with
    int best;
    bool any;
};

MR_DEV void task_start_ray(CTask &T, float3 o, float3 d, int slot, int home, bool nosplit)
{
    T.r = make_ray(o, d);
    T.slot = slot;
    T.home = home;
    T.first = true;
    T.nosplit = nosplit;
    T.lo = 0u;
    T.hi = 0xffffffffu;
    T.sp = T.bot = 0;
    T.cur = 0;
    T.cur_t = 0.f;
    T.closest = 1e7f;
    T.best = -1;
    T.any = false;
}

#if defined(__CUDA_ARCH__)
#define MR_SPLIT_APPEND(ctr) atomicAdd((ctr), 1)
#define MR_SPLIT_VOLATILE(x) (*(volatile float *)&(x))
#else
#define MR_SPLIT_APPEND(ctr) ((*(ctr))++)
#define MR_SPLIT_VOLATILE(x) (x)
#endif

// One visit (wide record or leaf) of a task.  Returns true when the task has no work left.
template <class REC>
MR_DEV bool task_step(const BvhView &bvh, CTask &T, int *__restrict__ stack_ref, float *__restrict__ stack_t, REC &rec)
{
    const Rec32 *rp = ref_address(bvh, T.cur);
    const Rec32 e0 = load_rec(rp), e1 = load_rec(rp + 1);
    // later tasks also prune with the first task's current distance: it bounds `closest` everywhere later in visit order
    float bound = T.first ? T.closest : fminf(T.closest, MR_SPLIT_VOLATILE(rec.bound[T.home]));
    bool pop = false;
    if (T.cur >= 0) {
        const Rec32 e2 = load_rec(rp + 2), e3 = load_rec(rp + 3);
        WideHit w;
        wide_slabs(T.r, e0, e1, e2, e3, w);
        int next = 0;
        float next_t = 0.f;
        bool got = false;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            if (fminf(bound, w.tf[k]) > w.tn[k]) {
                if (got) {
                    stack_ref[T.sp] = next;
                    stack_t[T.sp] = next_t;
                    ++T.sp;
                }
                next = w.ref[k];
                next_t = w.tn[k];
                got = true;
            }
        }
        if (T.sp > MR_STACK) { // at most three entries were deferred into the slack of the arrays: dropped and reported
            T.sp = MR_STACK;
            rec.overflow = 1;
        }
        if (got) {
            T.cur = next;
            T.cur_t = next_t;
        } else {
            pop = true;
        }
    } else {
        const int leaf = ~T.cur;
        float t, u, v;
        if (tri_test(T.r, e0, e1, t, u, v)) {
            if (T.first) {
                if (t <= T.closest) T.best = leaf;
                T.closest = fminf(t, T.closest);
                T.any = true;
                // (unsynchronised on purpose: a later task that reads this word while it changes sees the old or the new
                // distance, both valid bounds for it; compute-sanitizer racecheck reports exactly this pair as a warning)
                rec.bound[T.home] = T.closest;
                bound = T.closest;
            } else {
                const int idx = MR_SPLIT_APPEND(&rec.nlog[T.home]);
                if (idx < REC::CAP) {
                    SplitLogEntry e;
                    e.key = T.lo;
                    e.tnear = T.cur_t;
                    e.t = t;
                    e.leaf = leaf;
                    rec.log[T.home][idx] = e;
                }
                if (t >= T.cur_t) {
                    T.closest = fminf(T.closest, t);
                    bound = fminf(bound, t);
                }
            }
        }
        pop = true;
    }
    if (pop) {
        while (T.sp > T.bot) {
            --T.sp;
            if (bound > stack_t[T.sp]) {
                T.cur = stack_ref[T.sp];
                T.cur_t = stack_t[T.sp];
                return false;
            }
        }
        return true;
    }
    return false;
}

// The ray's result from the first task's final state and the log; `overflow` = the log lost entries.
template <class REC>
MR_DEV bool task_replay(const REC &rec, int home, float &closest, int &best, bool &any)
{
    closest = rec.fin_closest[home];
    best = rec.fin_best[home];
    any = rec.fin_any[home] != 0;
    const int n = rec.nlog[home];
    if (n > REC::CAP) return false;
    unsigned int done = 0u;
    for (int k = 0; k < n; ++k) {
        // next entry in (key, append order)
        int pick = -1;
        unsigned int pick_key = 0u;
        for (int j = 0; j < n; ++j) {
            if (done & (1u << j)) continue;
            const unsigned int kj = rec.log[home][j].key;
            if (pick < 0 || kj < pick_key) {
                pick = j;
                pick_key = kj;
            }
        }
        done |= 1u << pick;
        const SplitLogEntry e = rec.log[home][pick];
        if (closest > e.tnear) {
            if (e.t <= closest) best = e.leaf;
            closest = fminf(e.t, closest);
            any = true;
        }
    }
    return true;
}

MR_DEV bool task_can_donate(const CTask &T) { return T.sp > T.bot && !T.nosplit && (T.hi - T.lo) >= 2u; }

} // namespace mr
