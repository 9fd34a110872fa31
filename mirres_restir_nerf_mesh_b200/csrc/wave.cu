// mirres-b200: workspace preparation = ordered compaction of the foreground pixels (occ >= 0.1).
//
// Every ray-casting stage of the reference early-outs on `occ_map[pixel] < 0.1` (e.g. InitialResampling.slang:166,
// SpatialResampling.slang:192, FinalShading.slang:166,760).  The list of pixels that survive that test is built
// once per frame here and shared by all wavefront stages (mr_wave.cuh); ascending pixel order keeps neighbouring
// lanes on neighbouring surface points.
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include "mr_wave.cuh"
#include "mr_split.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

#define CP_BLOCK 256
#define CP_ITEMS 4
#define CP_TILE (CP_BLOCK * CP_ITEMS)

#if !defined(MR_HOST_CHECK)
__global__ void __launch_bounds__(CP_BLOCK) k_compact_count(const float *__restrict__ occ, int n, int *__restrict__ block_counts)
{
    __shared__ int warp_sums[CP_BLOCK / 32];
    const int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    int c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j)
        if (base + j < n && !(__ldg(occ + base + j) < 0.1f)) ++c;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
#pragma unroll
        for (int w = 0; w < CP_BLOCK / 32; ++w) s += warp_sums[w];
        block_counts[blockIdx.x] = s;
    }
}

// exclusive scan of block_counts[nb] in place; total -> counters[0]; single block
__global__ void __launch_bounds__(1024) k_compact_scan(int *block_counts, int nb, int *counters)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_counts[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + x - v;
        if (i < nb) block_counts[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counters[0] = carry;
        counters[1] = 0;
    }
    if (threadIdx.x >= MR_CTR_ALIVE_SIZE && threadIdx.x <= MR_CTR_ALIVE_LAST) counters[threadIdx.x] = 0; // the lists refer to the old pixel list
}

__global__ void __launch_bounds__(CP_BLOCK) k_compact_write(const float *__restrict__ occ, int n, const int *__restrict__ block_offsets,
                                                            int *__restrict__ active)
{
    __shared__ int warp_sums[CP_BLOCK / 32];
    const int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
    bool f[CP_ITEMS];
    int c = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j) {
        f[j] = base + j < n && !(__ldg(occ + base + j) < 0.1f);
        c += f[j] ? 1 : 0;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_sums[w];
    int dst = block_offsets[blockIdx.x] + woff + x - c;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; ++j)
        if (f[j]) active[dst++] = base + j;
}
#endif

#if !defined(MR_HOST_CHECK)
// ---- persistent queue tracers ------------------------------------------------------------------------------------------
// Warps pull rays from a dense queue with one atomic per refill; a lane whose ray terminates (first hit, or stack empty)
// is refilled at the next check point, so the SIMD lanes stay busy although path lengths vary by > 10x.
#define MR_TRACE_BLOCK 256
// visits between two refill / steal check points of a warp (B200 sweep on the C2 step: 4: 4.33 ms, 6: 4.26, 10: 4.235,
// 16: 4.24, 24: 4.34)
#ifndef MR_TRACE_STEPS
#define MR_TRACE_STEPS 10
#endif
#ifndef MR_TRACE_STEPS_SHARED
#define MR_TRACE_STEPS_SHARED 4
#endif
#define MR_SPLIT_ROUNDS 4
#define MR_TRACE_STEPS_SPLIT 3

// Any-hit.  When a warp cannot refill all of its idle lanes (queue dry, or -- for small queues -- the per-warp grab limit
// is reached) the idle lanes take over the OLDEST deferred subtree of busy lanes.  bvh_hit's boolean result is the OR over
// all leaf tests (mr_bvh.cuh), so walking the subtrees of one ray on several lanes returns the same flag while the
// longest ray of a launch stops being a serial chain of ~10^3 dependent L2 loads.
// TOP: the first wide levels of the tree are read from `s_top`, the block's shared-memory copy of bvh.top (see TopTable)
template <bool TOP>
__device__ __forceinline__ void trace_any_worker(const BvhView &bvh, const Workspace &ws, int grab, const Rec32 *s_top)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[MR_CTR_ANY_SIZE];
    int *ticket = ws.counters + MR_CTR_ANY_TICKET;
    int stack[MR_STACK + MR_STACK_SLACK];
    int sp = 0, bot = 0;   // live entries: [bot, sp)
    int cur = 0;           // node reference being processed: >= 0 internal, < 0 leaf
    int slot = -1;
    bool have = false;      // this lane owns a live (ray, subtree)
    bool found = false;     // this lane has just resolved its ray with a hit
    bool exhausted = false; // warp-uniform: the queue has been drained
    bool shared = false;    // warp-uniform: some ray of this warp is being walked by more than one lane
    bool overflow = false;  // this lane dropped a stack entry (MR_STACK)
    Ray r;
    r.o = r.d = r.inv = f3(0.f);
    for (;;) {
        unsigned int need = __ballot_sync(FULL, !have);
        if (need != 0u && !exhausted) {
            const int n_take = min(__popc(need), grab);
            const int leader = __ffs(need) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(ticket, n_take);
            base = __shfl_sync(FULL, base, leader);
            const int rank = __popc(need & lt_mask);
            if (!have && rank < n_take) {
                const int s = base + rank;
                if (s < total) {
                    const float4 o = __ldg(ws.ray_o + s);
                    const float4 d = __ldg(ws.ray_d + s);
                    r = make_ray(make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z));
                    slot = __float_as_int(o.w);
                    sp = bot = 0;
                    cur = (TOP && s_top) ? MR_REF_TOP : 0; // table slot 0 = the root record
                    have = true;
                }
            }
            if (base + n_take >= total) exhausted = true;
        }
        const unsigned int busy = __ballot_sync(FULL, have);
        if (busy == 0u) {
            if (exhausted) break;
            continue;
        }
        if (busy != FULL) {
            // ---- steal: the k-th idle lane takes the oldest deferred subtree of the k-th lane that has one
            const unsigned int donors = __ballot_sync(FULL, have && sp > bot);
            if (donors != 0u) {
                const unsigned int idle = ~busy;
                const int n_pairs = min(__popc(donors), __popc(idle));
                const bool robbed = have && sp > bot && __popc(donors & lt_mask) < n_pairs;
                int give = 0;
                if (robbed) give = stack[bot++];
                const int my_rank = __popc(idle & lt_mask);
                const bool thief = !have && my_rank < n_pairs;
                int src = (int)lane;
                if (thief) {
                    unsigned int m = donors;
                    for (int j = 0; j < my_rank; ++j) m &= m - 1u;
                    src = __ffs(m) - 1;
                }
                const int g_node = __shfl_sync(FULL, give, src);
                const int g_slot = __shfl_sync(FULL, slot, src);
                const float ox = __shfl_sync(FULL, r.o.x, src), oy = __shfl_sync(FULL, r.o.y, src), oz = __shfl_sync(FULL, r.o.z, src);
                const float dx = __shfl_sync(FULL, r.d.x, src), dy = __shfl_sync(FULL, r.d.y, src), dz = __shfl_sync(FULL, r.d.z, src);
                const float ix = __shfl_sync(FULL, r.inv.x, src), iy = __shfl_sync(FULL, r.inv.y, src), iz = __shfl_sync(FULL, r.inv.z, src);
                if (thief) {
                    r.o = make_float3(ox, oy, oz);
                    r.d = make_float3(dx, dy, dz);
                    r.inv = make_float3(ix, iy, iz);
                    slot = g_slot;
                    cur = g_node;
                    sp = bot = 0;
                    have = true;
                }
                shared = true;
            }
        }
        const int steps = shared ? MR_TRACE_STEPS_SHARED : MR_TRACE_STEPS;
#pragma unroll 1
        for (int it = 0; it < steps; ++it) {
            if (!have) break;
            // load phase common to both record kinds, so that node lanes and leaf lanes of a diverged warp wait for
            // their L2 round trip at the same time
            const bool in_top = TOP && cur >= 0 && (cur & MR_REF_TOP) != 0;
            const Rec32 *rec = in_top ? s_top + 4 * (cur & 0x1ff) : ref_address(bvh, cur);
            Rec32 e0, e1;
            if (in_top) { e0 = rec[0]; e1 = rec[1]; }
            else { e0 = load_rec(rec); e1 = load_rec(rec + 1); }
            if (cur >= 0) {
                Rec32 e2, e3;
                if (in_top) { e2 = rec[2]; e3 = rec[3]; }
                else load_tail(rec, ref_missing(cur), e2, e3); // trailing sectors only when the reference says they are in use
                WideHit w;
                wide_slabs(r, e0, e1, e2, e3, w);
                int next = 0;
                bool got = false;
#pragma unroll
                for (int k = 3; k >= 0; --k) {
                    if (fminf(1e7f, w.tf[k]) > w.tn[k]) {
                        if (got) {
                            stack[sp++] = next;
                        }
                        next = w.ref[k];
                        got = true;
                    }
                }
                if (sp > MR_STACK) { // a visit defers at most three entries: the array has that much slack (MR_STACK)
                    sp = MR_STACK;
                    overflow = true;
                }
                if (got) {
                    cur = next;
                } else if (sp > bot) {
                    cur = stack[--sp];
                } else {
                    have = false;
                }
            } else {
                float t, u, v;
                if (tri_test(r, e0, e1, t, u, v)) {
                    ws.hit[slot] = MR_HIT_HIT;
                    have = false;
                    found = true;
                } else if (sp > bot) {
                    cur = stack[--sp];
                } else {
                    have = false;
                }
            }
        }
        if (shared) {
            // a hit found by one lane resolves the ray for every lane that walks a piece of it
            const unsigned int fmask = __ballot_sync(FULL, found);
            if (fmask != 0u) {
                const unsigned int peers = __match_any_sync(FULL, (have || found) ? slot : -1 - (int)lane);
                if (have && (peers & fmask) != 0u) have = false;
            }
        }
        found = false;
        __syncwarp();
    }
    if (overflow) ws.counters[MR_CTR_ERROR] = 1;
}

// Closest-hit: same refill scheme and the reference's visit order (closest_hit in mr_bvh.cuh); lanes that the queue
// cannot feed take over the oldest deferred subtree of a busy lane and the result is rebuilt exactly from their hit logs
// (mr_split.cuh).
__device__ __forceinline__ void closest_write_result(const BvhView &bvh, const Workspace &ws, const CTask &T, float closest,
                                                     int best, bool any)
{
    float3 pos = f3(0.f), n = f3(1.0f);
    int prim = -1;
    float bary[2] = {0.f, 0.f};
    if (any) {
        pos = T.r.o + closest * T.r.d;
        closest_finish(bvh, T.r, best, n, prim, bary);
    }
    ws.chit[3 * (size_t)T.slot] = make_float4(pos.x, pos.y, pos.z, any ? 1.0f : 0.0f);
    ws.chit[3 * (size_t)T.slot + 1] = make_float4(n.x, n.y, n.z, any ? closest : 0.f);
    ws.chit[3 * (size_t)T.slot + 2] = make_float4(__int_as_float(prim), bary[0], bary[1], 0.f);
}

template <bool SPLIT, class REC>
__device__ __forceinline__ void trace_closest_worker(const BvhView &bvh, const Workspace &ws, REC &rec, int grab)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[MR_CTR_CLOSEST_SIZE];
    int *ticket = ws.counters + MR_CTR_CLOSEST_TICKET;
    CTask T;
    int stack_ref[MR_STACK + MR_STACK_SLACK];
    float stack_t[MR_STACK + MR_STACK_SLACK];
    T.r.o = T.r.d = T.r.inv = f3(0.f);
    T.slot = -1;
    T.home = (int)lane;
    T.first = true;
    T.nosplit = false;
    T.lo = T.hi = 0u;
    T.sp = T.bot = 0;
    T.cur = 0;
    T.cur_t = 0.f;
    T.closest = 1e7f;
    T.best = -1;
    T.any = false;
    bool have = false, exhausted = false;
    rec.pending[lane] = 0;
    rec.nlog[lane] = 0;
    rec.bound[lane] = 1e7f;
    if (lane == 0) rec.overflow = 0;
    __syncwarp();
    for (;;) {
        // ---- refill from the queue: a lane whose previous ray still has tasks in flight keeps its record
        const unsigned int need = __ballot_sync(FULL, !have && rec.pending[lane] == 0);
        if (need != 0u && !exhausted) {
            const int n_take = min(__popc(need), grab);
            const int leader = __ffs(need) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(ticket, n_take);
            base = __shfl_sync(FULL, base, leader);
            const int rank = __popc(need & lt_mask);
            if (!have && rec.pending[lane] == 0 && rank < n_take) {
                const int s = base + rank;
                if (s < total) {
                    const float4 o = __ldg(ws.cray_o + s);
                    const float4 d = __ldg(ws.cray_d + s);
                    task_start_ray(T, make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), __float_as_int(o.w), (int)lane, false);
                    rec.pending[lane] = 1;
                    rec.nlog[lane] = 0;
                    rec.bound[lane] = 1e7f;
                    have = true;
                }
            }
            if (base + n_take >= total) exhausted = true;
        }
        const unsigned int busy = __ballot_sync(FULL, have);
        if (busy == 0u) {
            if (exhausted) break;
            continue;
        }
        if (SPLIT && busy != FULL) {
            // ---- steal: the k-th idle lane takes the bottom (last in visit order) entry of the k-th lane that can give one;
            // repeated while idle lanes and donors remain, so one long ray spreads over the warp in a single visit of this
            // block instead of one entry per round
            unsigned int busy_now = busy;
#pragma unroll 1
            for (int round = 0; round < MR_SPLIT_ROUNDS; ++round) {
                const unsigned int donors = __ballot_sync(FULL, have && task_can_donate(T));
                if (donors == 0u || busy_now == FULL) break;
                const unsigned int idle = ~busy_now;
                const int n_pairs = min(__popc(donors), __popc(idle));
                const bool robbed = have && task_can_donate(T) && __popc(donors & lt_mask) < n_pairs;
                int give_ref = 0;
                float give_t = 0.f;
                unsigned int mid = 0u, old_hi = 0u;
                if (robbed) {
                    give_ref = stack_ref[T.bot];
                    give_t = stack_t[T.bot];
                    ++T.bot;
                    mid = T.lo + ((T.hi - T.lo) >> 1);
                    old_hi = T.hi;
                    T.hi = mid;
                }
                const int my_rank = __popc(idle & lt_mask);
                const bool thief = !have && my_rank < n_pairs;
                int src = (int)lane;
                if (thief) {
                    unsigned int m = donors;
                    for (int j = 0; j < my_rank; ++j) m &= m - 1u;
                    src = __ffs(m) - 1;
                }
                const int g_ref = __shfl_sync(FULL, give_ref, src);
                const float g_t = __shfl_sync(FULL, give_t, src);
                const unsigned int g_mid = __shfl_sync(FULL, mid, src), g_hi = __shfl_sync(FULL, old_hi, src);
                const int g_slot = __shfl_sync(FULL, T.slot, src), g_home = __shfl_sync(FULL, T.home, src);
                const float g_bound = __shfl_sync(FULL, T.closest, src);
                const float ox = __shfl_sync(FULL, T.r.o.x, src), oy = __shfl_sync(FULL, T.r.o.y, src), oz = __shfl_sync(FULL, T.r.o.z, src);
                const float dx = __shfl_sync(FULL, T.r.d.x, src), dy = __shfl_sync(FULL, T.r.d.y, src), dz = __shfl_sync(FULL, T.r.d.z, src);
                const float ix = __shfl_sync(FULL, T.r.inv.x, src), iy = __shfl_sync(FULL, T.r.inv.y, src), iz = __shfl_sync(FULL, T.r.inv.z, src);
                if (thief) {
                    T.r.o = make_float3(ox, oy, oz);
                    T.r.d = make_float3(dx, dy, dz);
                    T.r.inv = make_float3(ix, iy, iz);
                    T.slot = g_slot;
                    T.home = g_home;
                    T.first = false;
                    T.nosplit = false;
                    T.lo = g_mid;
                    T.hi = g_hi;
                    T.sp = T.bot = 0;
                    T.cur = g_ref;
                    T.cur_t = g_t;
                    T.closest = g_bound; // stale by construction: >= every value the ray's closest distance takes later
                    T.best = -1;
                    T.any = false;
                    atomicAdd(&rec.pending[g_home], 1);
                    have = true;
                }
                busy_now = __ballot_sync(FULL, have);
            }
            __syncwarp();
        }
        bool retired = false;
        // shorter rounds while lanes are idle: a thief with a small subtree waits for the round to end before it can rob again
        const int steps = (SPLIT && busy != FULL) ? MR_TRACE_STEPS_SPLIT : MR_TRACE_STEPS;
#pragma unroll 1
        for (int it = 0; it < steps; ++it) {
            if (!have || retired) break;
            retired = task_step(bvh, T, stack_ref, stack_t, rec);
        }
        __syncwarp();
        // ---- retirement: the last task of a ray rebuilds the result from the first task's state and the thieves' logs
        if (__any_sync(FULL, retired)) {
            if (retired && T.first) {
                rec.fin_closest[T.home] = T.closest;
                rec.fin_best[T.home] = T.best;
                rec.fin_any[T.home] = T.any ? 1 : 0;
            }
            __syncwarp();
            bool last = false;
            if (retired) {
                last = atomicSub(&rec.pending[T.home], 1) == 1;
                have = false;
            }
            __syncwarp();
            if (last) {
                float closest;
                int best;
                bool any;
                if (task_replay(rec, T.home, closest, best, any)) {
                    closest_write_result(bvh, ws, T, closest, best, any);
                } else {
                    // the log lost entries: the ray starts over on this lane alone (the record stays the ray's)
                    const int home = T.home, slot = T.slot;
                    const float3 o = T.r.o, d = T.r.d;
                    task_start_ray(T, o, d, slot, home, true);
                    rec.pending[home] = 1;
                    rec.nlog[home] = 0;
                    rec.bound[home] = 1e7f;
                    have = true;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && rec.overflow) ws.counters[MR_CTR_ERROR] = 1;
}

// small queues: give every warp a few rays and let stealing spread each of them over the lanes
__device__ __forceinline__ int grab_limit(int total, int warps)
{
    return max(2, min(32, (total + warps - 1) / warps));
}

#define MR_TRACE_WARPS (MR_TRACE_BLOCK / 32)

template <bool TOP>
__global__ void __launch_bounds__(MR_TRACE_BLOCK) k_trace_any_persistent(BvhView bvh, Workspace ws)
{
    __shared__ Rec32 s_top[TOP ? MR_TOP_MAX * 4 : 1];
    const int n_top = TOP ? bvh.top->count : 0;
    if (TOP) {
        const int n4 = n_top * 8; // float4 words of the records in use
        const float4 *src = reinterpret_cast<const float4 *>(bvh.top->rec);
        float4 *dst = reinterpret_cast<float4 *>(s_top);
        for (int i = threadIdx.x; i < n4; i += MR_TRACE_BLOCK) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    trace_any_worker<TOP>(bvh, ws, grab_limit(ws.counters[MR_CTR_ANY_SIZE], gridDim.x * MR_TRACE_WARPS), n_top > 0 ? s_top : nullptr);
}
// SPLIT: idle lanes walk deferred subtrees of their warp's closest-hit rays (mr_split.cuh); the hit logs live in shared
// memory, 4.9 KB per warp.  The walker without splitting needs 0.9 KB per warp.
template <bool SPLIT>
__global__ void __launch_bounds__(MR_TRACE_BLOCK, 4) k_trace_closest_persistent(BvhView bvh, Workspace ws)
{
    typedef SplitRecT<SPLIT ? MR_SPLIT_LOG : 1> Rec;
    __shared__ Rec recs[MR_TRACE_WARPS];
    const int grab = SPLIT ? grab_limit(ws.counters[MR_CTR_CLOSEST_SIZE], gridDim.x * MR_TRACE_WARPS) : 32;
    trace_closest_worker<SPLIT>(bvh, ws, recs[threadIdx.x >> 5], grab);
}
// both queues in one launch (the two queues of process_path_tracing_divided_no_grad are independent,
// FinalShading.slang:745-977): the lower half of every block's warps walks the boolean rays, the upper half the
// closest-hit rays, so a block carries hit logs for half of its warps only
template <bool SPLIT>
__global__ void __launch_bounds__(MR_TRACE_BLOCK, 4) k_trace_mixed_persistent(BvhView bvh, Workspace ws)
{
    typedef SplitRecT<SPLIT ? MR_SPLIT_LOG : 1> Rec;
    __shared__ Rec recs[MR_TRACE_WARPS / 2];
    const int w = threadIdx.x >> 5;
    const int warps = gridDim.x * (MR_TRACE_WARPS / 2);
    if (w < MR_TRACE_WARPS / 2) {
        trace_any_worker<false>(bvh, ws, grab_limit(ws.counters[MR_CTR_ANY_SIZE], warps), nullptr);
    } else {
        const int grab = SPLIT ? grab_limit(ws.counters[MR_CTR_CLOSEST_SIZE], warps) : 32;
        trace_closest_worker<SPLIT>(bvh, ws, recs[w - MR_TRACE_WARPS / 2], grab);
    }
}
#endif

// Launch-shape tuning (mirres_set_tuning): values are per HOST THREAD, so two threads that drive different streams do
// not see each other's settings; 0 = library default.  Results never depend on them.
static thread_local int t_tune[MIRRES_TUNE_COUNT_] = {0};

int tuning_value(int key) { return key >= 0 && key < MIRRES_TUNE_COUNT_ ? t_tune[key] : 0; }

void queue_reset(const Workspace &ws, cudaStream_t st)
{
    zero_async(ws.counters + 1, 4 * sizeof(int), st);
}

#if !defined(MR_HOST_CHECK)
// what the tracers need to know about a device, looked up once per device (not per process)
struct DeviceInfo {
    std::atomic<int> ready;
    int sm_count, occ_any, occ_any_top, occ_closest, occ_closest_split, occ_mixed, occ_mixed_split;
};
#define MR_MAX_DEVICES 64
static DeviceInfo g_devices[MR_MAX_DEVICES];
static std::mutex g_devices_mutex;

static const DeviceInfo &device_info()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MR_MAX_DEVICES) dev = 0;
    DeviceInfo &d = g_devices[dev];
    if (!d.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(g_devices_mutex);
        if (!d.ready.load(std::memory_order_relaxed)) {
            cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
            if (d.sm_count <= 0) d.sm_count = 148;
            auto occ = [](auto kernel) {
                int o = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, MR_TRACE_BLOCK, 0);
                return o < 1 ? 1 : o;
            };
            d.occ_any = occ(k_trace_any_persistent<false>);
            d.occ_any_top = occ(k_trace_any_persistent<true>);
            d.occ_closest = occ(k_trace_closest_persistent<false>);
            d.occ_closest_split = occ(k_trace_closest_persistent<true>);
            d.occ_mixed = occ(k_trace_mixed_persistent<false>);
            d.occ_mixed_split = occ(k_trace_mixed_persistent<true>);
            d.ready.store(1, std::memory_order_release);
        }
    }
    return d;
}
#endif

int trace_queues(const BvhView &bvh, const Workspace &ws, bool any, bool closest, int sm_count, cudaStream_t st)
{
#if defined(MR_HOST_CHECK)
    (void)sm_count;
    QueueTraceParams p = {bvh, ws};
    int rc = 0;
    if (any) rc = foreach_item<QueueTraceParams, queue_any_item, 128>(p, ws.capacity * MR_MAX_RAYS_PER_PIXEL, st);
    if (!rc && closest) rc = foreach_item<QueueTraceParams, queue_closest_item, 128>(p, ws.capacity, st);
    return rc;
#else
    // persistent grids: at most what is resident at once -- and deliberately fewer blocks per SM than would fit: a
    // grid that fills the register file keeps every other stream's kernels off the SMs for its whole duration, while a
    // traversal warp is latency-bound and loses little from lower occupancy (measured on the C2 step: 3 / 2 blocks
    // per SM for boolean / closest-hit queues 6.4 ms, full occupancy 5 / 4 blocks 6.7 ms)
    (void)sm_count;
    const DeviceInfo &d = device_info();
    const bool split = t_tune[MIRRES_TUNE_CLOSEST_SPLIT] != 2; // 0 = default (on), 1 = on, 2 = off
    const bool top = t_tune[MIRRES_TUNE_ANY_TOP] == 1; // default off: measured, see profiles/README.md
    const int any_blocks = min(t_tune[MIRRES_TUNE_ANY_BLOCKS] > 0 ? t_tune[MIRRES_TUNE_ANY_BLOCKS] : 3, top ? d.occ_any_top : d.occ_any);
    const int closest_blocks = min(t_tune[MIRRES_TUNE_CLOSEST_BLOCKS] > 0 ? t_tune[MIRRES_TUNE_CLOSEST_BLOCKS] : 2,
                                   split ? d.occ_closest_split : d.occ_closest);
    const int mixed_blocks = min(t_tune[MIRRES_TUNE_MIXED_BLOCKS] > 0 ? t_tune[MIRRES_TUNE_MIXED_BLOCKS] : 4,
                                 split ? d.occ_mixed_split : d.occ_mixed);
    const int g_mixed = d.sm_count * mixed_blocks, g_any = d.sm_count * any_blocks, g_closest = d.sm_count * closest_blocks;
    if (any && closest) {
        if (split) k_trace_mixed_persistent<true><<<g_mixed, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else k_trace_mixed_persistent<false><<<g_mixed, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
    } else if (any) {
        if (top) k_trace_any_persistent<true><<<g_any, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else k_trace_any_persistent<false><<<g_any, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
    } else if (closest) {
        if (split) k_trace_closest_persistent<true><<<g_closest, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else k_trace_closest_persistent<false><<<g_closest, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
#endif
}

int device_sm_count()
{
#if defined(MR_HOST_CHECK)
    return 1;
#else
    return device_info().sm_count;
#endif
}


} // namespace mr

using namespace mr;

extern "C" {

int mirres_set_tuning(int key, int value)
{
    if (key < 0 || key >= MIRRES_TUNE_COUNT_) return MIRRES_ERR_SHAPE;
    t_tune[key] = value > 0 ? value : 0;
    return 0;
}

int mirres_get_tuning(int key)
{
    if (key < 0 || key >= MIRRES_TUNE_COUNT_) return MIRRES_ERR_SHAPE;
    return t_tune[key];
}

size_t mirres_workspace_bytes(int n_pixels) { return n_pixels < 1 ? 0 : workspace_carve(nullptr, n_pixels, nullptr); }

int mirres_workspace_prepare(const float *occ, int n_pixels, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!occ || !workspace) return MIRRES_ERR_NULL;
    if (n_pixels < 1) return MIRRES_ERR_SHAPE;
    if ((uintptr_t)workspace & 255) return MIRRES_ERR_ALIGN;
    if (workspace_bytes < workspace_carve(nullptr, n_pixels, nullptr)) return MIRRES_ERR_SCRATCH;
    Workspace ws;
    workspace_carve(&ws, n_pixels, (char *)workspace);
#if defined(MR_HOST_CHECK)
    (void)stream;
    int c = 0;
    for (int i = 0; i < n_pixels; ++i)
        if (!(occ[i] < 0.1f)) ws.active[c++] = i;
    ws.counters[0] = c;
    for (int k = 1; k < 8; ++k) ws.counters[k] = 0;
    for (int k = MR_CTR_ALIVE_SIZE; k <= MR_CTR_ALIVE_LAST; ++k) ws.counters[k] = 0;
    return 0;
#else
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (n_pixels + CP_TILE - 1) / CP_TILE;
    k_compact_count<<<nb, CP_BLOCK, 0, st>>>(occ, n_pixels, ws.block_counts);
    k_compact_scan<<<1, 1024, 0, st>>>(ws.block_counts, nb, ws.counters);
    k_compact_write<<<nb, CP_BLOCK, 0, st>>>(occ, n_pixels, ws.block_counts, ws.active);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
#endif
}

#if defined(MR_HOST_CHECK)
} // extern "C"
// Test-only (host-check flavour, never in the shipped library): the split closest-hit walk of mr_split.cuh -- the same
// task_step / task_replay code the CUDA worker runs -- under a RANDOMISED scheduler: `lanes` virtual lanes share one
// record set; every round each busy lane advances a random number of visits, idle lanes rob the bottom stack entry of
// lanes that can give one (same pairing as the CUDA worker), rays retire in whatever order the schedule produces.  `cap`
// picks the log capacity (1 forces overflows and restarts).  Outputs in the layout of mirres_trace_closest.
template <int CAP>
static void closest_split_simulate(const BvhView &bvh, const float *org, const float *dir, int n, int lanes, unsigned int seed,
                                   int *hit, float *t, float *pos, float *normal, int *prim, int *stats)
{
    SplitRecT<CAP> rec;
    CTask T[MR_SPLIT_LANES];
    static thread_local int stack_ref[MR_SPLIT_LANES][MR_STACK + MR_STACK_SLACK];
    static thread_local float stack_t[MR_SPLIT_LANES][MR_STACK + MR_STACK_SLACK];
    bool have[MR_SPLIT_LANES];
    unsigned int rng = seed * 2654435761u + 12345u;
    auto rnd = [&]() { rng = 1664525u * rng + 1013904223u; return rng >> 8; };
    rec.overflow = 0;
    for (int l = 0; l < lanes; ++l) { have[l] = false; rec.pending[l] = 0; rec.nlog[l] = 0; rec.bound[l] = 1e7f; }
    int next_ray = 0;
    long steals = 0, restarts = 0, logged = 0;
    for (;;) {
        for (int l = 0; l < lanes; ++l) {
            if (!have[l] && rec.pending[l] == 0 && next_ray < n && (rnd() & 3u) != 0u) {
                const int s = next_ray++;
                task_start_ray(T[l], load3(org, (size_t)s), load3(dir, (size_t)s), s, l, false);
                rec.pending[l] = 1; rec.nlog[l] = 0; rec.bound[l] = 1e7f;
                have[l] = true;
            }
        }
        bool busy = false;
        for (int l = 0; l < lanes; ++l) busy = busy || have[l];
        if (!busy) { if (next_ray >= n) break; continue; }
        // steal: k-th idle lane <- k-th donor
        int donors[MR_SPLIT_LANES], nd = 0, idle[MR_SPLIT_LANES], ni = 0;
        for (int l = 0; l < lanes; ++l) {
            if (have[l] && task_can_donate(T[l])) donors[nd++] = l;
            if (!have[l]) idle[ni++] = l;
        }
        const int pairs = nd < ni ? nd : ni;
        for (int k = 0; k < pairs; ++k) {
            if ((rnd() & 1u) == 0u) continue; // not every opportunity is taken
            CTask &D = T[donors[k]], &N = T[idle[k]];
            const int dl = donors[k];
            const unsigned int mid = D.lo + ((D.hi - D.lo) >> 1), old_hi = D.hi;
            N.r = D.r; N.slot = D.slot; N.home = D.home; N.first = false; N.nosplit = false;
            N.lo = mid; N.hi = old_hi; N.sp = N.bot = 0;
            N.cur = stack_ref[dl][D.bot]; N.cur_t = stack_t[dl][D.bot]; ++D.bot;
            D.hi = mid;
            N.closest = D.closest; N.best = -1; N.any = false;
            rec.pending[N.home] += 1;
            have[idle[k]] = true;
            ++steals;
        }
        bool retired[MR_SPLIT_LANES];
        for (int l = 0; l < lanes; ++l) {
            retired[l] = false;
            if (!have[l]) continue;
            const int steps = (int)(rnd() % 7u);
            const int before = rec.nlog[T[l].home];
            for (int it = 0; it < steps && !retired[l]; ++it) retired[l] = task_step(bvh, T[l], stack_ref[l], stack_t[l], rec);
            logged += rec.nlog[T[l].home] - before;
        }
        for (int l = 0; l < lanes; ++l) {
            if (retired[l] && T[l].first) {
                rec.fin_closest[T[l].home] = T[l].closest; rec.fin_best[T[l].home] = T[l].best; rec.fin_any[T[l].home] = T[l].any ? 1 : 0;
            }
        }
        for (int l = 0; l < lanes; ++l) {
            if (!retired[l]) continue;
            have[l] = false;
            if (--rec.pending[T[l].home] != 0) continue;
            float closest; int best; bool any;
            if (task_replay(rec, T[l].home, closest, best, any)) {
                const size_t i = (size_t)T[l].slot;
                float3 p = f3(0.f), nn = f3(1.0f);
                int pr = -1;
                if (any) { p = T[l].r.o + closest * T[l].r.d; closest_finish(bvh, T[l].r, best, nn, pr, nullptr); }
                hit[i] = any ? 1 : 0;
                if (t) t[i] = any ? closest : 0.f;
                if (pos) store3(pos, i, p);
                if (normal) store3(normal, i, nn);
                if (prim) prim[i] = pr;
            } else {
                const int home = T[l].home, slot = T[l].slot;
                task_start_ray(T[l], load3(org, (size_t)slot), load3(dir, (size_t)slot), slot, home, true);
                rec.pending[home] = 1; rec.nlog[home] = 0; rec.bound[home] = 1e7f;
                have[l] = true;
                ++restarts;
            }
        }
    }
    if (stats) { stats[0] = (int)steals; stats[1] = (int)restarts; stats[2] = (int)logged; }
}

extern "C" int mirres_test_closest_split(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir,
                                         int n, int lanes, int cap, unsigned int seed, int *hit, float *t, float *pos,
                                         float *normal, int *prim, int *stats)
{
    if (!packed_nodes || !packed_tris || !org || !dir || !hit) return MIRRES_ERR_NULL;
    if (n < 0 || lanes < 1 || lanes > MR_SPLIT_LANES) return MIRRES_ERR_SHAPE;
    BvhView bvh = bvh_view(packed_nodes, packed_tris);
    if (cap == 1) closest_split_simulate<1>(bvh, org, dir, n, lanes, seed, hit, t, pos, normal, prim, stats);
    else if (cap == 2) closest_split_simulate<2>(bvh, org, dir, n, lanes, seed, hit, t, pos, normal, prim, stats);
    else closest_split_simulate<MR_SPLIT_LOG>(bvh, org, dir, n, lanes, seed, hit, t, pos, normal, prim, stats);
    return 0;
}
extern "C" {
#endif

// test / diagnostics helper: copies out the active list header is not needed -- the list lives in the workspace at a
// fixed offset: counters (64 ints, 256 B) then active[n_pixels].

} // extern "C"
