// mirres-b200: workspace preparation = ordered compaction of the foreground pixels (occ >= 0.1).
//
// Every ray-casting stage of the reference early-outs on `occ_map[pixel] < 0.1` (e.g. InitialResampling.slang:166,
// SpatialResampling.slang:192, FinalShading.slang:166,760).  The list of pixels that survive that test is built
// once per frame here and shared by all wavefront stages (mr_wave.cuh); ascending pixel order keeps neighbouring
// lanes on neighbouring surface points.
#include <stdlib.h>
#include "mr_wave.cuh"
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
#define MR_TRACE_STEPS 6
#define MR_TRACE_STEPS_SHARED 4

// Any-hit.  When a warp cannot refill all of its idle lanes (queue dry, or -- for small queues -- the per-warp grab limit
// is reached) the idle lanes take over the OLDEST deferred subtree of busy lanes.  bvh_hit's boolean result is the OR over
// all leaf tests (mr_bvh.cuh), so walking the subtrees of one ray on several lanes returns the same flag while the
// longest ray of a launch stops being a serial chain of ~10^3 dependent L2 loads.
__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <bool PREFETCH>
__device__ __forceinline__ void trace_any_worker(const BvhView &bvh, const Workspace &ws, int grab)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[MR_CTR_ANY_SIZE];
    int *ticket = ws.counters + MR_CTR_ANY_TICKET;
    int stack[MR_STACK];
    int sp = 0, bot = 0;   // live entries: [bot, sp)
    int cur = 0;           // node reference being processed: >= 0 internal, < 0 leaf
    int slot = -1;
    bool have = false;      // this lane owns a live (ray, subtree)
    bool found = false;     // this lane has just resolved its ray with a hit
    bool exhausted = false; // warp-uniform: the queue has been drained
    bool shared = false;    // warp-uniform: some ray of this warp is being walked by more than one lane
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
                    cur = 0;
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
            const Rec32 *rec = ref_address(bvh, cur);
            const Rec32 e0 = load_rec(rec), e1 = load_rec(rec + 1);
            if (cur >= 0) {
                Rec32 e2, e3;
                load_tail(rec, ref_missing(cur), e2, e3); // trailing sectors only when the reference says they are in use
                WideHit w;
                wide_slabs(r, e0, e1, e2, e3, w);
                int next = 0;
                bool got = false;
#pragma unroll
                for (int k = 3; k >= 0; --k) {
                    if (fminf(1e7f, w.tf[k]) > w.tn[k]) {
                        if (got) {
                            if (PREFETCH) prefetch_l1(ref_address(bvh, next));
                            stack[sp++] = next;
                        }
                        next = w.ref[k];
                        got = true;
                    }
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
}

// Closest-hit: same refill scheme, reference visit order (see closest_hit in mr_bvh.cuh), one lane per ray.
template <bool PREFETCH>
__device__ __forceinline__ void trace_closest_worker(const BvhView &bvh, const Workspace &ws)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[MR_CTR_CLOSEST_SIZE];
    int *ticket = ws.counters + MR_CTR_CLOSEST_TICKET;
    int stack_ref[MR_STACK];
    float stack_t[MR_STACK];
    int sp = 0, cur = 0, slot = -1, best_slot = -1;
    float closest = 1e7f;
    bool any = false, have = false, exhausted = false;
    Ray r;
    r.o = r.d = r.inv = f3(0.f);
    for (;;) {
        const unsigned int need = __ballot_sync(FULL, !have);
        if (need != 0u && !exhausted) {
            const int n_need = __popc(need);
            const int leader = __ffs(need) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(ticket, n_need);
            base = __shfl_sync(FULL, base, leader);
            if (!have) {
                const int s = base + __popc(need & lt_mask);
                if (s < total) {
                    const float4 o = __ldg(ws.cray_o + s);
                    const float4 d = __ldg(ws.cray_d + s);
                    r = make_ray(make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z));
                    slot = __float_as_int(o.w);
                    sp = 0;
                    cur = 0;
                    closest = 1e7f;
                    any = false;
                    best_slot = -1;
                    have = true;
                }
            }
            if (base + n_need >= total) exhausted = true;
        }
        if (!__any_sync(FULL, have)) {
            if (exhausted) break;
            continue;
        }
#pragma unroll 1
        for (int it = 0; it < MR_TRACE_STEPS; ++it) {
            if (!have) break;
            bool pop = false;
            const Rec32 *rec = ref_address(bvh, cur);
            const Rec32 e0 = load_rec(rec), e1 = load_rec(rec + 1);
            if (cur >= 0) {
                // all four sectors, unconditionally: this walk is bound by the latency of one ray's dependent fetches, and
                // making two of them conditional on the reference's entry count lengthened every step (+10 % measured)
                const Rec32 e2 = load_rec(rec + 2), e3 = load_rec(rec + 3);
                WideHit w;
                wide_slabs(r, e0, e1, e2, e3, w);
                int next = 0;
                float next_t = 0.f;
                bool got = false;
#pragma unroll
                for (int k = 3; k >= 0; --k) {
                    if (fminf(closest, w.tf[k]) > w.tn[k]) {
                        if (got) {
                            if (PREFETCH) prefetch_l1(ref_address(bvh, next));
                            stack_ref[sp] = next;
                            stack_t[sp] = next_t;
                            ++sp;
                        }
                        next = w.ref[k];
                        next_t = w.tn[k];
                        got = true;
                    }
                }
                if (got) cur = next;
                else pop = true;
            } else {
                const int leaf = ~cur;
                float t, u, v;
                if (tri_test(r, e0, e1, t, u, v)) {
                    if (t <= closest) best_slot = leaf;
                    closest = fminf(t, closest);
                    any = true;
                }
                pop = true;
            }
            if (pop) {
                bool found = false;
                while (sp > 0) {
                    --sp;
                    if (closest > stack_t[sp]) {
                        cur = stack_ref[sp];
                        found = true;
                        break;
                    }
                }
                if (!found) {
                    float3 pos = f3(0.f), n = f3(1.0f);
                    int prim = -1;
                    float bary[2] = {0.f, 0.f};
                    if (any) {
                        pos = r.o + closest * r.d;
                        closest_finish(bvh, r, best_slot, n, prim, bary);
                    }
                    ws.chit[3 * (size_t)slot] = make_float4(pos.x, pos.y, pos.z, any ? 1.0f : 0.0f);
                    ws.chit[3 * (size_t)slot + 1] = make_float4(n.x, n.y, n.z, any ? closest : 0.f);
                    ws.chit[3 * (size_t)slot + 2] = make_float4(__int_as_float(prim), bary[0], bary[1], 0.f);
                    have = false;
                }
            }
        }
        __syncwarp();
    }
}

// small queues: give every warp a few rays and let stealing spread each of them over the lanes
__device__ __forceinline__ int any_grab_limit(const Workspace &ws, int warps)
{
    const int total = ws.counters[MR_CTR_ANY_SIZE];
    return max(2, min(32, (total + warps - 1) / warps));
}

template <bool PREFETCH>
__global__ void __launch_bounds__(MR_TRACE_BLOCK) k_trace_any_persistent(BvhView bvh, Workspace ws)
{
    trace_any_worker<PREFETCH>(bvh, ws, any_grab_limit(ws, gridDim.x * (MR_TRACE_BLOCK / 32)));
}
template <bool PREFETCH>
__global__ void __launch_bounds__(MR_TRACE_BLOCK, 4) k_trace_closest_persistent(BvhView bvh, Workspace ws)
{
    trace_closest_worker<PREFETCH>(bvh, ws);
}
// both queues in one launch: even blocks walk the boolean rays, odd blocks the closest-hit rays (the two queues of
// process_path_tracing_divided_no_grad are independent, FinalShading.slang:745-977)
template <bool PREFETCH>
__global__ void __launch_bounds__(MR_TRACE_BLOCK, 4) k_trace_mixed_persistent(BvhView bvh, Workspace ws)
{
    if ((blockIdx.x & 1u) == 0u) trace_any_worker<PREFETCH>(bvh, ws, any_grab_limit(ws, (gridDim.x / 2) * (MR_TRACE_BLOCK / 32)));
    else trace_closest_worker<PREFETCH>(bvh, ws);
}
#endif

// tuning overrides of the persistent grids (blocks per SM), 0 = default; see mirres_set_tuning
static int g_tune_any_blocks = 0, g_tune_closest_blocks = 0;

void queue_reset(const Workspace &ws, cudaStream_t st)
{
    zero_async(ws.counters + 1, 4 * sizeof(int), st);
}

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
    static int occ_any = 0, occ_closest = 0, occ_mixed = 0, prefetch = 0;
    if (!occ_any) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_any, k_trace_any_persistent<false>, MR_TRACE_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_closest, k_trace_closest_persistent<false>, MR_TRACE_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_mixed, k_trace_mixed_persistent<false>, MR_TRACE_BLOCK, 0);
        if (occ_any < 1) occ_any = 4;
        if (occ_closest < 1) occ_closest = 4;
        if (occ_mixed < 2) occ_mixed = 4;
        const char *e = getenv("MIRRES_PREFETCH");
        prefetch = e ? atoi(e) : 0;
        if (occ_any > 3) occ_any = 3;
        if (occ_closest > 2) occ_closest = 2;
        if (occ_mixed > 4) occ_mixed = 4;
        const char *c = getenv("MIRRES_CLOSEST_BLOCKS"); // tuning overrides
        if (c && atoi(c) > 0) { occ_closest = atoi(c); occ_mixed = 2 * atoi(c); }
        const char *a = getenv("MIRRES_ANY_BLOCKS");
        if (a && atoi(a) > 0) occ_any = atoi(a);
    }
    // per-call override (mirres_set_tuning): the host raises the grid of the launches on its critical chain, which run on
    // a high-priority stream, and leaves the background chains at the small default
    static int occ_any_max = 0, occ_closest_max = 0;
    if (!occ_any_max) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_any_max, k_trace_any_persistent<false>, MR_TRACE_BLOCK, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_closest_max, k_trace_closest_persistent<false>, MR_TRACE_BLOCK, 0);
        if (occ_any_max < 1) occ_any_max = 4;
        if (occ_closest_max < 1) occ_closest_max = 4;
    }
    const int use_any = g_tune_any_blocks > 0 ? min(g_tune_any_blocks, occ_any_max) : occ_any;
    const int use_closest = g_tune_closest_blocks > 0 ? min(g_tune_closest_blocks, occ_closest_max) : occ_closest;
    const int g_mixed = sm_count * (occ_mixed & ~1), g_any = sm_count * use_any, g_closest = sm_count * use_closest;
    if (prefetch) {
        if (any && closest) k_trace_mixed_persistent<true><<<g_mixed, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else if (any) k_trace_any_persistent<true><<<g_any, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else if (closest) k_trace_closest_persistent<true><<<g_closest, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
    } else {
        if (any && closest) k_trace_mixed_persistent<false><<<g_mixed, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else if (any) k_trace_any_persistent<false><<<g_any, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
        else if (closest) k_trace_closest_persistent<false><<<g_closest, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
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
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
#endif
}


} // namespace mr

using namespace mr;

extern "C" {

int mirres_set_tuning(int key, int value)
{
    if (key == MIRRES_TUNE_ANY_BLOCKS) g_tune_any_blocks = value > 0 ? value : 0;
    else if (key == MIRRES_TUNE_CLOSEST_BLOCKS) g_tune_closest_blocks = value > 0 ? value : 0;
    else return MIRRES_ERR_SHAPE;
    return 0;
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

// test / diagnostics helper: copies out the active list header is not needed -- the list lives in the workspace at a
// fixed offset: counters (64 ints, 256 B) then active[n_pixels].

} // extern "C"
