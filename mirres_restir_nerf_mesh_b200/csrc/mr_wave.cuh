// mirres-b200 wavefront machinery: compacted foreground pixels, dense ray queues, persistent queue tracers.
//
// The reference casts rays from inside its per-pixel kernels (one thread = one pixel, up to ten sequential rays in
// process_SpatialResampling_, nerf/ScreenSpaceReSTIR/SpatialResampling.slang:258-284).  On a 21 %-covered 800x800
// frame ncu measured 4.8 of 32 lanes active per instruction for that shape.  Here every ray-casting entry point is a
// wavefront instead:
//     gen kernel      (one thread per ACTIVE pixel)  -> appends its rays to a DENSE queue (warp-aggregated ticket);
//                                                       every ray carries the id of the result slot it belongs to
//     queue tracer    persistent warps: idle lanes are refilled from the queue; once the queue runs dry, idle lanes
//                     of an any-hit warp STEAL deferred subtrees from the traversal stacks of busy lanes (a boolean
//                     query is an OR over subtrees, so the split cannot change the result); closest-hit rays keep the
//                     reference visit order, one lane per ray
//     resolve kernel  (one thread per active pixel)  -> consumes hit flags / hit records by slot
// Arithmetic per pixel and per ray is unchanged, so results stay bit-identical to the per-pixel formulation.
//
// Workspace (caller-allocated, mirres_workspace_bytes(N)): active pixel list, queues, results, per-pixel scratch.
#pragma once
#include "mr_bvh.cuh"

namespace mr {

#define MR_MAX_RAYS_PER_PIXEL 10
#define MR_PX_SCRATCH_FLOATS 32

// per-slot state of a boolean ray
#define MR_HIT_MISS 0u
#define MR_HIT_HIT 1u
#define MR_HIT_NONE 2u // no ray was queued for this slot

// counters[]: [0] active pixels  [1] any-hit ticket  [2] any-hit queue size  [3] closest ticket  [4] closest queue size
#define MR_CTR_ACTIVE 0
#define MR_CTR_ANY_TICKET 1
#define MR_CTR_ANY_SIZE 2
#define MR_CTR_CLOSEST_TICKET 3
#define MR_CTR_CLOSEST_SIZE 4
// [8] frame offset: added to the frame_index argument of every kernel that takes a workspace.  Written by the caller
// (never by the library), so a CUDA graph whose launches have their frame indices baked in can be replayed with fresh
// random streams (graphed.CapturedStep.set_frame_offset).  Zero = the plain reference behaviour.
#define MR_CTR_FRAME_OFFSET 8
// [9] row offset: added to a pixel's row when its random stream is seeded (Seed_Generator(pixel, frame), random.slang:2-39).
// Written by the caller like the frame offset.  A rank that renders a band of a larger frame passes the band's rows as
// a frame of its own (the [N, k] maps are row-major, so a band is a contiguous slice of every tensor) and sets this word
// to the band's first row: every pixel then draws the random numbers it draws in the full frame.  Zero = plain behaviour.
#define MR_CTR_ROW_OFFSET 9
// [10], [11] band of rows [lo, hi) of the frame handed in that the SPATIAL pass resamples; pixels of the list outside it
// only publish their sample for the band's pixels to reuse (row-band rendering: the pass reads neighbours up to 30 rows
// beyond the band).  hi <= lo (the zero-filled default) = every listed pixel.  Written by the caller.
#define MR_CTR_BAND_LO 10
#define MR_CTR_BAND_HI 11
// [12] error word: set to 1 by a queue tracer that had to drop a traversal-stack entry (MR_STACK); never cleared by the
// library (the caller zero-fills the workspace once and may inspect / clear the word whenever it synchronises)
#define MR_CTR_ERROR 12
// [16], [17] sizes of the two lists of paths that are still alive (the path kernels ping-pong between them), [18], [19]
// the signatures of the calls those lists were written for (see bounce_item in shade.cu; a kernel reads one word and
// writes the other); cleared by mirres_workspace_prepare
#define MR_CTR_ALIVE_SIZE 16
#define MR_CTR_ALIVE_SIG 18
#define MR_CTR_ALIVE_LAST 19

struct Workspace {
    int *counters;     // [64] see MR_CTR_*
    int *active;       // [N] pixel indices with occ >= 0.1, ascending
    int *block_counts; // [ceil(N/1024) + 1] compaction scratch
    float4 *ray_o;     // [N * 10] dense any-hit queue: origin.xyz, w = result slot (int bits)
    float4 *ray_d;     // [N * 10] direction.xyz (normalised again by the tracer, as bvh_hit does)
    unsigned int *hit; // [N * 10] per SLOT: MR_HIT_*
    float4 *cray_o;    // [N] dense closest-hit queue: origin.xyz, w = result slot (int bits)
    float4 *cray_d;    // [N]
    float4 *chit;      // [N * 3] per SLOT: (pos.xyz, -1 no ray / 0 miss / 1 hit) (normal.xyz, t) (prim bits, u, v, 0)
    float *px;         // [N * MR_PX_SCRATCH_FLOATS] per-active-pixel state carried from gen to resolve
    float *stop_in;    // [N] stop flag of every pixel as it was on entry to a bounce kernel
    float4 *lcache;    // [N * 2] per PIXEL: (emitted radiance, own target density) (direction, 0) of the pixel's reservoir sample (spatial pass)
    int *alive[2];     // [N] each: active-list positions of the paths that continue past a path vertex (unordered)
    int capacity;      // N
};

static inline size_t ws_align(size_t x) { return (x + 255) & ~(size_t)255; }
static inline size_t workspace_carve(Workspace *w, int N, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += ws_align(bytes); return base ? base + o : (char *)0; };
    char *p;
    p = take(64 * sizeof(int)); if (w) w->counters = (int *)p;
    p = take((size_t)N * sizeof(int)); if (w) w->active = (int *)p;
    p = take(((size_t)(N + 1023) / 1024 + 1) * sizeof(int)); if (w) w->block_counts = (int *)p;
    p = take((size_t)N * MR_MAX_RAYS_PER_PIXEL * sizeof(float4)); if (w) w->ray_o = (float4 *)p;
    p = take((size_t)N * MR_MAX_RAYS_PER_PIXEL * sizeof(float4)); if (w) w->ray_d = (float4 *)p;
    p = take((size_t)N * MR_MAX_RAYS_PER_PIXEL * sizeof(unsigned int)); if (w) w->hit = (unsigned int *)p;
    p = take((size_t)N * sizeof(float4)); if (w) w->cray_o = (float4 *)p;
    p = take((size_t)N * sizeof(float4)); if (w) w->cray_d = (float4 *)p;
    p = take((size_t)N * 3 * sizeof(float4)); if (w) w->chit = (float4 *)p;
    p = take((size_t)N * MR_PX_SCRATCH_FLOATS * sizeof(float)); if (w) w->px = (float *)p;
    p = take((size_t)N * sizeof(float)); if (w) w->stop_in = (float *)p;
    p = take((size_t)N * 2 * sizeof(float4)); if (w) w->lcache = (float4 *)p;
    p = take((size_t)N * sizeof(int)); if (w) w->alive[0] = (int *)p;
    p = take((size_t)N * sizeof(int)); if (w) w->alive[1] = (int *)p;
    if (w) w->capacity = N;
    return off;
}

MR_DEV bool in_band(const Workspace &w, unsigned int py)
{
    const int lo = w.counters[MR_CTR_BAND_LO], hi = w.counters[MR_CTR_BAND_HI];
    return hi <= lo || ((int)py >= lo && (int)py < hi);
}
MR_DEV unsigned int row_of(const Workspace &w, unsigned int py) { return py + (unsigned int)w.counters[MR_CTR_ROW_OFFSET]; }
MR_DEV unsigned int frame_of(const Workspace &w, unsigned int frame_index) { return frame_index + (unsigned int)w.counters[MR_CTR_FRAME_OFFSET]; }

// one ticket of a queue; lanes of a warp that arrive together share one atomic
MR_DEV int queue_alloc(int *ctr)
{
#if defined(__CUDA_ARCH__)
    const unsigned int m = __activemask();
    const unsigned int lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
#else
    int q;
#pragma omp atomic capture
    q = (*ctr)++;
    return q;
#endif
}

MR_DEV void queue_ray(const Workspace &w, size_t slot, float3 o, float3 d)
{
    w.hit[slot] = MR_HIT_MISS;
    const int q = queue_alloc(w.counters + MR_CTR_ANY_SIZE);
    w.ray_o[q] = make_float4(o.x, o.y, o.z, bits_float((int)slot));
    w.ray_d[q] = make_float4(d.x, d.y, d.z, 0.0f);
}
MR_DEV void queue_empty(const Workspace &w, size_t slot) { w.hit[slot] = MR_HIT_NONE; }
// entry q of the boolean-ray queue, for callers that reserve a block of tickets themselves
MR_DEV void queue_ray_at(const Workspace &w, int q, size_t slot, float3 o, float3 d)
{
    w.hit[slot] = MR_HIT_MISS;
    w.ray_o[q] = make_float4(o.x, o.y, o.z, bits_float((int)slot));
    w.ray_d[q] = make_float4(d.x, d.y, d.z, 0.0f);
}

MR_DEV void queue_closest_ray(const Workspace &w, size_t slot, float3 o, float3 d)
{
    const int q = queue_alloc(w.counters + MR_CTR_CLOSEST_SIZE);
    w.cray_o[q] = make_float4(o.x, o.y, o.z, bits_float((int)slot));
    w.cray_d[q] = make_float4(d.x, d.y, d.z, 0.0f);
}
MR_DEV void queue_closest_empty(const Workspace &w, size_t slot) { w.chit[3 * slot] = make_float4(0.f, 0.f, 0.f, -1.0f); }

// ---- one-thread-per-entry tracers: the host-check flavour of the queue tracers ------------------------------------
struct QueueTraceParams {
    BvhView bvh;
    Workspace ws;
};
MR_DEV void queue_any_item(const QueueTraceParams &p, int q)
{
    if (q >= p.ws.counters[MR_CTR_ANY_SIZE]) return;
    const float4 o = p.ws.ray_o[q], d = p.ws.ray_d[q];
    if (any_hit<false>(p.bvh, make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), nullptr)) p.ws.hit[float_bits(o.w)] = MR_HIT_HIT;
}
MR_DEV void queue_closest_item(const QueueTraceParams &p, int q)
{
    if (q >= p.ws.counters[MR_CTR_CLOSEST_SIZE]) return;
    const float4 o = p.ws.cray_o[q], d = p.ws.cray_d[q];
    const size_t slot = (size_t)float_bits(o.w);
    Hit h;
    h.t = 0.f;
    h.pos = f3(0.f);
    h.normal = f3(1.f);
    h.prim = -1;
    h.bary[0] = h.bary[1] = 0.f;
    bool found = closest_hit<false>(p.bvh, make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), h, nullptr);
    p.ws.chit[3 * slot] = make_float4(h.pos.x, h.pos.y, h.pos.z, found ? 1.0f : 0.0f);
    p.ws.chit[3 * slot + 1] = make_float4(h.normal.x, h.normal.y, h.normal.z, h.t);
    p.ws.chit[3 * slot + 2] = make_float4(bits_float(found ? h.prim : -1), h.bary[0], h.bary[1], 0.f);
}

// queue tracers and device query: defined once, in wave.cu.  queue_reset must be enqueued before the gen kernel of
// every ray-casting entry point; trace_queues walks the any-hit queue, the closest-hit queue, or both concurrently.
void queue_reset(const Workspace &ws, cudaStream_t st);
int trace_queues(const BvhView &bvh, const Workspace &ws, bool any, bool closest, int sm_count, cudaStream_t st);
int device_sm_count();

} // namespace mr
