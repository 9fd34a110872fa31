// mirres-b200 wavefront machinery: compacted foreground pixels, ray queues, queue tracers.
//
// The reference casts rays from inside its per-pixel kernels (one thread = one pixel, up to ten sequential rays in
// process_SpatialResampling_, nerf/ScreenSpaceReSTIR/SpatialResampling.slang:258-284).  On a 21 %-covered 800x800
// frame ncu measured 4.8 of 32 lanes active per instruction for that shape.  Here every ray-casting entry point is a
// wavefront instead:
//     gen kernel      (one thread per ACTIVE pixel)  -> ray queue, fixed R slots per active pixel, SoA float4 o / d
//     queue tracer    (any-hit: persistent warps that refill idle lanes from the queue; closest-hit: one thread/slot)
//     resolve kernel  (one thread per active pixel)  -> consumes hit flags / hit records
// Arithmetic per pixel and per ray is unchanged, so results stay bit-identical to the per-pixel formulation.
//
// Workspace (caller-allocated, mirres_workspace_bytes(N)): active pixel list, queue, results, per-pixel scratch.
#pragma once
#include "mr_bvh.cuh"

namespace mr {

#define MR_MAX_RAYS_PER_PIXEL 10
#define MR_PX_SCRATCH_FLOATS 32

struct Workspace {
    int *counters;     // [0] number of active pixels, [1] work counter of the running tracer, [2..15] spare
    int *active;       // [N] pixel indices with occ >= 0.1, ascending
    int *block_counts; // [ceil(N/1024) + 1] compaction scratch
    float4 *ray_o;     // [N * 10] origin.xyz, w = 1 valid / 0 empty slot
    float4 *ray_d;     // [N * 10] direction.xyz (normalised again by the tracer, as bvh_hit does)
    unsigned int *hit; // [N * 10] any-hit result per slot
    float4 *chit;      // [N * 2]  closest-hit record per active pixel: (pos.xyz, found) (normal.xyz, t)
    float *px;         // [N * MR_PX_SCRATCH_FLOATS] per-active-pixel state carried from gen to resolve
    float *stop_in;    // [N] stop flag of every pixel as it was on entry to a bounce kernel
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
    p = take((size_t)N * 2 * sizeof(float4)); if (w) w->chit = (float4 *)p;
    p = take((size_t)N * MR_PX_SCRATCH_FLOATS * sizeof(float)); if (w) w->px = (float *)p;
    p = take((size_t)N * sizeof(float)); if (w) w->stop_in = (float *)p;
    if (w) w->capacity = N;
    return off;
}

MR_DEV void queue_ray(const Workspace &w, size_t slot, float3 o, float3 d)
{
    w.ray_o[slot] = make_float4(o.x, o.y, o.z, 1.0f);
    w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.0f);
}
MR_DEV void queue_empty(const Workspace &w, size_t slot) { w.ray_o[slot] = make_float4(0.f, 0.f, 0.f, 0.0f); }

// ---- one-thread-per-slot tracers (closest-hit; also the host-check flavour of any-hit) ---------------------------
struct QueueTraceParams {
    BvhView bvh;
    Workspace ws;
    int rays_per_item;
};
MR_DEV void queue_any_item(const QueueTraceParams &p, int slot)
{
    if (slot >= p.ws.counters[0] * p.rays_per_item) return;
    float4 o = p.ws.ray_o[slot];
    if (o.w == 0.0f) return;
    float4 d = p.ws.ray_d[slot];
    p.ws.hit[slot] = any_hit<false>(p.bvh, make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), nullptr) ? 1u : 0u;
}
// closest-hit queue: one slot per active pixel, record in chit
MR_DEV void queue_closest_item(const QueueTraceParams &p, int slot)
{
    if (slot >= p.ws.counters[0]) return;
    float4 o = p.ws.ray_o[slot];
    if (o.w == 0.0f) return;
    float4 d = p.ws.ray_d[slot];
    Hit h;
    h.t = 0.f;
    h.pos = f3(0.f);
    h.normal = f3(1.f);
    bool found = closest_hit<false>(p.bvh, make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), h, nullptr);
    p.ws.chit[2 * (size_t)slot] = make_float4(h.pos.x, h.pos.y, h.pos.z, found ? 1.0f : 0.0f);
    p.ws.chit[2 * (size_t)slot + 1] = make_float4(h.normal.x, h.normal.y, h.normal.z, h.t);
}

// queue tracers and device query: defined once, in wave.cu
int trace_queue_any(const BvhView &bvh, const Workspace &ws, int rays_per_item, int sm_count, cudaStream_t st);
int trace_queue_closest(const BvhView &bvh, const Workspace &ws, int sm_count, cudaStream_t st);
int device_sm_count();

} // namespace mr
