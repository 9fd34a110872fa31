// mirres-b200: workspace preparation = ordered compaction of the foreground pixels (occ >= 0.1).
//
// Every ray-casting stage of the reference early-outs on `occ_map[pixel] < 0.1` (e.g. InitialResampling.slang:166,
// SpatialResampling.slang:192, FinalShading.slang:166,760).  The list of pixels that survive that test is built
// once per frame here and shared by all wavefront stages (mr_wave.cuh); ascending pixel order keeps neighbouring
// lanes on neighbouring surface points.
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
// ---- persistent any-hit tracer ----------------------------------------------------------------------------------------
// Warps pull ray slots from the queue with one atomic per refill; a lane whose ray terminates (first hit, or stack
// empty) is refilled at the next check point, so the SIMD lanes stay busy although path lengths vary by > 10x.
#define MR_TRACE_BLOCK 256
#define MR_TRACE_STEPS 6

__global__ void __launch_bounds__(MR_TRACE_BLOCK) k_trace_any_persistent(BvhView bvh, Workspace ws, int rays_per_item)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[0] * rays_per_item;
    int *work = ws.counters + 1;
    int stack[MR_STACK];
    int sp = 0;
    int cur = 0;           // node reference being processed: >= 0 internal, < 0 leaf
    int slot = -1;
    bool have = false;     // this lane owns a live ray
    bool exhausted = false; // warp-uniform: the queue has been drained
    Ray r;
    r.o = r.d = r.inv = f3(0.f);
    for (;;) {
        const unsigned int need = __ballot_sync(FULL, !have);
        if (need != 0u && !exhausted) {
            const int n_need = __popc(need);
            const int leader = __ffs(need) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(work, n_need);
            base = __shfl_sync(FULL, base, leader);
            if (!have) {
                const int s = base + __popc(need & lt_mask);
                if (s < total) {
                    const float4 o = __ldg(ws.ray_o + s);
                    if (o.w != 0.0f) {
                        const float4 d = __ldg(ws.ray_d + s);
                        r = make_ray(make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z));
                        slot = s;
                        sp = 0;
                        cur = 0;
                        have = true;
                    }
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
            if (cur >= 0) {
                const PackedNode *pn = bvh.nodes + cur;
                const float4 a = __ldg(&pn->a), b = __ldg(&pn->b), c = __ldg(&pn->c);
                const int4 d = __ldg(&pn->d);
                float ln, lf, rn, rf;
                slab(r, a.x, a.y, a.z, a.w, b.x, b.y, ln, lf);
                slab(r, b.z, b.w, c.x, c.y, c.z, c.w, rn, rf);
                const bool passL = fminf(1e7f, lf) > ln;
                const bool passR = fminf(1e7f, rf) > rn;
                if (passR) {
                    cur = d.y;
                    if (passL) stack[sp++] = d.x;
                } else if (passL) {
                    cur = d.x;
                } else if (sp > 0) {
                    cur = stack[--sp];
                } else {
                    ws.hit[slot] = 0u;
                    have = false;
                }
            } else {
                const float4 *tp = bvh.tris + 3 * (size_t)(~cur);
                float t, u, v;
                if (tri_test(r, __ldg(tp), __ldg(tp + 1), __ldg(tp + 2), t, u, v)) {
                    ws.hit[slot] = 1u;
                    have = false;
                } else if (sp > 0) {
                    cur = stack[--sp];
                } else {
                    ws.hit[slot] = 0u;
                    have = false;
                }
            }
        }
        __syncwarp();
    }
}
#endif

int trace_queue_any(const BvhView &bvh, const Workspace &ws, int rays_per_item, int sm_count, cudaStream_t st)
{
#if defined(MR_HOST_CHECK)
    (void)sm_count;
    QueueTraceParams p = {bvh, ws, rays_per_item};
    return foreach_item<QueueTraceParams, queue_any_item, 128>(p, ws.capacity * rays_per_item, st);
#else
    cudaMemsetAsync(ws.counters + 1, 0, sizeof(int), st);
    const int blocks = sm_count * (2048 / MR_TRACE_BLOCK);
    k_trace_any_persistent<<<blocks, MR_TRACE_BLOCK, 0, st>>>(bvh, ws, rays_per_item);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
#endif
}

#if !defined(MR_HOST_CHECK)
// ---- persistent closest-hit tracer: same refill scheme, reference visit order (see closest_hit in mr_bvh.cuh) --------
__global__ void __launch_bounds__(MR_TRACE_BLOCK) k_trace_closest_persistent(BvhView bvh, Workspace ws)
{
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int total = ws.counters[0];
    int *work = ws.counters + 1;
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
            if ((int)lane == leader) base = atomicAdd(work, n_need);
            base = __shfl_sync(FULL, base, leader);
            if (!have) {
                const int s = base + __popc(need & lt_mask);
                if (s < total) {
                    const float4 o = __ldg(ws.ray_o + s);
                    if (o.w != 0.0f) {
                        const float4 d = __ldg(ws.ray_d + s);
                        r = make_ray(make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z));
                        slot = s;
                        sp = 0;
                        cur = 0;
                        closest = 1e7f;
                        any = false;
                        best_slot = -1;
                        have = true;
                    }
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
            if (cur >= 0) {
                const PackedNode *pn = bvh.nodes + cur;
                const float4 a = __ldg(&pn->a), b = __ldg(&pn->b), c = __ldg(&pn->c);
                const int4 d = __ldg(&pn->d);
                float ln, lf, rn, rf;
                slab(r, a.x, a.y, a.z, a.w, b.x, b.y, ln, lf);
                slab(r, b.z, b.w, c.x, c.y, c.z, c.w, rn, rf);
                const bool passL = fminf(closest, lf) > ln;
                const bool passR = fminf(closest, rf) > rn;
                if (passR) {
                    cur = d.y;
                    if (passL) {
                        stack_ref[sp] = d.x;
                        stack_t[sp] = ln;
                        ++sp;
                    }
                } else if (passL) {
                    cur = d.x;
                } else {
                    pop = true;
                }
            } else {
                const int leaf = ~cur;
                const float4 *tp = bvh.tris + 3 * (size_t)leaf;
                float t, u, v;
                if (tri_test(r, __ldg(tp), __ldg(tp + 1), __ldg(tp + 2), t, u, v)) {
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
                    if (any) {
                        pos = r.o + closest * r.d;
                        if (best_slot >= 0) {
                            const float4 *tp = bvh.tris + 3 * (size_t)best_slot;
                            const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
                            float t, u, v;
                            tri_test(r, q0, q1, q2, t, u, v);
                            const float3 fn = normalize(cross(make_float3(q1.x, q1.y, q1.z), make_float3(q2.x, q2.y, q2.z)));
                            const float w = 1.0f - u - v;
                            n = u * fn + v * fn + w * fn;
                            if (dot(-r.d, n) < 0) n = -n;
                            n = normalize(n);
                        }
                    }
                    ws.chit[2 * (size_t)slot] = make_float4(pos.x, pos.y, pos.z, any ? 1.0f : 0.0f);
                    ws.chit[2 * (size_t)slot + 1] = make_float4(n.x, n.y, n.z, any ? closest : 0.f);
                    have = false;
                }
            }
        }
        __syncwarp();
    }
}
#endif

int trace_queue_closest(const BvhView &bvh, const Workspace &ws, int sm_count, cudaStream_t st)
{
#if defined(MR_HOST_CHECK)
    (void)sm_count;
    QueueTraceParams p = {bvh, ws, 1};
    return foreach_item<QueueTraceParams, queue_closest_item, 128>(p, ws.capacity, st);
#else
    cudaMemsetAsync(ws.counters + 1, 0, sizeof(int), st);
    const int blocks = sm_count * (1024 / MR_TRACE_BLOCK);
    k_trace_closest_persistent<<<blocks, MR_TRACE_BLOCK, 0, st>>>(bvh, ws);
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
    ws.counters[1] = 0;
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
