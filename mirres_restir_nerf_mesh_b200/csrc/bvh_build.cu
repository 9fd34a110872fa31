// mirres-b200: LBVH construction on sm_100a, one stream-ordered call, no host synchronisation.
//
// Replaces nerf/renderer_restir.py:25-89 (restirbvhWorker.update_bvh) and the kernels it launches:
//   nerf/bvhworkers/get_elements.slang:3-39          -> k_elements (+ fused scene-extent reduction,
//                                                        replacing six torch min()/max() + host syncs)
//   nerf/bvhworkers/lbvh_morton_codes.slang:46-79    -> k_morton
//   nerf/bvhworkers/lbvh_single_radixsort.slang      -> multi-block stable LSD radix sort (4 x 8 bit)
//   nerf/bvhworkers/lbvh_hierarchy.slang:111-245     -> k_hierarchy
//   nerf/bvhworkers/lbvh_bounding_boxes.slang:151-390-> k_refit (single bottom-up pass with arrival
//                                                        counters instead of ~tree-height launches)
// plus k_pack, which emits the traversal records of mr_bvh.cuh.  Outputs in the reference layout:
// info [2F-1,3] i32 (left,right,prim), aabb [2F-1,6] f32; leaves at [F-1,2F-2], root 0.
#include "mr_bvh.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

int device_sm_count(); // wave.cu: per-device cache

#define MR_DONLY __device__ __forceinline__

// ---- order-preserving float <-> uint (for atomicMin/Max on the scene extent) --------------------
MR_DONLY unsigned int f2ord(float f)
{
    unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
MR_DONLY float ord2f(unsigned int u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// everything a build needs zeroed / initialised, in one launch: scene extent (ordered uints), digit histograms, the
// sort's published counts and tickets
__global__ void __launch_bounds__(256) k_build_init(unsigned int *extent, unsigned int *zero_words, size_t n_zero)
{
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_zero; i += stride) zero_words[i] = 0u;
    if (blockIdx.x == 0 && extent) {
        if (threadIdx.x < 3) extent[threadIdx.x] = 0xffffffffu;
        else if (threadIdx.x < 6) extent[threadIdx.x] = 0u;
    }
}

__global__ void __launch_bounds__(256) k_elements(const float *__restrict__ vert, const int *__restrict__ tri, int F,
                                                  float *__restrict__ eaabb, int *__restrict__ prim_idx,
                                                  unsigned int *extent)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {1e9f, 1e9f, 1e9f}, mx[3] = {-1e9f, -1e9f, -1e9f};
    bool live = p < F;
    if (live) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int vi = __ldg(tri + 3 * (size_t)p + i);
            float3 v = load3(vert, (size_t)vi);
            mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
            mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
        }
        float lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(mn[k], mx[k]); hi[k] = fmaxf(mn[k], mx[k]); }
#pragma unroll
        for (int k = 0; k < 3; ++k) { eaabb[6 * (size_t)p + k] = lo[k]; eaabb[6 * (size_t)p + 3 + k] = hi[k]; mn[k] = lo[k]; mx[k] = hi[k]; }
        if (prim_idx) prim_idx[p] = p;
    }
    if (!extent) return;
    // warp reduce, block reduce in shared memory, then one atomic per block and component (all blocks hit the same six
    // words, so the number of global atomics is what this kernel's time is made of)
    __shared__ unsigned int s_mn[3][8], s_mx[3][8];
    unsigned int omn[3], omx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        omn[k] = live ? f2ord(mn[k]) : 0xffffffffu;
        omx[k] = live ? f2ord(mx[k]) : 0u;
        omn[k] = __reduce_min_sync(0xffffffffu, omn[k]);
        omx[k] = __reduce_max_sync(0xffffffffu, omx[k]);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_mn[k][wid] = omn[k]; s_mx[k][wid] = omx[k]; }
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            unsigned int a = lane < 8 ? s_mn[k][lane] : 0xffffffffu, b = lane < 8 ? s_mx[k][lane] : 0u;
            a = __reduce_min_sync(0xffffffffu, a);
            b = __reduce_max_sync(0xffffffffu, b);
            if (lane == 0) { atomicMin(extent + k, a); atomicMax(extent + 3 + k, b); }
        }
    }
}

MR_DONLY unsigned int expand_bits(unsigned int v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
MR_DONLY unsigned int morton3d(float x, float y, float z)
{
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    return expand_bits((unsigned int)x) * 4 + expand_bits((unsigned int)y) * 2 + expand_bits((unsigned int)z);
}

// extent_ord: device-resident ordered-uint extent (fused path) or nullptr with explicit floats (granular path)
__global__ void __launch_bounds__(256) k_morton(const float *__restrict__ eaabb, int F, const unsigned int *extent_ord,
                                                float gminx, float gminy, float gminz, float gmaxx, float gmaxy,
                                                float gmaxz, unsigned int *__restrict__ keys, int *__restrict__ vals,
                                                int *__restrict__ pairs)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= F) return;
    float g[6] = {gminx, gminy, gminz, gmaxx, gmaxy, gmaxz};
    if (extent_ord) {
#pragma unroll
        for (int k = 0; k < 6; ++k) g[k] = ord2f(extent_ord[k]);
    }
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo = __ldg(eaabb + 6 * (size_t)p + k), hi = __ldg(eaabb + 6 * (size_t)p + 3 + k);
        float center = lo + 0.5f * (hi - lo);
        c[k] = (center - g[k]) / (g[3 + k] - g[k]);
    }
    unsigned int code = morton3d(c[0], c[1], c[2]);
    if (keys) { keys[p] = code; vals[p] = p; }
    if (pairs) { pairs[2 * (size_t)p] = (int)code; pairs[2 * (size_t)p + 1] = p; }
}

// box of one triangle: the operations of generateElements (get_elements.slang:3-39)
MR_DONLY void tri_box(const float *__restrict__ vert, const int *__restrict__ tri, int p, float lo[3], float hi[3])
{
    float mn[3] = {1e9f, 1e9f, 1e9f}, mx[3] = {-1e9f, -1e9f, -1e9f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int vi = __ldg(tri + 3 * (size_t)p + i);
        const float3 v = load3(vert, (size_t)vi);
        mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
        mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = fminf(mn[k], mx[k]); hi[k] = fmaxf(mn[k], mx[k]); }
}

// scene extent only (the element boxes are recomputed where they are needed: the mesh is L2-resident and 24 bytes per
// triangle written and read back twice cost more than the three gathers)
__global__ void __launch_bounds__(256) k_extent(const float *__restrict__ vert, const int *__restrict__ tri, int F, unsigned int *extent)
{
    unsigned int omn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, omx[3] = {0u, 0u, 0u};
    for (int p = blockIdx.x * 256 + threadIdx.x; p < F; p += gridDim.x * 256) {
        float lo[3], hi[3];
        tri_box(vert, tri, p, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) { omn[k] = min(omn[k], f2ord(lo[k])); omx[k] = max(omx[k], f2ord(hi[k])); }
    }
    __shared__ unsigned int s_mn[3][8], s_mx[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        omn[k] = __reduce_min_sync(0xffffffffu, omn[k]);
        omx[k] = __reduce_max_sync(0xffffffffu, omx[k]);
        if (lane == 0) { s_mn[k][wid] = omn[k]; s_mx[k][wid] = omx[k]; }
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            unsigned int a = lane < 8 ? s_mn[k][lane] : 0xffffffffu, b = lane < 8 ? s_mx[k][lane] : 0u;
            a = __reduce_min_sync(0xffffffffu, a);
            b = __reduce_max_sync(0xffffffffu, b);
            if (lane == 0) { atomicMin(extent + k, a); atomicMax(extent + 3 + k, b); }
        }
    }
}

// Morton code of every triangle (lbvh_morton_codes.slang:46-79) + the digit histograms of all three sort passes
__global__ void __launch_bounds__(256) k_morton_hist(const float *__restrict__ vert, const int *__restrict__ tri, int F,
                                                     const unsigned int *__restrict__ extent_ord, unsigned int *__restrict__ keys,
                                                     unsigned int *ghist)
{
    __shared__ unsigned int h[3][1024];
    for (int k = threadIdx.x; k < 3 * 1024; k += 256) (&h[0][0])[k] = 0u;
    __syncthreads();
    float g[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = ord2f(extent_ord[k]);
    for (int p = blockIdx.x * 256 + threadIdx.x; p < F; p += gridDim.x * 256) {
        float lo[3], hi[3], c[3];
        tri_box(vert, tri, p, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float center = lo[k] + 0.5f * (hi[k] - lo[k]);
            c[k] = (center - g[k]) / (g[3 + k] - g[k]);
        }
        const unsigned int code = morton3d(c[0], c[1], c[2]);
        keys[p] = code;
        atomicAdd(&h[0][code & 1023u], 1u);
        atomicAdd(&h[1][(code >> 10) & 1023u], 1u);
        atomicAdd(&h[2][(code >> 20) & 1023u], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * 1024; k += 256) {
        const unsigned int v = (&h[0][0])[k];
        if (v) atomicAdd(ghist + k, v);
    }
}

// ---- stable LSD radix sort of (30-bit Morton code, index): 3 passes x 10 bits, ONE launch per pass --------------------
// The reference sorts inside a single thread block (lbvh_single_radixsort.slang:28-138); any stable sort by code gives
// the same order.  Round 1 used 4 x 8 bits with three launches per pass (histogram, scan, scatter).  Here:
//   * the digit histograms of all three passes are counted once, while the codes are produced (k_morton_hist);
//   * a pass is one persistent launch of G <= resident blocks; block b owns a contiguous chunk of the input.  It counts
//     the digits of its chunk, PUBLISHES the counts (one flagged word per digit), sums the counts of the blocks before
//     it (waiting for each word to appear: blocks take their chunk by ticket, so every block it waits for is already
//     running), and then scatters its chunk tile by tile with a stable in-block ranking (__match_any per warp round,
//     per-warp digit counters in shared memory).
#define RS_BITS 10
#define RS_BINS 1024
#define RS_PASSES 3
#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#ifndef RS_ROUNDS
#define RS_ROUNDS 8
#endif
#define RS_TILE (RS_THREADS * RS_ROUNDS) // 4096 keys
#define RS_FLAG 0x80000000u

// sort_state words: [0..2] chunk tickets of the three passes, [3] arrival counter of k_leaves, [4] error flag
#define SS_TICKET 0
#define SS_LEAVES_DONE 3
#define SS_ERROR 4
#define RS_SPIN_LIMIT (1u << 24)
#define SS_WORDS 8

__global__ void __launch_bounds__(RS_THREADS) k_sort_pass(const unsigned int *__restrict__ keys_in, const int *__restrict__ vals_in,
                                                          int n, int pass, int chunk, int nblocks, const unsigned int *__restrict__ ghist,
                                                          unsigned int *agg, unsigned int *state, unsigned int *__restrict__ keys_out,
                                                          int *__restrict__ vals_out)
{
    __shared__ unsigned short wcnt[RS_WARPS][RS_BINS]; // per warp: digit counts of the current tile, then exclusive warp offsets
    __shared__ unsigned int base[RS_BINS];             // output position of the next key of each digit from this block
    __shared__ unsigned int scan_tmp[RS_WARPS];
    __shared__ int s_block;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int shift = RS_BITS * pass;
    if (tid == 0) s_block = (int)atomicAdd(state + SS_TICKET + pass, 1u);
    base[tid] = 0u;
    base[tid + RS_THREADS] = 0u;
    __syncthreads();
    const int b = s_block;
    const int lo = min(n, b * chunk), hi = min(n, lo + chunk);
    // ---- digit counts of the chunk.  The eight keys a thread handles per tile are loaded with independent loads (a loop
    // of dependent L2 round trips otherwise), and when the chunk is a single tile they stay in registers for the scatter.
    const int slot0 = wid * (32 * RS_ROUNDS) + lane; // position of this thread's round-0 key inside a tile
    const bool single = hi - lo <= RS_TILE;
    unsigned int key[RS_ROUNDS];
    for (int tile = lo; tile < hi; tile += RS_TILE) {
#pragma unroll
        for (int j = 0; j < RS_ROUNDS; ++j) {
            const int i = tile + slot0 + j * 32;
            key[j] = i < hi ? keys_in[i] : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < RS_ROUNDS; ++j)
            if (tile + slot0 + j * 32 < hi) atomicAdd(&base[(key[j] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    // ---- publish; exclusive bin bases from the global histogram; counts of the blocks before this one
    unsigned int *my_agg = agg + ((size_t)pass * nblocks + b) * RS_BINS;
    const unsigned int c0 = base[2 * tid], c1 = base[2 * tid + 1]; // thread t owns bins 2t, 2t+1
    volatile unsigned int *my_pub = my_agg;
    my_pub[2 * tid] = c0 | RS_FLAG;
    my_pub[2 * tid + 1] = c1 | RS_FLAG;
    const unsigned int g0 = ghist[pass * RS_BINS + 2 * tid], g1 = ghist[pass * RS_BINS + 2 * tid + 1];
    unsigned int x = g0 + g1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) scan_tmp[wid] = x;
    __syncthreads();
    unsigned int before = x - (g0 + g1);
    for (int w = 0; w < wid; ++w) before += scan_tmp[w];
    unsigned int e0 = before, e1 = before + g0;
    {
        // sixteen blocks per round trip: the loads of a batch are independent, a word that is not there yet is re-read
        const volatile unsigned int *all = agg + (size_t)pass * nblocks * RS_BINS + 2 * tid;
        for (int t0 = 0; t0 < b; t0 += 16) {
            unsigned int vx[16], vy[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const bool in = t0 + k < b;
                vx[k] = in ? all[(size_t)(t0 + k) * RS_BINS] : RS_FLAG;
                vy[k] = in ? all[(size_t)(t0 + k) * RS_BINS + 1] : RS_FLAG;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                // (the wait is bounded: a word that never appears -- a foreign write into the scratch buffer -- raises the
                // error word instead of hanging the device; every block it waits for is running, see above)
                for (unsigned int spin = 0; !(vx[k] & RS_FLAG); ++spin) {
                    if (spin > RS_SPIN_LIMIT) { state[SS_ERROR] = 1u; break; }
                    vx[k] = all[(size_t)(t0 + k) * RS_BINS];
                }
                for (unsigned int spin = 0; !(vy[k] & RS_FLAG); ++spin) {
                    if (spin > RS_SPIN_LIMIT) { state[SS_ERROR] = 1u; break; }
                    vy[k] = all[(size_t)(t0 + k) * RS_BINS + 1];
                }
                e0 += vx[k] & ~RS_FLAG;
                e1 += vy[k] & ~RS_FLAG;
            }
        }
    }
    __syncthreads();
    base[2 * tid] = e0;
    base[2 * tid + 1] = e1;
    // ---- scatter, tile by tile
    for (int tile = lo; tile < hi; tile += RS_TILE) {
        for (int k = tid; k < RS_WARPS * RS_BINS / 2; k += RS_THREADS) reinterpret_cast<unsigned int *>(&wcnt[0][0])[k] = 0u;
        __syncthreads();
        unsigned short rank[RS_ROUNDS];
        int val[RS_ROUNDS];
        if (!single) {
#pragma unroll
            for (int j = 0; j < RS_ROUNDS; ++j) {
                const int i = tile + slot0 + j * 32;
                key[j] = i < hi ? keys_in[i] : 0xffffffffu;
            }
        }
#pragma unroll
        for (int j = 0; j < RS_ROUNDS; ++j) {
            const int i = tile + slot0 + j * 32;
            val[j] = vals_in ? (i < hi ? vals_in[i] : 0) : i;
        }
#pragma unroll
        for (int j = 0; j < RS_ROUNDS; ++j) {
            const int i = tile + slot0 + j * 32;
            const bool live = i < hi;
            const unsigned int digit = (key[j] >> shift) & (RS_BINS - 1);
            const unsigned int live_mask = __ballot_sync(0xffffffffu, live);
            const unsigned int peers = __match_any_sync(0xffffffffu, live ? digit : 0xffffffffu) & live_mask;
            unsigned int cnt = 0;
            if (live) cnt = wcnt[wid][digit];
            __syncwarp();
            rank[j] = (unsigned short)(cnt + __popc(peers & ((1u << lane) - 1u)));
            if (live && (peers & ((1u << lane) - 1u)) == 0u) wcnt[wid][digit] = (unsigned short)(cnt + __popc(peers));
            __syncwarp();
        }
        __syncthreads();
        // exclusive offsets of the warps within the tile, per digit; tile totals advance the block's bases afterwards
        unsigned int tot0 = 0, tot1 = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const unsigned int a0 = wcnt[w][2 * tid], a1 = wcnt[w][2 * tid + 1];
            wcnt[w][2 * tid] = (unsigned short)tot0;
            wcnt[w][2 * tid + 1] = (unsigned short)tot1;
            tot0 += a0;
            tot1 += a1;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ROUNDS; ++j) {
            const int i = tile + slot0 + j * 32;
            if (i < hi) {
                const unsigned int digit = (key[j] >> shift) & (RS_BINS - 1);
                const unsigned int dst = base[digit] + wcnt[wid][digit] + rank[j];
                keys_out[dst] = key[j];
                vals_out[dst] = val[j];
            }
        }
        __syncthreads();
        base[2 * tid] += tot0;
        base[2 * tid + 1] += tot1;
        __syncthreads();
    }
}

// ---- Karras hierarchy ------------------------------------------------------------------------------
MR_DONLY int delta_fn(int i, unsigned int codeI, int j, int n, const unsigned int *__restrict__ codes)
{
    if (j < 0 || j > n - 1) return -1;
    unsigned int codeJ = __ldg(codes + j);
    if (codeI == codeJ) return 32 + __clz((unsigned int)i ^ (unsigned int)j); // 31 - findMSB(x) == clz(x), x != 0
    return __clz(codeI ^ codeJ);
}

// ---- node boxes without a bottom-up walk ---------------------------------------------------------------
// The reference refits the tree level by level (lbvh_bounding_boxes.slang:151-390, ~tree-height launches); round 1 used
// one launch in which the second thread to arrive at a node unions its children (a chain of ~40 dependent atomics and
// L2 round trips from the deepest leaf to the root).  An internal node of a Karras tree covers a CONTIGUOUS range
// [first, last] of the sorted leaves, and its box is the exact min / max over the leaf boxes of that range -- min and max
// do not round, so the result does not depend on the order of the unions.  The boxes therefore come straight from the
// range: k_leaves lays the leaf boxes out in sorted order and reduces them over aligned groups of 4, 16, 64 ... leaves,
// and k_hierarchy, which knows [first, last] anyway, unions at most 3 + 3 items per level.  No atomics, no chain.
#define BX_LEVELS 14 // lv[j]: unions of 4^j consecutive sorted leaves; 4^13 > 2^26 leaves
struct BoxTables {
    float4 *lv[BX_LEVELS]; // two float4 per item (min, max); lv[0]: the leaf boxes
    int n[BX_LEVELS];
    int levels;            // levels in use: n[levels - 1] <= 4
};

struct Box6 {
    float mn[3], mx[3];
};
MR_DONLY void box_empty(Box6 &b)
{
    const float inf = __int_as_float(0x7f800000);
    b.mn[0] = b.mn[1] = b.mn[2] = inf;
    b.mx[0] = b.mx[1] = b.mx[2] = -inf;
}
MR_DONLY void box_add(Box6 &b, const float4 lo, const float4 hi)
{
    b.mn[0] = fminf(b.mn[0], lo.x); b.mn[1] = fminf(b.mn[1], lo.y); b.mn[2] = fminf(b.mn[2], lo.z);
    b.mx[0] = fmaxf(b.mx[0], hi.x); b.mx[1] = fmaxf(b.mx[1], hi.y); b.mx[2] = fmaxf(b.mx[2], hi.z);
}
MR_DONLY void box_xor_step(Box6 &b, int o)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.mn[k] = fminf(b.mn[k], __shfl_xor_sync(0xffffffffu, b.mn[k], o));
        b.mx[k] = fmaxf(b.mx[k], __shfl_xor_sync(0xffffffffu, b.mx[k], o));
    }
}
MR_DONLY void box_store(float4 *dst, const Box6 &b)
{
    dst[0] = make_float4(b.mn[0], b.mn[1], b.mn[2], 0.f);
    dst[1] = make_float4(b.mx[0], b.mx[1], b.mx[2], 0.f);
}

// one block = 1024 consecutive sorted leaves: leaf records in the reference layout (hierarchy, lbvh_hierarchy.slang:
// 121-141), sorted (code, index) pairs on request, packed triangles on request, leaf boxes for the range unions and their
// unions over aligned groups of 4, 16, 64, 256 and 1024 leaves; the last block to finish reduces the upper levels
__global__ void __launch_bounds__(1024) k_leaves(int F, const unsigned int *__restrict__ codes, const int *__restrict__ sorted_idx,
                                                 const float *__restrict__ vert, const int *__restrict__ tri,
                                                 const float *__restrict__ eaabb, int *__restrict__ info, float *__restrict__ aabb,
                                                 int *__restrict__ sorted_pairs, PackedTri *__restrict__ ptris, BoxTables tb,
                                                 unsigned int *done)
{
    __shared__ float4 s_lo[64], s_hi[64];
    __shared__ unsigned int s_last;
    // the block's leaf records are two contiguous pieces of the output (1024 x 12 bytes of info, 1024 x 24 bytes of boxes):
    // they are staged here and written out as one coalesced stream each instead of nine strided 4-byte stores per thread
    // (ncu: the kernel was throttled by its store queue)
    __shared__ int s_info[1024 * 3];
    __shared__ float s_box[1024 * 6];
    const int gid = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Box6 b;
    box_empty(b);
    if (gid < F) {
        const int e = __ldg(sorted_idx + gid);
        float lo[3], hi[3];
        if (eaabb) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = __ldg(eaabb + 6 * (size_t)e + k); hi[k] = __ldg(eaabb + 6 * (size_t)e + 3 + k); }
        } else {
            tri_box(vert, tri, e, lo, hi);
        }
        s_info[3 * threadIdx.x] = 0; s_info[3 * threadIdx.x + 1] = 0; s_info[3 * threadIdx.x + 2] = e;
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_box[6 * threadIdx.x + k] = lo[k]; s_box[6 * threadIdx.x + 3 + k] = hi[k]; b.mn[k] = lo[k]; b.mx[k] = hi[k]; }
        box_store(tb.lv[0] + 2 * (size_t)gid, b);
        if (sorted_pairs) *reinterpret_cast<int2 *>(sorted_pairs + 2 * (size_t)gid) = make_int2((int)__ldg(codes + gid), e);
        if (ptris) {
            const int i0 = __ldg(tri + 3 * (size_t)e), i1 = __ldg(tri + 3 * (size_t)e + 1), i2 = __ldg(tri + 3 * (size_t)e + 2);
            const float3 v0 = load3(vert, (size_t)i0), v1 = load3(vert, (size_t)i1), v2 = load3(vert, (size_t)i2);
            const float3 e1 = v1 - v0, e2 = v2 - v0;
            float4 *q = reinterpret_cast<float4 *>(ptris + gid);
            q[0] = make_float4(v0.x, v0.y, v0.z, __int_as_float(e));
            q[1] = make_float4(e1.x, e1.y, e1.z, 0.f);
            q[2] = make_float4(e2.x, e2.y, e2.z, 0.f);
            q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
    {
        const int first = blockIdx.x * 1024, count = min(1024, F - first);
        int *gi = info + 3 * (size_t)(F - 1 + first);
        float *gb = aabb + 6 * (size_t)(F - 1 + first);
        for (int k = threadIdx.x; k < 3 * count; k += 1024) gi[k] = s_info[k];
        for (int k = threadIdx.x; k < 6 * count; k += 1024) gb[k] = s_box[k];
    }
    // groups of 4 and 16 inside the warp (a butterfly step k leaves every lane with the union of its aligned 2^k group)
    box_xor_step(b, 1);
    box_xor_step(b, 2);
    if ((lane & 3) == 0 && (gid >> 2) < tb.n[1]) box_store(tb.lv[1] + 2 * (size_t)(gid >> 2), b);
    box_xor_step(b, 4);
    box_xor_step(b, 8);
    if ((lane & 15) == 0) {
        if ((gid >> 4) < tb.n[2]) box_store(tb.lv[2] + 2 * (size_t)(gid >> 4), b);
        s_lo[threadIdx.x >> 4] = make_float4(b.mn[0], b.mn[1], b.mn[2], 0.f);
        s_hi[threadIdx.x >> 4] = make_float4(b.mx[0], b.mx[1], b.mx[2], 0.f);
    }
    __syncthreads();
    // groups of 64, 256, 1024 from the block's 64 groups of 16: warp 0, the same butterfly over 64 -> 16 -> 4 -> 1 items
    if (wid == 0) {
        Box6 c;
        box_empty(c);
        box_add(c, s_lo[2 * lane], s_hi[2 * lane]);
        box_add(c, s_lo[2 * lane + 1], s_hi[2 * lane + 1]); // lane = group of 32 leaves
        box_xor_step(c, 1);                                  // 64 leaves
        if ((lane & 1) == 0) {
            const int i3 = blockIdx.x * 16 + (lane >> 1);
            if (i3 < tb.n[3]) box_store(tb.lv[3] + 2 * (size_t)i3, c);
        }
        box_xor_step(c, 2);
        box_xor_step(c, 4); // 256 leaves
        if ((lane & 7) == 0) {
            const int i4 = blockIdx.x * 4 + (lane >> 3);
            if (i4 < tb.n[4]) box_store(tb.lv[4] + 2 * (size_t)i4, c);
        }
        box_xor_step(c, 8);
        box_xor_step(c, 16); // 1024 leaves
        if (lane == 0) {
            if (blockIdx.x < tb.n[5]) box_store(tb.lv[5] + 2 * (size_t)blockIdx.x, c);
            __threadfence();
            s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // upper levels (a few hundred items at most), by the last block alone
    for (int j = 6; j < tb.levels; ++j) {
        for (int i = threadIdx.x; i < tb.n[j]; i += 1024) {
            Box6 c;
            box_empty(c);
            for (int k = 0; k < 4; ++k) {
                const int src = 4 * i + k;
                if (src < tb.n[j - 1]) box_add(c, __ldcg(tb.lv[j - 1] + 2 * (size_t)src), __ldcg(tb.lv[j - 1] + 2 * (size_t)src + 1));
            }
            box_store(tb.lv[j] + 2 * (size_t)i, c);
        }
        __threadfence();
        __syncthreads();
    }
    if (threadIdx.x == 0) *done = 0u;
}

// union of the leaf boxes [l, r] (inclusive): at most 3 + 3 ragged items per level, whole groups of four from the next
MR_DONLY void range_box(const BoxTables &tb, int l, int r, Box6 &b)
{
    box_empty(b);
#pragma unroll 1
    for (int j = 0; j < BX_LEVELS; ++j) {
        const float4 *t = tb.lv[j];
        const int bl = l >> 2, br = r >> 2;
        if (bl == br || j == tb.levels - 1) {
            for (int i = l; i <= r; ++i) box_add(b, __ldg(t + 2 * (size_t)i), __ldg(t + 2 * (size_t)i + 1));
            return;
        }
        int nl = bl, nr = br;
        if ((l & 3) != 0) {
            for (int i = l; i <= (bl << 2) + 3; ++i) box_add(b, __ldg(t + 2 * (size_t)i), __ldg(t + 2 * (size_t)i + 1));
            nl = bl + 1;
        }
        if ((r & 3) != 3) {
            for (int i = br << 2; i <= r; ++i) box_add(b, __ldg(t + 2 * (size_t)i), __ldg(t + 2 * (size_t)i + 1));
            nr = br - 1;
        }
        if (nl > nr) return;
        l = nl;
        r = nr;
    }
}

__global__ void __launch_bounds__(256) k_hierarchy(int F, const unsigned int *__restrict__ codes, int *__restrict__ info,
                                                   float *__restrict__ aabb, BoxTables tb)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F - 1) return;
    const int LEAF = F - 1;
    unsigned int code = __ldg(codes + idx);
    int deltaL = delta_fn(idx, code, idx - 1, F, codes);
    int deltaR = delta_fn(idx, code, idx + 1, F, codes);
    int d = (deltaR >= deltaL) ? 1 : -1;
    int deltaMin = min(deltaL, deltaR);
    int lMax = 2;
    while (delta_fn(idx, code, idx + lMax * d, F, codes) > deltaMin) lMax <<= 1;
    int l = 0;
    for (int t = lMax >> 1; t > 0; t >>= 1)
        if (delta_fn(idx, code, idx + (l + t) * d, F, codes) > deltaMin) l += t;
    int jdx = idx + l * d;
    int first = min(idx, jdx), last = max(idx, jdx);
    unsigned int firstCode = __ldg(codes + first);
    int commonPrefix = delta_fn(first, firstCode, last, F, codes);
    int split = first, stride = last - first;
    do {
        stride = (stride + 1) >> 1;
        int newSplit = split + stride;
        if (newSplit < last) {
            int splitPrefix = delta_fn(first, firstCode, newSplit, F, codes);
            if (splitPrefix > commonPrefix) split = newSplit;
        }
    } while (stride > 1);
    int childA = (split == first) ? LEAF + split : split;
    int childB = (split + 1 == last) ? LEAF + split + 1 : split + 1;
    info[3 * (size_t)idx] = childA; info[3 * (size_t)idx + 1] = childB; info[3 * (size_t)idx + 2] = 0;
    Box6 b;
    range_box(tb, first, last, b);
#pragma unroll
    for (int k = 0; k < 3; ++k) { aabb[6 * (size_t)idx + k] = b.mn[k]; aabb[6 * (size_t)idx + 3 + k] = b.mx[k]; }
}

__global__ void k_unzip_pairs(const int *__restrict__ pairs, int n, unsigned int *__restrict__ keys, int *__restrict__ vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (unsigned int)pairs[2 * (size_t)i]; vals[i] = pairs[2 * (size_t)i + 1]; }
}
__global__ void k_zip_pairs(const unsigned int *__restrict__ keys, const int *__restrict__ vals, int n, int *__restrict__ pairs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pairs[2 * (size_t)i] = (int)keys[i]; pairs[2 * (size_t)i + 1] = vals[i]; }
}
// digit histograms of keys that did not come from k_morton_hist (granular sort entry point)
__global__ void __launch_bounds__(256) k_key_hist(const unsigned int *__restrict__ keys, int n, unsigned int *ghist)
{
    __shared__ unsigned int h[3][1024];
    for (int k = threadIdx.x; k < 3 * 1024; k += 256) (&h[0][0])[k] = 0u;
    __syncthreads();
    for (int p = blockIdx.x * 256 + threadIdx.x; p < n; p += gridDim.x * 256) {
        const unsigned int code = keys[p];
        atomicAdd(&h[0][code & 1023u], 1u);
        atomicAdd(&h[1][(code >> 10) & 1023u], 1u);
        atomicAdd(&h[2][(code >> 20) & 1023u], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * 1024; k += 256) {
        const unsigned int v = (&h[0][0])[k];
        if (v) atomicAdd(ghist + k, v);
    }
}

// scratch carving -----------------------------------------------------------------------------------------------------
#define RS_MAX_BLOCKS 512
struct BuildScratch {
    unsigned int *keys[2];  // F each
    int *vals[2];           // F each
    unsigned int *zeroed;   // start of the words k_build_init clears: ghist, state, agg (in this order)
    unsigned int *ghist;    // 3 x 1024 digit histograms
    unsigned int *state;    // SS_WORDS
    unsigned int *agg;      // 3 x blocks x 1024 published chunk counts
    unsigned int *extent;   // 8
    BoxTables tb;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t carve(BuildScratch *s, int F, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return base ? base + o : (char *)0; };
    char *p;
    for (int k = 0; k < 2; ++k) { p = take((size_t)F * 4); if (s) s->keys[k] = (unsigned int *)p; }
    for (int k = 0; k < 2; ++k) { p = take((size_t)F * 4); if (s) s->vals[k] = (int *)p; }
    p = take((size_t)RS_PASSES * RS_BINS * 4); if (s) { s->ghist = (unsigned int *)p; s->zeroed = s->ghist; }
    p = take(256); if (s) s->state = (unsigned int *)p;
    p = take((size_t)RS_PASSES * RS_MAX_BLOCKS * RS_BINS * 4); if (s) s->agg = (unsigned int *)p;
    p = take(64); if (s) s->extent = (unsigned int *)p;
    int n = F, levels = 0;
    for (int j = 0; j < BX_LEVELS; ++j) {
        p = take((size_t)n * 2 * sizeof(float4));
        if (s) { s->tb.lv[j] = (float4 *)p; s->tb.n[j] = n; }
        if (!levels && (n <= 4 || j == BX_LEVELS - 1)) levels = j + 1;
        n = (n + 3) / 4;
    }
    if (s) s->tb.levels = levels;
    return off;
}

// chunk of the input one sort block owns, and the number of blocks that makes
static void sort_shape(int F, int &chunk, int &blocks)
{
    int most = 2 * device_sm_count();
    if (most > RS_MAX_BLOCKS) most = RS_MAX_BLOCKS;
    if (most < 1) most = 1;
    chunk = (F + most - 1) / most;
    chunk = (chunk + RS_TILE - 1) / RS_TILE * RS_TILE;
    blocks = (F + chunk - 1) / chunk;
}

// words k_build_init has to clear for a sort of `blocks` chunks: histograms, state, the published counts in use
static size_t zero_words(const BuildScratch &s, int blocks)
{
    return (size_t)(s.agg - s.zeroed) + (size_t)RS_PASSES * blocks * RS_BINS;
}

// keys[0] (+ vals[0] unless `iota`: the values are then the element indices) -> keys[1] / vals[1], sorted and stable.
// ghist, state and agg must have been cleared and ghist filled (k_morton_hist / k_key_hist).
static int sort_pairs(BuildScratch &s, int F, bool iota, cudaStream_t st)
{
    int chunk, blocks;
    sort_shape(F, chunk, blocks);
    // pass 0: 0 -> 1, pass 1: 1 -> 0, pass 2: 0 -> 1
    for (int pass = 0; pass < RS_PASSES; ++pass) {
        const int in = pass & 1, out = in ^ 1;
        k_sort_pass<<<blocks, RS_THREADS, 0, st>>>(s.keys[in], (pass == 0 && iota) ? nullptr : s.vals[in], F, pass, chunk, blocks,
                                                   s.ghist, s.agg, s.state, s.keys[out], s.vals[out]);
    }
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

static int init_grid(size_t words)
{
    size_t b = (words + 1023) / 1024;
    return (int)(b < 1 ? 1 : (b > 592 ? 592 : b));
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_abi_version(void) { return MIRRES_ABI_VERSION; }

size_t mirres_bvh_scratch_bytes(int F) { return F < 1 ? 0 : carve(nullptr, F, nullptr); }
size_t mirres_bvh_packed_node_bytes(int F) { return F < 1 ? 0 : MR_TOP_BYTES + sizeof(PackedNode) * (size_t)(F > 1 ? F - 1 : 1); }
size_t mirres_bvh_packed_tri_bytes(int F) { return F < 1 ? 0 : sizeof(PackedTri) * (size_t)F; }

// Nine launches (round 1: twenty): init | scene extent | codes + digit histograms | 3 sort passes | leaf records, leaf
// boxes and their block unions | hierarchy + node boxes from the leaf ranges | traversal records.
int mirres_bvh_build(const float *vert, int V, const int *tri, int F, int *info, float *aabb, void *packed_nodes,
                     void *packed_tris, int *sorted_codes, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!vert || !tri || !info || !aabb || !scratch) return MIRRES_ERR_NULL;
    if (F < 1 || V < 1) return MIRRES_ERR_SHAPE;
    if (scratch_bytes < carve(nullptr, F, nullptr)) return MIRRES_ERR_SCRATCH;
    if (((uintptr_t)scratch & 255) || ((uintptr_t)packed_nodes & 31) || ((uintptr_t)packed_tris & 31)) return MIRRES_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    carve(&s, F, (char *)scratch);
    int chunk, blocks;
    sort_shape(F, chunk, blocks);
    const size_t zw = zero_words(s, blocks);
    const int sweep = min((F + 255) / 256, 4 * device_sm_count());
    k_build_init<<<init_grid(zw), 256, 0, st>>>(s.extent, s.zeroed, zw);
    k_extent<<<sweep, 256, 0, st>>>(vert, tri, F, s.extent);
    k_morton_hist<<<sweep, 256, 0, st>>>(vert, tri, F, s.extent, s.keys[0], s.ghist);
    int rc = sort_pairs(s, F, true, st);
    if (rc) return rc;
    const bool pack = packed_nodes && packed_tris;
    k_leaves<<<(F + 1023) / 1024, 1024, 0, st>>>(F, s.keys[1], s.vals[1], vert, tri, nullptr, info, aabb, sorted_codes,
                                                 pack ? (PackedTri *)packed_tris : nullptr, s.tb, s.state + SS_LEAVES_DONE);
    if (F > 1) k_hierarchy<<<(F - 1 + 255) / 256, 256, 0, st>>>(F, s.keys[1], info, aabb, s.tb);
    MR_CUDA_CHECK_LAUNCH();
    if (pack) return pack_traversal(F, info, aabb, vert, tri, packed_nodes, nullptr, st);
    return 0;
}

// Granular entry points with the reference kernels' argument meaning (for the slangpy-protocol shim).
int mirres_bvh_elements(const float *vert, const int *tri, int F, int *ele_primitiveIdx, float *ele_aabb, void *stream)
{
    if (!vert || !tri || !ele_aabb) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    k_elements<<<(F + 255) / 256, 256, 0, (cudaStream_t)stream>>>(vert, tri, F, ele_aabb, ele_primitiveIdx, nullptr);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

int mirres_bvh_morton(const float *ele_aabb, int F, float min_x, float min_y, float min_z, float max_x, float max_y,
                      float max_z, int *morton_codes_ele, void *stream)
{
    if (!ele_aabb || !morton_codes_ele) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    k_morton<<<(F + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ele_aabb, F, nullptr, min_x, min_y, min_z, max_x, max_y,
                                                                 max_z, nullptr, nullptr, morton_codes_ele);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

// Stable sort of (code, idx) pairs by code; result in `pairs` (the reference's g_elements_in).  Codes are Morton codes
// (30 bits, lbvh_morton_codes.slang:24-44); three 10-bit passes cover them.
int mirres_bvh_sort(int *pairs, int F, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!pairs || !scratch) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    if (scratch_bytes < carve(nullptr, F, nullptr)) return MIRRES_ERR_SCRATCH;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    carve(&s, F, (char *)scratch);
    int chunk, blocks;
    sort_shape(F, chunk, blocks);
    const size_t zw = zero_words(s, blocks);
    const int grid = (F + 255) / 256;
    k_build_init<<<init_grid(zw), 256, 0, st>>>(nullptr, s.zeroed, zw);
    k_unzip_pairs<<<grid, 256, 0, st>>>(pairs, F, s.keys[0], s.vals[0]);
    k_key_hist<<<min(grid, 4 * device_sm_count()), 256, 0, st>>>(s.keys[0], F, s.ghist);
    int rc = sort_pairs(s, F, false, st);
    if (rc) return rc;
    k_zip_pairs<<<grid, 256, 0, st>>>(s.keys[1], s.vals[1], F, pairs);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

// hierarchy + node boxes from sorted pairs and element boxes (reference kernels hierarchy, get_bbox*, set_root)
int mirres_bvh_hierarchy_refit(const int *sorted_pairs, const float *ele_aabb, int F, int *info, float *aabb,
                               void *scratch, size_t scratch_bytes, void *stream)
{
    if (!sorted_pairs || !ele_aabb || !info || !aabb || !scratch) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    if (scratch_bytes < carve(nullptr, F, nullptr)) return MIRRES_ERR_SCRATCH;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    carve(&s, F, (char *)scratch);
    const int grid = (F + 255) / 256;
    k_build_init<<<1, 256, 0, st>>>(nullptr, s.state, SS_WORDS);
    k_unzip_pairs<<<grid, 256, 0, st>>>(sorted_pairs, F, s.keys[0], s.vals[0]);
    k_leaves<<<(F + 1023) / 1024, 1024, 0, st>>>(F, s.keys[0], s.vals[0], nullptr, nullptr, ele_aabb, info, aabb, nullptr, nullptr,
                                                 s.tb, s.state + SS_LEAVES_DONE);
    if (F > 1) k_hierarchy<<<(F - 1 + 255) / 256, 256, 0, st>>>(F, s.keys[0], info, aabb, s.tb);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

} // extern "C"
