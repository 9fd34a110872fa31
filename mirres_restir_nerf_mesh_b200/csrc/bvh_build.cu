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

__global__ void k_init_extent(unsigned int *extent)
{
    if (threadIdx.x < 3) extent[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) extent[threadIdx.x] = 0u;
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

// ---- stable LSD radix sort, 8 bits per pass ------------------------------------------------------
#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const unsigned int *__restrict__ keys, int n, int shift,
                                                            unsigned int *__restrict__ hist, int num_blocks)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int i = base + j * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * num_blocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist[256][num_blocks] (bin-major order): block b (one warp) scans the row of bin b in place with
// coalesced 32-wide chunks and publishes the row total; the last block to finish scans the 256 totals into bin_base,
// which k_sort_scatter adds to the row offsets.  `done` must be zero on entry and is reset for the next pass.
__global__ void __launch_bounds__(32) k_sort_scan(unsigned int *hist, int num_blocks, unsigned int *bin_base, unsigned int *done)
{
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x, bin = blockIdx.x;
    unsigned int *row = hist + (size_t)bin * num_blocks;
    unsigned int carry = 0;
    for (int base = 0; base < num_blocks; base += 32) {
        const int i = base + lane;
        const unsigned int v = i < num_blocks ? row[i] : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (i < num_blocks) row[i] = carry + x - v;
        carry += __shfl_sync(FULL, x, 31);
    }
    unsigned int last = 0;
    if (lane == 0) {
        bin_base[256 + bin] = carry; // row totals live behind the 256 bases
        __threadfence();
        last = atomicAdd(done, 1u) == 255u ? 1u : 0u;
    }
    last = __shfl_sync(FULL, last, 0);
    if (!last) return;
    __threadfence();
    // exclusive scan of the 256 totals: 8 per lane
    unsigned int loc[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { loc[k] = sum; sum += __ldcg(bin_base + 256 + lane * 8 + k); }
    unsigned int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    const unsigned int before = x - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) bin_base[lane * 8 + k] = before + loc[k];
    if (lane == 0) *done = 0u;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const unsigned int *__restrict__ keys_in,
                                                               const int *__restrict__ vals_in, int n, int shift,
                                                               const unsigned int *__restrict__ hist, int num_blocks,
                                                               const unsigned int *__restrict__ bin_base,
                                                               unsigned int *__restrict__ keys_out,
                                                               int *__restrict__ vals_out)
{
    __shared__ unsigned int running[256];                     // keys of this digit already placed by this block
    __shared__ unsigned short warp_cnt[SORT_THREADS / 32][256]; // per-round, per-warp digit counts
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    running[threadIdx.x] = hist[threadIdx.x * num_blocks + blockIdx.x] + bin_base[threadIdx.x];
    int base = blockIdx.x * SORT_TILE;
    for (int j = 0; j < SORT_ITEMS; ++j) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        int i = base + j * SORT_THREADS + threadIdx.x;
        bool live = i < n;
        unsigned int key = live ? keys_in[i] : 0xffffffffu;
        int val = live ? vals_in[i] : 0;
        unsigned int digit = (key >> shift) & 255u;
        unsigned int live_mask = __ballot_sync(0xffffffffu, live);
        unsigned int peers = __match_any_sync(0xffffffffu, live ? digit : 0xffffffffu) & live_mask;
        unsigned int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (live && rank_in_warp == 0) warp_cnt[wid][digit] = (unsigned short)__popc(peers);
        __syncthreads();
        if (live) {
            unsigned int before = running[digit];
            for (int w = 0; w < wid; ++w) before += warp_cnt[w][digit];
            unsigned int dst = before + rank_in_warp;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
        {
            unsigned int add = 0;
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) add += warp_cnt[w][threadIdx.x];
            running[threadIdx.x] += add;
        }
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

__global__ void __launch_bounds__(256) k_hierarchy(int F, const unsigned int *__restrict__ codes,
                                                   const int *__restrict__ sorted_idx, const float *__restrict__ eaabb,
                                                   int *__restrict__ info, float *__restrict__ aabb,
                                                   int *__restrict__ parent)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= F) return;
    const int LEAF = F - 1;
    {
        int e = __ldg(sorted_idx + gid);
        size_t n = (size_t)(LEAF + gid);
        info[3 * n] = 0; info[3 * n + 1] = 0; info[3 * n + 2] = e;
#pragma unroll
        for (int k = 0; k < 6; ++k) aabb[6 * n + k] = __ldg(eaabb + 6 * (size_t)e + k);
    }
    if (gid == 0) parent[0] = 0;
    if (gid >= F - 1) return;
    const int idx = gid;
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
    parent[childA] = idx;
    parent[childB] = idx;
}

// ---- bottom-up refit: the second thread to arrive at a node unions its children -----------------------------
// A level costs dependent L2 round trips, so the walk keeps them to two: the node's static data (children, parent) is
// fetched BEFORE the arrival atomic, the walker carries the box it has just produced in registers and, once the atomic
// has told it that it is the second arrival, fetches only its sibling's box.  Operand order of the union is (left,
// right) as before.
__global__ void __launch_bounds__(256) k_refit(int F, const int *__restrict__ info, float *aabb,
                                               const int *__restrict__ parent, int *visits)
{
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= F || F < 2) return;
    int me = F - 1 + gid;
    float box[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) box[k] = __ldcg(aabb + 6 * (size_t)me + k);
    int node = __ldg(parent + me);
    for (;;) {
        const int l = __ldg(info + 3 * (size_t)node), r = __ldg(info + 3 * (size_t)node + 1);
        const int up = node ? __ldg(parent + node) : 0;
        __threadfence();
        int old = atomicAdd(visits + node, 1);
        if (old == 0) return;
        const bool left = l == me;
        const int sib = left ? r : l;
        float o[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = __ldcg(aabb + 6 * (size_t)sib + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lmn = left ? box[k] : o[k], rmn = left ? o[k] : box[k];
            const float lmx = left ? box[3 + k] : o[3 + k], rmx = left ? o[3 + k] : box[3 + k];
            box[k] = fminf(lmn, rmn);
            box[3 + k] = fmaxf(lmx, rmx);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcg(aabb + 6 * (size_t)node + k, box[k]);
        if (node == 0) return;
        me = node;
        node = up;
    }
}

__global__ void k_zip_pairs(const unsigned int *__restrict__ keys, const int *__restrict__ vals, int n, int *__restrict__ pairs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pairs[2 * (size_t)i] = (int)keys[i]; pairs[2 * (size_t)i + 1] = vals[i]; }
}
__global__ void k_unzip_pairs(const int *__restrict__ pairs, int n, unsigned int *__restrict__ keys, int *__restrict__ vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (unsigned int)pairs[2 * (size_t)i]; vals[i] = pairs[2 * (size_t)i + 1]; }
}

// scratch carving -----------------------------------------------------------------------------------------------------
struct BuildScratch {
    float *eaabb;           // F*6
    unsigned int *keys[2];  // F each
    int *vals[2];           // F each
    unsigned int *hist;     // 256*num_blocks
    unsigned int *bin_base; // 512: exclusive bin bases, row totals
    unsigned int *done;     // arrival counter of the scan
    int *parent;            // 2F-1
    int *visits;            // F
    unsigned int *extent;   // 8
    int num_blocks;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t carve(BuildScratch *s, int F, char *base)
{
    int nb = (F + SORT_TILE - 1) / SORT_TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return base ? base + o : (char *)0; };
    char *p;
    p = take((size_t)F * 6 * 4); if (s) s->eaabb = (float *)p;
    for (int k = 0; k < 2; ++k) { p = take((size_t)F * 4); if (s) s->keys[k] = (unsigned int *)p; }
    for (int k = 0; k < 2; ++k) { p = take((size_t)F * 4); if (s) s->vals[k] = (int *)p; }
    p = take((size_t)256 * nb * 4); if (s) s->hist = (unsigned int *)p;
    p = take(512 * 4); if (s) s->bin_base = (unsigned int *)p;
    p = take(64); if (s) s->done = (unsigned int *)p;
    p = take((size_t)(2 * F) * 4); if (s) s->parent = (int *)p;
    p = take((size_t)F * 4); if (s) s->visits = (int *)p;
    p = take(64); if (s) s->extent = (unsigned int *)p;
    if (s) s->num_blocks = nb;
    return off;
}

static int sort_pairs(BuildScratch &s, int F, cudaStream_t st)
{
    // 4 passes; result ends in keys[0]/vals[0]
    zero_async(s.done, sizeof(unsigned int), st);
    for (int pass = 0; pass < 4; ++pass) {
        int in = pass & 1, out = in ^ 1;
        k_sort_hist<<<s.num_blocks, SORT_THREADS, 0, st>>>(s.keys[in], F, 8 * pass, s.hist, s.num_blocks);
        k_sort_scan<<<256, 32, 0, st>>>(s.hist, s.num_blocks, s.bin_base, s.done);
        k_sort_scatter<<<s.num_blocks, SORT_THREADS, 0, st>>>(s.keys[in], s.vals[in], F, 8 * pass, s.hist, s.num_blocks,
                                                             s.bin_base, s.keys[out], s.vals[out]);
    }
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_abi_version(void) { return MIRRES_ABI_VERSION; }

size_t mirres_bvh_scratch_bytes(int F) { return F < 1 ? 0 : carve(nullptr, F, nullptr); }
size_t mirres_bvh_packed_node_bytes(int F) { return F < 1 ? 0 : sizeof(PackedNode) * (size_t)(F > 1 ? F - 1 : 1); }
size_t mirres_bvh_packed_tri_bytes(int F) { return F < 1 ? 0 : sizeof(PackedTri) * (size_t)F; }

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
    const int grid = (F + 255) / 256;
    k_init_extent<<<1, 32, 0, st>>>(s.extent);
    zero_async(s.visits, sizeof(int) * (size_t)F, st);
    k_elements<<<grid, 256, 0, st>>>(vert, tri, F, s.eaabb, nullptr, s.extent);
    k_morton<<<grid, 256, 0, st>>>(s.eaabb, F, s.extent, 0, 0, 0, 0, 0, 0, s.keys[0], s.vals[0], nullptr);
    int rc = sort_pairs(s, F, st);
    if (rc) return rc;
    if (sorted_codes) k_zip_pairs<<<grid, 256, 0, st>>>(s.keys[0], s.vals[0], F, sorted_codes);
    k_hierarchy<<<grid, 256, 0, st>>>(F, s.keys[0], s.vals[0], s.eaabb, info, aabb, s.parent);
    k_refit<<<grid, 256, 0, st>>>(F, info, aabb, s.parent, s.visits);
    MR_CUDA_CHECK_LAUNCH();
    if (packed_nodes && packed_tris) {
        PackParams pp = {F, info, aabb, vert, tri, (PackedNode *)packed_nodes, (PackedTri *)packed_tris};
        return foreach_item<PackParams, pack_item, 256>(pp, F, st);
    }
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

// Stable sort of (code, idx) pairs by code; result in `pairs` (the reference's g_elements_in).
int mirres_bvh_sort(int *pairs, int F, void *scratch, size_t scratch_bytes, void *stream)
{
    if (!pairs || !scratch) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    if (scratch_bytes < carve(nullptr, F, nullptr)) return MIRRES_ERR_SCRATCH;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    carve(&s, F, (char *)scratch);
    const int grid = (F + 255) / 256;
    k_unzip_pairs<<<grid, 256, 0, st>>>(pairs, F, s.keys[0], s.vals[0]);
    int rc = sort_pairs(s, F, st);
    if (rc) return rc;
    k_zip_pairs<<<grid, 256, 0, st>>>(s.keys[0], s.vals[0], F, pairs);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

// hierarchy + refit from sorted pairs and element boxes (reference kernels hierarchy, get_bbox*, set_root)
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
    k_unzip_pairs<<<grid, 256, 0, st>>>(sorted_pairs, F, s.keys[0], s.vals[0]);
    zero_async(s.visits, sizeof(int) * (size_t)F, st);
    k_hierarchy<<<grid, 256, 0, st>>>(F, s.keys[0], s.vals[0], ele_aabb, info, aabb, s.parent);
    k_refit<<<grid, 256, 0, st>>>(F, info, aabb, s.parent, s.visits);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}

} // extern "C"
