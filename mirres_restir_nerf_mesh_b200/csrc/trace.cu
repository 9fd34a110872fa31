// mirres-b200: standalone ray queries and traversal-record packing (also the G-buffer producer of the
// synthetic benchmark).  Semantics: nerf/ScreenSpaceReSTIR/utils/helperDi.slang:197-274 (bvh_hit) and
// :313-395 (bvh_hit_with_normal) with t_min = 0, t_max = 1e7; the primitive id and visit counters are
// extensions the reference does not output.
#include "mr_bvh.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

struct TraceParams {
    BvhView bvh;
    const float *__restrict__ org; // [n,3]
    const float *__restrict__ dir; // [n,3]
    int *__restrict__ hit;         // [n]
    float *__restrict__ t;         // [n] or null
    float *__restrict__ pos;       // [n,3] or null
    float *__restrict__ normal;    // [n,3] or null
    int *__restrict__ prim;        // [n] or null
    unsigned int *__restrict__ visits; // [n,2] nodes, triangles -- or null
};

MR_DEV void trace_closest_item(const TraceParams &p, int idx)
{
    const size_t i = (size_t)idx;
    Hit h;
    h.t = 0.f;
    h.pos = f3(0.f);
    h.normal = f3(1.f);
    h.prim = -1;
    TraceStats st = {0u, 0u};
    bool found = p.visits ? closest_hit<true>(p.bvh, load3(p.org, i), load3(p.dir, i), h, &st)
                          : closest_hit<false>(p.bvh, load3(p.org, i), load3(p.dir, i), h, nullptr);
    p.hit[i] = found ? 1 : 0;
    if (p.t) p.t[i] = found ? h.t : 0.f;
    if (p.pos) store3(p.pos, i, found ? h.pos : f3(0.f));
    if (p.normal) store3(p.normal, i, found ? h.normal : f3(1.f));
    if (p.prim) p.prim[i] = found ? h.prim : -1;
    if (p.visits) { p.visits[2 * i] = st.nodes; p.visits[2 * i + 1] = st.tris; }
}

MR_DEV void trace_any_item(const TraceParams &p, int idx)
{
    const size_t i = (size_t)idx;
    TraceStats st = {0u, 0u};
    bool found = p.visits ? any_hit<true>(p.bvh, load3(p.org, i), load3(p.dir, i), &st)
                          : any_hit<false>(p.bvh, load3(p.org, i), load3(p.dir, i), nullptr);
    p.hit[i] = found ? 1 : 0;
    if (p.visits) { p.visits[2 * i] = st.nodes; p.visits[2 * i + 1] = st.tris; }
}

#if !defined(MR_HOST_CHECK)
// breadth-first copy of the first MR_TOP_LEVELS wide levels into the table in front of the node array; references from a
// table entry to a table entry are rewritten to MR_REF_TOP | slot (see TopTable)
__global__ void __launch_bounds__(256) k_top_table(const PackedNode *__restrict__ nodes, TopTable *top)
{
    __shared__ int idx[MR_TOP_MAX];
    __shared__ int count;
    if (threadIdx.x == 0) {
        idx[0] = 0;
        count = 1;
    }
    __syncthreads();
    const float inf = __int_as_float(0x7f800000);
    int start = 0;
    for (int level = 0; level < MR_TOP_LEVELS; ++level) {
        const int end = count;
        __syncthreads();
        for (int w = threadIdx.x; w < (end - start) * 4; w += 256) {
            const int s = start + (w >> 2), k = w & 3;
            Rec32 e = nodes[idx[s]].e[k];
            int ref = __float_as_int(e.v[6]);
            const bool used = !(e.v[0] == inf); // unused entries hold an empty box
            if (level < MR_TOP_LEVELS - 1 && used && ref >= 0) {
                const int c = atomicAdd(&count, 1);
                idx[c] = ref_node(ref);
                e.v[6] = __int_as_float(MR_REF_TOP | c | (ref & 0x30000000));
            }
            top->rec[s * 4 + k] = e;
        }
        __syncthreads();
        start = end;
    }
    if (threadIdx.x == 0) top->count = count;
}
#endif

int pack_traversal(int F, const int *info, const float *aabb, const float *vert, const int *tri, void *packed_nodes,
                   void *packed_tris, cudaStream_t st)
{
    TopTable *top = (TopTable *)packed_nodes;
    PackedNode *nodes = (PackedNode *)((char *)packed_nodes + MR_TOP_BYTES);
    PackParams pp = {F, info, aabb, vert, tri, nodes, (PackedTri *)packed_tris, top};
    int rc = foreach_item<PackParams, pack_item, 256>(pp, packed_tris ? F : (F > 1 ? F - 1 : 1), st);
    if (rc) return rc;
#if !defined(MR_HOST_CHECK)
    // the table is built only for callers that have switched the staged walker on (it is off by default: measured slower,
    // profiles/README.md); an empty table makes that walker start at the root record in global memory
    if (tuning_value(MIRRES_TUNE_ANY_TOP) == 1) {
        k_top_table<<<1, 256, 0, st>>>(nodes, top);
        MR_CUDA_CHECK_LAUNCH();
    }
#endif
    return 0;
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_trace_closest(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                         int *hit, float *t, float *pos, float *normal, int *prim, unsigned int *visits, void *stream)
{
    if (!packed_nodes || !packed_tris || !org || !dir || !hit) return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    TraceParams p = {bvh_view(packed_nodes, packed_tris), org, dir, hit, t, pos, normal, prim, visits};
    return foreach_item<TraceParams, trace_closest_item, 128>(p, n, (cudaStream_t)stream);
}

int mirres_trace_any(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                     int *hit, unsigned int *visits, void *stream)
{
    if (!packed_nodes || !packed_tris || !org || !dir || !hit) return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    TraceParams p = {bvh_view(packed_nodes, packed_tris), org, dir, hit, nullptr, nullptr, nullptr, nullptr, visits};
    return foreach_item<TraceParams, trace_any_item, 128>(p, n, (cudaStream_t)stream);
}

// Traversal records from reference-layout tensors (any LBVH with leaves at [F-1,2F-2], root 0).
int mirres_bvh_pack(const int *info, const float *aabb, const float *vert, const int *tri, int F, void *packed_nodes,
                    void *packed_tris, void *stream)
{
    if (!info || !aabb || !vert || !tri || !packed_nodes || !packed_tris) return MIRRES_ERR_NULL;
    if (F < 1) return MIRRES_ERR_SHAPE;
    if (((uintptr_t)packed_nodes & 31) || ((uintptr_t)packed_tris & 31)) return MIRRES_ERR_ALIGN;
    return pack_traversal(F, info, aabb, vert, tri, packed_nodes, packed_tris, (cudaStream_t)stream);
}

} // extern "C"
