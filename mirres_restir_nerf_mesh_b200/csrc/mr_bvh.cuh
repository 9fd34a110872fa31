// mirres-b200 BVH traversal for sm_100a.
//
// The reference walks its LBVH with one 16-byte stack record per node, six scalar AABB loads per
// pop and no early-out for boolean rays (nerf/ScreenSpaceReSTIR/utils/helperDi.slang:136-395).
// Here the hierarchy is re-packed at build time into 128-byte WIDE records -- for every internal
// node the boxes of its (up to four) GRANDCHILDREN, listed in the order the reference would visit
// them -- and 48-byte leaf-ordered triangle records (edges pre-subtracted).  One record fetch
// (seven independent LDG.128) replaces up to three dependent fetches of the binary walk, which
// halves the chain of dependent L2 loads that bounds a ray's latency.  Two traversals:
//
//   any_hit      boolean rays (shadow / visibility).  helperDi.slang:197-274 returns any_hit = "some
//                leaf triangle was line-hit before pruning could start", which is independent of the
//                visit order, so the walk stops at the first hit.
//   closest_hit  rays whose t / normal are consumed (helperDi.slang:313-395).  The result depends on the
//                visit order (negative-t hits, ties), so the walk keeps the reference order: right
//                subtree before left subtree at every level, a deferred subtree is re-tested against
//                the current closest distance when it is popped.
//
// Why skipping a level is exact.  The reference enters a node iff min(closest, t_far) > t_near for the
// node's box.  A child box is contained in its parent's box (refit takes exact min/max), and the slab
// arithmetic (subtract, multiply by the same 1/d, min/max) is monotone in fp32, so t_near(child) >=
// t_near(parent) and t_far(child) <= t_far(parent): a grandchild that passes implies its parent passes,
// a parent that fails implies its children fail.  `closest` does not change between the visit of a node
// and the visit of its right child, and a deferred entry is re-tested with the closest distance at pop
// time exactly as the reference re-tests a popped node.  Entry distances do not depend on `closest`, so
// they are computed once, when the record is fetched, and kept on the stack.
//
// Bug-compatible details kept on purpose: no t-range test on triangles, `t_max <= t_min` rejects,
// zero direction components become 1e-6, the direction is re-normalised on entry.
#pragma once
#include "mr_common.cuh"

namespace mr {

// One 32-byte sector, fetched with a single 256-bit load (LDG.E.256, new on sm_100): a divergent warp pays one L1
// tag lookup per lane and instruction whatever the access width, and the shadow-ray kernels are bound by exactly that
// (ncu: l1tex throughput 87 %), so a 128-byte record costs 4 lookups instead of 7-8 with 128-bit loads.
struct alignas(32) Rec32 {
    float v[8];
};
MR_DEV Rec32 load_rec(const Rec32 *p)
{
#if defined(__CUDA_ARCH__)
    Rec32 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
#else
    return *p;
#endif
}

// entries in reference visit order: expand(right child) then expand(left child), expand(X) = [X] if X is a leaf,
// else [right(X), left(X)].  Unused entries hold an empty box (+inf, -inf) and can never pass the slab test.
//
// A reference to an internal node carries, in bits 28-29, how many of the node's trailing entries are unused (0-2): near
// the leaves most records hold two or three boxes, and the traversal kernels are bound by L1 wavefronts (one per lane
// and 32-byte sector), so a walker that knows the count BEFORE the fetch skips the sectors that hold nothing.
struct alignas(32) PackedNode {
    Rec32 e[4]; // entry k: min.xyz, max.xyz, ref bits (>= 0 internal node index, < 0: ~leaf slot), 0
};
struct alignas(32) PackedTri {
    Rec32 a; // v0.xyz, primitive index bits, e1.xyz, 0
    Rec32 b; // e2.xyz, 0 ...
};

#define MR_REF_NODE_MASK 0x0fffffff
MR_DEV int ref_node(int ref) { return ref & MR_REF_NODE_MASK; }       // ref >= 0 only
MR_DEV int ref_missing(int ref) { return (ref >> 28) & 3; }           // unused trailing entries of the referenced record

// The first wide levels of the tree as a table of their own, in front of the node array (north-star item "top BVH levels
// staged in shared memory"): entry 0 is the root record, the children follow breadth first; references between table
// entries carry MR_REF_TOP and the table slot instead of a node index, so a walker that has copied the table into shared
// memory follows them without touching global memory.  Five wide levels = ten binary levels: at most 1 + 4 + 16 + 64 + 256
// records, 43 KB.  Walkers that do not stage the table start at node 0 and never meet a table reference.
#define MR_REF_TOP 0x08000000
#define MR_TOP_LEVELS 5
#define MR_TOP_MAX 341
struct alignas(32) TopTable {
    int count;      // records in use (0: no table, e.g. host-check flavour)
    int pad[31];
    Rec32 rec[MR_TOP_MAX * 4];
};
#define MR_TOP_BYTES ((size_t)sizeof(TopTable))

struct BvhView {
    const PackedNode *__restrict__ nodes; // [max(F-1,1)]
    const PackedTri *__restrict__ tris;   // [F], leaf (sorted-Morton) order
    const TopTable *__restrict__ top;     // in front of the nodes in the packed buffer
};
// the packed node buffer of the C ABI: [TopTable][PackedNode x max(F-1,1)]
static inline BvhView bvh_view(const void *packed_nodes, const void *packed_tris)
{
    BvhView v;
    v.top = (const TopTable *)packed_nodes;
    v.nodes = (const PackedNode *)((const char *)packed_nodes + MR_TOP_BYTES);
    v.tris = (const PackedTri *)packed_tris;
    return v;
}

// ---- building the traversal records from the reference-layout tensors ---------------------------------
struct PackParams {
    int F;
    const int *__restrict__ info;   // [2F-1,3]
    const float *__restrict__ aabb; // [2F-1,6]
    const float *__restrict__ vert; // [V,3]
    const int *__restrict__ tri;    // [F,3]
    PackedNode *__restrict__ nodes;
    PackedTri *__restrict__ tris;
    TopTable *__restrict__ top; // its record count is cleared here; k_top_table fills the table when it is wanted
};
MR_DEV void pack_item(const PackParams &p, int gid)
{
    const int F = p.F, LEAF = F - 1;
    if (gid == 0 && p.top) p.top->count = 0;
    if (p.tris) { // null: the caller has written the triangle records itself (mirres_bvh_build does, with the leaf records)
        int prim = MR_LDG(p.info + 3 * (size_t)(LEAF + gid) + 2);
        int i0 = MR_LDG(p.tri + 3 * (size_t)prim), i1 = MR_LDG(p.tri + 3 * (size_t)prim + 1), i2 = MR_LDG(p.tri + 3 * (size_t)prim + 2);
        float3 v0 = load3(p.vert, (size_t)i0), v1 = load3(p.vert, (size_t)i1), v2 = load3(p.vert, (size_t)i2);
        float3 e1 = v1 - v0, e2 = v2 - v0;
        PackedTri t;
        t.a.v[0] = v0.x; t.a.v[1] = v0.y; t.a.v[2] = v0.z; t.a.v[3] = bits_float(prim);
        t.a.v[4] = e1.x; t.a.v[5] = e1.y; t.a.v[6] = e1.z; t.a.v[7] = 0.f;
        t.b.v[0] = e2.x; t.b.v[1] = e2.y; t.b.v[2] = e2.z; t.b.v[3] = 0.f;
        t.b.v[4] = 0.f; t.b.v[5] = 0.f; t.b.v[6] = 0.f; t.b.v[7] = 0.f;
        p.tris[gid] = t;
    }
    if (gid >= F - 1 && !(F == 1 && gid == 0)) return;
    const float inf = bits_float(0x7f800000);
    float box[24];
    int ref[4] = {0, 0, 0, 0};
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        box[6 * k] = inf; box[6 * k + 1] = inf; box[6 * k + 2] = inf;
        box[6 * k + 3] = -inf; box[6 * k + 4] = -inf; box[6 * k + 5] = -inf;
    }
    auto put = [&](int node) {
        const float *b = p.aabb + 6 * (size_t)node;
#pragma unroll
        for (int k = 0; k < 6; ++k) box[6 * cnt + k] = MR_LDG(b + k);
        if (node < LEAF) {
            // entries of the referenced record: one per leaf child, two per internal child
            const int l = MR_LDG(p.info + 3 * (size_t)node), rr = MR_LDG(p.info + 3 * (size_t)node + 1);
            const int missing = (l >= LEAF ? 1 : 0) + (rr >= LEAF ? 1 : 0);
            ref[cnt] = node | (missing << 28);
        } else {
            ref[cnt] = ~(node - LEAF);
        }
        ++cnt;
    };
    auto expand = [&](int node) {
        if (node >= LEAF) { put(node); return; }
        put(MR_LDG(p.info + 3 * (size_t)node + 1)); // right child is visited first
        put(MR_LDG(p.info + 3 * (size_t)node));
    };
    if (F == 1) {
        put(LEAF); // a single triangle: the root is the leaf
    } else {
        expand(MR_LDG(p.info + 3 * (size_t)gid + 1));
        expand(MR_LDG(p.info + 3 * (size_t)gid));
    }
    PackedNode n;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 6; ++j) n.e[k].v[j] = box[6 * k + j];
        n.e[k].v[6] = bits_float(ref[k]);
        n.e[k].v[7] = 0.f;
    }
    p.nodes[gid] = n;
}

// Depth of the traversal stacks: the reference's 64 (helperDi.slang:136), which it never checks.  Here the arrays carry
// MR_STACK_SLACK spare entries -- a visit of a wide record defers at most three -- and after every visit a stack that
// has grown beyond 64 is cut back: entries are DROPPED instead of written out of bounds (one compare per visit; a
// compare per push cost the boolean-ray walker 9 registers and 8 % of its speed), and the queue tracers raise the
// workspace's error word
// (MIRRES_WORKSPACE_ERROR_BYTES, include/mirres_b200.h) so that the caller can tell that a launch lost a subtree.  The deepest
// stack of the BASELINE scenes is 22 entries (profiles/oracle_counters_*.json); only adversarial meshes (thousands of
// triangles with one Morton code) reach 64.
#define MR_STACK 64
#define MR_STACK_SLACK 4

// traversal records + top table from reference-layout tensors (trace.cu); packed_tris may be null when the caller has
// written the triangle records itself
int tuning_value(int key); // wave.cu: the calling thread's mirres_set_tuning value
int pack_traversal(int F, const int *info, const float *aabb, const float *vert, const int *tri, void *packed_nodes,
                   void *packed_tris, cudaStream_t st);

struct Ray {
    float3 o, d, inv;
};

MR_DEV Ray make_ray(float3 o, float3 d)
{
    Ray r;
    r.o = o;
    r.d = normalize(d);
    float dx = r.d.x == 0.f ? 0.000001f : r.d.x;
    float dy = r.d.y == 0.f ? 0.000001f : r.d.y;
    float dz = r.d.z == 0.f ? 0.000001f : r.d.z;
    r.inv = make_float3(1.0f / dx, 1.0f / dy, 1.0f / dz);
    return r;
}

// slab interval of one box: entry = max(0, near planes), exit = min(far planes)
MR_DEV void slab(const Ray &r, float minx, float miny, float minz, float maxx, float maxy, float maxz, float &tnear,
                 float &tfar)
{
    float ax = (minx - r.o.x) * r.inv.x, bx = (maxx - r.o.x) * r.inv.x;
    float ay = (miny - r.o.y) * r.inv.y, by = (maxy - r.o.y) * r.inv.y;
    float az = (minz - r.o.z) * r.inv.z, bz = (maxz - r.o.z) * r.inv.z;
    float nx = r.inv.x < 0.0f ? bx : ax, fx = r.inv.x < 0.0f ? ax : bx;
    float ny = r.inv.y < 0.0f ? by : ay, fy = r.inv.y < 0.0f ? ay : by;
    float nz = r.inv.z < 0.0f ? bz : az, fz = r.inv.z < 0.0f ? az : bz;
    tnear = fmaxf(fmaxf(fmaxf(0.0f, nx), ny), nz);
    tfar = fminf(fminf(fx, fy), fz);
}

// Moeller-Trumbore on a packed triangle; returns the line parameter with no range test.
MR_DEV bool tri_test(const Ray &r, const Rec32 &a, const Rec32 &b, float &t, float &u, float &v)
{
    const float epsilon = 1e-15f;
    float3 v0 = make_float3(a.v[0], a.v[1], a.v[2]);
    float3 E1 = make_float3(a.v[4], a.v[5], a.v[6]);
    float3 E2 = make_float3(b.v[0], b.v[1], b.v[2]);
    float3 P = cross(r.d, E2);
    float det = dot(E1, P);
    if (det > -epsilon && det < epsilon) return false;
    float invDet = 1.0f / det;
    float3 T = r.o - v0;
    u = dot(T, P) * invDet;
    if (u < 0 || u > 1) return false;
    float3 Q = cross(T, E1);
    v = dot(r.d, Q) * invDet;
    if (v < 0 || u + v > 1) return false;
    t = dot(E2, Q) * invDet;
    return true;
}

struct TraceStats {
    unsigned int nodes, tris; // wide records fetched, triangles tested
};

// one wide record: entry distances of its four entries
struct WideHit {
    float tn[4], tf[4];
    int ref[4];
};
MR_DEV void wide_slabs(const Ray &r, const Rec32 &e0, const Rec32 &e1, const Rec32 &e2, const Rec32 &e3, WideHit &w)
{
    slab(r, e0.v[0], e0.v[1], e0.v[2], e0.v[3], e0.v[4], e0.v[5], w.tn[0], w.tf[0]);
    slab(r, e1.v[0], e1.v[1], e1.v[2], e1.v[3], e1.v[4], e1.v[5], w.tn[1], w.tf[1]);
    slab(r, e2.v[0], e2.v[1], e2.v[2], e2.v[3], e2.v[4], e2.v[5], w.tn[2], w.tf[2]);
    slab(r, e3.v[0], e3.v[1], e3.v[2], e3.v[3], e3.v[4], e3.v[5], w.tn[3], w.tf[3]);
    w.ref[0] = float_bits(e0.v[6]);
    w.ref[1] = float_bits(e1.v[6]);
    w.ref[2] = float_bits(e2.v[6]);
    w.ref[3] = float_bits(e3.v[6]);
}
MR_DEV Rec32 empty_rec()
{
    Rec32 e;
    const float inf = bits_float(0x7f800000);
    e.v[0] = e.v[1] = e.v[2] = inf;
    e.v[3] = e.v[4] = e.v[5] = -inf;
    e.v[6] = e.v[7] = 0.f;
    return e;
}
// entries 2 and 3 of a record are fetched only when the reference says they are in use
MR_DEV void load_tail(const Rec32 *rec, int missing, Rec32 &e2, Rec32 &e3)
{
    e2 = e3 = empty_rec();
    if (missing < 2) e2 = load_rec(rec + 2);
    if (missing < 1) e3 = load_rec(rec + 3);
}
MR_DEV void wide_fetch(const Ray &r, const PackedNode *nodes, int ref, WideHit &w)
{
    const Rec32 *rec = reinterpret_cast<const Rec32 *>(nodes + ref_node(ref));
    const Rec32 e0 = load_rec(rec), e1 = load_rec(rec + 1), e2 = load_rec(rec + 2), e3 = load_rec(rec + 3);
    wide_slabs(r, e0, e1, e2, e3, w);
}
// record address of a traversal reference: wide node (>= 0) or packed triangle (< 0)
MR_DEV const Rec32 *ref_address(const BvhView &bvh, int ref)
{
    return ref >= 0 ? reinterpret_cast<const Rec32 *>(bvh.nodes + ref_node(ref)) : reinterpret_cast<const Rec32 *>(bvh.tris + (size_t)(~ref));
}
MR_DEV bool tri_test_at(const BvhView &bvh, const Ray &r, int leaf_slot, float &t, float &u, float &v)
{
    const PackedTri *tp = bvh.tris + (size_t)leaf_slot;
    const Rec32 a = load_rec(&tp->a), b = load_rec(&tp->b);
    return tri_test(r, a, b, t, u, v);
}

// Boolean query: true iff the reference's bvh_hit(rayo, rayd, 0, 1e7) returns true.
template <bool STATS>
MR_DEV bool any_hit(const BvhView &bvh, float3 origin, float3 dir, TraceStats *st)
{
    Ray r = make_ray(origin, dir);
    const float t_max = 1e7f;
    int stack[MR_STACK + MR_STACK_SLACK];
    int sp = 0;
    int node = 0;
    for (;;) {
        WideHit w;
        wide_fetch(r, bvh.nodes, node, w);
        if (STATS) st->nodes += 1;
        if (sp > MR_STACK) sp = MR_STACK; // see MR_STACK
        int next = 0;
        bool have = false;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            if (fminf(t_max, w.tf[k]) > w.tn[k]) {
                if (have) stack[sp++] = next;
                next = w.ref[k];
                have = true;
            }
        }
        for (;;) {
            if (!have) {
                if (sp == 0) return false;
                next = stack[--sp];
            }
            have = false;
            if (next >= 0) break;
            float t, u, v;
            if (STATS) st->tris += 1;
            if (tri_test_at(bvh, r, ~next, t, u, v)) return true;
        }
        node = next;
    }
}

struct Hit {
    float t;
    float3 pos;
    float3 normal;
    int prim;
    float bary[2]; // (u, v): weights of the triangle's 2nd and 3rd vertex; the 1st has 1 - u - v
};

// face normal of the recorded triangle, flipped towards the ray origin (helperDi.slang:299-307).
// best_slot can only stay -1 when every hit returned NaN; the reference then keeps float3(1).
MR_DEV void closest_finish(const BvhView &bvh, const Ray &r, int best_slot, float3 &n, int &prim, float *bary_uv = nullptr)
{
    n = f3(1.0f);
    prim = -1;
    if (bary_uv) { bary_uv[0] = 0.f; bary_uv[1] = 0.f; }
    if (best_slot >= 0) {
        const PackedTri *tp = bvh.tris + (size_t)best_slot;
        const Rec32 a = load_rec(&tp->a), b = load_rec(&tp->b);
        float t, u, v;
        tri_test(r, a, b, t, u, v);
        float3 fn = normalize(cross(make_float3(a.v[4], a.v[5], a.v[6]), make_float3(b.v[0], b.v[1], b.v[2])));
        float w = 1.0f - u - v;
        n = u * fn + v * fn + w * fn;
        if (dot(-r.d, n) < 0) n = -n;
        n = normalize(n);
        prim = float_bits(a.v[3]);
        if (bary_uv) { bary_uv[0] = u; bary_uv[1] = v; }
    }
}

// Closest-hit with the reference's visit order and update rules (helperDi.slang:313-395).
template <bool STATS>
MR_DEV bool closest_hit(const BvhView &bvh, float3 origin, float3 dir, Hit &out, TraceStats *st)
{
    Ray r = make_ray(origin, dir);
    float closest = 1e7f;
    bool any = false;
    int best_slot = -1;
    int stack_ref[MR_STACK + MR_STACK_SLACK];
    float stack_t[MR_STACK + MR_STACK_SLACK];
    int sp = 0;
    int node = 0;
    for (;;) {
        WideHit w;
        wide_fetch(r, bvh.nodes, node, w);
        if (STATS) st->nodes += 1;
        if (sp > MR_STACK) sp = MR_STACK; // see MR_STACK
        int next = 0;
        float next_t = 0.f;
        bool have = false;
        // entries that pass now, lowest index = visited next, the others deferred so that they pop in index order
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            if (fminf(closest, w.tf[k]) > w.tn[k]) {
                if (have) {
                    stack_ref[sp] = next;
                    stack_t[sp] = next_t;
                    ++sp;
                }
                next = w.ref[k];
                next_t = w.tn[k];
                have = true;
            }
        }
        for (;;) {
            if (!have) {
                // pop: a deferred subtree is entered only if it still beats the current closest distance
                bool found = false;
                while (sp > 0) {
                    --sp;
                    if (closest > stack_t[sp]) {
                        next = stack_ref[sp];
                        found = true;
                        break;
                    }
                }
                if (!found) goto done;
            }
            have = false;
            if (next >= 0) break;
            const int slot = ~next;
            float t, u, v;
            if (STATS) st->tris += 1;
            if (tri_test_at(bvh, r, slot, t, u, v)) {
                // closest = min(t, closest); the normal follows the latest hit with t <= previous closest
                if (t <= closest) best_slot = slot;
                closest = fminf(t, closest);
                any = true;
            }
        }
        node = next;
    }
done:
    if (!any) {
        out.prim = -1;
        return false;
    }
    out.t = closest;
    out.pos = r.o + closest * r.d;
    closest_finish(bvh, r, best_slot, out.normal, out.prim, out.bary);
    return true;
}

} // namespace mr
