// mirres-b200 BVH traversal for sm_100a.
//
// The reference walks its LBVH with one 16-byte stack record per node, six scalar AABB loads per
// pop and no early-out for boolean rays (nerf/ScreenSpaceReSTIR/utils/helperDi.slang:136-395).
// Here the hierarchy is re-packed at build time into 64-byte "two children per record" nodes
// (four LDG.128) and 48-byte leaf-ordered triangle records (three LDG.128, edges pre-subtracted),
// and two traversals are provided:
//
//   any_hit      boolean rays (shadow / visibility).  helperDi.slang:197-274 returns any_hit = "some
//                leaf triangle was line-hit before pruning could start", which is independent of the
//                visit order, so the walk stops at the first hit.
//   closest_hit  rays whose t / normal are consumed (helperDi.slang:313-395).  The result depends on the
//                visit order (negative-t hits, ties), so the walk keeps the reference order: the right
//                child is visited before the left one, a popped subtree is re-tested against the
//                current closest distance.  The slab test of a child is done when its parent is
//                visited (entry distance kept on the stack) -- identical to testing it when popped,
//                because the entry distance does not depend on the closest distance.
//
// Bug-compatible details kept on purpose: no t-range test on triangles, `t_max <= t_min` rejects,
// zero direction components become 1e-6, the direction is re-normalised on entry.
#pragma once
#include "mr_common.cuh"

namespace mr {

struct alignas(16) PackedNode {
    float4 a; // L.min.x L.min.y L.min.z L.max.x
    float4 b; // L.max.y L.max.z R.min.x R.min.y
    float4 c; // R.min.z R.max.x R.max.y R.max.z
    int4 d;   // left ref, right ref, -, -      ref >= 0: internal node index; ref < 0: ~leaf slot
};

struct BvhView {
    const PackedNode *__restrict__ nodes; // [max(F-1,1)]
    const float4 *__restrict__ tris;      // [3*F]: (v0, prim bits) (e1, 0) (e2, 0), leaf order
};

// ---- building the traversal records from the reference-layout tensors ---------------------------------
struct PackParams {
    int F;
    const int *__restrict__ info;   // [2F-1,3]
    const float *__restrict__ aabb; // [2F-1,6]
    const float *__restrict__ vert; // [V,3]
    const int *__restrict__ tri;    // [F,3]
    PackedNode *__restrict__ nodes;
    float4 *__restrict__ tris;
};
MR_DEV void pack_item(const PackParams &p, int gid)
{
    const int F = p.F, LEAF = F - 1;
    {
        int prim = MR_LDG(p.info + 3 * (size_t)(LEAF + gid) + 2);
        int i0 = MR_LDG(p.tri + 3 * (size_t)prim), i1 = MR_LDG(p.tri + 3 * (size_t)prim + 1), i2 = MR_LDG(p.tri + 3 * (size_t)prim + 2);
        float3 v0 = load3(p.vert, (size_t)i0), v1 = load3(p.vert, (size_t)i1), v2 = load3(p.vert, (size_t)i2);
        float3 e1 = v1 - v0, e2 = v2 - v0;
        p.tris[3 * (size_t)gid] = make_float4(v0.x, v0.y, v0.z, bits_float(prim));
        p.tris[3 * (size_t)gid + 1] = make_float4(e1.x, e1.y, e1.z, 0.f);
        p.tris[3 * (size_t)gid + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
    }
    if (F == 1) {
        // a single triangle: node 0 holds the leaf on the left and an empty box on the right
        const float *b = p.aabb;
        const float inf = bits_float(0x7f800000);
        PackedNode n;
        n.a = make_float4(b[0], b[1], b[2], b[3]);
        n.b = make_float4(b[4], b[5], inf, inf);
        n.c = make_float4(inf, -inf, -inf, -inf);
        n.d = make_int4(~0, ~0, 0, 0);
        p.nodes[0] = n;
        return;
    }
    if (gid >= F - 1) return;
    int l = MR_LDG(p.info + 3 * (size_t)gid), r = MR_LDG(p.info + 3 * (size_t)gid + 1);
    const float *bl = p.aabb + 6 * (size_t)l, *br = p.aabb + 6 * (size_t)r;
    PackedNode n;
    n.a = make_float4(MR_LDG(bl), MR_LDG(bl + 1), MR_LDG(bl + 2), MR_LDG(bl + 3));
    n.b = make_float4(MR_LDG(bl + 4), MR_LDG(bl + 5), MR_LDG(br), MR_LDG(br + 1));
    n.c = make_float4(MR_LDG(br + 2), MR_LDG(br + 3), MR_LDG(br + 4), MR_LDG(br + 5));
    n.d = make_int4(l < LEAF ? l : ~(l - LEAF), r < LEAF ? r : ~(r - LEAF), 0, 0);
    p.nodes[gid] = n;
}

#define MR_STACK 64

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
MR_DEV bool tri_test(const Ray &r, float4 q0, float4 q1, float4 q2, float &t, float &u, float &v)
{
    const float epsilon = 1e-15f;
    float3 v0 = make_float3(q0.x, q0.y, q0.z);
    float3 E1 = make_float3(q1.x, q1.y, q1.z);
    float3 E2 = make_float3(q2.x, q2.y, q2.z);
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
    unsigned int nodes, tris;
};

// Boolean query: true iff the reference's bvh_hit(rayo, rayd, 0, 1e7) returns true.
template <bool STATS>
MR_DEV bool any_hit(const BvhView &bvh, float3 origin, float3 dir, TraceStats *st)
{
    Ray r = make_ray(origin, dir);
    const float t_max = 1e7f;
    int stack[MR_STACK];
    int sp = 0;
    int node = 0;
    for (;;) {
        const PackedNode *pn = bvh.nodes + node;
        float4 a = MR_LDG(&pn->a), b = MR_LDG(&pn->b), c = MR_LDG(&pn->c);
        int4 d = MR_LDG(&pn->d);
        if (STATS) st->nodes += 2;
        float ln, lf, rn, rf;
        slab(r, a.x, a.y, a.z, a.w, b.x, b.y, ln, lf);
        slab(r, b.z, b.w, c.x, c.y, c.z, c.w, rn, rf);
        bool passL = fminf(t_max, lf) > ln;
        bool passR = fminf(t_max, rf) > rn;
        int next;
        bool have = false;
        if (passR) {
            next = d.y;
            have = true;
            if (passL) stack[sp++] = d.x;
        } else if (passL) {
            next = d.x;
            have = true;
        }
        for (;;) {
            if (!have) {
                if (sp == 0) return false;
                next = stack[--sp];
            }
            have = false;
            if (next >= 0) break;
            const float4 *tp = bvh.tris + 3 * (size_t)(~next);
            float t, u, v;
            if (STATS) st->tris += 1;
            if (tri_test(r, MR_LDG(tp), MR_LDG(tp + 1), MR_LDG(tp + 2), t, u, v)) return true;
        }
        node = next;
    }
}

struct Hit {
    float t;
    float3 pos;
    float3 normal;
    int prim;
};

// Closest-hit with the reference's visit order and update rules (helperDi.slang:313-395).
template <bool STATS>
MR_DEV bool closest_hit(const BvhView &bvh, float3 origin, float3 dir, Hit &out, TraceStats *st)
{
    Ray r = make_ray(origin, dir);
    float closest = 1e7f;
    bool any = false;
    int best_slot = -1;
    int stack_ref[MR_STACK];
    float stack_t[MR_STACK];
    int sp = 0;
    int node = 0;
    for (;;) {
        const PackedNode *pn = bvh.nodes + node;
        float4 a = MR_LDG(&pn->a), b = MR_LDG(&pn->b), c = MR_LDG(&pn->c);
        int4 d = MR_LDG(&pn->d);
        if (STATS) st->nodes += 2;
        float ln, lf, rn, rf;
        slab(r, a.x, a.y, a.z, a.w, b.x, b.y, ln, lf);
        slab(r, b.z, b.w, c.x, c.y, c.z, c.w, rn, rf);
        bool passL = fminf(closest, lf) > ln;
        bool passR = fminf(closest, rf) > rn;
        int next;
        bool have = false;
        if (passR) {
            next = d.y;
            have = true;
            if (passL) {
                stack_ref[sp] = d.x;
                stack_t[sp] = ln;
                ++sp;
            }
        } else if (passL) {
            next = d.x;
            have = true;
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
            const float4 *tp = bvh.tris + 3 * (size_t)slot;
            float t, u, v;
            if (STATS) st->tris += 1;
            if (tri_test(r, MR_LDG(tp), MR_LDG(tp + 1), MR_LDG(tp + 2), t, u, v)) {
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
    {
        // face normal of the recorded triangle, flipped towards the ray origin (helperDi.slang:299-307).
        // best_slot can only stay -1 when every hit returned NaN; the reference then keeps float3(1).
        float3 n = f3(1.0f);
        int prim = -1;
        if (best_slot >= 0) {
            const float4 *tp = bvh.tris + 3 * (size_t)best_slot;
            float4 q0 = MR_LDG(tp), q1 = MR_LDG(tp + 1), q2 = MR_LDG(tp + 2);
            float t, u, v;
            tri_test(r, q0, q1, q2, t, u, v);
            float3 fn = normalize(cross(make_float3(q1.x, q1.y, q1.z), make_float3(q2.x, q2.y, q2.z)));
            float w = 1.0f - u - v;
            n = u * fn + v * fn + w * fn;
            if (dot(-r.d, n) < 0) n = -n;
            n = normalize(n);
            prim = float_bits(q0.w);
        }
        out.normal = n;
        out.prim = prim;
    }
    return true;
}

} // namespace mr
