// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).  PARITY UNPINNED.
//
// orc_bvh.h : LBVH construction and stack traversal.
//   nerf/bvhworkers/get_elements.slang:3-39          generateElements
//   nerf/bvhworkers/lbvh_morton_codes.slang:24-79    expandBits / morton3D / morton_codes
//   nerf/bvhworkers/lbvh_single_radixsort.slang:28-138   (semantics = stable sort by code)
//   nerf/bvhworkers/lbvh_hierarchy.slang:31-245      delta / determineRange / findSplit / hierarchy
//   nerf/bvhworkers/lbvh_bounding_boxes.slang:151-390  get_bvh_height / get_bbox / set_root
//   nerf/ScreenSpaceReSTIR/utils/helperDi.slang:136-395  aabb_hit / triangle_hit / bvh_hit(_with_normal)
#ifndef ORC_BVH_H
#define ORC_BVH_H

#include "orc_common.h"

namespace orc {

struct Bvh {
    const int *info;   // [2F-1,3] left,right,prim
    const float *aabb; // [2F-1,6]
    const float *vert; // [V,3]
    const int *tri;    // [F,3]
};

struct TraceCounters {
    // reference schedule: full DFS with closest-so-far pruning (helperDi.slang:197-274)
    long long nodes_ref, tris_ref;
    // contract schedule for boolean rays: same DFS, stop at the first triangle hit
    long long nodes_any, tris_any;
    int max_stack;
};

// helperDi.slang:149-170
static inline bool aabb_hit(f3 o, f3 d, float t_min, float t_max, const float *bb)
{
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int i = 0; i < 3; ++i) {
        float ray_d_i = dd[i];
        if (ray_d_i == 0.f) ray_d_i = 0.000001f;
        float inverse_direction = 1.0f / ray_d_i;
        float t0 = (bb[i] - oo[i]) * inverse_direction;
        float t1 = (bb[3 + i] - oo[i]) * inverse_direction;
        if (inverse_direction < 0.0f) { float tmp = t1; t1 = t0; t0 = tmp; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    return true;
}

// helperDi.slang:172-195 and :277-310 (the normal is only formed when want_normal)
static inline bool triangle_hit(f3 o, f3 d, f3 v0, f3 v1, f3 v2, float &t_hit, f3 *normal)
{
    const float epsilon = 1e-15f;
    f3 E1 = v1 - v0;
    f3 E2 = v2 - v0;
    f3 P = cross(d, E2);
    float det = dot(E1, P);
    if (det > -epsilon && det < epsilon) return false;
    float invDet = 1.0f / det;
    f3 T = o - v0;
    float u = dot(T, P) * invDet;
    if (u < 0 || u > 1) return false;
    f3 Q = cross(T, E1);
    float v = dot(d, Q) * invDet;
    if (v < 0 || u + v > 1) return false;
    float t = dot(E2, Q) * invDet;
    t_hit = t;
    if (normal) {
        f3 fn = normalize(cross(E1, E2));
        float r = 1.0f - u - v;
        f3 n = u * fn + v * fn + r * fn;
        if (dot(-d, n) < 0) n = -n;
        *normal = normalize(n);
    }
    return true;
}

// helperDi.slang:197-274 (bvh_hit) and :313-395 (bvh_hit_with_normal).  The 64-entry stack of the
// reference has no overflow check; the oracle uses 256 entries and reports the depth reached.
static inline bool bvh_hit(const Bvh &b, f3 rayo, f3 rayd, float t_min, float t_max, float &t_hit, f3 &pos,
                           f3 *normal, int *prim_out, TraceCounters *tc)
{
    rayd = normalize(rayd);
    int stack[256];
    int count = 0;
    stack[count++] = 0;
    float closest_so_far = t_max;
    bool any_hit = false;
    float hit_t = 0.f;
    f3 hit_pos = mk3(0.f);
    f3 hit_normal = mk3(1.f);
    int hit_prim = -1;
    long long nodes = 0, tris = 0;
    while (count > 0) {
        int node = stack[--count];
        ++nodes;
        if (!aabb_hit(rayo, rayd, t_min, closest_so_far, b.aabb + 6 * (size_t)node)) continue;
        int left = b.info[3 * (size_t)node + 0], right = b.info[3 * (size_t)node + 1];
        if (left != 0 && right != 0) {
            if (count + 2 > 256) return any_hit; // cannot happen for the tested scenes
            stack[count++] = left;
            stack[count++] = right;
            if (tc && count > tc->max_stack) tc->max_stack = count;
        } else if (left == 0 && right == 0) {
            int prim = b.info[3 * (size_t)node + 2];
            const int *vi = b.tri + 3 * (size_t)prim;
            f3 v0 = mk3(b.vert[3 * (size_t)vi[0]], b.vert[3 * (size_t)vi[0] + 1], b.vert[3 * (size_t)vi[0] + 2]);
            f3 v1 = mk3(b.vert[3 * (size_t)vi[1]], b.vert[3 * (size_t)vi[1] + 1], b.vert[3 * (size_t)vi[1] + 2]);
            f3 v2 = mk3(b.vert[3 * (size_t)vi[2]], b.vert[3 * (size_t)vi[2] + 1], b.vert[3 * (size_t)vi[2] + 2]);
            float now_t = 0.f;
            f3 now_n = mk3(1.f);
            ++tris;
            bool hit = triangle_hit(rayo, rayd, v0, v1, v2, now_t, normal ? &now_n : 0);
            closest_so_far = hit ? smin(now_t, closest_so_far) : closest_so_far;
            if (hit) {
                if (!any_hit && tc) { tc->nodes_any += nodes; tc->tris_any += tris; }
                any_hit = true;
                hit_t = closest_so_far;
                hit_pos = rayo + hit_t * rayd;
                if (now_t <= closest_so_far) { hit_normal = now_n; hit_prim = prim; }
            }
        }
    }
    if (tc) {
        tc->nodes_ref += nodes;
        tc->tris_ref += tris;
        if (!any_hit) { tc->nodes_any += nodes; tc->tris_any += tris; }
    }
    if (any_hit) {
        t_hit = hit_t;
        pos = hit_pos;
        if (normal) *normal = hit_normal;
    }
    if (prim_out) *prim_out = any_hit ? hit_prim : -1;
    return any_hit;
}

} // namespace orc
#endif
