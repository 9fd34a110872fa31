// mirres-b200: screen-space ReSTIR kernels for sm_100a.
//
// One thread per pixel, pixelIndex = y * framedim_x + x, RNG keyed on the global pixel coordinate
// (so any tiling / sharding of the frame reproduces the single-GPU image bit for bit).
// Replaces:
//   nerf/ScreenSpaceReSTIR/InitialResampling.slang:151-295     -> k_initial
//   nerf/ScreenSpaceReSTIR/TemporalResampling.slang:23-135     -> k_temporal
//   nerf/ScreenSpaceReSTIR/SpatialResampling.slang:178-322     -> k_spatial
//   nerf/ScreenSpaceReSTIR/EvaluateFinalSamples.slang:84-124   -> k_final_visibility
//   nerf/ScreenSpaceReSTIR/EvaluateFinalSamples.slang:129-188  -> k_eval_final_fwd / k_eval_final_bwd (hand-derived)
//   nerf/ScreenSpaceReSTIR/utils/res.slang:53-232              -> the streaming RIS steps inlined below
// Visibility rays use mr::any_hit (first-hit exit), see mr_bvh.cuh for why that is result-identical.
#include "mr_wave.cuh"
#include "mr_light.cuh"
#include "mr_brdf.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

struct ResView {
    float *ld;  // [N,3] valid flag, oct.u, oct.v
    float *pdf; // [N]
    int *M;     // [N]
    float *w;   // [N]
};
struct ResConst {
    const float *__restrict__ ld;
    const float *__restrict__ pdf;
    const int *__restrict__ M;
    const float *__restrict__ w;
};
struct Reservoir {
    float3 ld;
    float pdf;
    int M;
    float w;
};
struct Ris {
    float3 ld;
    float pdf;
    float wsum, M, weight, canonical;
};
MR_DEV Ris ris_empty()
{
    Ris s;
    s.ld = f3(0.f);
    s.pdf = 0.f;
    s.wsum = 0.f; s.M = 0.f; s.weight = 0.f; s.canonical = 0.f;
    return s;
}
MR_DEV void res_zero(const ResView &r, size_t i)
{
    store3(r.ld, i, f3(0.f));
    r.pdf[i] = 0.f; r.M[i] = 0; r.w[i] = 0.f;
}
MR_DEV void res_store(const ResView &r, size_t i, const Ris &s)
{
    if (isinf(s.weight) || isnan(s.weight)) { res_zero(r, i); return; }
    store3(r.ld, i, s.ld);
    r.pdf[i] = s.pdf;
    r.M[i] = to_int(s.M);
    r.w[i] = s.weight;
}
MR_DEV Reservoir res_load(const ResConst &r, size_t i)
{
    Reservoir o;
    o.ld = load3(r.ld, i);
    o.pdf = MR_LDG(r.pdf + i);
    o.M = MR_LDG(r.M + i);
    o.w = MR_LDG(r.w + i);
    return o;
}
MR_DEV Reservoir res_load_rw(const ResView &r, size_t i)
{
    Reservoir o;
    o.ld = load3_rw(r.ld, i);
    o.pdf = r.pdf[i];
    o.M = r.M[i];
    o.w = r.w[i];
    return o;
}
MR_DEV bool ris_step_reservoir(Ris &st, const Reservoir &r, float targetPdf, uint32_t &sg)
{
    float sampleWeight = targetPdf * r.w * (float)r.M;
    st.wsum += sampleWeight;
    st.M += (float)r.M;
    bool sel = rnd(sg) * st.wsum < sampleWeight;
    if (sel) { st.ld = r.ld; st.pdf = r.pdf; st.weight = targetPdf; }
    return sel;
}
MR_DEV float m_factor(float q0, float q1) { return q0 == 0.f ? 1.f : clampf(mr_pow8f(fminf(q1 / q0, 1.f)), 0.f, 1.f); }
MR_DEV float pairwise_mis(float q0, float q1, float N0, float N1) { return (q1 == 0.f) ? 0.f : (N0 * q0) / (q0 * N0 + q1 * N1); }
MR_DEV bool neighbor_ok(float3 n0, float d0, float3 n1, float d1) { return dot(n0, n1) >= 0.5f && fabsf(d0 - d1) <= 0.1f * d0; }

struct GBuf {
    const float *__restrict__ occ;          // [N]
    const float *__restrict__ normal_depth; // [N,4]
    const float *__restrict__ brdf;         // [N,3]
    const float *__restrict__ ray_dir;      // [N,3]
};
MR_DEV float4 load_nd(const float *__restrict__ nd, size_t i) { return MR_LDG(reinterpret_cast<const float4 *>(nd) + i); }

#define VIS_NEAR 0.01f

// ------------------------------------------------------------------------------------------------------------------
struct InitialParams {
    BvhView bvh;
    EnvView env;
    GBuf g;
    const float *__restrict__ pos_map;
    ResView res;
    const float *__restrict__ light_data;
    const float *__restrict__ light_pdf;
    const float4 *__restrict__ light_cache; // optional (direction, radiance) per tile slot, see mirres_light_tiles
    int fx, fy;
    unsigned int frame;
    unsigned int tile_count, tile_size, screen_tile, n_light, n_brdf;
    unsigned char *vis_tag; // optional, see mirres_set_visibility_tags
    Workspace ws;
};

// gen: the whole RIS stream of one active pixel; stores the reservoir as if the sample were visible and queues the
// visibility ray (InitialResampling.slang:255-270); initial_resolve_px applies the occlusion test result.
MR_DEV void initial_gen_px(const InitialParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    uint32_t tileSg = seed_of(px / p.screen_tile, row_of(p.ws, py) / p.screen_tile, frame_of(p.ws, p.frame));
    uint32_t tileIndex = minu(to_uint(rnd(tileSg) * (float)p.tile_count), p.tile_count - 1u);
    const uint32_t tileOffset = tileIndex * p.tile_size;
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    const uint32_t stride = (p.tile_size + p.n_light - 1u) / p.n_light;
    const uint32_t offset = minu(to_uint(rnd(sg) * (float)stride), stride - 1u);
    float4 nd = load_nd(p.g.normal_depth, i);
    const float3 N = make_float3(nd.x, nd.y, nd.z);
    const float3 rd = load3(p.g.ray_dir, i);
    const float3 brdf = load3(p.g.brdf, i);
    const RisSurface surf = ris_surface(N, rd, brdf);
    const float ratio = (float)p.n_brdf / (float)(p.n_light + p.n_brdf);
    Ris st = ris_empty();
    for (uint32_t k = 0; k < p.n_light; ++k) {
        const uint32_t slot = tileOffset + (offset + k * stride) % p.tile_size;
        const float3 cand = load3(p.light_data, slot);
        const float cand_pdf = MR_LDG(p.light_pdf + slot);
        float3 Le, L;
        if (p.light_cache) {
            const float4 c0 = MR_LDG(p.light_cache + 2 * (size_t)slot), c1 = MR_LDG(p.light_cache + 2 * (size_t)slot + 1);
            L = make_float3(c0.x, c0.y, c0.z);
            Le = make_float3(c1.x, c1.y, c1.z);
        } else {
            light_of(p.env, cand.y, cand.z, Le, L);
        }
        float targetPdf = target_pdf(surf, Le, L);
        float sourcePdf = lerpf(cand_pdf, ris_brdf_pdf(surf, L), ratio);
        float sampleWeight = targetPdf / sourcePdf;
        st.wsum += sampleWeight;
        st.M += 1.f;
        if (rnd(sg) * st.wsum < sampleWeight) { st.ld = cand; st.pdf = cand_pdf; st.weight = targetPdf; }
    }
    for (uint32_t k = 0; k < p.n_brdf; ++k) {
        float x0 = rnd(sg), x1 = rnd(sg), x2 = rnd(sg);
        float3 dir;
        if (!ris_brdf_sample(surf, x0, x1, x2, dir)) { st.M += 1.f; continue; }
        float cand_pdf = env_pdf(p.env, dir);
        float2 o = oct_encode(dir);
        float3 Le = env_radiance(p.env, ngp_dir(dir));
        float targetPdf = target_pdf(surf, Le, dir);
        float sourcePdf = lerpf(cand_pdf, ris_brdf_pdf(surf, dir), ratio);
        float sampleWeight = targetPdf / sourcePdf;
        st.wsum += sampleWeight;
        st.M += 1.f;
        if (rnd(sg) * st.wsum < sampleWeight) { st.ld = make_float3(1.0f, o.x, o.y); st.pdf = cand_pdf; st.weight = targetPdf; }
    }
    if (st.ld.x > 0.1f) {
        float3 L = oct_decode(st.ld.y, st.ld.z);
        queue_ray(p.ws, (size_t)a, load3(p.pos_map, i) + VIS_NEAR * L, L);
    } else {
        queue_empty(p.ws, (size_t)a);
    }
    // the sample that survives this pass has been seen unoccluded from this pixel (the resolve pass takes the tag back
    // together with the sample when the ray hits)
    if (p.vis_tag) p.vis_tag[i] = st.ld.x > 0.1f ? 1 : 0;
    st.weight = st.weight > 0.f ? (st.wsum / st.M) / st.weight : 0.f;
    st.M = 1.f;
    res_store(p.res, i, st);
}

// an occluded sample resets the RIS state before the final weight is formed: (0,0,0), pdf 0, M 1, W 0
MR_DEV void initial_resolve_px(const InitialParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
    if (p.ws.hit[a] != MR_HIT_HIT) return;
    const size_t i = (size_t)p.ws.active[a];
    store3(p.res.ld, i, f3(0.f));
    p.res.pdf[i] = 0.f;
    p.res.M[i] = 1;
    p.res.w[i] = 0.f;
    if (p.vis_tag) p.vis_tag[i] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
struct TemporalParams {
    EnvView env;
    GBuf g, prev_g;
    ResView res;
    ResConst prev;
    const float *__restrict__ motion; // [N,2] or null (zeros)
    int fx, fy;
    unsigned int frame;
    unsigned int max_history;
    unsigned char *vis_tag;            // optional: tag of res (in / out)
    const unsigned char *prev_vis_tag; // optional: tag of prev, valid for the SAME pos_map and BVH as this frame's
    Workspace ws;
};

MR_DEV void temporal_px(const TemporalParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    if (MR_LDG(p.g.occ + i) < 0.1f) return;
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    float u0 = rnd(sg), u1 = rnd(sg);
    float mvx = p.motion ? MR_LDG(p.motion + 2 * i) : 0.f, mvy = p.motion ? MR_LDG(p.motion + 2 * i + 1) : 0.f;
    int ppx = to_int((float)px + mvx * (float)(uint32_t)p.fx + (u0 * 1.f - 0.f));
    // the jittered row is rounded at the magnitude of the FULL frame's row number (int(pixel + u) rounds up to the next
    // pixel for large coordinates), then brought back into the rows this launch was handed
    const uint32_t gy = row_of(p.ws, py);
    int ppy = to_int((float)gy + mvy * (float)(uint32_t)p.fy + (u1 * 1.f - 0.f)) - (int)(gy - py);
    if (ppx >= p.fx || ppx < 0 || ppy >= p.fy || ppy < 0) return;
    const size_t pi = (size_t)ppy * p.fx + ppx;
    if (MR_LDG(p.prev_g.occ + pi) < 0.1f) return;
    float4 nd = load_nd(p.g.normal_depth, i), pnd = load_nd(p.prev_g.normal_depth, pi);
    const float3 N = make_float3(nd.x, nd.y, nd.z), pN = make_float3(pnd.x, pnd.y, pnd.z);
    const RisSurface cur_s = ris_surface(N, load3(p.g.ray_dir, i), load3(p.g.brdf, i));
    const RisSurface prev_s = ris_surface(pN, load3(p.prev_g.ray_dir, pi), load3(p.prev_g.brdf, pi));
    Reservoir cur = res_load_rw(p.res, i);
    Reservoir prev = res_load(p.prev, pi);
    prev.M = (int)minu((uint32_t)prev.M, (uint32_t)cur.M * p.max_history);
    if (!neighbor_ok(N, nd.w, pN, pnd.w)) return;
    Ris st = ris_empty();
    float3 Le, L;
    light_of(p.env, cur.ld.y, cur.ld.z, Le, L);
    ris_step_reservoir(st, cur, target_pdf(cur_s, Le, L), sg);
    light_of(p.env, prev.ld.y, prev.ld.z, Le, L);
    bool usedPrev = ris_step_reservoir(st, prev, target_pdf(cur_s, Le, L), sg);
    light_of(p.env, st.ld.y, st.ld.z, Le, L);
    float currentPdf = target_pdf(cur_s, Le, L);
    float prevPdf = target_pdf(prev_s, Le, L);
    float normalization = (usedPrev ? prevPdf : currentPdf) / ((float)cur.M * currentPdf + (float)prev.M * prevPdf);
    st.weight = st.weight > 0.f ? (st.wsum * normalization) / st.weight : 0.f;
    res_store(p.res, i, st);
    // a history sample brings its tag along only if it comes from this very pixel (same position, same ray); the
    // pixel's own sample keeps the tag it has
    if (p.vis_tag && usedPrev) p.vis_tag[i] = (p.prev_vis_tag && pi == i) ? p.prev_vis_tag[pi] : 0;
}

// ------------------------------------------------------------------------------------------------------------------
struct SpatialParams {
    BvhView bvh;
    EnvView env;
    GBuf g;
    const float *__restrict__ pos_map;
    ResView res;
    ResConst prev;
    const float *__restrict__ offsets; // [count,2] in [-1,1]
    int fx, fy;
    unsigned int frame;
    unsigned int offset_count, neighbor_count;
    float radius;
    unsigned char *vis_tag;            // optional: tag of res (out)
    const unsigned char *prev_vis_tag; // optional: tag of prev
    Workspace ws;
};

MR_DEV bool spatial_neighbor(const SpatialParams &p, uint32_t px, uint32_t py, uint32_t startIndex, uint32_t k, size_t &n)
{
    const uint32_t ni = (startIndex + k) & (p.offset_count - 1u);
    int npx = (int)px + to_int(MR_LDG(p.offsets + 2 * (size_t)ni) * p.radius);
    int npy = (int)py + to_int(MR_LDG(p.offsets + 2 * (size_t)ni + 1) * p.radius);
    if (!(npx >= 0 && npx < p.fx && npy >= 0 && npy < p.fy)) return false;
    n = (size_t)npy * p.fx + npx;
    return true;
}

// gen: picks the neighbours, applies the reuse heuristics and queues the two visibility rays of every accepted one
// (SpatialResampling.slang:229-277): slot 2k = own surface -> neighbour's light, slot 2k+1 = neighbour surface -> own light
#define MR_MAX_NEIGHBORS (MR_MAX_RAYS_PER_PIXEL / 2)
MR_DEV void spatial_gen_px(const SpatialParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
#if defined(__CUDA_ARCH__)
    const unsigned int act = __activemask(); // the lanes of this warp that have a pixel
#endif
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    float4 nd = load_nd(p.g.normal_depth, i);
    const float3 N = make_float3(nd.x, nd.y, nd.z);
    const uint32_t startIndex = to_uint(rnd(sg) * (float)p.offset_count);
    const float3 cur_ld = load3(p.prev.ld, i);
    // The kernel was bound by latency (ncu: issue active 20 %, 26 warps stalled on the long scoreboard per issue): per
    // neighbour a chain of gathers and then two queue tickets, i.e. ten atomic round trips to L2 per warp one after the
    // other.  Here every field of every in-frame neighbour is requested at once, the tests run on registers in the
    // reference's order, and the warp draws ALL its tickets with one atomic; inside the reserved block the rays keep the
    // order the ticket-per-ray version produced (neighbour-major, lanes ascending), so the tracer sees the same queue.
    size_t nb[MR_MAX_NEIGHBORS];
    bool ok[MR_MAX_NEIGHBORS];
    float4 nnd[MR_MAX_NEIGHBORS];
    int nM[MR_MAX_NEIGHBORS];
    float nocc[MR_MAX_NEIGHBORS], nw[MR_MAX_NEIGHBORS];
    float3 nld[MR_MAX_NEIGHBORS], npos[MR_MAX_NEIGHBORS];
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        nb[k] = 0;
        ok[k] = k < p.neighbor_count && spatial_neighbor(p, px, py, startIndex, k, nb[k]);
    }
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        nnd[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        nM[k] = 0;
        nocc[k] = nw[k] = 0.f;
        nld[k] = npos[k] = f3(0.f);
        if (ok[k]) {
            nnd[k] = load_nd(p.g.normal_depth, nb[k]);
            nM[k] = MR_LDG(p.prev.M + nb[k]);
            nw[k] = MR_LDG(p.prev.w + nb[k]);
            nocc[k] = MR_LDG(p.g.occ + nb[k]);
            nld[k] = load3(p.prev.ld, nb[k]);
            npos[k] = load3(p.pos_map, nb[k]);
        }
    }
    // radiance and direction of this pixel's sample, evaluated ONCE per pixel here: the resolve pass needs them for the
    // pixel itself and for each of its neighbours (six env lookups with their trigonometry per pixel otherwise)
    float3 cLe, cL;
    light_of(p.env, cur_ld.y, cur_ld.z, cLe, cL);
    // ... and so is the target density of that sample at its own surface, which every pixel that picks this one as a
    // neighbour needs as well (candAtOwn of the pairwise MIS)
    const float own_target = target_pdf(ris_surface(N, load3(p.g.ray_dir, i), load3(p.g.brdf, i)), cLe, cL);
    p.ws.lcache[2 * i] = make_float4(cLe.x, cLe.y, cLe.z, own_target);
    p.ws.lcache[2 * i + 1] = make_float4(cL.x, cL.y, cL.z, 0.f);
    // a pixel outside the band of rows this launch resamples (row-band rendering, MR_CTR_BAND_LO) is on the list only
    // for the entry above -- the band's pixels read it when they pick this one as a neighbour: it queues no ray (it stays
    // in the warp's votes below) and the resolve pass skips it
    const bool band = in_band(p.ws, py);
    const float3 cur_pos = load3(p.pos_map, i);
    const size_t base = (size_t)a * MR_MAX_RAYS_PER_PIXEL;
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        if (ok[k]) ok[k] = neighbor_ok(N, nd.w, make_float3(nnd[k].x, nnd[k].y, nnd[k].z), nnd[k].w);
        if (ok[k]) ok[k] = nM[k] != 0;
        if (ok[k]) ok[k] = !(nocc[k] < 0.1f);
        ok[k] = ok[k] && band;
    }
    // A visibility only ever MULTIPLIES a target density in the resolve pass (candAtCur *= vis(own surface -> neighbour's
    // light), canonAtNb *= vis(neighbour's surface -> own light), SpatialResampling.slang:262-284), and that density is
    // max(0, lum * brdf) with brdf = 0 unless N.L > 0.  A ray whose light lies at or below the horizon of the surface
    // it starts from therefore multiplies +0: it is not cast, its slot reads "unoccluded", the product is the same +0.
    // Likewise the products only reach the result through weights that carry the reservoir weight W as a factor:
    //   candAtCur  -> sampleWeight = candAtCur * W_neighbour * m0   (and m_factor, which feeds st.M, overwritten at the end)
    //   canonAtNb  -> m1 -> canonical sample weight = p-hat * W_own * canonicalWeight
    // A reservoir whose sample was found occluded carries W = 0 exactly (InitialResampling.slang:258-270), the weight is
    // then +0 for either visibility and a zero weight is never selected (u * wSum < 0 is false): the ray towards a
    // neighbour's dead sample and, for a pixel whose own sample is dead, all rays towards its own light are not cast.
    const float cur_w = MR_LDG(p.prev.w + i);
    bool cast0[MR_MAX_NEIGHBORS], cast1[MR_MAX_NEIGHBORS];
    float3 nLk[MR_MAX_NEIGHBORS];
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        nLk[k] = f3(0.f);
        cast0[k] = cast1[k] = false;
        if (ok[k]) {
            nLk[k] = oct_decode(nld[k].y, nld[k].z);
            cast0[k] = dot(N, nLk[k]) > 0.f && nw[k] != 0.f;
            cast1[k] = dot(make_float3(nnd[k].x, nnd[k].y, nnd[k].z), cL) > 0.f && cur_w != 0.f;
        }
    }
#if defined(__CUDA_ARCH__)
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int m0[MR_MAX_NEIGHBORS], m1[MR_MAX_NEIGHBORS];
    int total = 0;
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        m0[k] = __ballot_sync(act, cast0[k]);
        m1[k] = __ballot_sync(act, cast1[k]);
        total += __popc(m0[k]) + __popc(m1[k]);
    }
    int q0 = 0;
    const int leader = __ffs(act) - 1;
    if ((int)lane == leader && total > 0) q0 = atomicAdd(p.ws.counters + MR_CTR_ANY_SIZE, total);
    q0 = __shfl_sync(act, q0, leader);
#pragma unroll
    for (uint32_t k = 0; k < MR_MAX_NEIGHBORS; ++k) {
        if (k < p.neighbor_count) {
            if (ok[k]) {
                if (cast0[k]) queue_ray_at(p.ws, q0 + __popc(m0[k] & lt_mask), base + 2 * k, cur_pos + VIS_NEAR * nLk[k], nLk[k]);
                else p.ws.hit[base + 2 * k] = MR_HIT_MISS;
                if (cast1[k]) queue_ray_at(p.ws, q0 + __popc(m0[k]) + __popc(m1[k] & lt_mask), base + 2 * k + 1, npos[k] + VIS_NEAR * cL, cL);
                else p.ws.hit[base + 2 * k + 1] = MR_HIT_MISS;
            } else {
                queue_empty(p.ws, base + 2 * k);
                queue_empty(p.ws, base + 2 * k + 1);
            }
            q0 += __popc(m0[k]) + __popc(m1[k]);
        }
    }
#else
    for (uint32_t k = 0; k < p.neighbor_count; ++k) {
        if (ok[k]) {
            if (cast0[k]) queue_ray(p.ws, base + 2 * k, cur_pos + VIS_NEAR * nLk[k], nLk[k]);
            else p.ws.hit[base + 2 * k] = MR_HIT_MISS;
            if (cast1[k]) queue_ray(p.ws, base + 2 * k + 1, npos[k] + VIS_NEAR * cL, cL);
            else p.ws.hit[base + 2 * k + 1] = MR_HIT_MISS;
        } else {
            queue_empty(p.ws, base + 2 * k);
            queue_empty(p.ws, base + 2 * k + 1);
        }
    }
#endif
}

// resolve: the pairwise-MIS streaming pass with the traced visibilities (res.slang:173-232)
MR_DEV void spatial_resolve_px(const SpatialParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    if (!in_band(p.ws, py)) return;
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    float4 nd = load_nd(p.g.normal_depth, i);
    const float3 N = make_float3(nd.x, nd.y, nd.z);
    const RisSurface cur_s = ris_surface(N, load3(p.g.ray_dir, i), load3(p.g.brdf, i));
    Ris st = ris_empty();
    const uint32_t startIndex = to_uint(rnd(sg) * (float)p.offset_count);
    const Reservoir cur = res_load(p.prev, i);
    const float4 c0 = p.ws.lcache[2 * i], c1 = p.ws.lcache[2 * i + 1]; // light_of(cur.ld), written by the gen pass
    const float3 cLe = make_float3(c0.x, c0.y, c0.z), cL = make_float3(c1.x, c1.y, c1.z);
    const float currentTargetPdf = c0.w; // target_pdf(cur_s, cLe, cL), from the gen pass
    st.canonical = 1.f;
    uint32_t validNeighbors = 1;
    unsigned char tag = 0; // of the sample selected so far
    const size_t base = (size_t)a * MR_MAX_RAYS_PER_PIXEL;
    for (uint32_t k = 0; k < p.neighbor_count; ++k) {
        if (p.ws.hit[base + 2 * k] == MR_HIT_NONE) continue;
        size_t n;
        spatial_neighbor(p, px, py, startIndex, k, n);
        float4 nnd = load_nd(p.g.normal_depth, n);
        const float3 nN = make_float3(nnd.x, nnd.y, nnd.z);
        const Reservoir nr = res_load(p.prev, n);
        const RisSurface nb_s = ris_surface(nN, load3(p.g.ray_dir, n), load3(p.g.brdf, n));
        ++validNeighbors;
        const float4 n0 = p.ws.lcache[2 * n], n1 = p.ws.lcache[2 * n + 1]; // light_of(nr.ld): n is on the pixel list
        const float3 nLe = make_float3(n0.x, n0.y, n0.z), nL = make_float3(n1.x, n1.y, n1.z);
        const bool canonical_hit = p.ws.hit[base + 2 * k] == MR_HIT_HIT;
        const bool candidate_hit = p.ws.hit[base + 2 * k + 1] == MR_HIT_HIT;
        float candidateVisibility = candidate_hit ? 0.f : 1.0f;
        float canonicalVisibility = canonical_hit ? 0.f : 1.0f;
        float candAtOwn = n0.w; // target_pdf(nb_s, nLe, nL): the neighbour's own target density, from the gen pass
        float candAtCur = target_pdf(cur_s, nLe, nL);
        float canonAtNb = target_pdf(nb_s, cLe, cL);
        candAtCur *= canonicalVisibility;
        canonAtNb *= candidateVisibility;
        float N0 = (float)((uint32_t)nr.M * p.neighbor_count);
        float N1 = (float)cur.M;
        float m0 = pairwise_mis(candAtOwn, candAtCur, N0, N1);
        float m1 = 1.f - pairwise_mis(canonAtNb, currentTargetPdf, N0, N1);
        float sampleWeight = candAtCur * nr.w * m0;
        st.M += (float)nr.M * fminf(m_factor(candAtOwn, candAtCur), m_factor(canonAtNb, currentTargetPdf));
        st.wsum += sampleWeight;
        st.canonical += m1;
        // a neighbour's sample is only ever selected with a positive weight, i.e. with candAtCur > 0: slot 2k was cast
        // from this pixel's position towards that sample and came back unoccluded
        if (rnd(sg) * st.wsum < sampleWeight) { st.ld = nr.ld; st.pdf = nr.pdf; st.weight = candAtCur; tag = 1; }
    }
    {
        float sampleWeight = currentTargetPdf * cur.w * st.canonical;
        st.M += (float)cur.M;
        st.wsum += sampleWeight;
        if (rnd(sg) * st.wsum < sampleWeight) {
            st.ld = cur.ld; st.pdf = cur.pdf; st.weight = currentTargetPdf;
            tag = p.prev_vis_tag ? p.prev_vis_tag[i] : 0;
        }
    }
    st.M = (float)cur.M;
    st.weight = st.weight > 0.f ? (st.wsum / (float)validNeighbors) / st.weight : 0.f;
    res_store(p.res, i, st);
    if (p.vis_tag) p.vis_tag[i] = tag;
}

// ------------------------------------------------------------------------------------------------------------------
struct VisParams {
    BvhView bvh;
    const float *__restrict__ res_ld;
    const float *__restrict__ pos_map;
    float *__restrict__ vis;
    int n;
    const unsigned char *vis_tag; // optional: 1 = the stored sample is known to be unoccluded from this pixel
    Workspace ws;
};
// thread t: vis[t] = 1 for every pixel (EvaluateFinalSamples.slang:103); the first n_active threads also queue a ray
MR_DEV void final_visibility_gen_px(const VisParams &p, int t)
{
    p.vis[t] = 1.0f;
    if (t >= p.ws.counters[0]) return;
    const size_t i = (size_t)p.ws.active[t];
    float3 ld = load3(p.res_ld, i);
    // a sample tagged 1 has already been through THIS ray (same origin pos + 0.01 L, same direction, same tree) in the
    // pass that selected it and was not occluded: vis stays 1 without a second cast
    if (ld.x > 0.1f && !(p.vis_tag && p.vis_tag[i] == 1)) {
        float3 L = oct_decode(ld.y, ld.z);
        queue_ray(p.ws, (size_t)t, load3(p.pos_map, i) + VIS_NEAR * L, L);
    } else {
        queue_empty(p.ws, (size_t)t);
    }
}
MR_DEV void final_visibility_resolve_px(const VisParams &p, int a)
{
    if (a >= p.ws.counters[0]) return;
    if (p.ws.hit[a] == MR_HIT_NONE) return;
    p.vis[p.ws.active[a]] = p.ws.hit[a] == MR_HIT_HIT ? 0.0f : 1.0f;
}
struct EvalParams {
    ResConst res;
    EnvView env;
    const float *__restrict__ vis;
    float *__restrict__ fs_dir;
    float *__restrict__ fs_dist;
    float *__restrict__ fs_Li;
    const float *__restrict__ grad_Li; // backward only
    float *grad_env;                   // backward only
};
MR_DEV void eval_final_fwd_px(const EvalParams &p, int idx)
{
    const ResConst &res = p.res;
    const EnvView &env = p.env;
    const float *vis = p.vis;
    float *fs_dir = p.fs_dir, *fs_dist = p.fs_dist, *fs_Li = p.fs_Li;
    const size_t i = (size_t)idx;
    float3 ld = load3(res.ld, i);
    float3 dir = f3(0.f), Li = f3(0.f);
    float dist = 0.f;
    if (ld.x > 0.1f) {
        float3 Le, L;
        light_of(env, ld.y, ld.z, Le, L);
        if (MR_LDG(vis + i) > 0.f) {
            dir = L;
            dist = 1e6f;
            Li = MR_LDG(res.w + i) * Le;
        }
    }
    store3(fs_dir, i, dir);
    fs_dist[i] = dist;
    store3(fs_Li, i, Li);
}

// Backward of Li = W * bilinear(env)(dir): grad_env[tap] += w_tap * W * grad_Li (4 taps x 3 channels).
// Many pixels select the same bright texels, so lanes of a warp that hit the same tap quad are summed with a
// shuffle tree first (__match_any_sync on the texel index) and a single lane issues the atomics.
// per-pixel part shared by both builds: which quad, which weights
MR_DEV bool eval_final_bwd_terms(const EvalParams &p, int idx, Taps &t, float vals[12])
{
    const size_t i = (size_t)idx;
    float3 ld = load3(p.res.ld, i);
    if (!(ld.x > 0.1f) || !(MR_LDG(p.vis + i) > 0.f)) return false;
    float3 L = oct_decode(ld.y, ld.z);
    float2 uv;
    if (!env_uv_of(ngp_dir(L), uv)) return false;
    t = bilinear_taps(uv, p.env.W, p.env.H);
    float3 g = MR_LDG(p.res.w + i) * load3(p.grad_Li, i);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f) return false;
    float iu = 1.0f - t.u, iv = 1.0f - t.v;
    float w00 = iu * iv, w10 = t.u * iv, w01 = iu * t.v, w11 = t.u * t.v;
    vals[0] = w00 * g.x; vals[1] = w00 * g.y; vals[2] = w00 * g.z;
    vals[3] = w10 * g.x; vals[4] = w10 * g.y; vals[5] = w10 * g.z;
    vals[6] = w01 * g.x; vals[7] = w01 * g.y; vals[8] = w01 * g.z;
    vals[9] = w11 * g.x; vals[10] = w11 * g.y; vals[11] = w11 * g.z;
    return true;
}

#if !defined(MR_HOST_CHECK)
// Two levels of aggregation before the global atomics.  All lit pixels of a frame pick the same few bright texels (the
// sun lobe of the C2 envmap collects most of them), and 12 atomics per warp on the same addresses serialise in L2
// (measured 43 us per launch).  Level 1: lanes of a warp with the same tap quad are summed with shuffles
// (__match_any_sync).  Level 2: the warp leaders add their sums into a per-block table in shared memory (open
// addressing on the quad key, shared-memory atomics); the table is flushed with one set of global atomics per
// distinct quad and block.
#define EVB_SLOTS 128
#define EVB_EMPTY 0xffffffffu
__global__ void __launch_bounds__(256) k_eval_final_bwd(EvalParams p, int n)
{
    __shared__ unsigned int s_key[EVB_SLOTS];
    __shared__ float s_val[EVB_SLOTS][12];
    for (int i = threadIdx.x; i < EVB_SLOTS; i += 256) s_key[i] = EVB_EMPTY;
    for (int i = threadIdx.x; i < EVB_SLOTS * 12; i += 256) (&s_val[0][0])[i] = 0.f;
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    Taps t;
    t.i00 = t.i10 = t.i01 = t.i11 = -1;
    t.u = t.v = 0.f;
    float vals[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) vals[q] = 0.f;
    const bool active = idx < n && eval_final_bwd_terms(p, idx, t, vals);
    float *grad_env = p.grad_env;
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int act_mask = __ballot_sync(0xffffffffu, active);
    if (act_mask != 0u) {
        // quad key: origin texel + which of the two clamped neighbours differ from it (x1 - x0, y1 - y0 are 0 or 1)
        const unsigned int key = active ? ((unsigned int)t.i00 << 2) | (t.i10 != t.i00 ? 1u : 0u) | (t.i01 != t.i00 ? 2u : 0u) : 0u;
        unsigned int peers = __match_any_sync(0xffffffffu, active ? (int)key : -1 - (int)lane);
        peers &= act_mask;
        const bool leader = active && (__ffs(peers) - 1 == (int)lane);
        // level 1: every lane pulls the contributions of the other members of its group
        float acc[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) acc[q] = vals[q];
        unsigned int rest = peers & ~(1u << lane);
        const int rounds = (int)__reduce_max_sync(0xffffffffu, (unsigned int)__popc(peers)) - 1;
        for (int r = 0; r < rounds; ++r) {
            const int src = rest ? __ffs(rest) - 1 : (int)lane;
            const bool take = rest != 0u;
            rest &= rest - 1u;
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const float x = __shfl_sync(0xffffffffu, vals[q], src);
                if (take) acc[q] += x;
            }
        }
        // level 2: per-block table
        if (leader) {
            unsigned int slot = (key * 2654435761u) >> 25; // top 7 bits
            bool placed = false;
            for (int probe = 0; probe < 8 && !placed; ++probe) {
                const unsigned int prev = atomicCAS(&s_key[slot], EVB_EMPTY, key);
                if (prev == EVB_EMPTY || prev == key) {
#pragma unroll
                    for (int q = 0; q < 12; ++q) atomicAdd(&s_val[slot][q], acc[q]);
                    placed = true;
                } else {
                    slot = (slot + 1u) & (EVB_SLOTS - 1u);
                }
            }
            if (!placed) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    atomicAdd(grad_env + 3 * (size_t)t.i00 + c, acc[c]);
                    atomicAdd(grad_env + 3 * (size_t)t.i10 + c, acc[3 + c]);
                    atomicAdd(grad_env + 3 * (size_t)t.i01 + c, acc[6 + c]);
                    atomicAdd(grad_env + 3 * (size_t)t.i11 + c, acc[9 + c]);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < EVB_SLOTS && s_key[threadIdx.x] != EVB_EMPTY) {
        const unsigned int key = s_key[threadIdx.x];
        const int W = p.env.W;
        const int i00 = (int)(key >> 2);
        const int i10 = i00 + (int)(key & 1u);
        const int i01 = i00 + ((key & 2u) ? W : 0);
        const int i11 = i01 + (int)(key & 1u);
        const float *v = s_val[threadIdx.x];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicAdd(grad_env + 3 * (size_t)i00 + c, v[c]);
            atomicAdd(grad_env + 3 * (size_t)i10 + c, v[3 + c]);
            atomicAdd(grad_env + 3 * (size_t)i01 + c, v[6 + c]);
            atomicAdd(grad_env + 3 * (size_t)i11 + c, v[9 + c]);
        }
    }
}
static int eval_final_bwd_launch(const EvalParams &p, int n, cudaStream_t st)
{
    k_eval_final_bwd<<<(n + 255) / 256, 256, 0, st>>>(p, n);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
}
#else
static int eval_final_bwd_launch(const EvalParams &p, int n, cudaStream_t)
{
    for (int idx = 0; idx < n; ++idx) {
        Taps t;
        float vals[12];
        if (!eval_final_bwd_terms(p, idx, t, vals)) continue;
        for (int c = 0; c < 3; ++c) {
            p.grad_env[3 * (size_t)t.i00 + c] += vals[c];
            p.grad_env[3 * (size_t)t.i10 + c] += vals[3 + c];
            p.grad_env[3 * (size_t)t.i01 + c] += vals[6 + c];
            p.grad_env[3 * (size_t)t.i11 + c] += vals[9 + c];
        }
    }
    return 0;
}
#endif

} // namespace mr

using namespace mr;

extern "C" {

#define MR_WS_ARGS void *workspace, size_t workspace_bytes
static int ws_open(Workspace &ws, int n, void *workspace, size_t workspace_bytes)
{
    if (!workspace) return MIRRES_ERR_NULL;
    if ((uintptr_t)workspace & 255) return MIRRES_ERR_ALIGN;
    if (workspace_bytes < workspace_carve(nullptr, n, nullptr)) return MIRRES_ERR_SCRATCH;
    workspace_carve(&ws, n, (char *)workspace);
    return 0;
}
// zero reservoirs for the background pixels and (one launch) the queue counters of the stage that follows
static void res_zero_all(const ResView &r, int n, const Workspace &ws, cudaStream_t st)
{
    ZeroRegions z = {{r.ld, r.pdf, r.M, r.w, ws.counters + 1}, {3 * (size_t)n, (size_t)n, (size_t)n, (size_t)n, 4}};
    zero_regions_async(z, st);
}

// visibility tags of the calling host thread (mirres_set_visibility_tags): consumed by the four passes below
static thread_local unsigned char *t_vis_tag = nullptr;
static thread_local const unsigned char *t_prev_vis_tag = nullptr;
int mirres_set_visibility_tags(unsigned char *res_tag, const unsigned char *prev_tag)
{
    t_vis_tag = res_tag;
    t_prev_vis_tag = prev_tag;
    return 0;
}

int mirres_initial_resampling(const void *packed_nodes, const void *packed_tris, const float *pos_map, float *res_ld,
                              float *res_pdf, int *res_M, float *res_w, const float *env_tex, int env_w, int env_h,
                              int fx, int fy, unsigned int frame_index, const float *occ, const float *normal_depth,
                              const float *brdf_map, const float *ray_dir, const float *pdf_, const float *mpdf_,
                              const float *light_data, const float *light_pdf, const float *light_cache,
                              int tile_count, int tile_size, int screen_tile, int n_light, int n_brdf, MR_WS_ARGS,
                              void *stream)
{
    if (!packed_nodes || !packed_tris || !pos_map || !res_ld || !res_pdf || !res_M || !res_w || !env_tex || !occ ||
        !normal_depth || !brdf_map || !ray_dir || !pdf_ || !mpdf_ || !light_data || !light_pdf)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1 || tile_count < 1 || tile_size < 1 || screen_tile < 1 || n_light < 1 || n_brdf < 0)
        return MIRRES_ERR_SHAPE;
    if (((uintptr_t)normal_depth & 15) || ((uintptr_t)light_cache & 15)) return MIRRES_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = fx * fy;
    InitialParams p;
    int rc = ws_open(p.ws, n, workspace, workspace_bytes);
    if (rc) return rc;
    p.bvh = bvh_view(packed_nodes, packed_tris);
    p.env = {env_tex, env_w, env_h, pdf_, nullptr, mpdf_, nullptr};
    p.g = {occ, normal_depth, brdf_map, ray_dir};
    p.pos_map = pos_map;
    p.res = {res_ld, res_pdf, res_M, res_w};
    p.light_data = light_data;
    p.light_pdf = light_pdf;
    p.light_cache = (const float4 *)light_cache;
    p.fx = fx; p.fy = fy; p.frame = frame_index;
    p.tile_count = tile_count; p.tile_size = tile_size; p.screen_tile = screen_tile; p.n_light = n_light; p.n_brdf = n_brdf;
    p.vis_tag = t_vis_tag;
    res_zero_all(p.res, n, p.ws, st); // background pixels (InitialResampling.slang:166-176) + queue_reset
    if ((rc = foreach_item<InitialParams, initial_gen_px, 128>(p, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, p.ws, true, false, device_sm_count(), st))) return rc;
    return foreach_item<InitialParams, initial_resolve_px, 256>(p, n, st);
}

int mirres_temporal_resampling(float *res_ld, float *res_pdf, int *res_M, float *res_w, const float *prev_ld,
                               const float *prev_pdf, const int *prev_M, const float *prev_w, const float *env_tex,
                               int env_w, int env_h, int fx, int fy, unsigned int frame_index, const float *occ,
                               const float *normal_depth, const float *brdf_map, const float *ray_dir,
                               const float *prev_occ, const float *prev_normal_depth, const float *prev_brdf_map,
                               const float *prev_ray_dir, const float *motion, int max_history, MR_WS_ARGS, void *stream)
{
    if (!res_ld || !res_pdf || !res_M || !res_w || !prev_ld || !prev_pdf || !prev_M || !prev_w || !env_tex || !occ ||
        !normal_depth || !brdf_map || !ray_dir || !prev_occ || !prev_normal_depth || !prev_brdf_map || !prev_ray_dir)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1 || max_history < 0) return MIRRES_ERR_SHAPE;
    if (((uintptr_t)normal_depth & 15) || ((uintptr_t)prev_normal_depth & 15)) return MIRRES_ERR_ALIGN;
    TemporalParams p;
    int rc = ws_open(p.ws, fx * fy, workspace, workspace_bytes);
    if (rc) return rc;
    p.env = {env_tex, env_w, env_h, nullptr, nullptr, nullptr, nullptr};
    p.g = {occ, normal_depth, brdf_map, ray_dir};
    p.prev_g = {prev_occ, prev_normal_depth, prev_brdf_map, prev_ray_dir};
    p.res = {res_ld, res_pdf, res_M, res_w};
    p.prev = {prev_ld, prev_pdf, prev_M, prev_w};
    p.motion = motion;
    p.fx = fx; p.fy = fy; p.frame = frame_index; p.max_history = max_history;
    p.vis_tag = t_vis_tag; p.prev_vis_tag = t_prev_vis_tag;
    return foreach_item<TemporalParams, temporal_px, 128>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_spatial_resampling(const void *packed_nodes, const void *packed_tris, const float *pos_map, float *res_ld,
                              float *res_pdf, int *res_M, float *res_w, const float *prev_ld, const float *prev_pdf,
                              const int *prev_M, const float *prev_w, const float *neighbor_offsets, const float *env_tex,
                              int env_w, int env_h, int fx, int fy, unsigned int frame_index, const float *occ,
                              const float *normal_depth, const float *brdf_map, const float *ray_dir, int offset_count,
                              int neighbor_count, float gather_radius, MR_WS_ARGS, void *stream)
{
    if (!packed_nodes || !packed_tris || !pos_map || !res_ld || !res_pdf || !res_M || !res_w || !prev_ld || !prev_pdf ||
        !prev_M || !prev_w || !neighbor_offsets || !env_tex || !occ || !normal_depth || !brdf_map || !ray_dir)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1 || offset_count < 1 || (offset_count & (offset_count - 1)) ||
        neighbor_count < 0 || 2 * neighbor_count > MR_MAX_RAYS_PER_PIXEL)
        return MIRRES_ERR_SHAPE;
    if ((uintptr_t)normal_depth & 15) return MIRRES_ERR_ALIGN;
    if (res_ld == prev_ld) return MIRRES_ERR_ALIAS;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = fx * fy;
    SpatialParams p;
    int rc = ws_open(p.ws, n, workspace, workspace_bytes);
    if (rc) return rc;
    p.bvh = bvh_view(packed_nodes, packed_tris);
    p.env = {env_tex, env_w, env_h, nullptr, nullptr, nullptr, nullptr};
    p.g = {occ, normal_depth, brdf_map, ray_dir};
    p.pos_map = pos_map;
    p.res = {res_ld, res_pdf, res_M, res_w};
    p.prev = {prev_ld, prev_pdf, prev_M, prev_w};
    p.offsets = neighbor_offsets;
    p.fx = fx; p.fy = fy; p.frame = frame_index;
    p.offset_count = offset_count; p.neighbor_count = neighbor_count; p.radius = gather_radius;
    p.vis_tag = t_vis_tag; p.prev_vis_tag = t_prev_vis_tag;
    res_zero_all(p.res, n, p.ws, st); // background pixels (SpatialResampling.slang:192-201) + queue_reset
    if ((rc = foreach_item<SpatialParams, spatial_gen_px, 128>(p, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, p.ws, true, false, device_sm_count(), st))) return rc;
    return foreach_item<SpatialParams, spatial_resolve_px, 128>(p, n, st);
}

int mirres_final_visibility(const void *packed_nodes, const void *packed_tris, const float *res_ld, int fx, int fy,
                            const float *pos_map, float *vis_map, MR_WS_ARGS, void *stream)
{
    if (!packed_nodes || !packed_tris || !res_ld || !pos_map || !vis_map) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1) return MIRRES_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = fx * fy;
    VisParams p;
    int rc = ws_open(p.ws, n, workspace, workspace_bytes);
    if (rc) return rc;
    p.bvh = bvh_view(packed_nodes, packed_tris);
    p.res_ld = res_ld; p.pos_map = pos_map; p.vis = vis_map; p.n = n;
    p.vis_tag = t_vis_tag;
    queue_reset(p.ws, st);
    if ((rc = foreach_item<VisParams, final_visibility_gen_px, 256>(p, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, p.ws, true, false, device_sm_count(), st))) return rc;
    return foreach_item<VisParams, final_visibility_resolve_px, 256>(p, n, st);
}

int mirres_eval_final_fwd(const float *res_ld, const float *res_pdf, const int *res_M, const float *res_w,
                          const float *env_tex, int env_w, int env_h, int fx, int fy, float *fs_dir, float *fs_dist,
                          float *fs_Li, const float *vis_map, void *stream)
{
    if (!res_ld || !res_w || !env_tex || !fs_dir || !fs_dist || !fs_Li || !vis_map) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1) return MIRRES_ERR_SHAPE;
    EvalParams p = {{res_ld, res_pdf, res_M, res_w}, {env_tex, env_w, env_h, nullptr, nullptr, nullptr, nullptr},
                    vis_map, fs_dir, fs_dist, fs_Li, nullptr, nullptr};
    return foreach_item<EvalParams, eval_final_fwd_px, 256>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_eval_final_bwd(const float *res_ld, const float *res_pdf, const int *res_M, const float *res_w, int env_w,
                          int env_h, int fx, int fy, const float *vis_map, const float *grad_Li, float *grad_env,
                          void *stream)
{
    if (!res_ld || !res_w || !vis_map || !grad_Li || !grad_env) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1) return MIRRES_ERR_SHAPE;
    EvalParams p = {{res_ld, res_pdf, res_M, res_w}, {nullptr, env_w, env_h, nullptr, nullptr, nullptr, nullptr},
                    vis_map, nullptr, nullptr, nullptr, grad_Li, grad_env};
    return eval_final_bwd_launch(p, fx * fy, (cudaStream_t)stream);
}

} // extern "C"
