// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).  PARITY UNPINNED.
//
// orc_kernels.cpp : the per-pixel kernels of the ReSTIR + path-tracing loop, one function per
// live Slang kernel:
//   nerf/ScreenSpaceReSTIR/utils/res.slang:53-232
//   nerf/ScreenSpaceReSTIR/InitialResampling.slang:151-295
//   nerf/ScreenSpaceReSTIR/TemporalResampling.slang:23-135
//   nerf/ScreenSpaceReSTIR/SpatialResampling.slang:32-39,178-322
//   nerf/ScreenSpaceReSTIR/EvaluateFinalSamples.slang:84-188
//   nerf/ScreenSpaceReSTIR/FinalShading.slang:14-265,641-1009
#include <string.h>
#include "orc_bvh.h"
#include "orc_light.h"
#include "orc_brdf.h"

using namespace orc;

namespace {

struct Res { float *ld; float *pdf; int *M; float *w; };   // res.slang:5-11 (SoA tensors)
struct reservoir { f3 light_data; float light_pdf; int M; float weight; }; // res.slang:12-19
struct RisState { f3 light_data; float inv_pdf; float weightSum, M, weight, canonicalWeight; }; // res.slang:20-30

inline RisState empty_ris()
{
    RisState s;
    s.light_data = mk3(0.f);
    s.inv_pdf = 0.f;
    s.weightSum = 0.f; s.M = 0.f; s.weight = 0.f; s.canonicalWeight = 0.f;
    return s;
}
inline reservoir load_res(const Res &r, size_t i)
{
    reservoir o;
    o.light_data = mk3(r.ld[3 * i], r.ld[3 * i + 1], r.ld[3 * i + 2]);
    o.light_pdf = r.pdf[i];
    o.M = r.M[i];
    o.weight = r.w[i];
    return o;
}
inline void zero_res(const Res &r, size_t i)
{
    r.ld[3 * i] = 0.f; r.ld[3 * i + 1] = 0.f; r.ld[3 * i + 2] = 0.f;
    r.pdf[i] = 0.f; r.M[i] = 0; r.w[i] = 0.f;
}
// common tail of the three resampling kernels (e.g. InitialResampling.slang:277-293)
inline void store_res(const Res &r, size_t i, const RisState &s)
{
    r.ld[3 * i] = s.light_data.x; r.ld[3 * i + 1] = s.light_data.y; r.ld[3 * i + 2] = s.light_data.z;
    r.pdf[i] = s.inv_pdf;
    r.M[i] = f2i(s.M);
    r.w[i] = s.weight;
    if (isinf(s.weight) || isnan(s.weight)) zero_res(r, i);
}
inline f3 ld3(const float *p, size_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void st3(float *p, size_t i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// res.slang:53-56
inline float mFactor(float q0, float q1) { return q0 == 0.f ? 1.f : sclamp(mr_pow8f(smin(q1 / q0, 1.f)), 0.f, 1.f); }
// res.slang:58-61
inline float pairwiseMisWeight(float q0, float q1, float N0, float N1) { return (q1 == 0.f) ? 0.f : (N0 * q0) / (q0 * N0 + q1 * N1); }
// res.slang:63-68
inline bool isValidNeighbor(f3 cn, float cd, f3 on, float od, float nt, float dt)
{
    return dot(cn, on) >= nt && fabsf(cd - od) <= dt * cd;
}
// res.slang:70-77
inline float evalTargetFunction(f3 Le, f3 L, f3 normal, f3 ray_dir, f3 brdf)
{
    float brdfWeight = eval_brdf(L, -ray_dir, normal, brdf.z, brdf.x, brdf.y);
    return smax(0.f, luminance(Le) * brdfWeight);
}
// res.slang:79-91 with kInitialBRDFSampleCount = 1 > 0
inline float evalInitialSamplePdf(float ratio, f3 L, float light_pdf, f3 V, f3 N, float a, float dW, float sW)
{
    float brdfPdf = eval_pdf_brdf(false, L, V, N, a, dW, sW);
    return slerp(light_pdf, brdfPdf, ratio);
}
// res.slang:93-114 (light-sample overload)
inline void stepLight(RisState &st, f3 light_data, float inv_pdf, float targetPdf, float sourcePdf, uint32_t &sg)
{
    float sampleWeight = targetPdf / sourcePdf;
    st.weightSum += sampleWeight;
    st.M += 1.f;
    bool sel = next1d(sg) * st.weightSum < sampleWeight;
    if (sel) { st.light_data = light_data; st.inv_pdf = inv_pdf; st.weight = targetPdf; }
}
// res.slang:116-135 (reservoir overload)
inline bool stepRes(RisState &st, const reservoir &r, float targetPdf, uint32_t &sg)
{
    float sampleWeight = targetPdf * r.weight * (float)r.M;
    st.weightSum += sampleWeight;
    st.M += (float)r.M;
    bool sel = next1d(sg) * st.weightSum < sampleWeight;
    if (sel) { st.light_data = r.light_data; st.inv_pdf = r.light_pdf; st.weight = targetPdf; }
    return sel;
}

inline bool shadow_ray(const Bvh &b, f3 pos, f3 dir, TraceCounters *tc)
{
    // every call site: origin = pos + VIS_near * dir, t in [0,1e7] (e.g. InitialResampling.slang:258-267)
    const float vis_near = 0.01f;
    f3 o = pos + vis_near * dir;
    float t_hit = 0.f;
    f3 p = mk3(0.f);
    return bvh_hit(b, o, dir, 0.f, 1e7f, t_hit, p, 0, 0, tc);
}

void merge_counters(long long *dst, const TraceCounters &a, const TraceCounters &c)
{
    if (!dst) return;
#pragma omp critical
    {
        dst[0] += a.nodes_ref; dst[1] += a.tris_ref; dst[2] += a.nodes_any; dst[3] += a.tris_any;
        dst[4] += c.nodes_ref; dst[5] += c.tris_ref;
        int ms = a.max_stack > c.max_stack ? a.max_stack : c.max_stack;
        if (ms > dst[6]) dst[6] = ms;
    }
}

} // namespace

// Sample provenance (test infrastructure for the product's visibility tags, include/mirres_b200.h
// mirres_set_visibility_tags): one byte per pixel beside a reservoir set, 1 = the stored sample has been through the
// final-visibility ray of this pixel in an earlier pass and was not occluded.  The oracle only RECORDS it -- results do
// not depend on it -- so that (a) orc_final_visibility can check the claim "tagged => unoccluded" on every tagged ray
// (g_tag_violations) and (b) the node / triangle counters can tell the rays whose answer was known ([9..11]).
static unsigned char *g_res_tag = nullptr;
static const unsigned char *g_prev_tag = nullptr;
static long long g_tag_violations = 0;

extern "C" {

void orc_set_provenance(unsigned char *res_tag, const unsigned char *prev_tag) { g_res_tag = res_tag; g_prev_tag = prev_tag; }
long long orc_provenance_violations() { return g_tag_violations; }

// counters (optional, accumulated): [0..3] shadow rays nodes_ref,tris_ref,nodes_any,tris_any;
// [4..5] closest rays nodes,tris; [6] max stack; [7] #shadow rays; [8] #closest rays;
// [9..11] (spatial pass only) nodes_any, tris_any, #rays of the shadow rays whose result cannot reach the output: the
// visibility multiplies a target density that is +0 because the light is not above the horizon of the surface the ray
// leaves, or only feeds a weight that has the W = 0 of a dead reservoir as a factor.  They are part of [0..3] / [7];
// bench.py subtracts them when it charges the product (which does not cast them) with algorithmic bytes.

// InitialResampling.slang:151-295
int orc_initial_resampling(const int *info, const float *aabb, const float *vert, const int *tri, const float *pos_map,
                           float *res_ld, float *res_pdf, int *res_M, float *res_w, const float *env_tex, int env_w,
                           int env_h, int fx, int fy, uint32_t frameIndex, const float *occ, const float *normal_depth,
                           const float *brdf_map, const float *ray_dir, const float *pdf_, const float *cdf_,
                           const float *mpdf_, const float *mcdf_, const float *light_data, const int *light_uv,
                           const float *light_pdf, int tile_count, int tile_size, int screen_tile, int n_light,
                           int n_brdf, long long *counters)
{
    (void)light_uv;
    Bvh b = {info, aabb, vert, tri};
    Env e = {env_tex, env_w, env_h, pdf_, cdf_, mpdf_, mcdf_};
    Res R = {res_ld, res_pdf, res_M, res_w};
    const float ratio = (float)n_brdf / (float)(n_light + n_brdf);
#pragma omp parallel
    {
        TraceCounters tcs = {0, 0, 0, 0, 0}, tcc = {0, 0, 0, 0, 0};
        long long nrays = 0;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < fx * fy; ++idx) {
            const uint32_t px = (uint32_t)(idx % fx), py = (uint32_t)(idx / fx);
            const size_t i = (size_t)idx;
            if (occ[i] < 0.1f) { zero_res(R, i); continue; }
            uint32_t tileSg = seed_generator(px / (uint32_t)screen_tile, py / (uint32_t)screen_tile, frameIndex);
            uint32_t tileIndex = f2u(next1d(tileSg) * (float)(uint32_t)tile_count);
            if (tileIndex > (uint32_t)tile_count - 1) tileIndex = (uint32_t)tile_count - 1;
            uint32_t lightTileOffset = tileIndex * (uint32_t)tile_size;
            uint32_t sg = seed_generator(px, py, frameIndex);
            uint32_t stride = ((uint32_t)tile_size + (uint32_t)n_light - 1) / (uint32_t)n_light;
            uint32_t offset = f2u(next1d(sg) * (float)stride);
            if (offset > stride - 1) offset = stride - 1;
            f3 N = mk3(normal_depth[4 * i], normal_depth[4 * i + 1], normal_depth[4 * i + 2]);
            f3 rd = ld3(ray_dir, i);
            f3 brdf = ld3(brdf_map, i);
            RisState st = empty_ris();
            for (uint32_t k = 0; k < (uint32_t)n_light; ++k) {
                uint32_t index = lightTileOffset + (offset + k * stride) % (uint32_t)tile_size;
                f3 lsd = ld3(light_data, index);
                float ls_pdf = light_pdf[index];
                f3 Le, L;
                get_light_info(e, mk2(lsd.y, lsd.z), Le, L);
                float targetPdf = evalTargetFunction(Le, L, N, rd, brdf);
                float sourcePdf = evalInitialSamplePdf(ratio, L, ls_pdf, -rd, N, brdf.z, brdf.x, brdf.y);
                stepLight(st, lsd, ls_pdf, targetPdf, sourcePdf, sg);
            }
            for (int k = 0; k < n_brdf; ++k) {
                f3 lsd = mk3(0.f);
                float ls_pdf = 0.f;
                f3 dir;
                f3 xi;
                xi.x = next1d(sg); xi.y = next1d(sg); xi.z = next1d(sg);
                if (sample_brdf(false, xi, dir, -rd, N, brdf.z, brdf.x, brdf.y)) {
                    ls_pdf = pdf_li(e, dir);
                    f2 o = oct_encode(dir);
                    lsd = mk3(1.0f, o.x, o.y);
                }
                if (lsd.x < 0.1f) { st.M += 1.f; continue; }
                f3 Le = env_le(ngp_dir(dir), e);
                float targetPdf = evalTargetFunction(Le, dir, N, rd, brdf);
                float sourcePdf = evalInitialSamplePdf(ratio, dir, ls_pdf, -rd, N, brdf.z, brdf.x, brdf.y);
                stepLight(st, lsd, ls_pdf, targetPdf, sourcePdf, sg);
            }
            if (st.light_data.x > 0.1f) {
                f3 L = oct_decode(mk2(st.light_data.y, st.light_data.z));
                ++nrays;
                if (shadow_ray(b, ld3(pos_map, i), L, &tcs)) st = empty_ris();
            }
            if (g_res_tag) g_res_tag[i] = st.light_data.x > 0.1f ? 1 : 0; // what is left has passed its ray
            st.weight = st.weight > 0.f ? (st.weightSum / st.M) / st.weight : 0.f;
            st.M = 1.f;
            store_res(R, i, st);
        }
        merge_counters(counters, tcs, tcc);
        if (counters) {
#pragma omp atomic
            counters[7] += nrays;
        }
    }
    return 0;
}

// TemporalResampling.slang:23-135 (kUsePairwiseMIS = 0, kUnbiased = 0; motion vectors explicit)
int orc_temporal_resampling(float *res_ld, float *res_pdf, int *res_M, float *res_w, const float *prev_ld,
                            const float *prev_pdf, const int *prev_M, const float *prev_w, const float *env_tex,
                            int env_w, int env_h, int fx, int fy, uint32_t frameIndex, const float *occ,
                            const float *normal_depth, const float *brdf_map, const float *ray_dir,
                            const float *prev_occ, const float *prev_normal_depth, const float *prev_brdf_map,
                            const float *prev_ray_dir, const float *motion, int max_history)
{
    Env e = {env_tex, env_w, env_h, 0, 0, 0, 0};
    Res R = {res_ld, res_pdf, res_M, res_w};
    Res P = {(float *)prev_ld, (float *)prev_pdf, (int *)prev_M, (float *)prev_w};
#pragma omp parallel for schedule(dynamic, 256)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const uint32_t px = (uint32_t)(idx % fx), py = (uint32_t)(idx / fx);
        const size_t i = (size_t)idx;
        if (occ[i] < 0.1f) continue;
        uint32_t sg = seed_generator(px, py, frameIndex);
        float u0 = next1d(sg), u1 = next1d(sg);
        float mvx = motion ? motion[2 * i] : 0.f, mvy = motion ? motion[2 * i + 1] : 0.f;
        int ppx = f2i((float)px + mvx * (float)(uint32_t)fx + (u0 * 1.f - 0.f));
        int ppy = f2i((float)py + mvy * (float)(uint32_t)fy + (u1 * 1.f - 0.f));
        if (ppx >= fx || ppx < 0) continue;
        if (ppy >= fy || ppy < 0) continue;
        const size_t pi = (size_t)ppy * fx + ppx;
        if (prev_occ[pi] < 0.1f) continue;
        f3 N = mk3(normal_depth[4 * i], normal_depth[4 * i + 1], normal_depth[4 * i + 2]);
        float depth = normal_depth[4 * i + 3];
        f3 rd = ld3(ray_dir, i);
        f3 brdf = ld3(brdf_map, i);
        f3 pN = mk3(prev_normal_depth[4 * pi], prev_normal_depth[4 * pi + 1], prev_normal_depth[4 * pi + 2]);
        float pdepth = prev_normal_depth[4 * pi + 3];
        f3 prd_ = ld3(prev_ray_dir, pi);
        f3 pbrdf = ld3(prev_brdf_map, pi);
        reservoir cur = load_res(R, i);
        reservoir prev = load_res(P, pi);
        {
            uint32_t cap = (uint32_t)cur.M * (uint32_t)max_history; // int * uint -> uint; min(int,uint) -> uint
            prev.M = (int)((uint32_t)prev.M < cap ? (uint32_t)prev.M : cap);
        }
        RisState st = empty_ris();
        if (!isValidNeighbor(N, depth, pN, pdepth, 0.5f, 0.1f)) continue;
        f3 Le, L;
        get_light_info(e, mk2(cur.light_data.y, cur.light_data.z), Le, L);
        float targetPdf = evalTargetFunction(Le, L, N, rd, brdf);
        stepRes(st, cur, targetPdf, sg);
        f3 pLe, pL;
        get_light_info(e, mk2(prev.light_data.y, prev.light_data.z), pLe, pL);
        float pre_targetPdf = evalTargetFunction(pLe, pL, N, rd, brdf);
        bool usedPrev = stepRes(st, prev, pre_targetPdf, sg);
        f3 sLe, sL;
        get_light_info(e, mk2(st.light_data.y, st.light_data.z), sLe, sL);
        float currentPdf = evalTargetFunction(sLe, sL, N, rd, brdf);
        float prevPdf = evalTargetFunction(sLe, sL, pN, prd_, pbrdf);
        float normalization = (usedPrev ? prevPdf : currentPdf) / ((float)cur.M * currentPdf + (float)prev.M * prevPdf);
        st.weight = st.weight > 0.f ? (st.weightSum * normalization) / st.weight : 0.f;
        store_res(R, i, st);
        if (g_res_tag && usedPrev) g_res_tag[i] = (g_prev_tag && pi == i) ? g_prev_tag[pi] : 0;
    }
    return 0;
}

// SpatialResampling.slang:178-322 (kUsePairwiseMIS = 1, kUnbiased = 1)
int orc_spatial_resampling(const int *info, const float *aabb, const float *vert, const int *tri, const float *pos_map,
                           float *res_ld, float *res_pdf, int *res_M, float *res_w, const float *prev_ld,
                           const float *prev_pdf, const int *prev_M, const float *prev_w, const float *neighborOffsets,
                           const float *env_tex, int env_w, int env_h, int fx, int fy, uint32_t frameIndex,
                           const float *occ, const float *normal_depth, const float *brdf_map, const float *ray_dir,
                           int offset_count, int neighbor_count, float gather_radius, long long *counters)
{
    Bvh b = {info, aabb, vert, tri};
    Env e = {env_tex, env_w, env_h, 0, 0, 0, 0};
    Res R = {res_ld, res_pdf, res_M, res_w};
    Res P = {(float *)prev_ld, (float *)prev_pdf, (int *)prev_M, (float *)prev_w};
    const uint32_t mask = (uint32_t)offset_count - 1;
#pragma omp parallel
    {
        TraceCounters tcs = {0, 0, 0, 0, 0}, tcc = {0, 0, 0, 0, 0}, tcd = {0, 0, 0, 0, 0};
        long long nrays = 0, ndead = 0;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < fx * fy; ++idx) {
            const uint32_t px = (uint32_t)(idx % fx), py = (uint32_t)(idx / fx);
            const size_t i = (size_t)idx;
            if (occ[i] < 0.1f) { zero_res(R, i); continue; }
            uint32_t sg = seed_generator(px, py, frameIndex);
            f3 N = mk3(normal_depth[4 * i], normal_depth[4 * i + 1], normal_depth[4 * i + 2]);
            float depth = normal_depth[4 * i + 3];
            f3 rd = ld3(ray_dir, i);
            f3 brdf = ld3(brdf_map, i);
            RisState st = empty_ris();
            const uint32_t startIndex = f2u(next1d(sg) * (float)(uint32_t)offset_count);
            reservoir cur = load_res(P, i);
            f3 cLe, cL;
            get_light_info(e, mk2(cur.light_data.y, cur.light_data.z), cLe, cL);
            float currentTargetPdf = evalTargetFunction(cLe, cL, N, rd, brdf);
            f3 curr_pos = ld3(pos_map, i);
            st.canonicalWeight = 1.f;
            uint32_t validNeighbors = 1;
            unsigned char tag = 0;
            for (uint32_t k = 0; k < (uint32_t)neighbor_count; ++k) {
                uint32_t ni = (startIndex + k) & mask; // getNextNeighborPixel :32-39
                int npx = (int)px + f2i(neighborOffsets[2 * (size_t)ni] * gather_radius);
                int npy = (int)py + f2i(neighborOffsets[2 * (size_t)ni + 1] * gather_radius);
                if (!(npx >= 0 && npx < fx && npy >= 0 && npy < fy)) continue;
                const size_t nidx = (size_t)npy * fx + npx;
                f3 nN = mk3(normal_depth[4 * nidx], normal_depth[4 * nidx + 1], normal_depth[4 * nidx + 2]);
                float nDepth = normal_depth[4 * nidx + 3];
                if (!isValidNeighbor(N, depth, nN, nDepth, 0.5f, 0.1f)) continue;
                reservoir nr = load_res(P, nidx);
                if (nr.M == 0) continue;
                f3 nrd = ld3(ray_dir, nidx);
                f3 nbrdf = ld3(brdf_map, nidx);
                if (occ[nidx] < 0.1f) continue;
                ++validNeighbors;
                f3 nLe, nL;
                get_light_info(e, mk2(nr.light_data.y, nr.light_data.z), nLe, nL);
                f3 neighbor_pos = ld3(pos_map, nidx);
                nrays += 2;
                const bool dead0 = !(dot(N, nL) > 0.f) || nr.weight == 0.f, dead1 = !(dot(nN, cL) > 0.f) || cur.weight == 0.f;
                ndead += (dead0 ? 1 : 0) + (dead1 ? 1 : 0);
                bool canonical_hit = shadow_ray(b, curr_pos, nL, dead0 ? &tcd : &tcs);
                bool candidate_hit = shadow_ray(b, neighbor_pos, cL, dead1 ? &tcd : &tcs);
                float candidateVisibility = candidate_hit ? 0.f : 1.0f;
                float canonicalVisibility = canonical_hit ? 0.f : 1.0f;
                // streamingResampleStepMisUnbiased res.slang:173-213
                float candidateTargetPdf = evalTargetFunction(nLe, nL, nN, nrd, nbrdf);
                float candidateTargetPdfAtOther = evalTargetFunction(nLe, nL, N, rd, brdf);
                float canonicalTargetPdfAtOther = evalTargetFunction(cLe, cL, nN, nrd, nbrdf);
                candidateTargetPdfAtOther *= canonicalVisibility;
                canonicalTargetPdfAtOther *= candidateVisibility;
                float N0 = (float)((uint32_t)nr.M * (uint32_t)neighbor_count);
                float N1 = (float)cur.M;
                float m0 = pairwiseMisWeight(candidateTargetPdf, candidateTargetPdfAtOther, N0, N1);
                float m1 = 1.f - pairwiseMisWeight(canonicalTargetPdfAtOther, currentTargetPdf, N0, N1);
                float sampleWeight = candidateTargetPdfAtOther * nr.weight * m0;
                st.M += (float)nr.M * smin(mFactor(candidateTargetPdf, candidateTargetPdfAtOther),
                                           mFactor(canonicalTargetPdfAtOther, currentTargetPdf));
                st.weightSum += sampleWeight;
                st.canonicalWeight += m1;
                bool sel = next1d(sg) * st.weightSum < sampleWeight;
                if (sel) { st.light_data = nr.light_data; st.inv_pdf = nr.light_pdf; st.weight = candidateTargetPdfAtOther; }
                if (sel) tag = canonical_hit ? 2 : 1; // 2 cannot happen (a hit zeroes the weight); orc_final_visibility would flag it
            }
            // streamingResampleFinalizeMis res.slang:215-232
            {
                float sampleWeight = currentTargetPdf * cur.weight * st.canonicalWeight;
                st.M += (float)cur.M;
                st.weightSum += sampleWeight;
                bool sel = next1d(sg) * st.weightSum < sampleWeight;
                if (sel) { st.light_data = cur.light_data; st.inv_pdf = cur.light_pdf; st.weight = currentTargetPdf; }
                if (sel) tag = g_prev_tag ? g_prev_tag[i] : 0;
            }
            if (g_res_tag) g_res_tag[i] = tag;
            st.M = (float)cur.M;
            st.weight = st.weight > 0.f ? (st.weightSum / (float)validNeighbors) / st.weight : 0.f;
            store_res(R, i, st);
        }
        merge_counters(counters, tcs, tcc);
        merge_counters(counters, tcd, TraceCounters{0, 0, 0, 0, 0});
        if (counters) {
#pragma omp atomic
            counters[7] += nrays;
#pragma omp atomic
            counters[9] += tcd.nodes_any;
#pragma omp atomic
            counters[10] += tcd.tris_any;
#pragma omp atomic
            counters[11] += ndead;
        }
    }
    return 0;
}

// EvaluateFinalSamples.slang:84-124
int orc_final_visibility(const int *info, const float *aabb, const float *vert, const int *tri, const float *res_ld,
                         int fx, int fy, const float *pos_map, float *vis_map, long long *counters)
{
    Bvh b = {info, aabb, vert, tri};
#pragma omp parallel
    {
        TraceCounters tcs = {0, 0, 0, 0, 0}, tcc = {0, 0, 0, 0, 0}, tcd = {0, 0, 0, 0, 0};
        long long nrays = 0, nknown = 0, bad = 0;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < fx * fy; ++idx) {
            const size_t i = (size_t)idx;
            f3 ld = ld3(res_ld, i);
            vis_map[i] = 1.0f;
            if (ld.x > 0.1f) {
                f3 L = oct_decode(mk2(ld.y, ld.z));
                ++nrays;
                const bool known = g_res_tag && g_res_tag[i] != 0;
                const bool hit = shadow_ray(b, ld3(pos_map, i), L, known ? &tcd : &tcs);
                vis_map[i] = hit ? 0.0f : 1.0f;
                if (known) { ++nknown; if (hit || g_res_tag[i] != 1) ++bad; }
            }
        }
        merge_counters(counters, tcs, tcc);
        merge_counters(counters, tcd, TraceCounters{0, 0, 0, 0, 0});
#pragma omp atomic
        g_tag_violations += bad;
        if (counters) {
#pragma omp atomic
            counters[7] += nrays;
#pragma omp atomic
            counters[9] += tcd.nodes_any;
#pragma omp atomic
            counters[10] += tcd.tris_any;
#pragma omp atomic
            counters[11] += nknown;
        }
    }
    return 0;
}

// EvaluateFinalSamples.slang:129-188 (forward).  taps (optional) records the bilinear taps used
// for Li so the float64 backward oracle (oracle/backward.py) can rebuild the scatter.
int orc_eval_final_fwd(const float *res_ld, const float *res_pdf, const int *res_M, const float *res_w,
                       const float *env_tex, int env_w, int env_h, int fx, int fy, float *fs_dir, float *fs_dist,
                       float *fs_Li, const float *vis_map)
{
    (void)res_pdf; (void)res_M;
    Env e = {env_tex, env_w, env_h, 0, 0, 0, 0};
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const size_t i = (size_t)idx;
        f3 ld = ld3(res_ld, i);
        st3(fs_dir, i, mk3(0.f));
        fs_dist[i] = 0.f;
        f3 Li = mk3(0.f);
        if (ld.x > 0.1f) {
            f3 Le, L;
            get_light_info(e, mk2(ld.y, ld.z), Le, L);
            float visibility = vis_map[i];
            if (visibility > 0.f) {
                st3(fs_dir, i, L);
                fs_dist[i] = 1e6f;
                Li = res_w[i] * Le;
            }
        }
        st3(fs_Li, i, Li);
    }
    return 0;
}

// FinalShading.slang:14-109 (forward)
int orc_final_shading_fwd(const float *fs_dir, const float *fs_dist, const float *fs_Li, const float *env_tex, int env_w,
                          int env_h, int fx, int fy, const float *occ, const float *normal, const float *ray_dir,
                          const float *diffuse_map, const float *rough_metal, float *color, float *diff_light,
                          float *spec_light)
{
    Env e = {env_tex, env_w, env_h, 0, 0, 0, 0};
    const float F0 = 0.04f;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const size_t i = (size_t)idx;
        f3 N = ld3(normal, i), rd = ld3(ray_dir, i), diffuse = ld3(diffuse_map, i);
        float linearRoughness = rough_metal[2 * i], metallic = rough_metal[2 * i + 1];
        f3 specular = mk3(F0) * (1.0f - metallic) + diffuse * metallic;
        f3 color_val = mk3(0.f), light_diffuse = mk3(0.f), light_spec = mk3(0.f);
        if (occ[i] > 0.1f) {
            f3 dir = ld3(fs_dir, i);
            float distance = fs_dist[i];
            f3 Li = ld3(fs_Li, i);
            f3 diffuse_val = mk3(0.f), specular_val = mk3(0.f);
            if (distance > 0.f) {
                Frame fr = create_frame(N);
                f3 wiLocal = frame_to_local(fr, -rd);
                f3 woLocal = frame_to_local(fr, dir);
                float ROUGHNESS_THRESHOLD = 0.01f;
                float kMinGGXAlpha = ROUGHNESS_THRESHOLD * ROUGHNESS_THRESHOLD;
                float alpha = linearRoughness * linearRoughness;
                if (alpha < kMinGGXAlpha) alpha = 0.f;
                float pD, pS;
                lobe_probs(diffuse, metallic, specular, rd, N, pD, pS);
                if (pD > 0.f) diffuse_val = mk3(diffuse_light(wiLocal, woLocal)) * Li;
                if (pS > 0.f) specular_val = specular_eval(wiLocal, woLocal, specular, alpha, true) * Li;
            }
            color_val += diffuse * (1.0f - metallic) * diffuse_val + specular_val;
            light_diffuse += diffuse_val;
            light_spec += specular_val;
        } else {
            color_val = env_le(ngp_dir(rd), e);
        }
        st3(color, i, color_val);
        st3(diff_light, i, light_diffuse);
        st3(spec_light, i, light_spec);
    }
    return 0;
}

// Shared tail of FinalShading.slang:190-262 and :907-977: sample a continuation direction, trace it.
static inline void continue_path(const Bvh &b, size_t i, f3 surf_pos, const Frame &fr, f3 wiLocal, float pD, float pS,
                                 float alpha, f3 specular, f3 diffuse_col, uint32_t &sg, uint32_t bounce_count,
                                 int max_bounce, f3 &thr, float *prd, float *new_pos_map, float *new_ray_d,
                                 float *new_occ_map, float *new_normal, TraceCounters *tcc, long long &nrays)
{
    f3 out_dir, out_weight = mk3(1.0f);
    float out_pdf;
    uint32_t sampledSpecular = 0;
    bool valid = falcor_sample(pD, pS, wiLocal, out_dir, out_pdf, sampledSpecular, out_weight, sg, alpha, specular,
                               diffuse_col, true, true);
    if (!valid) return;
    if (is_black(out_weight) || out_pdf == 0.f) {
        prd[5 * i + 4] = 1.f;
    } else if (bounce_count + 1 <= (uint32_t)max_bounce) {
        out_dir = normalize(frame_to_global(fr, out_dir));
        const float vis_near = 0.01f;
        f3 bounce_pos = surf_pos + vis_near * out_dir;
        float t_hit = 0.f;
        f3 t_pos = mk3(0.f), hit_normal = mk3(1.f);
        ++nrays;
        bool hit = bvh_hit(b, bounce_pos, out_dir, 0.f, 1e7f, t_hit, t_pos, &hit_normal, 0, tcc);
        float specularBounce = (float)sampledSpecular;
        thr *= out_weight;
        prd[5 * i + 0] = thr.x; prd[5 * i + 1] = thr.y; prd[5 * i + 2] = thr.z;
        prd[5 * i + 3] = specularBounce;
        st3(new_ray_d, i, out_dir);
        if (hit) {
            prd[5 * i + 4] = 0.f;
            st3(new_pos_map, i, t_pos);
            st3(new_normal, i, hit_normal);
            new_occ_map[i] = 1.f;
        } else if (specularBounce > 0.f) {
            prd[5 * i + 4] = 0.f;
        }
    }
}

// FinalShading.slang:113-265
int orc_bounce_first(const int *info, const float *aabb, const float *vert, const int *tri, uint32_t frameIndex,
                     uint32_t bounce_count, int max_bounce, int fx, int fy, const float *occ, const float *pos_map,
                     const float *normal, const float *ray_dir, float *prd, const float *diffuse_map,
                     const float *rough_metal, float *new_pos_map, float *new_ray_d, float *new_occ_map,
                     float *new_normal, long long *counters)
{
    Bvh b = {info, aabb, vert, tri};
    const float F0 = 0.04f;
#pragma omp parallel
    {
        TraceCounters tcs = {0, 0, 0, 0, 0}, tcc = {0, 0, 0, 0, 0};
        long long nrays = 0;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < fx * fy; ++idx) {
            const uint32_t px = (uint32_t)(idx % fx), py = (uint32_t)(idx / fx);
            const size_t i = (size_t)idx;
            f3 thr = mk3(prd[5 * i], prd[5 * i + 1], prd[5 * i + 2]);
            float is_stop = prd[5 * i + 4];
            new_occ_map[i] = 0.f;
            prd[5 * i + 4] = 1.f;
            if (bounce_count == 0) {
                thr = mk3(1.0f);
                is_stop = 0.f;
                prd[5 * i] = 1.f; prd[5 * i + 1] = 1.f; prd[5 * i + 2] = 1.f;
                prd[5 * i + 3] = 0.f;
            }
            if (is_stop > 0.f) continue;
            f3 N = ld3(normal, i), rd = ld3(ray_dir, i), P = ld3(pos_map, i), diffuse = ld3(diffuse_map, i);
            float linearRoughness = rough_metal[2 * i], metallic = rough_metal[2 * i + 1];
            f3 specular = mk3(F0) * (1.0f - metallic) + diffuse * metallic;
            uint32_t sg = seed_generator(px, py, frameIndex);
            if (occ[i] > 0.1f) {
                float ROUGHNESS_THRESHOLD = 0.01f;
                float kMinGGXAlpha = ROUGHNESS_THRESHOLD * ROUGHNESS_THRESHOLD;
                float alpha = linearRoughness * linearRoughness;
                if (alpha < kMinGGXAlpha) alpha = 0.f;
                float pD, pS;
                lobe_probs(diffuse, metallic, specular, rd, N, pD, pS);
                Frame fr = create_frame(N);
                f3 wiLocal = frame_to_local(fr, -rd);
                f3 diffuse_col = diffuse * (1.0f - metallic);
                continue_path(b, i, P, fr, wiLocal, pD, pS, alpha, specular, diffuse_col, sg, bounce_count, max_bounce,
                              thr, prd, new_pos_map, new_ray_d, new_occ_map, new_normal, &tcc, nrays);
            }
        }
        merge_counters(counters, tcs, tcc);
        if (counters) {
#pragma omp atomic
            counters[8] += nrays;
        }
    }
    return 0;
}

// FinalShading.slang:641-1009
int orc_bounce_shade(const int *info, const float *aabb, const float *vert, const int *tri, uint32_t frameIndex,
                     uint32_t bounce_count, int max_bounce, int fx, int fy, const float *env_tex, int env_w, int env_h,
                     const float *pdf_, const float *cdf_, const float *mpdf_, const float *mcdf_, const float *occ,
                     const float *pos_map, const float *normal, const float *ray_dir, float *prd,
                     const float *diffuse_map, const float *rough_metal, float *color, float *diff_color,
                     float *spec_color, float *new_pos_map, float *new_ray_d, float *new_occ_map, float *new_normal,
                     long long *counters)
{
    Bvh b = {info, aabb, vert, tri};
    Env e = {env_tex, env_w, env_h, pdf_, cdf_, mpdf_, mcdf_};
    const float F0 = 0.04f;
#pragma omp parallel
    {
        TraceCounters tcs = {0, 0, 0, 0, 0}, tcc = {0, 0, 0, 0, 0};
        long long nshadow = 0, nclosest = 0;
#pragma omp for schedule(dynamic, 256)
        for (int idx = 0; idx < fx * fy; ++idx) {
            const uint32_t px = (uint32_t)(idx % fx), py = (uint32_t)(idx / fx);
            const size_t i = (size_t)idx;
            f3 thr = mk3(prd[5 * i], prd[5 * i + 1], prd[5 * i + 2]);
            float specularBounce = prd[5 * i + 3];
            float is_stop = prd[5 * i + 4];
            new_occ_map[i] = 0.f;
            prd[5 * i + 4] = 1.f;
            if (bounce_count == 0) {
                thr = mk3(1.0f);
                specularBounce = 0.f;
                is_stop = 0.f;
                prd[5 * i] = 1.f; prd[5 * i + 1] = 1.f; prd[5 * i + 2] = 1.f;
                prd[5 * i + 3] = 0.f;
            }
            if (is_stop > 0.f) {
                st3(color, i, mk3(0.f)); st3(diff_color, i, mk3(0.f)); st3(spec_color, i, mk3(0.f));
                continue;
            }
            f3 N = ld3(normal, i), rd = ld3(ray_dir, i), P = ld3(pos_map, i), diffuse = ld3(diffuse_map, i);
            float linearRoughness = rough_metal[2 * i], metallic = rough_metal[2 * i + 1];
            f3 specular = mk3(F0) * (1.0f - metallic) + diffuse * metallic;
            f3 color_val = mk3(0.f), diff_color_val = mk3(0.f), spec_color_val = mk3(0.f);
            uint32_t sg = seed_generator(px, py, frameIndex);
            if (occ[i] > 0.1f) {
                float ROUGHNESS_THRESHOLD = 0.01f;
                float kMinGGXAlpha = ROUGHNESS_THRESHOLD * ROUGHNESS_THRESHOLD;
                float alpha = linearRoughness * linearRoughness;
                if (alpha < kMinGGXAlpha) alpha = 0.f;
                float pD, pS;
                lobe_probs(diffuse, metallic, specular, rd, N, pD, pS);
                // --- light sample (NEE) :737-815
                float lightPdf = 0.0f, scatteringPdf = 0.0f;
                bool samp_valid = false;
                f3 samp_dir_s = mk3(0.f), samp_weight_s = mk3(0.f);
                float samp_pdf_s = 0.f;
                {
                    f2 rnd;
                    rnd.x = next1d(sg);
                    rnd.y = next1d(sg);
                    f3 sdir;
                    float spdf;
                    f2 luv;
                    bool res = sample_li(e, rnd, sdir, spdf, luv);
                    if (res) {
                        samp_valid = true;
                        samp_dir_s = sdir;
                        samp_pdf_s = spdf;
                        samp_weight_s = env_le(ngp_dir(sdir), e) / spdf;
                    }
                }
                Frame fr = create_frame(N);
                f3 wiLocal = frame_to_local(fr, -rd);
                lightPdf = samp_pdf_s;
                f3 Li = samp_weight_s;
                if (samp_valid && lightPdf > 0 && !is_black(Li)) {
                    f3 diff_f = mk3(0.f), spec_f = mk3(0.f), total_f = mk3(0.f);
                    f3 woLocal = frame_to_local(fr, samp_dir_s);
                    if (!is_black(N)) {
                        if (pD > 0.f) diff_f = mk3(diffuse_light(wiLocal, woLocal));
                        if (pS > 0.f) spec_f = specular_eval(wiLocal, woLocal, specular, alpha, true);
                        f3 diffuse_col = diffuse * (1.0f - metallic);
                        total_f = diffuse_col * diff_f + spec_f;
                        diff_f = diffuse_col * diff_f;
                        scatteringPdf = falcor_eval_pdf(pD, pS, wiLocal, woLocal, alpha, true);
                    }
                    if (!is_black(total_f)) {
                        f3 light_dir = normalize(samp_dir_s);
                        ++nshadow;
                        bool hit = shadow_ray(b, P, light_dir, &tcs);
                        f3 tr = hit ? mk3(0.f) : mk3(1.f);
                        Li *= tr;
                        if (!is_black(Li)) {
                            float mis_weight = power_heuristic(lightPdf, scatteringPdf);
                            color_val += thr * total_f * Li * mis_weight;
                            diff_color_val += thr * diff_f * Li * mis_weight;
                            spec_color_val += thr * spec_f * Li * mis_weight;
                        }
                    }
                }
                // --- BSDF sample with MIS :817-905
                if (!is_black(N)) {
                    f3 bsdf_weight = mk3(1.0f), bsdf_diff_weight = mk3(1.0f), bsdf_spec_weight = mk3(1.0f);
                    f3 diffuse_col = diffuse * (1.0f - metallic);
                    f3 m_wi, dummy_w;
                    float m_pdf;
                    uint32_t sampledSpecular = 0;
                    bool valid = falcor_sample(pD, pS, wiLocal, m_wi, m_pdf, sampledSpecular, dummy_w, sg, alpha,
                                               specular, diffuse_col, true, false);
                    if (valid) {
                        if (pD > 0.f) bsdf_diff_weight = mk3(diffuse_light(wiLocal, m_wi));
                        if (pS > 0.f) bsdf_spec_weight = specular_eval(wiLocal, m_wi, specular, alpha, true);
                        bsdf_weight = diffuse_col * bsdf_diff_weight + bsdf_spec_weight;
                        m_wi = frame_to_global(fr, m_wi);
                        scatteringPdf = m_pdf;
                        f3 f = bsdf_weight / m_pdf;
                        f3 diff_f = diffuse_col * bsdf_diff_weight / m_pdf;
                        f3 spec_f = bsdf_spec_weight / m_pdf;
                        f *= m_pdf;
                        diff_f *= m_pdf;
                        spec_f *= m_pdf;
                        f3 m_safe_wi = normalize(m_wi);
                        if (!is_black(f) && scatteringPdf > 0) {
                            float weight = 1.0f;
                            bool islightpdfZero = false;
                            if (sampledSpecular == 0) {
                                float pl = pdf_li(e, m_safe_wi);
                                lightPdf = pl;
                                if (lightPdf == 0.0f) islightpdfZero = true;
                                weight = power_heuristic(scatteringPdf, lightPdf);
                            }
                            ++nshadow;
                            bool found = shadow_ray(b, P, m_safe_wi, &tcs);
                            f3 Tr = mk3(1.f);
                            Li = mk3(0.f);
                            if (!found) Li = env_le(ngp_dir(m_safe_wi), e);
                            if (!is_black(Li) && !islightpdfZero) {
                                color_val += thr * f * Li * Tr * weight / scatteringPdf;
                                diff_color_val += thr * diff_f * Li * Tr * weight / scatteringPdf;
                                spec_color_val += thr * spec_f * Li * Tr * weight / scatteringPdf;
                            }
                        }
                    }
                }
                // --- continuation :907-977
                f3 diffuse_col = diffuse * (1.0f - metallic);
                continue_path(b, i, P, fr, wiLocal, pD, pS, alpha, specular, diffuse_col, sg, bounce_count, max_bounce,
                              thr, prd, new_pos_map, new_ray_d, new_occ_map, new_normal, &tcc, nclosest);
            } else {
                if (bounce_count == 0) {
                    color_val += thr * env_le(ngp_dir(rd), e);
                } else if (specularBounce > 0.f) {
                    color_val += thr * env_le(ngp_dir(rd), e);
                    spec_color_val += thr * env_le(ngp_dir(rd), e);
                }
                prd[5 * i + 4] = 1.f;
            }
            st3(color, i, color_val);
            st3(diff_color, i, diff_color_val);
            st3(spec_color, i, spec_color_val);
        }
        merge_counters(counters, tcs, tcc);
        if (counters) {
#pragma omp atomic
            counters[7] += nshadow;
#pragma omp atomic
            counters[8] += nclosest;
        }
    }
    return 0;
}

} // extern "C"

// ---- EAW denoiser (nerf/ScreenSpaceReSTIR/EAWDenoise.slang:50-174 == :178-302) and normal AO (:591-647) ----------
extern "C" int orc_eaw_fwd(float c_phi, float n_phi, float p_phi, int fx, int fy, int stepWidth, const float *occ,
                           const float *color, const float *normal_map, const float *pos_map, float *out_color)
{
    static const float kern[25] = {1.0f / 256.0f, 1.0f / 64.0f, 3.0f / 128.0f, 1.0f / 64.0f, 1.0f / 256.0f,
                                   1.0f / 64.0f,  1.0f / 16.0f, 3.0f / 32.0f,  1.0f / 16.0f, 1.0f / 64.0f,
                                   3.0f / 128.0f, 3.0f / 32.0f, 9.0f / 64.0f,  3.0f / 32.0f, 3.0f / 128.0f,
                                   1.0f / 64.0f,  1.0f / 16.0f, 3.0f / 32.0f,  1.0f / 16.0f, 1.0f / 64.0f,
                                   1.0f / 256.0f, 1.0f / 64.0f, 3.0f / 128.0f, 1.0f / 64.0f, 1.0f / 256.0f};
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const size_t i = (size_t)idx;
        const int px = idx % fx, py = idx / fx;
        if (occ[i] < 0.1f) { st3(out_color, i, ld3(color, i)); continue; }
        f3 nval = ld3(normal_map, i), pval = ld3(pos_map, i), cval = ld3(color, i);
        f3 sum = mk3(0.f);
        float cum_w = 0.0f;
        for (int k = 0; k < 25; k++) {
            int ox = k % 5 - 2, oy = k / 5 - 2;
            int ux = px + ox * stepWidth, uy = py + oy * stepWidth;
            if (!(ux >= 0 && ux < fx && uy >= 0 && uy < fy)) continue;
            size_t u = (size_t)uy * fx + ux;
            f3 ctmp = ld3(color, u);
            f3 t = cval - ctmp;
            float dist2 = dot(t, t);
            float c_w = smin(mr_expf(-(dist2) / c_phi), 1.0f);
            f3 ntmp = ld3(normal_map, u);
            t = nval - ntmp;
            dist2 = smax(dot(t, t), 0.0f);
            float n_w = smin(mr_expf(-(dist2) / n_phi), 1.0f);
            f3 ptmp = ld3(pos_map, u);
            t = pval - ptmp;
            dist2 = smax(dot(t, t), 0.0f);
            float p_w = smin(mr_expf(-(dist2) / p_phi), 1.0f);
            float weight = c_w * n_w * p_w;
            sum += ctmp * weight * kern[k];
            cum_w += weight * kern[k];
        }
        st3(out_color, i, sum / cum_w);
    }
    return 0;
}

extern "C" int orc_normal_ao(int fx, int fy, const float *occ, const float *normal_map, float *out_ao)
{
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const size_t i = (size_t)idx;
        const int px = idx % fx, py = idx / fx;
        if (occ[i] < 0.1f) { st3(out_ao, i, mk3(0.f)); continue; }
        f3 nval = ld3(normal_map, i);
        int count = 0;
        float sum = 0.f;
        const int width = 4;
        for (int a = -width; a < width; a++)
            for (int b = -width; b < width; b++) {
                int ux = px + a, uy = py + b;
                if (!(ux >= 0 && ux < fx && uy >= 0 && uy < fy)) continue;
                size_t u = (size_t)uy * fx + ux;
                if (occ[u] < 0.1f) continue;
                float dist2 = smax(dot(ld3(normal_map, u), nval), 0.0f);
                dist2 = smin(1.0f, dist2);
                sum += dist2;
                count++;
            }
        float normal_weight = 1 - sum / (float)count;
        float sum_val = sclamp(normal_weight * 50, 0.f, 1.f);
        st3(out_ao, i, mk3(sum_val));
    }
    return 0;
}

// Bilinear tap addresses / weights of the Li lookup of EvaluateFinalSamples.slang:151-158 (get_light_info_di ->
// env_le_di -> eval_bi_di, helper.slang:72-99) for the backward oracle (oracle/backward.py).
extern "C" int orc_eval_final_taps(const float *res_ld, int env_w, int env_h, int n, int *taps /*[n,4]*/, float *uv /*[n,2]*/,
                                   int *valid /*[n]*/)
{
    for (int i = 0; i < n; ++i) {
        f3 L = oct_decode(mk2(res_ld[3 * (size_t)i + 1], res_ld[3 * (size_t)i + 2]));
        f2 q;
        bool ok = env_dir_to_uv(ngp_dir(L), q);
        valid[i] = ok ? 1 : 0;
        BiTaps t = {0, 0, 0, 0, 0.f, 0.f};
        if (ok) t = eval_bi_taps(q, env_w, env_h);
        taps[4 * (size_t)i + 0] = t.y0 * env_w + t.x0;
        taps[4 * (size_t)i + 1] = t.y0 * env_w + t.x1;
        taps[4 * (size_t)i + 2] = t.y1 * env_w + t.x0;
        taps[4 * (size_t)i + 3] = t.y1 * env_w + t.x1;
        uv[2 * (size_t)i] = t.u;
        uv[2 * (size_t)i + 1] = t.v;
    }
    return 0;
}


// ---- cross-bilateral denoiser (--use_bi_de), nerf/renderutils/c_src/denoising.cu:14-130 --------------------------------
// col [N,3], nrm [N,3] (already normalised by ops.py:197), zdz [N,2] (depth, depth gradient); out [N,4] = (sum of
// weighted colours, max(sum of weights, 1e-4)).  expf / powf(.,128) of the reference -> mr_expf / mr_pow128f.
static inline float bilateral_weight(int ox, int oy, f3 t_nrm, f3 c_nrm, float t_z, float c_z, float dz, float variance)
{
    const float FLT_EPS = 0.0001f;
    float dist_sqr = (float)(ox * ox + oy * oy);
    float dist = sqrtf(dist_sqr);
    float w_xy = mr_expf(-dist_sqr / (2.0f * variance));
    float w_normal = mr_pow128f(smin(smax(dot(t_nrm, c_nrm), FLT_EPS), 1.0f));
    float w_depth = mr_expf(-(fabsf(t_z - c_z) / smax(dz * dist, FLT_EPS)));
    return w_xy * w_normal * w_depth;
}
extern "C" int orc_bilateral_fwd(int fx, int fy, float sigma, const float *col, const float *nrm, const float *zdz, float *out)
{
    const float variance = sigma * sigma;
    const int rad = 2 * (int)ceilf(sigma * 2.5f) + 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const int px = idx % fx, py = idx / fx;
        f3 c_nrm = ld3(nrm, (size_t)idx);
        float c_z = zdz[2 * (size_t)idx], c_dz = zdz[2 * (size_t)idx + 1];
        float accum_w = 0.0f;
        f3 accum = mk3(0.f);
        for (int oy = -rad; oy <= rad; ++oy)
            for (int ox = -rad; ox <= rad; ++ox) {
                int y = py + oy, x = px + ox;
                if (y < 0 || x < 0 || y >= fy || x >= fx) continue;
                size_t t = (size_t)y * fx + x;
                float w = bilateral_weight(ox, oy, ld3(nrm, t), c_nrm, zdz[2 * t], c_z, c_dz, variance);
                accum = accum + ld3(col, t) * w;
                accum_w += w;
            }
        out[4 * (size_t)idx] = accum.x; out[4 * (size_t)idx + 1] = accum.y; out[4 * (size_t)idx + 2] = accum.z;
        out[4 * (size_t)idx + 3] = smax(accum_w, 0.0001f);
    }
    return 0;
}
// reverse mode w.r.t. col: the transposed gather of denoising.cu:72-130 (depth term with the TAP's dz)
extern "C" int orc_bilateral_bwd(int fx, int fy, float sigma, const float *nrm, const float *zdz, const float *out_grad /*[N,4]*/,
                                 float *col_grad /*[N,3]*/)
{
    const float variance = sigma * sigma;
    const int rad = 2 * (int)ceilf(sigma * 2.5f) + 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int idx = 0; idx < fx * fy; ++idx) {
        const int px = idx % fx, py = idx / fx;
        f3 c_nrm = ld3(nrm, (size_t)idx);
        float c_z = zdz[2 * (size_t)idx];
        f3 accum = mk3(0.f);
        for (int oy = -rad; oy <= rad; ++oy)
            for (int ox = -rad; ox <= rad; ++ox) {
                int y = py + oy, x = px + ox;
                if (y < 0 || x < 0 || y >= fy || x >= fx) continue;
                size_t t = (size_t)y * fx + x;
                float w = bilateral_weight(ox, oy, ld3(nrm, t), c_nrm, zdz[2 * t], c_z, zdz[2 * t + 1], variance);
                accum += mk3(out_grad[4 * t], out_grad[4 * t + 1], out_grad[4 * t + 2]) * w;
            }
        st3(col_grad, (size_t)idx, accum);
    }
    return 0;
}
