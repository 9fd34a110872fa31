// mirres-b200: final shading (forward + hand-derived backward) and the multi-bounce integrator for sm_100a.
//
// Replaces:
//   nerf/ScreenSpaceReSTIR/FinalShading.slang:14-109    process_FinalShading      -> final_shading_px
//   Slang autodiff of the same kernel (`.bwd`, called from nerf/ScreenSpaceReSTIR/Resampling.py:179-214)
//                                                                                 -> final_shading_bwd_px
//   nerf/ScreenSpaceReSTIR/FinalShading.slang:113-265   process_new_dir_for_pt    -> bounce_first_px
//   nerf/ScreenSpaceReSTIR/FinalShading.slang:641-1009  process_path_tracing_divided_no_grad -> bounce_shade_px
// `max_bounce` generalises the reference's compile-time MAX_Bounce = 2 (FinalShading.slang:7).
#include "mr_wave.cuh"
#include "mr_light.cuh"
#include "mr_brdf.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

#define VIS_NEAR 0.01f

// ---------------------------------------------------------------------------------------------------------------
struct ShadeParams {
    const float *__restrict__ fs_dir;  // [N,3]
    const float *__restrict__ fs_dist; // [N]
    const float *__restrict__ fs_Li;   // [N,3]
    EnvView env;
    const float *__restrict__ occ;     // [N]
    const float *__restrict__ normal;  // [N,3]
    const float *__restrict__ ray_dir; // [N,3]
    const float *__restrict__ kd;      // [N,3]
    const float *__restrict__ rm;      // [N,2] roughness, metallic
    float *__restrict__ color;
    float *__restrict__ diff_light;
    float *__restrict__ spec_light;
    // backward only
    const float *__restrict__ g_color;
    const float *__restrict__ g_diff;
    const float *__restrict__ g_spec;
    float *__restrict__ g_normal; // [N,3]
    float *__restrict__ g_kd;     // [N,3]
    float *__restrict__ g_rm;     // [N,2]
    float *__restrict__ g_Li;     // [N,3]
};

MR_DEV void final_shading_px(const ShadeParams &p, int idx)
{
    const size_t i = (size_t)idx;
    const float3 rd = load3(p.ray_dir, i);
    float3 color_val = f3(0.f), light_diffuse = f3(0.f), light_spec = f3(0.f);
    if (MR_LDG(p.occ + i) > 0.1f) {
        const float3 kd = load3(p.kd, i);
        const float metallic = MR_LDG(p.rm + 2 * i + 1);
        float3 diffuse_val = f3(0.f), specular_val = f3(0.f);
        if (MR_LDG(p.fs_dist + i) > 0.f) {
            const Surface s = surface_of(load3(p.normal, i), rd, kd, MR_LDG(p.rm + 2 * i), metallic);
            const float3 wi = to_frame(s.frame, load3(p.fs_dir, i));
            const float3 Li = load3(p.fs_Li, i);
            if (s.pD > 0.f) diffuse_val = f3(lambert_light(s.wo, wi)) * Li;
            if (s.pS > 0.f) specular_val = specular_f(s.wo, wi, s.spec, s.alpha) * Li;
        }
        color_val += kd * (1.0f - metallic) * diffuse_val + specular_val;
        light_diffuse += diffuse_val;
        light_spec += specular_val;
    } else {
        color_val = env_radiance(p.env, ngp_dir(rd));
    }
    store3(p.color, i, color_val);
    store3(p.diff_light, i, light_diffuse);
    store3(p.spec_light, i, light_spec);
}

// Reverse mode of final_shading_px with respect to normal, kd, (roughness, metallic) and Li.  Branch predicates
// (occupancy, distance, lobe probabilities, the 1e-6 / alpha == 0 gates, clamps) are frozen, as Slang's autodiff
// does.  The lobe probabilities only gate branches, so no gradient flows through them.
struct ShadeGrads {
    float3 gN, gKd, gLi;
    float gR, gM;
};
// gradients of one lit pixel (occ > 0.1 and sample distance > 0); gC / gD / gS: upstream on color / diff_light / spec_light
MR_DEV ShadeGrads final_shading_grads(float3 N, float3 rd, float3 kd, float rough, float metallic, float3 L, float3 Li,
                                      float3 gC, float3 gD, float3 gS)
{
    float3 gN = f3(0.f), gKd = f3(0.f), gLi = f3(0.f);
    float gR = 0.f, gM = 0.f;
    const float INV_PI = 0.31830988f;
    const float PI = 3.141592653589793f;
    const float F0 = 0.04f;
    const Surface s = surface_of(N, rd, kd, rough, metallic);
    const float3 V = -rd;
    const float3 wo = s.wo;
    const float3 wi = to_frame(s.frame, L);
    // forward values
    const bool gate = !(fminf(wo.z, wi.z) < 1e-6f);
    float Dl = 0.f;
    if (s.pD > 0.f && gate) Dl = fmaxf(INV_PI * wi.z, 0.0f);
    const float3 diffuse_val = f3(Dl) * Li;
    // upstream on diffuse_val / specular_val
    const float3 gDv = gC * (kd * (1.0f - metallic)) + gD;
    const float3 gSv = gC + gS;
    // color = kd (1-m) diffuse_val + specular_val
    gKd += gC * diffuse_val * (1.0f - metallic);
    gM += -(gC.x * kd.x * diffuse_val.x + gC.y * kd.y * diffuse_val.y + gC.z * kd.z * diffuse_val.z);
    gLi += gDv * Dl;
    float3 gwo = f3(0.f), gwi = f3(0.f);
    if (s.pD > 0.f && gate && INV_PI * wi.z > 0.0f) gwi.z += INV_PI * (gDv.x * Li.x + gDv.y * Li.y + gDv.z * Li.z);
    if (s.pS > 0.f && gate && s.alpha != 0.f) {
        const float alpha = s.alpha;
        const float3 sum = wo + wi;
        const float len = sqrtf(dot(sum, sum));
        const float3 h = sum / len;
        const float c = dot(wo, h);
        const float a2 = alpha * alpha;
        const float dd = ((h.z * a2 - h.z) * h.z + 1);
        const float D = a2 / (dd * dd * PI);
        const float lI = ggx_lambda(a2, wo.z), lO = ggx_lambda(a2, wi.z);
        const float G = 1 / (1 + lI + lO);
        const float om = fmaxf(1 - c, 0);
        const float p5 = mr_pow5f(om);
        const float3 F = make_float3(s.spec.x + (1 - s.spec.x) * p5, s.spec.y + (1 - s.spec.y) * p5, s.spec.z + (1 - s.spec.z) * p5);
        const float sc = D * G * 0.25f / wo.z; // Fs = F * sc
        gLi += gSv * (F * sc);
        const float3 gFs = gSv * Li;
        const float3 gF = gFs * sc;
        const float gsc = gFs.x * F.x + gFs.y * F.y + gFs.z * F.z;
        // F = spec + (1 - spec) p5
        const float3 gspec = gF * (1 - p5);
        const float gp5 = gF.x * (1 - s.spec.x) + gF.y * (1 - s.spec.y) + gF.z * (1 - s.spec.z);
        float gc = 0.f;
        if (1 - c > 0) gc = -gp5 * 5.0f * (om * om) * (om * om);
        // sc = D G / (4 wo.z)
        const float gD_ = gsc * G * 0.25f / wo.z;
        const float gG = gsc * D * 0.25f / wo.z;
        gwo.z += -gsc * sc / wo.z;
        // D(a2, hz)
        float ga2 = gD_ * (1 / (dd * dd * PI) - 2 * a2 * (h.z * h.z) / (dd * dd * dd * PI));
        float ghz = gD_ * (-2 * a2 / (dd * dd * dd * PI)) * (2 * h.z * (a2 - 1));
        // G = 1 / (1 + lI + lO)
        const float gl = -gG * G * G;
        {
            const float cs[2] = {wo.z, wi.z};
            float gcs[2] = {0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float cz = cs[k];
                if (cz > 0) {
                    const float c2 = cz * cz;
                    const float num = fmaxf(1 - c2, 0);
                    const float t = num / c2;
                    const float root = sqrtf(1 + a2 * t);
                    ga2 += gl * 0.25f * t / root;
                    if (1 - c2 > 0) gcs[k] = gl * (0.25f * a2 / root) * (-2.0f / (c2 * cz));
                }
            }
            gwo.z += gcs[0];
            gwi.z += gcs[1];
        }
        // c = dot(wo, h), hz = h.z, h = normalize(wo + wi)
        float3 gh = gc * wo;
        gh.z += ghz;
        gwo += gc * h;
        const float3 gsum = (gh - h * dot(h, gh)) / len;
        gwo += gsum;
        gwi += gsum;
        // alpha = rough^2 (frozen to 0 below the threshold, in which case this branch is not taken)
        gR += ga2 * 2 * alpha * 2 * rough;
        // spec = F0 (1-m) + kd m
        gKd += gspec * metallic;
        gM += gspec.x * (kd.x - F0) + gspec.y * (kd.y - F0) + gspec.z * (kd.z - F0);
    }
    // wo = frame(N)^T V, wi = frame(N)^T L
    const float3 gfx = gwo.x * V + gwi.x * L;
    const float3 gfy = gwo.y * V + gwi.y * L;
    const float3 gfz = gwo.z * V + gwi.z * L;
    gN += gfz;
    {
        const float sign = N.z > 0 ? 1.0f : -1.0f;
        const float a = -1.0f / (sign + N.z);
        const float gb = gfx.y * sign + gfy.x;
        const float ga = gfx.x * sign * N.x * N.x + gfy.y * N.y * N.y + gb * N.x * N.y;
        gN.x += gfx.x * sign * 2 * N.x * a + gb * N.y * a - gfx.z * sign;
        gN.y += gfy.y * 2 * N.y * a + gb * N.x * a - gfy.z;
        gN.z += ga * (a * a); // a = -1/(sign+nz)  =>  da/dnz = 1/(sign+nz)^2 = a^2
    }
    ShadeGrads r;
    r.gN = gN; r.gKd = gKd; r.gLi = gLi; r.gR = gR; r.gM = gM;
    return r;
}

MR_DEV void final_shading_bwd_px(const ShadeParams &p, int idx)
{
    const size_t i = (size_t)idx;
    ShadeGrads g;
    g.gN = g.gKd = g.gLi = f3(0.f);
    g.gR = g.gM = 0.f;
    if (MR_LDG(p.occ + i) > 0.1f && MR_LDG(p.fs_dist + i) > 0.f)
        g = final_shading_grads(load3(p.normal, i), load3(p.ray_dir, i), load3(p.kd, i), MR_LDG(p.rm + 2 * i), MR_LDG(p.rm + 2 * i + 1),
                                load3(p.fs_dir, i), load3(p.fs_Li, i), load3(p.g_color, i), load3(p.g_diff, i), load3(p.g_spec, i));
    store3(p.g_normal, i, g.gN);
    store3(p.g_kd, i, g.gKd);
    p.g_rm[2 * i] = g.gR;
    p.g_rm[2 * i + 1] = g.gM;
    store3(p.g_Li, i, g.gLi);
}

// The same gradients for the K shading passes of an spp loop in one pass: the K passes share the surface inputs and the
// upstream gradients (the loop adds their outputs, nerf/renderer_restir.py:443-459) and differ in the final sample
// (direction, distance, Li).  Per-pass gradients are added in the order the autograd engine would add them (last pass
// first).  g_Li[k] receives the radiance gradient of pass k; with `sum_gli` all passes add into g_Li[0] instead (valid
// when the K passes evaluated the SAME reservoir buffer, which is what the reference's saved aliases amount to,
// the env-gradient scatter is linear in grad_Li).
#define MR_SHADE_MULTI_MAX 16
struct ShadeMultiParams {
    const float *fs_dir[MR_SHADE_MULTI_MAX];
    const float *fs_dist[MR_SHADE_MULTI_MAX];
    const float *fs_Li[MR_SHADE_MULTI_MAX];
    float *g_Li[MR_SHADE_MULTI_MAX];
    int K, sum_gli, accumulate;
    int same_sample; // every pass has the same (fs_dir, fs_dist): the gradients are linear in Li, one evaluation on the summed Li
    float g_div; // != 0: upstream gradients are divided by it first (the loop's `total / mFrameIndex`)
    const float *__restrict__ occ;
    const float *__restrict__ normal;
    const float *__restrict__ ray_dir;
    const float *__restrict__ kd;
    const float *__restrict__ rm;
    const float *__restrict__ g_color; // may be null (= zeros)
    const float *__restrict__ g_diff;
    const float *__restrict__ g_spec;
    float *g_normal, *g_kd, *g_rm;
};
MR_DEV void final_shading_bwd_multi_px(const ShadeMultiParams &p, int idx)
{
    const size_t i = (size_t)idx;
    float3 aN = f3(0.f), aKd = f3(0.f), aLi = f3(0.f);
    float aR = 0.f, aM = 0.f;
    if (p.accumulate) {
        aN = load3_rw(p.g_normal, i);
        aKd = load3_rw(p.g_kd, i);
        aR = p.g_rm[2 * i];
        aM = p.g_rm[2 * i + 1];
        if (p.sum_gli) aLi = load3_rw(p.g_Li[0], i);
    }
    const bool lit = MR_LDG(p.occ + i) > 0.1f;
    float3 N = f3(0.f), rd = f3(0.f), kd = f3(0.f), gC = f3(0.f), gD = f3(0.f), gS = f3(0.f);
    float rough = 0.f, metallic = 0.f;
    if (lit) {
        N = load3(p.normal, i); rd = load3(p.ray_dir, i); kd = load3(p.kd, i);
        rough = MR_LDG(p.rm + 2 * i); metallic = MR_LDG(p.rm + 2 * i + 1);
        if (p.g_color) gC = load3(p.g_color, i);
        gD = load3(p.g_diff, i); gS = load3(p.g_spec, i);
        if (p.g_div != 0.0f) {
            gC = make_float3(div_by_scalar(gC.x, p.g_div), div_by_scalar(gC.y, p.g_div), div_by_scalar(gC.z, p.g_div));
            gD = make_float3(div_by_scalar(gD.x, p.g_div), div_by_scalar(gD.y, p.g_div), div_by_scalar(gD.z, p.g_div));
            gS = make_float3(div_by_scalar(gS.x, p.g_div), div_by_scalar(gS.y, p.g_div), div_by_scalar(gS.z, p.g_div));
        }
    }
    if (p.same_sample && p.sum_gli) {
        // The passes saved aliases of ONE final-sample buffer (the reference's behaviour): same direction, same distance,
        // different radiance.  Every gradient but grad_Li is linear in Li and grad_Li does not depend on it, so the K passes
        // collapse into one evaluation on the summed radiance (rounding differs from K separate evaluations by ~1e-7).
        if (lit && MR_LDG(p.fs_dist[0] + i) > 0.f) {
            float3 Li = f3(0.f);
            for (int k = p.K - 1; k >= 0; --k) Li += load3(p.fs_Li[k], i);
            const ShadeGrads g = final_shading_grads(N, rd, kd, rough, metallic, load3(p.fs_dir[0], i), Li, gC, gD, gS);
            aN += g.gN;
            aKd += g.gKd;
            aR += g.gR;
            aM += g.gM;
            aLi += g.gLi * (float)p.K;
        }
        store3(p.g_normal, i, aN);
        store3(p.g_kd, i, aKd);
        p.g_rm[2 * i] = aR;
        p.g_rm[2 * i + 1] = aM;
        store3(p.g_Li[0], i, aLi);
        return;
    }
    for (int k = p.K - 1; k >= 0; --k) {
        float3 gLi = f3(0.f);
        if (lit && MR_LDG(p.fs_dist[k] + i) > 0.f) {
            const ShadeGrads g = final_shading_grads(N, rd, kd, rough, metallic, load3(p.fs_dir[k], i), load3(p.fs_Li[k], i), gC, gD, gS);
            aN += g.gN;
            aKd += g.gKd;
            aR += g.gR;
            aM += g.gM;
            gLi = g.gLi;
        }
        if (p.sum_gli) aLi += gLi;
        else store3(p.g_Li[k], i, gLi);
    }
    store3(p.g_normal, i, aN);
    store3(p.g_kd, i, aKd);
    p.g_rm[2 * i] = aR;
    p.g_rm[2 * i + 1] = aM;
    if (p.sum_gli) store3(p.g_Li[0], i, aLi);
}

// ---------------------------------------------------------------------------------------------------------------
struct BounceParams {
    BvhView bvh;
    EnvView env;
    unsigned int frame, bounce_count;
    int max_bounce;
    int fx, fy;
    const float *__restrict__ occ;
    const float *__restrict__ pos_map;
    const float *__restrict__ normal;
    const float *__restrict__ ray_dir;
    float *prd; // [N,5] throughput rgb, specular flag, stop flag
    const float *__restrict__ kd;
    const float *__restrict__ rm;
    float *__restrict__ color;      // shade only
    float *__restrict__ diff_color; // shade only
    float *__restrict__ spec_color; // shade only
    float *__restrict__ new_pos;
    float *__restrict__ new_ray_d;
    float *__restrict__ new_occ;
    float *__restrict__ new_normal;
    Workspace ws;   // shadow-ray slots (2 per active pixel), continuation-ray slot (1 per active pixel), per-pixel scratch
    int sig_in, sig_out; // signatures of this call and of the call that follows it on the same path state (bounce_item)
};

// Which foreground pixel does item j of a path-kernel launch work on?  At the first vertex every foreground pixel carries a
// path; at vertex b >= 1 only the paths whose continuation ray found something (or left the scene on a specular bounce) are
// alive -- 13 % and 3 % of the foreground at the second and third vertex of config C2 -- and a launch over all foreground
// pixels ran 4.5 of 32 lanes (ncu, profiles/r3z_ncu_bounce_shade_gen_px.txt).  The resolve pass of vertex b - 1 therefore
// appends every path it keeps alive to a list, and the kernels of vertex b walk that list.  The list is only trusted if
// the workspace says it was written for THIS call (same path-state buffer, this vertex number): any other calling
// sequence -- a vertex evaluated on its own, a repeated call -- finds a foreign signature and falls back to all foreground
// pixels, for which the stop flag decides as in the reference (FinalShading.slang:657-690).  Dead paths have nothing to
// write: the prologue has zeroed their outputs and raised their stop flag.
MR_DEV int bounce_item(const BounceParams &p, int j)
{
    if (p.bounce_count > 0u && p.ws.counters[MR_CTR_ALIVE_SIG + (int)(p.bounce_count & 1u)] == p.sig_in) {
        const int k = (int)((p.bounce_count - 1u) & 1u);
        return j < p.ws.counters[MR_CTR_ALIVE_SIZE + k] ? p.ws.alive[k][j] : -1;
    }
    return j < p.ws.counters[MR_CTR_ACTIVE] ? j : -1;
}
// the resolve pass of vertex b signs the list for vertex b + 1 in the OTHER signature word: its own threads are still
// comparing the word of vertex b
MR_DEV void bounce_sign(const BounceParams &p) { p.ws.counters[MR_CTR_ALIVE_SIG + (int)((p.bounce_count + 1u) & 1u)] = p.sig_out; }
MR_DEV void bounce_keep_alive(const BounceParams &p, int a)
{
    const int k = (int)(p.bounce_count & 1u);
    p.ws.alive[k][queue_alloc(p.ws.counters + MR_CTR_ALIVE_SIZE + k)] = a;
}

// Entry sequence of both bounce kernels for EVERY pixel of the frame (FinalShading.slang:132-148, 657-690): remember
// the incoming stop flag, clear new_occ, raise the stop flag, reset the path state at bounce 0, zero the outputs.
MR_DEV void bounce_prologue_px(const BounceParams &p, int idx)
{
    const size_t i = (size_t)idx;
    if (idx == 0) p.ws.counters[MR_CTR_ALIVE_SIZE + (int)(p.bounce_count & 1u)] = 0; // the list this vertex's resolve pass fills
    p.ws.stop_in[i] = p.bounce_count == 0 ? 0.f : p.prd[5 * i + 4];
    p.new_occ[i] = 0.f;
    p.prd[5 * i + 4] = 1.f;
    if (p.bounce_count == 0) {
        p.prd[5 * i] = 1.f; p.prd[5 * i + 1] = 1.f; p.prd[5 * i + 2] = 1.f;
        p.prd[5 * i + 3] = 0.f;
    }
    if (p.color) {
        store3(p.color, i, f3(0.f));
        store3(p.diff_color, i, f3(0.f));
        store3(p.spec_color, i, f3(0.f));
    }
}

// sample a continuation direction and queue its closest-hit ray (FinalShading.slang:190-262 and :907-977)
MR_DEV void continue_path_gen(const BounceParams &p, int a, size_t i, const Surface &s, float3 surf_pos, uint32_t &sg, float3 thr)
{
    float3 out_dir, out_weight;
    float out_pdf;
    uint32_t sampledSpecular;
    bool valid = bsdf_sample<true>(s, sg, out_dir, out_pdf, sampledSpecular, out_weight);
    if (!valid) return;
    if (is_black(out_weight) || out_pdf == 0.f) {
        p.prd[5 * i + 4] = 1.f;
    } else if (p.bounce_count + 1u <= (unsigned int)p.max_bounce) {
        out_dir = normalize(from_frame(s.frame, out_dir));
        queue_closest_ray(p.ws, (size_t)a, surf_pos + VIS_NEAR * out_dir, out_dir);
        thr *= out_weight;
        p.prd[5 * i + 0] = thr.x;
        p.prd[5 * i + 1] = thr.y;
        p.prd[5 * i + 2] = thr.z;
        p.prd[5 * i + 3] = (float)sampledSpecular;
        store3(p.new_ray_d, i, out_dir);
    }
}

MR_DEV void continue_path_resolve(const BounceParams &p, int a)
{
    const float4 h0 = p.ws.chit[3 * (size_t)a];
    if (h0.w < 0.0f) return; // no continuation ray
    const size_t i = (size_t)p.ws.active[a];
    const float4 h1 = p.ws.chit[3 * (size_t)a + 1];
    if (h0.w != 0.0f) {
        p.prd[5 * i + 4] = 0.f;
        store3(p.new_pos, i, make_float3(h0.x, h0.y, h0.z));
        store3(p.new_normal, i, make_float3(h1.x, h1.y, h1.z));
        p.new_occ[i] = 1.f;
        bounce_keep_alive(p, a);
    } else if (p.prd[5 * i + 3] > 0.f) {
        p.prd[5 * i + 4] = 0.f; // a specular bounce that leaves the scene picks up the envmap in the next kernel
        bounce_keep_alive(p, a);
    }
}

MR_DEV void bounce_first_gen_px(const BounceParams &p, int j)
{
    const int a = bounce_item(p, j);
    if (a < 0) return;
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    queue_closest_empty(p.ws, (size_t)a);
    if (p.ws.stop_in[i] > 0.f) return;
    if (!(MR_LDG(p.occ + i) > 0.1f)) return;
    const float3 thr = make_float3(p.prd[5 * i], p.prd[5 * i + 1], p.prd[5 * i + 2]);
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    const Surface s = surface_of(load3(p.normal, i), load3(p.ray_dir, i), load3(p.kd, i), MR_LDG(p.rm + 2 * i), MR_LDG(p.rm + 2 * i + 1));
    continue_path_gen(p, a, i, s, load3(p.pos_map, i), sg, thr);
}

MR_DEV void bounce_first_resolve_px(const BounceParams &p, int j)
{
    if (j == 0) bounce_sign(p);
    const int a = bounce_item(p, j);
    if (a < 0) return;
    continue_path_resolve(p, a);
}

// per-active-pixel scratch of the shade kernel (floats): [0..8] contributions that need no ray (colour, diffuse, specular),
// [9..17] the light sample's contribution, [18..26] the BSDF sample's contribution, both pending their shadow ray
MR_DEV void put9(float *q, float3 a, float3 b, float3 c)
{
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = b.x; q[4] = b.y; q[5] = b.z; q[6] = c.x; q[7] = c.y; q[8] = c.z;
}

MR_DEV void bounce_shade_gen_px(const BounceParams &p, int j)
{
    const int a = bounce_item(p, j);
    if (a < 0) return;
    const int idx = p.ws.active[a];
    const size_t i = (size_t)idx;
    const uint32_t px = (uint32_t)(idx % p.fx), py = (uint32_t)(idx / p.fx);
    float *scratch = p.ws.px + (size_t)a * MR_PX_SCRATCH_FLOATS;
    queue_empty(p.ws, 2 * (size_t)a);
    queue_empty(p.ws, 2 * (size_t)a + 1);
    queue_closest_empty(p.ws, (size_t)a);
    put9(scratch, f3(0.f), f3(0.f), f3(0.f));
    if (p.ws.stop_in[i] > 0.f) return;
    const float3 thr = make_float3(p.prd[5 * i], p.prd[5 * i + 1], p.prd[5 * i + 2]);
    const float specularBounce = p.prd[5 * i + 3];
    const float3 rd = load3(p.ray_dir, i);
    if (!(MR_LDG(p.occ + i) > 0.1f)) {
        // the path left the scene: primary misses and specular chains collect the envmap (FinalShading.slang:979-996)
        if (p.bounce_count == 0) {
            put9(scratch, thr * env_radiance(p.env, ngp_dir(rd)), f3(0.f), f3(0.f));
        } else if (specularBounce > 0.f) {
            const float3 Le = env_radiance(p.env, ngp_dir(rd));
            put9(scratch, thr * Le, f3(0.f), thr * Le);
        }
        p.prd[5 * i + 4] = 1.f;
        return;
    }
    uint32_t sg = seed_of(px, row_of(p.ws, py), frame_of(p.ws, p.frame));
    const float3 N = load3(p.normal, i);
    const float3 P = load3(p.pos_map, i);
    const Surface s = surface_of(N, rd, load3(p.kd, i), MR_LDG(p.rm + 2 * i), MR_LDG(p.rm + 2 * i + 1));
    const bool has_normal = !is_black(N);
    // ---- next-event estimation: one env sample, power heuristic against the BSDF pdf
    float lightPdf = 0.0f, scatteringPdf = 0.0f;
    {
        float2 u;
        u.x = rnd(sg);
        u.y = rnd(sg);
        float3 sdir = f3(0.f), Li = f3(0.f);
        float spdf = 0.f;
        float2 luv;
        bool ok = sample_env(p.env, u, sdir, spdf, luv);
        if (ok) {
            lightPdf = spdf;
            Li = env_radiance(p.env, ngp_dir(sdir)) / spdf;
        }
        if (ok && lightPdf > 0 && !is_black(Li)) {
            float3 diff_f = f3(0.f), spec_f = f3(0.f), total_f = f3(0.f);
            const float3 wi = to_frame(s.frame, sdir);
            if (has_normal) {
                if (s.pD > 0.f) diff_f = f3(lambert_light(s.wo, wi));
                if (s.pS > 0.f) spec_f = specular_f(s.wo, wi, s.spec, s.alpha);
                total_f = s.kd_diff * diff_f + spec_f;
                diff_f = s.kd_diff * diff_f;
                scatteringPdf = bsdf_pdf(s, wi);
            }
            if (!is_black(total_f)) {
                const float3 ldir = normalize(sdir);
                queue_ray(p.ws, 2 * (size_t)a, P + VIS_NEAR * ldir, ldir);
                // the reference multiplies Li by the transmittance (1 when unoccluded) before using it
                Li = Li * f3(1.0f);
                const float mis = power_heuristic(lightPdf, scatteringPdf);
                put9(scratch + 9, thr * total_f * Li * mis, thr * diff_f * Li * mis, thr * spec_f * Li * mis);
            }
        }
    }
    // ---- BSDF sample with MIS against the light pdf
    if (has_normal) {
        float3 m_wi, unused_w;
        float m_pdf;
        uint32_t sampledSpecular;
        bool valid = bsdf_sample<false>(s, sg, m_wi, m_pdf, sampledSpecular, unused_w);
        if (valid) {
            float3 wd = f3(1.0f), wsp = f3(1.0f);
            if (s.pD > 0.f) wd = f3(lambert_light(s.wo, m_wi));
            if (s.pS > 0.f) wsp = specular_f(s.wo, m_wi, s.spec, s.alpha);
            const float3 wt = s.kd_diff * wd + wsp;
            m_wi = from_frame(s.frame, m_wi);
            scatteringPdf = m_pdf;
            // the reference divides by the pdf and multiplies it back (FinalShading.slang:853-859); kept for rounding
            float3 f = wt / m_pdf, diff_f = s.kd_diff * wd / m_pdf, spec_f = wsp / m_pdf;
            f *= m_pdf;
            diff_f *= m_pdf;
            spec_f *= m_pdf;
            const float3 dirw = normalize(m_wi);
            if (!is_black(f) && scatteringPdf > 0) {
                float weight = 1.0f;
                bool light_pdf_zero = false;
                if (sampledSpecular == 0) {
                    lightPdf = env_pdf(p.env, dirw);
                    if (lightPdf == 0.0f) light_pdf_zero = true;
                    weight = power_heuristic(scatteringPdf, lightPdf);
                }
                // the radiance that arrives if the ray escapes is looked up now; it only counts if the ray does escape
                const float3 Li = env_radiance(p.env, ngp_dir(dirw));
                if (!is_black(Li) && !light_pdf_zero) {
                    queue_ray(p.ws, 2 * (size_t)a + 1, P + VIS_NEAR * dirw, dirw);
                    const float3 Tr = f3(1.0f);
                    put9(scratch + 18, thr * f * Li * Tr * weight / scatteringPdf, thr * diff_f * Li * Tr * weight / scatteringPdf,
                         thr * spec_f * Li * Tr * weight / scatteringPdf);
                }
            }
        }
    }
    // ---- continuation
    continue_path_gen(p, a, i, s, P, sg, thr);
}

MR_DEV void bounce_shade_resolve_px(const BounceParams &p, int j)
{
    if (j == 0) bounce_sign(p);
    const int a = bounce_item(p, j);
    if (a < 0) return;
    const size_t i = (size_t)p.ws.active[a];
    const float *q = p.ws.px + (size_t)a * MR_PX_SCRATCH_FLOATS;
    float3 c = make_float3(q[0], q[1], q[2]), d = make_float3(q[3], q[4], q[5]), sp = make_float3(q[6], q[7], q[8]);
    if (p.ws.hit[2 * (size_t)a] == MR_HIT_MISS) {
        c += make_float3(q[9], q[10], q[11]);
        d += make_float3(q[12], q[13], q[14]);
        sp += make_float3(q[15], q[16], q[17]);
    }
    if (p.ws.hit[2 * (size_t)a + 1] == MR_HIT_MISS) {
        c += make_float3(q[18], q[19], q[20]);
        d += make_float3(q[21], q[22], q[23]);
        sp += make_float3(q[24], q[25], q[26]);
    }
    store3(p.color, i, c);
    store3(p.diff_color, i, d);
    store3(p.spec_color, i, sp);
    continue_path_resolve(p, a);
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_final_shading_fwd(const float *fs_dir, const float *fs_dist, const float *fs_Li, const float *env_tex,
                             int env_w, int env_h, int fx, int fy, const float *occ, const float *normal,
                             const float *ray_dir, const float *diffuse_map, const float *rough_metal, float *color,
                             float *diff_light, float *spec_light, void *stream)
{
    if (!fs_dir || !fs_dist || !fs_Li || !env_tex || !occ || !normal || !ray_dir || !diffuse_map || !rough_metal ||
        !color || !diff_light || !spec_light)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || env_w < 1 || env_h < 1) return MIRRES_ERR_SHAPE;
    ShadeParams p = {};
    p.fs_dir = fs_dir; p.fs_dist = fs_dist; p.fs_Li = fs_Li;
    p.env = {env_tex, env_w, env_h, nullptr, nullptr, nullptr, nullptr};
    p.occ = occ; p.normal = normal; p.ray_dir = ray_dir; p.kd = diffuse_map; p.rm = rough_metal;
    p.color = color; p.diff_light = diff_light; p.spec_light = spec_light;
    return foreach_item<ShadeParams, final_shading_px, 256>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_final_shading_bwd(const float *fs_dir, const float *fs_dist, const float *fs_Li, int fx, int fy,
                             const float *occ, const float *normal, const float *ray_dir, const float *diffuse_map,
                             const float *rough_metal, const float *grad_color, const float *grad_diff_light,
                             const float *grad_spec_light, float *grad_normal, float *grad_diffuse,
                             float *grad_rough_metal, float *grad_Li, void *stream)
{
    if (!fs_dir || !fs_dist || !fs_Li || !occ || !normal || !ray_dir || !diffuse_map || !rough_metal || !grad_color ||
        !grad_diff_light || !grad_spec_light || !grad_normal || !grad_diffuse || !grad_rough_metal || !grad_Li)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1) return MIRRES_ERR_SHAPE;
    ShadeParams p = {};
    p.fs_dir = fs_dir; p.fs_dist = fs_dist; p.fs_Li = fs_Li;
    p.occ = occ; p.normal = normal; p.ray_dir = ray_dir; p.kd = diffuse_map; p.rm = rough_metal;
    p.g_color = grad_color; p.g_diff = grad_diff_light; p.g_spec = grad_spec_light;
    p.g_normal = grad_normal; p.g_kd = grad_diffuse; p.g_rm = grad_rough_metal; p.g_Li = grad_Li;
    return foreach_item<ShadeParams, final_shading_bwd_px, 256>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_final_shading_bwd_multi(int n_passes, const float *const *fs_dir, const float *const *fs_dist,
                                   const float *const *fs_Li, int fx, int fy, const float *occ, const float *normal,
                                   const float *ray_dir, const float *diffuse_map, const float *rough_metal,
                                   const float *grad_color, const float *grad_diff_light, const float *grad_spec_light,
                                   float grad_divisor, int accumulate, float *grad_normal, float *grad_diffuse, float *grad_rough_metal,
                                   int sum_grad_Li, float *const *grad_Li, void *stream)
{
    if (!fs_dir || !fs_dist || !fs_Li || !occ || !normal || !ray_dir || !diffuse_map || !rough_metal || !grad_diff_light ||
        !grad_spec_light || !grad_normal || !grad_diffuse || !grad_rough_metal || !grad_Li)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || n_passes < 1 || n_passes > MR_SHADE_MULTI_MAX) return MIRRES_ERR_SHAPE;
    ShadeMultiParams p = {};
    for (int k = 0; k < n_passes; ++k) {
        if (!fs_dir[k] || !fs_dist[k] || !fs_Li[k] || (!sum_grad_Li && !grad_Li[k])) return MIRRES_ERR_NULL;
        p.fs_dir[k] = fs_dir[k]; p.fs_dist[k] = fs_dist[k]; p.fs_Li[k] = fs_Li[k];
        p.g_Li[k] = sum_grad_Li ? grad_Li[0] : grad_Li[k];
    }
    if (!p.g_Li[0]) return MIRRES_ERR_NULL;
    p.K = n_passes; p.sum_gli = sum_grad_Li ? 1 : 0; p.accumulate = accumulate ? 1 : 0; p.g_div = grad_divisor;
    p.same_sample = 1;
    for (int k = 1; k < n_passes; ++k)
        if (fs_dir[k] != fs_dir[0] || fs_dist[k] != fs_dist[0]) p.same_sample = 0;
    p.occ = occ; p.normal = normal; p.ray_dir = ray_dir; p.kd = diffuse_map; p.rm = rough_metal;
    p.g_color = grad_color; p.g_diff = grad_diff_light; p.g_spec = grad_spec_light;
    p.g_normal = grad_normal; p.g_kd = grad_diffuse; p.g_rm = grad_rough_metal;
    return foreach_item<ShadeMultiParams, final_shading_bwd_multi_px, 256>(p, fx * fy, (cudaStream_t)stream);
}

static int fill_bounce(BounceParams &p, const void *packed_nodes, const void *packed_tris, unsigned int frame_index,
                       unsigned int bounce_count, int max_bounce, int fx, int fy, const float *occ, const float *pos_map,
                       const float *normal, const float *ray_dir, float *prd, const float *diffuse_map,
                       const float *rough_metal, float *new_pos, float *new_ray_d, float *new_occ, float *new_normal,
                       void *workspace, size_t workspace_bytes)
{
    if (!packed_nodes || !packed_tris || !occ || !pos_map || !normal || !ray_dir || !prd || !diffuse_map || !rough_metal ||
        !new_pos || !new_ray_d || !new_occ || !new_normal || !workspace)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || max_bounce < 0) return MIRRES_ERR_SHAPE;
    if (new_pos == pos_map || new_occ == occ || new_normal == normal || new_ray_d == ray_dir) return MIRRES_ERR_ALIAS;
    if ((uintptr_t)workspace & 255) return MIRRES_ERR_ALIGN;
    const int n = fx * fy;
    if (workspace_bytes < workspace_carve(nullptr, n, nullptr)) return MIRRES_ERR_SCRATCH;
    workspace_carve(&p.ws, n, (char *)workspace);
    p.bvh = bvh_view(packed_nodes, packed_tris);
    p.frame = frame_index; p.bounce_count = bounce_count; p.max_bounce = max_bounce; p.fx = fx; p.fy = fy;
    p.occ = occ; p.pos_map = pos_map; p.normal = normal; p.ray_dir = ray_dir; p.prd = prd; p.kd = diffuse_map; p.rm = rough_metal;
    p.new_pos = new_pos; p.new_ray_d = new_ray_d; p.new_occ = new_occ; p.new_normal = new_normal;
    // never 0 (a cleared workspace matches nothing); the path state buffer and the vertex number identify the sequence
    const unsigned int h = (unsigned int)((uintptr_t)prd >> 4) * 2654435761u;
    p.sig_in = (int)((h ^ (bounce_count * 0x9e3779b9u)) | 1u);
    p.sig_out = (int)((h ^ ((bounce_count + 1u) * 0x9e3779b9u)) | 1u);
    return 0;
}

int mirres_bounce_first(const void *packed_nodes, const void *packed_tris, unsigned int frame_index,
                        unsigned int bounce_count, int max_bounce, int fx, int fy, const float *occ, const float *pos_map,
                        const float *normal, const float *ray_dir, float *prd, const float *diffuse_map,
                        const float *rough_metal, float *new_pos, float *new_ray_d, float *new_occ, float *new_normal,
                        void *workspace, size_t workspace_bytes, void *stream)
{
    BounceParams p = {};
    int rc = fill_bounce(p, packed_nodes, packed_tris, frame_index, bounce_count, max_bounce, fx, fy, occ, pos_map, normal,
                         ray_dir, prd, diffuse_map, rough_metal, new_pos, new_ray_d, new_occ, new_normal, workspace,
                         workspace_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = fx * fy;
    queue_reset(p.ws, st);
    if ((rc = foreach_item<BounceParams, bounce_prologue_px, 256>(p, n, st))) return rc;
    if ((rc = foreach_item<BounceParams, bounce_first_gen_px, 128>(p, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, p.ws, false, true, device_sm_count(), st))) return rc;
    return foreach_item<BounceParams, bounce_first_resolve_px, 256>(p, n, st);
}

int mirres_bounce_shade(const void *packed_nodes, const void *packed_tris, unsigned int frame_index,
                        unsigned int bounce_count, int max_bounce, int fx, int fy, const float *env_tex, int env_w,
                        int env_h, const float *pdf_, const float *cdf_, const float *mpdf_, const float *mcdf_,
                        const float *occ, const float *pos_map, const float *normal, const float *ray_dir, float *prd,
                        const float *diffuse_map, const float *rough_metal, float *color, float *diff_color,
                        float *spec_color, float *new_pos, float *new_ray_d, float *new_occ, float *new_normal,
                        void *workspace, size_t workspace_bytes, void *stream)
{
    if (!env_tex || !pdf_ || !cdf_ || !mpdf_ || !mcdf_ || !color || !diff_color || !spec_color) return MIRRES_ERR_NULL;
    if (env_w < 1 || env_h < 1) return MIRRES_ERR_SHAPE;
    BounceParams p = {};
    int rc = fill_bounce(p, packed_nodes, packed_tris, frame_index, bounce_count, max_bounce, fx, fy, occ, pos_map, normal,
                         ray_dir, prd, diffuse_map, rough_metal, new_pos, new_ray_d, new_occ, new_normal, workspace,
                         workspace_bytes);
    if (rc) return rc;
    p.env = {env_tex, env_w, env_h, pdf_, cdf_, mpdf_, mcdf_};
    p.color = color; p.diff_color = diff_color; p.spec_color = spec_color;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = fx * fy;
    queue_reset(p.ws, st);
    if ((rc = foreach_item<BounceParams, bounce_prologue_px, 256>(p, n, st))) return rc;
    if ((rc = foreach_item<BounceParams, bounce_shade_gen_px, 128>(p, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, p.ws, true, true, device_sm_count(), st))) return rc;
    return foreach_item<BounceParams, bounce_shade_resolve_px, 256>(p, n, st);
}

} // extern "C"
