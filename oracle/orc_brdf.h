// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).  PARITY UNPINNED.
//
// orc_brdf.h : BRDF evaluation / sampling.
//   nerf/ScreenSpaceReSTIR/utils/brdf.slang:1-211    (RIS target function side: evalBRDF, evalPdfBRDF, sampleBRDF)
//   nerf/ScreenSpaceReSTIR/utils/brdfDi.slang:1-457  (shading side: Diffuse_light, SpecularReflection_*, FalcorBRDF_*)
//   nerf/ScreenSpaceReSTIR/utils/helperDi.slang:1-40 (create_frame, frame_to_local/global)
#ifndef ORC_BRDF_H
#define ORC_BRDF_H

#include "orc_common.h"

namespace orc {

// brdf.slang:1-13
static inline f3 perp_stark(f3 u)
{
    f3 a = mk3(fabsf(u.x), fabsf(u.y), fabsf(u.z));
    uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0;
    uint32_t uzx = (a.x - a.z) < 0 ? 1 : 0;
    uint32_t uzy = (a.y - a.z) < 0 ? 1 : 0;
    uint32_t xm = uyx & uzx;
    uint32_t ym = (1 ^ xm) & uzy;
    uint32_t zm = 1 ^ (xm | ym);
    return normalize(cross(u, mk3((float)xm, (float)ym, (float)zm)));
}
// brdf.slang:15-20
static inline f3 to_local(f3 w, f3 N)
{
    f3 B = perp_stark(N);
    f3 T = cross(B, N);
    return mk3(dot(B, w), dot(T, w), dot(N, w));
}
// brdf.slang:110-115
static inline f3 to_global(f3 w, f3 N)
{
    f3 B = perp_stark(N);
    f3 T = cross(B, N);
    return B * w.x + T * w.y + N * w.z;
}
// brdf.slang:22-32 (pow(x,5) -> mr_pow5f, see include/mirres_fpmath.h)
static inline float fresnel_schlick(float f0, float f90, float cosTheta)
{
    return f0 + (f90 - f0) * mr_pow5f(smax(1 - cosTheta, 0));
}
static inline f3 fresnel_schlick3(f3 f0, float f90, float cosTheta)
{
    float p = mr_pow5f(smax(1 - cosTheta, 0));
    return mk3(f0.x + (f90 - f0.x) * p, f0.y + (f90 - f0.y) * p, f0.z + (f90 - f0.z) * p);
}
// brdf.slang:34-40
static inline float lambda_ggx(float alphaSqr, float cosTheta)
{
    if (cosTheta <= 0) return 0;
    float cosThetaSqr = cosTheta * cosTheta;
    float tanThetaSqr = smax(1 - cosThetaSqr, 0) / cosThetaSqr;
    return 0.5f * (-1 + sqrtf(1 + alphaSqr * tanThetaSqr));
}
// brdf.slang:42-49
static inline float ndf_ggx(float alpha, float cosTheta)
{
    const float M_PI_F = 3.141592653589793f;
    float a2 = alpha * alpha;
    float d = ((cosTheta * a2 - cosTheta) * cosTheta + 1);
    return a2 / (d * d * M_PI_F);
}
// brdf.slang:51-57
static inline float smith_separable(float alpha, float cosI, float cosO)
{
    float alphaSqr = alpha * alpha;
    float lI = lambda_ggx(alphaSqr, cosI), lO = lambda_ggx(alphaSqr, cosO);
    return 1 / ((1 + lI) * (1 + lO));
}
// brdf.slang:59-66
static inline float smith_correlated(float alpha, float cosI, float cosO)
{
    float alphaSqr = alpha * alpha;
    float lI = lambda_ggx(alphaSqr, cosI), lO = lambda_ggx(alphaSqr, cosO);
    return 1 / (1 + lI + lO);
}
// brdf.slang:68-71
static inline float pdf_ggx_ndf(float alpha, float cosTheta) { return ndf_ggx(alpha, cosTheta) * cosTheta; }
// brdf.slang:73-93
static inline f2 sample_disk_concentric(f2 u)
{
    const float M_PI_4_F = 0.785398163397448309616f;
    const float M_PI_2_F = 1.57079632679489661923f;
    u = mk2(2.f * u.x - 1.f, 2.f * u.y - 1.f);
    if (u.x == 0.f && u.y == 0.f) return u;
    float phi, r;
    if (fabsf(u.x) > fabsf(u.y)) { r = u.x; phi = (u.y / u.x) * M_PI_4_F; }
    else { r = u.y; phi = M_PI_2_F - (u.x / u.y) * M_PI_4_F; }
    float s, c;
    mr_sincosf(phi, &s, &c);
    return mk2(r * c, r * s);
}
// brdf.slang:95-102
static inline f3 sample_cosine_hemisphere_concentric(f2 u, float &pdf)
{
    const float M_1_PI_F = 0.31830988f;
    f2 d = sample_disk_concentric(u);
    float z = sqrtf(smax(0.f, 1.f - (d.x * d.x + d.y * d.y)));
    pdf = z * M_1_PI_F;
    return mk3(d.x, d.y, z);
}
// brdf.slang:117-129
static inline f3 sample_ggx_ndf(float alpha, f2 u, float &pdf)
{
    const float M_PI_F = 3.141592653589793f;
    float alphaSqr = alpha * alpha;
    float phi = u.y * (2 * M_PI_F);
    float tanThetaSqr = alphaSqr * u.x / (1 - u.x);
    float cosTheta = 1 / sqrtf(1 + tanThetaSqr);
    float r = sqrtf(smax(1 - cosTheta * cosTheta, 0));
    pdf = pdf_ggx_ndf(alpha, cosTheta);
    float s, c;
    mr_sincosf(phi, &s, &c);
    return mk3(c * r, s * r, cosTheta);
}

// ---- RIS side (brdf.slang:155-211) ------------------------------------------------------------
static inline float eval_brdf(f3 L, f3 V, f3 N, float ggxAlpha, float diffuseWeight, float specularWeight)
{
    const float M_1_PI_F = 0.31830988f;
    float weightSum = diffuseWeight + specularWeight;
    float mix = weightSum > 1e-7f ? (diffuseWeight / weightSum) : 1.f;
    float NdotV = saturate(dot(N, V));
    float NdotL = saturate(dot(N, L));
    f3 H = normalize(V + L);
    float NdotH = saturate(dot(N, H));
    float LdotH = saturate(dot(L, H));
    float D = ndf_ggx(ggxAlpha, NdotH);
    float G = smith_separable(ggxAlpha, NdotV, NdotL);
    float F = specularWeight < 1e-8f ? 0.f : fresnel_schlick(specularWeight, 1.f, LdotH) / specularWeight;
    float diffuse = NdotL * M_1_PI_F;
    float specular = smax(0.f, D * G * F / (4.f * NdotV));
    return NdotL > 0.f ? slerp(specular, diffuse, mix) : 0.f;
}
static inline float eval_pdf_brdf(bool specularOnly, f3 dir, f3 V, f3 N, float ggxAlpha, float diffuseWeight,
                                  float specularWeight)
{
    const float M_1_PI_F = 0.31830988f;
    float weightSum = diffuseWeight + specularWeight;
    float mix = weightSum > 1e-7f ? (diffuseWeight / weightSum) : 1.f;
    float cosTheta = saturate(dot(N, dir));
    float diffusePdf = specularOnly ? 0.f : cosTheta * M_1_PI_F;
    f3 h = normalize(to_local(dir + V, N));
    float specularPdf = pdf_ggx_ndf(ggxAlpha, h.z) / (4.f * saturate(dot(h, to_local(V, N))));
    return cosTheta > 0.f ? slerp(specularPdf, diffusePdf, mix) : 0.f;
}
static inline bool sample_brdf(bool specularOnly, f3 xi, f3 &dir, f3 V, f3 N, float ggxAlpha, float diffuseWeight,
                               float specularWeight)
{
    float weightSum = diffuseWeight + specularWeight;
    float mix = weightSum > 1e-7f ? (diffuseWeight / weightSum) : 1.f;
    dir = mk3(0.f);
    float pdf;
    if (xi.x < mix) {
        if (specularOnly) return false;
        dir = to_global(sample_cosine_hemisphere_concentric(mk2(xi.y, xi.z), pdf), N);
    } else {
        f3 h = sample_ggx_ndf(ggxAlpha, mk2(xi.y, xi.z), pdf);
        dir = reflect(-V, to_global(h, N));
    }
    return dot(N, dir) > 0.f;
}

// ---- shading side (helperDi.slang:1-40) -------------------------------------------------------
struct Frame { f3 x, y, z; };
static inline Frame create_frame(f3 normal)
{
    Frame fr;
    fr.z = normal;
    float sign = normal.z > 0 ? 1.0f : -1.0f; // copysignf(1, n.z) as defined at helperDi.slang:11-14
    const float a = -1.0f / (sign + normal.z);
    const float b = normal.x * normal.y * a;
    fr.x = mk3(1.0f + sign * normal.x * normal.x * a, sign * b, -sign * normal.x);
    fr.y = mk3(b, sign + normal.y * normal.y * a, -normal.y);
    return fr;
}
static inline f3 frame_to_local(const Frame &f, f3 v) { return mk3(dot(f.x, v), dot(f.y, v), dot(f.z, v)); }
static inline f3 frame_to_global(const Frame &f, f3 v) { return f.x * v.x + f.y * v.y + f.z * v.z; }

// brdfDi.slang:147-158 (first draw is burned)
static inline float diffuse_eval_pdf(f3 wo, f3 wi)
{
    const float M_1_PI_F = 0.31830988f;
    if (smin(wo.z, wi.z) < 1e-6f) return 0.f;
    return M_1_PI_F * wi.z;
}
// brdfDi.slang:134-145 -- NB the weight/albedo result is overwritten by every caller; only wi/pdf/valid matter
static inline bool diffuse_sample(f3 wo, f3 &wi, float &pdf, uint32_t &sg)
{
    next1d(sg);
    f2 u;
    u.x = next1d(sg);
    u.y = next1d(sg);
    wi = sample_cosine_hemisphere_concentric(u, pdf);
    if (smin(wo.z, wi.z) < 1e-6f) return false;
    return true;
}
// brdfDi.slang:160-168 (returns a scalar broadcast to float3)
static inline float diffuse_light(f3 wo, f3 wi)
{
    const float M_1_PI_F = 0.31830988f;
    if (smin(wo.z, wi.z) < 1e-6f) return 0.f;
    return smax(M_1_PI_F * wi.z, 0.0f);
}
// brdfDi.slang:170-192 with allowDeltaEval=false (every live call site)
static inline f3 specular_eval(f3 wo, f3 wi, f3 albedo, float alpha, bool activeLobes)
{
    if (smin(wo.z, wi.z) < 1e-6f) return mk3(0.f);
    if (alpha == 0.f) return mk3(0.f);
    if (!activeLobes) return mk3(0.f);
    f3 h = normalize(wo + wi);
    float woDotH = dot(wo, h);
    float D = ndf_ggx(alpha, h.z);
    float G = smith_correlated(alpha, wo.z, wi.z);
    f3 F = fresnel_schlick3(albedo, 1, woDotH);
    return F * D * G * 0.25f / wo.z;
}
// brdfDi.slang:194-214 with allowDeltaEval=false
static inline float specular_eval_pdf(f3 wo, f3 wi, float alpha, bool activeLobes)
{
    if (smin(wo.z, wi.z) < 1e-6f) return 0.f;
    if (alpha == 0.f) return 0.f;
    if (!activeLobes) return 0.f;
    f3 h = normalize(wo + wi);
    float woDotH = dot(wo, h);
    float pdf = pdf_ggx_ndf(alpha, h.z);
    return pdf / (4.f * woDotH);
}
// brdfDi.slang:216-257 with allowDeltaEval=false; weight is overwritten by callers, so not returned
static inline bool specular_sample(float alpha, f3 wo, f3 &wi, float &pdf, uint32_t &sg, bool activeLobes)
{
    wi = mk3(0.f);
    pdf = 0.f;
    if (wo.z < 1e-6f) return false;
    next1d(sg);
    if (alpha == 0.f) return false;
    if (!activeLobes) return false;
    f2 u;
    u.x = next1d(sg);
    u.y = next1d(sg);
    f3 h = sample_ggx_ndf(alpha, u, pdf);
    float woDotH = dot(wo, h);
    wi = 2.f * woDotH * h - wo;
    if (wi.z < 1e-6f) return false;
    pdf = specular_eval_pdf(wo, wi, alpha, activeLobes);
    return true;
}
// brdfDi.slang:259-266 with activeLobes=true, allowDeltaEval=false
static inline f3 falcor_eval(float pD, float pS, float alpha, f3 spec_albedo, f3 diff_albedo, f3 wo, f3 wi)
{
    const float M_1_PI_F = 0.31830988f;
    f3 result = mk3(0.f);
    if (pD > 0.f) {
        // DiffuseReflection_eval brdfDi.slang:122-130
        if (!(smin(wo.z, wi.z) < 1e-6f)) result += M_1_PI_F * diff_albedo * wi.z;
    }
    if (pS > 0.f) result += specular_eval(wo, wi, spec_albedo, alpha, true);
    return result;
}
// brdfDi.slang:268-275
static inline float falcor_eval_pdf(float pD, float pS, f3 wo, f3 wi, float alpha, bool activeLobes)
{
    float pdf = 0.f;
    if (pD > 0.f) pdf += pD * diffuse_eval_pdf(wo, wi);
    if (pS > 0.f) pdf += pS * specular_eval_pdf(wo, wi, alpha, activeLobes);
    return pdf;
}
// brdfDi.slang:277-329 (FalcorBRDF_sample) and :393-457 (FalcorBRDF_sample_no_weight, with_weight=false)
static inline bool falcor_sample(float pD, float pS, f3 wo, f3 &wi, float &pdf, uint32_t &specularBounce, f3 &weight,
                                 uint32_t &sg, float alpha, f3 spec_albedo, f3 diff_albedo, bool activeLobes,
                                 bool with_weight)
{
    wi = mk3(0.f);
    weight = mk3(0.f);
    pdf = 0.f;
    specularBounce = 0;
    bool valid = false;
    float uSelect = next1d(sg);
    if (uSelect < pD) {
        valid = diffuse_sample(wo, wi, pdf, sg);
        if (with_weight) weight = falcor_eval(pD, pS, alpha, spec_albedo, diff_albedo, wo, wi);
        pdf *= pD;
        if (pS > 0.f) pdf += pS * specular_eval_pdf(wo, wi, alpha, activeLobes);
        if (with_weight) weight = weight / pdf;
    } else if (uSelect < pD + pS) {
        valid = specular_sample(alpha, wo, wi, pdf, sg, activeLobes);
        if (with_weight) weight = falcor_eval(pD, pS, alpha, spec_albedo, diff_albedo, wo, wi);
        pdf *= pS;
        float test_roughness = sqrtf(alpha);
        if (test_roughness > 0.15f) {
            if (pD > 0.f) pdf += pD * diffuse_eval_pdf(wo, wi);
        } else {
            specularBounce = 1;
        }
        if (with_weight) weight = weight / pdf;
    }
    return valid;
}

// Lobe probabilities shared by FinalShading.slang:60-77, :172-188, :715-731
static inline void lobe_probs(f3 diffuse, float metallic, f3 specular, f3 ray_dir, f3 normal, float &pD, float &pS)
{
    const float specTrans = 0.f, diffTrans = 0.f;
    float diffuseWeight = luminance(diffuse);
    float dielectricBSDF = (1.f - metallic) * (1.f - specTrans);
    pD = diffuseWeight * dielectricBSDF * (1.f - diffTrans);
    float metallicBRDF = metallic;
    float specularWeight = luminance(fresnel_schlick3(specular, 1.f, dot(-ray_dir, normal)));
    pS = specularWeight * (metallicBRDF + dielectricBSDF);
    float normFactor = pD + pS;
    if (normFactor > 0.f) {
        normFactor = 1.f / normFactor;
        pD *= normFactor;
        pS *= normFactor;
    }
}

} // namespace orc
#endif
