// mirres-b200 BRDF evaluation and sampling (GGX microfacet + Lambert, Schlick Fresnel).
//
// Reference semantics restated here:
//   nerf/ScreenSpaceReSTIR/utils/brdf.slang:1-211     scalar RIS target BRDF, its pdf and sampler
//   nerf/ScreenSpaceReSTIR/utils/brdfDi.slang:1-457   shading BRDF: Diffuse_light, SpecularReflection_*, FalcorBRDF_*
//   nerf/ScreenSpaceReSTIR/utils/helperDi.slang:1-40  shading frame
// Every live call site passes activeLobes = true and allowDeltaEval = false; those are folded in.
#pragma once
#include "mr_common.cuh"

namespace mr {

// ---- tangent frame used by the RIS-side functions (perp_stark) ------------------------------------
MR_DEV float3 perp_stark(float3 u)
{
    float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    unsigned int uyx = (ax - ay) < 0 ? 1u : 0u;
    unsigned int uzx = (ax - az) < 0 ? 1u : 0u;
    unsigned int uzy = (ay - az) < 0 ? 1u : 0u;
    unsigned int xm = uyx & uzx;
    unsigned int ym = (1u ^ xm) & uzy;
    unsigned int zm = 1u ^ (xm | ym);
    return normalize(cross(u, make_float3((float)xm, (float)ym, (float)zm)));
}
struct Basis { float3 B, T, N; };
MR_DEV Basis basis_of(float3 N)
{
    Basis b;
    b.N = N;
    b.B = perp_stark(N);
    b.T = cross(b.B, N);
    return b;
}
MR_DEV float3 to_local(const Basis &b, float3 w) { return make_float3(dot(b.B, w), dot(b.T, w), dot(b.N, w)); }
MR_DEV float3 to_global(const Basis &b, float3 w) { return b.B * w.x + b.T * w.y + b.N * w.z; }

// ---- microfacet terms ----------------------------------------------------------------------------------
MR_DEV float schlick(float f0, float f90, float cosTheta) { return f0 + (f90 - f0) * mr_pow5f(fmaxf(1 - cosTheta, 0)); }
MR_DEV float3 schlick3(float3 f0, float f90, float cosTheta)
{
    float p = mr_pow5f(fmaxf(1 - cosTheta, 0));
    return make_float3(f0.x + (f90 - f0.x) * p, f0.y + (f90 - f0.y) * p, f0.z + (f90 - f0.z) * p);
}
MR_DEV float ggx_lambda(float alphaSqr, float cosTheta)
{
    if (cosTheta <= 0) return 0;
    float c2 = cosTheta * cosTheta;
    float tan2 = fmaxf(1 - c2, 0) / c2;
    return 0.5f * (-1 + sqrtf(1 + alphaSqr * tan2));
}
MR_DEV float ggx_ndf(float alpha, float cosTheta)
{
    const float PI = 3.141592653589793f;
    float a2 = alpha * alpha;
    float d = ((cosTheta * a2 - cosTheta) * cosTheta + 1);
    return a2 / (d * d * PI);
}
MR_DEV float ggx_ndf_pdf(float alpha, float cosTheta) { return ggx_ndf(alpha, cosTheta) * cosTheta; }
MR_DEV float smith_separable(float alpha, float cI, float cO)
{
    float a2 = alpha * alpha;
    float lI = ggx_lambda(a2, cI), lO = ggx_lambda(a2, cO);
    return 1 / ((1 + lI) * (1 + lO));
}
MR_DEV float smith_correlated(float alpha, float cI, float cO)
{
    float a2 = alpha * alpha;
    float lI = ggx_lambda(a2, cI), lO = ggx_lambda(a2, cO);
    return 1 / (1 + lI + lO);
}

// ---- samplers ---------------------------------------------------------------------------------------------
MR_DEV float3 cosine_hemisphere(float u0, float u1, float &pdf)
{
    const float PI_4 = 0.785398163397448309616f;
    const float PI_2 = 1.57079632679489661923f;
    const float INV_PI = 0.31830988f;
    float ux = 2.f * u0 - 1.f, uy = 2.f * u1 - 1.f;
    float dx, dy;
    if (ux == 0.f && uy == 0.f) {
        dx = ux; dy = uy;
    } else {
        float phi, r;
        if (fabsf(ux) > fabsf(uy)) { r = ux; phi = (uy / ux) * PI_4; }
        else { r = uy; phi = PI_2 - (ux / uy) * PI_4; }
        float s, c;
        mr_sincosf(phi, &s, &c);
        dx = r * c; dy = r * s;
    }
    float z = sqrtf(fmaxf(0.f, 1.f - (dx * dx + dy * dy)));
    pdf = z * INV_PI;
    return make_float3(dx, dy, z);
}
MR_DEV float3 ggx_ndf_sample(float alpha, float u0, float u1, float &pdf)
{
    const float PI = 3.141592653589793f;
    float a2 = alpha * alpha;
    float phi = u1 * (2 * PI);
    float tan2 = a2 * u0 / (1 - u0);
    float cosTheta = 1 / sqrtf(1 + tan2);
    float r = sqrtf(fmaxf(1 - cosTheta * cosTheta, 0));
    pdf = ggx_ndf_pdf(alpha, cosTheta);
    float s, c;
    mr_sincosf(phi, &s, &c);
    return make_float3(c * r, s * r, cosTheta);
}

// ---- RIS target function pieces (brdf.slang:155-211) ------------------------------------------------
struct RisSurface {
    float3 N, V;       // shading normal, direction to the eye (= -ray_dir)
    float alpha;       // brdf_map.z  (clamped roughness squared)
    float kd_w, ks_w;  // brdf_map.x, brdf_map.y
    float mix;         // kd_w / (kd_w + ks_w) or 1
    // per-surface terms the reference re-derives for every candidate; hoisted here (same operations, same values)
    Basis basis;       // perp_stark frame of N
    float3 Vlocal;     // V in that frame
    float NdotV;       // saturate(N.V)
    float lambdaV;     // Smith lambda of the view direction
};
MR_DEV RisSurface ris_surface(float3 N, float3 ray_dir, float3 brdf)
{
    RisSurface s;
    s.N = N;
    s.V = -ray_dir;
    s.alpha = brdf.z;
    s.kd_w = brdf.x;
    s.ks_w = brdf.y;
    float sum = brdf.x + brdf.y;
    s.mix = sum > 1e-7f ? (brdf.x / sum) : 1.f;
    s.basis = basis_of(N);
    s.Vlocal = to_local(s.basis, s.V);
    s.NdotV = saturate(dot(N, s.V));
    s.lambdaV = ggx_lambda(s.alpha * s.alpha, s.NdotV);
    return s;
}
MR_DEV float ris_brdf(const RisSurface &s, float3 L)
{
    const float INV_PI = 0.31830988f;
    const float NdotV = s.NdotV;
    float NdotL = saturate(dot(s.N, L));
    if (s.ks_w < 1e-8f) {
        // No specular weight (the reference's default material, --me_max 0): F is the constant 0 below, so the specular
        // term is max(0, D G 0 / (4 N.V)) = +0 whatever D and G are (both finite and non-negative for alpha >= 1e-4; a
        // 0 / 0 at N.V = 0 is a NaN that max() drops), and the result is the lerp of +0 and the diffuse term: the same
        // bits without the half vector, the NDF and the shadowing term (a quarter of a candidate's instructions).
        const float diffuse0 = NdotL * INV_PI;
        return NdotL > 0.f ? lerpf(0.f, diffuse0, s.mix) : 0.f;
    }
    float3 H = normalize(s.V + L);
    float NdotH = saturate(dot(s.N, H));
    float LdotH = saturate(dot(L, H));
    float D = ggx_ndf(s.alpha, NdotH);
    float G = 1 / ((1 + s.lambdaV) * (1 + ggx_lambda(s.alpha * s.alpha, NdotL))); // separable Smith
    float F = s.ks_w < 1e-8f ? 0.f : schlick(s.ks_w, 1.f, LdotH) / s.ks_w;
    float diffuse = NdotL * INV_PI;
    float specular = fmaxf(0.f, D * G * F / (4.f * NdotV));
    return NdotL > 0.f ? lerpf(specular, diffuse, s.mix) : 0.f;
}
// target function p-hat = max(0, lum(Le) * brdf)   (res.slang:70-77)
MR_DEV float target_pdf(const RisSurface &s, float3 Le, float3 L) { return fmaxf(0.f, luminance(Le) * ris_brdf(s, L)); }

MR_DEV float ris_brdf_pdf(const RisSurface &s, float3 dir)
{
    const float INV_PI = 0.31830988f;
    float cosTheta = saturate(dot(s.N, dir));
    float diffusePdf = cosTheta * INV_PI;
    float3 h = normalize(to_local(s.basis, dir + s.V));
    float specularPdf = ggx_ndf_pdf(s.alpha, h.z) / (4.f * saturate(dot(h, s.Vlocal)));
    return cosTheta > 0.f ? lerpf(specularPdf, diffusePdf, s.mix) : 0.f;
}
MR_DEV bool ris_brdf_sample(const RisSurface &s, float x0, float x1, float x2, float3 &dir)
{
    float pdf;
    const Basis &b = s.basis;
    if (x0 < s.mix) {
        dir = to_global(b, cosine_hemisphere(x1, x2, pdf));
    } else {
        float3 h = ggx_ndf_sample(s.alpha, x1, x2, pdf);
        dir = reflect(-s.V, to_global(b, h));
    }
    return dot(s.N, dir) > 0.f;
}

// ---- shading-side BSDF (brdfDi.slang) -----------------------------------------------------------------------
struct Frame { float3 x, y, z; };
MR_DEV Frame frame_of(float3 n)
{
    Frame f;
    f.z = n;
    float sign = n.z > 0 ? 1.0f : -1.0f;
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    f.x = make_float3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    f.y = make_float3(b, sign + n.y * n.y * a, -n.y);
    return f;
}
MR_DEV float3 to_frame(const Frame &f, float3 v) { return make_float3(dot(f.x, v), dot(f.y, v), dot(f.z, v)); }
MR_DEV float3 from_frame(const Frame &f, float3 v) { return f.x * v.x + f.y * v.y + f.z * v.z; }

struct Surface {
    Frame frame;
    float3 wo;        // local direction to the previous vertex (the reference calls it wiLocal)
    float3 spec;      // F0 = 0.04 (1-m) + kd m
    float3 kd_diff;   // kd (1-m)
    float alpha;      // roughness^2, 0 below 1e-4
    float pD, pS;     // lobe selection probabilities
};
MR_DEV Surface surface_of(float3 N, float3 ray_dir, float3 kd, float roughness, float metallic)
{
    const float F0 = 0.04f;
    Surface s;
    s.spec = f3(F0) * (1.0f - metallic) + kd * metallic;
    s.kd_diff = kd * (1.0f - metallic);
    float alpha = roughness * roughness;
    const float kMinGGXAlpha = 0.01f * 0.01f;
    if (alpha < kMinGGXAlpha) alpha = 0.f;
    s.alpha = alpha;
    float diffuseWeight = luminance(kd);
    float dielectric = (1.f - metallic) * (1.f - 0.f);
    float pD = diffuseWeight * dielectric * (1.f - 0.f);
    float specularWeight = luminance(schlick3(s.spec, 1.f, dot(-ray_dir, N)));
    float pS = specularWeight * (metallic + dielectric);
    float norm = pD + pS;
    if (norm > 0.f) {
        norm = 1.f / norm;
        pD *= norm;
        pS *= norm;
    }
    s.pD = pD;
    s.pS = pS;
    s.frame = frame_of(N);
    s.wo = to_frame(s.frame, -ray_dir);
    return s;
}

MR_DEV float lambert_light(float3 wo, float3 wi)
{
    const float INV_PI = 0.31830988f;
    if (fminf(wo.z, wi.z) < 1e-6f) return 0.f;
    return fmaxf(INV_PI * wi.z, 0.0f);
}
MR_DEV float lambert_pdf(float3 wo, float3 wi)
{
    const float INV_PI = 0.31830988f;
    if (fminf(wo.z, wi.z) < 1e-6f) return 0.f;
    return INV_PI * wi.z;
}
MR_DEV float3 specular_f(float3 wo, float3 wi, float3 albedo, float alpha)
{
    if (fminf(wo.z, wi.z) < 1e-6f) return f3(0.f);
    if (alpha == 0.f) return f3(0.f);
    float3 h = normalize(wo + wi);
    float woDotH = dot(wo, h);
    float D = ggx_ndf(alpha, h.z);
    float G = smith_correlated(alpha, wo.z, wi.z);
    float3 F = schlick3(albedo, 1, woDotH);
    return F * D * G * 0.25f / wo.z;
}
MR_DEV float specular_pdf(float3 wo, float3 wi, float alpha)
{
    if (fminf(wo.z, wi.z) < 1e-6f) return 0.f;
    if (alpha == 0.f) return 0.f;
    float3 h = normalize(wo + wi);
    float woDotH = dot(wo, h);
    return ggx_ndf_pdf(alpha, h.z) / (4.f * woDotH);
}
// f = kd(1-m) lambert + specular, as FalcorBRDF_eval assembles it (brdfDi.slang:259-266)
MR_DEV float3 bsdf_f(const Surface &s, float3 wi)
{
    const float INV_PI = 0.31830988f;
    float3 r = f3(0.f);
    if (s.pD > 0.f) {
        if (!(fminf(s.wo.z, wi.z) < 1e-6f)) r += INV_PI * s.kd_diff * wi.z;
    }
    if (s.pS > 0.f) r += specular_f(s.wo, wi, s.spec, s.alpha);
    return r;
}
MR_DEV float bsdf_pdf(const Surface &s, float3 wi)
{
    float pdf = 0.f;
    if (s.pD > 0.f) pdf += s.pD * lambert_pdf(s.wo, wi);
    if (s.pS > 0.f) pdf += s.pS * specular_pdf(s.wo, wi, s.alpha);
    return pdf;
}

// One BSDF sample (FalcorBRDF_sample / _no_weight, brdfDi.slang:277-329,393-457).  RNG draws:
// 1 lobe select; diffuse: 1 burned + 2; specular: (return before drawing if wo.z < 1e-6) 1 burned, then 2 unless alpha == 0.
template <bool WITH_WEIGHT>
MR_DEV bool bsdf_sample(const Surface &s, uint32_t &sg, float3 &wi, float &pdf, uint32_t &specular_bounce, float3 &weight)
{
    wi = f3(0.f);
    weight = f3(0.f);
    pdf = 0.f;
    specular_bounce = 0;
    bool valid = false;
    float uSelect = rnd(sg);
    if (uSelect < s.pD) {
        rnd(sg);
        float u0 = rnd(sg);
        float u1 = rnd(sg);
        wi = cosine_hemisphere(u0, u1, pdf);
        valid = !(fminf(s.wo.z, wi.z) < 1e-6f);
        if (WITH_WEIGHT) weight = bsdf_f(s, wi);
        pdf *= s.pD;
        if (s.pS > 0.f) pdf += s.pS * specular_pdf(s.wo, wi, s.alpha);
        if (WITH_WEIGHT) weight = weight / pdf;
    } else if (uSelect < s.pD + s.pS) {
        if (!(s.wo.z < 1e-6f)) {
            rnd(sg);
            if (s.alpha != 0.f) {
                float u0 = rnd(sg);
                float u1 = rnd(sg);
                float3 h = ggx_ndf_sample(s.alpha, u0, u1, pdf);
                float woDotH = dot(s.wo, h);
                wi = 2.f * woDotH * h - s.wo;
                if (!(wi.z < 1e-6f)) {
                    pdf = specular_pdf(s.wo, wi, s.alpha);
                    valid = true;
                }
            }
        }
        if (WITH_WEIGHT) weight = bsdf_f(s, wi);
        pdf *= s.pS;
        float test_roughness = sqrtf(s.alpha);
        if (test_roughness > 0.15f) {
            if (s.pD > 0.f) pdf += s.pD * lambert_pdf(s.wo, wi);
        } else {
            specular_bounce = 1;
        }
        if (WITH_WEIGHT) weight = weight / pdf;
    }
    return valid;
}

} // namespace mr
