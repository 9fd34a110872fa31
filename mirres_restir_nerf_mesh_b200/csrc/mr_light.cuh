// mirres-b200 environment light: equirect lookup, 2-D distribution sampling, solid-angle pdf.
//
// Reference semantics restated here:
//   nerf/ScreenSpaceReSTIR/utils/helper.slang:30-71      uv2xy, eval_bi (truncate, then clamp, then lerp)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:41-98     FindInterval_*, warp(_continue), pdf(_continue)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:119-209   env_le, InfiniteAreaLight_Sample_Li(_no_env)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:291-330   get_light_info, InfiniteAreaLight_pdf_li
// Transcendentals come from include/mirres_fpmath.h (the numerical contract shared with the oracle).
#pragma once
#include "mr_common.cuh"

namespace mr {

struct EnvView {
    const float *__restrict__ tex; // [H*W,3] (host already flipped it vertically)
    int W, H;
    const float *__restrict__ pdf;  // [H*W]      conditional mass per texel
    const float *__restrict__ cdf;  // [H*(W+1)]  per-row CDF with leading 0
    const float *__restrict__ mpdf; // [H]
    const float *__restrict__ mcdf; // [H+1]
};

struct Taps {
    int i00, i10, i01, i11; // texel indices (x0,y0) (x1,y0) (x0,y1) (x1,y1)
    float u, v;
};

MR_DEV Taps bilinear_taps(float2 uv, int W, int H)
{
    float x = uv.x * (float)W - 0.5f;
    float y = uv.y * (float)H - 0.5f;
    int x0 = to_int(x), y0 = to_int(y);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = clampi(x0, 0, W - 1);
    x1 = clampi(x1, 0, W - 1);
    y0 = clampi(y0, 0, H - 1);
    y1 = clampi(y1, 0, H - 1);
    Taps t;
    t.u = x - (float)x0;
    t.v = y - (float)y0;
    t.i00 = y0 * W + x0;
    t.i10 = y0 * W + x1;
    t.i01 = y1 * W + x0;
    t.i11 = y1 * W + x1;
    return t;
}

MR_DEV float3 env_bilinear(const EnvView &e, float2 uv)
{
    Taps t = bilinear_taps(uv, e.W, e.H);
    float3 t00 = load3(e.tex, t.i00), t01 = load3(e.tex, t.i10), t10 = load3(e.tex, t.i01), t11 = load3(e.tex, t.i11);
    float iu = 1.0f - t.u, iv = 1.0f - t.v;
    return (t00 * iu + t01 * t.u) * iv + (t10 * iu + t11 * t.u) * t.v;
}

// direction (already in env space) -> equirect uv; false when |sin(theta)| < 1e-4
MR_DEV bool env_uv_of(float3 dir, float2 &uv)
{
    const float TWO_PI = 6.2831853f;
    const float INV_TWO_PI = 0.1591549f;
    const float INV_PI = 0.31830988f;
    float theta = mr_acosf(dir.y);
    float sin_theta = mr_sinf(theta);
    if (fabsf(sin_theta) < 1e-4f) return false;
    float phi = mr_atan2f(dir.z, dir.x);
    if (phi < 0) phi += TWO_PI;
    uv = make_float2(phi * INV_TWO_PI, 1 - theta * INV_PI);
    return true;
}

MR_DEV float3 env_radiance(const EnvView &e, float3 dir)
{
    float2 uv;
    if (!env_uv_of(dir, uv)) return f3(0.f);
    return env_bilinear(e, uv);
}

// stored (octahedral) light sample -> world direction and emitted radiance
MR_DEV void light_of(const EnvView &e, float ox, float oy, float3 &emission, float3 &dir)
{
    dir = oct_decode(ox, oy);
    emission = env_radiance(e, ngp_dir(dir));
}

MR_DEV int upper_bound_minus_one(const float *__restrict__ table, int left, int right, float val)
{
    int l = left, r = right;
    while (l < r) {
        int mid = (l + r) / 2;
        if (MR_LDG(table + mid) <= val) l = mid + 1; else r = mid;
    }
    return clampi(l - left - 1, 0, right - left);
}

MR_DEV float texel_pdf(const EnvView &e, int row, int col)
{
    row = clampi(row, 0, e.H - 1);
    col = clampi(col, 0, e.W - 1);
    return MR_LDG(e.pdf + row * e.W + col) * MR_LDG(e.mpdf + row) * (float)e.W * (float)e.H;
}

// importance-sample the env map: world-space direction (before ngp_dir), solid-angle pdf, (u, 1-v)
MR_DEV bool sample_env(const EnvView &e, float2 rnd2, float3 &dir, float &pdf_out, float2 &uv_out)
{
    const float PI = 3.141592653589793f;
    float2 uv = rnd2;
    const int w_ = e.W, h_ = e.H;
    int row = upper_bound_minus_one(e.mcdf, 0, h_ + 1, uv.y);
    uv.y = clampf((uv.y - MR_LDG(e.mcdf + row)) / MR_LDG(e.mpdf + row), 0.0f, 1.0f);
    int row_start = row * (w_ + 1);
    int col = upper_bound_minus_one(e.cdf, row_start, row_start + (w_ + 1), uv.x);
    uv.x = clampf((uv.x - MR_LDG(e.cdf + row_start + col)) / MR_LDG(e.pdf + row * w_ + col), 0.0f, 1.0f);
    uv.x = clampf((uv.x + (float)col) / (float)w_, 0.0f, 1.0f);
    uv.y = clampf((uv.y + (float)row) / (float)h_, 0.0f, 1.0f);
    float pdf = texel_pdf(e, row, col);
    float theta = uv.y * PI, phi = uv.x * 2 * PI;
    float st, ct, sp, cp;
    mr_sincosf(theta, &st, &ct);
    mr_sincosf(phi, &sp, &cp);
    dir = make_float3(st * cp, ct, st * sp);
    if (fabsf(st) >= 1e-4f) pdf = pdf / (2 * PI * PI * st);
    else pdf = 0.0f;
    pdf_out = pdf;
    uv_out = make_float2(uv.x, 1 - uv.y);
    return !(pdf == 0);
}

MR_DEV int2 texel_of_uv(float2 uv, int W, int H)
{
    float x = uv.x * (float)W;
    float y = uv.y * (float)H;
    int x0 = x < 0.f ? to_int(x) - 1 : to_int(x);
    int y0 = y < 0.f ? to_int(y) - 1 : to_int(y);
    x0 = ((x0 % W) + W) % W;
    y0 = ((y0 % H) + H) % H;
    return make_int2(x0, y0);
}

// solid-angle pdf of an arbitrary direction (evaluated on the world direction, as the reference does)
MR_DEV float env_pdf(const EnvView &e, float3 dir)
{
    const float TWO_PI = 6.2831853f;
    const float INV_TWO_PI = 0.1591549f;
    const float INV_PI = 0.31830988f;
    const float PI = 3.141592653589793f;
    float wx = clampf(dir.x, -1.0f, 1.0f), wy = clampf(dir.y, -1.0f, 1.0f), wz = clampf(dir.z, -1.0f, 1.0f);
    float theta = mr_acosf(wy);
    float sin_theta = mr_sinf(theta);
    if (fabsf(sin_theta) < 1e-4f) return 0;
    float phi = mr_atan2f(wz, wx);
    if (phi < 0) phi += TWO_PI;
    int col = to_int(phi * INV_TWO_PI * (float)e.W);
    int row = to_int(theta * INV_PI * (float)e.H);
    return texel_pdf(e, row, col) / (2 * PI * PI * sin_theta);
}

} // namespace mr
