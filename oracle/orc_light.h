// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).  PARITY UNPINNED.
//
// orc_light.h : environment light lookup, Distribution2D sampling, pdf.
//   nerf/ScreenSpaceReSTIR/utils/helper.slang:46-71     eval_bi (truncate-then-clamp addressing)
//   nerf/ScreenSpaceReSTIR/utils/helper.slang:30-43     uv2xy
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:41-98    FindInterval_*, warp, warp_continue, pdf(_continue)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:119-132  env_le
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:150-209  InfiniteAreaLight_Sample_Li(_no_env)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:291-330  get_light_info, InfiniteAreaLight_pdf_li
#ifndef ORC_LIGHT_H
#define ORC_LIGHT_H

#include "orc_common.h"

namespace orc {

struct Env {
    const float *tex; // [H*W,3], already flipped vertically by the host (renderer_restir.py:311)
    int W, H;
    const float *pdf_;  // [H*W]
    const float *cdf_;  // [H*(W+1)]
    const float *mpdf_; // [H]
    const float *mcdf_; // [H+1]
};

static inline f3 env_texel(const Env &e, int x, int y)
{
    const float *p = e.tex + 3 * ((size_t)y * e.W + x);
    return mk3(p[0], p[1], p[2]);
}

// bilinear taps of helper.slang:46-71; shared by the forward lookup and the backward scatter
struct BiTaps { int x0, y0, x1, y1; float u, v; };
static inline BiTaps eval_bi_taps(f2 uv, int width, int height)
{
    float x = uv.x * (float)width - 0.5f;
    float y = uv.y * (float)height - 0.5f;
    int x0 = f2i(x), y0 = f2i(y);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = iclamp(x0, 0, width - 1);  // max(0,min(x0,w-1))
    x1 = iclamp(x1, 0, width - 1);
    y0 = iclamp(y0, 0, height - 1);
    y1 = iclamp(y1, 0, height - 1);
    BiTaps t;
    t.u = x - (float)x0;
    t.v = y - (float)y0;
    t.x0 = x0; t.y0 = y0; t.x1 = x1; t.y1 = y1;
    return t;
}
// helper.slang:2-7 math_lerp
static inline f3 math_lerp(f3 t00, f3 t01, f3 t10, f3 t11, float u, float v)
{
    return (t00 * (1.0f - u) + t01 * u) * (1.0f - v) + (t10 * (1.0f - u) + t11 * u) * v;
}
static inline f3 eval_bi(const Env &e, f2 uv)
{
    BiTaps t = eval_bi_taps(uv, e.W, e.H);
    return math_lerp(env_texel(e, t.x0, t.y0), env_texel(e, t.x1, t.y0), env_texel(e, t.x0, t.y1),
                     env_texel(e, t.x1, t.y1), t.u, t.v);
}
// helper.slang:30-43
static inline i2 uv2xy(f2 uv, int width, int height)
{
    float x = uv.x * (float)width;
    float y = uv.y * (float)height;
    int x0 = x < 0.f ? f2i(x) - 1 : f2i(x);
    int y0 = y < 0.f ? f2i(y) - 1 : f2i(y);
    x0 = ((x0 % width) + width) % width;
    y0 = ((y0 % height) + height) % height;
    i2 r = {x0, y0};
    return r;
}

// lightDi.slang:119-132.  Returns false (and uv untouched) on the |sin theta| < 1e-4 early-out.
static inline bool env_dir_to_uv(f3 dir, f2 &uv)
{
    const float TWO_PI = 6.2831853f;
    const float INV_TWO_PI = 0.1591549f;
    const float INV_PI = 0.31830988f;
    float theta = mr_acosf(dir.y);
    float sin_theta = mr_sinf(theta);
    if (fabsf(sin_theta) < 1e-4f) return false;
    float phi = mr_atan2f(dir.z, dir.x);
    if (phi < 0) phi += TWO_PI;
    uv = mk2(phi * INV_TWO_PI, 1 - theta * INV_PI);
    return true;
}
static inline f3 env_le(f3 dir, const Env &e)
{
    f2 uv;
    if (!env_dir_to_uv(dir, uv)) return mk3(0.f);
    return eval_bi(e, uv);
}
// lightDi.slang:291-298
static inline void get_light_info(const Env &e, f2 light_uv, f3 &emission, f3 &dir)
{
    dir = oct_decode(light_uv);
    emission = env_le(ngp_dir(dir), e);
}

// lightDi.slang:41-66: first index with table[idx] > val in [left,right), minus one, clamped
static inline int find_interval(const float *table, int left, int right, float val)
{
    int l = left, r = right;
    while (l < r) {
        int mid = (l + r) / 2;
        if (table[mid] <= val) l = mid + 1; else r = mid;
    }
    return iclamp(l - left - 1, 0, right - left);
}
// lightDi.slang:67-88 warp + warp_continue
static inline void warp_continue(const Env &e, f2 &uv, int &row, int &col)
{
    int w_ = e.W, h_ = e.H;
    row = find_interval(e.mcdf_, 0, h_ + 1, uv.y);
    uv.y = sclamp((uv.y - e.mcdf_[row]) / e.mpdf_[row], 0.0f, 1.0f);
    int row_start = row * (w_ + 1);
    int row_end = row_start + (w_ + 1);
    col = find_interval(e.cdf_, row_start, row_end, uv.x);
    int ic = row * (w_ + 1) + col;
    int ip = row * w_ + col;
    uv.x = sclamp((uv.x - e.cdf_[ic]) / e.pdf_[ip], 0.0f, 1.0f);
    uv.x = sclamp((uv.x + (float)col) / (float)w_, 0.0f, 1.0f);
    uv.y = sclamp((uv.y + (float)row) / (float)h_, 0.0f, 1.0f);
}
// lightDi.slang:89-98
static inline float pdf_continue(const Env &e, int row, int col)
{
    row = iclamp(row, 0, e.H - 1);
    col = iclamp(col, 0, e.W - 1);
    return e.pdf_[row * e.W + col] * e.mpdf_[row] * (float)e.W * (float)e.H;
}
// lightDi.slang:150-209 (both variants share this body; light.slang:105-137 differs only in the
// dead `weight` value).  Returns res; out: dir, pdf, uv (uv.x, 1-uv.y) as floats.
static inline bool sample_li(const Env &e, f2 rnd, f3 &dir, float &out_pdf, f2 &light_uv)
{
    f2 uv = rnd;
    const float PI = 3.141592653589793f;
    int row = 0, col = 0;
    warp_continue(e, uv, row, col);
    float pdf = pdf_continue(e, row, col);
    float theta = uv.y * PI, phi = uv.x * 2 * PI;
    float sin_theta, cos_theta, sin_phi, cos_phi;
    mr_sincosf(theta, &sin_theta, &cos_theta);
    mr_sincosf(phi, &sin_phi, &cos_phi);
    dir = mk3(sin_theta * cos_phi, cos_theta, sin_theta * sin_phi);
    if (fabsf(sin_theta) >= 1e-4f) pdf = pdf / (2 * PI * PI * sin_theta);
    else pdf = 0.0f;
    out_pdf = pdf;
    light_uv = mk2(uv.x, 1 - uv.y);
    if (pdf == 0) return false;
    return true;
}
// lightDi.slang:311-330 (sample_light_pdf :367-374 forwards to it)
static inline float pdf_li(const Env &e, f3 dir)
{
    const float TWO_PI = 6.2831853f;
    const float INV_TWO_PI = 0.1591549f;
    const float INV_PI = 0.31830988f;
    const float PI = 3.141592653589793f;
    f3 w = mk3(sclamp(dir.x, -1.0f, 1.0f), sclamp(dir.y, -1.0f, 1.0f), sclamp(dir.z, -1.0f, 1.0f));
    float theta = mr_acosf(w.y);
    float sin_theta = mr_sinf(theta);
    if (fabsf(sin_theta) < 1e-4f) return 0;
    float phi = mr_atan2f(w.z, w.x);
    if (phi < 0) phi += TWO_PI;
    int col = f2i(phi * INV_TWO_PI * (float)e.W);
    int row = f2i(theta * INV_PI * (float)e.H);
    return pdf_continue(e, row, col) / (2 * PI * PI * sin_theta);
}

} // namespace orc
#endif
