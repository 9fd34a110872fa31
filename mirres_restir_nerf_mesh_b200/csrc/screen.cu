// mirres-b200: per-pixel stream kernels of the spp loop's host side.
//
// Between its Slang launches the reference's driver runs chains of small elementwise torch kernels on full-frame
// tensors: the material lookup + torch.where merge between bounces (nerf/renderer_restir.py:398-408, 428-438), the
// running sums of the per-iteration outputs and their division by the frame count (:443-459, :505-515), and the
// final composite (:543-549).  On B200 a step is launch-bound in those stretches (2-5 us per launch, ~25 launches per
// material query, 24 adds at the end of the loop), so each chain is ONE kernel here with the same operations in the
// same order (fp32, no contraction), i.e. bit-identical results:
//   mirres_material_procedural   procedural kd / roughness / metallic texture evaluated at the path vertices and merged
//                                into the material maps under the occupancy mask (the in-kernel texture fetch of 8f-4)
//   mirres_sum_images            dst = (((dst|0) + src_0) + src_1) + ... [/ divisor]
//   mirres_composite_fwd/_bwd    final_color = nan_to_num(where(occ <= 0.1, 1, kd (1 - metallic) dd + ds + di))
#include <float.h>
#include "mr_common.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

MR_DEV float clamp_t(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); } // torch.clamp

// triangle wave in [0,1]: |2 frac(x) - 1| from exactly-rounded operations (synth.material / ProceduralMaterial)
MR_DEV float tri_wave(float x)
{
    const float f = x - floorf(x);
    return fabsf(2.0f * f - 1.0f);
}

struct MatParams {
    const float *__restrict__ pos; // [n,3]
    const float *__restrict__ occ; // [n] or null
    float *kd;                     // [n,3]
    float *rs;                     // [n,2] roughness, metallic
    float metallic;
    int mode;                      // 0: value * occ (occ null: value), 1: merge where occ >= 0.5
    int use_scale;
    float sx, sy, sz;
};
MR_DEV void material_px(const MatParams &p, int idx)
{
    const size_t i = (size_t)idx;
    const float3 x = load3(p.pos, i);
    float3 kd = make_float3(0.1f + 0.8f * tri_wave(x.x * 1.7f), 0.1f + 0.8f * tri_wave(x.y * 1.7f), 0.1f + 0.8f * tri_wave(x.z * 1.7f));
    float rough = 0.08f + 0.92f * tri_wave(x.x * 0.8f + x.y * 0.5f);
    float met = p.metallic;
    if (p.mode == 0) {
        if (p.occ) {
            const float o = MR_LDG(p.occ + i);
            kd = kd * o;
            rough = rough * o;
            met = met * o;
        }
        store3(p.kd, i, kd);
        p.rs[2 * i] = rough;
        p.rs[2 * i + 1] = met;
        return;
    }
    const bool hit = MR_LDG(p.occ + i) >= 0.5f;
    if (hit) {
        if (p.use_scale) kd = make_float3(kd.x * p.sx, kd.y * p.sy, kd.z * p.sz);
        p.rs[2 * i] = rough;
        p.rs[2 * i + 1] = met;
    } else {
        kd = load3_rw(p.kd, i);
    }
    if (p.use_scale) kd = make_float3(clamp_t(kd.x, 0.f, 1.f), clamp_t(kd.y, 0.f, 1.f), clamp_t(kd.z, 0.f, 1.f));
    if (hit || p.use_scale) store3(p.kd, i, kd);
}

#define MR_SUM_MAX 32
struct SumParams {
    const float *src[MR_SUM_MAX];
    float *dst;
    int n_src, accumulate;
    float divisor; // 0: no division
};
MR_DEV void sum_item(const SumParams &p, int idx)
{
    const size_t i = (size_t)idx;
    float acc = p.accumulate ? p.dst[i] : 0.0f;
    for (int k = 0; k < p.n_src; ++k) acc += MR_LDG(p.src[k] + i);
    if (p.divisor != 0.0f) acc = div_by_scalar(acc, p.divisor);
    p.dst[i] = acc;
}

struct CompositeParams {
    const float *__restrict__ occ; // [n]
    const float *__restrict__ kd;  // [n,3]
    const float *__restrict__ rm;  // [n,2]
    const float *__restrict__ dd;  // [n,3] denoised diffuse light
    const float *__restrict__ ds;  // [n,3] denoised specular light
    const float *__restrict__ di;  // [n,3] denoised indirect light
    float *__restrict__ out;       // [n,3]
    const float *__restrict__ g_out;
    float *__restrict__ g_kd, *__restrict__ g_rm, *__restrict__ g_dd, *__restrict__ g_ds;
};
MR_DEV float nan_to_num0(float x)
{
    if (x != x) return 0.0f;
    if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
    return x;
}
MR_DEV void composite_px(const CompositeParams &p, int idx)
{
    const size_t i = (size_t)idx;
    float3 c = f3(1.0f);
    if (!(MR_LDG(p.occ + i) <= 0.1f)) {
        const float om = 1.0f - MR_LDG(p.rm + 2 * i + 1);
        const float3 diffuse = load3(p.kd, i) * om;
        c = diffuse * load3(p.dd, i) + load3(p.ds, i) + load3(p.di, i);
    }
    store3(p.out, i, make_float3(nan_to_num0(c.x), nan_to_num0(c.y), nan_to_num0(c.z)));
}
MR_DEV bool finite_f(float x) { return !(x != x) && !isinf(x); }
MR_DEV void composite_bwd_px(const CompositeParams &p, int idx)
{
    const size_t i = (size_t)idx;
    float3 gkd = f3(0.f), gdd = f3(0.f), gds = f3(0.f);
    float gm = 0.f;
    if (!(MR_LDG(p.occ + i) <= 0.1f)) {
        const float om = 1.0f - MR_LDG(p.rm + 2 * i + 1);
        const float3 kd = load3(p.kd, i), dd = load3(p.dd, i);
        const float3 diffuse = kd * om;
        const float3 c = diffuse * dd + load3(p.ds, i) + load3(p.di, i);
        float3 g = load3(p.g_out, i);
        g = make_float3(finite_f(c.x) ? g.x : 0.f, finite_f(c.y) ? g.y : 0.f, finite_f(c.z) ? g.z : 0.f);
        gds = g;
        gdd = g * diffuse;
        const float3 gdiff = g * dd;
        gkd = gdiff * om;
        const float3 t = gdiff * kd;
        gm = -((t.x + t.y) + t.z);
    }
    store3(p.g_kd, i, gkd);
    p.g_rm[2 * i] = 0.f;
    p.g_rm[2 * i + 1] = gm;
    store3(p.g_dd, i, gdd);
    store3(p.g_ds, i, gds);
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_material_procedural(int n, const float *pos, const float *occ, int mode, float metallic, const float *scale_xyz,
                               float *kd, float *rough_metal, void *stream)
{
    if (!pos || !kd || !rough_metal) return MIRRES_ERR_NULL;
    if (mode != 0 && mode != 1) return MIRRES_ERR_SHAPE;
    if (mode == 1 && !occ) return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    MatParams p = {pos, occ, kd, rough_metal, metallic, mode, scale_xyz ? 1 : 0,
                   scale_xyz ? scale_xyz[0] : 1.f, scale_xyz ? scale_xyz[1] : 1.f, scale_xyz ? scale_xyz[2] : 1.f};
    return foreach_item<MatParams, material_px, 256>(p, n, (cudaStream_t)stream);
}

int mirres_sum_images(int n_floats, int n_src, const float *const *src, float divisor, int accumulate, float *dst, void *stream)
{
    if (!dst || (n_src > 0 && !src)) return MIRRES_ERR_NULL;
    if (n_floats < 0 || n_src < 0 || n_src > MR_SUM_MAX) return MIRRES_ERR_SHAPE;
    if (n_floats == 0) return 0;
    SumParams p = {};
    for (int k = 0; k < n_src; ++k) {
        if (!src[k]) return MIRRES_ERR_NULL;
        if (src[k] == dst) return MIRRES_ERR_ALIAS;
        p.src[k] = src[k];
    }
    p.dst = dst; p.n_src = n_src; p.accumulate = accumulate ? 1 : 0; p.divisor = divisor;
    return foreach_item<SumParams, sum_item, 256>(p, n_floats, (cudaStream_t)stream);
}

int mirres_composite_fwd(int n, const float *occ, const float *diffuse_map, const float *rough_metal, const float *denoised_diffuse,
                         const float *denoised_spec, const float *denoised_indirect, float *final_color, void *stream)
{
    if (!occ || !diffuse_map || !rough_metal || !denoised_diffuse || !denoised_spec || !denoised_indirect || !final_color)
        return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    CompositeParams p = {};
    p.occ = occ; p.kd = diffuse_map; p.rm = rough_metal; p.dd = denoised_diffuse; p.ds = denoised_spec; p.di = denoised_indirect;
    p.out = final_color;
    return foreach_item<CompositeParams, composite_px, 256>(p, n, (cudaStream_t)stream);
}

int mirres_composite_bwd(int n, const float *occ, const float *diffuse_map, const float *rough_metal, const float *denoised_diffuse,
                         const float *denoised_spec, const float *denoised_indirect, const float *grad_final_color,
                         float *grad_diffuse_map, float *grad_rough_metal, float *grad_denoised_diffuse,
                         float *grad_denoised_spec, void *stream)
{
    if (!occ || !diffuse_map || !rough_metal || !denoised_diffuse || !denoised_spec || !denoised_indirect || !grad_final_color ||
        !grad_diffuse_map || !grad_rough_metal || !grad_denoised_diffuse || !grad_denoised_spec)
        return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    CompositeParams p = {};
    p.occ = occ; p.kd = diffuse_map; p.rm = rough_metal; p.dd = denoised_diffuse; p.ds = denoised_spec; p.di = denoised_indirect;
    p.g_out = grad_final_color; p.g_kd = grad_diffuse_map; p.g_rm = grad_rough_metal; p.g_dd = grad_denoised_diffuse;
    p.g_ds = grad_denoised_spec;
    return foreach_item<CompositeParams, composite_bwd_px, 256>(p, n, (cudaStream_t)stream);
}

} // extern "C"
