// mirres-b200: edge-avoiding a-trous wavelet denoiser (forward + backward) and the normal-variation AO proxy.
//
// Replaces:
//   nerf/ScreenSpaceReSTIR/EAWDenoise.slang:50-174    process_EAWDenoise (and :178-302 _no_di, same arithmetic)
//   its Slang-autodiff `.bwd` (nerf/ScreenSpaceReSTIR/Denoising.py:30-48)
//   nerf/ScreenSpaceReSTIR/EAWDenoise.slang:591-647   process_normal_ao (nerf/renderer.py:1153-1158)
//
// Backward design: Slang's reverse mode scatters 25 taps x 9 channels of atomics per pixel.  The edge-stopping
// weight w(q,r) and the 5x5 kernel are symmetric in (q,r), so the same gradient is a GATHER: every pixel q visits its
// 25 neighbours r once and collects (a) what it receives as the centre of its own filter footprint and (b) what it
// receives as a tap of r's footprint.  No atomics, deterministic, one pass after a normalisation pre-pass.
#include "mr_common.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

struct EawParams {
    float c_phi, n_phi, p_phi;
    int fx, fy, step;
    const float *__restrict__ occ;    // [N]
    const float *__restrict__ color;  // [N,3]
    const float *__restrict__ normal; // [N,3]
    const float *__restrict__ pos;    // [N,3]
    float *__restrict__ out_color;    // [N,3] (forward: written; backward: read)
    // backward
    float *__restrict__ cum_w;        // [N] normalisation of each footprint (pre-pass output)
    const float *__restrict__ g_out;  // [N,3]
    float *__restrict__ g_color;      // [N,3]
    float *__restrict__ g_normal;     // [N,3]
    float *__restrict__ g_pos;        // [N,3]
};

MR_DEV float eaw_kernel(int i)
{
    // separable binomial-like 5x5: outer product of (1/16, 1/4, 3/8, 1/4, 1/16)
    const int x = i % 5, y = i / 5;
    const float k1[5] = {1.0f, 4.0f, 6.0f, 4.0f, 1.0f};
    // the reference tabulates the products as exact binary fractions (EAWDenoise.slang:111-139)
    return (k1[x] * k1[y]) / 256.0f;
}

MR_DEV float edge_weight(float3 a, float3 b, float phi, bool clamp_zero)
{
    float3 t = a - b;
    float dist2 = dot(t, t);
    if (clamp_zero) dist2 = fmaxf(dist2, 0.0f);
    return fminf(mr_expf(-(dist2) / phi), 1.0f);
}

// forward footprint of pixel (px,py): returns sum and cum_w
MR_DEV void eaw_footprint(const EawParams &p, int px, int py, float3 cval, float3 nval, float3 pval, float3 &sum, float &cum_w)
{
    sum = f3(0.f);
    cum_w = 0.0f;
#pragma unroll 5
    for (int i = 0; i < 25; ++i) {
        const int ux = px + (i % 5 - 2) * p.step, uy = py + (i / 5 - 2) * p.step;
        if (!(ux >= 0 && ux < p.fx && uy >= 0 && uy < p.fy)) continue;
        const size_t r = (size_t)uy * p.fx + ux;
        const float3 ctmp = load3(p.color, r);
        const float c_w = edge_weight(cval, ctmp, p.c_phi, false);
        const float n_w = edge_weight(nval, load3(p.normal, r), p.n_phi, true);
        const float p_w = edge_weight(pval, load3(p.pos, r), p.p_phi, true);
        const float weight = c_w * n_w * p_w;
        const float k = eaw_kernel(i);
        sum += ctmp * weight * k;
        cum_w += weight * k;
    }
}

MR_DEV void eaw_fwd_px(const EawParams &p, int idx)
{
    const size_t q = (size_t)idx;
    const float3 cval = load3(p.color, q);
    if (MR_LDG(p.occ + q) < 0.1f) { store3(p.out_color, q, cval); return; }
    float3 sum;
    float cum_w;
    eaw_footprint(p, idx % p.fx, idx / p.fx, cval, load3(p.normal, q), load3(p.pos, q), sum, cum_w);
    store3(p.out_color, q, sum / cum_w);
}

// backward pre-pass: normalisation of every footprint
MR_DEV void eaw_norm_px(const EawParams &p, int idx)
{
    const size_t q = (size_t)idx;
    float cum_w = 0.f;
    if (!(MR_LDG(p.occ + q) < 0.1f)) {
        float3 sum;
        eaw_footprint(p, idx % p.fx, idx / p.fx, load3(p.color, q), load3(p.normal, q), load3(p.pos, q), sum, cum_w);
    }
    p.cum_w[q] = cum_w;
}

MR_DEV void eaw_bwd_px(const EawParams &p, int idx)
{
    const size_t q = (size_t)idx;
    const int px = idx % p.fx, py = idx / p.fx;
    const float3 cq = load3(p.color, q), nq = load3(p.normal, q), pq = load3(p.pos, q);
    const bool q_center = !(MR_LDG(p.occ + q) < 0.1f);
    const float3 gq = load3(p.g_out, q), oq = load3(p.out_color, q);
    const float Wq = p.cum_w[q];
    float3 gc = f3(0.f), gn = f3(0.f), gp = f3(0.f);
#pragma unroll 5
    for (int i = 0; i < 25; ++i) {
        const int ux = px + (i % 5 - 2) * p.step, uy = py + (i / 5 - 2) * p.step;
        if (!(ux >= 0 && ux < p.fx && uy >= 0 && uy < p.fy)) continue;
        const size_t r = (size_t)uy * p.fx + ux;
        const bool r_center = !(MR_LDG(p.occ + r) < 0.1f);
        if (!q_center && !r_center) continue; // neither footprint exists: nothing flows between q and r
        const float3 cr = load3(p.color, r), nr = load3(p.normal, r), pr = load3(p.pos, r);
        const float w = edge_weight(cq, cr, p.c_phi, false) * edge_weight(nq, nr, p.n_phi, true) * edge_weight(pq, pr, p.p_phi, true);
        const float k = eaw_kernel(i);
        const float3 dc = cq - cr, dn = nq - nr, dp = pq - pr;
        if (q_center) {
            // q is the centre, r its tap:  d out_q / d w = k (c_r - out_q) / W_q ;  dw/d|t|^2 = -w/phi ; d|t|^2/d c_q = 2 t
            const float gw = k * dot(gq, cr - oq) / Wq;
            const float s = -gw * w * 2.0f;
            gc += dc * (s / p.c_phi);
            gn += dn * (s / p.n_phi);
            gp += dp * (s / p.p_phi);
        }
        if (r_center) {
            // r is the centre, q its tap (kernel and weight are symmetric)
            const float3 gr = load3(p.g_out, r), orr = load3(p.out_color, r);
            const float Wr = p.cum_w[r];
            gc += gr * (w * k / Wr);
            const float gw = k * dot(gr, cq - orr) / Wr;
            const float s = -gw * w * 2.0f; // d|t|^2 / d c_q = -2 (c_r - c_q) = 2 (c_q - c_r)
            gc += dc * (s / p.c_phi);
            gn += dn * (s / p.n_phi);
            gp += dp * (s / p.p_phi);
        }
    }
    store3(p.g_color, q, gc);
    store3(p.g_normal, q, gn);
    store3(p.g_pos, q, gp);
}

// ---- batched variants: several colour images filtered over the SAME guide buffers in one pass --------------------------
// run_restir_di_with_pt denoises five images per a-trous level (diffuse, specular, indirect, indirect diffuse,
// indirect specular; nerf/renderer_restir.py:517-541) with identical occ / normal / pos.  The normal and position
// edge-stopping weights (two of the three exps per tap) and the guide loads are shared; per-image arithmetic and its
// order are unchanged, so every output is bit-identical to the single-image entry point.
#define MR_EAW_MAX_IMAGES 8
struct EawMultiParams {
    float c_phi, n_phi, p_phi;
    int fx, fy, step, n_images;
    const float *__restrict__ occ;
    const float *__restrict__ normal;
    const float *__restrict__ pos;
    const float *color[MR_EAW_MAX_IMAGES];
    float *out_color[MR_EAW_MAX_IMAGES]; // forward: written; backward: read
    float *cum_w[MR_EAW_MAX_IMAGES];     // forward: optional output; backward: input (normalisation of each footprint)
    const float *g_out[MR_EAW_MAX_IMAGES];
    float *g_color[MR_EAW_MAX_IMAGES];
    float *g_normal[MR_EAW_MAX_IMAGES];
    float *g_pos[MR_EAW_MAX_IMAGES];
};

// (tap loops are NOT unrolled in the batched kernels: unrolled by 5 they stalled on instruction fetch -- ncu
// stall_no_instruction 3.8 per issue in the backward, profiles/r1r_*)
template <int NI>
MR_DEV void eaw_fwd_multi_px(const EawMultiParams &p, int idx)
{
    const size_t q = (size_t)idx;
    float3 cval[NI];
#pragma unroll
    for (int m = 0; m < NI; ++m) cval[m] = load3(p.color[m], q);
    if (MR_LDG(p.occ + q) < 0.1f) {
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            store3(p.out_color[m], q, cval[m]);
            if (p.cum_w[m]) p.cum_w[m][q] = 0.f;
        }
        return;
    }
    const int px = idx % p.fx, py = idx / p.fx;
    const float3 nval = load3(p.normal, q), pval = load3(p.pos, q);
    float3 sum[NI];
    float cum[NI];
#pragma unroll
    for (int m = 0; m < NI; ++m) { sum[m] = f3(0.f); cum[m] = 0.f; }
#pragma unroll 1
    for (int i = 0; i < 25; ++i) {
        const int ux = px + (i % 5 - 2) * p.step, uy = py + (i / 5 - 2) * p.step;
        if (!(ux >= 0 && ux < p.fx && uy >= 0 && uy < p.fy)) continue;
        const size_t r = (size_t)uy * p.fx + ux;
        const float n_w = edge_weight(nval, load3(p.normal, r), p.n_phi, true);
        const float p_w = edge_weight(pval, load3(p.pos, r), p.p_phi, true);
        const float k = eaw_kernel(i);
#pragma unroll
        for (int m = 0; m < NI; ++m) {
            const float3 ctmp = load3(p.color[m], r);
            const float c_w = edge_weight(cval[m], ctmp, p.c_phi, false);
            const float weight = c_w * n_w * p_w;
            sum[m] += ctmp * weight * k;
            cum[m] += weight * k;
        }
    }
#pragma unroll
    for (int m = 0; m < NI; ++m) {
        store3(p.out_color[m], q, sum[m] / cum[m]);
        if (p.cum_w[m]) p.cum_w[m][q] = cum[m];
    }
}

// Backward of the batched level.  Gradients carry a tolerance, not bit-exactness (include/mirres_b200.h), and the first
// version of this kernel was instruction-bound on IEEE divisions and the contract exp (ncu: 20 k instructions per
// foreground pixel, profiles/r1r_*), so the arithmetic here is the cheap kind: reciprocals of phi / of the footprint
// normalisations are formed once and multiplied, exp is the hardware ex2 path, the three edge-stopping scale factors of
// a tap are summed before they are applied, and the normal / position gradients of all images are accumulated into ONE
// pair of outputs (autograd would add them anyway).  Relative deviation from the exact-arithmetic version: < 1e-5.
MR_DEV float eaw_fast_exp(float x)
{
#if defined(__CUDA_ARCH__)
    return __expf(x);
#else
    return expf(x);
#endif
}
MR_DEV float eaw_fast_rcp(float x)
{
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}
template <int NI>
MR_DEV void eaw_bwd_multi_px(const EawMultiParams &p, int idx)
{
    const size_t q = (size_t)idx;
    const int px = idx % p.fx, py = idx / p.fx;
    const float3 nq = load3(p.normal, q), pq = load3(p.pos, q);
    const bool q_center = !(MR_LDG(p.occ + q) < 0.1f);
    const float nic = -eaw_fast_rcp(p.c_phi), nin = -eaw_fast_rcp(p.n_phi), nip = -eaw_fast_rcp(p.p_phi);
    float3 cq[NI], gqw[NI], oq[NI], gc[NI];
    float3 gn = f3(0.f), gp = f3(0.f);
#pragma unroll
    for (int m = 0; m < NI; ++m) {
        cq[m] = load3(p.color[m], q);
        oq[m] = load3(p.out_color[m], q);
        // upstream gradient of q's own footprint, already divided by its normalisation
        const float Wq = p.cum_w[m][q];
        gqw[m] = q_center ? load3(p.g_out[m], q) * eaw_fast_rcp(Wq) : f3(0.f);
        gc[m] = f3(0.f);
    }
#pragma unroll 1
    for (int iy = 0; iy < 5; ++iy) {
        const int uy = py + (iy - 2) * p.step;
        if (uy < 0 || uy >= p.fy) continue;
        const float ky = (float)((0x14641 >> (4 * iy)) & 15) * (1.0f / 256.0f);
#pragma unroll 1
        for (int ix = 0; ix < 5; ++ix) {
            const int ux = px + (ix - 2) * p.step;
            if (ux < 0 || ux >= p.fx) continue;
            const size_t r = (size_t)uy * p.fx + ux;
            const bool r_center = !(MR_LDG(p.occ + r) < 0.1f);
            if (!q_center && !r_center) continue; // neither footprint exists: nothing flows between q and r
            const float k = ky * (float)((0x14641 >> (4 * ix)) & 15);
            const float3 dn = nq - load3(p.normal, r), dp = pq - load3(p.pos, r);
            const float wk = fminf(eaw_fast_exp(fmaxf(dot(dn, dn), 0.f) * nin), 1.0f) * fminf(eaw_fast_exp(fmaxf(dot(dp, dp), 0.f) * nip), 1.0f) * k;
            float s_all = 0.f;
#pragma unroll
            for (int m = 0; m < NI; ++m) {
                const float3 cr = load3(p.color[m], r);
                const float3 dc = cq[m] - cr;
                const float w = fminf(eaw_fast_exp(dot(dc, dc) * nic), 1.0f) * wk; // weight x kernel of this tap
                // q as the centre, r as its tap:  d out_q / d w = k (c_r - out_q) / W_q
                float gw = dot(gqw[m], cr - oq[m]);
                if (r_center) {
                    // r as the centre, q as its tap (kernel and weight are symmetric)
                    const float3 grw = load3(p.g_out[m], r) * eaw_fast_rcp(p.cum_w[m][r]);
                    gc[m] += grw * w;
                    gw += dot(grw, cq[m] - load3(p.out_color[m], r));
                }
                // dw / d|t|^2 = -w / phi,  d|t|^2 / d t = 2 t
                const float s = gw * w * 2.0f;
                gc[m] += dc * (s * nic);
                s_all += s;
            }
            gn += dn * (s_all * nin);
            gp += dp * (s_all * nip);
        }
    }
#pragma unroll
    for (int m = 0; m < NI; ++m) store3(p.g_color[m], q, gc[m]);
    if (p.g_normal[0]) store3(p.g_normal[0], q, gn);
    if (p.g_pos[0]) store3(p.g_pos[0], q, gp);
}

template <int NI>
static int eaw_multi_launch(const EawMultiParams &p, bool backward, cudaStream_t st)
{
    if (backward) return foreach_item<EawMultiParams, eaw_bwd_multi_px<NI>, 128>(p, p.fx * p.fy, st);
    return foreach_item<EawMultiParams, eaw_fwd_multi_px<NI>, 128>(p, p.fx * p.fy, st);
}
static int eaw_multi_dispatch(const EawMultiParams &p, bool backward, cudaStream_t st)
{
    switch (p.n_images) {
    case 1: return eaw_multi_launch<1>(p, backward, st);
    case 2: return eaw_multi_launch<2>(p, backward, st);
    case 3: return eaw_multi_launch<3>(p, backward, st);
    case 4: return eaw_multi_launch<4>(p, backward, st);
    case 5: return eaw_multi_launch<5>(p, backward, st);
    case 6: return eaw_multi_launch<6>(p, backward, st);
    case 7: return eaw_multi_launch<7>(p, backward, st);
    default: return eaw_multi_launch<8>(p, backward, st);
    }
}

// ---- cross-bilateral denoiser (--use_bi_de) ---------------------------------------------------------------
// nerf/renderutils/c_src/denoising.cu:14-130 (nvdiffrec's bilateral_denoiser_fwd/bwd_kernel), called from
// nerf/renderer_restir.py:529-541 through nerf/renderutils/ops.py:173-212.  Same tap order and arithmetic; expf and
// powf(., 128) are the contract functions mr_expf / mr_pow128f.  Both directions are gathers, as in the reference.
struct BilateralParams {
    int fx, fy, rad;
    float variance;
    const float *__restrict__ col;      // [N,3] forward only
    const float *__restrict__ nrm;      // [N,3]
    const float *__restrict__ zdz;      // [N,2]
    const float *__restrict__ out_grad; // [N,4] backward only
    float *__restrict__ out;            // forward: [N,4]; backward: col_grad [N,3]
};
template <bool BACKWARD>
MR_DEV void bilateral_px(const BilateralParams &p, int idx)
{
    const float FLT_EPS = 0.0001f;
    const int px = idx % p.fx, py = idx / p.fx;
    const float3 c_nrm = load3(p.nrm, (size_t)idx);
    const float c_z = MR_LDG(p.zdz + 2 * (size_t)idx), c_dz = MR_LDG(p.zdz + 2 * (size_t)idx + 1);
    float accum_w = 0.0f;
    float3 accum = f3(0.f);
    for (int oy = -p.rad; oy <= p.rad; ++oy) {
        const int y = py + oy;
        if (y < 0 || y >= p.fy) continue;
        for (int ox = -p.rad; ox <= p.rad; ++ox) {
            const int x = px + ox;
            if (x < 0 || x >= p.fx) continue;
            const size_t t = (size_t)y * p.fx + x;
            const float dist_sqr = (float)(ox * ox + oy * oy);
            const float dist = sqrtf(dist_sqr);
            const float w_xy = mr_expf(-dist_sqr / (2.0f * p.variance));
            const float w_normal = mr_pow128f(fminf(fmaxf(dot(load3(p.nrm, t), c_nrm), FLT_EPS), 1.0f));
            const float dz = BACKWARD ? MR_LDG(p.zdz + 2 * t + 1) : c_dz; // the transposed gather uses the tap's gradient
            const float w_depth = mr_expf(-(fabsf(MR_LDG(p.zdz + 2 * t) - c_z) / fmaxf(dz * dist, FLT_EPS)));
            const float w = w_xy * w_normal * w_depth;
            if (BACKWARD) {
                accum += make_float3(MR_LDG(p.out_grad + 4 * t), MR_LDG(p.out_grad + 4 * t + 1), MR_LDG(p.out_grad + 4 * t + 2)) * w;
            } else {
                accum = accum + load3(p.col, t) * w;
                accum_w += w;
            }
        }
    }
    if (BACKWARD) {
        store3(p.out, (size_t)idx, accum);
    } else {
        p.out[4 * (size_t)idx] = accum.x; p.out[4 * (size_t)idx + 1] = accum.y; p.out[4 * (size_t)idx + 2] = accum.z;
        p.out[4 * (size_t)idx + 3] = fmaxf(accum_w, 0.0001f);
    }
}

struct AoParams {
    int fx, fy;
    const float *__restrict__ occ;
    const float *__restrict__ normal;
    float *__restrict__ out_ao;
};
MR_DEV void normal_ao_px(const AoParams &p, int idx)
{
    const size_t q = (size_t)idx;
    if (MR_LDG(p.occ + q) < 0.1f) { store3(p.out_ao, q, f3(0.f)); return; }
    const int px = idx % p.fx, py = idx / p.fx;
    const float3 nval = load3(p.normal, q);
    int count = 0;
    float sum = 0.f;
    const int width = 4;
    for (int i = -width; i < width; i++)
        for (int j = -width; j < width; j++) {
            const int ux = px + i, uy = py + j;
            if (!(ux >= 0 && ux < p.fx && uy >= 0 && uy < p.fy)) continue;
            const size_t r = (size_t)uy * p.fx + ux;
            if (MR_LDG(p.occ + r) < 0.1f) continue;
            float d = fmaxf(dot(load3(p.normal, r), nval), 0.0f);
            d = fminf(1.0f, d);
            sum += d;
            count++;
        }
    float normal_weight = 1 - sum / (float)count;
    float v = clampf(normal_weight * 50, 0.f, 1.f);
    store3(p.out_ao, q, f3(v));
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_eaw_fwd(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                   const float *color, const float *normal, const float *pos, float *out_color, void *stream)
{
    if (!occ || !color || !normal || !pos || !out_color) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1) return MIRRES_ERR_SHAPE;
    if (out_color == color) return MIRRES_ERR_ALIAS;
    EawParams p = {};
    p.c_phi = c_phi; p.n_phi = n_phi; p.p_phi = p_phi; p.fx = fx; p.fy = fy; p.step = (int)step_width;
    p.occ = occ; p.color = color; p.normal = normal; p.pos = pos; p.out_color = out_color;
    return foreach_item<EawParams, eaw_fwd_px, 128>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_eaw_bwd(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                   const float *color, const float *normal, const float *pos, const float *out_color,
                   const float *grad_out, float *grad_color, float *grad_normal, float *grad_pos, float *cum_w_scratch,
                   void *stream)
{
    if (!occ || !color || !normal || !pos || !out_color || !grad_out || !grad_color || !grad_normal || !grad_pos || !cum_w_scratch)
        return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1) return MIRRES_ERR_SHAPE;
    EawParams p = {};
    p.c_phi = c_phi; p.n_phi = n_phi; p.p_phi = p_phi; p.fx = fx; p.fy = fy; p.step = (int)step_width;
    p.occ = occ; p.color = color; p.normal = normal; p.pos = pos; p.out_color = (float *)out_color;
    p.cum_w = cum_w_scratch; p.g_out = grad_out; p.g_color = grad_color; p.g_normal = grad_normal; p.g_pos = grad_pos;
    int rc = foreach_item<EawParams, eaw_norm_px, 128>(p, fx * fy, (cudaStream_t)stream);
    if (rc) return rc;
    return foreach_item<EawParams, eaw_bwd_px, 128>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_eaw_fwd_multi(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                         const float *normal, const float *pos, int n_images, const float *const *colors,
                         float *const *out_colors, float *const *cum_w, void *stream)
{
    if (!occ || !normal || !pos || !colors || !out_colors) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || n_images < 1 || n_images > MR_EAW_MAX_IMAGES) return MIRRES_ERR_SHAPE;
    EawMultiParams p = {};
    p.c_phi = c_phi; p.n_phi = n_phi; p.p_phi = p_phi; p.fx = fx; p.fy = fy; p.step = (int)step_width; p.n_images = n_images;
    p.occ = occ; p.normal = normal; p.pos = pos;
    for (int m = 0; m < n_images; ++m) {
        if (!colors[m] || !out_colors[m]) return MIRRES_ERR_NULL;
        for (int j = 0; j < n_images; ++j)
            if (out_colors[m] == colors[j]) return MIRRES_ERR_ALIAS;
        p.color[m] = colors[m];
        p.out_color[m] = out_colors[m];
        p.cum_w[m] = cum_w ? cum_w[m] : nullptr;
    }
    return eaw_multi_dispatch(p, false, (cudaStream_t)stream);
}

int mirres_eaw_bwd_multi(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                         const float *normal, const float *pos, int n_images, const float *const *colors,
                         const float *const *out_colors, const float *const *cum_w, const float *const *grad_outs,
                         float *const *grad_colors, float *grad_normal_sum, float *grad_pos_sum, void *stream)
{
    if (!occ || !normal || !pos || !colors || !out_colors || !cum_w || !grad_outs || !grad_colors) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || n_images < 1 || n_images > MR_EAW_MAX_IMAGES) return MIRRES_ERR_SHAPE;
    EawMultiParams p = {};
    p.c_phi = c_phi; p.n_phi = n_phi; p.p_phi = p_phi; p.fx = fx; p.fy = fy; p.step = (int)step_width; p.n_images = n_images;
    p.occ = occ; p.normal = normal; p.pos = pos;
    for (int m = 0; m < n_images; ++m) {
        if (!colors[m] || !out_colors[m] || !cum_w[m] || !grad_outs[m] || !grad_colors[m]) return MIRRES_ERR_NULL;
        p.color[m] = colors[m];
        p.out_color[m] = (float *)out_colors[m];
        p.cum_w[m] = (float *)cum_w[m];
        p.g_out[m] = grad_outs[m];
        p.g_color[m] = grad_colors[m];
    }
    p.g_normal[0] = grad_normal_sum;
    p.g_pos[0] = grad_pos_sum;
    return eaw_multi_dispatch(p, true, (cudaStream_t)stream);
}

int mirres_bilateral_fwd(int fx, int fy, float sigma, const float *col, const float *nrm, const float *zdz, float *out,
                         void *stream)
{
    if (!col || !nrm || !zdz || !out) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || !(sigma > 0.f)) return MIRRES_ERR_SHAPE;
    BilateralParams p = {fx, fy, 2 * (int)ceilf(sigma * 2.5f) + 1, sigma * sigma, col, nrm, zdz, nullptr, out};
    return foreach_item<BilateralParams, bilateral_px<false>, 128>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_bilateral_bwd(int fx, int fy, float sigma, const float *nrm, const float *zdz, const float *out_grad,
                         float *col_grad, void *stream)
{
    if (!nrm || !zdz || !out_grad || !col_grad) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1 || !(sigma > 0.f)) return MIRRES_ERR_SHAPE;
    BilateralParams p = {fx, fy, 2 * (int)ceilf(sigma * 2.5f) + 1, sigma * sigma, nullptr, nrm, zdz, out_grad, col_grad};
    return foreach_item<BilateralParams, bilateral_px<true>, 128>(p, fx * fy, (cudaStream_t)stream);
}

int mirres_normal_ao(int fx, int fy, const float *occ, const float *normal, float *out_ao, void *stream)
{
    if (!occ || !normal || !out_ao) return MIRRES_ERR_NULL;
    if (fx < 1 || fy < 1) return MIRRES_ERR_SHAPE;
    AoParams p = {fx, fy, occ, normal, out_ao};
    return foreach_item<AoParams, normal_ao_px, 128>(p, fx * fy, (cudaStream_t)stream);
}

} // extern "C"
