// mirres-b200: G-buffer producer and the gradient scatter that closes the backward pass.
//
// In the reference the G-buffer comes from nvdiffrast (rasterise + interpolate, nerf/renderer.py:979-1030) and the
// per-pixel gradients the path emits (grad_normal / grad_diffuse / grad_ks, nerf/ScreenSpaceReSTIR/Resampling.py:193-214)
// are scattered to mesh vertices and texels by nvdiffrast's / tiny-cuda-nn's own backward.  Here:
//   mirres_gbuffer_primary   one closest-hit ray per pixel through the same LBVH (bvh_hit_with_normal semantics,
//                            helperDi.slang:313-395): occupancy, position, normal, depth and -- what rasterisation
//                            gives the reference -- the triangle id and barycentrics of every pixel; with per-vertex
//                            normals the shading normal is their barycentric interpolation (the role of
//                            dr.interpolate over auto_normals, nerf/meshutils.py:14-39)
//   mirres_interpolate_bwd   reverse of barycentric interpolation: out[tri[prim][k], c] += w_k * grad[pixel, c].
//                            Neighbouring pixels see the same triangle, so lanes of a warp with the same triangle id
//                            are summed with a shuffle tree (__match_any_sync) and one lane issues the atomics.
#include "mr_wave.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

struct GbufParams {
    BvhView bvh;
    const float *__restrict__ org;     // [n,3]
    const float *__restrict__ dir;     // [n,3]
    const float *__restrict__ vnormal; // [V,3] or null
    const int *__restrict__ tri;       // [F,3] (only with vnormal)
    float *__restrict__ occ;           // [n]
    float *__restrict__ pos;           // [n,3]
    float *__restrict__ normal;        // [n,3]
    float *__restrict__ depth;         // [n]
    int *__restrict__ prim;            // [n]
    float *__restrict__ bary;          // [n,2]
    float *__restrict__ gnormal;       // [n,3] or null: the face normal, kept when `normal` receives the interpolated one
};

// shared tail of both launch shapes
MR_DEV void gbuffer_write(const GbufParams &p, size_t i, bool found, float3 o, float3 x, float3 n, int prim, float u, float v)
{
    float d = 0.f;
    if (p.gnormal) store3(p.gnormal, i, found ? n : f3(0.f));
    if (found) {
        const float3 dv = x - o;
        d = sqrtf(dot(dv, dv));
        if (p.vnormal && prim >= 0) {
            const int i0 = MR_LDG(p.tri + 3 * (size_t)prim), i1 = MR_LDG(p.tri + 3 * (size_t)prim + 1), i2 = MR_LDG(p.tri + 3 * (size_t)prim + 2);
            const float w = 1.0f - u - v;
            n = w * load3(p.vnormal, (size_t)i0) + u * load3(p.vnormal, (size_t)i1) + v * load3(p.vnormal, (size_t)i2);
        }
    } else {
        x = f3(0.f);
        n = f3(0.f);
    }
    p.occ[i] = found ? 1.0f : 0.0f;
    store3(p.pos, i, x);
    store3(p.normal, i, n);
    p.depth[i] = d;
    if (p.prim) p.prim[i] = found ? prim : -1;
    if (p.bary) { p.bary[2 * i] = found ? u : 0.f; p.bary[2 * i + 1] = found ? v : 0.f; }
}

// wavefront shape: every pixel queues its primary ray (slot = pixel), the persistent closest-hit tracer walks the queue,
// the resolve kernel assembles the maps
struct GbufWaveParams {
    GbufParams g;
    Workspace ws;
    int n;
};
// Most primary rays of a view miss the mesh altogether.  The gen pass takes the walk's own first step for them -- the
// four boxes of the root record against closest = 1e7, the test closest_hit starts with -- and hands only the rays that
// pass one of them to the tracer; the others are the misses the tracer would have reported after the same test.
MR_DEV void gbuffer_gen_px(const GbufWaveParams &p, int idx)
{
    const float3 o = load3(p.g.org, (size_t)idx), d = load3(p.g.dir, (size_t)idx);
    const Ray r = make_ray(o, d);
    WideHit w;
    wide_fetch(r, p.g.bvh.nodes, 0, w);
    bool pass = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) pass = pass || fminf(1e7f, w.tf[k]) > w.tn[k];
    if (pass) queue_closest_ray(p.ws, (size_t)idx, o, d);
    else queue_closest_empty(p.ws, (size_t)idx);
}
MR_DEV void gbuffer_resolve_px(const GbufWaveParams &p, int idx)
{
    const size_t i = (size_t)idx;
    const float4 h0 = p.ws.chit[3 * i], h1 = p.ws.chit[3 * i + 1], h2 = p.ws.chit[3 * i + 2];
    gbuffer_write(p.g, i, h0.w > 0.5f, load3(p.g.org, i), make_float3(h0.x, h0.y, h0.z), make_float3(h1.x, h1.y, h1.z),
                  float_bits(h2.x), h2.y, h2.z);
}

MR_DEV void gbuffer_item(const GbufParams &p, int idx)
{
    const size_t i = (size_t)idx;
    Hit h;
    h.t = 0.f;
    h.pos = f3(0.f);
    h.normal = f3(1.f);
    h.prim = -1;
    h.bary[0] = h.bary[1] = 0.f;
    const float3 o = load3(p.org, i);
    const bool found = closest_hit<false>(p.bvh, o, load3(p.dir, i), h, nullptr);
    gbuffer_write(p, i, found, o, h.pos, h.normal, h.prim, h.bary[0], h.bary[1]);
}

// ---- derived per-pixel maps of run_restir_di_with_pt in one pass -----------------------------------------------------
// nerf/renderer_restir.py:279-287 (normal_depth, brdf_map) and :484-486 (occupancy threshold, ray normalisation) are
// ~25 elementwise torch launches in the reference; same operations, same order, one kernel.
struct PrepParams {
    float *occ;                        // [n] in place: occ <= 0.5 -> 0
    const float *__restrict__ normal;  // [n,3]
    const float *__restrict__ depth;   // [n]
    const float *__restrict__ kd;      // [n,3]
    const float *__restrict__ rs;      // [n,2] roughness, metallic
    const float *__restrict__ ray_in;  // [n,3]
    float *__restrict__ normal_depth;  // [n,4]
    float *__restrict__ brdf_map;      // [n,3] luminance(kd), metallic weight, alpha
    float *__restrict__ ray_out;       // [n,3]
};
MR_DEV float clamp_nan(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); } // torch.clamp
MR_DEV void prepare_maps_px(const PrepParams &p, int idx)
{
    const size_t i = (size_t)idx;
    if (p.occ[i] <= 0.5f) p.occ[i] = 0.f;
    const float3 nrm = load3(p.normal, i);
    reinterpret_cast<float4 *>(p.normal_depth)[i] = make_float4(nrm.x, nrm.y, nrm.z, MR_LDG(p.depth + i));
    const float3 kd = load3(p.kd, i);
    const float r = MR_LDG(p.rs + 2 * i), m = MR_LDG(p.rs + 2 * i + 1);
    const float lum = kd.x * 0.2126f + kd.y * 0.7152f + kd.z * 0.0722f;
    const float met = m * 0.2126f + m * 0.7152f + m * 0.0722f;
    const float a = clamp_nan(r, 0.01f, 1.0f);
    store3(p.brdf_map, i, make_float3(lum, met, a * a));
    const float3 d = load3(p.ray_in, i);
    const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    const float den = len != len ? len : fmaxf(len, 1e-6f);
    store3(p.ray_out, i, make_float3(d.x / den, d.y / den, d.z / den));
}

// ---- shading-normal set-up (the prepare_shading_normal step of the G-buffer stage) ----------------------------------
// Reference: PrepareShadingNormalFwdKernel / BwdKernel, nerf/renderutils/c_src/normal.cu:95-178, called at
// nerf/renderer.py:1013 between rasterisation and run_restir_di_with_pt.  Per pixel:
//   N  = nrm(smooth_nrm), T = nrm(smooth_tng), V = nrm(view_pos - pos), B = nrm(T x N)          nrm(0) = 0
//   S  = nrm(T p.x + s B p.y + N max(p.z, 0))      p = perturbed_nrm, s = -1 (OpenGL) / +1
//   two-sided: dot(V, G) < 0 flips S and G
//   out = G (1 - t) + S t,  t = clamp(dot(V, S) / 0.1, 0, 1)
// The backward below is a hand-derived reverse sweep of exactly that expression (branch predicates frozen).
struct ShadingNormalParams {
    const float *in[6];  // pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm
    int rs[6];           // row strides in floats (0 = one row broadcast to every pixel)
    int two_sided, opengl;
    float *out;          // fwd: [n,3]
    const float *gout;   // bwd: [n,3]
    float *gin[6];       // bwd: [n,3] each, optional
};
MR_DEV float3 sn_load(const ShadingNormalParams &p, int k, size_t i)
{
    const float *q = p.in[k] + i * (size_t)p.rs[k];
    return make_float3(MR_LDG(q), MR_LDG(q + 1), MR_LDG(q + 2));
}
MR_DEV float3 nrm0(float3 v)
{
    const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return l > 0.0f ? v / l : f3(0.f);
}
// reverse of y = v / |v|:  dv = (dy - y (y . dy)) / |v|   (0 where |v| = 0, as the forward is constant there)
MR_DEV float3 nrm0_bwd(float3 v, float3 dy)
{
    const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (!(l > 0.0f)) return f3(0.f);
    const float3 y = v / l;
    return (dy - y * dot(y, dy)) / l;
}
struct ShadingNormalFrame {
    float3 N, T, V, Bu, B, Su, S, G;  // Bu / Su: before normalisation
    float sgn, dp, t;
    bool flip;
};
MR_DEV ShadingNormalFrame shading_normal_frame(const ShadingNormalParams &p, size_t i, float3 &pert, float3 &view_raw,
                                               float3 &n_raw, float3 &t_raw)
{
    ShadingNormalFrame f;
    const float3 pos = sn_load(p, 0, i), vp = sn_load(p, 1, i);
    pert = sn_load(p, 2, i);
    n_raw = sn_load(p, 3, i);
    t_raw = sn_load(p, 4, i);
    const float3 g = sn_load(p, 5, i);
    view_raw = vp - pos;
    f.N = nrm0(n_raw);
    f.T = nrm0(t_raw);
    f.V = nrm0(view_raw);
    f.Bu = cross(f.T, f.N);
    f.B = nrm0(f.Bu);
    f.sgn = p.opengl ? -1.0f : 1.0f;
    f.Su = f.T * pert.x + (f.sgn * f.B) * pert.y + f.N * fmaxf(pert.z, 0.0f);
    const float3 s = nrm0(f.Su);
    f.flip = p.two_sided && dot(f.V, g) < 0.0f;
    f.S = f.flip ? -s : s;
    f.G = f.flip ? -g : g;
    f.dp = dot(f.V, f.S);
    f.t = clampf(f.dp / 0.1f, 0.0f, 1.0f);
    return f;
}
MR_DEV void shading_normal_fwd_px(const ShadingNormalParams &p, int idx)
{
    float3 pert, vr, nr, tr;
    const ShadingNormalFrame f = shading_normal_frame(p, (size_t)idx, pert, vr, nr, tr);
    store3(p.out, (size_t)idx, f.G * (1.0f - f.t) + f.S * f.t);
}
MR_DEV void shading_normal_bwd_px(const ShadingNormalParams &p, int idx)
{
    const size_t i = (size_t)idx;
    float3 pert, vr, nr, tr;
    const ShadingNormalFrame f = shading_normal_frame(p, i, pert, vr, nr, tr);
    const float3 go = load3(p.gout, i);
    // bend
    float3 dS, dG = f3(0.f), dV = f3(0.f);
    if (f.dp > 0.1f) {
        dS = go;
    } else {
        dG = go * (1.0f - f.t);
        dS = go * f.t;
        const float dt = dot(go, f.S - f.G);
        const float ddp = (f.dp < 0.0f || f.dp > 0.1f) ? 0.0f : dt / 0.1f;
        dV = f.S * ddp;
        dS += f.V * ddp;
    }
    if (f.flip) { dS = -dS; dG = -dG; }
    // perturb
    const float3 dSu = nrm0_bwd(f.Su, dS);
    float3 dN = f3(0.f), dT, dpert = f3(0.f);
    if (pert.z > 0.0f) {
        dN = dSu * pert.z;
        dpert.z = dot(dSu, f.N);
    }
    const float3 dB = (f.sgn * dSu) * pert.y;
    dpert.y = f.sgn * dot(dSu, f.B);
    dT = dSu * pert.x;
    dpert.x = dot(dSu, f.T);
    const float3 dBu = nrm0_bwd(f.Bu, dB);
    dT += cross(f.N, dBu);   // Bu = T x N
    dN += cross(dBu, f.T);
    const float3 dview = nrm0_bwd(vr, dV);
    if (p.gin[0]) store3(p.gin[0], i, -dview);
    if (p.gin[1]) store3(p.gin[1], i, dview);
    if (p.gin[2]) store3(p.gin[2], i, dpert);
    if (p.gin[3]) store3(p.gin[3], i, nrm0_bwd(nr, dN));
    if (p.gin[4]) store3(p.gin[4], i, nrm0_bwd(tr, dT));
    if (p.gin[5]) store3(p.gin[5], i, dG);
}

#define MR_SCATTER_MAX_C 8

struct ScatterParams {
    const float *__restrict__ grad; // [n,C]
    const int *__restrict__ prim;   // [n], -1 = background
    const float *__restrict__ bary; // [n,2] or null (null: 1/3 each)
    const int *__restrict__ tri;    // [F,3]
    float *out;                     // [V,C]
    int n, C, F;
};

// per-pixel part shared by both builds
MR_DEV bool scatter_terms(const ScatterParams &p, int idx, int &prim, float w[3], float g[MR_SCATTER_MAX_C])
{
    prim = MR_LDG(p.prim + idx);
    if (prim < 0 || prim >= p.F) return false;
    bool nz = false;
    for (int c = 0; c < p.C; ++c) {
        g[c] = MR_LDG(p.grad + (size_t)idx * p.C + c);
        nz = nz || g[c] != 0.f;
    }
    if (!nz) return false;
    if (p.bary) {
        const float u = MR_LDG(p.bary + 2 * (size_t)idx), v = MR_LDG(p.bary + 2 * (size_t)idx + 1);
        w[0] = 1.0f - u - v; w[1] = u; w[2] = v;
    } else {
        w[0] = w[1] = w[2] = 1.0f / 3.0f;
    }
    return true;
}

#if !defined(MR_HOST_CHECK)
template <int C>
__global__ void __launch_bounds__(256) k_interpolate_bwd(ScatterParams p)
{
    const unsigned int FULL = 0xffffffffu;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    int prim = -1;
    float w[3] = {0.f, 0.f, 0.f}, g[MR_SCATTER_MAX_C];
#pragma unroll
    for (int c = 0; c < MR_SCATTER_MAX_C; ++c) g[c] = 0.f;
    const bool active = idx < p.n && scatter_terms(p, idx, prim, w, g);
    const unsigned int act = __ballot_sync(FULL, active);
    if (act == 0u) return;
    // lanes that see the same triangle form a group; its first lane is the leader
    const unsigned int peers = __match_any_sync(FULL, active ? prim : -1 - (int)lane) & act;
    const bool leader = active && (__ffs(peers) - 1 == (int)lane);
    const unsigned int leaders = __ballot_sync(FULL, leader);
    int v[3] = {0, 0, 0};
    if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = __ldg(p.tri + 3 * (size_t)prim + k);
    }
    if (__popc(leaders) == __popc(act)) {
        // every lane has its own triangle: nothing to aggregate
        if (active) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) atomicAdd(p.out + (size_t)v[k] * C + c, w[k] * g[c]);
        }
        return;
    }
    // sum inside each group by walking the group mask
    float acc[3][C];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[k][c] = active ? w[k] * g[c] : 0.f;
    // every lane pulls the contributions of the other members of ITS group; after the loop the leader holds the
    // group total (members hold the same total, which they discard)
    unsigned int rest = peers & ~(1u << lane);
    const int rounds = (int)__reduce_max_sync(FULL, (unsigned int)__popc(peers)) - 1; // largest group of this warp
    float mine[3][C];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) mine[k][c] = acc[k][c];
    for (int r = 0; r < rounds; ++r) {
        const int src = rest ? __ffs(rest) - 1 : (int)lane;
        const bool take = rest != 0u;
        rest &= rest - 1u;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float x = __shfl_sync(FULL, mine[k][c], src);
                if (take) acc[k][c] += x;
            }
    }
    if (leader) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) atomicAdd(p.out + (size_t)v[k] * C + c, acc[k][c]);
    }
}
#endif

// ---- vertex normals (auto_normals, meshutils.py:14-39; called at nerf/renderer.py:979-1030 on the offset vertices) ----
//   face normal  n_f = (v1 - v0) x (v2 - v0)           (area-weighted, not normalised)
//   vertex sum   s_v = sum of n_f over incident faces   (scatter_add_ in the reference: float atomics, any order)
//   vertex normal     = s_v / |s_v|  if  s_v . s_v > 1e-20,  else (0, 0, 1)
// Backward: through the normalisation (no gradient where the fallback was taken), gathered per face, through the cross
// product, scattered to the three corner positions.
struct VertexNormalParams {
    const float *__restrict__ vert; // [V,3]
    const int *__restrict__ tri;    // [F,3]
    float *vsum;                    // [V,3] un-normalised sums (written by fwd, read by bwd)
    float *vnrm;                    // [V,3]
    const float *__restrict__ gnrm; // [V,3] upstream gradient (bwd)
    float *gvert;                   // [V,3] accumulated into (bwd)
    int V, F;
};
MR_DEV bool vn_face(const VertexNormalParams &p, int f, int i[3], float3 &a, float3 &b)
{
    i[0] = MR_LDG(p.tri + 3 * (size_t)f), i[1] = MR_LDG(p.tri + 3 * (size_t)f + 1), i[2] = MR_LDG(p.tri + 3 * (size_t)f + 2);
    if ((unsigned)i[0] >= (unsigned)p.V || (unsigned)i[1] >= (unsigned)p.V || (unsigned)i[2] >= (unsigned)p.V) return false;
    const float3 v0 = load3(p.vert, (size_t)i[0]);
    a = load3(p.vert, (size_t)i[1]) - v0;
    b = load3(p.vert, (size_t)i[2]) - v0;
    return true;
}
MR_DEV void vn_add3(float *dst, size_t row, float3 v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(dst + 3 * row, v.x), atomicAdd(dst + 3 * row + 1, v.y), atomicAdd(dst + 3 * row + 2, v.z);
#else
    dst[3 * row] += v.x, dst[3 * row + 1] += v.y, dst[3 * row + 2] += v.z;
#endif
}
MR_DEV void vertex_normal_face(const VertexNormalParams &p, int f)
{
    int i[3];
    float3 a, b;
    if (!vn_face(p, f, i, a, b)) return;
    const float3 n = cross(a, b);
    for (int k = 0; k < 3; ++k) vn_add3(p.vsum, (size_t)i[k], n);
}
MR_DEV void vertex_normal_finish(const VertexNormalParams &p, int v)
{
    float3 s = load3_rw(p.vsum, (size_t)v);
    float d = dot(s, s);
    if (!(d > 1e-20f)) { s = f3(0.f, 0.f, 1.f); d = 1.0f; }
    store3(p.vnrm, (size_t)v, s / sqrtf(fmaxf(d, 1e-20f)));
}
MR_DEV void vertex_normal_bwd_face(const VertexNormalParams &p, int f)
{
    int i[3];
    float3 a, b;
    if (!vn_face(p, f, i, a, b)) return;
    float3 dn = f3(0.f);
    for (int k = 0; k < 3; ++k) {
        const float3 s = load3(p.vsum, (size_t)i[k]);
        const float d = dot(s, s);
        if (!(d > 1e-20f)) continue;
        const float l = sqrtf(d);
        const float3 y = s / l, g = load3(p.gnrm, (size_t)i[k]);
        dn += (g - y * dot(y, g)) / l;
    }
    if (is_black(dn)) return;
    const float3 da = cross(b, dn), db = cross(dn, a);  // n = a x b
    vn_add3(p.gvert, (size_t)i[1], da);
    vn_add3(p.gvert, (size_t)i[2], db);
    vn_add3(p.gvert, (size_t)i[0], -(da + db));
}
#if defined(MR_HOST_CHECK)
template <void (*BODY)(const VertexNormalParams &, int)>
static int vn_foreach(const VertexNormalParams &p, int n, cudaStream_t)
{
    for (int i = 0; i < n; ++i) BODY(p, i);  // sequential: the bodies accumulate without atomics on the host
    return 0;
}
#else
template <void (*BODY)(const VertexNormalParams &, int)>
static int vn_foreach(const VertexNormalParams &p, int n, cudaStream_t st)
{
    return foreach_item<VertexNormalParams, BODY, 256>(p, n, st);
}
#endif

static int interpolate_bwd_launch(const ScatterParams &p, cudaStream_t st)
{
#if defined(MR_HOST_CHECK)
    (void)st;
    for (int idx = 0; idx < p.n; ++idx) {
        int prim;
        float w[3], g[MR_SCATTER_MAX_C];
        if (!scatter_terms(p, idx, prim, w, g)) continue;
        for (int k = 0; k < 3; ++k) {
            const int v = p.tri[3 * (size_t)prim + k];
            for (int c = 0; c < p.C; ++c) p.out[(size_t)v * p.C + c] += w[k] * g[c];
        }
    }
    return 0;
#else
    const int grid = (p.n + 255) / 256;
    switch (p.C) {
    case 1: k_interpolate_bwd<1><<<grid, 256, 0, st>>>(p); break;
    case 2: k_interpolate_bwd<2><<<grid, 256, 0, st>>>(p); break;
    case 3: k_interpolate_bwd<3><<<grid, 256, 0, st>>>(p); break;
    case 4: k_interpolate_bwd<4><<<grid, 256, 0, st>>>(p); break;
    case 5: k_interpolate_bwd<5><<<grid, 256, 0, st>>>(p); break;
    case 6: k_interpolate_bwd<6><<<grid, 256, 0, st>>>(p); break;
    case 7: k_interpolate_bwd<7><<<grid, 256, 0, st>>>(p); break;
    default: k_interpolate_bwd<8><<<grid, 256, 0, st>>>(p); break;
    }
    MR_CUDA_CHECK_LAUNCH();
    return 0;
#endif
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_gbuffer_primary(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                           const float *vnormal, const int *tri, float *occ, float *pos, float *normal, float *depth,
                           int *prim, float *bary, float *geom_normal, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!packed_nodes || !packed_tris || !org || !dir || !occ || !pos || !normal || !depth) return MIRRES_ERR_NULL;
    if (vnormal && !tri) return MIRRES_ERR_NULL;
    if (n < 0) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    GbufParams p = {bvh_view(packed_nodes, packed_tris), org, dir, vnormal, tri, occ, pos, normal, depth, prim, bary, geom_normal};
    cudaStream_t st = (cudaStream_t)stream;
    if (!workspace) return foreach_item<GbufParams, gbuffer_item, 128>(p, n, st);
    if ((uintptr_t)workspace & 255) return MIRRES_ERR_ALIGN;
    if (workspace_bytes < workspace_carve(nullptr, n, nullptr)) return MIRRES_ERR_SCRATCH;
    GbufWaveParams w;
    w.g = p;
    w.n = n;
    workspace_carve(&w.ws, n, (char *)workspace);
    queue_reset(w.ws, st);
    int rc;
    if ((rc = foreach_item<GbufWaveParams, gbuffer_gen_px, 256>(w, n, st))) return rc;
    if ((rc = trace_queues(p.bvh, w.ws, false, true, device_sm_count(), st))) return rc;
    return foreach_item<GbufWaveParams, gbuffer_resolve_px, 256>(w, n, st);
}

int mirres_prepare_maps(int n, float *occ, const float *normal, const float *depth, const float *diffuse_map,
                        const float *rough_metal, const float *ray_dir, float *normal_depth, float *brdf_map,
                        float *ray_dir_normalized, void *stream)
{
    if (!occ || !normal || !depth || !diffuse_map || !rough_metal || !ray_dir || !normal_depth || !brdf_map || !ray_dir_normalized)
        return MIRRES_ERR_NULL;
    if (n < 1) return MIRRES_ERR_SHAPE;
    if ((uintptr_t)normal_depth & 15) return MIRRES_ERR_ALIGN;
    PrepParams p = {occ, normal, depth, diffuse_map, rough_metal, ray_dir, normal_depth, brdf_map, ray_dir_normalized};
    return foreach_item<PrepParams, prepare_maps_px, 256>(p, n, (cudaStream_t)stream);
}

int mirres_vertex_normals_fwd(const float *vert, int V, const int *tri, int F, float *vsum, float *vnrm, void *stream)
{
    if (!vert || !tri || !vsum || !vnrm) return MIRRES_ERR_NULL;
    if (V < 0 || F < 0) return MIRRES_ERR_SHAPE;
    if (V == 0) return 0;
    VertexNormalParams p = {vert, tri, vsum, vnrm, nullptr, nullptr, V, F};
    zero_async(vsum, (size_t)V * 12, (cudaStream_t)stream);
    if (F > 0)
        if (int rc = vn_foreach<vertex_normal_face>(p, F, (cudaStream_t)stream)) return rc;
    return vn_foreach<vertex_normal_finish>(p, V, (cudaStream_t)stream);
}

int mirres_vertex_normals_bwd(const float *vert, int V, const int *tri, int F, const float *vsum, const float *grad_vnrm,
                              float *grad_vert, void *stream)
{
    if (!vert || !tri || !vsum || !grad_vnrm || !grad_vert) return MIRRES_ERR_NULL;
    if (V < 0 || F < 0) return MIRRES_ERR_SHAPE;
    if (V == 0 || F == 0) return 0;
    VertexNormalParams p = {vert, tri, const_cast<float *>(vsum), nullptr, grad_vnrm, grad_vert, V, F};
    return vn_foreach<vertex_normal_bwd_face>(p, F, (cudaStream_t)stream);
}

static int shading_normal_params(mr::ShadingNormalParams &p, int n, const float *const in[6], const int rs[6], int two_sided, int opengl)
{
    if (n < 1) return MIRRES_ERR_SHAPE;
    for (int k = 0; k < 6; ++k) {
        if (!in[k]) return MIRRES_ERR_NULL;
        if (rs[k] < 0) return MIRRES_ERR_SHAPE;
        p.in[k] = in[k];
        p.rs[k] = rs[k];
        p.gin[k] = nullptr;
    }
    p.two_sided = two_sided != 0;
    p.opengl = opengl != 0;
    p.out = nullptr;
    p.gout = nullptr;
    return 0;
}

int mirres_shading_normal_fwd(int n, const float *pos, int pos_rs, const float *view_pos, int view_rs,
                              const float *perturbed_nrm, int perturbed_rs, const float *smooth_nrm, int smooth_nrm_rs,
                              const float *smooth_tng, int smooth_tng_rs, const float *geom_nrm, int geom_rs, int two_sided,
                              int opengl, float *out, void *stream)
{
    if (n == 0) return 0;
    if (!out) return MIRRES_ERR_NULL;
    const float *in[6] = {pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm};
    const int rs[6] = {pos_rs, view_rs, perturbed_rs, smooth_nrm_rs, smooth_tng_rs, geom_rs};
    ShadingNormalParams p;
    if (int rc = shading_normal_params(p, n, in, rs, two_sided, opengl)) return rc;
    p.out = out;
    return foreach_item<ShadingNormalParams, shading_normal_fwd_px, 256>(p, n, (cudaStream_t)stream);
}

int mirres_shading_normal_bwd(int n, const float *pos, int pos_rs, const float *view_pos, int view_rs,
                              const float *perturbed_nrm, int perturbed_rs, const float *smooth_nrm, int smooth_nrm_rs,
                              const float *smooth_tng, int smooth_tng_rs, const float *geom_nrm, int geom_rs, int two_sided,
                              int opengl, const float *grad_out, float *grad_pos, float *grad_view_pos,
                              float *grad_perturbed_nrm, float *grad_smooth_nrm, float *grad_smooth_tng, float *grad_geom_nrm,
                              void *stream)
{
    if (n == 0) return 0;
    if (!grad_out) return MIRRES_ERR_NULL;
    const float *in[6] = {pos, view_pos, perturbed_nrm, smooth_nrm, smooth_tng, geom_nrm};
    const int rs[6] = {pos_rs, view_rs, perturbed_rs, smooth_nrm_rs, smooth_tng_rs, geom_rs};
    ShadingNormalParams p;
    if (int rc = shading_normal_params(p, n, in, rs, two_sided, opengl)) return rc;
    p.gout = grad_out;
    float *g[6] = {grad_pos, grad_view_pos, grad_perturbed_nrm, grad_smooth_nrm, grad_smooth_tng, grad_geom_nrm};
    for (int k = 0; k < 6; ++k) p.gin[k] = g[k];
    return foreach_item<ShadingNormalParams, shading_normal_bwd_px, 256>(p, n, (cudaStream_t)stream);
}

int mirres_interpolate_bwd(const float *grad, int n, int C, const int *prim, const float *bary, const int *tri, int F,
                           float *out, void *stream)
{
    if (!grad || !prim || !tri || !out) return MIRRES_ERR_NULL;
    if (n < 0 || C < 1 || C > MR_SCATTER_MAX_C || F < 1) return MIRRES_ERR_SHAPE;
    if (n == 0) return 0;
    ScatterParams p = {grad, prim, bary, tri, out, n, C, F};
    return interpolate_bwd_launch(p, (cudaStream_t)stream);
}

} // extern "C"
