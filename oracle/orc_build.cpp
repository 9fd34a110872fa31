// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).  PARITY UNPINNED.
//
// orc_build.cpp : LBVH construction, env distribution, light tiles, neighbour offsets, standalone rays.
//   nerf/renderer_restir.py:25-89 (update_bvh driver), nerf/bvhworkers/*.slang
//   nerf/ScreenSpaceReSTIR/GenerateLightTiles.py:4-29 + make_sampleable.slang:34-86,186-205
//   nerf/ScreenSpaceReSTIR/GenerateLightTiles.slang:16-62 + utils/light.slang:105-137,221-241
#include <algorithm>
#include <vector>
#include <string.h>
#include "orc_bvh.h"
#include "orc_light.h"

using namespace orc;

// lbvh_morton_codes.slang:24-45
static inline uint32_t expand_bits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3d(float x, float y, float z)
{
    x = smin(smax(x * 1024.0f, 0.0f), 1023.0f);
    y = smin(smax(y * 1024.0f, 0.0f), 1023.0f);
    z = smin(smax(z * 1024.0f, 0.0f), 1023.0f);
    uint32_t xx = expand_bits(f2u(x)), yy = expand_bits(f2u(y)), zz = expand_bits(f2u(z));
    return xx * 4 + yy * 2 + zz;
}
// lbvh_hierarchy.slang:31-52
static inline int find_msb(uint32_t v)
{
    if (v == 0) return -1;
    int msb = 31;
    while (!((v >> msb) & 1)) msb--;
    return msb;
}
static inline int delta(int i, uint32_t codeI, int j, int n, const int *codes)
{
    if (j < 0 || j > n - 1) return -1;
    uint32_t codeJ = (uint32_t)codes[2 * (size_t)j];
    if (codeI == codeJ) return 32 + 31 - find_msb((uint32_t)i ^ (uint32_t)j);
    return 31 - find_msb(codeI ^ codeJ);
}

extern "C" {

// Full rebuild.  Outputs follow renderer_restir.py:61-64: info [2F-1,3] i32, aabb [2F-1,6] f32.
// sorted_codes [F,2] (code, element idx) and parent [2F-1] are extra outputs for the tests.
int orc_bvh_build(const float *vert, int V, const int *tri, int F, int *info, float *aabb, int *sorted_codes,
                  int *parent_out, float *extent_out)
{
    (void)V;
    if (F < 1) return -1;
    std::vector<float> eaabb((size_t)F * 6);
    // get_elements.slang:3-39
    for (int p = 0; p < F; ++p) {
        float mn[3] = {1e9f, 1e9f, 1e9f}, mx[3] = {-1e9f, -1e9f, -1e9f};
        for (int i = 0; i < 3; ++i) {
            const float *v = vert + 3 * (size_t)tri[3 * (size_t)p + i];
            for (int k = 0; k < 3; ++k) { mn[k] = smin(mn[k], v[k]); mx[k] = smax(mx[k], v[k]); }
        }
        for (int k = 0; k < 3; ++k) {
            eaabb[6 * (size_t)p + k] = smin(mn[k], mx[k]);
            eaabb[6 * (size_t)p + 3 + k] = smax(mn[k], mx[k]);
        }
    }
    // renderer_restir.py:34-40 scene extent
    float g[6];
    for (int k = 0; k < 3; ++k) { g[k] = eaabb[k]; g[3 + k] = eaabb[3 + k]; }
    for (int p = 1; p < F; ++p)
        for (int k = 0; k < 3; ++k) {
            g[k] = smin(g[k], eaabb[6 * (size_t)p + k]);
            g[3 + k] = smax(g[3 + k], eaabb[6 * (size_t)p + 3 + k]);
        }
    if (extent_out) memcpy(extent_out, g, sizeof(g));
    // lbvh_morton_codes.slang:46-79
    std::vector<std::pair<uint32_t, int> > mc((size_t)F);
    for (int p = 0; p < F; ++p) {
        float c[3];
        for (int k = 0; k < 3; ++k) {
            float mnv = eaabb[6 * (size_t)p + k], mxv = eaabb[6 * (size_t)p + 3 + k];
            float center = mnv + 0.5f * (mxv - mnv);
            c[k] = (center - g[k]) / (g[3 + k] - g[k]);
        }
        mc[p] = std::make_pair(morton3d(c[0], c[1], c[2]), p);
    }
    // lbvh_single_radixsort.slang: 4 stable LSD passes == stable sort by the 32-bit key
    std::stable_sort(mc.begin(), mc.end(),
                     [](const std::pair<uint32_t, int> &a, const std::pair<uint32_t, int> &b) { return a.first < b.first; });
    std::vector<int> codes((size_t)F * 2);
    for (int p = 0; p < F; ++p) { codes[2 * (size_t)p] = (int)mc[p].first; codes[2 * (size_t)p + 1] = mc[p].second; }
    if (sorted_codes) memcpy(sorted_codes, codes.data(), sizeof(int) * 2 * (size_t)F);

    const int N = 2 * F - 1, LEAF = F - 1;
    std::vector<int> parent((size_t)N, 0);
    // lbvh_hierarchy.slang:111-245
    for (int gid = 0; gid < F; ++gid) {
        int e = codes[2 * (size_t)gid + 1];
        info[3 * (size_t)(LEAF + gid) + 0] = 0;
        info[3 * (size_t)(LEAF + gid) + 1] = 0;
        info[3 * (size_t)(LEAF + gid) + 2] = e; // ele_primitiveIdx[e] == e
        memcpy(aabb + 6 * (size_t)(LEAF + gid), &eaabb[6 * (size_t)e], 6 * sizeof(float));
    }
    for (int idx = 0; idx < F - 1; ++idx) {
        // determineRange :54-84
        uint32_t code = (uint32_t)codes[2 * (size_t)idx];
        int deltaL = delta(idx, code, idx - 1, F, codes.data());
        int deltaR = delta(idx, code, idx + 1, F, codes.data());
        int d = (deltaR >= deltaL) ? 1 : -1;
        int deltaMin = deltaL < deltaR ? deltaL : deltaR;
        int lMax = 2;
        while (delta(idx, code, idx + lMax * d, F, codes.data()) > deltaMin) lMax <<= 1;
        int l = 0;
        for (int t = lMax >> 1; t > 0; t >>= 1)
            if (delta(idx, code, idx + (l + t) * d, F, codes.data()) > deltaMin) l += t;
        int jdx = idx + l * d;
        int first = idx < jdx ? idx : jdx, last = idx < jdx ? jdx : idx;
        // findSplit :86-109
        uint32_t firstCode = (uint32_t)codes[2 * (size_t)first];
        int commonPrefix = delta(first, firstCode, last, F, codes.data());
        int split = first, stride = last - first;
        do {
            stride = (stride + 1) >> 1;
            int newSplit = split + stride;
            if (newSplit < last) {
                int splitPrefix = delta(first, firstCode, newSplit, F, codes.data());
                if (splitPrefix > commonPrefix) split = newSplit;
            }
        } while (stride > 1);
        int childA = (split == first) ? LEAF + split : split;
        int childB = (split + 1 == last) ? LEAF + split + 1 : split + 1;
        info[3 * (size_t)idx + 0] = childA;
        info[3 * (size_t)idx + 1] = childB;
        info[3 * (size_t)idx + 2] = 0;
        for (int k = 0; k < 3; ++k) { aabb[6 * (size_t)idx + k] = 1e9f; aabb[6 * (size_t)idx + 3 + k] = -1e9f; }
        parent[childA] = idx;
        parent[childB] = idx;
    }
    parent[0] = 0;
    if (parent_out) memcpy(parent_out, parent.data(), sizeof(int) * (size_t)N);
    // lbvh_bounding_boxes.slang:151-390.  The level-by-level schedule of get_bbox + set_root ends
    // with every internal node holding the exact min/max union of its two children (min/max are
    // exact, so the schedule does not matter); the oracle evaluates that fixed point bottom-up.
    if (F >= 2) {
        std::vector<int> order;
        order.reserve((size_t)F - 1);
        std::vector<int> st;
        st.push_back(0);
        while (!st.empty()) {
            int n = st.back();
            st.pop_back();
            order.push_back(n);
            int l = info[3 * (size_t)n], r = info[3 * (size_t)n + 1];
            if (l < LEAF) st.push_back(l);
            if (r < LEAF) st.push_back(r);
        }
        for (size_t q = order.size(); q-- > 0;) {
            int n = order[q];
            const float *a = aabb + 6 * (size_t)info[3 * (size_t)n];
            const float *b = aabb + 6 * (size_t)info[3 * (size_t)n + 1];
            for (int k = 0; k < 3; ++k) {
                aabb[6 * (size_t)n + k] = smin(a[k], b[k]);
                aabb[6 * (size_t)n + 3 + k] = smax(a[3 + k], b[3 + k]);
            }
        }
    }
    return 0;
}

// Standalone rays with the semantics of bvh_hit_with_normal (closest) -- helperDi.slang:313-395.
// counters[5] = nodes_ref, tris_ref, nodes_any, tris_any, max_stack (accumulated over all rays).
int orc_trace(const int *info, const float *aabb, const float *vert, const int *tri, const float *org, const float *dir,
              int n, int *hit, float *t, float *pos, float *normal, int *prim, long long *counters)
{
    Bvh b = {info, aabb, vert, tri};
    TraceCounters total = {0, 0, 0, 0, 0};
#pragma omp parallel
    {
        TraceCounters tc = {0, 0, 0, 0, 0};
#pragma omp for schedule(dynamic, 1024)
        for (int i = 0; i < n; ++i) {
            f3 o = mk3(org[3 * (size_t)i], org[3 * (size_t)i + 1], org[3 * (size_t)i + 2]);
            f3 d = mk3(dir[3 * (size_t)i], dir[3 * (size_t)i + 1], dir[3 * (size_t)i + 2]);
            float th = 0.f;
            f3 p = mk3(0.f), nn = mk3(1.f);
            int pr = -1;
            bool h = bvh_hit(b, o, d, 0.f, 1e7f, th, p, &nn, &pr, &tc);
            hit[i] = h ? 1 : 0;
            if (t) t[i] = th;
            if (pos) { pos[3 * (size_t)i] = p.x; pos[3 * (size_t)i + 1] = p.y; pos[3 * (size_t)i + 2] = p.z; }
            if (normal) { normal[3 * (size_t)i] = nn.x; normal[3 * (size_t)i + 1] = nn.y; normal[3 * (size_t)i + 2] = nn.z; }
            if (prim) prim[i] = pr;
        }
#pragma omp critical
        {
            total.nodes_ref += tc.nodes_ref; total.tris_ref += tc.tris_ref;
            total.nodes_any += tc.nodes_any; total.tris_any += tc.tris_any;
            if (tc.max_stack > total.max_stack) total.max_stack = tc.max_stack;
        }
    }
    if (counters) {
        counters[0] = total.nodes_ref; counters[1] = total.tris_ref;
        counters[2] = total.nodes_any; counters[3] = total.tris_any; counters[4] = total.max_stack;
    }
    return 0;
}

// make_sampleable.slang:34-58 (weights) + GenerateLightTiles.py:10-28 (scans; the torch parallel
// cumsum is restated as a sequential left-to-right fp32 prefix sum -- the build's defined order)
// + make_sampleable.slang:62-86 (Distribution2D row normalisation).
int orc_env_build_distribution(const float *env_tex, int W, int H, float *pdf_, float *cdf_, float *mpdf_, float *mcdf_)
{
    const float PI = 3.141592653589793f;
    Env e = {env_tex, W, H, 0, 0, 0, 0};
    for (int h = 0; h < H; ++h) {
        for (int w = 0; w < W; ++w) {
            float v = ((float)h + .5f) / (float)H;
            float sin_theta = mr_sinf(PI * v);
            f2 uv = mk2(((float)w + .5f) / (float)W, v);
            float theta = uv.y * PI, phi = uv.x * 2 * PI;
            float st, ct, sp, cp;
            mr_sincosf(theta, &st, &ct);
            mr_sincosf(phi, &sp, &cp);
            f3 raw_dir = mk3(st * cp, ct, st * sp);
            float wgt = luminance(env_le(ngp_dir(raw_dir), e));
            wgt *= sin_theta;
            pdf_[(size_t)h * W + w] = wgt;
        }
    }
    mcdf_[0] = 0.f;
    float macc = 0.f;
    for (int h = 0; h < H; ++h) {
        float acc = 0.f;
        cdf_[(size_t)h * (W + 1)] = 0.f;
        for (int w = 0; w < W; ++w) {
            acc += pdf_[(size_t)h * W + w];
            cdf_[(size_t)h * (W + 1) + w + 1] = acc;
        }
        mpdf_[h] = acc; // pdf_.sum(1): same sequential order
        macc += acc;
        mcdf_[h + 1] = macc;
    }
    for (int h = 0; h < H; ++h) {
        float row_weight = cdf_[(size_t)h * (W + 1) + W];
        for (int w = 0; w < W; ++w) {
            size_t ip = (size_t)h * W + w, ic = (size_t)h * (W + 1) + w;
            if (row_weight < 1e-4f) {
                pdf_[ip] = 1.0f / (float)W;
                cdf_[ic] = (float)w / (float)W;
            } else {
                pdf_[ip] /= row_weight;
                cdf_[ic] /= row_weight;
            }
        }
        cdf_[(size_t)h * (W + 1) + W] = 1.f;
    }
    float total = mcdf_[H];
    for (int h = 0; h < H; ++h) mpdf_[h] = mpdf_[h] / total;
    for (int h = 0; h <= H; ++h) mcdf_[h] = mcdf_[h] / total;
    mcdf_[H] = 1.f;
    return 0;
}

// make_sampleable.slang:186-205 (raw values; the host divides by 127, renderer_restir.py:221)
int orc_neighbor_offsets(int sampleCount, float *out)
{
    const int R = 254;
    const float phi2 = 1.f / 1.3247179572447f;
    float u = 0.5f, v = 0.5f;
    for (uint32_t index = 0; index < (uint32_t)sampleCount * 2;) {
        u += phi2;
        v += phi2 * phi2;
        if (u >= 1.f) u -= 1.f;
        if (v >= 1.f) v -= 1.f;
        float rSq = (u - 0.5f) * (u - 0.5f) + (v - 0.5f) * (v - 0.5f);
        if (rSq > 0.25f) continue;
        out[index++] = (float)f2i((u - 0.5f) * (float)R);
        out[index++] = (float)f2i((v - 0.5f) * (float)R);
    }
    return 0;
}

// GenerateLightTiles.slang:16-62; presampled_light light.slang:221-241.
int orc_light_tiles(const float *env_tex, int W, int H, const float *pdf_, const float *cdf_, const float *mpdf_,
                    const float *mcdf_, uint32_t frameIndex, int tile_count, int tile_size, float *light_data,
                    int *light_uv, float *light_pdf)
{
    Env e = {env_tex, W, H, pdf_, cdf_, mpdf_, mcdf_};
    const int n = tile_count * tile_size;
#pragma omp parallel for schedule(static)
    for (int bufferIndex = 0; bufferIndex < n; ++bufferIndex) {
        // Seed_Generator(uint -> uint2 splat): (b,b); the per-tile offset draw is a dead value.
        uint32_t sg = seed_generator((uint32_t)bufferIndex, (uint32_t)bufferIndex, frameIndex + 1);
        f2 rnd;
        rnd.x = next1d(sg);
        rnd.y = next1d(sg);
        f3 dir;
        float pdf;
        f2 luv;
        bool res = sample_li(e, rnd, dir, pdf, luv);
        float ld[3] = {0.f, 0.f, 0.f};
        int uvx = 0, uvy = 0;
        float ip = 0.f;
        if (res) {
            f2 o = oct_encode(dir);
            ld[0] = 1.f; ld[1] = o.x; ld[2] = o.y;
            i2 xy = uv2xy(luv, W, H);
            uvx = xy.x; uvy = xy.y;
            ip = pdf;
        }
        light_data[3 * (size_t)bufferIndex] = ld[0];
        light_data[3 * (size_t)bufferIndex + 1] = ld[1];
        light_data[3 * (size_t)bufferIndex + 2] = ld[2];
        light_uv[2 * (size_t)bufferIndex] = uvx;
        light_uv[2 * (size_t)bufferIndex + 1] = uvy;
        light_pdf[bufferIndex] = ip;
    }
    return 0;
}

} // extern "C"

// Elementwise access to include/mirres_fpmath.h for tests/test_fpmath.py.
// op: 0 sin, 1 cos, 2 acos, 3 atan2(x,y), 4 pow5, 5 pow8, 6 exp; also a contraction self-test (op 100).
extern "C" int orc_fpmath_eval(int op, const float *x, const float *y, float *out, int n)
{
    if (op == 100) {
        // a*b+c differs between fused and unfused evaluation for these values
        volatile float a = 1.0f + 0x1p-12f, b = 1.0f + 0x1p-12f, c = -(1.0f + 0x1p-11f);
        float r = a * b + c;
        out[0] = r;
        return r == 0.0f ? 0 : 1; // unfused: a*b rounds to 1+2^-11 -> 0; fused would give 2^-24
    }
    for (int i = 0; i < n; ++i) {
        switch (op) {
        case 0: out[i] = mr_sinf(x[i]); break;
        case 1: out[i] = mr_cosf(x[i]); break;
        case 2: out[i] = mr_acosf(x[i]); break;
        case 3: out[i] = mr_atan2f(x[i], y[i]); break;
        case 4: out[i] = mr_pow5f(x[i]); break;
        case 5: out[i] = mr_pow8f(x[i]); break;
        case 6: out[i] = mr_expf(x[i]); break;
        default: return -1;
        }
    }
    return 0;
}

// Host threads of the OpenMP regions.  Launchers such as torch.distributed.run export OMP_NUM_THREADS=1 into every rank;
// a caller that times the oracle as the CPU baseline (bench.py --impl reference) sets the count it means explicitly and
// reports the count that actually ran.
#include <omp.h>
extern "C" void orc_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}
extern "C" int orc_threads_in_use(void)
{
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
