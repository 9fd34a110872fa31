// mirres-b200: environment-map distribution, light-tile presampling, neighbour-offset table.
//
// Replaces nerf/ScreenSpaceReSTIR/GenerateLightTiles.py:4-29 (make_sampleable: 2 Slang kernels + 5 torch
// reductions/scans) with two launches, nerf/ScreenSpaceReSTIR/GenerateLightTiles.slang:16-62 (launched by the
// reference with a (1024,128)x256 grid of which 99.6 % of the threads exit) with one dense launch, and
// nerf/ScreenSpaceReSTIR/make_sampleable.slang:186-205 (createNeighborOffsetTexture).
//
// Scan order is part of the contract: row CDFs and the marginal CDF are sequential left-to-right fp32 prefix
// sums (the reference uses torch.cumsum, whose GPU order is unspecified); the oracle uses the same order.
#include "mr_light.cuh"
#include "../../include/mirres_b200.h"

namespace mr {

MR_DEV float env_weight(const EnvView &e, int h, int w)
{
    const float PI = 3.141592653589793f;
    float v = ((float)h + .5f) / (float)e.H;
    float sin_theta = mr_sinf(PI * v);
    float2 uv = make_float2(((float)w + .5f) / (float)e.W, v);
    float theta = uv.y * PI, phi = uv.x * 2 * PI;
    float st, ct, sp, cp;
    mr_sincosf(theta, &st, &ct);
    mr_sincosf(phi, &sp, &cp);
    float3 raw_dir = make_float3(st * cp, ct, st * sp);
    float wgt = luminance(env_radiance(e, ngp_dir(raw_dir)));
    wgt *= sin_theta;
    return wgt;
}

// granular: make_sampleable.slang:34-58
struct WeightParams { EnvView e; float *__restrict__ weight; };
MR_DEV void env_weight_px(const WeightParams &p, int i) { p.weight[i] = env_weight(p.e, i / p.e.W, i % p.e.W); }

// granular: make_sampleable.slang:62-86 (cdf_ tails hold the raw row sums on entry)
struct Dist2dParams { int w, h; float *pdf_; float *cdf_; };
MR_DEV void distribution2d_px(const Dist2dParams &p, int i)
{
    const int w = p.w;
    float *pdf_ = p.pdf_, *cdf_ = p.cdf_;
    int y = i / w, x = i % w;
    float row_weight = cdf_[(size_t)y * (w + 1) + w];
    size_t ip = (size_t)y * w + x, ic = (size_t)y * (w + 1) + x;
    if (row_weight < 1e-4f) {
        pdf_[ip] = 1.0f / (float)w;
        cdf_[ic] = (float)x / (float)w;
    } else {
        pdf_[ip] /= row_weight;
        cdf_[ic] /= row_weight;
    }
}

#if !defined(MR_HOST_CHECK)
// fused: one block per env row.  weights -> shared, sequential prefix sum, row normalisation.
__global__ void __launch_bounds__(256) k_env_rows(EnvView e, float *__restrict__ pdf_, float *__restrict__ cdf_,
                                                  float *__restrict__ row_sum)
{
    extern __shared__ float sh[]; // W weights, then W+1 prefix values
    const int W = e.W, h = blockIdx.x;
    float *wgt = sh, *pre = sh + W;
    for (int w = threadIdx.x; w < W; w += blockDim.x) wgt[w] = env_weight(e, h, w);
    __syncthreads();
    if (threadIdx.x == 0) {
        float acc = 0.f;
        pre[0] = 0.f;
        for (int w = 0; w < W; ++w) {
            acc += wgt[w];
            pre[w + 1] = acc;
        }
        row_sum[h] = acc;
    }
    __syncthreads();
    const float row_weight = pre[W];
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        float p, c;
        if (row_weight < 1e-4f) {
            p = 1.0f / (float)W;
            c = (float)w / (float)W;
        } else {
            p = wgt[w] / row_weight;
            c = pre[w] / row_weight;
        }
        pdf_[(size_t)h * W + w] = p;
        cdf_[(size_t)h * (W + 1) + w] = c;
    }
    if (threadIdx.x == 0) cdf_[(size_t)h * (W + 1) + W] = 1.f;
}

// fused: marginal distribution over rows (GenerateLightTiles.py:14-28)
__global__ void __launch_bounds__(256) k_env_marginal(int H, const float *__restrict__ row_sum, float *__restrict__ mpdf_,
                                                      float *__restrict__ mcdf_)
{
    extern __shared__ float sh[]; // H+1 prefix values
    if (threadIdx.x == 0) {
        float acc = 0.f;
        sh[0] = 0.f;
        for (int h = 0; h < H; ++h) {
            acc += row_sum[h];
            sh[h + 1] = acc;
        }
    }
    __syncthreads();
    const float total = sh[H];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        mpdf_[h] = row_sum[h] / total;
        mcdf_[h] = sh[h] / total;
    }
    if (threadIdx.x == 0) mcdf_[H] = 1.f;
}

#else
// host-check flavour of the two fused kernels above: same per-texel weight routine, same sequential scans
static void env_rows_host(EnvView e, float *pdf_, float *cdf_, float *row_sum)
{
    const int W = e.W;
    for (int h = 0; h < e.H; ++h) {
        float acc = 0.f;
        for (int w = 0; w < W; ++w) {
            pdf_[(size_t)h * W + w] = env_weight(e, h, w);
            cdf_[(size_t)h * (W + 1) + w] = acc;
            acc += pdf_[(size_t)h * W + w];
        }
        row_sum[h] = acc;
        for (int w = 0; w < W; ++w) {
            if (acc < 1e-4f) {
                pdf_[(size_t)h * W + w] = 1.0f / (float)W;
                cdf_[(size_t)h * (W + 1) + w] = (float)w / (float)W;
            } else {
                pdf_[(size_t)h * W + w] = pdf_[(size_t)h * W + w] / acc;
                cdf_[(size_t)h * (W + 1) + w] = cdf_[(size_t)h * (W + 1) + w] / acc;
            }
        }
        cdf_[(size_t)h * (W + 1) + W] = 1.f;
    }
}
static void env_marginal_host(int H, const float *row_sum, float *mpdf_, float *mcdf_)
{
    float acc = 0.f;
    for (int h = 0; h < H; ++h) { mcdf_[h] = acc; acc += row_sum[h]; }
    const float total = acc;
    for (int h = 0; h < H; ++h) { mpdf_[h] = row_sum[h] / total; mcdf_[h] = mcdf_[h] / total; }
    mcdf_[H] = 1.f;
}
#endif

struct OffsetParams { int sampleCount; float *out; };
MR_DEV void neighbor_offsets_one(const OffsetParams &p, int)
{
    const int sampleCount = p.sampleCount;
    float *out = p.out;
    const int R = 254;
    const float phi2 = 1.f / 1.3247179572447f;
    float u = 0.5f, v = 0.5f;
    for (unsigned int index = 0; index < (unsigned int)sampleCount * 2;) {
        u += phi2;
        v += phi2 * phi2;
        if (u >= 1.f) u -= 1.f;
        if (v >= 1.f) v -= 1.f;
        float rSq = (u - 0.5f) * (u - 0.5f) + (v - 0.5f) * (v - 0.5f);
        if (rSq > 0.25f) continue;
        out[index++] = (float)to_int((u - 0.5f) * (float)R);
        out[index++] = (float)to_int((v - 0.5f) * (float)R);
    }
}

// GenerateLightTiles.slang:16-62 -- the per-tile stratification offset is a dead value in the reference
// (light.slang:221-229 ignores `random`), so only the per-sample stream Seed((b,b), frame+1) is drawn.
struct TileParams {
    EnvView e;
    unsigned int frameIndex;
    const unsigned int *__restrict__ frame_offset; // optional device-resident offset added to frameIndex
    float *__restrict__ light_data;
    int *__restrict__ light_uv;
    float *__restrict__ light_pdf;
    float4 *__restrict__ cache; // optional [T,2]: world direction and emitted radiance of every slot
};
MR_DEV void light_tile_px(const TileParams &p, int b)
{
    const EnvView &e = p.e;
    const unsigned int frameIndex = p.frameIndex + (p.frame_offset ? MR_LDG(p.frame_offset) : 0u);
    float *light_data = p.light_data;
    int *light_uv = p.light_uv;
    float *light_pdf = p.light_pdf;
    uint32_t sg = seed_of((uint32_t)b, (uint32_t)b, frameIndex + 1u);
    float2 u;
    u.x = rnd(sg);
    u.y = rnd(sg);
    float3 dir;
    float pdf;
    float2 luv;
    bool ok = sample_env(e, u, dir, pdf, luv);
    float3 ld = f3(0.f);
    int2 xy = make_int2(0, 0);
    float ip = 0.f;
    if (ok) {
        float2 o = oct_encode(dir);
        ld = make_float3(1.f, o.x, o.y);
        xy = texel_of_uv(luv, e.W, e.H);
        ip = pdf;
    }
    store3(light_data, (size_t)b, ld);
    light_uv[2 * (size_t)b] = xy.x;
    light_uv[2 * (size_t)b + 1] = xy.y;
    light_pdf[b] = ip;
    if (p.cache) {
        // what every pixel that draws this slot would compute from it (InitialResampling.slang:208-209)
        float3 Le, L;
        light_of(e, ld.y, ld.z, Le, L);
        p.cache[2 * (size_t)b] = make_float4(L.x, L.y, L.z, 0.f);
        p.cache[2 * (size_t)b + 1] = make_float4(Le.x, Le.y, Le.z, 0.f);
    }
}

} // namespace mr

using namespace mr;

extern "C" {

int mirres_env_weights(const float *env_tex, int W, int H, float *weight, void *stream)
{
    if (!env_tex || !weight) return MIRRES_ERR_NULL;
    if (W < 1 || H < 1) return MIRRES_ERR_SHAPE;
    WeightParams p = {{env_tex, W, H, nullptr, nullptr, nullptr, nullptr}, weight};
    return foreach_item<WeightParams, env_weight_px, 256>(p, W * H, (cudaStream_t)stream);
}

int mirres_env_distribution2d(int W, int H, float *pdf_, float *cdf_, void *stream)
{
    if (!pdf_ || !cdf_) return MIRRES_ERR_NULL;
    if (W < 1 || H < 1) return MIRRES_ERR_SHAPE;
    Dist2dParams p = {W, H, pdf_, cdf_};
    return foreach_item<Dist2dParams, distribution2d_px, 256>(p, W * H, (cudaStream_t)stream);
}

int mirres_env_build_distribution(const float *env_tex, int W, int H, float *pdf_, float *cdf_, float *mpdf_,
                                  float *mcdf_, float *row_scratch, void *stream)
{
    if (!env_tex || !pdf_ || !cdf_ || !mpdf_ || !mcdf_ || !row_scratch) return MIRRES_ERR_NULL;
    if (W < 1 || H < 1 || W > 8192 || H > 8192) return MIRRES_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    EnvView e = {env_tex, W, H, nullptr, nullptr, nullptr, nullptr};
#if defined(MR_HOST_CHECK)
    (void)st;
    env_rows_host(e, pdf_, cdf_, row_scratch);
    env_marginal_host(H, row_scratch, mpdf_, mcdf_);
    return 0;
#else
    size_t sh_rows = sizeof(float) * (size_t)(2 * W + 1);
    if (sh_rows > 48 * 1024) cudaFuncSetAttribute(k_env_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_rows);
    k_env_rows<<<H, 256, sh_rows, st>>>(e, pdf_, cdf_, row_scratch);
    k_env_marginal<<<1, 256, sizeof(float) * (size_t)(H + 1), st>>>(H, row_scratch, mpdf_, mcdf_);
    MR_CUDA_CHECK_LAUNCH();
    return 0;
#endif
}

int mirres_neighbor_offsets(int sample_count, float *out, void *stream)
{
    if (!out) return MIRRES_ERR_NULL;
    if (sample_count < 1) return MIRRES_ERR_SHAPE;
    OffsetParams p = {sample_count, out};
    return foreach_item<OffsetParams, neighbor_offsets_one, 32>(p, 1, (cudaStream_t)stream);
}

int mirres_light_tiles(const float *env_tex, int W, int H, const float *pdf_, const float *cdf_, const float *mpdf_,
                       const float *mcdf_, unsigned int frame_index, int tile_count, int tile_size, float *light_data,
                       int *light_uv, float *light_pdf, float *light_cache, const unsigned int *frame_offset, void *stream)
{
    if (!env_tex || !pdf_ || !cdf_ || !mpdf_ || !mcdf_ || !light_data || !light_uv || !light_pdf) return MIRRES_ERR_NULL;
    if (W < 1 || H < 1 || tile_count < 1 || tile_size < 1) return MIRRES_ERR_SHAPE;
    if ((uintptr_t)light_cache & 15) return MIRRES_ERR_ALIGN;
    TileParams p = {{env_tex, W, H, pdf_, cdf_, mpdf_, mcdf_}, frame_index, frame_offset, light_data, light_uv, light_pdf, (float4 *)light_cache};
    return foreach_item<TileParams, light_tile_px, 256>(p, tile_count * tile_size, (cudaStream_t)stream);
}

} // extern "C"
