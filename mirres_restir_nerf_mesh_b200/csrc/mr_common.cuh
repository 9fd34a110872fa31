// mirres-b200 device primitives: float3 algebra with the evaluation order of the Slang intrinsics,
// the TEA/LCG random stream, octahedral light codec.  sm_100a only; compiled with -fmad=false so that
// every fp32 result is the IEEE single operation the numerical contract (include/mirres_fpmath.h) names.
//
// Reference semantics restated here:
//   nerf/ScreenSpaceReSTIR/utils/random.slang:1-73      (interleave_32bit, blockCipherTEA, LCG, 24-bit floats)
//   nerf/ScreenSpaceReSTIR/utils/helperDi.slang:96-134  (luminance, oct_encode, oct_decode)
//   nerf/ScreenSpaceReSTIR/utils/lightDi.slang:430-435  (ngp_dir)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/mirres_fpmath.h"

// Every per-pixel routine is __host__ __device__: the shipped library (libmirres_b200.so) only ever runs the
// device side; a second, test-only build of the same sources (-DMR_HOST_CHECK, tests/hostcheck.py) runs the
// identical bodies in host loops so parity against the oracle can be checked without a GPU.
#define MR_DEV __host__ __device__ __forceinline__

#if defined(__CUDA_ARCH__)
#define MR_LDG(p) __ldg(p)
#else
#define MR_LDG(p) (*(p))
#endif

#define MR_CUDA_CHECK_LAUNCH()                                  \
    do {                                                        \
        cudaError_t e__ = cudaGetLastError();                   \
        if (e__ != cudaSuccess) return -100 - (int)e__;         \
    } while (0)

namespace mr {

MR_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
MR_DEV float3 f3(float s) { return make_float3(s, s, s); }
MR_DEV float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
MR_DEV float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
MR_DEV float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
MR_DEV float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
MR_DEV float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
MR_DEV float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
MR_DEV float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
MR_DEV void operator+=(float3 &a, float3 b) { a = a + b; }
MR_DEV void operator*=(float3 &a, float3 b) { a = a * b; }
MR_DEV void operator*=(float3 &a, float s) { a = a * s; }

MR_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MR_DEV float3 cross(float3 a, float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
MR_DEV float3 normalize(float3 a) { return a / sqrtf(dot(a, a)); }
MR_DEV float3 reflect(float3 i, float3 n) { return i - (2.0f * dot(n, i)) * n; }
MR_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
MR_DEV float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
MR_DEV float lerpf(float a, float b, float t) { return a + (b - a) * t; }
MR_DEV int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
MR_DEV uint32_t minu(uint32_t a, uint32_t b) { return a < b ? a : b; }
MR_DEV bool is_black(float3 v) { return !(v.x != 0.0f) && !(v.y != 0.0f) && !(v.z != 0.0f); }
MR_DEV float luminance(float3 v) { return v.x * 0.212671f + v.y * 0.715160f + v.z * 0.072169f; }
MR_DEV float3 ngp_dir(float3 d) { return make_float3(-d.x, d.z, d.y); }
MR_DEV float power_heuristic(float a, float b) { return a * a / (a * a + b * b); }

// float -> integer conversions with the device semantics (cvt.rzi: saturating, NaN -> 0) on both sides
MR_DEV int to_int(float x)
{
#if defined(__CUDA_ARCH__)
    return (int)x;
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)x;
#endif
}
MR_DEV uint32_t to_uint(float x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)x;
#else
    if (x != x || x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
#endif
}

// `tensor / python_scalar` as torch evaluates it: CUDA multiplies by the fp32 reciprocal of the scalar
// (BinaryDivTrueKernel.cu, "a * reciprocal(b)" for a CPU-scalar divisor), the CPU kernel divides.  The driver's
// `total / mFrameIndex` (nerf/renderer_restir.py:505-515) and its autograd backward are such divisions.
MR_DEV float div_by_scalar(float x, float d)
{
#if defined(__CUDA_ARCH__)
    return x * (1.0f / d);
#else
    return x / d;
#endif
}

// [n,3] fp32 rows (12-byte stride, the reference's external layout)
MR_DEV float3 load3(const float *__restrict__ p, size_t i) { return make_float3(MR_LDG(p + 3 * i), MR_LDG(p + 3 * i + 1), MR_LDG(p + 3 * i + 2)); }
MR_DEV float3 load3_rw(const float *p, size_t i) { return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
MR_DEV void store3(float *p, size_t i, float3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// ---- random stream -------------------------------------------------------------------------
MR_DEV uint32_t spread16(uint32_t x)
{
    x &= 0x0000ffffu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
MR_DEV uint32_t seed_of(uint32_t px, uint32_t py, uint32_t sample_number)
{
    uint32_t v0 = spread16(px) | (spread16(py) << 1);
    uint32_t v1 = sample_number;
    uint32_t sum = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
MR_DEV float rnd(uint32_t &state)
{
    state = 1664525u * state + 1013904223u;
    return (float)(state >> 8) * 0x1p-24f;
}

// ---- octahedral codec ------------------------------------------------------------------------
MR_DEV float2 oct_encode(float3 n)
{
    float s = fabsf(n.x) + fabsf(n.y) + fabsf(n.z);
    n = n / s;
    float wx = (1.0f - fabsf(n.y)) * (n.x >= 0.0f ? 1.0f : -1.0f);
    float wy = (1.0f - fabsf(n.x)) * (n.y >= 0.0f ? 1.0f : -1.0f);
    float ox = n.z >= 0.0f ? n.x : wx;
    float oy = n.z >= 0.0f ? n.y : wy;
    return make_float2(ox * 0.5f + 0.5f, oy * 0.5f + 0.5f);
}
MR_DEV float3 oct_decode(float ex, float ey)
{
    float fx = ex * 2.0f - 1.0f, fy = ey * 2.0f - 1.0f;
    float3 n = make_float3(fx, fy, 1.0f - fabsf(fx) - fabsf(fy));
    float t = clampf(-n.z, 0.0f, 1.0f);
    n.x += (n.x >= 0.0f ? -t : t);
    n.y += (n.y >= 0.0f ? -t : t);
    return normalize(n);
}

// ---- bit casts usable on both sides ------------------------------------------------------------
MR_DEV int float_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
    return i;
#endif
}
MR_DEV float bits_float(int i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

// ---- one-thread-per-item launcher -----------------------------------------------------------------
template <class P, void (*BODY)(const P &, int), int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_foreach(const P p, int n)
{
    int idx = blockIdx.x * BLOCK + threadIdx.x;
    if (idx < n) BODY(p, idx);
}

#if defined(MR_HOST_CHECK)
template <class P, void (*BODY)(const P &, int), int BLOCK>
static inline int foreach_item(const P &p, int n, cudaStream_t)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) BODY(p, i);
    return 0;
}
#else
template <class P, void (*BODY)(const P &, int), int BLOCK>
static inline int foreach_item(const P &p, int n, cudaStream_t st)
{
    k_foreach<P, BODY, BLOCK><<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(p, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -100 - (int)e;
}
#endif

// Zero-fill that is stream-ordered on the device and immediate in the host-check flavour.  A KERNEL, not
// cudaMemsetAsync: inside a captured CUDA graph a memset node runs on a copy engine, and the hand-over between that
// engine and the SMs cost 45-90 us per memset on the serial front of the step (B200 timeline, profiles/README.md);
// a kernel node costs ~2 us.  Up to five regions per launch (a reservoir is four arrays, plus the queue counters).
#define MR_ZERO_REGIONS 5
struct ZeroRegions {
    void *ptr[MR_ZERO_REGIONS];
    size_t words[MR_ZERO_REGIONS]; // 32-bit words
};
#if !defined(MR_HOST_CHECK)
static __global__ void __launch_bounds__(256) k_zero_regions(ZeroRegions z)
{
    const size_t stride = (size_t)gridDim.x * 256;
#pragma unroll
    for (int k = 0; k < MR_ZERO_REGIONS; ++k) {
        uint32_t *p = static_cast<uint32_t *>(z.ptr[k]);
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < z.words[k]; i += stride) p[i] = 0u;
    }
}
#endif
static inline void zero_regions_async(const ZeroRegions &z, cudaStream_t st)
{
#if defined(MR_HOST_CHECK)
    (void)st;
    for (int k = 0; k < MR_ZERO_REGIONS; ++k)
        if (z.words[k]) memset(z.ptr[k], 0, z.words[k] * 4);
#else
    size_t most = 0;
    for (int k = 0; k < MR_ZERO_REGIONS; ++k) most = z.words[k] > most ? z.words[k] : most;
    if (!most) return;
    size_t blocks = (most + 1023) / 1024; // four words per thread
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_zero_regions<<<(unsigned int)blocks, 256, 0, st>>>(z);
#endif
}
static inline void zero_async(void *ptr, size_t bytes, cudaStream_t st)
{
    ZeroRegions z = {{ptr, nullptr, nullptr, nullptr, nullptr}, {bytes / 4, 0, 0, 0, 0}};
    zero_regions_async(z, st);
}

} // namespace mr
