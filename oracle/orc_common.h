// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed
// by the product path (mirres_restir_nerf_mesh_b200/); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and its
// Slang sources cannot be compiled here (no slangc / slangpy).  This file
// set is a line-by-line restatement of the reference's .slang sources in scalar C++; every
// function cites the reference file:line it follows.
//
// orc_common.h : Slang intrinsic semantics, float3 helpers, RNG.
//   nerf/ScreenSpaceReSTIR/utils/random.slang:1-73
//   nerf/ScreenSpaceReSTIR/utils/helper.slang / helperDi.slang (math_lerp, luminance, oct codec)
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <stdint.h>
#include <math.h>
#include "../include/mirres_fpmath.h"

namespace orc {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct i2 { int x, y; };

static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 mk3(float s) { f3 r = {s, s, s}; return r; }
static inline f2 mk2(float x, float y) { f2 r = {x, y}; return r; }
static inline f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
static inline f3 &operator+=(f3 &a, f3 b) { a = a + b; return a; }
static inline f3 &operator*=(f3 &a, f3 b) { a = a * b; return a; }
static inline f3 &operator*=(f3 &a, float s) { a = a * s; return a; }

// Slang/HLSL intrinsics as lowered for the CUDA target (fminf/fmaxf NaN semantics).
static inline float smin(float a, float b) { return fminf(a, b); }
static inline float smax(float a, float b) { return fmaxf(a, b); }
static inline float sclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float saturate(float x) { return sclamp(x, 0.0f, 1.0f); }
static inline int iclamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float slerp(float a, float b, float t) { return a + (b - a) * t; }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross(f3 a, f3 b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(f3 a) { return sqrtf(dot(a, a)); }
static inline f3 normalize(f3 a) { return a / length(a); }
static inline f3 reflect(f3 i, f3 n) { return i - (2.0f * dot(n, i)) * n; }
// float -> int conversion with CUDA's saturating semantics (NaN -> 0) so the oracle has no UB.
static inline int f2i(float x)
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)x;
}
static inline uint32_t f2u(float x)
{
    if (x != x || x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
}

// ---- random.slang:2-39 ---------------------------------------------------------------------
static inline uint32_t interleave_32bit(uint32_t vx, uint32_t vy)
{
    uint32_t x = vx & 0x0000ffffu;
    uint32_t y = vy & 0x0000ffffu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu;
    y = (y | (y << 4)) & 0x0F0F0F0Fu;
    y = (y | (y << 2)) & 0x33333333u;
    y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}
static inline uint32_t seed_generator(uint32_t px, uint32_t py, uint32_t sampleNumber)
{
    uint32_t v0 = interleave_32bit(px, py), v1 = sampleNumber, sum = 0;
    const uint32_t delta = 0x9e3779b9u;
    const uint32_t k[4] = {0xa341316cu, 0xc8013ea4u, 0xad90777du, 0x7e95761eu};
    for (int i = 0; i < 16; i++) {
        sum += delta;
        v0 += ((v1 << 4) + k[0]) ^ (v1 + sum) ^ ((v1 >> 5) + k[1]);
        v1 += ((v0 << 4) + k[2]) ^ (v0 + sum) ^ ((v0 >> 5) + k[3]);
    }
    return v0;
}
// random.slang:41-55
static inline float next1d(uint32_t &sg)
{
    sg = 1664525u * sg + 1013904223u;
    return (float)(sg >> 8) * 0x1p-24f;
}

// ---- helper.slang:100-130 (identical in helperDi.slang:100-134) -------------------------------
static inline float luminance(f3 v) { return v.x * 0.212671f + v.y * 0.715160f + v.z * 0.072169f; }
static inline f2 oct_encode(f3 n)
{
    float s = fabsf(n.x) + fabsf(n.y) + fabsf(n.z);
    n = n / s;
    // oct_wrap(v) = (1 - abs(v.yx)) * sign
    float wx = (1.0f - fabsf(n.y)) * (n.x >= 0.0f ? 1.0f : -1.0f);
    float wy = (1.0f - fabsf(n.x)) * (n.y >= 0.0f ? 1.0f : -1.0f);
    float ox = n.z >= 0.0f ? n.x : wx;
    float oy = n.z >= 0.0f ? n.y : wy;
    return mk2(ox * 0.5f + 0.5f, oy * 0.5f + 0.5f);
}
static inline f3 oct_decode(f2 f)
{
    float fx = f.x * 2.0f - 1.0f, fy = f.y * 2.0f - 1.0f;
    f3 n = mk3(fx, fy, 1.0f - fabsf(fx) - fabsf(fy));
    float t = sclamp(-n.z, 0.0f, 1.0f);
    n.x += (n.x >= 0.0f ? -t : t);
    n.y += (n.y >= 0.0f ? -t : t);
    return normalize(n);
}
// lightDi.slang:430-435
static inline f3 ngp_dir(f3 d) { return mk3(-d.x, d.z, d.y); }
// helperDi.slang:396-409
static inline bool is_black(f3 v) { return !(v.x != 0.0f) && !(v.y != 0.0f) && !(v.z != 0.0f); }
static inline float power_heuristic(float p1, float p2) { return p1 * p1 / (p1 * p1 + p2 * p2); }

} // namespace orc
#endif
