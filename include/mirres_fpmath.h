/*
 * mirres_fpmath.h -- the numerical contract of the mirres-b200 hot path.
 *
 * The reference (nerf/ScreenSpaceReSTIR/utils/{lightDi,brdf,brdfDi}.slang) calls sin / cos / acos /
 * atan2 / pow through slangc -> CUDA libdevice.  Their last-ulp behaviour is implementation
 * defined, yet the path feeds them into integer decisions (env texel indices lightDi.slang:326-327,
 * CDF bins lightDi.slang:64-77, reservoir selections res.slang:101).  To make "bit-exact hit ids /
 * reservoir indices" a testable statement, BOTH the sm_100a kernels and the CPU oracle evaluate
 * the functions below: argument reduction and polynomials in IEEE double using only
 * mul / add / div / sqrt / rint (never contracted into FMA), rounded once to fp32.  The result is
 * within 0.5000001 ulp of the exact value, i.e. as close to libdevice as libdevice is to glibc.
 * (Exception: mr_expf, which feeds no integer decision, is evaluated in fp32 and is within 1.5 ulp.)
 *
 * Coefficients: tools/gen_fpmath.py (Chebyshev interpolation, max abs error < 4e-14).
 *
 * Compile rules (enforced by the build scripts and checked by tests/test_fpmath.py):
 *   host  : gcc/g++ -O2 -ffp-contract=off   (no -ffast-math)
 *   device: nvcc -fmad=false (default -prec-div=true -prec-sqrt=true -ftz=false)
 */
#ifndef MIRRES_FPMATH_H
#define MIRRES_FPMATH_H

#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define MR_HD __host__ __device__ __forceinline__
#else
#define MR_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define MR_DMUL(a, b) __dmul_rn((a), (b))
#define MR_DADD(a, b) __dadd_rn((a), (b))
#else
#define MR_DMUL(a, b) ((a) * (b))
#define MR_DADD(a, b) ((a) + (b))
#endif

#define MR_PI_D 3.14159265358979323846
#define MR_PIO2_D 1.57079632679489661923
#define MR_PIO4_D 0.78539816339744830962

/* sin(r)/r on r^2 in [0,(pi/4)^2] */
#define MR_SIN_C0 9.99999999999995226e-01
#define MR_SIN_C1 -1.66666666666116764e-01
#define MR_SIN_C2 8.33333332312805736e-03
#define MR_SIN_C3 -1.98412629122893318e-04
#define MR_SIN_C4 2.75551919526888805e-06
#define MR_SIN_C5 -2.47508350196426756e-08
/* cos(r) on r^2 */
#define MR_COS_C0 1.00000000000000044e+00
#define MR_COS_C1 -5.00000000000035194e-01
#define MR_COS_C2 4.16666666671404659e-02
#define MR_COS_C3 -1.38888889099774196e-03
#define MR_COS_C4 2.48015895633122784e-05
#define MR_COS_C5 -2.75566435482120100e-07
#define MR_COS_C6 2.06980159174020301e-09
/* asin(x)/x on x^2 in [0,0.25] */
#define MR_ASIN_C0 9.99999999999949707e-01
#define MR_ASIN_C1 1.66666666707388028e-01
#define MR_ASIN_C2 7.49999946556089703e-02
#define MR_ASIN_C3 4.46431278000407813e-02
#define MR_ASIN_C4 3.03750265117076514e-02
#define MR_ASIN_C5 2.24728839165493875e-02
#define MR_ASIN_C6 1.64708756961239183e-02
#define MR_ASIN_C7 1.86464085835412741e-02
#define MR_ASIN_C8 -2.87434826992541415e-03
#define MR_ASIN_C9 3.19610173731254277e-02
/* atan(t)/t on t^2 in [0,tan(pi/8)^2] */
#define MR_ATAN_C0 9.99999999999971023e-01
#define MR_ATAN_C1 -3.33333333306098600e-01
#define MR_ATAN_C2 1.99999995788744694e-01
#define MR_ATAN_C3 -1.42856891030847366e-01
#define MR_ATAN_C4 1.11103523705183130e-01
#define MR_ATAN_C5 -9.07795587446554170e-02
#define MR_ATAN_C6 7.56049521386349954e-02
#define MR_ATAN_C7 -5.86227534442604128e-02
#define MR_ATAN_C8 3.04738117243655501e-02

#define MR_H(p, z, c) MR_DADD(MR_DMUL((p), (z)), (c))

/* sin and cos of x (radians), |x| < 1e6; NaN otherwise. */
MR_HD void mr_sincosf(float x, float *s_out, float *c_out)
{
    double xd = (double)x;
    if (!(fabs(xd) < 1.0e6)) { /* inf, NaN or out-of-contract magnitude -> NaN */
        float q = x - x;
        *s_out = q / q;
        *c_out = q / q;
        return;
    }
    double kd = rint(MR_DMUL(xd, 0.63661977236758134308));
    int k = (int)kd;
    double r = MR_DADD(MR_DADD(xd, -MR_DMUL(kd, 1.57079632679489655800e+00)), -MR_DMUL(kd, 6.12323399573676603587e-17));
    double z = MR_DMUL(r, r);
    double ps = MR_SIN_C5;
    ps = MR_H(ps, z, MR_SIN_C4);
    ps = MR_H(ps, z, MR_SIN_C3);
    ps = MR_H(ps, z, MR_SIN_C2);
    ps = MR_H(ps, z, MR_SIN_C1);
    ps = MR_H(ps, z, MR_SIN_C0);
    double s = MR_DMUL(r, ps);
    double pc = MR_COS_C6;
    pc = MR_H(pc, z, MR_COS_C5);
    pc = MR_H(pc, z, MR_COS_C4);
    pc = MR_H(pc, z, MR_COS_C3);
    pc = MR_H(pc, z, MR_COS_C2);
    pc = MR_H(pc, z, MR_COS_C1);
    pc = MR_H(pc, z, MR_COS_C0);
    double c = pc;
    double so, co;
    switch (k & 3) {
    case 0: so = s; co = c; break;
    case 1: so = c; co = -s; break;
    case 2: so = -s; co = -c; break;
    default: so = -c; co = s; break;
    }
    *s_out = (float)so;
    *c_out = (float)co;
}

MR_HD float mr_sinf(float x)
{
    float s, c;
    mr_sincosf(x, &s, &c);
    return s;
}

MR_HD float mr_cosf(float x)
{
    float s, c;
    mr_sincosf(x, &s, &c);
    return c;
}

/* asin(x)/x polynomial in z = x*x */
MR_HD double mr_asin_poly(double z)
{
    double p = MR_ASIN_C9;
    p = MR_H(p, z, MR_ASIN_C8);
    p = MR_H(p, z, MR_ASIN_C7);
    p = MR_H(p, z, MR_ASIN_C6);
    p = MR_H(p, z, MR_ASIN_C5);
    p = MR_H(p, z, MR_ASIN_C4);
    p = MR_H(p, z, MR_ASIN_C3);
    p = MR_H(p, z, MR_ASIN_C2);
    p = MR_H(p, z, MR_ASIN_C1);
    p = MR_H(p, z, MR_ASIN_C0);
    return p;
}

/* acos(x); NaN outside [-1,1]. */
MR_HD float mr_acosf(float x)
{
    double xd = (double)x;
    double ax = fabs(xd);
    if (ax <= 0.5) {
        double z = MR_DMUL(xd, xd);
        return (float)MR_DADD(MR_PIO2_D, -MR_DMUL(xd, mr_asin_poly(z)));
    }
    double z = MR_DMUL(MR_DADD(1.0, -ax), 0.5);
    double s = sqrt(z); /* NaN when |x| > 1 or x is NaN */
    double h = MR_DMUL(2.0, MR_DMUL(s, mr_asin_poly(z)));
    if (xd > 0.0) return (float)h;
    return (float)MR_DADD(MR_PI_D, -h);
}

/* atan2(y, x) in (-pi, pi]. */
MR_HD float mr_atan2f(float y, float x)
{
    double xd = (double)x, yd = (double)y;
    if (xd != xd || yd != yd) return x + y;
    double ax = fabs(xd), ay = fabs(yd);
    double mx = ax > ay ? ax : ay;
    double mn = ax > ay ? ay : ax;
    double r;
    if (mx == 0.0) {
        r = 0.0;
    } else {
        double a = (mx == INFINITY) ? ((mn == INFINITY) ? 1.0 : 0.0) : mn / mx;
        double t = a, off = 0.0;
        if (a > 0.41421356237309503) {
            t = MR_DADD(a, -1.0) / MR_DADD(a, 1.0);
            off = MR_PIO4_D;
        }
        double z = MR_DMUL(t, t);
        double p = MR_ATAN_C8;
        p = MR_H(p, z, MR_ATAN_C7);
        p = MR_H(p, z, MR_ATAN_C6);
        p = MR_H(p, z, MR_ATAN_C5);
        p = MR_H(p, z, MR_ATAN_C4);
        p = MR_H(p, z, MR_ATAN_C3);
        p = MR_H(p, z, MR_ATAN_C2);
        p = MR_H(p, z, MR_ATAN_C1);
        p = MR_H(p, z, MR_ATAN_C0);
        r = MR_DADD(off, MR_DMUL(t, p));
        if (ay > ax) r = MR_DADD(MR_PIO2_D, -r);
    }
    if (signbit(xd)) r = MR_DADD(MR_PI_D, -r);
    if (signbit(yd)) r = -r;
    return (float)r;
}

/* exp(x) for the edge-stopping weights of the a-trous filter (EAWDenoise.slang:157-167).  The filter evaluates three
 * exps per tap, 25 taps per pixel, and no integer decision depends on them, so this one function trades the
 * "correctly rounded" property of the others for speed: fp32 arithmetic only (the double-precision version cost
 * 0.35 ms of a 6.7 ms step on B200).  x = k ln2 + r with a two-constant Cody-Waite reduction, degree-7 Taylor
 * polynomial of exp(r) on |r| <= ln2/2 (truncation < 5e-9), 2^k applied in two exact steps.  Every operation is an
 * IEEE fp32 multiply, add or rint that is never contracted, so host and device agree bit for bit; the result is
 * within 1.5 ulp of exp(x) (measured max 1.16; libdevice expf, which the reference binary calls, is specified to 2 ulp). */
#if defined(__CUDA_ARCH__)
#define MR_FMUL(a, b) __fmul_rn((a), (b))
#define MR_FADD(a, b) __fadd_rn((a), (b))
#else
#define MR_FMUL(a, b) ((a) * (b))
#define MR_FADD(a, b) ((a) + (b))
#endif
#define MR_HF(p, z, c) MR_FADD(MR_FMUL((p), (z)), (c))
MR_HD float mr_pow2i(int k) /* 2^k for k in [-126, 127] */
{
    int bits = (127 + k) << 23;
    float f;
#if defined(__CUDA_ARCH__)
    f = __int_as_float(bits);
#else
    memcpy(&f, &bits, sizeof(f));
#endif
    return f;
}
MR_HD float mr_expf(float x)
{
    if (x != x) return x;
    if (x > 88.8f) return INFINITY;
    if (x < -104.0f) return 0.0f;
    float kf = rintf(MR_FMUL(x, 1.44269504088896341f));
    float r = MR_FADD(x, -MR_FMUL(kf, 0.693145751953125f));       /* ln2 high part: 11 significant bits, kf * hi is exact */
    r = MR_FADD(r, -MR_FMUL(kf, 1.42860682030941723e-06f));       /* ln2 low part */
    float p = 1.98412698412698413e-04f;                           /* 1/7! */
    p = MR_HF(p, r, 1.38888888888888894e-03f);
    p = MR_HF(p, r, 8.33333333333333322e-03f);
    p = MR_HF(p, r, 4.16666666666666644e-02f);
    p = MR_HF(p, r, 1.66666666666666657e-01f);
    p = MR_HF(p, r, 0.5f);
    p = MR_HF(p, r, 1.0f);
    p = MR_HF(p, r, 1.0f);
    int k = (int)kf;                                              /* [-151, 129] */
    int k1 = k / 2;
    return MR_FMUL(MR_FMUL(p, mr_pow2i(k1)), mr_pow2i(k - k1));   /* first product exact, second rounds once */
}

/* pow(x,5) and pow(x,8) as the reference uses them (brdf.slang:27, res.slang:55), by squaring. */
MR_HD float mr_pow5f(float x)
{
    float x2 = x * x;
    float x4 = x2 * x2;
    return x4 * x;
}

/* pow(x,128) of the cross-bilateral normal weight (nerf/renderutils/c_src/denoising.cu:56), by seven squarings. */
MR_HD float mr_pow128f(float x)
{
    float x2 = x * x;
    float x4 = x2 * x2;
    float x8 = x4 * x4;
    float x16 = x8 * x8;
    float x32 = x16 * x16;
    float x64 = x32 * x32;
    return x64 * x64;
}

MR_HD float mr_pow8f(float x)
{
    float x2 = x * x;
    float x4 = x2 * x2;
    return x4 * x4;
}

/*
 * TEST-ONLY flavour (never compiled into the product): -DMR_FPMATH_LIBM redirects the functions above to the C library,
 * and the oracle is then built with -ffp-contract=fast -mfma.  That is the closest stand-in available here for the
 * numerics of the reference's real binary (slangc -> nvcc: FMA contraction on, libdevice transcendentals, pow() for the
 * Fresnel / falloff powers), which cannot be built in this environment.  tools/numerics_sensitivity.py runs the oracle in
 * both flavours on the same inputs and reports which integer decisions flip and how far the float outputs move
 * (profiles/numerics_sensitivity.json): a bound on what "parity unpinned" can hide.
 */
#if defined(MR_FPMATH_LIBM) && !defined(__CUDACC__)
static inline float mr_libm_sinf(float x) { return sinf(x); }
static inline float mr_libm_cosf(float x) { return cosf(x); }
static inline void mr_libm_sincosf(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
static inline float mr_libm_acosf(float x) { return acosf(x); }
static inline float mr_libm_atan2f(float y, float x) { return atan2f(y, x); }
static inline float mr_libm_expf(float x) { return expf(x); }
static inline float mr_libm_pow5f(float x) { return powf(x, 5.0f); }
static inline float mr_libm_pow8f(float x) { return powf(x, 8.0f); }
static inline float mr_libm_pow128f(float x) { return powf(x, 128.0f); }
#define mr_sinf mr_libm_sinf
#define mr_cosf mr_libm_cosf
#define mr_sincosf mr_libm_sincosf
#define mr_acosf mr_libm_acosf
#define mr_atan2f mr_libm_atan2f
#define mr_expf mr_libm_expf
#define mr_pow5f mr_libm_pow5f
#define mr_pow8f mr_libm_pow8f
#define mr_pow128f mr_libm_pow128f
#endif

#endif /* MIRRES_FPMATH_H */
