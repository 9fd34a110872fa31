"""Generate the polynomial coefficients used by include/mirres_fpmath.h.

The path's integer decisions (texel indices, CDF bins, reservoir selections) depend on
sin/cos/acos/atan2.  CUDA libdevice and glibc disagree in the last ulp, so both the CUDA kernels
and the CPU oracle evaluate the SAME double-precision polynomials (mul/add only, no FMA
contraction) and round once to fp32.  This script fits them (Chebyshev interpolation on the
reduced argument) and prints C initialisers; the numbers in the header were pasted from its output.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

def fit(f, lo, hi, deg):
    # interpolate at Chebyshev nodes of [lo,hi]; return monomial coefficients in z
    k = np.arange(deg + 1)
    x = np.cos(np.pi * (k + 0.5) / (deg + 1))
    z = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
    c = C.chebfit(x, f(z), deg)
    # convert cheb series in x to power series in z
    px = C.cheb2poly(c)
    # x = (2z - (hi+lo))/(hi-lo)
    a = 2.0 / (hi - lo); b = -(hi + lo) / (hi - lo)
    out = np.zeros(1)
    base = np.ones(1)
    for ci in px:
        out = P.polyadd(out, ci * base)
        base = P.polymul(base, np.array([b, a]))
    return out

def sinc(z):
    r = np.sqrt(z)
    return np.where(z > 0, np.sin(r) / np.where(r == 0, 1, r), 1.0)
def cosz(z):
    return np.cos(np.sqrt(z))
def asinc(z):
    r = np.sqrt(z)
    return np.where(z > 0, np.arcsin(r) / np.where(r == 0, 1, r), 1.0)
def atanc(z):
    r = np.sqrt(z)
    return np.where(z > 0, np.arctan(r) / np.where(r == 0, 1, r), 1.0)

def emit(name, c):
    print("static const double %s[%d] = {" % (name, len(c)))
    for v in c:
        print("    %.17e," % v)
    print("};")

if __name__ == "__main__":
    q = (np.pi / 4) ** 2 * 1.02
    S = fit(sinc, 0.0, q, 5); emit("MR_SIN_C", S)
    Cc = fit(cosz, 0.0, q, 6); emit("MR_COS_C", Cc)
    A = fit(asinc, 0.0, 0.2501, 9); emit("MR_ASIN_C", A)
    t = np.tan(np.pi / 8) ** 2 * 1.01
    T = fit(atanc, 0.0, t, 8); emit("MR_ATAN_C", T)
    # accuracy report
    r = np.linspace(-np.pi / 4, np.pi / 4, 200001)
    z = r * r
    print("// sin err", np.max(np.abs(r * P.polyval(z, S) - np.sin(r))))
    print("// cos err", np.max(np.abs(P.polyval(z, Cc) - np.cos(r))))
    x = np.linspace(-0.5, 0.5, 200001)
    print("// asin err", np.max(np.abs(x * P.polyval(x * x, A) - np.arcsin(x))))
    x = np.linspace(-np.tan(np.pi / 8), np.tan(np.pi / 8), 200001)
    print("// atan err", np.max(np.abs(x * P.polyval(x * x, T) - np.arctan(x))))
