// Portable fp64 transcendentals for the PARITY build (-DCILQR_PARITY, libcilqr_b200_parity.so).
//
// Why this exists.  The reference algorithm amplifies last-bit differences into different line-search
// decisions ("Parity"), so "the GPU solve matches the CPU solve" can only be shown per
// instance if both sides execute the same IEEE operations.  Add, multiply, divide, sqrt and fma are
// correctly rounded on both x86-64 and sm_100a; sin / cos / tan / atan / exp / hypot are not the same
// functions in glibc and in CUDA's libdevice.  The functions below are written with nothing but
// those correctly rounded operations and integer bit manipulation, so gcc (-ffp-contract=off) and
// nvcc (-fmad=false) produce identical bits from them.  The parity build of the CUDA library
// and the "pm" flavour of the tests' CPU restatement both take their transcendentals from here; the
// default (fast) build keeps CUDA's libdevice.
//
// Accuracy class: that of a good libm (sin, cos, exp, atan, hypot < 1 ulp, tan < 1 ulp); held to glibc
// on the CPU by tests/test_pmath_cpu.py.  Algorithms are the classical ones (Cody-Waite three-part
// pi/2 reduction, minimax kernels on [-pi/4, pi/4], four-interval atan); arguments beyond ~1e9 lose
// accuracy (the reduction is not Payne-Hanek) but stay deterministic and identical on both sides.
#pragma once

#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define CILQR_PM_FN __host__ __device__ __forceinline__
#else
#define CILQR_PM_FN inline
#endif
#include <math.h>
#include <string.h>

namespace cilqr_pm {

CILQR_PM_FN double pm_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return ::fma(a, b, c);  // correctly rounded: hardware FMA or glibc's exact software path
#endif
}
CILQR_PM_FN double pm_sqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return ::sqrt(a);
#endif
}
CILQR_PM_FN int64_t pm_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t i;
    memcpy(&i, &x, 8);
    return i;
#endif
}
CILQR_PM_FN double pm_from_bits(int64_t i) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(i);
#else
    double x;
    memcpy(&x, &i, 8);
    return x;
#endif
}
CILQR_PM_FN double pm_abs(double x) { return pm_from_bits(pm_bits(x) & 0x7fffffffffffffffll); }

// exp: k = rint(x log2 e) by the shifter trick, r = x - k ln2 (two-part ln2, fma), degree-13 polynomial
// on |r| <= ln2/2, scaling by 2^k in two halves (overflow -> inf, underflow -> 0), NaN kept.
CILQR_PM_FN double pm_exp(double x) {
    if (!(x == x)) return x;
    const double xc = x < -1100.0 ? -1100.0 : (x > 1100.0 ? 1100.0 : x);
    const double shifter = 6755399441055744.0;  // 1.5 * 2^52
    const double t = pm_fma(xc, 1.4426950408889634, shifter);
    const double kf = t - shifter;
    const int k = int(uint32_t(uint64_t(pm_bits(t))));
    double r = pm_fma(kf, -6.93147180369123816490e-01, xc);
    r = pm_fma(kf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = pm_fma(p, r, 2.0876756987868100e-09);
    p = pm_fma(p, r, 2.5052108385441720e-08);
    p = pm_fma(p, r, 2.7557319223985888e-07);
    p = pm_fma(p, r, 2.7557319223985893e-06);
    p = pm_fma(p, r, 2.4801587301587302e-05);
    p = pm_fma(p, r, 1.9841269841269841e-04);
    p = pm_fma(p, r, 1.3888888888888889e-03);
    p = pm_fma(p, r, 8.3333333333333332e-03);
    p = pm_fma(p, r, 4.1666666666666664e-02);
    p = pm_fma(p, r, 1.6666666666666666e-01);
    p = pm_fma(p, r, 0.5);
    p = pm_fma(p, r, 1.0);
    p = pm_fma(p, r, 1.0);
    const int k1 = k >> 1, k2 = k - k1;
    const double s1 = pm_from_bits(int64_t(k1 + 1023) << 52), s2 = pm_from_bits(int64_t(k2 + 1023) << 52);
    return (p * s1) * s2;
}

// x = n pi/2 + (hi + lo), |hi + lo| <= pi/4 (+ rounding); returns n mod 4.
CILQR_PM_FN int pm_rem_pio2(double x, double* hi, double* lo) {
    const double P1 = 1.5707963267948966e+00;   // pi/2 rounded to 53 bits
    const double P2 = 6.123233995736766e-17;    // next 53 bits
    const double P3 = -1.4973849048591698e-33;  // and the next
    const double shifter = 6755399441055744.0;
    const double t = pm_fma(x, 6.36619772367581382433e-01, shifter);
    const double fn = t - shifter;
    const int n = int(uint32_t(uint64_t(pm_bits(t))));
    const double r1 = pm_fma(-fn, P1, x);  // exact for |x| < 2^20 or so
    const double h = pm_fma(-fn, P2, r1);
    double l = pm_fma(-fn, P2, r1 - h);
    l = pm_fma(-fn, P3, l);
    *hi = h;
    *lo = l;
    return n & 3;
}
// sin on [-pi/4, pi/4] of x + y (y the tail of the reduced argument)
CILQR_PM_FN double pm_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
// cos on [-pi/4, pi/4] of x + y
CILQR_PM_FN double pm_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    double w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}
CILQR_PM_FN double pm_sin(double x) {
    double h, l;
    const int n = pm_rem_pio2(x, &h, &l);
    const double s = pm_ksin(h, l), c = pm_kcos(h, l);
    const double v = (n & 1) ? c : s;
    return (n & 2) ? -v : v;
}
CILQR_PM_FN double pm_cos(double x) {
    double h, l;
    const int n = pm_rem_pio2(x, &h, &l);
    const double s = pm_ksin(h, l), c = pm_kcos(h, l);
    const double v = (n & 1) ? s : c;
    return ((n + 1) & 2) ? -v : v;
}
// tan on [-pi/4, pi/4] of x + y (iy = 1), or -1 / tan (iy = -1): odd minimax polynomial of degree 27; beyond
// 0.6744 the identity tan(x) = (1 - t) / (1 + t)-style reflection about pi/4 keeps the polynomial's argument small
CILQR_PM_FN double pm_ktan(double x, double y, int iy) {
    const double T[13] = {3.33333333333334091986e-01, 1.33333333333201242699e-01, 5.39682539762260521377e-02,
                          2.18694882948595424599e-02, 8.86323982359930005737e-03, 3.59207910759131235356e-03,
                          1.45620945432529025516e-03, 5.88041240820264096874e-04, 2.46463134818469906812e-04,
                          7.81794442939557092300e-05, 7.14072491382608190305e-05, -1.85586374855275456654e-05,
                          2.59073051863633712884e-05};
    const double pio4 = 7.85398163397448278999e-01, pio4lo = 3.06161699786838301793e-17;
    const bool neg = pm_bits(x) < 0;
    const bool big = pm_abs(x) >= 0.6744;
    if (big) {
        if (neg) {
            x = -x;
            y = -y;
        }
        const double z0 = pio4 - x, w0 = pio4lo - y;
        x = z0 + w0;
        y = 0.0;
    }
    double z = x * x;
    double w = z * z;
    double r = T[1] + w * (T[3] + w * (T[5] + w * (T[7] + w * (T[9] + w * T[11]))));
    double v = z * (T[2] + w * (T[4] + w * (T[6] + w * (T[8] + w * (T[10] + w * T[12])))));
    double s = z * x;
    r = y + z * (s * (r + v) + y);
    r += T[0] * s;
    w = x + r;
    if (big) {
        v = double(iy);
        const double t = v - 2.0 * (x - (w * w / (w + v) - r));
        return neg ? -t : t;
    }
    if (iy == 1) return w;
    // -1 / (x + r), with the head of w split off so that the quotient's error is corrected
    const int64_t hi_mask = int64_t(0xffffffff00000000ull);
    z = pm_from_bits(pm_bits(w) & hi_mask);
    v = r - (z - x);
    const double a = -1.0 / w;
    const double t = pm_from_bits(pm_bits(a) & hi_mask);
    s = 1.0 + t * z;
    return t + a * (s + t * v);
}
CILQR_PM_FN double pm_tan(double x) {
    double h, l;
    const int n = pm_rem_pio2(x, &h, &l);
    return pm_ktan(h, l, 1 - ((n & 1) << 1));
}
CILQR_PM_FN double pm_atan(double x) {
    const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                          1.57079632679489655800e+00};
    const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                          6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                           -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                           6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                           -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    if (!(x == x)) return x;
    const bool neg = pm_bits(x) < 0;
    double a = pm_abs(x);
    if (a >= 7.378697629483821e19) return neg ? -(hi[3] + lo[3]) : (hi[3] + lo[3]);  // 2^66
    int id = -1;
    if (a < 0.4375) {
        if (a < 1.862645149230957e-09) return x;  // 2^-29
    } else if (a < 1.1875) {
        if (a < 0.6875) {
            id = 0;
            a = (2.0 * a - 1.0) / (2.0 + a);
        } else {
            id = 1;
            a = (a - 1.0) / (a + 1.0);
        }
    } else if (a < 2.4375) {
        id = 2;
        a = (a - 1.5) / (1.0 + 1.5 * a);
    } else {
        id = 3;
        a = -1.0 / a;
    }
    const double z = a * a, w = z * z;
    const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) {
        const double r = a - a * (s1 + s2);
        return neg ? -r : r;
    }
    const double r = hi[id] - ((a * (s1 + s2) - lo[id]) - a);
    return neg ? -r : r;
}
// hypot: sqrt of the fused sum of squares, one Newton correction from the exact residual (fma), scaled by a
// power of two when the squares would overflow, or when both would underflow.
CILQR_PM_FN double pm_hypot(double x, double y) {
    double a = pm_abs(x), b = pm_abs(y);
    const double inf = pm_from_bits(0x7ff0000000000000ll);
    if (a == inf || b == inf) return inf;
    if (!(a == a) || !(b == b)) return a + b;
    if (a < b) {
        const double t = a;
        a = b;
        b = t;
    }
    if (a == 0.0) return 0.0;
    double scale = 1.0, unscale = 1.0;
    if (a > 1e150) {
        scale = 5.527147875260445e-181;  // 2^-600
        unscale = 1.8092513943330656e+180;
    } else if (a < 1e-150) {
        scale = 1.8092513943330656e+180;  // 2^600
        unscale = 5.527147875260445e-181;
    }
    a *= scale;
    b *= scale;
    const double xh = a * a, xl = pm_fma(a, a, -xh);
    const double yh = b * b, yl = pm_fma(b, b, -yh);
    double h = pm_sqrt(xh + yh);
    const double hh = h * h, hl = pm_fma(h, h, -hh);
    const double err = ((xh - hh) + yh) + ((xl + yl) - hl);
    h = h + err / (2.0 * h);
    return h * unscale;
}

// float flavours: evaluated in fp64 and rounded once (the same double rounding on both sides)
CILQR_PM_FN float pm_sin(float x) { return float(pm_sin(double(x))); }
CILQR_PM_FN float pm_cos(float x) { return float(pm_cos(double(x))); }
CILQR_PM_FN float pm_tan(float x) { return float(pm_tan(double(x))); }
CILQR_PM_FN float pm_atan(float x) { return float(pm_atan(double(x))); }
CILQR_PM_FN float pm_exp(float x) { return float(pm_exp(double(x))); }
CILQR_PM_FN float pm_hypot(float x, float y) { return float(pm_hypot(double(x), double(y))); }

}  // namespace cilqr_pm
