/* psy_detmath.h — deterministic elementary functions for the STRICT parity mode.
 *
 * The render path calls eight libm functions: f32 sin / cos / acos / asin (monte_carlo.rs:10-32,
 * surface_closure.rs:346-377,535-591) and f64 log / sin / cos / acos (math.rs:67-132,
 * monte_carlo.rs:64-199, sphere_light.rs:102-171, distant_disk_light.rs:59-85).  glibc (the CPU
 * side) and CUDA libdevice (the GPU side) round them differently in the last place, which is enough
 * to send a few paths per million through another branch.  In strict mode both sides call the
 * functions below instead: plain IEEE + - * / sqrt in a fixed order (both builds forbid FMA
 * contraction: -ffp-contract=off / -fmad=false), so CPU and GPU produce the same bits and every
 * remaining difference between the CUDA path and the restated reference path is a real one.
 *
 * Accuracy (tests/test_detmath.py, against glibc): the f32 functions are evaluated in double and rounded
 * once, i.e. correctly rounded except in ~1e-9 of the cases; the f64 functions are within 1 ulp.
 * Domain: |x| <= 1e6 for sin / cos (the path never exceeds 4*pi), [-1, 1] for acos / asin, x > 0 for log;
 * outside they return NaN like libm (log(0) = -inf).
 *
 * Polynomial coefficients are the classic fdlibm minimax ones (public algorithm; restated, not copied
 * from any source file of the reference, which contains no libm).
 */
#ifndef PSY_DETMATH_H
#define PSY_DETMATH_H

#include <stdint.h>
#ifndef __CUDACC__
#include <string.h>
#endif

#if defined(__CUDACC__)
#define PSY_DM __host__ __device__ __forceinline__
#else
#define PSY_DM static inline
#endif

PSY_DM uint64_t psy_dm_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
PSY_DM double psy_dm_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
PSY_DM double psy_dm_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
PSY_DM double psy_dm_nan(void) { return psy_dm_from_bits(0x7ff8000000000000ull); }

/* sin / cos kernels on [-pi/4, pi/4] */
PSY_DM double psy_dm_ksin(double x) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double r = S2 + (z * (S3 + (z * (S4 + (z * (S5 + (z * S6)))))));
    return x + ((x * z) * (S1 + (z * r)));
}
PSY_DM double psy_dm_kcos(double x) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    const double r = z * (C1 + (z * (C2 + (z * (C3 + (z * (C4 + (z * (C5 + (z * C6))))))))));
    return (1.0 - (0.5 * z)) + (z * r);
}
/* x = k * pi/2 + r, |r| <= pi/4 (+ rounding); three-part Cody-Waite, exact for |k| < 2^20 */
PSY_DM double psy_dm_reduce(double x, int* quadrant) {
    const double INV_PIO2 = 6.36619772367581382433e-01;
    const double P1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double P2 = 6.07710050630396597660e-11;  /* next 33 bits */
    const double P3 = 2.02226624871116645580e-21;  /* third 33 bits */
    const double P3T = 8.47842766036889956997e-32; /* pi/2 - (P1 + P2 + P3) */
    const double y = x * INV_PIO2;
    const int k = (int)(y + (y >= 0.0 ? 0.5 : -0.5));
    const double fk = (double)k;
    double r = x - (fk * P1);
    r = r - (fk * P2);
    r = r - (fk * P3);
    r = r - (fk * P3T);
    *quadrant = k & 3;
    return r;
}
PSY_DM double psy_det_sin(double x) {
    if (!(x >= -1.0e6 && x <= 1.0e6)) return psy_dm_nan();
    int q;
    const double r = psy_dm_reduce(x, &q);
    const double s = psy_dm_ksin(r), c = psy_dm_kcos(r);
    return q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c));
}
PSY_DM double psy_det_cos(double x) {
    if (!(x >= -1.0e6 && x <= 1.0e6)) return psy_dm_nan();
    int q;
    const double r = psy_dm_reduce(x, &q);
    const double s = psy_dm_ksin(r), c = psy_dm_kcos(r);
    return q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s));
}

/* asin(x) = x + x * R(x^2) on |x| <= 0.5 */
PSY_DM double psy_dm_asin_r(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01,
                 qS4 = 7.70381505559019352791e-02;
    const double p = z * (pS0 + (z * (pS1 + (z * (pS2 + (z * (pS3 + (z * (pS4 + (z * pS5))))))))));
    const double q = 1.0 + (z * (qS1 + (z * (qS2 + (z * (qS3 + (z * qS4)))))));
    return p / q;
}
PSY_DM double psy_det_acos(double x) {
    const double PIO2_HI = 1.57079632679489655800e+00, PIO2_LO = 6.12323399573676603587e-17;
    const double PI = 3.14159265358979311600e+00;
    if (!(x >= -1.0 && x <= 1.0)) return psy_dm_nan();
    if (x == 1.0) return 0.0;
    if (x == -1.0) return PI + (2.0 * PIO2_LO);
    const double ax = x < 0.0 ? -x : x;
    if (ax < 0.5) {
        const double r = psy_dm_asin_r(x * x);
        return PIO2_HI - (x - (PIO2_LO - (x * r)));
    }
    if (x < 0.0) {
        const double z = (1.0 + x) * 0.5;
        const double s = psy_dm_sqrt(z);
        const double w = (psy_dm_asin_r(z) * s) - PIO2_LO;
        return PI - (2.0 * (s + w));
    }
    const double z = (1.0 - x) * 0.5;
    const double s = psy_dm_sqrt(z);
    const double df = psy_dm_from_bits(psy_dm_bits(s) & 0xffffffff00000000ull);
    const double c = (z - (df * df)) / (s + df);
    const double w = (psy_dm_asin_r(z) * s) + c;
    return 2.0 * (df + w);
}
PSY_DM double psy_det_asin(double x) {
    const double PIO2_HI = 1.57079632679489655800e+00, PIO2_LO = 6.12323399573676603587e-17;
    if (!(x >= -1.0 && x <= 1.0)) return psy_dm_nan();
    const double ax = x < 0.0 ? -x : x;
    if (ax < 0.5) return x + (x * psy_dm_asin_r(x * x));
    const double z = (1.0 - ax) * 0.5;
    const double s = psy_dm_sqrt(z);
    const double r = psy_dm_asin_r(z);
    const double t = PIO2_HI - ((2.0 * (s + (s * r))) - PIO2_LO);
    return x < 0.0 ? -t : t;
}

/* natural logarithm, x finite and > 0 (subnormals included) */
PSY_DM double psy_det_log(double x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    if (x != x || x < 0.0) return psy_dm_nan();
    if (x == 0.0) return psy_dm_from_bits(0xfff0000000000000ull);
    uint64_t b = psy_dm_bits(x);
    if (b >= 0x7ff0000000000000ull) return x; /* +inf */
    int k = 0;
    if (b < 0x0010000000000000ull) { /* subnormal: scale by 2^54 */
        x = x * 18014398509481984.0;
        b = psy_dm_bits(x);
        k = -54;
    }
    k += (int)(b >> 52) - 1023;
    uint64_t m = b & 0x000fffffffffffffull;
    /* mantissa in [sqrt(2)/2, sqrt(2)) */
    if (m >= 0x0006a09e667f3bcdull) {
        m |= 0x3fe0000000000000ull;
        k += 1;
    } else {
        m |= 0x3ff0000000000000ull;
    }
    const double f = psy_dm_from_bits(m) - 1.0;
    const double hfsq = (0.5 * f) * f;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + (w * (Lg4 + (w * Lg6))));
    const double t2 = z * (Lg1 + (w * (Lg3 + (w * (Lg5 + (w * Lg7))))));
    const double R = t2 + t1;
    const double dk = (double)k;
    return (dk * LN2_HI) - ((hfsq - ((s * (hfsq + R)) + (dk * LN2_LO))) - f);
}

/* the f32 functions: one double evaluation, one rounding */
PSY_DM float psy_det_sinf(float x) { return (float)psy_det_sin((double)x); }
PSY_DM float psy_det_cosf(float x) { return (float)psy_det_cos((double)x); }
PSY_DM float psy_det_acosf(float x) { return (float)psy_det_acos((double)x); }
PSY_DM float psy_det_asinf(float x) { return (float)psy_det_asin((double)x); }

#endif /* PSY_DETMATH_H */
