/* aps_math.h — deterministic fp64 exp/log shared by the CUDA kernels and the CPU oracle.
 *
 * Why this exists: the reference computes flip rates with numpy's `np.exp`
 * (PARTICLE_solver_CLASS.py:60) whose last-ulp behaviour depends on the numpy build and the
 * CPU's SIMD dispatch, and CUDA's exp() is a third implementation.  To make "GPU == oracle"
 * a bit-exact statement for every output (including event times), both sides evaluate exp/log
 * through the SAME sequence of IEEE-754 operations written below: only +, -, *, / on doubles
 * (each correctly rounded on x86-64 SSE2 and on sm_100a), no FMA contraction, no libm.
 * The algorithms are the classic fdlibm ones (argument reduction by ln2 hi/lo split plus a
 * degree-5 / degree-7 minimax polynomial); error < 1 ulp, checked against libm in
 * tests/test_math.py.
 *
 * Build rules: host code must be compiled with -ffp-contract=off; device code goes through
 * the __d*_rn intrinsics, which the compiler never fuses.
 */
#ifndef APS_MATH_H
#define APS_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define APS_HD __host__ __device__ __forceinline__
#else
#define APS_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define APS_MUL(a, b) __dmul_rn((a), (b))
#define APS_ADD(a, b) __dadd_rn((a), (b))
#define APS_SUB(a, b) __dsub_rn((a), (b))
#define APS_DIV(a, b) __ddiv_rn((a), (b))
#else
#define APS_MUL(a, b) ((a) * (b))
#define APS_ADD(a, b) ((a) + (b))
#define APS_SUB(a, b) ((a) - (b))
#define APS_DIV(a, b) ((a) / (b))
#endif

APS_HD uint64_t aps_d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
APS_HD double aps_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

/* exp(x) for finite x; saturates to 0 / +inf outside the double range. */
APS_HD double aps_exp(double x) {
    const double ln2HI = 6.93147180369123816490e-01; /* 0x3fe62e42fee00000 */
    const double ln2LO = 1.90821492927058770002e-10; /* 0x3dea39ef35793c76 */
    const double invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;

    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return aps_u2d(0x7ff0000000000000ULL);
    if (x < -7.08e+02) return 0.0; /* flush the subnormal tail; never reached by beta*m */

    uint32_t hx = (uint32_t)(aps_d2u(x) >> 32) & 0x7fffffffu;
    double hi = 0.0, lo = 0.0;
    int k = 0;
    if (hx > 0x3fd62e42u) { /* |x| > 0.5 ln2 */
        double half = (x < 0.0) ? -0.5 : 0.5;
        k = (int)APS_ADD(APS_MUL(invln2, x), half);
        double t = (double)k;
        hi = APS_SUB(x, APS_MUL(t, ln2HI)); /* t*ln2HI exact: ln2HI has 21 trailing zero bits */
        lo = APS_MUL(t, ln2LO);
        x = APS_SUB(hi, lo);
    } else if (hx < 0x3e300000u) { /* |x| < 2^-28 */
        return APS_ADD(1.0, x);
    }
    double t = APS_MUL(x, x);
    double p = APS_ADD(P4, APS_MUL(t, P5));
    p = APS_ADD(P3, APS_MUL(t, p));
    p = APS_ADD(P2, APS_MUL(t, p));
    p = APS_ADD(P1, APS_MUL(t, p));
    double c = APS_SUB(x, APS_MUL(t, p));
    if (k == 0) {
        return APS_SUB(1.0, APS_SUB(APS_DIV(APS_MUL(x, c), APS_SUB(c, 2.0)), x));
    }
    double y = APS_SUB(1.0, APS_SUB(APS_SUB(lo, APS_DIV(APS_MUL(x, c), APS_SUB(2.0, c))), hi));
    /* scale by 2^k through the exponent field; y is in [0.5, 2), |k| <= 1024 */
    uint64_t u = aps_d2u(y);
    u += ((uint64_t)(int64_t)k) << 52;
    return aps_u2d(u);
}

/* log(x) for x > 0 (normal range). Returns -inf for 0, NaN for negatives. */
APS_HD double aps_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01;
    const double Lg2 = 3.999999999940941908e-01;
    const double Lg3 = 2.857142874366239149e-01;
    const double Lg4 = 2.222219843214978396e-01;
    const double Lg5 = 1.818357216161805012e-01;
    const double Lg6 = 1.531383769920937332e-01;
    const double Lg7 = 1.479819860511658591e-01;

    uint64_t ux = aps_d2u(x);
    int32_t hx = (int32_t)(ux >> 32);
    uint32_t lx = (uint32_t)ux;
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return aps_u2d(0xfff0000000000000ULL);
        if (hx < 0) return aps_u2d(0x7ff8000000000000ULL);
        k -= 54;
        x = APS_MUL(x, 1.80143985094819840000e+16);
        ux = aps_d2u(x);
        hx = (int32_t)(ux >> 32);
        lx = (uint32_t)ux;
    }
    if (hx >= 0x7ff00000) return APS_ADD(x, x);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    x = aps_u2d(((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx);
    k += (i >> 20);
    double f = APS_SUB(x, 1.0);
    double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return APS_ADD(APS_MUL(dk, ln2_hi), APS_MUL(dk, ln2_lo));
        }
        double R = APS_MUL(APS_MUL(f, f), APS_SUB(0.5, APS_MUL(0.33333333333333333, f)));
        if (k == 0) return APS_SUB(f, R);
        return APS_SUB(APS_MUL(dk, ln2_hi), APS_SUB(APS_SUB(R, APS_MUL(dk, ln2_lo)), f));
    }
    double s = APS_DIV(f, APS_ADD(2.0, f));
    double z = APS_MUL(s, s);
    i = hx - 0x6147a;
    double w = APS_MUL(z, z);
    int32_t j = 0x6b851 - hx;
    double t1 = APS_MUL(w, APS_ADD(Lg2, APS_MUL(w, APS_ADD(Lg4, APS_MUL(w, Lg6)))));
    double t2 = APS_MUL(z, APS_ADD(Lg1, APS_MUL(w, APS_ADD(Lg3, APS_MUL(w, APS_ADD(Lg5, APS_MUL(w, Lg7)))))));
    i |= j;
    double R = APS_ADD(t2, t1);
    if (i > 0) {
        double hfsq = APS_MUL(APS_MUL(0.5, f), f);
        if (k == 0) return APS_SUB(f, APS_SUB(hfsq, APS_MUL(s, APS_ADD(hfsq, R))));
        return APS_SUB(APS_MUL(dk, ln2_hi),
                       APS_SUB(APS_SUB(hfsq, APS_ADD(APS_MUL(s, APS_ADD(hfsq, R)), APS_MUL(dk, ln2_lo))), f));
    }
    if (k == 0) return APS_SUB(f, APS_MUL(s, APS_SUB(f, R)));
    return APS_SUB(APS_MUL(dk, ln2_hi), APS_SUB(APS_SUB(APS_MUL(s, APS_SUB(f, R)), APS_MUL(dk, ln2_lo)), f));
}

/* Tabulated flip rate (custom `flip_rate_fn`, PARTICLE_solver_CLASS.py:59-62,262): tab[s*(G+1) + k] holds
 * flip_rate_fn(sigma, m_k) evaluated ON THE HOST by the caller's Python callable at m_k = -1 + 2k/G, s = 0 for
 * sigma = +1 and 1 for sigma = -1; between grid points the rate is interpolated linearly (single-rounding operations
 * only, same code in the kernels and in the oracle).  Interpolation error <= (2/G)^2 * max|d2f/dm2| / 8
 * (G = 8192: 7.5e-9 * max|f''|). */
APS_HD double aps_flip_interp(const double* tab, int64_t G, int sg, double m) {
    double x = APS_MUL(APS_ADD(m, 1.0), APS_MUL(0.5, (double)G));
    int64_t k = (int64_t)x;
    if (k < 0) k = 0;
    if (k > G - 1) k = G - 1;
    const double fr = APS_SUB(x, (double)k);
    const double* t = tab + (sg == 1 ? 0 : (G + 1)) + k;
    return APS_ADD(t[0], APS_MUL(fr, APS_SUB(t[1], t[0])));
}

#endif /* APS_MATH_H */
