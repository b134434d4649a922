/* aps_math.h — deterministic fp64 exp/log shared by the CUDA kernels and the CPU oracle.
 *
 * Why this exists: the reference computes flip rates with numpy's `np.exp`
 * (PARTICLE_solver_CLASS.py:60) whose last-ulp behaviour depends on the numpy build and the
 * CPU's SIMD dispatch, and CUDA's exp() is a third implementation.  To make "GPU == oracle"
 * a bit-exact statement for every output (including event times), both sides evaluate exp/log
 * through the SAME sequence of IEEE-754 operations written below: only +, -, *, / on doubles
 * (each correctly rounded on x86-64 SSE2 and on sm_100a), no FMA contraction, no libm.
 * log: the classic fdlibm algorithm; exp: table-driven and division-free (see aps_exp); error <= 1 ulp,
 * checked against libm in tests/test_oracle_units.py.
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

/* exp(x) for finite x; saturates to 0 / +inf outside the double range.  fdlibm formulation (one division); since
 * round 2 only the out-of-range / NaN path of aps_exp below uses it. */
APS_HD double aps_exp_fdlibm(double x) {
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

/* 2^(j/64), j = 0..63, correctly rounded (generated with 60-digit decimal arithmetic) */
#define APS_EXP_TAB_BODY \
    0x3ff0000000000000ULL, 0x3ff02c9a3e778061ULL, 0x3ff059b0d3158574ULL, 0x3ff0874518759bc8ULL, \
    0x3ff0b5586cf9890fULL, 0x3ff0e3ec32d3d1a2ULL, 0x3ff11301d0125b51ULL, 0x3ff1429aaea92de0ULL, \
    0x3ff172b83c7d517bULL, 0x3ff1a35beb6fcb75ULL, 0x3ff1d4873168b9aaULL, 0x3ff2063b88628cd6ULL, \
    0x3ff2387a6e756238ULL, 0x3ff26b4565e27cddULL, 0x3ff29e9df51fdee1ULL, 0x3ff2d285a6e4030bULL, \
    0x3ff306fe0a31b715ULL, 0x3ff33c08b26416ffULL, 0x3ff371a7373aa9cbULL, 0x3ff3a7db34e59ff7ULL, \
    0x3ff3dea64c123422ULL, 0x3ff4160a21f72e2aULL, 0x3ff44e086061892dULL, 0x3ff486a2b5c13cd0ULL, \
    0x3ff4bfdad5362a27ULL, 0x3ff4f9b2769d2ca7ULL, 0x3ff5342b569d4f82ULL, 0x3ff56f4736b527daULL, \
    0x3ff5ab07dd485429ULL, 0x3ff5e76f15ad2148ULL, 0x3ff6247eb03a5585ULL, 0x3ff6623882552225ULL, \
    0x3ff6a09e667f3bcdULL, 0x3ff6dfb23c651a2fULL, 0x3ff71f75e8ec5f74ULL, 0x3ff75feb564267c9ULL, \
    0x3ff7a11473eb0187ULL, 0x3ff7e2f336cf4e62ULL, 0x3ff82589994cce13ULL, 0x3ff868d99b4492edULL, \
    0x3ff8ace5422aa0dbULL, 0x3ff8f1ae99157736ULL, 0x3ff93737b0cdc5e5ULL, 0x3ff97d829fde4e50ULL, \
    0x3ff9c49182a3f090ULL, 0x3ffa0c667b5de565ULL, 0x3ffa5503b23e255dULL, 0x3ffa9e6b5579fdbfULL, \
    0x3ffae89f995ad3adULL, 0x3ffb33a2b84f15fbULL, 0x3ffb7f76f2fb5e47ULL, 0x3ffbcc1e904bc1d2ULL, \
    0x3ffc199bdd85529cULL, 0x3ffc67f12e57d14bULL, 0x3ffcb720dcef9069ULL, 0x3ffd072d4a07897cULL, \
    0x3ffd5818dcfba487ULL, 0x3ffda9e603db3285ULL, 0x3ffdfc97337b9b5fULL, 0x3ffe502ee78b3ff6ULL, \
    0x3ffea4afa2a490daULL, 0x3ffefa1bee615a27ULL, 0x3fff50765b6e4540ULL, 0x3fffa7c1819e90d8ULL,

static const unsigned long long aps_exp_tab_h[64] = {
APS_EXP_TAB_BODY
};
#if defined(__CUDACC__)
static __device__ const unsigned long long aps_exp_tab_d[64] = {
APS_EXP_TAB_BODY
};
#endif

/* exp(x), division-free (round 2: the rate refresh of K1 evaluates one exp per particle in the update window, and
 * the fdlibm form spends ~40 instructions and ~200 cycles of latency in its division):
 *   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128,   exp(x) = 2^m * T_j * (1 + r + r^2 (1/2 + r/6 + r^2/24 + r^3/120 + r^4/720))
 * with k = 64 m + j rounded through the 1.5*2^52 shift, the reduction in two exact-product steps (the high part of
 * ln2/64 has 24 trailing zero bits), T_j from the table above (read through the L1 on the device) and the scaling
 * applied to the exponent field.  Only correctly rounded +, -, * — the same operation sequence on the GPU and in the
 * oracle.  Error <= 1 ulp (truncation r^7/5040 < 3e-20; table and final rounding 0.5 ulp each); checked against libm
 * in tests/test_oracle_units.py.  |x| > 700 and NaN go through the fdlibm form. */
#define APS_EXP_INV 92.33248261689366                    /* 64 / ln2 */
#define APS_EXP_C_HI 0.010830424667801708                /* 0x3f862e42fe000000: ln2/64, 24 trailing zero bits */
#define APS_EXP_C_LO 2.8447437476627285e-11              /* 0x3dbf473de6af278f */
#define APS_EXP_P6 1.3888888888888889e-03                /* 1/720 */
#define APS_EXP_P5 8.3333333333333332e-03                /* 1/120 */
#define APS_EXP_P4 4.1666666666666664e-02                /* 1/24 */
#define APS_EXP_P3 1.6666666666666666e-01                /* 1/6 */
#if defined(__CUDACC__)
/* the same literals as constant-bank operands of the DMUL / DADD (as immediates each costs two UMOVs per use) */
static __constant__ double aps_exp_cst_d[8] = {APS_EXP_INV, APS_EXP_C_HI, APS_EXP_C_LO, APS_EXP_P6, APS_EXP_P5, APS_EXP_P4, APS_EXP_P3, 0.0};
#endif
APS_HD double aps_exp(double x) {
#if defined(__CUDA_ARCH__)
    const double INV = aps_exp_cst_d[0], C_HI = aps_exp_cst_d[1], C_LO = aps_exp_cst_d[2];
    const double P6 = aps_exp_cst_d[3], P5 = aps_exp_cst_d[4], P4 = aps_exp_cst_d[5], P3 = aps_exp_cst_d[6];
#else
    const double INV = APS_EXP_INV, C_HI = APS_EXP_C_HI, C_LO = APS_EXP_C_LO;
    const double P6 = APS_EXP_P6, P5 = APS_EXP_P5, P4 = APS_EXP_P4, P3 = APS_EXP_P3;
#endif
    const double SHIFT = 6755399441055744.0;            /* 1.5 * 2^52 */
    if (!(x >= -700.0 && x <= 700.0)) return aps_exp_fdlibm(x);
    const double t = APS_ADD(APS_MUL(x, INV), SHIFT);
    const double kd = APS_SUB(t, SHIFT);
    const int32_t ki = (int32_t)(uint32_t)aps_d2u(t);   /* low word of the shifted value = k (two's complement) */
    double r = APS_SUB(x, APS_MUL(kd, C_HI));
    r = APS_SUB(r, APS_MUL(kd, C_LO));
    const int j = ki & 63, m = ki >> 6;
#if defined(__CUDA_ARCH__)
    const double T = __longlong_as_double((long long)__ldg(&aps_exp_tab_d[j]));
#else
    const double T = aps_u2d(aps_exp_tab_h[j]);
#endif
    const double r2 = APS_MUL(r, r);
    double q = APS_ADD(P5, APS_MUL(r, P6));              /* 1/120 + r/720 */
    q = APS_ADD(P4, APS_MUL(r, q));                      /* 1/24 */
    q = APS_ADD(P3, APS_MUL(r, q));                      /* 1/6 */
    q = APS_ADD(0.5, APS_MUL(r, q));
    const double p = APS_ADD(r, APS_MUL(r2, q));
    const double y = APS_ADD(T, APS_MUL(T, p));          /* in [0.99, 2.01) */
    return aps_u2d(aps_d2u(y) + (((uint64_t)(int64_t)m) << 52));
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

/* Total rate R in NATIVE (Philox) mode — round 2.
 * In replay mode R = rates.sum() has to be numpy's pairwise sum bit for bit (CLASS.py:352): it feeds the event clock, and the clock
 * decides which observation row a state lands in, which is compared with the reference.  In native mode there is no reference clock to
 * compare with (different random stream), and the exact pairwise tree was 19 % of the per-event dependency chain of K1 (ncu, profiles/
 * r2_k1.md) for a quantity that differs from any other summation order by ~1e-16 relative.  Native mode therefore DEFINES R as the total
 * the particle-selection scan produces anyway:
 *     chunks of CS(n) = 16 * 2^k rates, k the smallest with at most 32 chunks (rates beyond n count as 0.0);
 *     T16(b) = adjacent pairwise tree over 16 rates: (((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))) + (((r8+r9)+...)+...)  [aps_tree16];
 *     c_j = T16 of the chunk's first 16 rates, plus, left to right, the T16 of every further block of 16 that starts below n  [aps_native_chunk];
 *     R = S(31, 32) with S(l, 1) = c_l (0.0 for absent chunks), S(l, 2w) = S(l, w) + S(l - w, w)   [32-lane Hillis-Steele scan, last lane].
 * (The tree replaced a serial left-to-right chunk sum: one lane re-sums a dirty chunk with 4 dependent additions instead of 16.)
 * Every K1 kernel and the oracle (mode 1) use this definition, so "GPU == oracle" stays bit-exact in native mode, clock included. */
APS_HD int aps_native_cs_shift(int n) {
    int s = 4;
    while (((n + (1 << s) - 1) >> s) > 32) ++s;
    return s;
}
APS_HD double aps_tree16(const double* r, int lo, int n) {
    double a[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k) a[k] = (lo + k < n) ? r[lo + k] : 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 1; w < 16; w <<= 1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 16; k += 2 * w) a[k] = APS_ADD(a[k], a[k + w]);
    }
    return a[0];
}
APS_HD double aps_native_chunk(const double* r, int lo, int cs, int n) {
    double c = aps_tree16(r, lo, n);
    for (int b = lo + 16; b < lo + cs && b < n; b += 16) c = APS_ADD(c, aps_tree16(r, b, n));
    return c;
}
#if !defined(__CUDA_ARCH__)
static inline double aps_native_total(const double* rates, int n) {
    const int sh = aps_native_cs_shift(n), cs = 1 << sh;
    double v[32];
    for (int j = 0; j < 32; ++j) v[j] = ((j << sh) < n) ? aps_native_chunk(rates, j << sh, cs, n) : 0.0;
    for (int o = 1; o < 32; o <<= 1)                 /* in-place Hillis-Steele: high lanes first so that v[l - o] is still the old value */
        for (int l = 31; l >= o; --l) v[l] = v[l] + v[l - o];
    return v[31];
}
#endif

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
