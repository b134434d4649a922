/* aps_sampling.h — initial-condition sampling primitives shared by the CUDA init kernel
 * (csrc/aps_init.cuh) and the CPU oracle, so that both draw bit-identical initial states from the
 * Philox streams of aps_philox.h.  Distributions follow ParticleSystem._init_poisson
 * (PARTICLE_solver_CLASS.py:160-189). */
#ifndef APS_SAMPLING_H
#define APS_SAMPLING_H
#include "aps_math.h"
#include "aps_philox.h"

/* Poisson(lam) by CDF inversion of one uniform. */
APS_HD int poisson_inv(double u, double lam) {
    if (!(lam > 0.0)) return 0;
    double p = aps_exp(-lam), F = p;
    int k = 0;
    while (u > F && k < 1000) {
        ++k;
        p = APS_DIV(APS_MUL(p, lam), (double)k);
        F = APS_ADD(F, p);
    }
    return k;
}

// Site sample shared by the kernel and (through the same header) the oracle: returns the number of
// particles kept at site x and a bit mask of their labels in emission order (bit j set = '+').
APS_HD int sample_site(uint32_t x, double lam_p, double lam_m, int K, uint32_t k0, uint32_t k1, uint64_t* mask) {
    aps_u32x4 a = aps_philox4x32_10(x, 0u, APS_RNG_INIT_SITE, 0u, k0, k1);
    int cp = poisson_inv(aps_u53(a.v[0], a.v[1]), lam_p);
    int cm = poisson_inv(aps_u53(a.v[2], a.v[3]), lam_m);
    if (cp + cm <= K) {
        *mask = cp >= 64 ? ~0ULL : ((1ULL << cp) - 1ULL);
        return cp + cm;
    }
    uint64_t m = 0;
    for (int j = 0; j < K; ++j) {   // sequential draws without replacement == random K-subset in random order
        aps_u32x4 b = aps_philox4x32_10(x, (uint32_t)j, APS_RNG_INIT_TRUNC, 0u, k0, k1);
        double u = aps_u53(b.v[0], b.v[1]);
        if (APS_MUL(u, (double)(cp + cm)) < (double)cp) { m |= 1ULL << j; --cp; } else --cm;
    }
    *mask = m;
    return K;
}


#endif /* APS_SAMPLING_H */
