/* aps_k2_model.h — the sublattice-parallel update rule of K2, written once and shared by the CUDA
 * kernel (csrc/aps_k2.cuh) and the CPU oracle so that both are bit-identical by construction.
 *
 * K2 is NOT a restatement of a reference function: the reference's Gillespie chain is serial and
 * costs O(L) per event, so it cannot run a lattice of 2^26 sites at all.  K2 is
 * the discrete-time, synchronous-sublattice version of the same model (same rates: diffusive hops D
 * to either free neighbour, active hop lambda to the right for sigma=+1, flip exp(-beta*sigma*m),
 * reflecting walls, K = 1; PARTICLE_solver_CLASS.py:276-319,59-60,438-446):
 *   - the lattice is cut into segments of APS_K2_SEG sites; in the pass of parity q only the half
 *     [q*H, q*H+H) of every segment is ACTIVE (H = SEG/2), so active regions are separated by
 *     inactive halves and can be updated concurrently without conflicts (hops reach +-1 site);
 *   - during a pass every active half runs continuous-time kinetic Monte Carlo for a time dt by
 *     uniformisation: n ~ Poisson(B*H*dt) trials, each picks a site of the half uniformly and one
 *     of the rate slots [0,D) hop left, [D,2D) hop right, [2D,2D+lambda) active hop (sigma=+1),
 *     [2D+lambda, B) flip accepted with probability exp(-beta*sigma*m)/exp(beta), B = 2D+lambda+exp(beta);
 *   - the magnetisation m is frozen at the start of the pass: global (sum sigma / N) or local with
 *     integer (2^-16 fixed-point) Gaussian taps over +-r sites with reflect padding at the walls, the
 *     local value being quantised to 1/512 so that the acceptance threshold is a table lookup;
 *   - every decision is an integer comparison of a 32-bit Philox word with a precomputed threshold;
 *   - two passes (q = 0, 1) advance every particle by dt.
 * Site encoding (one byte per site): 0 empty, 1 = '+', 2 = '-'.
 * Random numbers: Philox4x32-10, key = seed, counter = (segment, pass, call, APS_RNG_SUBLATTICE).  ONE 32-bit word per
 * trial: call 0 -> word 0: Poisson trial count, words 1..3: trials 0..2; call c >= 1 -> words 0..3: trials 4c-1 .. 4c+2.
 * In a trial word the top 5 bits give the site, the remaining 27 bits (shifted up) the rate slot; a flip trial is
 * accepted iff its position inside the flip slot, slot - t_active in [0, 2^32 - t_active), is below
 * exp(-beta*sigma*m)/exp(|beta|) * (2^32 - t_active)  — the same uniform variate decides slot and acceptance, which is
 * exact (conditional on landing in the flip slot the position is uniform there) and halves the Philox work per trial.
 */
#ifndef APS_K2_MODEL_H
#define APS_K2_MODEL_H

#include "aps_math.h"
#include "aps_philox.h"

#define APS_K2_SEG 64
#define APS_K2_HALF 32
#define APS_K2_EMPTY 0
#define APS_K2_PLUS 1
#define APS_K2_MINUS 2
#define APS_K2_MAX_TRIALS 64

#define APS_K2_MQ 512                   /* local field is quantised to multiples of 1/512 for the flip table */

typedef struct aps_k2_rates {
    uint32_t t_left, t_right, t_active; /* cumulative 32-bit thresholds of the rate slots           */
    uint32_t n_cdf;                     /* entries of cdf32 in use                                  */
    double inv_cmax;                    /* exp(-|beta|)                                             */
    double beta;
    double mu;                          /* B*H*dt, mean trials per active half per pass             */
    uint32_t cdf32[APS_K2_MAX_TRIALS];  /* floor(2^32 * Poisson(mu) cdf): n = #{k : w >= cdf32[k]}   */
} aps_k2_rates;

/* acceptance threshold of a flip with rate c = exp(-beta*sigma*m): accept iff (slot - t_active) < thr */
APS_HD uint32_t aps_k2_flip_thr(double beta, int sigma, double m, double inv_cmax, uint32_t t_active) {
    double c = aps_exp(APS_MUL(APS_MUL(-beta, (double)sigma), m));
    double v = APS_MUL(APS_MUL(c, inv_cmax), APS_SUB(4294967296.0, (double)t_active));
    return v >= 4294967295.0 ? 4294967295u : (uint32_t)v;
}
/* Philox call and word (0..3) that hold trial `tr` of a segment */
APS_HD uint32_t aps_k2_trial_call(int tr) { return tr < 3 ? 0u : 1u + (uint32_t)((tr - 3) >> 2); }
APS_HD int aps_k2_trial_word(int tr) { return tr < 3 ? tr + 1 : ((tr - 3) & 3); }
/* quantised local field index in [0, 2*MQ]: round(MQ * sw / tw) + MQ (half away from zero), integer only */
APS_HD int aps_k2_mq_index(int sw, int tw) {
    if (tw <= 0) return APS_K2_MQ;
    int num = sw * APS_K2_MQ;                       /* |sw| <= sum of taps ~ 2^16 -> fits in 32 bits */
#if defined(__CUDA_ARCH__)
    /* same truncated quotient as the C division below, without the ~50-instruction integer-division sequence: fp32
     * estimate (|quotient| <= ~1024, so it is within +-1) corrected by the exact remainder */
    const int an = (num >= 0 ? num : -num) + tw / 2;
    int q = (int)(__int2float_rz(an) * __frcp_rn(__int2float_rn(tw)));
    int rem = an - q * tw;
    if (rem < 0) { --q; rem += tw; }
    if (rem >= tw) ++q;
    if (num < 0) q = -q;
#else
    int q = (num >= 0 ? num + tw / 2 : num - tw / 2) / tw;
#endif
    if (q < -APS_K2_MQ) q = -APS_K2_MQ;
    if (q > APS_K2_MQ) q = APS_K2_MQ;
    return (int)q + APS_K2_MQ;
}

/* Host-side (and oracle) construction of the thresholds; plain double arithmetic, done once. */
static inline int aps_k2_make_rates(double D, double lam, double beta, double dt, aps_k2_rates* r) {
    double cmax = aps_exp(beta < 0 ? -beta : beta);
    double B = 2.0 * D + lam + cmax;
    double two32 = 4294967296.0;
    double a = D / B * two32, b = 2.0 * D / B * two32, c = (2.0 * D + lam) / B * two32;
    r->t_left = (uint32_t)(a < 4294967295.0 ? a : 4294967295.0);
    r->t_right = (uint32_t)(b < 4294967295.0 ? b : 4294967295.0);
    r->t_active = (uint32_t)(c < 4294967295.0 ? c : 4294967295.0);
    r->inv_cmax = 1.0 / cmax;
    r->beta = beta;
    r->mu = B * (double)APS_K2_HALF * dt;
    if (!(r->mu > 0.0) || r->mu > 24.0) return -1;   /* keep the truncated Poisson tail < 1e-10 */
    double p = aps_exp(-r->mu), F = p;
    r->n_cdf = 0;
    for (int k = 0; k < APS_K2_MAX_TRIALS; ++k) {
        double v = F * two32;
        r->cdf32[k] = v >= 4294967295.0 ? 4294967295u : (uint32_t)v;
        if (r->cdf32[k] < 4294967295u) r->n_cdf = (uint32_t)(k + 1);
        p = p * r->mu / (double)(k + 1);
        F += p;
    }
    if (r->n_cdf >= APS_K2_MAX_TRIALS) r->n_cdf = APS_K2_MAX_TRIALS - 1;
    return 0;
}

#endif /* APS_K2_MODEL_H */
