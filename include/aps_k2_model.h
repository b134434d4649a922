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
 *     integer (2^-16 fixed-point) Gaussian taps over +-r sites with reflect padding at the walls;
 *   - two passes (q = 0, 1) advance every particle by dt.
 * Site encoding (one byte per site): 0 empty, 1 = '+', 2 = '-'.
 * Random numbers: Philox4x32-10, key = seed, counter = (segment, pass, trial pair, APS_RNG_SUBLATTICE).
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

typedef struct aps_k2_rates {
    uint32_t t_left, t_right, t_active; /* cumulative 32-bit thresholds of the rate slots           */
    double inv_cmax;                    /* exp(-beta)                                               */
    double beta;
    double mu;                          /* B*H*dt, mean trials per active half per pass             */
    double cdf[APS_K2_MAX_TRIALS];      /* Poisson(mu) cdf for inversion                            */
} aps_k2_rates;

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
    for (int k = 0; k < APS_K2_MAX_TRIALS; ++k) {
        r->cdf[k] = F;
        p = p * r->mu / (double)(k + 1);
        F += p;
    }
    return 0;
}

#endif /* APS_K2_MODEL_H */
