/* aps_philox.h — Philox4x32-10 counter-based generator (Salmon et al., SC'11), shared by the
 * CUDA kernels and the CPU oracle.  It replaces the reference's numpy Generator stream
 * (PARTICLE_solver_CLASS.py:75-78,358-362,378) in native mode.
 *
 * Stream layout used by the particle stepper (one independent stream per replica):
 *   key     = (seed_lo, seed_hi)              -- 64-bit per-replica seed
 *   counter = (event_lo, event_hi, purpose, lane)
 *   purpose APS_RNG_EVENT_A: words 0,1 -> uniform for the waiting time (tau = -log(1-u)/R)
 *                            words 2,3 -> uniform for the particle choice   (draw #2, :360)
 *   purpose APS_RNG_EVENT_B: words 0,1 -> uniform for the event type        (draw #3, :362)
 *                            words 2,3 -> uniform for the hop direction     (draw #4, :378)
 *   purpose APS_RNG_INIT_*  : initial-condition sampling (counter word 0 = site / particle).
 * A 53-bit uniform in [0,1) is ((w0<<32 | w1) >> 11) * 2^-53.
 */
#ifndef APS_PHILOX_H
#define APS_PHILOX_H

#include <stdint.h>

#if defined(__CUDACC__)
#define APS_PHD __host__ __device__ __forceinline__
#else
#define APS_PHD static inline
#endif

#define APS_RNG_EVENT_A 0u
#define APS_RNG_EVENT_B 1u
#define APS_RNG_INIT_SITE 2u
#define APS_RNG_INIT_TRUNC 3u
#define APS_RNG_INIT_POS 4u
#define APS_RNG_INIT_SIGMA 5u
#define APS_RNG_SUBLATTICE 6u

typedef struct { uint32_t v[4]; } aps_u32x4;

APS_PHD void aps_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
#endif
}

APS_PHD aps_u32x4 aps_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        aps_mulhilo32(M0, c0, &hi0, &lo0);
        aps_mulhilo32(M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    aps_u32x4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

APS_PHD double aps_u53(uint32_t w0, uint32_t w1) {
    uint64_t x = (((uint64_t)w0 << 32) | (uint64_t)w1) >> 11;
    return (double)x * 1.1102230246251565404e-16; /* 2^-53, exact */
}

#endif /* APS_PHILOX_H */
