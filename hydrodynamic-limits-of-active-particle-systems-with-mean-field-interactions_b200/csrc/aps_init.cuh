// aps_init.cuh — device-side initial conditions for native-mode ensembles (K3 streams).
//
// Same distributions as ParticleSystem.init_particles (PARTICLE_solver_CLASS.py:141-195), sampled
// from the counter-based Philox streams of include/aps_philox.h instead of numpy's Generator
// (numpy's bit stream cannot be reproduced on the device; statistical parity is tested instead,
// and the same algorithm is restated in the oracle for bit-exact GPU<->oracle checks):
//   'poisson'  per site c± ~ Poisson(rho0±[x]) by CDF inversion; a site holding more than K keeps a
//              uniformly random K-subset in random order (:169-176); particles are emitted in site
//              order, + labels before - labels (:173).
//   'fixed'    N particles, each placed uniformly among the sites that still have room (:149-156);
//              for K = 1 this is the same law as rng.choice(L, N, replace=False) (:145).
//              sigma = +1 / -1 with probability 1/2 (:146,157).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aps.h"
#include "../../include/aps_math.h"
#include "../../include/aps_philox.h"
#include "../../include/aps_sampling.h"

namespace aps {

#if defined(__CUDACC__)
__global__ void init_kernel(aps_init_args a) {
    extern __shared__ __align__(8) unsigned char ik_raw[];
    const int rep = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int L = a.L, K = a.K, n_max = a.n_max;
    const uint32_t k0 = (uint32_t)a.seeds[rep], k1 = (uint32_t)(a.seeds[rep] >> 32);
    int32_t* gpos = a.pos0 + (size_t)rep * n_max;
    int8_t* gsig = a.sigma0 + (size_t)rep * n_max;
    if (a.mode == 1) {
        // ---- poisson ----
        uint64_t* masks = reinterpret_cast<uint64_t*>(ik_raw);          // [L]
        int32_t* cnt = reinterpret_cast<int32_t*>(masks + L);           // [L] -> exclusive offsets
        __shared__ int32_t wsum[32];
        __shared__ int32_t carry;
        const int prof = a.profile_of ? a.profile_of[rep] : 0;
        const double* rp = a.rho0_plus + (size_t)prof * L;
        const double* rm = a.rho0_minus + (size_t)prof * L;
        for (int x = tid; x < L; x += NT) { uint64_t m; cnt[x] = sample_site((uint32_t)x, rp[x], rm[x], K, k0, k1, &m); masks[x] = m; }
        if (tid == 0) carry = 0;
        __syncthreads();
        // exclusive scan of cnt[] in tiles of NT sites
        for (int base = 0; base < L; base += NT) {
            int x = base + tid, v = x < L ? cnt[x] : 0, inc = v;
            for (int o = 1; o < 32; o <<= 1) { int up = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += up; }
            if ((tid & 31) == 31) wsum[tid >> 5] = inc;
            __syncthreads();
            int off = carry;
            for (int w = 0; w < (tid >> 5); ++w) off += wsum[w];
            if (x < L) cnt[x] = off + inc - v;
            __syncthreads();
            if (tid == NT - 1) carry = off + inc;
            __syncthreads();
        }
        const int n = carry;
        if (n > n_max) { if (tid == 0) a.n[rep] = -1; return; }
        for (int x = tid; x < L; x += NT) {
            int o = cnt[x], c = ((x + 1 < L) ? cnt[x + 1] : n) - o;
            uint64_t m = masks[x];
            for (int j = 0; j < c; ++j) { gpos[o + j] = x; gsig[o + j] = ((m >> j) & 1ULL) ? 1 : -1; }
        }
        if (tid == 0) a.n[rep] = n;
    } else {
        // ---- fixed: sequential fill by one thread (N iterations, once per replica) ----
        uint16_t* avail = reinterpret_cast<uint16_t*>(ik_raw);          // [L]
        uint8_t* fill = reinterpret_cast<uint8_t*>(avail + L + (L & 1));  // [L]
        const int N = a.N_of ? a.N_of[rep] : a.N_fixed;
        if (N > n_max || (long long)N > (long long)L * K) { if (tid == 0) a.n[rep] = -1; return; }
        for (int x = tid; x < L; x += NT) { avail[x] = (uint16_t)x; fill[x] = 0; }
        __syncthreads();
        if (tid == 0) {
            int navail = L;
            for (int i = 0; i < N; ++i) {
                aps_u32x4 r4 = aps_philox4x32_10((uint32_t)i, 0u, APS_RNG_INIT_POS, 0u, k0, k1);
                int j = (int)APS_MUL(aps_u53(r4.v[0], r4.v[1]), (double)navail);
                if (j >= navail) j = navail - 1;
                int site = avail[j];
                gpos[i] = site;
                if (++fill[site] >= K) avail[j] = avail[--navail];
            }
            a.n[rep] = N;
        }
        for (int i = tid; i < N; i += NT) {
            aps_u32x4 r4 = aps_philox4x32_10((uint32_t)i, 0u, APS_RNG_INIT_SIGMA, 0u, k0, k1);
            gsig[i] = (r4.v[0] & 1u) ? 1 : -1;
        }
    }
}
#endif

}  // namespace aps
