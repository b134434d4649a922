// aps_k2.cuh — K2: sublattice-parallel kernel for lattices too large for shared memory.
// Update rule: include/aps_k2_model.h (synchronous sublattice KMC; shared with the oracle).
//
// HBM design: one byte per site, one pass = one streamed read + one streamed write of the lattice
// (2 B per site-visit).  A CTA owns a window of APS_K2_TILE sites whose borders sit in the middle of
// the inactive halves of that parity (so no update ever crosses a window border), stages it into
// shared memory with 16-byte vector loads, lets each of its 128 threads run the trials of one
// 64-site segment (shared-memory layout padded by one word per segment -> conflict-free lanes),
// and writes the window back with 16-byte vector stores to the ping-pong buffer.  The frozen copy
// used by the local magnetisation (+-r halo) is a second shared-memory array.  Random bits are
// spent per EVENT (Poisson number of trials per segment), not per site, which is what keeps the
// kernel memory-bound at small rate*dt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aps.h"
#include "../../include/aps_k2_model.h"

namespace aps {

constexpr int kK2Tile = 8192;                       // sites per CTA window
constexpr int kK2Threads = kK2Tile / APS_K2_SEG;    // one thread per segment = 128
constexpr int kK2Margin = APS_K2_HALF / 2;          // window shift: borders in the middle of inactive halves

__device__ __forceinline__ int k2_pad(int p) { return p + (p >> 6) * 4; }   // one pad word per 64-site segment

__device__ __forceinline__ long long k2_reflect(long long i, long long L) {
    long long per = 2 * L, m = i % per;
    if (m < 0) m += per;
    if (m >= L) m = per - 1 - m;
    return m;
}

// One pass.  LOCAL: Gaussian local field from the frozen snapshot; otherwise the global magnetisation.
template <bool LOCAL>
__global__ void __launch_bounds__(kK2Threads) k2_pass_kernel(const __grid_constant__ aps_k2_args a) {
    extern __shared__ __align__(16) unsigned char k2_raw[];
    const int tid = threadIdx.x;
    const long long L = a.L;                         // sites held by this call (slab incl. ghosts)
    const int qpar = (int)(a.pass & 1ULL);           // parity; slabs start on tile boundaries, so it is global
    const int r = LOCAL ? a.radius : 0;
    const long long t0 = (long long)blockIdx.x * kK2Tile;
    const int sh = qpar * APS_K2_HALF - kK2Margin;
    long long lo = t0 + sh, hi = t0 + kK2Tile + sh;
    if (blockIdx.x == 0) lo = 0;
    if (blockIdx.x == gridDim.x - 1) hi = L;
    // shared memory: work (padded, origin = t0 - 64), snap (unpadded, origin = lo - r)
    unsigned char* work = k2_raw;
    const int work_bytes = k2_pad(kK2Tile + 128) + 16;
    unsigned char* snap = k2_raw + ((work_bytes + 15) & ~15);
    const uint8_t* __restrict__ in = a.in;
    uint8_t* __restrict__ out = a.out;

    // ---- stage the window (16-byte vector loads; lo/hi are multiples of 16) ----
    for (long long i = lo + 16LL * tid; i < hi; i += 16LL * kK2Threads) {
        uint4 v = *reinterpret_cast<const uint4*>(in + i);
        int p = (int)(i - t0) + 64;
        uint32_t* w = reinterpret_cast<uint32_t*>(work + k2_pad(p));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        if (LOCAL) {
            uint32_t* s4 = reinterpret_cast<uint32_t*>(snap + (i - lo) + ((r + 3) & ~3));
            s4[0] = v.x; s4[1] = v.y; s4[2] = v.z; s4[3] = v.w;
        }
    }
    if (LOCAL) {   // +-r halo of the frozen copy (reflect only at true walls; slab ends behave like walls)
        const int ro = (r + 3) & ~3;
        for (int j = tid; j < r; j += kK2Threads) {
            snap[ro - 1 - j] = in[k2_reflect(lo - 1 - j, L)];
            snap[ro + (hi - lo) + j] = in[k2_reflect(hi + j, L)];
        }
    }
    __syncthreads();

    // ---- trials of this thread's segment ----
    const long long seg_local = (long long)blockIdx.x * kK2Threads + tid;          // segment index in this slab
    const uint64_t seg_global = (uint64_t)(a.global_offset / APS_K2_SEG) + (uint64_t)seg_local;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t c0 = (uint32_t)seg_global, c1 = (uint32_t)a.pass;
    const uint32_t chi = (uint32_t)(seg_global >> 32) * 0x9E3779B9u + APS_RNG_SUBLATTICE;   // folds high bits into word 3
    const long long abase = t0 + (long long)tid * APS_K2_SEG + qpar * APS_K2_HALF;     // first active site (slab index)
    int dsig = 0;
    if (abase + APS_K2_HALF <= L) {
        aps_u32x4 pn = aps_philox4x32_10(c0, c1, 0xFFFFFFFFu, chi, k0, k1);
        const double un = aps_u53(pn.v[0], pn.v[1]);
        int ntr = 0;
        while (ntr < APS_K2_MAX_TRIALS - 1 && un >= a.rates.cdf[ntr]) ++ntr;
        const int pbase = (int)(abase - t0) + 64;      // padded-layout position of the first active site
        const double mg = LOCAL ? 0.0 : APS_DIV((double)(*a.msum_in), (double)a.n_particles);
        for (int pair = 0; 2 * pair < ntr; ++pair) {
            aps_u32x4 w4 = aps_philox4x32_10(c0, c1, (uint32_t)pair, chi, k0, k1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (2 * pair + h >= ntr) break;
                const uint32_t wa = w4.v[2 * h], wb = w4.v[2 * h + 1];
                const int x = (int)(wa >> 27);
                const uint32_t slot = wa << 5;
                const int p = pbase + x;
                const long long lx = abase + x;                        // slab-local site index (slab ends act as walls)
                const unsigned char v = work[k2_pad(p)];
                if (v == APS_K2_EMPTY) continue;
                if (slot < a.rates.t_left) {
                    if (lx > 0 && work[k2_pad(p - 1)] == APS_K2_EMPTY) { work[k2_pad(p - 1)] = v; work[k2_pad(p)] = APS_K2_EMPTY; }
                } else if (slot < a.rates.t_right || (slot < a.rates.t_active && v == APS_K2_PLUS)) {
                    if (lx < L - 1 && work[k2_pad(p + 1)] == APS_K2_EMPTY) { work[k2_pad(p + 1)] = v; work[k2_pad(p)] = APS_K2_EMPTY; }
                } else if (slot >= a.rates.t_active) {
                    const int sg = (v == APS_K2_PLUS) ? 1 : -1;
                    double m;
                    if (LOCAL) {
                        const unsigned char* c = snap + ((r + 3) & ~3) + (abase + x - lo);
                        int sw, tw;
                        { const int cv = c[0]; const int wj = a.w16[0]; sw = wj * ((cv == APS_K2_PLUS) - (cv == APS_K2_MINUS)); tw = wj * (cv != 0); }
                        for (int j = 1; j <= r; ++j) {
                            const int cl = c[-j], cr = c[j], wj = a.w16[j];
                            sw += wj * (((cl == APS_K2_PLUS) - (cl == APS_K2_MINUS)) + ((cr == APS_K2_PLUS) - (cr == APS_K2_MINUS)));
                            tw += wj * ((cl != 0) + (cr != 0));
                        }
                        m = tw > 0 ? APS_DIV((double)sw, (double)tw) : 0.0;
                    } else m = mg;
                    const double cflip = aps_exp(APS_MUL(APS_MUL(-a.rates.beta, (double)sg), m));
                    if (APS_MUL((double)wb, 2.3283064365386963e-10) < APS_MUL(cflip, a.rates.inv_cmax)) {
                        work[k2_pad(p)] = (v == APS_K2_PLUS) ? APS_K2_MINUS : APS_K2_PLUS;
                        dsig -= 2 * sg;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- write the window back ----
    for (long long i = lo + 16LL * tid; i < hi; i += 16LL * kK2Threads) {
        int p = (int)(i - t0) + 64;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(work + k2_pad(p));
        *reinterpret_cast<uint4*>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (a.msum_out) {   // running sum(sigma) for the global magnetisation of the next pass
        for (int o = 16; o > 0; o >>= 1) dsig += __shfl_xor_sync(0xffffffffu, dsig, o);
        if ((tid & 31) == 0 && dsig != 0) atomicAdd(reinterpret_cast<unsigned long long*>(a.msum_out), (unsigned long long)(long long)dsig);
    }
}

// Bernoulli initial condition: site occupied with probability `density`, '+' with probability frac_plus.
__global__ void k2_init_kernel(uint8_t* __restrict__ state, long long L, long long global_offset, uint64_t seed,
                               uint32_t t_occ, uint32_t t_plus) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;   // 2 sites per thread (4 words)
    if (i4 >= L) return;
    const uint64_t g = (uint64_t)(global_offset + i4);
    aps_u32x4 w = aps_philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, APS_RNG_INIT_SITE, (uint32_t)seed, (uint32_t)(seed >> 32));
    state[i4] = (w.v[0] < t_occ) ? ((w.v[1] < t_plus) ? APS_K2_PLUS : APS_K2_MINUS) : APS_K2_EMPTY;
    if (i4 + 1 < L) state[i4 + 1] = (w.v[2] < t_occ) ? ((w.v[3] < t_plus) ? APS_K2_PLUS : APS_K2_MINUS) : APS_K2_EMPTY;
}

// Coarse-grained profile: counts of '+' and '-' per bin (bin = site * nbins / L_global), int64 accumulators.
__global__ void k2_profile_kernel(const uint8_t* __restrict__ state, long long L, long long global_offset, long long L_global,
                                  int nbins, unsigned long long* __restrict__ cnt_plus, unsigned long long* __restrict__ cnt_minus) {
    const long long chunk = 4096;
    const long long start = (long long)blockIdx.x * chunk;
    if (start >= L) return;
    const long long end = start + chunk < L ? start + chunk : L;
    // a 4096-site chunk spans at most two bins when L_global/nbins >= 4096; handle the general case per thread
    int cp = 0, cm = 0; long long cur = -1;
    for (long long i = start + threadIdx.x; i < end; i += blockDim.x) {
        const uint8_t v = state[i];
        const long long b = (long long)(((__int128)(global_offset + i) * nbins) / L_global);
        if (b != cur) {
            if (cur >= 0) { if (cp) atomicAdd(&cnt_plus[cur], (unsigned long long)cp); if (cm) atomicAdd(&cnt_minus[cur], (unsigned long long)cm); }
            cur = b; cp = 0; cm = 0;
        }
        cp += (v == APS_K2_PLUS); cm += (v == APS_K2_MINUS);
    }
    if (cur >= 0) { if (cp) atomicAdd(&cnt_plus[cur], (unsigned long long)cp); if (cm) atomicAdd(&cnt_minus[cur], (unsigned long long)cm); }
}

}  // namespace aps
