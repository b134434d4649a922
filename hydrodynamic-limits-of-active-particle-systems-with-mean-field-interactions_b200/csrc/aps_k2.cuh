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


__device__ __forceinline__ long long k2_reflect(long long i, long long L) {
    long long per = 2 * L, m = i % per;
    if (m < 0) m += per;
    if (m >= L) m = per - 1 - m;
    return m;
}

// ---- TMA (cp.async.bulk) + mbarrier helpers: 1-D bulk copies global <-> shared memory ----
__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k2_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void k2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void k2_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(k2_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void k2_tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(k2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(k2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k2_tma_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(k2_smem_u32(ssrc)), "r"(bytes) : "memory");
}

constexpr int kK2StagesGlobal = 4;   // global field: copy-bound, deep prefetch
constexpr int kK2StagesLocal = 2;    // local field: compute-bound, favour resident CTAs over prefetch depth

// One pass, persistent CTAs.  Each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... through a
// kK2Stages-deep ring of shared-memory buffers: TMA bulk loads (mbarrier complete_tx) bring the window
// (and, for the local field, a second read-only copy with the +-r halo) in, the 128 threads run the
// trials of their segments in place, and a TMA bulk store writes the window to the ping-pong buffer.
// LOCAL: Gaussian local field from the frozen copy; otherwise the global magnetisation.
template <bool LOCAL>
__global__ void __launch_bounds__(kK2Threads) k2_pass_kernel(const __grid_constant__ aps_k2_args a) {
    extern __shared__ __align__(128) unsigned char k2_raw[];
    const int tid = threadIdx.x;
    const long long L = a.L;                         // sites held by this call (slab incl. ghosts)
    const int qpar = (int)(a.pass & 1ULL);           // parity; slabs start on tile boundaries, so it is global
    const int r = LOCAL ? a.radius : 0;
    const int R16 = LOCAL ? ((r + 15) & ~15) : 0;
    const int sh = qpar * APS_K2_HALF - kK2Margin;
    const int ntiles = (int)(L / kK2Tile);
    constexpr int WB = kK2Tile + 32;                 // largest window
    constexpr int kK2Stages = LOCAL ? kK2StagesLocal : kK2StagesGlobal;
    const int stride = WB + (LOCAL ? WB + 2 * R16 : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(k2_raw);
    unsigned char* bufs = k2_raw + 128;
    __shared__ uint32_t thr_glob[2];
    const uint8_t* __restrict__ in = a.in;
    uint8_t* __restrict__ out = a.out;

    if (tid == 0) {
        for (int s2 = 0; s2 < kK2Stages; ++s2) k2_mbar_init(&bars[s2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!LOCAL) {
            const double m = APS_DIV((double)(*a.msum_in), (double)a.n_particles);
            thr_glob[0] = aps_k2_flip_thr(a.rates.beta, +1, m, a.rates.inv_cmax);
            thr_glob[1] = aps_k2_flip_thr(a.rates.beta, -1, m, a.rates.inv_cmax);
        }
    }
    __syncthreads();

    auto window = [&](int t, long long& lo, long long& hi) {
        lo = (long long)t * kK2Tile + sh; hi = lo + kK2Tile;
        if (t == 0) lo = 0;
        if (t == ntiles - 1) hi = L;
    };
    auto issue_load = [&](int t, int s2) {          // thread 0 only
        long long lo, hi; window(t, lo, hi);
        unsigned char* work = bufs + (size_t)s2 * stride;
        uint32_t bytes = (uint32_t)(hi - lo);
        long long slo = lo - R16, shi = hi + R16;
        if (slo < 0) slo = 0;
        if (shi > L) shi = L;
        k2_mbar_expect_tx(&bars[s2], bytes + (LOCAL ? (uint32_t)(shi - slo) : 0u));
        k2_tma_load(work, in + lo, bytes, &bars[s2]);
        if (LOCAL) k2_tma_load(work + WB + (slo - (lo - R16)), in + slo, (uint32_t)(shi - slo), &bars[s2]);
    };

    const int my_first = blockIdx.x, step_t = gridDim.x;
    if (tid == 0) {
        for (int k = 0; k < kK2Stages - 1; ++k) { int t = my_first + k * step_t; if (t < ntiles) issue_load(t, k); }
    }
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    int dsig = 0;
    int it = 0;
    for (int t = my_first; t < ntiles; t += step_t, ++it) {
        const int s2 = it % kK2Stages;
        long long lo, hi; window(t, lo, hi);
        unsigned char* work = bufs + (size_t)s2 * stride;
        unsigned char* snap = work + WB;             // snap[i - (lo - R16)] = site i (LOCAL only)
        k2_mbar_wait(&bars[s2], (uint32_t)((it / kK2Stages) & 1));
        if (LOCAL && (t == 0 || t == ntiles - 1)) {   // reflect padding at the ends of this slab
            for (int j = tid; j < r; j += kK2Threads) {
                if (t == 0) snap[R16 - 1 - j] = snap[R16 + k2_reflect(-1 - j, L)];
                if (t == ntiles - 1) snap[R16 + (L - lo) + j] = snap[R16 + (k2_reflect(L + j, L) - lo)];
            }
            __syncthreads();
        }
        // ---- trials of this thread's segment ----
        const long long t0 = (long long)t * kK2Tile;
        const long long seg_local = (long long)t * kK2Threads + tid;
        const uint64_t seg_global = (uint64_t)(a.global_offset / APS_K2_SEG) + (uint64_t)seg_local;
        const uint32_t c0 = (uint32_t)seg_global, c1 = (uint32_t)a.pass;
        const uint32_t chi = (uint32_t)(seg_global >> 32) * 0x9E3779B9u + APS_RNG_SUBLATTICE;
        const long long abase = t0 + (long long)tid * APS_K2_SEG + qpar * APS_K2_HALF;     // first active site (slab index)
        // The trial loop runs in warp lock-step up to the largest trial count of the warp so that the local
        // field of any lane can be evaluated COOPERATIVELY: the 2r+1 integer taps are spread over the 32 lanes
        // and summed with redux.sync, instead of one lane walking them while 31 wait.
        const bool seg_ok = abase + APS_K2_HALF <= L;
        aps_u32x4 w4 = aps_philox4x32_10(c0, c1, 0u, chi, k0, k1);
        int ntr = 0;
        if (seg_ok) while (ntr < (int)a.rates.n_cdf && w4.v[0] >= a.rates.cdf32[ntr]) ++ntr;
        const int ntr_max = LOCAL ? __reduce_max_sync(0xffffffffu, ntr) : ntr;
        int wreg0 = 0, wreg1 = 0;       // this lane's Gaussian taps (k = lane and lane + 32) when r <= 31
        if (LOCAL && r <= 31) {
            const int l2 = tid & 31;
            if (l2 <= 2 * r) wreg0 = a.w16[l2 < r ? r - l2 : l2 - r];
            if (l2 + 32 <= 2 * r) wreg1 = a.w16[l2 + 32 - r];
        }
        unsigned char* act = work + (abase - lo);
        for (int tr = 0; tr < ntr_max; ++tr) {
            const bool live = tr < ntr;
            uint32_t wa = 0, wb = 0;
            if (live) {
                if (tr == 0) { wa = w4.v[2]; wb = w4.v[3]; }
                else {
                    if (tr & 1) w4 = aps_philox4x32_10(c0, c1, (uint32_t)((tr + 1) >> 1), chi, k0, k1);
                    wa = (tr & 1) ? w4.v[0] : w4.v[2]; wb = (tr & 1) ? w4.v[1] : w4.v[3];
                }
            }
            const int x = (int)(wa >> 27);
            const uint32_t slot = wa << 5;
            const long long lx = abase + x;
            const unsigned char v = live ? act[x] : (unsigned char)APS_K2_EMPTY;
            bool want_flip = false;
            if (v != APS_K2_EMPTY) {
                if (slot < a.rates.t_left) {
                    if (lx > 0 && act[x - 1] == APS_K2_EMPTY) { act[x - 1] = v; act[x] = APS_K2_EMPTY; }
                } else if (slot < a.rates.t_right || (slot < a.rates.t_active && v == APS_K2_PLUS)) {
                    if (lx < L - 1 && act[x + 1] == APS_K2_EMPTY) { act[x + 1] = v; act[x] = APS_K2_EMPTY; }
                } else if (slot >= a.rates.t_active) want_flip = true;
            }
            const int sg = (v == APS_K2_PLUS) ? 1 : -1;
            uint32_t thr = 0;
            if (LOCAL) {
                int my_sw = 0, my_tw = 0;
                unsigned need = __ballot_sync(0xffffffffu, want_flip);
                const int lane = tid & 31;
                const int centre = (int)(lx - lo) + R16;                  // snap index of this lane's site
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int cidx = __shfl_sync(0xffffffffu, centre, src);
                    int sw, tw;
                    if (r <= 31) {          // at most two taps per lane, weights held in registers
                        const int cv0 = snap[cidx - r + lane];
                        const int cv1 = (lane + 32 <= 2 * r) ? snap[cidx - r + lane + 32] : 0;
                        sw = wreg0 * ((cv0 & 1) - (cv0 >> 1)) + wreg1 * ((cv1 & 1) - (cv1 >> 1));
                        tw = wreg0 * (cv0 != 0) + wreg1 * (cv1 != 0);
                    } else {
                        sw = 0; tw = 0;
                        for (int k = lane; k <= 2 * r; k += 32) {
                            const int cv = snap[cidx - r + k];
                            const int wj = a.w16[k < r ? r - k : k - r];
                            sw += wj * ((cv & 1) - (cv >> 1));
                            tw += wj * (cv != 0);
                        }
                    }
                    sw = __reduce_add_sync(0xffffffffu, sw);
                    tw = __reduce_add_sync(0xffffffffu, tw);
                    if (lane == src) { my_sw = sw; my_tw = tw; }
                }
                if (want_flip) thr = a.flip_tab[(sg == 1 ? 0 : (2 * APS_K2_MQ + 1)) + aps_k2_mq_index(my_sw, my_tw)];
            } else if (want_flip) thr = thr_glob[sg == 1 ? 0 : 1];
            if (want_flip && wb < thr) { act[x] = (v == APS_K2_PLUS) ? APS_K2_MINUS : APS_K2_PLUS; dsig -= 2 * sg; }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            k2_tma_store(out + lo, work, (uint32_t)(hi - lo));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the buffer used one iteration ago is free once its store has finished READING shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            const int tn = t + (kK2Stages - 1) * step_t;
            if (tn < ntiles) issue_load(tn, (it + kK2Stages - 1) % kK2Stages);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
