// aps_k2.cuh — K2: sublattice-parallel kernel for lattices too large for shared memory.
// Update rule: include/aps_k2_model.h (synchronous sublattice KMC; shared with the oracle).
//
// HBM design: one byte per site, one pass = one streamed read + one streamed write of the lattice
// (2 B per site-visit).  A persistent CTA walks windows of APS_K2_TILE sites whose borders sit in the middle
// of the inactive halves of that parity (so no update ever crosses a window border): a TMA bulk load
// (mbarrier complete_tx) brings the window — for the local field together with its +-r halo — into a ring
// of shared-memory stages, each of the 128 threads runs the trials of one 64-site segment in place, and a
// TMA bulk store writes the window to the ping-pong buffer.  Random bits are spent per EVENT (Poisson
// number of trials per segment), not per site, which is what keeps the kernel memory-bound at small
// rate*dt.  The local magnetisation needs the state at the START of the pass: all fields of a
// tile are evaluated before a block barrier and only then is the tile modified (see k2_pass_kernel).
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
constexpr int kK2StagesLocal = 3;    // local field: window + halo per stage, plus one frozen copy per CTA

// ---- local-field mode: per-CTA scratch for the three-phase tile update (see k2_pass_kernel) ----
// Trials of a segment beyond the stash capacity (and flip candidates beyond the list capacity) are rare and are
// replayed inline from the Philox counters, so the capacities only affect speed, never results.
__host__ __device__ inline int k2_stash_cap(double mu) { return mu <= 2.6 ? 8 : (mu <= 8.0 ? 16 : 32); }
__host__ __device__ inline bool k2_packed_taps(int radius) { return radius >= 1 && radius <= 127; }
__host__ __device__ inline int k2_tap_words(int radius) { return (2 * radius + 7) / 4; }   // covers 2r+1 bytes at any alignment
__host__ __device__ inline size_t k2_scratch_bytes(int radius, int cap) {
    size_t b = k2_packed_taps(radius) ? (size_t)k2_tap_words(radius) * 4 * 8 : 0;   // taps[word][shift] (2 x u32)
    b += (size_t)kK2Threads * cap * (4 + 4 + 2);     // per warp: acceptance words, candidate list, 16-bit trial codes
    return (b + 15) & ~(size_t)15;
}

// sum of the integer taps over the '+' and '-' sites of the window centred at snap[c]: four sites per step with dp2a
// (two 16-bit taps x two site bytes per instruction); taps[k*4 + s] holds the taps of word k for window alignment s.
__device__ __forceinline__ void k2_field_packed(const unsigned char* snap, int c, int r, const uint2* taps, int nwords,
                                                int& sw, int& tw) {
    const int start = c - r, sft = start & 3;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(snap + (start - sft));
    unsigned P = 0, Mi = 0;
    for (int k = 0; k < nwords; ++k) {
        const uint32_t w = wp[k];
        const uint2 t = taps[k * 4 + sft];
        const uint32_t pl = w & 0x01010101u, mi = (w >> 1) & 0x01010101u;
        P = __dp2a_lo(t.x, pl, P); P = __dp2a_hi(t.y, pl, P);
        Mi = __dp2a_lo(t.x, mi, Mi); Mi = __dp2a_hi(t.y, mi, Mi);
    }
    sw = (int)P - (int)Mi; tw = (int)(P + Mi);
}
__device__ __forceinline__ void k2_field_scalar(const unsigned char* snap, int c, int r, const int32_t* __restrict__ w16, int& sw, int& tw) {
    sw = 0; tw = 0;
    for (int k = 0; k <= 2 * r; ++k) {
        const int cv = snap[c - r + k];
        const int wj = w16[k < r ? r - k : k - r];
        sw += wj * ((cv & 1) - (cv >> 1));
        tw += wj * (cv != 0);
    }
}

// ---- many passes per launch + slab decomposition over peer memory (NVLink) ---------------------------------------
// One cooperative launch runs `n_passes` passes: the persistent CTAs meet at a grid barrier (device-scope atomic
// counter) between passes instead of returning to the host, so a 2^26-site lattice no longer pays a launch and a
// pipeline ramp-up per ~20 us pass.  With several ranks (one process per GPU, contiguous slabs with a ghost zone of
// `ghost` sites per interior side, see sublattice.py) the same kernel does the exchange itself through peer memory
// mapped with CUDA IPC: system-scope release stores of a tag into the neighbour's flag words, acquire loads on the
// own ones — no NCCL call and no host round trip inside the time stepping:
//   * ghost refresh (every `refresh_every` passes): a rank copies its two owned edges into its staging area, the
//     neighbours pull them over NVLink into their ghost zones (double-buffered staging, tags = absolute pass number);
//   * global-magnetisation mode: every pass each rank stores its cumulative flip increment into every peer's mailbox
//     (8 bytes per peer) in the same flag round, so all ranks use the lattice-wide sum(sigma) of the single-slab run.
// Every spin has a time-out that raises `*err` and makes all CTAs leave the kernel (no hung GPU on a lost peer).
constexpr int kK2MaxRanks = 8;
constexpr int kK2GhostMax = 65536;
struct K2PeerRegion {                          // one per rank, in IPC-shared device memory, zero-initialised
    unsigned long long mail_tag[2][kK2MaxRanks];   // [pass parity][source rank]   tag = absolute pass number + 1
    long long mail_val[2][kK2MaxRanks];            // cumulative sum of the source rank's own flip increments
    unsigned long long ready_tag[2][2];            // [refresh parity][side]: the neighbour's edge for my ghost `side` is staged
    unsigned long long pad[4];
    unsigned char edge[2][2][kK2GhostMax];         // staging [refresh parity][0 = my left edge, 1 = my right edge]
};
struct K2Multi {
    int32_t n_passes, world, rank, refresh_every;
    long long ghost, own_lo, own_hi;           // buffer coordinates of the owned range; ghost sites per interior side
    uint8_t* buf[2];                           // pass j of this launch reads buf[j & 1] and writes buf[(j + 1) & 1]
    unsigned* gbar;                            // grid-barrier counter, zero at launch
    int32_t* err;                              // time-out flag
    long long* acc;                            // global field: cumulative own flip increments since the lattice was created
    long long* msum0;                          // global field: lattice-wide sum(sigma) at creation
    long long* msum_cur;                       // global field: lattice-wide sum(sigma) before the current pass (kept up to date)
    K2PeerRegion* peer[kK2MaxRanks];           // IPC-mapped regions of all ranks (peer[rank] = own), world > 1 only
};

__device__ __forceinline__ void k2_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long k2_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned k2_ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long kK2SpinLimit = 6000000000LL;   // ~3 s of SM clocks: a peer that far behind is lost
// one thread spins until cond() or until the error flag is up / the time limit is hit (then it raises the flag)
template <class F>
__device__ __forceinline__ bool k2_spin(F cond, int32_t* err) {
    const long long t0 = clock64();
    while (!cond()) {
        if (*reinterpret_cast<volatile int32_t*>(err)) return false;
        if (clock64() - t0 > kK2SpinLimit) { atomicExch(err, 1); return false; }
    }
    return true;
}
// grid barrier of the co-resident CTAs (cooperative launch): monotone counter, round r waits for r * gridDim.x arrivals
__device__ __forceinline__ void k2_grid_sync(const K2Multi& m, unsigned& round) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++round;
        __threadfence();
        atomicAdd(m.gbar, 1u);
        const unsigned target = round * gridDim.x;
        k2_spin([&] { return k2_ld_acquire_gpu(m.gbar) >= target; }, m.err);
        __threadfence();
    }
    __syncthreads();
}

// One pass, persistent CTAs.  Each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... through a
// kK2Stages-deep ring of shared-memory buffers: TMA bulk loads (mbarrier complete_tx) bring the window
// (and, for the local field, a second read-only copy with the +-r halo) in, the 128 threads run the
// trials of their segments in place, and a TMA bulk store writes the window to the ping-pong buffer.
// LOCAL: Gaussian local field from the frozen copy; otherwise the global magnetisation.
//
// Global field: one thread walks the trials of its segment (Philox on the fly, two integer compares per trial).
// Local field: the field of a flip trial costs 2r+1 taps and depends only on the FROZEN copy, so each warp splits
// the update of its 32 segments into three warp-synchronous phases instead of letting one lane walk taps while
// 31 wait:
//   A  the trials are generated up front and stashed as 16 bits (site, rate slot) + the acceptance word; the flip
//      candidates are compacted into a per-warp list with ballots;
//   B  the list is evaluated densely, one candidate per lane (dp2a over the frozen bytes, four sites per step),
//      and both acceptance bits (particle '+' / '-') are written back into the stash;
//   -- block barrier: every field of the tile has been read from the still unmodified buffer, which therefore IS the
//      frozen state (no second copy of the window in shared memory, one bulk load per tile incl. the +-r halo) --
//   C  each lane replays its stashed trials in order against the live tile: no RNG, no taps, a few integer ops.
// The decisions are the same function of (Philox words, state, frozen field) as in the oracle's serial loop;
// trials beyond the stash capacity keep their two acceptance bits in register masks (Philox regenerated in C).
// MULTI = false: one pass a.in -> a.out (aps_k2_pass_device).  MULTI = true: m.n_passes passes ping-ponging between
// m.buf[0] and m.buf[1] with grid barriers and, for world > 1, the peer-memory exchanges described above.
template <bool LOCAL, bool MULTI>
__global__ void __launch_bounds__(kK2Threads) k2_pass_kernel(const __grid_constant__ aps_k2_args a, const int stash_cap,
                                                             const __grid_constant__ K2Multi m, const int nstages) {
    extern __shared__ __align__(128) unsigned char k2_raw[];
    const int tid = threadIdx.x;
    const long long L = a.L;                         // sites held by this call (slab incl. ghosts)
    const int r = LOCAL ? a.radius : 0;
    const int R16 = LOCAL ? ((r + 15) & ~15) : 0;
    const int ntiles = (int)(L / kK2Tile);
    constexpr int WB = kK2Tile + 32;                 // largest window
    // ring depth: kK2StagesLocal / kK2StagesGlobal for long tile walks; 2 for short slabs (a tile or two per CTA: more CTAs per SM
    // beat a deeper ring there, see k2_plan in aps_capi.cu)
    const int kK2Stages = nstages;
    const int stride = WB + 2 * R16;                 // LOCAL: [lo - R16, hi + R16) in one buffer, the window at offset R16
    uint64_t* bars = reinterpret_cast<uint64_t*>(k2_raw);
    unsigned char* bufs = k2_raw + 128;
    __shared__ uint32_t thr_glob[2];
    __shared__ long long mail_sum[kK2MaxRanks];
    __shared__ long long msum_sh;                    // MULTI, global field: lattice-wide sum(sigma) before the current pass
    __shared__ uint32_t cdf_sh[APS_K2_MAX_TRIALS];   // Poisson cdf thresholds for the per-lane binary search of the trial count

    // local-field scratch behind the ring: taps, then per warp [cap][32] acceptance words, candidate list, trial codes
    const int cap = stash_cap;
    const bool packed = LOCAL && k2_packed_taps(r) && a.w16[0] <= 65535;
    const int nwords = k2_tap_words(r);
    unsigned char* scr = bufs + (size_t)kK2Stages * stride;
    uint2* taps = reinterpret_cast<uint2*>(scr);
    uint32_t* base32 = reinterpret_cast<uint32_t*>(scr + (k2_packed_taps(r) ? (size_t)nwords * 32 : 0));
    const int per = 32 * cap, wv = tid >> 5;
    uint32_t* wb32 = base32 + wv * per;
    uint32_t* list = base32 + (kK2Threads / 32) * per + wv * per;
    uint16_t* code16 = reinterpret_cast<uint16_t*>(base32 + 2 * (kK2Threads / 32) * per) + wv * per;

    if (tid == 0) {
        for (int s2 = 0; s2 < kK2Stages; ++s2) k2_mbar_init(&bars[s2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (MULTI && !LOCAL) msum_sh = *m.msum_cur;
    }
    if (tid < APS_K2_MAX_TRIALS) cdf_sh[tid] = a.rates.cdf32[tid];
    if (LOCAL && packed) {
        for (int e = tid; e < nwords * 4; e += kK2Threads) {
            const int k = e >> 2, sft = e & 3;
            uint32_t tw4[4];
            for (int j = 0; j < 4; ++j) {
                const int ti = 4 * k + j - sft;
                tw4[j] = (ti >= 0 && ti <= 2 * r) ? (uint32_t)a.w16[ti < r ? r - ti : ti - r] : 0u;
            }
            taps[e] = make_uint2(tw4[0] | (tw4[1] << 16), tw4[2] | (tw4[3] << 16));
        }
    }
    __syncthreads();

    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const uint32_t t_left = a.rates.t_left, t_right = a.rates.t_right, t_active = a.rates.t_active;
    const int my_first = blockIdx.x, step_t = gridDim.x;
    int s2 = 0;                                      // ring slot and mbarrier phase of the next tile: they run on across the passes of a
    uint32_t ring_phase = 0;                         // launch (kept as counters: `it % nstages` with a run-time depth cost 32 instructions per tile)
    unsigned bar_round = 0;
    const int n_passes = MULTI ? m.n_passes : 1;
  for (int pj = 0; pj < n_passes; ++pj) {
    const uint64_t pass_abs = a.pass + (uint64_t)pj;
    const uint8_t* __restrict__ in = MULTI ? m.buf[pj & 1] : a.in;
    uint8_t* __restrict__ out = MULTI ? m.buf[(pj + 1) & 1] : a.out;
    const int qpar = (int)(pass_abs & 1ULL);         // parity; slabs start on tile boundaries, so it is global
    const int sh = qpar * APS_K2_HALF - kK2Margin;
    if (!LOCAL) {
        if (tid == 0) {
            const long long msum = MULTI ? msum_sh : (long long)*a.msum_in;
            const double mg = APS_DIV((double)msum, (double)a.n_particles);
            thr_glob[0] = aps_k2_flip_thr(a.rates.beta, +1, mg, a.rates.inv_cmax, a.rates.t_active);
            thr_glob[1] = aps_k2_flip_thr(a.rates.beta, -1, mg, a.rates.inv_cmax, a.rates.t_active);
        }
        __syncthreads();
    }

    auto window = [&](int t, long long& lo, long long& hi) {
        lo = (long long)t * kK2Tile + sh; hi = lo + kK2Tile;
        if (t == 0) lo = 0;
        if (t == ntiles - 1) hi = L;
    };
    auto issue_load = [&](int t, int s2) {          // thread 0 only
        long long lo, hi; window(t, lo, hi);
        unsigned char* work_b = bufs + (size_t)s2 * stride;
        uint32_t bytes = (uint32_t)(hi - lo);
        long long slo = lo - R16, shi = hi + R16;
        if (slo < 0) slo = 0;
        if (shi > L) shi = L;
        if (LOCAL) {       // window and +-R16 halo in one bulk copy
            k2_mbar_expect_tx(&bars[s2], (uint32_t)(shi - slo));
            k2_tma_load(work_b + (slo - (lo - R16)), in + slo, (uint32_t)(shi - slo), &bars[s2]);
        } else {
            k2_mbar_expect_tx(&bars[s2], bytes);
            k2_tma_load(work_b, in + lo, bytes, &bars[s2]);
        }
    };

    if (tid == 0) {
        for (int k = 0; k < kK2Stages - 1; ++k) { int t = my_first + k * step_t; if (t < ntiles) issue_load(t, s2 + k < kK2Stages ? s2 + k : s2 + k - kK2Stages); }
    }
    int dsig = 0;
    for (int t = my_first; t < ntiles; t += step_t) {
        long long lo, hi; window(t, lo, hi);
        unsigned char* work_b = bufs + (size_t)s2 * stride;
        // LOCAL: work_b[i - (lo - R16)] = site i; every field is read in phase B, before the barrier that lets phase C
        // modify the tile, so the live buffer IS the frozen state the model asks for
        unsigned char* snap = work_b;
        k2_mbar_wait(&bars[s2], ring_phase);
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
        const uint32_t c0 = (uint32_t)seg_global, c1 = (uint32_t)pass_abs;
        const uint32_t chi = (uint32_t)(seg_global >> 32) * 0x9E3779B9u + APS_RNG_SUBLATTICE;
        const long long abase = t0 + (long long)tid * APS_K2_SEG + qpar * APS_K2_HALF;     // first active site (slab index)
        const bool seg_ok = abase + APS_K2_HALF <= L;
        // slab decomposition: flips of ghost segments are recomputed by the neighbour rank and must not be counted twice
        const int dsig_on = (a.count_hi <= a.count_lo) || (abase >= a.count_lo && abase < a.count_hi);
        aps_u32x4 w4 = aps_philox4x32_10(c0, c1, 0u, chi, k0, k1);
        // Poisson count n = #{k < n_cdf : w >= cdf32[k]} (cdf32 is non-decreasing): branch-free binary search in shared memory
        int ntr = 0;
#pragma unroll
        for (int stp = APS_K2_MAX_TRIALS / 2; stp >= 1; stp >>= 1) {
            const int k = ntr + stp;
            if (k <= (int)a.rates.n_cdf && w4.v[0] >= cdf_sh[k - 1]) ntr = k;
        }
        if (!seg_ok) ntr = 0;
        const uint32_t thr_p_glob = LOCAL ? 0u : thr_glob[0], thr_m_glob = LOCAL ? 0u : thr_glob[1];
        unsigned char* act = work_b + R16 + (abase - lo);

        // one trial against the live tile, written on predicates (ncu, round 2: the four rate slots as four divergent paths cost
        // ~125 warp-instructions per trial step, the first branch-free form with an integer slot category still ~70):
        //   left / flip / active = position of the rate slot; d = hop direction (0 for a flip); the neighbour byte is read
        //   unconditionally (it lies inside the staged window); acc_p / acc_m = a flip of a '+' / '-' particle is accepted.
        // Reflecting walls: the byte beyond the first / last lattice site is a non-empty SENTINEL in the staged window, so
        // "neighbour empty" fails there by itself (no wall test per trial); see set_wall_sentinels below.
        const int dsig2 = 2 * dsig_on;
        auto apply = [&](uint32_t x, bool is_left, bool is_flip, bool is_act, bool acc_p, bool acc_m, bool live) {
            unsigned char* px = act + x;
            const uint32_t v = *px;
            const int d = is_left ? -1 : (is_flip ? 0 : 1);
            const uint32_t nb = px[d];
            const bool plus = v == APS_K2_PLUS;
            const bool part = live && v != APS_K2_EMPTY;
            const bool mv = part && !is_flip && (!is_act || plus) && nb == APS_K2_EMPTY;
            const bool fl = part && is_flip && (plus ? acc_p : acc_m);
            if (mv) { px[d] = (unsigned char)v; *px = APS_K2_EMPTY; }
            if (fl) { *px = (unsigned char)(v ^ 3u); dsig += plus ? -dsig2 : dsig2; }     // '+' -> '-': -2, '-' -> '+': +2
        };
        auto set_wall_sentinels = [&]() {        // only the two threads that own the wall segments touch these bytes
            if (t == 0 && tid == 0 && qpar == 0) act[-1] = 0xFF;
            if (t == ntiles - 1 && tid == kK2Threads - 1 && qpar == 1) act[APS_K2_HALF] = 0xFF;
        };
        auto category = [&](uint32_t slot) { return (int)(slot >= t_left) + (int)(slot >= t_right) + (int)(slot >= t_active); };

        // Philox words of the trials: call 0 holds the count and trials 0..2, call c >= 1 trials 4c-1 .. 4c+2 (one word per
        // trial); the loops below walk the calls and address the four words of a call at compile-time positions
        if constexpr (!LOCAL) {
            set_wall_sentinels();
            auto do_trial = [&](uint32_t wa, bool live) {
                const uint32_t slot = wa << 5, wb = slot - t_active;
                apply(wa >> 27, slot < t_left, slot >= t_active, slot >= t_right && slot < t_active, wb < thr_p_glob, wb < thr_m_glob, live);
            };
            if (ntr > 0) {
                do_trial(w4.v[1], true); do_trial(w4.v[2], ntr > 1); do_trial(w4.v[3], ntr > 2);
            }
            for (int base = 3; base < ntr; base += 4) {
                w4 = aps_philox4x32_10(c0, c1, 1u + (uint32_t)((base - 3) >> 2), chi, k0, k1);
                do_trial(w4.v[0], true); do_trial(w4.v[1], base + 1 < ntr); do_trial(w4.v[2], base + 2 < ntr); do_trial(w4.v[3], base + 3 < ntr);
            }
        } else {
            const int lane = tid & 31;
            const int cbase = (int)(abase - lo) + R16;                     // snap index of this segment's first active site
            auto field_thresholds = [&](int centre, uint32_t& thr_p, uint32_t& thr_m) {
                int sw, tw;
                if (packed) k2_field_packed(snap, centre, r, taps, nwords, sw, tw);
                else k2_field_scalar(snap, centre, r, a.w16, sw, tw);
                const int mq = aps_k2_mq_index(sw, tw);
                thr_p = a.flip_tab[mq]; thr_m = a.flip_tab[2 * APS_K2_MQ + 1 + mq];
            };
            // ---- phase A: generate, stash ([trial][lane] layout) and compact the flip candidates ----
            const int tmax = ntr < cap ? ntr : cap;
            const int it_max = __reduce_max_sync(0xffffffffu, tmax);
            int nl = 0;                                                    // warp-uniform length of the candidate list
            auto stash_trial = [&](int tr, uint32_t wa) {                  // warp-uniform call sites (tr < it_max)
                const bool live = tr < tmax;
                const uint32_t wb = (wa << 5) - t_active;                  // position inside the flip slot (flip trials only)
                const int x = (int)(wa >> 27), cat = category(wa << 5), slot = tr * 32 + lane;
                const bool cand = live && cat == 3;
                const unsigned mask = __ballot_sync(0xffffffffu, cand);
                if (live) code16[slot] = (uint16_t)((uint32_t)x | ((uint32_t)cat << 5));
                if (cand) {
                    wb32[slot] = wb;
                    list[nl + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(cbase + x) | ((uint32_t)slot << 14);
                }
                nl += __popc(mask);
            };
            if (it_max > 0) stash_trial(0, w4.v[1]);
            if (it_max > 1) stash_trial(1, w4.v[2]);
            if (it_max > 2) stash_trial(2, w4.v[3]);
            for (int base = 3; base < it_max; base += 4) {
                w4 = aps_philox4x32_10(c0, c1, 1u + (uint32_t)((base - 3) >> 2), chi, k0, k1);
                stash_trial(base, w4.v[0]);
                if (base + 1 < it_max) stash_trial(base + 1, w4.v[1]);
                if (base + 2 < it_max) stash_trial(base + 2, w4.v[2]);
                if (base + 3 < it_max) stash_trial(base + 3, w4.v[3]);
            }
            __syncwarp();
            // ---- phase B: dense evaluation of the flip candidates, one per lane ----
            for (int e = lane; e < nl; e += 32) {
                const uint32_t ent = list[e];
                const int slot = (int)(ent >> 14);
                uint32_t thr_p, thr_m;
                field_thresholds((int)(ent & 0x3fffu), thr_p, thr_m);
                const uint32_t wb = wb32[slot];
                code16[slot] = (uint16_t)(code16[slot] | (wb < thr_p ? 1u << 7 : 0u) | (wb < thr_m ? 1u << 8 : 0u));
            }
            // trials beyond the stash capacity (rare): acceptance bits into two register masks, still before the barrier
            unsigned long long ov_p = 0ULL, ov_m = 0ULL;
            for (int tr = cap; tr < ntr; ++tr) {
                const aps_u32x4 wq = aps_philox4x32_10(c0, c1, aps_k2_trial_call(tr), chi, k0, k1);
                const int wsel = aps_k2_trial_word(tr);
                const uint32_t wa = wsel == 0 ? wq.v[0] : (wsel == 1 ? wq.v[1] : (wsel == 2 ? wq.v[2] : wq.v[3]));
                const uint32_t wb = (wa << 5) - t_active;
                if (category(wa << 5) == 3) {
                    uint32_t thr_p, thr_m;
                    field_thresholds(cbase + (int)(wa >> 27), thr_p, thr_m);
                    ov_p |= (unsigned long long)(wb < thr_p) << (tr - cap);
                    ov_m |= (unsigned long long)(wb < thr_m) << (tr - cap);
                }
            }
            __syncthreads();                                               // every field of the tile has been read
            set_wall_sentinels();                                          // (the reflect padding of phase B is no longer needed)
            // ---- phase C: replay the trials in order against the live tile ----
            for (int tr = 0; tr < ntr; ++tr) {
                if (tr < cap) {
                    const uint32_t code = code16[tr * 32 + lane];
                    const uint32_t cat = code & 0x60u;
                    apply(code & 31u, cat == 0u, cat == 0x60u, cat == 0x40u, (code & 0x80u) != 0u, (code & 0x100u) != 0u, true);
                } else {
                    const aps_u32x4 wq = aps_philox4x32_10(c0, c1, aps_k2_trial_call(tr), chi, k0, k1);
                    const int wsel = aps_k2_trial_word(tr);
                    const uint32_t wa = wsel == 0 ? wq.v[0] : (wsel == 1 ? wq.v[1] : (wsel == 2 ? wq.v[2] : wq.v[3]));
                    const uint32_t slot = wa << 5;
                    apply(wa >> 27, slot < t_left, slot >= t_active, slot >= t_right && slot < t_active,
                          ((ov_p >> (tr - cap)) & 1ULL) != 0ULL, ((ov_m >> (tr - cap)) & 1ULL) != 0ULL, true);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            k2_tma_store(out + lo, work_b + R16, (uint32_t)(hi - lo));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the buffer used one iteration ago is free once its store has finished READING shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            const int tn = t + (kK2Stages - 1) * step_t;
            if (tn < ntiles) issue_load(tn, s2 == 0 ? kK2Stages - 1 : s2 - 1);
        }
        if (++s2 == kK2Stages) { s2 = 0; ring_phase ^= 1u; }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    long long* const sum_target = MULTI ? (LOCAL ? nullptr : m.acc) : reinterpret_cast<long long*>(a.msum_out);
    if (sum_target) {   // running sum(sigma) for the global magnetisation of the next pass
        for (int o = 16; o > 0; o >>= 1) dsig += __shfl_xor_sync(0xffffffffu, dsig, o);
        if ((tid & 31) == 0 && dsig != 0) atomicAdd(reinterpret_cast<unsigned long long*>(sum_target), (unsigned long long)(long long)dsig);
    }
    if (!MULTI) break;

    // ================= between two passes of one launch =================
    // (A) every CTA has written its tiles (bulk stores complete) and added its flips: the pass is finished on this GPU
    asm volatile("fence.proxy.async;" ::: "memory");
    k2_grid_sync(m, bar_round);
    if (*reinterpret_cast<volatile int32_t*>(m.err)) break;
    const unsigned long long tag = (unsigned long long)pass_abs + 1ULL;
    const bool refresh = m.world > 1 && ((pass_abs + 1ULL) % (uint64_t)m.refresh_every) == 0ULL;
    if (!LOCAL) {
        // lattice-wide sum(sigma) = value at creation + cumulative flip increments of every rank.  The own increments are
        // complete (barrier A).  CTA 0 posts them into every peer's mailbox; EVERY CTA then polls the own mailboxes (local
        // memory, written remotely over NVLink) and forms the sum itself, so no second grid barrier is needed.
        const int par = (int)(pass_abs & 1ULL);
        if (tid < m.world) {
            const long long mine = *reinterpret_cast<volatile long long*>(m.acc);
            if (tid == m.rank) mail_sum[tid] = mine;
            else {
                if (blockIdx.x == 0) {
                    K2PeerRegion* pr = m.peer[tid];
                    *reinterpret_cast<volatile long long*>(&pr->mail_val[par][m.rank]) = mine;
                    k2_st_release_sys(&pr->mail_tag[par][m.rank], tag);
                }
                K2PeerRegion* me = m.peer[m.rank];
                k2_spin([&] { return k2_ld_acquire_sys(&me->mail_tag[par][tid]) == tag; }, m.err);
                mail_sum[tid] = *reinterpret_cast<volatile long long*>(&me->mail_val[par][tid]);
            }
        }
        __syncthreads();
        if (tid == 0) {
            long long sg = *m.msum0;
            for (int q = 0; q < m.world; ++q) sg += mail_sum[q];
            msum_sh = sg;
            if (blockIdx.x == 0) *reinterpret_cast<volatile long long*>(m.msum_cur) = sg;     // for the host / the next launch
        }
        __syncthreads();
    }
    if (refresh) {
        // (B) stage the two owned edges, tell the neighbours, pull theirs into the ghost zones of the new state
        const int rpar = (int)(((pass_abs + 1ULL) / (uint64_t)m.refresh_every) & 1ULL);
        K2PeerRegion* me = m.peer[m.rank];
        const long long g = m.ghost;
        const int ncopy = gridDim.x < 16 ? gridDim.x : 16;
        if ((int)blockIdx.x < ncopy) {
            const long long nvec = g / 16;
            for (int side = 0; side < 2; ++side) {
                if ((side == 0 && m.rank == 0) || (side == 1 && m.rank == m.world - 1)) continue;
                const uint4* src = reinterpret_cast<const uint4*>(out + (side == 0 ? m.own_lo : m.own_hi - g));
                uint4* dst = reinterpret_cast<uint4*>(me->edge[rpar][side]);
                for (long long v = (long long)blockIdx.x * kK2Threads + tid; v < nvec; v += (long long)ncopy * kK2Threads) dst[v] = src[v];
            }
            __threadfence_system();
        }
        k2_grid_sync(m, bar_round);
        if (blockIdx.x == 0 && tid < 2) {       // my left edge is the RIGHT ghost of rank-1, my right edge the LEFT ghost of rank+1
            const int nb = tid == 0 ? m.rank - 1 : m.rank + 1;
            if (nb >= 0 && nb < m.world) k2_st_release_sys(&m.peer[nb]->ready_tag[rpar][tid == 0 ? 1 : 0], tag);
        }
        if ((int)blockIdx.x < ncopy) {
            const long long nvec = g / 16;
            for (int side = 0; side < 2; ++side) {
                const int nb = side == 0 ? m.rank - 1 : m.rank + 1;
                if (nb < 0 || nb >= m.world) continue;
                if (tid == 0) k2_spin([&] { return k2_ld_acquire_sys(&me->ready_tag[rpar][side]) == tag; }, m.err);
                __syncthreads();
                const uint4* src = reinterpret_cast<const uint4*>(m.peer[nb]->edge[rpar][1 - side]);   // NVLink peer loads
                uint4* dst = reinterpret_cast<uint4*>(out + (side == 0 ? m.own_lo - g : m.own_hi));
                for (long long v = (long long)blockIdx.x * kK2Threads + tid; v < nvec; v += (long long)ncopy * kK2Threads) dst[v] = src[v];
            }
            __threadfence();
        }
    }
    if (refresh) {                           // (C) the refreshed ghosts are visible to every CTA
        k2_grid_sync(m, bar_round);
    }
    if (*reinterpret_cast<volatile int32_t*>(m.err)) break;
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
