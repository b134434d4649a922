// aps_k1.cuh — K1: replica-batched exact Gillespie kernel (one CTA per replica).
//
// What it replaces: one launch == n_replicas calls of ParticleSystem.run()
// (PARTICLE_solver_CLASS.py:450-558), each event being one step_gillespie()
// (:254-448) preceded by compute_local_m_field() (:216-246).
//
// B200-first design (not a translation of the numpy code):
//   * the reference recomputes the whole field (2 Gaussian filters over L sites) and all n
//     rates for every single-particle event: O(L*taps + n).  Here the lattice lives in shared
//     memory as one packed uint16 per site (c_plus | c_minus<<8) with the reflect halo
//     materialised, the per-particle total rates live in shared memory, and an event only
//     re-evaluates the rates of the particles inside the window its <=2 changed sites can
//     influence (filter radius r): O(rho*r*r) work, all of it on-chip.
//   * the filter value at a particle's site is re-evaluated from scratch in the reference's
//     exact tap order (outermost pair first, separate mul and add, no FMA), so every rate is
//     bit-identical to the oracle's full recomputation.
//   * particle selection: the reference does searchsorted(cumsum(rates/R)/last, u).  A serial
//     cumsum is a 100s-of-adds dependent chain, so the kernel selects with a two-level parallel
//     prefix sum (per-thread chunk + warp shuffle scan) and proves the decision equal to the
//     serial one with a rounding-error guard band; if u lies within the band of a bin edge
//     (probability ~1e-10 per event) one thread redoes that selection in the exact serial order.
//   * R = rates.sum() is numpy's pairwise summation; it is reproduced exactly and in parallel
//     (8 strided accumulators per <=128-element leaf == 8 lanes, butterfly combine, then the
//     recursion tree), so the event clock t is bit-identical as well.
//   * HBM is touched only for the observation rows (M per run) and, in replay mode, for the
//     variate log.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aps.h"
#include "../../include/aps_math.h"
#include "../../include/aps_philox.h"

namespace aps {

constexpr int kMaxNodes = 512;  // nodes of the pairwise-sum tree (n_max <= 8192)

struct K1Args {
    aps_params p;
    aps_batch b;
    double guard_scale;   // multiplies the selection guard band (test hook; 1.0 in production)
    int32_t max_nodes;    // tree nodes the shared-memory plan reserves
    int32_t pad;          // halo width = max(radius, 0)
    int32_t use_lut;      // 1: filter taps come from the product table (see local_m_lut)
    int32_t only_retry;   // 1: run only the replicas the fast kernel flagged (status == 100)
    int32_t n_lo;         // lean kernels: a launch owns the replicas with n_lo < n <= n_hi and leaves the others untouched (size classes
    int32_t bcode;        // 2K+1: radix of the per-site code c_plus + bcode*c_minus
    int32_t wt_valid;     // 1: wt[] holds the taps w[0..radius] (radius <= 83): kernel-parameter = constant-bank operands
    int32_t n_hi;         // of one batch run as independent, concurrent launches); -1 / INT_MAX = all replicas
    double wt[84];
};

constexpr int kRing = 128;      // doubles of variate look-ahead kept in shared memory
constexpr size_t kLutMaxBytes = 24 * 1024;

// Shared-memory plan (bytes) — keep in sync with carve() below.
__host__ __device__ inline size_t k1_lut_bytes(int K, int radius) {
    if (radius < 0) return 0;
    size_t b = (size_t)(2 * K + 1);
    return (size_t)(radius + 1) * b * b * 16;
}
__host__ __device__ inline size_t k1_smem_bytes(int L, int n_max, int radius, int max_nodes, int nwarps, int use_lut, int K) {
    int pad = radius > 0 ? radius : 0;
    size_t d = 0;
    if (use_lut) d += k1_lut_bytes(K, radius) / 8;  // product table (16-byte aligned: first)
    d += (size_t)kRing;                  // variate ring
    d += (size_t)n_max;                  // rates
    d += (size_t)(pad + 1);              // taps w[0..r]
    d += (size_t)max_nodes;              // tree node values
    d += (size_t)nwarps;                 // warp totals
    d += 8;                              // draws[4], t_new, R, spare
    size_t bytes = d * 8;
    bytes += (size_t)max_nodes * 3 * 4;  // node_a, node_b (children or leaf start/len), kind
    bytes += 16 * 4;                     // desc + flags
    bytes += (((size_t)(L + 2 * pad) + 1) / 2) * 4;  // packed lattice (uint16, 4-byte granules)
    if (use_lut) bytes += (((size_t)(L + 2 * pad) + 1) / 2) * 4;  // per-site codes
    bytes += (((size_t)n_max + 1) / 2) * 4;          // pos (uint16)
    bytes += (((size_t)n_max + 3) / 4) * 4;          // sigma (int8)
    bytes += (((size_t)n_max + 3) / 4) * 4;          // bound flags (int8)
    bytes += (((size_t)L + 3) / 4) * 4;              // anchor mask (uint8)
    return bytes;
}

struct Smem {
    double2* lut; double* ring;
    double* rates; double* w; double* node_val; double* wtot; double* misc;
    int32_t* node_a; int32_t* node_b; int32_t* node_kind; int32_t* desc;
    uint16_t* pk; uint16_t* code; uint16_t* pos; int8_t* sigma;
    int8_t* bound;        // bound flags (anchors only; all zero otherwise)
    uint8_t* anchor;      // is_anchor_site, or nullptr when the batch has no anchors
    const double* m_in;   // optional injected field (global memory), see aps_batch.m_field_in
    const double* flip_tab;   // optional tabulated flip_rate_fn (global memory, read through L1), see aps_batch.flip_tab
    int64_t flip_G;
    int periodic;         // APS_FLAG_PERIODIC: the halo holds wrapped images and hops wrap
};

__device__ __forceinline__ Smem carve(unsigned char* base, int L, int n_max, int pad, int max_nodes, int nwarps,
                                      int use_lut, int K, int radius) {
    Smem s;
    double* d = reinterpret_cast<double*>(base);
    s.lut = reinterpret_cast<double2*>(d);
    if (use_lut) d += k1_lut_bytes(K, radius) / 8;
    s.ring = d; d += kRing;
    s.rates = d; d += n_max;
    s.w = d; d += pad + 1;
    s.node_val = d; d += max_nodes;
    s.wtot = d; d += nwarps;
    s.misc = d; d += 8;
    int32_t* q = reinterpret_cast<int32_t*>(d);
    s.node_a = q; q += max_nodes;
    s.node_b = q; q += max_nodes;
    s.node_kind = q; q += max_nodes;
    s.desc = q; q += 16;
    s.pk = reinterpret_cast<uint16_t*>(q); q += ((L + 2 * pad) + 1) / 2;
    s.code = reinterpret_cast<uint16_t*>(q); if (use_lut) q += ((L + 2 * pad) + 1) / 2;
    s.pos = reinterpret_cast<uint16_t*>(q); q += (n_max + 1) / 2;
    s.sigma = reinterpret_cast<int8_t*>(q); q += (n_max + 3) / 4;
    s.bound = reinterpret_cast<int8_t*>(q); q += (n_max + 3) / 4;
    s.anchor = reinterpret_cast<uint8_t*>(q);
    return s;
}

// desc[] slots
enum { D_SEQ = 0, D_PART, D_KIND, D_OLD, D_NEW, D_STOP, D_EXACT, D_NCROSS, D_END, D_AVAIL, D_NNODES, D_BADR, D_R12, D_R13, D_SG };
// misc[] slots
enum { X_TNEW = 0, X_R };

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// target of a hop by d = -1, 0, +1 from p: reflecting walls clamp, a ring wraps (CLASS.py:278-288)
__device__ __forceinline__ int hop_target(int p, int d, int L, bool per) {
    int q = p + d;
    return per ? (q < 0 ? q + L : (q >= L ? q - L : q)) : clampi(q, 0, L - 1);
}

// Add `delta` to the cell of site x and to every reflect (or, on a ring, wrapped) image inside the halo.
__device__ __forceinline__ void cell_add(uint16_t* pk, int L, int pad, int x, int delta, bool per) {
    if (per) {                           // pad <= L/2 on a ring (validated at the boundary)
        pk[pad + x] = (uint16_t)(pk[pad + x] + delta);
        if (x < pad) pk[pad + L + x] = (uint16_t)(pk[pad + L + x] + delta);
        if (x >= L - pad) pk[pad + x - L] = (uint16_t)(pk[pad + x - L] + delta);
    } else if (pad < L) {
        pk[pad + x] = (uint16_t)(pk[pad + x] + delta);
        if (x < pad) pk[pad - 1 - x] = (uint16_t)(pk[pad - 1 - x] + delta);
        if (x >= L - pad) pk[pad + 2 * L - 1 - x] = (uint16_t)(pk[pad + 2 * L - 1 - x] + delta);
    } else {
        const int twoL = 2 * L;
        const int kmax = (pad + L) / twoL + 1;
        for (int k = -kmax; k <= kmax; ++k) {
            int i1 = k * twoL + x, i2 = k * twoL - 1 - x;
            if (i1 >= -pad && i1 < L + pad) pk[pad + i1] = (uint16_t)(pk[pad + i1] + delta);
            if (i2 >= -pad && i2 < L + pad) pk[pad + i2] = (uint16_t)(pk[pad + i2] + delta);
        }
    }
}

// packed counts (c_plus | c_minus<<8) and, when the product table is in use, the radix code
__device__ __forceinline__ void pk_add(const Smem& s, int L, int pad, int x, int dplus, int dminus, int bcode, bool lut) {
    cell_add(s.pk, L, pad, x, dplus + 256 * dminus, s.periodic != 0);
    if (lut) cell_add(s.code, L, pad, x, dplus + bcode * dminus, s.periodic != 0);
}

__device__ __forceinline__ int occ_of(uint16_t v) { return (v & 0xff) + (v >> 8); }

// Hop rates of one particle, CLASS.py:276-336 (anchors absent).
__device__ __forceinline__ void hop_rates(const uint16_t* pk, int pad, int L, int K, double D, double lam, bool crowd, bool per,
                                          int p, int sg, double& rl, double& rr, double& ra) {
    int fwd = hop_target(p, sg == 1, L, per), lt = hop_target(p, -1, L, per), rt = hop_target(p, 1, L, per);
    int occ_l = occ_of(pk[pad + lt]), occ_r = occ_of(pk[pad + rt]);
    int occ_f = (fwd == rt) ? occ_r : occ_of(pk[pad + fwd]);
    bool f_free = (occ_f < K) && (fwd != p);
    bool l_free = (occ_l < K) && (lt != p);
    bool r_free = (occ_r < K) && (rt != p);
    rl = l_free ? D : APS_MUL(D, 0.0);
    rr = r_free ? D : APS_MUL(D, 0.0);
    ra = (sg == 1 && f_free) ? lam : 0.0;
    if (crowd) {
        double kd = (double)K;
        double ff = APS_SUB(1.0, APS_DIV((double)occ_f, kd));
        ff = ff < 0.0 ? 0.0 : (ff > 1.0 ? 1.0 : ff);
        ra = APS_MUL(ra, ff);
        double lf = APS_SUB(1.0, APS_DIV((double)occ_l, kd)), rf = APS_SUB(1.0, APS_DIV((double)occ_r, kd));
        lf = lf < 0.0 ? 0.0 : (lf > 1.0 ? 1.0 : lf);
        rf = rf < 0.0 ? 0.0 : (rf > 1.0 ? 1.0 : rf);
        rl = APS_MUL(rl, lf);
        rr = APS_MUL(rr, rf);
    }
}

// Every additive part of rates[i] (CLASS.py:259-351): hop parts, bind / unbind / exit (anchors) and whether
// the flip term is suppressed.  Without anchors the last three are +0.0 and the sum is unchanged.
struct RateParts { double rl, rr, rdiff, ract, rbind, runbind, rexit; bool no_flip; };

__device__ __forceinline__ RateParts rate_parts(const Smem& s, const aps_params& P, int pad, bool crowd, int i, int p, int sg) {
    RateParts r;
    hop_rates(s.pk, pad, P.L, P.K, P.rate_diffusion, P.rate_active, crowd, s.periodic != 0, p, sg, r.rl, r.rr, r.ract);
    r.rbind = 0.0; r.runbind = 0.0; r.rexit = 0.0; r.no_flip = false;
    r.rdiff = APS_ADD(r.rl, r.rr);
    if (s.anchor) {
        const bool bnd = s.bound[i] != 0, on_anchor = s.anchor[p] != 0;
        if ((P.flags & APS_FLAG_IMMOBILIZE) && sg == -1 && on_anchor && bnd) {    // anchored: immobile, may exit (:307-312,338-340)
            r.rdiff = 0.0; r.ract = 0.0; r.rexit = P.k_exit;
        }
        r.no_flip = (P.flags & APS_FLAG_SUPPRESS_FLIP_BOUND) && bnd;                 // :266-267
        if (!bnd && sg == -1 && on_anchor && occ_of(s.pk[pad + p]) < P.K) r.rbind = P.k_on;   // :343-345
        if (bnd) r.runbind = P.k_off;                                                // :347-348
    }
    return r;
}

// Local magnetisation at site p: the two Gaussian filters of CLASS.py:229-245 evaluated at one
// output in scipy's symmetric-correlate order.
__device__ __forceinline__ double local_m(const uint16_t* pk, const double* w, int pad, int r, int p) {
    const uint16_t* c = pk + pad + p;
    uint32_t v0 = c[0];
    double sc = APS_MUL((double)((int)(v0 & 0xff) - (int)(v0 >> 8)), w[r]);
    double tc = APS_MUL((double)((int)(v0 & 0xff) + (int)(v0 >> 8)), w[r]);
    for (int jj = -r; jj < 0; ++jj) {
        uint32_t a = (uint32_t)c[jj] + (uint32_t)c[-jj];
        int acp = (int)(a & 0xff), acm = (int)(a >> 8);
        double wj = w[r + jj];
        sc = APS_ADD(sc, APS_MUL((double)(acp - acm), wj));
        tc = APS_ADD(tc, APS_MUL((double)(acp + acm), wj));
    }
    double m = 0.0;
    if (tc > 0.0) m = APS_DIV(sc, tc);
    m = m < -1.0 ? -1.0 : (m > 1.0 ? 1.0 : m);
    return m;
}

// Same value, fewer instructions: per tap the pair (x[l+j]+x[l-j]) is a small integer pair
// (a_plus, a_minus) in [0,2K]^2, so both products (a_plus-a_minus)*w_j and (a_plus+a_minus)*w_j
// are looked up from a table built once per CTA with the same __dmul_rn; the radix codes of the
// two sites add up to the table index.  Adds stay in the reference order.
__device__ __forceinline__ double local_m_lut(const uint16_t* code, const double2* lut, int pad, int r, int b2, int p) {
    const uint16_t* c = code + pad + p;
    double2 v = lut[r * b2 + c[0]];
    double sc = v.x, tc = v.y;
    const double2* row = lut;
#pragma unroll 4
    for (int jj = -r; jj < 0; ++jj) {
        int idx = (int)c[jj] + (int)c[-jj];
        double2 t2 = row[idx];
        row += b2;
        sc = APS_ADD(sc, t2.x);
        tc = APS_ADD(tc, t2.y);
    }
    double m = 0.0;
    if (tc > 0.0) m = APS_DIV(sc, tc);
    m = m < -1.0 ? -1.0 : (m > 1.0 ? 1.0 : m);
    return m;
}

// rates[i] of CLASS.py:351 for one particle (bind/unbind/exit terms are +0.0).
__device__ __forceinline__ double site_m(const Smem& s, const aps_params& P, int pad, bool lut, int b2, int p) {
    return lut ? local_m_lut(s.code, s.lut, pad, P.radius, b2, p) : local_m(s.pk, s.w, pad, P.radius, p);
}
__device__ __forceinline__ double particle_rate(const Smem& s, const aps_params& P, int pad, bool crowd, double beta,
                                                double m_global, int i, int p, int sg, bool lut, int b2) {
    const RateParts r = rate_parts(s, P, pad, crowd, i, p, sg);
    double cv = 0.0;
    if (!r.no_flip) {
        double m = s.m_in ? s.m_in[p] : ((P.radius < 0) ? m_global : site_m(s, P, pad, lut, b2, p));
        cv = s.flip_tab ? aps_flip_interp(s.flip_tab, s.flip_G, sg, m) : aps_exp(APS_MUL(APS_MUL(-beta, (double)sg), m));
    }
    double rate = APS_ADD(APS_ADD(r.rdiff, r.ract), cv);
    if (s.anchor) rate = APS_ADD(APS_ADD(APS_ADD(rate, r.rbind), r.runbind), r.rexit);
    return rate;
}

// numpy pairwise-sum tree for n elements, built once per replica by one thread.
// Nodes are stored children-before-parents; kind 0 = leaf (a=start, b=len), 1 = add(a, b).
__device__ inline int build_sum_tree(int n, int32_t* na, int32_t* nb, int32_t* nk, int cap) {
    // explicit stack of (start, len, state, left_node)
    int st_start[32], st_len[32], st_state[32], st_left[32];
    int sp = 0, nn = 0, ret = -1;
    st_start[0] = 0; st_len[0] = n; st_state[0] = 0; st_left[0] = -1; sp = 1;
    while (sp > 0) {
        int top = sp - 1;
        int start = st_start[top], len = st_len[top];
        if (len <= 128) {
            if (nn >= cap) return -1;
            na[nn] = start; nb[nn] = len; nk[nn] = 0; ret = nn++; --sp;
            continue;
        }
        int n2 = len / 2; n2 -= n2 % 8;
        if (st_state[top] == 0) {
            st_state[top] = 1;
            st_start[sp] = start; st_len[sp] = n2; st_state[sp] = 0; st_left[sp] = -1; ++sp;
        } else if (st_state[top] == 1) {
            st_left[top] = ret; st_state[top] = 2;
            st_start[sp] = start + n2; st_len[sp] = len - n2; st_state[sp] = 0; st_left[sp] = -1; ++sp;
        } else {
            if (nn >= cap) return -1;
            na[nn] = st_left[top]; nb[nn] = ret; nk[nn] = 1; ret = nn++; --sp;
        }
    }
    return nn;
}

// Apply one event to the shared-memory state.
__device__ __forceinline__ void apply_event(const Smem& s, int L, int pad, int i, int kind, int oldp, int newp, int sg,
                                            int bcode, bool lut) {
    if (kind == APS_EV_BIND) s.bound[i] = 1;
    else if (kind == APS_EV_UNBIND) s.bound[i] = 0;
    else if (kind == APS_EV_EXIT) {                     // lattice only; the particle arrays are compacted by the whole CTA
        pk_add(s, L, pad, oldp, sg == 1 ? -1 : 0, sg == 1 ? 0 : -1, bcode, lut);
    } else if (kind == APS_EV_FLIP) {
        s.sigma[i] = (int8_t)(-sg);
        pk_add(s, L, pad, oldp, -sg, sg, bcode, lut);   // c_plus-1,c_minus+1  or the reverse
    } else {
        s.pos[i] = (uint16_t)newp;
        const int dp = sg == 1 ? 1 : 0, dm = 1 - dp;
        pk_add(s, L, pad, oldp, -dp, -dm, bcode, lut);
        pk_add(s, L, pad, newp, dp, dm, bcode, lut);
    }
}
// Exact inverse of apply_event on the lattice arrays only (used to expose the pre-event field).
__device__ __forceinline__ void lattice_delta(const Smem& s, int L, int pad, int kind, int oldp, int newp, int sg_old,
                                              int sign, int bcode, bool lut) {
    if (kind == APS_EV_BIND || kind == APS_EV_UNBIND) return;
    if (kind == APS_EV_EXIT) {
        pk_add(s, L, pad, oldp, sg_old == 1 ? -sign : 0, sg_old == 1 ? 0 : -sign, bcode, lut);
    } else if (kind == APS_EV_FLIP) {
        pk_add(s, L, pad, oldp, -sign * sg_old, sign * sg_old, bcode, lut);
    } else {
        const int dp = sg_old == 1 ? 1 : 0, dm = 1 - dp;
        pk_add(s, L, pad, oldp, -sign * dp, -sign * dm, bcode, lut);
        pk_add(s, L, pad, newp, sign * dp, sign * dm, bcode, lut);
    }
}

// Decide the event of particle `sel` from u_event/u_dir (CLASS.py:362-446) and apply it.
// Returns false (nothing applied) if a direction draw is needed but not available.
__device__ __forceinline__ bool decode_and_apply(const Smem& s, const aps_params& P, int pad, bool crowd, int sel,
                                                 double ue, double ud, int avail, int seq, int bcode, bool lut) {
    int p = s.pos[sel], sg = s.sigma[sel];
    const RateParts rp = rate_parts(s, P, pad, crowd, sel, p, sg);
    const double rl = rp.rl, rr = rp.rr;
    double v = APS_MUL(ue, s.rates[sel]);
    double diff_thresh = rp.rdiff;
    double act_thresh = APS_ADD(diff_thresh, rp.ract);
    double bind_thresh = APS_ADD(act_thresh, rp.rbind);          // :365-367
    double unbind_thresh = APS_ADD(bind_thresh, rp.runbind);
    double exit_thresh = APS_ADD(unbind_thresh, rp.rexit);
    int kind, newp = p;
    if (v < diff_thresh) {
        if (avail < 4) { s.desc[D_STOP] = 1; s.desc[D_SEQ] = seq; return false; }
        if (ud < APS_DIV(rl, APS_ADD(rl, rr))) { kind = APS_EV_DIFF_LEFT; newp = hop_target(p, -1, P.L, s.periodic != 0); }
        else { kind = APS_EV_DIFF_RIGHT; newp = hop_target(p, 1, P.L, s.periodic != 0); }
    } else if (v < act_thresh) {
        kind = APS_EV_ACTIVE; newp = hop_target(p, sg == 1, P.L, s.periodic != 0);
    } else if (v < bind_thresh) kind = APS_EV_BIND;
    else if (v < unbind_thresh) kind = APS_EV_UNBIND;
    else if (v < exit_thresh) kind = APS_EV_EXIT;
    else kind = APS_EV_FLIP;
    apply_event(s, P.L, pad, sel, kind, p, newp, sg, bcode, lut);
    s.desc[D_PART] = sel; s.desc[D_KIND] = kind; s.desc[D_OLD] = p; s.desc[D_NEW] = newp; s.desc[D_SG] = sg;
    s.desc[D_STOP] = 0; s.desc[D_SEQ] = seq;
    return true;
}

template <int NT>
__device__ __forceinline__ void bsync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

template <int NT, bool PHILOX>
__global__ void __launch_bounds__(NT, 1024 / NT) k1_kernel(const __grid_constant__ K1Args A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const aps_params& P = A.p;
    const aps_batch& B = A.b;
    const int rep = blockIdx.x;
    if (A.only_retry && B.status[rep] != 100) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int L = P.L, pad = A.pad, r = P.radius, n_max = B.n_max, M = B.M;
    const bool crowd = (P.flags & APS_FLAG_CROWDING) != 0;
    const bool global_m = r < 0;
    bool lut = A.use_lut != 0;
    const int bcode = A.bcode, b2 = A.bcode * A.bcode;
    int n = B.n[rep];
    const double beta = B.beta[rep], T = P.T;
    Smem s = carve(smem_raw, L, n_max, pad, A.max_nodes, NW, A.use_lut, P.K, r);
    const bool periodic = (P.flags & APS_FLAG_PERIODIC) != 0;
    s.periodic = periodic ? 1 : 0;
    s.m_in = B.m_field_in ? B.m_field_in + (size_t)rep * (size_t)L : nullptr;
    s.flip_tab = B.flip_tab; s.flip_G = B.flip_G;
    if (B.anchor_mask) { for (int l = tid; l < L; l += NT) s.anchor[l] = B.anchor_mask[l]; } else s.anchor = nullptr;
    for (int i = tid; i < n_max; i += NT) s.bound[i] = (B.bound0 && i < n) ? B.bound0[(size_t)rep * n_max + i] : 0;

    // ---------------- prologue: stage the replica into shared memory ----------------
    {
        uint32_t* pk32 = reinterpret_cast<uint32_t*>(s.pk);
        uint32_t* cd32 = reinterpret_cast<uint32_t*>(s.code);
        for (int i = tid; i < (L + 2 * pad + 1) / 2; i += NT) { pk32[i] = 0u; if (lut) cd32[i] = 0u; }
        if (r >= 0) for (int i = tid; i <= r; i += NT) s.w[i] = B.weights[i];
        if (lut) {
            // table[j][a_minus*bcode + a_plus] = ((a_plus-a_minus)*w_j, (a_plus+a_minus)*w_j)
            for (int e = tid; e < (r + 1) * b2; e += NT) {
                int j = e / b2, idx = e - j * b2, am = idx / bcode, ap = idx - am * bcode;
                double wj = B.weights[j];
                s.lut[e] = make_double2(APS_MUL((double)(ap - am), wj), APS_MUL((double)(ap + am), wj));
            }
        }
        if (tid < 16) s.desc[tid] = 0;
    }
    bsync<NT>();
    int S = 0;  // sum(sigma), replicated in every thread
    {
        const int32_t* gp = B.pos0 + (size_t)rep * n_max;
        const int8_t* gs = B.sigma0 + (size_t)rep * n_max;
        uint32_t* pk32 = reinterpret_cast<uint32_t*>(s.pk);
        uint32_t* cd32 = reinterpret_cast<uint32_t*>(s.code);
        int part = 0;
        for (int i = tid; i < n; i += NT) {
            int p = gp[i], sg = gs[i];
            s.pos[i] = (uint16_t)p; s.sigma[i] = (int8_t)sg;
            part += sg;
            const uint32_t d = sg == 1 ? 1u : 256u, dc = sg == 1 ? 1u : (uint32_t)bcode;
            // every reflect image of p inside the halo (atomic: several particles may share a word)
            const int twoL = 2 * L, kmax = periodic ? 1 : (pad + L) / twoL + 1;
            for (int k = -kmax; k <= kmax; ++k) {
                // reflect images: p + 2kL and -1 - p + 2kL; ring images: p + kL (i2 unused)
                int i1 = periodic ? k * L + p : k * twoL + p, i2 = periodic ? -pad - 1 : k * twoL - 1 - p;
                if (i1 >= -pad && i1 < L + pad) {
                    int q = pad + i1; atomicAdd(&pk32[q >> 1], d << (16 * (q & 1)));
                    if (lut) atomicAdd(&cd32[q >> 1], dc << (16 * (q & 1)));
                }
                if (i2 >= -pad && i2 < L + pad) {
                    int q = pad + i2; atomicAdd(&pk32[q >> 1], d << (16 * (q & 1)));
                    if (lut) atomicAdd(&cd32[q >> 1], dc << (16 * (q & 1)));
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s.ring[wid] = (double)part;   // the variate ring is free until the loop starts
        bsync<NT>();
        for (int w2 = 0; w2 < NW; ++w2) S += (int)s.ring[w2];
        // The product table indexes pair sums up to 2K per species; an over-capacity initial state
        // (the reference accepts one) makes this replica fall back to the arithmetic taps.
        if (lut) {
            int bad = 0;
            for (int l = tid; l < L; l += NT) { uint16_t v = s.pk[pad + l]; bad |= ((v & 0xff) > P.K) | ((v >> 8) > P.K); }
            bad = (NT == 32) ? __any_sync(0xffffffffu, bad) : __syncthreads_or(bad);
            if (bad) lut = false;
        }
        bsync<NT>();
        if (tid == 0) s.desc[D_NNODES] = (n > 0) ? build_sum_tree(n, s.node_a, s.node_b, s.node_kind, A.max_nodes) : 0;
    }
    bsync<NT>();
    int nnodes = s.desc[D_NNODES];
    int n_exit = (B.n_exit && B.ev_start) ? B.n_exit[rep] : 0;

    int64_t n_done = 0;
    const int64_t ev_base = B.ev_start ? B.ev_start[rep] : 0;
    int64_t cursor = 0, n_guard = 0, rbase = -(int64_t)kRing - 8;
    const int64_t draws_len = PHILOX ? 0 : (B.draw_off[rep + 1] - B.draw_off[rep]);
    const double* gdraws = PHILOX ? nullptr : (B.draws + B.draw_off[rep]);
    const uint32_t k0 = PHILOX ? (uint32_t)B.seeds[rep] : 0u, k1 = PHILOX ? (uint32_t)(B.seeds[rep] >> 32) : 0u;
    double t = B.t_start ? B.t_start[rep] : 0.0;
    int obs_idx = B.obs_start ? B.obs_start[rep] : 0;
    int status = APS_RUN_DONE;
    double m_glob = (global_m && n > 0) ? APS_DIV((double)S, (double)n) : 0.0;

    // observation row writer (post-event state; field written separately)
    auto write_rows = [&](int first, int count) {
        for (int m = first; m < first + count; ++m) {
            size_t row = (size_t)rep * (size_t)M + (size_t)m;
            if ((B.record & APS_REC_COUNTS) && B.obs_cp && B.obs_cm) {
                int8_t* ocp = B.obs_cp + row * (size_t)L; int8_t* ocm = B.obs_cm + row * (size_t)L;
                for (int l = tid; l < L; l += NT) { uint16_t v = s.pk[pad + l]; ocp[l] = (int8_t)(v & 0xff); ocm[l] = (int8_t)(v >> 8); }
            }
            if ((B.record & APS_REC_POS) && B.obs_pos) {
                int32_t* op = B.obs_pos + row * (size_t)n_max;
                for (int i = tid; i < n; i += NT) op[i] = (int32_t)s.pos[i];
            }
            if (B.obs_sigma_sum && tid == 0) B.obs_sigma_sum[row] = S;
            if (B.obs_n && tid == 0) B.obs_n[row] = n;
            if (B.obs_bound) { int8_t* ob = B.obs_bound + row * (size_t)n_max; for (int i = tid; i < n; i += NT) ob[i] = s.bound[i]; }
        }
    };
    auto write_field = [&](int first, int count) {
        if (!((B.record & APS_REC_MLOCAL) && B.obs_m_local)) return;
        for (int l = tid; l < L; l += NT) {
            double m = global_m ? m_glob : site_m(s, P, pad, lut, b2, l);
            for (int mm = first; mm < first + count; ++mm)
                B.obs_m_local[((size_t)rep * (size_t)M + (size_t)mm) * (size_t)L + l] = m;
        }
    };

    if (n == 0 || nnodes <= 0) {
        status = APS_RUN_EMPTY;
    } else {
        // initial rates (every particle) and observation row 0
        for (int i = tid; i < n; i += NT)
            s.rates[i] = particle_rate(s, P, pad, crowd, beta, m_glob, i, s.pos[i], s.sigma[i], lut, b2);
        if (obs_idx == 0 && M > 0) { write_field(0, 1); write_rows(0, 1); obs_idx = 1; }
        bsync<NT>();
        double next_obs = (obs_idx < M) ? B.times_obs[obs_idx] : 0.0;
        double guard = A.guard_scale * 4.0 * (double)(n + 32) * 1.1102230246251565e-16;
        int chunk = (n + NT - 1) / NT;

        while (true) {
            if (!(t < T)) { status = APS_RUN_DONE; break; }
            if (n == 0) { status = APS_RUN_EMPTY; break; }          // every particle has exited (CLASS.py:256-257)
            if (B.max_events > 0 && n_done >= B.max_events) { status = APS_RUN_MAX_EVENTS; break; }

            // ---- variates of this event from the shared-memory ring (refilled by a whole warp) ----
            int avail; double e, uc, ue, ud;
            if (PHILOX) {
                const int slot = (int)(n_done & 31);
                if (slot == 0) {
                    if (tid < 32) {   // 32 events ahead, one per lane
                        uint64_t ev = (uint64_t)(ev_base + n_done + tid);
                        aps_u32x4 a = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_A, 0u, k0, k1);
                        aps_u32x4 b = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_B, 0u, k0, k1);
                        s.ring[4 * tid + 0] = -aps_log(APS_SUB(1.0, aps_u53(a.v[0], a.v[1])));
                        s.ring[4 * tid + 1] = aps_u53(a.v[2], a.v[3]);
                        s.ring[4 * tid + 2] = aps_u53(b.v[0], b.v[1]);
                        s.ring[4 * tid + 3] = aps_u53(b.v[2], b.v[3]);
                    }
                    bsync<NT>();
                }
                avail = 4;
                e = s.ring[4 * slot]; uc = s.ring[4 * slot + 1]; ue = s.ring[4 * slot + 2]; ud = s.ring[4 * slot + 3];
            } else {
                if (cursor + 4 > rbase + kRing) {
                    bsync<NT>();
                    rbase = cursor;
                    for (int i = tid; i < kRing; i += NT) s.ring[i] = (rbase + i < draws_len) ? __ldg(gdraws + rbase + i) : 0.0;
                    bsync<NT>();
                }
                const int64_t left = draws_len - cursor;
                avail = left >= 4 ? 4 : (int)(left < 0 ? 0 : left);
                if (B.spec_from >= 0 && cursor >= B.spec_from && avail > 3) avail = 3;   // no direction variate there
                const int o = (int)(cursor - rbase);
                e = s.ring[o]; uc = s.ring[o + 1]; ue = s.ring[o + 2]; ud = s.ring[o + 3];
            }
            if (avail < 3) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
            const int seq = (int)(n_done & 0x3fffffff) + 1;

            // ---- A1: chunked prefix sums (approximate order) + exact pairwise leaves ----
            const int c0 = tid * chunk;
            double cs = 0.0;
            for (int j = 0; j < chunk; ++j) { int i = c0 + j; if (i < n) cs = APS_ADD(cs, s.rates[i]); }
            double incl = cs;
            for (int o = 1; o < 32; o <<= 1) {
                double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = APS_ADD(incl, up);
            }
            double prev = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) prev = 0.0;
            if (lane == 31) s.wtot[wid] = incl;
            if (PHILOX && wid == 0) {
                // native mode: R is DEFINED as the 32-lane scan total of the chunk sums with the chunking rule of aps_math.h
                // (aps_native_total) — the quantity the specialised kernels' selection produces anyway; same bits in every kernel
                const int sh = aps_native_cs_shift(n), c_lo = lane << sh;
                double c = c_lo < n ? aps_native_chunk(s.rates, c_lo, 1 << sh, n) : 0.0;
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, c, o);
                    if (lane >= o) c = APS_ADD(c, up);
                }
                if (lane == 31) s.misc[2] = c;
            }
            // exact leaves (replay mode): 8 lanes per leaf == numpy's 8 strided accumulators
            for (int g = tid >> 3; !PHILOX && g < nnodes; g += NT / 8) {
                const unsigned gmask = 0xffu << (lane & 24);
                if (s.node_kind[g] != 0) continue;   // uniform within the 8-lane group
                const int k = tid & 7, start = s.node_a[g], len = s.node_b[g];
                double res;
                if (len < 8) {
                    res = 0.0;
                    for (int i = 0; i < len; ++i) res = APS_ADD(res, s.rates[start + i]);
                } else {
                    const int body = len - (len % 8);
                    double acc = s.rates[start + k];
                    for (int i = 8; i < body; i += 8) acc = APS_ADD(acc, s.rates[start + i + k]);
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 1));
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 2));
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 4));
                    res = acc;
                    for (int i = body; i < len; ++i) res = APS_ADD(res, s.rates[start + i]);
                }
                if (k == 0) s.node_val[g] = res;
            }
            bsync<NT>();  // BAR1

            // ---- A2: locate the selected particle; one thread forms R, tau and the clock ----
            double base = 0.0, atot = 0.0;
            for (int w2 = 0; w2 < NW; ++w2) { if (w2 == wid) base = atot; atot = APS_ADD(atot, s.wtot[w2]); }
            const double target = APS_MUL(uc, atot);
            const double excl = APS_ADD(base, prev), inc2 = APS_ADD(base, incl);
            if (excl <= target && target < inc2 && c0 < n) {
                // walk the chunk; interior edges use the same association as the scan inputs and the
                // last edge is pinned to inc2 so that neighbouring threads share their boundary exactly
                double run = 0.0, lo = excl, hi = excl; int sel = -1;
                for (int j = 0; j < chunk; ++j) {
                    int i = c0 + j; if (i >= n) break;
                    run = APS_ADD(run, s.rates[i]);
                    const bool last = (j == chunk - 1) || (i == n - 1);
                    hi = last ? inc2 : APS_ADD(base, APS_ADD(prev, run));
                    if (target < hi) { sel = i; break; }
                    lo = hi;
                }
                const double band = APS_MUL(guard, atot);
                if (sel < 0 || (target - lo) < band || (hi - target) < band) {
                    s.desc[D_EXACT] = 1;
                } else {
                    decode_and_apply(s, P, pad, crowd, sel, ue, ud, avail, seq, bcode, lut);
                }
            }
            if (tid == NT - 1) {
                if (!PHILOX)
                    for (int g = 0; g < nnodes; ++g)
                        if (s.node_kind[g] != 0) s.node_val[g] = APS_ADD(s.node_val[s.node_a[g]], s.node_val[s.node_b[g]]);
                const double R = PHILOX ? s.misc[2] : s.node_val[nnodes - 1];
                const double tau = APS_MUL(APS_DIV(1.0, R), e);
                const double tn = APS_ADD(t, tau);
                s.misc[X_R] = R; s.misc[X_TNEW] = tn;
                s.desc[D_BADR] = !(R > 0.0);
                s.desc[D_END] = tn > T;
                int nc = 0;
                if (!(tn > T) && obs_idx < M && next_obs <= tn) {
                    nc = 1;
                    while (obs_idx + nc < M && B.times_obs[obs_idx + nc] <= tn) ++nc;
                }
                s.desc[D_NCROSS] = nc;
            }
            bsync<NT>();  // BAR2

            if (s.desc[D_BADR]) { status = APS_RUN_EMPTY; break; }
            if (s.desc[D_EXACT] || s.desc[D_SEQ] != seq) {
                // exact serial selection (CLASS.py:359-360 literally); rare
                bsync<NT>();
                if (tid == 0) {
                    const double R = s.misc[X_R];
                    double acc = 0.0;
                    for (int i = 0; i < n; ++i) acc = APS_ADD(acc, APS_DIV(s.rates[i], R));
                    const double last = acc;
                    int sel = n - 1; acc = 0.0;
                    for (int i = 0; i < n; ++i) { acc = APS_ADD(acc, APS_DIV(s.rates[i], R)); if (APS_DIV(acc, last) > uc) { sel = i; break; } }
                    decode_and_apply(s, P, pad, crowd, sel, ue, ud, avail, seq, bcode, lut);
                    s.desc[D_EXACT] = 0;
                }
                ++n_guard;
                bsync<NT>();
            }
            if (s.desc[D_STOP]) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
            const int kind = s.desc[D_KIND], part = s.desc[D_PART], oldp = s.desc[D_OLD], newp = s.desc[D_NEW];
            const int ncross = s.desc[D_NCROSS], endflag = s.desc[D_END];
            const double tnew = s.misc[X_TNEW];
            if (B.trace && tid == 0 && n_done < B.trace_cap) {
                int32_t* tr = B.trace + ((size_t)rep * (size_t)B.trace_cap + (size_t)n_done) * 3;
                tr[0] = part; tr[1] = kind; tr[2] = (kind <= APS_EV_ACTIVE) ? newp : (kind == APS_EV_EXIT ? oldp : -1);
            }
            ++n_done;
            cursor += 3 + (kind < 2 ? 1 : 0);
            const int sg_old = s.desc[D_SG];                        // sign of the particle before the event
            if (kind == APS_EV_FLIP) S -= 2 * sg_old;
            if (kind == APS_EV_EXIT) {
                // np.delete(pos/sigma/bound, i) (CLASS.py:434-436): compact the particle arrays (and their rates)
                if (tid == 0 && B.exit_t && n_exit < B.exit_cap) {
                    B.exit_t[(size_t)rep * (size_t)B.exit_cap + n_exit] = t;          // clock before this step's tau (:425)
                    B.exit_pos[(size_t)rep * (size_t)B.exit_cap + n_exit] = oldp;
                }
                ++n_exit;
                S -= sg_old;
                for (int base = part; base < n - 1; base += NT) {
                    const int i = base + tid;
                    const bool ok = i < n - 1;
                    uint16_t vp = 0; int8_t vs = 0, vb = 0; double vr = 0.0;
                    if (ok) { vp = s.pos[i + 1]; vs = s.sigma[i + 1]; vb = s.bound[i + 1]; vr = s.rates[i + 1]; }
                    bsync<NT>();
                    if (ok) { s.pos[i] = vp; s.sigma[i] = vs; s.bound[i] = vb; s.rates[i] = vr; }
                    bsync<NT>();
                }
                --n;
                if (tid == 0) s.desc[D_NNODES] = (n > 0) ? build_sum_tree(n, s.node_a, s.node_b, s.node_kind, A.max_nodes) : 0;
                bsync<NT>();
                nnodes = s.desc[D_NNODES];
                chunk = (n + NT - 1) / NT;
                guard = A.guard_scale * 4.0 * (double)(n + 32) * 1.1102230246251565e-16;
            }
            t = tnew;
            if (endflag) { status = APS_RUN_DONE; break; }
            if (ncross > 0) {
                // rows get the field computed BEFORE this event (CLASS.py:512,525) and the state after it
                if ((B.record & APS_REC_MLOCAL) && B.obs_m_local) {
                    bsync<NT>();
                    if (tid == 0) lattice_delta(s, L, pad, kind, oldp, newp, sg_old, -1, bcode, lut);   // undo
                    bsync<NT>();
                    write_field(obs_idx, ncross);   // m_glob still holds the pre-event value
                    bsync<NT>();
                    if (tid == 0) lattice_delta(s, L, pad, kind, oldp, newp, sg_old, +1, bcode, lut);   // redo
                    bsync<NT>();
                }
                write_rows(obs_idx, ncross);
                obs_idx += ncross;
                if (obs_idx < M) next_obs = B.times_obs[obs_idx];
            }
            if (obs_idx >= M) { status = APS_RUN_DONE; break; }

            // ---- B: refresh the rates the event can have changed ----
            int wlo, whi;
            if (global_m) {
                if (kind == APS_EV_FLIP || kind == APS_EV_EXIT) {      // sum(sigma) or n changed: every flip rate changes
                    if (n > 0) m_glob = APS_DIV((double)S, (double)n);
                    wlo = 0; whi = L - 1;
                } else { wlo = (oldp < newp ? oldp : newp) - 1; whi = (oldp < newp ? newp : oldp) + 1; }
            } else {
                const int reach = r > 1 ? r : 1;
                wlo = (oldp < newp ? oldp : newp) - reach; whi = (oldp < newp ? newp : oldp) + reach;
            }
            if (periodic && (newp - oldp > 1 || oldp - newp > 1)) { wlo = 0; whi = L - 1; }   // a hop across the seam
            for (int i = tid; i < n; i += NT) {
                int p = s.pos[i];
                if ((p >= wlo && p <= whi) || (periodic && (p + L <= whi || p - L >= wlo))) s.rates[i] = particle_rate(s, P, pad, crowd, beta, m_glob, i, p, s.sigma[i], lut, b2);
            }
            bsync<NT>();  // BAR3
        }
    }

    // ---------------- epilogue ----------------
    if (tid == 0) {
        if (B.n_obs) B.n_obs[rep] = obs_idx;
        if (B.n_events) B.n_events[rep] = ev_base + n_done;
        if (B.t_end) B.t_end[rep] = t;
        if (B.status) B.status[rep] = status;
        if (B.n_guard) B.n_guard[rep] = n_guard;
        if (B.draws_used) B.draws_used[rep] = PHILOX ? 0 : cursor;
        if (B.n_end) B.n_end[rep] = n;
        if (B.n_exit) B.n_exit[rep] = n_exit;
    }
    bsync<NT>();
    if (B.bound_end) for (int i = tid; i < n; i += NT) B.bound_end[(size_t)rep * n_max + i] = s.bound[i];
    if (B.pos_end) for (int i = tid; i < n; i += NT) B.pos_end[(size_t)rep * n_max + i] = (int32_t)s.pos[i];
    if (B.sigma_end) for (int i = tid; i < n; i += NT) B.sigma_end[(size_t)rep * n_max + i] = s.sigma[i];
}

// compute_local_m_field for one lattice (API parity entry; one CTA).
static __global__ void field_kernel(aps_params P, const double* __restrict__ weights, const int32_t* __restrict__ cp,
                             const int32_t* __restrict__ cm, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char fk_raw[];
    const int L = P.L, r = P.radius, pad = r > 0 ? r : 0;
    double* w = reinterpret_cast<double*>(fk_raw);
    uint16_t* pk = reinterpret_cast<uint16_t*>(w + pad + 1);
    for (int i = threadIdx.x; i <= pad && r >= 0; i += blockDim.x) w[i] = weights[i];
    for (int i = threadIdx.x; i < L + 2 * pad; i += blockDim.x) {
        long long q = (long long)i - pad, per = 2LL * L;
        long long m = q % per; if (m < 0) m += per; if (m >= L) m = per - 1 - m;
        if (P.flags & APS_FLAG_PERIODIC) { m = q % L; if (m < 0) m += L; }
        pk[i] = (uint16_t)((cp[m] & 0xff) | ((cm[m] & 0xff) << 8));
    }
    __syncthreads();
    if (r < 0) {
        __shared__ long long tot[2];
        if (threadIdx.x == 0) {
            long long s = 0, t = 0;
            for (int l = 0; l < L; ++l) { s += cp[l] - cm[l]; t += cp[l] + cm[l]; }
            tot[0] = s; tot[1] = t;
        }
        __syncthreads();
        const double mg = APS_DIV((double)tot[0], (double)tot[1]);
        for (int l = threadIdx.x; l < L; l += blockDim.x) out[l] = mg;
        return;
    }
    for (int l = threadIdx.x; l < L; l += blockDim.x) out[l] = local_m(pk, w, pad, r, l);
}

}  // namespace aps
