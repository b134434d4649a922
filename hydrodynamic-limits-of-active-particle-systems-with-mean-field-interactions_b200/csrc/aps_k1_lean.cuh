// aps_k1_lean.cuh — K1 for the sweep configurations with HALF the shared-memory image of aps_k1_fast.cuh.
//
// Why: measured on B200 (profiles/r1_k1_lean.md) one replica-warp takes ~17-19 ms for config 2 whether 3 or 14 of them
// share an SM — the kernel is a latency chain with issue slots to spare — so throughput is (#replicas resident per SM)
// and run time is (#waves of CTAs) x (replica latency).  The 15.5 KB image of the fast kernel gives 14 replicas per SM
// = 2 waves for 4096 replicas; at <= 7.3 KB all 28 replicas of an SM are resident and the batch runs in ONE wave.
// What was dropped to get there (same arithmetic, bit-identical outputs; checked against the oracle like the others):
//   * the 3 KB product table: a tap pair is (multiplier[code_l + code_r]) * w_d with two 9-entry multiplier tables
//     (the products equal the table entries bit for bit);
//   * the site->particle map (2 KB): K = 1 hops cannot reorder particles, so when the initial positions are strictly
//     increasing (device Poisson init, any sorted input) the particles inside an update window are a contiguous index
//     range around the event particle, found with two ballots;  unsorted replicas are handed to aps_k1_fast.cuh
//     (status APS_RUN_RETRY_FAST, second launch);
//   * sigma / hop-flag / accumulator-slot arrays (1.5 KB): orientation is the lattice code at the particle's site,
//     hop flags are two neighbour reads, the accumulator slot is recomputed from the index;
//   * cached pairwise accumulators and chunk sums in shared memory: the 8 accumulators of a dirty leaf are re-summed
//     by 8 lanes in lock-step anyway, the chunk sum of lane c lives in a register of lane c.
// Instantiations (aps_fast.cu): RCAP = 21 (r <= 20) and RCAP = 81 (r <= 80, config 4) x
//   <NCAP = 512, WHO = false> as described (sorted inputs, n <= 488: <= 4 leaves in numpy's pairwise tree; 28 / 26 replicas per SM),
//   <NCAP = 1024, WHO = false> sorted inputs up to n = 968 (<= 8 leaves; 17 / 16 per SM; config 3),
//   <NCAP = 512 | 1024, WHO = true> which keep the site map and therefore take any particle order (22 / 15 per SM; the fast kernel: 10).
// Round 2 (profiles/r2_k1.md): native-mode R from the selection scan, site codes pre-scaled to table offsets, taps as kernel
// parameters, tree-of-16 chunk sums, the decided event broadcast by shuffles, size classes of one batch launched concurrently.
// Limits: single-warp CTAs, r + 1 <= RCAP, L + 2r <= LPCAP, n within the class; anything else falls through to the fast kernel.
#pragma once
#include <type_traits>

#include "aps_k1_fast.cuh"

namespace aps {

constexpr int APS_RUN_RETRY_FAST = 101;
constexpr int kLeanN = 512;            // capacity of the per-particle arrays
constexpr int kLeanNMax = 488;         // largest n whose numpy pairwise-sum tree has <= 4 leaves (489 has 5)
constexpr int kLeanRing = 32;          // 8 events of variate look-ahead

template <int RCAP>
struct LeanFixed {
    double ring[kLeanRing];
    double hop_tab[8];
    double2 mst[9];                    // multipliers (a_plus - a_minus, a_plus + a_minus) of a pair code: one 16-byte load
    double wtab[RCAP];
    double leafsum[8];
    double misc[6];                    // [0..3] P(left | diffusive hop) per free-neighbour pair, [4] guard band factor
    int32_t desc[16];
    int8_t node_a[16], node_b[16], node_kind[16], node_level[16], node_leaf[16];
    int16_t leaf_start[8], leaf_len[8];
    uint16_t list[64];                 // compacted particle list of an update window (site-map variant only)
    uint8_t dirty_c[32];
    uint8_t dirty_leaf[8];
};

// NCAP = 512, WHO = false: particles must come sorted (no site map), n <= 488, 28 replicas per SM.
// NCAP = 512, WHO = true: any particle order, n <= 488, 22 replicas per SM (second link of the chain for small replicas).
// NCAP = 1024, WHO = false: sorted particles, n <= 968 (<= 8 leaves), 17 replicas per SM (config 3: N = 900).
// NCAP = 1024, WHO = true: site->particle map kept (any particle order), n <= 968, 15 replicas per SM.
template <int RCAP, int LPCAP, int NCAP, bool WHO>
__host__ __device__ inline size_t k1_lean_smem_bytes() {
    return ((sizeof(LeanFixed<RCAP>) + 15) & ~(size_t)15) + (size_t)NCAP * 8 + (size_t)NCAP * 2 + (size_t)LPCAP + (WHO ? (size_t)LPCAP * 2 : 0);
}

template <bool PHILOX, int RCAP, int LPCAP, int NCAP, bool WHO>
__global__ void __launch_bounds__(32, NCAP <= 512 ? 28 : (WHO ? 15 : 17)) k1_lean_kernel(const __grid_constant__ K1Args A) {
    constexpr int kNMax = NCAP <= 512 ? kLeanNMax : 968;          // largest n with <= 4 (8) leaves in numpy's pairwise tree
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const aps_params& P = A.p;
    const aps_batch& B = A.b;
    const int rep = blockIdx.x;
    const int lane = threadIdx.x;
    const int L = P.L, r = P.radius, pad = r, n_max = B.n_max, M = B.M;
    const int n = B.n[rep];
    const double beta = B.beta[rep], T = P.T, D = P.rate_diffusion, lam = P.rate_active;

    if (A.only_retry == 2 && B.status[rep] != APS_RUN_RETRY_FAST) return;   // later launch of the chain: earlier rejects only
    if (n <= A.n_lo || n > A.n_hi) return;                                   // another size class of this batch owns the replica
    LeanFixed<RCAP>& F = *reinterpret_cast<LeanFixed<RCAP>*>(smem_raw);
    unsigned char* dyn = smem_raw + ((sizeof(LeanFixed<RCAP>) + 15) & ~(size_t)15);
    double* const rates = reinterpret_cast<double*>(dyn); dyn += (size_t)NCAP * 8;
    uint16_t* const pos = reinterpret_cast<uint16_t*>(dyn); dyn += (size_t)NCAP * 2;
    uint16_t* const who = reinterpret_cast<uint16_t*>(dyn); dyn += WHO ? (size_t)LPCAP * 2 : 0;
    uint8_t* const code = dyn;
    // site codes are stored PRE-SCALED by the size of a multiplier-table entry (16 B): the sum of two codes is the byte offset
    // into F.mst, so a tap costs one 3-input add instead of an add and a scaled-index multiply (ncu round 2: 7.5 -> 6.5 per tap)
    constexpr int kCP = 16, kCM = 48;                               // '+' particle, '-' particle (1 and 3 entries)
    auto mst_at = [&](int byte_off) { return *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(F.mst) + byte_off); };

    // ---------------- prologue ----------------
    for (int i = lane; i < L + 2 * pad; i += 32) code[i] = 0;
    if (WHO) for (int i = lane; i < L; i += 32) who[i] = 0xFFFFu;
    for (int j = lane; j <= r; j += 32) F.wtab[j] = B.weights[j];
    if (lane < 9) { const int am = lane / 3, ap = lane - am * 3; F.mst[lane] = make_double2((double)(ap - am), (double)(ap + am)); }
    if (lane < 16) F.desc[lane] = 0;
    F.dirty_c[lane] = 1;
    if (lane < 8) F.dirty_leaf[lane] = 1;
    if (lane < 8) {
        const double dz = APS_MUL(D, 0.0);
        F.hop_tab[lane] = APS_ADD(APS_ADD((lane & 1) ? D : dz, (lane & 2) ? D : dz), (lane & 4) ? lam : 0.0);
        // P(left | diffusive hop) per pair of free neighbours (CLASS.py:392): rl / (rl + rr); NaN when neither is free (never used:
        // the diffusive threshold is 0 then)
        if (lane < 4) F.misc[lane] = APS_DIV((lane & 1) ? D : dz, APS_ADD((lane & 1) ? D : dz, (lane & 2) ? D : dz));
    }
    __syncwarp();
    int S = 0;
    bool unsorted = false;
    if (n > 0 && n <= kNMax) {
        const int32_t* gp = B.pos0 + (size_t)rep * n_max;
        const int8_t* gs = B.sigma0 + (size_t)rep * n_max;
        int part = 0, bad = 0;
        for (int i = lane; i < n; i += 32) {
            const int p = gp[i], sg = gs[i];
            pos[i] = (uint16_t)p;
            part += sg;
            if (!WHO && i + 1 < n && gp[i + 1] <= p) bad = 1;         // needs strictly increasing positions (K = 1, sorted)
            code[pad + p] = (uint8_t)(sg == 1 ? kCP : kCM);              // distinct sites when valid; garbage otherwise (we bail out)
            if (WHO) who[p] = (uint16_t)i;
        }
        if (WHO) {                                                    // two particles on one site: only the last writer is in who[]
            __syncwarp();
            for (int i = lane; i < n; i += 32) bad |= (who[pos[i]] != (uint16_t)i);
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        S = part;
        unsorted = __any_sync(0xffffffffu, bad);
    }
    if (unsorted || n > kNMax || n == 0) {
        if (lane == 0) {
            if (n == 0) {
                if (B.n_obs) B.n_obs[rep] = B.obs_start ? B.obs_start[rep] : 0;
                if (B.n_events) B.n_events[rep] = B.ev_start ? B.ev_start[rep] : 0;
                if (B.t_end) B.t_end[rep] = B.t_start ? B.t_start[rep] : 0.0;
                if (B.n_guard) B.n_guard[rep] = 0;
                if (B.draws_used) B.draws_used[rep] = 0;
                if (B.n_end) B.n_end[rep] = 0;
                if (B.n_exit && !B.ev_start) B.n_exit[rep] = 0;
                B.status[rep] = APS_RUN_EMPTY;
            } else B.status[rep] = APS_RUN_RETRY_FAST;
        }
        return;
    }
    __syncwarp();
    // reflect images of the halo (sites within `pad` of a wall)
    for (int i = lane; i < n; i += 32) {
        const int p = pos[i];
        const uint8_t c = code[pad + p];
        if (p < pad) code[pad - 1 - p] = c;
        if (p >= L - pad) code[pad + 2 * L - 1 - p] = c;
    }
    if (lane == 0) {
        // numpy pairwise tree (<= 4 leaves for n <= 512), built in the still unused rates area
        int32_t* na = reinterpret_cast<int32_t*>(rates); int32_t* nb = na + 16; int32_t* nk = nb + 16; int32_t* lv = nk + 16;
        const int nn = build_sum_tree(n, na, nb, nk, 16);
        int nl = 0;
        for (int g = 0; g < nn; ++g) {
            F.node_kind[g] = (int8_t)nk[g];
            if (nk[g] == 0) { F.leaf_start[nl] = (int16_t)na[g]; F.leaf_len[nl] = (int16_t)nb[g]; F.node_leaf[g] = (int8_t)nl++; F.node_a[g] = 0; F.node_b[g] = 0; }
            else { F.node_leaf[g] = -1; F.node_a[g] = (int8_t)na[g]; F.node_b[g] = (int8_t)nb[g]; }
        }
        lv[nn - 1] = 0;
        int maxlev = 0;
        for (int g = nn - 1; g >= 0; --g) if (nk[g] == 1) {
            const int l2 = lv[g] + 1;
            lv[na[g]] = l2; lv[nb[g]] = l2;
            if (l2 > maxlev) maxlev = l2;
        }
        for (int g = 0; g < nn; ++g) F.node_level[g] = (int8_t)lv[g];
        F.desc[D_NNODES] = nn; F.desc[12] = nl; F.desc[13] = maxlev;
    }
    __syncwarp();
    const int nnodes = F.desc[D_NNODES], nleaf = F.desc[12], maxlev = F.desc[13];
    // loop invariants of the clock in registers: node `lane` of the tree, the leaf of this lane's 8-lane group
    const bool have_node = lane < nnodes;
    const int nd_kind = have_node ? F.node_kind[lane] : 0, nd_lev = have_node ? F.node_level[lane] : -1;
    const int nd_a = (have_node && nd_kind) ? F.node_a[lane] : 0, nd_b = (have_node && nd_kind) ? F.node_b[lane] : 0;
    const int nd_leaf = (have_node && !nd_kind) ? F.node_leaf[lane] : 0;
    const int my_g = lane >> 3;
    const int g_start = my_g < nleaf ? F.leaf_start[my_g] : 0, g_len = my_g < nleaf ? F.leaf_len[my_g] : 0;
    const int g2_start = my_g + 4 < nleaf ? F.leaf_start[my_g + 4] : 0, g2_len = my_g + 4 < nleaf ? F.leaf_len[my_g + 4] : 0;   // NCAP = 1024 only
    const int ls1 = nleaf > 1 ? F.leaf_start[1] : 0x7fff, ls2 = nleaf > 2 ? F.leaf_start[2] : 0x7fff, ls3 = nleaf > 3 ? F.leaf_start[3] : 0x7fff;
    const int ls4 = nleaf > 4 ? F.leaf_start[4] : 0x7fff, ls5 = nleaf > 5 ? F.leaf_start[5] : 0x7fff;
    const int ls6 = nleaf > 6 ? F.leaf_start[6] : 0x7fff, ls7 = nleaf > 7 ? F.leaf_start[7] : 0x7fff;
    int cs_shift = 4;
    while (((n + (1 << cs_shift) - 1) >> cs_shift) > 32) ++cs_shift;
    const int CS = 1 << cs_shift;
    const int nchunks = (n + CS - 1) >> cs_shift;
    double my_cs = 0.0;                                           // cached sum of chunk `lane`

    int64_t n_done = 0;
    const int64_t ev_base = B.ev_start ? B.ev_start[rep] : 0;
    int64_t cursor = 0, n_guard = 0, rbase = -(int64_t)kLeanRing - 8;
    const int64_t draws_len = PHILOX ? 0 : (B.draw_off[rep + 1] - B.draw_off[rep]);
    const double* gdraws = PHILOX ? nullptr : (B.draws + B.draw_off[rep]);
    const uint32_t k0 = PHILOX ? (uint32_t)B.seeds[rep] : 0u, k1 = PHILOX ? (uint32_t)(B.seeds[rep] >> 32) : 0u;
    double t = B.t_start ? B.t_start[rep] : 0.0;
    int obs_idx = B.obs_start ? B.obs_start[rep] : 0;
    int status = APS_RUN_DONE;
    const int64_t max_events = B.max_events > 0 ? B.max_events : 0x7fffffffffffffffLL;

    // local magnetisation at site p: scipy's symmetric tap order, products formed as multiplier * w (== the table entries).
    // The multipliers are 0, +-1, +-2 (K = 1: a pair of sites holds at most two particles), so multiplier * w is EXACT and
    // fma(multiplier, w, acc) rounds once, to the same value as the reference's separate multiply and add — one DFMA per
    // accumulator and tap instead of DMUL + DADD.  For the radius of the capacity class (r = RCAP - 1, every shipped sweep
    // with sigma = 0.005) the tap loop has a compile-time trip count: immediate offsets, loads hoisted ahead of the chain.
    // (`hot` selects the unrolled form: only the rate refresh of the event loop uses it, to keep the loop body in the i-cache)
    auto local_m = [&](int p, auto hot) {
        const uint8_t* c = code + pad + p;
        const double w0 = F.wtab[r];
        const double2 m0 = mst_at(c[0]);
        double sc = APS_MUL(m0.x, w0), tc = APS_MUL(m0.y, w0);
        if (decltype(hot)::value && r == RCAP - 1 && RCAP <= 84 && A.wt_valid) {
#pragma unroll
            for (int jj = -(RCAP - 1); jj < 0; ++jj) {
                const double2 mm = mst_at((int)c[jj] + (int)c[-jj]);
                const double wj = A.wt[RCAP <= 84 ? RCAP - 1 + jj : 0];  // kernel parameter: a constant-bank operand of the DFMA, no load
                sc = __fma_rn(mm.x, wj, sc);
                tc = __fma_rn(mm.y, wj, tc);
            }
        } else {
#pragma unroll 4
            for (int jj = -r; jj < 0; ++jj) {
                const double2 mm = mst_at((int)c[jj] + (int)c[-jj]);
                const double wj = F.wtab[r + jj];
                sc = __fma_rn(mm.x, wj, sc);
                tc = __fma_rn(mm.y, wj, tc);
            }
        }
        double m = 0.0;
        if (tc > 0.0) m = APS_DIV(sc, tc);
        return m < -1.0 ? -1.0 : (m > 1.0 ? 1.0 : m);
    };
    auto write_rows = [&](int first, int count) {
        for (int m = first; m < first + count; ++m) {
            const size_t row = (size_t)rep * (size_t)M + (size_t)m;
            if ((B.record & APS_REC_COUNTS) && B.obs_cp && B.obs_cm) {
                int8_t* ocp = B.obs_cp + row * (size_t)L; int8_t* ocm = B.obs_cm + row * (size_t)L;
                for (int l = lane; l < L; l += 32) { const uint8_t v = code[pad + l]; ocp[l] = (int8_t)(v == kCP); ocm[l] = (int8_t)(v == kCM); }
            }
            if ((B.record & APS_REC_POS) && B.obs_pos) {
                int32_t* op = B.obs_pos + row * (size_t)n_max;
                for (int i = lane; i < n; i += 32) op[i] = (int32_t)pos[i];
            }
            if (B.obs_sigma_sum && lane == 0) B.obs_sigma_sum[row] = S;
            if (B.obs_n && lane == 0) B.obs_n[row] = n;
            if (B.obs_bound) { int8_t* ob = B.obs_bound + row * (size_t)n_max; for (int i = lane; i < n; i += 32) ob[i] = 0; }
        }
    };
    auto write_field = [&](int first, int count) {
        if (!((B.record & APS_REC_MLOCAL) && B.obs_m_local)) return;
        for (int l = lane; l < L; l += 32) {
            const double m = local_m(l, std::false_type{});
            for (int mm = first; mm < first + count; ++mm)
                B.obs_m_local[((size_t)rep * (size_t)M + (size_t)mm) * (size_t)L + l] = m;
        }
    };
    auto hop_flags = [&](int p, int cd) {
        const bool l_free = (p > 0) && code[pad + p - 1] == 0, r_free = (p < L - 1) && code[pad + p + 1] == 0;
        return (l_free ? 1 : 0) | (r_free ? 2 : 0) | ((cd == kCP && r_free) ? 4 : 0);
    };
    // full rate of particle i (CLASS.py:351)
    auto refresh = [&](int i, auto hot) {
        const int p = pos[i];
        const int cd = code[pad + p];
        const double sgd = cd == kCP ? 1.0 : -1.0;
        const double h = F.hop_tab[hop_flags(p, cd)];
        const double m = local_m(p, hot);
        rates[i] = APS_ADD(h, aps_exp(APS_MUL(APS_MUL(-beta, sgd), m)));
        F.dirty_c[i >> cs_shift] = 1;
        if (!PHILOX) {                                  // the pairwise leaves are only summed in replay mode (see the clock below)
            int lf = (i >= ls1) + (i >= ls2) + (i >= ls3);
            if (NCAP > 512) lf += (i >= ls4) + (i >= ls5) + (i >= ls6) + (i >= ls7);
            F.dirty_leaf[lf] = 1;
        }
    };
    auto code_put = [&](int x, int delta) { code_add(code, L, pad, x, delta); };
    auto tree16 = [&](const double* q) {                            // aps_tree16 without the i < n tests: rates[n ..] are 0.0
        const double2* v = reinterpret_cast<const double2*>(q);
        const double2 v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3], v4 = v[4], v5 = v[5], v6 = v[6], v7 = v[7];
        const double a0 = APS_ADD(APS_ADD(v0.x, v0.y), APS_ADD(v1.x, v1.y)), a1 = APS_ADD(APS_ADD(v2.x, v2.y), APS_ADD(v3.x, v3.y));
        const double a2 = APS_ADD(APS_ADD(v4.x, v4.y), APS_ADD(v5.x, v5.y)), a3 = APS_ADD(APS_ADD(v6.x, v6.y), APS_ADD(v7.x, v7.y));
        return APS_ADD(APS_ADD(a0, a1), APS_ADD(a2, a3));
    };
    for (int i = n + lane; i < NCAP; i += 32) rates[i] = 0.0;      // zero padding of the last chunk (tree16)

    for (int i = lane; i < n; i += 32) refresh(i, std::false_type{});
    if (obs_idx == 0 && M > 0) { write_field(0, 1); write_rows(0, 1); obs_idx = 1; }
    __syncwarp();
    double next_obs = (obs_idx < M) ? B.times_obs[obs_idx] : 0.0;
    if (lane == 0) F.misc[4] = A.guard_scale * 4.0 * (double)(n + 32) * 1.1102230246251565e-16;   // read by the deciding lane only
    __syncwarp();

    while (true) {
        if (!(t < T)) { status = APS_RUN_DONE; break; }
        if (n_done >= max_events) { status = APS_RUN_MAX_EVENTS; break; }
        int avail; double e, uc, ue, ud;
        if (PHILOX) {
            const int slot = (int)(n_done & 7);
            if (slot == 0) {
                if (lane < 8) {
                    const uint64_t ev = (uint64_t)(ev_base + n_done + lane);
                    const aps_u32x4 a = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_A, 0u, k0, k1);
                    const aps_u32x4 b = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_B, 0u, k0, k1);
                    F.ring[4 * lane + 0] = -aps_log(APS_SUB(1.0, aps_u53(a.v[0], a.v[1])));
                    F.ring[4 * lane + 1] = aps_u53(a.v[2], a.v[3]);
                    F.ring[4 * lane + 2] = aps_u53(b.v[0], b.v[1]);
                    F.ring[4 * lane + 3] = aps_u53(b.v[2], b.v[3]);
                }
                __syncwarp();
            }
            avail = 4;
            e = F.ring[4 * slot]; uc = F.ring[4 * slot + 1]; ue = F.ring[4 * slot + 2]; ud = F.ring[4 * slot + 3];
        } else {
            if (cursor + 4 > rbase + kLeanRing) {
                __syncwarp();
                rbase = cursor;
                F.ring[lane] = (rbase + lane < draws_len) ? __ldg(gdraws + rbase + lane) : 0.0;
                __syncwarp();
            }
            const int64_t left = draws_len - cursor;
            avail = left >= 4 ? 4 : (int)(left < 0 ? 0 : left);
            if (B.spec_from >= 0 && cursor >= B.spec_from && avail > 3) avail = 3;
            const int o = (int)(cursor - rbase);
            e = F.ring[o]; uc = F.ring[o + 1]; ue = F.ring[o + 2]; ud = F.ring[o + 3];
        }
        if (avail < 3) { status = APS_RUN_DRAWS_EXHAUSTED; break; }

        // The event of particle `sel`, decided and applied by ONE lane; it is handed to the warp in two packed words that the
        // caller broadcasts with shuffles (round 2: the shared-memory descriptor + __syncwarp round trip cost ~45 instructions per event):
        //   pa = part | kind << 10 | (sigma == +1) << 12 | stop << 13 | (guard band hit: redo exactly) << 14,  pb = old site | new site << 16
        auto decode_apply = [&](int sel, int& pa, int& pb) {
            const int p = pos[sel];
            const int cd = code[pad + p], sg = cd == kCP ? 1 : -1;
            const int hf = hop_flags(p, cd);
            const double v = APS_MUL(ue, rates[sel]);
            // thresholds (rl + rr) and (rl + rr) + ra with rl, rr in {D, D*0}, ra in {lam, 0}: entries of the hop-rate table
            const double diff_thresh = F.hop_tab[hf & 3], act_thresh = F.hop_tab[hf];
            int kind, newp = p;
            if (v < diff_thresh) {
                if (avail < 4) { pa = 1 << 13; pb = 0; return; }
                if (ud < F.misc[hf & 3]) { kind = APS_EV_DIFF_LEFT; newp = clampi(p - 1, 0, L - 1); }
                else { kind = APS_EV_DIFF_RIGHT; newp = clampi(p + 1, 0, L - 1); }
            } else if (v < act_thresh) { kind = APS_EV_ACTIVE; newp = clampi(p + (sg == 1), 0, L - 1); }
            else kind = APS_EV_FLIP;
            if (kind == APS_EV_FLIP) code_put(p, sg == 1 ? kCM - kCP : kCP - kCM);
            else if (newp != p) {
                pos[sel] = (uint16_t)newp; code_put(p, -cd); code_put(newp, cd);
                if (WHO) { who[p] = 0xFFFFu; who[newp] = (uint16_t)sel; }
            }
            pa = sel | (kind << 10) | ((sg == 1) << 12);
            pb = p | (newp << 16);
        };

        // ---- selection: chunk sums (dirty ones re-summed, cached in a register), warp scan, walk of the winning chunk ----
        double r_scan, inv_r_scan = 0.0;                // total of the scan = R in native mode (aps_math.h, aps_native_total), and 1/R
        int pa = 0, pb = 0;                             // the packed event (decode_apply)
        bool exact;                                     // warp-uniform: the scan could not decide, lane 0 redoes the selection serially
        {
            if (lane < nchunks && F.dirty_c[lane]) {
                // aps_native_chunk (aps_math.h) on the zero-padded image: adjacent pairwise tree of 16 per block, 16-byte loads
                const int c0 = lane << cs_shift;
                double cs = tree16(rates + c0);
                for (int b = c0 + 16; b < c0 + CS && b < n; b += 16) cs = APS_ADD(cs, tree16(rates + b));
                my_cs = cs; F.dirty_c[lane] = 0;
            }
            double incl = lane < nchunks ? my_cs : 0.0;
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = APS_ADD(incl, up);
            }
            double prev = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) prev = 0.0;
            const double atot = __shfl_sync(0xffffffffu, incl, 31);
            r_scan = atot;
            if (PHILOX) inv_r_scan = APS_DIV(1.0, atot);    // every lane, straight after the scan: the division's latency hides behind
                                                            // the chunk walk instead of sitting in the clock lane's serial section
            const double target = APS_MUL(uc, atot);
            const unsigned wmask = __ballot_sync(0xffffffffu, lane < nchunks && prev <= target && target < incl);
            exact = (__popc(wmask) != 1);
            if (!exact) {
                const int wl = __ffs(wmask) - 1;
                const double prev_w = __shfl_sync(0xffffffffu, prev, wl), inc_w = __shfl_sync(0xffffffffu, incl, wl);
                const int i = (wl << cs_shift) + lane;
                const bool mine = lane < CS && i < n;
                double run = mine ? rates[i] : 0.0;
                for (int o = 1; o < CS; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, run, o);
                    if (lane >= o) run = APS_ADD(run, up);
                }
                const bool last = mine && (lane == CS - 1 || i == n - 1);
                const double hi = last ? inc_w : APS_ADD(prev_w, run);
                double lo = __shfl_up_sync(0xffffffffu, hi, 1);
                if (lane == 0) lo = prev_w;
                const unsigned smask = __ballot_sync(0xffffffffu, mine && lo <= target && target < hi);
                if (__popc(smask) != 1) exact = true;
                else {
                    const int src = __ffs(smask) - 1;
                    if (lane == src) {
                        const double band = APS_MUL(F.misc[4], atot);
                        if ((target - lo) < band || (hi - target) < band) pa = 1 << 14;
                        else decode_apply(i, pa, pb);
                    }
                    pa = __shfl_sync(0xffffffffu, pa, src); pb = __shfl_sync(0xffffffffu, pb, src);
                    exact = (pa & (1 << 14)) != 0;
                }
            }
        }
        // ---- clock.  Replay mode: numpy's pairwise sum exactly (8 lanes per leaf, <= 4 leaves; CLASS.py:352), the clock the reference
        //      reports.  Native mode: R is the total of the selection scan (defined in aps_math.h; the oracle does the same), so the
        //      leaf sums, their dirty flags and the tree drop out of the per-event chain (19 % of it, ncu profiles/r2_k1.md). ----
        double R;
        {
            double val = r_scan;
            if (!PHILOX) {
#pragma unroll
            for (int half = 0; half < (NCAP > 512 ? 2 : 1); ++half) {
                const int gg = my_g + 4 * half;
                const int gs = half ? g2_start : g_start, gl = half ? g2_len : g_len;
                if (gg < nleaf && F.dirty_leaf[gg]) {                 // uniform within the 8-lane group
                    const unsigned gmask = 0xffu << (lane & 24);
                    const int k = lane & 7;
                    double res;
                    if (gl < 8) {
                        res = 0.0;
                        for (int i = 0; i < gl; ++i) res = APS_ADD(res, rates[gs + i]);
                    } else {
                        const int body = gl - (gl & 7);
                        double acc = rates[gs + k];
                        for (int i = 8; i < body; i += 8) acc = APS_ADD(acc, rates[gs + i + k]);
                        acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 1));
                        acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 2));
                        acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 4));
                        res = acc;
                        for (int i = body; i < gl; ++i) res = APS_ADD(res, rates[gs + i]);
                    }
                    __syncwarp(gmask);
                    if (k == 0) { F.leafsum[gg] = res; F.dirty_leaf[gg] = 0; }
                }
            }
            __syncwarp();
            val = (have_node && !nd_kind) ? F.leafsum[nd_leaf] : 0.0;
            for (int lev = maxlev - 1; lev >= 0; --lev) {
                const double va = __shfl_sync(0xffffffffu, val, nd_a), vb = __shfl_sync(0xffffffffu, val, nd_b);
                if (nd_kind && nd_lev == lev) val = APS_ADD(va, vb);
            }
            }
            if (!PHILOX) val = __shfl_sync(0xffffffffu, val, nnodes - 1);    // the root of the pairwise tree
            R = val;
        }
        // every lane advances the clock itself (same operands in all lanes: no broadcast)
        const double tau = APS_MUL(PHILOX ? inv_r_scan : APS_DIV(1.0, R), e);
        const double tnew = APS_ADD(t, tau);
        const bool endflag = tnew > T;
        int ncross = 0;
        if (!endflag && obs_idx < M && next_obs <= tnew) {
            ncross = 1;
            while (obs_idx + ncross < M && B.times_obs[obs_idx + ncross] <= tnew) ++ncross;
        }
        if (!(R > 0.0)) { status = APS_RUN_EMPTY; break; }
        if (exact) {
            if (lane == 0) {
                double acc = 0.0;
                for (int i = 0; i < n; ++i) acc = APS_ADD(acc, APS_DIV(rates[i], R));
                const double last = acc;
                int sel = n - 1; acc = 0.0;
                for (int i = 0; i < n; ++i) { acc = APS_ADD(acc, APS_DIV(rates[i], R)); if (APS_DIV(acc, last) > uc) { sel = i; break; } }
                decode_apply(sel, pa, pb);
            }
            pa = __shfl_sync(0xffffffffu, pa, 0); pb = __shfl_sync(0xffffffffu, pb, 0);
            ++n_guard;
        }
        __syncwarp();                                                        // the event's writes to pos / code are visible to all lanes
        if (pa & (1 << 13)) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
        const int kind = (pa >> 10) & 3, part = pa & 1023, oldp = pb & 0xffff, newp = (int)((unsigned)pb >> 16);
        const int sg_now = (pa & (1 << 12)) ? 1 : -1;                // orientation of the particle BEFORE the event
        if (lane == 0 && n_done < B.trace_cap) {            // trace_cap is 0 without a trace buffer (launch_k1)
            int32_t* tr = B.trace + ((size_t)rep * (size_t)B.trace_cap + (size_t)n_done) * 3;
            tr[0] = part; tr[1] = kind; tr[2] = (kind == APS_EV_FLIP) ? -1 : newp;
        }
        ++n_done;
        cursor += 3 + (kind < 2 ? 1 : 0);
        if (kind == APS_EV_FLIP) S -= 2 * sg_now;
        t = tnew;
        if (endflag) { status = APS_RUN_DONE; break; }
        if (ncross > 0) {
            if ((B.record & APS_REC_MLOCAL) && B.obs_m_local) {
                const int cd = sg_now == 1 ? kCP : kCM;
                __syncwarp();
                if (lane == 0) {   // undo on the code array only: the recorded field is the pre-event one
                    if (kind == APS_EV_FLIP) code_put(oldp, sg_now == 1 ? kCP - kCM : kCM - kCP);
                    else if (newp != oldp) { code_put(newp, -cd); code_put(oldp, cd); }
                }
                __syncwarp();
                write_field(obs_idx, ncross);
                __syncwarp();
                if (lane == 0) {   // redo
                    if (kind == APS_EV_FLIP) code_put(oldp, sg_now == 1 ? kCM - kCP : kCP - kCM);
                    else if (newp != oldp) { code_put(newp, cd); code_put(oldp, -cd); }
                }
                __syncwarp();
            }
            write_rows(obs_idx, ncross);
            obs_idx += ncross;
            if (obs_idx < M) next_obs = B.times_obs[obs_idx];
        }
        if (obs_idx >= M) { status = APS_RUN_DONE; break; }

        // ---- refresh the rates inside the window ----
        if (WHO) {                                                     // any particle order: site->particle map + ballot compaction
            const int reach = r > 1 ? r : 1;
            const int mn = oldp < newp ? oldp : newp, mx = oldp < newp ? newp : oldp;
            int wlo = mn - reach, whi = mx + reach;
            if (wlo < 0) wlo = 0;
            if (whi > L - 1) whi = L - 1;
            int count = 0;
            for (int s0 = wlo; s0 <= whi; s0 += 32) {
                const int site = s0 + lane;
                const unsigned v = (site <= whi) ? who[site] : 0xFFFFu;
                const unsigned mask = __ballot_sync(0xffffffffu, v != 0xFFFFu);
                if (v != 0xFFFFu) F.list[count + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)v;
                count += __popc(mask);
                if (count >= 32 || s0 + 32 > whi) {                    // the list holds at most 64 entries
                    __syncwarp();
                    for (int j0 = 0; j0 < count; j0 += 32) { const int j = j0 + lane; if (j < count) refresh(F.list[j], std::true_type{}); }
                    __syncwarp();
                    count = 0;
                }
            }
        } else {   // sorted particles: the window is a contiguous index range around `part`
            const int reach = r > 1 ? r : 1;
            const int mn = oldp < newp ? oldp : newp, mx = oldp < newp ? newp : oldp;
            const int wlo = mn - reach, whi = mx + reach;
            // lowest particle index inside the window: blocks of 32 candidates part-31 .. part, part-63 .. part-32, ... (the test is
            // monotone in the lane because the positions are sorted, so the first set lane of the last non-empty block is the answer;
            // one block for r = 20, up to three for r = 80)
            int ilo = part;
            for (int base = part - 31;; base -= 32) {
                const int ia = base + lane;
                const bool in_a = ia >= 0 && (int)pos[ia >= 0 ? ia : 0] >= wlo;
                const unsigned ma = __ballot_sync(0xffffffffu, in_a);
                if (ma) ilo = base + (__ffs(ma) - 1);
                if (ma != 0xffffffffu || base <= 0) break;
            }
            for (int i0 = ilo; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const bool in = i < n && (int)pos[i < n ? i : n - 1] <= whi;
                if (in) refresh(i, std::true_type{});
                if (!__all_sync(0xffffffffu, in)) break;
            }
        }
        __syncwarp();
    }

    if (lane == 0) {
        if (B.n_obs) B.n_obs[rep] = obs_idx;
        if (B.n_events) B.n_events[rep] = ev_base + n_done;
        if (B.t_end) B.t_end[rep] = t;
        B.status[rep] = status;
        if (B.n_guard) B.n_guard[rep] = n_guard;
        if (B.draws_used) B.draws_used[rep] = PHILOX ? 0 : cursor;
        if (B.n_end) B.n_end[rep] = n;
        if (B.n_exit && !B.ev_start) B.n_exit[rep] = 0;
    }
    __syncwarp();
    if (B.bound_end) for (int i = lane; i < n; i += 32) B.bound_end[(size_t)rep * n_max + i] = 0;
    if (B.pos_end) for (int i = lane; i < n; i += 32) B.pos_end[(size_t)rep * n_max + i] = (int32_t)pos[i];
    if (B.sigma_end) for (int i = lane; i < n; i += 32) B.sigma_end[(size_t)rep * n_max + i] = (int8_t)(code[pad + pos[i]] == kCP ? 1 : -1);
}

}  // namespace aps
