// aps_k1_fast.cuh — K1 specialised for the configurations every sweep driver ships:
// site_capacity K = 1, local Gaussian field (radius r < L), no crowding, n <= 2048.
// Same arithmetic as aps_k1.cuh (bit-identical outputs; both are checked against the oracle) but
// restructured around what the first ncu capture showed (profiles/r1_k1_baseline_ncu.md: 2100
// warp-instructions per event, half-empty lanes, O(n) rescans, parameter reloads):
//   * the selection prefix sums and numpy's pairwise accumulators are kept in shared memory and only
//     the entries whose rates changed (dirty flags set by the rate update) are re-summed;
//   * the particles inside the update window are found through a site->particle map and compacted
//     with ballots, so the filter loop runs with dense lanes in one warp pass;
//   * hop contributions are cached per particle, so only the particles next to the changed sites
//     re-read occupancies; the lattice is one uint8 radix code per site (0 empty, 1 '+', 3 '-');
//   * the pairwise tree is combined level-synchronously with shuffles instead of a serial loop;
//   * small fixed-size arrays sit at compile-time offsets, the rest behind 7 base offsets.
// Replicas whose initial state violates K = 1 (the reference accepts such states) are flagged
// APS_RUN_RETRY_GENERIC and re-run by the generic kernel in a second launch.
#pragma once
#include "aps_k1.cuh"

namespace aps {

constexpr int kFastLeafCap = 8;       // n <= 968 (kFastMaxN); larger replicas are re-run by the generic kernel
constexpr int kFastNodeCap = 16;
constexpr int kFastRing = 64;         // doubles of variate look-ahead (16 events in native mode)
constexpr int kFastMaxN = 968;        // largest n whose numpy pairwise-sum tree has <= 8 leaves / 15 nodes (969 has 9 leaves)
constexpr int APS_RUN_RETRY_GENERIC = 100;

struct FastFixed {                    // compile-time-offset part of the shared-memory image
    double ring[kFastRing];
    double acc[kFastLeafCap * 8];
    double leafsum[kFastLeafCap];
    double csum[64];
    double hop_tab[8];
    double wtot[2];
    double misc[8];
    int32_t desc[16];
    int32_t leaf_start[kFastLeafCap], leaf_len[kFastLeafCap];
    int32_t node_a[kFastNodeCap], node_b[kFastNodeCap], node_kind[kFastNodeCap], node_level[kFastNodeCap], node_leaf[kFastNodeCap];
    uint16_t list[2][96];             // per-warp compacted particle lists (first two warps)
    uint8_t dirty_c[64];
    uint8_t dirty_a[kFastLeafCap * 8];
    uint8_t dirty_leaf[kFastLeafCap];
};

// rcap/ncap/lpcap = capacity class (compile-time offsets) or, with rcap == 0, the actual sizes
__host__ __device__ inline size_t k1_fast_smem_bytes(int L, int n_max, int radius, int rcap, int ncap_, int lpcap) {
    const bool st = rcap > 0;
    const size_t ncap = st ? (size_t)ncap_ : (size_t)n_max;
    size_t b = (sizeof(FastFixed) + 15) & ~(size_t)15;
    b += (st ? (size_t)rcap : (size_t)(radius + 1)) * 9 * 16;          // lut
    b += ncap * 8;                                                     // rates
    b += ((ncap + 7) & ~(size_t)7) * 2;                                // hop flags, accumulator slots
    b += st ? (size_t)lpcap : ((((size_t)L + 2 * (size_t)radius) + 7) & ~(size_t)7);    // codes (uint8, with halo)
    b += st ? (size_t)lpcap * 2 : ((((size_t)L * 2) + 7) & ~(size_t)7);                 // who
    b += (ncap * 2 + 7) & ~(size_t)7;                                  // pos
    b += (ncap + 7) & ~(size_t)7;                                      // sigma
    return b;
}

__device__ __forceinline__ void code_add(uint8_t* code, int L, int pad, int x, int delta) {
    code[pad + x] = (uint8_t)(code[pad + x] + delta);
    if (x < pad || x >= L - pad) {                    // within `pad` of a wall (rare): the reflect images of the site
        if (x < pad) code[pad - 1 - x] = (uint8_t)(code[pad - 1 - x] + delta);
        if (x >= L - pad) code[pad + 2 * L - 1 - x] = (uint8_t)(code[pad + 2 * L - 1 - x] + delta);
    }
}

__device__ __forceinline__ double fast_local_m(const uint8_t* code, const double2* lut, int pad, int r, int p) {
    const uint8_t* c = code + pad + p;
    double2 v = lut[r * 9 + c[0]];
    double sc = v.x, tc = v.y;
    const double2* row = lut;
#pragma unroll 4
    for (int jj = -r; jj < 0; ++jj) {
        int idx = (int)c[jj] + (int)c[-jj];
        double2 t2 = row[idx];
        row += 9;
        sc = APS_ADD(sc, t2.x);
        tc = APS_ADD(tc, t2.y);
    }
    double m = 0.0;
    if (tc > 0.0) m = APS_DIV(sc, tc);
    m = m < -1.0 ? -1.0 : (m > 1.0 ? 1.0 : m);
    return m;
}

// hop rates for K = 1 from the code array (CLASS.py:276-319): left/right free <=> neighbour empty and inside
__device__ __forceinline__ void fast_hops(const uint8_t* code, int pad, int L, double D, double lam, int p, int sg,
                                          double& rl, double& rr, double& ra) {
    const bool l_free = (p > 0) && code[pad + p - 1] == 0;
    const bool r_free = (p < L - 1) && code[pad + p + 1] == 0;
    const double dz = APS_MUL(D, 0.0);
    rl = l_free ? D : dz;
    rr = r_free ? D : dz;
    ra = (sg == 1 && r_free) ? lam : 0.0;
}

// RCAP/NCAP/LPCAP > 0: capacity class with compile-time shared-memory offsets (r+1 <= RCAP, n_max <= NCAP,
// L+2r <= LPCAP); RCAP == 0: offsets computed at run time from the actual sizes.
template <int NT, bool PHILOX, int RCAP, int NCAP, int LPCAP>
__global__ void __launch_bounds__(NT, 1024 / NT) k1_fast_kernel(const __grid_constant__ K1Args A) {
    constexpr bool STATIC = RCAP > 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const aps_params& P = A.p;
    const aps_batch& B = A.b;
    const int rep = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int L = P.L, r = P.radius, pad = r, n_max = B.n_max, M = B.M;
    const int n = B.n[rep];
    const double beta = B.beta[rep], T = P.T, D = P.rate_diffusion, lam = P.rate_active;

    if (A.only_retry == 2 && B.status[rep] != 101) return;   // second launch after aps_k1_lean.cuh: its rejects only
    FastFixed& F = *reinterpret_cast<FastFixed*>(smem_raw);
    unsigned char* dyn = smem_raw + ((sizeof(FastFixed) + 15) & ~(size_t)15);
    const size_t ncap = STATIC ? (size_t)NCAP : (size_t)n_max;
    double2* const lut = reinterpret_cast<double2*>(dyn); dyn += (STATIC ? (size_t)RCAP : (size_t)(r + 1)) * 9 * 16;
    double* const rates = reinterpret_cast<double*>(dyn); dyn += ncap * 8;
    uint8_t* const hopf = dyn; dyn += (ncap + 7) & ~(size_t)7;
    uint8_t* const accslot = dyn; dyn += (ncap + 7) & ~(size_t)7;
    uint8_t* const code = dyn; dyn += STATIC ? (size_t)LPCAP : ((((size_t)L + 2 * (size_t)r) + 7) & ~(size_t)7);
    uint16_t* const who = reinterpret_cast<uint16_t*>(dyn); dyn += STATIC ? (size_t)LPCAP * 2 : ((((size_t)L * 2) + 7) & ~(size_t)7);
    uint16_t* const pos = reinterpret_cast<uint16_t*>(dyn); dyn += (ncap * 2 + 7) & ~(size_t)7;
    int8_t* const sigma = reinterpret_cast<int8_t*>(dyn);

    // ---------------- prologue ----------------
    for (int i = tid; i < L + 2 * pad; i += NT) code[i] = 0;
    for (int i = tid; i < L; i += NT) who[i] = 0xFFFFu;
    for (int e = tid; e < (r + 1) * 9; e += NT) {
        int j = e / 9, idx = e - j * 9, am = idx / 3, ap = idx - am * 3;
        double wj = B.weights[j];
        lut[e] = make_double2(APS_MUL((double)(ap - am), wj), APS_MUL((double)(ap + am), wj));
    }
    if (tid < 16) F.desc[tid] = 0;
    for (int i = tid; i < 64; i += NT) F.dirty_c[i] = 1;
    if (tid < 8) {   // (rl + rr) + ra for every combination of (left free, right free, active hop possible)
        const double dz = APS_MUL(D, 0.0);
        F.hop_tab[tid] = APS_ADD(APS_ADD((tid & 1) ? D : dz, (tid & 2) ? D : dz), (tid & 4) ? lam : 0.0);
    }
    for (int i = tid; i < kFastLeafCap * 8; i += NT) F.dirty_a[i] = 1;
    if (tid < kFastLeafCap) F.dirty_leaf[tid] = 1;
    bsync<NT>();
    int S = 0;
    int bad = 0;
    {
        const int32_t* gp = B.pos0 + (size_t)rep * n_max;
        const int8_t* gs = B.sigma0 + (size_t)rep * n_max;
        int part = 0;
        for (int i = tid; i < n; i += NT) {
            int p = gp[i], sg = gs[i];
            pos[i] = (uint16_t)p; sigma[i] = (int8_t)sg;
            part += sg;
            who[p] = (uint16_t)i;   // K = 1: one particle per site; a collision is detected below
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) F.ring[wid] = (double)part;
        bsync<NT>();
        for (int w2 = 0; w2 < NW; ++w2) S += (int)F.ring[w2];
        // a site claimed by two particles keeps only the last writer in who[] -> detect and fall back
        for (int i = tid; i < n; i += NT) bad |= (who[pos[i]] != (uint16_t)i);
        bad = (NT == 32) ? __any_sync(0xffffffffu, bad) : __syncthreads_or(bad);
    }
    if (bad || n > kFastMaxN || n == 0) {
        if (tid == 0) {
            if (n == 0) {
                if (B.n_obs) B.n_obs[rep] = B.obs_start ? B.obs_start[rep] : 0;
                if (B.n_events) B.n_events[rep] = B.ev_start ? B.ev_start[rep] : 0;
                if (B.t_end) B.t_end[rep] = B.t_start ? B.t_start[rep] : 0.0;
                if (B.n_guard) B.n_guard[rep] = 0;
                if (B.draws_used) B.draws_used[rep] = 0;
                if (B.n_end) B.n_end[rep] = 0;
                if (B.n_exit && !B.ev_start) B.n_exit[rep] = 0;
                B.status[rep] = APS_RUN_EMPTY;
            } else B.status[rep] = APS_RUN_RETRY_GENERIC;
        }
        return;
    }
    for (int i = tid; i < n; i += NT) {
        const int p = pos[i];
        code_add(code, L, pad, p, sigma[i] == 1 ? 1 : 3);    // distinct sites: no two threads touch the same cell
    }
    if (tid == 0) {
        // numpy pairwise tree -> leaves (in order) + post-order node program with levels
        int nn = build_sum_tree(n, F.node_a, F.node_b, F.node_kind, kFastNodeCap);
        int nl = 0;
        for (int g = 0; g < nn; ++g) {
            if (F.node_kind[g] == 0) { F.leaf_start[nl] = F.node_a[g]; F.leaf_len[nl] = F.node_b[g]; F.node_leaf[g] = nl++; }
            else F.node_leaf[g] = -1;
        }
        F.node_level[nn - 1] = 0;
        int maxlev = 0;
        for (int g = nn - 1; g >= 0; --g) if (F.node_kind[g] == 1) {
            int lv = F.node_level[g] + 1;
            F.node_level[F.node_a[g]] = lv; F.node_level[F.node_b[g]] = lv;
            if (lv > maxlev) maxlev = lv;
        }
        F.desc[D_NNODES] = nn; F.desc[12] = nl; F.desc[13] = maxlev;
    }
    bsync<NT>();
    // per particle: which pairwise accumulator its rate feeds (bits 0-2 lane k, bits 3-6 leaf, bit 7 = tail element)
    for (int i = tid; i < n; i += NT) {
        int lf = 0;
        while (lf + 1 < F.desc[12] && i >= F.leaf_start[lf + 1]) ++lf;
        const int j = i - F.leaf_start[lf], len = F.leaf_len[lf];
        const bool body = len >= 8 && j < len - (len & 7);
        accslot[i] = (uint8_t)((lf << 3) | (body ? (j & 7) : 0) | (body ? 0 : 0x80));
    }
    const int nnodes = F.desc[D_NNODES], nleaf = F.desc[12], maxlev = F.desc[13];
    const int cs_shift = aps_native_cs_shift(n);     // chunks of 16 * 2^k rates, at most one per lane of the selection warp; the
                                                     // same rule in every K1 kernel: the scan total is R in native mode (aps_math.h)
    const int CS = 1 << cs_shift;
    const int nchunks = (n + CS - 1) >> cs_shift;

    int64_t n_done = 0;
    const int64_t ev_base = B.ev_start ? B.ev_start[rep] : 0;
    int64_t cursor = 0, n_guard = 0, rbase = -(int64_t)kFastRing - 8;
    const int64_t draws_len = PHILOX ? 0 : (B.draw_off[rep + 1] - B.draw_off[rep]);
    const double* gdraws = PHILOX ? nullptr : (B.draws + B.draw_off[rep]);
    const uint32_t k0 = PHILOX ? (uint32_t)B.seeds[rep] : 0u, k1 = PHILOX ? (uint32_t)(B.seeds[rep] >> 32) : 0u;
    double t = B.t_start ? B.t_start[rep] : 0.0;
    int obs_idx = B.obs_start ? B.obs_start[rep] : 0;
    int status = APS_RUN_DONE;
    const int64_t max_events = B.max_events > 0 ? B.max_events : 0x7fffffffffffffffLL;

    auto write_rows = [&](int first, int count) {
        for (int m = first; m < first + count; ++m) {
            size_t row = (size_t)rep * (size_t)M + (size_t)m;
            if ((B.record & APS_REC_COUNTS) && B.obs_cp && B.obs_cm) {
                int8_t* ocp = B.obs_cp + row * (size_t)L; int8_t* ocm = B.obs_cm + row * (size_t)L;
                for (int l = tid; l < L; l += NT) { uint8_t v = code[pad + l]; ocp[l] = (int8_t)(v == 1); ocm[l] = (int8_t)(v == 3); }
            }
            if ((B.record & APS_REC_POS) && B.obs_pos) {
                int32_t* op = B.obs_pos + row * (size_t)n_max;
                for (int i = tid; i < n; i += NT) op[i] = (int32_t)pos[i];
            }
            if (B.obs_sigma_sum && tid == 0) B.obs_sigma_sum[row] = S;
            if (B.obs_n && tid == 0) B.obs_n[row] = n;
            if (B.obs_bound) { int8_t* ob = B.obs_bound + row * (size_t)n_max; for (int i = tid; i < n; i += NT) ob[i] = 0; }
        }
    };
    auto write_field = [&](int first, int count) {
        if (!((B.record & APS_REC_MLOCAL) && B.obs_m_local)) return;
        for (int l = tid; l < L; l += NT) {
            double m = fast_local_m(code, lut, pad, r, l);
            for (int mm = first; mm < first + count; ++mm)
                B.obs_m_local[((size_t)rep * (size_t)M + (size_t)mm) * (size_t)L + l] = m;
        }
    };
    // full rate of particle i (CLASS.py:351) and its cached hop part
    auto refresh = [&](int i, bool redo_hop) {
        const int p = pos[i], sg = sigma[i];
        int hf;
        if (redo_hop) {
            const bool l_free = (p > 0) && code[pad + p - 1] == 0, r_free = (p < L - 1) && code[pad + p + 1] == 0;
            hf = (l_free ? 1 : 0) | (r_free ? 2 : 0) | ((sg == 1 && r_free) ? 4 : 0);
            hopf[i] = (uint8_t)hf;
        } else hf = hopf[i];
        const double h = F.hop_tab[hf];
        const double m = fast_local_m(code, lut, pad, r, p);
        const double cv = aps_exp(APS_MUL(APS_MUL(-beta, (double)sg), m));
        rates[i] = APS_ADD(h, cv);
        F.dirty_c[i >> cs_shift] = 1;
        if (!PHILOX) {                                  // pairwise accumulators: replay mode only
            const int as = accslot[i];
            F.dirty_leaf[(as >> 3) & 15] = 1;
            if (!(as & 0x80)) F.dirty_a[as] = 1;
        }
    };

    for (int i = tid; i < n; i += NT) refresh(i, true);
    if (obs_idx == 0 && M > 0) { write_field(0, 1); write_rows(0, 1); obs_idx = 1; }
    bsync<NT>();
    double next_obs = (obs_idx < M) ? B.times_obs[obs_idx] : 0.0;
    const double guard = A.guard_scale * 4.0 * (double)(n + 32) * 1.1102230246251565e-16;

    while (true) {
        if (!(t < T)) { status = APS_RUN_DONE; break; }
        if (n_done >= max_events) { status = APS_RUN_MAX_EVENTS; break; }
        int avail; double e, uc, ue, ud;
        if (PHILOX) {
            const int slot = (int)(n_done & 15);
            if (slot == 0) {
                if (tid < 16) {
                    uint64_t ev = (uint64_t)(ev_base + n_done + tid);
                    aps_u32x4 a = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_A, 0u, k0, k1);
                    aps_u32x4 b = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_B, 0u, k0, k1);
                    F.ring[4 * tid + 0] = -aps_log(APS_SUB(1.0, aps_u53(a.v[0], a.v[1])));
                    F.ring[4 * tid + 1] = aps_u53(a.v[2], a.v[3]);
                    F.ring[4 * tid + 2] = aps_u53(b.v[0], b.v[1]);
                    F.ring[4 * tid + 3] = aps_u53(b.v[2], b.v[3]);
                }
                bsync<NT>();
            }
            avail = 4;
            e = F.ring[4 * slot]; uc = F.ring[4 * slot + 1]; ue = F.ring[4 * slot + 2]; ud = F.ring[4 * slot + 3];
        } else {
            if (cursor + 4 > rbase + kFastRing) {
                bsync<NT>();
                rbase = cursor;
                for (int i = tid; i < kFastRing; i += NT) F.ring[i] = (rbase + i < draws_len) ? __ldg(gdraws + rbase + i) : 0.0;
                bsync<NT>();
            }
            const int64_t left = draws_len - cursor;
            avail = left >= 4 ? 4 : (int)(left < 0 ? 0 : left);
            if (B.spec_from >= 0 && cursor >= B.spec_from && avail > 3) avail = 3;
            const int o = (int)(cursor - rbase);
            e = F.ring[o]; uc = F.ring[o + 1]; ue = F.ring[o + 2]; ud = F.ring[o + 3];
        }
        if (avail < 3) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
        const int seq = (int)(n_done & 0x3fffffff) + 1;

        auto decode_apply = [&](int sel) {
            const int p = pos[sel], sg = sigma[sel];
            double rl, rr, ra;
            fast_hops(code, pad, L, D, lam, p, sg, rl, rr, ra);
            const double v = APS_MUL(ue, rates[sel]);
            const double diff_thresh = APS_ADD(rl, rr), act_thresh = APS_ADD(diff_thresh, ra);
            int kind, newp = p;
            if (v < diff_thresh) {
                if (avail < 4) { F.desc[D_STOP] = 1; F.desc[D_SEQ] = seq; return; }
                if (ud < APS_DIV(rl, APS_ADD(rl, rr))) { kind = APS_EV_DIFF_LEFT; newp = clampi(p - 1, 0, L - 1); }
                else { kind = APS_EV_DIFF_RIGHT; newp = clampi(p + 1, 0, L - 1); }
            } else if (v < act_thresh) { kind = APS_EV_ACTIVE; newp = clampi(p + (sg == 1), 0, L - 1); }
            else kind = APS_EV_FLIP;
            const int cd = sg == 1 ? 1 : 3;
            if (kind == APS_EV_FLIP) {
                sigma[sel] = (int8_t)(-sg);
                code_add(code, L, pad, p, sg == 1 ? 2 : -2);
            } else if (newp != p) {
                pos[sel] = (uint16_t)newp;
                code_add(code, L, pad, p, -cd); code_add(code, L, pad, newp, cd);
                who[p] = 0xFFFFu; who[newp] = (uint16_t)sel;
            }
            F.desc[D_PART] = sel; F.desc[D_KIND] = kind; F.desc[D_OLD] = p; F.desc[D_NEW] = newp;
            F.desc[D_STOP] = 0; F.desc[D_SEQ] = seq;
        };

        // ---- warp 0 = SELECTION warp: chunk sums (<= 32 chunks, dirty ones re-summed), warp scan, cooperative
        //      walk of the winning chunk, event decode + apply.  The prefix sums are an approximation of the
        //      reference's serial cumsum in any association; the guard band below makes the decision exact. ----
        if (wid == 0) {
            double cs = 0.0;
            if (lane < nchunks) {
                if (F.dirty_c[lane]) {
                    cs = aps_native_chunk(rates, lane << cs_shift, CS, n);      // tree of 16 per block (aps_math.h)
                    F.csum[lane] = cs; F.dirty_c[lane] = 0;
                } else cs = F.csum[lane];
            }
            double incl = cs;
            for (int o = 1; o < 32; o <<= 1) {
                double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = APS_ADD(incl, up);
            }
            double prev = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) prev = 0.0;
            const double atot = __shfl_sync(0xffffffffu, incl, 31);
            const double target = APS_MUL(uc, atot);
            const unsigned wmask = __ballot_sync(0xffffffffu, lane < nchunks && prev <= target && target < incl);
            bool exact = (__popc(wmask) != 1);
            if (!exact) {
                const int wl = __ffs(wmask) - 1;
                const double prev_w = __shfl_sync(0xffffffffu, prev, wl), inc_w = __shfl_sync(0xffffffffu, incl, wl);
                const int i = (wl << cs_shift) + lane;
                const bool mine = lane < CS && i < n;
                double run = mine ? rates[i] : 0.0;
                for (int o = 1; o < CS; o <<= 1) {
                    double up = __shfl_up_sync(0xffffffffu, run, o);
                    if (lane >= o) run = APS_ADD(run, up);
                }
                const bool last = mine && (lane == CS - 1 || i == n - 1);
                const double hi = last ? inc_w : APS_ADD(prev_w, run);
                double lo = __shfl_up_sync(0xffffffffu, hi, 1);
                if (lane == 0) lo = prev_w;
                const unsigned smask = __ballot_sync(0xffffffffu, mine && lo <= target && target < hi);
                if (__popc(smask) != 1) exact = true;
                else if (lane == __ffs(smask) - 1) {
                    const double band = APS_MUL(guard, atot);
                    if ((target - lo) < band || (hi - target) < band) F.desc[D_EXACT] = 1;
                    else decode_apply(i);
                }
            }
            if (exact && lane == 0) F.desc[D_EXACT] = 1;
            if (PHILOX && lane == 0) {                  // native mode: R = total of the selection scan (aps_math.h, aps_native_total)
                const double R = atot;
                const double tau = APS_MUL(APS_DIV(1.0, R), e);
                const double tn = APS_ADD(t, tau);
                F.misc[X_R] = R; F.misc[X_TNEW] = tn;
                F.desc[D_BADR] = !(R > 0.0);
                F.desc[D_END] = tn > T;
                int nc = 0;
                if (!(tn > T) && obs_idx < M && next_obs <= tn) {
                    nc = 1;
                    while (obs_idx + nc < M && B.times_obs[obs_idx + nc] <= tn) ++nc;
                }
                F.desc[D_NCROSS] = nc;
            }
        }
        // ---- last warp = CLOCK warp (replay mode): numpy's pairwise sum exactly (dirty accumulators re-summed, 8 lanes per
        //      leaf, level-synchronous tree), R, tau, the event clock and the observation-crossing count ----
        if (!PHILOX && wid == NW - 1) {
            for (int g = lane >> 3; g < nleaf; g += 4) {
                if (!F.dirty_leaf[g]) continue;                       // uniform within the 8-lane group
                const unsigned gmask = 0xffu << (lane & 24);
                const int k = lane & 7, start = F.leaf_start[g], len = F.leaf_len[g];
                double res;
                if (len < 8) {
                    res = 0.0;
                    for (int i = 0; i < len; ++i) res = APS_ADD(res, rates[start + i]);
                } else {
                    const int body = len - (len & 7);
                    double acc;
                    if (F.dirty_a[g * 8 + k]) {
                        acc = rates[start + k];
                        for (int i = 8; i < body; i += 8) acc = APS_ADD(acc, rates[start + i + k]);
                        F.acc[g * 8 + k] = acc; F.dirty_a[g * 8 + k] = 0;
                    } else acc = F.acc[g * 8 + k];
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 1));
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 2));
                    acc = APS_ADD(acc, __shfl_xor_sync(gmask, acc, 4));
                    res = acc;
                    for (int i = body; i < len; ++i) res = APS_ADD(res, rates[start + i]);
                }
                __syncwarp(gmask);
                if (k == 0) { F.leafsum[g] = res; F.dirty_leaf[g] = 0; }
            }
            __syncwarp();
            // level-synchronous evaluation of the tree: lane g holds node g
            const bool have = lane < nnodes;
            const int kd = have ? F.node_kind[lane] : 0, lv = have ? F.node_level[lane] : -1;
            const int ca = (have && kd) ? F.node_a[lane] : 0, cb = (have && kd) ? F.node_b[lane] : 0;
            double val = (have && !kd) ? F.leafsum[F.node_leaf[lane]] : 0.0;
            for (int lev = maxlev - 1; lev >= 0; --lev) {
                const double va = __shfl_sync(0xffffffffu, val, ca), vb = __shfl_sync(0xffffffffu, val, cb);
                if (kd && lv == lev) val = APS_ADD(va, vb);
            }
            if (lane == nnodes - 1) {
                const double R = val;
                const double tau = APS_MUL(APS_DIV(1.0, R), e);
                const double tn = APS_ADD(t, tau);
                F.misc[X_R] = R; F.misc[X_TNEW] = tn;
                F.desc[D_BADR] = !(R > 0.0);
                F.desc[D_END] = tn > T;
                int nc = 0;
                if (!(tn > T) && obs_idx < M && next_obs <= tn) {
                    nc = 1;
                    while (obs_idx + nc < M && B.times_obs[obs_idx + nc] <= tn) ++nc;
                }
                F.desc[D_NCROSS] = nc;
            }
        }
        bsync<NT>();  // BAR2

        if (F.desc[D_BADR]) { status = APS_RUN_EMPTY; break; }
        if (F.desc[D_EXACT] || F.desc[D_SEQ] != seq) {
            bsync<NT>();
            if (tid == 0) {
                const double R = F.misc[X_R];
                double acc = 0.0;
                for (int i = 0; i < n; ++i) acc = APS_ADD(acc, APS_DIV(rates[i], R));
                const double last = acc;
                int sel = n - 1; acc = 0.0;
                for (int i = 0; i < n; ++i) { acc = APS_ADD(acc, APS_DIV(rates[i], R)); if (APS_DIV(acc, last) > uc) { sel = i; break; } }
                decode_apply(sel);
                F.desc[D_EXACT] = 0;
            }
            ++n_guard;
            bsync<NT>();
        }
        if (F.desc[D_STOP]) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
        const int kind = F.desc[D_KIND], part = F.desc[D_PART], oldp = F.desc[D_OLD], newp = F.desc[D_NEW];
        const int ncross = F.desc[D_NCROSS], endflag = F.desc[D_END];
        const double tnew = F.misc[X_TNEW];
        if (B.trace && tid == 0 && n_done < B.trace_cap) {
            int32_t* tr = B.trace + ((size_t)rep * (size_t)B.trace_cap + (size_t)n_done) * 3;
            tr[0] = part; tr[1] = kind; tr[2] = (kind == APS_EV_FLIP) ? -1 : newp;
        }
        ++n_done;
        cursor += 3 + (kind < 2 ? 1 : 0);
        int sg_old = sigma[part];
        if (kind == APS_EV_FLIP) { S += 2 * sg_old; sg_old = -sg_old; }
        t = tnew;
        if (endflag) { status = APS_RUN_DONE; break; }
        if (ncross > 0) {
            if ((B.record & APS_REC_MLOCAL) && B.obs_m_local) {
                const int cd = sg_old == 1 ? 1 : 3;
                bsync<NT>();
                if (tid == 0) {   // undo on the code array only
                    if (kind == APS_EV_FLIP) code_add(code, L, pad, oldp, sg_old == 1 ? -2 : 2);
                    else if (newp != oldp) { code_add(code, L, pad, newp, -cd); code_add(code, L, pad, oldp, cd); }
                }
                bsync<NT>();
                write_field(obs_idx, ncross);
                bsync<NT>();
                if (tid == 0) {   // redo
                    if (kind == APS_EV_FLIP) code_add(code, L, pad, oldp, sg_old == 1 ? 2 : -2);
                    else if (newp != oldp) { code_add(code, L, pad, newp, cd); code_add(code, L, pad, oldp, -cd); }
                }
                bsync<NT>();
            }
            write_rows(obs_idx, ncross);
            obs_idx += ncross;
            if (obs_idx < M) next_obs = B.times_obs[obs_idx];
        }
        if (obs_idx >= M) { status = APS_RUN_DONE; break; }

        // ---- B: refresh the rates inside the window; each of the first two warps compacts and handles
        //         alternate 32-site strips through the site->particle map ----
        const int reach = r > 1 ? r : 1;
        const int mn = oldp < newp ? oldp : newp, mx = oldp < newp ? newp : oldp;
        int wlo = mn - reach, whi = mx + reach;
        if (wlo < 0) wlo = 0;
        if (whi > L - 1) whi = L - 1;
        // narrow windows (<= 64 sites, i.e. r <= 30) are handled by warp 0 alone: one compaction, dense lanes;
        // wide windows are compacted by both warps, which then take alternate 32-particle blocks
        const int NB = (NW > 1 && whi - wlo + 1 > 64) ? 2 : 1;
        if (wid < NB) {
            int count = 0, done = 0;
            uint16_t* lst = F.list[wid];
            for (int s0 = wlo; s0 <= whi; s0 += 32) {
                const int site = s0 + lane;
                const unsigned v = (site <= whi) ? who[site] : 0xFFFFu;
                const unsigned mask = __ballot_sync(0xffffffffu, v != 0xFFFFu);
                if (v != 0xFFFFu) lst[count + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)v;
                count += __popc(mask);
                if (count >= 64 || s0 + 32 > whi) {     // process what is queued (the list holds at most 96 entries)
                    __syncwarp();
                    for (int j0 = 32 * ((done + wid) % NB); j0 < count; j0 += 32 * NB) {
                        const int j = j0 + lane;
                        if (j < count) { const int i = lst[j]; const int p = pos[i]; refresh(i, p >= mn - 1 && p <= mx + 1); }
                    }
                    done += (count + 31) / 32;
                    __syncwarp();
                    count = 0;
                }
            }
        }
        bsync<NT>();  // BAR3
    }

    if (tid == 0) {
        if (B.n_obs) B.n_obs[rep] = obs_idx;
        if (B.n_events) B.n_events[rep] = ev_base + n_done;
        if (B.t_end) B.t_end[rep] = t;
        B.status[rep] = status;
        if (B.n_guard) B.n_guard[rep] = n_guard;
        if (B.draws_used) B.draws_used[rep] = PHILOX ? 0 : cursor;
        if (B.n_end) B.n_end[rep] = n;
        if (B.n_exit && !B.ev_start) B.n_exit[rep] = 0;
    }
    bsync<NT>();
    if (B.bound_end) for (int i = tid; i < n; i += NT) B.bound_end[(size_t)rep * n_max + i] = 0;
    if (B.pos_end) for (int i = tid; i < n; i += NT) B.pos_end[(size_t)rep * n_max + i] = (int32_t)pos[i];
    if (B.sigma_end) for (int i = tid; i < n; i += NT) B.sigma_end[(size_t)rep * n_max + i] = sigma[i];
}

}  // namespace aps
