// aps_obs.cuh — K4: on-device observables.  Nothing per-event or per-site crosses PCIe for an
// ensemble: the observation rows K1 wrote (int8 counts, int32 positions) are reduced here.
//
//   expand_kernel    counts -> rho_plus / rho_minus / total density rows (+ np.var of total),
//                    ParticleSystem.empirical_densities_from_particles (CLASS.py:198-214) and
//                    run()'s per-row bookkeeping (:489-507, :518-535)
//   reduce_kernel    the sweep drivers' per-run reducers
//                    compute_v_eff_and_window      sweep_beta.py:123-162
//                    compute_rho_eff               sweep_beta.py:165-194
//                    compute_blocking_probability  sweep_beta.py:197-229
//                    compute_mean_magnetizatoin    sweep_beta.py:316-319
//                    compute_D_eff_active          sweep_beta.py:500-525
//   profile_kernel   ensemble sums (over the replicas of one grid point and a row window) of the
//                    density / magnetisation profiles — the payload of the NCCL allreduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aps.h"
#include "../../include/aps_math.h"

namespace aps {

// ---- block reduction helpers (sum of doubles / ints over a CTA, result broadcast) ----------
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T tot = 0;
    for (int w = 0; w < nw; ++w) tot += scratch[w];
    return tot;
}

// numpy pairwise sum of a shared-memory array, executed by one thread (used for np.var parity;
// rows are short and this runs M times per replica, far off the hot path)
__device__ inline double pairwise_serial(const double* a, int n) {
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r = APS_ADD(r, a[i]); return r; }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] = APS_ADD(r[k], a[i + k]);
        double res = APS_ADD(APS_ADD(APS_ADD(r[0], r[1]), APS_ADD(r[2], r[3])), APS_ADD(APS_ADD(r[4], r[5]), APS_ADD(r[6], r[7])));
        for (; i < n; ++i) res = APS_ADD(res, a[i]);
        return res;
    }
    int n2 = n / 2; n2 -= n2 % 8;
    return APS_ADD(pairwise_serial(a, n2), pairwise_serial(a + n2, n - n2));
}

// One CTA per (replica, row).  rho = counts / (max(1,n)*dx)  (CLASS.py:208-213), total = rho_p+rho_m,
// var = np.var(total) with numpy's evaluation order (mean by pairwise sum, then pairwise sum of squares).
__global__ void expand_kernel(aps_expand_args a) {
    extern __shared__ double sh[];   // [L] when var is requested
    const int rep = blockIdx.y, m = blockIdx.x;
    if (m >= a.n_obs[rep]) return;
    const int L = a.L;
    const size_t row = (size_t)rep * a.M + m;
    const int8_t* cp = a.obs_cp + row * L;
    const int8_t* cm = a.obs_cm + row * L;
    const int n = a.obs_n ? a.obs_n[row] : a.n[rep];
    const double denom = APS_MUL((double)(n > 1 ? n : 1), a.dx);
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        double rp = APS_DIV((double)cp[l], denom), rm = APS_DIV((double)cm[l], denom);
        double tot = APS_ADD(rp, rm);
        if (a.rho_p) a.rho_p[row * L + l] = rp;
        if (a.rho_m) a.rho_m[row * L + l] = rm;
        if (a.total) a.total[row * L + l] = tot;
        if (a.var) sh[l] = tot;
    }
    if (a.var) {
        __syncthreads();
        __shared__ double mean_s;
        if (threadIdx.x == 0) mean_s = APS_DIV(pairwise_serial(sh, L), (double)L);
        __syncthreads();
        const double mean = mean_s;
        for (int l = threadIdx.x; l < L; l += blockDim.x) { double x = APS_SUB(sh[l], mean); sh[l] = APS_MUL(x, x); }
        __syncthreads();
        if (threadIdx.x == 0) a.var[row] = APS_DIV(pairwise_serial(sh, L), (double)L);
    }
}

// x_grid = np.linspace(0, 1, L): arange(L) * (1/(L-1)), last point forced to 1.0
__device__ __forceinline__ double xgrid(int l, int L, double step) { return (l == L - 1 && L > 1) ? 1.0 : APS_MUL((double)l, step); }

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// v1 of the reducer (round 1): per-site double arithmetic.  Kept as the A/B reference of reduce_kernel below
// (aps_debug_set_reduce_impl(1), tests/test_dropin_gpu.py); not launched otherwise.
// One CTA per replica; every warp owns whole observation rows (lane-strided over the lattice, shuffle
// reductions only), so the row loops run without block barriers.  Shared: per-row scalars [M] x 3.
__global__ void reduce_kernel_v1(aps_reduce_args a) {
    extern __shared__ double sh[];
    const int rep = blockIdx.x, tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const int M = a.M, L = a.L, n = a.n[rep], nobs = a.n_obs[rep];
    double* mean_x = sh;            // [M]
    double* frac_b = sh + M;        // [M]
    double* aux = sh + 2 * M;       // [M] scratch (S_k of the MSD fit)
    double* scr = sh + 3 * M;       // [32] reduction scratch
    double* out = a.out + (size_t)rep * APS_RED_N;
    const double step = L > 1 ? APS_DIV(1.0, (double)(L - 1)) : 0.0;
    const double dxg = L > 1 ? APS_SUB(xgrid(1, L, step), xgrid(0, L, step)) : 0.0;
    const double denom = APS_MUL((double)(n > 1 ? n : 1), a.dx);

    // density of a site holding c particles of one species: c / (max(1,n)*dx) (CLASS.py:208-213), tabulated once
    double* dtab = scr + 32;        // [128]: the observation rows hold int8 counts, so the table covers every value
    for (int c = tid; c < 128; c += NT) dtab[c] = APS_DIV((double)c, denom);
    __syncthreads();
    auto dens = [&](int c) { return dtab[c & 127]; };
    const bool vec = (L % 4) == 0;  // rows start 4-byte aligned: four sites per load
    // ---- per-row sums over the lattice (rows never reached are all-zero in the reference) ----
    for (int m = wid; m < M; m += NW) {
        double sx = 0.0, st = 0.0, sb = 0.0;
        if (m < nobs) {
            const int8_t* cp = a.obs_cp + ((size_t)rep * M + m) * L;
            const int8_t* cm = a.obs_cm + ((size_t)rep * M + m) * L;
            auto site = [&](int l, int p, int q) {
                double d = APS_ADD(dens(p), dens(q));
                double x = xgrid(l, L, step);
                st += d; sx += d * x;
                if (x >= a.boundary_xmin) sb += d;
            };
            if (vec) {
                const uint32_t* cp4 = reinterpret_cast<const uint32_t*>(cp);
                const uint32_t* cm4 = reinterpret_cast<const uint32_t*>(cm);
                for (int w = lane; w < L / 4; w += 32) {
                    const uint32_t pw = cp4[w], qw = cm4[w];
                    if (pw | qw) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int pj = (pw >> (8 * j)) & 0xff, qj = (qw >> (8 * j)) & 0xff;
                            if (pj | qj) site(4 * w + j, pj, qj);
                        }
                    }
                }
            } else {
                for (int l = lane; l < L; l += 32) { const int pj = cp[l], qj = cm[l]; if (pj | qj) site(l, pj, qj); }
            }
        }
        st = warp_sum(st); sx = warp_sum(sx); sb = warp_sum(sb);
        if (lane == 0) {
            double Nt = st * dxg;
            frac_b[m] = (sb * dxg) / (Nt + 1e-12);
            mean_x[m] = sx / (st + 1e-12);
        }
    }
    __syncthreads();

    // ---- window selection, verbatim quirks of compute_v_eff_and_window (:139-154) ----
    int start_idx = (int)(0.65 * (double)M), end_idx = M;
    {
        int unsafe = 0;
        for (int m = 0; m < M; ++m) unsafe += (frac_b[m] >= a.max_boundary_fraction);
        if (unsafe > 0 && unsafe > start_idx) {   // `safe[start_idx:]` non-empty -> `~idx` is always truthy
            end_idx = start_idx;
            int min_len = (int)(a.min_window_fraction * (double)M);
            if (min_len < 3) min_len = 3;
            if (end_idx - start_idx < min_len) end_idx = (start_idx + min_len < M) ? start_idx + min_len : M;
        }
    }
    const int wlen = end_idx - start_idx;
    // ---- v_eff = np.gradient(mean_x, times) averaged over the window; mean magnetisation ----
    const double* t = a.times_obs;
    bool uniform = true;
    for (int m = 1; m + 1 < M; ++m) if ((t[m + 1] - t[m]) != (t[1] - t[0])) uniform = false;
    double acc_v = 0.0, acc_m = 0.0;
    for (int m = start_idx + tid; m < end_idx; m += NT) {
        double g;
        if (M < 2) g = 0.0;
        else if (m == 0) g = (mean_x[1] - mean_x[0]) / (t[1] - t[0]);
        else if (m == M - 1) g = (mean_x[M - 1] - mean_x[M - 2]) / (t[M - 1] - t[M - 2]);
        else if (uniform) g = (mean_x[m + 1] - mean_x[m - 1]) / (2.0 * (t[1] - t[0]));
        else {
            double hd = t[m + 1] - t[m], hs = t[m] - t[m - 1];
            double ca = -hd / (hs * (hd + hs)), cb = (hd - hs) / (hd * hs), cc = hs / (hd * (hd + hs));
            g = ca * mean_x[m - 1] + cb * mean_x[m] + cc * mean_x[m + 1];
        }
        if (a.v_eff) a.v_eff[(size_t)rep * M + m] = g;
        acc_v += g;
        acc_m += (m < nobs) ? (double)a.obs_sigma_sum[(size_t)rep * M + m] / (double)n : 0.0;
    }
    acc_v = block_sum(acc_v, scr); acc_m = block_sum(acc_m, scr);
    const double mean_v = wlen > 0 ? acc_v / (double)wlen : 0.0;
    const double m_mean = wlen > 0 ? acc_m / (double)wlen : 0.0;

    // ---- rho_eff (front density) and blocking probability: one warp per window row ----
    double rsum = 0.0, rcnt = 0.0, attempts = 0.0, blocked = 0.0;
    for (int m = start_idx + wid; m < end_idx; m += NW) {
        if (m >= nobs) continue;
        const int8_t* cp = a.obs_cp + ((size_t)rep * M + m) * L;
        const int8_t* cm = a.obs_cm + ((size_t)rep * M + m) * L;
        int jmax = -1;
        double att = 0.0, blk = 0.0;
        auto site2 = [&](int l, int pj, int qj, int pn, int qn) {      // pn, qn: counts of the right neighbour
            if ((pj + qj) > 0 && l > jmax) jmax = l;
            if (pj > 0 && l + 1 < L) {
                double rp = dens(pj);
                att += rp;
                if (APS_ADD(dens(pn), dens(qn)) >= 1.0) blk += rp;
            }
        };
        if (vec) {
            const uint32_t* cp4 = reinterpret_cast<const uint32_t*>(cp);
            const uint32_t* cm4 = reinterpret_cast<const uint32_t*>(cm);
            for (int w = lane; w < L / 4; w += 32) {
                const uint32_t pw = cp4[w], qw = cm4[w];
                if (pw | qw) {
                    const int l0 = 4 * w;
                    const int pnext = l0 + 4 < L ? cp[l0 + 4] : 0, qnext = l0 + 4 < L ? cm[l0 + 4] : 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pj = (pw >> (8 * j)) & 0xff, qj = (qw >> (8 * j)) & 0xff;
                        const int pn = j < 3 ? (int)((pw >> (8 * j + 8)) & 0xff) : pnext, qn = j < 3 ? (int)((qw >> (8 * j + 8)) & 0xff) : qnext;
                        if (pj | qj) site2(l0 + j, pj, qj, pn, qn);
                    }
                }
            }
        } else {
            for (int l = lane; l < L; l += 32) {
                const int pj = cp[l], qj = cm[l];
                if (pj | qj) site2(l, pj, qj, l + 1 < L ? cp[l + 1] : 0, l + 1 < L ? cm[l + 1] : 0);
            }
        }
        for (int o = 16; o > 0; o >>= 1) { int v = __shfl_xor_sync(0xffffffffu, jmax, o); jmax = v > jmax ? v : jmax; }
        attempts += warp_sum(att); blocked += warp_sum(blk);
        if (jmax >= 0) {
            const double xmax = xgrid(jmax, L, step), lo = xmax - a.window_fraction;
            // only the sites of the front window can pass the test below: start two grid points left of lo
            int lfirst = step > 0.0 ? (int)(lo / step) - 2 : 0;
            if (lfirst < 0) lfirst = 0;
            double s2 = 0.0;
            for (int l = lfirst + lane; l <= jmax; l += 32) {
                double x = xgrid(l, L, step);
                if (x >= lo && x <= xmax) s2 += APS_ADD(dens(cp[l]), dens(cm[l]));
            }
            s2 = warp_sum(s2);
            rsum += s2 * dxg / a.window_fraction; rcnt += 1.0;
        }
    }
    // every lane of a warp holds the same partials: count each warp once
    rsum = block_sum(lane == 0 ? rsum : 0.0, scr); rcnt = block_sum(lane == 0 ? rcnt : 0.0, scr);
    attempts = block_sum(lane == 0 ? attempts : 0.0, scr); blocked = block_sum(lane == 0 ? blocked : 0.0, scr);
    const double rho_eff = rcnt > 0.0 ? rsum / rcnt : nan("");
    const double block = attempts > 0.0 ? blocked / attempts : 0.0;

    // ---- D_eff: slope of the per-particle MSD against time (np.polyfit degree 1), one warp per row ----
    double d_eff = nan("");
    if (a.obs_pos && wlen >= 3 && n >= 2 && nobs >= end_idx) {
        const int32_t* p0 = a.obs_pos + ((size_t)rep * M + start_idx) * a.n_max;
        for (int k = start_idx + 1 + wid; k < end_idx; k += NW) {
            const int32_t* pk = a.obs_pos + ((size_t)rep * M + k) * a.n_max;
            double s1 = 0.0;
            for (int i = lane; i < n; i += 32) s1 += (double)pk[i] * a.dx - (double)p0[i] * a.dx;
            s1 = warp_sum(s1);
            const double rbar = s1 / (double)n;
            double s2 = 0.0;
            for (int i = lane; i < n; i += 32) { double ri = ((double)pk[i] * a.dx - (double)p0[i] * a.dx) - rbar; s2 += ri * ri; }
            s2 = warp_sum(s2);
            if (lane == 0) aux[k] = s2 / (double)(n - 1);
        }
        __syncthreads();
        const int cnt = wlen - 1;
        double tb = 0.0, sb2 = 0.0;
        for (int k = start_idx + 1; k < end_idx; ++k) { tb += t[k] - t[start_idx]; sb2 += aux[k]; }
        tb /= cnt; sb2 /= cnt;
        double num = 0.0, den = 0.0;
        for (int k = start_idx + 1; k < end_idx; ++k) {
            double dt = (t[k] - t[start_idx]) - tb;
            num += dt * (aux[k] - sb2); den += dt * dt;
        }
        d_eff = num / den;
    }
    if (tid == 0) {
        out[APS_RED_V_EFF] = mean_v; out[APS_RED_D_EFF] = d_eff; out[APS_RED_M_MEAN] = m_mean;
        out[APS_RED_RHO_EFF] = rho_eff; out[APS_RED_BLOCK] = block;
        out[APS_RED_START] = (double)start_idx; out[APS_RED_END] = (double)end_idx; out[APS_RED_NOBS] = (double)nobs;
    }
}

// ---- reduce_kernel (round 2): the same reducers on INTEGER row sums ---------------------------------------------------
// The observation rows hold small non-negative counts, so every lattice sum of the drivers' reducers is a sum of integers
// times one constant: sum_l rho(l) = C / denom, sum_l rho(l) x_l = step * (sum_l c_l l) / denom, the attempts / blocked
// sums of compute_blocking_probability are counts over 1 / denom, the MSD of compute_D_eff_active is
// dx^2 (n S2 - S1^2) / (n (n - 1)) with S1 = sum d_i, S2 = sum d_i^2 of the integer displacements.  The sums are formed
// exactly (dp4a: four sites per instruction; 8-byte row loads) and converted once per row, instead of a table lookup, a
// grid-point product and three double additions per site (v1: 2.5 ms of the 64 ms bench step, ~10x its HBM time).
// Against v1 / numpy only the association of the roundings differs (<= 1e-15 relative; tests: 1e-9 against the reference's
// own functions, 1e-10 against v1 on random rows with counts up to 3).
template <class F>
__device__ __forceinline__ void row_words(const int8_t* cp, const int8_t* cm, int L, int lane, F&& f) {
    if ((L & 7) == 0) {                       // rows start 8-byte aligned
        const uint2* c2 = reinterpret_cast<const uint2*>(cp);
        const uint2* m2 = reinterpret_cast<const uint2*>(cm);
        for (int w = lane; w < (L >> 3); w += 32) {
            const uint2 x = c2[w], y = m2[w];
            f(8 * w, x.x, y.x, x.y & 0xffu, y.y & 0xffu, true);
            f(8 * w + 4, x.y, y.y, 0u, 0u, false);
        }
    } else if ((L & 3) == 0) {
        const uint32_t* c4 = reinterpret_cast<const uint32_t*>(cp);
        const uint32_t* m4 = reinterpret_cast<const uint32_t*>(cm);
        for (int w = lane; w < (L >> 2); w += 32) f(4 * w, c4[w], m4[w], 0u, 0u, false);
    } else {
        for (int w = lane; w < ((L + 3) >> 2); w += 32) {
            uint32_t pw = 0, qw = 0;
            for (int j = 0; j < 4; ++j) if (4 * w + j < L) { pw |= (uint32_t)(uint8_t)cp[4 * w + j] << (8 * j); qw |= (uint32_t)(uint8_t)cm[4 * w + j] << (8 * j); }
            f(4 * w, pw, qw, 0u, 0u, false);
        }
    }
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void reduce_kernel(aps_reduce_args a) {
    extern __shared__ double sh[];
    const int rep = blockIdx.x, tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const int M = a.M, L = a.L, n = a.n[rep], nobs = a.n_obs[rep];
    double* mean_x = sh;            // [M]
    double* frac_b = sh + M;        // [M]
    double* aux = sh + 2 * M;       // [M] scratch (S_k of the MSD fit)
    double* scr = sh + 3 * M;       // [32] reduction scratch
    double* out = a.out + (size_t)rep * APS_RED_N;
    const double step = L > 1 ? APS_DIV(1.0, (double)(L - 1)) : 0.0;
    const double dxg = L > 1 ? APS_SUB(xgrid(1, L, step), xgrid(0, L, step)) : 0.0;
    const double denom = APS_MUL((double)(n > 1 ? n : 1), a.dx);

    double* dtab = scr + 32;        // [128] density of a site holding c particles of one species: c / (max(1,n)*dx) (CLASS.py:208-213)
    for (int c = tid; c < 128; c += NT) dtab[c] = APS_DIV((double)c, denom);
    __syncthreads();
    auto dens = [&](int c) { return dtab[c & 127]; };
    // first grid point with x_l >= boundary_xmin (x_l is non-decreasing in l): v1's per-site test, located once
    int lmin = L;
    if (a.boundary_xmin == a.boundary_xmin) {
        double g0 = step > 0.0 ? a.boundary_xmin / step - 2.0 : 0.0;
        int g = g0 < 0.0 ? 0 : (g0 > (double)L ? L : (int)g0);
        while (g < L && !(xgrid(g, L, step) >= a.boundary_xmin)) ++g;
        lmin = g;
    }
    // ---- per-row sums over the lattice (rows never reached are all-zero in the reference) ----
    for (int m = wid; m < M; m += NW) {
        uint32_t C = 0, CB = 0, CLlo = 0;
        unsigned long long CW = 0;
        if (m < nobs) {
            const int8_t* cp = a.obs_cp + ((size_t)rep * M + m) * L;
            const int8_t* cm = a.obs_cm + ((size_t)rep * M + m) * L;
            row_words(cp, cm, L, lane, [&](int l0, uint32_t pw, uint32_t qw, uint32_t, uint32_t, bool) {
                const uint32_t tw = pw + qw;                               // counts per site (< 256: no carry between the bytes)
                const uint32_t cw = __dp4a(tw, 0x01010101u, 0u);
                C += cw;
                CLlo = __dp4a(tw, 0x03020100u, CLlo);
                CW += (unsigned long long)((uint32_t)l0 * cw);             // l0 * cw < 2^32 for L < 4e6
                if (l0 >= lmin) CB += cw;
                else if (l0 + 3 >= lmin) CB = __dp4a(tw, 0x01010101u << (8 * (lmin - l0)), CB);
            });
        }
        C = __reduce_add_sync(0xffffffffu, C); CB = __reduce_add_sync(0xffffffffu, CB); CLlo = __reduce_add_sync(0xffffffffu, CLlo);
        CW = warp_sum_u64(CW);
        if (lane == 0) {
            const double st = APS_DIV((double)C, denom), sb = APS_DIV((double)CB, denom);
            const double sx = APS_DIV(APS_MUL((double)(CW + CLlo), step), denom);
            const double Nt = st * dxg;
            frac_b[m] = (sb * dxg) / (Nt + 1e-12);
            mean_x[m] = sx / (st + 1e-12);
        }
    }
    __syncthreads();

    // ---- window selection, verbatim quirks of compute_v_eff_and_window (:139-154) ----
    int start_idx = (int)(0.65 * (double)M), end_idx = M;
    {
        int unsafe = 0;
        for (int m = 0; m < M; ++m) unsafe += (frac_b[m] >= a.max_boundary_fraction);
        if (unsafe > 0 && unsafe > start_idx) {   // `safe[start_idx:]` non-empty -> `~idx` is always truthy
            end_idx = start_idx;
            int min_len = (int)(a.min_window_fraction * (double)M);
            if (min_len < 3) min_len = 3;
            if (end_idx - start_idx < min_len) end_idx = (start_idx + min_len < M) ? start_idx + min_len : M;
        }
    }
    const int wlen = end_idx - start_idx;
    // ---- v_eff = np.gradient(mean_x, times) averaged over the window; mean magnetisation ----
    const double* t = a.times_obs;
    bool uniform = true;
    for (int m = 1; m + 1 < M; ++m) if ((t[m + 1] - t[m]) != (t[1] - t[0])) uniform = false;
    double acc_v = 0.0, acc_m = 0.0;
    for (int m = start_idx + tid; m < end_idx; m += NT) {
        double g;
        if (M < 2) g = 0.0;
        else if (m == 0) g = (mean_x[1] - mean_x[0]) / (t[1] - t[0]);
        else if (m == M - 1) g = (mean_x[M - 1] - mean_x[M - 2]) / (t[M - 1] - t[M - 2]);
        else if (uniform) g = (mean_x[m + 1] - mean_x[m - 1]) / (2.0 * (t[1] - t[0]));
        else {
            double hd = t[m + 1] - t[m], hs = t[m] - t[m - 1];
            double ca = -hd / (hs * (hd + hs)), cb = (hd - hs) / (hd * hs), cc = hs / (hd * (hd + hs));
            g = ca * mean_x[m - 1] + cb * mean_x[m] + cc * mean_x[m + 1];
        }
        if (a.v_eff) a.v_eff[(size_t)rep * M + m] = g;
        acc_v += g;
        acc_m += (m < nobs) ? (double)a.obs_sigma_sum[(size_t)rep * M + m] / (double)n : 0.0;
    }
    acc_v = block_sum(acc_v, scr); acc_m = block_sum(acc_m, scr);
    const double mean_v = wlen > 0 ? acc_v / (double)wlen : 0.0;
    const double m_mean = wlen > 0 ? acc_m / (double)wlen : 0.0;

    // ---- rho_eff (front density) and blocking probability: one warp per window row ----
    // attempts = sum of rho_plus over the sites with a right neighbour, blocked = those whose neighbour holds total density >= 1
    // (sweep_beta.py:197-229): both are (integer count) / denom, so the ratio is the ratio of the counts.
    const bool one_blocks = APS_ADD(dens(1), dens(0)) >= 1.0;
    double rsum = 0.0, rcnt = 0.0;
    unsigned long long att_tot = 0, blk_tot = 0;
    for (int m = start_idx + wid; m < end_idx; m += NW) {
        if (m >= nobs) continue;
        const int8_t* cp = a.obs_cp + ((size_t)rep * M + m) * L;
        const int8_t* cm = a.obs_cm + ((size_t)rep * M + m) * L;
        int jmax = -1;
        uint32_t att = 0, blk = 0;
        row_words(cp, cm, L, lane, [&](int l0, uint32_t pw, uint32_t qw, uint32_t pnx, uint32_t qnx, bool have_next) {
            const uint32_t tw = pw + qw;
            if (tw == 0) return;
            jmax = l0 + ((31 - __clz(tw)) >> 3);                          // words arrive in increasing order per lane
            if (pw == 0) return;
            if (!have_next && (pw >> 24) != 0 && l0 + 4 < L) { pnx = (uint32_t)(uint8_t)cp[l0 + 4]; qnx = (uint32_t)(uint8_t)cm[l0 + 4]; }
            uint32_t pv = pw;                                             // the last site has no right neighbour: not an attempt
            if (L - 1 - l0 < 4) pv &= ~(0xFFu << (8 * (L - 1 - l0)));
            const uint32_t pn = (pw >> 8) | (pnx << 24), qn = (qw >> 8) | (qnx << 24);   // counts of the right neighbours
            if ((((pw | qw | pnx | qnx) & 0xFEFEFEFEu) | (pn & qn)) == 0u) {           // counts 0 / 1, no neighbour with both species
                att += __popc(pv);
                if (one_blocks) blk += __popc(pv & (pn | qn));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t pj = (pv >> (8 * j)) & 0xffu;
                    if (pj) {
                        att += pj;
                        if (APS_ADD(dens((int)((pn >> (8 * j)) & 0xffu)), dens((int)((qn >> (8 * j)) & 0xffu))) >= 1.0) blk += pj;
                    }
                }
            }
        });
        jmax = __reduce_max_sync(0xffffffffu, jmax);
        att_tot += __reduce_add_sync(0xffffffffu, att); blk_tot += __reduce_add_sync(0xffffffffu, blk);
        if (jmax >= 0) {
            const double xmax = xgrid(jmax, L, step), lo = xmax - a.window_fraction;
            // only the sites of the front window can pass the test below: start two grid points left of lo
            int lfirst = step > 0.0 ? (int)(lo / step) - 2 : 0;
            if (lfirst < 0) lfirst = 0;
            double s2 = 0.0;
            for (int l = lfirst + lane; l <= jmax; l += 32) {
                double x = xgrid(l, L, step);
                if (x >= lo && x <= xmax) s2 += APS_ADD(dens(cp[l]), dens(cm[l]));
            }
            s2 = warp_sum(s2);
            rsum += s2 * dxg / a.window_fraction; rcnt += 1.0;
        }
    }
    // every lane of a warp holds the same partials: count each warp once (the counts are < 2^53: exact in double)
    rsum = block_sum(lane == 0 ? rsum : 0.0, scr); rcnt = block_sum(lane == 0 ? rcnt : 0.0, scr);
    const double attempts = block_sum(lane == 0 ? (double)att_tot : 0.0, scr), blocked = block_sum(lane == 0 ? (double)blk_tot : 0.0, scr);
    const double rho_eff = rcnt > 0.0 ? rsum / rcnt : nan("");
    const double block = attempts > 0.0 ? blocked / attempts : 0.0;

    // ---- D_eff: slope of the per-particle MSD against time (np.polyfit degree 1), one warp per row ----
    double d_eff = nan("");
    if (a.obs_pos && wlen >= 3 && n >= 2 && nobs >= end_idx) {
        const int32_t* p0 = a.obs_pos + ((size_t)rep * M + start_idx) * a.n_max;
        const double dx2 = a.dx * a.dx;
        for (int k = start_idx + 1 + wid; k < end_idx; k += NW) {
            const int32_t* pk = a.obs_pos + ((size_t)rep * M + k) * a.n_max;
            long long s1 = 0;
            unsigned long long s2 = 0;
            for (int i = lane; i < n; i += 32) { const long long d = (long long)pk[i] - (long long)p0[i]; s1 += d; s2 += (unsigned long long)(d * d); }
            s1 = (long long)warp_sum_u64((unsigned long long)s1); s2 = warp_sum_u64(s2);
            // sample variance of the displacements r_i = d_i dx:  dx^2 (n S2 - S1^2) / (n (n - 1))
            if (lane == 0) aux[k] = ((double)((long long)n * (long long)s2 - s1 * s1) * dx2) / ((double)n * (double)(n - 1));
        }
        __syncthreads();
        const int cnt = wlen - 1;
        double tb = 0.0, sb2 = 0.0;
        for (int k = start_idx + 1; k < end_idx; ++k) { tb += t[k] - t[start_idx]; sb2 += aux[k]; }
        tb /= cnt; sb2 /= cnt;
        double num = 0.0, den = 0.0;
        for (int k = start_idx + 1; k < end_idx; ++k) {
            double dt = (t[k] - t[start_idx]) - tb;
            num += dt * (aux[k] - sb2); den += dt * dt;
        }
        d_eff = num / den;
    }
    if (tid == 0) {
        out[APS_RED_V_EFF] = mean_v; out[APS_RED_D_EFF] = d_eff; out[APS_RED_M_MEAN] = m_mean;
        out[APS_RED_RHO_EFF] = rho_eff; out[APS_RED_BLOCK] = block;
        out[APS_RED_START] = (double)start_idx; out[APS_RED_END] = (double)end_idx; out[APS_RED_NOBS] = (double)nobs;
    }
}

// Ensemble profile sums.  Replicas are laid out grid-point-major: rep = g*reps_per_point + j.
// For every grid point g and site l:   prof[g][q][l]   = sum_j  mean_{m in [row_lo,row_hi)} q_j(m,l)
//                                      prof2[g][q][l]  = sum_j (mean_m q_j(m,l))^2
// with q in {rho_plus, rho_minus}; dividing by reps_per_point (after the cross-GPU allreduce) gives the
// ensemble mean profile and its standard error.
__global__ void profile_kernel(aps_profile_args a) {
    const int g = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= a.L) return;
    const int L = a.L, M = a.M;
    double sp = 0.0, sm = 0.0, sp2 = 0.0, sm2 = 0.0;
    const int rows = a.row_hi - a.row_lo;
    const int j_lo = a.point_start ? a.point_start[g] : 0, j_hi = a.point_start ? a.point_start[g + 1] : a.reps_per_point;
    for (int j = j_lo; j < j_hi; ++j) {
        const int rep = a.point_start ? a.point_reps[j] : g * a.reps_per_point + j;
        const int n = a.n[rep];
        const double denom = (double)(n > 1 ? n : 1) * a.dx;
        int cp = 0, cm = 0;
        for (int m = a.row_lo; m < a.row_hi; ++m) {
            if (m >= a.n_obs[rep]) break;
            cp += a.obs_cp[((size_t)rep * M + m) * L + l];
            cm += a.obs_cm[((size_t)rep * M + m) * L + l];
        }
        const double mp = (double)cp / denom / (double)rows, mm = (double)cm / denom / (double)rows;
        sp += mp; sm += mm; sp2 += mp * mp; sm2 += mm * mm;
    }
    double* o = a.prof + ((size_t)g * 4) * L;
    o[l] = sp; o[L + l] = sm; o[2 * L + l] = sp2; o[3 * L + l] = sm2;
}

// Per-point sums from per-replica rows: out[g][q][l] = sum over the replicas of point g (CSR list, ascending: a fixed
// summation order, so the result does not depend on the launch geometry) of per_rep[rep][q][l].  q = 0,1: time-averaged
// rho_plus / rho_minus of the replica (profile_kernel with reps_per_point = 1), q = 2,3: their squares are formed here.
// profile_kernel for L % 4 == 0: four sites per thread (one 32-bit load per row and species), the counts of the even / odd
// bytes accumulated in 16-bit lanes and flushed every 256 rows (counts <= 127).  Same integers, hence the same doubles, as profile_kernel.
__global__ void profile_kernel_w4(aps_profile_args a) {
    const int g = blockIdx.y;
    const int l4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l4 >= a.L) return;
    const int L = a.L, M = a.M;
    double sp[4] = {0.0, 0.0, 0.0, 0.0}, sm[4] = {0.0, 0.0, 0.0, 0.0}, sp2[4] = {0.0, 0.0, 0.0, 0.0}, sm2[4] = {0.0, 0.0, 0.0, 0.0};
    const int rows = a.row_hi - a.row_lo;
    const int j_lo = a.point_start ? a.point_start[g] : 0, j_hi = a.point_start ? a.point_start[g + 1] : a.reps_per_point;
    for (int j = j_lo; j < j_hi; ++j) {
        const int rep = a.point_start ? a.point_reps[j] : g * a.reps_per_point + j;
        const int n = a.n[rep], nobs = a.n_obs[rep];
        const double denom = (double)(n > 1 ? n : 1) * a.dx;
        const int m_hi = a.row_hi < nobs ? a.row_hi : nobs;
        int cp[4] = {0, 0, 0, 0}, cm[4] = {0, 0, 0, 0};
        uint32_t pe = 0, po = 0, qe = 0, qo = 0;
        int pending = 0;
        auto flush = [&]() {
            cp[0] += pe & 0xffffu; cp[2] += pe >> 16; cp[1] += po & 0xffffu; cp[3] += po >> 16;
            cm[0] += qe & 0xffffu; cm[2] += qe >> 16; cm[1] += qo & 0xffffu; cm[3] += qo >> 16;
            pe = po = qe = qo = 0u; pending = 0;
        };
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(a.obs_cp + ((size_t)rep * M + a.row_lo) * L + l4);
        const uint32_t* rq = reinterpret_cast<const uint32_t*>(a.obs_cm + ((size_t)rep * M + a.row_lo) * L + l4);
        const size_t stride = (size_t)L / 4;
#pragma unroll 4
        for (int m = a.row_lo; m < m_hi; ++m, rp += stride, rq += stride) {
            const uint32_t pw = *rp, qw = *rq;
            pe += pw & 0x00ff00ffu; po += (pw >> 8) & 0x00ff00ffu;
            qe += qw & 0x00ff00ffu; qo += (qw >> 8) & 0x00ff00ffu;
            if (++pending == 256) flush();
        }
        flush();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double mp = (double)cp[k] / denom / (double)rows, mm = (double)cm[k] / denom / (double)rows;
            sp[k] += mp; sm[k] += mm; sp2[k] += mp * mp; sm2[k] += mm * mm;
        }
    }
    double* o = a.prof + ((size_t)g * 4) * L + l4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { o[k] = sp[k]; o[L + k] = sm[k]; o[2 * L + k] = sp2[k]; o[3 * L + k] = sm2[k]; }
}
__global__ void profile_gather_kernel(const double* __restrict__ per_rep, const int32_t* __restrict__ point_start,
                                      const int32_t* __restrict__ point_reps, double* __restrict__ out, int L) {
    const int g = blockIdx.y;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    double sp = 0.0, sm = 0.0, sp2 = 0.0, sm2 = 0.0;
    for (int j = point_start[g]; j < point_start[g + 1]; ++j) {
        const double* row = per_rep + (size_t)point_reps[j] * 4 * L;
        const double mp = row[l], mm = row[L + l];
        sp += mp; sm += mm; sp2 += mp * mp; sm2 += mm * mm;
    }
    double* o = out + (size_t)g * 4 * L;
    o[l] = sp; o[L + l] = sm; o[2 * L + l] = sp2; o[3 * L + l] = sm2;
}

// One thread per replica: time-averaged m_global over a row window -> per-grid-point histogram (integer atomics).
__global__ void hist_kernel(aps_hist_args a) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= a.n_replicas) return;
    const int hi_row = a.row_hi < a.n_obs[rep] ? a.row_hi : a.n_obs[rep];
    double acc = 0.0;
    int rows = 0;
    for (int m = a.row_lo; m < hi_row; ++m, ++rows) {
        const int n = a.obs_n ? a.obs_n[(size_t)rep * a.M + m] : a.n[rep];
        acc += (double)a.obs_sigma_sum[(size_t)rep * a.M + m] / (double)(n > 1 ? n : 1);
    }
    if (rows == 0) { if (a.mbar) a.mbar[rep] = nan(""); return; }
    const double mb = acc / (double)rows;
    if (a.mbar) a.mbar[rep] = mb;
    int bin = (int)floor((mb - a.lo) / (a.hi - a.lo) * (double)a.n_bins);
    bin = bin < 0 ? 0 : (bin >= a.n_bins ? a.n_bins - 1 : bin);
    const int g = a.point_of ? a.point_of[rep] : 0;
    atomicAdd(a.hist + (size_t)g * a.n_bins + bin, 1ull);
}

}  // namespace aps
