// aps_pde.cu — batched IMEX hydrodynamic-PDE stepper (include/aps_pde.h; IMEX_PDE_solver_class.py:157-289).
//
// One CTA of 1024 threads owns one solver instance and runs ALL its time steps inside one launch; rho_plus, rho_minus,
// the diffused fields, the local magnetisation and the ring kernel live in shared memory for the whole run, so HBM
// sees only the per-step diagnostics (two doubles), the snapshot rows and the tracer ring.  Per step:
//   * local magnetisation: pointwise, or ring-kernel convolution as a direct sum over +-radius taps (:157-169);
//   * implicit diffusion (I - gamma*dt/dx^2 * Lap) x = rho (:189-190): the cyclic tridiagonal system with constant
//     coefficients factorises into a causal and an anticausal first-order recursive filter,
//         (1+2a) - a(S + S^-1) = (a/r)(1 - r S)(1 - r S^-1),   r + 1/r = 2 + 1/a,  0 < r < 1,
//     each evaluated as a block-wide scan of affine maps with the periodic closure y_-1 = B_tot / (1 - A_tot);
//     Neumann boundaries (rows "2, -2" of the reference matrix) are the same solve on the even extension (2L-2 points);
//   * upwind advection, Curie-Weiss reaction with clipped rates, clipping at zero, mass renormalisation (:192-233);
//   * tracers (:255-282): flip with probability rate*dt, drift lam*state*dt plus sqrt(2 gamma dt) N(0,1) from Philox,
//     displacement statistics over a window of past positions kept in a global-memory ring.
// This translation unit is compiled WITHOUT --fmad=false: parity with the reference is by tolerance here (its solve is
// SuperLU, its convolution an FFT), not bitwise.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/aps_pde.h"
#include "../../include/aps_philox.h"

namespace aps {

constexpr int kPdeThreads = 1024;
constexpr int kPdeMaxL = 2048;
constexpr int kPdeMaxTracers = 2048;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kRngTracer = 0x7DE0u;

struct Affine { double A, B; };                                   // y -> A*y + B
__device__ __forceinline__ Affine then(Affine f, Affine g) { return Affine{f.A * g.A, g.A * f.B + g.B}; }   // g after f

struct PdeSmem {
    double *p, *m, *dp, *dm, *mf, *kern, *ext, *tpos;
    int8_t* tstate;
    Affine* wbuf;      // [32]
    double* red;       // [32]
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = red[lane];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    __syncthreads();
    return t;
}

// exclusive scan of affine maps in thread order; *tot = composition of all maps
__device__ __forceinline__ Affine block_scan_affine(Affine v, Affine* wbuf, Affine* tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Affine inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        Affine up{__shfl_up_sync(kFull, inc.A, o), __shfl_up_sync(kFull, inc.B, o)};
        if (lane >= o) inc = then(up, inc);
    }
    if (lane == 31) wbuf[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        Affine w = wbuf[lane];
        for (int o = 1; o < 32; o <<= 1) {
            Affine up{__shfl_up_sync(kFull, w.A, o), __shfl_up_sync(kFull, w.B, o)};
            if (lane >= o) w = then(up, w);
        }
        wbuf[lane] = w;
    }
    __syncthreads();
    Affine prev{__shfl_up_sync(kFull, inc.A, 1), __shfl_up_sync(kFull, inc.B, 1)};
    if (lane == 0) prev = Affine{1.0, 0.0};
    const Affine base = wid ? wbuf[wid - 1] : Affine{1.0, 0.0};
    *tot = wbuf[31];
    __syncthreads();
    return then(base, prev);
}

// cyclic first-order filter in place: y_k = e_k + r * y_{k-1} around the ring of n points, k running forwards
// (reverse = false) or backwards (reverse = true) through the array
__device__ __forceinline__ void cyclic_filter(double* e, int n, double r, bool reverse, Affine* wbuf) {
    const int C = (n + kPdeThreads - 1) / kPdeThreads;
    const int k0 = threadIdx.x * C, k1 = (k0 + C < n) ? k0 + C : n;
    double run = 0.0, A = 1.0;
    for (int k = k0; k < k1; ++k) { run = fma(r, run, e[reverse ? n - 1 - k : k]); A *= r; }
    Affine tot;
    const Affine ex = block_scan_affine(Affine{A, run}, wbuf, &tot);
    const double y_init = tot.B / (1.0 - tot.A);                  // periodic closure
    double y = fma(ex.A, y_init, ex.B);
    for (int k = k0; k < k1; ++k) { const int i = reverse ? n - 1 - k : k; y = fma(r, y, e[i]); e[i] = y; }
    __syncthreads();
}

// dst = (I - a * Lap)^-1 src with the reference's periodic or Neumann matrix (:71-85)
__device__ __forceinline__ void diffuse(const double* src, double* dst, int L, int bc, double a, const PdeSmem& S) {
    if (!(a > 0.0)) {
        for (int i = threadIdx.x; i < L; i += kPdeThreads) dst[i] = src[i];
        __syncthreads();
        return;
    }
    const double q = 2.0 + 1.0 / a;
    const double r = 2.0 / (q + sqrt(q * q - 4.0));               // the root below 1 of r + 1/r = q, without cancellation
    const double scale = r / a;
    if (bc == APS_PDE_BC_PERIODIC) {
        for (int i = threadIdx.x; i < L; i += kPdeThreads) dst[i] = src[i];
        __syncthreads();
        cyclic_filter(dst, L, r, false, S.wbuf);
        cyclic_filter(dst, L, r, true, S.wbuf);
        for (int i = threadIdx.x; i < L; i += kPdeThreads) dst[i] *= scale;
        __syncthreads();
    } else {                                                      // even extension: x_-1 = x_1, x_L = x_{L-2}
        const int n = 2 * L - 2;
        for (int i = threadIdx.x; i < n; i += kPdeThreads) S.ext[i] = src[i < L ? i : n - i];
        __syncthreads();
        cyclic_filter(S.ext, n, r, false, S.wbuf);
        cyclic_filter(S.ext, n, r, true, S.wbuf);
        for (int i = threadIdx.x; i < L; i += kPdeThreads) dst[i] = S.ext[i] * scale;
        __syncthreads();
    }
}

__device__ __forceinline__ double cw_rate(double beta, double sigma, double m) {   // :64-66
    const double r = exp(-beta * sigma * m);
    return fmin(fmax(r, 1e-8), 1e8);
}

__global__ void __launch_bounds__(kPdeThreads, 1) pde_kernel(const __grid_constant__ aps_pde_args a) {
    extern __shared__ __align__(16) unsigned char pde_raw[];
    const int run = blockIdx.x, tid = threadIdx.x;
    const int L = a.L;
    PdeSmem S;
    {
        double* d = reinterpret_cast<double*>(pde_raw);
        S.p = d; d += L; S.m = d; d += L; S.dp = d; d += L; S.dm = d; d += L; S.mf = d; d += L; S.kern = d; d += L;
        S.ext = d; d += (a.bc == APS_PDE_BC_NEUMANN ? 2 * L : 0);
        S.tpos = d; d += a.n_tracers;
        S.red = d; d += 32;
        S.wbuf = reinterpret_cast<Affine*>(d); d += 64;
        S.tstate = reinterpret_cast<int8_t*>(d);
    }
    const double beta = a.beta[run], lam = a.lam[run], gamma = a.gamma[run];
    const double dt = a.dt, dx = a.dx, xlim = a.xlim;
    const double adiff = gamma * dt / (dx * dx);
    const int rad = a.field == APS_PDE_FIELD_KERNEL ? a.radius[run] : 0;
    const bool full_ring = rad >= L / 2;
    const int npair = full_ring ? (L - 1) / 2 : rad;              // symmetric tap pairs d = 1..npair
    const bool antipode = full_ring && (L % 2 == 0);              // single tap at ring distance L/2
    const int64_t nsteps = a.nsteps;
    const int n_snap_rows = (int)(nsteps / a.snapshot_interval) + 1;
    const uint32_t k0 = (uint32_t)(a.seeds ? a.seeds[run] : 0), k1 = (uint32_t)((a.seeds ? a.seeds[run] : 0) >> 32);
    const int ntr = a.n_tracers;

    for (int i = tid; i < L; i += kPdeThreads) {
        S.p[i] = a.rho_p[(size_t)run * L + i];
        S.m[i] = a.rho_m[(size_t)run * L + i];
        if (a.field == APS_PDE_FIELD_KERNEL) S.kern[i] = a.kernel[(size_t)run * L + i];
    }
    for (int j = tid; j < ntr; j += kPdeThreads) {
        S.tpos[j] = a.tracer_pos[(size_t)run * ntr + j];
        S.tstate[j] = a.tracer_state[(size_t)run * ntr + j];
    }
    __syncthreads();

    for (int64_t n = 0; n <= nsteps; ++n) {
        // ---- local magnetisation of the current state (:157-169) ----
        for (int i = tid; i < L; i += kPdeThreads) {
            double num, den;
            if (a.field == APS_PDE_FIELD_POINTWISE) { num = S.p[i] - S.m[i]; den = S.p[i] + S.m[i]; }
            else {
                const double k0w = S.kern[0];
                num = k0w * (S.p[i] - S.m[i]); den = k0w * (S.p[i] + S.m[i]);
                int jl = i, jr = i;
                for (int d = 1; d <= npair; ++d) {
                    jl = jl == 0 ? L - 1 : jl - 1; jr = jr == L - 1 ? 0 : jr + 1;
                    const double kw = S.kern[d];
                    const double pl = S.p[jl], ml = S.m[jl], pr = S.p[jr], mr = S.m[jr];
                    num = fma(kw, (pl - ml) + (pr - mr), num);
                    den = fma(kw, (pl + ml) + (pr + mr), den);
                }
                if (antipode) {
                    const int j = i + L / 2 >= L ? i - L / 2 : i + L / 2;
                    num = fma(S.kern[L / 2], S.p[j] - S.m[j], num);
                    den = fma(S.kern[L / 2], S.p[j] + S.m[j], den);
                }
            }
            S.mf[i] = num / (den + 1e-12);
        }
        __syncthreads();
        // ---- diagnostics (:244-245) ----
        {
            double sm = 0.0, st = 0.0;
            for (int i = tid; i < L; i += kPdeThreads) { sm += S.mf[i]; st += S.p[i] + S.m[i]; }
            const double m_mean = block_sum(sm, S.red) / L;
            const double t_mean = block_sum(st, S.red) / L;
            double sv = 0.0;
            for (int i = tid; i < L; i += kPdeThreads) { const double dlt = S.p[i] + S.m[i] - t_mean; sv = fma(dlt, dlt, sv); }
            const double var = block_sum(sv, S.red) / L;
            if (tid == 0) {
                a.m_series[(size_t)run * (nsteps + 1) + n] = m_mean;
                a.var_series[(size_t)run * (nsteps + 1) + n] = var;
            }
        }
        if (a.tot_series) {                                       // rows of the per-step spectra (:247-249)
            double* row = a.tot_series + ((size_t)run * (size_t)(nsteps + 1) + (size_t)n) * L;
            for (int i = tid; i < L; i += kPdeThreads) row[i] = S.p[i] + S.m[i];
        }
        if (n % a.snapshot_interval == 0) {                       // :251-254
            const size_t row = ((size_t)run * n_snap_rows + (size_t)(n / a.snapshot_interval)) * L;
            for (int i = tid; i < L; i += kPdeThreads) {
                if (a.snapshots) a.snapshots[row + i] = S.p[i] + S.m[i];
                if (a.m_snapshots) a.m_snapshots[row + i] = S.p[i] - S.m[i];
            }
        }
        // ---- tracers (:255-282) ----
        if (ntr > 0) {
            double s_dr = 0.0;
            const bool have = n >= a.window;
            const double amp = sqrt(2.0 * gamma * dt);
            for (int j = tid; j < ntr; j += kPdeThreads) {
                double unw = S.tpos[j];
                double wrapped = unw - floor(unw / xlim) * xlim;
                int idx = (int)(wrapped / dx) % L;
                if (idx < 0) idx += L;
                int st = S.tstate[j];
                const aps_u32x4 w = aps_philox4x32_10((uint32_t)n, (uint32_t)((uint64_t)n >> 32), (uint32_t)j, kRngTracer, k0, k1);
                const double u = aps_u53(w.v[0], w.v[1]);
                if (u < cw_rate(beta, (double)st, S.mf[idx]) * dt) st = -st;
                const double u1 = ((double)w.v[2] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)w.v[3] + 0.5) * (1.0 / 4294967296.0);
                const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
                unw += lam * (double)st * dt + amp * z;
                S.tpos[j] = unw; S.tstate[j] = (int8_t)st;
                double* ring = a.tracer_hist + (size_t)run * a.window * ntr;
                ring[(size_t)(n % a.window) * ntr + j] = unw;
                if (have) s_dr += unw - ring[(size_t)((n + 1) % a.window) * ntr + j];
            }
            const double mean_dr = block_sum(s_dr, S.red) / ntr;
            double s_v = 0.0;
            if (have) {
                const double* ring = a.tracer_hist + (size_t)run * a.window * ntr;
                for (int j = tid; j < ntr; j += kPdeThreads) {
                    const double dr = S.tpos[j] - ring[(size_t)((n + 1) % a.window) * ntr + j] - mean_dr;
                    s_v = fma(dr, dr, s_v);
                }
            }
            const double var_dr = block_sum(s_v, S.red) / ntr;
            if (tid == 0) {
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                a.v_eff_series[(size_t)run * (nsteps + 1) + n] = have ? mean_dr / (a.window * dt) : nanv;
                a.D_eff_series[(size_t)run * (nsteps + 1) + n] = have ? var_dr / (2.0 * a.window * dt) : nanv;
            }
        }
        if (n == nsteps) break;

        // ---- step() (:187-233) ----
        diffuse(S.p, S.dp, L, a.bc, adiff, S);
        diffuse(S.m, S.dm, L, a.bc, adiff, S);
        double s0 = 0.0, s1 = 0.0;
        const bool neu = a.bc == APS_PDE_BC_NEUMANN;
        if (a.model == APS_PDE_MODEL_BIDIRECTIONAL) {
            for (int i = tid; i < L; i += kPdeThreads) {
                const double dpi = S.dp[i], dmi = S.dm[i], mi = S.mf[i];
                const double der_p = (neu && i == 0) ? 0.0 : (dpi - S.dp[i == 0 ? L - 1 : i - 1]) / dx;          // right movers, upwind
                const double der_m = (neu && i == L - 1) ? 0.0 : (S.dm[i == L - 1 ? 0 : i + 1] - dmi) / dx;       // left movers
                const double Rp = cw_rate(beta, -1.0, mi) * dmi - cw_rate(beta, 1.0, mi) * dpi;
                const double np_ = fmax(dpi + dt * (-lam * der_p + Rp), 0.0);
                const double nm_ = fmax(dmi + dt * (lam * der_m - Rp), 0.0);
                S.p[i] = np_; S.m[i] = nm_;
                s0 += dpi + dmi; s1 += np_ + nm_;
            }
        } else {                                                  // anchored_minus: reaction first, then advection of '+' only
            for (int i = tid; i < L; i += kPdeThreads) {
                const double dpi = S.dp[i], dmi = S.dm[i], mi = S.mf[i];
                const double Rp = cw_rate(beta, -1.0, mi) * dmi - cw_rate(beta, 1.0, mi) * dpi;
                S.p[i] = fmax(dpi + dt * Rp, 0.0);                // rho_p_star
                S.m[i] = fmax(dmi - dt * Rp, 0.0);                // rho_m_star (final)
                s0 += dpi + dmi;
            }
            __syncthreads();
            for (int i = tid; i < L; i += kPdeThreads) {
                const double ps = S.p[i];
                const double der_p = (neu && i == 0) ? 0.0 : (ps - S.p[i == 0 ? L - 1 : i - 1]) / dx;
                S.dp[i] = fmax(ps + dt * (-lam * der_p), 0.0);
            }
            __syncthreads();
            for (int i = tid; i < L; i += kPdeThreads) { S.p[i] = S.dp[i]; s1 += S.p[i] + S.m[i]; }
        }
        const double M0 = block_sum(s0, S.red), M1 = block_sum(s1, S.red);
        const double sc = M0 / M1;
        for (int i = tid; i < L; i += kPdeThreads) { S.p[i] *= sc; S.m[i] *= sc; }
        __syncthreads();
    }

    for (int i = tid; i < L; i += kPdeThreads) {
        a.rho_p[(size_t)run * L + i] = S.p[i];
        a.rho_m[(size_t)run * L + i] = S.m[i];
    }
    for (int j = tid; j < ntr; j += kPdeThreads) {
        a.tracer_pos[(size_t)run * ntr + j] = S.tpos[j];
        a.tracer_state[(size_t)run * ntr + j] = S.tstate[j];
    }
}

size_t pde_smem_bytes(int L, int bc, int n_tracers) {
    if (L < 8 || L > kPdeMaxL || n_tracers < 0 || n_tracers > kPdeMaxTracers) return 0;
    size_t d = (size_t)6 * L + (bc == APS_PDE_BC_NEUMANN ? (size_t)2 * L : 0) + (size_t)n_tracers + 32 + 64;
    return d * 8 + (((size_t)n_tracers + 15) & ~(size_t)15) + 16;
}

cudaError_t pde_launch(const aps_pde_args& a, cudaStream_t st) {
    const size_t smem = pde_smem_bytes(a.L, a.bc, a.n_tracers);
    cudaError_t e = cudaFuncSetAttribute(pde_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pde_kernel<<<a.n_runs, kPdeThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace aps
