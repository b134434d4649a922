/* aps_pde.h — C ABI of the batched IMEX hydrodynamic-PDE stepper.
 *
 * Replaces, for a BATCH of independent solver instances (one CTA per instance, all time steps inside one launch),
 *     IMEXPDE.step()           IMEX_PDE_solver_class.py:187-233   implicit diffusion, upwind advection, reaction,
 *                                                                  clipping, mass renormalisation
 *     IMEXPDE.magnetization()  :157-169                            pointwise or kernel-convolved local magnetisation
 *     IMEXPDE.solve()          :236-289                            per-step diagnostics (m_series, var_series),
 *                                                                  snapshots, tracer particles (v_eff / D_eff series)
 * The reference is Python (numpy / scipy.sparse spsolve / numpy.fft); a maintainer binds this with ctypes exactly
 * like include/aps.h.  Same conventions: plain C types, caller-owned DEVICE buffers, a cudaStream_t
 * passed as void*, integer status codes (aps_status of aps.h), no CPU fallback.
 *
 * Numerics: fp64.  The cyclic tridiagonal solve of the implicit diffusion is done by two first-order recursive
 * filters (parallel scan), the kernel convolution by a direct ring sum; results agree with the reference's
 * spsolve / FFT formulation to rounding (tests state 1e-9 relative over O(1e3) steps), not bit for bit.
 * The tracer noise comes from Philox4x32-10 (counter = (step, tracer), key = seed): the reference draws it from
 * numpy's global stream inside the time loop, which no parallel kernel can replay; tracer statistics
 * (v_eff_series, D_eff_series) are therefore equal in distribution only.
 */
#ifndef APS_PDE_H
#define APS_PDE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APS_PDE_BC_PERIODIC 0
#define APS_PDE_BC_NEUMANN 1
#define APS_PDE_MODEL_BIDIRECTIONAL 0      /* active_model="bidirectional"  (:192-206) */
#define APS_PDE_MODEL_ANCHORED_MINUS 1     /* active_model="anchored_minus" (:207-229) */
#define APS_PDE_FIELD_POINTWISE 0          /* gaussian_kernel=False         (:158-161) */
#define APS_PDE_FIELD_KERNEL 1             /* ring-kernel convolution       (:166-169) */

typedef struct aps_pde_args {
    int32_t L;                  /* grid points, 8 <= L <= 4096                                              */
    int32_t n_runs;             /* solver instances in this launch (one CTA each)                           */
    int32_t bc, model, field;   /* APS_PDE_BC_*, APS_PDE_MODEL_*, APS_PDE_FIELD_*                            */
    int32_t snapshot_interval;  /* snapshot rows are written for n % snapshot_interval == 0                  */
    int32_t n_tracers;          /* 0 = no tracers                                                           */
    int32_t window;             /* tracer displacement window in steps, int(0.05/dt) in the reference (:238) */
    int64_t nsteps;             /* the loop runs n = 0..nsteps (nsteps calls of step())                      */
    double dt, dx, xlim;
    const double* beta;         /* [n_runs]                                                                 */
    const double* lam;          /* [n_runs]                                                                 */
    const double* gamma;        /* [n_runs]                                                                 */
    const double* kernel;       /* [n_runs][L] normalised ring kernel (kernel[j] at ring distance j), FIELD_KERNEL */
    const int32_t* radius;      /* [n_runs] taps kept on each side; >= L/2 means the full ring                */
    const uint64_t* seeds;      /* [n_runs] tracer noise keys                                               */
    double* rho_p;              /* [n_runs][L] in: initial state, out: final state                          */
    double* rho_m;              /* [n_runs][L]                                                              */
    double* m_series;           /* [n_runs][nsteps+1] mean of the local magnetisation (:244)                */
    double* var_series;         /* [n_runs][nsteps+1] np.var(rho_p + rho_m) (:245)                          */
    double* snapshots;          /* [n_runs][n_snap][L] rho_p + rho_m, n_snap = nsteps/snapshot_interval + 1, may be NULL */
    double* m_snapshots;        /* [n_runs][n_snap][L] rho_p - rho_m, may be NULL                           */
    double* tracer_pos;         /* [n_runs][n_tracers] unwrapped positions, in/out                          */
    int8_t* tracer_state;       /* [n_runs][n_tracers] +-1, in/out                                          */
    double* tracer_hist;        /* [n_runs][window][n_tracers] scratch ring of unwrapped positions          */
    double* v_eff_series;       /* [n_runs][nsteps+1], NaN where undefined (:274-282)                       */
    double* D_eff_series;       /* [n_runs][nsteps+1]                                                       */
    double* tot_series;         /* [n_runs][nsteps+1][L] rho_p + rho_m of EVERY step, or NULL: input of the per-step
                                   spectra fft_amp / fft_phase of solve() (:247-249), transformed by the caller    */
} aps_pde_args;

/* Runs every instance from its initial state through nsteps steps (device pointers, enqueued on `stream`). */
int aps_pde_solve_device(const aps_pde_args* args, void* stream);
/* Dynamic shared memory one instance needs for a given L (0 if unsupported). */
int64_t aps_pde_smem_bytes(int32_t L, int32_t bc);

#ifdef __cplusplus
}
#endif
#endif /* APS_PDE_H */
