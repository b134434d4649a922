/* aps.h — C ABI of the B200-native active-exclusion-process stepper.
 *
 * This is the drop-in boundary for ONE hot path of the reference
 * (StandeHaas/Hydrodynamic-Limits-of-Active-Particle-Systems-with-Mean-Field-Interactions):
 * the Gillespie time-stepping loop of `ParticleSystem`
 *     run()                     PARTICLE_solver_CLASS.py:450-558
 *     step_gillespie()          PARTICLE_solver_CLASS.py:254-448
 *     compute_local_m_field()   PARTICLE_solver_CLASS.py:216-246
 *     init_particles()          PARTICLE_solver_CLASS.py:141-195
 *     empirical_densities_from_particles()   PARTICLE_solver_CLASS.py:198-214
 * and the ensemble loops / per-run reducers of the sweep drivers
 *     sweep_beta_ensemble()     PARTICLE_solver_BIOLOGY_EXCLUSION_sweep_beta.py:56-117
 *     compute_v_eff_and_window / compute_blocking_probability / compute_mean_magnetizatoin /
 *     compute_rho_eff / compute_D_eff_active     ...sweep_beta.py:123-229,316-319,500-525
 *
 * The reference is pure Python, so there is no existing FFI to mirror; a maintainer binds this
 * library with ctypes.  Conventions:
 *   - plain C types, caller-allocated buffers, no exceptions; every entry point returns an
 *     `aps_status` (0 = OK) and leaves a message retrievable with aps_last_error();
 *   - `*_device` entry points take DEVICE pointers and a cudaStream_t (as void*), enqueue
 *     kernels and return without synchronising;
 *   - `*_host` entry points take HOST pointers, stage through device memory, synchronise and
 *     copy results back (this is the path the end-to-end benchmark times);
 *   - the library never falls back to the CPU: with no usable sm_100 device every compute call
 *     fails with APS_ERR_NO_DEVICE.
 */
#ifndef APS_H
#define APS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APS_ABI_VERSION 5

typedef enum aps_status {
    APS_OK = 0,
    APS_ERR_INVALID = 1,      /* bad argument / unsupported configuration            */
    APS_ERR_NO_DEVICE = 2,    /* no CUDA device (the product has no CPU path)         */
    APS_ERR_CUDA = 3,         /* CUDA runtime error, text in aps_last_error()         */
    APS_ERR_CAPACITY = 4      /* replica does not fit the kernel's shared-memory plan */
} aps_status;

/* per-replica completion codes written to aps_batch.status[] */
#define APS_RUN_DONE 0          /* loop ended as the reference's does (t>=T or all M rows filled) */
#define APS_RUN_EMPTY 1         /* n == 0 or R <= 0: the reference raises here (CLASS.py:257,355) */
#define APS_RUN_DRAWS_EXHAUSTED 2 /* replay log ran out before the loop ended                     */
#define APS_RUN_MAX_EVENTS 3    /* stopped by aps_batch.max_events                                */

/* aps_params.flags */
#define APS_FLAG_CROWDING 1u    /* crowding_suppresses_rates=True (CLASS.py:322-336)              */
#define APS_FLAG_SUPPRESS_FLIP_BOUND 2u  /* suppress_flip_when_bound=True (CLASS.py:266-267)      */
#define APS_FLAG_IMMOBILIZE 4u  /* immobilize_when_anchored=True (CLASS.py:307-312,338-340)       */
#define APS_FLAG_PERIODIC 8u    /* periodic=True: hops wrap (CLASS.py:278-288), the field is the   */
                                /* circular convolution with `weights` (truncated periodic kernel,  */
                                /* :108-122,224-227; needs 2*radius+1 <= L)                         */

/* aps_batch.record */
#define APS_REC_COUNTS 1u       /* obs_cp / obs_cm                                                */
#define APS_REC_POS 2u          /* obs_pos                                                        */
#define APS_REC_MLOCAL 4u       /* obs_m_local (pre-event field, CLASS.py:512,525)                */

/* event kinds in the optional trace */
#define APS_EV_DIFF_LEFT 0
#define APS_EV_DIFF_RIGHT 1
#define APS_EV_ACTIVE 2
#define APS_EV_FLIP 3
#define APS_EV_BIND 4           /* anchors only: bound[i] = True   (CLASS.py:418-419)             */
#define APS_EV_UNBIND 5         /* bound[i] = False                (CLASS.py:421-422)             */
#define APS_EV_EXIT 6           /* particle leaves the system      (CLASS.py:424-436)             */

/* Model parameters shared by every replica of one launch (constructor arguments of
 * ParticleSystem after the optional dx-rescaling, CLASS.py:41-50,84). */
typedef struct aps_params {
    int32_t L;              /* lattice sites                                                      */
    int32_t K;              /* site_capacity                                                      */
    int32_t radius;         /* Gaussian filter radius lw = int(4*sigma/dx + 0.5); -1 selects the  */
                            /* global mean field (local_kernel_sigma <= 0, CLASS.py:219-221)      */
    uint32_t flags;
    double rate_diffusion;  /* D                                                                  */
    double rate_active;     /* lambda (sigma=+1 particles hop right only, CLASS.py:276,317-319)   */
    double T;               /* run(T=...)                                                         */
    double k_on;            /* binding rate at anchor sites (only used when aps_batch.anchor_mask) */
    double k_off;           /* unbinding rate                                                     */
    double k_exit;          /* exit rate of bound '-' particles on anchor sites                   */
} aps_params;

/* One launch = n_replicas independent ParticleSystem.run() calls.  Optional pointers may be NULL. */
typedef struct aps_batch {
    int32_t n_replicas;
    int32_t n_max;              /* row stride of the per-particle arrays                          */
    int32_t M;                  /* len(np.arange(0, T, obs_dt))                                   */
    uint32_t record;            /* APS_REC_* mask                                                 */
    int64_t max_events;         /* 0 = unlimited; otherwise stop each replica after this many     */
    int64_t trace_cap;          /* events per replica the trace can hold (0 = no trace)           */
    int64_t spec_from;          /* replay: draws at offsets >= spec_from are SPECULATIVE triples  */
                                /* (e,u,u) with no direction variate behind them: a diffusive hop */
                                /* starting there stops the replica (APS_RUN_DRAWS_EXHAUSTED) so  */
                                /* the host can draw the 4th variate in rng order. -1 = log is    */
                                /* exact (default for recorded logs).                             */

    const double* times_obs;    /* [M]  exactly np.arange(0, T, obs_dt)                           */
    const double* weights;      /* [2*radius+1] normalised Gaussian taps (scipy _gaussian_kernel1d)*/
    const double* beta;         /* [n_replicas]                                                   */
    const int32_t* n;           /* [n_replicas] particle count of each replica (<= n_max)         */
    const int32_t* pos0;        /* [n_replicas][n_max] initial sites, reference particle order    */
    const int8_t* sigma0;       /* [n_replicas][n_max] +1 / -1                                    */

    /* replay mode: the variates the reference's rng returned, in call order:
     * per event  e (standard exponential), u_choice, u_event [, u_dir if the event is diffusive] */
    const double* draws;        /* flat log                                                       */
    const int64_t* draw_off;    /* [n_replicas+1] offsets into draws                              */
    /* native mode */
    const uint64_t* seeds;      /* [n_replicas] Philox keys                                       */
    /* optional resume point (all NULL = fresh run: t=0, observation row 0 recorded at start)    */
    const double* t_start;      /* [n_replicas] simulation clock to resume from                   */
    const int32_t* obs_start;   /* [n_replicas] next observation row; 0 = record row 0 first      */
    const int64_t* ev_start;    /* [n_replicas] events already executed (Philox counter base)     */

    /* observation outputs (row m of replica r is written when observation m is reached) */
    int8_t* obs_cp;             /* [n_replicas][M][L] + particles per site                        */
    int8_t* obs_cm;             /* [n_replicas][M][L] - particles per site                        */
    int32_t* obs_pos;           /* [n_replicas][M][n_max]                                         */
    int32_t* obs_sigma_sum;     /* [n_replicas][M]  sum(sigma)  (m_global = sum/n, CLASS.py:526)   */
    double* obs_m_local;        /* [n_replicas][M][L]                                             */

    /* per-replica results */
    int32_t* n_obs;             /* [n_replicas] rows written                                      */
    int64_t* n_events;          /* [n_replicas] ev_start + events executed (incl. the one past T) */
    double* t_end;              /* [n_replicas] simulation clock after the last event             */
    int32_t* status;            /* [n_replicas] APS_RUN_*                                         */
    int64_t* n_guard;           /* [n_replicas] selections resolved by the exact slow path        */
    int64_t* draws_used;        /* [n_replicas] replay variates consumed by completed events      */
    int32_t* pos_end;           /* [n_replicas][n_max] optional final state                       */
    int8_t* sigma_end;          /* [n_replicas][n_max]                                            */
    int32_t* trace;             /* [n_replicas][trace_cap][3] = (particle, kind, new_site)        */
    /* optional: use this magnetisation field instead of computing it (step_gillespie takes m_field
     * as an argument, CLASS.py:254,261); only meaningful with max_events = 1 */
    const double* m_field_in;   /* [n_replicas][L]                                                */
    /* anchors / binding / exit (CLASS.py:88-104,307-312,343-348,418-436); all NULL = no anchors         */
    const uint8_t* anchor_mask; /* [L] is_anchor_site, shared by all replicas                     */
    const int8_t* bound0;       /* [n_replicas][n_max] optional initial bound flags (resume)      */
    int32_t* n_end;             /* [n_replicas] particle count at the end (exits shrink it)       */
    int8_t* bound_end;          /* [n_replicas][n_max]                                            */
    int32_t* obs_n;             /* [n_replicas][M] particle count at each observation             */
    int8_t* obs_bound;          /* [n_replicas][M][n_max] bound flags at each observation         */
    double* exit_t;             /* [n_replicas][exit_cap] clock at the start of the exit step     */
    int32_t* exit_pos;          /* [n_replicas][exit_cap] site the particle left from             */
    int32_t* n_exit;            /* [n_replicas] exits recorded (in: previous count when resuming) */
    int64_t exit_cap;
    /* custom flip_rate_fn (CLASS.py:59-62): NULL = the default exp(-beta*sigma*m).  Otherwise the callable tabulated
     * by the caller on the grid m_k = -1 + 2k/flip_G, [2][flip_G+1] (sigma = +1 row first), shared by all replicas;
     * the kernels interpolate linearly (aps_flip_interp, aps_math.h: tolerance stated there).                        */
    const double* flip_tab;
    int64_t flip_G;
    /* optional HOST copy of `weights` (same 2*radius+1 values) for the `*_device` entry points: the specialised K1 kernel takes
     * the taps as kernel-parameter (constant-bank) operands of its unrolled filter instead of shared-memory loads.  NULL is
     * allowed (the taps are then read from shared memory); the `*_host` entry points fill it in themselves.             */
    const double* weights_host;
} aps_batch;

int aps_abi_version(void);
const char* aps_last_error(void);

/* Number of visible CUDA devices with compute capability 10.x; 0 means every compute call fails. */
int aps_device_count(void);
/* Bind the calling thread to a device (one process per GPU: call once with LOCAL_RANK). */
int aps_set_device(int device);

/* ---- K1: replica-batched exact Gillespie kernel ------------------------------------------- */
/* Replaces ParticleSystem.run (CLASS.py:450) for a batch of replicas, replaying injected draws. */
int aps_run_replay_device(const aps_params* p, const aps_batch* b, void* stream);
/* Same loop with the counter-based Philox stream of aps_philox.h. */
int aps_run_philox_device(const aps_params* p, const aps_batch* b, void* stream);
/* Host-buffer variants: allocate, copy in, run, copy out, free. */
int aps_run_replay_host(const aps_params* p, const aps_batch* b);
int aps_run_philox_host(const aps_params* p, const aps_batch* b);
/* Number of kernels the library has launched since load (for gpu_launches accounting). */
int64_t aps_launch_count(void);
/* Shared-memory bytes and threads the K1 plan uses for (L, n_max, radius); <0 if it cannot fit. */
int64_t aps_replica_smem_bytes(const aps_params* p, int32_t n_max);

/* ---- K3: device-side initial conditions for native-mode ensembles (init_particles, CLASS.py:141-195) */
typedef struct aps_init_args {
    int32_t n_replicas, L, K, n_max;
    int32_t mode;               /* 0 = 'fixed', 1 = 'poisson'                                     */
    int32_t N_fixed;            /* 'fixed': particles per replica (unless N_of is given)          */
    int32_t n_profiles, reserved;
    const double* rho0_plus;    /* 'poisson': [n_profiles][L] intensities rho0_plus(i/L)          */
    const double* rho0_minus;
    const int32_t* profile_of;  /* [n_replicas] profile index per replica, NULL = profile 0       */
    const int32_t* N_of;        /* [n_replicas] optional per-replica N for 'fixed'                */
    const uint64_t* seeds;      /* [n_replicas] Philox keys (same keys as the run)                */
    int32_t* pos0;              /* [n_replicas][n_max] out                                        */
    int8_t* sigma0;             /* [n_replicas][n_max] out                                        */
    int32_t* n;                 /* [n_replicas] out; -1 if the sample does not fit n_max          */
} aps_init_args;
int aps_init_particles_device(const aps_init_args* a, void* stream);

/* compute_local_m_field (CLASS.py:216-246) for one lattice; host buffers: counts int32[L] -> out double[L] */
int aps_m_field_host(const aps_params* p, const double* weights, const int32_t* counts_p, const int32_t* counts_m,
                     double* out);

/* ---- K4: on-device observables ------------------------------------------------------------ */
/* counts -> density rows, replaces empirical_densities_from_particles (CLASS.py:198-214) and the
 * per-row bookkeeping of run() (:489-507,:518-535).  Rows >= n_obs[rep] are left untouched. */
typedef struct aps_expand_args {
    int32_t n_replicas, M, L, reserved;
    double dx;                  /* ParticleSystem.dx = xlim / L                                   */
    const int32_t* n;           /* [n_replicas]                                                   */
    const int32_t* n_obs;       /* [n_replicas]                                                   */
    const int8_t* obs_cp;       /* [n_replicas][M][L]                                             */
    const int8_t* obs_cm;
    double* rho_p;              /* [n_replicas][M][L] optional                                    */
    double* rho_m;              /* optional                                                       */
    double* total;              /* optional                                                       */
    double* var;                /* [n_replicas][M] optional: np.var(total row), numpy's order     */
    const int32_t* obs_n;       /* [n_replicas][M] optional per-row particle count (exits); else n */
} aps_expand_args;
int aps_expand_obs_device(const aps_expand_args* a, void* stream);

/* per-run reducers of the sweep drivers (sweep_beta.py:123-229,316-319,500-525) */
#define APS_RED_V_EFF 0     /* mean of np.gradient(mean_x, times) over the window                 */
#define APS_RED_D_EFF 1     /* slope of the per-particle MSD (NaN if obs_pos is NULL)             */
#define APS_RED_M_MEAN 2    /* mean of m_global over the window                                   */
#define APS_RED_RHO_EFF 3   /* front density                                                      */
#define APS_RED_BLOCK 4     /* blocking probability                                               */
#define APS_RED_START 5     /* window [start, end)                                                */
#define APS_RED_END 6
#define APS_RED_NOBS 7
#define APS_RED_N 8
typedef struct aps_reduce_args {
    int32_t n_replicas, M, L, n_max;
    double dx;
    double boundary_xmin;           /* 0.99  (sweep_beta.py:85)                                   */
    double max_boundary_fraction;   /* 0.06                                                       */
    double min_window_fraction;     /* 0.10                                                       */
    double window_fraction;         /* 0.05  (compute_rho_eff default)                            */
    const double* times_obs;        /* [M]                                                        */
    const int32_t* n;
    const int32_t* n_obs;
    const int8_t* obs_cp;
    const int8_t* obs_cm;
    const int32_t* obs_pos;         /* optional                                                   */
    const int32_t* obs_sigma_sum;
    double* out;                    /* [n_replicas][APS_RED_N]                                    */
    double* v_eff;                  /* [n_replicas][M] optional (window rows only)                */
} aps_reduce_args;
int aps_reduce_runs_device(const aps_reduce_args* a, void* stream);

/* ensemble profile sums per grid point; prof is [n_points][4][L]: sum of time-averaged rho_plus, rho_minus
 * over the point's replicas, and the sums of their squares.  Replica membership: either grid-point-major
 * (replica = g*reps_per_point + j; point_start == NULL) or an explicit list per point (any replica order,
 * e.g. a rank's strided shard of a sweep): replicas point_reps[point_start[g] .. point_start[g+1]). */
typedef struct aps_profile_args {
    int32_t n_points, reps_per_point, M, L;
    int32_t row_lo, row_hi;
    double dx;
    const int32_t* n;
    const int32_t* n_obs;
    const int8_t* obs_cp;
    const int8_t* obs_cm;
    double* prof;
    const int32_t* point_start; /* [n_points+1] optional                                          */
    const int32_t* point_reps;  /* [point_start[n_points]] replica indices, ascending per point   */
    double* scratch;            /* optional [n_replicas][4][L]: with point lists, the replicas are reduced in parallel */
    int32_t n_replicas, reserved; /* into these rows first and then summed per point (same values, more parallelism) */
} aps_profile_args;
int aps_profile_sums_device(const aps_profile_args* a, void* stream);

/* Histogram of the per-replica time-averaged magnetisation (BASELINE north_star: "density and magnetisation
 * profile and histogram accumulation stays on device"; the sample of the KS test of the native-mode parity check).
 * mbar[r] = mean over rows [row_lo, min(row_hi, n_obs[r])) of m_global = sum(sigma)/n (CLASS.py:526, the quantity
 * compute_mean_magnetizatoin averages, sweep_beta.py:316-319); bin = floor((mbar - lo) / (hi - lo) * n_bins)
 * clamped to [0, n_bins-1]; hist[point_of[r]][bin] += 1 (integer atomics: order-independent, all-reducible). */
typedef struct aps_hist_args {
    int32_t n_replicas, M, n_points, n_bins;
    int32_t row_lo, row_hi;
    int32_t accumulate, reserved;   /* 0: the call zeroes hist first (cudaMemsetAsync); 1: add to its contents */
    double lo, hi;                  /* histogram range, normally [-1, 1]                           */
    const int32_t* n;               /* [n_replicas]                                               */
    const int32_t* n_obs;           /* [n_replicas]                                               */
    const int32_t* obs_sigma_sum;   /* [n_replicas][M]                                            */
    const int32_t* obs_n;           /* [n_replicas][M] optional per-row particle count (exits)    */
    const int32_t* point_of;        /* [n_replicas] optional grid point of every replica (else 0) */
    double* mbar;                   /* [n_replicas] optional                                      */
    unsigned long long* hist;       /* [n_points][n_bins]                                          */
} aps_hist_args;
int aps_m_histogram_device(const aps_hist_args* a, void* stream);

/* ---- K2: sublattice-parallel kernel for one lattice too large for shared memory -------------- */
#include "aps_k2_model.h"
typedef struct aps_k2_args {
    int64_t L;                  /* sites in this call's buffers (multiple of 8192); a slab incl. ghosts */
    int64_t L_global;           /* sites of the whole lattice (profile binning)                     */
    int64_t global_offset;      /* global index of buffer site 0 (multiple of 8192); keys the RNG   */
    int64_t n_particles;        /* global-field mode: N                                             */
    uint64_t seed;
    uint64_t pass;              /* pass counter; parity = pass & 1; two passes advance time by dt   */
    int32_t radius;             /* local Gaussian field radius in sites; -1 = global magnetisation  */
    int32_t reserved;
    aps_k2_rates rates;         /* aps_k2_make_rates(D, lambda, beta, dt)                            */
    const int32_t* w16;         /* [radius+1] taps round(w_j * 65536), centre first (device)         */
    const uint32_t* flip_tab;   /* local field: [2][2*512+1] acceptance thresholds (aps_k2_flip_table) */
    const uint8_t* in;          /* [L] site bytes 0/1/2                                              */
    uint8_t* out;               /* [L] ping-pong target                                              */
    const int64_t* msum_in;     /* global-field mode: sum(sigma) at pass start (device)             */
    int64_t* msum_out;          /* must hold *msum_in on entry; receives the flips' increments      */
    int64_t count_lo, count_hi; /* slab decomposition: only flips at buffer sites [count_lo, count_hi) are added to   */
                                /* msum_out (the caller all-reduces the increments of the ranks); 0, 0 = whole buffer */
} aps_k2_args;
/* thresholds / Poisson table for (D, lambda, beta, dt); returns non-zero if B*32*dt is out of (0, 24] */
int aps_k2_rates_init(double D, double lam, double beta, double dt, aps_k2_rates* out);
/* local-field acceptance thresholds: out[s][i] for sigma = +1 (s=0) / -1 (s=1) and m = (i-512)/512; host buffer of 2*1025 */
int aps_k2_flip_table(const aps_k2_rates* rates, uint32_t* out);   /* needs the rate thresholds: acceptance is read off the flip slot */
/* one pass (in -> out) */
int aps_k2_pass_device(const aps_k2_args* a, void* stream);
/* n_passes passes ping-ponging between a->in and a->out (a->pass, in/out and msum are advanced in the
 * struct); the final state is in a->in after the call.  msum buffers are handled internally. */
int aps_k2_run_device(aps_k2_args* a, int n_passes, void* stream);
/* ---- K2, many passes per launch and slab decomposition over peer memory (NVLink) ----------------------------
 * One cooperative launch runs n_passes passes (grid barriers between passes instead of kernel boundaries).  With
 * world > 1 (one process per GPU; contiguous slabs with `ghost` redundant sites per interior side) the kernel itself
 * exchanges the ghost zones (every refresh_every passes) and, in global-field mode, the 8-byte sum(sigma) increments
 * (every pass) through peer memory mapped with CUDA IPC: release/acquire flag words, no NCCL call, no host round trip.
 * The result is bit-identical to the single-slab run.  See csrc/aps_k2.cuh (K2PeerRegion, K2Multi). */
typedef struct aps_k2_multi {
    int32_t n_passes, world, rank, refresh_every;   /* refresh_every: passes between ghost refreshes (world > 1)      */
    int64_t ghost, own_lo, own_hi;                  /* ghost sites per interior side; owned range in buffer indices  */
    uint8_t* buf0;                                  /* pass j of the launch reads buf[j & 1], writes buf[(j+1) & 1]  */
    uint8_t* buf1;
    void* sync;                                     /* device int64[8], zero at creation: [0] grid-barrier counter,   */
                                                    /* [1] error flag (peer time-out), [2] cumulative own flip sum,  */
                                                    /* [3] sum(sigma) at creation, [4] current lattice-wide sum(sigma) */
    void* peer[8];                                  /* exchange regions of all ranks (aps_k2_peer_*), world > 1 only  */
} aps_k2_multi;
/* a->pass is advanced by n_passes; the final state is in buf[n_passes & 1].  a->in/out/msum_* are ignored. */
int aps_k2_run_persistent_device(aps_k2_args* a, const aps_k2_multi* m, void* stream);
/* exchange region of this rank: device memory + its 64-byte CUDA IPC handle; peers map it with aps_k2_peer_open */
int aps_k2_peer_region_bytes(void);
int aps_k2_peer_alloc(void** region, void* ipc_handle_out_64_bytes);
int aps_k2_peer_open(const void* ipc_handle_64_bytes, void** region);
int aps_k2_peer_close(void* region);
int aps_k2_peer_free(void* region);

/* Bernoulli(density) occupancy, '+' with probability frac_plus */
int aps_k2_init_device(uint8_t* state, int64_t L, int64_t global_offset, uint64_t seed, double density, double frac_plus,
                       void* stream);
/* counts of '+' / '-' per coarse bin (uint64 accumulators, caller zeroes them) */
int aps_k2_profile_device(const uint8_t* state, int64_t L, int64_t global_offset, int64_t L_global, int32_t nbins,
                          uint64_t* cnt_plus, uint64_t* cnt_minus, void* stream);

/* Test / tuning hooks: widen the selection guard band (forces the exact serial slow path) and
 * override the K1 block size (32, 64, 128, 256; 0 = heuristic). Not needed in production.
 * Environment knobs read ONCE when the library is loaded, for A/B measurements only (results never depend on them):
 *   APS_K1_THREADS=32|64|128|256   block size of the K1 kernels
 *   APS_K1_NO_LEAN=1               skip the half-image K1 kernel (aps_k1_lean.cuh), use the full-size one
 *   APS_K1_EXTRA_SMEM=<bytes>      pad the full-size K1 kernel's shared memory (occupancy experiments)
 *   APS_K2_NO_SHORT_PLAN=1         K2: always the deep TMA ring at 6 CTAs per SM (no 2-deep ring for short slabs)
 *   APS_K2_STAGES=<n>              K2: ring depth of the first launch plan (default 3 local / 4 global field) */
void aps_debug_set_guard_scale(double scale);
void aps_debug_set_k1_threads(int threads);
void aps_debug_set_use_lut(int on); /* 0: evaluate filter taps arithmetically instead of by table */
void aps_debug_set_k2_ctas_per_sm(int n); /* persistent K2 CTAs per SM (default 6) */
void aps_debug_set_reduce_impl(int v);    /* 0: integer row sums (default); 1: the round-1 per-site double kernel (A/B tests) */
void aps_debug_set_reduce_threads(int n); /* threads per CTA of the reducer kernel (multiple of 32) */
void aps_debug_set_k2_stash_cap(int n); /* local-field K2: stashed trials per segment (1..32, power of two; 0 = automatic) */
void aps_debug_set_use_fast(int on); /* 0: always use the generic K1 kernel (no K=1 specialisation) */

#ifdef __cplusplus
}
#endif
#endif /* APS_H */
