/* aps_oracle.c — CPU restatement of the reference's particle time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may link, import or execute this
 * file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm as the checker and the CPU baseline.
 *
 * It follows the reference's algorithm literally — every event recomputes the full
 * magnetisation field over all L sites and all n per-particle rates, exactly as
 * ParticleSystem.run does — and reproduces the reference's floating-point evaluation order:
 *   compute_local_m_field   PARTICLE_solver_CLASS.py:216-246   (scipy correlate1d symmetric branch)
 *   step_gillespie          PARTICLE_solver_CLASS.py:254-448   (rates :351, R :352 numpy pairwise
 *                                                               sum, choice :360 = cumsum/searchsorted)
 *   run                     PARTICLE_solver_CLASS.py:450-558   (observation semantics :511-539)
 * Parity pin: the reference has no tests or golden vectors, so this restatement
 * is pinned against outputs of the unmodified reference class executed in the build container
 * (tools/gen_golden.py -> the npz fixtures under tests/golden/; checked in tests/test_oracle_golden.py).
 * The only deliberate deviation: exp() is include/aps_math.h's aps_exp instead of numpy's
 * np.exp (< 1 ulp apart; the discrete trajectories are identical on every fixture).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (see oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/aps.h"
#include "../include/aps_math.h"
#include "../include/aps_philox.h"
#include "../include/aps_sampling.h"

/* ---- numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE),
 *      what `rates.sum()` does at CLASS.py:352 ------------------------------------------------ */
static double pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}
double aps_oracle_pairwise_sum(const double* a, int64_t n) { return pairwise_sum(a, n); }

double aps_oracle_native_total(const double* rates, int n) { return aps_native_total(rates, n); }   /* native-mode R (aps_math.h) */
double aps_oracle_exp(double x) { return aps_exp(x); }
double aps_oracle_log(double x) { return aps_log(x); }
void aps_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    aps_u32x4 r = aps_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

/* scipy NI_EXTEND_REFLECT ("d c b a | a b c d | d c b a"), repeated when radius >= L */
static inline int reflect_index(int64_t i, int L) {
    int64_t p = 2 * (int64_t)L;
    int64_t m = i % p;
    if (m < 0) m += p;
    if (m >= L) m = p - 1 - m;
    return (int)m;
}

/* scipy.ndimage.correlate1d, symmetric branch (ni_filters.c), with mode='reflect':
 *   out[l] = x[l]*w[r];  for j = -r..-1: out[l] += (x[l+j] + x[l-j]) * w[r+j]   (no FMA)      */
static void gaussian_filter_reflect(const double* x, int L, int r, const double* w,
                                    double* pad, double* out, int periodic) {
    for (int i = 0; i < L + 2 * r; ++i) {
        int64_t q = (int64_t)i - r;
        pad[i] = periodic ? x[((q % L) + L) % L] : x[reflect_index(q, L)];
    }
    const double wc = w[r];
    for (int l = 0; l < L; ++l) out[l] = pad[l + r] * wc;
    for (int jj = -r; jj < 0; ++jj) {
        const double wj = w[r + jj];
        const double* a = pad + r + jj;
        const double* b = pad + r - jj;
        for (int l = 0; l < L; ++l) {
            double pair = a[l] + b[l];
            double term = pair * wj;
            out[l] = out[l] + term;
        }
    }
}

typedef struct work {
    int L, r;
    double *s, *tot, *pad, *sconv, *tconv, *m;
    double *r_left, *r_right, *r_diff, *r_act, *cvec, *rates, *cdf;
    double *r_bind, *r_unbind, *r_exit;
    int32_t *cp, *cm;
    int64_t* pos;
    int8_t* sigma;
    int8_t* bound;
} work;

static int work_alloc(work* w, int L, int r, int n_max) {
    memset(w, 0, sizeof(*w));
    w->L = L; w->r = r;
    int rr = r < 0 ? 0 : r;
    size_t nl = (size_t)L, np = (size_t)(n_max > 0 ? n_max : 1);
    w->s = malloc(8 * nl); w->tot = malloc(8 * nl); w->pad = malloc(8 * (nl + 2 * (size_t)rr));
    w->sconv = malloc(8 * nl); w->tconv = malloc(8 * nl); w->m = malloc(8 * nl);
    w->r_left = malloc(8 * np); w->r_right = malloc(8 * np); w->r_diff = malloc(8 * np);
    w->r_act = malloc(8 * np); w->cvec = malloc(8 * np); w->rates = malloc(8 * np); w->cdf = malloc(8 * np);
    w->cp = malloc(4 * nl); w->cm = malloc(4 * nl);
    w->pos = malloc(8 * np); w->sigma = malloc(np); w->bound = calloc(np, 1);
    w->r_bind = malloc(8 * np); w->r_unbind = malloc(8 * np); w->r_exit = malloc(8 * np);
    return (w->bound && w->r_bind && w->r_unbind && w->r_exit && w->s && w->tot && w->pad && w->sconv && w->tconv && w->m && w->r_left && w->r_right &&
            w->r_diff && w->r_act && w->cvec && w->rates && w->cdf && w->cp && w->cm && w->pos && w->sigma) ? 0 : -1;
}
static void work_free(work* w) {
    free(w->s); free(w->tot); free(w->pad); free(w->sconv); free(w->tconv); free(w->m);
    free(w->r_left); free(w->r_right); free(w->r_diff); free(w->r_act); free(w->cvec); free(w->rates);
    free(w->cdf); free(w->cp); free(w->cm); free(w->pos); free(w->sigma);
    free(w->bound); free(w->r_bind); free(w->r_unbind); free(w->r_exit);
}

/* compute_local_m_field, CLASS.py:216-246 */
static void m_field(work* w, const aps_params* P, const double* weights) {
    const int L = P->L;
    for (int l = 0; l < L; ++l) {
        w->s[l] = (double)w->cp[l] - (double)w->cm[l];
        w->tot[l] = (double)w->cp[l] + (double)w->cm[l];
    }
    if (P->radius < 0) {
        double mg = pairwise_sum(w->s, L) / pairwise_sum(w->tot, L);
        for (int l = 0; l < L; ++l) w->m[l] = mg;
        return;
    }
    const int periodic = (P->flags & APS_FLAG_PERIODIC) != 0;   /* circular convolution, same tap order */
    gaussian_filter_reflect(w->s, L, P->radius, weights, w->pad, w->sconv, periodic);
    gaussian_filter_reflect(w->tot, L, P->radius, weights, w->pad, w->tconv, periodic);
    for (int l = 0; l < L; ++l) {
        double m = 0.0;
        if (w->tconv[l] > 0) m = w->sconv[l] / w->tconv[l];
        if (m < -1.0) m = -1.0;
        if (m > 1.0) m = 1.0;
        w->m[l] = m;
    }
}
/* exported for unit tests: counts (int32) -> m_field */
int aps_oracle_m_field(const aps_params* P, const double* weights, const int32_t* cp, const int32_t* cm,
                       double* out) {
    work w;
    if (work_alloc(&w, P->L, P->radius, 1)) return -1;
    memcpy(w.cp, cp, 4 * (size_t)P->L);
    memcpy(w.cm, cm, 4 * (size_t)P->L);
    m_field(&w, P, weights);
    memcpy(out, w.m, 8 * (size_t)P->L);
    work_free(&w);
    return 0;
}

static inline int clipi(int64_t v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }

typedef struct draw_src {
    int mode;               /* 0 replay, 1 philox */
    const double* d; int64_t left;
    uint32_t k0, k1; uint64_t ev;
    double a_choice, b_event, b_dir;
} draw_src;

/* per-particle rates, CLASS.py:259-351 restricted to anchors=None (bind/unbind/exit rates are 0) */
static double build_rates(work* w, const aps_params* P, int n, double beta, const uint8_t* anchor, const double* flip_tab,
                          int64_t flip_G) {
    const int L = P->L, K = P->K;
    const int suppress = (P->flags & APS_FLAG_SUPPRESS_FLIP_BOUND) != 0, immob = (P->flags & APS_FLAG_IMMOBILIZE) != 0;
    const double D = P->rate_diffusion, lam = P->rate_active;
    const int crowd = (P->flags & APS_FLAG_CROWDING) != 0;
    const int periodic = (P->flags & APS_FLAG_PERIODIC) != 0;
    for (int i = 0; i < n; ++i) {
        int p = (int)w->pos[i];
        int sg = w->sigma[i];
        int fwd = periodic ? (p + (sg == 1)) % L : clipi((int64_t)p + (sg == 1), 0, L - 1);   /* :278-291 */
        int lt = periodic ? (p - 1 + L) % L : clipi((int64_t)p - 1, 0, L - 1);
        int rt = periodic ? (p + 1) % L : clipi((int64_t)p + 1, 0, L - 1);
        int occ_f = w->cp[fwd] + w->cm[fwd], occ_l = w->cp[lt] + w->cm[lt], occ_r = w->cp[rt] + w->cm[rt];
        int f_free = (occ_f < K) && (fwd != p);
        int l_free = (occ_l < K) && (lt != p);
        int r_free = (occ_r < K) && (rt != p);
        double rl = D * (double)l_free, rr = D * (double)r_free;
        double ra = (sg == 1 && f_free) ? lam : 0.0;
        const int bnd = anchor ? w->bound[i] : 0, on_anchor = anchor ? anchor[p] : 0;
        const int anchored = immob && sg == -1 && on_anchor && bnd;        /* :308 */
        double rex = 0.0;
        if (anchored) { ra = 0.0; rl = 0.0; rr = 0.0; rex = P->k_exit; }   /* :309-312 */
        if (crowd) {
            double ff = 1.0 - ((double)occ_f / (double)K);
            ff = ff < 0.0 ? 0.0 : (ff > 1.0 ? 1.0 : ff);
            ra = ra * ff;
            double lf = 1.0 - ((double)occ_l / (double)K), rf = 1.0 - ((double)occ_r / (double)K);
            lf = lf < 0.0 ? 0.0 : (lf > 1.0 ? 1.0 : lf);
            rf = rf < 0.0 ? 0.0 : (rf > 1.0 ? 1.0 : rf);
            rl = (D * (double)l_free) * lf;
            rr = (D * (double)r_free) * rf;
        }
        w->r_left[i] = rl; w->r_right[i] = rr;
        w->r_diff[i] = anchored ? 0.0 : rl + rr;                           /* :315,336,339 */
        w->r_act[i] = anchored ? 0.0 : ra;                                 /* :340 */
        /* flip_rate_fn default: np.exp(-beta * sigma * m), CLASS.py:60 */
        double arg = ((-beta) * (double)sg) * w->m[p];
        /* custom flip_rate_fn: the caller's callable tabulated on the host, linear interpolation (aps_flip_interp) */
        const double cv = flip_tab ? aps_flip_interp(flip_tab, flip_G, sg, w->m[p]) : aps_exp(arg);
        w->cvec[i] = (suppress && bnd) ? 0.0 : cv;                         /* :266-267 */
        int occ_here = w->cp[p] + w->cm[p];
        w->r_bind[i] = (anchor && !bnd && sg == -1 && on_anchor && occ_here < K) ? P->k_on : 0.0;   /* :343-345 */
        w->r_unbind[i] = bnd ? P->k_off : 0.0;                             /* :347-348 */
        w->r_exit[i] = rex;
        w->rates[i] = ((((w->r_diff[i] + w->r_act[i]) + w->cvec[i]) + w->r_bind[i]) + w->r_unbind[i]) + w->r_exit[i];
    }
    return pairwise_sum(w->rates, n);
}

static void record_obs(const aps_params* P, const aps_batch* B, int rep, int m, const work* w, int n,
                       const double* mfield_pre) {
    const size_t L = (size_t)P->L;
    size_t row = ((size_t)rep * (size_t)B->M + (size_t)m);
    if ((B->record & APS_REC_COUNTS) && B->obs_cp && B->obs_cm) {
        for (size_t l = 0; l < L; ++l) {
            B->obs_cp[row * L + l] = (int8_t)w->cp[l];
            B->obs_cm[row * L + l] = (int8_t)w->cm[l];
        }
    }
    if ((B->record & APS_REC_POS) && B->obs_pos) {
        int32_t* dst = B->obs_pos + row * (size_t)B->n_max;
        for (int i = 0; i < n; ++i) dst[i] = (int32_t)w->pos[i];
    }
    if (B->obs_sigma_sum) {
        int32_t s = 0;
        for (int i = 0; i < n; ++i) s += w->sigma[i];
        B->obs_sigma_sum[row] = s;
    }
    if ((B->record & APS_REC_MLOCAL) && B->obs_m_local) memcpy(B->obs_m_local + row * L, mfield_pre, 8 * L);
    if (B->obs_n) B->obs_n[row] = n;
    if (B->obs_bound) for (int i = 0; i < n; ++i) B->obs_bound[row * (size_t)B->n_max + i] = w->bound[i];
}

/* one ParticleSystem.run(), CLASS.py:450-558 */
static void run_one(const aps_params* P, const aps_batch* B, int rep, int mode, work* w) {
    const int L = P->L, M = B->M;
    int n = B->n[rep];
    const uint8_t* anchor = B->anchor_mask;
    int n_exit = (B->n_exit && B->ev_start) ? B->n_exit[rep] : 0;
    const double beta = B->beta[rep], T = P->T;
    int64_t n_events = B->ev_start ? B->ev_start[rep] : 0;
    const int64_t ev_base = n_events;
    int status = APS_RUN_DONE;
    double t = B->t_start ? B->t_start[rep] : 0.0;
    int obs_idx = B->obs_start ? B->obs_start[rep] : 0;
    const double* draws_begin = NULL;

    draw_src ds; memset(&ds, 0, sizeof(ds));
    ds.mode = mode;
    if (mode == 0) { ds.d = B->draws + B->draw_off[rep]; ds.left = B->draw_off[rep + 1] - B->draw_off[rep]; draws_begin = ds.d; }
    else { ds.k0 = (uint32_t)B->seeds[rep]; ds.k1 = (uint32_t)(B->seeds[rep] >> 32); ds.ev = (uint64_t)n_events; }

    memset(w->cp, 0, 4 * (size_t)L); memset(w->cm, 0, 4 * (size_t)L);
    for (int i = 0; i < n; ++i) {
        w->pos[i] = B->pos0[(size_t)rep * B->n_max + i];
        w->sigma[i] = B->sigma0[(size_t)rep * B->n_max + i];
        w->bound[i] = B->bound0 ? B->bound0[(size_t)rep * B->n_max + i] : 0;
        if (w->sigma[i] == 1) w->cp[w->pos[i]]++; else w->cm[w->pos[i]]++;
    }
    if (n == 0) { status = APS_RUN_EMPTY; goto done; }

    /* observation 0, CLASS.py:489-508 */
    m_field(w, P, B->weights);
    if (M > 0 && obs_idx == 0) { record_obs(P, B, rep, 0, w, n, w->m); obs_idx = 1; }

    while (t < T) {                                                      /* :511 */
        if (B->max_events > 0 && n_events - ev_base >= B->max_events) { status = APS_RUN_MAX_EVENTS; break; }
        if (B->m_field_in) memcpy(w->m, B->m_field_in + (size_t)rep * (size_t)L, 8 * (size_t)L);
        else m_field(w, P, B->weights);                                  /* :512 */
        if (n == 0) { status = APS_RUN_EMPTY; break; }                   /* :256-257 (every particle has exited) */
        double R = build_rates(w, P, n, beta, anchor, B->flip_tab, B->flip_G);   /* :259-352 */
        if (mode != 0) R = aps_native_total(w->rates, n);                /* native mode: the selection scan's total (aps_math.h) */
        if (!(R > 0)) { status = APS_RUN_EMPTY; break; }                 /* :353-355 */
        double e, u_choice, u_event;
        if (mode == 0) {
            if (ds.left < 3) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
            e = ds.d[0]; u_choice = ds.d[1]; u_event = ds.d[2];
        } else {
            aps_u32x4 a = aps_philox4x32_10((uint32_t)ds.ev, (uint32_t)(ds.ev >> 32), APS_RNG_EVENT_A, 0, ds.k0, ds.k1);
            aps_u32x4 b = aps_philox4x32_10((uint32_t)ds.ev, (uint32_t)(ds.ev >> 32), APS_RNG_EVENT_B, 0, ds.k0, ds.k1);
            e = -aps_log(1.0 - aps_u53(a.v[0], a.v[1]));
            u_choice = aps_u53(a.v[2], a.v[3]);
            u_event = aps_u53(b.v[0], b.v[1]);
            ds.b_dir = aps_u53(b.v[2], b.v[3]);
            ds.ev++;
        }
        double tau = (1.0 / R) * e;                                      /* rng.exponential(1/R), :358 */
        /* rng.choice(n, p=rates/R), :359-360 */
        double acc = 0.0;
        for (int i = 0; i < n; ++i) { acc = acc + w->rates[i] / R; w->cdf[i] = acc; }
        double last = w->cdf[n - 1];
        int sel = n - 1;
        for (int i = 0; i < n; ++i) { if (w->cdf[i] / last > u_choice) { sel = i; break; } }
        double v = u_event * w->rates[sel];                              /* :362 */
        double diff_thresh = w->r_diff[sel];
        double act_thresh = diff_thresh + w->r_act[sel];
        double bind_thresh = act_thresh + w->r_bind[sel];                /* :365-367 */
        double unbind_thresh = bind_thresh + w->r_unbind[sel];
        double exit_thresh = unbind_thresh + w->r_exit[sel];
        int old_pos = (int)w->pos[sel], new_pos = old_pos, kind;
        if (v < diff_thresh) {                                           /* :371-398 */
            double rl = w->r_left[sel], rr = w->r_right[sel];
            double u_dir;
            if (mode == 0) {
                int64_t at = (int64_t)(ds.d - draws_begin);
                if (ds.left < 4 || (B->spec_from >= 0 && at >= B->spec_from)) { status = APS_RUN_DRAWS_EXHAUSTED; break; }
                u_dir = ds.d[3]; ds.d += 1; ds.left -= 1;
            } else u_dir = ds.b_dir;
            const int per = (P->flags & APS_FLAG_PERIODIC) != 0;
            if (u_dir < rl / (rl + rr)) { new_pos = per ? (old_pos - 1 + L) % L : clipi((int64_t)old_pos - 1, 0, L - 1); kind = APS_EV_DIFF_LEFT; }
            else { new_pos = per ? (old_pos + 1) % L : clipi((int64_t)old_pos + 1, 0, L - 1); kind = APS_EV_DIFF_RIGHT; }
        } else if (v < act_thresh) {                                     /* :400-416 */
            const int per = (P->flags & APS_FLAG_PERIODIC) != 0;
            new_pos = per ? (old_pos + (w->sigma[sel] == 1)) % L : clipi((int64_t)old_pos + (w->sigma[sel] == 1), 0, L - 1);
            kind = APS_EV_ACTIVE;
        } else if (v < bind_thresh) kind = APS_EV_BIND;                  /* :418-419 */
        else if (v < unbind_thresh) kind = APS_EV_UNBIND;                /* :421-422 */
        else if (v < exit_thresh) kind = APS_EV_EXIT;                    /* :424-436 */
        else kind = APS_EV_FLIP;                                         /* :438-446 */
        if (kind == APS_EV_BIND) w->bound[sel] = 1;
        else if (kind == APS_EV_UNBIND) w->bound[sel] = 0;
        else if (kind == APS_EV_EXIT) {
            if (B->exit_t && n_exit < B->exit_cap) {
                B->exit_t[(size_t)rep * (size_t)B->exit_cap + n_exit] = t;               /* clock BEFORE this step's tau */
                B->exit_pos[(size_t)rep * (size_t)B->exit_cap + n_exit] = old_pos;
            }
            n_exit++;
            if (w->sigma[sel] == 1) w->cp[old_pos]--; else w->cm[old_pos]--;
            for (int i = sel; i + 1 < n; ++i) { w->pos[i] = w->pos[i + 1]; w->sigma[i] = w->sigma[i + 1]; w->bound[i] = w->bound[i + 1]; }
            n--;                                                                          /* np.delete, :434-436 */
        } else if (kind == APS_EV_FLIP) {
            if (w->sigma[sel] == 1) { w->sigma[sel] = -1; w->cp[old_pos]--; w->cm[old_pos]++; }
            else { w->sigma[sel] = 1; w->cm[old_pos]--; w->cp[old_pos]++; }
        } else {
            w->pos[sel] = new_pos;
            if (w->sigma[sel] == 1) { w->cp[old_pos]--; w->cp[new_pos]++; }
            else { w->cm[old_pos]--; w->cm[new_pos]++; }
        }
        if (mode == 0) { ds.d += 3; ds.left -= 3; }
        if (B->trace && n_events - ev_base < B->trace_cap) {
            int32_t* tr = B->trace + ((size_t)rep * (size_t)B->trace_cap + (size_t)(n_events - ev_base)) * 3;
            tr[0] = sel; tr[1] = kind; tr[2] = (kind <= APS_EV_ACTIVE) ? new_pos : (kind == APS_EV_EXIT ? old_pos : -1);
        }
        n_events++;
        t += tau;                                                        /* :514 */
        if (t > T) break;                                                /* :515 */
        while (obs_idx < M && B->times_obs[obs_idx] <= t) {              /* :517-536 */
            record_obs(P, B, rep, obs_idx, w, n, w->m);
            obs_idx++;
        }
        if (obs_idx >= M) break;                                         /* :538 */
    }
done:
    if (B->n_obs) B->n_obs[rep] = obs_idx;
    if (B->n_events) B->n_events[rep] = n_events;
    if (B->t_end) B->t_end[rep] = t;
    if (B->status) B->status[rep] = status;
    if (B->n_guard) B->n_guard[rep] = 0;
    if (B->draws_used) B->draws_used[rep] = (mode == 0 && draws_begin) ? (int64_t)(ds.d - draws_begin) : 0;
    if (B->pos_end) for (int i = 0; i < n; ++i) B->pos_end[(size_t)rep * B->n_max + i] = (int32_t)w->pos[i];
    if (B->sigma_end) for (int i = 0; i < n; ++i) B->sigma_end[(size_t)rep * B->n_max + i] = w->sigma[i];
    if (B->bound_end) for (int i = 0; i < n; ++i) B->bound_end[(size_t)rep * B->n_max + i] = w->bound[i];
    if (B->n_end) B->n_end[rep] = n;
    if (B->n_exit) B->n_exit[rep] = n_exit;
}

typedef struct targ { const aps_params* P; const aps_batch* B; int mode; int tid, nthreads; int rc; } targ;

static void* worker(void* vp) {
    targ* a = (targ*)vp;
    work w;
    if (work_alloc(&w, a->P->L, a->P->radius, a->B->n_max)) { a->rc = -1; return NULL; }
    for (int rep = a->tid; rep < a->B->n_replicas; rep += a->nthreads) run_one(a->P, a->B, rep, a->mode, &w);
    work_free(&w);
    a->rc = 0;
    return NULL;
}

/* mode 0 = replay, 1 = philox.  n_threads replicas are run concurrently (one pthread each). */
int aps_oracle_run(const aps_params* P, const aps_batch* B, int mode, int n_threads) {
    if (!P || !B || P->L <= 0 || P->K <= 0 || B->n_replicas < 0) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B->n_replicas) n_threads = B->n_replicas > 0 ? B->n_replicas : 1;
    pthread_t* th = malloc(sizeof(pthread_t) * (size_t)n_threads);
    targ* args = malloc(sizeof(targ) * (size_t)n_threads);
    for (int i = 0; i < n_threads; ++i) {
        args[i] = (targ){P, B, mode, i, n_threads, 0};
        if (i > 0) pthread_create(&th[i], NULL, worker, &args[i]);
    }
    worker(&args[0]);
    int rc = args[0].rc;
    for (int i = 1; i < n_threads; ++i) { pthread_join(th[i], NULL); if (args[i].rc) rc = args[i].rc; }
    free(th); free(args);
    return rc;
}

/* ---- initial conditions: same sampling algorithm as csrc/aps_init.cuh (host pointers) ---- */
int aps_oracle_init(const aps_init_args* a) {
    for (int rep = 0; rep < a->n_replicas; ++rep) {
        const uint32_t k0 = (uint32_t)a->seeds[rep], k1 = (uint32_t)(a->seeds[rep] >> 32);
        int32_t* gpos = a->pos0 + (size_t)rep * a->n_max;
        int8_t* gsig = a->sigma0 + (size_t)rep * a->n_max;
        if (a->mode == 1) {
            const int prof = a->profile_of ? a->profile_of[rep] : 0;
            const double* rp = a->rho0_plus + (size_t)prof * a->L;
            const double* rm = a->rho0_minus + (size_t)prof * a->L;
            int n = 0, overflow = 0;
            for (int x = 0; x < a->L; ++x) {
                uint64_t m;
                int c = sample_site((uint32_t)x, rp[x], rm[x], a->K, k0, k1, &m);
                for (int j = 0; j < c; ++j) {
                    if (n < a->n_max) { gpos[n] = x; gsig[n] = ((m >> j) & 1ULL) ? 1 : -1; } else overflow = 1;
                    ++n;
                }
            }
            a->n[rep] = overflow ? -1 : n;
        } else {
            const int N = a->N_of ? a->N_of[rep] : a->N_fixed;
            if (N > a->n_max || (long long)N > (long long)a->L * a->K) { a->n[rep] = -1; continue; }
            uint16_t* avail = malloc(2 * (size_t)a->L);
            uint8_t* fill = calloc((size_t)a->L, 1);
            for (int x = 0; x < a->L; ++x) avail[x] = (uint16_t)x;
            int navail = a->L;
            for (int i = 0; i < N; ++i) {
                aps_u32x4 r4 = aps_philox4x32_10((uint32_t)i, 0u, APS_RNG_INIT_POS, 0u, k0, k1);
                int j = (int)(aps_u53(r4.v[0], r4.v[1]) * (double)navail);
                if (j >= navail) j = navail - 1;
                int site = avail[j];
                gpos[i] = site;
                if (++fill[site] >= a->K) avail[j] = avail[--navail];
                aps_u32x4 s4 = aps_philox4x32_10((uint32_t)i, 0u, APS_RNG_INIT_SIGMA, 0u, k0, k1);
                gsig[i] = (s4.v[0] & 1u) ? 1 : -1;
            }
            free(avail); free(fill);
            a->n[rep] = N;
        }
    }
    return 0;
}

/* ---- K2: sublattice-parallel update rule (include/aps_k2_model.h), sequential restatement -------
 * Host pointers.  Processes segments one after another; active regions are disjoint, so the order
 * does not matter and the result must equal the CUDA kernel's bit for bit. */
static long long k2_reflect_idx(long long i, long long L) {
    long long per = 2 * L, m = i % per;
    if (m < 0) m += per;
    if (m >= L) m = per - 1 - m;
    return m;
}

int aps_oracle_k2_pass(const aps_k2_args* a) {
    const long long L = a->L;
    const int q = (int)(a->pass & 1ULL), r = a->radius;
    const uint8_t* in = a->in;
    uint8_t* out = a->out;
    memcpy(out, in, (size_t)L);
    const uint32_t k0 = (uint32_t)a->seed, k1 = (uint32_t)(a->seed >> 32);
    long long dsig = 0;
    uint32_t thr_glob[2] = {0, 0};
    if (r < 0) {
        const double m = (double)(*a->msum_in) / (double)a->n_particles;
        thr_glob[0] = aps_k2_flip_thr(a->rates.beta, +1, m, a->rates.inv_cmax, a->rates.t_active);
        thr_glob[1] = aps_k2_flip_thr(a->rates.beta, -1, m, a->rates.inv_cmax, a->rates.t_active);
    }
    for (long long seg = 0; seg * APS_K2_SEG < L; ++seg) {
        const long long abase = seg * APS_K2_SEG + q * APS_K2_HALF;
        if (abase + APS_K2_HALF > L) continue;
        const int dsig_on = (a->count_hi <= a->count_lo) || (abase >= a->count_lo && abase < a->count_hi);
        const uint64_t sg64 = (uint64_t)(a->global_offset / APS_K2_SEG) + (uint64_t)seg;
        const uint32_t c0 = (uint32_t)sg64, c1 = (uint32_t)a->pass;
        const uint32_t chi = (uint32_t)(sg64 >> 32) * 0x9E3779B9u + APS_RNG_SUBLATTICE;
        aps_u32x4 w4 = aps_philox4x32_10(c0, c1, 0u, chi, k0, k1);
        int ntr = 0;
        while (ntr < (int)a->rates.n_cdf && w4.v[0] >= a->rates.cdf32[ntr]) ++ntr;
        for (int t = 0; t < ntr; ++t) {
            uint32_t wa;                                                  /* one word per trial (aps_k2_model.h) */
            if (t < 3) wa = w4.v[t + 1];
            else {
                aps_u32x4 c4 = aps_philox4x32_10(c0, c1, aps_k2_trial_call(t), chi, k0, k1);
                wa = c4.v[aps_k2_trial_word(t)];
            }
            const long long x = abase + (long long)(wa >> 27);
            const uint32_t slot = wa << 5;
            const uint8_t v = out[x];
            if (v == APS_K2_EMPTY) continue;
            if (slot < a->rates.t_left) {
                if (x > 0 && out[x - 1] == APS_K2_EMPTY) { out[x - 1] = v; out[x] = APS_K2_EMPTY; }
            } else if (slot < a->rates.t_right || (slot < a->rates.t_active && v == APS_K2_PLUS)) {
                if (x < L - 1 && out[x + 1] == APS_K2_EMPTY) { out[x + 1] = v; out[x] = APS_K2_EMPTY; }
            } else if (slot >= a->rates.t_active) {
                const int sg = (v == APS_K2_PLUS) ? 1 : -1;
                uint32_t thr;
                if (r >= 0) {
                    int sw = 0, tw = 0;
                    for (int j = -r; j <= r; ++j) {
                        const int cv = in[k2_reflect_idx(x + j, L)];      /* frozen pre-pass state */
                        const int wj = a->w16[j < 0 ? -j : j];
                        sw += wj * ((cv == APS_K2_PLUS) - (cv == APS_K2_MINUS));
                        tw += wj * (cv != 0);
                    }
                    thr = a->flip_tab[(sg == 1 ? 0 : (2 * APS_K2_MQ + 1)) + aps_k2_mq_index(sw, tw)];
                } else thr = thr_glob[sg == 1 ? 0 : 1];
                if (slot - a->rates.t_active < thr) { out[x] = (v == APS_K2_PLUS) ? APS_K2_MINUS : APS_K2_PLUS; dsig -= dsig_on * 2 * sg; }
            }
        }
    }
    if (a->msum_out) *a->msum_out += dsig;
    return 0;
}

int aps_oracle_k2_run(aps_k2_args* a, int n_passes) {
    for (int p = 0; p < n_passes; ++p) {
        if (a->radius < 0) *a->msum_out = *a->msum_in;
        aps_oracle_k2_pass(a);
        const uint8_t* t = a->in; a->in = a->out; a->out = (uint8_t*)t;
        if (a->radius < 0) { const int64_t* m = a->msum_in; a->msum_in = a->msum_out; a->msum_out = (int64_t*)m; }
        a->pass += 1;
    }
    return 0;
}

int aps_oracle_k2_rates(double D, double lam, double beta, double dt, aps_k2_rates* out) {
    return aps_k2_make_rates(D, lam, beta, dt, out);
}

void aps_oracle_k2_flip_table(const aps_k2_rates* r, uint32_t* out) {
    for (int sgi = 0; sgi < 2; ++sgi)
        for (int i = 0; i <= 2 * APS_K2_MQ; ++i)
            out[sgi * (2 * APS_K2_MQ + 1) + i] = aps_k2_flip_thr(r->beta, sgi == 0 ? 1 : -1, (double)(i - APS_K2_MQ) / (double)APS_K2_MQ,
                                                                 r->inv_cmax, r->t_active);
}

void aps_oracle_k2_init(uint8_t* state, int64_t L, int64_t global_offset, uint64_t seed, double density, double frac_plus) {
    double vo = density * 4294967296.0, vp = frac_plus * 4294967296.0;
    uint32_t t_occ = (uint32_t)(vo >= 4294967295.0 ? 4294967295.0 : vo), t_plus = (uint32_t)(vp >= 4294967295.0 ? 4294967295.0 : vp);
    for (int64_t i = 0; i < L; i += 2) {
        uint64_t g = (uint64_t)(global_offset + i);
        aps_u32x4 w = aps_philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, APS_RNG_INIT_SITE, (uint32_t)seed, (uint32_t)(seed >> 32));
        state[i] = (w.v[0] < t_occ) ? ((w.v[1] < t_plus) ? APS_K2_PLUS : APS_K2_MINUS) : APS_K2_EMPTY;
        if (i + 1 < L) state[i + 1] = (w.v[2] < t_occ) ? ((w.v[3] < t_plus) ? APS_K2_PLUS : APS_K2_MINUS) : APS_K2_EMPTY;
    }
}

/* The native-mode variate stream of one replica as a replay log: for every event e = -log(1-u), u_choice,
 * u_event and, only where kinds[ev] is a diffusive hop, u_dir (what aps_run_replay_* consumes).
 * Returns the number of doubles written. */
int64_t aps_oracle_philox_log(uint64_t seed, int64_t ev0, int64_t n_events, const int32_t* kinds, double* out) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    int64_t w = 0;
    for (int64_t i = 0; i < n_events; ++i) {
        uint64_t ev = (uint64_t)(ev0 + i);
        aps_u32x4 a = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_A, 0, k0, k1);
        aps_u32x4 b = aps_philox4x32_10((uint32_t)ev, (uint32_t)(ev >> 32), APS_RNG_EVENT_B, 0, k0, k1);
        out[w++] = -aps_log(1.0 - aps_u53(a.v[0], a.v[1]));
        out[w++] = aps_u53(a.v[2], a.v[3]);
        out[w++] = aps_u53(b.v[0], b.v[1]);
        if (kinds[i] == APS_EV_DIFF_LEFT || kinds[i] == APS_EV_DIFF_RIGHT) out[w++] = aps_u53(b.v[2], b.v[3]);
    }
    return w;
}
