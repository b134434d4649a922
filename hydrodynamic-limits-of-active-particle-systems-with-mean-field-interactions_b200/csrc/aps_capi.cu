// aps_capi.cu — implementation of the C ABI in include/aps.h (host side of the kernels).
// No torch types, no CPU compute path: without an sm_100 device every compute entry fails.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aps_k1.cuh"
#include "aps_obs.cuh"
#include "aps_init.cuh"
#include "aps_k2.cuh"

#include "../../include/aps_pde.h"
namespace aps {
size_t pde_smem_bytes(int L, int bc, int n_tracers);
cudaError_t pde_launch(const aps_pde_args& a, cudaStream_t st);
}
namespace aps { cudaError_t launch_fast(const K1Args& a, bool philox, cudaStream_t st, int allow_static, int nt, int auto_threads, int* launched); }

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
double g_guard_scale = 1.0;
int g_k1_threads = 0;  // 0 = heuristic
int g_use_lut = 1;
int g_use_fast = 1;
int g_k2_ctas_per_sm = 6;
int g_reduce_threads = 128;
int g_reduce_impl = 0;          // 0: integer row sums (round 2); 1: the round-1 kernel (A/B reference of the tests)
int g_k2_stash_cap = 0;        // 0 = from the mean trial count (aps::k2_stash_cap)

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return APS_ERR_CUDA;
}
#define CU(call)                                         \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

int count_sm100() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int tree_nodes(int n) {  // nodes of numpy's pairwise-sum recursion for n elements
    if (n <= 128) return 1;
    int n2 = n / 2; n2 -= n2 % 8;
    return tree_nodes(n2) + tree_nodes(n - n2) + 1;
}
int max_tree_nodes(int n_max) {
    int best = 1;
    for (int n = 1; n <= n_max; ++n) { int t = tree_nodes(n); if (t > best) best = t; }
    return best;
}

int validate(const aps_params* p, const aps_batch* b, bool philox) {
    if (!p || !b) return fail(APS_ERR_INVALID, "null params/batch");
    if (p->L < 1 || p->L > 65535) return fail(APS_ERR_INVALID, "K1 needs 1 <= L <= 65535 (shared-memory resident lattice)");
    if (p->K < 1 || p->K > 63) return fail(APS_ERR_INVALID, "site capacity K must be in [1, 63]");
    if (b->n_replicas < 0 || b->n_max < 1 || b->n_max > 8192) return fail(APS_ERR_INVALID, "need 1 <= n_max <= 8192");
    if (b->M < 1) return fail(APS_ERR_INVALID, "need at least one observation time (M >= 1)");
    if (!b->times_obs || !b->beta || !b->n || !b->pos0 || !b->sigma0) return fail(APS_ERR_INVALID, "missing required input pointer");
    if (p->radius >= 0 && !b->weights) return fail(APS_ERR_INVALID, "weights required when radius >= 0");
    if ((p->flags & APS_FLAG_PERIODIC) && p->radius >= 0 && 2 * (int64_t)p->radius + 1 > p->L)
        return fail(APS_ERR_INVALID, "periodic field needs 2*radius+1 <= L (truncate the ring kernel)");
    if (philox && !b->seeds) return fail(APS_ERR_INVALID, "seeds required in native (Philox) mode");
    if (!philox && (!b->draws || !b->draw_off)) return fail(APS_ERR_INVALID, "draws/draw_off required in replay mode");
    if (b->flip_tab && b->flip_G < 1) return fail(APS_ERR_INVALID, "flip_tab needs flip_G >= 1 (grid m_k = -1 + 2k/flip_G)");
    return APS_OK;
}

template <int NT>
int launch_nt(const aps::K1Args& a, bool philox, size_t smem, cudaStream_t st) {
    auto kr = aps::k1_kernel<NT, false>;
    auto kp = aps::k1_kernel<NT, true>;
    if (philox) {
        CU(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kp<<<a.b.n_replicas, NT, smem, st>>>(a);
    } else {
        CU(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kr<<<a.b.n_replicas, NT, smem, st>>>(a);
    }
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

// APS_K1_THREADS (A/B knob of include/aps.h): read once at load time, never on the launch path
const int g_env_k1_threads = [] {
    const char* e = getenv("APS_K1_THREADS");
    const int v = e ? atoi(e) : 0;
    return (v == 32 || v == 64 || v == 128 || v == 256) ? v : 0;
}();

int pick_threads(int n_max) {
    if (g_k1_threads) return g_k1_threads;
    if (g_env_k1_threads) return g_env_k1_threads;
    return n_max > 2048 ? 128 : 64;
}

int run_device(const aps_params* p, const aps_batch* b, void* stream, bool philox) {
    int rc = validate(p, b, philox);
    if (rc) return rc;
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (b->n_replicas == 0) return APS_OK;
    aps::K1Args a;
    a.p = *p; a.b = *b;
    if (!a.b.trace) a.b.trace_cap = 0;        // internal copy: lets the kernels test the capacity alone
    a.guard_scale = g_guard_scale;
    a.max_nodes = max_tree_nodes(b->n_max);
    a.pad = p->radius > 0 ? p->radius : 0;
    const int nt = pick_threads(b->n_max);
    a.bcode = 2 * p->K + 1;
    a.use_lut = (g_use_lut && p->radius >= 0 && aps::k1_lut_bytes(p->K, p->radius) <= aps::kLutMaxBytes) ? 1 : 0;
    const size_t smem = aps::k1_smem_bytes(p->L, b->n_max, p->radius, a.max_nodes, nt / 32, a.use_lut, p->K);
    if (smem > 227 * 1024) return fail(APS_ERR_CAPACITY, "replica does not fit in 227 KB of shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    a.only_retry = 0; a.n_lo = -1;
    a.wt_valid = 0; a.n_hi = 0x7fffffff;
    for (int j = 0; j < 84; ++j) a.wt[j] = 0.0;
    if (b->weights_host && p->radius >= 0 && p->radius <= 83) {      // taps w[0..radius] (outermost first) as kernel parameters
        for (int j = 0; j <= p->radius; ++j) a.wt[j] = b->weights_host[j];
        a.wt_valid = 1;
    }
    // specialised kernel for K = 1 with a local field (every shipped sweep configuration)
    const bool fast_ok = g_use_fast && p->K == 1 && p->radius >= 0 && p->radius < p->L && !(p->flags & (APS_FLAG_CROWDING | APS_FLAG_PERIODIC)) &&
                         !b->m_field_in && !b->anchor_mask && !b->flip_tab && b->n_max <= 1024 && b->status != nullptr;
    if (fast_ok) {
        int launched = 0;
        // single-warp CTAs (no block barriers) win for narrow update windows; wide windows (r > 30) use two warps
        const bool forced = g_k1_threads || g_env_k1_threads;
        const int fnt = forced ? nt : (p->radius <= 30 ? 32 : 64);
        CU(aps::launch_fast(a, philox, st, g_use_fast == 1, fnt, forced ? 0 : 1, &launched));
        if (launched) { g_launches.fetch_add(launched); a.only_retry = 1; }   // last launch: replicas violating K = 1 only
    }
    switch (nt) {
        case 32: return launch_nt<32>(a, philox, smem, st);
        case 64: return launch_nt<64>(a, philox, smem, st);
        case 256: return launch_nt<256>(a, philox, smem, st);
        default: return launch_nt<128>(a, philox, smem, st);
    }
}

// ---- host-buffer staging -------------------------------------------------------------------
struct DevBuf {
    void* d = nullptr;
    ~DevBuf() { if (d) cudaFree(d); }
};

struct Stager {
    std::vector<DevBuf*> bufs;
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;
    cudaStream_t st = nullptr;
    ~Stager() { for (auto* b : bufs) delete b; }
    // copy a host input to the device (returns nullptr for nullptr)
    int in(const void* h, size_t bytes, const void** dptr) {
        *dptr = nullptr;
        if (!h || bytes == 0) return APS_OK;
        auto* b = new DevBuf(); bufs.push_back(b);
        CU(cudaMalloc(&b->d, bytes));
        CU(cudaMemcpyAsync(b->d, h, bytes, cudaMemcpyHostToDevice, st));
        *dptr = b->d;
        return APS_OK;
    }
    // allocate a device output mirrored back to `h` at the end
    int out(void* h, size_t bytes, void** dptr, bool copy_in_first = true) {
        *dptr = nullptr;
        if (!h || bytes == 0) return APS_OK;
        auto* b = new DevBuf(); bufs.push_back(b);
        CU(cudaMalloc(&b->d, bytes));
        if (copy_in_first) CU(cudaMemcpyAsync(b->d, h, bytes, cudaMemcpyHostToDevice, st));
        outs.push_back({h, b->d, bytes});
        *dptr = b->d;
        return APS_OK;
    }
    int finish() {
        for (auto& o : outs) CU(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return APS_OK;
    }
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

int run_host(const aps_params* p, const aps_batch* hb, bool philox) {
    int rc = validate(p, hb, philox);
    if (rc) return rc;
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    const size_t R = (size_t)hb->n_replicas, NM = (size_t)hb->n_max, M = (size_t)hb->M, L = (size_t)p->L;
    if (R == 0) return APS_OK;
    Stager s;
    aps_batch d = *hb;
    TRY(s.in(hb->times_obs, M * 8, (const void**)&d.times_obs));
    TRY(s.in(hb->weights, p->radius >= 0 ? (size_t)(2 * p->radius + 1) * 8 : 0, (const void**)&d.weights));
    d.weights_host = hb->weights;                  // host-buffer entry point: the taps are host memory already
    TRY(s.in(hb->beta, R * 8, (const void**)&d.beta));
    TRY(s.in(hb->n, R * 4, (const void**)&d.n));
    TRY(s.in(hb->pos0, R * NM * 4, (const void**)&d.pos0));
    TRY(s.in(hb->sigma0, R * NM, (const void**)&d.sigma0));
    if (!philox) {
        TRY(s.in(hb->draw_off, (R + 1) * 8, (const void**)&d.draw_off));
        TRY(s.in(hb->draws, (size_t)hb->draw_off[R] * 8, (const void**)&d.draws));
    } else {
        TRY(s.in(hb->seeds, R * 8, (const void**)&d.seeds));
    }
    TRY(s.in(hb->t_start, R * 8, (const void**)&d.t_start));
    TRY(s.in(hb->obs_start, R * 4, (const void**)&d.obs_start));
    TRY(s.in(hb->ev_start, R * 8, (const void**)&d.ev_start));
    // observation rows not reached keep whatever the caller put there (the reference leaves zeros)
    TRY(s.out((hb->record & APS_REC_COUNTS) ? hb->obs_cp : nullptr, R * M * L, (void**)&d.obs_cp));
    TRY(s.out((hb->record & APS_REC_COUNTS) ? hb->obs_cm : nullptr, R * M * L, (void**)&d.obs_cm));
    TRY(s.out((hb->record & APS_REC_POS) ? hb->obs_pos : nullptr, R * M * NM * 4, (void**)&d.obs_pos));
    TRY(s.out(hb->obs_sigma_sum, R * M * 4, (void**)&d.obs_sigma_sum));
    TRY(s.out((hb->record & APS_REC_MLOCAL) ? hb->obs_m_local : nullptr, R * M * L * 8, (void**)&d.obs_m_local));
    TRY(s.out(hb->n_obs, R * 4, (void**)&d.n_obs, false));
    TRY(s.out(hb->n_events, R * 8, (void**)&d.n_events, false));
    TRY(s.out(hb->t_end, R * 8, (void**)&d.t_end, false));
    TRY(s.out(hb->status, R * 4, (void**)&d.status, false));
    TRY(s.out(hb->n_guard, R * 8, (void**)&d.n_guard, false));
    TRY(s.out(hb->draws_used, R * 8, (void**)&d.draws_used, false));
    TRY(s.out(hb->pos_end, R * NM * 4, (void**)&d.pos_end));
    TRY(s.out(hb->sigma_end, R * NM, (void**)&d.sigma_end));
    TRY(s.out(hb->trace_cap > 0 ? hb->trace : nullptr, R * (size_t)hb->trace_cap * 12, (void**)&d.trace));
    if (hb->trace_cap <= 0) d.trace = nullptr;
    TRY(s.in(hb->m_field_in, R * L * 8, (const void**)&d.m_field_in));
    TRY(s.in(hb->anchor_mask, L, (const void**)&d.anchor_mask));
    TRY(s.in(hb->bound0, R * NM, (const void**)&d.bound0));
    TRY(s.in(hb->flip_tab, hb->flip_tab ? 2 * (size_t)(hb->flip_G + 1) * 8 : 0, (const void**)&d.flip_tab));
    TRY(s.out(hb->n_end, R * 4, (void**)&d.n_end, false));
    TRY(s.out(hb->bound_end, R * NM, (void**)&d.bound_end));
    TRY(s.out(hb->obs_n, R * M * 4, (void**)&d.obs_n));
    TRY(s.out(hb->obs_bound, R * M * NM, (void**)&d.obs_bound));
    const size_t EC = hb->exit_cap > 0 ? (size_t)hb->exit_cap : 0;
    TRY(s.out(EC ? hb->exit_t : nullptr, R * EC * 8, (void**)&d.exit_t));
    TRY(s.out(EC ? hb->exit_pos : nullptr, R * EC * 4, (void**)&d.exit_pos));
    TRY(s.out(hb->n_exit, R * 4, (void**)&d.n_exit));
    if (!EC) { d.exit_t = nullptr; d.exit_pos = nullptr; }
    TRY(run_device(p, &d, nullptr, philox));
    return s.finish();
}

}  // namespace

extern "C" {

int aps_abi_version(void) { return APS_ABI_VERSION; }
const char* aps_last_error(void) { return g_err.c_str(); }
int aps_device_count(void) { return count_sm100(); }
int aps_set_device(int device) {
    CU(cudaSetDevice(device));
    return APS_OK;
}
int64_t aps_launch_count(void) { return g_launches.load(); }

int64_t aps_replica_smem_bytes(const aps_params* p, int32_t n_max) {
    if (!p || n_max < 1 || n_max > 8192) return -1;
    int nt = pick_threads(n_max);
    int use_lut = (g_use_lut && p->radius >= 0 && aps::k1_lut_bytes(p->K, p->radius) <= aps::kLutMaxBytes) ? 1 : 0;
    size_t b = aps::k1_smem_bytes(p->L, n_max, p->radius, max_tree_nodes(n_max), nt / 32, use_lut, p->K);
    return b > 227 * 1024 ? -1 : (int64_t)b;
}

int aps_run_replay_device(const aps_params* p, const aps_batch* b, void* stream) { return run_device(p, b, stream, false); }
int aps_run_philox_device(const aps_params* p, const aps_batch* b, void* stream) { return run_device(p, b, stream, true); }
int aps_run_replay_host(const aps_params* p, const aps_batch* b) { return run_host(p, b, false); }
int aps_run_philox_host(const aps_params* p, const aps_batch* b) { return run_host(p, b, true); }

int aps_init_particles_device(const aps_init_args* a, void* stream) {
    if (!a || a->L < 1 || a->L > 65535 || a->K < 1 || a->K > 63 || a->n_max < 1 || !a->seeds || !a->pos0 || !a->sigma0 || !a->n)
        return fail(APS_ERR_INVALID, "aps_init_particles: bad argument");
    if (a->mode == 1 && (!a->rho0_plus || !a->rho0_minus)) return fail(APS_ERR_INVALID, "poisson init needs rho0_plus/rho0_minus");
    if (a->mode != 0 && a->mode != 1) return fail(APS_ERR_INVALID, "mode must be 0 (fixed) or 1 (poisson)");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (a->n_replicas == 0) return APS_OK;
    size_t smem = a->mode == 1 ? (size_t)a->L * 12 : (size_t)a->L * 3 + 8;
    if (smem > 200 * 1024) return fail(APS_ERR_CAPACITY, "L too large for the init kernel");
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(aps::init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aps::init_kernel<<<a->n_replicas, 128, smem, (cudaStream_t)stream>>>(*a);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_m_field_host(const aps_params* p, const double* weights, const int32_t* cp, const int32_t* cm, double* out) {
    if (!p || !cp || !cm || !out || p->L < 1 || p->L > 65535 || (p->radius >= 0 && !weights))
        return fail(APS_ERR_INVALID, "aps_m_field_host: bad argument");
    if ((p->flags & APS_FLAG_PERIODIC) && p->radius >= 0 && 2 * (int64_t)p->radius + 1 > p->L)
        return fail(APS_ERR_INVALID, "periodic field needs 2*radius+1 <= L (truncate the ring kernel)");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    const int pad = p->radius > 0 ? p->radius : 0;
    const size_t L = (size_t)p->L, smem = (size_t)(pad + 1) * 8 + (L + 2 * (size_t)pad) * 2 + 16;
    if (smem > 227 * 1024) return fail(APS_ERR_CAPACITY, "lattice + halo do not fit in shared memory");
    Stager s;
    const void *dw = nullptr, *dcp = nullptr, *dcm = nullptr; void* dout = nullptr;
    TRY(s.in(weights, p->radius >= 0 ? (size_t)(2 * p->radius + 1) * 8 : 0, &dw));
    TRY(s.in(cp, L * 4, &dcp));
    TRY(s.in(cm, L * 4, &dcm));
    TRY(s.out(out, L * 8, &dout, false));
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(aps::field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aps::field_kernel<<<1, 256, smem>>>(*p, (const double*)dw, (const int32_t*)dcp, (const int32_t*)dcm, (double*)dout);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return s.finish();
}

int aps_expand_obs_device(const aps_expand_args* a, void* stream) {
    if (!a || a->L < 1 || a->M < 1 || !a->n || !a->n_obs || !a->obs_cp || !a->obs_cm)
        return fail(APS_ERR_INVALID, "aps_expand_obs: missing argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (a->n_replicas == 0) return APS_OK;
    size_t smem = a->var ? (size_t)a->L * 8 : 0;
    if (smem > 200 * 1024) return fail(APS_ERR_CAPACITY, "L too large for the variance pass");
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(aps::expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(a->M, a->n_replicas);
    aps::expand_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(*a);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_reduce_runs_device(const aps_reduce_args* a, void* stream) {
    if (!a || a->L < 1 || a->M < 1 || !a->n || !a->n_obs || !a->obs_cp || !a->obs_cm || !a->obs_sigma_sum || !a->out ||
        !a->times_obs)
        return fail(APS_ERR_INVALID, "aps_reduce_runs: missing argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (a->n_replicas == 0) return APS_OK;
    size_t smem = ((size_t)3 * a->M + 32 + 128) * 8;     // row scalars, reduction scratch, density table
    if (smem > 200 * 1024) return fail(APS_ERR_CAPACITY, "too many observation rows for the reducer");
    auto kern = g_reduce_impl == 1 ? aps::reduce_kernel_v1 : aps::reduce_kernel;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a->n_replicas, g_reduce_threads, smem, (cudaStream_t)stream>>>(*a);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_profile_sums_device(const aps_profile_args* a, void* stream) {
    if (!a || a->L < 1 || a->M < 1 || !a->n || !a->n_obs || !a->obs_cp || !a->obs_cm || !a->prof || a->row_hi <= a->row_lo ||
        a->row_lo < 0 || a->row_hi > a->M || (a->reps_per_point < 1 && !a->point_start) || (a->point_start && !a->point_reps))
        return fail(APS_ERR_INVALID, "aps_profile_sums: bad argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (a->n_points == 0) return APS_OK;
    if (a->point_start && a->scratch) {
        // replica lists: (1) one CTA row per replica (all replicas in parallel) into the scratch rows, (2) per-point gather-sum
        const int R = a->n_replicas;
        if (R < 1) return fail(APS_ERR_INVALID, "aps_profile_sums: n_replicas needed with point lists and scratch");
        aps_profile_args one = *a;
        one.n_points = R; one.reps_per_point = 1; one.point_start = nullptr; one.point_reps = nullptr; one.prof = a->scratch;
        if (a->L % 4 == 0) {
            dim3 g1((a->L / 4 + 63) / 64, R);
            aps::profile_kernel_w4<<<g1, 64, 0, (cudaStream_t)stream>>>(one);
        } else {
            dim3 g1((a->L + 127) / 128, R);
            aps::profile_kernel<<<g1, 128, 0, (cudaStream_t)stream>>>(one);
        }
        CU(cudaGetLastError());
        dim3 g2((a->L + 127) / 128, a->n_points);
        aps::profile_gather_kernel<<<g2, 128, 0, (cudaStream_t)stream>>>(a->scratch, a->point_start, a->point_reps, a->prof, a->L);
        CU(cudaGetLastError());
        g_launches.fetch_add(2);
        return APS_OK;
    }
    if (a->L % 4 == 0) {
        dim3 grid((a->L / 4 + 63) / 64, a->n_points);
        aps::profile_kernel_w4<<<grid, 64, 0, (cudaStream_t)stream>>>(*a);
    } else {
        dim3 grid((a->L + 127) / 128, a->n_points);
        aps::profile_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*a);
    }
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_m_histogram_device(const aps_hist_args* a, void* stream) {
    if (!a || a->M < 1 || a->n_bins < 1 || a->n_points < 1 || !a->n || !a->n_obs || !a->obs_sigma_sum || !a->hist ||
        a->row_lo < 0 || a->row_hi > a->M || a->row_hi <= a->row_lo || !(a->hi > a->lo))
        return fail(APS_ERR_INVALID, "aps_m_histogram: bad argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (!a->accumulate) CU(cudaMemsetAsync(a->hist, 0, (size_t)a->n_points * (size_t)a->n_bins * 8, (cudaStream_t)stream));
    if (a->n_replicas == 0) return APS_OK;
    aps::hist_kernel<<<(a->n_replicas + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_k2_rates_init(double D, double lam, double beta, double dt, aps_k2_rates* out) {
    if (!out || !(D >= 0) || !(lam >= 0) || !(dt > 0)) return fail(APS_ERR_INVALID, "aps_k2_rates_init: bad argument");
    if (aps_k2_make_rates(D, lam, beta, dt, out)) return fail(APS_ERR_INVALID, "B*32*dt must be in (0, 24]: reduce dt");
    return APS_OK;
}

static size_t k2_smem(int radius, int cap, int stages = 0) {
    const size_t WB = aps::kK2Tile + 32;
    const size_t R16 = radius >= 0 ? (size_t)((radius + 15) & ~15) : 0;
    const size_t stride = WB + 2 * R16;              // local field: window + halo in one buffer per stage
    if (stages <= 0) stages = radius >= 0 ? aps::kK2StagesLocal : aps::kK2StagesGlobal;
    return 128 + (size_t)stages * stride + (radius >= 0 ? aps::k2_scratch_bytes(radius, cap) + 16 : 0);
}

// Launch plan of a K2 pass: ring depth, shared memory and grid.  Long tile walks (many tiles per CTA) use the deep ring at
// g_k2_ctas_per_sm CTAs per SM.  SHORT slabs — a 2^26-site lattice cut over 8 GPUs leaves 1026 tiles per rank against 888
// persistent CTAs, i.e. two rounds of which the second is 15 % full — take a 2-deep ring instead when that lets enough CTAs be
// resident (co-resident for the grid barrier) to walk the slab in ONE round.
struct K2Plan { int stages; size_t smem; int grid; };
const bool g_env_k2_no_short_plan = getenv("APS_K2_NO_SHORT_PLAN") != nullptr;   // A/B knob, read once at load
const int g_env_k2_stages = [] { const char* e = getenv("APS_K2_STAGES"); return e ? atoi(e) : 0; }();   // A/B knob: ring depth of the first plan
static int k2_plan(const void* fn, int radius, int cap, int ntiles, int n_sm, K2Plan* out) {
    // one-entry cache per thread: a run is thousands of launches with the same plan (the occupancy queries are host-side work)
    struct Key { const void* fn; int radius, cap, ntiles, per_sm_cap; K2Plan plan; };
    thread_local Key last{nullptr, 0, 0, 0, 0, {}};
    if (last.fn == fn && last.radius == radius && last.cap == cap && last.ntiles == ntiles && last.per_sm_cap == g_k2_ctas_per_sm) { *out = last.plan; return 0; }
    int best_rounds = 1 << 30;
    for (int pass = 0; pass < (g_env_k2_no_short_plan ? 1 : 2); ++pass) {
        const int stages = pass == 0 ? (g_env_k2_stages >= 2 ? g_env_k2_stages : (radius >= 0 ? aps::kK2StagesLocal : aps::kK2StagesGlobal)) : 2;
        const size_t smem = k2_smem(radius, cap, stages);
        if (smem > 227 * 1024) continue;
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, aps::kK2Threads, smem) != cudaSuccess) { cudaGetLastError(); continue; }
        const int cap_sm = pass == 0 ? g_k2_ctas_per_sm : 2 * g_k2_ctas_per_sm;
        if (per_sm > cap_sm) per_sm = cap_sm;
        if (per_sm < 1) continue;
        int grid = n_sm * per_sm;
        if (grid > ntiles) grid = ntiles;
        const int rounds = (ntiles + grid - 1) / grid;
        // the 2-deep ring only pays when it makes the walk a SINGLE round (measured: 1024 tiles 14.3 -> 12.9 us per pass; with two
        // rounds of 1776 CTAs the grid barrier of the persistent kernel costs more than the saved round: 20.4 -> 29.0 us at 2048 tiles)
        if (rounds < best_rounds && (pass == 0 || rounds == 1)) { best_rounds = rounds; out->stages = stages; out->smem = smem; out->grid = grid; }
    }
    if (best_rounds == (1 << 30)) return -1;
    last = Key{fn, radius, cap, ntiles, g_k2_ctas_per_sm, *out};
    return 0;
}

int aps_k2_flip_table(const aps_k2_rates* r, uint32_t* out) {
    if (!r || !out) return fail(APS_ERR_INVALID, "aps_k2_flip_table: null argument");
    for (int sgi = 0; sgi < 2; ++sgi)
        for (int i = 0; i <= 2 * APS_K2_MQ; ++i)
            out[sgi * (2 * APS_K2_MQ + 1) + i] = aps_k2_flip_thr(r->beta, sgi == 0 ? 1 : -1, (double)(i - APS_K2_MQ) / (double)APS_K2_MQ,
                                                                 r->inv_cmax, r->t_active);
    return APS_OK;
}

int aps_k2_pass_device(const aps_k2_args* a, void* stream) {
    if (!a || a->L < aps::kK2Tile || a->L % aps::kK2Tile || a->global_offset % aps::kK2Tile || !a->in || !a->out || a->in == a->out)
        return fail(APS_ERR_INVALID, "aps_k2_pass: L and global_offset must be multiples of 8192, in != out");
    if (a->radius > 1024) return fail(APS_ERR_INVALID, "aps_k2_pass: radius too large");
    if (a->radius >= 0 && (!a->w16 || !a->flip_tab)) return fail(APS_ERR_INVALID, "aps_k2_pass: local field needs w16 taps and flip_tab");
    if (a->radius < 0 && (!a->msum_in || a->n_particles < 1)) return fail(APS_ERR_INVALID, "aps_k2_pass: global field needs msum_in and n_particles");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    const int cap = g_k2_stash_cap > 0 ? g_k2_stash_cap : aps::k2_stash_cap(a->rates.mu);
    if (k2_smem(a->radius, cap) > 227 * 1024) return fail(APS_ERR_CAPACITY, "aps_k2_pass: radius too large for the shared-memory ring");
    const int ntiles = (int)(a->L / aps::kK2Tile);
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; CU(cudaGetDevice(&dev)); CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
    const void* fn = a->radius >= 0 ? (const void*)aps::k2_pass_kernel<true, false> : (const void*)aps::k2_pass_kernel<false, false>;
    K2Plan plan{};
    if (k2_plan(fn, a->radius, cap, ntiles, n_sm, &plan)) return fail(APS_ERR_CAPACITY, "aps_k2_pass: kernel does not fit on an SM");
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    const aps::K2Multi none{};
    if (a->radius >= 0) aps::k2_pass_kernel<true, false><<<plan.grid, aps::kK2Threads, plan.smem, (cudaStream_t)stream>>>(*a, cap, none, plan.stages);
    else aps::k2_pass_kernel<false, false><<<plan.grid, aps::kK2Threads, plan.smem, (cudaStream_t)stream>>>(*a, cap, none, plan.stages);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

// ---- many passes per launch (grid barriers) + slab exchange over peer memory ----
int aps_k2_peer_region_bytes(void) { return (int)sizeof(aps::K2PeerRegion); }

int aps_k2_peer_alloc(void** region, void* ipc_handle_out) {
    if (!region || !ipc_handle_out) return fail(APS_ERR_INVALID, "aps_k2_peer_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "aps.h promises a 64-byte handle");
    void* p = nullptr;
    CU(cudaMalloc(&p, sizeof(aps::K2PeerRegion)));
    CU(cudaMemset(p, 0, sizeof(aps::K2PeerRegion)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *region = p;
    return APS_OK;
}
int aps_k2_peer_open(const void* ipc_handle, void** region) {
    if (!region || !ipc_handle) return fail(APS_ERR_INVALID, "aps_k2_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(region, h, cudaIpcMemLazyEnablePeerAccess));
    return APS_OK;
}
int aps_k2_peer_close(void* region) { if (region) CU(cudaIpcCloseMemHandle(region)); return APS_OK; }
int aps_k2_peer_free(void* region) { if (region) CU(cudaFree(region)); return APS_OK; }

int aps_k2_run_persistent_device(aps_k2_args* a, const aps_k2_multi* hm, void* stream) {
    if (!a || !hm || hm->n_passes < 0 || !hm->buf0 || !hm->buf1 || hm->buf0 == hm->buf1 || !hm->sync)
        return fail(APS_ERR_INVALID, "aps_k2_run_persistent: missing buffers / sync scratch");
    if (a->L < aps::kK2Tile || a->L % aps::kK2Tile || a->global_offset % aps::kK2Tile)
        return fail(APS_ERR_INVALID, "aps_k2_run_persistent: L and global_offset must be multiples of 8192");
    if (a->radius > 1024) return fail(APS_ERR_INVALID, "aps_k2_run_persistent: radius too large");
    if (a->radius >= 0 && (!a->w16 || !a->flip_tab)) return fail(APS_ERR_INVALID, "aps_k2_run_persistent: local field needs w16 taps and flip_tab");
    if (a->radius < 0 && a->n_particles < 1) return fail(APS_ERR_INVALID, "aps_k2_run_persistent: global field needs n_particles");
    if (hm->world < 1 || hm->world > aps::kK2MaxRanks || hm->rank < 0 || hm->rank >= hm->world)
        return fail(APS_ERR_INVALID, "aps_k2_run_persistent: need 1 <= world <= 8 and 0 <= rank < world");
    if (hm->world > 1) {
        if (hm->ghost < 16 || hm->ghost > aps::kK2GhostMax || hm->ghost % 16 || hm->refresh_every < 1)
            return fail(APS_ERR_INVALID, "aps_k2_run_persistent: ghost must be a multiple of 16 in [16, 65536], refresh_every >= 1");
        for (int q = 0; q < hm->world; ++q) if (!hm->peer[q]) return fail(APS_ERR_INVALID, "aps_k2_run_persistent: missing peer region");
    }
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    if (hm->n_passes == 0) return APS_OK;
    const int cap = g_k2_stash_cap > 0 ? g_k2_stash_cap : aps::k2_stash_cap(a->rates.mu);
    if (k2_smem(a->radius, cap) > 227 * 1024) return fail(APS_ERR_CAPACITY, "aps_k2_run_persistent: radius too large for the shared-memory ring");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, n_sm = 0, coop = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return fail(APS_ERR_CUDA, "device does not support cooperative launches");
    const void* fn = a->radius >= 0 ? (const void*)aps::k2_pass_kernel<true, true> : (const void*)aps::k2_pass_kernel<false, true>;
    const int ntiles = (int)(a->L / aps::kK2Tile);
    K2Plan plan{};                                  // all CTAs co-resident (grid barrier): grid <= SMs x resident CTAs per SM
    if (k2_plan(fn, a->radius, cap, ntiles, n_sm, &plan)) return fail(APS_ERR_CAPACITY, "aps_k2_run_persistent: kernel does not fit on an SM");
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    const int grid = plan.grid;
    const size_t smem = plan.smem;
    aps::K2Multi m{};
    m.n_passes = hm->n_passes; m.world = hm->world; m.rank = hm->rank; m.refresh_every = hm->refresh_every > 0 ? hm->refresh_every : 1;
    m.ghost = hm->ghost; m.own_lo = hm->own_lo; m.own_hi = hm->own_hi;
    m.buf[0] = hm->buf0; m.buf[1] = hm->buf1;
    long long* sync = (long long*)hm->sync;         // [0] grid barrier, [1] error flag, [2] acc, [3] msum0, [4] msum_cur
    m.gbar = (unsigned*)(sync + 0); m.err = (int32_t*)(sync + 1); m.acc = sync + 2; m.msum0 = sync + 3; m.msum_cur = sync + 4;
    for (int q = 0; q < aps::kK2MaxRanks; ++q) m.peer[q] = q < hm->world ? (aps::K2PeerRegion*)hm->peer[q] : nullptr;
    CU(cudaMemsetAsync(sync, 0, 8, st));            // barrier counter restarts with every launch
    aps_k2_args ka = *a;
    if (hm->world > 1 && a->radius < 0) { ka.count_lo = hm->own_lo; ka.count_hi = hm->own_hi; }   // own flips only
    int cap_arg = cap, stages_arg = plan.stages;
    void* params[] = {(void*)&ka, (void*)&cap_arg, (void*)&m, (void*)&stages_arg};
    CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(aps::kK2Threads), params, smem, st));
    g_launches.fetch_add(1);
    a->pass += (uint64_t)hm->n_passes;
    return APS_OK;
}

int aps_k2_run_device(aps_k2_args* a, int n_passes, void* stream) {
    if (!a || n_passes < 0) return fail(APS_ERR_INVALID, "aps_k2_run: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int p = 0; p < n_passes; ++p) {
        if (a->radius < 0) {
            if (!a->msum_out) return fail(APS_ERR_INVALID, "aps_k2_run: global field needs msum_in/msum_out");
            CU(cudaMemcpyAsync(a->msum_out, a->msum_in, 8, cudaMemcpyDeviceToDevice, st));
        }
        TRY(aps_k2_pass_device(a, stream));
        const uint8_t* t = a->in; a->in = a->out; a->out = const_cast<uint8_t*>(t);
        if (a->radius < 0) { const int64_t* m = a->msum_in; a->msum_in = a->msum_out; a->msum_out = const_cast<int64_t*>(m); }
        a->pass += 1;
    }
    return APS_OK;
}

int aps_k2_init_device(uint8_t* state, int64_t L, int64_t global_offset, uint64_t seed, double density, double frac_plus,
                       void* stream) {
    if (!state || L < 1 || !(density >= 0 && density <= 1) || !(frac_plus >= 0 && frac_plus <= 1))
        return fail(APS_ERR_INVALID, "aps_k2_init: bad argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    auto thr = [](double p) { double v = p * 4294967296.0; return (uint32_t)(v >= 4294967295.0 ? 4294967295.0 : v); };
    const long long threads = (L + 1) / 2;
    aps::k2_init_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(state, L, global_offset, seed, thr(density), thr(frac_plus));
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

int aps_k2_profile_device(const uint8_t* state, int64_t L, int64_t global_offset, int64_t L_global, int32_t nbins,
                          uint64_t* cnt_plus, uint64_t* cnt_minus, void* stream) {
    if (!state || L < 1 || nbins < 1 || !cnt_plus || !cnt_minus || L_global < L) return fail(APS_ERR_INVALID, "aps_k2_profile: bad argument");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    aps::k2_profile_kernel<<<(unsigned)((L + 4095) / 4096), 256, 0, (cudaStream_t)stream>>>(
        state, L, global_offset, L_global, nbins, (unsigned long long*)cnt_plus, (unsigned long long*)cnt_minus);
    CU(cudaGetLastError());
    g_launches.fetch_add(1);
    return APS_OK;
}

// ---- batched IMEX PDE stepper (include/aps_pde.h) ----
int64_t aps_pde_smem_bytes(int32_t L, int32_t bc) { return (int64_t)aps::pde_smem_bytes(L, bc, 0); }

int aps_pde_solve_device(const aps_pde_args* a, void* stream) {
    if (!a || a->n_runs < 1 || a->nsteps < 0 || !(a->dt > 0) || !(a->dx > 0) || !(a->xlim > 0) || a->snapshot_interval < 1)
        return fail(APS_ERR_INVALID, "aps_pde_solve: bad scalar argument");
    if (a->bc != APS_PDE_BC_PERIODIC && a->bc != APS_PDE_BC_NEUMANN) return fail(APS_ERR_INVALID, "aps_pde_solve: bc");
    if (a->model != APS_PDE_MODEL_BIDIRECTIONAL && a->model != APS_PDE_MODEL_ANCHORED_MINUS) return fail(APS_ERR_INVALID, "aps_pde_solve: model");
    if (a->field != APS_PDE_FIELD_POINTWISE && a->field != APS_PDE_FIELD_KERNEL) return fail(APS_ERR_INVALID, "aps_pde_solve: field");
    if (!a->beta || !a->lam || !a->gamma || !a->rho_p || !a->rho_m || !a->m_series || !a->var_series)
        return fail(APS_ERR_INVALID, "aps_pde_solve: missing required pointer");
    if (a->field == APS_PDE_FIELD_KERNEL && (!a->kernel || !a->radius)) return fail(APS_ERR_INVALID, "aps_pde_solve: kernel field needs kernel and radius");
    if (a->n_tracers > 0 && (a->window < 1 || !a->seeds || !a->tracer_pos || !a->tracer_state || !a->tracer_hist || !a->v_eff_series || !a->D_eff_series))
        return fail(APS_ERR_INVALID, "aps_pde_solve: tracers need window >= 1, seeds, state, ring and output series");
    const size_t smem = aps::pde_smem_bytes(a->L, a->bc, a->n_tracers);
    if (smem == 0) return fail(APS_ERR_CAPACITY, "aps_pde_solve: need 8 <= L <= 2048 and n_tracers <= 2048");
    if (smem > 227 * 1024) return fail(APS_ERR_CAPACITY, "aps_pde_solve: instance does not fit in shared memory");
    if (count_sm100() == 0) return fail(APS_ERR_NO_DEVICE, "no sm_100 CUDA device visible; this library has no CPU path");
    CU(aps::pde_launch(*a, (cudaStream_t)stream));
    g_launches.fetch_add(1);
    return APS_OK;
}

// Test / tuning hooks (declared in include/aps.h).
void aps_debug_set_guard_scale(double s) { g_guard_scale = s; }
void aps_debug_set_k1_threads(int nt) { g_k1_threads = (nt == 32 || nt == 64 || nt == 128 || nt == 256) ? nt : 0; }
void aps_debug_set_use_lut(int on) { g_use_lut = on ? 1 : 0; }
void aps_debug_set_k2_ctas_per_sm(int n) { g_k2_ctas_per_sm = n > 0 ? n : 6; }
void aps_debug_set_k2_stash_cap(int n) { g_k2_stash_cap = (n == 1 || n == 2 || n == 4 || n == 8 || n == 16 || n == 32) ? n : 0; }
void aps_debug_set_reduce_impl(int v) { g_reduce_impl = v == 1 ? 1 : 0; }
void aps_debug_set_reduce_threads(int n) { g_reduce_threads = (n >= 32 && n <= 1024 && n % 32 == 0) ? n : 128; }
void aps_debug_set_use_fast(int on) { g_use_fast = on; }   // 0 generic only, 1 capacity classes, 2 run-time layout

}  // extern "C"
