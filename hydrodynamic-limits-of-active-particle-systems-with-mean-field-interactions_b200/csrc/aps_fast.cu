// aps_fast.cu — instantiations and dispatch of the K = 1 specialised K1 kernel (aps_k1_fast.cuh).
// Separate translation unit so that the two halves of the library compile in parallel.
#include <cstdlib>
#include <cuda_runtime.h>

#include "aps_k1_lean.cuh"

namespace aps {

// A/B knobs of include/aps.h, read ONCE when the library is loaded (never on the launch path)
static const int g_env_extra_smem = [] { const char* e = getenv("APS_K1_EXTRA_SMEM"); return e ? atoi(e) : 0; }();   // occupancy experiments only
static const bool g_env_no_lean = getenv("APS_K1_NO_LEAN") != nullptr;
template <int NT, int RCAP, int NCAP, int LPCAP>
static cudaError_t launch_nt(const K1Args& a, bool philox, cudaStream_t st) {
    const size_t smem = k1_fast_smem_bytes(a.p.L, a.b.n_max, a.p.radius, RCAP, NCAP, LPCAP) + (size_t)g_env_extra_smem;
    if (philox) {
        auto k = k1_fast_kernel<NT, true, RCAP, NCAP, LPCAP>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<a.b.n_replicas, NT, smem, st>>>(a);
    } else {
        auto k = k1_fast_kernel<NT, false, RCAP, NCAP, LPCAP>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<a.b.n_replicas, NT, smem, st>>>(a);
    }
    return cudaGetLastError();
}
template <int RCAP, int NCAP, int LPCAP>
static cudaError_t launch_class(const K1Args& a, bool philox, cudaStream_t st, int nt) {
    return nt == 32 ? launch_nt<32, RCAP, NCAP, LPCAP>(a, philox, st) : launch_nt<64, RCAP, NCAP, LPCAP>(a, philox, st);
}

// side stream for size classes of one batch that run concurrently (one per host thread and device, created on first use)
struct SideStream { int dev; cudaStream_t stream; cudaEvent_t fork, join; };
static SideStream* side_stream() {
    thread_local SideStream s{-1, nullptr, nullptr, nullptr};
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (s.dev == dev && s.stream) return &s;
    SideStream n{dev, nullptr, nullptr, nullptr};
    if (cudaStreamCreateWithFlags(&n.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&n.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&n.join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    s = n;                                              // (a thread that changes device leaks one stream and two events: by design rare)
    return &s;
}

template <int RCAP, int LPCAP, int NCAP, bool WHO>
static cudaError_t launch_lean(const K1Args& a, bool philox, cudaStream_t st) {
    const size_t smem = k1_lean_smem_bytes<RCAP, LPCAP, NCAP, WHO>();
    if (philox) {
        auto k = k1_lean_kernel<true, RCAP, LPCAP, NCAP, WHO>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<a.b.n_replicas, 32, smem, st>>>(a);
    } else {
        auto k = k1_lean_kernel<false, RCAP, LPCAP, NCAP, WHO>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<a.b.n_replicas, 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// Returns cudaSuccess and *launched = number of specialised kernels enqueued; *launched = 0 if the
// configuration does not qualify (the caller then uses the generic kernel only).
cudaError_t launch_fast(const K1Args& a, bool philox, cudaStream_t st, int allow_static, int nt, int auto_threads, int* launched) {
    *launched = 0;
    nt = nt == 32 ? 32 : 64;
    const int r1 = a.p.radius + 1, nm = a.b.n_max, lp = a.p.L + 2 * a.p.radius;
    if (k1_fast_smem_bytes(a.p.L, nm, a.p.radius, 0, 0, 0) > 227 * 1024) return cudaSuccess;
    *launched = 1;
    if (allow_static) {
        if (r1 <= 21 && nm <= 512 && lp <= 1056) {
            if (nt == 32 && !g_env_no_lean) {
                // half-size shared-memory image: all replicas of an SM resident at once; its rejects (unsorted initial
                // positions) go through the full-size kernel in a second, otherwise empty launch
                cudaError_t e = launch_lean<21, 1056, 512, false>(a, philox, st);
                if (e != cudaSuccess) return e;
                K1Args b = a;
                b.only_retry = 2;
                e = launch_lean<21, 1056, 512, true>(b, philox, st);      // unsorted rejects: same image + site map, 22 per SM
                if (e != cudaSuccess) return e;
                *launched = 3;
                return launch_class<21, 512, 1056>(b, philox, st, nt);        // what is left (n > 488)
            }
            return launch_class<21, 512, 1056>(a, philox, st, nt);
        }
        if (r1 <= 21 && nm <= 1024 && lp <= 1056) {
            if (nt == 32 && !g_env_no_lean) {
                // mid-size replicas (config 3: N = 900): trimmed image, sorted particles first (17 replicas per SM), then the
                // variant WITH the site map for unsorted ones (15 per SM; the full-size kernel holds 10); what both reject
                // (n > 968, K = 1 violated) falls through to the full-size and generic kernels
                cudaError_t e = launch_lean<21, 1056, 1024, false>(a, philox, st);
                if (e != cudaSuccess) return e;
                K1Args b = a;
                b.only_retry = 2;
                e = launch_lean<21, 1056, 1024, true>(b, philox, st);
                if (e != cudaSuccess) return e;
                *launched = 3;
                return launch_class<21, 1024, 1056>(b, philox, st, nt);
            }
            return launch_class<21, 1024, 1056>(a, philox, st, nt);
        }
        if (r1 <= 81 && nm <= 1024 && lp <= 1184) {
            if (auto_threads && !g_env_no_lean) {            // (a forced thread count selects the two-warp kernel: A/B and its tests)
                // wide filters (config 4: r = 80, 30 .. 700 particles per replica in one batch): the trimmed one-warp image by
                // particle count — n <= 488 sorted (26 replicas per SM), then n <= 968 sorted (16 per SM), then the site-map variant
                // for unsorted replicas; the full-size two-warp kernel (an 11.6 KB pre-multiplied tap table per replica) takes the rest
                // The two sorted classes own disjoint replicas (by n), so they run CONCURRENTLY (fork / join on a side stream):
                // a launch lasts as long as its longest replica chain, and with few replicas per GPU (strong scaling) two partial
                // waves in sequence would double the step.
                SideStream* ss = side_stream();
                K1Args a1 = a, a2 = a;
                a1.n_hi = kLeanNMax;
                a2.n_lo = kLeanNMax;
                cudaError_t e = ss ? cudaEventRecord(ss->fork, st) : cudaSuccess;
                if (e != cudaSuccess) return e;
                e = launch_lean<81, 1184, 512, false>(a1, philox, st);
                if (e != cudaSuccess) return e;
                if (ss) {
                    if ((e = cudaStreamWaitEvent(ss->stream, ss->fork, 0)) != cudaSuccess) return e;
                    if ((e = launch_lean<81, 1184, 1024, false>(a2, philox, ss->stream)) != cudaSuccess) return e;
                    if ((e = cudaEventRecord(ss->join, ss->stream)) != cudaSuccess) return e;
                    if ((e = cudaStreamWaitEvent(st, ss->join, 0)) != cudaSuccess) return e;
                } else if ((e = launch_lean<81, 1184, 1024, false>(a2, philox, st)) != cudaSuccess) return e;
                K1Args b = a;
                b.only_retry = 2;
                e = launch_lean<81, 1184, 1024, true>(b, philox, st);
                if (e != cudaSuccess) return e;
                *launched = 4;
                return nm <= 512 ? launch_class<81, 512, 1184>(b, philox, st, nt) : launch_class<81, 1024, 1184>(b, philox, st, nt);
            }
            return nm <= 512 ? launch_class<81, 512, 1184>(a, philox, st, nt) : launch_class<81, 1024, 1184>(a, philox, st, nt);
        }
    }
    return launch_class<0, 0, 0>(a, philox, st, nt);
}

}  // namespace aps
