// smem_bw.cu — measured shared-memory read bandwidth of one GPU (the denominator of K1's roofline; MEASURED_PEAKS.json has no such figure).
// Every warp issues conflict-free 16-byte loads (LDS.128: 512 B per warp-instruction = 4 wavefronts of 128 B) from a shared-memory
// buffer in an unrolled loop with 8 independent loads in flight per thread; the loaded words are folded with XOR so the loads stay live.
// Reported: bytes loaded / elapsed (CUDA events), best of `reps` launches.  Test tool, not product code.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kThreads = 1024;
constexpr int kVecs = 2048;          // 32 KB of uint4 per CTA

__global__ void __launch_bounds__(kThreads) smem_read_kernel(uint32_t* out, int iters) {
    __shared__ uint4 buf[kVecs];
    for (int i = threadIdx.x; i < kVecs; i += kThreads) buf[i] = make_uint4(i, i * 3u, i * 5u, i * 7u);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 v = buf[(idx + k * 256) & (kVecs - 1)];        // consecutive lanes -> consecutive 16-byte vectors: no bank conflicts
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        idx = (idx + 32) & (kVecs - 1);
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) out[blockIdx.x] = acc.x;      // practically never: keeps the loop alive
}

extern "C" int smem_bw_measure(int reps, double* gbs_out, double* bytes_per_clk_per_sm_out, int* sm_count_out, int* clock_khz_out) {
    int dev = 0, sms = 0, khz = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t* out = nullptr;
    const int grid = sms * 2, iters = 20000;
    if (cudaMalloc(&out, grid * sizeof(uint32_t)) != cudaSuccess) return 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    smem_read_kernel<<<grid, kThreads>>>(out, 2000);                       // warm-up (clocks)
    smem_read_kernel<<<grid, kThreads>>>(out, iters);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        smem_read_kernel<<<grid, kThreads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 3;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * kThreads * (double)iters * 8.0 * 16.0;
        const double gbs = bytes / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gbs_out = best; *sm_count_out = sms; *clock_khz_out = khz;
    *bytes_per_clk_per_sm_out = best * 1e9 / ((double)sms * (double)khz * 1e3);
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
