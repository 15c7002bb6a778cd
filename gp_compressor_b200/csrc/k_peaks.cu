// k_peaks.cu — micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not hold:
// FP64 FMA throughput (independent DFMA chains) and shared-memory bandwidth (conflict-free 128-bit
// loads).  Called through gpc_debug_peak(); results go to profiles/ and bench.py.
#include "gpc_internal.h"

namespace gpc {

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(256) lds_kernel(double* out, int iters) {
    __shared__ __align__(16) double buf[4096];  // 32 KB
    for (int i = threadIdx.x; i < 4096; i += 256) buf[i] = (double)i;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    double2 acc0 = make_double2(0, 0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    unsigned off = threadIdx.x * 16;  // consecutive 16-byte words: conflict-free
    for (int i = 0; i < iters; i++) {
        const double2 v0 = lds128(base + ((off) & 32767u));
        const double2 v1 = lds128(base + ((off + 4096u) & 32767u));
        const double2 v2 = lds128(base + ((off + 8192u) & 32767u));
        const double2 v3 = lds128(base + ((off + 12288u) & 32767u));
        acc0.x += v0.x; acc0.y += v0.y; acc1.x += v1.x; acc1.y += v1.y;
        acc2.x += v2.x; acc2.y += v2.y; acc3.x += v3.x; acc3.y += v3.y;
        off += 16384u;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (acc0.x + acc0.y) + (acc1.x + acc1.y) + (acc2.x + acc2.y) + (acc3.x + acc3.y);
}

__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(256) lds64_kernel(double* out, int iters) {
    __shared__ __align__(16) double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) buf[i] = (double)i;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned off = threadIdx.x * 8;
    for (int i = 0; i < iters; i++) {
        a0 += lds64(base + ((off) & 32767u));
        a1 += lds64(base + ((off + 2048u) & 32767u));
        a2 += lds64(base + ((off + 4096u) & 32767u));
        a3 += lds64(base + ((off + 6144u) & 32767u));
        off += 8192u;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
}

}  // namespace

// kind 0: FP64 FLOP/s (2 per DFMA); kind 1: shared-memory bytes/s.  Best of `reps` timed launches.
cudaError_t measure_peak(int kind, int sm_count, cudaStream_t s, double* value) {
    double* out = nullptr;
    const int blocks = sm_count * 8, threads = 256;
    cudaError_t e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = kind == 0 ? 20000 : 4000;
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, s);
        if (kind == 0) dfma_kernel<<<blocks, threads, 0, s>>>(out, iters, 1.0 + rep);
        else if (kind == 1) lds_kernel<<<blocks, threads, 0, s>>>(out, iters);
        else lds64_kernel<<<blocks, threads, 0, s>>>(out, iters);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double work = kind == 0 ? (double)blocks * threads * iters * 8 * 2.0
                            : (double)blocks * threads * iters * 4 * (kind == 1 ? 16.0 : 8.0);
        const double v = work / (ms * 1e-3);
        if (rep > 0 && v > best) best = v;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e = cudaGetLastError();
    cudaFree(out);
    *value = best;
    return e;
}

}  // namespace gpc
