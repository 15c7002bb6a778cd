// gpc_device.cuh — device-side canonical arithmetic shared by every kernel.
//
// All translation units are compiled with -fmad=false: the compiler never contracts a*b+c,
// every fused multiply-add is an explicit fma().  FP64 add/mul/fma/div/sqrt are IEEE
// round-to-nearest on sm_100a, so a scalar CPU program performing the same operation
// sequence produces the same bits (that program is the test oracle, which keeps its own
// independent copy of these definitions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Device-side bounds traps for the indices the kernels compute themselves (hand-off slots, overflow queues, tile slots):
// compiled in with `GPC_DEBUG=1 python -m gp_compressor_b200.build --force` (-DGPC_DEBUG_ASSERTS); compute-sanitizer is closed on
// the GPU pool, so this is the memory-safety check that can actually be run there.
#ifdef GPC_DEBUG_ASSERTS
#include <cassert>
#define GPC_DASSERT(c) assert(c)
#else
#define GPC_DASSERT(c) ((void)0)
#endif

namespace gpc {

// float literals of the reference widened to double (sparse_gp.hpp:124,146,229,236)
__device__ __forceinline__ double tiny12() { return (double)1e-12f; }
__device__ __forceinline__ double geo9() { return (double)1e-9f; }

// Constants of the canonical exp in constant memory: loaded as uniform-register pairs (LDCU.128) instead of two UMOV per
// 64-bit literal and use.
static __constant__ double GPC_EXPC[18] = {
    1.4426950408889634,          // 0  1/ln2
    6755399441055744.0,          // 1  1.5 * 2^52
    -0x1.62e42fefa39efp-1,       // 2  -ln2_hi
    -0x1.abc9e3b39803fp-56,      // 3  -ln2_lo
    1.0 / 6227020800.0,          // 4  1/13!
    1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,                     // 5..15
    -6755399441055744.0,         // 16
    0x1p-1020                    // 17
};

// Canonical exp: replaces glibc exp at rbf_kernel.cpp:17.  n = rint(x/ln2) by the
// 1.5*2^52 trick, Cody-Waite reduction with two fma, degree-13 Taylor Horner with fma,
// exponent insertion.  <= 1 ulp from libm.
__device__ __forceinline__ double gpc_exp(double x) {
    // the special cases (NaN, overflow, underflow) are selected at the end: no branches on the hot path
    const double xc = fmin(fmax(x, -746.0), 710.0);  // NaN -> -746 (fmax/fmin return the number)
    double t = __dmul_rn(xc, GPC_EXPC[0]);
    double kd = __dadd_rn(t, GPC_EXPC[1]);
    int n = (int)(unsigned int)(unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, GPC_EXPC[16]);
    double r = fma(kd, GPC_EXPC[2], xc);
    r = fma(kd, GPC_EXPC[3], r);
    double p = GPC_EXPC[4];
    p = fma(p, r, GPC_EXPC[5]);
    p = fma(p, r, GPC_EXPC[6]);
    p = fma(p, r, GPC_EXPC[7]);
    p = fma(p, r, GPC_EXPC[8]);
    p = fma(p, r, GPC_EXPC[9]);
    p = fma(p, r, GPC_EXPC[10]);
    p = fma(p, r, GPC_EXPC[11]);
    p = fma(p, r, GPC_EXPC[12]);
    p = fma(p, r, GPC_EXPC[13]);
    p = fma(p, r, GPC_EXPC[14]);
    p = fma(p, r, GPC_EXPC[15]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const bool sub = n < -1020;           // subnormal result: scale in two exact steps
    n += sub ? 1020 : 0;
    long long pb = __double_as_longlong(p) + ((long long)n << 52);
    p = __longlong_as_double(pb);
    p = __dmul_rn(p, sub ? 0x1p-1020 : 1.0);
    if (x > 709.0) p = __longlong_as_double(0x7ff0000000000000LL);
    if (x < -745.0) p = 0.0;
    if (x != x) p = x;
    return p;
}

// gpc_exp for arguments that are <= 0 or NaN (the RBF exponent): same arithmetic, fewer selects
__device__ __forceinline__ double gpc_exp_nonpos(double x) {
    const bool special = !(x >= -745.0);          // underflow to 0, or NaN
    const double xc = special ? -745.0 : x;
    double t = __dmul_rn(xc, GPC_EXPC[0]);
    double kd = __dadd_rn(t, GPC_EXPC[1]);
    int n = (int)(unsigned int)(unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, GPC_EXPC[16]);
    double r = fma(kd, GPC_EXPC[2], xc);
    r = fma(kd, GPC_EXPC[3], r);
    double p = GPC_EXPC[4];
    p = fma(p, r, GPC_EXPC[5]);
    p = fma(p, r, GPC_EXPC[6]);
    p = fma(p, r, GPC_EXPC[7]);
    p = fma(p, r, GPC_EXPC[8]);
    p = fma(p, r, GPC_EXPC[9]);
    p = fma(p, r, GPC_EXPC[10]);
    p = fma(p, r, GPC_EXPC[11]);
    p = fma(p, r, GPC_EXPC[12]);
    p = fma(p, r, GPC_EXPC[13]);
    p = fma(p, r, GPC_EXPC[14]);
    p = fma(p, r, GPC_EXPC[15]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const bool sub = n < -1020;
    n += sub ? 1020 : 0;
    p = __longlong_as_double(__double_as_longlong(p) + ((long long)n << 52));
    p = __dmul_rn(p, sub ? 0x1p-1020 : 1.0);
    if (special) p = (x != x) ? x : 0.0;
    return p;
}

// rbf_kernel::kernel_function, rbf_kernel.cpp:15-18 : p0 * exp((-0.5f/p1) * ||xi-xj||^2)
__device__ __forceinline__ double rbf(double x1, double x2, double b1, double b2, double p0, double cl) {
    double d1 = __dadd_rn(x1, -b1), d2 = __dadd_rn(x2, -b2);
    double sq = __dadd_rn(__dmul_rn(d1, d1), __dmul_rn(d2, d2));
    return __dmul_rn(p0, gpc_exp_nonpos(__dmul_rn(cl, sq)));  // cl < 0 (l_sq > 0 is checked at gpc_create), sq >= 0
}

__device__ __forceinline__ double shfl_xor_d(double v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}
// 5-step xor butterfly; every lane ends with the same bits (a+b == b+a)
__device__ __forceinline__ double butterfly32(double p) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, shfl_xor_d(p, off));
    return p;
}
// canonical dot: 32 lane-strided fma partials + butterfly.  a, b in shared or global memory.
__device__ __forceinline__ double warp_dot32(const double* a, const double* b, int n, int lane) {
    double p = 0.0;
    for (int j = lane; j < n; j += 32) p = fma(a[j], b[j], p);
    return butterfly32(p);
}
// canonical row product: 4 strided fma partials, (a0+a1)+(a2+a3).  m[j*stride] is element j.
__device__ __forceinline__ double row4(const double* m, int stride, const double* k, int n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 3 < n; j += 4) {
        a0 = fma(m[(size_t)j * stride], k[j], a0);
        a1 = fma(m[(size_t)(j + 1) * stride], k[j + 1], a1);
        a2 = fma(m[(size_t)(j + 2) * stride], k[j + 2], a2);
        a3 = fma(m[(size_t)(j + 3) * stride], k[j + 3], a3);
    }
    if (j < n) a0 = fma(m[(size_t)j * stride], k[j], a0);
    if (j + 1 < n) a1 = fma(m[(size_t)(j + 1) * stride], k[j + 1], a1);
    if (j + 2 < n) a2 = fma(m[(size_t)(j + 2) * stride], k[j + 2], a2);
    return __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
}

}  // namespace gpc
