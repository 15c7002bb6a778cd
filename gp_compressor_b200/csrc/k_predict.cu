// k_predict.cu — K8: decompressor grid evaluation K(x*, BV) . alpha and reprojection.
//
// Replaces gp_compressor::load_compressed (/root/reference/src/gp_compressor.cpp:267-386)
// and sparse_gp::predict_measurements / predict (sparse_gp.hpp:299-351).  The reference
// also computes the predictive variance k' C k per grid point and throws it away
// (gp_compressor.cpp:333 never reads V_star); this kernel computes the mean only.
//
// One CTA per non-empty patch: alpha and BV staged in shared memory, each thread owns
// grid points m = t, t + NT, ... (y outer, x inner as in :320-332), evaluates N RBF
// kernels with the canonical 4-partial dot, applies the patch frame (rotation rebuilt
// from the stored quaternion as Eigen's toRotationMatrix does, :339) and writes one
// 32-byte PointXYZRGB record with two 16-byte stores.
#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace gpc {

namespace {

constexpr int PRED_T = 128;

// gp_compressor::flatten_colors, gp_compressor.cpp:251-265 (x86 double -> short cast)
__device__ __forceinline__ unsigned int flatten_color(double x) {
    if (x != x || fabs(x) == __longlong_as_double(0x7ff0000000000000LL)) return 255u;
    int v = (fabs(x) < 2147483648.0) ? (int)x : 0;
    short s = (short)v;
    if (s < 0) return 0u;
    if (s > 255) return 255u;
    return (unsigned int)s;
}

__device__ __forceinline__ void quat_to_rot(const double* q, double* R) {
    double tx = __dmul_rn(2.0, q[0]), ty = __dmul_rn(2.0, q[1]), tz = __dmul_rn(2.0, q[2]);
    double twx = __dmul_rn(tx, q[3]), twy = __dmul_rn(ty, q[3]), twz = __dmul_rn(tz, q[3]);
    double txx = __dmul_rn(tx, q[0]), txy = __dmul_rn(ty, q[0]), txz = __dmul_rn(tz, q[0]);
    double tyy = __dmul_rn(ty, q[1]), tyz = __dmul_rn(tz, q[1]), tzz = __dmul_rn(tz, q[2]);
    R[0] = __dadd_rn(1.0, -__dadd_rn(tyy, tzz)); R[1] = __dadd_rn(txy, -twz); R[2] = __dadd_rn(txz, twy);
    R[3] = __dadd_rn(txy, twz); R[4] = __dadd_rn(1.0, -__dadd_rn(txx, tzz)); R[5] = __dadd_rn(tyz, -twx);
    R[6] = __dadd_rn(txz, -twy); R[7] = __dadd_rn(tyz, twx); R[8] = __dadd_rn(1.0, -__dadd_rn(txx, tyy));
}

__global__ void __launch_bounds__(PRED_T) predict_grid_kernel(PredictArgs a) {
    extern __shared__ double sm[];
    const int64_t p = blockIdx.x;
    const int N = a.nbv[p];
    if (N == 0) return;  // gp_compressor.cpp:299
    double* al = sm;
    double* b1 = al + N;
    double* b2 = b1 + N;
    // RGB field GP of the patch (sparse_gp_field::predict, sparse_gp_field.hpp:284-320): its own BVs, 3 alphas
    const int NR = a.rgb_nbv ? a.rgb_nbv[p] : 0;
    double* ra0 = b2 + N;
    double* ra1 = ra0 + NR;
    double* ra2 = ra1 + NR;
    double* rb1 = ra2 + NR;
    double* rb2 = rb1 + NR;
    __shared__ double R[9], mean[3], cmean[3];
    __shared__ unsigned int rgba;
    const int t = threadIdx.x;
    const int64_t pb = p * a.stride;
    for (int i = t; i < N; i += PRED_T) { al[i] = a.alpha[pb + i]; b1[i] = a.b1[pb + i]; b2[i] = a.b2[pb + i]; }
    for (int i = t; i < NR; i += PRED_T) {
        ra0[i] = a.rgb_alpha[0][pb + i]; ra1[i] = a.rgb_alpha[1][pb + i]; ra2[i] = a.rgb_alpha[2][pb + i];
        rb1[i] = a.rgb_b1[pb + i]; rb2[i] = a.rgb_b2[pb + i];
    }
    if (t == 0) {
        if (a.quat) {
            quat_to_rot(a.quat + 4 * p, R);
            for (int d = 0; d < 3; d++) { mean[d] = a.mean[3 * p + d]; cmean[d] = a.rgbmean[3 * p + d]; }
            unsigned int r = flatten_color(a.rgbmean[3 * p + 0]), g = flatten_color(a.rgbmean[3 * p + 1]),
                         b = flatten_color(a.rgbmean[3 * p + 2]);
            rgba = b | (g << 8) | (r << 16) | (255u << 24);
        } else {
            for (int d = 0; d < 9; d++) R[d] = (d % 4 == 0) ? 1.0 : 0.0;
            mean[0] = mean[1] = mean[2] = 0.0;
            cmean[0] = cmean[1] = cmean[2] = 0.0;
            rgba = 255u << 24;
        }
    }
    __syncthreads();
    const int sz = a.sz, g2 = sz * sz;
    const int64_t base = a.slot[p] * g2;
    const double dsz = (double)sz;
    for (int m = t; m < g2; m += PRED_T) {
        const int yy = m / sz, xx = m - yy * sz;
        // res*((double(x) + 0.5f)/double(sz) - 0.5f), gp_compressor.cpp:326-327
        const double X0 = __dmul_rn(a.res, __dadd_rn(__ddiv_rn(__dadd_rn((double)xx, 0.5), dsz), -0.5));
        const double X1 = __dmul_rn(a.res, __dadd_rn(__ddiv_rn(__dadd_rn((double)yy, 0.5), dsz), -0.5));
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = 0;
        for (; i + 3 < N; i += 4) {
            a0 = fma(al[i], rbf(X0, X1, b1[i], b2[i], a.p0, a.cl), a0);
            a1 = fma(al[i + 1], rbf(X0, X1, b1[i + 1], b2[i + 1], a.p0, a.cl), a1);
            a2 = fma(al[i + 2], rbf(X0, X1, b1[i + 2], b2[i + 2], a.p0, a.cl), a2);
            a3 = fma(al[i + 3], rbf(X0, X1, b1[i + 3], b2[i + 3], a.p0, a.cl), a3);
        }
        if (i < N) a0 = fma(al[i], rbf(X0, X1, b1[i], b2[i], a.p0, a.cl), a0);
        if (i + 1 < N) a1 = fma(al[i + 1], rbf(X0, X1, b1[i + 1], b2[i + 1], a.p0, a.cl), a1);
        if (i + 2 < N) a2 = fma(al[i + 2], rbf(X0, X1, b1[i + 2], b2[i + 2], a.p0, a.cl), a2);
        const double f = __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
        if (a.heights) a.heights[base + m] = f;
        if (a.out32) {
            float o[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                double v = __dadd_rn(__dadd_rn(__dmul_rn(R[d * 3 + 0], f), __dmul_rn(R[d * 3 + 1], X0)), __dmul_rn(R[d * 3 + 2], X1));
                o[d] = (float)__dadd_rn(v, mean[d]);
            }
            unsigned int col = rgba;
            if (a.rgb_nbv) {  // c = C_star.row(m) + RGB_means[i], gp_compressor.cpp:367
                double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
                for (int i = 0; i < NR; i++) {
                    const double k = rbf(X0, X1, rb1[i], rb2[i], a.p0, a.cl);
                    c0[i & 3] = fma(ra0[i], k, c0[i & 3]);
                    c1[i & 3] = fma(ra1[i], k, c1[i & 3]);
                    c2[i & 3] = fma(ra2[i], k, c2[i & 3]);
                }
                const double fr = __dadd_rn(__dadd_rn(c0[0], c0[1]), __dadd_rn(c0[2], c0[3]));
                const double fg = __dadd_rn(__dadd_rn(c1[0], c1[1]), __dadd_rn(c1[2], c1[3]));
                const double fb = __dadd_rn(__dadd_rn(c2[0], c2[1]), __dadd_rn(c2[2], c2[3]));
                const unsigned int rr = flatten_color(__dadd_rn(fr, cmean[0])), gg = flatten_color(__dadd_rn(fg, cmean[1])),
                                   bb = flatten_color(__dadd_rn(fb, cmean[2]));
                col = bb | (gg << 8) | (rr << 16) | (255u << 24);
            }
            float4* dst = reinterpret_cast<float4*>(a.out32 + (size_t)(base + m) * GPC_POINT_BYTES);
            dst[0] = make_float4(o[0], o[1], o[2], 1.0f);
            dst[1] = make_float4(__uint_as_float(col), 0.0f, 0.0f, 0.0f);
        }
    }
}

// sparse_gp::predict at arbitrary points of one patch (sparse_gp.hpp:312-351), optional sigma
__global__ void __launch_bounds__(128) predict_points_kernel(const double* __restrict__ alpha, const double* __restrict__ b1,
                                                             const double* __restrict__ b2, int N, const double* __restrict__ C,
                                                             double p0, double cl, double s20, const double* __restrict__ X,
                                                             int64_t m, double* __restrict__ f, double* __restrict__ sigma) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const double x1 = X[2 * t], x2 = X[2 * t + 1];
    if (N == 0) {
        f[t] = 0.0;
        if (sigma) sigma[t] = sqrt(__dadd_rn(p0, s20));
        return;
    }
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < N; i++) a[i & 3] = fma(alpha[i], rbf(x1, x2, b1[i], b2[i], p0, cl), a[i & 3]);
    f[t] = __dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3]));
    if (sigma) {
        // sqrt(s20 + kstar + k'Ck): canonical order = dot32 over i of k_i * row4(C_i, k)
        double part[32];
        for (int l = 0; l < 32; l++) part[l] = 0.0;
        for (int i = 0; i < N; i++) {
            double r[4] = {0.0, 0.0, 0.0, 0.0};
            for (int j = 0; j < N; j++) r[j & 3] = fma(C[(size_t)i * N + j], rbf(x1, x2, b1[j], b2[j], p0, cl), r[j & 3]);
            double ck = __dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3]));
            part[i & 31] = fma(rbf(x1, x2, b1[i], b2[i], p0, cl), ck, part[i & 31]);
        }
        for (int off = 16; off >= 1; off >>= 1) {
            double tmp[32];
            for (int l = 0; l < 32; l++) tmp[l] = __dadd_rn(part[l], part[l ^ off]);
            for (int l = 0; l < 32; l++) part[l] = tmp[l];
        }
        double sg = __dadd_rn(__dadd_rn(s20, p0), part[0]);
        if (sg < 0.0) sg = 0.0;
        sigma[t] = sqrt(sg);
    }
}

}  // namespace

void launch_predict_grid(const PredictArgs& a, cudaStream_t s) {
    if (a.n_patches <= 0) return;
    size_t smem = (size_t)8 * a.stride * sizeof(double);
    predict_grid_kernel<<<(unsigned)a.n_patches, PRED_T, smem, s>>>(a);
    g_launches++;
}

void launch_predict_points(const double* alpha, const double* b1, const double* b2, int N, const double* C, double p0,
                           double cl, double s20, const double* X, int64_t m, double* f, double* sigma, cudaStream_t s) {
    if (m <= 0) return;
    predict_points_kernel<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(alpha, b1, b2, N, C, p0, cl, s20, X, m, f, sigma);
}

}  // namespace gpc
