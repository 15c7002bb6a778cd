// k_predict.cu — K8: decompressor grid evaluation K(x*, BV) . alpha and reprojection.
//
// Replaces gp_compressor::load_compressed (/root/reference/src/gp_compressor.cpp:267-386)
// and sparse_gp::predict_measurements / predict (sparse_gp.hpp:299-351).  The reference
// also computes the predictive variance k' C k per grid point and throws it away
// (gp_compressor.cpp:333 never reads V_star); this kernel computes the mean only.
//
// The grid is a lattice, so the RBF kernel is evaluated separably: per patch and basis vector i the kernel
// builds the tables  Ex_i[a] = p0 * exp(cl (X0_a - b1_i)^2)  and  Ey_i[b] = exp(cl (X1_b - b2_i)^2)  (N (sz + rows)
// exps instead of N sz^2) and takes  k_i(a, b) = Ex_i[a] * Ey_i[b]  — within 3 ulp of the reference's
// p0 * exp(cl ((X0-b1)^2 + (X1-b2)^2)); the oracle restates the same arithmetic, so outputs are bit-identical to it.
// One CTA per (patch, block of grid rows): tables in shared memory, each thread owns R consecutive rows of one
// column (y outer, x inner as in :320-332), accumulates the canonical 4-partial dot with one DMUL + one DFMA per
// basis vector and point, applies the patch frame (rotation rebuilt from the stored quaternion as Eigen's
// toRotationMatrix does, :339) and writes one 32-byte PointXYZRGB record with two 16-byte stores.
#include <algorithm>

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

// x, y, z, 1.0f | rgba, 0, 0, 0 : 32 bytes, 32-byte aligned, written with st.global.v8.b32 (sm_100)
__device__ __forceinline__ void store_record(uint8_t* dst, float x, float y, float z, unsigned int rgba) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)),
                 "r"(__float_as_uint(z)), "r"(0x3f800000u), "r"(rgba), "r"(0u), "r"(0u), "r"(0u)
                 : "memory");
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

// One (patch, block of grid rows) per group of G threads: G = 32 (four independent warps per CTA, for small grids)
// or G = PRED_T.  Each thread owns R consecutive rows of one grid column.
// MINB: resident CTAs per SM the register budget is cut for (large-grid shape: 4 when the tables leave room for 4 CTAs, else 3)
// DIRECT: k_i = rbf_kernel::kernel_function(x*, BV_i) = p0 exp(cl (dx^2 + dy^2)) per (grid point, BV), the reference's
// arithmetic (rbf_kernel.cpp:15-18); otherwise the separable tables (gpc_config.decode_separable).
template <int R, int G, int MINB, bool DIRECT>
__global__ void __launch_bounds__(PRED_T, MINB) predict_grid_kernel(PredictArgs a) {
    extern __shared__ double sm_all[];
    const int64_t gi = (int64_t)blockIdx.x * (PRED_T / G) + threadIdx.x / G;
    if (gi >= a.n_groups) return;
    const int64_t p = gi / a.nblk;
    const int rb = (int)(gi - p * a.nblk);
    const int N = a.nbv[p];
    if (N == 0) return;  // gp_compressor.cpp:299
    // RGB field GP of the patch (sparse_gp_field::predict, sparse_gp_field.hpp:284-320): its own BVs, 3 alphas
    const int NR = a.rgb_nbv ? a.rgb_nbv[p] : 0;
    const int sz = a.sz, rowsP = a.rowsP;
    const int row0 = rb * a.rows;
    const int rows_here = min(a.rows, sz - row0);
    double* sm = sm_all + (size_t)(threadIdx.x / G) * a.group_doubles;
    double* Rm = sm;                 // 9 rotation entries, mean[3] at 9, cmean[3] at 12, packed mean colour at 15
    double* mean = Rm + 9;
    double* cmean = Rm + 12;
    double* Ey = sm + 16;            // first: 16-byte aligned rows for the paired loads (rowsP is even when R > 1)
    double* Xs = Ey + (size_t)N * rowsP;  // X0 per grid column
    double* Ys = Xs + sz;            // X1 per grid row of this block
    double* al = Ys + rowsP;
    double* sb1 = al + N;
    double* sb2 = sb1 + N;
    double* Ex = sb2 + N;
    double* ra0 = Ex + (size_t)N * sz;
    double* ra1 = ra0 + NR;
    double* ra2 = ra1 + NR;
    double* rb1 = ra2 + NR;
    double* rb2 = rb1 + NR;
    double* REx = rb2 + NR;
    double* REy = REx + (size_t)NR * sz;
    const int t = threadIdx.x % G;
    const int64_t pb = p * a.stride;
    const double dsz = (double)sz;
    // res*((double(x) + 0.5f)/double(sz) - 0.5f), gp_compressor.cpp:326-327
    for (int i = t; i < sz; i += G) Xs[i] = __dmul_rn(a.res, __dadd_rn(__ddiv_rn(__dadd_rn((double)i, 0.5), dsz), -0.5));
    for (int i = t; i < rowsP; i += G)
        Ys[i] = __dmul_rn(a.res, __dadd_rn(__ddiv_rn(__dadd_rn((double)(row0 + i), 0.5), dsz), -0.5));
    for (int i = t; i < N; i += G) { al[i] = a.alpha[pb + i]; sb1[i] = a.b1[pb + i]; sb2[i] = a.b2[pb + i]; }
    for (int i = t; i < NR; i += G) {
        ra0[i] = a.rgb_alpha[0][pb + i]; ra1[i] = a.rgb_alpha[1][pb + i]; ra2[i] = a.rgb_alpha[2][pb + i];
        rb1[i] = a.rgb_b1[pb + i]; rb2[i] = a.rgb_b2[pb + i];
    }
    if (t == G - 1) {
        unsigned int rgba = 255u << 24;
        if (a.quat) {
            quat_to_rot(a.quat + 4 * p, Rm);
            for (int d = 0; d < 3; d++) { mean[d] = a.mean[3 * p + d]; cmean[d] = a.rgbmean[3 * p + d]; }
            unsigned int r = flatten_color(a.rgbmean[3 * p + 0]), g = flatten_color(a.rgbmean[3 * p + 1]),
                         b = flatten_color(a.rgbmean[3 * p + 2]);
            rgba = b | (g << 8) | (r << 16) | (255u << 24);
        } else {
            for (int d = 0; d < 9; d++) Rm[d] = (d % 4 == 0) ? 1.0 : 0.0;
            mean[0] = mean[1] = mean[2] = 0.0;
            cmean[0] = cmean[1] = cmean[2] = 0.0;
        }
        reinterpret_cast<unsigned int*>(Rm + 15)[0] = rgba;
    }
    if (G == 32) __syncwarp(); else __syncthreads();
    // separable kernel tables
    if (!DIRECT) {
    for (int e = t; e < N * sz; e += G) {
        const int i = e / sz, c = e - i * sz;
        const double d = __dadd_rn(Xs[c], -sb1[i]);
        Ex[e] = __dmul_rn(a.p0, gpc_exp_nonpos(__dmul_rn(a.cl, __dmul_rn(d, d))));
    }
    for (int e = t; e < N * rowsP; e += G) {
        const int i = e / rowsP, c = e - i * rowsP;
        const double d = __dadd_rn(Ys[c], -sb2[i]);
        Ey[e] = (c < rows_here) ? gpc_exp_nonpos(__dmul_rn(a.cl, __dmul_rn(d, d))) : 0.0;
    }
    for (int e = t; e < NR * sz; e += G) {
        const int i = e / sz, c = e - i * sz;
        const double d = __dadd_rn(Xs[c], -rb1[i]);
        REx[e] = __dmul_rn(a.p0, gpc_exp_nonpos(__dmul_rn(a.cl, __dmul_rn(d, d))));
    }
    for (int e = t; e < NR * rowsP; e += G) {
        const int i = e / rowsP, c = e - i * rowsP;
        const double d = __dadd_rn(Ys[c], -rb2[i]);
        REy[e] = (c < rows_here) ? gpc_exp_nonpos(__dmul_rn(a.cl, __dmul_rn(d, d))) : 0.0;
    }
    if (G == 32) __syncwarp(); else __syncthreads();
    }
    const unsigned int rgba = reinterpret_cast<const unsigned int*>(Rm + 15)[0];
    const int g2 = sz * sz;
    const int64_t base = a.slot[p] * g2 + (int64_t)row0 * sz;
    const int ngroups = sz * ((rows_here + R - 1) / R);
    for (int g = t; g < ngroups; g += G) {
        const int yq = g / sz, xx = g - yq * sz;
        const int yl0 = yq * R;
        double acc[R][4];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0;
        if (DIRECT) {
            const double gx = Xs[xx];
            double gy[R];
#pragma unroll
            for (int r = 0; r < R; r++) gy[r] = Ys[min(yl0 + r, rowsP - 1)];
            int i = 0;
            for (; i + 3 < N; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double w = al[i + u], bx = sb1[i + u], by = sb2[i + u];
#pragma unroll
                    for (int r = 0; r < R; r++) acc[r][u] = fma(w, rbf(gx, gy[r], bx, by, a.p0, a.cl), acc[r][u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 3; u++)
                if (i + u < N) {
                    const double w = al[i + u], bx = sb1[i + u], by = sb2[i + u];
#pragma unroll
                    for (int r = 0; r < R; r++) acc[r][u] = fma(w, rbf(gx, gy[r], bx, by, a.p0, a.cl), acc[r][u]);
                }
        } else {
        const double* ex = Ex + xx;
        const double* ey = Ey + yl0;
        int i = 0;
        for (; i + 3 < N; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double e = ex[(i + u) * sz], w = al[i + u];
                double eyv[R];
                if (R % 2 == 0) {  // rowsP and yl0 are multiples of R: 16-byte aligned pairs
#pragma unroll
                    for (int r = 0; r < R; r += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(ey + (i + u) * rowsP + r);
                        eyv[r] = v.x; eyv[r + 1 < R ? r + 1 : r] = v.y;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) eyv[r] = ey[(i + u) * rowsP + r];
                }
#pragma unroll
                for (int r = 0; r < R; r++) acc[r][u] = fma(w, __dmul_rn(e, eyv[r]), acc[r][u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 3; u++)
            if (i + u < N) {
                const double e = ex[(i + u) * sz], w = al[i + u];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r][u] = fma(w, __dmul_rn(e, ey[(i + u) * rowsP + r]), acc[r][u]);
            }
        }
        const double X0 = Xs[xx];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (yl0 + r >= rows_here) break;
            const int m = (yl0 + r) * sz + xx;
            const double X1 = Ys[yl0 + r];
            const double f = __dadd_rn(__dadd_rn(acc[r][0], acc[r][1]), __dadd_rn(acc[r][2], acc[r][3]));
            if (a.heights) a.heights[base + m] = f;
            if (a.out32) {
                float o[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    double v = __dadd_rn(__dadd_rn(__dmul_rn(Rm[d * 3 + 0], f), __dmul_rn(Rm[d * 3 + 1], X0)), __dmul_rn(Rm[d * 3 + 2], X1));
                    o[d] = (float)__dadd_rn(v, mean[d]);
                }
                unsigned int col = rgba;
                if (a.rgb_nbv) {  // c = C_star.row(m) + RGB_means[i], gp_compressor.cpp:367
                    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
                    if (DIRECT) {
                        int j = 0;
                        for (; j + 3 < NR; j += 4) {
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const double k = rbf(X0, X1, rb1[j + u], rb2[j + u], a.p0, a.cl);
                                c0[u] = fma(ra0[j + u], k, c0[u]);
                                c1[u] = fma(ra1[j + u], k, c1[u]);
                                c2[u] = fma(ra2[j + u], k, c2[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 3; u++)
                            if (j + u < NR) {
                                const double k = rbf(X0, X1, rb1[j + u], rb2[j + u], a.p0, a.cl);
                                c0[u] = fma(ra0[j + u], k, c0[u]);
                                c1[u] = fma(ra1[j + u], k, c1[u]);
                                c2[u] = fma(ra2[j + u], k, c2[u]);
                            }
                    } else {
                    const double* rex = REx + xx;
                    const double* rey = REy + yl0 + r;
                    int j = 0;
                    for (; j + 3 < NR; j += 4) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const double k = __dmul_rn(rex[(j + u) * sz], rey[(j + u) * rowsP]);
                            c0[u] = fma(ra0[j + u], k, c0[u]);
                            c1[u] = fma(ra1[j + u], k, c1[u]);
                            c2[u] = fma(ra2[j + u], k, c2[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 3; u++)
                        if (j + u < NR) {
                            const double k = __dmul_rn(rex[(j + u) * sz], rey[(j + u) * rowsP]);
                            c0[u] = fma(ra0[j + u], k, c0[u]);
                            c1[u] = fma(ra1[j + u], k, c1[u]);
                            c2[u] = fma(ra2[j + u], k, c2[u]);
                        }
                    }
                    const double fr = __dadd_rn(__dadd_rn(c0[0], c0[1]), __dadd_rn(c0[2], c0[3]));
                    const double fg = __dadd_rn(__dadd_rn(c1[0], c1[1]), __dadd_rn(c1[2], c1[3]));
                    const double fb = __dadd_rn(__dadd_rn(c2[0], c2[1]), __dadd_rn(c2[2], c2[3]));
                    const unsigned int rr = flatten_color(__dadd_rn(fr, cmean[0])), gg = flatten_color(__dadd_rn(fg, cmean[1])),
                                       bb = flatten_color(__dadd_rn(fb, cmean[2]));
                    col = bb | (gg << 8) | (rr << 16) | (255u << 24);
                }
                // one 256-bit store per PointXYZRGB record (a whole 32-byte sector: no partial-sector write traffic)
                store_record(a.out32 + (size_t)(base + m) * GPC_POINT_BYTES, o[0], o[1], o[2], col);
            }
        }
    }
}

// sparse_gp::predict at arbitrary points of one patch (sparse_gp.hpp:312-327): the mean.  Sigma / conf come from K9.
__global__ void __launch_bounds__(128) predict_points_kernel(const double* __restrict__ alpha, const double* __restrict__ b1,
                                                             const double* __restrict__ b2, int N, double p0, double cl,
                                                             const double* __restrict__ X, int64_t m, double* __restrict__ f) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const double x1 = X[2 * t], x2 = X[2 * t + 1];
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < N; i++) a[i & 3] = fma(alpha[i], rbf(x1, x2, b1[i], b2[i], p0, cl), a[i & 3]);
    f[t] = __dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3]));
}

}  // namespace

template <int R, int G, int MINB>
static cudaError_t launch_grid_variant(const PredictArgs& a, size_t smem, cudaStream_t s) {
    const int per_cta = PRED_T / G;
    const int64_t grid = (a.n_groups + per_cta - 1) / per_cta;
    if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
    if (a.separable) {
        cudaError_t e = GPC_FUNC_ATTR_ONCE((predict_grid_kernel<R, G, MINB, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        predict_grid_kernel<R, G, MINB, false><<<(unsigned)grid, PRED_T, smem, s>>>(a);
    } else {
        cudaError_t e = GPC_FUNC_ATTR_ONCE((predict_grid_kernel<R, G, MINB, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        predict_grid_kernel<R, G, MINB, true><<<(unsigned)grid, PRED_T, smem, s>>>(a);
    }
    return cudaGetLastError();
}

// Picks the thread-group shape and the rows-per-block split so that the tables of nmax (+ nrmax) basis vectors fit in
// shared memory; returns cudaErrorInvalidConfiguration if even one row block does not fit.
cudaError_t launch_predict_grid(const PredictArgs& a0, cudaStream_t s) {
    if (a0.n_patches <= 0) return cudaSuccess;
    PredictArgs a = a0;
    const int nb = a.nmax + a.nrmax;
    const bool small = a.sz <= 16;           // warp per patch, four patches per CTA
    const int R = (a.sz > 32) ? 8 : 1;
    const int per_cta = small ? PRED_T / 32 : 1;
    const int64_t budget = (int64_t)(small ? 48 : 220) * 1024 / (int64_t)sizeof(double) / per_cta;
    // doubles per block: frame[16] + Xs[sz] + Ys[rowsP] + (alpha, b1, b2)[nmax] + (3 alpha, b1, b2)[nrmax] + nb (sz + rowsP)
    const int64_t fixed = 16 + (int64_t)a.sz + 3 * (int64_t)a.nmax + 5 * (int64_t)a.nrmax + (int64_t)nb * a.sz;
    int64_t rmax = (budget - fixed) / (nb + 1);
    if (small && rmax < a.sz) {  // tables too large for four patches per CTA: use the CTA-wide shape
        PredictArgs b = a0;
        b.sz = a0.sz;
        const int64_t budget2 = (int64_t)220 * 1024 / (int64_t)sizeof(double);
        rmax = (budget2 - fixed) / (nb + 1);
        if (rmax < 1) return cudaErrorInvalidConfiguration;
        a.rows = (int)std::min<int64_t>(rmax, a.sz);
        a.rowsP = a.rows;
        a.nblk = (a.sz + a.rows - 1) / a.rows;
        a.group_doubles = (int)(fixed + (int64_t)(nb + 1) * a.rowsP);
        a.n_groups = a.n_patches * a.nblk;
        g_launches++;
        return launch_grid_variant<1, PRED_T, 8>(a, (size_t)a.group_doubles * sizeof(double), s);
    }
    if (rmax < 1) return cudaErrorInvalidConfiguration;
    int rows = (int)std::min<int64_t>(rmax, a.sz);
    if (R > 1 && rows >= R) rows -= rows % R;
    a.rows = rows;
    a.nblk = (a.sz + rows - 1) / rows;
    a.rowsP = (rows + R - 1) / R * R;
    a.group_doubles = (int)(fixed + (int64_t)(nb + 1) * a.rowsP);
    a.group_doubles += a.group_doubles & 1;  // keep every block 16-byte aligned
    a.n_groups = a.n_patches * a.nblk;
    const size_t smem = (size_t)a.group_doubles * per_cta * sizeof(double);
    g_launches++;
    if (small) return launch_grid_variant<1, 32, 8>(a, smem, s);
    if (R == 8) return smem <= 56 * 1024 ? launch_grid_variant<8, PRED_T, 4>(a, smem, s) : launch_grid_variant<8, PRED_T, 3>(a, smem, s);
    return launch_grid_variant<1, PRED_T, 8>(a, smem, s);
}

void launch_predict_points(const double* alpha, const double* b1, const double* b2, int N, double p0, double cl, const double* X,
                           int64_t m, double* f, cudaStream_t s) {
    if (m <= 0) return;
    predict_points_kernel<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(alpha, b1, b2, N, p0, cl, X, m, f);
}

}  // namespace gpc
