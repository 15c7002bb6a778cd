// k_evaluate.cu — K9: batched GP evaluation with uncertainty, likelihood and likelihood gradient (next rows N2 / N4).
//
// Replaces, for every fitted patch and a ragged set of query points per patch:
//   sparse_gp::predict_measurements with sigma / conf   /root/reference/src/sparse_gp.hpp:299-351
//   sparse_gp::compute_likelihoods -> likelihood          sparse_gp.hpp:407-425
//   sparse_gp::compute_derivatives -> likelihood_dx       sparse_gp.hpp:459-502 (+ rbf_kernel::kernel_dx, rbf_kernel.cpp:38-46)
// which the reference runs one point at a time (gp_registration.cpp:175-194).
//
// One CTA per patch, tiles of 32 query points.  Per tile: (1) K = p0 E, E = exp(cl |x - BV_i|^2) for the N x 32
// pairs; (2) CK = C K as a register-tiled product (each thread: 4 rows x 4 points, or 1 x 4 for small N), C read
// by columns from shared memory (from global memory when N^2 does not fit); (3) the six O(N) sums per point
// (k'Ck, kdx'Ck, kdy'Ck, alpha'k, alpha'kdx, alpha'kdy) as canonical row4 dots, one warp per pair of sums;
// (4) the scalar epilogue of predict / likelihood / likelihood_dx.  Same operation order as oracle Sogp::evaluate.
#include <algorithm>

#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace gpc {

namespace {

constexpr int EV_T = 32;     // points per tile
constexpr int EV_NT = 128;   // threads per CTA

// DOUT = 1: sparse_gp (heights).  DOUT = 3: sparse_gp_field (RGB), sparse_gp_field.hpp:267-393: three alpha columns, the
// residual is a 3-vector (squaredNorm in the exponent, pow(2 pi, 3) in the density), dX[0] = 0.
template <int RPT, bool CSMEM, int DOUT>
__global__ void __launch_bounds__(EV_NT) evaluate_kernel(EvalArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int64_t p = blockIdx.x;
    const int N = a.nbv[p];
    const int64_t o = a.off[p];
    const int n = (int)(a.off[p + 1] - o);
    if (n == 0) return;
    const int t = threadIdx.x;
    const int LDC = (N + 3) & ~3;
    // layout: K, E, CK [N x 32] | red [(3 + 3 DOUT) x 32] | xs, ys [32], yv [DOUT x 32] | al [DOUT x N], b1, b2 [N] | C [N x LDC]
    double* K = sm;
    double* E = K + (size_t)N * EV_T;
    double* CK = E + (size_t)N * EV_T;
    double* red = CK + (size_t)N * EV_T;
    constexpr int NRED = 3 + 3 * DOUT;
    double* xs = red + NRED * EV_T;
    double* ys = xs + EV_T;
    double* yv = ys + EV_T;
    double* al = yv + DOUT * EV_T;
    double* b1 = al + DOUT * N;
    double* b2 = b1 + N;
    double* Cs = b2 + N + (((DOUT + 2) * N + NRED * EV_T + DOUT * EV_T) & 1);  // 16-byte aligned
    const int64_t pb = p * a.stride;
    const double* Cg = a.C + p * (int64_t)a.stride * a.stride;  // packed N x N
    for (int i = t; i < N; i += EV_NT) {
#pragma unroll
        for (int c = 0; c < DOUT; c++) al[c * N + i] = a.alpha[c][pb + i];
        b1[i] = a.b1[pb + i]; b2[i] = a.b2[pb + i];
    }
    const bool have_c = a.C != nullptr;   // mean only (no sigma / likelihood / gradient): C is neither needed nor read
    if (CSMEM && have_c) {
        for (int e = t; e < N * LDC; e += EV_NT) {
            const int j = e / LDC, i = e - j * LDC;
            Cs[e] = (i < N) ? Cg[(size_t)j * N + i] : 0.0;
        }
    }
    const double kstar = a.p0, s20 = a.s20, c1 = a.c1;
    for (int tile = 0; tile < n; tile += EV_T) {
        __syncthreads();
        if (t < EV_T) {
            const int q = tile + t;
            xs[t] = (q < n) ? a.x1[o + q] : 0.0;
            ys[t] = (q < n) ? a.x2[o + q] : 0.0;
#pragma unroll
            for (int c = 0; c < DOUT; c++) yv[c * EV_T + t] = (q < n && a.y) ? a.y[(o + q) * DOUT + c] : 0.0;
        }
        __syncthreads();
        // (1) kernel values
        for (int e = t; e < N * EV_T; e += EV_NT) {
            const int i = e / EV_T, tt = e - i * EV_T;
            const double d1 = __dadd_rn(xs[tt], -b1[i]), d2 = __dadd_rn(ys[tt], -b2[i]);
            const double ex = gpc_exp_nonpos(__dmul_rn(a.cl, __dadd_rn(__dmul_rn(d1, d1), __dmul_rn(d2, d2))));
            E[e] = ex;
            K[e] = __dmul_rn(a.p0, ex);
        }
        __syncthreads();
        // (2) CK(i, tt) = sum_j C(j, i) K(j, tt), j ascending, one fma chain per entry
        if (have_c) {
            const int pg = t & 7, rg = t >> 3;   // 8 point groups of 4, 16 row groups
            for (int i0 = rg * RPT; i0 < N; i0 += 16 * RPT) {
                double acc[RPT][4];
#pragma unroll
                for (int r = 0; r < RPT; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0;
                for (int j = 0; j < N; j++) {
                    const double2 k01 = *reinterpret_cast<const double2*>(K + j * EV_T + 4 * pg);
                    const double2 k23 = *reinterpret_cast<const double2*>(K + j * EV_T + 4 * pg + 2);
                    double cv[RPT];
                    if (CSMEM) {
                        if (RPT == 4) {
                            const double2 c01 = *reinterpret_cast<const double2*>(Cs + j * LDC + i0);
                            const double2 c23 = *reinterpret_cast<const double2*>(Cs + j * LDC + i0 + 2);
                            cv[0] = c01.x; cv[RPT > 1 ? 1 : 0] = c01.y; cv[RPT > 2 ? 2 : 0] = c23.x; cv[RPT > 3 ? 3 : 0] = c23.y;
                        } else {
#pragma unroll
                            for (int r = 0; r < RPT; r++) cv[r] = Cs[j * LDC + i0 + r];
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < RPT; r++) cv[r] = (i0 + r < N) ? __ldg(Cg + (size_t)j * N + i0 + r) : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        acc[r][0] = fma(cv[r], k01.x, acc[r][0]);
                        acc[r][1] = fma(cv[r], k01.y, acc[r][1]);
                        acc[r][2] = fma(cv[r], k23.x, acc[r][2]);
                        acc[r][3] = fma(cv[r], k23.y, acc[r][3]);
                    }
                }
#pragma unroll
                for (int r = 0; r < RPT; r++)
                    if (i0 + r < N) {
                        *reinterpret_cast<double2*>(CK + (i0 + r) * EV_T + 4 * pg) = make_double2(acc[r][0], acc[r][1]);
                        *reinterpret_cast<double2*>(CK + (i0 + r) * EV_T + 4 * pg + 2) = make_double2(acc[r][2], acc[r][3]);
                    }
            }
        }
        __syncthreads();
        // (3) the six sums: warps 0-2 take a pair each (k'Ck + alpha'k, kdx'Ck + alpha'kdx, kdy'Ck + alpha'kdy), sharing
        // the loads of their first factor; every sum is its own canonical row4
        {
            const int w = t >> 5, tt = t & 31;
            const double x1 = xs[tt], x2 = ys[tt];
            if (w < 3) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                double bb[DOUT][4];
#pragma unroll
                for (int c = 0; c < DOUT; c++) bb[c][0] = bb[c][1] = bb[c][2] = bb[c][3] = 0.0;
                auto fac = [&](int i) {
                    if (w == 0) return K[i * EV_T + tt];
                    const double d = (w == 1) ? __dadd_rn(x1, -b1[i]) : __dadd_rn(x2, -b2[i]);
                    return __dmul_rn(__dmul_rn(c1, d), E[i * EV_T + tt]);
                };
                int i = 0;
                for (; i + 3 < N; i += 4) {
                    const double f0 = fac(i), f1 = fac(i + 1), f2 = fac(i + 2), f3 = fac(i + 3);
                    a0 = fma(f0, CK[i * EV_T + tt], a0); a1 = fma(f1, CK[(i + 1) * EV_T + tt], a1);
                    a2 = fma(f2, CK[(i + 2) * EV_T + tt], a2); a3 = fma(f3, CK[(i + 3) * EV_T + tt], a3);
#pragma unroll
                    for (int c = 0; c < DOUT; c++) {
                        const double* ac = al + c * N;
                        bb[c][0] = fma(ac[i], f0, bb[c][0]); bb[c][1] = fma(ac[i + 1], f1, bb[c][1]);
                        bb[c][2] = fma(ac[i + 2], f2, bb[c][2]); bb[c][3] = fma(ac[i + 3], f3, bb[c][3]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 3; u++)
                    if (i + u < N) {
                        const double f = fac(i + u);
                        if (u == 0) a0 = fma(f, CK[(i + u) * EV_T + tt], a0);
                        if (u == 1) a1 = fma(f, CK[(i + u) * EV_T + tt], a1);
                        if (u == 2) a2 = fma(f, CK[(i + u) * EV_T + tt], a2);
#pragma unroll
                        for (int c = 0; c < DOUT; c++) bb[c][u] = fma(al[c * N + i + u], f, bb[c][u]);
                    }
                // red rows: 0..2 = k'Ck, kdx'Ck, kdy'Ck; 3 + 3 c + w = alpha_c' (k | kdx | kdy)
                red[w * EV_T + tt] = __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
#pragma unroll
                for (int c = 0; c < DOUT; c++)
                    red[(3 + 3 * c + w) * EV_T + tt] = __dadd_rn(__dadd_rn(bb[c][0], bb[c][1]), __dadd_rn(bb[c][2], bb[c][3]));
            }
        }
        __syncthreads();
        // (4) epilogue
        if (t < EV_T && tile + t < n) {
            const int64_t q = o + tile + t;
            const double kCk = red[t], sx = red[EV_T + t], sy = red[2 * EV_T + t];
            double mu[DOUT], ax[DOUT], ay[DOUT], off[DOUT];
#pragma unroll
            for (int c = 0; c < DOUT; c++) {
                mu[c] = red[(3 + 3 * c) * EV_T + t]; ax[c] = red[(4 + 3 * c) * EV_T + t]; ay[c] = red[(5 + 3 * c) * EV_T + t];
                off[c] = __dadd_rn(yv[c * EV_T + t], -mu[c]);
            }
            // predict, sparse_gp.hpp:329-349 / sparse_gp_field.hpp:296-318
            double var = __dadd_rn(__dadd_rn(s20, kstar), kCk);
            const double var_l = var;
            if (var < 0.0) var = 0.0;
            if (a.f) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) a.f[q * DOUT + c] = mu[c];
            }
            if (a.sigma)
                a.sigma[q] = a.conf ? __dmul_rn(100.0, __dadd_rn(1.0, -__ddiv_rn(var, __dadd_rn(kstar, s20)))) : __dsqrt_rn(var);
            // the residual enters as off^2 (scalar: ((-0.5/var) off) off, :424 / :495) or as squaredNorm (field, :350 / :383)
            double sqn = 0.0;
            if (DOUT == 3) sqn = __dadd_rn(__dadd_rn(__dmul_rn(off[0], off[0]), __dmul_rn(off[DOUT > 1 ? 1 : 0], off[DOUT > 1 ? 1 : 0])),
                                           __dmul_rn(off[DOUT > 2 ? 2 : 0], off[DOUT > 2 ? 2 : 0]));
            if (a.lik) {
                const double arg = (DOUT == 1) ? __dmul_rn(__dmul_rn(__ddiv_rn(-0.5, var_l), off[0]), off[0])
                                               : __dmul_rn(__ddiv_rn(-0.5, var_l), sqn);
                const double norm = (DOUT == 1) ? 6.283185307179586 : a.pow2pi3;
                a.lik[q] = __dmul_rn(__ddiv_rn(1.0, __dsqrt_rn(__dmul_rn(norm, var_l))), gpc_exp(arg));
            }
            if (a.dX) {  // sparse_gp.hpp:487-499 / sparse_gp_field.hpp:378-390
                const double var_d = __dadd_rn(__dadd_rn(s20, kCk), kstar);
                const double sdx[2] = {__dmul_rn(2.0, sx), __dmul_rn(2.0, sy)};
                const double sq = __dsqrt_rn(var_d);
                const double vs = __dmul_rn(var_d, sq);
                const double arg = (DOUT == 1) ? __dmul_rn(__dmul_rn(__ddiv_rn(-0.5, var_d), off[0]), off[0])
                                               : __dmul_rn(__ddiv_rn(-0.5, var_d), sqn);
                const double exppart = __dmul_rn(__ddiv_rn(0.5, vs), gpc_exp(arg));
                a.dX[3 * q] = (DOUT == 1) ? __dmul_rn(__dmul_rn(__ddiv_rn(-1.0, vs), off[0]), exppart) : 0.0;
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    const double* A = d ? ay : ax;
                    const double first = -sdx[d];
                    double second, third;
                    if (DOUT == 1) {
                        second = __dmul_rn(__dmul_rn(2.0, A[0]), off[0]);
                        third = __dmul_rn(__dmul_rn(__ddiv_rn(sdx[d], var_d), off[0]), off[0]);
                    } else {
                        second = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(A[0], off[0]), __dmul_rn(A[DOUT > 1 ? 1 : 0], off[DOUT > 1 ? 1 : 0])),
                                                          __dmul_rn(A[DOUT > 2 ? 2 : 0], off[DOUT > 2 ? 2 : 0])));
                        third = __dmul_rn(__ddiv_rn(sdx[d], var_d), sqn);
                    }
                    a.dX[3 * q + 1 + d] = __dmul_rn(exppart, __dadd_rn(__dadd_rn(first, second), third));
                }
            }
        }
    }
}

template <int RPT, bool CSMEM, int DOUT>
cudaError_t launch_variant(const EvalArgs& a, size_t smem, cudaStream_t s) {
    cudaError_t e = GPC_FUNC_ATTR_ONCE((evaluate_kernel<RPT, CSMEM, DOUT>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    evaluate_kernel<RPT, CSMEM, DOUT><<<(unsigned)a.n_patches, EV_NT, smem, s>>>(a);
    return cudaGetLastError();
}

template <int DOUT>
cudaError_t launch_evaluate_d(const EvalArgs& a, cudaStream_t s) {
    const int64_t N = std::max(a.nmax, 1);
    const int64_t LDC = (N + 3) & ~(int64_t)3;
    const int64_t base = 3 * N * EV_T + (3 + 3 * DOUT) * EV_T + (2 + DOUT) * EV_T + (DOUT + 2) * N + 1;
    const int64_t with_c = base + N * LDC;
    const int64_t budget = 220 * 1024 / (int64_t)sizeof(double);
    if (with_c <= budget) {
        const size_t smem = (size_t)with_c * sizeof(double);
        return N > 32 ? launch_variant<4, true, DOUT>(a, smem, s) : launch_variant<1, true, DOUT>(a, smem, s);
    }
    if (base > budget) return cudaErrorInvalidConfiguration;
    return launch_variant<4, false, DOUT>(a, (size_t)base * sizeof(double), s);
}

}  // namespace

cudaError_t launch_evaluate(const EvalArgs& a, cudaStream_t s) {
    if (a.n_patches <= 0) return cudaSuccess;
    g_launches++;
    return a.dout == 3 ? launch_evaluate_d<3>(a, s) : launch_evaluate_d<1>(a, s);
}

}  // namespace gpc
