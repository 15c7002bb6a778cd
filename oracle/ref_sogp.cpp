// oracle/ref_sogp.cpp — TEST INFRASTRUCTURE.  Harness around THE REFERENCE'S OWN sparse_gp.hpp,
// rbf_kernel.cpp and gaussian_noise.cpp, compiled from where they lie under /root/reference/src (never
// copied) against oracle/eigen_shim.  It exposes the reference's SOGP (add_measurements incl. its rand()
// shuffle, predict_measurements) and the private state (alpha, BV, C, Q) to the tests and to
// tests/golden/make_golden.py.  Private members are reached with the usual test-harness trick.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <sstream>
#include <vector>

#include <Eigen/Dense>

#include "octave_convenience.h"  // std headers first: the access trick below must not reach them
#define private public
#define protected public
#include "sparse_gp.h"
#include "sparse_gp_field.h"
#undef private
#undef protected
#include "gaussian_noise.h"
#include "gaussian_noise_3d.h"
#include "rbf_kernel.h"

typedef sparse_gp<rbf_kernel, gaussian_noise> ref_gp;
typedef sparse_gp_field<rbf_kernel, gaussian_noise_3d> ref_gp_field;

extern "C" {

// Fits one process.  rand_offset: number of rand() draws consumed before this call (srand(1) = the unseeded
// stream).  Returns current_size; fills alpha/bv (2 x N, column j = BV j)/C/Q (N x N row-major) up to max_n.
int ref_sogp_fit(int n, const double* x1, const double* x2, const double* y, int capacity, double s0, double sigmaf_sq,
                 double l_sq, double eps_tol, unsigned long long rand_offset, int max_n, double* alpha, double* bv1, double* bv2,
                 double* C, double* Q, int n_pred, const double* px1, const double* px2, double* f_star, double* sigma) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp gp(capacity, s0);
    gp.kernel.param()(0) = sigmaf_sq;
    gp.kernel.param()(1) = l_sq;
    gp.eps_tol = eps_tol;
    Eigen::MatrixXd X(n, 2);
    Eigen::VectorXd Y(n);
    for (int i = 0; i < n; i++) { X(i, 0) = x1[i]; X(i, 1) = x2[i]; Y(i) = y[i]; }
    gp.add_measurements(X, Y);
    const int N = gp.size();
    if (N > max_n) return -N;
    for (int i = 0; i < N; i++) {
        alpha[i] = gp.alpha(i);
        bv1[i] = gp.BV(0, i);
        bv2[i] = gp.BV(1, i);
        for (int j = 0; j < N; j++) { C[i * N + j] = gp.C(i, j); Q[i * N + j] = gp.Q(i, j); }
    }
    if (n_pred > 0) {
        Eigen::MatrixXd Xs(n_pred, 2);
        for (int i = 0; i < n_pred; i++) { Xs(i, 0) = px1[i]; Xs(i, 1) = px2[i]; }
        Eigen::VectorXd f, sg;
        gp.predict_measurements(f, Xs, sg);
        for (int i = 0; i < n_pred; i++) { f_star[i] = f(i); sigma[i] = sg(i); }
    }
    return N;
}

// The RGB field GP of the reference (sparse_gp_field.hpp): Y is n x 3 row-major.  alpha out: 3 per BV.
int ref_field_fit(int n, const double* x1, const double* x2, const double* Y, int capacity, double s0, double sigmaf_sq,
                  double l_sq, double eps_tol, unsigned long long rand_offset, int max_n, double* alpha3, double* bv1, double* bv2,
                  int n_pred, const double* px1, const double* px2, double* f_star3) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp_field gp(capacity, s0);
    gp.kernel.param()(0) = sigmaf_sq;
    gp.kernel.param()(1) = l_sq;
    gp.eps_tol = eps_tol;
    Eigen::MatrixXd X(n, 2), C(n, 3);
    for (int i = 0; i < n; i++) { X(i, 0) = x1[i]; X(i, 1) = x2[i]; for (int c = 0; c < 3; c++) C(i, c) = Y[3 * i + c]; }
    gp.add_measurements(X, C);
    const int N = gp.size();
    if (N > max_n) return -N;
    for (int i = 0; i < N; i++) {
        for (int c = 0; c < 3; c++) alpha3[3 * i + c] = gp.alpha(i, c);
        bv1[i] = gp.BV(0, i);
        bv2[i] = gp.BV(1, i);
    }
    if (n_pred > 0) {
        Eigen::MatrixXd Xs(n_pred, 2), F;
        for (int i = 0; i < n_pred; i++) { Xs(i, 0) = px1[i]; Xs(i, 1) = px2[i]; }
        Eigen::VectorXd sg;
        gp.predict_measurements(F, Xs, sg);
        for (int i = 0; i < n_pred; i++) for (int c = 0; c < 3; c++) f_star3[3 * i + c] = F(i, c);
    }
    return N;
}

// Fit, then the reference's own predict_measurements (sigma and conf modes, sparse_gp.hpp:299-351),
// compute_likelihoods (:414-420) and compute_derivatives (:459-468) at the given points.
int ref_sogp_evaluate(int n, const double* x1, const double* x2, const double* y, int capacity, double s0, double sigmaf_sq,
                      double l_sq, double eps_tol, unsigned long long rand_offset, int m, const double* ex1, const double* ex2,
                      const double* ey, double* f, double* sigma, double* conf, double* lik, double* dX) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp gp(capacity, s0);
    gp.kernel.param()(0) = sigmaf_sq;
    gp.kernel.param()(1) = l_sq;
    gp.eps_tol = eps_tol;
    Eigen::MatrixXd X(n, 2);
    Eigen::VectorXd Y(n);
    for (int i = 0; i < n; i++) { X(i, 0) = x1[i]; X(i, 1) = x2[i]; Y(i) = y[i]; }
    if (n > 0) gp.add_measurements(X, Y);
    Eigen::MatrixXd Xs(m, 2);
    Eigen::VectorXd Ys(m);
    for (int i = 0; i < m; i++) { Xs(i, 0) = ex1[i]; Xs(i, 1) = ex2[i]; Ys(i) = ey[i]; }
    Eigen::VectorXd fs, sg, cf, l;
    Eigen::MatrixXd D;
    gp.predict_measurements(fs, Xs, sg, false);
    gp.predict_measurements(fs, Xs, cf, true);
    gp.compute_likelihoods(l, Xs, Ys);
    gp.compute_derivatives(D, Xs, Ys);
    for (int i = 0; i < m; i++) {
        f[i] = fs(i); sigma[i] = sg(i); conf[i] = cf(i); lik[i] = l(i);
        for (int c = 0; c < 3; c++) dX[3 * i + c] = D(i, c);
    }
    return gp.size();
}

// add_measurements called twice on one process (the reference accumulates, sparse_gp.hpp:59-86): state after both
int ref_sogp_fit_twice(int n1, int n2, const double* x1, const double* x2, const double* y, int capacity, double s0, double sigmaf_sq,
                       double l_sq, double eps_tol, unsigned long long rand_offset, int max_n, double* alpha, double* bv1, double* bv2,
                       double* C, double* Q) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp gp(capacity, s0);
    gp.kernel.param()(0) = sigmaf_sq;
    gp.kernel.param()(1) = l_sq;
    gp.eps_tol = eps_tol;
    int done = 0;
    for (int part = 0; part < 2; part++) {
        const int n = part ? n2 : n1;
        if (n > 0) {
            Eigen::MatrixXd X(n, 2);
            Eigen::VectorXd Y(n);
            for (int i = 0; i < n; i++) { X(i, 0) = x1[done + i]; X(i, 1) = x2[done + i]; Y(i) = y[done + i]; }
            gp.add_measurements(X, Y);
        }
        done += n;
    }
    const int N = gp.size();
    if (N > max_n) return -N;
    for (int i = 0; i < N; i++) {
        alpha[i] = gp.alpha(i);
        bv1[i] = gp.BV(0, i);
        bv2[i] = gp.BV(1, i);
        for (int j = 0; j < N; j++) { C[i * N + j] = gp.C(i, j); Q[i * N + j] = gp.Q(i, j); }
    }
    return N;
}

// Field GP: fit, then the reference's own predict_measurements (sigma and conf), compute_likelihoods and
// compute_derivatives (sparse_gp_field.hpp:267-393) at the given points.  Y / EY: 3 per point, row-major.
int ref_field_evaluate(int n, const double* x1, const double* x2, const double* Y, int capacity, double s0, double sigmaf_sq,
                       double l_sq, double eps_tol, unsigned long long rand_offset, int m, const double* ex1, const double* ex2,
                       const double* EY, double* f3, double* sigma, double* conf, double* lik, double* dX) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp_field gp(capacity, s0);
    gp.kernel.param()(0) = sigmaf_sq;
    gp.kernel.param()(1) = l_sq;
    gp.eps_tol = eps_tol;
    Eigen::MatrixXd X(n, 2), C(n, 3);
    for (int i = 0; i < n; i++) { X(i, 0) = x1[i]; X(i, 1) = x2[i]; for (int c = 0; c < 3; c++) C(i, c) = Y[3 * i + c]; }
    if (n > 0) gp.add_measurements(X, C);
    Eigen::MatrixXd Xs(m, 2), Ys(m, 3), F, D;
    for (int i = 0; i < m; i++) { Xs(i, 0) = ex1[i]; Xs(i, 1) = ex2[i]; for (int c = 0; c < 3; c++) Ys(i, c) = EY[3 * i + c]; }
    Eigen::VectorXd sg, cf, l;
    gp.predict_measurements(F, Xs, sg, false);
    gp.predict_measurements(F, Xs, cf, true);
    gp.compute_likelihoods(l, Xs, Ys);
    gp.compute_derivatives(D, Xs, Ys);
    for (int i = 0; i < m; i++) {
        sigma[i] = sg(i); conf[i] = cf(i); lik[i] = l(i);
        for (int c = 0; c < 3; c++) { f3[3 * i + c] = F(i, c); dX[3 * i + c] = D(i, c); }
    }
    return gp.size();
}

// sparse_gp::shuffle alone (sparse_gp.hpp:42-56) over the real rand()
void ref_shuffle(int n, unsigned long long rand_offset, int* out) {
    srand(1);
    for (unsigned long long i = 0; i < rand_offset; i++) rand();
    ref_gp gp(10, 0.1);
    std::vector<int> ind;
    gp.shuffle(ind, n);
    for (int i = 0; i < n; i++) out[i] = ind[i];
}

double ref_kernel(double sigmaf_sq, double l_sq, double a1, double a2, double b1, double b2) {
    rbf_kernel k(sigmaf_sq, l_sq);
    Eigen::Vector2d a, b;
    a(0) = a1; a(1) = a2; b(0) = b1; b(1) = b2;
    return k.kernel_function(a, b);
}

}  // extern "C"
