// gpc_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A dependency-free C++17 restatement of the gp_compressor compress/decompress hot
// path of nilsbore/gp_compressor (reference at /root/reference, read-only).  It is the
// checker for the CUDA product in gp_compressor_b200/: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product never
// links, imports or calls anything in this directory.
//
// PARITY STATUS: "parity unpinned" against a real PCL+Eigen build of the reference.
// The reference cannot be compiled here (PCL, Eigen, Boost absent; no network) and
// ships no golden vectors, tests or data.  What IS pinned (tests/test_oracle_*.py):
//   * glibc rand() stream and sparse_gp::shuffle     — against the real libc in this image
//   * canonical exp                                  — against libm exp, <= 1 ulp
//   * the SOGP recursion                             — against an independent numpy
//     restatement that follows matlab/sogp.m (second witness shipped by the reference)
//     and, when built, against the reference's own sparse_gp.hpp compiled over a
//     minimal Eigen shim (oracle/_ref)
//   * PCL octree semantics                           — [RECALLED PCL 1.7], listed below
// Arithmetic that lives in un-vendored third-party code (Eigen's summation order, the
// JacobiSVD rotation sequence, glibc's exp table) is replaced by the CANONICAL forms
// defined in this file; the CUDA kernels implement exactly the same forms, so the
// oracle<->GPU comparison is bit-exact while oracle<->real-reference agreement is
// limited to rounding-level differences.
//
// Reference map (file:line under /root/reference/src):
//   lattice / keys / leaf order / radius search : gp_compressor.cpp:179-182,204-207,220
//                                                 gp_octree.cpp:3-11,138-140  [+PCL 1.7]
//   compute_rotation                            : gp_compressor.cpp:29-64
//   project_points (claim + local frame)        : gp_compressor.cpp:66-118
//   project_cloud / train_processes drivers     : gp_compressor.cpp:121-249
//   shuffle / add_measurements                  : sparse_gp.hpp:42-86
//   add                                         : sparse_gp.hpp:89-249
//   delete_bv                                   : sparse_gp.hpp:252-295
//   predict / predict_measurements              : sparse_gp.hpp:299-351
//   construct_covariance / rbf kernel           : sparse_gp.hpp:522-530, rbf_kernel.cpp:15-18
//   gaussian noise derivatives                  : gaussian_noise.cpp:9-18
//   load_compressed (grid + reprojection)       : gp_compressor.cpp:267-386
//   RGB field GP shuffle (rand accounting only) : sparse_gp_field.hpp:29-55
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -mfma -fPIC -shared (see oracle/Makefile).
// -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit fma().

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------
// Constants: the reference writes float literals that widen to double.
// ------------------------------------------------------------------------------------
const double TINY12 = (double)1e-12f;  // sparse_gp.hpp:124,146
const double GEO9 = (double)1e-9f;     // sparse_gp.hpp:229,236

// ------------------------------------------------------------------------------------
// Canonical exp (replaces glibc exp at rbf_kernel.cpp:17).  Pure IEEE double ops with
// explicit fma; <= 1 ulp from libm.  The CUDA side implements the same sequence.
// ------------------------------------------------------------------------------------
inline double orc_exp_impl(double x) {
    if (x != x) return x;
    if (x > 709.0) return std::numeric_limits<double>::infinity();
    if (x < -745.0) return 0.0;
    const double INV_LN2 = 1.4426950408889634;
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52 : round-to-nearest-even integer
    const double LN2_HI = 0x1.62e42fefa39efp-1;
    const double LN2_LO = 0x1.abc9e3b39803fp-56;
    double t = x * INV_LN2;
    double kd = t + MAGIC;
    uint64_t kb;
    std::memcpy(&kb, &kd, 8);
    int n = (int)(uint32_t)kb;
    kd = kd - MAGIC;
    double r = std::fma(kd, -LN2_HI, x);
    r = std::fma(kd, -LN2_LO, r);
    // Taylor coefficients 1/k!, k = 13..0, Horner with fma.
    double p = 1.0 / 6227020800.0;
    p = std::fma(p, r, 1.0 / 479001600.0);
    p = std::fma(p, r, 1.0 / 39916800.0);
    p = std::fma(p, r, 1.0 / 3628800.0);
    p = std::fma(p, r, 1.0 / 362880.0);
    p = std::fma(p, r, 1.0 / 40320.0);
    p = std::fma(p, r, 1.0 / 5040.0);
    p = std::fma(p, r, 1.0 / 720.0);
    p = std::fma(p, r, 1.0 / 120.0);
    p = std::fma(p, r, 1.0 / 24.0);
    p = std::fma(p, r, 1.0 / 6.0);
    p = std::fma(p, r, 0.5);
    p = std::fma(p, r, 1.0);
    p = std::fma(p, r, 1.0);
    int adj = 0;
    if (n < -1020) { adj = 1; n += 1020; }
    uint64_t pb;
    std::memcpy(&pb, &p, 8);
    pb += (uint64_t)((int64_t)n << 52);
    std::memcpy(&p, &pb, 8);
    if (adj) p = p * 0x1p-1020;
    return p;
}

// ------------------------------------------------------------------------------------
// glibc rand(): TYPE_3 additive feedback generator, unseeded (= srand(1)).
// [RECALLED glibc random_r.c]; pinned against the real libc in tests.
// Output k (k = 0,1,...) is r[344+k] >> 1 with r[i] = r[i-31] + r[i-3] (mod 2^32).
// ------------------------------------------------------------------------------------
struct GlibcRand {
    uint32_t w[34];  // sliding window of the last 34 r values; w[33] is r[pos-1]
    uint64_t pos;    // index of the next r to produce
    GlibcRand() { reset(); }
    void reset() {
        uint32_t r[344];
        r[0] = 1;
        for (int i = 1; i < 31; i++) {
            int64_t v = (16807LL * (int32_t)r[i - 1]) % 2147483647;
            if (v < 0) v += 2147483647;
            r[i] = (uint32_t)v;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
        for (int i = 0; i < 34; i++) w[i] = r[310 + i];
        pos = 344;
    }
    uint32_t next() {
        uint32_t v = w[34 - 31] + w[34 - 3];
        for (int i = 0; i < 33; i++) w[i] = w[i + 1];
        w[33] = v;
        pos++;
        return v >> 1;
    }
};

// ------------------------------------------------------------------------------------
// Canonical reductions.
//   dot32 : 32 lane-strided fma partial sums + 5-step xor butterfly (a warp can do it;
//           addition is commutative so every lane ends with the same bits).
//   row4  : 4 strided fma partial sums, (a0+a1)+(a2+a3)  (one matvec row / predict dot).
//   sum32 : like dot32 for plain sums.
// These replace Eigen's (unknowable) packet order for alpha'*k, k'*C*k, k'*e_hat, C*k, Q*k.
// ------------------------------------------------------------------------------------
inline double butterfly32(double* p) {
    for (int off = 16; off >= 1; off >>= 1) {
        double t[32];
        for (int l = 0; l < 32; l++) t[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; l++) p[l] = t[l];
    }
    return p[0];
}
inline double dot32(const double* a, const double* b, int n) {
    double p[32];
    for (int l = 0; l < 32; l++) p[l] = 0.0;
    for (int j = 0; j < n; j++) p[j & 31] = std::fma(a[j], b[j], p[j & 31]);
    return butterfly32(p);
}
inline double row4(const double* row, const double* k, int n) {
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < n; j++) a[j & 3] = std::fma(row[j], k[j], a[j & 3]);
    return (a[0] + a[1]) + (a[2] + a[3]);
}

struct SogpParams {
    int capacity;
    double s20, eps_tol, p0, p1, cl;  // cl = -0.5f / p1
    int ref_order = 0;  // 1: the height GP evaluates every expression in the order of the reference source over
                        // oracle/eigen_shim (sequential sums, libm exp, per-element divisions): bit-equal to oracle/_ref
};

struct SogpStats {
    uint64_t n_add = 0, n_first = 0, n_sparse = 0, n_full = 0, n_del_cap = 0, n_del_geo = 0;
    double sumN2_common = 0, sumN2_sparse = 0, sumN2_full = 0, sumN2_del = 0, sumN = 0;
    void merge(const SogpStats& o) {
        n_add += o.n_add; n_first += o.n_first; n_sparse += o.n_sparse; n_full += o.n_full;
        n_del_cap += o.n_del_cap; n_del_geo += o.n_del_geo;
        sumN2_common += o.sumN2_common; sumN2_sparse += o.sumN2_sparse;
        sumN2_full += o.sumN2_full; sumN2_del += o.sumN2_del; sumN += o.sumN;
    }
};

// One sparse online GP with D outputs (D = 1: sparse_gp, heights; D = 3: sparse_gp_field, RGB) and 2-D input.
// Dense symmetric C and Q with leading dimension ld = capacity + 2; alpha is stored per output channel.
// The recursion for C, Q and BV does not depend on the outputs; the two classes differ in alpha's width, in the
// capacity score (sparse_gp_field.hpp:187) and in delete_bv's alpha update (sparse_gp_field.hpp:250-253).
struct Sogp {
    SogpParams P;
    int N = 0, ld = 0, D = 1;
    bool field = false;
    std::vector<double> alpha, C, Q, b1, b2, k, ck, e, sv, qs, cs, qc;
    std::vector<int> idx;
    SogpStats st;
    int flags = 0;  // bit0: NaN in C(0,0) seen (sparse_gp.hpp:245-247)

    void init(const SogpParams& p, int maxn, int dout = 1) {
        P = p;
        D = dout;
        field = dout > 1;
        ld = maxn + 2;
        N = 0;
        st = SogpStats();
        flags = 0;
        alpha.assign((size_t)D * ld, 0.0);
        C.assign((size_t)ld * ld, 0.0);
        Q.assign((size_t)ld * ld, 0.0);
        b1.assign(ld, 0.0); b2.assign(ld, 0.0); k.assign(ld, 0.0); ck.assign(ld, 0.0);
        e.assign(ld, 0.0); sv.assign(ld, 0.0); qs.assign(ld, 0.0); cs.assign(ld, 0.0);
        qc.assign(ld, 0.0);
        idx.assign(ld, -1);
    }
    double& c(int i, int j) { return C[(size_t)i * ld + j]; }
    double& q(int i, int j) { return Q[(size_t)i * ld + j]; }
    double& al(int ch, int i) { return alpha[(size_t)ch * ld + i]; }

    // rbf_kernel.cpp:15-18 : p0 * exp(-0.5f/p1 * ||xi-xj||^2)
    inline double kern(double x1, double x2, double y1, double y2) const {
        double d1 = x1 - y1, d2 = x2 - y2;
        double sq = d1 * d1 + d2 * d2;
        return P.p0 * orc_exp_impl(P.cl * sq);
    }

    // sparse_gp.hpp:252-295 / sparse_gp_field.hpp:224-263
    void delete_bv(int loc) {
        const int L = N - 1, M = N - 1;
        st.sumN2_del += (double)M * M;
        double astar[3] = {0, 0, 0};
        for (int ch = 0; ch < D; ch++) { astar[ch] = al(ch, loc); al(ch, loc) = al(ch, L); }
        double cstar = c(loc, loc);
        for (int i = 0; i < N; i++) cs[i] = c(i, loc);
        cs[loc] = cs[L];
        // Crep = C.col(L); Crep(loc) = Crep(L); row/col loc <- Crep
        {
            std::vector<double>& rep = sv;  // reuse
            for (int i = 0; i < N; i++) rep[i] = c(i, L);
            rep[loc] = rep[L];
            for (int i = 0; i < N; i++) { c(loc, i) = rep[i]; c(i, loc) = rep[i]; }
        }
        double qstar = q(loc, loc);
        for (int i = 0; i < N; i++) qs[i] = q(i, loc);
        qs[loc] = qs[L];
        {
            std::vector<double>& rep = sv;
            for (int i = 0; i < N; i++) rep[i] = q(i, L);
            rep[loc] = rep[L];
            for (int i = 0; i < N; i++) { q(loc, i) = rep[i]; q(i, loc) = rep[i]; }
        }
        // Appendix G section g (sparse_gp.hpp:283-288); per-element divisions of the
        // reference are restated as multiplications by the rounded reciprocals iq, iqc.
        double qcs = qstar + cstar;
        for (int i = 0; i < M; i++) qc[i] = qs[i] + cs[i];
        if (!field) {
            double coef = astar[0] / qcs;   // alpha -= alphastar/(qstar + cstar)*(Qstar + Cstar)
            for (int i = 0; i < M; i++) al(0, i) = al(0, i) - coef * qc[i];
        } else {
            // sparse_gp_field.hpp:250-253: qc = (qstar + cstar)*(Qstar + Cstar) — multiplied, as the reference has it
            for (int i = 0; i < M; i++) {
                double w = qcs * qc[i];
                for (int ch = 0; ch < D; ch++) al(ch, i) = al(ch, i) - astar[ch] * w;
            }
        }
        double iq = 1.0 / qstar, iqc = 1.0 / qcs;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) {
                double u = qs[i] * qs[j];
                double v = qc[i] * qc[j];
                double t = v * iqc;
                double w = std::fma(u, iq, -t);
                c(i, j) = c(i, j) + w;
                q(i, j) = std::fma(-u, iq, q(i, j));
            }
        b1[loc] = b1[L]; b2[loc] = b2[L]; idx[loc] = idx[L];
        // clear the dropped row/col so padded reads stay zero
        for (int i = 0; i < N; i++) { c(L, i) = 0; c(i, L) = 0; q(L, i) = 0; q(i, L) = 0; }
        for (int ch = 0; ch < D; ch++) al(ch, L) = 0;
        b1[L] = 0; b2[L] = 0; idx[L] = -1;
        N = M;
    }

    // ================= reference-order arithmetic (P.ref_order, D == 1) =================
    // The same recursion, but every expression is evaluated exactly as the reference source does when it is compiled over
    // oracle/eigen_shim (the build behind oracle/_ref and tests/golden/ref_sogp.npz): Eigen products are plain sequential
    // sums s += a*b without fma, exp is libm's, delete_bv divides element by element, r*s*s' is (r s_i) s_j.  C and Q
    // are NOT symmetrised (k'C and C k are different sums).  The file:line of each statement is given.
    inline double kern_ref(double x1, double x2, double y1, double y2) const {  // rbf_kernel.cpp:15-18
        const double d1 = x1 - y1, d2 = x2 - y2;
        double sq = 0.0;
        sq += d1 * d1;
        sq += d2 * d2;
        return P.p0 * std::exp(P.cl * sq);
    }
    void delete_bv_ref(int loc) {  // sparse_gp.hpp:252-295
        const int L = N - 1, M = N - 1;
        st.sumN2_del += (double)M * M;
        const double alphastar = al(0, loc);                     // :256-258
        al(0, loc) = al(0, L);
        const double cstar = c(loc, loc);                        // :261
        for (int i = 0; i < N; i++) cs[i] = c(i, loc);           // :262 Cstar = C.col(loc)
        cs[loc] = cs[L];                                         // :263
        for (int i = 0; i < N; i++) sv[i] = c(i, L);             // :266 Crep = C.col(last)
        sv[loc] = sv[L];                                         // :267
        for (int j = 0; j < N; j++) c(loc, j) = sv[j];           // :268 C.row(loc) = Crep'
        for (int i = 0; i < N; i++) c(i, loc) = sv[i];           // :269 C.col(loc) = Crep
        const double qstar = q(loc, loc);                        // :273
        for (int i = 0; i < N; i++) qs[i] = q(i, loc);           // :274
        qs[loc] = qs[L];                                         // :275
        for (int i = 0; i < N; i++) sv[i] = q(i, L);             // :277
        sv[loc] = sv[L];                                         // :278
        for (int j = 0; j < N; j++) q(loc, j) = sv[j];           // :279
        for (int i = 0; i < N; i++) q(i, loc) = sv[i];           // :280
        const double qcs = qstar + cstar;
        const double coef = alphastar / qcs;                     // :285 alphastar/(qstar + cstar)*(Qstar + Cstar)
        for (int i = 0; i < M; i++) qc[i] = qs[i] + cs[i];
        for (int i = 0; i < M; i++) al(0, i) = al(0, i) - coef * qc[i];
        for (int i = 0; i < M; i++)                              // :286-288
            for (int j = 0; j < M; j++) {
                const double A = (0.0 + qs[i] * qs[j]) / qstar;
                const double B = (0.0 + qc[i] * qc[j]) / qcs;
                c(i, j) = c(i, j) + (A - B);
                q(i, j) = q(i, j) - A;
            }
        b1[loc] = b1[L]; b2[loc] = b2[L]; idx[loc] = idx[L];     // :291-292
        for (int i = 0; i < N; i++) { c(L, i) = 0; c(i, L) = 0; q(L, i) = 0; q(i, L) = 0; }
        al(0, L) = 0; b1[L] = 0; b2[L] = 0; idx[L] = -1;
        N = M;
    }
    void add_ref(double x1, double x2, double y, int orig) {  // sparse_gp.hpp:89-249
        st.n_add++;
        const double kstar = P.p0 * std::exp(P.cl * 0.0);        // :98
        if (N == 0) {                                            // :100-110
            al(0, 0) = y / (kstar + P.s20);
            c(0, 0) = -1.0 / (kstar + P.s20);
            q(0, 0) = 1.0 / kstar;
            b1[0] = x1; b2[0] = x2; idx[0] = orig;
            N = 1;
            st.n_first++;
            if (std::isnan(c(0, 0))) flags |= 1;
            return;
        }
        st.sumN += N;
        st.sumN2_common += (double)N * N;
        for (int i = 0; i < N; i++) k[i] = kern_ref(x1, x2, b1[i], b2[i]);   // :119
        double m = 0.0;
        for (int i = 0; i < N; i++) m += al(0, i) * k[i];                    // :121
        double kCk = 0.0;                                                    // :122 (k'C) k
        for (int j = 0; j < N; j++) {
            double t = 0.0;
            for (int i = 0; i < N; i++) t += k[i] * c(i, j);
            sv[j] = t;
        }
        for (int j = 0; j < N; j++) kCk += sv[j] * k[j];
        const double s2 = kstar + kCk;
        const double r = -1.0 / (P.s20 + s2);                                // :134, gaussian_noise.cpp:15-18
        const double qq = (y - m) / (P.s20 + s2);                            // :137, gaussian_noise.cpp:9-12
        for (int i = 0; i < N; i++) {                                        // :140 e_hat = Q k
            double t = 0.0;
            for (int j = 0; j < N; j++) t += q(i, j) * k[j];
            e[i] = t;
        }
        double ke = 0.0;
        for (int i = 0; i < N; i++) ke += k[i] * e[i];
        double gamma = kstar - ke;                                           // :144
        if (gamma < TINY12) gamma = 0;                                       // :146-151
        for (int i = 0; i < N; i++) {                                        // C k (:160 / :171)
            double t = 0.0;
            for (int j = 0; j < N; j++) t += c(i, j) * k[j];
            ck[i] = t;
        }
        if (gamma < P.eps_tol && P.capacity != -1) {                         // :155-163
            st.n_sparse++;
            st.sumN2_sparse += (double)N * N;
            const double eta = 1 / (1 + gamma * r);
            for (int i = 0; i < N; i++) sv[i] = ck[i] + e[i];
            const double qe = qq * eta;
            for (int i = 0; i < N; i++) al(0, i) = al(0, i) + sv[i] * qe;
            const double re = r * eta;
            for (int i = 0; i < N; i++) {
                const double t = re * sv[i];
                for (int j = 0; j < N; j++) c(i, j) = c(i, j) + (0.0 + t * sv[j]);
            }
        } else {                                                             // :164-203
            st.n_full++;
            st.sumN2_full += (double)(N + 1) * (N + 1);
            for (int i = 0; i < N; i++) sv[i] = ck[i];
            sv[N] = 1.0;
            al(0, N) = 0;
            for (int i = 0; i <= N; i++) al(0, i) = al(0, i) + qq * sv[i];
            for (int i = 0; i <= N; i++) {
                const double t = r * sv[i];
                for (int j = 0; j <= N; j++) c(i, j) = c(i, j) + (0.0 + t * sv[j]);
            }
            b1[N] = x1; b2[N] = x2; idx[N] = orig;
            e[N] = -1.0;
            const double ig = 1.0 / gamma;
            for (int i = 0; i <= N; i++) {
                const double t = ig * e[i];
                for (int j = 0; j <= N; j++) q(i, j) = q(i, j) + (0.0 + t * e[j]);
            }
            N++;
        }
        while (N > P.capacity && P.capacity > 0) {                           // :206-223
            double minscore = 0;
            int minloc = -1;
            for (int i = 0; i < N; i++) {
                const double score = al(0, i) * al(0, i) / (q(i, i) + c(i, i));
                if (i == 0 || score < minscore) { minscore = score; minloc = i; }
            }
            delete_bv_ref(minloc);
            st.n_del_cap++;
        }
        double minscore = 0;                                                 // :226-242
        int minloc = -1;
        while (minscore < GEO9 && N > 1) {
            for (int i = 0; i < N; i++) {
                const double score = 1.0 / q(i, i);
                if (i == 0 || score < minscore) { minscore = score; minloc = i; }
            }
            if (minscore < GEO9) {
                delete_bv_ref(minloc);
                st.n_del_geo++;
            }
        }
        if (std::isnan(c(0, 0))) flags |= 1;
    }
    // sparse_gp.hpp:312-351: f* and sqrt(s20 + k** + k'Ck), reference order
    double predict_ref(double x1, double x2, double* sigma) {
        const double kstar = P.p0 * std::exp(P.cl * 0.0);
        if (N == 0) {
            if (sigma) *sigma = std::sqrt(kstar + P.s20);
            return 0.0;
        }
        for (int i = 0; i < N; i++) k[i] = kern_ref(x1, x2, b1[i], b2[i]);
        double f = 0.0;
        for (int i = 0; i < N; i++) f += al(0, i) * k[i];
        if (sigma) {
            double kCk = 0.0;
            for (int j = 0; j < N; j++) {
                double t = 0.0;
                for (int i = 0; i < N; i++) t += k[i] * c(i, j);
                kCk += t * k[j];
            }
            double sg = P.s20 + kstar + kCk;
            if (sg < 0) sg = 0;
            *sigma = std::sqrt(sg);
        }
        return f;
    }

    // sparse_gp.hpp:89-249 / sparse_gp_field.hpp:59-222 ; y has D entries
    void add(double x1, double x2, const double* y, int orig) {
        if (P.ref_order && D == 1) { add_ref(x1, x2, y[0], orig); return; }
        st.n_add++;
        const double kstar = P.p0;  // kernel(X,X) = p0*exp(-0) exactly
        if (N == 0) {
            double d = kstar + P.s20;
            for (int ch = 0; ch < D; ch++) al(ch, 0) = y[ch] / d;
            c(0, 0) = -1.0 / d;
            q(0, 0) = 1.0 / kstar;
            b1[0] = x1; b2[0] = x2; idx[0] = orig;
            N = 1;
            st.n_first++;
            if (std::isnan(c(0, 0))) flags |= 1;
            return;
        }
        st.sumN += N;
        st.sumN2_common += (double)N * N;
        for (int i = 0; i < N; i++) k[i] = kern(x1, x2, b1[i], b2[i]);
        double m[3] = {0, 0, 0};
        for (int ch = 0; ch < D; ch++) m[ch] = dot32(&alpha[(size_t)ch * ld], k.data(), N);
        for (int i = 0; i < N; i++) ck[i] = row4(&C[(size_t)i * ld], k.data(), N);
        double s2 = kstar + dot32(k.data(), ck.data(), N);
        double den = P.s20 + s2;
        double r = -1.0 / den;      // gaussian_noise.cpp:15-18 / gaussian_noise_3d.cpp:16-19
        double qq[3] = {0, 0, 0};
        for (int ch = 0; ch < D; ch++) qq[ch] = (y[ch] - m[ch]) / den;  // gaussian_noise.cpp:9-12 / gaussian_noise_3d.cpp:10-13
        for (int i = 0; i < N; i++) e[i] = row4(&Q[(size_t)i * ld], k.data(), N);
        double gamma = kstar - dot32(k.data(), e.data(), N);
        if (gamma < TINY12) gamma = 0;
        if (gamma < P.eps_tol && P.capacity != -1) {
            st.n_sparse++;
            st.sumN2_sparse += (double)N * N;
            double eta = 1.0 / (1.0 + gamma * r);
            for (int i = 0; i < N; i++) sv[i] = ck[i] + e[i];
            for (int ch = 0; ch < D; ch++) {
                double qe = qq[ch] * eta;
                for (int i = 0; i < N; i++) al(ch, i) = al(ch, i) + sv[i] * qe;
            }
            double re = r * eta;
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) c(i, j) = std::fma(re, sv[i] * sv[j], c(i, j));
        } else {
            st.n_full++;
            st.sumN2_full += (double)(N + 1) * (N + 1);
            for (int i = 0; i < N; i++) sv[i] = ck[i];
            sv[N] = 1.0;
            for (int ch = 0; ch < D; ch++) {
                for (int i = 0; i < N; i++) al(ch, i) = al(ch, i) + qq[ch] * sv[i];
                al(ch, N) = 0.0 + qq[ch] * sv[N];
            }
            for (int i = 0; i <= N; i++)
                for (int j = 0; j <= N; j++) c(i, j) = std::fma(r, sv[i] * sv[j], c(i, j));
            b1[N] = x1; b2[N] = x2; idx[N] = orig;
            e[N] = -1.0;
            double ig = 1.0 / gamma;
            for (int i = 0; i <= N; i++)
                for (int j = 0; j <= N; j++) q(i, j) = std::fma(ig, e[i] * e[j], q(i, j));
            N++;
        }
        // capacity deletions (sparse_gp.hpp:206-223 / sparse_gp_field.hpp:181-197)
        while (N > P.capacity && P.capacity > 0) {
            double minscore = 0;
            int minloc = -1;
            for (int i = 0; i < N; i++) {
                double num = al(0, i) * al(0, i);
                if (D == 3) num = num + (al(1, i) * al(1, i) + al(2, i) * al(2, i));  // alpha.row(i).squaredNorm()
                double score = num / (q(i, i) + c(i, i));
                if (i == 0 || score < minscore) { minscore = score; minloc = i; }
            }
            delete_bv(minloc);
            st.n_del_cap++;
        }
        // geometric deletions (sparse_gp.hpp:226-242)
        double minscore = 0;
        int minloc = -1;
        while (minscore < GEO9 && N > 1) {
            for (int i = 0; i < N; i++) {
                double score = 1.0 / q(i, i);
                if (i == 0 || score < minscore) { minscore = score; minloc = i; }
            }
            if (minscore < GEO9) {
                delete_bv(minloc);
                st.n_del_geo++;
            }
        }
        if (std::isnan(c(0, 0))) flags |= 1;
    }
    void add(double x1, double x2, double y, int orig) { add(x1, x2, &y, orig); }

    // sparse_gp.hpp:312-351 (mean only; the caller discards sigma, gp_compressor.cpp:333)
    double predict(double x1, double x2) {
        if (P.ref_order && D == 1) return predict_ref(x1, x2, nullptr);
        if (N == 0) return 0.0;
        for (int i = 0; i < N; i++) k[i] = kern(x1, x2, b1[i], b2[i]);
        return row4(alpha.data(), k.data(), N);
    }
    // sparse_gp_field.hpp:284-320 (mean of every channel)
    void predict_field(double x1, double x2, double* f) {
        for (int ch = 0; ch < D; ch++) f[ch] = 0.0;
        if (N == 0) return;
        for (int i = 0; i < N; i++) k[i] = kern(x1, x2, b1[i], b2[i]);
        for (int ch = 0; ch < D; ch++) f[ch] = row4(&alpha[(size_t)ch * ld], k.data(), N);
    }
    // Grid decode (gp_compressor.cpp:320-334): the query points form a lattice, so the product path evaluates the RBF
    // kernel separably, k_i(a, b) = (p0 exp(cl (X0_a - b1_i)^2)) * exp(cl (X1_b - b2_i)^2), which is within 3 ulp of
    // kern(); this is its restatement (test_oracle_kat pins the distance to the direct form).
    void grid_tables(const double* Xs, int nx, const double* Ys, int ny, std::vector<double>& Ex, std::vector<double>& Ey) const {
        Ex.resize((size_t)std::max(N, 1) * nx);
        Ey.resize((size_t)std::max(N, 1) * ny);
        for (int i = 0; i < N; i++) {
            for (int a = 0; a < nx; a++) { const double d = Xs[a] - b1[i]; Ex[(size_t)i * nx + a] = P.p0 * orc_exp_impl(P.cl * (d * d)); }
            for (int b = 0; b < ny; b++) { const double d = Ys[b] - b2[i]; Ey[(size_t)i * ny + b] = orc_exp_impl(P.cl * (d * d)); }
        }
    }
    void grid_k(const std::vector<double>& Ex, int nx, int a, const std::vector<double>& Ey, int ny, int b) {
        for (int i = 0; i < N; i++) k[i] = Ex[(size_t)i * nx + a] * Ey[(size_t)i * ny + b];
    }
    double predict_grid() const { return N == 0 ? 0.0 : row4(alpha.data(), k.data(), N); }
    void predict_field_grid(double* f) const {
        for (int ch = 0; ch < D; ch++) f[ch] = (N == 0) ? 0.0 : row4(&alpha[(size_t)ch * ld], k.data(), N);
    }
    // ---- next rows N2 / N4: sparse_gp::predict with sigma / conf (sparse_gp.hpp:312-351), likelihood (:407-425) and
    // likelihood_dx (:472-502, with rbf_kernel::kernel_dx rbf_kernel.cpp:38-46) at one point.  Canonical order:
    // (C k)_i = sequential fma over j of C(j, i) k_j; every O(N) sum is a row4.  out: f, sigma, conf, lik, dx[3].
    void evaluate(double x1, double x2, double y, double* out) {
        std::vector<double> kx(std::max(N, 1)), ky(std::max(N, 1));
        const double kstar = P.p0;
        const double c1 = (-P.p0) / P.p1;                       // -p(0)/p(1), rbf_kernel.cpp:44
        for (int i = 0; i < N; i++) {
            const double d1 = x1 - b1[i], d2 = x2 - b2[i];
            const double e = orc_exp_impl(P.cl * (d1 * d1 + d2 * d2));
            k[i] = P.p0 * e;
            kx[i] = (c1 * d1) * e;
            ky[i] = (c1 * d2) * e;
        }
        for (int i = 0; i < N; i++) {
            double a = 0.0;
            for (int j = 0; j < N; j++) a = std::fma(C[(size_t)j * ld + i], k[j], a);
            ck[i] = a;
        }
        const double kCk = N ? row4(k.data(), ck.data(), N) : 0.0;
        const double sx = N ? row4(kx.data(), ck.data(), N) : 0.0, sy = N ? row4(ky.data(), ck.data(), N) : 0.0;
        const double mu = N ? row4(alpha.data(), k.data(), N) : 0.0;
        const double ax = N ? row4(alpha.data(), kx.data(), N) : 0.0, ay = N ? row4(alpha.data(), ky.data(), N) : 0.0;
        // predict (:329-349)
        double var = (P.s20 + kstar) + kCk;
        const double var_l = var;                                 // likelihood (:421) has no clamp
        if (var < 0) var = 0;
        out[0] = mu;
        out[1] = std::sqrt(var);
        out[2] = 100.0 * (1.0 - var / (kstar + P.s20));
        // likelihood (:424)
        const double off = y - mu;
        out[3] = (1.0 / std::sqrt((2.0 * M_PI) * var_l)) * orc_exp_impl(((-0.5 / var_l) * off) * off);
        // likelihood_dx (:487-499)
        const double var_d = (P.s20 + kCk) + kstar;
        const double sdx[2] = {2.0 * sx, 2.0 * sy};
        const double sq = std::sqrt(var_d);
        const double vs = var_d * sq;
        const double exppart = (0.5 / vs) * orc_exp_impl(((-0.5 / var_d) * off) * off);
        const double ad[2] = {ax, ay};
        out[4] = ((-1.0 / vs) * off) * exppart;
        for (int d = 0; d < 2; d++) {
            const double first = -sdx[d], second = (2.0 * ad[d]) * off, third = ((sdx[d] / var_d) * off) * off;
            out[5 + d] = exppart * ((first + second) + third);
        }
    }
    // The same for the RGB field GP (D = 3): sparse_gp_field::predict with sigma / conf (sparse_gp_field.hpp:284-320),
    // likelihood (:333-351) and likelihood_dx (:367-393).  y: 3 values.  out: f[3], sigma, conf, lik, dx[3] (dx[0] = 0).
    void evaluate_field(double x1, double x2, const double* y, double* out) {
        std::vector<double> kx(std::max(N, 1)), ky(std::max(N, 1));
        const double kstar = P.p0;
        const double c1 = (-P.p0) / P.p1;
        for (int i = 0; i < N; i++) {
            const double d1 = x1 - b1[i], d2 = x2 - b2[i];
            const double e = orc_exp_impl(P.cl * (d1 * d1 + d2 * d2));
            k[i] = P.p0 * e;
            kx[i] = (c1 * d1) * e;
            ky[i] = (c1 * d2) * e;
        }
        for (int i = 0; i < N; i++) {
            double a = 0.0;
            for (int j = 0; j < N; j++) a = std::fma(C[(size_t)j * ld + i], k[j], a);
            ck[i] = a;
        }
        const double kCk = N ? row4(k.data(), ck.data(), N) : 0.0;
        const double sx = N ? row4(kx.data(), ck.data(), N) : 0.0, sy = N ? row4(ky.data(), ck.data(), N) : 0.0;
        double mu[3], ax[3], ay[3], off[3];
        for (int c = 0; c < 3; c++) {
            const double* a = &alpha[(size_t)c * ld];
            mu[c] = N ? row4(a, k.data(), N) : 0.0;
            ax[c] = N ? row4(a, kx.data(), N) : 0.0;
            ay[c] = N ? row4(a, ky.data(), N) : 0.0;
            off[c] = y[c] - mu[c];
            out[c] = mu[c];
        }
        double var = (P.s20 + kstar) + kCk;
        const double var_l = var;
        if (var < 0) var = 0;
        out[3] = std::sqrt(var);
        out[4] = 100.0 * (1.0 - var / (kstar + P.s20));
        const double sqn = (off[0] * off[0] + off[1] * off[1]) + off[2] * off[2];      // squaredNorm
        const double pow2pi3 = std::pow(2.0 * M_PI, 3.0);                               // pow(2.0f*M_PI, double(y.rows()))
        out[5] = (1.0 / std::sqrt(pow2pi3 * var_l)) * orc_exp_impl((-0.5 / var_l) * sqn);
        const double var_d = (P.s20 + kCk) + kstar;
        const double sdx[2] = {2.0 * sx, 2.0 * sy};
        const double sq = std::sqrt(var_d);
        const double exppart = (0.5 / (var_d * sq)) * orc_exp_impl((-0.5 / var_d) * sqn);
        out[6] = 0.0;
        for (int d = 0; d < 2; d++) {
            const double* A = d ? ay : ax;
            const double second = 2.0 * ((A[0] * off[0] + A[1] * off[1]) + A[2] * off[2]);
            const double third = (sdx[d] / var_d) * sqn;
            out[7 + d] = exppart * ((-sdx[d] + second) + third);
        }
    }
    // predictive sigma as sparse_gp.hpp:329-347 (conf = false): sqrt(s20 + kstar + k'Ck), the evaluate() arithmetic
    double predict_sigma(double x1, double x2) {
        double r[7];
        evaluate(x1, x2, 0.0, r);
        return r[1];
    }
};

// sparse_gp.hpp:42-56 : ind = 0..n-1; for i = n-1..1: r = rand() % i; swap(ind[i], ind[r])
void shuffle_ref(std::vector<int>& ind, int n, GlibcRand& g) {
    ind.resize(n);
    for (int i = 0; i < n; i++) ind[i] = i;
    for (int i = n - 1; i > 0; --i) {
        int r = (int)(g.next() % (uint32_t)i);
        std::swap(ind[i], ind[r]);
    }
}

// ------------------------------------------------------------------------------------
// Lattice = PCL OctreePointCloud bounding-box growth  [RECALLED PCL 1.7
// octree_pointcloud.hpp: adoptBoundingBoxToPoint / getKeyBitSize / genOctreeKeyforPoint /
// genLeafNodeCenterFromOctreeKey].  min/max are doubles, point coordinates floats.
// ------------------------------------------------------------------------------------
struct Lattice {
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    double res = 0;
    unsigned depth = 0;
    bool defined = false;
};

inline bool finite3(const float* p) {
    return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
}

void lattice_add_point(Lattice& L, const float* p) {
    const float minValue = std::numeric_limits<float>::epsilon();
    while (true) {
        bool lo[3], up[3];
        for (int a = 0; a < 3; a++) { lo[a] = (double)p[a] < L.mn[a]; up[a] = (double)p[a] >= L.mx[a]; }
        bool viol = lo[0] || lo[1] || lo[2] || up[0] || up[1] || up[2] || !L.defined;
        if (!viol) break;
        if (L.defined) {
            double side = (double)(1u << L.depth) * L.res;
            for (int a = 0; a < 3; a++)
                if (!up[a]) L.mn[a] -= side;
            L.depth++;
            side = (double)(1u << L.depth) * L.res - minValue;
            for (int a = 0; a < 3; a++) L.mx[a] = L.mn[a] + side;
        } else {
            for (int a = 0; a < 3; a++) {
                L.mn[a] = (double)p[a] - L.res / 2;
                L.mx[a] = (double)p[a] + L.res / 2;
            }
            // getKeyBitSize()
            unsigned mk[3];
            for (int a = 0; a < 3; a++) mk[a] = (unsigned)((L.mx[a] - L.mn[a]) / L.res);
            unsigned maxv = std::max(std::max(std::max(mk[0], mk[1]), mk[2]), 2u);
            double lg = std::log((double)maxv) / std::log(2.0);
            L.depth = (unsigned)std::ceil(lg - minValue);
            double side = (double)(1u << L.depth) * L.res - minValue;
            for (int a = 0; a < 3; a++) {
                double over = (side - (L.mx[a] - L.mn[a])) / 2.0;
                L.mn[a] -= over;
                L.mx[a] += over;
            }
            L.defined = true;
        }
    }
}

inline uint64_t morton_encode(uint32_t kx, uint32_t ky, uint32_t kz, unsigned depth) {
    uint64_t c = 0;
    for (int b = (int)depth - 1; b >= 0; --b) {
        // child index = (x_bit<<2)|(y_bit<<1)|z_bit  (gp_octree.cpp:138-140)
        c = (c << 3) | (uint64_t)((((kx >> b) & 1u) << 2) | (((ky >> b) & 1u) << 1) | ((kz >> b) & 1u));
    }
    return c;
}

// ------------------------------------------------------------------------------------
// Plane-fit rotation (gp_compressor.cpp:29-64).  The reference takes the right singular
// vector of the smallest singular value of the m x 4 matrix A = [x y z 1] via Eigen
// JacobiSVD.  Canonical restatement: A = A_c * S with A_c = [p - c, 1] (c = voxel
// centre, exact in double) and S = [[I,0],[c',1]]; G_c = A_c' A_c from canonical sums;
// G_c = L D L' ; M = sqrt(D) L' S has the same right singular vectors as A; one-sided
// (Hestenes) Jacobi on the 4x4 M.
// ------------------------------------------------------------------------------------
struct GramSums {
    double s[10];  // xx xy xz yy yz zz x y z n
};

void smallest_right_singular_vector(const GramSums& g, const double c[3], double v[4]) {
    double G[4][4];
    G[0][0] = g.s[0]; G[0][1] = g.s[1]; G[0][2] = g.s[2]; G[0][3] = g.s[6];
    G[1][1] = g.s[3]; G[1][2] = g.s[4]; G[1][3] = g.s[7];
    G[2][2] = g.s[5]; G[2][3] = g.s[8];
    G[3][3] = g.s[9];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < i; j++) G[i][j] = G[j][i];
    // LDL' (no pivoting; non-positive pivots are clamped to zero)
    double Lm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double D[4];
    for (int j = 0; j < 4; j++) {
        double d = G[j][j];
        for (int t = 0; t < j; t++) d = d - (Lm[j][t] * Lm[j][t]) * D[t];
        if (!(d > 0.0)) d = 0.0;
        D[j] = d;
        for (int i = j + 1; i < 4; i++) {
            double a = G[i][j];
            for (int t = 0; t < j; t++) a = a - (Lm[i][t] * Lm[j][t]) * D[t];
            Lm[i][j] = (d > 0.0) ? a / d : 0.0;
        }
    }
    // U = sqrt(D) L'  (upper triangular), M = U S,  S = [[I,0],[c',1]]
    double M[4][4];
    for (int i = 0; i < 4; i++) {
        double sd = std::sqrt(D[i]);
        double u[4];
        for (int j = 0; j < 4; j++) u[j] = (j >= i) ? sd * Lm[j][i] : 0.0;
        for (int j = 0; j < 3; j++) M[i][j] = std::fma(u[3], c[j], u[j]);
        M[i][3] = u[3];
    }
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double a = 0, b = 0, gpq = 0;
                for (int t = 0; t < 4; t++) {
                    a = std::fma(M[t][p], M[t][p], a);
                    b = std::fma(M[t][q], M[t][q], b);
                    gpq = std::fma(M[t][p], M[t][q], gpq);
                }
                if (gpq == 0.0) continue;
                if (gpq * gpq <= 0x1p-106 * (a * b)) continue;
                rotated = true;
                double zeta = (b - a) / (2.0 * gpq);
                double t = 1.0 / (std::fabs(zeta) + std::sqrt(std::fma(zeta, zeta, 1.0)));
                if (zeta < 0.0) t = -t;
                double cs = 1.0 / std::sqrt(std::fma(t, t, 1.0));
                double sn = cs * t;
                for (int r = 0; r < 4; r++) {
                    double mp = M[r][p], mq = M[r][q];
                    M[r][p] = cs * mp - sn * mq;
                    M[r][q] = sn * mp + cs * mq;
                    double vp = V[r][p], vq = V[r][q];
                    V[r][p] = cs * vp - sn * vq;
                    V[r][q] = sn * vp + cs * vq;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bestn = 0;
    for (int j = 0; j < 4; j++) {
        double nn = 0;
        for (int t = 0; t < 4; t++) nn = std::fma(M[t][j], M[t][j], nn);
        if (j == 0 || nn < bestn) { bestn = nn; best = j; }
    }
    for (int r = 0; r < 4; r++) v[r] = V[r][best];
}

inline void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void normalize3(double v[3]) {
    double n2 = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);  // Eigen fixed-size redux tree [RECALLED]
    double n = std::sqrt(n2);
    v[0] = v[0] / n; v[1] = v[1] / n; v[2] = v[2] / n;
}

// gp_compressor.cpp:38-63 ; R stored row-major R[r*3+c], columns are the patch axes.
void rotation_from_normal(double nrm[3], double R[9]) {
    normalize3(nrm);
    double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
    double c1[3];
    double ax = std::fabs(nrm[0]), ay = std::fabs(nrm[1]), az = std::fabs(nrm[2]);
    if (ax > ay && ax > az) {
        if (nrm[0] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ez, nrm, c1);
    } else if (ay > ax && ay > az) {
        if (nrm[1] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ex, nrm, c1);
    } else {
        if (nrm[2] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ey, nrm, c1);
    }
    normalize3(c1);
    double c2[3];
    cross3(nrm, c1, c2);
    for (int r = 0; r < 3; r++) { R[r * 3 + 0] = nrm[r]; R[r * 3 + 1] = c1[r]; R[r * 3 + 2] = c2[r]; }
}

// Eigen Quaterniond <- Matrix3d and toRotationMatrix()  [RECALLED Eigen 3.2
// Geometry/Quaternion.h]; gp_compressor.cpp:240 stores R as a quaternion and :339
// rebuilds the matrix, so the decoder sees the round-tripped matrix.  q = (x,y,z,w).
void rot_to_quat(const double R[9], double q[4]) {
    auto m = [&](int r, int c) { return R[r * 3 + c]; };
    double t = (m(0, 0) + m(1, 1)) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t;
        q[1] = (m(0, 2) - m(2, 0)) * t;
        q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(((m(i, i) - m(j, j)) - m(k, k)) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t;
        q[j] = (m(j, i) + m(i, j)) * t;
        q[k] = (m(k, i) + m(i, k)) * t;
    }
}
void quat_to_rot(const double q[4], double R[9]) {
    double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// gp_compressor.cpp:251-265 ; Eigen cast<short> of a double is a C cast.
inline int flatten_color(double x) {
    if (std::isnan(x) || std::isinf(x)) return 255;
    // x86-64 double -> short: cvttsd2si to int32 then truncation to 16 bits
    short s = (std::fabs(x) < 2147483648.0) ? (short)(int)x : (short)0;
    if (s < 0) return 0;
    if (s > 255) return 255;
    return s;
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) { f(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    std::atomic<int64_t> next(0);
    int64_t chunk = std::max<int64_t>(1, n / (nthreads * 16));
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            while (true) {
                int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(t, b, std::min(n, b + chunk));
            }
        });
    for (auto& x : th) x.join();
}

struct Config {
    double res = (double)0.1f;
    int sz = 10;
    int capacity = 100;
    double s0 = (double)1e-1f;
    double eps_tol = (double)1e-6f;
    double sigmaf_sq = (double)100e-0f;
    double l_sq = 1.0;
    int leaf_order = 0;  // 0 = reverse Morton (PCL 1.7/1.8 depth-first iterator), 1 = Morton
    int shuffle = 1;     // 0 disables sparse_gp::shuffle (matlab/sogp.m has none)
    int rgb_rand = 1;    // account for the RGB field GP's shuffle in the rand stream
    int threads = 1;
    int rgb = 0;         // 1: also fit / decode the RGB field GP (sparse_gp_field, gp_compressor.cpp:163,334)
    int ref_order = 0;   // 1: height GPs in the reference source's own evaluation order (SogpParams::ref_order)
    int decode_separable = 0;  // 0: grid decode evaluates rbf_kernel::kernel_function per (grid point, BV) as the reference does
                               // (sparse_gp.hpp:320-327); 1: the product's flagged fast mode, separable on the lattice
    double rgb_s0 = (double)1e2f;       // sparse_gp_field.h:43
    double rgb_eps_tol = (double)1e-4f; // sparse_gp_field.hpp:16
};

struct Oracle {
    Config cfg;
    std::string err;
    uint64_t rand_offset = 0;  // number of rand() values consumed so far by this handle

    // ---- binning results ----
    Lattice lat;
    int64_t n_in = 0, n_leaves = 0, n_claimed = 0;
    std::vector<uint64_t> leaf_code;           // gp_index order
    std::vector<float> leaf_center;            // 3 per leaf
    std::vector<int32_t> leaf_ncand;           // candidates per leaf
    std::vector<double> leaf_R, leaf_quat, leaf_mean, leaf_rgbmean;  // 9,4,3,3
    std::vector<int64_t> patch_off;            // n_leaves+1 offsets into the stream
    std::vector<int32_t> owner;                // per input point: gp_index or -1
    std::vector<int32_t> st_idx;               // stream: original point index
    std::vector<double> st_x1, st_x2, st_y;    // stream: local coordinates (candidate order)
    std::vector<int32_t> st_perm;              // per patch shuffle permutation (stream-local)
    std::vector<double> st_c;                  // stream: centred colours (r,g,b per point), gp_compressor.cpp:105
    std::vector<int32_t> st_perm_rgb;          // the RGB field GP's own shuffle
    std::vector<int32_t> rgb_nbv, rgb_bv_idx;  // RGB field GP results
    std::vector<int64_t> rgb_bv_off;
    std::vector<double> rgb_bv1, rgb_bv2, rgb_alpha;  // rgb_alpha: 3 per BV (r,g,b)
    std::vector<double> rgb_dumpC;             // optional dense C of the field GPs (N*N per patch)
    std::vector<int64_t> rgb_dump_off;
    SogpStats stats_rgb;
    // ---- fit results ----
    std::vector<int32_t> nbv;                  // per patch
    std::vector<int64_t> fed;                  // points fed to each patch's height GP so far (continued fits)
    std::vector<int64_t> bv_off;               // n_leaves+1
    std::vector<int32_t> bv_idx;               // patch-local stream position of each BV
    std::vector<double> bv1, bv2, alpha;
    std::vector<int32_t> fit_flags;
    std::vector<double> dumpC, dumpQ;          // optional dense dumps (N*N per patch, row-major)
    std::vector<int64_t> dump_off;
    SogpStats stats;
    double t_project = 0, t_train = 0, t_decode = 0;

    SogpParams sogp_params() const {
        SogpParams p;
        p.capacity = cfg.capacity;
        p.s20 = cfg.s0;
        p.eps_tol = cfg.eps_tol;
        p.p0 = cfg.sigmaf_sq;
        p.p1 = cfg.l_sq;
        p.cl = (double)(-0.5f) / cfg.l_sq;
        p.ref_order = cfg.ref_order;
        return p;
    }
    SogpParams rgb_params() const {  // sparse_gp_field(capacity, s0 = 1e2f), eps_tol 1e-4f, same default kernel
        SogpParams p = sogp_params();
        p.s20 = cfg.rgb_s0;
        p.eps_tol = cfg.rgb_eps_tol;
        return p;
    }

    // ---------------- project_cloud (gp_compressor.cpp:177-249) ----------------
    int project(const uint8_t* cloud32, int64_t n) {
        auto t0 = std::chrono::steady_clock::now();
        n_in = n;
        lat = Lattice();
        lat.res = cfg.res;
        auto P = [&](int64_t i) { return reinterpret_cast<const float*>(cloud32 + 32 * i); };
        auto RGB = [&](int64_t i, int ch) { return cloud32[32 * i + 16 + (2 - ch)]; };  // b,g,r,a
        // octree.addPointsFromInputCloud(): sequential bbox growth over finite points
        for (int64_t i = 0; i < n; i++)
            if (finite3(P(i))) lattice_add_point(lat, P(i));
        owner.assign(n, -1);
        leaf_code.clear(); leaf_center.clear(); leaf_ncand.clear(); leaf_R.clear();
        leaf_quat.clear(); leaf_mean.clear(); leaf_rgbmean.clear(); patch_off.assign(1, 0);
        st_idx.clear(); st_x1.clear(); st_x2.clear(); st_y.clear(); st_c.clear();
        n_leaves = 0; n_claimed = 0;
        if (!lat.defined) return 0;
        if (3 * lat.depth > 63) { err = "octree depth too large for a 64-bit Morton code"; return 2; }
        const int nth = cfg.threads;
        // keys + Morton sort (stable in index): ascending Morton = radiusSearch leaf order
        std::vector<uint64_t> code(n);
        std::vector<int32_t> order;
        order.reserve(n);
        for (int64_t i = 0; i < n; i++) {
            const float* p = P(i);
            if (!finite3(p)) continue;
            uint32_t kk[3];
            for (int a = 0; a < 3; a++) kk[a] = (uint32_t)(((double)p[a] - lat.mn[a]) / lat.res);
            code[i] = morton_encode(kk[0], kk[1], kk[2], lat.depth);
            order.push_back((int32_t)i);
        }
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return code[a] < code[b]; });
        // leaves in ascending Morton
        std::vector<uint64_t> lcode;
        std::vector<int64_t> lstart;
        for (int64_t s = 0; s < (int64_t)order.size(); s++)
            if (s == 0 || code[order[s]] != code[order[s - 1]]) { lcode.push_back(code[order[s]]); lstart.push_back(s); }
        lstart.push_back((int64_t)order.size());
        const int64_t NL = (int64_t)lcode.size();
        n_leaves = NL;
        std::unordered_map<uint64_t, int32_t> lmap;
        lmap.reserve(NL * 2);
        for (int64_t l = 0; l < NL; l++) lmap[lcode[l]] = (int32_t)l;
        auto decode = [&](uint64_t c, uint32_t kk[3]) {
            kk[0] = kk[1] = kk[2] = 0;
            for (unsigned b = 0; b < lat.depth; b++) {
                uint32_t ch = (uint32_t)((c >> (3 * b)) & 7u);
                kk[0] |= ((ch >> 2) & 1u) << b; kk[1] |= ((ch >> 1) & 1u) << b; kk[2] |= (ch & 1u) << b;
            }
        };
        // per-leaf (ascending index a): centre, neighbours (ascending Morton), rotation
        std::vector<float> cen(NL * 3);
        std::vector<int32_t> nbr(NL * 27, -1);
        std::vector<int32_t> nnbr(NL, 0);
        std::vector<double> Rl(NL * 9);
        std::vector<int32_t> ncand(NL, 0);
        const double radius = (double)(std::sqrt(3.0f) / 2.0f) * cfg.res;  // gp_compressor.cpp:194
        const double r2 = radius * radius;
        const uint32_t kmax = (lat.depth >= 32) ? 0xffffffffu : ((1u << lat.depth) - 1u);
        parallel_for(NL, nth, [&](int, int64_t b, int64_t e) {
            for (int64_t a = b; a < e; a++) {
                uint32_t kk[3];
                decode(lcode[a], kk);
                // genLeafNodeCenterFromOctreeKey: float((double(k)+0.5f)*res + min)
                for (int d = 0; d < 3; d++) cen[a * 3 + d] = (float)(((double)kk[d] + 0.5f) * lat.res + lat.mn[d]);
                int cnt = 0;
                int32_t* nb = &nbr[a * 27];
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dz = -1; dz <= 1; dz++) {
                            int64_t x = (int64_t)kk[0] + dx, y = (int64_t)kk[1] + dy, z = (int64_t)kk[2] + dz;
                            if (x < 0 || y < 0 || z < 0 || x > kmax || y > kmax || z > kmax) continue;
                            auto it = lmap.find(morton_encode((uint32_t)x, (uint32_t)y, (uint32_t)z, lat.depth));
                            if (it != lmap.end()) nb[cnt++] = it->second;
                        }
                std::sort(nb, nb + cnt);
                nnbr[a] = cnt;
            }
        });
        parallel_for(NL, nth, [&](int, int64_t b, int64_t e) {
            for (int64_t a = b; a < e; a++) {
                const float* c = &cen[a * 3];
                double part[10][32];
                for (int q = 0; q < 10; q++)
                    for (int l = 0; l < 32; l++) part[q][l] = 0.0;
                int m = 0;
                for (int t = 0; t < nnbr[a]; t++) {
                    int32_t v = nbr[a * 27 + t];
                    for (int64_t s = lstart[v]; s < lstart[v + 1]; s++) {
                        const float* p = P(order[s]);
                        // PCL pointSquaredDist: float (p - c).squaredNorm(), accept iff <= radius^2 (double)
                        float fx = p[0] - c[0], fy = p[1] - c[1], fz = p[2] - c[2];
                        float d2 = fx * fx + (fy * fy + fz * fz);
                        if ((double)d2 > r2) continue;
                        int l = (int)((s - lstart[v]) & 31);
                        double ux = (double)p[0] - (double)c[0], uy = (double)p[1] - (double)c[1], uz = (double)p[2] - (double)c[2];
                        part[0][l] = std::fma(ux, ux, part[0][l]); part[1][l] = std::fma(ux, uy, part[1][l]);
                        part[2][l] = std::fma(ux, uz, part[2][l]); part[3][l] = std::fma(uy, uy, part[3][l]);
                        part[4][l] = std::fma(uy, uz, part[4][l]); part[5][l] = std::fma(uz, uz, part[5][l]);
                        part[6][l] = part[6][l] + ux; part[7][l] = part[7][l] + uy; part[8][l] = part[8][l] + uz;
                        part[9][l] = part[9][l] + 1.0;
                        m++;
                    }
                }
                ncand[a] = m;
                double* R = &Rl[a * 9];
                if (m < 4) {  // gp_compressor.cpp:31-34
                    for (int q = 0; q < 9; q++) R[q] = (q % 4 == 0) ? 1.0 : 0.0;
                    continue;
                }
                GramSums g;
                for (int q = 0; q < 10; q++) g.s[q] = butterfly32(part[q]);
                double cd[3] = {(double)c[0], (double)c[1], (double)c[2]};
                double v[4];
                smallest_right_singular_vector(g, cd, v);
                double nrm[3] = {v[0], v[1], v[2]};
                rotation_from_normal(nrm, R);
            }
        });
        // visiting order: gp_index i  <->  ascending index a
        auto a_of_i = [&](int64_t i) { return cfg.leaf_order == 0 ? NL - 1 - i : i; };
        auto i_of_a = [&](int64_t a) { return cfg.leaf_order == 0 ? NL - 1 - a : a; };
        // owner(p) = first leaf in visiting order whose candidate list holds p and whose box
        // test accepts it (gp_compressor.cpp:80-89 with occupied_indices)
        const double half = cfg.res / 2.0f;
        std::vector<double> pt0(order.size()), pt1(order.size()), pt2(order.size());
        std::vector<int32_t> own_s(order.size(), -1);
        parallel_for(NL, nth, [&](int, int64_t b, int64_t e) {
            for (int64_t a = b; a < e; a++) {
                int cnt = nnbr[a];
                for (int64_t s = lstart[a]; s < lstart[a + 1]; s++) {
                    const float* p = P(order[s]);
                    for (int t = 0; t < cnt; t++) {
                        int32_t v = (cfg.leaf_order == 0) ? nbr[a * 27 + (cnt - 1 - t)] : nbr[a * 27 + t];
                        const float* c = &cen[v * 3];
                        float fx = p[0] - c[0], fy = p[1] - c[1], fz = p[2] - c[2];
                        float d2 = fx * fx + (fy * fy + fz * fz);
                        if ((double)d2 > r2) continue;
                        if (ncand[v] == 0) continue;
                        const double* R = &Rl[v * 9];
                        double d0 = (double)p[0] - (double)c[0], d1 = (double)p[1] - (double)c[1], dd2 = (double)p[2] - (double)c[2];
                        // pt = R' (p - centre)
                        double q0 = (R[0] * d0 + R[3] * d1) + R[6] * dd2;
                        double q1 = (R[1] * d0 + R[4] * d1) + R[7] * dd2;
                        double q2 = (R[2] * d0 + R[5] * d1) + R[8] * dd2;
                        if (q1 > half || q1 < -half || q2 > half || q2 < -half) continue;
                        own_s[s] = v; pt0[s] = q0; pt1[s] = q1; pt2[s] = q2;
                        break;
                    }
                }
            }
        });
        // group by owner in visiting order, keeping (Morton, index) order inside a patch
        std::vector<int64_t> cnt(NL + 1, 0);
        for (size_t s = 0; s < order.size(); s++)
            if (own_s[s] >= 0) cnt[i_of_a(own_s[s]) + 1]++;
        patch_off.assign(NL + 1, 0);
        for (int64_t i = 0; i < NL; i++) patch_off[i + 1] = patch_off[i] + cnt[i + 1];
        n_claimed = patch_off[NL];
        st_idx.assign(n_claimed, 0); st_x1.assign(n_claimed, 0); st_x2.assign(n_claimed, 0); st_y.assign(n_claimed, 0);
        st_c.assign(3 * n_claimed, 0);
        std::vector<double> st_h(n_claimed);
        {
            std::vector<int64_t> cur(patch_off.begin(), patch_off.end() - 1);
            for (size_t s = 0; s < order.size(); s++) {
                if (own_s[s] < 0) continue;
                int64_t i = i_of_a(own_s[s]);
                int64_t d = cur[i]++;
                st_idx[d] = order[s]; st_h[d] = pt0[s]; st_x1[d] = pt1[s]; st_x2[d] = pt2[s];
                owner[order[s]] = (int32_t)i;
            }
        }
        leaf_code.resize(NL); leaf_center.resize(NL * 3); leaf_ncand.resize(NL); leaf_R.resize(NL * 9);
        leaf_quat.resize(NL * 4); leaf_mean.resize(NL * 3); leaf_rgbmean.resize(NL * 3);
        parallel_for(NL, nth, [&](int, int64_t b, int64_t e) {
            for (int64_t i = b; i < e; i++) {
                int64_t a = a_of_i(i);
                leaf_code[i] = lcode[a];
                for (int d = 0; d < 3; d++) leaf_center[i * 3 + d] = cen[a * 3 + d];
                leaf_ncand[i] = ncand[a];
                for (int q = 0; q < 9; q++) leaf_R[i * 9 + q] = Rl[a * 9 + q];
                rot_to_quat(&Rl[a * 9], &leaf_quat[i * 4]);
                int64_t lo = patch_off[i], hi = patch_off[i + 1];
                // mn = mean of pt(0); colours mean (gp_compressor.cpp:88-102), canonical lane sums
                double ph[32], pc[3][32];
                for (int l = 0; l < 32; l++) { ph[l] = 0; pc[0][l] = pc[1][l] = pc[2][l] = 0; }
                for (int64_t s = lo; s < hi; s++) {
                    int l = (int)((s - lo) & 31);
                    ph[l] = ph[l] + st_h[s];
                    for (int ch = 0; ch < 3; ch++) pc[ch][l] = pc[ch][l] + (double)RGB(st_idx[s], ch);
                }
                double cntd = (double)(hi - lo);
                double mn = butterfly32(ph) / cntd;  // NaN for an empty patch, as in the reference
                for (int ch = 0; ch < 3; ch++) leaf_rgbmean[i * 3 + ch] = butterfly32(pc[ch]) / cntd;
                for (int64_t s = lo; s < hi; s++) {
                    st_y[s] = st_h[s] - mn;
                    for (int ch = 0; ch < 3; ch++) st_c[3 * s + ch] = (double)RGB(st_idx[s], ch) - leaf_rgbmean[i * 3 + ch];
                }
                for (int d = 0; d < 3; d++)
                    leaf_mean[i * 3 + d] = (double)cen[a * 3 + d] + mn * Rl[a * 9 + d * 3 + 0];  // centre += mn*R.col(0)
            }
        });
        t_project = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    }

    // ------------- train_processes (gp_compressor.cpp:121-175) on any patch stream -------------
    // cont: sparse_gp::add_measurements called again on the fitted processes (sparse_gp.hpp:59-86 accumulates): the new
    // points of every patch are shuffled with the next rand() draws and added to the kept state (needs dump on both calls)
    int train(int64_t NP, const int64_t* off, const double* x1, const double* x2, const double* y, int dump,
              const double* colours = nullptr, bool cont = false) {
        auto t0 = std::chrono::steady_clock::now();
        const SogpParams sp = sogp_params();
        const int64_t total = off[NP];
        const bool do_rgb = cfg.rgb && colours != nullptr;
        std::vector<int32_t> o_nbv, o_idx, o_flags;
        std::vector<int64_t> o_bvoff, o_dumpoff;
        std::vector<double> o_alpha, o_b1, o_b2, o_C, o_Q;
        if (cont) {
            if (do_rgb || !dump || (int64_t)nbv.size() != NP || dump_off.empty() || (int64_t)fed.size() != NP) { err = "continued fit needs a previous height-only fit of the same patches with dump"; return 4; }
            o_nbv = nbv; o_idx = bv_idx; o_flags = fit_flags; o_bvoff = bv_off; o_dumpoff = dump_off;
            o_alpha = alpha; o_b1 = bv1; o_b2 = bv2; o_C = dumpC; o_Q = dumpQ;
        } else {
            fed.assign(NP, 0);
        }
        if (do_rgb && !(cfg.shuffle && cfg.rgb_rand)) { err = "rgb needs shuffle and rgb_rand"; return 3; }
        st_perm_rgb.assign(do_rgb ? total : 0, 0);
        // rand stream: patch p's height shuffle starts after 2*(n_q-1) draws for every earlier
        // non-empty patch q (height GP then RGB field GP each shuffle; gp_compressor.cpp:162-163)
        st_perm.assign(total, 0);
        {
            GlibcRand g;
            for (uint64_t t = 0; t < rand_offset; t++) g.next();
            std::vector<int> ind;
            for (int64_t p = 0; p < NP; p++) {
                int n = (int)(off[p + 1] - off[p]);
                if (n == 0) continue;
                if (cfg.shuffle) {
                    shuffle_ref(ind, n, g);
                    for (int i = 0; i < n; i++) st_perm[off[p] + i] = ind[i];
                    if (do_rgb) {  // RGB_gps[i].add_measurements(X, C): its own shuffle, gp_compressor.cpp:163
                        shuffle_ref(ind, n, g);
                        for (int i = 0; i < n; i++) st_perm_rgb[off[p] + i] = ind[i];
                    } else if (cfg.rgb_rand) {
                        for (int i = n - 1; i > 0; --i) g.next();
                    }
                    rand_offset += (uint64_t)(n - 1) * (cfg.rgb_rand ? 2 : 1);
                } else {
                    for (int i = 0; i < n; i++) st_perm[off[p] + i] = i;
                }
            }
        }
        nbv.assign(NP, 0);
        fit_flags.assign(NP, 0);
        const int cap = cfg.capacity;
        std::vector<std::vector<double>> outA(NP), outB1(NP), outB2(NP), outC(NP), outQ(NP);
        std::vector<std::vector<int32_t>> outI(NP);
        const int nth = cfg.threads;
        std::vector<SogpStats> tstats(std::max(1, nth));
        parallel_for(NP, nth, [&](int tid, int64_t b, int64_t e) {
            Sogp gp;
            for (int64_t p = b; p < e; p++) {
                int n = (int)(off[p + 1] - off[p]);
                if (n == 0 && !(cont && o_nbv[p] > 0)) continue;
                int maxn = (cap > 0) ? std::min(cap, n) : n;
                int base = 0;
                if (cont) {  // resume from the kept state
                    const int N0 = o_nbv[p];
                    maxn = (cap > 0) ? cap : N0 + n;
                    gp.init(sp, std::max(maxn, 1));
                    gp.N = N0;
                    gp.flags = o_flags[p];
                    for (int i = 0; i < N0; i++) {
                        gp.alpha[i] = o_alpha[o_bvoff[p] + i]; gp.b1[i] = o_b1[o_bvoff[p] + i]; gp.b2[i] = o_b2[o_bvoff[p] + i];
                        gp.idx[i] = o_idx[o_bvoff[p] + i];
                        for (int j = 0; j < N0; j++) {
                            gp.c(i, j) = o_C[o_dumpoff[p] + (size_t)i * N0 + j];
                            gp.q(i, j) = o_Q[o_dumpoff[p] + (size_t)i * N0 + j];
                        }
                    }
                    base = (int)fed[p];
                } else {
                    gp.init(sp, maxn);
                }
                const int64_t o = off[p];
                for (int t = 0; t < n; t++) {
                    int s = st_perm[o + t];
                    gp.add(x1[o + s], x2[o + s], y[o + s], base + s);
                }
                nbv[p] = gp.N;
                fit_flags[p] = gp.flags;
                outA[p].assign(gp.alpha.begin(), gp.alpha.begin() + gp.N);
                outB1[p].assign(gp.b1.begin(), gp.b1.begin() + gp.N);
                outB2[p].assign(gp.b2.begin(), gp.b2.begin() + gp.N);
                outI[p].assign(gp.idx.begin(), gp.idx.begin() + gp.N);
                if (dump) {
                    outC[p].resize((size_t)gp.N * gp.N); outQ[p].resize((size_t)gp.N * gp.N);
                    for (int i = 0; i < gp.N; i++)
                        for (int j = 0; j < gp.N; j++) { outC[p][(size_t)i * gp.N + j] = gp.c(i, j); outQ[p][(size_t)i * gp.N + j] = gp.q(i, j); }
                }
                tstats[tid].merge(gp.st);
            }
        });
        stats = SogpStats();
        for (auto& s : tstats) stats.merge(s);
        for (int64_t p = 0; p < NP; p++) fed[p] += off[p + 1] - off[p];
        // ---- RGB field GP (sparse_gp_field<rbf_kernel, gaussian_noise_3d>), same points, own shuffle ----
        rgb_nbv.assign(do_rgb ? NP : 0, 0);
        rgb_bv_off.assign(NP + 1, 0);
        rgb_bv_idx.clear(); rgb_bv1.clear(); rgb_bv2.clear(); rgb_alpha.clear();
        stats_rgb = SogpStats();
        if (do_rgb) {
            const SogpParams rp = rgb_params();
            std::vector<std::vector<double>> rA(NP), rB1(NP), rB2(NP), rC(NP);
            std::vector<std::vector<int32_t>> rI(NP);
            std::vector<SogpStats> rstats(std::max(1, nth));
            parallel_for(NP, nth, [&](int tid, int64_t b, int64_t e) {
                Sogp gp;
                for (int64_t p = b; p < e; p++) {
                    int n = (int)(off[p + 1] - off[p]);
                    if (n == 0) continue;
                    int maxn = (cap > 0) ? std::min(cap, n) : n;
                    gp.init(rp, maxn, 3);
                    const int64_t o = off[p];
                    for (int t = 0; t < n; t++) {
                        int s = st_perm_rgb[o + t];
                        gp.add(x1[o + s], x2[o + s], &colours[3 * (o + s)], s);
                    }
                    rgb_nbv[p] = gp.N;
                    rA[p].resize((size_t)3 * gp.N);
                    for (int i = 0; i < gp.N; i++)
                        for (int ch = 0; ch < 3; ch++) rA[p][3 * i + ch] = gp.al(ch, i);
                    rB1[p].assign(gp.b1.begin(), gp.b1.begin() + gp.N);
                    rB2[p].assign(gp.b2.begin(), gp.b2.begin() + gp.N);
                    rI[p].assign(gp.idx.begin(), gp.idx.begin() + gp.N);
                    if (dump) {
                        rC[p].resize((size_t)gp.N * gp.N);
                        for (int i = 0; i < gp.N; i++)
                            for (int j = 0; j < gp.N; j++) rC[p][(size_t)i * gp.N + j] = gp.c(i, j);
                    }
                    rstats[tid].merge(gp.st);
                }
            });
            for (auto& s : rstats) stats_rgb.merge(s);
            for (int64_t p = 0; p < NP; p++) rgb_bv_off[p + 1] = rgb_bv_off[p] + rgb_nbv[p];
            rgb_dump_off.assign(NP + 1, 0);
            for (int64_t p = 0; p < NP; p++) rgb_dump_off[p + 1] = rgb_dump_off[p] + (dump ? (int64_t)rgb_nbv[p] * rgb_nbv[p] : 0);
            rgb_dumpC.resize(rgb_dump_off[NP]);
            if (dump)
                for (int64_t p = 0; p < NP; p++) std::copy(rC[p].begin(), rC[p].end(), rgb_dumpC.begin() + rgb_dump_off[p]);
            rgb_bv_idx.resize(rgb_bv_off[NP]); rgb_bv1.resize(rgb_bv_off[NP]); rgb_bv2.resize(rgb_bv_off[NP]);
            rgb_alpha.resize(3 * rgb_bv_off[NP]);
            for (int64_t p = 0; p < NP; p++) {
                std::copy(rA[p].begin(), rA[p].end(), rgb_alpha.begin() + 3 * rgb_bv_off[p]);
                std::copy(rB1[p].begin(), rB1[p].end(), rgb_bv1.begin() + rgb_bv_off[p]);
                std::copy(rB2[p].begin(), rB2[p].end(), rgb_bv2.begin() + rgb_bv_off[p]);
                std::copy(rI[p].begin(), rI[p].end(), rgb_bv_idx.begin() + rgb_bv_off[p]);
            }
        }
        bv_off.assign(NP + 1, 0);
        dump_off.assign(NP + 1, 0);
        for (int64_t p = 0; p < NP; p++) { bv_off[p + 1] = bv_off[p] + nbv[p]; dump_off[p + 1] = dump_off[p] + (dump ? (int64_t)nbv[p] * nbv[p] : 0); }
        bv_idx.resize(bv_off[NP]); bv1.resize(bv_off[NP]); bv2.resize(bv_off[NP]); alpha.resize(bv_off[NP]);
        dumpC.resize(dump_off[NP]); dumpQ.resize(dump_off[NP]);
        for (int64_t p = 0; p < NP; p++) {
            std::copy(outA[p].begin(), outA[p].end(), alpha.begin() + bv_off[p]);
            std::copy(outB1[p].begin(), outB1[p].end(), bv1.begin() + bv_off[p]);
            std::copy(outB2[p].begin(), outB2[p].end(), bv2.begin() + bv_off[p]);
            std::copy(outI[p].begin(), outI[p].end(), bv_idx.begin() + bv_off[p]);
            if (dump) {
                std::copy(outC[p].begin(), outC[p].end(), dumpC.begin() + dump_off[p]);
                std::copy(outQ[p].begin(), outQ[p].end(), dumpQ.begin() + dump_off[p]);
            }
        }
        t_train = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    }

    // ------------- load_compressed (gp_compressor.cpp:267-386) -------------
    // out32: 32-byte PointXYZRGB records; heights (optional): f* per grid point.
    int64_t decode_into(uint8_t* out32, double* heights, int with_sigma) {
        auto t0 = std::chrono::steady_clock::now();
        const int sz = cfg.sz;
        const int64_t g2 = (int64_t)sz * sz;
        const int64_t NP = (int64_t)nbv.size();
        std::vector<int64_t> slot(NP + 1, 0);
        for (int64_t p = 0; p < NP; p++) slot[p + 1] = slot[p] + (nbv[p] > 0 ? 1 : 0);
        const SogpParams sp = sogp_params();
        const bool have_frames = (int64_t)leaf_R.size() == NP * 9;
        std::atomic<double> sink(0.0);
        parallel_for(NP, cfg.threads, [&](int, int64_t b, int64_t e) {
            Sogp gp, gc;
            std::vector<double> Xg(sz), Ex, Ey, REx, REy;
            double acc = 0;
            for (int64_t p = b; p < e; p++) {
                int N = nbv[p];
                if (N == 0) continue;  // gp_compressor.cpp:299
                gp.init(sp, N);
                gp.N = N;
                for (int i = 0; i < N; i++) {
                    gp.alpha[i] = alpha[bv_off[p] + i]; gp.b1[i] = bv1[bv_off[p] + i]; gp.b2[i] = bv2[bv_off[p] + i];
                }
                if (with_sigma && !dumpC.empty())
                    for (int i = 0; i < N; i++)
                        for (int j = 0; j < N; j++) gp.c(i, j) = dumpC[dump_off[p] + (size_t)i * N + j];
                const bool have_rgb = (int64_t)rgb_nbv.size() == NP;
                if (have_rgb) {
                    const int NR = rgb_nbv[p];
                    gc.init(rgb_params(), std::max(NR, 1), 3);
                    gc.N = NR;
                    for (int i = 0; i < NR; i++) {
                        gc.b1[i] = rgb_bv1[rgb_bv_off[p] + i]; gc.b2[i] = rgb_bv2[rgb_bv_off[p] + i];
                        for (int ch = 0; ch < 3; ch++) gc.al(ch, i) = rgb_alpha[3 * (rgb_bv_off[p] + i) + ch];
                    }
                }
                double Rq[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, mean[3] = {0, 0, 0}, cm[3] = {0, 0, 0};
                if (have_frames) {
                    quat_to_rot(&leaf_quat[p * 4], Rq);
                    for (int d = 0; d < 3; d++) { mean[d] = leaf_mean[p * 3 + d]; cm[d] = leaf_rgbmean[p * 3 + d]; }
                }
                int64_t base = slot[p] * g2;
                int64_t m = 0;
                for (int a = 0; a < sz; a++) Xg[a] = cfg.res * (((double)a + 0.5f) / (double)sz - 0.5f);
                const bool sep = cfg.decode_separable != 0;
                if (sep) {
                    gp.grid_tables(Xg.data(), sz, Xg.data(), sz, Ex, Ey);
                    if (have_rgb) gc.grid_tables(Xg.data(), sz, Xg.data(), sz, REx, REy);
                }
                for (int yy = 0; yy < sz; yy++)
                    for (int xx = 0; xx < sz; xx++, m++) {
                        double X0 = Xg[xx];
                        double X1 = Xg[yy];
                        double f;
                        if (sep) { gp.grid_k(Ex, sz, xx, Ey, sz, yy); f = gp.predict_grid(); }
                        else f = gp.predict(X0, X1);
                        if (with_sigma) acc += gp.predict_sigma(X0, X1);  // the work the reference discards
                        if (heights) heights[base + m] = f;
                        if (out32) {
                            float* o = reinterpret_cast<float*>(out32 + 32 * (base + m));
                            for (int d = 0; d < 3; d++) {
                                double v = ((Rq[d * 3 + 0] * f + Rq[d * 3 + 1] * X0) + Rq[d * 3 + 2] * X1) + mean[d];
                                o[d] = (float)v;
                            }
                            o[3] = 1.0f;
                            uint8_t* cb = out32 + 32 * (base + m) + 16;
                            // c = C_star.row(m) + RGB_means[i] (gp_compressor.cpp:367); without the field GP: mean only
                            double cf[3] = {0, 0, 0};
                            if (have_rgb) {
                                if (sep) { gc.grid_k(REx, sz, xx, REy, sz, yy); gc.predict_field_grid(cf); }
                                else gc.predict_field(X0, X1, cf);
                            }
                            cb[2] = (uint8_t)flatten_color(cf[0] + cm[0]); cb[1] = (uint8_t)flatten_color(cf[1] + cm[1]);
                            cb[0] = (uint8_t)flatten_color(cf[2] + cm[2]); cb[3] = 255;
                            std::memset(out32 + 32 * (base + m) + 20, 0, 12);
                        }
                    }
            }
            double old = sink.load();
            while (!sink.compare_exchange_weak(old, old + acc)) {}
        });
        t_decode = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return slot[NP] * g2;
    }
};

}  // namespace

// ======================================================================================
// C API (ctypes)
// ======================================================================================
extern "C" {

struct orc_config {
    double res; int sz; int capacity; double s0; double eps_tol; double sigmaf_sq; double l_sq;
    int leaf_order; int shuffle; int rgb_rand; int threads;
    int rgb; int ref_order; double rgb_s0; double rgb_eps_tol;
    int decode_separable; int pad2;
};

double orc_exp(double x) { return orc_exp_impl(x); }
void orc_exp_array(const double* x, double* out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = orc_exp_impl(x[i]); }

void orc_rand_stream(uint64_t offset, int64_t n, uint32_t* out) {
    GlibcRand g;
    for (uint64_t t = 0; t < offset; t++) g.next();
    for (int64_t i = 0; i < n; i++) out[i] = g.next();
}
// sequence of shuffles of the given sizes starting at stream offset `offset`; out is the concatenation
void orc_shuffles(uint64_t offset, const int32_t* sizes, int64_t count, int32_t* out) {
    GlibcRand g;
    for (uint64_t t = 0; t < offset; t++) g.next();
    std::vector<int> ind;
    int64_t o = 0;
    for (int64_t c = 0; c < count; c++) {
        shuffle_ref(ind, sizes[c], g);
        for (int i = 0; i < sizes[c]; i++) out[o++] = ind[i];
    }
}

void orc_config_default(orc_config* c) {
    Config d;
    c->res = d.res; c->sz = d.sz; c->capacity = d.capacity; c->s0 = d.s0; c->eps_tol = d.eps_tol;
    c->sigmaf_sq = d.sigmaf_sq; c->l_sq = d.l_sq; c->leaf_order = d.leaf_order; c->shuffle = d.shuffle;
    c->rgb_rand = d.rgb_rand; c->threads = d.threads;
    c->rgb = d.rgb; c->ref_order = d.ref_order; c->rgb_s0 = d.rgb_s0; c->rgb_eps_tol = d.rgb_eps_tol;
    c->decode_separable = d.decode_separable; c->pad2 = 0;
}

void* orc_create(const orc_config* c) {
    Oracle* o = new Oracle();
    o->cfg.res = c->res; o->cfg.sz = c->sz; o->cfg.capacity = c->capacity; o->cfg.s0 = c->s0;
    o->cfg.eps_tol = c->eps_tol; o->cfg.sigmaf_sq = c->sigmaf_sq; o->cfg.l_sq = c->l_sq;
    o->cfg.leaf_order = c->leaf_order; o->cfg.shuffle = c->shuffle; o->cfg.rgb_rand = c->rgb_rand;
    o->cfg.threads = c->threads < 1 ? 1 : c->threads;
    o->cfg.rgb = c->rgb; o->cfg.ref_order = c->ref_order; o->cfg.rgb_s0 = c->rgb_s0; o->cfg.rgb_eps_tol = c->rgb_eps_tol;
    o->cfg.decode_separable = c->decode_separable;
    return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
const char* orc_last_error(void* h) { return ((Oracle*)h)->err.c_str(); }
void orc_set_rand_offset(void* h, uint64_t off) { ((Oracle*)h)->rand_offset = off; }
uint64_t orc_get_rand_offset(void* h) { return ((Oracle*)h)->rand_offset; }

// project_cloud only
int orc_project(void* h, const void* cloud32, int64_t n) { return ((Oracle*)h)->project((const uint8_t*)cloud32, n); }
// train on the projected stream (save_compressed = project + train)
int orc_train_projected(void* h, int dump) {
    Oracle* o = (Oracle*)h;
    return o->train(o->n_leaves, o->patch_off.data(), o->st_x1.data(), o->st_x2.data(), o->st_y.data(), dump, o->st_c.data());
}
int orc_compress(void* h, const void* cloud32, int64_t n, int dump) {
    int rc = orc_project(h, cloud32, n);
    if (rc) return rc;
    return orc_train_projected(h, dump);
}
// sparse_gp::add_measurements on caller-provided patch streams (frames stay undefined)
int orc_fit_patches(void* h, int64_t NP, const int64_t* off, const double* x1, const double* x2, const double* y, int dump) {
    Oracle* o = (Oracle*)h;
    o->leaf_R.clear();
    return o->train(NP, off, x1, x2, y, dump);
}
// sparse_gp::add_measurements again on the fitted patches (continues the kept state; both fits with dump)
int orc_add_measurements(void* h, int64_t NP, const int64_t* off, const double* x1, const double* x2, const double* y) {
    Oracle* o = (Oracle*)h;
    return o->train(NP, off, x1, x2, y, 1, nullptr, true);
}
// the same with per-point colours (3 per point, centred by the caller): also fits the RGB field GPs when cfg.rgb
int orc_fit_patches_rgb(void* h, int64_t NP, const int64_t* off, const double* x1, const double* x2, const double* y,
                        const double* colours, int dump) {
    Oracle* o = (Oracle*)h;
    o->leaf_R.clear();
    o->n_claimed = off[NP];
    return o->train(NP, off, x1, x2, y, dump, colours);
}
// inject fitted parameters (decode-only configurations)
int orc_set_params(void* h, int64_t NP, const int32_t* nbv, const double* bv1, const double* bv2, const double* alpha,
                   const double* quat, const double* mean, const double* rgbmean) {
    Oracle* o = (Oracle*)h;
    o->nbv.assign(nbv, nbv + NP);
    o->bv_off.assign(NP + 1, 0);
    for (int64_t p = 0; p < NP; p++) o->bv_off[p + 1] = o->bv_off[p] + nbv[p];
    int64_t T = o->bv_off[NP];
    o->bv1.assign(bv1, bv1 + T); o->bv2.assign(bv2, bv2 + T); o->alpha.assign(alpha, alpha + T);
    o->dumpC.clear(); o->dumpQ.clear(); o->dump_off.assign(NP + 1, 0);
    if (quat) {
        o->leaf_quat.assign(quat, quat + NP * 4); o->leaf_mean.assign(mean, mean + NP * 3);
        o->leaf_rgbmean.assign(rgbmean, rgbmean + NP * 3); o->leaf_R.assign(NP * 9, 0.0);
    } else {
        o->leaf_R.clear();
    }
    return 0;
}
int64_t orc_decode(void* h, void* out32, double* heights, int with_sigma) {
    return ((Oracle*)h)->decode_into((uint8_t*)out32, heights, with_sigma);
}
// predict at arbitrary local coordinates for one patch
// Batched evaluation over fitted patches (needs the C dump): out arrays may be null.  conf: 0 sigma, 1 confidence.
int orc_evaluate(void* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y, int conf,
                 double* f, double* sigma, double* lik, double* dX) {
    Oracle* o = (Oracle*)h;
    if (P < 0 || P > (int64_t)o->nbv.size()) return 1;
    Sogp gp;
    for (int64_t p = 0; p < P; p++) {
        const int N = o->nbv[p];
        if (N > 0 && o->dumpC.empty()) return 1;
        gp.init(o->sogp_params(), std::max(N, 1));
        gp.N = N;
        for (int i = 0; i < N; i++) {
            gp.alpha[i] = o->alpha[o->bv_off[p] + i]; gp.b1[i] = o->bv1[o->bv_off[p] + i]; gp.b2[i] = o->bv2[o->bv_off[p] + i];
            for (int j = 0; j < N; j++) gp.c(i, j) = o->dumpC[o->dump_off[p] + (size_t)i * N + j];
        }
        for (int64_t t = off[p]; t < off[p + 1]; t++) {
            double r[7];
            gp.evaluate(x1[t], x2[t], y ? y[t] : 0.0, r);
            if (f) f[t] = r[0];
            if (sigma) sigma[t] = conf ? r[2] : r[1];
            if (lik) lik[t] = r[3];
            if (dX) { dX[3 * t] = r[4]; dX[3 * t + 1] = r[5]; dX[3 * t + 2] = r[6]; }
        }
    }
    return 0;
}

// The RGB field GPs of the previous fit (with colours and dump): y = 3 values per point; f = 3 per point, dX = 3 per point
int orc_evaluate_rgb(void* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y3, int conf,
                     double* f3, double* sigma, double* lik, double* dX) {
    Oracle* o = (Oracle*)h;
    if (P < 0 || P > (int64_t)o->rgb_nbv.size()) return 1;
    Sogp gp;
    const double zero3[3] = {0, 0, 0};
    for (int64_t p = 0; p < P; p++) {
        const int N = o->rgb_nbv[p];
        if (N > 0 && o->rgb_dumpC.empty()) return 1;
        gp.init(o->rgb_params(), std::max(N, 1), 3);
        gp.N = N;
        for (int i = 0; i < N; i++) {
            gp.b1[i] = o->rgb_bv1[o->rgb_bv_off[p] + i]; gp.b2[i] = o->rgb_bv2[o->rgb_bv_off[p] + i];
            for (int ch = 0; ch < 3; ch++) gp.al(ch, i) = o->rgb_alpha[3 * (o->rgb_bv_off[p] + i) + ch];
            for (int j = 0; j < N; j++) gp.c(i, j) = o->rgb_dumpC[o->rgb_dump_off[p] + (size_t)i * N + j];
        }
        for (int64_t t = off[p]; t < off[p + 1]; t++) {
            double r[9];
            gp.evaluate_field(x1[t], x2[t], y3 ? &y3[3 * t] : zero3, r);
            if (f3) { f3[3 * t] = r[0]; f3[3 * t + 1] = r[1]; f3[3 * t + 2] = r[2]; }
            if (sigma) sigma[t] = conf ? r[4] : r[3];
            if (lik) lik[t] = r[5];
            if (dX) { dX[3 * t] = r[6]; dX[3 * t + 1] = r[7]; dX[3 * t + 2] = r[8]; }
        }
    }
    return 0;
}

int orc_predict(void* h, int64_t patch, const double* X, int64_t m, double* f, double* sigma) {
    Oracle* o = (Oracle*)h;
    if (patch < 0 || patch >= (int64_t)o->nbv.size()) return 1;
    Sogp gp;
    int N = o->nbv[patch];
    gp.init(o->sogp_params(), std::max(N, 1));
    gp.N = N;
    for (int i = 0; i < N; i++) {
        gp.alpha[i] = o->alpha[o->bv_off[patch] + i]; gp.b1[i] = o->bv1[o->bv_off[patch] + i]; gp.b2[i] = o->bv2[o->bv_off[patch] + i];
    }
    if (sigma && !o->dumpC.empty())
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) gp.c(i, j) = o->dumpC[o->dump_off[patch] + (size_t)i * N + j];
    for (int64_t t = 0; t < m; t++) {
        if (gp.P.ref_order) {  // sparse_gp.hpp:312-351 as the reference evaluates it
            f[t] = gp.predict_ref(X[2 * t], X[2 * t + 1], sigma ? sigma + t : nullptr);
            continue;
        }
        f[t] = gp.predict(X[2 * t], X[2 * t + 1]);
        if (sigma) sigma[t] = gp.predict_sigma(X[2 * t], X[2 * t + 1]);
    }
    return 0;
}

// ---- getters: sizes then raw pointers (valid until the next call on the handle) ----
struct orc_sizes { int64_t n_in, n_leaves, n_claimed, n_bv_total, n_dump; uint32_t depth; uint32_t pad; double mn[3]; };
void orc_get_sizes(void* h, orc_sizes* s) {
    Oracle* o = (Oracle*)h;
    s->n_in = o->n_in; s->n_leaves = o->n_leaves; s->n_claimed = o->n_claimed;
    s->n_bv_total = o->bv_off.empty() ? 0 : o->bv_off.back();
    s->n_dump = o->dump_off.empty() ? 0 : o->dump_off.back();
    s->depth = o->lat.depth; s->pad = 0;
    for (int a = 0; a < 3; a++) s->mn[a] = o->lat.mn[a];
}
const void* orc_ptr(void* h, const char* name) {
    Oracle* o = (Oracle*)h;
    std::string n(name);
    if (n == "leaf_code") return o->leaf_code.data();
    if (n == "leaf_center") return o->leaf_center.data();
    if (n == "leaf_ncand") return o->leaf_ncand.data();
    if (n == "leaf_R") return o->leaf_R.data();
    if (n == "leaf_quat") return o->leaf_quat.data();
    if (n == "leaf_mean") return o->leaf_mean.data();
    if (n == "leaf_rgbmean") return o->leaf_rgbmean.data();
    if (n == "patch_off") return o->patch_off.data();
    if (n == "owner") return o->owner.data();
    if (n == "st_idx") return o->st_idx.data();
    if (n == "st_x1") return o->st_x1.data();
    if (n == "st_x2") return o->st_x2.data();
    if (n == "st_y") return o->st_y.data();
    if (n == "st_perm") return o->st_perm.data();
    if (n == "nbv") return o->nbv.data();
    if (n == "bv_off") return o->bv_off.data();
    if (n == "bv_idx") return o->bv_idx.data();
    if (n == "bv1") return o->bv1.data();
    if (n == "bv2") return o->bv2.data();
    if (n == "alpha") return o->alpha.data();
    if (n == "fit_flags") return o->fit_flags.data();
    if (n == "dumpC") return o->dumpC.data();
    if (n == "dumpQ") return o->dumpQ.data();
    if (n == "dump_off") return o->dump_off.data();
    if (n == "st_c") return o->st_c.data();
    if (n == "st_perm_rgb") return o->st_perm_rgb.data();
    if (n == "rgb_nbv") return o->rgb_nbv.data();
    if (n == "rgb_bv_off") return o->rgb_bv_off.data();
    if (n == "rgb_bv_idx") return o->rgb_bv_idx.data();
    if (n == "rgb_bv1") return o->rgb_bv1.data();
    if (n == "rgb_bv2") return o->rgb_bv2.data();
    if (n == "rgb_alpha") return o->rgb_alpha.data();
    return nullptr;
}
struct orc_stats {
    uint64_t n_add, n_first, n_sparse, n_full, n_del_cap, n_del_geo;
    double sumN, sumN2_common, sumN2_sparse, sumN2_full, sumN2_del;
    double t_project, t_train, t_decode;
};
void orc_get_stats(void* h, orc_stats* s) {
    Oracle* o = (Oracle*)h;
    s->n_add = o->stats.n_add; s->n_first = o->stats.n_first; s->n_sparse = o->stats.n_sparse; s->n_full = o->stats.n_full;
    s->n_del_cap = o->stats.n_del_cap; s->n_del_geo = o->stats.n_del_geo; s->sumN = o->stats.sumN;
    s->sumN2_common = o->stats.sumN2_common; s->sumN2_sparse = o->stats.sumN2_sparse; s->sumN2_full = o->stats.sumN2_full;
    s->sumN2_del = o->stats.sumN2_del; s->t_project = o->t_project; s->t_train = o->t_train; s->t_decode = o->t_decode;
}

void orc_get_stats_rgb(void* h, orc_stats* s) {
    Oracle* o = (Oracle*)h;
    std::memset(s, 0, sizeof(*s));
    s->n_add = o->stats_rgb.n_add; s->n_first = o->stats_rgb.n_first; s->n_sparse = o->stats_rgb.n_sparse; s->n_full = o->stats_rgb.n_full;
    s->n_del_cap = o->stats_rgb.n_del_cap; s->n_del_geo = o->stats_rgb.n_del_geo; s->sumN = o->stats_rgb.sumN;
    s->sumN2_common = o->stats_rgb.sumN2_common; s->sumN2_sparse = o->stats_rgb.sumN2_sparse; s->sumN2_full = o->stats_rgb.sumN2_full;
    s->sumN2_del = o->stats_rgb.sumN2_del;
}
int64_t orc_rgb_bv_total(void* h) { Oracle* o = (Oracle*)h; return o->rgb_bv_off.empty() ? 0 : o->rgb_bv_off.back(); }

// helpers exported for unit tests
void orc_rotation_from_sums(const double sums[10], const double c[3], double R[9]) {
    GramSums g;
    for (int i = 0; i < 10; i++) g.s[i] = sums[i];
    double v[4];
    smallest_right_singular_vector(g, c, v);
    double nrm[3] = {v[0], v[1], v[2]};
    rotation_from_normal(nrm, R);
}
void orc_quat_roundtrip(const double R[9], double q[4], double R2[9]) { rot_to_quat(R, q); quat_to_rot(q, R2); }
void orc_lattice(const void* cloud32, int64_t n, double res, double mn[3], uint32_t* depth) {
    Lattice L;
    L.res = res;
    const uint8_t* c = (const uint8_t*)cloud32;
    for (int64_t i = 0; i < n; i++) {
        const float* p = reinterpret_cast<const float*>(c + 32 * i);
        if (finite3(p)) lattice_add_point(L, p);
    }
    for (int a = 0; a < 3; a++) mn[a] = L.mn[a];
    *depth = L.depth;
}

}  // extern "C"
