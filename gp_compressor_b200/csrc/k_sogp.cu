// k_sogp.cu — K7: sparse online GP fit, one CTA per patch, state in shared memory.
//
// Replaces sparse_gp<rbf_kernel,gaussian_noise>::add / delete_bv
// (/root/reference/src/sparse_gp.hpp:89-295) driven by add_measurements (:59-86) and
// train_processes (gp_compressor.cpp:121-175).  The point stream of a patch is strictly
// sequential (each add depends on the previous state); the parallelism is over patches
// (grid) and over the N x N state (threads of the CTA).
//
// Shared-memory state per CTA: C and Q (ld x ld doubles each, bitwise symmetric, so
// "row i" is read as the contiguous column i), alpha, BV, and the work vectors.
// Thread mapping with NT = 2*RB threads: for the two matvecs thread t computes row
// (t % RB) of matrix (t / RB) with the canonical 4-partial order; for rank-1/rank-2
// updates thread t owns row (t % RB) and the columns j = (t / RB), (t / RB) + 2, ...
// The three dot products use the canonical 32-lane order and are computed redundantly by
// every warp (no broadcast barrier).  Buckets: RB = 16/32/64/128; a patch whose BV count
// outgrows its bucket is pushed to the next bucket's queue and refitted there.
#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace gpc {

namespace {

template <int NT>
__device__ __forceinline__ void cta_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

struct Smem {
    double *C, *Q, *alpha, *b1, *b2, *kv, *ck, *ev, *sv, *qsv, *qcv;
    int* bidx;
};

__device__ __forceinline__ Smem carve(double* base, int ld) {
    Smem s;
    s.C = base;
    s.Q = s.C + ld * ld;
    s.alpha = s.Q + ld * ld;
    s.b1 = s.alpha + ld;
    s.b2 = s.b1 + ld;
    s.kv = s.b2 + ld;
    s.ck = s.kv + ld;
    s.ev = s.ck + ld;
    s.sv = s.ev + ld;
    s.qsv = s.sv + ld;
    s.qcv = s.qsv + ld;
    s.bidx = reinterpret_cast<int*>(s.qcv + ld);
    return s;
}

// argmin with the reference's "first strict minimum" scan (sparse_gp.hpp:210-217,230-236).
// MODE 0: score_i = alpha_i^2 / (Q_ii + C_ii);  MODE 1: score_i = 1 / Q_ii.
// Every lane returns the same (loc, score).  A NaN score never wins unless it is score_0,
// in which case nothing is ever "< minscore" and loc stays 0 with a NaN minimum.
template <int MODE>
__device__ __forceinline__ int warp_argmin(const Smem& s, int ld, int N, int lane, double* minscore) {
    double best = 0.0;
    int bi = 0x7fffffff;
    int nan0 = 0;
    for (int i = lane; i < N; i += 32) {
        double qii = s.Q[i * ld + i];
        double sc;
        if (MODE == 0) {
            double a = s.alpha[i];
            sc = __ddiv_rn(__dmul_rn(a, a), __dadd_rn(qii, s.C[i * ld + i]));
        } else {
            sc = __ddiv_rn(1.0, qii);
        }
        if (sc != sc) {
            if (i == 0) nan0 = 1;
            continue;
        }
        if (bi == 0x7fffffff || sc < best) { best = sc; bi = i; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        double ob = shfl_xor_d(best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob < best || (ob == best && oi < bi))) { best = ob; bi = oi; }
    }
    nan0 = __shfl_sync(0xffffffffu, nan0, 0);
    if (nan0) {
        *minscore = __longlong_as_double(0x7ff8000000000000LL);
        return 0;
    }
    *minscore = best;
    return bi;
}

// sparse_gp::delete_bv, sparse_gp.hpp:252-295.  Uniform across the CTA; ends with a barrier.
template <int RB, int NT>
__device__ __forceinline__ void delete_bv(const Smem& s, int ld, int& N, int loc, int t) {
    const int L = N - 1, M = N - 1;
    const int irow = t & (RB - 1), jg = t / RB;
    double csi = 0, qsi = 0, repc = 0, repq = 0, ai = 0, nb1 = 0, nb2 = 0;
    int nidx = -1;
    if (t < N) {
        const int i = t;
        const int src = (i == loc) ? L : i;  // Cs(loc) = Cs(L), Crep(loc) = Crep(L), alpha(loc) = alpha(L)
        csi = s.C[loc * ld + src];
        qsi = s.Q[loc * ld + src];
        repc = s.C[L * ld + src];
        repq = s.Q[L * ld + src];
        ai = s.alpha[src];
        if (i == loc) { nb1 = s.b1[L]; nb2 = s.b2[L]; nidx = s.bidx[L]; }
    }
    const double cstar = s.C[loc * ld + loc], qstar = s.Q[loc * ld + loc], astar = s.alpha[loc];
    cta_sync<NT>();
    const double qcs = __dadd_rn(qstar, cstar);
    const double coef = __ddiv_rn(astar, qcs);
    const double iq = __ddiv_rn(1.0, qstar), iqc = __ddiv_rn(1.0, qcs);
    if (t < N) {
        const int i = t;
        if (i < M) {
            if (loc != L) {
                s.C[loc * ld + i] = repc; s.C[i * ld + loc] = repc;
                s.Q[loc * ld + i] = repq; s.Q[i * ld + loc] = repq;
                if (i == loc) { s.b1[loc] = nb1; s.b2[loc] = nb2; s.bidx[loc] = nidx; }
            }
            double qci = __dadd_rn(qsi, csi);
            s.alpha[i] = __dadd_rn(ai, -__dmul_rn(coef, qci));
            s.qsv[i] = qsi;
            s.qcv[i] = qci;
        }
        s.C[L * ld + i] = 0.0; s.C[i * ld + L] = 0.0;
        s.Q[L * ld + i] = 0.0; s.Q[i * ld + L] = 0.0;
        if (i == L) { s.alpha[L] = 0.0; s.b1[L] = 0.0; s.b2[L] = 0.0; s.bidx[L] = -1; }
    }
    cta_sync<NT>();
    if (irow < M) {
        const double qi = s.qsv[irow], ci = s.qcv[irow];
        for (int j = jg; j < M; j += 2) {
            const int idx = j * ld + irow;
            double u = __dmul_rn(qi, s.qsv[j]);
            double v = __dmul_rn(ci, s.qcv[j]);
            double tt = __dmul_rn(v, iqc);
            double w = fma(u, iq, -tt);
            s.C[idx] = __dadd_rn(s.C[idx], w);
            s.Q[idx] = fma(-u, iq, s.Q[idx]);
        }
    }
    N = M;
    cta_sync<NT>();
}

template <int RB, int NT>
__global__ void __launch_bounds__(NT) sogp_fit_kernel(SogpArgs a) {
    extern __shared__ double smem_d[];
    const int ld = a.ld;
    const Smem s = carve(smem_d, ld);
    const int t = threadIdx.x, lane = t & 31;
    const int irow = t & (RB - 1), jg = t / RB;
    const int64_t patch = a.patch_ids ? (int64_t)a.patch_ids[blockIdx.x] : a.first_patch + blockIdx.x;
    const int64_t o = a.off[patch];
    const int n = (int)(a.off[patch + 1] - o);
    const int64_t op = patch - a.out_first;
    if (n == 0) {
        if (t == 0) { a.nbv[op] = 0; a.flags[op] = 0; }
        return;
    }
    for (int i = t; i < 2 * ld * ld + 9 * ld; i += NT) smem_d[i] = 0.0;
    for (int i = t; i < ld; i += NT) s.bidx[i] = -1;
    cta_sync<NT>();

    const double kstar = a.p0, s20 = a.s20, p0 = a.p0, cl = a.cl, eps_tol = a.eps_tol;
    const int cap = a.capacity;
    int N = 0;
    unsigned long long c_first = 0, c_sparse = 0, c_full = 0, c_dcap = 0, c_dgeo = 0;
    unsigned long long c_sumn = 0, c_n2c = 0, c_n2s = 0, c_n2f = 0, c_n2d = 0;

    double nx1 = a.fx1[o], nx2 = a.fx2[o], ny = a.fy[o];
    int norig = a.forig[o];
    for (int tt = 0; tt < n; ++tt) {
        const double x1 = nx1, x2 = nx2, y = ny;
        const int orig = norig;
        if (tt + 1 < n) {  // prefetch the next point of the stream
            nx1 = a.fx1[o + tt + 1]; nx2 = a.fx2[o + tt + 1]; ny = a.fy[o + tt + 1];
            norig = a.forig[o + tt + 1];
        }
        if (N == 0) {  // sparse_gp.hpp:100-110
            if (t == 0) {
                double d = __dadd_rn(kstar, s20);
                s.alpha[0] = __ddiv_rn(y, d);
                s.C[0] = __ddiv_rn(-1.0, d);
                s.Q[0] = __ddiv_rn(1.0, kstar);
                s.b1[0] = x1; s.b2[0] = x2; s.bidx[0] = orig;
            }
            N = 1;
            c_first++;
            cta_sync<NT>();
            continue;
        }
        c_sumn += N;
        c_n2c += (unsigned long long)N * N;
        // k = K(x, BV)  (sparse_gp.hpp:119)
        if (t < N) s.kv[t] = rbf(x1, x2, s.b1[t], s.b2[t], p0, cl);
        cta_sync<NT>();
        // C k and e_hat = Q k (:122,:140), m = alpha' k (:121)
        if (irow < N) {
            const double* Mx = jg ? s.Q : s.C;
            double r = row4(Mx + irow, ld, s.kv, N);
            (jg ? s.ev : s.ck)[irow] = r;
        }
        const double m = warp_dot32(s.alpha, s.kv, N, lane);
        cta_sync<NT>();
        double kck = 0.0, ke = 0.0;
        for (int j = lane; j < N; j += 32) {
            double kj = s.kv[j];
            kck = fma(kj, s.ck[j], kck);
            ke = fma(kj, s.ev[j], ke);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            kck = __dadd_rn(kck, shfl_xor_d(kck, off));
            ke = __dadd_rn(ke, shfl_xor_d(ke, off));
        }
        const double s2 = __dadd_rn(kstar, kck);
        const double den = __dadd_rn(s20, s2);
        const double r = __ddiv_rn(-1.0, den);                // gaussian_noise.cpp:15-18
        const double q = __ddiv_rn(__dadd_rn(y, -m), den);    // gaussian_noise.cpp:9-12
        double gamma = __dadd_rn(kstar, -ke);                 // sparse_gp.hpp:144
        if (gamma < tiny12()) gamma = 0.0;
        if (gamma < eps_tol) {
            // sparse update (sparse_gp.hpp:155-163)
            c_sparse++;
            c_n2s += (unsigned long long)N * N;
            const double eta = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(gamma, r)));
            if (t < N) {
                double sh = __dadd_rn(s.ck[t], s.ev[t]);
                s.sv[t] = sh;
                s.alpha[t] = __dadd_rn(s.alpha[t], __dmul_rn(sh, __dmul_rn(q, eta)));
            }
            cta_sync<NT>();
            const double re = __dmul_rn(r, eta);
            if (irow < N) {
                const double si = s.sv[irow];
                for (int j = jg; j < N; j += 2) {
                    const int idx = j * ld + irow;
                    s.C[idx] = fma(re, __dmul_rn(si, s.sv[j]), s.C[idx]);
                }
            }
            continue;  // Q and N unchanged: neither deletion loop can fire
        }
        // full update (sparse_gp.hpp:164-203)
        if (N + 1 > ld) {  // does not fit this bucket: hand the patch to the next one
            if (t == 0) {
                int pos = atomicAdd(a.queue_count, 1);
                a.queue[pos] = (int32_t)patch;
            }
            return;
        }
        c_full++;
        c_n2f += (unsigned long long)(N + 1) * (N + 1);
        if (t < N) {
            double sc = s.ck[t];
            s.sv[t] = sc;
            s.alpha[t] = __dadd_rn(s.alpha[t], __dmul_rn(q, sc));
        }
        if (t == N) {
            s.sv[N] = 1.0;
            s.alpha[N] = __dadd_rn(0.0, __dmul_rn(q, 1.0));
            s.ev[N] = -1.0;
            s.b1[N] = x1; s.b2[N] = x2; s.bidx[N] = orig;
        }
        cta_sync<NT>();
        {
            const double ig = __ddiv_rn(1.0, gamma);
            const int N1 = N + 1;
            if (irow < N1) {
                const double si = s.sv[irow], ei = s.ev[irow];
                for (int j = jg; j < N1; j += 2) {
                    const int idx = j * ld + irow;
                    s.C[idx] = fma(r, __dmul_rn(si, s.sv[j]), s.C[idx]);
                    s.Q[idx] = fma(ig, __dmul_rn(ei, s.ev[j]), s.Q[idx]);
                }
            }
            N = N1;
        }
        cta_sync<NT>();
        // capacity deletions (sparse_gp.hpp:206-223)
        while (N > cap) {
            double ms;
            int loc = warp_argmin<0>(s, ld, N, lane, &ms);
            c_n2d += (unsigned long long)(N - 1) * (N - 1);
            delete_bv<RB, NT>(s, ld, N, loc, t);
            c_dcap++;
        }
        // geometric deletions (sparse_gp.hpp:226-242)
        {
            double minscore = 0.0;
            while (minscore < geo9() && N > 1) {
                int loc = warp_argmin<1>(s, ld, N, lane, &minscore);
                if (minscore < geo9()) {
                    c_n2d += (unsigned long long)(N - 1) * (N - 1);
                    delete_bv<RB, NT>(s, ld, N, loc, t);
                    c_dgeo++;
                }
            }
        }
    }
    cta_sync<NT>();
    // results
    if (t == 0) {
        a.nbv[op] = N;
        double c00 = s.C[0];
        a.flags[op] = (c00 != c00) ? 1 : 0;
        unsigned long long* st = a.stats;
        atomicAdd(st + 0, (unsigned long long)n);
        atomicAdd(st + 1, c_first); atomicAdd(st + 2, c_sparse); atomicAdd(st + 3, c_full);
        atomicAdd(st + 4, c_dcap); atomicAdd(st + 5, c_dgeo); atomicAdd(st + 6, c_sumn);
        atomicAdd(st + 7, c_n2c); atomicAdd(st + 8, c_n2s); atomicAdd(st + 9, c_n2f); atomicAdd(st + 10, c_n2d);
    }
    const int64_t ob = op * cap;
    for (int i = t; i < N; i += NT) {
        a.o_alpha[ob + i] = s.alpha[i];
        a.o_b1[ob + i] = s.b1[i];
        a.o_b2[ob + i] = s.b2[i];
        a.o_idx[ob + i] = s.bidx[i];
    }
    if (a.dumpC) {
        const int64_t od = op * (int64_t)cap * cap;
        for (int e = t; e < N * N; e += NT) {
            int i = e / N, j = e - i * N;
            a.dumpC[od + e] = s.C[j * ld + i];
            a.dumpQ[od + e] = s.Q[j * ld + i];
        }
    }
}

}  // namespace

int sogp_bucket_ld(int bucket) {
    static const int lds[4] = {16, 32, 64, 118};
    return lds[bucket];
}

size_t sogp_smem_bytes(int ld) { return (size_t)(2 * ld * ld + 9 * ld) * sizeof(double) + (size_t)ld * sizeof(int); }

template <int RB, int NT>
static cudaError_t launch_bucket(const SogpArgs& a, cudaStream_t st) {
    size_t smem = sogp_smem_bytes(a.ld);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(sogp_fit_kernel<RB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    sogp_fit_kernel<RB, NT><<<a.n_work, NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_sogp_fit(int bucket, const SogpArgs& a, cudaStream_t st) {
    if (a.n_work <= 0) return cudaSuccess;
    switch (bucket) {
        case 0: return launch_bucket<16, 32>(a, st);
        case 1: return launch_bucket<32, 64>(a, st);
        case 2: return launch_bucket<64, 128>(a, st);
        default: return launch_bucket<128, 256>(a, st);
    }
}

}  // namespace gpc
