// k_sogp.cu — K7: sparse online GP fit, one CTA per patch, state in shared memory.
//
// Replaces sparse_gp<rbf_kernel,gaussian_noise>::add / delete_bv
// (/root/reference/src/sparse_gp.hpp:89-295) driven by add_measurements (:59-86) and
// train_processes (gp_compressor.cpp:121-175).  The point stream of a patch is strictly
// sequential (each add depends on the previous state); the parallelism is over patches
// (grid) and over the N x N state (threads of the CTA).
//
// State per CTA in shared memory: C and Q (LD x LD doubles each, bitwise symmetric, so
// "row i" is read as the contiguous column i: conflict-free), plus work vectors.
//
// Bucket 0 (N + 1 <= 16): TWO PATCHES per warp, one per half-warp.  Lane r of a half-warp owns index r of
// alpha / BV / k and row r of C in registers, and row r of Q in shared memory; the three dot products are one product
// per lane and a 4-step butterfly; patches are visited by decreasing size so that the two halves of a warp have
// (almost) equal streams.  No block barrier, only __syncwarp; the point stream arrives by bulk-async (TMA) tile copies.
// Bucket 1 (N + 1 <= 32): two warps per patch, warp 0 owns C and warp 1 owns Q (sogp_fit_pair_kernel).
// Buckets 2-4 (N + 1 <= 64 / 104 / 202), height GP: sogp_fit_fused_kernel -- the update of a point and the two matvecs
// of the next one share ONE pass over C and Q (in shared memory; in a global-memory slice for bucket 4), and a capacity
// deletion is folded into that pass.  The step-by-step kernel sogp_fit_kernel (NT = 4*RB threads, a warp covers 16
// rows for the matvec, a thread one row and every fourth column for the updates) serves the RGB field GP (three
// outputs) and capacity 104..117.
//
// Every patch starts in bucket 0.  When a full update would not fit the bucket, the CTA
// writes its complete state (N, next point, counters, alpha, BV, C, Q) to a hand-off slot and
// queues the patch; the next bucket's kernel resumes from that state — no work is redone
// and the arithmetic is the same in every bucket.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace gpc {

namespace {

constexpr int NCNT = 10;  // first sparse full dcap dgeo sumn n2c n2s n2f n2d

// hand-off slot: header (2) | counters | alpha[dout][ld] | b1[ld] | b2[ld] | C[ld*ld] | Q[ld*ld] | bidx[ld] (ints)
__host__ __device__ constexpr int slot_doubles(int ld, int dout = 1) { return 2 + NCNT + (2 + dout) * ld + 2 * ld * ld + (ld + 1) / 2; }

struct Counters {
    unsigned long long c[NCNT];
    unsigned int run;  // sparse points seen at the current N, not yet folded into c[]
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < NCNT; i++) c[i] = 0;
        run = 0;
    }
    __device__ __forceinline__ void flush(int N) {
        const unsigned long long n2 = (unsigned long long)N * N;
        c[1] += run; c[5] += (unsigned long long)run * N; c[6] += run * n2; c[7] += run * n2;
        run = 0;
    }
    __device__ __forceinline__ void full(int N) {
        c[2]++; c[5] += N; c[6] += (unsigned long long)N * N; c[8] += (unsigned long long)(N + 1) * (N + 1);
    }
};

__device__ __forceinline__ void publish(const SogpArgs& a, const Counters& k, int n) {
    unsigned long long* st = a.stats;
    atomicAdd(st + 0, (unsigned long long)n);
#pragma unroll
    for (int i = 0; i < NCNT; i++)
        if (k.c[i]) atomicAdd(st + 1 + i, k.c[i]);
}

// argmin with the reference's "first strict minimum" scan (sparse_gp.hpp:210-217, 230-236) over one
// score per lane (idx = the element the lane holds, 0x7fffffff if none).  A NaN score never wins unless
// it is element 0, in which case nothing is ever "< minscore": loc stays 0 and the minimum is NaN.
// Min by butterfly on the value alone, then the lowest index among the lanes that hold that value.
__device__ __forceinline__ int warp_first_min(double sc, int idx, double* minscore) {
    const bool valid = idx != 0x7fffffff;
    const bool isn = sc != sc;
    const int nan0 = __any_sync(0xffffffffu, valid && isn && idx == 0);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double v = (valid && !isn) ? sc : inf;
    double m = v;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmin(m, shfl_xor_d(m, off));
    int cand = (valid && !isn && v == m) ? idx : 0x7fffffff;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, off));
    if (nan0) {
        *minscore = __longlong_as_double(0x7ff8000000000000LL);
        return 0;
    }
    *minscore = m;
    return cand;
}

// =====================================================================================
// Bucket 0 layout: rows are contiguous (row r at r * W_LD, W_LD = 18: the two pad columns keep
// 16-byte alignment and make the 128-bit row loads of a quarter-warp hit distinct banks).
// C and Q are bitwise symmetric, so row r == column r.
// =====================================================================================
constexpr int W_N = 16;    // N + 1 <= 16
constexpr int W_LD = 18;

template <int DOUT>
struct __align__(16) StagedPointT { double x1, x2, y[DOUT]; int orig, pad; };
typedef StagedPointT<1> StagedPoint;

// canonical row4 over a contiguous row with 128-bit loads: identical operation order
__device__ __forceinline__ double row4_contig(const double* row, const double* k, int n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int j = 0;
    for (; j + 3 < n; j += 4) {
        const double2 m01 = *reinterpret_cast<const double2*>(row + j), m23 = *reinterpret_cast<const double2*>(row + j + 2);
        const double2 k01 = *reinterpret_cast<const double2*>(k + j), k23 = *reinterpret_cast<const double2*>(k + j + 2);
        a0 = fma(m01.x, k01.x, a0);
        a1 = fma(m01.y, k01.y, a1);
        a2 = fma(m23.x, k23.x, a2);
        a3 = fma(m23.y, k23.y, a3);
    }
    if (j < n) {
        const double2 m01 = *reinterpret_cast<const double2*>(row + j), k01 = *reinterpret_cast<const double2*>(k + j);
        a0 = fma(m01.x, k01.x, a0);
        if (j + 1 < n) a1 = fma(m01.y, k01.y, a1);
        if (j + 2 < n) a2 = fma(row[j + 2], k[j + 2], a2);
    }
    return __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
}

// =====================================================================================
// Bucket 0, packed: TWO PATCHES per warp, one per half-warp.  With N + 1 <= 16 a full warp per patch leaves most
// lanes of every scalar step (exp, divisions, butterflies) idle; here lane r of a half-warp owns index r of
// alpha / BV / k AND row r of both C and Q, so the per-point instruction stream serves two patches.  The halves run
// independently (own masks for every shuffle / sync, own control flow); they share instructions whenever they are
// on the same path, which is the sparse update almost always.  Same arithmetic and operation order as every bucket:
// the 16 partial products of a dot reduce with a 4-step butterfly (the canonical dot32 with zeros in partials 16-31).
// =====================================================================================
__device__ __forceinline__ double shfl16_xor(unsigned gm, double v, int off) { return __shfl_xor_sync(gm, v, off, 16); }
__device__ __forceinline__ double shfl16(unsigned gm, double v, int src) { return __shfl_sync(gm, v, src, 16); }

// first strict minimum over the 16 lanes of a half-warp (see warp_first_min)
__device__ __forceinline__ int half_first_min(unsigned gm, double sc, int idx, double* minscore) {
    const bool valid = idx != 0x7fffffff;
    const bool isn = sc != sc;
    const int nan0 = __any_sync(gm, valid && isn && idx == 0);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double v = (valid && !isn) ? sc : inf;
    double m = v;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) m = fmin(m, shfl16_xor(gm, m, off));
    int cand = (valid && !isn && v == m) ? idx : 0x7fffffff;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) cand = min(cand, __shfl_xor_sync(gm, cand, off, 16));
    if (nan0) {
        *minscore = __longlong_as_double(0x7ff8000000000000LL);
        return 0;
    }
    *minscore = m;
    return cand;
}

// ---- 1-D bulk-async copies (the TMA unit without a tensor map: UBLKCP in SASS) completing on an mbarrier --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// One staged tile of the fit stream: 16 consecutive points of a patch.  Every array is filled by ONE bulk copy from the
// 16-byte aligned address at or below the tile's first element (the copy engine needs 16-byte alignment and sizes), so
// element j of the tile sits at index j + shift, shift = (o & 1) for the doubles and (o & 3) for the indices.
template <int DOUT>
struct __align__(16) HalfTile {
    double x1[18], x2[18], y[DOUT][18];
    int orig[20];
};
constexpr int TILE_PAD_BYTES = 256;  // the stream buffers end with at least this much slack (gpc_api.cu): tiles read past the end

template <int DOUT>
struct __align__(16) HalfSmem {
    double Q[W_N * W_LD];  // lane r reads and updates row r in place
    double kv[W_N], sv[W_N], ev[W_N];
    HalfTile<DOUT> tile[2];
    unsigned long long mbar[2];
};

// One warp = two patches.  C lives in registers; its home in shared memory is needed only by the rare paths (deletions,
// hand-off, final dump), so the two halves share ONE and take turns there: 2.3 KB less per warp, 20 warps per SM instead of 17.
template <int DOUT>
struct __align__(16) HalfWarpSmem {
    HalfSmem<DOUT> h[2];
    double Chome[W_N * W_LD];
};

template <int DOUT>
__device__ __forceinline__ void issue_tile(HalfTile<DOUT>& t, unsigned long long* bar, const SogpArgs& a, int64_t pos) {
    mbar_expect_tx(bar, (uint32_t)((2 + DOUT) * 144 + 80));
    const int64_t pd = pos & ~(int64_t)1, pi = pos & ~(int64_t)3;
    bulk_g2s(t.x1, a.fx1 + pd, 144, bar);
    bulk_g2s(t.x2, a.fx2 + pd, 144, bar);
#pragma unroll
    for (int c = 0; c < DOUT; c++) bulk_g2s(t.y[c], a.fy[c] + pd, 144, bar);
    bulk_g2s(t.orig, a.forig + pi, 80, bar);
}

constexpr int HALF_BLOCKS = 20;   // 96 registers; 20 x (shared memory of one warp + 1 KB) <= 227 KB

template <int V>
struct IntC { static constexpr int value = V; };

// DOUT = 1: sparse_gp (heights).  DOUT = 3: sparse_gp_field (RGB): alpha has three columns, the capacity score is
// |alpha_i|^2 / (Q_ii + C_ii) (sparse_gp_field.hpp:187) and delete_bv updates alpha with alphastar * ((q*+c*)(Qs+Cs)) (:250-253).
//
// Structure (sparse_gp.hpp:119-203 is the loop; the arithmetic and its order are those of every other bucket):
//  * Lane r keeps row r of C in registers and reads / updates row r of Q in place in shared memory (Q changes only at full
//    updates, 7 % of the points under the reference hyper-parameters; C changes at every point).  A sparse update (:155-163)
//    and a full update (:164-203) touch no other shared memory than the broadcasts of k, s and e_hat (16 doubles each): the
//    rank-1 updates run over the zero-padded width, so the new row / column of a full update needs no indexing.
//    96 registers and 4.9 KB of shared memory per patch: 20 warps (40 patches) per SM.
//    The step is compiled once per number of occupied 4-column groups (N <= 8, 12, 16) so that it has no per-group
//    predicates, and runs converged for both halves of the warp (full-mask shuffles).
//  * Deletions (rare under the reference hyper-parameters) need full matrices: the rows of C are spilled to a home in shared
//    memory that the two halves of the warp SHARE and use in turns; the hand-off to the next bucket and the final dump are
//    written straight from the registers.
//  * k of the next point is computed while this point's chain runs (it depends only on the BV set).
//  * The point stream is staged 16 points at a time by bulk-async copies (TMA unit) into a double buffer, one tile ahead.
template <int DOUT>
__global__ void __launch_bounds__(32, HALF_BLOCKS) sogp_fit_half_kernel(SogpArgs a) {
    __shared__ __align__(16) HalfWarpSmem<DOUT> smw;
    const int lane = threadIdx.x;
    const int half = lane >> 4, r = lane & 15;
    const unsigned gm = 0xffffu << (16 * half);
    const unsigned FULL = 0xffffffffu;
    const int64_t w = 2 * (int64_t)blockIdx.x + half;
    HalfSmem<DOUT>& sm = smw.h[half];
    double* const C = smw.Chome;   // valid only inside take_turns()
    double* const Q = sm.Q;
    int64_t patch = 0, o = 0, op = 0;
    int n = 0;
    if (w < a.n_work) {
        patch = a.patch_ids ? (int64_t)a.patch_ids[w] : a.first_patch + w;
        o = a.off[patch];
        n = (int)(a.off[patch + 1] - o);
        op = patch - a.out_first;
        if (n == 0 && r == 0) { a.nbv[op] = 0; a.flags[op] = 0; }
    }
    // a half without work stays in the loop with n = 0: every warp-level operation of the step names all 32 lanes
#pragma unroll
    for (int i = 0; i < W_N * W_LD / 16; i++) Q[r + 16 * i] = 0.0;
    if (half == 0) {
#pragma unroll
        for (int i = 0; i < W_N * W_LD / 16; i++) C[r + 16 * i] = 0.0;
    }
    sm.kv[r] = 0.0; sm.sv[r] = 0.0; sm.ev[r] = 0.0;
    if (r == 0) { mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1); }
    __syncwarp();
    int issued = 0, waited = 0;  // tiles issued / consumed (half-uniform)
    if (n > 0) {
        if (r == 0) issue_tile<DOUT>(sm.tile[0], &sm.mbar[0], a, o);
        issued = 1;
    }
    const int sh1 = (int)(o & 1), sh3 = (int)(o & 3);
    const int nmax = max(n, __shfl_xor_sync(FULL, n, 16));
    const double kstar = a.p0, s20 = a.s20, p0 = a.p0, cl = a.cl, eps_tol = a.eps_tol;
    const int cap = a.capacity, ldmax = a.ld;
    double alpha[DOUT], b1 = 0.0, b2 = 0.0;  // lane r owns entry r
#pragma unroll
    for (int c = 0; c < DOUT; c++) alpha[c] = 0.0;
    int bidx = -1;
    int N = 0, Nw = 1;
    unsigned int run = 0;          // sparse points at the current N, folded into the counters when N changes
    unsigned long long cnt = 0;    // lane r < NCNT owns event counter r
    double creg[16];               // row r of C
    double qd = 0.0;               // Q(r, r)
#pragma unroll
    for (int j = 0; j < 16; j++) creg[j] = 0.0;
    int nl = n;                    // points of this half still to be processed here: 0 after a hand-off
    bool have_next = false;
    double kl_next = 0.0;
    double* const Crow = C + r * W_LD;
    double* const Qrow = Q + r * W_LD;
    const HalfTile<DOUT>* tlp = &sm.tile[0];

    auto spill_rows = [&]() {
#pragma unroll
        for (int p = 0; p < 8; p++) {
            *reinterpret_cast<double2*>(Crow + 2 * p) = make_double2(creg[2 * p], creg[2 * p + 1]);
        }
    };
    auto flush_run = [&]() {
        const unsigned long long n2 = (unsigned long long)N * N;
        cnt += (r == 1) ? (unsigned long long)run : (r == 5) ? (unsigned long long)run * N : (r == 6 || r == 7) ? run * n2 : 0ull;
        run = 0;
    };

    for (int tt = 0; tt < nmax; ++tt) {
        const int s = tt & 15;
        const bool live = tt < nl;
        if (s == 0) {
            __syncwarp();  // every lane is done with the buffer the next tile will land in
            if (__all_sync(FULL, !live)) break;
            if (live) {
                const int kt = tt >> 4;
                if (tt + 16 < n) {
                    if (r == 0) issue_tile<DOUT>(sm.tile[(kt + 1) & 1], &sm.mbar[(kt + 1) & 1], a, o + tt + 16);
                    issued++;
                }
                mbar_wait(&sm.mbar[kt & 1], (uint32_t)((kt >> 1) & 1));
                waited++;
                tlp = &sm.tile[kt & 1];
                have_next = false;
            }
        }
        const HalfTile<DOUT>& tl = *tlp;
        GPC_DASSERT(s + 1 + sh1 < 18 && s + sh3 < 20);   // staged slots stay inside the tile arrays
        if (tt == 0) {  // sparse_gp.hpp:100-110 (both halves: a half has n == 0 or starts here)
            if (live) {
                if (r == 0) {
                    const double d = __dadd_rn(kstar, s20);
#pragma unroll
                    for (int c = 0; c < DOUT; c++) alpha[c] = __ddiv_rn(tl.y[c][sh1], d);
                    creg[0] = __ddiv_rn(-1.0, d);
                    qd = __ddiv_rn(1.0, kstar);
                    Qrow[0] = qd;
                    b1 = tl.x1[sh1]; b2 = tl.x2[sh1]; bidx = tl.orig[sh3];
                    cnt++;
                }
                N = 1;
            }
            continue;
        }
        const bool act = r < N;
        // k = K(x, BV) (:119); lanes >= N hold zeros.  Usually computed during the previous point's chain.
        double kl = kl_next;
        if (live && !have_next) kl = rbf(tl.x1[s + sh1], tl.x2[s + sh1], b1, b2, p0, cl);
        kl = act ? kl : 0.0;
        sm.kv[r] = kl;
        __syncwarp();
        const double nx1 = tl.x1[s + 1 + sh1], nx2 = tl.x2[s + 1 + sh1];  // the next point of the tile (a stale slot at s == 15: value unused)
        double rv = 0.0, el = 0.0, q[DOUT], rr = 0.0, gamma = 0.0;
        bool sp = false;
        // one point of the recursion up to and including a sparse update, for NG occupied groups of four columns
        auto step = [&](auto ngc) {
            constexpr int NG = decltype(ngc)::value;
            {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const double2 k01 = *reinterpret_cast<const double2*>(sm.kv + 4 * g), k23 = *reinterpret_cast<const double2*>(sm.kv + 4 * g + 2);
                    a0 = fma(creg[4 * g], k01.x, a0); a1 = fma(creg[4 * g + 1], k01.y, a1);
                    a2 = fma(creg[4 * g + 2], k23.x, a2); a3 = fma(creg[4 * g + 3], k23.y, a3);
                    const double2 q01 = *reinterpret_cast<const double2*>(Qrow + 4 * g), q23 = *reinterpret_cast<const double2*>(Qrow + 4 * g + 2);
                    e0 = fma(q01.x, k01.x, e0); e1 = fma(q01.y, k01.y, e1);
                    e2 = fma(q23.x, k23.x, e2); e3 = fma(q23.y, k23.y, e3);
                }
                rv = __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));   // (C k)_r  (:122); rows >= N are zero
                el = __dadd_rn(__dadd_rn(e0, e1), __dadd_rn(e2, e3));   // (Q k)_r = e_hat_r  (:140)
            }
            // k of the next point: independent of everything below, overlaps the butterflies and the divisions
            kl_next = rbf(nx1, nx2, b1, b2, p0, cl);
            // m = alpha'k, k'Ck, k'e_hat: one product per lane, 4-step butterflies over the half-warp
            double pm[DOUT];
#pragma unroll
            for (int c = 0; c < DOUT; c++) pm[c] = fma(alpha[c], kl, 0.0);
            double pc = fma(kl, rv, 0.0);
            double pe = fma(kl, el, 0.0);
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) pm[c] = __dadd_rn(pm[c], __shfl_xor_sync(FULL, pm[c], off));
                pc = __dadd_rn(pc, __shfl_xor_sync(FULL, pc, off));
                pe = __dadd_rn(pe, __shfl_xor_sync(FULL, pe, off));
            }
            const double s2 = __dadd_rn(kstar, pc);
            const double den = __dadd_rn(s20, s2);
            // one division for two quotients: lanes 8-15 compute r = -1/den (gaussian_noise.cpp:15-18), lanes 0-7
            // q = (y - m)/den (gaussian_noise.cpp:9-12 / gaussian_noise_3d.cpp:10-13)
            {
                const double t = __ddiv_rn((r & 8) ? -1.0 : __dadd_rn(tl.y[0][s + sh1], -pm[0]), den);
                q[0] = __shfl_sync(FULL, t, lane & 16);
                rr = __shfl_sync(FULL, t, (lane & 16) | 8);
#pragma unroll
                for (int c = 1; c < DOUT; c++) q[c] = __ddiv_rn(__dadd_rn(tl.y[c][s + sh1], -pm[c]), den);
            }
            gamma = __dadd_rn(kstar, -pe);  // :144
            if (gamma < tiny12()) gamma = 0.0;
            sp = live && gamma < eps_tol;
            double shv = 0.0, re = 0.0;
            if (sp) {
                // sparse update (:155-163)
                run++;
                const double eta = __drcp_rn(__dadd_rn(1.0, __dmul_rn(gamma, rr)));  // == 1 / (1 + gamma r), correctly rounded
                shv = act ? __dadd_rn(rv, el) : 0.0;
                sm.sv[r] = shv;  // zero beyond N
                if (act) {
#pragma unroll
                    for (int c = 0; c < DOUT; c++) alpha[c] = __dadd_rn(alpha[c], __dmul_rn(shv, __dmul_rn(q[c], eta)));
                }
                re = __dmul_rn(rr, eta);
            }
            __syncwarp();
            if (sp && act) {
                // finite factors: a column beyond N gets fma(re, sh * 0, 0) = +0, so the pad columns stay exact zeros without
                // per-column predicates; otherwise (state already inf / NaN) the columns are guarded one by one
                const bool fin = ((__double2hiint(re) & 0x7ff00000) != 0x7ff00000) && ((__double2hiint(shv) & 0x7ff00000) != 0x7ff00000);
                if (fin) {
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        const double2 s01 = *reinterpret_cast<const double2*>(sm.sv + 4 * g), s23 = *reinterpret_cast<const double2*>(sm.sv + 4 * g + 2);
                        creg[4 * g] = fma(re, __dmul_rn(shv, s01.x), creg[4 * g]);
                        creg[4 * g + 1] = fma(re, __dmul_rn(shv, s01.y), creg[4 * g + 1]);
                        creg[4 * g + 2] = fma(re, __dmul_rn(shv, s23.x), creg[4 * g + 2]);
                        creg[4 * g + 3] = fma(re, __dmul_rn(shv, s23.y), creg[4 * g + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4 * NG; j++)
                        if (j < N) creg[j] = fma(re, __dmul_rn(shv, sm.sv[j]), creg[j]);
                }
            }
        };
        if (Nw <= 8) step(IntC<2>());
        else if (Nw <= 12) step(IntC<3>());
        else step(IntC<4>());
        have_next = (s != 15);
        const bool fu = live && !sp;
        if (!__any_sync(FULL, fu)) continue;  // Q and N unchanged: neither deletion loop can fire
        // ---- at least one half has a full update (:164-203) ----
        bool fur = fu;
        if (fu) {
            flush_run();
            if (N + 1 > ldmax) {  // does not fit this bucket: hand the state to the next one
                int pos = 0;
                if (r == 0) pos = atomicAdd(a.queue_count, 1);
                pos = __shfl_sync(gm, pos, 0, 16);
                GPC_DASSERT(a.queue != nullptr && a.handoff_out != nullptr && pos >= 0 && pos < a.n_work);  // one slot per patch of this launch at most
                double* slot = a.handoff_out + (size_t)pos * slot_doubles(W_N, DOUT);
                if (r == 0) {
                    a.queue[pos] = (int32_t)patch;
                    reinterpret_cast<int*>(slot)[0] = N;
                    reinterpret_cast<int*>(slot)[1] = tt;
                }
                if (r < NCNT) reinterpret_cast<unsigned long long*>(slot + 2)[r] = cnt;
                double* v = slot + 2 + NCNT;
#pragma unroll
                for (int c = 0; c < DOUT; c++) v[c * W_N + r] = alpha[c];
                v[DOUT * W_N + r] = b1; v[(DOUT + 1) * W_N + r] = b2;
                reinterpret_cast<int*>(v + (DOUT + 2) * W_N + 2 * W_N * W_N)[r] = bidx;
#pragma unroll
                for (int j = 0; j < W_N; j++) {   // column-major slots: lane r owns row r
                    v[(DOUT + 2) * W_N + j * W_N + r] = creg[j];
                    v[(DOUT + 2) * W_N + W_N * W_N + j * W_N + r] = Qrow[j];
                }
                nl = 0;
                fur = false;
            }
        }
        double si = 0.0, ei = 0.0;
        if (fur) {
            cnt += (r == 2) ? 1ull : (r == 5) ? (unsigned long long)N : (r == 6) ? (unsigned long long)N * N
                            : (r == 8) ? (unsigned long long)(N + 1) * (N + 1) : 0ull;
            si = act ? rv : (r == N ? 1.0 : 0.0);    // s = [C k; 1], e_hat' = [Q k; -1]; zeros beyond N
            ei = act ? el : (r == N ? -1.0 : 0.0);
            sm.sv[r] = si;
            sm.ev[r] = ei;
            if (act) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) alpha[c] = __dadd_rn(alpha[c], __dmul_rn(q[c], rv));
            }
            if (r == N) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) alpha[c] = __dadd_rn(0.0, __dmul_rn(q[c], 1.0));
                b1 = tl.x1[s + sh1]; b2 = tl.x2[s + sh1]; bidx = tl.orig[s + sh3];
            }
        }
        __syncwarp();
        bool need_del = false;
        if (fur) {
            const double ig = __ddiv_rn(1.0, gamma);
            const int N1 = N + 1;
            if (r < N1) {
                auto finite = [](double v) { return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000; };
                // finite factors: the columns beyond N1 get +0 (see the sparse update); otherwise guarded one by one
                if (finite(rr) && finite(ig) && finite(si) && finite(ei)) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (4 * g < N1) {
                            const double2 s01 = *reinterpret_cast<const double2*>(sm.sv + 4 * g), s23 = *reinterpret_cast<const double2*>(sm.sv + 4 * g + 2);
                            const double2 e01 = *reinterpret_cast<const double2*>(sm.ev + 4 * g), e23 = *reinterpret_cast<const double2*>(sm.ev + 4 * g + 2);
                            creg[4 * g] = fma(rr, __dmul_rn(si, s01.x), creg[4 * g]);
                            creg[4 * g + 1] = fma(rr, __dmul_rn(si, s01.y), creg[4 * g + 1]);
                            creg[4 * g + 2] = fma(rr, __dmul_rn(si, s23.x), creg[4 * g + 2]);
                            creg[4 * g + 3] = fma(rr, __dmul_rn(si, s23.y), creg[4 * g + 3]);
                            double2 q01 = *reinterpret_cast<double2*>(Qrow + 4 * g), q23 = *reinterpret_cast<double2*>(Qrow + 4 * g + 2);
                            q01.x = fma(ig, __dmul_rn(ei, e01.x), q01.x); q01.y = fma(ig, __dmul_rn(ei, e01.y), q01.y);
                            q23.x = fma(ig, __dmul_rn(ei, e23.x), q23.x); q23.y = fma(ig, __dmul_rn(ei, e23.y), q23.y);
                            *reinterpret_cast<double2*>(Qrow + 4 * g) = q01; *reinterpret_cast<double2*>(Qrow + 4 * g + 2) = q23;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if (j < N1) {
                            creg[j] = fma(rr, __dmul_rn(si, sm.sv[j]), creg[j]);
                            Qrow[j] = fma(ig, __dmul_rn(ei, sm.ev[j]), Qrow[j]);
                        }
                    }
                }
                qd = fma(ig, __dmul_rn(ei, ei), qd);  // the same operations as column r of the row above
            }
            N = N1;
            have_next = false;  // the BV set changed
            // capacity deletions (:206-223) then geometric deletions (:226-242): decided here, done in shared memory
            need_del = N > cap;
            if (!need_del && N > 1) {
                // exact shortcut: the geometric scan deletes iff score_0 is not NaN and some score 1 / Q_ii is < 1e-9f
                const double sc = (r < N) ? __ddiv_rn(1.0, qd) : 0.0;
                const bool hit = __any_sync(gm, r < N && sc < geo9());
                const bool nan0 = __any_sync(gm, r == 0 && sc != sc);
                need_del = hit && !nan0;
            }
        }
        // deletions work on full matrices in shared memory; the two halves share one home for C and take turns
        if (!__any_sync(FULL, need_del)) { Nw = max(N, __shfl_xor_sync(FULL, N, 16)); continue; }
        for (int hh = 0; hh < 2; hh++) {
            if (half == hh && need_del) {
                spill_rows();
                __syncwarp(gm);
                double minscore = 0.0;
                for (int phase = 0; phase < 2; phase++) {
                    for (;;) {
                        if (phase == 0 ? !(N > cap) : !(minscore < geo9() && N > 1)) break;
                        double sc = 0.0;
                        if (r < N) {
                            const double qii = Q[r * W_LD + r];
                            double num = __dmul_rn(alpha[0], alpha[0]);
                            if (DOUT == 3) num = __dadd_rn(num, __dadd_rn(__dmul_rn(alpha[DOUT > 1 ? 1 : 0], alpha[DOUT > 1 ? 1 : 0]), __dmul_rn(alpha[DOUT > 2 ? 2 : 0], alpha[DOUT > 2 ? 2 : 0])));
                            sc = (phase == 0) ? __ddiv_rn(num, __dadd_rn(qii, C[r * W_LD + r])) : __ddiv_rn(1.0, qii);
                        }
                        if (phase == 1) {
                            const bool hit = __any_sync(gm, r < N && sc < geo9());
                            const bool nan0 = __any_sync(gm, r == 0 && sc != sc);
                            if (!hit || nan0) { minscore = geo9(); break; }
                        }
                        double best;
                        const int loc = half_first_min(gm, sc, r < N ? r : 0x7fffffff, &best);
                        if (phase == 1) minscore = best;
                        // ---- delete_bv(loc), :252-295 ----
                        const int L = N - 1, M = N - 1;
                        cnt += (r == 9) ? (unsigned long long)M * M : (r == (phase == 0 ? 3 : 4)) ? 1ull : 0ull;
                        double csi = 0, qsi = 0, repc = 0, repq = 0;
                        const int src = (r == loc) ? L : r;
                        if (r < N) {
                            csi = C[loc * W_LD + src]; qsi = Q[loc * W_LD + src];
                            repc = C[L * W_LD + src]; repq = Q[L * W_LD + src];
                        }
                        const double cstar = C[loc * W_LD + loc], qstar = Q[loc * W_LD + loc];
                        double astar[DOUT], aL[DOUT];
#pragma unroll
                        for (int c = 0; c < DOUT; c++) { astar[c] = shfl16(gm, alpha[c], loc); aL[c] = shfl16(gm, alpha[c], L); }
                        const double b1L = shfl16(gm, b1, L), b2L = shfl16(gm, b2, L);
                        const int idL = __shfl_sync(gm, bidx, L, 16);
                        __syncwarp(gm);
                        const double qcs = __dadd_rn(qstar, cstar);
                        const double coef = (DOUT == 1) ? __ddiv_rn(astar[0], qcs) : 0.0;
                        const double iq = __ddiv_rn(1.0, qstar), iqc = __ddiv_rn(1.0, qcs);
                        if (r < N) {
                            if (r < M) {
                                if (loc != L) {
                                    C[loc * W_LD + r] = repc; C[r * W_LD + loc] = repc;
                                    Q[loc * W_LD + r] = repq; Q[r * W_LD + loc] = repq;
                                    if (r == loc) { b1 = b1L; b2 = b2L; bidx = idL; }
                                }
                                const double qci = __dadd_rn(qsi, csi);
#pragma unroll
                                for (int c = 0; c < DOUT; c++) {
                                    const double ai = (r == loc) ? aL[c] : alpha[c];
                                    alpha[c] = (DOUT == 1) ? __dadd_rn(ai, -__dmul_rn(coef, qci))
                                                           : __dadd_rn(ai, -__dmul_rn(astar[c], __dmul_rn(qcs, qci)));
                                }
                                sm.sv[r] = qsi;   // Qstar
                                sm.ev[r] = qci;   // Qstar + Cstar
                            }
                            C[L * W_LD + r] = 0.0; C[r * W_LD + L] = 0.0;
                            Q[L * W_LD + r] = 0.0; Q[r * W_LD + L] = 0.0;
                            if (r == L) {
#pragma unroll
                                for (int c = 0; c < DOUT; c++) alpha[c] = 0.0;
                                b1 = 0.0; b2 = 0.0; bidx = -1;
                            }
                        }
                        __syncwarp(gm);
                        if (r < M) {
                            const double qi = sm.sv[r], ci = sm.ev[r];
                            for (int j = 0; j < M; j += 2) {  // column pairs; the pad column (j + 1 == M) stays zero
                                double2 c = *reinterpret_cast<double2*>(Crow + j), qq = *reinterpret_cast<double2*>(Qrow + j);
                                const double2 s2v = *reinterpret_cast<const double2*>(sm.sv + j), e2v = *reinterpret_cast<const double2*>(sm.ev + j);
                                const double u0 = __dmul_rn(qi, s2v.x), v0 = __dmul_rn(ci, e2v.x);
                                c.x = __dadd_rn(c.x, fma(u0, iq, -__dmul_rn(v0, iqc)));
                                qq.x = fma(-u0, iq, qq.x);
                                if (j + 1 < M) {
                                    const double u1 = __dmul_rn(qi, s2v.y), v1 = __dmul_rn(ci, e2v.y);
                                    c.y = __dadd_rn(c.y, fma(u1, iq, -__dmul_rn(v1, iqc)));
                                    qq.y = fma(-u1, iq, qq.y);
                                }
                                *reinterpret_cast<double2*>(Crow + j) = c;
                                *reinterpret_cast<double2*>(Qrow + j) = qq;
                            }
                        }
                        N = M;
                        __syncwarp(gm);
                    }
                }
                // rows back into registers
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const double2 c2 = *reinterpret_cast<const double2*>(Crow + 2 * p);
                    creg[2 * p] = c2.x; creg[2 * p + 1] = c2.y;
                }
                qd = Q[r * W_LD + r];
            }
            __syncwarp();
        }
        Nw = max(N, __shfl_xor_sync(FULL, N, 16));
    }
    __syncwarp();
    if (issued > waited) mbar_wait(&sm.mbar[waited & 1], (uint32_t)((waited >> 1) & 1));  // never leave with a copy in flight
    if (w >= a.n_work || n == 0 || nl == 0) return;
    flush_run();
    if (r == 0) {
        a.nbv[op] = N;
        const double c00 = creg[0];
        a.flags[op] = (c00 != c00) ? 1 : 0;
        atomicAdd(a.stats + 0, (unsigned long long)n);
    }
    if (r < NCNT && cnt) atomicAdd(a.stats + 1 + r, cnt);
    const int64_t ob = op * cap;
    if (r < N) {
#pragma unroll
        for (int c = 0; c < DOUT; c++) a.o_alpha[c][ob + r] = alpha[c];
        a.o_b1[ob + r] = b1;
        a.o_b2[ob + r] = b2;
        a.o_idx[ob + r] = bidx;
    }
    if (a.dumpC) {
        const int64_t od = op * (int64_t)cap * cap;
        if (r < N) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j < N) {
                    a.dumpC[od + r * N + j] = creg[j];
                    a.dumpQ[od + r * N + j] = Qrow[j];
                }
            }
        }
    }
}

// =====================================================================================
// Bucket 1 (N + 1 <= 32): TWO WARPS per patch.  Warp 0 owns C, warp 1 owns Q (rows contiguous,
// P_LD = 34).  Both warps keep alpha / BV / k for index `lane` in registers and compute every scalar
// redundantly, so a sparse point needs ONE block barrier (the exchange of C k and Q k); the rank-1
// update of C is done by warp 0 alone while warp 1 already runs ahead to the next point's Q k.
// A full update costs two barriers, each deletion two more.  Same arithmetic as every other bucket.
// =====================================================================================
constexpr int P_N = 32;
constexpr int P_LD = 34;

struct PairSmem {
    double C[P_N * P_LD], Q[P_N * P_LD];
    double kv[2][P_N];          // per-warp copy of k
    double xv[2][2][P_N];       // [parity of the point][matrix][row]: C k and Q k exchange
    double sv[2][P_N], ev[2][P_N];
    StagedPoint pts[2][32];
    unsigned long long cnt[NCNT];
    int spos;
};

template <int NG>
__device__ __forceinline__ double row4_padded(const double* row, const double* k, int n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (4 * g < n) {
            const double2 m01 = *reinterpret_cast<const double2*>(row + 4 * g), m23 = *reinterpret_cast<const double2*>(row + 4 * g + 2);
            const double2 k01 = *reinterpret_cast<const double2*>(k + 4 * g), k23 = *reinterpret_cast<const double2*>(k + 4 * g + 2);
            a0 = fma(m01.x, k01.x, a0);
            a1 = fma(m01.y, k01.y, a1);
            a2 = fma(m23.x, k23.x, a2);
            a3 = fma(m23.y, k23.y, a3);
        }
    }
    return __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
}

// one patch of the bucket-1 work list (work = position in the list / the hand-off slot array)
__device__ __forceinline__ void pair_patch(const SogpArgs& a, PairSmem& sm, const int work) {

    const int t = threadIdx.x, lane = t & 31, mat = t >> 5;
    double* const Mown = mat ? sm.Q : sm.C;
    double* const Mrow = Mown + lane * P_LD;
    const int64_t patch = a.patch_ids ? (int64_t)a.patch_ids[work] : a.first_patch + work;
    const int64_t o = a.off[patch];
    const int n = (int)(a.off[patch + 1] - o);
    const int64_t op = patch - a.out_first;
    if (n == 0) {
        if (t == 0) { a.nbv[op] = 0; a.flags[op] = 0; }
        return;
    }
    for (int i = t; i < P_N * P_LD; i += 64) { sm.C[i] = 0.0; sm.Q[i] = 0.0; }
    sm.kv[mat][lane] = 0.0; sm.sv[mat][lane] = 0.0; sm.ev[mat][lane] = 0.0;
    if (t < NCNT) sm.cnt[t] = 0;
    __syncthreads();
    const double kstar = a.p0, s20 = a.s20, p0 = a.p0, cl = a.cl, eps_tol = a.eps_tol;
    const int cap = a.capacity, ldmax = a.ld;
    double alpha = 0.0, b1 = 0.0, b2 = 0.0;
    int bidx = -1;
    int N = 0, tt0 = 0;
    unsigned int run = 0;
    if (a.handoff_in) {  // resume a patch that outgrew bucket 0 (slots hold W_N x W_N column-major matrices)
        const double* slot = a.handoff_in + (size_t)work * slot_doubles(W_N);
        N = reinterpret_cast<const int*>(slot)[0];
        tt0 = reinterpret_cast<const int*>(slot)[1];
        if (t < NCNT) sm.cnt[t] = reinterpret_cast<const unsigned long long*>(slot + 2)[t];
        const double* v = slot + 2 + NCNT;
        if (lane < W_N) {
            alpha = v[lane]; b1 = v[W_N + lane]; b2 = v[2 * W_N + lane];
            bidx = reinterpret_cast<const int*>(v + 3 * W_N + 2 * W_N * W_N)[lane];
        }
        for (int e = t; e < W_N * W_N; e += 64) {
            const int j = e / W_N, i = e - j * W_N;
            sm.C[i * P_LD + j] = v[3 * W_N + e];
            sm.Q[i * P_LD + j] = v[3 * W_N + W_N * W_N + e];
        }
        __syncthreads();
    }
    const double* const gx1 = a.fx1 + o;
    const double* const gx2 = a.fx2 + o;
    const double* const gy = a.fy[0] + o;
    const int32_t* const go = a.forig + o;
    const bool w0l0 = (t == 0);

    for (int tt = tt0; tt < n; ++tt) {
        if ((tt & 31) == 0 || tt == tt0) {  // each warp stages its own copy of the next 32 points
            __syncwarp();
            const int i = (tt & ~31) + lane;
            if (i < n) {
                StagedPoint sp;
                sp.x1 = gx1[i]; sp.x2 = gx2[i]; sp.y[0] = gy[i]; sp.orig = go[i]; sp.pad = 0;
                sm.pts[mat][lane] = sp;
            }
            __syncwarp();
        }
        const StagedPoint pt = sm.pts[mat][tt & 31];
        const double x1 = pt.x1, x2 = pt.x2, y = pt.y[0];
        const int orig = pt.orig;
        if (N == 0) {  // sparse_gp.hpp:100-110
            const double d = __dadd_rn(kstar, s20);
            if (lane == 0) { alpha = __ddiv_rn(y, d); b1 = x1; b2 = x2; bidx = orig; }
            if (t == 0) { sm.C[0] = __ddiv_rn(-1.0, d); sm.cnt[0]++; }
            if (t == 32) sm.Q[0] = __ddiv_rn(1.0, kstar);
            N = 1;
            __syncthreads();
            continue;
        }
        const bool act = lane < N;
        double kl = rbf(x1, x2, b1, b2, p0, cl);
        kl = act ? kl : 0.0;
        sm.kv[mat][lane] = kl;
        __syncwarp();
        double rv = 0.0;
        if (act) rv = row4_padded<8>(Mrow, sm.kv[mat], N);
        const int par = tt & 1;
        sm.xv[par][mat][lane] = rv;
        __syncthreads();  // the one barrier of a sparse point
        const double other = sm.xv[par][1 - mat][lane];
        const double ck = mat ? other : rv, el = mat ? rv : other;
        double pm = act ? fma(alpha, kl, 0.0) : 0.0;
        double pc = act ? fma(kl, ck, 0.0) : 0.0;
        double pe = act ? fma(kl, el, 0.0) : 0.0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            pm = __dadd_rn(pm, shfl_xor_d(pm, off));
            pc = __dadd_rn(pc, shfl_xor_d(pc, off));
            pe = __dadd_rn(pe, shfl_xor_d(pe, off));
        }
        const double s2 = __dadd_rn(kstar, pc);
        const double den = __dadd_rn(s20, s2);
        const double rr = __ddiv_rn(-1.0, den);
        const double q = __ddiv_rn(__dadd_rn(y, -pm), den);
        double gamma = __dadd_rn(kstar, -pe);
        if (gamma < tiny12()) gamma = 0.0;
        if (gamma < eps_tol) {
            // sparse update (:155-163): Q untouched, warp 1 goes straight on to the next point
            run++;
            const double eta = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(gamma, rr)));
            const double sh = act ? __dadd_rn(ck, el) : 0.0;
            if (act) alpha = __dadd_rn(alpha, __dmul_rn(sh, __dmul_rn(q, eta)));
            if (mat == 0) {
                sm.sv[0][lane] = sh;
                __syncwarp();
                const double re = __dmul_rn(rr, eta);
                if (act) {
#pragma unroll
                    for (int p2 = 0; p2 < P_N / 2; p2++) {
                        const int j = 2 * p2;
                        if (j < N) {
                            double2 c = *reinterpret_cast<double2*>(Mrow + j);
                            const double2 s2v = *reinterpret_cast<const double2*>(sm.sv[0] + j);
                            c.x = fma(re, __dmul_rn(sh, s2v.x), c.x);
                            if (j + 1 < N) c.y = fma(re, __dmul_rn(sh, s2v.y), c.y);
                            *reinterpret_cast<double2*>(Mrow + j) = c;
                        }
                    }
                }
            }
            continue;
        }
        // full update (:164-203)
        if (w0l0) {
            const unsigned long long n2 = (unsigned long long)N * N;
            sm.cnt[1] += run; sm.cnt[5] += (unsigned long long)run * N; sm.cnt[6] += run * n2; sm.cnt[7] += run * n2;
        }
        run = 0;
        if (N + 1 > ldmax) {  // hand the state to bucket 2 (slots hold P_N x P_N column-major matrices)
            if (t == 0) sm.spos = atomicAdd(a.queue_count, 1);
            __syncthreads();
            const int pos = sm.spos;
            GPC_DASSERT(a.queue != nullptr && a.handoff_out != nullptr && pos >= 0 && pos < a.n_work);  // one slot per patch of this launch at most
                double* slot = a.handoff_out + (size_t)pos * slot_doubles(P_N);
            if (t == 0) {
                a.queue[pos] = (int32_t)patch;
                reinterpret_cast<int*>(slot)[0] = N;
                reinterpret_cast<int*>(slot)[1] = tt;
                for (int i = 0; i < NCNT; i++) reinterpret_cast<unsigned long long*>(slot + 2)[i] = sm.cnt[i];
            }
            double* v = slot + 2 + NCNT;
            if (mat == 0) {
                v[lane] = alpha; v[P_N + lane] = b1; v[2 * P_N + lane] = b2;
                reinterpret_cast<int*>(v + 3 * P_N + 2 * P_N * P_N)[lane] = bidx;
            }
            for (int e = t; e < P_N * P_N; e += 64) {
                const int j = e / P_N, i = e - j * P_N;
                v[3 * P_N + e] = sm.C[i * P_LD + j];
                v[3 * P_N + P_N * P_N + e] = sm.Q[i * P_LD + j];
            }
            return;
        }
        if (w0l0) {
            sm.cnt[2]++; sm.cnt[5] += N; sm.cnt[6] += (unsigned long long)N * N; sm.cnt[8] += (unsigned long long)(N + 1) * (N + 1);
        }
        {
            double svl = act ? ck : 0.0, evl = act ? el : 0.0;
            if (act) alpha = __dadd_rn(alpha, __dmul_rn(q, ck));
            if (lane == N) {
                svl = 1.0; evl = -1.0;
                alpha = __dadd_rn(0.0, __dmul_rn(q, 1.0));
                b1 = x1; b2 = x2; bidx = orig;
            }
            sm.sv[mat][lane] = svl;
            sm.ev[mat][lane] = evl;
            __syncwarp();
            const int N1 = N + 1;
            const double coef = mat ? __ddiv_rn(1.0, gamma) : rr;       // Q += (1/gamma) e' e'^T ; C += r s s^T
            const double* const vec = mat ? sm.ev[1] : sm.sv[0];
            const double vi = mat ? evl : svl;
            if (lane < N1) {
#pragma unroll 4
                for (int j = 0; j < N1; j += 2) {
                    double2 c = *reinterpret_cast<double2*>(Mrow + j);
                    const double2 w2 = *reinterpret_cast<const double2*>(vec + j);
                    c.x = fma(coef, __dmul_rn(vi, w2.x), c.x);
                    if (j + 1 < N1) c.y = fma(coef, __dmul_rn(vi, w2.y), c.y);
                    *reinterpret_cast<double2*>(Mrow + j) = c;
                }
            }
            N = N1;
        }
        // capacity deletions (:206-223) then geometric deletions (:226-242)
        double minscore = 0.0;
        for (int phase = 0; phase < 2; phase++) {
            for (;;) {
                if (phase == 0 ? !(N > cap) : !(minscore < geo9() && N > 1)) break;
                __syncthreads();  // both matrices up to date before their diagonals / columns are read
                double sc = 0.0;
                if (lane < N) {
                    const double qii = sm.Q[lane * P_LD + lane];
                    sc = (phase == 0) ? __ddiv_rn(__dmul_rn(alpha, alpha), __dadd_rn(qii, sm.C[lane * P_LD + lane])) : __ddiv_rn(1.0, qii);
                }
                if (phase == 1) {
                    // exact shortcut: the scan deletes iff score_0 is not NaN and some score is < 1e-9f
                    const bool hit = __any_sync(0xffffffffu, lane < N && sc < geo9());
                    const bool nan0 = __any_sync(0xffffffffu, lane == 0 && sc != sc);
                    if (!hit || nan0) { minscore = geo9(); break; }
                }
                double best;
                const int loc = warp_first_min(sc, lane < N ? lane : 0x7fffffff, &best);
                if (phase == 1) minscore = best;
                // ---- delete_bv(loc), :252-295 ----
                const int L = N - 1, M = N - 1;
                if (w0l0) { sm.cnt[9] += (unsigned long long)M * M; sm.cnt[phase == 0 ? 3 : 4]++; }
                double csi = 0, qsi = 0, rep = 0;
                const int src = (lane == loc) ? L : lane;
                if (lane < N) {
                    csi = sm.C[loc * P_LD + src]; qsi = sm.Q[loc * P_LD + src];
                    rep = Mown[L * P_LD + src];
                }
                const double cstar = sm.C[loc * P_LD + loc], qstar = sm.Q[loc * P_LD + loc];
                const double astar = __shfl_sync(0xffffffffu, alpha, loc);
                const double aL = __shfl_sync(0xffffffffu, alpha, L);
                const double b1L = __shfl_sync(0xffffffffu, b1, L), b2L = __shfl_sync(0xffffffffu, b2, L);
                const int idL = __shfl_sync(0xffffffffu, bidx, L);
                __syncthreads();  // every read of the old matrices is done
                const double qcs = __dadd_rn(qstar, cstar);
                const double coef = __ddiv_rn(astar, qcs);
                const double iq = __ddiv_rn(1.0, qstar), iqc = __ddiv_rn(1.0, qcs);
                double qci = 0.0;
                if (lane < N) {
                    if (lane < M) {
                        if (loc != L) {
                            Mown[loc * P_LD + lane] = rep; Mown[lane * P_LD + loc] = rep;
                            if (lane == loc) { b1 = b1L; b2 = b2L; bidx = idL; }
                        }
                        qci = __dadd_rn(qsi, csi);
                        const double ai = (lane == loc) ? aL : alpha;
                        alpha = __dadd_rn(ai, -__dmul_rn(coef, qci));
                    } else {
                        qsi = 0.0;
                    }
                    Mown[L * P_LD + lane] = 0.0; Mown[lane * P_LD + L] = 0.0;
                    if (lane == L) { alpha = 0.0; b1 = 0.0; b2 = 0.0; bidx = -1; }
                } else {
                    qsi = 0.0;
                }
                sm.sv[mat][lane] = qsi;   // Qstar (zero beyond M)
                sm.ev[mat][lane] = qci;   // Qstar + Cstar
                __syncwarp();
                if (lane < M) {
                    const double* const qv = sm.sv[mat];
                    const double* const cv = sm.ev[mat];
#pragma unroll 4
                    for (int j = 0; j < M; j += 2) {
                        double2 c = *reinterpret_cast<double2*>(Mrow + j);
                        const double2 q2 = *reinterpret_cast<const double2*>(qv + j);
                        const double2 c2 = *reinterpret_cast<const double2*>(cv + j);
                        const double u0 = __dmul_rn(qsi, q2.x), u1 = __dmul_rn(qsi, q2.y);
                        if (mat == 0) {
                            const double w0 = fma(u0, iq, -__dmul_rn(__dmul_rn(qci, c2.x), iqc));
                            const double w1 = fma(u1, iq, -__dmul_rn(__dmul_rn(qci, c2.y), iqc));
                            c.x = __dadd_rn(c.x, w0);
                            if (j + 1 < M) c.y = __dadd_rn(c.y, w1);
                        } else {
                            c.x = fma(-u0, iq, c.x);
                            if (j + 1 < M) c.y = fma(-u1, iq, c.y);
                        }
                        *reinterpret_cast<double2*>(Mrow + j) = c;
                    }
                }
                N = M;
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        const unsigned long long n2 = (unsigned long long)N * N;
        sm.cnt[1] += run; sm.cnt[5] += (unsigned long long)run * N; sm.cnt[6] += run * n2; sm.cnt[7] += run * n2;
        a.nbv[op] = N;
        const double c00 = sm.C[0];
        a.flags[op] = (c00 != c00) ? 1 : 0;
        unsigned long long* st = a.stats;
        atomicAdd(st + 0, (unsigned long long)n);
        for (int i = 0; i < NCNT; i++)
            if (sm.cnt[i]) atomicAdd(st + 1 + i, sm.cnt[i]);
    }
    const int64_t ob = op * cap;
    if (mat == 0 && lane < N) {
        a.o_alpha[0][ob + lane] = alpha;
        a.o_b1[ob + lane] = b1;
        a.o_b2[ob + lane] = b2;
        a.o_idx[ob + lane] = bidx;
    }
    if (a.dumpC) {
        const int64_t od = op * (int64_t)cap * cap;
        for (int e = t; e < N * N; e += 64) {
            const int i = e / N, j = e - i * N;
            a.dumpC[od + e] = sm.C[i * P_LD + j];
            a.dumpQ[od + e] = sm.Q[i * P_LD + j];
        }
    }
}
__global__ void __launch_bounds__(64, 10) sogp_fit_pair_kernel(SogpArgs a) {
    __shared__ __align__(16) PairSmem sm;
    pair_patch(a, sm, (int)blockIdx.x);
}

// =====================================================================================
// Buckets 1-3: NT = 2*RB threads per patch, LD constexpr
// =====================================================================================
template <int LD, int DOUT>
struct Smem {
    double* base;
    static constexpr int V = 2 * LD * LD;                      // start of the vectors
    static constexpr int kDoubles = V + (8 + DOUT) * LD + 16;
    __device__ __forceinline__ double* C() const { return base; }
    __device__ __forceinline__ double* Q() const { return base + LD * LD; }
    __device__ __forceinline__ double* alpha(int c = 0) const { return base + V + c * LD; }
    __device__ __forceinline__ double* b1() const { return base + V + DOUT * LD; }
    __device__ __forceinline__ double* b2() const { return base + V + (DOUT + 1) * LD; }
    __device__ __forceinline__ double* kv() const { return base + V + (DOUT + 2) * LD; }
    __device__ __forceinline__ double* ck() const { return base + V + (DOUT + 3) * LD; }
    __device__ __forceinline__ double* ev() const { return base + V + (DOUT + 4) * LD; }
    __device__ __forceinline__ double* sv() const { return base + V + (DOUT + 5) * LD; }
    __device__ __forceinline__ double* qsv() const { return base + V + (DOUT + 6) * LD; }
    __device__ __forceinline__ double* qcv() const { return base + V + (DOUT + 7) * LD; }
    __device__ __forceinline__ double* scv() const { return kv(); }                               // deletion scores (k is dead by then)
    __device__ __forceinline__ double* scal() const { return base + V + (8 + DOUT) * LD; }        // 16 CTA-wide scalars
    __device__ __forceinline__ int* bidx() const { return reinterpret_cast<int*>(base + kDoubles); }
};

// One score per BV, computed ONCE by thread i (a division each), then every warp reduces the scores from
// shared memory with the "first strict minimum" rule.  MODE 0: alpha_i^2 / (Q_ii + C_ii) (sparse_gp.hpp:213);
// MODE 1: 1 / Q_ii with the exact shortcut "delete iff score_0 is not NaN and some score < 1e-9f" (:229-236).
// Contains block barriers: must be reached by all threads.  Returns loc, or -1 when MODE 1 finds nothing to delete.
template <int LD, int DOUT, int MODE, class S>
__device__ __forceinline__ int block_argmin(const S& s, int N, int t, int lane) {
    double sc = 0.0;
    if (t < N) {
        const double qii = s.Q()[t * LD + t];
        if (MODE == 0) {
            const double a0 = s.alpha(0)[t];
            double num = __dmul_rn(a0, a0);
            if (DOUT == 3) {  // alpha.row(i).squaredNorm(), sparse_gp_field.hpp:187
                const double a1 = s.alpha(DOUT > 1 ? 1 : 0)[t], a2 = s.alpha(DOUT > 2 ? 2 : 0)[t];
                num = __dadd_rn(num, __dadd_rn(__dmul_rn(a1, a1), __dmul_rn(a2, a2)));
            }
            sc = __ddiv_rn(num, __dadd_rn(qii, s.C()[t * LD + t]));
        } else {
            sc = __ddiv_rn(1.0, qii);
        }
        s.scv()[t] = sc;
    }
    if (MODE == 1) {
        if (!__syncthreads_or(t < N && sc < geo9())) return -1;
        if (s.scv()[0] != s.scv()[0]) return -1;  // score_0 is NaN: nothing is ever "< minscore"
    } else {
        __syncthreads();
    }
    double best = 0.0;
    int bi = 0x7fffffff;
    bool nan0 = false;
    for (int i = lane; i < N; i += 32) {
        const double v = s.scv()[i];
        if (v != v) {
            if (i == 0) nan0 = true;
            continue;
        }
        if (bi == 0x7fffffff || v < best) { best = v; bi = i; }
    }
    if (nan0) { best = __longlong_as_double(0x7ff8000000000000LL); bi = 0; }
    double ms;
    return warp_first_min(best, bi, &ms);
}

// sparse_gp::delete_bv, sparse_gp.hpp:252-295.  Uniform across the CTA; ends with a barrier.
// The three divisions are done by warp 0 only and published through scal[4..6].
template <int LD, int RB, int NT, int DOUT>
__device__ __forceinline__ void delete_bv(const Smem<LD, DOUT>& s, int& N, int loc, int t) {
    const int L = N - 1, M = N - 1;
    const int lane_ = t & 31, w_ = t >> 5;
    const int irow = 16 * (w_ % (RB / 16)) + (lane_ & 15);
    const int jq = 2 * (w_ / (RB / 16)) + (lane_ >> 4);
    double* const C = s.C();
    double* const Q = s.Q();
    double csi = 0, qsi = 0, repc = 0, repq = 0, nb1 = 0, nb2 = 0;
    double ai[DOUT];
#pragma unroll
    for (int c = 0; c < DOUT; c++) ai[c] = 0.0;
    int nidx = -1;
    if (t < N) {
        const int src = (t == loc) ? L : t;  // Cs(loc) = Cs(L), Crep(loc) = Crep(L), alpha(loc) = alpha(L)
        csi = C[loc * LD + src];
        qsi = Q[loc * LD + src];
        repc = C[L * LD + src];
        repq = Q[L * LD + src];
#pragma unroll
        for (int c = 0; c < DOUT; c++) ai[c] = s.alpha(c)[src];
        if (t == loc) { nb1 = s.b1()[L]; nb2 = s.b2()[L]; nidx = s.bidx()[L]; }
    }
    // scal[8] = 1/q*, [9] = 1/(q*+c*), [10] = q*+c*, [11..] = alpha* / (q*+c*) (sparse_gp) or alpha* (sparse_gp_field)
    double coef0[DOUT], iq0 = 0.0, iqc0 = 0.0, qcs0 = 0.0;
#pragma unroll
    for (int c = 0; c < DOUT; c++) coef0[c] = 0.0;
    if (w_ == 0) {
        const double cstar = C[loc * LD + loc], qstar = Q[loc * LD + loc];
        qcs0 = __dadd_rn(qstar, cstar);
#pragma unroll
        for (int c = 0; c < DOUT; c++) {
            const double astar = s.alpha(c)[loc];
            coef0[c] = (DOUT == 1) ? __ddiv_rn(astar, qcs0) : astar;
        }
        iq0 = __ddiv_rn(1.0, qstar);
        iqc0 = __ddiv_rn(1.0, qcs0);
    }
    __syncthreads();  // every read of the old state is done
    if (t == 0) {
        s.scal()[8] = iq0; s.scal()[9] = iqc0; s.scal()[10] = qcs0;
#pragma unroll
        for (int c = 0; c < DOUT; c++) s.scal()[11 + c] = coef0[c];
    }
    double qci = 0.0;
    if (t < N) {
        if (t < M) {
            if (loc != L) {
                C[loc * LD + t] = repc; C[t * LD + loc] = repc;
                Q[loc * LD + t] = repq; Q[t * LD + loc] = repq;
                if (t == loc) { s.b1()[loc] = nb1; s.b2()[loc] = nb2; s.bidx()[loc] = nidx; }
            }
            qci = __dadd_rn(qsi, csi);
            s.qsv()[t] = qsi;
            s.qcv()[t] = qci;
        }
        C[L * LD + t] = 0.0; C[t * LD + L] = 0.0;
        Q[L * LD + t] = 0.0; Q[t * LD + L] = 0.0;
        if (t == L) {
#pragma unroll
            for (int c = 0; c < DOUT; c++) s.alpha(c)[L] = 0.0;
            s.b1()[L] = 0.0; s.b2()[L] = 0.0; s.bidx()[L] = -1;
        }
    }
    __syncthreads();
    const double iq = s.scal()[8], iqc = s.scal()[9];
    if (t < M) {
        const double qcs = s.scal()[10];
#pragma unroll
        for (int c = 0; c < DOUT; c++) {
            const double cf = s.scal()[11 + c];
            s.alpha(c)[t] = (DOUT == 1) ? __dadd_rn(ai[c], -__dmul_rn(cf, qci))                       // sparse_gp.hpp:285
                                        : __dadd_rn(ai[c], -__dmul_rn(cf, __dmul_rn(qcs, qci)));      // sparse_gp_field.hpp:250-253
        }
    }
    if (irow < M) {
        const double qi = s.qsv()[irow], ci = s.qcv()[irow];
        for (int j = jq; j < M; j += 4) {
            const int idx = j * LD + irow;
            const double u = __dmul_rn(qi, s.qsv()[j]);
            const double v = __dmul_rn(ci, s.qcv()[j]);
            const double w = fma(u, iq, -__dmul_rn(v, iqc));
            C[idx] = __dadd_rn(C[idx], w);
            Q[idx] = fma(-u, iq, Q[idx]);
        }
    }
    N = M;
    __syncthreads();
}

// delete_bv for any thread count (used by the rare paths of the fused kernel): the matrix update strides over elements.
// The three divisions are done by warp 0 only and published through scal[4..6].
template <int LD, int NT, int DOUT, class S>
__device__ __forceinline__ void delete_bv_any(const S& s, int& N, int loc, int t) {
    const int L = N - 1, M = N - 1;
    const int lane_ = t & 31, w_ = t >> 5;
    double* const C = s.C();
    double* const Q = s.Q();
    double csi = 0, qsi = 0, repc = 0, repq = 0, nb1 = 0, nb2 = 0;
    double ai[DOUT];
#pragma unroll
    for (int c = 0; c < DOUT; c++) ai[c] = 0.0;
    int nidx = -1;
    if (t < N) {
        const int src = (t == loc) ? L : t;  // Cs(loc) = Cs(L), Crep(loc) = Crep(L), alpha(loc) = alpha(L)
        csi = C[loc * LD + src];
        qsi = Q[loc * LD + src];
        repc = C[L * LD + src];
        repq = Q[L * LD + src];
#pragma unroll
        for (int c = 0; c < DOUT; c++) ai[c] = s.alpha(c)[src];
        if (t == loc) { nb1 = s.b1()[L]; nb2 = s.b2()[L]; nidx = s.bidx()[L]; }
    }
    // scal[8] = 1/q*, [9] = 1/(q*+c*), [10] = q*+c*, [11..] = alpha* / (q*+c*) (sparse_gp) or alpha* (sparse_gp_field)
    double coef0[DOUT], iq0 = 0.0, iqc0 = 0.0, qcs0 = 0.0;
#pragma unroll
    for (int c = 0; c < DOUT; c++) coef0[c] = 0.0;
    if (w_ == 0) {
        const double cstar = C[loc * LD + loc], qstar = Q[loc * LD + loc];
        qcs0 = __dadd_rn(qstar, cstar);
#pragma unroll
        for (int c = 0; c < DOUT; c++) {
            const double astar = s.alpha(c)[loc];
            coef0[c] = (DOUT == 1) ? __ddiv_rn(astar, qcs0) : astar;
        }
        iq0 = __ddiv_rn(1.0, qstar);
        iqc0 = __ddiv_rn(1.0, qcs0);
    }
    __syncthreads();  // every read of the old state is done
    if (t == 0) {
        s.scal()[8] = iq0; s.scal()[9] = iqc0; s.scal()[10] = qcs0;
#pragma unroll
        for (int c = 0; c < DOUT; c++) s.scal()[11 + c] = coef0[c];
    }
    double qci = 0.0;
    if (t < N) {
        if (t < M) {
            if (loc != L) {
                C[loc * LD + t] = repc; C[t * LD + loc] = repc;
                Q[loc * LD + t] = repq; Q[t * LD + loc] = repq;
                if (t == loc) { s.b1()[loc] = nb1; s.b2()[loc] = nb2; s.bidx()[loc] = nidx; }
            }
            qci = __dadd_rn(qsi, csi);
            s.qsv()[t] = qsi;
            s.qcv()[t] = qci;
        }
        C[L * LD + t] = 0.0; C[t * LD + L] = 0.0;
        Q[L * LD + t] = 0.0; Q[t * LD + L] = 0.0;
        if (t == L) {
#pragma unroll
            for (int c = 0; c < DOUT; c++) s.alpha(c)[L] = 0.0;
            s.b1()[L] = 0.0; s.b2()[L] = 0.0; s.bidx()[L] = -1;
        }
    }
    __syncthreads();
    const double iq = s.scal()[8], iqc = s.scal()[9];
    if (t < M) {
        const double qcs = s.scal()[10];
#pragma unroll
        for (int c = 0; c < DOUT; c++) {
            const double cf = s.scal()[11 + c];
            s.alpha(c)[t] = (DOUT == 1) ? __dadd_rn(ai[c], -__dmul_rn(cf, qci))                       // sparse_gp.hpp:285
                                        : __dadd_rn(ai[c], -__dmul_rn(cf, __dmul_rn(qcs, qci)));      // sparse_gp_field.hpp:250-253
        }
    }
    for (int e = t; e < M * M; e += NT) {
        const int j = e / M, irow = e - j * M;
        const int idx = j * LD + irow;
        const double u = __dmul_rn(s.qsv()[irow], s.qsv()[j]);
        const double v = __dmul_rn(s.qcv()[irow], s.qcv()[j]);
        const double w = fma(u, iq, -__dmul_rn(v, iqc));
        C[idx] = __dadd_rn(C[idx], w);
        Q[idx] = fma(-u, iq, Q[idx]);
    }
    N = M;
    __syncthreads();
}

template <int LD, int RB, int NT, int LD_IN, bool SPILL, int DOUT>
__global__ void __launch_bounds__(NT) sogp_fit_kernel(SogpArgs a) {
    extern __shared__ double smem_dyn[];
    // SPILL: the state lives in a per-CTA slice of global memory (L2-resident) instead of shared memory;
    // block barriers order the accesses exactly as they do for shared memory.
    double* const smem_d = SPILL ? a.spill + (size_t)blockIdx.x * (Smem<LD, DOUT>::kDoubles + LD) : smem_dyn;
    const Smem<LD, DOUT> s{smem_d};
    double* const C = s.C();
    double* const Q = s.Q();
    // NT = 4*RB threads.  A warp covers 16 rows: lanes 0-15 and 16-31 hold the same rows and split the columns.
    // Matvec: warps [0, RB/16) take C, the rest Q; a half-warp accumulates the canonical partials a_{2h}, a_{2h+1}.
    // Updates: row irow, columns j = jq (mod 4) with jq = 2*mat + half.
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int half = lane >> 4;
    const int irow = 16 * (w % (RB / 16)) + (lane & 15);
    const int mat = w / (RB / 16);
    const int jq = 2 * mat + half;
    const int64_t patch = a.patch_ids ? (int64_t)a.patch_ids[blockIdx.x] : a.first_patch + blockIdx.x;
    const int64_t o = a.off[patch];
    const int n = (int)(a.off[patch + 1] - o);
    const int64_t op = patch - a.out_first;
    if (n == 0) {
        if (t == 0) { a.nbv[op] = 0; a.flags[op] = 0; }
        return;
    }
    for (int i = t; i < Smem<LD, DOUT>::kDoubles; i += NT) smem_d[i] = 0.0;
    for (int i = t; i < LD; i += NT) s.bidx()[i] = -1;
    __syncthreads();

    const double kstar = a.p0, s20 = a.s20, p0 = a.p0, cl = a.cl, eps_tol = a.eps_tol;
    const int cap = a.capacity, ldmax = a.ld;
    int N = 0, tt0 = 0;
    Counters cnt;
    cnt.init();
    if (a.handoff_in) {  // resume a patch that outgrew the previous bucket
        const double* slot = a.handoff_in + (size_t)blockIdx.x * slot_doubles(LD_IN, DOUT);
        N = reinterpret_cast<const int*>(slot)[0];
        tt0 = reinterpret_cast<const int*>(slot)[1];
#pragma unroll
        for (int i = 0; i < NCNT; i++) cnt.c[i] = reinterpret_cast<const unsigned long long*>(slot + 2)[i];
        const double* v = slot + 2 + NCNT;
        for (int i = t; i < LD_IN; i += NT) {
#pragma unroll
            for (int c = 0; c < DOUT; c++) s.alpha(c)[i] = v[c * LD_IN + i];
            s.b1()[i] = v[DOUT * LD_IN + i]; s.b2()[i] = v[(DOUT + 1) * LD_IN + i];
            s.bidx()[i] = reinterpret_cast<const int*>(v + (DOUT + 2) * LD_IN + 2 * LD_IN * LD_IN)[i];
        }
        for (int e = t; e < LD_IN * LD_IN; e += NT) {
            const int j = e / LD_IN, i = e - j * LD_IN;
            C[j * LD + i] = v[(DOUT + 2) * LD_IN + e];
            Q[j * LD + i] = v[(DOUT + 2) * LD_IN + LD_IN * LD_IN + e];
        }
        __syncthreads();
    }

    double nx1 = a.fx1[o + tt0], nx2 = a.fx2[o + tt0], ny[DOUT];
#pragma unroll
    for (int c = 0; c < DOUT; c++) ny[c] = a.fy[c][o + tt0];
    int norig = a.forig[o + tt0];
    for (int tt = tt0; tt < n; ++tt) {
        const double x1 = nx1, x2 = nx2;
        double y[DOUT];
#pragma unroll
        for (int c = 0; c < DOUT; c++) y[c] = ny[c];
        const int orig = norig;
        if (tt + 1 < n) {  // prefetch the next point of the stream
            nx1 = a.fx1[o + tt + 1]; nx2 = a.fx2[o + tt + 1];
#pragma unroll
            for (int c = 0; c < DOUT; c++) ny[c] = a.fy[c][o + tt + 1];
            norig = a.forig[o + tt + 1];
        }
        if (N == 0) {  // sparse_gp.hpp:100-110
            if (t == 0) {
                const double d = __dadd_rn(kstar, s20);
#pragma unroll
                for (int c = 0; c < DOUT; c++) s.alpha(c)[0] = __ddiv_rn(y[c], d);
                C[0] = __ddiv_rn(-1.0, d);
                Q[0] = __ddiv_rn(1.0, kstar);
                s.b1()[0] = x1; s.b2()[0] = x2; s.bidx()[0] = orig;
            }
            N = 1;
            cnt.c[0]++;
            __syncthreads();
            continue;
        }
        // k = K(x, BV)  (sparse_gp.hpp:119)
        if (t < N) s.kv()[t] = rbf(x1, x2, s.b1()[t], s.b2()[t], p0, cl);
        __syncthreads();
        // C k and e_hat = Q k (:122,:140); warp 0 also computes m = alpha' k (:121)
        {
            const double* const Mx = (mat ? Q : C) + irow;
            double a0 = 0.0, a1 = 0.0;
            if (irow < N) {
                for (int j = 2 * half; j < N; j += 4) {
                    a0 = fma(Mx[j * LD], s.kv()[j], a0);
                    if (j + 1 < N) a1 = fma(Mx[(j + 1) * LD], s.kv()[j + 1], a1);
                }
            }
            const double pr = __dadd_rn(a0, a1);                       // (a0+a1) in half 0, (a2+a3) in half 1
            const double rv = __dadd_rn(pr, shfl_xor_d(pr, 16));       // canonical (a0+a1)+(a2+a3)
            if (irow < N && half == 0) (mat ? s.ev() : s.ck())[irow] = rv;
        }
        double m[DOUT];
#pragma unroll
        for (int c = 0; c < DOUT; c++) m[c] = 0.0;
        if (w == 0) {
#pragma unroll
            for (int c = 0; c < DOUT; c++) m[c] = warp_dot32(s.alpha(c), s.kv(), N, lane);
        }
        __syncthreads();
        // every warp: k'Ck and k'e_hat (cheap, no division) -> gamma and the sparse / full decision
        double kck = 0.0, ke = 0.0;
        for (int j = lane; j < N; j += 32) {
            const double kj = s.kv()[j];
            kck = fma(kj, s.ck()[j], kck);
            ke = fma(kj, s.ev()[j], ke);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            kck = __dadd_rn(kck, shfl_xor_d(kck, off));
            ke = __dadd_rn(ke, shfl_xor_d(ke, off));
        }
        double gamma = __dadd_rn(kstar, -ke);                 // sparse_gp.hpp:144
        if (gamma < tiny12()) gamma = 0.0;
        const bool sparse = gamma < eps_tol;
        // warp 0 alone does the divisions and publishes them:
        // scal[0] = r, [1] = r*eta (sparse) | 1/gamma (full), [2..2+DOUT) = q_c, [5..5+DOUT) = q_c*eta (sparse)
        if (w == 0) {
            const double s2 = __dadd_rn(kstar, kck);
            const double den = __dadd_rn(s20, s2);
            const double rr = __ddiv_rn(-1.0, den);               // gaussian_noise.cpp:15-18
            double q[DOUT], qe[DOUT];
#pragma unroll
            for (int c = 0; c < DOUT; c++) q[c] = __ddiv_rn(__dadd_rn(y[c], -m[c]), den);    // gaussian_noise.cpp:9-12
            double c3;
            if (sparse) {
                const double eta = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(gamma, rr)));
#pragma unroll
                for (int c = 0; c < DOUT; c++) qe[c] = __dmul_rn(q[c], eta);
                c3 = __dmul_rn(rr, eta);
            } else {
#pragma unroll
                for (int c = 0; c < DOUT; c++) qe[c] = 0.0;
                c3 = __ddiv_rn(1.0, gamma);
            }
            if (lane == 0) {
                s.scal()[0] = rr; s.scal()[1] = c3;
#pragma unroll
                for (int c = 0; c < DOUT; c++) { s.scal()[2 + c] = q[c]; s.scal()[5 + c] = qe[c]; }
            }
        }
        if (sparse) {
            // sparse update (sparse_gp.hpp:155-163)
            cnt.run++;
            double sh = 0.0;
            if (t < N) {
                sh = __dadd_rn(s.ck()[t], s.ev()[t]);
                s.sv()[t] = sh;
            }
            __syncthreads();
            const double re = s.scal()[1];
            if (t < N) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) s.alpha(c)[t] = __dadd_rn(s.alpha(c)[t], __dmul_rn(sh, s.scal()[5 + c]));
            }
            if (irow < N) {
                const double si = s.sv()[irow];
                for (int j = jq; j < N; j += 4) {
                    const int idx = j * LD + irow;
                    C[idx] = fma(re, __dmul_rn(si, s.sv()[j]), C[idx]);
                }
            }
            continue;  // Q and N unchanged: neither deletion loop can fire
        }
        // full update (sparse_gp.hpp:164-203)
        cnt.flush(N);
        if (N + 1 > ldmax) {  // does not fit this bucket: hand the state to the next one
            __shared__ int spos;
            if (t == 0) spos = atomicAdd(a.queue_count, 1);
            __syncthreads();
            const int pos = spos;
            GPC_DASSERT(a.queue != nullptr && a.handoff_out != nullptr && pos >= 0 && pos < a.n_work);  // one slot per patch of this launch at most
                double* slot = a.handoff_out + (size_t)pos * slot_doubles(LD, DOUT);
            if (t == 0) {
                a.queue[pos] = (int32_t)patch;
                reinterpret_cast<int*>(slot)[0] = N;
                reinterpret_cast<int*>(slot)[1] = tt;
                for (int i = 0; i < NCNT; i++) reinterpret_cast<unsigned long long*>(slot + 2)[i] = cnt.c[i];
            }
            double* v = slot + 2 + NCNT;
            for (int i = t; i < LD; i += NT) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) v[c * LD + i] = s.alpha(c)[i];
                v[DOUT * LD + i] = s.b1()[i]; v[(DOUT + 1) * LD + i] = s.b2()[i];
                reinterpret_cast<int*>(v + (DOUT + 2) * LD + 2 * LD * LD)[i] = s.bidx()[i];
            }
            for (int i = t; i < LD * LD; i += NT) { v[(DOUT + 2) * LD + i] = C[i]; v[(DOUT + 2) * LD + LD * LD + i] = Q[i]; }
            return;
        }
        cnt.full(N);
        double sct = 0.0;
        if (t < N) {
            sct = s.ck()[t];
            s.sv()[t] = sct;
        }
        if (t == N) {
            s.sv()[N] = 1.0;
            s.ev()[N] = -1.0;
            s.b1()[N] = x1; s.b2()[N] = x2; s.bidx()[N] = orig;
        }
        __syncthreads();
        {
            const double rr = s.scal()[0], ig = s.scal()[1];
            if (t <= N) {
#pragma unroll
                for (int c = 0; c < DOUT; c++) {
                    const double q = s.scal()[2 + c];
                    if (t < N) s.alpha(c)[t] = __dadd_rn(s.alpha(c)[t], __dmul_rn(q, sct));
                    else s.alpha(c)[N] = __dadd_rn(0.0, __dmul_rn(q, 1.0));
                }
            }
            const int N1 = N + 1;
            if (irow < N1) {
                const double si = s.sv()[irow], ei = s.ev()[irow];
                for (int j = jq; j < N1; j += 4) {
                    const int idx = j * LD + irow;
                    C[idx] = fma(rr, __dmul_rn(si, s.sv()[j]), C[idx]);
                    Q[idx] = fma(ig, __dmul_rn(ei, s.ev()[j]), Q[idx]);
                }
            }
            N = N1;
        }
        __syncthreads();
        // capacity deletions (sparse_gp.hpp:206-223)
        while (N > cap) {
            const int loc = block_argmin<LD, DOUT, 0>(s, N, t, lane);
            cnt.c[9] += (unsigned long long)(N - 1) * (N - 1);
            delete_bv<LD, RB, NT, DOUT>(s, N, loc, t);
            cnt.c[3]++;
        }
        // geometric deletions (sparse_gp.hpp:226-242)
        while (N > 1) {
            const int loc = block_argmin<LD, DOUT, 1>(s, N, t, lane);
            if (loc < 0) break;
            cnt.c[9] += (unsigned long long)(N - 1) * (N - 1);
            delete_bv<LD, RB, NT, DOUT>(s, N, loc, t);
            cnt.c[4]++;
        }
    }
    cnt.flush(N);
    __syncthreads();
    // results
    if (t == 0) {
        a.nbv[op] = N;
        const double c00 = C[0];
        a.flags[op] = (c00 != c00) ? 1 : 0;
        publish(a, cnt, n);
    }
    const int64_t ob = op * cap;
    for (int i = t; i < N; i += NT) {
#pragma unroll
        for (int c = 0; c < DOUT; c++) a.o_alpha[c][ob + i] = s.alpha(c)[i];
        a.o_b1[ob + i] = s.b1()[i];
        a.o_b2[ob + i] = s.b2()[i];
        a.o_idx[ob + i] = s.bidx()[i];
    }
    if (a.dumpC) {
        const int64_t od = op * (int64_t)cap * cap;
        for (int e = t; e < N * N; e += NT) {
            const int i = e / N, j = e - i * N;
            a.dumpC[od + e] = C[j * LD + i];
            a.dumpQ[od + e] = Q[j * LD + i];
        }
    }
}

// =====================================================================================
// Buckets 2-3, fused (height GP, state in shared memory).  The update of point t and the two matvecs C k, Q k of
// point t + 1 share ONE pass over the matrices: every element is read once, brought up to date, written once and
// multiplied into the next point's row sums while it is still in a register.  A full update that is followed by a
// capacity deletion never materialises the (N+1) x (N+1) matrices: the deletion index is decided first from the
// O(N) quantities (updated alpha and updated diagonals), then each element takes the rank-1 update and the three
// outer products of delete_bv in one go.  Per capacity-bound point this moves 32 N^2 bytes of shared memory instead
// of 80 N^2 and needs 6 block barriers instead of 11.  Every element still receives exactly the operations, in the
// order, of sparse_gp::add / delete_bv as every other bucket performs them, so results are bit-identical.
//
// A warp covers 8 rows x 4 column classes: lane = 8 * jq + rp owns row 8 w + rp and the columns j = jq (mod 4), i.e.
// the canonical row4 partial a_jq of its row; two shuffles give (a0 + a1) + (a2 + a3).
// Rare events (geometric deletions, the hand-off to a larger bucket) first bring the matrices up to date with a
// pass without matvec and then run the step-by-step code of the generic kernel.
// =====================================================================================
enum { OP_NONE = 0, OP_SPARSE = 1, OP_FULL = 2, OP_FULL_DEL = 3 };

// Smem<LD, 1> with the matrices elsewhere (bucket 4: C and Q in a per-CTA slice of global memory, vectors in shared memory)
template <int LD>
struct SmemSplit {
    double* base;   // vectors, laid out as in Smem<LD, 1> behind its matrices
    double* mat;    // C | Q
    static constexpr int kVecDoubles = (8 + 1) * LD + 16;
    __device__ __forceinline__ double* C() const { return mat; }
    __device__ __forceinline__ double* Q() const { return mat + LD * LD; }
    __device__ __forceinline__ double* alpha(int c = 0) const { return base + c * LD; }
    __device__ __forceinline__ double* b1() const { return base + LD; }
    __device__ __forceinline__ double* b2() const { return base + 2 * LD; }
    __device__ __forceinline__ double* kv() const { return base + 3 * LD; }
    __device__ __forceinline__ double* ck() const { return base + 4 * LD; }
    __device__ __forceinline__ double* ev() const { return base + 5 * LD; }
    __device__ __forceinline__ double* sv() const { return base + 6 * LD; }
    __device__ __forceinline__ double* qsv() const { return base + 7 * LD; }
    __device__ __forceinline__ double* qcv() const { return base + 8 * LD; }
    __device__ __forceinline__ double* scv() const { return kv(); }
    __device__ __forceinline__ double* scal() const { return base + 9 * LD; }
    __device__ __forceinline__ int* bidx() const { return reinterpret_cast<int*>(base + kVecDoubles); }
};

template <int LD>
struct FusedVecs {   // beyond Smem<LD, 1>: k of the next point, permuted update vectors
    double* base;
    __device__ __forceinline__ double* kn() const { return base; }
    __device__ __forceinline__ double* svp() const { return base + LD; }
    __device__ __forceinline__ double* evp() const { return base + 2 * LD; }
    static constexpr int kDoubles = 3 * LD;
};

// first strict minimum of scv[0..n) (sparse_gp.hpp:210-217); every warp computes it redundantly
__device__ __forceinline__ int scan_first_min(const double* scv, int n, int lane) {
    double best = 0.0;
    int bi = 0x7fffffff;
    bool nan0 = false;
    for (int i = lane; i < n; i += 32) {
        const double v = scv[i];
        if (v != v) {
            if (i == 0) nan0 = true;
            continue;
        }
        if (bi == 0x7fffffff || v < best) { best = v; bi = i; }
    }
    if (nan0) { best = __longlong_as_double(0x7ff8000000000000LL); bi = 0; }
    double ms;
    return warp_first_min(best, bi, &ms);
}

struct FusedOp {
    double c0, c1, iq, iqc;   // SPARSE: c0 = r eta.  FULL / FULL_DEL: c0 = r, c1 = 1/gamma.  FULL_DEL: 1/q*, 1/(q*+c*)
    int loc;
};

// One pass over the leading nf x nf blocks of C and Q: apply OP, optionally accumulate the row sums with kn.
// sv / ev: update vectors indexed by the FINAL position (already permuted for FULL_DEL); qs / qc: the deletion's
// Qstar and Qstar + Cstar.  Writes ck / ek (the row sums) when do_mv.
// One row per thread: lane = 8 * jq + rp owns row 8 w + rp and the columns j = jq (mod 4); with LD = 8 (mod 16) the
// four 64-byte segments of a warp access fall on disjoint banks in pairs (two wavefronts for 256 bytes).
template <int LD, int OP>
__device__ __forceinline__ void fused_pass(double* __restrict__ C, double* __restrict__ Q, int nf, const FusedOp op,
                                           const double* __restrict__ sv, const double* __restrict__ ev,
                                           const double* __restrict__ qs, const double* __restrict__ qc,
                                           const double* __restrict__ kn, bool do_mv, double* __restrict__ ck,
                                           double* __restrict__ ek, int t) {
    const int lane = t & 31, w = t >> 5;
    const int rp = lane & 7, jq = lane >> 3;
    const int r0 = 8 * w + rp;
    double ac = 0.0, aq = 0.0;
    if (r0 < nf) {
        double s0 = 0, e0 = 0, qs0 = 0, qc0 = 0;
        if (OP != OP_NONE) s0 = sv[r0];
        if (OP == OP_FULL || OP == OP_FULL_DEL) e0 = ev[r0];
        if (OP == OP_FULL_DEL) { qs0 = qs[r0]; qc0 = qc[r0]; }
        const bool keep0 = r0 != op.loc;
#pragma unroll 4
        for (int j = jq; j < nf; j += 4) {
            const int idx = j * LD + r0;
            double c = C[idx];
            double q = Q[idx];
            if (OP == OP_SPARSE) {
                c = fma(op.c0, __dmul_rn(s0, sv[j]), c);
            } else if (OP == OP_FULL) {
                c = fma(op.c0, __dmul_rn(s0, sv[j]), c);
                q = fma(op.c1, __dmul_rn(e0, ev[j]), q);
            } else if (OP == OP_FULL_DEL) {
                const double sj = sv[j], ej = ev[j], qsj = qs[j], qcj = qc[j];
                // the source of row / column loc is the new point's row, which is zero before the update
                const bool keep = keep0 && j != op.loc;
                const double c1 = fma(op.c0, __dmul_rn(s0, sj), keep ? c : 0.0);
                const double q1 = fma(op.c1, __dmul_rn(e0, ej), keep ? q : 0.0);
                const double u = __dmul_rn(qs0, qsj);
                const double v = __dmul_rn(qc0, qcj);
                c = __dadd_rn(c1, fma(u, op.iq, -__dmul_rn(v, op.iqc)));
                q = fma(-u, op.iq, q1);
            }
            if (OP != OP_NONE) C[idx] = c;
            if (OP == OP_FULL || OP == OP_FULL_DEL) Q[idx] = q;
            if (do_mv) {
                const double kj = kn[j];
                ac = fma(c, kj, ac);
                aq = fma(q, kj, aq);
            }
        }
    }
    if (do_mv) {  // (a0 + a1) + (a2 + a3): classes 0/1 and 2/3 differ in lane bit 3, the pairs in lane bit 4
        ac = __dadd_rn(ac, shfl_xor_d(ac, 8));
        aq = __dadd_rn(aq, shfl_xor_d(aq, 8));
        ac = __dadd_rn(ac, shfl_xor_d(ac, 16));
        aq = __dadd_rn(aq, shfl_xor_d(aq, 16));
        if (jq == 0 && r0 < nf) { ck[r0] = ac; ek[r0] = aq; }
    }
}

// LD: storage leading dimension (= 8 mod 16, see fused_pass); NT = 32 * ceil(LD / 8) threads; LD_IN: leading dimension
// of the hand-off slots this bucket resumes; LD_OUT: of the slots it writes when a patch outgrows a.ld.
template <int LD, int NT, int LD_IN, int LD_OUT, bool SPILL>
__global__ void __launch_bounds__(NT) sogp_fit_fused_kernel(SogpArgs a) {
    constexpr int DOUT = 1;
    static_assert(LD % 16 == 8 && NT * 8 >= 32 * LD && NT >= LD + 1, "thread mapping");
    extern __shared__ double smem_dyn[];
    // SPILL (bucket 4): C and Q live in a per-CTA slice of global memory (L2-resident); the fused pass touches every
    // element once per point there too, and the block barriers order the accesses as they do for shared memory
    typedef typename std::conditional<SPILL, SmemSplit<LD>, Smem<LD, 1>>::type State;
    constexpr int kStateDoubles = SPILL ? SmemSplit<LD>::kVecDoubles : Smem<LD, 1>::kDoubles;   // what lives in shared memory
    State s;
    s.base = smem_dyn;
    if constexpr (SPILL) s.mat = a.spill + (size_t)blockIdx.x * (2 * LD * LD);
    // vectors after the state block and its bidx ints (LD ints = LD/2 doubles)
    const FusedVecs<LD> fv{smem_dyn + kStateDoubles + LD / 2};
    double* const C = s.C();
    double* const Q = s.Q();
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t patch = a.patch_ids ? (int64_t)a.patch_ids[blockIdx.x] : a.first_patch + blockIdx.x;
    const int64_t o = a.off[patch];
    const int n = (int)(a.off[patch + 1] - o);
    const int64_t op_ = patch - a.out_first;
    if (n == 0) {
        if (t == 0) { a.nbv[op_] = 0; a.flags[op_] = 0; }
        return;
    }
    for (int i = t; i < kStateDoubles + LD / 2 + FusedVecs<LD>::kDoubles; i += NT) smem_dyn[i] = 0.0;
    if constexpr (SPILL)
        for (int i = t; i < 2 * LD * LD; i += NT) s.mat[i] = 0.0;
    __syncthreads();
    for (int i = t; i < LD; i += NT) s.bidx()[i] = -1;
    __syncthreads();

    const double kstar = a.p0, s20 = a.s20, p0 = a.p0, cl = a.cl, eps_tol = a.eps_tol;
    const int cap = a.capacity, ldmax = a.ld;
    int N = 0, tt0 = 0;
    Counters cnt;
    cnt.init();
    if (a.handoff_in) {  // resume a patch that outgrew the previous bucket
        const double* slot = a.handoff_in + (size_t)blockIdx.x * slot_doubles(LD_IN, DOUT);
        N = reinterpret_cast<const int*>(slot)[0];
        tt0 = reinterpret_cast<const int*>(slot)[1];
#pragma unroll
        for (int i = 0; i < NCNT; i++) cnt.c[i] = reinterpret_cast<const unsigned long long*>(slot + 2)[i];
        const double* v = slot + 2 + NCNT;
        for (int i = t; i < LD_IN; i += NT) {
            s.alpha(0)[i] = v[i];
            s.b1()[i] = v[DOUT * LD_IN + i]; s.b2()[i] = v[(DOUT + 1) * LD_IN + i];
            s.bidx()[i] = reinterpret_cast<const int*>(v + (DOUT + 2) * LD_IN + 2 * LD_IN * LD_IN)[i];
        }
        for (int e = t; e < LD_IN * LD_IN; e += NT) {
            const int j = e / LD_IN, i = e - j * LD_IN;
            C[j * LD + i] = v[(DOUT + 2) * LD_IN + e];
            Q[j * LD + i] = v[(DOUT + 2) * LD_IN + LD_IN * LD_IN + e];
        }
        __syncthreads();
    }
    // k of the current point and of the next one alternate between two buffers
    double* kcur = s.kv();
    double* knext = fv.kn();
    bool have_mv = false;
    FusedOp fo;
    fo.c0 = fo.c1 = fo.iq = fo.iqc = 0.0;
    fo.loc = -1;

    double nx1 = a.fx1[o + tt0], nx2 = a.fx2[o + tt0], ny = a.fy[0][o + tt0];
    int norig = a.forig[o + tt0];
    for (int tt = tt0; tt < n; ++tt) {
        const double x1 = nx1, x2 = nx2, y = ny;
        const int orig = norig;
        const bool last = tt + 1 == n;
        if (!last) {  // prefetch the next point of the stream
            nx1 = a.fx1[o + tt + 1]; nx2 = a.fx2[o + tt + 1]; ny = a.fy[0][o + tt + 1];
            norig = a.forig[o + tt + 1];
        }
        if (N == 0) {  // sparse_gp.hpp:100-110
            if (t == 0) {
                const double d = __dadd_rn(kstar, s20);
                s.alpha(0)[0] = __ddiv_rn(y, d);
                C[0] = __ddiv_rn(-1.0, d);
                Q[0] = __ddiv_rn(1.0, kstar);
                s.b1()[0] = x1; s.b2()[0] = x2; s.bidx()[0] = orig;
            }
            N = 1;
            cnt.c[0]++;
            have_mv = false;
            __syncthreads();
            continue;
        }
        if (!have_mv) {  // first point after the start, a resume or a rare event: k and the matvecs on their own
            if (t < N) kcur[t] = rbf(x1, x2, s.b1()[t], s.b2()[t], p0, cl);
            __syncthreads();
            fused_pass<LD, OP_NONE>(C, Q, N, fo, nullptr, nullptr, nullptr, nullptr, kcur, true, s.ck(), s.ev(), t);
            __syncthreads();
        }
        // every warp: k'Ck and k'e_hat -> gamma and the sparse / full decision; warp 0: m = alpha'k and the divisions
        double kck = 0.0, ke = 0.0;
        for (int j = lane; j < N; j += 32) {
            const double kj = kcur[j];
            kck = fma(kj, s.ck()[j], kck);
            ke = fma(kj, s.ev()[j], ke);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            kck = __dadd_rn(kck, shfl_xor_d(kck, off));
            ke = __dadd_rn(ke, shfl_xor_d(ke, off));
        }
        double gamma = __dadd_rn(kstar, -ke);                 // sparse_gp.hpp:144
        if (gamma < tiny12()) gamma = 0.0;
        const bool sparse = gamma < eps_tol;
        // scal[0] = r, [1] = r*eta (sparse) | 1/gamma (full), [2] = q, [5] = q*eta (sparse)
        if (w == 0) {
            const double m = warp_dot32(s.alpha(0), kcur, N, lane);
            const double s2 = __dadd_rn(kstar, kck);
            const double den = __dadd_rn(s20, s2);
            const double rr = __ddiv_rn(-1.0, den);               // gaussian_noise.cpp:15-18
            const double q = __ddiv_rn(__dadd_rn(y, -m), den);    // gaussian_noise.cpp:9-12
            double qe = 0.0, c3;
            if (sparse) {
                const double eta = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(gamma, rr)));
                qe = __dmul_rn(q, eta);
                c3 = __dmul_rn(rr, eta);
            } else {
                c3 = __ddiv_rn(1.0, gamma);
            }
            if (lane == 0) { s.scal()[0] = rr; s.scal()[1] = c3; s.scal()[2] = q; s.scal()[5] = qe; }
        }
        if (sparse) {
            // sparse update (sparse_gp.hpp:155-163) fused with the next point's matvecs
            cnt.run++;
            double sh = 0.0;
            if (t < N) {
                sh = __dadd_rn(s.ck()[t], s.ev()[t]);
                s.sv()[t] = sh;
                if (!last) knext[t] = rbf(nx1, nx2, s.b1()[t], s.b2()[t], p0, cl);
            }
            __syncthreads();
            fo.c0 = s.scal()[1];
            if (t < N) s.alpha(0)[t] = __dadd_rn(s.alpha(0)[t], __dmul_rn(sh, s.scal()[5]));
            fused_pass<LD, OP_SPARSE>(C, Q, N, fo, s.sv(), nullptr, nullptr, nullptr, knext, !last, s.ck(), s.ev(), t);
            __syncthreads();
            have_mv = !last;
            { double* tmp = kcur; kcur = knext; knext = tmp; }
            continue;  // Q and N unchanged: neither deletion loop can fire
        }
        // full update (sparse_gp.hpp:164-203)
        cnt.flush(N);
        if (N + 1 > ldmax) {  // does not fit this bucket: hand the state to the next one (the matrices are up to date)
            __shared__ int spos;
            if (t == 0) spos = atomicAdd(a.queue_count, 1);
            __syncthreads();
            const int pos = spos;
            GPC_DASSERT(a.queue != nullptr && a.handoff_out != nullptr && pos >= 0 && pos < a.n_work);  // one slot per patch of this launch at most
                double* slot = a.handoff_out + (size_t)pos * slot_doubles(LD_OUT, DOUT);
            if (t == 0) {
                a.queue[pos] = (int32_t)patch;
                reinterpret_cast<int*>(slot)[0] = N;
                reinterpret_cast<int*>(slot)[1] = tt;
                for (int i = 0; i < NCNT; i++) reinterpret_cast<unsigned long long*>(slot + 2)[i] = cnt.c[i];
            }
            double* v = slot + 2 + NCNT;
            for (int i = t; i < LD_OUT; i += NT) {
                v[i] = s.alpha(0)[i];
                v[DOUT * LD_OUT + i] = s.b1()[i]; v[(DOUT + 1) * LD_OUT + i] = s.b2()[i];
                reinterpret_cast<int*>(v + (DOUT + 2) * LD_OUT + 2 * LD_OUT * LD_OUT)[i] = s.bidx()[i];
            }
            for (int e = t; e < LD_OUT * LD_OUT; e += NT) {
                const int j = e / LD_OUT, i = e - j * LD_OUT;
                v[(DOUT + 2) * LD_OUT + e] = C[j * LD + i];
                v[(DOUT + 2) * LD_OUT + LD_OUT * LD_OUT + e] = Q[j * LD + i];
            }
            return;
        }
        cnt.full(N);
        const int N1 = N + 1;
        double sct = 0.0;
        if (t < N) {
            sct = s.ck()[t];
            s.sv()[t] = sct;
        }
        if (t == N) {
            s.sv()[N] = 1.0;
            s.ev()[N] = -1.0;
            s.b1()[N] = x1; s.b2()[N] = x2; s.bidx()[N] = orig;
        }
        __syncthreads();
        const double rr = s.scal()[0], ig = s.scal()[1];
        fo.c0 = rr; fo.c1 = ig;
        // alpha after the full update, and the diagonals of the updated C and Q (O(N), no matrix pass yet)
        double a1 = 0.0, dc = 0.0, dq = 0.0;
        if (t < N1) {
            const double q = s.scal()[2];
            a1 = (t < N) ? __dadd_rn(s.alpha(0)[t], __dmul_rn(q, sct)) : __dadd_rn(0.0, __dmul_rn(q, 1.0));
            s.alpha(0)[t] = a1;
            const double st_ = s.sv()[t], et_ = s.ev()[t];
            dc = fma(rr, __dmul_rn(st_, st_), C[t * LD + t]);
            dq = fma(ig, __dmul_rn(et_, et_), Q[t * LD + t]);
        }
        if (N1 <= cap) {
            // no capacity deletion; geometric deletions (sparse_gp.hpp:226-242) are decided on the updated diagonal
            const int geo = __syncthreads_or(t < N1 && __ddiv_rn(1.0, dq) < geo9());
            N = N1;
            if (geo) {  // rare: bring the matrices up to date, then the step-by-step loop
                fused_pass<LD, OP_FULL>(C, Q, N, fo, s.sv(), s.ev(), nullptr, nullptr, knext, false, s.ck(), s.ev(), t);
                __syncthreads();
                while (N > 1) {
                    const int loc = block_argmin<LD, DOUT, 1>(s, N, t, lane);
                    if (loc < 0) break;
                    cnt.c[9] += (unsigned long long)(N - 1) * (N - 1);
                    delete_bv_any<LD, NT, DOUT>(s, N, loc, t);
                    cnt.c[4]++;
                }
                have_mv = false;
                continue;
            }
            if (!last && t < N) knext[t] = rbf(nx1, nx2, s.b1()[t], s.b2()[t], p0, cl);
            __syncthreads();
            // ek aliases ev: the pass reads ev[j] for all j while lanes jq == 0 write ek[r0] at its end, after the
            // shuffles that every lane of the warp reaches only when its own loop is done; other warps may still be
            // reading ev, so the row sums go to sv / qsv first and are copied below
            fused_pass<LD, OP_FULL>(C, Q, N, fo, s.sv(), s.ev(), nullptr, nullptr, knext, !last, s.qsv(), s.qcv(), t);
            __syncthreads();
            if (!last && t < N) { s.ck()[t] = s.qsv()[t]; s.ev()[t] = s.qcv()[t]; }
            __syncthreads();
            have_mv = !last;
            { double* tmp = kcur; kcur = knext; knext = tmp; }
            continue;
        }
        // capacity deletion (sparse_gp.hpp:206-223): scores from the updated alpha and diagonals, then ONE pass
        if (t < N1) s.scv()[t] = __ddiv_rn(__dmul_rn(a1, a1), __dadd_rn(dq, dc));
        __syncthreads();
        const int loc = scan_first_min(s.scv(), N1, lane);
        const int L = N, M = N;   // the new point sits at index L; M entries remain
        cnt.c[9] += (unsigned long long)M * M;
        cnt.c[3]++;
        fo.loc = loc;
        double ai = 0.0, qsi = 0.0, qci = 0.0, dqn = 0.0;
        if (t < M) {
            const int src = (t == loc) ? L : t;  // the last entry moves into the deleted slot
            const double cold = (src == L || loc == L) ? 0.0 : C[src * LD + loc];
            const double qold = (src == L || loc == L) ? 0.0 : Q[src * LD + loc];
            const double ssrc = s.sv()[src], esrc = s.ev()[src];
            const double csi = fma(rr, __dmul_rn(s.sv()[loc], ssrc), cold);   // row loc of the updated C, permuted
            qsi = fma(ig, __dmul_rn(s.ev()[loc], esrc), qold);
            qci = __dadd_rn(qsi, csi);
            ai = s.alpha(0)[src];
            // diagonal of the updated Q at the source position: feeds the geometric test on the final Q
            const double qdd = (src == L) ? 0.0 : Q[src * LD + src];
            dqn = fma(ig, __dmul_rn(esrc, esrc), qdd);
            fv.svp()[t] = ssrc;
            fv.evp()[t] = esrc;
        }
        if (w == 0) {
            const double cll = (loc == L) ? 0.0 : C[loc * LD + loc], qll = (loc == L) ? 0.0 : Q[loc * LD + loc];
            const double sl = s.sv()[loc], el = s.ev()[loc];
            const double cstar = fma(rr, __dmul_rn(sl, sl), cll), qstar = fma(ig, __dmul_rn(el, el), qll);
            const double qcs = __dadd_rn(qstar, cstar);
            const double coef = __ddiv_rn(s.alpha(0)[loc], qcs);
            const double iq = __ddiv_rn(1.0, qstar), iqc = __ddiv_rn(1.0, qcs);
            if (lane == 0) { s.scal()[8] = iq; s.scal()[9] = iqc; s.scal()[11] = coef; }
        }
        __syncthreads();  // every read of alpha / sv / ev by index src is done
        fo.iq = s.scal()[8]; fo.iqc = s.scal()[9];
        if (t < M) {
            s.alpha(0)[t] = __dadd_rn(ai, -__dmul_rn(s.scal()[11], qci));   // sparse_gp.hpp:285
            s.qsv()[t] = qsi;
            s.qcv()[t] = qci;
            if (t == loc && loc != L) { s.b1()[loc] = x1; s.b2()[loc] = x2; s.bidx()[loc] = orig; }
            dqn = fma(-__dmul_rn(qsi, qsi), fo.iq, dqn);   // Q_tt after the deletion
        }
        if (t == L) { s.alpha(0)[L] = 0.0; s.b1()[L] = 0.0; s.b2()[L] = 0.0; s.bidx()[L] = -1; }
        const int geo = __syncthreads_or(t < M && M > 1 && __ddiv_rn(1.0, dqn) < geo9());
        if (!last && !geo && t < M) knext[t] = rbf(nx1, nx2, s.b1()[t], s.b2()[t], p0, cl);
        __syncthreads();
        fused_pass<LD, OP_FULL_DEL>(C, Q, M, fo, fv.svp(), fv.evp(), s.qsv(), s.qcv(), knext, !last && !geo, s.ck(), s.ev(), t);
        N = M;
        fo.loc = -1;
        __syncthreads();
        if (geo) {  // rare: geometric deletions on the up-to-date state (sparse_gp.hpp:226-242)
            while (N > 1) {
                const int l2 = block_argmin<LD, DOUT, 1>(s, N, t, lane);
                if (l2 < 0) break;
                cnt.c[9] += (unsigned long long)(N - 1) * (N - 1);
                delete_bv_any<LD, NT, DOUT>(s, N, l2, t);
                cnt.c[4]++;
            }
            have_mv = false;
            continue;
        }
        have_mv = !last;
        { double* tmp = kcur; kcur = knext; knext = tmp; }
    }
    cnt.flush(N);
    __syncthreads();
    // results
    if (t == 0) {
        a.nbv[op_] = N;
        const double c00 = C[0];
        a.flags[op_] = (c00 != c00) ? 1 : 0;
        publish(a, cnt, n);
    }
    const int64_t ob = op_ * cap;
    for (int i = t; i < N; i += NT) {
        a.o_alpha[0][ob + i] = s.alpha(0)[i];
        a.o_b1[ob + i] = s.b1()[i];
        a.o_b2[ob + i] = s.b2()[i];
        a.o_idx[ob + i] = s.bidx()[i];
    }
    if (a.dumpC) {
        const int64_t od = op_ * (int64_t)cap * cap;
        for (int e = t; e < N * N; e += NT) {
            const int i = e / N, j = e - i * N;
            a.dumpC[od + e] = C[j * LD + i];
            a.dumpQ[od + e] = Q[j * LD + i];
        }
    }
}

template <int LD, bool SPILL>
constexpr size_t fused_smem_bytes() {
    return (size_t)((SPILL ? SmemSplit<LD>::kVecDoubles : Smem<LD, 1>::kDoubles) + LD / 2 + FusedVecs<LD>::kDoubles) * sizeof(double);
}

template <int LD, int NT, int LD_IN, int LD_OUT, bool SPILL = false>
cudaError_t launch_fused_bucket(const SogpArgs& a, cudaStream_t st) {
    constexpr size_t smem = fused_smem_bytes<LD, SPILL>();
    if (smem > 48 * 1024) {
        cudaError_t e = GPC_FUNC_ATTR_ONCE((sogp_fit_fused_kernel<LD, NT, LD_IN, LD_OUT, SPILL>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    sogp_fit_fused_kernel<LD, NT, LD_IN, LD_OUT, SPILL><<<a.n_work, NT, smem, st>>>(a);
    return cudaGetLastError();
}

template <int LD, int DOUT>
constexpr size_t cta_smem_bytes() { return (size_t)Smem<LD, DOUT>::kDoubles * sizeof(double) + (size_t)LD * sizeof(int); }

template <int LD, int RB, int NT, int LD_IN, bool SPILL, int DOUT>
cudaError_t launch_cta_bucket(const SogpArgs& a, cudaStream_t st) {
    constexpr size_t smem = SPILL ? 0 : cta_smem_bytes<LD, DOUT>();
    if (smem > 48 * 1024) {
        cudaError_t e = GPC_FUNC_ATTR_ONCE((sogp_fit_kernel<LD, RB, NT, LD_IN, SPILL, DOUT>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    sogp_fit_kernel<LD, RB, NT, LD_IN, SPILL, DOUT><<<a.n_work, NT, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

// ---- continued fits (sparse_gp::add_measurements called again, sparse_gp.hpp:59-86) ---------------------------------
// The kept state of a patch (alpha, BV, C, Q of the previous fit) is written as a hand-off slot of the smallest format
// that holds it; the bucket that resumes that format continues the recursion on the new points.
__global__ void continue_partition_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ nbv, int64_t lo, int64_t n,
                                          int32_t* __restrict__ queues, int32_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (off[lo + i + 1] == off[lo + i]) return;  // no new points: the patch keeps its state
    const int N = nbv[i];
    const int b0 = N <= 16 ? 1 : (N <= 32 ? 2 : (N <= 64 ? 3 : 4));
    const int pos = atomicAdd(counts + (b0 - 1), 1);
    queues[(int64_t)(b0 - 1) * n + pos] = (int32_t)(lo + i);
}

// one CTA per queued patch: slot q of format LDS <- state of patch ids[q]
template <int LDS>
__global__ void __launch_bounds__(128) state_to_slot_kernel(const int32_t* __restrict__ ids, int64_t out_first, int cap,
                                                            const int32_t* __restrict__ nbv, const double* __restrict__ alpha,
                                                            const double* __restrict__ b1, const double* __restrict__ b2,
                                                            const int32_t* __restrict__ bidx, const double* __restrict__ dumpC,
                                                            const double* __restrict__ dumpQ, double* __restrict__ slots) {
    const int64_t op = (int64_t)ids[blockIdx.x] - out_first;
    const int N = nbv[op];
    double* slot = slots + (size_t)blockIdx.x * slot_doubles(LDS, 1);
    const int t = threadIdx.x;
    if (t == 0) { reinterpret_cast<int*>(slot)[0] = N; reinterpret_cast<int*>(slot)[1] = 0; }
    if (t < NCNT) reinterpret_cast<unsigned long long*>(slot + 2)[t] = 0ull;
    double* v = slot + 2 + NCNT;
    const int64_t ob = op * cap, od = op * (int64_t)cap * cap;
    for (int i = t; i < LDS; i += 128) {
        v[i] = i < N ? alpha[ob + i] : 0.0;
        v[LDS + i] = i < N ? b1[ob + i] : 0.0;
        v[2 * LDS + i] = i < N ? b2[ob + i] : 0.0;
        reinterpret_cast<int*>(v + 3 * LDS + 2 * LDS * LDS)[i] = i < N ? bidx[ob + i] : -1;
    }
    for (int e = t; e < LDS * LDS; e += 128) {
        const int j = e / LDS, i = e - j * LDS;   // slot element e = (row i, column j)
        const bool in = i < N && j < N;
        v[3 * LDS + e] = in ? dumpC[od + (size_t)i * N + j] : 0.0;
        v[3 * LDS + LDS * LDS + e] = in ? dumpQ[od + (size_t)i * N + j] : 0.0;
    }
}

void launch_continue_partition(const int64_t* off, const int32_t* nbv, int64_t lo, int64_t n, int32_t* queues, int32_t* counts4,
                               cudaStream_t s) {
    cudaMemsetAsync(counts4, 0, 4 * sizeof(int32_t), s);
    if (n <= 0) return;
    continue_partition_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(off, nbv, lo, n, queues, counts4);
    g_launches++;
}

// start bucket b0 in 1..4 resumes slots of bucket b0 - 1's format
cudaError_t launch_state_to_slots(int b0, const int32_t* ids, int64_t n_work, int64_t out_first, int cap, const int32_t* nbv,
                                  const double* alpha, const double* b1, const double* b2, const int32_t* bidx, const double* dumpC,
                                  const double* dumpQ, double* slots, cudaStream_t s) {
    if (n_work <= 0) return cudaSuccess;
    g_launches++;
    switch (b0) {
        case 1: state_to_slot_kernel<16><<<(unsigned)n_work, 128, 0, s>>>(ids, out_first, cap, nbv, alpha, b1, b2, bidx, dumpC, dumpQ, slots); break;
        case 2: state_to_slot_kernel<32><<<(unsigned)n_work, 128, 0, s>>>(ids, out_first, cap, nbv, alpha, b1, b2, bidx, dumpC, dumpQ, slots); break;
        case 3: state_to_slot_kernel<64><<<(unsigned)n_work, 128, 0, s>>>(ids, out_first, cap, nbv, alpha, b1, b2, bidx, dumpC, dumpQ, slots); break;
        default: state_to_slot_kernel<118><<<(unsigned)n_work, 128, 0, s>>>(ids, out_first, cap, nbv, alpha, b1, b2, bidx, dumpC, dumpQ, slots); break;
    }
    return cudaGetLastError();
}

int sogp_bucket_ld(int bucket) {
    static const int lds[5] = {16, 32, 64, 118, 202};
    return lds[bucket];
}


size_t sogp_handoff_slot_bytes(int bucket, int dout) { return (size_t)slot_doubles(sogp_bucket_ld(bucket), dout) * sizeof(double); }

size_t sogp_spill_bytes_per_patch() { return (size_t)(2 * 216 * 216) * sizeof(double); }  // >= the generic kernel's (2 * 202^2 + 13 * 202 + 16)

// Height GPs (dout 1) use buckets 0,1,2,3,4; the RGB field GPs (dout 3) use 0,2,4 (bucket 2 resumes bucket-0 slots,
// bucket 4 resumes bucket-2 slots): under the reference's field hyper-parameters N stays below 16 anyway.
int sogp_next_bucket(int bucket, int dout) {
    if (dout == 1) return bucket + 1;
    return bucket == 0 ? 2 : 4;
}

cudaError_t launch_sogp_fit(int bucket, const SogpArgs& a, cudaStream_t st) {
    if (a.n_work <= 0) return cudaSuccess;
    g_launches++;
    if (bucket == 0) {  // 20 one-warp blocks per SM need the largest shared-memory carve-out
        cudaError_t e = a.dout == 3 ? GPC_FUNC_ATTR_ONCE((sogp_fit_half_kernel<3>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
                                    : GPC_FUNC_ATTR_ONCE((sogp_fit_half_kernel<1>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    if (a.dout == 3) {
        switch (bucket) {
            case 0:
                sogp_fit_half_kernel<3><<<(a.n_work + 1) / 2, 32, 0, st>>>(a);
                return cudaGetLastError();
            case 2: return launch_cta_bucket<64, 64, 256, 16, false, 3>(a, st);
            default: return launch_cta_bucket<202, 256, 1024, 64, true, 3>(a, st);
        }
    }
    switch (bucket) {
        case 0:
            sogp_fit_half_kernel<1><<<(a.n_work + 1) / 2, 32, 0, st>>>(a);
            return cudaGetLastError();
        case 1:
            sogp_fit_pair_kernel<<<a.n_work, 64, 0, st>>>(a);
            return cudaGetLastError();
        case 2: return launch_fused_bucket<72, 288, 32, 64>(a, st);
        case 3:
            if (a.ld <= 104) return launch_fused_bucket<104, 416, 64, 104>(a, st);
            return launch_cta_bucket<118, 128, 512, 64, false, 1>(a, st);
        default: return launch_fused_bucket<216, 864, 118, 216, true>(a, st);
    }
}

}  // namespace gpc
