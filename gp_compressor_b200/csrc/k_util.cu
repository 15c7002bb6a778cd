// k_util.cu — scans, the per-patch shuffle (K6b) and the gather into the fit stream (K6c).
#include "gpc_device.cuh"
#include <algorithm>

#include "gpc_internal.h"

namespace gpc {

// ------------------------------------------------------------------------------------
// Exclusive scan of int64 (three launches: tile sums, scan of sums, apply).
// ------------------------------------------------------------------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_E = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_E;

__device__ inline int64_t block_exclusive_scan(int64_t v, int64_t* total) {
    __shared__ int64_t warp_sums[SCAN_T / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int64_t s = (lane < SCAN_T / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += y;
        }
        if (lane < SCAN_T / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    int64_t base = (wid > 0) ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_T / 32 - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_sums(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ sums) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_E;
    int64_t s = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; e++)
        if (base + e < n) s += in[base + e];
    int64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) scan_sums_inplace(int64_t* sums, int64_t n_tiles, int64_t* grand_total) {
    int64_t carry = 0;
    for (int64_t b = 0; b < n_tiles; b += SCAN_T) {
        int64_t i = b + threadIdx.x;
        int64_t v = (i < n_tiles) ? sums[i] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_T) scan_apply(const int64_t* __restrict__ in, int64_t n, const int64_t* __restrict__ sums,
                                                     int64_t* __restrict__ out) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_E;
    int64_t v[SCAN_E];
    int64_t s = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) {
        v[e] = (base + e < n) ? in[base + e] : 0;
        s += v[e];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, &total) + sums[blockIdx.x];
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) {
        if (base + e < n) out[base + e] = ex;
        ex += v[e];
    }
}

size_t scan_tmp_bytes(int64_t n) {
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(tiles + 1) * sizeof(int64_t);
}

void launch_exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* tmp, cudaStream_t s) {
    if (n <= 0) {
        cudaMemsetAsync(out, 0, sizeof(int64_t), s);
        return;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* sums = reinterpret_cast<int64_t*>(tmp);
    scan_tile_sums<<<(unsigned)tiles, SCAN_T, 0, s>>>(in, n, sums);
    scan_sums_inplace<<<1, SCAN_T, 0, s>>>(sums, tiles, out + n);
    scan_apply<<<(unsigned)tiles, SCAN_T, 0, s>>>(in, n, sums, out);
    g_launches += 3;
}

// ------------------------------------------------------------------------------------
// rand() draws per patch: the height GP shuffles n_p points with n_p - 1 draws, then the
// RGB field GP does the same (gp_compressor.cpp:162-163), so mult = 2 in the full path.
// ------------------------------------------------------------------------------------
__global__ void patch_draws_kernel(const int64_t* __restrict__ off, int64_t n_patches, int mult, int64_t* __restrict__ draws,
                                   unsigned long long* __restrict__ max_np) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = 0;
    if (p < n_patches) {
        n = off[p + 1] - off[p];
        draws[p] = (n > 0) ? (n - 1) * mult : 0;
    }
    // largest patch (selects the shuffle kernel): warp max, one atomic per warp
    unsigned long long m = (unsigned long long)n;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_np, m);
}

// Shard bounds and the rand-stream window of the shard, on the device (no per-patch work on the host):
// plan = { n_claimed, lo, hi, off[lo], off[hi], roff[lo], roff[hi], roff[P] }.  Same rule as gpc_shard_range:
// bound k = first p with off[p] >= floor(total * k / count).
__global__ void fit_plan_kernel(const int64_t* __restrict__ off, const int64_t* __restrict__ roff, int64_t P, int rank, int count,
                                int64_t fixed_lo, int64_t fixed_hi, int64_t* __restrict__ plan) {
    const int64_t total = off[P];
    int64_t b[2];
    for (int e = 0; e < 2; e++) {
        if (fixed_lo >= 0) { b[e] = e ? fixed_hi : fixed_lo; continue; }  // sharded binning: the owned range is given
        const int k = rank + e;
        if (k <= 0) { b[e] = 0; continue; }
        if (k >= count) { b[e] = P; continue; }
        const int64_t target = (total / count) * k + ((total % count) * k) / count;
        int64_t lo = 0, hi = P;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (off[mid] >= target) hi = mid; else lo = mid + 1;
        }
        b[e] = lo;
    }
    plan[0] = total; plan[1] = b[0]; plan[2] = b[1]; plan[3] = off[b[0]]; plan[4] = off[b[1]];
    plan[5] = roff[b[0]]; plan[6] = roff[b[1]]; plan[7] = roff[P];
}

void launch_fit_plan(const int64_t* off, int64_t n_patches, int mult, int rank, int count, int64_t fixed_lo, int64_t fixed_hi,
                     int64_t* draws, int64_t* roff, void* scan_tmp, int64_t* plan9, cudaStream_t s) {
    cudaMemsetAsync(plan9 + 8, 0, sizeof(int64_t), s);
    if (n_patches > 0) {
        patch_draws_kernel<<<(unsigned)((n_patches + 255) / 256), 256, 0, s>>>(off, n_patches, mult, draws,
                                                                             reinterpret_cast<unsigned long long*>(plan9 + 8));
        g_launches++;
    }
    launch_exclusive_scan_i64(draws, roff, n_patches, scan_tmp, s);
    fit_plan_kernel<<<1, 1, 0, s>>>(off, roff, n_patches, rank, count, fixed_lo, fixed_hi, plan9);
    g_launches++;
}

// patch_of[s] = the patch whose range holds stream element s; perm[s] = identity (patch-local)
// (off points at the first patch of the shard; stream indices stay absolute)
__global__ void patch_of_kernel(const int64_t* __restrict__ off, int64_t n_patches, int64_t s_begin, int64_t s_count,
                                int32_t* __restrict__ patch_of, int32_t* __restrict__ perm) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s_count) return;
    s += s_begin;
    // last p with off[p] <= s
    int64_t lo = 0, hi = n_patches;  // invariant: off[lo] <= s < off[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= s) lo = mid; else hi = mid;
    }
    patch_of[s] = (int32_t)lo;
    perm[s] = (int32_t)(s - off[lo]);
}

// sparse_gp::shuffle, sparse_gp.hpp:42-56: for i = n-1..1: r = rand() % i; swap(ind[i], ind[r]).
// rnd[roff[p] - roff[0] + t] is draw t of patch p.  Two kernels share the work by patch size:
//  * n <= SHUF_SMEM_MAX: one warp per patch, index array and r_i = draw_i % i in shared memory; every lane
//    computes its share of the modulos, lane 0 walks the (inherently serial) swap chain at shared-memory latency;
//  * larger patches: one thread per patch, swaps in global memory.
constexpr int SHUF_SMEM_MAX = 1024;    // per-patch limit of the separate shuffle_warp_kernel (32-bit indices)
constexpr int SHUF_FUSED_MAX = 8192;   // per-patch limit of the fused shuffle + gather kernel (16-bit indices, <= 32 KB)

__global__ void __launch_bounds__(32) shuffle_warp_kernel(const int64_t* __restrict__ off, int64_t n_patches,
                                                          const int64_t* __restrict__ roff, const uint32_t* __restrict__ rnd,
                                                          int second, int32_t* __restrict__ perm) {
    __shared__ int ind[SHUF_SMEM_MAX];
    __shared__ int rr[SHUF_SMEM_MAX];
    const int64_t p = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t o = off[p];
    const int n = (int)(off[p + 1] - o);
    if (n > SHUF_SMEM_MAX) return;
    if (n < 2) {
        if (n == 1 && lane == 0) perm[o] = 0;
        return;
    }
    // the RGB field GP shuffles right after the height GP: its draws follow the patch's first n - 1
    const uint32_t* r = rnd + (roff[p] - roff[0]) + (second ? (n - 1) : 0);
    for (int i = lane; i < n; i += 32) {
        ind[i] = i;
        if (i > 0) rr[i] = (int)(r[n - 1 - i] % (uint32_t)i);
    }
    __syncwarp();
    if (lane == 0) {
        for (int i = n - 1; i > 0; --i) {
            const int j = rr[i];
            const int a = ind[i], b = ind[j];
            ind[i] = b;
            ind[j] = a;
        }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) perm[o + i] = ind[i];
}

__global__ void shuffle_kernel(const int64_t* __restrict__ off, int64_t n_patches, const int64_t* __restrict__ roff,
                               const uint32_t* __restrict__ rnd, int second, int32_t* __restrict__ perm) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_patches) return;
    const int64_t o = off[p];
    const int n = (int)(off[p + 1] - o);
    if (n <= SHUF_SMEM_MAX) return;  // handled by shuffle_warp_kernel
    const uint32_t* r = rnd + (roff[p] - roff[0]) + (second ? (n - 1) : 0);
    int32_t* ind = perm + o;
    if (second)
        for (int i = 0; i < n; i++) ind[i] = i;
    for (int i = n - 1; i > 0; --i) {
        uint32_t rr = r[n - 1 - i] % (uint32_t)i;
        int32_t a = ind[i], b = ind[rr];
        ind[i] = b;
        ind[rr] = a;
    }
}

// fit stream in add order: element t of patch p is point perm[off[p] + t]
__global__ void gather_stream_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ patch_of,
                                     const int32_t* __restrict__ perm, const double* __restrict__ x1,
                                     const double* __restrict__ x2, const double* __restrict__ y, int64_t s_begin,
                                     int64_t s_count, double* __restrict__ fx1, double* __restrict__ fx2,
                                     double* __restrict__ fy) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s_count) return;
    s += s_begin;
    int64_t src = off[patch_of[s]] + perm[s];
    fx1[s] = x1[src];
    fx2[s] = x2[src];
    fy[s] = y[src];
}
void launch_gather_stream(const int64_t* off, const int32_t* patch_of, const int32_t* perm, const double* x1,
                          const double* x2, const double* y, int64_t s_begin, int64_t s_count, double* fx1, double* fx2,
                          double* fy, cudaStream_t s) {
    if (s_count <= 0) return;
    gather_stream_kernel<<<(unsigned)((s_count + 255) / 256), 256, 0, s>>>(off, patch_of, perm, x1, x2, y, s_begin, s_count, fx1, fx2, fy);
    g_launches++;
}

// RGB field GP stream in its own add order: local coordinates and colours centred on the patch mean
// (p.second -= c_mn, gp_compressor.cpp:105).  rgb holds the b,g,r,a bytes of each claimed point.
__global__ void gather_rgb_stream_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ patch_of,
                                         const int32_t* __restrict__ perm, const double* __restrict__ x1,
                                         const double* __restrict__ x2, const uint32_t* __restrict__ rgb,
                                         const double* __restrict__ rgbmean, int64_t first_patch, int64_t s_begin,
                                         int64_t s_count, double* __restrict__ fx1, double* __restrict__ fx2,
                                         double* __restrict__ fr, double* __restrict__ fg, double* __restrict__ fb) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s_count) return;
    s += s_begin;
    const int64_t pl = patch_of[s];
    const int64_t src = off[pl] + perm[s];
    const double* mean = rgbmean + 3 * (first_patch + pl);
    const uint32_t c = rgb[src];
    fx1[s] = x1[src];
    fx2[s] = x2[src];
    fr[s] = __dadd_rn((double)((c >> 16) & 255u), -mean[0]);
    fg[s] = __dadd_rn((double)((c >> 8) & 255u), -mean[1]);
    fb[s] = __dadd_rn((double)(c & 255u), -mean[2]);
}
void launch_gather_rgb_stream(const int64_t* off, const int32_t* patch_of, const int32_t* perm, const double* x1, const double* x2,
                              const uint32_t* rgb, const double* rgbmean, int64_t first_patch, int64_t s_begin, int64_t s_count,
                              double* fx1, double* fx2, double* fr, double* fg, double* fb, cudaStream_t s) {
    if (s_count <= 0) return;
    gather_rgb_stream_kernel<<<(unsigned)((s_count + 255) / 256), 256, 0, s>>>(off, patch_of, perm, x1, x2, rgb, rgbmean, first_patch,
                                                                             s_begin, s_count, fx1, fx2, fr, fg, fb);
    g_launches++;
}

// flags[i] = patch i holds a GP; maxes[0/1] = largest BV count of the height / RGB GPs (sizes the decode tables)
__global__ void __launch_bounds__(256) flag_nonempty_kernel(const int32_t* __restrict__ nbv, const int32_t* __restrict__ rgb_nbv,
                                                            int64_t n, int64_t* __restrict__ flags, int32_t* __restrict__ maxes) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int m0 = 0, m1 = 0;
    if (i < n) {
        m0 = nbv[i];
        flags[i] = m0 > 0 ? 1 : 0;
        if (rgb_nbv) m1 = rgb_nbv[i];
    }
    m0 = __reduce_max_sync(0xffffffffu, m0);
    m1 = __reduce_max_sync(0xffffffffu, m1);
    if ((threadIdx.x & 31) == 0) {
        if (m0 > 0) atomicMax(maxes, m0);
        if (m1 > 0) atomicMax(maxes + 1, m1);
    }
}
void launch_flag_nonempty(const int32_t* nbv, const int32_t* rgb_nbv, int64_t n, int64_t* flags, int32_t* maxes, cudaStream_t s) {
    cudaMemsetAsync(maxes, 0, 2 * sizeof(int32_t), s);
    if (n <= 0) return;
    flag_nonempty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(nbv, rgb_nbv, n, flags, maxes);
    g_launches++;
}

// out34[0] = max nbv, out34[1 + k] = patches with k basis vectors (k < 32), out34[33] = patches with >= 32 (gpc_stats.bv_hist)
__global__ void __launch_bounds__(256) bv_hist_kernel(const int32_t* __restrict__ nbv, int64_t n, unsigned long long* __restrict__ out34) {
    __shared__ unsigned int h[34];
    if (threadIdx.x < 34) h[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = nbv[i];
        atomicAdd(&h[1 + (v < 32 ? v : 32)], 1u);
        atomicMax(&h[0], (unsigned)v);
    }
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(out34, (unsigned long long)h[0]);
    else if (threadIdx.x < 34 && h[threadIdx.x]) atomicAdd(out34 + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}
void launch_bv_hist(const int32_t* nbv, int64_t n, unsigned long long* out34, cudaStream_t s) {
    cudaMemsetAsync(out34, 0, 34 * sizeof(unsigned long long), s);
    if (n <= 0) return;
    bv_hist_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 592), 256, 0, s>>>(nbv, n, out34);
    g_launches++;
}

// points fed to each patch so far: fed[p] = (accumulate ? fed[p] : 0) + off[p + 1] - off[p]
__global__ void fed_update_kernel(const int64_t* __restrict__ off, int64_t n, int accumulate, int64_t* __restrict__ fed) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) fed[p] = (accumulate ? fed[p] : 0) + off[p + 1] - off[p];
}
void launch_fed_update(const int64_t* off, int64_t n, int accumulate, int64_t* fed, cudaStream_t s) {
    if (n <= 0) return;
    fed_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(off, n, accumulate, fed);
    g_launches++;
}
__global__ void forig_kernel(const int32_t* __restrict__ patch_of, const int32_t* __restrict__ perm, const int64_t* __restrict__ orig_base,
                             int64_t s_begin, int64_t s_count, int32_t* __restrict__ forig) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s_count) return;
    s += s_begin;
    forig[s] = perm[s] + (int32_t)orig_base[patch_of[s]];
}

// Fused shuffle + gather for patches of up to SHUF_SMEM_MAX points (the common case): the permutation never leaves
// shared memory before the fit streams are written in add order.  RGB: the field GP's own shuffle (the draws after the
// patch's first n - 1) and its colour stream centred on the patch mean (p.second -= c_mn, gp_compressor.cpp:105).
// cap: capacity of the index arrays (16-bit entries in dynamic shared memory: ind[cap], rr[cap]); patches of more than cap
// points are left to the global-memory kernels.
template <bool RGB>
__global__ void __launch_bounds__(32) shuffle_gather_warp_kernel(ShuffleGatherArgs a, int cap) {
    extern __shared__ unsigned short shuf_smem[];
    unsigned short* const ind = shuf_smem;
    unsigned short* const rr = shuf_smem + cap;
    const int64_t p = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t o = a.off[p];
    const int n = (int)(a.off[p + 1] - o);
    if (n == 0 || n > cap) return;
    if (a.do_shuffle && n >= 2) {
        const uint32_t* r = a.rnd + (a.roff[p] - a.roff[0]) + (RGB ? (n - 1) : 0);
        for (int i = lane; i < n; i += 32) {
            ind[i] = (unsigned short)i;
            if (i > 0) rr[i] = (unsigned short)(r[n - 1 - i] % (uint32_t)i);
        }
        __syncwarp();
        if (lane == 0) {
            for (int i = n - 1; i > 0; --i) {
                const int j = rr[i];
                const unsigned short x = ind[i], y = ind[j];
                ind[i] = y;
                ind[j] = x;
            }
        }
        __syncwarp();
    } else {
        for (int i = lane; i < n; i += 32) ind[i] = (unsigned short)i;
        __syncwarp();
    }
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    if (RGB) { m0 = a.rgbmean[3 * (a.first_patch + p)]; m1 = a.rgbmean[3 * (a.first_patch + p) + 1]; m2 = a.rgbmean[3 * (a.first_patch + p) + 2]; }
    for (int i = lane; i < n; i += 32) {
        const int k = ind[i];
        const int64_t src = o + k, dst = o + i;
        a.perm[dst] = k;
        if (a.forig) a.forig[dst] = k + (int32_t)a.orig_base[p];
        a.fx1[dst] = a.x1[src];
        a.fx2[dst] = a.x2[src];
        if (RGB) {
            const uint32_t c = a.rgb[src];
            a.f0[dst] = __dadd_rn((double)((c >> 16) & 255u), -m0);
            a.f1[dst] = __dadd_rn((double)((c >> 8) & 255u), -m1);
            a.f2[dst] = __dadd_rn((double)(c & 255u), -m2);
        } else {
            a.f0[dst] = a.y[src];
        }
    }
}

// Shuffle (sparse_gp::shuffle, sparse_gp.hpp:42-56) and fit-stream gather of one shard.  Patches of more than
// SHUF_SMEM_MAX points (rare) take the separate global-memory kernels.
void launch_shuffle_gather(const ShuffleGatherArgs& a, int64_t max_patch_points, int32_t* patch_of, cudaStream_t s) {
    if (a.s_count <= 0 || a.n_patches <= 0) return;
    if (max_patch_points <= SHUF_FUSED_MAX) {
        // index arrays sized for the largest patch of the call: 4 KB for the common case (32 CTAs / SM), up to 32 KB
        int cap = 1024;
        while (cap < max_patch_points) cap *= 2;
        const size_t smem = (size_t)cap * 2 * sizeof(unsigned short);
        if (a.is_rgb) shuffle_gather_warp_kernel<true><<<(unsigned)a.n_patches, 32, smem, s>>>(a, cap);
        else shuffle_gather_warp_kernel<false><<<(unsigned)a.n_patches, 32, smem, s>>>(a, cap);
        g_launches++;
        return;
    }
    const int second = a.is_rgb ? 1 : 0;
    patch_of_kernel<<<(unsigned)((a.s_count + 255) / 256), 256, 0, s>>>(a.off, a.n_patches, a.s_begin, a.s_count, patch_of, a.perm);
    g_launches++;
    if (a.do_shuffle) {
        shuffle_warp_kernel<<<(unsigned)a.n_patches, 32, 0, s>>>(a.off, a.n_patches, a.roff, a.rnd, second, a.perm);
        shuffle_kernel<<<(unsigned)((a.n_patches + 127) / 128), 128, 0, s>>>(a.off, a.n_patches, a.roff, a.rnd, second, a.perm);
        g_launches += 2;
    }
    if (a.is_rgb)
        launch_gather_rgb_stream(a.off, patch_of, a.perm, a.x1, a.x2, a.rgb, a.rgbmean, a.first_patch, a.s_begin, a.s_count, a.fx1, a.fx2,
                                 a.f0, a.f1, a.f2, s);
    else
        launch_gather_stream(a.off, patch_of, a.perm, a.x1, a.x2, a.y, a.s_begin, a.s_count, a.fx1, a.fx2, a.f0, s);
    if (a.forig) {
        forig_kernel<<<(unsigned)((a.s_count + 255) / 256), 256, 0, s>>>(patch_of, a.perm, a.orig_base, a.s_begin, a.s_count, a.forig);
        g_launches++;
    }
}

// ---- patch ids of [lo, lo + n) ordered by decreasing point count (counting sort over min(count, 1023)) ----------
// The bucket-0 SOGP kernel packs two patches per warp: neighbours in this order have (almost) equal lengths, and the
// longest patches start first.  The order within one count is arbitrary; results do not depend on it.
__global__ void __launch_bounds__(256) size_hist_kernel(const int64_t* __restrict__ off, int64_t lo, int64_t n, int32_t* __restrict__ hist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = off[lo + i + 1] - off[lo + i];
    atomicAdd(hist + (1023 - (int)(c < 1023 ? c : 1023)), 1);
}
__global__ void __launch_bounds__(1024) size_scan_kernel(int32_t* __restrict__ hist) {
    __shared__ int32_t s[1024];
    const int t = threadIdx.x;
    const int32_t v = hist[t];
    s[t] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int32_t u = (t >= d) ? s[t - d] : 0;
        __syncthreads();
        s[t] += u;
        __syncthreads();
    }
    hist[t] = s[t] - v;  // exclusive: the bin's cursor
}
__global__ void __launch_bounds__(256) size_scatter_kernel(const int64_t* __restrict__ off, int64_t lo, int64_t n, int32_t* __restrict__ cursor,
                                                           int32_t* __restrict__ ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = off[lo + i + 1] - off[lo + i];
    const int pos = atomicAdd(cursor + (1023 - (int)(c < 1023 ? c : 1023)), 1);
    ids[pos] = (int32_t)(lo + i);
}
void launch_size_order(const int64_t* off, int64_t lo, int64_t n, int32_t* hist1024, int32_t* ids, cudaStream_t s) {
    if (n <= 0) return;
    cudaMemsetAsync(hist1024, 0, 1024 * sizeof(int32_t), s);
    size_hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(off, lo, n, hist1024);
    size_scan_kernel<<<1, 1024, 0, s>>>(hist1024);
    size_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(off, lo, n, hist1024, ids);
    g_launches += 3;
}

// nbv (int32) -> int64 for the scan
__global__ void widen_i32_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
// strided per-patch parameters (stride = capacity) -> packed arrays at bv_off[p]; one warp per patch
__global__ void __launch_bounds__(256) compact_params_kernel(const int32_t* __restrict__ nbv, const int64_t* __restrict__ bv_off,
                                                             int64_t n_patches, int stride, const double* __restrict__ alpha,
                                                             const double* __restrict__ b1, const double* __restrict__ b2,
                                                             const int32_t* __restrict__ idx, double* __restrict__ palpha,
                                                             double* __restrict__ pb1, double* __restrict__ pb2,
                                                             int32_t* __restrict__ pidx) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_patches) return;
    const int n = nbv[p];
    const int64_t src = p * stride, dst = bv_off[p];
    for (int i = lane; i < n; i += 32) {
        palpha[dst + i] = alpha[src + i];
        pb1[dst + i] = b1[src + i];
        pb2[dst + i] = b2[src + i];
        pidx[dst + i] = idx[src + i];
    }
}
void launch_compact_params(const int32_t* nbv, int64_t n_patches, int stride, const double* alpha, const double* b1,
                           const double* b2, const int32_t* idx, int64_t* widened, int64_t* bv_off, void* scan_tmp,
                           double* palpha, double* pb1, double* pb2, int32_t* pidx, cudaStream_t s) {
    if (n_patches <= 0) {
        cudaMemsetAsync(bv_off, 0, sizeof(int64_t), s);
        return;
    }
    widen_i32_kernel<<<(unsigned)((n_patches + 255) / 256), 256, 0, s>>>(nbv, n_patches, widened);
    launch_exclusive_scan_i64(widened, bv_off, n_patches, scan_tmp, s);
    compact_params_kernel<<<(unsigned)((n_patches * 32 + 255) / 256), 256, 0, s>>>(nbv, bv_off, n_patches, stride, alpha, b1, b2, idx,
                                                                                 palpha, pb1, pb2, pidx);
    g_launches += 2;
}

__global__ void debug_exp_kernel(const double* __restrict__ x, double* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gpc_exp(x[i]);
}
void launch_debug_exp(const double* x, double* out, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    debug_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, out, n);
}

}  // namespace gpc
