// k_sort.cu — K2: stable LSD radix sort of (64-bit key, 32-bit value) pairs.
//
// Replaces the ordering the reference gets from PCL's octree: insertion into voxels,
// depth-first leaf iteration (gp_compressor.cpp:204-205) and the search order of
// radiusSearch (:220).  Sorting (Morton code, point index) ascending gives both; a second
// sort on the owning patch index groups claimed points per patch while keeping that order.
//
// HBM-bound integer work, so the design minimises bytes per element and pass:
//  * digits of up to 9 bits: the 3*depth+1 key bits of a room-sized cloud (25) take 3 passes, not 4;
//  * keys that fit 32 bits are NARROWED by the first pass and widened again by the last one, so the
//    passes in between move 8 bytes per element instead of 12 (the interface stays 64-bit);
//  * ONE histogram kernel up front counts the digits of every pass (the keys are read once); each pass is then a
//    single scatter kernel: a tile's offset inside every digit comes from a chained scan with decoupled look-back
//    over the tiles before it (tile numbers are handed out by an atomic counter, so predecessors are running).
//    The scatter ranks every element inside its tile with warp match + per-warp digit counters (warps own contiguous
//    runs, so the rank is stable) and ONE block barrier, reorders the tile by digit in shared memory and writes it
//    out run by run: consecutive threads write consecutive addresses.
//    (Inputs of 2^30 elements or more take the older three-kernel pass: per-tile histogram, scan, scatter.)
#include <algorithm>
#include <cstdlib>

#include "gpc_internal.h"

namespace gpc {

namespace {

constexpr int RS_T = 256;              // threads per block
constexpr int RS_ITEMS = 8;            // elements per thread
constexpr int RS_TILE = RS_T * RS_ITEMS;
constexpr int RS_WARPS = RS_T / 32;
constexpr int RS_MAXBITS = 9;
constexpr int RS_MAXBINS = 1 << RS_MAXBITS;

template <class K>
__global__ void __launch_bounds__(RS_T) radix_hist_kernel(const K* __restrict__ keys, int64_t n, int shift, int bits, int64_t n_tiles,
                                                          uint32_t* __restrict__ hist) {
    __shared__ unsigned int h[RS_MAXBINS];
    const int bins = 1 << bits;
    for (int i = threadIdx.x; i < bins; i += RS_T) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    const unsigned mask = (unsigned)bins - 1u;
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; r++) {
        const int64_t i = base + r * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += RS_T) hist[(int64_t)i * n_tiles + blockIdx.x] = h[i];
}

// digit counts of EVERY pass in one read of the keys: hist_all[p * 512 + digit]
constexpr int RS_MAXPASSES = 16;
struct PassPlan { int passes; int shift[RS_MAXPASSES]; int bits[RS_MAXPASSES]; };

template <class K>
__global__ void __launch_bounds__(RS_T) radix_hist_all_kernel(const K* __restrict__ keys, int64_t n, PassPlan pl, uint32_t* __restrict__ hist_all) {
    extern __shared__ unsigned int hall[];   // passes * 512
    for (int i = threadIdx.x; i < pl.passes * RS_MAXBINS; i += RS_T) hall[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * RS_T + threadIdx.x; i < n; i += (int64_t)gridDim.x * RS_T) {
        const K k = keys[i];
        for (int p = 0; p < pl.passes; p++) atomicAdd(&hall[p * RS_MAXBINS + ((unsigned)(k >> pl.shift[p]) & ((1u << pl.bits[p]) - 1u))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pl.passes * RS_MAXBINS; i += RS_T)
        if (hall[i]) atomicAdd(&hist_all[i], hall[i]);
}
// digit_base[p * 512 + d] = number of keys whose digit of pass p is below d (one block per pass)
__global__ void __launch_bounds__(RS_MAXBINS) radix_digit_base_kernel(const uint32_t* __restrict__ hist_all, uint32_t* __restrict__ digit_base) {
    __shared__ uint32_t ws[RS_MAXBINS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t v = hist_all[blockIdx.x * RS_MAXBINS + t];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    uint32_t b = 0;
    for (int ww = 0; ww < w; ww++) b += ws[ww];
    digit_base[blockIdx.x * RS_MAXBINS + t] = b + x - v;
}

constexpr uint32_t ST_AGG = 1u << 30, ST_INCL = 2u << 30, ST_MASK = (1u << 30) - 1u;

// exclusive scan of m uint32 counts (bin-major tile histograms) in three small kernels
constexpr int SC_T = 512, SC_ITEMS = 8, SC_TILE = SC_T * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[SC_T / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SC_T / 32 ? wsum[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SC_T / 32) wsum[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = w ? wsum[w - 1] : 0u;
    if (total) *total = wsum[SC_T / 32 - 1];
    __syncthreads();
    return before + x - v;
}

__global__ void __launch_bounds__(SC_T) scan_u32_sums_kernel(const uint32_t* __restrict__ in, int64_t m, uint32_t* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++)
        if (base + i < m) s += in[base + i];
    uint32_t tot;
    block_exclusive_scan_u32(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SC_T) scan_u32_top_kernel(uint32_t* __restrict__ sums, int64_t nb) {
    // nb <= SC_T * SC_ITEMS block sums, one block
    uint32_t v[SC_ITEMS];
    uint32_t s = 0;
    const int64_t base = (int64_t)threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { v[i] = base + i < nb ? sums[base + i] : 0u; s += v[i]; }
    uint32_t ex = block_exclusive_scan_u32(s, nullptr);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        if (base + i < nb) sums[base + i] = ex;
        ex += v[i];
    }
}
__global__ void __launch_bounds__(SC_T) scan_u32_apply_kernel(const uint32_t* __restrict__ in, int64_t m, const uint32_t* __restrict__ sums,
                                                              uint32_t* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) { v[i] = base + i < m ? in[base + i] : 0u; s += v[i]; }
    uint32_t ex = block_exclusive_scan_u32(s, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        if (base + i < m) out[base + i] = ex;
        ex += v[i];
    }
}

template <class KS>
struct ScatterSmem {
    KS key[RS_TILE];                     // the tile reordered by digit (already narrowed when the pass writes 32-bit keys)
    uint32_t val[RS_TILE];
    unsigned short wcnt[RS_WARPS][RS_MAXBINS]; // per-warp digit counts (<= 512), then exclusive bases over the warps
    unsigned short toff[RS_MAXBINS];     // start of every digit inside the reordered tile
    uint32_t gbase[RS_MAXBINS];          // global offset of the tile's run of every digit
    uint32_t wtot[RS_T / 32];
};

// KIN / KOUT: uint64_t or uint32_t (keys that fit 32 bits travel narrow between the first and the last pass)
// LOOKBACK: offs = digit_base of this pass (512 entries), status = n_tiles x 512 words (zeroed), counter = tile dispenser;
// otherwise offs = the scanned bin-major (digit, tile) table of the three-kernel pass.
template <class KIN, class KOUT, bool LOOKBACK>
__global__ void __launch_bounds__(RS_T, 4) radix_scatter_kernel(const KIN* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
                                                             int shift, int bits, int64_t n_tiles, const uint32_t* __restrict__ offs,
                                                             KOUT* __restrict__ okeys, uint32_t* __restrict__ ovals,
                                                             uint32_t* __restrict__ status, unsigned int* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    ScatterSmem<KOUT>& sm = *reinterpret_cast<ScatterSmem<KOUT>*>(rs_smem);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int bins = 1 << bits;
    const unsigned mask = (unsigned)bins - 1u;
    if (LOOKBACK && t == 0) sm.wtot[0] = atomicAdd(counter, 1u);   // tiles start in the order of their numbers
    for (int i = t; i < RS_WARPS * RS_MAXBINS / 2; i += RS_T) reinterpret_cast<uint32_t*>(&sm.wcnt[0][0])[i] = 0u;
    __syncthreads();
    const int64_t tile = LOOKBACK ? (int64_t)sm.wtot[0] : (int64_t)blockIdx.x;
    if (!LOOKBACK)
        for (int i = t; i < bins; i += RS_T) sm.gbase[i] = offs[(int64_t)i * n_tiles + tile];
    // warp w owns the elements [base + w * 32 * RS_ITEMS, + 32 * RS_ITEMS) in RS_ITEMS rounds of 32: ranks grow in input order (stable)
    const int64_t wbase = tile * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    KIN kreg[RS_ITEMS];
    uint32_t vreg[RS_ITEMS];
    unsigned short rank[RS_ITEMS];
    // all loads of the thread first (independent, in flight together), then the ranking rounds
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        kreg[r] = i < n ? keys[i] : (KIN)0;
        vreg[r] = i < n ? vals[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned d = valid ? ((unsigned)(kreg[r] >> shift) & mask) : (unsigned)bins;  // invalid lanes match only each other
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned below = peers & ((1u << lane) - 1u);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (valid && lane == leader) {
            old = sm.wcnt[w][d];
            sm.wcnt[w][d] = (unsigned short)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = (unsigned short)(old + __popc(below));
        __syncwarp();   // the next round's leaders read the counters this round's leaders wrote
    }
    __syncthreads();
    // per digit: exclusive bases over the warps and the tile total; then an exclusive scan of the totals over the digits
    uint32_t tot2[2] = {0u, 0u};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int d = t + q * RS_T;
        if (d < bins) {
            uint32_t acc = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) {
                const uint32_t c = sm.wcnt[ww][d];
                sm.wcnt[ww][d] = (unsigned short)acc;
                acc += c;
            }
            tot2[q] = acc;
            // publish the tile's count of this digit at once; the prefix over the earlier tiles is looked up below
            if (LOOKBACK) __stcg(&status[tile * RS_MAXBINS + d], (tile == 0 ? ST_INCL : ST_AGG) | acc);
        }
    }
    {
        // digits t (first half) and t + 256 (second half): scan each half across the threads, chain the halves
        uint32_t x0 = tot2[0], x1 = tot2[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
            if (lane >= o) { x0 += y0; x1 += y1; }
        }
        __shared__ uint32_t ws0[RS_WARPS], ws1[RS_WARPS];
        if (lane == 31) { ws0[w] = x0; ws1[w] = x1; }
        __syncthreads();
        uint32_t b0 = 0, b1 = 0, all0 = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            if (ww < w) { b0 += ws0[ww]; b1 += ws1[ww]; }
            all0 += ws0[ww];
        }
        if (t < bins) sm.toff[t] = (unsigned short)(b0 + x0 - tot2[0]);
        if (t + RS_T < bins) sm.toff[t + RS_T] = (unsigned short)(all0 + b1 + x1 - tot2[1]);
    }
    __syncthreads();
    // reorder the tile by digit in shared memory
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)(kreg[r] >> shift) & mask;
            const uint32_t pos = (uint32_t)sm.toff[d] + sm.wcnt[w][d] + rank[r];
            sm.key[pos] = (KOUT)kreg[r];
            sm.val[pos] = vreg[r];
        }
    }
    if (LOOKBACK) {
        // decoupled look-back: sum the counts of the tiles before this one, stopping at the first inclusive prefix
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int d = t + q * RS_T;
            if (d < bins) {
                uint32_t excl = 0;
                for (int64_t j = tile - 1; j >= 0; j--) {
                    uint32_t sw;
                    do { sw = *reinterpret_cast<volatile const uint32_t*>(&status[j * RS_MAXBINS + d]); } while ((sw >> 30) == 0u);  // re-read until published
                    excl += sw & ST_MASK;
                    if (sw & ST_INCL) break;
                }
                if (tile > 0) __stcg(&status[tile * RS_MAXBINS + d], ST_INCL | (excl + tot2[q]));
                sm.gbase[d] = offs[d] + excl;
            }
        }
    }
    __syncthreads();
    const int64_t tile_n = min((int64_t)RS_TILE, n - tile * RS_TILE);
    for (int i = t; i < tile_n; i += RS_T) {
        const KOUT k = sm.key[i];
        const unsigned d = (unsigned)(k >> shift) & mask;
        const int64_t dst = (int64_t)sm.gbase[d] + (i - (int)sm.toff[d]);
        okeys[dst] = k;
        ovals[dst] = sm.val[i];
    }
}

__global__ void iota_u32_kernel(uint32_t* v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

template <class KIN, class KOUT, bool LOOKBACK>
cudaError_t launch_scatter(const void* ki, const uint32_t* vi, int64_t n, int shift, int bits, int64_t tiles, const uint32_t* offs, void* ko,
                           uint32_t* vo, uint32_t* status, unsigned int* counter, cudaStream_t s) {
    {
        cudaError_t e = GPC_FUNC_ATTR_ONCE((radix_scatter_kernel<KIN, KOUT, LOOKBACK>), cudaFuncAttributeMaxDynamicSharedMemorySize, sizeof(ScatterSmem<KOUT>));
        if (e != cudaSuccess) return e;
    }
    radix_scatter_kernel<KIN, KOUT, LOOKBACK><<<(unsigned)tiles, RS_T, sizeof(ScatterSmem<KOUT>), s>>>(reinterpret_cast<const KIN*>(ki), vi, n, shift, bits, tiles,
                                                                                                 offs, reinterpret_cast<KOUT*>(ko), vo, status, counter);
    return cudaGetLastError();
}
template <bool LOOKBACK>
cudaError_t launch_scatter_typed(bool in32, bool out32, const void* ki, const uint32_t* vi, int64_t n, int shift, int bits, int64_t tiles,
                                 const uint32_t* offs, void* ko, uint32_t* vo, uint32_t* status, unsigned int* counter, cudaStream_t s) {
    if (in32 && out32) return launch_scatter<uint32_t, uint32_t, LOOKBACK>(ki, vi, n, shift, bits, tiles, offs, ko, vo, status, counter, s);
    if (in32) return launch_scatter<uint32_t, uint64_t, LOOKBACK>(ki, vi, n, shift, bits, tiles, offs, ko, vo, status, counter, s);
    if (out32) return launch_scatter<uint64_t, uint32_t, LOOKBACK>(ki, vi, n, shift, bits, tiles, offs, ko, vo, status, counter, s);
    return launch_scatter<uint64_t, uint64_t, LOOKBACK>(ki, vi, n, shift, bits, tiles, offs, ko, vo, status, counter, s);
}

}  // namespace

size_t radix_sort_tmp_bytes(int64_t n) {
    const int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const int64_t m = (int64_t)RS_MAXBINS * tiles;
    const int64_t nb = (m + SC_TILE - 1) / SC_TILE;
    // the larger of: three-kernel pass (2 m + nb words) and look-back pass (2 * 16 * 512 + 64 + m words)
    return (size_t)(2 * m + nb + 2 * RS_MAXPASSES * RS_MAXBINS + 128) * sizeof(uint32_t);
}

// Sorts on key bits [0, nbits).  Buffers ping-pong; returns which buffer (0: keys/vals, 1: keys2/vals2)
// holds the sorted result (64-bit keys again).  n < 2^32.
int launch_radix_sort(uint64_t* keys, uint32_t* vals, uint64_t* keys2, uint32_t* vals2, int64_t n, int nbits, void* tmp,
                      cudaStream_t s) {
    if (n <= 0 || nbits <= 0) return 0;
    const int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const bool lookback = n < ((int64_t)1 << 30);   // status words carry 30-bit counts
    // digit width: up to 9 bits, fewer for huge inputs so that the (digit, tile) count table scans in one top-level block
    int maxbits = RS_MAXBITS;
    while (!lookback && maxbits > 4 && (((int64_t)1 << maxbits) * tiles + SC_TILE - 1) / SC_TILE > SC_TILE) maxbits--;
    const int passes = (nbits + maxbits - 1) / maxbits;
    const bool narrow = nbits <= 32 && passes > 1;   // 32-bit keys between the first and the last pass
    PassPlan pl;
    pl.passes = passes;
    for (int p = 0, sh = 0; p < passes; p++) {
        pl.bits[p] = (nbits - sh + (passes - p) - 1) / (passes - p);   // spread the bits evenly over the passes
        pl.shift[p] = sh;
        sh += pl.bits[p];
    }
    uint32_t* w = reinterpret_cast<uint32_t*>(tmp);
    int cur = 0;
    if (lookback) {
        uint32_t* hist_all = w;                                   // passes * 512
        uint32_t* digit_base = w + RS_MAXPASSES * RS_MAXBINS;     // passes * 512
        unsigned int* counters = digit_base + RS_MAXPASSES * RS_MAXBINS;   // one tile dispenser per pass
        uint32_t* status = counters + 64;                         // tiles * 512, zeroed before every pass
        cudaMemsetAsync(w, 0, (size_t)(2 * RS_MAXPASSES * RS_MAXBINS + 64) * sizeof(uint32_t), s);
        const unsigned hb = (unsigned)std::min<int64_t>((n + RS_T - 1) / RS_T, 148 * 8);
        radix_hist_all_kernel<uint64_t><<<hb, RS_T, (size_t)passes * RS_MAXBINS * sizeof(unsigned int), s>>>(keys, n, pl, hist_all);
        radix_digit_base_kernel<<<passes, RS_MAXBINS, 0, s>>>(hist_all, digit_base);
        g_launches += 2;
        for (int p = 0; p < passes; p++) {
            void* ki = cur ? (void*)keys2 : (void*)keys;
            const uint32_t* vi = cur ? vals2 : vals;
            void* ko = cur ? (void*)keys : (void*)keys2;
            uint32_t* vo = cur ? vals : vals2;
            const bool in32 = narrow && p > 0, out32 = narrow && p + 1 < passes;
            cudaMemsetAsync(status, 0, (size_t)tiles * RS_MAXBINS * sizeof(uint32_t), s);
            launch_scatter_typed<true>(in32, out32, ki, vi, n, pl.shift[p], pl.bits[p], tiles, digit_base + p * RS_MAXBINS, ko, vo, status,
                                       counters + p, s);
            g_launches++;
            cur ^= 1;
        }
        return cur;
    }
    uint32_t* hist = w;
    const int64_t mmax = (int64_t)RS_MAXBINS * tiles;
    uint32_t* offs = hist + mmax;
    uint32_t* sums = offs + mmax;
    for (int p = 0; p < passes; p++) {
        const int bits = pl.bits[p], shift = pl.shift[p];
        void* ki = cur ? (void*)keys2 : (void*)keys;
        const uint32_t* vi = cur ? vals2 : vals;
        void* ko = cur ? (void*)keys : (void*)keys2;
        uint32_t* vo = cur ? vals : vals2;
        const bool in32 = narrow && p > 0, out32 = narrow && p + 1 < passes;
        const int64_t m = ((int64_t)1 << bits) * tiles;
        if (in32) radix_hist_kernel<uint32_t><<<(unsigned)tiles, RS_T, 0, s>>>(reinterpret_cast<const uint32_t*>(ki), n, shift, bits, tiles, hist);
        else radix_hist_kernel<uint64_t><<<(unsigned)tiles, RS_T, 0, s>>>(reinterpret_cast<const uint64_t*>(ki), n, shift, bits, tiles, hist);
        const int64_t nb = (m + SC_TILE - 1) / SC_TILE;
        scan_u32_sums_kernel<<<(unsigned)nb, SC_T, 0, s>>>(hist, m, sums);
        scan_u32_top_kernel<<<1, SC_T, 0, s>>>(sums, nb);
        scan_u32_apply_kernel<<<(unsigned)nb, SC_T, 0, s>>>(hist, m, sums, offs);
        launch_scatter_typed<false>(in32, out32, ki, vi, n, shift, bits, tiles, offs, ko, vo, nullptr, nullptr, s);
        g_launches += 5;
        cur ^= 1;
    }
    return cur;
}

void launch_iota_u32(uint32_t* v, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, n);
    g_launches++;
}

}  // namespace gpc
