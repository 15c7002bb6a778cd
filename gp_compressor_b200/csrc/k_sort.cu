// k_sort.cu — K2: stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass.
//
// Replaces the ordering the reference gets from PCL's octree: insertion into voxels,
// depth-first leaf iteration (gp_compressor.cpp:204-205) and the search order of
// radiusSearch (:220).  Sorting (Morton code, point index) ascending gives both; a second
// sort on the owning patch index groups claimed points per patch while keeping that order.
//
// Per pass: (1) per-tile digit histograms, bin-major; (2) exclusive scan (k_util.cu);
// (3) scatter with a stable in-tile rank: the tile is walked in rounds of 256 elements,
// each warp ranks equal digits with match_any, warp counts are prefix-summed per digit.
#include "gpc_internal.h"

namespace gpc {

namespace {

constexpr int RS_T = 256;            // threads per block
constexpr int RS_ROUNDS = 16;        // elements per thread
constexpr int RS_TILE = RS_T * RS_ROUNDS;
constexpr int RS_WARPS = RS_T / 32;

__global__ void __launch_bounds__(RS_T) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                          int64_t n_tiles, int64_t* __restrict__ hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; r++) {
        int64_t i = base + r * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_T) radix_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                             int64_t n, int shift, int64_t n_tiles,
                                                             const int64_t* __restrict__ offs, uint64_t* __restrict__ okeys,
                                                             uint32_t* __restrict__ ovals) {
    __shared__ int64_t gbase[256];            // global offset of this tile's run of each digit
    __shared__ unsigned int run[256];         // elements of each digit already placed by earlier rounds
    __shared__ unsigned int wcnt[RS_WARPS][256];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    gbase[t] = offs[(int64_t)t * n_tiles + blockIdx.x];
    run[t] = 0;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; r++) {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) wcnt[ww][t] = 0;
        __syncthreads();
        const int64_t i = base + r * RS_T + t;
        const bool valid = i < n;
        uint64_t k = 0;
        uint32_t v = 0;
        unsigned d = 256;  // invalid lanes match nobody's digit
        if (valid) { k = keys[i]; v = vals[i]; d = (unsigned)(k >> shift) & 255u; }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned below = peers & ((1u << lane) - 1u);
        if (valid && below == 0) wcnt[w][d] = __popc(peers);
        __syncthreads();
        // exclusive prefix over warps for digit t, then advance the running count
        {
            unsigned acc = run[t];
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) {
                unsigned c = wcnt[ww][t];
                wcnt[ww][t] = acc;
                acc += c;
            }
            run[t] = acc;
        }
        __syncthreads();
        if (valid) {
            int64_t dst = gbase[d] + wcnt[w][d] + __popc(below);
            okeys[dst] = k;
            ovals[dst] = v;
        }
        __syncthreads();
    }
}

__global__ void iota_u32_kernel(uint32_t* v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

}  // namespace

size_t radix_sort_tmp_bytes(int64_t n) {
    int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    int64_t m = 256 * tiles;
    return (size_t)(m + 1) * sizeof(int64_t) * 2 + scan_tmp_bytes(m) + 64;
}

// Sorts on key bits [0, nbits).  Buffers ping-pong; returns which buffer (0: keys/vals, 1: keys2/vals2)
// holds the sorted result.
int launch_radix_sort(uint64_t* keys, uint32_t* vals, uint64_t* keys2, uint32_t* vals2, int64_t n, int nbits, void* tmp,
                      cudaStream_t s) {
    if (n <= 0 || nbits <= 0) return 0;
    const int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    const int64_t m = 256 * tiles;
    int64_t* hist = reinterpret_cast<int64_t*>(tmp);
    int64_t* offs = hist + (m + 1);
    void* scan_tmp = offs + (m + 1);
    int cur = 0;
    for (int shift = 0; shift < nbits; shift += 8) {
        const uint64_t* ki = cur ? keys2 : keys;
        const uint32_t* vi = cur ? vals2 : vals;
        uint64_t* ko = cur ? keys : keys2;
        uint32_t* vo = cur ? vals : vals2;
        radix_hist_kernel<<<(unsigned)tiles, RS_T, 0, s>>>(ki, n, shift, tiles, hist);
        launch_exclusive_scan_i64(hist, offs, m, scan_tmp, s);
        radix_scatter_kernel<<<(unsigned)tiles, RS_T, 0, s>>>(ki, vi, n, shift, tiles, offs, ko, vo);
        g_launches += 2;
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
