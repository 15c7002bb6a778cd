// k_rand.cu — K6a: the glibc rand() stream on the device, with jump-ahead.
//
// Replaces: the process-global unseeded rand() consumed by sparse_gp::shuffle
// (/root/reference/src/sparse_gp.hpp:51) and by sparse_gp_field::shuffle
// (sparse_gp_field.hpp:38).  glibc's default generator (TYPE_3) is the additive lagged
// Fibonacci recurrence r[i] = r[i-31] + r[i-3] (mod 2^32) with output r[i] >> 1, valid
// from i = 34; rand() call number k (k = 0,1,...) returns r[344 + k] >> 1.
// With s[j] = r[3 + j] the recurrence s[j] = s[j-31] + s[j-3] holds for every j >= 31 and
// output k is s[341 + k] >> 1.
//
// Jump-ahead: in Z_{2^32}[x] / (x^31 - x^28 - 1), x^n = sum_k c_k x^k  implies
// s[n + j] = sum_k c_k s[k + j].  One WARP produces a chunk of 32 x 512 consecutive draws:
//   1. lanes = polynomial coefficients: c = product of the tabulated x^(2^k) selected by the bits of the
//      chunk's start (warp-parallel convolution; the fold of degrees 31..60 is a tabulated 30 x 31 matrix);
//   2. the 61-word window at the chunk start from s[0..90];
//   3. lane l jumps a further 512 l draws with the tabulated x^(512 l) applied to that window;
//   4. every lane runs the recurrence for its 512 draws (31-word ring in registers), staging 32 draws per
//      lane in shared memory so that the stores to the stream are coalesced.
#include "gpc_internal.h"

namespace gpc {

constexpr int RW_PER_LANE = 512;
constexpr int RW_CHUNK = 32 * RW_PER_LANE;

__constant__ uint32_t c_pow2[48][31];   // x^(2^k) mod P
__device__ uint32_t g_base[91];         // s[0..90]
__device__ uint32_t g_red[30][31];      // x^(31+k) mod P, k = 0..29
__device__ uint32_t g_lanepow[32][31];  // x^(512 l) mod P, l = 0..31

static void polymulmod_host(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    uint32_t t[61];
    for (int i = 0; i < 61; i++) t[i] = 0;
    for (int i = 0; i < 31; i++) {
        uint32_t ai = a[i];
        if (ai == 0) continue;
        for (int j = 0; j < 31; j++) t[i + j] += ai * b[j];
    }
    for (int d = 60; d >= 31; --d) {  // x^d = x^(d-3) + x^(d-31)
        t[d - 3] += t[d];
        t[d - 31] += t[d];
    }
    for (int i = 0; i < 31; i++) out[i] = t[i];
}

void rand_tables_init(RandTables* T) {
    uint32_t r[100];
    r[0] = 1;
    for (int i = 1; i < 31; i++) {
        int64_t v = (16807LL * (int32_t)r[i - 1]) % 2147483647;
        if (v < 0) v += 2147483647;
        r[i] = (uint32_t)v;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 94; i++) r[i] = r[i - 31] + r[i - 3];
    for (int j = 0; j < 91; j++) T->base[j] = r[3 + j];
    for (int i = 0; i < 31; i++) T->pow2[0][i] = (i == 1) ? 1u : 0u;  // x
    for (int k = 1; k < 48; k++) polymulmod_host(T->pow2[k - 1], T->pow2[k - 1], T->pow2[k]);
    // x^(31+k) mod P: start from x^31 = x^28 + 1 and multiply by x
    uint32_t cur[31];
    for (int i = 0; i < 31; i++) cur[i] = (i == 28 || i == 0) ? 1u : 0u;
    for (int k = 0; k < 30; k++) {
        for (int i = 0; i < 31; i++) T->red[k][i] = cur[i];
        uint32_t nx[31];
        polymulmod_host(cur, T->pow2[0], nx);
        for (int i = 0; i < 31; i++) cur[i] = nx[i];
    }
    // x^(512 l): x^512 = pow2[9]
    for (int i = 0; i < 31; i++) T->lanepow[0][i] = (i == 0) ? 1u : 0u;
    for (int l = 1; l < 32; l++) polymulmod_host(T->lanepow[l - 1], T->pow2[9], T->lanepow[l]);
}

cudaError_t rand_upload_tables(const RandTables* T) {
    cudaError_t e = cudaMemcpyToSymbol(c_pow2, T->pow2, sizeof(T->pow2));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(g_base, T->base, sizeof(T->base));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(g_red, T->red, sizeof(T->red));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_lanepow, T->lanepow, sizeof(T->lanepow));
}

namespace {

struct RandSmem {
    uint32_t pa[32], pb[32], th[32];
    uint32_t W[64];
    uint32_t base[96];
    uint32_t red[30][31];
    uint32_t tile[32][33];
};

// pa <- pa * pb mod P.  Lane i < 31 holds coefficient i.
__device__ __forceinline__ void polymul_warp(RandSmem& sm, int lane) {
    uint32_t lo = 0, hi = 0;
    if (lane < 31) {
        for (int j = 0; j <= lane; j++) lo += sm.pa[j] * sm.pb[lane - j];            // degree lane
        for (int j = lane + 1; j < 31; j++) hi += sm.pa[j] * sm.pb[31 + lane - j];   // degree 31 + lane
    }
    __syncwarp();
    sm.th[lane] = (lane < 30) ? hi : 0u;
    __syncwarp();
    if (lane < 31) {
        uint32_t acc = lo;
        for (int k = 0; k < 30; k++) acc += sm.th[k] * sm.red[k][lane];
        sm.pa[lane] = acc;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) rand_stream_kernel(uint64_t offset, int64_t n, uint32_t* __restrict__ out) {
    __shared__ RandSmem sm;
    const int lane = threadIdx.x;
    const int64_t first = (int64_t)blockIdx.x * RW_CHUNK;
    if (first >= n) return;
    for (int i = lane; i < 30 * 31; i += 32) (&sm.red[0][0])[i] = (&g_red[0][0])[i];
    for (int i = lane; i < 91; i += 32) sm.base[i] = g_base[i];
    // the first output of the chunk is s[n0], n0 = 341 + offset + first; the window starts 31 earlier
    const uint64_t m = 341ull + offset + (uint64_t)first - 31ull;
    sm.pa[lane] = (lane == 0) ? 1u : 0u;
    __syncwarp();
    for (int k = 0; k < 48; k++)
        if ((m >> k) & 1ull) {
            sm.pb[lane] = (lane < 31) ? c_pow2[k][lane] : 0u;
            __syncwarp();
            polymul_warp(sm, lane);
        }
    // window at the chunk start: W[j] = s[m + j] = sum_k c_k s[k + j], j = 0..60
    {
        uint32_t a0 = 0, a1 = 0;
        for (int k = 0; k < 31; k++) {
            const uint32_t ck = sm.pa[k];
            a0 += ck * sm.base[k + lane];
            if (lane + 32 < 61) a1 += ck * sm.base[k + lane + 32];
        }
        sm.W[lane] = a0;
        sm.W[lane + 32] = a1;
    }
    __syncwarp();
    // lane l: window at m + 512 l :  w[j] = sum_k T_l[k] * W[k + j]
    uint32_t w[31];
#pragma unroll
    for (int j = 0; j < 31; j++) w[j] = 0;
    for (int k = 0; k < 31; k++) {
        const uint32_t tk = g_lanepow[lane][k];
#pragma unroll
        for (int j = 0; j < 31; j++) w[j] += tk * sm.W[k + j];
    }
    __syncwarp();
    // 512 draws per lane = 16 rounds of 32.  The ring w[0..30] holds the last 31 values with the oldest at
    // position 0; s[i] = s[i-31] + s[i-3] = w[p] + w[(p + 28) % 31] for the p-th draw of a round.
    for (int round = 0; round < RW_PER_LANE / 32; round++) {
#pragma unroll
        for (int p = 0; p < 31; p++) {
            const uint32_t v = w[p] + w[(p + 28) % 31];
            w[p] = v;
            sm.tile[lane][p] = v >> 1;
        }
        {
            const uint32_t v = w[0] + w[28];  // the 32nd draw wraps to position 0
            w[0] = v;
            sm.tile[lane][31] = v >> 1;
        }
        {   // rotate left by one: the oldest value is at position 0 again
            const uint32_t t0 = w[0];
#pragma unroll
            for (int j = 0; j < 30; j++) w[j] = w[j + 1];
            w[30] = t0;
        }
        __syncwarp();
        // row r of the tile holds lane r's 32 draws of this round: coalesced 128-byte stores
#pragma unroll 4
        for (int r = 0; r < 32; r++) {
            const int64_t idx = first + (int64_t)r * RW_PER_LANE + round * 32 + lane;
            if (idx < n) out[idx] = sm.tile[r][lane];
        }
        __syncwarp();
    }
}

}  // namespace

void launch_rand_stream(uint64_t offset, int64_t n, uint32_t* out, cudaStream_t s) {
    if (n <= 0) return;
    const int64_t chunks = (n + RW_CHUNK - 1) / RW_CHUNK;
    rand_stream_kernel<<<(unsigned)chunks, 32, 0, s>>>(offset, n, out);
    g_launches++;
}

}  // namespace gpc
