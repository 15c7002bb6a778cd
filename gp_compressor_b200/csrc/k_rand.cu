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
// s[n + j] = sum_k c_k s[k + j].  x^(2^k) are tabulated once on the host; a thread
// multiplies the table entries selected by the bits of n, rebuilds its 31-word window
// from s[0..60], then runs the recurrence for its block of outputs.
#include "gpc_internal.h"

namespace gpc {

__constant__ uint32_t c_pow2[48][31];
__constant__ uint32_t c_base[61];

template <class T>
__host__ __device__ inline void polymulmod(const T* a, const T* b, uint32_t* out) {
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
    uint32_t r[64 + 3];
    r[0] = 1;
    for (int i = 1; i < 31; i++) {
        int64_t v = (16807LL * (int32_t)r[i - 1]) % 2147483647;
        if (v < 0) v += 2147483647;
        r[i] = (uint32_t)v;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 64; i++) r[i] = r[i - 31] + r[i - 3];
    for (int j = 0; j < 61; j++) T->base[j] = r[3 + j];
    for (int i = 0; i < 31; i++) T->pow2[0][i] = (i == 1) ? 1u : 0u;  // x
    for (int k = 1; k < 48; k++) polymulmod(T->pow2[k - 1], T->pow2[k - 1], T->pow2[k]);
}

cudaError_t rand_upload_tables(const RandTables* T) {
    cudaError_t e = cudaMemcpyToSymbol(c_pow2, T->pow2, sizeof(T->pow2));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_base, T->base, sizeof(T->base));
}

constexpr int RAND_PER_THREAD = 2048;

__global__ void __launch_bounds__(32) rand_stream_kernel(uint64_t offset, int64_t n, uint32_t* __restrict__ out) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t first = b * RAND_PER_THREAD;
    if (first >= n) return;
    // first output is s[n0], n0 = 341 + offset + first; window w[j] = s[m + j], m = n0 - 31
    uint64_t m = 341ull + offset + (uint64_t)first - 31ull;
    uint32_t c[31];
    for (int i = 0; i < 31; i++) c[i] = (i == 0) ? 1u : 0u;
    for (int k = 0; k < 48; k++)
        if ((m >> k) & 1ull) {
            uint32_t o[31];
            polymulmod(c, c_pow2[k], o);
            for (int i = 0; i < 31; i++) c[i] = o[i];
        }
    uint32_t w[31];
    for (int j = 0; j < 31; j++) {
        uint32_t acc = 0;
        for (int k = 0; k < 31; k++) acc += c[k] * c_base[k + j];
        w[j] = acc;
    }
    int64_t cnt = n - first;
    if (cnt > RAND_PER_THREAD) cnt = RAND_PER_THREAD;
    int p = 0;
    for (int64_t i = 0; i < cnt; i++) {
        int q = p + 28;
        if (q >= 31) q -= 31;
        uint32_t v = w[p] + w[q];
        w[p] = v;
        out[first + i] = v >> 1;
        if (++p == 31) p = 0;
    }
}

void launch_rand_stream(uint64_t offset, int64_t n, uint32_t* out, cudaStream_t s) {
    if (n <= 0) return;
    int64_t threads = (n + RAND_PER_THREAD - 1) / RAND_PER_THREAD;
    int blocks = (int)((threads + 31) / 32);  // small blocks: the few thousand threads spread over all SMs
    rand_stream_kernel<<<blocks, 32, 0, s>>>(offset, n, out);
    g_launches++;
}

}  // namespace gpc
