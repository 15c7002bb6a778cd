// gpc_internal.h — handle layout and kernel launch prototypes (host side of libgpc_b200.so)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include <string>
#include <vector>

#include "../../include/gpc.h"

// cudaFuncSetAttribute is not a cheap call: it serialises against the work other host threads have in flight on the device
// (measured: with two handles in flight a per-launch call cost 30 % of the pipelined throughput).  This sets an integer
// attribute of a kernel once per device (again only if a larger value is asked for) and is safe to call from several threads.
#define GPC_FUNC_ATTR_ONCE(fn, attr, value)                                                                  \
    ([&]() -> cudaError_t {                                                                                  \
        static std::atomic<int> have__[64];                                                                  \
        int dev__ = 0;                                                                                       \
        cudaGetDevice(&dev__);                                                                               \
        dev__ = (dev__ < 0 || dev__ >= 64) ? 0 : dev__;                                                      \
        const int want__ = (int)(value);                                                                     \
        if (have__[dev__].load(std::memory_order_acquire) >= want__) return cudaSuccess;                     \
        const cudaError_t e__ = cudaFuncSetAttribute(fn, attr, want__);                                      \
        if (e__ == cudaSuccess) have__[dev__].store(want__, std::memory_order_release);                      \
        return e__;                                                                                          \
    }())

namespace gpc {

extern thread_local uint64_t g_launches;  // kernels launched by the current API call


// Grow-only device buffer owned by the handle.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- K1..K5: binning ----------------------------------------------------------------------
struct LatticeDev {          // PCL octree bounding box (doubles), voxel size and depth
    double mn[3], mx[3];
    double res;
    uint32_t depth;
};
struct LatticeState {        // device-resident state of the lattice replay
    LatticeDev lat;
    unsigned long long best; // index of the point found by the current search (~0: none)
    int64_t start;           // first index the next search looks at
    int32_t defined;         // PCL bounding_box_defined_
    int32_t found;           // 1 if the last step adopted a point
};
cudaError_t launch_lattice_replay(const uint8_t* cloud, int64_t n, LatticeState* st, cudaStream_t s);  // whole replay, one cooperative launch
// sharded binning (k_binning.cu): key sample -> splitters -> halo selection -> compaction -> owned patch range
void launch_shard_sample(const uint64_t* keys, int64_t n, int64_t stride, int64_t m, uint64_t* sample, uint32_t* dummy, cudaStream_t s);
void launch_shard_splitters(const uint64_t* sorted, int64_t m, int64_t stride, int depth, int rank, int count, int leaf_order, int64_t* weight,
                            int64_t* prefix, void* scan_tmp, uint64_t* range2, cudaStream_t s);
void launch_shard_select(const uint64_t* keys, int64_t n, int depth, const uint64_t* range2, int64_t* flags, cudaStream_t s);
void launch_shard_compact(const uint8_t* cloud, const int64_t* ex, int64_t n, uint8_t* sel_cloud, int32_t* sel_idx, cudaStream_t s);
size_t shard_select_tmp_bytes(int64_t n);
void launch_shard_select_compact(const uint64_t* keys, int64_t n, int depth, const uint64_t* range2, const uint8_t* cloud, uint8_t* sel_cloud,
                                 int32_t* sel_idx, unsigned long long* n_sel_out, void* tmp, cudaStream_t s);
void launch_owned_range(const uint64_t* code, int64_t P, int leaf_order, const uint64_t* range2, int64_t* out2, cudaStream_t s);
void launch_point_keys(const uint8_t* cloud, int64_t n, const LatticeDev& lat, uint64_t* keys, uint32_t* vals, void* cpt16, cudaStream_t s);
size_t radix_sort_tmp_bytes(int64_t n);
int launch_radix_sort(uint64_t* keys, uint32_t* vals, uint64_t* keys2, uint32_t* vals2, int64_t n, int nbits, void* tmp,
                      cudaStream_t s);
void launch_iota_u32(uint32_t* v, int64_t n, cudaStream_t s);
size_t leaves_fused_tmp_bytes(int64_t n);
void launch_leaves_fused(const uint64_t* keys, const uint32_t* vals, int64_t n, uint32_t depth, const void* cpt16, int32_t* leaf_of,
                         int64_t* leaf_start, uint64_t* leaf_code, void* spt, unsigned long long* counts2, void* tmp, cudaStream_t s);
void launch_leaf_neighbours(const uint64_t* leaf_code, int64_t P, const LatticeDev& lat, int32_t* nbr, int32_t* nnbr,
                            float* center, cudaStream_t s);
void launch_leaf_rotation(const void* spt, const int64_t* leaf_start, const int32_t* nbr, const int32_t* nnbr,
                          const float* center, int64_t P, double r2, double* sums, double* Rm, int32_t* ncand, cudaStream_t s);
void launch_claim(const void* spt, const int32_t* leaf_of, const int32_t* nbr, const int32_t* nnbr, const float* center,
                  const double* Rm, const int32_t* ncand, int64_t n_valid, int64_t P, double r2, double half, int leaf_order,
                  uint64_t* okey, uint32_t* oval, double* pt0, double* pt1, double* pt2, cudaStream_t s);
void launch_patch_bounds(const uint64_t* okey, int64_t n_valid, int64_t P, int64_t* patch_off, cudaStream_t s);
void launch_group_gather(const uint64_t* okey, const uint32_t* oval, const uint32_t* sorted_idx, const void* spt,
                         const double* pt0, const double* pt1, const double* pt2, int64_t n_claimed, int32_t* st_idx, double* h,
                         double* x1, double* x2, uint32_t* rgb, int32_t* owner, cudaStream_t s);
void launch_patch_frames(const int64_t* patch_off, int64_t P, int leaf_order, const double* h, const uint32_t* rgb,
                         const uint64_t* leaf_code_a, const float* center_a, const double* Rm_a, const int32_t* ncand_a, double* y,
                         uint64_t* code, float* center, int32_t* ncand, double* Rm, double* quat, double* mean, double* rgbmean,
                         cudaStream_t s);

// ---- K6: glibc rand stream + shuffle -------------------------------------------------
struct RandTables {       // polynomials are mod (x^31 - x^28 - 1) over Z/2^32
    uint32_t pow2[48][31];    // x^(2^k)
    uint32_t base[91];        // s[0..90]
    uint32_t red[30][31];     // x^(31+k)
    uint32_t lanepow[32][31]; // x^(512 l)
};
void rand_tables_init(RandTables* t);                       // host, once per process
cudaError_t rand_upload_tables(const RandTables* t);        // to __constant__
// out[i] = rand() value number (offset + i), i < n
void launch_rand_stream(uint64_t offset, int64_t n, uint32_t* out, cudaStream_t s);

// ---- scans and small utilities ---------------------------------------------------------
// exclusive scan of n int64 values; out has n+1 entries (out[n] = total). tmp >= scan_tmp_bytes(n)
size_t scan_tmp_bytes(int64_t n);
void launch_exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* tmp, cudaStream_t s);

// ---- K6b/c: per-patch shuffle and gather into the fit stream -------------------------------
// draws[p] = (n_p > 0 ? (n_p - 1) * mult : 0)
// draws[p] + exclusive scan roff[0..P] + plan9 = { n_claimed, lo, hi, off[lo], off[hi], roff[lo], roff[hi], roff[P], max n_p }
// ids[0..n): patches lo..lo+n-1 by decreasing point count; hist1024: 1024 ints of scratch
void launch_bv_hist(const int32_t* nbv, int64_t n, unsigned long long* out34, cudaStream_t s);
void launch_fed_update(const int64_t* off, int64_t n, int accumulate, int64_t* fed, cudaStream_t s);
void launch_size_order(const int64_t* off, int64_t lo, int64_t n, int32_t* hist1024, int32_t* ids, cudaStream_t s);
void launch_fit_plan(const int64_t* off, int64_t n_patches, int mult, int rank, int count, int64_t fixed_lo, int64_t fixed_hi,
                     int64_t* draws, int64_t* roff, void* scan_tmp, int64_t* plan9, cudaStream_t s);
// perm (patch-local) for every patch; rnd holds the stream starting at the handle's offset
struct ShuffleGatherArgs {
    const int64_t* off;          // patch offsets of this shard's first patch onwards (n_patches + 1)
    int64_t n_patches;
    const int64_t* roff;         // rand-draw offsets, same base
    const uint32_t* rnd;         // the shard's window of the rand() stream
    int do_shuffle;
    int is_rgb;                  // 0: height stream (f0 = y); 1: the field GP's second shuffle and colour streams (f0..f2)
    int64_t first_patch;         // absolute index of off[0]'s patch (rgbmean lookup)
    int64_t s_begin, s_count;    // the shard's window of the claimed-point stream
    const double *x1, *x2, *y;   // grouped points
    const uint32_t* rgb;         // b,g,r,a bytes per grouped point
    const double* rgbmean;       // 3 per patch (absolute patch index)
    int32_t* perm;               // out: local index of the point added at each stream position
    int32_t* forig;              // optional: perm + orig_base[patch] (continued fits: index in the concatenated streams)
    const int64_t* orig_base;    // points fed to each patch before this call (same base as off), with forig
    double *fx1, *fx2, *f0, *f1, *f2;
};
void launch_shuffle_gather(const ShuffleGatherArgs& a, int64_t max_patch_points, int32_t* patch_of, cudaStream_t s);
void launch_gather_rgb_stream(const int64_t* off, const int32_t* patch_of, const int32_t* perm, const double* x1, const double* x2,
                              const uint32_t* rgb, const double* rgbmean, int64_t first_patch, int64_t s_begin, int64_t s_count,
                              double* fx1, double* fx2, double* fr, double* fg, double* fb, cudaStream_t s);
void launch_gather_stream(const int64_t* off, const int32_t* patch_of, const int32_t* perm, const double* x1,
                          const double* x2, const double* y, int64_t s_begin, int64_t s_count, double* fx1,
                          double* fx2, double* fy, cudaStream_t s);

// ---- K7: SOGP fit ---------------------------------------------------------------------
struct SogpArgs {
    const int64_t* off;      // n_patches + 1 stream offsets (global patch numbering)
    const double *fx1, *fx2;       // fit stream in add order (already shuffled)
    const double* fy[3];           // outputs: heights (dout 1) or centred r,g,b (dout 3)
    int dout;
    const int32_t* forig;    // patch-local original position of each stream element
    const int32_t* patch_ids;  // patches to process (nullptr: first_patch + blockIdx.x)
    int64_t first_patch;
    int n_work;              // number of CTAs
    int ld;                  // storage leading dimension (N + 1 <= ld at all times)
    int capacity;
    double s20, eps_tol, p0, cl;
    // outputs, indexed by (patch - out_first) * capacity
    int64_t out_first;
    int32_t* nbv;
    int32_t* flags;
    double* o_alpha[3];
    double *o_b1, *o_b2;
    int32_t* o_idx;
    double *dumpC, *dumpQ;   // optional, (patch - out_first) * capacity^2
    int32_t* queue;          // overflow queue for the next bucket (nullptr: overflow impossible)
    int32_t* queue_count;
    const double* handoff_in;  // state slots written by the previous bucket (slot = blockIdx.x), or nullptr
    double* handoff_out;       // state slots for patches that outgrow this bucket (slot = queue position)
    double* spill;             // bucket 4 only: per-CTA state slices in global memory
    unsigned long long* stats;  // 11 counters, see gpc_stats
};
// bucket b supports ld <= {16, 32, 64, 118, 202}; bucket 4 keeps its state in global memory
int sogp_bucket_ld(int bucket);
size_t sogp_handoff_slot_bytes(int bucket, int dout);
int sogp_next_bucket(int bucket, int dout);
size_t sogp_spill_bytes_per_patch();
void launch_continue_partition(const int64_t* off, const int32_t* nbv, int64_t lo, int64_t n, int32_t* queues, int32_t* counts4,
                               cudaStream_t s);
cudaError_t launch_state_to_slots(int b0, const int32_t* ids, int64_t n_work, int64_t out_first, int cap, const int32_t* nbv,
                                  const double* alpha, const double* b1, const double* b2, const int32_t* bidx, const double* dumpC,
                                  const double* dumpQ, double* slots, cudaStream_t s);
cudaError_t launch_sogp_fit(int bucket, const SogpArgs& a, cudaStream_t s);

// ---- K8: grid prediction ----------------------------------------------------------------
struct PredictArgs {
    int64_t n_patches;           // patches of this shard
    const int32_t* nbv;          // per patch
    const int64_t* slot;         // exclusive scan of (nbv > 0), n_patches + 1
    int stride;                  // parameter stride per patch (capacity)
    const double *alpha, *b1, *b2;
    const double *quat, *mean, *rgbmean;  // may be nullptr (identity frame)
    // RGB field GP (may be nullptr): per patch nbv, alpha[3], BVs with the same stride
    const int32_t* rgb_nbv;
    const double *rgb_alpha[3], *rgb_b1, *rgb_b2;
    double res;
    int sz;
    double p0, cl;
    bool separable;              // gpc_config.decode_separable: kernel tables instead of one exp per (grid point, BV)
    uint8_t* out32;              // may be nullptr
    double* heights;             // may be nullptr
    int nmax, nrmax;             // largest BV count of the height / RGB GPs (table sizing)
    int rows, rowsP, nblk;       // set by launch_predict_grid: grid rows per block, padded, blocks per patch
    int group_doubles;           // shared-memory doubles per patch block
    int64_t n_groups;            // n_patches * nblk
};
void launch_compact_params(const int32_t* nbv, int64_t n_patches, int stride, const double* alpha, const double* b1,
                           const double* b2, const int32_t* idx, int64_t* widened, int64_t* bv_off, void* scan_tmp,
                           double* palpha, double* pb1, double* pb2, int32_t* pidx, cudaStream_t s);
void launch_flag_nonempty(const int32_t* nbv, const int32_t* rgb_nbv, int64_t n, int64_t* flags, int32_t* maxes, cudaStream_t s);
cudaError_t launch_predict_grid(const PredictArgs& a, cudaStream_t s);

// K9 (k_evaluate.cu): batched predict with sigma / conf, likelihood and likelihood gradient
struct EvalArgs {
    int64_t n_patches;
    const int32_t* nbv;
    const int64_t* off;            // n_patches + 1 offsets into the query arrays
    int stride;                    // parameter stride per patch (capacity); C at patch * stride^2, packed N x N
    int nmax;                      // largest nbv among the patches
    int dout;                      // 1: height GP, 3: RGB field GP (y, f: dout values per point)
    const double* alpha[3];
    const double *b1, *b2, *C;
    const double *x1, *x2, *y;     // y may be nullptr
    double p0, cl, c1, s20;        // c1 = -p0 / p1 (rbf_kernel.cpp:44)
    double pow2pi3;                // pow(2 pi, 3): the field GP's density normalisation (sparse_gp_field.hpp:350)
    int conf;
    double *f, *sigma, *lik, *dX;  // any may be nullptr
};
cudaError_t launch_evaluate(const EvalArgs& a, cudaStream_t s);
void launch_predict_points(const double* alpha, const double* b1, const double* b2, int N, double p0, double cl, const double* X,
                           int64_t m, double* f, cudaStream_t s);
cudaError_t measure_peak(int kind, int sm_count, cudaStream_t s, double* value);
void launch_debug_exp(const double* x, double* out, int64_t n, cudaStream_t s);

}  // namespace gpc
