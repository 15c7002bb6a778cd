/* gpc.h — C ABI of libgpc_b200.so, the B200 (sm_100a) implementation of the
 * gp_compressor compress / decompress hot path.
 *
 * The reference (nilsbore/gp_compressor, /root/reference) has no FFI or plugin layer: the
 * boundary of this path is the public surface of three C++ classes.  Each entry point
 * below names the reference member it replaces (file:line under /root/reference/src).
 * Plain pointers and sizes only; the caller owns every host buffer, the handle owns all
 * device memory and streams; no allocation crosses the ABI.  There is no CPU fallback:
 * every compute entry point fails with GPC_ERR_CUDA when no sm_100 device is usable.
 *
 * Point clouds are arrays of PCL PointXYZRGB records, 32 bytes each:
 *   float x, y, z, 1.0f | uint8 b, g, r, a | 12 bytes padding.
 *
 * Threading: one handle = one host thread = one CUDA device (gpc_config.device); several
 * handles may work on one device from different threads (two clouds in flight hide the
 * host-to-device copy of one behind the kernels of the other).
 * Multi-GPU runs use one process (or thread) and one handle per GPU; patches shard with
 * gpc_config.shard_rank / shard_count and no collective, or -- one cloud, binning sharded as
 * well -- with gpc_compress_shard_begin / _finish and one all-gather of two integers
 * (see DESIGN.md section 5).
 */
#ifndef GPC_H
#define GPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPC_OK 0
#define GPC_ERR_INVALID 1   /* bad argument / unsupported configuration */
#define GPC_ERR_CUDA 2      /* CUDA runtime failure or no usable device  */
#define GPC_ERR_STATE 3     /* call sequence error (e.g. decompress before compress) */
#define GPC_ERR_OVERFLOW 4  /* octree deeper than 21 levels (63-bit Morton code)     */

#define GPC_POINT_BYTES 32

typedef struct gpc_handle gpc_handle;

/* Constructor arguments of the reference classes on this path, plus the knobs the
 * reference hard-codes.  gpc_config_default() fills the reference's values, with ONE deviation: rgb = 0 (the reference always
 * fits the colour field GP next to the height GP, gp_compressor.cpp:163, and decodes C* + RGB mean, :367); set rgb = 1 for the
 * reference's behaviour -- with rgb = 0 the decoder paints every grid point with its patch's mean colour. */
typedef struct gpc_config {
    double res;         /* gp_compressor(cloud, res, sz): voxel size, gp_compressor.h:65 (default 0.1f) */
    int32_t sz;         /* gp_compressor(cloud, res, sz): grid side at decode, gp_compressor.h:65 (10)  */
    int32_t capacity;   /* sparse_gp(capacity, s0): max basis vectors, sparse_gp.h:48 (100)             */
    double s0;          /* sparse_gp(capacity, s0): noise variance s20, sparse_gp.h:48 (1e-1f)          */
    double eps_tol;     /* sparse_gp ctor literal, sparse_gp.hpp:31 (1e-6f)                             */
    double sigmaf_sq;   /* rbf_kernel(sigmaf_sq, l_sq), rbf_kernel.h:24 (100)                           */
    double l_sq;        /* rbf_kernel(sigmaf_sq, l_sq), rbf_kernel.h:24 (1)                             */
    int32_t leaf_order; /* 0: reverse Morton (PCL 1.7/1.8 leaf iterator), 1: Morton (PCL >= 1.9)        */
    int32_t shuffle;    /* 1: sparse_gp::shuffle before adding (sparse_gp.hpp:62-63); 0: input order    */
    int32_t rgb_rand;   /* 1: skip the rand() draws the RGB field GP consumes (gp_compressor.cpp:163)   */
    int32_t device;     /* CUDA device ordinal                                                          */
    int32_t shard_rank; /* this handle fits / decodes patches of shard shard_rank ...                   */
    int32_t shard_count;/* ... out of shard_count contiguous ranges of the patch visiting order (1)     */
    int32_t keep_state; /* 1: keep dense C and Q per patch (gpc_get_state, gpc_evaluate_patches, gpc_add_measurements) */
    int32_t rgb;        /* 1: also fit / decode the RGB field GP, sparse_gp_field (gp_compressor.cpp:163,334)  */
    double rgb_s0;      /* sparse_gp_field(capacity, s0), sparse_gp_field.h:43 (1e2f)                        */
    double rgb_eps_tol; /* sparse_gp_field ctor literal, sparse_gp_field.hpp:16 (1e-4f)                      */
    int32_t decode_separable; /* 0 (default): the grid decode evaluates rbf_kernel::kernel_function p0*exp(cl*(dx^2+dy^2)) per
                                 (grid point, BV) as the reference does (rbf_kernel.cpp:15-18 from sparse_gp.hpp:320-327 on the
                                 lattice of gp_compressor.cpp:320-328).  1: flagged fast mode -- the kernel separated on the
                                 lattice, (p0*exp(cl*dx^2))*exp(cl*dy^2) from per-patch tables: N(sz+rows) exps instead of
                                 N*sz^2, heights within a few ulp of the kernel value of the direct form               */
    int32_t pad0;
} gpc_config;

typedef struct gpc_sizes {
    int64_t n_in;        /* points given to the last compress                                  */
    int64_t n_patches;   /* octree leaves = patches (gp_compressor.cpp:182), all shards        */
    int64_t n_claimed;   /* points claimed by some patch (gp_compressor.cpp:88-97); after
                            gpc_compress_shard_*: by the patches this shard binned (own + halo) */
    int64_t n_bv_total;  /* sum of basis-vector counts over this shard's patches               */
    int64_t patch_lo;    /* first patch (gp_index) of this shard                               */
    int64_t patch_hi;    /* one past the last patch of this shard                              */
    int64_t n_decoded;   /* points produced by the last decompress                             */
    uint64_t rand_offset;/* rand() draws consumed so far by this handle                        */
    double lattice_min[3]; /* octree bounding-box minimum (PCL min_x_/min_y_/min_z_)           */
    uint32_t depth;      /* octree depth                                                       */
    uint32_t pad;
} gpc_sizes;

/* Event counters of the SOGP fit (the flop / byte accounting of DESIGN.md uses them) and
 * per-stage device times of the last call, in milliseconds (CUDA events). */
typedef struct gpc_stats {
    uint64_t n_add, n_first, n_sparse, n_full, n_del_cap, n_del_geo;
    uint64_t sum_n, sum_n2_common, sum_n2_sparse, sum_n2_full, sum_n2_del;
    uint64_t escalated[4];      /* patches that left SOGP bucket b for a larger one */
    uint64_t kernel_launches;   /* kernels launched by the last compress / decompress */
    /* ms_evaluate: the kernel of the last gpc_evaluate_patches(_rgb) */
    uint64_t rgb_n_sparse, rgb_n_full, rgb_n_del_cap, rgb_n_del_geo, rgb_sum_n2_common;  /* RGB field GP events */
    float ms_h2d, ms_lattice, ms_keys, ms_sort, ms_leaves, ms_rotation, ms_claim, ms_group,
          ms_shuffle, ms_fit, ms_d2h, ms_predict, ms_total, ms_fit_rgb, ms_evaluate, pad_;
    /* basis vectors per patch after the last fit of this shard: the largest count ("Max added", gp_compressor.cpp:165-174)
     * and a histogram -- bv_hist[k] = patches with k basis vectors for k < 32, bv_hist[32] = patches with 32 or more */
    uint64_t max_bv;
    uint64_t bv_hist[33];
} gpc_stats;

/* ---- lifetime ---------------------------------------------------------------------- */
int gpc_config_default(gpc_config* cfg);
int gpc_create(const gpc_config* cfg, gpc_handle** out);     /* gp_compressor ctor, gp_compressor.cpp:12-19 */
void gpc_destroy(gpc_handle* h);
const char* gpc_last_error(const gpc_handle* h);             /* replaces print + exit(0), gp_compressor.cpp:136-139 */
const char* gpc_version(void);

/* ---- compress: gp_compressor::save_compressed, gp_compressor.cpp:21-27 ---------------
 * = project_cloud (:177-249) + train_processes (:121-175).  Host buffer in; the timed
 * end-to-end call.  gpc_upload_cloud + gpc_compress_resident split the same work so that
 * the device part can be timed with the cloud already in HBM. */
int gpc_compress(gpc_handle* h, const void* cloud32_host, int64_t n_points);
int gpc_upload_cloud(gpc_handle* h, const void* cloud32_host, int64_t n_points);
int gpc_compress_resident(gpc_handle* h);

/* ---- sparse_gp<rbf_kernel,gaussian_noise>::add_measurements, sparse_gp.hpp:59-86 ------
 * on n_patches independent processes at once: patch p owns rows off[p]..off[p+1] of
 * X = (x1, x2) and y.  Includes the shuffle (rand stream continues from the handle's
 * offset).  Replaces any previous fit held by the handle. */
int gpc_fit_patches(gpc_handle* h, int64_t n_patches, const int64_t* off,
                    const double* x1, const double* x2, const double* y);

/* sparse_gp::add_measurements called AGAIN on fitted processes: the reference accumulates (sparse_gp.hpp:59-86 appends to
 * the existing BV / alpha / C / Q state; gp_mapping.cpp:338 adds every new cloud to the patch GPs).  The new points of
 * each patch are shuffled with the next rand() draws and the recursion continues from the kept state (the previous fit
 * and this handle need gpc_config.keep_state; same P as that fit, shard_count 1, rgb 0, capacity <= 117).  Patches
 * without new points keep their state.  BV indices (gpc_get_params) of points chosen from this call count on from the
 * points fed to the patch before. */
int gpc_add_measurements(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y);

/* Strong scaling of ONE cloud over several GPUs (replaces nothing in the reference, which is
 * single-threaded): every rank passes the same cloud (or NULL to use the cloud of gpc_upload_cloud) with
 * gpc_config.shard_rank / shard_count set.  begin replays the lattice on the whole cloud, cuts the visiting order
 * (gp_compressor.cpp:204-205) into shard_count ranges of coarse octree cells -- identically on every rank, from a
 * coarse Morton-key histogram -- and bins only the points within three voxels of the rank's own range (a patch's
 * claimed set depends on the leaf frames within two rings and those on the points within three), so the sort /
 * rotation / claim stages shrink with the shard count.  It returns how many patches and rand() draws the rank owns.
 * The ranks then all-gather those two integers (the one exchange of the path) and call finish with the sums over
 * the earlier ranks and over all ranks; finish shuffles and fits the owned patches.  Afterwards gpc_get_sizes
 * reports global patch indices; gpc_get_params / gpc_get_params_rgb / gpc_decompress / gpc_save cover the owned range,
 * and concatenating the ranks' outputs in rank order equals the single-GPU result bit for bit. */
int gpc_compress_shard_begin(gpc_handle* h, const void* cloud, int64_t n, int64_t* owned_patches, uint64_t* owned_draws);
int gpc_compress_shard_finish(gpc_handle* h, int64_t patches_before, uint64_t draws_before, int64_t patches_total,
                              uint64_t draws_total);

/* ---- decompress: gp_compressor::load_compressed, gp_compressor.cpp:267-386 ------------
 * Writes sz*sz points per non-empty patch, patches in gp_index order.  out may be NULL
 * (device-only run, result stays in HBM).  capacity_points guards the host buffer. */
int gpc_decompress(gpc_handle* h, void* out_cloud32_host, int64_t capacity_points, int64_t* n_out);
int gpc_decompress_resident(gpc_handle* h, int64_t* n_out);
/* f* of sparse_gp::predict_measurements for every grid point of the last decompress */
int gpc_get_heights(gpc_handle* h, double* heights_host, int64_t capacity);

/* ---- sparse_gp::predict_measurements, sparse_gp.hpp:299-351 ---------------------------
 * on one patch (gp_index) at m arbitrary local coordinates X (m x 2, row-major).
 * sigma may be NULL; when given it needs keep_state (conf = false branch, :346). */
int gpc_predict(gpc_handle* h, int64_t patch, const double* X, int64_t m, double* f, double* sigma);

/* Next rows N2 / N4: batched evaluation of the fitted height GPs at ragged per-patch point sets -- what
 * gp_registration::compute_transformation (gp_registration.cpp:175-194) asks of every patch, one point at a time:
 *   f      sparse_gp::predict_measurements mean                         sparse_gp.hpp:299-351
 *   sigma  its sigconf output: conf == 0 sqrt(s20 + k** + k'Ck), conf != 0 the 0..100 confidence (:340-346)
 *   lik    sparse_gp::compute_likelihoods -> likelihood                 sparse_gp.hpp:407-425
 *   dX     sparse_gp::compute_derivatives -> likelihood_dx, 3 per point sparse_gp.hpp:459-502, rbf_kernel.cpp:38-46
 * Patches are the first P of this shard (local indices), points of patch p are [off[p], off[p+1]) of x1 / x2 / y
 * (host arrays; y may be NULL when lik and dX are NULL).  Outputs are host arrays and may be NULL.  Needs a fit made
 * with gpc_config.keep_state (the C matrices).  Device time of the kernel: gpc_stats.ms_evaluate. */
int gpc_evaluate_patches(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y,
                         int conf, double* f, double* sigma, double* lik, double* dX);

/* The same for the RGB field GPs (gpc_config.rgb = 1 and keep_state = 1 at compress time): sparse_gp_field::
 * predict_measurements with sigconf / conf (sparse_gp_field.hpp:267-320), compute_likelihoods (:322-351) and
 * compute_derivatives (:353-393), which gp_registration.cpp:177,194 calls beside the height GP's.  y3 and f3 hold three
 * values per point (the colours centred on the patch mean, as the fit saw them); dX three per point with dX[0] = 0. */
int gpc_evaluate_patches_rgb(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y3,
                             int conf, double* f3, double* sigma, double* lik, double* dX);

/* ---- results ---------------------------------------------------------------------------
 * Any output pointer may be NULL.  Per-patch arrays are indexed by gp_index over ALL
 * patches (frames are computed on every shard); fitted parameters cover this shard only
 * and are addressed relative to patch_lo. */
int gpc_get_sizes(gpc_handle* h, gpc_sizes* s);
int gpc_get_stats(gpc_handle* h, gpc_stats* s);
/* rotations / means / RGB_means (gp_compressor.h:36-40), leaf keys and centres */
int gpc_get_patches(gpc_handle* h, uint64_t* code, float* center3, int32_t* n_candidates,
                    double* R9, double* quat4, double* mean3, double* rgbmean3, int64_t* patch_off);
/* owner[i] = gp_index of the patch that claimed input point i, or -1;  stream arrays in
 * patch-major claim order: original index and local (x1, x2, y) of gp_compressor.cpp:146-155 */
int gpc_get_assignment(gpc_handle* h, int32_t* owner, int32_t* stream_index,
                       double* x1, double* x2, double* y, int32_t* perm);
/* alpha, BV (sparse_gp.h:22,26) per patch: bv_off has (patch_hi - patch_lo + 1) entries */
int gpc_get_params(gpc_handle* h, int32_t* nbv, int64_t* bv_off, int32_t* bv_index,
                   double* bv1, double* bv2, double* alpha, int32_t* flags);
int gpc_get_state(gpc_handle* h, int64_t patch, double* C, double* Q);  /* N x N, row-major */
/* RGB field GP (next-row N1, gpc_config.rgb): per patch nbv, BVs and alpha (3 per BV: r, g, b), packed like gpc_get_params;
 * perm = the field GP's own shuffle.  n_bv_total_rgb is returned through the last argument. */
int gpc_get_params_rgb(gpc_handle* h, int32_t* nbv, int64_t* bv_off, int32_t* bv_index, double* bv1, double* bv2, double* alpha3,
                       int32_t* perm, int64_t* n_bv_total_rgb);
/* decode-only use: install fitted parameters (and optionally frames) from the host */
int gpc_set_params(gpc_handle* h, int64_t n_patches, const int32_t* nbv, const double* bv1,
                   const double* bv2, const double* alpha, const double* quat4,
                   const double* mean3, const double* rgbmean3);
int gpc_set_rand_offset(gpc_handle* h, uint64_t offset);
/* the cudaStream_t every kernel and copy of this handle is issued on (for CUDA-event timing by the caller) */
int gpc_get_stream(gpc_handle* h, void** stream);

/* ---- wire format (next-row N3): what gp_compressor::save_compressed(name) ignores ---------------
 * The reference's GP path has no on-disk format (save_compressed drops its file name,
 * gp_compressor.cpp:21-27); the K-SVD codec's .pccode layout (dictionary_representation.cpp:173-248) is the
 * model: little-endian header, then per-patch records.  gpc_save writes this shard's fitted patches
 * (N, quaternion, mean, RGB mean, BV, alpha); gpc_load installs them into a handle so that gpc_decompress
 * can run in another process.  File layout: "GPCB200\0" | u32 version | config (res, sz, capacity, s0,
 * eps_tol, sigmaf_sq, l_sq) | i64 n_patches | i64 n_bv_total | nbv[i32] | quat[4 f64] | mean[3 f64] |
 * rgbmean[3 f64] | bv1, bv2, alpha [f64, packed] | (version 2) i32 has_rgb | rgb_s0, rgb_eps_tol | i64 n_bv_rgb |
 * rgb nbv[i32] | rgb bv1, bv2 [f64] | rgb alpha [3 f64 per BV]. */
int gpc_save(gpc_handle* h, const char* path, int64_t* bytes_written);
int gpc_get_config(const gpc_handle* h, gpc_config* cfg);  /* the handle's current configuration (gpc_load updates it) */
int gpc_load(gpc_handle* h, const char* path);

/* ---- test hooks: the device versions of the canonical primitives ----------------------- */
int gpc_debug_exp(gpc_handle* h, const double* x, double* out, int64_t n);
int gpc_debug_rand(gpc_handle* h, uint64_t offset, int64_t n, uint32_t* out);
/* roofline denominators measured on the device: kind 0 = FP64 FLOP/s (DFMA), kind 1 = shared-memory bytes/s with 128-bit loads, kind 2 = with 64-bit loads */
int gpc_debug_peak(gpc_handle* h, int kind, double* value);

/* ---- sharding rule (pure host arithmetic, no device needed) -------------------------------
 * Patches [lo, hi) that shard_rank of shard_count owns, given the patch offsets off[0..n_patches]:
 * contiguous ranges of the visiting order holding about equal numbers of claimed points.  This
 * is the rule gpc_compress / gpc_fit_patches / gpc_set_params apply internally. */
int gpc_shard_range(const int64_t* off, int64_t n_patches, int32_t shard_rank, int32_t shard_count, int64_t* lo, int64_t* hi);

#ifdef __cplusplus
}
#endif
#endif /* GPC_H */
