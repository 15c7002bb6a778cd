// gpc_api.cu — the extern "C" layer of libgpc_b200.so (see include/gpc.h).
// Host-side orchestration only: buffer management, stage sequencing, sharding, timing.
// All arithmetic of the path runs in the kernels of k_*.cu; there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>

#include "gpc_internal.h"

namespace gpc {
thread_local uint64_t g_launches = 0;
}

using namespace gpc;

struct gpc_handle {
    gpc_config cfg;
    std::string err;
    cudaStream_t stream = nullptr;
    uint64_t rand_offset = 0;
    gpc_stats stats;
    // sizes
    int64_t n_in = 0, n_patches = 0, n_claimed = 0, patch_lo = 0, patch_hi = 0, n_decoded = 0, n_bv_total = -1;
    uint32_t depth = 0;
    double lattice_min[3] = {0, 0, 0};
    bool have_fit = false, have_frames = false, have_binning = false, have_cloud = false, params_packed = false;
    bool have_state = false;     // dumpC / dumpQ (and the fed counts) belong to the fit the handle holds
    bool heights_valid = false;  // heights hold the grid of the last decompress
    int64_t s_begin = 0, s_count = 0;  // stream range of this shard
    // device buffers
    DevBuf cloud;                               // input cloud (n_in * 32)
    DevBuf off, x1, x2, y, perm, patch_of;      // claimed stream, patch-major
    DevBuf fx1, fx2, fy;                        // fit stream (add order)
    DevBuf draws, roff, rnd, scan_tmp, small;   // rand bookkeeping, scan scratch, small readbacks
    DevBuf nbv, flags, alpha, b1, b2, bidx, dumpC, dumpQ, queue0, queue1, qcount, kstats, hand0, hand1, spill;
    DevBuf nonempty, slot, out32, heights;
    DevBuf bv_off, palpha, pb1, pb2, pidx;      // packed copies of the fitted parameters (what leaves for the host)
    // RGB field GP (gpc_config.rgb): its own shuffle, fit stream and parameters
    DevBuf perm_rgb, fcr, fcg, fcb, r_nbv, r_flags, r_alpha0, r_alpha1, r_alpha2, r_b1, r_b2, r_bidx, kstats_rgb;
    bool have_rgb = false;
    DevBuf quat, mean, rgbmean, Rm, center, code, ncand, owner, st_idx;
    DevBuf tmpA, tmpB, tmpC, lat_state, ev_in, ev_out, size_ids, size_hist, sel_cloud, sel_idx, coarse_hist, fed, forig, cont_q, cont_slots, queueB, handB, queueC0, queueC1, handC0, handC1, r_dumpC, r_dumpQ;
    cudaStream_t stream2 = nullptr;   // side streams: the bucket chains of the larger patches run beside bucket 0 of the rest
    cudaStream_t stream3 = nullptr;
    // two or three bucket chains: measured per handle on its first calls (run_fit, chain_tune_after)
    int chain_pick = 3, chain_calls = 0, chain_last = 0;   // mode in use, calls at this size, mode of the call in flight
    int64_t chain_PL = 0;
    float chain_ms[2] = {0.f, 0.f};                        // last fit-stage time with two / three chains
    cudaEvent_t ev_a = nullptr, ev_a2 = nullptr, ev_a3 = nullptr;
    int32_t* pinned_counts = nullptr;  // 16 pinned host words for asynchronous read-backs of device counters
    // sharded binning (gpc_compress_shard_begin / _finish): this shard's patches are local indices [own_lo, own_hi) of a
    // binning that holds the shard's key range plus its halo; global patch index = local + gshift
    bool shard_mode = false;
    int64_t own_lo = 0, own_hi = 0, gshift = 0, patches_total = 0, n_sel = 0;
    uint64_t draws_before = 0, draws_total = 0, draws_owned = 0;

    // binning scratch
    DevBuf cpt;  // 16-byte point records in input order (point_keys_kernel)
    DevBuf chunk_queue, chunk_hand, chunk_queue1, chunk_hand1;  // second ping-pong pair of the main-stream bucket chain (chain B)
    DevBuf keys, keys2, vals, vals2, ovals, ovals2, sort_tmp, flags64, ex, leaf_of, leaf_start, leaf_code_a, spt, nbr, nnbr,
        center_a, Rm_a, ncand_a, pt0, pt1, pt2, hbuf, rgb, leaf_sums;
    std::vector<cudaEvent_t> ev;
};

namespace {

int fail(gpc_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(h, GPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    } while (0)

// Layout of gpc_handle::small, the 256-byte device block for the few scalars that travel between kernels and the host.
struct SmallScratch {
    unsigned long long pad0[4];
    unsigned long long n_valid;      // leaves_fused_kernel: finite points ...
    unsigned long long n_leaves;     // ... and leaves (read back together)
    unsigned long long pad1[2];
    int64_t plan[9];                 // fit_plan_kernel: n_claimed, lo, hi, off[lo], off[hi], roff[lo], roff[hi], roff[P], max patch
    int64_t pad2[3];
    int64_t owned_range[2];          // owned_range_kernel (sharded binning): first / one-past-last owned patch
    uint64_t key_range[2];           // shard_splitters_kernel: [klo, khi) of this rank
    int32_t maxes[2];                // flag_nonempty_kernel: largest BV count of the height / RGB GPs
    int32_t pad3[14];
};
static_assert(sizeof(SmallScratch) == 256, "scratch layout");

// Two or three bucket chains for the fit (run_fit)?  Three win on a GPU whose process runs alone (C2: 1.85 against 1.91 ms).  In a
// process that holds an NCCL communicator set up with NVLink SHARP (NVLS, NCCL's default on NVSwitch nodes) three were measured
// SLOWER (2.13 against 1.92 ms at 2 and 8 ranks; NCCL_NVLS_ENABLE=0 removes it; the chains' own event timeline is unchanged).  The
// library cannot know what else its process has set up, so the handle keeps measuring: it remembers the last fit time of either
// mode, uses the faster, and tries the other one every 8th call.  Results do not depend on the mode.
void chain_tune_after(gpc_handle* h) {
    const int mode = h->chain_last;
    h->chain_last = 0;
    if (mode == 0) return;
    if (h->chain_calls > 1) h->chain_ms[mode == 3] = h->stats.ms_fit;   // the first call of a size is cold (allocations): not used
    const float t2 = h->chain_ms[0], t3 = h->chain_ms[1];
    if (t2 > 0.f && t3 > 0.f) {
        if (h->chain_pick == 3 && t2 < 0.98f * t3) h->chain_pick = 2;
        else if (h->chain_pick == 2 && t3 < 0.98f * t2) h->chain_pick = 3;
    }
}

struct StageTimer {
    gpc_handle* h;
    size_t used = 0;
    std::vector<std::pair<float*, std::pair<size_t, size_t>>> spans;
    explicit StageTimer(gpc_handle* hh) : h(hh) {}
    size_t mark() {
        if (used == h->ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            h->ev.push_back(e);
        }
        cudaEventRecord(h->ev[used], h->stream);
        return used++;
    }
    void span(float* dst, size_t a, size_t b) { spans.push_back({dst, {a, b}}); }
    void resolve() {
        for (auto& s : spans) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[s.second.first], h->ev[s.second.second]);
            *s.first += ms;
        }
        chain_tune_after(h);
    }
};

// the bucket-0 SOGP kernel stages 16-point tiles with bulk copies from 16-byte aligned addresses and may read up to
// 136 bytes past the last stream element (k_sogp.cu, HalfTile): every stream buffer ends with this much slack
constexpr size_t STREAM_PAD = 256;

double kernel_cl(const gpc_config& c) { return (double)(-0.5f) / c.l_sq; }  // -0.5f / p(1), rbf_kernel.cpp:17

// Patches [lo, hi) of shard r out of c: contiguous ranges of the visiting order with about
// equal numbers of claimed points.  off is the host copy of the patch offsets.
void shard_range(const std::vector<int64_t>& off, int r, int c, int64_t* lo, int64_t* hi) {
    const int64_t P = (int64_t)off.size() - 1;
    const int64_t total = off[P];
    auto bound = [&](int k) -> int64_t {
        if (k <= 0) return 0;
        if (k >= c) return P;
        int64_t target = (int64_t)((__int128)total * k / c);
        int64_t a = 0, b = P;  // first p with off[p] >= target
        while (a < b) {
            int64_t m = (a + b) / 2;
            if (off[m] >= target) b = m; else a = m + 1;
        }
        return a;
    };
    *lo = bound(r);
    *hi = bound(r + 1);
}

// Runs the SOGP bucket chain for one family of processes (dout 1: heights, dout 3: RGB field).
// One chain of buckets: `work` patches (ids) start in bucket b0 -- from scratch (b0 = 0) or from the state slots
// hand_in of bucket b0 - 1's format (continued fits) -- and climb to larger buckets as they outgrow them.
int run_buckets(gpc_handle* h, SogpArgs& a, int need_ld, int64_t lo, uint64_t* escalated, int b0, int64_t work, const int32_t* ids,
                const double* hand_in, int step0 = 0, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    a.spill = nullptr;
    int step = step0;   // parity picks the queue / hand-off buffers a bucket writes (the other pair is being read)
    for (int b = b0; b < 5 && work > 0; b = sogp_next_bucket(b, a.dout), step++) {
        const int bl = sogp_bucket_ld(b);
        const bool final_bucket = need_ld <= bl;
        a.ld = final_bucket ? need_ld : bl;
        a.patch_ids = ids;
        a.first_patch = lo;
        a.n_work = (int)work;
        DevBuf& q = (step & 1) ? h->queue1 : h->queue0;
        DevBuf& ho = (step & 1) ? h->hand1 : h->hand0;
        DevBuf& hi_ = (step & 1) ? h->hand0 : h->hand1;
        a.queue = final_bucket ? nullptr : q.as<int32_t>();
        a.queue_count = h->qcount.as<int32_t>() + b;
        a.handoff_in = (step > step0) ? hi_.as<double>() : hand_in;
        a.handoff_out = nullptr;
        if (!final_bucket) {
            CK(ho.reserve((size_t)work * sogp_handoff_slot_bytes(b, a.dout)));
            a.handoff_out = ho.as<double>();
        }
        if (b == 4) {
            CK(h->spill.reserve((size_t)work * sogp_spill_bytes_per_patch()));
            a.spill = h->spill.as<double>();
        }
        CK(launch_sogp_fit(b, a, st));
        if (final_bucket) break;
        int32_t qn = 0;
        CK(cudaMemcpyAsync(&qn, h->qcount.as<int32_t>() + b, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (escalated && b < 4) escalated[b] += (uint64_t)qn;
        work = qn;
        ids = q.as<int32_t>();
    }
    return GPC_OK;
}

// One chain of buckets advanced ONE LEVEL per call, so that several chains (on different streams) can be walked breadth
// first: a patch is a strictly sequential recursion, the largest patches climb through several buckets, and the host must not
// sit in one chain's synchronisation while another chain's next kernel could already be launched.
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield");
#endif
}

struct BucketChain {
    cudaStream_t st = nullptr;
    int b = 0;                     // next bucket to launch
    int64_t work = 0;              // patches of that launch
    const int32_t* ids = nullptr;  // their ids
    const double* hand_in = nullptr;
    DevBuf *q[2] = {nullptr, nullptr}, *hand[2] = {nullptr, nullptr};  // ping-pong overflow queues / hand-off slots of this chain
    int32_t* qcount = nullptr;     // 8 device counters of this chain
    int step = 0;
    bool pending = false;          // a launched bucket whose overflow count has not been read yet
    int32_t* qn = nullptr;         // pinned host word the count is copied to (a pageable target would make the copy block)
};
// launches the chain's next bucket (if any); returns 0 / error.  After it, chain.pending tells whether a count is outstanding.
int chain_launch(gpc_handle* h, SogpArgs a, int need_ld, int64_t lo, BucketChain& ch) {
    ch.pending = false;
    if (ch.b >= 5 || ch.work <= 0) { ch.work = 0; return GPC_OK; }
    const int bl = sogp_bucket_ld(ch.b);
    const bool final_bucket = need_ld <= bl;
    a.spill = nullptr;
    a.ld = final_bucket ? need_ld : bl;
    a.patch_ids = ch.ids;
    a.first_patch = lo;
    a.n_work = (int)ch.work;
    DevBuf& q = *ch.q[ch.step & 1];
    DevBuf& ho = *ch.hand[ch.step & 1];
    a.queue = nullptr; a.handoff_out = nullptr;
    a.queue_count = ch.qcount + ch.b;
    a.handoff_in = ch.hand_in;
    if (!final_bucket) {
        CK(q.reserve((size_t)ch.work * sizeof(int32_t)));
        CK(ho.reserve((size_t)ch.work * sogp_handoff_slot_bytes(ch.b, a.dout)));
        a.queue = q.as<int32_t>();
        a.handoff_out = ho.as<double>();
    }
    if (ch.b == 4) {
        CK(h->spill.reserve((size_t)ch.work * sogp_spill_bytes_per_patch()));
        a.spill = h->spill.as<double>();
    }
    CK(launch_sogp_fit(ch.b, a, ch.st));
    if (final_bucket) { ch.work = 0; return GPC_OK; }
    CK(cudaMemcpyAsync(ch.qn, ch.qcount + ch.b, sizeof(int32_t), cudaMemcpyDeviceToHost, ch.st));
    ch.pending = true;
    return GPC_OK;
}
// waits for the outstanding count and moves the chain to its next bucket
int chain_collect(gpc_handle* h, int dout, uint64_t* escalated, BucketChain& ch) {
    if (!ch.pending) return GPC_OK;
    CK(cudaStreamSynchronize(ch.st));
    ch.pending = false;
    if (escalated && ch.b < 4) escalated[ch.b] += (uint64_t)*ch.qn;
    ch.ids = ch.q[ch.step & 1]->as<int32_t>();
    ch.hand_in = ch.hand[ch.step & 1]->as<double>();
    ch.work = *ch.qn;
    ch.b = sogp_next_bucket(ch.b, dout);
    ch.step++;
    return GPC_OK;
}

// Shuffle + SOGP fit of patches [patch_lo, patch_hi) over the stream held in h->off/x1/x2/y.
// cont: continue the kept state of every patch with the new stream (gpc_add_measurements)
int run_fit(gpc_handle* h, StageTimer& tm, bool cont = false) {
    const gpc_config& c = h->cfg;
    cudaStream_t st = h->stream;
    const int64_t P = h->n_patches;
    if (c.capacity < 1) return fail(h, GPC_ERR_INVALID, "capacity must be >= 1 (the reference's -1 / 0 modes are not built)");
    const int need_ld = c.capacity + 1;
    if (need_ld > sogp_bucket_ld(4)) return fail(h, GPC_ERR_INVALID, "capacity > 201 is not supported");
    // shard bounds, rand-stream window and the largest patch: computed on the device, nine scalars come back
    size_t t0 = tm.mark();
    const int mult = c.shuffle ? (c.rgb_rand ? 2 : 1) : 0;
    CK(h->draws.reserve((P + 1) * sizeof(int64_t)));
    CK(h->roff.reserve((P + 2) * sizeof(int64_t)));
    CK(h->scan_tmp.reserve(scan_tmp_bytes(P + 1)));
    CK(h->small.reserve(sizeof(SmallScratch)));
    int64_t* d_plan = h->small.as<SmallScratch>()->plan;
    launch_fit_plan(h->off.as<int64_t>(), P, mult, c.shard_rank, c.shard_count, h->shard_mode ? h->own_lo : -1, h->own_hi,
                    h->draws.as<int64_t>(), h->roff.as<int64_t>(), h->scan_tmp.p, d_plan, st);
    int64_t plan[9];
    CK(cudaMemcpyAsync(plan, d_plan, sizeof(plan), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->n_claimed = plan[0];
    h->patch_lo = plan[1];
    h->patch_hi = plan[2];
    const int64_t lo = h->patch_lo, hi = h->patch_hi, PL = hi - lo;
    h->s_begin = plan[3];
    h->s_count = plan[4] - plan[3];
    const int64_t S = h->n_claimed;
    // replicated binning: the shard's window of the global draw sequence comes from the prefix over all patches;
    // sharded binning: the draws of the earlier shards come from the all-gather between begin and finish
    const uint64_t draws_lo = h->shard_mode ? h->draws_before : (uint64_t)plan[5];
    const uint64_t draws_hi = draws_lo + (uint64_t)(plan[6] - plan[5]);
    const uint64_t draws_all = h->shard_mode ? h->draws_total : (uint64_t)plan[7];
    const int64_t max_np = plan[8];
    CK(h->perm.reserve(std::max<int64_t>(S, 1) * sizeof(int32_t) + STREAM_PAD));
    CK(h->patch_of.reserve(std::max<int64_t>(S, 1) * sizeof(int32_t)));
    CK(h->fx1.reserve(std::max<int64_t>(S, 1) * sizeof(double) + STREAM_PAD));
    CK(h->fx2.reserve(std::max<int64_t>(S, 1) * sizeof(double) + STREAM_PAD));
    CK(h->fy.reserve(std::max<int64_t>(S, 1) * sizeof(double) + STREAM_PAD));
    if (PL > 0 && h->s_count > 0) {
        const int64_t nd = (int64_t)(draws_hi - draws_lo);
        if (c.shuffle && nd > 0) {
            CK(h->rnd.reserve(nd * sizeof(uint32_t)));
            launch_rand_stream(h->rand_offset + draws_lo, nd, h->rnd.as<uint32_t>(), st);
        }
        ShuffleGatherArgs sg;
        sg.off = h->off.as<int64_t>() + lo; sg.n_patches = PL; sg.roff = h->roff.as<int64_t>() + lo; sg.rnd = h->rnd.as<uint32_t>();
        sg.do_shuffle = c.shuffle; sg.is_rgb = 0; sg.first_patch = lo; sg.s_begin = h->s_begin; sg.s_count = h->s_count;
        sg.x1 = h->x1.as<double>(); sg.x2 = h->x2.as<double>(); sg.y = h->y.as<double>(); sg.rgb = nullptr; sg.rgbmean = nullptr;
        sg.perm = h->perm.as<int32_t>();
        sg.forig = nullptr; sg.orig_base = nullptr;
        if (cont) {  // BV indices of a continued fit count from the points fed before
            CK(h->forig.reserve(std::max<int64_t>(S, 1) * sizeof(int32_t) + STREAM_PAD));
            sg.forig = h->forig.as<int32_t>();
            sg.orig_base = h->fed.as<int64_t>() + lo;
        }
        sg.fx1 = h->fx1.as<double>(); sg.fx2 = h->fx2.as<double>(); sg.f0 = h->fy.as<double>(); sg.f1 = sg.f2 = nullptr;
        launch_shuffle_gather(sg, max_np, h->patch_of.as<int32_t>(), st);
    }
    CK(h->fed.reserve(std::max<int64_t>(P, 1) * sizeof(int64_t)));
    launch_fed_update(h->off.as<int64_t>(), P, cont ? 1 : 0, h->fed.as<int64_t>(), st);
    h->rand_offset += draws_all;
    size_t t1 = tm.mark();
    tm.span(&h->stats.ms_shuffle, t0, t1);
    // outputs
    const int cap = c.capacity;
    const int64_t PLa = std::max<int64_t>(PL, 1);
    CK(h->nbv.reserve(PLa * sizeof(int32_t)));
    CK(h->flags.reserve(PLa * sizeof(int32_t)));
    CK(h->alpha.reserve(PLa * cap * sizeof(double)));
    CK(h->b1.reserve(PLa * cap * sizeof(double)));
    CK(h->b2.reserve(PLa * cap * sizeof(double)));
    CK(h->bidx.reserve(PLa * cap * sizeof(int32_t)));
    if (c.keep_state) {
        CK(h->dumpC.reserve(PLa * (size_t)cap * cap * sizeof(double)));
        CK(h->dumpQ.reserve(PLa * (size_t)cap * cap * sizeof(double)));
    }
    CK(h->queue0.reserve(PLa * sizeof(int32_t)));
    CK(h->queue1.reserve(PLa * sizeof(int32_t)));
    CK(h->qcount.reserve(64 * sizeof(int32_t)));
    CK(h->kstats.reserve(64 * sizeof(unsigned long long)));   // 16 event counters, then max BV count + histogram (34)
    CK(cudaMemsetAsync(h->kstats.p, 0, 64 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(h->qcount.p, 0, 8 * sizeof(int32_t), st));
    SogpArgs a;
    a.off = h->off.as<int64_t>();
    a.fx1 = h->fx1.as<double>(); a.fx2 = h->fx2.as<double>();
    a.fy[0] = h->fy.as<double>(); a.fy[1] = a.fy[2] = nullptr;
    a.dout = 1;
    a.forig = cont ? h->forig.as<int32_t>() : h->perm.as<int32_t>();
    a.capacity = cap;
    a.s20 = c.s0; a.eps_tol = c.eps_tol; a.p0 = c.sigmaf_sq; a.cl = kernel_cl(c);
    a.out_first = lo;
    a.nbv = h->nbv.as<int32_t>(); a.flags = h->flags.as<int32_t>();
    a.o_alpha[0] = h->alpha.as<double>(); a.o_alpha[1] = a.o_alpha[2] = nullptr;
    a.o_b1 = h->b1.as<double>(); a.o_b2 = h->b2.as<double>();
    a.o_idx = h->bidx.as<int32_t>();
    a.dumpC = c.keep_state ? h->dumpC.as<double>() : nullptr;
    a.dumpQ = c.keep_state ? h->dumpQ.as<double>() : nullptr;
    a.stats = h->kstats.as<unsigned long long>();
    if (!cont) {
        CK(h->size_ids.reserve((size_t)PL * sizeof(int32_t)));
        CK(h->size_hist.reserve(1024 * sizeof(int32_t)));
        launch_size_order(h->off.as<int64_t>(), lo, PL, h->size_hist.as<int32_t>(), h->size_ids.as<int32_t>(), st);
        if (need_ld > sogp_bucket_ld(0) && PL >= 4096) {
            // A patch is a strictly sequential recursion, the patches that outgrow bucket 0 continue in the two-warp kernel (and
            // the largest climb further), and a continuation can only be launched when the kernel that handed it off has ended:
            // run as one chain, the stage would end with the continuations of the whole cloud after the last bucket-0 block.
            // So the size-ordered list (largest first) is cut in three chains on three streams, advanced breadth first:
            //   A  = the largest patches (1/32, at least one wave of 8 warps per SM): longest continuations, started first;
            //   B1 = the next ones up to 3/5 of the list (about three quarters of the points): their continuations run beside
            //        bucket 0 of B2 instead of after it;
            //   B2 = the small patches on the main stream: what it hands off has few points left, so the tail is short.
            if (h->chain_PL == 0 || PL > 2 * h->chain_PL || 2 * PL < h->chain_PL) {   // another workload: measure again
                h->chain_pick = 3; h->chain_calls = 0; h->chain_PL = PL; h->chain_ms[0] = h->chain_ms[1] = 0.f;
            }
            h->chain_calls++;
            const bool explore = (h->chain_calls & 7) == 2;   // calls 2, 10, 18, ...: the mode not in use
            const int chains = explore ? (h->chain_pick == 3 ? 2 : 3) : h->chain_pick;
            h->chain_last = chains;
            const int64_t nA = std::min<int64_t>(PL, std::max<int64_t>(2368, (PL / 32))) & ~(int64_t)1;
            const int64_t nB1 = chains == 3 ? std::max<int64_t>(0, (PL * 3 / 5 - nA)) & ~(int64_t)1 : 0;   // cut at 0.5 .. 0.7 of the list: within noise on C2   // cut at 0.5 .. 0.7 of the list: within noise on C2
            const int64_t nB2 = PL - nA - nB1;
            CK(h->qcount.reserve(64 * sizeof(int32_t)));
            CK(cudaMemsetAsync(h->qcount.p, 0, 64 * sizeof(int32_t), st));
            BucketChain ch[3];
            const int32_t* ids = h->size_ids.as<int32_t>();
            ch[0].st = h->stream2; ch[0].work = nA; ch[0].ids = ids;
            ch[0].q[0] = &h->queue0; ch[0].q[1] = &h->queue1; ch[0].hand[0] = &h->hand0; ch[0].hand[1] = &h->hand1;
            ch[1].st = h->stream3; ch[1].work = nB1; ch[1].ids = ids + nA;
            ch[1].q[0] = &h->queueC0; ch[1].q[1] = &h->queueC1; ch[1].hand[0] = &h->handC0; ch[1].hand[1] = &h->handC1;
            ch[2].st = st; ch[2].work = nB2; ch[2].ids = ids + nA + nB1;
            ch[2].q[0] = &h->queueB; ch[2].q[1] = &h->chunk_queue1; ch[2].hand[0] = &h->handB; ch[2].hand[1] = &h->chunk_hand1;
            for (int i = 0; i < 3; i++) { ch[i].b = 0; ch[i].qn = h->pinned_counts + i; ch[i].qcount = h->qcount.as<int32_t>() + 8 * i; }
            CK(cudaEventRecord(h->ev_a, st));
            CK(cudaStreamWaitEvent(h->stream2, h->ev_a, 0));
            CK(cudaStreamWaitEvent(h->stream3, h->ev_a, 0));
            int rc;
            for (int i = 0; i < 3; i++)
                if ((rc = chain_launch(h, a, need_ld, lo, ch[i]))) return rc;
            while (ch[0].pending || ch[1].pending || ch[2].pending) {
                // whichever chain's count has arrived moves on first.  The host thread has nothing else to do, so it polls --
                // but not back to back: every query takes the driver's lock, which another handle's thread needs to launch.
                bool moved = false;
                for (int i = 0; i < 3; i++) {
                    if (!ch[i].pending) continue;
                    const cudaError_t qe = cudaStreamQuery(ch[i].st);
                    if (qe == cudaErrorNotReady) continue;
                    if (qe != cudaSuccess) CK(qe);
                    if ((rc = chain_collect(h, 1, h->stats.escalated, ch[i]))) return rc;
                    if ((rc = chain_launch(h, a, need_ld, lo, ch[i]))) return rc;
                    moved = true;
                }
                if (!moved) {
                    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(4);
                    while (std::chrono::steady_clock::now() < until) cpu_relax();
                }
            }
            CK(cudaEventRecord(h->ev_a2, h->stream2));
            CK(cudaStreamWaitEvent(st, h->ev_a2, 0));
            CK(cudaEventRecord(h->ev_a3, h->stream3));
            CK(cudaStreamWaitEvent(st, h->ev_a3, 0));
        } else {
            int rc = run_buckets(h, a, need_ld, lo, h->stats.escalated, 0, PL, h->size_ids.as<int32_t>(), nullptr);
            if (rc) return rc;
        }
    } else {
        // patches with new points, by the size of their kept state: one chain of buckets per slot format
        CK(h->cont_q.reserve((size_t)(4 * PLa + 8) * sizeof(int32_t)));
        int32_t* d_cnt = h->cont_q.as<int32_t>() + 4 * PLa;
        launch_continue_partition(h->off.as<int64_t>(), h->nbv.as<int32_t>(), lo, PL, h->cont_q.as<int32_t>(), d_cnt, st);
        int32_t cnt4[4] = {0, 0, 0, 0};
        CK(cudaMemcpyAsync(cnt4, d_cnt, sizeof(cnt4), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int b0 = 1; b0 <= 4; b0++) {
            const int64_t work = cnt4[b0 - 1];
            if (work == 0) continue;
            const int32_t* ids = h->cont_q.as<int32_t>() + (int64_t)(b0 - 1) * PL;
            CK(h->cont_slots.reserve((size_t)work * sogp_handoff_slot_bytes(b0 - 1, 1)));
            CK(launch_state_to_slots(b0, ids, work, lo, cap, h->nbv.as<int32_t>(), h->alpha.as<double>(), h->b1.as<double>(),
                                     h->b2.as<double>(), h->bidx.as<int32_t>(), h->dumpC.as<double>(), h->dumpQ.as<double>(),
                                     h->cont_slots.as<double>(), st));
            CK(cudaMemsetAsync(h->qcount.p, 0, 8 * sizeof(int32_t), st));
            int rc = run_buckets(h, a, need_ld, lo, nullptr, b0, work, ids, h->cont_slots.as<double>());
            if (rc) return rc;
        }
    }
    size_t t2 = tm.mark();
    tm.span(&h->stats.ms_fit, t1, t2);
    // ---- RGB field GP (sparse_gp_field<rbf_kernel, gaussian_noise_3d>): same points, own shuffle, 3 outputs ----
    h->have_rgb = false;
    size_t t_pack0 = t2;
    if (c.rgb && h->have_binning) {
        if (!(c.shuffle && c.rgb_rand)) return fail(h, GPC_ERR_INVALID, "rgb = 1 needs shuffle = 1 and rgb_rand = 1");
        const int64_t Sa = std::max<int64_t>(S, 1);
        CK(h->perm_rgb.reserve(Sa * sizeof(int32_t) + STREAM_PAD));
        CK(h->fcr.reserve(Sa * sizeof(double) + STREAM_PAD));
        CK(h->fcg.reserve(Sa * sizeof(double) + STREAM_PAD));
        CK(h->fcb.reserve(Sa * sizeof(double) + STREAM_PAD));
        CK(h->r_nbv.reserve(PLa * sizeof(int32_t)));
        CK(h->r_flags.reserve(PLa * sizeof(int32_t)));
        CK(h->r_alpha0.reserve(PLa * cap * sizeof(double)));
        CK(h->r_alpha1.reserve(PLa * cap * sizeof(double)));
        CK(h->r_alpha2.reserve(PLa * cap * sizeof(double)));
        CK(h->r_b1.reserve(PLa * cap * sizeof(double)));
        CK(h->r_b2.reserve(PLa * cap * sizeof(double)));
        CK(h->r_bidx.reserve(PLa * cap * sizeof(int32_t)));
        CK(h->kstats_rgb.reserve(16 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(h->kstats_rgb.p, 0, 16 * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(h->qcount.p, 0, 8 * sizeof(int32_t), st));
        if (PL > 0 && h->s_count > 0) {
            // the height fit is done with fx1 / fx2: reuse them for the field GP's order
            ShuffleGatherArgs sg;
            sg.off = h->off.as<int64_t>() + lo; sg.n_patches = PL; sg.roff = h->roff.as<int64_t>() + lo; sg.rnd = h->rnd.as<uint32_t>();
            sg.do_shuffle = c.shuffle && c.rgb_rand; sg.is_rgb = 1; sg.first_patch = lo; sg.s_begin = h->s_begin; sg.s_count = h->s_count;
            sg.x1 = h->x1.as<double>(); sg.x2 = h->x2.as<double>(); sg.y = nullptr; sg.rgb = h->rgb.as<uint32_t>();
            sg.rgbmean = h->rgbmean.as<double>();
            sg.perm = h->perm_rgb.as<int32_t>();
            sg.forig = nullptr; sg.orig_base = nullptr;
            sg.fx1 = h->fx1.as<double>(); sg.fx2 = h->fx2.as<double>();
            sg.f0 = h->fcr.as<double>(); sg.f1 = h->fcg.as<double>(); sg.f2 = h->fcb.as<double>();
            launch_shuffle_gather(sg, max_np, h->patch_of.as<int32_t>(), st);
        }
        SogpArgs r = a;
        r.fy[0] = h->fcr.as<double>(); r.fy[1] = h->fcg.as<double>(); r.fy[2] = h->fcb.as<double>();
        r.dout = 3;
        r.forig = h->perm_rgb.as<int32_t>();
        r.s20 = c.rgb_s0; r.eps_tol = c.rgb_eps_tol;
        r.nbv = h->r_nbv.as<int32_t>(); r.flags = h->r_flags.as<int32_t>();
        r.o_alpha[0] = h->r_alpha0.as<double>(); r.o_alpha[1] = h->r_alpha1.as<double>(); r.o_alpha[2] = h->r_alpha2.as<double>();
        r.o_b1 = h->r_b1.as<double>(); r.o_b2 = h->r_b2.as<double>();
        r.o_idx = h->r_bidx.as<int32_t>();
        r.dumpC = r.dumpQ = nullptr;
        if (c.keep_state) {  // C of the field GPs (sigma / likelihood / gradient of the colours, gpc_evaluate_patches_rgb)
            CK(h->r_dumpC.reserve(PLa * (size_t)cap * cap * sizeof(double)));
            CK(h->r_dumpQ.reserve(PLa * (size_t)cap * cap * sizeof(double)));
            r.dumpC = h->r_dumpC.as<double>(); r.dumpQ = h->r_dumpQ.as<double>();
        }
        r.stats = h->kstats_rgb.as<unsigned long long>();
        {
            int rc = run_buckets(h, r, need_ld, lo, nullptr, 0, PL, h->size_ids.as<int32_t>(), nullptr);
            if (rc) return rc;
        }
        h->have_rgb = true;
        size_t t2b = tm.mark();
        tm.span(&h->stats.ms_fit_rgb, t2, t2b);
        t_pack0 = t2b;
    }
    // pack the parameters on the device: what the host fetches is sum(nbv) entries, not capacity per patch
    CK(h->bv_off.reserve((PLa + 1) * sizeof(int64_t)));
    CK(h->nonempty.reserve((PLa + 1) * sizeof(int64_t)));
    CK(h->scan_tmp.reserve(scan_tmp_bytes(PLa + 1)));
    CK(h->palpha.reserve(PLa * cap * sizeof(double)));
    CK(h->pb1.reserve(PLa * cap * sizeof(double)));
    CK(h->pb2.reserve(PLa * cap * sizeof(double)));
    CK(h->pidx.reserve(PLa * cap * sizeof(int32_t)));
    launch_compact_params(h->nbv.as<int32_t>(), PL, cap, h->alpha.as<double>(), h->b1.as<double>(), h->b2.as<double>(),
                          h->bidx.as<int32_t>(), h->nonempty.as<int64_t>(), h->bv_off.as<int64_t>(), h->scan_tmp.p,
                          h->palpha.as<double>(), h->pb1.as<double>(), h->pb2.as<double>(), h->pidx.as<int32_t>(), st);
    launch_bv_hist(h->nbv.as<int32_t>(), PL, h->kstats.as<unsigned long long>() + 16, st);
    int64_t tot = 0;
    CK(cudaMemcpyAsync(&tot, h->bv_off.as<int64_t>() + PL, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->n_bv_total = tot;
    size_t t3 = tm.mark();
    tm.span(&h->stats.ms_group, t_pack0, t3);
    h->have_fit = true;
    h->have_state = c.keep_state != 0;
    h->params_packed = true;
    return GPC_OK;
}

int read_fit_stats(gpc_handle* h) {
    unsigned long long k[50];
    CK(cudaMemcpyAsync(k, h->kstats.p, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    gpc_stats& s = h->stats;
    s.max_bv = k[16];
    for (int i = 0; i < 33; i++) s.bv_hist[i] = k[17 + i];
    s.n_add = k[0]; s.n_first = k[1]; s.n_sparse = k[2]; s.n_full = k[3]; s.n_del_cap = k[4]; s.n_del_geo = k[5];
    s.sum_n = k[6]; s.sum_n2_common = k[7]; s.sum_n2_sparse = k[8]; s.sum_n2_full = k[9]; s.sum_n2_del = k[10];
    if (h->have_rgb) {
        CK(cudaMemcpyAsync(k, h->kstats_rgb.p, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        s.rgb_n_sparse = k[2]; s.rgb_n_full = k[3]; s.rgb_n_del_cap = k[4]; s.rgb_n_del_geo = k[5]; s.rgb_sum_n2_common = k[7];
    }
    return GPC_OK;
}

void reset_stats(gpc_handle* h) {
    std::memset(&h->stats, 0, sizeof(h->stats));
    g_launches = 0;
}

int run_decode(gpc_handle* h, bool want_cloud, bool want_heights, StageTimer& tm) {
    const gpc_config& c = h->cfg;
    cudaStream_t st = h->stream;
    if (!h->have_fit) return fail(h, GPC_ERR_STATE, "decompress before compress / fit / set_params");
    const int64_t PL = h->patch_hi - h->patch_lo;
    const int64_t g2 = (int64_t)c.sz * c.sz;
    size_t t0 = tm.mark();
    CK(h->nonempty.reserve((PL + 1) * sizeof(int64_t)));
    CK(h->slot.reserve((PL + 2) * sizeof(int64_t)));
    CK(h->scan_tmp.reserve(scan_tmp_bytes(PL + 1)));
    const bool with_rgb = h->have_rgb && h->have_frames;
    CK(h->small.reserve(sizeof(SmallScratch)));
    int32_t* d_maxes = h->small.as<SmallScratch>()->maxes;
    launch_flag_nonempty(h->nbv.as<int32_t>(), with_rgb ? h->r_nbv.as<int32_t>() : nullptr, PL, h->nonempty.as<int64_t>(), d_maxes, st);
    launch_exclusive_scan_i64(h->nonempty.as<int64_t>(), h->slot.as<int64_t>(), PL, h->scan_tmp.p, st);
    int64_t n_nonempty = 0;
    int32_t maxes[2] = {0, 0};
    CK(cudaMemcpyAsync(&n_nonempty, h->slot.as<int64_t>() + PL, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(maxes, d_maxes, sizeof(maxes), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->n_decoded = n_nonempty * g2;
    if (want_cloud) CK(h->out32.reserve(std::max<int64_t>(h->n_decoded, 1) * GPC_POINT_BYTES));
    if (want_heights) CK(h->heights.reserve(std::max<int64_t>(h->n_decoded, 1) * sizeof(double)));
    h->heights_valid = want_heights;
    PredictArgs a;
    a.n_patches = PL;
    a.nbv = h->nbv.as<int32_t>();
    a.slot = h->slot.as<int64_t>();
    a.stride = c.capacity;
    a.alpha = h->alpha.as<double>(); a.b1 = h->b1.as<double>(); a.b2 = h->b2.as<double>();
    if (h->have_frames) {
        a.quat = h->quat.as<double>() + 4 * h->patch_lo;
        a.mean = h->mean.as<double>() + 3 * h->patch_lo;
        a.rgbmean = h->rgbmean.as<double>() + 3 * h->patch_lo;
    } else {
        a.quat = a.mean = a.rgbmean = nullptr;
    }
    a.rgb_nbv = nullptr;
    a.rgb_alpha[0] = a.rgb_alpha[1] = a.rgb_alpha[2] = a.rgb_b1 = a.rgb_b2 = nullptr;
    if (with_rgb) {
        a.rgb_nbv = h->r_nbv.as<int32_t>();
        a.rgb_alpha[0] = h->r_alpha0.as<double>(); a.rgb_alpha[1] = h->r_alpha1.as<double>(); a.rgb_alpha[2] = h->r_alpha2.as<double>();
        a.rgb_b1 = h->r_b1.as<double>(); a.rgb_b2 = h->r_b2.as<double>();
    }
    a.nmax = std::max(maxes[0], 1); a.nrmax = with_rgb ? std::max(maxes[1], 1) : 0;
    a.res = c.res; a.sz = c.sz; a.p0 = c.sigmaf_sq; a.cl = kernel_cl(c);
    a.separable = c.decode_separable != 0;
    a.out32 = want_cloud ? h->out32.as<uint8_t>() : nullptr;
    a.heights = want_heights ? h->heights.as<double>() : nullptr;
    if (launch_predict_grid(a, st) != cudaSuccess)
        return fail(h, GPC_ERR_INVALID, "decode: sz x capacity too large for the shared-memory kernel tables");
    CK(cudaGetLastError());
    size_t t1 = tm.mark();
    tm.span(&h->stats.ms_predict, t0, t1);
    return GPC_OK;
}

int bits_for(uint64_t v) {  // bits needed to represent values 0..v
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

// project_cloud (gp_compressor.cpp:177-249) on the resident cloud: fills off/x1/x2/y and frames.
// sharded: bin only the points within three voxels of this rank's range of the visiting order (gpc_compress_shard_begin)
int run_binning(gpc_handle* h, StageTimer& tm, bool sharded = false) {
    const gpc_config& c = h->cfg;
    cudaStream_t st = h->stream;
    int64_t n = h->n_in;
    const uint8_t* cloud = h->cloud.as<uint8_t>();
    h->shard_mode = false;
    if (n > 0x7fffffff) return fail(h, GPC_ERR_INVALID, "more than 2^31-1 points");
    h->have_binning = h->have_frames = h->have_fit = false;
    CK(h->small.reserve(sizeof(SmallScratch)));
    unsigned long long* d_nvalid = &h->small.as<SmallScratch>()->n_valid;
    // ---- lattice replay: search + adopt on the device, the host only polls "found" ----
    size_t t0 = tm.mark();
    CK(h->lat_state.reserve(sizeof(LatticeState)));
    LatticeState L0;
    std::memset(&L0, 0, sizeof(L0));
    L0.lat.res = c.res;
    L0.best = ~0ull;
    CK(cudaMemcpyAsync(h->lat_state.p, &L0, sizeof(L0), cudaMemcpyHostToDevice, st));
    LatticeState Lh = L0;
    if (n > 0) CK(launch_lattice_replay(cloud, n, h->lat_state.as<LatticeState>(), st));
    CK(cudaMemcpyAsync(&Lh, h->lat_state.p, sizeof(Lh), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (Lh.defined && Lh.lat.depth > 21) return fail(h, GPC_ERR_OVERFLOW, "octree deeper than 21 levels: res too small for the cloud extent");
    struct { unsigned depth; bool defined; double mn[3]; } L;
    L.depth = Lh.lat.depth;
    L.defined = Lh.defined != 0;
    for (int a = 0; a < 3; a++) L.mn[a] = Lh.lat.mn[a];
    h->depth = L.depth;
    for (int a = 0; a < 3; a++) h->lattice_min[a] = L.mn[a];
    size_t t1 = tm.mark();
    tm.span(&h->stats.ms_lattice, t0, t1);
    h->n_patches = 0;
    h->n_claimed = 0;
    CK(h->off.reserve(2 * sizeof(int64_t)));
    CK(h->owner.reserve(std::max<int64_t>(n, 1) * sizeof(int32_t)));
    if (n > 0) CK(cudaMemsetAsync(h->owner.p, 0xff, n * sizeof(int32_t), st));
    if (!L.defined) {  // no finite point: no leaves
        CK(cudaMemsetAsync(h->off.p, 0, sizeof(int64_t), st));
        h->have_binning = h->have_frames = true;
        return GPC_OK;
    }
    const LatticeDev lat = Lh.lat;
    if (sharded) {
        // ---- the rank's key range: quantiles of a sorted key sample of the WHOLE cloud (identical on every rank), then
        // the points within the halo of that range, compacted into a private cloud.  No host round trip until n_sel. ----
        const int depth = (int)L.depth;
        CK(h->keys.reserve(n * sizeof(uint64_t)));
        CK(h->vals.reserve(n * sizeof(uint32_t)));
        launch_point_keys(cloud, n, lat, h->keys.as<uint64_t>(), h->vals.as<uint32_t>(), nullptr, st);
        const int64_t stride = std::max<int64_t>(1, n >> 20);
        const int64_t m = (n + stride - 1) / stride;
        CK(h->coarse_hist.reserve((size_t)m * 2 * (sizeof(uint64_t) + sizeof(uint32_t)) + 64));
        uint64_t* smp = h->coarse_hist.as<uint64_t>();
        uint64_t* smp2 = smp + m;
        uint32_t* dv = reinterpret_cast<uint32_t*>(smp2 + m);
        uint32_t* dv2 = dv + m;
        CK(h->sort_tmp.reserve(radix_sort_tmp_bytes(std::max(n, m))));
        launch_shard_sample(h->keys.as<uint64_t>(), n, stride, m, smp, dv, st);
        const int which_s = launch_radix_sort(smp, dv, smp2, dv2, m, 3 * depth + 1, h->sort_tmp.p, st);
        CK(h->small.reserve(sizeof(SmallScratch)));
        uint64_t* d_range = h->small.as<SmallScratch>()->key_range;
        CK(h->flags64.reserve((m + 1) * sizeof(int64_t)));
        CK(h->ex.reserve((m + 1) * sizeof(int64_t)));
        CK(h->scan_tmp.reserve(scan_tmp_bytes(m)));
        launch_shard_splitters(which_s ? smp2 : smp, m, stride, depth, c.shard_rank, c.shard_count, c.leaf_order, h->flags64.as<int64_t>(),
                               h->ex.as<int64_t>(), h->scan_tmp.p, d_range, st);
        // halo test + scan + compaction of the selected records in one pass (sel_cloud / sel_idx sized for the whole cloud)
        CK(h->sel_cloud.reserve(std::max<int64_t>(n, 1) * GPC_POINT_BYTES));
        CK(h->sel_idx.reserve(std::max<int64_t>(n, 1) * sizeof(int32_t)));
        CK(h->scan_tmp.reserve(std::max(shard_select_tmp_bytes(n), scan_tmp_bytes(n))));
        launch_shard_select_compact(h->keys.as<uint64_t>(), n, depth, d_range, cloud, h->sel_cloud.as<uint8_t>(), h->sel_idx.as<int32_t>(),
                                    d_nvalid, h->scan_tmp.p, st);
        unsigned long long nsel_u = 0;
        CK(cudaMemcpyAsync(&nsel_u, d_nvalid, sizeof(nsel_u), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int64_t n_sel = (int64_t)nsel_u;
        h->n_sel = n_sel;
        h->shard_mode = true;
        cloud = h->sel_cloud.as<uint8_t>();
        n = n_sel;
        if (n == 0) {
            CK(cudaMemsetAsync(h->off.p, 0, sizeof(int64_t), st));
            h->have_binning = h->have_frames = true;
            return GPC_OK;
        }
    }
    // ---- keys + Morton sort ----
    CK(h->keys.reserve(n * sizeof(uint64_t)));
    CK(h->keys2.reserve(n * sizeof(uint64_t)));
    CK(h->vals.reserve(n * sizeof(uint32_t)));
    CK(h->vals2.reserve(n * sizeof(uint32_t)));
    CK(h->sort_tmp.reserve(radix_sort_tmp_bytes(n)));
    CK(h->cpt.reserve(n * 16));
    launch_point_keys(cloud, n, lat, h->keys.as<uint64_t>(), h->vals.as<uint32_t>(), h->cpt.p, st);
    size_t t2 = tm.mark();
    tm.span(&h->stats.ms_keys, t1, t2);
    int which = launch_radix_sort(h->keys.as<uint64_t>(), h->vals.as<uint32_t>(), h->keys2.as<uint64_t>(), h->vals2.as<uint32_t>(), n,
                                  3 * (int)L.depth + 1, h->sort_tmp.p, st);
    uint64_t* skeys = which ? h->keys2.as<uint64_t>() : h->keys.as<uint64_t>();
    uint32_t* svals = which ? h->vals2.as<uint32_t>() : h->vals.as<uint32_t>();
    uint64_t* fkeys = which ? h->keys.as<uint64_t>() : h->keys2.as<uint64_t>();  // free buffer for the owner keys
    size_t t3 = tm.mark();
    tm.span(&h->stats.ms_sort, t2, t3);
    // ---- leaves: valid count, leaf numbering, leaf tables and the sorted-point gather in one pass; ONE round trip ----
    CK(h->leaf_of.reserve(n * sizeof(int32_t)));
    CK(h->leaf_start.reserve((n + 1) * sizeof(int64_t)));
    CK(h->leaf_code_a.reserve(n * sizeof(uint64_t)));
    CK(h->spt.reserve(n * 16));
    CK(h->scan_tmp.reserve(std::max(leaves_fused_tmp_bytes(n), scan_tmp_bytes(n))));
    launch_leaves_fused(skeys, svals, n, L.depth, h->cpt.p, h->leaf_of.as<int32_t>(), h->leaf_start.as<int64_t>(),
                        h->leaf_code_a.as<uint64_t>(), h->spt.p, d_nvalid, h->scan_tmp.p, st);
    unsigned long long nvp[2] = {0, 0};
    CK(cudaMemcpyAsync(nvp, d_nvalid, sizeof(nvp), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t n_valid = (int64_t)nvp[0];
    const int64_t P = (int64_t)nvp[1];
    h->n_patches = P;
    size_t t4 = tm.mark();
    tm.span(&h->stats.ms_leaves, t3, t4);
    // ---- neighbours + rotation ----
    CK(h->nbr.reserve(P * 27 * sizeof(int32_t)));
    CK(h->nnbr.reserve(P * sizeof(int32_t)));
    CK(h->center_a.reserve(P * 3 * sizeof(float)));
    CK(h->Rm_a.reserve(P * 9 * sizeof(double)));
    CK(h->ncand_a.reserve(P * sizeof(int32_t)));
    CK(h->leaf_sums.reserve(P * 10 * sizeof(double)));
    const double radius = (double)(std::sqrt(3.0f) / 2.0f) * c.res;  // gp_compressor.cpp:194
    const double r2 = radius * radius;
    const double half = c.res / 2.0f;                                // gp_compressor.cpp:85
    launch_leaf_neighbours(h->leaf_code_a.as<uint64_t>(), P, lat, h->nbr.as<int32_t>(), h->nnbr.as<int32_t>(), h->center_a.as<float>(), st);
    launch_leaf_rotation(h->spt.p, h->leaf_start.as<int64_t>(), h->nbr.as<int32_t>(), h->nnbr.as<int32_t>(), h->center_a.as<float>(), P,
                         r2, h->leaf_sums.as<double>(), h->Rm_a.as<double>(), h->ncand_a.as<int32_t>(), st);
    size_t t5 = tm.mark();
    tm.span(&h->stats.ms_rotation, t4, t5);
    // ---- claim ----
    CK(h->ovals.reserve(n_valid * sizeof(uint32_t)));
    CK(h->ovals2.reserve(n_valid * sizeof(uint32_t)));
    CK(h->pt0.reserve(n_valid * sizeof(double)));
    CK(h->pt1.reserve(n_valid * sizeof(double)));
    CK(h->pt2.reserve(n_valid * sizeof(double)));
    launch_claim(h->spt.p, h->leaf_of.as<int32_t>(), h->nbr.as<int32_t>(), h->nnbr.as<int32_t>(), h->center_a.as<float>(),
                 h->Rm_a.as<double>(), h->ncand_a.as<int32_t>(), n_valid, P, r2, half, c.leaf_order, fkeys, h->ovals.as<uint32_t>(),
                 h->pt0.as<double>(), h->pt1.as<double>(), h->pt2.as<double>(), st);
    size_t t6 = tm.mark();
    tm.span(&h->stats.ms_claim, t5, t6);
    // ---- group by owner (stable) ----
    // the Morton-sorted keys are dead now (leaf codes extracted): reuse their buffer as the sort's second key buffer
    int which2 = launch_radix_sort(fkeys, h->ovals.as<uint32_t>(), skeys, h->ovals2.as<uint32_t>(), n_valid, bits_for((uint64_t)P),
                                   h->sort_tmp.p, st);
    const uint64_t* gkeys = which2 ? skeys : fkeys;
    const uint32_t* gvals = which2 ? h->ovals2.as<uint32_t>() : h->ovals.as<uint32_t>();
    size_t t7 = tm.mark();
    tm.span(&h->stats.ms_sort, t6, t7);
    CK(h->off.reserve((P + 1) * sizeof(int64_t)));
    launch_patch_bounds(gkeys, n_valid, P, h->off.as<int64_t>(), st);
    int64_t S = 0;
    CK(cudaMemcpyAsync(&S, h->off.as<int64_t>() + P, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->n_claimed = S;
    const int64_t Sa = std::max<int64_t>(S, 1);
    CK(h->st_idx.reserve(Sa * sizeof(int32_t)));
    CK(h->hbuf.reserve(Sa * sizeof(double)));
    CK(h->x1.reserve(Sa * sizeof(double)));
    CK(h->x2.reserve(Sa * sizeof(double)));
    CK(h->y.reserve(Sa * sizeof(double)));
    CK(h->rgb.reserve(Sa * sizeof(uint32_t)));
    launch_group_gather(gkeys, gvals, svals, h->spt.p, h->pt0.as<double>(), h->pt1.as<double>(), h->pt2.as<double>(), S,
                        h->st_idx.as<int32_t>(), h->hbuf.as<double>(), h->x1.as<double>(), h->x2.as<double>(), h->rgb.as<uint32_t>(),
                        h->owner.as<int32_t>(), st);
    CK(h->code.reserve(P * sizeof(uint64_t)));
    CK(h->center.reserve(P * 3 * sizeof(float)));
    CK(h->ncand.reserve(P * sizeof(int32_t)));
    CK(h->Rm.reserve(P * 9 * sizeof(double)));
    CK(h->quat.reserve(P * 4 * sizeof(double)));
    CK(h->mean.reserve(P * 3 * sizeof(double)));
    CK(h->rgbmean.reserve(P * 3 * sizeof(double)));
    launch_patch_frames(h->off.as<int64_t>(), P, c.leaf_order, h->hbuf.as<double>(), h->rgb.as<uint32_t>(), h->leaf_code_a.as<uint64_t>(),
                        h->center_a.as<float>(), h->Rm_a.as<double>(), h->ncand_a.as<int32_t>(), h->y.as<double>(),
                        h->code.as<uint64_t>(), h->center.as<float>(), h->ncand.as<int32_t>(), h->Rm.as<double>(),
                        h->quat.as<double>(), h->mean.as<double>(), h->rgbmean.as<double>(), st);
    size_t t8 = tm.mark();
    tm.span(&h->stats.ms_group, t7, t8);
    CK(cudaGetLastError());
    h->have_binning = h->have_frames = true;
    return GPC_OK;
}

int compress_resident_impl(gpc_handle* h, StageTimer& tm) {
    int rc = run_binning(h, tm);
    if (rc) return rc;
    if (h->n_patches == 0) {
        h->patch_lo = h->patch_hi = 0;
        h->s_begin = h->s_count = 0;
        h->have_fit = true;
        h->n_bv_total = 0;
        h->params_packed = false;
        CK(h->nbv.reserve(sizeof(int32_t)));
        return GPC_OK;
    }
    return run_fit(h, tm);
}

}  // namespace

extern "C" {

const char* gpc_version(void) { return "gpc_b200 0.1 (sm_100a)"; }

int gpc_config_default(gpc_config* c) {
    if (!c) return GPC_ERR_INVALID;
    c->res = (double)0.1f;         // gp_compressor.h:65
    c->sz = 10;                    // gp_compressor.h:65
    c->capacity = 100;             // sparse_gp.h:48
    c->s0 = (double)1e-1f;         // sparse_gp.h:48
    c->eps_tol = (double)1e-6f;    // sparse_gp.hpp:31
    c->sigmaf_sq = (double)100e-0f;  // rbf_kernel.h:24
    c->l_sq = 1.0;                 // rbf_kernel.h:24
    c->leaf_order = 0;
    c->shuffle = 1;
    c->rgb_rand = 1;
    c->device = 0;
    c->shard_rank = 0;
    c->shard_count = 1;
    c->keep_state = 0;
    c->rgb = 0;                      // next-row N1, opt-in
    c->rgb_s0 = (double)1e2f;        // sparse_gp_field.h:43
    c->rgb_eps_tol = (double)1e-4f;  // sparse_gp_field.hpp:16
    c->decode_separable = 0;         // the reference's direct kernel evaluation
    c->pad0 = 0;
    return GPC_OK;
}

int gpc_create(const gpc_config* cfg, gpc_handle** out) {
    if (!cfg || !out) return GPC_ERR_INVALID;
    *out = nullptr;
    if (!(cfg->res > 0) || cfg->sz < 1 || cfg->shard_count < 1 || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count ||
        !(cfg->l_sq > 0) || cfg->capacity < 1 || cfg->capacity > 201 || !(cfg->s0 > 0) || !(cfg->sigmaf_sq > 0))
        return GPC_ERR_INVALID;  // (the reference's unbounded capacity 0 / -1 modes are not built: DESIGN.md section 6)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return GPC_ERR_CUDA;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return GPC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return GPC_ERR_CUDA;
    if (prop.major != 10) return GPC_ERR_CUDA;  // sm_100a SASS only
    gpc_handle* h = new gpc_handle();
    h->cfg = *cfg;
    std::memset(&h->stats, 0, sizeof(h->stats));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return GPC_ERR_CUDA; }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    auto bail = [&]() {
        if (h->ev_a) cudaEventDestroy(h->ev_a);
        if (h->ev_a2) cudaEventDestroy(h->ev_a2);
        if (h->ev_a3) cudaEventDestroy(h->ev_a3);
        if (h->stream2) cudaStreamDestroy(h->stream2);
        if (h->stream3) cudaStreamDestroy(h->stream3);
        if (h->pinned_counts) cudaFreeHost(h->pinned_counts);
        cudaStreamDestroy(h->stream);
        delete h;
        return GPC_ERR_CUDA;
    };
    if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_a3, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_a2, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&h->pinned_counts), 16 * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess)
        return bail();
    // process-wide jump-ahead tables of the glibc rand() generator: built once, whichever thread creates a handle first
    static RandTables tables;
    static std::once_flag tables_once;
    std::call_once(tables_once, [] { rand_tables_init(&tables); });
    if (rand_upload_tables(&tables) != cudaSuccess) return bail();
    *out = h;
    return GPC_OK;
}

void gpc_destroy(gpc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->cloud, &h->off, &h->x1, &h->x2, &h->y, &h->perm, &h->patch_of, &h->fx1, &h->fx2, &h->fy, &h->draws,
                      &h->roff, &h->rnd, &h->scan_tmp, &h->small, &h->nbv, &h->flags, &h->alpha, &h->b1, &h->b2, &h->bidx,
                      &h->dumpC, &h->dumpQ, &h->queue0, &h->queue1, &h->hand0, &h->hand1, &h->spill, &h->qcount, &h->kstats, &h->bv_off, &h->palpha, &h->pb1, &h->pb2, &h->pidx, &h->perm_rgb, &h->fcr, &h->fcg, &h->fcb, &h->r_nbv, &h->r_flags,
                      &h->r_alpha0, &h->r_alpha1, &h->r_alpha2, &h->r_b1, &h->r_b2, &h->r_bidx, &h->kstats_rgb, &h->nonempty, &h->slot, &h->out32,
                      &h->heights, &h->quat, &h->mean, &h->rgbmean, &h->Rm, &h->center, &h->code, &h->ncand, &h->owner,
                      &h->st_idx, &h->tmpA, &h->tmpB, &h->tmpC, &h->lat_state, &h->ev_in, &h->ev_out, &h->size_ids, &h->size_hist, &h->sel_cloud, &h->sel_idx, &h->coarse_hist, &h->fed, &h->forig, &h->cont_q, &h->cont_slots, &h->queueB, &h->handB, &h->queueC0, &h->queueC1, &h->handC0, &h->handC1, &h->r_dumpC, &h->r_dumpQ, &h->keys, &h->keys2, &h->vals, &h->vals2, &h->ovals,
                      &h->cpt, &h->chunk_queue, &h->chunk_hand, &h->chunk_queue1, &h->chunk_hand1, &h->ovals2, &h->sort_tmp, &h->flags64, &h->ex, &h->leaf_of, &h->leaf_start, &h->leaf_code_a, &h->spt,
                      &h->nbr, &h->nnbr, &h->center_a, &h->Rm_a, &h->ncand_a, &h->pt0, &h->pt1, &h->pt2, &h->hbuf, &h->rgb, &h->leaf_sums};
    for (DevBuf* b : bufs) b->release();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_a2) cudaEventDestroy(h->ev_a2);
    if (h->ev_a3) cudaEventDestroy(h->ev_a3);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->stream3) { cudaStreamSynchronize(h->stream3); cudaStreamDestroy(h->stream3); }
    if (h->pinned_counts) cudaFreeHost(h->pinned_counts);
    cudaStreamDestroy(h->stream);
    delete h;
}

const char* gpc_last_error(const gpc_handle* h) { return h ? h->err.c_str() : "null handle"; }

int gpc_get_stream(gpc_handle* h, void** stream) {
    if (!h || !stream) return GPC_ERR_INVALID;
    *stream = (void*)h->stream;
    return GPC_OK;
}

int gpc_set_rand_offset(gpc_handle* h, uint64_t offset) {
    if (!h) return GPC_ERR_INVALID;
    h->rand_offset = offset;
    return GPC_OK;
}

static int fit_patches_impl(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y, bool cont) {
    if (!h || P < 0 || !off) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (cont) {
        if (!h->have_fit || !h->cfg.keep_state || !h->have_state || !h->dumpC.p)
            return fail(h, GPC_ERR_STATE, "gpc_add_measurements needs a previous fit on this handle made with gpc_config.keep_state");
        if (h->shard_mode || h->cfg.shard_count != 1 || P != h->n_patches || h->patch_lo != 0 || h->patch_hi != h->n_patches)
            return fail(h, GPC_ERR_INVALID, "gpc_add_measurements: same patches as the previous fit, one shard");
        if (h->cfg.rgb && h->have_rgb) return fail(h, GPC_ERR_INVALID, "gpc_add_measurements continues the height GPs only (rgb = 0)");
        if (h->cfg.capacity > 117) return fail(h, GPC_ERR_INVALID, "gpc_add_measurements supports capacity <= 117");
    }
    h->shard_mode = false;
    if (off[0] != 0) return fail(h, GPC_ERR_INVALID, "off[0] must be 0");
    for (int64_t p = 0; p < P; p++)
        if (off[p + 1] < off[p] || off[p + 1] - off[p] > 0x7fffffff) return fail(h, GPC_ERR_INVALID, "offsets must be non-decreasing");
    const int64_t S = off[P];
    if (S > 0 && (!x1 || !x2 || !y)) return GPC_ERR_INVALID;
    reset_stats(h);
    StageTimer tm(h);
    cudaStream_t st = h->stream;
    size_t tA = tm.mark();
    CK(h->off.reserve((P + 1) * sizeof(int64_t)));
    CK(h->x1.reserve(std::max<int64_t>(S, 1) * sizeof(double)));
    CK(h->x2.reserve(std::max<int64_t>(S, 1) * sizeof(double)));
    CK(h->y.reserve(std::max<int64_t>(S, 1) * sizeof(double)));
    CK(cudaMemcpyAsync(h->off.p, off, (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (S > 0) {
        CK(cudaMemcpyAsync(h->x1.p, x1, S * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->x2.p, x2, S * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->y.p, y, S * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    size_t tB = tm.mark();
    tm.span(&h->stats.ms_h2d, tA, tB);
    h->n_patches = P;
    h->n_in = S;
    if (!cont) {
        h->have_frames = false;
        h->have_binning = false;
    }
    h->have_binning = false;   // the point-level arrays of a compress no longer describe the stream
    h->have_rgb = false;
    int rc = run_fit(h, tm, cont);
    if (rc) return rc;
    size_t tC = tm.mark();
    tm.span(&h->stats.ms_total, tA, tC);
    rc = read_fit_stats(h);
    if (rc) return rc;
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    return GPC_OK;
}

int gpc_fit_patches(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y) {
    return fit_patches_impl(h, P, off, x1, x2, y, false);
}

int gpc_add_measurements(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y) {
    return fit_patches_impl(h, P, off, x1, x2, y, true);
}

int gpc_upload_cloud(gpc_handle* h, const void* cloud, int64_t n) {
    if (!h || n < 0 || (n > 0 && !cloud)) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    CK(h->cloud.reserve(std::max<int64_t>(n, 1) * GPC_POINT_BYTES));
    if (n > 0) CK(cudaMemcpyAsync(h->cloud.p, cloud, n * GPC_POINT_BYTES, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_in = n;
    h->have_cloud = true;
    return GPC_OK;
}

int gpc_compress_resident(gpc_handle* h) {
    if (!h) return GPC_ERR_INVALID;
    if (!h->have_cloud) return fail(h, GPC_ERR_STATE, "gpc_compress_resident before gpc_upload_cloud");
    CK(cudaSetDevice(h->cfg.device));
    reset_stats(h);
    StageTimer tm(h);
    size_t tA = tm.mark();
    int rc = compress_resident_impl(h, tm);
    if (rc) return rc;
    size_t tB = tm.mark();
    tm.span(&h->stats.ms_total, tA, tB);
    if (h->n_patches > 0 && (rc = read_fit_stats(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    return GPC_OK;
}

int gpc_compress(gpc_handle* h, const void* cloud, int64_t n) {
    if (!h || n < 0 || (n > 0 && !cloud)) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    reset_stats(h);
    StageTimer tm(h);
    size_t tA = tm.mark();
    CK(h->cloud.reserve(std::max<int64_t>(n, 1) * GPC_POINT_BYTES));
    if (n > 0) CK(cudaMemcpyAsync(h->cloud.p, cloud, n * GPC_POINT_BYTES, cudaMemcpyHostToDevice, h->stream));
    h->n_in = n;
    h->have_cloud = true;
    size_t tB = tm.mark();
    tm.span(&h->stats.ms_h2d, tA, tB);
    int rc = compress_resident_impl(h, tm);
    if (rc) return rc;
    size_t tC = tm.mark();
    tm.span(&h->stats.ms_total, tA, tC);
    if (h->n_patches > 0 && (rc = read_fit_stats(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    return GPC_OK;
}

int gpc_compress_shard_begin(gpc_handle* h, const void* cloud, int64_t n, int64_t* owned_patches, uint64_t* owned_draws) {
    if (!h || n < 0 || !owned_patches || !owned_draws) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    reset_stats(h);
    StageTimer tm(h);
    size_t tA = tm.mark();
    if (cloud) {
        CK(h->cloud.reserve(std::max<int64_t>(n, 1) * GPC_POINT_BYTES));
        if (n > 0) CK(cudaMemcpyAsync(h->cloud.p, cloud, n * GPC_POINT_BYTES, cudaMemcpyHostToDevice, h->stream));
        h->n_in = n;
        h->have_cloud = true;
    } else if (!h->have_cloud) {
        return fail(h, GPC_ERR_STATE, "gpc_compress_shard_begin without a cloud: pass one or call gpc_upload_cloud first");
    }
    size_t tB = tm.mark();
    tm.span(&h->stats.ms_h2d, tA, tB);
    int rc = run_binning(h, tm, true);
    if (rc) return rc;
    h->shard_mode = true;
    h->own_lo = h->own_hi = 0;
    h->draws_owned = 0;
    const gpc_config& c = h->cfg;
    cudaStream_t st = h->stream;
    const int64_t P = h->n_patches;
    if (P > 0) {
        CK(h->small.reserve(sizeof(SmallScratch)));
        int64_t* d_rng = h->small.as<SmallScratch>()->owned_range;
        launch_owned_range(h->code.as<uint64_t>(), P, c.leaf_order, h->small.as<SmallScratch>()->key_range, d_rng, st);
        int64_t rng[2] = {0, 0};
        CK(cudaMemcpyAsync(rng, d_rng, sizeof(rng), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        h->own_lo = rng[0]; h->own_hi = rng[1];
        const int mult = c.shuffle ? (c.rgb_rand ? 2 : 1) : 0;
        CK(h->draws.reserve((P + 1) * sizeof(int64_t)));
        CK(h->roff.reserve((P + 2) * sizeof(int64_t)));
        CK(h->scan_tmp.reserve(scan_tmp_bytes(P + 1)));
        int64_t* d_plan = h->small.as<SmallScratch>()->plan;
        launch_fit_plan(h->off.as<int64_t>(), P, mult, c.shard_rank, c.shard_count, h->own_lo, h->own_hi, h->draws.as<int64_t>(),
                        h->roff.as<int64_t>(), h->scan_tmp.p, d_plan, st);
        int64_t plan[9];
        CK(cudaMemcpyAsync(plan, d_plan, sizeof(plan), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        h->draws_owned = (uint64_t)(plan[6] - plan[5]);
    }
    *owned_patches = h->own_hi - h->own_lo;
    *owned_draws = h->draws_owned;
    size_t tC = tm.mark();
    tm.span(&h->stats.ms_total, tA, tC);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    return GPC_OK;
}

int gpc_compress_shard_finish(gpc_handle* h, int64_t patches_before, uint64_t draws_before, int64_t patches_total, uint64_t draws_total) {
    if (!h || patches_before < 0 || patches_total < patches_before) return GPC_ERR_INVALID;
    if (!h->shard_mode || !h->have_binning) return fail(h, GPC_ERR_STATE, "gpc_compress_shard_finish without gpc_compress_shard_begin");
    CK(cudaSetDevice(h->cfg.device));
    StageTimer tm(h);
    size_t tA = tm.mark();
    h->gshift = patches_before - h->own_lo;
    h->patches_total = patches_total;
    h->draws_before = draws_before;
    h->draws_total = draws_total;
    int rc;
    if (h->n_patches == 0) {
        h->patch_lo = h->patch_hi = 0;
        h->s_begin = h->s_count = 0;
        h->have_fit = true;
        h->n_bv_total = 0;
        h->params_packed = false;
        h->have_rgb = false;
        CK(h->nbv.reserve(sizeof(int32_t)));
        h->rand_offset += draws_total;
    } else if ((rc = run_fit(h, tm))) {
        return rc;
    }
    size_t tC = tm.mark();
    {   // the two phases add up in ms_total
        float before = h->stats.ms_total;
        h->stats.ms_total = 0;
        tm.span(&h->stats.ms_total, tA, tC);
        if (h->n_patches > 0 && (rc = read_fit_stats(h))) return rc;
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        tm.resolve();
        h->stats.ms_total += before;
    }
    h->stats.kernel_launches = g_launches;
    return GPC_OK;
}

int gpc_decompress_resident(gpc_handle* h, int64_t* n_out) {
    if (!h) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    reset_stats(h);
    StageTimer tm(h);
    size_t tA = tm.mark();
    int rc = run_decode(h, true, true, tm);
    if (rc) return rc;
    size_t tB = tm.mark();
    tm.span(&h->stats.ms_total, tA, tB);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    if (n_out) *n_out = h->n_decoded;
    return GPC_OK;
}

int gpc_decompress(gpc_handle* h, void* out, int64_t capacity_points, int64_t* n_out) {
    if (!h) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    reset_stats(h);
    StageTimer tm(h);
    size_t tA = tm.mark();
    int rc = run_decode(h, true, false, tm);
    if (rc) return rc;
    size_t tB = tm.mark();
    if (out) {
        if (capacity_points < h->n_decoded) return fail(h, GPC_ERR_INVALID, "output buffer too small");
        if (h->n_decoded > 0)
            CK(cudaMemcpyAsync(out, h->out32.p, h->n_decoded * GPC_POINT_BYTES, cudaMemcpyDeviceToHost, h->stream));
    }
    size_t tC = tm.mark();
    tm.span(&h->stats.ms_d2h, tB, tC);
    tm.span(&h->stats.ms_total, tA, tC);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    tm.resolve();
    h->stats.kernel_launches = g_launches;
    if (n_out) *n_out = h->n_decoded;
    return GPC_OK;
}

int gpc_get_heights(gpc_handle* h, double* out, int64_t capacity) {
    if (!h || !out) return GPC_ERR_INVALID;
    if (!h->heights_valid) return fail(h, GPC_ERR_STATE, "heights are produced by gpc_decompress_resident: call it before gpc_get_heights");
    if (capacity < h->n_decoded) return fail(h, GPC_ERR_INVALID, "heights buffer too small");
    CK(cudaSetDevice(h->cfg.device));
    if (h->n_decoded > 0) CK(cudaMemcpy(out, h->heights.p, h->n_decoded * sizeof(double), cudaMemcpyDeviceToHost));
    return GPC_OK;
}

static int evaluate_impl(gpc_handle* h, int64_t op0, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y,
                         int conf, double* f, double* sigma, double* lik, double* dX, int dout = 1);

int gpc_predict(gpc_handle* h, int64_t patch, const double* X, int64_t m, double* f, double* sigma) {
    if (!h || m < 0 || (m > 0 && (!X || !f))) return GPC_ERR_INVALID;
    if (!h->have_fit) return fail(h, GPC_ERR_STATE, "predict before fit");
    if (h->shard_mode) patch -= h->gshift;  // global -> local index of the sharded binning
    if (patch < h->patch_lo || patch >= h->patch_hi) return fail(h, GPC_ERR_INVALID, "patch outside this shard");
    if (sigma && !h->cfg.keep_state) return fail(h, GPC_ERR_INVALID, "sigma needs gpc_config.keep_state");
    CK(cudaSetDevice(h->cfg.device));
    if (m == 0) return GPC_OK;
    const gpc_config& c = h->cfg;
    const int64_t op = patch - h->patch_lo;
    if (sigma) {  // mean and sigma from the batched evaluation kernel (K9)
        std::vector<double> x1(m), x2(m);
        for (int64_t i = 0; i < m; i++) { x1[i] = X[2 * i]; x2[i] = X[2 * i + 1]; }
        const int64_t off[2] = {0, m};
        return evaluate_impl(h, op, 1, off, x1.data(), x2.data(), nullptr, 0, f, sigma, nullptr, nullptr);
    }
    int32_t N = 0;
    CK(cudaMemcpy(&N, h->nbv.as<int32_t>() + op, sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(h->tmpA.reserve(2 * m * sizeof(double)));
    CK(h->tmpB.reserve(m * sizeof(double)));
    CK(cudaMemcpyAsync(h->tmpA.p, X, 2 * m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_predict_points(h->alpha.as<double>() + op * c.capacity, h->b1.as<double>() + op * c.capacity,
                          h->b2.as<double>() + op * c.capacity, N, c.sigmaf_sq, kernel_cl(c), h->tmpA.as<double>(), m,
                          h->tmpB.as<double>(), h->stream);
    CK(cudaMemcpyAsync(f, h->tmpB.p, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return GPC_OK;
}

// patches [op0, op0 + P) of this shard (local indices)
static int evaluate_impl(gpc_handle* h, int64_t op0, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y,
                         int conf, double* f, double* sigma, double* lik, double* dX, int dout) {
    if (!h || P < 0 || (P > 0 && !off)) return GPC_ERR_INVALID;
    const bool need_c = sigma || lik || dX;   // the mean alone needs no state
    if (!h->have_fit) return fail(h, GPC_ERR_STATE, "evaluation before compress / fit");
    if (need_c && (!h->cfg.keep_state || !h->have_state || !h->dumpC.p))
        return fail(h, GPC_ERR_STATE, "sigma / likelihood evaluation needs a fit made with gpc_config.keep_state");
    if (dout == 3 && (!h->have_rgb || (need_c && !h->r_dumpC.p)))
        return fail(h, GPC_ERR_STATE, "gpc_evaluate_patches_rgb needs a compress with gpc_config.rgb = 1 (and keep_state = 1 for sigma / likelihood)");
    if (op0 < 0 || op0 + P > h->patch_hi - h->patch_lo) return fail(h, GPC_ERR_INVALID, "more patches than this shard holds");
    if (P == 0) return GPC_OK;
    const int64_t m = off[P];
    if (off[0] != 0 || m < 0) return fail(h, GPC_ERR_INVALID, "offsets must start at 0 and be non-decreasing");
    for (int64_t p = 0; p < P; p++)
        if (off[p + 1] < off[p]) return fail(h, GPC_ERR_INVALID, "offsets must start at 0 and be non-decreasing");
    if (m > 0 && (!x1 || !x2 || ((lik || dX) && !y))) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (m == 0) return GPC_OK;
    const gpc_config& c = h->cfg;
    cudaStream_t st = h->stream;
    const size_t in_bytes = (size_t)(P + 1) * sizeof(int64_t) + (2 + (size_t)dout) * (size_t)m * sizeof(double);
    CK(h->ev_in.reserve(in_bytes));
    CK(h->ev_out.reserve((5 + (size_t)dout) * (size_t)m * sizeof(double)));
    int64_t* d_off = h->ev_in.as<int64_t>();
    double* d_x1 = reinterpret_cast<double*>(d_off + P + 1);
    double* d_x2 = d_x1 + m;
    double* d_y = d_x2 + m;
    double* d_f = h->ev_out.as<double>();
    double* d_sg = d_f + (size_t)dout * m;
    double* d_lk = d_sg + m;
    double* d_dx = d_lk + m;
    CK(cudaMemcpyAsync(d_off, off, (size_t)(P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_x1, x1, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_x2, x2, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
    if (y) CK(cudaMemcpyAsync(d_y, y, (size_t)dout * m * sizeof(double), cudaMemcpyHostToDevice, st));
    // largest BV count (sizes the shared-memory tiles)
    CK(h->small.reserve(sizeof(SmallScratch)));
    CK(h->nonempty.reserve((P + 1) * sizeof(int64_t)));
    int32_t* d_maxes = h->small.as<SmallScratch>()->maxes;
    const int32_t* nbv_d = (dout == 3 ? h->r_nbv.as<int32_t>() : h->nbv.as<int32_t>()) + op0;
    launch_flag_nonempty(nbv_d, nullptr, P, h->nonempty.as<int64_t>(), d_maxes, st);
    int32_t maxes[2] = {0, 0};
    CK(cudaMemcpyAsync(maxes, d_maxes, sizeof(maxes), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    EvalArgs a;
    a.n_patches = P;
    a.nbv = nbv_d;
    a.off = d_off;
    a.stride = c.capacity;
    a.nmax = maxes[0];
    a.dout = dout;
    const int64_t oo = op0 * c.capacity, oc = op0 * (int64_t)c.capacity * c.capacity;
    if (dout == 3) {
        a.alpha[0] = h->r_alpha0.as<double>() + oo; a.alpha[1] = h->r_alpha1.as<double>() + oo; a.alpha[2] = h->r_alpha2.as<double>() + oo;
        a.b1 = h->r_b1.as<double>() + oo; a.b2 = h->r_b2.as<double>() + oo; a.C = need_c ? h->r_dumpC.as<double>() + oc : nullptr;
        a.s20 = c.rgb_s0;
    } else {
        a.alpha[0] = h->alpha.as<double>() + oo; a.alpha[1] = a.alpha[2] = nullptr;
        a.b1 = h->b1.as<double>() + oo; a.b2 = h->b2.as<double>() + oo; a.C = need_c ? h->dumpC.as<double>() + oc : nullptr;
        a.s20 = c.s0;
    }
    a.x1 = d_x1; a.x2 = d_x2; a.y = y ? d_y : nullptr;
    a.p0 = c.sigmaf_sq; a.cl = kernel_cl(c); a.c1 = (-c.sigmaf_sq) / c.l_sq;
    a.pow2pi3 = std::pow(2.0 * M_PI, 3.0);   // pow(2.0f*M_PI, double(y.rows())), sparse_gp_field.hpp:350
    a.conf = conf;
    a.f = f ? d_f : nullptr; a.sigma = sigma ? d_sg : nullptr; a.lik = lik ? d_lk : nullptr; a.dX = dX ? d_dx : nullptr;
    StageTimer tm(h);
    h->stats.ms_evaluate = 0;
    size_t t0 = tm.mark();
    if (launch_evaluate(a, st) != cudaSuccess) return fail(h, GPC_ERR_CUDA, "evaluate kernel launch failed");
    size_t t1 = tm.mark();
    tm.span(&h->stats.ms_evaluate, t0, t1);
    if (f) CK(cudaMemcpyAsync(f, d_f, (size_t)dout * m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sigma) CK(cudaMemcpyAsync(sigma, d_sg, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (lik) CK(cudaMemcpyAsync(lik, d_lk, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dX) CK(cudaMemcpyAsync(dX, d_dx, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    tm.resolve();
    return GPC_OK;
}

int gpc_evaluate_patches(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y,
                         int conf, double* f, double* sigma, double* lik, double* dX) {
    return evaluate_impl(h, 0, P, off, x1, x2, y, conf, f, sigma, lik, dX);
}

int gpc_evaluate_patches_rgb(gpc_handle* h, int64_t P, const int64_t* off, const double* x1, const double* x2, const double* y3,
                             int conf, double* f3, double* sigma, double* lik, double* dX) {
    return evaluate_impl(h, 0, P, off, x1, x2, y3, conf, f3, sigma, lik, dX, 3);
}

int gpc_get_sizes(gpc_handle* h, gpc_sizes* s) {
    if (!h || !s) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (h->have_fit && !h->params_packed) {
        int rc = gpc_get_params(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    s->n_in = h->n_in; s->n_patches = h->shard_mode ? h->patches_total : h->n_patches; s->n_claimed = h->n_claimed;
    s->n_bv_total = h->have_fit ? h->n_bv_total : 0;
    const int64_t gs = h->shard_mode ? h->gshift : 0;
    s->patch_lo = h->patch_lo + gs; s->patch_hi = h->patch_hi + gs; s->n_decoded = h->n_decoded;
    s->rand_offset = h->rand_offset;
    for (int a = 0; a < 3; a++) s->lattice_min[a] = h->lattice_min[a];
    s->depth = h->depth; s->pad = 0;
    return GPC_OK;
}

int gpc_get_stats(gpc_handle* h, gpc_stats* s) {
    if (!h || !s) return GPC_ERR_INVALID;
    *s = h->stats;
    return GPC_OK;
}

int gpc_get_params(gpc_handle* h, int32_t* nbv, int64_t* bv_off, int32_t* bv_index, double* bv1, double* bv2, double* alpha,
                   int32_t* flags) {
    if (!h) return GPC_ERR_INVALID;
    if (!h->have_fit) return fail(h, GPC_ERR_STATE, "no fit held by the handle");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const int64_t PL = h->patch_hi - h->patch_lo;
    const int cap = h->cfg.capacity;
    if (!h->params_packed) {  // parameters installed by gpc_set_params: pack them now
        const int64_t PLa = std::max<int64_t>(PL, 1);
        CK(h->bv_off.reserve((PLa + 1) * sizeof(int64_t)));
        CK(h->nonempty.reserve((PLa + 1) * sizeof(int64_t)));
        CK(h->scan_tmp.reserve(scan_tmp_bytes(PLa + 1)));
        CK(h->palpha.reserve(PLa * cap * sizeof(double)));
        CK(h->pb1.reserve(PLa * cap * sizeof(double)));
        CK(h->pb2.reserve(PLa * cap * sizeof(double)));
        CK(h->pidx.reserve(PLa * cap * sizeof(int32_t)));
        CK(h->bidx.reserve(PLa * cap * sizeof(int32_t)));
        launch_compact_params(h->nbv.as<int32_t>(), PL, cap, h->alpha.as<double>(), h->b1.as<double>(), h->b2.as<double>(),
                              h->bidx.as<int32_t>(), h->nonempty.as<int64_t>(), h->bv_off.as<int64_t>(), h->scan_tmp.p,
                              h->palpha.as<double>(), h->pb1.as<double>(), h->pb2.as<double>(), h->pidx.as<int32_t>(), st);
        int64_t tot = 0;
        CK(cudaMemcpyAsync(&tot, h->bv_off.as<int64_t>() + PL, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        h->n_bv_total = tot;
        h->params_packed = true;
    }
    const int64_t T = h->n_bv_total;
    if (nbv && PL > 0) CK(cudaMemcpyAsync(nbv, h->nbv.p, PL * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (flags && PL > 0) CK(cudaMemcpyAsync(flags, h->flags.p, PL * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (bv_off) CK(cudaMemcpyAsync(bv_off, h->bv_off.p, (PL + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (bv_index && T > 0) CK(cudaMemcpyAsync(bv_index, h->pidx.p, T * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (bv1 && T > 0) CK(cudaMemcpyAsync(bv1, h->pb1.p, T * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (bv2 && T > 0) CK(cudaMemcpyAsync(bv2, h->pb2.p, T * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (alpha && T > 0) CK(cudaMemcpyAsync(alpha, h->palpha.p, T * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GPC_OK;
}

int gpc_get_params_rgb(gpc_handle* h, int32_t* nbv, int64_t* bv_off, int32_t* bv_index, double* bv1, double* bv2, double* alpha3,
                       int32_t* perm, int64_t* n_bv_total_rgb) {
    if (!h) return GPC_ERR_INVALID;
    if (!h->have_rgb) return fail(h, GPC_ERR_STATE, "no RGB field GP held by the handle (gpc_config.rgb = 1 and a compress are needed)");
    CK(cudaSetDevice(h->cfg.device));
    const int64_t PL = h->patch_hi - h->patch_lo;
    const int cap = h->cfg.capacity;
    std::vector<int32_t> nb(PL);
    if (PL > 0) CK(cudaMemcpy(nb.data(), h->r_nbv.p, PL * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<int64_t> bo(PL + 1, 0);
    for (int64_t p = 0; p < PL; p++) bo[p + 1] = bo[p] + nb[p];
    if (n_bv_total_rgb) *n_bv_total_rgb = bo[PL];
    if (nbv) std::memcpy(nbv, nb.data(), PL * sizeof(int32_t));
    if (bv_off) std::memcpy(bv_off, bo.data(), (PL + 1) * sizeof(int64_t));
    // results of a next-row feature: a plain strided copy + host compaction is enough here
    auto fetch = [&](const DevBuf& src, size_t esz, std::vector<uint8_t>& tmp) -> int {
        tmp.resize((size_t)std::max<int64_t>(PL, 1) * cap * esz);
        if (PL > 0) CK(cudaMemcpy(tmp.data(), src.p, (size_t)PL * cap * esz, cudaMemcpyDeviceToHost));
        return GPC_OK;
    };
    std::vector<uint8_t> t0, t1, t2;
    int rc;
    if (bv_index) {
        if ((rc = fetch(h->r_bidx, 4, t0))) return rc;
        for (int64_t p = 0; p < PL; p++) std::memcpy(bv_index + bo[p], t0.data() + (size_t)p * cap * 4, (size_t)nb[p] * 4);
    }
    if (bv1) {
        if ((rc = fetch(h->r_b1, 8, t0))) return rc;
        for (int64_t p = 0; p < PL; p++) std::memcpy(bv1 + bo[p], t0.data() + (size_t)p * cap * 8, (size_t)nb[p] * 8);
    }
    if (bv2) {
        if ((rc = fetch(h->r_b2, 8, t0))) return rc;
        for (int64_t p = 0; p < PL; p++) std::memcpy(bv2 + bo[p], t0.data() + (size_t)p * cap * 8, (size_t)nb[p] * 8);
    }
    if (alpha3) {
        if ((rc = fetch(h->r_alpha0, 8, t0)) || (rc = fetch(h->r_alpha1, 8, t1)) || (rc = fetch(h->r_alpha2, 8, t2))) return rc;
        const double *a0 = (const double*)t0.data(), *a1 = (const double*)t1.data(), *a2 = (const double*)t2.data();
        for (int64_t p = 0; p < PL; p++)
            for (int i = 0; i < nb[p]; i++) {
                alpha3[3 * (bo[p] + i) + 0] = a0[p * cap + i];
                alpha3[3 * (bo[p] + i) + 1] = a1[p * cap + i];
                alpha3[3 * (bo[p] + i) + 2] = a2[p * cap + i];
            }
    }
    if (perm && h->s_count) {
        std::memset(perm, 0xff, h->n_claimed * sizeof(int32_t));
        CK(cudaMemcpy(perm + h->s_begin, h->perm_rgb.as<int32_t>() + h->s_begin, h->s_count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    return GPC_OK;
}

int gpc_get_state(gpc_handle* h, int64_t patch, double* C, double* Q) {
    if (!h) return GPC_ERR_INVALID;
    if (!h->have_fit || !h->cfg.keep_state || !h->have_state) return fail(h, GPC_ERR_STATE, "gpc_get_state needs a fit made with keep_state");
    if (h->shard_mode) patch -= h->gshift;
    if (patch < h->patch_lo || patch >= h->patch_hi) return fail(h, GPC_ERR_INVALID, "patch outside this shard");
    CK(cudaSetDevice(h->cfg.device));
    const int64_t op = patch - h->patch_lo;
    const int cap = h->cfg.capacity;
    int32_t N = 0;
    CK(cudaMemcpy(&N, h->nbv.as<int32_t>() + op, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (C && N > 0) CK(cudaMemcpy(C, h->dumpC.as<double>() + op * (int64_t)cap * cap, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToHost));
    if (Q && N > 0) CK(cudaMemcpy(Q, h->dumpQ.as<double>() + op * (int64_t)cap * cap, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToHost));
    return GPC_OK;
}

int gpc_set_params(gpc_handle* h, int64_t P, const int32_t* nbv, const double* bv1, const double* bv2, const double* alpha,
                   const double* quat4, const double* mean3, const double* rgbmean3) {
    if (!h || P < 0 || !nbv || !bv1 || !bv2 || !alpha) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    h->shard_mode = false;
    const int cap = h->cfg.capacity;
    std::vector<int64_t> off(P + 1, 0);
    for (int64_t p = 0; p < P; p++) {
        if (nbv[p] < 0 || nbv[p] > cap) return fail(h, GPC_ERR_INVALID, "nbv exceeds capacity");
        off[p + 1] = off[p] + 1;  // shard by patch count
    }
    int64_t lo, hi;
    shard_range(off, h->cfg.shard_rank, h->cfg.shard_count, &lo, &hi);
    const int64_t PL = hi - lo, PLa = std::max<int64_t>(PL, 1);
    std::vector<int64_t> bo(P + 1, 0);
    for (int64_t p = 0; p < P; p++) bo[p + 1] = bo[p] + nbv[p];
    CK(h->nbv.reserve(PLa * sizeof(int32_t)));
    CK(h->flags.reserve(PLa * sizeof(int32_t)));
    CK(h->alpha.reserve(PLa * cap * sizeof(double)));
    CK(h->b1.reserve(PLa * cap * sizeof(double)));
    CK(h->b2.reserve(PLa * cap * sizeof(double)));
    auto expand = [&](const double* src, DevBuf& dst) -> int {
        std::vector<double> tmp((size_t)PLa * cap, 0.0);
        for (int64_t p = lo; p < hi; p++) std::memcpy(&tmp[(size_t)(p - lo) * cap], src + bo[p], (size_t)nbv[p] * sizeof(double));
        CK(cudaMemcpy(dst.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
        return GPC_OK;
    };
    int rc;
    if ((rc = expand(alpha, h->alpha))) return rc;
    if ((rc = expand(bv1, h->b1))) return rc;
    if ((rc = expand(bv2, h->b2))) return rc;
    if (PL > 0) CK(cudaMemcpy(h->nbv.p, nbv + lo, PL * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemset(h->flags.p, 0, PLa * sizeof(int32_t)));
    h->have_frames = false;
    if (quat4 && mean3 && rgbmean3 && P > 0) {
        CK(h->quat.reserve(P * 4 * sizeof(double)));
        CK(h->mean.reserve(P * 3 * sizeof(double)));
        CK(h->rgbmean.reserve(P * 3 * sizeof(double)));
        CK(cudaMemcpy(h->quat.p, quat4, P * 4 * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->mean.p, mean3, P * 3 * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->rgbmean.p, rgbmean3, P * 3 * sizeof(double), cudaMemcpyHostToDevice));
        h->have_frames = true;
    }
    h->n_patches = P;
    h->patch_lo = lo;
    h->patch_hi = hi;
    h->have_fit = true;
    h->have_binning = false;
    h->have_rgb = false;
    h->n_bv_total = -1;
    h->params_packed = false;
    // nothing of an earlier fit survives: no claimed stream, no kept C / Q, no fed counts, no BV indices
    h->n_claimed = 0; h->s_begin = 0; h->s_count = 0;
    h->have_state = false;
    h->heights_valid = false;
    CK(h->bidx.reserve((size_t)PLa * cap * sizeof(int32_t)));
    CK(cudaMemset(h->bidx.p, 0xff, (size_t)PLa * cap * sizeof(int32_t)));
    return GPC_OK;
}

int gpc_get_patches(gpc_handle* h, uint64_t* code, float* center3, int32_t* n_candidates, double* R9, double* quat4,
                    double* mean3, double* rgbmean3, int64_t* patch_off) {
    if (!h) return GPC_ERR_INVALID;
    if (h->shard_mode) return fail(h, GPC_ERR_STATE, "patch-level arrays are local to the shard after gpc_compress_shard_*: use gpc_get_params / gpc_decompress");
    CK(cudaSetDevice(h->cfg.device));
    const int64_t P = h->n_patches;
    if (patch_off) {
        if (!h->have_fit && !h->have_binning) return fail(h, GPC_ERR_STATE, "no patches held by the handle");
        CK(cudaMemcpy(patch_off, h->off.p, (P + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    if (code || center3 || n_candidates || R9) {
        if (!h->have_binning) return fail(h, GPC_ERR_STATE, "patch frames need a compress on this handle");
        if (code && P) CK(cudaMemcpy(code, h->code.p, P * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (center3 && P) CK(cudaMemcpy(center3, h->center.p, P * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        if (n_candidates && P) CK(cudaMemcpy(n_candidates, h->ncand.p, P * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (R9 && P) CK(cudaMemcpy(R9, h->Rm.p, P * 9 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (quat4 || mean3 || rgbmean3) {
        if (!h->have_frames) return fail(h, GPC_ERR_STATE, "no patch frames held by the handle");
        if (quat4 && P) CK(cudaMemcpy(quat4, h->quat.p, P * 4 * sizeof(double), cudaMemcpyDeviceToHost));
        if (mean3 && P) CK(cudaMemcpy(mean3, h->mean.p, P * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        if (rgbmean3 && P) CK(cudaMemcpy(rgbmean3, h->rgbmean.p, P * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return GPC_OK;
}

int gpc_get_assignment(gpc_handle* h, int32_t* owner, int32_t* stream_index, double* x1, double* x2, double* y, int32_t* perm) {
    if (!h) return GPC_ERR_INVALID;
    if (h->shard_mode) return fail(h, GPC_ERR_STATE, "point-level arrays are local to the shard after gpc_compress_shard_*");
    CK(cudaSetDevice(h->cfg.device));
    const int64_t S = h->n_claimed;
    if (owner || stream_index) {
        if (!h->have_binning) return fail(h, GPC_ERR_STATE, "assignment needs a compress on this handle");
        if (owner && h->n_in) CK(cudaMemcpy(owner, h->owner.p, h->n_in * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (stream_index && S) CK(cudaMemcpy(stream_index, h->st_idx.p, S * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    if (!h->have_fit && (x1 || x2 || y || perm)) return fail(h, GPC_ERR_STATE, "no stream held by the handle");
    if (x1 && S) CK(cudaMemcpy(x1, h->x1.p, S * sizeof(double), cudaMemcpyDeviceToHost));
    if (x2 && S) CK(cudaMemcpy(x2, h->x2.p, S * sizeof(double), cudaMemcpyDeviceToHost));
    if (y && S) CK(cudaMemcpy(y, h->y.p, S * sizeof(double), cudaMemcpyDeviceToHost));
    if (perm && h->s_count) {
        // only this shard's range of the permutation is defined
        std::memset(perm, 0xff, S * sizeof(int32_t));
        CK(cudaMemcpy(perm + h->s_begin, h->perm.as<int32_t>() + h->s_begin, h->s_count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    return GPC_OK;
}

int gpc_debug_exp(gpc_handle* h, const double* x, double* out, int64_t n) {
    if (!h || n < 0 || (n > 0 && (!x || !out))) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (n == 0) return GPC_OK;
    CK(h->tmpA.reserve(n * sizeof(double)));
    CK(h->tmpB.reserve(n * sizeof(double)));
    CK(cudaMemcpyAsync(h->tmpA.p, x, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_debug_exp(h->tmpA.as<double>(), h->tmpB.as<double>(), n, h->stream);
    CK(cudaMemcpyAsync(out, h->tmpB.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return GPC_OK;
}

int gpc_save(gpc_handle* h, const char* path, int64_t* bytes_written) {
    if (!h || !path) return GPC_ERR_INVALID;
    if (!h->have_fit) return fail(h, GPC_ERR_STATE, "gpc_save before compress / fit");
    if (!h->have_frames) return fail(h, GPC_ERR_STATE, "gpc_save needs patch frames (a compress on this handle)");
    CK(cudaSetDevice(h->cfg.device));
    const int64_t PL = h->patch_hi - h->patch_lo;
    int rc = gpc_get_params(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);  // packs the parameters if needed
    if (rc) return rc;
    const int64_t T = h->n_bv_total;
    std::vector<int32_t> nbv(PL);
    std::vector<double> quat(PL * 4), mean(PL * 3), rgbm(PL * 3), b1(T), b2(T), al(T);
    if (PL > 0) {
        CK(cudaMemcpy(nbv.data(), h->nbv.p, PL * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(quat.data(), h->quat.as<double>() + 4 * h->patch_lo, PL * 4 * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(mean.data(), h->mean.as<double>() + 3 * h->patch_lo, PL * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rgbm.data(), h->rgbmean.as<double>() + 3 * h->patch_lo, PL * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (T > 0) {
        CK(cudaMemcpy(b1.data(), h->pb1.p, T * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b2.data(), h->pb2.p, T * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(al.data(), h->palpha.p, T * sizeof(double), cudaMemcpyDeviceToHost));
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(h, GPC_ERR_INVALID, std::string("cannot open ") + path);
    const char magic[8] = {'G', 'P', 'C', 'B', '2', '0', '0', 0};
    const uint32_t version = 2;
    const gpc_config& c = h->cfg;
    // RGB field GP block (version 2): strided device arrays -> packed host arrays
    const int32_t has_rgb = h->have_rgb ? 1 : 0;
    std::vector<int32_t> rnbv(has_rgb ? PL : 0);
    std::vector<double> rb1, rb2, ra;
    int64_t TR = 0;
    if (has_rgb && PL > 0) {
        const int cap = c.capacity;
        CK(cudaMemcpy(rnbv.data(), h->r_nbv.p, PL * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int32_t v : rnbv) TR += v;
        std::vector<double> t((size_t)PL * cap);
        auto pack = [&](const DevBuf& src, std::vector<double>& dst, int stride3, int ch) -> int {
            CK(cudaMemcpy(t.data(), src.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
            int64_t o = 0;
            for (int64_t p = 0; p < PL; p++)
                for (int i = 0; i < rnbv[p]; i++, o++) dst[(size_t)o * stride3 + ch] = t[(size_t)p * cap + i];
            return GPC_OK;
        };
        rb1.resize(TR); rb2.resize(TR); ra.resize(3 * TR);
        int rc2;
        if ((rc2 = pack(h->r_b1, rb1, 1, 0)) || (rc2 = pack(h->r_b2, rb2, 1, 0)) || (rc2 = pack(h->r_alpha0, ra, 3, 0)) ||
            (rc2 = pack(h->r_alpha1, ra, 3, 1)) || (rc2 = pack(h->r_alpha2, ra, 3, 2)))
            return rc2;
    }
    const double dcfg[5] = {c.res, c.s0, c.eps_tol, c.sigmaf_sq, c.l_sq};
    const int32_t icfg[2] = {c.sz, c.capacity};
    size_t ok = 1;
    ok &= std::fwrite(magic, 8, 1, f) == 1;
    ok &= std::fwrite(&version, 4, 1, f) == 1;
    ok &= std::fwrite(dcfg, sizeof(dcfg), 1, f) == 1;
    ok &= std::fwrite(icfg, sizeof(icfg), 1, f) == 1;
    ok &= std::fwrite(&PL, 8, 1, f) == 1;
    ok &= std::fwrite(&T, 8, 1, f) == 1;
    auto put = [&](const void* p, size_t n) { if (n) ok &= std::fwrite(p, n, 1, f) == 1; };
    put(nbv.data(), PL * sizeof(int32_t));
    put(quat.data(), PL * 4 * sizeof(double));
    put(mean.data(), PL * 3 * sizeof(double));
    put(rgbm.data(), PL * 3 * sizeof(double));
    put(b1.data(), T * sizeof(double));
    put(b2.data(), T * sizeof(double));
    put(al.data(), T * sizeof(double));
    const double rcfg[2] = {c.rgb_s0, c.rgb_eps_tol};
    put(&has_rgb, sizeof(has_rgb));
    put(rcfg, sizeof(rcfg));
    put(&TR, sizeof(TR));
    put(rnbv.data(), rnbv.size() * sizeof(int32_t));
    put(rb1.data(), rb1.size() * sizeof(double));
    put(rb2.data(), rb2.size() * sizeof(double));
    put(ra.data(), ra.size() * sizeof(double));
    const long pos = std::ftell(f);
    std::fclose(f);
    if (!ok) return fail(h, GPC_ERR_INVALID, std::string("short write to ") + path);
    if (bytes_written) *bytes_written = (int64_t)pos;
    return GPC_OK;
}

int gpc_get_config(const gpc_handle* h, gpc_config* cfg) {
    if (!h || !cfg) return GPC_ERR_INVALID;
    *cfg = h->cfg;
    return GPC_OK;
}

static int gpc_load_impl(gpc_handle* h, const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(h, GPC_ERR_INVALID, std::string("cannot open ") + path);
    struct Closer { FILE* f; ~Closer() { if (f) std::fclose(f); } } closer{f};
    std::fseek(f, 0, SEEK_END);
    const int64_t fsize = (int64_t)std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    char magic[8];
    uint32_t version = 0;
    double dcfg[5];
    int32_t icfg[2];
    int64_t PL = 0, T = 0;
    bool ok = std::fread(magic, 8, 1, f) == 1 && std::memcmp(magic, "GPCB200", 8) == 0;
    ok = ok && std::fread(&version, 4, 1, f) == 1 && (version == 1 || version == 2);
    ok = ok && std::fread(dcfg, sizeof(dcfg), 1, f) == 1 && std::fread(icfg, sizeof(icfg), 1, f) == 1;
    ok = ok && std::fread(&PL, 8, 1, f) == 1 && std::fread(&T, 8, 1, f) == 1 && PL >= 0 && T >= 0;
    if (!ok) return fail(h, GPC_ERR_INVALID, "not a gpc_b200 parameter file (or unsupported version)");
    // nothing of the handle is touched before the whole file has been read and checked
    const double f_res = dcfg[0], f_s0 = dcfg[1], f_eps = dcfg[2], f_p0 = dcfg[3], f_lsq = dcfg[4];
    const int32_t f_sz = icfg[0], f_cap = icfg[1];
    if (!(f_res > 0) || !(f_lsq > 0) || f_sz < 1 || f_cap < 1 || f_cap > 201 || !(f_s0 == f_s0) || !(f_p0 == f_p0))
        return fail(h, GPC_ERR_INVALID, "parameter file holds an invalid configuration (res, l_sq > 0, sz >= 1, 1 <= capacity <= 201)");
    // sizes against the file length: 4 PL + 80 PL + 24 T bytes follow the header
    if (PL > fsize / 84 + 1 || T > fsize / 24 + 1 || T > PL * (int64_t)f_cap)
        return fail(h, GPC_ERR_INVALID, "parameter file: patch / basis-vector counts exceed the file size");
    std::vector<int32_t> nbv(PL);
    std::vector<double> quat(PL * 4), mean(PL * 3), rgbm(PL * 3), b1(T), b2(T), al(T);
    auto get = [&](void* p, size_t n) { if (n) ok = ok && std::fread(p, n, 1, f) == 1; };
    get(nbv.data(), PL * sizeof(int32_t));
    get(quat.data(), PL * 4 * sizeof(double));
    get(mean.data(), PL * 3 * sizeof(double));
    get(rgbm.data(), PL * 3 * sizeof(double));
    get(b1.data(), T * sizeof(double));
    get(b2.data(), T * sizeof(double));
    get(al.data(), T * sizeof(double));
    int32_t has_rgb = 0;
    double rcfg[2] = {h->cfg.rgb_s0, h->cfg.rgb_eps_tol};
    int64_t TR = 0;
    std::vector<int32_t> rnbv;
    std::vector<double> rb1, rb2, ra;
    if (ok && version >= 2) {
        get(&has_rgb, sizeof(has_rgb));
        get(rcfg, sizeof(rcfg));
        get(&TR, sizeof(TR));
        if (ok && has_rgb) {
            if (TR < 0 || TR > fsize / 40 + 1 || TR > PL * (int64_t)f_cap)
                return fail(h, GPC_ERR_INVALID, "inconsistent RGB block in the parameter file");
            rnbv.resize(PL); rb1.resize(TR); rb2.resize(TR); ra.resize(3 * TR);
            get(rnbv.data(), PL * sizeof(int32_t));
            get(rb1.data(), TR * sizeof(double));
            get(rb2.data(), TR * sizeof(double));
            get(ra.data(), 3 * TR * sizeof(double));
        }
    }
    if (!ok) return fail(h, GPC_ERR_INVALID, "truncated parameter file");
    int64_t sum = 0;
    for (int32_t v : nbv) {
        if (v < 0 || v > f_cap) return fail(h, GPC_ERR_INVALID, "inconsistent parameter file");
        sum += v;
    }
    if (sum != T) return fail(h, GPC_ERR_INVALID, "inconsistent parameter file");
    std::vector<int64_t> ro(PL + 1, 0);
    if (has_rgb) {
        for (int64_t p = 0; p < PL; p++) {
            if (rnbv[p] < 0 || rnbv[p] > f_cap) return fail(h, GPC_ERR_INVALID, "inconsistent RGB block in the parameter file");
            ro[p + 1] = ro[p] + rnbv[p];
        }
        if (ro[PL] != TR) return fail(h, GPC_ERR_INVALID, "inconsistent RGB block in the parameter file");
    }
    // the decoder's configuration comes from the file (the shard layout and device stay the handle's own); committed only
    // when the parameters are installed
    const gpc_config saved = h->cfg;
    h->cfg.res = f_res; h->cfg.s0 = f_s0; h->cfg.eps_tol = f_eps; h->cfg.sigmaf_sq = f_p0; h->cfg.l_sq = f_lsq;
    h->cfg.sz = f_sz; h->cfg.capacity = f_cap;
    h->cfg.rgb_s0 = rcfg[0]; h->cfg.rgb_eps_tol = rcfg[1];
    h->cfg.rgb = has_rgb;
    h->have_fit = false;  // whatever the handle held is gone from here on: its arrays are strided by the old capacity
    const double one = 0.0;
    int rc = gpc_set_params(h, PL, nbv.data(), T ? b1.data() : &one, T ? b2.data() : &one, T ? al.data() : &one, quat.data(), mean.data(),
                            rgbm.data());
    if (rc) { h->cfg = saved; h->have_fit = false; return rc; }
    if (!has_rgb) return GPC_OK;
    // install the RGB field GP of this shard's patch range (strided by capacity, like the fitted arrays)
    const int cap = h->cfg.capacity;
    const int64_t lo = h->patch_lo, hi = h->patch_hi, PLs = hi - lo, PLa = std::max<int64_t>(PLs, 1);
    DevBuf* dst[5] = {&h->r_b1, &h->r_b2, &h->r_alpha0, &h->r_alpha1, &h->r_alpha2};
    std::vector<double> t((size_t)PLa * cap);
    for (int k = 0; k < 5; k++) {
        std::fill(t.begin(), t.end(), 0.0);
        for (int64_t p = lo; p < hi; p++)
            for (int i = 0; i < rnbv[p]; i++)
                t[(size_t)(p - lo) * cap + i] = (k == 0) ? rb1[ro[p] + i] : (k == 1) ? rb2[ro[p] + i] : ra[3 * (ro[p] + i) + (k - 2)];
        CK(dst[k]->reserve(t.size() * sizeof(double)));
        CK(cudaMemcpy(dst[k]->p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CK(h->r_nbv.reserve(PLa * sizeof(int32_t)));
    if (PLs > 0) CK(cudaMemcpy(h->r_nbv.p, rnbv.data() + lo, PLs * sizeof(int32_t), cudaMemcpyHostToDevice));
    // a loaded file has no BV indices: gpc_get_params_rgb(bv_index) reports -1
    CK(h->r_bidx.reserve((size_t)PLa * cap * sizeof(int32_t)));
    CK(cudaMemset(h->r_bidx.p, 0xff, (size_t)PLa * cap * sizeof(int32_t)));
    h->have_rgb = true;
    return GPC_OK;
}

int gpc_load(gpc_handle* h, const char* path) {
    if (!h || !path) return GPC_ERR_INVALID;
    try {
        return gpc_load_impl(h, path);
    } catch (const std::exception& e) {  // e.g. bad_alloc on a hostile header: nothing crosses the C ABI
        h->have_fit = false;
        return fail(h, GPC_ERR_INVALID, std::string("gpc_load: ") + e.what());
    }
}

int gpc_debug_peak(gpc_handle* h, int kind, double* value) {
    if (!h || !value || kind < 0 || kind > 2) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    h->shard_mode = false;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    CK(measure_peak(kind, sms, h->stream, value));
    return GPC_OK;
}

int gpc_shard_range(const int64_t* off, int64_t n_patches, int32_t shard_rank, int32_t shard_count, int64_t* lo, int64_t* hi) {
    if (!off || n_patches < 0 || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count || !lo || !hi) return GPC_ERR_INVALID;
    std::vector<int64_t> o(off, off + n_patches + 1);
    shard_range(o, shard_rank, shard_count, lo, hi);
    return GPC_OK;
}

int gpc_debug_rand(gpc_handle* h, uint64_t offset, int64_t n, uint32_t* out) {
    if (!h || n < 0 || (n > 0 && !out)) return GPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (n == 0) return GPC_OK;
    CK(h->tmpA.reserve(n * sizeof(uint32_t)));
    launch_rand_stream(offset, n, h->tmpA.as<uint32_t>(), h->stream);
    CK(cudaMemcpyAsync(out, h->tmpA.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return GPC_OK;
}

}  // extern "C"
