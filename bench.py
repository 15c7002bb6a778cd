#!/usr/bin/env python
"""bench.py — headline benchmark of the gp_compressor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2bind|c1|c5] [--points M]

Metric (BASELINE.json): compress pts/s (+ decompress pts/s), quoted on configs[1]:
synthetic 5M-point indoor-scan cloud, 0.1 m patches, SOGP capacity 30, 1 x B200.
A step = one compress (K1..K7) + one decompress (K8) of one batch (one cloud per GPU).

  value         compress pts/s with the cloud already resident in HBM (device timeline, CUDA
                events on the handle's stream, max over ranks)
  e2e           the same metric through the host-buffer C-ABI calls gpc_compress +
                gpc_get_params (+ gpc_decompress): H2D of the cloud from pinned memory and D2H
                of the results inside the timed region
  roofline      for the stage that dominates the step (see DESIGN.md for the byte / flop models)
  cpu_baseline  the CPU oracle (a port of the reference path; the reference itself cannot be
                built here) timed on this box's host cores on rank 0

--impl reference times the CPU oracle with all host threads on the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F32 = lambda v: float(np.float32(v))


RGB = False


def workload(name, points, seed_shift=0):
    cloud, cfg, desc = _workload(name, points, seed_shift)
    if RGB:
        cfg = dict(cfg, rgb=1)
        desc += ", + RGB field GP (N1)"
    return cloud, cfg, desc


def _workload(name, points, seed_shift=0):
    from gp_compressor_b200 import synth
    if name in ("c2", "c2bind"):
        n = points or 5_000_000
        cfg = dict(res=F32(0.1), sz=10, capacity=30)
        if name == "c2bind":
            cfg.update(synth.hyper_bind(F32(0.1)))
        desc = "C2 synthetic indoor scan, %d pts, res 0.1f, sz 10, capacity 30, hyper=%s" % (n, "BIND" if name == "c2bind" else "REF (reference defaults)")
        return synth.c2_indoor(n, seed=2 + seed_shift), cfg, desc
    if name == "c1":
        n = points or 100_000
        return synth.c1_planar_bumps(n, seed=1 + seed_shift), dict(res=F32(0.15), sz=20, capacity=100), "C1 planar+bumps, %d pts, res 0.15f, sz 20, capacity 100, hyper=REF" % n
    if name == "c5":
        n = points or 50_000_000
        return synth.c5_outdoor(n, seed=5 + seed_shift), dict(res=F32(0.2), sz=10, capacity=100), "C5 outdoor LiDAR-like, %d pts, res 0.2f, sz 10, capacity 100, hyper=REF" % n
    raise SystemExit("unknown workload " + name)


def config_dict(desc, n):
    """The `config` object of the JSON line: identical in both arms (--impl ours / reference) for the same workload."""
    mb = n * 32 / 1e6
    l2 = ("inputs larger than L2 (%.0f MB cloud + sort buffers per step)" % mb if mb > 126.0 else
          "inputs SMALLER than the 126 MB L2 (%.0f MB cloud) and no flush between steps: a secondary workload, L2-warm numbers" % mb)
    return {"workload": desc, "per_gpu_points": int(n), "l2": l2,
            "parallelism": "one cloud per GPU, patches independent, no collective"}


def strong_block(args, G, torch, dist, rank, world, local):
    """Strong scaling of ONE cloud (north star: C5, 50 M points, capacity 100) over the N GPUs of the run, through
    gpc_compress_shard_begin / _finish: every rank holds the whole cloud, bins its key range plus halo, ONE all-gather of two
    integers per rank places its patches and its window of the rand() stream, then it fits / decodes the patches it owns
    (gp_compressor.cpp:132-172, 204-243 are the loops being sharded).  Wall clock around begin + all-gather + finish
    bracketed by barriers, max over ranks; the cloud is resident (uploaded before the timed region)."""
    cloud, cfg, desc = _workload("c5", args.strong_points)      # same seed on every rank: the same cloud
    n = cloud.shape[0]
    h = G.Handle(device=local, shard_rank=rank, shard_count=world, **cfg)
    h.upload_cloud(cloud)
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    allc = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
    walls, gath, dev, dec, fit, decw = [], [], [], [], [], []
    reps, skip = 6, 3   # the first calls of a handle are untimed: allocation, then the handle's own two-vs-three-chains measurement
    for it in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        p, d = h.compress_shard_begin()
        sb = h.stats()
        g0 = time.perf_counter()
        counts[0] = p; counts[1] = d
        if world > 1:
            dist.all_gather_into_tensor(allc, counts)
            a = allc.cpu().numpy().reshape(world, 2)
        else:
            a = np.array([[p, d]], dtype=np.int64)
        tg = time.perf_counter() - g0
        h.compress_shard_finish(*G.binding.shard_prefix(a, rank))
        st = h.stats()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        nd = h.decompress_resident()
        sd = h.stats()
        torch.cuda.synchronize()
        w2 = time.perf_counter()
        if it >= skip:
            walls.append(1e3 * (w1 - w0)); gath.append(1e3 * tg); dev.append(st["ms_total"]); dec.append(sd["ms_predict"])
            fit.append(st["ms_fit"]); decw.append(1e3 * (w2 - w1))
    sz = h.sizes()
    own_pts = float(st["n_add"])   # points of the patches this rank owns (fed to its SOGP kernels)
    mx = torch.tensor([float(np.mean(walls)), float(np.mean(gath)), float(np.mean(dev)), float(np.mean(dec)), float(np.mean(fit)),
                       float(sz.n_claimed), own_pts, float(np.mean(decw))], dtype=torch.float64, device="cuda")
    sm = torch.tensor([float(np.mean(fit)), float(sz.n_claimed), own_pts, float(nd), float(sz.patch_hi - sz.patch_lo)], dtype=torch.float64, device="cuda")
    per_rank = None
    mine = torch.tensor([float(np.mean(fit)), float(sz.n_claimed), own_pts, float(sb["ms_total"]), float(np.mean(dev)), float(sz.patch_hi - sz.patch_lo),
                         float(st["ms_shuffle"]), float(np.mean(walls))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        allm = torch.zeros(world * mine.numel(), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allm, mine)
        per_rank = allm.cpu().numpy().reshape(world, -1)
    else:
        per_rank = mine.cpu().numpy().reshape(1, -1)
    h.close()
    wm, gm, dm, pm, fm, cl, op_, dw = mx.tolist()
    fs, cs, os_, nds, ps = sm.tolist()
    return {"workload": desc, "points": int(n), "n_gpus": world, "scaling": "strong",
            "mode": "binning sharded by key range + halo, one all-gather of 2 integers, patches fitted where they are owned",
            "compress_wall_ms": wm, "value": n / (wm * 1e-3), "unit": "pts/s", "allgather_wall_ms": gm,
            "max_rank_device_ms": dm, "max_rank_fit_ms": fm, "mean_rank_fit_ms": fs / world,
            "max_rank_binned_points": cl, "mean_rank_binned_points": cs / world,
            "max_rank_owned_points": op_, "mean_rank_owned_points": os_ / world,
            "decompress_wall_ms": dw, "max_rank_predict_ms": pm, "decompress_value": nds / (dw * 1e-3), "patches": int(ps),
            "per_rank": {k: [round(float(v), 3) for v in per_rank[:, i]] for i, k in enumerate(
                ["fit_ms", "binned_points", "owned_points", "begin_device_ms", "device_ms", "owned_patches", "shuffle_ms", "wall_ms"])},
            "timed_reps": reps - skip}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- algorithmic byte / flop models ------------------------------------
def stage_models(n_pts, n_valid, n_claimed, depth, n_patches, st, cap):
    sort_passes = -(-(3 * depth + 1) // 8) + -(-max(1, int(n_patches).bit_length()) // 8)
    return {
        "ms_sort": ("radix sort (hist + scatter, both sorts)", "hbm", sort_passes * n_valid * (8.0 + 12.0 + 12.0)),
        "ms_keys": ("point_keys", "hbm", n_pts * (32.0 + 12.0)),
        "ms_leaves": ("mark_heads + scan + fill_leaves", "hbm", n_valid * (8 + 8 + 8 + 8 + 8 + 12 + 4 + 32 + 16.0)),
        "ms_claim": ("claim", "hbm", n_valid * (16 + 4 + 12 + 24.0)),
        "ms_group": ("group_gather + patch_frames", "hbm", n_claimed * (12 + 4 + 16 + 24 + 4 + 8 + 8 + 8 + 4 + 8 + 8.0)),
        "ms_shuffle": ("rand + shuffle + gather", "hbm", n_claimed * (8 + 4 + 4 + 4 + 24 + 24.0)),
    }


def sogp_flops(st):
    E = 28.0  # flops per exp by convention
    return (4 * st["sum_n2_common"] + (13 + E) * st["sum_n"] + 20 * (st["n_add"] - st["n_first"])
            + 2 * st["sum_n2_sparse"] + 4 * st["n_sparse"] * 0 + 4 * st["sum_n2_full"] + 6 * st["sum_n2_del"])


def sogp_smem_bytes(st):
    # matvecs read C and Q once (16 N^2); sparse rank-1 reads+writes C (16 N^2); full update reads+writes
    # C and Q (32 (N+1)^2); a deletion reads+writes C and Q (32 (M-1)^2)
    return 16.0 * st["sum_n2_common"] + 16.0 * st["sum_n2_sparse"] + 32.0 * st["sum_n2_full"] + 32.0 * st["sum_n2_del"]


def run_reference(args, rank, world):
    """The reference arm: the reference's CPU implementation of the path (here: its port, the oracle,
    because the reference needs PCL + Eigen which are absent) on all host threads."""
    if rank != 0:
        return
    from oracle import oracle as O
    cloud, cfg, desc = workload(args.workload, args.points)
    cores = os.cpu_count() or 1
    o = O.Oracle(threads=cores, **cfg)
    n = cloud.shape[0]
    tc, td, nd = [], [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.compress(cloud)
        t1 = time.perf_counter()
        cl, _ = o.decode(want_heights=False)
        t2 = time.perf_counter()
        nd = cl.shape[0]
        if it >= args.warmup:
            tc.append(t1 - t0); td.append(t2 - t1)
    sec = float(np.sum(tc))
    v = n * len(tc) / sec
    line = {"impl": "reference", "metric": "compress pts/s", "value": v, "unit": "pts/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (sec + float(np.sum(td))) / len(tc), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(desc, n),
            "decompress": {"value": nd * len(td) / float(np.sum(td)), "unit": "grid pts/s"},
            "cpu_baseline": {"value": v, "unit": "pts/s", "cores": cores, "kind": "port", "sample": "%d pts (whole workload) per step" % n},
            "e2e": {"value": v, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rgb", action="store_true", help="also fit / decode the RGB field GP (next-row N1), in both arms")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (C5, one cloud sharded over the N GPUs)")
    ap.add_argument("--strong-points", type=int, default=0, help="points of the strong-scaling cloud (default: C5's 50 M)")
    args = ap.parse_args()
    global RGB
    RGB = args.rgb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gp_compressor_b200 as G
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The collectives here are a barrier and a few integers: NVLink SHARP (NVLS) brings nothing, and a communicator set up
        # WITH it makes three concurrent kernel chains of the fit stage slower on the same GPU (measured at 2 ranks: fit 2.11 ms
        # with NVLS, 1.85 ms without).  A process that keeps NVLS is covered by the handle's own choice
        # between two and three chains.
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cloud_np, cfg, desc = workload(args.workload, args.points, seed_shift=rank)  # weak scaling: one cloud per GPU
    n = cloud_np.shape[0]
    # pinned host buffers for the end-to-end arm
    pin_in = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = cloud_np
    h = G.Handle(device=local, **cfg)
    ext = torch.cuda.ExternalStream(h.stream(), device=torch.device("cuda", local))

    # ---- device-resident arm -----------------------------------------------------------
    h.upload_cloud_ptr(pin_in.data_ptr(), n)
    evs, stage_acc = [], {}
    n_dec = 0
    launches = 0
    sampler = ClockSampler(local) if rank == 0 else None   # runs over warm-up, the timed loops and the e2e loops
    e_all0 = torch.cuda.Event(enable_timing=True)
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
            e_all0.record(ext)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        h.compress_resident()
        st_c = h.stats()
        e1.record(ext)
        n_dec = h.decompress_resident()
        st_d = h.stats()
        e2.record(ext)
        if it >= args.warmup:
            evs.append((e0, e1, e2))
            launches += st_c["kernel_launches"] + st_d["kernel_launches"]
            for k, v in st_c.items():
                if k.startswith("ms_"):
                    stage_acc[k] = stage_acc.get(k, 0.0) + v
            stage_acc["ms_predict"] = stage_acc.get("ms_predict", 0.0) + st_d["ms_predict"]
    e_all1 = torch.cuda.Event(enable_timing=True); e_all1.record(ext)
    barrier()
    total_ms = e_all0.elapsed_time(e_all1)
    comp_ms = [a.elapsed_time(b) for a, b, _ in evs]
    dec_ms = [b.elapsed_time(c) for _, b, c in evs]
    sizes = h.sizes()
    fit_stats = st_c
    K = args.steps
    t = torch.tensor([float(np.sum(comp_ms)), float(np.sum(dec_ms)), total_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n), float(n_dec)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    comp_s, dec_s, tot_s = (t / 1e3).tolist()
    n_all, ndec_all = cnt.tolist()
    value = n_all * K / comp_s
    dec_value = ndec_all * K / dec_s

    # ---- end-to-end arm: host buffers through the C ABI --------------------------------------
    out_cap = int(sizes.patch_hi - sizes.patch_lo) * cfg["sz"] * cfg["sz"]
    pin_out = torch.empty((max(out_cap, 1), 32), dtype=torch.uint8, pin_memory=True)
    e2e_c, e2e_d = [], []
    d2h = 0
    # pinned result buffers for the compressed parameters (worst case: capacity entries per patch)
    PLn = int(sizes.patch_hi - sizes.patch_lo)
    capn = PLn * cfg["capacity"]
    pin_nbv = torch.empty(max(PLn, 1), dtype=torch.int32, pin_memory=True)
    pin_off = torch.empty(PLn + 1, dtype=torch.int64, pin_memory=True)
    pin_idx = torch.empty(max(capn, 1), dtype=torch.int32, pin_memory=True)
    pin_b1 = torch.empty(max(capn, 1), dtype=torch.float64, pin_memory=True)
    pin_b2 = torch.empty(max(capn, 1), dtype=torch.float64, pin_memory=True)
    pin_al = torch.empty(max(capn, 1), dtype=torch.float64, pin_memory=True)
    pin_fl = torch.empty(max(PLn, 1), dtype=torch.int32, pin_memory=True)
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
        t0 = time.perf_counter()
        h.compress_ptr(pin_in.data_ptr(), n)
        # compressed parameters back on the host (alpha, BV, indices, counts)
        h.params_into(pin_nbv.data_ptr(), pin_off.data_ptr(), pin_idx.data_ptr(), pin_b1.data_ptr(), pin_b2.data_ptr(), pin_al.data_ptr(), pin_fl.data_ptr())
        t1 = time.perf_counter()
        nd = h.decompress_ptr(pin_out.data_ptr(), pin_out.shape[0])
        t2 = time.perf_counter()
        if it >= args.warmup:
            e2e_c.append(t1 - t0); e2e_d.append(t2 - t1)
            nbt = int(pin_off[PLn])
            d2h = PLn * 4 * 2 + (PLn + 1) * 8 + nbt * (4 + 8 * 3) + nd * 32
    barrier()
    # ---- pipelined end-to-end arm: a stream of clouds, two handles in flight ------------------------
    # Same calls and bytes per step as the end-to-end arm; a second handle (own stream, own buffers) lets the H2D
    # copy of one cloud overlap the kernels of the other.  Reported beside e2e, never instead of it.
    import threading
    h2 = G.Handle(device=local, **cfg)
    pin_in2 = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    pin_in2.numpy()[:] = cloud_np
    bufs2 = [torch.empty_like(b).pin_memory() for b in (pin_nbv, pin_off, pin_idx, pin_b1, pin_b2, pin_al, pin_fl)]
    lanes = [(h, pin_in, (pin_nbv, pin_off, pin_idx, pin_b1, pin_b2, pin_al, pin_fl)), (h2, pin_in2, tuple(bufs2))]
    per_lane = max(K // 2, 1)

    lane_errors = []

    def lane_loop(hh, pin, outs, reps):
        try:
            torch.cuda.set_device(local)
            for _ in range(reps):
                hh.compress_ptr(pin.data_ptr(), n)
                hh.params_into(*(o.data_ptr() for o in outs))
        except Exception as e:  # a failed lane must not pass as a fast one
            lane_errors.append(repr(e))

    for hh, pin, outs in lanes:
        lane_loop(hh, pin, outs, 3)  # warm-up (allocations of the second handle and its chain measurement)
    barrier()
    tp0 = time.perf_counter()
    ths = [threading.Thread(target=lane_loop, args=(hh, pin, outs, per_lane)) for hh, pin, outs in lanes]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    tp1 = time.perf_counter()
    tpipe = torch.tensor([tp1 - tp0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tpipe, op=dist.ReduceOp.MAX)
    pipe_s = float(tpipe.item())
    h2.close()
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + device-resident timed loop + end-to-end loop (nvidia-smi -lms 20)"
    te = torch.tensor([float(np.sum(e2e_c)), float(np.sum(e2e_d))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_cs, e2e_ds = te.tolist()

    # ---- strong scaling of one cloud over the N GPUs (reported beside the weak headline, never instead of it) ----
    strong = None
    if not args.no_strong:
        try:
            h_keep = h  # the headline handle stays alive for the RMSE / roofline legs below
            strong = strong_block(args, G, torch, dist, rank, world, local)
        except Exception as e:  # reported, never fatal for the headline
            strong = {"error": repr(e)}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- reconstruction RMSE: sqrt(mean over the claimed points of (y_local - f*(x_local))^2), metres, in the
    # patch frames; the means of the fitted GPs at their own training points, outside every timed region, rank 0 only ----
    rmse = None
    try:
        h.set_rand_offset(0)        # the reference's first call: the unseeded rand() stream from its start
        h.compress_resident()
        asg = h.assignment(binning=False)
        poff = h.patches(frames=False, binning=False)["patch_off"]
        fhat = h.evaluate(poff, asg["st_x1"], asg["st_x2"], want=("f",))["f"]
        rmse = float(np.sqrt(np.mean((fhat - asg["st_y"]) ** 2))) if fhat.size else 0.0
    except Exception as e:  # reported, never fatal for the timing line
        rmse = "unavailable: %s" % e

    # ---- roofline of the dominant stage --------------------------------------------------------
    peak, peak_src = measured_peaks()
    stage_ms = {k: v / K for k, v in stage_acc.items() if k not in ("ms_total", "ms_h2d", "ms_d2h", "ms_evaluate", "pad_") and (v > 0 or k != "ms_fit_rgb")}
    dom = max(stage_ms, key=stage_ms.get)
    n_valid = n  # all points finite in the synthetic clouds
    models = stage_models(n, n_valid, sizes.n_claimed, sizes.depth, sizes.n_patches, fit_stats, cfg["capacity"])
    roof = None
    if dom in models:
        name, bound, nbytes = models[dom]
        ach = nbytes / (stage_ms[dom] * 1e-3) / 1e9
        roof = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "stage_ms": stage_ms[dom]}
    elif dom == "ms_fit":
        fl = sogp_flops(fit_stats)
        sm = sogp_smem_bytes(fit_stats)
        sec = stage_ms[dom] * 1e-3
        # FP64 / shared-memory peaks measured on this device by the library's micro-benchmarks (csrc/k_peaks.cu)
        fp64_peak, smem_peak = h.debug_peak(0), h.debug_peak(1)
        f1, f2 = fl / sec / fp64_peak, sm / sec / smem_peak
        roof = {"kernel": "sogp_fit_kernel (K7)", "bound": "smem" if f2 >= f1 else "fp64", "achieved": (sm if f2 >= f1 else fl) / sec / 1e9,
                "peak": (smem_peak if f2 >= f1 else fp64_peak) / 1e9, "unit": "GB/s" if f2 >= f1 else "GFLOP/s", "frac": max(f1, f2),
                "traffic": None, "peak_source": "measured here: independent DFMA chains %.1f TFLOP/s, conflict-free LDS.128 %.1f TB/s (gpc_debug_peak)" % (fp64_peak / 1e12, smem_peak / 1e12), "stage_ms": stage_ms[dom],
                "fp64_frac": f1, "smem_frac": f2, "mean_n": fit_stats["sum_n"] / max(1, fit_stats["n_add"] - fit_stats["n_first"])}
        if args.workload == "c2" and not args.points and not RGB:
            # NOT live: one `ncu --set full` capture of this same command, kept under profiles/ with its grid and duration so
            # that it can be told apart from the live numbers above (traffic = dram bytes read + written by the bucket-0 launches)
            cap_path = os.path.join(ROOT, "profiles", "r2_k7_capture.json")
            if os.path.exists(cap_path):
                roof["captured_once"] = json.load(open(cap_path))
                roof["traffic"] = roof["captured_once"].get("dram_bytes_per_step")
            roof["note"] = ("achieved / frac are ALGORITHMIC bytes and flops (event counters of the run) over the live stage time; "
                            "traffic and captured_once come from one ncu capture of the same command, not from this run")
    elif dom == "ms_predict":
        npatch = max(1, n_dec / (cfg["sz"] ** 2))
        fl = n_dec * (3.0 * (sizes.n_bv_total / npatch) + 36) + 2.0 * cfg["sz"] * sizes.n_bv_total * 37.0   # separable tables
        sec = stage_ms[dom] * 1e-3
        fp64_peak = h.debug_peak(0)
        roof = {"kernel": "predict_grid_kernel (K8)", "bound": "fp64", "achieved": fl / sec / 1e9, "peak": fp64_peak / 1e9, "unit": "GFLOP/s",
                "frac": fl / sec / fp64_peak, "traffic": None, "peak_source": "measured here (gpc_debug_peak)", "stage_ms": stage_ms[dom]}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}

    # ---- CPU baseline (oracle, 1 thread = faithful to the single-threaded reference) -------------
    cpu = None
    if not args.no_cpu_baseline and world >= 1:
        from oracle import oracle as O
        sample = min(n, 2_000_000)
        sub = np.ascontiguousarray(cloud_np[:sample])
        o = O.Oracle(threads=1, **cfg)
        t0 = time.perf_counter()
        o.compress(sub)
        t1 = time.perf_counter()
        cpu = {"value": sample / (t1 - t0), "unit": "pts/s", "cores": 1, "kind": "port",
               "sample": "first %d points of the same cloud, 1 thread (the reference is single-threaded)" % sample}

    line = {
        "metric": "compress pts/s", "value": value, "unit": "pts/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(desc, n),
        "decompress": {"value": dec_value, "unit": "grid pts/s", "points_per_step": ndec_all},
        "rmse_m": rmse,
        "compress_ms": 1e3 * comp_s / K, "decompress_ms": 1e3 * dec_s / K,
        "stages_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items())},
        "patches": int(sizes.n_patches), "claimed": int(sizes.n_claimed), "mean_bv": sizes.n_bv_total / max(1, sizes.n_patches),
        "escalated": fit_stats["escalated"], "max_bv": fit_stats["max_bv"],
        "fit_events": {k: int(fit_stats[k]) for k in ("n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo", "sum_n", "sum_n2_common",
                                                      "sum_n2_sparse", "sum_n2_full", "sum_n2_del")},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": n_all * K / e2e_cs, "unit": "pts/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": int(d2h),
                "decompress_value": ndec_all * K / e2e_ds, "compress_ms": 1e3 * e2e_cs / K, "decompress_ms": 1e3 * e2e_ds / K},
        "e2e_pipelined": {"value": None if lane_errors else n_all * 2 * per_lane / pipe_s, "errors": lane_errors, "unit": "pts/s", "in_flight": 2, "steps": 2 * per_lane,
                          "ms_per_step": 1e3 * pipe_s / (2 * per_lane),
                          "note": "same calls and bytes per step as e2e, two handles / host threads per GPU so that the H2D copy of one cloud overlaps the kernels of the other; wall clock"},
        "gpu_launches": int(launches), "clocks": clocks, "strong": strong,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
