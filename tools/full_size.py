"""BASELINE configs[2] and [3] AT THE STATED SIZES.

  c3 : SOGP capacity sweep 10..200 BVs on a 10 M-point cloud (c3_dense_floor, ~1000 points per patch, hyper-set BIND so that
       the capacity binds), 1 x B200: dense shared-memory buckets (<= 117), spill bucket (> 117).
  c4 : decompression only: 1 M patches x 64 x 64 grid (4.096e9 grid points = 131 GB of PointXYZRGB records), N = 30 and 100
       BVs per patch, at the GPUs of the run (python -m torch.distributed.run --nproc-per-node G tools/full_size.py c4): every
       rank decodes its 1/G of the patches in chunks of 125 k patches (16 GB of output per chunk, reused), both decode modes.
One JSON line per row on rank 0; `profiles/r2_full_size.md` is assembled from them."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_compressor_b200 as G  # noqa: E402
from gp_compressor_b200 import synth  # noqa: E402

F32 = lambda v: float(np.float32(v))


def smem_bytes(st):
    return 16.0 * st["sum_n2_common"] + 16.0 * st["sum_n2_sparse"] + 32.0 * st["sum_n2_full"] + 32.0 * st["sum_n2_del"]


def fp64_flops(st):
    return (4 * st["sum_n2_common"] + 41 * st["sum_n"] + 20 * (st["n_add"] - st["n_first"]) + 2 * st["sum_n2_sparse"]
            + 4 * st["sum_n2_full"] + 6 * st["sum_n2_del"])


def c3(points=10_000_000, caps=(10, 20, 30, 50, 75, 100, 117, 118, 150, 200)):
    cloud = synth.c3_dense_floor(points, seed=3, side=10.0)
    h0 = G.Handle()
    fp64_peak, smem_peak = h0.debug_peak(0), h0.debug_peak(1)
    h0.close()
    for cap in caps:
        h = G.Handle(res=F32(0.1), sz=10, capacity=cap, **synth.hyper_bind(F32(0.1)))
        h.upload_cloud(cloud)
        best = None
        for it in range(2 if cap <= 118 else 1):
            h.compress_resident()
            st = h.stats()
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        s = h.sizes()
        sec = best["ms_fit"] * 1e-3
        regime = ("half-warp, registers" if cap <= 15 else "two warps, smem" if cap <= 31 else "CTA fused, smem" if cap <= 103
                  else "CTA step-by-step, smem" if cap <= 117 else "CTA fused, global/L2 spill")
        print(json.dumps({"config": "C3", "points": points, "capacity": cap, "regime": regime, "patches": int(s.n_patches),
                          "compress_ms": best["ms_total"], "fit_ms": best["ms_fit"], "compress_pts_per_s": points / (best["ms_total"] * 1e-3),
                          "mean_bv": s.n_bv_total / max(1, s.n_patches), "n_full": best["n_full"], "n_sparse": best["n_sparse"],
                          "n_del_cap": best["n_del_cap"], "alg_gflops": fp64_flops(best) / sec / 1e9, "alg_smem_gbs": smem_bytes(best) / sec / 1e9,
                          "frac_fp64": fp64_flops(best) / sec / fp64_peak, "frac_lds128": smem_bytes(best) / sec / smem_peak}), flush=True)
        h.close()


def c4(total_patches=1_000_000, chunk=125_000, sz=64):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = total_patches // world
    for nbv in (30, 100):
        for separable in (0, 1):
            h = G.Handle(device=local, res=F32(0.1), sz=sz, capacity=nbv, decode_separable=separable)
            n_chunks = (mine + chunk - 1) // chunk
            ms, pts = 0.0, 0
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for c in range(n_chunks):
                m = min(chunk, mine - c * chunk)
                prm = synth.c4_patch_params(n_patches=m, nbv=nbv, seed=4 + 1000 * rank + c)
                h.set_params(**prm)          # host -> device parameters of the chunk (outside the kernel time, inside the wall time)
                pts += h.decompress_resident()
                ms += h.stats()["ms_predict"]
            torch.cuda.synchronize()
            wall = time.perf_counter() - w0
            t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            n = torch.tensor([float(pts)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(n, op=dist.ReduceOp.SUM)
            if rank == 0:
                kms, wl = t.tolist()
                tot = n.item()
                print(json.dumps({"config": "C4", "patches": total_patches, "n_gpus": world, "bv_per_patch": nbv, "grid": "%dx%d" % (sz, sz),
                                  "mode": "separable (flagged)" if separable else "direct (reference arithmetic, default)",
                                  "grid_points": tot, "output_GB": tot * 32 / 1e9, "max_rank_kernel_ms": kms,
                                  "grid_pts_per_s": tot / (kms * 1e-3), "output_GBps_per_gpu": tot * 32 / world / (kms * 1e-3) / 1e9,
                                  "frac_hbm_write_peak": tot * 32 / world / (kms * 1e-3) / 1e9 / 6451.5,
                                  "wall_s_incl_param_generation": wl}), flush=True)
            h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    if which == "c3":
        c3(int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000)
    else:
        c4(int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000)
