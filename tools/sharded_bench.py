"""Strong scaling of ONE cloud over the GPUs of a box (north star: C5, 50 M points, capacity 100).
Every rank receives the whole cloud.  Two modes:
  replicated  (argv[3] == "replicated"): every rank bins the whole cloud (K1-K5) and fits / decodes its patch range;
  sharded     (default): gpc_compress_shard_begin bins only the rank's key range plus a three-voxel halo, one NCCL
              all-gather of two integers per rank, gpc_compress_shard_finish fits the owned patches.
Launch with torchrun; prints one JSON line on rank 0.  Time = max over ranks of the device times (CUDA events inside the
library: gpc_stats.ms_total of begin + finish) plus, for the sharded mode, the wall time of the all-gather."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
import gp_compressor_b200 as G
from bench import workload


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c5"
    points = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    mode = sys.argv[3] if len(sys.argv) > 3 else "sharded"
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cloud, cfg, desc = workload(name, points)        # same seed on every rank: the same cloud
    n = cloud.shape[0]
    h = G.Handle(device=local, shard_rank=rank, shard_count=world, **cfg)
    h.upload_cloud(cloud)
    times, dec, walls, gathers, st = [], [], [], [], None
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    allc = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
    for it in range(4):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        if mode == "replicated":
            h.compress_resident()
            st = h.stats()
            tg = 0.0
        else:
            p, d = h.compress_shard_begin()
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
        if it >= 1:
            times.append(st["ms_total"]); dec.append(sd["ms_predict"]); walls.append(1e3 * (w1 - w0)); gathers.append(1e3 * tg)
    sz = h.sizes()
    t = torch.tensor([float(np.mean(times)), float(np.mean(dec)), float(st["ms_fit"]), float(st["ms_shuffle"]), float(np.mean(walls)),
                      float(np.mean(gathers)), float(sz.n_claimed)], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(nd), float(sz.patch_hi - sz.patch_lo)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        cm, dm, fm, sm, wm, gm, cl = t.tolist()
        print(json.dumps({"workload": desc, "n_gpus": world, "mode": mode,
                          "scaling": "strong (one cloud; " + ("binning replicated" if mode == "replicated" else "binning sharded by key range + 3-voxel halo, one all-gather of 2 integers") + ")",
                          "compress_device_ms": cm, "compress_wall_ms": wm, "allgather_wall_ms": gm,
                          "compress_pts_per_s": n / (wm * 1e-3), "decompress_ms": dm, "decompress_grid_pts_per_s": cnt[0].item() / (dm * 1e-3),
                          "max_rank_fit_ms": fm, "max_rank_shuffle_ms": sm, "max_rank_binned_points": cl, "patches": cnt[1].item()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
