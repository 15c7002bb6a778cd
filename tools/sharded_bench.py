"""Strong scaling of ONE cloud sharded by patch range over the GPUs of a box (north star: C5, 50 M points, capacity 100).
Every rank receives the whole cloud, bins it redundantly (K1-K5) and fits / decodes only its patch range (K6-K8).
Launch with torchrun; prints one JSON line on rank 0.  Time = max over ranks of the device-resident compress."""
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
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cloud, cfg, desc = workload(name, points)        # same seed on every rank: the same cloud
    n = cloud.shape[0]
    h = G.Handle(device=local, shard_rank=rank, shard_count=world, **cfg)
    h.upload_cloud(cloud)
    times, dec, st = [], [], None
    for it in range(4):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        h.compress_resident()
        st = h.stats()
        nd = h.decompress_resident()
        sd = h.stats()
        if it >= 1:
            times.append(st["ms_total"]); dec.append(sd["ms_predict"])
    t = torch.tensor([float(np.mean(times)), float(np.mean(dec)), float(st["ms_fit"]), float(st["ms_shuffle"])], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(nd)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        cm, dm, fm, sm = t.tolist()
        print(json.dumps({"workload": desc, "n_gpus": world, "scaling": "strong (one cloud, patches sharded, binning redundant)", "compress_ms": cm,
                          "compress_pts_per_s": n / (cm * 1e-3), "decompress_ms": dm, "decompress_grid_pts_per_s": cnt.item() / (dm * 1e-3),
                          "max_rank_fit_ms": fm, "max_rank_shuffle_ms": sm, "redundant_binning_ms": cm - fm - sm}))
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
