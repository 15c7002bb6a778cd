"""Per-rank view of the strong-scaling run WITHOUT 8 GPUs: the 8 shards of one cloud are run one after the other on ONE GPU
(begin for every shard, the all-gather done on the host, then finish for every shard).  Gives each rank's begin / fit / finish
device times as an 8-GPU run would see them (no contention).  Development tooling for DESIGN.md section 5."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_compressor_b200 as G
from bench import _workload

def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    points = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cloud, cfg, desc = _workload("c5", points)
    rows = []
    hs, counts = [], []
    for r in range(world):
        h = G.Handle(shard_rank=r, shard_count=world, **cfg)
        h.upload_cloud(cloud)
        hs.append(h)
    for rep in range(2):
        counts, begins = [], []
        for h in hs:
            p, d = h.compress_shard_begin()
            counts.append((p, d)); begins.append(h.stats())
        a = np.array(counts, dtype=np.int64)
        rows = []
        for r, h in enumerate(hs):
            h.compress_shard_finish(*G.binding.shard_prefix(a, r))
            st = h.stats(); sz = h.sizes()
            rows.append(dict(rank=r, begin_ms=round(begins[r]["ms_total"], 2), finish_ms=round(st["ms_total"], 2), fit_ms=round(st["ms_fit"], 2),
                             shuffle_ms=round(st["ms_shuffle"], 2), binned=int(sz.n_claimed), owned=int(st["n_add"]), patches=int(sz.patch_hi - sz.patch_lo),
                             max_bv=st["max_bv"], escalated=st["escalated"][:2]))
    for row in rows:
        print(json.dumps(row))
    print(json.dumps({"max_begin": max(r["begin_ms"] for r in rows), "max_finish": max(r["finish_ms"] for r in rows),
                      "sum": max(r["begin_ms"] for r in rows) + max(r["finish_ms"] for r in rows)}))

if __name__ == "__main__":
    main()
