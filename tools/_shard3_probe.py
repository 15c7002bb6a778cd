import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import gp_compressor_b200 as G
from bench import _workload
cloud, cfg, desc = _workload("c5", 0)
h = G.Handle(shard_rank=3, shard_count=8, **cfg)
h.upload_cloud(cloud)
for rep in range(2):
    p, d = h.compress_shard_begin()
    h.compress_shard_finish(0, 0, p, d)
    print(h.stats()["ms_fit"], h.stats()["escalated"])
