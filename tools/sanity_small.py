"""Small run that touches every kernel (all SOGP buckets, RGB field GP, decode, wire format) — used under compute-sanitizer."""
import sys; sys.path.insert(0, '.')
import numpy as np
import gp_compressor_b200 as G
from gp_compressor_b200 import synth
F32 = lambda v: float(np.float32(v))
cloud = synth.c2_indoor(8000, seed=1)
h = G.Handle(res=F32(0.1), sz=4, capacity=30, rgb=1)
h.compress(cloud); h.decompress(); h.params(); h.params_rgb(); h.assignment(); h.patches()
dense = synth.c3_dense_floor(6000, seed=2, side=0.3)   # ~660 points per patch
for cap in (8, 20, 40, 100, 130):
    hh = G.Handle(res=F32(0.1), sz=3, capacity=cap, rgb=1, rgb_s0=1e-2, **synth.hyper_bind(F32(0.1)))
    hh.compress(dense); hh.decompress()
    print(cap, hh.stats()["escalated"], hh.sizes().n_bv_total)
rng = np.random.default_rng(0)
off = np.array([0, 50, 50, 400]); n = 400
hf = G.Handle(capacity=64, keep_state=1, **synth.hyper_bind(0.1))
hf.fit_patches(off, rng.uniform(-.05, .05, n), rng.uniform(-.05, .05, n), rng.normal(0, .01, n))
hf.predict(2, rng.uniform(-.05, .05, (10, 2)), sigma=True); hf.decompress_resident(); hf.heights()
print("sanity ok")
