import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import gp_compressor_b200 as G
from gp_compressor_b200 import synth
F32=lambda v: float(np.float32(v))
cloud = synth.c2_indoor(5_000_000, seed=2)
pin = torch.empty((cloud.shape[0],32), dtype=torch.uint8, pin_memory=True); pin.numpy()[:] = cloud
h = G.Handle(res=F32(0.1), sz=10, capacity=30)
for it in range(4):
    t0=time.perf_counter(); h.compress_ptr(pin.data_ptr(), cloud.shape[0]); t1=time.perf_counter()
    st=h.stats(); prm=h.params(); t2=time.perf_counter()
    print('compress call %.2f ms (h2d %.2f, device total %.2f, lattice %.2f fit %.2f) get_params %.2f ms' % (1e3*(t1-t0), st['ms_h2d'], st['ms_total'], st['ms_lattice'], st['ms_fit'], 1e3*(t2-t1)))
# pageable
t0=time.perf_counter(); h.compress(cloud); t1=time.perf_counter(); print('pageable compress %.2f ms'%(1e3*(t1-t0)), h.stats()['ms_h2d'])
