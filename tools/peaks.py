import sys; sys.path.insert(0, '.')
import gp_compressor_b200 as G
h = G.Handle()
for i in range(3):
    print("fp64 TFLOP/s %.2f   smem LDS.128 TB/s %.2f   smem LDS.64 TB/s %.2f" % (h.debug_peak(0) / 1e12, h.debug_peak(1) / 1e12, h.debug_peak(2) / 1e12))
