"""BASELINE configs[2] and [3] at single-GPU scale: SOGP capacity sweep (shared-memory-resident vs spill regime) on a
dense cloud, and decode-only grid evaluation.  Prints markdown tables (stored in profiles/)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import gp_compressor_b200 as G
from gp_compressor_b200 import synth
F32 = lambda v: float(np.float32(v))

def fp64_flops(st):
    return (4 * st["sum_n2_common"] + 41 * st["sum_n"] + 20 * (st["n_add"] - st["n_first"]) + 2 * st["sum_n2_sparse"]
            + 4 * st["sum_n2_full"] + 6 * st["sum_n2_del"])
def smem_bytes(st):
    return 16.0 * st["sum_n2_common"] + 16.0 * st["sum_n2_sparse"] + 32.0 * st["sum_n2_full"] + 32.0 * st["sum_n2_del"]

def c3(points, side, caps=(10, 20, 30, 50, 75, 100, 117, 150, 200)):
    cloud = synth.c3_dense_floor(points, seed=3, side=side)
    print(f"\n### C3 capacity sweep: {points} pts on {side}x{side} m (~{points / (side / 0.1) ** 2:.0f} pts/patch), res 0.1f, hyper BIND, 1 x B200\n")
    print("| capacity | bucket regime | fit ms | compress pts/s | mean BV | full / sparse / cap-del | GFLOP/s (alg.) | SMEM GB/s (alg.) | frac of LDS.128 peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    h0 = G.Handle()
    smem_peak = h0.debug_peak(1)
    for cap in caps:
        h = G.Handle(res=F32(0.1), sz=10, capacity=cap, **synth.hyper_bind(F32(0.1)))
        h.upload_cloud(cloud)
        best = None
        for it in range(2):
            h.compress_resident()
            st = h.stats()
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        s = h.sizes()
        regime = "warp (<=15)" if cap <= 15 else "pair (<=31)" if cap <= 31 else "CTA smem (<=63)" if cap <= 63 else "CTA smem (<=117)" if cap <= 117 else "spill (global/L2)"
        sec = best["ms_fit"] * 1e-3
        print(f"| {cap} | {regime} | {best['ms_fit']:.2f} | {points / (best['ms_total'] * 1e-3):.3e} | {s.n_bv_total / max(1, s.n_patches):.1f} | "
              f"{best['n_full']} / {best['n_sparse']} / {best['n_del_cap']} | {fp64_flops(best) / sec / 1e9:.0f} | {smem_bytes(best) / sec / 1e9:.0f} | {smem_bytes(best) / sec / smem_peak:.3f} |")
        h.close()

def c4(n_patches, nbv, sz=64):
    prm = synth.c4_patch_params(n_patches=n_patches, nbv=nbv, seed=4)
    h = G.Handle(res=F32(0.1), sz=sz, capacity=nbv)
    h.set_params(**prm)
    fp64_peak = h.debug_peak(0)
    best = 1e9
    for it in range(3):
        n = h.decompress_resident()
        best = min(best, h.stats()["ms_predict"])
    # separable decode: per grid point 1 DMUL + 1 DFMA (3 flop) per BV + 36 for the frame; per patch 2 sz N exps (34 flop each)
    flops = n * (3.0 * nbv + 36) + n_patches * 2.0 * sz * nbv * 37.0
    hbm = float(G.measured_peaks().get("hbm_gbs", 6451.5)) if hasattr(G, "measured_peaks") else 6451.5
    gbs = n * 32 / (best * 1e-3) / 1e9
    print(f"| {n_patches} | {nbv} | {sz}x{sz} | {n} | {best:.2f} | {n / (best * 1e-3):.3e} | {flops / (best * 1e-3) / 1e12:.2f} | {flops / (best * 1e-3) / fp64_peak:.3f} | {gbs:.0f} | {gbs / hbm:.3f} |")
    h.close()

def ev(n_patches, cap, n_fit=300, n_query=256, res=0.1):
    """K9 (rows N2 / N4): batched predict + sigma + likelihood + gradient at n_query points per patch."""
    rng = np.random.default_rng(cap)
    off = np.arange(n_patches + 1, dtype=np.int64) * n_fit
    x1 = rng.uniform(-res / 2, res / 2, n_patches * n_fit); x2 = rng.uniform(-res / 2, res / 2, n_patches * n_fit)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, x1.size)
    h = G.Handle(capacity=cap, keep_state=1, sigmaf_sq=1.0, l_sq=(res / 12.0) ** 2, s0=1e-4)
    h.fit_patches(off, x1, x2, y)
    nbv = h.params()["nbv"].astype(np.float64)
    qoff = np.arange(n_patches + 1, dtype=np.int64) * n_query
    q1 = rng.uniform(-res / 2, res / 2, n_patches * n_query); q2 = rng.uniform(-res / 2, res / 2, n_patches * n_query)
    qy = 0.02 * np.sin(40 * q1) * np.cos(30 * q2)
    fp64_peak = h.debug_peak(0)
    best = 1e9
    for it in range(3):
        h.evaluate(qoff, q1, q2, qy)
        best = min(best, h.stats()["ms_evaluate"])
    flops = float((n_query * (2 * nbv * nbv + 49 * nbv + 60)).sum())
    m = n_patches * n_query
    print(f"| {n_patches} | {cap} | {nbv.mean():.1f} | {m} | {best:.2f} | {m / (best * 1e-3):.3e} | {flops / (best * 1e-3) / 1e12:.2f} | {flops / (best * 1e-3) / fp64_peak:.3f} |")
    h.close()

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eval":
        print("| patches | capacity | mean BV | query pts | evaluate ms | pts/s | TFLOP/s (2N^2+49N+60 per pt) | frac of FP64 peak |")
        print("|---|---|---|---|---|---|---|---|")
        ev(20000, 12); ev(20000, 30); ev(10000, 60); ev(5000, 100); ev(2000, 150)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c4":
        c4(125_000, 30); c4(125_000, 100); c4(250_000, 30); c4(125_000, 30, sz=10)
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "c3":
        c3(1_000_000, 3.2, caps=tuple(int(v) for v in sys.argv[2:]))
        sys.exit(0)
    c3(1_000_000, 3.2)
    print("\n### C4 decode only (per-GPU share of the 1M-patch configuration), REF kernel, 1 x B200\n")
    print("| patches | BVs/patch | grid | output pts | predict ms | grid pts/s | TFLOP/s (alg.) | frac of FP64 peak | output GB/s | frac of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    c4(125_000, 30)
    c4(125_000, 100)
    c4(250_000, 30)
