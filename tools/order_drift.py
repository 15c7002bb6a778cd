"""How far does the product's canonical arithmetic (kernel-shaped reduction order, canonical exp) drift from the
reference's own evaluation order on a real workload?  Both run on the CPU oracle: `ref_order=1` is bit-equal to the
reference source compiled over oracle/eigen_shim (tests/test_oracle_reference_order.py), `ref_order=0` is bit-equal to the
CUDA path (tests/ -m gpu).  Same cloud, same binning, same shuffles; reports per-patch BV-set agreement and the distance of
the decoded heights.  Test / documentation tooling: python tools/order_drift.py [--points N] [--workload c2|c1]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gp_compressor_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=5_000_000)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    F = np.float32
    if a.workload == "c2":
        cloud, cfg = synth.c2_indoor(a.points, seed=2), dict(res=float(F(0.1)), sz=10, capacity=30)
    else:
        cloud, cfg = synth.c1_planar_bumps(a.points, seed=1), dict(res=float(F(0.15)), sz=20, capacity=100)
    out = {}
    t0 = time.time()
    for mode in (0, 1):
        o = O.Oracle(threads=a.threads, ref_order=mode, **cfg)   # reference default hyper-parameters (rbf 100 / 1, s0 1e-1f)
        r = o.compress(cloud)
        _, h = o.decode(want_cloud=False)
        out[mode] = (r, h)
        if mode == 0:
            b = o.binning()
            poff, sx1, sx2, sy = b["patch_off"], b["st_x1"], b["st_x2"], b["st_y"]
    (rc, hc), (rr, hr) = out[0], out[1]
    P = rc["nbv"].size
    same_n = rc["nbv"] == rr["nbv"]
    same_set = np.zeros(P, dtype=bool)
    for p in np.nonzero(same_n)[0]:
        a0, a1 = rc["bv_off"][p], rc["bv_off"][p + 1]
        same_set[p] = np.array_equal(np.sort(rc["bv_idx"][a0:a1]), np.sort(rr["bv_idx"][a0:a1]))
    nonempty = rc["nbv"] > 0
    g2 = cfg["sz"] ** 2
    assert hc.size == hr.size == int(nonempty.sum()) * g2
    d = np.abs(hc - hr).reshape(-1, g2)
    scale = np.maximum(np.abs(hr).reshape(-1, g2).max(axis=1), 1e-300)
    ident = same_set[nonempty]

    def fit_err(r):
        """per-patch sum of squared residuals at the patch's own training points: f = sum_i alpha_i k(x, BV_i)"""
        sse = np.zeros(P)
        for p in np.nonzero(nonempty)[0]:
            a0, a1 = r["bv_off"][p], r["bv_off"][p + 1]
            s0, s1 = poff[p], poff[p + 1]
            d2 = (sx1[s0:s1, None] - r["bv1"][None, a0:a1]) ** 2 + (sx2[s0:s1, None] - r["bv2"][None, a0:a1]) ** 2
            f = (100.0 * np.exp(-0.5 * d2)) @ r["alpha"][a0:a1]
            sse[p] = ((f - sy[s0:s1]) ** 2).sum()
        return sse
    sse_c, sse_r = fit_err(rc), fit_err(rr)
    npts = float(poff[-1])
    dm = d.max(axis=1)
    rep = {
        "rmse_m_canonical": float(np.sqrt(sse_c.sum() / npts)), "rmse_m_reference_order": float(np.sqrt(sse_r.sum() / npts)),
        "patches_dheight_gt_1mm": int((dm > 1e-3).sum()), "patches_dheight_gt_1cm": int((dm > 1e-2).sum()),
        "patches_dheight_gt_10cm": int((dm > 1e-1).sum()),
        "worst_patch_rmse_m": [float(np.sqrt(sse_c.max() / max(1, np.diff(poff)[sse_c.argmax()]))),
                               float(np.sqrt(sse_r.max() / max(1, np.diff(poff)[sse_r.argmax()])))],
        "workload": a.workload, "points": int(a.points), "patches": int(P), "nonempty": int(nonempty.sum()),
        "frac_same_bv_count": float(same_n[nonempty].mean()), "frac_same_bv_set": float(ident.mean()),
        "mean_bv_canonical": float(rc["nbv"][nonempty].mean()), "mean_bv_reference_order": float(rr["nbv"][nonempty].mean()),
        "max_abs_dheight_m_all": float(d.max()), "max_abs_dheight_m_same_set": float(d[ident].max()) if ident.any() else None,
        "p99_abs_dheight_m_all": float(np.quantile(d.max(axis=1), 0.99)),
        "max_rel_dheight_same_set": float((d[ident].max(axis=1) / scale[ident]).max()) if ident.any() else None,
        "median_rel_dheight_same_set": float(np.median(d[ident].max(axis=1) / scale[ident])) if ident.any() else None,
        "seconds": round(time.time() - t0, 1),
    }
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
