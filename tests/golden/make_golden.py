"""Generates tests/golden/ref_sogp.npz from THE REFERENCE'S OWN SOURCE (oracle/_ref, see oracle/ref_build.py):
inputs and outputs of sparse_gp<rbf_kernel,gaussian_noise>::add_measurements / predict_measurements for a
few seeded patches.  Run in the container that has /root/reference:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_source as R  # noqa: E402

CASES = [  # name, n, capacity, hyper-parameters
    ("ref_defaults_cap100", 300, 100, dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1)))),
    ("ref_defaults_cap30", 500, 30, dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1)))),
    ("bind_cap12", 400, 12, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
    ("bind_cap30", 600, 30, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
    ("bind_cap60", 900, 60, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
    ("tiny_n3", 3, 100, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
]


def main():
    assert R.available(), "needs /root/reference (or a prebuilt oracle/_ref)"
    out = {}
    for k, (name, n, cap, hyp) in enumerate(CASES):
        rng = np.random.default_rng(100 + k)
        x1 = rng.uniform(-0.05, 0.05, n)
        x2 = rng.uniform(-0.05, 0.05, n)
        y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + 0.5 * x1 + rng.normal(0, 0.003, n)
        pred = rng.uniform(-0.05, 0.05, (25, 2))
        r = R.fit(x1, x2, y, capacity=cap, rand_offset=7 * k, pred=pred, **hyp)
        for key, v in dict(x1=x1, x2=x2, y=y, pred=pred, alpha=r["alpha"], bv1=r["bv1"], bv2=r["bv2"], C=r["C"], Q=r["Q"], f=r["f"],
                           sigma=r["sigma"], meta=np.array([n, cap, 7 * k, r["N"]], dtype=np.int64),
                           hyper=np.array([hyp["sigmaf_sq"], hyp["l_sq"], hyp["s0"]])).items():
            out[f"{name}/{key}"] = v
    # RGB field GP (sparse_gp_field.hpp) of the reference: defaults (capacity never binds) and a capacity-bound case
    for name, n, cap, hyp in [("field_ref_defaults", 400, 100, dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e2)))),
                              ("field_cap8", 300, 8, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-2))]:
        rng = np.random.default_rng(len(name))
        x1 = rng.uniform(-0.05, 0.05, n)
        x2 = rng.uniform(-0.05, 0.05, n)
        Y = np.stack([50 * np.sin(40 * x1), 30 * np.cos(30 * x2), 20 * np.sin(30 * (x1 + x2))], axis=1) + rng.normal(0, 1, (n, 3))
        pred = rng.uniform(-0.05, 0.05, (25, 2))
        r = R.field_fit(x1, x2, Y, capacity=cap, rand_offset=n - 1, pred=pred, eps_tol=float(np.float32(1e-4)), **hyp)
        for key, v in dict(x1=x1, x2=x2, Y=Y, pred=pred, alpha=r["alpha"], bv1=r["bv1"], bv2=r["bv2"], f=r["f"],
                           meta=np.array([n, cap, n - 1, r["N"]], dtype=np.int64),
                           hyper=np.array([hyp["sigmaf_sq"], hyp["l_sq"], hyp["s0"]])).items():
            out[f"{name}/{key}"] = v
    # rows N2 / N4: the reference's predict_measurements (sigma, conf), compute_likelihoods and compute_derivatives
    for name, n, cap, hyp in [("eval_cap12", 300, 12, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
                              ("eval_cap40", 700, 40, dict(sigmaf_sq=2.0, l_sq=4e-4, s0=1e-3)),
                              ("eval_empty", 0, 10, dict(sigmaf_sq=1.5, l_sq=1e-3, s0=1e-2))]:
        rng = np.random.default_rng(len(name) + n)
        x1 = rng.uniform(-0.05, 0.05, n)
        x2 = rng.uniform(-0.05, 0.05, n)
        fn = lambda a, b: 0.02 * np.sin(40 * a) * np.cos(30 * b) + 0.5 * a
        y = fn(x1, x2) + rng.normal(0, 0.003, n)
        ex = rng.uniform(-0.05, 0.05, (40, 2))
        ey = fn(ex[:, 0], ex[:, 1]) + rng.normal(0, 0.01, 40)
        r = R.evaluate(x1, x2, y, ex, ey, capacity=cap, rand_offset=3, **hyp)
        for key, v in dict(x1=x1, x2=x2, y=y, ex=ex, ey=ey, f=r["f"], sigma=r["sigma"], conf=r["conf"], lik=r["lik"], dX=r["dX"],
                           meta=np.array([n, cap, 3, r["N"]], dtype=np.int64),
                           hyper=np.array([hyp["sigmaf_sq"], hyp["l_sq"], hyp["s0"]])).items():
            out[f"{name}/{key}"] = v
    # RGB field GP: the reference's predict_measurements (sigma, conf), compute_likelihoods, compute_derivatives
    for name, n, cap, s0 in [("feval_cap8", 200, 8, 1e-2), ("feval_nodel", 60, 100, 1e-1)]:   # capacity-bound (the divergent delete_bv) / never deleting
        rng = np.random.default_rng(len(name) + n)
        x1 = rng.uniform(-0.05, 0.05, n)
        x2 = rng.uniform(-0.05, 0.05, n)
        Y = np.stack([50 * np.sin(40 * x1), 30 * np.cos(30 * x2), 20 * np.sin(30 * (x1 + x2))], axis=1) + rng.normal(0, 1, (n, 3))
        ex = rng.uniform(-0.05, 0.05, (30, 2))
        EY = np.stack([50 * np.sin(40 * ex[:, 0]), 30 * np.cos(30 * ex[:, 1]), 20 * np.sin(30 * ex.sum(1))], axis=1) + rng.normal(0, 3, (30, 3))
        hyp = dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2)
        r = R.field_evaluate(x1, x2, Y, ex, EY, capacity=cap, s0=s0, rand_offset=n - 1, **hyp)
        for key, v in dict(x1=x1, x2=x2, Y=Y, ex=ex, EY=EY, f=r["f"], sigma=r["sigma"], conf=r["conf"], lik=r["lik"], dX=r["dX"],
                           meta=np.array([n, cap, n - 1, r["N"]], dtype=np.int64),
                           hyper=np.array([hyp["sigmaf_sq"], hyp["l_sq"], s0])).items():
            out[f"{name}/{key}"] = v
    # add_measurements called twice on one process (the reference accumulates)
    for name, n, n1, cap in [("cont_cap25", 400, 150, 25), ("cont_cap12", 300, 5, 12)]:
        rng = np.random.default_rng(len(name) + n)
        x1 = rng.uniform(-0.05, 0.05, n)
        x2 = rng.uniform(-0.05, 0.05, n)
        y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
        hyp = dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)
        r = R.fit_twice(x1, x2, y, n1, capacity=cap, rand_offset=3, **hyp)
        for key, v in dict(x1=x1, x2=x2, y=y, alpha=r["alpha"], bv1=r["bv1"], bv2=r["bv2"],
                           meta=np.array([n, cap, 3, r["N"], n1], dtype=np.int64),
                           hyper=np.array([hyp["sigmaf_sq"], hyp["l_sq"], hyp["s0"]])).items():
            out[f"{name}/{key}"] = v
    out["shuffle_n57_off3"] = R.shuffle(57, 3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_sogp.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
