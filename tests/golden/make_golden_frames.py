"""Generates tests/golden/ref_frames.npz: patch frames and greedy claims produced by THE REFERENCE'S OWN
gp_compressor::compute_rotation and ::project_points (gp_compressor.cpp:29-118, compiled from /root/reference by
oracle/ref_build.py::build_frames over oracle/eigen_shim with a one-sided-Jacobi JacobiSVD stand-in), driven over small
clouds by the sequential loop of project_cloud (gp_compressor.cpp:204-243) as restated in tests/binning_numpy.py
(lattice, leaf order and brute-force float32 radius search follow the [RECALLED] PCL semantics of SURVEY.md 8a).
Run in the container that has /root/reference:  python tests/golden/make_golden_frames.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import binning_numpy as B  # noqa: E402
from gp_compressor_b200 import synth  # noqa: E402
from oracle import ref_source as R  # noqa: E402


def clouds():
    F = np.float32
    yield "dense_floor_res015", synth.c3_dense_floor(4000, seed=5, side=1.2), float(F(0.15)), 0   # ~60 points per patch
    rng = np.random.default_rng(6)
    # a room corner: floor + two walls inside a 0.9 m box, 1400 points each, 2 mm noise: z-, x- and y-dominant normals
    u, v, e = rng.uniform(0, 0.9, (3, 4200)), rng.uniform(0, 0.9, (3, 4200)), rng.normal(0, 0.002, (3, 4200))
    corner = np.concatenate([np.stack([u[0], v[0], e[0]], 1)[:1400], np.stack([e[1], u[1], v[1]], 1)[:1400],
                             np.stack([u[2], e[2], v[2]], 1)[:1400]]) + [2.0, -1.0, 0.5]
    corner = corner[rng.permutation(corner.shape[0])]
    yield "corner_res01", synth.pack_cloud(corner, rng.integers(0, 256, (corner.shape[0], 3))), float(F(0.1)), 0
    yield "corner_morton_res02", synth.pack_cloud(corner[:2500], rng.integers(0, 256, (2500, 3))), float(F(0.2)), 1   # PCL >= 1.9 leaf order
    yield "c2_indoor_sparse", synth.c2_indoor(3000, seed=6), float(F(0.1)), 0          # ~1 point per leaf: identity frames
    rng = np.random.default_rng(9)
    xyz = rng.uniform(0, 1.2, (1500, 3)) * [1, 1, 0.04]
    xyz[100:140] = xyz[0:40]                 # exact duplicates
    xyz[200] = [np.nan, 0.5, 0.0]            # dropped by addPointsFromInputCloud
    xyz[201:204] += [30.0, 0.0, 0.0]         # an isolated leaf with fewer than 4 candidates: identity frame (:31-34)
    yield "edge_cases", synth.pack_cloud(xyz, rng.integers(0, 256, (1500, 3))), float(F(0.15)), 0


def main():
    assert R.frames_available(), "needs /root/reference (or a prebuilt oracle/_ref/libref_frames.so)"
    out = {}
    for name, cloud, res, order in clouds():
        xyz = cloud[:, :12].copy().view(np.float32).reshape(-1, 3)
        bgr = cloud[:, 16:19].astype(np.float64)
        rgb = bgr[:, ::-1].copy()            # colours as (r, g, b) doubles, gp_compressor.cpp:233-235
        w = B.project_cloud(xyz, res, leaf_order=order, ref=R, rgb=rgb, sz=10)
        n_p = np.array([len(s) for s in w["stream"]], dtype=np.int64)
        out[name + "/cloud"] = cloud
        out[name + "/meta"] = np.array([order, w["depth"]], dtype=np.int64)
        out[name + "/res"] = np.array([res])
        out[name + "/lattice_min"] = w["lattice_min"]
        out[name + "/leaf_code"] = w["leaf_code"].astype(np.uint64)
        out[name + "/ncand"] = np.array(w["ncand"], dtype=np.int32)
        out[name + "/R"] = np.array(w["R"])
        out[name + "/owner"] = w["owner"]
        out[name + "/patch_off"] = np.concatenate([[0], np.cumsum(n_p)])
        out[name + "/stream"] = np.concatenate(w["stream"]).astype(np.int32)
        out[name + "/x1"] = np.concatenate(w["x1"])
        out[name + "/x2"] = np.concatenate(w["x2"])
        out[name + "/y"] = np.concatenate(w["y"])
        out[name + "/center"] = np.array(w["center"])
        out[name + "/rgb_mean"] = np.array(w["rgb_mean"])
        out[name + "/colour"] = np.concatenate(w["colour"])
        print(name, "leaves", len(n_p), "claimed", int(n_p.sum()), "of", xyz.shape[0])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_frames.npz"), **out)


if __name__ == "__main__":
    main()
