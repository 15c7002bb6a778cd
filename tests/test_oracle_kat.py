"""Pins for the CPU oracle that do not need a GPU: glibc rand()/shuffle against the real
libc of this image, canonical exp against libm, the SOGP recursion against the independent
numpy restatement of matlab/sogp.m, and the known-answer vectors of SURVEY.md section 4."""
import ctypes
import math

import numpy as np
import pytest

from sogp_numpy import SogpNumpy


def test_rand_stream_matches_real_glibc(oracle_mod):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)  # unseeded rand() == srand(1)
    real = np.array([libc.rand() for _ in range(20000)], dtype=np.uint32)
    assert np.array_equal(oracle_mod.rand_stream(0, 20000), real)
    assert np.array_equal(oracle_mod.rand_stream(777, 100), real[777:877])
    assert real[:4].tolist() == [1804289383, 846930886, 1681692777, 1714636915]


def test_shuffle_kat_and_real_glibc(oracle_mod):
    got = oracle_mod.shuffles(0, [1, 2, 5, 8])
    assert got.tolist() == [0, 1, 0, 4, 3, 1, 0, 2, 7, 5, 0, 6, 1, 2, 4, 3]
    # sparse_gp.hpp:42-56 run over the real libc rand()
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    sizes = [3, 1, 40, 17, 2, 256]
    want = []
    for n in sizes:
        ind = list(range(n))
        for i in range(n - 1, 0, -1):
            r = libc.rand() % i
            ind[i], ind[r] = ind[r], ind[i]
        want += ind
    assert oracle_mod.shuffles(0, sizes).tolist() == want


def test_exp_within_one_ulp_of_libm(oracle_mod):
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 60, 200000), -rng.uniform(0, 1e-2, 200000), -rng.uniform(0, 1e-6, 50000),
                        rng.uniform(0, 5, 50000), np.array([0.0, -0.0, -1e-300, -700.0, -744.0])])
    got = oracle_mod.exp_array(x)
    ref = np.array([math.exp(v) for v in x])  # libm (np.exp is numpy's own SIMD kernel)
    ulp = np.abs(got.view(np.int64) - ref.view(np.int64))
    assert ulp.max() <= 1
    assert (ulp > 0).mean() < 0.06
    assert oracle_mod.exp_array(np.array([0.0]))[0] == 1.0
    assert oracle_mod.exp_array(np.array([-800.0]))[0] == 0.0
    assert math.isnan(oracle_mod.exp_array(np.array([np.nan]))[0])


def test_float_literal_constants():
    assert float(np.float32(1e-6)) == 9.9999999747524271e-07
    assert float(np.float32(1e-12)) == 9.999999960041972e-13
    assert float(np.float32(1e-9)) == 9.9999997171806854e-10
    assert float(np.float32(1e-1)) == 0.10000000149011612
    assert float(np.float32(np.sqrt(np.float32(3.0))) / np.float32(2.0)) * float(np.float32(0.15)) == 0.12990381339803569


def test_sogp_survey_smoke_vector(oracle_mod):
    o = oracle_mod.Oracle(capacity=2, shuffle=0)
    r = o.fit_patches([0, 4], [0, .05, 0, .05], [0, 0, .05, .05], [.01, .02, -.01, 0])
    assert r["nbv"].tolist() == [2]
    assert r["bv_idx"].tolist() == [2, 1]
    np.testing.assert_allclose(r["alpha"], [-0.04287038525668329, 0.04292043524805538], rtol=1e-9)
    np.testing.assert_allclose(o.predict(0, [[0.025, 0.025]])[0], 0.0050018719900836684, rtol=1e-8)
    st = o.stats()
    assert (st["n_full"], st["n_del_cap"], st["n_sparse"], st["n_first"]) == (3, 2, 0, 1)


def _patch(rng, n, res=0.1):
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
    return x1, x2, y


@pytest.mark.parametrize("cap,hyper", [(100, dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(0.1)))),
                                      (12, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)),
                                      (30, dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4))])
def test_sogp_matches_numpy_witness(oracle_mod, cap, hyper):
    """Oracle (canonical arithmetic) vs numpy restatement of matlab/sogp.m: same BV sets and
    event sequences, alpha to rounding level scaled by the conditioning of the hyper-set."""
    rng = np.random.default_rng(cap)
    n = 300
    x1, x2, y = _patch(rng, n)
    o = oracle_mod.Oracle(capacity=cap, shuffle=0, **hyper)
    r = o.fit_patches([0, n], x1, x2, y, dump=True)
    w = SogpNumpy(capacity=cap, s20=hyper["s0"], sigmaf_sq=hyper["sigmaf_sq"], l_sq=hyper["l_sq"])
    for i in range(n):
        w.add([x1[i], x2[i]], y[i], i)
    st = o.stats()
    well_conditioned = hyper["l_sq"] < 1.0
    if well_conditioned:
        assert sorted(r["bv_idx"].tolist()) == sorted(w.idx)
        assert r["bv_idx"].tolist() == w.idx  # same slot order too
        assert st["n_full"] == w.events.count("full") and st["n_sparse"] == w.events.count("sparse")
        assert st["n_del_cap"] == w.events.count("delcap") and st["n_del_geo"] == w.events.count("delgeo")
        np.testing.assert_allclose(r["alpha"], w.alpha, rtol=1e-7, atol=1e-9)
        N = int(r["nbv"][0])
        np.testing.assert_allclose(r["C"].reshape(N, N), w.C, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r["Q"].reshape(N, N), w.Q, rtol=1e-6, atol=1e-6)
    # predictions agree in both regimes (the ill-conditioned REF set hides BV-set noise)
    X = np.stack(_patch(rng, 50)[:2], axis=1)
    f = o.predict(0, X)
    fw = np.array([w.predict(x) for x in X])
    np.testing.assert_allclose(f, fw, atol=2e-4 if not well_conditioned else 1e-8)


def test_sogp_invariants(oracle_mod):
    """Q K_BV ~ I, symmetry of C and Q, size <= capacity."""
    rng = np.random.default_rng(5)
    n, cap, res = 400, 20, 0.1
    x1, x2, y = _patch(rng, n)
    hyper = dict(sigmaf_sq=1.0, l_sq=(res / 12) ** 2, s0=1e-4)
    o = oracle_mod.Oracle(capacity=cap, shuffle=1, **hyper)
    r = o.fit_patches([0, n], x1, x2, y, dump=True)
    N = int(r["nbv"][0])
    assert 1 <= N <= cap
    Cm, Qm = r["C"].reshape(N, N), r["Q"].reshape(N, N)
    assert np.array_equal(Cm, Cm.T) and np.array_equal(Qm, Qm.T)  # canonical updates are bitwise symmetric
    b = np.stack([r["bv1"], r["bv2"]])
    d = b[:, :, None] - b[:, None, :]
    K = np.exp(-0.5 / hyper["l_sq"] * (d * d).sum(axis=0))
    np.testing.assert_allclose(Qm @ K, np.eye(N), atol=1e-6)
    # BVs are points of the patch, addressed through the shuffle permutation
    assert np.array_equal(r["bv1"], x1[r["bv_idx"]]) and np.array_equal(r["bv2"], x2[r["bv_idx"]])
    perm = r["perm"]
    assert sorted(perm.tolist()) == list(range(n))
    assert perm.tolist() == oracle_mod.shuffles(0, [n]).tolist()


def test_rand_stream_accounting_across_patches(oracle_mod):
    """Patch p's height shuffle starts after 2*(n_q - 1) draws per earlier non-empty patch
    (gp_compressor.cpp:162-163: height GP then RGB field GP both shuffle)."""
    rng = np.random.default_rng(9)
    sizes = [5, 0, 9, 1, 7]
    off = np.concatenate([[0], np.cumsum(sizes)])
    x1, x2, y = _patch(rng, off[-1])
    o = oracle_mod.Oracle(capacity=4)
    r = o.fit_patches(off, x1, x2, y)
    pos = 0
    for p, n in enumerate(sizes):
        if n == 0:
            continue
        want = oracle_mod.shuffles(pos, [n])
        assert r["perm"][off[p]:off[p + 1]].tolist() == want.tolist()
        pos += 2 * (n - 1)
    assert o.rand_offset() == pos
    assert r["nbv"][1] == 0


def test_separable_grid_decode_matches_direct_kernel(oracle_mod):
    """Grid decode evaluates the RBF separably, k = (p0 e^{cl dx^2}) e^{cl dy^2}; the reference computes
    p0 exp(cl (dx^2 + dy^2)) with libm (rbf_kernel.cpp:15-18, sparse_gp.hpp:320-327).  Pin the distance: a kernel
    value with exponent argument z differs by at most (3 + 2|z|) ulp -- the conditioning of exp at z, which the direct
    form's own rounding of z has too -- and heights by < 1e-12 of sum|alpha_i| p0 (tolerance of the path: 1e-9)."""
    rng = np.random.default_rng(11)
    res, sz = float(np.float32(0.15)), 20
    p0, l_sq, s0 = 1.0, (res / 6.0) ** 2, 1e-3
    P, npts = 6, 120
    off = np.arange(P + 1, dtype=np.int64) * npts
    x1 = rng.uniform(-res / 2, res / 2, P * npts)
    x2 = rng.uniform(-res / 2, res / 2, P * npts)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + 1e-4 * rng.standard_normal(P * npts)
    o = oracle_mod.Oracle(res=res, sz=sz, capacity=25, sigmaf_sq=p0, l_sq=l_sq, s0=s0, decode_separable=1)  # the flagged fast mode
    o.fit_patches(off, x1, x2, y)
    r = o.fit_result()
    _, heights = o.decode(want_cloud=False)
    heights = heights.reshape(P, sz, sz)
    grid = res * ((np.arange(sz, dtype=np.float64) + 0.5) / sz - 0.5)          # gp_compressor.cpp:326-327
    cl = float(np.float32(-0.5)) / l_sq
    bo = np.concatenate([[0], np.cumsum(r["nbv"])])
    worst_ulp, worst_h = 0.0, 0.0
    for p in range(P):
        sl = slice(bo[p], bo[p + 1])
        b1, b2, al = r["bv1"][sl], r["bv2"][sl], r["alpha"][sl]
        assert len(al) >= 5
        ref = np.zeros((sz, sz))
        for yy in range(sz):
            for xx in range(sz):
                acc = 0.0
                for i in range(len(al)):
                    d1, d2 = grid[xx] - b1[i], grid[yy] - b2[i]
                    kd = p0 * math.exp(cl * (d1 * d1 + d2 * d2))
                    ks = (p0 * math.exp(cl * (d1 * d1))) * math.exp(cl * (d2 * d2))
                    z = abs(cl * (d1 * d1 + d2 * d2))
                    worst_ulp = max(worst_ulp, abs(ks - kd) / math.ulp(kd) / (3.0 + 2.0 * z))
                    acc += al[i] * kd
                ref[yy, xx] = acc
        scale = np.abs(al).sum() * p0
        worst_h = max(worst_h, float(np.abs(heights[p] - ref).max() / scale))
    assert worst_ulp <= 1.0
    assert worst_h < 1e-12
    # and the point-wise predict (direct form) agrees with the grid decode to the same level
    f = o.predict(0, np.stack(np.meshgrid(grid, grid), -1).reshape(-1, 2))
    assert np.abs(f - heights[0].ravel()).max() <= 1e-12 * np.abs(r["alpha"][:r["nbv"][0]]).sum() * p0


@pytest.mark.parametrize("separable", [0, 1])
def test_grid_decode_modes_under_reference_hyperparameters(oracle_mod, separable):
    """Both decode modes under the REFERENCE DEFAULTS (rbf 100 / 1, s0 1e-1f: sum|alpha| p0 >> |f|, the ill-conditioned
    case the bench runs) against the reference expression evaluated in numpy with libm: f* = sum_i alpha_i p0
    exp(cl ((X0-b1)^2 + (X1-b2)^2)) on the lattice of gp_compressor.cpp:320-328 (rbf_kernel.cpp:15-18, sparse_gp.hpp:320-327).
    Contract of the path: 1e-9 relative to max|f| of the patch.  Direct mode (the default) meets it with orders of
    magnitude to spare; the separable fast mode is measured with the same yardstick."""
    rng = np.random.default_rng(23)
    res, sz = float(np.float32(0.1)), 10
    P, npts = 8, 400
    off = np.arange(P + 1, dtype=np.int64) * npts
    x1 = rng.uniform(-res / 2, res / 2, P * npts)
    x2 = rng.uniform(-res / 2, res / 2, P * npts)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + 0.3 * x1 + 0.003 * rng.standard_normal(P * npts)
    o = oracle_mod.Oracle(res=res, sz=sz, capacity=30, decode_separable=separable)  # reference defaults
    o.fit_patches(off, x1, x2, y)
    r = o.fit_result()
    _, heights = o.decode(want_cloud=False)
    heights = heights.reshape(P, sz * sz)
    grid = res * ((np.arange(sz, dtype=np.float64) + np.float32(0.5)) / sz - np.float32(0.5))
    X0, X1 = (g.ravel() for g in np.meshgrid(grid, grid))
    cl = float(np.float32(-0.5)) / 1.0
    bo = np.concatenate([[0], np.cumsum(r["nbv"])])
    worst = 0.0
    for p in range(P):
        sl = slice(bo[p], bo[p + 1])
        b1, b2, al = r["bv1"][sl], r["bv2"][sl], r["alpha"][sl]
        K = 100.0 * np.exp(cl * ((X0[:, None] - b1[None, :]) ** 2 + (X1[:, None] - b2[None, :]) ** 2))
        want = np.array([math.fsum(K[m] * al) for m in range(sz * sz)])   # exactly rounded sum: the yardstick
        worst = max(worst, float(np.abs(heights[p] - want).max() / np.abs(want).max()))
        assert np.abs(al).sum() * 100.0 > 1e3 * np.abs(want).max()        # the cancellation this test is about
    assert worst < 1e-9, worst


@pytest.mark.parametrize("seed,res,order", [(1, 0.15, 0), (2, 0.1, 1), (3, 0.3, 0), (4, 0.07, 0)])
def test_binning_matches_sequential_numpy_witness(oracle_mod, seed, res, order):
    """Rows a-1 .. a-5: the oracle's order-free restatement of project_cloud against an independent, sequential numpy
    restatement (tests/binning_numpy.py: point-by-point bounding-box growth, brute-force float32 radius search, numpy SVD
    for the plane normal, greedy claim with an occupied array).  Lattice, leaves, visiting order, candidate counts, owners
    and per-patch point order must be identical; local coordinates agree to the rounding of the two SVDs."""
    from binning_numpy import project_cloud
    from gp_compressor_b200 import synth
    rng = np.random.default_rng(seed)
    n = 1500
    ext = np.array([1.2, 0.9, 0.12])
    xyz = (rng.uniform(-1, 1, (n, 3)) * ext + rng.uniform(-5, 5, 3)).astype(np.float32)
    xyz[:, 2] += (0.05 * np.sin(3 * xyz[:, 0]) * np.cos(2 * xyz[:, 1])).astype(np.float32)
    xyz[rng.random(n) < 0.01] = np.nan
    cloud = synth.pack_cloud(xyz, rng.integers(0, 256, (n, 3)).astype(np.uint8))
    resf = float(np.float32(res))
    o = oracle_mod.Oracle(res=resf, leaf_order=order, capacity=5)
    b = o.project(cloud)
    w = project_cloud(xyz, resf, leaf_order=order)
    assert int(b["depth"]) == w["depth"]
    assert np.array_equal(b["lattice_min"], w["lattice_min"])
    assert np.array_equal(b["leaf_code"].astype(np.uint64), w["leaf_code"])
    assert np.array_equal(b["leaf_ncand"], np.array(w["ncand"]))
    assert np.array_equal(b["owner"], w["owner"])
    off = b["patch_off"]
    for p in range(len(w["stream"])):
        sl = slice(off[p], off[p + 1])
        assert np.array_equal(b["st_idx"][sl], w["stream"][p]), p
        if w["stream"][p].size:
            assert np.abs(b["st_x1"][sl] - w["x1"][p]).max() < 1e-9 and np.abs(b["st_x2"][sl] - w["x2"][p]).max() < 1e-9
            assert np.abs(b["st_y"][sl] - w["y"][p]).max() < 1e-9
            Rw = w["R"][p]
            assert np.abs(b["leaf_R"].reshape(-1, 3, 3)[p] - Rw).max() < 1e-7


def test_decode_frames_match_numpy(oracle_mod):
    """load_compressed (gp_compressor.cpp:320-371): every decoded point is R [f, X0, X1]' + mean on the sz x sz lattice,
    with R the patch rotation (through its stored quaternion) and the patch's mean colour; recomputed here in numpy from
    the oracle's own frames and heights, independent of its quaternion code."""
    from gp_compressor_b200 import synth
    cloud = synth.c1_planar_bumps(6000, seed=5)
    res, sz = float(np.float32(0.15)), 6
    o = oracle_mod.Oracle(res=res, sz=sz, capacity=20)
    o.compress(cloud)
    b = o.binning()
    dec, heights = o.decode()
    nbv = o.fit_result()["nbv"]
    keep = np.nonzero(nbv > 0)[0]
    assert dec.shape[0] == keep.size * sz * sz
    grid = res * ((np.arange(sz, dtype=np.float64) + np.float32(0.5)) / sz - np.float32(0.5))
    X0, X1 = np.meshgrid(grid, grid)                       # y outer, x inner (:320-328)
    R = b["leaf_R"].reshape(-1, 3, 3)[keep]
    mean = b["leaf_mean"].reshape(-1, 3)[keep]
    f = heights.reshape(keep.size, sz * sz)
    local = np.stack([f, np.broadcast_to(X0.ravel(), f.shape), np.broadcast_to(X1.ravel(), f.shape)], axis=2)
    want = np.einsum("pij,pmj->pmi", R, local) + mean[:, None, :]
    got = dec[:, :12].copy().view(np.float32).reshape(keep.size, sz * sz, 3)
    assert np.abs(got - want).max() < 2e-6 * max(1.0, np.abs(want).max())
    rgb = dec[:, 16:19].reshape(keep.size, sz * sz, 3)
    cm = np.clip(b["leaf_rgbmean"].reshape(-1, 3)[keep].astype(np.int64), 0, 255)     # flatten_colors on the mean colour
    assert np.array_equal(rgb[:, 0, ::-1], cm)                                          # stored b, g, r
    assert (rgb == rgb[:, :1, :]).all()
