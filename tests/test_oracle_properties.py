"""Property tests of the CPU oracle (hypothesis): the invariants SURVEY.md section 4 item 3 lists, on random patches
and random clouds.  They guard the checker itself; the GPU is compared against it bit for bit elsewhere."""
import numpy as np
from hypothesis import given, settings, strategies as st

from gp_compressor_b200 import synth

F32 = lambda v: float(np.float32(v))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 300), cap=st.integers(1, 40), shuffle=st.booleans())
def test_sogp_state_invariants(oracle_mod, seed, n, cap, shuffle):
    rng = np.random.default_rng(seed)
    res = 0.1
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
    hyp = synth.hyper_bind(res)
    o = oracle_mod.Oracle(capacity=cap, shuffle=int(shuffle), **hyp)
    r = o.fit_patches([0, n], x1, x2, y, dump=True)
    N = int(r["nbv"][0])
    assert 1 <= N <= min(cap, n)
    Cm, Qm = r["C"].reshape(N, N), r["Q"].reshape(N, N)
    assert np.array_equal(Cm, Cm.T) and np.array_equal(Qm, Qm.T)           # bitwise symmetric canonical updates
    assert len(set(r["bv_idx"].tolist())) == N                              # BVs are distinct input points
    assert np.array_equal(r["bv1"], x1[r["bv_idx"]]) and np.array_equal(r["bv2"], x2[r["bv_idx"]])
    b = np.stack([r["bv1"], r["bv2"]])
    d = b[:, :, None] - b[:, None, :]
    K = hyp["sigmaf_sq"] * np.exp(-0.5 / hyp["l_sq"] * (d * d).sum(axis=0))
    assert np.abs(Qm @ K - np.eye(N)).max() < 1e-5                          # Q is the inverse Gram matrix of the BVs
    assert sorted(r["perm"].tolist()) == list(range(n))                     # the shuffle is a permutation
    stt = o.stats()
    assert stt["n_first"] == 1 and stt["n_first"] + stt["n_sparse"] + stt["n_full"] == n
    assert stt["n_full"] - stt["n_del_cap"] - stt["n_del_geo"] == N - 1


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 1000), n=st.integers(1, 4000), order=st.integers(0, 1))
def test_binning_invariants(oracle_mod, seed, n, order):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-2, 2, (n, 3)) * [1, 1, 0.1]
    xyz[rng.random(n) < 0.02] = np.nan
    cloud = synth.pack_cloud(xyz, rng.integers(0, 256, (n, 3)).astype(np.uint8))
    res = F32(0.3)
    o = oracle_mod.Oracle(res=res, leaf_order=order, capacity=5)
    b = o.project(cloud)
    off = b["patch_off"]
    P = b["n_leaves"]
    finite = np.isfinite(xyz).all(axis=1)
    # claimed at most once, only finite points, owner consistent with the stream
    assert np.unique(b["st_idx"]).size == b["st_idx"].size
    assert finite[b["st_idx"]].all() and (b["owner"][~finite] == -1).all()
    assert np.array_equal(b["owner"][b["st_idx"]], np.repeat(np.arange(P), np.diff(off)))
    # visiting order is monotone in the Morton code (descending for the PCL 1.7 iterator, ascending otherwise)
    code = b["leaf_code"].astype(np.int64)
    assert np.all(np.diff(code) < 0) if order == 0 else np.all(np.diff(code) > 0)
    # every leaf holds at least one finite point, so there are at most as many leaves as finite points
    assert 0 < P <= finite.sum() or finite.sum() == 0
    # local coordinates respect the closed patch box; rotations are orthonormal
    half = res / 2.0
    assert np.all(np.abs(b["st_x1"]) <= half) and np.all(np.abs(b["st_x2"]) <= half)
    if P > 0:
        R = b["leaf_R"].reshape(P, 3, 3)
        assert np.abs(np.einsum("pij,pkj->pik", R, R) - np.eye(3)).max() < 1e-12


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(2, 250), cut=st.floats(0.0, 1.0), cap=st.integers(1, 40))
def test_continued_fit_equals_one_fit_without_shuffle(oracle_mod, seed, n, cut, cap):
    """sparse_gp::add_measurements accumulates (sparse_gp.hpp:59-86): with the shuffle off, feeding the points in two calls
    gives exactly the state of one call, whatever the cut."""
    rng = np.random.default_rng(seed)
    res = 0.1
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
    hyp = synth.hyper_bind(res)
    n1 = int(round(cut * n))
    a = oracle_mod.Oracle(capacity=cap, shuffle=0, **hyp)
    ra = a.fit_patches([0, n], x1, x2, y, dump=True)
    b = oracle_mod.Oracle(capacity=cap, shuffle=0, **hyp)
    b.fit_patches([0, n1], x1[:n1], x2[:n1], y[:n1], dump=True)
    rb = b.add_measurements([0, n - n1], x1[n1:], x2[n1:], y[n1:])
    for k in ("nbv", "bv_idx", "bv1", "bv2", "alpha", "C", "Q"):
        assert np.array_equal(ra[k], rb[k]), k


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(0, 200), cap=st.integers(1, 30), m=st.integers(1, 40))
def test_evaluate_invariants(oracle_mod, seed, n, cap, m):
    """Rows N2 / N4: sigma = sqrt(variance) >= 0 and never above the prior's, confidence in [0, 100], likelihood a
    density value, the mean equals predict, and at a BV with tiny noise the fit interpolates."""
    rng = np.random.default_rng(seed)
    res = 0.1
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
    hyp = synth.hyper_bind(res)
    o = oracle_mod.Oracle(capacity=cap, **hyp)
    o.fit_patches([0, n], x1, x2, y, dump=True)
    q1 = rng.uniform(-res / 2, res / 2, m)
    q2 = rng.uniform(-res / 2, res / 2, m)
    qy = rng.normal(0, 0.02, m)
    e = o.evaluate([0, m], q1, q2, qy)
    c = o.evaluate([0, m], q1, q2, qy, conf=True)
    prior = np.sqrt(hyp["sigmaf_sq"] + hyp["s0"])
    assert np.all(e["sigma"] >= 0) and np.all(e["sigma"] <= prior * (1 + 1e-12))
    assert np.all(c["sigma"] >= -1e-9) and np.all(c["sigma"] <= 100.0)
    assert np.all(np.isfinite(e["lik"])) and np.all(e["lik"] >= 0)
    assert np.all(np.isfinite(e["dX"]))
    assert np.array_equal(e["f"], o.predict(0, np.stack([q1, q2], 1)))
    if n == 0:
        assert np.all(e["f"] == 0) and np.all(e["sigma"] == prior) and np.all(c["sigma"] == 0)
