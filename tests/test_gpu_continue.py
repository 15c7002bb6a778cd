"""GPU parity tests of gpc_add_measurements (sparse_gp::add_measurements called again on fitted processes,
sparse_gp.hpp:59-86; gp_mapping.cpp:338): the kept state of every patch is resumed by the bucket that reads its slot format
and the recursion continues on the new points.  Bit-equal to the CPU oracle; the oracle's continuation is pinned against the
reference's own source (two successive add_measurements calls) in test_oracle_continue_matches_reference_source."""
import numpy as np
import pytest


def bind(res=0.1):
    return dict(sigmaf_sq=1.0, l_sq=(res / 12.0) ** 2, s0=1e-4)


REF = dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1)))


def make(seed, sizes, res=0.1, noise=0.003):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + 0.5 * x1 + rng.normal(0, noise, n)
    return off, x1, x2, y


def test_oracle_continue_matches_reference_source(oracle_mod):
    from oracle import ref_source as R
    if not R.available():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt library)")
    rng = np.random.default_rng(5)
    for cap, n, n1 in ((25, 400, 150), (12, 300, 5), (40, 500, 300)):
        x1 = rng.uniform(-0.05, 0.05, n); x2 = rng.uniform(-0.05, 0.05, n)
        y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + rng.normal(0, 0.003, n)
        hy = dict(capacity=cap, **bind())
        ref = R.fit_twice(x1, x2, y, n1, rand_offset=3, **hy)
        o = oracle_mod.Oracle(rgb_rand=0, **hy)
        o.set_rand_offset(3)
        o.fit_patches([0, n1], x1[:n1], x2[:n1], y[:n1], dump=True)
        r = o.add_measurements([0, n - n1], x1[n1:], x2[n1:], y[n1:])
        assert int(r["nbv"][0]) == ref["N"]
        assert np.array_equal(r["bv1"], ref["bv1"]) and np.array_equal(r["bv2"], ref["bv2"])      # same BVs in the same slots
        assert np.abs(r["alpha"] - ref["alpha"]).max() <= 5e-6 * np.abs(ref["alpha"]).max()
        assert o.rand_offset() == 3 + (n1 - 1) + (n - n1 - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("cap,hyper", [(8, "bind"), (15, "bind"), (30, "bind"), (60, "bind"), (100, "bind"), (117, "bind"), (30, "ref")])
def test_add_measurements_matches_oracle(oracle_mod, cap, hyper):
    import gp_compressor_b200 as G
    hy = bind() if hyper == "bind" else REF
    cfg = dict(capacity=cap, **hy)
    # first fit: states of every size (empty, tiny, mid, at capacity) -> every slot format is resumed
    sizes1 = [0, 1, 3, 12, 20, 40, 70, 130, 260, 0, 33]
    sizes2 = [5, 0, 40, 9, 0, 100, 1, 50, 200, 0, 17]
    sizes3 = [0, 7, 0, 30, 2, 0, 60, 0, 10, 4, 0]
    h = G.Handle(keep_state=1, **cfg)
    o = oracle_mod.Oracle(**cfg)
    fed = np.zeros(len(sizes1), dtype=np.int64)
    for k, sizes in enumerate((sizes1, sizes2, sizes3)):
        off, x1, x2, y = make(10 * cap + k, sizes)
        if k == 0:
            h.fit_patches(off, x1, x2, y)
            want = o.fit_patches(off, x1, x2, y, dump=True)
        else:
            h.add_measurements(off, x1, x2, y)
            want = o.add_measurements(off, x1, x2, y)
        got = h.params()
        assert np.array_equal(got["nbv"], want["nbv"]), k
        assert np.array_equal(got["bv_idx"], want["bv_idx"]), k
        assert np.array_equal(got["bv1"], want["bv1"]) and np.array_equal(got["bv2"], want["bv2"]), k
        np.testing.assert_allclose(got["alpha"], want["alpha"], rtol=1e-9, atol=0)
        assert np.array_equal(got["alpha"], want["alpha"]), k
        assert h.sizes().rand_offset == o.rand_offset()
        for p in (3, 7, 8):
            N = int(want["nbv"][p])
            Cg, Qg = h.state(p, N)
            lo, hi = want["dump_off"][p], want["dump_off"][p + 1]
            assert np.array_equal(Cg.reshape(-1), want["C"][lo:hi]) and np.array_equal(Qg.reshape(-1), want["Q"][lo:hi])
        fed += np.asarray(sizes)
    # BV indices of later calls count on from the points fed before: below the total fed, some beyond the first call's
    bo = got["bv_off"]
    later = False
    for p in range(len(sizes1)):
        idx = got["bv_idx"][bo[p]:bo[p + 1]]
        assert (idx < fed[p]).all() and (idx >= 0).all() and len(set(idx.tolist())) == idx.size
        later |= bool((idx >= sizes1[p]).any())
    assert later


@pytest.mark.gpu
def test_add_measurements_equals_one_fit_without_shuffle(oracle_mod):
    """With shuffle = 0 two successive calls feed the points in the same order as one call: identical state."""
    import gp_compressor_b200 as G
    cfg = dict(capacity=20, shuffle=0, **bind())
    off, x1, x2, y = make(77, [300])
    a = G.Handle(keep_state=1, **cfg)
    a.fit_patches(off, x1, x2, y)
    b = G.Handle(keep_state=1, **cfg)
    b.fit_patches([0, 120], x1[:120], x2[:120], y[:120])
    b.add_measurements([0, 180], x1[120:], x2[120:], y[120:])
    pa, pb = a.params(), b.params()
    for k in ("nbv", "bv_idx", "bv1", "bv2", "alpha"):
        assert np.array_equal(pa[k], pb[k]), k
    X = np.stack([x1[:50], x2[:50]], 1)
    fa, sa = a.predict(0, X, sigma=True)
    fb, sb = b.predict(0, X, sigma=True)
    assert np.array_equal(fa, fb) and np.array_equal(sa, sb)


@pytest.mark.gpu
def test_add_measurements_errors():
    import gp_compressor_b200 as G
    off, x1, x2, y = make(1, [50, 60])
    h = G.Handle(capacity=10, **bind())                       # no keep_state
    h.fit_patches(off, x1, x2, y)
    with pytest.raises(RuntimeError):
        h.add_measurements(off, x1, x2, y)
    h = G.Handle(capacity=10, keep_state=1, **bind())
    with pytest.raises(RuntimeError):                          # nothing fitted yet
        h.add_measurements(off, x1, x2, y)
    h.fit_patches(off, x1, x2, y)
    with pytest.raises(RuntimeError):                          # different number of patches
        h.add_measurements(off[:2], x1[:50], x2[:50], y[:50])


@pytest.mark.gpu
def test_add_measurements_with_large_patches(oracle_mod):
    """More than 1024 new points in a patch: the shuffle takes the global-memory path and the BV indices are offset there."""
    import gp_compressor_b200 as G
    cfg = dict(capacity=20, **bind())
    h = G.Handle(keep_state=1, **cfg)
    o = oracle_mod.Oracle(**cfg)
    off, x1, x2, y = make(5, [1300, 40])
    h.fit_patches(off, x1, x2, y)
    o.fit_patches(off, x1, x2, y, dump=True)
    off2, a1, a2, ay = make(6, [1500, 0])
    h.add_measurements(off2, a1, a2, ay)
    want = o.add_measurements(off2, a1, a2, ay)
    got = h.params()
    for k in ("nbv", "bv_idx", "bv1", "bv2", "alpha"):
        assert np.array_equal(got[k], want[k]), k
    assert got["bv_idx"].max() >= 1300
