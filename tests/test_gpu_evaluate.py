"""GPU parity tests of the next rows N2 / N4 (run with -m gpu): batched predict with sigma / conf, likelihood and
likelihood gradient (gpc_evaluate_patches, kernel K9) against the CPU oracle -- bit-equal, because both implement the
same canonical arithmetic -- and against the reference's own predict_measurements / compute_likelihoods /
compute_derivatives (oracle/_ref, sparse_gp.hpp:299-351, 407-502) within 1e-9 relative on well-conditioned fits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bind(res=0.1):
    return dict(sigmaf_sq=1.0, l_sq=(res / 12.0) ** 2, s0=1e-4)


def make(seed, sizes, qsizes, res=0.1, noise=0.003):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    fn = lambda a, b: 0.02 * np.sin(40 * a) * np.cos(30 * b) + 0.5 * a
    y = fn(x1, x2) + rng.normal(0, noise, n)
    qoff = np.concatenate([[0], np.cumsum(qsizes)]).astype(np.int64)
    m = int(qoff[-1])
    q1 = rng.uniform(-res / 2, res / 2, m)
    q2 = rng.uniform(-res / 2, res / 2, m)
    qy = fn(q1, q2) + rng.normal(0, 3 * noise, m)
    return (off, x1, x2, y), (qoff, q1, q2, qy)


@pytest.fixture(scope="module")
def G():
    import gp_compressor_b200 as G
    G.load()
    return G


def run_both(G, O, fit, query, **cfg):
    h = G.Handle(keep_state=1, **cfg)
    h.fit_patches(*fit)
    o = O.Oracle(**cfg)
    o.fit_patches(*fit, dump=True)
    out = {}
    for conf in (False, True):
        g = h.evaluate(*query, conf=conf)
        w = o.evaluate(*query, conf=conf)
        for k in ("f", "sigma", "lik", "dX"):
            assert np.array_equal(g[k], w[k]), (k, conf, np.abs(g[k] - w[k]).max())
        out[conf] = g
    assert h.stats()["ms_evaluate"] > 0
    return h, o, out


@pytest.mark.parametrize("cap,sizes", [(12, [0, 1, 5, 40, 200, 33, 0, 150]),      # N <= 32: one row per thread
                                       (60, [300, 17, 0, 450, 260]),              # 4 x 4 register tiles, C in shared memory
                                       (140, [900, 1200])])                        # C read from global memory
def test_evaluate_matches_oracle(G, oracle_mod, cap, sizes):
    rng = np.random.default_rng(cap)
    qsizes = [int(v) for v in rng.integers(0, 90, len(sizes))]
    qsizes[0] = 37                                                    # an empty GP is still evaluated (prior)
    fit, query = make(100 + cap, sizes, qsizes)
    h, o, out = run_both(G, oracle_mod, fit, query, capacity=cap, **bind())
    nbv = h.params()["nbv"]
    assert nbv.max() > (32 if cap > 32 else 4)
    if cap == 140:
        assert nbv.max() > 124
    # the mean agrees with the point-wise predict entry (same row4 dot)
    qoff, q1, q2, _ = query
    p = int(np.argmax(np.diff(qoff) > 0))
    sl = slice(qoff[p], qoff[p + 1])
    assert np.array_equal(out[False]["f"][sl], h.predict(p, np.stack([q1[sl], q2[sl]], 1)))
    # empty GP: prior mean 0, sigma sqrt(k** + s20), confidence 0 (sparse_gp.hpp:322-327, 340-346)
    e = slice(qoff[0], qoff[1])
    if sizes[0] == 0:
        assert np.all(out[False]["f"][e] == 0.0)
        assert np.all(out[False]["sigma"][e] == np.sqrt(1.0 + 1e-4))
        assert np.all(out[True]["sigma"][e] == 0.0)


def test_evaluate_outputs_optional_and_errors(G, oracle_mod):
    fit, query = make(7, [80, 90], [10, 12])
    cfg = dict(capacity=10, **bind())
    h = G.Handle(keep_state=1, **cfg)
    h.fit_patches(*fit)
    full = h.evaluate(*query)
    only = h.evaluate(query[0], query[1], query[2], None, want=("f", "sigma"))
    assert set(only) == {"f", "sigma"}
    assert np.array_equal(only["f"], full["f"]) and np.array_equal(only["sigma"], full["sigma"])
    h2 = G.Handle(**cfg)                      # no keep_state: the C matrices are not kept
    h2.fit_patches(*fit)
    with pytest.raises(RuntimeError):
        h2.evaluate(*query)
    with pytest.raises(RuntimeError):         # more patches than fitted
        h.evaluate(np.array([0, 1, 2, 3]), query[1][:3], query[2][:3], query[3][:3])


def test_evaluate_matches_reference_source(G, oracle_mod):
    """The reference's own code (oracle/_ref) on one patch: same BV set on this well-conditioned problem, so its
    predict / likelihood / likelihood_dx outputs must agree with the kernel's within the path tolerance."""
    from oracle import ref_source
    if not ref_source.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine and no prebuilt library)")
    res = 0.15
    hy = dict(capacity=20, s0=1e-3, sigmaf_sq=1.0, l_sq=(res / 6) ** 2)
    fit, query = make(3, [80], [60], res=res, noise=1e-4)
    off, x1, x2, y = fit
    qoff, q1, q2, qy = query
    ref = ref_source.evaluate(x1, x2, y, np.stack([q1, q2], 1), qy, **hy)
    h = G.Handle(keep_state=1, res=res, **hy)
    h.fit_patches(off, x1, x2, y)
    assert h.params()["nbv"][0] == ref["N"]
    g = h.evaluate(qoff, q1, q2, qy)
    gc = h.evaluate(qoff, q1, q2, qy, conf=True)
    np.testing.assert_allclose(g["f"], ref["f"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["sigma"], ref["sigma"], rtol=1e-9)
    np.testing.assert_allclose(gc["sigma"], ref["conf"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(g["lik"], ref["lik"], rtol=1e-9)
    np.testing.assert_allclose(g["dX"], ref["dX"], rtol=1e-8, atol=1e-9 * np.abs(ref["dX"]).max())
