"""GPU parity tests (run with -m gpu on the B200 box): K6 rand/shuffle, K7 SOGP fit and K8
grid prediction through the C ABI against the CPU oracle on the same seeded inputs.
Bar: BV counts, BV index sets, permutations and event counters identical; alpha and
heights within 1e-9 relative (north star) — and, because oracle and kernels implement the
same canonical arithmetic, bit-equality is asserted as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REF = dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1)))


def bind(res=0.1):
    return dict(sigmaf_sq=1.0, l_sq=(res / 12.0) ** 2, s0=1e-4)


def make_patches(seed, sizes, res=0.1, noise=0.003):
    rng = np.random.default_rng(seed)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    x1 = rng.uniform(-res / 2, res / 2, n)
    x2 = rng.uniform(-res / 2, res / 2, n)
    y = 0.02 * np.sin(40 * x1) * np.cos(30 * x2) + 0.5 * x1 + rng.normal(0, noise, n)
    return off, x1, x2, y


@pytest.fixture(scope="module")
def G():
    import gp_compressor_b200 as G
    G.load()
    return G


def check_fit(G, O, off, x1, x2, y, exact=True, **cfg):
    h = G.Handle(keep_state=1, **cfg)
    h.fit_patches(off, x1, x2, y)
    got = h.params()
    o = O.Oracle(**cfg)
    want = o.fit_patches(off, x1, x2, y, dump=True)
    assert np.array_equal(got["nbv"], want["nbv"])
    assert np.array_equal(got["bv_off"], want["bv_off"])
    assert np.array_equal(got["bv_idx"], want["bv_idx"])
    assert np.array_equal(got["bv1"], want["bv1"]) and np.array_equal(got["bv2"], want["bv2"])
    np.testing.assert_allclose(got["alpha"], want["alpha"], rtol=1e-9, atol=0)
    if exact:
        assert np.array_equal(got["alpha"], want["alpha"])
    a = h.assignment(binning=False)
    assert np.array_equal(a["perm"], want["perm"])
    gs, os_ = h.stats(), o.stats()
    for k in ("n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo"):
        assert gs[k] == os_[k], k
    assert gs["sum_n"] == os_["sumN"] and gs["sum_n2_common"] == os_["sumN2_common"]
    assert gs["sum_n2_sparse"] == os_["sumN2_sparse"] and gs["sum_n2_full"] == os_["sumN2_full"]
    assert gs["sum_n2_del"] == os_["sumN2_del"]
    assert h.sizes().rand_offset == o.rand_offset()
    return h, o, got, want


def test_device_exp_bit_equal_to_oracle(G, oracle_mod):
    rng = np.random.default_rng(1)
    x = np.concatenate([-rng.uniform(0, 80, 300000), -rng.uniform(0, 1e-3, 100000), rng.uniform(0, 3, 50000),
                        np.array([0.0, -0.0, -1e-300, -700.0, -730.0, -744.9, -746.0, 800.0, np.nan])])
    h = G.Handle()
    got = h.debug_exp(x)
    want = oracle_mod.exp_array(x)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_device_rand_stream(G, oracle_mod):
    h = G.Handle()
    assert np.array_equal(h.debug_rand(0, 100000), oracle_mod.rand_stream(0, 100000))
    for off in (1, 30, 31, 511, 512, 513, 12345678, 3_000_000_019):
        if off < 20_000_000:
            assert np.array_equal(h.debug_rand(off, 3000), oracle_mod.rand_stream(off, 3000)), off
    # far jump: consistency between two device jumps (the CPU oracle walks the stream)
    a = h.debug_rand(3_000_000_019, 2000)
    b = h.debug_rand(3_000_000_019 + 700, 1300)
    assert np.array_equal(a[700:], b)


@pytest.mark.parametrize("cap", [1, 2, 5, 15, 16, 20, 31, 32, 40, 63, 64, 100, 117, 118, 150, 200])
def test_fit_bind_hyperset_all_buckets(G, oracle_mod, cap):
    """Capacity binds: exercises full updates, capacity deletions and every SOGP bucket."""
    sizes = [0, 1, 2, 3, 17, 64, 150, 0, 333, 40] + ([700] if cap > 100 else [])
    off, x1, x2, y = make_patches(100 + cap, sizes)
    check_fit(G, oracle_mod, off, x1, x2, y, capacity=cap, **bind())


@pytest.mark.parametrize("cap", [30, 100])
def test_fit_reference_hyperset(G, oracle_mod, cap):
    """Reference defaults (rbf 100/1, s0 1e-1f): ill-conditioned, sparse updates and geometric deletions."""
    rng = np.random.default_rng(7)
    sizes = rng.integers(0, 400, 200)
    off, x1, x2, y = make_patches(8, sizes)
    h, o, got, want = check_fit(G, oracle_mod, off, x1, x2, y, capacity=cap, **REF)
    assert o.stats()["n_sparse"] > 0


def test_fit_no_shuffle_and_rand_offset(G, oracle_mod):
    off, x1, x2, y = make_patches(3, [50, 60, 0, 70])
    check_fit(G, oracle_mod, off, x1, x2, y, capacity=10, shuffle=0, **bind())
    # second call on the same handle continues the rand stream
    h = G.Handle(capacity=10, **bind())
    o = oracle_mod.Oracle(capacity=10, **bind())
    for _ in range(2):
        h.fit_patches(off, x1, x2, y)
        want = o.fit_patches(off, x1, x2, y)
        assert np.array_equal(h.assignment(binning=False)["perm"], want["perm"])
        assert np.array_equal(h.params()["bv_idx"], want["bv_idx"])
    assert h.sizes().rand_offset == o.rand_offset() == 2 * 2 * (49 + 59 + 69)


def test_state_matches_oracle(G, oracle_mod):
    off, x1, x2, y = make_patches(11, [300, 5])
    h, o, got, want = check_fit(G, oracle_mod, off, x1, x2, y, capacity=25, **bind())
    for p in range(2):
        N = int(want["nbv"][p])
        Cg, Qg = h.state(p, N)
        lo, hi = want["dump_off"][p], want["dump_off"][p + 1]
        assert np.array_equal(Cg.reshape(-1), want["C"][lo:hi])
        assert np.array_equal(Qg.reshape(-1), want["Q"][lo:hi])
        assert np.array_equal(Cg, Cg.T) and np.array_equal(Qg, Qg.T)


def test_duplicate_points_geometric_deletion(G, oracle_mod):
    """Exact duplicates make gamma ~ 0 and Q ill-conditioned: sparse path + geometric deletes."""
    off, x1, x2, y = make_patches(5, [120])
    x1[40:80] = x1[0:40]
    x2[40:80] = x2[0:40]
    check_fit(G, oracle_mod, off, x1, x2, y, capacity=12, sigmaf_sq=1.0, l_sq=1e-4, s0=1e-6, exact=False)


@pytest.mark.parametrize("separable", [0, 1])
@pytest.mark.parametrize("sz", [1, 10, 20, 40])
def test_decode_heights_and_cloud(G, oracle_mod, sz, separable):
    """separable = 0: rbf_kernel::kernel_function per (grid point, BV) as the reference (the default);
    1: the flagged fast mode.  Both bit-equal to the oracle's restatement of the same mode."""
    off, x1, x2, y = make_patches(21, [0, 90, 200, 0, 31])
    cfg = dict(capacity=20, sz=sz, decode_separable=separable, **bind())
    h = G.Handle(**cfg)
    h.fit_patches(off, x1, x2, y)
    o = oracle_mod.Oracle(**cfg)
    o.fit_patches(off, x1, x2, y)
    cloud_o, heights_o = o.decode()
    n = h.decompress_resident()
    assert n == heights_o.size == 3 * sz * sz
    hg = h.heights()
    np.testing.assert_allclose(hg, heights_o, rtol=1e-9, atol=0)
    assert np.array_equal(hg, heights_o)
    cloud_g = h.decompress()
    assert np.array_equal(cloud_g, cloud_o)  # identity frames: xyz = (f, X0, X1) as float
    X = np.stack([x1[:64], x2[:64]], axis=1)
    assert np.array_equal(h.predict(1, X), o.predict(1, X))


@pytest.mark.parametrize("separable,sz", [(0, 16), (1, 16), (0, 64), (1, 64)])
def test_decode_only_with_frames(G, oracle_mod, separable, sz):
    from gp_compressor_b200 import synth
    prm = synth.c4_patch_params(n_patches=500 if sz == 16 else 40, nbv=30, seed=4)
    cfg = dict(capacity=30, sz=sz, decode_separable=separable, **REF)
    h = G.Handle(**cfg)
    h.set_params(**prm)
    o = oracle_mod.Oracle(**cfg)
    o.set_params(**prm)
    cloud_o, heights_o = o.decode()
    assert h.decompress_resident() == heights_o.size
    assert np.array_equal(h.heights(), heights_o)
    cloud_g = h.decompress()
    assert np.array_equal(cloud_g, cloud_o)


def test_predict_sigma(G, oracle_mod):
    off, x1, x2, y = make_patches(31, [150])
    cfg = dict(capacity=15, **bind())
    h = G.Handle(keep_state=1, **cfg)
    h.fit_patches(off, x1, x2, y)
    o = oracle_mod.Oracle(**cfg)
    o.fit_patches(off, x1, x2, y, dump=True)
    X = np.stack([x1[:40], x2[:40]], axis=1)
    fg, sg = h.predict(0, X, sigma=True)
    fo, so = o.predict(0, X, sigma=True)
    assert np.array_equal(fg, fo) and np.array_equal(sg, so)


def test_sharded_fit_is_world_size_invariant(G, oracle_mod):
    rng = np.random.default_rng(13)
    sizes = rng.integers(0, 120, 97)
    off, x1, x2, y = make_patches(14, sizes)
    cfg = dict(capacity=12, **bind())
    o = oracle_mod.Oracle(**cfg)
    want = o.fit_patches(off, x1, x2, y)
    for world in (1, 2, 4, 8):
        nbv, idx, alpha = [], [], []
        covered = 0
        for r in range(world):
            h = G.Handle(shard_rank=r, shard_count=world, **cfg)
            h.fit_patches(off, x1, x2, y)
            p = h.params()
            assert p["patch_lo"] == covered
            covered = p["patch_hi"]
            nbv.append(p["nbv"]); idx.append(p["bv_idx"]); alpha.append(p["alpha"])
            assert h.sizes().rand_offset == o.rand_offset()
        assert covered == len(sizes)
        assert np.array_equal(np.concatenate(nbv), want["nbv"])
        assert np.array_equal(np.concatenate(idx), want["bv_idx"])
        assert np.array_equal(np.concatenate(alpha), want["alpha"])


def test_errors(G):
    with pytest.raises(G.GpcError):
        G.Handle(capacity=0).fit_patches([0, 1], [0.0], [0.0], [0.0])
    with pytest.raises(G.GpcError):
        G.Handle(capacity=202).fit_patches([0, 1], [0.0], [0.0], [0.0])
    with pytest.raises(G.GpcError):
        G.Handle().decompress_resident()


@pytest.mark.parametrize("cap", [50, 90])
def test_geometric_deletions_in_the_fused_buckets(G, oracle_mod, cap):
    """Near-duplicate points with a tiny novelty threshold: full updates with huge 1/gamma, so 1/Q_ii drops below 1e-9f
    and the geometric deletion loop (sparse_gp.hpp:226-242) fires while N is in the fused buckets (N > 32) -- the rare
    path of sogp_fit_fused_kernel, which first brings C and Q up to date and then runs the step-by-step deletion."""
    off, x1, x2, y = make_patches(40 + cap, [500])
    rng = np.random.default_rng(cap)
    src = rng.integers(0, 250, 120)
    x1[250:370] = x1[src] + rng.uniform(-1, 1, 120) * 3e-7
    x2[250:370] = x2[src] + rng.uniform(-1, 1, 120) * 3e-7
    cfg = dict(capacity=cap, sigmaf_sq=1.0, l_sq=(0.1 / 14.0) ** 2, s0=1e-6, eps_tol=1e-14, shuffle=0)
    h, o, got, want = check_fit(G, oracle_mod, off, x1, x2, y, **cfg)
    st = o.stats()
    assert st["n_del_geo"] > 0 and int(want["nbv"][0]) > 32
