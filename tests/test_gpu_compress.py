"""GPU parity tests of the whole path through the C ABI (gpc_compress / gpc_decompress)
against the CPU oracle: identical lattice, leaves, visiting order, candidate counts,
rotations, patch assignment and per-patch point order (bit-exact, integer / index work),
identical BV index sets, alpha and reconstructed cloud (tolerance 1e-9 relative required by
the north star; bit-equality asserted because both sides use the same canonical arithmetic)."""
import numpy as np
import pytest

from gp_compressor_b200 import synth

pytestmark = pytest.mark.gpu

F32 = lambda v: float(np.float32(v))


@pytest.fixture(scope="module")
def G():
    import gp_compressor_b200 as G
    G.load()
    return G


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


def compare_all(G, O, cloud, decode=True, **cfg):
    h = G.Handle(**cfg)
    h.compress(cloud)
    o = O.Oracle(**cfg)
    want = o.compress(cloud)
    bo = o.binning()
    bg = h.patches()
    ag = h.assignment()
    assert bg["depth"] == bo["depth"]
    assert eq(bg["lattice_min"], bo["lattice_min"])
    assert bg["n_leaves"] == bo["n_leaves"] and bg["n_claimed"] == bo["n_claimed"]
    for k in ("leaf_code", "leaf_center", "leaf_ncand", "patch_off"):
        assert eq(bg[k], bo[k]), k
    assert eq(ag["owner"], bo["owner"]), "patch assignment differs"
    assert eq(ag["st_idx"], bo["st_idx"]), "per-patch point order differs"
    for k in ("leaf_R", "leaf_quat", "leaf_mean", "leaf_rgbmean"):
        assert eq(bg[k], bo[k]), k
    for k in ("st_x1", "st_x2", "st_y"):
        assert eq(ag[k], bo[k]), k
    assert eq(ag["perm"], want["perm"])
    got = h.params()
    assert eq(got["nbv"], want["nbv"]) and eq(got["bv_idx"], want["bv_idx"])
    np.testing.assert_allclose(got["alpha"], want["alpha"], rtol=1e-9, atol=0)
    assert eq(got["alpha"], want["alpha"])
    gs, os_ = h.stats(), o.stats()
    for k in ("n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo"):
        assert gs[k] == os_[k], k
    assert h.sizes().rand_offset == o.rand_offset()
    # "Max added" (gp_compressor.cpp:165-174) and the BV histogram of gpc_stats
    assert gs["max_bv"] == int(want["nbv"].max(initial=0))
    assert gs["bv_hist"] == np.bincount(np.minimum(want["nbv"], 32), minlength=33).tolist()
    if decode:
        cloud_o, heights_o = o.decode()
        assert h.decompress_resident() == heights_o.size
        hg = h.heights()
        np.testing.assert_allclose(hg, heights_o, rtol=1e-9, atol=0)
        cloud_g = h.decompress()
        assert eq(cloud_g, cloud_o)
    return h, o


def test_c1_planar_bumps_reference_config(G, oracle_mod):
    """BASELINE configs[0] at reduced size: gp_compressor(cloud, 0.15f, 20), capacity 100, reference hyper-parameters."""
    cloud = synth.c1_planar_bumps(30000, seed=1)
    h, o = compare_all(G, oracle_mod, cloud, res=F32(0.15), sz=20, capacity=100)
    assert h.sizes().n_patches > 1000


def test_c2_indoor_capacity30_both_hypersets(G, oracle_mod):
    cloud = synth.c2_indoor(60000, seed=2)
    compare_all(G, oracle_mod, cloud, res=F32(0.1), sz=10, capacity=30)
    compare_all(G, oracle_mod, cloud, res=F32(0.1), sz=10, capacity=30, **synth.hyper_bind(F32(0.1)))


def test_c5_outdoor_small(G, oracle_mod):
    cloud = synth.c5_outdoor(80000, seed=5)
    compare_all(G, oracle_mod, cloud, res=F32(0.2), sz=10, capacity=100)


def test_dense_patches_capacity_binds(G, oracle_mod):
    cloud = synth.c3_dense_floor(40000, seed=3, side=1.0)  # ~400 points per patch
    compare_all(G, oracle_mod, cloud, res=F32(0.1), sz=10, capacity=20, **synth.hyper_bind(F32(0.1)))


def test_morton_leaf_order(G, oracle_mod):
    cloud = synth.c1_planar_bumps(8000, seed=3)
    compare_all(G, oracle_mod, cloud, res=F32(0.15), sz=5, capacity=30, leaf_order=1)


def test_edge_clouds(G, oracle_mod):
    rng = np.random.default_rng(0)
    # empty, single point, non-finite points, all points in one voxel, exact duplicates, < 4 candidates
    empty = np.zeros((0, 32), dtype=np.uint8)
    h = G.Handle()
    h.compress(empty)
    assert h.sizes().n_patches == 0 and h.decompress_resident() == 0
    one = synth.pack_cloud(np.array([[1.0, 2.0, 3.0]]), np.array([[10, 20, 30]], dtype=np.uint8))
    compare_all(G, oracle_mod, one, capacity=5)
    xyz = rng.uniform(-1, 1, (3000, 3))
    xyz[::7, 0] = np.nan
    xyz[5::11, 2] = np.inf
    rgb = rng.integers(0, 256, (3000, 3)).astype(np.uint8)
    compare_all(G, oracle_mod, synth.pack_cloud(xyz, rgb), res=F32(0.25), capacity=10)
    allnan = synth.pack_cloud(np.full((10, 3), np.nan))
    h = G.Handle()
    h.compress(allnan)
    assert h.sizes().n_patches == 0
    tiny = synth.pack_cloud(rng.uniform(0, 0.01, (500, 3)), rgb[:500])
    compare_all(G, oracle_mod, tiny, capacity=8, **synth.hyper_bind(F32(0.1)))
    dup = np.repeat(rng.uniform(0, 1, (300, 3)), 4, axis=0)
    compare_all(G, oracle_mod, synth.pack_cloud(dup, np.repeat(rgb[:300], 4, axis=0)), res=F32(0.2), capacity=10)
    sparse = synth.pack_cloud(rng.uniform(0, 50, (200, 3)), rgb[:200])  # isolated points: m < 4 -> identity rotation
    compare_all(G, oracle_mod, sparse, res=F32(0.1), capacity=5)


def test_lattice_growth_directions_and_exact_planes(G, oracle_mod):
    rng = np.random.default_rng(1)
    # first point in the middle, later points far away in every direction: growth on all sides
    xyz = np.concatenate([[[0.0, 0.0, 0.0]], rng.uniform(-40, 40, (5000, 3)) * [1, 1, 0.02]])
    compare_all(G, oracle_mod, synth.pack_cloud(xyz), res=F32(0.5), capacity=10)
    # noise-free axis-aligned planes (singular Gram matrix, exact ties in the dominant-axis test)
    u = rng.uniform(0, 2, (4000, 2))
    planes = np.concatenate([np.c_[u[:1500], np.zeros(1500)], np.c_[np.zeros(1500), u[1500:3000]],
                             np.c_[u[3000:, 0], np.full(1000, 1.0), u[3000:, 1]]])
    compare_all(G, oracle_mod, synth.pack_cloud(planes), res=F32(0.1), capacity=10)


def test_second_compress_continues_rand_stream(G, oracle_mod):
    cloud = synth.c1_planar_bumps(5000, seed=9)
    cfg = dict(res=F32(0.15), sz=4, capacity=20)
    h = G.Handle(**cfg)
    o = oracle_mod.Oracle(**cfg)
    for _ in range(2):
        h.compress(cloud)
        want = o.compress(cloud)
        assert eq(h.assignment()["perm"], want["perm"])
        assert eq(h.params()["alpha"], want["alpha"])
    assert h.sizes().rand_offset == o.rand_offset() > 0


def test_upload_then_resident_equals_host_call(G):
    cloud = synth.c2_indoor(20000, seed=12)
    a = G.Handle(capacity=30)
    a.compress(cloud)
    b = G.Handle(capacity=30)
    b.upload_cloud(cloud)
    b.compress_resident()
    assert eq(a.params()["alpha"], b.params()["alpha"])
    assert eq(a.decompress(), b.decompress())


def test_sharded_compress_is_world_size_invariant(G, oracle_mod):
    cloud = synth.c2_indoor(30000, seed=7)
    cfg = dict(res=F32(0.1), sz=6, capacity=30)
    o = oracle_mod.Oracle(**cfg)
    want = o.compress(cloud)
    cloud_o, _ = o.decode()
    for world in (2, 4, 8):
        nbv, alpha, clouds = [], [], []
        for r in range(world):
            h = G.Handle(shard_rank=r, shard_count=world, **cfg)
            h.compress(cloud)
            p = h.params()
            nbv.append(p["nbv"]); alpha.append(p["alpha"])
            clouds.append(h.decompress())
        assert eq(np.concatenate(nbv), want["nbv"])
        assert eq(np.concatenate(alpha), want["alpha"])
        assert eq(np.concatenate(clouds), cloud_o)


def test_reconstruction_rmse_is_small(G):
    """Round-trip property at a larger size: the decoded surface stays close to the data."""
    cloud = synth.c1_planar_bumps(60000, seed=4)
    h = G.Handle(res=F32(0.15), sz=20, capacity=100)
    h.compress(cloud)
    a = h.assignment()
    off = h.patches(binning=False, frames=False)["patch_off"]
    se, cnt = 0.0, 0
    for p in range(0, len(off) - 1, 37):
        lo, hi = off[p], off[p + 1]
        if hi == lo:
            continue
        X = np.stack([a["st_x1"][lo:hi], a["st_x2"][lo:hi]], axis=1)
        f = h.predict(p, X)
        se += float(((f - a["st_y"][lo:hi]) ** 2).sum())
        cnt += hi - lo
    rmse = np.sqrt(se / cnt)
    assert rmse < 0.01


def test_wire_format_round_trip(G, oracle_mod, tmp_path):
    """Next-row N3: save the fitted patches, load them into a fresh handle (another process would do the same),
    decode: the cloud must be identical, and the file must be much smaller than the input cloud."""
    cloud = synth.c3_dense_floor(40000, seed=8, side=1.0)  # ~400 points per patch: the regime the codec is meant for
    cfg = dict(res=F32(0.1), sz=12, capacity=40)
    a = G.Handle(**cfg)
    a.compress(cloud)
    want = a.decompress()
    path = tmp_path / "patches.gpc"
    nbytes = a.save(path)
    assert nbytes == path.stat().st_size and 0 < nbytes < cloud.nbytes
    b = G.Handle()            # default config: everything the decoder needs comes from the file
    b.load_file(path)
    assert b.cfg.sz == 12 and b.cfg.capacity == 40
    got = b.decompress()
    assert eq(got, want)
    with pytest.raises(G.GpcError):
        bad = tmp_path / "bad.gpc"
        bad.write_bytes(b"not a parameter file")
        G.Handle().load_file(bad)
    # hostile / damaged files: every one is refused with an error (nothing thrown across the C ABI, nothing half-installed)
    raw = bytearray(path.read_bytes())
    hdr = 8 + 4 + 5 * 8          # magic, version, five doubles; then sz, capacity (int32), PL, T (int64)
    def variant(name, edit):
        b2 = bytearray(raw)
        edit(b2)
        q = tmp_path / name
        q.write_bytes(bytes(b2))
        return q
    import struct
    cases = {
        "sz0.gpc": lambda d: d.__setitem__(slice(hdr, hdr + 4), struct.pack("<i", 0)),
        "cap0.gpc": lambda d: d.__setitem__(slice(hdr + 4, hdr + 8), struct.pack("<i", 0)),
        "cap_huge.gpc": lambda d: d.__setitem__(slice(hdr + 4, hdr + 8), struct.pack("<i", 1 << 30)),
        "lsq_neg.gpc": lambda d: d.__setitem__(slice(8 + 4 + 32, 8 + 4 + 40), struct.pack("<d", -1.0)),
        "pl_huge.gpc": lambda d: d.__setitem__(slice(hdr + 8, hdr + 16), struct.pack("<q", 1 << 60)),
        "t_huge.gpc": lambda d: d.__setitem__(slice(hdr + 16, hdr + 24), struct.pack("<q", 1 << 59)),
        "truncated.gpc": lambda d: d.__delitem__(slice(len(d) // 2, len(d))),
    }
    for name, edit in cases.items():
        c = G.Handle(**cfg)
        c.compress(cloud)
        with pytest.raises(G.GpcError):
            c.load_file(variant(name, edit))
        assert c.cfg.sz == 12 and c.cfg.capacity == 40
        assert eq(c.decompress(), want)   # a refused file leaves the handle's own fit and configuration untouched
        c.load_file(path)
        assert eq(c.decompress(), want)


def test_heights_need_a_resident_decompress(G):
    """gpc_get_heights returns the grid of the last gpc_decompress_resident; after gpc_decompress (which produces no
    heights) it is a state error, not stale memory."""
    cloud = synth.c2_indoor(20000, seed=4)
    h = G.Handle(res=F32(0.1), sz=5, capacity=20)
    h.compress(cloud)
    n = h.decompress_resident()
    hg = h.heights()
    assert hg.size == n and np.isfinite(hg).all()
    h.decompress()
    with pytest.raises(G.GpcError):
        h.heights()
    h.decompress_resident()
    assert eq(h.heights(), hg)


def test_handle_reuse_across_different_clouds(G, oracle_mod):
    """One handle, clouds of different sizes and extents in sequence (buffers are grow-only and reused)."""
    cfg = dict(res=F32(0.1), sz=4, capacity=30)
    h = G.Handle(**cfg)
    o = oracle_mod.Oracle(**cfg)
    for cloud in (synth.c2_indoor(30000, seed=1), synth.c1_planar_bumps(4000, seed=2), synth.c2_indoor(60000, seed=3),
                  synth.pack_cloud(np.array([[5.0, 5.0, 5.0]])), synth.c3_dense_floor(20000, seed=4, side=1.0)):
        h.compress(cloud)
        want = o.compress(cloud)
        assert eq(h.assignment()["owner"], o.binning()["owner"])
        assert eq(h.params()["alpha"], want["alpha"])
        co, _ = o.decode(want_heights=False)
        assert eq(h.decompress(), co)


def test_extreme_coordinates_and_depth_overflow(G, oracle_mod):
    rng = np.random.default_rng(2)
    # far from the origin and negative: float spacing ~ 0.06 at 1e6, still a valid lattice
    xyz = rng.uniform(-1, 1, (4000, 3)) * [3, 3, 0.05] + [-1.0e5, 2.0e5, 50.0]
    compare_all(G, oracle_mod, synth.pack_cloud(xyz), res=F32(0.25), capacity=10)
    # an extent / resolution ratio beyond 2^21 voxels per axis cannot be keyed in 63 bits: clean error, no crash
    far = synth.pack_cloud(np.array([[0.0, 0.0, 0.0], [1.0e6, 0.0, 0.0]]))
    with pytest.raises(G.GpcError):
        G.Handle(res=F32(0.01)).compress(far)


@pytest.mark.parametrize("seed", range(6))
def test_random_clouds_whole_path(G, oracle_mod, seed):
    """Random slabs with NaNs, random resolution / grid / capacity / hyper-set / leaf order: every intermediate and the
    decoded cloud bit-equal to the oracle (compare_all)."""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(500, 40000))
    ext = rng.uniform(0.3, 4.0, 3) * [1, 1, 0.15]
    xyz = (rng.uniform(-1, 1, (n, 3)) * ext + rng.uniform(-20, 20, 3)).astype(np.float32)
    xyz[:, 2] += (0.04 * np.sin(4 * xyz[:, 0]) * np.cos(3 * xyz[:, 1])).astype(np.float32)
    xyz[rng.random(n) < 0.01] = np.nan
    cloud = synth.pack_cloud(xyz, rng.integers(0, 256, (n, 3)).astype(np.uint8))
    res = F32(rng.choice([0.05, 0.1, 0.2, 0.4]))
    cfg = dict(res=res, sz=int(rng.integers(1, 9)), capacity=int(rng.integers(2, 70)), leaf_order=int(rng.integers(0, 2)))
    if rng.random() < 0.5:
        cfg.update(synth.hyper_bind(res))
    compare_all(G, oracle_mod, cloud, **cfg)
