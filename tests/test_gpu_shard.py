"""GPU parity tests of the sharded-binning compress (gpc_compress_shard_begin / _finish, strong scaling of one cloud):
R ranks -- here R handles on one GPU, visited in rank order with the all-gather done in Python -- must reproduce the
single-handle result bit for bit: patch counts, BV counts, BV index sets, alpha, rand() accounting and the decoded cloud."""
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


def sharded(G, cloud, world, resident=False, **cfg):
    hs = [G.Handle(shard_rank=r, shard_count=world, **cfg) for r in range(world)]
    counts = []
    for h in hs:
        if resident:
            h.upload_cloud(cloud)
            counts.append(h.compress_shard_begin())
        else:
            counts.append(h.compress_shard_begin(cloud))
    pt, dt = sum(c[0] for c in counts), sum(c[1] for c in counts)
    pb = db = 0
    for h, (p, d) in zip(hs, counts):          # what the ranks learn from one all-gather of two integers
        h.compress_shard_finish(pb, db, pt, dt)
        pb += p; db += d
    return hs, counts


def check(G, cloud, worlds, **cfg):
    ref = G.Handle(**cfg)
    ref.compress(cloud)
    want = ref.params()
    want_cloud = ref.decompress()
    want_sizes = ref.sizes()
    for world in worlds:
        hs, counts = sharded(G, cloud, world, **cfg)
        assert sum(c[0] for c in counts) == want_sizes.n_patches
        nbv, idx, alpha, b1, clouds = [], [], [], [], []
        nxt = 0
        sel = 0
        for h in hs:
            s = h.sizes()
            assert s.patch_lo == nxt and s.n_patches == want_sizes.n_patches
            nxt = s.patch_hi
            assert s.rand_offset == want_sizes.rand_offset
            p = h.params()
            nbv.append(p["nbv"]); idx.append(p["bv_idx"]); alpha.append(p["alpha"]); b1.append(p["bv1"])
            clouds.append(h.decompress())
        assert nxt == want_sizes.n_patches
        assert np.array_equal(np.concatenate(nbv), want["nbv"])
        assert np.array_equal(np.concatenate(idx), want["bv_idx"])
        assert np.array_equal(np.concatenate(b1), want["bv1"])
        assert np.array_equal(np.concatenate(alpha), want["alpha"])
        assert np.array_equal(np.concatenate(clouds), want_cloud)
    return ref


@pytest.mark.parametrize("leaf_order", [0, 1])
def test_sharded_binning_matches_single_gpu_indoor(G, leaf_order):
    cloud = synth.c2_indoor(120000, seed=9)
    check(G, cloud, (1, 2, 3, 8), res=F32(0.1), sz=6, capacity=30, leaf_order=leaf_order)


def test_sharded_binning_matches_single_gpu_outdoor_with_rgb(G):
    cloud = synth.c5_outdoor(150000, seed=3)
    check(G, cloud, (2, 5), res=F32(0.2), sz=5, capacity=20, rgb=1)


def test_sharded_binning_small_and_degenerate_clouds(G):
    # a lattice too small to cut (everything goes to rank 0), NaN points, a single point, an empty cloud
    cloud = synth.c1_planar_bumps(3000, seed=2)
    cloud[::97, :4] = np.frombuffer(np.float32(np.nan).tobytes(), dtype=np.uint8)
    check(G, cloud, (2, 4), res=F32(0.15), sz=4, capacity=12)
    check(G, cloud[:1].copy(), (2,), res=F32(0.15), sz=4, capacity=12)
    hs, counts = sharded(G, np.zeros((0, 32), dtype=np.uint8), 2, res=F32(0.15), sz=4, capacity=12)
    assert counts == [(0, 0), (0, 0)] and all(h.sizes().n_patches == 0 for h in hs)


def test_sharded_binning_bins_fewer_points_per_rank(G):
    """The point of the exercise: with 8 ranks a rank sorts / rotates / claims a fraction of the cloud."""
    cloud = synth.c5_outdoor(400000, seed=5)
    cfg = dict(res=F32(0.2), sz=4, capacity=20)
    hs, counts = sharded(G, cloud, 8, resident=True, **cfg)
    ref = G.Handle(**cfg)
    ref.compress(cloud)
    claimed = [h.sizes().n_claimed for h in hs]
    assert max(claimed) < 0.45 * ref.sizes().n_claimed          # own range + halo, not the whole cloud
    with pytest.raises(RuntimeError):
        hs[0].patches()                                          # patch-level arrays are shard-local in this mode


@pytest.mark.parametrize("seed", range(8))
def test_sharded_binning_random_clouds(G, seed):
    """Random slabs with NaNs, random resolution, shard count and leaf order: the halo / ownership logic at lattice edges."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2000, 60000))
    ext = rng.uniform(0.5, 6.0, 3) * [1, 1, 0.2]
    xyz = (rng.uniform(-1, 1, (n, 3)) * ext + rng.uniform(-50, 50, 3)).astype(np.float32)
    xyz[:, 2] += (0.05 * np.sin(3 * xyz[:, 0]) * np.cos(2 * xyz[:, 1])).astype(np.float32)
    xyz[rng.random(n) < 0.01] = np.nan
    cloud = synth.pack_cloud(xyz, rng.integers(0, 256, (n, 3)).astype(np.uint8))
    res = F32(rng.choice([0.05, 0.1, 0.2, 0.4]))
    world = int(rng.integers(2, 7))
    check(G, cloud, (world,), res=res, sz=3, capacity=int(rng.integers(4, 30)), leaf_order=int(rng.integers(0, 2)))
