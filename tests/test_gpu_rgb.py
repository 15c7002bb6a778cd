"""Next-row N1 on the GPU: the RGB field GP (sparse_gp_field<rbf_kernel, gaussian_noise_3d>) fitted next to the height
GP (gp_compressor.cpp:163) and evaluated at decode (gp_compressor.cpp:334,367-371), against the CPU oracle (which is itself
pinned against the reference's own sparse_gp_field.hpp, tests/test_oracle_vs_reference_source.py).  Bit-exact bar."""
import numpy as np
import pytest

from gp_compressor_b200 import synth

pytestmark = pytest.mark.gpu
F32 = lambda v: float(np.float32(v))


def run(cloud, **cfg):
    import gp_compressor_b200 as G
    from oracle import oracle as O
    h = G.Handle(rgb=1, **cfg)
    h.compress(cloud)
    o = O.Oracle(rgb=1, **cfg)
    want_h = o.compress(cloud)
    want = o.rgb_result()
    got = h.params_rgb()
    assert np.array_equal(got["perm"], want["perm"])
    assert np.array_equal(got["nbv"], want["nbv"]) and np.array_equal(got["bv_idx"], want["bv_idx"])
    assert np.array_equal(got["bv1"], want["bv1"]) and np.array_equal(got["bv2"], want["bv2"])
    assert np.array_equal(got["alpha"], want["alpha"], equal_nan=True)
    # the height GP is untouched by the extra work
    assert np.array_equal(h.params()["alpha"], want_h["alpha"])
    gs, os_ = h.stats(), o.stats_rgb()
    assert (gs["rgb_n_sparse"], gs["rgb_n_full"], gs["rgb_n_del_cap"], gs["rgb_n_del_geo"]) == \
           (os_["n_sparse"], os_["n_full"], os_["n_del_cap"], os_["n_del_geo"])
    assert gs["rgb_sum_n2_common"] == os_["sumN2_common"]
    assert h.sizes().rand_offset == o.rand_offset()
    cloud_o, _ = o.decode(want_heights=False)
    cloud_g = h.decompress()
    assert np.array_equal(cloud_g, cloud_o)
    return h, o, cloud_g


def test_rgb_reference_defaults():
    """Reference configuration: capacity 100, s0 = 1e2f, eps_tol = 1e-4f (the field GP keeps ~4-6 BVs)."""
    cloud = synth.c1_planar_bumps(30000, seed=1)
    h, o, dec = run(cloud, res=F32(0.15), sz=8, capacity=100)
    # colours are no longer constant inside a patch
    rgb = dec[:, 16:19].reshape(-1, 64, 3).astype(int)
    assert (rgb.max(axis=1) - rgb.min(axis=1)).max() > 0


def test_rgb_capacity_bound_and_bucket_chain():
    """A hyper-set under which the field GP grows: exercises its delete_bv variant and the 0 -> 2 bucket hand-off."""
    cloud = synth.c3_dense_floor(30000, seed=5, side=1.0)
    hyp = synth.hyper_bind(F32(0.1))
    run(cloud, res=F32(0.1), sz=4, capacity=12, rgb_s0=1e-2, sigmaf_sq=hyp["sigmaf_sq"], l_sq=hyp["l_sq"], s0=hyp["s0"])
    run(cloud, res=F32(0.1), sz=4, capacity=40, rgb_s0=1e-2, sigmaf_sq=hyp["sigmaf_sq"], l_sq=hyp["l_sq"], s0=hyp["s0"])


def test_rgb_sharded_matches_single():
    import gp_compressor_b200 as G
    cloud = synth.c2_indoor(20000, seed=3)
    cfg = dict(res=F32(0.1), sz=5, capacity=30, rgb=1)
    whole = G.Handle(**cfg)
    whole.compress(cloud)
    want = whole.decompress()
    parts = []
    for r in range(3):
        h = G.Handle(shard_rank=r, shard_count=3, **cfg)
        h.compress(cloud)
        parts.append(h.decompress())
    assert np.array_equal(np.concatenate(parts), want)


def test_rgb_survives_the_wire_format(tmp_path):
    import gp_compressor_b200 as G
    cloud = synth.c1_planar_bumps(15000, seed=2)
    a = G.Handle(res=F32(0.15), sz=6, capacity=50, rgb=1)
    a.compress(cloud)
    want = a.decompress()
    a.save(tmp_path / "p.gpc")
    b = G.Handle()
    b.load_file(tmp_path / "p.gpc")
    assert b.cfg.rgb == 1
    assert np.array_equal(b.decompress(), want)


def test_large_patches_use_the_global_shuffle_path():
    """Patches with more than 1024 points take the thread-per-patch global-memory shuffle (both the height GP's and the
    RGB field GP's), smaller ones the shared-memory warp shuffle: mix both in one cloud."""
    rng = np.random.default_rng(3)
    dense = rng.uniform(0.0, 0.19, (9000, 3)) * [1, 1, 0.05]          # ~4 voxels with > 2000 points each
    sparse = rng.uniform(1.0, 3.0, (3000, 3)) * [1, 1, 0.02]
    xyz = np.concatenate([dense, sparse])[rng.permutation(12000)]
    rgb = rng.integers(0, 256, (12000, 3)).astype(np.uint8)
    h, o, _ = run(synth.pack_cloud(xyz, rgb), res=F32(0.1), sz=3, capacity=20)
    assert np.diff(h.patches(frames=False, binning=False)["patch_off"]).max() > 1024


@pytest.mark.parametrize("regime", ["reference", "grown"])
def test_rgb_field_gp_evaluation_matches_oracle(regime):
    """sparse_gp_field::predict_measurements (sigma / conf), compute_likelihoods, compute_derivatives
    (sparse_gp_field.hpp:267-393; gp_registration.cpp:177,194) through gpc_evaluate_patches_rgb: bit-equal to the oracle,
    whose restatement is pinned against the reference's own source in test_oracle_field_evaluate_matches_reference_source."""
    import gp_compressor_b200 as G
    from oracle import oracle as O
    if regime == "reference":
        cloud = synth.c1_planar_bumps(20000, seed=3)
        cfg = dict(res=F32(0.15), sz=4, capacity=100)
    else:
        cloud = synth.c3_dense_floor(30000, seed=5, side=1.0)
        hyp = synth.hyper_bind(F32(0.1))
        cfg = dict(res=F32(0.1), sz=4, capacity=40, rgb_s0=1e-2, sigmaf_sq=hyp["sigmaf_sq"], l_sq=hyp["l_sq"], s0=hyp["s0"])
    h = G.Handle(rgb=1, keep_state=1, **cfg)
    h.compress(cloud)
    o = O.Oracle(rgb=1, **cfg)
    o.compress(cloud, dump=True)
    P = int(h.sizes().n_patches)
    rng = np.random.default_rng(7)
    cnt = rng.integers(0, 12, P)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    m = int(off[-1])
    half = cfg["res"] / 2
    q1, q2 = rng.uniform(-half, half, m), rng.uniform(-half, half, m)
    Y = rng.normal(0, 20, (m, 3))
    for conf in (False, True):
        g = h.evaluate_rgb(off, q1, q2, Y, conf=conf)
        w = o.evaluate_rgb(off, q1, q2, Y, conf=conf)
        for k in ("f", "sigma", "lik", "dX"):
            assert np.array_equal(g[k], w[k], equal_nan=True), (k, conf)
    assert np.all(g["dX"][:, 0] == 0)
    # the height GPs of the same handle still evaluate as before
    ge = h.evaluate(off, q1, q2, Y[:, 0].copy())
    we = o.evaluate(off, q1, q2, Y[:, 0].copy())
    assert np.array_equal(ge["dX"], we["dX"]) and np.array_equal(ge["sigma"], we["sigma"])
    no_state = G.Handle(rgb=1, **cfg)
    no_state.compress(cloud)
    with pytest.raises(RuntimeError):
        no_state.evaluate_rgb(off, q1, q2, Y)
