"""Properties that do not need the oracle, checked at BASELINE.json's full single-GPU size (configs[1]: 5 M points,
res 0.1, capacity 30): every point is claimed at most once and exactly by the patch the stream says; per-patch
permutations are permutations; BV counts respect the capacity; BVs are points of their patch; the decoder emits
sz^2 points per non-empty patch; the fitted surfaces reproduce the data (round trip); results do not depend on how
the patches are sharded (checksum of checksums)."""
import hashlib

import numpy as np
import pytest

from gp_compressor_b200 import synth

pytestmark = pytest.mark.gpu
F32 = lambda v: float(np.float32(v))
N = 5_000_000


@pytest.fixture(scope="module")
def run():
    import gp_compressor_b200 as G
    cloud = synth.c2_indoor(N, seed=2)
    h = G.Handle(res=F32(0.1), sz=10, capacity=30)
    h.compress(cloud)
    return G, cloud, h


def test_assignment_invariants(run):
    G, cloud, h = run
    s = h.sizes()
    a = h.assignment()
    p = h.patches(frames=False, binning=False)
    off = p["patch_off"]
    assert s.n_in == N and 0 < s.n_claimed <= N and off[0] == 0 and off[-1] == s.n_claimed
    assert np.all(np.diff(off) >= 0)
    # claimed once: stream indices are distinct, and owner[] agrees with the stream
    st = a["st_idx"]
    assert np.unique(st).size == st.size
    patch_of = np.repeat(np.arange(s.n_patches, dtype=np.int32), np.diff(off))
    assert np.array_equal(a["owner"][st], patch_of)
    assert int((a["owner"] >= 0).sum()) == s.n_claimed
    # local coordinates inside the patch box (closed), gp_compressor.cpp:85
    half = F32(0.1) / 2.0
    assert np.abs(a["st_x1"]).max() <= half and np.abs(a["st_x2"]).max() <= half
    # per-patch mean of y is ~0 (heights are mean-subtracted, gp_compressor.cpp:103-106)
    sums = np.add.reduceat(a["st_y"], off[:-1][np.diff(off) > 0])
    assert np.abs(sums).max() < 1e-9
    # permutations: sorting perm inside each patch gives 0..n-1
    perm = a["perm"].astype(np.int64)
    key = patch_of.astype(np.int64) * (perm.max() + 1) + perm
    srt = np.sort(key) - patch_of.astype(np.int64) * (perm.max() + 1)
    assert np.array_equal(srt, np.arange(s.n_claimed) - np.repeat(off[:-1], np.diff(off)))


def test_fit_invariants_and_round_trip(run):
    G, cloud, h = run
    s = h.sizes()
    prm = h.params()
    a = h.assignment()
    off = h.patches(frames=False, binning=False)["patch_off"]
    n_p = np.diff(off)
    assert prm["nbv"].max() <= 30 and np.all((prm["nbv"] > 0) == (n_p > 0))
    assert np.all(prm["nbv"] <= np.maximum(n_p, 0))
    assert np.isfinite(prm["alpha"]).all()
    # BVs are points of their own patch
    bv_patch = np.repeat(np.arange(s.n_patches), prm["nbv"])
    src = off[bv_patch] + prm["bv_idx"]
    assert np.array_equal(prm["bv1"], a["st_x1"][src]) and np.array_equal(prm["bv2"], a["st_x2"][src])
    st = h.stats()
    assert st["n_add"] == s.n_claimed and st["n_first"] == int((n_p > 0).sum())
    assert st["n_sparse"] + st["n_full"] + st["n_first"] == st["n_add"]
    # decoder: sz^2 points per non-empty patch
    assert h.decompress_resident() == int((prm["nbv"] > 0).sum()) * 100
    # round trip on a sample of patches: the fitted surface reproduces the (mean-subtracted) heights
    rng = np.random.default_rng(0)
    se, cnt = 0.0, 0
    for p in rng.choice(np.nonzero(n_p > 50)[0], 60, replace=False):
        lo, hi = off[p], off[p + 1]
        f = h.predict(int(p), np.stack([a["st_x1"][lo:hi], a["st_x2"][lo:hi]], axis=1))
        se += float(((f - a["st_y"][lo:hi]) ** 2).sum())
        cnt += hi - lo
    assert np.sqrt(se / cnt) < 0.01  # 3 mm noise on planar patches


def test_checksum_is_independent_of_sharding(run):
    G, cloud, h = run
    def digest(parts):
        return hashlib.sha256(b"".join(hashlib.sha256(np.ascontiguousarray(p).tobytes()).digest() for p in parts)).hexdigest()
    whole = h.decompress()
    alpha = h.params()["alpha"]
    pieces, apieces = [], []
    for r in range(4):
        hs = G.Handle(res=F32(0.1), sz=10, capacity=30, shard_rank=r, shard_count=4)
        hs.compress(cloud)
        pieces.append(hs.decompress())
        apieces.append(hs.params()["alpha"])
    assert digest([np.concatenate(pieces)]) == digest([whole])
    assert np.array_equal(np.concatenate(apieces), alpha)


def test_results_do_not_depend_on_the_bucket_chain_mode(run):
    """The fit stage cuts the size-ordered patch list into two or three bucket chains and switches between the two modes on
    measured time (call 1 of a handle: three chains, call 2: two, call 3: three; gpc_api.cu chain_tune_after).  Scheduling only:
    parameters, BV index sets and event counters of the three calls are identical."""
    G, cloud, _ = run
    h = G.Handle(res=F32(0.1), sz=10, capacity=30)
    h.upload_cloud(cloud)
    got = []
    for _ in range(3):
        h.set_rand_offset(0)
        h.compress_resident()
        p, st = h.params(), h.stats()
        got.append((p["nbv"].copy(), p["bv_idx"].copy(), p["alpha"].copy(), [st[k] for k in ("n_add", "n_sparse", "n_full", "n_del_cap", "n_del_geo")], list(st["escalated"])))
    for g in got[1:]:
        assert np.array_equal(g[0], got[0][0]) and np.array_equal(g[1], got[0][1]) and np.array_equal(g[2], got[0][2])
        assert g[3] == got[0][3] and g[4] == got[0][4]


# ---- full-size BIT-COMPARES against the oracle (every array of the path, not just properties) ----------------------
def _oracle_threads():
    import os
    return max(1, os.cpu_count() or 1)


def test_c1_full_size_bit_compare(oracle_mod):
    """BASELINE configs[0] as stated: 100 000 points, gp_compressor(cloud, 0.15f, 20) (test_gp_compress.cpp:21), capacity 100,
    reference hyper-parameters, with the RGB field GP as the reference runs it: lattice, patch assignment, frames, shuffles,
    BV index sets, alpha, event counters, decoded cloud -- identical to the oracle."""
    import gp_compressor_b200 as G
    from test_gpu_compress import compare_all
    cloud = synth.c1_planar_bumps(100_000, seed=1)
    compare_all(G, oracle_mod, cloud, res=F32(0.15), sz=20, capacity=100)
    h = G.Handle(res=F32(0.15), sz=20, capacity=100, rgb=1)
    o = oracle_mod.Oracle(res=F32(0.15), sz=20, capacity=100, rgb=1, threads=_oracle_threads())
    h.compress(cloud)
    want = o.compress(cloud)
    assert np.array_equal(h.params()["bv_idx"], want["bv_idx"]) and np.array_equal(h.params()["alpha"], want["alpha"])
    rg, ro = h.params_rgb(), o.rgb_result()
    assert np.array_equal(rg["nbv"], ro["nbv"]) and np.array_equal(rg["bv_idx"], ro["bv_idx"]) and np.array_equal(rg["alpha"], ro["alpha"])
    co, _ = o.decode(want_heights=False)
    assert np.array_equal(h.decompress(), co)


def test_c2_full_size_bit_compare(run, oracle_mod):
    """BASELINE configs[1] as stated (5 M points, res 0.1, capacity 30, reference hyper-parameters): the whole path against
    the oracle run on all host cores -- identical patch assignment, per-patch order, BV index sets, alpha and decoded cloud."""
    G, cloud, h = run
    h.set_rand_offset(0)
    h.compress(cloud)
    o = oracle_mod.Oracle(res=F32(0.1), sz=10, capacity=30, threads=_oracle_threads())
    want = o.compress(cloud)
    bo, bg, ag = o.binning(), h.patches(), h.assignment()
    assert bg["depth"] == bo["depth"] and np.array_equal(bg["lattice_min"], bo["lattice_min"])
    for k in ("leaf_code", "leaf_center", "leaf_ncand", "patch_off", "leaf_R", "leaf_quat", "leaf_mean", "leaf_rgbmean"):
        assert np.array_equal(bg[k], bo[k], equal_nan=True), k   # patches without claimed points hold NaN means (gp_compressor.cpp:101-102)
    for k in ("owner", "st_idx", "st_x1", "st_x2", "st_y"):
        assert np.array_equal(ag[k], bo[k]), k
    assert np.array_equal(ag["perm"], want["perm"])
    got = h.params()
    assert np.array_equal(got["nbv"], want["nbv"]) and np.array_equal(got["bv_idx"], want["bv_idx"])
    assert np.array_equal(got["alpha"], want["alpha"])
    gs, os_ = h.stats(), o.stats()
    for k in ("n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo"):
        assert gs[k] == os_[k], k
    co, ho = o.decode()
    assert h.decompress_resident() == ho.size and np.array_equal(h.heights(), ho)
    assert np.array_equal(h.decompress(), co)
