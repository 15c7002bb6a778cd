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
