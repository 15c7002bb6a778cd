"""Reference-order mode of the oracle (`ref_order=1`) against THE REFERENCE'S OWN SOURCE.

oracle/_ref is sparse_gp.hpp / rbf_kernel.cpp / gaussian_noise.cpp compiled from /root/reference over oracle/eigen_shim
(sequential sums, libm exp, per-element divisions).  In reference-order mode the oracle evaluates every expression of
sparse_gp::add / delete_bv / predict (sparse_gp.hpp:89-351) in that same order, so the two must agree BIT FOR BIT -- on
every committed golden case, including the reference's default hyper-parameters (rbf 100 / 1), where the canonical
(kernel-shaped) order of the product selects a different BV set because cond(Q) ~ 1e8.  This is what demonstrates that
the canonical mode differs from the reference only by summation order / exp rounding and not by the recursion itself;
tools/order_drift.py measures how far the two orders drift on the C2 streams.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_source as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sogp.npz"))
FIT_CASES = ["ref_defaults_cap100", "ref_defaults_cap30", "bind_cap12", "bind_cap30", "bind_cap60", "tiny_n3"]


def _fit(name, **kw):
    n, cap, roff, N = (int(v) for v in G[name + "/meta"][:4])
    hyp = G[name + "/hyper"]
    o = O.Oracle(capacity=cap, sigmaf_sq=hyp[0], l_sq=hyp[1], s0=hyp[2], ref_order=1, **kw)
    o.set_rand_offset(roff)
    r = o.fit_patches(np.array([0, n]), G[name + "/x1"], G[name + "/x2"], G[name + "/y"], dump=True)
    return o, r, N


@pytest.mark.parametrize("name", FIT_CASES)
def test_reference_order_is_bit_equal_to_reference_source_vectors(name):
    o, r, N = _fit(name)
    assert int(r["nbv"][0]) == N
    assert np.array_equal(r["bv1"][:N], G[name + "/bv1"]) and np.array_equal(r["bv2"][:N], G[name + "/bv2"])
    assert np.array_equal(r["alpha"][:N], G[name + "/alpha"])
    assert np.array_equal(r["C"][:N * N], G[name + "/C"].ravel())
    assert np.array_equal(r["Q"][:N * N], G[name + "/Q"].ravel())
    f, sg = o.predict(0, G[name + "/pred"], sigma=True)
    assert np.array_equal(f, G[name + "/f"])
    assert np.array_equal(sg, G[name + "/sigma"])


@pytest.mark.parametrize("name", ["cont_cap25", "cont_cap12"])
def test_reference_order_continued_fit_bit_equal(name):
    n, cap, roff, N, n1 = (int(v) for v in G[name + "/meta"])
    hyp = G[name + "/hyper"]
    o = O.Oracle(capacity=cap, sigmaf_sq=hyp[0], l_sq=hyp[1], s0=hyp[2], ref_order=1, rgb_rand=0)  # the harness has no field GP drawing in between
    o.set_rand_offset(roff)
    x1, x2, y = G[name + "/x1"], G[name + "/x2"], G[name + "/y"]
    o.fit_patches(np.array([0, n1]), x1[:n1], x2[:n1], y[:n1], dump=True)
    r = o.add_measurements(np.array([0, n - n1]), x1[n1:], x2[n1:], y[n1:])
    assert int(r["nbv"][0]) == N
    assert np.array_equal(r["bv1"][:N], G[name + "/bv1"])
    assert np.array_equal(r["alpha"][:N], G[name + "/alpha"])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref is built only where /root/reference exists")
@pytest.mark.parametrize("seed,n,cap,hyper", [
    (1, 700, 100, "ref"), (2, 1500, 30, "ref"), (3, 64, 100, "ref"), (4, 900, 15, "ref"),
    (5, 500, 20, "bind"), (6, 800, 45, "bind"), (7, 2, 5, "bind"), (8, 1, 5, "ref")])
def test_reference_order_live_against_reference_source(seed, n, cap, hyper):
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(-0.05, 0.05, n)
    x2 = rng.uniform(-0.05, 0.05, n)
    y = 0.03 * np.sin(35 * x1) * np.cos(28 * x2) - 0.4 * x2 + rng.normal(0, 0.003, n)
    hyp = dict(sigmaf_sq=100.0, l_sq=1.0, s0=float(np.float32(1e-1))) if hyper == "ref" else dict(sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)
    pred = rng.uniform(-0.05, 0.05, (30, 2))
    want = R.fit(x1, x2, y, capacity=cap, rand_offset=11 * seed, pred=pred, **hyp)
    o = O.Oracle(capacity=cap, ref_order=1, **hyp)
    o.set_rand_offset(11 * seed)
    r = o.fit_patches(np.array([0, n]), x1, x2, y, dump=True)
    N = want["N"]
    assert int(r["nbv"][0]) == N
    assert np.array_equal(r["bv1"][:N], want["bv1"]) and np.array_equal(r["alpha"][:N], want["alpha"])
    assert np.array_equal(r["C"][:N * N], want["C"].ravel()) and np.array_equal(r["Q"][:N * N], want["Q"].ravel())
    f, sg = o.predict(0, pred, sigma=True)
    assert np.array_equal(f, want["f"]) and np.array_equal(sg, want["sigma"])
