"""CUDA path (through the C ABI) against the golden vectors produced by the reference's own SOGP source
(tests/golden/ref_sogp.npz, see tests/golden/make_golden.py).  Discrete results must be identical on the
well-conditioned hyper-set; parameters agree to rounding level scaled by conditioning (the reference run
sums sequentially and uses libm exp, the kernels use the canonical order and exp)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_sogp.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if "/" in k and not k.startswith(("field_", "eval_", "feval_", "cont_"))})
EVAL_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.startswith("eval_")})


@pytest.mark.parametrize("name", CASES)
def test_cuda_sogp_matches_reference_source_vectors(name):
    import gp_compressor_b200 as G
    g = {k.split("/")[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    n, cap, roff, N = (int(v) for v in g["meta"])
    p0, l_sq, s0 = (float(v) for v in g["hyper"])
    h = G.Handle(capacity=cap, sigmaf_sq=p0, l_sq=l_sq, s0=s0, rgb_rand=0)
    h.set_rand_offset(roff)
    h.fit_patches([0, n], g["x1"], g["x2"], g["y"])
    p = h.params()
    f = h.predict(0, g["pred"])
    if l_sq < 1.0:
        assert int(p["nbv"][0]) == N
        assert np.array_equal(p["bv1"], g["bv1"]) and np.array_equal(p["bv2"], g["bv2"])  # same BVs in the same slots
        assert np.abs(p["alpha"] - g["alpha"]).max() <= 5e-6 * np.abs(g["alpha"]).max()
        assert np.abs(f - g["f"]).max() <= 1e-6 * np.abs(g["f"]).max()
    else:
        assert np.abs(f - g["f"]).max() <= 2e-4


@pytest.mark.parametrize("name", EVAL_CASES)
def test_cuda_evaluate_matches_reference_source_vectors(name):
    """Rows N2 / N4 through the C ABI (gpc_evaluate_patches) against the reference's own predict_measurements /
    compute_likelihoods / compute_derivatives outputs."""
    import gp_compressor_b200 as G
    from test_oracle_vs_reference_source import check_eval, eval_tol
    g = {k.split("/")[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    n, cap, roff, N = (int(v) for v in g["meta"])
    p0, l_sq, s0 = (float(v) for v in g["hyper"])
    h = G.Handle(capacity=cap, sigmaf_sq=p0, l_sq=l_sq, s0=s0, rgb_rand=0, keep_state=1)
    h.set_rand_offset(roff)
    h.fit_patches([0, n], g["x1"], g["x2"], g["y"])
    assert int(h.params()["nbv"][0]) == N
    q = ([0, g["ex"].shape[0]], g["ex"][:, 0].copy(), g["ex"][:, 1].copy(), g["ey"])
    check_eval(h.evaluate(*q), h.evaluate(*q, conf=True), g, eval_tol(name))


FEVAL_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.startswith("feval_")})
CONT_CASES = sorted({k.split("/")[0] for k in GOLD.files if k.startswith("cont_")})


@pytest.mark.parametrize("name", CONT_CASES)
def test_cuda_continued_fit_matches_reference_source_vectors(name):
    """gpc_fit_patches + gpc_add_measurements against the reference's own add_measurements called twice."""
    import gp_compressor_b200 as G
    g = {k.split("/")[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    n, cap, roff, N, n1 = (int(v) for v in g["meta"])
    p0, l_sq, s0 = (float(v) for v in g["hyper"])
    h = G.Handle(capacity=cap, sigmaf_sq=p0, l_sq=l_sq, s0=s0, rgb_rand=0, keep_state=1)
    h.set_rand_offset(roff)
    h.fit_patches([0, n1], g["x1"][:n1], g["x2"][:n1], g["y"][:n1])
    h.add_measurements([0, n - n1], g["x1"][n1:], g["x2"][n1:], g["y"][n1:])
    p = h.params()
    assert int(p["nbv"][0]) == N
    assert np.array_equal(p["bv1"], g["bv1"]) and np.array_equal(p["bv2"], g["bv2"])
    assert np.abs(p["alpha"] - g["alpha"]).max() <= 5e-6 * np.abs(g["alpha"]).max()
