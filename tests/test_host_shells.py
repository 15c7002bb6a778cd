"""The host C++ shells (gp_compressor_b200/host/*.h) keep the reference's class names and signatures
over the C ABI.  CPU: the driver modelled on src/test_gp_compress.cpp compiles, links against the shared
library and fails loudly without a device.  GPU: it reproduces what the ctypes path produces."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_api")


def build_driver():
    import __graft_entry__ as ge
    ge.build()
    libdir = os.path.join(ROOT, "gp_compressor_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(libdir, "host"), os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
                           "-o", BIN, "-L" + libdir, "-lgpc_b200", "-Wl,-rpath," + libdir])
    return BIN


def test_driver_compiles_and_has_no_cpu_fallback():
    import torch
    exe = build_driver()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe, os.devnull], capture_output=True, text=True)
    assert out.returncode == 1 and "no usable sm_100 CUDA device" in out.stdout


@pytest.mark.gpu
def test_driver_matches_ctypes_path(tmp_path):
    import gp_compressor_b200 as G
    from gp_compressor_b200 import synth
    exe = build_driver()
    cloud = synth.c1_planar_bumps(20000, seed=6)
    path = tmp_path / "cloud.bin"
    cloud.tofile(path)
    dec_path = tmp_path / "decoded.bin"
    out = subprocess.run([exe, str(path), str(dec_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.splitlines() if l.split(" ")[0] in ("CLOUD", "SOGP", "SOGP2", "EVAL", "KSVD_SHELL"))
    h = G.Handle(res=float(np.float32(0.15)), sz=20, rgb=1)  # the literals of test_gp_compress.cpp:21; the shell enables the RGB GP
    h.compress(cloud)
    dec = h.decompress()
    dec_cpp = np.fromfile(dec_path, dtype=np.uint8).reshape(-1, 32)
    assert int(lines["CLOUD"].split()[0]) == dec.shape[0]
    assert np.array_equal(dec_cpp, dec)
    nb, f = lines["SOGP"].split()
    assert int(nb) == 2 and abs(float(f) - 0.0050018719900836684) < 1e-10   # SURVEY.md section 4 known answer
    assert lines["KSVD_SHELL"].strip() == "1"
    # predict_measurements (sigma, conf) / compute_likelihoods / compute_derivatives of the shell == the ctypes path
    g = G.Handle(capacity=2, s0=float(np.float32(1e-1)), shuffle=0, rgb_rand=0, keep_state=1)
    pts = np.array([[0, 0, .01], [.05, 0, .02], [0, .05, -.01], [.05, .05, 0]])
    g.fit_patches(np.array([0, 4]), pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy())
    q = (np.array([0, 1]), np.array([0.025]), np.array([0.025]), np.array([0.004]))
    e, ec = g.evaluate(*q), g.evaluate(*q, conf=True)
    want = [e["sigma"][0], ec["sigma"][0], e["lik"][0], *e["dX"][0]]
    got = [float(v) for v in lines["EVAL"].split()]
    assert got == [float(repr(float(w))) for w in want]
    # the shell's second add_measurements == gpc_add_measurements through ctypes
    more = np.array([[.02, .03, .015], [.04, .01, .005], [.01, .04, -.002]])
    g.add_measurements(np.array([0, 3]), more[:, 0].copy(), more[:, 1].copy(), more[:, 2].copy())
    f2, s2 = g.predict(0, np.array([[0.025, 0.025]]), sigma=True)
    nb2, ff, ss = lines["SOGP2"].split()
    assert int(nb2) == int(g.params()["nbv"][0]) and float(ff) == float(repr(float(f2[0]))) and float(ss) == float(repr(float(s2[0])))
