"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol include/gpc.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from gp_compressor_b200 import binding
    return binding


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "gpc.h")).read()
    declared = set(re.findall(r"\b(gpc_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"gpc_config", "gpc_handle", "gpc_sizes", "gpc_stats"}
    assert declared == set(built.SYMBOLS), declared ^ set(built.SYMBOLS)
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_defaults_are_the_reference_constructor_defaults(built):
    import numpy as np
    cfg = built.default_config()
    assert cfg.res == float(np.float32(0.1)) and cfg.sz == 10          # gp_compressor.h:65
    assert cfg.capacity == 100 and cfg.s0 == float(np.float32(1e-1))    # sparse_gp.h:48
    assert cfg.eps_tol == float(np.float32(1e-6))                       # sparse_gp.hpp:31
    assert cfg.sigmaf_sq == 100.0 and cfg.l_sq == 1.0                   # rbf_kernel.h:24
    assert ctypes.sizeof(built.GpcConfig) == 104
    assert cfg.decode_separable == 0  # the reference's direct kernel evaluation is the default (rbf_kernel.cpp:15-18)
    assert cfg.rgb == 0 and cfg.rgb_s0 == float(np.float32(1e2)) and cfg.rgb_eps_tol == float(np.float32(1e-4))  # sparse_gp_field.h:43, .hpp:16


def test_no_cpu_fallback(built):
    """Without a CUDA device gpc_create must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.GpcError):
        built.Handle()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "gp_compressor_b200")
    for dp, _, fns in os.walk(pkg):
        if "build" in dp.split(os.sep)[-1:]:
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libgpc_oracle" not in txt, fn
