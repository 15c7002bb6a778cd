"""ctypes wrapper of oracle/_ref/libref_sogp.so (the reference's own sparse_gp.hpp / rbf_kernel.cpp /
gaussian_noise.cpp compiled over oracle/eigen_shim).  Test infrastructure."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_sogp.so")
_LIB = None


def available():
    if os.path.isdir("/root/reference/src"):
        from . import ref_build
        ref_build.build()
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.ref_sogp_fit.restype = C.c_int
        L.ref_sogp_fit.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_ulonglong, C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4
        L.ref_shuffle.argtypes = [C.c_int, C.c_ulonglong, C.c_void_p]
        L.ref_field_fit.restype = C.c_int
        L.ref_field_fit.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_ulonglong, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
        L.ref_sogp_evaluate.restype = C.c_int
        L.ref_sogp_evaluate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_ulonglong, C.c_int] + [C.c_void_p] * 8
        L.ref_sogp_fit_twice.restype = C.c_int
        L.ref_sogp_fit_twice.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_ulonglong, C.c_int] + [C.c_void_p] * 5
        L.ref_field_evaluate.restype = C.c_int
        L.ref_field_evaluate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_ulonglong, C.c_int] + [C.c_void_p] * 8
        L.ref_kernel.restype = C.c_double
        L.ref_kernel.argtypes = [C.c_double] * 6
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fit(x1, x2, y, capacity=100, s0=float(np.float32(1e-1)), sigmaf_sq=100.0, l_sq=1.0, eps_tol=float(np.float32(1e-6)),
        rand_offset=0, pred=None):
    """sparse_gp<rbf_kernel, gaussian_noise>(capacity, s0).add_measurements(X, y) of the reference itself."""
    x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
    n = x1.size
    mx = (capacity if capacity > 0 else n) + 2
    alpha, b1, b2 = np.zeros(mx), np.zeros(mx), np.zeros(mx)
    Cm, Qm = np.zeros(mx * mx), np.zeros(mx * mx)
    if pred is None:
        pred = np.zeros((0, 2))
    pred = np.ascontiguousarray(pred, dtype=np.float64).reshape(-1, 2)
    p1, p2 = np.ascontiguousarray(pred[:, 0]), np.ascontiguousarray(pred[:, 1])
    f, sg = np.zeros(pred.shape[0]), np.zeros(pred.shape[0])
    N = lib().ref_sogp_fit(n, _p(x1), _p(x2), _p(y), capacity, s0, sigmaf_sq, l_sq, eps_tol, rand_offset, mx, _p(alpha), _p(b1), _p(b2),
                           _p(Cm), _p(Qm), pred.shape[0], _p(p1), _p(p2), _p(f), _p(sg))
    assert N >= 0
    return dict(N=N, alpha=alpha[:N].copy(), bv1=b1[:N].copy(), bv2=b2[:N].copy(), C=Cm[:N * N].reshape(N, N).copy(),
                Q=Qm[:N * N].reshape(N, N).copy(), f=f, sigma=sg)


def evaluate(x1, x2, y, ex, ey, capacity=100, s0=float(np.float32(1e-1)), sigmaf_sq=100.0, l_sq=1.0,
             eps_tol=float(np.float32(1e-6)), rand_offset=0):
    """Fit, then the reference's predict_measurements (sigma / conf), compute_likelihoods and compute_derivatives at ex (m x 2), ey."""
    x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
    ex = np.ascontiguousarray(ex, dtype=np.float64).reshape(-1, 2)
    ey = np.ascontiguousarray(ey, dtype=np.float64)
    m = ex.shape[0]
    e1, e2 = np.ascontiguousarray(ex[:, 0]), np.ascontiguousarray(ex[:, 1])
    f, sg, cf, lk, dX = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(3 * m)
    N = lib().ref_sogp_evaluate(x1.size, _p(x1), _p(x2), _p(y), capacity, s0, sigmaf_sq, l_sq, eps_tol, rand_offset, m, _p(e1), _p(e2),
                                _p(ey), _p(f), _p(sg), _p(cf), _p(lk), _p(dX))
    return dict(N=N, f=f, sigma=sg, conf=cf, lik=lk, dX=dX.reshape(m, 3))


def fit_twice(x1, x2, y, n1, capacity=100, s0=float(np.float32(1e-1)), sigmaf_sq=100.0, l_sq=1.0, eps_tol=float(np.float32(1e-6)),
              rand_offset=0):
    """Two successive add_measurements calls (the first n1 points, then the rest) on one process of the reference."""
    x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
    n = x1.size
    mx = (capacity if capacity > 0 else n) + 2
    alpha, b1, b2 = np.zeros(mx), np.zeros(mx), np.zeros(mx)
    Cm, Qm = np.zeros(mx * mx), np.zeros(mx * mx)
    N = lib().ref_sogp_fit_twice(n1, n - n1, _p(x1), _p(x2), _p(y), capacity, s0, sigmaf_sq, l_sq, eps_tol, rand_offset, mx, _p(alpha),
                                 _p(b1), _p(b2), _p(Cm), _p(Qm))
    assert N >= 0
    return dict(N=N, alpha=alpha[:N].copy(), bv1=b1[:N].copy(), bv2=b2[:N].copy(), C=Cm[:N * N].reshape(N, N).copy(),
                Q=Qm[:N * N].reshape(N, N).copy())


def field_evaluate(x1, x2, Y, ex, EY, capacity=100, s0=float(np.float32(1e2)), sigmaf_sq=100.0, l_sq=1.0,
                   eps_tol=float(np.float32(1e-4)), rand_offset=0):
    """Field GP: fit, then the reference's predict_measurements (sigma / conf), compute_likelihoods, compute_derivatives."""
    x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2))
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, 3)
    ex = np.ascontiguousarray(ex, dtype=np.float64).reshape(-1, 2)
    EY = np.ascontiguousarray(EY, dtype=np.float64).reshape(-1, 3)
    m = ex.shape[0]
    e1, e2 = np.ascontiguousarray(ex[:, 0]), np.ascontiguousarray(ex[:, 1])
    f, sg, cf, lk, dX = np.zeros(3 * m), np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(3 * m)
    N = lib().ref_field_evaluate(x1.size, _p(x1), _p(x2), _p(Y), capacity, s0, sigmaf_sq, l_sq, eps_tol, rand_offset, m, _p(e1), _p(e2),
                                 _p(EY), _p(f), _p(sg), _p(cf), _p(lk), _p(dX))
    return dict(N=N, f=f.reshape(m, 3), sigma=sg, conf=cf, lik=lk, dX=dX.reshape(m, 3))


def field_fit(x1, x2, Y, capacity=100, s0=float(np.float32(1e2)), sigmaf_sq=100.0, l_sq=1.0, eps_tol=float(np.float32(1e-4)),
              rand_offset=0, pred=None):
    """sparse_gp_field<rbf_kernel, gaussian_noise_3d>(capacity, s0).add_measurements(X, Y) of the reference itself."""
    x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2))
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, 3)
    n = x1.size
    mx = (capacity if capacity > 0 else n) + 2
    alpha, b1, b2 = np.zeros(3 * mx), np.zeros(mx), np.zeros(mx)
    if pred is None:
        pred = np.zeros((0, 2))
    pred = np.ascontiguousarray(pred, dtype=np.float64).reshape(-1, 2)
    p1, p2 = np.ascontiguousarray(pred[:, 0]), np.ascontiguousarray(pred[:, 1])
    f = np.zeros(3 * pred.shape[0])
    N = lib().ref_field_fit(n, _p(x1), _p(x2), _p(Y), capacity, s0, sigmaf_sq, l_sq, eps_tol, rand_offset, mx, _p(alpha), _p(b1), _p(b2),
                            pred.shape[0], _p(p1), _p(p2), _p(f))
    assert N >= 0
    return dict(N=N, alpha=alpha[:3 * N].reshape(N, 3).copy(), bv1=b1[:N].copy(), bv2=b2[:N].copy(), f=f.reshape(-1, 3))


def shuffle(n, rand_offset=0):
    out = np.zeros(n, dtype=np.int32)
    lib().ref_shuffle(n, rand_offset, _p(out))
    return out


def kernel(sigmaf_sq, l_sq, a, b):
    return lib().ref_kernel(sigmaf_sq, l_sq, a[0], a[1], b[0], b[1])


# ---- gp_compressor::compute_rotation / project_points (gp_compressor.cpp:29-118) from the reference's own text ----
SO_FRAMES = os.path.join(_HERE, "_ref", "libref_frames.so")
_LIBF = None


def frames_available():
    if os.path.isdir("/root/reference/src"):
        from . import ref_build
        ref_build.build_frames()
    return os.path.exists(SO_FRAMES)


def _libf():
    global _LIBF
    if _LIBF is None:
        L = C.CDLL(SO_FRAMES)
        L.ref_compute_rotation.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_project_points.restype = C.c_int
        L.ref_project_points.argtypes = [C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 10
        _LIBF = L
    return _LIBF


def compute_rotation(pts):
    """pts: m x 3 doubles (the candidates of one leaf).  Returns R (3 x 3) as the reference computes it."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    R = np.zeros(9)
    _libf().ref_compute_rotation(pts.shape[0], _p(pts), _p(R))
    return R.reshape(3, 3)


def project_points(res, sz, pts, cols, index_search, occupied, R, center):
    """One call of the reference's project_points.  occupied (int32 array over the cloud) is updated in place.
    Returns dict(claimed = positions in the candidate list, local = n x 3 (height - mean, x1, x2), colour = n x 3 centred,
    center = updated voxel centre, rgb_mean)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    cols = np.ascontiguousarray(cols, dtype=np.float64).reshape(-1, 3)
    idx = np.ascontiguousarray(index_search, dtype=np.int32)
    assert occupied.dtype == np.int32 and occupied.flags.c_contiguous
    m = pts.shape[0]
    Rr = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    c3 = np.ascontiguousarray(center, dtype=np.float64).copy()
    claimed = np.zeros(max(m, 1), dtype=np.int32)
    local, colour, rgbm = np.zeros(3 * max(m, 1)), np.zeros(3 * max(m, 1)), np.zeros(3)
    n = _libf().ref_project_points(float(res), int(sz), m, _p(pts), _p(cols), _p(idx), _p(occupied), _p(Rr), _p(c3), _p(claimed),
                                   _p(local), _p(colour), _p(rgbm))
    assert n >= 0
    return dict(claimed=claimed[:n].copy(), local=local[:3 * n].reshape(n, 3).copy(), colour=colour[:3 * n].reshape(n, 3).copy(),
                center=c3, rgb_mean=rgbm)
