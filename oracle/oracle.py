"""ctypes wrapper around the CPU oracle (oracle/libgpc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (gp_compressor_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcConfig(C.Structure):
    _fields_ = [("res", C.c_double), ("sz", C.c_int), ("capacity", C.c_int), ("s0", C.c_double),
                ("eps_tol", C.c_double), ("sigmaf_sq", C.c_double), ("l_sq", C.c_double),
                ("leaf_order", C.c_int), ("shuffle", C.c_int), ("rgb_rand", C.c_int), ("threads", C.c_int),
                ("rgb", C.c_int), ("ref_order", C.c_int), ("rgb_s0", C.c_double), ("rgb_eps_tol", C.c_double),
                ("decode_separable", C.c_int), ("pad2", C.c_int)]


class OrcSizes(C.Structure):
    _fields_ = [("n_in", C.c_int64), ("n_leaves", C.c_int64), ("n_claimed", C.c_int64),
                ("n_bv_total", C.c_int64), ("n_dump", C.c_int64), ("depth", C.c_uint32),
                ("pad", C.c_uint32), ("mn", C.c_double * 3)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo")] + \
               [(n, C.c_double) for n in ("sumN", "sumN2_common", "sumN2_sparse", "sumN2_full", "sumN2_del",
                                          "t_project", "t_train", "t_decode")]


def build(force=False):
    so = os.path.join(_HERE, "libgpc_oracle.so")
    src = os.path.join(_HERE, "gpc_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_exp.restype = C.c_double
        L.orc_exp.argtypes = [C.c_double]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_rand_offset.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_get_rand_offset.restype = C.c_uint64
        L.orc_get_rand_offset.argtypes = [C.c_void_p]
        L.orc_project.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_train_projected.argtypes = [C.c_void_p, C.c_int]
        L.orc_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_fit_patches.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int]
        L.orc_fit_patches_rgb.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + [C.c_int]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
        L.orc_decode.restype = C.c_int64
        L.orc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_predict.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_add_measurements.restype = C.c_int
        L.orc_add_measurements.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        L.orc_evaluate_rgb.restype = C.c_int
        L.orc_evaluate_rgb.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
        L.orc_evaluate.restype = C.c_int
        L.orc_evaluate.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
        L.orc_get_sizes.argtypes = [C.c_void_p, C.POINTER(OrcSizes)]
        L.orc_ptr.restype = C.c_void_p
        L.orc_ptr.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_stats.argtypes = [C.c_void_p, C.POINTER(OrcStats)]
        L.orc_get_stats_rgb.argtypes = [C.c_void_p, C.POINTER(OrcStats)]
        L.orc_rgb_bv_total.restype = C.c_int64
        L.orc_rgb_bv_total.argtypes = [C.c_void_p]
        L.orc_rand_stream.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
        L.orc_shuffles.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_exp_array.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_rotation_from_sums.argtypes = [C.c_void_p] * 3
        L.orc_quat_roundtrip.argtypes = [C.c_void_p] * 3
        L.orc_lattice.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def exp_array(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().orc_exp_array(_p(x), _p(out), x.size)
    return out


def rand_stream(offset, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_rand_stream(offset, n, _p(out))
    return out


def shuffles(offset, sizes):
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    out = np.empty(int(sizes.sum()), dtype=np.int32)
    lib().orc_shuffles(offset, _p(sizes), sizes.size, _p(out))
    return out


def lattice(cloud32, res):
    mn = np.zeros(3)
    depth = C.c_uint32(0)
    lib().orc_lattice(_p(cloud32), cloud32.shape[0], float(res), _p(mn), C.byref(depth))
    return mn, depth.value


def rotation_from_sums(sums, c):
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    R = np.zeros(9)
    lib().orc_rotation_from_sums(_p(sums), _p(c), _p(R))
    return R.reshape(3, 3)


def quat_roundtrip(R):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    q = np.zeros(4)
    R2 = np.zeros(9)
    lib().orc_quat_roundtrip(_p(R), _p(q), _p(R2))
    return q, R2.reshape(3, 3)


_FIELDS = {
    # name: (dtype, size function of sizes)
    "leaf_code": (np.uint64, lambda s: s.n_leaves), "leaf_center": (np.float32, lambda s: s.n_leaves * 3),
    "leaf_ncand": (np.int32, lambda s: s.n_leaves), "leaf_R": (np.float64, lambda s: s.n_leaves * 9),
    "leaf_quat": (np.float64, lambda s: s.n_leaves * 4), "leaf_mean": (np.float64, lambda s: s.n_leaves * 3),
    "leaf_rgbmean": (np.float64, lambda s: s.n_leaves * 3), "patch_off": (np.int64, lambda s: s.n_leaves + 1),
    "owner": (np.int32, lambda s: s.n_in), "st_idx": (np.int32, lambda s: s.n_claimed),
    "st_x1": (np.float64, lambda s: s.n_claimed), "st_x2": (np.float64, lambda s: s.n_claimed),
    "st_y": (np.float64, lambda s: s.n_claimed),
}


class Oracle:
    """One oracle handle = one gp_compressor object of the reference."""

    def __init__(self, res=float(np.float32(0.1)), sz=10, capacity=100, s0=float(np.float32(1e-1)),
                 eps_tol=float(np.float32(1e-6)), sigmaf_sq=100.0, l_sq=1.0, leaf_order=0, shuffle=1,
                 rgb_rand=1, threads=1, rgb=0, rgb_s0=float(np.float32(1e2)), rgb_eps_tol=float(np.float32(1e-4)),
                 ref_order=0, decode_separable=0):
        """ref_order=1: the height GPs run in the reference source's own evaluation order (bit-equal to oracle/_ref);
        decode_separable=1: the grid decode follows the product's flagged separable mode instead of the direct kernel."""
        self.cfg = OrcConfig(res, sz, capacity, s0, eps_tol, sigmaf_sq, l_sq, leaf_order, shuffle, rgb_rand, threads, rgb, ref_order,
                             rgb_s0, rgb_eps_tol, decode_separable, 0)
        self.h = lib().orc_create(C.byref(self.cfg))
        self._np = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + lib().orc_last_error(self.h).decode())

    def set_rand_offset(self, off):
        lib().orc_set_rand_offset(self.h, off)

    def rand_offset(self):
        return lib().orc_get_rand_offset(self.h)

    def sizes(self):
        s = OrcSizes()
        lib().orc_get_sizes(self.h, C.byref(s))
        return s

    def stats(self):
        s = OrcStats()
        lib().orc_get_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in OrcStats._fields_}

    def stats_rgb(self):
        s = OrcStats()
        lib().orc_get_stats_rgb(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in OrcStats._fields_}

    def rgb_result(self):
        """RGB field GP (sparse_gp_field) parameters of the last compress with rgb=1."""
        NP = self._np
        T = lib().orc_rgb_bv_total(self.h)
        S = self.sizes().n_claimed
        return {"nbv": self._arr("rgb_nbv", np.int32, NP), "bv_off": self._arr("rgb_bv_off", np.int64, NP + 1),
                "bv_idx": self._arr("rgb_bv_idx", np.int32, T), "bv1": self._arr("rgb_bv1", np.float64, T),
                "bv2": self._arr("rgb_bv2", np.float64, T), "alpha": self._arr("rgb_alpha", np.float64, 3 * T),
                "perm": self._arr("st_perm_rgb", np.int32, S), "colours": self._arr("st_c", np.float64, 3 * S)}

    def _arr(self, name, dtype, n):
        ptr = lib().orc_ptr(self.h, name.encode())
        if n == 0 or not ptr:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(int(n),)).copy()

    def project(self, cloud32):
        assert cloud32.dtype == np.uint8 and cloud32.ndim == 2 and cloud32.shape[1] == 32 and cloud32.flags.c_contiguous
        self._check(lib().orc_project(self.h, _p(cloud32), cloud32.shape[0]))
        return self.binning()

    def binning(self):
        s = self.sizes()
        out = {k: self._arr(k, dt, fn(s)) for k, (dt, fn) in _FIELDS.items()}
        out["depth"] = s.depth
        out["lattice_min"] = np.array(list(s.mn))
        out["n_leaves"] = s.n_leaves
        out["n_claimed"] = s.n_claimed
        return out

    def train_projected(self, dump=False):
        self._check(lib().orc_train_projected(self.h, int(dump)))
        self._np = self.sizes().n_leaves
        return self.fit_result(dump)

    def compress(self, cloud32, dump=False):
        assert cloud32.dtype == np.uint8 and cloud32.shape[1] == 32 and cloud32.flags.c_contiguous
        self._check(lib().orc_compress(self.h, _p(cloud32), cloud32.shape[0], int(dump)))
        self._np = self.sizes().n_leaves
        return self.fit_result(dump)

    def fit_patches(self, off, x1, x2, y, dump=False, colours=None):
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
        self._np = off.size - 1
        self._ntot = int(off[-1])
        if colours is not None:
            colours = np.ascontiguousarray(colours, dtype=np.float64).reshape(-1, 3)
            self._check(lib().orc_fit_patches_rgb(self.h, self._np, _p(off), _p(x1), _p(x2), _p(y), _p(colours), int(dump)))
            return self.fit_result(dump, ntot=self._ntot)
        self._check(lib().orc_fit_patches(self.h, self._np, _p(off), _p(x1), _p(x2), _p(y), int(dump)))
        return self.fit_result(dump, ntot=self._ntot)

    def add_measurements(self, off, x1, x2, y):
        """sparse_gp::add_measurements again on the patches of the previous fit_patches(dump=True): continues the state."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
        assert off.size - 1 == self._np
        self._ntot = int(off[-1])
        self._check(lib().orc_add_measurements(self.h, self._np, _p(off), _p(x1), _p(x2), _p(y)))
        return self.fit_result(True, ntot=self._ntot)

    def fit_result(self, dump=False, ntot=None):
        s = self.sizes()
        NP = self._np
        T = s.n_bv_total
        r = {
            "nbv": self._arr("nbv", np.int32, NP), "bv_off": self._arr("bv_off", np.int64, NP + 1),
            "bv_idx": self._arr("bv_idx", np.int32, T), "bv1": self._arr("bv1", np.float64, T),
            "bv2": self._arr("bv2", np.float64, T), "alpha": self._arr("alpha", np.float64, T),
            "flags": self._arr("fit_flags", np.int32, NP),
            "perm": self._arr("st_perm", np.int32, s.n_claimed if ntot is None else ntot),
        }
        if dump:
            r["dump_off"] = self._arr("dump_off", np.int64, NP + 1)
            r["C"] = self._arr("dumpC", np.float64, s.n_dump)
            r["Q"] = self._arr("dumpQ", np.float64, s.n_dump)
        return r

    def set_params(self, nbv, bv1, bv2, alpha, quat=None, mean=None, rgbmean=None):
        nbv = np.ascontiguousarray(nbv, dtype=np.int32)
        bv1, bv2, alpha = (np.ascontiguousarray(a, dtype=np.float64) for a in (bv1, bv2, alpha))
        self._np = nbv.size
        if quat is not None:
            quat, mean, rgbmean = (np.ascontiguousarray(a, dtype=np.float64) for a in (quat, mean, rgbmean))
            self._check(lib().orc_set_params(self.h, nbv.size, _p(nbv), _p(bv1), _p(bv2), _p(alpha), _p(quat), _p(mean), _p(rgbmean)))
        else:
            self._check(lib().orc_set_params(self.h, nbv.size, _p(nbv), _p(bv1), _p(bv2), _p(alpha), None, None, None))

    def decode(self, want_cloud=True, want_heights=True, with_sigma=False):
        nbv = self._arr("nbv", np.int32, self._np)
        n = int((nbv > 0).sum()) * self.cfg.sz * self.cfg.sz
        cloud = np.zeros((n, 32), dtype=np.uint8) if want_cloud else None
        heights = np.zeros(n, dtype=np.float64) if want_heights else None
        m = lib().orc_decode(self.h, _p(cloud) if want_cloud else None, _p(heights) if want_heights else None, int(with_sigma))
        assert m == n
        return cloud, heights

    def evaluate(self, off, x1, x2, y, conf=False):
        """Batched predict (sigma or conf) / likelihood / likelihood_dx over the fitted patches; needs fit(dump=True)."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
        m = x1.size
        f, sg, lk, dX = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(3 * m)
        rc = lib().orc_evaluate(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(y), int(conf), _p(f), _p(sg), _p(lk), _p(dX))
        assert rc == 0, "orc_evaluate needs the state dump (fit with dump=True)"
        return dict(f=f, sigma=sg, lik=lk, dX=dX.reshape(m, 3))

    def evaluate_rgb(self, off, x1, x2, Y, conf=False):
        """The same for the RGB field GPs of the previous fit_patches(colours=..., dump=True): Y is m x 3."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2))
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, 3)
        m = x1.size
        f, sg, lk, dX = np.zeros(3 * m), np.zeros(m), np.zeros(m), np.zeros(3 * m)
        rc = lib().orc_evaluate_rgb(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(Y), int(conf), _p(f), _p(sg), _p(lk), _p(dX))
        assert rc == 0, "orc_evaluate_rgb needs the field GPs' state dump (fit with colours and dump=True)"
        return dict(f=f.reshape(m, 3), sigma=sg, lik=lk, dX=dX.reshape(m, 3))

    def predict(self, patch, X, sigma=False):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 2)
        f = np.zeros(X.shape[0])
        sg = np.zeros(X.shape[0]) if sigma else None
        rc = lib().orc_predict(self.h, patch, _p(X), X.shape[0], _p(f), _p(sg) if sigma else None)
        assert rc == 0
        return (f, sg) if sigma else f
