"""ctypes binding of libgpc_b200.so (the C ABI declared in include/gpc.h).

This is the thin Python face used by the tests and bench.py; the product is the shared
library.  There is no fallback: if the library is missing or no B200 is usable the calls
raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpc_b200.so")
_LIB = None

GPC_OK = 0

# every symbol include/gpc.h declares (tests check the library exports all of them)
SYMBOLS = [
    "gpc_config_default", "gpc_create", "gpc_destroy", "gpc_last_error", "gpc_version", "gpc_compress",
    "gpc_upload_cloud", "gpc_compress_resident", "gpc_add_measurements", "gpc_compress_shard_begin", "gpc_compress_shard_finish", "gpc_fit_patches", "gpc_decompress", "gpc_decompress_resident",
    "gpc_get_heights", "gpc_predict", "gpc_evaluate_patches", "gpc_evaluate_patches_rgb", "gpc_get_sizes", "gpc_get_stats", "gpc_get_patches", "gpc_get_assignment",
    "gpc_get_params", "gpc_get_params_rgb", "gpc_get_state", "gpc_set_params", "gpc_set_rand_offset", "gpc_get_stream", "gpc_debug_exp", "gpc_debug_rand", "gpc_debug_peak", "gpc_shard_range", "gpc_save", "gpc_load", "gpc_get_config",
]


class GpcConfig(C.Structure):
    _fields_ = [("res", C.c_double), ("sz", C.c_int32), ("capacity", C.c_int32), ("s0", C.c_double),
                ("eps_tol", C.c_double), ("sigmaf_sq", C.c_double), ("l_sq", C.c_double), ("leaf_order", C.c_int32),
                ("shuffle", C.c_int32), ("rgb_rand", C.c_int32), ("device", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("keep_state", C.c_int32), ("rgb", C.c_int32), ("rgb_s0", C.c_double),
                ("rgb_eps_tol", C.c_double), ("decode_separable", C.c_int32), ("pad0", C.c_int32)]


class GpcSizes(C.Structure):
    _fields_ = [("n_in", C.c_int64), ("n_patches", C.c_int64), ("n_claimed", C.c_int64), ("n_bv_total", C.c_int64),
                ("patch_lo", C.c_int64), ("patch_hi", C.c_int64), ("n_decoded", C.c_int64), ("rand_offset", C.c_uint64),
                ("lattice_min", C.c_double * 3), ("depth", C.c_uint32), ("pad", C.c_uint32)]


_STAT_U64 = ["n_add", "n_first", "n_sparse", "n_full", "n_del_cap", "n_del_geo", "sum_n", "sum_n2_common",
             "sum_n2_sparse", "sum_n2_full", "sum_n2_del"]
_STAT_MS = ["ms_h2d", "ms_lattice", "ms_keys", "ms_sort", "ms_leaves", "ms_rotation", "ms_claim", "ms_group",
            "ms_shuffle", "ms_fit", "ms_d2h", "ms_predict", "ms_total", "ms_fit_rgb", "ms_evaluate", "pad_"]
_STAT_RGB = ["rgb_n_sparse", "rgb_n_full", "rgb_n_del_cap", "rgb_n_del_geo", "rgb_sum_n2_common"]


class GpcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in _STAT_U64] + [("escalated", C.c_uint64 * 4), ("kernel_launches", C.c_uint64)] + \
               [(n, C.c_uint64) for n in _STAT_RGB] + [(n, C.c_float) for n in _STAT_MS] + \
               [("max_bv", C.c_uint64), ("bv_hist", C.c_uint64 * 33)]


def load():
    """Loads the shared library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m gp_compressor_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.gpc_version.restype = C.c_char_p
    L.gpc_last_error.restype = C.c_char_p
    L.gpc_last_error.argtypes = [vp]
    L.gpc_config_default.argtypes = [C.POINTER(GpcConfig)]
    L.gpc_create.argtypes = [C.POINTER(GpcConfig), C.POINTER(vp)]
    L.gpc_destroy.argtypes = [vp]
    L.gpc_destroy.restype = None
    L.gpc_compress.argtypes = [vp, vp, i64]
    L.gpc_upload_cloud.argtypes = [vp, vp, i64]
    L.gpc_compress_resident.argtypes = [vp]
    L.gpc_fit_patches.argtypes = [vp, i64, vp, vp, vp, vp]
    L.gpc_decompress.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.gpc_decompress_resident.argtypes = [vp, C.POINTER(i64)]
    L.gpc_get_heights.argtypes = [vp, vp, i64]
    L.gpc_predict.argtypes = [vp, i64, vp, i64, vp, vp]
    L.gpc_add_measurements.argtypes = [vp, i64, vp, vp, vp, vp]
    L.gpc_compress_shard_begin.argtypes = [vp, vp, i64, vp, vp]
    L.gpc_compress_shard_finish.argtypes = [vp, i64, C.c_uint64, i64, C.c_uint64]
    L.gpc_evaluate_patches.argtypes = [vp, i64, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.gpc_evaluate_patches_rgb.argtypes = [vp, i64, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.gpc_get_sizes.argtypes = [vp, C.POINTER(GpcSizes)]
    L.gpc_get_stats.argtypes = [vp, C.POINTER(GpcStats)]
    L.gpc_get_patches.argtypes = [vp] + [vp] * 8
    L.gpc_get_assignment.argtypes = [vp] + [vp] * 6
    L.gpc_get_params.argtypes = [vp] + [vp] * 7
    L.gpc_get_state.argtypes = [vp, i64, vp, vp]
    L.gpc_get_params_rgb.argtypes = [vp] + [vp] * 7 + [C.POINTER(i64)]
    L.gpc_set_params.argtypes = [vp, i64] + [vp] * 7
    L.gpc_set_rand_offset.argtypes = [vp, u64]
    L.gpc_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.gpc_debug_exp.argtypes = [vp, vp, vp, i64]
    L.gpc_debug_rand.argtypes = [vp, u64, i64, vp]
    L.gpc_debug_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.gpc_save.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    L.gpc_load.argtypes = [vp, C.c_char_p]
    L.gpc_get_config.argtypes = [vp, C.POINTER(GpcConfig)]
    L.gpc_shard_range.argtypes = [vp, i64, C.c_int32, C.c_int32, C.POINTER(i64), C.POINTER(i64)]
    _LIB = L
    return L


def default_config():
    cfg = GpcConfig()
    load().gpc_config_default(C.byref(cfg))
    return cfg


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def shard_prefix(all_counts, rank):
    """The arguments of gpc_compress_shard_finish from the all-gathered (owned_patches, owned_draws) pairs of every rank
    (rank order = visiting order): (patches_before, draws_before, patches_total, draws_total).  Host arithmetic only."""
    a = np.asarray(all_counts, dtype=np.int64).reshape(-1, 2)
    return int(a[:rank, 0].sum()), int(a[:rank, 1].sum()), int(a[:, 0].sum()), int(a[:, 1].sum())


def shard_range(off, rank, count):
    """Patches [lo, hi) owned by shard `rank` of `count` (host arithmetic only, no GPU needed)."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    lo, hi = C.c_int64(0), C.c_int64(0)
    rc = load().gpc_shard_range(_p(off), off.size - 1, rank, count, C.byref(lo), C.byref(hi))
    if rc != GPC_OK:
        raise GpcError(f"gpc_shard_range failed with code {rc}")
    return lo.value, hi.value


class GpcError(RuntimeError):
    pass


class Handle:
    """RAII wrapper of a gpc_handle."""

    def __init__(self, **kw):
        L = load()
        cfg = default_config()
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown gpc_config field {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = L.gpc_create(C.byref(cfg), C.byref(self.h))
        if rc != GPC_OK:
            self.h = None
            raise GpcError(f"gpc_create failed with code {rc} (2 = no usable sm_100 CUDA device; there is no CPU fallback)")

    def close(self):
        if getattr(self, "h", None):
            load().gpc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != GPC_OK:
            raise GpcError(f"gpc error {rc}: {load().gpc_last_error(self.h).decode()}")

    # ---- compress -------------------------------------------------------------------
    def compress(self, cloud32):
        assert cloud32.dtype == np.uint8 and cloud32.ndim == 2 and cloud32.shape[1] == 32 and cloud32.flags.c_contiguous
        self._ck(load().gpc_compress(self.h, _p(cloud32), cloud32.shape[0]))

    def compress_shard_begin(self, cloud32=None):
        """First phase of the sharded-binning compress; returns (owned_patches, owned_draws) for the all-gather."""
        op, od = C.c_int64(0), C.c_uint64(0)
        if cloud32 is None:
            rc = load().gpc_compress_shard_begin(self.h, None, 0, C.byref(op), C.byref(od))
        else:
            assert cloud32.dtype == np.uint8 and cloud32.ndim == 2 and cloud32.shape[1] == 32 and cloud32.flags.c_contiguous
            rc = load().gpc_compress_shard_begin(self.h, _p(cloud32), cloud32.shape[0], C.byref(op), C.byref(od))
        self._ck(rc)
        return int(op.value), int(od.value)

    def compress_shard_finish(self, patches_before, draws_before, patches_total, draws_total):
        self._ck(load().gpc_compress_shard_finish(self.h, patches_before, draws_before, patches_total, draws_total))

    def upload_cloud(self, cloud32):
        assert cloud32.dtype == np.uint8 and cloud32.ndim == 2 and cloud32.shape[1] == 32 and cloud32.flags.c_contiguous
        self._ck(load().gpc_upload_cloud(self.h, _p(cloud32), cloud32.shape[0]))

    def upload_cloud_ptr(self, ptr, n):
        self._ck(load().gpc_upload_cloud(self.h, C.c_void_p(ptr), n))

    def compress_ptr(self, ptr, n):
        self._ck(load().gpc_compress(self.h, C.c_void_p(ptr), n))

    def compress_resident(self):
        self._ck(load().gpc_compress_resident(self.h))

    def fit_patches(self, off, x1, x2, y):
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
        self._ck(load().gpc_fit_patches(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(y)))

    def add_measurements(self, off, x1, x2, y):
        """gpc_add_measurements: more points for the patches of the previous fit (keep_state=1), state continued."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2, y))
        self._ck(load().gpc_add_measurements(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(y)))

    # ---- decompress -----------------------------------------------------------------
    def decompress(self, out=None):
        n = C.c_int64(0)
        if out is None:
            s = self.sizes()
            # upper bound: every patch of the shard non-empty
            out = np.zeros(((s.patch_hi - s.patch_lo) * self.cfg.sz * self.cfg.sz, 32), dtype=np.uint8)
        self._ck(load().gpc_decompress(self.h, _p(out), out.shape[0], C.byref(n)))
        return out[: n.value]

    def decompress_ptr(self, ptr, capacity):
        n = C.c_int64(0)
        self._ck(load().gpc_decompress(self.h, C.c_void_p(ptr), capacity, C.byref(n)))
        return n.value

    def decompress_resident(self):
        n = C.c_int64(0)
        self._ck(load().gpc_decompress_resident(self.h, C.byref(n)))
        return n.value

    def heights(self):
        n = self.sizes().n_decoded
        out = np.zeros(n, dtype=np.float64)
        self._ck(load().gpc_get_heights(self.h, _p(out), n))
        return out

    def predict(self, patch, X, sigma=False):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 2)
        f = np.zeros(X.shape[0])
        sg = np.zeros(X.shape[0]) if sigma else None
        self._ck(load().gpc_predict(self.h, patch, _p(X), X.shape[0], _p(f), _p(sg)))
        return (f, sg) if sigma else f

    def evaluate(self, off, x1, x2, y=None, conf=False, want=("f", "sigma", "lik", "dX")):
        """gpc_evaluate_patches: batched predict (sigma / conf), likelihood and likelihood gradient over the first
        len(off) - 1 patches of this shard; needs keep_state=1."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2))
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        m = x1.size
        out = dict(f=np.zeros(m) if "f" in want else None, sigma=np.zeros(m) if "sigma" in want else None,
                   lik=np.zeros(m) if "lik" in want and yy is not None else None,
                   dX=np.zeros((m, 3)) if "dX" in want and yy is not None else None)
        self._ck(load().gpc_evaluate_patches(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(yy), int(conf), _p(out["f"]),
                                             _p(out["sigma"]), _p(out["lik"]), _p(out["dX"])))
        return {k: v for k, v in out.items() if v is not None}

    def evaluate_rgb(self, off, x1, x2, Y, conf=False):
        """gpc_evaluate_patches_rgb: the RGB field GPs (rgb=1, keep_state=1); Y is m x 3 (colours centred on the patch mean)."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        x1, x2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x1, x2))
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, 3)
        m = x1.size
        out = dict(f=np.zeros((m, 3)), sigma=np.zeros(m), lik=np.zeros(m), dX=np.zeros((m, 3)))
        self._ck(load().gpc_evaluate_patches_rgb(self.h, off.size - 1, _p(off), _p(x1), _p(x2), _p(Y), int(conf), _p(out["f"]),
                                                 _p(out["sigma"]), _p(out["lik"]), _p(out["dX"])))
        return out

    # ---- results ----------------------------------------------------------------------
    def sizes(self):
        s = GpcSizes()
        self._ck(load().gpc_get_sizes(self.h, C.byref(s)))
        return s

    def stats(self):
        s = GpcStats()
        self._ck(load().gpc_get_stats(self.h, C.byref(s)))
        d = {n: getattr(s, n) for n in _STAT_U64 + _STAT_MS + _STAT_RGB + ["kernel_launches"]}
        d["escalated"] = list(s.escalated)
        d["max_bv"] = int(s.max_bv)
        d["bv_hist"] = list(s.bv_hist)   # [k] = patches with k basis vectors (k < 32), [32] = 32 or more
        return d

    def params(self):
        s = self.sizes()
        PL = s.patch_hi - s.patch_lo
        T = s.n_bv_total
        r = dict(nbv=np.zeros(PL, np.int32), bv_off=np.zeros(PL + 1, np.int64), bv_idx=np.zeros(T, np.int32),
                 bv1=np.zeros(T), bv2=np.zeros(T), alpha=np.zeros(T), flags=np.zeros(PL, np.int32))
        self._ck(load().gpc_get_params(self.h, _p(r["nbv"]), _p(r["bv_off"]), _p(r["bv_idx"]), _p(r["bv1"]),
                                       _p(r["bv2"]), _p(r["alpha"]), _p(r["flags"])))
        r["patch_lo"], r["patch_hi"] = s.patch_lo, s.patch_hi
        return r

    def params_rgb(self):
        """RGB field GP parameters of this shard (gpc_config.rgb = 1)."""
        s = self.sizes()
        PL = s.patch_hi - s.patch_lo
        tot = C.c_int64(0)
        self._ck(load().gpc_get_params_rgb(self.h, None, None, None, None, None, None, None, C.byref(tot)))
        T = tot.value
        r = dict(nbv=np.zeros(PL, np.int32), bv_off=np.zeros(PL + 1, np.int64), bv_idx=np.zeros(T, np.int32), bv1=np.zeros(T), bv2=np.zeros(T),
                 alpha=np.zeros(3 * T), perm=np.zeros(s.n_claimed, np.int32))
        self._ck(load().gpc_get_params_rgb(self.h, _p(r["nbv"]), _p(r["bv_off"]), _p(r["bv_idx"]), _p(r["bv1"]), _p(r["bv2"]), _p(r["alpha"]),
                                           _p(r["perm"]), C.byref(tot)))
        return r

    def params_into(self, nbv_ptr, bv_off_ptr, idx_ptr, bv1_ptr, bv2_ptr, alpha_ptr, flags_ptr):
        """gpc_get_params into caller-owned (e.g. pinned) buffers given as raw addresses."""
        vp = lambda p: C.c_void_p(p) if p else None
        self._ck(load().gpc_get_params(self.h, vp(nbv_ptr), vp(bv_off_ptr), vp(idx_ptr), vp(bv1_ptr), vp(bv2_ptr), vp(alpha_ptr), vp(flags_ptr)))

    def state(self, patch, n):
        Cm = np.zeros((n, n))
        Qm = np.zeros((n, n))
        self._ck(load().gpc_get_state(self.h, patch, _p(Cm), _p(Qm)))
        return Cm, Qm

    def patches(self, frames=True, binning=True):
        s = self.sizes()
        P = s.n_patches
        r = dict(patch_off=np.zeros(P + 1, np.int64))
        if binning:
            r.update(leaf_code=np.zeros(P, np.uint64), leaf_center=np.zeros(P * 3, np.float32),
                     leaf_ncand=np.zeros(P, np.int32), leaf_R=np.zeros(P * 9))
        if frames:
            r.update(leaf_quat=np.zeros(P * 4), leaf_mean=np.zeros(P * 3), leaf_rgbmean=np.zeros(P * 3))
        self._ck(load().gpc_get_patches(self.h, _p(r.get("leaf_code")), _p(r.get("leaf_center")), _p(r.get("leaf_ncand")),
                                        _p(r.get("leaf_R")), _p(r.get("leaf_quat")), _p(r.get("leaf_mean")),
                                        _p(r.get("leaf_rgbmean")), _p(r["patch_off"])))
        r["depth"] = s.depth
        r["lattice_min"] = np.array(list(s.lattice_min))
        r["n_leaves"] = P
        r["n_claimed"] = s.n_claimed
        return r

    def assignment(self, binning=True):
        s = self.sizes()
        S = s.n_claimed
        r = dict(st_x1=np.zeros(S), st_x2=np.zeros(S), st_y=np.zeros(S), perm=np.zeros(S, np.int32))
        if binning:
            r.update(owner=np.zeros(s.n_in, np.int32), st_idx=np.zeros(S, np.int32))
        self._ck(load().gpc_get_assignment(self.h, _p(r.get("owner")), _p(r.get("st_idx")), _p(r["st_x1"]), _p(r["st_x2"]),
                                           _p(r["st_y"]), _p(r["perm"])))
        return r

    def set_params(self, nbv, bv1, bv2, alpha, quat=None, mean=None, rgbmean=None):
        nbv = np.ascontiguousarray(nbv, dtype=np.int32)
        bv1, bv2, alpha = (np.ascontiguousarray(a, dtype=np.float64) for a in (bv1, bv2, alpha))
        if quat is not None:
            quat, mean, rgbmean = (np.ascontiguousarray(a, dtype=np.float64) for a in (quat, mean, rgbmean))
        self._ck(load().gpc_set_params(self.h, nbv.size, _p(nbv), _p(bv1), _p(bv2), _p(alpha), _p(quat), _p(mean), _p(rgbmean)))

    def stream(self):
        s = C.c_void_p()
        self._ck(load().gpc_get_stream(self.h, C.byref(s)))
        return s.value or 0

    def set_rand_offset(self, off):
        self._ck(load().gpc_set_rand_offset(self.h, off))

    # ---- test hooks -------------------------------------------------------------------
    def debug_exp(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        self._ck(load().gpc_debug_exp(self.h, _p(x), _p(out), x.size))
        return out

    def save(self, path):
        n = C.c_int64(0)
        self._ck(load().gpc_save(self.h, str(path).encode(), C.byref(n)))
        return n.value

    def load_file(self, path):
        self._ck(load().gpc_load(self.h, str(path).encode()))
        self._ck(load().gpc_get_config(self.h, C.byref(self.cfg)))  # the file carries the decoder's configuration
        return self.sizes()

    def debug_peak(self, kind):
        v = C.c_double(0)
        self._ck(load().gpc_debug_peak(self.h, kind, C.byref(v)))
        return v.value

    def debug_rand(self, offset, n):
        out = np.zeros(n, dtype=np.uint32)
        self._ck(load().gpc_debug_rand(self.h, offset, n, _p(out)))
        return out
