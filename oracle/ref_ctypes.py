"""ctypes bindings to the two CPU checkers (ORACLE / TEST INFRASTRUCTURE ONLY):

  RefTree    -> oracle/_ref/libikd_ref.so   the UNMODIFIED reference behind oracle/ref_harness.cpp ("ref_" symbols)
  OracleTree -> oracle/libikd_oracle.so     the plain-C restatement oracle/ikd_oracle.c            ("ikdo_" symbols)

Both expose the same methods. May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libikd_ref.so")
ORACLE_SO = os.path.join(_HERE, "libikd_oracle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


def available():
    """True when the compiled reference (oracle/_ref) is present."""
    return os.path.exists(REF_SO)


def oracle_available():
    return os.path.exists(ORACLE_SO)


_SIGS = {
    "create": (_vp, [C.c_float, C.c_float, C.c_float]),
    "destroy": (None, [_vp]),
    "set_params": (None, [_vp, C.c_float, C.c_float, C.c_float]),
    "build": (None, [_vp, _f32p, C.c_long]),
    "knn": (C.c_int, [_vp, _f32p, C.c_int, C.c_double, _f32p, _f32p]),
    "knn_batch": (C.c_int, [_vp, _f32p, C.c_long, C.c_int, C.c_double, _vp, _vp, _vp, C.c_int]),
    "box_search": (C.c_long, [_vp, _f32p, _f32p, C.c_long]),
    "radius_search": (C.c_long, [_vp, _f32p, C.c_float, _f32p, C.c_long]),
    "add_points": (C.c_int, [_vp, _f32p, C.c_long, C.c_int]),
    "delete_points": (None, [_vp, _f32p, C.c_long]),
    "delete_boxes": (C.c_int, [_vp, _f32p, C.c_long]),
    "add_boxes": (None, [_vp, _f32p, C.c_long]),
    "size": (C.c_int, [_vp]),
    "validnum": (C.c_int, [_vp]),
    "root_alpha": (None, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tree_range": (None, [_vp, _f32p]),
    "flatten": (C.c_long, [_vp, _f32p, C.c_long]),
    "acquire_removed": (C.c_long, [_vp, _f32p, C.c_long]),
    "dump_tree": (C.c_long, [_vp, _f32p, C.c_long]),
    "max_depth": (C.c_int, [_vp]),
    "mean_visits": (C.c_double, [_vp, _f32p, C.c_long, C.c_int, C.c_double]),
    "num_threads": (C.c_int, []),
}

_libs = {}


def _load(path, prefix):
    key = (path, prefix)
    if key not in _libs:
        L = C.CDLL(path)
        fns = {}
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, prefix + name)
            fn.restype = res
            fn.argtypes = args
            fns[name] = fn
        # optional / differing symbols
        if prefix == "ref_":
            fns["last_result"] = L.ref_last_result
            fns["last_result"].restype = C.c_long
            fns["last_result"].argtypes = [_f32p, C.c_long]
            fns["wait_rebuild"] = L.ref_wait_rebuild
            fns["wait_rebuild"].argtypes = [_vp]
            for name in ("add_points", "delete_points", "delete_boxes", "add_boxes"):
                fn = getattr(L, "ref_" + name + "_serial", None)  # absent from a libikd_ref.so built before they existed
                if fn is not None:
                    fn.restype, fn.argtypes = _SIGS[name]
                    fns[name + "_serial"] = fn
        else:
            fns["last_result_h"] = L.ikdo_last_result
            fns["last_result_h"].restype = C.c_long
            fns["last_result_h"].argtypes = [_vp, _f32p, C.c_long]
            fns["rebuild_count"] = L.ikdo_rebuild_count
            fns["rebuild_count"].restype = C.c_int
            fns["rebuild_count"].argtypes = [_vp]
        _libs[key] = fns
    return _libs[key]


def lib():
    return _load(REF_SO, "ref_")


def plane_batch(q, nbr, sqd, cnt, max_kth_sqdist=5.0, threshold=0.1):
    """oracle/ikd_oracle.c ikdo_plane_batch: gate + plane fit + residual for queries q[nq,3] whose neighbours are
    nbr[nq,k,3] (ascending distance), sqd[nq,k], cnt[nq]. Returns (plane[nq,4], resid[nq], valid[nq] uint8)."""
    L = C.CDLL(ORACLE_SO)
    q = _pts(q)
    nq = len(q)
    nbr = np.ascontiguousarray(nbr, dtype=np.float32)
    k = nbr.shape[1]
    sqd = np.ascontiguousarray(sqd, dtype=np.float32)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    plane = np.zeros((nq, 4), dtype=np.float32)
    resid = np.zeros(nq, dtype=np.float32)
    valid = np.zeros(nq, dtype=np.uint8)
    L.ikdo_plane_batch.restype = None
    L.ikdo_plane_batch.argtypes = [_vp, C.c_long, C.c_int, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp]
    L.ikdo_plane_batch(q.ctypes.data, nq, k, nbr.ctypes.data, sqd.ctypes.data, cnt.ctypes.data, max_kth_sqdist,
                       threshold, plane.ctypes.data, resid.ctypes.data, valid.ctypes.data)
    return plane, resid, valid


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 3)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


class _CpuTree:
    """One CPU k-d tree with the reference's public interface (ikd_Tree.h:225-249)."""

    _path = None
    _prefix = None

    def __init__(self, delete_param=0.5, balance_param=0.6, box_length=0.2):
        self.F = _load(self._path, self._prefix)
        self.h = self.F["create"](delete_param, balance_param, box_length)

    def close(self):
        if getattr(self, "h", None):
            self.F["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, pts):
        pts = _pts(pts)
        self.F["build"](self.h, pts, len(pts))

    def knn(self, q, k, max_dist=float("inf"), nthreads=1, want_points=True):
        """Returns (xyz[nq,k,3] or None, sqdist[nq,k] (inf padded), count[nq])."""
        q = _pts(q)
        nq = len(q)
        d = np.full((nq, k), np.inf, dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        xyz = np.full((nq, k, 3), np.nan, dtype=np.float32) if want_points else None
        self.F["knn_batch"](self.h, q, nq, k, float(max_dist), xyz.ctypes.data if want_points else None,
                            d.ctypes.data, cnt.ctypes.data, nthreads)
        return xyz, d, cnt

    def _collect(self, n, first, cap):
        if n <= cap:
            return first[:n].copy()
        out = np.empty((n, 3), dtype=np.float32)
        if "last_result" in self.F:
            self.F["last_result"](out, n)
        else:
            self.F["last_result_h"](self.h, out, n)
        return out

    def box_search(self, box6, cap=4096):
        box6 = np.ascontiguousarray(box6, dtype=np.float32).reshape(6)
        buf = np.empty((cap, 3), dtype=np.float32)
        n = self.F["box_search"](self.h, box6, buf, cap)
        return self._collect(n, buf, cap)

    def radius_search(self, c, r, cap=4096):
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(3)
        buf = np.empty((cap, 3), dtype=np.float32)
        n = self.F["radius_search"](self.h, c, np.float32(r), buf, cap)
        return self._collect(n, buf, cap)

    _serial = False  # RefTree: route updates through the one-element-per-call harness entry points

    def _upd(self, name):
        if not self._serial:
            return self.F[name]
        if name + "_serial" in self.F:
            return self.F[name + "_serial"]

        def one_by_one(h, arr, n, *rest):  # same thing from Python for an older harness build
            total = 0
            for i in range(n):
                r = self.F[name](h, np.ascontiguousarray(arr[i:i + 1]), 1, *rest)
                self.F["wait_rebuild"](h)
                total += r or 0
            return total
        return one_by_one

    def add_points(self, pts, downsample_on):
        pts = _pts(pts)
        return self._upd("add_points")(self.h, pts, len(pts), 1 if downsample_on else 0)

    def delete_points(self, pts):
        pts = _pts(pts)
        self._upd("delete_points")(self.h, pts, len(pts))

    def delete_boxes(self, boxes):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        return self._upd("delete_boxes")(self.h, boxes, len(boxes))

    def add_boxes(self, boxes):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        self._upd("add_boxes")(self.h, boxes, len(boxes))

    def size(self):
        return self.F["size"](self.h)

    def validnum(self):
        return self.F["validnum"](self.h)

    def root_alpha(self):
        b, d = C.c_float(), C.c_float()
        self.F["root_alpha"](self.h, C.byref(b), C.byref(d))
        return b.value, d.value

    def tree_range(self):
        out = np.zeros(6, dtype=np.float32)
        self.F["tree_range"](self.h, out)
        return out

    def wait_rebuild(self):
        if "wait_rebuild" in self.F:
            self.F["wait_rebuild"](self.h)

    def flatten(self):
        n = max(self.size(), 1)
        buf = np.empty((n, 3), dtype=np.float32)
        m = self.F["flatten"](self.h, buf, n)
        return self._collect(m, buf, n)

    def acquire_removed(self, cap=1 << 20):
        buf = np.empty((cap, 3), dtype=np.float32)
        m = self.F["acquire_removed"](self.h, buf, cap)
        return buf[:min(m, cap)].copy()

    def dump_tree(self):
        """Pre-order structure dump [n,16]; see ref_harness.cpp:ref_dump_tree for the columns."""
        n = max(self.size(), 1)
        buf = np.empty((n, 16), dtype=np.float32)
        m = self.F["dump_tree"](self.h, buf.reshape(-1), n)
        assert m <= n
        return buf[:m].copy()

    def max_depth(self):
        return self.F["max_depth"](self.h)

    def mean_visits(self, q, k, max_dist=float("inf")):
        q = _pts(q)
        return self.F["mean_visits"](self.h, q, len(q), k, float(max_dist))

    def num_threads(self):
        return self.F["num_threads"]()


class RefTree(_CpuTree):
    """The unmodified reference (oracle/_ref/libikd_ref.so).

    serial=True (the default, for parity checks): every update call is fed to the reference one element per public call
    with a wait for its background rebuild thread in between -- same result as the batch call whenever no rebuild is in
    flight, and reproducible, which the batch call is not (see ref_harness.cpp). serial=False (timing legs of bench.py):
    the plain batch calls, i.e. the reference exactly as a caller runs it."""
    _path = REF_SO
    _prefix = "ref_"

    def __init__(self, delete_param=0.5, balance_param=0.6, box_length=0.2, serial=True):
        super().__init__(delete_param, balance_param, box_length)
        self._serial = bool(serial)


class OracleTree(_CpuTree):
    """The C restatement (oracle/libikd_oracle.so)."""
    _path = ORACLE_SO
    _prefix = "ikdo_"

    def rebuild_count(self):
        return self.F["rebuild_count"](self.h)
