"""ctypes binding to oracle/_ref/libikd_ref.so (the UNMODIFIED reference behind oracle/ref_harness.cpp).

ORACLE / TEST INFRASTRUCTURE ONLY. May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libikd_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_float, C.c_float, C.c_float]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.ref_build.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_knn.restype = C.c_int
        L.ref_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_double, _f32p, _f32p]
        L.ref_knn_batch.restype = C.c_int
        L.ref_knn_batch.argtypes = [C.c_void_p, _f32p, C.c_long, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int]
        L.ref_box_search.restype = C.c_long
        L.ref_box_search.argtypes = [C.c_void_p, _f32p, _f32p, C.c_long]
        L.ref_radius_search.restype = C.c_long
        L.ref_radius_search.argtypes = [C.c_void_p, _f32p, C.c_float, _f32p, C.c_long]
        L.ref_last_result.restype = C.c_long
        L.ref_last_result.argtypes = [_f32p, C.c_long]
        L.ref_add_points.restype = C.c_int
        L.ref_add_points.argtypes = [C.c_void_p, _f32p, C.c_long, C.c_int]
        L.ref_delete_points.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_delete_boxes.restype = C.c_int
        L.ref_delete_boxes.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_add_boxes.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_size.restype = C.c_int
        L.ref_size.argtypes = [C.c_void_p]
        L.ref_validnum.restype = C.c_int
        L.ref_validnum.argtypes = [C.c_void_p]
        L.ref_root_alpha.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_tree_range.argtypes = [C.c_void_p, _f32p]
        L.ref_wait_rebuild.argtypes = [C.c_void_p]
        L.ref_flatten.restype = C.c_long
        L.ref_flatten.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_acquire_removed.restype = C.c_long
        L.ref_acquire_removed.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_dump_tree.restype = C.c_long
        L.ref_dump_tree.argtypes = [C.c_void_p, _f32p, C.c_long]
        L.ref_max_depth.restype = C.c_int
        L.ref_max_depth.argtypes = [C.c_void_p]
        L.ref_mean_visits.restype = C.c_double
        L.ref_mean_visits.argtypes = [C.c_void_p, _f32p, C.c_long, C.c_int, C.c_double]
        L.ref_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


class RefTree:
    """Thin object wrapper: one reference KD_TREE<ikdTree_PointType> (ikd_Tree.h:225-249)."""

    def __init__(self, delete_param=0.5, balance_param=0.6, box_length=0.2):
        self.L = lib()
        self.h = self.L.ref_create(delete_param, balance_param, box_length)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, pts):
        pts = _pts(pts)
        self.L.ref_build(self.h, pts, len(pts))

    def knn(self, q, k, max_dist=float("inf"), nthreads=1, want_points=True):
        """Returns (xyz[nq,k,3] or None, sqdist[nq,k] (inf padded), count[nq])."""
        q = _pts(q)
        nq = len(q)
        d = np.full((nq, k), np.inf, dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        xyz = np.full((nq, k, 3), np.nan, dtype=np.float32) if want_points else None
        self.L.ref_knn_batch(self.h, q, nq, k, float(max_dist), xyz.ctypes.data if want_points else None,
                             d.ctypes.data, cnt.ctypes.data, nthreads)
        return xyz, d, cnt

    def _collect(self, n, first, cap):
        if n <= cap:
            return first[:n].copy()
        out = np.empty((n, 3), dtype=np.float32)
        self.L.ref_last_result(out, n)
        return out

    def box_search(self, box6, cap=4096):
        box6 = np.ascontiguousarray(box6, dtype=np.float32).reshape(6)
        buf = np.empty((cap, 3), dtype=np.float32)
        n = self.L.ref_box_search(self.h, box6, buf, cap)
        return self._collect(n, buf, cap)

    def radius_search(self, c, r, cap=4096):
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(3)
        buf = np.empty((cap, 3), dtype=np.float32)
        n = self.L.ref_radius_search(self.h, c, np.float32(r), buf, cap)
        return self._collect(n, buf, cap)

    def add_points(self, pts, downsample_on):
        pts = _pts(pts)
        return self.L.ref_add_points(self.h, pts, len(pts), 1 if downsample_on else 0)

    def delete_points(self, pts):
        pts = _pts(pts)
        self.L.ref_delete_points(self.h, pts, len(pts))

    def delete_boxes(self, boxes):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        return self.L.ref_delete_boxes(self.h, boxes, len(boxes))

    def add_boxes(self, boxes):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        self.L.ref_add_boxes(self.h, boxes, len(boxes))

    def size(self):
        return self.L.ref_size(self.h)

    def validnum(self):
        return self.L.ref_validnum(self.h)

    def root_alpha(self):
        b, d = C.c_float(), C.c_float()
        self.L.ref_root_alpha(self.h, C.byref(b), C.byref(d))
        return b.value, d.value

    def tree_range(self):
        out = np.zeros(6, dtype=np.float32)
        self.L.ref_tree_range(self.h, out)
        return out

    def wait_rebuild(self):
        self.L.ref_wait_rebuild(self.h)

    def flatten(self):
        n = max(self.size(), 1)
        buf = np.empty((n, 3), dtype=np.float32)
        m = self.L.ref_flatten(self.h, buf, n)
        return self._collect(m, buf, n)

    def acquire_removed(self, cap=1 << 20):
        buf = np.empty((cap, 3), dtype=np.float32)
        m = self.L.ref_acquire_removed(self.h, buf, cap)
        return self._collect(m, buf, cap)

    def dump_tree(self):
        """Pre-order structure dump [n,16]; see ref_harness.cpp:ref_dump_tree for the columns."""
        n = max(self.size(), 1)
        buf = np.empty((n, 16), dtype=np.float32)
        m = self.L.ref_dump_tree(self.h, buf.reshape(-1), n)
        assert m <= n
        return buf[:m].copy()

    def max_depth(self):
        return self.L.ref_max_depth(self.h)

    def mean_visits(self, q, k, max_dist=float("inf")):
        q = _pts(q)
        return self.L.ref_mean_visits(self.h, q, len(q), k, float(max_dist))

    def num_threads(self):
        return self.L.ref_num_threads()
