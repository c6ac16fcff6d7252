"""ctypes binding to libikd_b200.so (the C ABI in include/ikd_b200.h).

Plumbing for tests/ and bench.py only: numpy in, numpy out, every call goes through the C ABI.
There is no Python or CPU implementation behind it; if the library or a B200 is missing the calls
raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IKD_LIB_PATH") or os.path.join(_HERE, "libikd_b200.so")  # override: kernel A/B experiments

_vp = C.c_void_p
_i64 = C.c_int64


class IkdError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("node_slots_used", C.c_int64), ("node_slots_cap", C.c_int64), ("max_depth", C.c_int32),
                ("rebuilds_partial", C.c_int32), ("rebuilds_full", C.c_int32), ("rebuilds_async", C.c_int32),
                ("rebuilt_points", C.c_int64), ("last_knn_visits", C.c_int64),
                ("add_points_in", C.c_int64), ("add_points_inserted", C.c_int64), ("add_vox_visits", C.c_int64),
                ("add_descend_levels", C.c_int64),
                ("rebuild_inline_ms", C.c_double), ("rebuild_inline_n", C.c_int64),
                ("rebuild_async_ms", C.c_double), ("rebuild_async_n", C.c_int64),
                ("rebuild_full_ms", C.c_double), ("rebuild_full_n", C.c_int64), ("rebuild_max_ms", C.c_double)]


class ReplicaDesc(C.Structure):
    _fields_ = [("header_dev", _vp), ("header_bytes", _i64), ("search_dev", _vp), ("search_bytes", _i64),
                ("update_dev", _vp), ("update_bytes", _i64), ("walk_dev", _vp), ("walk_bytes", _i64),
                ("points_dev", _vp), ("points_bytes", _i64),
                ("slots", _i64), ("npoints", _i64)]


# name -> (restype, argtypes); every symbol include/ikd_b200.h declares
SIGNATURES = {
    "ikd_last_error": (C.c_char_p, []),
    "ikd_abi_version": (C.c_int, []),
    "ikd_launch_count": (C.c_longlong, []),
    "ikd_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_float, C.c_float, C.c_float]),
    "ikd_destroy": (C.c_int, [_vp]),
    "ikd_set_delete_param": (C.c_int, [_vp, C.c_float]),
    "ikd_set_balance_param": (C.c_int, [_vp, C.c_float]),
    "ikd_set_downsample_param": (C.c_int, [_vp, C.c_float]),
    "ikd_size": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "ikd_validnum": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "ikd_root_alpha": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ikd_tree_range": (C.c_int, [_vp, _vp]),
    "ikd_has_root": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "ikd_build": (C.c_int, [_vp, _vp, _i64, _i64]),
    "ikd_knn_batch": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, C.c_double, _vp, _vp, _vp]),
    "ikd_knn_batch_dev": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_double, _vp, _vp, _vp]),
    "ikd_knn_plane_batch": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, C.c_double, C.c_float, C.c_float, _vp, _vp, _vp,
                                      _vp]),
    "ikd_knn_plane_batch_dev": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_double, C.c_float, C.c_float, _vp, _vp, _vp]),
    "ikd_box_search_batch": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ikd_radius_search_batch": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "ikd_search_fetch": (C.c_int, [_vp, _vp, _i64]),
    "ikd_get_points": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ikd_add_points": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32),
                                 C.POINTER(_i64), _vp]),
    "ikd_add_points_dev": (C.c_int, [_vp, _vp, _i64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32),
                                     C.POINTER(_i64), _vp]),
    "ikd_delete_points": (C.c_int, [_vp, _vp, _i64, _i64]),
    "ikd_delete_boxes": (C.c_int, [_vp, _vp, _i64, C.POINTER(C.c_int)]),
    "ikd_delete_points_dev": (C.c_int, [_vp, _vp, _i64]),
    "ikd_delete_boxes_dev": (C.c_int, [_vp, _vp, _i64, C.POINTER(C.c_int)]),
    "ikd_add_boxes_dev": (C.c_int, [_vp, _vp, _i64]),
    "ikd_next_id": (C.c_int, [_vp, C.POINTER(_i64)]),
    "ikd_id_epoch": (C.c_int, [_vp, C.POINTER(_i64)]),
    "ikd_compact_ids": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64), _vp, _i64, C.POINTER(_i64)]),
    "ikd_set_rebuild_timing": (C.c_int, [_vp, C.c_int]),
    "ikd_add_boxes": (C.c_int, [_vp, _vp, _i64]),
    "ikd_flatten": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "ikd_acquire_removed": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "ikd_synchronize": (C.c_int, [_vp]),
    "ikd_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "ikd_set_visit_counting": (C.c_int, [_vp, C.c_int]),
    "ikd_set_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "ikd_get_kernel_time": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "ikd_dump_tree": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "ikd_replica_export": (C.c_int, [_vp, C.POINTER(ReplicaDesc)]),
    "ikd_replica_prepare": (C.c_int, [_vp, _i64, _i64, C.POINTER(ReplicaDesc)]),
    "ikd_replica_commit": (C.c_int, [_vp]),
    "ikd_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
}

_lib = None


def load():
    """Load libikd_b200.so and attach signatures. Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IkdError(f"{LIB_PATH} not built: run `make -C ikd-tree_b200` (no fallback exists)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _chk(L, status):
    if status != 0:
        raise IkdError(f"ikd status {status}: {L.ikd_last_error().decode()}")


def _f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, cols)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


def launch_count():
    return load().ikd_launch_count()


class Tree:
    """One ikd_tree handle (one replica on one GPU)."""

    def __init__(self, delete_param=0.5, balance_param=0.6, box_length=0.2, device=-1):
        self.L = load()
        h = _vp()
        _chk(self.L, self.L.ikd_create(C.byref(h), device, delete_param, balance_param, box_length))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.ikd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters / counters
    def set_params(self, delete_param=None, balance_param=None, box_length=None):
        if delete_param is not None:
            _chk(self.L, self.L.ikd_set_delete_param(self.h, delete_param))
        if balance_param is not None:
            _chk(self.L, self.L.ikd_set_balance_param(self.h, balance_param))
        if box_length is not None:
            _chk(self.L, self.L.ikd_set_downsample_param(self.h, box_length))

    def size(self):
        v = C.c_int()
        _chk(self.L, self.L.ikd_size(self.h, C.byref(v)))
        return v.value

    def validnum(self):
        v = C.c_int()
        _chk(self.L, self.L.ikd_validnum(self.h, C.byref(v)))
        return v.value

    def has_root(self):
        v = C.c_int()
        _chk(self.L, self.L.ikd_has_root(self.h, C.byref(v)))
        return bool(v.value)

    def root_alpha(self):
        b, d = C.c_float(), C.c_float()
        _chk(self.L, self.L.ikd_root_alpha(self.h, C.byref(b), C.byref(d)))
        return b.value, d.value

    def tree_range(self):
        out = np.zeros(6, dtype=np.float32)
        _chk(self.L, self.L.ikd_tree_range(self.h, out.ctypes.data))
        return out

    # -- build / queries
    def build(self, pts):
        pts = _f32(pts, 3)
        _chk(self.L, self.L.ikd_build(self.h, pts.ctypes.data, len(pts), 12))

    def knn(self, q, k, max_dist=float("inf")):
        """Returns (idx[nq,k] int32 (-1 padded), sqdist[nq,k] (inf padded), count[nq])."""
        q = _f32(q, 3)
        nq = len(q)
        idx = np.empty((nq, k), dtype=np.int32)
        d = np.empty((nq, k), dtype=np.float32)
        cnt = np.empty(nq, dtype=np.int32)
        _chk(self.L, self.L.ikd_knn_batch(self.h, q.ctypes.data, nq, 12, k, float(max_dist), idx.ctypes.data,
                                          d.ctypes.data, cnt.ctypes.data))
        return idx, d, cnt

    def knn_dev(self, q_ptr, nq, k, max_dist, idx_ptr, d_ptr, cnt_ptr):
        """Device-pointer variant (float4 queries); asynchronous on the tree's stream."""
        _chk(self.L, self.L.ikd_knn_batch_dev(self.h, q_ptr, nq, k, float(max_dist), idx_ptr, d_ptr, cnt_ptr))

    def knn_plane(self, q, k=5, max_dist=float("inf"), max_kth_sqdist=5.0, threshold=0.1, want_idx=False):
        """kNN + plane fit on the device (ikd_knn_plane_batch). Returns (plane[nq,4], resid[nq], valid[nq] uint8)
        and, with want_idx, the neighbour ids [nq,k]."""
        q = _f32(q, 3)
        nq = len(q)
        plane = np.empty((nq, 4), dtype=np.float32)
        resid = np.empty(nq, dtype=np.float32)
        valid = np.empty(nq, dtype=np.uint8)
        idx = np.empty((nq, k), dtype=np.int32) if want_idx else None
        _chk(self.L, self.L.ikd_knn_plane_batch(self.h, q.ctypes.data, nq, 12, k, float(max_dist), max_kth_sqdist,
                                                threshold, plane.ctypes.data, resid.ctypes.data, valid.ctypes.data,
                                                idx.ctypes.data if want_idx else None))
        return (plane, resid, valid, idx) if want_idx else (plane, resid, valid)

    def knn_plane_dev(self, q_ptr, nq, k, max_dist, max_kth_sqdist, threshold, plane_ptr, resid_ptr, valid_ptr):
        """Device-pointer variant; asynchronous on the tree's stream."""
        _chk(self.L, self.L.ikd_knn_plane_batch_dev(self.h, q_ptr, nq, k, float(max_dist), max_kth_sqdist, threshold,
                                                    plane_ptr, resid_ptr, valid_ptr))

    def _fetch(self, offsets):
        total = int(offsets[-1])
        ids = np.empty(total, dtype=np.int32)
        if total:
            _chk(self.L, self.L.ikd_search_fetch(self.h, ids.ctypes.data, total))
        return ids

    def box_search(self, boxes):
        """boxes [nb,6] -> (offsets[nb+1] int64, ids[total] int32)."""
        boxes = _f32(boxes, 6)
        off = np.zeros(len(boxes) + 1, dtype=np.int64)
        _chk(self.L, self.L.ikd_box_search_batch(self.h, boxes.ctypes.data, len(boxes), off.ctypes.data))
        return off, self._fetch(off)

    def radius_search(self, centers, radii):
        centers = _f32(centers, 3)
        radii = np.ascontiguousarray(radii, dtype=np.float32).reshape(-1)
        assert len(radii) == len(centers)
        off = np.zeros(len(centers) + 1, dtype=np.int64)
        _chk(self.L, self.L.ikd_radius_search_batch(self.h, centers.ctypes.data, radii.ctypes.data, len(centers),
                                                    off.ctypes.data))
        return off, self._fetch(off)

    def get_points(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty((ids.size, 3), dtype=np.float32)
        _chk(self.L, self.L.ikd_get_points(self.h, ids.ctypes.data, ids.size, out.ctypes.data))
        return out.reshape(ids.shape + (3,))

    # -- updates
    def add_points(self, pts, downsample_on):
        """Returns (added, first_id, src[ninserted])."""
        pts = _f32(pts, 3)
        added, first, nins = C.c_int(), C.c_int32(), _i64()
        src = np.empty(max(len(pts), 1), dtype=np.int32)
        _chk(self.L, self.L.ikd_add_points(self.h, pts.ctypes.data, len(pts), 12, 1 if downsample_on else 0,
                                           C.byref(added), C.byref(first), C.byref(nins), src.ctypes.data))
        return added.value, first.value, src[:nins.value].copy()

    def add_points_dev(self, pts_ptr, n, downsample_on):
        """Device-pointer variant (float4 points). Returns (added, first_id, ninserted)."""
        added, first, nins = C.c_int(), C.c_int32(), _i64()
        _chk(self.L, self.L.ikd_add_points_dev(self.h, pts_ptr, n, 1 if downsample_on else 0, C.byref(added),
                                               C.byref(first), C.byref(nins), None))
        return added.value, first.value, nins.value

    def delete_points(self, pts):
        pts = _f32(pts, 3)
        _chk(self.L, self.L.ikd_delete_points(self.h, pts.ctypes.data, len(pts), 12))

    def delete_boxes(self, boxes):
        boxes = _f32(boxes, 6)
        n = C.c_int()
        _chk(self.L, self.L.ikd_delete_boxes(self.h, boxes.ctypes.data, len(boxes), C.byref(n)))
        return n.value

    def add_boxes(self, boxes):
        boxes = _f32(boxes, 6)
        _chk(self.L, self.L.ikd_add_boxes(self.h, boxes.ctypes.data, len(boxes)))

    # -- device-resident update inputs (replica delta sync: the broadcast buffer is applied without a host bounce)
    def delete_points_dev(self, pts_ptr, n):
        _chk(self.L, self.L.ikd_delete_points_dev(self.h, pts_ptr, n))

    def delete_boxes_dev(self, boxes_ptr, nb):
        n = C.c_int()
        _chk(self.L, self.L.ikd_delete_boxes_dev(self.h, boxes_ptr, nb, C.byref(n)))
        return n.value

    def add_boxes_dev(self, boxes_ptr, nb):
        _chk(self.L, self.L.ikd_add_boxes_dev(self.h, boxes_ptr, nb))

    # -- point ids
    def next_id(self):
        v = _i64()
        _chk(self.L, self.L.ikd_next_id(self.h, C.byref(v)))
        return v.value

    def id_epoch(self):
        v = _i64()
        _chk(self.L, self.L.ikd_id_epoch(self.h, C.byref(v)))
        return v.value

    def compact_ids(self):
        """Renumber the valid points 0..M-1. Returns (old_of_new[M], removed_old_ids)."""
        nrem = _i64()
        _chk(self.L, self.L.ikd_acquire_removed(self.h, None, 0, C.byref(nrem)))
        valid, size = self.validnum(), self.size()
        old = np.empty(max(valid, 1), dtype=np.int32)
        rem = np.empty(nrem.value + max(size - valid, 0) + 16, dtype=np.int32)
        na, nr = _i64(), _i64()
        _chk(self.L, self.L.ikd_compact_ids(self.h, old.ctypes.data, old.size, C.byref(na), rem.ctypes.data, rem.size,
                                            C.byref(nr)))
        return old[:na.value].copy(), rem[:nr.value].copy()

    def flatten(self):
        n = _i64()
        _chk(self.L, self.L.ikd_flatten(self.h, None, 0, C.byref(n)))
        ids = np.empty(n.value, dtype=np.int32)
        if n.value:
            _chk(self.L, self.L.ikd_flatten(self.h, ids.ctypes.data, n.value, C.byref(n)))
        return ids

    def acquire_removed(self):
        n = _i64()
        _chk(self.L, self.L.ikd_acquire_removed(self.h, None, 0, C.byref(n)))
        ids = np.empty(n.value, dtype=np.int32)
        if n.value:
            _chk(self.L, self.L.ikd_acquire_removed(self.h, ids.ctypes.data, n.value, C.byref(n)))
        return ids

    # -- misc
    def synchronize(self):
        _chk(self.L, self.L.ikd_synchronize(self.h))

    def stats(self):
        s = Stats()
        _chk(self.L, self.L.ikd_get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def set_kernel_timing(self, on):
        _chk(self.L, self.L.ikd_set_kernel_timing(self.h, 1 if on else 0))

    def kernel_time(self):
        """(accumulated kNN traversal-kernel milliseconds, launches) since the last call."""
        ms, n = C.c_double(), _i64()
        _chk(self.L, self.L.ikd_get_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_visit_counting(self, on):
        _chk(self.L, self.L.ikd_set_visit_counting(self.h, 1 if on else 0))

    def set_rebuild_timing(self, on):
        _chk(self.L, self.L.ikd_set_rebuild_timing(self.h, 1 if on else 0))

    def dump_tree(self):
        n = _i64()
        cap = max(self.size(), 1)
        buf = np.empty((cap, 16), dtype=np.float32)
        _chk(self.L, self.L.ikd_dump_tree(self.h, buf.ctypes.data, cap, C.byref(n)))
        assert n.value <= cap
        return buf[:n.value].copy()

    def stream(self):
        s = _vp()
        _chk(self.L, self.L.ikd_stream(self.h, C.byref(s)))
        return s.value

    def replica_export(self):
        d = ReplicaDesc()
        _chk(self.L, self.L.ikd_replica_export(self.h, C.byref(d)))
        return d

    def replica_prepare(self, slots, npoints):
        d = ReplicaDesc()
        _chk(self.L, self.L.ikd_replica_prepare(self.h, slots, npoints, C.byref(d)))
        return d

    def replica_commit(self):
        _chk(self.L, self.L.ikd_replica_commit(self.h))
