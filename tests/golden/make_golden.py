"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/libikd_ref.so).

Run in the dev container (where /root/reference exists and `make -C oracle ref` has been run):
    python tests/golden/make_golden.py
Writes tests/golden/ikd_golden_v1.npz (small, committed). The reference ships no golden vectors of its
own, so these outputs of the reference itself are what pins the oracle and the
CUDA path on machines where the reference sources are absent.

Scenario (seeded, sizes kept small so the file stays < 1 MB): Build on 4,000 uniform points, 5-NN /
1-NN / 12-NN with and without max_dist, box and radius searches, Delete_Point_Boxes, Delete_Points,
Add_Points without and with downsample, Add_Point_Boxes, then the same queries again. All inputs use
distinct coordinates and sizes below the reference's background-rebuild threshold interplay, so the
outputs are reproducible run to run.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_ctypes as R  # noqa: E402


def rows(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


def scenario(tree_cls):
    rng = np.random.default_rng(20240611)
    out = {}
    P = (rng.random((4000, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    Q = (rng.random((300, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    ctr = (rng.random((40, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    half = (rng.random(40, dtype=np.float32) * 1.2 + 0.2).astype(np.float32)
    boxes = np.concatenate([ctr - half[:, None], ctr + half[:, None]], axis=1).astype(np.float32)
    del_boxes = boxes[:3].copy()
    del_pts = P[rng.choice(len(P), 150, replace=False)].copy()
    add1 = (rng.random((600, 3), dtype=np.float32) * 10 - 5).astype(np.float32)
    add2 = (rng.random((1500, 3), dtype=np.float32) * 11 - 5.5).astype(np.float32)
    out.update(P=P, Q=Q, ctr=ctr, half=half, boxes=boxes, del_boxes=del_boxes, del_pts=del_pts, add1=add1, add2=add2)
    params = (0.5, 0.6, 0.4)
    out["params"] = np.asarray(params, dtype=np.float32)
    t = tree_cls(*params)
    t.build(P)
    out["build_dump"] = t.dump_tree()
    out["build_range"] = t.tree_range()

    def queries(tag):
        for k, md in ((5, np.inf), (1, np.inf), (12, np.inf), (5, 0.5)):
            _, d, c = t.knn(Q, k, md, want_points=False)
            out[f"{tag}_knn_k{k}_md{md}_d"] = d
            out[f"{tag}_knn_k{k}_md{md}_c"] = c
        bs, rs, bo, ro = [], [], [0], [0]
        for i in range(len(boxes)):
            b = rows(t.box_search(boxes[i], cap=8192))
            r = rows(t.radius_search(ctr[i], half[i], cap=8192))
            bs.append(b); rs.append(r); bo.append(bo[-1] + len(b)); ro.append(ro[-1] + len(r))
        out[f"{tag}_box_pts"] = np.concatenate(bs) if bs else np.zeros((0, 3), np.float32)
        out[f"{tag}_box_off"] = np.asarray(bo, dtype=np.int64)
        out[f"{tag}_rad_pts"] = np.concatenate(rs) if rs else np.zeros((0, 3), np.float32)
        out[f"{tag}_rad_off"] = np.asarray(ro, dtype=np.int64)

    queries("s0")
    out["del_boxes_count"] = np.int64(t.delete_boxes(del_boxes))
    t.delete_points(del_pts)
    out["s1_validnum"] = np.int64(t.validnum())
    queries("s1")
    out["add1_ret"] = np.int64(t.add_points(add1, False))
    out["add2_ret"] = np.int64(t.add_points(add2, True))
    t.wait_rebuild()
    out["s2_validnum"] = np.int64(t.validnum())
    out["s2_valid_set"] = rows(t.flatten())
    queries("s2")
    t.add_boxes(del_boxes)
    t.wait_rebuild()
    out["s3_validnum"] = np.int64(t.validnum())
    out["s3_valid_set"] = rows(t.flatten())
    t.close()
    return out


if __name__ == "__main__":
    if not R.available():
        raise SystemExit("oracle/_ref/libikd_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    a = scenario(R.RefTree)
    b = scenario(R.RefTree)  # reproducibility of the reference itself on this scenario
    for k in a:
        assert np.array_equal(a[k], b[k]), f"reference not reproducible on {k}"
    path = os.path.join(HERE, "ikd_golden_v1.npz")
    np.savez_compressed(path, **a)
    print("wrote", path, os.path.getsize(path), "bytes;", len(a), "arrays")
