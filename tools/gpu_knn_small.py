"""Small-batch kNN sweep: 1M-point map, nq in a list, timing of the whole ikd_knn_batch_dev call and of the
traversal kernel alone (L2 flushed between calls), plus a digest of the results so that runs with different
IKD_KNN_G (lanes per query; 0 = one thread per query) can be compared for equality."""
import hashlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ikd-tree_b200"))
import ikd_ctypes as I
k = int(os.environ.get("K", "5"))
n = 1_000_000
rng = np.random.default_rng(4)
P = (rng.random((n, 3), dtype=np.float32) * 100 - 50).astype(np.float32)
t = I.Tree(0.5, 0.6, 0.5); t.build(P)
# unbalance it a little: deletes + inserts
t.delete_boxes(np.array([[-10, -10, -10, 5, 5, 5]], np.float32))
A = (rng.random((50000, 3), dtype=np.float32) * 30 - 15).astype(np.float32)
t.add_points(A, False)
ts = torch.cuda.ExternalStream(t.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {"G": os.environ.get("IKD_KNN_G", "auto"), "k": k, "rows": []}
for nq in [int(x) for x in os.environ.get("NQ", "2000,5000,10000,20000,50000,100000,200000").split(",")]:
    g = torch.Generator(device="cuda").manual_seed(nq)
    q4 = torch.zeros((nq, 4), dtype=torch.float32, device="cuda")
    q4[:, :3] = torch.rand((nq, 3), generator=g, device="cuda") * 104 - 52
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda"); od = torch.empty((nq, k), dtype=torch.float32, device="cuda"); oc = torch.empty(nq, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t.set_kernel_timing(True)
    call_ms = []
    for it in range(8):
        flush.fill_(it); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ts):
            e0.record(ts)
            t.knn_dev(q4.data_ptr(), nq, k, 3.0 if it % 2 else float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr())
            e1.record(ts)
        t.synchronize(); torch.cuda.synchronize()
        if it >= 2: call_ms.append(e0.elapsed_time(e1))
        if it == 0: t.kernel_time()
    kms, kn = t.kernel_time()
    t.set_kernel_timing(False)
    t.set_visit_counting(True)
    t.knn_dev(q4.data_ptr(), nq, k, float("inf"), oi.data_ptr(), od.data_ptr(), oc.data_ptr()); t.synchronize()
    vis = t.stats()["last_knn_visits"] / nq
    t.set_visit_counting(False)
    torch.cuda.synchronize()
    h = hashlib.sha1(od.cpu().numpy().tobytes() + oi.cpu().numpy().tobytes() + oc.cpu().numpy().tobytes()).hexdigest()[:12]
    res["rows"].append({"nq": nq, "call_us": round(1e3 * float(np.median(call_ms)), 1), "kernel_us": round(1e3 * kms / max(kn, 1), 1),
                        "visits": round(vis, 1), "digest": h})
print(json.dumps(res))
t.close()
