/*
 * ikd_Tree.h -- drop-in replacement for the reference header of hku-mars/ikd-Tree (ikd-Tree/ikd_Tree.h),
 * backed by the B200-native library libikd_b200.so through the C ABI in ikd_b200.h.
 *
 * Same class template, same public names, signatures and defaults as the reference's public section
 * (reference ikd_Tree.h:225-249); callers such as FAST-LIO2 keep `#include "ikd_Tree.h"` and link
 * -likd_b200 instead of compiling ikd_Tree.cpp. Header-only: every method is a thin inline wrapper that
 * packs arguments, calls one extern "C" function and unpacks the result. The tree itself (nodes, boxes,
 * delete flags) lives in GPU memory; this object only keeps the caller's PointType values by point id so
 * that searches return full points (payload fields of pcl::PointXYZINormal etc. round-trip).
 *
 * Additions over the reference (batched overloads; the per-query methods are kept for compatibility but
 * a batch is what the GPU is for):
 *   Nearest_Search(const PointVector& queries, int k, vector<PointVector>&, vector<vector<float>>&, double)
 *   Nearest_Search_Batch(queries, k, idx, sqdist, count, max_dist)   -- flat arrays, no per-query vectors
 *   Nearest_Plane_Batch(queries, k, plane, residual, valid, ...)     -- kNN + the caller's plane fit, on the device
 *   Box_Search(const vector<BoxPointType>&, vector<PointVector>&)
 *   Radius_Search(const PointVector& centers, const vector<float>& radii, vector<PointVector>&)
 *
 * Differences from the reference, by design:
 *   - no background pthread: rebuilds triggered by the alpha criteria run as CUDA work; there is no window
 *     in which size()/validnum()/tree_range() return stale or -1 values (reference :93, :122, :142);
 *   - Add_Points on a never-built tree builds it (the reference dereferences null, :447/:472);
 *   - errors (no GPU, CUDA failure) throw std::runtime_error instead of being silently ignored;
 *   - Root_Node is an opaque non-null token when a root exists; flatten() accepts only Root_Node.
 *
 * Concurrency (reference ikd_Tree.cpp:371-387, :875-884: any number of threads may call Nearest_Search at once, which
 * is how FAST-LIO2's `#pragma omp parallel for` drives it): concurrent single-query calls are COMBINED -- the first
 * caller to arrive becomes the leader, collects the requests of the other threads for a few microseconds, issues ONE
 * ikd_knn_batch for all of them and hands the results out (flat combining) -- so an unmodified caller gets one kernel
 * launch per round of its thread team instead of one per query.
 *
 * Point ids and memory: the payload array is indexed by point id, and ids only grow (every inserted node gets a fresh
 * one). When dead ids outnumber live points by `id_compaction_slack`, the next mutating call renumbers the live points
 * (ikd_compact_ids) and compacts the payload array, so host and device memory stay proportional to the live map.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ikd_b200.h"

#if defined(IKD_USE_EIGEN_ALLOCATOR) || (defined(__has_include) && __has_include(<Eigen/Core>) && !defined(IKD_NO_EIGEN))
#include <Eigen/Core>
#include <Eigen/StdVector>
#define IKD_POINT_ALLOCATOR(T) Eigen::aligned_allocator<T>
#else
#define IKD_POINT_ALLOCATOR(T) std::allocator<T>
#endif

// Compile-time constants of the reference (ikd_Tree.h:13-18), kept for callers that read them.
#define EPSS 1e-6
#define Minimal_Unbalanced_Tree_Size 10
#define Multi_Thread_Rebuild_Point_Num 1500
#define DOWNSAMPLE_SWITCH true
#define ForceRebuildPercentage 0.2
#define Q_LEN 1000000

// API types, verbatim layout (reference ikd_Tree.h:22-39).
struct ikdTree_PointType {
    float x, y, z;
    ikdTree_PointType(float px = 0.0f, float py = 0.0f, float pz = 0.0f) {
        x = px;
        y = py;
        z = pz;
    }
};

struct BoxPointType {
    float vertex_min[3];
    float vertex_max[3];
};

enum operation_set { ADD_POINT, DELETE_POINT, DELETE_BOX, ADD_BOX, DOWNSAMPLE_DELETE, PUSH_DOWN };

enum delete_point_storage_set { NOT_RECORD, DELETE_POINTS_REC, MULTI_THREAD_REC };

template <typename PointType>
class KD_TREE {
public:
    using PointVector = std::vector<PointType, IKD_POINT_ALLOCATOR(PointType)>;
    using Ptr = std::shared_ptr<KD_TREE<PointType>>;
    // The tree lives on the GPU; this is only a token so that `Root_Node == nullptr` keeps its meaning.
    struct KD_TREE_NODE {
        int unused = 0;
    };

    KD_TREE(float delete_param = 0.5, float balance_param = 0.6, float box_length = 0.2) {
        check(ikd_create(&h_, -1, delete_param, balance_param, box_length), "ikd_create");
    }
    ~KD_TREE() {
        if (h_) ikd_destroy(h_);
    }
    KD_TREE(const KD_TREE&) = delete;
    KD_TREE& operator=(const KD_TREE&) = delete;

    void Set_delete_criterion_param(float delete_param) { check(ikd_set_delete_param(h_, delete_param), "set"); }
    void Set_balance_criterion_param(float balance_param) { check(ikd_set_balance_param(h_, balance_param), "set"); }
    void set_downsample_param(float box_length) { check(ikd_set_downsample_param(h_, box_length), "set"); }
    void InitializeKDTree(float delete_param = 0.5, float balance_param = 0.7, float box_length = 0.2) {
        Set_delete_criterion_param(delete_param);
        Set_balance_criterion_param(balance_param);
        set_downsample_param(box_length);
    }

    int size() {
        int v = 0;
        check(ikd_size(h_, &v), "size");
        return v;
    }
    int validnum() {
        int v = 0;
        check(ikd_validnum(h_, &v), "validnum");
        return v;
    }
    void root_alpha(float& alpha_bal, float& alpha_del) { check(ikd_root_alpha(h_, &alpha_bal, &alpha_del), "root_alpha"); }
    BoxPointType tree_range() {
        BoxPointType b;
        float r[6];
        check(ikd_tree_range(h_, r), "tree_range");
        for (int a = 0; a < 3; a++) {
            b.vertex_min[a] = r[a];
            b.vertex_max[a] = r[3 + a];
        }
        return b;
    }

    // Build(point_cloud): reference ikd_Tree.cpp:353. The cloud is copied (payload by point id).
    void Build(PointVector point_cloud) {
        std::lock_guard<std::mutex> g(mu_);
        payload_.assign(point_cloud.begin(), point_cloud.end());
        const float* p = payload_.empty() ? nullptr : &payload_[0].x;
        check(ikd_build(h_, p, (int64_t)payload_.size(), (int64_t)sizeof(PointType)), "Build");
        removed_backlog_.clear();
        refresh_root();
    }

    // Nearest_Search: reference ikd_Tree.cpp:367. Outputs are cleared, then filled in ascending distance.
    // Safe to call from any number of threads at once; concurrent calls are combined into one GPU launch (see the
    // header comment). A single caller pays one small launch per call; use the batched overloads for speed.
    void Nearest_Search(PointType point, int k_nearest, PointVector& Nearest_Points, std::vector<float>& Point_Distance,
                        double max_dist = INFINITY) {
        PointVector().swap(Nearest_Points);
        std::vector<float>().swap(Point_Distance);
        if (k_nearest < 1) return;
        NsRequest r;
        r.q = &point;
        r.k = k_nearest;
        r.max_dist = max_dist;
        r.pts = &Nearest_Points;
        r.dist = &Point_Distance;
        bool leader;
        {
            std::lock_guard<std::mutex> g(cq_mu_);
            cq_.push_back(&r);
            leader = !cq_leader_;
            if (leader) cq_leader_ = true;
        }
        if (leader) {
            combine_and_serve();
        } else {
            int spins = 0;
            while (r.state.load(std::memory_order_acquire) == 0) {
                if (++spins > 2000) { std::this_thread::yield(); spins = 0; }
            }
        }
        if (r.err) std::rethrow_exception(r.err);
    }

    // Batched overload: one GPU launch for all queries.
    void Nearest_Search(const PointVector& queries, int k_nearest, std::vector<PointVector>& Nearest_Points,
                        std::vector<std::vector<float>>& Point_Distance, double max_dist = INFINITY) {
        size_t nq = queries.size();
        Nearest_Points.assign(nq, PointVector());
        Point_Distance.assign(nq, std::vector<float>());
        if (nq == 0 || k_nearest < 1) return;
        std::vector<int32_t> idx, cnt;
        std::vector<float> d;
        std::lock_guard<std::mutex> g(mu_);  // held across the payload gather: a concurrent Add_Points may move payload_
        knn_batch_locked(queries, k_nearest, idx, d, cnt, max_dist);
        for (size_t i = 0; i < nq; i++) {
            for (int j = 0; j < cnt[i]; j++) {
                Nearest_Points[i].push_back(at(idx[i * k_nearest + j]));
                Point_Distance[i].push_back(d[i * k_nearest + j]);
            }
        }
    }

    // Flat batched form: idx/sqdist are nq*k (row-major, -1 / +inf padded), count is nq. idx values are
    // point ids; Point(id) returns the stored point. k_nearest < 1 returns empty rows (count 0), like the single-query form.
    void Nearest_Search_Batch(const PointVector& queries, int k_nearest, std::vector<int32_t>& idx,
                              std::vector<float>& sqdist, std::vector<int32_t>& count, double max_dist = INFINITY) {
        std::lock_guard<std::mutex> g(mu_);
        knn_batch_locked(queries, k_nearest, idx, sqdist, count, max_dist);
    }
    const PointType& Point(int32_t id) const { return at(id); }

    // kNN followed on the device by the plane fit FAST-LIO2 runs on the neighbours (h_share_model / esti_plane):
    // plane = nq*4 (unit normal a,b,c and d), residual = a*x+b*y+c*z+d of the query, valid = 1 when k neighbours were
    // found within max_dist, the k-th squared distance is <= max_kth_sqdist and every neighbour is within
    // plane_threshold of the plane. The neighbours stay in GPU memory (see ikd_knn_plane_batch in ikd_b200.h).
    void Nearest_Plane_Batch(const PointVector& queries, int k_nearest, std::vector<float>& plane,
                             std::vector<float>& residual, std::vector<uint8_t>& valid, double max_dist = INFINITY,
                             float max_kth_sqdist = 5.0f, float plane_threshold = 0.1f) {
        std::lock_guard<std::mutex> g(mu_);
        size_t nq = queries.size();
        plane.assign(nq * 4, 0.f);
        residual.assign(nq, 0.f);
        valid.assign(nq, 0);
        if (nq == 0) return;
        if (k_nearest < IKD_PLANE_MIN_K || k_nearest > IKD_PLANE_MAX_K) throw std::invalid_argument("Nearest_Plane_Batch: k out of range");
        check(ikd_knn_plane_batch(h_, &queries[0].x, (int64_t)nq, (int64_t)sizeof(PointType), k_nearest, max_dist,
                                  max_kth_sqdist, plane_threshold, plane.data(), residual.data(), valid.data(), nullptr),
              "Nearest_Plane_Batch");
    }

    // Box_Search: reference ikd_Tree.cpp:400 (half-open box [min,max)).
    void Box_Search(const BoxPointType& Box_of_Point, PointVector& Storage) {
        std::vector<BoxPointType> b(1, Box_of_Point);
        std::vector<PointVector> out;
        Box_Search(b, out);
        Storage.swap(out[0]);
    }
    void Box_Search(const std::vector<BoxPointType>& Boxes, std::vector<PointVector>& Storage) {
        std::lock_guard<std::mutex> g(mu_);
        size_t nb = Boxes.size();
        std::vector<int64_t> off(nb + 1, 0);
        check(ikd_box_search_batch(h_, nb ? Boxes[0].vertex_min : nullptr, (int64_t)nb, off.data()), "Box_Search");
        unpack_search(off, nb, Storage);
    }

    // Radius_Search: reference ikd_Tree.cpp:407.
    void Radius_Search(PointType point, const float radius, PointVector& Storage) {
        PointVector c(1, point);
        std::vector<float> r(1, radius);
        std::vector<PointVector> out;
        Radius_Search(c, r, out);
        Storage.swap(out[0]);
    }
    void Radius_Search(const PointVector& centers, const std::vector<float>& radii, std::vector<PointVector>& Storage) {
        std::lock_guard<std::mutex> g(mu_);
        size_t nq = centers.size();
        if (radii.size() != nq) throw std::invalid_argument("Radius_Search: centers and radii differ in length");
        std::vector<float> c(3 * nq);
        for (size_t i = 0; i < nq; i++) {
            c[3 * i] = centers[i].x;
            c[3 * i + 1] = centers[i].y;
            c[3 * i + 2] = centers[i].z;
        }
        std::vector<int64_t> off(nq + 1, 0);
        check(ikd_radius_search_batch(h_, c.data(), radii.data(), (int64_t)nq, off.data()), "Radius_Search");
        unpack_search(off, nq, Storage);
    }

    // Add_Points: reference ikd_Tree.cpp:414. Returns what the reference returns (number of inserts done by
    // the downsample branch; 0 when downsample_on is false).
    int Add_Points(PointVector& PointToAdd, bool downsample_on) {
        std::lock_guard<std::mutex> g(mu_);
        size_t n = PointToAdd.size();
        if (n == 0) return 0;
        int added = 0;
        int32_t first = 0;
        int64_t nins = 0;
        std::vector<int32_t> src(n);
        check(ikd_add_points(h_, &PointToAdd[0].x, (int64_t)n, (int64_t)sizeof(PointType), downsample_on ? 1 : 0, &added,
                             &first, &nins, src.data()),
              "Add_Points");
        if ((size_t)first != payload_.size()) throw std::runtime_error("ikd_Tree: point id bookkeeping out of sync");
        payload_.reserve(payload_.size() + (size_t)nins);
        for (int64_t i = 0; i < nins; i++) {
            if (src[i] >= 0) payload_.push_back(PointToAdd[(size_t)src[i]]);
            else payload_.push_back(PointType(at(~src[i])));  // downsample winner already in the tree (:441, :447)
        }
        refresh_root();
        maybe_compact_ids();
        return added;
    }

    // Add_Point_Boxes: reference ikd_Tree.cpp:492.
    void Add_Point_Boxes(std::vector<BoxPointType>& BoxPoints) {
        std::lock_guard<std::mutex> g(mu_);
        if (BoxPoints.empty()) return;
        check(ikd_add_boxes(h_, BoxPoints[0].vertex_min, (int64_t)BoxPoints.size()), "Add_Point_Boxes");
        refresh_root();
        maybe_compact_ids();
    }

    // Delete_Points: reference ikd_Tree.cpp:514.
    void Delete_Points(PointVector& PointToDel) {
        std::lock_guard<std::mutex> g(mu_);
        if (PointToDel.empty()) return;
        check(ikd_delete_points(h_, &PointToDel[0].x, (int64_t)PointToDel.size(), (int64_t)sizeof(PointType)),
              "Delete_Points");
        refresh_root();
        maybe_compact_ids();
    }

    // Delete_Point_Boxes: reference ikd_Tree.cpp:536. Returns the number of newly deleted points.
    int Delete_Point_Boxes(std::vector<BoxPointType>& BoxPoints) {
        std::lock_guard<std::mutex> g(mu_);
        if (BoxPoints.empty()) return 0;
        int n = 0;
        check(ikd_delete_boxes(h_, BoxPoints[0].vertex_min, (int64_t)BoxPoints.size(), &n), "Delete_Point_Boxes");
        refresh_root();
        maybe_compact_ids();
        return n;
    }

    // flatten(Root_Node, Storage, storage_type): reference ikd_Tree.cpp:1326. Appends every valid point.
    void flatten(KD_TREE_NODE* root, PointVector& Storage, delete_point_storage_set /*storage_type*/) {
        if (root == nullptr) return;
        std::lock_guard<std::mutex> g(mu_);
        int64_t n = 0;
        check(ikd_flatten(h_, nullptr, 0, &n), "flatten");
        std::vector<int32_t> ids((size_t)n);
        if (n) check(ikd_flatten(h_, ids.data(), n, &n), "flatten");
        for (int64_t i = 0; i < n; i++) Storage.push_back(at(ids[i]));
    }

    // acquire_removed_points: reference ikd_Tree.cpp:559.
    void acquire_removed_points(PointVector& removed_points) {
        std::lock_guard<std::mutex> g(mu_);
        for (const PointType& p : removed_backlog_) removed_points.push_back(p);  // drained before an id compaction
        removed_backlog_.clear();
        int64_t n = 0;
        check(ikd_acquire_removed(h_, nullptr, 0, &n), "acquire_removed_points");
        std::vector<int32_t> ids((size_t)std::max<int64_t>(n, 1));
        check(ikd_acquire_removed(h_, ids.data(), n, &n), "acquire_removed_points");
        for (int64_t i = 0; i < n; i++) removed_points.push_back(at(ids[i]));
    }

    ikd_tree* handle() { return h_; }
    // Dead point ids tolerated before the payload array and the id space are compacted (see the header comment).
    int64_t id_compaction_slack = (int64_t)1 << 20;
    size_t payload_size() const { return payload_.size(); }

    PointVector PCL_Storage;
    KD_TREE_NODE* Root_Node = nullptr;
    int max_queue_size = 0;  // the reference's op-log high-water mark; there is no op log here

private:
    void check(int status, const char* what) {
        if (status != IKD_OK)
            throw std::runtime_error(std::string("ikd_Tree(B200) ") + what + " failed: " + ikd_last_error());
    }
    void refresh_root() {
        int has = 0;
        check(ikd_has_root(h_, &has), "has_root");
        Root_Node = has ? &root_token_ : nullptr;
    }
    void unpack_search(const std::vector<int64_t>& off, size_t nq, std::vector<PointVector>& Storage) {
        int64_t total = off[nq];
        std::vector<int32_t> ids((size_t)std::max<int64_t>(total, 1));
        if (total) check(ikd_search_fetch(h_, ids.data(), total), "search_fetch");
        Storage.assign(nq, PointVector());
        for (size_t i = 0; i < nq; i++) {
            Storage[i].reserve((size_t)(off[i + 1] - off[i]));
            for (int64_t j = off[i]; j < off[i + 1]; j++) Storage[i].push_back(at(ids[(size_t)j]));
        }
    }
    // payload of a point id handed out by the library (bounds-checked: a stale id must not index past the array)
    const PointType& at(int32_t id) const {
        if (id < 0 || (size_t)id >= payload_.size()) throw std::runtime_error("ikd_Tree: point id out of range (stale id?)");
        return payload_[(size_t)id];
    }
    // caller holds mu_
    void knn_batch_locked(const PointVector& queries, int k_nearest, std::vector<int32_t>& idx, std::vector<float>& sqdist,
                          std::vector<int32_t>& count, double max_dist) {
        size_t nq = queries.size();
        count.assign(nq, 0);
        if (k_nearest < 1) { idx.clear(); sqdist.clear(); return; }
        idx.assign(nq * (size_t)k_nearest, -1);
        sqdist.assign(nq * (size_t)k_nearest, INFINITY);
        if (nq == 0) return;
        check(ikd_knn_batch(h_, &queries[0].x, (int64_t)nq, (int64_t)sizeof(PointType), k_nearest, max_dist, idx.data(),
                            sqdist.data(), count.data()),
              "Nearest_Search_Batch");
    }
    // caller holds mu_. Renumber the live points when dead ids dominate (ikd_compact_ids) and compact payload_.
    void maybe_compact_ids() {
        int64_t next = 0;
        check(ikd_next_id(h_, &next), "next_id");
        int valid = 0, sz = 0;
        check(ikd_validnum(h_, &valid), "validnum");
        check(ikd_size(h_, &sz), "size");
        if (next <= 2 * (int64_t)std::max(valid, 0) + id_compaction_slack) return;
        int64_t nrem = 0;
        check(ikd_acquire_removed(h_, nullptr, 0, &nrem), "acquire_removed_points");
        std::vector<int32_t> old_of_new((size_t)std::max(valid, 1));
        std::vector<int32_t> rem((size_t)(nrem + std::max(sz - valid, 0) + 16));
        int64_t na = 0, nr = 0;
        check(ikd_compact_ids(h_, old_of_new.data(), (int64_t)old_of_new.size(), &na, rem.data(), (int64_t)rem.size(), &nr),
              "compact_ids");
        for (int64_t i = 0; i < nr; i++) removed_backlog_.push_back(at(rem[(size_t)i]));
        std::vector<PointType, IKD_POINT_ALLOCATOR(PointType)> np;
        np.reserve((size_t)na);
        for (int64_t i = 0; i < na; i++) np.push_back(at(old_of_new[(size_t)i]));
        payload_.swap(np);
        refresh_root();
    }

    // ---- flat combining of concurrent single-query Nearest_Search calls ------------------------------------
    struct NsRequest {
        const PointType* q = nullptr;
        int k = 0;
        double max_dist = 0;
        PointVector* pts = nullptr;
        std::vector<float>* dist = nullptr;
        std::atomic<int> state{0};  // 0 waiting, 1 served
        std::exception_ptr err;
    };
    // The leader: gather what the other threads have queued (waiting a few microseconds for a team of the size seen in
    // the last round), serve everything with one ikd_knn_batch per (k, max_dist) group, repeat while requests keep
    // coming, then give the leadership up.
    void combine_and_serve() {
        std::vector<NsRequest*> batch;
        for (;;) {
            if (cq_expected_ > 1) {  // let the rest of the team arrive (bounded: ~20 us)
                auto t0 = std::chrono::steady_clock::now();
                for (;;) {
                    size_t have;
                    { std::lock_guard<std::mutex> g(cq_mu_); have = cq_.size(); }
                    if (have >= cq_expected_) break;
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(20)) break;
                }
            }
            {
                std::lock_guard<std::mutex> g(cq_mu_);
                batch.swap(cq_);
                if (batch.empty()) { cq_leader_ = false; return; }
            }
            cq_expected_ = batch.size();
            serve(batch);
            batch.clear();
        }
    }
    void serve(std::vector<NsRequest*>& batch) {
        std::vector<char> done(batch.size(), 0);
        std::vector<float> q;
        std::vector<int32_t> idx, cnt;
        std::vector<float> d;
        std::vector<size_t> members;
        for (size_t i = 0; i < batch.size(); i++) {
            if (done[i]) continue;
            const int k = batch[i]->k;
            const double md = batch[i]->max_dist;
            members.clear();
            for (size_t j = i; j < batch.size(); j++)
                if (!done[j] && batch[j]->k == k && (batch[j]->max_dist == md || (md != md && batch[j]->max_dist != batch[j]->max_dist))) {
                    members.push_back(j);
                    done[j] = 1;
                }
            const size_t n = members.size();
            q.resize(3 * n);
            for (size_t m = 0; m < n; m++) {
                const PointType& p = *batch[members[m]]->q;
                q[3 * m] = p.x; q[3 * m + 1] = p.y; q[3 * m + 2] = p.z;
            }
            idx.assign(n * (size_t)k, -1);
            d.assign(n * (size_t)k, INFINITY);
            cnt.assign(n, 0);
            std::exception_ptr err;
            try {
                std::lock_guard<std::mutex> g(mu_);
                check(ikd_knn_batch(h_, q.data(), (int64_t)n, 12, k, md, idx.data(), d.data(), cnt.data()), "Nearest_Search");
                for (size_t m = 0; m < n; m++) {
                    NsRequest* r = batch[members[m]];
                    for (int j = 0; j < cnt[m]; j++) {
                        r->pts->push_back(at(idx[m * (size_t)k + j]));
                        r->dist->push_back(d[m * (size_t)k + j]);
                    }
                }
            } catch (...) {
                err = std::current_exception();
            }
            for (size_t m = 0; m < n; m++) {
                NsRequest* r = batch[members[m]];
                r->err = err;
                r->state.store(1, std::memory_order_release);  // (the leader's own request is in the batch too)
            }
        }
    }

    ikd_tree* h_ = nullptr;
    std::mutex mu_;
    std::mutex cq_mu_;                 // guards cq_ and cq_leader_
    std::vector<NsRequest*> cq_;       // queued single-query requests
    bool cq_leader_ = false;
    size_t cq_expected_ = 1;           // size of the last combined round (only the leader touches it)
    PointVector removed_backlog_;      // removed points whose ids were retired by an id compaction
    std::vector<PointType, IKD_POINT_ALLOCATOR(PointType)> payload_;  // PointType by point id
    KD_TREE_NODE root_token_;
};
