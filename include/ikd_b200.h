/*
 * ikd_b200.h -- C ABI of the B200-native incremental k-d tree (libikd_b200.so).
 *
 * This is the drop-in boundary for the hot path of hku-mars/ikd-Tree. The reference has no FFI
 * boundary of its own: callers include ikd_Tree.h and link ikd_Tree.cpp (CMakeLists.txt:15-22), so the
 * boundary is the public section of KD_TREE<PointType> (ikd_Tree.h:225-249). include/ikd_Tree.h in this
 * repo re-creates that class template with the same names and signatures; its inline methods call only
 * the functions declared here. Each entry point cites the reference interface it replaces.
 *
 * Conventions: plain C types; all pointers are caller-owned HOST buffers unless the name says "_dev";
 * every function returns an int status (IKD_OK == 0) and reports results through out-parameters;
 * nothing throws across the boundary; one handle == one tree replica on one GPU. Points are identified
 * by a stable 32-bit point id: Build assigns 0..n-1 in input order, Add_Points assigns the next ids
 * to the points it actually inserts, in order. Ids let the host-side wrapper keep the non-xyz payload of
 * PointType. There is NO CPU fallback: every call fails with IKD_ERR_CUDA if no sm_100-class device is
 * usable.
 */
#ifndef IKD_B200_H_
#define IKD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ikd_tree ikd_tree; /* opaque handle */

enum {
    IKD_OK = 0,
    IKD_ERR_CUDA = 1,      /* CUDA runtime/driver failure or no device; ikd_last_error() has the text */
    IKD_ERR_ARG = 2,       /* bad argument (null pointer, negative size, k out of range) */
    IKD_ERR_NOT_BUILT = 3, /* operation needs a built tree (reference would dereference null: ikd_Tree.cpp:447,472) */
    IKD_ERR_CAPACITY = 4,  /* result buffer too small / device memory exhausted */
    IKD_ERR_INTERNAL = 5
};

#define IKD_MAX_K 128 /* largest k accepted by ikd_knn_batch */

/* Human-readable text of the last error on this thread. */
const char* ikd_last_error(void);
/* Library/ABI version, bumped when a signature changes. */
int ikd_abi_version(void);
/* Number of kernels of this library launched by the calling process so far (library kernels such as CUB's
 * sorts and scans are not counted). bench.py reports the difference over its timed region. */
long long ikd_launch_count(void);

/* KD_TREE(delete_param, balance_param, box_length) ikd_Tree.cpp:10 ; ~KD_TREE ikd_Tree.cpp:20.
 * device < 0 selects the current CUDA device. */
int ikd_create(ikd_tree** out, int device, float delete_param, float balance_param, float box_length);
int ikd_destroy(ikd_tree* t);

/* Set_delete_criterion_param / Set_balance_criterion_param / set_downsample_param  ikd_Tree.cpp:30-42 */
int ikd_set_delete_param(ikd_tree* t, float v);
int ikd_set_balance_param(ikd_tree* t, float v);
int ikd_set_downsample_param(ikd_tree* t, float v);

/* size() :79, validnum() :129, root_alpha() :148, tree_range() :99. range6 = min[3], max[3]. */
int ikd_size(ikd_tree* t, int* out);
int ikd_validnum(ikd_tree* t, int* out);
int ikd_root_alpha(ikd_tree* t, float* alpha_bal, float* alpha_del);
int ikd_tree_range(ikd_tree* t, float* range6);
/* 1 if a root exists (reference: Root_Node != nullptr). */
int ikd_has_root(ikd_tree* t, int* out);

/* Build(point_cloud) ikd_Tree.cpp:353 -> BuildTree :574. xyz points to n points, stride_bytes apart
 * (first three floats of each are x,y,z). Point ids 0..n-1. Replaces any previous tree. n == 0 leaves
 * the tree empty (reference: Root_Node stays null, :357). */
int ikd_build(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes);

/* Batched Nearest_Search(point, k, pts, dist, max_dist) ikd_Tree.cpp:367 -> Search :869.
 * q: nq queries, stride_bytes apart. out_idx / out_sqdist: nq*k entries, row i holds the neighbours of
 * query i in ascending (squared distance, point id) order; unused tail entries are -1 / +inf.
 * out_count[i] <= k is the number found (fewer when fewer than k points lie within max_dist).
 * max_dist follows the reference: candidates need dist <= max_dist*max_dist evaluated in double. */
int ikd_knn_batch(ikd_tree* t, const float* q, int64_t nq, int64_t stride_bytes, int k, double max_dist,
                  int32_t* out_idx, float* out_sqdist, int32_t* out_count);

/* Device-resident variant of the same call: q_dev is float4-packed (x,y,z,unused), outputs are device
 * pointers. Enqueued on the tree's stream; returns without synchronising. This is the kernel-only path
 * bench.py times for `value`. */
int ikd_knn_batch_dev(ikd_tree* t, const void* q_dev_float4, int64_t nq, int k, double max_dist,
                      int32_t* out_idx_dev, float* out_sqdist_dev, int32_t* out_count_dev);

/* Nearest_Search followed, on the device, by the step its known caller runs next (FAST-LIO2
 * laserMapping.cpp h_share_model / esti_plane -- external to the reference tree, so an extension of the path):
 * for every query take its k nearest neighbours (IKD_PLANE_MIN_K <= k <= IKD_PLANE_MAX_K, the caller uses 5),
 * solve the k x 3 least-squares system A n = -1 (rows of A = neighbour coordinates) by column-pivoted Householder
 * QR in fp32, and return the unit normal and offset (a, b, c, d = 1/|n|) in out_plane (nq*4 floats) and the query's
 * signed distance a*x + b*y + c*z + d in out_resid (nq floats). out_valid[i] = 1 when k neighbours were found within
 * max_dist, the k-th squared distance is <= max_kth_sqdist (caller: 5), the system has full rank and every
 * neighbour lies within plane_threshold (caller: 0.1) of the plane; otherwise 0. Rows that fail the first two
 * tests (no fit attempted) or whose fit is not finite hold zeros. out_idx (nq*k, may be null) receives the neighbour
 * ids as ikd_knn_batch would. The neighbours themselves never leave the device. */
#define IKD_PLANE_MIN_K 3
#define IKD_PLANE_MAX_K 8
int ikd_knn_plane_batch(ikd_tree* t, const float* q, int64_t nq, int64_t stride_bytes, int k, double max_dist,
                        float max_kth_sqdist, float plane_threshold, float* out_plane, float* out_resid,
                        uint8_t* out_valid, int32_t* out_idx);
/* Device-resident variant (q_dev float4-packed, outputs device pointers; enqueued on the tree's stream, no sync). */
int ikd_knn_plane_batch_dev(ikd_tree* t, const void* q_dev_float4, int64_t nq, int k, double max_dist,
                            float max_kth_sqdist, float plane_threshold, float* out_plane_dev, float* out_resid_dev,
                            uint8_t* out_valid_dev);

/* Batched Box_Search(box, storage) ikd_Tree.cpp:400 -> Search_by_range :1016. boxes: nb * 6 floats
 * (min[3], max[3]), half-open [min,max). Two-phase: phase 1 returns per-box counts and exclusive
 * offsets (out_offsets has nb+1 entries, last = total); phase 2 copies the point ids of the last
 * search (cap entries available). Results of box i are out_idx[offsets[i] .. offsets[i+1]). */
int ikd_box_search_batch(ikd_tree* t, const float* boxes, int64_t nb, int64_t* out_offsets);
/* Batched Radius_Search(point, radius, storage) ikd_Tree.cpp:407 -> Search_by_radius :1047.
 * centers: nq * 3 floats; radii: nq floats. Same two-phase protocol. */
int ikd_radius_search_batch(ikd_tree* t, const float* centers, const float* radii, int64_t nq,
                            int64_t* out_offsets);
/* Phase 2 of either search: copy the ids found by the last search call. */
int ikd_search_fetch(ikd_tree* t, int32_t* out_idx, int64_t cap);

/* Coordinates of points by id (ids from kNN / search results). out_xyz: n*3 floats. */
int ikd_get_points(ikd_tree* t, const int32_t* ids, int64_t n, float* out_xyz);

/* Add_Points(PointToAdd, downsample_on) ikd_Tree.cpp:414. Returns the reference's return value (number
 * of insert branches taken) in *out_added. out_src (n entries, may be null) receives, for each point
 * that ended up inserted as a node that survives the call, where its payload comes from:
 * value >= 0: index into xyz; value < 0: ~(existing point id) (a downsample winner that was already in
 * the tree and is re-inserted as a new node, :445-447). *out_first_id is the id of the first such point;
 * ids are consecutive. *out_ninserted is their number. */
int ikd_add_points(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes, int downsample_on,
                   int* out_added, int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src);

/* Device-resident variant of ikd_add_points: pts_dev is float4-packed (x,y,z,unused) in device memory.
 * out_src may be null. This is the call bench.py times for `value` (inputs already in HBM). */
int ikd_add_points_dev(ikd_tree* t, const void* pts_dev_float4, int64_t n, int downsample_on, int* out_added,
                       int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src);

/* Delete_Points(PointToDel) ikd_Tree.cpp:514 -> Delete_by_point :713. */
int ikd_delete_points(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes);
/* Device-resident variants of the three calls around this comment (inputs already in HBM: float4-packed points,
 * 6-float boxes). Used to apply a broadcast update batch on every replica without a host bounce. */
int ikd_delete_points_dev(ikd_tree* t, const void* pts_dev_float4, int64_t n);
int ikd_delete_boxes_dev(ikd_tree* t, const float* boxes_dev, int64_t nb, int* out_deleted);
int ikd_add_boxes_dev(ikd_tree* t, const float* boxes_dev, int64_t nb);
/* Delete_Point_Boxes(BoxPoints) ikd_Tree.cpp:536 -> Delete_by_range :648. *out_deleted = newly deleted points. */
int ikd_delete_boxes(ikd_tree* t, const float* boxes, int64_t nb, int* out_deleted);
/* Add_Point_Boxes(BoxPoints) ikd_Tree.cpp:492 -> Add_by_range :763 (SURVEY 8f "next" #1). */
int ikd_add_boxes(ikd_tree* t, const float* boxes, int64_t nb);

/* flatten(Root_Node, storage, NOT_RECORD) ikd_Tree.cpp:1326: ids of all valid points. Two-phase:
 * *out_n = count; if out_idx != null, copies min(count, cap) ids. */
int ikd_flatten(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n);
/* acquire_removed_points ikd_Tree.cpp:559: ids of points dropped by rebuilds since the last call
 * (lazy-deleted, not downsample-deleted). Same two-phase protocol; the list is cleared when copied. */
int ikd_acquire_removed(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n);

/* Point ids grow with every inserted point (a re-inserted downsample winner gets a fresh id too), so a map that runs
 * for hours accumulates dead ids although its valid point count stays bounded. *out_next_id = ids handed out so far
 * (compare with ikd_validnum). Inserts fail with IKD_ERR_CAPACITY before the 31-bit id space overflows. */
int ikd_next_id(ikd_tree* t, int64_t* out_next_id);
/* Renumber the valid points 0..M-1 in increasing old-id order, rebuild the tree on them (this also compacts the node
 * pool) and shrink the id space to M. old_of_new[i] = old id of new id i (cap_alive >= ikd_validnum). The removed-point
 * log is drained into removed_old (OLD ids; may hold up to the current log length + size - validnum entries) because
 * its ids would be void afterwards. Fails with IKD_ERR_CAPACITY, changing nothing, when a buffer is too small.
 * Ids obtained before the call are void afterwards (ikd_id_epoch changes). include/ikd_Tree.h calls this on its own
 * when dead ids outnumber live ones and compacts its payload array with the result. */
int ikd_compact_ids(ikd_tree* t, int32_t* old_of_new, int64_t cap_alive, int64_t* out_alive, int32_t* removed_old,
                    int64_t cap_removed, int64_t* out_removed);
int ikd_id_epoch(ikd_tree* t, int64_t* out_epoch);

/* Wait for all work enqueued on this tree (including a side-stream rebuild) to finish. */
int ikd_synchronize(ikd_tree* t);

/* Introspection for tests and DESIGN.md figures. */
typedef struct ikd_stats {
    int64_t node_slots_used;  /* bump pointer of the node pool */
    int64_t node_slots_cap;
    int32_t max_depth;        /* upper bound on the depth of the tree */
    int32_t rebuilds_partial; /* subtree rebuilds triggered by the alpha criteria since create */
    int32_t rebuilds_full;    /* whole-tree rebuilds (criteria at the root, or pool compaction) */
    int32_t rebuilds_async;   /* of the above, how many ran on the side stream */
    int64_t rebuilt_points;   /* points passed through rebuilds */
    int64_t last_knn_visits;  /* node visits of the last ikd_knn_batch* call if visit counting is on, else -1 */
    /* Add_Points(downsample) work counters, accumulated while visit counting is on (the algorithmic-bytes figure of
     * 12 + 64*(V_box + 2*depth) per added point): */
    int64_t add_points_in;        /* input points of downsampled Add_Points calls */
    int64_t add_points_inserted;  /* points that became nodes */
    int64_t add_vox_visits;       /* sum over input points of the nodes visited by their voxel's box search */
    int64_t add_descend_levels;   /* sum over inserted points of the levels descended to the insert position */
    /* rebuild timing (ikd_set_rebuild_timing): device milliseconds and counts by kind */
    double rebuild_inline_ms;  int64_t rebuild_inline_n;   /* subtrees rebuilt on the tree's stream (one entry per batch) */
    double rebuild_async_ms;   int64_t rebuild_async_n;    /* side-stream rebuilds (>= 2049 points) */
    double rebuild_full_ms;    int64_t rebuild_full_n;     /* whole-tree rebuilds */
    double rebuild_max_ms;                                  /* longest single entry of any kind */
} ikd_stats;
int ikd_get_stats(ikd_tree* t, ikd_stats* out);
/* Kernel timing for the roofline figure: when on, every ikd_knn_batch* call brackets its traversal kernel
 * (not the Morton sort) with CUDA events on the tree's stream. ikd_get_kernel_time synchronises, returns
 * the accumulated milliseconds and launch count since the last call, and resets them. */
int ikd_set_kernel_timing(ikd_tree* t, int on);
int ikd_get_kernel_time(ikd_tree* t, double* out_ms, int64_t* out_launches);
/* Turn the per-launch node-visit counter on/off (off by default; it costs an atomic per query). */
int ikd_set_visit_counting(ikd_tree* t, int on);
/* Bracket every rebuild with CUDA events; ikd_get_stats resolves the finished ones into the rebuild_* fields. */
int ikd_set_rebuild_timing(ikd_tree* t, int on);
/* Pre-order structure dump for parity tests: 16 floats per node, columns as oracle/ref_harness.cpp
 * ref_dump_tree. *out_n = number of nodes; copies min(n, cap). */
int ikd_dump_tree(ikd_tree* t, float* out, int64_t cap, int64_t* out_n);

/* Replica sync (multi-GPU query sharding). The tree state is a small header plus four device
 * arrays (search records, update records, walk records, point coordinates by id). ikd_replica_export gives their device
 * pointers and byte sizes on the source replica; the caller sends `slots` and `npoints` to the peers, each
 * peer calls ikd_replica_prepare (allocates and returns its own pointers, same byte sizes), the caller
 * broadcasts the five buffers with NCCL (torch.distributed) and each peer calls ikd_replica_commit. */
typedef struct ikd_replica_desc {
    void* header_dev;  int64_t header_bytes;
    void* search_dev;  int64_t search_bytes;
    void* update_dev;  int64_t update_bytes;
    void* walk_dev;    int64_t walk_bytes;    /* 16-byte enumeration records (child links, deleted bit, point id, ids of leaf children) */
    void* points_dev;  int64_t points_bytes;
    int64_t slots;     /* node slots covered by the two record arrays */
    int64_t npoints;   /* point ids covered by points_dev */
} ikd_replica_desc;
int ikd_replica_export(ikd_tree* t, ikd_replica_desc* out);
int ikd_replica_prepare(ikd_tree* t, int64_t slots, int64_t npoints, ikd_replica_desc* out);
int ikd_replica_commit(ikd_tree* t);

/* The CUDA stream the tree enqueues on (cudaStream_t as void*), for callers that time with events. */
int ikd_stream(ikd_tree* t, void** out_stream);

#ifdef __cplusplus
}
#endif
#endif /* IKD_B200_H_ */
