// Host-side tree object behind the opaque ikd_tree handle, plus small utilities shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ikd_b200.h"
#include "ikd_node.cuh"

#include <atomic>

namespace ikd {

void set_error(const char* fmt, ...);

// Number of kernels of this library launched by the process (CUB kernels not counted); bench.py reports it.
extern std::atomic<long long> g_launches;
inline int count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); return 0; }
#define IKD_LAUNCH (void)ikd::count_launch(),

// Launch with programmatic dependent launch allowed (the kernel must start with pdl_wait(), ikd_node.cuh).
// IKD_NO_PDL=1 in the environment turns the attribute off (A/B measurements).
extern bool g_use_pdl;
template <typename... Exp, typename... Act>
inline cudaError_t launch_pdl(void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Exp>(args)...);
}
#define IKD_LAUNCH_PDL(kern, grid, block, smem, stream, ...) \
    ((void)ikd::count_launch(), (void)ikd::launch_pdl(kern, dim3(grid), dim3(block), smem, stream, __VA_ARGS__))

#define IKD_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            ikd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return IKD_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define IKD_TRY(call)                  \
    do {                               \
        int s_ = (call);               \
        if (s_ != IKD_OK) return s_;   \
    } while (0)

// Grow-only device buffer (scratch or persistent). Never shrinks; contents are lost on growth unless
// `preserve` is asked for.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need, cudaStream_t s, bool preserve = false);
    void release();
    template <class T> T* as() const { return (T*)p; }
};

// Forest description for the level-by-level builder: R disjoint point segments, each becoming one
// balanced subtree whose root lands in root_slot[r] and whose other nodes live in a heap-ordered block.
struct ForestDev {
    int R = 0;
    const int* seg_begin = nullptr;    // R+1 (device)
    const int* root_slot = nullptr;    // R
    const int* block_base = nullptr;   // R, even; local heap index h>=2 lives at block_base+h
    const int* root_parent = nullptr;  // R, parent slot of the root (0 = tree root)
    const int* root_depth = nullptr;   // R
    const int* single_axis = nullptr;  // R, axis for a 1-point subtree (Add_by_point leaf rule) or -1
    const int* elem_root = nullptr;    // M, root index of every element/position (null when R == 1)
};

}  // namespace ikd

namespace ikd {
// per-lane scratch of the kNN path (two lanes so that host-buffer calls can pipeline H2D / search / D2H)
struct KnnScratch {
    DevBuf mkeys, mkeys2, perm, perm2, cubtmp, counter, hist;
    int hist_sel = 0;  // which of the two query-ordering histograms the next call counts into
    DevBuf q3, q4, out_idx, out_d, out_cnt;   // host-path staging on the device
    DevBuf plane, resid, valid;               // plane-fit outputs of the host path
    cudaStream_t stream = nullptr;            // lane stream of the host path
    cudaEvent_t done = nullptr;
    void* pin_in = nullptr;  size_t pin_in_bytes = 0;
    void* pin_out = nullptr; size_t pin_out_bytes = 0;
};
}  // namespace ikd

struct ikd_tree {
    int device = 0;
    static constexpr int KNN_LANES = 4;
    ikd::KnnScratch knn_scr[KNN_LANES];
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t side_done = nullptr;
    // helper streams of the forest builder (size classes build concurrently); [0] for `stream`, [1] for `side`
    cudaStream_t aux[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    cudaEvent_t aux_ev[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    cudaEvent_t aux_fork[2] = {nullptr, nullptr};
    float delete_param = 0.5f, balance_param = 0.6f, downsample = 0.2f;

    // node pool
    ikd::SearchRec* srec = nullptr;
    ikd::UpdateRec* urec = nullptr;
    ikd::WalkRec* wrec = nullptr;  // WalkRec per slot (ikd_node.cuh)
    size_t cap_slots = 0;
    bool pool_from_malloc = false;  // the node arrays came from cudaMalloc (large pools) rather than the stream-ordered pool
    size_t pool_reserved = 0;  // scratch reserved in the stream-ordered pool at Build time
    ikd::TreeHeader* hdr_dev = nullptr;
    ikd::TreeHeader* hdr_pin = nullptr;  // pinned host mirror
    // mapped pinned page for small device -> host reads that the host polls (fetch_small); last 64 B = sequence word
    static constexpr size_t MAPPED_BYTES = 4096;
    void* map_host = nullptr;
    void* map_dev = nullptr;
    uint32_t map_seq = 0;
    ikd::TreeHeader hdr;                 // last synced copy

    // coordinates by point id (float4: xyz + unused), device
    ikd::DevBuf pid_xyz;
    int64_t pid_cap = 0;
    int next_pid = 0;

    // scratch
    ikd::DevBuf b_p4, b_keys0, b_keys1, b_ord[3], b_ord_alt[3], b_cubtmp, b_pos, b_cls, b_scan, b_mpos, b_flag,
        b_segaxis, b_forest, b_q, b_perm, b_mkeys, b_mkeys2, b_perm2, b_out_idx, b_out_d, b_out_cnt, b_misc[8];
    ikd::DevBuf u[32];  // scratch of the update path (indices: enum in ikd_update.cu)
    int64_t rinfo_stride = 0;  // layout of the rebuild planner's arrays inside u[U_RINFO]
    // side-stream rebuild of large subtrees (replaces the reference's rebuild pthread, ikd_Tree.cpp:201-315)
    struct AsyncRebuild {
        bool pending = false;
        int R = 0, S = 0;
        int64_t stride = 0;
        ikd::DevBuf roots, plan, p4, eroot, stack, forest, visited, split;
    } async;
    // Subtrees with at least this many valid points rebuild on the side stream (0 = never). Since a finished side-stream
    // rebuild is swapped in by the next mutation's own kernels (commit_async), everything above the 512-point in-block
    // builder class goes there (measured on the scan loop: 2049 / 1025 / 513 / 257 -> 0.421 / 0.394 / 0.392 / 0.392 ms).
    int async_min = 513;
    cudaEvent_t main_ev = nullptr;
    // The side stream's adoption kernel rewrites size / invalid of live ancestors; range searches (which read them in
    // their count pass) wait for this event while it is outstanding.
    cudaEvent_t adopt_ev = nullptr;
    bool adopt_in_flight = false;
    // point-id numbering: bumped by ikd_compact_ids (ids handed out before are void afterwards)
    int64_t id_epoch = 0;
    // mapped pinned buffers of the micro-batch kNN path (queries in, results out, no copy-engine call, no stream sync)
    static constexpr int MICRO_MAX_Q = 256;
    void* micro_host = nullptr;
    void* micro_dev = nullptr;
    size_t micro_bytes = 0;
    // last search result (device) for the two-phase protocol
    ikd::DevBuf b_search_ids;
    ikd::DevBuf b_range_pool;  // chunk pool of the single-pass range search
    ikd::DevBuf b_range_ord, b_range_tmp;  // longest-first query order (keys / values, sort scratch)
    int64_t search_total = 0;
    // removed-point log (acquire_removed_points)
    ikd::DevBuf b_removed;
    int64_t removed_n = 0, removed_cap = 0;

    // stats
    ikd_stats stats{};
    bool count_visits = false;
    bool time_kernels = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    // optional rebuild timing (ikd_set_rebuild_timing): event pairs around inline (0), side-stream (1) and whole-tree (2) rebuilds
    bool time_rebuilds = false;
    struct RebuildTiming { cudaEvent_t a, b; int kind; int64_t points; };
    std::vector<RebuildTiming> rebuild_events;
    int64_t launches_total = 0;  // kernels launched by this library (all kinds), for bench.py's gpu_launches
    // optional phase timing of the update path (env IKD_PHASES=1): events between named marks
    bool phase_on = false;
    std::vector<std::pair<const char*, cudaEvent_t>> phase_marks;
    std::vector<std::pair<std::string, std::pair<double, long>>> phase_acc;
    ikd::DevBuf b_visits;

    // pinned staging for small D2H reads
    void* pin = nullptr;
    size_t pin_bytes = 0;
    // big pinned staging for e2e H2D/D2H
    void* pin_io = nullptr;
    size_t pin_io_bytes = 0;
};

namespace ikd {
void phase_mark(ikd_tree* t, const char* name);   // ikd_capi.cu
void phase_flush(ikd_tree* t);
#define IKD_PHASE(t, name) do { if ((t)->phase_on) ikd::phase_mark((t), (name)); } while (0)

// ---- implemented in ikd_build.cu -----------------------------------------------------------------
// Build R balanced subtrees level by level from M float4 points (xyz + pid bits) already on the device.
// max_seg = largest segment size (decides the number of levels).
int forest_build(ikd_tree* t, const float4* p4, int M, const ForestDev& f, int max_seg, cudaStream_t s);
// Whole-tree build from device float4 points: resets the pool, root at slot 1.
int full_build(ikd_tree* t, const float4* p4, int M, cudaStream_t s);
int ensure_pool(ikd_tree* t, size_t slots, bool preserve);
int sync_header(ikd_tree* t);       // read the device header into t->hdr (waits for the stream's earlier work)
// read up to ~4 KB from one or two device locations once everything enqueued on t->stream so far is done
int fetch_small(ikd_tree* t, void* host0, const void* dev0, size_t bytes0, void* host1 = nullptr, const void* dev1 = nullptr,
                size_t bytes1 = 0);
// Same read without the extra launch: the LAST kernel of a sequence publishes the words itself (device side:
// publish_words() at its very end), the host takes a ticket before launching it and waits on the ticket afterwards.
struct PublishTicket {
    uint32_t* dst = nullptr;            // mapped pinned memory (device view)
    volatile uint32_t* flag = nullptr;  // sequence word (device view)
    uint32_t seq = 0;
};
PublishTicket publish_ticket(ikd_tree* t);
int publish_wait(ikd_tree* t, const PublishTicket& tk, void* host_dst, size_t bytes);
int push_header(ikd_tree* t);       // H2D copy of t->hdr
int ensure_pin(ikd_tree* t, size_t bytes);
int ensure_pin_io(ikd_tree* t, size_t bytes);
int ensure_pid_cap(ikd_tree* t, int64_t n);
// Pack strided host points into device float4; w_mode 0: w = point id (first_id + i), 1: w = 0.
int upload_points_f4(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, float4* dst, int first_id, int w_mode);

// ---- implemented in ikd_knn.cu -------------------------------------------------------------------
int knn_launch(ikd_tree* t, const float4* q_dev, int64_t nq, int k, double max_dist, int32_t* out_idx,
               float* out_d, int32_t* out_cnt, cudaStream_t s, int lane = 0);
int pack_queries(const float* q3_dev, int64_t n, float4* q4_dev, cudaStream_t s);

// ---- implemented in ikd_plane.cu -----------------------------------------------------------------
// Plane fit over the kNN results of the same queries (idx/sqd/cnt as written by knn_launch); outputs on the device.
int plane_fit_launch(ikd_tree* t, const float4* q_dev, int64_t nq, int k, const int32_t* idx, const float* sqd,
                     const int32_t* cnt, float max_kth_sqdist, float thr, float* out_plane, float* out_resid,
                     uint8_t* out_valid, cudaStream_t s);

// ---- implemented in ikd_range.cu -----------------------------------------------------------------
int box_search_launch(ikd_tree* t, const float* boxes_dev, int64_t nb, int64_t* offsets_host);
int radius_search_launch(ikd_tree* t, const float4* cr_dev, int64_t nq, int64_t* offsets_host);
int box_add_launch(ikd_tree* t, const float* boxes_dev, int64_t nb, int32_t* changed_dev, unsigned int* nchanged_dev,
                   int* err_dev);
int box_delete_launch(ikd_tree* t, const float* boxes_dev, int64_t nb, bool downsample, int32_t* changed_dev,
                      unsigned int* nchanged_dev, unsigned long long* count_dev, int* err_dev, cudaStream_t stream = nullptr);

// ---- implemented in ikd_update.cu ----------------------------------------------------------------
int delete_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb, int* out_deleted);
int add_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb);
int delete_points_impl(ikd_tree* t, const float* xyz, int64_t n, int64_t stride);
int add_points_impl(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, int downsample_on, int* out_added,
                    int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src);
int add_points_dev_impl(ikd_tree* t, const float4* pts_dev, int64_t n, int downsample_on, int* out_added,
                        int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src);
int flatten_impl(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n);
int dump_tree_impl(ikd_tree* t, float* out, int64_t cap, int64_t* out_n);
int rebuild_all(ikd_tree* t);  // whole-tree rebuild (compaction)
void preload_update_kernels();
void preload_range_kernels();
void preload_build_kernels();
void preload_knn_kernels();
int finish_async(ikd_tree* t);  // wait for a side-stream rebuild and swap its result in (no-op when none is pending)
int pool_alloc(void** p, size_t bytes, cudaStream_t s);  // from the library's private stream-ordered pool (ikd_capi.cu)
int delete_boxes_dev_impl(ikd_tree* t, const float* boxes_dev, int64_t nb, int* out_deleted);
int add_boxes_dev_impl(ikd_tree* t, const float* boxes_dev, int64_t nb);
int delete_points_dev_impl(ikd_tree* t, const float4* pts_dev, int64_t n);
int compact_ids_impl(ikd_tree* t, int32_t* old_of_new, int64_t cap_alive, int64_t* out_alive, int32_t* removed_old,
                     int64_t cap_removed, int64_t* out_removed);
int reset_removed_log(ikd_tree* t);
int read_update_stats(ikd_tree* t);  // device-side Add_Points statistics -> t->stats
void rebuild_time_begin(ikd_tree* t, int kind, int64_t points, cudaStream_t s);
void rebuild_time_end(ikd_tree* t, cudaStream_t s);
}  // namespace ikd
