// extern "C" entry points of libikd_b200.so (declared in include/ikd_b200.h) and the host-side plumbing
// they share: error text, grow-only device buffers, pinned staging, the node pool, header mirroring.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>

#include "ikd_host.h"

namespace ikd {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::ensure(size_t need, cudaStream_t s, bool preserve) {
    if (need <= bytes) return IKD_OK;
    size_t nb = std::max(need, bytes + bytes / 2);
    nb = (nb + 255) & ~(size_t)255;
    void* np = nullptr;
    IKD_CUDA(cudaMalloc(&np, nb));
    if (p) {
        if (preserve) IKD_CUDA(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, s));
        IKD_CUDA(cudaStreamSynchronize(s));
        IKD_CUDA(cudaFree(p));
    }
    p = np;
    bytes = nb;
    return IKD_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}

int ensure_pin(ikd_tree* t, size_t bytes) {
    if (bytes <= t->pin_bytes) return IKD_OK;
    if (t->pin) cudaFreeHost(t->pin);
    t->pin = nullptr;
    t->pin_bytes = 0;
    size_t nb = std::max(bytes, (size_t)1 << 16);
    IKD_CUDA(cudaMallocHost(&t->pin, nb));
    t->pin_bytes = nb;
    return IKD_OK;
}

int ensure_pin_io(ikd_tree* t, size_t bytes) {
    if (bytes <= t->pin_io_bytes) return IKD_OK;
    if (t->pin_io) cudaFreeHost(t->pin_io);
    t->pin_io = nullptr;
    t->pin_io_bytes = 0;
    size_t nb = std::max(bytes, (size_t)1 << 20);
    nb += nb / 4;
    IKD_CUDA(cudaMallocHost(&t->pin_io, nb));
    t->pin_io_bytes = nb;
    return IKD_OK;
}

int ensure_pool(ikd_tree* t, size_t slots, bool preserve) {
    if (slots <= t->cap_slots) return IKD_OK;
    if (slots >= ((size_t)1 << 28)) { set_error("node pool limit exceeded (%zu slots)", slots); return IKD_ERR_CAPACITY; }
    size_t ns = slots;
    if (preserve) ns = std::max(slots, t->cap_slots + t->cap_slots / 2);
    ns = (ns + 1023) & ~(size_t)1023;
    SearchRec* nsr = nullptr;
    UpdateRec* nur = nullptr;
    IKD_CUDA(cudaMalloc((void**)&nsr, ns * sizeof(SearchRec)));
    IKD_CUDA(cudaMalloc((void**)&nur, ns * sizeof(UpdateRec)));
    if (t->srec) {
        if (preserve) {
            IKD_CUDA(cudaMemcpyAsync(nsr, t->srec, t->cap_slots * sizeof(SearchRec), cudaMemcpyDeviceToDevice, t->stream));
            IKD_CUDA(cudaMemcpyAsync(nur, t->urec, t->cap_slots * sizeof(UpdateRec), cudaMemcpyDeviceToDevice, t->stream));
        }
        IKD_CUDA(cudaStreamSynchronize(t->stream));
        cudaFree(t->srec);
        cudaFree(t->urec);
    }
    t->srec = nsr;
    t->urec = nur;
    t->cap_slots = ns;
    return IKD_OK;
}

int ensure_pid_cap(ikd_tree* t, int64_t n) {
    if (n <= t->pid_cap) return IKD_OK;
    int64_t nc = std::max<int64_t>(n, t->pid_cap + t->pid_cap / 2);
    nc = std::max<int64_t>(nc, 1 << 16);
    IKD_TRY(t->pid_xyz.ensure((size_t)nc * sizeof(float4), t->stream, true));
    t->pid_cap = (int64_t)(t->pid_xyz.bytes / sizeof(float4));
    return IKD_OK;
}

int sync_header(ikd_tree* t) {
    IKD_CUDA(cudaMemcpyAsync(t->hdr_pin, t->hdr_dev, sizeof(TreeHeader), cudaMemcpyDeviceToHost, t->stream));
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    t->hdr = *t->hdr_pin;
    return IKD_OK;
}

int push_header(ikd_tree* t) {
    *t->hdr_pin = t->hdr;
    IKD_CUDA(cudaMemcpyAsync(t->hdr_dev, t->hdr_pin, sizeof(TreeHeader), cudaMemcpyHostToDevice, t->stream));
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    return IKD_OK;
}

// Pack n strided host points into float4 (xyz + `w`) on the device, through pinned chunks.
// w_mode 0: w = first_id + i (point id bits); 1: w = 0.
static int upload_f4(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, float4* dst, int first_id, int w_mode) {
    const int64_t CH = 1 << 22;  // 4M points = 64 MB per chunk
    int64_t chunk = std::min<int64_t>(n, CH);
    IKD_TRY(ensure_pin_io(t, (size_t)chunk * sizeof(float4) * (n > CH ? 2 : 1)));
    float4* stage[2] = {(float4*)t->pin_io, (float4*)t->pin_io + (n > CH ? chunk : 0)};
    cudaEvent_t ev[2];
    IKD_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    IKD_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int b = 0;
    for (int64_t off = 0; off < n; off += chunk, b ^= 1) {
        int64_t m = std::min(chunk, n - off);
        if (off >= 2 * chunk) IKD_CUDA(cudaEventSynchronize(ev[b]));
        float4* sp = stage[b];
        const char* base = (const char*)xyz + off * stride;
        for (int64_t i = 0; i < m; i++) {
            const float* p = (const float*)(base + i * stride);
            float4 v;
            v.x = p[0]; v.y = p[1]; v.z = p[2];
            int w = w_mode == 0 ? (int)(first_id + off + i) : 0;
            memcpy(&v.w, &w, 4);
            sp[i] = v;
        }
        IKD_CUDA(cudaMemcpyAsync(dst + off, sp, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, t->stream));
        IKD_CUDA(cudaEventRecord(ev[b], t->stream));
    }
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return IKD_OK;
}

int upload_points_f4(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, float4* dst, int first_id, int w_mode) {
    return upload_f4(t, xyz, n, stride, dst, first_id, w_mode);
}

}  // namespace ikd

using namespace ikd;

#define CHECK_T(t)                                            \
    do {                                                      \
        if (!(t)) { set_error("null tree handle"); return IKD_ERR_ARG; } \
        cudaError_t e0_ = cudaSetDevice((t)->device);         \
        if (e0_ != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e0_)); return IKD_ERR_CUDA; } \
    } while (0)

extern "C" {

const char* ikd_last_error(void) { return g_err; }
int ikd_abi_version(void) { return 1; }
long long ikd_launch_count(void) { return ikd::g_launches.load(); }

int ikd_create(ikd_tree** out, int device, float delete_param, float balance_param, float box_length) {
    if (!out) { set_error("out is null"); return IKD_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); libikd_b200 has no CPU fallback", cudaGetErrorString(e));
        return IKD_ERR_CUDA;
    }
    if (device < 0) IKD_CUDA(cudaGetDevice(&device));
    if (device >= ndev) { set_error("device %d out of range", device); return IKD_ERR_ARG; }
    IKD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    IKD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libikd_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return IKD_ERR_CUDA;
    }
    ikd_tree* t = new ikd_tree();
    t->device = device;
    t->delete_param = delete_param;
    t->balance_param = balance_param;
    t->downsample = box_length;
    IKD_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    IKD_CUDA(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
    IKD_CUDA(cudaEventCreateWithFlags(&t->side_done, cudaEventDisableTiming));
    IKD_CUDA(cudaMalloc((void**)&t->hdr_dev, sizeof(TreeHeader)));
    IKD_CUDA(cudaMallocHost((void**)&t->hdr_pin, sizeof(TreeHeader)));
    memset(&t->hdr, 0, sizeof(t->hdr));
    t->hdr.alpha_bal = 0.5f;
    IKD_TRY(push_header(t));
    t->stats.last_knn_visits = -1;
    *out = t;
    return IKD_OK;
}

int ikd_destroy(ikd_tree* t) {
    if (!t) return IKD_OK;
    cudaSetDevice(t->device);
    cudaStreamSynchronize(t->stream);
    cudaStreamSynchronize(t->side);
    DevBuf* bufs[] = {&t->pid_xyz, &t->b_p4, &t->b_keys0, &t->b_keys1, &t->b_cubtmp, &t->b_pos, &t->b_cls, &t->b_scan,
                      &t->b_mpos, &t->b_flag, &t->b_segaxis, &t->b_forest, &t->b_q, &t->b_perm, &t->b_mkeys, &t->b_mkeys2,
                      &t->b_perm2, &t->b_out_idx, &t->b_out_d, &t->b_out_cnt, &t->b_search_ids, &t->b_removed, &t->b_visits};
    for (DevBuf* b : bufs) b->release();
    for (int a = 0; a < 3; a++) { t->b_ord[a].release(); t->b_ord_alt[a].release(); }
    for (auto& b : t->b_misc) b.release();
    for (auto& b : t->u) b.release();
    if (t->srec) cudaFree(t->srec);
    if (t->urec) cudaFree(t->urec);
    if (t->hdr_dev) cudaFree(t->hdr_dev);
    if (t->hdr_pin) cudaFreeHost(t->hdr_pin);
    if (t->pin) cudaFreeHost(t->pin);
    if (t->pin_io) cudaFreeHost(t->pin_io);
    cudaEventDestroy(t->side_done);
    cudaStreamDestroy(t->stream);
    cudaStreamDestroy(t->side);
    delete t;
    return IKD_OK;
}

int ikd_set_delete_param(ikd_tree* t, float v) { CHECK_T(t); t->delete_param = v; return IKD_OK; }
int ikd_set_balance_param(ikd_tree* t, float v) { CHECK_T(t); t->balance_param = v; return IKD_OK; }
int ikd_set_downsample_param(ikd_tree* t, float v) { CHECK_T(t); t->downsample = v; return IKD_OK; }

int ikd_size(ikd_tree* t, int* out) { CHECK_T(t); *out = t->hdr.root_exists ? t->hdr.size : 0; return IKD_OK; }
int ikd_validnum(ikd_tree* t, int* out) {
    CHECK_T(t);
    *out = t->hdr.root_exists ? t->hdr.size - t->hdr.invalid : 0;
    return IKD_OK;
}
int ikd_root_alpha(ikd_tree* t, float* alpha_bal, float* alpha_del) {
    CHECK_T(t);
    *alpha_bal = t->hdr.alpha_bal;
    *alpha_del = t->hdr.alpha_del;
    return IKD_OK;
}
int ikd_tree_range(ikd_tree* t, float* range6) {
    CHECK_T(t);
    for (int i = 0; i < 6; i++) range6[i] = t->hdr.root_exists ? t->hdr.range[i] : 0.f;
    return IKD_OK;
}
int ikd_has_root(ikd_tree* t, int* out) { CHECK_T(t); *out = t->hdr.root_exists; return IKD_OK; }

int ikd_build(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes) {
    CHECK_T(t);
    if (n < 0 || (n > 0 && !xyz) || stride_bytes < 12) { set_error("bad build arguments"); return IKD_ERR_ARG; }
    if (n > 200000000) { set_error("n too large"); return IKD_ERR_ARG; }
    IKD_CUDA(cudaStreamSynchronize(t->side));
    t->next_pid = 0;
    t->removed_n = 0;
    IKD_TRY(ensure_pid_cap(t, n));
    if (n > 0) IKD_TRY(upload_points_f4(t, xyz, n, stride_bytes, t->pid_xyz.as<float4>(), 0, 0));
    t->next_pid = (int)n;
    IKD_TRY(full_build(t, t->pid_xyz.as<float4>(), (int)n, t->stream));
    IKD_TRY(sync_header(t));
    return IKD_OK;
}

int ikd_knn_batch_dev(ikd_tree* t, const void* q_dev_float4, int64_t nq, int k, double max_dist, int32_t* out_idx_dev,
                      float* out_sqdist_dev, int32_t* out_count_dev) {
    CHECK_T(t);
    if (nq < 0 || (nq > 0 && (!q_dev_float4 || !out_idx_dev || !out_sqdist_dev || !out_count_dev))) {
        set_error("bad knn arguments");
        return IKD_ERR_ARG;
    }
    IKD_TRY(knn_launch(t, (const float4*)q_dev_float4, nq, k, max_dist, out_idx_dev, out_sqdist_dev, out_count_dev,
                       t->stream));
    return IKD_OK;
}

int ikd_knn_batch(ikd_tree* t, const float* q, int64_t nq, int64_t stride_bytes, int k, double max_dist,
                  int32_t* out_idx, float* out_sqdist, int32_t* out_count) {
    CHECK_T(t);
    if (nq < 0 || (nq > 0 && (!q || !out_idx || !out_sqdist || !out_count)) || stride_bytes < 12) {
        set_error("bad knn arguments");
        return IKD_ERR_ARG;
    }
    if (k < 1 || k > IKD_MAX_K) { set_error("k=%d out of range [1,%d]", k, IKD_MAX_K); return IKD_ERR_ARG; }
    // chunk so that staging stays bounded
    const int64_t CH = 1 << 24;
    for (int64_t off = 0; off < nq; off += CH) {
        int64_t m = std::min(CH, nq - off);
        IKD_TRY(t->b_q.ensure((size_t)m * sizeof(float4), t->stream));
        IKD_TRY(t->b_out_idx.ensure((size_t)m * k * 4, t->stream));
        IKD_TRY(t->b_out_d.ensure((size_t)m * k * 4, t->stream));
        IKD_TRY(t->b_out_cnt.ensure((size_t)m * 4, t->stream));
        IKD_TRY(upload_points_f4(t, (const float*)((const char*)q + off * stride_bytes), m, stride_bytes,
                                 t->b_q.as<float4>(), 0, 1));
        IKD_TRY(knn_launch(t, t->b_q.as<float4>(), m, k, max_dist, t->b_out_idx.as<int32_t>(), t->b_out_d.as<float>(),
                           t->b_out_cnt.as<int32_t>(), t->stream));
        IKD_CUDA(cudaMemcpyAsync(out_idx + off * k, t->b_out_idx.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, t->stream));
        IKD_CUDA(cudaMemcpyAsync(out_sqdist + off * k, t->b_out_d.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, t->stream));
        IKD_CUDA(cudaMemcpyAsync(out_count + off, t->b_out_cnt.p, (size_t)m * 4, cudaMemcpyDeviceToHost, t->stream));
        IKD_CUDA(cudaStreamSynchronize(t->stream));
    }
    if (t->count_visits && nq > 0) {
        unsigned long long v = 0;
        IKD_CUDA(cudaMemcpy(&v, t->b_visits.p, sizeof(v), cudaMemcpyDeviceToHost));
        t->stats.last_knn_visits = (int64_t)v;
    }
    return IKD_OK;
}

int ikd_get_points(ikd_tree* t, const int32_t* ids, int64_t n, float* out_xyz) {
    CHECK_T(t);
    if (n < 0 || (n > 0 && (!ids || !out_xyz))) { set_error("bad arguments"); return IKD_ERR_ARG; }
    if (n == 0) return IKD_OK;
    // gather on the host from a D2H copy of the id->xyz table segment that is needed
    int32_t lo = INT32_MAX, hi = -1;
    for (int64_t i = 0; i < n; i++) {
        if (ids[i] < 0) continue;
        lo = std::min(lo, ids[i]);
        hi = std::max(hi, ids[i]);
    }
    if (hi >= t->next_pid) { set_error("point id %d out of range", hi); return IKD_ERR_ARG; }
    std::vector<float4> tmp;
    if (hi >= 0) {
        tmp.resize((size_t)(hi - lo + 1));
        IKD_CUDA(cudaStreamSynchronize(t->stream));
        IKD_CUDA(cudaMemcpy(tmp.data(), t->pid_xyz.as<float4>() + lo, tmp.size() * sizeof(float4), cudaMemcpyDeviceToHost));
    }
    for (int64_t i = 0; i < n; i++) {
        if (ids[i] < 0) { out_xyz[3 * i] = out_xyz[3 * i + 1] = out_xyz[3 * i + 2] = NAN; continue; }
        const float4& v = tmp[(size_t)(ids[i] - lo)];
        out_xyz[3 * i] = v.x; out_xyz[3 * i + 1] = v.y; out_xyz[3 * i + 2] = v.z;
    }
    return IKD_OK;
}

int ikd_synchronize(ikd_tree* t) {
    CHECK_T(t);
    IKD_CUDA(cudaStreamSynchronize(t->side));
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    return IKD_OK;
}

int ikd_get_stats(ikd_tree* t, ikd_stats* out) {
    CHECK_T(t);
    if (!out) return IKD_ERR_ARG;
    t->stats.node_slots_used = t->hdr.pool_top;
    t->stats.node_slots_cap = (int64_t)t->cap_slots;
    t->stats.max_depth = t->hdr.max_depth;
    if (t->count_visits && t->b_visits.p) {
        unsigned long long v = 0;
        IKD_CUDA(cudaStreamSynchronize(t->stream));
        IKD_CUDA(cudaMemcpy(&v, t->b_visits.p, sizeof(v), cudaMemcpyDeviceToHost));
        t->stats.last_knn_visits = (int64_t)v;
    }
    *out = t->stats;
    return IKD_OK;
}

int ikd_set_visit_counting(ikd_tree* t, int on) {
    CHECK_T(t);
    t->count_visits = on != 0;
    if (!on) t->stats.last_knn_visits = -1;
    return IKD_OK;
}

int ikd_set_kernel_timing(ikd_tree* t, int on) {
    CHECK_T(t);
    t->time_kernels = on != 0;
    return IKD_OK;
}

int ikd_get_kernel_time(ikd_tree* t, double* out_ms, int64_t* out_launches) {
    CHECK_T(t);
    if (!out_ms || !out_launches) return IKD_ERR_ARG;
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    double ms = 0;
    for (auto& pr : t->timing_events) {
        float e = 0;
        IKD_CUDA(cudaEventElapsedTime(&e, pr.first, pr.second));
        ms += e;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    *out_ms = ms;
    *out_launches = (int64_t)t->timing_events.size();
    t->timing_events.clear();
    return IKD_OK;
}

int ikd_stream(ikd_tree* t, void** out_stream) {
    CHECK_T(t);
    *out_stream = (void*)t->stream;
    return IKD_OK;
}

int ikd_dump_tree(ikd_tree* t, float* out, int64_t cap, int64_t* out_n) {
    CHECK_T(t);
    return dump_tree_impl(t, out, cap, out_n);
}

}  // extern "C"

// ---- replica sync -----------------------------------------------------------------------------------
extern "C" {

static void fill_desc(ikd_tree* t, int64_t slots, int64_t npoints, ikd_replica_desc* d) {
    d->header_dev = t->hdr_dev;  d->header_bytes = (int64_t)sizeof(TreeHeader);
    d->search_dev = t->srec;     d->search_bytes = slots * (int64_t)sizeof(SearchRec);
    d->update_dev = t->urec;     d->update_bytes = slots * (int64_t)sizeof(UpdateRec);
    d->points_dev = t->pid_xyz.p; d->points_bytes = npoints * (int64_t)sizeof(float4);
    d->slots = slots;
    d->npoints = npoints;
}

int ikd_replica_export(ikd_tree* t, ikd_replica_desc* out) {
    CHECK_T(t);
    if (!out) return IKD_ERR_ARG;
    IKD_CUDA(cudaStreamSynchronize(t->side));
    IKD_TRY(sync_header(t));
    t->hdr.next_pid = t->next_pid;
    IKD_TRY(push_header(t));
    fill_desc(t, (int64_t)t->hdr.pool_top, (int64_t)t->next_pid, out);
    return IKD_OK;
}

int ikd_replica_prepare(ikd_tree* t, int64_t slots, int64_t npoints, ikd_replica_desc* out) {
    CHECK_T(t);
    if (!out || slots < 0 || npoints < 0) return IKD_ERR_ARG;
    IKD_CUDA(cudaStreamSynchronize(t->side));
    size_t extra = (size_t)std::max<int64_t>(npoints / 4, 1 << 20);
    IKD_TRY(ensure_pool(t, (size_t)slots + extra, false));
    IKD_TRY(ensure_pid_cap(t, std::max<int64_t>(npoints, 1)));
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    fill_desc(t, slots, npoints, out);
    return IKD_OK;
}

int ikd_replica_commit(ikd_tree* t) {
    CHECK_T(t);
    IKD_CUDA(cudaDeviceSynchronize());
    IKD_TRY(sync_header(t));
    t->next_pid = t->hdr.next_pid;
    t->hdr.pool_cap = (unsigned int)t->cap_slots;
    IKD_TRY(push_header(t));
    return IKD_OK;
}

}  // extern "C"
