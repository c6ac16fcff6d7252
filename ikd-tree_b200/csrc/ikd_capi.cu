// extern "C" entry points of libikd_b200.so (declared in include/ikd_b200.h) and the host-side plumbing
// they share: error text, grow-only device buffers, pinned staging, the node pool, header mirroring.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>

#include "ikd_host.h"

namespace ikd {

static thread_local char g_err[512] = "";
static bool g_trace_alloc = getenv("IKD_PHASES") && atoi(getenv("IKD_PHASES")) != 0;
static double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
std::atomic<long long> g_launches{0};
bool g_use_pdl = !(getenv("IKD_NO_PDL") && atoi(getenv("IKD_NO_PDL")) != 0);

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Grow-only buffer from the device's stream-ordered memory pool (cudaMallocAsync). Plain cudaMalloc/cudaFree
// cost 2-100 ms each on this platform and showed up as the p99 latency of a streaming map whose batch sizes
// creep upwards; pool allocations are microseconds once the pool has grown (its release threshold is set to
// "never" in ikd_create) and the old buffer is freed in stream order, so growth needs no synchronisation.
// The library allocates its scratch from a PRIVATE stream-ordered pool per device (release threshold "never"), so the
// host application's default pool keeps its own settings.
static cudaMemPool_t g_pool[64] = {};
static int library_pool(cudaMemPool_t* out) {
    int dev = 0;
    IKD_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return IKD_ERR_ARG; }
    if (!g_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        IKD_CUDA(cudaMemPoolCreate(&g_pool[dev], &props));
        uint64_t never = UINT64_MAX;
        IKD_CUDA(cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &never));
    }
    *out = g_pool[dev];
    return IKD_OK;
}
int pool_alloc(void** p, size_t bytes, cudaStream_t s) {
    cudaMemPool_t pool;
    IKD_TRY(library_pool(&pool));
    IKD_CUDA(cudaMallocFromPoolAsync(p, bytes, pool, s));
    return IKD_OK;
}

int DevBuf::ensure(size_t need, cudaStream_t s, bool preserve) {
    if (need <= bytes) return IKD_OK;
    size_t nb = std::max(need, bytes * 2);
    nb = (nb + 255) & ~(size_t)255;
    void* np = nullptr;
    double t0 = g_trace_alloc ? now_ms() : 0;
    IKD_TRY(pool_alloc(&np, nb, s));
    if (p) {
        if (preserve) IKD_CUDA(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, s));
        IKD_CUDA(cudaFreeAsync(p, s));
    }
    p = np;
    bytes = nb;
    if (g_trace_alloc && now_ms() - t0 > 0.05) fprintf(stderr, "[ikd alloc] DevBuf grow to %zu bytes took %.2f ms\n", nb, now_ms() - t0);
    return IKD_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}

void phase_mark(ikd_tree* t, const char* name) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, t->stream);
    t->phase_marks.emplace_back(name, e);
}

void phase_flush(ikd_tree* t) {
    if (t->phase_marks.empty()) return;
    cudaStreamSynchronize(t->stream);
    for (size_t i = 0; i + 1 < t->phase_marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t->phase_marks[i].second, t->phase_marks[i + 1].second);
        std::string nm = t->phase_marks[i].first;
        if (nm == "end") continue;
        bool found = false;
        for (auto& a : t->phase_acc)
            if (a.first == nm) { a.second.first += ms; a.second.second++; found = true; break; }
        if (!found) t->phase_acc.push_back({nm, {ms, 1}});
    }
    for (auto& m : t->phase_marks) cudaEventDestroy(m.second);
    t->phase_marks.clear();
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
    size_t nb = std::max(std::max(bytes, t->pin_io_bytes * 2), (size_t)4 << 20);
    t->pin_io_bytes = 0;
    double t0 = g_trace_alloc ? now_ms() : 0;
    IKD_CUDA(cudaMallocHost(&t->pin_io, nb));
    if (g_trace_alloc) fprintf(stderr, "[ikd alloc] pinned io buffer %zu bytes took %.2f ms\n", nb, now_ms() - t0);
    t->pin_io_bytes = nb;
    return IKD_OK;
}

int ensure_pool(ikd_tree* t, size_t slots, bool preserve) {
    if (slots <= t->cap_slots) return IKD_OK;
    if (slots >= ((size_t)1 << 28)) { set_error("node pool limit exceeded (%zu slots)", slots); return IKD_ERR_CAPACITY; }
    // Growth is geometric whatever the caller (a streaming map that grows through whole-tree rebuilds crossed a power of
    // two of points every few hundred scans and paid a 6 ms reallocation each time it asked for exactly what it needed)
    size_t ns = slots;
    if (t->cap_slots) ns = std::max(slots, t->cap_slots + t->cap_slots / 2);
    ns = (ns + 1023) & ~(size_t)1023;
    SearchRec* nsr = nullptr;
    UpdateRec* nur = nullptr;
    WalkRec* nwr = nullptr;
    double t0 = g_trace_alloc ? now_ms() : 0;
    // Stream-ordered allocation from the library's pool, old arrays freed in stream order: no device-wide
    // synchronisation (cudaFree) and no driver allocation once the pool has grown. This call sits inside whole-tree
    // rebuilds of a growing map; with cudaMalloc / cudaFree it showed up as a 3-24 ms update spike (configs[4], scan 62).
    // (Above 1 GB -- Build of a large map, replica preparation: not latency-critical calls -- plain cudaMalloc is used: growing
    // the stream-ordered pool maps physical memory at ~10 ms per GB, measured as +0.35 s on a 34 GB replica.)
    cudaStream_t s = t->stream;
    const bool big = ns * (sizeof(SearchRec) + sizeof(UpdateRec) + sizeof(WalkRec)) > ((size_t)1 << 30);
    if (big) {
        IKD_CUDA(cudaMalloc((void**)&nsr, ns * sizeof(SearchRec)));
        IKD_CUDA(cudaMalloc((void**)&nur, ns * sizeof(UpdateRec)));
        IKD_CUDA(cudaMalloc((void**)&nwr, ns * sizeof(WalkRec)));
    } else {
        IKD_TRY(pool_alloc((void**)&nsr, ns * sizeof(SearchRec), s));
        IKD_TRY(pool_alloc((void**)&nur, ns * sizeof(UpdateRec), s));
        IKD_TRY(pool_alloc((void**)&nwr, ns * sizeof(WalkRec), s));
    }
    if (t->srec) {
        // a side-stream rebuild in flight still reads / writes the old arrays: it ends first (rare: growth is geometric)
        IKD_CUDA(cudaStreamSynchronize(t->side));
        if (preserve) {
            IKD_CUDA(cudaMemcpyAsync(nsr, t->srec, t->cap_slots * sizeof(SearchRec), cudaMemcpyDeviceToDevice, s));
            IKD_CUDA(cudaMemcpyAsync(nur, t->urec, t->cap_slots * sizeof(UpdateRec), cudaMemcpyDeviceToDevice, s));
            IKD_CUDA(cudaMemcpyAsync(nwr, t->wrec, t->cap_slots * sizeof(WalkRec), cudaMemcpyDeviceToDevice, s));
        }
        // searches of the host pipeline run on lane streams; every call that can reach this point has ordered the tree's
        // stream behind them (knn_host_batch), so freeing in the order of that stream is safe
        if (t->pool_from_malloc) {
            IKD_CUDA(cudaStreamSynchronize(s));
            cudaFree(t->srec); cudaFree(t->urec); cudaFree(t->wrec);
        } else {
            IKD_CUDA(cudaFreeAsync(t->srec, s));
            IKD_CUDA(cudaFreeAsync(t->urec, s));
            IKD_CUDA(cudaFreeAsync(t->wrec, s));
        }
    }
    t->pool_from_malloc = big;
    t->srec = nsr;
    t->urec = nur;
    t->wrec = nwr;
    t->cap_slots = ns;
    if (g_trace_alloc) fprintf(stderr, "[ikd alloc] node pool -> %zu slots took %.2f ms\n", ns, now_ms() - t0);
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

// ---- small device -> host reads without a copy engine + stream synchronisation --------------------------------
// The update path needs a handful of counters on the host between kernels (to size the next launches). A
// cudaMemcpyAsync + cudaStreamSynchronize pair leaves the GPU idle for ~17 us per read (measured, four reads per
// Add_Points); here a one-block kernel stores the words into mapped pinned memory, fences, and bumps a sequence
// word the host spins on.
__global__ void publish_kernel(const uint32_t* __restrict__ src0, int n0, const uint32_t* __restrict__ src1, int n1,
                               uint32_t* __restrict__ dst, volatile uint32_t* flag, uint32_t seq) {
    for (int i = threadIdx.x; i < n0; i += blockDim.x) dst[i] = __ldcg(src0 + i);
    for (int i = threadIdx.x; i < n1; i += blockDim.x) dst[n0 + i] = __ldcg(src1 + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;
}

int fetch_small(ikd_tree* t, void* host0, const void* dev0, size_t bytes0, void* host1, const void* dev1, size_t bytes1) {
    const int n0 = (int)((bytes0 + 3) / 4), n1 = (int)((bytes1 + 3) / 4);
    if ((size_t)(n0 + n1) * 4 > ikd_tree::MAPPED_BYTES - 64) { set_error("fetch_small: request too large"); return IKD_ERR_INTERNAL; }
    const uint32_t seq = ++t->map_seq;
    volatile uint32_t* flag_h = reinterpret_cast<volatile uint32_t*>((char*)t->map_host + ikd_tree::MAPPED_BYTES - 64);
    uint32_t* flag_d = reinterpret_cast<uint32_t*>((char*)t->map_dev + ikd_tree::MAPPED_BYTES - 64);
    IKD_LAUNCH publish_kernel<<<1, 64, 0, t->stream>>>((const uint32_t*)dev0, n0, (const uint32_t*)dev1, n1,
                                                      (uint32_t*)t->map_dev, flag_d, seq);
    unsigned long long spins = 0;
    while (*flag_h != seq) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0xffff) == 0) {  // a failed launch / sticky error must not hang the caller
            cudaError_t e = cudaStreamQuery(t->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                set_error("fetch_small: %s", cudaGetErrorString(e));
                return IKD_ERR_CUDA;
            }
            if (e == cudaSuccess && *flag_h != seq) {  // stream drained but the flag is not visible yet: settle it
                IKD_CUDA(cudaStreamSynchronize(t->stream));
                if (*flag_h != seq) { set_error("fetch_small: publish kernel did not run"); return IKD_ERR_INTERNAL; }
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(host0, t->map_host, bytes0);
    if (bytes1) memcpy(host1, (char*)t->map_host + 4 * (size_t)n0, bytes1);
    return IKD_OK;
}

PublishTicket publish_ticket(ikd_tree* t) {
    PublishTicket tk;
    tk.dst = (uint32_t*)t->map_dev;
    tk.flag = reinterpret_cast<volatile uint32_t*>((char*)t->map_dev + ikd_tree::MAPPED_BYTES - 64);
    tk.seq = ++t->map_seq;
    return tk;
}

int publish_wait(ikd_tree* t, const PublishTicket& tk, void* host_dst, size_t bytes) {
    volatile uint32_t* flag_h = reinterpret_cast<volatile uint32_t*>((char*)t->map_host + ikd_tree::MAPPED_BYTES - 64);
    unsigned long long spins = 0;
    while (*flag_h != tk.seq) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0xffff) == 0) {
            cudaError_t e = cudaStreamQuery(t->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                set_error("publish_wait: %s", cudaGetErrorString(e));
                return IKD_ERR_CUDA;
            }
            if (e == cudaSuccess && *flag_h != tk.seq) {
                IKD_CUDA(cudaStreamSynchronize(t->stream));
                if (*flag_h != tk.seq) { set_error("publish_wait: publishing kernel did not run"); return IKD_ERR_INTERNAL; }
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(host_dst, t->map_host, bytes);
    return IKD_OK;
}

int sync_header(ikd_tree* t) {
    IKD_TRY(fetch_small(t, &t->hdr, t->hdr_dev, sizeof(TreeHeader)));
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
// device pointer of a pinned (page-locked, mapped) host buffer, or nullptr for pageable memory
static void* mapped_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// strided xyz (possibly in mapped host memory, read over PCIe) -> float4 on the device
__global__ void pack_strided_kernel(const char* __restrict__ src, int64_t stride, int n, float4* __restrict__ dst, int first_id,
                                    int w_mode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(src + (int64_t)i * stride);
    dst[i] = make_float4(p[0], p[1], p[2], __int_as_float(w_mode == 0 ? first_id + i : 0));
}

static int upload_f4(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, float4* dst, int first_id, int w_mode) {
    const int64_t CH = 1 << 22;  // 4M points = 64 MB per chunk
    if (n <= (1 << 18)) {
        // scan-sized batch. Pinned caller buffer: the pack kernel reads it in place (no staging copy, no H2D call).
        // Pageable: pack into the pinned staging buffer and copy; the staging buffer is not touched again before the
        // caller's next stream wait, which every public entry point performs before it returns.
        if (void* dp = (stride % 4 == 0) ? mapped_device_pointer(xyz) : nullptr) {
            IKD_LAUNCH_PDL((pack_strided_kernel), (int)((n + 255) / 256), 256, 0, t->stream, (const char*)dp, stride, (int)n, dst, first_id, w_mode);
            return IKD_OK;
        }
        IKD_TRY(ensure_pin_io(t, (size_t)n * sizeof(float4)));
        float4* sp = (float4*)t->pin_io;
        for (int64_t i = 0; i < n; i++) {
            const float* p = (const float*)((const char*)xyz + i * stride);
            float4 v;
            v.x = p[0]; v.y = p[1]; v.z = p[2];
            int w = w_mode == 0 ? (int)(first_id + i) : 0;
            memcpy(&v.w, &w, 4);
            sp[i] = v;
        }
        IKD_CUDA(cudaMemcpyAsync(dst, sp, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, t->stream));
        return IKD_OK;
    }
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
int ikd_abi_version(void) { return 2; }
long long ikd_launch_count(void) { return ikd::g_launches.load(); }

__global__ void warm_stream_kernel() {}
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
    {
        cudaMemPool_t pool;  // created on first use; freed blocks stay in it instead of going back to the driver
        IKD_TRY(library_pool(&pool));
    }
    IKD_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    IKD_CUDA(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
    IKD_CUDA(cudaEventCreateWithFlags(&t->side_done, cudaEventDisableTiming));
    IKD_CUDA(cudaEventCreateWithFlags(&t->main_ev, cudaEventDisableTiming));
    IKD_CUDA(cudaEventCreateWithFlags(&t->adopt_ev, cudaEventDisableTiming));
    // Helper streams of the forest builder. They are created AND used once here: the first launch on a new stream
    // allocates its hardware channel, which was measured at 2-15 ms in the middle of the first side-stream rebuild.
    for (int w = 0; w < 2; w++) {
        for (int i = 0; i < 3; i++) {
            IKD_CUDA(cudaStreamCreateWithFlags(&t->aux[w][i], cudaStreamNonBlocking));
            IKD_CUDA(cudaEventCreateWithFlags(&t->aux_ev[w][i], cudaEventDisableTiming));
        }
        IKD_CUDA(cudaEventCreateWithFlags(&t->aux_fork[w], cudaEventDisableTiming));
    }
    {
        static bool preloaded = false;  // once per process
        if (!preloaded) {
            preload_build_kernels(); preload_knn_kernels(); preload_range_kernels(); preload_update_kernels();
            preloaded = true;
        }
    }
    {
        cudaStream_t all[8] = {t->stream, t->side, t->aux[0][0], t->aux[0][1], t->aux[0][2], t->aux[1][0], t->aux[1][1], t->aux[1][2]};
        for (cudaStream_t st : all) IKD_LAUNCH warm_stream_kernel<<<1, 32, 0, st>>>();
        for (cudaStream_t st : all) IKD_CUDA(cudaStreamSynchronize(st));
    }
    if (getenv("IKD_ASYNC_MIN")) t->async_min = atoi(getenv("IKD_ASYNC_MIN"));
    // (the side-stream path is written for subtrees that need the multi-kernel builder; 0 = off, otherwise at least 65 points)
    if (t->async_min > 0 && t->async_min < 65) t->async_min = 65;
    IKD_CUDA(cudaMalloc((void**)&t->hdr_dev, sizeof(TreeHeader)));
    IKD_CUDA(cudaMallocHost((void**)&t->hdr_pin, sizeof(TreeHeader)));
    IKD_CUDA(cudaHostAlloc(&t->map_host, ikd_tree::MAPPED_BYTES, cudaHostAllocMapped));
    memset(t->map_host, 0, ikd_tree::MAPPED_BYTES);
    IKD_CUDA(cudaHostGetDevicePointer(&t->map_dev, t->map_host, 0));
    // micro-batch kNN path: queries (float4) | ids | squared distances | counts, for up to MICRO_MAX_Q queries, k <= 8
    t->micro_bytes = (size_t)ikd_tree::MICRO_MAX_Q * (16 + 8 * 8 + 4) + 256;
    IKD_CUDA(cudaHostAlloc(&t->micro_host, t->micro_bytes, cudaHostAllocMapped));
    memset(t->micro_host, 0, t->micro_bytes);
    IKD_CUDA(cudaHostGetDevicePointer(&t->micro_dev, t->micro_host, 0));
    memset(&t->hdr, 0, sizeof(t->hdr));
    t->hdr.alpha_bal = 0.5f;
    IKD_TRY(push_header(t));
    t->stats.last_knn_visits = -1;
    t->phase_on = getenv("IKD_PHASES") && atoi(getenv("IKD_PHASES")) != 0;
    *out = t;
    return IKD_OK;
}

int ikd_destroy(ikd_tree* t) {
    if (!t) return IKD_OK;
    cudaSetDevice(t->device);
    cudaStreamSynchronize(t->stream);
    if (t->phase_on) {
        phase_flush(t);
        double tot = 0;
        for (auto& a : t->phase_acc) tot += a.second.first;
        fprintf(stderr, "[ikd phases] total %.3f ms\n", tot);
        for (auto& a : t->phase_acc)
            fprintf(stderr, "[ikd phases] %-22s %9.3f ms  n=%6ld  mean %8.2f us  %5.1f%%\n", a.first.c_str(), a.second.first,
                    a.second.second, 1e3 * a.second.first / a.second.second, 100.0 * a.second.first / (tot > 0 ? tot : 1));
    }
    cudaStreamSynchronize(t->side);
    DevBuf* bufs[] = {&t->pid_xyz, &t->b_p4, &t->b_keys0, &t->b_keys1, &t->b_cubtmp, &t->b_pos, &t->b_cls, &t->b_scan,
                      &t->b_mpos, &t->b_flag, &t->b_segaxis, &t->b_forest, &t->b_q, &t->b_perm, &t->b_mkeys, &t->b_mkeys2,
                      &t->b_perm2, &t->b_out_idx, &t->b_out_d, &t->b_out_cnt, &t->b_search_ids, &t->b_range_pool, &t->b_range_ord, &t->b_range_tmp, &t->b_removed, &t->b_visits};
    for (DevBuf* b : bufs) b->release();
    for (int a = 0; a < 3; a++) { t->b_ord[a].release(); t->b_ord_alt[a].release(); }
    for (auto& b : t->b_misc) b.release();
    for (auto& b : t->u) b.release();
    for (auto& L : t->knn_scr) {
        DevBuf* lb[] = {&L.mkeys, &L.mkeys2, &L.perm, &L.perm2, &L.cubtmp, &L.counter, &L.hist, &L.q3, &L.q4, &L.out_idx, &L.out_d, &L.out_cnt, &L.plane, &L.resid, &L.valid};
        for (DevBuf* b : lb) b->release();
        if (L.pin_in) cudaFreeHost(L.pin_in);
        if (L.pin_out) cudaFreeHost(L.pin_out);
        if (L.done) cudaEventDestroy(L.done);
        if (L.stream && L.stream != t->stream) cudaStreamDestroy(L.stream);
    }
    if (t->srec) cudaFree(t->srec);
    if (t->urec) cudaFree(t->urec);
    if (t->wrec) cudaFree(t->wrec);
    if (t->hdr_dev) cudaFree(t->hdr_dev);
    if (t->hdr_pin) cudaFreeHost(t->hdr_pin);
    if (t->map_host) cudaFreeHost(t->map_host);
    if (t->pin) cudaFreeHost(t->pin);
    if (t->pin_io) cudaFreeHost(t->pin_io);
    cudaEventDestroy(t->side_done);
    for (int w = 0; w < 2; w++) {
        for (int i = 0; i < 3; i++) {
            if (t->aux[w][i]) cudaStreamDestroy(t->aux[w][i]);
            if (t->aux_ev[w][i]) cudaEventDestroy(t->aux_ev[w][i]);
        }
        if (t->aux_fork[w]) cudaEventDestroy(t->aux_fork[w]);
    }
    cudaEventDestroy(t->main_ev);
    cudaEventDestroy(t->adopt_ev);
    for (auto& r : t->rebuild_events) { if (r.a) cudaEventDestroy(r.a); if (r.b) cudaEventDestroy(r.b); }
    if (t->micro_host) cudaFreeHost(t->micro_host);
    { DevBuf* ab[] = {&t->async.roots, &t->async.plan, &t->async.p4, &t->async.eroot, &t->async.stack, &t->async.forest, &t->async.visited, &t->async.split}; for (DevBuf* b : ab) b->release(); }
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

static int lane_pin(void** p, size_t* have, size_t need);
int ikd_build(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes) {
    CHECK_T(t);
    if (n < 0 || (n > 0 && !xyz) || stride_bytes < 12) { set_error("bad build arguments"); return IKD_ERR_ARG; }
    if (n > 200000000) { set_error("n too large"); return IKD_ERR_ARG; }
    IKD_CUDA(cudaStreamSynchronize(t->side));
    t->async.pending = false;  // whatever was being rebuilt is replaced by the new tree
    t->adopt_in_flight = false;
    t->next_pid = 0;
    t->removed_n = 0;
    t->id_epoch++;                    // ids restart at 0: everything handed out before is void,
    IKD_TRY(reset_removed_log(t));    // including the removed-point log of the previous tree
    IKD_TRY(ensure_pid_cap(t, n));
    if (n > 0) IKD_TRY(upload_points_f4(t, xyz, n, stride_bytes, t->pid_xyz.as<float4>(), 0, 0));
    t->next_pid = (int)n;
    IKD_TRY(full_build(t, t->pid_xyz.as<float4>(), (int)n, t->stream));
    IKD_TRY(sync_header(t));
    // Reserve scratch for the update path inside the stream-ordered pool now, so that the first Add_Points /
    // Delete_Point_Boxes after Build does not pay the driver's allocation latency for a dozen fresh buffers
    // (measured: 3.4 ms for the first 4-box delete on a 100k-point tree, 0.4 ms afterwards).
    static const bool no_reserve = getenv("IKD_NO_RESERVE") && atoi(getenv("IKD_NO_RESERVE"));
    // (256 MB at least: in a 400-scan streaming run a 16 MB scratch buffer that did not fit the 64 MB reserved before
    // took 18 ms to come from the driver -- the largest update latency of the whole run)
    size_t reserve = std::min<size_t>(std::max<size_t>((size_t)256 << 20, (size_t)64 * t->cap_slots), (size_t)2 << 30);
    if (!no_reserve && n > 0 && reserve > t->pool_reserved) {
        void* p = nullptr;
        IKD_TRY(pool_alloc(&p, reserve, t->stream));
        IKD_CUDA(cudaFreeAsync(p, t->stream));
        t->pool_reserved = reserve;
    }
    // page-locked staging of the host-buffer calls (4-11 ms per cudaMallocHost): here, not inside the first scan
    if (!no_reserve && n > 0) {
        IKD_TRY(ensure_pin_io(t, (size_t)4 << 20));
        for (int ln = 0; ln < 2; ln++) {
            KnnScratch& L = t->knn_scr[ln];
            IKD_TRY(lane_pin(&L.pin_in, &L.pin_in_bytes, (size_t)4 << 20));
            IKD_TRY(lane_pin(&L.pin_out, &L.pin_out_bytes, (size_t)4 << 20));
        }
    }
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

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static int lane_pin(void** p, size_t* have, size_t need) {
    if (need <= *have) return IKD_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    // pinned allocations are slow (milliseconds): grow geometrically so that a slowly growing batch size does not
    // reallocate on every call
    size_t nb = std::max<size_t>(std::max(need, *have * 2), (size_t)4 << 20);
    *have = 0;
    double t0 = g_trace_alloc ? now_ms() : 0;
    IKD_CUDA(cudaMallocHost(p, nb));
    if (g_trace_alloc) fprintf(stderr, "[ikd alloc] pinned lane buffer %zu bytes took %.2f ms\n", nb, now_ms() - t0);
    *have = nb;
    return IKD_OK;
}

// Host-buffer batched kNN. Chunks of up to 1M queries alternate between two lanes (own stream, own device
// and pinned staging), so the H2D copy of chunk i+1, the search of chunk i and the D2H copy of chunk i-1
// overlap. Caller buffers that are already pinned (cudaHostAlloc / cudaHostRegister) are used directly.
// `pf` (optional): run the plane fit behind the search of every chunk and return its outputs instead of the
// distances and counts (out_idx stays optional then).
struct PlaneStage {
    float max_kth_sqdist, thr;
    float* out_plane;
    float* out_resid;
    uint8_t* out_valid;
};
static int knn_host_batch(ikd_tree* t, const float* q, int64_t nq, int64_t stride_bytes, int k, double max_dist,
                          int32_t* out_idx, float* out_sqdist, int32_t* out_count, const PlaneStage* pf) {
    if (nq <= ikd_tree::MICRO_MAX_Q && k <= 8 && !pf) {
        // Micro batch (the combined single-query Nearest_Search calls of a caller's OpenMP loop, include/ikd_Tree.h):
        // latency is everything. Queries are packed into mapped pinned memory by the host, the search kernel reads them
        // and writes its result rows straight back into mapped memory, and the host spins on a sequence word -- no
        // copy-engine call and no stream synchronisation (about two kernel launches of latency per call).
        float4* hq = (float4*)t->micro_host;
        for (int64_t i = 0; i < nq; i++) {
            const float* p = (const float*)((const char*)q + i * stride_bytes);
            hq[i] = make_float4(p[0], p[1], p[2], 0.f);
        }
        char* hb = (char*)t->micro_host;
        char* db = (char*)t->micro_dev;
        const size_t o_idx = (size_t)ikd_tree::MICRO_MAX_Q * 16, o_d = o_idx + (size_t)ikd_tree::MICRO_MAX_Q * 32,
                     o_c = o_d + (size_t)ikd_tree::MICRO_MAX_Q * 32;
        IKD_TRY(knn_launch(t, (const float4*)db, nq, k, max_dist, (int32_t*)(db + o_idx), (float*)(db + o_d),
                           (int32_t*)(db + o_c), t->stream, 0));
        uint32_t dummy = 0;
        IKD_TRY(fetch_small(t, &dummy, t->hdr_dev, 4));  // ordered behind the search kernel: its rows are in host memory
        memcpy(out_idx, hb + o_idx, (size_t)nq * k * 4);
        memcpy(out_sqdist, hb + o_d, (size_t)nq * k * 4);
        memcpy(out_count, hb + o_c, (size_t)nq * 4);
        if (t->count_visits && t->b_visits.p) {
            unsigned long long v = 0;
            IKD_CUDA(cudaMemcpy(&v, t->b_visits.p, sizeof(v), cudaMemcpyDeviceToHost));
            t->stats.last_knn_visits = (int64_t)v;
        }
        return IKD_OK;
    }
    if (nq <= 65536) {
        // Scan-sized batch with pinned caller buffers: the pack kernel reads the queries in place (mapped host memory,
        // coalesced reads over PCIe) and the results go back with three copy-engine transfers straight into the
        // caller's arrays -- no staging copies, no per-call events. (Writing the results from the search kernel into
        // mapped memory was not done: its per-query 20-byte rows would become ~100k small PCIe writes.)
        void* qd = (stride_bytes % 4 == 0) ? mapped_device_pointer(q) : nullptr;
        const bool outs_pinned = pf ? (is_pinned(pf->out_plane) && is_pinned(pf->out_resid) && is_pinned(pf->out_valid) &&
                                       (!out_idx || is_pinned(out_idx)))
                                    : (is_pinned(out_idx) && is_pinned(out_sqdist) && is_pinned(out_count));
        if (qd && outs_pinned) {
            KnnScratch& L = t->knn_scr[0];
            cudaStream_t s = t->stream;
            IKD_TRY(L.q4.ensure((size_t)nq * 16, s));
            IKD_TRY(L.out_idx.ensure((size_t)nq * k * 4, s));
            IKD_TRY(L.out_d.ensure((size_t)nq * k * 4, s));
            IKD_TRY(L.out_cnt.ensure((size_t)nq * 4, s));
            IKD_LAUNCH_PDL((pack_strided_kernel), (int)((nq + 255) / 256), 256, 0, s, (const char*)qd, stride_bytes, (int)nq, L.q4.as<float4>(), 0, 1);
            IKD_TRY(knn_launch(t, L.q4.as<float4>(), nq, k, max_dist, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                               L.out_cnt.as<int32_t>(), s, 0));
            if (pf) {
                IKD_TRY(L.plane.ensure((size_t)nq * 16, s));
                IKD_TRY(L.resid.ensure((size_t)nq * 4, s));
                IKD_TRY(L.valid.ensure((size_t)nq, s));
                IKD_TRY(plane_fit_launch(t, L.q4.as<float4>(), nq, k, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                                         L.out_cnt.as<int32_t>(), pf->max_kth_sqdist, pf->thr, L.plane.as<float>(),
                                         L.resid.as<float>(), L.valid.as<uint8_t>(), s));
                IKD_CUDA(cudaMemcpyAsync(pf->out_plane, L.plane.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, s));
                IKD_CUDA(cudaMemcpyAsync(pf->out_resid, L.resid.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
                IKD_CUDA(cudaMemcpyAsync(pf->out_valid, L.valid.p, (size_t)nq, cudaMemcpyDeviceToHost, s));
                if (out_idx) IKD_CUDA(cudaMemcpyAsync(out_idx, L.out_idx.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
            } else {
                IKD_CUDA(cudaMemcpyAsync(out_idx, L.out_idx.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
                IKD_CUDA(cudaMemcpyAsync(out_sqdist, L.out_d.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
                IKD_CUDA(cudaMemcpyAsync(out_count, L.out_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
            }
            IKD_CUDA(cudaStreamSynchronize(s));
            if (t->count_visits && t->b_visits.p) {
                unsigned long long v = 0;
                IKD_CUDA(cudaMemcpy(&v, t->b_visits.p, sizeof(v), cudaMemcpyDeviceToHost));
                t->stats.last_knn_visits = (int64_t)v;
            }
            return IKD_OK;
        }
    }
    // chunk = one pipeline stage. Larger chunks are denser after Morton ordering (the lanes of a warp share more of their
    // path), which matters on big maps: 100M-point map, 100M queries: 0.70 G q/s with 1M-query chunks, 0.76 G with 4M,
    // 0.79 G with 8M (device-resident single batch: 1.24 G q/s).
    static const int64_t chunk_env = getenv("IKD_KNN_CHUNK") ? atoll(getenv("IKD_KNN_CHUNK")) : 0;
    // measured again in round 2 (100M/100M, e2e): 4M 0.766, 8M 0.786, 16M 0.797, 32M 0.759 G q/s
    const bool big = t->hdr.size >= (16 << 20) && nq >= ((int64_t)32 << 20);  // (on a 1M-point map 1M-query chunks are faster)
    // measured again after the kernel got faster (58 ms per 100M queries, the result copy alone needs ~80 ms): with the ramp
    // below 16M-query chunks 94.8 ms, 8M 89.9 ms, 4M 89.7 ms per 100M-query call
    const int64_t CH = chunk_env > 0 ? chunk_env : (big ? ((int64_t)8 << 20) : ((int64_t)1 << 20));
    const bool in_direct = stride_bytes == 12 && is_pinned(q);
    const bool out_direct = pf ? (is_pinned(pf->out_plane) && is_pinned(pf->out_resid) && is_pinned(pf->out_valid) &&
                                  (!out_idx || is_pinned(out_idx)))
                               : (is_pinned(out_idx) && is_pinned(out_sqdist) && is_pinned(out_count));
    // bytes per query staged in a lane's pinned output buffer
    const size_t out_row = pf ? (size_t)21 + (out_idx ? (size_t)k * 4 : 0) : (size_t)k * 8 + 4;
    static const int max_lanes = getenv("IKD_KNN_LANES") ? std::max(1, std::min((int)ikd_tree::KNN_LANES, atoi(getenv("IKD_KNN_LANES")))) : (int)ikd_tree::KNN_LANES;
    const int nlanes = (int)std::min<int64_t>(max_lanes, (nq + CH - 1) / CH);
    struct EventGuard {  // destroyed on every return path, including the error ones
        cudaEvent_t ev = nullptr;
        ~EventGuard() { if (ev) cudaEventDestroy(ev); }
    } start_guard;
    IKD_CUDA(cudaEventCreateWithFlags(&start_guard.ev, cudaEventDisableTiming));
    const cudaEvent_t start_ev = start_guard.ev;
    IKD_CUDA(cudaEventRecord(start_ev, t->stream));  // searches are ordered after earlier work on the tree
    struct Pending { int64_t off = -1, m = 0; } pend[ikd_tree::KNN_LANES];
    auto drain = [&](int ln) -> int {  // wait for the lane's last chunk and hand its results to the caller
        KnnScratch& L = t->knn_scr[ln];
        if (pend[ln].off < 0) return IKD_OK;
        IKD_CUDA(cudaEventSynchronize(L.done));
        if (!out_direct) {
            int64_t off = pend[ln].off, m = pend[ln].m;
            char* po = (char*)L.pin_out;
            if (pf) {
                memcpy(pf->out_plane + off * 4, po, (size_t)m * 16);
                memcpy(pf->out_resid + off, po + (size_t)m * 16, (size_t)m * 4);
                memcpy(pf->out_valid + off, po + (size_t)m * 20, (size_t)m);
                if (out_idx) memcpy(out_idx + off * k, po + (((size_t)m * 21 + 15) & ~(size_t)15), (size_t)m * k * 4);
            } else {
                memcpy(out_idx + off * k, po, (size_t)m * k * 4);
                memcpy(out_sqdist + off * k, po + (size_t)m * k * 4, (size_t)m * k * 4);
                memcpy(out_count + off, po + (size_t)m * k * 8, (size_t)m * 4);
            }
        }
        pend[ln].off = -1;
        return IKD_OK;
    };
    int ci = 0;
    // Ramp: on a big batch the result copy (8k + 4 bytes per query over PCIe) is what bounds the call, so the first chunks
    // are short (CH/4, CH/2) to get the device-to-host engine going early, and the last full-size chunk is split the same
    // way so that little copying is left once the last search has finished.
    static const bool ramp = !(getenv("IKD_KNN_NO_RAMP") && atoi(getenv("IKD_KNN_NO_RAMP")));
    auto chunk_len = [&](int64_t off) -> int64_t {
        const int64_t left = nq - off;
        if (!ramp || !big || nq < 4 * CH) return std::min(CH, left);
        if (off == 0) return CH / 4;
        if (off == CH / 4) return CH / 2;
        if (left <= CH / 4) return left;
        if (left <= CH) return std::max<int64_t>(left / 2, CH / 4);  // tail: halves
        return CH;
    };
    for (int64_t off = 0, m = 0; off < nq; off += m, ci++) {
        m = chunk_len(off);
        int ln = ci % nlanes;
        KnnScratch& L = t->knn_scr[ln];
        if (!L.stream) {
            if (ln == 0) L.stream = t->stream;
            else IKD_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
            IKD_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
        }
        IKD_TRY(drain(ln));
        cudaStream_t s = L.stream;
        if (ln != 0) IKD_CUDA(cudaStreamWaitEvent(s, start_ev, 0));
        IKD_TRY(L.q3.ensure((size_t)m * 12, s));
        IKD_TRY(L.q4.ensure((size_t)m * 16, s));
        IKD_TRY(L.out_idx.ensure((size_t)m * k * 4, s));
        IKD_TRY(L.out_d.ensure((size_t)m * k * 4, s));
        IKD_TRY(L.out_cnt.ensure((size_t)m * 4, s));
        const float* src = (const float*)((const char*)q + off * stride_bytes);
        if (!in_direct) {
            IKD_TRY(lane_pin(&L.pin_in, &L.pin_in_bytes, (size_t)std::min(CH, nq) * 12));
            float* pi = (float*)L.pin_in;
            if (stride_bytes == 12) memcpy(pi, src, (size_t)m * 12);
            else
                for (int64_t i = 0; i < m; i++) {
                    const float* p = (const float*)((const char*)src + i * stride_bytes);
                    pi[3 * i] = p[0]; pi[3 * i + 1] = p[1]; pi[3 * i + 2] = p[2];
                }
            src = pi;
        }
        IKD_CUDA(cudaMemcpyAsync(L.q3.p, src, (size_t)m * 12, cudaMemcpyHostToDevice, s));
        IKD_TRY(pack_queries(L.q3.as<float>(), m, L.q4.as<float4>(), s));
        IKD_TRY(knn_launch(t, L.q4.as<float4>(), m, k, max_dist, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                           L.out_cnt.as<int32_t>(), s, ln));
        if (pf) {
            IKD_TRY(L.plane.ensure((size_t)m * 16, s));
            IKD_TRY(L.resid.ensure((size_t)m * 4, s));
            IKD_TRY(L.valid.ensure((size_t)m, s));
            IKD_TRY(plane_fit_launch(t, L.q4.as<float4>(), m, k, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                                     L.out_cnt.as<int32_t>(), pf->max_kth_sqdist, pf->thr, L.plane.as<float>(),
                                     L.resid.as<float>(), L.valid.as<uint8_t>(), s));
        }
        if (!out_direct) IKD_TRY(lane_pin(&L.pin_out, &L.pin_out_bytes, (size_t)std::min(CH, nq) * out_row + 64));
        char* po = (char*)L.pin_out;
        if (pf) {
            IKD_CUDA(cudaMemcpyAsync(out_direct ? (void*)(pf->out_plane + off * 4) : (void*)po, L.plane.p, (size_t)m * 16,
                                     cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(out_direct ? (void*)(pf->out_resid + off) : (void*)(po + (size_t)m * 16), L.resid.p,
                                     (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(out_direct ? (void*)(pf->out_valid + off) : (void*)(po + (size_t)m * 20), L.valid.p,
                                     (size_t)m, cudaMemcpyDeviceToHost, s));
            if (out_idx)
                IKD_CUDA(cudaMemcpyAsync(out_direct ? (void*)(out_idx + off * k) : (void*)(po + (((size_t)m * 21 + 15) & ~(size_t)15)),
                                         L.out_idx.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
        } else if (out_direct) {
            IKD_CUDA(cudaMemcpyAsync(out_idx + off * k, L.out_idx.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(out_sqdist + off * k, L.out_d.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(out_count + off, L.out_cnt.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        } else {
            IKD_CUDA(cudaMemcpyAsync(po, L.out_idx.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(po + (size_t)m * k * 4, L.out_d.p, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
            IKD_CUDA(cudaMemcpyAsync(po + (size_t)m * k * 8, L.out_cnt.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        }
        IKD_CUDA(cudaEventRecord(L.done, s));
        pend[ln].off = off;
        pend[ln].m = m;
    }
    for (int ln = 0; ln < nlanes; ln++) IKD_TRY(drain(ln));
    for (int ln = 1; ln < nlanes; ln++) IKD_CUDA(cudaStreamWaitEvent(t->stream, t->knn_scr[ln].done, 0));  // later updates wait for the searches
    if (t->count_visits && t->b_visits.p) {
        unsigned long long v = 0;
        IKD_CUDA(cudaMemcpy(&v, t->b_visits.p, sizeof(v), cudaMemcpyDeviceToHost));
        t->stats.last_knn_visits = (int64_t)v;
    }
    return IKD_OK;
}

int ikd_knn_plane_batch_dev(ikd_tree* t, const void* q_dev_float4, int64_t nq, int k, double max_dist,
                            float max_kth_sqdist, float plane_threshold, float* out_plane_dev, float* out_resid_dev,
                            uint8_t* out_valid_dev) {
    CHECK_T(t);
    if (nq < 0 || (nq > 0 && (!q_dev_float4 || !out_plane_dev || !out_resid_dev || !out_valid_dev))) {
        set_error("bad knn_plane arguments");
        return IKD_ERR_ARG;
    }
    if (k < IKD_PLANE_MIN_K || k > IKD_PLANE_MAX_K) {
        set_error("plane fit needs %d <= k <= %d (got %d)", IKD_PLANE_MIN_K, IKD_PLANE_MAX_K, k);
        return IKD_ERR_ARG;
    }
    if (nq == 0) return IKD_OK;
    KnnScratch& L = t->knn_scr[0];
    cudaStream_t s = t->stream;
    IKD_TRY(L.out_idx.ensure((size_t)nq * k * 4, s));
    IKD_TRY(L.out_d.ensure((size_t)nq * k * 4, s));
    IKD_TRY(L.out_cnt.ensure((size_t)nq * 4, s));
    IKD_TRY(knn_launch(t, (const float4*)q_dev_float4, nq, k, max_dist, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                       L.out_cnt.as<int32_t>(), s, 0));
    IKD_TRY(plane_fit_launch(t, (const float4*)q_dev_float4, nq, k, L.out_idx.as<int32_t>(), L.out_d.as<float>(),
                             L.out_cnt.as<int32_t>(), max_kth_sqdist, plane_threshold, out_plane_dev, out_resid_dev,
                             out_valid_dev, s));
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
    if (nq == 0) return IKD_OK;
    return knn_host_batch(t, q, nq, stride_bytes, k, max_dist, out_idx, out_sqdist, out_count, nullptr);
}

// Host-buffer kNN + plane fit: same chunked, multi-lane pipeline as ikd_knn_batch; per query 12 bytes go up and 21
// come back (plus 4k when the caller wants the neighbour ids) instead of 8k + 4.
int ikd_knn_plane_batch(ikd_tree* t, const float* q, int64_t nq, int64_t stride_bytes, int k, double max_dist,
                        float max_kth_sqdist, float plane_threshold, float* out_plane, float* out_resid,
                        uint8_t* out_valid, int32_t* out_idx) {
    CHECK_T(t);
    if (nq < 0 || (nq > 0 && (!q || !out_plane || !out_resid || !out_valid)) || stride_bytes < 12) {
        set_error("bad knn_plane arguments");
        return IKD_ERR_ARG;
    }
    if (k < IKD_PLANE_MIN_K || k > IKD_PLANE_MAX_K) {
        set_error("plane fit needs %d <= k <= %d (got %d)", IKD_PLANE_MIN_K, IKD_PLANE_MAX_K, k);
        return IKD_ERR_ARG;
    }
    if (nq == 0) return IKD_OK;
    PlaneStage pf = {max_kth_sqdist, plane_threshold, out_plane, out_resid, out_valid};
    return knn_host_batch(t, q, nq, stride_bytes, k, max_dist, out_idx, nullptr, nullptr, &pf);
}

// out[3i..3i+2] = coordinates of point id ids[i] (NaN for a negative id); *bad is set when an id is out of range
__global__ void gather_points_kernel(const int32_t* __restrict__ ids, int64_t n, const float4* __restrict__ pid_xyz, int next_pid,
                                     float* __restrict__ out, int* __restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int id = ids[i];
        float x = CUDART_NAN_F, y = CUDART_NAN_F, z = CUDART_NAN_F;
        if (id >= next_pid) *bad = id;
        else if (id >= 0) { const float4 v = pid_xyz[id]; x = v.x; y = v.y; z = v.z; }
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
}

int ikd_get_points(ikd_tree* t, const int32_t* ids, int64_t n, float* out_xyz) {
    CHECK_T(t);
    if (n < 0 || (n > 0 && (!ids || !out_xyz))) { set_error("bad arguments"); return IKD_ERR_ARG; }
    if (n == 0) return IKD_OK;
    // gather on the device (the id table stays in HBM): ids up, 12 bytes per point back, in chunks of 4M ids
    cudaStream_t s = t->stream;
    const int64_t CH = (int64_t)4 << 20;
    DevBuf& b_ids = t->b_out_idx;
    DevBuf& b_xyz = t->b_out_d;
    DevBuf& b_bad = t->b_out_cnt;
    const int64_t m0 = std::min(n, CH);
    IKD_TRY(b_ids.ensure((size_t)m0 * 4, s));
    IKD_TRY(b_xyz.ensure((size_t)m0 * 12, s));
    IKD_TRY(b_bad.ensure(16, s));
    IKD_CUDA(cudaMemsetAsync(b_bad.p, 0xFF, 4, s));
    for (int64_t off = 0; off < n; off += CH) {
        const int64_t m = std::min(CH, n - off);
        IKD_CUDA(cudaMemcpyAsync(b_ids.p, ids + off, (size_t)m * 4, cudaMemcpyHostToDevice, s));
        IKD_LAUNCH gather_points_kernel<<<(int)std::min<int64_t>((m + 255) / 256, 148 * 16), 256, 0, s>>>(
            b_ids.as<int32_t>(), m, t->pid_xyz.as<float4>(), t->next_pid, b_xyz.as<float>(), b_bad.as<int>());
        IKD_CUDA(cudaMemcpyAsync(out_xyz + 3 * off, b_xyz.p, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
        if (off + CH < n) IKD_CUDA(cudaStreamSynchronize(s));  // the staging buffers are reused by the next chunk
    }
    int bad = -1;
    IKD_CUDA(cudaMemcpyAsync(&bad, b_bad.p, 4, cudaMemcpyDeviceToHost, s));
    IKD_CUDA(cudaStreamSynchronize(s));
    if (bad >= 0) { set_error("point id %d out of range (ids handed out: %d)", bad, t->next_pid); return IKD_ERR_ARG; }
    return IKD_OK;
}

int ikd_next_id(ikd_tree* t, int64_t* out_next_id) {
    CHECK_T(t);
    if (!out_next_id) return IKD_ERR_ARG;
    *out_next_id = t->next_pid;
    return IKD_OK;
}

int ikd_id_epoch(ikd_tree* t, int64_t* out_epoch) {
    CHECK_T(t);
    if (!out_epoch) return IKD_ERR_ARG;
    *out_epoch = t->id_epoch;
    return IKD_OK;
}

int ikd_set_rebuild_timing(ikd_tree* t, int on) {
    CHECK_T(t);
    t->time_rebuilds = on != 0;
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
    if (t->count_visits) IKD_TRY(read_update_stats(t));
    // finished rebuild timings -> stats
    for (size_t i = 0; i < t->rebuild_events.size();) {
        auto& r = t->rebuild_events[i];
        float ms = 0;
        cudaError_t e = r.b ? cudaEventQuery(r.b) : cudaErrorNotReady;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            if (r.kind == 0) { t->stats.rebuild_inline_ms += ms; t->stats.rebuild_inline_n++; }
            else if (r.kind == 1) { t->stats.rebuild_async_ms += ms; t->stats.rebuild_async_n++; }
            else { t->stats.rebuild_full_ms += ms; t->stats.rebuild_full_n++; }
            if (ms > t->stats.rebuild_max_ms) t->stats.rebuild_max_ms = ms;
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
            t->rebuild_events.erase(t->rebuild_events.begin() + i);
        } else {
            cudaGetLastError();
            i++;
        }
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
    d->walk_dev = t->wrec;       d->walk_bytes = slots * (int64_t)sizeof(WalkRec);
    d->points_dev = t->pid_xyz.p; d->points_bytes = npoints * (int64_t)sizeof(float4);
    d->slots = slots;
    d->npoints = npoints;
}

int ikd_replica_export(ikd_tree* t, ikd_replica_desc* out) {
    CHECK_T(t);
    if (!out) return IKD_ERR_ARG;
    IKD_TRY(finish_async(t));
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
