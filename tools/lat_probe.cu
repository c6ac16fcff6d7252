// Dependent-access latency probe (B200): pointer chase through a table of 64-byte records with the access flavours the
// update chain uses (ld.cg 16 B, ld.nc 64 B as 2x256-bit, atomicAdd, store + __threadfence + atomic), one thread per
// warp, few warps -- the regime of the refit / descend / mark kernels. Build on the GPU box: nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
struct __align__(64) Rec { unsigned next; unsigned cnt; float pad[14]; };
template <int MODE>
__global__ void chase(Rec* t, unsigned start_stride, int hops, unsigned long long* out, unsigned* sink, unsigned nwarps) {
    if (threadIdx.x & 31) return;
    unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nwarps) return;
    unsigned cur = w * start_stride + 1;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < hops; i++) {
        if (MODE == 0) { uint4 v = __ldcg(reinterpret_cast<const uint4*>(t + cur)); cur = v.x; acc += v.y; }
        if (MODE == 1) { uint4 v = __ldg(reinterpret_cast<const uint4*>(t + cur)); cur = v.x; acc += v.y; }
        if (MODE == 2) { unsigned nx = __ldcg(&t[cur].next); unsigned old = atomicAdd(&t[cur].cnt, 1u); acc += old; cur = nx; }
        if (MODE == 3) {  // store the record, fence, atomic on it, reload (one "pair level" of the refit)
            unsigned nx = __ldcg(&t[cur].next);
            reinterpret_cast<float4*>(t + cur)[1] = make_float4(1.f, 2.f, 3.f, (float)i);
            __threadfence();
            unsigned old = atomicAdd(&t[cur].cnt, 1u);
            uint4 v = __ldcg(reinterpret_cast<const uint4*>(t + cur));
            acc += old + v.y; cur = nx;
        }
        if (MODE == 4) {  // same without the fence
            unsigned nx = __ldcg(&t[cur].next);
            reinterpret_cast<float4*>(t + cur)[1] = make_float4(1.f, 2.f, 3.f, (float)i);
            unsigned old = atomicAdd(&t[cur].cnt, 1u);
            acc += old; cur = nx;
        }
    }
    long long t1 = clock64();
    out[w] = (unsigned long long)(t1 - t0);
    sink[w] = acc + cur;
}
// same-address atomic throughput: every thread appends 1..8 items to ONE list through a global counter (what mark_kernel /
// the delete kernels do with ndirty / nchanged), plain vs. aggregated over the lanes that are active together
template <bool AGG>
__global__ void append(unsigned* counter, unsigned* list, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int items = (i * 2654435761u >> 29) + 1;
    for (int j = 0; j < items; j++) {
        unsigned pos;
        if (AGG) {
            const unsigned m = __activemask();
            const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            pos = base + __popc(m & ((1u << lane) - 1u));
        } else {
            pos = atomicAdd(counter, 1u);
        }
        list[pos] = (unsigned)i;
    }
}
static void append_probe() {
    unsigned *cnt, *list;
    cudaMalloc(&cnt, 4); cudaMalloc(&list, 4 * 8 * 400000);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int n : {5000, 40000, 400000}) {
        for (int agg = 0; agg < 2; agg++) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; rep++) {
                cudaMemset(cnt, 0, 4);
                cudaEventRecord(a);
                if (agg) append<true><<<(n + 255) / 256, 256>>>(cnt, list, n); else append<false><<<(n + 255) / 256, 256>>>(cnt, list, n);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
            }
            unsigned total; cudaMemcpy(&total, cnt, 4, cudaMemcpyDeviceToHost);
            printf("append: %6d threads, %7u items, %s: %.1f us (%.2f ns per item)\n", n, total, agg ? "warp-aggregated" : "one atomic per item", best * 1e3, best * 1e6 / total);
        }
    }
}
// random-gather bandwidth: every thread reads `per` records of BYTES bytes at pseudo-random (independent) positions of a
// table much larger than L2 -- the access pattern of the partially covered nodes of a range search / of kNN visits
template <int BYTES>
__global__ void gather(const uint4* __restrict__ t, size_t nrec, int per, unsigned* sink) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned long long x = i * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int j = 0; j < per; j++) {
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        size_t r = (x * 0x2545F4914F6CDD1Dull >> 11) % nrec;
        const uint4* p = t + r * (BYTES / 16);
#pragma unroll
        for (int q = 0; q < BYTES / 16; q++) { uint4 v = __ldg(p + q); acc += v.x ^ v.w; }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
static void gather_probe() {
    const size_t bytes = (size_t)8 << 30;
    uint4* t; cudaMalloc(&t, bytes); cudaMemset(t, 1, bytes);
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int threads = 148 * 2048, per = 64;
    for (int bytes_per : {16, 32, 64, 128}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(a);
            if (bytes_per == 16) gather<16><<<threads / 256, 256>>>(t, bytes / 16, per, sink);
            if (bytes_per == 32) gather<32><<<threads / 256, 256>>>(t, bytes / 32, per, sink);
            if (bytes_per == 64) gather<64><<<threads / 256, 256>>>(t, bytes / 64, per, sink);
            if (bytes_per == 128) gather<128><<<threads / 256, 256>>>(t, bytes / 128, per, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
        }
        double n = (double)threads * per;
        printf("random gather of %3d-byte records from an 8 GB table: %.2f G records/s, %.0f GB/s useful\n", bytes_per, n / best / 1e6, n * bytes_per / best / 1e6);
    }
    cudaFree(t);
}
int main(int argc, char** argv) {
    gather_probe();
    append_probe();
    size_t n = argc > 1 ? atol(argv[1]) : (1u << 20);  // records (64 B each)
    int hops = 2000;
    std::vector<unsigned> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = (unsigned)i;
    std::mt19937 g(1);
    std::shuffle(perm.begin() + 1, perm.end(), g);
    std::vector<Rec> h(n);
    for (size_t i = 1; i < n; i++) { h[perm[i]].next = perm[i + 1 < n ? i + 1 : 1]; h[perm[i]].cnt = 0; }
    Rec* d; cudaMalloc(&d, n * sizeof(Rec)); cudaMemcpy(d, h.data(), n * sizeof(Rec), cudaMemcpyHostToDevice);
    unsigned long long* out; unsigned* sink; cudaMalloc(&out, 8 * 4096); cudaMalloc(&sink, 4 * 4096);
    const char* names[5] = {"ld.cg 16B", "ld.nc 16B", "ld.cg + atomicAdd", "store+fence+atomic+reload", "store+atomic (no fence)"};
    for (int warps : {1, 32, 592, 4096}) {
        for (int mode = 0; mode < 5; mode++) {
            for (int rep = 0; rep < 2; rep++) {  // rep 0 warms L2 (table <= L2) -- rep 1 is reported
                int blocks = (warps + 3) / 4;
                unsigned stride = (unsigned)((n - 2) / warps);
                if (mode == 0) chase<0><<<blocks, 128>>>(d, stride, hops, out, sink, (unsigned)warps);
                if (mode == 1) chase<1><<<blocks, 128>>>(d, stride, hops, out, sink, (unsigned)warps);
                if (mode == 2) chase<2><<<blocks, 128>>>(d, stride, hops, out, sink, (unsigned)warps);
                if (mode == 3) chase<3><<<blocks, 128>>>(d, stride, hops, out, sink, (unsigned)warps);
                if (mode == 4) chase<4><<<blocks, 128>>>(d, stride, hops, out, sink, (unsigned)warps);
                cudaDeviceSynchronize();
                if (rep == 1) {
                    std::vector<unsigned long long> o(warps);
                    cudaMemcpy(o.data(), out, 8 * warps, cudaMemcpyDeviceToHost);
                    double s = 0; for (auto v : o) s += (double)v;
                    printf("table %zu MB  warps %5d  %-28s %8.1f cycles/hop\n", n * 64 >> 20, warps, names[mode], s / warps / hops);
                }
            }
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
