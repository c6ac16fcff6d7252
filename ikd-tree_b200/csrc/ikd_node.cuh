// Node storage and exact-fp32 helpers shared by every kernel of libikd_b200.
//
// Replaces KD_TREE_NODE (reference ikd_Tree.h:64-86, 136 B, one `new` per point) with two flat
// 64-byte record arrays in HBM, indexed by node slot:
//   SearchRec  - everything a read-only traversal needs for ONE visit in two 32 B sectors:
//                the node's point, its split axis / deleted bit / child-pair index, and the
//                AABBs of BOTH children (so a visit decides both subtrees without touching them).
//   UpdateRec  - everything the mutating kernels need (own AABB, TreeSize, invalid_point_num,
//                down_del_num, parent, point id, flag bits, refit counter).
// Children of a node always sit in one adjacent slot pair (2*cp, 2*cp+1); cp == 0 means "no children".
// Slot 0 is never used, the root is slot 1, so a fresh Build is exactly the implicit heap layout
// (cp(i) == i) and partial rebuilds are heap-ordered blocks hanging off their root slot.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ikd {

// ---- SearchRec.meta bit layout -------------------------------------------------------------------
constexpr uint32_t META_PDEL = 1u;        // point_deleted (ikd_Tree.h:70)
constexpr uint32_t META_AXIS_SHIFT = 1;   // 2 bits division_axis (ikd_Tree.h:66)
// Which child slots hold a node. A traversal that must tell an EMPTY child position from a dead subtree (both have an
// inverted search box) -- the insert descent, Delete_by_point -- reads these instead of the children's UpdateRec flags, so
// it touches nothing but the 16-byte head of SearchRec (the lines the preceding kNN batch has just pulled into L2).
// Written by emit_node and recompute_core, like the rest of the record.
constexpr uint32_t META_LEX = 8u;         // left child exists
constexpr uint32_t META_REX = 16u;        // right child exists
constexpr uint32_t META_CP_SHIFT = 5;     // child pair index (27 bits: 2^28 node slots)
__host__ __device__ __forceinline__ uint32_t meta_cp(uint32_t m) { return m >> META_CP_SHIFT; }
__host__ __device__ __forceinline__ int meta_axis(uint32_t m) { return (m >> META_AXIS_SHIFT) & 3; }

struct __align__(16) SearchRec {  // 64 B
    float x, y, z;
    uint32_t meta;
    // search-effective child boxes: inverted (min=+inf,max=-inf) when the child is absent or its whole
    // subtree is deleted, so traversals never enter it (reference tests tree_deleted at :870, :649).
    float lmin[3], lmax[3];
    float rmin[3], rmax[3];
};
static_assert(sizeof(SearchRec) == 64, "SearchRec must be 64 bytes");

// ---- UpdateRec.flags -----------------------------------------------------------------------------
constexpr uint32_t F_EXISTS = 1u;      // slot holds a node
constexpr uint32_t F_PDEL = 2u;        // point_deleted
constexpr uint32_t F_TDEL = 4u;        // tree_deleted
constexpr uint32_t F_PDS = 8u;         // point_downsample_deleted
constexpr uint32_t F_TDS = 16u;        // tree_downsample_deleted
constexpr uint32_t F_VIOL = 32u;       // Criterion_Check (ikd_Tree.cpp:1090) failed at the last refit
constexpr uint32_t F_ASYNC = 64u;      // a rebuild of this subtree is in flight on the side stream (Rebuild_Ptr, ikd_Tree.h:184)
constexpr uint32_t F_SOLO = 128u;      // refit bookkeeping: this dirty node is the only dirty child of its parent
constexpr uint32_t F_AXIS_SHIFT = 8;   // 2 bits, copy of the split axis

struct __align__(16) UpdateRec {  // 64 B
    float bmin[3], bmax[3];  // node_range_{x,y,z} (Update, ikd_Tree.cpp:1184): box over non-deleted content
    int size;                // TreeSize
    int invalid;             // invalid_point_num
    int down_del;            // down_del_num
    int parent;              // father slot, 0 for the root
    int pid;                 // stable point id
    uint32_t flags;
    int pending;             // refit bookkeeping: -1 clean, else number of dirty children not yet refit
    int depth;               // root = 0
    // size / invalid count this subtree WILL have once the rebuilds already decided below it are done (equal
    // to size / invalid whenever no rebuild is pending); lets one refit pass evaluate Criterion_Check bottom-up
    // the way the reference does (children are rebuilt before the parent is checked, ikd_Tree.cpp:704-707)
    int eff_size, eff_invalid;
};
static_assert(sizeof(UpdateRec) == 64, "UpdateRec must be 64 bytes");

// ---- WalkRec: 16 bytes per slot for walks that only ENUMERATE points (the reference's flatten, ikd_Tree.cpp:1326-1352,
// which Search_by_range / Search_by_radius call for fully contained subtrees).
//   x: child-pair index, which children exist, point_deleted        y: the node's point id
//   z / w: what the LEFT / RIGHT child is, when it is a single node (a leaf): its point id (>= 0), or W_LEAF_DEAD when that
//          leaf is deleted; W_NOT_LEAF when the child is a subtree that has to be walked (or does not exist).
// Half the nodes of a balanced tree are leaves: with their ids kept by the parent a walk never visits them -- one 16-byte
// fetch reports up to three points, and the number of dependent fetch rounds of a range search halves. A range search
// reports a point of a contained subtree at <= 16 bytes of node traffic instead of 128 (SearchRec + the id in UpdateRec).
// Written wherever a node's records are (re)written: emit_node (builds; x / y by the node itself, z / w of the PARENT by
// each child, the same way a child publishes its box into the parent's SearchRec) and recompute_core (the refit touches
// every node an update changed and all its ancestors, and rewrites all four words from the children's UpdateRecs), so
// it needs no maintenance code of its own.
typedef uint4 WalkRec;
constexpr uint32_t W_PDEL = 1u, W_RIGHT = 2u, W_LEFT = 4u;
constexpr uint32_t W_CP_SHIFT = 3;
constexpr int W_NOT_LEAF = -1, W_LEAF_DEAD = -2;
__host__ __device__ __forceinline__ uint32_t walk_head(uint32_t cp, bool has_l, bool has_r, bool pdel) {
    return (cp << W_CP_SHIFT) | (has_l ? W_LEFT : 0u) | (has_r ? W_RIGHT : 0u) | (pdel ? W_PDEL : 0u);
}

// Device-resident tree header (mirrored on the host after every mutating call).
struct TreeHeader {
    int root_exists;     // Root_Node != nullptr
    int root_searchable; // root exists and is not tree_deleted
    int size;            // Root_Node->TreeSize
    int invalid;         // Root_Node->invalid_point_num
    float range[6];      // root AABB min[3], max[3]
    float alpha_bal, alpha_del;
    unsigned int pool_top;   // bump pointer (slots, always even)
    unsigned int pool_cap;
    int max_depth;       // upper bound of node depth
    int next_pid;
    // scratch counters used by kernels (reset by the host wrapper before use)
    unsigned long long counter0;
    unsigned long long counter1;
    int flag0;
    int flag1;
    int plan[8];  // results of the device-side rebuild planner, read back with the header in one copy
    int plan2[8]; // same for the subtrees handed to the side-stream rebuild
};

constexpr int ROOT_SLOT = 1;

// ---- exact fp32 arithmetic (FMA contraction must never happen; the reference binary has no FMA) ---
// calc_dist, ikd_Tree.cpp:1374-1378: (ax-bx)^2 + (ay-by)^2 + (az-bz)^2, left to right.
__device__ __forceinline__ float sq_dist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// calc_box_dist, ikd_Tree.cpp:1381-1391: terms added in the order x-lo, x-hi, y-lo, y-hi, z-lo, z-hi.
// Per axis at most one of the two terms is non-zero for a valid box (min <= max), (q-min)^2 == (min-q)^2 exactly, and
// adding +0 to a non-negative sum is exact, so  t = max(min-q, q-max, 0); d += t*t  yields the reference's bits with six
// instructions per axis instead of eight predicated ones. An inverted box (min=+inf, max=-inf: child absent or entirely
// deleted) gives +inf as before; a NaN query gives 0 as before (both comparisons of the reference are false, fmaxf drops NaN).
__device__ __forceinline__ float box_sq_dist(float qx, float qy, float qz, float minx, float miny, float minz,
                                             float maxx, float maxy, float maxz) {
    float t = fmaxf(fmaxf(__fsub_rn(minx, qx), __fsub_rn(qx, maxx)), 0.0f);
    float d = __fmul_rn(t, t);
    t = fmaxf(fmaxf(__fsub_rn(miny, qy), __fsub_rn(qy, maxy)), 0.0f);
    d = __fadd_rn(d, __fmul_rn(t, t));
    t = fmaxf(fmaxf(__fsub_rn(minz, qz), __fsub_rn(qz, maxz)), 0.0f);
    d = __fadd_rn(d, __fmul_rn(t, t));
    return d;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute
// (IKD_LAUNCH_PDL, ikd_host.h) may be scheduled while the previous kernel of its stream is still draining; it must
// call this before it touches anything that kernel wrote. The wait returns once the previous grid has completed and its
// memory operations are visible, so semantics are those of a plain stream-ordered launch -- what is saved is the launch
// latency of each link of the 20-40-kernel update chains (a no-op when the kernel was launched the ordinary way).
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// One 64-byte record as two 256-bit loads (LDG.E.256, new with sm_100) instead of four 128-bit ones. A tree traversal's
// record fetches are divergent -- every lane reads its own line -- so the L1TEX unit spends one wavefront per lane and
// load instruction; ncu on the 100M-query kNN launch showed that unit at 97% of its throughput with four LDG.128 per
// visit (profiles/r02_knn_large_100M_ncu_details.txt). Halving the instruction count halves the wavefronts.
#ifndef IKD_LDG256
#define IKD_LDG256 1
#endif
struct Rec64 { float4 a, b, c, e; };
__device__ __forceinline__ Rec64 load_rec64_nc(const void* p) {  // read-only path (searches)
    Rec64 r;
#if IKD_LDG256
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];"
                 : "=f"(r.c.x), "=f"(r.c.y), "=f"(r.c.z), "=f"(r.c.w), "=f"(r.e.x), "=f"(r.e.y), "=f"(r.e.z), "=f"(r.e.w)
                 : "l"(p));
#else
    const float4* q = reinterpret_cast<const float4*>(p);
    r.a = __ldg(q); r.b = __ldg(q + 1); r.c = __ldg(q + 2); r.e = __ldg(q + 3);
#endif
    return r;
}
__device__ __forceinline__ Rec64 load_rec64_cg(const void* p) {  // records that kernels of the same launch modify (L2 only)
    Rec64 r;
#if IKD_LDG256
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p) : "memory");
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];"
                 : "=f"(r.c.x), "=f"(r.c.y), "=f"(r.c.z), "=f"(r.c.w), "=f"(r.e.x), "=f"(r.e.y), "=f"(r.e.z), "=f"(r.e.w)
                 : "l"(p) : "memory");
#else
    const float4* q = reinterpret_cast<const float4*>(p);
    r.a = __ldcg(q); r.b = __ldcg(q + 1); r.c = __ldcg(q + 2); r.e = __ldcg(q + 3);
#endif
    return r;
}

// order-preserving float -> uint32 map for radix sorting by `a < b` on floats
__host__ __device__ __forceinline__ uint32_t float_order_key(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Single-pass chained scan across the blocks of ONE launch (replaces single-block scan kernels, whose 20-60 us were
// pure latency): every block publishes the total of its items in part[blockIdx.x] and then adds up the totals of
// the blocks before it, waiting for those that have not published yet (blocks are dispatched in index order, so
// the ones waited for are always running). `part` must be zero before the launch; bit 63 marks "published".
constexpr int CHAIN_MAX_BLOCKS = 1024;
__device__ __forceinline__ unsigned long long chain_base(unsigned long long* part, unsigned long long my_total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, nb = blockIdx.x;
    if (tid == 0) {
        __threadfence();
        atomicExch(&part[nb], my_total | (1ull << 63));
    }
    unsigned long long sum = 0;
    for (int j = tid; j < nb; j += blockDim.x) {
        unsigned long long v;
        do { v = *reinterpret_cast<volatile unsigned long long*>(&part[j]); } while (!(v >> 63));
        sum += v & ~(1ull << 63);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) s_warp[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = 0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) b += s_warp[w];
        s_base = b;
    }
    __syncthreads();
    return s_base;
}

// Called by every thread of ONE block at the end of a kernel: copy `nwords` 32-bit words to mapped host memory, fence,
// then raise the sequence word the host polls (fetch_small / publish_wait in ikd_capi.cu).
__device__ __forceinline__ void publish_words(const void* src, int nwords, uint32_t* dst, volatile uint32_t* flag, uint32_t seq,
                                              const void* src1 = nullptr, int nwords1 = 0) {
    __syncthreads();  // everything this block wrote to `src` is visible to its threads
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = __ldcg(s + i);
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src1);
    for (int i = threadIdx.x; i < nwords1; i += blockDim.x) dst[nwords + i] = __ldcg(s1 + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;
}

// two independent chains at once (same launch)
__device__ __forceinline__ void chain_base2(unsigned long long* partA, unsigned long long totalA, unsigned long long* partB,
                                            unsigned long long totalB, unsigned long long* baseA,
                                            unsigned long long* baseB) {
    __shared__ unsigned long long s_w[2][32];
    __shared__ unsigned long long s_b[2];
    const int tid = threadIdx.x, nb = blockIdx.x;
    if (tid == 0) {
        atomicExch(&partA[nb], totalA | (1ull << 63));
        atomicExch(&partB[nb], totalB | (1ull << 63));
    }
    unsigned long long sa = 0, sb = 0;
    for (int j = tid; j < nb; j += blockDim.x) {
        unsigned long long v;
        do { v = *reinterpret_cast<volatile unsigned long long*>(&partA[j]); } while (!(v >> 63));
        sa += v & ~(1ull << 63);
        do { v = *reinterpret_cast<volatile unsigned long long*>(&partB[j]); } while (!(v >> 63));
        sb += v & ~(1ull << 63);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_down_sync(0xffffffffu, sa, o);
        sb += __shfl_down_sync(0xffffffffu, sb, o);
    }
    if ((tid & 31) == 0) { s_w[0][tid >> 5] = sa; s_w[1][tid >> 5] = sb; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) { a += s_w[0][w]; b += s_w[1][w]; }
        s_b[0] = a; s_b[1] = b;
    }
    __syncthreads();
    *baseA = s_b[0];
    *baseB = s_b[1];
}

}  // namespace ikd
