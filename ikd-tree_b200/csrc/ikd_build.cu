// Level-by-level median-split builder (replaces BuildTree, reference ikd_Tree.cpp:574-622, and the
// Update pull-up it calls at :620 for freshly built nodes).
//
// The reference recursion picks, for a point segment [l,r]: axis = largest fp32 range with the lowest
// axis winning ties (:582-595), node = element of rank mid=(l+r)>>1 along that axis (nth_element,
// :599-613), children = [l,mid-1] and [mid+1,r] (:616-617). The segment boundaries depend only on the
// segment size, so the whole recursion is a fixed positional schedule. We keep three index lists, one
// per axis, each sorted by that axis inside every live segment ("presorted lists" k-d construction):
//   - a segment's AABB is read off the first/last element of the three lists (O(1), no reduction);
//   - the median is the element at position mid of the split-axis list;
//   - the other two lists are stably partitioned around it (flag by point, prefix-sum, scatter), which
//     keeps them sorted inside the two child segments.
// Every level is therefore a handful of streaming passes over the point positions, for all segments
// of all subtrees of a rebuild forest at once. Ties on the split coordinate are resolved by list
// position (stable sort by input order), which is one of the outcomes nth_element may produce.
//
// Two implementations of the same schedule:
//   - global: one kernel per phase per level over all positions (whole-tree Build, big subtrees);
//   - in-block: a subtree of <= 2048 points is built by ONE thread block entirely in shared memory
//     (block radix sort, block scans), all levels inside one launch. Incremental updates rebuild
//     hundreds of tiny subtrees per batch; this keeps that to a single launch per size class.
#include <cub/cub.cuh>
#include <stdlib.h>
#include <time.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>

#include "ikd_host.h"

namespace ikd {

namespace {

constexpr int TPB = 256;
inline int nblk(int64_t n, int tpb = TPB) { return (int)std::max<int64_t>(1, (n + tpb - 1) / tpb); }

constexpr int SMALL_MAX = 2048;  // largest subtree handled by the in-block builder
constexpr int SMALL_MID = 512;   // boundary between the two largest in-block size classes
#ifndef IKD_SMALL_BT
#define IKD_SMALL_BT 256
#endif
#ifndef IKD_SMALL_RADIX_BITS
#define IKD_SMALL_RADIX_BITS 6
#endif
// The 2048-point in-block builder spends most of its time in its three block radix sorts. With 256 threads (8 keys
// per thread) the digit counters of 6-bit passes fit in shared memory (33 KB), so the ~20 key bits that differ
// inside a subtree take 4 passes instead of 5-6 (block size itself made no difference: 256 / 512 / 1024 threads
// measured within 5%).
constexpr int SMALL_BT = IKD_SMALL_BT;  // threads of the 2048-point in-block builder
constexpr int SMALL_RADIX_BITS = IKD_SMALL_RADIX_BITS;

__device__ __forceinline__ int seg_size_of(const ForestDev& F, int r) { return F.seg_begin[r + 1] - F.seg_begin[r]; }

// axis = largest range, lowest axis on ties (ikd_Tree.cpp:594-595)
__device__ __forceinline__ int pick_axis(const float* mn, const float* mx) {
    float rg0 = __fsub_rn(mx[0], mn[0]), rg1 = __fsub_rn(mx[1], mn[1]), rg2 = __fsub_rn(mx[2], mn[2]);
    int axis = 0;
    float best = rg0;
    if (rg1 > best) { axis = 1; best = rg1; }
    if (rg2 > best) { axis = 2; }
    return axis;
}

// Per-subtree constants of the forest description (one global read per block instead of one per node).
struct RootCtx {
    int base, root_slot, root_parent, root_depth;
    int levels;  // levels of the whole subtree (32 - clz(points)): fixes the slot layout of its block
};
// Slot of local heap index h (>= 2) inside a block of a subtree with `levels` levels. Heap (level) order keeps the nodes of
// one level together, so the three lowest levels under a node lie in three different DRAM rows. With IKD_TREELET the levels
// below the root are grouped in threes (the remainder stays in heap order at the top) and every group is stored treelet by
// treelet: a child pair, its two child pairs and their four child pairs -- 14 nodes, 896 contiguous bytes -- so that the
// last levels of a search touch one row instead of three. Each group of levels occupies exactly the slots its levels occupy
// in heap order (a permutation inside the group), children stay an adjacent even-aligned pair, block sizes do not change.
// MEASURED and NOT adopted (default 0; `tools/build_variant.sh treelet -DIKD_TREELET=1`, all parity tests green): 100M-point
// map, 100M queries 57.57 vs 57.78 ms (12.5M queries: 9.08 vs 9.07 ms) -- the kNN kernel is not bound by DRAM row
// activations --, while the range searches get slower (10M points: box 5.91 vs 4.65 ms, radius 4.29 vs 3.51 ms): the
// enumeration of a contained subtree profits from level order, where every level of a subtree is one contiguous run.
#ifndef IKD_TREELET
#define IKD_TREELET 0
#endif
__device__ __forceinline__ uint32_t block_slot(uint32_t h, int levels) {
#if IKD_TREELET
    const int d = 31 - __clz(h);  // depth of h below the block's root (root = 0)
    const int r = (levels - 1) % 3;
    if (d <= r) return h;
    const int rel = d - r - 1;
    const int t = rel / 3, j = rel - 3 * t;
    const int p0 = r + 1 + 3 * t;                 // depth of the treelet's root pair
    const uint32_t a0 = (h >> j) & ~1u;           // first node of that pair (ancestor of h)
    const uint32_t tl = (a0 - (1u << p0)) >> 1;   // treelet number inside the group
    return (1u << p0) + 14u * tl + ((2u << j) - 2u) + (h - (a0 << j));
#else
    (void)levels;
    return h;
#endif
}
__device__ __forceinline__ RootCtx load_root_ctx(const ForestDev& F, int root) {
    RootCtx rc;
    {
        const int npts = F.seg_begin[root + 1] - F.seg_begin[root];
        rc.levels = npts > 0 ? 32 - __clz(npts) : 1;
    }
    rc.base = F.block_base[root];
    rc.root_slot = F.root_slot[root];
    rc.root_parent = F.root_parent[root];
    rc.root_depth = F.root_depth[root];
    return rc;
}

// Write the node of segment [l,r] (n = r-l+1 points, all valid) of a subtree, local heap index h.
// A node writes its own two records, the records / search boxes of children that do NOT exist, and its own box into
// its parent's search record; it never writes anything a child or the parent writes. The nodes of a subtree can
// therefore be emitted in any order and in parallel (the in-block builders emit all of them in one pass at the end).
__device__ __forceinline__ void emit_node(const RootCtx& rc, uint32_t h, int level, int n, int nleft,
                                          const float* mn, const float* mx, int axis, float4 pt,
                                          SearchRec* __restrict__ srec, UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec,
                                          TreeHeader* __restrict__ hdr) {
    const int base = rc.base;
    int slot = (h == 1u) ? rc.root_slot : base + (int)block_slot(h, rc.levels);
    uint32_t cp = (n >= 2) ? ((uint32_t)base + block_slot(2u * h, rc.levels)) >> 1 : 0u;
    int parent;
    if (h == 1u) parent = rc.root_parent;
    else parent = ((h >> 1) == 1u) ? rc.root_slot : base + (int)block_slot(h >> 1, rc.levels);
    const bool has_l = nleft > 0, has_r = n - 1 - nleft > 0;

    SearchRec* sr = srec + slot;
    uint32_t meta = (cp << META_CP_SHIFT) | ((uint32_t)axis << META_AXIS_SHIFT) | (has_l ? META_LEX : 0u) | (has_r ? META_REX : 0u);
    reinterpret_cast<float4*>(sr)[0] = make_float4(pt.x, pt.y, pt.z, __uint_as_float(meta));
    const float pi = CUDART_INF_F, ni = -CUDART_INF_F;
    if (!has_l) { sr->lmin[0] = pi; sr->lmin[1] = pi; sr->lmin[2] = pi; sr->lmax[0] = ni; sr->lmax[1] = ni; sr->lmax[2] = ni; }
    if (!has_r) { sr->rmin[0] = pi; sr->rmin[1] = pi; sr->rmin[2] = pi; sr->rmax[0] = ni; sr->rmax[1] = ni; sr->rmax[2] = ni; }

    UpdateRec u;
    u.bmin[0] = mn[0]; u.bmin[1] = mn[1]; u.bmin[2] = mn[2];
    u.bmax[0] = mx[0]; u.bmax[1] = mx[1]; u.bmax[2] = mx[2];
    u.size = n; u.invalid = 0; u.down_del = 0; u.parent = parent;
    u.pid = __float_as_int(pt.w);
    u.flags = F_EXISTS | ((uint32_t)axis << F_AXIS_SHIFT);
    u.pending = -1;
    u.depth = rc.root_depth + level;
    u.eff_size = n; u.eff_invalid = 0;
    int4* uq = reinterpret_cast<int4*>(urec + slot);
    const int4* us = reinterpret_cast<const int4*>(&u);
    uq[0] = us[0]; uq[1] = us[1]; uq[2] = us[2]; uq[3] = us[3];
    // walk record: own head and id (8-byte store); the leaf words belong to the children (below), a missing child's word
    // is never consulted (its exists bit is off)
    reinterpret_cast<uint2*>(wrec + slot)[0] = make_uint2(walk_head(cp, has_l, has_r, false), (uint32_t)u.pid);

    if (cp && !(has_l && has_r)) {
        // the child slot that stays empty gets a defined record (n >= 2, so at most one child is missing)
        UpdateRec z;
        memset(&z, 0, sizeof(z));
        z.pending = -1;
        const int4* zs = reinterpret_cast<const int4*>(&z);
        int4* cq = reinterpret_cast<int4*>(urec + 2 * cp + (has_l ? 1 : 0));
        cq[0] = zs[0]; cq[1] = zs[1]; cq[2] = zs[2]; cq[3] = zs[3];
    }
    if (h > 1u) {
        // publish into the parent's walk record what this child is: a live leaf (its id) or a subtree
        reinterpret_cast<int*>(wrec + parent)[2 + (int)(h & 1u)] = (n == 1) ? u.pid : W_NOT_LEAF;
        // publish own box into the parent's search record
        float* dst = (h & 1u) ? srec[parent].rmin : srec[parent].lmin;
        dst[0] = mn[0]; dst[1] = mn[1]; dst[2] = mn[2];
        dst[3] = mx[0]; dst[4] = mx[1]; dst[5] = mx[2];
    } else if (parent == 0) {
        // whole-tree root: header
        hdr->root_exists = 1;
        hdr->root_searchable = 1;
        hdr->size = n;
        hdr->invalid = 0;
        hdr->range[0] = mn[0]; hdr->range[1] = mn[1]; hdr->range[2] = mn[2];
        hdr->range[3] = mx[0]; hdr->range[4] = mx[1]; hdr->range[5] = mx[2];
        // Update(), ikd_Tree.cpp:1315-1321 (son = left child, or right if there is none)
        float ab = 0.5f, ad = 0.0f;
        if (n > 3) {
            int son = nleft > 0 ? nleft : (n - 1 - nleft);
            float tb = (float)son / (float)(n - 1);
            ab = ((double)tb >= 0.5 - 1e-6) ? tb : 1.0f - tb;
        }
        hdr->alpha_bal = ab;
        hdr->alpha_del = ad;
    }
}
__device__ __forceinline__ void emit_node(const ForestDev& F, int root, uint32_t h, int level, int n, int nleft,
                                          const float* mn, const float* mx, int axis, float4 pt,
                                          SearchRec* __restrict__ srec, UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec,
                                          TreeHeader* __restrict__ hdr) {
    emit_node(load_root_ctx(F, root), h, level, n, nleft, mn, mx, axis, pt, srec, urec, wrec, hdr);
}

// ================================================================================================
// global builder
// ================================================================================================
__global__ void init_pos_kernel(int M, ForestDev F, int skip_upto, int* __restrict__ posl, int* __restrict__ posr,
                                uint32_t* __restrict__ posh) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    int root = F.elem_root ? F.elem_root[p] : 0;
    int b = F.seg_begin[root], e = F.seg_begin[root + 1];
    if (e - b <= skip_upto) { posl[p] = 1; posr[p] = 0; }  // built by the in-block builder
    else { posl[p] = b; posr[p] = e - 1; }
    posh[p] = 1u;
}

template <typename KeyT>
__global__ void make_keys_kernel(const float4* __restrict__ p4, int M, int axis, const int* __restrict__ elem_root,
                                 KeyT* __restrict__ keys, int* __restrict__ vals) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    float4 v = p4[e];
    float c = axis == 0 ? v.x : (axis == 1 ? v.y : v.z);
    KeyT k = (KeyT)float_order_key(c);
    if constexpr (sizeof(KeyT) == 8) {
        if (elem_root) k |= ((KeyT)(uint32_t)elem_root[e]) << 32;
    }
    keys[e] = k;
    vals[e] = e;
}

struct BuildArrays {
    const float4* p4;
    int M;
    int* ord[3];      // current lists
    int* ord_out[3];  // next lists
    int* posl;
    int* posr;
    uint32_t* posh;
    uint8_t* segaxis;  // by mid position
    uint8_t* flag;     // by element: 0 left, 1 median, 2 right
    uint8_t* cls;      // 3*M: class of position p in list a (3 = stays in place)
    uint32_t* scan;    // 3*M exclusive sum of (cls == 0)
    int* mpos;         // 3*M: position of the median element in list a, indexed by a*M + mid
};

// One thread per position; only the thread sitting on the median position of a live segment works.
__global__ void build_nodes_kernel(BuildArrays A, ForestDev F, int level, SearchRec* __restrict__ srec,
                                   UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec, TreeHeader* __restrict__ hdr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.M) return;
    int l = A.posl[p], r = A.posr[p];
    if (l > r) return;
    int mid = (l + r) >> 1;
    if (p != mid) return;
    uint32_t h = A.posh[p];
    int root = F.elem_root ? F.elem_root[p] : 0;
    // AABB from the list extremes (all points of a fresh segment are valid)
    float mn[3], mx[3];
    {
        float4 a = A.p4[A.ord[0][l]], b = A.p4[A.ord[0][r]];
        mn[0] = a.x; mx[0] = b.x;
        a = A.p4[A.ord[1][l]]; b = A.p4[A.ord[1][r]];
        mn[1] = a.y; mx[1] = b.y;
        a = A.p4[A.ord[2][l]]; b = A.p4[A.ord[2][r]];
        mn[2] = a.z; mx[2] = b.z;
    }
    int axis = pick_axis(mn, mx);
    int n = r - l + 1;
    float4 pt = A.p4[A.ord[axis][mid]];
    A.segaxis[mid] = (uint8_t)axis;
    emit_node(F, root, h, level, n, mid - l, mn, mx, axis, pt, srec, urec, wrec, hdr);
}

// Fused level kernel (replaces build_nodes + flag + class, and the device-wide scan when CHAIN): one thread per
// position. Every thread derives its segment's split axis from the list extremes itself (same loads for the whole
// segment, served by the cache), the thread on the median position emits the node, and the class of an element in
// the two other lists follows from comparing (coordinate key, element index) with the median's -- the lists are
// sorted by exactly that pair (stable radix sort of the keys with ascending element indices), so this is the same
// decision as "position in the split list < mid" without a flag array and without a kernel boundary.
// CHAIN: the exclusive scan of "goes left" per list is done in the same launch (block scan + chained block totals).
constexpr int LV_TPB = 256;
template <bool CHAIN>
__global__ void __launch_bounds__(LV_TPB)
level_kernel(BuildArrays A, ForestDev F, int level, SearchRec* __restrict__ srec, UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec,
             TreeHeader* __restrict__ hdr, unsigned long long* __restrict__ chain01, unsigned long long* __restrict__ chain2) {
    const int p = blockIdx.x * LV_TPB + threadIdx.x;
    uint8_t c[3] = {3, 3, 3};
    if (p < A.M) {
        const int l = A.posl[p], r = A.posr[p];
        if (l <= r) {
            const int mid = (l + r) >> 1;
            float mn[3], mx[3];
            {
                float4 a = A.p4[A.ord[0][l]], b = A.p4[A.ord[0][r]];
                mn[0] = a.x; mx[0] = b.x;
                a = A.p4[A.ord[1][l]]; b = A.p4[A.ord[1][r]];
                mn[1] = a.y; mx[1] = b.y;
                a = A.p4[A.ord[2][l]]; b = A.p4[A.ord[2][r]];
                mn[2] = a.z; mx[2] = b.z;
            }
            const int ax = pick_axis(mn, mx);
            const int em = A.ord[ax][mid];
            const float4 pm = A.p4[em];
            const uint32_t km = float_order_key(ax == 0 ? pm.x : (ax == 1 ? pm.y : pm.z));
            if (p == mid) {
                const int root = F.elem_root ? F.elem_root[p] : 0;
                emit_node(F, root, A.posh[p], level, r - l + 1, mid - l, mn, mx, ax, pm, srec, urec, wrec, hdr);
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (a == ax) continue;
                const int e = A.ord[a][p];
                const float4 pe = A.p4[e];
                const uint32_t ke = float_order_key(ax == 0 ? pe.x : (ax == 1 ? pe.y : pe.z));
                uint8_t cc = (ke < km || (ke == km && e < em)) ? 0 : (e == em ? 1 : 2);
                if (cc == 1) A.mpos[(size_t)a * A.M + mid] = p;
                c[a] = cc;
            }
        }
#pragma unroll
        for (int a = 0; a < 3; a++) A.cls[(size_t)a * A.M + p] = c[a];
    }
    if (CHAIN) {
        typedef cub::BlockScan<unsigned long long, LV_TPB> Scan;
        __shared__ typename Scan::TempStorage tmp;
        // lists 0 and 1 share one 64-bit scan (32 bits each), list 2 has its own
        unsigned long long v01 = (unsigned long long)(c[0] == 0 ? 1u : 0u) | ((unsigned long long)(c[1] == 0 ? 1u : 0u) << 32);
        unsigned long long v2 = c[2] == 0 ? 1ull : 0ull;
        unsigned long long o01, t01, o2, t2;
        Scan(tmp).ExclusiveSum(v01, o01, t01);
        __syncthreads();
        Scan(tmp).ExclusiveSum(v2, o2, t2);
        unsigned long long b01, b2;
        chain_base2(chain01, t01, chain2, t2, &b01, &b2);
        if (p < A.M) {
            const unsigned long long s01 = b01 + o01;
            A.scan[p] = (uint32_t)(s01 & 0xffffffffu);
            A.scan[(size_t)A.M + p] = (uint32_t)(s01 >> 32);
            A.scan[2 * (size_t)A.M + p] = (uint32_t)(b2 + o2);
        }
    }
}

struct IsLeft {
    __host__ __device__ __forceinline__ uint32_t operator()(const uint8_t& c) const { return c == 0 ? 1u : 0u; }
};

__global__ void scatter_kernel(BuildArrays A) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.M) return;
    int l = A.posl[p], r = A.posr[p];
    bool live = l <= r;
    int mid = (l + r) >> 1;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        size_t off = (size_t)a * A.M;
        uint8_t c = A.cls[off + p];
        int e = A.ord[a][p];
        int dest = p;
        if (c != 3) {
            int cntL = (int)(A.scan[off + p] - A.scan[off + l]);
            if (c == 0) dest = l + cntL;
            else if (c == 1) dest = mid;
            else {
                int pm = A.mpos[off + mid];
                dest = mid + 1 + (p - l - cntL) - (pm < p ? 1 : 0);
            }
        }
        A.ord_out[a][dest] = e;
    }
    if (live) {
        uint32_t h = A.posh[p];
        if (p < mid) { A.posr[p] = mid - 1; A.posh[p] = 2u * h; }
        else if (p > mid) { A.posl[p] = mid + 1; A.posh[p] = 2u * h + 1u; }
        else { A.posl[p] = 1; A.posr[p] = 0; }
    }
}

__global__ void forest_depth_kernel(ForestDev F, TreeHeader* hdr) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= F.R) return;
    int n = seg_size_of(F, r);
    if (n <= 0) return;
    int levels = 32 - __clz(n);  // ceil(log2(n+1))
    atomicMax(&hdr->max_depth, F.root_depth[r] + levels - 1);
}

template <typename KeyT>
int presort(ikd_tree* t, const float4* p4, int M, const ForestDev& f, int end_bit, cudaStream_t s) {
    IKD_TRY(t->b_keys0.ensure(sizeof(KeyT) * (size_t)M, s));
    IKD_TRY(t->b_keys1.ensure(sizeof(KeyT) * (size_t)M, s));
    IKD_TRY(t->b_perm.ensure(sizeof(int) * (size_t)M, s));
    size_t tmp = 0;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<KeyT, int>(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, M, 0,
                                                          end_bit, s)));
    IKD_TRY(t->b_cubtmp.ensure(tmp, s));
    for (int a = 0; a < 3; a++) {
        IKD_TRY(t->b_ord[a].ensure(sizeof(int) * (size_t)M, s));
        IKD_TRY(t->b_ord_alt[a].ensure(sizeof(int) * (size_t)M, s));
        IKD_LAUNCH make_keys_kernel<KeyT><<<nblk(M), TPB, 0, s>>>(p4, M, a, f.elem_root, t->b_keys0.as<KeyT>(),
                                                                    t->b_perm.as<int>());
        size_t tb = t->b_cubtmp.bytes;
        IKD_CUDA((cub::DeviceRadixSort::SortPairs<KeyT, int>(t->b_cubtmp.p, tb, t->b_keys0.as<KeyT>(),
                                                              t->b_keys1.as<KeyT>(), t->b_perm.as<int>(),
                                                              t->b_ord[a].as<int>(), M, 0, end_bit, s)));
    }
    return IKD_OK;
}

// Small and medium builds (side-stream rebuilds): ONE radix sort of the 3M (axis, subtree, coordinate) keys instead
// of three sorts of M keys -- at these sizes every sort pass is a latency-bound launch, and 8 launches replace 21.
__global__ void make_keys3_kernel(const float4* __restrict__ p4, int M, const int* __restrict__ elem_root, int root_bits,
                                  unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * M) return;
    int a = i / M, e = i - a * M;
    float4 v = p4[e];
    float c = a == 0 ? v.x : (a == 1 ? v.y : v.z);
    unsigned long long k = (unsigned long long)float_order_key(c);
    if (elem_root) k |= ((unsigned long long)(uint32_t)elem_root[e]) << 32;
    k |= (unsigned long long)a << (32 + root_bits);
    keys[i] = k;
    vals[i] = e;
}

int presort_combined(ikd_tree* t, const float4* p4, int M, const ForestDev& f, int root_bits, cudaStream_t s) {
    const int N3 = 3 * M;
    IKD_TRY(t->b_keys0.ensure(sizeof(unsigned long long) * (size_t)N3, s));
    IKD_TRY(t->b_keys1.ensure(sizeof(unsigned long long) * (size_t)N3, s));
    IKD_TRY(t->b_perm.ensure(sizeof(int) * (size_t)N3, s));
    IKD_TRY(t->b_ord[0].ensure(sizeof(int) * (size_t)N3, s));  // the three lists back to back
    for (int a = 0; a < 3; a++) IKD_TRY(t->b_ord_alt[a].ensure(sizeof(int) * (size_t)M, s));
    const int end_bit = 32 + root_bits + 2;
    size_t tmp = 0;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<unsigned long long, int>(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, N3, 0,
                                                                        end_bit, s)));
    IKD_TRY(t->b_cubtmp.ensure(tmp, s));
    IKD_LAUNCH make_keys3_kernel<<<nblk(N3), TPB, 0, s>>>(p4, M, (f.R > 1) ? f.elem_root : nullptr, root_bits,
                                                         t->b_keys0.as<unsigned long long>(), t->b_perm.as<int>());
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<unsigned long long, int>(t->b_cubtmp.p, tb, t->b_keys0.as<unsigned long long>(),
                                                                        t->b_keys1.as<unsigned long long>(), t->b_perm.as<int>(),
                                                                        t->b_ord[0].as<int>(), N3, 0, end_bit, s)));
    return IKD_OK;
}

template <int NMAX, int BT>
int launch_finish(ikd_tree* t, const float4* p4, const ForestDev& f, int level0, int skip_upto, const int* o0, const int* o1,
                  const int* o2, int* local_id, cudaStream_t s);

int global_build(ikd_tree* t, const float4* p4, int M, const ForestDev& f, int max_seg, int skip_upto, cudaStream_t s) {
    // 1. three lists sorted by (subtree, coordinate); stable w.r.t. element order
    double t0_ = t->phase_on ? (double)clock() / CLOCKS_PER_SEC * 1e3 : 0;
    const bool combined = M <= (1 << 18);
    if (combined) {
        int rb = 0;
        if (f.R > 1 && f.elem_root) { rb = 1; while ((1ll << rb) < f.R) rb++; }
        IKD_TRY(presort_combined(t, p4, M, f, rb, s));
    } else if (f.R == 1 || !f.elem_root) {
        IKD_TRY(presort<uint32_t>(t, p4, M, f, 32, s));
    } else {
        int rb = 1;
        while ((1ll << rb) < f.R) rb++;
        IKD_TRY(presort<uint64_t>(t, p4, M, f, 32 + rb, s));
    }
    if (t->phase_on) fprintf(stderr, "[ikd host] global_build presort (R=%d) cpu %.3f ms\n", f.R, (double)clock() / CLOCKS_PER_SEC * 1e3 - t0_);
    // 2. positional state
    IKD_TRY(t->b_pos.ensure(sizeof(int) * 3 * (size_t)M, s));
    IKD_TRY(t->b_cls.ensure(3 * (size_t)M, s));
    IKD_TRY(t->b_scan.ensure(sizeof(uint32_t) * 3 * (size_t)M, s));
    IKD_TRY(t->b_mpos.ensure(sizeof(int) * 3 * (size_t)M, s));
    IKD_TRY(t->b_flag.ensure((size_t)M * 4, s));  // flags now, element -> local index map in the finish kernel
    IKD_TRY(t->b_segaxis.ensure((size_t)M, s));
    BuildArrays A;
    A.p4 = p4;
    A.M = M;
    for (int a = 0; a < 3; a++) {
        A.ord[a] = combined ? t->b_ord[0].as<int>() + (size_t)a * M : t->b_ord[a].as<int>();
        A.ord_out[a] = t->b_ord_alt[a].as<int>();
    }
    A.posl = t->b_pos.as<int>();
    A.posr = A.posl + M;
    A.posh = reinterpret_cast<uint32_t*>(A.posr + M);
    A.segaxis = t->b_segaxis.as<uint8_t>();
    A.flag = t->b_flag.as<uint8_t>();
    A.cls = t->b_cls.as<uint8_t>();
    A.scan = t->b_scan.as<uint32_t>();
    A.mpos = t->b_mpos.as<int>();
    IKD_LAUNCH init_pos_kernel<<<nblk(M), TPB, 0, s>>>(M, f, skip_upto, A.posl, A.posr, A.posh);

    int levels = 0;
    while ((1ll << levels) < (long long)max_seg + 1) levels++;  // ceil(log2(max_seg+1))
    // global levels only while segments are larger than the in-block builder can take
    int glevels = 0;
    while ((max_seg >> glevels) > SMALL_MAX) glevels++;
    auto it = thrust::make_transform_iterator((const uint8_t*)A.cls, IsLeft());
    size_t tmp = 0;
    IKD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, A.scan, 3 * (int64_t)M, s));
    IKD_TRY(t->b_cubtmp.ensure(tmp, s));
    // small and medium builds (side-stream rebuilds, maps up to 256k points) scan inside the level kernel
    const int nb_lv = nblk(M, LV_TPB);
    const bool chained = nb_lv <= CHAIN_MAX_BLOCKS;
    unsigned long long* chain_mem = nullptr;
    if (chained) {
        DevBuf& cb = (s == t->side) ? t->b_misc[7] : t->b_misc[6];
        size_t bytes = sizeof(unsigned long long) * 2 * (size_t)nb_lv * (size_t)(glevels + 1);
        IKD_TRY(cb.ensure(bytes, s));
        IKD_CUDA(cudaMemsetAsync(cb.p, 0, bytes, s));
        chain_mem = cb.as<unsigned long long>();
    }
    for (int lv = 0; lv < levels; lv++) {
        if (lv == glevels) {
            // every live segment now fits one block: finish all remaining levels in shared memory
            IKD_TRY(t->b_flag.ensure((size_t)M * 4, s));  // reused as the element -> local index map
            IKD_TRY((launch_finish<SMALL_MAX, 1024>(t, p4, f, glevels, skip_upto, A.ord[0], A.ord[1], A.ord[2],
                                                    t->b_flag.as<int>(), s)));
            break;
        }
        if (lv + 1 == levels) {  // last level: every live segment has one point, nothing to split
            IKD_LAUNCH build_nodes_kernel<<<nblk(M), TPB, 0, s>>>(A, f, lv, t->srec, t->urec, t->wrec, t->hdr_dev);
            break;
        }
        if (chained) {
            unsigned long long* ch = chain_mem + (size_t)lv * 2 * nb_lv;
            IKD_LAUNCH level_kernel<true><<<nb_lv, LV_TPB, 0, s>>>(A, f, lv, t->srec, t->urec, t->wrec, t->hdr_dev, ch, ch + nb_lv);
        } else {
            IKD_LAUNCH level_kernel<false><<<nb_lv, LV_TPB, 0, s>>>(A, f, lv, t->srec, t->urec, t->wrec, t->hdr_dev, nullptr, nullptr);
            size_t tb = t->b_cubtmp.bytes;
            IKD_CUDA(cub::DeviceScan::ExclusiveSum(t->b_cubtmp.p, tb, it, A.scan, 3 * (int64_t)M, s));
        }
        IKD_LAUNCH scatter_kernel<<<nblk(M), TPB, 0, s>>>(A);
        for (int a = 0; a < 3; a++) { int* x = A.ord[a]; A.ord[a] = A.ord_out[a]; A.ord_out[a] = x; }
    }
    return IKD_OK;
}

// ================================================================================================
// in-block builder (one thread block per subtree, everything in shared memory)
// ================================================================================================
// one thread per single-point subtree: a leaf (Add_by_point :819-825)
__global__ void leaf_build_kernel(const float4* __restrict__ p4, ForestDev F, SearchRec* __restrict__ srec,
                                  UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec, TreeHeader* __restrict__ hdr) {
    pdl_wait();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= F.R) return;
    int b = F.seg_begin[r];
    const int nseg = F.seg_begin[r + 1] - b;
    if (nseg > 0) atomicMax(&hdr->max_depth, F.root_depth[r] + (32 - __clz(nseg)) - 1);  // depth bound (was forest_depth_kernel)
    if (nseg != 1) return;
    float4 pt = p4[b];
    float mn[3] = {pt.x, pt.y, pt.z};
    int axis = (F.single_axis && F.single_axis[r] >= 0) ? F.single_axis[r] : 0;
    emit_node(F, r, 1u, 0, 1, 0, mn, mn, axis, pt, srec, urec, wrec, hdr);
}

template <int NMAX>
struct SmallSmem {
    float4 pts[NMAX];
    uint16_t ord[2][3][NMAX];
    uint16_t posl[NMAX], posr[NMAX], posh[NMAX];
    uint16_t scan[3][NMAX];
    uint16_t mpos[3][NMAX];
    uint8_t cls[3][NMAX];
    uint8_t segaxis[NMAX];
    int tie[NMAX];  // tie-break rank of an element among equal coordinates: the lists are sorted by (key, tie)
    // deferred node emission: what position p needs to write its node once all levels are done
    uint16_t nl[NMAX], nr[NMAX];  // its segment when it was the median
    uint16_t npt[NMAX];           // the element that became its node
    uint16_t ext[6][NMAX];        // elements holding the segment's min x, max x, min y, max y, min z, max z
};

// Level loop of the in-block builder: the block's segment (n points in S.pts, three sorted lists in S.ord[0])
// is the node with local heap index h0 at depth level0 of subtree `root`.
template <int NMAX, int BT, class Temp>
__device__ __forceinline__ void block_levels(SmallSmem<NMAX>& S, Temp& tmp, int n, const ForestDev& F, int root, uint32_t h0,
                                             int level0, SearchRec* __restrict__ srec, UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec,
                                             TreeHeader* __restrict__ hdr) {
    constexpr int IT = NMAX / BT;
    typedef cub::BlockScan<unsigned long long, BT> Scan;
    const int tid = threadIdx.x;
    int cur = 0;
    const int levels = 32 - __clz(n);
    for (int lv = 0; lv < levels; lv++) {
        // nodes: the median position of every live segment picks the split axis and notes what it needs to write its
        // node later (emitting here would make every level wait for a few threads' global stores)
        for (int p = tid; p < n; p += BT) {
            int l = S.posl[p], r = S.posr[p];
            if (l > r) continue;
            int mid = (l + r) >> 1;
            if (p != mid) continue;
            uint16_t e0 = S.ord[cur][0][l], e1 = S.ord[cur][0][r], e2 = S.ord[cur][1][l], e3 = S.ord[cur][1][r],
                     e4 = S.ord[cur][2][l], e5 = S.ord[cur][2][r];
            float mn[3], mx[3];
            mn[0] = S.pts[e0].x; mx[0] = S.pts[e1].x;
            mn[1] = S.pts[e2].y; mx[1] = S.pts[e3].y;
            mn[2] = S.pts[e4].z; mx[2] = S.pts[e5].z;
            int axis = pick_axis(mn, mx);
            S.segaxis[mid] = (uint8_t)axis;
            S.nl[p] = (uint16_t)l; S.nr[p] = (uint16_t)r;
            S.ext[0][p] = e0; S.ext[1][p] = e1; S.ext[2][p] = e2; S.ext[3][p] = e3; S.ext[4][p] = e4; S.ext[5][p] = e5;
            S.npt[p] = S.ord[cur][axis][mid];
        }
        __syncthreads();
        if (lv + 1 == levels) break;
        // classes + median positions. The lists are sorted by (coordinate key, tie rank), so an element's side of the
        // split follows from comparing that pair with the median's -- the same decision as "position in the split list
        // < mid", without a flag pass over the split list
        for (int p = tid; p < NMAX; p += BT) {
            int l = S.posl[p], r = S.posr[p];
            bool live = p < n && l <= r;
            int mid = (l + r) >> 1;
            int ax = -1, em = 0, tm = 0;
            uint32_t km = 0;
            if (live) {
                ax = S.segaxis[mid];
                em = S.ord[cur][ax][mid];
                float4 pm = S.pts[em];
                km = float_order_key(ax == 0 ? pm.x : (ax == 1 ? pm.y : pm.z));
                tm = S.tie[em];
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
                uint8_t c = 3;
                if (live && a != ax) {
                    int e = S.ord[cur][a][p];
                    float4 pe = S.pts[e];
                    uint32_t ke = float_order_key(ax == 0 ? pe.x : (ax == 1 ? pe.y : pe.z));
                    c = (ke < km || (ke == km && S.tie[e] < tm)) ? 0 : (e == em ? 1 : 2);
                    if (c == 1) S.mpos[a][mid] = (uint16_t)p;
                }
                S.cls[a][p] = c;
            }
        }
        __syncthreads();
        // exclusive counts of "left" per list: one scan of the three counters packed into 64 bits
        {
            unsigned long long ind[IT], out[IT];
#pragma unroll
            for (int j = 0; j < IT; j++) {
                int p = tid * IT + j;
                ind[j] = (unsigned long long)(S.cls[0][p] == 0 ? 1 : 0) | ((unsigned long long)(S.cls[1][p] == 0 ? 1 : 0) << 20) |
                         ((unsigned long long)(S.cls[2][p] == 0 ? 1 : 0) << 40);
            }
            Scan(tmp.scan).ExclusiveSum(ind, out);
#pragma unroll
            for (int j = 0; j < IT; j++) {
                int p = tid * IT + j;
                S.scan[0][p] = (uint16_t)(out[j] & 0xfffff);
                S.scan[1][p] = (uint16_t)((out[j] >> 20) & 0xfffff);
                S.scan[2][p] = (uint16_t)((out[j] >> 40) & 0xfffff);
            }
            __syncthreads();
        }
        // stable partition of the non-split lists, next level's segments
        for (int p = tid; p < n; p += BT) {
            int l = S.posl[p], r = S.posr[p];
            bool live = l <= r;
            int mid = (l + r) >> 1;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                uint8_t c = S.cls[a][p];
                uint16_t e = S.ord[cur][a][p];
                int dest = p;
                if (c != 3) {
                    int cntL = (int)S.scan[a][p] - (int)S.scan[a][l];
                    if (c == 0) dest = l + cntL;
                    else if (c == 1) dest = mid;
                    else dest = mid + 1 + (p - l - cntL) - ((int)S.mpos[a][mid] < p ? 1 : 0);
                }
                S.ord[cur ^ 1][a][dest] = e;
            }
            if (live) {
                uint16_t h = S.posh[p];
                if (p < mid) { S.posr[p] = (uint16_t)(mid - 1); S.posh[p] = (uint16_t)(2 * h); }
                else if (p > mid) { S.posl[p] = (uint16_t)(mid + 1); S.posh[p] = (uint16_t)(2 * h + 1); }
                else { S.posl[p] = 1; S.posr[p] = 0; }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    // emit all nodes of the segment at once (every position was the median of exactly one segment)
    __syncthreads();
    const RootCtx rc = load_root_ctx(F, root);
    for (int p = tid; p < n; p += BT) {
        const int l = S.nl[p], r = S.nr[p];
        float mn[3], mx[3];
        mn[0] = S.pts[S.ext[0][p]].x; mx[0] = S.pts[S.ext[1][p]].x;
        mn[1] = S.pts[S.ext[2][p]].y; mx[1] = S.pts[S.ext[3][p]].y;
        mn[2] = S.pts[S.ext[4][p]].z; mx[2] = S.pts[S.ext[5][p]].z;
        const uint32_t hr = S.posh[p];  // heap index relative to this block's segment root
        const int hd = 31 - __clz(hr);
        const uint32_t h = (h0 << hd) | (hr ^ (1u << hd));
        emit_node(rc, h, level0 + hd, r - l + 1, p - l, mn, mx, (int)S.segaxis[p], S.pts[S.npt[p]], srec, urec, wrec, hdr);
    }
}

template <int NMAX, int BT>
__global__ void __launch_bounds__(BT)
small_build_kernel(const float4* __restrict__ p4, ForestDev F, int nmin, SearchRec* __restrict__ srec,
                   UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec, TreeHeader* __restrict__ hdr) {
    constexpr int IT = NMAX / BT;
    typedef cub::BlockRadixSort<uint32_t, BT, IT, uint16_t, (NMAX > 256 ? SMALL_RADIX_BITS : 4)> Sort;
    typedef cub::BlockScan<unsigned long long, BT> Scan;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem<NMAX>& S = *reinterpret_cast<SmallSmem<NMAX>*>(smem_raw);
    union Temp {
        typename Sort::TempStorage sort;
        typename Scan::TempStorage scan;
    };
    Temp& tmp = *reinterpret_cast<Temp*>(smem_raw + ((sizeof(SmallSmem<NMAX>) + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    pdl_wait();
    for (int root = blockIdx.x; root < F.R; root += gridDim.x) {
        const int beg = F.seg_begin[root];
        const int n = F.seg_begin[root + 1] - beg;
        if (n <= nmin || n > NMAX) continue;  // other size class (uniform across the block)
        __syncthreads();
        for (int i = tid; i < n; i += BT) S.pts[i] = p4[beg + i];
        for (int i = tid; i < NMAX; i += BT) {
            S.posl[i] = i < n ? 0 : 1;
            S.posr[i] = i < n ? (uint16_t)(n - 1) : 0;
            S.posh[i] = 1;
            S.tie[i] = i;
        }
        __syncthreads();
        // three lists sorted by coordinate, stable w.r.t. element order
        if constexpr (NMAX <= 256) {
            // tiny subtrees: rank by counting (n comparisons per element, all operands in shared memory);
            // rank = #{j : key_j < key_i or (key_j == key_i and j < i)} is the stable sorted position
            for (int a = 0; a < 3; a++) {
                for (int i = tid; i < n; i += BT) {
                    float4 v = S.pts[i];
                    uint32_t ki = float_order_key(a == 0 ? v.x : (a == 1 ? v.y : v.z));
                    int rank = 0;
                    for (int j = 0; j < n; j++) {
                        float4 w = S.pts[j];
                        uint32_t kj = float_order_key(a == 0 ? w.x : (a == 1 ? w.y : w.z));
                        rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
                    }
                    S.ord[0][a][rank] = (uint16_t)i;
                }
            }
            __syncthreads();
        } else {
            // The points of one subtree are close together, so their order keys share the leading bits: sort only
            // the bits below the highest bit in which the smallest and the largest key differ (4 bits per pass).
            __shared__ uint32_t kmin[3], kmax[3];
            if (tid < 3) { kmin[tid] = 0xFFFFFFFFu; kmax[tid] = 0u; }
            __syncthreads();
            for (int i = tid; i < n; i += BT) {
                float4 v = S.pts[i];
                uint32_t kx = float_order_key(v.x), ky = float_order_key(v.y), kz = float_order_key(v.z);
                atomicMin(&kmin[0], kx); atomicMax(&kmax[0], kx);
                atomicMin(&kmin[1], ky); atomicMax(&kmax[1], ky);
                atomicMin(&kmin[2], kz); atomicMax(&kmax[2], kz);
            }
            __syncthreads();
            for (int a = 0; a < 3; a++) {
                uint32_t keys[IT];
                uint16_t vals[IT];
#pragma unroll
                for (int j = 0; j < IT; j++) {
                    int i = tid * IT + j;
                    float c = 0.f;
                    if (i < n) { float4 v = S.pts[i]; c = a == 0 ? v.x : (a == 1 ? v.y : v.z); }
                    keys[j] = i < n ? float_order_key(c) : 0xFFFFFFFFu;
                    vals[j] = (uint16_t)i;
                }
                const uint32_t diff = kmin[a] ^ kmax[a];
                const int nbits = diff ? 32 - __clz(diff) : 1;
                Sort(tmp.sort).Sort(keys, vals, 0, nbits);
#pragma unroll
                for (int j = 0; j < IT; j++) S.ord[0][a][tid * IT + j] = vals[j];
                __syncthreads();
            }
        }
        block_levels<NMAX, BT>(S, tmp, n, F, root, 1u, 0, srec, urec, wrec, hdr);
    }
}

// Finish kernel of the global builder: after `level0` global levels every live segment holds <= NMAX points
// and its three lists are already sorted; one block per (subtree, segment index) builds the rest in shared memory.
template <int NMAX, int BT>
__global__ void __launch_bounds__(BT)
finish_build_kernel(const float4* __restrict__ p4, ForestDev F, int level0, int skip_upto, const int* __restrict__ ord0,
                    const int* __restrict__ ord1, const int* __restrict__ ord2, int* __restrict__ local_id,
                    SearchRec* __restrict__ srec, UpdateRec* __restrict__ urec, WalkRec* __restrict__ wrec, TreeHeader* __restrict__ hdr) {
    constexpr int IT = NMAX / BT;
    typedef cub::BlockRadixSort<uint32_t, BT, IT, uint16_t> Sort;
    typedef cub::BlockScan<unsigned long long, BT> Scan;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem<NMAX>& S = *reinterpret_cast<SmallSmem<NMAX>*>(smem_raw);
    union Temp {
        typename Sort::TempStorage sort;
        typename Scan::TempStorage scan;
    };
    Temp& tmp = *reinterpret_cast<Temp*>(smem_raw + ((sizeof(SmallSmem<NMAX>) + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const unsigned int nseg = 1u << level0;
    const unsigned long long total = (unsigned long long)F.R * nseg;
    for (unsigned long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int root = (int)(w >> level0);
        const unsigned int j = (unsigned int)(w & (nseg - 1));
        int l = F.seg_begin[root], r = F.seg_begin[root + 1] - 1;
        if (r - l + 1 <= skip_upto) continue;
        // positional descent: segment j of level `level0` (bits of j from the top: 0 = left, 1 = right)
        for (int b = level0 - 1; b >= 0 && l <= r; b--) {
            int mid = (l + r) >> 1;
            if ((j >> b) & 1u) l = mid + 1; else r = mid - 1;
        }
        if (l > r) continue;
        const int n = r - l + 1;
        __syncthreads();
        for (int i = tid; i < n; i += BT) {
            int e = ord0[l + i];
            S.pts[i] = p4[e];
            local_id[e] = i;
            S.tie[i] = e;  // the global lists are sorted by (key, element index)
            S.ord[0][0][i] = (uint16_t)i;
        }
        for (int i = tid; i < NMAX; i += BT) {
            S.posl[i] = i < n ? 0 : 1;
            S.posr[i] = i < n ? (uint16_t)(n - 1) : 0;
            S.posh[i] = 1;
        }
        __syncthreads();
        for (int i = tid; i < n; i += BT) {
            S.ord[0][1][i] = (uint16_t)local_id[ord1[l + i]];
            S.ord[0][2][i] = (uint16_t)local_id[ord2[l + i]];
        }
        __syncthreads();
        block_levels<NMAX, BT>(S, tmp, n, F, root, nseg + j, level0, srec, urec, wrec, hdr);
    }
}

template <int NMAX, int BT>
int launch_small(ikd_tree* t, const float4* p4, const ForestDev& f, int nmin, cudaStream_t s) {
    typedef cub::BlockRadixSort<uint32_t, BT, NMAX / BT, uint16_t, (NMAX > 256 ? SMALL_RADIX_BITS : 4)> Sort;
    typedef cub::BlockScan<unsigned long long, BT> Scan;
    size_t smem = ((sizeof(SmallSmem<NMAX>) + 15) & ~(size_t)15) +
                  std::max(sizeof(typename Sort::TempStorage), sizeof(typename Scan::TempStorage)) + 16;
    auto kern = small_build_kernel<NMAX, BT>;
    static bool attr_set = false;
    if (!attr_set) {
        IKD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int grid = std::min(f.R, 148 * 16);
    IKD_LAUNCH_PDL(kern, grid, BT, smem, s, p4, f, nmin, t->srec, t->urec, t->wrec, t->hdr_dev);
    return IKD_OK;
}

template <int NMAX, int BT>
int launch_finish(ikd_tree* t, const float4* p4, const ForestDev& f, int level0, int skip_upto, const int* o0, const int* o1,
                  const int* o2, int* local_id, cudaStream_t s) {
    typedef cub::BlockRadixSort<uint32_t, BT, NMAX / BT, uint16_t> Sort;
    typedef cub::BlockScan<unsigned long long, BT> Scan;
    size_t smem = ((sizeof(SmallSmem<NMAX>) + 15) & ~(size_t)15) +
                  std::max(sizeof(typename Sort::TempStorage), sizeof(typename Scan::TempStorage)) + 16;
    auto kern = finish_build_kernel<NMAX, BT>;
    static bool attr_set = false;
    if (!attr_set) {
        IKD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    unsigned long long total = (unsigned long long)f.R << level0;
    int grid = (int)std::min<unsigned long long>(total, 148ull * 64ull);
    IKD_LAUNCH kern<<<grid, BT, smem, s>>>(p4, f, level0, skip_upto, o0, o1, o2, local_id, t->srec, t->urec, t->wrec, t->hdr_dev);
    return IKD_OK;
}

}  // namespace

// Build R balanced subtrees; max_seg = largest segment size or an upper bound of it (host-known).
int forest_build(ikd_tree* t, const float4* p4, int M, const ForestDev& f, int max_seg, cudaStream_t s) {
    if (M <= 0 || f.R <= 0) return IKD_OK;
    bool whole = (f.R == 1 && max_seg > SMALL_MAX);
    if (whole) IKD_LAUNCH forest_depth_kernel<<<nblk(f.R), TPB, 0, s>>>(f, t->hdr_dev);  // otherwise leaf_build_kernel does it
    // The size classes write disjoint subtrees and only read the forest description, so the in-block builders of
    // the larger classes run on two helper streams next to the small ones (and next to the global levels).
    // Size classes: 1 | 2..32 | 33..256 | 257..512 | 513..2048 points. The time of an in-block build is
    // ~2 us per level plus ~3 ns per point and level plus its sorts (measured with %globaltimer), so the common
    // 257..512 class has its own, four times smaller instance instead of riding in the 2048-point one (48 -> ~30 us).
    cudaStream_t sx[3] = {s, s, s};
    const int w = (s == t->side) ? 1 : 0;
    static const int fork_min = getenv("IKD_FORK_MIN") ? atoi(getenv("IKD_FORK_MIN")) : 32;
    // (no fork on the side stream: nothing waits for that rebuild, while every event call costs the host, which is on the
    // caller's critical path, about a microsecond)
    const bool fork = !whole && max_seg > fork_min && w == 0;
    if (fork) {
        sx[0] = t->aux[w][0];
        if (max_seg > 256) sx[1] = t->aux[w][1];
        if (max_seg > SMALL_MID) sx[2] = t->aux[w][2];
        IKD_CUDA(cudaEventRecord(t->aux_fork[w], s));
        for (int i = 0; i < 3; i++)
            if (sx[i] != s) IKD_CUDA(cudaStreamWaitEvent(sx[i], t->aux_fork[w], 0));
    }
    if (!whole) {
        if (max_seg > SMALL_MID) IKD_TRY((launch_small<SMALL_MAX, SMALL_BT>(t, p4, f, SMALL_MID, sx[2])));
        if (max_seg > 256) IKD_TRY((launch_small<SMALL_MID, 256>(t, p4, f, 256, sx[1])));
        if (max_seg > 32) IKD_TRY((launch_small<256, 256>(t, p4, f, 32, sx[0])));
        IKD_LAUNCH_PDL((leaf_build_kernel), nblk(f.R), TPB, 0, s, p4, f, t->srec, t->urec, t->wrec, t->hdr_dev);
        if (max_seg >= 2) IKD_TRY((launch_small<32, 32>(t, p4, f, 1, s)));
    }
    if (max_seg > SMALL_MAX) IKD_TRY(global_build(t, p4, M, f, max_seg, whole ? 0 : SMALL_MAX, s));
    if (fork) {
        for (int i = 0; i < 3; i++) {
            if (sx[i] == s) continue;
            IKD_CUDA(cudaEventRecord(t->aux_ev[w][i], sx[i]));
            IKD_CUDA(cudaStreamWaitEvent(s, t->aux_ev[w][i], 0));
        }
    }
    IKD_CUDA(cudaGetLastError());
    return IKD_OK;
}

namespace {
__global__ void set_single_forest_kernel(int* arr, int M, int root_slot, int base, int parent, int depth) {
    // layout: seg_begin[2], root_slot, block_base, root_parent, root_depth, single_axis
    arr[0] = 0; arr[1] = M; arr[2] = root_slot; arr[3] = base; arr[4] = parent; arr[5] = depth; arr[6] = -1;
}
__global__ void reset_header_kernel(TreeHeader* h, unsigned int pool_top, unsigned int pool_cap, int next_pid) {
    h->root_exists = 0; h->root_searchable = 0; h->size = 0; h->invalid = 0;
    for (int i = 0; i < 6; i++) h->range[i] = 0.f;
    h->alpha_bal = 0.5f; h->alpha_del = 0.f;
    h->pool_top = pool_top; h->pool_cap = pool_cap; h->max_depth = 0; h->next_pid = next_pid;
    h->counter0 = 0; h->counter1 = 0; h->flag0 = 0; h->flag1 = 0;
    for (int i = 0; i < 8; i++) { h->plan[i] = 0; h->plan2[i] = 0; }
}
}  // namespace

int full_build(ikd_tree* t, const float4* p4, int M, cudaStream_t s) {
    int levels = 0;
    while ((1ll << levels) < (long long)M + 1) levels++;
    size_t heap_slots = (size_t)1 << levels;  // slots 1 .. 2^levels-1 in heap order
    if (heap_slots < 2) heap_slots = 2;
    size_t extra = (size_t)M > ((size_t)1 << 20) ? (size_t)M : ((size_t)1 << 20);
    IKD_TRY(ensure_pool(t, heap_slots + extra, false));
    IKD_LAUNCH reset_header_kernel<<<1, 1, 0, s>>>(t->hdr_dev, (unsigned)heap_slots, (unsigned)t->cap_slots, t->next_pid);
    // every slot below pool_top carries defined flags
    IKD_CUDA(cudaMemsetAsync(t->urec, 0, heap_slots * sizeof(UpdateRec), s));
    if (M == 0) return IKD_OK;
    IKD_TRY(t->b_forest.ensure(sizeof(int) * 16, s));
    int* fa = t->b_forest.as<int>();
    IKD_LAUNCH set_single_forest_kernel<<<1, 1, 0, s>>>(fa, M, ROOT_SLOT, 0, 0, 0);
    ForestDev f;
    f.R = 1;
    f.seg_begin = fa; f.root_slot = fa + 2; f.block_base = fa + 3; f.root_parent = fa + 4; f.root_depth = fa + 5;
    f.single_axis = fa + 6; f.elem_root = nullptr;
    return forest_build(t, p4, M, f, M, s);
}

// Load every kernel of this file now (CUDA loads kernels lazily at their first launch, 0.1-0.3 ms each, which
// showed up as milliseconds of extra latency in the first update after Build).
#define IKD_PRELOAD(fn) do { cudaFuncAttributes a_; if (cudaFuncGetAttributes(&a_, fn) != cudaSuccess) cudaGetLastError(); } while (0)
void preload_build_kernels() {
    IKD_PRELOAD(build_nodes_kernel); IKD_PRELOAD((finish_build_kernel<SMALL_MAX, 1024>)); IKD_PRELOAD(forest_depth_kernel);
    IKD_PRELOAD(init_pos_kernel); IKD_PRELOAD(leaf_build_kernel); IKD_PRELOAD(level_kernel<true>); IKD_PRELOAD(level_kernel<false>);
    IKD_PRELOAD(make_keys3_kernel); IKD_PRELOAD(make_keys_kernel<uint32_t>); IKD_PRELOAD(make_keys_kernel<uint64_t>);
    IKD_PRELOAD(scatter_kernel); IKD_PRELOAD((small_build_kernel<32, 32>)); IKD_PRELOAD((small_build_kernel<256, 256>));
    IKD_PRELOAD((small_build_kernel<SMALL_MAX, SMALL_BT>)); IKD_PRELOAD((small_build_kernel<SMALL_MID, 256>));
}
#undef IKD_PRELOAD

}  // namespace ikd
