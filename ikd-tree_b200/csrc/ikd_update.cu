// Mutating operations: Add_Points (with voxel downsample), Delete_Points, Delete_Point_Boxes, the
// Update pull-up ("refit"), Criterion_Check + Rebuild, flatten. Reference: ikd_Tree.cpp:414-489,
// :514-556, :625-760, :818-866, :1090-1107, :1184-1352.
//
// The reference mutates one point at a time and repairs the path it walked (Update, Criterion_Check,
// Rebuild on the way back up). Here every public call is a BATCH:
//   1. a kernel applies the whole batch (flag bits set with atomics / new subtrees written),
//      recording the node slots it touched;
//   2. refit: the ancestors of the touched slots are marked dirty and recomputed bottom-up in one
//      kernel (last-arriving child continues to the parent), which is Update() applied once per dirty
//      node instead of once per point, and evaluates Criterion_Check on each of them;
//   3. the topmost violating nodes are rebuilt together by the level-by-level forest builder
//      (ikd_build.cu); their ancestors are refit again; repeat until no node violates the criteria.
// Deletes are eager (the flag is written on every affected node), so there is no Push_Down and
// searches never mutate. Lazy "tree_deleted" still exists as a derived bit and makes searches skip
// dead subtrees through inverted child boxes.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <vector>

#include "ikd_host.h"

namespace ikd {

namespace {

constexpr int TPB = 256;
inline int nblk(int64_t n, int tpb = TPB) { return (int)std::max<int64_t>(1, (n + tpb - 1) / tpb); }

enum {
    U_CHANGED = 0, U_NCHANGED, U_DIRTY, U_START, U_ROOTS, U_RINFO, U_STACK, U_P4, U_EROOT, U_FOREST, U_BOXES, U_ERR,
    U_PTS, U_KEYS, U_KEYS2, U_IDX, U_IDX2, U_GROUP, U_GINFO, U_VOX, U_ALIVE, U_SEL, U_CNT, U_TMP, U_TMP2, U_SURV, U_SRC
};

struct Ctx {
    SearchRec* srec;
    UpdateRec* urec;
    TreeHeader* hdr;
};

__device__ __forceinline__ UpdateRec load_urec_cg(const UpdateRec* p) {
    UpdateRec u;
    const int4* s = reinterpret_cast<const int4*>(p);
    int4* d = reinterpret_cast<int4*>(&u);
    d[0] = __ldcg(s); d[1] = __ldcg(s + 1); d[2] = __ldcg(s + 2); d[3] = __ldcg(s + 3);
    return u;
}
__device__ __forceinline__ void store_urec(UpdateRec* p, const UpdateRec& u) {
    int4* d = reinterpret_cast<int4*>(p);
    const int4* s = reinterpret_cast<const int4*>(&u);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// ================================================================================================
// refit (Update, ikd_Tree.cpp:1184-1323, + Criterion_Check :1090-1107)
// ================================================================================================
__global__ void mark_kernel(Ctx c, const int32_t* __restrict__ changed, const unsigned int* __restrict__ nchanged,
                            int32_t* __restrict__ dirty, unsigned int* __restrict__ ndirty) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *nchanged) return;
    int n = changed[i];
    if (n <= 0) return;
    if (atomicCAS(&c.urec[n].pending, -1, 0) != -1) return;  // already dirty: its marker walks the ancestors
    dirty[atomicAdd(ndirty, 1u)] = n;
    while (true) {
        int p = c.urec[n].parent;
        if (p == 0) break;
        int prev = atomicCAS(&c.urec[p].pending, -1, 0);
        atomicAdd(&c.urec[p].pending, 1);
        if (prev != -1) break;
        dirty[atomicAdd(ndirty, 1u)] = p;
        n = p;
    }
}

__global__ void starters_kernel(Ctx c, const int32_t* __restrict__ dirty, const unsigned int* __restrict__ ndirty,
                                uint8_t* __restrict__ start) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *ndirty) return;
    start[i] = c.urec[dirty[i]].pending == 0 ? 1 : 0;
}

__device__ void recompute_node(Ctx c, int n, float del_param, float bal_param) {
    SearchRec* sr = c.srec + n;
    float4 a = __ldcg(reinterpret_cast<const float4*>(sr));
    uint32_t meta = __float_as_uint(a.w);
    UpdateRec u = load_urec_cg(c.urec + n);
    const uint32_t cp = meta_cp(meta);
    const bool pdel = (u.flags & F_PDEL) != 0, pds = (u.flags & F_PDS) != 0;
    int size = 1, invalid = pdel ? 1 : 0, dd = pds ? 1 : 0;
    bool tds = pds, tdel = pdel;
    bool cex[2] = {false, false}, ctdel[2] = {false, false};
    float cmn[2][3], cmx[2][3];
    int csize[2] = {0, 0};
    if (cp) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
            UpdateRec ch = load_urec_cg(c.urec + 2 * cp + s);
            if (ch.flags & F_EXISTS) {
                cex[s] = true;
                csize[s] = ch.size;
                size += ch.size; invalid += ch.invalid; dd += ch.down_del;
                tds = tds && (ch.flags & F_TDS);
                ctdel[s] = (ch.flags & F_TDEL) != 0;
                tdel = tdel && ctdel[s];
#pragma unroll
                for (int k = 0; k < 3; k++) { cmn[s][k] = ch.bmin[k]; cmx[s][k] = ch.bmax[k]; }
            }
        }
    }
    // range over non-deleted content, or over everything when nothing / all is deleted (:1197-1229)
    const bool all = tdel || (!pdel && !(cex[0] && ctdel[0]) && !(cex[1] && ctdel[1]));
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int s = 0; s < 2; s++) {
        if (cex[s] && (all || !ctdel[s])) {
#pragma unroll
            for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], cmn[s][k]); mx[k] = fmaxf(mx[k], cmx[s][k]); }
        }
    }
    if (all || !pdel) {
        mn[0] = fminf(mn[0], a.x); mx[0] = fmaxf(mx[0], a.x);
        mn[1] = fminf(mn[1], a.y); mx[1] = fmaxf(mx[1], a.y);
        mn[2] = fminf(mn[2], a.z); mx[2] = fmaxf(mx[2], a.z);
    }
    // Criterion_Check (:1090-1107)
    bool viol = false;
    if (size > 10) {
        int son = cex[0] ? csize[0] : csize[1];
        float de = (float)invalid / (float)size;
        float be = (float)son / (float)(size - 1);
        if (de > del_param) viol = true;
        if (be > bal_param || be < 1.0f - bal_param) viol = true;
    }
    uint32_t fl = u.flags & ~(F_TDEL | F_TDS | F_VIOL);
    if (tdel) fl |= F_TDEL;
    if (tds) fl |= F_TDS;
    if (viol) fl |= F_VIOL;
    u.flags = fl;
    u.size = size; u.invalid = invalid; u.down_del = dd;
#pragma unroll
    for (int k = 0; k < 3; k++) { u.bmin[k] = mn[k]; u.bmax[k] = mx[k]; }
    store_urec(c.urec + n, u);
    // search record: deleted bit + search-effective child boxes
    meta = pdel ? (meta | META_PDEL) : (meta & ~META_PDEL);
    float b[12];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        bool vis = cex[s] && !ctdel[s];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            b[6 * s + k] = vis ? cmn[s][k] : CUDART_INF_F;
            b[6 * s + 3 + k] = vis ? cmx[s][k] : -CUDART_INF_F;
        }
    }
    float4* q = reinterpret_cast<float4*>(sr);
    q[0] = make_float4(a.x, a.y, a.z, __uint_as_float(meta));
    q[1] = make_float4(b[0], b[1], b[2], b[3]);
    q[2] = make_float4(b[4], b[5], b[6], b[7]);
    q[3] = make_float4(b[8], b[9], b[10], b[11]);
    if (u.parent == 0) {
        TreeHeader* h = c.hdr;
        h->root_exists = 1;
        h->root_searchable = tdel ? 0 : 1;
        h->size = size;
        h->invalid = invalid;
#pragma unroll
        for (int k = 0; k < 3; k++) { h->range[k] = mn[k]; h->range[3 + k] = mx[k]; }
        if (size > 3) {  // :1315-1321
            int son = cex[0] ? csize[0] : csize[1];
            float tb = (float)son / (float)(size - 1);
            h->alpha_del = (float)invalid / (float)size;
            h->alpha_bal = ((double)tb >= 0.5 - 1e-6) ? tb : 1.0f - tb;
        }
    }
}

__global__ void refit_kernel(Ctx c, const int32_t* __restrict__ dirty, const unsigned int* __restrict__ ndirty,
                             const uint8_t* __restrict__ start, float del_param, float bal_param) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *ndirty || !start[i]) return;
    int n = dirty[i];
    while (true) {
        recompute_node(c, n, del_param, bal_param);
        c.urec[n].pending = -1;
        int p = c.urec[n].parent;
        if (p == 0) break;
        __threadfence();
        int old = atomicSub(&c.urec[p].pending, 1);
        if (old != 1) break;  // a sibling subtree is still being refit; its thread will take the parent
        n = p;
    }
}

// topmost violating nodes among the dirty set -> rebuild roots
__global__ void collect_viol_kernel(Ctx c, const int32_t* __restrict__ dirty, const unsigned int* __restrict__ ndirty,
                                    int32_t* __restrict__ roots, unsigned int* __restrict__ nroots) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *ndirty) return;
    int n = dirty[i];
    if (!(c.urec[n].flags & F_VIOL)) return;
    int p = c.urec[n].parent;
    while (p) {
        if (c.urec[p].flags & F_VIOL) return;
        p = c.urec[p].parent;
    }
    roots[atomicAdd(nroots, 1u)] = n;
}

// ================================================================================================
// rebuild (Rebuild :625-645, flatten :1326-1352)
// ================================================================================================
// per root: [0] valid count, [1] TreeSize, [2] block slots (0 if <2 valid), [3] parent, [4] depth
__global__ void root_info_kernel(Ctx c, const int32_t* __restrict__ roots, int R, int* __restrict__ nvalid,
                                 int* __restrict__ tsize, int* __restrict__ bslots) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const UpdateRec& u = c.urec[roots[r]];
    int nv = u.size - u.invalid;
    nvalid[r] = nv;
    tsize[r] = u.size;
    int levels = nv > 0 ? 32 - __clz(nv) : 0;
    bslots[r] = nv >= 2 ? (1 << levels) : 0;
}

__global__ void forest_setup_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ seg_begin,
                                    const int* __restrict__ boff, unsigned int pool_base, int* __restrict__ root_slot,
                                    int* __restrict__ block_base, int* __restrict__ root_parent,
                                    int* __restrict__ root_depth, int* __restrict__ single_axis) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int s = roots[r];
    const UpdateRec& u = c.urec[s];
    root_slot[r] = s;
    block_base[r] = (int)pool_base + boff[r];
    root_parent[r] = u.parent;
    root_depth[r] = u.depth;
    single_axis[r] = -1;
    if (seg_begin[r + 1] == seg_begin[r]) {
        // no valid point left: the subtree vanishes (BuildTree on an empty range leaves *root null, :575)
        c.urec[s].flags = 0;
        c.urec[s].pending = -1;
        if (u.parent == 0) { c.hdr->root_exists = 0; c.hdr->root_searchable = 0; c.hdr->size = 0; c.hdr->invalid = 0; }
    }
}

// One block per rebuild root: pre-order flatten of the valid points with exact output offsets
// (offset of a node = offset of its parent + [parent valid] (+ valid count of the left sibling)),
// so the point order is the reference's flatten order without atomics. Old nodes are released.
constexpr int FL_TPB = 256;
__global__ void __launch_bounds__(FL_TPB)
flatten_kernel(Ctx c, const int32_t* __restrict__ roots, const int* __restrict__ seg_begin,
               const long long* __restrict__ stack_off, uint2* __restrict__ stack_mem, float4* __restrict__ p4,
               int* __restrict__ eroot, int32_t* __restrict__ removed, unsigned int* __restrict__ nremoved,
               unsigned int removed_cap) {
    typedef cub::BlockScan<int, FL_TPB> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_top;
    const int r = blockIdx.x, tid = threadIdx.x;
    uint2* stack = stack_mem + stack_off[r];
    const int root = roots[r];
    if (tid == 0) { stack[0] = make_uint2((unsigned)root, (unsigned)seg_begin[r]); s_top = 1; }
    __syncthreads();
    while (true) {
        int top = s_top;
        if (top == 0) break;
        int take = top < FL_TPB ? top : FL_TPB;
        bool active = tid < take;
        uint2 ent = active ? stack[top - 1 - tid] : make_uint2(0, 0);
        __syncthreads();
        int npush = 0;
        uint2 pu[2];
        if (active) {
            int slot = (int)ent.x;
            int off = (int)ent.y;
            float4 a = reinterpret_cast<const float4*>(c.srec + slot)[0];
            UpdateRec u = c.urec[slot];
            bool valid = !(u.flags & F_PDEL);
            if (valid) {
                p4[off] = make_float4(a.x, a.y, a.z, __int_as_float(u.pid));
                eroot[off] = r;
            } else if (!(u.flags & F_PDS)) {
                unsigned int k = atomicAdd(nremoved, 1u);  // Points_deleted (:1339-1341)
                if (k < removed_cap) removed[k] = u.pid;
            }
            uint32_t cp = meta_cp(__float_as_uint(a.w));
            int coff = off + (valid ? 1 : 0);
            if (cp) {
                const UpdateRec& L = c.urec[2 * cp];
                const UpdateRec& Rr = c.urec[2 * cp + 1];
                if (L.flags & F_EXISTS) { pu[npush++] = make_uint2(2 * cp, (unsigned)coff); coff += L.size - L.invalid; }
                if (Rr.flags & F_EXISTS) { pu[npush++] = make_uint2(2 * cp + 1, (unsigned)coff); }
            }
            if (slot != root) { c.urec[slot].flags = 0; c.urec[slot].pending = -1; }
        }
        int pos, total;
        Scan(tmp).ExclusiveSum(npush, pos, total);
        int base = top - take;
        // push in reverse so that the left child is popped first (not required for correctness)
        if (npush >= 1) stack[base + pos] = pu[0];
        if (npush == 2) stack[base + pos + 1] = pu[1];
        __syncthreads();
        if (tid == 0) s_top = base + total;
        __syncthreads();
    }
}

__global__ void gather_roots_kernel(const int32_t* __restrict__ roots, int R, Ctx c, int32_t* __restrict__ changed,
                                    unsigned int* __restrict__ nchanged) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int s = roots[r];
    // existing roots are refit themselves (harmless recompute) so that their ancestors follow; a vanished root
    // hands over to its parent
    int v = (c.urec[s].flags & F_EXISTS) ? s : c.urec[s].parent;
    if (v > 0) changed[atomicAdd(nchanged, 1u)] = v;
}

// alive[pid] = 1 for every valid point; logs removed points (whole-tree rebuild / flatten export)
__global__ void alive_kernel(Ctx c, unsigned int pool_top, uint8_t* __restrict__ alive, bool log_removed,
                             int32_t* __restrict__ removed, unsigned int* __restrict__ nremoved, unsigned int removed_cap) {
    unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= pool_top || s == 0) return;
    const UpdateRec& u = c.urec[s];
    if (!(u.flags & F_EXISTS)) return;
    if (!(u.flags & F_PDEL)) alive[u.pid] = 1;
    else if (log_removed && !(u.flags & F_PDS)) {
        unsigned int k = atomicAdd(nremoved, 1u);
        if (k < removed_cap) removed[k] = u.pid;
    }
}

__global__ void gather_pid_kernel(const int32_t* __restrict__ pids, int n, const float4* __restrict__ pid_xyz,
                                  float4* __restrict__ p4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pid = pids[i];
    float4 v = pid_xyz[pid];
    p4[i] = make_float4(v.x, v.y, v.z, __int_as_float(pid));
}

// ================================================================================================
// Delete_Points (Delete_by_point :713-760)
// ================================================================================================
__device__ __forceinline__ bool same_point_d(float ax, float ay, float az, float bx, float by, float bz) {
    // same_point :1369-1371: fabs(float diff) promoted to double against EPSS = 1e-6
    return (double)fabsf(__fsub_rn(ax, bx)) < 1e-6 && (double)fabsf(__fsub_rn(ay, by)) < 1e-6 &&
           (double)fabsf(__fsub_rn(az, bz)) < 1e-6;
}

__global__ void delete_points_kernel(Ctx c, const float4* __restrict__ pts, int n, int32_t* __restrict__ changed,
                                     unsigned int* __restrict__ nchanged) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !c.hdr->root_exists) return;
    float4 p = pts[i];
    uint32_t cur = ROOT_SLOT;
    while (cur) {
        uint32_t fl = __ldcg(&c.urec[cur].flags);
        if (fl & F_TDEL) return;  // :714
        float4 a = __ldcg(reinterpret_cast<const float4*>(c.srec + cur));
        uint32_t meta = __float_as_uint(a.w);
        if (same_point_d(a.x, a.y, a.z, p.x, p.y, p.z)) {
            uint32_t old = atomicOr(&c.urec[cur].flags, F_PDEL);
            if (!(old & F_PDEL)) {  // this thread deleted it (:717-722)
                atomicOr(&c.srec[cur].meta, META_PDEL);
                changed[atomicAdd(nchanged, 1u)] = (int32_t)cur;
                return;
            }
        }
        int ax = meta_axis(meta);
        float pc = ax == 0 ? p.x : (ax == 1 ? p.y : p.z);
        float nc = ax == 0 ? a.x : (ax == 1 ? a.y : a.z);
        uint32_t cp = meta_cp(meta);
        if (!cp) return;
        uint32_t ch = 2 * cp + (pc < nc ? 0u : 1u);
        if (!(c.urec[ch].flags & F_EXISTS)) return;
        cur = ch;
    }
}

// ================================================================================================
// Add_by_point as a bulk insert (:818-866)
// ================================================================================================
// descend to the empty child position each point would be appended at; key = parent slot * 2 + side
__global__ void descend_kernel(Ctx c, const float4* __restrict__ pts, int n, uint32_t* __restrict__ keys,
                               int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pts[i];
    uint32_t cur = ROOT_SLOT;
    uint32_t key;
    while (true) {
        float4 a = reinterpret_cast<const float4*>(c.srec + cur)[0];
        uint32_t meta = __float_as_uint(a.w);
        int ax = meta_axis(meta);
        float pc = ax == 0 ? p.x : (ax == 1 ? p.y : p.z);
        float nc = ax == 0 ? a.x : (ax == 1 ? a.y : a.z);
        uint32_t side = pc < nc ? 0u : 1u;  // :833
        uint32_t cp = meta_cp(meta);
        key = cur * 2 + side;
        if (!cp) break;
        uint32_t ch = 2 * cp + side;
        if (!(c.urec[ch].flags & F_EXISTS)) break;
        cur = ch;
    }
    keys[i] = key;
    idx[i] = i;
}

__global__ void group_flag_kernel(const uint32_t* __restrict__ keys, int n, int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// after an inclusive scan of head flags: gid[i]-1 is the group of sorted element i
__global__ void group_bounds_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ gid, int n,
                                    int* __restrict__ seg_begin, uint32_t* __restrict__ gkey, int* __restrict__ eroot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = gid[i] - 1;
    eroot[i] = g;
    if (i == 0 || keys[i] != keys[i - 1]) { seg_begin[g] = i; gkey[g] = keys[i]; }
    if (i == n - 1) seg_begin[g + 1] = n;
}

// one thread per group: the first group of each distinct parent allocates the child pair if missing
__global__ void alloc_pairs_kernel(Ctx c, const uint32_t* __restrict__ gkey, int R) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= R) return;
    uint32_t parent = gkey[g] >> 1;
    if (g > 0 && (gkey[g - 1] >> 1) == parent) return;
    uint32_t meta = c.srec[parent].meta;
    if (meta_cp(meta)) return;
    uint32_t slot = atomicAdd(&c.hdr->pool_top, 2u);
    uint32_t cp = slot >> 1;
    UpdateRec z;
    memset(&z, 0, sizeof(z));
    z.pending = -1;
    store_urec(c.urec + slot, z);
    store_urec(c.urec + slot + 1, z);
    c.srec[parent].meta = meta | (cp << META_CP_SHIFT);
}

// block slots needed by each insert group (a heap-ordered block of 2^levels slots when it has >= 2 points)
__global__ void insert_sizes_kernel(const int* __restrict__ seg_begin, int R, int* __restrict__ bslots,
                                    int* __restrict__ maxseg) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= R) return;
    int n = seg_begin[g + 1] - seg_begin[g];
    bslots[g] = n >= 2 ? (1 << (32 - __clz(n))) : 0;
    atomicMax(maxseg, n);
}

__global__ void insert_forest_kernel(Ctx c, const uint32_t* __restrict__ gkey, int R, const int* __restrict__ boff,
                                     unsigned int pool_base, int* __restrict__ root_slot, int* __restrict__ block_base,
                                     int* __restrict__ root_parent, int* __restrict__ root_depth,
                                     int* __restrict__ single_axis) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= R) return;
    uint32_t parent = gkey[g] >> 1, side = gkey[g] & 1u;
    uint32_t meta = c.srec[parent].meta;
    uint32_t cp = meta_cp(meta);
    root_slot[g] = (int)(2 * cp + side);
    root_parent[g] = (int)parent;
    root_depth[g] = c.urec[parent].depth + 1;
    single_axis[g] = (meta_axis(meta) + 1) % 3;  // :823
    block_base[g] = (int)pool_base + boff[g];
}

__global__ void gather_sorted_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, int n, int first_pid,
                                     float4* __restrict__ p4, float4* __restrict__ pid_xyz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = idx[i];
    float4 v = pts[j];
    int pid = first_pid + j;  // ids follow the order of the survivor list, not the sorted order
    p4[i] = make_float4(v.x, v.y, v.z, __int_as_float(pid));
    pid_xyz[pid] = make_float4(v.x, v.y, v.z, 0.f);
}

__global__ void roots_to_changed_kernel(const int* __restrict__ root_slot, int R, int32_t* __restrict__ changed,
                                        unsigned int* __restrict__ nchanged) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= R) return;
    changed[atomicAdd(nchanged, 1u)] = root_slot[g];
}

// ================================================================================================
// Add_Points voxel downsample (:423-469)
// ================================================================================================
struct VoxOut {
    int del_box;   // 1: downsample-delete the voxel box before inserting
    int kind;      // 0 nothing, 1 insert new point `ref` (batch index), 2 re-insert existing point id `ref`
    int ref;
    int acts;      // insert branches taken (the reference's return value contribution)
};

__device__ __forceinline__ void voxel_box(float v, float ds, float& lo, float& hi, float& mid) {
    // :424-432  floor(p/ds)*ds in fp32, centre through double
    lo = __fmul_rn(floorf(__fdiv_rn(v, ds)), ds) + 0.0f;  // +0 canonicalises -0
    hi = __fadd_rn(lo, ds);
    mid = (float)((double)lo + (double)__fsub_rn(hi, lo) / 2.0);
}

__global__ void voxel_key_kernel(const float4* __restrict__ pts, int n, float ds, uint32_t* __restrict__ kx,
                                 uint32_t* __restrict__ ky, uint32_t* __restrict__ kz, int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pts[i];
    float lo, hi, mid;
    voxel_box(p.x, ds, lo, hi, mid); kx[i] = __float_as_uint(lo);
    voxel_box(p.y, ds, lo, hi, mid); ky[i] = __float_as_uint(lo);
    voxel_box(p.z, ds, lo, hi, mid); kz[i] = __float_as_uint(lo);
    idx[i] = i;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const int* __restrict__ idx, int n,
                                  uint32_t* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void voxel_head_kernel(const uint32_t* __restrict__ kx, const uint32_t* __restrict__ ky,
                                  const uint32_t* __restrict__ kz, const int* __restrict__ idx, int n,
                                  int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = idx[i];
    bool h = true;
    if (i > 0) { int b = idx[i - 1]; h = kx[a] != kx[b] || ky[a] != ky[b] || kz[a] != kz[b]; }
    head[i] = h ? 1 : 0;
}

__global__ void voxel_bounds_kernel(const int* __restrict__ head, const int* __restrict__ gid, int n,
                                    int* __restrict__ seg_begin) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = gid[i] - 1;
    if (head[i]) seg_begin[g] = i;
    if (i == n - 1) seg_begin[g + 1] = n;
}

// a coordinate is "regular" for voxel index nf if it lies in box nf and in neither neighbour box, so the
// voxel groups of one batch touch disjoint point sets and can be processed independently
__device__ __forceinline__ bool regular_coord(float x, float nf, float ds) {
    float lo = __fmul_rn(nf, ds), hi = __fadd_rn(lo, ds);
    float lo_next = __fmul_rn(nf + 1.0f, ds);
    float hi_prev = __fadd_rn(__fmul_rn(nf - 1.0f, ds), ds);
    return x >= lo && x < hi && x < lo_next && x >= hi_prev;
}

// One thread per voxel group: box-search the tree (existing points), then replay the reference's
// per-point decisions for the new points of this voxel in input order.
__global__ void voxel_decide_kernel(Ctx c, const float4* __restrict__ pts, const int* __restrict__ idx,
                                    const int* __restrict__ seg_begin, int G, float ds, VoxOut* __restrict__ out,
                                    float* __restrict__ boxes, int* __restrict__ irregular) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int b = seg_begin[g], e = seg_begin[g + 1];
    float4 p0 = pts[idx[b]];
    float lo[3], hi[3], mid[3], nf[3];
    voxel_box(p0.x, ds, lo[0], hi[0], mid[0]);
    voxel_box(p0.y, ds, lo[1], hi[1], mid[1]);
    voxel_box(p0.z, ds, lo[2], hi[2], mid[2]);
    nf[0] = floorf(__fdiv_rn(p0.x, ds)); nf[1] = floorf(__fdiv_rn(p0.y, ds)); nf[2] = floorf(__fdiv_rn(p0.z, ds));
    bool reg = true;
    // existing points in the half-open box (Search_by_range :1016-1044)
    int cnt = 0;
    float best_d = CUDART_INF_F;
    int best_pid = 0x7fffffff;
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (c.hdr->root_exists) {
        uint32_t st[64];
        int sp = 0;
        const float* rg = c.hdr->range;
        bool dis = hi[0] <= rg[0] || lo[0] > rg[3] || hi[1] <= rg[1] || lo[1] > rg[4] || hi[2] <= rg[2] || lo[2] > rg[5];
        if (!dis) st[sp++] = ROOT_SLOT;
        while (sp > 0) {
            uint32_t cur = st[--sp];
            const float4* r = reinterpret_cast<const float4*>(c.srec + cur);
            float4 a = r[0], q1 = r[1], q2 = r[2], q3 = r[3];
            uint32_t meta = __float_as_uint(a.w);
            if (!(meta & META_PDEL) && lo[0] <= a.x && hi[0] > a.x && lo[1] <= a.y && hi[1] > a.y && lo[2] <= a.z && hi[2] > a.z) {
                cnt++;
                float d = sq_dist3(a.x, a.y, a.z, mid[0], mid[1], mid[2]);
                int pid = c.urec[cur].pid;
                if (d < best_d || (d == best_d && pid < best_pid)) { best_d = d; best_pid = pid; bx = a.x; by = a.y; bz = a.z; }
                reg = reg && regular_coord(a.x, nf[0], ds) && regular_coord(a.y, nf[1], ds) && regular_coord(a.z, nf[2], ds);
            }
            uint32_t cp = meta_cp(meta);
            if (cp) {
                bool dl = hi[0] <= q1.x || lo[0] > q1.w || hi[1] <= q1.y || lo[1] > q2.x || hi[2] <= q1.z || lo[2] > q2.y;
                bool dr = hi[0] <= q2.z || lo[0] > q3.y || hi[1] <= q2.w || lo[1] > q3.z || hi[2] <= q3.x || lo[2] > q3.w;
                if (!dr && sp < 64) st[sp++] = 2 * cp + 1;
                if (!dl && sp < 64) st[sp++] = 2 * cp;
            }
        }
    }
    // replay (:435-449)
    int c_exist = cnt;                 // points of the box currently in the tree (as the reference would see it)
    bool have_inc = cnt >= 1;
    float inc_d = best_d, ix = bx, iy = by, iz = bz;
    int inc_kind = 2, inc_ref = best_pid;
    int acts = 0;
    for (int k = b; k < e; k++) {
        int j = idx[k];
        float4 p = pts[j];
        reg = reg && regular_coord(p.x, nf[0], ds) && regular_coord(p.y, nf[1], ds) && regular_coord(p.z, nf[2], ds);
        float dp = sq_dist3(p.x, p.y, p.z, mid[0], mid[1], mid[2]);
        bool use_inc = have_inc && inc_d < dp;  // strict: the new point wins ties (:439)
        bool act = c_exist > 1 || (use_inc ? same_point_d(p.x, p.y, p.z, ix, iy, iz) : true);  // :445
        if (act) {
            acts++;
            if (!use_inc) { inc_d = dp; ix = p.x; iy = p.y; iz = p.z; inc_kind = 1; inc_ref = j; }
            have_inc = true;
            c_exist = 1;
        }
    }
    VoxOut o;
    o.acts = acts;
    o.del_box = (acts > 0 && cnt > 0) ? 1 : 0;
    o.kind = acts > 0 ? inc_kind : 0;
    o.ref = inc_ref;
    out[g] = o;
    float* bb = boxes + 6 * (size_t)g;
    bb[0] = lo[0]; bb[1] = lo[1]; bb[2] = lo[2]; bb[3] = hi[0]; bb[4] = hi[1]; bb[5] = hi[2];
    if (!reg) atomicExch(irregular, 1);
}

// compact the voxel decisions: delete boxes, survivors (coordinates + payload source), act count
__global__ void voxel_apply_kernel(const VoxOut* __restrict__ vo, int G, const float* __restrict__ boxes,
                                   const float4* __restrict__ pts, const float4* __restrict__ pid_xyz,
                                   const int* __restrict__ del_pos, const int* __restrict__ ins_pos,
                                   float* __restrict__ del_boxes, float4* __restrict__ surv, int32_t* __restrict__ src,
                                   int src_base) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    VoxOut o = vo[g];
    if (o.del_box) {
        int k = del_pos[g];
        for (int a = 0; a < 6; a++) del_boxes[6 * (size_t)k + a] = boxes[6 * (size_t)g + a];
    }
    if (o.kind) {
        int k = ins_pos[g];
        float4 v = o.kind == 1 ? pts[o.ref] : pid_xyz[o.ref];
        surv[k] = make_float4(v.x, v.y, v.z, 0.f);
        src[k] = o.kind == 1 ? src_base + o.ref : ~o.ref;
    }
}

struct VoxDel { __host__ __device__ int operator()(const VoxOut& o) const { return o.del_box; } };
struct VoxIns { __host__ __device__ int operator()(const VoxOut& o) const { return o.kind ? 1 : 0; } };
struct VoxAct { __host__ __device__ int operator()(const VoxOut& o) const { return o.acts; } };

__global__ void iota_src_kernel(int32_t* src, int n, int base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) src[i] = base + i;
}

// ------------------------------------------------------------------------------------------------
template <class T>
int d2h(ikd_tree* t, T* host, const void* dev, size_t count) {
    IKD_TRY(ensure_pin(t, sizeof(T) * count));
    IKD_CUDA(cudaMemcpyAsync(t->pin, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, t->stream));
    IKD_CUDA(cudaStreamSynchronize(t->stream));
    memcpy(host, t->pin, sizeof(T) * count);
    return IKD_OK;
}

Ctx ctx_of(ikd_tree* t) { return Ctx{t->srec, t->urec, t->hdr_dev}; }

int cub_exclusive_sum_int(ikd_tree* t, const int* in, int* out, int n) {
    size_t tmp = 0;
    IKD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, t->stream));
    IKD_TRY(t->b_cubtmp.ensure(tmp, t->stream));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA(cub::DeviceScan::ExclusiveSum(t->b_cubtmp.p, tb, in, out, n, t->stream));
    return IKD_OK;
}
int cub_inclusive_sum_int(ikd_tree* t, const int* in, int* out, int n) {
    size_t tmp = 0;
    IKD_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, n, t->stream));
    IKD_TRY(t->b_cubtmp.ensure(tmp, t->stream));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA(cub::DeviceScan::InclusiveSum(t->b_cubtmp.p, tb, in, out, n, t->stream));
    return IKD_OK;
}
int cub_sort_pairs_u32(ikd_tree* t, const uint32_t* kin, uint32_t* kout, const int* vin, int* vout, int n, int end_bit = 32) {
    size_t tmp = 0;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, int>(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, n, 0, end_bit, t->stream)));
    IKD_TRY(t->b_cubtmp.ensure(tmp, t->stream));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, int>(t->b_cubtmp.p, tb, kin, kout, vin, vout, n, 0, end_bit, t->stream)));
    return IKD_OK;
}

int ensure_removed_cap(ikd_tree* t) {
    int64_t need = std::max<int64_t>(t->next_pid, 1024);
    if (need > t->removed_cap) {
        IKD_TRY(t->b_removed.ensure((size_t)need * 4 + 16, t->stream, true));
        t->removed_cap = (int64_t)((t->b_removed.bytes - 16) / 4);
    }
    return IKD_OK;
}
// the removed-point counter lives in the last 4 bytes... keep it simple: a dedicated small buffer
unsigned int* removed_counter(ikd_tree* t) { return t->u[U_CNT].as<unsigned int>() + 8; }

int ensure_counters(ikd_tree* t) {
    if (!t->u[U_CNT].p) {
        IKD_TRY(t->u[U_CNT].ensure(64 * sizeof(unsigned int), t->stream));
        IKD_CUDA(cudaMemsetAsync(t->u[U_CNT].p, 0, 64 * sizeof(unsigned int), t->stream));
    }
    return IKD_OK;
}

// select alive point ids in increasing id order into U_SEL; returns their number
int select_alive(ikd_tree* t, bool log_removed, int* out_n) {
    cudaStream_t s = t->stream;
    *out_n = 0;
    int np = t->next_pid;
    if (np == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(ensure_counters(t));
    IKD_TRY(ensure_removed_cap(t));
    IKD_TRY(t->u[U_ALIVE].ensure((size_t)np, s));
    IKD_TRY(t->u[U_SEL].ensure((size_t)np * 4, s));
    IKD_TRY(t->u[U_TMP].ensure(16, s));
    IKD_CUDA(cudaMemsetAsync(t->u[U_ALIVE].p, 0, (size_t)np, s));
    unsigned int top = t->hdr.pool_top;
    IKD_LAUNCH alive_kernel<<<nblk(top), TPB, 0, s>>>(ctx_of(t), top, t->u[U_ALIVE].as<uint8_t>(), log_removed,
                                           t->b_removed.as<int32_t>(), removed_counter(t), (unsigned)t->removed_cap);
    size_t tmp = 0;
    thrust::counting_iterator<int> it(0);
    IKD_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp, it, t->u[U_ALIVE].as<uint8_t>(), t->u[U_SEL].as<int32_t>(),
                                        t->u[U_TMP].as<int>(), np, s));
    IKD_TRY(t->b_cubtmp.ensure(tmp, s));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA(cub::DeviceSelect::Flagged(t->b_cubtmp.p, tb, it, t->u[U_ALIVE].as<uint8_t>(), t->u[U_SEL].as<int32_t>(),
                                        t->u[U_TMP].as<int>(), np, s));
    IKD_TRY(d2h(t, out_n, t->u[U_TMP].p, 1));
    return IKD_OK;
}

// refit the ancestors of U_CHANGED[0..*U_NCHANGED), then collect rebuild roots into U_ROOTS; returns their number
int refit_and_collect(ikd_tree* t, int64_t changed_cap, int* out_R) {
    cudaStream_t s = t->stream;
    Ctx c = ctx_of(t);
    // dirty set is bounded by changed * (depth+1) and by the number of slots in use
    int64_t dcap = std::min<int64_t>(changed_cap * (int64_t)(t->hdr.max_depth + 36), (int64_t)t->cap_slots);
    dcap = std::max<int64_t>(dcap, 64);
    IKD_TRY(t->u[U_DIRTY].ensure((size_t)dcap * 4, s));
    IKD_TRY(t->u[U_START].ensure((size_t)dcap, s));
    IKD_TRY(t->u[U_ROOTS].ensure((size_t)dcap * 4, s));
    unsigned int* cnt = t->u[U_CNT].as<unsigned int>();  // [0] nchanged, [1] ndirty, [2] nroots
    IKD_CUDA(cudaMemsetAsync(cnt + 1, 0, 2 * sizeof(unsigned int), s));
    int32_t* changed = t->u[U_CHANGED].as<int32_t>();
    int32_t* dirty = t->u[U_DIRTY].as<int32_t>();
    IKD_LAUNCH mark_kernel<<<nblk(changed_cap), TPB, 0, s>>>(c, changed, cnt, dirty, cnt + 1);
    IKD_LAUNCH starters_kernel<<<nblk(dcap), TPB, 0, s>>>(c, dirty, cnt + 1, t->u[U_START].as<uint8_t>());
    IKD_LAUNCH refit_kernel<<<nblk(dcap), TPB, 0, s>>>(c, dirty, cnt + 1, t->u[U_START].as<uint8_t>(), t->delete_param, t->balance_param);
    IKD_LAUNCH collect_viol_kernel<<<nblk(dcap), TPB, 0, s>>>(c, dirty, cnt + 1, t->u[U_ROOTS].as<int32_t>(), cnt + 2);
    unsigned int h[3];
    IKD_TRY(d2h(t, h, cnt, 3));
    IKD_CUDA(cudaGetLastError());
    *out_R = (int)h[2];
    return IKD_OK;
}

int rebuild_forest(ikd_tree* t, int R);

// After a batch touched the slots in U_CHANGED: refit, rebuild violating subtrees, repeat (:704-707).
int settle(ikd_tree* t, int64_t changed_cap) {
    for (int iter = 0; iter < 64; iter++) {
        int R = 0;
        IKD_TRY(refit_and_collect(t, changed_cap, &R));
        if (R == 0) break;
        IKD_TRY(rebuild_forest(t, R));
        changed_cap = R;
    }
    IKD_TRY(sync_header(t));
    if (t->hdr.max_depth >= 60) IKD_TRY(rebuild_all(t));  // keep traversal stacks bounded
    return IKD_OK;
}

// Rebuild the R subtrees rooted at U_ROOTS. Leaves the next changed list in U_CHANGED / cnt[0].
int rebuild_forest(ikd_tree* t, int R) {
    cudaStream_t s = t->stream;
    Ctx c = ctx_of(t);
    unsigned int* cnt = t->u[U_CNT].as<unsigned int>();
    int32_t* roots = t->u[U_ROOTS].as<int32_t>();
    // whole tree? then rebuild from the id table and compact the pool
    if (R == 1) {
        int32_t r0;
        IKD_TRY(d2h(t, &r0, roots, 1));
        if (r0 == ROOT_SLOT) {
            IKD_TRY(rebuild_all(t));
            IKD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), s));
            return IKD_OK;
        }
    }
    // make the order of the roots deterministic (atomic append order is not)
    {
        IKD_TRY(t->u[U_TMP].ensure((size_t)R * 4 + 16, s));
        size_t tmp = 0;
        IKD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const int32_t*)roots, t->u[U_TMP].as<int32_t>(), R, 0, 32, s));
        IKD_TRY(t->b_cubtmp.ensure(tmp, s));
        size_t tb = t->b_cubtmp.bytes;
        IKD_CUDA(cub::DeviceRadixSort::SortKeys(t->b_cubtmp.p, tb, (const int32_t*)roots, t->u[U_TMP].as<int32_t>(), R, 0, 32, s));
        IKD_CUDA(cudaMemcpyAsync(roots, t->u[U_TMP].p, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
    }
    // per-root sizes -> segment / stack / block offsets
    IKD_TRY(t->u[U_RINFO].ensure((size_t)(R + 1) * 4 * 8, s));
    int* nvalid = t->u[U_RINFO].as<int>();
    int* tsize = nvalid + (R + 1);
    int* bslots = tsize + (R + 1);
    int* seg_begin = bslots + (R + 1);
    int* soff32 = seg_begin + (R + 1);
    int* boff = soff32 + (R + 1);
    IKD_CUDA(cudaMemsetAsync(nvalid, 0, (size_t)(R + 1) * 4 * 3, s));
    IKD_LAUNCH root_info_kernel<<<nblk(R), TPB, 0, s>>>(c, roots, R, nvalid, tsize, bslots);
    IKD_TRY(cub_exclusive_sum_int(t, nvalid, seg_begin, R + 1));
    IKD_TRY(cub_exclusive_sum_int(t, tsize, soff32, R + 1));
    IKD_TRY(cub_exclusive_sum_int(t, bslots, boff, R + 1));
    int tot[3];
    {
        IKD_TRY(ensure_pin(t, 64));
        int* pp = (int*)t->pin;
        IKD_CUDA(cudaMemcpyAsync(pp, seg_begin + R, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaMemcpyAsync(pp + 1, soff32 + R, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaMemcpyAsync(pp + 2, boff + R, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaStreamSynchronize(s));
        tot[0] = pp[0]; tot[1] = pp[1]; tot[2] = pp[2];
    }
    const int M = tot[0], S = tot[1], B = tot[2];
    // pool room for the new blocks (grows by reallocation; slot numbers stay valid)
    IKD_TRY(sync_header(t));
    unsigned int pool_base = t->hdr.pool_top;
    if ((size_t)pool_base + (size_t)B + 2 > t->cap_slots) {
        IKD_TRY(ensure_pool(t, (size_t)pool_base + (size_t)B + 1024, true));
        c = ctx_of(t);
    }
    // stack offsets as 64-bit
    IKD_TRY(t->u[U_STACK].ensure((size_t)std::max(S, 1) * sizeof(uint2) + (size_t)(R + 1) * 8, s));
    IKD_TRY(t->u[U_P4].ensure((size_t)std::max(M, 1) * sizeof(float4), s));
    IKD_TRY(t->u[U_EROOT].ensure((size_t)std::max(M, 1) * 4, s));
    IKD_TRY(ensure_removed_cap(t));
    // widen soff32 to long long on device: reuse a tiny kernel-free path via cudaMemcpy2D is overkill; do it on host
    std::vector<int> h_soff(R + 1);
    IKD_TRY(d2h(t, h_soff.data(), soff32, R + 1));
    std::vector<long long> h_soff64(R + 1);
    for (int i = 0; i <= R; i++) h_soff64[i] = h_soff[i];
    long long* soff64 = reinterpret_cast<long long*>(t->u[U_STACK].as<char>() + (size_t)std::max(S, 1) * sizeof(uint2));
    IKD_CUDA(cudaMemcpyAsync(soff64, h_soff64.data(), (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, s));
    IKD_CUDA(cudaStreamSynchronize(s));
    IKD_LAUNCH flatten_kernel<<<R, FL_TPB, 0, s>>>(c, roots, seg_begin, soff64, t->u[U_STACK].as<uint2>(), t->u[U_P4].as<float4>(),
                                        t->u[U_EROOT].as<int>(), t->b_removed.as<int32_t>(), removed_counter(t),
                                        (unsigned)t->removed_cap);
    // forest description
    IKD_TRY(t->u[U_FOREST].ensure((size_t)R * 4 * 5 + 64, s));
    int* root_slot = t->u[U_FOREST].as<int>();
    int* block_base = root_slot + R;
    int* root_parent = block_base + R;
    int* root_depth = root_parent + R;
    int* single_axis = root_depth + R;
    IKD_LAUNCH forest_setup_kernel<<<nblk(R), TPB, 0, s>>>(c, roots, R, seg_begin, boff, pool_base, root_slot, block_base, root_parent,
                                                root_depth, single_axis);
    if (B > 0) {
        // every slot below pool_top carries defined flags
        IKD_CUDA(cudaMemsetAsync(t->urec + pool_base, 0, (size_t)B * sizeof(UpdateRec), s));
        t->hdr.pool_top = pool_base + (unsigned)B;
        IKD_CUDA(cudaMemcpyAsync(&t->hdr_dev->pool_top, &t->hdr.pool_top, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
    }
    if (M > 0) {
        ForestDev f;
        f.R = R; f.seg_begin = seg_begin; f.root_slot = root_slot; f.block_base = block_base; f.root_parent = root_parent;
        f.root_depth = root_depth; f.single_axis = single_axis; f.elem_root = R > 1 ? t->u[U_EROOT].as<int>() : nullptr;
        std::vector<int> h_nv(R);
        IKD_TRY(d2h(t, h_nv.data(), nvalid, R));
        int max_seg = 0;
        for (int v : h_nv) max_seg = std::max(max_seg, v);
        IKD_TRY(forest_build(t, t->u[U_P4].as<float4>(), M, f, max_seg, s));
    }
    t->stats.rebuilds_partial += R;
    t->stats.rebuilt_points += M;
    // next round: the rebuilt roots (or the parents of vanished ones)
    IKD_TRY(t->u[U_CHANGED].ensure((size_t)R * 4 + 16, s, false));
    IKD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int), s));
    IKD_LAUNCH gather_roots_kernel<<<nblk(R), TPB, 0, s>>>(roots, R, c, t->u[U_CHANGED].as<int32_t>(), cnt);
    IKD_CUDA(cudaGetLastError());
    return IKD_OK;
}

int prepare_changed(ikd_tree* t, int64_t cap) {
    IKD_TRY(ensure_counters(t));
    IKD_TRY(t->u[U_CHANGED].ensure((size_t)std::max<int64_t>(cap, 16) * 4, t->stream));
    IKD_CUDA(cudaMemsetAsync(t->u[U_CNT].p, 0, 8 * sizeof(unsigned int), t->stream));
    return IKD_OK;
}

// Bulk insert of n device points (float4 xyz) that all become nodes; ids first_pid + i.
int insert_points(ikd_tree* t, const float4* pts, int n) {
    cudaStream_t s = t->stream;
    if (n <= 0) return IKD_OK;
    IKD_TRY(ensure_counters(t));
    IKD_TRY(ensure_pid_cap(t, (int64_t)t->next_pid + n));
    int first_pid = t->next_pid;
    if (!t->hdr.root_exists) {
        // empty tree: the batch becomes the tree (the reference would dereference null here, :447/:472)
        IKD_TRY(t->u[U_P4].ensure((size_t)n * sizeof(float4), s));
        IKD_TRY(t->u[U_IDX].ensure((size_t)n * 4, s));
        IKD_LAUNCH iota_src_kernel<<<nblk(n), TPB, 0, s>>>(t->u[U_IDX].as<int32_t>(), n, 0);
        IKD_LAUNCH gather_sorted_kernel<<<nblk(n), TPB, 0, s>>>(pts, t->u[U_IDX].as<int>(), n, first_pid, t->u[U_P4].as<float4>(),
                                                     t->pid_xyz.as<float4>());
        t->next_pid += n;
        IKD_TRY(full_build(t, t->u[U_P4].as<float4>(), n, s));
        IKD_TRY(sync_header(t));
        return IKD_OK;
    }
    // room: child pairs (<= 2n) + subtree blocks (<= 4n)
    IKD_TRY(sync_header(t));
    if ((size_t)t->hdr.pool_top + 6 * (size_t)n + 64 > t->cap_slots) IKD_TRY(ensure_pool(t, (size_t)t->hdr.pool_top + 6 * (size_t)n + 4096, true));
    Ctx c = ctx_of(t);
    IKD_TRY(t->u[U_KEYS].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_KEYS2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_GROUP].ensure((size_t)n * 4 * 2, s));
    IKD_TRY(t->u[U_GINFO].ensure(((size_t)n + 1) * 4 * 2 + 64, s));
    IKD_TRY(t->u[U_P4].ensure((size_t)n * sizeof(float4), s));
    IKD_TRY(t->u[U_EROOT].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_FOREST].ensure((size_t)n * 4 * 5 + 64, s));
    uint32_t* keys = t->u[U_KEYS].as<uint32_t>();
    uint32_t* keys_s = t->u[U_KEYS2].as<uint32_t>();
    int* idx = t->u[U_IDX].as<int>();
    int* idx_s = t->u[U_IDX2].as<int>();
    int* head = t->u[U_GROUP].as<int>();
    int* gid = head + n;
    int* seg_begin = t->u[U_GINFO].as<int>();
    uint32_t* gkey = reinterpret_cast<uint32_t*>(seg_begin + n + 1);
    IKD_LAUNCH descend_kernel<<<nblk(n), TPB, 0, s>>>(c, pts, n, keys, idx);
    IKD_TRY(cub_sort_pairs_u32(t, keys, keys_s, idx, idx_s, n, 30));
    IKD_LAUNCH group_flag_kernel<<<nblk(n), TPB, 0, s>>>(keys_s, n, head);
    IKD_TRY(cub_inclusive_sum_int(t, head, gid, n));
    IKD_LAUNCH group_bounds_kernel<<<nblk(n), TPB, 0, s>>>(keys_s, gid, n, seg_begin, gkey, t->u[U_EROOT].as<int>());
    int R;
    IKD_TRY(d2h(t, &R, gid + (n - 1), 1));
    IKD_LAUNCH alloc_pairs_kernel<<<nblk(R), TPB, 0, s>>>(c, gkey, R);
    int* root_slot = t->u[U_FOREST].as<int>();
    int* block_base = root_slot + R;
    int* root_parent = block_base + R;
    int* root_depth = root_parent + R;
    int* single_axis = root_depth + R;
    int* maxseg = t->u[U_CNT].as<int>() + 16;
    IKD_CUDA(cudaMemsetAsync(maxseg, 0, 4, s));
    IKD_TRY(t->u[U_RINFO].ensure((size_t)(R + 1) * 4 * 2, s));
    int* bslots = t->u[U_RINFO].as<int>();
    int* boff = bslots + (R + 1);
    IKD_CUDA(cudaMemsetAsync(bslots, 0, (size_t)(R + 1) * 4, s));
    IKD_LAUNCH insert_sizes_kernel<<<nblk(R), TPB, 0, s>>>(seg_begin, R, bslots, maxseg);
    IKD_TRY(cub_exclusive_sum_int(t, bslots, boff, R + 1));
    int max_seg, B;
    unsigned int pool_base;
    {
        IKD_TRY(ensure_pin(t, 64));
        int* pp = (int*)t->pin;
        IKD_CUDA(cudaMemcpyAsync(pp, maxseg, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaMemcpyAsync(pp + 1, boff + R, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaMemcpyAsync(pp + 2, &t->hdr_dev->pool_top, 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaStreamSynchronize(s));
        max_seg = pp[0]; B = pp[1]; pool_base = (unsigned int)pp[2];
    }
    if (B > 0) {
        IKD_CUDA(cudaMemsetAsync(t->urec + pool_base, 0, (size_t)B * sizeof(UpdateRec), s));  // defined flags below pool_top
        t->hdr.pool_top = pool_base + (unsigned)B;
        IKD_CUDA(cudaMemcpyAsync(&t->hdr_dev->pool_top, &t->hdr.pool_top, sizeof(unsigned int), cudaMemcpyHostToDevice, s));
    }
    IKD_LAUNCH insert_forest_kernel<<<nblk(R), TPB, 0, s>>>(c, gkey, R, boff, pool_base, root_slot, block_base, root_parent, root_depth,
                                                 single_axis);
    IKD_LAUNCH gather_sorted_kernel<<<nblk(n), TPB, 0, s>>>(pts, idx_s, n, first_pid, t->u[U_P4].as<float4>(), t->pid_xyz.as<float4>());
    t->next_pid += n;
    ForestDev f;
    f.R = R; f.seg_begin = seg_begin; f.root_slot = root_slot; f.block_base = block_base; f.root_parent = root_parent;
    f.root_depth = root_depth; f.single_axis = single_axis; f.elem_root = R > 1 ? t->u[U_EROOT].as<int>() : nullptr;
    IKD_TRY(forest_build(t, t->u[U_P4].as<float4>(), n, f, max_seg, s));
    // refit from the new subtree roots
    IKD_TRY(prepare_changed(t, R));
    IKD_LAUNCH roots_to_changed_kernel<<<nblk(R), TPB, 0, s>>>(root_slot, R, t->u[U_CHANGED].as<int32_t>(), t->u[U_CNT].as<unsigned int>());
    IKD_TRY(settle(t, R));
    return IKD_OK;
}

}  // namespace

// ================================================================================================
// public implementations
// ================================================================================================
int rebuild_all(ikd_tree* t) {
    cudaStream_t s = t->stream;
    int M = 0;
    IKD_TRY(sync_header(t));
    IKD_TRY(select_alive(t, true, &M));
    IKD_TRY(t->u[U_P4].ensure((size_t)std::max(M, 1) * sizeof(float4), s));
    if (M > 0)
        IKD_LAUNCH gather_pid_kernel<<<nblk(M), TPB, 0, s>>>(t->u[U_SEL].as<int32_t>(), M, t->pid_xyz.as<float4>(), t->u[U_P4].as<float4>());
    IKD_TRY(full_build(t, t->u[U_P4].as<float4>(), M, s));
    IKD_TRY(sync_header(t));
    t->stats.rebuilds_full += 1;
    t->stats.rebuilt_points += M;
    return IKD_OK;
}

int flatten_impl(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n) {
    int M = 0;
    IKD_TRY(sync_header(t));
    IKD_TRY(select_alive(t, false, &M));
    *out_n = M;
    int64_t m = std::min<int64_t>(M, cap);
    if (out_idx && m > 0) {
        IKD_CUDA(cudaMemcpyAsync(out_idx, t->u[U_SEL].p, (size_t)m * 4, cudaMemcpyDeviceToHost, t->stream));
        IKD_CUDA(cudaStreamSynchronize(t->stream));
    }
    return IKD_OK;
}

int acquire_removed_impl(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n) {
    IKD_TRY(ensure_counters(t));
    unsigned int n = 0;
    IKD_TRY(d2h(t, &n, removed_counter(t), 1));
    n = (unsigned)std::min<int64_t>(n, t->removed_cap);
    *out_n = n;
    if (out_idx) {
        int64_t m = std::min<int64_t>(n, cap);
        if (m > 0) {
            IKD_CUDA(cudaMemcpyAsync(out_idx, t->b_removed.p, (size_t)m * 4, cudaMemcpyDeviceToHost, t->stream));
            IKD_CUDA(cudaStreamSynchronize(t->stream));
        }
        IKD_CUDA(cudaMemsetAsync(removed_counter(t), 0, 4, t->stream));  // Points_deleted.clear() (:567)
    }
    return IKD_OK;
}

int delete_boxes_dev(ikd_tree* t, const float* boxes_dev, int64_t nb, bool downsample, int* out_deleted) {
    *out_deleted = 0;
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    int64_t cap = (int64_t)t->hdr.size + 16;
    IKD_TRY(prepare_changed(t, cap));
    unsigned int* cnt = t->u[U_CNT].as<unsigned int>();
    unsigned long long* dcount = reinterpret_cast<unsigned long long*>(cnt + 4);
    int* err = reinterpret_cast<int*>(cnt + 6);
    IKD_TRY(box_delete_launch(t, boxes_dev, nb, downsample, t->u[U_CHANGED].as<int32_t>(), cnt, dcount, err));
    unsigned int h[8];
    IKD_TRY(d2h(t, h, cnt, 8));
    if (h[6]) { set_error("box delete traversal stack overflow"); return IKD_ERR_INTERNAL; }
    unsigned long long dc;
    memcpy(&dc, &h[4], 8);
    *out_deleted = (int)dc;
    if (h[0] > 0) IKD_TRY(settle(t, h[0]));
    return IKD_OK;
}

int delete_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb, int* out_deleted) {
    *out_deleted = 0;
    if (nb == 0) return IKD_OK;
    IKD_TRY(t->u[U_BOXES].ensure((size_t)nb * 24, t->stream));
    IKD_CUDA(cudaMemcpyAsync(t->u[U_BOXES].p, boxes_host, (size_t)nb * 24, cudaMemcpyHostToDevice, t->stream));
    return delete_boxes_dev(t, t->u[U_BOXES].as<float>(), nb, false, out_deleted);
}

int delete_points_impl(ikd_tree* t, const float* xyz, int64_t n, int64_t stride) {
    cudaStream_t s = t->stream;
    if (n == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(t->u[U_PTS].ensure((size_t)n * sizeof(float4), s));
    IKD_TRY(upload_points_f4(t, xyz, n, stride, t->u[U_PTS].as<float4>(), 0, 1));
    IKD_TRY(prepare_changed(t, n));
    unsigned int* cnt = t->u[U_CNT].as<unsigned int>();
    IKD_LAUNCH delete_points_kernel<<<nblk(n), TPB, 0, s>>>(ctx_of(t), t->u[U_PTS].as<float4>(), (int)n, t->u[U_CHANGED].as<int32_t>(), cnt);
    unsigned int h;
    IKD_TRY(d2h(t, &h, cnt, 1));
    if (h > 0) IKD_TRY(settle(t, h));
    return IKD_OK;
}

namespace {
// one contiguous piece [off, off+n) of an Add_Points batch with downsampling; returns IKD_OK and sets
// *irregular when the piece must be split (nothing has been modified in that case)
int add_downsample_piece(ikd_tree* t, const float4* pts, int n, int src_base, bool force, int* acts_out,
                         int64_t* nins_out, int32_t* src_host, int* irregular) {
    cudaStream_t s = t->stream;
    *irregular = 0;
    *acts_out = 0;
    *nins_out = 0;
    Ctx c = ctx_of(t);
    float ds = t->downsample;
    // 1. group the new points by voxel (stable three-pass sort on the voxel min corner bits)
    IKD_TRY(t->u[U_VOX].ensure((size_t)n * 4 * 3, s));
    IKD_TRY(t->u[U_KEYS].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_KEYS2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_GROUP].ensure((size_t)n * 4 * 2, s));
    IKD_TRY(t->u[U_GINFO].ensure(((size_t)n + 1) * 4 * 3 + 64, s));
    uint32_t* kx = t->u[U_VOX].as<uint32_t>();
    uint32_t* ky = kx + n;
    uint32_t* kz = ky + n;
    int* idx_a = t->u[U_IDX].as<int>();
    int* idx_b = t->u[U_IDX2].as<int>();
    uint32_t* k1 = t->u[U_KEYS].as<uint32_t>();
    uint32_t* k2 = t->u[U_KEYS2].as<uint32_t>();
    IKD_LAUNCH voxel_key_kernel<<<nblk(n), TPB, 0, s>>>(pts, n, ds, kx, ky, kz, idx_a);
    const uint32_t* comp[3] = {kz, ky, kx};
    for (int pass = 0; pass < 3; pass++) {
        IKD_LAUNCH gather_u32_kernel<<<nblk(n), TPB, 0, s>>>(comp[pass], idx_a, n, k1);
        IKD_TRY(cub_sort_pairs_u32(t, k1, k2, idx_a, idx_b, n));
        std::swap(idx_a, idx_b);
    }
    int* head = t->u[U_GROUP].as<int>();
    int* gid = head + n;
    int* seg_begin = t->u[U_GINFO].as<int>();
    IKD_LAUNCH voxel_head_kernel<<<nblk(n), TPB, 0, s>>>(kx, ky, kz, idx_a, n, head);
    IKD_TRY(cub_inclusive_sum_int(t, head, gid, n));
    IKD_LAUNCH voxel_bounds_kernel<<<nblk(n), TPB, 0, s>>>(head, gid, n, seg_begin);
    int G;
    IKD_TRY(d2h(t, &G, gid + (n - 1), 1));
    // 2. per-voxel decision
    IKD_TRY(t->u[U_TMP].ensure((size_t)G * sizeof(VoxOut) + 64, s));
    IKD_TRY(t->u[U_TMP2].ensure((size_t)G * 24 + 64, s));
    IKD_TRY(t->u[U_ERR].ensure(64, s));
    VoxOut* vo = t->u[U_TMP].as<VoxOut>();
    float* vboxes = t->u[U_TMP2].as<float>();
    int* irr = t->u[U_ERR].as<int>();
    IKD_CUDA(cudaMemsetAsync(irr, 0, 4, s));
    IKD_LAUNCH voxel_decide_kernel<<<nblk(G, 128), 128, 0, s>>>(c, pts, idx_a, seg_begin, G, ds, vo, vboxes, irr);
    // 3. totals: delete boxes, survivors, acts
    int* del_pos = seg_begin + (n + 1);
    int* ins_pos = del_pos + (n + 1);
    // transform + exclusive scan via cub transform iterators
    {
        auto itd = thrust::make_transform_iterator((const VoxOut*)vo, VoxDel());
        auto iti = thrust::make_transform_iterator((const VoxOut*)vo, VoxIns());
        auto ita = thrust::make_transform_iterator((const VoxOut*)vo, VoxAct());
        size_t tmp = 0, t2 = 0;
        IKD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, itd, del_pos, G, s));
        IKD_CUDA(cub::DeviceReduce::Sum(nullptr, t2, ita, (int*)nullptr, G, s));
        IKD_TRY(t->b_cubtmp.ensure(std::max(tmp, t2), s));
        size_t tb = t->b_cubtmp.bytes;
        IKD_CUDA(cub::DeviceScan::ExclusiveSum(t->b_cubtmp.p, tb, itd, del_pos, G, s));
        tb = t->b_cubtmp.bytes;
        IKD_CUDA(cub::DeviceScan::ExclusiveSum(t->b_cubtmp.p, tb, iti, ins_pos, G, s));
        tb = t->b_cubtmp.bytes;
        IKD_CUDA(cub::DeviceReduce::Sum(t->b_cubtmp.p, tb, ita, irr + 1, G, s));
    }
    int h_irr[2];
    IKD_TRY(d2h(t, h_irr, irr, 2));
    if (h_irr[0] && !force) { *irregular = 1; return IKD_OK; }
    *acts_out = h_irr[1];
    int last[2];
    VoxOut vlast;
    IKD_TRY(d2h(t, &last[0], del_pos + (G - 1), 1));
    IKD_TRY(d2h(t, &last[1], ins_pos + (G - 1), 1));
    IKD_TRY(d2h(t, &vlast, vo + (G - 1), 1));
    int ndel = last[0] + vlast.del_box, nins = last[1] + (vlast.kind ? 1 : 0);
    IKD_TRY(t->u[U_BOXES].ensure((size_t)std::max(ndel, 1) * 24, s));
    IKD_TRY(t->u[U_SURV].ensure((size_t)std::max(nins, 1) * sizeof(float4), s));
    IKD_TRY(t->u[U_SRC].ensure((size_t)std::max(nins, 1) * 4, s));
    IKD_LAUNCH voxel_apply_kernel<<<nblk(G), TPB, 0, s>>>(vo, G, vboxes, pts, t->pid_xyz.as<float4>(), del_pos, ins_pos,
                                               t->u[U_BOXES].as<float>(), t->u[U_SURV].as<float4>(),
                                               t->u[U_SRC].as<int32_t>(), src_base);
    if (src_host && nins > 0) {
        IKD_CUDA(cudaMemcpyAsync(src_host, t->u[U_SRC].p, (size_t)nins * 4, cudaMemcpyDeviceToHost, s));
        IKD_CUDA(cudaStreamSynchronize(s));
    }
    // 4. apply: downsample-delete the boxes, then insert the survivors
    if (ndel > 0) {
        int dummy;
        IKD_TRY(delete_boxes_dev(t, t->u[U_BOXES].as<float>(), ndel, true, &dummy));
    }
    if (nins > 0) IKD_TRY(insert_points(t, t->u[U_SURV].as<float4>(), nins));
    *nins_out = nins;
    return IKD_OK;
}

int add_downsample_range(ikd_tree* t, const float4* pts, int off, int n, int* acts, int64_t* nins, int32_t* src_host) {
    if (n <= 0) return IKD_OK;
    int irregular = 0, a = 0;
    int64_t k = 0;
    // a single point is processed exactly as the reference does whatever its geometry (one voxel group cannot
    // conflict with itself), so pieces of size 1 are forced through
    IKD_TRY(add_downsample_piece(t, pts + off, n, off, n == 1, &a, &k, src_host ? src_host + *nins : nullptr, &irregular));
    if (!irregular) {
        *acts += a;
        *nins += k;
        return IKD_OK;
    }
    int h = n / 2;
    IKD_TRY(add_downsample_range(t, pts, off, h, acts, nins, src_host));
    IKD_TRY(add_downsample_range(t, pts, off + h, n - h, acts, nins, src_host));
    return IKD_OK;
}
}  // namespace

int add_points_dev_impl(ikd_tree* t, const float4* pts_dev, int64_t n, int downsample_on, int* out_added,
                        int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src) {
    *out_added = 0;
    *out_first_id = t->next_pid;
    *out_ninserted = 0;
    if (n == 0) return IKD_OK;
    if (n > 0x3fffffff) { set_error("batch too large"); return IKD_ERR_ARG; }
    IKD_TRY(ensure_pid_cap(t, (int64_t)t->next_pid + n));
    if (!downsample_on) {
        IKD_TRY(insert_points(t, pts_dev, (int)n));
        *out_added = 0;  // the reference only counts inserts of the downsample branch (tmp_counter, :448 vs :472)
        *out_ninserted = n;
        if (out_src) for (int64_t i = 0; i < n; i++) out_src[i] = (int32_t)i;
        return IKD_OK;
    }
    int acts = 0;
    int64_t nins = 0;
    IKD_TRY(add_downsample_range(t, pts_dev, 0, (int)n, &acts, &nins, out_src));
    *out_added = acts;
    *out_ninserted = nins;
    return IKD_OK;
}

int add_points_impl(ikd_tree* t, const float* xyz, int64_t n, int64_t stride, int downsample_on, int* out_added,
                    int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src) {
    *out_added = 0;
    *out_first_id = t->next_pid;
    *out_ninserted = 0;
    if (n == 0) return IKD_OK;
    if (n > 0x3fffffff) { set_error("batch too large"); return IKD_ERR_ARG; }
    IKD_TRY(t->u[U_PTS].ensure((size_t)n * sizeof(float4), t->stream));
    IKD_TRY(upload_points_f4(t, xyz, n, stride, t->u[U_PTS].as<float4>(), 0, 1));
    return add_points_dev_impl(t, t->u[U_PTS].as<float4>(), n, downsample_on, out_added, out_first_id, out_ninserted, out_src);
}

int add_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb) {
    (void)t; (void)boxes_host; (void)nb;
    set_error("Add_Point_Boxes is not implemented yet (SURVEY 8f next #1)");
    return IKD_ERR_INTERNAL;
}

// Pre-order structure dump for parity tests (columns as oracle/ref_harness.cpp ref_dump_tree).
int dump_tree_impl(ikd_tree* t, float* out, int64_t cap, int64_t* out_n) {
    IKD_TRY(sync_header(t));
    *out_n = 0;
    if (!t->hdr.root_exists) return IKD_OK;
    size_t used = t->hdr.pool_top;
    std::vector<SearchRec> sr(used);
    std::vector<UpdateRec> ur(used);
    IKD_CUDA(cudaMemcpy(sr.data(), t->srec, used * sizeof(SearchRec), cudaMemcpyDeviceToHost));
    IKD_CUDA(cudaMemcpy(ur.data(), t->urec, used * sizeof(UpdateRec), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> stack;
    stack.push_back(ROOT_SLOT);
    int64_t k = 0;
    while (!stack.empty()) {
        uint32_t s = stack.back();
        stack.pop_back();
        const SearchRec& a = sr[s];
        const UpdateRec& u = ur[s];
        uint32_t cp = meta_cp(a.meta);
        bool hl = cp && (ur[2 * cp].flags & F_EXISTS), hr = cp && (ur[2 * cp + 1].flags & F_EXISTS);
        if (k < cap) {
            float* o = out + 16 * k;
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = (float)meta_axis(a.meta);
            o[4] = (float)u.size; o[5] = (float)u.invalid;
            o[6] = (float)(((u.flags & F_PDEL) ? 1 : 0) | ((u.flags & F_TDEL) ? 2 : 0) | ((u.flags & F_PDS) ? 4 : 0) |
                           ((u.flags & F_TDS) ? 8 : 0));
            o[7] = u.bmin[0]; o[8] = u.bmax[0]; o[9] = u.bmin[1]; o[10] = u.bmax[1]; o[11] = u.bmin[2]; o[12] = u.bmax[2];
            o[13] = hl ? 1.f : 0.f; o[14] = hr ? 1.f : 0.f; o[15] = (float)u.down_del;
        }
        k++;
        if (hr) stack.push_back(2 * cp + 1);
        if (hl) stack.push_back(2 * cp);
    }
    *out_n = k;
    return IKD_OK;
}

}  // namespace ikd

using namespace ikd;

#define CHECK_T2(t)                                                        \
    do {                                                                   \
        if (!(t)) { set_error("null tree handle"); return IKD_ERR_ARG; }   \
        IKD_CUDA(cudaSetDevice((t)->device));                              \
        IKD_CUDA(cudaStreamSynchronize((t)->side));                        \
    } while (0)

extern "C" {

int ikd_add_points(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes, int downsample_on, int* out_added,
                   int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src) {
    CHECK_T2(t);
    if (n < 0 || (n > 0 && !xyz) || stride_bytes < 12 || !out_added || !out_first_id || !out_ninserted) {
        set_error("bad add_points arguments");
        return IKD_ERR_ARG;
    }
    return add_points_impl(t, xyz, n, stride_bytes, downsample_on, out_added, out_first_id, out_ninserted, out_src);
}

int ikd_add_points_dev(ikd_tree* t, const void* pts_dev_float4, int64_t n, int downsample_on, int* out_added,
                       int32_t* out_first_id, int64_t* out_ninserted, int32_t* out_src) {
    CHECK_T2(t);
    if (n < 0 || (n > 0 && !pts_dev_float4) || !out_added || !out_first_id || !out_ninserted) {
        set_error("bad add_points_dev arguments");
        return IKD_ERR_ARG;
    }
    return add_points_dev_impl(t, (const float4*)pts_dev_float4, n, downsample_on, out_added, out_first_id, out_ninserted,
                               out_src);
}

int ikd_delete_points(ikd_tree* t, const float* xyz, int64_t n, int64_t stride_bytes) {
    CHECK_T2(t);
    if (n < 0 || (n > 0 && !xyz) || stride_bytes < 12) { set_error("bad delete_points arguments"); return IKD_ERR_ARG; }
    return delete_points_impl(t, xyz, n, stride_bytes);
}

int ikd_delete_boxes(ikd_tree* t, const float* boxes, int64_t nb, int* out_deleted) {
    CHECK_T2(t);
    if (nb < 0 || (nb > 0 && !boxes) || !out_deleted) { set_error("bad delete_boxes arguments"); return IKD_ERR_ARG; }
    return delete_boxes_impl(t, boxes, nb, out_deleted);
}

int ikd_add_boxes(ikd_tree* t, const float* boxes, int64_t nb) {
    CHECK_T2(t);
    if (nb < 0 || (nb > 0 && !boxes)) { set_error("bad add_boxes arguments"); return IKD_ERR_ARG; }
    return add_boxes_impl(t, boxes, nb);
}

int ikd_flatten(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n) {
    CHECK_T2(t);
    if (!out_n || cap < 0) { set_error("bad flatten arguments"); return IKD_ERR_ARG; }
    return flatten_impl(t, out_idx, cap, out_n);
}

int ikd_acquire_removed(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n) {
    CHECK_T2(t);
    if (!out_n || cap < 0) { set_error("bad acquire_removed arguments"); return IKD_ERR_ARG; }
    return acquire_removed_impl(t, out_idx, cap, out_n);
}

}  // extern "C"
