// Mutating operations: Add_Points (with voxel downsample), Delete_Points, Delete_Point_Boxes, the
// Update pull-up ("refit"), Criterion_Check + Rebuild, flatten. Reference: ikd_Tree.cpp:414-489,
// :514-556, :625-760, :818-866, :1090-1107, :1184-1352.
//
// The reference mutates one point at a time and repairs the path it walked (Update, Criterion_Check,
// Rebuild on the way back up). Here every public call is a BATCH:
//   1. kernels apply the whole batch (flag bits set with atomics / new subtrees written), recording
//      the node slots they touched;
//   2. refit: the ancestors of the touched slots are marked dirty and recomputed bottom-up in one
//      kernel (the last-arriving child continues to the parent). That is Update() applied once per
//      dirty node instead of once per point. Criterion_Check is evaluated on every dirty node with
//      "effective" sizes, i.e. the sizes its children will have after the rebuilds already decided
//      below it, which is what the reference sees because it rebuilds children before it checks the
//      parent (:704-707);
//   3. the topmost violating nodes are rebuilt together (pre-order flatten with exact offsets, then
//      the forest builder of ikd_build.cu); their ancestors already carry the post-rebuild criteria
//      and boxes and adopt the post-rebuild sizes in one flat kernel (no second refit pass).
// Deletes are eager (the flag is written on every affected node), so there is no Push_Down and
// searches never mutate. "tree_deleted" still exists as a derived bit and makes searches skip dead
// subtrees through inverted child boxes.
// Host <-> device round trips are kept to a few small reads per call (counts the host needs to size
// the next launch), done by polling mapped pinned memory (fetch_small); everything else is enqueued
// back to back on the tree's stream, with independent pieces on helper streams.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "ikd_host.h"

namespace ikd {

namespace {

constexpr int TPB = 256;
constexpr int MAX_GRID = 148 * 8;
inline int nblk(int64_t n, int tpb = TPB) { return (int)std::max<int64_t>(1, (n + tpb - 1) / tpb); }
inline int sgrid(int64_t n, int tpb = TPB) { return std::min(nblk(n, tpb), MAX_GRID); }  // grid-stride launches

enum {
    U_CHANGED = 0, U_DIRTY, U_START, U_ROOTS, U_RINFO, U_STACK, U_P4, U_EROOT, U_FOREST, U_BOXES, U_PTS, U_KEYS, U_KEYS2,
    U_IDX, U_IDX2, U_GROUP, U_GINFO, U_VOX, U_ALIVE, U_SEL, U_CNT, U_TMP, U_TMP2, U_SURV, U_SRC, U_K64A, U_K64B, U_HT, U_NEXT
};

// device-side counters of the update path (one small struct, read back in one copy)
constexpr int CHAIN_SURV = 16;    // chain slots of the survivor compaction (<= 65536 / 4096 blocks)
constexpr int CHAIN_INS = 256;    // chain slots of the insert grouping (<= 65536 / 256 blocks)
struct Counters {
    // ---- reset by ONE memset at the start of every mutating operation (begin_changes) ----
    unsigned int nchanged, ndirty, nroots, nroots_big;
    unsigned long long delcount;
    int err, irregular;
    int maxseg, G, acts, ndel, nins, R_ins, B_ins, oor;
    unsigned long long chain_surv[CHAIN_SURV];  // single-pass chained scans (chain_base) of the scan-sized kernels
    unsigned long long chain_ins[CHAIN_INS];
    unsigned int collect_done;  // blocks of collect_plan_kernel that have finished their part of the dirty list
    unsigned int pad_cd;
    // ---- kept across operations ----
    unsigned int nremoved;  // removed-point log (acquire_removed_points)
    int pad0;
    // work counters of Add_Points (only touched while visit counting is on; read by ikd_get_stats)
    unsigned long long vox_visits;   // sum over input points of the nodes their voxel's box search visited
    unsigned long long desc_levels;  // sum over inserted points of the levels descended
};
static_assert(offsetof(Counters, chain_surv) == 64, "Counters layout");

struct Ctx {
    SearchRec* srec;
    UpdateRec* urec;
    TreeHeader* hdr;
    WalkRec* wrec;
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
#define GRID_STRIDE(i, n) for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

// ================================================================================================
// refit (Update, ikd_Tree.cpp:1184-1323, + Criterion_Check :1090-1107)
// ================================================================================================
// UpdateRec.pending while a refit is being prepared / run: -1 = clean, otherwise (dirty children at mark time << 16) |
// (dirty children not yet refit). The upper half is final once mark_kernel has finished, so the refit can tell an only
// dirty child (hand-over to the parent without fence and atomic) from one of two.
constexpr int PEND_ONE = 0x10001;
__global__ void mark_kernel(Ctx c, const int32_t* __restrict__ changed, Counters* __restrict__ k,
                            int32_t* __restrict__ dirty) {
    pdl_wait();
    const unsigned int nch = k->nchanged;
    GRID_STRIDE(i, nch) {
        int n = changed[i];
        if (n <= 0) continue;
        if (atomicCAS(&c.urec[n].pending, -1, 0) != -1) continue;  // already dirty: its marker walks the ancestors
        dirty[atomicAdd(&k->ndirty, 1u)] = n;
        int p = c.urec[n].parent;
        while (p) {
            // one memory round trip per level: the grandparent index is fetched next to the atomic that decides
            // whether the walk goes on
            const int gp = __ldcg(&c.urec[p].parent);
            const int prev = atomicAdd(&c.urec[p].pending, PEND_ONE);
            if (prev != -1) break;                   // p was dirty already: whoever made it so walks on from there
            atomicAdd(&c.urec[p].pending, 1);        // -1 + PEND_ONE + 1 == PEND_ONE (nobody waits for this one)
            dirty[atomicAdd(&k->ndirty, 1u)] = p;
            p = gp;
        }
    }
}

// Update() for one node (ikd_Tree.cpp:1184-1323) + Criterion_Check (:1090-1107), from records the caller has loaded:
// `a` = first 16 bytes of the node's SearchRec, `u` = its UpdateRec (updated in place and stored with pending = -1),
// `c0` / `c1` = the UpdateRecs of the two child slots (flags == 0: no child).
__device__ __forceinline__ void recompute_core(const Ctx& c, int n, float4 a, UpdateRec& u, const UpdateRec& c0,
                                               const UpdateRec& c1, float del_param, float bal_param) {
    uint32_t meta = __float_as_uint(a.w);
    const uint32_t cp = meta_cp(meta);
    const bool pdel = (u.flags & F_PDEL) != 0, pds = (u.flags & F_PDS) != 0;
    int size = 1, invalid = pdel ? 1 : 0, dd = pds ? 1 : 0;
    int esize = 1, einvalid = pdel ? 1 : 0;
    bool tds = pds, tdel = pdel;
    bool cex[2] = {false, false}, ctdel[2] = {false, false};
    float cmn[2][3], cmx[2][3];
    int cesize[2] = {0, 0};
    if (cp) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const UpdateRec& ch = s == 0 ? c0 : c1;
            // a child that fails the criteria is about to be rebuilt (now, or on the side stream): it will hold
            // exactly its valid points and no downsample-deleted ones; with no valid point left it vanishes
            // (BuildTree on an empty range, :575) and is treated as absent here already
            const bool cviol = (ch.flags & F_VIOL) != 0;
            // size / invalid stay PHYSICAL node counts until the rebuild has happened (they bound the flatten
            // stack and the visited list of a rebuild rooted above this node); adopt_effective_kernel switches
            // the surviving ancestors to the effective values afterwards
            if (ch.flags & F_EXISTS) { size += ch.size; invalid += ch.invalid; }
            if ((ch.flags & F_EXISTS) && !(cviol && ch.eff_size == 0)) {
                cex[s] = true;
                cesize[s] = ch.eff_size;
                dd += cviol ? 0 : ch.down_del;
                esize += ch.eff_size; einvalid += ch.eff_invalid;
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
    // Criterion_Check (:1090-1107) on effective sizes; a child that will vanish counts as absent
    bool viol = false;
    if (esize > 10) {
        int son = cesize[0] > 0 ? cesize[0] : cesize[1];
        float de = (float)einvalid / (float)esize;
        float be = (float)son / (float)(esize - 1);
        if (de > del_param) viol = true;
        if (be > bal_param || be < 1.0f - bal_param) viol = true;
    }
    uint32_t fl = u.flags & ~(F_TDEL | F_TDS | F_VIOL | F_SOLO);
    if (tdel) fl |= F_TDEL;
    if (tds) fl |= F_TDS;
    if (viol) fl |= F_VIOL;
    u.flags = fl;
    u.size = size; u.invalid = invalid; u.down_del = dd;
    u.eff_size = viol ? (esize - einvalid) : esize;   // a rebuild keeps exactly the valid points
    u.eff_invalid = viol ? 0 : einvalid;
    u.pending = -1;
#pragma unroll
    for (int k = 0; k < 3; k++) { u.bmin[k] = mn[k]; u.bmax[k] = mx[k]; }
    store_urec(c.urec + n, u);
    // search record: deleted bit + search-effective child boxes
    meta = pdel ? (meta | META_PDEL) : (meta & ~META_PDEL);
    meta = (meta & ~(META_LEX | META_REX)) | (cex[0] ? META_LEX : 0u) | (cex[1] ? META_REX : 0u);
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
    // walk record (enumeration-only walks of range searches): children as they will be once the rebuilds decided
    // below are done (a vanishing child is already absent), deleted bit, id
    {
        // (a child is a leaf when it is a single node; `size` is the physical node count)
        int lw[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const UpdateRec& ch = s == 0 ? c0 : c1;
            lw[s] = (cex[s] && ch.size == 1) ? ((ch.flags & F_PDEL) ? W_LEAF_DEAD : ch.pid) : W_NOT_LEAF;
        }
        c.wrec[n] = make_uint4(walk_head(cp, cex[0], cex[1], pdel), (uint32_t)u.pid, (uint32_t)lw[0], (uint32_t)lw[1]);
    }
    float4* q = reinterpret_cast<float4*>(c.srec + n);
    q[0] = make_float4(a.x, a.y, a.z, __uint_as_float(meta));
    q[1] = make_float4(b[0], b[1], b[2], b[3]);
    q[2] = make_float4(b[4], b[5], b[6], b[7]);
    q[3] = make_float4(b[8], b[9], b[10], b[11]);
    if (u.parent == 0) {
        TreeHeader* h = c.hdr;
        h->root_exists = 1;
        h->root_searchable = tdel ? 0 : 1;
        // the header reports the tree as it is once the rebuilds decided in this pass are done (effective sizes)
        h->size = esize;
        h->invalid = einvalid;
#pragma unroll
        for (int k = 0; k < 3; k++) { h->range[k] = mn[k]; h->range[3 + k] = mx[k]; }
        if (esize > 3) {  // :1315-1321
            int son = cex[0] ? cesize[0] : cesize[1];
            float tb = (float)son / (float)(esize - 1);
            h->alpha_del = (float)einvalid / (float)esize;
            h->alpha_bal = ((double)tb >= 0.5 - 1e-6) ? tb : 1.0f - tb;
        }
    }
}

// Bottom-up refit. A thread starts at a dirty node without dirty children and climbs: the last child to finish takes
// the parent. One memory round trip per level on the way up: the thread keeps the record it has just computed in
// registers and fetches the parent's two records and the SIBLING's record together (the sibling slot is n ^ 1), instead
// of loading the parent first and its two children afterwards. An only dirty child (mark_kernel's count in the upper
// half of `pending`) goes on without fence and atomic; one of two hands over through the counter, and the one that
// arrives last re-reads the sibling's record, which is final by then.
__global__ void refit_kernel(Ctx c, const int32_t* __restrict__ dirty, const Counters* __restrict__ k, float del_param,
                             float bal_param) {
    pdl_wait();
    const unsigned int nd = k->ndirty;
    GRID_STRIDE(i, nd) {
        int n = dirty[i];
        UpdateRec u = load_urec_cg(c.urec + n);
        // Starters are the dirty nodes that had NO dirty child when mark_kernel finished (upper half of `pending`). The
        // lower half must not be used here: it also reaches 0 when the last of two children hands over, a moment before
        // that child's thread (not this one) recomputes the node. A finished node reads -1.
        if ((u.pending >> 16) != 0) continue;
        float4 a = __ldcg(reinterpret_cast<const float4*>(c.srec + n));
        UpdateRec c0, c1;
        c0.flags = 0; c1.flags = 0;
        {
            const uint32_t cp = meta_cp(__float_as_uint(a.w));
            if (cp) { c0 = load_urec_cg(c.urec + 2 * cp); c1 = load_urec_cg(c.urec + 2 * cp + 1); }
        }
        while (true) {
            recompute_core(c, n, a, u, c0, c1, del_param, bal_param);
            const int p = u.parent;
            if (p == 0) break;
            const float4 ap = __ldcg(reinterpret_cast<const float4*>(c.srec + p));
            UpdateRec up = load_urec_cg(c.urec + p);
            UpdateRec us = load_urec_cg(c.urec + (n ^ 1));
            if ((up.pending >> 16) != 1) {
                __threadfence();
                const int old = atomicSub(&c.urec[p].pending, 1);
                if ((old & 0xffff) != 1) break;  // the sibling subtree is still being refit; its thread will take the parent
                us = load_urec_cg(c.urec + (n ^ 1));
            }
            if (n & 1) { c1 = u; c0 = us; } else { c0 = u; c1 = us; }
            n = p; a = ap; u = up;
        }
    }
}

// Per-root sizes, the three exclusive scans (point segments, flatten stacks, node blocks) and the totals the host needs,
// written into plan_out[] (a member of the header, so that one header read fetches them). Called by all NT threads of ONE block.
template <int NT>
__device__ void plan_block(const Ctx& c, const int32_t* roots, int R, const Counters* k, int* __restrict__ seg_begin,
                           int* __restrict__ soff, int* __restrict__ boff, int* __restrict__ plan_out) {
    typedef cub::BlockScan<int, NT> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int carry[3];
    __shared__ int smax, sroot, sdepth;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) { carry[0] = carry[1] = carry[2] = 0; smax = 0; sroot = 0; sdepth = 0; }
    __syncthreads();
    for (int base = 0; base < R; base += NT) {
        int r = base + tid;
        int nv = 0, ts = 0, bs = 0;
        if (r < R) {
            int s = __ldcg(roots + r);  // written by other blocks of this launch
            const UpdateRec& u = c.urec[s];
            nv = u.size - u.invalid;
            ts = u.size;
            bs = nv >= 2 ? (1 << (32 - __clz(nv))) : 0;
            atomicMax(&smax, nv);
            if (nv > 0) atomicMax(&sdepth, u.depth + (32 - __clz(nv)) - 1);  // depth bound after the rebuild (forest_depth_kernel)
            if (s == ROOT_SLOT) sroot = 1;
        }
        int o0, o1, o2, t0, t1, t2;
        Scan(tmp).ExclusiveSum(nv, o0, t0);
        __syncthreads();
        Scan(tmp).ExclusiveSum(ts, o1, t1);
        __syncthreads();
        Scan(tmp).ExclusiveSum(bs, o2, t2);
        __syncthreads();
        if (r < R) { seg_begin[r] = carry[0] + o0; soff[r] = carry[1] + o1; boff[r] = carry[2] + o2; }
        __syncthreads();
        if (tid == 0) { carry[0] += t0; carry[1] += t1; carry[2] += t2; }
        __syncthreads();
    }
    if (tid == 0) {
        seg_begin[R] = carry[0]; soff[R] = carry[1]; boff[R] = carry[2];
        int* p = plan_out;
        p[0] = R; p[1] = carry[0]; p[2] = carry[1]; p[3] = carry[2]; p[4] = smax; p[5] = sroot; p[6] = (int)__ldcg(&k->ndirty); p[7] = sdepth;
    }
    __syncthreads();
}

// Topmost violating nodes among the dirty set -> rebuild roots. Subtrees with at least `async_min` valid points go to a
// second list (rebuilt on the side stream) and are flagged F_ASYNC so that later passes leave them alone. The block that
// finishes LAST then plans the rebuilds of both lists (plan_block) and hands the whole header (root counters from the refit,
// plan[] / plan2[]) to the host -- one launch instead of three (collect, plan, plan of the side-stream list).
struct PlanArrays { int* seg_begin; int* soff; int* boff; int* out; };
__global__ void __launch_bounds__(TPB)
collect_plan_kernel(Ctx c, const int32_t* __restrict__ dirty, Counters* __restrict__ k, int32_t* __restrict__ roots,
                    int32_t* __restrict__ roots_big, int async_min, PlanArrays pa, PlanArrays pb, PublishTicket pub) {
    pdl_wait();
    const unsigned int nd = k->ndirty;
    GRID_STRIDE(i, nd) {
        int n = dirty[i];
        uint32_t fl = c.urec[n].flags;
        if (!(fl & F_VIOL) || (fl & F_ASYNC)) continue;
        int p = c.urec[n].parent;
        bool top = true;
        while (p) {
            if (c.urec[p].flags & F_VIOL) { top = false; break; }
            p = c.urec[p].parent;
        }
        if (!top) continue;
        int nv = c.urec[n].size - c.urec[n].invalid;
        if (async_min > 0 && nv >= async_min && n != ROOT_SLOT) {
            roots_big[atomicAdd(&k->nroots_big, 1u)] = n;
            atomicOr(&c.urec[n].flags, F_ASYNC);
        } else {
            roots[atomicAdd(&k->nroots, 1u)] = n;
        }
    }
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&k->collect_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    plan_block<TPB>(c, roots, (int)__ldcg(&k->nroots), k, pa.seg_begin, pa.soff, pa.boff, pa.out);
    if (pb.out) plan_block<TPB>(c, roots_big, (int)__ldcg(&k->nroots_big), k, pb.seg_begin, pb.soff, pb.boff, pb.out);
    if (pub.dst) publish_words(c.hdr, (int)(sizeof(TreeHeader) / 4), pub.dst, pub.flag, pub.seq);
}

// After the rebuilds planned by a refit pass have been enqueued: the ancestors of the rebuilt subtrees take the
// sizes the pass already computed for them ("effective" = after the rebuilds). This replaces a second
// mark/refit round (a 25-level chain of dependent atomics, ~110 us) by one flat pass over the dirty list.
// (`nd` by value: on the side stream this kernel can run after the next operation has already reset the counters)
__global__ void adopt_effective_kernel(Ctx c, const int32_t* __restrict__ dirty, unsigned int nd) {
    pdl_wait();
    GRID_STRIDE(i, nd) {
        UpdateRec* u = c.urec + dirty[i];
        const uint32_t fl = u->flags;
        if (!(fl & F_EXISTS) || (fl & F_VIOL)) continue;  // released by a rebuild / waiting for the side stream
        const int es = u->eff_size, ei = u->eff_invalid;
        if (u->size != es) u->size = es;
        if (u->invalid != ei) u->invalid = ei;
    }
}

// ================================================================================================
// rebuild (Rebuild :625-645, flatten :1326-1352)
// ================================================================================================
__global__ void set_pool_top_kernel(TreeHeader* hdr, unsigned int v) { hdr->pool_top = v; }

__global__ void forest_setup_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ seg_begin,
                                    const int* __restrict__ boff, unsigned int pool_base, int* __restrict__ root_slot,
                                    int* __restrict__ block_base, int* __restrict__ root_parent,
                                    int* __restrict__ root_depth, int* __restrict__ single_axis, unsigned int new_pool_top) {
    pdl_wait();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) c.hdr->pool_top = new_pool_top;  // the host mirror already has it
    if (r >= R) return;
    int s = roots[r];
    UpdateRec u = c.urec[s];
    root_slot[r] = s;
    block_base[r] = (int)pool_base + boff[r];
    root_parent[r] = u.parent;
    root_depth[r] = u.depth;
    single_axis[r] = -1;
    if (seg_begin[r + 1] == seg_begin[r]) {
        // no valid point left: the subtree vanishes (BuildTree on an empty range leaves *root null, :575)
        c.urec[s].flags = 0;
        c.urec[s].pending = -1;
        c.urec[s].size = 0; c.urec[s].invalid = 0; c.urec[s].eff_size = 0; c.urec[s].eff_invalid = 0;
        if (u.parent == 0) { c.hdr->root_exists = 0; c.hdr->root_searchable = 0; c.hdr->size = 0; c.hdr->invalid = 0; }
    }
}

// One block per rebuild root: pre-order flatten of the valid points with exact output offsets
// (offset of a node = offset of its parent + [parent valid] (+ valid count of the left sibling)),
// so the point order is the reference's flatten order without atomics. Old nodes are released.
constexpr int FL_TPB = 256;
template <int NT>
__global__ void __launch_bounds__(NT)
flatten_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ seg_begin,
               const int* __restrict__ stack_off, uint2* __restrict__ stack_mem, float4* __restrict__ p4,
               int* __restrict__ eroot, int32_t* __restrict__ removed, Counters* __restrict__ k,
               unsigned int removed_cap, bool emit, bool release, int32_t* __restrict__ visited,
               const int* __restrict__ limit_arr = nullptr, const int* __restrict__ root_of = nullptr) {
    // emit: write the valid points (rebuild input); release: free the old nodes and log removed points.
    // A synchronous rebuild does both at once; a side-stream rebuild emits first and also records every node it
    // visited (`visited`, ~slot for the root), so that the commit releases them with one flat kernel
    // (release_list_kernel) instead of walking the subtree a second time.
    typedef cub::BlockScan<int, NT> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_top, s_done;
    const int tid = threadIdx.x;
    pdl_wait();
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const int root = roots[r];
        if (root == 0) continue;  // unused sub-root entry of a split (uniform across the block)
        uint2* stack = stack_mem + stack_off[r];
        // node count of the subtree: bounds stack and visited list
        const int limit = limit_arr ? limit_arr[r] : stack_off[r + 1] - stack_off[r];
        const int out_root = root_of ? root_of[r] : r;  // subtree index the builder sees
        __syncthreads();
        if (tid == 0) { stack[0] = make_uint2((unsigned)root, (unsigned)seg_begin[r]); s_top = 1; s_done = 0; }
        __syncthreads();
        while (true) {
            int top = s_top;
            if (top == 0) break;
            const int done = s_done;
            int take = top < NT ? top : NT;
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
                    if (emit) {
                        p4[off] = make_float4(a.x, a.y, a.z, __int_as_float(u.pid));
                        eroot[off] = out_root;
                    }
                } else if (release && !(u.flags & F_PDS)) {
                    unsigned int q = atomicAdd(&k->nremoved, 1u);  // Points_deleted (:1339-1341)
                    if (q < removed_cap) removed[q] = u.pid;
                }
                if (visited) {
                    if (done + tid < limit) visited[stack_off[r] + done + tid] = (slot == root && !root_of) ? ~slot : slot;
                    else c.hdr->flag1 = 1;  // size bookkeeping broken: reported by the next header read
                }
                uint32_t cp = meta_cp(__float_as_uint(a.w));
                int coff = off + (valid ? 1 : 0);
                if (cp) {
                    const UpdateRec& L = c.urec[2 * cp];
                    const UpdateRec& Rr = c.urec[2 * cp + 1];
                    if (L.flags & F_EXISTS) { pu[npush++] = make_uint2(2 * cp, (unsigned)coff); coff += L.size - L.invalid; }
                    if (Rr.flags & F_EXISTS) { pu[npush++] = make_uint2(2 * cp + 1, (unsigned)coff); }
                }
                if (release && slot != root) { c.urec[slot].flags = 0; c.urec[slot].pending = -1; }
            }
            int pos, total;
            Scan(tmp).ExclusiveSum(npush, pos, total);
            int base = top - take;
            if (base + pos + npush > limit) { c.hdr->flag1 = 1; npush = 0; }
            if (npush >= 1) stack[base + pos] = pu[0];
            if (npush == 2) stack[base + pos + 1] = pu[1];
            __syncthreads();
            if (tid == 0) { s_top = base + total; s_done = done + take; }
            __syncthreads();
        }
    }
}

// Side-stream rebuilds are few and large: one block per root would walk a 76k-node subtree in ~100 rounds of
// dependent loads (135 us measured). This kernel walks only the top SPLIT_LEVELS levels of every root (emitting those
// nodes) and hands the up to 64 subtrees below to flatten_kernel as independent sub-roots, each with its exact output
// offset and its own slice of the stack / visited regions, so the rest of the walk runs on up to 64 SMs per root.
constexpr int SPLIT_LEVELS = 6;
constexpr int SPLIT_MAX = 1 << SPLIT_LEVELS;
__global__ void __launch_bounds__(SPLIT_MAX)
split_roots_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ seg_begin,
                   const int* __restrict__ stack_off, float4* __restrict__ p4, int* __restrict__ eroot,
                   int32_t* __restrict__ visited, int32_t* __restrict__ sub_root, int* __restrict__ sub_seg,
                   int* __restrict__ sub_stack, int* __restrict__ sub_limit, int* __restrict__ sub_of) {
    __shared__ uint2 cur[SPLIT_MAX], nxt[SPLIT_MAX];
    __shared__ int ncur, nnxt, nvis;
    const int r = blockIdx.x, tid = threadIdx.x;
    if (r >= R) return;
    const int root = roots[r];
    const int vbase = stack_off[r];
    if (tid == 0) { cur[0] = make_uint2((unsigned)root, (unsigned)seg_begin[r]); ncur = 1; nnxt = 0; nvis = 0; }
    __syncthreads();
    for (int lv = 0; lv < SPLIT_LEVELS; lv++) {
        const int n = ncur;
        if (tid < n) {
            const int slot = (int)cur[tid].x, off = (int)cur[tid].y;
            float4 a = reinterpret_cast<const float4*>(c.srec + slot)[0];
            UpdateRec u = c.urec[slot];
            const bool valid = !(u.flags & F_PDEL);
            if (valid) { p4[off] = make_float4(a.x, a.y, a.z, __int_as_float(u.pid)); eroot[off] = r; }
            visited[vbase + atomicAdd(&nvis, 1)] = slot == root ? ~slot : slot;
            uint32_t cp = meta_cp(__float_as_uint(a.w));
            int coff = off + (valid ? 1 : 0);
            if (cp) {
                const UpdateRec& L = c.urec[2 * cp];
                const UpdateRec& Rr = c.urec[2 * cp + 1];
                if (L.flags & F_EXISTS) { nxt[atomicAdd(&nnxt, 1)] = make_uint2(2 * cp, (unsigned)coff); coff += L.size - L.invalid; }
                if (Rr.flags & F_EXISTS) nxt[atomicAdd(&nnxt, 1)] = make_uint2(2 * cp + 1, (unsigned)coff);
            }
        }
        __syncthreads();
        if (tid < nnxt) cur[tid] = nxt[tid];
        __syncthreads();
        if (tid == 0) { ncur = nnxt; nnxt = 0; }
        __syncthreads();
    }
    // the remaining frontier becomes the sub-roots; their regions follow the top nodes inside this root's region
    const int n = ncur;
    if (tid == 0) {
        int pos = vbase + nvis;
        for (int i = 0; i < n; i++) {
            const int slot = (int)cur[i].x;
            const int sz = c.urec[slot].size;
            const int o = r * SPLIT_MAX + i;
            sub_root[o] = slot; sub_seg[o] = (int)cur[i].y; sub_stack[o] = pos; sub_limit[o] = sz; sub_of[o] = r;
            pos += sz;
        }
        for (int i = n; i < SPLIT_MAX; i++) sub_root[r * SPLIT_MAX + i] = 0;
    }
}

// Commit of a side-stream rebuild: release the old nodes recorded by the emit pass and log the removed points.
__global__ void release_list_kernel(Ctx c, const int32_t* __restrict__ visited, int n, int32_t* __restrict__ removed,
                                    Counters* __restrict__ k, unsigned int removed_cap) {
    pdl_wait();
    GRID_STRIDE(i, (unsigned)n) {
        int v = visited[i];
        const bool is_root = v < 0;
        const int slot = is_root ? ~v : v;
        const uint32_t fl = c.urec[slot].flags;
        if ((fl & F_PDEL) && !(fl & F_PDS)) {
            unsigned int q = atomicAdd(&k->nremoved, 1u);  // Points_deleted (:1339-1341)
            if (q < removed_cap) removed[q] = c.urec[slot].pid;
        }
        if (!is_root) { c.urec[slot].flags = 0; c.urec[slot].pending = -1; }
    }
}

// Side-stream rebuild, step 1 (forest description): like forest_setup_kernel, but the new subtree gets a TEMPORARY
// root slot (the unused slot 1 of its own node block), so nothing reachable from the live tree is written.
__global__ void forest_setup_async_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ boff,
                                          unsigned int pool_base, int* __restrict__ root_slot, int* __restrict__ block_base,
                                          int* __restrict__ root_parent, int* __restrict__ root_depth,
                                          int* __restrict__ single_axis) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const UpdateRec& u = c.urec[roots[r]];
    int base = (int)pool_base + boff[r];
    block_base[r] = base;
    root_slot[r] = base + 1;
    root_parent[r] = u.parent;
    root_depth[r] = u.depth;
    single_axis[r] = -1;
}

// Side-stream rebuild, step 2 (commit, on the main stream once the build has finished and the old nodes have been
// released): move the new root record into the old root slot -- the parent's child-pair link stays valid -- and
// re-parent its two children (the reference swaps the subtree pointer in the father, ikd_Tree.cpp:277-285).
__global__ void commit_async_kernel(Ctx c, const int32_t* __restrict__ roots, int R, const int* __restrict__ new_root,
                                    int32_t* __restrict__ changed, Counters* __restrict__ k) {
    pdl_wait();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int old = roots[r], nw = new_root[r];
    const float4* sn = reinterpret_cast<const float4*>(c.srec + nw);
    float4* so = reinterpret_cast<float4*>(c.srec + old);
    float4 a = sn[0], b = sn[1], cc = sn[2], d = sn[3];
    so[0] = a; so[1] = b; so[2] = cc; so[3] = d;
    UpdateRec u = c.urec[nw];
    u.parent = c.urec[old].parent;
    u.depth = c.urec[old].depth;
    store_urec(c.urec + old, u);
    c.wrec[old] = c.wrec[nw];
    uint32_t cp = meta_cp(__float_as_uint(a.w));
    if (cp) {
        if (c.urec[2 * cp].flags & F_EXISTS) c.urec[2 * cp].parent = old;
        if (c.urec[2 * cp + 1].flags & F_EXISTS) c.urec[2 * cp + 1].parent = old;
    }
    c.urec[nw].flags = 0;
    c.urec[nw].pending = -1;
    changed[atomicAdd(&k->nchanged, 1u)] = old;
}

// alive[pid] = 1 for every valid point; logs removed points (whole-tree rebuild / flatten export)
__global__ void alive_kernel(Ctx c, unsigned int pool_top, uint8_t* __restrict__ alive, bool log_removed,
                             int32_t* __restrict__ removed, Counters* __restrict__ k, unsigned int removed_cap) {
    GRID_STRIDE(s, pool_top) {
        if (s == 0) continue;
        const UpdateRec& u = c.urec[s];
        if (!(u.flags & F_EXISTS)) continue;
        if (!(u.flags & F_PDEL)) alive[u.pid] = 1;
        else if (log_removed && !(u.flags & F_PDS)) {
            unsigned int q = atomicAdd(&k->nremoved, 1u);
            if (q < removed_cap) removed[q] = u.pid;
        }
    }
}

// id compaction: element i = point with old id pids[i], new id i
__global__ void gather_newid_kernel(const int32_t* __restrict__ pids, int n, const float4* __restrict__ pid_xyz,
                                    float4* __restrict__ p4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = pid_xyz[pids[i]];
    p4[i] = make_float4(v.x, v.y, v.z, __int_as_float(i));
}
__global__ void scatter_xyz_kernel(const float4* __restrict__ p4, int n, float4* __restrict__ pid_xyz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = p4[i];
    pid_xyz[i] = make_float4(v.x, v.y, v.z, 0.f);
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
                                     Counters* __restrict__ k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !c.hdr->root_exists) return;
    float4 p = pts[i];
    uint32_t cur = ROOT_SLOT;
    while (cur) {
        uint32_t fl = __ldcg(&c.urec[cur].flags);
        if (!(fl & F_EXISTS)) return;  // stepped into an empty child position (noticed here: one round trip per level)
        if (fl & F_TDEL) return;  // :714
        float4 a = __ldcg(reinterpret_cast<const float4*>(c.srec + cur));
        uint32_t meta = __float_as_uint(a.w);
        if (same_point_d(a.x, a.y, a.z, p.x, p.y, p.z)) {
            uint32_t old = atomicOr(&c.urec[cur].flags, F_PDEL);
            if (!(old & F_PDEL)) {  // this thread deleted it (:717-722)
                atomicOr(&c.srec[cur].meta, META_PDEL);
                changed[atomicAdd(&k->nchanged, 1u)] = (int32_t)cur;
                return;
            }
        }
        int ax = meta_axis(meta);
        float pc = ax == 0 ? p.x : (ax == 1 ? p.y : p.z);
        float nc = ax == 0 ? a.x : (ax == 1 ? a.y : a.z);
        uint32_t cp = meta_cp(meta);
        if (!cp) return;
        cur = 2 * cp + (pc < nc ? 0u : 1u);
    }
}

// ================================================================================================
// Add_by_point as a bulk insert (:818-866)
// ================================================================================================
// Descend to the empty child position a point would be appended at (:818-866); key = parent slot * 2 + side.
// One 16-byte fetch per level (split coordinate, axis, child pair, which children exist): the lines the preceding kNN
// batch has pulled into L2; the UpdateRec array is not touched.
__device__ __forceinline__ uint32_t descend_to_insert_position(const Ctx& c, float4 p, unsigned int* levels) {
    uint32_t cur = ROOT_SLOT, key = 0;
    while (true) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(c.srec + cur));
        (*levels)++;
        const uint32_t meta = __float_as_uint(a.w);
        const int ax = meta_axis(meta);
        const float pc = ax == 0 ? p.x : (ax == 1 ? p.y : p.z);
        const float nc = ax == 0 ? a.x : (ax == 1 ? a.y : a.z);
        const uint32_t side = pc < nc ? 0u : 1u;  // :833
        key = cur * 2 + side;
        if (!(meta & (side ? META_REX : META_LEX))) break;  // empty child position
        cur = 2 * meta_cp(meta) + side;
    }
    return key;
}

__global__ void descend_kernel(Ctx c, const float4* __restrict__ pts, int n, uint32_t* __restrict__ keys,
                               int* __restrict__ idx, Counters* __restrict__ k, bool count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int levels = 0;
    keys[i] = descend_to_insert_position(c, pts[i], &levels);
    idx[i] = i;
    if (count) atomicAdd(&k->desc_levels, (unsigned long long)levels);
}

template <typename KeyT>
__global__ void head_flag_kernel(const KeyT* __restrict__ keys, int n, int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// after an inclusive scan of head flags: gid[i]-1 is the group of sorted element i
template <typename KeyT>
__global__ void group_bounds_kernel(const KeyT* __restrict__ keys, const int* __restrict__ gid, int n,
                                    int* __restrict__ seg_begin, KeyT* __restrict__ gkey, int* __restrict__ eroot,
                                    int* __restrict__ ngroups) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = gid[i] - 1;
    if (eroot) eroot[i] = g;
    if (i == 0 || keys[i] != keys[i - 1]) { seg_begin[g] = i; if (gkey) gkey[g] = keys[i]; }
    if (i == n - 1) { seg_begin[g + 1] = n; *ngroups = g + 1; }
}

// the first group of each distinct parent allocates the child pair if it is missing
__global__ void alloc_pairs_kernel(Ctx c, const uint32_t* __restrict__ gkey, const int* __restrict__ ngroups) {
    const int R = *ngroups;
    GRID_STRIDE(g, R) {
        uint32_t parent = gkey[g] >> 1;
        if (g > 0 && (gkey[g - 1] >> 1) == parent) continue;
        uint32_t meta = c.srec[parent].meta;
        if (meta_cp(meta)) continue;
        uint32_t slot = atomicAdd(&c.hdr->pool_top, 2u);
        uint32_t cp = slot >> 1;
        UpdateRec z;
        memset(&z, 0, sizeof(z));
        z.pending = -1;
        store_urec(c.urec + slot, z);
        store_urec(c.urec + slot + 1, z);
        c.srec[parent].meta = meta | (cp << META_CP_SHIFT);
    }
}

// Single block: node-block sizes of the insert groups, their exclusive scan and totals.
__global__ void __launch_bounds__(1024)
insert_plan_kernel(const int* __restrict__ seg_begin, Counters* __restrict__ k, int* __restrict__ boff) {
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int carry, smax;
    const int R = k->R_ins;
    const int tid = threadIdx.x;
    if (tid == 0) { carry = 0; smax = 0; }
    __syncthreads();
    for (int base = 0; base < R; base += 1024) {
        int g = base + tid;
        int bs = 0;
        if (g < R) {
            int n = seg_begin[g + 1] - seg_begin[g];
            bs = n >= 2 ? (1 << (32 - __clz(n))) : 0;
            atomicMax(&smax, n);
        }
        int o, tot;
        Scan(tmp).ExclusiveSum(bs, o, tot);
        __syncthreads();
        if (g < R) boff[g] = carry + o;
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) { boff[R] = carry; k->B_ins = carry; k->maxseg = smax; }
}

// ---- sort-free grouping for small batches: a hash table keyed by the group key links the members of a
// group into a list (head per table slot, next per element); the element that created the slot registers
// the group. Replaces radix sort + head flags + scan + bounds (a dozen launches) by one kernel.
constexpr unsigned long long HT_EMPTY = ~0ull;
struct HashTab {
    unsigned long long* keys;  // HT_EMPTY when free
    int* head;                 // -1 when empty
    uint32_t mask;
};
__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}
__device__ __forceinline__ uint32_t ht_find_or_insert(const HashTab& h, unsigned long long key, bool& created) {
    uint32_t s = hash64(key) & h.mask;
    while (true) {
        unsigned long long prev = atomicCAS(&h.keys[s], HT_EMPTY, key);
        if (prev == HT_EMPTY) { created = true; return s; }
        if (prev == key) { created = false; return s; }
        s = (s + 1) & h.mask;
    }
}

// descend (as descend_kernel) and link the point into the list of its target position
__global__ void descend_link_kernel(Ctx c, const float4* __restrict__ pts, int n, HashTab ht, int* __restrict__ next,
                                    int* __restrict__ slot_of, int* __restrict__ glist, Counters* __restrict__ k, bool count) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int levels = 0;
    const uint32_t key = descend_to_insert_position(c, pts[i], &levels);
    if (count) atomicAdd(&k->desc_levels, (unsigned long long)levels);
    // group bookkeeping without linked lists (walking them made the largest group's thread a chain of dependent
    // loads): the table slot counts its members (head = count - 1), every point keeps its slot and arrival number
    bool created;
    uint32_t slot = ht_find_or_insert(ht, (unsigned long long)key, created);
    next[i] = atomicAdd(&ht.head[slot], 1) + 1;  // arrival number inside the group
    slot_of[i] = (int)slot;
    if (created) glist[atomicAdd(&k->R_ins, 1)] = (int)slot;
}

// Sizes of the insert groups, child-pair allocation, the scans for point segments and node blocks (chained across
// the blocks of the launch). One thread per group; launched with enough blocks for the upper bound n of the
// group count. Leaves, per table slot, the group's number and the start of its point segment.
constexpr int IG_TPB = 256;
__global__ void __launch_bounds__(IG_TPB)
insert_group_kernel(Ctx c, HashTab ht, const int* __restrict__ glist, Counters* __restrict__ k,
                    int* __restrict__ seg_begin, uint32_t* __restrict__ gkey, int* __restrict__ boff,
                    int* __restrict__ slot_begin, int* __restrict__ slot_gid, unsigned long long* __restrict__ chain) {
    pdl_wait();
    typedef cub::BlockScan<unsigned long long, IG_TPB> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int smax;
    const int R = k->R_ins;
    const int tid = threadIdx.x;
    const int g = blockIdx.x * IG_TPB + tid;
    if (tid == 0) smax = 0;
    __syncthreads();
    int slot = 0, cnt = 0;
    unsigned long long v = 0;
    uint32_t parent = 0, meta = 0;
    bool need_pair = false;
    if (g < R) {
        slot = glist[g];
        unsigned long long key = ht.keys[slot];
        cnt = ht.head[slot] + 1;
        gkey[g] = (uint32_t)key;
        parent = (uint32_t)key >> 1;
        meta = __ldcg(&c.srec[parent].meta);
        need_pair = !meta_cp(meta);
        unsigned long long bs = cnt >= 2 ? (1ull << (32 - __clz(cnt))) : 0ull;
        v = ((unsigned long long)cnt << 32) | bs;
        atomicMax(&smax, cnt);
    }
    {
        // child pairs for leaf positions: one pool allocation per block (thousands of per-thread atomics on the one
        // pool_top word were serialising this kernel)
        typedef cub::BlockScan<int, IG_TPB> ScanI;
        __shared__ typename ScanI::TempStorage tmp_i;
        __shared__ uint32_t s_pair_base;
        int ppos, ptot;
        ScanI(tmp_i).ExclusiveSum(need_pair ? 1 : 0, ppos, ptot);
        if (tid == 0 && ptot > 0) s_pair_base = atomicAdd(&c.hdr->pool_top, 2u * (uint32_t)ptot);
        __syncthreads();
        if (need_pair) {
            uint32_t ns = s_pair_base + 2u * (uint32_t)ppos;
            UpdateRec z;
            memset(&z, 0, sizeof(z));
            z.pending = -1;
            store_urec(c.urec + ns, z);
            store_urec(c.urec + ns + 1, z);
            __threadfence();
            // the sibling position's group may install its pair first; then this one is simply left unused
            atomicCAS(&c.srec[parent].meta, meta, meta | ((ns >> 1) << META_CP_SHIFT));
        }
    }
    unsigned long long o, tot;
    Scan(tmp).ExclusiveSum(v, o, tot);
    const unsigned long long base = chain_base(chain, tot);
    const unsigned long long mine = base + o;
    if (tid == 0 && smax > 0) atomicMax(&k->maxseg, smax);
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const unsigned long long all = base + tot;
        seg_begin[R] = (int)(all >> 32);
        boff[R] = (int)(all & 0xffffffffu);
        k->B_ins = (int)(all & 0xffffffffu);
    }
    if (g >= R) return;
    seg_begin[g] = (int)(mine >> 32);
    boff[g] = (int)(mine & 0xffffffffu);
    slot_begin[slot] = (int)(mine >> 32);
    slot_gid[slot] = g;
}

// every point drops its index into its group's segment (arrival order) ...
__global__ void insert_scatter_kernel(int n, const int* __restrict__ arrival, const int* __restrict__ slot_of,
                                      const int* __restrict__ slot_begin, int* __restrict__ members) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    members[slot_begin[slot_of[i]] + arrival[i]] = i;
}
// ... and then takes the position of its index among the group's members (ascending input index: a stable order
// for coordinate ties; groups above INS_RANK_MAX members keep the arrival order, as the builder's sort is stable
// either way) and writes the point where the forest builder expects it.
constexpr int INS_RANK_MAX = 64;
__global__ void insert_place_kernel(int n, const float4* __restrict__ pts, HashTab ht, const int* __restrict__ arrival,
                                    const int* __restrict__ slot_of, const int* __restrict__ slot_begin,
                                    const int* __restrict__ slot_gid, const int* __restrict__ members, int first_pid,
                                    int* __restrict__ eroot, float4* __restrict__ p4, float4* __restrict__ pid_xyz,
                                    const Counters* __restrict__ k, const TreeHeader* __restrict__ hdr, PublishTicket pub) {
    pdl_wait();
    // the group counters and the pool top were final before this launch: block 0 hands them to the host
    if (blockIdx.x == 0 && pub.dst)
        publish_words(k, (int)(offsetof(Counters, chain_surv) / 4), pub.dst, pub.flag, pub.seq, &hdr->pool_top, 1);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = slot_of[i];
    const int b = slot_begin[slot];
    const int cnt = ht.head[slot] + 1;
    int rank = arrival[i];
    if (cnt > 1 && cnt <= INS_RANK_MAX) {
        rank = 0;
        for (int j = 0; j < cnt; j++) rank += members[b + j] < i ? 1 : 0;
    }
    float4 w = pts[i];
    int pid = first_pid + i;
    p4[b + rank] = make_float4(w.x, w.y, w.z, __int_as_float(pid));
    pid_xyz[pid] = make_float4(w.x, w.y, w.z, 0.f);
    eroot[b + rank] = slot_gid[slot];
}

__global__ void insert_forest_kernel(Ctx c, const uint32_t* __restrict__ gkey, int R, const int* __restrict__ boff,
                                     unsigned int pool_base, int* __restrict__ root_slot, int* __restrict__ block_base,
                                     int* __restrict__ root_parent, int* __restrict__ root_depth,
                                     int* __restrict__ single_axis, int32_t* __restrict__ changed,
                                     Counters* __restrict__ k, unsigned int new_pool_top) {
    pdl_wait();
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) c.hdr->pool_top = new_pool_top;  // the host mirror already has it
    if (g >= R) return;
    uint32_t parent = gkey[g] >> 1, side = gkey[g] & 1u;
    uint32_t meta = c.srec[parent].meta;
    uint32_t cp = meta_cp(meta);
    int slot = (int)(2 * cp + side);
    root_slot[g] = slot;
    root_parent[g] = (int)parent;
    root_depth[g] = c.urec[parent].depth + 1;
    single_axis[g] = (meta_axis(meta) + 1) % 3;  // :823
    block_base[g] = (int)pool_base + boff[g];
    changed[atomicAdd(&k->nchanged, 1u)] = slot;  // refit starts from the new subtree roots
}

__global__ void gather_sorted_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, int n, int first_pid,
                                     float4* __restrict__ p4, float4* __restrict__ pid_xyz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = idx ? idx[i] : i;
    float4 v = pts[j];
    int pid = first_pid + j;  // ids follow the order of the input list, not the sorted order
    p4[i] = make_float4(v.x, v.y, v.z, __int_as_float(pid));
    pid_xyz[pid] = make_float4(v.x, v.y, v.z, 0.f);
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

// Packed voxel key from the three floor indices, each taken relative to `org[a]` with `bits[a]` bits
// (chosen by the host from the tree's range plus a margin, so that few radix passes are needed);
// sets k->oor when an index does not fit and the wide three-pass grouping has to be used instead.
struct VoxPack { float org[3]; int bits[3]; };
__global__ void voxel_key64_kernel(const float4* __restrict__ pts, int n, float ds, VoxPack vp,
                                   unsigned long long* __restrict__ keys, int* __restrict__ idx, Counters* __restrict__ k) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pts[i];
    float f[3] = {floorf(__fdiv_rn(p.x, ds)), floorf(__fdiv_rn(p.y, ds)), floorf(__fdiv_rn(p.z, ds))};
    unsigned long long key = 0;
    bool bad = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float rel = f[a] - vp.org[a];  // exact: both are integers well below 2^24 whenever the test below passes
        float lim = (float)(1u << vp.bits[a]);
        if (!(rel >= 0.f && rel < lim && fabsf(f[a]) < 8388608.f)) { bad = true; rel = 0.f; }
        key = (key << vp.bits[a]) | (unsigned long long)rel;
    }
    if (bad) k->oor = 1;
    keys[i] = key;
    idx[i] = i;
}

// fallback grouping for out-of-range voxel indices: stable three-pass sort on the box corner bits
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
__global__ void voxel_head3_kernel(const uint32_t* __restrict__ kx, const uint32_t* __restrict__ ky,
                                   const uint32_t* __restrict__ kz, const int* __restrict__ idx, int n,
                                   int* __restrict__ head) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = idx[i];
    bool h = true;
    if (i > 0) { int b = idx[i - 1]; h = kx[a] != kx[b] || ky[a] != ky[b] || kz[a] != kz[b]; }
    head[i] = h ? 1 : 0;
}
__global__ void voxel_bounds3_kernel(const int* __restrict__ head, const int* __restrict__ gid, int n,
                                     int* __restrict__ seg_begin, int* __restrict__ ngroups) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = gid[i] - 1;
    if (head[i]) seg_begin[g] = i;
    if (i == n - 1) { seg_begin[g + 1] = n; *ngroups = g + 1; }
}

// a coordinate is "regular" for voxel index nf if it lies in box nf and in neither neighbour box, so the
// voxel groups of one batch touch disjoint point sets and can be processed independently
__device__ __forceinline__ bool regular_coord(float x, float nf, float ds) {
    float lo = __fmul_rn(nf, ds), hi = __fadd_rn(lo, ds);
    float lo_next = __fmul_rn(nf + 1.0f, ds);
    float hi_prev = __fadd_rn(__fmul_rn(nf - 1.0f, ds), ds);
    return x >= lo && x < hi && x < lo_next && x >= hi_prev;
}

// Core of the per-voxel decision: box-search the tree (existing points of the voxel), then replay the
// reference's per-point decisions for the voxel's new points `members[0..cnt)` (ascending input order).
__device__ VoxOut vox_decide_core(const Ctx& c, const float4* __restrict__ pts, const int* members, int nmem, float ds,
                                  float* box6, bool& reg_out, unsigned int* nvisit = nullptr) {
    float4 p0 = pts[members[0]];
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
            const Rec64 rec = load_rec64_cg(c.srec + cur);
            const float4 a = rec.a, q1 = rec.b, q2 = rec.c, q3 = rec.e;
            if (nvisit) (*nvisit)++;
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
    int c_exist = cnt;  // points of the box currently in the tree (as the reference would see it)
    bool have_inc = cnt >= 1;
    float inc_d = best_d, ix = bx, iy = by, iz = bz;
    int inc_kind = 2, inc_ref = best_pid;
    int acts = 0;
    for (int q = 0; q < nmem; q++) {
        int j = members[q];
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
    box6[0] = lo[0]; box6[1] = lo[1]; box6[2] = lo[2]; box6[3] = hi[0]; box6[4] = hi[1]; box6[5] = hi[2];
    reg_out = reg;
    return o;
}

// One thread per voxel group (groups = segments of the sorted index list).
__global__ void voxel_decide_kernel(Ctx c, const float4* __restrict__ pts, const int* __restrict__ idx,
                                    const int* __restrict__ seg_begin, Counters* __restrict__ k, float ds,
                                    VoxOut* __restrict__ out, float* __restrict__ boxes, bool count) {
    const int G = k->G;
    GRID_STRIDE(g, G) {
        int b = seg_begin[g], e = seg_begin[g + 1];
        bool reg;
        unsigned int nvis = 0;
        out[g] = vox_decide_core(c, pts, idx + b, e - b, ds, boxes + 6 * (size_t)g, reg, count ? &nvis : nullptr);
        if (count) atomicAdd(&k->vox_visits, (unsigned long long)nvis * (unsigned long long)(e - b));
        if (!reg) atomicExch(&k->irregular, 1);
    }
}

// ---- sort-free variant for scan-sized batches (hash-linked voxel groups) -----------------------------
// First kernel of a scan-sized Add_Points: clears the per-operation counters, the voxel hash table and the survivor
// flags in one launch (three memset nodes before).
__global__ void vox_prep_kernel(uint32_t* __restrict__ counters_words, int ncnt, uint4* __restrict__ ht16, int nht16,
                                int* __restrict__ surv_flag, int n) {
    pdl_wait();  // (the previous operation's last kernels may still be draining)
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = i0; i < nht16; i += stride) ht16[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    for (int i = i0; i < n; i += stride) surv_flag[i] = 0;
    for (int i = i0; i < ncnt; i += stride) counters_words[i] = 0u;
}
__global__ void vox_link_kernel(const float4* __restrict__ pts, int n, float ds, VoxPack vp, HashTab ht,
                                int* __restrict__ next, int* __restrict__ glist, Counters* __restrict__ k) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pts[i];
    float f[3] = {floorf(__fdiv_rn(p.x, ds)), floorf(__fdiv_rn(p.y, ds)), floorf(__fdiv_rn(p.z, ds))};
    unsigned long long key = 0;
    bool bad = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float rel = f[a] - vp.org[a];
        float lim = (float)(1u << vp.bits[a]);
        if (!(rel >= 0.f && rel < lim && fabsf(f[a]) < 8388608.f)) { bad = true; rel = 0.f; }
        key = (key << vp.bits[a]) | (unsigned long long)rel;
    }
    if (bad) { k->oor = 1; return; }
    bool created;
    uint32_t slot = ht_find_or_insert(ht, key, created);
    next[i] = atomicExch(&ht.head[slot], i);
    if (created) glist[atomicAdd(&k->G, 1)] = (int)slot;
}

// one thread per voxel group; members come from the group's list and are put in ascending input order
__global__ void vox_decide_linked_kernel(Ctx c, const float4* __restrict__ pts, HashTab ht, const int* __restrict__ next,
                                         const int* __restrict__ glist, Counters* __restrict__ k, float ds,
                                         VoxOut* __restrict__ out, float* __restrict__ del_boxes,
                                         int* __restrict__ surv_flag, bool count) {
    pdl_wait();
    const int G = k->G;
    GRID_STRIDE(g, G) {
        int slot = glist[g];
        int m[32];
        int cnt = 0;
        bool overflow = false;
        for (int j = ht.head[slot]; j >= 0; j = next[j]) {
            if (cnt == 32) { overflow = true; break; }
            int x = cnt++;
            while (x > 0 && m[x - 1] > j) { m[x] = m[x - 1]; x--; }
            m[x] = j;
        }
        if (overflow) { k->oor = 1; continue; }  // very crowded voxel: the host falls back to the sorted path
        bool reg;
        float box[6];
        unsigned int nvis = 0;
        VoxOut o = vox_decide_core(c, pts, m, cnt, ds, box, reg, count ? &nvis : nullptr);
        if (count) atomicAdd(&k->vox_visits, (unsigned long long)nvis * (unsigned long long)cnt);
        out[g] = o;
        if (!reg) atomicExch(&k->irregular, 1);
        if (o.acts) atomicAdd(&k->acts, o.acts);
        if (o.del_box) {
            int q = atomicAdd(&k->ndel, 1);
            for (int a = 0; a < 6; a++) del_boxes[6 * (size_t)q + a] = box[a];
        }
        // survivors are ordered by input index: a new point by its own index, a re-inserted existing point by
        // the first input index of its voxel (ids stay deterministic although group numbers are not)
        if (o.kind) surv_flag[o.kind == 1 ? o.ref : m[0]] = (int)g + 1;
    }
}

// Compact the survivors in input order (4096 flags per block, chained scan across the blocks).
__global__ void __launch_bounds__(1024)
surv_scan_kernel(const int* __restrict__ surv_flag, int n, const VoxOut* __restrict__ vo, const float4* __restrict__ pts,
                 const float4* __restrict__ pid_xyz, float4* __restrict__ surv, int32_t* __restrict__ src, int src_base,
                 Counters* __restrict__ k, unsigned long long* __restrict__ chain, PublishTicket pub) {
    pdl_wait();
    constexpr int IT = 4;
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int tid = threadIdx.x;
    const int base = blockIdx.x * 1024 * IT;
    int f[IT], v[IT], o[IT], tot;
#pragma unroll
    for (int j = 0; j < IT; j++) {
        int i = base + tid * IT + j;
        f[j] = i < n ? surv_flag[i] : 0;
        v[j] = f[j] ? 1 : 0;
    }
    Scan(tmp).ExclusiveSum(v, o, tot);
    const int c0 = (int)chain_base(chain, (unsigned long long)tot);
#pragma unroll
    for (int j = 0; j < IT; j++) {
        if (f[j]) {
            VoxOut x = vo[f[j] - 1];
            float4 p = x.kind == 1 ? pts[x.ref] : pid_xyz[x.ref];
            surv[c0 + o[j]] = make_float4(p.x, p.y, p.z, 0.f);
            src[c0 + o[j]] = x.kind == 1 ? src_base + x.ref : ~x.ref;
        }
    }
    if (blockIdx.x == gridDim.x - 1) {
        // the last block knows the survivor total; every other counter the host wants was final before this launch
        if (tid == 0) k->nins = c0 + tot;
        if (pub.dst) publish_words(k, (int)(offsetof(Counters, chain_surv) / 4), pub.dst, pub.flag, pub.seq);
    }
}

// Single block: positions of the delete boxes / survivors among the voxel groups and the act total.
// The two position counters are packed into one 64-bit scan; 8 groups per thread per round.
__global__ void __launch_bounds__(1024)
voxel_plan_kernel(const VoxOut* __restrict__ vo, Counters* __restrict__ k, int* __restrict__ del_pos,
                  int* __restrict__ ins_pos) {
    constexpr int IT = 8;
    typedef cub::BlockScan<unsigned long long, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ unsigned long long carry;
    __shared__ int acts_total;
    const int G = k->G;
    const int tid = threadIdx.x;
    if (tid == 0) { carry = 0; acts_total = 0; }
    __syncthreads();
    int my_acts = 0;
    for (int base = 0; base < G; base += 1024 * IT) {
        unsigned long long v[IT], o[IT], tot;
#pragma unroll
        for (int j = 0; j < IT; j++) {
            int g = base + tid * IT + j;
            v[j] = 0;
            if (g < G) { VoxOut x = vo[g]; v[j] = ((unsigned long long)(x.del_box ? 1 : 0) << 32) | (unsigned long long)(x.kind ? 1 : 0); my_acts += x.acts; }
        }
        Scan(tmp).ExclusiveSum(v, o, tot);
        unsigned long long c0 = carry;
#pragma unroll
        for (int j = 0; j < IT; j++) {
            int g = base + tid * IT + j;
            if (g < G) { unsigned long long w = c0 + o[j]; del_pos[g] = (int)(w >> 32); ins_pos[g] = (int)(w & 0xffffffffu); }
        }
        __syncthreads();
        if (tid == 0) carry = c0 + tot;
        __syncthreads();
    }
    atomicAdd(&acts_total, my_acts);
    __syncthreads();
    if (tid == 0) { k->ndel = (int)(carry >> 32); k->nins = (int)(carry & 0xffffffffu); k->acts = acts_total; }
}

// compact the voxel decisions: delete boxes, survivors (coordinates + payload source)
__global__ void voxel_apply_kernel(const VoxOut* __restrict__ vo, const Counters* __restrict__ k,
                                   const float* __restrict__ boxes, const float4* __restrict__ pts,
                                   const float4* __restrict__ pid_xyz, const int* __restrict__ del_pos,
                                   const int* __restrict__ ins_pos, float* __restrict__ del_boxes,
                                   float4* __restrict__ surv, int32_t* __restrict__ src, int src_base) {
    const int G = k->G;
    GRID_STRIDE(g, G) {
        VoxOut o = vo[g];
        if (o.del_box) {
            int q = del_pos[g];
            for (int a = 0; a < 6; a++) del_boxes[6 * (size_t)q + a] = boxes[6 * (size_t)g + a];
        }
        if (o.kind) {
            int q = ins_pos[g];
            float4 v = o.kind == 1 ? pts[o.ref] : pid_xyz[o.ref];
            surv[q] = make_float4(v.x, v.y, v.z, 0.f);
            src[q] = o.kind == 1 ? src_base + o.ref : ~o.ref;
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <class T>
int d2h(ikd_tree* t, T* host, const void* dev, size_t count) {
    return fetch_small(t, host, dev, sizeof(T) * count);
}

Ctx ctx_of(ikd_tree* t) { return Ctx{t->srec, t->urec, t->hdr_dev, t->wrec}; }

// host wall-clock trace of one public call (env IKD_PHASES=1): elapsed ms since the previous mark
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point last;
    explicit HostTrace(bool o) : on(o), last(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(now - last).count();
        if (ms > 0.2) fprintf(stderr, "[ikd host] %-22s %.3f ms\n", what, ms);
        last = now;
    }
};
Counters* counters(ikd_tree* t) { return t->u[U_CNT].as<Counters>(); }

int cub_inclusive_sum_int(ikd_tree* t, const int* in, int* out, int n) {
    size_t tmp = 0;
    IKD_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, n, t->stream));
    IKD_TRY(t->b_cubtmp.ensure(tmp, t->stream));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA(cub::DeviceScan::InclusiveSum(t->b_cubtmp.p, tb, in, out, n, t->stream));
    return IKD_OK;
}
template <typename KeyT>
int cub_sort_pairs(ikd_tree* t, const KeyT* kin, KeyT* kout, const int* vin, int* vout, int n, int end_bit) {
    size_t tmp = 0;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<KeyT, int>(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, n, 0, end_bit, t->stream)));
    IKD_TRY(t->b_cubtmp.ensure(tmp, t->stream));
    size_t tb = t->b_cubtmp.bytes;
    IKD_CUDA((cub::DeviceRadixSort::SortPairs<KeyT, int>(t->b_cubtmp.p, tb, kin, kout, vin, vout, n, 0, end_bit, t->stream)));
    return IKD_OK;
}

int ensure_counters(ikd_tree* t) {
    if (!t->u[U_CNT].p) {
        IKD_TRY(t->u[U_CNT].ensure(sizeof(Counters), t->stream));
        IKD_CUDA(cudaMemsetAsync(t->u[U_CNT].p, 0, sizeof(Counters), t->stream));
    }
    return IKD_OK;
}
int ensure_removed_cap(ikd_tree* t) {
    int64_t need = std::max<int64_t>(t->next_pid, 1024);
    if (need > t->removed_cap) {
        IKD_TRY(t->b_removed.ensure((size_t)need * 4, t->stream, true));
        t->removed_cap = (int64_t)(t->b_removed.bytes / 4);
    }
    return IKD_OK;
}
// the scalar counters (first 64 bytes) and the removed-point count; the chain slots stay on the device
int read_counters(ikd_tree* t, Counters* out) {
    Counters* k = t->u[U_CNT].as<Counters>();
    return fetch_small(t, out, k, offsetof(Counters, chain_surv), &out->nremoved, &k->nremoved, sizeof(unsigned int));
}

// Reset the per-operation counters (everything but the removed-point log count) and size the changed list.
// keep_results: leave delcount and err alone (the commit of a side-stream rebuild / a whole-tree rebuild can run inside
// settle(), i.e. between an operation's kernels and the read of its results).
int begin_changes(ikd_tree* t, int64_t changed_cap, bool keep_results = false, bool reset_by_caller = false) {
    IKD_TRY(ensure_counters(t));
    // (+ room for the roots of a side-stream rebuild that the operation commits before its own kernels, commit_async)
    IKD_TRY(t->u[U_CHANGED].ensure((size_t)(std::max<int64_t>(changed_cap, 16) + t->async.R + 16) * 4, t->stream));
    char* base = (char*)t->u[U_CNT].p;
    if (reset_by_caller) return IKD_OK;  // the caller's first kernel clears the counters (vox_prep_kernel)
    if (!keep_results) {
        IKD_CUDA(cudaMemsetAsync(base, 0, offsetof(Counters, nremoved), t->stream));
    } else {
        IKD_CUDA(cudaMemsetAsync(base, 0, offsetof(Counters, delcount), t->stream));
        IKD_CUDA(cudaMemsetAsync(base + offsetof(Counters, irregular), 0, offsetof(Counters, nremoved) - offsetof(Counters, irregular), t->stream));
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
    IKD_LAUNCH alive_kernel<<<sgrid(top), TPB, 0, s>>>(ctx_of(t), top, t->u[U_ALIVE].as<uint8_t>(), log_removed,
                                                      t->b_removed.as<int32_t>(), counters(t), (unsigned)t->removed_cap);
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

}  // namespace

void rebuild_time_begin(ikd_tree* t, int kind, int64_t points, cudaStream_t s) {
    if (!t->time_rebuilds) return;
    ikd_tree::RebuildTiming r;
    r.kind = kind; r.points = points; r.a = nullptr; r.b = nullptr;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { cudaGetLastError(); return; }
    cudaEventRecord(r.a, s);
    t->rebuild_events.push_back(r);
}
void rebuild_time_end(ikd_tree* t, cudaStream_t s) {
    if (!t->time_rebuilds || t->rebuild_events.empty()) return;
    cudaEventRecord(t->rebuild_events.back().b, s);
}

// device-side Add_Points work counters -> t->stats (one small read; called by ikd_get_stats)
int read_update_stats(ikd_tree* t) {
    if (!t->u[U_CNT].p) return IKD_OK;
    unsigned long long v[2] = {0, 0};
    IKD_TRY(fetch_small(t, v, &t->u[U_CNT].as<Counters>()->vox_visits, sizeof(v)));
    t->stats.add_vox_visits = (int64_t)v[0];
    t->stats.add_descend_levels = (int64_t)v[1];
    return IKD_OK;
}

namespace {

// Enqueue: refit the ancestors of the changed list, find the rebuild roots, plan their rebuild. Results land
// in the header's plan[] (fetched by the caller with sync_header).
int enqueue_refit_and_plan(ikd_tree* t, int64_t changed_cap, PublishTicket* ticket) {
    cudaStream_t s = t->stream;
    Ctx c = ctx_of(t);
    int64_t dcap = std::min<int64_t>(changed_cap * (int64_t)(t->hdr.max_depth + 36), (int64_t)t->cap_slots);
    dcap = std::max<int64_t>(dcap, 64);
    IKD_TRY(t->u[U_DIRTY].ensure((size_t)dcap * 4, s));
    IKD_TRY(t->u[U_ROOTS].ensure((size_t)dcap * 4, s));
    IKD_TRY(t->u[U_RINFO].ensure(((size_t)dcap + 1) * 4 * 3, s));
    Counters* k = counters(t);
    // (ndirty, nroots, nroots_big are zero here: begin_changes precedes every settle and the pass runs once)
    int32_t* changed = t->u[U_CHANGED].as<int32_t>();
    int32_t* dirty = t->u[U_DIRTY].as<int32_t>();
    int* seg_begin = t->u[U_RINFO].as<int>();
    int* soff = seg_begin + (dcap + 1);
    int* boff = soff + (dcap + 1);
    IKD_PHASE(t, "mark");
    IKD_LAUNCH_PDL((mark_kernel), sgrid(changed_cap), TPB, 0, s, c, changed, k, dirty);
    IKD_PHASE(t, "refit");
    IKD_LAUNCH_PDL((refit_kernel), sgrid(dcap), TPB, 0, s, c, dirty, k, t->delete_param, t->balance_param);
    IKD_PHASE(t, "collect+plan");
    const bool can_defer = t->async_min > 0 && !t->async.pending;
    if (can_defer) {
        IKD_TRY(t->async.roots.ensure((size_t)dcap * 4, s));
        IKD_TRY(t->async.plan.ensure(((size_t)dcap + 1) * 4 * 3, s));
        t->async.stride = dcap + 1;
    }
    *ticket = publish_ticket(t);
    PlanArrays pa = {seg_begin, soff, boff, t->hdr_dev->plan};
    PlanArrays pb = {nullptr, nullptr, nullptr, nullptr};
    if (can_defer) {
        int* ap = t->async.plan.as<int>();
        pb = PlanArrays{ap, ap + t->async.stride, ap + 2 * t->async.stride, t->hdr_dev->plan2};
    }
    IKD_LAUNCH_PDL((collect_plan_kernel), sgrid(dcap), TPB, 0, s, c, dirty, k, t->u[U_ROOTS].as<int32_t>(),
                   t->async.roots.as<int32_t>(), can_defer ? t->async_min : 0, pa, pb, *ticket);
    t->rinfo_stride = dcap + 1;
    return IKD_OK;
}

// Rebuild the R subtrees rooted at U_ROOTS as planned (plan[] already on the host). Leaves the next changed
// list (rebuilt roots, or parents of vanished ones) in U_CHANGED.
int rebuild_forest(ikd_tree* t, int R, int M, int S, int B, int max_seg, bool adopt_now = true) {
    cudaStream_t s = t->stream;
    if (t->async.pending) IKD_CUDA(cudaStreamWaitEvent(s, t->side_done, 0));  // the builder's scratch is shared
    Counters* k = counters(t);
    int32_t* roots = t->u[U_ROOTS].as<int32_t>();
    int* seg_begin = t->u[U_RINFO].as<int>();
    int* soff = seg_begin + t->rinfo_stride;
    int* boff = soff + t->rinfo_stride;
    // pool room for the new blocks (grows by reallocation; slot numbers stay valid)
    unsigned int pool_base = t->hdr.pool_top;
    if ((size_t)pool_base + (size_t)B + 2 > t->cap_slots) IKD_TRY(ensure_pool(t, (size_t)pool_base + (size_t)B + 1024, true));
    Ctx c = ctx_of(t);
    IKD_TRY(t->u[U_STACK].ensure((size_t)std::max(S, 1) * sizeof(uint2), s));
    IKD_TRY(t->u[U_P4].ensure((size_t)std::max(M, 1) * sizeof(float4), s));
    IKD_TRY(t->u[U_EROOT].ensure((size_t)std::max(M, 1) * 4, s));
    IKD_TRY(t->u[U_FOREST].ensure((size_t)R * 4 * 5 + 64, s));
    IKD_TRY(ensure_removed_cap(t));
    IKD_PHASE(t, "flatten");
    rebuild_time_begin(t, 0, M, s);
    IKD_LAUNCH_PDL((flatten_kernel<FL_TPB>), std::min(R, MAX_GRID * 2), FL_TPB, 0, s,
                   c, roots, R, seg_begin, soff, t->u[U_STACK].as<uint2>(), t->u[U_P4].as<float4>(), t->u[U_EROOT].as<int>(),
                   t->b_removed.as<int32_t>(), k, (unsigned)t->removed_cap, true, true, nullptr, nullptr, nullptr);
    IKD_PHASE(t, "rebuild_build");
    int* root_slot = t->u[U_FOREST].as<int>();
    int* block_base = root_slot + R;
    int* root_parent = block_base + R;
    int* root_depth = root_parent + R;
    int* single_axis = root_depth + R;
    IKD_LAUNCH_PDL((forest_setup_kernel), nblk(R), TPB, 0, s, c, roots, R, seg_begin, boff, pool_base, root_slot, block_base,
                                                          root_parent, root_depth, single_axis, pool_base + (unsigned)B);
    if (B > 0) {
        IKD_CUDA(cudaMemsetAsync(t->urec + pool_base, 0, (size_t)B * sizeof(UpdateRec), s));  // defined flags below pool_top
        t->hdr.pool_top = pool_base + (unsigned)B;
    }
    IKD_PHASE(t, "rebuild_forest_build");
    if (M > 0) {
        ForestDev f;
        f.R = R; f.seg_begin = seg_begin; f.root_slot = root_slot; f.block_base = block_base; f.root_parent = root_parent;
        f.root_depth = root_depth; f.single_axis = single_axis; f.elem_root = R > 1 ? t->u[U_EROOT].as<int>() : nullptr;
        IKD_TRY(forest_build(t, t->u[U_P4].as<float4>(), M, f, max_seg, s));
    }
    IKD_PHASE(t, "rebuild_adopt");
    t->stats.rebuilds_partial += R;
    t->stats.rebuilt_points += M;
    if (t->phase_on) fprintf(stderr, "[ikd rebuild] R=%d M=%d S=%d B=%d max_seg=%d\n", R, M, S, B, max_seg);
    // the ancestors of the rebuilt roots already carry their post-rebuild criteria and boxes; sizes follow here
    // (when large subtrees of the same pass go to the side stream, the adoption runs there after their flatten: the
    // flatten sizes its sub-root regions with the physical sizes of nodes inside those subtrees)
    if (adopt_now)
        IKD_LAUNCH_PDL((adopt_effective_kernel), sgrid(std::max(t->hdr.plan[6], 1)), TPB, 0, s, c, t->u[U_DIRTY].as<int32_t>(),
                       (unsigned int)t->hdr.plan[6]);
    rebuild_time_end(t, s);
    IKD_PHASE(t, "after_rebuild");
    IKD_CUDA(cudaGetLastError());
    return IKD_OK;
}

int settle(ikd_tree* t, int64_t changed_cap);
}  // namespace
int commit_async(ikd_tree* t);
namespace {
// pool hygiene: when most of the pool is garbage left behind by rebuilds, compact
inline bool pool_hygiene_due(const ikd_tree* t) {
    return t->hdr.root_exists && (size_t)t->hdr.pool_top > t->cap_slots / 2 &&
           (size_t)t->hdr.pool_top > 4 * (size_t)t->hdr.size + (1u << 16);
}

// Start the rebuild of the R large subtrees listed in async.roots on the side stream (plan arrays in async.plan).
// The old subtrees stay in place and searchable; finish_async() swaps the results in before the next mutation.
int enqueue_async_rebuild(ikd_tree* t, int R, int M, int S, int B, int max_seg, bool adopt_after_flatten = false) {
    cudaStream_t ms = t->stream, ss = t->side;
    Ctx c = ctx_of(t);
    int* seg_begin = t->async.plan.as<int>();
    int* soff = seg_begin + t->async.stride;
    int* boff = soff + t->async.stride;
    unsigned int pool_base = t->hdr.pool_top;
    if ((size_t)pool_base + (size_t)B + 2 > t->cap_slots) {
        IKD_TRY(ensure_pool(t, (size_t)pool_base + (size_t)B + 1024, true));
        c = ctx_of(t);
    }
    IKD_TRY(t->async.stack.ensure((size_t)std::max(S, 1) * sizeof(uint2), ms));
    IKD_TRY(t->async.visited.ensure((size_t)std::max(S, 1) * sizeof(int32_t), ms));
    IKD_TRY(t->async.p4.ensure((size_t)std::max(M, 1) * sizeof(float4), ms));
    IKD_TRY(t->async.eroot.ensure((size_t)std::max(M, 1) * 4, ms));
    IKD_TRY(t->async.forest.ensure((size_t)R * 4 * 5 + 64, ms));
    IKD_TRY(t->async.split.ensure((size_t)R * SPLIT_MAX * 4 * 5, ms));
    HostTrace tr(t->phase_on);
    t->hdr.pool_top = pool_base + (unsigned)B;
    IKD_LAUNCH set_pool_top_kernel<<<1, 1, 0, ms>>>(t->hdr_dev, t->hdr.pool_top);
    IKD_CUDA(cudaEventRecord(t->main_ev, ms));
    IKD_CUDA(cudaStreamWaitEvent(ss, t->main_ev, 0));  // everything enqueued so far (refit, small rebuilds) comes first
    rebuild_time_begin(t, 1, M, ss);
    IKD_CUDA(cudaMemsetAsync(t->urec + pool_base, 0, (size_t)B * sizeof(UpdateRec), ss));
    {
        const int NS = R * SPLIT_MAX;
        int32_t* sub_root = t->async.split.as<int32_t>();
        int* sub_seg = sub_root + NS;
        int* sub_stack = sub_seg + NS;
        int* sub_limit = sub_stack + NS;
        int* sub_of = sub_limit + NS;
        IKD_LAUNCH split_roots_kernel<<<R, SPLIT_MAX, 0, ss>>>(c, t->async.roots.as<int32_t>(), R, seg_begin, soff,
                                                              t->async.p4.as<float4>(), t->async.eroot.as<int>(),
                                                              t->async.visited.as<int32_t>(), sub_root, sub_seg, sub_stack,
                                                              sub_limit, sub_of);
        IKD_LAUNCH_PDL((flatten_kernel<FL_TPB>), std::min(NS, MAX_GRID * 2), FL_TPB, 0, ss,
                       c, sub_root, NS, sub_seg, sub_stack, t->async.stack.as<uint2>(), t->async.p4.as<float4>(), t->async.eroot.as<int>(),
                       nullptr, counters(t), 0u, true, false, t->async.visited.as<int32_t>(), sub_limit, sub_of);
        if (adopt_after_flatten) {  // U_DIRTY and the counters stay untouched until finish_async (the next mutation) waits for this stream
            IKD_LAUNCH_PDL((adopt_effective_kernel), sgrid(std::max(t->hdr.plan[6], 1)), TPB, 0, ss, c, t->u[U_DIRTY].as<int32_t>(),
                           (unsigned int)t->hdr.plan[6]);
            // the adoption stores size and invalid of live ancestors separately; a reader of both (the two-pass range search
            // of round 1 was one; the single-pass search reads neither) must not run next to it: run_search still waits
            // for this event, which costs nothing when no adoption is in flight
            IKD_CUDA(cudaEventRecord(t->adopt_ev, ss));
            t->adopt_in_flight = true;
        }
    }
    int* root_slot = t->async.forest.as<int>();
    int* block_base = root_slot + R;
    int* root_parent = block_base + R;
    int* root_depth = root_parent + R;
    int* single_axis = root_depth + R;
    IKD_LAUNCH forest_setup_async_kernel<<<nblk(R), TPB, 0, ss>>>(c, t->async.roots.as<int32_t>(), R, boff, pool_base, root_slot,
                                                                 block_base, root_parent, root_depth, single_axis);
    ForestDev f;
    f.R = R; f.seg_begin = seg_begin; f.root_slot = root_slot; f.block_base = block_base; f.root_parent = root_parent;
    f.root_depth = root_depth; f.single_axis = single_axis; f.elem_root = R > 1 ? t->async.eroot.as<int>() : nullptr;
    tr.mark("async: flatten enqueue");
    IKD_TRY(forest_build(t, t->async.p4.as<float4>(), M, f, max_seg, ss));
    tr.mark("async: forest_build");
    rebuild_time_end(t, ss);
    IKD_CUDA(cudaEventRecord(t->side_done, ss));
    t->async.pending = true;
    t->async.R = R;
    t->async.S = S;
    t->stats.rebuilds_partial += R;
    t->stats.rebuilds_async += R;
    t->stats.rebuilt_points += M;
    if (t->phase_on) fprintf(stderr, "[ikd async rebuild] R=%d M=%d B=%d max_seg=%d\n", R, M, B, max_seg);
    return IKD_OK;
}

// After a batch touched the slots in U_CHANGED: refit, rebuild violating subtrees, refit their ancestors
// (:704-707). Ends with the host header mirror up to date.
int settle(ikd_tree* t, int64_t changed_cap) {
    for (int iter = 0; iter < 64; iter++) {
        const bool planned_async = t->async_min > 0 && !t->async.pending;
        PublishTicket ticket;
        IKD_TRY(enqueue_refit_and_plan(t, changed_cap, &ticket));
        IKD_PHASE(t, "settle_d2h");
        IKD_TRY(publish_wait(t, ticket, &t->hdr, sizeof(TreeHeader)));  // the planner kernel published the header itself
        if (t->hdr.flag1) { set_error("internal: subtree size bookkeeping inconsistent (flatten overflow)"); return IKD_ERR_INTERNAL; }
        const int* p = t->hdr.plan;
        const int* p2 = t->hdr.plan2;
        IKD_PHASE(t, "host_gap");
        int R = p[0];
        const int Rb = planned_async ? p2[0] : 0;
        if (Rb > 0) t->hdr.max_depth = std::max(t->hdr.max_depth, p2[7]);
        // a whole-tree rebuild that is due anyway (depth bound, pool hygiene, see below) makes the partial ones pointless
        const bool whole_due = std::max(t->hdr.max_depth, std::max(p[7], Rb > 0 ? p2[7] : 0)) >= 60 || pool_hygiene_due(t);
        if (whole_due || p[5]) {  // p[5]: the criteria fail at the tree root: rebuild everything (also compacts the node pool)
            IKD_TRY(rebuild_all(t));
            return IKD_OK;
        }
        if (R == 0) {
            HostTrace tr(t->phase_on);
            if (Rb > 0) IKD_TRY(enqueue_async_rebuild(t, Rb, p2[1], p2[2], p2[3], p2[4]));
            tr.mark("enqueue_async_rebuild");
            break;
        }
        HostTrace tr(t->phase_on);
        IKD_TRY(rebuild_forest(t, R, p[1], p[2], p[3], p[4], /*adopt_now=*/Rb == 0));
        tr.mark("rebuild_forest");
        t->hdr.max_depth = std::max(t->hdr.max_depth, p[7]);  // what forest_depth_kernel writes on the device
        if (Rb > 0) IKD_TRY(enqueue_async_rebuild(t, Rb, p2[1], p2[2], p2[3], p2[4], /*adopt_after_flatten=*/true));
        tr.mark("enqueue_async_rebuild");
        // One pass is enough: Criterion_Check was evaluated on effective sizes, so no ancestor of a rebuilt
        // subtree can start to violate because of the rebuild (tested on every node in tests/).
        break;
    }
    if (t->hdr.max_depth >= 60 || pool_hygiene_due(t)) {  // keep traversal stacks bounded / compact a pool that is mostly garbage
        // (reached when the rebuilds of this very pass pushed the depth bound or the pool past the limit)
        IKD_TRY(rebuild_all(t));
    }
    return IKD_OK;
}

// Enqueue a lazy box delete over device boxes; touched slots are appended to U_CHANGED. No sync.
int enqueue_box_delete(ikd_tree* t, const float* boxes_dev, int64_t nb, bool downsample, cudaStream_t stream = nullptr) {
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    Counters* k = counters(t);
    return box_delete_launch(t, boxes_dev, nb, downsample, t->u[U_CHANGED].as<int32_t>(), &k->nchanged, &k->delcount, &k->err,
                             stream);
}

// Bulk insert of n device points (float4 xyz) that all become nodes, ids next_pid + i; the new subtree
// roots are appended to U_CHANGED. One host round trip (group count / block sizes).
// join_before_group: work on another stream (the box delete of the same batch) that must be finished before the
// child-pair links are installed (both modify SearchRec.meta), or nullptr.
int enqueue_insert(ikd_tree* t, const float4* pts, int n, bool* built_whole_tree, cudaEvent_t join_before_group = nullptr) {
    cudaStream_t s = t->stream;
    *built_whole_tree = false;
    if (n <= 0) return IKD_OK;
    IKD_TRY(ensure_pid_cap(t, (int64_t)t->next_pid + n));
    int first_pid = t->next_pid;
    IKD_TRY(t->u[U_P4].ensure((size_t)n * sizeof(float4), s));
    if (!t->hdr.root_exists) {
        // empty tree: the batch becomes the tree (the reference would dereference null here, :447/:472)
        IKD_LAUNCH gather_sorted_kernel<<<nblk(n), TPB, 0, s>>>(pts, nullptr, n, first_pid, t->u[U_P4].as<float4>(),
                                                               t->pid_xyz.as<float4>());
        t->next_pid += n;
        IKD_TRY(full_build(t, t->u[U_P4].as<float4>(), n, s));
        IKD_TRY(sync_header(t));
        *built_whole_tree = true;
        return IKD_OK;
    }
    // room: child pairs (<= 2n) + subtree blocks (<= 4n)
    if ((size_t)t->hdr.pool_top + 6 * (size_t)n + 64 > t->cap_slots)
        IKD_TRY(ensure_pool(t, (size_t)t->hdr.pool_top + 6 * (size_t)n + 4096, true));
    Ctx c = ctx_of(t);
    Counters* k = counters(t);
    IKD_TRY(t->u[U_KEYS].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_KEYS2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_GROUP].ensure((size_t)n * 4 * 2, s));
    IKD_TRY(t->u[U_GINFO].ensure(((size_t)n + 1) * 4 * 3 + 64, s));
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
    int* boff = reinterpret_cast<int*>(gkey + n + 1);
    const bool fused = n <= 65536;  // sort-free grouping for scan-sized batches
    PublishTicket ins_ticket;
    if (fused) {
        uint32_t hsz = 1024;
        while (hsz < 2u * (uint32_t)n) hsz <<= 1;
        IKD_TRY(t->u[U_HT].ensure((size_t)hsz * 20, s));  // keys (8 B), head/count (4 B) | segment start, group number
        IKD_TRY(t->u[U_NEXT].ensure((size_t)n * 4 * 4, s));
        HashTab ht;
        ht.keys = t->u[U_HT].as<unsigned long long>();
        ht.head = reinterpret_cast<int*>(ht.keys + hsz);
        ht.mask = hsz - 1;
        int* slot_begin = ht.head + hsz;
        int* slot_gid = slot_begin + hsz;
        int* arrival = t->u[U_NEXT].as<int>();
        int* glist = arrival + n;
        int* slot_of = glist + n;
        int* members = slot_of + n;
        IKD_PHASE(t, "ins_descend");
        IKD_CUDA(cudaMemsetAsync(t->u[U_HT].p, 0xFF, (size_t)hsz * 12, s));
        IKD_LAUNCH_PDL((descend_link_kernel), nblk(n), TPB, 0, s, c, pts, n, ht, arrival, slot_of, glist, k, t->count_visits);
        IKD_PHASE(t, "ins_group");
        static_assert(65536 / IG_TPB <= CHAIN_INS, "chain slots");
        if (join_before_group) { IKD_CUDA(cudaStreamWaitEvent(s, join_before_group, 0)); join_before_group = nullptr; }
        IKD_LAUNCH_PDL((insert_group_kernel), nblk(n, IG_TPB), IG_TPB, 0, s, c, ht, glist, k, seg_begin, gkey, boff, slot_begin,
                                                                         slot_gid, k->chain_ins);
        IKD_LAUNCH_PDL((insert_scatter_kernel), nblk(n), TPB, 0, s, n, arrival, slot_of, slot_begin, members);
        IKD_LAUNCH_PDL((insert_place_kernel), nblk(n), TPB, 0, s, n, pts, ht, arrival, slot_of, slot_begin, slot_gid, members,
                                                              first_pid, t->u[U_EROOT].as<int>(), t->u[U_P4].as<float4>(),
                                                              t->pid_xyz.as<float4>(), k, t->hdr_dev,
                                                              ins_ticket = publish_ticket(t));
    } else {
        if (join_before_group) { IKD_CUDA(cudaStreamWaitEvent(s, join_before_group, 0)); join_before_group = nullptr; }
        IKD_PHASE(t, "ins_descend");
        IKD_LAUNCH descend_kernel<<<nblk(n), TPB, 0, s>>>(c, pts, n, keys, idx, k, t->count_visits);
        IKD_PHASE(t, "ins_group");
        int key_bits = 1;
        while (key_bits < 31 && (1ull << key_bits) <= 2ull * (unsigned long long)t->hdr.pool_top + 1ull) key_bits++;
        IKD_TRY(cub_sort_pairs<uint32_t>(t, keys, keys_s, idx, idx_s, n, key_bits));
        IKD_LAUNCH head_flag_kernel<uint32_t><<<nblk(n), TPB, 0, s>>>(keys_s, n, head);
        IKD_TRY(cub_inclusive_sum_int(t, head, gid, n));
        IKD_LAUNCH group_bounds_kernel<uint32_t><<<nblk(n), TPB, 0, s>>>(keys_s, gid, n, seg_begin, gkey, t->u[U_EROOT].as<int>(),
                                                                         &k->R_ins);
        IKD_LAUNCH alloc_pairs_kernel<<<sgrid(n), TPB, 0, s>>>(c, gkey, &k->R_ins);
        IKD_LAUNCH insert_plan_kernel<<<1, 1024, 0, s>>>(seg_begin, k, boff);
    }
    IKD_PHASE(t, "ins_d2h");
    // one round trip: group count, block total, largest group, pool top after the pair allocations
    int R, B, max_seg;
    unsigned int pool_base;
    {
        Counters hk;
        if (fused) {
            struct { unsigned char c[offsetof(Counters, chain_surv)]; unsigned int pool_top; } buf;
            IKD_TRY(publish_wait(t, ins_ticket, &buf, sizeof(buf)));
            memcpy(&hk, buf.c, sizeof(buf.c));
            pool_base = buf.pool_top;
        } else {
            IKD_TRY(fetch_small(t, &hk, k, offsetof(Counters, chain_surv), &pool_base, &t->hdr_dev->pool_top, 4));
        }
        R = hk.R_ins; B = hk.B_ins; max_seg = hk.maxseg;
    }
    IKD_PHASE(t, "ins_build");
    if (B > 0) IKD_CUDA(cudaMemsetAsync(t->urec + pool_base, 0, (size_t)B * sizeof(UpdateRec), s));
    t->hdr.pool_top = pool_base + (unsigned)B;
    if (t->phase_on) fprintf(stderr, "[ikd insert] n=%d R=%d B=%d max_seg=%d\n", n, R, B, max_seg);
    int* root_slot = t->u[U_FOREST].as<int>();
    int* block_base = root_slot + R;
    int* root_parent = block_base + R;
    int* root_depth = root_parent + R;
    int* single_axis = root_depth + R;
    IKD_LAUNCH_PDL((insert_forest_kernel), nblk(R), TPB, 0, s, c, gkey, R, boff, pool_base, root_slot, block_base, root_parent,
                                                           root_depth, single_axis, t->u[U_CHANGED].as<int32_t>(), k,
                                                           t->hdr.pool_top);
    if (!fused)
        IKD_LAUNCH gather_sorted_kernel<<<nblk(n), TPB, 0, s>>>(pts, idx_s, n, first_pid, t->u[U_P4].as<float4>(),
                                                               t->pid_xyz.as<float4>());
    t->next_pid += n;
    ForestDev f;
    f.R = R; f.seg_begin = seg_begin; f.root_slot = root_slot; f.block_base = block_base; f.root_parent = root_parent;
    f.root_depth = root_depth; f.single_axis = single_axis; f.elem_root = R > 1 ? t->u[U_EROOT].as<int>() : nullptr;
    IKD_TRY(forest_build(t, t->u[U_P4].as<float4>(), n, f, max_seg, s));
    return IKD_OK;
}

}  // namespace

// ================================================================================================
// public implementations
// ================================================================================================
// Swap the result of the side-stream rebuild (if any) in: wait for it, release the old nodes, move the new root records
// into the old root slots and append those roots to the changed list of the operation that is being assembled (the
// caller has run begin_changes). Enqueue only -- the refit of the ancestors is the calling operation's own settle(), so
// a rebuild that finished in the shadow of the previous scan costs the next one two short kernels.
int commit_async(ikd_tree* t) {
    if (!t->async.pending) return IKD_OK;
    cudaStream_t s = t->stream;
    IKD_CUDA(cudaStreamWaitEvent(s, t->side_done, 0));
    const int R = t->async.R;
    Ctx c = ctx_of(t);
    t->adopt_in_flight = false;  // the wait above covers the adoption kernel too
    IKD_TRY(ensure_removed_cap(t));
    Counters* k = counters(t);
    IKD_PHASE(t, "commit_async");
    IKD_LAUNCH_PDL((release_list_kernel), sgrid(std::max(t->async.S, 1)), TPB, 0, s, c, t->async.visited.as<int32_t>(), t->async.S,
                                                                                 t->b_removed.as<int32_t>(), k,
                                                                                 (unsigned)t->removed_cap);
    IKD_LAUNCH_PDL((commit_async_kernel), nblk(R), TPB, 0, s, c, t->async.roots.as<int32_t>(), R, t->async.forest.as<int>(),
                                                          t->u[U_CHANGED].as<int32_t>(), k);
    t->async.pending = false;
    return IKD_OK;
}

// Same, complete: commit and refit now. For calls that export the tree (flatten, dump, replica export, id compaction,
// the removed-point log) rather than mutate it.
int finish_async(ikd_tree* t) {
    if (!t->async.pending) return IKD_OK;
    const int R = t->async.R;
    IKD_TRY(begin_changes(t, R + 16, /*keep_results=*/true));
    IKD_TRY(commit_async(t));
    const int keep = t->async_min;
    t->async_min = 0;  // nothing new goes to the side stream from here: the caller wants a settled tree
    int st = settle(t, R + 16);
    t->async_min = keep;
    return st;
}

int rebuild_all(ikd_tree* t) {
    cudaStream_t s = t->stream;
    int M = 0;
    IKD_TRY(finish_async(t));
    IKD_TRY(sync_header(t));
    rebuild_time_begin(t, 2, t->hdr.size - t->hdr.invalid, s);
    IKD_TRY(select_alive(t, true, &M));
    IKD_TRY(t->u[U_P4].ensure((size_t)std::max(M, 1) * sizeof(float4), s));
    if (M > 0)
        IKD_LAUNCH gather_pid_kernel<<<nblk(M), TPB, 0, s>>>(t->u[U_SEL].as<int32_t>(), M, t->pid_xyz.as<float4>(),
                                                            t->u[U_P4].as<float4>());
    IKD_TRY(full_build(t, t->u[U_P4].as<float4>(), M, s));
    rebuild_time_end(t, s);
    IKD_TRY(sync_header(t));
    t->stats.rebuilds_full += 1;
    t->stats.rebuilt_points += M;
    return IKD_OK;
}

int flatten_impl(ikd_tree* t, int32_t* out_idx, int64_t cap, int64_t* out_n) {
    int M = 0;
    IKD_TRY(finish_async(t));
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
    IKD_TRY(finish_async(t));  // points dropped by a side-stream rebuild are logged when it is committed
    IKD_TRY(ensure_counters(t));
    Counters hk;
    IKD_TRY(read_counters(t, &hk));
    unsigned int n = (unsigned)std::min<int64_t>(hk.nremoved, t->removed_cap);
    *out_n = n;
    if (out_idx) {
        int64_t m = std::min<int64_t>(n, cap);
        if (m > 0) {
            IKD_CUDA(cudaMemcpyAsync(out_idx, t->b_removed.p, (size_t)m * 4, cudaMemcpyDeviceToHost, t->stream));
            IKD_CUDA(cudaStreamSynchronize(t->stream));
        }
        IKD_CUDA(cudaMemsetAsync(&counters(t)->nremoved, 0, 4, t->stream));  // Points_deleted.clear() (:567)
    }
    return IKD_OK;
}

int delete_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb, int* out_deleted) {
    *out_deleted = 0;
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(t->u[U_BOXES].ensure((size_t)nb * 24, t->stream));
    IKD_CUDA(cudaMemcpyAsync(t->u[U_BOXES].p, boxes_host, (size_t)nb * 24, cudaMemcpyHostToDevice, t->stream));
    return delete_boxes_dev_impl(t, t->u[U_BOXES].as<float>(), nb, out_deleted);
}

int delete_boxes_dev_impl(ikd_tree* t, const float* boxes_dev, int64_t nb, int* out_deleted) {
    *out_deleted = 0;
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    int64_t cap = (int64_t)t->hdr.size + 16;
    IKD_TRY(begin_changes(t, cap));
    IKD_TRY(commit_async(t));
    IKD_TRY(enqueue_box_delete(t, boxes_dev, nb, false));
    IKD_TRY(settle(t, cap));
    Counters hk;
    IKD_TRY(read_counters(t, &hk));
    if (hk.err) { set_error("box delete traversal stack overflow"); return IKD_ERR_INTERNAL; }
    *out_deleted = (int)hk.delcount;
    return IKD_OK;
}

int delete_points_impl(ikd_tree* t, const float* xyz, int64_t n, int64_t stride) {
    cudaStream_t s = t->stream;
    if (n == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(t->u[U_PTS].ensure((size_t)n * sizeof(float4), s));
    IKD_TRY(upload_points_f4(t, xyz, n, stride, t->u[U_PTS].as<float4>(), 0, 1));
    return delete_points_dev_impl(t, t->u[U_PTS].as<float4>(), n);
}

int delete_points_dev_impl(ikd_tree* t, const float4* pts_dev, int64_t n) {
    cudaStream_t s = t->stream;
    if (n == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(begin_changes(t, n));
    IKD_TRY(commit_async(t));
    IKD_LAUNCH delete_points_kernel<<<nblk(n), TPB, 0, s>>>(ctx_of(t), pts_dev, (int)n,
                                                           t->u[U_CHANGED].as<int32_t>(), counters(t));
    IKD_TRY(settle(t, n));
    return IKD_OK;
}

namespace {
// key layout of the packed voxel key: indices relative to the tree's range (the batch normally lies in or near
// the map) with a margin of 256 voxels; returns the total number of key bits
int voxel_pack_layout(ikd_tree* t, float ds, VoxPack* vp) {
    int total_bits = 0;
    for (int a = 0; a < 3; a++) {
        float lo = t->hdr.root_exists ? t->hdr.range[a] : 0.f, hi = t->hdr.root_exists ? t->hdr.range[3 + a] : 0.f;
        double o = floor((double)lo / ds) - 256.0, e = floor((double)hi / ds) + 256.0;
        if (!(fabs(o) < 8.0e6 && fabs(e) < 8.0e6)) { o = -1048576.0; e = 1048575.0; }
        vp->org[a] = (float)o;
        int b = 1;
        while ((double)(1u << b) < e - o + 1.0 && b < 21) b++;
        vp->bits[a] = b;
        total_bits += b;
    }
    return total_bits;
}

// group the batch by voxel; leaves idx (sorted element order), seg_begin and k->G on the device
int group_by_voxel(ikd_tree* t, const float4* pts, int n, int** idx_out, int** seg_begin_out) {
    cudaStream_t s = t->stream;
    Counters* k = counters(t);
    float ds = t->downsample;
    IKD_TRY(t->u[U_K64A].ensure((size_t)n * 8, s));
    IKD_TRY(t->u[U_K64B].ensure((size_t)n * 8, s));
    IKD_TRY(t->u[U_IDX].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_IDX2].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_GROUP].ensure((size_t)n * 4 * 2, s));
    IKD_TRY(t->u[U_GINFO].ensure(((size_t)n + 1) * 4 * 3 + 64, s));
    unsigned long long* ka = t->u[U_K64A].as<unsigned long long>();
    unsigned long long* kb = t->u[U_K64B].as<unsigned long long>();
    int* idx_a = t->u[U_IDX].as<int>();
    int* idx_b = t->u[U_IDX2].as<int>();
    int* head = t->u[U_GROUP].as<int>();
    int* gid = head + n;
    int* seg_begin = t->u[U_GINFO].as<int>();
    VoxPack vp;
    int total_bits = voxel_pack_layout(t, ds, &vp);
    IKD_LAUNCH voxel_key64_kernel<<<nblk(n), TPB, 0, s>>>(pts, n, ds, vp, ka, idx_a, k);
    IKD_TRY(cub_sort_pairs<unsigned long long>(t, ka, kb, idx_a, idx_b, n, total_bits));
    IKD_LAUNCH head_flag_kernel<unsigned long long><<<nblk(n), TPB, 0, s>>>(kb, n, head);
    IKD_TRY(cub_inclusive_sum_int(t, head, gid, n));
    IKD_LAUNCH group_bounds_kernel<unsigned long long><<<nblk(n), TPB, 0, s>>>(kb, gid, n, seg_begin, nullptr, nullptr, &k->G);
    *idx_out = idx_b;
    *seg_begin_out = seg_begin;
    return IKD_OK;
}

// same, for coordinates whose voxel index does not fit the packed key (three stable 32-bit passes)
int group_by_voxel_wide(ikd_tree* t, const float4* pts, int n, int** idx_out, int** seg_begin_out) {
    cudaStream_t s = t->stream;
    Counters* k = counters(t);
    float ds = t->downsample;
    IKD_TRY(t->u[U_VOX].ensure((size_t)n * 4 * 3, s));
    IKD_TRY(t->u[U_KEYS].ensure((size_t)n * 4, s));
    IKD_TRY(t->u[U_KEYS2].ensure((size_t)n * 4, s));
    uint32_t* kx = t->u[U_VOX].as<uint32_t>();
    uint32_t* ky = kx + n;
    uint32_t* kz = ky + n;
    int* idx_a = t->u[U_IDX].as<int>();
    int* idx_b = t->u[U_IDX2].as<int>();
    uint32_t* k1 = t->u[U_KEYS].as<uint32_t>();
    uint32_t* k2 = t->u[U_KEYS2].as<uint32_t>();
    int* head = t->u[U_GROUP].as<int>();
    int* gid = head + n;
    int* seg_begin = t->u[U_GINFO].as<int>();
    IKD_LAUNCH voxel_key_kernel<<<nblk(n), TPB, 0, s>>>(pts, n, ds, kx, ky, kz, idx_a);
    const uint32_t* comp[3] = {kz, ky, kx};
    for (int pass = 0; pass < 3; pass++) {
        IKD_LAUNCH gather_u32_kernel<<<nblk(n), TPB, 0, s>>>(comp[pass], idx_a, n, k1);
        IKD_TRY(cub_sort_pairs<uint32_t>(t, k1, k2, idx_a, idx_b, n, 32));
        std::swap(idx_a, idx_b);
    }
    IKD_LAUNCH voxel_head3_kernel<<<nblk(n), TPB, 0, s>>>(kx, ky, kz, idx_a, n, head);
    IKD_TRY(cub_inclusive_sum_int(t, head, gid, n));
    IKD_LAUNCH voxel_bounds3_kernel<<<nblk(n), TPB, 0, s>>>(head, gid, n, seg_begin, &k->G);
    *idx_out = idx_a;
    *seg_begin_out = seg_begin;
    return IKD_OK;
}

// one contiguous piece of an Add_Points batch with downsampling; sets *irregular (and modifies nothing)
// when the piece must be split
int add_downsample_piece(ikd_tree* t, const float4* pts, int n, int src_base, bool force, int* acts_out,
                         int64_t* nins_out, int32_t* src_host, int* irregular) {
    cudaStream_t s = t->stream;
    *irregular = 0;
    *acts_out = 0;
    *nins_out = 0;
    float ds = t->downsample;
    HostTrace tr(t->phase_on);
    int64_t changed_cap = (int64_t)(t->hdr.root_exists ? t->hdr.size : 0) + n + 16;
    IKD_TRY(begin_changes(t, changed_cap, false, /*reset_by_caller=*/n <= 65536));
    Counters* k = counters(t);
    Ctx c = ctx_of(t);
    IKD_TRY(t->u[U_TMP].ensure((size_t)n * sizeof(VoxOut) + 64, s));
    IKD_TRY(t->u[U_TMP2].ensure((size_t)n * 24 + 64, s));
    IKD_TRY(t->u[U_BOXES].ensure((size_t)n * 24 + 64, s));
    IKD_TRY(t->u[U_SURV].ensure((size_t)n * sizeof(float4), s));
    IKD_TRY(t->u[U_SRC].ensure((size_t)n * 4, s));
    VoxOut* vo = t->u[U_TMP].as<VoxOut>();
    float* vboxes = t->u[U_TMP2].as<float>();
    Counters hk;
    PublishTicket vox_ticket;
    // attempt 0: sort-free hash-linked grouping (scan-sized batches); attempt 1: packed-key radix sort;
    // attempt 2: wide three-pass sort. A later attempt is only needed when an earlier one reports `oor`
    // (voxel index out of the packed range, or a voxel with more than 32 new points).
    for (int attempt = (n <= 65536 ? 0 : 1); attempt < 3; attempt++) {
        if (attempt > (n <= 65536 ? 0 : 1)) IKD_TRY(begin_changes(t, changed_cap));
        IKD_PHASE(t, "vox_group");
        if (attempt == 0) {
            VoxPack vp;
            voxel_pack_layout(t, ds, &vp);
            uint32_t hsz = 1024;
            while (hsz < 2u * (uint32_t)n) hsz <<= 1;
            IKD_TRY(t->u[U_HT].ensure((size_t)hsz * 12, s));
            IKD_TRY(t->u[U_NEXT].ensure((size_t)n * 4 * 3, s));
            HashTab ht;
            ht.keys = t->u[U_HT].as<unsigned long long>();
            ht.head = reinterpret_cast<int*>(ht.keys + hsz);
            ht.mask = hsz - 1;
            int* next = t->u[U_NEXT].as<int>();
            int* glist = next + n;
            int* surv_flag = glist + n;
            static_assert(offsetof(Counters, nremoved) % 4 == 0, "counter reset in words");
            IKD_LAUNCH_PDL((vox_prep_kernel), 148, 1024, 0, s, reinterpret_cast<uint32_t*>(k), (int)(offsetof(Counters, nremoved) / 4),
                           t->u[U_HT].as<uint4>(), (int)((size_t)hsz * 12 / 16), surv_flag, n);
            IKD_LAUNCH_PDL((vox_link_kernel), nblk(n), TPB, 0, s, pts, n, ds, vp, ht, next, glist, k);
            IKD_PHASE(t, "vox_decide");
            IKD_LAUNCH_PDL((vox_decide_linked_kernel), sgrid(n, 128), 128, 0, s, c, pts, ht, next, glist, k, ds, vo,
                                                                            t->u[U_BOXES].as<float>(), surv_flag, t->count_visits);
            IKD_PHASE(t, "vox_plan+apply");
            IKD_LAUNCH_PDL((surv_scan_kernel), nblk(n, 4096), 1024, 0, s, surv_flag, n, vo, pts, t->pid_xyz.as<float4>(),
                                                                      t->u[U_SURV].as<float4>(), t->u[U_SRC].as<int32_t>(),
                                                                      src_base, k, k->chain_surv, vox_ticket = publish_ticket(t));
        } else {
            int* idx = nullptr;
            int* seg_begin = nullptr;
            if (attempt == 1) IKD_TRY(group_by_voxel(t, pts, n, &idx, &seg_begin));
            else IKD_TRY(group_by_voxel_wide(t, pts, n, &idx, &seg_begin));
            int* del_pos = seg_begin + (n + 1);
            int* ins_pos = del_pos + (n + 1);
            IKD_PHASE(t, "vox_decide");
            IKD_LAUNCH voxel_decide_kernel<<<sgrid(n, 128), 128, 0, s>>>(c, pts, idx, seg_begin, k, ds, vo, vboxes, t->count_visits);
            IKD_PHASE(t, "vox_plan+apply");
            IKD_LAUNCH voxel_plan_kernel<<<1, 1024, 0, s>>>(vo, k, del_pos, ins_pos);
            IKD_LAUNCH voxel_apply_kernel<<<sgrid(n), TPB, 0, s>>>(vo, k, vboxes, pts, t->pid_xyz.as<float4>(), del_pos, ins_pos,
                                                                  t->u[U_BOXES].as<float>(), t->u[U_SURV].as<float4>(),
                                                                  t->u[U_SRC].as<int32_t>(), src_base);
        }
        IKD_PHASE(t, "vox_d2h");
        // round trip 1: G, irregular, oor, acts, ndel, nins (published by the compaction kernel itself on the fused path)
        if (attempt == 0) IKD_TRY(publish_wait(t, vox_ticket, &hk, offsetof(Counters, chain_surv)));
        else IKD_TRY(read_counters(t, &hk));
        if (!hk.oor || attempt == 2) break;
    }
    tr.mark("voxel phase");
    if (hk.irregular && !force) { *irregular = 1; return IKD_OK; }
    *acts_out = hk.acts;
    int ndel = hk.ndel, nins = hk.nins;
    if (src_host && nins > 0) IKD_CUDA(cudaMemcpyAsync(src_host, t->u[U_SRC].p, (size_t)nins * 4, cudaMemcpyDeviceToHost, s));
    // The voxel phase above only read the tree; everything below mutates it. A side-stream rebuild handed off by the
    // previous call has had the whole voxel phase (and the caller's searches before it) to finish: swap it in now.
    if (ndel > 0 || nins > 0) IKD_TRY(commit_async(t));
    // apply: downsample-delete the boxes, insert the survivors, then ONE refit / rebuild pass for both
    IKD_PHASE(t, "box_delete");
    bool whole = false;
    if (ndel > 0 && nins > 0 && nins <= 65536 && t->hdr.root_exists) {
        // The downsample delete and the descent of the survivors are two full-depth tree walks that do not depend on
        // each other (the descent only looks at split planes and at which children exist): run them side by side and
        // join before the insert installs child links.
        cudaStream_t ds = t->aux[0][0];
        // (the node pool may not move while the delete is in flight: grow it now if the insert will need room)
        if ((size_t)t->hdr.pool_top + 6 * (size_t)nins + 64 > t->cap_slots)
            IKD_TRY(ensure_pool(t, (size_t)t->hdr.pool_top + 6 * (size_t)nins + 4096, true));
        IKD_CUDA(cudaEventRecord(t->aux_fork[0], s));
        IKD_CUDA(cudaStreamWaitEvent(ds, t->aux_fork[0], 0));
        IKD_TRY(enqueue_box_delete(t, t->u[U_BOXES].as<float>(), ndel, true, ds));
        IKD_CUDA(cudaEventRecord(t->aux_ev[0][0], ds));
        IKD_TRY(enqueue_insert(t, t->u[U_SURV].as<float4>(), nins, &whole, t->aux_ev[0][0]));  // round trip 2
    } else {
        if (ndel > 0) IKD_TRY(enqueue_box_delete(t, t->u[U_BOXES].as<float>(), ndel, true));
        if (nins > 0) IKD_TRY(enqueue_insert(t, t->u[U_SURV].as<float4>(), nins, &whole));  // round trip 2
    }
    tr.mark("insert phase");
    if (!whole && (ndel > 0 || nins > 0)) IKD_TRY(settle(t, changed_cap));                // round trips 3 (+1 per rebuild round)
    tr.mark("settle");
    IKD_PHASE(t, "end");
    if (t->phase_on) phase_flush(t);
    *nins_out = nins;
    return IKD_OK;
}

int add_downsample_range(ikd_tree* t, const float4* pts, int off, int n, int* acts, int64_t* nins, int32_t* src_host) {
    if (n <= 0) return IKD_OK;
    int irregular = 0, a = 0;
    int64_t q = 0;
    // a single point is processed exactly as the reference does whatever its geometry (one voxel group cannot
    // conflict with itself), so pieces of size 1 are forced through
    IKD_TRY(add_downsample_piece(t, pts + off, n, off, n == 1, &a, &q, src_host ? src_host + *nins : nullptr, &irregular));
    if (!irregular) {
        *acts += a;
        *nins += q;
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
    if ((int64_t)t->next_pid + n > (int64_t)0x7ffffff0) {
        set_error("point ids exhausted (%d handed out, %d valid points): call ikd_compact_ids", t->next_pid,
                  t->hdr.root_exists ? t->hdr.size - t->hdr.invalid : 0);
        return IKD_ERR_CAPACITY;
    }
    IKD_TRY(ensure_pid_cap(t, (int64_t)t->next_pid + n));
    if (t->count_visits && downsample_on) t->stats.add_points_in += n;
    if (!downsample_on) {
        IKD_TRY(begin_changes(t, n + 16));
        IKD_TRY(commit_async(t));
        bool whole = false;
        IKD_TRY(enqueue_insert(t, pts_dev, (int)n, &whole));
        if (!whole) IKD_TRY(settle(t, n + 16));
        *out_added = 0;  // the reference only counts inserts of the downsample branch (tmp_counter, :448 vs :472)
        *out_ninserted = n;
        if (out_src) for (int64_t i = 0; i < n; i++) out_src[i] = (int32_t)i;
        return IKD_OK;
    }
    int acts = 0;
    int64_t nins = 0;
    IKD_TRY(add_downsample_range(t, pts_dev, 0, (int)n, &acts, &nins, out_src));
    // No stream wait here: every value handed back (counts, payload sources) was read after the kernels that produce
    // it, and the caller's batch is not read after the first host read. What may still be running are the rebuild
    // kernels; later calls are ordered behind them on the tree's stream (ikd_synchronize waits explicitly).
    *out_added = acts;
    *out_ninserted = nins;
    if (t->count_visits) t->stats.add_points_inserted += nins;
    return IKD_OK;
}

// Renumber the valid points 0..M-1 (increasing old id), rebuild the whole tree on them, shrink the id space
// (ikd_compact_ids). The removed-point log is drained with its OLD ids first: they are void afterwards.
int compact_ids_impl(ikd_tree* t, int32_t* old_of_new, int64_t cap_alive, int64_t* out_alive, int32_t* removed_old,
                     int64_t cap_removed, int64_t* out_removed) {
    cudaStream_t s = t->stream;
    *out_alive = 0;
    *out_removed = 0;
    IKD_TRY(finish_async(t));
    IKD_TRY(sync_header(t));
    IKD_TRY(ensure_counters(t));
    Counters hk;
    IKD_TRY(read_counters(t, &hk));
    const int64_t valid = t->hdr.root_exists ? (int64_t)t->hdr.size - t->hdr.invalid : 0;
    const int64_t invalid = t->hdr.root_exists ? (int64_t)t->hdr.invalid : 0;
    if (cap_alive < valid || cap_removed < (int64_t)hk.nremoved + invalid || (valid > 0 && !old_of_new) ||
        ((int64_t)hk.nremoved + invalid > 0 && !removed_old)) {
        set_error("ikd_compact_ids: buffers too small (need %lld alive, %lld removed)", (long long)valid,
                  (long long)((int64_t)hk.nremoved + invalid));
        return IKD_ERR_CAPACITY;
    }
    int M = 0;
    rebuild_time_begin(t, 2, valid, s);
    IKD_TRY(select_alive(t, true, &M));  // alive ids ascending in U_SEL; lazily deleted points join the removed log
    IKD_TRY(read_counters(t, &hk));
    const int64_t nrem = std::min<int64_t>(hk.nremoved, t->removed_cap);
    if (nrem > 0) IKD_CUDA(cudaMemcpyAsync(removed_old, t->b_removed.p, (size_t)nrem * 4, cudaMemcpyDeviceToHost, s));
    if (M > 0) IKD_CUDA(cudaMemcpyAsync(old_of_new, t->u[U_SEL].p, (size_t)M * 4, cudaMemcpyDeviceToHost, s));
    IKD_CUDA(cudaMemsetAsync(&counters(t)->nremoved, 0, 4, s));
    IKD_TRY(t->u[U_P4].ensure((size_t)std::max(M, 1) * sizeof(float4), s));
    if (M > 0) {
        IKD_LAUNCH gather_newid_kernel<<<nblk(M), TPB, 0, s>>>(t->u[U_SEL].as<int32_t>(), M, t->pid_xyz.as<float4>(),
                                                              t->u[U_P4].as<float4>());
        // the id table in the new numbering (through the gathered copy: source and destination ranges overlap)
        IKD_LAUNCH scatter_xyz_kernel<<<nblk(M), TPB, 0, s>>>(t->u[U_P4].as<float4>(), M, t->pid_xyz.as<float4>());
    }
    t->next_pid = M;
    IKD_TRY(full_build(t, t->u[U_P4].as<float4>(), M, s));
    rebuild_time_end(t, s);
    IKD_TRY(sync_header(t));
    IKD_CUDA(cudaStreamSynchronize(s));  // the two result copies
    t->stats.rebuilds_full += 1;
    t->stats.rebuilt_points += M;
    t->id_epoch++;
    *out_alive = M;
    *out_removed = nrem;
    return IKD_OK;
}

// zero the device-side removed-point log (a new Build starts a new id numbering)
int reset_removed_log(ikd_tree* t) {
    if (t->u[U_CNT].p) IKD_CUDA(cudaMemsetAsync(&counters(t)->nremoved, 0, 4, t->stream));
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

// Add_Point_Boxes (:492-511)
int add_boxes_impl(ikd_tree* t, const float* boxes_host, int64_t nb) {
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    IKD_TRY(t->u[U_BOXES].ensure((size_t)nb * 24, t->stream));
    IKD_CUDA(cudaMemcpyAsync(t->u[U_BOXES].p, boxes_host, (size_t)nb * 24, cudaMemcpyHostToDevice, t->stream));
    return add_boxes_dev_impl(t, t->u[U_BOXES].as<float>(), nb);
}

int add_boxes_dev_impl(ikd_tree* t, const float* boxes_dev, int64_t nb) {
    if (nb == 0 || !t->hdr.root_exists) return IKD_OK;
    int64_t cap = (int64_t)t->hdr.size + 16;
    IKD_TRY(begin_changes(t, cap));
    IKD_TRY(commit_async(t));
    Counters* k = counters(t);
    IKD_TRY(box_add_launch(t, boxes_dev, nb, t->u[U_CHANGED].as<int32_t>(), &k->nchanged, &k->err));
    IKD_TRY(settle(t, cap));
    Counters hk;
    IKD_TRY(read_counters(t, &hk));
    if (hk.err) { set_error("box re-insert traversal stack overflow"); return IKD_ERR_INTERNAL; }
    return IKD_OK;
}

// Pre-order structure dump for parity tests (columns as oracle/ref_harness.cpp ref_dump_tree).
int dump_tree_impl(ikd_tree* t, float* out, int64_t cap, int64_t* out_n) {
    IKD_TRY(finish_async(t));
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

// Load every kernel of this file now (CUDA loads kernels lazily at their first launch, 0.1-0.3 ms each, which
// showed up as milliseconds of extra latency in the first update after Build).
#define IKD_PRELOAD(fn) do { cudaFuncAttributes a_; if (cudaFuncGetAttributes(&a_, fn) != cudaSuccess) cudaGetLastError(); } while (0)
void preload_update_kernels() {
    IKD_PRELOAD(adopt_effective_kernel); IKD_PRELOAD(alive_kernel); IKD_PRELOAD(alloc_pairs_kernel);
    IKD_PRELOAD(collect_plan_kernel); IKD_PRELOAD(commit_async_kernel); IKD_PRELOAD(delete_points_kernel);
    IKD_PRELOAD(descend_kernel); IKD_PRELOAD(descend_link_kernel); IKD_PRELOAD(flatten_kernel<FL_TPB>);
    IKD_PRELOAD(split_roots_kernel); IKD_PRELOAD(set_pool_top_kernel); IKD_PRELOAD(forest_setup_async_kernel); IKD_PRELOAD(forest_setup_kernel);
    IKD_PRELOAD(gather_pid_kernel); IKD_PRELOAD(gather_sorted_kernel); IKD_PRELOAD(gather_u32_kernel);
    IKD_PRELOAD(group_bounds_kernel<uint32_t>); IKD_PRELOAD(group_bounds_kernel<unsigned long long>);
    IKD_PRELOAD(head_flag_kernel<uint32_t>); IKD_PRELOAD(head_flag_kernel<unsigned long long>);
    IKD_PRELOAD(insert_forest_kernel); IKD_PRELOAD(insert_group_kernel); IKD_PRELOAD(insert_place_kernel);
    IKD_PRELOAD(insert_plan_kernel); IKD_PRELOAD(insert_scatter_kernel); IKD_PRELOAD(mark_kernel);
    IKD_PRELOAD(refit_kernel); IKD_PRELOAD(release_list_kernel); IKD_PRELOAD(surv_scan_kernel);
    IKD_PRELOAD(vox_decide_linked_kernel); IKD_PRELOAD(vox_link_kernel); IKD_PRELOAD(voxel_apply_kernel);
    IKD_PRELOAD(voxel_bounds3_kernel); IKD_PRELOAD(voxel_decide_kernel); IKD_PRELOAD(voxel_head3_kernel);
    IKD_PRELOAD(voxel_key64_kernel); IKD_PRELOAD(voxel_key_kernel); IKD_PRELOAD(voxel_plan_kernel); IKD_PRELOAD(vox_prep_kernel);
}
#undef IKD_PRELOAD

}  // namespace ikd

using namespace ikd;

// (a pending side-stream rebuild is committed inside the operation, right before its first mutating kernel: commit_async)
#define CHECK_T2(t)                                                        \
    do {                                                                   \
        if (!(t)) { set_error("null tree handle"); return IKD_ERR_ARG; }   \
        IKD_CUDA(cudaSetDevice((t)->device));                              \
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

int ikd_delete_points_dev(ikd_tree* t, const void* pts_dev_float4, int64_t n) {
    CHECK_T2(t);
    if (n < 0 || (n > 0 && !pts_dev_float4)) { set_error("bad delete_points_dev arguments"); return IKD_ERR_ARG; }
    return delete_points_dev_impl(t, (const float4*)pts_dev_float4, n);
}

int ikd_delete_boxes_dev(ikd_tree* t, const float* boxes_dev, int64_t nb, int* out_deleted) {
    CHECK_T2(t);
    if (nb < 0 || (nb > 0 && !boxes_dev) || !out_deleted) { set_error("bad delete_boxes_dev arguments"); return IKD_ERR_ARG; }
    return delete_boxes_dev_impl(t, boxes_dev, nb, out_deleted);
}

int ikd_add_boxes_dev(ikd_tree* t, const float* boxes_dev, int64_t nb) {
    CHECK_T2(t);
    if (nb < 0 || (nb > 0 && !boxes_dev)) { set_error("bad add_boxes_dev arguments"); return IKD_ERR_ARG; }
    return add_boxes_dev_impl(t, boxes_dev, nb);
}

int ikd_compact_ids(ikd_tree* t, int32_t* old_of_new, int64_t cap_alive, int64_t* out_alive, int32_t* removed_old,
                    int64_t cap_removed, int64_t* out_removed) {
    CHECK_T2(t);
    if (!out_alive || !out_removed || cap_alive < 0 || cap_removed < 0) { set_error("bad compact_ids arguments"); return IKD_ERR_ARG; }
    return compact_ids_impl(t, old_of_new, cap_alive, out_alive, removed_old, cap_removed, out_removed);
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
