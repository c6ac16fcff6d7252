// Batched exact k-nearest-neighbour search (replaces Nearest_Search / Search / MANUAL_HEAP,
// reference ikd_Tree.cpp:367-397, :869-1013, ikd_Tree.h:95-172).
//
// Three traversal kernels share one visit routine and return identical bits:
//   knn_reg_persist_kernel  k <= 8, large batches: one thread per query, persistent warps with dynamic query
//                           hand-out, top-k in registers, traversal stack in shared memory;
//   knn_coop_kernel         k <= 8, scan-sized batches (<= 64k queries): 4 / 16 / 32 lanes per query that pop
//                           several stack entries per memory round trip (latency-bound regime);
//   knn_heap_kernel         k <= 128: one thread per query, binary heap in shared memory.
// Queries are visited in Morton order (radix sort, or a 2-launch counting sort for small batches) so the 32
// lanes of a warp walk nearly the same root-to-leaf paths (their 64 B SearchRec fetches share sectors / L1).
// A visit is ONE 64 B record: the node's point plus both children's AABBs, so the thread scores the
// point, computes calc_box_dist for both children (:1381) and picks nearer-first (:897) without a
// second dependent load. The far child goes on a short per-thread stack together with its box
// distance and is re-checked against the current k-th distance when popped (:915, :952).
// Distances use the reference's exact fp32 operation order with FMA contraction off, so the returned
// squared distances are bit-identical to the reference's; ties at the k-th distance are broken by
// (distance, point id), and a subtree is entered when its box distance EQUALS the current bound so
// that this rule is independent of traversal order.
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "ikd_host.h"

namespace ikd {
namespace {

constexpr int KNN_TPB = 128;
constexpr int STACK_MAX = 64;

__device__ __forceinline__ bool cand_less(float d, int s, float hd, int hs, const UpdateRec* __restrict__ urec) {
    if (d < hd) return true;
    if (d == hd && hs >= 0) return urec[s].pid < urec[hs].pid;  // exact tie: smaller point id first
    return false;
}

struct QueryCtx {
    float qx, qy, qz, T;
};

// ---- register-resident top-K (K == k, exact) -----------------------------------------------------
template <int K, bool COUNT>
__global__ void __launch_bounds__(KNN_TPB)
knn_reg_kernel(const SearchRec* __restrict__ srec, const UpdateRec* __restrict__ urec,
               const TreeHeader* __restrict__ hdr, const float4* __restrict__ q, const int* __restrict__ perm,
               int nq, float T, int32_t* __restrict__ out_idx, float* __restrict__ out_d,
               int32_t* __restrict__ out_cnt, unsigned long long* __restrict__ visits) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int qi = perm ? perm[i] : i;
    float4 qq = q[qi];
    const float qx = qq.x, qy = qq.y, qz = qq.z;
    float hd[K];
    int hs[K];
#pragma unroll
    for (int j = 0; j < K; j++) { hd[j] = CUDART_INF_F; hs[j] = -1; }
    uint32_t st_s[STACK_MAX];
    float st_d[STACK_MAX];
    int sp = 0;
    unsigned int nvis = 0;

    uint32_t cur = 0;
    if (hdr->root_searchable) {
        float d0 = box_sq_dist(qx, qy, qz, hdr->range[0], hdr->range[1], hdr->range[2], hdr->range[3],
                               hdr->range[4], hdr->range[5]);
        if (d0 <= T) cur = ROOT_SLOT;  // reference: cur_dist > max_dist_sqr -> return (:873)
    }
    while (cur) {
        const Rec64 rec = load_rec64_nc(srec + cur);
        const float4 a = rec.a, b = rec.b, c = rec.c, e = rec.e;
        if (COUNT) nvis++;
        uint32_t meta = __float_as_uint(a.w);
        if (!(meta & META_PDEL)) {
            float d = sq_dist3(qx, qy, qz, a.x, a.y, a.z);
            if (d <= T && cand_less(d, (int)cur, hd[K - 1], hs[K - 1], urec)) {
                float cd = d;
                int cs = (int)cur;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    bool sw = cand_less(cd, cs, hd[j], hs[j], urec);
                    float td = hd[j];
                    int ts = hs[j];
                    hd[j] = sw ? cd : td;
                    hs[j] = sw ? cs : ts;
                    cd = sw ? td : cd;
                    cs = sw ? ts : cs;
                }
            }
        }
        float bound = fminf(T, hd[K - 1]);
        uint32_t cp = meta_cp(meta);
        uint32_t next = 0;
        if (cp) {
            float dl = box_sq_dist(qx, qy, qz, b.x, b.y, b.z, b.w, c.x, c.y);
            float dr = box_sq_dist(qx, qy, qz, c.z, c.w, e.x, e.y, e.z, e.w);
            bool okl = dl <= bound && dl < CUDART_INF_F;
            bool okr = dr <= bound && dr < CUDART_INF_F;
            if (okl && okr) {
                if (dl <= dr) { st_s[sp] = 2 * cp + 1; st_d[sp] = dr; sp++; next = 2 * cp; }
                else { st_s[sp] = 2 * cp; st_d[sp] = dl; sp++; next = 2 * cp + 1; }
            } else if (okl) next = 2 * cp;
            else if (okr) next = 2 * cp + 1;
        }
        if (!next) {
            while (sp > 0) {
                --sp;
                if (st_d[sp] <= bound) { next = st_s[sp]; break; }
            }
        }
        cur = next;
    }
    int cnt = 0;
    size_t o = (size_t)qi * K;
#pragma unroll
    for (int j = 0; j < K; j++) {
        int s = hs[j];
        out_idx[o + j] = s >= 0 ? urec[s].pid : -1;
        out_d[o + j] = hd[j];
        cnt += s >= 0 ? 1 : 0;
    }
    out_cnt[qi] = cnt;
    if (COUNT) atomicAdd(visits, (unsigned long long)nvis);
}

// ---- persistent variant: dynamic query fetch + shared-memory traversal stack ---------------------
// ncu on the static kernel above (10M-point map, 4M queries): 12 of 32 lanes active on average and 1.2 GB
// of DRAM writes that are nothing but the per-thread stacks spilling out of L1. Here
//  - every warp claims chunks of consecutive (Morton-ordered) queries from a global counter and a lane that
//    finishes its query immediately takes the next one of the chunk, so lanes do not idle until the slowest
//    query of the warp is done;
//  - the first KNN_SDEPTH stack levels live in shared memory laid out [level][thread] (conflict-free); only
//    deeper levels (unbalanced trees) fall back to a small local array.
constexpr int KNN_SDEPTH = 24;
constexpr int KNN_ODEPTH = 40;  // overflow levels in local memory (KNN_SDEPTH + KNN_ODEPTH >= 64 = depth bound)
#ifndef IKD_KNN_REFILL
#define IKD_KNN_REFILL 8
#endif
constexpr int KNN_REFILL = IKD_KNN_REFILL;   // flush results / hand out new queries once this many lanes of a warp wait
#ifndef IKD_KNN_POPS
#define IKD_KNN_POPS 2
#endif
constexpr int KNN_POPS = IKD_KNN_POPS;  // stack entries a lane without a node may examine per iteration

// OCC > 0: the box distances on the shared stack are kept as bf16 rounded towards zero (a LOWER bound of the distance, so
// the re-check at pop time can only let a few more subtrees through -- results are unaffected); 6 instead of 8 bytes per
// entry and a register cap let OCC blocks live on an SM instead of 9 (which the 24.6 KB of stack per block allow).
template <int K, bool COUNT, int OCC>
__global__ void __launch_bounds__(KNN_TPB, OCC > 0 ? OCC : 1)
knn_reg_persist_kernel(const SearchRec* __restrict__ srec, const UpdateRec* __restrict__ urec,
                       const TreeHeader* __restrict__ hdr, const float4* __restrict__ q, const int* __restrict__ perm,
                       int nq, int chunk, float T, int32_t* __restrict__ out_idx, float* __restrict__ out_d,
                       int32_t* __restrict__ out_cnt, unsigned long long* __restrict__ visits,
                       unsigned int* __restrict__ next_chunk) {
    __shared__ uint32_t sm_s[KNN_SDEPTH][KNN_TPB];
    constexpr bool COMPACT = OCC > 0;
    __shared__ typename std::conditional<COMPACT, unsigned short, float>::type sm_d[KNN_SDEPTH][KNN_TPB];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t ov_s[KNN_ODEPTH];
    float ov_d[KNN_ODEPTH];
    const bool root_ok = hdr->root_searchable != 0;
    const float r0 = hdr->range[0], r1 = hdr->range[1], r2 = hdr->range[2], r3 = hdr->range[3], r4 = hdr->range[4],
                r5 = hdr->range[5];
    int chunk_pos = 0, chunk_end = 0;  // warp-uniform
    bool exhausted = false;            // warp-uniform: the global counter ran past nq
    int qi = -1;                       // query owned by this lane (-1: none)
    float qx = 0.f, qy = 0.f, qz = 0.f;
    float hd[K];
    int hs[K];
    int sp = 0;
    uint32_t cur = 0;
    unsigned int nvis = 0;
    while (true) {
        // A lane whose traversal ended keeps its result in registers and waits; results are flushed and new
        // queries handed out for several lanes at once (when KNN_REFILL lanes wait, or all of them), so the
        // divergent flush/refill code runs rarely and with many lanes active.
        const bool waiting = cur == 0 && sp == 0;
        const unsigned wmask = __ballot_sync(0xffffffffu, waiting);
        const bool more = !(exhausted && chunk_pos >= chunk_end);
        if (wmask == 0xffffffffu || (more && __popc(wmask) >= KNN_REFILL)) {
            if (qi >= 0 && waiting) {
                int cnt = 0;
                size_t o = (size_t)qi * K;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    int s = hs[j];
                    out_idx[o + j] = s >= 0 ? urec[s].pid : -1;
                    out_d[o + j] = hd[j];
                    cnt += s >= 0 ? 1 : 0;
                }
                out_cnt[qi] = cnt;
                qi = -1;
            }
            unsigned idle = __ballot_sync(0xffffffffu, qi < 0);
            if (idle) {
                if (!exhausted && chunk_pos >= chunk_end) {  // claim the next chunk of queries for this warp
                    unsigned int base = 0;
                    if (lane == 0) base = atomicAdd(next_chunk, (unsigned int)chunk);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= (unsigned int)nq) { exhausted = true; chunk_pos = chunk_end = nq; }
                    else { chunk_pos = (int)base; chunk_end = min((int)base + chunk, nq); }
                }
                int avail = chunk_end - chunk_pos;
                if (avail > 0) {
                    int rank = __popc(idle & lt_mask);
                    if (qi < 0 && rank < avail) {
                        int i = chunk_pos + rank;
                        qi = perm ? perm[i] : i;
                        float4 qq = q[qi];
                        qx = qq.x; qy = qq.y; qz = qq.z;
#pragma unroll
                        for (int j = 0; j < K; j++) { hd[j] = CUDART_INF_F; hs[j] = -1; }
                        sp = 0;
                        cur = 0;
                        if (root_ok) {
                            float d0 = box_sq_dist(qx, qy, qz, r0, r1, r2, r3, r4, r5);
                            if (d0 <= T) cur = ROOT_SLOT;  // reference: cur_dist > max_dist_sqr -> return (:873)
                        }
                    }
                    int taken = min(__popc(idle), avail);
                    chunk_pos += taken;
                } else if (exhausted && idle == 0xffffffffu) {
                    break;  // nothing left anywhere and every lane of the warp is done
                }
            }
        }
        // One step per lane and iteration, in lock-step: a lane without a current node takes ONE entry off its
        // stack (entries whose box distance no longer beats the k-th distance are dropped, one per iteration),
        // then every lane that has a node visits it. (A per-lane `while` over stale entries was executed one
        // lane at a time and held 37% of the stall samples of the first version of this kernel.)
        float bound = fminf(T, hd[K - 1]);
        uint32_t node = cur;
        // Up to KNN_POPS entries per iteration: the far siblings deferred during the first descent (bound still infinite)
        // are mostly stale by the time they are popped, and a lane that pops a stale entry idles for the whole visit the
        // other lanes execute; a fixed, unrolled number of tries keeps the popping lanes on one code path.
        // Measured on B200, 100M-point map, same box: 100M queries 61.8 / 58.2 / 59.3 / 61.2 ms with 1 / 2 / 3 / 4 tries,
        // 12.5M queries 9.41 / 9.11 / 9.16 / 9.32 ms.
#pragma unroll
        for (int tr = 0; tr < KNN_POPS; tr++) {
            if (node == 0 && sp > 0) {
                --sp;
                float sd;
                if constexpr (COMPACT) sd = sp < KNN_SDEPTH ? __uint_as_float((uint32_t)sm_d[sp][tid] << 16) : ov_d[sp - KNN_SDEPTH];
                else sd = sp < KNN_SDEPTH ? sm_d[sp][tid] : ov_d[sp - KNN_SDEPTH];
                uint32_t ss = sp < KNN_SDEPTH ? sm_s[sp][tid] : ov_s[sp - KNN_SDEPTH];
                node = sd <= bound ? ss : 0u;
            }
        }
        cur = 0;
        if (node) {
            const Rec64 rec = load_rec64_nc(srec + node);
            const float4 a = rec.a, b = rec.b, c = rec.c, e = rec.e;
            if (COUNT) nvis++;
            uint32_t meta = __float_as_uint(a.w);
            float d = sq_dist3(qx, qy, qz, a.x, a.y, a.z);
            bool live = !(meta & META_PDEL) && d <= T;
            // exact distance tie with a kept neighbour (practically never on real data): id-aware slow path
            bool tie = false;
#pragma unroll
            for (int j = 0; j < K; j++) tie = tie || d == hd[j];  // (an empty slot holds +inf: only an infinite d can match it, and the slow path copes)
            if (live && tie) {
                if (cand_less(d, (int)node, hd[K - 1], hs[K - 1], urec)) {
                    float cd = d;
                    int cs = (int)node;
#pragma unroll
                    for (int j = 0; j < K; j++) {
                        bool sw = cand_less(cd, cs, hd[j], hs[j], urec);
                        float td = hd[j];
                        int ts = hs[j];
                        hd[j] = sw ? cd : td;
                        hs[j] = sw ? cs : ts;
                        cd = sw ? td : cd;
                        cs = sw ? ts : cs;
                    }
                }
            } else {
                // branch-free sorted insertion; a rejected candidate is +inf and falls through unchanged
                float cd = (live && d < hd[K - 1]) ? d : CUDART_INF_F;
                int cs = (int)node;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    bool sw = cd < hd[j];
                    float td = hd[j];
                    int ts = hs[j];
                    hd[j] = sw ? cd : td;
                    hs[j] = sw ? cs : ts;
                    cd = sw ? td : cd;
                    cs = sw ? ts : cs;
                }
            }
            bound = fminf(T, hd[K - 1]);
            uint32_t cp = meta_cp(meta);
            if (cp) {
                float dl = box_sq_dist(qx, qy, qz, b.x, b.y, b.z, b.w, c.x, c.y);
                float dr = box_sq_dist(qx, qy, qz, c.z, c.w, e.x, e.y, e.z, e.w);
                bool okl = dl <= bound && dl < CUDART_INF_F;
                bool okr = dr <= bound && dr < CUDART_INF_F;
                bool left_first = dl <= dr;
                if (okl && okr) {
                    uint32_t fs = left_first ? 2 * cp + 1 : 2 * cp;
                    float fd = left_first ? dr : dl;
                    if (sp < KNN_SDEPTH) {
                        sm_s[sp][tid] = fs;
                        if constexpr (COMPACT) sm_d[sp][tid] = (unsigned short)(__float_as_uint(fd) >> 16);
                        else sm_d[sp][tid] = fd;
                    }
                    else { ov_s[sp - KNN_SDEPTH] = fs; ov_d[sp - KNN_SDEPTH] = fd; }
                    sp++;  // (prefetching the deferred sibling into L2 here was measured: 0.70 -> 0.63 of roofline; not done)
                }
                cur = (okl && (left_first || !okr)) ? 2 * cp : (okr ? 2 * cp + 1 : 0u);
            }
        }
    }
    if (COUNT) atomicAdd(visits, (unsigned long long)nvis);
}

// ---- cooperative variant for small batches: G lanes per query ------------------------------------
// A scan-sized batch (~20k queries) cannot fill the machine with one thread per query: the persistent kernel
// above then runs ~600 warps, each a chain of 35-70 *dependent* 64 B fetches (ncu: 9% of the warp slots active,
// DRAM <1% busy, time = chain length x L2/DRAM latency). Here a query is owned by a group of G lanes that
// share one traversal stack in shared memory and pop up to G entries per iteration: lane 0 follows the
// sequential depth-first order, the other lanes visit the next-best deferred subtrees in the same memory
// round trip, so the dependent chain per query shrinks from "visits" to about "tree depth" and G x more
// fetches are in flight. The k best are kept replicated in every lane's registers; the candidates of one
// iteration are broadcast one after the other with warp shuffles and inserted by all lanes of the group.
// The result does not depend on the visiting order (ties broken by (distance, point id), subtrees entered on
// equality), so it is bit-identical to the one-thread-per-query kernels.
constexpr int COOP_TPB = 128;
constexpr int COOP_SPEC = 64;            // stack fill up to which a group pops G entries at once (beyond: 1)
constexpr int COOP_CAP_BASE = 128;       // COOP_SPEC + depth bound 64; capacity = COOP_CAP_BASE + 2 * G

template <int K, int G, bool COUNT, int VAR>
__global__ void __launch_bounds__(COOP_TPB)
knn_coop_kernel(const SearchRec* __restrict__ srec, const UpdateRec* __restrict__ urec,
                const TreeHeader* __restrict__ hdr, const float4* __restrict__ q, const int* __restrict__ perm,
                int nq, float T, int32_t* __restrict__ out_idx, float* __restrict__ out_d,
                int32_t* __restrict__ out_cnt, unsigned long long* __restrict__ visits) {
    constexpr int CAP = COOP_CAP_BASE + 2 * G;
    constexpr int GPB = COOP_TPB / G;  // groups per block
    constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    __shared__ uint2 stack[GPB][CAP + 2];  // (slot, box distance bits); +2 staggers the groups over the banks
    const int tid = threadIdx.x, lane = tid & 31;
    const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
    const int grp = tid / G;
    const int gq = blockIdx.x * GPB + grp;  // position of the group's query in the (ordered) batch
    uint2* st = stack[grp];
    pdl_wait();
    float qx = 0.f, qy = 0.f, qz = 0.f;
    float hd[K];
    int hs[K];
#pragma unroll
    for (int j = 0; j < K; j++) { hd[j] = CUDART_INF_F; hs[j] = -1; }
    int size = 0;  // entries on the group's stack (replicated in every lane of the group)
    int qi = -1;
    unsigned int nvis = 0;
    if (gq < nq) {
        qi = perm ? perm[gq] : gq;
        float4 qq = q[qi];
        qx = qq.x; qy = qq.y; qz = qq.z;
        if (hdr->root_searchable) {
            float d0 = box_sq_dist(qx, qy, qz, hdr->range[0], hdr->range[1], hdr->range[2], hdr->range[3],
                                   hdr->range[4], hdr->range[5]);
            if (d0 <= T) {  // reference: cur_dist > max_dist_sqr -> return (:873)
                if (gl == 0) st[0] = make_uint2(ROOT_SLOT, __float_as_uint(d0));
                size = 1;
            }
        }
    }
    __syncwarp();
    while (__any_sync(0xffffffffu, size > 0)) {
        float bound = fminf(T, hd[K - 1]);
        // pop: lane i of the group takes the i-th entry from the top
        const int npop = min(size <= COOP_SPEC ? G : 1, size);
        uint32_t node = 0;
        if (gl < npop) {
            uint2 e = st[size - 1 - gl];
            node = __uint_as_float(e.y) <= bound ? e.x : 0u;
        }
        size -= npop;
        float d = CUDART_INF_F, dl = CUDART_INF_F, dr = CUDART_INF_F;
        uint32_t cp = 0;
        if (node) {
            const Rec64 rec = load_rec64_nc(srec + node);
            const float4 a = rec.a, b = rec.b, c = rec.c, e = rec.e;
            if (COUNT) nvis++;
            uint32_t meta = __float_as_uint(a.w);
            float dd = sq_dist3(qx, qy, qz, a.x, a.y, a.z);
            if (!(meta & META_PDEL) && dd <= T) d = dd;
            cp = meta_cp(meta);
            if (cp) {
                dl = box_sq_dist(qx, qy, qz, b.x, b.y, b.z, b.w, c.x, c.y);
                dr = box_sq_dist(qx, qy, qz, c.z, c.w, e.x, e.y, e.z, e.w);
            }
        }
        // merge this iteration's candidates into the replicated top-k, one broadcast at a time
        bool has = d < hd[K - 1] || (d == hd[K - 1] && hs[K - 1] >= 0 && d < CUDART_INF_F);
        unsigned m = (__ballot_sync(0xffffffffu, has) >> gbase) & GM;
        while (__any_sync(0xffffffffu, m != 0)) {
            const int src = m ? __ffs(m) - 1 : gl;
            float cd = __shfl_sync(0xffffffffu, d, gbase + src);
            int cs = (int)__shfl_sync(0xffffffffu, node, gbase + src);
            if (m) {
                m &= m - 1;
                bool tie = false;
#pragma unroll
                for (int j = 0; j < K; j++) tie = tie || cd == hd[j];
                if (tie) {  // exact distance tie with a kept neighbour (practically never on real data): order by point id
                    if (cand_less(cd, cs, hd[K - 1], hs[K - 1], urec)) {
#pragma unroll
                        for (int j = 0; j < K; j++) {
                            bool sw = cand_less(cd, cs, hd[j], hs[j], urec);
                            float td = hd[j];
                            int ts = hs[j];
                            hd[j] = sw ? cd : td;
                            hs[j] = sw ? cs : ts;
                            cd = sw ? td : cd;
                            cs = sw ? ts : cs;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < K; j++) {  // branch-free sorted insertion (a candidate that no longer fits falls through)
                        bool sw = cd < hd[j];
                        float td = hd[j];
                        int ts = hs[j];
                        hd[j] = sw ? cd : td;
                        hs[j] = sw ? cs : ts;
                        cd = sw ? td : cd;
                        cs = sw ? ts : cs;
                    }
                }
            }
        }
        // push the children that can still hold a neighbour: lane G-1's first, lane 0's last (on top), far child
        // below near child, so that lane 0 continues in the sequential nearer-child-first order (:897)
        bound = fminf(T, hd[K - 1]);
        const bool okl = dl <= bound && dl < CUDART_INF_F;
        const bool okr = dr <= bound && dr < CUDART_INF_F;
        const unsigned ml = (__ballot_sync(0xffffffffu, okl) >> gbase) & GM;
        const unsigned mr = (__ballot_sync(0xffffffffu, okr) >> gbase) & GM;
        if (VAR & 1) {
            // order this iteration's children by box distance, nearest on top: position = number of new entries
            // that are farther (ties: left before right, lower lane first)
            const float kl = okl ? dl : -1.f, kr = okr ? dr : -1.f;  // -1: no entry
            int pl = 0, pr = 0;
#pragma unroll
            for (int j = 0; j < G; j++) {
                const float ol = __shfl_sync(0xffffffffu, kl, gbase + j);
                const float orr = __shfl_sync(0xffffffffu, kr, gbase + j);
                // entry order key: (dist, lane, side); "farther" = greater key
                pl += (ol > kl || (ol == kl && j > gl)) ? 1 : 0;
                pl += (orr > kl || (orr == kl && j >= gl)) ? 1 : 0;
                pr += (ol > kr || (ol == kr && j > gl)) ? 1 : 0;
                pr += (orr > kr || (orr == kr && j > gl)) ? 1 : 0;
            }
            __syncwarp();
            if (okl) st[size + pl] = make_uint2(2 * cp, __float_as_uint(dl));
            if (okr) st[size + pr] = make_uint2(2 * cp + 1, __float_as_uint(dr));
        } else {
            const unsigned above = gl == G - 1 ? 0u : (GM >> (gl + 1)) << (gl + 1);  // lanes of the group after this one
            int pos = size + __popc(ml & above) + __popc(mr & above);
            __syncwarp();  // every pop of this iteration is done before entries are overwritten
            if (okl && okr) {
                const bool left_first = dl <= dr;
                st[pos] = left_first ? make_uint2(2 * cp + 1, __float_as_uint(dr)) : make_uint2(2 * cp, __float_as_uint(dl));
                st[pos + 1] = left_first ? make_uint2(2 * cp, __float_as_uint(dl)) : make_uint2(2 * cp + 1, __float_as_uint(dr));
            } else if (okl) {
                st[pos] = make_uint2(2 * cp, __float_as_uint(dl));
            } else if (okr) {
                st[pos] = make_uint2(2 * cp + 1, __float_as_uint(dr));
            }
        }
        size += __popc(ml) + __popc(mr);
        __syncwarp();  // pushes visible to the group's next pops
    }
    if (qi >= 0) {
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < K; j++) cnt += hs[j] >= 0 ? 1 : 0;
        size_t o = (size_t)qi * K;
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (gl == (j % G)) {
                int s = hs[j];
                out_idx[o + j] = s >= 0 ? urec[s].pid : -1;
                out_d[o + j] = hd[j];
            }
        }
        if (gl == 0) out_cnt[qi] = cnt;
    }
    if (COUNT) atomicAdd(visits, (unsigned long long)nvis);
}

// ---- shared-memory binary max-heap for larger k --------------------------------------------------
// heap element j of thread t lives at [j * blockDim.x + t] (conflict-free columns)
template <bool COUNT>
__global__ void knn_heap_kernel(const SearchRec* __restrict__ srec, const UpdateRec* __restrict__ urec,
                                const TreeHeader* __restrict__ hdr, const float4* __restrict__ q,
                                const int* __restrict__ perm, int nq, int k, float T,
                                int32_t* __restrict__ out_idx, float* __restrict__ out_d,
                                int32_t* __restrict__ out_cnt, unsigned long long* __restrict__ visits) {
    extern __shared__ unsigned char smem_raw[];
    const int B = blockDim.x, tid = threadIdx.x;
    float* hd = reinterpret_cast<float*>(smem_raw);
    int* hs = reinterpret_cast<int*>(smem_raw + sizeof(float) * (size_t)k * B);
#define HD(j) hd[(j)*B + tid]
#define HS(j) hs[(j)*B + tid]
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int qi = perm ? perm[i] : i;
    float4 qq = q[qi];
    const float qx = qq.x, qy = qq.y, qz = qq.z;
    int cnt = 0;
    uint32_t st_s[STACK_MAX];
    float st_d[STACK_MAX];
    int sp = 0;
    unsigned int nvis = 0;
    uint32_t cur = 0;
    if (hdr->root_searchable) {
        float d0 = box_sq_dist(qx, qy, qz, hdr->range[0], hdr->range[1], hdr->range[2], hdr->range[3],
                               hdr->range[4], hdr->range[5]);
        if (d0 <= T) cur = ROOT_SLOT;
    }
    // lock-step traversal: a lane without a current node takes ONE stack entry per iteration (see the note in
    // knn_reg_persist_kernel); stale entries are dropped one per iteration instead of in a per-lane loop
    while (cur || sp > 0) {
        if (!cur) {
            --sp;
            float bnd = (cnt >= k) ? fminf(T, HD(0)) : T;
            if (st_d[sp] <= bnd) cur = st_s[sp];
            if (!cur) continue;
        }
        const Rec64 rec = load_rec64_nc(srec + cur);
        const float4 a = rec.a, b = rec.b, c = rec.c, e = rec.e;
        if (COUNT) nvis++;
        uint32_t meta = __float_as_uint(a.w);
        if (!(meta & META_PDEL)) {
            float d = sq_dist3(qx, qy, qz, a.x, a.y, a.z);
            if (d <= T) {
                if (cnt < k) {
                    int j = cnt++;  // sift up
                    while (j > 0) {
                        int pj = (j - 1) >> 1;
                        float pd = HD(pj);
                        int ps = HS(pj);
                        if (cand_less(pd, ps, d, (int)cur, urec)) { HD(j) = pd; HS(j) = ps; j = pj; }
                        else break;
                    }
                    HD(j) = d; HS(j) = (int)cur;
                } else if (cand_less(d, (int)cur, HD(0), HS(0), urec)) {
                    int j = 0;  // replace the maximum, sift down
                    while (true) {
                        int l = 2 * j + 1;
                        if (l >= k) break;
                        float ld = HD(l);
                        int ls = HS(l);
                        if (l + 1 < k) {
                            float rd = HD(l + 1);
                            int rs = HS(l + 1);
                            if (cand_less(ld, ls, rd, rs, urec)) { l = l + 1; ld = rd; ls = rs; }
                        }
                        if (cand_less(d, (int)cur, ld, ls, urec)) { HD(j) = ld; HS(j) = ls; j = l; }
                        else break;
                    }
                    HD(j) = d; HS(j) = (int)cur;
                }
            }
        }
        float bound = (cnt >= k) ? fminf(T, HD(0)) : T;
        uint32_t cp = meta_cp(meta);
        uint32_t next = 0;
        if (cp) {
            float dl = box_sq_dist(qx, qy, qz, b.x, b.y, b.z, b.w, c.x, c.y);
            float dr = box_sq_dist(qx, qy, qz, c.z, c.w, e.x, e.y, e.z, e.w);
            bool okl = dl <= bound && dl < CUDART_INF_F;
            bool okr = dr <= bound && dr < CUDART_INF_F;
            if (okl && okr) {
                if (dl <= dr) { st_s[sp] = 2 * cp + 1; st_d[sp] = dr; sp++; next = 2 * cp; }
                else { st_s[sp] = 2 * cp; st_d[sp] = dl; sp++; next = 2 * cp + 1; }
            } else if (okl) next = 2 * cp;
            else if (okr) next = 2 * cp + 1;
        }
        cur = next;
    }
    // heap-sort: pop the maximum into the tail so the row ends up ascending
    size_t o = (size_t)qi * k;
    int m = cnt;
    for (int j = cnt; j < k; j++) { out_idx[o + j] = -1; out_d[o + j] = CUDART_INF_F; }
    while (m > 0) {
        float td = HD(0);
        int ts = HS(0);
        out_idx[o + m - 1] = urec[ts].pid;
        out_d[o + m - 1] = td;
        m--;
        if (m == 0) break;
        float d = HD(m);
        int s = HS(m);
        int j = 0;
        while (true) {
            int l = 2 * j + 1;
            if (l >= m) break;
            float ld = HD(l);
            int ls = HS(l);
            if (l + 1 < m) {
                float rd = HD(l + 1);
                int rs = HS(l + 1);
                if (cand_less(ld, ls, rd, rs, urec)) { l = l + 1; ld = rd; ls = rs; }
            }
            if (cand_less(d, s, ld, ls, urec)) { HD(j) = ld; HS(j) = ls; j = l; }
            else break;
        }
        HD(j) = d; HS(j) = s;
    }
    out_cnt[qi] = cnt;
    if (COUNT) atomicAdd(visits, (unsigned long long)nvis);
#undef HD
#undef HS
}

// ---- Morton ordering of the queries --------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void morton_kernel(const float4* __restrict__ q, int nq, const TreeHeader* __restrict__ hdr,
                              uint32_t* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float4 v = q[i];
    float c[3] = {v.x, v.y, v.z};
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = hdr->range[a], hi = hdr->range[3 + a];
        float ext = hi - lo;
        float u = ext > 0.f ? (c[a] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);  // NaN -> 0
        uint32_t g = (uint32_t)(u * 1023.f);
        code |= spread10(g) << a;
    }
    keys[i] = code;
    vals[i] = i;
}

// ---- cheap query ordering for small batches: counting sort on a 4096-cell grid over the tree's range --------
// (a full radix sort of 20k keys costs ~45 us in six latency-bound launches; coarse cells are all a small batch
// needs -- scan points arrive spatially coherent already -- and the order inside a cell does not influence any
// result). The 12 cell-index bits are split over the axes by the host so that cells are roughly cubic. Two
// launches: count (atomic rank inside the cell), then scatter, where every block redoes the 4096-entry exclusive
// scan in shared memory instead of waiting for a separate single-block scan kernel.
constexpr int BIN_BITS = 12;
constexpr int NBINS = 1 << BIN_BITS;
struct BinGrid { int bits[3]; };
__global__ void bin_count_kernel(const float4* __restrict__ q, int nq, const TreeHeader* __restrict__ hdr, BinGrid bg,
                                 unsigned int* __restrict__ hist, uint32_t* __restrict__ bin_of, int* __restrict__ rank) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float4 v = q[i];
    float c[3] = {v.x, v.y, v.z};
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = hdr->range[a], hi = hdr->range[3 + a];
        float ext = hi - lo;
        float u = ext > 0.f ? (c[a] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);  // NaN -> 0
        uint32_t cells = 1u << bg.bits[a];
        uint32_t g = min((uint32_t)(u * (float)cells), cells - 1u);
        code = (code << bg.bits[a]) | g;
    }
    bin_of[i] = code;
    // scan points crowd into few cells: the lanes of a warp that hit the same cell share one atomic
    const unsigned peers = __match_any_sync(__activemask(), code);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(&hist[code], (unsigned int)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    rank[i] = (int)(base + __popc(peers & ((1u << lane) - 1u)));
}
__global__ void __launch_bounds__(256)
bin_scatter_kernel(const uint32_t* __restrict__ bin_of, const int* __restrict__ rank, const unsigned int* __restrict__ hist,
                   int nq, int* __restrict__ perm, unsigned int* __restrict__ hist_next) {
    pdl_wait();
    __shared__ unsigned int off[NBINS];
    __shared__ unsigned int wsum[8];
    const int t = threadIdx.x;
    // the two histograms alternate between calls: clear the one the next call will count into (saves a memset node)
    if (blockIdx.x == 0)
        for (int j = 0; j < NBINS / 256 / 4; j++) reinterpret_cast<uint4*>(hist_next)[j * 256 + t] = make_uint4(0u, 0u, 0u, 0u);
    // exclusive scan of the cell counts: 16 consecutive cells per thread
    uint4 v[4];
    unsigned int tot = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] = reinterpret_cast<const uint4*>(hist)[t * 4 + j]; tot += v[j].x + v[j].y + v[j].z + v[j].w; }
    unsigned int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int n = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += n;
    }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    unsigned int run = inc - tot;
    for (int w = 0; w < (t >> 5); w++) run += wsum[w];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint4 o;
        o.x = run; run += v[j].x;
        o.y = run; run += v[j].y;
        o.z = run; run += v[j].z;
        o.w = run; run += v[j].w;
        reinterpret_cast<uint4*>(off)[t * 4 + j] = o;
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + t;
    if (i < nq) perm[off[bin_of[i]] + rank[i]] = i;
}

template <int K, int G>
void launch_coop(bool count, int nq, cudaStream_t s, const SearchRec* srec, const UpdateRec* urec,
                 const TreeHeader* hdr, const float4* q, const int* perm, float T, int32_t* oi, float* od, int32_t* oc,
                 unsigned long long* vis) {
    constexpr int GPB = COOP_TPB / G;
    int blocks = (nq + GPB - 1) / GPB;
    constexpr int V = G == 4 ? 1 : 0;  // distance-ordered pushes pay off only for narrow groups (measured)
    if (count) IKD_LAUNCH_PDL((knn_coop_kernel<K, G, true, V>), blocks, COOP_TPB, 0, s, srec, urec, hdr, q, perm, nq, T, oi, od, oc, vis);
    else IKD_LAUNCH_PDL((knn_coop_kernel<K, G, false, V>), blocks, COOP_TPB, 0, s, srec, urec, hdr, q, perm, nq, T, oi, od, oc, vis);
}

// lanes per query for a batch of n queries: enough groups to fill ~150k thread slots, 0 = one thread per query
int coop_group(int n) {
    static int forced = getenv("IKD_KNN_G") ? atoi(getenv("IKD_KNN_G")) : -1;
    if (forced >= 0) return forced;
    // measured on B200 (1M-point map, L2 flushed, tools/gpu_knn_small.py): 2k queries 121 -> 53 us with 32 lanes,
    // 5k 112 -> 56 us with 16, 10k 121 -> 73 us with 16, 20k 121 -> 74 us with 4, 50k 132 -> 119 us with 4
    if (n <= 3000) return 32;
    if (n <= 12000) return 16;
    if (n <= 64000) return 4;
    return 0;
}

template <int K>
void launch_reg(bool count, int nq, cudaStream_t s, const SearchRec* srec, const UpdateRec* urec,
                const TreeHeader* hdr, const float4* q, const int* perm, float T, int32_t* oi, float* od, int32_t* oc,
                unsigned long long* vis, unsigned int* next_chunk) {
    static int v1 = getenv("IKD_KNN_V1") ? atoi(getenv("IKD_KNN_V1")) : 0;
    const int G = coop_group(nq);
    if (G == 4) return launch_coop<K, 4>(count, nq, s, srec, urec, hdr, q, perm, T, oi, od, oc, vis);
    if (G == 8) return launch_coop<K, 8>(count, nq, s, srec, urec, hdr, q, perm, T, oi, od, oc, vis);
    if (G == 16) return launch_coop<K, 16>(count, nq, s, srec, urec, hdr, q, perm, T, oi, od, oc, vis);
    if (G == 32) return launch_coop<K, 32>(count, nq, s, srec, urec, hdr, q, perm, T, oi, od, oc, vis);
    if (v1 || !next_chunk) {
        int blocks = (nq + KNN_TPB - 1) / KNN_TPB;
        if (count) IKD_LAUNCH knn_reg_kernel<K, true><<<blocks, KNN_TPB, 0, s>>>(srec, urec, hdr, q, perm, nq, T, oi, od, oc, vis);
        else IKD_LAUNCH knn_reg_kernel<K, false><<<blocks, KNN_TPB, 0, s>>>(srec, urec, hdr, q, perm, nq, T, oi, od, oc, vis);
        return;
    }
    // persistent launch: at most 148 SMs x resident blocks; chunk = queries a warp claims at a time
    // 10 blocks per SM with the compact stack (measured on the 100M / 100M batch, same box: 9 blocks 1.333 G q/s, 10
    // blocks 1.434, 11 and 12 blocks (40 registers, spills) 1.146 / 1.131); IKD_KNN_OCC=0 selects the 9-block kernel
    static int occ = getenv("IKD_KNN_OCC") ? atoi(getenv("IKD_KNN_OCC")) : 10;
    const int max_blocks = 148 * (occ == 10 ? 10 : 9);
    int blocks = std::min((nq + KNN_TPB - 1) / KNN_TPB, max_blocks);
    int warps = blocks * (KNN_TPB / 32);
    int chunk = (nq + warps - 1) / warps;
    chunk = std::max(32, std::min(256, (chunk + 31) / 32 * 32));
#define PERSIST_LAUNCH(CNT, OC) \
    IKD_LAUNCH knn_reg_persist_kernel<K, CNT, OC><<<blocks, KNN_TPB, 0, s>>>(srec, urec, hdr, q, perm, nq, chunk, T, oi, od, oc, vis, next_chunk)
    if (count) PERSIST_LAUNCH(true, 0);
    else if (occ == 10) PERSIST_LAUNCH(false, 10);
    else PERSIST_LAUNCH(false, 0);
#undef PERSIST_LAUNCH
}

}  // namespace

// largest float T with (double)T <= max_dist*max_dist, so that `d <= T` in fp32 equals the reference's
// `dist <= max_dist_sqr` in double (ikd_Tree.cpp:872, :887)
static float max_dist_threshold(double max_dist) {
    double m2 = max_dist * max_dist;
    if (isnan(m2)) return NAN;
    if (isinf(m2)) return INFINITY;
    float T = (float)m2;
    if ((double)T > m2) T = nextafterf(T, -INFINITY);
    return T;
}

// raw xyz triples (stride 12) -> float4 queries
__global__ void pack_queries_kernel(const float* __restrict__ q3, int n, float4* __restrict__ q4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q4[i] = make_float4(q3[3 * (size_t)i], q3[3 * (size_t)i + 1], q3[3 * (size_t)i + 2], 0.f);
}
int pack_queries(const float* q3_dev, int64_t n, float4* q4_dev, cudaStream_t s) {
    if (n <= 0) return IKD_OK;
    IKD_LAUNCH pack_queries_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(q3_dev, (int)n, q4_dev);
    return IKD_OK;
}

int knn_launch(ikd_tree* t, const float4* q_dev, int64_t nq, int k, double max_dist, int32_t* out_idx,
               float* out_d, int32_t* out_cnt, cudaStream_t s, int lane) {
    if (nq <= 0) return IKD_OK;
    if (k < 1 || k > IKD_MAX_K) { set_error("k=%d out of range [1,%d]", k, IKD_MAX_K); return IKD_ERR_ARG; }
    if (nq > 0x7fffffff) { set_error("nq too large for one call"); return IKD_ERR_ARG; }
    float T = max_dist_threshold(max_dist);
    int n = (int)nq;
    static int no_morton = getenv("IKD_NO_MORTON") ? atoi(getenv("IKD_NO_MORTON")) : 0;
    const int* perm = nullptr;
    KnnScratch& sc = t->knn_scr[lane];
    static int bin_max = getenv("IKD_KNN_BIN_MAX") ? atoi(getenv("IKD_KNN_BIN_MAX")) : (1 << 18);
    if (!no_morton && n >= 256 && n < bin_max) {
        // small batch: counting sort on a 4096-cell grid (memset + 2 short kernels)
        IKD_TRY(sc.mkeys.ensure(sizeof(uint32_t) * (size_t)n, s));
        IKD_TRY(sc.perm.ensure(sizeof(int) * (size_t)n, s));
        IKD_TRY(sc.perm2.ensure(sizeof(int) * (size_t)n, s));
        if (!sc.hist.p) {  // two histograms, zeroed once; afterwards each call clears the other one
            IKD_TRY(sc.hist.ensure(sizeof(unsigned int) * NBINS * 2, s));
            IKD_CUDA(cudaMemsetAsync(sc.hist.p, 0, sizeof(unsigned int) * NBINS * 2, s));
            sc.hist_sel = 0;
        }
        unsigned int* hist_cur = sc.hist.as<unsigned int>() + (size_t)sc.hist_sel * NBINS;
        unsigned int* hist_nxt = sc.hist.as<unsigned int>() + (size_t)(sc.hist_sel ^ 1) * NBINS;
        sc.hist_sel ^= 1;
        BinGrid bg = {{0, 0, 0}};
        {
            double ext[3];
            for (int a = 0; a < 3; a++) ext[a] = std::max((double)t->hdr.range[3 + a] - (double)t->hdr.range[a], 1e-30);
            for (int b = 0; b < BIN_BITS; b++) {  // next bit to the axis whose cells are currently the longest
                int best = 0;
                for (int a = 1; a < 3; a++)
                    if (ext[a] / (double)(1 << bg.bits[a]) > ext[best] / (double)(1 << bg.bits[best])) best = a;
                bg.bits[best]++;
            }
        }
        IKD_LAUNCH_PDL((bin_count_kernel), (n + 255) / 256, 256, 0, s, q_dev, n, t->hdr_dev, bg, hist_cur,
                                                                    sc.mkeys.as<uint32_t>(), sc.perm.as<int>());
        IKD_LAUNCH_PDL((bin_scatter_kernel), (n + 255) / 256, 256, 0, s, sc.mkeys.as<uint32_t>(), sc.perm.as<int>(),
                                                                      hist_cur, n, sc.perm2.as<int>(), hist_nxt);
        perm = sc.perm2.as<int>();
    } else if (!no_morton && n >= 1024) {
        IKD_TRY(sc.mkeys.ensure(sizeof(uint32_t) * (size_t)n, s));
        IKD_TRY(sc.mkeys2.ensure(sizeof(uint32_t) * (size_t)n, s));
        IKD_TRY(sc.perm.ensure(sizeof(int) * (size_t)n, s));
        IKD_TRY(sc.perm2.ensure(sizeof(int) * (size_t)n, s));
        IKD_LAUNCH morton_kernel<<<(n + 255) / 256, 256, 0, s>>>(q_dev, n, t->hdr_dev, sc.mkeys.as<uint32_t>(),
                                                                 sc.perm.as<int>());
        size_t tmp = 0;
        // small batches only need coarse coherence: sort on the top 24 of the 30 Morton bits (3 radix passes)
        const int lo_bit = n < (1 << 20) ? 6 : 0;
        IKD_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, int>(nullptr, tmp, nullptr, nullptr, nullptr, nullptr, n,
                                                                  lo_bit, 30, s)));
        IKD_TRY(sc.cubtmp.ensure(tmp, s));
        size_t tb = sc.cubtmp.bytes;
        IKD_CUDA((cub::DeviceRadixSort::SortPairs<uint32_t, int>(sc.cubtmp.p, tb, sc.mkeys.as<uint32_t>(),
                                                                  sc.mkeys2.as<uint32_t>(), sc.perm.as<int>(),
                                                                  sc.perm2.as<int>(), n, lo_bit, 30, s)));
        perm = sc.perm2.as<int>();
    }
    unsigned long long* vis = nullptr;
    if (t->count_visits) {
        IKD_TRY(t->b_visits.ensure(sizeof(unsigned long long), s));
        vis = t->b_visits.as<unsigned long long>();
        IKD_CUDA(cudaMemsetAsync(vis, 0, sizeof(unsigned long long), s));
    }
    bool cv = t->count_visits;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (t->time_kernels) {
        IKD_CUDA(cudaEventCreate(&ev0));
        IKD_CUDA(cudaEventCreate(&ev1));
        IKD_CUDA(cudaEventRecord(ev0, s));
    }
    IKD_TRY(sc.counter.ensure(sizeof(unsigned int), s));
    if (k <= 8 && coop_group(n) == 0) IKD_CUDA(cudaMemsetAsync(sc.counter.p, 0, sizeof(unsigned int), s));  // persistent kernel only
    unsigned int* next_chunk = sc.counter.as<unsigned int>();
#define REG_CASE(KK) \
    case KK: launch_reg<KK>(cv, n, s, t->srec, t->urec, t->hdr_dev, q_dev, perm, T, out_idx, out_d, out_cnt, vis, next_chunk); break;
    switch (k) {
        REG_CASE(1) REG_CASE(2) REG_CASE(3) REG_CASE(4) REG_CASE(5) REG_CASE(6) REG_CASE(7) REG_CASE(8)
        default: {
            // Persistent scheduling (dynamic query hand-out as in knn_reg_persist_kernel) was measured twice for
            // k = 32 and lost both times: 1.9x slower with a shared-memory stack (occupancy halves), 1.7x slower with
            // the stack in local memory (2M queries on a 1M-point map: 10.9 ms vs 6.35 ms). Static assignment stays.
            int tpb = k <= 32 ? 128 : 64;
            size_t smem = (size_t)k * tpb * 8;
            auto kern = cv ? knn_heap_kernel<true> : knn_heap_kernel<false>;
            IKD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            IKD_LAUNCH kern<<<(n + tpb - 1) / tpb, tpb, smem, s>>>(t->srec, t->urec, t->hdr_dev, q_dev, perm, n, k, T, out_idx,
                                                       out_d, out_cnt, vis);
        }
    }
#undef REG_CASE
    if (t->time_kernels) {
        IKD_CUDA(cudaEventRecord(ev1, s));
        t->timing_events.emplace_back(ev0, ev1);
    }
    IKD_CUDA(cudaGetLastError());
    return IKD_OK;
}

// Load every kernel of this file now (CUDA loads kernels lazily at their first launch, 0.1-0.3 ms each, which
// showed up as milliseconds of extra latency in the first update after Build).
#define IKD_PRELOAD(fn) do { cudaFuncAttributes a_; if (cudaFuncGetAttributes(&a_, fn) != cudaSuccess) cudaGetLastError(); } while (0)
void preload_knn_kernels() {
    // the k = 5 family (FAST-LIO2's query) and the ordering kernels; other k load at first use
    IKD_PRELOAD((knn_coop_kernel<5, 4, false, 1>)); IKD_PRELOAD((knn_coop_kernel<5, 16, false, 0>));
    IKD_PRELOAD((knn_coop_kernel<5, 32, false, 0>)); IKD_PRELOAD((knn_reg_persist_kernel<5, false, 10>));
    IKD_PRELOAD(bin_count_kernel); IKD_PRELOAD(bin_scatter_kernel); IKD_PRELOAD(morton_kernel); IKD_PRELOAD(pack_queries_kernel);
}
#undef IKD_PRELOAD

}  // namespace ikd
