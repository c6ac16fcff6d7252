/*
 * ikd_oracle.c -- ORACLE / TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from the
 * product path (libikd_b200.so, include/ikd_Tree.h). Users: tests/, __graft_entry__.smoke(), and
 * bench.py's cpu_baseline leg when oracle/_ref is unavailable.
 *
 * A plain-C, single-threaded restatement of the reference algorithm of hku-mars/ikd-Tree for the hot
 * path (ikd-Tree/ikd_Tree.cpp, cited per function as :line). Nodes live in an index-addressed pool
 * instead of `new`-ed structs; there is no background rebuild thread, no locks and no operation log:
 * every Rebuild (:625) runs inline, whatever the subtree size. Arithmetic follows the reference
 * exactly (fp32, left-to-right, compiled with -ffp-contract=off, the two double-precision spots kept).
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this restatement
 * is pinned against the reference ITSELF: tests/test_oracle.py compares it with
 * oracle/_ref/libikd_ref.so (the unmodified ikd_Tree.cpp compiled in place) on seeded inputs -- tree
 * structure after Build, kNN distances, box/radius result sets, delete counts, Add_Points return values
 * and valid point sets -- and tests/golden/ holds vectors generated from that library by
 * tests/golden/make_golden.py for boxes where the reference sources are absent.
 * Known, documented divergences: (1) nth_element (:602-611) is libstdc++ introselect in the reference and
 * a median-of-three quickselect here, so inputs with duplicate split coordinates may place tied points
 * on different sides; (2) subtrees of >= 1500 points are rebuilt inline here and by a racing background
 * thread in the reference (:627-633), so TREE STRUCTURE after streaming updates is not comparable
 * (it is not reproducible between two runs of the reference either); point sets,
 * counters and query results are.
 */
#include "ikd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPSS 1e-6                       /* ikd_Tree.h:13 */
#define MIN_UNBALANCED_TREE_SIZE 10     /* ikd_Tree.h:14 */
#define NIL (-1)

typedef struct {
    float p[3];
    int axis;
    int size, invalid, down_del;        /* TreeSize, invalid_point_num, down_del_num (ikd_Tree.h:67-69) */
    unsigned char pdel, tdel, pds, tds; /* point_deleted, tree_deleted, *_downsample_deleted (:70-73) */
    unsigned char pushl, pushr;         /* need_push_down_to_left / right (:74-75) */
    float rmin[3], rmax[3];             /* node_range_{x,y,z} (:79) */
    float radius_sq;
    int left, right, father;
    float alpha_del, alpha_bal;
} Node;

typedef struct { float p[3]; } Pt;
typedef struct { Pt* v; long n, cap; } PtVec;

struct ikdo_tree {
    Node* nd;
    int ncap, nused;
    int* freelist;
    int nfree, freecap;
    int root;
    float del_param, bal_param, ds;
    PtVec scratch;      /* PCL_Storage */
    PtVec down;         /* Downsample_Storage */
    PtVec removed;      /* Points_deleted */
    PtVec last;         /* last search result */
    int rebuilds;
};

/* ---------------------------------------------------------------- small containers */
static void pv_push(PtVec* v, const float* p) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 256;
        v->v = (Pt*)realloc(v->v, sizeof(Pt) * (size_t)v->cap);
    }
    v->v[v->n].p[0] = p[0]; v->v[v->n].p[1] = p[1]; v->v[v->n].p[2] = p[2];
    v->n++;
}
static long pv_copy(const PtVec* v, float* out, long cap) {
    long m = v->n < cap ? v->n : cap;
    for (long i = 0; i < m; i++) { out[3 * i] = v->v[i].p[0]; out[3 * i + 1] = v->v[i].p[1]; out[3 * i + 2] = v->v[i].p[2]; }
    return v->n;
}

static int node_alloc(ikdo_tree* t) {
    int i;
    if (t->nfree > 0) i = t->freelist[--t->nfree];
    else {
        if (t->nused == t->ncap) {
            t->ncap = t->ncap ? t->ncap * 2 : 1024;
            t->nd = (Node*)realloc(t->nd, sizeof(Node) * (size_t)t->ncap);
        }
        i = t->nused++;
    }
    /* InitTreeNode :52-76 */
    Node* n = &t->nd[i];
    memset(n, 0, sizeof(*n));
    n->left = n->right = n->father = NIL;
    n->size = 0;
    return i;
}
static void node_free(ikdo_tree* t, int i) {
    if (t->nfree == t->freecap) {
        t->freecap = t->freecap ? t->freecap * 2 : 1024;
        t->freelist = (int*)realloc(t->freelist, sizeof(int) * (size_t)t->freecap);
    }
    t->freelist[t->nfree++] = i;
}

/* ---------------------------------------------------------------- arithmetic */
/* calc_dist :1374-1378 */
static float calc_dist(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return dx * dx + dy * dy + dz * dz;
}
/* calc_box_dist :1381-1391 (node == null -> INFINITY) */
static float calc_box_dist(const ikdo_tree* t, int ni, const float* q) {
    if (ni == NIL) return INFINITY;
    const Node* n = &t->nd[ni];
    float d = 0.0f;
    for (int a = 0; a < 3; a++) {
        if (q[a] < n->rmin[a]) d += (q[a] - n->rmin[a]) * (q[a] - n->rmin[a]);
        if (q[a] > n->rmax[a]) d += (q[a] - n->rmax[a]) * (q[a] - n->rmax[a]);
    }
    return d;
}
/* same_point :1369-1371 (float difference, compared against the double 1e-6) */
static int same_point(const float* a, const float* b) {
    return (double)fabsf(a[0] - b[0]) < EPSS && (double)fabsf(a[1] - b[1]) < EPSS && (double)fabsf(a[2] - b[2]) < EPSS;
}
static float fmin2(float a, float b) { return b < a ? b : a; }  /* std::min */
static float fmax2(float a, float b) { return a < b ? b : a; }  /* std::max */

/* ---------------------------------------------------------------- Push_Down :1110-1181 */
static void push_to_child(ikdo_tree* t, const Node* r, int ci) {
    Node* c = &t->nd[ci];
    c->tds |= r->tds;
    c->pds |= r->tds;
    c->tdel = r->tdel || c->tds;
    c->pdel = c->tdel || c->pds;
    if (r->tds) c->down_del = c->size;
    if (r->tdel) c->invalid = c->size;
    else c->invalid = c->down_del;
    c->pushl = 1;
    c->pushr = 1;
}
static void push_down(ikdo_tree* t, int ri) {
    if (ri == NIL) return;
    Node* r = &t->nd[ri];
    if (r->pushl && r->left != NIL) { push_to_child(t, r, r->left); r->pushl = 0; }
    if (r->pushr && r->right != NIL) { push_to_child(t, r, r->right); r->pushr = 0; }
}

/* ---------------------------------------------------------------- Update :1184-1323 */
static void update(ikdo_tree* t, int ri) {
    Node* r = &t->nd[ri];
    Node* L = r->left != NIL ? &t->nd[r->left] : NULL;
    Node* R = r->right != NIL ? &t->nd[r->right] : NULL;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    r->size = 1 + (L ? L->size : 0) + (R ? R->size : 0);
    r->invalid = (L ? L->invalid : 0) + (R ? R->invalid : 0) + (r->pdel ? 1 : 0);
    r->down_del = (L ? L->down_del : 0) + (R ? R->down_del : 0) + (r->pds ? 1 : 0);
    r->tds = (unsigned char)((L ? L->tds : 1) & (R ? R->tds : 1) & r->pds);
    r->tdel = (unsigned char)((L ? L->tdel : 1) && (R ? R->tdel : 1) && r->pdel);
    if (!L && !R) {
        for (int a = 0; a < 3; a++) { mn[a] = r->p[a]; mx[a] = r->p[a]; }   /* :1299-1304 */
    } else {
        int all = r->tdel || (!(L && L->tdel) && !(R && R->tdel) && !r->pdel); /* :1197, :1236, :1268 */
        if (L && (all || !L->tdel)) for (int a = 0; a < 3; a++) { mn[a] = fmin2(mn[a], L->rmin[a]); mx[a] = fmax2(mx[a], L->rmax[a]); }
        if (R && (all || !R->tdel)) for (int a = 0; a < 3; a++) { mn[a] = fmin2(mn[a], R->rmin[a]); mx[a] = fmax2(mx[a], R->rmax[a]); }
        if (all || !r->pdel) for (int a = 0; a < 3; a++) { mn[a] = fmin2(mn[a], r->p[a]); mx[a] = fmax2(mx[a], r->p[a]); }
    }
    memcpy(r->rmin, mn, sizeof(mn));
    memcpy(r->rmax, mx, sizeof(mx));
    {   /* :1309-1312 */
        float xl = (r->rmax[0] - r->rmin[0]) * 0.5f, yl = (r->rmax[1] - r->rmin[1]) * 0.5f, zl = (r->rmax[2] - r->rmin[2]) * 0.5f;
        r->radius_sq = xl * xl + yl * yl + zl * zl;
    }
    if (L) L->father = ri;
    if (R) R->father = ri;
    if (ri == t->root && r->size > 3) {   /* :1315-1321 */
        Node* son = L ? L : R;
        float tb = (float)son->size / (float)(r->size - 1);
        r->alpha_del = (float)r->invalid / (float)r->size;
        r->alpha_bal = ((double)tb >= 0.5 - EPSS) ? tb : 1 - tb;
    }
}

/* ---------------------------------------------------------------- BuildTree :574-622 */
static void swap_pt(Pt* a, Pt* b) { Pt t = *a; *a = *b; *b = t; }
/* nth_element replacement: iterative quickselect, median-of-three pivot, on v[l..r] by coordinate `ax` */
static void select_nth(Pt* v, long l, long r, long nth, int ax) {
    while (l < r) {
        long m = l + (r - l) / 2;
        if (v[m].p[ax] < v[l].p[ax]) swap_pt(&v[m], &v[l]);
        if (v[r].p[ax] < v[l].p[ax]) swap_pt(&v[r], &v[l]);
        if (v[r].p[ax] < v[m].p[ax]) swap_pt(&v[r], &v[m]);
        float piv = v[m].p[ax];
        long i = l, j = r;
        while (i <= j) {
            while (v[i].p[ax] < piv) i++;
            while (piv < v[j].p[ax]) j--;
            if (i <= j) { swap_pt(&v[i], &v[j]); i++; j--; }
        }
        if (nth <= j) r = j;
        else if (nth >= i) l = i;
        else return;
    }
}
static int build_tree(ikdo_tree* t, long l, long r, Pt* st) {
    if (l > r) return NIL;
    int ri = node_alloc(t);
    long mid = (l + r) >> 1;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long i = l; i <= r; i++)
        for (int a = 0; a < 3; a++) { mn[a] = fmin2(mn[a], st[i].p[a]); mx[a] = fmax2(mx[a], st[i].p[a]); }
    int ax = 0;
    float rg[3];
    for (int a = 0; a < 3; a++) rg[a] = mx[a] - mn[a];
    for (int a = 1; a < 3; a++) if (rg[a] > rg[ax]) ax = a;   /* :594-595 */
    select_nth(st, l, r, mid, ax);
    int li = build_tree(t, l, mid - 1, st);
    int rri = build_tree(t, mid + 1, r, st);
    Node* n = &t->nd[ri];
    n->axis = ax;
    memcpy(n->p, st[mid].p, sizeof(n->p));
    n->left = li;
    n->right = rri;
    update(t, ri);
    return ri;
}

/* ---------------------------------------------------------------- flatten :1326-1352, delete_tree_nodes :1355-1366 */
static void flatten(ikdo_tree* t, int ri, PtVec* out, int record) {
    if (ri == NIL) return;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    if (!r->pdel) pv_push(out, r->p);
    flatten(t, r->left, out, record);
    flatten(t, t->nd[ri].right, out, record);
    r = &t->nd[ri];
    if (record && r->pdel && !r->pds) pv_push(&t->removed, r->p);   /* DELETE_POINTS_REC :1339-1341 */
}
static void delete_nodes(ikdo_tree* t, int ri) {
    if (ri == NIL) return;
    delete_nodes(t, t->nd[ri].left);
    delete_nodes(t, t->nd[ri].right);
    node_free(t, ri);
}

/* ---------------------------------------------------------------- Criterion_Check :1090-1107, Rebuild :625-645 */
static int criterion_check(const ikdo_tree* t, int ri) {
    const Node* r = &t->nd[ri];
    if (r->size <= MIN_UNBALANCED_TREE_SIZE) return 0;
    int si = r->left != NIL ? r->left : r->right;
    float de = (float)r->invalid / (float)r->size;
    float be = (float)t->nd[si].size / (float)(r->size - 1);
    if (de > t->del_param) return 1;
    if (be > t->bal_param || be < 1 - t->bal_param) return 1;
    return 0;
}
/* returns the index of the rebuilt subtree root (NIL if it became empty) */
static int rebuild(ikdo_tree* t, int ri) {
    int father = t->nd[ri].father;
    int was_root = (ri == t->root);
    t->scratch.n = 0;
    flatten(t, ri, &t->scratch, 1);
    delete_nodes(t, ri);
    if (was_root) t->root = NIL;
    int ni = build_tree(t, 0, t->scratch.n - 1, t->scratch.v);
    if (ni != NIL) t->nd[ni].father = father;
    if (was_root) { t->root = ni; if (ni != NIL) update(t, ni); }
    t->rebuilds++;
    return ni;
}

/* ---------------------------------------------------------------- Add_by_point :818-866 */
static int add_by_point(ikdo_tree* t, int ri, const float* p, int allow_rebuild, int father_axis) {
    if (ri == NIL) {
        int ni = node_alloc(t);
        Node* n = &t->nd[ni];
        memcpy(n->p, p, sizeof(n->p));
        n->axis = (father_axis + 1) % 3;
        update(t, ni);
        return ni;
    }
    push_down(t, ri);
    Node* r = &t->nd[ri];
    int ax = r->axis;
    if (p[ax] < r->p[ax]) {
        int c = add_by_point(t, r->left, p, allow_rebuild, ax);
        t->nd[ri].left = c;
    } else {
        int c = add_by_point(t, r->right, p, allow_rebuild, ax);
        t->nd[ri].right = c;
    }
    update(t, ri);
    if (allow_rebuild && criterion_check(t, ri)) return rebuild(t, ri);
    return ri;
}

/* ---------------------------------------------------------------- Delete_by_range :648-710 */
static int box_disjoint(const Node* r, const float* b) {
    for (int a = 0; a < 3; a++) if (b[3 + a] <= r->rmin[a] || b[a] > r->rmax[a]) return 1;
    return 0;
}
static int box_contains_range(const Node* r, const float* b) {
    for (int a = 0; a < 3; a++) if (!(b[a] <= r->rmin[a] && b[3 + a] > r->rmax[a])) return 0;
    return 1;
}
static int box_has_point(const float* p, const float* b) {
    for (int a = 0; a < 3; a++) if (!(b[a] <= p[a] && b[3 + a] > p[a])) return 0;
    return 1;
}
/* `ri` is the subtree root; *out receives the (possibly rebuilt) root index the parent must link to */
static int delete_by_range(ikdo_tree* t, int ri, const float* b, int allow_rebuild, int is_ds, int* out) {
    *out = ri;
    if (ri == NIL || t->nd[ri].tdel) return 0;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    int cnt = 0;
    if (box_disjoint(r, b)) return 0;
    if (box_contains_range(r, b)) {
        r->tdel = 1; r->pdel = 1; r->pushl = 1; r->pushr = 1;
        cnt = r->size - r->invalid;
        r->invalid = r->size;
        if (is_ds) { r->tds = 1; r->pds = 1; r->down_del = r->size; }
        return cnt;
    }
    if (!r->pdel && box_has_point(r->p, b)) {
        r->pdel = 1;
        cnt += 1;
        if (is_ds) r->pds = 1;
    }
    int c;
    cnt += delete_by_range(t, t->nd[ri].left, b, allow_rebuild, is_ds, &c);
    t->nd[ri].left = c;
    cnt += delete_by_range(t, t->nd[ri].right, b, allow_rebuild, is_ds, &c);
    t->nd[ri].right = c;
    update(t, ri);
    if (allow_rebuild && criterion_check(t, ri)) *out = rebuild(t, ri);
    return cnt;
}

/* ---------------------------------------------------------------- Delete_by_point :713-760 */
static int delete_by_point(ikdo_tree* t, int ri, const float* p, int allow_rebuild) {
    if (ri == NIL || t->nd[ri].tdel) return ri;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    if (same_point(r->p, p) && !r->pdel) {
        r->pdel = 1;
        r->invalid += 1;
        if (r->invalid == r->size) r->tdel = 1;
        return ri;
    }
    int ax = r->axis;
    if (p[ax] < r->p[ax]) { int c = delete_by_point(t, r->left, p, allow_rebuild); t->nd[ri].left = c; }
    else { int c = delete_by_point(t, r->right, p, allow_rebuild); t->nd[ri].right = c; }
    update(t, ri);
    if (allow_rebuild && criterion_check(t, ri)) return rebuild(t, ri);
    return ri;
}

/* ---------------------------------------------------------------- Add_by_range :763-815 */
static int add_by_range(ikdo_tree* t, int ri, const float* b, int allow_rebuild) {
    if (ri == NIL) return ri;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    if (box_disjoint(r, b)) return ri;
    if (box_contains_range(r, b)) {
        r->tdel = r->tds;
        r->pdel = r->pds;
        r->pushl = 1; r->pushr = 1;
        r->invalid = r->down_del;
        return ri;
    }
    if (box_has_point(r->p, b)) r->pdel = r->pds;
    { int c = add_by_range(t, t->nd[ri].left, b, allow_rebuild); t->nd[ri].left = c; }
    { int c = add_by_range(t, t->nd[ri].right, b, allow_rebuild); t->nd[ri].right = c; }
    update(t, ri);
    if (allow_rebuild && criterion_check(t, ri)) return rebuild(t, ri);
    return ri;
}

/* ---------------------------------------------------------------- Search_by_range :1016-1044, Search_by_radius :1047-1087 */
static void search_by_range(ikdo_tree* t, int ri, const float* b, PtVec* out) {
    if (ri == NIL) return;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    if (box_disjoint(r, b)) return;
    if (box_contains_range(r, b)) { flatten(t, ri, out, 0); return; }
    if (box_has_point(r->p, b) && !r->pdel) pv_push(out, r->p);
    search_by_range(t, r->left, b, out);
    search_by_range(t, t->nd[ri].right, b, out);
}
static void search_by_radius(ikdo_tree* t, int ri, const float* q, float radius, PtVec* out) {
    if (ri == NIL) return;
    push_down(t, ri);
    Node* r = &t->nd[ri];
    float c[3];
    for (int a = 0; a < 3; a++) c[a] = (r->rmin[a] + r->rmax[a]) * 0.5f;
    float dist = sqrtf(calc_dist(c, q));
    if (dist > radius + sqrtf(r->radius_sq)) return;
    if (dist <= radius - sqrtf(r->radius_sq)) { flatten(t, ri, out, 0); return; }
    if (!r->pdel && calc_dist(r->p, q) <= radius * radius) pv_push(out, r->p);
    search_by_radius(t, r->left, q, radius, out);
    search_by_radius(t, t->nd[ri].right, q, radius, out);
}

/* ---------------------------------------------------------------- MANUAL_HEAP ikd_Tree.h:95-172 */
typedef struct { float p[3]; float dist; } HeapEnt;
typedef struct { HeapEnt* h; int n, cap; } Heap;
static int ent_less(const HeapEnt* a, const HeapEnt* b) {   /* PointType_CMP::operator< :102-105 */
    if ((double)fabsf(a->dist - b->dist) < 1e-10) return a->p[0] < b->p[0];
    return a->dist < b->dist;
}
static void heap_push(Heap* q, const HeapEnt* e) {
    if (q->n >= q->cap) return;
    int i = q->n++;
    HeapEnt tmp = *e;
    while (i > 0) {
        int a = (i - 1) / 2;
        if (ent_less(&q->h[a], &tmp)) { q->h[i] = q->h[a]; i = a; }
        else break;
    }
    q->h[i] = tmp;
}
static void heap_pop(Heap* q) {
    if (q->n == 0) return;
    q->h[0] = q->h[q->n - 1];
    q->n--;
    int i = 0, l = 1;
    HeapEnt tmp = q->h[0];
    while (l < q->n) {
        if (l + 1 < q->n && ent_less(&q->h[l], &q->h[l + 1])) l++;
        if (ent_less(&tmp, &q->h[l])) { q->h[i] = q->h[l]; i = l; l = 2 * i + 1; }
        else break;
    }
    q->h[i] = tmp;
}

/* ---------------------------------------------------------------- Search :869-1013 (read-only: see settle_flags) */
static void search(const ikdo_tree* t, int ri, int k, const float* q, Heap* hp, double max_dist, long* visits) {
    if (ri == NIL || t->nd[ri].tdel) return;
    double cur = calc_box_dist(t, ri, q);
    double md2 = max_dist * max_dist;
    if (cur > md2) return;
    if (visits) (*visits)++;
    const Node* r = &t->nd[ri];
    if (!r->pdel) {
        float d = calc_dist(q, r->p);
        if (d <= md2 && (hp->n < k || d < hp->h[0].dist)) {
            if (hp->n >= k) heap_pop(hp);
            HeapEnt e; memcpy(e.p, r->p, sizeof(e.p)); e.dist = d;
            heap_push(hp, &e);
        }
    }
    float dl = calc_box_dist(t, r->left, q), dr = calc_box_dist(t, r->right, q);
    if (hp->n < k || (dl < hp->h[0].dist && dr < hp->h[0].dist)) {
        if (dl <= dr) {
            search(t, r->left, k, q, hp, max_dist, visits);
            if (hp->n < k || dr < hp->h[0].dist) search(t, r->right, k, q, hp, max_dist, visits);
        } else {
            search(t, r->right, k, q, hp, max_dist, visits);
            if (hp->n < k || dl < hp->h[0].dist) search(t, r->left, k, q, hp, max_dist, visits);
        }
    } else {
        if (dl < hp->h[0].dist) search(t, r->left, k, q, hp, max_dist, visits);
        if (dr < hp->h[0].dist) search(t, r->right, k, q, hp, max_dist, visits);
    }
}
/* The reference pushes pending delete flags down lazily inside Search (:875-884). Doing all pending
 * pushes up front gives the same flags on every node a search can reach and keeps search() read-only
 * (hence callable from several OpenMP threads). */
static void settle_flags(ikdo_tree* t, int ri) {
    if (ri == NIL) return;
    Node* r = &t->nd[ri];
    if (r->tdel) return;   /* never entered by Search (:870) */
    if (r->pushl || r->pushr) push_down(t, ri);
    settle_flags(t, t->nd[ri].left);
    settle_flags(t, t->nd[ri].right);
}

/* ================================================================= public API */
ikdo_tree* ikdo_create(float del, float bal, float ds) {
    ikdo_tree* t = (ikdo_tree*)calloc(1, sizeof(*t));
    t->root = NIL;
    t->del_param = del; t->bal_param = bal; t->ds = ds;
    return t;
}
void ikdo_destroy(ikdo_tree* t) {
    if (!t) return;
    free(t->nd); free(t->freelist); free(t->scratch.v); free(t->down.v); free(t->removed.v); free(t->last.v);
    free(t);
}
void ikdo_set_params(ikdo_tree* t, float del, float bal, float ds) { t->del_param = del; t->bal_param = bal; t->ds = ds; }

/* Build :353-364 */
void ikdo_build(ikdo_tree* t, const float* xyz, long n) {
    t->nused = 0; t->nfree = 0; t->root = NIL;
    if (n == 0) return;
    Pt* st = (Pt*)malloc(sizeof(Pt) * (size_t)n);
    for (long i = 0; i < n; i++) { st[i].p[0] = xyz[3 * i]; st[i].p[1] = xyz[3 * i + 1]; st[i].p[2] = xyz[3 * i + 2]; }
    /* the reference runs Update on the root before Root_Node is assigned, so the root alphas are not set by Build */
    t->root = build_tree(t, 0, n - 1, st);
    t->nd[t->root].father = NIL;
    free(st);
}

/* Nearest_Search :367-397 */
static int knn_one(const ikdo_tree* t, const float* q, int k, double max_dist, float* out_xyz, float* out_d, long* visits) {
    Heap hp;
    hp.cap = 2 * k; hp.n = 0;
    hp.h = (HeapEnt*)malloc(sizeof(HeapEnt) * (size_t)(hp.cap > 0 ? hp.cap : 1));
    search(t, t->root, k, q, &hp, max_dist, visits);
    int found = hp.n < k ? hp.n : k;
    for (int i = found - 1; i >= 0; i--) {   /* pop max, insert at the front -> ascending */
        if (out_xyz) { out_xyz[3 * i] = hp.h[0].p[0]; out_xyz[3 * i + 1] = hp.h[0].p[1]; out_xyz[3 * i + 2] = hp.h[0].p[2]; }
        if (out_d) out_d[i] = hp.h[0].dist;
        heap_pop(&hp);
    }
    free(hp.h);
    return found;
}
int ikdo_knn(ikdo_tree* t, const float* q, int k, double max_dist, float* out_xyz, float* out_d) {
    settle_flags(t, t->root);
    return knn_one(t, q, k, max_dist, out_xyz, out_d, NULL);
}
int ikdo_knn_batch(ikdo_tree* t, const float* q, long nq, int k, double max_dist, float* out_xyz, float* out_d,
                   int* out_cnt, int nthreads) {
    settle_flags(t, t->root);
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (long i = 0; i < nq; i++) {
        int c = knn_one(t, q + 3 * i, k, max_dist, out_xyz ? out_xyz + 3 * i * k : NULL, out_d ? out_d + i * k : NULL, NULL);
        if (out_cnt) out_cnt[i] = c;
    }
    return used;
}
double ikdo_mean_visits(ikdo_tree* t, const float* q, long nq, int k, double max_dist) {
    settle_flags(t, t->root);
    long total = 0;
    for (long i = 0; i < nq; i++) knn_one(t, q + 3 * i, k, max_dist, NULL, NULL, &total);
    return nq ? (double)total / (double)nq : 0.0;
}

/* ---- plane fit on the k nearest neighbours -------------------------------------------------------
 * NOT a restatement of reference code: the step lives in the reference's known caller (hku-mars/FAST_LIO
 * src/laserMapping.cpp h_share_model + include/common_lib.h esti_plane, absent from /root/reference), which hands the
 * 5 neighbours to Eigen's ColPivHouseholderQR (Eigen 3.3.x, also absent). PARITY UNPINNED against Eigen: this is the
 * published algorithm -- Householder QR with column pivoting on the largest remaining column norm, Q^T applied to
 * b = -1, back substitution -- in fp32 with a fixed operation order, which the product kernel (csrc/ikd_plane.cu)
 * must reproduce bit for bit; tests also hold it to numpy's float64 least squares within a stated tolerance. */
#define IKDO_PLANE_MAX_K 8
int ikdo_plane_fit(const float* nbr, int k, float thr, float* pl) {
    float A[IKDO_PLANE_MAX_K][3], b[IKDO_PLANE_MAX_K], ess[IKDO_PLANE_MAX_K];
    int perm[3] = {0, 1, 2};
    int ok = 1;
    if (k < 3 || k > IKDO_PLANE_MAX_K) return 0;
    for (int j = 0; j < k; j++) {
        b[j] = -1.f;
        for (int c = 0; c < 3; c++) A[j][c] = nbr[3 * j + c];
    }
    for (int s = 0; s < 3; s++) {
        int best = s;
        float bestn = -1.f;
        for (int c = s; c < 3; c++) {
            float n = 0.f;
            for (int j = s; j < k; j++) n = n + A[j][c] * A[j][c];
            if (n > bestn) { bestn = n; best = c; }
        }
        if (best != s) {
            for (int j = 0; j < k; j++) { float tmp = A[j][s]; A[j][s] = A[j][best]; A[j][best] = tmp; }
            int tp = perm[s]; perm[s] = perm[best]; perm[best] = tp;
        }
        float c0 = A[s][s], tail = 0.f;
        for (int j = s + 1; j < k; j++) tail = tail + A[j][s] * A[j][s];
        float beta = c0, tau = 0.f;
        for (int j = 0; j < k; j++) ess[j] = 0.f;
        if (tail != 0.f) {
            beta = sqrtf(c0 * c0 + tail);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            for (int j = s + 1; j < k; j++) ess[j] = A[j][s] / den;
            tau = (beta - c0) / beta;
        }
        A[s][s] = beta;
        if (!(beta != 0.f)) ok = 0;
        for (int c = s + 1; c < 3; c++) {
            float w = A[s][c];
            for (int j = s + 1; j < k; j++) w = w + ess[j] * A[j][c];
            w = w * tau;
            A[s][c] = A[s][c] - w;
            for (int j = s + 1; j < k; j++) A[j][c] = A[j][c] - ess[j] * w;
        }
        float w = b[s];
        for (int j = s + 1; j < k; j++) w = w + ess[j] * b[j];
        w = w * tau;
        b[s] = b[s] - w;
        for (int j = s + 1; j < k; j++) b[j] = b[j] - ess[j] * w;
    }
    float x[3], nv[3];
    x[2] = b[2] / A[2][2];
    x[1] = (b[1] - A[1][2] * x[2]) / A[1][1];
    x[0] = ((b[0] - A[0][1] * x[1]) - A[0][2] * x[2]) / A[0][0];
    for (int s = 0; s < 3; s++) nv[perm[s]] = x[s];
    float nn = sqrtf((nv[0] * nv[0] + nv[1] * nv[1]) + nv[2] * nv[2]);
    pl[0] = nv[0] / nn;
    pl[1] = nv[1] / nn;
    pl[2] = nv[2] / nn;
    pl[3] = 1.f / nn;
    for (int c = 0; c < 4; c++)
        if (!(fabsf(pl[c]) <= 3.0e38f)) ok = 0;
    if (ok)
        for (int j = 0; j < k; j++) {
            float r = ((pl[0] * nbr[3 * j] + pl[1] * nbr[3 * j + 1]) + pl[2] * nbr[3 * j + 2]) + pl[3];
            if (fabsf(r) > thr) ok = 0;
        }
    return ok;
}

/* Gate (k found, k-th squared distance <= max_kth_sqdist), fit and residual for nq queries whose neighbours are given
 * (nbr: nq*k*3, sqd: nq*k, cnt: nq). Rows that are gated out or whose fit is not finite hold zeros. */
void ikdo_plane_batch(const float* q, long nq, int k, const float* nbr, const float* sqd, const int* cnt,
                      float max_kth_sqdist, float thr, float* out_plane, float* out_resid, unsigned char* out_valid) {
    for (long i = 0; i < nq; i++) {
        float pl[4] = {0.f, 0.f, 0.f, 0.f}, resid = 0.f;
        unsigned char valid = 0;
        if (cnt[i] == k && sqd[i * k + (k - 1)] <= max_kth_sqdist) {
            int ok = ikdo_plane_fit(nbr + 3 * i * k, k, thr, pl);
            resid = ((pl[0] * q[3 * i] + pl[1] * q[3 * i + 1]) + pl[2] * q[3 * i + 2]) + pl[3];
            if (!(fabsf(resid) <= 3.0e38f)) { ok = 0; resid = 0.f; pl[0] = pl[1] = pl[2] = pl[3] = 0.f; }
            valid = (unsigned char)ok;
        }
        memcpy(out_plane + 4 * i, pl, sizeof pl);
        out_resid[i] = resid;
        out_valid[i] = valid;
    }
}

/* Box_Search :400-404, Radius_Search :407-411 */
long ikdo_box_search(ikdo_tree* t, const float* box6, float* out_xyz, long cap) {
    t->last.n = 0;
    search_by_range(t, t->root, box6, &t->last);
    return pv_copy(&t->last, out_xyz, cap);
}
long ikdo_radius_search(ikdo_tree* t, const float* c, float r, float* out_xyz, long cap) {
    t->last.n = 0;
    search_by_radius(t, t->root, c, r, &t->last);
    return pv_copy(&t->last, out_xyz, cap);
}
long ikdo_last_result(ikdo_tree* t, float* out_xyz, long cap) { return pv_copy(&t->last, out_xyz, cap); }

/* Add_Points :414-489 */
int ikdo_add_points(ikdo_tree* t, const float* xyz, long n, int downsample_on) {
    int counter = 0;
    float ds = t->ds;
    for (long i = 0; i < n; i++) {
        const float* p = xyz + 3 * i;
        if (downsample_on) {
            float b[6], mid[3];
            for (int a = 0; a < 3; a++) {
                b[a] = floorf(p[a] / ds) * ds;                                  /* :424-429 */
                b[3 + a] = b[a] + ds;
                mid[a] = (float)((double)b[a] + (double)(b[3 + a] - b[a]) / 2.0);   /* :430-432 */
            }
            t->down.n = 0;
            search_by_range(t, t->root, b, &t->down);
            float min_dist = calc_dist(p, mid);
            float res[3] = {p[0], p[1], p[2]};
            for (long j = 0; j < t->down.n; j++) {
                float d = calc_dist(t->down.v[j].p, mid);
                if (d < min_dist) { min_dist = d; memcpy(res, t->down.v[j].p, sizeof(res)); }
            }
            if (t->down.n > 1 || same_point(p, res)) {                          /* :445 */
                if (t->down.n > 0) { int nr; delete_by_range(t, t->root, b, 1, 1, &nr); t->root = nr; }
                if (t->root == NIL) {   /* the reference would dereference null here; keep going sensibly */
                    t->root = add_by_point(t, NIL, res, 1, 2);
                } else {
                    t->root = add_by_point(t, t->root, res, 1, t->nd[t->root].axis);
                }
                t->nd[t->root].father = NIL;
                counter++;
            }
        } else {
            if (t->root == NIL) t->root = add_by_point(t, NIL, p, 1, 2);
            else t->root = add_by_point(t, t->root, p, 1, t->nd[t->root].axis);
            t->nd[t->root].father = NIL;
        }
    }
    return counter;
}

/* Delete_Points :514-533 */
void ikdo_delete_points(ikdo_tree* t, const float* xyz, long n) {
    for (long i = 0; i < n; i++) t->root = delete_by_point(t, t->root, xyz + 3 * i, 1);
}
/* Delete_Point_Boxes :536-556 */
int ikdo_delete_boxes(ikdo_tree* t, const float* boxes, long nb) {
    int c = 0;
    for (long i = 0; i < nb; i++) { int nr; c += delete_by_range(t, t->root, boxes + 6 * i, 1, 0, &nr); t->root = nr; }
    return c;
}
/* Add_Point_Boxes :492-511 */
void ikdo_add_boxes(ikdo_tree* t, const float* boxes, long nb) {
    for (long i = 0; i < nb; i++) t->root = add_by_range(t, t->root, boxes + 6 * i, 1);
}

int ikdo_size(ikdo_tree* t) { return t->root == NIL ? 0 : t->nd[t->root].size; }                       /* :79 */
int ikdo_validnum(ikdo_tree* t) { return t->root == NIL ? 0 : t->nd[t->root].size - t->nd[t->root].invalid; }  /* :129 */
void ikdo_root_alpha(ikdo_tree* t, float* bal, float* del) {                                            /* :148 */
    if (t->root == NIL) { *bal = 0; *del = 0; return; }
    *bal = t->nd[t->root].alpha_bal; *del = t->nd[t->root].alpha_del;
}
void ikdo_tree_range(ikdo_tree* t, float* box6) {                                                       /* :99 */
    memset(box6, 0, sizeof(float) * 6);
    if (t->root == NIL) return;
    for (int a = 0; a < 3; a++) { box6[a] = t->nd[t->root].rmin[a]; box6[3 + a] = t->nd[t->root].rmax[a]; }
}
long ikdo_flatten(ikdo_tree* t, float* out_xyz, long cap) {
    t->last.n = 0;
    flatten(t, t->root, &t->last, 0);
    return pv_copy(&t->last, out_xyz, cap);
}
long ikdo_acquire_removed(ikdo_tree* t, float* out_xyz, long cap) {                                     /* :559-571 */
    long n = pv_copy(&t->removed, out_xyz, cap);
    t->removed.n = 0;
    return n;
}

static void dump_rec(const ikdo_tree* t, int ri, float* out, long cap, long* k) {
    if (ri == NIL) return;
    const Node* n = &t->nd[ri];
    if (*k < cap) {
        float* o = out + 16 * (*k);
        o[0] = n->p[0]; o[1] = n->p[1]; o[2] = n->p[2]; o[3] = (float)n->axis;
        o[4] = (float)n->size; o[5] = (float)n->invalid;
        o[6] = (float)((n->pdel ? 1 : 0) | (n->tdel ? 2 : 0) | (n->pds ? 4 : 0) | (n->tds ? 8 : 0) | (n->pushl ? 16 : 0) | (n->pushr ? 32 : 0));
        o[7] = n->rmin[0]; o[8] = n->rmax[0]; o[9] = n->rmin[1]; o[10] = n->rmax[1]; o[11] = n->rmin[2]; o[12] = n->rmax[2];
        o[13] = n->left != NIL ? 1.f : 0.f; o[14] = n->right != NIL ? 1.f : 0.f; o[15] = (float)n->down_del;
    }
    (*k)++;
    dump_rec(t, n->left, out, cap, k);
    dump_rec(t, n->right, out, cap, k);
}
long ikdo_dump_tree(ikdo_tree* t, float* out, long cap) {
    long k = 0;
    dump_rec(t, t->root, out, cap, &k);
    return k;
}
static int depth_rec(const ikdo_tree* t, int ri) {
    if (ri == NIL) return 0;
    int a = depth_rec(t, t->nd[ri].left), b = depth_rec(t, t->nd[ri].right);
    return 1 + (a > b ? a : b);
}
int ikdo_max_depth(ikdo_tree* t) { return depth_rec(t, t->root); }
int ikdo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int ikdo_rebuild_count(ikdo_tree* t) { return t->rebuilds; }
