// Drop-in behaviour of include/ikd_Tree.h that the reference's own callers rely on:
//   1. any number of threads may call the single-query Nearest_Search at once (reference ikd_Tree.cpp:371-387,
//      :875-884; FAST-LIO2 does it inside `#pragma omp parallel for`): results must equal the batched call bit for bit,
//      whatever mix of k / max_dist the threads use, and the calls must be combined (timing printed);
//   2. a long-running map must not grow without bound: ids / payload are compacted (ikd_compact_ids) while every
//      result keeps returning the caller's full PointType (payload tag), including acquire_removed_points.
// Exit code 0 and a final "PASS" line on success. Build with -fopenmp.
#include <ikd_Tree.h>
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>

struct PointXYZI16 {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    float intensity = 0;
    float pad2[3] = {0, 0, 0};
};

static int failures = 0;
#define EXPECT(cond, ...)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            _Pragma("omp critical")                         \
            {                                               \
                failures++;                                 \
                if (failures < 30) {                        \
                    printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                    printf(__VA_ARGS__);                    \
                    printf("\n");                           \
                }                                           \
            }                                               \
        }                                                   \
    } while (0)

using Tree = KD_TREE<PointXYZI16>;
using PV = Tree::PointVector;

static PointXYZI16 rnd(std::mt19937& g, float lo, float hi, int tag) {
    std::uniform_real_distribution<float> u(lo, hi);
    PointXYZI16 p;
    p.x = u(g); p.y = u(g); p.z = u(g);
    p.intensity = (float)tag;
    return p;
}
static float sqd(const PointXYZI16& a, const PointXYZI16& b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

static void concurrent_queries(int nthreads) {
    std::mt19937 g(7);
    Tree tree(0.5, 0.6, 0.2);
    PV cloud;
    const int N = 200000, NQ = 20000;
    for (int i = 0; i < N; i++) cloud.push_back(rnd(g, -20.f, 20.f, i));
    tree.Build(cloud);
    PV queries;
    for (int i = 0; i < NQ; i++) queries.push_back(rnd(g, -20.f, 20.f, 0));
    std::vector<PV> bp;
    std::vector<std::vector<float>> bd;
    tree.Nearest_Search(queries, 5, bp, bd, 3.0);
    // warm-up round, then the timed loop: the unmodified caller's pattern
    for (int rep = 0; rep < 2; rep++) {
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int i = 0; i < NQ; i++) {
            PV sp;
            std::vector<float> sd;
            tree.Nearest_Search(queries[i], 5, sp, sd, 3.0);
            EXPECT(sd == bd[i], "thread result of query %d differs from the batched result", i);
            for (size_t j = 0; j < sp.size(); j++)
                EXPECT(sqd(sp[j], queries[i]) == sd[j] && sp[j].intensity == bp[i][j].intensity, "payload / distance mismatch, query %d", i);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rep == 1)
            printf("[concurrent] %d single-query Nearest_Search calls from %d threads: %.2f ms (%.1f us per call, %.2f M q/s)\n", NQ,
                   nthreads, ms, 1e3 * ms / NQ, NQ / ms * 1e-3);
    }
    // one caller thread, no team to wait for
    {
        auto t0 = std::chrono::steady_clock::now();
        const int M = 2000;
        for (int i = 0; i < M; i++) {
            PV sp;
            std::vector<float> sd;
            tree.Nearest_Search(queries[i], 5, sp, sd, 3.0);
            EXPECT(sd == bd[i], "serial single-query result %d differs", i);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        printf("[concurrent] %d single-query calls from 1 thread: %.2f ms (%.1f us per call)\n", M, ms, 1e3 * ms / M);
    }
    // mixed k and max_dist across the team (k = 20 takes the heap kernel, k <= 0 returns nothing)
    const int ks[5] = {1, 5, 8, 20, 0};
    const double mds[3] = {INFINITY, 1.5, 0.0};
    std::vector<std::vector<float>> expect[15];
    for (int a = 0; a < 5; a++)
        for (int b = 0; b < 3; b++) {
            std::vector<PV> p;
            std::vector<std::vector<float>> d;
            PV q(queries.begin(), queries.begin() + 600);
            tree.Nearest_Search(q, ks[a], p, d, mds[b]);
            expect[a * 3 + b] = d;
        }
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int i = 0; i < 600 * 15; i++) {
        int qi = i % 600, a = (i / 600) % 5, b = (i / 3000) % 3;
        PV sp;
        std::vector<float> sd;
        tree.Nearest_Search(queries[qi], ks[a], sp, sd, mds[b]);
        EXPECT(sd == expect[a * 3 + b][qi], "mixed team: query %d k %d max_dist %g differs", qi, ks[a], mds[b]);
    }
}

static void long_running_map() {
    std::mt19937 g(11);
    Tree tree(0.5, 0.6, 0.4);
    tree.id_compaction_slack = 20000;
    std::map<int, PointXYZI16> live;  // bookkeeping by tag (brute force)
    std::set<int> box_deleted, reported_removed;
    PV cloud;
    int tag = 0;
    for (int i = 0; i < 30000; i++) { cloud.push_back(rnd(g, -10.f, 10.f, tag)); live[tag] = cloud.back(); tag++; }
    tree.Build(cloud);
    size_t max_payload = 0;
    int compactions = 0;
    size_t last_payload = tree.payload_size();
    for (int round = 0; round < 60; round++) {
        // plain inserts (every point becomes a node) ...
        PV inc;
        for (int i = 0; i < 3000; i++) { inc.push_back(rnd(g, -10.f, 10.f, tag)); live[tag] = inc.back(); tag++; }
        tree.Add_Points(inc, false);
        // ... and a moving slab delete that keeps the live count bounded
        std::vector<BoxPointType> boxes(1);
        float lo = -10.f + (float)(round % 10) * 2.f;
        boxes[0].vertex_min[0] = lo; boxes[0].vertex_max[0] = lo + 2.f;
        for (int a = 1; a < 3; a++) { boxes[0].vertex_min[a] = -100.f; boxes[0].vertex_max[a] = 100.f; }
        int expect = 0;
        for (auto it = live.begin(); it != live.end();) {
            if (it->second.x >= lo && it->second.x < lo + 2.f) { box_deleted.insert(it->first); it = live.erase(it); expect++; }
            else ++it;
        }
        int got = tree.Delete_Point_Boxes(boxes);
        EXPECT(got == expect, "round %d: Delete_Point_Boxes %d expected %d", round, got, expect);
        EXPECT(tree.validnum() == (int)live.size(), "round %d: validnum %d expected %zu", round, tree.validnum(), live.size());
        if (tree.payload_size() < last_payload) compactions++;
        last_payload = tree.payload_size();
        max_payload = std::max(max_payload, tree.payload_size());
        // removed points keep their payload through compactions; each is reported once and was deleted by a box
        PV rem;
        tree.acquire_removed_points(rem);
        for (auto& p : rem) {
            int t = (int)p.intensity;
            EXPECT(box_deleted.count(t) == 1, "removed point with tag %d was never box-deleted", t);
            EXPECT(reported_removed.insert(t).second, "removed point with tag %d reported twice", t);
        }
        // results still carry the right payload
        if (round % 10 == 9) {
            PV q;
            std::vector<int> tags;
            int c = 0;
            for (auto& kv : live) { if ((c++ % 97) == 0) { q.push_back(kv.second); tags.push_back(kv.first); } }
            std::vector<PV> p;
            std::vector<std::vector<float>> d;
            tree.Nearest_Search(q, 1, p, d);
            for (size_t i = 0; i < q.size(); i++)
                EXPECT(p[i].size() == 1 && d[i][0] == 0.f && live.count((int)p[i][0].intensity) == 1 && sqd(p[i][0], q[i]) == 0.f,
                       "round %d: self query of tag %d returned tag %d dist %g", round, tags[i], p[i].empty() ? -1 : (int)p[i][0].intensity,
                       d[i].empty() ? -1.f : d[i][0]);
            PV all;
            tree.flatten(tree.Root_Node, all, NOT_RECORD);
            EXPECT(all.size() == live.size(), "flatten %zu expected %zu", all.size(), live.size());
            for (auto& pt : all) EXPECT(live.count((int)pt.intensity) == 1, "flatten returned dead tag %d", (int)pt.intensity);
        }
    }
    // 210k ids were handed out in total; without compaction the payload array would hold all of them
    printf("[long-running] ids handed out %d, live %zu, payload array now %zu (max %zu), compactions %d, removed reported %zu of %zu deleted\n",
           tag, live.size(), tree.payload_size(), max_payload, compactions, reported_removed.size(), box_deleted.size());
    EXPECT(compactions >= 2, "no id compaction happened");
    EXPECT(max_payload <= 2 * 60000 + 20000 + 3000 + 30000, "payload array grew to %zu", max_payload);
}

int main(int argc, char** argv) {
    int nthreads = argc > 1 ? atoi(argv[1]) : 16;
    try {
        concurrent_queries(nthreads);
        long_running_map();
    } catch (const std::exception& e) {
        printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    printf(failures == 0 ? "PASS\n" : "FAILED (%d)\n", failures);
    return failures == 0 ? 0 : 1;
}
