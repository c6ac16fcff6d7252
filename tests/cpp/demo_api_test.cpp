// Demo-style test of the drop-in C++ API (include/ikd_Tree.h), modelled on the reference's
// examples/ikd_Tree_demo.cpp workload (:16-30, :184-286): build a random cloud, then rounds of
// Add_Points / Delete_Points / Delete_Point_Boxes / Nearest_Search -- but unlike the reference demo,
// every result is asserted against a brute-force fp32 scan of the bookkeeping cloud.
// Exit code 0 and a final "PASS" line on success.
#include <ikd_Tree.h>

#include <cstdio>
#include <cstdlib>
#include <random>

struct PointXYZI16 {  // a payload-carrying point type (like pcl::PointXYZI): x,y,z first, extra fields opaque
    float x = 0, y = 0, z = 0, pad = 1.0f;
    float intensity = 0;
    float pad2[3] = {0, 0, 0};
};

template <class P> static float sqd(const P& a, const P& b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

static int failures = 0;
#define EXPECT(cond, ...)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            failures++;                                     \
            printf("FAIL %s:%d: ", __FILE__, __LINE__);    \
            printf(__VA_ARGS__);                            \
            printf("\n");                                   \
        }                                                   \
    } while (0)

template <class P> static P make_point(std::mt19937& g, float lo, float hi, int tag) {
    std::uniform_real_distribution<float> u(lo, hi);
    P p;
    p.x = u(g); p.y = u(g); p.z = u(g);
    (void)tag;
    return p;
}
template <> PointXYZI16 make_point<PointXYZI16>(std::mt19937& g, float lo, float hi, int tag) {
    std::uniform_real_distribution<float> u(lo, hi);
    PointXYZI16 p;
    p.x = u(g); p.y = u(g); p.z = u(g);
    p.intensity = (float)tag;
    return p;
}
static float payload(const ikdTree_PointType&) { return 0.f; }
static float payload(const PointXYZI16& p) { return p.intensity; }

template <class P> int run(const char* name) {
    using Tree = KD_TREE<P>;
    using PV = typename Tree::PointVector;
    std::mt19937 g(1);
    Tree tree(0.3, 0.6, 0.2);
    EXPECT(tree.Root_Node == nullptr && tree.size() == 0, "fresh tree not empty");
    PV cloud;  // bookkeeping copy of the valid points
    for (int i = 0; i < 20000; i++) cloud.push_back(make_point<P>(g, -5.f, 5.f, i));
    tree.Build(cloud);
    EXPECT(tree.Root_Node != nullptr, "Root_Node null after Build");
    EXPECT(tree.size() == 20000 && tree.validnum() == 20000, "size %d validnum %d", tree.size(), tree.validnum());
    int tag = 20000;
    for (int round = 0; round < 40; round++) {
        // + 200 points
        PV inc;
        for (int i = 0; i < 200; i++) { inc.push_back(make_point<P>(g, -5.f, 5.f, tag++)); cloud.push_back(inc.back()); }
        int r = tree.Add_Points(inc, false);
        EXPECT(r == 0, "Add_Points(no downsample) returned %d, reference returns 0", r);
        // - 100 points
        PV dec;
        for (int i = 0; i < 100; i++) {
            std::uniform_int_distribution<size_t> u(0, cloud.size() - 1);
            size_t j = u(g);
            dec.push_back(cloud[j]);
            cloud[j] = cloud.back();
            cloud.pop_back();
        }
        tree.Delete_Points(dec);
        // box delete every 10 rounds
        if (round % 10 == 5) {
            std::vector<BoxPointType> boxes;
            for (int b = 0; b < 4; b++) {
                P c = make_point<P>(g, -5.f, 5.f, 0);
                BoxPointType bx;
                float cc[3] = {c.x, c.y, c.z};
                for (int a = 0; a < 3; a++) { bx.vertex_min[a] = cc[a] - 0.75f; bx.vertex_max[a] = cc[a] + 0.75f; }
                boxes.push_back(bx);
            }
            int expect = 0;
            PV keep;
            for (auto& p : cloud) {
                bool in = false;
                float pc[3] = {p.x, p.y, p.z};
                for (auto& bx : boxes) {
                    bool i2 = true;
                    for (int a = 0; a < 3; a++) i2 = i2 && bx.vertex_min[a] <= pc[a] && bx.vertex_max[a] > pc[a];
                    in = in || i2;
                }
                if (in) expect++; else keep.push_back(p);
            }
            cloud.swap(keep);
            int got = tree.Delete_Point_Boxes(boxes);
            EXPECT(got == expect, "Delete_Point_Boxes %d expected %d", got, expect);
        }
        EXPECT(tree.validnum() == (int)cloud.size(), "round %d validnum %d expected %zu", round, tree.validnum(), cloud.size());
        // 200 x 5-NN, batched and single, against brute force
        PV queries;
        for (int i = 0; i < 200; i++) queries.push_back(make_point<P>(g, -5.f, 5.f, 0));
        std::vector<PV> npts;
        std::vector<std::vector<float>> nd;
        tree.Nearest_Search(queries, 5, npts, nd);
        for (int i = 0; i < 200; i++) {
            std::vector<float> all;
            all.reserve(cloud.size());
            for (auto& p : cloud) all.push_back(sqd(p, queries[i]));
            std::partial_sort(all.begin(), all.begin() + 5, all.end());
            EXPECT(nd[i].size() == 5, "query %d returned %zu", i, nd[i].size());
            for (size_t j = 0; j < nd[i].size() && j < 5; j++) {
                EXPECT(nd[i][j] == all[j], "round %d query %d nn %zu dist %.9g expected %.9g", round, i, j, nd[i][j], all[j]);
                EXPECT(sqd(npts[i][j], queries[i]) == nd[i][j], "returned point does not match its distance");
            }
            if (i < 3) {
                PV sp;
                std::vector<float> sd;
                tree.Nearest_Search(queries[i], 5, sp, sd);
                EXPECT(sd == nd[i], "single-query result differs from batched");
            }
        }
    }
    // kNN + plane fit on the device: for rows flagged valid the plane is a unit normal whose distance to each of the 5
    // neighbours is within the threshold, and the residual is the query's signed distance to it
    {
        PV queries;
        for (int i = 0; i < 300; i++) queries.push_back(make_point<P>(g, -5.f, 5.f, 0));
        std::vector<PV> npts;
        std::vector<std::vector<float>> nd;
        tree.Nearest_Search(queries, 5, npts, nd);
        std::vector<float> plane, resid;
        std::vector<uint8_t> valid;
        tree.Nearest_Plane_Batch(queries, 5, plane, resid, valid, INFINITY, 5.0f, 0.1f);
        EXPECT(plane.size() == 1200 && resid.size() == 300 && valid.size() == 300, "Nearest_Plane_Batch sizes");
        int nvalid = 0;
        for (int i = 0; i < 300; i++) {
            const float* pl = &plane[4 * (size_t)i];
            if (!valid[i]) continue;
            nvalid++;
            double nn = sqrt((double)pl[0] * pl[0] + (double)pl[1] * pl[1] + (double)pl[2] * pl[2]);
            EXPECT(fabs(nn - 1.0) < 1e-5, "plane normal of query %d has norm %.9g", i, nn);
            for (auto& p : npts[i]) {
                double r = (double)pl[0] * p.x + (double)pl[1] * p.y + (double)pl[2] * p.z + pl[3];
                EXPECT(fabs(r) <= 0.1 + 1e-5, "neighbour of query %d is %.6g from its plane", i, r);
            }
            double rq = (double)pl[0] * queries[i].x + (double)pl[1] * queries[i].y + (double)pl[2] * queries[i].z + pl[3];
            EXPECT(fabs(rq - resid[i]) < 1e-5, "residual of query %d: %.9g vs %.9g", i, (double)resid[i], rq);
        }
        EXPECT(nvalid > 0, "no valid plane among 300 queries");
    }
    // payload round-trip: every returned point must be bit-identical to one we inserted (tag preserved)
    {
        PV q(1, cloud[123]);
        PV sp;
        std::vector<float> sd;
        tree.Nearest_Search(q[0], 1, sp, sd);
        EXPECT(sp.size() == 1 && sd[0] == 0.f && payload(sp[0]) == payload(cloud[123]), "payload did not round-trip");
    }
    // box + radius search vs brute force
    {
        BoxPointType bx;
        for (int a = 0; a < 3; a++) { bx.vertex_min[a] = -1.f; bx.vertex_max[a] = 1.5f; }
        PV res;
        tree.Box_Search(bx, res);
        size_t expect = 0;
        for (auto& p : cloud) expect += (p.x >= -1.f && p.x < 1.5f && p.y >= -1.f && p.y < 1.5f && p.z >= -1.f && p.z < 1.5f);
        EXPECT(res.size() == expect, "Box_Search %zu expected %zu", res.size(), expect);
        P c; c.x = 0.5f; c.y = -0.5f; c.z = 0.25f;
        tree.Radius_Search(c, 1.25f, res);
        size_t lo = 0, hi = 0;  // the reference's whole-subtree shortcut has ULP slack
        for (auto& p : cloud) { float d = sqd(p, c); lo += d <= 1.25f * 1.25f * 0.9999f; hi += d <= 1.25f * 1.25f * 1.0001f; }
        EXPECT(res.size() >= lo && res.size() <= hi, "Radius_Search %zu not in [%zu,%zu]", res.size(), lo, hi);
    }
    // downsampled insert: at most one point per voxel afterwards among touched voxels
    {
        PV inc;
        for (int i = 0; i < 3000; i++) inc.push_back(make_point<P>(g, -5.f, 5.f, tag++));
        int before = tree.validnum();
        int added = tree.Add_Points(inc, true);
        EXPECT(added > 0 && tree.validnum() <= before + added, "downsample add: added %d validnum %d before %d", added, tree.validnum(), before);
        PV all;
        tree.flatten(tree.Root_Node, all, NOT_RECORD);
        EXPECT((int)all.size() == tree.validnum(), "flatten %zu validnum %d", all.size(), tree.validnum());
    }
    float ab, ad;
    tree.root_alpha(ab, ad);
    BoxPointType rg = tree.tree_range();
    EXPECT(rg.vertex_min[0] >= -5.f && rg.vertex_max[0] <= 5.f, "tree_range");
    printf("[%s] size %d validnum %d alpha_bal %.3f alpha_del %.3f failures %d\n", name, tree.size(), tree.validnum(), ab, ad, failures);
    return failures;
}

int main() {
    int f = 0;
    try {
        f += run<ikdTree_PointType>("ikdTree_PointType");
        f += run<PointXYZI16>("PointXYZI16");
    } catch (const std::exception& e) {
        printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    printf(f == 0 ? "PASS\n" : "FAILED\n");
    return f == 0 ? 0 : 1;
}
