// ORACLE (test infrastructure, not product code).
// Exact k-nearest-neighbour search standing in for pcl::KdTreeFLANN<PointXYZI> ->
// FLANN 1.9.1 KDTreeSingleIndex<L2_Simple<float>> (call sites ltremovert/src/Session.cpp:404, 457,
// 471, 489, 592, 627).  FLANN/PCL are not under /root/reference (UNPINNED): restated
// from their published behaviour:
//   * xyz only (PointXYZI's default representation is 3-D), exact search (eps = 0);
//   * L2_Simple<float>: d2 = ((dx*dx) + dy*dy) + dz*dz accumulated in f32 in x,y,z order, where
//     dx = query.x - point.x; SQUARED distances are returned, ascending;
//   * k is clamped to the number of indexed points.
// The decisions downstream depend only on the multiset of the k smallest d2, so any exact search
// structure reproduces them.  Pruning uses a lower bound evaluated with the same f32 formula on
// the query-to-box offsets, which is conservative because every rounding step is monotone.
#pragma once
#include <vector>
#include <algorithm>
#include <cstdint>
#include <cmath>
#include <limits>

namespace ltr_oracle {

struct KdTree {
    struct Node {
        int left, right;      // point range [left,right) for leaves
        int child0, child1;   // -1 for leaves
        float lo[3], hi[3];   // bounding box of the points below
    };
    std::vector<float> pts;   // reordered xyz
    std::vector<Node> nodes;
    int n = 0;
    static constexpr int kLeaf = 15;

    static inline float dist2(const float* a, const float* b) {
        float r = 0.0f, d;
        d = a[0] - b[0]; r += d * d;
        d = a[1] - b[1]; r += d * d;
        d = a[2] - b[2]; r += d * d;
        return r;
    }

    void build(const float* xyz, int stride, int count) {
        n = count;
        pts.resize((size_t)n * 3);
        for (int i = 0; i < n; ++i) {
            pts[3 * (size_t)i + 0] = xyz[(size_t)i * stride + 0];
            pts[3 * (size_t)i + 1] = xyz[(size_t)i * stride + 1];
            pts[3 * (size_t)i + 2] = xyz[(size_t)i * stride + 2];
        }
        nodes.clear();
        if (n > 0) {
            nodes.reserve((size_t)(2 * n / kLeaf + 4));
            buildRec(0, n);
        }
    }

    int buildRec(int l, int r) {
        const int id = (int)nodes.size();
        nodes.push_back(Node());
        Node nd;
        nd.left = l; nd.right = r; nd.child0 = nd.child1 = -1;
        for (int d = 0; d < 3; ++d) { nd.lo[d] = std::numeric_limits<float>::infinity(); nd.hi[d] = -nd.lo[d]; }
        for (int i = l; i < r; ++i)
            for (int d = 0; d < 3; ++d) {
                const float v = pts[3 * (size_t)i + d];
                nd.lo[d] = std::min(nd.lo[d], v);
                nd.hi[d] = std::max(nd.hi[d], v);
            }
        if (r - l > kLeaf) {
            int dim = 0;
            float ext = nd.hi[0] - nd.lo[0];
            for (int d = 1; d < 3; ++d) if (nd.hi[d] - nd.lo[d] > ext) { ext = nd.hi[d] - nd.lo[d]; dim = d; }
            if (ext > 0.0f) {
                const int mid = (l + r) / 2;
                // nth_element on triples
                struct P3 { float v[3]; };
                P3* base = reinterpret_cast<P3*>(pts.data());
                std::nth_element(base + l, base + mid, base + r, [dim](const P3& a, const P3& b) { return a.v[dim] < b.v[dim]; });
                const int c0 = buildRec(l, mid);
                const int c1 = buildRec(mid, r);
                nd.child0 = c0; nd.child1 = c1;
            }
        }
        nodes[id] = nd;
        return id;
    }

    // conservative lower bound of dist2(q, p) for any p in the node box (same f32 formula)
    static inline float boxLB(const float* q, const Node& nd) {
        float c[3];
        for (int d = 0; d < 3; ++d) {
            if (q[d] < nd.lo[d]) c[d] = q[d] - nd.lo[d];
            else if (q[d] > nd.hi[d]) c[d] = q[d] - nd.hi[d];
            else c[d] = 0.0f;
        }
        float r = 0.0f;
        r += c[0] * c[0]; r += c[1] * c[1]; r += c[2] * c[2];
        return r;
    }

    // Fills out[0..kk) with the kk = min(k, n) smallest squared distances, ascending. Returns kk.
    int knn(const float* q, int k, float* out) const {
        const int kk = std::min(k, n);
        if (kk <= 0) return 0;
        for (int i = 0; i < kk; ++i) out[i] = std::numeric_limits<float>::infinity();
        search(0, q, kk, out);
        return kk;
    }

    void search(int id, const float* q, int kk, float* best) const {
        const Node& nd = nodes[id];
        if (nd.child0 < 0) {
            for (int i = nd.left; i < nd.right; ++i) {
                const float d = dist2(q, &pts[3 * (size_t)i]);
                if (d < best[kk - 1]) {
                    int j = kk - 1;
                    while (j > 0 && best[j - 1] > d) { best[j] = best[j - 1]; --j; }
                    best[j] = d;
                }
            }
            return;
        }
        const float lb0 = boxLB(q, nodes[nd.child0]);
        const float lb1 = boxLB(q, nodes[nd.child1]);
        const int first = lb0 <= lb1 ? nd.child0 : nd.child1;
        const int second = lb0 <= lb1 ? nd.child1 : nd.child0;
        const float lbf = std::min(lb0, lb1), lbs = std::max(lb0, lb1);
        if (lbf < best[kk - 1]) search(first, q, kk, best);
        if (lbs < best[kk - 1]) search(second, q, kk, best);
    }
};

}  // namespace ltr_oracle
